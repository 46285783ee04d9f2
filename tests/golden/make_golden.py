#!/usr/bin/env python
"""Generates tests/golden/chain_stats.json from the reference's recorded chains and writes
tests/golden/known_answers.json (values printed in the reference's notebooks).

Run in the authoring container only (needs /root/reference):
    python tests/golden/make_golden.py
The reference itself cannot be imported here (qulacs/matplotlib/seaborn missing), so its *recorded
outputs* are the fixtures: pickled MCMCChain lists and notebook cell outputs.
"""
import glob
import json
import os
import pickle
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import qumcmc  # noqa: E402,F401  (registers qumcmc.basic_utils for the unpickler)

REF = "/root/reference"


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("numpy.core"):
            module = module.replace("numpy.core", "numpy._core")
        return super().find_class(module, name)


def load(path):
    with open(path, "rb") as f:
        return _Unpickler(f).load()


def chain_stats(chain):
    idx = chain.proposal_indices
    acc = chain.accepted_mask
    n = chain.num_spins
    mc = chain.markov_chain_indices()
    cur = mc[:-1]
    prop = idx[1:]
    ham = np.array([bin(int(a) ^ int(b)).count("1") for a, b in zip(cur, prop)])
    hist = np.bincount(ham, minlength=n + 1) / len(ham)
    return {"n_hops": int(len(prop)), "num_spins": int(n), "initial": int(idx[0]),
            "acceptance": float(acc[1:].mean()), "hamming": hist.tolist(),
            "visit_top": [[int(k), int(c)] for k, c in
                          sorted(zip(*np.unique(mc, return_counts=True)), key=lambda kc: -kc[1])[:16]]}


def main():
    out = {"bas": {}, "gamma_variation": {}}
    bas_dir = os.path.join(REF, "final_results_notebooks", "results_with_new_code")
    for path in sorted(glob.glob(os.path.join(bas_dir, "BAS_new_code_final_plots_wt1*.pkl"))):
        chains = load(path)
        stats = [chain_stats(c) for c in chains]
        out["bas"][os.path.basename(path)] = {
            "source": os.path.relpath(path, REF), "n_chains": len(stats),
            "acceptance_mean": float(np.mean([s["acceptance"] for s in stats])),
            "acceptance_std": float(np.std([s["acceptance"] for s in stats])),
            "hamming_mean": np.mean([s["hamming"] for s in stats], axis=0).tolist(),
            "chains": stats}
    for n in (8, 9):
        for seed in (1, 2, 3):
            path = os.path.join(REF, "gamma_variation", f"num_spins_{n}_Gamma_var_num_spins_{n}_beta_1.5_seed_{seed}.pkl")
            if not os.path.exists(path):
                continue
            d = load(path)
            out["gamma_variation"][f"n{n}_seed{seed}"] = {
                "source": os.path.relpath(path, REF), "num_spins": n, "seed": seed, "beta": 1.5,
                "ranges": {k: chain_stats(v) for k, v in d.items()}}
    with open(os.path.join(ROOT, "tests", "golden", "chain_stats.json"), "w") as f:
        json.dump(out, f, indent=1)

    known = {
        "_comment": "Values printed in the reference's notebooks (SURVEY.md section 4); sources relative to the reference checkout.",
        "bas3_bars_stripes": {"source": "tutorial.ipynb cells 10,15; final_results_notebooks/review_code_bas_new_code.ipynb cells 4,12",
                              "alpha": 0.125, "state": "111000111", "energy": -48.0, "n_states_gt_0.01_beta1": 12,
                              "entropy_beta_1.0": 3.5849606926186035, "entropy_beta_1.5": 3.5849625001146053},
        "bas3_bars_only": {"source": "final_results_notebooks/review_code.ipynb cells 7,18; example_higher_pauli_weight_mixer.ipynb cell 9",
                           "alpha": 0.14433756729740643, "state": "111000111", "energy": -72.0,
                           "n_states_gt_0.01": 6, "entropy_beta_1.5": 2.5849625007211547},
        "bas3_times5": {"source": "gamma-variation-v0.1.ipynb cells 7,8,13", "alpha": 0.025, "energy_data_states": -240.0,
                        "entropy_beta_1.0": 3.584962500721156},
        "cardinality_4_2": {"source": "gamma-variation-v0.2.ipynb cells 8,9", "alpha": 0.4082482904638631,
                            "entropy_beta_1.0": 2.7573388201725257},
        "recipe_b_n10_seed4": {"source": "example_exact_time_evolution.ipynb cells 1,3", "n_states_gt_0.01": 12,
                               "entropy_beta_1.0": 2.839932901889561},
        "random_ising_model_n5": {"source": "sampling_checks.ipynb cells 8-11", "beta": 1.0134,
                                  "seeds": [23564, 202687, 742410, 156407], "n_states_gt_0.01": [16, 18, 19, 20],
                                  "entropy": None},
        "rough_check_n4_seed6": {"source": "qumcmc/rough_check.ipynb cell 1", "alpha": 1.0885895376867156},
        "qulacs_rotation_sign": {"source": "rough.ipynb cells 24,26",
                                 "description": "exp(i*angle/2*(X0+X1)), angle=1, applied to ket '10'",
                                 "amplitudes_re_im": [[0.0, 0.42073549240394786], [-0.22984884706593006, 0.0],
                                                      [0.7701511529340689, 0.0], [0.0, 0.420735492403948]]},
        "readme_model_summary": {"source": "README.md:33-44 (unseeded; alpha follows from the printed means only loosely)",
                                 "alpha": 0.5606804251097042},
    }
    with open(os.path.join(ROOT, "tests", "golden", "known_answers.json"), "w") as f:
        json.dump(known, f, indent=1)
    print("wrote chain_stats.json, known_answers.json")


if __name__ == "__main__":
    main()
