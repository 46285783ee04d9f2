"""Distribution helpers of the reference's ``qumcmc.prob_dist`` (host-side analysis; kept so that
``Exact_Sampling.boltzmann_pd`` and user code see the same types and the same numbers).

Reference: /root/reference/qumcmc/prob_dist.py.
"""
from __future__ import annotations

from typing import Dict

import numpy as np

from .basic_utils import value_sorted_dict  # noqa: F401  (re-exported like the reference)

EPS = 1e-8


class DiscreteProbabilityDistribution(dict):
    """dict {bitstring: probability}; renormalised on construction when the Python ``sum`` of the
    values is not exactly 1.0 (prob_dist.py:14-19, quirk Q8)."""

    def __init__(self, distribution: dict) -> None:
        super().__init__(distribution)
        if sum(list(distribution.values())) != 1.0:
            self._normalise()

    def _normalise(self, print_normalisation: bool = False):
        total = np.sum(list(self.values()))
        if print_normalisation:
            print("Normalisation : ", total)
        for key in list(self.keys()):
            self[key] = self[key] / total

    normalise = _normalise

    def value_sorted_dict(self, reverse: bool = False) -> dict:
        return dict(sorted(self.items(), key=lambda kv: kv[1], reverse=reverse))

    def index_sorted_dict(self, reverse: bool = False) -> dict:
        return dict(sorted(self.items(), key=lambda kv: kv[0], reverse=reverse))

    def get_truncated_distribution(self, epsilon: float = 0.00001, inplace: bool = False):
        kept = {k: v for k, v in self.items() if v > epsilon}
        if inplace:
            self.clear()
            self.__init__(kept)
            return None
        return DiscreteProbabilityDistribution(kept)

    def expectation(self, dict_observable_val_at_states: dict):
        return np.sum([self[k] * dict_observable_val_at_states[k] for k in self.keys()])

    def get_entropy(self):
        """-sum p log2 p over p > 1e-5 in descending order; None if the cut is never reached
        (prob_dist.py:89-96, quirk Q7)."""
        entropy = 0
        for val in sorted(np.array(list(self.values())), reverse=True):
            if val > 0.00001:
                entropy += -1 * val * np.log2(val)
            else:
                return entropy

    def get_observable_expectation(self, observable) -> float:
        return np.sum([self[config] * observable(config) for config in self.keys()])

    def get_sample(self, num_samples, seed=None) -> list:
        return list(np.random.choice(list(self.keys()), p=list(self.values()), size=num_samples))


def kl_divergence(dict_p: dict, dict_q: dict, prelim_check: bool = False):
    """KL(p||q), natural log, q floored at EPS (prob_dist.py:130-151)."""
    kl = 0
    for bitstring, p in dict_p.items():
        q = dict_q[bitstring] if bitstring in dict_q else 0.0
        kl += p * np.log(p) - p * np.log(max(EPS, q))
    return kl


def vectoried_KL(target_vector, model_vector):
    """Array form used by the running-KL curves (prob_dist.py:154-161)."""
    keep = target_vector > 1e-10
    t = target_vector[keep]
    m = model_vector[keep]
    m = np.where(m < EPS, EPS, m)
    return np.sum(t * np.log(t) - t * np.log(m))


def js_divergence(dict_p: dict, dict_q: dict, prelim_check=True):
    """JS(p, q) built from the EPS-floored KL above (prob_dist.py:164-211)."""
    if prelim_check:
        if sorted(dict_p.keys()) != sorted(dict_q.keys()):
            for key in set(dict_p.keys()) | set(dict_q.keys()):
                dict_p.setdefault(key, 0.0)
                dict_q.setdefault(key, 0.0)
        eps = 1e-6
        assert np.abs(np.sum(list(dict_p.values())) - 1.0) <= eps, "sum of values of dict_p must be 1."
        assert np.abs(np.sum(list(dict_q.values())) - 1.0) <= eps, "sum of values of dict_q must be 1."
    p = DiscreteProbabilityDistribution(dict_p).index_sorted_dict()
    q = DiscreteProbabilityDistribution(dict_q).index_sorted_dict()
    p_arr = np.array(list(p.values()))
    q_arr = np.array(list(q.values()))
    m = dict(zip(p.keys(), np.round(0.5 * (p_arr + q_arr), decimals=12)))
    return 0.5 * (kl_divergence(p, m) + kl_divergence(q, m))
