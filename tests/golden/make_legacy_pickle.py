#!/usr/bin/env python
"""Writes tests/golden/legacy_chains.pkl: the first 300 states of two chains recorded by the reference
(final_results_notebooks/results_with_new_code/BAS_new_code_final_plots_wt1.pkl), re-pickled in the REFERENCE's class
layout (module qumcmc.basic_utils, dataclass MCMCState(bitstring, accepted), MCMCChain with name / _states /
_current_state / _states_accepted / markov_chain: basic_utils.py:54-89) so the legacy loader is tested against the
byte format the reference writes.  Run in the authoring container only (needs /root/reference); run it in a fresh
interpreter: it registers stand-in classes under the reference's module path."""
import os
import pickle
import sys
import types
from dataclasses import dataclass

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from qumcmc_b200.io import load_legacy_pickle  # noqa: E402

src = "/root/reference/final_results_notebooks/results_with_new_code/BAS_new_code_final_plots_wt1.pkl"
chains = load_legacy_pickle(src)
rows = [[(s.bitstring, bool(s.accepted)) for s in c.states[:300]] for c in list(chains)[:2]]
names = [c.name for c in list(chains)[:2]]

# stand-ins with the reference's attribute layout, registered under its module path
mod = types.ModuleType("qumcmc.basic_utils")


@dataclass
class MCMCState:
    bitstring: str
    accepted: bool


class MCMCChain:
    pass


for cls in (MCMCState, MCMCChain):
    cls.__module__ = "qumcmc.basic_utils"
    cls.__qualname__ = cls.__name__
    setattr(mod, cls.__name__, cls)
sys.modules["qumcmc.basic_utils"] = mod

out = []
for name, row in zip(names, rows):
    ch = MCMCChain()
    ch.name = name
    ch._states = [MCMCState(b, a) for b, a in row]
    ch._current_state = next((s for s in ch._states[::-1] if s.accepted), None)
    ch._states_accepted = [s for s in ch._states if s.accepted]
    mc = [ch._states[0].bitstring]
    for s in ch._states[1:]:
        mc.append(s.bitstring if s.accepted else mc[-1])
    ch.markov_chain = mc
    out.append(ch)
with open(os.path.join(ROOT, "tests", "golden", "legacy_chains.pkl"), "wb") as f:
    pickle.dump({"source": os.path.relpath(src, "/root/reference"), "chains": out}, f, protocol=4)
print("wrote", len(out), "chains,", [len(c._states) for c in out], "states")
