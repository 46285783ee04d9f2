"""Drop-in alias: ``import qumcmc`` resolves to the B200-native implementation.

The reference's module names (``qumcmc.basic_utils``, ``qumcmc.energy_models``, ...) are registered
as aliases of the ``qumcmc_b200`` modules, so existing user code and the reference's pickled
``MCMCChain`` objects (which name ``qumcmc.basic_utils.MCMCChain``) load unchanged."""
import importlib
import sys

import qumcmc_b200 as _impl
from qumcmc_b200 import *  # noqa: F401,F403

for _name in ("basic_utils", "energy_models", "mixers", "classical_mixers", "classical_mcmc_routines",
              "quantum_mcmc_routines", "mcmc_sampler_base", "prob_dist", "trajectory_processing", "training"):
    _mod = importlib.import_module(f"qumcmc_b200.{_name}")
    sys.modules[f"{__name__}.{_name}"] = _mod
    globals()[_name] = _mod
