"""qumcmc_b200 -- B200-native (sm_100a) implementation of quMCMC's quantum-proposal hot path behind
the reference's Python API.  ``import qumcmc`` (the shim package at the repo root) exposes the same
module names as the reference: basic_utils, energy_models, mixers, classical_mixers,
classical_mcmc_routines, quantum_mcmc_routines, mcmc_sampler_base, prob_dist, trajectory_processing, training."""
from .basic_utils import *  # noqa: F401,F403
from .basic_utils import MCMCChain, MCMCState, bas_dataset, get_cardinality_dataset, get_random_state, hebbing_learning
from .classical_mcmc_routines import classical_mcmc, test_accept
from .classical_mixers import ClassicalMixer, CombineProposals, CustomProposals, FixedWeightProposals, UniformProposals
from .energy_models import Exact_Sampling, IsingEnergyFunction, random_ising_model
from .mcmc_sampler_base import ClassicalMCMCSampler, MCMCSampler, QuantumMCMCSampler
from .mixers import CoherentMixerSum, CustomMixer, GenericMixer, IncoherentMixerSum, Mixer
from .prob_dist import DiscreteProbabilityDistribution, js_divergence, kl_divergence, vectoried_KL
from .trajectory_processing import (batch_trajectory_statistics, calculate_running_kl_divergence,
                                    calculate_runnning_magnetisation, get_trajectory_statistics)
from .training import CDTraining
from .io import load_chains, load_legacy_pickle, save_chains
from .quantum_mcmc_routines import (ChainBatch, quantum_enhanced_mcmc, run_chains, run_qmcmc_quantum_ckt,
                                    transition_probabilities)

__version__ = "0.1.0"
