"""CDTraining on the device samplers: the vectorised gradient equals the reference's
per-parameter cd_J / cd_h (training.py:70-99), and training lowers the KL divergence to the data."""
import numpy as np
import pytest

from tests import models


def _bas_distribution(q, grid=2):
    ds = q.bas_dataset(grid_size=grid)
    states = ds.dataset["stripes"] + ds.dataset["bars"] if isinstance(ds.dataset, dict) else list(ds.dataset)
    states = sorted(set(states))
    return q.DiscreteProbabilityDistribution({s: 1.0 / len(states) for s in states})


def test_data_moments_equal_the_dictionary_expectations():
    import qumcmc_b200 as q
    from qumcmc_b200.training import correlation_spins, data_moments
    rng = np.random.default_rng(0)
    keys = [f"{i:05b}" for i in rng.choice(32, size=9, replace=False)]
    p = rng.random(9)
    dist = q.DiscreteProbabilityDistribution(dict(zip(keys, p / p.sum())))
    m1, m2 = data_moments(dist)
    for i in range(5):
        assert abs(m1[i] - dist.get_observable_expectation(lambda s: correlation_spins(s, [i]))) < 1e-14
        for j in range(5):
            assert abs(m2[i, j] - dist.get_observable_expectation(lambda s: correlation_spins(s, [i, j]))) < 1e-14


@pytest.mark.gpu
@pytest.mark.parametrize("n,count", [(5, 1000), (20, 70000), (36, 5000)])
def test_state_moments_are_exact_counts(n, count):
    import qumcmc_b200 as q
    from qumcmc_b200.training import spin_moments
    J, h = models.packaged_random_model(n, 1)
    m = q.IsingEnergyFunction(J, h)
    rng = np.random.default_rng(n)
    states = rng.integers(0, 1 << n, size=count, dtype=np.uint64)
    mask = (rng.random(count) < 0.6).astype(np.uint8)
    N, m1, m2 = spin_moments(m, states, mask)
    sel = states[mask != 0]
    S = np.array([[1.0 if (int(s) >> (n - 1 - i)) & 1 else -1.0 for i in range(n)] for s in sel[:4000]])
    assert N == len(sel)
    if len(sel) <= 4000:
        assert np.array_equal(m1, S.mean(axis=0)) or np.allclose(m1, S.mean(axis=0), atol=1e-15)
        assert np.allclose(m2, S.T @ S / len(sel), atol=1e-14)
    else:  # bit-matrix form for the large case
        bits = ((sel[:, None] >> (n - 1 - np.arange(n, dtype=np.uint64))[None, :]) & np.uint64(1)).astype(np.float64)
        Sb = 2.0 * bits - 1.0
        assert np.allclose(m1, Sb.mean(axis=0), atol=1e-14) and np.allclose(m2, Sb.T @ Sb / len(sel), atol=1e-13)
    N2, _, _ = spin_moments(m, states)
    assert N2 == count


@pytest.mark.gpu
def test_vectorised_gradient_equals_reference_cd_functions():
    import qumcmc_b200 as q
    from qumcmc_b200.training import CDTraining
    n = 4
    data = _bas_distribution(q)
    m = q.random_ising_model(n, 2)
    tr = CDTraining(m, 1.0, data)
    np.random.seed(0)
    chain = q.classical_mcmc(400, tr.model, q.FixedWeightProposals(n, 1), initial_state="0101", temperature=1.0, seed=3)
    gh, gJ = tr.gradients(chain)
    for i in range(n):
        assert abs(gh[i] - tr.cd_h(i, chain)) < 1e-12
        for j in range(i):
            assert abs(gJ[i, j] - tr.cd_J([i, j], chain)) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("sampler", ["classical", "quantum", "quantum-batch"])
def test_training_lowers_the_kl_divergence(sampler):
    import qumcmc_b200 as q
    from qumcmc_b200.training import CDTraining
    n = 4
    data = _bas_distribution(q)
    np.random.seed(1)
    m = q.IsingEnergyFunction(np.zeros((n, n)), np.zeros(n))
    tr = CDTraining(m, 1.0, data)
    if sampler == "classical":
        sm = q.ClassicalMCMCSampler(100, tr.model, q.UniformProposals(n), temperature=1.0)
    elif sampler == "quantum":
        sm = q.QuantumMCMCSampler(100, tr.model, q.GenericMixer.OneBodyMixer(n), (0.25, 0.6), temperature=1.0)
    else:
        sm = q.QuantumMCMCSampler(100, tr.model, q.GenericMixer.OneBodyMixer(n), (0.25, 0.6), temperature=1.0, n_chains=32)
    tr.train(sm, mcmc_steps=600, lr=0.05, epochs=40, verbose=False)
    kl = tr.training_history["kl_div"]
    assert len(kl) == 40 and len(tr.training_history["max-min-gradient"]) == 40
    assert kl[-1] < 0.6 * kl[0]
    assert np.allclose(tr.model.J, tr.model.J.T) and np.all(np.diag(tr.model.J) == 0)
    tr.train(sm, mcmc_steps=200, lr=0.05, epochs=2, verbose=False,
             update_strategy=["random", {"num_random_bias": 2, "num_random_interactions": 2}],
             save_training_data={"kl_div": [True, "last-mcmc-chain"], "max-min-gradient": True})
    assert len(tr.training_history["kl_div"]) == 42
