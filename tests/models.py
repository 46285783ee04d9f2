"""Model recipes used by the reference's notebooks/tests (host-side numpy only)."""
import json
import os

import numpy as np

from qumcmc_b200.basic_utils import bas_dataset, get_cardinality_dataset, hebbing_learning

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def bas3(which="both", scale=1.0):
    """BAS 3x3: J = -scale * hebb(data), h = 0 (testing.py:16-27, tutorial.ipynb)."""
    bas = bas_dataset(grid_size=3)
    data = {"both": bas.bas_dict["bars"] + bas.bas_dict["stripes"], "bars": bas.bas_dict["bars"],
            "stripes": bas.bas_dict["stripes"]}[which]
    return -scale * hebbing_learning(data), np.zeros(9), data


def cardinality_model(n=4, card=2):
    data = get_cardinality_dataset(n, card)
    return -1 * hebbing_learning(data), np.zeros(n), data


def recipe_b(n, seed):
    """define_random_spin_ising_model (example_higher_pauli_weight_mixer.ipynb cell 4)."""
    np.random.seed(seed)
    np.random.choice([+1, 0, -1], size=(n, n))  # discarded draw present in the notebook recipe
    J = np.random.uniform(low=-2, high=2, size=(n, n))
    J = 0.5 * (J + J.transpose())
    J = np.round(J - np.diag(np.diag(J)), decimals=3)
    h = np.round(np.random.uniform(low=-1, high=1, size=n), decimals=2)
    return J, h


def readme_model(n=10, seed=610358):
    """README.md:16-32 recipe; the README is unseeded, seed fixed per SURVEY section 8(d)."""
    np.random.seed(seed)
    J = np.random.uniform(low=-2, high=2, size=(n, n))
    J = 0.5 * (J + J.transpose())
    J = np.round(J - np.diag(np.diag(J)), decimals=3)
    h = np.round(0.5 * np.random.randn(n), decimals=2)
    return J, h


def packaged_random_model(n, seed):
    """random_ising_model (qumcmc/energy_models.py:419-446)."""
    np.random.seed(seed)
    J = np.random.uniform(low=-1, high=1, size=(n, n))
    J = 0.5 * (J + J.transpose())
    J = np.round(J - np.diag(np.diag(J)), decimals=3)
    h = np.round(0.4 * np.random.randn(n), decimals=2)
    return J, h


def one_body(n):
    return [1 << q for q in range(n)], [1.0] * n
