import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# app/params/amhmcl.yaml of the reference
YAML_PARAMS = dict(alpha1=0.002, alpha2=0.03, alpha3=0.08, alpha4=0.002, sigma_hit=0.3,
                   z_hit=0.75, z_rand=0.25, max_range=5.0, step=1)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def oracle_map_world():
    from oracle import node_glue as ng
    g = golden("map_world.npz")
    return ng.load_map(g["occ"], float(g["resolution"]), float(g["origin"][0]), float(g["origin"][1]))


@pytest.fixture(scope="session")
def oracle_map_house():
    from oracle import node_glue as ng
    g = golden("map_house.npz")
    return ng.load_map(g["occ"], float(g["resolution"]), float(g["origin"][0]), float(g["origin"][1]))


def split_normals(seed, counts, pad_to=None):
    """Regenerate the reference's MT19937 normal stream (RandomState legacy) and split it per
    particle by the attempt counts recorded in the golden file -> (N, A, 3)."""
    counts = np.asarray(counts)
    z = np.random.RandomState(int(seed)).normal(0, 1, int(counts.sum()) * 3).reshape(-1, 3)
    A = int(pad_to or counts.max())
    out = np.zeros((len(counts), A, 3))
    off = 0
    for i, c in enumerate(counts):
        out[i, :c] = z[off:off + c]
        off += c
    return out
