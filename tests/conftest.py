import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_oracle():
    """The CPU oracle (test infrastructure only)."""
    odir = os.path.join(ROOT, "oracle")
    if odir not in sys.path:
        sys.path.insert(0, odir)
    import fem_np
    return fem_np


@pytest.fixture(scope="session")
def oracle():
    return load_oracle()


@pytest.fixture(scope="session")
def golden_meshes(tmp_path_factory):
    """Re-emit the reference's test meshes (stored as arrays) as gmsh 2.2 files with the repo's own writer."""
    from scatter_b200 import gmsh_io
    G = np.load(os.path.join(GOLDEN, "meshes.npz"))
    out = tmp_path_factory.mktemp("meshes")
    keys = sorted({k.split("__")[0] for k in G.files})
    paths = {}
    for key in keys:
        phys = []
        for p in G[key + "__phys"]:
            d, t, n = str(p).split("|")
            phys.append([float(d), int(t), n])
        et = gmsh_io.GMSH_TO_TYPE[int(G[key + "__gmsh_type"])]
        path = os.path.join(out, key + ".msh")
        gmsh_io.write_msh(path, G[key + "__nodes"], G[key + "__elem"], G[key + "__tags"], phys, et)
        paths[key + ".msh"] = path
    return paths


@pytest.fixture(scope="session")
def golden_matrices():
    return np.load(os.path.join(GOLDEN, "matrices.npz"))


@pytest.fixture(scope="session")
def golden_histories():
    return np.load(os.path.join(GOLDEN, "histories.npz"))


@pytest.fixture(scope="session")
def golden_elements():
    return np.load(os.path.join(GOLDEN, "elements.npz"))


def probe_vector(n):
    i = np.arange(n, dtype=np.float64)
    return np.sin(0.37 * i + 0.11) + 0.25 * np.cos(1.3 * i)


def rel_l2(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    nb = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (nb if nb > 0 else 1.0))


def max_rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    s = np.abs(b).max()
    return float(np.abs(a - b).max() / (s if s > 0 else 1.0))
