"""Shared fixtures.  GPU tests are marked @pytest.mark.gpu and call the product through the C ABI;
everything else runs on CPU (oracle vs golden vectors, host logic, ABI symbol checks)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


class Golden:
    """Lazy access to tests/golden/*.npz (outputs of the live reference; make_golden.py)."""

    def __init__(self):
        self._cache = {}

    def file(self, name):
        if name not in self._cache:
            self._cache[name] = np.load(os.path.join(GOLDEN, name + ".npz"))
        return self._cache[name]

    def case(self, file, prefix):
        z = self.file(file)
        p = prefix + "__"
        return {k[len(p):]: z[k] for k in z.files if k.startswith(p)}

    def points(self):
        z = self.file("fixture_points")
        return z["ptsA"], z["ptsB"]


@pytest.fixture(scope="session")
def golden():
    return Golden()


FIXTURE_CASES = ["A_s0", "A_s1", "A_s2", "B_s0", "B_s1", "B_s2", "R_s0", "K_s0", "E_s1"]
CASE_METHOD = {"R_s0": "reproj", "K_s0": "backward"}


def case_method(name):
    return CASE_METHOD.get(name, "fwd")


def synth_pair(seed, M, outlier_frac=0.5, noise=0.5):
    """Synthetic correspondences of SURVEY.md 8(d) C3 (same generator as make_golden.py)."""
    rs = np.random.RandomState(seed)
    X = (rs.rand(2, M) * np.array([[1920.0], [1080.0]])).astype(np.float32)
    Ht = np.array([[0.9, 0.05, 30], [-0.04, 1.1, -20], [1e-5, -2e-5, 1]])
    q = Ht @ np.vstack([X.astype(np.float64), np.ones(M)])
    Y = q[:2] / q[2] + rs.randn(2, M) * noise
    out = rs.rand(M) < outlier_frac
    Y[:, out] = rs.rand(2, int(out.sum())) * np.array([[1920.0], [1080.0]])
    return X, Y.astype(np.float32)


def descriptors(rs, n, nbytes=32, dup=0.3):
    """Random descriptors with many near-duplicates so that equal distances (tie-breaking) actually occur."""
    d = rs.randint(0, 256, (n, nbytes)).astype(np.uint8)
    for i in range(1, n):
        if rs.rand() < dup:
            d[i] = d[rs.randint(0, i)]
            flips = rs.randint(0, 3)
            for _ in range(flips):
                d[i, rs.randint(0, nbytes)] ^= np.uint8(1 << rs.randint(0, 8))
    return d
