import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

GOLDEN = os.path.join(ROOT, "tests", "golden")
MESHES = os.path.join(GOLDEN, "meshes")
GOLDEN_CASES = ["3d_2uni_n8", "3d_2refine_n8", "3d_2refine_d1_n4", "3d_multi_refine_n4",
                "2d_2d2ref_d1_n8", "2d_multi_refine_8_n4"]


NEUMANN_CASES = ["3d_2refine_n8_neumann", "2d_2d2ref_d1_n8_neumann", "3d_2uni_n16_neumann"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def meshes():
    return MESHES


def load_golden(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def rel_l2(a, b):
    import numpy as np
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
