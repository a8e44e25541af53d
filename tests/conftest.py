import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def lin_prob():
    """The reference's golden vector (algo/linear_solver/test/lin_prob.xml.bz2 -> tools/iotk_read.py)."""
    d = np.load(ROOT / "tests" / "golden" / "lin_prob.npz")
    A = np.asfortranarray(d["A_real"].astype(np.complex128))
    return {"n": int(d["n"]), "ns": int(d["ns"]), "A": A, "b": d["b"].astype(np.complex128),
            "sigma": d["sigma"].astype(np.complex128), "x_bad": d["x_bad"]}


@pytest.fixture(scope="session")
def tiny_sys():
    import synth
    return synth.preset("tiny")
