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


def _gpu_usable():
    """True if a CUDA device and the in-tree library are both usable (the GPU tests go through libsgw_b200.so)."""
    try:
        import torch
        if not torch.cuda.is_available():
            return False
        from sternheimergw_b200 import _lib
        _lib.load()
        return True
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """a plain `pytest tests` on a GPU-less machine skips the gpu-marked tests instead of erroring in
    Context(0).  On a GPU box nothing is skipped: a missing library still fails loudly there (`_lib.load()` raises)."""
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items:
        return
    import torch
    if torch.cuda.is_available():
        return                                        # GPU box: run everything, no silent skips
    skip = pytest.mark.skip(reason="no CUDA device in this machine (run on the B200 box: pytest -m gpu)")
    for it in gpu_items:
        it.add_marker(skip)
