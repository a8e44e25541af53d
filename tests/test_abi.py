"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/sgw_b200.h declares; the GPU-free entry point (parallel_task) matches the reference's unit test;
creating a context without a GPU fails loudly (no CPU fallback)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "sgw_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sgw_[A-Za-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from sternheimergw_b200 import _lib
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/sgw_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared


@pytest.mark.parametrize("ntask,nproc,expect", [
    (47, 4, [11, 12, 12, 12]), (2, 4, [0, 0, 1, 1]), (32, 4, [8, 8, 8, 8]), (32, 3, [10, 11, 11]),
    (47, 2, [23, 24]), (32, 1, [32])])
def test_parallel_task_matches_reference_cases(ntask, nproc, expect):
    """data/parallel/test/parallel.pf:25-205 (npes = 1..4; 32, 47 and 2 tasks)."""
    import oracle
    from sternheimergw_b200 import parallel_task
    nxt = 1
    for r in range(nproc):
        first, last, num = parallel_task(nproc, r, ntask)
        assert num == expect
        assert (first, last) == (nxt, nxt + expect[r] - 1)
        assert (first, last, num) == oracle.parallel_task(nproc, r, ntask)
        nxt = last + 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sternheimergw_b200 import Context, SgwError
    with pytest.raises(SgwError):
        Context(0)
