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


def test_freqbins_num_freq_host_entry_point():
    """sgw_freqbins_num_freq is pure host logic (freqbins_symm, freqbins.f90:243-305): same answers as the oracle, and the
    reference's error for two zero frequencies."""
    import numpy as np
    from oracle import sigma as osg
    from sternheimergw_b200 import SgwError, freqbins_type
    for solver in ([0.0, 0.3j, 0.9j], [0.2j, 0.5j], [0.0], [1e-15, 0.4j, 2.0 + 0.1j]):
        for symm in (0, 1, 2):
            assert freqbins_type(np.array(solver, dtype=complex), freq_symm_coul=symm).num_freq() == \
                osg.freqbins_symm(np.array(solver, dtype=complex), symm).size
    with pytest.raises(SgwError):
        freqbins_type(np.array([0.0, 0.0, 0.5j])).num_freq()
    f = freqbins_type(np.array([0.0, 0.5j]), coul=np.array([0.1j, 0.3j]), weight=np.ones(2), sigma=np.array([0j, 1j]))
    assert f.num_coul() == 2 and f.num_sigma() == 2
    assert np.allclose(f.green(1.0), [1 + 0.1j, 1 + 0.3j, 1 - 0.1j, 1 - 0.3j])
    assert f.symmetrize(2j) == 2j and freqbins_type(np.array([0j]), freq_symm_coul=2).symmetrize(2j) == -4


def test_host_frequency_meshes_match_oracle_and_numpy():
    """Host mirror of freqbins / gauleg_grid (freqbins.f90:109-180, gauleg_grid.f90:23) vs the oracle and numpy's leggauss."""
    import numpy as np
    from oracle import sigma as osg
    from sternheimergw_b200.host import freqbins, gauleg_grid
    for n in (1, 2, 5, 12, 35, 51):
        x, w = gauleg_grid(0.0, 7.3, n)
        xo, wo = osg.gauleg_grid(0.0, 7.3, n)
        xr, wr = np.polynomial.legendre.leggauss(n)
        assert np.array_equal(x, xo) and np.array_equal(w, wo)
        assert np.allclose(x, 3.65 + 3.65 * xr, atol=1e-12) and np.allclose(w, 3.65 * wr, atol=1e-12)
    for imag, eta in ((True, 0.0), (False, 0.07)):
        a = freqbins(imag, -0.5 if not imag else 0.0, 1.0, 6, 3.0, 9, [0.0, 0.3j, 0.9j], eta=eta)
        b = osg.freqbins(imag, -0.5 if not imag else 0.0, 1.0, 6, 3.0, 9, [0.0, 0.3j, 0.9j], eta=eta)
        assert np.array_equal(a.coul, b.coul) and np.array_equal(a.weight, b.weight) and np.array_equal(a.sigma, b.sigma)
        assert a.imag_sigma == b.imag_sigma and a.num_freq() == b.num_freq() == 5
