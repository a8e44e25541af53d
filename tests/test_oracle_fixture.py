"""Pin the oracle's solver half against the reference's own golden vector and unit-test assertions.

Mirrors algo/linear_solver/test/linear_solver.pf (test_bicgstab :126-257, test_linear_solver :265-345),
data/algebra/test/gram_schmidt.pf and data/parallel/test/parallel.pf (parallel_task cases).
"""
import numpy as np
import pytest

import oracle


def _res(A, s, x, b):
    return np.linalg.norm(A @ x + s * x - b)


def test_fixture_facts(lin_prob):
    A, b, sigma = lin_prob["A"], lin_prob["b"], lin_prob["sigma"]
    assert lin_prob["n"] == 283 and lin_prob["ns"] == 70
    assert np.flatnonzero(b).tolist() == [0] and b[0] == -1.0
    ev = np.linalg.eigvalsh(A)
    assert abs(ev[0] + 0.41541) < 1e-5 and abs(ev[1] - 0.46803) < 1e-5 and abs(ev[-1] - 14.7515) < 1e-4
    # second half of the shifts = conjugate of the first half (mu +- i w)
    assert np.allclose(sigma[35:], np.conj(sigma[:35]))


@pytest.mark.parametrize("lmax", range(1, 16))
def test_bicgstab_single_shift(lin_prob, lmax):
    """linear_solver.pf:196-221: every shift alone, ierr==0, no NaN, |(A+sigma)x-b| <= 1e-6."""
    A, b, sigma = lin_prob["A"], lin_prob["b"], lin_prob["sigma"]
    thr = 1e-6
    for s in sigma:
        x, ierr, _ = oracle.bicgstab_dense(A, b, [s], lmax=lmax, threshold=thr)
        assert ierr == 0
        assert not np.isnan(x).any()
        assert _res(A, s, x[:, 0], b) <= thr


@pytest.mark.parametrize("lmax", range(1, 16))
def test_bicgstab_multishift(lin_prob, lmax):
    """linear_solver.pf:228-252: all 70 shifts together, residual <= 10*threshold."""
    A, b, sigma = lin_prob["A"], lin_prob["b"], lin_prob["sigma"]
    thr = 1e-6
    x, ierr, _ = oracle.bicgstab_dense(A, b, sigma, lmax=lmax, threshold=thr)
    assert ierr == 0 and not np.isnan(x).any()
    for i, s in enumerate(sigma):
        assert _res(A, s, x[:, i], b) <= 10 * thr


def test_bicgstab_known_answers(lin_prob):
    """Known-answer numbers recorded in SURVEY.md section 8c for the fixture."""
    A, b, sigma = lin_prob["A"], lin_prob["b"], lin_prob["sigma"]
    x, ierr, st = oracle.bicgstab_dense(A, b, sigma, lmax=4, threshold=1e-6)
    assert (ierr, st["n_op"], st["n_outer"]) == (0, 24, 3)
    assert max(_res(A, s, x[:, i], b) for i, s in enumerate(sigma)) == pytest.approx(8.4e-8, rel=0.05)
    for L, nop in ((1, 20), (2, 20), (8, 32)):
        assert oracle.bicgstab_dense(A, b, sigma, lmax=L, threshold=1e-6)[2]["n_op"] == nop
    x, ierr, st = oracle.bicgstab_dense(A, b, sigma, lmax=4, threshold=1e-12)
    assert st["n_op"] == 32
    xd = np.stack([np.linalg.solve(A + s * np.eye(283), b) for s in sigma], axis=1)
    assert np.max(np.linalg.norm(x - xd, axis=0) / np.linalg.norm(xd, axis=0)) < 1e-13


def test_linear_solver(lin_prob):
    """linear_solver.pf:265-345: default config (relative threshold 1e-4)."""
    A, b, sigma = lin_prob["A"], lin_prob["b"], lin_prob["sigma"]
    x, ierr, st = oracle.linear_solver_dense(A, b, sigma, threshold=1e-4)
    assert ierr == 0
    nb = np.linalg.norm(b)
    for i, s in enumerate(sigma):
        assert _res(A, s, x[:, i], b) / nb <= 1e-4


def test_select_solver_fallback(lin_prob):
    """select_solver.f90:121-159: bicgstab with max_iter=1 fails (ierr=1), the chain falls through to solver 3."""
    A, b, sigma = lin_prob["A"], lin_prob["b"], lin_prob["sigma"][:5]
    cfg = oracle.make_cfg(priority=(1,), max_iter=1, threshold=1e-10)
    _, ierr, st = oracle.select_solver_dense(A, b, sigma, cfg)
    assert ierr == 1
    cfg = oracle.make_cfg(priority=(1, 3), max_iter=1, threshold=1e-10)
    _, ierr, st = oracle.select_solver_dense(A, b, sigma, cfg)
    assert ierr == 1 and st["solver_used"] == 3          # max_iter=1 is also too few for solver 3
    cfg = oracle.make_cfg(priority=(2,), max_iter=100, threshold=1e-8)
    x, ierr, st = oracle.select_solver_dense(A, b, sigma, cfg)
    assert ierr == 0 and st["solver_used"] == 2
    for i, s in enumerate(sigma):
        assert _res(A, s, x[:, i], b) <= 1e-7


def test_gram_schmidt():
    """data/algebra/test/gram_schmidt.pf: 300x300 random complex, half basis then extension with first=151."""
    rng = np.random.default_rng(7)
    n = 300
    op = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    vec = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    basis = op @ vec
    half = n // 2
    b1, v1 = oracle.gram_schmidt(1, basis[:, :half], vec[:, :half])
    assert np.abs(b1.conj().T @ b1 - np.eye(half)).max() < 1e-12
    assert np.abs(op @ v1 - b1).max() < 1e-10 * np.abs(b1).max() * n
    b2 = np.concatenate([b1, basis[:, half:]], axis=1)
    v2 = np.concatenate([v1, vec[:, half:]], axis=1)
    b3, v3 = oracle.gram_schmidt(half + 1, b2, v2)
    assert np.abs(b3.conj().T @ b3 - np.eye(n)).max() < 1e-11
    assert np.abs(b3[:, :half] - b1).max() == 0.0
    assert np.abs(op @ v3 - b3).max() < 1e-8


def test_norm():
    rng = np.random.default_rng(3)
    v = rng.standard_normal(1000) + 1j * rng.standard_normal(1000)
    assert oracle.norm(v) == pytest.approx(np.linalg.norm(v), rel=1e-14)
    assert oracle.norm(v * 1e-200) == pytest.approx(np.linalg.norm(v) * 1e-200, rel=1e-13)   # scaled ssq: no underflow


@pytest.mark.parametrize("ntask,nproc,expect", [
    (47, 4, [11, 12, 12, 12]), (2, 4, [0, 0, 1, 1]), (32, 4, [8, 8, 8, 8]), (32, 3, [10, 11, 11]), (47, 1, [47])])
def test_parallel_task(ntask, nproc, expect):
    """data/parallel/test/parallel.pf: remainder goes to the LAST ranks; contiguous 1-based blocks."""
    nxt = 1
    for r in range(nproc):
        first, last, num = oracle.parallel_task(nproc, r, ntask)
        assert num == expect
        assert first == nxt and last == first + expect[r] - 1
        nxt = last + 1
    assert nxt == ntask + 1
