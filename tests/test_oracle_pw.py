"""Anchor the plane-wave half of the oracle (QE semantics, unpinned by reference tests) with numpy formulas."""
import numpy as np
import pytest

import oracle
import sos
import synth


@pytest.mark.parametrize("nr", [(15, 15, 15), (18, 18, 60), (20, 20, 20), (24, 24, 24), (8, 9, 10), (72, 4, 6)])
def test_fft_conventions(nr):
    rng = np.random.default_rng(11)
    f = rng.standard_normal(nr) + 1j * rng.standard_normal(nr)
    assert np.abs(oracle.fft3d(f, +1) - np.fft.ifftn(f) * f.size).max() < 1e-11    # invfft: unscaled e^{+iGr}
    assert np.abs(oracle.fft3d(f, -1) - np.fft.fftn(f) / f.size).max() < 1e-14     # fwfft: scaled 1/nnr


def test_linear_op_vs_dense(tiny_sys):
    s = tiny_sys
    ps = oracle.PwSystem(s)
    rng = np.random.default_rng(1)
    for ik in (0, 3):
        kq = s.kpairs[ik].kq
        H = synth.dense_h(s, kq.mill, kq.g2kin, kq.vkb[:kq.npw], kq.dion)
        assert np.abs(H @ kq.evq[:kq.npw] - kq.evq[:kq.npw] * kq.et[:4]).max() < 1e-12
        psi = np.zeros(kq.npwx, complex)
        psi[:kq.npw] = rng.standard_normal(kq.npw) + 1j * rng.standard_normal(kq.npw)
        for om, apv in ((0.3 + 0.2j, kq.alpha_pv), (-0.7j, 0.0)):
            out = ps.linear_op(ik, om, apv, psi)
            P = kq.evq[:kq.npw] @ kq.evq[:kq.npw].conj().T
            ref = (H + om * np.eye(kq.npw) + apv * P) @ psi[:kq.npw]
            assert np.abs(out[:kq.npw] - ref).max() < 1e-13 * np.abs(ref).max()
            assert np.abs(out[kq.npw:]).max(initial=0.0) == 0.0


def test_pw_select_solver_vs_dense(tiny_sys):
    s = tiny_sys
    ps = oracle.PwSystem(s)
    kq = s.kpairs[1].kq
    H = synth.dense_h(s, kq.mill, kq.g2kin, kq.vkb[:kq.npw], kq.dion)
    P = kq.evq[:kq.npw] @ kq.evq[:kq.npw].conj().T
    rng = np.random.default_rng(5)
    b = rng.standard_normal(kq.npw) + 1j * rng.standard_normal(kq.npw)
    sig = -(kq.et[0] + np.array([0.0, 0.5j, 1.5j, -0.5j, -1.5j]))
    for prio in ((1,), (3,)):
        x, ierr, st = ps.select_solver(1, b, sig, oracle.make_cfg(priority=prio, threshold=1e-12))
        assert ierr == 0
        for i, sg in enumerate(sig):
            xd = np.linalg.solve(H + sg * np.eye(kq.npw) + kq.alpha_pv * P, b)
            assert np.linalg.norm(x[:, i] - xd) < 1e-9 * np.linalg.norm(xd)


def test_coulomb_vs_sum_over_states(tiny_sys):
    """eps_{G'G}(q,w) from the Sternheimer pipeline == independent sum-over-states chi0 formula."""
    s = tiny_sys
    ps = oracle.PwSystem(s)
    fiu = np.array([0.0, 1.2j])
    ngc = 9
    cfg = oracle.make_cfg(priority=(1, 3), threshold=1e-12)
    scr, ierr, st = ps.coulomb(1, ngc, ngc, np.arange(1, ngc + 1), fiu, cfg, nthreads=4)
    assert ierr == 0
    for ig in (1, 2, 5, 9):
        ref = sos.eps_sos(s, ig, ngc, fiu)
        assert np.abs(scr[:, :, ig - 1] - ref).max() < 1e-10
    # partial block through igstart (do_stern.f90:199-209 hands every image a contiguous block)
    scr2, ierr, _ = ps.coulomb(4, ngc, 3, np.arange(1, ngc + 1), fiu, cfg)
    assert np.abs(scr2 - scr[:, :, 3:6]).max() < 1e-12
    # head element: coulomb_q0G0 == the (1,1) element of the first perturbation
    eps_m, ierr, _ = ps.coulomb_q0G0(fiu, cfg)
    assert np.abs(eps_m - scr[0, :, 0]).max() < 1e-12
    assert eps_m[0].real > 1.0 and abs(eps_m[0].imag) < 1e-10


def test_invert_epsilon_and_unfold():
    rng = np.random.default_rng(2)
    ngc, nfs = 7, 3
    scr_in = rng.standard_normal((ngc, nfs, ngc)) + 1j * rng.standard_normal((ngc, nfs, ngc))
    scr_in += 4 * np.eye(ngc)[:, None, :]
    full = oracle.unfold_w(ngc, nfs, np.arange(1, ngc + 1), scr_in)
    for iw in range(nfs):
        assert np.array_equal(full[:, :, iw], np.conj(scr_in[:, iw, :]).T)         # unfold_w.f90:84
    inv, info = oracle.invert_epsilon(full)
    assert info == 0
    for iw in range(nfs):
        assert np.abs(inv[:, :, iw] + np.eye(ngc) - np.linalg.inv(full[:, :, iw])).max() < 1e-13
    inv_g, _ = oracle.invert_epsilon(full, lgamma=True)
    for iw in range(nfs):
        a = full[:, :, iw].copy()
        a[1:, 0] = 0
        a[0, 1:] = 0
        r = np.linalg.inv(a)
        r[1:, 0] = 0
        r[0, 1:] = 0
        assert np.abs(inv_g[:, :, iw] + np.eye(ngc) - r).max() < 1e-13


def test_green_function(tiny_sys):
    """(H - w) G = -delta  (green.f90:105-226) against dense inverses, including the strict '<' mask of :212."""
    s = tiny_sys
    ps = oracle.PwSystem(s)
    kq = s.kpairs[0].kq
    H = synth.dense_h(s, kq.mill, kq.g2kin, kq.vkb[:kq.npw], kq.dion)
    ngc = 9
    pos = {int(g): i + 1 for i, g in enumerate(kq.igk)}
    map_ = np.array([pos.get(ig, 0) for ig in range(1, ngc + 1)], dtype=np.int32)
    fft_map = np.arange(1, ngc + 1, dtype=np.int32)
    mu = 0.5 * (kq.et[3] + kq.et[4])
    omega = mu + np.array([0.3j, 2.0j, -0.3j, -2.0j])
    green, ierr, st = ps.green_function(0, map_, fft_map, omega, oracle.make_cfg(priority=(1, 3), threshold=1e-12))
    assert ierr == 0
    for ifr, w in enumerate(omega):
        Gd = -np.linalg.inv(H - w * np.eye(kq.npw))
        for igp in range(ngc):
            if map_[igp] == 0:
                continue
            for ig in range(ngc):
                want = Gd[map_[ig] - 1, map_[igp] - 1] if 0 < map_[ig] < kq.npw else 0.0
                assert abs(green[ig, igp, ifr] - want) < 1e-10


def test_solve_linter_selfconsistent_vs_dense_inverse(tiny_sys):
    """the self-consistent branch (solve_linter.f90:376-460,:564-582 + mix_potential_c) must converge to
    dV_scf = (eps^-1 - 1) dV_bare, where eps is the FULL direct dielectric matrix over the density sphere -- an
    independent formula (dense inverse of the direct branch's output) for the same quantity."""
    syn = tiny_sys
    ps = oracle.PwSystem(syn)
    fiu = np.array([0.0, 1.2j])
    ngm = syn.ngm
    igu = np.arange(1, ngm + 1, dtype=np.int32)
    E, ierr, _ = ps.coulomb(1, ngm, ngm, igu, fiu, oracle.make_cfg(priority=(1, 3), threshold=1e-12), nthreads=8)
    assert ierr == 0
    nnr = int(np.prod(syn.nr))
    for ig in (3, 10):
        dv = np.zeros(nnr, dtype=complex)
        dv[syn.nl[ig - 1] - 1] = 1.0
        dvr = (np.fft.ifftn(dv.reshape(syn.nr, order="F")) * nnr).reshape(-1, order="F")      # invfft (coulomb.f90:134)
        out, ierr, st = ps.solve_linter_iter(40, 0.7, 1e-22, 4, dvr, fiu, oracle.make_cfg(priority=(1, 3), threshold=1e-4),
                                             nthreads=8)
        assert ierr == 0 and 3 < st["iter"] < 40
        for iw in range(fiu.size):
            W = (np.fft.fftn(out[:, iw].reshape(syn.nr, order="F")) / nnr).reshape(-1, order="F")[syn.nl - 1]
            ref = np.linalg.inv(E[:, iw, :])[:, ig - 1].copy()
            ref[ig - 1] -= 1.0
            assert np.abs(W - ref).max() < 1e-8 * max(1.0, np.abs(ref).max())
    # too few iterations: the reference aborts (solve_linter.f90:588-591) -> code 10
    _, ierr, st = ps.solve_linter_iter(3, 0.7, 1e-22, 4, dvr, fiu, oracle.make_cfg(priority=(1, 3), threshold=1e-4), nthreads=8)
    assert ierr == 10 and st["iter"] == 3
