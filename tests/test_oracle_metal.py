"""Metals (klist lgauss): the smeared projector of [QE] LR_Modules/orthogonalize.f90 and the wg/wk scaling of solve_linter.f90:373
in the oracle.  QE's sources are not part of the reference tree and the reference has no unit test for this path, so no reference
vector can pin the restatement; it is anchored here by checks that do not depend on it:
  * the static eps column equals FINITE-TEMPERATURE PERTURBATION THEORY summed over all pairs of eigenstates (dense
    diagonalisation, tests/sos.py) for Gaussian, Methfessel-Paxton, cold and Fermi-Dirac smearing -- the physics-level anchor;
  * the static density response does not depend on alpha_pv: with De = e_j - e_i the |j> component of the solution is
    [theta~_i - w(j,i)] / (De + alpha) and w = theta~_i (1 - theta) + theta~_j theta + alpha theta (theta~_j - theta~_i) / De makes
    the numerator theta (theta~_i - theta~_j) (De + alpha) / De, i.e. the result theta_ji (theta~_i - theta~_j) / De for every alpha;
  * w0gauss is the derivative of wgauss for every smearing type, with the right limits;
  * with the Fermi level in the gap and a tiny smearing the metallic branch reproduces the insulator branch."""
import numpy as np
import pytest

import oracle
import synth


from metal_util import metal_system as _metal_system


@pytest.mark.parametrize("n", [-99, -1, 0, 1, 2])
def test_w0gauss_is_the_derivative_of_wgauss(n):
    h = 1e-5
    for x in np.linspace(-4.0, 4.0, 33):
        d = (oracle.wgauss(x + h, n) - oracle.wgauss(x - h, n)) / (2 * h)
        assert abs(d - oracle.w0gauss(x, n)) < 1e-8, (n, x)
    assert abs(oracle.wgauss(-45.0, n)) < 1e-12 and abs(oracle.wgauss(45.0, n) - 1.0) < 1e-12
    if n in (-99, 0):
        assert abs(oracle.wgauss(0.0, n) - 0.5) < 1e-15


def test_metal_weight_limits():
    # both states far below the Fermi level: the insulator projector (weight 1); state j far above: 0
    assert abs(oracle.metal_weight(-1.0, -0.8, True, 0.7, 0.0, 0.01, 0) - 1.0) < 1e-12
    assert abs(oracle.metal_weight(-1.0, +0.9, False, 0.7, 0.0, 0.01, 0)) < 1e-12
    # degenerate pair: the 0/0 limit is continuous
    a = oracle.metal_weight(0.01, 0.01 + 2e-5, True, 0.7, 0.0, 0.05, 0)
    b = oracle.metal_weight(0.01, 0.01 + 0.5e-5, True, 0.7, 0.0, 0.05, 0)
    assert abs(a - b) < 5e-3 * abs(a)


def test_insulator_limit_of_the_metallic_branch():
    s = synth.build_lattice("tiny", 10.26, synth.FCC, [[0.125] * 3, [-0.125] * 3], ["Si", "Si"], 6.0)
    ins = synth.attach_kpoints(s, synth.mp_grid(s.bg, 1), [0.5, 0.5, 0.5])
    nocc = ins.nbnd_occ
    s2 = synth.build_lattice("tiny", 10.26, synth.FCC, [[0.125] * 3, [-0.125] * 3], ["Si", "Si"], 6.0)
    met = synth.attach_kpoints(s2, synth.mp_grid(s2.bg, 1), [0.5, 0.5, 0.5], nbnd=nocc + 3)
    top = max(max(kp.et[nocc - 1], kp.kq.et[nocc - 1]) for kp in met.kpairs)
    bot = min(min(kp.et[nocc], kp.kq.et[nocc]) for kp in met.kpairs)
    assert bot - top > 0.02
    for kp in met.kpairs:                      # same alpha_pv as the insulator (attach_kpoints derives it from the band count)
        kp.kq.alpha_pv = ins.kpairs[0].kq.alpha_pv
    synth.make_metal(met, 0.5 * (top + bot), (bot - top) / 40.0, 0, target=3.0)
    assert all(m.nocc_k == nocc for m in met.metal.pairs)
    ngc, igu, fiu = 5, np.arange(1, 6, dtype=np.int32), np.array([0.0, 0.7j])
    cfg = oracle.make_cfg(priority=(1, 3), threshold=1e-12)
    a, ierr_a, _ = oracle.PwSystem(ins).coulomb(1, ngc, ngc, igu, fiu, cfg, nthreads=2)
    b, ierr_b, _ = oracle.PwSystem(met).coulomb(1, ngc, ngc, igu, fiu, cfg, nthreads=2)
    assert ierr_a == 0 and ierr_b == 0
    assert np.abs(a - b).max() < 1e-9 * np.abs(a).max()


@pytest.mark.parametrize("ngauss", [0, -99, 1])
def test_static_response_of_a_metal_does_not_depend_on_alpha_pv(ngauss):
    ngc, igu, fiu = 4, np.arange(1, 5, dtype=np.int32), np.array([0.0])
    cfg = oracle.make_cfg(priority=(1, 3), threshold=1e-12)
    res = []
    for scale in (1.0, 1.9):
        syn = _metal_system(ngauss=ngauss, alpha_scale=scale)
        assert any(abs(w) > 0.01 and abs(w - 1.0) > 0.01 for m in syn.metal.pairs for w in m.wg_over_wk)   # partial occupations
        scr, ierr, _ = oracle.PwSystem(syn).coulomb(1, ngc, ngc, igu, fiu, cfg, nthreads=2)
        assert ierr == 0
        res.append(scr)
    assert np.abs(res[0] - np.eye(ngc)[:, None, :]).max() > 1e-3                        # a response is there
    assert np.abs(res[0] - res[1]).max() < 1e-8 * np.abs(res[0]).max()


def _zincblende(nbnd=None):
    """2-atom zincblende stand-in (no inversion centre: the response matrix is genuinely complex) on the full 2x2x2 mesh, which
    contains -k-q for every k (q is a mesh vector)."""
    s = synth.build_lattice("zb", 10.26, synth.FCC, [[0.125] * 3, [-0.125] * 3], ["Si", "C"], 6.0)
    return synth.attach_kpoints(s, synth.mp_grid(s.bg, 2), [0.5, 0.5, 0.5], nbnd=nbnd)


@pytest.mark.parametrize("ngauss,degauss,nbnd,tol", [(0, 0.05, 8, 1e-9), (1, 0.05, 8, 1e-8), (-1, 0.05, 8, 1e-9), (-99, 0.02, 10, 2e-5)])
def test_metallic_branch_equals_the_finite_temperature_sum_over_states(ngauss, degauss, nbnd, tol):
    """The physics-level anchor: with wg/wk = 1 (the scheme as published; solve_linter.f90:373 is the reference's own extra
    factor) the static eps column of the smeared Sternheimer scheme equals first-order perturbation theory with smeared
    occupations summed over ALL pairs of eigenstates (tests/sos.py: dense eigh, no projector, no pair splitting).  This fixes
    the weights of the projector as a whole -- theta~_i (1 - theta_ji) + theta~_j theta_ji and the alpha_pv term.
    Fermi-Dirac agrees to the occupation QE's band cutoff leaves out (setup_nbnd_occ: w0gauss < 7e-5)."""
    import sos
    syn = _zincblende(nbnd)
    ef = 0.39                                   # between the bands at 0.36 Ry and 0.42 Ry: occupations 0.79 and 0.21
    synth.make_metal(syn, ef, degauss, ngauss)
    for m in syn.metal.pairs:
        m.wg_over_wk[:] = 1.0
    ngc, igu, fiu = 4, np.arange(1, 5, dtype=np.int32), np.array([0.0])
    scr, ierr, _ = oracle.PwSystem(syn).coulomb(1, ngc, 1, igu, fiu, oracle.make_cfg(priority=(1, 3), threshold=1e-12), nthreads=4)
    assert ierr == 0
    ref = sos.eps_sos_smeared(syn, 1, ngc, lambda e: oracle.wgauss((ef - e) / degauss, ngauss),
                              lambda e: -oracle.w0gauss((ef - e) / degauss, ngauss) / degauss)
    assert np.abs(scr[:, 0, 0] - np.eye(ngc)[:, 0]).max() > 0.1            # a metallic-size response
    assert np.abs(scr[:, 0, 0] - ref).max() < tol, np.abs(scr[:, 0, 0] - ref).max()


def test_insulator_equals_the_full_double_sum():
    """The same double sum with step-function occupations reproduces the insulator branch: it checks the normalisation of
    tests/sos.py's finite-temperature formula (the factor 2 of incdrhoscf = the time-reversed partner of every pair)."""
    import sos
    syn = _zincblende()
    top = max(max(kp.et[3], kp.kq.et[3]) for kp in syn.kpairs)
    ngc, igu, fiu = 4, np.arange(1, 5, dtype=np.int32), np.array([0.0])
    scr, ierr, _ = oracle.PwSystem(syn).coulomb(1, ngc, 1, igu, fiu, oracle.make_cfg(priority=(1, 3), threshold=1e-12), nthreads=4)
    assert ierr == 0
    ref = sos.eps_sos_smeared(syn, 1, ngc, lambda e: 1.0 if e < top + 0.05 else 0.0, lambda e: 0.0)
    assert np.abs(scr[:, 0, 0] - ref).max() < 1e-10
