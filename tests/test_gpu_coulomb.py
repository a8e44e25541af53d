"""GPU parity of the screened-Coulomb / Green's-function pipelines (phys/coul, phys/green) against the oracle,
through the C ABI: solve_linter, coulomb, coulomb_q0G0, unfold_w, invert_epsilon, green_function."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from sternheimergw_b200 import Context
    c = Context(0)
    yield c
    c.close()


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("name,nk,ngc,fiu", [
    ("tiny", 2, 9, [0.0, 1.2j]),
    ("tiny", 2, 5, [0.4j, 1.2j, 3.0j]),          # no static frequency: num_omega = 2 nfreq (solve_linter.f90:238-242)
    ("si", 1, 11, [0.0, 16j / 13.605698066]),     # gw_si: FREQUENCIES 0, 16i eV
])
def test_coulomb_matches_oracle(ctx, name, nk, ngc, fiu):
    """eps_{G'G}(q,w) for a block of perturbations: converged (thr 1e-12) <= 1e-8 relative to the oracle."""
    import oracle
    import synth
    from sternheimergw_b200 import select_solver_type
    syn = synth.preset(name, nk=nk)
    ctx.install_system(syn)
    ps = oracle.PwSystem(syn)
    fiu = np.asarray(fiu, dtype=complex)
    igu = np.arange(1, ngc + 1, dtype=np.int32)
    cfg = select_solver_type(priority=(1, 3), threshold=1e-12)
    scr = ctx.coulomb(cfg, 1, ngc, ngc, igu, fiu)
    st = ctx.stats()
    ref, ierr, so = ps.coulomb(1, ngc, ngc, igu, fiu, oracle.make_cfg(priority=(1, 3), threshold=1e-12), nthreads=4)
    assert ierr == 0
    assert _rel(scr, ref) < 1e-8, _rel(scr, ref)
    assert st["n_kernel_launch"] > 0 and st["n_linear_op"] > 0
    # a contiguous sub-block through igstart (do_stern.f90:199-209: every image gets one)
    scr2 = ctx.coulomb(cfg, 3, ngc, 3, igu, fiu)
    assert _rel(scr2, ref[:, :, 2:5]) < 1e-8
    # head: coulomb_q0G0 == the (1,1) element of the first perturbation
    eps_m = ctx.coulomb_q0G0(cfg, fiu)
    assert _rel(eps_m, ref[0, :, 0]) < 1e-8


def test_coulomb_reduced_rho_grid_matches_oracle_and_full_box(ctx, monkeypatch):
    """Si8 supercell (36^3 box): Delta-rho is accumulated on the alias-free reduced box; eps must equal the oracle's
    (full box, [QE] incdrhoscf order) to 1e-8 and the library's own full-box result to rounding."""
    import oracle
    import synth
    from sternheimergw_b200 import select_solver_type
    syn = synth.preset("si8")
    ctx.install_system(syn)
    fiu = np.array([0.0, 0.9j])
    ngc = 7
    igu = np.arange(1, ngc + 1, dtype=np.int32)
    cfg = select_solver_type(priority=(1, 3), threshold=1e-11)
    scr = ctx.coulomb(cfg, 2, ngc, 2, igu, fiu)
    reduced, dims = ctx.rho_grid()
    assert reduced and np.prod(dims) < 0.7 * np.prod(syn.nr), (reduced, dims)
    monkeypatch.setenv("SGW_RHO_GRID", "fine")
    scr_full = ctx.coulomb(cfg, 2, ngc, 2, igu, fiu)
    assert ctx.rho_grid() == (False, tuple(syn.nr))
    monkeypatch.delenv("SGW_RHO_GRID")
    assert _rel(scr, scr_full) < 1e-11, _rel(scr, scr_full)
    ps = oracle.PwSystem(syn)
    ref, ierr, _ = ps.coulomb(2, ngc, 2, igu, fiu, oracle.make_cfg(priority=(1, 3), threshold=1e-11), nthreads=8)
    assert ierr == 0
    assert _rel(scr, ref) < 1e-8, _rel(scr, ref)


@pytest.mark.parametrize("name,fiu,nmix", [("tiny", [0.0, 1.2j], 4), ("tiny", [0.5j, 2.0j], 2)])
def test_solve_linter_selfconsistent_matches_oracle(ctx, name, fiu, nmix):
    """self-consistent W branch (solve_linter.f90:376-460, :564-582) + complex Broyden mixing
    (mix_pot_c.f90:25): same number of self-consistency iterations as the oracle and dV_scf(r, w) equal to 1e-7
    (both sides stop at the same tr2; the per-iteration solver threshold min(0.1 sqrt(dr2), 1e-2) follows dr2)."""
    import oracle
    import synth
    from sternheimergw_b200 import SgwError, select_solver_type
    syn = synth.preset(name)
    ctx.install_system(syn)
    ps = oracle.PwSystem(syn)
    fiu = np.asarray(fiu, dtype=complex)
    nnr = int(np.prod(syn.nr))
    ig = 4
    dv = np.zeros(nnr, dtype=complex)
    dv[syn.nl[ig - 1] - 1] = 1.0
    dvr = (np.fft.ifftn(dv.reshape(syn.nr, order="F")) * nnr).reshape(-1, order="F")
    niter, alpha, tr2 = 40, 0.7, 1e-22
    ref, ierr, st = ps.solve_linter_iter(niter, alpha, tr2, nmix, dvr, fiu, oracle.make_cfg(priority=(1, 3), threshold=1e-4), nthreads=8)
    assert ierr == 0
    ctx.set_mixing(niter, alpha, tr2, nmix)
    out = ctx.solve_linter(select_solver_type(priority=(1, 3), threshold=1e-4), niter, dvr, fiu)
    assert abs(ctx.scf_iterations() - st["iter"]) <= 1, (ctx.scf_iterations(), st["iter"])
    assert _rel(out, ref) < 1e-7, _rel(out, ref)
    # not enough iterations -> the reference's errore (solve_linter.f90:588-591)
    ctx.set_mixing(3, alpha, tr2, nmix)
    with pytest.raises(SgwError, match="Iterative solver did not converge"):
        ctx.solve_linter(select_solver_type(priority=(1, 3), threshold=1e-4), 3, dvr, fiu)
    # without sgw_set_mixing for enough iterations the call is refused, not silently defaulted
    with pytest.raises(SgwError):
        ctx.solve_linter(select_solver_type(priority=(1, 3), threshold=1e-4), 12, dvr, fiu)


def test_coulomb_selfconsistent_equals_inverse_of_direct(ctx):
    """solve_coul = 'iter' (coulomb.f90:104-110,:149-151): the columns returned by the self-consistent branch are
    (eps^-1 - 1) of the FULL direct dielectric matrix -- both computed on the GPU, related by a dense numpy inverse."""
    import synth
    from sternheimergw_b200 import select_solver_type
    syn = synth.preset("tiny")
    ctx.install_system(syn)
    fiu = np.array([0.0, 1.2j])
    ngm = syn.ngm
    igu = np.arange(1, ngm + 1, dtype=np.int32)
    eps = ctx.coulomb(select_solver_type(priority=(1, 3), threshold=1e-12), 1, ngm, ngm, igu, fiu)
    ctx.set_mixing(40, 0.7, 1e-22, 4)
    ctx.set_solve_direct(False)
    try:
        w = ctx.coulomb(select_solver_type(priority=(1, 3), threshold=1e-4), 2, ngm, 3, igu, fiu)
    finally:
        ctx.set_solve_direct(True)
    assert 3 < ctx.scf_iterations() < 40
    for iw in range(fiu.size):
        inv = np.linalg.inv(eps[:, iw, :])
        for t in range(3):
            ref = inv[:, 1 + t].copy()
            ref[1 + t] -= 1.0
            assert np.abs(w[:, iw, t] - ref).max() < 1e-8, (iw, t, np.abs(w[:, iw, t] - ref).max())


def test_coulomb_production_threshold_and_sos(ctx):
    """Production threshold (1e-4): eps within 10*thr of the converged one; converged one equals sum-over-states."""
    import sos
    import synth
    from sternheimergw_b200 import select_solver_type
    syn = synth.preset("tiny")
    ctx.install_system(syn)
    fiu = np.array([0.0, 1.2j])
    ngc = 9
    igu = np.arange(1, ngc + 1, dtype=np.int32)
    tight = ctx.coulomb(select_solver_type(priority=(1, 3), threshold=1e-12), 1, ngc, ngc, igu, fiu)
    loose = ctx.coulomb(select_solver_type(priority=(1, 3), threshold=1e-4), 1, ngc, ngc, igu, fiu)
    assert np.abs(loose - tight).max() < 1e-3
    for ig in (1, 2, 5, 9):
        assert np.abs(tight[:, :, ig - 1] - sos.eps_sos(syn, ig, ngc, fiu)).max() < 1e-9
    # the skipped perturbation rule |q+G|^2 < 1e-8 (coulomb.f90:126): at q = 0 the G = 0 column stays zero
    syn0 = synth.preset("tiny", xq=[0.0, 0.0, 0.0])
    ctx.install_system(syn0)
    scr0 = ctx.coulomb(select_solver_type(priority=(1, 3), threshold=1e-10), 1, 5, 3, igu, fiu)
    assert np.abs(scr0[:, :, 0]).max() == 0.0 and np.abs(scr0[:, :, 1]).max() > 0.0


def test_solve_linter_matches_oracle(ctx):
    """drhoscf(nnr, nfreq) = -dV_H(r) for a generic (non-delta) perturbation, incl. the subspace solver (priority 3)."""
    import oracle
    import synth
    from sternheimergw_b200 import SgwError, select_solver_type
    syn = synth.preset("tiny")
    ctx.install_system(syn)
    ps = oracle.PwSystem(syn)
    rng = np.random.default_rng(4)
    dv = np.zeros(syn.nnr, complex)
    sel = syn.nl[:15] - 1
    dv[sel] = rng.standard_normal(15) + 1j * rng.standard_normal(15)
    dvr = oracle.fft3d(dv.reshape(syn.nr, order="F"), +1).reshape(-1, order="F")
    freq = np.array([0.0, 0.7j, 2.0j])
    for prio in ((1, 3), (3,)):
        thr = 1e-12 if prio[0] == 1 else 1e-10
        out = ctx.solve_linter(select_solver_type(priority=prio, threshold=thr), 1, dvr, freq)
        ref, ierr, _ = ps.solve_linter(dvr, freq, oracle.make_cfg(priority=prio, threshold=thr), nthreads=4)
        assert ierr == 0
        assert out.shape == ref.shape
        assert _rel(out, ref) < 1e-8, (prio, _rel(out, ref))
    with pytest.raises(SgwError):
        ctx.solve_linter(select_solver_type(), 3, dvr, freq)        # num_iter > 1 without sgw_set_mixing for 3 iterations: refused, loud


@pytest.mark.parametrize("ngc,nfs", [(7, 3), (59, 2), (130, 2)])
def test_unfold_and_invert_epsilon(ctx, ngc, nfs):
    import oracle
    rng = np.random.default_rng(2)
    scr_in = rng.standard_normal((ngc, nfs, ngc)) + 1j * rng.standard_normal((ngc, nfs, ngc))
    scr_in += 4 * np.sqrt(ngc) * np.eye(ngc)[:, None, :]
    igu = np.arange(1, ngc + 1, dtype=np.int32)
    full = ctx.unfold_w(ngc, igu, scr_in)
    assert np.array_equal(full, oracle.unfold_w(ngc, nfs, igu, scr_in))
    for lgamma in (False, True):
        inv = ctx.invert_epsilon(full, lgamma=lgamma)
        ref, info = oracle.invert_epsilon(full, lgamma=lgamma)
        assert info == 0
        assert _rel(inv, ref) < 1e-11
    # pivoting is needed: a matrix with a zero leading element
    a = full.copy()
    a[0, 0, :] = 0.0
    inv = ctx.invert_epsilon(a)
    for iw in range(nfs):
        assert np.abs(inv[:, :, iw] + np.eye(ngc) - np.linalg.inv(a[:, :, iw])).max() < 1e-10


@pytest.mark.parametrize("n,nfs,env", [
    (300, 3, {}),                                              # default plan: NB = 64, SB = 16, sub-panels in shared memory
    (333, 2, {"SGW_GJ_NB": "48", "SGW_GJ_SB": "4"}),           # ragged panels and sub-panels
    (150, 2, {"SGW_GJ_PANEL": "global", "SGW_GJ_NB": "32"}),   # sub-panel steps in global memory
    (97, 2, {"SGW_GJ_UNBLOCKED": "1"}),                        # a single panel
    (1100, 2, {}),                                             # SB = 8
])
def test_invert_epsilon_blocked(ctx, monkeypatch, n, nfs, env):
    """The blocked Gauss-Jordan of csrc/invert.cu against LAPACK (numpy), which is what invert_epsilon.f90:59-66 calls;
    rows are scaled so that partial pivoting interchanges rows in almost every step."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    rng = np.random.default_rng(n)
    a = rng.standard_normal((n, n, nfs)) + 1j * rng.standard_normal((n, n, nfs))
    a[np.arange(n), np.arange(n), :] += 2.0 * np.sqrt(n) * (1.0 + np.abs(rng.standard_normal((n, 1))))
    a = np.asfortranarray(a[rng.permutation(n)])
    inv = ctx.invert_epsilon(a)
    for iw in range(nfs):
        ref = np.linalg.inv(a[:, :, iw])
        err = np.abs(inv[:, :, iw] + np.eye(n) - ref).max() / np.abs(ref).max()
        assert err < 1e-11, (iw, err)


def test_invert_epsilon_singular_is_loud(ctx):
    from sternheimergw_b200 import SgwError
    a = np.zeros((4, 4, 1), complex)
    a[0, 0, 0] = 1.0
    with pytest.raises(SgwError):
        ctx.invert_epsilon(a)


@pytest.mark.parametrize("name", ["tiny", "si"])
def test_green_function_matches_oracle(ctx, name):
    """green(ngc, ngp, nfreq): b = -e_{G'}, shifts -omega, strict '<' mask (green.f90:196-213)."""
    import oracle
    import synth
    from sternheimergw_b200 import select_solver_type
    syn = synth.preset(name, nk=1 if name == "si" else 2)
    ctx.install_system(syn)
    ps = oracle.PwSystem(syn)
    kq = syn.kpairs[0].kq
    ngc = 9 if name == "tiny" else 15
    pos = {int(g): i + 1 for i, g in enumerate(kq.igk)}
    map_ = np.array([pos.get(ig, 0) for ig in range(1, ngc + 1)], dtype=np.int32)
    map_[-1] = kq.npw                      # hits the strict '<' of green.f90:212
    map_[3] = 0                            # a G' outside the k sphere: skipped right-hand side
    fft_map = np.arange(1, ngc + 1, dtype=np.int32)[::-1].copy()
    mu = 0.5 * (kq.et[3] + kq.et[4])
    w = np.array([0.3j, 2.0j, 7.0j])
    omega = np.concatenate([mu + w, mu - w])
    cfg = select_solver_type(priority=(1, 3), threshold=1e-12)
    green = ctx.green_function(cfg, 0, map_, fft_map, omega)
    ref, ierr, _ = ps.green_function(0, map_, fft_map, omega, oracle.make_cfg(priority=(1, 3), threshold=1e-12), nthreads=4)
    assert ierr == 0
    assert _rel(green, ref) < 1e-8
    assert np.abs(green[:, list(fft_map).index(4), :]).max() == 0.0


def test_smoke_entry():
    import __graft_entry__ as g
    g.smoke()


def test_unfold_w_symmetric_matches_oracle(ctx):
    """unfold_w with use_symm = .TRUE. (unfold_w.f90:86-131) through sgw_unfold_w_symm against oracle/symm.py (which is pinned by
    the invariance property in tests/test_oracle_symm.py): the full cubic group, with a fractional translation on some
    operations so that the eigv phases (gmap_sym.f90:112-134) take part; and the invariant matrix is recovered on the GPU too."""
    from oracle import symm as osy
    from symm_util import cubic_group, g_shell_list
    from sternheimergw_b200 import SgwError
    ops, invs = cubic_group()
    mill = g_shell_list(6)
    ngc, nsym, nfs = len(mill), len(ops), 3
    rng = np.random.default_rng(4)
    for frac in (False, True):
        ftau = np.zeros((nsym, 3), int)
        if frac:
            ftau[3] = (6, 0, 12); ftau[17] = (0, 12, 6); ftau[40] = (12, 12, 12)
        gmapsym, eigv = osy.gmap_sym(mill, ops, ftau, (24, 24, 24))
        ig_unique, sym_ig, sym_friend = osy.stern_symm(ngc, nsym, gmapsym, invs)
        scr_in = np.asfortranarray(rng.standard_normal((ngc, nfs, ig_unique.size)) + 1j * rng.standard_normal((ngc, nfs, ig_unique.size)))
        ref = osy.unfold_w(ngc, nfs, ig_unique, scr_in, use_symm=True, nsymq=nsym, sym_ig=sym_ig, sym_friend=sym_friend,
                           gmapsym=gmapsym, eigv=eigv, invs=invs)
        got = ctx.unfold_w_symm(ngc, ig_unique, sym_ig, sym_friend, gmapsym, eigv, invs, scr_in)
        assert np.abs(got - ref).max() <= 1e-14 * np.abs(ref).max(), (frac, np.abs(got - ref).max())   # two complex products
        assert np.abs(ref).min() > 0                         # every element of W was filled
    # a table that maps outside the list is refused, not silently dropped
    bad = gmapsym.copy()
    bad[5, :] = 0                                            # G number 6 has no image under any operation
    with pytest.raises(SgwError):
        ctx.unfold_w_symm(ngc, ig_unique, sym_ig, sym_friend, bad, eigv, invs, scr_in)


@pytest.mark.parametrize("nr,env", [
    ((125, 128, 27), {"SGW_RHO_GRID": "fine"}),                       # planes in global memory, also for the Delta-rho accumulation
    ((24, 25, 27), {"SGW_PLANE_GMEM": "1", "SGW_RHO_GRID": "fine"}),
])
def test_coulomb_on_general_grids(ctx, monkeypatch, nr, env):
    """The whole screened-Coulomb column (dV psi, solves, Delta-rho, Hartree) on boxes that take the large-plane path."""
    import oracle
    import synth
    from sternheimergw_b200 import select_solver_type
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    s = synth.build_lattice("tiny", 10.26, synth.FCC, [[0.125] * 3, [-0.125] * 3], ["Si", "Si"], 6.0, nr=nr)
    syn = synth.attach_kpoints(s, synth.mp_grid(s.bg, 1), [0.5, 0.5, 0.5])
    ctx.install_system(syn)
    ps = oracle.PwSystem(syn)
    ngc = 4
    igu = np.arange(1, ngc + 1, dtype=np.int32)
    fiu = synth.imag_freqs(2)
    cfg = select_solver_type(priority=(1, 3), threshold=1e-12)
    scr = ctx.coulomb(cfg, 1, ngc, ngc, igu, fiu)
    ref, ierr, _ = ps.coulomb(1, ngc, ngc, igu, fiu, oracle.make_cfg(priority=(1, 3), threshold=1e-12), nthreads=4)
    assert ierr == 0
    assert _rel(scr, ref) < 1e-8, _rel(scr, ref)


@pytest.mark.parametrize("ngauss,nk", [(0, 1), (-99, 2), (1, 2), (-1, 1)])
def test_coulomb_metal_matches_oracle(ctx, ngauss, nk):
    """Metals (sgw_set_smearing / sgw_set_kpair_metal): the smeared projector of [QE] orthogonalize (lgauss) and the wg/wk scaling
    of solve_linter.f90:373 against the oracle, whose restatement is anchored by tests/test_oracle_metal.py: <= 1e-8 at thr 1e-12."""
    import oracle
    from metal_util import metal_system
    from sternheimergw_b200 import select_solver_type
    syn = metal_system(ngauss=ngauss, nk=nk)
    ctx.install_system(syn)
    ps = oracle.PwSystem(syn)
    ngc = 5
    igu = np.arange(1, ngc + 1, dtype=np.int32)
    fiu = np.array([0.0, 0.9j])
    try:
        scr = ctx.coulomb(select_solver_type(priority=(1, 3), threshold=1e-12), 1, ngc, ngc, igu, fiu)
    finally:
        ctx.set_smearing(False)
    ref, ierr, _ = ps.coulomb(1, ngc, ngc, igu, fiu, oracle.make_cfg(priority=(1, 3), threshold=1e-12), nthreads=4)
    assert ierr == 0
    assert np.abs(ref - np.eye(ngc)[:, None, :]).max() > 1e-3
    assert _rel(scr, ref) < 1e-8, (ngauss, nk, _rel(scr, ref))


def test_padded_box_gives_the_compact_result(ctx):
    """dffts%nr1x > nr1 (padded FFT descriptors): the potential, the index arrays and dvbarein / drhoscf cross the ABI in the
    padded layout; the result is the compact one, entry for entry, and the padding comes back as zeros."""
    import synth
    from sternheimergw_b200 import select_solver_type
    syn = synth.preset("tiny", nk=2)
    nr = tuple(int(x) for x in syn.nr)
    nrx = (nr[0] + 1, nr[1], nr[2] + 2)
    rng = np.random.default_rng(5)
    dv = rng.standard_normal(nr) + 1j * rng.standard_normal(nr)
    freq = np.array([0.0, 0.8j])
    cfg = select_solver_type(priority=(1, 3), threshold=1e-10)
    igu = np.arange(1, 6, dtype=np.int32)
    ctx.install_system(syn)
    a = ctx.solve_linter(cfg, 1, dv.ravel(order="F"), freq).reshape(nr + (2,), order="F")
    ca = ctx.coulomb(cfg, 1, 5, 5, igu, freq)
    ctx.install_system(syn, nrx=nrx)
    dvp = np.zeros(nrx, dtype=complex, order="F")
    dvp[:nr[0], :nr[1], :nr[2]] = dv
    b = ctx.solve_linter(cfg, 1, dvp.ravel(order="F"), freq).reshape(nrx + (2,), order="F")
    cb = ctx.coulomb(cfg, 1, 5, 5, igu, freq)
    assert np.array_equal(b[:nr[0], :nr[1], :nr[2]], a)
    assert np.abs(b[nr[0]:]).max() == 0.0 and np.abs(b[:, :, nr[2]:]).max() == 0.0
    assert np.array_equal(ca, cb)


def test_solve_linter_selfconsistent_metal_matches_oracle(ctx):
    """The self-consistent branch for a metal: both orthogonalize calls (solve_linter.f90:337 and :409) take the smeared
    projector and the solutions of every iteration are scaled by wg/wk (:373); same iteration count and dV_scf as the oracle."""
    import oracle
    from metal_util import metal_system
    from sternheimergw_b200 import select_solver_type
    syn = metal_system(ngauss=0, nk=2)
    ctx.install_system(syn)
    ps = oracle.PwSystem(syn)
    fiu = np.array([0.0, 0.6j])
    nnr = int(np.prod(syn.nr))
    dv = np.zeros(nnr, dtype=complex)
    dv[syn.nl[3] - 1] = 1.0
    dvr = (np.fft.ifftn(dv.reshape(syn.nr, order="F")) * nnr).reshape(-1, order="F")
    niter, alpha, tr2, nmix = 60, 0.5, 1e-22, 4
    try:
        ref, ierr, st = ps.solve_linter_iter(niter, alpha, tr2, nmix, dvr, fiu, oracle.make_cfg(priority=(1, 3), threshold=1e-4), nthreads=8)
        assert ierr == 0
        ctx.set_mixing(niter, alpha, tr2, nmix)
        out = ctx.solve_linter(select_solver_type(priority=(1, 3), threshold=1e-4), niter, dvr, fiu)
    finally:
        ctx.set_smearing(False)
    assert abs(ctx.scf_iterations() - st["iter"]) <= 1, (ctx.scf_iterations(), st["iter"])
    assert _rel(out, ref) < 1e-7, _rel(out, ref)
