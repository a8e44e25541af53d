"""GPU parity of the batched multishift BiCGStab(l) / SGW subspace solver / select_solver chain.

The first half replays the reference's own unit test (algo/linear_solver/test/linear_solver.pf) on its golden
vector through the CUDA path (dense fake backend, linear_solver.pf:106); the second half compares the
plane-wave solves with the oracle on seeded synthetic inputs.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from sternheimergw_b200 import Context
    c = Context(0)
    yield c
    c.close()


def _res(A, s, x, b):
    return np.linalg.norm(A @ x + s * x - b)


def test_fixture_bicgstab_multishift_all_lmax(ctx, lin_prob):
    """linear_solver.pf:228-252 for lmax = 1..15, plus parity with the oracle (same exit iteration, 1e-8)."""
    import oracle
    from sternheimergw_b200 import select_solver_type
    A, b, sigma = lin_prob["A"], lin_prob["b"], lin_prob["sigma"]
    ctx.set_dense_operator(0, A)
    thr = 1e-6
    for lmax in range(1, 16):
        x, ierr = ctx.select_solver(select_solver_type(priority=(1,), threshold=thr, bicg_lmax=lmax), 0, b, sigma)
        st = ctx.stats()
        assert ierr == 0
        assert not np.isnan(x).any()
        for i, s in enumerate(sigma):
            assert _res(A, s, x[:, i], b) <= 10 * thr
        xo, ierr_o, so = oracle.bicgstab_dense(A, b, sigma, lmax=lmax, threshold=thr)
        assert st["n_outer_max"] == so["n_outer"], (lmax, st, so)
        assert st["n_linear_op"] == so["n_op"]
        assert np.abs(x - xo).max() < 1e-8 * np.abs(xo).max()


def test_fixture_bicgstab_single_shifts_batched(ctx, lin_prob):
    """linear_solver.pf:196-221: every shift alone -- here as 70 independent right-hand sides in ONE batch."""
    from sternheimergw_b200 import select_solver_type
    A, b, sigma = lin_prob["A"], lin_prob["b"], lin_prob["sigma"]
    ctx.set_dense_operator(0, A)
    thr = 1e-6
    bb = np.asfortranarray(np.repeat(b.reshape(-1, 1), sigma.size, axis=1))
    for lmax in (1, 2, 4, 7, 15):
        x, ierr = ctx.select_solver(select_solver_type(priority=(1,), threshold=thr, bicg_lmax=lmax), 0, bb,
                                    sigma.reshape(1, -1))
        assert (ierr == 0).all()
        for i, s in enumerate(sigma):
            assert _res(A, s, x[:, 0, i], b) <= thr


def test_fixture_known_answers(ctx, lin_prob):
    import oracle
    from sternheimergw_b200 import select_solver_type
    A, b, sigma = lin_prob["A"], lin_prob["b"], lin_prob["sigma"]
    ctx.set_dense_operator(0, A)
    x, ierr = ctx.select_solver(select_solver_type(priority=(1,), threshold=1e-6, bicg_lmax=4), 0, b, sigma)
    st = ctx.stats()
    assert (ierr, st["n_linear_op"], st["n_outer_max"]) == (0, 24, 3)
    x, ierr = ctx.select_solver(select_solver_type(priority=(1,), threshold=1e-12, bicg_lmax=4), 0, b, sigma)
    assert ctx.stats()["n_linear_op"] == 32
    xd = np.stack([np.linalg.solve(A + s * np.eye(283), b) for s in sigma], axis=1)
    assert np.max(np.linalg.norm(x - xd, axis=0) / np.linalg.norm(xd, axis=0)) < 1e-12


@pytest.mark.parametrize("path", ["cholqr", "mgs"])
def test_fixture_subspace_solver(ctx, lin_prob, monkeypatch, path):
    """linear_solver.pf:265-345 (default config: relative threshold 1e-4) + parity with the oracle, for the default path
    (Cholesky-QR re-orthonormalisation, incremental residual) and for the kernels that keep the reference's MGS order."""
    import oracle
    from sternheimergw_b200 import select_solver_type
    if path == "mgs":
        monkeypatch.setenv("SGW_SUB", "mgs")
    A, b, sigma = lin_prob["A"], lin_prob["b"], lin_prob["sigma"]
    ctx.set_dense_operator(0, A)
    x, ierr = ctx.select_solver(select_solver_type(priority=(3,), threshold=1e-4), 0, b, sigma)
    st = ctx.stats()
    assert ierr == 0
    nb = np.linalg.norm(b)
    for i, s in enumerate(sigma):
        assert _res(A, s, x[:, i], b) / nb <= 1e-4
    xo, ierr_o, so = oracle.linear_solver_dense(A, b, sigma, threshold=1e-4)
    assert st["n_linear_op"] == so["n_op"]
    assert np.abs(x - xo).max() < 1e-8 * np.abs(xo).max()
    # tight threshold: more basis vectors than the initial capacity (exercises the growth path)
    x, ierr = ctx.select_solver(select_solver_type(priority=(3,), threshold=1e-11), 0, b, sigma)
    assert ierr == 0
    xd = np.stack([np.linalg.solve(A + s * np.eye(283), b) for s in sigma], axis=1)
    assert np.max(np.linalg.norm(x - xd, axis=0) / np.linalg.norm(xd, axis=0)) < 1e-9


def test_select_solver_error_codes_and_fallback(ctx, lin_prob):
    """select_solver.f90:121-159: ierr=1 when max_iter is hit, fall through to the next solver in the list."""
    from sternheimergw_b200 import SgwError, select_solver_type
    A, b, sigma = lin_prob["A"], lin_prob["b"], lin_prob["sigma"][:6]
    ctx.set_dense_operator(0, A)
    x, ierr = ctx.select_solver(select_solver_type(priority=(1,), threshold=1e-10, max_iter=1), 0, b, sigma)
    assert ierr == 1
    assert np.abs(x).max() > 0          # the reference still copies the unconverged x out (bicgstab.f90:258)
    x, ierr = ctx.select_solver(select_solver_type(priority=(1, 3), threshold=1e-10, max_iter=40), 0, b, sigma)
    assert ierr == 0
    msgs = []
    ctx.set_message_callback(msgs.append)        # the reference's stdout warnings arrive through the callback
    x, ierr = ctx.select_solver(select_solver_type(priority=(1, 3), threshold=1e-10, max_iter=2), 0, b, sigma)
    st = ctx.stats()
    ctx.set_message_callback(None)
    assert ierr == 1 and st["n_fallback"] == 1
    assert len(msgs) == 3 and msgs[0].startswith("WARNING: BiCGstab algorithm did not converge in 2 iterations.")   # bicgstab.f90:250
    assert msgs[1] == "First choice of solver did not converge, try a different one"                                 # select_solver.f90:126
    assert msgs[2].startswith("WARNING: SternheimerGW linear solver did not converge")                              # linear_solver.f90:177
    x2, ierr2 = ctx.select_solver(select_solver_type(priority=(2,), threshold=1e-8), 0, b, sigma)
    assert ierr2 == 0
    for i, s in enumerate(sigma):
        assert _res(A, s, x2[:, i], b) <= 1e-7
    with pytest.raises(SgwError):
        ctx.select_solver(select_solver_type(priority=()), 0, b, sigma)
    # NaN in the right-hand side -> ierr = 2 (bicgstab.f90:264-267)
    bn = b.copy()
    bn[3] = np.nan
    _, ierr = ctx.select_solver(select_solver_type(priority=(1,), threshold=1e-8, max_iter=3), 0, bn, sigma)
    assert ierr == 2


def test_mixed_batch_freezes_converged_rhs(ctx, lin_prob):
    """Right-hand sides converge at different outer iterations; each must equal its stand-alone solve."""
    from sternheimergw_b200 import select_solver_type
    A, b, sigma = lin_prob["A"], lin_prob["b"], lin_prob["sigma"]
    ctx.set_dense_operator(0, A)
    rng = np.random.default_rng(5)
    nrhs = 9
    bb = np.asfortranarray(rng.standard_normal((283, nrhs)) + 1j * rng.standard_normal((283, nrhs)))
    bb[:, 0] = b
    bb[:, 4] = 1e-9 * bb[:, 4]            # converged before the first iteration ends up below the threshold early
    sg = np.asfortranarray(np.stack([sigma[i:i + 5] for i in range(nrhs)], axis=1))
    cfg = select_solver_type(priority=(1,), threshold=1e-7, bicg_lmax=4)
    xb, ierr = ctx.select_solver(cfg, 0, bb, sg)
    assert (ierr == 0).all()
    for r in range(nrhs):
        xs, ie = ctx.select_solver(cfg, 0, bb[:, r], sg[:, r])
        assert ie == 0
        assert np.array_equal(xs, xb[:, :, r])


@pytest.mark.parametrize("name", ["tiny", "si"])
def test_planewave_solves_match_oracle(ctx, name):
    """Converged dpsi (thr 1e-12): <= 1e-8 relative; production threshold: same outer-iteration count +-1."""
    import oracle
    import synth
    from sternheimergw_b200 import select_solver_type
    syn = synth.preset(name, nk=1 if name == "si" else 2)
    ctx.install_system(syn) if False else None
    ctx.set_grid(*syn.nr)
    ctx.set_vloc(syn.vrs)
    kq = syn.kpairs[0].kq
    ctx.set_kpoint(0, kq.npw, kq.npwx, kq.nl_igk, kq.g2kin, kq.vkb, kq.dion, kq.evq, kq.alpha_pv)
    ps = oracle.PwSystem(syn)
    rng = np.random.default_rng(9)
    nrhs = 4
    bb = np.zeros((kq.npwx, nrhs), dtype=complex, order="F")
    raw = rng.standard_normal((kq.npw, nrhs)) + 1j * rng.standard_normal((kq.npw, nrhs))
    ev = kq.evq[:kq.npw]
    bb[:kq.npw] = -(raw - ev @ (ev.conj().T @ raw))            # -P_c b, like orthogonalize
    freq = synth.imag_freqs(4)
    omega = np.concatenate([freq, -freq[1:]])
    sg = np.asfortranarray(np.stack([-(syn.kpairs[0].et[r] + omega) for r in range(nrhs)], axis=1))
    for thr, tol in ((1e-12, 1e-8), (1e-4, None)):
        for use_pv in (True, False):
            if not use_pv:
                sgm = np.asfortranarray(sg + 0.3 + 0.2j)         # green-like: no projector, shifted off the spectrum
            else:
                sgm = sg
            x, ierr = ctx.select_solver(select_solver_type(priority=(1, 3), threshold=thr), 0, bb, sgm, use_alpha_pv=use_pv)
            st = ctx.stats()
            assert (ierr == 0).all()
            nouter = 0
            for r in range(nrhs):
                xo, ie, so = ps.select_solver(0, bb[:kq.npw, r], sgm[:, r],
                                              oracle.make_cfg(priority=(1, 3), threshold=thr),
                                              alpha_pv=None if use_pv else 0.0)
                assert ie == 0
                nouter = max(nouter, so["n_outer"])
                err = np.abs(x[:kq.npw, :, r] - xo).max() / np.abs(xo).max()
                assert err < (tol if tol else 10 * thr * 50), (thr, use_pv, r, err)
            assert abs(st["n_outer_max"] - nouter) <= 1


@pytest.mark.parametrize("path", ["cholqr", "mgs"])
def test_planewave_subspace_many_shifts(ctx, monkeypatch, path):
    """SGW subspace solver (priority 3) on the gw_licl stand-in with a frequency mesh like its 102 shifts: the basis is
    carried from shift to shift and re-orthonormalised at each one (linear_solver.f90:272-300).  Both device paths against
    the oracle: <= 1e-8 at threshold 1e-11, same operator count at the production threshold."""
    import oracle
    import synth
    from sternheimergw_b200 import select_solver_type
    if path == "mgs":
        monkeypatch.setenv("SGW_SUB", "mgs")
    syn = synth.preset("licl", nk=1)
    ctx.install_system(syn)
    kq = syn.kpairs[0].kq
    ps = oracle.PwSystem(syn)
    rng = np.random.default_rng(11)
    nrhs = 3
    bb = np.zeros((kq.npwx, nrhs), dtype=complex, order="F")
    raw = rng.standard_normal((kq.npw, nrhs)) + 1j * rng.standard_normal((kq.npw, nrhs))
    ev = kq.evq[:kq.npw]
    bb[:kq.npw] = -(raw - ev @ (ev.conj().T @ raw))
    freq = (np.linspace(2.5, 12.5, 12) + 0.3j) / 13.605698066
    omega = np.concatenate([freq, -freq])
    sg = np.asfortranarray(np.stack([-(syn.kpairs[0].et[r] + omega) for r in range(nrhs)], axis=1))
    for thr, tol in ((1e-11, 1e-8), (1e-4, None)):
        x, ierr = ctx.select_solver(select_solver_type(priority=(3,), threshold=thr), 0, bb, sg)
        nop = ctx.stats()["n_linear_op"]
        assert (ierr == 0).all()
        nop_o = 0
        for r in range(nrhs):
            xo, ie, so = ps.select_solver(0, bb[:kq.npw, r], sg[:, r], oracle.make_cfg(priority=(3,), threshold=thr))
            assert ie == 0
            nop_o += so["n_op"]
            err = np.abs(x[:kq.npw, :, r] - xo).max() / np.abs(xo).max()
            assert err < (tol if tol else 10 * thr * 50), (thr, r, err)
        assert abs(nop - nop_o) <= (0 if tol else 2), (nop, nop_o)
