"""GPU parity of the batched H.psi / linear_op kernels against the oracle, through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from sternheimergw_b200 import Context
    c = Context(0)
    yield c
    c.close()


def _install_kpoints(ctx, syn):
    ctx.set_grid(*syn.nr)
    ctx.set_vloc(syn.vrs)
    for ik, kp in enumerate(syn.kpairs):
        kq = kp.kq
        ctx.set_kpoint(ik, kq.npw, kq.npwx, kq.nl_igk, kq.g2kin, kq.vkb, kq.dion, kq.evq, kq.alpha_pv)


@pytest.mark.parametrize("name,nk", [("tiny", 2), ("si", 2), ("c", 1), ("licl", 1), ("bn", 1)])
def test_linear_op_matches_oracle(ctx, name, nk):
    """(H + omega S + alpha_pv P_v) psi for a batch of vectors with per-vector omega: <= 1e-12 relative."""
    import oracle
    import synth
    syn = synth.preset(name, nk=nk)
    _install_kpoints(ctx, syn)
    ps = oracle.PwSystem(syn)
    rng = np.random.default_rng(42)
    for ik in range(min(2, len(syn.kpairs))):
        kq = syn.kpairs[ik].kq
        nvec = 7
        psi = np.zeros((kq.npwx, nvec), dtype=complex, order="F")
        psi[:kq.npw] = rng.standard_normal((kq.npw, nvec)) + 1j * rng.standard_normal((kq.npw, nvec))
        omega = rng.standard_normal(nvec) + 1j * rng.standard_normal(nvec)
        for apv in (kq.alpha_pv, 0.0):
            out = ctx.linear_op(ik, omega, apv, psi)
            for v in range(nvec):
                ref = ps.linear_op(ik, omega[v], apv, psi[:, v])
                err = np.abs(out[:, v] - ref).max() / np.abs(ref).max()
                assert err < 1e-12, (name, ik, v, apv, err)
            assert np.abs(out[kq.npw:]).max(initial=0.0) == 0.0


def test_linear_op_is_linear_and_hermitian(ctx):
    """Size-independent properties at a larger batch: linearity and <x|H y> = <H x|y> for real omega."""
    import synth
    syn = synth.preset("si", nk=1)
    _install_kpoints(ctx, syn)
    kq = syn.kpairs[0].kq
    rng = np.random.default_rng(3)
    nvec = 64
    x = np.zeros((kq.npwx, nvec), dtype=complex, order="F")
    y = np.zeros_like(x)
    x[:kq.npw] = rng.standard_normal((kq.npw, nvec)) + 1j * rng.standard_normal((kq.npw, nvec))
    y[:kq.npw] = rng.standard_normal((kq.npw, nvec)) + 1j * rng.standard_normal((kq.npw, nvec))
    om = np.full(nvec, 0.37 + 0j)
    hx = ctx.linear_op(0, om, kq.alpha_pv, x)
    hy = ctx.linear_op(0, om, kq.alpha_pv, y)
    hxy = ctx.linear_op(0, om, kq.alpha_pv, np.asfortranarray(2.0 * x - 1j * y))
    assert np.abs(hxy - (2.0 * hx - 1j * hy)).max() < 1e-11 * np.abs(hx).max()
    a = np.einsum("ij,ij->j", x.conj(), hy)
    b = np.einsum("ij,ij->j", hx.conj(), y)
    assert np.abs(a - b).max() < 1e-10 * np.abs(a).max()
    # eigenvectors: H evq = et evq
    ev = np.asfortranarray(kq.evq)
    hev = ctx.linear_op(0, np.zeros(ev.shape[1], complex), 0.0, ev)
    assert np.abs(hev - ev * kq.et[:ev.shape[1]]).max() < 1e-10


def test_errors_are_loud(ctx):
    from sternheimergw_b200 import SgwError
    with pytest.raises(SgwError):
        ctx.set_grid(13, 13, 13)                   # 13 is not a supported radix product
    with pytest.raises(SgwError):
        ctx.linear_op(999, [0.0], 0.0, np.zeros(10, complex))   # slot not set


def test_caller_owned_stream_gives_identical_results():
    """sgw_set_stream: the kernels run on a caller-owned CUDA stream; same bits as on the context's own stream."""
    import torch
    import synth
    from sternheimergw_b200 import Context
    syn = synth.preset("tiny")
    c = Context(0)
    try:
        c.install_system(syn)
        kq = syn.kpairs[0].kq
        rng = np.random.default_rng(2)
        psi = np.zeros((kq.npwx, 3), complex, order="F")
        psi[:kq.npw] = rng.standard_normal((kq.npw, 3)) + 1j * rng.standard_normal((kq.npw, 3))
        om = np.array([0.1, 0.2j, -0.3])
        ref = c.linear_op(0, om, kq.alpha_pv, psi)
        st = torch.cuda.Stream()
        c.set_stream(st.cuda_stream)
        got = c.linear_op(0, om, kq.alpha_pv, psi)
        c.set_stream(None)
        again = c.linear_op(0, om, kq.alpha_pv, psi)
        assert np.array_equal(got, ref) and np.array_equal(again, ref)
    finally:
        c.close()


def test_persistent_plane_kernel_is_bit_identical_to_the_generic_one(monkeypatch):
    """k_plane_vloc (persistent CTAs, TMA-staged rows, fused scatter/gather, x-column masks; csrc/fft.cu) does the arithmetic
    of the generic k_plane<PLANE_VLOC> in the same order: H.psi at the Si64 size must agree BIT FOR BIT between the generic
    kernel (SGW_PLANE=0), the persistent kernel with a zero-filled plane (1) and the masked variant (2, default), also with
    an `active` mask (multishift solve) -- and the generic kernel is the one test_linear_op_matches_oracle pins to 1e-12."""
    import synth
    from sternheimergw_b200 import Context, select_solver_type
    syn = synth.preset("si64")
    c = Context(0)
    try:
        c.install_system(syn)
        kq = syn.kpairs[0].kq
        rng = np.random.default_rng(5)
        nvec = 37
        psi = np.zeros((kq.npwx, nvec), complex, order="F")
        psi[:kq.npw] = rng.standard_normal((kq.npw, nvec)) + 1j * rng.standard_normal((kq.npw, nvec))
        om = rng.standard_normal(nvec) + 0.3j
        outs = {}
        for variant in ("0", "1", "2"):
            monkeypatch.setenv("SGW_PLANE", variant)
            outs[variant] = c.linear_op(0, om, kq.alpha_pv, psi)
        assert np.array_equal(outs["0"], outs["1"]) and np.array_equal(outs["0"], outs["2"])
        # through the solver (per-vector active masks, converged right-hand sides drop out of the batch)
        b = np.asfortranarray(psi[:, :6] - kq.evq @ (kq.evq.conj().T @ psi[:, :6]))
        b /= np.linalg.norm(b, axis=0)
        sigma = np.asfortranarray(-(kq.et[:6][None, :] + np.array([0.0, 0.2j, -0.2j])[:, None]))
        xs = {}
        for variant in ("0", "2"):
            monkeypatch.setenv("SGW_PLANE", variant)
            xs[variant], ierr = c.select_solver(select_solver_type(priority=(1,), threshold=1e-6), 0, b, sigma)
            assert np.all(ierr == 0)
        assert np.array_equal(xs["0"], xs["2"])
        monkeypatch.delenv("SGW_PLANE")
        # the persistent TMA z passes (bulk-copy staged input, tensor-map stores / loads) against the generic kernels:
        # same butterflies on a transposed tile -> same bits, on the 72^3 box (H.psi) and on the reduced 45^3 Delta-rho box
        zo = {}
        for variant in ("0", "1", "2"):
            monkeypatch.setenv("SGW_ZPASS", variant)
            zo[variant] = c.linear_op(0, om, kq.alpha_pv, psi)
        assert np.array_equal(zo["0"], zo["1"]) and np.array_equal(zo["1"], outs["2"]) and np.array_equal(zo["2"], zo["0"])
        fiu = synth.imag_freqs(3)
        igu = np.arange(1, 41, dtype=np.int32)
        sc = {}
        for variant in ("0", "1"):
            monkeypatch.setenv("SGW_ZPASS", variant)
            sc[variant] = c.coulomb(select_solver_type(priority=(1, 3), threshold=1e-6), 3, 40, 2, igu, fiu)
            assert c.rho_grid()[0]
        assert np.array_equal(sc["0"], sc["1"])
    finally:
        c.close()


def test_rho_plane_v2_is_bit_identical_to_the_generic_kernel(monkeypatch):
    """k_plane_rho_v2 (TMA-staged rows, index-table gather, register accumulator, y-fastest psi_v(r); csrc/fft.cu) against
    the generic k_plane_rho on the reduced 45^3 Delta-rho box of the Si64 workload: the same eps columns bit for bit."""
    import synth
    from sternheimergw_b200 import Context, select_solver_type
    syn = synth.preset("si64")
    c = Context(0)
    try:
        c.install_system(syn)
        fiu = synth.imag_freqs(3)
        igu = np.arange(1, 1901, dtype=np.int32)
        out = {}
        for variant in ("0", "1"):
            monkeypatch.setenv("SGW_RHO_V2", variant)
            out[variant] = c.coulomb(select_solver_type(priority=(1, 3), threshold=1e-6), 5, 1900, 3, igu, fiu)
            assert c.rho_grid()[0] and tuple(c.rho_grid()[1]) == (45, 45, 45)
        assert np.array_equal(out["0"], out["1"]), np.abs(out["0"] - out["1"]).max() / np.abs(out["0"]).max()
    finally:
        c.close()


def _tiny_on_grid(nr, nk=1):
    """The 2-atom 'tiny' stand-in on an FFT box chosen by the test (any box that holds its spheres is valid input)."""
    import synth
    s = synth.build_lattice("tiny", 10.26, synth.FCC, [[0.125] * 3, [-0.125] * 3], ["Si", "Si"], 6.0, nr=nr)
    return synth.attach_kpoints(s, synth.mp_grid(s.bg, nk), [0.5, 0.5, 0.5])


@pytest.mark.parametrize("nr,env", [
    ((125, 128, 27), {}),                       # 125 = 5 x 25; the 125 x 128 plane does not fit in shared memory
    ((162, 20, 200), {}),                       # 162 = 9 x 18 inside a shared-memory plane, 200 = 10 x 20 in the z pass
    ((216, 243, 16), {}),                       # 216 = 12 x 18, 243 = 9 x 27, plane in global memory
    ((24, 25, 27), {"SGW_PLANE_GMEM": "1"}),    # the global-memory plane path on a box the default path also handles
    ((28, 22, 21), {}),                         # factors 7 and 11 (good_fft_order hands such lengths out): radices 7, 11 and 3 x 7
    ((42, 44, 63), {}),
])
def test_linear_op_on_general_grids(ctx, monkeypatch, nr, env):
    """FFT generality (VERDICT r1 missing 4): boxes with lengths that need the radices 18..32 and planes larger than an
    SM's shared memory, against the oracle (whose FFT takes any length): <= 1e-12 relative."""
    import oracle
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    syn = _tiny_on_grid(nr)
    _install_kpoints(ctx, syn)
    ps = oracle.PwSystem(syn)
    rng = np.random.default_rng(7)
    kq = syn.kpairs[0].kq
    nvec = 3
    psi = np.zeros((kq.npwx, nvec), dtype=complex, order="F")
    psi[:kq.npw] = rng.standard_normal((kq.npw, nvec)) + 1j * rng.standard_normal((kq.npw, nvec))
    omega = rng.standard_normal(nvec) + 1j * rng.standard_normal(nvec)
    out = ctx.linear_op(0, omega, kq.alpha_pv, psi)
    for v in range(nvec):
        ref = ps.linear_op(0, omega[v], kq.alpha_pv, psi[:, v])
        err = np.abs(out[:, v] - ref).max() / np.abs(ref).max()
        assert err < 1e-12, (nr, v, err)


def test_unsupported_grid_is_loud(ctx):
    from sternheimergw_b200 import SgwError
    with pytest.raises(SgwError):
        ctx.set_grid(26, 20, 20)                # 2 x 13: no plan
