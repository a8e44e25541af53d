"""The numpy oracle of SURVEY 8 rows f2/f3 (oracle/sigma.py) against independent formulas.  The reference has no unit test
for pade_coeff/pade_eval, godby_needs, fft6 or sigma_prod (algo/analytic/test/pade.pf covers pade_robust only), so
these properties are what anchors the restatement."""
import numpy as np
import pytest

from oracle import sigma as osg


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def test_gauleg_matches_numpy():
    for n in (1, 2, 5, 12, 35):
        x, w = osg.gauleg_grid(0.0, 7.3, n)
        xr, wr = np.polynomial.legendre.leggauss(n)
        assert np.allclose(x, 3.65 + 3.65 * xr, atol=1e-12)
        assert np.allclose(w, 3.65 * wr, atol=1e-12)


def test_freqbins_symm():
    solver = np.array([0.0, 0.3j, 0.9j])
    arr = np.zeros((2, 2, 5), complex)
    arr[:, :, :3] = np.arange(12).reshape(2, 2, 3)
    z = osg.freqbins_symm(solver, osg.EVEN_SYMMETRY, arr)
    assert np.allclose(z, [0, 0.3j, 0.9j, -0.3j, -0.9j])
    assert np.allclose(arr[:, :, 3], arr[:, :, 1]) and np.allclose(arr[:, :, 4], arr[:, :, 2])
    assert np.allclose(osg.freqbins_symm(solver, osg.SQUARE_SYMMETRY), solver ** 2)
    assert np.allclose(osg.freqbins_symm(solver, osg.NO_SYMMETRY), solver)
    z2 = osg.freqbins_symm(np.array([0.2j, 0.5j]), osg.EVEN_SYMMETRY)          # no zero frequency: mesh doubles
    assert np.allclose(z2, [0.2j, 0.5j, -0.2j, -0.5j])
    with pytest.raises(ValueError):
        osg.freqbins_symm(np.array([0.0, 0.0]), osg.EVEN_SYMMETRY)
    f = osg.freqbins_type(solver, np.array([0.1j, 0.2j]), np.ones(2), np.array([0.0j]))
    assert f.num_freq() == 5
    assert np.allclose(f.green(1.0 + 0j), [1 + 0.1j, 1 + 0.2j, 1 - 0.1j, 1 - 0.2j])


def test_pade_interpolates_and_continues():
    """The continued fraction passes through its nodes, and recovers a rational function off the axis."""
    rng = np.random.default_rng(3)
    z = 1j * np.array([0.0, 0.2, 0.5, 0.9, 1.4, 2.0, 2.7, 3.5])
    poles = np.array([1.1, 1.9])                     # 2 pole pairs: representable by the 8-node fraction
    res = rng.standard_normal((4, 3, 2)) + 0.3
    f = lambda w: (res[..., None] * 2 * poles[:, None] / (np.asarray(w) ** 2 - poles[:, None] ** 2)).sum(axis=-2)
    u = f(z)
    a = osg.pade_coeff(z, u)
    for i, zi in enumerate(z):
        assert _rel(osg.pade_eval(z, a, zi), u[..., i]) < 1e-10
    w = 0.7 + 0.4j
    assert _rel(osg.pade_eval(z, a, w), f([w])[..., 0]) < 1e-6
    # scalar transliteration of pade.f90 against the vectorised form
    N = z.size
    g = np.zeros((N, N), complex)
    uu = u[1, 2]
    prot = lambda x: x if abs(x) > 1e-24 else 1e-24 + 0j
    for p in range(N):
        for i in range(p, N):
            g[p, i] = prot(uu[i]) if p == 0 else prot((g[p - 1, p - 1] / g[p - 1, i] - g[p - 1, i] / g[p - 1, i]) / (z[i] - z[p - 1]))
    assert np.array_equal(np.diag(g), a[1, 2])


def test_godby_needs_reproduces_inputs():
    rng = np.random.default_rng(5)
    wp = 1.3
    w0 = -(rng.random((5, 5)) + 0.5)
    w1 = w0 * (0.2 + 0.5 * rng.random((5, 5)))       # same sign, smaller magnitude: a valid plasmon-pole pair
    w0[0, 1] = w1[0, 1] = 0.3                        # W(0) = W(i wp): coefficient set to zero (godby_needs.f90:62)
    c = np.stack([w0, w1], axis=2).astype(complex)
    osg.godby_needs_coeffs(wp, c)
    m0 = osg.godby_needs_model(0.0j, c)
    m1 = osg.godby_needs_model(1j * wp, c)
    mask = np.ones((5, 5), bool)
    mask[0, 1] = False
    assert _rel(m0[mask], w0[mask]) < 1e-12 and _rel(m1[mask], w1[mask]) < 1e-12
    assert m0[0, 1] == 0 and c[0, 1, 0] == 0 and c[0, 1, 1] == 0


def _grid(nr, ngc, seed=0):
    rng = np.random.default_rng(seed)
    nnr = int(np.prod(nr))
    nl = np.sort(rng.choice(nnr, ngc, replace=False)) + 1
    return osg.corr_fft_type(tuple(nr), nl.astype(np.int32))


def test_fft6_matches_numpy_6d():
    d = _grid((3, 4, 5), 17)
    rng = np.random.default_rng(1)
    nnr, ng = d.nnr, d.ngm
    omega = 2.7
    fg = rng.standard_normal((ng, ng)) + 1j * rng.standard_normal((ng, ng))
    f = np.zeros((nnr, nnr), complex)
    f[:ng, :ng] = fg
    osg.invfft6(f, d, d, omega)
    # independent: f(r,r') = 1/omega sum_{G,G'} e^{-iGr} f(G,G') e^{+iG'r'}  as one 6-D transform
    box = np.zeros((nnr, nnr), complex)
    box[np.ix_(d.nl - 1, d.nl - 1)] = fg
    b6 = box.reshape(d.nr + d.nr, order="F")
    r6 = np.fft.fftn(b6, axes=(0, 1, 2))
    r6 = np.fft.ifftn(r6, axes=(3, 4, 5)) * nnr / omega
    assert _rel(f, r6.reshape(nnr, nnr, order="F")) < 1e-12
    osg.fwfft6(f, d, d, omega)
    assert _rel(f[:ng, :ng], fg) < 1e-12            # round trip


def test_sigma_prod_is_a_convolution():
    """alpha * fwfft6(G(r,r') W(r,r')) = alpha/omega sum G(G1,G1') W(G-G1, G'-G1') on the (aliased) box."""
    nr = (3, 3, 2)
    d = osg.corr_fft_type(nr, np.arange(1, 19, dtype=np.int32))       # every box point is a G vector: no pruning
    nnr = d.nnr
    rng = np.random.default_rng(2)
    G = rng.standard_normal((nnr, nnr)) + 1j * rng.standard_normal((nnr, nnr))
    W = rng.standard_normal((nnr, nnr)) + 1j * rng.standard_normal((nnr, nnr))
    omega, alpha = 1.9, 0.3 - 0.2j
    Gr = G.copy()
    osg.invfft6(Gr, d, d, omega)
    work = W.copy()
    osg.sigma_prod(omega, d, d, alpha, Gr, work)
    idx = np.array([[i, j, k] for k in range(nr[2]) for j in range(nr[1]) for i in range(nr[0])])
    lin = lambda m: (m[..., 0] % nr[0]) + nr[0] * ((m[..., 1] % nr[1]) + nr[1] * (m[..., 2] % nr[2]))
    ref = np.zeros((nnr, nnr), complex)
    for a in range(nnr):
        da = lin(idx[a][None, :] - idx)              # G - G1 for all G1
        for b in range(nnr):
            db = lin(idx[b][None, :] - idx)
            ref[a, b] = (G * W[np.ix_(da, db)]).sum()
    assert _rel(work, alpha / omega * ref) < 1e-12


def test_qp_eigval_linearisation():
    w = np.linspace(-2, 2, 41)
    sig = 0.1 - 0.25 * w
    e, z = osg.qp_eigval(w, sig, 0.33)
    assert abs(z - 1 / 1.25) < 1e-12 and abs(e - (0.33 + z * (0.1 - 0.25 * 0.33))) < 1e-12
    assert osg.qp_eigval(w, sig, 5.0) == (5.0, 1.0)


def test_aaa_recovers_a_rational_function():
    """vendor/analytic/src/aaa.f90 restated: degree-3 rational data are fitted with 4 support points, interpolated at the
    support points and continued off the mesh; the packed coefficient layout of analytic.f90:160-167 round-trips."""
    z = 1j * 0.07 * np.arange(14) * (np.arange(14) + 1)
    poles = np.array([0.9 + 0.3j, -1.7 + 0.2j, 0.4 - 2.9j])
    res = np.array([1.0, -0.4 + 0.2j, 0.3])
    f = lambda w: (res / (np.atleast_1d(w)[:, None] - poles)).sum(-1)
    p, v, w = osg.aaa_generate(1e-10, z.size // 3, z, f(z))
    assert p.size == 4 and np.all(np.diff(np.abs(p)) > 0)                  # mesh order (PACK)
    assert _rel(osg.aaa_evaluate(p, v, w, z), f(z)) < 1e-12
    assert _rel(osg.aaa_evaluate(p, v, w, [0.3 + 0.8j, -1.0j]), f([0.3 + 0.8j, -1.0j])) < 1e-11
    assert np.array_equal(osg.aaa_evaluate(p, v, w, p), v)                # tabulated value at a support point
    fo = osg.freqbins_type(z, np.array([0.1j]), np.ones(1), np.array([0j]), osg.NO_SYMMETRY)
    scr = np.zeros((2, 2, z.size), complex, order="F")
    scr[:, :, :] = f(z)[None, None, :] * np.array([[1.0, 2.0], [0.5j, -1.0]])[:, :, None]
    data = scr.copy()
    osg.analytic_coeff(osg.AAA_APPROX, 1e-10, fo, scr)
    mmax = z.size // 3
    assert np.count_nonzero(np.abs(scr[0, 1, 2 * mmax:]) > 1e-12) == 4 and np.all(scr[:, :, 3 * mmax:] == 0)
    got = osg.analytic_eval(osg.AAA_APPROX, np.array([1, 2]), fo, scr, z[5])
    assert _rel(got, data[:, :, 5]) < 1e-12


def test_aaa_pole_recovers_poles_and_residues():
    """'aaa pole' restated (aaa.f90 aaa_pole_residual via ZGGEV, analytic.f90 pole_correction / aaa_pole_eval): the three poles
    and residues of a rational function come back, the stored sum reproduces the function."""
    z = 1j * 0.07 * np.arange(14) * (np.arange(14) + 1)
    poles = np.array([0.9 + 0.3j, -1.7 + 0.2j, 0.4 - 2.9j])
    res = np.array([1.0, -0.4 + 0.2j, 0.3])
    f = lambda w: (res / (np.atleast_1d(w)[:, None] - poles)).sum(-1)
    p, v, w = osg.aaa_generate(1e-10, z.size, z, f(z))
    pl, rs = osg.aaa_pole_residual(p, v, w)
    o, o0 = np.argsort(pl.real), np.argsort(poles.real)
    assert np.abs(pl[o] - poles[o0]).max() < 1e-9 and np.abs(rs[o] - res[o0]).max() < 1e-8
    c = osg.pole_correction(1e-8, p, v, w, z.size)
    assert np.count_nonzero(c[z.size // 2:]) == 3
    assert abs(osg.aaa_pole_eval(0.3 + 0.8j, c) - f(0.3 + 0.8j)[0]) < 1e-8
    with pytest.raises(ValueError):                    # analytic.f90:362: more relevant poles than half the array can hold
        osg.pole_correction(1e-8, p, v, w, 4)


def test_pade_robust_reproduces_the_reference_golden_numbers():
    """algo/analytic/test/pade.pf:120-200 -- the reference's own known-answer test of pade_robust (exp on the unit circle,
    degrees (4, 4); cos on the circle of radius 2, degrees (5, 11) reduced to (4, 10)), same numbers, same thresholds.
    This is the one golden vector the reference holds in algo/analytic; it pins the restated pade_robust.f90."""
    circle = lambda radius, n: radius * np.exp(2j * np.pi * np.arange(n) / n)       # pade_problem_evaluate
    dn, dd, cn, cd = osg.pade_robust(1.0, np.exp(circle(1.0, 25)), 4, 4)
    assert (dn, dd, cn.size, cd.size) == (4, 4, 5, 5)
    for got, want in zip(cn, [1.000000000000000, 0.499999999987559, 0.107142857136564, 0.011904761903482, 5.952380951286663e-4]):
        assert abs(got - want) < 1e-10
    for got, want in zip(cd, [1.000000000000000, -0.500000000012441, 0.107142857149005, -0.011904761905969, 5.952380953354424e-4]):
        assert abs(got - want) < 1e-10
    dn, dd, cn, cd = osg.pade_robust(2.0, np.cos(circle(2.0, 25)), 5, 11, 1e-10, 1e-15)
    assert (dn, dd, cn.size, cd.size) == (4, 10, 5, 11)
    for got, want in zip(cn, [1.0, -5.736142091971364e-19, -0.450639141234688, 2.325721027857943e-19, 0.018381449098078]):
        assert abs(got - want) < 1e-8
    for got, want in zip(cd, [1.0, -5.736142091971364e-19, 0.049360858765312, -5.423500181277404e-20, 0.001395211814067,
                              -3.222944976341087e-21, 2.979234736799848e-05, -1.468353270738426e-22, 5.175090814695919e-07,
                              -3.717349899090172e-24, 6.546464226759871e-09]):
        assert abs(got - want) < 1e-8
    # the coefficient layout of pade_coeff_robust and its evaluation
    z = circle(1.5, 24)
    f = np.zeros((1, 2, 24), complex)
    f[0, 0, :] = np.exp(z)
    g = lambda w: (1.0 + w + w * w) / (w - 3.0)       # [2/1]: (a numerator of degree <= 1 is undefined in the reference, :386)
    f[0, 1, :] = g(z)
    osg.pade_coeff_robust(z, f)
    for w in (0.3 + 0.2j, -0.7j):
        assert abs(osg.pade_eval_robust(f[0, 0, :], w) - np.exp(w)) < 1e-9
        assert abs(osg.pade_eval_robust(f[0, 1, :], w) - g(w)) < 1e-5      # FFT aliasing (r / 3)^24 of the Taylor coefficients
    assert (int(f[0, 1, 0].real), int(f[0, 1, 1].real)) == (2, 1)      # the robust algorithm finds the true degrees
