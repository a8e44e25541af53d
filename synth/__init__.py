"""Synthetic plane-wave inputs for the Sternheimer hot path.

Quantum ESPRESSO cannot run here, so the inputs a Fortran host would hand to the C-ABI
(FFT grid, local potential ``vrs``, G lists and ``nl`` maps, ``g2kin``, ``vkb``/``dion``, eigenpairs
``evc``/``evq``/``et``, ``alpha_pv``) are generated directly, in QE's conventions:

* lattice vectors ``at`` in units of ``alat``; reciprocal ``bg`` in units of 2pi/alat; ``tpiba2=(2pi/alat)^2``
* global G list sorted by |G|^2 (G=0 first) inside the density sphere ``ecutrho = 4 ecutwfc``
* ``nl(ig)`` = 1-based column-major index of the wrapped Miller indices in the (nr1,nr2,nr3) box
* k-point spheres ``|k+G|^2 tpiba2 <= ecutwfc`` sorted by kinetic energy; ``npwx`` = max over k-points
* ``invfft`` = unscaled sum_G f(G) e^{+iGr};  ``fwfft`` = (1/nnr) sum_r f(r) e^{-iGr}

The wavefunctions are exact eigenvectors (dense ``eigh``) of the very operator the oracle and the CUDA
library apply -- the local potential is *defined on the FFT grid* and its G-space matrix elements are
taken from ``fwfft(vrs)`` with the box's own aliasing, so H_dense == H_fft to rounding.

This module is harness code (tests + bench); it is not part of the product path.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

SEED = 20261017
RYTOEV = 13.605698066


# ----------------------------------------------------------------------------- helpers
def good_fft_order(n: int) -> int:
    """Smallest m >= n whose only prime factors are 2, 3, 5 (QE good_fft_order without 7/11)."""
    m = max(int(n), 1)
    while True:
        k = m
        for p in (2, 3, 5):
            while k % p == 0:
                k //= p
        if k == 1:
            return m
        m += 1


def _form_factor(x):
    """Smooth Cohen-Bergstresser-like local form factor v(x) in Ry, x = |G|^2 in (2pi/a_Si)^2 units.

    Passes through the Si EPM values v(3)=-0.21, v(8)=+0.04, v(11)=+0.08 and decays to 0.
    """
    xs = np.array([0.0, 1.0, 2.0, 3.0, 4.0, 6.0, 8.0, 11.0, 14.0, 17.0, 20.0, 24.0, 1e9])
    vs = np.array([-1.10, -0.80, -0.47, -0.21, -0.11, -0.01, 0.04, 0.08, 0.05, 0.02, 0.005, 0.0, 0.0])
    return np.interp(x, xs, vs)


# species -> (reference lattice constant a_s [bohr], amplitude) of the synthetic local pseudopotential
SPECIES_VLOC = {"Si": (10.26, 1.0), "C": (6.74, 1.0), "B": (6.83, 0.55), "N": (6.83, 1.45),
                "Li": (9.693, 0.25), "Cl": (9.693, 1.75)}


@dataclass
class KQ:
    """Operator data at one k-point (what init_us_2 + g2_kin + evq/alpha_pv globals hold)."""
    xk: np.ndarray
    npw: int
    npwx: int
    igk: np.ndarray        # (npw,) 1-based index into the global G list
    mill: np.ndarray       # (npw,3)
    nl_igk: np.ndarray     # (npw,) int32, 1-based FFT index
    g2kin: np.ndarray      # (npw,)
    vkb: np.ndarray        # (npwx, nkb) complex, zero padded
    dion: np.ndarray       # (nkb, nkb) real
    evq: np.ndarray        # (npwx, nbnd_occ)
    et: np.ndarray         # (nbnd,) all computed eigenvalues (Ry)
    alpha_pv: float = 0.0


@dataclass
class KPairS:
    kq: KQ
    npw_k: int
    nl_igk_k: np.ndarray
    evc: np.ndarray        # (npwx, nbnd)
    et: np.ndarray         # (nbnd,)
    wk: float
    k: KQ | None = None


@dataclass
class SynthSystem:
    name: str
    alat: float
    at: np.ndarray
    bg: np.ndarray
    nr: tuple
    ecutwfc: float
    tpiba2: float
    omega_cell: float
    tau: np.ndarray         # (nat,3) cartesian, alat units
    species: list
    vrs: np.ndarray         # (nnr,) real, F-order flattening of (nr1,nr2,nr3)
    ngm: int
    mill: np.ndarray        # (ngm,3)
    g: np.ndarray           # (3,ngm) cartesian, 2pi/alat
    gg: np.ndarray          # (ngm,)
    nl: np.ndarray          # (ngm,) int32 1-based
    nbnd_occ: int
    xq: np.ndarray = field(default_factory=lambda: np.zeros(3))
    kpairs: list = field(default_factory=list)
    npwx: int = 0
    alpha_pv: float = 0.0
    proj: dict = field(default_factory=dict)

    @property
    def nnr(self):
        return int(np.prod(self.nr))


# ----------------------------------------------------------------------------- lattice / G vectors
def _nl_of_mill(mill, nr):
    m = np.mod(mill, np.asarray(nr)[None, :])
    return (1 + m[:, 0] + nr[0] * (m[:, 1] + nr[1] * m[:, 2])).astype(np.int32)


def build_lattice(name, alat, at, tau, species, ecutwfc, nr=None, nbnd_occ=4, vloc_scale=None, seed=SEED):
    at = np.asarray(at, dtype=float)
    bg = np.linalg.inv(at).T                      # bg[i] . at[j] = delta_ij   (units 2pi/alat)
    tpiba2 = (2.0 * np.pi / alat) ** 2
    omega_cell = abs(np.linalg.det(at)) * alat ** 3
    gcutm = 4.0 * ecutwfc / tpiba2                # density sphere |G|^2 (2pi/alat)^2
    mmax = [int(np.floor(np.sqrt(gcutm) * np.linalg.norm(at[i]) + 1e-9)) for i in range(3)]
    if nr is None:
        nr = tuple(good_fft_order(2 * m + 1) for m in mmax)
    nr = tuple(int(x) for x in nr)
    rng = [np.arange(-m, m + 1) for m in mmax]
    M = np.stack(np.meshgrid(*rng, indexing="ij"), axis=-1).reshape(-1, 3)
    G = M @ bg
    gg = np.einsum("ij,ij->i", G, G)
    keep = gg <= gcutm + 1e-10
    M, G, gg = M[keep], G[keep], gg[keep]
    # QE sorts by |G|^2; ties broken deterministically here (Miller index lexicographic)
    order = np.lexsort((M[:, 2], M[:, 1], M[:, 0], np.round(gg, 8)))
    M, G, gg = M[order], G[order], gg[order]
    assert np.all(M[0] == 0)
    for i in range(3):
        assert 2 * np.abs(M[:, i]).max() + 1 <= nr[i], (name, nr, np.abs(M).max(0))
    nl = _nl_of_mill(M, nr)

    # local potential on the density sphere -> real-space grid.  Each species carries a length scale a_s
    # (the lattice constant of its diamond/zincblende-like reference crystal) and an amplitude:
    #   V(G) = sum_atoms e^{-iG.tau} * amp_s (a_ref/a_s)^2 v(|G|^2 (a_s/2pi)^2) * (a_s^3/8) / omega_cell
    tau = np.asarray(tau, dtype=float)
    a_ref = 10.26
    if vloc_scale is None:
        vloc_scale = {}
    Vg = np.zeros(len(gg), dtype=complex)
    for t, s in zip(tau, species):
        a_s, amp = SPECIES_VLOC[s]
        amp = amp * vloc_scale.get(s, 1.0)
        x = gg * tpiba2 / (2.0 * np.pi / a_s) ** 2
        phase = np.exp(-2j * np.pi * (G @ t))
        Vg += amp * (a_ref / a_s) ** 2 * _form_factor(x) * phase * (a_s ** 3 / 8.0) / omega_cell
    Vg[0] = 0.0
    box = np.zeros(nr, dtype=complex)
    mm = np.mod(M, np.asarray(nr)[None, :])
    box[mm[:, 0], mm[:, 1], mm[:, 2]] = Vg
    vr = np.fft.ifftn(box) * box.size                          # invfft
    assert np.abs(vr.imag).max() < 1e-10 * max(1.0, np.abs(vr.real).max())
    vrs = np.ascontiguousarray(vr.real.reshape(-1, order="F"))
    return SynthSystem(name=name, alat=alat, at=at, bg=bg, nr=nr, ecutwfc=ecutwfc, tpiba2=tpiba2,
                       omega_cell=omega_cell, tau=tau, species=list(species), vrs=vrs, ngm=len(gg),
                       mill=M, g=np.ascontiguousarray(G.T), gg=gg, nl=nl, nbnd_occ=nbnd_occ)


# ----------------------------------------------------------------------------- k-point operator data
# projector channels per species: list of (l, D_l [Ry], r_c [bohr])
DEFAULT_PROJ = {"Si": [(0, 0.45, 1.1), (1, -0.25, 1.1)], "C": [(0, 0.60, 0.8)], "B": [(0, 0.40, 0.9)],
                "N": [(0, 0.70, 0.8)], "Li": [(0, 0.30, 1.4)], "Cl": [(0, 0.50, 1.0), (1, -0.30, 1.0)]}


def kpoint_sphere(sys: SynthSystem, xk):
    xk = np.asarray(xk, dtype=float)
    kg = sys.g.T + xk[None, :]
    q2 = np.einsum("ij,ij->i", kg, kg)
    sel = np.flatnonzero(q2 * sys.tpiba2 <= sys.ecutwfc + 1e-10)
    order = np.argsort(q2[sel], kind="stable")
    sel = sel[order]
    return sel, q2[sel] * sys.tpiba2


def build_vkb(sys: SynthSystem, xk, igk, proj=None):
    """vkb(G, beta) = f_l(|k+G|) Y_lm(k+G) (-i)^l e^{-i(k+G).tau}  (init_us_2 semantics, synthetic radial part)."""
    proj = proj or sys.proj or DEFAULT_PROJ
    kg = (sys.g.T[igk] + np.asarray(xk)[None, :])             # 2pi/alat
    tpiba = np.sqrt(sys.tpiba2)
    q = kg * tpiba                                            # bohr^-1
    qn = np.linalg.norm(q, axis=1)
    cols, dvals = [], []
    for t, s in zip(sys.tau, sys.species):
        sk = np.exp(-2j * np.pi * (kg @ t))
        for (l, D, rc) in proj.get(s, []):
            rad = np.exp(-0.25 * (qn * rc) ** 2)
            # <k+G|beta> of a unit-norm gaussian projector exp(-r^2/rc^2): cell-size independent norm
            norm = (2.0 * np.pi * rc ** 2) ** 0.75 / np.sqrt(sys.omega_cell)
            if l == 0:
                cols.append(norm * rad * sk)
                dvals.append(D)
            elif l == 1:
                for m in range(3):
                    cols.append((-1j) * norm * rc * q[:, m] * rad * sk)
                    dvals.append(D)
            else:
                raise NotImplementedError
    if not cols:
        return np.zeros((len(igk), 0), dtype=complex), np.zeros((0, 0))
    return np.stack(cols, axis=1), np.diag(np.asarray(dvals, dtype=float))


def dense_h(sys: SynthSystem, mill, g2kin, vkb, dion):
    """Dense H in the plane-wave basis, with V_loc matrix elements taken from the FFT box (aliasing included)."""
    nr = np.asarray(sys.nr)
    box = sys.vrs.reshape(sys.nr, order="F")
    Vg = np.fft.fftn(box) / box.size                          # fwfft
    d = mill[:, None, :] - mill[None, :, :]
    d = np.mod(d, nr[None, None, :])
    H = Vg[d[..., 0], d[..., 1], d[..., 2]]
    H = H + np.diag(g2kin)
    if vkb.shape[1]:
        H = H + vkb @ dion @ vkb.conj().T
    return 0.5 * (H + H.conj().T)


def make_kq(sys: SynthSystem, xk, nbnd, npwx=None) -> KQ:
    igk, g2kin = kpoint_sphere(sys, xk)
    mill = sys.mill[igk]
    vkb, dion = build_vkb(sys, xk, igk)
    H = dense_h(sys, mill, g2kin, vkb, dion)
    et, ev = np.linalg.eigh(H)
    # fix the gauge deterministically (largest component real positive) so fixtures are reproducible
    for b in range(nbnd):
        j = np.argmax(np.abs(ev[:, b]))
        ev[:, b] *= np.exp(-1j * np.angle(ev[:, b][j]))
    npw = len(igk)
    return KQ(xk=np.asarray(xk, float), npw=npw, npwx=npwx or npw, igk=(igk + 1).astype(np.int32), mill=mill,
              nl_igk=sys.nl[igk].astype(np.int32), g2kin=g2kin, vkb=vkb, dion=dion,
              evq=ev[:, :nbnd].copy(), et=et[:max(nbnd, min(len(et), nbnd + 4))].copy())


def make_kq_blocks(sys: SynthSystem, xk, nbnd_blk, key_fn, nextra=1) -> KQ:
    """Like make_kq for a supercell whose potential has the periodicity of a smaller primitive cell.

    H couples two plane waves only if their Miller indices differ by a primitive reciprocal-lattice vector,
    so H is block diagonal over the cosets ``key_fn(mill)`` (one block per folded k-point of the primitive
    cell).  Each block is diagonalised densely; the ``nbnd_blk`` lowest states of every block together are the
    occupied manifold (checked: the highest of them lies below the lowest unoccupied state of every block).
    """
    igk, g2kin = kpoint_sphere(sys, xk)
    mill = sys.mill[igk]
    vkb, dion = build_vkb(sys, xk, igk)
    keys = key_fn(mill)
    npw = len(igk)
    cols, ets, e_unocc = [], [], []
    for key in np.unique(keys):
        idx = np.flatnonzero(keys == key)
        H = dense_h(sys, mill[idx], g2kin[idx], vkb[idx], dion)
        e, v = np.linalg.eigh(H)
        for b in range(nbnd_blk):
            j = np.argmax(np.abs(v[:, b]))
            col = np.zeros(npw, dtype=complex)
            col[idx] = v[:, b] * np.exp(-1j * np.angle(v[j, b]))
            cols.append(col)
            ets.append(e[b])
        e_unocc.append(e[nbnd_blk:nbnd_blk + nextra])
    ets = np.asarray(ets)
    order = np.argsort(ets, kind="stable")
    evq = np.stack([cols[i] for i in order], axis=1)
    e_un = np.sort(np.concatenate(e_unocc))
    assert ets.max() < e_un[0], ("occupied manifold is not the union of the block manifolds", ets.max(), e_un[0])
    et = np.concatenate([ets[order], e_un[:4]])
    return KQ(xk=np.asarray(xk, float), npw=npw, npwx=npw, igk=(igk + 1).astype(np.int32), mill=mill,
              nl_igk=sys.nl[igk].astype(np.int32), g2kin=g2kin, vkb=vkb, dion=dion, evq=evq, et=et)


def attach_kpoints_blocks(sys: SynthSystem, klist, xq, nbnd_blk, key_fn, weights=None):
    """attach_kpoints for block-diagonal supercells (see make_kq_blocks)."""
    xq = np.asarray(xq, dtype=float)
    klist = [np.asarray(k, dtype=float) for k in klist]
    nk = len(klist)
    weights = weights if weights is not None else [2.0 / nk] * nk
    ks = [make_kq_blocks(sys, k, nbnd_blk, key_fn) for k in klist]
    kqs = [make_kq_blocks(sys, k + xq, nbnd_blk, key_fn) for k in klist]
    nbnd = ks[0].evq.shape[1]
    npwx = max(max(k.npw for k in ks), max(k.npw for k in kqs))
    emax_occ = max(k.et[nbnd - 1] for k in ks + kqs)
    emin = min(k.et[0] for k in ks + kqs)
    alpha_pv = max(2.0 * (emax_occ - emin), 1e-2)
    sys.kpairs = []
    for k, kq, w in zip(ks, kqs, weights):
        for o in (k, kq):
            o.npwx = npwx
            o.vkb = _pad(o.vkb, npwx)
            o.evq = _pad(o.evq, npwx)
            o.alpha_pv = alpha_pv
        sys.kpairs.append(KPairS(kq=kq, npw_k=k.npw, nl_igk_k=k.nl_igk, evc=k.evq, et=k.et[:nbnd].copy(), wk=w, k=k))
    sys.npwx, sys.alpha_pv, sys.xq, sys.nbnd_occ = npwx, alpha_pv, xq, nbnd
    sys.gap = min(k.et[nbnd] for k in ks + kqs) - emax_occ
    return sys


def si_supercell(ncell=2, ecutwfc=None, nr=None, xq=None, name=None):
    """Diamond Si in a simple-cubic supercell of ncell^3 conventional cells (8 ncell^3 atoms); ncell = 2 is the
    64-atom cell of BASELINE.json's scaling config (72^3 grid, ~24 k plane waves, 128 occupied bands, nkb 256)."""
    a0 = 10.26
    alat = a0 * ncell
    fcc = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0]])
    tau = []
    for i in range(ncell):
        for j in range(ncell):
            for k in range(ncell):
                for f in fcc:
                    for sgn in (+1, -1):
                        tau.append((np.array([i, j, k]) + f + sgn * 0.125) / ncell)
    tau = np.array(tau)
    if ecutwfc is None:
        # radius-(9 ncell - 0.1) sphere in units of 2pi/alat: fills the (36 ncell)^3 box the way 16 Ry fills 20^3
        ecutwfc = (9.0 * ncell - 0.1) ** 2 * (2.0 * np.pi / alat) ** 2
    s = build_lattice(name or f"si{8 * ncell ** 3}", alat, np.eye(3), tau, ["Si"] * len(tau), ecutwfc, nr=nr,
                      nbnd_occ=4 * 4 * ncell ** 3)
    m = 2 * ncell    # primitive (fcc) reciprocal lattice in supercell Miller units: all = 0 or all = ncell (mod 2 ncell)

    def key_fn(mill):
        r = np.mod(mill, m)
        flip = r[:, 0] >= ncell
        r[flip] = np.mod(r[flip] + ncell, m)
        return (r[:, 0] * m + r[:, 1]) * m + r[:, 2]

    q = [0.25, 0.25, 0.25] if xq is None else xq
    return attach_kpoints_blocks(s, [np.zeros(3)], q, 4, key_fn)


def _pad(a, npwx):
    out = np.zeros((npwx,) + a.shape[1:], dtype=a.dtype, order="F")
    out[:a.shape[0]] = a
    return out


def attach_kpoints(sys: SynthSystem, klist, xq, nbnd=None, weights=None):
    """Build (k, k+q) pairs like setup_nscf/initialize_gw do (ikks/ikqs), with wk summing to 2."""
    nbnd = nbnd or sys.nbnd_occ
    xq = np.asarray(xq, dtype=float)
    klist = [np.asarray(k, dtype=float) for k in klist]
    nk = len(klist)
    weights = weights if weights is not None else [2.0 / nk] * nk
    ks = [make_kq(sys, k, nbnd) for k in klist]
    kqs = [make_kq(sys, k + xq, nbnd) for k in klist]
    npwx = max(max(k.npw for k in ks), max(k.npw for k in kqs))
    emax_occ = max(max(k.et[nbnd - 1] for k in ks), max(k.et[nbnd - 1] for k in kqs))
    emin = min(min(k.et[0] for k in ks), min(k.et[0] for k in kqs))
    gap = min(min(k.et[nbnd] for k in ks + kqs) - emax_occ, 1e9) if all(len(k.et) > nbnd for k in ks + kqs) else np.nan
    alpha_pv = max(2.0 * (emax_occ - emin), 1e-2)              # setup_alpha_pv (insulator)
    sys.kpairs = []
    for k, kq, w in zip(ks, kqs, weights):
        for o in (k, kq):
            o.npwx = npwx
            o.vkb = _pad(o.vkb, npwx)
            o.evq = _pad(o.evq, npwx)
            o.alpha_pv = alpha_pv
        sys.kpairs.append(KPairS(kq=kq, npw_k=k.npw, nl_igk_k=k.nl_igk, evc=k.evq, et=k.et[:nbnd].copy(), wk=w, k=k))
    sys.npwx, sys.alpha_pv, sys.xq = npwx, alpha_pv, xq
    sys.gap = gap
    return sys


@dataclass
class MetalPair:
    evq_all: np.ndarray     # (npwx, nbnd) all bands at k+q
    et_q: np.ndarray        # (nbnd,)
    nocc_k: int             # nbnd_occ(ikk)
    wg_over_wk: np.ndarray  # (nocc_k,) wg(ibnd, ikk) / wk(ikk)


@dataclass
class Metal:
    ef: float
    degauss: float
    ngauss: int
    pairs: list


def wgauss(x, n):
    """[QE] Modules/wgauss.f90 for ngauss = 0 (Gaussian) and -99 (Fermi-Dirac); the others live in the oracle and the library."""
    from math import erfc, exp
    if n == -99:
        return 0.0 if x < -200 else (1.0 if x > 200 else 1.0 / (1.0 + exp(-x)))
    assert n == 0
    return 0.5 * erfc(-x)


def make_metal(sys: SynthSystem, ef, degauss, ngauss=0, occ_fn=None, target=None):
    """Turn a system built by attach_kpoints(sys, klist, xq, nbnd=nbnd_all) into a metallic one (klist lgauss): the Fermi level
    `ef` lies inside the computed bands, nbnd_occ(ik) = number of bands below ef + target * degauss (setup_nbnd_occ), the
    operator's projector keeps the first nbnd_occ(ikq) bands of evq, and `sys.metal` carries what orthogonalize's lgauss
    branch and solve_linter.f90:373 read.  occ_fn(x, ngauss) = theta~(x); default: the oracle's orc_wgauss."""
    if occ_fn is None:
        import oracle
        occ_fn = oracle.wgauss
    if target is None:
        # [QE] setup_nbnd_occ: bands up to ef + xmax * degauss, xmax = 3 (w0gauss < 6.96e-5 for a Gaussian), 9.57 for Fermi-Dirac
        target = 9.57 if ngauss == -99 else 3.0
    pairs = []
    for kp in sys.kpairs:
        kq = kp.kq
        nb = kq.evq.shape[1]
        et_q = np.asarray(kq.et[:nb], dtype=float)
        et_k = np.asarray(kp.et[:nb], dtype=float)
        nocc_q = max(1, int(np.sum(et_q < ef + target * degauss)))
        nocc_k = max(1, int(np.sum(et_k < ef + target * degauss)))
        assert nocc_q < nb and nocc_k < nb, "raise nbnd: every computed band is (partially) occupied"
        pairs.append(MetalPair(evq_all=kq.evq.copy(), et_q=et_q.copy(), nocc_k=nocc_k,
                               wg_over_wk=np.array([occ_fn((ef - e) / degauss, ngauss) for e in et_k[:nocc_k]])))
        kq.evq = np.ascontiguousarray(kq.evq[:, :nocc_q])
    sys.metal = Metal(ef=float(ef), degauss=float(degauss), ngauss=int(ngauss), pairs=pairs)
    return sys


def mp_grid(bg, n):
    """Unshifted n1 x n2 x n3 Monkhorst-Pack grid in cartesian 2pi/alat units (no symmetry reduction)."""
    n = (n, n, n) if np.isscalar(n) else n
    ks = []
    for i in range(n[0]):
        for j in range(n[1]):
            for k in range(n[2]):
                ks.append((i / n[0]) * bg[0] + (j / n[1]) * bg[1] + (k / n[2]) * bg[2])
    return ks


# ----------------------------------------------------------------------------- presets
FCC = 0.5 * np.array([[-1.0, 0.0, 1.0], [0.0, 1.0, 1.0], [-1.0, 1.0, 0.0]])


def preset(name: str, xq=None, nk=2):
    """Synthetic stand-ins for the BASELINE.json configs: 'tiny', 'si' (gw_si), 'c' (gw_c), 'bn', 'licl'."""
    if name == "tiny":          # 2-atom fcc, very low cutoff: fast CPU tests
        s = build_lattice("tiny", 10.26, FCC, [[0.125] * 3, [-0.125] * 3], ["Si", "Si"], 6.0)
        q = [0.5, 0.5, 0.5] if xq is None else xq
        return attach_kpoints(s, mp_grid(s.bg, 1) if nk == 1 else mp_grid(s.bg, nk), q)
    if name == "si":
        s = build_lattice("si", 10.26, FCC, [[0.125] * 3, [-0.125] * 3], ["Si", "Si"], 16.0)
        q = [-0.5, 0.5, -0.5] if xq is None else xq          # L point, one of gw_si's q
        return attach_kpoints(s, mp_grid(s.bg, nk), q)
    if name == "c":
        s = build_lattice("c", 6.74, FCC, [[0.125] * 3, [-0.125] * 3], ["C", "C"], 27.0)
        q = [-0.5, 0.5, -0.5] if xq is None else xq
        return attach_kpoints(s, mp_grid(s.bg, nk), q)
    if name == "licl":
        s = build_lattice("licl", 9.693, FCC, [[0.0] * 3, [0.5, 0.5, 0.5]], ["Li", "Cl"], 27.0)
        q = [-0.5, 0.5, -0.5] if xq is None else xq
        return attach_kpoints(s, mp_grid(s.bg, nk), q)
    if name == "bn":
        c_a = 17.008 / 4.748
        at = np.array([[1.0, 0.0, 0.0], [-0.5, np.sqrt(3) / 2, 0.0], [0.0, 0.0, c_a]])
        tau = np.array([[0.0, 0.0, 0.0], [0.5, 1.0 / (2 * np.sqrt(3)), 0.0]])
        s = build_lattice("bn", 4.748, at, tau, ["B", "N"], 30.0)
        q = [0.2, 0.0, 0.0] if xq is None else xq
        ks = [(i / 5) * s.bg[0] + (j / 5) * s.bg[1] for i in range(5) for j in range(5)] if nk >= 5 \
            else [(i / nk) * s.bg[0] + (j / nk) * s.bg[1] for i in range(nk) for j in range(nk)]
        return attach_kpoints(s, ks, q)
    if name == "si64":          # BASELINE.json configs[4]: 64-atom Si supercell, 72^3 grid, 1 q / 1 k
        return si_supercell(2, xq=xq)
    if name == "si8":           # same construction, one conventional cell (36^3 grid): CPU-sized parity case
        return si_supercell(1, xq=xq)
    raise KeyError(name)


def imag_freqs(n):
    """gw_c/gw.in imaginary grid: i * 0.15 n (n+1) eV -> Ry."""
    return np.array([1j * 0.15 * k * (k + 1) for k in range(n)]) / RYTOEV


def corr_grid(sys: SynthSystem, ngc: int, nr=None):
    """The custom FFT type grid%corr_fft of the correlation cutoff (algo/grid/src/sigma_grid.f90): the smallest good
    FFT box that holds the first `ngc` G vectors of the global list (or the box `nr` given), and their 1-based
    positions `nl` in it."""
    mill = sys.mill[:ngc]
    if nr is None:
        nr = tuple(good_fft_order(2 * int(np.abs(mill[:, i]).max()) + 1) for i in range(3))
    assert all(nr[i] >= 2 * int(np.abs(mill[:, i]).max()) + 1 for i in range(3)), "box too small for these G vectors"
    return tuple(nr), _nl_of_mill(mill, np.asarray(nr))
