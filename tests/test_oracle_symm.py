"""oracle/symm.py (gmap_sym.f90, stern_symm.f90, unfold_w.f90) anchored by an independent property: a W(G, G', w) that is
invariant under the small group of q, W(SG, SG') = W(G, G'), is recovered exactly from its symmetry-unique rows."""
import numpy as np

from oracle import symm as osy
from symm_util import cubic_group, g_shell_list


def _invariant_w(mill, nfs, seed=0):
    """W(G, G', w) = f_w(|G|^2, |G'|^2, G.G') with an asymmetric dependence on (G, G') and complex values."""
    g2 = (mill ** 2).sum(axis=1).astype(float)
    dot = (mill @ mill.T).astype(float)
    w = np.zeros((len(mill), len(mill), nfs), dtype=complex, order="F")
    for iw in range(nfs):
        w[:, :, iw] = np.exp(-0.1 * (iw + 1) * g2[:, None]) * (1.0 + 0.3 * g2[None, :]) + 1j * np.sin(0.37 * dot + 0.2 * iw) \
            + 0.05 * dot * g2[:, None]
    return w


def test_unfold_recovers_an_invariant_matrix():
    ops, invs = cubic_group()
    mill = g_shell_list(6)                      # 81 G vectors, closed under the group
    ngc, nsym, nfs = len(mill), len(ops), 3
    gmapsym, eigv = osy.gmap_sym(mill, ops, np.zeros((nsym, 3), int), (24, 24, 24))
    assert gmapsym.min() >= 1 and np.all(eigv == 1.0)
    ig_unique, sym_ig, sym_friend = osy.stern_symm(ngc, nsym, gmapsym, invs)
    assert 1 < ig_unique.size < ngc // 4         # the group really reduces the list (one representative per star)
    full = _invariant_w(mill, nfs)
    # what `coulomb` delivers for the unique perturbations: scrcoul_in(igp, iw, ig) = CONJG(W(ig_unique(ig), igp, iw))
    scr_in = np.zeros((ngc, nfs, ig_unique.size), dtype=complex, order="F")
    for i, iu in enumerate(ig_unique):
        scr_in[:, :, i] = np.conj(full[iu - 1, :, :])
    out = osy.unfold_w(ngc, nfs, ig_unique, scr_in, use_symm=True, nsymq=nsym, sym_ig=sym_ig, sym_friend=sym_friend,
                       gmapsym=gmapsym, eigv=eigv, invs=invs)
    assert np.abs(out - full).max() < 1e-13
    # identity-symmetry branch (use_symm = .FALSE.): only the unique rows are filled
    out0 = osy.unfold_w(ngc, nfs, ig_unique, scr_in)
    rest = np.setdiff1d(np.arange(1, ngc + 1), ig_unique) - 1
    assert np.abs(out0[ig_unique - 1] - full[ig_unique - 1]).max() == 0.0 and np.abs(out0[rest]).max() == 0.0


def test_gmap_sym_phases_of_fractional_translations():
    """eigv(ig, isym) = exp(-i 2 pi (m . ftau / nr)) (gmap_sym.f90:112-134), 1 for symmorphic operations."""
    ops, invs = cubic_group()
    mill = g_shell_list(3)
    ftau = np.zeros((len(ops), 3), int)
    ftau[5] = (6, 0, 12)
    gm, eigv = osy.gmap_sym(mill, ops, ftau, (24, 24, 48))
    want = np.exp(-2j * np.pi * (mill[:, 0] * 6 / 24 + mill[:, 2] * 12 / 48))
    assert np.abs(eigv[:, 5] - want).max() < 1e-12 and np.all(eigv[:, 4] == 1.0)
