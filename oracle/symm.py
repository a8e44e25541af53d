"""CPU restatement (numpy, reference loop order) of the symmetry reduction of the G-perturbations and of the unfolding of W --
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

  gmap_sym    algo/symmetry/src/gmap_sym.f90:25-150     S(G) index table and the phases e^{-i G.v} of fractional translations
  stern_symm  algo/symmetry/src/stern_symm.f90:23-106   the unique G list ig_unique, sym_ig, sym_friend
  unfold_w    algo/symmetry/src/unfold_w.f90:23-131     full W(G, G', w) from the symmetry-reduced rows

`s` holds the rotation matrices in crystal (Miller) coordinates as QE's s(3, 3, nsym); everything is 1-based where the
reference is (index arrays), 0 marks "not found".  Pinned by an independent property (tests/test_oracle_symm.py): for a
W(G, G') that is invariant under the group, unfolding the unique rows gives back the full matrix.
"""
import numpy as np


def gmap_sym(mill, s, ftau, nr, num_g_corr=None):
    """mill: (ngm, 3) Miller indices of the global G list; s: (nsym, 3, 3) with s[isym][a, b] = s(a+1, b+1, isym+1);
    ftau: (nsym, 3) fractional translations in FFT-grid units; nr: (nr1, nr2, nr3).  Returns gmapsym (num_g_corr, nsym)
    1-based (0 = not in the list) and eigv (num_g_corr, nsym)."""
    mill = np.asarray(mill, dtype=np.int64)
    ngm = mill.shape[0]
    ngc = ngm if num_g_corr is None else num_g_corr
    nsym = len(s)
    index = {tuple(m): i + 1 for i, m in reversed(list(enumerate(map(tuple, mill))))}     # first occurrence, like the DO WHILE scan
    gmapsym = np.zeros((ngc, nsym), dtype=np.int32)
    eigv = np.ones((ngc, nsym), dtype=np.complex128)
    for isym in range(nsym):
        for ig in range(ngc):
            rot = tuple(int(x) for x in np.asarray(s[isym], dtype=np.int64) @ mill[ig])      # gmap_sym.f90:92-94
            gmapsym[ig, isym] = index.get(rot, 0)
            if np.any(np.asarray(ftau[isym]) != 0):                                          # :112-134
                rdotk = sum(float(mill[ig, d] * ftau[isym][d]) / float(nr[d]) for d in range(3))
                eigv[ig, isym] = np.exp(-1j * 6.28318530717959 * rdotk)                      # the reference's single-precision-looking twopi
    return gmapsym, eigv


def stern_symm(num_g_corr, nsymq, gmapsym, invs):
    """stern_symm.f90:81-103.  invs: 1-based inverse-operation table.  Returns (ig_unique[:ngmunique], sym_ig, sym_friend), 1-based."""
    ig_unique = [1]
    sym_ig = np.zeros(num_g_corr, dtype=np.int32)
    sym_friend = np.zeros(num_g_corr, dtype=np.int32)
    for ig in range(2, num_g_corr + 1):
        unique = True
        for isym in range(1, nsymq + 1):
            for igp in range(len(ig_unique)):
                if ig == gmapsym[ig_unique[igp] - 1, invs[isym - 1] - 1]:
                    unique = False
                    sym_ig[ig - 1] = isym                  # no EXIT in the reference: the LAST match stays
                    sym_friend[ig - 1] = ig_unique[igp]
        if unique:
            ig_unique.append(ig)
    return np.asarray(ig_unique, dtype=np.int32), sym_ig, sym_friend


def unfold_w(num_g_corr, nfs, ig_unique, scrcoul_in, use_symm=False, nsymq=1, sym_ig=None, sym_friend=None, gmapsym=None,
             eigv=None, invs=None):
    """unfold_w.f90:84-131.  scrcoul_in(num_g_corr, nfs, ngmunique) -> scrcoul_out(num_g_corr, num_g_corr, nfs)."""
    out = np.zeros((num_g_corr, num_g_corr, nfs), dtype=np.complex128, order="F")
    for ig, iu in enumerate(ig_unique):
        out[iu - 1, :, :] = np.conj(scrcoul_in[:, :, ig])                                   # :86-88
    if not use_symm or nsymq == 1:                                                          # :91-93
        return out
    done = set(int(i) for i in ig_unique)
    for ig in range(1, num_g_corr + 1):                                                     # :104-129
        if ig in done:
            continue
        done.add(ig)
        fr, isym = int(sym_friend[ig - 1]), int(sym_ig[ig - 1])
        tmp = out[fr - 1, :, :].copy()                                                      # scrcoul_tmp(igp, iwim)
        ism1 = int(invs[isym - 1])
        for iw in range(nfs):
            for igp in range(1, num_g_corr + 1):
                phase = eigv[fr - 1, isym - 1] * np.conj(eigv[igp - 1, isym - 1])           # :123
                out[ig - 1, gmapsym[igp - 1, ism1 - 1] - 1, iw] = tmp[igp - 1, iw] * phase  # :124
    return out
