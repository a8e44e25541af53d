"""Independent numpy formulas used to anchor the (reference-unpinned) plane-wave half of the oracle.

Sum-over-states evaluation of the direct-solver dielectric matrix: the Sternheimer solution of
    (H + sigma + alpha_pv P_v) dpsi = -P_c dV psi_v ,  sigma = -(e_v + w)
is dpsi(w) = - sum_c |c><c|dV|v> / (e_c - e_v - w); the reference averages +w and -w
(solve_linter.f90:464-480), accumulates drho = sum_k (2 wk/Omega) sum_v conj(psi_v(r)) dpsi_v(r)
(incdrhoscf), applies the Hartree kernel (dv_of_drho, lrpa) and forms eps = delta - dV_H (coulomb.f90:143-157).
Here everything is done with a full dense eigh at k+q and numpy FFTs -- no iterative solver involved.
"""
import numpy as np

import synth


def _to_box(sys, nl, coef):
    box = np.zeros(sys.nnr, dtype=complex)
    box[nl - 1] = coef
    return box.reshape(sys.nr, order="F")


def eps_sos(sys: "synth.SynthSystem", ig_pert: int, ngc: int, freqs):
    """eps_{G', G}(q, w) for the perturbation G = ig_pert (1-based), rows G' = 1..ngc; returns (ngc, nfreq)."""
    nnr = sys.nnr
    nocc = sys.nbnd_occ
    dv = np.zeros(nnr, dtype=complex)
    dv[sys.nl[ig_pert - 1] - 1] = 1.0
    dvr = np.fft.ifftn(dv.reshape(sys.nr, order="F")) * nnr
    freqs = np.asarray(freqs, dtype=complex)
    drho = np.zeros((len(freqs),) + tuple(sys.nr), dtype=complex)
    for kp in sys.kpairs:
        kq, k = kp.kq, kp.k
        H = synth.dense_h(sys, kq.mill, kq.g2kin, kq.vkb[:kq.npw], kq.dion)
        e, U = np.linalg.eigh(H)
        Uc, ec = U[:, nocc:], e[nocc:]
        for v in range(nocc):
            psir = np.fft.ifftn(_to_box(sys, kp.nl_igk_k, kp.evc[:kp.npw_k, v])) * nnr
            prod = np.fft.fftn(psir * dvr) / nnr
            dvpsi = prod.reshape(-1, order="F")[kq.nl_igk - 1]
            c = Uc.conj().T @ dvpsi
            for iw, w in enumerate(freqs):
                den = 0.5 * (1.0 / (ec - kp.et[v] - w) + 1.0 / (ec - kp.et[v] + w))
                dpsi = -Uc @ (c * den)
                dpsir = np.fft.ifftn(_to_box(sys, kq.nl_igk, dpsi)) * nnr
                drho[iw] += (2.0 * kp.wk / sys.omega_cell) * np.conj(psir) * dpsir
    out = np.zeros((ngc, len(freqs)), dtype=complex)
    qg2 = np.sum((sys.g + sys.xq[:, None]) ** 2, axis=0)
    for iw in range(len(freqs)):
        dg = (np.fft.fftn(drho[iw]) / nnr).reshape(-1, order="F")
        for igp in range(ngc):
            vc = 2.0 * 4.0 * np.pi / (sys.tpiba2 * qg2[igp]) if qg2[igp] > 1e-8 else 0.0
            out[igp, iw] = -vc * dg[sys.nl[igp] - 1]
        if ig_pert <= ngc:
            out[ig_pert - 1, iw] += 1.0
    return out


def eps_sos_smeared(sys: "synth.SynthSystem", ig_pert: int, ngc: int, occ, docc, tol=1e-7):
    """Static eps_{G', G}(q, 0) from FINITE-TEMPERATURE first-order perturbation theory with the full double sum over states,
        drho(r) = sum_k (wk / Omega) sum_{i at k, j at k+q} F_ij conj(psi_i(r)) psi_j(r) <j|dV|i>,
        F_ij = (f_i - f_j) / (e_i - e_j)   (-> f'(e_i) for degenerate pairs),   f = occ(e), f' = docc(e),
    over ALL eigenstates at k and k+q (dense eigh of both).  No Sternheimer equation, no projector, no pair splitting: this is what
    de Gironcoli's smeared Sternheimer scheme (and its factor 2 for the time-reversed partner, incdrhoscf's weight 2 wk) must
    reproduce on a k mesh that contains -k-q for every k.  Returns (ngc,)."""
    nnr = sys.nnr
    dv = np.zeros(nnr, dtype=complex)
    dv[sys.nl[ig_pert - 1] - 1] = 1.0
    dvr = np.fft.ifftn(dv.reshape(sys.nr, order="F")) * nnr
    drho = np.zeros(tuple(sys.nr), dtype=complex)
    for kp in sys.kpairs:
        kq, k = kp.kq, kp.k
        eq, Uq = np.linalg.eigh(synth.dense_h(sys, kq.mill, kq.g2kin, kq.vkb[:kq.npw], kq.dion))
        ek, Uk = np.linalg.eigh(synth.dense_h(sys, k.mill, k.g2kin, k.vkb[:k.npw], k.dion))
        fk, fq = np.array([occ(e) for e in ek]), np.array([occ(e) for e in eq])
        psi_q_r = [np.fft.ifftn(_to_box(sys, kq.nl_igk, Uq[:, j])) * nnr for j in range(len(eq))]
        for i in range(len(ek)):
            de = ek[i] - eq
            F = np.where(np.abs(de) > tol, (fk[i] - fq) / np.where(np.abs(de) > tol, de, 1.0), docc(ek[i]))
            if np.abs(F).max() < 1e-14:
                continue
            psir = np.fft.ifftn(_to_box(sys, k.nl_igk, Uk[:, i])) * nnr
            prod = np.fft.fftn(psir * dvr) / nnr
            m = Uq.conj().T @ prod.reshape(-1, order="F")[kq.nl_igk - 1]          # <j|dV|i>
            for j in np.flatnonzero(np.abs(F) > 1e-14):
                drho += (kp.wk / sys.omega_cell) * F[j] * m[j] * np.conj(psir) * psi_q_r[j]
    out = np.zeros(ngc, dtype=complex)
    qg2 = np.sum((sys.g + sys.xq[:, None]) ** 2, axis=0)
    dg = (np.fft.fftn(drho) / nnr).reshape(-1, order="F")
    for igp in range(ngc):
        vc = 2.0 * 4.0 * np.pi / (sys.tpiba2 * qg2[igp]) if qg2[igp] > 1e-8 else 0.0
        out[igp] = -vc * dg[sys.nl[igp] - 1]
    if ig_pert <= ngc:
        out[ig_pert - 1] += 1.0
    return out
