"""ctypes front-end of the CPU oracle (oracle/solver.c, oracle/pw.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  The product (sternheimergw_b200/) never imports this package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

c_int, c_double, c_void_p, c_long = C.c_int, C.c_double, C.c_void_p, C.c_long


class SolverCfg(C.Structure):
    """select_solver_type (select_solver.f90:48-62)."""
    _fields_ = [("npriority", c_int), ("priority", c_int * 4), ("max_iter", c_int),
                ("threshold", c_double), ("bicg_lmax", c_int)]


class Stats(C.Structure):
    _fields_ = [("n_op", c_long), ("n_outer", c_int), ("solver_used", c_int)]


class ZC(C.Structure):
    """double _Complex passed by value (SysV: two SSE eightbytes, same as struct{double,double})."""
    _fields_ = [("re", c_double), ("im", c_double)]


class DenseOp(C.Structure):
    _fields_ = [("n", c_int), ("A", c_void_p)]


class Grid(C.Structure):
    _fields_ = [("nr1", c_int), ("nr2", c_int), ("nr3", c_int), ("vrs", c_void_p)]


class KPoint(C.Structure):
    _fields_ = [("npw", c_int), ("npwx", c_int), ("nl_igk", c_void_p), ("g2kin", c_void_p),
                ("nkb", c_int), ("vkb", c_void_p), ("dion", c_void_p), ("nbnd_occ", c_int),
                ("evq", c_void_p), ("alpha_pv", c_double)]


class PwOp(C.Structure):
    _fields_ = [("grid", C.POINTER(Grid)), ("kp", C.POINTER(KPoint)), ("alpha_pv", c_double),
                ("work", c_void_p), ("becp", c_void_p)]


class KPair(C.Structure):
    _fields_ = [("kq", KPoint), ("npw_k", c_int), ("nl_igk_k", c_void_p), ("nbnd", c_int),
                ("evc", c_void_p), ("et", c_void_p), ("wk", c_double)]


class System(C.Structure):
    _fields_ = [("grid", Grid), ("nks", c_int), ("kp", C.POINTER(KPair)), ("omega_cell", c_double),
                ("tpiba2", c_double), ("xq", c_double * 3), ("ngm", c_int), ("g", c_void_p),
                ("nl", c_void_p)]


def build(native: bool = False, force: bool = False) -> Path:
    out = "liboracle_native.so" if native else "liboracle.so"
    target = _HERE / out
    srcs = [_HERE / "solver.c", _HERE / "pw.c", _HERE / "sgw_oracle.h"]
    if force or not target.exists() or any(s.stat().st_mtime > target.stat().st_mtime for s in srcs):
        args = ["make", "-C", str(_HERE), f"OUT={out}"] + (["ARCH=native"] if native else [])
        subprocess.run(args, check=True, capture_output=True)
    return target


def lib(native: bool = False):
    global _LIB
    key = "native" if native else "portable"
    if _LIB is None:
        _LIB = {}
    if key not in _LIB:
        path = build(native=native)
        L = C.CDLL(str(path))
        L.orc_norm.restype = c_double
        L.orc_dnrm2.restype = c_double
        _LIB[key] = L
    return _LIB[key]


def _p(a):
    return a.ctypes.data_as(c_void_p)


def _c16(a):
    return np.ascontiguousarray(a, dtype=np.complex128)


def make_cfg(priority=(1, 3), max_iter=10000, threshold=1e-4, lmax=4) -> SolverCfg:
    cfg = SolverCfg()
    cfg.npriority = len(priority)
    for i, p in enumerate(priority):
        cfg.priority[i] = p
    cfg.max_iter, cfg.threshold, cfg.bicg_lmax = max_iter, threshold, lmax
    return cfg


# ------------------------------------------------------------------ solver half, dense fake backend
def _dense_ctx(A):
    A = np.asfortranarray(A, dtype=np.complex128)
    op = DenseOp(A.shape[0], _p(A))
    return op, A


def bicgstab_dense(A, b, sigma, lmax=4, threshold=1e-4, max_iter=10000, native=False):
    """bicgstab(config, AA, bb, sigma, xx, ierr) with the dense operator of linear_solver.pf:106."""
    L = lib(native)
    op, keep = _dense_ctx(A)
    b, sigma = _c16(b), _c16(np.atleast_1d(sigma))
    n, ns = b.size, sigma.size
    xx = np.zeros((n, ns), dtype=np.complex128, order="F")
    st = Stats()
    ierr = L.orc_bicgstab(c_int(lmax), c_double(threshold), c_int(max_iter), L.orc_dense_apply, C.byref(op),
                          c_int(n), _p(b), c_int(ns), _p(sigma), _p(xx), C.byref(st))
    return xx, ierr, {"n_op": st.n_op, "n_outer": st.n_outer}


def linear_solver_dense(A, b, sigma, threshold=1e-4, max_iter=10000, native=False):
    L = lib(native)
    op, keep = _dense_ctx(A)
    b, sigma = _c16(b), _c16(np.atleast_1d(sigma))
    n, ns = b.size, sigma.size
    xx = np.zeros((n, ns), dtype=np.complex128, order="F")
    st = Stats()
    ierr = L.orc_linear_solver(c_double(threshold), c_int(max_iter), L.orc_dense_apply, C.byref(op),
                               c_int(n), _p(b), c_int(ns), _p(sigma), _p(xx), C.byref(st))
    return xx, ierr, {"n_op": st.n_op, "n_outer": st.n_outer}


def select_solver_dense(A, b, sigma, cfg: SolverCfg, native=False):
    L = lib(native)
    op, keep = _dense_ctx(A)
    b, sigma = _c16(b), _c16(np.atleast_1d(sigma))
    n, ns = b.size, sigma.size
    xx = np.zeros((n, ns), dtype=np.complex128, order="F")
    st = Stats()
    ierr = L.orc_select_solver(C.byref(cfg), L.orc_dense_apply, C.byref(op), c_int(n), _p(b), c_int(ns),
                               _p(sigma), _p(xx), C.byref(st))
    return xx, ierr, {"n_op": st.n_op, "n_outer": st.n_outer, "solver_used": st.solver_used}


def gram_schmidt(first, basis, vector=None):
    L = lib()
    basis = np.asfortranarray(basis, dtype=np.complex128).copy(order="F")
    n, nb = basis.shape
    if vector is not None:
        vector = np.asfortranarray(vector, dtype=np.complex128).copy(order="F")
    L.orc_gram_schmidt(c_int(first), c_int(n), c_int(nb), _p(basis), _p(vector) if vector is not None else None)
    return basis, vector


def norm(v):
    v = _c16(v)
    return lib().orc_norm(_p(v), c_int(v.size))


def parallel_task(nproc, rank, ntotal):
    first, last = c_int(), c_int()
    num = (c_int * nproc)()
    lib().orc_parallel_task(c_int(nproc), c_int(rank), c_int(ntotal), C.byref(first), C.byref(last), num)
    return first.value, last.value, list(num)


def fft3d(f, sign):
    """sign=+1: QE invfft (unscaled, e^{+iGr}); sign=-1: QE fwfft (scaled 1/nnr). f: (nr1,nr2,nr3) F-order."""
    f = np.asfortranarray(f, dtype=np.complex128).copy(order="F")
    lib().orc_fft3d(_p(f), c_int(f.shape[0]), c_int(f.shape[1]), c_int(f.shape[2]), c_int(sign))
    return f


# ------------------------------------------------------------------ plane-wave half
class MetalPair(C.Structure):
    _fields_ = [("nbnd", c_int), ("evq_all", c_void_p), ("et_q", c_void_p), ("nocc_k", c_int), ("wg_over_wk", c_void_p)]


def wgauss(x, n):
    L = lib()
    L.orc_wgauss.restype = c_double
    return float(L.orc_wgauss(c_double(x), c_int(n)))


def w0gauss(x, n):
    L = lib()
    L.orc_w0gauss.restype = c_double
    return float(L.orc_w0gauss(c_double(x), c_int(n)))


def metal_weight(e_i, e_j, j_in_projector, alpha_pv, ef, degauss, ngauss):
    L = lib()
    L.orc_metal_weight.restype = c_double
    return float(L.orc_metal_weight(c_double(e_i), c_double(e_j), c_int(1 if j_in_projector else 0), c_double(alpha_pv),
                                    c_double(ef), c_double(degauss), c_int(ngauss)))


class PwSystem:
    """Keeps numpy arrays alive and exposes the C structs for a synthetic system (see synth/)."""

    def __init__(self, syn, native=False):
        self.syn = syn
        self.native = native
        self._keep = []
        g = Grid(syn.nr[0], syn.nr[1], syn.nr[2], self._k(np.ascontiguousarray(syn.vrs, dtype=np.float64)))
        self.grid = g
        self.kpairs = (KPair * len(syn.kpairs))()
        for i, kp in enumerate(syn.kpairs):
            self.kpairs[i].kq = self._kpoint(kp.kq)
            self.kpairs[i].npw_k = kp.npw_k
            self.kpairs[i].nl_igk_k = self._k(np.ascontiguousarray(kp.nl_igk_k, dtype=np.int32))
            self.kpairs[i].nbnd = kp.evc.shape[1]
            self.kpairs[i].evc = self._k(np.asfortranarray(kp.evc, dtype=np.complex128))
            self.kpairs[i].et = self._k(np.ascontiguousarray(kp.et, dtype=np.float64))
            self.kpairs[i].wk = kp.wk
        s = System()
        s.grid = g
        s.nks = len(syn.kpairs)
        s.kp = C.cast(self.kpairs, C.POINTER(KPair))
        s.omega_cell = syn.omega_cell
        s.tpiba2 = syn.tpiba2
        for i in range(3):
            s.xq[i] = syn.xq[i]
        s.ngm = syn.ngm
        s.g = self._k(np.ascontiguousarray(syn.g.T, dtype=np.float64))  # (ngm,3) row-major == 3 x ngm col-major
        s.nl = self._k(np.ascontiguousarray(syn.nl, dtype=np.int32))
        self.sys = s
        # metals: the oracle reads klist/ener like the reference does, from globals that `_smearing()` sets before every call
        self.metal = getattr(syn, "metal", None)
        if self.metal is not None:
            self.metal_pairs = (MetalPair * len(syn.kpairs))()
            for i, m in enumerate(self.metal.pairs):
                self.metal_pairs[i].nbnd = m.evq_all.shape[1]
                self.metal_pairs[i].evq_all = self._k(np.asfortranarray(m.evq_all, dtype=np.complex128))
                self.metal_pairs[i].et_q = self._k(np.ascontiguousarray(m.et_q, dtype=np.float64))
                self.metal_pairs[i].nocc_k = m.nocc_k
                self.metal_pairs[i].wg_over_wk = self._k(np.ascontiguousarray(m.wg_over_wk, dtype=np.float64))

    def _smearing(self):
        L = lib(self.native)
        L.orc_set_smearing.argtypes = [c_int, c_double, c_double, c_int, c_int, c_void_p]
        if self.metal is None:
            L.orc_set_smearing(0, 0.0, 0.0, 0, 0, None)
        else:
            L.orc_set_smearing(1, self.metal.ef, self.metal.degauss, self.metal.ngauss, len(self.metal.pairs),
                               C.cast(self.metal_pairs, c_void_p))

    def _k(self, a):
        self._keep.append(a)
        return _p(a)

    def _kpoint(self, kq) -> KPoint:
        k = KPoint()
        k.npw, k.npwx = kq.npw, kq.npwx
        k.nl_igk = self._k(np.ascontiguousarray(kq.nl_igk, dtype=np.int32))
        k.g2kin = self._k(np.ascontiguousarray(kq.g2kin, dtype=np.float64))
        k.nkb = kq.vkb.shape[1]
        k.vkb = self._k(np.asfortranarray(kq.vkb, dtype=np.complex128))
        k.dion = self._k(np.asfortranarray(kq.dion, dtype=np.float64))
        k.nbnd_occ = kq.evq.shape[1]
        k.evq = self._k(np.asfortranarray(kq.evq, dtype=np.complex128))
        k.alpha_pv = kq.alpha_pv
        return k

    # linear_op on one vector (npwx-padded)
    def linear_op(self, ik, omega, alpha_pv, psi):
        L = lib(self.native)
        kq = self.kpairs[ik].kq
        nnr = int(np.prod(self.syn.nr))
        psi = _c16(psi)
        out = np.zeros(kq.npwx, dtype=np.complex128)
        work = np.zeros(nnr, dtype=np.complex128)
        becp = np.zeros(2 * (kq.nkb + kq.nbnd_occ + 1), dtype=np.complex128)
        L.orc_linear_op.argtypes = [C.POINTER(Grid), C.POINTER(KPoint), ZC, c_double, c_void_p,
                                    c_void_p, c_void_p, c_void_p]
        om = ZC(complex(omega).real, complex(omega).imag)
        L.orc_linear_op(C.byref(self.grid), C.byref(kq), om, c_double(alpha_pv), _p(psi), _p(out), _p(work), _p(becp))
        return out

    def select_solver(self, ik, b, sigma, cfg: SolverCfg, alpha_pv=None):
        """select_solver(config, coulomb_operator|green_operator, bb, sigma, xx, ierr) on the plane-wave operator."""
        L = lib(self.native)
        kq = self.kpairs[ik].kq
        nnr = int(np.prod(self.syn.nr))
        work = np.zeros(nnr, dtype=np.complex128)
        becp = np.zeros(2 * (kq.nkb + kq.nbnd_occ + 1), dtype=np.complex128)
        op = PwOp(C.pointer(self.grid), C.pointer(kq), kq.alpha_pv if alpha_pv is None else alpha_pv, _p(work), _p(becp))
        b, sigma = _c16(b), _c16(np.atleast_1d(sigma))
        n, ns = b.size, sigma.size
        xx = np.zeros((n, ns), dtype=np.complex128, order="F")
        st = Stats()
        ierr = L.orc_select_solver(C.byref(cfg), L.orc_pw_apply, C.byref(op), c_int(n), _p(b), c_int(ns), _p(sigma),
                                   _p(xx), C.byref(st))
        return xx, ierr, {"n_op": st.n_op, "n_outer": st.n_outer, "solver_used": st.solver_used}

    def solve_linter(self, dvbare, freq, cfg: SolverCfg, nthreads=1):
        self._smearing()
        L = lib(self.native)
        nnr = int(np.prod(self.syn.nr))
        dvbare, freq = _c16(np.ravel(dvbare, order="F")), _c16(freq)
        drho = np.zeros((nnr, freq.size), dtype=np.complex128, order="F")
        st = Stats()
        ierr = L.orc_solve_linter(C.byref(self.sys), C.byref(cfg), _p(dvbare), c_int(freq.size), _p(freq), _p(drho),
                                  C.byref(st), c_int(nthreads))
        return drho, ierr, {"n_op": st.n_op, "n_outer": st.n_outer}

    def solve_linter_iter(self, num_iter, alpha_mix, tr2_gw, nmix_gw, dvbare, freq, cfg: SolverCfg, nthreads=1):
        """solve_linter.f90 with num_iter > 1 (self-consistent branch + mix_potential_c): returns dvscfin(nnr, nfreq)."""
        self._smearing()
        L = lib(self.native)
        nnr = int(np.prod(self.syn.nr))
        dvbare, freq = _c16(np.ravel(dvbare, order="F")), _c16(freq)
        am = np.ascontiguousarray(np.broadcast_to(np.asarray(alpha_mix, dtype=np.float64), (num_iter,)))
        out = np.zeros((nnr, freq.size), dtype=np.complex128, order="F")
        st = Stats()
        it = c_int(0)
        ierr = L.orc_solve_linter_iter(C.byref(self.sys), C.byref(cfg), c_int(num_iter), _p(am), c_double(tr2_gw),
                                       c_int(nmix_gw), _p(dvbare), c_int(freq.size), _p(freq), _p(out), C.byref(st),
                                       c_int(nthreads), C.byref(it))
        return out, ierr, {"n_op": st.n_op, "n_outer": st.n_outer, "iter": it.value}

    def coulomb(self, igstart, ngc, ntask, ig_unique, fiu, cfg: SolverCfg, nthreads=1):
        self._smearing()
        L = lib(self.native)
        fiu = _c16(fiu)
        ig_unique = np.ascontiguousarray(ig_unique, dtype=np.int32)
        scr = np.zeros((ngc, fiu.size, ntask), dtype=np.complex128, order="F")
        st = Stats()
        ierr = L.orc_coulomb(C.byref(self.sys), C.byref(cfg), c_int(igstart), c_int(ngc), c_int(ntask), _p(ig_unique),
                             c_int(fiu.size), _p(fiu), _p(scr), C.byref(st), c_int(nthreads))
        return scr, ierr, {"n_op": st.n_op, "n_outer": st.n_outer}

    def set_band_window(self, lo=0, hi=2147483647):
        """bench only: bounded CPU-baseline samples solve bands lo <= ibnd < hi (default: all, as the reference)."""
        lib(self.native).orc_set_band_window(c_int(lo), c_int(hi))

    def coulomb_q0G0(self, fiu, cfg: SolverCfg):
        self._smearing()
        L = lib(self.native)
        fiu = _c16(fiu)
        eps = np.zeros(fiu.size, dtype=np.complex128)
        st = Stats()
        ierr = L.orc_coulomb_q0G0(C.byref(self.sys), C.byref(cfg), c_int(fiu.size), _p(fiu), _p(eps), C.byref(st))
        return eps, ierr, {"n_op": st.n_op, "n_outer": st.n_outer}

    def green_function(self, ik, map_, fft_map, omega, cfg: SolverCfg, nthreads=1):
        L = lib(self.native)
        kq = self.kpairs[ik].kq
        omega = _c16(omega)
        map_ = np.ascontiguousarray(map_, dtype=np.int32)
        fft_map = np.ascontiguousarray(fft_map, dtype=np.int32)
        ngc, ngp = map_.size, fft_map.size
        green = np.zeros((ngc, ngp, omega.size), dtype=np.complex128, order="F")
        st = Stats()
        ierr = L.orc_green_function(C.byref(self.grid), C.byref(kq), C.byref(cfg), c_int(ngc), _p(map_), c_int(ngp),
                                    _p(fft_map), c_int(omega.size), _p(omega), _p(green), C.byref(st), c_int(nthreads))
        return green, ierr, {"n_op": st.n_op, "n_outer": st.n_outer}


def unfold_w(ngc, nfs, ig_unique, scr_in):
    ig_unique = np.ascontiguousarray(ig_unique, dtype=np.int32)
    scr_in = np.asfortranarray(scr_in, dtype=np.complex128)
    out = np.zeros((ngc, ngc, nfs), dtype=np.complex128, order="F")
    lib().orc_unfold_w(c_int(ngc), c_int(nfs), c_int(ig_unique.size), _p(ig_unique), _p(scr_in), _p(out))
    return out


def invert_epsilon(scr, lgamma=False):
    scr = np.asfortranarray(scr, dtype=np.complex128).copy(order="F")
    ngc, _, nfs = scr.shape
    info = lib().orc_invert_epsilon(c_int(ngc), c_int(nfs), _p(scr), c_int(1 if lgamma else 0))
    return scr, info
