"""Host-side mirror of the reference's operator/plugin interface for the Sternheimer path (see __init__)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib


class SgwError(RuntimeError):
    pass


@dataclass
class select_solver_type:
    """Configuration of the linear solver (select_solver.f90:48-62); same names and defaults."""
    priority: tuple = (1, 3)       # main/src/gw_input.yml:151-159 default priority_coul / priority_green
    max_iter: int = 10000
    threshold: float = 1e-4
    bicg_lmax: int = 4

    def c(self) -> _lib.SolverCfg:
        if not self.priority:
            raise SgwError("priority of the solvers not specified")     # select_solver.f90:116-118
        cfg = _lib.SolverCfg()
        cfg.npriority = len(self.priority)
        for i, p in enumerate(self.priority):
            cfg.priority[i] = int(p)
        cfg.max_iter, cfg.threshold, cfg.bicg_lmax = int(self.max_iter), float(self.threshold), int(self.bicg_lmax)
        return cfg


# model_coul (analytic.f90:39-60) and freq_symm_coul (freqbins.f90:31-35)
godby_needs, pade_approx, pade_robust, aaa_approx, aaa_pole = 1, 2, 3, 4, 5
no_symmetry, even_symmetry, square_symmetry = 0, 1, 2


@dataclass
class freqbins_type:
    """The members of freqbins_type (algo/grid/src/freqbins.f90:42-105) the continuation and the G W convolution read;
    same names.  `num_freq`, `green` and `symmetrize` follow freqbins.f90:190,222,322."""
    solver: np.ndarray
    coul: np.ndarray = None
    weight: np.ndarray = None
    sigma: np.ndarray = None
    freq_symm_coul: int = even_symmetry
    imag_sigma: bool = True

    def c(self):
        """(sgw_freqbins, keep-alive list of the arrays it points to)."""
        keep = [np.ascontiguousarray(self.solver, dtype=np.complex128)]
        fb = _lib.Freqbins()
        fb.imag_sigma, fb.freq_symm_coul = int(bool(self.imag_sigma)), int(self.freq_symm_coul)
        fb.num_solver, fb.solver = keep[0].size, keep[0].ctypes.data
        if self.coul is not None:
            coul = np.ascontiguousarray(self.coul, dtype=np.complex128)
            weight = np.ascontiguousarray(self.weight, dtype=np.float64)
            if weight.size != coul.size:
                raise SgwError("freqbins: one weight per integration frequency")
            fb.num_coul, fb.coul, fb.weight = coul.size, coul.ctypes.data, weight.ctypes.data
            keep += [coul, weight]
        if self.sigma is not None:
            sig = np.ascontiguousarray(self.sigma, dtype=np.complex128)
            fb.num_sigma, fb.sigma = sig.size, sig.ctypes.data
            keep.append(sig)
        return fb, keep

    def num_freq(self):
        fb, _keep = self.c()
        n = _lib.load().sgw_freqbins_num_freq(C.byref(fb))
        if n < 0:
            raise SgwError("only a single frequency may be smaller than 1e-14")       # freqbins.f90:276
        return n

    def num_coul(self):
        return int(np.size(self.coul))

    def num_sigma(self):
        return int(np.size(self.sigma))

    def green(self, freq_sigma):
        return np.concatenate([freq_sigma + np.asarray(self.coul), freq_sigma - np.asarray(self.coul)])

    def symmetrize(self, freq):
        return freq ** 2 if self.freq_symm_coul == square_symmetry else freq


def gauleg_grid(x1, x2, n):
    """Gauss-Legendre abscissas and weights on [x1, x2] (algo/grid/src/gauleg_grid.f90:23: Newton iteration on P_n with the
    Chebyshev starting guess, relative precision 3e-14); host logic of the frequency meshes, numpy only."""
    i = np.arange(1, (n + 1) // 2 + 1)
    z = np.cos(np.pi * (i - 0.25) / (n + 0.5))
    pp = np.ones_like(z)
    todo = np.ones(z.shape, dtype=bool)              # every root stops on its own criterion, like the reference's scalar loop
    for _ in range(100):
        p1, p2 = np.ones_like(z), np.zeros_like(z)
        for j in range(1, n + 1):
            p1, p2 = ((2.0 * j - 1.0) * z * p1 - (j - 1.0) * p2) / j, p1
        ppn = n * (z * p1 - p2) / (z * z - 1.0)
        zn = z - p1 / ppn
        step = np.abs(zn - z)
        pp = np.where(todo, ppn, pp)
        z = np.where(todo, zn, z)
        todo = todo & (step > 3e-14)
        if not todo.any():
            break
    xm, xl = 0.5 * (x2 + x1), 0.5 * (x2 - x1)
    x, w = np.zeros(n), np.zeros(n)
    wz = 2.0 * xl / ((1.0 - z * z) * pp * pp)
    x[i - 1], x[n - i] = xm - xl * z, xm + xl * z
    w[i - 1], w[n - i] = wz, wz
    return x, w


def freqbins(imag_sigma, min_sigma, max_sigma, num_sigma, max_coul, num_coul, solver, freq_symm_coul=even_symmetry, eta=0.0):
    """freqbins (algo/grid/src/freqbins.f90:109-180): the self-energy mesh (equidistant) and the integration mesh of the
    G W convolution -- Gauss-Legendre nodes on the imaginary axis, or an equidistant mesh shifted by i eta on the real axis
    with the trapezoid-like weights of :156-158 -- as a `freqbins_type`."""
    grid = min_sigma + (max_sigma - min_sigma) / (num_sigma - 1) * np.arange(num_sigma)
    if imag_sigma:
        g, weight = gauleg_grid(0.0, max_coul, num_coul)
        return freqbins_type(np.asarray(solver, dtype=complex), 1j * g, weight, 1j * grid, freq_symm_coul, True)
    g = max_coul / (num_coul - 1) * np.arange(num_coul)
    weight = np.full(num_coul, 2.0 * max_coul / float(2 * num_coul - 1))
    weight[0] *= 0.5
    return freqbins_type(np.asarray(solver, dtype=complex), g + 1j * eta, weight, grid.astype(complex), freq_symm_coul, False)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _c16(a):
    return np.require(a, dtype=np.complex128, requirements=["F_CONTIGUOUS", "ALIGNED"])


def parallel_task(nproc: int, rank: int, num_task_total: int):
    """parallel.f90:80-138: returns (first_task, last_task, num_task[nproc]); first/last 1-based, rank 0-based."""
    L = _lib.load()
    first, last = C.c_int32(), C.c_int32()
    num = (C.c_int32 * nproc)()
    rc = L.sgw_parallel_task(nproc, rank, num_task_total, C.byref(first), C.byref(last), num)
    if rc != 0:
        raise SgwError(f"parallel_task: invalid argument ({rc})")
    return first.value, last.value, list(num)


class Context:
    """One GPU context = the state the reference keeps in QE module globals for one MPI rank."""

    def __init__(self, device: int = 0):
        self._L = _lib.load()
        self._h = C.c_void_p()
        rc = self._L.sgw_create(device, C.byref(self._h))
        if rc != 0:
            self._h = None
            raise SgwError(f"sgw_create(device={device}) failed ({rc}): no usable CUDA device -- there is no CPU fallback")
        self.npw = {}
        self.npwx = {}

    def close(self):
        if getattr(self, "_h", None):
            self._L.sgw_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc, what):
        if rc < 0:
            raise SgwError(f"{what} failed ({rc}): {self._L.sgw_last_error(self._h).decode()}")
        return rc

    def stats(self) -> dict:
        st = _lib.Stats()
        self._L.sgw_get_stats(self._h, C.byref(st))
        return {k: getattr(st, k) for k, _ in st._fields_}

    def set_profiling(self, on: bool):
        self._chk(self._L.sgw_set_profiling(self._h, 1 if on else 0), "set_profiling")

    def profile(self) -> dict:
        """Per-kernel-class device milliseconds and region counts of the last solver-level call."""
        ms = (C.c_double * 16)()
        cnt = (C.c_int64 * 16)()
        n = C.c_int(0)
        self._chk(self._L.sgw_get_profile(self._h, 16, ms, cnt, C.byref(n)), "get_profile")
        return {self._L.sgw_profile_class_name(i).decode(): {"ms": ms[i], "regions": int(cnt[i])} for i in range(n.value)}

    def set_stream(self, cuda_stream):
        """Run on a caller-owned CUDA stream (an integer handle such as torch.cuda.Stream.cuda_stream); None restores the
        context's own stream."""
        self._chk(self._L.sgw_set_stream(self._h, C.c_void_p(int(cuda_stream)) if cuda_stream else None), "set_stream")

    def set_message_callback(self, fn):
        """fn(str) receives the solver warnings the reference writes to stdout (bicgstab.f90:250, select_solver.f90:126,
        linear_solver.f90:177); None removes the callback."""
        self._msg_cb = _lib.MESSAGE_FN(lambda msg, _user: fn(msg.decode())) if fn is not None else _lib.MESSAGE_FN()
        self._chk(self._L.sgw_set_message_callback(self._h, self._msg_cb, None), "set_message_callback")

    def synchronize(self):
        self._chk(self._L.sgw_device_synchronize(self._h), "synchronize")

    # ---- L0 / L1: operator installation -------------------------------------------------------------
    def set_grid(self, nr1, nr2, nr3, nr1x=None, nr2x=None, nr3x=None):
        self._chk(self._L.sgw_set_grid(self._h, nr1, nr2, nr3, nr1x or nr1, nr2x or nr2, nr3x or nr3), "set_grid")

    def set_vloc(self, vrs):
        vrs = np.ascontiguousarray(vrs, dtype=np.float64)
        self._chk(self._L.sgw_set_vloc(self._h, _p(vrs)), "set_vloc")

    def set_kpoint(self, slot, npw, npwx, nl_igk, g2kin, vkb, dion, evq, alpha_pv):
        nl_igk = np.ascontiguousarray(nl_igk, dtype=np.int32)
        g2kin = np.ascontiguousarray(g2kin, dtype=np.float64)
        vkb, evq = _c16(vkb), _c16(evq)
        dion = np.asfortranarray(dion, dtype=np.float64)
        nkb = vkb.shape[1] if vkb.ndim == 2 else 0
        nb = evq.shape[1] if evq.ndim == 2 else 0
        assert nkb == 0 or vkb.shape[0] == npwx
        assert nb == 0 or evq.shape[0] == npwx
        self._chk(self._L.sgw_set_kpoint(self._h, slot, npw, npwx, _p(nl_igk), _p(g2kin), nkb, _p(vkb), _p(dion), nb,
                                         _p(evq), float(alpha_pv)), "set_kpoint")
        self.npw[slot], self.npwx[slot] = npw, npwx

    def set_dense_operator(self, slot, A):
        A = _c16(A)
        n = A.shape[0]
        self._chk(self._L.sgw_set_dense_operator(self._h, slot, n, _p(A), n), "set_dense_operator")
        self.npw[slot], self.npwx[slot] = n, n

    def install_system(self, syn, kpairs=None, nrx=None):
        """Install a synth.SynthSystem the way the Fortran host would (gwq_setup, solve_linter.f90:300-316).
        kpairs: indices of the (k, k+q) pairs THIS context works on (a pool of the reference, solve_linter.f90:521); default all.
        nrx: physical dimensions (dffts%nr1x, nr2x, nr3x) of a padded box: the potential and every index array are then handed
        over in the padded layout, as a host whose FFT descriptor pads would hold them."""
        nr = tuple(int(x) for x in syn.nr)
        if nrx is None:
            padi = lambda a: np.ascontiguousarray(a, dtype=np.int32)
            self.set_grid(*nr)
            self.set_vloc(syn.vrs)
        else:
            nrx = tuple(int(x) for x in nrx)

            def padi(a):
                i = np.asarray(a, dtype=np.int64) - 1
                x, y, z = i % nr[0], (i // nr[0]) % nr[1], i // (nr[0] * nr[1])
                return np.ascontiguousarray(x + nrx[0] * (y + nrx[1] * z) + 1, dtype=np.int32)
            vp = np.zeros(nrx, dtype=np.float64, order="F")
            vp[:nr[0], :nr[1], :nr[2]] = np.asarray(syn.vrs).reshape(nr, order="F")
            self.set_grid(*nr, *nrx)
            self.set_vloc(vp.ravel(order="F"))
        self._chk(self._L.sgw_set_system(self._h, syn.omega_cell, syn.tpiba2, syn.ngm,
                                         _p(np.ascontiguousarray(syn.g.T, dtype=np.float64)),
                                         _p(padi(syn.nl))), "set_system")
        self.set_q(syn.xq)
        pairs = list(syn.kpairs) if kpairs is None else [syn.kpairs[i] for i in kpairs]
        self._chk(self._L.sgw_set_nksq(self._h, len(pairs)), "set_nksq")
        for ik, kp in enumerate(pairs):
            kq = kp.kq
            self.set_kpoint(ik, kq.npw, kq.npwx, padi(kq.nl_igk), kq.g2kin, kq.vkb, kq.dion, kq.evq, kq.alpha_pv)
            evc = _c16(kp.evc)
            et = np.ascontiguousarray(kp.et, dtype=np.float64)
            nl = padi(kp.nl_igk_k)
            self._chk(self._L.sgw_set_kpair(self._h, ik, ik, kp.npw_k, _p(nl), evc.shape[1], _p(evc), _p(et), float(kp.wk)),
                      "set_kpair")
        metal = getattr(syn, "metal", None)
        if metal is None:
            self.set_smearing(False)
        else:                                   # klist lgauss/degauss/ngauss, ener ef + what orthogonalize's lgauss branch reads
            self.set_smearing(True, metal.ef, metal.degauss, metal.ngauss)
            sel = range(len(syn.kpairs)) if kpairs is None else kpairs
            for ik, i in enumerate(sel):
                m = metal.pairs[i]
                self.set_kpair_metal(ik, m.evq_all, m.et_q, m.nocc_k, m.wg_over_wk)

    def set_smearing(self, lgauss, ef=0.0, degauss=0.0, ngauss=0):
        self._chk(self._L.sgw_set_smearing(self._h, 1 if lgauss else 0, float(ef), float(degauss), int(ngauss)), "set_smearing")

    def set_kpair_metal(self, ik, evq_all, et_q, nocc_k, wg_over_wk):
        evq_all = _c16(evq_all)
        et_q = np.ascontiguousarray(et_q, dtype=np.float64)
        w = np.ascontiguousarray(wg_over_wk, dtype=np.float64)
        assert et_q.size == evq_all.shape[1] and w.size == nocc_k
        self._chk(self._L.sgw_set_kpair_metal(self._h, ik, evq_all.shape[1], _p(evq_all), _p(et_q), int(nocc_k), _p(w)),
                  "set_kpair_metal")

    def set_q(self, xq):
        xq = np.ascontiguousarray(xq, dtype=np.float64)
        self._chk(self._L.sgw_set_q(self._h, _p(xq)), "set_q")

    # ---- linear_op(current_k, num_g, omega, alpha_pv, psi, A_psi) ------------------------------------
    def linear_op(self, slot, omega, alpha_pv, psi):
        psi = _c16(psi)
        if psi.ndim == 1:
            psi = psi.reshape(-1, 1, order="F")
        omega = _c16(np.atleast_1d(omega))
        nvec = psi.shape[1]
        if omega.size != nvec:
            raise SgwError("Second dimension of vector psi should be num_band")      # linear_op.f90:101-102
        out = np.zeros_like(psi, order="F")
        self._chk(self._L.sgw_linear_op(self._h, slot, nvec, _p(omega), float(alpha_pv), _p(psi), psi.shape[0], _p(out),
                                        out.shape[0]), "linear_op")
        return out

    # ---- select_solver(config, AA, bb, sigma, xx, ierr), batched over right-hand sides ----------------
    def select_solver(self, config: select_solver_type, slot, bb, sigma, use_alpha_pv=True):
        """bb: (n,) or (n, nrhs); sigma: (nshift,) or (nshift, nrhs).  Returns xx (n, nshift[, nrhs]) and ierr."""
        bb = _c16(bb)
        single = bb.ndim == 1
        if single:
            bb = bb.reshape(-1, 1, order="F")
        n, nrhs = bb.shape
        sigma = _c16(sigma)
        if sigma.ndim == 1:
            sigma = np.asfortranarray(np.repeat(sigma.reshape(-1, 1), nrhs, axis=1))
        nshift = sigma.shape[0]
        if sigma.shape[1] != nrhs:
            raise SgwError("we need one shift list per right-hand side")
        xx = np.zeros((n, nshift, nrhs), dtype=np.complex128, order="F")
        ierr = np.zeros(nrhs, dtype=np.int32)
        cfg = config.c()
        self._chk(self._L.sgw_solve_multishift(self._h, slot, C.byref(cfg), 1 if use_alpha_pv else 0, nrhs, _p(bb), n,
                                               nshift, _p(sigma), _p(xx), n, n * nshift, _p(ierr)), "select_solver")
        if single:
            return xx[:, :, 0], int(ierr[0])
        return xx, ierr

    # ---- phys/coul -------------------------------------------------------------------------------------
    def set_mixing(self, niter_gw, alpha_mix, tr2_gw, nmix_gw):
        """control_gw globals of the self-consistent W branch (num_iter_coul, alpha_mix, tr2_gw, num_mix_coul)."""
        am = np.ascontiguousarray(np.broadcast_to(np.asarray(alpha_mix, dtype=np.float64), (niter_gw,)))
        self._chk(self._L.sgw_set_mixing(self._h, int(niter_gw), _p(am), float(tr2_gw), int(nmix_gw)), "set_mixing")

    def set_solve_direct(self, solve_direct: bool):
        """control_gw solve_direct: False makes `coulomb` run the self-consistent branch (needs set_mixing)."""
        self._chk(self._L.sgw_set_solve_direct(self._h, 1 if solve_direct else 0), "set_solve_direct")

    def scf_iterations(self):
        return int(self._L.sgw_get_scf_iterations(self._h))

    def solve_linter(self, config: select_solver_type, num_iter, dvbarein, freq):
        dvbarein, freq = _c16(np.ravel(dvbarein, order="F")), _c16(freq)
        drho = np.zeros((dvbarein.size, freq.size), dtype=np.complex128, order="F")
        ierr = C.c_int32(0)
        cfg = config.c()
        self._chk(self._L.sgw_solve_linter(self._h, C.byref(cfg), num_iter, _p(dvbarein), freq.size, _p(freq), _p(drho),
                                           C.byref(ierr)), "solve_linter")
        if ierr.value == 10:
            raise SgwError("Iterative solver did not converge within given number of iterations")   # solve_linter.f90:588-591
        if ierr.value != 0:
            raise SgwError(f"solver did not converge (ierr={ierr.value})")          # solve_linter.f90:370
        return drho

    def coulomb(self, config: select_solver_type, igstart, num_g_corr, num_task, ig_unique, fiu, check=True):
        fiu = _c16(fiu)
        ig_unique = np.ascontiguousarray(ig_unique, dtype=np.int32)
        scr = np.zeros((num_g_corr, fiu.size, num_task), dtype=np.complex128, order="F")
        ierr = C.c_int32(0)
        cfg = config.c()
        self._chk(self._L.sgw_coulomb(self._h, C.byref(cfg), igstart, num_g_corr, num_task, _p(ig_unique), fiu.size,
                                      _p(fiu), _p(scr), C.byref(ierr)), "coulomb")
        if check and ierr.value == 10:
            raise SgwError("Iterative solver did not converge within given number of iterations")
        if check and ierr.value != 0:
            raise SgwError(f"solver did not converge (ierr={ierr.value})")
        return scr

    def rho_grid(self):
        """(reduced, (n1, n2, n3)): the box the last `coulomb` accumulated Delta-rho on (see sgw_get_rho_grid)."""
        dims = (C.c_int * 3)()
        rc = self._chk(self._L.sgw_get_rho_grid(self._h, dims), "get_rho_grid")
        return bool(rc), tuple(dims)

    def coulomb_q0G0(self, config: select_solver_type, fiu):
        fiu = _c16(fiu)
        eps = np.zeros(fiu.size, dtype=np.complex128)
        ierr = C.c_int32(0)
        cfg = config.c()
        self._chk(self._L.sgw_coulomb_q0G0(self._h, C.byref(cfg), fiu.size, _p(fiu), _p(eps), C.byref(ierr)), "coulomb_q0G0")
        if ierr.value != 0:
            raise SgwError(f"solver did not converge (ierr={ierr.value})")
        return eps

    def unfold_w(self, num_g_corr, ig_unique, scrcoul_in):
        scr_in = _c16(scrcoul_in)
        ig_unique = np.ascontiguousarray(ig_unique, dtype=np.int32)
        nfs = scr_in.shape[1]
        out = np.zeros((num_g_corr, num_g_corr, nfs), dtype=np.complex128, order="F")
        self._chk(self._L.sgw_unfold_w(self._h, num_g_corr, nfs, ig_unique.size, _p(ig_unique), _p(scr_in), _p(out)), "unfold_w")
        return out

    def unfold_w_symm(self, num_g_corr, ig_unique, sym_ig, sym_friend, gmapsym, eigv, invs, scrcoul_in):
        """unfold_w with use_symm = .TRUE. (unfold_w.f90:23-131); the symmetry tables are gmap_sym's / stern_symm's."""
        scr_in = _c16(scrcoul_in)
        ig_unique = np.ascontiguousarray(ig_unique, dtype=np.int32)
        sym_ig = np.ascontiguousarray(sym_ig, dtype=np.int32)
        sym_friend = np.ascontiguousarray(sym_friend, dtype=np.int32)
        gm = np.asfortranarray(gmapsym, dtype=np.int32)
        ev = np.asfortranarray(eigv, dtype=np.complex128)
        invs = np.ascontiguousarray(invs, dtype=np.int32)
        nfs = scr_in.shape[1]
        out = np.zeros((num_g_corr, num_g_corr, nfs), dtype=np.complex128, order="F")
        self._chk(self._L.sgw_unfold_w_symm(self._h, num_g_corr, nfs, ig_unique.size, _p(ig_unique), invs.size, _p(sym_ig),
                                            _p(sym_friend), _p(gm), _p(ev), _p(invs), _p(scr_in), _p(out)), "unfold_w_symm")
        return out

    def invert_epsilon(self, scrcoul_g, lgamma=False):
        scr = _c16(scrcoul_g).copy(order="F")
        ngc, _, nfs = scr.shape
        self._chk(self._L.sgw_invert_epsilon(self._h, ngc, nfs, _p(scr), 1 if lgamma else 0), "invert_epsilon")
        return scr

    # ---- phys/green ------------------------------------------------------------------------------------
    def green_function(self, config: select_solver_type, slot, map_, fft_map, omega):
        map_ = np.ascontiguousarray(map_, dtype=np.int32)
        fft_map = np.ascontiguousarray(fft_map, dtype=np.int32)
        omega = _c16(omega)
        green = np.zeros((map_.size, fft_map.size, omega.size), dtype=np.complex128, order="F")
        ierr = C.c_int32(0)
        cfg = config.c()
        self._chk(self._L.sgw_green_function(self._h, slot, C.byref(cfg), map_.size, _p(map_), fft_map.size, _p(fft_map),
                                             omega.size, _p(omega), _p(green), C.byref(ierr)), "green_function")
        if ierr.value != 0:
            raise SgwError(f"the linear solver for G did not converge (ierr={ierr.value})")   # green.f90:208
        return green

    # ---- algo/analytic -------------------------------------------------------------------
    def coulpade(self, factor, scrcoul_g):
        """coulpade.f90:36: scrcoul_g(ig,:,:) *= factor(ig); returns the scaled copy."""
        scr = _c16(scrcoul_g).copy(order="F")
        factor = np.ascontiguousarray(factor, dtype=np.float64)
        ngc, ngc2, nf = scr.shape
        if ngc2 != ngc:
            raise SgwError("input array should have same dimension for G and G'")
        self._chk(self._L.sgw_coulpade(self._h, ngc, nf, _p(factor), _p(scr)), "coulpade")
        return scr

    def pade_robust(self, radius, func, deg_num, deg_den, tol_coeff=None, tol_fft=None):
        """pade_robust (pade_robust.f90:177): returns (deg_num, deg_den, coeff_num, coeff_den)."""
        func = _c16(func)
        dn, dd = C.c_int(int(deg_num)), C.c_int(int(deg_den))
        cn = np.zeros(int(deg_num) + 1, dtype=np.complex128)
        cd = np.zeros(int(deg_den) + 1, dtype=np.complex128)
        self._chk(self._L.sgw_pade_robust(self._h, float(radius), func.size, _p(func), C.byref(dn), C.byref(dd), _p(cn), _p(cd),
                                          float(tol_coeff or 0.0), float(tol_fft or 0.0)), "pade_robust")
        return dn.value, dd.value, cn[:dn.value + 1].copy(), cd[:dd.value + 1].copy()

    def aaa_pole_residual(self, position, value, weight):
        """aaa_pole_residual (vendor/analytic/src/aaa.f90:93) of a given barycentric approximant: returns (pole, residual)."""
        position, value, weight = _c16(position), _c16(value), _c16(weight)
        m = position.size
        if value.size != m or weight.size != m:
            raise SgwError("position, value and weight must have the same size")
        pole = np.zeros(max(m - 1, 1), dtype=np.complex128)
        res = np.zeros(max(m - 1, 1), dtype=np.complex128)
        n = C.c_int(0)
        self._chk(self._L.sgw_aaa_pole_residual(self._h, m, _p(position), _p(value), _p(weight), _p(pole), _p(res), C.byref(n)),
                  "aaa_pole_residual")
        return pole[:n.value].copy(), res[:n.value].copy()

    def analytic_coeff(self, model_coul, thres, freq: "freqbins_type", scrcoul_g):
        """analytic.f90:50: returns the coefficient array (the reference overwrites scrcoul_g in place)."""
        scr = _c16(scrcoul_g).copy(order="F")
        ngc = scr.shape[0]
        if scr.shape[1] != ngc:
            raise SgwError("input array should have same dimension for G and G'")
        if scr.shape[2] != freq.num_freq():
            raise SgwError("frequency dimension of Coulomb inconsistent with frequency mesh")
        fb, _keep = freq.c()
        self._chk(self._L.sgw_analytic_coeff(self._h, int(model_coul), float(thres), C.byref(fb), ngc, _p(scr)), "analytic_coeff")
        return scr

    def analytic_eval(self, model_coul, gmapsym, freq_in: "freqbins_type", scrcoul_coeff, freq_out):
        """analytic.f90:211 for one or several output frequencies: (ngc, ngc[, nout])."""
        coeff = _c16(scrcoul_coeff)
        gmapsym = np.ascontiguousarray(gmapsym, dtype=np.int32)
        fo = _c16(np.atleast_1d(freq_out))
        ngc = gmapsym.size
        out = np.zeros((ngc, ngc, fo.size), dtype=np.complex128, order="F")
        fb, _keep = freq_in.c()
        self._chk(self._L.sgw_analytic_eval(self._h, int(model_coul), C.byref(fb), ngc, _p(gmapsym), _p(coeff), fo.size, _p(fo),
                                            _p(out)), "analytic_eval")
        return out[:, :, 0] if np.ndim(freq_out) == 0 else out

    # ---- data/fft fft6 + phys/corr sigma ----------------------------------------------------
    def set_corr_grid(self, nr, nl):
        nl = np.ascontiguousarray(nl, dtype=np.int32)
        self._chk(self._L.sgw_set_corr_grid(self._h, int(nr[0]), int(nr[1]), int(nr[2]), nl.size, _p(nl)), "set_corr_grid")

    def invfft6(self, f, omega):
        """fft6.f90:231 in place on f(nnr_c, nnr_c) (F-ordered complex128)."""
        assert f.flags.f_contiguous and f.dtype == np.complex128
        self._chk(self._L.sgw_invfft6(self._h, float(omega), _p(f)), "invfft6")

    def fwfft6(self, f, omega):
        """fft6.f90:84 in place; the result is f[:ngm_c, :ngm_c]."""
        assert f.flags.f_contiguous and f.dtype == np.complex128
        self._chk(self._L.sgw_fwfft6(self._h, float(omega), _p(f)), "fwfft6")

    def sigma_correlation(self, omega, config: select_solver_type, slot, mu, alpha, model_coul, freq: "freqbins_type", map_,
                          gmapsym, coulomb, sigma):
        """sigma.f90:528: sigma(ngm_c, ngm_c, num_sigma) += G W convolution for the operator in `slot`; in place."""
        map_ = np.ascontiguousarray(map_, dtype=np.int32)
        gmapsym = np.ascontiguousarray(gmapsym, dtype=np.int32)
        coulomb = _c16(coulomb)
        ngc = map_.size
        assert sigma.flags.f_contiguous and sigma.dtype == np.complex128
        if sigma.shape[2] != freq.num_sigma():
            raise SgwError("frequency dimension of self energy not correct size")         # sigma.f90:633
        if coulomb.shape[0] != ngc or coulomb.shape[1] != ngc:
            raise SgwError("screened Coulomb not a square matrix")                        # sigma.f90:626
        fb, _keep = freq.c()
        cfg = config.c()
        ierr = C.c_int32(0)
        a = _lib.Cplx(float(np.real(alpha)), float(np.imag(alpha)))
        self._chk(self._L.sgw_sigma_correlation(self._h, slot, C.byref(cfg), float(omega), float(mu), a, int(model_coul),
                                                C.byref(fb), ngc, _p(map_), _p(gmapsym), _p(coulomb), _p(sigma), C.byref(ierr)),
                  "sigma_correlation")
        if ierr.value != 0:
            raise SgwError(f"the linear solver for G did not converge (ierr={ierr.value})")   # green.f90:208
        return sigma

    def release_workspace(self):
        """sgw_release_workspace: give the solver scratch memory back to the driver."""
        self._chk(self._L.sgw_release_workspace(self._h), "release_workspace")

    def bench_linear_op(self, slot, nvec, reps=10):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self._chk(self._L.sgw_bench_linear_op(self._h, slot, nvec, reps, C.byref(a), C.byref(b), C.byref(c)), "bench_linear_op")
        return {"ms_total": a.value, "ms_fft": b.value, "ms_gemm": c.value}
