"""ctypes binding of libsgw_b200.so -- exactly the entry points include/sgw_b200.h declares."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libsgw_b200.so"

c_int, c_double, c_void_p, c_int64 = C.c_int, C.c_double, C.c_void_p, C.c_int64

# every symbol of include/sgw_b200.h (tests/test_abi.py checks the list against the header)
SYMBOLS = [
    "sgw_create", "sgw_destroy", "sgw_set_stream", "sgw_last_error", "sgw_get_stats", "sgw_set_profiling", "sgw_release_workspace", "sgw_unfold_w_symm", "sgw_device_synchronize",
    "sgw_get_profile", "sgw_profile_class_name", "sgw_set_message_callback",
    "sgw_set_grid", "sgw_set_vloc", "sgw_set_kpoint", "sgw_set_dense_operator", "sgw_linear_op",
    "sgw_solve_multishift", "sgw_set_system", "sgw_set_q", "sgw_set_nksq", "sgw_set_kpair", "sgw_set_smearing", "sgw_set_kpair_metal", "sgw_set_mixing", "sgw_set_solve_direct", "sgw_get_scf_iterations", "sgw_solve_linter",
    "sgw_coulomb", "sgw_get_rho_grid", "sgw_coulomb_q0G0", "sgw_unfold_w", "sgw_invert_epsilon", "sgw_green_function",
    "sgw_parallel_task", "sgw_bench_linear_op",
    "sgw_freqbins_num_freq", "sgw_pade_robust", "sgw_aaa_pole_residual", "sgw_coulpade", "sgw_analytic_coeff", "sgw_analytic_eval", "sgw_set_corr_grid", "sgw_invfft6",
    "sgw_fwfft6", "sgw_sigma_correlation",
]


class SolverCfg(C.Structure):
    """sgw_solver_cfg == select_solver_type (select_solver.f90:48-62)."""
    _fields_ = [("npriority", C.c_int32), ("priority", C.c_int32 * 4), ("max_iter", C.c_int32),
                ("threshold", c_double), ("bicg_lmax", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("n_linear_op", c_int64), ("n_kernel_launch", c_int64), ("n_outer_max", C.c_int32),
                ("n_fallback", C.c_int32), ("ms_solver", c_double), ("ms_linear_op", c_double), ("ms_total", c_double)]


MESSAGE_FN = C.CFUNCTYPE(None, C.c_char_p, c_void_p)      # sgw_message_fn


class Cplx(C.Structure):
    _fields_ = [("re", c_double), ("im", c_double)]


class Freqbins(C.Structure):
    """sgw_freqbins == the members of freqbins_type (freqbins.f90:42-105) read by analytic.f90 / sigma.f90."""
    _fields_ = [("imag_sigma", C.c_int32), ("freq_symm_coul", C.c_int32), ("num_solver", C.c_int32), ("solver", c_void_p),
                ("num_coul", C.c_int32), ("coul", c_void_p), ("weight", c_void_p), ("num_sigma", C.c_int32),
                ("sigma", c_void_p)]


_lib = None


def load():
    """Load the CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        L = C.CDLL(str(LIB_PATH))
        L.sgw_last_error.restype = C.c_char_p
        L.sgw_last_error.argtypes = [c_void_p]
        L.sgw_create.argtypes = [c_int, C.POINTER(c_void_p)]
        L.sgw_destroy.argtypes = [c_void_p]
        L.sgw_set_stream.argtypes = [c_void_p, c_void_p]
        L.sgw_get_stats.argtypes = [c_void_p, C.POINTER(Stats)]
        L.sgw_set_profiling.argtypes = [c_void_p, c_int]
        L.sgw_device_synchronize.argtypes = [c_void_p]
        L.sgw_set_message_callback.argtypes = [c_void_p, MESSAGE_FN, c_void_p]
        L.sgw_get_profile.argtypes = [c_void_p, c_int, c_void_p, c_void_p, C.POINTER(c_int)]
        L.sgw_profile_class_name.restype = C.c_char_p
        L.sgw_profile_class_name.argtypes = [c_int]
        L.sgw_set_grid.argtypes = [c_void_p] + [c_int] * 6
        L.sgw_set_vloc.argtypes = [c_void_p, c_void_p]
        L.sgw_set_kpoint.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                     c_int, c_void_p, c_double]
        L.sgw_set_dense_operator.argtypes = [c_void_p, c_int, c_int, c_void_p, c_int]
        L.sgw_linear_op.argtypes = [c_void_p, c_int, c_int, c_void_p, c_double, c_void_p, c_int, c_void_p, c_int]
        L.sgw_solve_multishift.argtypes = [c_void_p, c_int, C.POINTER(SolverCfg), c_int, c_int, c_void_p, c_int, c_int,
                                           c_void_p, c_void_p, c_int64, c_int64, c_void_p]
        L.sgw_set_system.argtypes = [c_void_p, c_double, c_double, c_int, c_void_p, c_void_p]
        L.sgw_set_q.argtypes = [c_void_p, c_void_p]
        L.sgw_set_nksq.argtypes = [c_void_p, c_int]
        L.sgw_set_kpair.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_double]
        L.sgw_set_smearing.argtypes = [c_void_p, c_int, c_double, c_double, c_int]
        L.sgw_set_kpair_metal.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p]
        L.sgw_set_mixing.argtypes = [c_void_p, c_int, c_void_p, c_double, c_int]
        L.sgw_set_solve_direct.argtypes = [c_void_p, c_int]
        L.sgw_get_scf_iterations.argtypes = [c_void_p]
        L.sgw_solve_linter.argtypes = [c_void_p, C.POINTER(SolverCfg), c_int, c_void_p, c_int, c_void_p, c_void_p,
                                       c_void_p]
        L.sgw_coulomb.argtypes = [c_void_p, C.POINTER(SolverCfg), c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                                  c_void_p, c_void_p]
        L.sgw_get_rho_grid.argtypes = [c_void_p, c_void_p]
        L.sgw_coulomb_q0G0.argtypes = [c_void_p, C.POINTER(SolverCfg), c_int, c_void_p, c_void_p, c_void_p]
        L.sgw_unfold_w.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]
        L.sgw_invert_epsilon.argtypes = [c_void_p, c_int, c_int, c_void_p, c_int]
        L.sgw_green_function.argtypes = [c_void_p, c_int, C.POINTER(SolverCfg), c_int, c_void_p, c_int, c_void_p,
                                         c_int, c_void_p, c_void_p, c_void_p]
        L.sgw_parallel_task.argtypes = [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]
        L.sgw_bench_linear_op.argtypes = [c_void_p, c_int, c_int, c_int, C.POINTER(c_double), C.POINTER(c_double),
                                          C.POINTER(c_double)]
        L.sgw_freqbins_num_freq.argtypes = [C.POINTER(Freqbins)]
        L.sgw_pade_robust.argtypes = [c_void_p, c_double, c_int, c_void_p, C.POINTER(c_int), C.POINTER(c_int), c_void_p, c_void_p,
                                      c_double, c_double]
        L.sgw_unfold_w_symm.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p]
        L.sgw_aaa_pole_residual.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, C.POINTER(c_int)]
        L.sgw_coulpade.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p]
        L.sgw_analytic_coeff.argtypes = [c_void_p, c_int, c_double, C.POINTER(Freqbins), c_int, c_void_p]
        L.sgw_analytic_eval.argtypes = [c_void_p, c_int, C.POINTER(Freqbins), c_int, c_void_p, c_void_p, c_int, c_void_p,
                                        c_void_p]
        L.sgw_set_corr_grid.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_void_p]
        L.sgw_invfft6.argtypes = [c_void_p, c_double, c_void_p]
        L.sgw_fwfft6.argtypes = [c_void_p, c_double, c_void_p]
        L.sgw_sigma_correlation.argtypes = [c_void_p, c_int, C.POINTER(SolverCfg), c_double, c_double, Cplx, c_int,
                                            C.POINTER(Freqbins), c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
        _lib = L
    return _lib
