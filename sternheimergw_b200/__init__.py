"""sternheimergw_b200 -- B200-native (sm_100a) Sternheimer linear-response hot path of SternheimerGW.

Host-side mirror (Python, over the C ABI of include/sgw_b200.h) of the reference interface for this path:

    select_solver_type / select_solver   algo/linear_solver/src/select_solver.f90:48,67
    linear_op                            algo/linear_solver/src/linear_op.f90:46
    solve_linter                         phys/coul/src/solve_linter.f90:55
    coulomb / coulomb_q0G0               phys/coul/src/coulomb.f90:29, coulomb_q0G0.f90:31
    unfold_w / invert_epsilon            algo/symmetry/src/unfold_w.f90:23, phys/coul/src/invert_epsilon.f90:23
    green_function                       phys/green/src/green.f90:105
    parallel_task                        data/parallel/src/parallel.f90:80
    freqbins_type / freqbins             algo/grid/src/freqbins.f90:42,109 (+ gauleg_grid.f90)
    coulpade                             phys/coul/src/coulpade.f90:36
    analytic_coeff / analytic_eval       algo/analytic/src/analytic.f90:50,211 (all five model_coul values)
    invfft6 / fwfft6                     data/fft/src/fft6.f90:231,84
    sigma_correlation                    phys/corr/src/sigma.f90:528

Everything computes on the GPU through libsgw_b200.so; there is no CPU fallback -- importing works
without a GPU, creating a `Context` does not.
"""
from .host import Context, SgwError, freqbins, freqbins_type, parallel_task, select_solver_type  # noqa: F401

__all__ = ["Context", "SgwError", "select_solver_type", "parallel_task", "freqbins_type", "freqbins"]
