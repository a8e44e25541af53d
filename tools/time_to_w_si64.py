"""time-to-W(q, omega) for ONE full q-point of the Si64 synthetic (BASELINE.json configs[4], the north star's headline):
all `ngc` G-perturbations of the q (do_stern.f90:199-236 on one rank): operator tables host -> device, `coulomb` for every
perturbation (the library chunks them by free memory), gather, unfold_w, invert_epsilon, result eps^-1 - 1 on the host.

  python tools/time_to_w_si64.py [ngc] > gpurun_out/time_to_w_si64.json

The CPU side cannot be run in full (about a day on the box's 16 cores): it is extrapolated from the bounded oracle
sample `bench.py` times in the same run (solves/s), and labelled as an extrapolation.
"""
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402

import synth  # noqa: E402
from sternheimergw_b200 import Context, select_solver_type  # noqa: E402


def main():
    ngc = int(sys.argv[1]) if len(sys.argv) > 1 else 1900
    nfs = 32
    syn = synth.preset("si64")
    ngc = min(ngc, syn.ngm)
    fiu = synth.imag_freqs(nfs)
    igu = np.arange(1, ngc + 1, dtype=np.int32)
    cfg = select_solver_type(priority=(1, 3), threshold=1e-4)
    ctx = Context(0)
    ctx.install_system(syn)
    ctx.coulomb(cfg, 1, ngc, 8, igu, fiu)                 # warm-up: allocations, attribute opt-ins
    t0 = time.perf_counter()
    ctx.install_system(syn)                               # H2D of every table, as the Fortran host would
    t1 = time.perf_counter()
    scr = ctx.coulomb(cfg, 1, ngc, ngc, igu, fiu)
    st = ctx.stats()
    t2 = time.perf_counter()
    eps = ctx.unfold_w(ngc, igu, scr)
    t3 = time.perf_counter()
    w = ctx.invert_epsilon(eps)
    t4 = time.perf_counter()
    nocc, nshift = syn.nbnd_occ, 2 * nfs - 1
    solves = ngc * nocc * nshift
    # sanity: eps^-1 computed must invert eps (checked on the static frequency with numpy)
    e0 = eps[:, :, 0]
    resid = float(np.abs((w[:, :, 0] + np.eye(ngc)) @ e0 - np.eye(ngc)).max())
    rec = {"config": "Si64 synthetic, one full q-point", "fft_grid": list(syn.nr), "npw": int(syn.kpairs[0].kq.npw),
           "nbnd_occ": int(nocc), "ngc": int(ngc), "nfreq": nfs, "nshift": nshift, "solves": int(solves),
           "time_to_W_s": t4 - t0, "install_s": t1 - t0, "coulomb_s": t2 - t1, "unfold_s": t3 - t2, "invert_epsilon_s": t4 - t3,
           "coulomb_device_ms": st["ms_total"], "h_psi": int(st["n_linear_op"]), "gpu_launches": int(st["n_kernel_launch"]),
           "fallbacks": int(st["n_fallback"]), "solves_per_s": solves / (t2 - t1),
           "eps_inverse_residual_w0": resid, "eps_00_w0": [float(e0[0, 0].real), float(e0[0, 0].imag)],
           "rho_grid": list(ctx.rho_grid()[1])}
    print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
