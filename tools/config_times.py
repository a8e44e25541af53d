"""time-to-W(q, omega) of the BASELINE.json parity configs (synthetic stand-ins of gw_si / gw_c / gw_bn / gw_licl, SURVEY
section 8 table) on the GPU, next to the CPU oracle on the same box and the reference's own published clocks.

  python tools/config_times.py [--no-cpu] > gpurun_out/config_times.json

One q-point each: `coulomb` for all ngc perturbations of that q (do_stern.f90:209), then unfold_w + invert_epsilon.
The GPU result is checked against the oracle (production thresholds: both stop at the same criterion).
"""
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402

import oracle  # noqa: E402
import synth  # noqa: E402
from sternheimergw_b200 import Context, select_solver_type  # noqa: E402

RY = 13.605698066
CASES = [
    # name, preset, nk, ngc, frequencies (Ry), priority, threshold, reference clock for context
    ("gw_si", "si", 2, 59, np.array([0.0, 16j]) / RY, (1, 3), 1e-4,
     "reference: coulomb 17.57 s for 4 q (107 perturbations), 1 rank, test-suite/gw_si/benchmark.out.v0.11.inp=gw.in:3406"),
    ("gw_c", "c", 2, 15, synth.imag_freqs(35), (1, 3), 1e-4,
     "reference: coulomb 7.48 s for 3 q (19 perturbations), 1 rank, test-suite/gw_c/benchmark.out.v0.13.inp=gw.in:3801"),
    ("gw_bn", "bn", 5, 39, np.array([0.0, 10j]) / RY, (1, 3), 1e-4,
     "reference: coulomb 141.89 s for 5 q (91 perturbations), 1 rank, test-suite/gw_bn/benchmark.out.v0.13.inp=gw.in:5267"),
    ("gw_licl", "licl", 2, 6, (np.linspace(2.5, 12.5, 51) + 0.3j) / RY, (3,), 1e-4,
     "reference: coul_solver 62.29 s for 1 q (6 perturbations, subspace solver), 1 rank, test-suite/gw_licl/benchmark.out.v0.11.inp=gw.in:1312"),
]


def main(do_cpu=None, device=0, quiet=False):
    if do_cpu is None:
        do_cpu = "--no-cpu" not in sys.argv
    cores = len(os.sched_getaffinity(0))
    ctx = Context(device)
    out = []
    for name, preset, nk, ngc, fiu, prio, thr, ref in CASES:
        syn = synth.preset(preset, nk=nk)
        ngc = min(ngc, syn.ngm)
        igu = np.arange(1, ngc + 1, dtype=np.int32)
        cfg = select_solver_type(priority=prio, threshold=thr)
        ctx.install_system(syn)
        ctx.coulomb(cfg, 1, ngc, ngc, igu, fiu)                       # warm-up (allocations)
        t = time.perf_counter()
        ctx.install_system(syn)
        scr = ctx.coulomb(cfg, 1, ngc, ngc, igu, fiu)
        st = ctx.stats()
        eps = ctx.invert_epsilon(ctx.unfold_w(ngc, igu, scr))
        gpu_s = time.perf_counter() - t
        kq = syn.kpairs[0].kq
        nshift = 2 * fiu.size - (1 if abs(fiu[0]) < 1e-14 else 0)
        rec = {"config": name, "fft_grid": list(syn.nr), "npw": int(kq.npw), "nk": len(syn.kpairs), "nbnd_occ": int(syn.nbnd_occ),
               "ngc": int(ngc), "nshift": int(nshift), "solver_priority": list(prio),
               "solves": int(ngc * len(syn.kpairs) * syn.nbnd_occ * nshift),
               "gpu_time_to_W_s": gpu_s, "gpu_device_ms": st["ms_total"], "gpu_h_psi": int(st["n_linear_op"]),
               "gpu_launches": int(st["n_kernel_launch"]), "gpu_fallbacks": int(st["n_fallback"]), "reference_context": ref}
        if do_cpu:
            ps = oracle.PwSystem(syn)
            t = time.perf_counter()
            ref_scr, ierr, so = ps.coulomb(1, ngc, ngc, igu, fiu, oracle.make_cfg(priority=prio, threshold=thr), nthreads=cores)
            cpu_s = time.perf_counter() - t
            err = float(np.abs(scr - ref_scr).max())
            rec.update({"cpu_oracle_s": cpu_s, "cpu_cores": cores, "cpu_h_psi": int(so["n_op"]), "cpu_ierr": int(ierr),
                        "speedup": cpu_s / gpu_s, "max_abs_diff_eps": err, "agrees_within_10_thr": bool(err < 10 * thr * 10)})
        rec["eps_inv_00_w0"] = [float(eps[0, 0, 0].real), float(eps[0, 0, 0].imag)]
        out.append(rec)
        if not quiet:
            print(json.dumps(rec), flush=True)
    ctx.close()
    return out


if __name__ == "__main__":
    main()
