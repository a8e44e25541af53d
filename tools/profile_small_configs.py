import sys, time, json
sys.path.insert(0, '.')
import numpy as np, synth
from sternheimergw_b200 import Context, select_solver_type
RY = 13.605698066
ctx = Context(0)
for name, preset, nk, ngc, fiu in [("gw_si", "si", 2, 59, np.array([0.0, 16j]) / RY), ("gw_c", "c", 2, 15, synth.imag_freqs(35)),
                                   ("gw_bn", "bn", 5, 39, np.array([0.0, 10j]) / RY)]:
    syn = synth.preset(preset, nk=nk)
    igu = np.arange(1, ngc + 1, dtype=np.int32)
    cfg = select_solver_type(priority=(1, 3), threshold=1e-4)
    ctx.install_system(syn)
    ctx.coulomb(cfg, 1, ngc, ngc, igu, fiu)
    for prof in (False, True):
        ctx.set_profiling(prof)
        t = time.perf_counter(); ctx.coulomb(cfg, 1, ngc, ngc, igu, fiu); w = time.perf_counter() - t
        st = ctx.stats()
        p = ctx.profile()
        print(name, "prof" if prof else "noprof", "wall %.1f ms" % (1e3 * w), "dev %.1f" % st["ms_total"], "solver %.1f" % st["ms_solver"], "launches", st["n_kernel_launch"],
              "kernel-sum %.1f" % sum(v["ms"] for v in p.values()), {k: round(v["ms"], 1) for k, v in p.items() if v["ms"] > 0}, flush=True)
    ctx.set_profiling(False)
