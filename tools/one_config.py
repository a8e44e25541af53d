"""One `coulomb` call of a small BASELINE stand-in (for ncu launch lists): python tools/one_config.py si|c|bn|licl [ncalls]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, synth
from sternheimergw_b200 import Context, select_solver_type
RY = 13.605698066
CASES = {"si": ("si", 2, 59, np.array([0.0, 16j]) / RY, (1, 3)), "c": ("c", 2, 15, synth.imag_freqs(35), (1, 3)),
         "bn": ("bn", 5, 39, np.array([0.0, 10j]) / RY, (1, 3)), "licl": ("licl", 2, 6, (np.linspace(2.5, 12.5, 51) + 0.3j) / RY, (3,))}
preset, nk, ngc, fiu, prio = CASES[sys.argv[1]]
ncalls = int(sys.argv[2]) if len(sys.argv) > 2 else 1
syn = synth.preset(preset, nk=nk)
igu = np.arange(1, ngc + 1, dtype=np.int32)
cfg = select_solver_type(priority=prio, threshold=1e-4)
ctx = Context(0)
ctx.install_system(syn)
for _ in range(ncalls):
    t = time.perf_counter(); scr = ctx.coulomb(cfg, 1, ngc, ngc, igu, fiu); dt = time.perf_counter() - t
    st = ctx.stats()
    print("%s coulomb %.4f s, device %.2f ms, launches %d, H.psi %d, checksum %.12e" % (sys.argv[1], dt, st["ms_total"], st["n_kernel_launch"], st["n_linear_op"], np.abs(scr).sum()), flush=True)
