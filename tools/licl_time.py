"""gw_licl stand-in (subspace solver, 102 shifts): GPU time of one q-point, for tuning (SGW_SUB_THREADS, ...)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, synth
from sternheimergw_b200 import Context, select_solver_type
RY = 13.605698066
syn = synth.preset("licl", nk=2); ngc = 6; igu = np.arange(1, ngc + 1, dtype=np.int32)
fiu = (np.linspace(2.5, 12.5, 51) + 0.3j) / RY
cfg = select_solver_type(priority=(3,), threshold=1e-4)
ctx = Context(0); ctx.install_system(syn); ctx.coulomb(cfg, 1, ngc, ngc, igu, fiu)
t = time.perf_counter(); scr = ctx.coulomb(cfg, 1, ngc, ngc, igu, fiu); dt = time.perf_counter() - t
st = ctx.stats()
print("licl coulomb %.3f s, device %.1f ms, launches %d, H.psi %d, checksum %.12e" % (dt, st["ms_total"], st["n_kernel_launch"], st["n_linear_op"], np.abs(scr).sum()))
