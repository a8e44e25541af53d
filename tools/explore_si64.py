"""Scratch exploration of the Si64 workload on the GPU box (not part of the product or the bench contract)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import synth
from sternheimergw_b200 import Context, select_solver_type

P = int(sys.argv[1]) if len(sys.argv) > 1 else 2
nfs = int(sys.argv[2]) if len(sys.argv) > 2 else 32
t = time.time(); syn = synth.preset("si64"); print("synth", time.time() - t, flush=True)
ctx = Context(0)
t = time.time(); ctx.install_system(syn); print("install", time.time() - t, flush=True)
for nvec in (128, 1024):
    print("bench_linear_op", nvec, ctx.bench_linear_op(0, nvec, reps=5), flush=True)
fiu = synth.imag_freqs(nfs)
ngc = 1900
igu = np.arange(1, ngc + 1, dtype=np.int32)
cfg = select_solver_type(priority=(1, 3), threshold=1e-4)
for rep in range(2):
    t = time.time()
    scr = ctx.coulomb(cfg, 1 + rep * P, ngc, P, igu, fiu, check=False)
    dt = time.time() - t
    st = ctx.stats()
    nsol = P * 128 * (2 * nfs - 1)
    print(f"coulomb P={P} nfs={nfs}: wall {dt:.3f}s  solves/s {nsol/dt:.0f}  stats {st}", flush=True)
print("eps diag", scr[:3, 0, 0])
