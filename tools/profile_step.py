"""One bench step of the Si64 workload for ncu (launch list / --set full captures); prints nothing a bench may quote.

  python tools/profile_step.py [P] [steps]     P perturbations per step (default 2), `steps` identical steps (default 2)

Identical steps (same perturbations) so that the launch list of step 2 can be cut out by launch count.
"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402

import synth  # noqa: E402
from sternheimergw_b200 import Context, select_solver_type  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
profiling = len(sys.argv) > 3 and sys.argv[3] == "prof"
syn = synth.preset("si64")
ctx = Context(0)
ctx.install_system(syn)
ctx.set_profiling(profiling)
fiu = synth.imag_freqs(32)
ngc = 1900
igu = np.arange(1, ngc + 1, dtype=np.int32)
cfg = select_solver_type(priority=(1, 3), threshold=1e-4)
for s in range(steps):
    t = time.time()
    ctx.coulomb(cfg, 2, ngc, P, igu, fiu)
    st = ctx.stats()
    if profiling:
        print({k: round(v["ms"], 1) for k, v in ctx.profile().items()})
    print(f"step {s}: wall {time.time() - t:.3f} s, launches {st['n_kernel_launch']}, ms_total {st['ms_total']:.1f}, "
          f"ms_solver {st['ms_solver']:.1f}, H.psi {st['n_linear_op']}", flush=True)
