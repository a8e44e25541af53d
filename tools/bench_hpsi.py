"""Micro-benchmark of the batched H.psi pipeline at the Si64 size (sgw_bench_linear_op): per-kernel-class device times
(CUDA events around every launch) for a batch of `nvec` vectors.  Tuning knobs come from the environment (SGW_PLANE, ...).

  python tools/bench_hpsi.py [nvec] [reps] [preset]
"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import synth  # noqa: E402
from sternheimergw_b200 import Context  # noqa: E402

nvec = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
name = sys.argv[3] if len(sys.argv) > 3 else "si64"
syn = synth.preset(name)
ctx = Context(0)
ctx.install_system(syn)
out = ctx.bench_linear_op(0, nvec, reps)            # unprofiled totals
ctx.set_profiling(True)
ctx.bench_linear_op(0, nvec, reps)
prof = {k: round(v["ms"] / v["regions"], 4) for k, v in ctx.profile().items() if v["regions"]}
print(json.dumps({"preset": name, "nvec": nvec, **{k: round(v, 4) for k, v in out.items()}, "per_launch_ms": prof}))
