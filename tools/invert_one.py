"""One sgw_invert_epsilon at the size of a full Si64 q-point (for ncu captures of the blocked Gauss-Jordan kernels)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from sternheimergw_b200 import Context  # noqa: E402

r = bench.invert_epsilon_full(Context(0), nfs=int(sys.argv[1]) if len(sys.argv) > 1 else 32)
print({k: r[k] for k in ("elimination_ms", "tflops_elimination", "gpu_launches", "residual_max")})
