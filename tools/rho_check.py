"""A/B of k_plane_rho_v2 against the generic k_plane_rho on the Si64 workload (SGW_RHO_V2=0/1): prints the differences."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, synth
from sternheimergw_b200 import Context, select_solver_type
syn = synth.preset("si64"); c = Context(0); c.install_system(syn)
fiu = synth.imag_freqs(3); igu = np.arange(1, 1901, dtype=np.int32)
out = {}
for v in ("0", "1", "0", "1"):
    os.environ["SGW_RHO_V2"] = v
    o = c.coulomb(select_solver_type(priority=(1, 3), threshold=1e-6), 5, 1900, 3, igu, fiu)
    if v in out:
        print("repeat", v, "identical to first run:", np.array_equal(out[v], o))
    out[v] = o
d = np.abs(out["0"] - out["1"])
print("rho grid", c.rho_grid())
print("max abs diff", d.max(), "max abs", np.abs(out["0"]).max(), "n differing", int((d > 0).sum()), "of", d.size)
