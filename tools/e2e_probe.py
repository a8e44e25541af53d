"""Where does the host-side table installation (the H2D part of bench.py's e2e leg) spend its time?  Prints per-call wall
times of the sgw_set_* sequence for a few repetitions on the Si64 workload."""
import ctypes as C
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402

import synth  # noqa: E402
from sternheimergw_b200 import Context, select_solver_type  # noqa: E402
from sternheimergw_b200.host import _c16, _p  # noqa: E402

syn = synth.preset("si64")
fiu = synth.imag_freqs(32)
ngc = 1900
igu = np.arange(1, ngc + 1, dtype=np.int32)
cfg = select_solver_type(priority=(1, 3), threshold=1e-4)
ctx = Context(0)
ctx.install_system(syn)
ctx.coulomb(cfg, 1, ngc, 8, igu, fiu)
L, h = ctx._L, ctx._h
for rep in range(6):
    t = [time.perf_counter()]
    ctx.set_grid(*syn.nr); t.append(time.perf_counter())
    ctx.set_vloc(syn.vrs); t.append(time.perf_counter())
    L.sgw_set_system(h, syn.omega_cell, syn.tpiba2, syn.ngm, _p(np.ascontiguousarray(syn.g.T, dtype=np.float64)),
                     _p(np.ascontiguousarray(syn.nl, dtype=np.int32))); t.append(time.perf_counter())
    ctx.set_q(syn.xq); L.sgw_set_nksq(h, 1); t.append(time.perf_counter())
    kp = syn.kpairs[0]; kq = kp.kq
    ctx.set_kpoint(0, kq.npw, kq.npwx, kq.nl_igk, kq.g2kin, kq.vkb, kq.dion, kq.evq, kq.alpha_pv); t.append(time.perf_counter())
    evc = _c16(kp.evc); et = np.ascontiguousarray(kp.et); nl = np.ascontiguousarray(kp.nl_igk_k, dtype=np.int32)
    L.sgw_set_kpair(h, 0, 0, kp.npw_k, _p(nl), evc.shape[1], _p(evc), _p(et), float(kp.wk)); t.append(time.perf_counter())
    ctx.coulomb(cfg, 1 + 8 * rep, ngc, 8, igu, fiu); t.append(time.perf_counter())
    names = ["set_grid", "set_vloc", "set_system", "set_q/nksq", "set_kpoint", "set_kpair", "coulomb"]
    print(rep, " ".join(f"{n} {1e3 * (b - a):.1f}" for n, a, b in zip(names, t[:-1], t[1:])), flush=True)
