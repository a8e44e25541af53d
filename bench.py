#!/usr/bin/env python
"""bench.py -- shifted Sternheimer solves/s on the Si64 synthetic (BASELINE.json configs[4]), one JSON line.

A "step" = one pass of the hot path over one block of G-perturbations: ``sgw_coulomb`` (phys/coul/src/coulomb.f90:29)
for P perturbations x 128 occupied bands x 63 shifts (32 imaginary frequencies -> +-omega) of the 64-atom Si
supercell (72^3 FFT grid, npw ~ 24 k, nkb 256), production threshold 1e-4 (thres_coul default).  One solve = one
(right-hand side, shift) pair converged to the reference's criterion.  Inputs are synthetic (synth/, seed 20261017):
exact eigenvectors of the synthetic Hamiltonian the operator applies.

  python bench.py [--gpus N --steps K --warmup W]         B200 path (under torchrun for N > 1: one rank per GPU)
  python bench.py --impl reference [...]                   the CPU restatement of the reference (oracle/), all host cores

value      : whole-job solves/s, operator tables resident in HBM, device time (CUDA events on the library's stream)
e2e        : same metric through the C ABI with HOST buffers: every step re-installs all operator tables (H2D),
             runs sgw_coulomb and reads scrcoul back (D2H), wall clock around the blocking calls
roofline   : the ONE kernel class with the largest share of the timed region: its own algorithmic bytes / flops per launch
              / CUDA-event time / measured peak; `traffic` = ncu DRAM bytes per launch
time_to_W  : BASELINE's second metric on a fixed block: tables H2D + sgw_coulomb of NTW perturbations (split over the ranks,
             do_stern.f90:199) + gather + unfold_w + invert_epsilon + result on the host, wall clock (strong scaling over N)
parity     : in-run check of the timed workload against oracle/ (H.psi and one 63-shift solve at the bench threshold)
cpu_baseline: the oracle (kind "port": the Fortran reference cannot be built here) on a bounded sample, host cores
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "shifted Sternheimer solves/s"
UNIT = "solves/s"
NFS = 32               # imaginary frequencies -> 63 shifts (solve_linter.f90:238-252)
NGC = 1900             # G-perturbations of the q-point (~59 x 32)
THRESHOLD = 1e-4       # thres_coul default (main/src/gw_input.yml)
NTW = 64               # perturbations (= G vectors kept) of the fixed time-to-W block
DMMA_PIPE_TFLOPS = 37.05   # FP64 MMA pipe of one B200, register-only mma.sync loop (tools/micro/dmma_peak.cu, round 1)


def workload(name="si64"):
    import synth
    syn = synth.preset(name)
    fiu = synth.imag_freqs(NFS)
    ngc = min(NGC, syn.ngm)
    igu = np.arange(1, ngc + 1, dtype=np.int32)
    return syn, fiu, ngc, igu


def workload_config(syn, P, world):
    kq = syn.kpairs[0].kq
    return {"workload": "Si64 synthetic (BASELINE.json configs[4]): 64-atom Si supercell, FFT 72^3, 1 q / 1 k",
            "fft_grid": list(syn.nr), "npw": int(kq.npw), "nbnd_occ": int(syn.nbnd_occ), "nkb": int(kq.vkb.shape[1]),
            "nfreq": NFS, "nshift": 2 * NFS - 1, "perturbations_per_step_per_gpu": P, "threshold": THRESHOLD,
            "bicg_lmax": 4, "solver": "multishift BiCGStab(l), priority (1,3)",
            "parallelism": f"perturbation blocks over {world} GPU(s) (do_stern.f90:199), gather of eps columns per step",
            "l2": "inputs larger than L2 (solver state of one step is > 10 GB)"}


# ----------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock / throttle-reason samples DURING the timed region.  NVML is polled from a thread every 250 ms
    (a fast `nvidia-smi -lms` loop takes driver locks often enough to stall kernel launches of the process it
    watches: measured +60% on this launch-bound-free but sync-per-iteration workload); falls back to
    `nvidia-smi -lms 500` when pynvml is missing."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index, period=0.25):
        self.index, self.period = index, period
        self.sm, self.mx, self.pw, self.reasons = [], [], [], set()
        self.stop_flag = threading.Event()
        self.thread = self.proc = self.nv = None

    def prepare(self):
        """nvmlInit + the device handle, OUTSIDE the timed region: initialising NVML takes the driver's locks for tens of
        milliseconds and was measured to stretch the first timed step (400 vs 423 ms per step over 4 steps)."""
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.nv = pynvml
        except Exception:
            self.nv = None

    def start(self):
        if self.nv is None and not getattr(self, "_prepared", False):
            self.prepare()
        if self.nv is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "500"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.pw.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1e3)
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self.stop_flag.wait(self.period)

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.proc.stdout:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                self.sm.append(float(f[0])); self.mx.append(float(f[1])); self.pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    self.reasons.add(nm)

    def stop(self):
        self.stop_flag.set()
        if self.proc:
            time.sleep(0.1)
            self.proc.terminate()
        if self.thread:
            self.thread.join(timeout=2)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples (nvml / nvidia-smi unavailable)"]}
        return {"sm_mhz": float(np.median(self.sm)), "sm_min_mhz": float(min(self.sm)), "sm_max_mhz": max(self.mx),
                "power_w_median": float(np.median(self.pw)) if self.pw else None, "samples": len(self.sm),
                "source": "nvml" if self.nv else "nvidia-smi", "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------------------------------- CPU arm
def cpu_sample(syn, fiu, ngc, igu, steps, warmup, first_ig=2):
    """The oracle in the reference's execution order (one band, one vector at a time, unfused BLAS-1), OpenMP over
    bands on all host cores; each step = ONE perturbation x `cores` bands x 63 shifts (a bounded sample)."""
    import oracle
    oracle.build(native=True, force=True)          # -march=native for THIS host
    cores = len(os.sched_getaffinity(0))
    ps = oracle.PwSystem(syn, native=True)
    nb = min(cores, syn.nbnd_occ)
    ps.set_band_window(0, nb)
    cfg = oracle.make_cfg(priority=(1, 3), threshold=THRESHOLD)
    times = []
    nop = 0
    for s in range(warmup + steps):
        t = time.perf_counter()
        scr, ierr, st = ps.coulomb(first_ig + s, ngc, 1, igu, fiu, cfg, nthreads=cores)
        dt = time.perf_counter() - t
        if ierr != 0:
            raise RuntimeError(f"oracle did not converge (ierr={ierr})")
        if s >= warmup:
            times.append(dt)
            nop += st["n_op"]
    ps.set_band_window()
    solves = nb * (2 * NFS - 1)
    total = sum(times)
    return {"value": solves * steps / total, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{steps} step(s) of 1 perturbation x {nb} bands x {2 * NFS - 1} shifts each (of 128 bands), "
                      f"OpenMP over bands, {nop} H.psi, {total:.1f} s",
            "ms_per_step": 1e3 * total / steps, "solves_per_step": solves}


def run_reference(args, rank, world):
    if rank != 0:
        return
    syn, fiu, ngc, igu = workload()
    cb = cpu_sample(syn, fiu, ngc, igu, args.steps, args.warmup)
    cfg = workload_config(syn, 1, 1)
    cfg["parallelism"] = f"OpenMP over bands, {cb['cores']} host threads (the reference's MPI axes are images/pools)"
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "CPU restatement of the reference algorithm (oracle/), not gw.x: no Fortran compiler / QE in this image"}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------- B200 arm
def zgemm_peak_tflops():
    """FP64 tensor denominator: MEASURED_PEAKS.json has no FP64 entry, so measure cuBLAS ZGEMM here (burst)."""
    import torch
    n = 4096
    a = torch.randn(n, n, dtype=torch.complex128, device="cuda")
    b = torch.randn(n, n, dtype=torch.complex128, device="cuda")
    for _ in range(2):
        (a @ b)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); (a @ b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 8.0 * n ** 3 / (best * 1e-3) / 1e12


def h2d_bytes(syn, fiu, igu):
    n = syn.vrs.nbytes + syn.g.nbytes + syn.nl.nbytes + fiu.nbytes + igu.nbytes
    for kp in syn.kpairs:
        kq = kp.kq
        n += kq.nl_igk.nbytes + kq.g2kin.nbytes + kq.vkb.nbytes + kq.dion.nbytes + kq.evq.nbytes
        n += kp.evc.nbytes + kp.et.nbytes + kp.nl_igk_k.nbytes
    return int(n)


# ----------------------------------------------------------------------------------------------------- Sigma_c leg
def sigma_c_leg(device, f64_peak, reps=3):
    """SURVEY 8 f2/f3 measured: one sigma_correlation call (sigma.f90:528) at the sizes of examples/example01_Si
    (20^3 grid, npw ~ 283, 59 correlation G vectors on a 6^3 box, 51 integration frequencies -> 102 Green's-function
    frequencies, 35 solver frequencies -> Pade on 69 points, 11 self-energy frequencies).  The reference prints
    'G: 3.9-5.3 s  G*W: 7.0-7.9 s' per (k, q) configuration for this case (examples/example01_Si/gw.ref:10328)."""
    import synth
    from sternheimergw_b200 import Context, freqbins, select_solver_type
    from sternheimergw_b200.host import pade_approx
    syn = synth.preset("si", nk=1)
    kq = syn.kpairs[0].kq
    ngc, ncoul, nsig, nsolver = 59, 51, 11, 35
    nr_c, nl_c = synth.corr_grid(syn, ngc, nr=(6, 6, 6))
    nnr = int(np.prod(nr_c))
    pos = {int(g): i + 1 for i, g in enumerate(kq.igk)}
    map_ = np.array([pos.get(ig, 0) for ig in range(1, ngc + 1)], dtype=np.int32)
    mu = 0.5 * (kq.et[syn.nbnd_occ - 1] + kq.et[syn.nbnd_occ])
    fh = freqbins(True, 0.0, 100.0 / synth.RYTOEV, nsig, 200.0 / synth.RYTOEV, ncoul, synth.imag_freqs(nsolver))
    nsym = fh.num_freq()
    rng = np.random.default_rng(synth.SEED)
    poles = np.array([0.9, 1.7, 2.9])
    res = rng.standard_normal((ngc, ngc, 3)) * 0.05 + np.eye(ngc)[:, :, None]
    coul = np.zeros((ngc, ngc, nsym), complex, order="F")
    coul[:, :, :nsolver] = -(res[..., None] * 2 * poles[:, None] / (fh.solver ** 2 - poles[:, None] ** 2)).sum(axis=-2)
    gmapsym = np.arange(1, ngc + 1, dtype=np.int32)
    alpha = -1.0 / (2 * np.pi)
    ctx = Context(device)
    ctx.install_system(syn)
    ctx.set_corr_grid(nr_c, nl_c)
    ctx.set_profiling(True)
    cfg = select_solver_type(priority=(1, 3), threshold=1e-5)            # thres_green default
    coeff = ctx.analytic_coeff(pade_approx, 1e-4, fh, coul)
    nb = 2 * ncoul
    best = None
    for _ in range(reps):
        sig = np.zeros((ngc, ngc, nsig), complex, order="F")
        t0 = time.perf_counter()
        ctx.sigma_correlation(syn.omega_cell, cfg, 0, mu, alpha, pade_approx, fh, map_, gmapsym, coeff, sig)
        wall = 1e3 * (time.perf_counter() - t0)
        st, prof = ctx.stats(), ctx.profile()
        if best is None or wall < best["wall_ms"]:
            best = {"wall_ms": wall, "ms_total": st["ms_total"], "ms_green_solver": st["ms_solver"], "prof": prof,
                    "launches": int(st["n_kernel_launch"]), "linear_op": int(st["n_linear_op"])}
    gw = best["prof"].get("gw_product", {"ms": 0.0, "regions": 0})
    flop = 8.0 * nnr * nnr * ngc * nb * nsig
    out = {"workload": f"examples/example01_Si sizes: FFT 20^3, npw {kq.npw}, {ngc} correlation G on a {nr_c[0]}^3 box, {nb} Green "
                       f"frequencies, {nsym}-point Pade of W, {nsig} Sigma frequencies; one (k, q) configuration",
           "products": nsig * nb, "wall_ms": best["wall_ms"], "device_ms": best["ms_total"],
           "green_solver_ms": best["ms_green_solver"], "gpu_launches": best["launches"], "linear_op": best["linear_op"],
           "gw_product": {"kernel": "k_gw_product (second half of invfft6 + product with G(r,r') + sum over omega_green)",
                          "bound": "tensor", "ms": gw["ms"], "launches": gw["regions"],
                          "achieved": flop / (gw["ms"] * 1e-3) / 1e12 if gw["ms"] > 0 else None, "peak": f64_peak, "unit": "TFLOP/s",
                          "frac": flop / (gw["ms"] * 1e-3) / 1e12 / f64_peak if gw["ms"] > 0 else None,
                          "algorithmic_flop_per_launch": flop / max(1, gw["regions"])},
           "reference_2017": "G: 3.9-5.3 s, G*W: 7.0-7.9 s per (k, q) configuration (examples/example01_Si/gw.ref:10328, CPU of that run)"}
    del ctx
    # CPU oracle (numpy restatement of sigma_prod with per-column FFTs, fft6.f90) on a bounded sample of the products: the
    # only place of this leg that touches oracle/ (as the timed CPU baseline, never on the GPU path)
    from oracle import sigma as osg
    fo = osg.freqbins_type(fh.solver, np.asarray(fh.coul), np.asarray(fh.weight), np.asarray(fh.sigma), fh.freq_symm_coul, True)
    d = osg.corr_fft_type(tuple(nr_c), nl_c)
    green_r = np.asfortranarray(rng.standard_normal((nnr, nnr)) + 1j * rng.standard_normal((nnr, nnr)))
    nsample, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < 5.0:
        work = np.zeros((nnr, nnr), complex, order="F")
        work[:ngc, :ngc] = osg.analytic_eval(osg.PADE_APPROX, gmapsym, fo, coeff, 0.3j + 0.01 * nsample)
        osg.sigma_prod(syn.omega_cell, d, d, alpha, green_r, work)
        nsample += 1
    dt = time.perf_counter() - t0
    out["cpu_oracle"] = {"products_per_s": nsample / dt, "kind": "port", "cores": 1,
                         "sample": f"{nsample} analytic_eval + sigma_prod products in {dt:.1f} s (numpy, one thread)"}
    gw_ms = best["ms_total"] - best["ms_green_solver"]
    out["gw_products_per_s"] = nsig * nb / (gw_ms * 1e-3) if gw_ms > 0 else None
    out["note"] = "gw_products_per_s counts everything after the Green's-function solve: 6-D transform of G, analytic_eval, products, forward transforms"
    return out



# ----------------------------------------------------------------------------------------------------- time-to-W, parity
def time_to_w_block(ctx, syn, cfg, fiu, rank, world, barrier):
    """BASELINE metric (ii), driver-visible: one q-point that keeps NTW G vectors (NTW perturbations, all 32 frequencies,
    128 bands x 63 shifts each): operator tables host -> device on every rank, `coulomb` on the rank's block of
    perturbations (parallel_task's rule, do_stern.f90:199), gather of the eps columns (:211), unfold_w + invert_epsilon
    with the frequencies shared among the ranks, result eps^-1 - 1 on the root's host.  Total work is FIXED, so the
    figure at N GPUs against N = 1 is the strong scaling of time-to-W.  Wall clock between two barriers."""
    from sternheimergw_b200.dist import do_stern_q
    ngc = NTW
    igu = np.arange(1, ngc + 1, dtype=np.int32)
    t = {}
    if world > 1:        # warm the collectives this block uses (communicator setup is not part of a q-point's time)
        from sternheimergw_b200.dist import gather_columns, gather_frequencies
        gather_columns(np.zeros((4, 2, 1), complex, order="F"), [1] * world, all_ranks=True)
        gather_frequencies(np.zeros((4, 4, 1), complex, order="F"), [1] * world)
    barrier()
    t0 = time.perf_counter()
    ctx.install_system(syn)
    t["install_s"] = time.perf_counter() - t0
    stat = {}

    def coulomb_fn(config, igstart, num_g_corr, num_task, ig_unique, fiu_):
        tc = time.perf_counter()
        scr = ctx.coulomb(config, igstart, num_g_corr, num_task, ig_unique, fiu_)
        stat["coulomb_s"] = time.perf_counter() - tc
        stat["h_psi"] = int(ctx.stats()["n_linear_op"])
        return scr

    tm = {}
    w, (first, last, num_task) = do_stern_q(coulomb_fn, cfg, ngc, igu, fiu, unfold_fn=ctx.unfold_w,
                                            invert_fn=lambda a, lgamma=False: ctx.invert_epsilon(a, lgamma=lgamma),
                                            shard_invert=True, timings=tm)
    t_rank = time.perf_counter() - t0
    barrier()
    total = time.perf_counter() - t0
    if rank != 0:
        return None
    solves = ngc * syn.nbnd_occ * (2 * NFS - 1)
    out = {"what": f"one q-point keeping {ngc} G vectors: {ngc} perturbations x {syn.nbnd_occ} bands x {2 * NFS - 1} shifts, "
                   "tables H2D + coulomb + gather + unfold_w + invert_epsilon + eps^-1 - 1 on the host (fixed total work: strong scaling)",
           "n_gpus": world, "perturbations": ngc, "solves": int(solves), "seconds": total, "solves_per_s": solves / total,
           "rank0": {"install_s": t["install_s"], "coulomb_s": stat.get("coulomb_s"), **tm, "total_s": t_rank},
           "tasks_per_rank": [int(x) for x in num_task],
           "eps_inv_minus_1_00_w0": [float(w[0, 0, 0].real), float(w[0, 0, 0].imag)]}
    return out


def invert_epsilon_full(ctx, ngc=NGC, nfs=NFS):
    """invert_epsilon.f90:23 at the size of a full Si64 q-point (ngc x ngc x nfs), on a synthetic diagonally dominant eps."""
    rng = np.random.default_rng(3)
    eps = np.asfortranarray((rng.standard_normal((ngc, ngc, nfs)) + 1j * rng.standard_normal((ngc, ngc, nfs))) * (0.3 / np.sqrt(ngc)))
    eps[np.arange(ngc), np.arange(ngc), :] += 1.5
    t0 = time.perf_counter()
    w = ctx.invert_epsilon(eps)
    wall = time.perf_counter() - t0
    st = ctx.stats()
    dev_ms, gj_ms = st["ms_total"], st["ms_solver"]
    resid = float(np.abs((w[:, :, 1] + np.eye(ngc)) @ eps[:, :, 1] - np.eye(ngc)).max())
    flop = 8.0 * ngc ** 3 * nfs
    return {"ngc": ngc, "nfs": nfs, "wall_s": wall, "device_ms": dev_ms, "elimination_ms": gj_ms,
            "tflops_elimination": flop / (gj_ms * 1e-3) / 1e12 if gj_ms > 0 else None, "gpu_launches": int(st["n_kernel_launch"]),
            "residual_max": resid,
            "note": "blocked Gauss-Jordan (csrc/invert.cu); wall and device_ms include the H2D / D2H of the 2 x 1.85 GB matrices "
                    "from pageable host memory, elimination_ms is the factorisation alone (8 n^3 flop per matrix)"}


def parity_check(ctx, syn, cfg):
    """In-run parity of the timed workload against oracle/ (the C restatement of the reference): (a) linear_op on two random
    vectors, (b) one right-hand side of the bench step -- all 63 shifts, bench threshold -- through select_solver on both
    sides.  Tolerances: SURVEY 8d (1e-12; same outer-iteration count +-1 and 10 x threshold at production thresholds)."""
    import oracle
    import synth
    ps = oracle.PwSystem(syn)
    kq = syn.kpairs[0].kq
    rng = np.random.default_rng(synth.SEED)
    psi = np.zeros((kq.npwx, 2), dtype=complex, order="F")
    psi[:kq.npw] = rng.standard_normal((kq.npw, 2)) + 1j * rng.standard_normal((kq.npw, 2))
    om = np.array([0.13 + 0.2j, -0.4 + 0.05j])
    out = ctx.linear_op(0, om, kq.alpha_pv, psi)
    err_op = 0.0
    for v in range(2):
        ref = ps.linear_op(0, om[v], kq.alpha_pv, psi[:, v])
        err_op = max(err_op, float(np.abs(out[:, v] - ref).max() / np.abs(ref).max()))
    b = psi[:, :1] - kq.evq @ (kq.evq.conj().T @ psi[:, :1])
    b = np.asfortranarray(b / np.linalg.norm(b))
    fiu = synth.imag_freqs(NFS)
    omega = np.concatenate([fiu, -fiu[1:]])
    sigma = np.asfortranarray(-(kq.et[0] + omega).reshape(-1, 1))
    x, ierr = ctx.select_solver(cfg, 0, b, sigma)
    it_gpu = int(ctx.stats()["n_outer_max"])
    xo, ierr_o, so = ps.select_solver(0, b[:kq.npw, 0], sigma[:, 0], oracle.make_cfg(priority=(1, 3), threshold=THRESHOLD))
    err_x = float(np.abs(x[:kq.npw, :, 0] - xo).max() / np.abs(xo).max())
    ok = bool(err_op < 1e-12 and int(ierr[0]) == 0 and ierr_o == 0 and abs(it_gpu - so["n_outer"]) <= 1 and
              (err_x < 10 * THRESHOLD or it_gpu != so["n_outer"]))
    return {"ok": ok, "linear_op_rel_err": err_op, "linear_op_tol": 1e-12, "solve_rel_err_63_shifts": err_x,
            "solve_tol": 10 * THRESHOLD, "outer_iterations": {"gpu": it_gpu, "oracle": int(so["n_outer"])},
            "against": "oracle/ (C restatement of the reference, kind 'port'), same inputs, in this run"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pert", type=int, default=8, help="perturbations per step per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sigma", action="store_true", help="skip the Sigma_c (SURVEY 8 f2/f3) leg")
    ap.add_argument("--no-small", action="store_true", help="skip the time-to-W of the small BASELINE configs (gw_si, gw_c, gw_bn, gw_licl)")
    ap.add_argument("--sigma-only", action="store_true", help="run only the Sigma_c leg and print its record (profiling)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # Native libraries write to the process's stdout (NCCL prints its version banner there whenever NCCL_DEBUG is set on the
    # box, and NCCL_DEBUG_FILE does not cover that line): point fd 1 at stderr for the whole run and keep the original
    # stdout for the ONE JSON line rank 0 prints at the end.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    json_out = os.fdopen(json_fd, "w")

    import torch
    import torch.distributed as dist
    from sternheimergw_b200 import Context, select_solver_type
    from sternheimergw_b200.dist import gather_columns
    torch.cuda.set_device(local_rank)
    if args.sigma_only:
        print(json.dumps({"sigma_c": sigma_c_leg(local_rank, zgemm_peak_tflops())}), file=json_out, flush=True)
        return
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its debug lines (the "NCCL version ..." banner at level >= VERSION) to stdout by default: route them
        # to stderr so that rank 0's stdout carries exactly ONE JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    syn, fiu, ngc, igu = workload()
    P = args.pert
    ctx = Context(local_rank)
    ctx.install_system(syn)
    cfg = select_solver_type(priority=(1, 3), threshold=THRESHOLD)
    nocc, nshift = syn.nbnd_occ, 2 * NFS - 1
    solves_per_step = P * nocc * nshift
    num_task = [P] * world

    def igstart(s):
        return 1 + ((s * world + rank) * P) % max(1, ngc - P)

    def step(s, e2e=False):
        if e2e:
            ctx.install_system(syn)                       # H2D of every operator table, as a Fortran host would
        scr = ctx.coulomb(cfg, igstart(s), ngc, P, igu, fiu)            # raises if a solve did not converge
        st = ctx.stats()
        prof = ctx.profile()
        t_coll = 0.0
        if world > 1:                                     # do_stern.f90:211 gather of the eps columns (NCCL)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            gather_columns(scr, num_task)
            e1.record(); torch.cuda.synchronize()
            t_coll = e0.elapsed_time(e1)
        return scr, st, prof, t_coll

    ctx.set_profiling(False)
    for s in range(args.warmup):
        step(s)
    # ---- timed region: tables resident, per-kernel profiling OFF
    sampler = ClockSampler(local_rank)
    sampler.prepare()
    sampler._prepared = True
    barrier()
    sampler.start()
    wall0 = time.perf_counter()
    dev_ms, launches, nop, solver_ms, coll_ms, step_ms = 0.0, 0, 0, 0.0, 0.0, []
    for s in range(args.warmup, args.warmup + args.steps):
        scr, st, prof, t_coll = step(s)
        dev_ms += st["ms_total"] + t_coll
        coll_ms += t_coll
        step_ms.append(st["ms_total"])
        launches += st["n_kernel_launch"]
        solver_ms += st["ms_solver"]
        nop += st["n_linear_op"]
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - wall0)
    clocks = sampler.stop()
    # ---- the same steps once more with CUDA events around every launch of the library's stream (single stream):
    #      per-kernel-class device time for `kernels` / `roofline`
    ctx.set_profiling(True)
    prof_tot = {}
    prof_dev_ms = 0.0
    for s in range(args.warmup, args.warmup + args.steps):
        scr, st, prof, t_coll = step(s)
        prof_dev_ms += st["ms_total"]
        for k, v in prof.items():
            a = prof_tot.setdefault(k, {"ms": 0.0, "regions": 0})
            a["ms"] += v["ms"]; a["regions"] += v["regions"]
    ctx.set_profiling(False)
    # ---- e2e: host buffers in, host buffers out, every step
    barrier()
    e0 = time.perf_counter()
    for s in range(args.warmup, args.warmup + args.steps):
        step(s, e2e=True)
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - e0)

    # ---- time-to-W(q, omega) of a FIXED block (strong scaling): NTW perturbations of a q-point that keeps NTW G vectors
    # the first q-point of a run also allocates the workspace of its (larger) perturbation blocks -- tens of GB of cudaMalloc,
    # 0.3-0.7 s -- which every later q-point re-uses: the second call is the figure, the first is reported next to it
    ttw_first = time_to_w_block(ctx, syn, cfg, fiu, rank, world, barrier)
    ttw = time_to_w_block(ctx, syn, cfg, fiu, rank, world, barrier)
    if ttw is not None and ttw_first is not None:
        ttw["first_call_seconds"] = ttw_first["seconds"]
    my_step = float(np.mean(step_ms))
    t = torch.tensor([dev_ms, wall_ms, e2e_ms, coll_ms, my_step, -my_step], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms, e2e_ms, coll_ms, step_max, step_min = [float(x) for x in t.cpu()]
    step_min = -step_min
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_solves = solves_per_step * args.steps * world
    value = total_solves / (dev_ms * 1e-3)
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
    f64_peak = zgemm_peak_tflops()
    f64_src = "cuBLAS ZGEMM 4096^3 via torch.matmul(complex128), measured in this run (no FP64 entry in MEASURED_PEAKS.json)"
    kq = syn.kpairs[0].kq
    n, npw, m = kq.npwx, kq.npw, kq.vkb.shape[1] + nocc
    nnr = syn.nnr
    L, ns = 4, nshift - 1
    ncol = int(len(set(((kq.nl_igk - 1) % (syn.nr[0] * syn.nr[1])).tolist())))
    outer_rhs = nop / (2.0 * L)                      # sum over RHS of outer iterations they took part in
    nrhs_step = P * nocc * args.steps
    nsnap = L * (L + 1) // 2 + 1
    kbar = (nsnap + L) * outer_rhs / max(1, nrhs_step)   # basis vectors per right-hand side at the materialisation
    # ALGORITHMIC work of every kernel class over the timed steps; one unit = one H.psi (vector)
    # unless stated.  HBM classes in bytes, tensor classes in flop (8 flop per complex multiply-add: the cuBLAS convention).
    alg = {
        "fft_plane": ("hbm", nop * (2.0 * syn.nr[2] * ncol * 16), "GB/s"),
        "fft_zpass": ("hbm", nop * (2.0 * syn.nr[2] * ncol * 16 + 4 * 16.0 * n), "GB/s"),
        "gemm_project": ("tensor", nop * 8.0 * npw * m, "TFLOP/s"),
        "gemm_expand": ("tensor", nop * 8.0 * n * m, "TFLOP/s"),
        # materialisation of the +-omega averaged solutions: (n x K)(K x nfreq) per right-hand side (AvgSpec, csrc/bicgstab.cu)
        "shift_gemm": ("tensor", nrhs_step * 8.0 * n * kbar * NFS, "TFLOP/s"),
        # the streaming shifted update exists only with SGW_SHIFT=stream; in the default (lazy) mode the class holds the tiny
        # coefficient kernels and has no roofline
        **({} if prof_tot.get("shift_gemm", {"ms": 0})["ms"] > 0 else {"shift_fused": ("hbm", outer_rhs * (4.0 * ns + nsnap + L) * 16.0 * n, "GB/s")}),
        # seed recurrences: SURVEY 8d fused floor of the seed part, 3L(L+1)+4L + (3L+7) = 95 vector passes at L = 4
        "seed_blas1": ("hbm", outer_rhs * (3.0 * L * (L + 1) + 4 * L + 3 * L + 7) * 16.0 * n, "GB/s"),
    }
    # measured DRAM bytes per unit (ncu --set full, profiles/traffic.json)
    traffic = {}
    try:
        traffic = json.loads((ROOT / "profiles" / "traffic.json").read_text())
    except Exception:
        pass
    units_per_class = {"fft_plane": nop, "fft_zpass": 2 * nop, "gemm_project": nop, "gemm_expand": nop, "shift_fused": outer_rhs,
                       "seed_blas1": outer_rhs, "shift_gemm": nrhs_step}
    tot_prof = sum(v["ms"] for v in prof_tot.values()) or 1.0
    kernels = {}
    for k, v in prof_tot.items():
        ent = {"ms_per_step": v["ms"] / args.steps, "share_of_profiled": v["ms"] / tot_prof, "regions": v["regions"]}
        if k in alg and v["ms"] > 0:
            bound, work, unit = alg[k]
            ach = work / (v["ms"] * 1e-3) / (1e9 if unit == "GB/s" else 1e12)
            peak = hbm_peak if bound == "hbm" else f64_peak
            per_launch = units_per_class[k] / max(1, v["regions"])
            tr = traffic.get(k, {}).get("dram_bytes_per_unit")
            ent.update({"bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                        "peak_source": hbm_src if bound == "hbm" else f64_src,
                        "algorithmic_per_launch": work / max(1, v["regions"]), "avg_launch_ms": v["ms"] / max(1, v["regions"]),
                        "traffic": tr * per_launch if tr else None})
            if bound == "tensor":
                # the kernels use the 3-multiplication complex product: 6 real flop per complex multiply-add actually run on the pipe
                ent["dmma_pipe"] = {"achieved_real_tflops": 0.75 * ach, "peak": DMMA_PIPE_TFLOPS, "frac": 0.75 * ach / DMMA_PIPE_TFLOPS,
                                    "peak_source": "register-only mma.sync f64 loop, tools/micro/dmma_peak.cu (round 1)"}
            for extra in ("fp64_pipe_pct", "tensor_pipe_pct", "dram_pct", "smem_pipe_pct"):
                if extra in traffic.get(k, {}):
                    ent[extra + "_ncu"] = traffic[k][extra]
        kernels[k] = ent
    fft_ms = prof_tot.get("fft_plane", {"ms": 0})["ms"] + prof_tot.get("fft_zpass", {"ms": 0})["ms"]
    hpsi_fft = None
    if fft_ms > 0:
        own = alg["fft_plane"][1] + alg["fft_zpass"][1]          # the pipeline's own algorithmic bytes (sphere columns only)
        tr = sum(traffic.get(k, {}).get("dram_bytes_per_unit", 0.0) * (2 if k == "fft_zpass" else 1) for k in ("fft_plane", "fft_zpass"))
        hpsi_fft = {"kernel": "hpsi_fft_pipeline (k_zpass_g2r + k_plane_vloc + k_zpass_r2g, 3 launches per batch)", "bound": "hbm",
                    "us_per_vector": 1e3 * fft_ms / nop, "share_of_step": fft_ms / prof_dev_ms,
                    "achieved": own / (fft_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": own / (fft_ms * 1e-3) / 1e9 / hbm_peak,
                    "peak_source": hbm_src, "algorithmic_bytes_per_vector": own / nop,
                    "measured_dram_bytes_per_vector": tr or None,
                    "frac_on_measured_dram_bytes": (tr * nop / (fft_ms * 1e-3) / 1e9 / hbm_peak) if tr else None,
                    "unfused_3d_fft": {"bytes_per_vector": 192.0 * nnr + 32.0 * npw, "us_per_vector_at_hbm_peak": (192.0 * nnr + 32.0 * npw) / hbm_peak / 1e3,
                                       "note": "SURVEY 8d figure for an UNFUSED 3-D FFT (six 1-D passes over the full box); quoted as a "
                                               "time, not as a fraction: the fused sphere-pruned pipeline does not move those bytes"},
                    "note": "the pipeline is bound by shared memory / FP64 in k_plane_vloc, not by HBM: see kernels.fft_plane"}
    # `roofline` = the single kernel class with the largest share of the profiled step
    top = max((k for k in kernels if "frac" in kernels[k]), key=lambda k: kernels[k]["ms_per_step"])
    roofline = {"kernel": {"fft_plane": "k_plane_vloc (persistent 2-D FFT x v(r) x 2-D FFT per z-plane)", "gemm_project": "k_zgemm<1,1> (coef = P^H psi)",
                           "gemm_expand": "k_zgemm<0,0> (out = P coef)", "fft_zpass": "k_zpass_g2r / k_zpass_r2g", "shift_gemm": "k_shift_gemm",
                           "seed_blas1": "seed BLAS-1 kernels", "shift_fused": "k_shift_apply"}.get(top, top), "class": top,
                **{k: kernels[top][k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "peak_source")},
                "share_of_step": kernels[top]["ms_per_step"] / (prof_dev_ms / args.steps),
                "algorithmic_per_launch": kernels[top]["algorithmic_per_launch"], "avg_launch_ms": kernels[top]["avg_launch_ms"]}
    if "dmma_pipe" in kernels[top]:
        roofline["dmma_pipe"] = kernels[top]["dmma_pipe"]
    if top == "fft_plane":
        roofline["note"] = ("HBM fraction of the kernel's own algorithmic bytes (it reads and writes only the sphere columns of every z-plane); "
                            "the kernel is bound by shared-memory bandwidth and the FP64 pipe (ncu: *_ncu fields in kernels.fft_plane), "
                            "its DRAM traffic is 2.3 MB per vector")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(syn, P, world),
            "wall_ms_per_step": wall_ms / args.steps,
            "solves_per_step": solves_per_step * world, "linear_op_per_step": nop / args.steps,
            "coul_solver_ms_per_step": solver_ms / args.steps,
            "collective_ms_per_step": coll_ms / args.steps,
            "rank_step_ms": {"min": step_min, "max": step_max, "note": "mean device time of sgw_coulomb per step, slowest and fastest rank"},
            "step_ms_rank0": [round(x, 2) for x in step_ms],
            "profiled_ms_per_step": prof_dev_ms / args.steps,
            "profiled_note": "`kernels`/`roofline` come from a second pass over the same steps with CUDA events around every launch "
                             "of the library's stream; the timed pass runs without them",
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": total_solves / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes(syn, fiu, igu),
                    "d2h_bytes_per_step": int(ngc * NFS * P * 16 + 4), "ms_per_step": e2e_ms / args.steps,
                    "note": "host (pageable, caller-owned) arrays -> sgw_set_* + sgw_coulomb -> scrcoul on host, wall clock"},
            "roofline": roofline, "kernels": kernels, "hpsi_fft": hpsi_fft, "fp64_zgemm_peak_tflops": f64_peak,
            "time_to_W": ttw,
            "rho_grid": {"reduced": bool(ctx.rho_grid()[0]), "dims": list(ctx.rho_grid()[1]),
                         "note": "Delta-rho accumulated on the alias-free reduced box (sgw_get_rho_grid, DESIGN.md section 4)"}}
    if world == 1:
        try:
            ctx.release_workspace()
            line["invert_epsilon_full"] = invert_epsilon_full(ctx)
            ctx.release_workspace()
        except Exception as e:
            line["invert_epsilon_full"] = {"error": repr(e)}
        if not args.no_cpu_baseline:
            try:
                ctx.install_system(syn)
                line["parity"] = parity_check(ctx, syn, cfg)
            except Exception as e:
                line["parity"] = {"ok": False, "error": repr(e)}
    if world == 1 and not args.no_sigma:
        del ctx
        ctx = None
        try:
            line["sigma_c"] = sigma_c_leg(local_rank, f64_peak)
        except Exception as e:                                  # the headline line must survive a failure of the extra leg
            line["sigma_c"] = {"error": repr(e)}
    if world == 1 and not args.no_small and not args.no_cpu_baseline:
        # BASELINE.json configs[0..3] (the reference's own CPU-runnable cases): one q-point each, time-to-W on this GPU next to
        # the oracle on the box's cores in the same run, with the agreement of the two results (tools/config_times.py)
        del ctx
        ctx = None
        try:
            sys.path.insert(0, str(ROOT / "tools"))
            import config_times
            keep = ("config", "fft_grid", "nk", "ngc", "nshift", "solver_priority", "solves", "gpu_time_to_W_s", "gpu_launches",
                    "cpu_oracle_s", "cpu_cores", "speedup", "max_abs_diff_eps", "agrees_within_10_thr")
            line["small_configs"] = [{k: r[k] for k in keep if k in r} for r in config_times.main(do_cpu=True, device=local_rank, quiet=True)]
        except Exception as e:
            line["small_configs"] = {"error": repr(e)}
    if world == 1 and not args.no_cpu_baseline:
        del ctx
        cb = cpu_sample(syn, fiu, ngc, igu, steps=1, warmup=0)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
