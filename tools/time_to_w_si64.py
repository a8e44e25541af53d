"""time-to-W(q, omega) for ONE full q-point of the Si64 synthetic (BASELINE.json configs[4], the north star's headline):
all `ngc` G-perturbations of the q (do_stern.f90:199-236 on one rank): operator tables host -> device, `coulomb` for every
perturbation (the library chunks them by free memory), gather, unfold_w, invert_epsilon, result eps^-1 - 1 on the host.

  python tools/time_to_w_si64.py [ngc] > gpurun_out/time_to_w_si64.json
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/time_to_w_si64.py [ngc]
      N ranks = N GPUs: the perturbations are split with parallel_task's rule (do_stern.f90:199), every rank runs `coulomb`
      on its block, ONE all_gather collects the columns (do_stern.f90:211) and the root unfolds and inverts (strong scaling)

The CPU side cannot be run in full (about a day on the box's 16 cores): it is extrapolated from the bounded oracle
sample `bench.py` times in the same run (solves/s), and labelled as an extrapolation.
"""
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402

import synth  # noqa: E402
from sternheimergw_b200 import Context, select_solver_type  # noqa: E402


def main_dist(ngc, rank, world, local_rank):
    import os
    import torch
    import torch.distributed as dist
    from sternheimergw_b200.dist import do_stern_q
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    nfs = 32
    syn = synth.preset("si64")
    ngc = min(ngc, syn.ngm)
    fiu = synth.imag_freqs(nfs)
    igu = np.arange(1, ngc + 1, dtype=np.int32)
    cfg = select_solver_type(priority=(1, 3), threshold=1e-4)
    ctx = Context(local_rank)
    ctx.install_system(syn)
    ctx.coulomb(cfg, 1, ngc, 8, igu, fiu)                 # warm-up
    h_psi = [0]

    tcs = [0.0]

    def coulomb_fn(config, igstart, num_g_corr, num_task, ig_unique, fiu_):
        tc = time.perf_counter()
        scr = ctx.coulomb(config, igstart, num_g_corr, num_task, ig_unique, fiu_)
        tcs[0] = time.perf_counter() - tc
        h_psi[0] = int(ctx.stats()["n_linear_op"])
        return scr

    # two q-points: the first one also creates the communicator's buffers, the pinned staging areas and the workspace of the
    # full-size perturbation blocks, which every later q-point of a run re-uses; the second one is the figure
    runs = []
    for rep in range(2):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.install_system(syn)
        tm = {}
        w, (first, last, num_task) = do_stern_q(coulomb_fn, cfg, ngc, igu, fiu, unfold_fn=ctx.unfold_w,
                                                invert_fn=lambda s, lgamma=False: ctx.invert_epsilon(s, lgamma=lgamma), shard_invert=True,
                                                timings=tm)
        t_rank = time.perf_counter() - t0
        dist.barrier(); torch.cuda.synchronize()
        t1 = time.perf_counter()
        runs.append({"time_to_W_s": t1 - t0, "rank0_coulomb_s": tcs[0], **tm})
    tt = torch.tensor([t_rank], dtype=torch.float64, device="cuda")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        nocc, nshift = syn.nbnd_occ, 2 * nfs - 1
        solves = ngc * nocc * nshift
        rec = {"config": "Si64 synthetic, one full q-point, strong scaling over GPUs", "n_gpus": world, "ngc": int(ngc), "nfreq": nfs,
               "solves": int(solves), "time_to_W_s": t1 - t0, "first_q_point": runs[0], "second_q_point": runs[1],
               "max_rank_s": float(tt.item()), "tasks_per_rank": list(num_task),
               "solves_per_s": solves / (t1 - t0), "h_psi_rank0": h_psi[0],
               "eps_inv_minus_1_00_w0": [float(w[0, 0, 0].real), float(w[0, 0, 0].imag)],
               "note": "tables H2D on every rank + coulomb on the rank's block + all_to_all of the columns by frequency + unfold_w + invert_epsilon of every rank's frequencies + gather of W on the root; wall clock between two barriers, second q-point of the run"}
        print(json.dumps(rec), flush=True)
    dist.destroy_process_group()


def main():
    import os
    ngc = int(sys.argv[1]) if len(sys.argv) > 1 else 1900
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        return main_dist(ngc, int(os.environ.get("RANK", "0")), world, int(os.environ.get("LOCAL_RANK", "0")))
    nfs = 32
    syn = synth.preset("si64")
    ngc = min(ngc, syn.ngm)
    fiu = synth.imag_freqs(nfs)
    igu = np.arange(1, ngc + 1, dtype=np.int32)
    cfg = select_solver_type(priority=(1, 3), threshold=1e-4)
    ctx = Context(0)
    ctx.install_system(syn)
    ctx.coulomb(cfg, 1, ngc, 8, igu, fiu)                 # warm-up: allocations, attribute opt-ins
    t0 = time.perf_counter()
    ctx.install_system(syn)                               # H2D of every table, as the Fortran host would
    t1 = time.perf_counter()
    scr = ctx.coulomb(cfg, 1, ngc, ngc, igu, fiu)
    st = ctx.stats()
    t2 = time.perf_counter()
    eps = ctx.unfold_w(ngc, igu, scr)
    t3 = time.perf_counter()
    w = ctx.invert_epsilon(eps)
    t4 = time.perf_counter()
    nocc, nshift = syn.nbnd_occ, 2 * nfs - 1
    solves = ngc * nocc * nshift
    # sanity: eps^-1 computed must invert eps (checked on the static frequency with numpy)
    e0 = eps[:, :, 0]
    resid = float(np.abs((w[:, :, 0] + np.eye(ngc)) @ e0 - np.eye(ngc)).max())
    rec = {"config": "Si64 synthetic, one full q-point", "fft_grid": list(syn.nr), "npw": int(syn.kpairs[0].kq.npw),
           "nbnd_occ": int(nocc), "ngc": int(ngc), "nfreq": nfs, "nshift": nshift, "solves": int(solves),
           "time_to_W_s": t4 - t0, "install_s": t1 - t0, "coulomb_s": t2 - t1, "unfold_s": t3 - t2, "invert_epsilon_s": t4 - t3,
           "coulomb_device_ms": st["ms_total"], "h_psi": int(st["n_linear_op"]), "gpu_launches": int(st["n_kernel_launch"]),
           "fallbacks": int(st["n_fallback"]), "solves_per_s": solves / (t2 - t1),
           "eps_inverse_residual_w0": resid, "eps_00_w0": [float(e0[0, 0].real), float(e0[0, 0].imag)],
           "rho_grid": list(ctx.rho_grid()[1])}
    print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
