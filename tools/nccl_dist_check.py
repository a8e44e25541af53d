"""Multi-rank check ON HARDWARE of the three collectives of the path (NCCL over NVLink), with the real GPU pipeline on every rank:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/nccl_dist_check.py

1. do_stern_q (do_stern.f90:199-232): perturbations split by parallel_task's rule, `coulomb` on every rank's GPU, all_gather of
   the eps columns, frequency-sharded unfold_w + invert_epsilon, gather of W -- compared with the single-rank answer computed on
   rank 0.
2. sigma_wrapper_k (sigma.f90:319-362): (k, q) configurations dealt to the ranks, Sigma_c = G W of every configuration by
   sgw_sigma_correlation on the rank's GPU, mp_root_sum of Sigma(k, omega) by ONE ncclReduce -- compared with the serial sum.
3. coulomb_pools: the k-points of one q shared among the ranks (the reference's pools), Delta-rho summed by ncclAllReduce
   (solve_linter.f90:521) -- compared with all k-points on one rank.
Prints one JSON record (rank 0)."""
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402

import synth  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    from sternheimergw_b200 import Context, freqbins, select_solver_type
    from sternheimergw_b200.dist import coulomb_pools, do_stern_q, sigma_wrapper_k
    from sternheimergw_b200.host import pade_approx
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    rec = {"world": world, "backend": dist.get_backend()}
    ctx = Context(lr)
    # ---- 1. W(q, omega) of one q-point over the ranks
    syn = synth.preset("si", nk=2)
    ctx.install_system(syn)
    ngc = 27
    igu = np.arange(1, ngc + 1, dtype=np.int32)
    fiu = synth.imag_freqs(5)
    cfg = select_solver_type(priority=(1, 3), threshold=1e-10)
    tm = {}
    t0 = time.perf_counter()
    w, (first, last, num_task) = do_stern_q(ctx.coulomb, cfg, ngc, igu, fiu, unfold_fn=ctx.unfold_w,
                                            invert_fn=lambda a, lgamma=False: ctx.invert_epsilon(a, lgamma=lgamma),
                                            shard_invert=True, timings=tm)
    rec["do_stern_q"] = {"tasks_per_rank": [int(x) for x in num_task], "seconds": time.perf_counter() - t0, **tm}
    if rank == 0:
        scr = ctx.coulomb(cfg, 1, ngc, ngc, igu, fiu)
        ref = ctx.invert_epsilon(ctx.unfold_w(ngc, igu, scr))
        rec["do_stern_q"]["rel_err_vs_single_rank"] = float(np.abs(w - ref).max() / np.abs(ref).max())
    # ---- 1b. k-points shared among the ranks (pools): ncclAllReduce of the density response (solve_linter.f90:521)
    t0 = time.perf_counter()
    scr_pool = coulomb_pools(ctx, syn, cfg, 2, ngc, 5, igu, fiu)
    rec["coulomb_pools"] = {"k_points": len(syn.kpairs), "seconds": time.perf_counter() - t0}
    if rank == 0:
        ctx.install_system(syn)
        ref = ctx.coulomb(cfg, 2, ngc, 5, igu, fiu)
        rec["coulomb_pools"]["rel_err_vs_all_k_on_one_rank"] = float(np.abs(scr_pool - ref).max() / np.abs(ref).max())
    # ---- 2. Sigma_c(k, omega) summed over the ranks
    syn1 = synth.preset("si", nk=1)
    ctx.install_system(syn1)
    kq = syn1.kpairs[0].kq
    ngs, ncoul, nsig, nsolver = 27, 7, 3, 9
    nr_c, nl_c = synth.corr_grid(syn1, ngs)
    ctx.set_corr_grid(nr_c, nl_c)
    pos = {int(g): i + 1 for i, g in enumerate(kq.igk)}
    map_ = np.array([pos.get(ig, 0) for ig in range(1, ngs + 1)], dtype=np.int32)
    mu = 0.5 * (kq.et[syn1.nbnd_occ - 1] + kq.et[syn1.nbnd_occ])
    fh = freqbins(True, 0.0, 2.0, nsig, 6.0, ncoul, synth.imag_freqs(nsolver))
    nsym = fh.num_freq()
    rng = np.random.default_rng(synth.SEED)
    poles = np.array([0.9, 1.7, 2.9])
    res = rng.standard_normal((ngs, ngs, 3)) * 0.05 + np.eye(ngs)[:, :, None]
    coul = np.zeros((ngs, ngs, nsym), complex, order="F")
    coul[:, :, :nsolver] = -(res[..., None] * 2 * poles[:, None] / (fh.solver ** 2 - poles[:, None] ** 2)).sum(axis=-2)
    coeff = ctx.analytic_coeff(pade_approx, 1e-4, fh, coul)
    gcfg = select_solver_type(priority=(1, 3), threshold=1e-10)
    configs = [{"weight": 0.25 * (c + 1), "perm": np.random.default_rng(c).permutation(ngs)} for c in range(2 * world + 1)]

    def one(config, sigma):
        gm = (config["perm"] + 1).astype(np.int32)
        ctx.sigma_correlation(syn1.omega_cell, gcfg, 0, mu, -config["weight"] / (2 * np.pi), pade_approx, fh, map_, gm, coeff, sigma)

    t0 = time.perf_counter()
    sig, (first, last, num_task) = sigma_wrapper_k(one, configs, ngs, nsig)
    rec["sigma_wrapper_k"] = {"configurations_per_rank": [int(x) for x in num_task], "seconds": time.perf_counter() - t0}
    if rank == 0:
        ref = np.zeros((ngs, ngs, nsig), complex, order="F")
        for c in configs:
            one(c, ref)
        rec["sigma_wrapper_k"]["rel_err_vs_serial_sum"] = float(np.abs(sig - ref).max() / np.abs(ref).max())
        rec["ok"] = bool(rec["do_stern_q"]["rel_err_vs_single_rank"] < 1e-9 and rec["sigma_wrapper_k"]["rel_err_vs_serial_sum"] < 1e-12 and
                         rec["coulomb_pools"]["rel_err_vs_all_k_on_one_rank"] < 1e-9)
        print(json.dumps(rec), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
