"""q-point level driver: the part of do_stern (phys/coul/src/do_stern.f90:189-236) that distributes the
G-perturbations over ranks and gathers the dielectric-matrix columns.

The reference splits ``ngmunique`` perturbations over MPI images with ``parallel_task`` (do_stern.f90:199), every image
runs ``coulomb`` on its contiguous block (:209) and ``mp_gatherv`` collects ``scrcoul_loc(ngc, nfs, ntask)`` on the
root image (:211), which unfolds, patches the head, inverts and writes (:220-236).  Here one rank = one GPU; the
blocks are independent, so there is NO data-path collective during the solves; the gather is a single
``torch.distributed`` call per q (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np

from .host import parallel_task


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def _device(dist, device):
    import torch
    if device is not None:
        return device
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


_PINNED = {}


def _pinned(key, shape):
    """Page-locked staging buffers are expensive to create (cudaHostAlloc): keep one per (role, shape)."""
    import torch
    t = _PINNED.get((key, shape))
    if t is None:
        t = torch.zeros(shape, dtype=torch.float64, pin_memory=True)
        _PINNED[(key, shape)] = t
    return t


def _gather_blocks(blocks_c: np.ndarray, counts, root: int, all_ranks: bool, device=None):
    """Gather of per-rank blocks that are C-contiguous along their first axis (counts[r] leading entries on rank r): ONE
    padded all_gather (or gather) of raw float64 pairs, no transposes and no per-element Python work; returns the concatenated
    (sum(counts), ...) complex array on the root (on every rank with all_ranks), None elsewhere."""
    import torch
    dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    tail = blocks_c.shape[1:]
    per = int(np.prod(tail)) if tail else 1
    nmax = max(counts)
    dev = _device(dist, device)
    pin = dev.type == "cuda"
    mine_h = _pinned("send", (nmax, per, 2)) if pin else torch.zeros((nmax, per, 2), dtype=torch.float64)
    if blocks_c.shape[0]:
        mine_h[:blocks_c.shape[0]] = torch.from_numpy(np.ascontiguousarray(blocks_c).view(np.float64).reshape(blocks_c.shape[0], per, 2))
    mine = mine_h.to(dev, non_blocking=True)
    if all_ranks or dist.get_backend() == "nccl":
        full = torch.empty((world, nmax, per, 2), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(full.view(-1), mine.view(-1))
        have = all_ranks or rank == root
    else:
        parts = [torch.empty_like(mine) for _ in range(world)] if rank == root else None
        dist.gather(mine, parts, dst=root)
        full = torch.stack(parts) if rank == root else None
        have = rank == root
    if not have:
        return None
    if pin:
        recv = _pinned("recv", (world, nmax, per, 2))
        recv.copy_(full, non_blocking=False)
        full_h = recv.numpy().view(np.complex128).reshape((world, nmax) + tail)
    else:
        full_h = full.cpu().numpy().view(np.complex128).reshape((world, nmax) + tail)
    if all(int(c) == nmax for c in counts):                # equal blocks: the received buffer already is the concatenation
        return full_h.reshape((world * nmax,) + tail)
    out = np.empty((int(sum(counts)),) + tail, dtype=np.complex128)
    off = 0
    for r in range(world):
        n = counts[r]
        if n:
            out[off:off + n] = full_h[r, :n]
        off += n
    return out


def gather_columns(scr_loc: np.ndarray, num_task, root: int = 0, device=None, all_ranks: bool = False):
    """mp_gatherv(inter_image_comm, root, num_task, scrcoul_loc, scrcoul_root) (do_stern.f90:211, parallel.f90:1130).

    scr_loc: (ngc, nfs, ntask_loc) complex128 of this rank; num_task: tasks of every rank.  Returns the
    (ngc, nfs, sum(num_task)) array on ``root`` (on every rank with ``all_ranks``) and None elsewhere.  Without an initialised
    process group (single rank) the input is returned unchanged.  A Fortran-ordered (ngc, nfs, ntask) array IS a C-ordered
    (ntask, nfs, ngc) one, so the blocks travel as they lie in memory.
    """
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return scr_loc
    blocks = np.ascontiguousarray(np.asfortranarray(scr_loc).T)          # a view for Fortran-ordered input
    out = _gather_blocks(blocks, [int(x) for x in num_task], root, all_ranks, device)
    return None if out is None else out.T                                 # (ngc, nfs, ntot), Fortran order, no copy


def exchange_frequencies(scr_loc: np.ndarray, num_task, num_freq, device=None):
    """Columns of all perturbations for THIS rank's share of the frequencies: rank r holds scr_loc(ngc, nfs, num_task[r]) and
    receives (ngc, num_freq[rank], sum(num_task)) -- one all_to_all (ncclSend/Recv pairs over NVLink on the GPU box) that moves
    1/world of what an all_gather of every column to every rank would.  The pieces are C-contiguous (ntask, nf, ngc) blocks, so
    the receive buffer IS the result: no assembly copy."""
    import torch
    dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = _device(dist, device)
    ngc, nfs, ntl = scr_loc.shape
    blocks = np.asfortranarray(scr_loc).T                                  # (ntask_loc, nfs, ngc), C order, a view
    f_off = np.concatenate([[0], np.cumsum(num_freq)]).astype(int)
    nf_me = int(num_freq[rank])
    send = np.empty(ntl * nfs * ngc, dtype=np.complex128)
    in_splits, off = [], 0
    for d in range(world):
        piece = blocks[:, f_off[d]:f_off[d + 1], :]
        send[off:off + piece.size] = piece.reshape(-1)
        in_splits.append(2 * piece.size)
        off += piece.size
    out_splits = [2 * int(num_task[r]) * nf_me * ngc for r in range(world)]
    ts = torch.from_numpy(send.view(np.float64)).to(dev)
    tr = torch.empty(sum(out_splits), dtype=torch.float64, device=dev)
    dist.all_to_all_single(tr, ts, out_splits, in_splits)
    recv = tr.cpu().numpy().view(np.complex128).reshape(int(sum(num_task)), nf_me, ngc)
    return recv.T                                                          # (ngc, nf_me, ntot), Fortran order


def gather_frequencies(w_loc: np.ndarray, num_freq, root: int = 0, device=None):
    """Gather of the frequency slices (ngc, ngc, nfs_loc) every rank inverted into (ngc, ngc, sum(num_freq)) on ``root``."""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return w_loc
    blocks = np.ascontiguousarray(np.asfortranarray(w_loc).T)
    out = _gather_blocks(blocks, [int(x) for x in num_freq], root, False, device)
    return None if out is None else out.T


def do_stern_q(coulomb_fn, config, num_g_corr, ig_unique, fiu, unfold_fn=None, invert_fn=None, eps_head=None,
               lgamma=False, root: int = 0, shard_invert: bool = False, timings: dict | None = None):
    """One q-point of do_stern: split -> coulomb on the local block -> gather -> (root) unfold, head, invert.

    coulomb_fn(config, igstart, num_g_corr, num_task, ig_unique, fiu) -> (ngc, nfs, num_task) is normally
    ``Context.coulomb``; unfold_fn / invert_fn are ``Context.unfold_w`` / ``Context.invert_epsilon``.
    Returns (scrcoul_g on root | None, (first_task, last_task, num_task)).
    """
    dist = _dist()
    rank = dist.get_rank() if dist else 0
    world = dist.get_world_size() if dist else 1
    ngmunique = len(ig_unique)
    first, last, num_task = parallel_task(world, rank, ngmunique)                 # do_stern.f90:199
    ntask_loc = num_task[rank]
    if ntask_loc > 0:
        scr_loc = coulomb_fn(config, first, num_g_corr, ntask_loc, ig_unique, fiu)   # :209
    else:
        scr_loc = np.zeros((num_g_corr, len(fiu), 0), dtype=np.complex128, order="F")
    import time
    t0 = time.perf_counter()
    if shard_invert and world > 1 and unfold_fn is not None:
        # The reference unfolds and inverts all frequencies on the root image (do_stern.f90:220-232) -- a serial tail of
        # O(ngc^3 nfs) that holds strong scaling back.  The frequencies are independent, so every rank takes a contiguous
        # share of them (parallel_task's rule again), unfolds and inverts its slices on its own GPU and the root gathers.
        f_first, f_last, num_freq = parallel_task(world, rank, len(fiu))
        sl = slice(f_first - 1, f_first - 1 + num_freq[rank])
        scr_me = exchange_frequencies(scr_loc, num_task, num_freq)           # this rank's frequencies of every column
        t1 = time.perf_counter()
        if num_freq[rank] > 0:
            w_loc = unfold_fn(num_g_corr, ig_unique, scr_me)
            if eps_head is not None:
                w_loc[0, 0, :] = np.asarray(eps_head)[sl]
            if invert_fn is not None:
                w_loc = invert_fn(w_loc, lgamma=lgamma)
        else:
            w_loc = np.zeros((num_g_corr, num_g_corr, 0), dtype=np.complex128, order="F")
        t2 = time.perf_counter()
        out = gather_frequencies(w_loc, num_freq, root=root)
        if timings is not None:
            timings.update(gather_s=t1 - t0, unfold_invert_s=t2 - t1, gather_w_s=time.perf_counter() - t2)
        return out, (first, last, num_task)
    scr_root = gather_columns(scr_loc, num_task, root=root)                        # :211
    t1 = time.perf_counter()
    if timings is not None:
        timings.update(gather_s=t1 - t0)
    if rank != root or unfold_fn is None:
        return (scr_root if rank == root else None), (first, last, num_task)
    scr_g = unfold_fn(num_g_corr, ig_unique, scr_root)                             # :220 (identity symmetry)
    if eps_head is not None:
        scr_g[0, 0, :] = eps_head                                                  # :224
    if invert_fn is not None:
        scr_g = invert_fn(scr_g, lgamma=lgamma)                                    # :232
    if timings is not None:
        timings.update(unfold_invert_s=time.perf_counter() - t1)
    return scr_g, (first, last, num_task)


def pool_sum_eps(scr_loc: np.ndarray, igstart: int, ig_unique, device=None):
    """mp_sum(drhoscf, inter_pool_comm) of solve_linter.f90:521 for k-points shared among ranks (the reference's pools).

    Every rank runs ``coulomb`` for the SAME perturbations with only its share of the (k, k+q) pairs installed
    (``Context.install_system(syn, kpairs=...)``, weights wk unchanged).  The eps column the library returns is affine in the
    density response, scrcoul = delta - v_c Delta-rho (coulomb.f90:143-157), and Delta-rho is a sum over k, so the allreduce
    can act on the columns: sum over ranks, then remove the (world - 1) surplus copies of the delta on the perturbation's own G.
    scr_loc: (ngc, nfs, ntask) of this rank, tasks igstart .. igstart + ntask - 1 of ig_unique (1-based).  Returns the full
    columns on every rank."""
    import torch
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return scr_loc
    world = dist.get_world_size()
    dev = _device(dist, device)
    a = np.asfortranarray(scr_loc, dtype=np.complex128)
    t = torch.from_numpy(np.ascontiguousarray(a.T).view(np.float64).copy()).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)                       # ncclAllReduce over NVLink on the GPU box
    out = np.ascontiguousarray(t.cpu().numpy()).view(np.complex128).reshape(a.T.shape).T
    out = np.asfortranarray(out)
    for it in range(out.shape[2]):
        ig = int(ig_unique[igstart - 1 + it])
        if ig <= out.shape[0]:
            out[ig - 1, :, it] -= float(world - 1)
    return out


def coulomb_pools(ctx, syn, config, igstart: int, num_g_corr: int, num_task: int, ig_unique, fiu):
    """`coulomb` with the k-points of `syn` shared among the ranks like the reference's pools (k loop of solve_linter.f90:300,
    mp_sum over inter_pool_comm at :521): rank r installs its contiguous share of the (k, k+q) pairs (parallel_task's rule),
    runs the same perturbations and ``pool_sum_eps`` adds the density responses.  A rank without k-points contributes the
    bare delta.  Returns the full eps columns on every rank."""
    dist = _dist()
    rank = dist.get_rank() if dist else 0
    world = dist.get_world_size() if dist else 1
    first, last, num = parallel_task(world, rank, len(syn.kpairs))
    if num[rank] > 0:
        ctx.install_system(syn, kpairs=list(range(first - 1, first - 1 + num[rank])))
        scr = ctx.coulomb(config, igstart, num_g_corr, num_task, ig_unique, fiu)
    else:
        scr = np.zeros((num_g_corr, len(fiu), num_task), dtype=np.complex128, order="F")
        for it in range(num_task):
            scr[int(ig_unique[igstart - 1 + it]) - 1, :, it] = 1.0
    return pool_sum_eps(scr, igstart, ig_unique)


def root_sum(a: np.ndarray, root: int = 0, device=None):
    """mp_root_sum(comm, root, array) (data/parallel/src/parallel.f90, used at sigma.f90:362,383): element-wise sum
    over ranks, result on ``root`` (None elsewhere).  One reduce of the Sigma(k, omega) block per k-point: the only
    collective of the Sigma stage (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
    import torch
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return a
    dev = _device(dist, device)
    t = torch.from_numpy(np.ascontiguousarray(a).view(np.float64).copy()).to(dev)
    dist.reduce(t, dst=root, op=dist.ReduceOp.SUM)
    if dist.get_rank() != root:
        return None
    return np.ascontiguousarray(t.cpu().numpy()).view(np.complex128).reshape(a.shape)


def sigma_wrapper_k(sigma_correlation_fn, configs, num_g_corr, num_sigma, root: int = 0):
    """The k-point loop body of sigma_wrapper (phys/corr/src/sigma.f90:319-362) over ranks.

    ``configs``: the (k, q) configurations of this k-point (sigma.f90 ``config(:)``: index_kq, index_q, sym_op, weight).
    The reference gives every pool the whole list and distributes G' over images inside each product; here the
    configurations themselves are dealt to the ranks with ``parallel_task``'s block rule (one rank = one GPU, every
    product runs entirely on one device), each rank accumulates its share into a local sigma and ``root_sum`` adds
    the shares on the root -- the allreduce of Sigma(k, omega) contributions.

    sigma_correlation_fn(config, sigma) accumulates one configuration in place (normally a closure around
    ``Context.sigma_correlation``).  Returns (sigma on root | None, (first, last, num_task)).
    """
    dist = _dist()
    rank = dist.get_rank() if dist else 0
    world = dist.get_world_size() if dist else 1
    first, last, num_task = parallel_task(world, rank, len(configs))
    sigma = np.zeros((num_g_corr, num_g_corr, num_sigma), dtype=np.complex128, order="F")
    for icon in range(first - 1, first - 1 + num_task[rank]):
        sigma_correlation_fn(configs[icon], sigma)
    total = root_sum(np.ascontiguousarray(np.transpose(sigma, (2, 1, 0))), root=root)
    if total is None:
        return None, (first, last, num_task)
    return np.asfortranarray(np.transpose(total, (2, 1, 0))), (first, last, num_task)
