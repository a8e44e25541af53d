"""N > 1 path on CPU: the do_stern split/gather logic (sternheimergw_b200/dist.py) with world_size 2 and 3 over gloo.

The data path has no collective (perturbation blocks are independent, do_stern.f90:199-209); what is tested here is
the host-side logic around it: parallel_task's block rule (parallel.f90:80-138, remainder to the LAST ranks), the
gather of scrcoul_loc(ngc, nfs, ntask_loc) in task order (do_stern.f90:211 / parallel.f90:1130) including ranks with
zero tasks and unequal blocks, and the root-only unfold (unfold_w.f90:84: conjugate transpose).
coulomb_fn is a deterministic stand-in (no GPU here); the GPU tests cover Context.coulomb itself.
"""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _fake_coulomb(config, igstart, ngc, ntask, ig_unique, fiu):
    out = np.zeros((ngc, len(fiu), ntask), dtype=np.complex128, order="F")
    for t in range(ntask):
        ig = int(ig_unique[igstart - 1 + t])
        for iw in range(len(fiu)):
            out[:, iw, t] = (np.arange(ngc) + 1) * (1.0 + 0.5j * iw) + 1000.0 * ig
    return out


def _unfold_identity(ngc, ig_unique, scr):
    out = np.zeros((ngc, ngc, scr.shape[1]), dtype=np.complex128, order="F")
    for i, ig in enumerate(ig_unique):
        out[ig - 1, :, :] = np.conj(scr[:, :, i])           # unfold_w.f90:84
    return out


def _worker(rank, world, port, ngmunique, ngc, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sternheimergw_b200.dist import do_stern_q
        ig_unique = np.arange(1, ngmunique + 1, dtype=np.int32)
        fiu = np.array([0.0, 0.5j, 1.0j])
        scr_g, (first, last, num_task) = do_stern_q(_fake_coulomb, None, ngc, ig_unique, fiu, unfold_fn=_unfold_identity)
        q.put((rank, first, last, num_task, None if scr_g is None else scr_g.copy()))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,ngmunique,expect", [(2, 7, [3, 4]), (3, 7, [2, 2, 3]), (3, 2, [0, 1, 1]), (2, 8, [4, 4])])
def test_do_stern_split_and_gather(world, ngmunique, expect):
    ngc = max(ngmunique, 5)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ngmunique, ngc, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r = q.get(timeout=120)
        res[r[0]] = r
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # serial answer
    ig_unique = np.arange(1, ngmunique + 1, dtype=np.int32)
    fiu = np.array([0.0, 0.5j, 1.0j])
    serial = _unfold_identity(ngc, ig_unique, _fake_coulomb(None, 1, ngc, ngmunique, ig_unique, fiu))
    off = 1
    for r in range(world):
        _, first, last, num_task, scr = res[r]
        assert list(num_task) == expect
        if expect[r]:
            assert (first, last) == (off, off + expect[r] - 1)
        off += expect[r]
        if r == 0:
            assert scr is not None and np.array_equal(scr, serial)          # bit-exact: a gather moves bytes only
        else:
            assert scr is None


def _fake_sigma_correlation(config, sigma):
    """Deterministic stand-in with integer-valued contributions, so that the rank-wise sum is bit-exact."""
    ngc, _, nsig = sigma.shape
    g = np.arange(ngc)
    for i in range(nsig):
        sigma[:, :, i] += (config["index_kq"] * 7 + i) * (np.add.outer(g, 2 * g) + 1j * np.subtract.outer(g, g)) * config["weight"]


def _sigma_worker(rank, world, port, ncon, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sternheimergw_b200.dist import sigma_wrapper_k
        configs = [{"index_kq": c + 1, "weight": 2 ** (c % 3)} for c in range(ncon)]
        sig, (first, last, num_task) = sigma_wrapper_k(_fake_sigma_correlation, configs, 5, 3)
        q.put((rank, first, last, num_task, None if sig is None else sig.copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,ncon,expect", [(2, 7, [3, 4]), (3, 2, [0, 1, 1])])
def test_sigma_configurations_dealt_and_summed(world, ncon, expect):
    """Sigma stage over ranks: (k, q) configurations dealt with parallel_task's rule, shares added by root_sum
    (mp_root_sum of sigma.f90:362) -- the only collective of that stage."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sigma_worker, args=(r, world, port, ncon, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r = q.get(timeout=120)
        res[r[0]] = r
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    serial = np.zeros((5, 5, 3), dtype=np.complex128, order="F")
    for c in range(ncon):
        _fake_sigma_correlation({"index_kq": c + 1, "weight": 2 ** (c % 3)}, serial)
    for r in range(world):
        _, first, last, num_task, sig = res[r]
        assert list(num_task) == expect
        if r == 0:
            assert sig is not None and np.array_equal(sig, serial)
        else:
            assert sig is None


def _invert_fake(scr, lgamma=False):
    out = np.zeros_like(scr)
    for iw in range(scr.shape[2]):
        out[:, :, iw] = np.linalg.inv(scr[:, :, iw] + 5000.0 * np.eye(scr.shape[0])) - np.eye(scr.shape[0])
    return np.asfortranarray(out)


def _worker_shard(rank, world, port, ngmunique, ngc, nfs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sternheimergw_b200.dist import do_stern_q
        ig_unique = np.arange(1, ngmunique + 1, dtype=np.int32)
        fiu = 0.5j * np.arange(nfs)
        tm = {}
        w, _ = do_stern_q(_fake_coulomb, None, ngc, ig_unique, fiu, unfold_fn=_unfold_identity, invert_fn=_invert_fake,
                          shard_invert=True, timings=tm)
        q.put((rank, None if w is None else w.copy(), sorted(tm)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nfs", [(2, 3), (3, 2), (3, 5)])
def test_do_stern_frequency_sharded_unfold_and_invert(world, nfs):
    """shard_invert: every rank unfolds + inverts a contiguous share of the frequencies (parallel_task's rule, ranks with
    zero frequencies included) and the root gathers; the result equals the root-only path of do_stern.f90:220-232."""
    ngmunique = ngc = 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_shard, args=(r, world, port, ngmunique, ngc, nfs, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r = q.get(timeout=120)
        res[r[0]] = r
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ig_unique = np.arange(1, ngmunique + 1, dtype=np.int32)
    fiu = 0.5j * np.arange(nfs)
    serial = _invert_fake(_unfold_identity(ngc, ig_unique, _fake_coulomb(None, 1, ngc, ngmunique, ig_unique, fiu)))
    assert res[0][1] is not None and np.allclose(res[0][1], serial, rtol=0, atol=1e-14)
    assert all(res[r][1] is None for r in range(1, world))
    assert res[0][2] == ["gather_s", "gather_w_s", "unfold_invert_s"]


def _fake_coulomb_k(kset, igstart, ngc, ntask, ig_unique, nfs):
    """eps columns of a fake system whose density response is a sum over k-points: scr = delta - sum_k drho_k."""
    out = np.zeros((ngc, nfs, ntask), dtype=np.complex128, order="F")
    for t in range(ntask):
        ig = int(ig_unique[igstart - 1 + t])
        out[ig - 1, :, t] = 1.0
        for k in kset:
            for iw in range(nfs):
                out[:, iw, t] -= (0.1 * (k + 1) + 0.01j * iw) * np.cos(np.arange(ngc) * (ig + k))
    return out


def _pool_worker(rank, world, port, nk, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sternheimergw_b200.dist import pool_sum_eps
        from sternheimergw_b200.host import parallel_task
        first, last, num = parallel_task(world, rank, nk)
        kset = list(range(first - 1, first - 1 + num[rank]))
        ig_unique = np.array([2, 5, 1, 4], dtype=np.int32)
        loc = _fake_coulomb_k(kset, 2, 6, 3, ig_unique, 2)
        q.put((rank, pool_sum_eps(loc, 2, ig_unique).copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nk", [(2, 5), (3, 4), (3, 2)])
def test_pool_sum_of_eps_columns_over_k_shards(world, nk):
    """k-points shared among ranks (pools, solve_linter.f90:521): allreduce of the affine eps columns == all k on one rank,
    including a rank that holds no k-point at all."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pool_worker, args=(r, world, port, nk, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    serial = _fake_coulomb_k(list(range(nk)), 2, 6, 3, np.array([2, 5, 1, 4], dtype=np.int32), 2)
    for r in range(world):
        assert np.abs(res[r] - serial).max() < 1e-14
