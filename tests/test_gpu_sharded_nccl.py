"""GPU, world_size 2 / 4 / 8 (whatever the box has): the row-sharded search on the real engine -- local exact top-k per
shard with global ids, then the exchange, either fused (all-gather inside the merge kernel over NVLink peer
memory, CUDA IPC + flag protocol, K5p) or NCCL all-gather + device k-way merge (K5) -- must equal the unsharded
answer.  Skipped on a one-GPU box (the gloo test covers the host logic)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, metric, with_mask, exchange, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), TRX_EXCHANGE_ENTRIES="1000")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from oracle import cpu_flat as oracle
        from tests import util
        from textreact_b200.sharded import ShardedIndexFlat, shard_bounds
        n, d, nq, k = 60001, 768, 300, 20                               # odd n: uneven shards
        xb, xq = util.gaussian(n, d, 5), util.gaussian(nq, d, 6)
        groups = (np.arange(n) // 4).astype(np.int32)
        excl = groups[np.random.default_rng(7).integers(0, n, nq)].astype(np.int32) if with_mask else None
        idx = ShardedIndexFlat(d, metric, device=rank, exchange=exchange)
        # shards proportional to measured speed for one of the cases (weights agree on every rank: all-gathered)
        weights = idx.calibrate(batch=256, rows=20000, seconds=0.05, k=k) if (with_mask and exchange == "peer") else None
        if weights is not None:
            assert len(weights) == world and abs(sum(weights) - 1.0) < 1e-9 and min(weights) > 0.5 / world
            weights = [w * (1.3 if r == 0 else 1.0) for r, w in enumerate(weights)]      # and visibly uneven
        idx.add_global(xb, weights=weights)
        lo, hi = shard_bounds(n, world, rank, weights)
        assert idx.local.ntotal == hi - lo and idx.ntotal == n
        if with_mask:
            idx.set_groups_global(groups)
        # device tensors in -> device tensors out, every rank holds the merged result
        D, I = idx.search(torch.from_numpy(xq).cuda(), k,
                          exclude=None if excl is None else torch.from_numpy(excl).cuda())
        assert D.is_cuda and I.is_cuda
        oracle.check_parity(D.cpu().numpy(), I.cpu().numpy(), xb, xq, k, metric,
                            groups if with_mask else None, excl)
        # host arrays in -> host arrays out
        D2, I2 = idx.search(xq[:33], k, exclude=None if excl is None else excl[:33])
        np.testing.assert_array_equal(I2, I.cpu().numpy()[:33])
        # sharded merge: each rank receives only its slice of the queries (1/world of the exchange traffic)
        Ds, Is = idx.search(torch.from_numpy(xq).cuda(), k, exclude=None if excl is None else torch.from_numpy(excl).cuda(),
                            result="slice")
        qlo, qhi = idx.query_slice(nq)
        assert tuple(Is.shape) == (qhi - qlo, k)
        np.testing.assert_array_equal(Is.cpu().numpy(), I.cpu().numpy()[qlo:qhi])
        np.testing.assert_array_equal(Ds.cpu().numpy(), D.cpu().numpy()[qlo:qhi])
        # pipelined form: several exchanges pending while later local searches run
        xq_t = torch.from_numpy(xq).cuda()
        ex_t = None if excl is None else torch.from_numpy(excl).cuda()
        hs = [idx.search_async(xq_t[i * 100:(i + 1) * 100].contiguous(), k,
                               exclude=None if ex_t is None else ex_t[i * 100:(i + 1) * 100].contiguous())
              for i in range(3)]
        for i, h in enumerate(hs):
            Da, Ia = h.result()
            torch.cuda.current_stream().synchronize()
            np.testing.assert_array_equal(Ia.cpu().numpy(), I.cpu().numpy()[i * 100:(i + 1) * 100])
            np.testing.assert_array_equal(Da.cpu().numpy(), D.cpu().numpy()[i * 100:(i + 1) * 100])
        assert idx._exchange_mode == exchange, "peer mapping failed: fell back to NCCL"
        # a larger exchange than the export buffers were sized for: they are rebuilt collectively
        if exchange == "peer":
            Db, Ib = idx.search(xq_t[:500].contiguous(), 256)
            oracle.check_parity(Db.cpu().numpy()[:20], Ib.cpu().numpy()[:20], xb, xq[:20], 256, metric)
        out[rank] = idx.local.stats()["last_path"]
        idx.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("exchange", ["peer", "nccl"])
@pytest.mark.parametrize("metric,with_mask", [(0, False), (1, False), (0, True)])
def test_sharded_search_equals_unsharded(metric, with_mask, exchange, world):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29900 + os.getpid() % 300 + metric * 7 + int(with_mask) + (13 if exchange == "peer" else 0) + 31 * world
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, metric, with_mask, exchange, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert len(out) == world


def test_two_devices_in_one_process():
    """Indexes on different GPUs of the same process (the C ABI takes a device ordinal; per-device kernel
    attributes, device guards around every call)."""
    import torch
    import textreact_b200 as trx
    from oracle import cpu_flat as oracle
    from tests import util
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    xb, xq = util.gaussian(30000, 256, 301), util.gaussian(200, 256, 302)
    res = []
    for dev in (0, 1):
        idx = trx.IndexFlatIP(256, device=dev)
        idx.add(xb)
        res.append(idx.search(xq, 10))                                   # numpy in / out, current device stays 0
        Dt, It = idx.search(torch.from_numpy(xq).to(f"cuda:{dev}"), 10)    # tensors on the index's device
        assert It.device.index == dev
        np.testing.assert_array_equal(It.cpu().numpy(), res[-1][1])
        idx.close()
    assert torch.cuda.current_device() == 0
    np.testing.assert_array_equal(res[0][1], res[1][1])
    oracle.check_parity(res[0][0], res[0][1], xb, xq, 10, 0)


def _fuzz_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from oracle import cpu_flat as oracle
        from tests import util
        from textreact_b200.sharded import ShardedIndexFlat
        ok = 0
        for seed in range(12):
            rng = np.random.default_rng(4000 + seed)                 # same draws on every rank
            d = int(rng.choice([32, 100, 256, 768]))
            n = int(rng.choice([1001, 9000, 30001, 50000]))
            nq = int(rng.choice([1, 5, 64, 200]))
            k = int(rng.choice([1, 10, 100]))
            metric = int(rng.integers(0, 2))
            xb, xq = util.gaussian(n, d, 5000 + seed), util.gaussian(nq, d, 6000 + seed)
            groups = (rng.permutation(n) // 3).astype(np.int32)
            excl = groups[rng.integers(0, n, nq)].astype(np.int32) if seed % 2 else None
            idx = ShardedIndexFlat(d, metric, device=rank, exchange="peer" if seed % 3 else "nccl")
            idx.add_global(xb)
            idx.set_groups_global(groups)
            idx.local.set_option("path", int(rng.integers(0, 4)) if nq <= 8 else int(rng.choice([0, 1, 3])))
            D, I = idx.search(xq, k, exclude=excl)                    # host arrays in / out
            oracle.check_parity(D, I, xb, xq, k, metric, groups if excl is not None else None, excl)
            idx.close()
            ok += 1
        out[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 8])
def test_sharded_fuzz(world):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = 29700 + os.getpid() % 200 + world
    procs = [ctx.Process(target=_fuzz_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    assert dict(out) == {r: 12 for r in range(world)}


def _replica_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from oracle import cpu_flat as oracle
        from tests import util
        from textreact_b200.sharded import ReplicatedIndexFlat
        xb, xq = util.gaussian(30000, 128, 401), util.gaussian(301, 128, 402)       # odd query count
        idx = ReplicatedIndexFlat(128, 0, device=rank)
        idx.add(xb)
        D, I = idx.search(torch.from_numpy(xq).cuda(), 10)
        oracle.check_parity(D.cpu().numpy(), I.cpu().numpy(), xb, xq, 10, 0)
        D2, I2 = idx.search(xq, 10)
        np.testing.assert_array_equal(I2, I.cpu().numpy())
        idx.close()
        out[rank] = True
    finally:
        dist.destroy_process_group()


def test_two_gpu_query_sharded_replicas():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = 29300 + os.getpid() % 200
    procs = [ctx.Process(target=_replica_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert len(out) == 2
