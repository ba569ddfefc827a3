"""CPU, world_size 2, gloo: the host logic of the row-sharded search (shard bounds, global id offsets,
all-gather plumbing, merge order).  The CUDA engine cannot run here, so the two seams of
ShardedIndexFlat are filled with test doubles built on the oracle; the product defaults are untouched."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleLocalIndex:
    """Stands in for textreact_b200.IndexFlat on a CPU box."""

    def __init__(self, d, metric):
        self.d, self.metric, self.xb, self.off, self.groups = d, metric, None, 0, None

    @property
    def ntotal(self):
        return 0 if self.xb is None else self.xb.shape[0]

    def add(self, x):
        self.xb = np.ascontiguousarray(x, dtype=np.float32)

    def set_id_offset(self, off):
        self.off = off

    def set_groups(self, g):
        self.groups = np.ascontiguousarray(g, dtype=np.int32)

    def search(self, xq, k, *, exclude=None):
        from oracle import cpu_flat as oracle
        D, I = oracle.search_seq(self.xb, xq, k, self.metric, self.groups, exclude)
        return D, np.where(I >= 0, I + self.off, -1)


class TwoPhaseOracleLocalIndex(OracleLocalIndex):
    """... with the two halves of the two-phase search (IndexFlat.search_begin / search_finish): the "prefilter" scores
    are the exact ones (eps = 0), in the larger-is-better domain the engine exchanges."""
    max_batch = 4          # several chunks per call

    def get_option(self, key):
        assert key == "max_batch"
        return self.max_batch

    def search_begin(self, xq, k, nb, *, exclude=None):
        D, I = self.search(np.asarray(xq), k, exclude=exclude)
        self._held = (D, I)
        s = np.where(I >= 0, D if self.metric == 0 else -D, -np.inf).astype(np.float32)
        top = np.full((len(s), nb), -np.inf, np.float32)
        top[:, :min(nb, k)] = s[:, :nb]
        return torch.from_numpy(np.concatenate([top, np.zeros((len(s), 1), np.float32)], axis=1))

    def search_finish(self, floor):
        D, I = self._held
        s = np.where(I >= 0, D if self.metric == 0 else -D, -np.inf)
        keep = s >= floor.numpy()[:, None]                       # rows under the floor cannot reach the global top-k
        self.dropped = getattr(self, "dropped", 0) + int((~keep & (I >= 0)).sum())
        fill = np.float32(-3.4028235e38 if self.metric == 0 else 3.4028235e38)
        return np.where(keep, D, fill), np.where(keep, I, -1)


def numpy_merge(Dg, Ig, metric):
    """Reference merge: (score, id) order over the concatenated shard lists."""
    Dg, Ig = Dg.numpy(), Ig.numpy()
    G, nq, k = Dg.shape
    D = np.empty((nq, k), np.float32)
    I = np.empty((nq, k), np.int64)
    for q in range(nq):
        d, i = Dg[:, q].reshape(-1), Ig[:, q].reshape(-1)
        key = np.where(i >= 0, -d if metric == 0 else d, np.inf)
        order = np.lexsort((np.where(i >= 0, i, np.iinfo(np.int64).max), key))[:k]
        D[q], I[q] = d[order], i[order]
    return torch.from_numpy(D), torch.from_numpy(I)


def _worker(rank, world, port, metric, with_mask, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import cpu_flat as oracle
        from textreact_b200.sharded import ShardedIndexFlat, shard_bounds
        rng = np.random.default_rng(5)
        n, d, nq, k = 1001, 24, 9, 7                                  # odd n: uneven shards
        xb = rng.standard_normal((n, d)).astype(np.float32)
        xq = rng.standard_normal((nq, d)).astype(np.float32)
        groups = (np.arange(n) // 4).astype(np.int32)
        excl = groups[rng.integers(0, n, nq)].astype(np.int32) if with_mask else None
        idx = ShardedIndexFlat(d, metric, local_factory=OracleLocalIndex, merge_fn=numpy_merge)
        idx.add_global(xb)
        lo, hi = shard_bounds(n, world, rank)
        assert idx.local.ntotal == hi - lo and idx.ntotal == n
        if with_mask:
            idx.set_groups_global(groups)
        D, I = idx.search(xq, k, exclude=excl)
        Do, Io = oracle.search_seq(xb, xq, k, metric, groups if with_mask else None, excl)
        np.testing.assert_array_equal(I, Io)
        np.testing.assert_allclose(D, Do, rtol=1e-6)
        # two-phase local search (bounds exchange over gloo, chunks of max_batch queries): same answer, rows dropped
        tp = ShardedIndexFlat(d, metric, local_factory=TwoPhaseOracleLocalIndex, merge_fn=numpy_merge)
        assert tp._two_phase
        tp.add_global(xb)
        if with_mask:
            tp.set_groups_global(groups)
        Dt, It = tp.search(xq, k, exclude=excl)
        np.testing.assert_array_equal(It, Io)
        np.testing.assert_allclose(Dt, Do, rtol=1e-6)
        assert tp.local.dropped > 0
        # sharded merge: this rank keeps the merged rows of its query slice only
        Ds, Is = idx.search(xq, k, exclude=excl, result="slice")
        qlo, qhi = idx.query_slice(nq)
        assert (qlo, qhi) == shard_bounds(nq, world, rank) and Is.shape == (qhi - qlo, k)
        np.testing.assert_array_equal(Is, Io[qlo:qhi])
        np.testing.assert_allclose(Ds, Do[qlo:qhi], rtol=1e-6)
        out[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("metric,with_mask", [(0, False), (1, False), (0, True)])
def test_two_rank_sharded_search_equals_unsharded(metric, with_mask):
    world = 2
    port = 29600 + os.getpid() % 300 + metric * 7 + int(with_mask)
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, metric, with_mask, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert len(out) == world


def test_weighted_shard_bounds_cover_rows_exactly():
    """Shards proportional to measured GPU speed (ShardedIndexFlat.calibrate): contiguous, complete, tile-aligned."""
    from textreact_b200.sharded import shard_bounds
    for n in (0, 1, 7, 1000, 60001, 16_000_000):
        for w in ([1, 1], [0.26, 0.24, 0.25, 0.25], [1.07, 1.0, 0.93, 1.0, 1.0, 1.0, 1.02, 0.98]):
            W = len(w)
            b = [shard_bounds(n, W, r, w) for r in range(W)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(W - 1)) and all(lo <= hi for lo, hi in b)
            if n >= W * 256:
                assert all(lo % 256 == 0 for lo, _ in b)
    lo, hi = shard_bounds(16_000_000, 4, 0, [0.26, 0.24, 0.25, 0.25])
    assert (lo, hi) == (0, 4_160_000)
    assert shard_bounds(1000, 4, 2, None) == (500, 750)


def test_shard_bounds_cover_rows_exactly():
    from textreact_b200.sharded import shard_bounds
    for n in (0, 1, 7, 1000, 16_000_000):
        for w in (1, 2, 4, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _replica_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import cpu_flat as oracle
        from textreact_b200.sharded import ReplicatedIndexFlat
        rng = np.random.default_rng(9)
        xb = rng.standard_normal((700, 16)).astype(np.float32)
        xq = rng.standard_normal((11, 16)).astype(np.float32)           # odd: uneven query slices
        idx = ReplicatedIndexFlat(16, 0, local_factory=OracleLocalIndex)
        idx.add(xb)
        D, I = idx.search(xq, 5)
        Do, Io = oracle.search_seq(xb, xq, 5, 0)
        np.testing.assert_array_equal(I, Io)
        np.testing.assert_allclose(D, Do, rtol=1e-6)
        out[rank] = True
    finally:
        dist.destroy_process_group()


def test_two_rank_query_sharded_replicas():
    world = 2
    port = 29400 + os.getpid() % 300
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    procs = [ctx.Process(target=_replica_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert len(out) == world


def test_floor_from_payloads_is_a_valid_lower_bound():
    """The bounds exchange of the two-phase search, host restatement (what trx_exchange_floor computes in its kernel):
    floor = k-th largest of the G * nb exchanged prefilter scores - 2 * max eps, -inf when fewer than k are finite.
    Property checked on random shards: no row of the true global top-k (by exact score, |exact - prefilter| <= eps)
    has a prefilter score below the floor."""
    from textreact_b200.sharded import bounds_width, floor_from_payloads
    rng = np.random.default_rng(3)
    G, nq, k, rows = 4, 50, 20, 400
    nb = bounds_width(k, G)
    assert nb * G >= k and nb <= k and bounds_width(100, 8) == 32 and bounds_width(100, 1) == 100
    exact = rng.standard_normal((G, nq, rows)).astype(np.float32) * 5
    eps = rng.uniform(0.05, 0.2, (G, nq)).astype(np.float32)
    pre = exact + rng.uniform(-1, 1, exact.shape).astype(np.float32) * eps[:, :, None]      # |pre - exact| <= eps
    top = -np.sort(-pre, axis=2)[:, :, :nb]
    payload = torch.from_numpy(np.concatenate([top, eps[:, :, None]], axis=2))
    floor = floor_from_payloads(payload, nb, k).numpy()
    allx = exact.transpose(1, 0, 2).reshape(nq, -1)
    allp = pre.transpose(1, 0, 2).reshape(nq, -1)
    for q in range(nq):
        topk = np.argsort(-allx[q])[:k]
        assert (allp[q][topk] >= floor[q]).all()
    # the bound is not vacuous: it cuts most of every shard's rows
    assert (allp < floor[:, None]).mean() > 0.8
    # too few finite scores -> no bound
    payload[:, 0, :nb] = float("-inf")
    assert floor_from_payloads(payload, nb, k)[0] == float("-inf")
