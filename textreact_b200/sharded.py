"""Row-sharded multi-GPU search: one process per GPU, one exchange step.

No reference counterpart (the reference's retrieval is single-process CPU FAISS,
retrieve/retrieve_faiss.py:62-74); this is SURVEY.md section 8e / north_star (4):
shard g of G holds corpus rows [g*N/G, (g+1)*N/G) and answers every query with a local exact
top-k carrying GLOBAL ids; one all-gather of the [nq, k] score and id lists (NCCL over
NVLink/NVSwitch) feeds the device k-way merge (K5, ``trx_merge_topk``).  Every rank ends up
with the full result.

``local_factory`` / ``merge_fn`` are seams for the CPU (gloo) tests of the host logic; the
defaults are the CUDA engine and there is no CPU fallback behind them.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

import ctypes
import os

from . import _lib
from .index import IndexFlat, _stream_handle, merge_topk


def shard_bounds(n, world, rank, weights=None, align=256):
    """Contiguous split: rows [lo, hi) of shard ``rank``.  Near-even by default; with ``weights`` (one positive number
    per rank, e.g. the measured relative speed of each GPU) the shard sizes are proportional to them, interior
    boundaries rounded to a multiple of ``align`` rows (the scoring kernel's corpus tile)."""
    if weights is None:
        return rank * n // world, (rank + 1) * n // world
    w = [float(v) for v in weights]
    assert len(w) == world and all(v > 0 for v in w), "one positive weight per rank"
    total = sum(w)

    def edge(r):
        if r <= 0:
            return 0
        if r >= world:
            return n
        e = int(round(n * sum(w[:r]) / total))
        if align > 1 and n >= world * align:
            e = int(round(e / align)) * align
        return min(max(e, 0), n)
    lo, hi = edge(rank), edge(rank + 1)
    return lo, max(lo, hi)


def floor_from_payloads(g, nb, k):
    """[G, nq, nb + 1] gathered bounds payloads (``IndexFlat.search_begin``) -> floor[nq]: the k-th largest of the
    G * nb exchanged prefilter scores is a lower bound on the global k-th prefilter score, the rows behind it score
    exactly at least that minus eps, so a row whose prefilter score is below  kth - 2 max(eps)  cannot be in the
    global top-k.  -inf where the shards hold fewer than k candidates between them.  (What ``trx_exchange_floor``
    computes inside its peer-memory kernel.)"""
    G, nq = g.shape[0], g.shape[1]
    scores = g[:, :, :nb].permute(1, 0, 2).reshape(nq, -1)
    eps = g[:, :, nb].max(dim=0).values
    if scores.shape[1] < k:
        return torch.full((nq,), float("-inf"), dtype=torch.float32, device=g.device)
    kth = torch.topk(scores, k, dim=1).values[:, k - 1]
    floor = kth - 2.0 * eps * 1.00001
    floor = floor - floor.abs() * 1e-6 - 1e-30
    return torch.where(torch.isfinite(kth), floor, torch.full_like(floor, float("-inf"))).contiguous()


def bounds_width(k, world):
    """Prefilter scores per query every shard contributes to the bounds exchange: its expected share of the global
    top-k (k / world) plus four standard deviations -- any width gives a valid (lower) bound; at least k in total."""
    share = k / world
    return int(min(k, max(8, -(-int(share + 4.0 * share ** 0.5 + 4.0) // 8) * 8)))


class PendingSearch:
    """Handle of a ``search_async``: the merged (D, I) become valid for the caller's stream in ``result()``."""

    def __init__(self, D, I, event, as_numpy):
        self._D, self._I, self._event, self._as_numpy = D, I, event, as_numpy

    def result(self):
        D, I = self._D, self._I
        if self._event is not None:
            cur = torch.cuda.current_stream(D.device)
            cur.wait_event(self._event)
            D.record_stream(cur); I.record_stream(cur)
            if self._as_numpy:
                self._event.synchronize()
        if self._as_numpy and isinstance(D, torch.Tensor):
            return D.cpu().numpy(), I.cpu().numpy()
        return D, I


class ReplicatedIndexFlat:
    """The other way to use G GPUs when the corpus fits on one (16M x 768 is 74 GB of a B200's 180 GB): every rank
    holds ALL rows and answers its 1/G slice of the queries; the slices are concatenated with one all-gather.
    No merge, no per-shard rescoring redundancy -- but G copies of the corpus.  north_star's contract is the
    row-sharded ``ShardedIndexFlat``; this is the throughput option beside it."""

    def __init__(self, d, metric, *, group=None, device=None, local_factory=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.d, self.metric_type = int(d), int(metric)
        self._on_cuda = local_factory is None          # the product path; test doubles stay on the host
        self.local = (local_factory or (lambda d_, m_: IndexFlat(d_, m_, device=device)))(d, metric)

    @property
    def ntotal(self):
        return self.local.ntotal

    def add(self, x):
        self.local.add(x)

    def set_groups(self, groups):
        self.local.set_groups(groups)

    def search(self, xq, k, *, exclude=None, **kw):
        """xq replicated on every rank; every rank returns the full (D, I)."""
        nq = xq.shape[0]
        lo, hi = shard_bounds(nq, self.world, self.rank)
        per = -(-nq // self.world)                                  # equal slices for the all-gather (last one padded)
        as_numpy = not (isinstance(xq, torch.Tensor) and xq.is_cuda)
        ex = None if exclude is None else exclude[lo:hi]
        D, I = self.local.search(xq[lo:hi].contiguous() if not as_numpy else xq[lo:hi], k, exclude=ex, **kw) \
            if hi > lo else (None, None)
        dev = torch.device("cuda", self.local.device) if self._on_cuda else torch.device("cpu")
        Dp = torch.full((per, k), float("nan"), dtype=torch.float32, device=dev)
        Ip = torch.full((per, k), -1, dtype=torch.int64, device=dev)
        if hi > lo:
            Dp[:hi - lo] = torch.as_tensor(D, device=dev)
            Ip[:hi - lo] = torch.as_tensor(I, device=dev)
        Dg = torch.empty((self.world, per, k), dtype=torch.float32, device=dev)
        Ig = torch.empty((self.world, per, k), dtype=torch.int64, device=dev)
        if dev.type == "cuda":
            dist.all_gather_into_tensor(Dg, Dp, group=self.group)
            dist.all_gather_into_tensor(Ig, Ip, group=self.group)
        else:
            dist.all_gather(list(Dg.unbind(0)), Dp, group=self.group)
            dist.all_gather(list(Ig.unbind(0)), Ip, group=self.group)
        parts_D, parts_I = [], []
        for g in range(self.world):
            glo, ghi = shard_bounds(nq, self.world, g)
            parts_D.append(Dg[g, :ghi - glo]); parts_I.append(Ig[g, :ghi - glo])
        Dm, Im = torch.cat(parts_D), torch.cat(parts_I)
        if as_numpy:
            return Dm.cpu().numpy(), Im.cpu().numpy()
        return Dm, Im

    def close(self):
        if hasattr(self.local, "close"):
            self.local.close()


class ShardedIndexFlat:
    def __init__(self, d, metric, *, group=None, local_factory=None, merge_fn=None, device=None, exchange=None,
                 two_phase=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.d, self.metric_type = int(d), int(metric)
        self._on_cuda = local_factory is None          # the product path; test doubles stay on the host
        self.local = (local_factory or (lambda d_, m_: IndexFlat(d_, m_, device=device)))(d, metric)
        self._merge = merge_fn or merge_topk
        self._ntotal_global = 0
        self._lo = 0
        self._xstream = None          # side stream of the exchange step (CUDA only)
        self._xev = None              # ... and the event of the last exchange queued on it
        # "peer": all-gather fused into the merge kernel over NVLink peer memory (CUDA IPC, K5p); "nccl": NCCL
        # all-gather + K5.  Default: peer on the CUDA engine, NCCL for the CPU test doubles / when mapping fails.
        self._exchange_mode = exchange or os.environ.get("TRX_EXCHANGE") or ("peer" if self._on_cuda and merge_fn is None else "nccl")
        self._peer = None             # trx_exchange handle
        self._peer_entries = 0
        # Two-phase local search (CUDA engine only): the shards first exchange their best PREFILTER scores, which
        # bounds the global k-th score from below, and each shard then rescores only the candidates that can still
        # reach the global top-k -- ~1/G of what certifying its own top-k would take (K4 does not shrink with G
        # otherwise).  TRX_TWO_PHASE=0 / two_phase=False: every shard computes its full local top-k.
        if two_phase is None:
            two_phase = os.environ.get("TRX_TWO_PHASE", "1") != "0"
        # (test doubles take part when they implement search_begin / search_finish: the host logic runs on gloo)
        self._two_phase = bool(two_phase) and ((self._on_cuda and merge_fn is None) or hasattr(self.local, "search_begin"))

    @property
    def ntotal(self):
        return self._ntotal_global

    def add_global(self, x, weights=None):
        """Every rank is handed the same [N, d] array (or a lazily sliced view); keeps its slice -- near-even, or
        proportional to ``weights`` (see ``calibrate``)."""
        n = x.shape[0]
        lo, hi = shard_bounds(n, self.world, self.rank, weights)
        self.add_shard(x[lo:hi], lo, n)

    def calibrate(self, batch=8192, rows=1_000_000, seconds=1.5, k=100):
        """Relative scoring speed of every rank's GPU, measured: each rank searches a throw-away random index for
        ``seconds`` and the device time of the scoring passes per corpus row is all-gathered.  Returns one weight per
        rank (sum 1), for ``shard_bounds(..., weights=)`` / ``add_global(..., weights=)``.  B200s under the 1 kW power
        cap differ by several per cent in sustained tensor throughput; with equal shards every step waits for the
        slowest one, with shards proportional to speed they finish together.  Collective (one all-gather)."""
        import time
        assert self._on_cuda, "calibrate() measures the CUDA engine"
        dev = torch.device("cuda", self.local.device)
        probe = IndexFlat(self.d, self.metric_type, device=self.local.device)
        try:
            g = torch.Generator(device=dev)
            g.manual_seed(97 + self.rank)
            for c0 in range(0, rows, 250_000):
                probe.add(torch.randn((min(250_000, rows - c0), self.d), generator=g, device=dev))
            q = torch.randn((batch, self.d), generator=g, device=dev)
            for _ in range(3):
                probe.search(q, k)
            s0, t0, n = probe.stats(), time.perf_counter(), 0
            while n < 3 or time.perf_counter() - t0 < seconds:
                probe.search(q, k)
                n += 1
            s1 = probe.stats()
            nb = max(1, s1["timed_batches"] - s0["timed_batches"])
            ms = (s1["sum_prefilter_ms"] - s0["sum_prefilter_ms"] + s1["sum_sample_ms"] - s0["sum_sample_ms"]) / nb
        finally:
            probe.close()
        mine = torch.tensor([ms / rows], dtype=torch.float64, device=dev)
        allr = torch.empty((self.world,), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allr, mine, group=self.group)
        speed = 1.0 / allr.clamp(min=1e-12)
        return [float(v) for v in (speed / speed.sum()).tolist()]

    def add_shard(self, x_local, lo, n_global):
        """This rank's rows are global rows [lo, lo + len(x_local))."""
        assert self.local.ntotal == 0, "one shard per index"
        self.local.add(x_local)
        self.local.set_id_offset(lo)
        self._lo, self._ntotal_global = int(lo), int(n_global)

    def set_groups_global(self, groups):
        lo, hi = self._lo, self._lo + self.local.ntotal
        self.local.set_groups(groups[lo:hi])

    def set_row_attr_global(self, attr):
        lo, hi = self._lo, self._lo + self.local.ntotal
        self.local.set_row_attr(attr[lo:hi])

    def query_slice(self, nq):
        """Rows [lo, hi) of an nq-query batch whose merged result this rank holds under ``result="slice"``."""
        return shard_bounds(nq, self.world, self.rank)

    def search(self, xq, k, *, exclude=None, attr_below=None, result="full"):
        """xq replicated on every rank.  ``result="full"``: (D, I) of every query on every rank (numpy in -> numpy
        out).  ``result="slice"``: the merge itself is sharded -- this rank gets the merged top-k of the queries
        ``query_slice(nq)`` only ([hi - lo, k]); 1/world of the exchange traffic, every answer exists on one rank."""
        return self.search_async(xq, k, exclude=exclude, attr_below=attr_below, result=result).result()

    def search_async(self, xq, k, *, exclude=None, attr_below=None, result="full"):
        """Local search now, exchange (all-gather + merge) queued on a side stream: the returned handle's
        ``result()`` orders the caller's stream after it.  Calling ``search_async`` for batch i+1 before
        ``result()`` of batch i lets the exchange of batch i -- and the wait for the slowest rank that comes
        with it -- overlap the local search of batch i+1.

        Measured caveat (2 x B200, 16M rows, batch 8192): with the current K2 -- persistent, one CTA per SM, corpus
        tiles statically assigned -- this is slower than the synchronous ``search`` (93 vs 72 ms per step): the NCCL
        kernel takes SMs and spins for the peer whose own NCCL kernel is queued behind its K2, and the displaced K2
        CTAs then run their whole static share late.  It pays only where the local search leaves SMs free (small
        corpora, exact path); a dynamic tile scheduler in K2 is what would make it pay in general."""
        as_numpy = not (isinstance(xq, torch.Tensor) and xq.is_cuda)
        if as_numpy and self._on_cuda:
            # the exchange runs over NCCL: keep the per-shard lists on the device, one H2D / D2H per call
            dev = torch.device("cuda", self.local.device)
            if isinstance(xq, torch.Tensor):
                xq = xq.detach().cpu().numpy()
            xq = torch.from_numpy(np.ascontiguousarray(xq, dtype=np.float32)).to(dev)
            if exclude is not None and not (isinstance(exclude, torch.Tensor) and exclude.is_cuda):
                exclude = torch.as_tensor(np.ascontiguousarray(exclude, dtype=np.int32)).to(dev)
        kw = {} if attr_below is None else {"attr_below": attr_below}
        if self._two_phase and self.world > 1 and (not self._on_cuda or (isinstance(xq, torch.Tensor) and xq.is_cuda)):
            D, I = self._local_search_two_phase(xq, k, exclude, kw)
        else:
            D, I = self.local.search(xq, k, exclude=exclude, **kw)   # complete on return (host-synchronised)
        if not isinstance(D, torch.Tensor):
            D, I = torch.from_numpy(D), torch.from_numpy(I)
        assert result in ("full", "slice")
        if not D.is_cuda:                                              # CPU test doubles: nothing to overlap
            Dm, Im = self.exchange(D, I, result=result)
            return PendingSearch(Dm, Im, None, as_numpy)
        if self._xstream is None:
            self._xstream = torch.cuda.Stream(device=D.device)
        cur = torch.cuda.current_stream(D.device)
        self._xstream.wait_stream(cur)
        with torch.cuda.stream(self._xstream):
            Dm, Im = self.exchange(D, I, result=result)
            ev = torch.cuda.Event()
            ev.record(self._xstream)
            self._xev = ev
        for t in (D, I):
            t.record_stream(self._xstream)
        return PendingSearch(Dm, Im, ev, as_numpy)

    def _local_search_two_phase(self, xq, k, exclude, kw):
        mb = int(self.local.get_option("max_batch"))
        nb = bounds_width(k, self.world)
        if self._xev is not None and self._on_cuda:      # the bounds exchange shares the export slots with the merge: keep them in order
            torch.cuda.current_stream(xq.device).wait_event(self._xev)
        outs = []
        for q0 in range(0, xq.shape[0], mb):
            xc = xq[q0:q0 + mb]
            xc = xc.contiguous() if isinstance(xc, torch.Tensor) else np.ascontiguousarray(xc)
            ec = None if exclude is None else exclude[q0:q0 + mb]
            payload = self.local.search_begin(xc, k, nb, exclude=ec, **kw)
            outs.append(self.local.search_finish(self.exchange_floor(payload, nb, k)))
        if len(outs) == 1:
            return outs[0]
        cat = torch.cat if isinstance(outs[0][0], torch.Tensor) else np.concatenate
        return cat([o[0] for o in outs]), cat([o[1] for o in outs])

    def exchange_floor(self, payload, nb, k):
        """The bounds exchange: every rank's [nq, nb + 1] payload -> floor[nq], a prefilter score below which no row
        of the global top-k can lie (k-th largest of the G * nb exchanged scores - 2 * the largest eps).  Peer mode:
        one kernel over NVLink peer memory (trx_exchange_floor); NCCL mode: all-gather + torch top-k."""
        nq = payload.shape[0]
        if self._exchange_mode == "peer":
            ex = self._peer_exchange(nq, k, payload.device)
            if ex is not None:
                floor = torch.empty((nq,), dtype=torch.float32, device=payload.device)
                _lib.check(_lib.lib().trx_exchange_floor(ex, payload.data_ptr(), nq, nb, k, floor.data_ptr(),
                                                         _stream_handle(payload.device)), "exchange_floor")
                return floor
        g = torch.empty((self.world,) + tuple(payload.shape), dtype=payload.dtype, device=payload.device)
        if payload.is_cuda:
            dist.all_gather_into_tensor(g, payload.contiguous(), group=self.group)
        else:             # gloo (CPU tests of the host logic)
            dist.all_gather(list(g.unbind(0)), payload.contiguous(), group=self.group)
        return floor_from_payloads(g, nb, k)

    def _peer_exchange(self, nq, k, device):
        """(Re)build the peer-memory exchange for at least nq*k entries.  Collective: every rank takes the same
        decisions (queries are replicated, so nq and k agree).  Returns None when peers cannot be mapped."""
        need = int(nq) * int(k)
        if self._peer is not None and need <= self._peer_entries:
            return self._peer
        L = _lib.lib()
        if self._peer is not None:
            torch.cuda.synchronize(device)
            dist.barrier(group=self.group)          # nobody still reads the old export buffers
            L.trx_exchange_destroy(self._peer)
            self._peer = None
        entries = max(need, int(os.environ.get("TRX_EXCHANGE_ENTRIES", 8192 * 100)))
        ex = ctypes.c_void_p()
        ok = L.trx_exchange_create(device.index, self.rank, self.world, entries, ctypes.byref(ex)) == 0
        handle = ctypes.create_string_buffer(64)
        ok = ok and L.trx_exchange_handle(ex, handle) == 0
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw) if ok else None, group=self.group)
        ok = ok and all(h is not None for h in handles)
        ok = ok and L.trx_exchange_connect(ex, b"".join(handles)) == 0
        flag = torch.tensor([1 if ok else 0], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:                   # some rank could not map its peers: everybody uses NCCL
            if ex:
                L.trx_exchange_destroy(ex)
            self._exchange_mode = "nccl"
            return None
        self._peer, self._peer_entries = ex, entries
        return ex

    def exchange(self, D, I, result="full"):
        """The one exchange step: the per-shard [nq, k] lists of every rank -> the merged top-k, on every rank
        (``result="full"``) or each rank its ``query_slice`` of it (``"slice"``).
        CUDA default: ONE kernel that gathers over NVLink peer memory and merges (K5p, k5_peer.cu);
        otherwise all-gather (NCCL / gloo) + k-way merge (K5)."""
        q_lo, q_hi = (0, D.shape[0]) if result == "full" else self.query_slice(D.shape[0])
        if D.is_cuda and self._exchange_mode == "peer" and self.world > 1:
            ex = self._peer_exchange(D.shape[0], D.shape[1], D.device)
            if ex is not None:
                D, I = D.contiguous(), I.contiguous()
                Dm = torch.empty((q_hi - q_lo, D.shape[1]), dtype=D.dtype, device=D.device)
                Im = torch.empty((q_hi - q_lo, I.shape[1]), dtype=I.dtype, device=I.device)
                _lib.check(_lib.lib().trx_exchange_merge_slice(ex, self.metric_type, D.data_ptr(), I.data_ptr(),
                                                               D.shape[0], D.shape[1], q_lo, q_hi - q_lo,
                                                               Dm.data_ptr(), Im.data_ptr(),
                                                               _stream_handle(D.device)), "exchange_merge")
                return Dm, Im
        Dg = torch.empty((self.world,) + tuple(D.shape), dtype=D.dtype, device=D.device)
        Ig = torch.empty((self.world,) + tuple(I.shape), dtype=I.dtype, device=I.device)
        if D.is_cuda:     # NCCL: one ncclAllGather each, straight into the [G, nq, k] buffers
            dist.all_gather_into_tensor(Dg, D.contiguous(), group=self.group)
            dist.all_gather_into_tensor(Ig, I.contiguous(), group=self.group)
        else:             # gloo (CPU tests of the host logic): list-of-views form
            dist.all_gather(list(Dg.unbind(0)), D.contiguous(), group=self.group)
            dist.all_gather(list(Ig.unbind(0)), I.contiguous(), group=self.group)
        # (the all-gather moves every list either way: a slice is the merged rows this rank keeps)
        return self._merge(Dg[:, q_lo:q_hi].contiguous(), Ig[:, q_lo:q_hi].contiguous(), self.metric_type)

    def close(self):
        if self._peer is not None:
            _lib.lib().trx_exchange_destroy(self._peer)
            self._peer = None
        if hasattr(self.local, "close"):
            self.local.close()
