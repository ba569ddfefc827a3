#!/usr/bin/env python
"""Small shapes through every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck):

  TRX_NO_TORCH=1 compute-sanitizer --tool memcheck python scripts/sanitize_driver.py

numpy only (torch is not imported: the sanitizer would instrument its start-up for minutes).  Results are checked
against the CPU oracle so a run that is "clean" but wrong still fails."""
import os
import sys

os.environ.setdefault("TRX_NO_TORCH", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import textreact_b200 as trx  # noqa: E402
from oracle import cpu_flat as oracle  # noqa: E402  (checker only)


def two_phase(idx, xb, rng, k, metric):
    """trx_search_begin -> floor (synthesised from the payload) -> trx_search_finish through the raw C ABI with device
    buffers from trx_device_malloc: the bounds kernel, K4 on presorted lists with a floor."""
    import ctypes
    from textreact_b200 import _lib
    L = _lib.lib()
    nq, nb, d = 150, 8, xb.shape[1]
    xq = rng.standard_normal((nq, d), dtype=np.float32)

    def dev(nbytes):
        p = ctypes.c_void_p()
        _lib.check(L.trx_device_malloc(0, nbytes, ctypes.byref(p)), "device_malloc")
        return p
    dq, dpay, dfl = dev(xq.nbytes), dev(nq * (nb + 1) * 4), dev(nq * 4)
    dD, dI = dev(nq * k * 4), dev(nq * k * 8)
    _lib.check(L.trx_device_copy(dq, xq.ctypes.data, xq.nbytes), "copy")
    _lib.check(L.trx_search_begin(idx._h, dq, nq, k, None, nb, dpay, None), "search_begin")
    pay = np.empty((nq, nb + 1), np.float32)
    _lib.check(L.trx_device_copy(pay.ctypes.data, dpay, pay.nbytes), "copy")
    assert (np.diff(pay[:, :nb], axis=1) <= 0).all()                       # best first
    floor = np.ascontiguousarray(pay[:, 4] - 2.0 * pay[:, nb], dtype=np.float32)   # as if ~k/5 rows per shard survive
    _lib.check(L.trx_device_copy(dfl, floor.ctypes.data, floor.nbytes), "copy")
    _lib.check(L.trx_search_finish(idx._h, dfl, dD, dI, None), "search_finish")
    D, I = np.empty((nq, k), np.float32), np.empty((nq, k), np.int64)
    _lib.check(L.trx_device_copy(D.ctypes.data, dD, D.nbytes), "copy")
    _lib.check(L.trx_device_copy(I.ctypes.data, dI, I.nbytes), "copy")
    Do, Io = oracle.search_blas(xb, xq, k, metric)
    s64 = oracle.scores_f64(xb, xq, metric)
    nv = (I >= 0).sum(1)
    assert (nv >= 5).all() and (nv < k).any()                              # short lists: only what is above the floor
    for i in range(nq):
        assert (I[i, :3] == Io[i, :3]).all() and (I[i, nv[i]:] == -1).all()  # the head of the exact answer ...
        got = D[i, :nv[i]]
        assert np.allclose(got, s64[i, I[i, :nv[i]]], rtol=1e-4, atol=1e-3)  # ... exact scores, best first
        assert (np.diff(got) <= 0).all() if metric == 0 else (np.diff(got) >= 0).all()
    for p in (dq, dpay, dfl, dD, dI):
        L.trx_device_free(p)


def big_groups(rng):
    """distinct-groups search with a group larger than the widest exact selection: the rounds of the exact path"""
    n, d, k = 9000, 32, 20
    xb = rng.integers(-3, 4, (n, d)).astype(np.float32)
    groups = (np.arange(n) + 10).astype(np.int32)
    big = rng.choice(n, 2500, replace=False)
    xb[big] = xb[big[0]]
    groups[big] = 1
    xq = (xb[big[0]][None, :] + rng.integers(-1, 2, (6, d))).astype(np.float32)
    idx = trx.IndexFlatIP(d, device=0)
    idx.add(xb)
    idx.set_groups(groups)
    D, I = idx.search(xq, k, dedup=True)
    g = groups[I]
    assert all(len(set(r.tolist())) == k for r in g) and (g[:, 0] == 1).all()
    idx.close()
    print("ok distinct groups, 2500-row group (rounds)", flush=True)


def main():
    rng = np.random.default_rng(7)
    n, d, k = 12000, 128, 10
    xb = rng.standard_normal((n, d), dtype=np.float32)
    groups = (np.arange(n) // 4).astype(np.int32)
    for metric in (0, 1):
        idx = trx.IndexFlat(d, metric, device=0)
        idx.add(xb[:5000]); idx.add(xb[5000:])
        idx.set_groups(groups)
        cases = [("umma_pair", trx.PATH_UMMA, 150), ("umma_single", trx.PATH_UMMA, 70), ("umma_small", trx.PATH_UMMA, 5),
                 ("stream", trx.PATH_STREAM, 3), ("exact", trx.PATH_EXACT, 9)]
        for name, path, nq in cases:
            xq = rng.standard_normal((nq, d), dtype=np.float32)
            excl = groups[rng.integers(0, n, nq)].astype(np.int32)
            idx.set_option("path", path)
            D, I = idx.search(xq, k, exclude=excl)
            oracle.check_parity(D, I, xb, xq, k, metric, groups, excl)
            print("ok", "L2" if metric else "IP", name, flush=True)
        idx.set_option("path", trx.PATH_UMMA)
        idx.set_option("target_candidates", 32)       # force uncertified queries -> every fallback route
        xq = rng.standard_normal((90, d), dtype=np.float32)
        for sp in (1, 0):                             # batched second tcgen05 pass / fp32 sweep per 4 queries
            idx.set_option("second_pass", sp)
            D, I = idx.search(xq, k)
            oracle.check_parity(D, I, xb, xq, k, metric)
            st = idx.stats()
            print("ok", "L2" if metric else "IP", "fallbacks second_pass=%d" % sp, st["queries_second_pass"], st["queries_exact"], flush=True)
        idx.set_option("second_pass", 1)
        idx.set_option("target_candidates", 768)
        two_phase(idx, xb, rng, k, metric)
        print("ok", "L2" if metric else "IP", "two-phase begin / finish", flush=True)
        D, I = idx.search_self(k, 100, 140)
        assert (I[:, 0] == np.arange(100, 140)).all()
        idx.close()
    xi = (rng.random((9000, 64)) < 0.1).astype(np.int8)          # typed ingestion
    idx = trx.IndexFlatL2(64, device=0)
    idx.add(xi)
    D, I = idx.search(xi[:20], 5)
    Do, Io = oracle.search_seq(xi, xi[:20], 5, 1)
    assert (I == Io).all() and (D == Do).all()
    idx.close()
    big_groups(rng)
    print("sanitize driver done")


if __name__ == "__main__":
    main()
