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
        idx.set_option("target_candidates", 32)       # force uncertified queries -> both fallback routes
        xq = rng.standard_normal((90, d), dtype=np.float32)
        D, I = idx.search(xq, k)
        oracle.check_parity(D, I, xb, xq, k, metric)
        print("ok", "L2" if metric else "IP", "fallbacks", idx.stats()["queries_exact"], flush=True)
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
    print("sanitize driver done")


if __name__ == "__main__":
    main()
