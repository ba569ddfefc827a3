"""Regenerates tests/golden/*.npz.  Run from the repo root:  python tests/golden/make_golden.py

The reference (thomas0809/textreact) ships no tests or golden vectors for its retrieval path and its
arithmetic lives in the un-vendored `faiss` wheel (not importable here), so these known-answer
vectors are produced by the oracle's float64 arbiter (oracle/cpu_flat.py:search_f64) on small seeded
inputs chosen so that every rank gap is far above fp32 rounding (or the arithmetic is exact integer
arithmetic): any correct flat search -- FAISS included -- must reproduce the ids exactly and the
scores to 1e-5.  Inputs are stored with the answers so the fixtures do not depend on RNG streams."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cpu_flat as oracle  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def well_separated(xb, xq, k, metric, groups=None, excl=None, min_gap=1e-3):
    """Keep only queries whose first k+1 fp64 scores are pairwise separated by > min_gap relative."""
    D, I = oracle.search_f64(xb, xq, k, metric, groups, excl, extra=1)
    keep = []
    for i in range(xq.shape[0]):
        d = D[i][I[i] >= 0]
        gaps = np.abs(np.diff(d)) / np.maximum(np.abs(d[:-1]), 1e-30)
        if len(gaps) == 0 or gaps.min() > min_gap:
            keep.append(i)
    return np.array(keep)


def save(name, xb, xq, k, metric, groups=None, excl=None, exact_ties=False):
    if not exact_ties:
        keep = well_separated(xb, xq, k, metric, groups, excl)
        xq = xq[keep]
        if excl is not None:
            excl = excl[keep]
        D, I = oracle.search_f64(xb, xq, k, metric, groups, excl, extra=0)
        D = np.where(I >= 0, D, -oracle.FLT_MAX if metric == 0 else oracle.FLT_MAX).astype(np.float32)
    else:  # integer-valued data: fp32 arithmetic is exact, ties resolved by ascending id
        D, I = oracle.search_seq(xb, xq, k, metric, groups, excl)
        D64, I64 = oracle.search_f64(xb, xq, k, metric, groups, excl, extra=0)
        assert (I == I64).all()
        assert np.allclose(D, np.where(I64 >= 0, D64, D), rtol=1e-6, atol=1e-6)
    arrs = dict(xb=xb, xq=xq, k=np.int64(k), metric=np.int64(metric), D=D, I=I)
    if groups is not None:
        arrs.update(groups=groups, excl=excl)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)
    print(name, "nq", xq.shape[0], "nb", xb.shape[0], "d", xb.shape[1], "k", k)


def main():
    rng = np.random.default_rng(20240607)
    xb = rng.standard_normal((600, 32)).astype(np.float32)
    xq = rng.standard_normal((24, 32)).astype(np.float32)
    save("kat_ip_gauss", xb, xq, 10, 0)
    save("kat_l2_gauss", xb, xq, 10, 1)
    # Morgan-bit-like 0/1 int8 rows, self retrieval, heavy ties (retrieve_faiss.py:36-44, :114-115)
    fb = (rng.random((400, 128)) < 0.08).astype(np.int8)
    save("kat_l2_bits_ties", fb, fb[:16], 20, 1, exact_ties=True)
    save("kat_ip_bits_ties", fb, fb[:16], 20, 0, exact_ties=True)
    # difference-fingerprint-like signed counts as int64 (retrieve_faiss.py:18-27)
    cb = (rng.integers(-2, 3, (300, 256)) * (rng.random((300, 256)) < 0.05)).astype(np.int64)
    save("kat_l2_counts_int64", cb, cb[:12], 20, 1, exact_ties=True)
    # k > ntotal: -1 / FLT_MAX padding
    save("kat_ip_k_gt_n", xb[:7], xq[:5], 12, 0)
    # gold-removed mode: groups of 5 rows, some queries without exclusion
    groups = (np.arange(600) // 5).astype(np.int32)
    excl = groups[rng.integers(0, 600, 24)].astype(np.int32)
    excl[::5] = -1
    save("kat_ip_masked", xb, xq, 10, 0, groups, excl)
    # duplicate rows: exact ties between different ids
    dup = np.concatenate([xb[:50], xb[:50], xb[:50]])
    save("kat_ip_duplicates", dup, xb[:8], 6, 0, exact_ties=True)


if __name__ == "__main__":
    main()
