"""oracle/cpu_flat.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the flat-index search TextReact's retrieval step delegates to the
third-party ``faiss`` wheel (reference call sites: retrieve/retrieve_faiss.py:65 ctor,
:66 ``index.add``, :71 ``index.search``; neural 768-d inner-product variant described at
README.md:44-47 whose output enters through retrieve/convert_format.py:7-16).

PARITY UNPINNED.  ``faiss`` is not vendored, not pinned (absent from environment.yml) and
not importable in this image; the reference has no tests or golden vectors for this path
(SURVEY.md section 4 / 8c); the one attempt to fetch faiss-cpu on a GPU box found no index
reachable either (profiles/round2_a_try_faiss_on_gpu_box.log).  tests/test_oracle.py checks
this restatement against two unrelated implementations that ARE here (torch/MKL matmul + topk,
scikit-learn's brute-force kNN) and, automatically, against the real faiss wherever it imports
(tests/golden/make_faiss_golden.py then produces real-FAISS known-answer vectors).
This file restates FAISS's published algorithm:

* ``search_blas``  -- the nq >= 20 path: fp32 ``sgemm`` over query-block x database-block
  tiles (FAISS defaults 4096 x 1024; any blocking gives the same scores because the
  reduction runs over d only), each score block fed to a per-query k-heap
  (``trx_oracle_heap_block`` in cpu_flat.c).  L2 uses ``|x|^2 + |y|^2 - 2 x.y`` clamped
  at 0 exactly as the BLAS path does.
* ``search_seq``   -- the nq < 20 path: direct fp32 dot / sum of squared differences.
* ``search_f64``   -- float64 scores; the arbiter used to compute the rank-k gap that
  decides where ids MUST agree (north_star: relative gap > 1e-5).
* ``post_filter``  -- the gold-removed / dedup consumer restated from
  textreact/dataset.py:46-56 (dedup) and :74-78 (gold skip, head num_neighbors).

Result convention everywhere: best first, ties by ascending id, unfilled slots
``I = -1`` with ``D = -FLT_MAX`` (IP) / ``+FLT_MAX`` (L2).

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference``
legs may import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1
FLT_MAX = float(np.finfo(np.float32).max)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    """Load (building on first use) the C half of the oracle."""
    global _LIB
    if _LIB is not None:
        return _LIB
    so = os.path.join(_HERE, "libtrx_oracle.so")
    src = os.path.join(_HERE, "cpu_flat.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    lib = ctypes.CDLL(so)
    f32p = ctypes.POINTER(ctypes.c_float)
    i64p = ctypes.POINTER(ctypes.c_int64)
    i32p = ctypes.POINTER(ctypes.c_int32)
    f64p = ctypes.POINTER(ctypes.c_double)
    lib.trx_oracle_heap_init.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int, f32p, i64p]
    lib.trx_oracle_heap_block.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                          ctypes.c_int, f32p, i32p, i32p, f32p, i64p]
    lib.trx_oracle_heap_reorder.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int, f32p, i64p]
    lib.trx_oracle_search_seq.argtypes = [ctypes.c_int, f32p, ctypes.c_int64, f32p, ctypes.c_int64,
                                          ctypes.c_int64, ctypes.c_int, i32p, i32p, f32p, i64p]
    lib.trx_oracle_scores_f64.argtypes = [ctypes.c_int, f32p, ctypes.c_int64, f32p, ctypes.c_int64,
                                          ctypes.c_int64, f64p]
    for fn in (lib.trx_oracle_heap_init, lib.trx_oracle_heap_block, lib.trx_oracle_heap_reorder,
               lib.trx_oracle_search_seq, lib.trx_oracle_scores_f64):
        fn.restype = None
    _LIB = lib
    return lib


def _p(a, ty):
    return a.ctypes.data_as(ctypes.POINTER(ty)) if a is not None else None


def coerce(x):
    """FAISS's python wrapper coerces every input with
    ``np.ascontiguousarray(x, dtype='float32')`` -- the reference feeds int64 and int8
    fingerprints (retrieve/retrieve_faiss.py:26, :40)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    assert x.ndim == 2
    return x


def _masks(groups, exclude, nb, nq):
    g = None if groups is None else np.ascontiguousarray(groups, dtype=np.int32)
    e = None if exclude is None else np.ascontiguousarray(exclude, dtype=np.int32)
    if e is not None:
        assert g is not None and g.shape == (nb,) and e.shape == (nq,)
    return g, e


def search_blas(xb, xq, k, metric=METRIC_INNER_PRODUCT, groups=None, exclude=None,
                bs_q=4096, bs_b=1024 * 16, gemm="numpy"):
    """FAISS BLAS path (exhaustive_inner_product_blas / exhaustive_L2sqr_blas).
    gemm: "numpy" (the BLAS numpy links, OpenBLAS here) or "torch" (torch.mm: MKL sgemm) -- same scores up to
    the summation order, used by bench.py's CPU arm to time the faster of the two."""
    xb, xq = coerce(xb), coerce(xq)
    if gemm == "torch":
        import torch
        tb = torch.from_numpy(xb)
    nb, d = xb.shape
    nq = xq.shape[0]
    assert xq.shape[1] == d and k > 0
    g, e = _masks(groups, exclude, nb, nq)
    lib = _lib()
    D = np.empty((nq, k), np.float32)
    I = np.empty((nq, k), np.int64)
    lib.trx_oracle_heap_init(metric, nq, k, _p(D, ctypes.c_float), _p(I, ctypes.c_int64))
    if metric == METRIC_L2:
        bn = np.einsum("ij,ij->i", xb, xb, dtype=np.float32)
        qn = np.einsum("ij,ij->i", xq, xq, dtype=np.float32)
    for q0 in range(0, nq, bs_q):
        q1 = min(nq, q0 + bs_q)
        Dq, Iq = D[q0:q1], I[q0:q1]        # contiguous row slices
        eq = None if e is None else np.ascontiguousarray(e[q0:q1])
        for b0 in range(0, nb, bs_b):
            b1 = min(nb, b0 + bs_b)
            if gemm == "torch":
                s = torch.mm(torch.from_numpy(xq[q0:q1]), tb[b0:b1].T).numpy()
            else:
                s = xq[q0:q1] @ xb[b0:b1].T     # fp32 sgemm
            if metric == METRIC_L2:
                s = qn[q0:q1, None] + bn[None, b0:b1] - 2.0 * s
                np.maximum(s, 0.0, out=s)
            s = np.ascontiguousarray(s, dtype=np.float32)
            lib.trx_oracle_heap_block(metric, q1 - q0, b1 - b0, b0, k, _p(s, ctypes.c_float),
                                      _p(g, ctypes.c_int32), _p(eq, ctypes.c_int32),
                                      _p(Dq, ctypes.c_float), _p(Iq, ctypes.c_int64))
    lib.trx_oracle_heap_reorder(metric, nq, k, _p(D, ctypes.c_float), _p(I, ctypes.c_int64))
    return D, I


def search_seq(xb, xq, k, metric=METRIC_INNER_PRODUCT, groups=None, exclude=None):
    """FAISS scalar path (nq < 20): direct per-pair arithmetic, no norm expansion."""
    xb, xq = coerce(xb), coerce(xq)
    nb, d = xb.shape
    nq = xq.shape[0]
    assert xq.shape[1] == d and k > 0
    g, e = _masks(groups, exclude, nb, nq)
    D = np.empty((nq, k), np.float32)
    I = np.empty((nq, k), np.int64)
    _lib().trx_oracle_search_seq(metric, _p(xb, ctypes.c_float), nb, _p(xq, ctypes.c_float), nq, d, k,
                                 _p(g, ctypes.c_int32), _p(e, ctypes.c_int32),
                                 _p(D, ctypes.c_float), _p(I, ctypes.c_int64))
    return D, I


def search(xb, xq, k, metric=METRIC_INNER_PRODUCT, groups=None, exclude=None):
    """Dispatch as FAISS does: nq < 20 -> scalar path, else BLAS path."""
    nq = np.asarray(xq).shape[0]
    fn = search_seq if nq < 20 else search_blas
    return fn(xb, xq, k, metric, groups, exclude)


def scores_f64(xb, xq, metric=METRIC_INNER_PRODUCT):
    xb, xq = coerce(xb), coerce(xq)
    out = np.empty((xq.shape[0], xb.shape[0]), np.float64)
    _lib().trx_oracle_scores_f64(metric, _p(xb, ctypes.c_float), xb.shape[0], _p(xq, ctypes.c_float),
                                 xq.shape[0], xb.shape[1], _p(out, ctypes.c_double))
    return out


def search_f64(xb, xq, k, metric=METRIC_INNER_PRODUCT, groups=None, exclude=None, extra=1):
    """float64 arbiter.  Returns (D64 [nq,k+extra], I [nq,k+extra]) so callers can read
    the gap between rank k and rank k+1.  Slots beyond the eligible rows: I=-1, D=nan."""
    s = scores_f64(xb, xq, metric)
    nq, nb = s.shape
    if exclude is not None:
        g = np.asarray(groups, dtype=np.int32)
        e = np.asarray(exclude, dtype=np.int32)
        bad = (g[None, :] == e[:, None]) & (e[:, None] >= 0)
        s = np.where(bad, -np.inf if metric == METRIC_INNER_PRODUCT else np.inf, s)
    kk = k + extra
    key = -s if metric == METRIC_INNER_PRODUCT else s
    D = np.full((nq, kk), np.nan)
    I = np.full((nq, kk), -1, np.int64)
    ids = np.arange(nb)
    for i in range(nq):
        order = np.lexsort((ids, key[i]))[:kk]          # primary key score, secondary id
        order = order[np.isfinite(key[i][order])]
        D[i, :len(order)] = s[i][order]
        I[i, :len(order)] = order
    return D, I


def check_parity(D, I, xb, xq, k, metric=METRIC_INNER_PRODUCT, groups=None, exclude=None,
                 rtol=1e-5, D64=None, I64=None):
    """The north_star acceptance rule, as an executable check.

    * scores: |D - D64| <= rtol * scale, scale = max(|D64|, rtol-floor) for IP; for L2 the
      scale is |q|^2 + |x|^2 (the magnitude the subtraction works at -- FAISS's own BLAS
      path has that absolute error).
    * ids: the SET of the first j ids must equal the fp64 set at every rank j where the
      fp64 relative gap between rank j and rank j+1 exceeds rtol; inside a tie group
      (gap <= rtol) any order / member is accepted.
    Returns a dict with counts; raises AssertionError on violation."""
    xb, xq = coerce(xb), coerce(xq)
    if D64 is None:
        D64, I64 = search_f64(xb, xq, k, metric, groups, exclude, extra=1)
    nq = xq.shape[0]
    D = np.asarray(D); I = np.asarray(I)
    assert D.shape == (nq, k) and I.shape == (nq, k), (D.shape, I.shape)
    assert D.dtype == np.float32 and I.dtype == np.int64
    qn = np.einsum("ij,ij->i", xq.astype(np.float64), xq.astype(np.float64))
    bn = np.einsum("ij,ij->i", xb.astype(np.float64), xb.astype(np.float64))
    n_forced = n_tied = 0
    for i in range(nq):
        valid = int((I64[i, :k] >= 0).sum())
        assert (I[i, valid:] == -1).all(), f"query {i}: padding ids"
        fill = -FLT_MAX if metric == METRIC_INNER_PRODUCT else FLT_MAX
        assert (D[i, valid:] == np.float32(fill)).all(), f"query {i}: padding scores"
        assert (I[i, :valid] >= 0).all() and len(set(I[i, :valid].tolist())) == valid, f"query {i}: ids not unique"
        ref = D64[i, :valid]
        if metric == METRIC_INNER_PRODUCT:
            scale = np.sqrt(qn[i] * bn[np.maximum(I[i, :valid], 0)])   # |q||x| >= |q.x|
        else:
            scale = qn[i] + bn[np.maximum(I[i, :valid], 0)]
        # score of the id we returned, in fp64, must match what we reported
        if metric == METRIC_INNER_PRODUCT:
            own = xb[I[i, :valid]].astype(np.float64) @ xq[i].astype(np.float64)
        else:
            diff = xb[I[i, :valid]].astype(np.float64) - xq[i].astype(np.float64)
            own = np.einsum("ij,ij->i", diff, diff)
        err = np.abs(D[i, :valid].astype(np.float64) - own)
        assert (err <= rtol * np.maximum(scale, 1e-30)).all(), \
            f"query {i}: score error {err.max():.3e} vs tol {rtol}*scale"
        # rank-wise agreement with the fp64 ordering
        assert (np.abs(D[i, :valid] - ref) <= rtol * np.maximum(scale, 1e-30) + rtol * np.abs(ref)).all(), \
            f"query {i}: score at rank differs from fp64 oracle"
        full = D64[i]
        for j in range(valid):
            nxt = full[j + 1] if j + 1 < full.shape[0] and I64[i, j + 1] >= 0 else None
            if nxt is None:
                gap_ok = True                    # nothing beyond: the set is forced
            else:
                denom = max(abs(full[j]), abs(nxt), 1e-30)
                gap_ok = abs(full[j] - nxt) / denom > rtol
            if gap_ok:
                n_forced += 1
                assert set(I[i, :j + 1].tolist()) == set(I64[i, :j + 1].tolist()), \
                    f"query {i}: top-{j + 1} id set differs from fp64 oracle across a gap > {rtol}"
            else:
                n_tied += 1
        # order inside our own result must be best-first
        dd = D[i, :valid]
        if metric == METRIC_INNER_PRODUCT:
            assert (dd[:-1] >= dd[1:]).all(), f"query {i}: D not descending"
        else:
            assert (dd[:-1] <= dd[1:]).all(), f"query {i}: D not ascending"
        same = dd[:-1] == dd[1:]
        assert (I[i, :valid][:-1][same] < I[i, :valid][1:][same]).all(), f"query {i}: equal scores not id-ascending"
    return {"forced_ranks": n_forced, "tied_ranks": n_tied}


def post_filter(nn_ids, corpus_text, gold_text=None, num_neighbors=None):
    """Consumer-side filter restated from textreact/dataset.py:
    keep ids present in the corpus (:60); if gold_text is given drop every id whose text
    equals it (:74-76); de-duplicate by text keeping first occurrence (:46-56, :77);
    take the head (:78)."""
    ids = [i for i in nn_ids if i in corpus_text]
    if gold_text is not None:
        ids = [i for i in ids if corpus_text[i] != gold_text]
    out = []
    for i in ids:
        if not any(corpus_text[i] == corpus_text[j] for j in out):
            out.append(i)
    return out if num_neighbors is None else out[:num_neighbors]
