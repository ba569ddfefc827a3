"""CPU: the oracle (FAISS flat-search restatement) against the committed golden vectors and against
itself (scalar path vs BLAS path vs float64 arbiter)."""
import glob
import os

import numpy as np
import pytest

from oracle import cpu_flat as oracle
from tests import util

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def load(path):
    z = np.load(path)
    g = z["groups"] if "groups" in z.files else None
    e = z["excl"] if "excl" in z.files else None
    return z["xb"], z["xq"], int(z["k"]), int(z["metric"]), g, e, z["D"], z["I"]


def test_golden_files_present():
    assert len(GOLDEN) >= 8


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
@pytest.mark.parametrize("fn", ["seq", "blas"])
def test_oracle_matches_golden(path, fn):
    xb, xq, k, metric, g, e, D, I = load(path)
    search = oracle.search_seq if fn == "seq" else oracle.search_blas
    Do, Io = search(xb, xq, k, metric, g, e)
    np.testing.assert_array_equal(Io, I)
    np.testing.assert_allclose(Do, D, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("metric", [0, 1])
def test_paths_agree_and_pass_parity_rule(metric):
    xb, xq = util.gaussian(8000, 96, 1), util.gaussian(40, 96, 2)
    D1, I1 = oracle.search_seq(xb, xq, 15, metric)
    D2, I2 = oracle.search_blas(xb, xq, 15, metric, bs_q=16, bs_b=1000)   # odd blocking on purpose
    r1 = oracle.check_parity(D1, I1, xb, xq, 15, metric)
    r2 = oracle.check_parity(D2, I2, xb, xq, 15, metric)
    assert r1["forced_ranks"] > 0 and r2["forced_ranks"] > 0
    assert (I1 == I2).mean() > 0.99


def test_dispatch_threshold_like_faiss():
    xb = util.gaussian(500, 16, 3)
    for nq in (1, 19, 20, 33):
        xq = util.gaussian(nq, 16, 4)
        D, I = oracle.search(xb, xq, 5)
        oracle.check_parity(D, I, xb, xq, 5)


def test_int_inputs_are_coerced_like_faiss_wrapper():
    xb = util.fingerprints(300, 64, 5)             # int8, as morgan_fingerprint produces
    xq = util.count_fingerprints(7, 64, 6)          # int64, as reaction_fingerprint_array produces
    D, I = oracle.search_seq(xb, xq, 4, 1)
    D2, I2 = oracle.search_seq(xb.astype(np.float32), xq.astype(np.float32), 4, 1)
    np.testing.assert_array_equal(I, I2)
    np.testing.assert_array_equal(D, D2)


def test_ties_resolved_by_ascending_id():
    xb = np.ones((10, 4), np.float32)
    D, I = oracle.search_seq(xb, np.ones((2, 4), np.float32), 4, 0)
    np.testing.assert_array_equal(I, [[0, 1, 2, 3]] * 2)
    D, I = oracle.search_blas(xb, np.ones((25, 4), np.float32), 4, 1, bs_b=3)
    np.testing.assert_array_equal(I, [[0, 1, 2, 3]] * 25)


def test_padding_when_k_exceeds_rows():
    xb, xq = util.gaussian(3, 8, 7), util.gaussian(2, 8, 8)
    D, I = oracle.search_seq(xb, xq, 5, 0)
    assert (I[:, 3:] == -1).all() and (D[:, 3:] == np.float32(-oracle.FLT_MAX)).all()
    D, I = oracle.search_seq(xb, xq, 5, 1)
    assert (I[:, 3:] == -1).all() and (D[:, 3:] == np.float32(oracle.FLT_MAX)).all()


def test_mask_equals_post_filter():
    """engine-side exclusion == the consumer filter of textreact/dataset.py:74-76 applied to a deeper list"""
    n, g = 2000, 5
    xb, xq = util.clustered_unit(n, 48, 9, ncent=20), util.clustered_unit(30, 48, 10, ncent=20)
    groups = (np.arange(n) // g).astype(np.int32)
    excl = groups[np.random.default_rng(11).integers(0, n, 30)].copy()
    excl[::4] = -1
    k = 10
    D, I = oracle.search_seq(xb, xq, k, 0, groups, excl)
    D2, I2 = oracle.search_seq(xb, xq, k + g, 0)
    corpus_text = {j: f"text-{groups[j]}" for j in range(n)}        # one paragraph per group
    for i in range(30):
        gold = None if excl[i] < 0 else f"text-{excl[i]}"
        ids = [j for j in I2[i].tolist() if gold is None or corpus_text[j] != gold][:k]
        np.testing.assert_array_equal(I[i], ids)


def test_post_filter_semantics():
    corpus = {"a": "T1", "b": "T1", "c": "T2", "d": "T3"}
    assert oracle.post_filter(["a", "zz", "b", "c", "d"], corpus) == ["a", "c", "d"]            # unknown id + dedup
    assert oracle.post_filter(["a", "b", "c", "d"], corpus, gold_text="T1") == ["c", "d"]        # gold skip
    assert oracle.post_filter(["a", "b", "c", "d"], corpus, gold_text="T2", num_neighbors=1) == ["a"]


def test_c1_shape_against_an_independent_blas_topk():
    """BASELINE.json configs[0] (the reference's own CPU-runnable case): 100K x 768 fp32 corpus, 1K queries, k=20.
    The oracle's FAISS-restatement (sgemm blocks + k-heap in C) against an independent route (torch/MKL matmul +
    torch.topk): same ids wherever the fp32 gap at the boundary is not a rounding-level tie, same scores."""
    import torch
    from tests import util
    xb, xq = util.gaussian(100_000, 768, 11), util.gaussian(1000, 768, 12)
    D, I = oracle.search_blas(xb, xq, 20, 0)
    s = torch.from_numpy(xq) @ torch.from_numpy(xb).T
    Dt, It = torch.topk(s, 21, dim=1)
    Dt, It = Dt.numpy(), It.numpy()
    np.testing.assert_allclose(D, Dt[:, :20], rtol=2e-5, atol=2e-4)
    gap = np.abs(Dt[:, :-1] - Dt[:, 1:]) / np.maximum(np.abs(Dt[:, :-1]), 1e-30)
    clear = gap > 1e-5                                     # rank j and j+1 are distinguishable in fp32
    same = I == It[:, :20]
    # a position may differ only if it sits next to a rounding-level tie
    near_tie = ~clear[:, :20] | np.concatenate([np.zeros((1000, 1), bool), ~clear[:, :19]], axis=1)
    assert (same | near_tie).all()
    assert same.mean() > 0.999
    oracle.check_parity(D[:16], I[:16], xb, xq[:16], 20, 0)


def _real_faiss():
    try:
        import faiss
    except Exception:
        return None
    return None if str(getattr(faiss, "__version__", "")).startswith("textreact_b200") else faiss


def test_oracle_against_real_faiss_when_it_is_installed():
    """Pins the restatement to the real thing wherever `faiss` is importable (it is not in the build image: the
    reference neither vendors nor pins it).  IndexFlatIP / IndexFlatL2, BLAS and scalar paths, int8 inputs."""
    faiss = _real_faiss()
    if faiss is None:
        pytest.skip("faiss is not installed (un-vendored, un-pinned by the reference; no network here)")
    from tests import util
    xb, xq = util.gaussian(20000, 96, 31), util.gaussian(64, 96, 32)
    for metric, cls in ((0, faiss.IndexFlatIP), (1, faiss.IndexFlatL2)):
        for q in (xq, xq[:7]):                                # nq >= 20 -> BLAS path, < 20 -> scalar path
            index = cls(96)
            index.add(xb)
            Df, If = index.search(q, 10)
            oracle.check_parity(Df, If, xb, q, 10, metric)     # FAISS itself satisfies the north_star rule ...
            Do, Io = oracle.search(xb, q, 10, metric)
            assert (Io == If).mean() > 0.999                   # ... and the restatement agrees with it
            np.testing.assert_allclose(Do, Df, rtol=2e-5, atol=2e-4)
    fb = util.fingerprints(5000, 128, 33)
    index = faiss.IndexFlatL2(128)
    index.add(np.ascontiguousarray(fb, dtype="float32"))
    Df, If = index.search(np.ascontiguousarray(fb[:30], dtype="float32"), 5)
    Do, Io = oracle.search(fb, fb[:30], 5, 1)
    np.testing.assert_array_equal(Do, Df)                      # integer distances: exact


def test_oracle_against_scikit_learn_brute_force():
    """A third, unrelated implementation (scikit-learn's brute-force kNN: pairwise distances in its own Cython/BLAS
    code + argpartition) agrees with the restatement on L2 -- ids wherever ranks are separated, distances to 1e-5 --
    and on 0/1 fingerprints, where distances are integers, exactly (ids compared as tie-free sets)."""
    sk = pytest.importorskip("sklearn.neighbors")
    xb, xq = util.gaussian(20000, 96, 41), util.gaussian(64, 96, 42)
    nn = sk.NearestNeighbors(n_neighbors=10, algorithm="brute", metric="sqeuclidean").fit(xb.astype(np.float64))
    Ds, Is = nn.kneighbors(xq.astype(np.float64))
    Do, Io = oracle.search(xb, xq, 10, 1)
    assert (Io == Is).mean() > 0.999
    np.testing.assert_allclose(Do, Ds, rtol=2e-5, atol=2e-4)
    fb = util.fingerprints(4000, 128, 43).astype(np.float64)
    nn = sk.NearestNeighbors(n_neighbors=5, algorithm="brute", metric="sqeuclidean").fit(fb)
    Ds, Is = nn.kneighbors(fb[:30])
    Do, Io = oracle.search(fb, fb[:30], 5, 1)
    np.testing.assert_array_equal(Do, Ds.astype(np.float32))              # integer distances: exact
