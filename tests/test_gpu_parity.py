"""GPU parity: the CUDA engine (through the C ABI) against the CPU oracle, north_star rule:
ids identical wherever the fp64 rank gap exceeds 1e-5 relative, scores within 1e-5."""
import numpy as np
import pytest

from oracle import cpu_flat as oracle
from tests import util

pytestmark = pytest.mark.gpu

IP, L2 = 0, 1


def _engine():
    import textreact_b200 as trx
    return trx


def _run(xb, xq, k, metric, path, groups=None, exclude=None, **opts):
    trx = _engine()
    idx = trx.IndexFlat(xb.shape[1], metric)
    idx.add(xb)
    idx.set_option("path", path)
    for key, v in opts.items():
        idx.set_option(key, v)
    if groups is not None:
        idx.set_groups(groups)
    D, I = idx.search(xq, k, exclude=exclude)
    st = idx.stats()
    idx.close()
    return D, I, st


@pytest.mark.parametrize("metric", [IP, L2])
@pytest.mark.parametrize("shape", [(1000, 64, 7, 10), (5000, 100, 33, 20), (3001, 768, 5, 100)])
def test_exact_path_small(metric, shape):
    n, d, nq, k = shape
    xb, xq = util.gaussian(n, d, 1), util.gaussian(nq, d, 2)
    D, I, st = _run(xb, xq, k, metric, _engine().PATH_EXACT)
    oracle.check_parity(D, I, xb, xq, k, metric)
    assert st["queries_exact"] == nq


@pytest.mark.parametrize("pair", [0, 1])
def test_umma_raw_scores_match_bf16_matmul(pair):
    """K2 in STORE mode against torch: validates TMA swizzle, UMMA descriptors, TMEM readout,
    for the single-CTA (cta_group::1) and the CTA-pair (cta_group::2) tilings."""
    import torch
    trx = _engine()
    n, d, nq = 20000, 768, 300
    xb, xq = util.gaussian(n, d, 3), util.gaussian(nq, d, 4)
    idx = trx.IndexFlatIP(d)
    idx.set_option("umma_pair", pair)
    idx.add(xb)
    got = idx.debug_scores_umma(torch.from_numpy(xq).cuda(), 256, n - 256 - 77).cpu().numpy()
    idx.close()
    ref = (torch.from_numpy(xq).bfloat16().double() @ torch.from_numpy(xb[256:n - 77]).bfloat16().double().T).numpy()
    err = np.abs(got - ref).max()
    assert err < 2e-3, err


@pytest.mark.parametrize("metric", [IP, L2])
@pytest.mark.parametrize("path_name", ["stream", "umma", "umma_single_cta"])
def test_prefilter_paths_gaussian(metric, path_name):
    trx = _engine()
    path = {"stream": trx.PATH_STREAM, "umma": trx.PATH_UMMA, "umma_single_cta": trx.PATH_UMMA}[path_name]
    n, d, k = 60000, 768, 20
    nq = 6 if path_name == "stream" else 300
    xb, xq = util.gaussian(n, d, 5), util.gaussian(nq, d, 6)
    D, I, st = _run(xb, xq, k, metric, path, umma_pair=0 if path_name == "umma_single_cta" else 1)
    oracle.check_parity(D, I, xb, xq, k, metric)
    assert st["last_path"] == path
    assert st["queries_exact"] <= nq // 10, st      # the certificate should hold for almost all


@pytest.mark.parametrize("metric", [IP, L2])
@pytest.mark.parametrize("second_pass", [1, 0])
def test_uncertified_queries_take_exact_second_chance(metric, second_pass):
    """Starve the candidate lists so the certificate fails for many queries: they must still come back
    exact -- through the batched second tcgen05 pass whose threshold (k-th exact score seen - eps) makes the
    candidate list complete (default), or the threshold-guided fp32 sweep (second_pass=0), or the generic scan."""
    trx = _engine()
    n, d, nq, k = 60000, 768, 200, 20
    xb, xq = util.gaussian(n, d, 15), util.gaussian(nq, d, 16)
    D, I, st = _run(xb, xq, k, metric, trx.PATH_UMMA, target_candidates=32, second_pass=second_pass)
    oracle.check_parity(D, I, xb, xq, k, metric)
    assert st["queries_uncert"] > 0, st
    if second_pass:
        # every query K4 left a bound for (>= k candidates seen) is answered by the second pass; the ones with fewer
        # than k candidates (target 32 < 2k) have no bound and scan
        assert st["queries_second_pass"] > 0 and st["queries_second_pass"] + st["queries_exact"] == st["queries_uncert"], st
    else:
        assert st["queries_second_pass"] == 0 and st["queries_exact"] > 0, st


def test_config1_c1_shape():
    """BASELINE.json configs[0]: 100K x 768 fp32 corpus, 1K queries, k=20, inner product."""
    trx = _engine()
    xb, xq = util.gaussian(100_000, 768, 11), util.gaussian(1000, 768, 12)
    D, I, st = _run(xb, xq, 20, IP, trx.PATH_AUTO)
    Do, Io = oracle.search_blas(xb, xq, 20, IP)
    # fp32-vs-fp32 agreement with the FAISS restatement, then the fp64 rule on a slice
    assert (I == Io).mean() > 0.999
    np.testing.assert_allclose(D, Do, rtol=2e-5, atol=2e-4)
    oracle.check_parity(D[:64], I[:64], xb, xq[:64], 20, IP)
    assert st["last_path"] == trx.PATH_UMMA


@pytest.mark.parametrize("metric", [IP, L2])
def test_clustered_unit_k100(metric):
    trx = _engine()
    xb, xq = util.clustered_unit(50000, 768, 21), util.clustered_unit(200, 768, 22)
    D, I, st = _run(xb, xq, 100, metric, trx.PATH_UMMA)
    oracle.check_parity(D, I, xb, xq, 100, metric)


@pytest.mark.parametrize("path_name", ["exact", "stream", "umma"])
def test_fingerprint_l2_ties(path_name):
    """In-tree workload: 0/1 Morgan bits, d=1024, L2, k=20 (retrieve_faiss.py:36-44, :65, :70)."""
    trx = _engine()
    path = {"exact": trx.PATH_EXACT, "stream": trx.PATH_STREAM, "umma": trx.PATH_UMMA}[path_name]
    xb = util.fingerprints(30000, 1024, 31)
    xq = xb[:40]                                   # train->train self retrieval (:114-115)
    D, I, st = _run(xb, xq, 20, L2, path)
    assert (I[:, 0] >= 0).all() and (D[:, 0] == 0).all()
    Do, Io = oracle.search_seq(xb, xq, 20, L2)
    np.testing.assert_array_equal(D, Do)           # integer distances: bit exact
    np.testing.assert_array_equal(I, Io)           # ties broken by ascending id on both sides


def test_int64_count_fingerprints_l2():
    trx = _engine()
    xb = util.count_fingerprints(20000, 2048, 41)
    xq = util.count_fingerprints(50, 2048, 42)
    D, I, st = _run(xb, xq, 20, L2, trx.PATH_AUTO)
    Do, Io = oracle.search_seq(xb, xq, 20, L2)
    np.testing.assert_array_equal(D, Do)
    np.testing.assert_array_equal(I, Io)


@pytest.mark.parametrize("path_name", ["exact", "umma"])
def test_exclusion_mask_equals_post_filter(path_name):
    """masked_search(k) == post_filter(search(k + g)) -- textreact/dataset.py:74-76 semantics."""
    trx = _engine()
    path = {"exact": trx.PATH_EXACT, "umma": trx.PATH_UMMA}[path_name]
    n, d, nq, k, g = 40000, 768, 64, 20, 5
    xb, xq = util.clustered_unit(n, d, 51), util.clustered_unit(nq, d, 52)
    groups = (np.arange(n) // g).astype(np.int32)
    rng = np.random.default_rng(53)
    excl = groups[rng.integers(0, n, nq)].copy()
    excl[::7] = -1
    D, I, _ = _run(xb, xq, k, IP, path, groups=groups, exclude=excl)
    D2, I2, _ = _run(xb, xq, k + g, IP, path)
    for i in range(nq):
        keep = [j for j in range(k + g) if groups[I2[i, j]] != excl[i]][:k]
        np.testing.assert_array_equal(I[i], I2[i, keep])
        np.testing.assert_array_equal(D[i], D2[i, keep])
    oracle.check_parity(D, I, xb, xq, k, IP, groups, excl)


def test_edge_cases():
    trx = _engine()
    d = 32
    xb = util.gaussian(50, d, 61)
    idx = trx.IndexFlatIP(d)
    D, I = idx.search(util.gaussian(3, d, 62), 5)          # empty index
    assert (I == -1).all() and (D == -oracle.FLT_MAX).all()
    idx.add(xb[:20]); idx.add(xb[20:])                     # incremental add
    assert idx.ntotal == 50
    xq = util.gaussian(4, d, 63)
    D, I = idx.search(xq, 64)                              # k > ntotal -> -1 padding
    oracle.check_parity(D, I, xb, xq, 64, IP)
    D, I = idx.search(xq[:1], 1)                           # nq = 1, k = 1
    oracle.check_parity(D, I, xb, xq[:1], 1, IP)
    with pytest.raises(AssertionError):
        idx.search(util.gaussian(2, d + 1, 64), 3)         # wrong dimension (FAISS: AssertionError)
    np.testing.assert_array_equal(idx.reconstruct_n(0, 50), xb)          # FAISS reconstruct: the stored fp32 rows
    np.testing.assert_array_equal(idx.reconstruct(7), xb[7])
    idx.reset()
    assert idx.ntotal == 0
    idx.close()


def test_duplicate_rows_tie_order():
    trx = _engine()
    base = util.gaussian(500, 64, 71)
    xb = np.concatenate([base, base, base])                # every row three times
    xq = base[:9]
    for path in (trx.PATH_EXACT,):
        D, I, _ = _run(xb, xq, 6, IP, path)
        Do, Io = oracle.search_seq(xb, xq, 6, IP)
        np.testing.assert_array_equal(I, Io)


def test_torch_cuda_tensors_roundtrip():
    import torch
    trx = _engine()
    xb, xq = util.gaussian(20000, 256, 81), util.gaussian(130, 256, 82)
    idx = trx.IndexFlatIP(256)
    idx.add(torch.from_numpy(xb).cuda())
    D, I = idx.search(torch.from_numpy(xq).cuda(), 10)
    assert D.is_cuda and I.is_cuda and I.dtype == torch.int64
    oracle.check_parity(D.cpu().numpy(), I.cpu().numpy(), xb, xq, 10, IP)
    idx.close()


def test_merge_topk_equals_unsharded():
    import torch
    trx = _engine()
    n, d, nq, k, G = 48000, 128, 50, 10, 4
    xb, xq = util.gaussian(n, d, 91), util.gaussian(nq, d, 92)
    Ds, Is = [], []
    for g in range(G):
        lo, hi = g * n // G, (g + 1) * n // G
        idx = trx.IndexFlatIP(d)
        idx.add(xb[lo:hi]); idx.set_id_offset(lo)
        Dg, Ig = idx.search(torch.from_numpy(xq).cuda(), k)
        Ds.append(Dg); Is.append(Ig); idx.close()
    D, I = trx.merge_topk(torch.stack(Ds), torch.stack(Is), IP)
    oracle.check_parity(D.cpu().numpy(), I.cpu().numpy(), xb, xq, k, IP)


@pytest.mark.parametrize("jitter", [0.0, 1e-3])
def test_dense_ties_at_the_top_come_back_exact(jitter):
    """6000 (near-)identical rows that are the best match of some queries.
    jitter == 0: exact ties -- the sampled threshold lands ON the tie score, nothing beats it, the certificate
    fails and the queries take the generic fp32 scan + radix select, ties in ascending id order.
    jitter > 0: more rows above the threshold than a candidate list holds (4096) -> overflow -> same exact scan.
    The other queries stay on the prefilter path."""
    trx = _engine()
    n, d, nq, k = 60000, 256, 130, 20
    xb, xq = util.gaussian(n, d, 101), util.gaussian(nq, d, 102)
    hot = xq[:8].sum(0)
    hot *= 3.0 / np.linalg.norm(hot) * np.sqrt(d)
    xb[1000:7000] = hot + jitter * util.gaussian(6000, d, 103)
    D, I, st = _run(xb, xq, k, IP, trx.PATH_UMMA)
    oracle.check_parity(D, I, xb, xq, k, IP)
    assert st["queries_exact"] >= 8, st
    assert ((I[:8] >= 1000) & (I[:8] < 7000)).all()
    if jitter == 0.0:
        assert (I[:8] == np.arange(1000, 1000 + k)).all()
    else:
        assert st["queries_overflow"] >= 1, st


@pytest.mark.parametrize("path_name", ["stream", "umma"])
def test_ascending_score_corpus(path_name):
    """Dist A of SURVEY 8d: row norms grow with the row id, so the best rows all sit at the end of the
    scan -- the worst case for a running-threshold design; the sampled threshold must not care."""
    trx = _engine()
    path = {"stream": trx.PATH_STREAM, "umma": trx.PATH_UMMA}[path_name]
    n, d, k = 80000, 768, 100
    nq = 5 if path_name == "stream" else 200
    xb = util.gaussian(n, d, 111) * (0.25 + 1.5 * np.arange(n, dtype=np.float32)[:, None] / n)
    xq = util.gaussian(nq, d, 112)
    D, I, st = _run(xb, xq, k, IP, path)
    oracle.check_parity(D, I, xb, xq, k, IP)
    assert st["last_path"] == path
    assert np.median(I) > 0.6 * n


@pytest.mark.parametrize("metric", [IP, L2])
@pytest.mark.parametrize("d", [100, 333, 1000])
def test_dimensions_that_are_not_tile_multiples(metric, d):
    """d not a multiple of 64 (bf16 rows are zero padded to the TMA tile) nor of 4 (fp32 rows are not
    16-byte aligned: scalar-load variants of the exact scorer and of the rescore)."""
    trx = _engine()
    n, nq, k = 30000, 150, 10
    xb, xq = util.gaussian(n, d, 121), util.gaussian(nq, d, 122)
    for path in (trx.PATH_UMMA, trx.PATH_STREAM, trx.PATH_EXACT):
        q = xq if path == trx.PATH_UMMA else xq[:5]
        D, I, st = _run(xb, q, k, metric, path)
        oracle.check_parity(D, I, xb, q, k, metric)
        assert st["last_path"] == path


@pytest.mark.parametrize("k", [1, 100, 256, 300, 512, 1000])
def test_k_range(k):
    """k up to 512 stays on the prefilter path (candidate lists grow with k); beyond that the exact scan."""
    trx = _engine()
    n, d, nq = 40000, 128, 140
    xb, xq = util.gaussian(n, d, 131), util.gaussian(nq, d, 132)
    D, I, st = _run(xb, xq, k, IP, trx.PATH_AUTO)
    oracle.check_parity(D, I, xb, xq, k, IP)
    assert st["last_path"] == (trx.PATH_UMMA if k <= 512 else trx.PATH_EXACT)


def test_more_queries_than_one_batch():
    """nq > max_batch: the call is cut into batches; every row of the output belongs to its query."""
    trx = _engine()
    n, d, nq, k = 30000, 64, 2500, 10
    xb, xq = util.gaussian(n, d, 141), util.gaussian(nq, d, 142)
    D, I, st = _run(xb, xq, k, IP, trx.PATH_AUTO, max_batch=1024)
    Do, Io = oracle.search_blas(xb, xq, k, IP)
    assert (I == Io).mean() > 0.999
    oracle.check_parity(D[::50], I[::50], xb, xq[::50], k, IP)
    oracle.check_parity(D[-3:], I[-3:], xb, xq[-3:], k, IP)


def test_pipelined_batches_match_serial_batches():
    """The double-buffered batch pipeline (upload / launch of batch i+1 overlapping batch i) returns exactly what
    the serial loop returns, for pageable host arrays, pinned host arrays and device tensors, with a mask."""
    import torch
    trx = _engine()
    n, d, nq, k = 30000, 64, 3000, 10
    xb, xq = util.gaussian(n, d, 151), util.gaussian(nq, d, 152)
    groups = (np.arange(n) // 3).astype(np.int32)
    excl = groups[np.random.default_rng(153).integers(0, n, nq)].astype(np.int32)
    idx = trx.IndexFlatIP(d)
    idx.add(xb)
    idx.set_groups(groups)
    idx.set_option("max_batch", 512)
    idx.set_option("target_candidates", 64)        # starve some lists: fallbacks inside the pipeline too
    idx.set_option("pipeline", 0)
    D0, I0 = idx.search(xq, k, exclude=excl)
    assert idx.stats()["queries_exact"] + idx.stats()["queries_second_pass"] > 0
    idx.set_option("pipeline", 1)
    D1, I1 = idx.search(xq, k, exclude=excl)                                    # pageable in / out
    np.testing.assert_array_equal(I1, I0); np.testing.assert_array_equal(D1, D0)
    hq = torch.from_numpy(xq).pin_memory()
    hD = torch.empty((nq, k), dtype=torch.float32).pin_memory()
    hI = torch.empty((nq, k), dtype=torch.int64).pin_memory()
    idx.search(hq.numpy(), k, exclude=excl, D=hD.numpy(), I=hI.numpy())         # pinned in / out
    np.testing.assert_array_equal(hI.numpy(), I0); np.testing.assert_array_equal(hD.numpy(), D0)
    Dt, It = idx.search(torch.from_numpy(xq).cuda(), k, exclude=torch.from_numpy(excl).cuda())   # device in / out
    np.testing.assert_array_equal(It.cpu().numpy(), I0); np.testing.assert_array_equal(Dt.cpu().numpy(), D0)
    oracle.check_parity(D0[::40], I0[::40], xb, xq[::40], k, IP, groups, excl[::40])
    idx.close()


@pytest.mark.parametrize("metric", [IP, L2])
def test_search_self_equals_search_of_the_same_rows(metric):
    """train->train mode (retrieve_faiss.py:114-115): queries are rows already in the index."""
    trx = _engine()
    xb = util.fingerprints(20000, 1024, 161) if metric == L2 else util.gaussian(20000, 256, 162)
    idx = trx.IndexFlat(xb.shape[1], metric)
    idx.add(xb)
    D0, I0 = idx.search(xb[3000:3700], 20)
    D1, I1 = idx.search_self(20, 3000, 3700)
    np.testing.assert_array_equal(I1, I0); np.testing.assert_array_equal(D1, D0)
    Dt, It = idx.search_self(20, 3000, 3700, device=True)
    np.testing.assert_array_equal(It.cpu().numpy(), I0)
    assert (I1[:, 0] == np.arange(3000, 3700)).all() or metric == L2     # L2 bits: duplicates may tie at distance 0
    with pytest.raises(AssertionError):
        idx.search_self(5, 10, 30000)
    idx.close()


@pytest.mark.parametrize("frac", [0.9, 0.5, 0.2, 0.04])
@pytest.mark.parametrize("path_name", ["auto", "exact"])
def test_row_attribute_filter_equals_filtered_corpus(frac, path_name):
    """`--before T` (retrieve_faiss.py:102-103) as a per-row attribute: searching the full index with
    attr_below=T equals searching an index built from the rows with year < T."""
    trx = _engine()
    n, d, nq, k = 50000, 128, 150, 20
    xb, xq = util.gaussian(n, d, 171), util.gaussian(nq, d, 172)
    years = np.random.default_rng(173).integers(1976, 2017, n).astype(np.int32)
    T = int(np.quantile(years, frac))
    keep = np.nonzero(years < T)[0]
    idx = trx.IndexFlatIP(d)
    idx.add(xb)
    idx.set_row_attr(years)
    if path_name == "exact":
        idx.set_option("path", trx.PATH_EXACT)
    D, I = idx.search(xq, k, attr_below=T)
    assert (years[I] < T).all()
    Do, Io = oracle.search_blas(xb[keep], xq, k, IP)
    assert (I == keep[Io]).mean() > 0.999
    oracle.check_parity(D[:40], np.searchsorted(keep, I[:40]), xb[keep], xq[:40], k, IP)
    D2, I2 = idx.search(xq, k)                      # the filter is per call: off again
    oracle.check_parity(D2[:20], I2[:20], xb, xq[:20], k, IP)
    idx.close()


@pytest.mark.parametrize("dtype", ["int8", "uint8", "bool", "int16", "int32", "int64", "float16", "float64"])
def test_typed_add_equals_host_side_float32_coercion(dtype):
    """index.add ships raw int8 / int64 / ... arrays and widens them on the device (trx_add_typed); the result
    must be what FAISS's wrapper would have stored: np.ascontiguousarray(x, dtype='float32')."""
    trx = _engine()
    rng = np.random.default_rng(181)
    n, d = 9000, 96
    if dtype == "bool":
        xb = rng.random((n, d)) < 0.1
    elif dtype.startswith("float"):
        xb = rng.standard_normal((n, d)).astype(dtype)
    else:
        lo = 0 if dtype == "uint8" else -3
        xb = (rng.integers(lo, 4, (n, d)) * (rng.random((n, d)) < 0.2)).astype(dtype)
    xq = np.ascontiguousarray(xb[:37], dtype=np.float32)
    res = []
    for arr in (xb, np.ascontiguousarray(xb, dtype=np.float32), np.asfortranarray(xb)):   # typed, fp32, non-contiguous
        idx = trx.IndexFlatL2(d)
        idx.add(arr[:5000]); idx.add(arr[5000:])
        assert idx.ntotal == n
        res.append(idx.search(xq, 10))
        idx.close()
    for D, I in res[1:]:
        np.testing.assert_array_equal(I, res[0][1]); np.testing.assert_array_equal(D, res[0][0])
    Do, Io = oracle.search_seq(xb, xq, 10, L2)
    np.testing.assert_array_equal(res[0][1], Io)
    np.testing.assert_allclose(res[0][0], Do, rtol=1e-5, atol=1e-5)


def test_certificate_bound_holds_for_every_pair():
    """The exactness certificate rests on |bf16-pipeline score - exact score| <= |q^||x^-x| + |x||q^-q| (+ fp32
    accumulation slack).  Check the inequality on every (query, row) pair of a sample, with the measured
    maxima the engine uses, against float64 scores -- and that adversarial inputs (all mantissas at the
    rounding midpoint, same sign) come close to it without crossing it."""
    import torch
    trx = _engine()
    n, d, nq = 20000, 768, 130
    rng = np.random.default_rng(191)
    xb, xq = util.gaussian(n, d, 192), util.gaussian(nq, d, 193)
    # adversarial block: values just below a bf16 rounding midpoint, all positive -> errors add up coherently
    adv = (1.0 + (2.0 ** -8) * 0.999) * np.exp2(rng.integers(-2, 3, (512, d))).astype(np.float32)
    xb[:512] = adv
    xq[:8] = (1.0 + (2.0 ** -8) * 0.999) * np.exp2(rng.integers(-2, 3, (8, d))).astype(np.float32)
    idx = trx.IndexFlatIP(d)
    idx.add(xb)
    got = idx.debug_scores_umma(torch.from_numpy(xq).cuda(), 0, n).cpu().numpy().astype(np.float64)
    idx.close()
    exact = xq.astype(np.float64) @ xb.astype(np.float64).T
    xh = torch.from_numpy(xb).bfloat16().float().numpy().astype(np.float64)
    qh = torch.from_numpy(xq).bfloat16().float().numpy().astype(np.float64)
    ex_max = np.linalg.norm(xh - xb, axis=1).max()
    x_max = np.linalg.norm(xb.astype(np.float64), axis=1).max()
    bound = (np.linalg.norm(qh, axis=1) * ex_max + x_max * np.linalg.norm(qh - xq, axis=1)
             + 2 * (d + 3) * 2.0 ** -23 * np.linalg.norm(xq.astype(np.float64), axis=1) * x_max)
    err = np.abs(got - exact)
    assert (err <= bound[:, None]).all(), float((err / bound[:, None]).max())
    # the adversarial pairs use a good part of the bound: the bound is not vacuous
    assert (err[:8, :512] / bound[:8, None]).max() > 0.2


def test_degenerate_queries_do_not_break_the_batch():
    """A zero query (every score ties at 0) and a NaN query in the same batch as ordinary ones: the ordinary
    queries keep their exact answers, the zero query returns the k lowest ids, nothing hangs or crashes."""
    trx = _engine()
    n, d, nq, k = 30000, 128, 140, 10
    xb, xq = util.gaussian(n, d, 201), util.gaussian(nq, d, 202)
    xq[5] = 0.0
    xq[9, 3] = np.nan
    for path in (trx.PATH_UMMA, trx.PATH_EXACT):
        D, I, st = _run(xb, xq, k, IP, path)
        good = np.array([i for i in range(nq) if i not in (5, 9)])
        oracle.check_parity(D[good], I[good], xb, xq[good], k, IP)
        assert (I[5] == np.arange(k)).all() and (D[5] == 0).all()


def test_wide_rows():
    """d = 4096 (wider than any fingerprint the reference builds): every path still works."""
    trx = _engine()
    n, d, k = 12000, 4096, 10
    xb, xq = util.gaussian(n, d, 211), util.gaussian(40, d, 212)
    for path, nq in ((trx.PATH_UMMA, 40), (trx.PATH_STREAM, 5), (trx.PATH_EXACT, 6)):
        D, I, st = _run(xb, xq[:nq], k, L2, path)
        oracle.check_parity(D, I, xb, xq[:nq], k, L2)
        assert st["last_path"] == path


def _dedup_post_filter(I, D, groups, k):
    """textreact/dataset.py:46-56 deduplicate_neighbors (keep the first row of every text group), then head k."""
    outI = np.full((I.shape[0], k), -1, np.int64)
    outD = np.zeros((I.shape[0], k), np.float32)
    for i in range(I.shape[0]):
        seen, j = set(), 0
        for idn, dn in zip(I[i], D[i]):
            if idn < 0 or groups[idn] in seen:
                continue
            seen.add(groups[idn])
            outI[i, j], outD[i, j] = idn, dn
            j += 1
            if j == k:
                break
        assert j == k
    return outD, outI


@pytest.mark.parametrize("path_name", ["umma", "stream", "exact", "starved"])
@pytest.mark.parametrize("metric", [IP, L2])
def test_distinct_groups_mode_equals_dedup_post_filter(path_name, metric):
    """dedup=True returns k rows of k different groups: identical to searching deeper (k x largest group) and
    running the consumer's deduplicate_neighbors -- with the gold-removed mask on top."""
    trx = _engine()
    n, d, k, gsz = 40000, 128, 20, 4
    nq = 5 if path_name == "stream" else 150
    # near-duplicate rows inside a group (same paragraph, slightly different reaction): they crowd the top
    base = util.clustered_unit(n // gsz, d, 221)
    xb = (np.repeat(base, gsz, axis=0) + 0.02 * util.gaussian(n, d, 222)).astype(np.float32)
    groups = (np.arange(n) // gsz).astype(np.int32)
    xq = util.clustered_unit(nq, d, 223)
    excl = groups[np.random.default_rng(224).integers(0, n, nq)].astype(np.int32)
    idx = trx.IndexFlat(d, metric)
    idx.add(xb)
    idx.set_groups(groups)
    path = {"umma": trx.PATH_UMMA, "stream": trx.PATH_STREAM, "exact": trx.PATH_EXACT, "starved": trx.PATH_UMMA}[path_name]
    idx.set_option("path", path)
    if path_name == "starved":
        idx.set_option("target_candidates", 32)         # too few candidates: the fallbacks must dedup as well
        idx.set_option("thr_bias", 0.03 if metric == IP else 0.06)
    D, I = idx.search(xq, k, exclude=excl, dedup=True)
    Dd, Id = idx.search(xq, k * gsz, exclude=excl)       # deeper, not deduplicated (dedup is per call: off again)
    Dr, Ir = _dedup_post_filter(Id, Dd, groups, k)
    np.testing.assert_array_equal(I, Ir)
    np.testing.assert_allclose(D, Dr, rtol=2e-6, atol=1e-6)   # K4 and the exact scan sum in different orders
    g = groups[I]
    assert all(len(set(row.tolist())) == k for row in g) and not (g == excl[:, None]).any()
    if path_name == "starved":
        assert idx.stats()["queries_exact"] + idx.stats()["queries_second_pass"] > 0
    idx.close()


def _dedup_reference(xb, xq, groups, k, metric, excl=None):
    """Group leaders in (score, id) order from float64 scores: the consumer's deduplicate_neighbors on the FULL ranking."""
    s = oracle.scores_f64(xb, xq, metric)
    key = -s if metric == IP else s
    out = np.full((xq.shape[0], k), -1, np.int64)
    ids = np.arange(xb.shape[0])
    for i in range(xq.shape[0]):
        seen, j = set(), 0
        for r in np.lexsort((ids, key[i])):
            g = groups[r]
            if g in seen or (excl is not None and g == excl[i]):
                continue
            seen.add(g)
            out[i, j] = r
            j += 1
            if j == k:
                break
    return out


@pytest.mark.parametrize("case", ["exact_small_corpus", "prefilter_fallback", "forced_exact"])
@pytest.mark.parametrize("metric", [IP, L2])
def test_distinct_groups_with_groups_larger_than_the_widest_selection(case, metric):
    """real corpora have hundreds of rows sharing one text.  k x (largest group) > 2048 used to make the
    exact path (and through it every fallback) fail the whole call; it now scans in rounds.  One group of 1500
    near-duplicates sits right at the top of every query, a second of 700 behind it; integer-valued data so that
    the expected ids are exact."""
    trx = _engine()
    rng = np.random.default_rng(5)
    n = 6000 if case == "exact_small_corpus" else 30000
    d, k, nq = 64, 20, 12
    xb = rng.integers(-3, 4, (n, d)).astype(np.float32)
    proto = rng.integers(-3, 4, d).astype(np.float32)
    groups = (np.arange(n) + 10).astype(np.int32)                # singletons ...
    big = rng.choice(n, 1500, replace=False)
    rest = np.setdiff1d(np.arange(n), big)
    mid = rng.choice(rest, 700, replace=False)
    xb[big] = proto; xb[big, rng.integers(0, d, 1500)] += rng.integers(-1, 2, 1500)    # ... but two large text groups
    xb[mid] = proto; xb[mid, :2] += 1; xb[mid, rng.integers(2, d, 700)] += rng.integers(-1, 2, 700)
    groups[big], groups[mid] = 1, 2
    xq = (proto[None, :] + rng.integers(-1, 2, (nq, d))).astype(np.float32)
    excl = np.full(nq, -1, np.int32); excl[::3] = 2
    idx = trx.IndexFlat(d, metric)
    idx.add(xb)
    idx.set_groups(groups)
    if case == "forced_exact":
        idx.set_option("path", trx.PATH_EXACT)
    elif case == "prefilter_fallback":
        idx.set_option("path", trx.PATH_UMMA)
    D, I = idx.search(xq, k, exclude=excl, dedup=True)
    st = idx.stats()
    want = _dedup_reference(xb, xq, groups, k, metric, excl)
    np.testing.assert_array_equal(I, want)
    assert st["queries_exact"] > 0                                # the exact scan (direct, or as the fallback) did run
    g = groups[I]
    assert all(len(set(row.tolist())) == k for row in g)
    idx.close()


def test_modes_without_their_side_data_are_errors():
    trx = _engine()
    xb = util.gaussian(2000, 32, 231)
    idx = trx.IndexFlatIP(32)
    idx.add(xb)
    for kw in ({"dedup": True}, {"attr_below": 5}, {"exclude": np.zeros(3, np.int32)}):
        with pytest.raises(RuntimeError, match="trx_set_groups|trx_set_row_attr"):
            idx.search(xb[:3], 4, **kw)
    D, I = idx.search(xb[:3], 4)                 # the failed calls left no option behind
    assert (I[:, 0] == np.arange(3)).all()
    idx.close()


def test_cuda_graph_replay_matches_plain_launches():
    """Small batches replay their kernel pipeline as one cached CUDA graph: same answers as plain launches, for
    changing query contents, after the index grows, with the mask on and off, on both prefilter paths."""
    trx = _engine()
    n, d, k = 30000, 128, 10
    xb = util.gaussian(n + 5000, d, 241)
    groups = (np.arange(n + 5000) // 3).astype(np.int32)
    idx = trx.IndexFlatIP(d)
    idx.add(xb[:n])
    idx.set_groups(groups[:n])
    assert idx.get_option("graphs") == 1
    for path, nq in ((trx.PATH_UMMA, 40), (trx.PATH_STREAM, 1), (trx.PATH_UMMA, 200)):
        idx.set_option("path", path)
        for rep in range(3):                               # the second and third call replay the cached graph
            xq = util.gaussian(nq, d, 250 + rep)
            excl = groups[np.random.default_rng(rep).integers(0, n, nq)] if rep == 2 else None
            idx.set_option("graphs", 1)
            D1, I1 = idx.search(xq, k, exclude=excl)
            idx.set_option("graphs", 0)
            D0, I0 = idx.search(xq, k, exclude=excl)
            np.testing.assert_array_equal(I1, I0); np.testing.assert_array_equal(D1, D0)
            oracle.check_parity(D1, I1, xb[:n], xq, k, IP, groups[:n] if excl is not None else None, excl)
    # graphs were really captured and replayed (a failed capture silently falls back to plain launches)
    assert idx.get_option("graph_captures") >= 3 and idx.get_option("graph_replays") >= 9
    idx.set_option("graphs", 1)
    idx.set_option("path", trx.PATH_UMMA)
    xq = util.gaussian(40, d, 260)
    idx.search(xq, k)
    idx.search(xq, k)
    c0 = idx.get_option("graph_captures")
    idx.add(xb[n:])                                        # buffers move, ntotal changes: the graph must not be reused
    D, I = idx.search(xq, k)
    oracle.check_parity(D, I, xb, xq, k, IP)
    assert idx.get_option("graph_captures") == c0 + 1 and idx.get_option("graphs") == 1
    st = idx.stats()
    assert st["launches"] > 0
    idx.close()


def test_engine_against_real_faiss_when_it_is_installed():
    """Wherever the real `faiss` imports (not in this image), the engine is compared with it directly:
    same ids wherever FAISS's own fp32 gap is not a rounding-level tie, scores within 1e-5 relative."""
    try:
        import faiss
    except Exception:
        faiss = None
    if faiss is None or str(getattr(faiss, "__version__", "")).startswith("textreact_b200"):
        pytest.skip("faiss is not installed (un-vendored, un-pinned by the reference; no network here)")
    trx = _engine()
    xb, xq = util.gaussian(60000, 768, 251), util.gaussian(300, 768, 252)
    for metric, cls in ((IP, faiss.IndexFlatIP), (L2, faiss.IndexFlatL2)):
        ref = cls(768)
        ref.add(xb)
        Df, If = ref.search(xq, 100)
        D, I, _ = _run(xb, xq, 100, metric, trx.PATH_AUTO)
        oracle.check_parity(D, I, xb, xq, 100, metric)
        assert (I == If).mean() > 0.995
        np.testing.assert_allclose(D, Df, rtol=2e-5, atol=2e-3)


@pytest.mark.parametrize("metric", [IP, L2])
@pytest.mark.parametrize("shards,with_mask", [(2, False), (4, True), (8, False)])
def test_two_phase_search_of_row_shards_on_one_gpu(metric, shards, with_mask):
    """The two-phase local search of the row-sharded mode (trx_search_begin -> bounds exchange -> trx_search_finish),
    with all the shards on ONE device so that it runs on a one-GPU box: every shard exports its best prefilter scores,
    the floor computed from them lets each shard rescore only what can reach the GLOBAL top-k, and the k-way merge of
    the (short, padded) per-shard results must be the exact unsharded answer.  Also: far fewer rows are rescored."""
    import torch
    trx = _engine()
    from textreact_b200.index import merge_topk
    from textreact_b200.sharded import bounds_width, floor_from_payloads, shard_bounds
    n, d, nq, k = 120_000, 256, 300, 50
    xb, xq = util.gaussian(n, d, 801), util.gaussian(nq, d, 802)
    groups = (np.arange(n) // 3).astype(np.int32)
    excl = groups[np.random.default_rng(803).integers(0, n, nq)].astype(np.int32) if with_mask else None
    nb = bounds_width(k, shards)
    assert nb * shards >= k
    idx, rescored = [], {}
    for g in range(shards):
        lo, hi = shard_bounds(n, shards, g)
        ix = trx.IndexFlat(d, metric)
        ix.add(xb[lo:hi])
        ix.set_id_offset(lo)
        if with_mask:
            ix.set_groups(groups[lo:hi])
        idx.append(ix)
    xq_t = torch.from_numpy(xq).cuda()
    ex_t = None if excl is None else torch.from_numpy(excl).cuda()
    for mode in ("two_phase", "plain"):
        r0 = sum(ix.stats()["rescored"] for ix in idx)
        if mode == "two_phase":
            payloads = torch.stack([ix.search_begin(xq_t, k, nb, exclude=ex_t) for ix in idx])
            assert payloads.shape == (shards, nq, nb + 1)
            floor = floor_from_payloads(payloads, nb, k)
            res = [ix.search_finish(floor) for ix in idx]
            padded = sum(int((I < 0).sum().item()) for _, I in res)
            assert padded > (1.0 - 1.7 / shards) * shards * nq * k     # most of a shard's local top-k is never produced
        else:
            res = [ix.search(xq_t, k, exclude=ex_t) for ix in idx]
        rescored[mode] = sum(ix.stats()["rescored"] for ix in idx) - r0
        Dm, Im = merge_topk(torch.stack([r[0] for r in res]), torch.stack([r[1] for r in res]), metric)
        oracle.check_parity(Dm.cpu().numpy(), Im.cpu().numpy(), xb, xq, k, metric, groups if with_mask else None, excl)
    assert rescored["two_phase"] < (0.9 if shards == 2 else 0.6) * rescored["plain"], rescored
    # a begin must be finished before anything else is searched; finish(None) is the plain local top-k
    idx[0].search_begin(xq_t, k, nb)
    with pytest.raises(RuntimeError, match="two-phase"):
        idx[0].search(xq_t, k)
    D0, I0 = idx[0].search_finish(None)
    D1, I1 = idx[0].search(xq_t, k)
    np.testing.assert_array_equal(I0.cpu().numpy(), I1.cpu().numpy())
    # a mutation between the halves abandons the pending search (its candidate lists describe the old rows)
    idx[0].search_begin(xq_t, k, nb)
    idx[0].add(xb[:10])
    with pytest.raises(RuntimeError, match="without search_begin"):
        idx[0].search_finish(None)
    for ix in idx:
        ix.close()


def test_add_npy_streams_the_reference_fingerprint_cache(tmp_path):
    """retrieve_faiss.py:106-110 caches the train fingerprints with np.save into `train_fp.pkl`; add_npy ingests such a
    file chunk by chunk (memory-mapped, raw int8 / int64 over PCIe) and gives the same index as add(np.load(...))."""
    trx = _engine()
    for fps in (util.fingerprints(7000, 256, 901), util.count_fingerprints(5000, 128, 902), util.gaussian(6000, 96, 903)):
        path = tmp_path / "train_fp.pkl"
        with open(path, "wb") as f:
            np.save(f, fps)
        a, b = trx.IndexFlatL2(fps.shape[1]), trx.IndexFlatL2(fps.shape[1])
        assert a.add_npy(str(path), chunk_rows=1500) == len(fps) and a.ntotal == len(fps)
        b.add(fps)
        Da, Ia = a.search(fps[:40], 7)
        Db, Ib = b.search(fps[:40], 7)
        np.testing.assert_array_equal(Ia, Ib)
        np.testing.assert_array_equal(Da, Db)
        np.testing.assert_array_equal(a.reconstruct_n(0, 50), np.ascontiguousarray(fps[:50], dtype=np.float32))
        a.close(); b.close()
