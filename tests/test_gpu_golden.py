"""GPU: the CUDA engine against the committed known-answer vectors (tests/golden), every path."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_engine_matches_golden(path):
    import textreact_b200 as trx
    z = np.load(path)
    xb, xq, k, metric = z["xb"], z["xq"], int(z["k"]), int(z["metric"])
    idx = trx.IndexFlat(xb.shape[1], metric)
    idx.add(xb)
    excl = None
    if "groups" in z.files:
        idx.set_groups(z["groups"])
        excl = z["excl"]
    D, I = idx.search(xq, k, exclude=excl)
    np.testing.assert_array_equal(I, z["I"])
    np.testing.assert_allclose(D, z["D"], rtol=1e-5, atol=1e-5)
    idx.close()


def test_engine_matches_golden_after_replication_on_prefilter_paths():
    """The golden corpora are tiny (exact path).  Tile them with far-away filler rows so the tcgen05 and
    streaming prefilter paths run, and require the same answers (ids of the original rows)."""
    import textreact_b200 as trx
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "kat_ip_gauss.npz"))
    xb, xq, k = z["xb"], z["xq"], int(z["k"])
    rng = np.random.default_rng(99)
    filler = (0.01 * rng.standard_normal((30000, xb.shape[1]))).astype(np.float32)   # scores ~0: never in the top-k
    big = np.concatenate([filler[:15000], xb, filler[15000:]])
    for path in (trx.PATH_UMMA, trx.PATH_STREAM):
        idx = trx.IndexFlatIP(xb.shape[1])
        idx.add(big)
        idx.set_option("path", path)
        D, I = idx.search(xq[:4] if path == trx.PATH_STREAM else xq, k)
        ref = z["I"][:4] if path == trx.PATH_STREAM else z["I"]
        np.testing.assert_array_equal(I - 15000, ref)
        assert idx.stats()["last_path"] == path
        idx.close()
