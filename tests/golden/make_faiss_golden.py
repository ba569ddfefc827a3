"""Known-answer vectors produced by the REAL faiss (faiss-cpu wheel), for the call sites
retrieve/retrieve_faiss.py:65-71 (`IndexFlatL2(d)` / `.add` / `.search`) and the neural retriever's
`IndexFlatIP` (README.md:44-47).

    python tests/golden/make_faiss_golden.py [--out tests/golden/faiss]

Needs `import faiss` to resolve to the real package (it is neither vendored nor pinned by the
reference and is not in the build image; `scripts/try_faiss.sh` records the attempt to obtain it on a
GPU box).  Fixtures hold the inputs' generator + seed (tests/util.py), not the inputs, when those are
large; `faiss.__version__` and the BLAS/heap path taken are recorded beside the answers.

Cases: C1 (100K x 768, 1K queries, k=20: the BLAS + heap path), the same corpus at k=100 (the
reservoir path, k >= 100), nq<20 (the scalar path), and the in-tree fingerprint shapes (0/1 int8 bits
d=1024 and int64 signed counts d=2048 under IndexFlatL2 with k=20, self retrieval)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import util  # noqa: E402

CASES = [
    # name, generator, (n, d, seed), (nq, seed) or "self:<nq>", k, metric
    ("c1_ip_k20", "gaussian", (100_000, 768, 11), (1000, 12), 20, 0),
    ("c1_ip_k100", "gaussian", (100_000, 768, 11), (256, 12), 100, 0),
    ("c1_l2_k20", "gaussian", (100_000, 768, 11), (256, 12), 20, 1),
    ("c1_ip_scalar_nq7", "gaussian", (100_000, 768, 11), (7, 12), 20, 0),
    ("unit_ip_k100", "clustered_unit", (50_000, 768, 21), (200, 22), 100, 0),
    ("bits_l2_k20_self", "fingerprints", (20_000, 1024, 33), "self:300", 20, 1),
    ("counts_l2_k20_self", "count_fingerprints", (10_000, 2048, 34), "self:200", 20, 1),
    ("bits_l2_k100_self", "fingerprints", (20_000, 1024, 33), "self:100", 100, 1),
]


def inputs(case):
    name, gen, (n, d, seed), q, k, metric = case
    xb = getattr(util, gen)(n, d, seed)
    if isinstance(q, str):
        xq = xb[: int(q.split(":")[1])]
    else:
        xq = getattr(util, gen)(q[0], d, q[1])
    return xb, xq, k, metric


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "faiss"))
    a = ap.parse_args()
    import faiss
    ver = str(getattr(faiss, "__version__", "?"))
    if ver.startswith("textreact_b200"):
        raise SystemExit("`import faiss` resolved to the textreact_b200 shim, not the real faiss")
    os.makedirs(a.out, exist_ok=True)
    for case in CASES:
        xb, xq, k, metric = inputs(case)
        # the FAISS python wrapper's coercion: np.ascontiguousarray(x, dtype='float32')
        index = (faiss.IndexFlatIP if metric == 0 else faiss.IndexFlatL2)(xb.shape[1])
        index.add(np.ascontiguousarray(xb, dtype="float32"))
        D, I = index.search(np.ascontiguousarray(xq, dtype="float32"), k)
        np.savez_compressed(os.path.join(a.out, case[0] + ".npz"), D=D, I=I, k=np.int64(k), metric=np.int64(metric),
                            faiss_version=np.array(ver), omp_threads=np.int64(faiss.omp_get_max_threads()),
                            checksum=np.float64(np.asarray(xb, dtype=np.float64).sum() + np.asarray(xq, dtype=np.float64).sum()))
        print(case[0], "faiss", ver, "D", D.shape, "I", I.shape)


if __name__ == "__main__":
    main()
