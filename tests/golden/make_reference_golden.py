"""Regenerates tests/golden/reference_consumer/world.{json,npz} by running the REFERENCE's own consumer code.

Run from the repo root, in the build container only (needs /root/reference):
    python tests/golden/make_reference_golden.py

What is pinned.  The retrieval output is consumed by `BaseDataset.load_corpus` / `get_neighbor_text` /
`deduplicate_neighbors` (reference textreact/dataset.py:40-44, :58-80, :46-56).  Those methods are imported
UNMODIFIED from /root/reference (the module's `rdkit` import is stubbed: none of the three methods touches it) and
run on a synthetic world:

  * a corpus of N rows with ids "US<patent>_<n>"; rows of one patent paragraph share the same text (1..6 rows per
    text); a few rows have no text at all (the reference drops neighbour ids that are not `in corpus`, :60);
  * queries that are perturbed copies of corpus rows and carry that row's id (their "gold"), some without gold;
  * depth-100 neighbour lists from a flat inner-product search (the oracle stands in for FAISS), written with
    textreact_b200.nnfile.write_nn_json and read back by the reference's load_corpus;
  * the reference's selection of `num_neighbors` = 3 texts per query, with skip_gold_neighbor off and on
    (main.py:336-340).

The fixture stores inputs and the reference's outputs.  tests/test_reference_consumer.py then requires
(CPU) the oracle's post_filter restatement and (GPU) the engine's in-engine masks -- exclude= (gold-removed),
dedup=True (distinct texts), attr_below= (`in corpus`) at k = 3 -- to select exactly those texts."""
import json
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_consumer")


def import_reference_dataset():
    for name in ("rdkit", "rdkit.Chem"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["rdkit"].Chem = sys.modules["rdkit.Chem"]
    sys.path.insert(0, "/root/reference")
    import textreact.dataset as ds   # the reference module, unmodified
    return ds


def main():
    from oracle import cpu_flat as oracle
    from textreact_b200 import nnfile
    ds = import_reference_dataset()

    rng = np.random.default_rng(20241017)
    n_text, d, depth, num_neighbors = 900, 48, 100, 3
    sizes = rng.integers(1, 7, n_text)                       # rows per text
    group = np.repeat(np.arange(n_text), sizes).astype(np.int32)
    n = int(group.shape[0])
    cent = rng.standard_normal((40, d)).astype(np.float32)
    base = cent[rng.integers(0, 40, n_text)] + 0.8 * rng.standard_normal((n_text, d)).astype(np.float32)
    xb = (base[group] + 0.15 * rng.standard_normal((n, d)).astype(np.float32)).astype(np.float32)
    corpus_ids = [f"US{20000000 + g}_{j}" for g, s in enumerate(sizes) for j in range(s)]
    texts = [f"Heading {g}. Paragraph describing the preparation of compound {g}." for g in group]
    has_text = rng.random(n) > 0.04                            # a few rows are missing from the corpus file
    corpus = {cid: t for cid, t, h in zip(corpus_ids, texts, has_text) if h}

    nq = 80
    src = rng.choice(n, nq, replace=False)
    xq = (xb[src] + 0.35 * rng.standard_normal((nq, d)).astype(np.float32)).astype(np.float32)
    query_ids = [corpus_ids[s] if i % 5 else f"US99{i:06d}_0" for i, s in enumerate(src)]   # every 5th: no gold

    # keep queries whose ranking is unambiguous in float64 over the depth we use
    D64, I64 = oracle.search_f64(xb, xq, depth, 0, extra=1)
    gaps = np.abs(np.diff(D64, axis=1)) / np.maximum(np.abs(D64[:, :-1]), 1e-30)
    keep = np.nonzero(gaps[:, :40].min(axis=1) > 1e-4)[0]    # the filters consume a short prefix of the list
    xq, query_ids, I64 = xq[keep], [query_ids[i] for i in keep], I64[keep, :depth]
    nq = len(keep)

    with tempfile.TemporaryDirectory() as tmp:
        nn_file = os.path.join(tmp, "test.json")
        nnfile.write_nn_json(nn_file, query_ids, corpus_ids, I64)         # our writer ...
        dset = object.__new__(ds.BaseDataset)                             # ... the reference's reader and filters
        dset.args = types.SimpleNamespace(num_neighbors=num_neighbors, max_num_neighbors=10, use_gold_neighbor=False,
                                          random_neighbor_ratio=0.0)
        dset.indices = list(query_ids)
        dset.split = "test"
        ds.BaseDataset.load_corpus(dset, corpus, nn_file)
        assert dset.neighbors == {q: [corpus_ids[j] for j in row] for q, row in zip(query_ids, I64)}
        out = {}
        for skip in (False, True):
            dset.skip_gold_neighbor = skip
            out["skip_gold" if skip else "plain"] = [ds.BaseDataset.get_neighbor_text(dset, i, return_list=True)
                                                     for i in range(nq)]

    fixture = {
        "about": "outputs of the reference's BaseDataset.load_corpus/get_neighbor_text (textreact/dataset.py:40-80), "
                 "generated by tests/golden/make_reference_golden.py",
        "d": d, "depth": depth, "num_neighbors": num_neighbors,
        "corpus_ids": corpus_ids, "group": group.tolist(), "has_text": has_text.astype(int).tolist(),
        "query_ids": query_ids, "texts_by_group": [f"Heading {g}. Paragraph describing the preparation of compound {g}."
                                                   for g in range(n_text)],
        "reference_plain": out["plain"], "reference_skip_gold": out["skip_gold"],
    }
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "world.json")
    with open(path, "w") as f:
        json.dump(fixture, f)
    np.savez_compressed(os.path.join(OUT, "world.npz"), xb=xb, xq=xq, rank=I64)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB (+ .npz);", nq, "queries,", n, "rows,",
          sum(len(t) < num_neighbors for t in out["skip_gold"]), "short lists")


if __name__ == "__main__":
    main()
