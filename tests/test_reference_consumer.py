"""The retrieval output as the REFERENCE consumes it.  tests/golden/reference_consumer/ holds what the reference's own
`BaseDataset.load_corpus` / `get_neighbor_text` / `deduplicate_neighbors` (textreact/dataset.py:40-80, imported
unmodified by tests/golden/make_reference_golden.py) selected on a synthetic world: 3 neighbour texts per query,
with skip_gold_neighbor off and on.

  CPU  the oracle's `post_filter` restatement reproduces the reference's selection from the same ranked lists;
       the `{id, nn}` file written by nnfile.write_nn_json is what the reference's reader parsed.
  GPU  the engine, asked for k = 3 with its in-engine masks (attr_below = "in corpus", exclude = gold group,
       dedup = distinct texts), returns rows whose texts ARE the reference's selection -- no depth-100 list needed."""
import json
import os

import numpy as np
import pytest

from oracle import cpu_flat as oracle
from textreact_b200 import nnfile

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_consumer")


@pytest.fixture(scope="module")
def world():
    with open(os.path.join(HERE, "world.json")) as f:
        w = json.load(f)
    z = np.load(os.path.join(HERE, "world.npz"))
    w.update(xb=z["xb"], xq=z["xq"], rank=z["rank"])
    w["group"] = np.asarray(w["group"], np.int32)
    w["has_text"] = np.asarray(w["has_text"], bool)
    w["corpus"] = {cid: w["texts_by_group"][g] for cid, g, h in zip(w["corpus_ids"], w["group"], w["has_text"]) if h}
    return w


def test_oracle_ranking_is_the_fixture_ranking(world):
    D, I = oracle.search_blas(world["xb"], world["xq"], 40, 0)
    np.testing.assert_array_equal(I, world["rank"][:, :40])


def test_post_filter_restatement_matches_the_reference(world):
    corpus, ids = world["corpus"], world["corpus_ids"]
    for qi, qid in enumerate(world["query_ids"]):
        nn = [ids[j] for j in world["rank"][qi]]
        plain = oracle.post_filter(nn, corpus, None, world["num_neighbors"])
        assert [corpus[i] for i in plain] == world["reference_plain"][qi]
        gold = corpus.get(qid)
        skip = oracle.post_filter(nn, corpus, gold, world["num_neighbors"])
        assert [corpus[i] for i in skip] == world["reference_skip_gold"][qi]


def test_nn_file_round_trip(world, tmp_path):
    p = tmp_path / "test.json"
    nnfile.write_nn_json(p, world["query_ids"], world["corpus_ids"], world["rank"])
    nn = nnfile.load_nn_json(p)
    assert list(nn) == world["query_ids"]
    assert nn[world["query_ids"][3]] == [world["corpus_ids"][j] for j in world["rank"][3]]


@pytest.mark.gpu
@pytest.mark.parametrize("path_name", ["exact", "umma"])
def test_engine_masks_select_what_the_reference_selects(world, path_name):
    import textreact_b200 as trx
    xb, xq, k = world["xb"], world["xq"], world["num_neighbors"]
    if path_name == "umma":      # pad with far-away filler rows (no text) so the tcgen05 prefilter path is taken
        filler = (0.01 * np.random.default_rng(5).standard_normal((20000, xb.shape[1]))).astype(np.float32)
        xb = np.concatenate([xb, filler])
    n0 = world["xb"].shape[0]
    group = np.concatenate([world["group"], 10 ** 6 + np.arange(xb.shape[0] - n0, dtype=np.int32)])
    in_corpus = np.concatenate([world["has_text"], np.zeros(xb.shape[0] - n0, bool)])
    idx = trx.IndexFlatIP(xb.shape[1])
    idx.add(xb)
    idx.set_groups(group)
    idx.set_row_attr((~in_corpus).astype(np.int32))            # 0 = has a corpus text, 1 = `i not in corpus`
    idx.set_option("path", trx.PATH_UMMA if path_name == "umma" else trx.PATH_EXACT)
    row_of = {cid: i for i, cid in enumerate(world["corpus_ids"])}
    gold = np.array([world["group"][row_of[q]] if q in world["corpus"] else -1 for q in world["query_ids"]], np.int32)
    texts = world["texts_by_group"]
    D, I = idx.search(xq, k, attr_below=1, dedup=True)
    assert [[texts[group[j]] for j in row] for row in I] == world["reference_plain"]
    D, I = idx.search(xq, k, attr_below=1, dedup=True, exclude=gold)
    assert [[texts[group[j]] for j in row] for row in I] == world["reference_skip_gold"]
    assert idx.stats()["last_path"] == (trx.PATH_UMMA if path_name == "umma" else trx.PATH_EXACT)
    idx.close()
