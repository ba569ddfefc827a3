"""CPU: the {id, nn} JSON hand-off (retrieve/retrieve_faiss.py:116-118, retrieve/convert_format.py,
textreact/dataset.py:40-44)."""
import json

import numpy as np

from textreact_b200 import nnfile


def reference_writer(query_id, train_id, rank):
    # the reference's own comprehension (retrieve_faiss.py:116), on plain lists instead of pandas Series
    return [{'id': query_id[i], 'nn': [train_id[n] for n in nn]} for i, nn in enumerate(rank)]


def test_writer_is_byte_identical_to_reference_dump(tmp_path):
    rng = np.random.default_rng(0)
    train_id = [f"US{1000 + i // 3}_{i % 3}" for i in range(500)]
    query_id = [f"US{9000 + i}_0" for i in range(40)]
    rank = rng.integers(0, 500, size=(40, 20)).astype(np.int64)
    p = tmp_path / "test.json"
    nnfile.write_nn_json(p, query_id, train_id, rank)
    assert p.read_text() == json.dumps(reference_writer(query_id, train_id, rank))
    loaded = nnfile.load_nn_json(p)                       # dataset.load_corpus view
    assert list(loaded) == query_id and loaded[query_id[3]] == [train_id[n] for n in rank[3]]


def test_padding_ids_are_dropped():
    recs = nnfile.rank_to_records(["q"], ["a", "b"], np.array([[1, 0, -1, -1]]))
    assert recs == [{"id": "q", "nn": ["b", "a"]}]


def test_convert_tevatron(tmp_path):
    src, dst = tmp_path / "in.jsonl", tmp_path / "out.json"
    rows = [{"query_id": "q1", "query": "CCO", "negative_passages": [{"docid": "d3", "text": "x"}, {"docid": "d1"}]},
            {"query_id": "q2", "query": "CCN", "negative_passages": []}]
    src.write_text("\n".join(json.dumps(r) for r in rows) + "\n")
    nnfile.convert_tevatron(src, dst)
    assert json.loads(dst.read_text()) == [{"id": "q1", "nn": ["d3", "d1"]}, {"id": "q2", "nn": []}]
