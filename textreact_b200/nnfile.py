"""The ``{id, nn}`` JSON hand-off between retrieval and the predictor.

Reference anchors:
  * writer: retrieve/retrieve_faiss.py:116-118 (and :122-124, :128-130)
        result = [{'id': query_id[i], 'nn': [train_id[n] for n in nn]} for i, nn in enumerate(rank)]
        json.dump(result, f)
  * Tevatron converter: retrieve/convert_format.py:7-16
  * reader: textreact/dataset.py:40-44   self.neighbors = {ex['id']: ex['nn'] for ex in nn_data}
The reference's writer is an O(nq*k) Python loop over pandas Series; ``write_nn_json`` does the
id mapping with one vectorised numpy take and produces byte-identical ``json.dump`` output.
"""
from __future__ import annotations

import json

import numpy as np


def rank_to_records(query_ids, corpus_ids, rank):
    """[{'id': qid, 'nn': [corpus ids, best first]}]; -1 slots (k > ntotal) are dropped."""
    rank = np.asarray(rank)
    corpus = np.asarray(list(corpus_ids), dtype=object)
    qids = list(query_ids)
    assert rank.ndim == 2 and rank.shape[0] == len(qids)
    mapped = corpus[np.maximum(rank, 0)]
    valid = rank >= 0
    out = []
    for i, qid in enumerate(qids):
        row = mapped[i].tolist() if valid[i].all() else mapped[i][valid[i]].tolist()
        out.append({"id": _py(qid), "nn": [_py(v) for v in row]})
    return out


def _py(v):
    return v.item() if isinstance(v, np.generic) else v


def write_nn_json(path, query_ids, corpus_ids, rank):
    with open(path, "w") as f:
        json.dump(rank_to_records(query_ids, corpus_ids, rank), f)


def convert_tevatron(input_path, output_path):
    """retrieve/convert_format.py:7-16: jsonl {query_id, negative_passages:[{docid}]} -> {id, nn}."""
    output = []
    with open(input_path) as f:
        for line in f:
            data = json.loads(line)
            output.append({"id": data["query_id"], "nn": [p["docid"] for p in data["negative_passages"]]})
    with open(output_path, "w") as f:
        json.dump(output, f)


def load_nn_json(path):
    """textreact/dataset.py:40-44."""
    with open(path) as f:
        return {ex["id"]: ex["nn"] for ex in json.load(f)}
