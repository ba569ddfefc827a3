"""The ``{id, nn}`` JSON hand-off between retrieval and the predictor.

Reference anchors:
  * writer: retrieve/retrieve_faiss.py:116-118 (and :122-124, :128-130)
        result = [{'id': query_id[i], 'nn': [train_id[n] for n in nn]} for i, nn in enumerate(rank)]
        json.dump(result, f)
  * Tevatron converter: retrieve/convert_format.py:7-16
  * reader: textreact/dataset.py:40-44   self.neighbors = {ex['id']: ex['nn'] for ex in nn_data}
The reference's writer is an O(nq*k) Python loop over pandas Series (13.5 s per 20K queries at k = 100, i.e.
minutes for the 700K-query job whose search takes 3 s here); ``write_nn_json`` encodes every corpus id once and
assembles rows with one numpy take + one join each, byte-identical to ``json.dump`` of the reference's list.
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


def dumps_nn_json(query_ids, corpus_ids, rank):
    """The text ``json.dumps(result)`` would produce for the reference's ``result`` list
    (retrieve/retrieve_faiss.py:116-118), byte for byte, without building the 100-strings-per-query Python
    objects: every corpus id is JSON-encoded once, rows are assembled with one take + one join each."""
    rank = np.asarray(rank)
    qids = _encode_ids(query_ids)
    assert rank.ndim == 2 and rank.shape[0] == len(qids)
    quoted = np.array(_encode_ids(corpus_ids), dtype=object)
    padded = bool((rank < 0).any())
    parts = []
    for i, qid in enumerate(qids):
        row = rank[i]
        if padded:
            row = row[row >= 0]
        parts.append('{"id": ' + qid + ', "nn": [' + ", ".join(quoted[row].tolist()) + "]}")
    return "[" + ", ".join(parts) + "]"


def _encode_ids(ids):
    """JSON text of every id (what json.dumps(id) returns), strings through the C string encoder directly."""
    lst = ids.tolist() if hasattr(ids, "tolist") else list(ids)     # pandas Series / numpy array / list
    enc = json.encoder.encode_basestring_ascii
    return [enc(c) if type(c) is str else json.dumps(_py(c)) for c in lst]


def write_nn_json(path, query_ids, corpus_ids, rank):
    with open(path, "w") as f:
        f.write(dumps_nn_json(query_ids, corpus_ids, rank))


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
