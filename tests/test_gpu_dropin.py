"""GPU: `import faiss` resolves to the shim and the reference's call sequence
(retrieve/retrieve_faiss.py:62-74, :114-130) runs unchanged, down to the JSON the predictor loads."""
import json
import os
import sys

import numpy as np
import pytest

from oracle import cpu_flat as oracle
from tests import util

pytestmark = pytest.mark.gpu
SHIM = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "textreact_b200", "shim")


def index_and_search(faiss, train_fps, query_fps):
    # verbatim call sequence of the reference function (prints and timer dropped)
    d = train_fps.shape[1]
    index = faiss.IndexFlatL2(d)
    index.add(train_fps)
    k = 20
    distance, rank = index.search(query_fps, k)
    return rank


def test_reference_call_sequence_through_the_shim(tmp_path):
    sys.path.insert(0, SHIM)
    try:
        sys.modules.pop("faiss", None)
        import faiss
        assert faiss.__version__.startswith("textreact_b200")
        train_fps = util.fingerprints(12000, 1024, 1)                # int8 Morgan bits (:36-44)
        val_fps = util.fingerprints(64, 1024, 2)
        train_id = [f"US{20000 + i // 4}_{i % 4}" for i in range(len(train_fps))]
        val_id = [f"US{90000 + i}_0" for i in range(len(val_fps))]
        for qfps, qid, name in ((train_fps[:200], train_id[:200], "train.json"), (val_fps, val_id, "val.json")):
            rank = index_and_search(faiss, train_fps, qfps)
            assert rank.dtype == np.int64 and rank.shape == (len(qfps), 20)
            result = [{'id': qid[i], 'nn': [train_id[n] for n in nn]} for i, nn in enumerate(rank)]   # :116
            with open(tmp_path / name, 'w') as f:
                json.dump(result, f)
            Do, Io = oracle.search_seq(train_fps, qfps, 20, 1)
            np.testing.assert_array_equal(rank, Io)
        nn = {ex['id']: ex['nn'] for ex in json.load(open(tmp_path / "train.json"))}          # dataset.py:40-44
        assert all(q in nn and len(nn[q]) == 20 for q in train_id[:200])
    finally:
        sys.path.remove(SHIM)
        sys.modules.pop("faiss", None)


def test_difference_fingerprint_int64_inputs():
    sys.path.insert(0, SHIM)
    try:
        sys.modules.pop("faiss", None)
        import faiss
        train = util.count_fingerprints(9000, 2048, 3)               # int64 counts (:18-27)
        rank = index_and_search(faiss, train, train[:50])
        Do, Io = oracle.search_seq(train, train[:50], 20, 1)
        np.testing.assert_array_equal(rank, Io)
    finally:
        sys.path.remove(SHIM)
        sys.modules.pop("faiss", None)
