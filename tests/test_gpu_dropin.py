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


# ---- the whole script, not only its index calls ------------------------------------------------
from tests.dropin import world  # noqa: E402

GOLDEN_DROPIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dropin", "golden.json")
REF_SCRIPT = os.environ.get("TRX_REFERENCE_SCRIPT", "/root/reference/retrieve/retrieve_faiss.py")


def _script_inputs(scenario, data):
    """What retrieve_faiss.py:89-112 builds on a cache miss: dataframes (year filter applied), fingerprint arrays."""
    import pandas as pd
    sys.path.insert(0, world.STUBS)
    try:
        sys.modules.pop("rdkit", None)
        import rdkit
        argv = [a.format(data=data, out="unused") for a in world.SCENARIOS[scenario]]
        opt = {argv[i][2:]: argv[i + 1] for i in range(0, len(argv), 2)}
        dfs = [pd.read_csv(os.path.join(opt["data_path"], opt[f]), keep_default_na=False)
               for f in ("train_file", "valid_file", "test_file")]
        if "before" in opt:
            dfs[0] = dfs[0][dfs[0]["year"] < int(opt["before"])].reset_index(drop=True)
        if opt["field"] == "canonical_rxn":
            fp = lambda col: np.array([np.array([x for x in rdkit.fake_difference_counts(s)]) for s in col])   # noqa: E731
        else:
            fp = lambda col: np.array([rdkit.fake_morgan_bits(s) for s in col])                                  # noqa: E731
        return dfs, [fp(df[opt["field"]]) for df in dfs]
    finally:
        sys.path.remove(world.STUBS)
        for m in [m for m in sys.modules if m == "rdkit" or m.startswith("rdkit.")]:
            sys.modules.pop(m)


@pytest.mark.parametrize("scenario", sorted(world.SCENARIOS))
def test_script_flow_on_the_engine_reproduces_the_unchanged_scripts_files(scenario, tmp_path):
    """tests/golden/dropin/golden.json records what the UNCHANGED reference script wrote when run in the build
    container (tests/golden/make_dropin_golden.py; `faiss` there = the CPU oracle, no GPU).  Here the same inputs go
    through the same three index_and_search calls on the B200 engine, the id mapping of :116 and json.dump: the
    train/val/test.json files must be byte-identical, through the reference's list comprehension AND through
    textreact_b200.nnfile."""
    from textreact_b200 import nnfile
    with open(GOLDEN_DROPIN) as f:
        gold = json.load(f)["scenarios"][scenario]["first_run"]
    data = str(tmp_path / "data")
    world.write_world(data)
    (train_df, val_df, test_df), (train_fps, val_fps, test_fps) = _script_inputs(scenario, data)
    assert list(train_fps.shape) == gold["calls"][1]["shape"] and str(train_fps.dtype) == gold["calls"][1]["dtype"]
    sys.path.insert(0, SHIM)
    try:
        sys.modules.pop("faiss", None)
        import faiss
        train_id = train_df["id"]
        for name, qfps, qdf in (("train.json", train_fps, train_df), ("val.json", val_fps, val_df),
                                ("test.json", test_fps, test_df)):
            rank = index_and_search(faiss, train_fps, qfps)
            query_id = qdf["id"]
            result = [{'id': query_id[i], 'nn': [train_id[n] for n in nn]} for i, nn in enumerate(rank)]      # :116
            with open(tmp_path / name, "w") as f:
                json.dump(result, f)
            assert world.sha256_file(tmp_path / name) == gold["files"][name]["sha256"], (scenario, name)
            nnfile.write_nn_json(tmp_path / ("fast_" + name), query_id, train_id, rank)
            assert world.sha256_file(tmp_path / ("fast_" + name)) == gold["files"][name]["sha256"]
    finally:
        sys.path.remove(SHIM)
        sys.modules.pop("faiss", None)


@pytest.mark.parametrize("scenario", sorted(world.SCENARIOS))
def test_unchanged_script_runs_on_the_engine_when_the_reference_tree_is_present(scenario, tmp_path):
    """`python retrieve_faiss.py <argv of condition_year.sh / retro_year.sh>` itself, `import faiss` = the shim.
    The reference tree does not travel to the driver's GPU box (skip there); profiles/ holds the log of a run where
    the script file was handed to the box out of band (scripts/run_reference_script_on_gpu.sh)."""
    if not os.path.exists(REF_SCRIPT):
        pytest.skip(f"{REF_SCRIPT} is not present on this box")
    with open(GOLDEN_DROPIN) as f:
        gold = json.load(f)["scenarios"][scenario]
    data, out = str(tmp_path / "data"), str(tmp_path / "out")
    world.write_world(data)
    for run in ("first_run", "cache_run"):
        p = world.run_script(REF_SCRIPT, scenario, data, out, SHIM)
        assert p.returncode == 0, p.stderr[-3000:]
        for name, f in gold[run]["files"].items():
            assert world.sha256_file(os.path.join(world.output_dir(scenario, out), name)) == f["sha256"], (run, name)
