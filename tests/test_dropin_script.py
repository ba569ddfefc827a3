"""CPU: the reference's retrieval script, UNCHANGED (retrieve/retrieve_faiss.py:77-130 with the argv of
retrieve/condition_year.sh:1-7 and retrieve/retro_year.sh:9-16), runs here against synthetic CSVs and reproduces
the committed record tests/golden/dropin/golden.json (made by tests/golden/make_dropin_golden.py).  `rdkit` is a
stub (featurisers are out of scope) and `import faiss` is the oracle-backed stand-in: there is no GPU in this
container.  tests/test_gpu_dropin.py runs the same thing against the B200 engine.

/root/reference does not exist on the GPU box: the runs are skipped there, the well-formedness checks are not."""
import json
import os

import pytest

from tests.dropin import world

SCRIPT = os.environ.get("TRX_REFERENCE_SCRIPT", "/root/reference/retrieve/retrieve_faiss.py")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dropin", "golden.json")


def golden():
    with open(GOLDEN) as f:
        return json.load(f)


def test_golden_record_is_well_formed():
    g = golden()
    assert set(g["scenarios"]) == set(world.SCENARIOS)
    for sc, runs in g["scenarios"].items():
        for run in ("first_run", "cache_run"):
            r = runs[run]
            assert set(r["files"]) == {"train.json", "val.json", "test.json"}
            calls = r["calls"]
            # three fresh indexes (train->train, val, test), each: IndexFlatL2(d); add(train_fps); search(query_fps, 20)
            assert [c["call"] for c in calls] == ["ctor", "add", "search"] * 3
            assert all(c["cls"] == "IndexFlatL2" for c in calls[0::3]) and all(c["k"] == 20 for c in calls[2::3])
            assert calls[1]["shape"] == calls[2]["shape"]                       # self retrieval: nq == N (:114-115)
            assert calls[1]["dtype"] == ("int64" if sc == "condition_year" else "int8")   # (:26, :40)
            for name, f in r["files"].items():
                assert len(f["sha256"]) == 64 and len(f["head"][0]["nn"]) == 20
    # the reference filters `train_df` by year only on a cache miss: with --before, a cache hit maps the same ranks
    # to other ids (recorded, not "fixed")
    ry = g["scenarios"]["retro_year_before_2012"]
    assert ry["first_run"]["files"] != ry["cache_run"]["files"] and ry["first_run"]["calls"] == ry["cache_run"]["calls"]
    cy = g["scenarios"]["condition_year"]
    assert cy["first_run"]["files"] == cy["cache_run"]["files"]


@pytest.mark.parametrize("scenario", sorted(world.SCENARIOS))
def test_unchanged_script_reproduces_the_golden_record(scenario, tmp_path):
    if not os.path.exists(SCRIPT):
        pytest.skip(f"{SCRIPT} is not present (the reference tree does not travel to the GPU box)")
    data, out = str(tmp_path / "data"), str(tmp_path / "out")
    world.write_world(data)
    g = golden()["scenarios"][scenario]
    for run in ("first_run", "cache_run"):          # the second run takes the train_fp.pkl branch (:100-110)
        log = str(tmp_path / f"{scenario}.{run}.jsonl")
        p = world.run_script(SCRIPT, scenario, data, out, world.ORACLE_FAISS, call_log=log)
        assert p.returncode == 0, p.stderr[-3000:]
        od = world.output_dir(scenario, out)
        assert os.path.exists(os.path.join(od, "train_fp.pkl"))
        for name, f in g[run]["files"].items():
            assert world.sha256_file(os.path.join(od, name)) == f["sha256"], (scenario, run, name)
        assert [json.loads(ln) for ln in open(log)] == g[run]["calls"]
