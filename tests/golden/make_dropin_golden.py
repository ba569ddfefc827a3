"""Runs the reference's retrieval script UNCHANGED (retrieve/retrieve_faiss.py, with the argv of
retrieve/condition_year.sh and retrieve/retro_year.sh) in this container and records what it produces:

    python tests/golden/make_dropin_golden.py        ->  tests/golden/dropin/golden.json

`rdkit` is stubbed (tests/dropin/stubs: deterministic fake fingerprints; featurisers are out of scope) and
`import faiss` resolves to the oracle-backed stand-in (tests/dropin/oracle_faiss: there is no GPU here and no real
faiss anywhere).  Recorded per scenario: sha256 + head of train/val/test.json, the call log of the index object
(constructor, add / search shapes and dtypes, k), and the script's stdout tail.  Each scenario runs twice: the second run
takes the `train_fp.pkl` cache branch (retrieve_faiss.py:100-110) and must produce the same files.
The GPU test (tests/test_gpu_dropin.py) feeds the same call sequence to the B200 engine and requires byte-identical JSON."""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.dropin import world  # noqa: E402

SCRIPT = os.environ.get("TRX_REFERENCE_SCRIPT", "/root/reference/retrieve/retrieve_faiss.py")
OUT = os.path.join(ROOT, "tests", "golden", "dropin", "golden.json")


def record(tmp):
    data, out = os.path.join(tmp, "data"), os.path.join(tmp, "out")
    world.write_world(data)
    gold = {"script": "retrieve/retrieve_faiss.py (unmodified)", "scenarios": {}}
    for sc in world.SCENARIOS:
        log = os.path.join(tmp, sc + ".calls.jsonl")
        runs = []
        for attempt in (1, 2):
            if os.path.exists(log):
                os.remove(log)
            p = world.run_script(SCRIPT, sc, data, out, world.ORACLE_FAISS, call_log=log)
            assert p.returncode == 0, p.stderr[-3000:]
            od = world.output_dir(sc, out)
            files = {}
            for name in ("train.json", "val.json", "test.json"):
                with open(os.path.join(od, name)) as f:
                    head = json.load(f)[:2]
                files[name] = {"sha256": world.sha256_file(os.path.join(od, name)), "head": head}
            calls = [json.loads(ln) for ln in open(log)]
            runs.append({"files": files, "calls": calls, "stdout_tail": p.stdout.strip().splitlines()[-3:]})
        assert runs[0]["calls"] == runs[1]["calls"]          # the index object sees the same arrays either way
        assert os.path.exists(os.path.join(world.output_dir(sc, out), "train_fp.pkl"))
        if "--before" not in world.SCENARIOS[sc]:
            assert runs[0]["files"] == runs[1]["files"], "cache branch changed the output"
        # With --before the reference filters train_df only on a cache MISS (retrieve_faiss.py:101-103 sits inside the
        # `if not os.path.exists(train_fp_file)` branch), so on a cache hit `train_id` is the UNFILTERED id column and the
        # same ranks map to different ids.  That is the script's behaviour, recorded as it is: first_run / cache_run.
        gold["scenarios"][sc] = {"first_run": runs[0], "cache_run": runs[1]}
    return gold


def main():
    with tempfile.TemporaryDirectory() as tmp:
        gold = record(tmp)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    for sc, g in gold["scenarios"].items():
        for run in ("first_run", "cache_run"):
            print(sc, run, {k: v["sha256"][:12] for k, v in g[run]["files"].items()}, g[run]["stdout_tail"][-1:])


if __name__ == "__main__":
    main()
