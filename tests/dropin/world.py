"""Synthetic inputs for running the reference's retrieval script unchanged (retrieve/retrieve_faiss.py:77-130):
the CSV layouts of retrieve/condition_year.sh:1-7 (reaction fingerprints, `canonical_rxn`) and
retrieve/retro_year.sh:9-16 (`product_smiles`, `--before 2012`, train file one directory up), with reaction ids of
the reference's form ``{patent_id}_{n}`` (preprocess/uspto_script/1.get_condition_from_uspto.py:117).
Deterministic; a few SMILES strings repeat so that exact distance ties occur as in the real data."""
import csv
import hashlib
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
STUBS = os.path.join(HERE, "stubs")                 # fake rdkit
ORACLE_FAISS = os.path.join(HERE, "oracle_faiss")   # CPU stand-in for `import faiss` (tests only)
SHIM = os.path.join(ROOT, "textreact_b200", "shim") # the product: `import faiss` -> B200 engine

CONDITION_ARGV = ["--data_path", "{data}/USPTO_condition_year", "--train_file", "USPTO_condition_train.csv",
                  "--valid_file", "USPTO_condition_val.csv", "--test_file", "USPTO_condition_test.csv",
                  "--field", "canonical_rxn", "--output_path", "{out}/USPTO_condition_year"]          # condition_year.sh
RETRO_YEAR_ARGV = ["--data_path", "{data}/USPTO_50K_year", "--train_file", "../USPTO_rxn_smiles.csv", "--before", "2012",
                   "--valid_file", "valid.csv", "--test_file", "test.csv", "--field", "product_smiles",
                   "--output_path", "{out}/USPTO_50K_year/corpus_before_2012"]                       # retro_year.sh:9-16
SCENARIOS = {"condition_year": CONDITION_ARGV, "retro_year_before_2012": RETRO_YEAR_ARGV}


def _smiles(rng, n, pool):
    atoms = ["C", "N", "O", "c1ccccc1", "Cl", "Br", "S", "F", "C(=O)", "C#N"]
    out = []
    for _ in range(n):
        if pool and rng.random() < 0.08:
            out.append(pool[rng.integers(0, len(pool))])          # a repeated structure: exact ties
        else:
            out.append("".join(atoms[j] for j in rng.integers(0, len(atoms), rng.integers(3, 9))))
            pool.append(out[-1])
    return out


def write_world(data_dir, n_train=2500, n_val=150, n_test=150, n_rxn=3000, n_q=120):
    rng = np.random.default_rng(20241017)
    pool = []
    cond = os.path.join(data_dir, "USPTO_condition_year")
    os.makedirs(cond, exist_ok=True)
    agents = ["", "CCO", "O", "ClCCl", "[Pd]", "CN(C)C=O"]
    rid = 0
    for name, n in (("train", n_train), ("val", n_val), ("test", n_test)):
        with open(os.path.join(cond, f"USPTO_condition_{name}.csv"), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["id", "canonical_rxn", "year", "catalyst1", "solvent1", "solvent2", "reagent1", "reagent2"])
            for s in _smiles(rng, n, pool):
                p = _smiles(rng, 1, pool)[0]
                w.writerow([f"US{6000000 + rid // 3:08d}_{rid % 3}", f"{s}>>{p}", int(rng.integers(1990, 2017))]
                           + [agents[j] for j in rng.integers(0, len(agents), 5)])
                rid += 1
    retro = os.path.join(data_dir, "USPTO_50K_year")
    os.makedirs(retro, exist_ok=True)
    with open(os.path.join(data_dir, "USPTO_rxn_smiles.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["id", "product_smiles", "year"])
        for s in _smiles(rng, n_rxn, pool):
            w.writerow([f"US{7000000 + rid // 2:08d}_{rid % 2}", s, int(rng.integers(2000, 2017))])
            rid += 1
    for name in ("valid", "test"):
        with open(os.path.join(retro, f"{name}.csv"), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["id", "product_smiles"])
            for s in _smiles(rng, n_q, pool):
                w.writerow([f"US{8000000 + rid:08d}_0", s])
                rid += 1


def run_script(script, scenario, data_dir, out_dir, faiss_path, call_log=None, extra_env=None):
    """Runs `python <script> <argv of the scenario's shell script>` from the script's directory, with the fake rdkit and
    the chosen `faiss` package first on PYTHONPATH.  -> CompletedProcess."""
    argv = [a.format(data=data_dir, out=out_dir) for a in SCENARIOS[scenario]]
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([STUBS, faiss_path, ROOT] + ([env["PYTHONPATH"]] if env.get("PYTHONPATH") else []))
    if call_log:
        env["TRX_DROPIN_CALL_LOG"] = call_log
    env.update(extra_env or {})
    return subprocess.run([sys.executable, script] + argv, cwd=os.path.dirname(script), env=env, capture_output=True,
                          text=True, timeout=900)


def output_dir(scenario, out_dir):
    return SCENARIOS[scenario][SCENARIOS[scenario].index("--output_path") + 1].format(out=out_dir)


def sha256_file(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()
