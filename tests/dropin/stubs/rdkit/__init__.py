"""TEST STUB of the `rdkit` names retrieve/retrieve_faiss.py:6-11 imports (rdkit is not installed here and the
featurisers are out of scope, SURVEY.md section 2 row 6: the engine takes dense arrays as given).

Deterministic fake fingerprints with the real ones' container behaviour, so that the UNCHANGED script can run:
  * CreateDifferenceFingerprintForReaction(rxn) -> an iterable of 2048 small signed ints (the script does
    ``np.array([x for x in fp])`` -> int64, retrieve_faiss.py:24-27)
  * GetMorganFingerprintAsBitVect(mol, 2, nBits=1024) + DataStructs.ConvertToNumpyArray(fp, array), which RESIZES the
    zero-length int8 array it is given in place (retrieve_faiss.py:38-41)
Values come from a hash of the SMILES string: similar strings do NOT give similar vectors, which is irrelevant to the
flat search being exercised; duplicates of a string give identical vectors (ties, as in the real data)."""
import hashlib

import numpy as np

from . import RDLogger  # noqa: F401


def _rng(text, salt):
    h = hashlib.sha256((salt + "|" + str(text)).encode()).digest()
    return np.random.default_rng(int.from_bytes(h[:8], "little"))


def fake_difference_counts(smiles, n=2048):
    rng = _rng(smiles, "diff")
    v = np.zeros(n, dtype=np.int64)
    on = rng.choice(n, size=24, replace=False)
    v[on] = rng.integers(-2, 3, size=24)
    return v


def fake_morgan_bits(smiles, n=1024):
    rng = _rng(smiles, "morgan")
    v = np.zeros(n, dtype=np.int8)
    v[rng.choice(n, size=40, replace=False)] = 1
    return v
