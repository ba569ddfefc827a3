"""TEST-ONLY stand-in for `import faiss`, backed by the CPU oracle (oracle/cpu_flat.py) -- used where there is no
GPU (the build container) to run the UNCHANGED reference script and record what it produces
(tests/golden/make_dropin_golden.py).  It also logs every call the script makes on the index object, so that the
GPU test can check the engine is fed the same thing.  Never importable from the product package."""
import json
import os

import numpy as np

from oracle import cpu_flat as _oracle

__version__ = "oracle-standin"
METRIC_INNER_PRODUCT, METRIC_L2 = 0, 1
_LOG = os.environ.get("TRX_DROPIN_CALL_LOG")


def _log(rec):
    if _LOG:
        with open(_LOG, "a") as f:
            f.write(json.dumps(rec) + "\n")


class IndexFlat:
    def __init__(self, d, metric=METRIC_L2):
        self.d, self.metric_type, self.ntotal, self.is_trained, self._x = int(d), metric, 0, True, None
        _log({"call": "ctor", "cls": type(self).__name__, "d": int(d)})

    def add(self, x):
        _log({"call": "add", "shape": list(x.shape), "dtype": str(x.dtype), "c_contiguous": bool(x.flags.c_contiguous)})
        x = np.ascontiguousarray(x, dtype="float32")
        assert x.shape[1] == self.d
        self._x = x if self._x is None else np.concatenate([self._x, x])
        self.ntotal = self._x.shape[0]

    def search(self, x, k):
        _log({"call": "search", "shape": list(x.shape), "dtype": str(x.dtype), "k": int(k)})
        return _oracle.search(self._x, np.ascontiguousarray(x, dtype="float32"), k, self.metric_type)


class IndexFlatL2(IndexFlat):
    def __init__(self, d):
        super().__init__(d, METRIC_L2)


class IndexFlatIP(IndexFlat):
    def __init__(self, d):
        super().__init__(d, METRIC_INNER_PRODUCT)
