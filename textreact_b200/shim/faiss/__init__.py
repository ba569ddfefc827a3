"""``import faiss`` shim: put ``textreact_b200/shim`` on PYTHONPATH and the reference's
retrieve/retrieve_faiss.py (line 14 ``import faiss``; lines 65-71) runs unchanged on the B200
engine.  Only the names that script (and the north_star IndexFlatIP variant) use exist."""
from textreact_b200.index import (IndexFlat, IndexFlatIP, IndexFlatL2,  # noqa: F401
                                  METRIC_INNER_PRODUCT, METRIC_L2)

__version__ = "textreact_b200-shim"
