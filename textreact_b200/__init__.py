"""textreact_b200 -- B200-native exact dense retrieval for TextReact's SMILES-to-text step.

A drop-in for the ``faiss`` flat index used by retrieve/retrieve_faiss.py (thomas0809/textreact):
``IndexFlatIP/IndexFlatL2(d)``, ``index.add(xb)``, ``index.search(xq, k) -> (D, I)``, plus the
gold-removed exclusion mask and the ``{id, nn}`` JSON hand-off.  All arithmetic runs in
hand-written sm_100a CUDA behind the C ABI of include/trx.h; there is no CPU fallback.
"""
from .index import (IndexFlat, IndexFlatIP, IndexFlatL2, METRIC_INNER_PRODUCT, METRIC_L2,  # noqa: F401
                    PATH_AUTO, PATH_EXACT, PATH_STREAM, PATH_UMMA, merge_topk)

__all__ = ["IndexFlat", "IndexFlatIP", "IndexFlatL2", "METRIC_INNER_PRODUCT", "METRIC_L2", "merge_topk",
           "PATH_AUTO", "PATH_EXACT", "PATH_STREAM", "PATH_UMMA"]
