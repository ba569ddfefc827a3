"""ctypes binding of libtrx.so (the C ABI declared in include/trx.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is no
Python or CPU implementation behind it: if the shared object is missing this module
raises at first use, and ``trx_create`` itself fails without a B200-class device.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TRX_LIBTRX") or os.path.join(_HERE, "libtrx.so")   # TRX_LIBTRX: a variant build (experiments)

TRX_OK, TRX_EINVAL, TRX_ENOMEM, TRX_ECUDA, TRX_ENODEV = 0, 1, 2, 3, 4
PATH_AUTO, PATH_EXACT, PATH_STREAM, PATH_UMMA = 0, 1, 2, 3
# TRX_DTYPE_* by numpy dtype name (trx_add_typed)
DTYPES = {"float32": 0, "float64": 1, "float16": 2, "int8": 3, "uint8": 4, "bool": 4, "int16": 5, "int32": 6, "int64": 7}


class TrxStats(ctypes.Structure):
    _fields_ = [
        ("searches", ctypes.c_int64), ("queries", ctypes.c_int64), ("queries_exact", ctypes.c_int64),
        ("queries_uncert", ctypes.c_int64), ("queries_overflow", ctypes.c_int64), ("rescored", ctypes.c_int64),
        ("candidates", ctypes.c_int64), ("last_path", ctypes.c_int32), ("sm_count", ctypes.c_int32),
        ("launches", ctypes.c_int64), ("last_prefilter_ms", ctypes.c_double), ("last_total_ms", ctypes.c_double),
        ("timed_batches", ctypes.c_int64), ("sum_sample_ms", ctypes.c_double), ("sum_prefilter_ms", ctypes.c_double),
        ("sum_rescore_ms", ctypes.c_double), ("sum_total_ms", ctypes.c_double),
        ("queries_second_pass", ctypes.c_int64),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class TrxSearchParams(ctypes.Structure):
    _fields_ = [("exclude", ctypes.c_void_p), ("attr_below", ctypes.c_int32), ("dedup_groups", ctypes.c_int32),
                ("self_row0", ctypes.c_int64)]


# every symbol include/trx.h declares, with its ctypes signature
_vp = ctypes.c_void_p
SIGNATURES = {
    "trx_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(_vp)]),
    "trx_add": (ctypes.c_int, [_vp, _vp, ctypes.c_int64]),
    "trx_add_typed": (ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int]),
    "trx_reserve": (ctypes.c_int, [_vp, ctypes.c_int64]),
    "trx_set_groups": (ctypes.c_int, [_vp, _vp, ctypes.c_int64]),
    "trx_search": (ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int, _vp, _vp, _vp, _vp]),
    "trx_search_ex": (ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int, ctypes.POINTER(TrxSearchParams), _vp, _vp, _vp]),
    "trx_search_begin": (ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int, ctypes.POINTER(TrxSearchParams), ctypes.c_int, _vp, _vp]),
    "trx_search_finish": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "trx_exchange_floor": (ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, _vp, _vp]),
    "trx_search_self": (ctypes.c_int, [_vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, _vp, _vp, _vp, _vp]),
    "trx_set_row_attr": (ctypes.c_int, [_vp, _vp, ctypes.c_int64]),
    "trx_reconstruct": (ctypes.c_int, [_vp, ctypes.c_int64, ctypes.c_int64, _vp]),
    "trx_reset": (ctypes.c_int, [_vp]),
    "trx_destroy": (None, [_vp]),
    "trx_ntotal": (ctypes.c_int64, [_vp]),
    "trx_dim": (ctypes.c_int, [_vp]),
    "trx_metric": (ctypes.c_int, [_vp]),
    "trx_set_id_offset": (ctypes.c_int, [_vp, ctypes.c_int64]),
    "trx_set_option": (ctypes.c_int, [_vp, ctypes.c_char_p, ctypes.c_double]),
    "trx_get_option": (ctypes.c_int, [_vp, ctypes.c_char_p, ctypes.POINTER(ctypes.c_double)]),
    "trx_stats": (ctypes.c_int, [_vp, ctypes.POINTER(TrxStats)]),
    "trx_merge_topk": (ctypes.c_int, [ctypes.c_int, _vp, _vp, ctypes.c_int, ctypes.c_int64, ctypes.c_int, _vp, _vp, _vp]),
    "trx_exchange_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.POINTER(_vp)]),
    "trx_exchange_handle": (ctypes.c_int, [_vp, ctypes.c_char_p]),
    "trx_exchange_connect": (ctypes.c_int, [_vp, ctypes.c_char_p]),
    "trx_exchange_merge": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, ctypes.c_int64, ctypes.c_int, _vp, _vp, _vp]),
    "trx_exchange_merge_slice": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int64,
                                                ctypes.c_int64, _vp, _vp, _vp]),
    "trx_exchange_destroy": (None, [_vp]),
    "trx_device_malloc": (ctypes.c_int, [ctypes.c_int, ctypes.c_size_t, ctypes.POINTER(_vp)]),
    "trx_device_free": (ctypes.c_int, [_vp]),
    "trx_device_copy": (ctypes.c_int, [_vp, _vp, ctypes.c_size_t]),
    "trx_debug_scores_umma": (ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _vp, _vp]),
    "trx_last_error": (ctypes.c_char_p, []),
    "trx_version": (ctypes.c_char_p, []),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C textreact_b200/csrc`).  textreact_b200 has no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    return lib().trx_last_error().decode("utf-8", "replace")


def check(rc: int, what: str):
    if rc == TRX_OK:
        return
    msg = f"{what}: {last_error()} (code {rc})"
    if rc == TRX_ENOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)
