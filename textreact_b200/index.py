"""FAISS flat-index surface over the B200 engine.

Mirrors exactly what TextReact's retrieval script calls on the ``faiss`` object
(reference: retrieve/retrieve_faiss.py:62-74):

    index = faiss.IndexFlatL2(d)            # :65   (IndexFlatIP for the 768-d neural retriever)
    index.add(train_fps)                    # :66
    distance, rank = index.search(query_fps, k)   # :71

Inputs of any numeric dtype / memory order are coerced with
``np.ascontiguousarray(x, dtype='float32')`` as FAISS's python wrapper does -- the script
passes int64 difference fingerprints (:26) and int8 Morgan bits (:40).  A wrong second
dimension raises ``AssertionError`` (FAISS behaviour); engine failures raise ``RuntimeError``
/ ``MemoryError``.  Extensions, all keyword-only and ignored by the reference script:
``exclude=`` (gold-removed mode, textreact/dataset.py:74-76 lifted into the engine),
``set_groups``, ``set_row_attr`` / ``attr_below=`` (the ``--before`` year restriction, :102-103),
``search_self`` (train->train search, :114-115), ``dedup=True`` (distinct text groups, the consumer's
``deduplicate_neighbors``, textreact/dataset.py:46-56), torch CUDA tensors in / out, ``device=``.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import _lib
from ._lib import PATH_AUTO, PATH_EXACT, PATH_STREAM, PATH_UMMA  # noqa: F401

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1

try:  # torch is plumbing only (device tensors, streams); numpy callers never need it
    if os.environ.get("TRX_NO_TORCH"):      # e.g. under compute-sanitizer: keep the process small
        raise ImportError("TRX_NO_TORCH set")
    import torch
except Exception:  # pragma: no cover
    torch = None


_CUDA_STREAM_LEGACY = 0x1   # cudaStreamLegacy: the C ABI reads a NULL stream as "use the index's own stream"


def _stream_handle(device):
    """torch's current stream on `device` as a non-NULL handle (the default stream has handle 0)."""
    return torch.cuda.current_stream(device).cuda_stream or _CUDA_STREAM_LEGACY


def _is_torch(x):
    return torch is not None and isinstance(x, torch.Tensor)


def _as_f32_matrix(x, d):
    """-> (pointer, n, keepalive, on_device)"""
    if _is_torch(x) and x.is_cuda:
        assert x.dim() == 2 and x.shape[1] == d, f"expected shape [n, {d}], got {tuple(x.shape)}"
        t = x.detach().to(torch.float32).contiguous()
        return t.data_ptr(), t.shape[0], t, True
    if _is_torch(x):
        x = x.detach().cpu().numpy()
    a = np.ascontiguousarray(x, dtype="float32")
    assert a.ndim == 2 and a.shape[1] == d, f"expected shape [n, {d}], got {a.shape}"
    return a.ctypes.data, a.shape[0], a, False


def _as_i32_vector(x, n, what):
    if _is_torch(x) and x.is_cuda:
        t = x.detach().to(torch.int32).contiguous()
        assert t.dim() == 1 and t.shape[0] == n, f"{what}: expected {n} entries, got {tuple(t.shape)}"
        return t.data_ptr(), t
    if _is_torch(x):
        x = x.detach().cpu().numpy()
    a = np.ascontiguousarray(x, dtype=np.int32)
    assert a.ndim == 1 and a.shape[0] == n, f"{what}: expected {n} entries, got {a.shape}"
    return a.ctypes.data, a


class IndexFlat:
    """Exact (brute-force) index; ``metric`` is METRIC_L2 (squared distances, ascending) or
    METRIC_INNER_PRODUCT (scores, descending)."""

    def __init__(self, d, metric=METRIC_L2, *, device=None):
        L = _lib.lib()
        if device is None:
            device = torch.cuda.current_device() if (torch is not None and torch.cuda.is_available()) else 0
        self._h = ctypes.c_void_p()
        self._L = L
        _lib.check(L.trx_create(int(d), int(metric), int(device), ctypes.byref(self._h)), "IndexFlat")
        self.d = int(d)
        self.metric_type = int(metric)
        self.is_trained = True
        self.device = int(device)

    # -- FAISS attributes -------------------------------------------------------------------
    @property
    def ntotal(self):
        return int(self._L.trx_ntotal(self._h))

    # -- FAISS methods ----------------------------------------------------------------------
    def add(self, x):
        if isinstance(x, np.ndarray) and x.dtype.name in _lib.DTYPES and x.dtype != np.float32 and x.ndim == 2 \
                and x.flags.c_contiguous:
            # int8 Morgan bits / int64 difference counts (retrieve_faiss.py:26, :40): ship the raw array,
            # widen on the device -- same values as FAISS's np.ascontiguousarray(x, dtype='float32')
            assert x.shape[1] == self.d, f"expected shape [n, {self.d}], got {x.shape}"
            _lib.check(self._L.trx_add_typed(self._h, x.ctypes.data, x.shape[0], _lib.DTYPES[x.dtype.name]), "add")
            return
        ptr, n, keep, on_dev = _as_f32_matrix(x, self.d)
        if on_dev:   # trx_add copies on the index's own stream: the producer of `x` must have finished
            torch.cuda.current_stream(keep.device).synchronize()
        _lib.check(self._L.trx_add(self._h, ptr, n), "add")
        del keep

    def add_npy(self, path, chunk_rows=262144):
        """``index.add(np.load(path))`` without holding the array in host memory: the file (what ``np.save`` wrote --
        the reference's fingerprint cache ``train_fp.pkl`` is one, retrieve/retrieve_faiss.py:106-110 -- or a
        Tevatron embedding dump) is memory-mapped and streamed to the device in chunks, any numeric dtype."""
        a = np.load(path, mmap_mode="r")
        assert a.ndim == 2 and a.shape[1] == self.d, f"expected shape [n, {self.d}], got {a.shape}"
        self.reserve(self.ntotal + a.shape[0])
        for r0 in range(0, a.shape[0], int(chunk_rows)):
            self.add(np.ascontiguousarray(a[r0:r0 + int(chunk_rows)]))
        return a.shape[0]

    def search(self, x, k, *, D=None, I=None, exclude=None, attr_below=None, dedup=False, params=None):
        assert k > 0
        ptr, nq, keep, on_dev = _as_f32_matrix(x, self.d)
        return self._search(ptr, nq, keep, on_dev, k, D, I, exclude, attr_below, None, dedup)

    def search_self(self, k, start=0, stop=None, *, D=None, I=None, exclude=None, attr_below=None, dedup=False,
                    device=False):
        """``search(xb[start:stop], k)`` with the rows already stored in the index as the queries -- the
        reference's train->train search (retrieve/retrieve_faiss.py:114-115, ``query_fps = train_fps``)
        without sending the corpus to the GPU a second time.  numpy out unless ``device=True``."""
        assert k > 0
        stop = self.ntotal if stop is None else int(stop)
        start = int(start)
        assert 0 <= start <= stop <= self.ntotal, f"rows [{start}, {stop}) outside [0, {self.ntotal})"
        keep = torch.empty(0, device=torch.device("cuda", self.device)) if device else None
        return self._search(None, stop - start, keep, bool(device), k, D, I, exclude, attr_below, start, dedup)

    def _search(self, ptr, nq, keep, on_dev, k, D, I, exclude, attr_below, self_row0, dedup=False):
        ex_ptr, ex_keep = (None, None)
        if exclude is not None:
            ex_ptr, ex_keep = _as_i32_vector(exclude, nq, "exclude")
        stream = None
        if on_dev:
            Dt = D if D is not None else torch.empty((nq, k), dtype=torch.float32, device=keep.device)
            It = I if I is not None else torch.empty((nq, k), dtype=torch.int64, device=keep.device)
            assert Dt.shape == (nq, k) and It.shape == (nq, k) and Dt.is_contiguous() and It.is_contiguous()
            assert Dt.dtype == torch.float32 and It.dtype == torch.int64 and Dt.is_cuda and It.is_cuda
            dptr, iptr = Dt.data_ptr(), It.data_ptr()
            stream = _stream_handle(keep.device)
            out = (Dt, It)
        else:
            Dn = D if D is not None else np.empty((nq, k), dtype=np.float32)
            In = I if I is not None else np.empty((nq, k), dtype=np.int64)
            assert Dn.shape == (nq, k) and In.shape == (nq, k)
            assert Dn.dtype == np.float32 and In.dtype == np.int64
            assert Dn.flags.c_contiguous and In.flags.c_contiguous
            dptr, iptr = Dn.ctypes.data, In.ctypes.data
            out = (Dn, In)
        # per-call modes travel in the call (trx_search_ex): nothing is left set on the index
        sp = _lib.TrxSearchParams(ex_ptr, 2147483647 if attr_below is None else int(attr_below), 1 if dedup else 0,
                                  -1 if self_row0 is None else int(self_row0))
        _lib.check(self._L.trx_search_ex(self._h, ptr, nq, int(k), ctypes.byref(sp), dptr, iptr, stream),
                   "search" if self_row0 is None else "search_self")
        del keep, ex_keep
        return out

    # -- two-phase search (row-sharded multi-GPU mode; see include/trx.h) ---------------------
    def search_begin(self, x, k, nb, *, exclude=None, attr_below=None):
        """First half for ONE batch of CUDA-tensor queries: prefilter up to the masked, sorted candidate lists.
        Returns the payload to exchange: float32 CUDA [nq, nb + 1] = the nb best prefilter scores per query, then eps."""
        ptr, nq, keep, on_dev = _as_f32_matrix(x, self.d)
        assert on_dev, "search_begin takes CUDA tensors"
        ex_ptr, ex_keep = (None, None)
        if exclude is not None:
            ex_ptr, ex_keep = _as_i32_vector(exclude, nq, "exclude")
        payload = torch.empty((nq, int(nb) + 1), dtype=torch.float32, device=keep.device)
        sp = _lib.TrxSearchParams(ex_ptr, 2147483647 if attr_below is None else int(attr_below), 0, -1)
        _lib.check(self._L.trx_search_begin(self._h, ptr, nq, int(k), ctypes.byref(sp), int(nb), payload.data_ptr(),
                                            _stream_handle(keep.device)), "search_begin")
        self._pending = (nq, int(k), keep.device)
        del keep, ex_keep
        return payload

    def search_finish(self, floor=None, *, D=None, I=None):
        """Second half: exact rescore of the candidates at or above ``floor`` (float32 CUDA [nq]; None = plain local
        top-k).  Returns (D, I) CUDA tensors; rows that cannot reach the global top-k are missing (-1 padded)."""
        assert getattr(self, "_pending", None) is not None, "search_finish without search_begin"
        nq, k, dev = self._pending
        self._pending = None
        Dt = D if D is not None else torch.empty((nq, k), dtype=torch.float32, device=dev)
        It = I if I is not None else torch.empty((nq, k), dtype=torch.int64, device=dev)
        fptr = None
        if floor is not None:
            assert floor.is_cuda and floor.dtype == torch.float32 and floor.shape == (nq,) and floor.is_contiguous()
            fptr = floor.data_ptr()
        _lib.check(self._L.trx_search_finish(self._h, fptr, Dt.data_ptr(), It.data_ptr(), _stream_handle(dev)),
                   "search_finish")
        return Dt, It

    def reset(self):
        _lib.check(self._L.trx_reset(self._h), "reset")

    def reconstruct(self, key):
        """FAISS ``index.reconstruct(i)``: the stored fp32 row."""
        return self.reconstruct_n(int(key), 1)[0]

    def reconstruct_n(self, n0=0, ni=-1):
        """FAISS ``index.reconstruct_n(n0, ni)``: rows [n0, n0+ni) as a float32 array."""
        ni = self.ntotal - n0 if ni < 0 else ni
        out = np.empty((ni, self.d), dtype=np.float32)
        _lib.check(self._L.trx_reconstruct(self._h, int(n0), int(ni), out.ctypes.data), "reconstruct")
        return out

    # -- extensions -------------------------------------------------------------------------
    def reserve(self, n):
        _lib.check(self._L.trx_reserve(self._h, int(n)), "reserve")

    def set_groups(self, groups):
        """Per-row exclusion group (text-dedup group / patent id), one int32 per added row."""
        if groups is None:
            _lib.check(self._L.trx_set_groups(self._h, None, 0), "set_groups")
            return
        ptr, keep = _as_i32_vector(groups, self.ntotal, "groups")
        if _is_torch(keep) and keep.is_cuda:
            torch.cuda.current_stream(keep.device).synchronize()
        _lib.check(self._L.trx_set_groups(self._h, ptr, self.ntotal), "set_groups")
        del keep

    def set_row_attr(self, attr):
        """Per-row integer attribute (e.g. year), one int32 per added row; ``search(..., attr_below=T)``
        then only returns rows with attr < T -- the ``--before`` restriction of
        retrieve/retrieve_faiss.py:102-103 without rebuilding the index per split."""
        if attr is None:
            _lib.check(self._L.trx_set_row_attr(self._h, None, 0), "set_row_attr")
            return
        ptr, keep = _as_i32_vector(attr, self.ntotal, "attr")
        if _is_torch(keep) and keep.is_cuda:
            torch.cuda.current_stream(keep.device).synchronize()
        _lib.check(self._L.trx_set_row_attr(self._h, ptr, self.ntotal), "set_row_attr")
        del keep

    def set_id_offset(self, offset):
        _lib.check(self._L.trx_set_id_offset(self._h, int(offset)), "set_id_offset")

    def set_option(self, key, value):
        _lib.check(self._L.trx_set_option(self._h, key.encode(), float(value)), f"set_option({key})")

    def get_option(self, key):
        v = ctypes.c_double()
        _lib.check(self._L.trx_get_option(self._h, key.encode(), ctypes.byref(v)), f"get_option({key})")
        return v.value

    def stats(self):
        s = _lib.TrxStats()
        _lib.check(self._L.trx_stats(self._h, ctypes.byref(s)), "stats")
        return s.as_dict()

    def debug_scores_umma(self, xq, row0, n):
        """bf16 tcgen05 scores of CUDA tensor ``xq`` against rows [row0, row0+n) (tests only)."""
        ptr, nq, keep, on_dev = _as_f32_matrix(xq, self.d)
        assert on_dev, "debug_scores_umma takes a CUDA tensor"
        out = torch.empty((nq, n), dtype=torch.float32, device=keep.device)
        _lib.check(self._L.trx_debug_scores_umma(self._h, ptr, nq, int(row0), int(n), out.data_ptr(),
                                                 _stream_handle(keep.device)), "debug_scores")
        return out

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.trx_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class IndexFlatIP(IndexFlat):
    def __init__(self, d, **kw):
        super().__init__(d, METRIC_INNER_PRODUCT, **kw)


class IndexFlatL2(IndexFlat):
    def __init__(self, d, **kw):
        super().__init__(d, METRIC_L2, **kw)


def merge_topk(Dg, Ig, metric):
    """K5: merge per-shard results.  Dg [G, nq, k] float32, Ig [G, nq, k] int64 CUDA tensors,
    every list best-first with global ids, shards in ascending id order."""
    assert _is_torch(Dg) and Dg.is_cuda and Ig.is_cuda and Dg.shape == Ig.shape and Dg.dim() == 3
    Dg, Ig = Dg.contiguous(), Ig.contiguous()
    G, nq, k = Dg.shape
    D = torch.empty((nq, k), dtype=torch.float32, device=Dg.device)
    I = torch.empty((nq, k), dtype=torch.int64, device=Dg.device)
    L = _lib.lib()
    _lib.check(L.trx_merge_topk(int(metric), Dg.data_ptr(), Ig.data_ptr(), G, nq, k, D.data_ptr(), I.data_ptr(),
                                _stream_handle(Dg.device)), "merge_topk")
    return D, I
