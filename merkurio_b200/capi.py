"""ctypes binding of include/merkurio_cuda.h — plumbing for the tests, bench.py and the Python
mirror of the host interface. All matching runs inside libmerkurio_cuda.so on the GPU; if the
library is missing this module raises (there is no fallback path)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Optional, Sequence

import numpy as np

import os

# MK_CUDA_LIB: another build of the same library (the -DMK_DEBUG_CHECKS build with device-side asserts)
LIB_PATH = Path(os.environ.get("MK_CUDA_LIB") or Path(__file__).resolve().parent / "lib" / "libmerkurio_cuda.so")

MK_ENC_ASCII, MK_ENC_BAM4 = 0, 1
MK_MODE_FLAG, MK_MODE_PATTERN_SET, MK_MODE_ALL_HITS = 0, 1, 2
MK_FEATURE_DUAL8, MK_FEATURE_GATE = 1, 2  # mk_engine_info.features (ASCII tables: bits 0..7, BAM4: bits 8..15)

EXPORTS = (
    "mk_engine_create", "mk_engine_destroy", "mk_engine_get_info", "mk_slot_buffers", "mk_scan_submit",
    "mk_scan_wait", "mk_scan_host", "mk_scan_device", "mk_scan_device_submit", "mk_scan_host_uniform", "mk_last_error", "mk_version",
    "mk_engine_scan_kernel", "mk_tables_create", "mk_engine_create_shared", "mk_tables_destroy",
)


class MkPatterns(C.Structure):
    _fields_ = [("bytes", C.c_void_p), ("off", C.c_void_p), ("n", C.c_uint32)]


class MkConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("case_insensitive", C.c_int32), ("n_slots", C.c_uint32),
                ("max_batch_records", C.c_uint32), ("max_batch_bytes", C.c_uint64), ("hit_capacity", C.c_uint64)]


class MkResult(C.Structure):
    _fields_ = [("record_flags", C.c_void_p), ("n_records", C.c_uint32), ("reserved", C.c_uint32),
                ("hits", C.c_void_p), ("n_hits", C.c_uint64), ("bases_scanned", C.c_uint64),
                ("device_ns", C.c_uint64), ("scan_ns", C.c_uint64), ("verify_ns", C.c_uint64),
                ("n_candidates", C.c_uint64), ("n_rescans", C.c_uint32),
                ("reserved2", C.c_uint32), ("d_record_flags", C.c_void_p), ("d_hits", C.c_void_p)]


class MkEngineInfo(C.Structure):
    _fields_ = [("n_patterns", C.c_uint32), ("min_len", C.c_uint32), ("max_len", C.c_uint32),
                ("seed_q", C.c_uint32 * 2), ("seed_d", C.c_uint32 * 2), ("n_seeds", C.c_uint32 * 2),
                ("filter_log2_bits", C.c_uint32 * 2), ("filter_hashes", C.c_uint32 * 2),
                ("filter_bytes", C.c_uint64 * 2),
                ("filter_in_smem", C.c_uint32 * 2), ("table_bytes", C.c_uint64 * 2),
                ("sm_count", C.c_uint32), ("features", C.c_uint32)]


HIT_DTYPE = np.dtype([("record", "<u4"), ("start", "<u4"), ("pattern", "<u4"), ("len", "<u4")])


class MkError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"merkurio_cuda error {code}: {msg}")
        self.code = code
        self.message = msg


_lib = None


def load():
    """Load libmerkurio_cuda.so (built by merkurio_b200.build). Raises if it is not there."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise FileNotFoundError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`; "
                                    "the matching engine has no non-CUDA implementation")
        L = C.CDLL(str(LIB_PATH))
        L.mk_engine_create.argtypes = [C.POINTER(MkPatterns), C.POINTER(MkConfig), C.POINTER(C.c_void_p)]
        L.mk_engine_create.restype = C.c_int
        L.mk_engine_destroy.argtypes = [C.c_void_p]
        L.mk_engine_destroy.restype = None
        L.mk_engine_get_info.argtypes = [C.c_void_p, C.POINTER(MkEngineInfo)]
        L.mk_engine_get_info.restype = C.c_int
        L.mk_slot_buffers.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.mk_slot_buffers.restype = C.c_int
        L.mk_scan_submit.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int, C.c_int, C.c_int]
        L.mk_scan_submit.restype = C.c_int
        L.mk_scan_wait.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(MkResult)]
        L.mk_scan_wait.restype = C.c_int
        L.mk_scan_host.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_int]
        L.mk_scan_host.restype = C.c_int
        L.mk_scan_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_int, C.c_int, C.POINTER(MkResult)]
        L.mk_scan_device.restype = C.c_int
        L.mk_scan_device_submit.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_int,
                                            C.c_int, C.c_int]
        L.mk_scan_device_submit.restype = C.c_int
        L.mk_scan_host_uniform.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int]
        L.mk_scan_host_uniform.restype = C.c_int
        L.mk_engine_scan_kernel.argtypes = [C.c_void_p, C.c_int]
        L.mk_engine_scan_kernel.restype = C.c_char_p
        L.mk_tables_create.argtypes = [C.POINTER(MkPatterns), C.c_int, C.POINTER(C.c_void_p)]
        L.mk_tables_create.restype = C.c_int
        L.mk_engine_create_shared.argtypes = [C.c_void_p, C.POINTER(MkConfig), C.POINTER(C.c_void_p)]
        L.mk_engine_create_shared.restype = C.c_int
        L.mk_tables_destroy.argtypes = [C.c_void_p]
        L.mk_tables_destroy.restype = None
        L.mk_last_error.argtypes = []
        L.mk_last_error.restype = C.c_char_p
        L.mk_version.argtypes = []
        L.mk_version.restype = C.c_char_p
        _lib = L
    return _lib


def _check(rc: int):
    if rc != 0:
        raise MkError(rc, load().mk_last_error().decode("utf-8", "replace"))


class ScanResult:
    """Host copy of one mk_result."""

    def __init__(self, r: MkResult, copy: bool = True):
        self.n_records = r.n_records
        self.n_hits = int(r.n_hits)
        self.bases_scanned = int(r.bases_scanned)
        self.device_ns = int(r.device_ns)
        self.scan_ns = int(r.scan_ns)
        self.verify_ns = int(r.verify_ns)
        self.n_candidates = int(r.n_candidates)
        self.n_rescans = int(r.n_rescans)
        self.d_record_flags = r.d_record_flags
        self.d_hits = r.d_hits
        nw = (r.n_records + 63) // 64
        if r.record_flags and nw:
            a = np.ctypeslib.as_array(C.cast(r.record_flags, C.POINTER(C.c_uint64)), shape=(nw,))
            self.flags = a.copy() if copy else a
        elif r.record_flags:
            self.flags = np.zeros(0, dtype=np.uint64)
        else:
            self.flags = None  # not fetched (device-resident scan without fetch)
        if r.hits and r.n_hits:
            buf = (C.c_uint8 * (self.n_hits * 16)).from_address(r.hits)
            a = np.frombuffer(buf, dtype=HIT_DTYPE)
            self.hits = a.copy() if copy else a
        else:
            self.hits = np.zeros(0, dtype=HIT_DTYPE)

    def flagged_records(self) -> np.ndarray:
        bits = np.unpackbits(self.flags.view(np.uint8), bitorder="little")[: self.n_records]
        return np.nonzero(bits)[0]


def _pattern_blob(patterns: Sequence[bytes]):
    pats = [bytes(p) for p in patterns]
    blob = np.frombuffer(b"".join(pats), dtype=np.uint8).copy() if sum(map(len, pats)) else np.zeros(1, np.uint8)
    off = np.zeros(len(pats) + 1, dtype=np.uint32)
    if pats:
        off[1:] = np.cumsum([len(p) for p in pats])
    return pats, blob, off


class Tables:
    """Host side of a query set (mk_tables_create): built once, shared by the engines created from it."""

    def __init__(self, patterns: Sequence[bytes], case_insensitive: bool = False):
        self.patterns, blob, off = _pattern_blob(patterns)
        self.case_insensitive = case_insensitive
        mp = MkPatterns(blob.ctypes.data, off.ctypes.data, len(self.patterns))
        h = C.c_void_p()
        _check(load().mk_tables_create(C.byref(mp), int(case_insensitive), C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            load().mk_tables_destroy(self._h)
            self._h = None

    __del__ = close


class Engine:
    """One engine per GPU (mk_engine_create). `patterns` is the sorted unique query list, or a `Tables` object
    (mk_engine_create_shared)."""

    def __init__(self, patterns, device: int = 0, case_insensitive: bool = False, n_slots: int = 2,
                 max_batch_bytes: int = 64 << 20, max_batch_records: int = 1 << 20, hit_capacity: int = 0):
        L = load()
        h = C.c_void_p()
        if isinstance(patterns, Tables):
            pats = patterns.patterns
            cfg = MkConfig(device, int(patterns.case_insensitive), n_slots, max_batch_records, max_batch_bytes, hit_capacity)
            _check(L.mk_engine_create_shared(patterns._h, C.byref(cfg), C.byref(h)))
        else:
            pats, self._blob, self._off = _pattern_blob(patterns)
            mp = MkPatterns(self._blob.ctypes.data, self._off.ctypes.data, len(pats))
            cfg = MkConfig(device, int(case_insensitive), n_slots, max_batch_records, max_batch_bytes, hit_capacity)
            _check(L.mk_engine_create(C.byref(mp), C.byref(cfg), C.byref(h)))
        self._h = h
        self.n_slots = n_slots
        self.max_batch_bytes = max_batch_bytes
        self.max_batch_records = max_batch_records
        self.patterns = pats

    def close(self):
        if getattr(self, "_h", None):
            load().mk_engine_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def info(self) -> MkEngineInfo:
        out = MkEngineInfo()
        _check(load().mk_engine_get_info(self._h, C.byref(out)))
        return out

    def scan_kernel(self, enc: int = MK_ENC_ASCII) -> str:
        """Name of the scan kernel of that encoding's tables ("" before the first batch)."""
        return load().mk_engine_scan_kernel(self._h, enc).decode()

    # -- pinned slot path (mk_slot_buffers / mk_scan_submit / mk_scan_wait) ------------------------
    def slot_arrays(self, slot: int):
        seq, off, lens = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(load().mk_slot_buffers(self._h, slot, C.byref(seq), C.byref(off), C.byref(lens)))
        a = np.ctypeslib.as_array(C.cast(seq, C.POINTER(C.c_uint8)), shape=(self.max_batch_bytes,))
        o = np.ctypeslib.as_array(C.cast(off, C.POINTER(C.c_uint64)), shape=(self.max_batch_records + 1,))
        l = np.ctypeslib.as_array(C.cast(lens, C.POINTER(C.c_uint32)), shape=(max(self.max_batch_records, 1),))
        return a, o, l

    def submit(self, slot: int, n_records: int, n_units: int, enc: int, mode: int, use_lens: bool = False):
        _check(load().mk_scan_submit(self._h, slot, n_records, n_units, int(use_lens), enc, mode))

    def wait(self, slot: int, copy: bool = True) -> ScanResult:
        r = MkResult()
        _check(load().mk_scan_wait(self._h, slot, C.byref(r)))
        return ScanResult(r, copy)

    def scan_host_async(self, slot: int, seq, off, lens, n_records: int, n_units: int, enc: int, mode: int):
        """mk_scan_host with raw addresses or numpy arrays (caller keeps them alive until wait)."""
        sp = seq.ctypes.data if isinstance(seq, np.ndarray) else seq
        op = off.ctypes.data if isinstance(off, np.ndarray) else off
        lp = None if lens is None else (lens.ctypes.data if isinstance(lens, np.ndarray) else lens)
        _check(load().mk_scan_host(self._h, slot, sp, op, lp, n_records, n_units, enc, mode))

    def scan(self, seq: np.ndarray, off: np.ndarray, mode: int = MK_MODE_ALL_HITS, enc: int = MK_ENC_ASCII,
             lens: Optional[np.ndarray] = None, n_units: Optional[int] = None, slot: int = 0) -> ScanResult:
        """One batch from host arrays through slot staging (H2D + scan + sort + D2H)."""
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        if lens is not None:
            lens = np.ascontiguousarray(lens, dtype=np.uint32)
        if n_units is None:
            n_units = int(seq.size) if enc == MK_ENC_ASCII else int(seq.size) * 2
        if seq.size == 0:
            seq = np.zeros(16, dtype=np.uint8)
        self.scan_host_async(slot, seq, off, lens, len(off) - 1, n_units, enc, mode)
        return self.wait(slot)

    # -- device-resident path (mk_scan_device) ------------------------------------------------------
    def scan_device(self, d_seq: int, d_off: int, n_records: int, n_units: int, mode: int = MK_MODE_FLAG,
                    enc: int = MK_ENC_ASCII, d_lens: Optional[int] = None, fetch: bool = False) -> ScanResult:
        r = MkResult()
        _check(load().mk_scan_device(self._h, d_seq, d_off, d_lens, n_records, n_units, enc, mode, int(fetch), C.byref(r)))
        return ScanResult(r, copy=True)

    def scan_host_uniform_async(self, slot: int, seq, n_records: int, record_len: int, enc: int, mode: int):
        """mk_scan_host_uniform: fixed-length records, no offset array (caller keeps seq alive until wait)."""
        sp = seq.ctypes.data if isinstance(seq, np.ndarray) else seq
        _check(load().mk_scan_host_uniform(self._h, slot, sp, n_records, record_len, enc, mode))

    def scan_device_submit(self, slot: int, d_seq: int, d_off: int, n_records: int, n_units: int, mode: int = MK_MODE_FLAG,
                           enc: int = MK_ENC_ASCII, d_lens: Optional[int] = None, fetch: bool = False):
        """Asynchronous mk_scan_device on a slot's stream; collect the result with wait(slot)."""
        _check(load().mk_scan_device_submit(self._h, slot, d_seq, d_off, d_lens, n_records, n_units, enc, mode, int(fetch)))


def version() -> str:
    return load().mk_version().decode()
