"""ctypes wrapper of libmerkurio_synth.so (mk_synth.h): deterministic synthetic reads of the
BASELINE shapes for the benchmark and the parity tests. Not part of the matching product."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

LIB_PATH = Path(__file__).resolve().parent.parent / "lib" / "libmerkurio_synth.so"


class MksParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("n_reads", C.c_uint64), ("read_len", C.c_uint32), ("k", C.c_uint32),
                ("n_queries", C.c_uint32), ("plant_per_65536", C.c_uint32), ("nrun_per_65536", C.c_uint32),
                ("reserved", C.c_uint32)]


_lib = None


def load():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise FileNotFoundError(f"{LIB_PATH} is missing: run __graft_entry__.build()")
        L = C.CDLL(str(LIB_PATH))
        L.mks_queries.argtypes = [C.POINTER(MksParams), C.c_void_p]
        L.mks_fill_host.argtypes = [C.POINTER(MksParams), C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
        L.mks_fill_device.argtypes = [C.POINTER(MksParams), C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mks_last_error.restype = C.c_char_p
        _lib = L
    return _lib


class Synth:
    """One synthetic data set (seed, shape, query sampling)."""

    def __init__(self, seed: int, n_reads: int, read_len: int = 150, k: int = 31, n_queries: int = 1000,
                 plant_per_65536: int = 655, nrun_per_65536: int = 328):
        self.p = MksParams(seed, n_reads, read_len, k, n_queries, plant_per_65536, nrun_per_65536, 0)
        self.queries = np.zeros(max(n_queries * k, 1), dtype=np.uint8)
        if load().mks_queries(C.byref(self.p), self.queries.ctypes.data) != 0:
            raise ValueError(load().mks_last_error().decode())
        self.read_len, self.k, self.n_queries, self.n_reads = read_len, k, n_queries, n_reads

    def query_list(self):
        q = self.queries[: self.n_queries * self.k].reshape(self.n_queries, self.k)
        return [row.tobytes() for row in q]

    def host_reads(self, r0: int, r1: int, enc: int = 0):
        """(seq bytes, offsets) of reads [r0, r1), offsets relative to r0."""
        n = r1 - r0
        nbytes = n * self.read_len // (2 if enc else 1)
        out = np.zeros(nbytes + 16, dtype=np.uint8)
        off = np.zeros(n + 1, dtype=np.uint64)
        if load().mks_fill_host(C.byref(self.p), self.queries.ctypes.data, r0, r1, enc, out.ctypes.data, off.ctypes.data) != 0:
            raise ValueError(load().mks_last_error().decode())
        return out[:nbytes], off

    def device_reads(self, d_queries: int, r0: int, r1: int, d_out: int, d_off: int, enc: int = 0, stream: int = 0):
        """Fill device memory (addresses as ints). d_out needs 16 bytes of slack after the data."""
        if load().mks_fill_device(C.byref(self.p), d_queries, r0, r1, enc, d_out, d_off, stream or None) != 0:
            raise RuntimeError(load().mks_last_error().decode())
