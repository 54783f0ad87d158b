"""CPU restatement of the reference's host-side logic around the matchers.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module; the product (merkurio_b200/) never does.

Paths cited are relative to the reference repository (lschoenm/MerKurio):
  * query list:        src/helpers.rs:76-163  (parse_pattern_list, read_kmers_from_file)
  * algorithm choice:  src/helpers.rs:203-211, src/cmd_extract.rs:165-171, src/cmd_tag.rs:183-189
  * flag conflicts and output naming: src/helpers.rs:29-68,172-200
  * extract loops, counters, summaries: src/cmd_extract.rs:230-255,285-290,321-406,463-714
  * tag loop: src/cmd_tag.rs:328-364,387-500,504-686
  * log formats: src/logger.rs:41-60,95-190
  * reverse complement / canonical: third-party crate needletail 0.6.3 (Cargo.lock:382-383),
    not in the reference tree; restated from its published behaviour and pinned by
    src/helpers.rs:363-397 (three 32-mers) plus every golden that uses -r.
  * FASTA/FASTQ record model: needletail 0.6.3 (id = header line without the marker, seq() strips
    line breaks, write() re-emits FASTA with its original wrapping, FASTQ as four lines with a
    bare '+'), pinned by tests/fixtures/extract/*.
  * SAM/BAM record model: crate bam 0.1.4 (Cargo.lock:77-78), pinned by tests/fixtures/tag/*.
The matchers themselves live in oracle/mk_oracle.c and are called through ctypes.
"""
from __future__ import annotations

import ctypes
import gzip
import json
import os
import struct
import sys
from dataclasses import dataclass, field
from pathlib import Path
from typing import List, Optional, Sequence, Tuple

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
PROGRAM = "merkurio"
VERSION = "1.0.0"  # crate_version!() of the reference tree (Cargo.toml:3); volatile in all comparisons


class RefError(Exception):
    """An `Err` bubbling to main (the reference prints `Error: <msg>` and exits 1)."""


# ----------------------------------------------------------------------------------------------
# C matchers
# ----------------------------------------------------------------------------------------------
_lib = None


def lib():
    global _lib
    if _lib is None:
        so = ROOT / "oracle" / "_build" / "libmk_oracle.so"
        src = ROOT / "oracle" / "mk_oracle.c"
        if not so.exists() or (src.exists() and src.stat().st_mtime > so.stat().st_mtime):
            sys.path.insert(0, str(ROOT))
            from merkurio_b200.build import build_oracle
            build_oracle()
        L = ctypes.CDLL(str(so))
        u8p, u32p, u64p = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p
        L.mko_generate_masks.argtypes = [u8p, ctypes.c_size_t, u64p, u64p]
        L.mko_generate_masks.restype = ctypes.c_int
        L.mko_tune_q_value.argtypes = [ctypes.c_size_t]
        L.mko_tune_q_value.restype = ctypes.c_int
        L.mko_bndmq_check.argtypes = [ctypes.c_size_t, ctypes.c_size_t]
        L.mko_bndmq_check.restype = ctypes.c_int
        L.mko_bndmq_find.argtypes = [u8p, ctypes.c_size_t, ctypes.c_size_t, u8p, ctypes.c_size_t, u64p, ctypes.c_size_t, ctypes.c_int]
        L.mko_bndmq_find.restype = ctypes.c_long
        L.mko_naive_find.argtypes = [u8p, ctypes.c_size_t, u8p, ctypes.c_size_t, u64p, ctypes.c_size_t]
        L.mko_naive_find.restype = ctypes.c_long
        L.mko_ac_build.argtypes = [u8p, u32p, ctypes.c_uint32, ctypes.c_int]
        L.mko_ac_build.restype = ctypes.c_void_p
        L.mko_ac_free.argtypes = [ctypes.c_void_p]
        L.mko_ac_free.restype = None
        L.mko_ac_n_states.argtypes = [ctypes.c_void_p]
        L.mko_ac_n_states.restype = ctypes.c_uint32
        L.mko_ac_table_bytes.argtypes = [ctypes.c_void_p]
        L.mko_ac_table_bytes.restype = ctypes.c_uint64
        L.mko_ac_find_overlapping.argtypes = [ctypes.c_void_p, u8p, ctypes.c_size_t, u32p, u64p, ctypes.c_size_t, ctypes.c_int]
        L.mko_ac_find_overlapping.restype = ctypes.c_long
        L.mko_ac_scan_batch.argtypes = [ctypes.c_void_p, u8p, u64p, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, u64p, u64p]
        L.mko_ac_scan_batch.restype = ctypes.c_uint64
        L.mko_ac_batch_hits.argtypes = [ctypes.c_void_p, u8p, u64p, u32p, ctypes.c_uint32, u32p, u32p, u32p, ctypes.c_uint64]
        L.mko_ac_batch_hits.restype = ctypes.c_uint64
        _lib = L
    return _lib


def _buf(b: bytes):
    """bytes -> (keepalive array, pointer)"""
    a = np.frombuffer(b, dtype=np.uint8) if len(b) else np.zeros(1, dtype=np.uint8)
    return a, a.ctypes.data


def generate_masks(pattern: bytes) -> Tuple[List[int], int]:
    """src/pattern_preprocessing.rs:24-43"""
    masks = np.zeros(256, dtype=np.uint64)
    acc = np.zeros(1, dtype=np.uint64)
    a, p = _buf(pattern)
    if lib().mko_generate_masks(p, len(pattern), masks.ctypes.data, acc.ctypes.data) != 0:
        raise RefError(f"Pattern length {len(pattern)} is too large for this architecture when using BNDM (max 64).")
    return [int(x) for x in masks], int(acc[0])


def tune_q_value(pattern: str) -> int:
    """src/pattern_matching.rs:213-225"""
    q = lib().mko_tune_q_value(len(pattern.encode()))
    if q == 0:
        raise RefError("Pattern length is too long for BNDMq.")
    return q


class BNDMq:
    """src/pattern_matching.rs:42-154"""

    def __init__(self, pattern: bytes, q: int):
        rc = lib().mko_bndmq_check(len(pattern), q)
        if rc == -2:
            raise RefError("Pattern is empty.")
        if rc == -3:
            raise RefError(f"Invalid q-gram length: {q}. Must be between 1 and pattern length.")
        if rc == -1:
            raise RefError(f"Pattern length {len(pattern)} is too large for this architecture when using BNDM (max 64).")
        self.pattern, self.q = bytes(pattern), q

    def find_all(self, text: bytes) -> List[int]:
        cap = max(len(text), 1)
        out = np.zeros(cap, dtype=np.uint64)
        a, p = _buf(self.pattern)
        t, tp = _buf(text)
        n = lib().mko_bndmq_find(p, len(self.pattern), self.q, tp, len(text), out.ctypes.data, cap, 0)
        return [int(x) for x in out[:n]]

    find_iter = find_all

    def find_match(self, text: bytes) -> bool:
        a, p = _buf(self.pattern)
        t, tp = _buf(text)
        return lib().mko_bndmq_find(p, len(self.pattern), self.q, tp, len(text), None, 0, 1) > 0


def naive_find_all(pattern: bytes, text: bytes) -> List[int]:
    cap = max(len(text), 1)
    out = np.zeros(cap, dtype=np.uint64)
    a, p = _buf(pattern)
    t, tp = _buf(text)
    n = lib().mko_naive_find(p, len(pattern), tp, len(text), out.ctypes.data, cap)
    return [int(x) for x in out[:n]]


def pack_patterns(patterns: Sequence[bytes]):
    off = np.zeros(len(patterns) + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(p) for p in patterns], dtype=np.uint64)
    blob = np.frombuffer(b"".join(patterns), dtype=np.uint8).copy() if patterns else np.zeros(1, dtype=np.uint8)
    return blob, off


class AhoCorasick:
    """aho-corasick 1.1.3 DFA, MatchKind::Standard (src/cmd_extract.rs:260-265)."""

    def __init__(self, patterns: Sequence[bytes], ascii_case_insensitive: bool = False):
        self.patterns = [bytes(p) for p in patterns]
        self._blob, self._off = pack_patterns(self.patterns)
        self._h = lib().mko_ac_build(self._blob.ctypes.data, self._off.ctypes.data, len(self.patterns), int(ascii_case_insensitive))
        if not self._h:
            raise MemoryError("mko_ac_build failed")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().mko_ac_free(self._h)
            self._h = None

    @property
    def handle(self):
        return self._h

    def n_states(self) -> int:
        return lib().mko_ac_n_states(self._h)

    def find_overlapping_iter(self, text: bytes) -> List[Tuple[int, int]]:
        """[(pattern index, start)] in report order."""
        t, tp = _buf(text)
        n = lib().mko_ac_find_overlapping(self._h, tp, len(text), None, None, 0, 0)
        if n == 0:
            return []
        pat = np.zeros(n, dtype=np.uint32)
        st = np.zeros(n, dtype=np.uint64)
        lib().mko_ac_find_overlapping(self._h, tp, len(text), pat.ctypes.data, st.ctypes.data, n, 0)
        return list(zip(pat.tolist(), st.tolist()))

    def is_match(self, text: bytes) -> bool:
        t, tp = _buf(text)
        return lib().mko_ac_find_overlapping(self._h, tp, len(text), None, None, 0, 1) > 0

    def batch_hits(self, seq: np.ndarray, off: np.ndarray, lens: Optional[np.ndarray] = None):
        """All hits of a packed batch as (record, start, pattern) arrays in AC report order."""
        n_rec = len(off) - 1
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        lp = None
        if lens is not None:
            lens = np.ascontiguousarray(lens, dtype=np.uint32)
            lp = lens.ctypes.data
        sp = seq.ctypes.data if seq.size else np.zeros(1, np.uint8).ctypes.data
        n = lib().mko_ac_batch_hits(self._h, sp, off.ctypes.data, lp, n_rec, None, None, None, 0)
        rec = np.zeros(max(n, 1), dtype=np.uint32)
        st = np.zeros(max(n, 1), dtype=np.uint32)
        pat = np.zeros(max(n, 1), dtype=np.uint32)
        lib().mko_ac_batch_hits(self._h, sp, off.ctypes.data, lp, n_rec, rec.ctypes.data, st.ctypes.data, pat.ctypes.data, n)
        return rec[:n], st[:n], pat[:n]

    def scan_batch(self, seq: np.ndarray, off: np.ndarray, n_threads: int = 1, count_all: bool = False):
        """(flags bitmap u64[], records hit, hits) — the CPU baseline loop."""
        n_rec = len(off) - 1
        flags = np.zeros((n_rec + 63) // 64, dtype=np.uint64)
        nh = np.zeros(1, dtype=np.uint64)
        rec = lib().mko_ac_scan_batch(self._h, seq.ctypes.data, off.ctypes.data, n_rec, n_threads, int(count_all),
                                      flags.ctypes.data, nh.ctypes.data)
        return flags, int(rec), int(nh[0])


# ----------------------------------------------------------------------------------------------
# needletail: complement / reverse_complement / canonical
# ----------------------------------------------------------------------------------------------
_COMP = {}
for a, b in ("AT", "CG", "RY", "KM", "BV", "DH", "SS", "WW"):
    _COMP[ord(a)] = ord(b)
    _COMP[ord(b)] = ord(a)
    _COMP[ord(a.lower())] = ord(b.lower())
    _COMP[ord(b.lower())] = ord(a.lower())
_COMP_TABLE = bytes(_COMP.get(i, i) for i in range(256))  # N and every other byte pass through


def reverse_complement(seq: bytes) -> bytes:
    return seq.translate(_COMP_TABLE)[::-1]


def canonical(seq: bytes) -> bytes:
    """lexicographic min(seq, revcomp); ties keep seq"""
    rc = reverse_complement(seq)
    return rc if rc < seq else seq


# ----------------------------------------------------------------------------------------------
# src/helpers.rs
# ----------------------------------------------------------------------------------------------
_RUST_WS = " \t\n\r\x0b\x0c\x85\xa0                　"


def _rust_lines(content: str) -> List[str]:
    """str::lines(): split on '\n', drop one trailing '\r' per line, no trailing empty line."""
    if not content:
        return []
    parts = content.split("\n")
    if parts[-1] == "":
        parts.pop()
    return [p[:-1] if p.endswith("\r") else p for p in parts]


def read_kmers_from_file(path) -> List[str]:
    """src/helpers.rs:139-163"""
    path = Path(path)
    if path.is_dir():
        raise RefError(f"K-mer file path '{path}' is a directory, not a file.")
    try:
        content = path.read_bytes().decode("utf-8")
    except FileNotFoundError:
        raise RefError("File not found.")
    except (OSError, UnicodeDecodeError):
        raise RefError(f"Error reading file: {path}")
    kmers = [ln.strip(_RUST_WS) for ln in _rust_lines(content) if ln and not ln.startswith("#") and not ln.startswith(">")]
    if not kmers:
        raise RefError("No k-mers found in the file.")
    return kmers


def parse_pattern_list(kmer_file, kmer_seq: Optional[Sequence[str]], reverse_complement_: bool, canonical_: bool,
                       lowercase: bool, uppercase: bool) -> List[str]:
    """src/helpers.rs:76-133. Returns the sorted, unique, non-empty list; index == pattern id."""
    if kmer_file is not None:
        try:
            pats = read_kmers_from_file(kmer_file)
        except RefError as e:
            raise RefError(f'Problem reading k-mers from file: "{kmer_file}"\n\nCaused by:\n    {e}')
    else:
        if kmer_seq is None:
            raise RefError("No k-mer sequence provided.")
        pats = list(kmer_seq)
    if lowercase:
        pats = [s.lower() for s in pats]
    elif uppercase:
        pats = [s.upper() for s in pats]
    if reverse_complement_:
        pats = pats + [reverse_complement(p.encode()).decode() for p in pats]
    if canonical_:
        pats = [canonical(p.encode()).decode() for p in pats]
    pats = sorted(set(p for p in pats if p != ""), key=lambda s: s.encode())
    if not pats:
        raise RefError("No k-mers found in file or provided sequence.")
    return pats


def recommend_aho_corasick(patterns: Sequence[str]) -> bool:
    """src/helpers.rs:203-211"""
    return len(patterns) >= 14 or max(len(p.encode()) for p in patterns) > 64


def choose_aho_corasick(patterns, case_insensitive: bool, q_size: Optional[int], aho_corasick: bool) -> bool:
    """src/cmd_extract.rs:165-171 == src/cmd_tag.rs:183-189"""
    if case_insensitive:
        return True
    if q_size is None and not aho_corasick:
        return recommend_aho_corasick(patterns)
    return aho_corasick


def check_log_flag_conflict(out_log, json_log, out_file, suppress_output: bool) -> None:
    """src/helpers.rs:172-200"""
    l_std = out_log is not None and str(out_log) == "STDOUT"
    j_std = json_log is not None and str(json_log) == "STDOUT"
    if l_std and j_std:
        raise RefError("Cannot use both -l/--out-log and -j/--json-log with no arguments (both to stdout). Please specify a file for at least one.")
    if (l_std or j_std) and out_file is None and not suppress_output:
        raise RefError("Cannot write log to stdout when normal output is also stdout. Specify an output file with -o or suppress output with -S.")


def add_suffix_to_file_prefix(path, suffix: str) -> Path:
    """src/helpers.rs:29-43 (suffix goes before the first dot of the file name)"""
    path = Path(path)
    parts = path.name.split(".")
    parts[0] = parts[0] + suffix
    return path.with_name(".".join(parts))


def identify_uncompressed_type(path) -> str:
    """src/helpers.rs:48-68"""
    path = Path(path)
    if path.is_dir():
        raise RefError("The path points to a directory.")
    name = path.name
    ext = _rust_extension(name)
    if ext is None:
        raise RefError("Path has no extension")
    if ext in ("gz", "bz", "bz2", "xz"):
        inner = _rust_extension(name[: -(len(ext) + 1)])
        if inner is None:
            raise RefError("Could not determine uncompressed file type")
        return inner
    return ext


def _rust_extension(name: str) -> Optional[str]:
    """Path::extension(): text after the last dot, None for no dot or a leading-dot-only name."""
    i = name.rfind(".")
    if i <= 0:
        return None
    return name[i + 1:]


def with_extension(path, ext: str) -> Path:
    """PathBuf::with_extension"""
    path = Path(path)
    name = path.name
    i = name.rfind(".")
    stem = name if i <= 0 else name[:i]
    return path.with_name(stem + ("." + ext if ext else ""))


# ----------------------------------------------------------------------------------------------
# needletail record model
# ----------------------------------------------------------------------------------------------
@dataclass
class FastxRecord:
    id: bytes           # header line without '>' / '@'
    seq: bytes          # line breaks removed
    raw_seq: bytes      # as in the file (FASTA: internal line breaks kept, final one dropped)
    qual: Optional[bytes]
    line_ending: bytes

    def num_bases(self) -> int:
        return len(self.seq)

    def write(self) -> bytes:
        le = self.line_ending
        if self.qual is None:
            return b">" + self.id + le + self.raw_seq + le
        return b"@" + self.id + le + self.raw_seq + le + b"+" + le + self.qual + le


def open_maybe_compressed(path) -> bytes:
    data = Path(path).read_bytes()
    if data[:2] == b"\x1f\x8b":
        return gzip.decompress(data)
    if data[:3] == b"BZh":
        import bz2
        return bz2.decompress(data)
    if data[:6] == b"\xfd7zXZ\x00":
        import lzma
        return lzma.decompress(data)
    return data


def parse_fastx(data: bytes) -> List[FastxRecord]:
    recs: List[FastxRecord] = []
    i, n = 0, len(data)
    # needletail skips leading blank lines
    while i < n and data[i:i + 1] in (b"\n", b"\r"):
        i += 1
    if i >= n:
        return recs
    first = data[i:i + 1]
    if first == b">":
        while i < n:
            if data[i:i + 1] != b">":
                raise RefError("Error during FASTQ/A record parsing.")
            j = data.find(b"\n", i)
            if j < 0:
                j = n
            head = data[i + 1:j]
            le = b"\n"
            if head.endswith(b"\r"):
                head, le = head[:-1], b"\r\n"
            k = data.find(b"\n>", j)
            end = n if k < 0 else k
            raw = data[j + 1:end] if j < n else b""
            # drop the final line break of the record
            if raw.endswith(b"\n"):
                raw = raw[:-1]
            if raw.endswith(b"\r"):
                raw = raw[:-1]
            seq = raw.replace(b"\n", b"").replace(b"\r", b"")
            recs.append(FastxRecord(head, seq, raw, None, le))
            i = n if k < 0 else k + 1
    elif first == b"@":
        lines = data[i:].split(b"\n")
        if lines and lines[-1] == b"":
            lines.pop()
        k = 0
        while k < len(lines):
            if lines[k] in (b"", b"\r") and all(x in (b"", b"\r") for x in lines[k:]):
                break
            if k + 3 >= len(lines) + 0 and len(lines) - k < 4:
                raise RefError("Error during FASTQ/A record parsing.")
            h, s, p, q = lines[k:k + 4]
            le = b"\n"
            if h.endswith(b"\r"):
                le = b"\r\n"
                h, s, p, q = (x[:-1] if x.endswith(b"\r") else x for x in (h, s, p, q))
            if not h.startswith(b"@") or not p.startswith(b"+") or len(s) != len(q):
                raise RefError("Error during FASTQ/A record parsing.")
            recs.append(FastxRecord(h[1:], s, s, q, le))
            k += 4
    else:
        raise RefError("Error during FASTQ/A record parsing.")
    return recs


# ----------------------------------------------------------------------------------------------
# src/logger.rs
# ----------------------------------------------------------------------------------------------
def _json_str(s: str) -> str:
    return json.dumps(s, ensure_ascii=False)


def _pretty(value, indent=0) -> str:
    """serde_json::to_string_pretty with sorted keys (serde_json's default map is a BTreeMap)."""
    return json.dumps(value, indent=2, sort_keys=True, ensure_ascii=False)


class TextLog:
    """BufferedLogger (src/logger.rs:11-83): header lines are written through, records buffered."""

    def __init__(self, enabled: bool):
        self.enabled = enabled
        self.parts: List[str] = []

    def write_header(self, s: str):
        if self.enabled:
            self.parts.append(s)

    def log_fields(self, prefix: str, record: bytes, pattern: str, index: int):
        if self.enabled:
            self.parts.append(f"{prefix}\t{record.decode('utf-8')}\t{pattern}\t{index}\n")

    def content(self) -> bytes:
        return "".join(self.parts).encode("utf-8")


class JsonLog:
    """JsonLogger (src/logger.rs:86-191)"""

    def __init__(self):
        self.parts: List[str] = ['{\n  "matching_records": [\n']
        self.first = True

    def log_fields(self, file: str, record: bytes, pattern: str, index: int):
        if not self.first:
            self.parts.append(",\n")
        self.first = False
        obj = {"file": file, "record_id": record.decode("utf-8"), "pattern": pattern, "position": str(index)}
        for line in _pretty(obj).split("\n"):
            self.parts.append("    " + line + "\n")

    def _indented(self, value, indent: int) -> str:
        lines = _pretty(value).split("\n")
        return "\n".join([lines[0]] + [" " * indent + ln for ln in lines[1:]])

    def finalize(self, meta, pattern_hit_counts, summary, paired) -> bytes:
        self.parts.append('  ],\n  "meta_information": ' + self._indented(meta, 2))
        if paired is not None:
            self.parts.append(',\n  "paired_end_reads_statistics": ' + self._indented(paired, 2))
        self.parts.append(',\n  "pattern_hit_counts": ' + self._indented(pattern_hit_counts, 2))
        self.parts.append(',\n  "summary_statistics": ' + self._indented(summary, 2))
        self.parts.append("\n}\n")
        return "".join(self.parts).encode("utf-8")


def _timestamp() -> str:
    import datetime
    now = datetime.datetime.now().astimezone().replace(microsecond=0)
    return now.isoformat()


# ----------------------------------------------------------------------------------------------
# extract  (src/cmd_extract.rs:143-717)
# ----------------------------------------------------------------------------------------------
@dataclass
class CmdExtract:
    in_fastx: str = ""
    in_fastq_2: Optional[str] = None
    kmer_seq: Optional[List[str]] = None
    kmer_file: Optional[str] = None
    out_fastx: Optional[str] = None
    reverse_complement: bool = False
    canonical: bool = False
    out_log: Optional[str] = None      # "STDOUT" when the flag has no value
    json_log: Optional[str] = None
    suppress_output: bool = False
    invert_match: bool = False
    case_insensitive: bool = False
    lowercase: bool = False
    uppercase: bool = False
    q_size: Optional[int] = None
    aho_corasick: bool = False
    argv: List[str] = field(default_factory=lambda: ["merkurio", "extract"])


@dataclass
class RunResult:
    files: dict          # path (str) -> bytes ; "STDOUT" for stdout
    search_algorithm: str = ""
    patterns: List[str] = field(default_factory=list)


def _build_matchers(patterns: List[str], use_ac: bool, case_insensitive: bool, q_size: Optional[int]):
    if use_ac:
        return AhoCorasick([p.encode() for p in patterns], case_insensitive), None
    coll = []
    for p in patterns:
        q = q_size if q_size is not None else tune_q_value(p)
        coll.append((p, BNDMq(p.encode(), q)))
    return None, coll


def _write_out(files: dict, key: str, data: bytes):
    files[key] = files.get(key, b"") + data


def extract_records(args: CmdExtract, write_files: bool = True) -> RunResult:
    check_log_flag_conflict(args.out_log, args.json_log, args.out_fastx, args.suppress_output)
    try:
        patterns = parse_pattern_list(args.kmer_file, args.kmer_seq, args.reverse_complement, args.canonical,
                                      args.lowercase, args.uppercase)
    except RefError as e:
        raise RefError(f"Problem parsing pattern list.\n\nCaused by:\n    {e}")
    use_ac = choose_aho_corasick(patterns, args.case_insensitive, args.q_size, args.aho_corasick)
    res = RunResult(files={}, search_algorithm="Aho-Corasick" if use_ac else "BNDMq", patterns=patterns)

    if Path(args.in_fastx).is_dir():
        raise RefError(f"Record file path '{args.in_fastx}' is a directory, not a file.")
    f1 = Path(args.in_fastx).name
    f2 = ""
    if args.in_fastq_2 is not None:
        if Path(args.in_fastq_2).is_dir():
            raise RefError(f"Second read file path '{args.in_fastq_2}' is a directory, not a file.")
        f2 = Path(args.in_fastq_2).name
    logging_active = args.out_log is not None or args.json_log is not None
    log = TextLog(args.out_log is not None)
    jl = JsonLog() if args.json_log is not None else None
    if logging_active:
        log.write_header("#SeqKatcher extract log\n")
        log.write_header(f"#{_timestamp()}\n")
        log.write_header(f"#Running {PROGRAM} version {VERSION}\n")
        log.write_header(f"#Command line: {' '.join(args.argv)}\n")
        log.write_header("#Searching for {} pattern{} {}\n".format(
            len(patterns), "s" if len(patterns) > 1 else "", "(inverted matching)" if args.invert_match else ""))
        log.write_header("#\n#File\tRecord\tPattern\tPosition (zero-based)\n")

    ac, coll = _build_matchers(patterns, use_ac, args.case_insensitive, args.q_size)

    try:
        recs1 = parse_fastx(open_maybe_compressed(args.in_fastx))
    except (OSError, RefError) as e:
        raise RefError(f"Invalid FASTQ/A input path or file: \"{args.in_fastx}\"")
    nb_records_tot = 0
    nb_bases = 0
    hits_tot = [0, 0]
    rec_hit = [0, 0]
    nb_extracted = 0
    counts = [0] * len(patterns)

    def emit(fname, rec, pat_idx, pos):
        log.log_fields(fname, rec.id, patterns[pat_idx], pos)
        if jl is not None:
            jl.log_fields(fname, rec.id, patterns[pat_idx], pos)

    if args.in_fastq_2 is None:
        out_key = "STDOUT"
        if args.out_fastx is not None:
            out_key = str(with_extension(args.out_fastx, identify_uncompressed_type(args.in_fastx)))
        if not args.suppress_output:
            res.files.setdefault(out_key, b"")
        for rec in recs1:
            found = False
            if logging_active:
                nb_records_tot += 1
                nb_bases += rec.num_bases()
            if ac is not None:
                if not logging_active:
                    found = ac.is_match(rec.seq)
                else:
                    for p, s in ac.find_overlapping_iter(rec.seq):
                        emit(f1, rec, p, s)
                        counts[p] += 1
                        hits_tot[0] += 1
                        found = True
                if found:
                    rec_hit[0] += 1
            elif logging_active:
                for idx, (pat, m) in enumerate(coll):
                    occ = m.find_iter(rec.seq)
                    for o in occ:
                        emit(f1, rec, idx, o)
                        hits_tot[0] += 1
                    if occ:
                        found = True
                        counts[idx] += 1
                if found:
                    rec_hit[0] += 1
            else:
                found = any(m.find_match(rec.seq) for _, m in coll)
            if found != args.invert_match:
                nb_extracted += 1
                if not args.suppress_output:
                    _write_out(res.files, out_key, rec.write())
    else:
        try:
            recs2 = parse_fastx(open_maybe_compressed(args.in_fastq_2))
        except (OSError, RefError):
            raise RefError(f"Invalid second FASTQ input path or file: Some(\"{args.in_fastq_2}\")")
        k1 = k2 = "STDOUT"
        if args.out_fastx is not None:
            base = with_extension(args.out_fastx, identify_uncompressed_type(args.in_fastx))
            k1 = str(add_suffix_to_file_prefix(base, "_1"))
            k2 = str(add_suffix_to_file_prefix(base, "_2"))
        if not args.suppress_output:
            res.files.setdefault(k1, b"")
            res.files.setdefault(k2, b"")
        for i, r1 in enumerate(recs1):
            if i >= len(recs2):
                raise RefError("Error during FASTQ record parsing of second file. Do the two input files contain the same number of records?")
            r2 = recs2[i]
            found = False
            if logging_active:
                nb_records_tot += 2
                nb_bases += r1.num_bases() + r2.num_bases()
            if ac is not None:
                if not logging_active:
                    found = ac.is_match(r1.seq) or ac.is_match(r2.seq)
                else:
                    rh = [0, 0]
                    for fi, (fname, rec) in enumerate(((f1, r1), (f2, r2))):
                        for p, s in ac.find_overlapping_iter(rec.seq):
                            emit(fname, rec, p, s)
                            counts[p] += 1
                            rh[fi] = 1
                            hits_tot[fi] += 1
                            found = True
                    rec_hit[0] += rh[0]
                    rec_hit[1] += rh[1]
            elif logging_active:
                rh = [0, 0]
                for idx, (pat, m) in enumerate(coll):
                    for fi, (fname, rec) in enumerate(((f1, r1), (f2, r2))):
                        occ = m.find_iter(rec.seq)
                        for o in occ:
                            emit(fname, rec, idx, o)
                            hits_tot[fi] += 1
                        if occ:
                            found = True
                            rh[fi] = 1
                            counts[idx] += 1
                rec_hit[0] += rh[0]
                rec_hit[1] += rh[1]
            else:
                found = any(m.find_match(r1.seq) or m.find_match(r2.seq) for _, m in coll)
            if found != args.invert_match:
                nb_extracted += 2
                if not args.suppress_output:
                    _write_out(res.files, k1, r1.write())
                    _write_out(res.files, k2, r2.write())
        if len(recs2) > len(recs1):
            raise RefError("The two input files have a different number of records. Please provide valid paired-end read files.")

    paired = args.in_fastq_2 is not None
    if logging_active:
        found_n = sum(1 for c in counts if c > 0)
        log.write_header("#\n#Number of patterns found: {}/{} ({:.2f} %)\n".format(found_n, len(counts), found_n / len(counts) * 100.0))
        log.write_header("#Pattern\tCount\n")
        for p, c in zip(patterns, counts):
            log.write_header(f"#{p}\t{c}\n")
        log.write_header(f"#\n#Total number of records searched: {nb_records_tot}\n")
        log.write_header(f"#Total number of characters searched: {nb_bases}\n")
        log.write_header(f"#Total number of hits: {hits_tot[0] + hits_tot[1]}\n")
        log.write_header(f"#Number of distinct records with a hit: {rec_hit[0] + rec_hit[1]}\n")
        if paired:
            log.write_header(f"#\n#Total number of hits in file 1: {hits_tot[0]}\n")
            log.write_header(f"#Total number of hits in file 2: {hits_tot[1]}\n")
            log.write_header(f"#Number of distinct records with a hit in file 1: {rec_hit[0]}\n")
            log.write_header(f"#Number of distinct records with a hit in file 2: {rec_hit[1]}\n")
            log.write_header(f"#Total number of extracted records: {nb_extracted}\n")
    if args.out_log is not None:
        _write_out(res.files, str(args.out_log), log.content())
    if jl is not None:
        meta = {
            "program": PROGRAM, "version": VERSION, "timestamp": _timestamp(), "subcommand": "extract",
            "command_line": list(args.argv), "search_algorithm": res.search_algorithm,
            "inverted_matching": args.invert_match, "case_insensitive": args.case_insensitive,
            "input_files": {"kmer_file": str(args.kmer_file) if args.kmer_file is not None else None,
                            "record_file_1": f1, "record_file_2": f2 if paired else None},
        }
        summary = {
            "number_of_patterns_searched": len(patterns),
            "number_of_patterns_found": sum(1 for c in counts if c > 0),
            "number_of_records_searched": nb_records_tot,
            "number_of_characters_searched": nb_bases,
            "number_of_matches": hits_tot[0] + hits_tot[1],
            "number_of_distinct_records_with_a_hit": rec_hit[0] + rec_hit[1],
        }
        pstats = {
            "searching_paired_end_reads": paired,
            "number_of_hits_in_file_1": hits_tot[0],
            "number_of_hits_in_file_2": hits_tot[1] if paired else None,
            "number_of_distinct_records_with_a_hit_in_file_1": rec_hit[0],
            "number_of_distinct_records_with_a_hit_in_file_2": rec_hit[1] if paired else None,
            "number_of_extracted_records": nb_extracted,
        }
        _write_out(res.files, str(args.json_log), jl.finalize(meta, dict(zip(patterns, counts)), summary, pstats))
    if write_files:
        _flush_files(res.files)
    return res


def _flush_files(files: dict):
    for k, v in files.items():
        if k == "STDOUT":
            continue
        Path(k).write_bytes(v)


# ----------------------------------------------------------------------------------------------
# SAM / BAM record model (crate bam 0.1.4)
# ----------------------------------------------------------------------------------------------
NIBBLE_CHARS = b"=ACMGRSVTWYHKDBN"
_CIGAR_OPS = "MIDNSHP=X"


@dataclass
class SamRecord:
    name: bytes
    seq: bytes        # decoded, upper-case as the bam crate yields it ("" for '*')
    line: bytes       # SAM text line without line break


def _sam_seq_through_nibbles(seq: bytes) -> bytes:
    """SAM text -> 4-bit codes -> text, as `record.sequence().to_vec()` sees it (src/cmd_tag.rs:395).
    Characters outside the 16 codes are stored as N; letters are case-insensitive."""
    if seq == b"*":
        return b""
    up = seq.upper()
    return bytes(c if c in NIBBLE_CHARS else ord("N") for c in up)


def parse_sam(data: bytes) -> Tuple[List[bytes], List[SamRecord]]:
    header, recs = [], []
    for ln in data.split(b"\n"):
        if ln.endswith(b"\r"):
            ln = ln[:-1]
        if not ln:
            continue
        if ln.startswith(b"@"):
            header.append(ln)
            continue
        f = ln.split(b"\t")
        if len(f) < 11:
            raise RefError("Error during SAM record parsing: truncated record")
        recs.append(SamRecord(f[0], _sam_seq_through_nibbles(f[9]), ln))
    return header, recs


def _bam_aux_to_sam(buf: bytes) -> List[bytes]:
    out, i = [], 0
    while i < len(buf):
        tag, typ = buf[i:i + 2], buf[i + 2:i + 3]
        i += 3
        if typ == b"A":
            out.append(tag + b":A:" + buf[i:i + 1]); i += 1
        elif typ in b"cCsSiI":
            fmt = {b"c": "<b", b"C": "<B", b"s": "<h", b"S": "<H", b"i": "<i", b"I": "<I"}[typ]
            sz = struct.calcsize(fmt)
            out.append(tag + b":i:" + str(struct.unpack_from(fmt, buf, i)[0]).encode()); i += sz
        elif typ == b"f":
            v = struct.unpack_from("<f", buf, i)[0]; i += 4
            out.append(tag + b":f:" + str(np.float32(v)).encode())  # shortest float32 form (unpinned by the reference)
        elif typ in b"ZH":
            j = buf.index(b"\0", i)
            out.append(tag + b":" + typ + b":" + buf[i:j]); i = j + 1
        elif typ == b"B":
            sub = buf[i:i + 1]; n = struct.unpack_from("<I", buf, i + 1)[0]; i += 5
            fmt = {b"c": "b", b"C": "B", b"s": "h", b"S": "H", b"i": "i", b"I": "I", b"f": "f"}[sub]
            vals = struct.unpack_from("<" + fmt * n, buf, i); i += struct.calcsize("<" + fmt * n)
            out.append(tag + b":B:" + sub + b"".join(b"," + str(v).encode() for v in vals))
        else:
            raise RefError("Error during BAM record parsing: bad tag type")
    return out


def parse_bam(data: bytes) -> Tuple[List[bytes], List[SamRecord]]:
    raw = gzip.decompress(data)
    if raw[:4] != b"BAM\1":
        raise RefError("Error reading BAM file: bad magic")
    l_text = struct.unpack_from("<i", raw, 4)[0]
    text = raw[8:8 + l_text].rstrip(b"\0")
    header = [ln for ln in text.split(b"\n") if ln]
    p = 8 + l_text
    n_ref = struct.unpack_from("<i", raw, p)[0]; p += 4
    refs = []
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", raw, p)[0]; p += 4
        refs.append(raw[p:p + l_name - 1]); p += l_name + 4
    recs = []
    while p < len(raw):
        bs = struct.unpack_from("<i", raw, p)[0]; p += 4
        b = raw[p:p + bs]; p += bs
        ref_id, pos, l_rn, mapq, _bin, n_cig, flag, l_seq, nref, npos, tlen = struct.unpack_from("<iiBBHHHiiii", b, 0)
        q = 32
        name = b[q:q + l_rn - 1]; q += l_rn
        cig = struct.unpack_from("<" + "I" * n_cig, b, q); q += 4 * n_cig
        packed = b[q:q + (l_seq + 1) // 2]; q += (l_seq + 1) // 2
        qual = b[q:q + l_seq]; q += l_seq
        seq = bytes(NIBBLE_CHARS[(packed[i >> 1] >> (4 if i % 2 == 0 else 0)) & 0xF] for i in range(l_seq))
        cigar = "".join(f"{c >> 4}{_CIGAR_OPS[c & 0xF]}" for c in cig) or "*"
        rname = refs[ref_id] if ref_id >= 0 else b"*"
        rnext = b"*" if nref < 0 else (b"=" if nref == ref_id else refs[nref])
        qtxt = b"*" if (l_seq == 0 or qual[:1] == b"\xff") else bytes(c + 33 for c in qual)
        fields = [name, str(flag).encode(), rname, str(pos + 1).encode(), str(mapq).encode(), cigar.encode(), rnext,
                  str(npos + 1).encode(), str(tlen).encode(), seq if l_seq else b"*", qtxt] + _bam_aux_to_sam(b[q:])
        recs.append(SamRecord(name, seq, b"\t".join(fields)))
    return header, recs


# ----------------------------------------------------------------------------------------------
# tag  (src/cmd_tag.rs:155-689)
# ----------------------------------------------------------------------------------------------
@dataclass
class CmdTag:
    in_file: str = ""
    out_file: Optional[str] = None
    kmer_seq: Optional[List[str]] = None
    kmer_file: Optional[str] = None
    reverse_complement: bool = False
    canonical: bool = False
    out_log: Optional[str] = None
    json_log: Optional[str] = None
    suppress_output: bool = False
    filter_matching: bool = False
    invert_match: bool = False
    case_insensitive: bool = False
    lowercase: bool = False
    uppercase: bool = False
    q_size: Optional[int] = None
    aho_corasick: bool = False
    tag: str = "km"
    threads: int = 1
    argv: List[str] = field(default_factory=lambda: ["merkurio", "tag"])


def _existing_tag_value(line: bytes, tag: bytes):
    """('missing' | 'string' | 'other', value) of an optional field of a SAM line"""
    for f in line.split(b"\t")[11:]:
        if f[:2] == tag and f[2:3] == b":":
            if f[3:5] == b"Z:":
                return "string", f[5:]
            return "other", b""
    return "missing", b""


def tag_records(args: CmdTag, write_files: bool = True) -> RunResult:
    check_log_flag_conflict(args.out_log, args.json_log, args.out_file, args.suppress_output)
    if Path(args.in_file).is_dir():
        raise RefError(f"Record file path '{args.in_file}' is a directory, not a file.")
    fname = Path(args.in_file).name
    try:
        patterns = parse_pattern_list(args.kmer_file, args.kmer_seq, args.reverse_complement, args.canonical,
                                      args.lowercase, args.uppercase)
    except RefError as e:
        raise RefError(f"Problem parsing pattern list.\n\nCaused by:\n    {e}")
    use_ac = choose_aho_corasick(patterns, args.case_insensitive, args.q_size, args.aho_corasick)
    res = RunResult(files={}, search_algorithm="Aho-Corasick" if use_ac else "BNDMq", patterns=patterns)
    logging_active = args.out_log is not None or args.json_log is not None
    if args.threads < 1:
        raise RefError("Number of threads must be at least 1.")
    if len(args.tag.encode()) != 2:
        raise RefError("Tag must be exactly two characters long.")
    tagb = args.tag.encode()
    ac, coll = _build_matchers(patterns, use_ac, args.case_insensitive, args.q_size)

    in_ext = _rust_extension(Path(args.in_file).name)
    if in_ext is None:
        raise RefError(f"Could not detect the file extension: \"{args.in_file}\"")
    if args.out_file is not None:
        out_ext = _rust_extension(Path(args.out_file).name) or in_ext
    else:
        out_ext = "STDOUT"

    log = TextLog(args.out_log is not None)
    jl = JsonLog() if args.json_log is not None else None
    if logging_active:
        log.write_header("#SeqKatcher tag log\n")
        log.write_header(f"#{_timestamp()}\n")
        log.write_header(f"#Running {PROGRAM} version {VERSION}\n")
        log.write_header(f"#Command line: {' '.join(args.argv)}\n")
        log.write_header(f"#Tag used for labeling records: {args.tag}\n")
        log.write_header("#Searching for {} pattern{} {}\n".format(
            len(patterns), "s" if len(patterns) > 1 else "", "(inverted matching)" if args.invert_match else ""))
        log.write_header("#\n#File\tRecord\tPattern\tPosition (zero-based)\n")

    if in_ext == "bam":
        header, recs = parse_bam(Path(args.in_file).read_bytes())
    elif in_ext == "sam":
        header, recs = parse_sam(Path(args.in_file).read_bytes())
    else:
        raise RefError("Input file must be a BAM or SAM file.")
    pg = f"@PG\tID:{PROGRAM}\tPN:{PROGRAM}\tCL:{' '.join(args.argv)}\tVN:{VERSION}".encode()
    header = header + [pg]
    if args.suppress_output:
        header = []
    if out_ext not in ("bam", "sam", "STDOUT"):
        raise RefError("Could not create writer.\n\nCaused by:\n    Output file must be a BAM or SAM file.")
    out_key = "STDOUT" if out_ext == "STDOUT" else str(with_extension(args.out_file, out_ext))
    out_lines: List[bytes] = list(header)

    nb_records_tot = nb_bases = nb_hits = nb_rec_hit = 0
    counts = [0] * len(patterns)
    for rec in recs:
        found: List[str] = []
        if ac is not None:
            for p, s in ac.find_overlapping_iter(rec.seq):
                found.append(patterns[p])
                if logging_active:
                    nb_hits += 1
                    counts[p] += 1
                    log.log_fields(fname, rec.name, patterns[p], s)
                    if jl is not None:
                        jl.log_fields(fname, rec.name, patterns[p], s)
        elif logging_active:
            for idx, (pat, m) in enumerate(coll):
                occ = m.find_iter(rec.seq)
                for o in occ:
                    log.log_fields(fname, rec.name, pat, o)
                    if jl is not None:
                        jl.log_fields(fname, rec.name, pat, o)
                    nb_hits += 1
                if occ:
                    found.append(pat)
                    counts[idx] += 1
        else:
            for pat, m in coll:
                if m.find_match(rec.seq):
                    found.append(pat)
        if logging_active:
            nb_records_tot += 1
            nb_bases += len(rec.seq)
            if found:
                nb_rec_hit += 1
        if args.filter_matching:
            keep = bool(found)
        elif args.invert_match:
            keep = not found
        else:
            keep = True
        if not keep:
            continue
        kind, val = _existing_tag_value(rec.line, tagb)
        if kind == "other":
            raise RefError("Invalid tag value format. Expected string value.")
        if kind == "string" and val != b"":
            found.extend(val.decode("utf-8").split(","))
        found = sorted(set(found), key=lambda s: s.encode())
        if not args.suppress_output:
            out_lines.append(rec.line + b"\t" + tagb + b":Z:" + ",".join(found).encode())

    if not args.suppress_output:
        if out_ext == "bam":
            res.files[out_key + "#as-sam"] = b"".join(ln + b"\n" for ln in out_lines)  # BGZF bytes are not modelled
        else:
            res.files[out_key] = b"".join(ln + b"\n" for ln in out_lines)

    if logging_active:
        found_n = sum(1 for c in counts if c > 0)
        log.write_header("#\n#Number of patterns found: {}/{} ({:.2f} %)\n".format(found_n, len(counts), found_n / len(counts) * 100.0))
        log.write_header("#Pattern\tCount\n")
        for p, c in zip(patterns, counts):
            log.write_header(f"#{p}\t{c}\n")
        log.write_header(f"#\n#Total number of records searched: {nb_records_tot}\n")
        log.write_header(f"#Total number of characters searched: {nb_bases}\n")
        log.write_header(f"#Total number of hits: {nb_hits}\n")
        log.write_header(f"#Number of distinct records with a hit: {nb_rec_hit}\n")
    if args.out_log is not None:
        _write_out(res.files, str(args.out_log), log.content())
    if jl is not None:
        meta = {
            "program": PROGRAM, "version": VERSION, "timestamp": _timestamp(), "subcommand": "tag",
            "command_line": list(args.argv), "search_algorithm": res.search_algorithm,
            "inverted_matching": args.invert_match, "case_insensitive": args.case_insensitive,
            "input_files": {"kmer_file": str(args.kmer_file) if args.kmer_file is not None else None, "record_file_1": fname},
            "tag": args.tag,
        }
        summary = {
            "number_of_patterns_searched": len(patterns),
            "number_of_patterns_found": sum(1 for c in counts if c > 0),
            "number_of_records_searched": nb_records_tot,
            "number_of_characters_searched": nb_bases,
            "number_of_matches": nb_hits,
            "number_of_distinct_records_with_a_hit": nb_rec_hit,
        }
        _write_out(res.files, str(args.json_log), jl.finalize(meta, dict(zip(patterns, counts)), summary, None))
    if write_files:
        _flush_files({k: v for k, v in res.files.items() if not k.endswith("#as-sam")})
    return res
