"""Access to the bundled reference vectors (tests/golden/reference_vectors.json.xz) and the
comparators the reference's own end-to-end tests use (src/cmd_extract.rs:730-881,
src/cmd_tag.rs:840-1006): text logs are compared from line 5 on, JSON logs on every section except
the volatile meta fields, record files byte for byte (SAM: every line except @PG)."""
import base64
import json
import lzma
from pathlib import Path

BUNDLE = Path(__file__).resolve().parent / "golden" / "reference_vectors.json.xz"
_cache = None


def vectors() -> dict:
    global _cache
    if _cache is None:
        blob = json.loads(lzma.decompress(BUNDLE.read_bytes()))
        _cache = {k: base64.b64decode(v) for k, v in blob["files"].items()}
    return _cache


def materialize(dest: Path) -> Path:
    for rel, data in vectors().items():
        p = dest / rel
        p.parent.mkdir(parents=True, exist_ok=True)
        p.write_bytes(data)
    return dest


def log_body(data: bytes):
    """Lines after the 4 volatile metadata lines (src/cmd_extract.rs:746-748)."""
    return data.decode().split("\n")[4:]


def assert_log_equal(actual: bytes, expected: bytes):
    assert log_body(actual) == log_body(expected)


STABLE_META = ("search_algorithm", "inverted_matching", "case_insensitive", "tag", "subcommand", "program")


def json_stable(data: bytes) -> dict:
    d = json.loads(data)
    meta = d.get("meta_information", {})
    d["meta_information"] = {k: meta[k] for k in STABLE_META if k in meta}
    return d


def assert_json_equal(actual: bytes, expected: bytes):
    assert json_stable(actual) == json_stable(expected)


def json_layout(data: bytes):
    """Byte layout of a JSON log with the volatile meta block cut out (checks indentation, the
    ',' separator lines, key order and section order exactly)."""
    text = data.decode()
    a = text.index('  "meta_information": {')
    b = text.index("\n  },\n", a)
    return text[:a], text[b:]


def sam_lines(data: bytes):
    return [ln for ln in data.decode().split("\n") if ln and not ln.startswith("@PG")]


def assert_sam_equal(actual: bytes, expected: bytes):
    assert sam_lines(actual) == sam_lines(expected)
