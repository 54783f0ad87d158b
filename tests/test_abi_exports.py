"""CPU tests of the boundary: libmerkurio_cuda.so loads without a GPU and exports exactly the entry
points include/merkurio_cuda.h declares; without a device it fails loudly (no CPU path)."""
import ctypes
import os
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def built():
    from merkurio_b200.build import build_cuda
    return build_cuda()


def declared_functions():
    text = (ROOT / "include" / "merkurio_cuda.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mk_[a-z_]+)\s*\(", text)))


def test_header_declares_the_boundary():
    names = declared_functions()
    for n in ("mk_engine_create", "mk_engine_destroy", "mk_slot_buffers", "mk_scan_submit", "mk_scan_wait",
              "mk_scan_host", "mk_scan_device", "mk_last_error"):
        assert n in names


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(str(built))
    for n in declared_functions():
        assert hasattr(lib, n), f"{n} declared in include/merkurio_cuda.h but not exported"
    from merkurio_b200 import capi
    assert sorted(capi.EXPORTS) == declared_functions()


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from merkurio_b200 import capi
    with pytest.raises(capi.MkError) as ei:
        capi.Engine([b"ACGT"])
    assert ei.value.code == -4 and "no CPU path" in ei.value.message


def test_product_does_not_touch_the_oracle():
    # only tests/, __graft_entry__.smoke() and bench.py may reference oracle/
    for p in list((ROOT / "merkurio_b200").rglob("*")) + list((ROOT / "include").rglob("*")):
        if p.is_file() and p.suffix in (".py", ".cu", ".cuh", ".h", ".cpp", ".c"):
            txt = p.read_text(errors="replace")
            if p.name == "build.py":
                continue  # builds the checker, does not use it
            assert "oracle" not in txt.lower() or "mk_oracle" not in txt, p
            assert "refmodel" not in txt, p


def test_table_builder_output_is_unchanged(tmp_path):
    """mk::build_tables is host code that only the GPU suite exercises end to end; its output for seeded
    query lists (perm / window / ordered / dual-key layouts, -I, queries with N, both encodings) must stay
    byte-identical to the digests recorded when that suite last ran against it."""
    import subprocess
    exe = tmp_path / "table_digest"
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I", str(ROOT / "merkurio_b200" / "csrc"), "-I", str(ROOT / "include"), "-o", str(exe),
                    str(ROOT / "tests" / "stub" / "table_digest.cpp")], check=True)
    env = {k: v for k, v in os.environ.items() if not k.startswith("MK_")}
    got = subprocess.run([str(exe), "9"], check=True, capture_output=True, env=env).stdout.decode().splitlines()
    want = (ROOT / "tests" / "golden" / "table_digests.txt").read_text().splitlines()
    assert len(got) == len(want) == 18
    for g, w in zip(got, want):
        assert g == w


def test_struct_layout_matches_the_binding_tables(tmp_path):
    """The sizes and offsets INTEGRATION.md lists for the #[repr(C)] structs of a Rust binding, checked from a C
    translation unit (the header's own _Static_asserts fire at compile time; the numbers are printed and compared
    with the documented table as well)."""
    import re
    import subprocess
    src = tmp_path / "layout.c"
    fields = {
        "mk_patterns": ["off", "n"], "mk_config": ["n_slots", "max_batch_records", "max_batch_bytes", "hit_capacity"],
        "mk_hit": ["start", "pattern", "len"],
        "mk_result": ["n_records", "hits", "n_hits", "bases_scanned", "device_ns", "scan_ns", "verify_ns", "n_candidates", "n_rescans", "d_record_flags", "d_hits"],
        "mk_engine_info": ["seed_q", "seed_d", "n_seeds", "filter_log2_bits", "filter_hashes", "filter_bytes", "filter_in_smem", "table_bytes", "sm_count", "features"],
    }
    body = "".join(f'printf("{s} %zu", sizeof({s}));' + "".join(f'printf(" {f} %zu", offsetof({s}, {f}));' for f in fs) + 'printf("\\n");' for s, fs in fields.items())
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "merkurio_cuda.h"\nint main(void){' + body + "return 0;}\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I", str(ROOT / "include"), "-o", str(exe), str(src)], check=True)
    got = {ln.split()[0]: ln.split()[1:] for ln in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines()}
    doc = (ROOT / "INTEGRATION.md").read_text()
    for s, fs in fields.items():
        row = next(ln for ln in doc.splitlines() if ln.startswith(f"| `{s}` |"))
        cells = [c.strip() for c in row.strip("|").split("|")]
        assert got[s][0] == cells[1], (s, got[s][0], cells[1])
        documented = dict(re.findall(r"`(\w+)` (\d+)", cells[2]))
        for f, off in zip(got[s][1::2], got[s][2::2]):
            assert documented[f] == off, (s, f, off, documented[f])


def test_exports_cover_the_header():
    """Every function the header declares is exported by the built library (and listed in capi.EXPORTS)."""
    import re
    import subprocess
    from merkurio_b200 import capi
    hdr = (ROOT / "include" / "merkurio_cuda.h").read_text()
    declared = set(re.findall(r"\b(mk_[a-z_0-9]+)\s*\(", hdr)) - {"mk_hit"}
    syms = subprocess.run(["nm", "-D", "--defined-only", str(capi.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (mk_[a-z_0-9]+)", syms))
    assert declared <= exported, declared - exported
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)


# ------------------------------------------------------------------------------------------------
# include/merkurio_io.h: the host's input streams behind a C ABI (libmerkurio_io.so, no CUDA)
def test_io_library_exports_its_header_and_reads_like_gzip(tmp_path):
    import gzip
    import random
    from merkurio_b200.build import build_io
    text = (ROOT / "include" / "merkurio_io.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = sorted(set(re.findall(r"\b(mk_input_[a-z_]+)\s*\(", text)))
    assert declared == ["mk_input_close", "mk_input_error", "mk_input_open", "mk_input_read"]
    lib = ctypes.CDLL(str(build_io()))
    for n in declared:
        assert hasattr(lib, n), n
    lib.mk_input_open.restype = ctypes.c_void_p
    lib.mk_input_open.argtypes = [ctypes.c_char_p]
    lib.mk_input_read.restype = ctypes.c_longlong
    lib.mk_input_read.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_ulonglong]
    lib.mk_input_error.restype = ctypes.c_char_p
    lib.mk_input_error.argtypes = [ctypes.c_void_p]
    lib.mk_input_close.argtypes = [ctypes.c_void_p]

    def slurp(path, piece=1 << 16):
        h = lib.mk_input_open(str(path).encode())
        assert h, lib.mk_input_error(None)
        buf = ctypes.create_string_buffer(piece)
        out = bytearray()
        while True:
            n = lib.mk_input_read(h, buf, piece)
            if n <= 0:
                break
            out += buf.raw[:n]
        err = lib.mk_input_error(h).decode()
        lib.mk_input_close(h)
        return n, bytes(out), err

    rng = random.Random(2)
    raw = "".join("@r%d\n%s\n+\n%s\n" % (i, "".join(rng.choices("ACGT", k=120)), "".join(rng.choices("FF:,#", k=120))) for i in range(40000)).encode()
    plain, zipped = tmp_path / "r.fastq", tmp_path / "r.fastq.gz"
    plain.write_bytes(raw)
    zipped.write_bytes(gzip.compress(raw, 6))
    assert slurp(plain) == (0, raw, "")
    assert slurp(zipped, 12345) == (0, raw, "")
    os.environ["MERKURIO_GZIP_PIECE_KB"] = "64"  # the same file on several threads
    try:
        assert slurp(zipped) == (0, raw, "")
    finally:
        del os.environ["MERKURIO_GZIP_PIECE_KB"]
    cut = tmp_path / "cut.fastq.gz"
    cut.write_bytes(zipped.read_bytes()[:100000])
    rc, out, err = slurp(cut)
    assert rc == -1 and raw.startswith(out) and len(out) > 0 and "truncated gzip stream" in err
    assert not lib.mk_input_open(str(tmp_path / "missing.fq").encode())
    assert b"No such file" in lib.mk_input_error(None)
