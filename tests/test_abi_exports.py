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
