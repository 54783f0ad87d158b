"""In-tree builds: the CUDA matching library (sm_100a), the synthetic-data helper, the C++ host
binary and — test infrastructure only — the CPU oracle under oracle/.

Everything is compiled with explicit nvcc / g++ command lines so that the built .so files sit in
the source tree and travel to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "merkurio_b200"
CSRC = PKG / "csrc"
HOST = PKG / "host"
LIBDIR = PKG / "lib"
ORACLE = ROOT / "oracle"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unknown-pragmas",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print("+", " ".join(map(str, cmd)), file=sys.stderr)
    subprocess.run(list(map(str, cmd)), check=True)


def cuda_sources():
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [ROOT / "include" / "merkurio_cuda.h"]


def build_cuda(force: bool = False, verbose: bool = False, ptxas_v: bool = False, debug_checks: bool = False) -> Path:
    """libmerkurio_cuda.so: kernels + C ABI (include/merkurio_cuda.h). debug_checks: libmerkurio_cuda_dbg.so with
    device-side asserts on every computed index (-DMK_DEBUG_CHECKS); load it with MK_CUDA_LIB=<path> (capi.py)."""
    LIBDIR.mkdir(exist_ok=True)
    out = LIBDIR / ("libmerkurio_cuda_dbg.so" if debug_checks else "libmerkurio_cuda.so")
    if force or _stale(out, cuda_sources()):
        cmd = [_nvcc(), *NVCC_FLAGS, "-shared", "-I", ROOT / "include", "-o", out, CSRC / "mk_engine.cu"]
        if debug_checks:
            cmd += ["-DMK_DEBUG_CHECKS"]
        if ptxas_v:
            cmd += ["-Xptxas", "-v"]
        if os.environ.get("MK_TUNE_BUILD"):  # every launch shape scripts/tune_scan.py sweeps
            cmd += ["-DMK_TUNE_BUILD"]
        _run(cmd, verbose)
    return out


def build_synth(force: bool = False, verbose: bool = False) -> Path:
    """libmerkurio_synth.so: synthetic read generator (bench / test support, not the product)."""
    LIBDIR.mkdir(exist_ok=True)
    out = LIBDIR / "libmerkurio_synth.so"
    src = PKG / "synth" / "mk_synth.cu"
    if not src.exists():
        return out
    if force or _stale(out, [src, PKG / "synth" / "mk_synth.h"]):
        _run([_nvcc(), *NVCC_FLAGS, "-shared", "-o", out, src], verbose)
    return out


def build_host(force: bool = False, verbose: bool = False) -> Path:
    """merkurio: the C++ host (CLI with the reference's flags) linked against the C ABI."""
    LIBDIR.mkdir(exist_ok=True)
    out = LIBDIR / "merkurio"
    srcs = sorted(HOST.glob("*.cpp"))
    if not srcs:
        return out
    deps = srcs + sorted(HOST.glob("*.h")) + [ROOT / "include" / "merkurio_cuda.h", build_cuda(force, verbose)]
    if force or _stale(out, deps):
        _run(["g++", "-O2", "-std=c++17", "-Wall", "-pthread", "-I", ROOT / "include", "-o", out, *srcs,
              "-L", LIBDIR, "-lmerkurio_cuda", "-Wl,-rpath,$ORIGIN", "-lz", "-ldl"], verbose)
    return out


def build_io(force: bool = False, verbose: bool = False) -> Path:
    """libmerkurio_io.so: the host's input streams behind a C ABI (include/merkurio_io.h); no CUDA."""
    LIBDIR.mkdir(exist_ok=True)
    out = LIBDIR / "libmerkurio_io.so"
    srcs = [HOST / n for n in ("io_capi.cpp", "codecs.cpp", "inflate.cpp", "pgzip.cpp", "bgzf.cpp", "io.cpp")]
    if force or _stale(out, srcs + sorted(HOST.glob("*.h")) + [ROOT / "include" / "merkurio_io.h"]):
        _run(["g++", "-O2", "-std=c++17", "-Wall", "-fPIC", "-shared", "-pthread", "-I", ROOT / "include", "-o", out, *srcs,
              "-lz", "-ldl", "-Wl,--no-undefined"], verbose)
    return out


def build_oracle(force: bool = False, verbose: bool = False) -> Path:
    """oracle/_build/libmk_oracle.so: CPU restatement of the reference matchers (checker only)."""
    outdir = ORACLE / "_build"
    outdir.mkdir(exist_ok=True)
    out = outdir / "libmk_oracle.so"
    src = ORACLE / "mk_oracle.c"
    if not src.exists():
        return out
    if force or _stale(out, [src]):
        _run(["gcc", "-O3", "-march=x86-64-v2", "-std=c11", "-Wall", "-fPIC", "-shared", "-pthread", "-o", out, src], verbose)
    return out


def build_all(force: bool = False, verbose: bool = False):
    return {
        "cuda": build_cuda(force, verbose),
        "synth": build_synth(force, verbose),
        "host": build_host(force, verbose),
        "io": build_io(force, verbose),
        "oracle": build_oracle(force, verbose),
    }


if __name__ == "__main__":
    for k, v in build_all(force="--force" in sys.argv, verbose=True).items():
        print(f"{k}: {v}")
