"""Host-only throughput of the ingest pipelines (reader -> indexer -> packer -> driver -> writer) without a GPU:
the host binary runs against the test double of the C ABI (tests/stub/stub_engine.c, preloaded), which answers
every batch at once, so what is timed is exactly the host work the GPU path has to keep up with. These are the
"build box, test double" figures of DESIGN.md section 6; the real file -> file numbers come from scripts/bench_cli.py.

    python scripts/bench_host_stub.py [--reads 4000000] [--fasta-mbp 900] [--sam-records 2000000]
"""
import argparse
import os
import re
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from merkurio_b200.build import build_host

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=4_000_000)
ap.add_argument("--fasta-mbp", type=int, default=900)
ap.add_argument("--sam-records", type=int, default=2_000_000)
ap.add_argument("--repeat", type=int, default=3)
ap.add_argument("--startup-ms", type=int, default=0,
                help="the test double takes this long to create an engine (MK_STUB_STARTUP_MS), as a CUDA context does: the readers run "
                     "ahead meanwhile and 'pipeline' then shows the packer / driver stages alone")
args = ap.parse_args()
exe = str(build_host())
L = 150
rng = np.random.default_rng(1)
acgt = np.frombuffer(b"ACGT", dtype=np.uint8)

with tempfile.TemporaryDirectory(prefix="mk_stub_") as tmp:
    tmp = Path(tmp)
    stub = tmp / "libstub_engine.so"
    subprocess.run(["gcc", "-O1", "-shared", "-fPIC", "-I", str(ROOT / "include"), "-o", str(stub), str(ROOT / "tests" / "stub" / "stub_engine.c")], check=True)
    (tmp / "q.txt").write_text("\n".join("".join(rng.choice(list("ACGT"), size=31)) for _ in range(1000)) + "\n")

    def fastq(n):
        names = np.char.add("@read", np.arange(n).astype(str)).astype("S")
        w = names.dtype.itemsize
        nm = np.frombuffer(np.char.ljust(names, w, b"x").tobytes(), dtype=np.uint8).reshape(n, w)
        rec = np.empty((n, w + 1 + L + 3 + L + 1), dtype=np.uint8)
        rec[:, :w] = nm
        rec[:, w] = 10
        rec[:, w + 1:w + 1 + L] = rng.choice(acgt, size=n * L).reshape(n, L)
        rec[:, w + 1 + L] = 10
        rec[:, w + 2 + L] = ord("+")
        rec[:, w + 3 + L] = 10
        rec[:, w + 4 + L:w + 4 + 2 * L] = ord("I")
        rec[:, -1] = 10
        return rec.tobytes()

    def fasta(mbp):
        out = bytearray()
        left, i = mbp * 1_000_000, 0
        while left > 0:
            n = min(left, 250_000_020) // 60 * 60
            if n == 0:
                break
            rows = np.empty((n // 60, 61), dtype=np.uint8)
            rows[:, :60] = rng.choice(acgt, size=n).reshape(-1, 60)
            rows[:, 60] = 10
            out += b">chr%d test\n" % i + rows.tobytes()
            left -= n
            i += 1
        return bytes(out)

    def sam(n):
        seq = rng.choice(acgt, size=n * L).reshape(n, L)
        parts = [b"@HD\tVN:1.6\tSO:unsorted\n@SQ\tSN:chr1\tLN:248956422\n"]
        q = b"I" * L
        for i in range(n):
            parts.append(b"read%d\t0\tchr1\t%d\t60\t150M\t*\t0\t0\t%s\t%s\tNM:i:0\n" % (i, 1000 + i, seq[i].tobytes(), q))
        return b"".join(parts)

    def run(label, units, unit, argv):
        best = None
        for _ in range(args.repeat):
            t0 = time.perf_counter()
            r = subprocess.run([exe, *argv], env=dict(os.environ, LD_PRELOAD=str(stub), MERKURIO_TIMING="1", MK_STUB_STARTUP_MS=str(args.startup_ms)), capture_output=True, text=True)
            dt = time.perf_counter() - t0
            assert r.returncode == 0, r.stderr
            m = re.search(r"pipeline ([0-9.]+) s", r.stderr)
            pipe = float(m.group(1)) if m else dt
            if best is None or pipe < best[0]:
                best = (pipe, dt, [ln for ln in r.stderr.splitlines() if "reader" in ln or "engine setup" in ln])
        print(f"{label}: pipeline {best[0]:.3f} s = {units / best[0] / 1e6:.1f} M {unit}/s (process {best[1]:.3f} s)")
        for ln in best[2]:
            print("   ", ln)

    (tmp / "r.fastq").write_bytes(fastq(args.reads))
    run(f"FASTQ extract, {args.reads} reads", args.reads, "reads", ["extract", "-i", str(tmp / "r.fastq"), "-f", str(tmp / "q.txt"), "-r", "-o", str(tmp / "o.fastq")])
    run(f"FASTQ extract -v (every read written), {args.reads} reads", args.reads, "reads",
        ["extract", "-i", str(tmp / "r.fastq"), "-f", str(tmp / "q.txt"), "-r", "-v", "-o", str(tmp / "o.fastq")])
    (tmp / "r.fastq").unlink()
    (tmp / "g.fa").write_bytes(fasta(args.fasta_mbp))
    run(f"FASTA extract, {args.fasta_mbp} Mbp", args.fasta_mbp * 1e6, "bases", ["extract", "-i", str(tmp / "g.fa"), "-f", str(tmp / "q.txt"), "-r", "-o", str(tmp / "o.fa")])
    (tmp / "g.fa").unlink()
    (tmp / "a.sam").write_bytes(sam(args.sam_records))
    run(f"SAM tag -m, {args.sam_records} records", args.sam_records, "records", ["tag", "-i", str(tmp / "a.sam"), "-f", str(tmp / "q.txt"), "-m", "-o", str(tmp / "o.sam")])
