"""Scan throughput by pattern length: cfg2-shaped reads (ASCII, FLAG mode), n_queries k-mers + reverse complements for
each k of --ks; prints the scan kernel, its CUDA-event time and GB/s, and checks the flag bitmap of a sample against the
oracle. The numbers of DESIGN.md section 4 for k < 31.

    python scripts/bench_k.py --ks 8,10,12,13,14,16,18,21,27,31 [--reads 20000000] [--queries 1000]
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from merkurio_b200 import capi, patterns as pt
from merkurio_b200.synth import Synth
from oracle import refmodel as rm

ap = argparse.ArgumentParser()
ap.add_argument("--ks", default="8,10,12,13,14,16,18,21,27,31")
ap.add_argument("--reads", type=int, default=20_000_000)
ap.add_argument("--queries", type=int, default=1000)
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--check-reads", type=int, default=100_000)
ap.add_argument("--out", default=None)
args = ap.parse_args()
L = 150
n = args.reads
rows = []
for k in map(int, args.ks.split(",")):
    syn = Synth(0x5EED0002, n, L, k, args.queries)
    pats = pt.parse_pattern_list(syn.query_list(), reverse_complement_=True)
    d_seq = torch.empty(n * L + 64, dtype=torch.uint8, device="cuda")
    d_off = torch.empty(n + 1, dtype=torch.int64, device="cuda")
    d_q = torch.from_numpy(syn.queries).cuda()
    syn.device_reads(d_q.data_ptr(), 0, n, d_seq.data_ptr(), d_off.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    d_seq[n * L:] = 0
    with capi.Engine(pats, n_slots=0) as e:
        for _ in range(3):
            e.scan_device(d_seq.data_ptr(), d_off.data_ptr(), n, n * L)
        rs = [e.scan_device(d_seq.data_ptr(), d_off.data_ptr(), n, n * L) for _ in range(args.steps)]
        f = e.scan_device(d_seq.data_ptr(), d_off.data_ptr(), n, n * L, fetch=True)
        m = min(args.check_reads, n)
        h_seq, h_off = syn.host_reads(0, m)
        want, _, _ = rm.AhoCorasick(pats).scan_batch(h_seq, h_off, 8, False)
        ok = bool(np.array_equal(want[: m // 64], f.flags[: m // 64]))
        scan = float(np.median([r.scan_ns for r in rs])) / 1e6
        dev = float(np.median([r.device_ns for r in rs])) / 1e6
        info = e.info()
        row = {"k": k, "patterns": len(pats), "kernel": e.scan_kernel(), "seed_d": int(info.seed_d[0]), "seed_q": int(info.seed_q[0]), "scan_ms": scan,
               "device_ms": dev, "scan_gb_per_s": n * L / scan / 1e6, "candidates": int(rs[-1].n_candidates), "flagged": int(np.bitwise_count(f.flags).sum()),
               "oracle_sample_equal": ok}
        rows.append(row)
        print(json.dumps(row), flush=True)
    del d_seq, d_off
if args.out:
    Path(args.out).write_text(json.dumps(rows, indent=1) + "\n")
