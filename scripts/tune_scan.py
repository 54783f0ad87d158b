"""Time scan-kernel launch shapes on a device-resident cfg2 data set (CUDA events inside the engine).
    MK_TUNE_BUILD=1 python -m merkurio_b200.build --force   # builds every launch shape
    python scripts/tune_scan.py [--reads N] [--variants "U,T;U,T;..."]"""
import argparse
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from merkurio_b200 import capi, patterns as pt
from merkurio_b200.synth import Synth

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=100_000_000)
ap.add_argument("--queries", type=int, default=1000)
ap.add_argument("--k", type=int, default=31)
ap.add_argument("--variants", default="2,1024;4,768;2,768;4,512;8,512;2,512", help="U,T pairs")
ap.add_argument("--env", default="", help="extra NAME=VALUE pairs, ';' separated, applied to every variant")
ap.add_argument("--steps", type=int, default=8)
args = ap.parse_args()

n, L = args.reads, 150
syn = Synth(0x5EED0002, n, L, args.k, args.queries)
pats = pt.parse_pattern_list(syn.query_list(), reverse_complement_=True)
d_seq = torch.empty(n * L + 64, dtype=torch.uint8, device="cuda")
d_off = torch.empty(n + 1, dtype=torch.int64, device="cuda")
d_q = torch.from_numpy(syn.queries).cuda()
syn.device_reads(d_q.data_ptr(), 0, n, d_seq.data_ptr(), d_off.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
for kv in filter(None, args.env.split(";")):
    k, v = kv.split("=")
    os.environ[k] = v
ref = None
for var in args.variants.split(";"):
    u, pf, *rest = var.split(",")
    os.environ["MK_TUNE_V8"] = rest[0] if rest else "0"
    os.environ["MK_TUNE_U"], os.environ["MK_TUNE_T"] = u, pf
    with capi.Engine(pats, n_slots=0) as e:
        for _ in range(3):
            e.scan_device(d_seq.data_ptr(), d_off.data_ptr(), n, n * L)
        ts = [e.scan_device(d_seq.data_ptr(), d_off.data_ptr(), n, n * L).scan_ns for _ in range(args.steps)]
        f = e.scan_device(d_seq.data_ptr(), d_off.data_ptr(), n, n * L, fetch=True)
        cnt = int(np.bitwise_count(f.flags).sum())
        if ref is None:
            ref = cnt
        ms = np.median(ts) / 1e6
        print(f"U={u} T={pf} V8={os.environ['MK_TUNE_V8']}: scan {ms:.3f} ms (min {min(ts) / 1e6:.3f})  {n * L / ms / 1e6:.0f} GB/s  flagged={cnt} {'ok' if cnt == ref else 'MISMATCH'}", flush=True)
