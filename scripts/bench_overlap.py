"""Throughput of a stream of device-resident batches under different orderings of their device work.

  base   the engine's default: batches run in submission order (batch i + 1 starts when batch i has ended)
  free   MK_FREE=1: the slots' streams run freely, so verification and sort of batch i compete with the scan of
         batch i + 1 for the SMs
(Round 2 also tried a scan queue — all scan kernels on one high-priority stream, everything else of a batch on its
own stream — and launch shapes that leave registers to the post-processing kernels; results in
profiles/r2_overlap_experiment/, code not kept.)

For one BASELINE config the workload is generated once and the tables are built once; every variant (a mode plus
optional launch-shape switches) gets its own engine on top of them. Per variant: K passes through
mk_scan_device_submit / mk_scan_wait with `depth` batches in flight, wall clock per pass (CUDA synchronised on
both sides), the per-batch device / scan / verify times, and the hit count (must not change between variants).

    python scripts/bench_overlap.py --config cfg4 [--steps 30] [--depth 3] [--variants base,free,free:MK_D16_SHAPE=2]
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

DEFAULT_VARIANTS = {c: "base,free" for c in ("cfg2", "cfg3", "cfg4", "cfg5")}
TUNING_VARS = ("MK_FREE", "MK_D16_SHAPE", "MK_DUAL_SHAPE", "MK_VERIFY_THREADS")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", choices=["cfg2", "cfg3", "cfg4", "cfg5", "cfg5_verbatim_case"], required=True)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--depth", type=int, default=3)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--variants", default=None)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()

    import torch
    from merkurio_b200 import capi
    from merkurio_b200 import patterns as pt
    from merkurio_b200.synth import Synth, workloads as wlm

    t0 = time.perf_counter()
    if args.config == "cfg2":
        n, L = int(100_000_000 * args.scale) // 64 * 64, 150
        syn = Synth(0x5EED0002, n, L, 31, 1000)
        pats = pt.parse_pattern_list(syn.query_list(), reverse_complement_=True)
        d_seq = torch.empty(n * L + 64, dtype=torch.uint8, device="cuda")
        d_off = torch.empty(n + 1, dtype=torch.int64, device="cuda")
        d_q = torch.from_numpy(syn.queries).cuda()
        syn.device_reads(d_q.data_ptr(), 0, n, d_seq.data_ptr(), d_off.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        wl = wlm.Workload("cfg2", d_seq, d_off, n, n * L, capi.MK_ENC_ASCII, capi.MK_MODE_FLAG, pats, n * L, 1 << 20, syn, L)
    elif args.config in ("cfg3", "cfg4"):
        wl = wlm.reads_workload(args.config, args.scale)
    else:
        wl = wlm.genome_workload(args.scale, upper_queries=(args.config == "cfg5"))
    print(f"[overlap] {wl.name}: generated in {time.perf_counter() - t0:.1f} s", file=sys.stderr)
    tables = capi.Tables(wl.pats)
    peak = 6554.2
    try:
        peak = float(json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
    except Exception:
        pass

    results = []
    for spec in (args.variants or DEFAULT_VARIANTS[args.config.replace("_verbatim_case", "")]).split(","):
        parts = spec.split(":")
        for v in TUNING_VARS:
            os.environ.pop(v, None)
        if parts[0] == "free":
            os.environ["MK_FREE"] = "1"
        for kv in parts[1:]:
            k, v = kv.split("=")
            os.environ[k] = v
        eng = capi.Engine(tables, n_slots=args.depth, max_batch_bytes=1 << 20, max_batch_records=1 << 10, hit_capacity=wl.hit_capacity)

        def run(k):
            out, pending = [], []
            for i in range(k):
                if len(pending) == args.depth:
                    out.append(eng.wait(pending.pop(0), copy=False))
                s = i % args.depth
                eng.scan_device_submit(s, wl.d_seq.data_ptr(), wl.d_off.data_ptr(), wl.n_records, wl.n_units, wl.mode, wl.enc)
                pending.append(s)
            while pending:
                out.append(eng.wait(pending.pop(0), copy=False))
            return out

        run(args.depth + 2)  # tables uploaded, lists grown
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rs = run(args.steps)
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
        single = wl.scan(eng)  # one batch alone (the `direct` workspace)
        entry = {"variant": spec, "kernel": eng.scan_kernel(wl.enc), "wall_ms_per_pass": wall_ms,
                 "frac_of_peak_stream": wl.algorithmic_bytes / wall_ms / 1e6 / peak,
                 "device_ms_median": float(np.median([r.device_ns for r in rs])) / 1e6,
                 "scan_ms_median": float(np.median([r.scan_ns for r in rs])) / 1e6,
                 "verify_ms_median": float(np.median([r.verify_ns for r in rs])) / 1e6,
                 "alone_device_ms": single.device_ns / 1e6, "alone_scan_ms": single.scan_ns / 1e6, "alone_verify_ms": single.verify_ns / 1e6,
                 "n_hits": int(rs[-1].n_hits), "n_candidates": int(rs[-1].n_candidates), "rescans": int(sum(r.n_rescans for r in rs))}
        results.append(entry)
        print("[overlap] " + json.dumps(entry), file=sys.stderr)
        eng.close()
    assert len({(r["n_hits"], r["n_candidates"]) for r in results}) == 1, "variants disagree on hits / candidates"
    out = {"config": wl.name, "steps": args.steps, "depth": args.depth, "algorithmic_bytes": int(wl.algorithmic_bytes), "peak_gbs": peak, "variants": results}
    print(json.dumps(out))
    if args.out:
        Path(args.out).write_text(json.dumps(out, indent=1) + "\n")


if __name__ == "__main__":
    main()
