#!/usr/bin/env python
"""bench.py — device-timed scan throughput of the matching hot path on BASELINE config 2
(synthetic single-end 100 M x 150 bp reads, 1 000 31-mers + reverse complements, `extract`).

  python bench.py [--gpus N] [--steps K] [--warmup W]           # our CUDA path (one rank per GPU)
  python bench.py --impl reference [--steps K] [--warmup W]     # the reference's CPU algorithm on host cores

A step is one pass of the scan over one batch = the whole per-GPU data set (100 M reads, 15 GB,
resident in HBM, far larger than the 126 MB L2 so no flush is needed). `value` is the whole-job
Gbases/s with inputs resident; `e2e` is the same workload pushed through the C ABI from pinned host
memory (H2D of every batch and D2H of the flags inside the timed region). Rank 0 prints one JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "Gbases/s scanned (device-timed)"
UNIT = "Gbases/s"
SEED = 0x5EED0002
K_MER = 31


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.05] or [r for _, r in self.rows[-3:]]
        sm, mx, reasons = [], [], set()
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


def workload_name(n_reads, read_len, n_queries):
    return (f"cfg2: synthetic single-end reads, {n_reads / 1e6:g}M x {read_len} bp per GPU, {n_queries} {K_MER}-mers + reverse "
            f"complements, extract (reference mode: Aho-Corasick; device mode: FLAG)")


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's CPU algorithm (Aho-Corasick DFA, any-hit per record as
    src/cmd_extract.rs:332-335) on the host cores. The reference itself is Rust and cannot be built in
    this image, so this times the oracle port (oracle/mk_oracle.c) — kind "port"."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from merkurio_b200 import patterns as pt
    from merkurio_b200.synth import Synth
    from oracle import refmodel as rm
    from concurrent.futures import ThreadPoolExecutor

    cores = os.cpu_count() or 1
    sample = args.ref_sample_reads
    syn = Synth(SEED, args.reads, args.read_len, K_MER, args.queries)
    pats = pt.parse_pattern_list(syn.query_list(), reverse_complement_=True)
    t0 = time.perf_counter()
    parts = min(cores, 32)
    step = (sample + parts - 1) // parts
    with ThreadPoolExecutor(parts) as ex:
        chunks = list(ex.map(lambda i: syn.host_reads(i * step, min(sample, (i + 1) * step))[0], range(parts)))
    seq = np.concatenate(chunks)
    off = np.arange(sample + 1, dtype=np.uint64) * np.uint64(args.read_len)
    log(f"[reference] generated {sample} reads on the host in {time.perf_counter() - t0:.1f} s")
    ac = rm.AhoCorasick(pats)
    log(f"[reference] AC DFA: {ac.n_states()} states, {rm.lib().mko_ac_table_bytes(ac.handle) / 1e6:.1f} MB")
    for _ in range(args.warmup):
        ac.scan_batch(seq, off, cores, False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flags, nrec, _ = ac.scan_batch(seq, off, cores, False)
    dt = time.perf_counter() - t0
    gb = sample * args.read_len * args.steps / dt / 1e9
    desc = f"first {sample} reads of the workload per step, {cores} threads, records partitioned over threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": gb, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(args.reads, args.read_len, args.queries), "sample": desc,
                   "note": "CPU restatement of the reference (aho-corasick DFA, any-hit per record); the Rust reference cannot be built in this image"},
        "cpu_baseline": {"value": gb, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": gb, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "records_hit_in_sample": int(nrec),
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch

    from merkurio_b200 import capi
    from merkurio_b200 import patterns as pt
    from merkurio_b200.shard import Dist, shard_range
    from merkurio_b200.synth import Synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the matching engine has no CPU path")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dctx = Dist("nccl", torch.device("cuda", local))  # barrier / max-over-ranks only: no data-path collective
    rank, world = dctx.rank, dctx.world

    def barrier():
        torch.cuda.synchronize()
        dctx.barrier()

    max_over_ranks, sum_over_ranks = dctx.max, dctx.sum

    n_reads, L = args.reads, args.read_len
    n_bytes = n_reads * L
    # the data set of the whole job is world * n_reads reads; this rank owns [rank*n_reads, (rank+1)*n_reads)
    syn = Synth(SEED, n_reads * world, L, K_MER, args.queries)
    pats = pt.parse_pattern_list(syn.query_list(), reverse_complement_=True)

    d_seq = torch.empty(n_bytes + 64, dtype=torch.uint8, device="cuda")
    d_off = torch.empty(n_reads + 1, dtype=torch.int64, device="cuda")
    d_q = torch.from_numpy(syn.queries).cuda()
    t0 = time.perf_counter()
    r_lo, r_hi = shard_range(n_reads * world, world, rank)  # this rank's records of the whole job
    assert r_hi - r_lo == n_reads
    syn.device_reads(d_q.data_ptr(), r_lo, r_hi, d_seq.data_ptr(), d_off.data_ptr(), 0,
                     torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    if rank == 0:
        log(f"[bench] generated {n_reads} reads x {L} bp on the device in {time.perf_counter() - t0:.2f} s")

    batch_reads = min(args.batch_reads, n_reads)
    eng = capi.Engine(pats, device=local, n_slots=args.slots, max_batch_bytes=batch_reads * L, max_batch_records=batch_reads)

    def step():
        return eng.scan_device(d_seq.data_ptr(), d_off.data_ptr(), n_reads, n_bytes, capi.MK_MODE_FLAG, capi.MK_ENC_ASCII)

    # The timed steps keep two passes in flight (mk_scan_device_submit on alternating slots, mk_scan_wait on
    # the older one), as the streaming path does: the GPU does not idle while the host collects a result.
    depth = 2 if args.slots >= 2 else 1

    def run_steps(k):
        out, pending = [], []
        for i in range(k):
            if len(pending) == depth:
                out.append(eng.wait(pending.pop(0), copy=False))
            eng.scan_device_submit(i % depth, d_seq.data_ptr(), d_off.data_ptr(), n_reads, n_bytes, capi.MK_MODE_FLAG, capi.MK_ENC_ASCII)
            pending.append(i % depth)
        while pending:
            out.append(eng.wait(pending.pop(0), copy=False))
        return out

    r = step()
    run_steps(max(args.warmup, 1))
    info = eng.info()
    # number of flagged reads of the resident pass (device bitmap -> host once, outside any timing)
    fl = eng.scan_device(d_seq.data_ptr(), d_off.data_ptr(), n_reads, n_bytes, capi.MK_MODE_FLAG, capi.MK_ENC_ASCII, fetch=True)
    flagged_resident = int(np.bitwise_count(fl.flags).sum())
    resident_flags = fl.flags.copy()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    t0 = time.perf_counter()
    scan_ns, dev_ns, ver_ns, n_cand = [], [], [], 0
    for r in run_steps(args.steps):
        scan_ns.append(r.scan_ns)
        dev_ns.append(r.device_ns)
        ver_ns.append(r.verify_ns)
        n_cand = r.n_candidates
    barrier()
    t1 = time.perf_counter()
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    elapsed = max_over_ranks(t1 - t0)
    value = world * n_bytes * args.steps / elapsed / 1e9
    scan_ms = float(np.mean(scan_ns)) / 1e6
    dev_ms_max = max_over_ranks(float(np.mean(dev_ns)) / 1e6)

    # ---- end to end through the C ABI from pinned host memory ---------------------------------
    e2e = None
    if not args.no_e2e:
        h_seq = torch.empty(n_bytes + 64, dtype=torch.uint8, pin_memory=True)
        h_seq.copy_(d_seq)
        rel_off = torch.empty(batch_reads + 1, dtype=torch.int64, pin_memory=True)
        rel_off.copy_(torch.arange(batch_reads + 1, dtype=torch.int64) * L)
        torch.cuda.synchronize()
        n_batches = (n_reads + batch_reads - 1) // batch_reads
        base_ptr, off_ptr = h_seq.data_ptr(), rel_off.data_ptr()

        def e2e_pass(collect=None):
            flagged, pending = 0, []
            for b in range(n_batches):
                slot = b % args.slots
                if len(pending) == args.slots:
                    s0, b0 = pending.pop(0)
                    res = eng.wait(s0, copy=False)
                    flagged += int(np.bitwise_count(res.flags).sum())
                    if collect is not None:
                        collect.append((b0, res.flags.copy()))
                nb = min(batch_reads, n_reads - b * batch_reads)
                # the reads have one common length: mk_scan_host_uniform, no offset array crosses the bus
                eng.scan_host_uniform_async(slot, base_ptr + b * batch_reads * L, nb, L, capi.MK_ENC_ASCII, capi.MK_MODE_FLAG)
                pending.append((slot, b))
            for s0, b0 in pending:
                res = eng.wait(s0, copy=False)
                flagged += int(np.bitwise_count(res.flags).sum())
                if collect is not None:
                    collect.append((b0, res.flags.copy()))
            return flagged

        coll = []
        got = e2e_pass(coll)  # warm-up pass, also checks the streamed result against the resident one
        assert got == flagged_resident, (got, flagged_resident)
        if batch_reads % 64 == 0:
            merged = np.concatenate([f for _, f in sorted(coll, key=lambda x: x[0])])
            assert np.array_equal(merged[: resident_flags.size], resident_flags), "streamed flags differ from resident flags"
        del coll
        barrier()
        t0e = time.perf_counter()
        for _ in range(args.e2e_steps):
            got = e2e_pass()
        barrier()
        t1e = time.perf_counter()
        el = max_over_ranks(t1e - t0e)
        e2e = {"value": world * n_bytes * args.e2e_steps / el / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": int(world * n_bytes),
               "d2h_bytes_per_step": int(world * n_batches * ((batch_reads + 63) // 64 * 8 + 16)),
               "steps": args.e2e_steps, "ms_per_step": el / args.e2e_steps * 1e3,
               "records_per_s": world * n_reads * args.e2e_steps / el,
               "path": f"mk_scan_host_uniform/mk_scan_wait, {args.slots} slots, batches of {batch_reads} reads from pinned host memory"}
        del h_seq

    # ---- CPU baseline beside it (rank 0, N == 1 only) ------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import refmodel as rm
        cores = os.cpu_count() or 1
        sample = min(args.cpu_sample_reads, n_reads)
        h = d_seq[: sample * L].cpu().numpy()
        off = np.arange(sample + 1, dtype=np.uint64) * np.uint64(L)
        ac = rm.AhoCorasick(pats)
        one = min(sample, max(sample // 8, 1))
        t0c = time.perf_counter()
        ac.scan_batch(h[: one * L], off[: one + 1], 1, False)
        t_one = time.perf_counter() - t0c
        ac.scan_batch(h, off, cores, False)
        t0c = time.perf_counter()
        cflags, crec, _ = ac.scan_batch(h, off, cores, False)
        t_all = time.perf_counter() - t0c
        # parity at scale: the oracle's flag bitmap of the sample equals the device's
        words = sample // 64
        assert np.array_equal(cflags[:words], resident_flags[:words]), "device flags differ from the oracle on the CPU sample"
        cpu = {"value": sample * L / t_all / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"first {sample} reads of the workload, Aho-Corasick DFA any-hit scan (oracle/mk_oracle.c), {cores} threads; "
                         f"single thread on {one} reads: {one * L / t_one / 1e9:.3f} Gbases/s",
               "single_thread_value": one * L / t_one / 1e9, "flags_equal_device": True}

    total_flagged = sum_over_ranks(float(flagged_resident))
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        algo_bytes = n_bytes + (n_reads + 7) // 8  # sequence bytes + flag bitmap; offsets are only read for hits
        achieved = algo_bytes / (scan_ms / 1e3) / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum of mk_scan_d16 from the ncu --set full capture of exactly this
        # workload (profiles/r1_ncu_full_cfg2.txt: 15.001007 GB read + 14.089984 MB written); null for any other size
        traffic = 15_001_007_000 + 14_089_984 if (n_reads, L, args.queries) == (100_000_000, 150, 1000) else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(n_reads, L, args.queries), "patterns": len(pats), "reads_per_gpu": n_reads,
                       "seed_q": int(info.seed_q[0]), "seed_d": int(info.seed_d[0]), "filter_hashes": int(info.filter_hashes[0]),
                       "filter_in_smem": int(info.filter_in_smem[0]), "table_bytes": int(info.table_bytes[0]),
                       "l2": "per-step input (15 GB) is far larger than the 126 MB L2; no flush needed",
                       "timing": "wall clock around K passes between barriers, two in flight (mk_scan_device_submit / mk_scan_wait on alternating slots); device_ms_per_step is the CUDA-event time of one pass on its stream",
                       "records_flagged": int(total_flagged)},
            "device_ms_per_step": dev_ms_max, "clocks": clocks, "e2e": e2e, "gpu_launches": 2 * args.steps * world,
            "kernels_per_step": {"mk_scan_d16": 1, "mk_verify_candidates": 1, "verify_ms": float(np.mean(ver_ns)) / 1e6, "candidates": int(n_cand)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "mk_scan_d16<ASCII, smem filter, U=4, T=896>", "kernel_ms": scan_ms, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(algo_bytes)},
            "cpu_baseline": cpu,
        }
        emit(line)
    eng.close()
    dctx.close()
    return 0


_real_stdout = None


def emit(line: dict):
    """The one JSON line, on the process's real stdout (see main)."""
    os.write(_real_stdout if _real_stdout is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    # Libraries chat on stdout (NCCL prints its version there when NCCL_DEBUG is set): route file
    # descriptor 1 to stderr for the whole run and keep the original for the JSON line alone.
    global _real_stdout
    sys.stdout.flush()
    _real_stdout = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--reads", type=int, default=100_000_000, help="reads per GPU (BASELINE config 2: 100 M)")
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--queries", type=int, default=1000)
    ap.add_argument("--batch-reads", type=int, default=1 << 20, help="reads per batch of the end-to-end leg")
    ap.add_argument("--slots", type=int, default=3)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-reads", type=int, default=16_000_000)
    ap.add_argument("--ref-sample-reads", type=int, default=8_000_000)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        log("[bench] note: fewer than 3 warm-up steps requested; using 3")
        args.warmup = 3
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
