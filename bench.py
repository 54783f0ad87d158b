#!/usr/bin/env python
"""bench.py — device-timed scan throughput of the matching hot path on BASELINE config 2
(synthetic single-end 100 M x 150 bp reads, 1 000 31-mers + reverse complements, `extract`).

  python bench.py [--gpus N] [--steps K] [--warmup W]           # our CUDA path (one rank per GPU)
  python bench.py --impl reference [--steps K] [--warmup W]     # the reference's CPU algorithm on host cores

A step is one pass of the scan over one batch = the whole per-GPU data set (100 M reads, 15 GB,
resident in HBM, far larger than the 126 MB L2 so no flush is needed). `value` is the whole-job
Gbases/s with inputs resident; `e2e` is the same workload pushed through the C ABI from pinned host
memory (H2D of every batch and D2H of the flags inside the timed region). Rank 0 prints one JSON line.

Beside the headline the line carries (N = 1 only, each can be switched off):
  sustained   >= 2 s of back-to-back cfg2 passes with per-pass device times and the clocks seen meanwhile
  configs     cfg3 / cfg4 / cfg5 at BASELINE size: device time, dominant kernel, fraction of the HBM peak,
              and parity with the oracle on a sample of the same data
  e2e_file    file -> file: the host binary `merkurio extract` on a FASTQ on tmpfs, wall clock, with the
              single-threaded oracle matcher timed on the same reads; `gzip_input`: the same on a gzip-compressed
              FASTQ through the host's own DEFLATE decoder (several threads / one thread) and through zlib
"""
from __future__ import annotations

import argparse
import hashlib
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
    """SM clock, power and throttle reasons of one GPU while a timed region runs: NVML (pynvml) polled from a
    thread every 10 ms; `nvidia-smi -lms` as the fallback when NVML cannot be loaded."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap", 0x80: "hw_power_brake"}
    SMI_Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, torch_device_index: int):
        self.rows = []  # (t, sm_mhz, power_w, reason bitmask)
        self.max_mhz = None
        self.source = None
        self._stop = threading.Event()
        self._th = None
        self._proc = None
        self._idx = torch_device_index

    def _nvml_handle(self):
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(self._idx).uuid)
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode() if not uuid.startswith("GPU-") else uuid.encode())
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self._idx]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self._idx
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)

    def start(self):
        try:
            nv, h = self._nvml_handle()
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            self.source = "nvml"

            def poll():
                while not self._stop.is_set():
                    try:
                        try:
                            reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                        except Exception:
                            reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        self.rows.append((time.perf_counter(), float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)),
                                          nv.nvmlDeviceGetPowerUsage(h) / 1000.0, int(reasons)))
                    except Exception:
                        pass
                    self._stop.wait(0.01)
            self._th = threading.Thread(target=poll, daemon=True)
            self._th.start()
            return
        except Exception as e:  # no NVML: nvidia-smi
            log(f"[bench] NVML unavailable ({e!r}); sampling clocks with nvidia-smi")
        try:
            self._proc = subprocess.Popen(["nvidia-smi", "-i", str(self._idx), f"--query-gpu={self.SMI_Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                          stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"

            def read():
                for line in self._proc.stdout:
                    f = [x.strip() for x in line.split(",")]
                    try:
                        mask = sum(bit for bit, v in zip((0x8, 0x40, 0x20, 0x4), f[3:7]) if v.lower().startswith("active"))
                        self.rows.append((time.perf_counter(), float(f[0]), float(f[2]), mask))
                        self.max_mhz = float(f[1])
                    except (ValueError, IndexError):
                        continue
            self._th = threading.Thread(target=read, daemon=True)
            self._th.start()
        except OSError:
            self.source = None

    def window(self, t0: float, t1: float) -> dict:
        """Summary of the samples taken in [t0, t1] (perf_counter times)."""
        rows = [r for r in self.rows if t0 <= r[0] <= t1]
        if not rows:  # a region shorter than the sampling period: the nearest samples around it
            rows = sorted(self.rows, key=lambda r: abs(r[0] - (t0 + t1) / 2))[:3]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"], "samples": 0, "source": self.source}
        mask = 0
        for r in rows:
            mask |= r[3]
        sm = [r[1] for r in rows]
        return {"sm_mhz": float(np.median(sm)), "sm_min_mhz": float(min(sm)), "sm_max_mhz": self.max_mhz, "power_w_max": float(max(r[2] for r in rows)),
                "reasons": sorted(name for bit, name in self.REASONS.items() if mask & bit), "samples": len(rows), "source": self.source}

    def stop(self):
        self._stop.set()
        if self._proc:
            self._proc.terminate()


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


def csrc_digest() -> str:
    """sha256 over the CUDA sources and the C ABI header: what an ncu capture of a kernel is valid for."""
    h = hashlib.sha256()
    for f in sorted((ROOT / "merkurio_b200" / "csrc").glob("*")) + [ROOT / "include" / "merkurio_cuda.h"]:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()[:16]


def ncu_traffic(kernel: str, workload_key: str):
    """DRAM bytes (read + write) of one launch of `kernel` from the committed ncu summary, or (None, why) when that
    capture was taken from other sources than the ones that are running now."""
    p = ROOT / "profiles" / "ncu_traffic.json"
    if not p.exists():
        return None, "no profiles/ncu_traffic.json"
    try:
        d = json.loads(p.read_text())
        e = d["kernels"][kernel][workload_key]
    except Exception:
        return None, f"profiles/ncu_traffic.json has no entry for {kernel} / {workload_key}"
    if e.get("csrc_sha16") != csrc_digest():
        return None, f"stale: captured at csrc {e.get('csrc_sha16')}, running {csrc_digest()}"
    return int(e["dram_bytes_read"]) + int(e["dram_bytes_write"]), f"{e['source']} (csrc {e['csrc_sha16']})"


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
def bench_config(name: str, steps: int, peak: float, oracle_check: bool = True, sampler=None):
    """One of cfg3 / cfg4 / cfg5 at BASELINE size: device-timed passes over the resident batch, the dominant (scan)
    kernel against the HBM peak, and parity: oracle on a sample of the same data + size-independent properties."""
    import torch
    from merkurio_b200 import capi
    from merkurio_b200.synth import workloads as wlm
    from oracle import checks

    t_gen = time.perf_counter()
    if name in ("cfg3", "cfg4"):
        wl = wlm.reads_workload(name)
    else:
        wl = wlm.genome_workload(1.0, upper_queries=(name == "cfg5"))
    t_gen = time.perf_counter() - t_gen
    entry = {"config": wl.name, "records": wl.n_records, "bases": wl.n_units, "patterns": len(wl.pats),
             "encoding": "BAM4" if wl.enc == capi.MK_ENC_BAM4 else "ASCII",
             "mode": {capi.MK_MODE_FLAG: "FLAG", capi.MK_MODE_PATTERN_SET: "PATTERN_SET", capi.MK_MODE_ALL_HITS: "ALL_HITS"}[wl.mode],
             "generate_s": t_gen}
    with capi.Engine(wl.pats, n_slots=0, hit_capacity=wl.hit_capacity) as eng:
        t0 = time.perf_counter()
        wl.scan(eng)  # first pass: waits for the seed tables (built on host threads while the engine starts)
        entry["table_build_and_first_pass_s"] = time.perf_counter() - t0
        # warm-up: at least 3 passes and 0.1 s of them. The legs in front of this one (the CPU baseline, the generators)
        # leave the GPU idle for seconds, and the first passes after that run at clocks that are still ramping up (cfg3, the
        # first config, measured 2.39 ms that way against 2.09 ms a moment later); half a second of passes, on the other
        # hand, runs a 1 kW part into its software power cap (SM clock 1.65-1.87 GHz: cfg3 2.64 ms, cfg4 0.66-0.70 ms), which
        # is the sustained regime, not that of 10 timed passes. `clocks` of the entry says which regime a run saw.
        t_warm = time.perf_counter() + 0.1
        n_warm = 0
        while n_warm < 3 or time.perf_counter() < t_warm:
            wl.scan(eng)
            n_warm += 1
        entry["warmup_passes"] = n_warm
        scan, dev, ver = [], [], []
        t_steps0 = time.perf_counter()
        for _ in range(steps):
            r = wl.scan(eng)
            scan.append(r.scan_ns / 1e6)
            dev.append(r.device_ns / 1e6)
            ver.append(r.verify_ns / 1e6)
        if sampler is not None:  # the clocks seen from the start of the warm-up to the last timed pass
            entry["clocks"] = sampler.window(t_warm - 0.1, time.perf_counter())
        entry["kernel_ms_per_step"] = [round(x, 4) for x in scan]
        entry["timed_steps_wall_s"] = time.perf_counter() - t_steps0
        full = wl.scan(eng, fetch=True)
        info = eng.info()
        enc = wl.enc
        scan_ms, dev_ms = float(np.median(scan)), float(np.median(dev))
        entry.update(
            device_ms=dev_ms, device_ms_min=float(min(dev)), kernel=eng.scan_kernel(enc),
            kernel_ms=scan_ms, verify_ms=float(np.median(ver)), sort_and_rest_ms=dev_ms - scan_ms - float(np.median(ver)),
            algorithmic_bytes=int(wl.algorithmic_bytes), achieved_gbs=wl.algorithmic_bytes / scan_ms / 1e6,
            frac=wl.algorithmic_bytes / scan_ms / 1e6 / peak, frac_whole_pass=wl.algorithmic_bytes / dev_ms / 1e6 / peak,
            gbases_per_s=wl.n_units / dev_ms / 1e6, n_hits=int(full.n_hits), candidates=int(full.n_candidates), rescans=int(full.n_rescans),
            seed_q=int(info.seed_q[enc]), seed_d=int(info.seed_d[enc]), seeds=int(info.n_seeds[enc]), filter_in_smem=int(info.filter_in_smem[enc]),
            filter_bytes=int(info.filter_bytes[enc]), table_bytes=int(info.table_bytes[enc]), steps=steps)
        # parity
        t0 = time.perf_counter()
        if wl.syn is not None:
            entry["oracle_sample"] = f"first 200000 reads, Aho-Corasick (oracle/mk_oracle.c), bit-exact {entry['mode']}"
            entry["oracle_sample_hits"] = checks.reads_sample_equal(full, wl, 200_000)
            entry["oracle_sample_equal"] = True
            if wl.mode == capi.MK_MODE_ALL_HITS:
                entry["hits_reverified"] = checks.reverify_hits(wl.d_seq, wl.d_off, full.hits, wl.pats, wl.enc == capi.MK_ENC_BAM4)
        else:
            assert info.filter_in_smem[enc] == 0, "cfg5 must run through the L2-resident filter path"
            entry["hits_reverified"] = checks.reverify_hits(wl.d_seq, wl.d_off, full.hits, wl.pats, False)
            entry["expected_queries_missing"] = checks.genome_expected_found(full, wl)
            entry["lower_case_frac"], entry["n_frac"] = wl.extra["lower_case_frac"], wl.extra["n_frac"]
            assert entry["hits_reverified"] and entry["expected_queries_missing"] == 0
            if name == "cfg5" and oracle_check:
                entry["oracle_sample"] = ("hits of the full run (same engine, all queries, L2 dual-key filter) inside the first 4 Mbp of record 0 "
                                          "vs the oracle's Aho-Corasick automaton of all queries over that slice")
                entry["oracle_sample_hits"] = checks.genome_slice_equal(full, wl, 4_000_000)
                entry["oracle_sample_equal"] = True
        entry["check_s"] = time.perf_counter() - t0
    del wl, full
    torch.cuda.empty_cache()
    return entry


def bench_e2e_file(args):
    """File -> file: `merkurio extract -i reads.fastq -f q.txt -r -o out.fastq` (the host binary, its own process, CUDA
    start-up included) on a synthetic FASTQ on tmpfs, via scripts/bench_cli.py (which also checks the extracted set
    against the oracle), and the oracle's single-threaded Aho-Corasick any-hit scan of the same reads beside it."""
    from merkurio_b200 import patterns as pt
    from merkurio_b200.synth import Synth
    from oracle import refmodel as rm
    n = args.file_reads
    import shutil
    need = n * 340 * 1.2  # the FASTQ (~320 bytes per read) plus what is extracted
    shm = next((Path(c) for c in ("/dev/shm", "/tmp", str(ROOT / "gpurun_out")) if Path(c).is_dir() and shutil.disk_usage(c).free > need), Path("/tmp"))
    d = shm / f"mk_bench_file_{os.getpid()}"
    out_json = d / "cli.json"
    try:
        d.mkdir(parents=True, exist_ok=True)
        pr = subprocess.run([sys.executable, str(ROOT / "scripts" / "bench_cli.py"), "--config", "cfg2", "--reads", str(n), "--dir", str(d), "--out", str(out_json)],
                            stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
        if pr.returncode != 0:
            return {"error": pr.stderr[-400:]}
        res = json.loads(out_json.read_text())
        # the same command on gzip-compressed input (how reads are stored in practice; binned random qualities, level 1):
        # the host's own DEFLATE decoder on several threads (host/pgzip.cpp), on one thread (host/inflate.cpp), and
        # zlib's inflate (MERKURIO_ZLIB_INFLATE=1, the round-1 path) beside it
        gz = {}
        if args.file_gz_reads > 0:
            for key, extra in (("own_decoder_parallel", {}), ("own_decoder_one_thread", {"MERKURIO_GZIP_THREADS": "1"}), ("zlib", {"MERKURIO_ZLIB_INFLATE": "1"})):
                pg = subprocess.run([sys.executable, str(ROOT / "scripts" / "bench_cli.py"), "--config", "cfg2", "--gz", "--reuse", "--reads", str(args.file_gz_reads),
                                     "--dir", str(d), "--out", str(out_json)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600,
                                    env={**os.environ, **extra})
                if pg.returncode != 0:
                    gz[key] = {"error": pg.stderr[-300:]}
                    continue
                rg = json.loads(out_json.read_text())
                gz[key] = {"records_per_s": rg["records_per_s"], "wall_s": rg["wall_s"], "cuda_startup_s": rg["engine_setup_s"][int(np.argmin(rg["runs_s"]))],
                           "input_bytes": rg["input_bytes"], "reads": rg["reads"]}
                # the ingest alone (decompress + index the records, no CUDA in the process): `merkurio records <file> count`.
                # The extract run above decodes while CUDA starts up, so its wall clock shows the decoder only on boxes
                # where start-up is short
                fq = d / "reads_1.fastq.gz"
                best = None
                for _ in range(3):
                    t0 = time.perf_counter()
                    pc = subprocess.run([str(ROOT / "merkurio_b200" / "lib" / "merkurio"), "records", str(fq), "count"], stdout=subprocess.PIPE,
                                        stderr=subprocess.PIPE, text=True, env={**os.environ, **extra})
                    dt = time.perf_counter() - t0
                    if pc.returncode == 0 and pc.stdout.split()[:1] == [str(args.file_gz_reads)]:
                        best = dt if best is None else min(best, dt)
                if best:
                    gz[key]["ingest_only_records_per_s"] = args.file_gz_reads / best
                    gz[key]["ingest_only_gb_per_s_decompressed"] = args.file_gz_reads * 320 / best / 1e9
    finally:
        shutil.rmtree(d, ignore_errors=True)
    # the reference's matcher on the same reads, one thread, no parsing and no output (it can only be faster than the reference)
    m = min(n, args.file_ref_reads)
    syn = Synth(SEED, n, 150, K_MER, 1000)
    pats = pt.parse_pattern_list(syn.query_list(), reverse_complement_=True)
    seq, off = syn.host_reads(0, m)
    ac = rm.AhoCorasick(pats)
    t0 = time.perf_counter()
    ac.scan_batch(seq, off, 1, False)
    t_ref = time.perf_counter() - t0
    best = int(np.argmin(res["runs_s"]))
    wall, startup = res["runs_s"][best], res["engine_setup_s"][best]
    return {"workload": f"cfg2 shape, {n} reads x 150 bp FASTQ ({res['input_bytes'] / 1e9:.2f} GB) on {shm}, extract -f q.txt -r -o out.fastq",
            "records_per_s": n / wall, "records_per_s_after_cuda_startup": n / (wall - startup),
            "gbases_per_s": n * 150 / wall / 1e9, "wall_s": wall, "cuda_startup_s": startup, "runs_s": res["runs_s"], "cuda_startup_of_runs_s": res["engine_setup_s"],
            "extracted_set_equals_oracle": True, "host_cores": res["host_cores"],
            "reference_matcher_single_thread": {"records_per_s": m / t_ref, "gbases_per_s": m * 150 / t_ref / 1e9, "reads": m, "kind": "port",
                                                "what": "oracle Aho-Corasick any-hit scan of the same reads in memory: no parsing, no output"},
            "speedup_vs_reference_matcher": (n / wall) / (m / t_ref), "gzip_input": gz, "stages_last_run": res.get("stages_last_run", [])}


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
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))

    def barrier():
        torch.cuda.synchronize()
        dctx.barrier()

    # file -> file first: the host binary is its own process with its own CUDA start-up, which is slower (and varies more)
    # while another process holds a context and tens of GB on the same GPU — so before this process creates its context
    e2e_file = None
    if rank == 0 and world == 1 and not args.no_e2e_file:
        try:
            e2e_file = bench_e2e_file(args)
        except Exception as ex:
            e2e_file = {"error": repr(ex)[:300]}
        log(f"[bench] e2e_file: {json.dumps({k: v for k, v in e2e_file.items() if k != 'stages_last_run'})}")
    torch.cuda.set_device(local)
    dctx = Dist("nccl", torch.device("cuda", local))  # barrier / max-over-ranks only: no data-path collective
    assert (rank, world) == (dctx.rank, dctx.world)
    max_over_ranks, sum_over_ranks = dctx.max, dctx.sum
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # well before the first timed region

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
    clocks = sampler.window(t0, t1) if rank == 0 else None
    elapsed = max_over_ranks(t1 - t0)
    value = world * n_bytes * args.steps / elapsed / 1e9
    scan_ms = float(np.mean(scan_ns)) / 1e6
    dev_ms_max = max_over_ranks(float(np.mean(dev_ns)) / 1e6)

    # ---- the other BASELINE configs (rank 0, N == 1) ----------------------------------------------------------------
    # Before the sustained run and the end-to-end legs, not after them: seconds at 1 kW leave the part in its software
    # power cap for a while, and the kernels that are not HBM-bound (cfg3's and cfg4's scans) then measure 10-15 % slower
    # than in a process of their own (cfg4: 0.66-0.74 ms after those legs, 0.606-0.608 ms in four separate processes).
    peak, peak_src = measured_peak_gbs()
    configs = None
    if rank == 0 and world == 1:
        if not args.no_configs:
            configs = []
            for name in args.configs.split(","):
                t0c = time.perf_counter()
                try:
                    configs.append(bench_config(name.strip(), args.config_steps, peak, sampler=sampler))
                except Exception as ex:  # reported, never hidden: a failed parity check must show in the line
                    configs.append({"config": name, "error": repr(ex)[:300], "oracle_sample_equal": False})
                log(f"[bench] {name}: {time.perf_counter() - t0c:.1f} s  {json.dumps({k: v for k, v in configs[-1].items() if k in ('kernel_ms', 'device_ms', 'frac', 'n_hits', 'oracle_sample_equal', 'error')})}")

    # ---- sustained: seconds of back-to-back passes (thermal / power evidence for the burst figure) -------
    sustained = None
    if args.sustained_s > 0:
        k_sus = max(int(args.sustained_s * 1e3 / max(dev_ms_max, 0.1)) + 1, args.steps)
        barrier()
        t0s = time.perf_counter()
        rs = run_steps(k_sus)
        barrier()
        t1s = time.perf_counter()
        el = max_over_ranks(t1s - t0s)
        dms = np.array([x.device_ns for x in rs], dtype=np.float64) / 1e6
        sms = np.array([x.scan_ns for x in rs], dtype=np.float64) / 1e6
        sustained = {"steps": k_sus, "seconds": el, "value": world * n_bytes * k_sus / el / 1e9, "unit": UNIT,
                     "device_ms_median": float(np.median(dms)), "device_ms_min": float(dms.min()), "device_ms_max": float(dms.max()),
                     "device_ms_last_tenth_median": float(np.median(dms[-max(k_sus // 10, 1):])),
                     "scan_ms_median": float(np.median(sms)),
                     "scan_gbs_median": (n_bytes + (n_reads + 7) // 8) / float(np.median(sms)) / 1e6,
                     "clocks": sampler.window(t0s, t1s) if rank == 0 else None}

    # ---- end to end through the C ABI from pinned host memory ---------------------------------
    e2e = None
    if not args.no_e2e:
        h_seq = torch.empty(n_bytes + 64, dtype=torch.uint8, pin_memory=True)
        h_seq.copy_(d_seq)
        torch.cuda.synchronize()
        n_batches = (n_reads + batch_reads - 1) // batch_reads
        base_ptr = h_seq.data_ptr()

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
        # the ceiling of that leg: the same pinned buffer copied to the device by bare cudaMemcpyAsync calls of the same
        # size, on every rank at once (what the host's memory and PCIe fabric give N concurrent streams)
        d_tmp = torch.empty(batch_reads * L, dtype=torch.uint8, device="cuda")
        st = torch.cuda.Stream()
        def h2d_pass():
            with torch.cuda.stream(st):
                for b in range(n_batches):
                    nb = min(batch_reads, n_reads - b * batch_reads) * L
                    d_tmp[:nb].copy_(h_seq[b * batch_reads * L: b * batch_reads * L + nb], non_blocking=True)
            st.synchronize()
        h2d_pass()
        barrier()
        t0c = time.perf_counter()
        h2d_pass()
        barrier()
        ceil_s = max_over_ranks(time.perf_counter() - t0c)
        ceiling = world * n_bytes / ceil_s / 1e9
        del d_tmp
        e2e = {"value": world * n_bytes * args.e2e_steps / el / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": int(world * n_bytes),
               "d2h_bytes_per_step": int(world * n_batches * ((batch_reads + 63) // 64 * 8 + 16)),
               "steps": args.e2e_steps, "ms_per_step": el / args.e2e_steps * 1e3,
               "records_per_s": world * n_reads * args.e2e_steps / el,
               "h2d_ceiling_gb_per_s": ceiling, "frac_of_h2d_ceiling": (world * n_bytes * args.e2e_steps / el / 1e9) / ceiling,
               "h2d_ceiling_how": f"bare cudaMemcpyAsync of the same {n_batches} x {batch_reads * L} byte pieces from the same pinned buffer, all {world} ranks at once",
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
        del h, ac

    total_flagged = sum_over_ranks(float(flagged_resident))
    eng.close()
    del d_seq, d_off
    torch.cuda.empty_cache()

    if rank == 0:
        algo_bytes = n_bytes + (n_reads + 7) // 8  # sequence bytes + flag bitmap; offsets are only read for hits
        achieved = algo_bytes / (scan_ms / 1e3) / 1e9
        kernel = "mk_scan_d16<ASCII, smem filter, U=4, T=896>"
        traffic, traffic_source = ncu_traffic("mk_scan_d16", f"cfg2:{n_reads}x{L}:{args.queries}")
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
            "device_ms_per_step": dev_ms_max, "clocks": clocks, "sustained": sustained, "e2e": e2e, "gpu_launches": 2 * args.steps * world,
            "kernels_per_step": {"mk_scan_d16": 1, "mk_verify_candidates": 1, "verify_ms": float(np.mean(ver_ns)) / 1e6, "candidates": int(n_cand)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_source, "kernel": kernel, "kernel_ms": scan_ms, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(algo_bytes)},
            "cpu_baseline": cpu, "configs": configs, "e2e_file": e2e_file, "csrc_sha16": csrc_digest(),
        }
        emit(line)
    sampler.stop()
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
    ap.add_argument("--sustained-s", type=float, default=2.5, help="seconds of back-to-back passes for the `sustained` entry (0: skip)")
    ap.add_argument("--configs", default="cfg3,cfg4,cfg5,cfg5_verbatim_case", help="the other BASELINE configs measured into `configs` (N = 1)")
    ap.add_argument("--config-steps", type=int, default=10)
    ap.add_argument("--no-configs", action="store_true")
    ap.add_argument("--no-e2e-file", action="store_true")
    ap.add_argument("--file-reads", type=int, default=12_000_000, help="reads of the FASTQ of the file -> file run")
    ap.add_argument("--file-gz-reads", type=int, default=2_000_000, help="reads of the gzip-compressed FASTQ of the file -> file run (0: skip)")
    ap.add_argument("--file-ref-reads", type=int, default=4_000_000, help="reads the single-threaded oracle matcher is timed on")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        log("[bench] note: fewer than 3 warm-up steps requested; using 3")
        args.warmup = 3
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
