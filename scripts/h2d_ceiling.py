"""What the host gives N GPUs at once: pinned host -> device copy bandwidth with 1, 2, 4, ... GPUs copying concurrently
(one pinned buffer and one stream per GPU, bare cudaMemcpyAsync through torch), the ceiling of bench.py's `e2e` leg
and of any host-fed pipeline. Also prints the topology the numbers belong to.

    python scripts/h2d_ceiling.py [--mb 1024] [--seconds 1.0] [--out profiles/r2_h2d_ceiling.json]
"""
import argparse
import json
import os
import subprocess
import time

import torch

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=int, default=1024)
ap.add_argument("--seconds", type=float, default=1.0)
ap.add_argument("--out", default=None)
args = ap.parse_args()
n_gpu = torch.cuda.device_count()
nbytes = args.mb << 20
host = [torch.empty(nbytes, dtype=torch.uint8, pin_memory=True) for _ in range(n_gpu)]
dev, streams = [], []
for g in range(n_gpu):
    with torch.cuda.device(g):
        dev.append(torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{g}"))
        streams.append(torch.cuda.Stream(device=g))


def run(gpus, seconds):
    def one_round():
        for g in gpus:
            with torch.cuda.device(g), torch.cuda.stream(streams[g]):
                dev[g].copy_(host[g], non_blocking=True)
    one_round()
    for g in gpus:
        streams[g].synchronize()
    t0 = time.perf_counter()
    rounds = 0
    while time.perf_counter() - t0 < seconds:
        one_round()
        for g in gpus:
            streams[g].synchronize()
        rounds += 1
    dt = time.perf_counter() - t0
    return rounds * len(gpus) * nbytes / dt / 1e9


res = {"gpus": n_gpu, "buffer_mb": args.mb, "host_cores": os.cpu_count(), "subsets": []}
sets = [[0]] + [list(range(k)) for k in (2, 4, 8) if k <= n_gpu]
if n_gpu >= 8:
    sets += [[0, 2, 4, 6], [0, 4], [4, 5, 6, 7]]
for s in sets:
    gbs = run(s, args.seconds)
    res["subsets"].append({"gpus": s, "total_gb_per_s": gbs, "per_gpu_gb_per_s": gbs / len(s)})
    print(f"GPUs {s}: {gbs:.1f} GB/s in total, {gbs / len(s):.1f} per GPU", flush=True)
for cmd in (["nvidia-smi", "topo", "-m"], ["lscpu"], ["numactl", "-H"]):
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=20).stdout
    except Exception as e:
        out = f"{cmd[0]}: {e}"
    res[" ".join(cmd)] = [ln for ln in out.splitlines() if ln.strip()][:60]
if args.out:
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)
