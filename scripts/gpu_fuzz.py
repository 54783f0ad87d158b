"""Randomised parity runs of the device path against the oracle (needs a GPU): query sets and texts drawn with random
lengths, alphabets, case modes, record shapes and filter flavours, each checked in ALL_HITS, PATTERN_SET and FLAG mode
with tests/test_gpu_parity.check_batch / check_bam4 (bit-exact hit lists; a third of the ACGT(N) cases as BAM 4-bit text). A soak test beside the fixed-seed cases of the suite.

    python scripts/gpu_fuzz.py [--seconds 120] [--seed 1]
"""
import argparse
import os
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from tests.test_gpu_parity import check_bam4, check_batch, planted_records, rand_seq  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=120)
ap.add_argument("--seed", type=int, default=1)
args = ap.parse_args()
rng = np.random.default_rng(args.seed)
ENV = ("MK_FILTER_MODE", "MK_NO_FILTER32", "MK_NO_WIN_SCAN", "MK_NO_LONG_SEEDS", "MK_NO_DUAL8", "MK_NO_GATE", "MK_NO_DIRECT", "MK_NO_BUCKET_SORT")
t0 = time.time()
n = 0
while time.time() - t0 < args.seconds:
    for v in ENV:
        os.environ.pop(v, None)
    kmin = int(rng.choice([1, 2, 3, 5, 8, 11, 12, 13, 14, 15, 16, 18, 19, 21, 25, 30, 31, 32, 40, 64, 100]))
    kmax = kmin + int(rng.choice([0, 0, 1, 3, 10, 40]))
    alpha = [b"ACGT", b"ACGT", b"ACGTN", b"ACGTacgt", b"ACDEFGHIKLMNPQRSTVWY", b"AC"][int(rng.integers(6))]
    n_pat = int(rng.choice([1, 2, 7, 60, 400, 3000]))
    pats = sorted({rand_seq(rng, int(rng.integers(kmin, kmax + 1)), alpha) for _ in range(n_pat)})
    ci = bool(rng.random() < 0.25)
    shape = int(rng.integers(4))
    if shape == 0:
        recs = planted_records(rng, pats, int(rng.integers(1, 3000)), 0, 300, alpha, plant_p=0.3)
    elif shape == 1:
        recs = planted_records(rng, pats, int(rng.integers(1, 6)), 50_000, 400_000, alpha, plant_p=1.0)
    elif shape == 2:
        recs = planted_records(rng, pats, int(rng.integers(1, 500)), 150, 150, alpha, plant_p=0.5)
    else:  # low complexity: long runs of one letter, many overlapping hits for short patterns
        recs = [bytes([alpha[int(rng.integers(len(alpha)))]]) * int(rng.integers(1, 5000)) for _ in range(int(rng.integers(1, 40)))]
    if ci:
        recs = [r.lower() if rng.random() < 0.5 else r for r in recs]
    knobs = {}
    if rng.random() < 0.3:
        knobs["MK_FILTER_MODE"] = str(rng.choice(["l2", "smem"]))
    for v in ENV[1:]:
        if rng.random() < 0.1:
            knobs[v] = "1"
    os.environ.update(knobs)
    if kmin <= 3 and shape in (1, 3):
        recs = [r[:20000] for r in recs[:3]]  # (millions of hits otherwise)
    bam = (not ci) and alpha in (b"ACGT", b"ACGTN") and rng.random() < 0.35  # BAM 4-bit text (upper-case IUPAC only)
    try:
        if bam:
            check_bam4(pats, [r for r in recs] or [b""])
        else:
            check_batch(pats, recs, case_insensitive=ci)
    except Exception as ex:
        print(f"FAILED case {n}: seed {args.seed} kmin {kmin} kmax {kmax} alphabet {alpha} patterns {len(pats)} ci {ci} bam {bam} shape {shape} knobs {knobs}: {repr(ex)[:500]}", flush=True)
        raise
    n += 1
print(f"{n} random cases in {time.time() - t0:.0f} s: all equal to the oracle")
