"""Device-timed scan of the other BASELINE configs (parity-test cases; bench.py reports the same entries in its
`configs` array):

  cfg3  paired-end 2 x 150 bp, 50 M pairs, 10 000 canonical 31-mers, ALL_HITS (JSON log)
  cfg4  BAM 4-bit, 50 M x 150 bp, 10 000 31-mers, PATTERN_SET (km tag, keep-only-matching)
  cfg5  3 Gbp multi-chromosome FASTA, 1 M mixed-length queries (21-63, 1 % with N), ALL_HITS;
        cfg5_verbatim_case: the queries as sampled (lower case where the text is soft-masked)

    python scripts/bench_configs.py --config cfg3 [--steps 10] [--out x.json]

Each run checks a bounded sample against the oracle (bit-exact) and size-independent properties at full size
(see bench.py: bench_config)."""
import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--config", choices=["cfg3", "cfg4", "cfg5", "cfg5_verbatim_case"], required=True)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--out", default=None)
args = ap.parse_args()
peak, _ = bench.measured_peak_gbs()
out = bench.bench_config(args.config, args.steps, peak)
print(json.dumps(out))
if args.out:
    Path(args.out).write_text(json.dumps(out, indent=1) + "\n")
