"""profiles/ncu_traffic.json from an `ncu --set full` capture of bench.py's scan kernel: the DRAM bytes of one launch,
keyed by the digest of the CUDA sources the capture was taken from (bench.py reports `roofline.traffic` only while that
digest matches what is running; otherwise null with the reason).

    python scripts/ncu_traffic.py gpurun_out/prof_cfg2.ncu-rep mk_scan_d16 cfg2:100000000x150:1000 profiles/r2_ncu_full_cfg2.txt
"""
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

rep, kernel, key, source = sys.argv[1:5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
row = next(r for r in data if kernel in r[hdr.index("Kernel Name")])


def metric(name):
    i = hdr.index(name)
    v, u = float(row[i]), units[i]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]


p = ROOT / "profiles" / "ncu_traffic.json"
d = json.loads(p.read_text()) if p.exists() else {"kernels": {}}
d["kernels"].setdefault(kernel, {})[key] = {
    "dram_bytes_read": int(metric("dram__bytes_read.sum")), "dram_bytes_write": int(metric("dram__bytes_write.sum")),
    "kernel_name": row[hdr.index("Kernel Name")], "duration_ms_under_ncu": float(row[hdr.index("gpu__time_duration.sum")]) * {"ms": 1, "us": 1e-3, "s": 1e3}[units[hdr.index("gpu__time_duration.sum")]],
    "csrc_sha16": bench.csrc_digest(), "source": source}
p.write_text(json.dumps(d, indent=1) + "\n")
print(json.dumps(d["kernels"][kernel][key]))
