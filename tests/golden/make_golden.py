"""Bundle the reference's own test vectors into tests/golden/reference_vectors.json.xz.

Run here (the build container), where /root/reference is mounted; the GPU box has no reference
tree, so tests read the committed bundle instead. Only data files are bundled (fixtures, example
inputs and their golden outputs, benchmark query lists) — no reference source code.

    python tests/golden/make_golden.py
"""
import base64
import json
import lzma
from pathlib import Path

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent / "reference_vectors.json.xz"

GLOBS = [
    "tests/fixtures/**/*",
    "tests/data/*",
    "example-minimal/kmers.txt",
    "example-minimal/sample.fasta",
    "example-minimal/sample.sam",
    "example-workflow/data/mutant_R1.fastq",
    "example-workflow/data/mutant_R2.fastq",
    "example-workflow/data/significant_kmers.txt",
    "example-workflow/significant_kmers.txt",
    "example-workflow/logs/*",
    "example-workflow/output/*",
    "benchmarks/patterns/*",
]


def main():
    files = {}
    for g in GLOBS:
        for p in sorted(REF.glob(g)):
            if p.is_file():
                files[str(p.relative_to(REF))] = base64.b64encode(p.read_bytes()).decode()
    blob = json.dumps({"source": "lschoenm/MerKurio test vectors", "files": files}, sort_keys=True).encode()
    OUT.write_bytes(lzma.compress(blob, preset=9 | lzma.PRESET_EXTREME))
    print(f"{len(files)} files, {len(blob)} bytes -> {OUT.stat().st_size} bytes")


if __name__ == "__main__":
    main()
