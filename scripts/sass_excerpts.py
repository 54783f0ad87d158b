"""SASS evidence for the hot kernels (no GPU needed): for each kernel of libmerkurio_cuda.so named below, registers /
spills / shared memory from `cuobjdump -res-usage`, an opcode histogram of its hottest loop (the innermost backward
branch that contains the streaming 16-byte loads, or the largest loop for the verify kernel) and that loop's SASS.

    python scripts/sass_excerpts.py            # writes profiles/sass/<kernel>.txt
"""
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "merkurio_b200" / "lib" / "libmerkurio_cuda.so"
OUT = ROOT / "profiles" / "sass"
KERNELS = {
    # file stem: (substring of the mangled name, note)
    "mk_scan_d16_ascii": ("mk_scan_d16ILi0ELi0ELi4ELi896ELb0ELb1E", "cfg2: ASCII, stride 16, shared-memory filter with 32-bit blocks, U=4, T=896"),
    "mk_scan_d16_ascii_f64": ("mk_scan_d16ILi0ELi0ELi4ELi896ELb0ELb0E", "cfg3: ASCII, stride 16, shared-memory filter with 64-bit blocks"),
    "mk_scan_d16_bam4": ("mk_scan_d16ILi1ELi0ELi4ELi768ELb0ELb0E", "cfg4: BAM 4-bit, stride 16, two seeds per 16-byte vector, T=768"),
    "mk_scan_win_ascii_d8": ("mk_scan_winILi0ELi8ELi4ELi896ELb1E", "k = 19..30, small query sets: window seeds, stride 8"),
    "mk_scan_dual8_gate": ("mk_scan_dual8ILi2ELi1024ELb1E", "cfg5: stride 8, L2-resident dual-key filter, alphabet gate, U=2, T=1024"),
    "mk_verify_candidates_ascii": ("mk_verify_candidatesILi0ELi128E", "candidate verification, one candidate per thread, CTAs of 128 threads"),
}


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", str(LIB)], capture_output=True, text=True, check=True).stdout
    funcs = {}
    cur = None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur is not None:
            funcs[cur].append(ln)
    usage = {}
    lines = res.splitlines()
    for i, ln in enumerate(lines):
        m = re.match(r"\s*Function (\S+):", ln)
        if m and i + 1 < len(lines):
            usage[m.group(1)] = lines[i + 1].strip()
    for stem, (key, note) in KERNELS.items():
        name = next((f for f in funcs if key in f), None)
        if name is None:
            print(f"{stem}: no function matching {key}", file=sys.stderr)
            continue
        ins = []  # (addr, text)
        for ln in funcs[name]:
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
        strip = lambda t: re.sub(r"^@!?U?P\d+\s+", "", t)
        op = lambda t: strip(t).split()[0]
        # the tile body: from the first streaming load (ld.global.nc.L1::no_allocate -> LDG.E.NA.128) on; the verify
        # kernel from its first global load
        first = next((i for i, (_, t) in enumerate(ins) if op(t).startswith("LDG") and ".NA." in op(t)), None)
        if first is None:
            first = next((i for i, (_, t) in enumerate(ins) if op(t).startswith("LDG")), 0)
        b = ins[first:first + 260]
        ops_all = Counter(op(t).split(".")[0] for _, t in ins)
        mem_all = Counter(op(t) for _, t in ins if re.match(r"(LDG|LDS|STS|STG|ATOMS|ATOMG|RED|LDL|STL|LD|ST)$", op(t).split(".")[0]))
        with open(OUT / f"{stem}.txt", "w") as f:
            f.write(f"{name}\n{note}\nresources: {usage.get(name, '?')}  (STACK / LOCAL 0 = no spills)\n")
            f.write(f"whole kernel: {len(ins)} instructions (both copies of the double-buffered tile body, the ragged tile and the queue paths)\n")
            f.write("opcodes: " + ", ".join(f"{k} {v}" for k, v in ops_all.most_common(24)) + "\n")
            f.write("memory instructions: " + ", ".join(f"{k} x{v}" for k, v in mem_all.most_common()) + "\n")
            f.write(f"\nexcerpt: {len(b)} instructions from the first streaming load of the tile body (0x{b[0][0]:x})\n")
            for a_, t in b:
                f.write(f"/*{a_:04x}*/ {t}\n")
        print(f"{stem}: {usage.get(name, '?')}; " + ", ".join(f"{k} x{v}" for k, v in mem_all.most_common(8)))


if __name__ == "__main__":
    main()
