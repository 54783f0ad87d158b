"""CPU test of the multi-process path (world_size 2, gloo): each rank scans its shard of a synthetic
data set — here with the oracle, since there is no GPU — and rank-ordered merging reproduces the
single-process result. The GPU run uses exactly this sharding and merging (bench.py)."""
import os
import socket
import subprocess
import sys
import textwrap
from pathlib import Path

import numpy as np

from merkurio_b200.shard import merge_flags, merge_hits, shard_range

ROOT = Path(__file__).resolve().parent.parent


def test_shard_ranges_partition():
    for n in (0, 1, 63, 64, 65, 1000, 100_000_000):
        for w in (1, 2, 3, 4, 8):
            r = [shard_range(n, w, k) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            for a, b in zip(r, r[1:]):
                assert a[1] == b[0]
            assert all(a[0] % 64 == 0 or a[0] == a[1] for a in r)  # non-empty shards start on a flag word


def test_merge_flags_and_hits():
    f = merge_flags([(64, np.array([5], dtype=np.uint64)), (0, np.array([1], dtype=np.uint64))], 100)
    assert f.tolist() == [1, 5]
    dt = np.dtype([("record", "<u4"), ("start", "<u4"), ("pattern", "<u4"), ("len", "<u4")])
    a = np.array([(0, 1, 2, 3)], dtype=dt)
    b = np.array([(1, 4, 5, 6)], dtype=dt)
    m = merge_hits([(64, b), (0, a)])
    assert m["record"].tolist() == [0, 65]


WORKER = textwrap.dedent("""
    import sys, pickle
    sys.path.insert(0, {root!r})
    import numpy as np
    from merkurio_b200 import patterns as pt
    from merkurio_b200.shard import Dist, shard_range, merge_flags, merge_hits
    from merkurio_b200.synth import Synth
    from oracle import refmodel as rm
    d = Dist("gloo")
    n, L = 20000, 150
    syn = Synth(0x5EED0002, n, L, 31, 200)
    pats = pt.parse_pattern_list(syn.query_list(), reverse_complement_=True)
    ac = rm.AhoCorasick(pats)
    lo, hi = shard_range(n, d.world, d.rank)
    seq, off = syn.host_reads(lo, hi)
    flags, nrec, _ = ac.scan_batch(seq, off, 1, False)
    rec, st, pat = ac.batch_hits(seq, off)
    hits = np.zeros(len(rec), dtype=[("record", "<u4"), ("start", "<u4"), ("pattern", "<u4"), ("len", "<u4")])
    hits["record"], hits["start"], hits["pattern"] = rec, st, pat
    d.barrier()
    t = d.max(float(d.rank + 1))
    total = d.sum(float(nrec))
    parts = d.gather_objects((lo, flags, hits))
    if d.rank == 0:
        assert t == float(d.world)
        merged = merge_flags([(p[0], p[1]) for p in parts], n)
        mh = merge_hits([(p[0], p[2]) for p in parts])
        seq_all, off_all = syn.host_reads(0, n)
        f_all, n_all, _ = ac.scan_batch(seq_all, off_all, 1, False)
        r_all, s_all, p_all = ac.batch_hits(seq_all, off_all)
        assert np.array_equal(merged, f_all) and total == n_all
        assert np.array_equal(mh["record"], r_all) and np.array_equal(mh["start"], s_all) and np.array_equal(mh["pattern"], p_all)
        print("OK", int(total), len(mh))
    d.close()
""")


def test_two_ranks_gloo(tmp_path):
    from merkurio_b200.build import build_oracle, build_synth
    build_synth()
    build_oracle()
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=str(ROOT)))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK" in r.stdout
