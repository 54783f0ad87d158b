"""Device-timed scan of the other BASELINE configs (parity-test cases, not bench.py lines):

  cfg3  paired-end 2 x 150 bp, 50 M pairs, 10 000 canonical 31-mers, ALL_HITS (JSON log)
  cfg4  BAM 4-bit, 50 M x 150 bp, 10 000 31-mers, PATTERN_SET (km tag, keep-only-matching)
  cfg5  3 Gbp multi-chromosome FASTA, 1 M mixed-length queries (21-63, 1 % with N), ALL_HITS

    python scripts/bench_configs.py --config cfg3 [--scale 0.1] [--check-reads 200000]

Each run checks a bounded sample against the oracle (bit-exact hit lists) and, at full size,
size-independent properties (every reported hit re-verified by direct byte comparison on the GPU;
for cfg5 every clean sampled query found at its sampling position)."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from merkurio_b200 import capi, patterns as pt
from merkurio_b200.synth import Synth
from oracle import refmodel as rm

ap = argparse.ArgumentParser()
ap.add_argument("--config", choices=["cfg3", "cfg4", "cfg5"], required=True)
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--check-reads", type=int, default=200_000)
ap.add_argument("--out", default=None)
args = ap.parse_args()
PEAK = 6554.2
p = Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json"
if p.exists():
    PEAK = float(json.loads(p.read_text())["hbm_gbs"])


def timed(engine, fn, steps):
    for _ in range(3):
        r = fn(False)
    scan, dev = [], []
    for _ in range(steps):
        r = fn(False)
        scan.append(r.scan_ns)
        dev.append(r.device_ns)
    return float(np.median(scan)) / 1e6, float(np.median(dev)) / 1e6, r


def reverify_hits(d_seq, d_off, hits, pats, enc):
    """Every reported hit, re-checked by direct comparison of text and pattern (on the GPU)."""
    if len(hits) == 0:
        return True
    lens = np.array([len(x) for x in pats], dtype=np.int64)
    maxlen = int(lens.max())
    pat_mat = np.zeros((len(pats), maxlen), dtype=np.uint8)
    for i, x in enumerate(pats):
        pat_mat[i, :len(x)] = np.frombuffer(x, dtype=np.uint8)
    pm = torch.from_numpy(pat_mat).cuda()
    pl = torch.from_numpy(lens).cuda()
    ok = True
    B = 2_000_000
    dec = torch.from_numpy(np.frombuffer(rm.NIBBLE_CHARS, dtype=np.uint8).copy()).cuda()
    for s in range(0, len(hits), B):
        h = hits[s:s + B]
        rec = torch.from_numpy(h["record"].astype(np.int64)).cuda()
        st = torch.from_numpy(h["start"].astype(np.int64)).cuda()
        pid = torch.from_numpy(h["pattern"].astype(np.int64)).cuda()
        base = d_off[rec] + st
        idx = base[:, None] + torch.arange(maxlen, device="cuda")[None, :]
        valid = torch.arange(maxlen, device="cuda")[None, :] < pl[pid][:, None]
        if enc == capi.MK_ENC_ASCII:
            idx = idx.clamp_(max=d_seq.numel() - 1)
            txt = d_seq[idx]
        else:
            b = d_seq[(idx >> 1).clamp_(max=d_seq.numel() - 1)]
            txt = dec[torch.where(idx & 1 == 1, b & 15, b >> 4).long()]
        ok = ok and bool(((txt == pm[pid]) | ~valid).all().item())
        # and inside the record
        ok = ok and bool((st + pl[pid] <= d_off[rec + 1] - d_off[rec]).all().item())
    return ok


def oracle_sample_check(eng, syn, pats, n_check, enc, mode, L):
    """Bit-exact comparison with the oracle on the first n_check reads."""
    h_seq, h_off = syn.host_reads(0, n_check, 0)
    ac = rm.AhoCorasick(pats)
    rec, st, pat = ac.batch_hits(h_seq, h_off)
    if enc == capi.MK_ENC_ASCII:
        d = torch.from_numpy(np.concatenate([h_seq, np.zeros(64, np.uint8)])).cuda()
        units = n_check * L
    else:
        p4, _ = syn.host_reads(0, n_check, 1)
        d = torch.from_numpy(np.concatenate([p4, np.zeros(64, np.uint8)])).cuda()
        units = n_check * L
    o = torch.from_numpy(h_off.astype(np.int64)).cuda()
    r = eng.scan_device(d.data_ptr(), o.data_ptr(), n_check, units, mode, enc, fetch=True)
    if mode == capi.MK_MODE_ALL_HITS:
        assert np.array_equal(r.hits["record"], rec) and np.array_equal(r.hits["start"], st) and np.array_equal(r.hits["pattern"], pat), "hit list differs from the oracle"
    else:
        pairs = sorted(set(zip(rec.tolist(), pat.tolist())))
        assert list(zip(r.hits["record"].tolist(), r.hits["pattern"].tolist())) == pairs, "pattern sets differ from the oracle"
    assert np.array_equal(r.flagged_records(), np.unique(rec)), "flags differ from the oracle"
    return len(rec)


out = {"config": args.config, "scale": args.scale}
if args.config in ("cfg3", "cfg4"):
    L = 150
    n = int((100_000_000 if args.config == "cfg3" else 50_000_000) * args.scale) // 64 * 64
    seed = 0x5EED0003 if args.config == "cfg3" else 0x5EED0004
    syn = Synth(seed, n, L, 31, 10000)
    if args.config == "cfg3":
        pats = pt.parse_pattern_list(syn.query_list(), canonical_=True)
        enc, mode = capi.MK_ENC_ASCII, capi.MK_MODE_ALL_HITS
        nbytes = n * L
    else:
        pats = pt.parse_pattern_list(syn.query_list())
        enc, mode = capi.MK_ENC_BAM4, capi.MK_MODE_PATTERN_SET
        nbytes = n * L // 2
    d_seq = torch.empty(nbytes + 64, dtype=torch.uint8, device="cuda")
    d_off = torch.empty(n + 1, dtype=torch.int64, device="cuda")
    d_q = torch.from_numpy(syn.queries).cuda()
    syn.device_reads(d_q.data_ptr(), 0, n, d_seq.data_ptr(), d_off.data_ptr(), 1 if enc else 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    with capi.Engine(pats, n_slots=0, hit_capacity=max(n // 20, 1 << 20)) as eng:
        checked = oracle_sample_check(eng, syn, pats, min(args.check_reads, n), enc, mode, L)
        fn = lambda fetch: eng.scan_device(d_seq.data_ptr(), d_off.data_ptr(), n, n * L, mode, enc, fetch=fetch)
        scan_ms, dev_ms, r = timed(eng, fn, args.steps)
        full = fn(True)
        info = eng.info()
        ok = reverify_hits(d_seq, d_off, full.hits, pats, enc) if mode == capi.MK_MODE_ALL_HITS else None
        if mode == capi.MK_MODE_ALL_HITS:
            key = full.hits["record"].astype(np.uint64) << np.uint64(20) | (full.hits["start"] + full.hits["len"]).astype(np.uint64)
            assert np.all(key[1:] >= key[:-1]), "hit list is not sorted"
        else:
            key = full.hits["record"].astype(np.uint64) << np.uint64(32) | full.hits["pattern"].astype(np.uint64)
            assert np.all(key[1:] > key[:-1]), "pattern sets are not sorted/unique"
        out.update(records=n, bases=n * L, patterns=len(pats), seeds=int(info.n_seeds[enc]), filter_bytes=int(info.filter_bytes[enc]),
                   scan_ms=scan_ms, device_ms=dev_ms, n_hits=int(full.n_hits), records_flagged=int(np.bitwise_count(full.flags).sum()),
                   oracle_checked_hits=checked, hits_reverified=ok, rescans=int(full.n_rescans),
                   verify_ms=r.verify_ns / 1e6, candidates=r.n_candidates, gbases_per_s=n * L / dev_ms / 1e6, scan_gbs=nbytes / scan_ms / 1e6, frac_of_peak=nbytes / scan_ms / 1e6 / PEAK)
else:
    # cfg5: 24 chromosomes, lengths proportional to hg38, total 3 Gbp; 2 % N runs; 30 % lower-case
    total = int(3_000_000_000 * args.scale)
    hg38 = [248.9, 242.2, 198.3, 190.2, 181.5, 170.8, 159.3, 145.1, 138.4, 133.8, 135.1, 133.3, 114.4, 107.0, 102.0, 90.3, 83.3, 80.4, 58.6, 64.4, 46.7, 50.8, 156.0, 57.2]
    lens = np.array([int(x / sum(hg38) * total) for x in hg38], dtype=np.int64)
    total = int(lens.sum())
    off = np.zeros(25, dtype=np.int64)
    off[1:] = np.cumsum(lens)
    g = torch.Generator(device="cuda").manual_seed(0x5EED0005)
    d_seq = torch.empty(total + 64, dtype=torch.uint8, device="cuda")
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device="cuda")
    CH = 1 << 28
    for s in range(0, total, CH):
        e = min(total, s + CH)
        d_seq[s:e] = lut[torch.randint(0, 4, (e - s,), generator=g, device="cuda", dtype=torch.uint8).long()]
    rng = np.random.default_rng(5)
    # lower-case (soft-masked) spans: 30 % of the length, N runs: 2 %
    pos = 0
    n_low = n_N = 0
    while pos < total:
        span = int(rng.integers(2000, 200000))
        kind = rng.random()
        e = min(total, pos + span)
        if kind < 0.30:
            d_seq[pos:e] |= 0x20
            n_low += e - pos
        elif kind < 0.32:
            d_seq[pos:e] = 78
            n_N += e - pos
        pos = e
    d_seq[total:] = 0
    # queries: 1 M, length uniform 21..63, sampled inside chromosomes
    nq = int(1_000_000 * min(1.0, max(args.scale, 0.02)))
    ql = rng.integers(21, 64, size=nq)
    chrom = rng.choice(24, size=nq, p=lens / lens.sum())
    qs = off[chrom] + (rng.random(nq) * (lens[chrom] - ql)).astype(np.int64)
    idx = torch.from_numpy(qs).cuda()[:, None] + torch.arange(63, device="cuda")[None, :]
    qmat = d_seq[idx.clamp_(max=total - 1)].cpu().numpy()
    # drop samples that fell inside an N run (more than 2 N); 1-2 N at a run edge stay: they occur verbatim
    n_count = ((qmat == 78) & (np.arange(63)[None, :] < ql[:, None])).sum(axis=1)
    keep = n_count <= 2
    qmat, ql, chrom, qs = qmat[keep], ql[keep], chrom[keep], qs[keep]
    nq = int(keep.sum())
    queries = [qmat[i, :ql[i]].tobytes() for i in range(nq)]
    # 1 % of the queries get one base replaced by N (these can only hit where the text has that N)
    for i in rng.choice(nq, size=nq // 100, replace=False):
        b = bytearray(queries[i])
        b[int(rng.integers(len(b)))] = 78
        queries[i] = bytes(b)
    t0 = time.perf_counter()
    pats = pt.parse_pattern_list(queries)
    d_off = torch.from_numpy(off).cuda()
    with capi.Engine(pats, n_slots=0, hit_capacity=4 * nq) as eng:
        t_build0 = time.perf_counter()
        fn = lambda fetch: eng.scan_device(d_seq.data_ptr(), d_off.data_ptr(), 24, total, capi.MK_MODE_ALL_HITS, capi.MK_ENC_ASCII, fetch=fetch)
        first = fn(False)
        t_build = time.perf_counter() - t_build0
        scan_ms, dev_ms, r = timed(eng, fn, args.steps)
        full = fn(True)
        info = eng.info()
        ok = reverify_hits(d_seq, d_off, full.hits, pats, capi.MK_ENC_ASCII)
        # every query must be reported at the position it was sampled from unless it was altered
        pid_of = {p_: i for i, p_ in enumerate(pats)}
        hs = set(zip((full.hits["record"].astype(np.int64)).tolist(), full.hits["start"].tolist(), full.hits["pattern"].tolist()))
        missing = 0
        for i in range(0, nq, max(nq // 20000, 1)):
            q = queries[i]
            if qmat[i, :ql[i]].tobytes() != q:
                continue  # N-substituted
            if (int(chrom[i]), int(qs[i] - off[chrom[i]]), pid_of[q]) not in hs:
                missing += 1
        # oracle parity on a slice: first 4 Mbp of chromosome 1 with the queries sampled from it
        sl = min(4_000_000, int(lens[0]))
        sub = [q for q, c, s_ in zip(queries, chrom, qs) if c == 0 and s_ + 63 < sl]
        sub_p = pt.parse_pattern_list(sub) if sub else None
        checked = None
        if sub_p:
            text = d_seq[:sl].cpu().numpy()
            rec, st, pat = rm.AhoCorasick(sub_p).batch_hits(text, np.array([0, sl], dtype=np.uint64))
            with capi.Engine(sub_p, n_slots=0) as e2:
                o2 = torch.tensor([0, sl], dtype=torch.int64, device="cuda")
                r2 = e2.scan_device(d_seq.data_ptr(), o2.data_ptr(), 1, sl, capi.MK_MODE_ALL_HITS, capi.MK_ENC_ASCII, fetch=True)
            assert np.array_equal(r2.hits["start"], st) and np.array_equal(r2.hits["pattern"], pat), "slice differs from the oracle"
            checked = len(st)
        out.update(records=24, bases=total, patterns=len(pats), min_len=int(info.min_len), max_len=int(info.max_len), seed_q=int(info.seed_q[0]),
                   seed_d=int(info.seed_d[0]), seeds=int(info.n_seeds[0]), filter_in_smem=int(info.filter_in_smem[0]), filter_bytes=int(info.filter_bytes[0]),
                   table_bytes=int(info.table_bytes[0]), table_build_s=t_build, scan_ms=scan_ms, device_ms=dev_ms, n_hits=int(full.n_hits),
                   hits_reverified=ok, sampled_queries_missing=missing, oracle_slice_hits=checked, lower_case_frac=n_low / total, n_frac=n_N / total,
                   verify_ms=r.verify_ns / 1e6, candidates=r.n_candidates, gbases_per_s=total / dev_ms / 1e6, scan_gbs=total / scan_ms / 1e6, frac_of_peak=total / scan_ms / 1e6 / PEAK)
print(json.dumps(out))
if args.out:
    Path(args.out).write_text(json.dumps(out, indent=1) + "\n")
