"""BASELINE configs at their FULL sizes on the device, checked through size-independent properties
(the oracle would need minutes for them): every reported hit re-verified by direct byte comparison,
report order, flags == records of the hit list, idempotence, the result of the whole equal to the
results of its halves, ASCII and BAM4 packings of the same reads agreeing — plus bit-exact parity
with the oracle on a sample of the same data. Needs ~20 GB of device memory (a B200 has 180)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

L = 150


def _device_reads(syn, n, enc):
    import torch
    nbytes = n * L // (2 if enc else 1)
    d_seq = torch.empty(nbytes + 64, dtype=torch.uint8, device="cuda")
    d_off = torch.empty(n + 1, dtype=torch.int64, device="cuda")
    d_q = torch.from_numpy(syn.queries).cuda()
    syn.device_reads(d_q.data_ptr(), 0, n, d_seq.data_ptr(), d_off.data_ptr(), enc, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    d_seq[nbytes:] = 0
    return d_seq, d_off


def _reverify(d_seq, d_off, hits, pats, bam4):
    """Every hit: text[start, start+len) == pattern, inside its record (gathers on the GPU)."""
    import torch
    from oracle import refmodel as rm
    if len(hits) == 0:
        return
    lens = np.array([len(x) for x in pats], dtype=np.int64)
    width = int(lens.max())
    pm = np.zeros((len(pats), width), dtype=np.uint8)
    for i, x in enumerate(pats):
        pm[i, :len(x)] = np.frombuffer(x, dtype=np.uint8)
    pm, pl = torch.from_numpy(pm).cuda(), torch.from_numpy(lens).cuda()
    dec = torch.from_numpy(np.frombuffer(rm.NIBBLE_CHARS, dtype=np.uint8).copy()).cuda()
    col = torch.arange(width, device="cuda")[None, :]
    for s in range(0, len(hits), 1_000_000):
        h = hits[s:s + 1_000_000]
        rec = torch.from_numpy(h["record"].astype(np.int64)).cuda()
        st = torch.from_numpy(h["start"].astype(np.int64)).cuda()
        pid = torch.from_numpy(h["pattern"].astype(np.int64)).cuda()
        idx = (d_off[rec] + st)[:, None] + col
        if bam4:
            b = d_seq[(idx >> 1).clamp_(max=d_seq.numel() - 1)]
            txt = dec[torch.where(idx & 1 == 1, b & 15, b >> 4).long()]
        else:
            txt = d_seq[idx.clamp_(max=d_seq.numel() - 1)]
        valid = col < pl[pid][:, None]
        assert bool(((txt == pm[pid]) | ~valid).all().item()), "a reported hit does not match its pattern"
        assert bool((st + pl[pid] <= d_off[rec + 1] - d_off[rec]).all().item()), "a reported hit leaves its record"


def _flag_bits(flags, n):
    return np.unpackbits(flags.view(np.uint8), bitorder="little")[:n]


def test_cfg2_cfg3_full_size_properties():
    """100 M x 150 bp reads (15 GB): FLAG (cfg2: extract without logs) and ALL_HITS (cfg3: logs)."""
    import torch
    from merkurio_b200 import capi, patterns as pt
    from merkurio_b200.synth import Synth
    from oracle import refmodel as rm
    n = 100_000_000
    syn = Synth(0x5EED0002, n, L, 31, 1000)
    pats = pt.parse_pattern_list(syn.query_list(), reverse_complement_=True)
    d_seq, d_off = _device_reads(syn, n, 0)
    with capi.Engine(pats, n_slots=0, hit_capacity=1 << 22) as e:
        f1 = e.scan_device(d_seq.data_ptr(), d_off.data_ptr(), n, n * L, capi.MK_MODE_FLAG, fetch=True)
        f2 = e.scan_device(d_seq.data_ptr(), d_off.data_ptr(), n, n * L, capi.MK_MODE_FLAG, fetch=True)
        assert np.array_equal(f1.flags, f2.flags)  # idempotent
        a = e.scan_device(d_seq.data_ptr(), d_off.data_ptr(), n, n * L, capi.MK_MODE_ALL_HITS, fetch=True)
        hits = a.hits
        assert a.n_hits == len(hits) > 900_000  # ~1 % of the reads carry a planted query
        # report order: (record, end, start) ascending; len == pattern length
        key = (hits["record"].astype(np.uint64) << np.uint64(20)) | ((hits["start"] + hits["len"]).astype(np.uint64) << np.uint64(10)) | hits["start"].astype(np.uint64)
        assert np.all(key[1:] >= key[:-1])
        assert np.array_equal(hits["len"], np.array([len(p) for p in pats], dtype=np.uint32)[hits["pattern"]])
        # the flag bitmap is exactly the set of records of the hit list, in both modes
        bits = _flag_bits(f1.flags, n)
        assert int(bits.sum()) == len(np.unique(hits["record"]))
        assert bits[hits["record"]].all()
        assert np.array_equal(a.flags, f1.flags)
        _reverify(d_seq, d_off, hits, pats, False)
        # the whole equals its halves (records are independent: what sharding over GPUs relies on)
        half = n // 2
        lo = e.scan_device(d_seq.data_ptr(), d_off.data_ptr(), half, half * L, capi.MK_MODE_FLAG, fetch=True)
        off_hi = (d_off[half:] - d_off[half]).contiguous()
        hi = e.scan_device(d_seq.data_ptr() + half * L, off_hi.data_ptr(), n - half, (n - half) * L, capi.MK_MODE_FLAG, fetch=True)
        assert np.array_equal(np.concatenate([_flag_bits(lo.flags, half), _flag_bits(hi.flags, n - half)]), bits)
        # bit-exact against the oracle on a sample of the same data
        m = 300_000
        h_seq, h_off = syn.host_reads(0, m)
        rec, st, pat = rm.AhoCorasick(pats).batch_hits(h_seq, h_off)
        sel = hits["record"] < m
        assert np.array_equal(hits["record"][sel], rec) and np.array_equal(hits["start"][sel], st) and np.array_equal(hits["pattern"][sel], pat)
    del d_seq, d_off
    torch.cuda.empty_cache()


def test_cfg4_full_size_bam4_properties():
    """50 M x 150 bp alignments in BAM's 4-bit packing (3.75 GB), 10 000 31-mers, PATTERN_SET (km tag)."""
    import torch
    from merkurio_b200 import capi, patterns as pt
    from merkurio_b200.synth import Synth
    from oracle import refmodel as rm
    n = 50_000_000
    syn = Synth(0x5EED0004, n, L, 31, 10000)
    pats = pt.parse_pattern_list(syn.query_list())
    d4, d_off = _device_reads(syn, n, 1)
    with capi.Engine(pats, n_slots=0, hit_capacity=1 << 21) as e:
        p = e.scan_device(d4.data_ptr(), d_off.data_ptr(), n, n * L, capi.MK_MODE_PATTERN_SET, capi.MK_ENC_BAM4, fetch=True)
        pairs = p.hits
        key = (pairs["record"].astype(np.uint64) << np.uint64(32)) | pairs["pattern"].astype(np.uint64)
        assert np.all(key[1:] > key[:-1])  # sorted, no duplicate (record, pattern)
        bits = _flag_bits(p.flags, n)
        assert int(bits.sum()) == len(np.unique(pairs["record"])) > 200_000
        a = e.scan_device(d4.data_ptr(), d_off.data_ptr(), n, n * L, capi.MK_MODE_ALL_HITS, capi.MK_ENC_BAM4, fetch=True)
        _reverify(d4, d_off, a.hits, pats, True)
        got = np.unique((a.hits["record"].astype(np.uint64) << np.uint64(32)) | a.hits["pattern"].astype(np.uint64))
        assert np.array_equal(got, key)  # the pattern sets are the distinct (record, pattern) of the hit list
        # the ASCII packing of the first 20 M of the same reads gives the same hits
        m = 20_000_000
        d1, d_off1 = _device_reads(Synth(0x5EED0004, n, L, 31, 10000), m, 0)
        b = e.scan_device(d1.data_ptr(), d_off1.data_ptr(), m, m * L, capi.MK_MODE_ALL_HITS, capi.MK_ENC_ASCII, fetch=True)
        sel = a.hits["record"] < m
        for f in ("record", "start", "pattern"):
            assert np.array_equal(a.hits[f][sel], b.hits[f])
        # oracle sample
        k = 200_000
        h_seq, h_off = syn.host_reads(0, k)
        rec, st, pat = rm.AhoCorasick(pats).batch_hits(h_seq, h_off)
        sel = a.hits["record"] < k
        assert np.array_equal(a.hits["record"][sel], rec) and np.array_equal(a.hits["start"][sel], st) and np.array_equal(a.hits["pattern"][sel], pat)
    del d4, d1
    torch.cuda.empty_cache()


def test_cfg5_shape_properties():
    """cfg5 at 1/10 scale (300 Mbp in 24 records, 100 k queries of 21-63 bases, 1 % with N, soft-masked
    and N spans in the text): the L2-resident dual-key filter path with records far larger than a tile."""
    import torch
    from merkurio_b200 import capi, patterns as pt
    from oracle import refmodel as rm
    rng = np.random.default_rng(5)
    total = 300_000_000
    lens = (rng.dirichlet(np.ones(24) * 3) * total).astype(np.int64) + 1000
    total = int(lens.sum())
    off = np.zeros(25, dtype=np.int64)
    off[1:] = np.cumsum(lens)
    g = torch.Generator(device="cuda").manual_seed(55)
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device="cuda")
    d_seq = torch.zeros(total + 64, dtype=torch.uint8, device="cuda")
    d_seq[:total] = lut[torch.randint(0, 4, (total,), generator=g, device="cuda", dtype=torch.uint8).long()]
    pos = 0
    while pos < total:  # 30 % lower case, 2 % N
        e = min(total, pos + int(rng.integers(2000, 200000)))
        kind = rng.random()
        if kind < 0.30:
            d_seq[pos:e] |= 0x20
        elif kind < 0.32:
            d_seq[pos:e] = 78
        pos = e
    nq = 100_000
    ql = rng.integers(21, 64, size=nq)
    chrom = rng.choice(24, size=nq, p=lens / lens.sum())
    qs = off[chrom] + (rng.random(nq) * (lens[chrom] - ql)).astype(np.int64)
    idx = torch.from_numpy(qs).cuda()[:, None] + torch.arange(63, device="cuda")[None, :]
    qmat = d_seq[idx.clamp_(max=total - 1)].cpu().numpy()
    keep = ((qmat == 78) & (np.arange(63)[None, :] < ql[:, None])).sum(axis=1) <= 2
    qmat, ql, chrom, qs = qmat[keep], ql[keep], chrom[keep], qs[keep]
    queries = [qmat[i, :ql[i]].tobytes() for i in range(len(ql))]
    altered = set(rng.choice(len(queries), size=len(queries) // 100, replace=False).tolist())
    for i in altered:
        b = bytearray(queries[i])
        b[int(rng.integers(len(b)))] = 78
        queries[i] = bytes(b)
    pats = pt.parse_pattern_list(queries)
    pid_of = {p: i for i, p in enumerate(pats)}
    d_off = torch.from_numpy(off).cuda()
    with capi.Engine(pats, n_slots=0, hit_capacity=1 << 19) as e:
        r = e.scan_device(d_seq.data_ptr(), d_off.data_ptr(), 24, total, capi.MK_MODE_ALL_HITS, fetch=True)
        info = e.info()
        assert info.filter_in_smem[0] == 0 and info.seed_d[0] == 8  # the L2-resident filter, stride 8
        hits = r.hits
        key = (hits["record"].astype(np.uint64) << np.uint64(40)) | ((hits["start"].astype(np.uint64) + hits["len"]) << np.uint64(8)) | (np.uint64(255) - hits["len"].astype(np.uint64))
        assert np.all(key[1:] >= key[:-1])  # (record, end, longer pattern first)
        _reverify(d_seq, d_off, hits, pats, False)
        found = set(zip(hits["record"].tolist(), hits["start"].tolist(), hits["pattern"].tolist()))
        for i, q in enumerate(queries):
            if i not in altered:
                assert (int(chrom[i]), int(qs[i] - off[chrom[i]]), pid_of[q]) in found
        assert np.array_equal(np.nonzero(_flag_bits(r.flags, 24))[0], np.unique(hits["record"]))
        # oracle parity on the first 3 Mbp of record 0 with the queries sampled from it
        sl = min(3_000_000, int(lens[0]))
        sub = pt.parse_pattern_list([q for q, c, s_ in zip(queries, chrom, qs) if c == 0 and s_ + 63 < sl])
    text = d_seq[:sl].cpu().numpy()
    rec, st, pat = rm.AhoCorasick(sub).batch_hits(text, np.array([0, sl], dtype=np.uint64))
    with capi.Engine(sub, n_slots=0) as e2:
        o2 = torch.tensor([0, sl], dtype=torch.int64, device="cuda")
        r2 = e2.scan_device(d_seq.data_ptr(), o2.data_ptr(), 1, sl, capi.MK_MODE_ALL_HITS, fetch=True)
    assert len(st) > 100 and np.array_equal(r2.hits["start"], st) and np.array_equal(r2.hits["pattern"], pat)
    del d_seq
    torch.cuda.empty_cache()
