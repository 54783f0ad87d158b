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


def _cfg5_checks(wl, expect_min_hits):
    """cfg5 through ONE engine (all queries, the L2-resident dual-key filter): report order, every hit re-verified,
    every query that still equals the text at its sampling position reported there, flags == records of the hit
    list, idempotence — and the hits inside the first 4 Mbp of record 0 against the oracle's Aho-Corasick automaton
    of the SAME full query set over that slice."""
    from merkurio_b200 import capi
    from oracle import checks
    with capi.Engine(wl.pats, n_slots=0, hit_capacity=wl.hit_capacity) as e:
        r = wl.scan(e, fetch=True)
        info = e.info()
        assert info.filter_in_smem[0] == 0 and info.seed_d[0] == 8  # the L2-resident filter, stride 8
        assert "dual-key" in e.scan_kernel(capi.MK_ENC_ASCII)
        hits = r.hits
        assert r.n_hits == len(hits) >= expect_min_hits
        key = (hits["record"].astype(np.uint64) << np.uint64(40)) | ((hits["start"].astype(np.uint64) + hits["len"]) << np.uint64(8)) | (np.uint64(255) - hits["len"].astype(np.uint64))
        assert np.all(key[1:] >= key[:-1])  # (record, end, longer pattern first)
        assert checks.reverify_hits(wl.d_seq, wl.d_off, hits, wl.pats, False)
        assert checks.genome_expected_found(r, wl) == 0
        assert np.array_equal(np.nonzero(_flag_bits(r.flags, 24))[0], np.unique(hits["record"]))
        r2 = wl.scan(e, fetch=True)
        assert np.array_equal(r2.hits, hits)  # idempotent, order included
        n_slice = checks.genome_slice_equal(r, wl, 4_000_000)  # same engine, same pattern set
        assert n_slice > 100
    return r


def test_cfg5_full_size_same_engine():
    """BASELINE cfg5 at full size: 3.0 Gbp in 24 records, ~980 k queries of 21-63 bases (upper case, 1 % with N)
    against text with 30 % lower-case soft-masked spans (which must not match) and 2 % N."""
    import torch
    from merkurio_b200.synth import workloads as wlm
    wl = wlm.genome_workload(1.0, upper_queries=True)
    assert len(wl.pats) > 900_000 and wl.n_units > 2_900_000_000
    r = _cfg5_checks(wl, 500_000)
    # upper-case queries never match inside a soft-masked span: no hit may hold a lower-case text byte
    h = r.hits[:: max(len(r.hits) // 200_000, 1)]
    st = torch.from_numpy((wl.extra["off"][h["record"]] + h["start"]).astype(np.int64)).cuda()
    ln = torch.from_numpy(h["len"].astype(np.int64)).cuda()
    first, last = wl.d_seq[st], wl.d_seq[st + ln - 1]
    assert bool(((first < 97) & (last < 97)).all().item())
    del wl
    torch.cuda.empty_cache()


def test_cfg5_verbatim_case_tenth_scale():
    """cfg5 at 1/10 scale with the queries kept as sampled (a query from a soft-masked span is lower case and matches
    there): lower-case and mixed-case patterns through the same path, the alphabet gate finding nothing to skip."""
    import torch
    from merkurio_b200.synth import workloads as wlm
    wl = wlm.genome_workload(0.1, upper_queries=False)
    _cfg5_checks(wl, 90_000)
    del wl
    torch.cuda.empty_cache()
