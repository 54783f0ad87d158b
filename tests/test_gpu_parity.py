"""GPU parity tests: the CUDA path, called through the C ABI (include/merkurio_cuda.h), against the
CPU oracle on the same seeded inputs. Bar: bit-exact hit lists, in the reference's report order."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from merkurio_b200 import capi
from oracle import refmodel as rm

NIB = rm.NIBBLE_CHARS


def pack_records(records):
    off = np.zeros(len(records) + 1, dtype=np.uint64)
    if records:
        off[1:] = np.cumsum([len(r) for r in records])
    seq = np.frombuffer(b"".join(records), dtype=np.uint8).copy() if off[-1] else np.zeros(0, dtype=np.uint8)
    return seq, off


def oracle_hits(patterns, records, case_insensitive=False):
    ac = rm.AhoCorasick(patterns, case_insensitive)
    seq, off = pack_records(records)
    rec, st, pat = ac.batch_hits(seq if seq.size else np.zeros(1, np.uint8), off)
    return rec, st, pat


def check_batch(patterns, records, case_insensitive=False, hit_capacity=0, engine=None):
    """ALL_HITS, PATTERN_SET and FLAG of one batch against the oracle."""
    patterns = sorted(set(patterns))
    rec, st, pat = oracle_hits(patterns, records, case_insensitive)
    seq, off = pack_records(records)
    own = engine is None
    if own:
        engine = capi.Engine(patterns, case_insensitive=case_insensitive, max_batch_bytes=max(int(seq.size), 16) + 64,
                             max_batch_records=max(len(records), 1), hit_capacity=hit_capacity)
    try:
        r = engine.scan(seq, off, capi.MK_MODE_ALL_HITS)
        assert r.n_hits == len(rec), (r.n_hits, len(rec))
        np.testing.assert_array_equal(r.hits["record"], rec)
        np.testing.assert_array_equal(r.hits["start"], st)
        np.testing.assert_array_equal(r.hits["pattern"], pat)
        np.testing.assert_array_equal(r.hits["len"], np.array([len(patterns[p]) for p in pat], dtype=np.uint32))
        want_flag = np.unique(rec)
        np.testing.assert_array_equal(r.flagged_records(), want_flag)
        f = engine.scan(seq, off, capi.MK_MODE_FLAG)
        np.testing.assert_array_equal(f.flagged_records(), want_flag)
        assert f.n_hits == 0
        p = engine.scan(seq, off, capi.MK_MODE_PATTERN_SET)
        pairs = sorted(set(zip(rec.tolist(), pat.tolist())))
        assert list(zip(p.hits["record"].tolist(), p.hits["pattern"].tolist())) == pairs
        np.testing.assert_array_equal(p.flagged_records(), want_flag)
        return r
    finally:
        if own:
            engine.close()


def rand_seq(rng, n, alphabet=b"ACGT"):
    return bytes(rng.choice(np.frombuffer(alphabet, dtype=np.uint8), size=n).tobytes())


def planted_records(rng, patterns, n_records, min_len, max_len, alphabet=b"ACGT", plant_p=0.3, noise=b"N"):
    recs = []
    for _ in range(n_records):
        n = int(rng.integers(min_len, max_len + 1))
        r = bytearray(rand_seq(rng, n, alphabet))
        if n and rng.random() < plant_p:
            for _k in range(int(rng.integers(1, 4))):
                p = patterns[int(rng.integers(len(patterns)))]
                if len(p) <= n:
                    s = int(rng.integers(0, n - len(p) + 1))
                    r[s:s + len(p)] = p
        if n and rng.random() < 0.2:
            s = int(rng.integers(0, n))
            e = min(n, s + int(rng.integers(1, 6)))
            r[s:e] = noise * (e - s)
        recs.append(bytes(r))
    return recs


# ------------------------------------------------------------------------------------------------
def test_version_and_info():
    assert "sm_100a" in capi.version()
    with capi.Engine([b"A" * 31, b"C" * 40]) as e:
        e.scan(np.frombuffer(b"ACGT" * 16, dtype=np.uint8), np.array([0, 64], dtype=np.uint64))
        info = e.info()
        assert info.n_patterns == 2 and info.min_len == 31 and info.max_len == 40
        assert info.seed_q[0] == 16 and info.seed_d[0] == 16 and info.sm_count > 0


def test_reference_toy_vectors():
    # src/pattern_matching.rs:353-392 through the device
    r = check_batch([b"abc"], [b"abcabcabc", b"xabcabcabcx", b"ab", b""])
    assert r.hits["start"].tolist() == [0, 3, 6, 1, 4, 7]
    # self-overlap and substring patterns
    r = check_batch([b"AAA"], [b"AAAAA"])
    assert r.hits["start"].tolist() == [0, 1, 2]
    r = check_batch([b"CTC", b"TC", b"C"], [b"GGCTCGG"])
    assert [(h["pattern"], h["start"]) for h in r.hits] == [(0, 2), (1, 2), (2, 3), (0, 4)]


def test_golden_fixture_texts(ref_tree):
    fx = ref_tree / "tests" / "fixtures" / "input"
    recs = [r.seq for r in rm.parse_fastx((fx / "simple.fasta").read_bytes())]
    check_batch([b"ACG", b"CGT"], recs)
    recs = [r.seq for r in rm.parse_fastx((fx / "fixed-width.faa").read_bytes())]
    r = check_batch([b"DKAT"], recs)
    assert r.hits["start"].tolist() == [79, 272]
    _, sam = rm.parse_bam((fx / "simple.bam").read_bytes())
    pats = rm.parse_pattern_list(None, ["CTC", "AC", "CT", "AA", "T", "A", "C", "G", "GA", "AG"], True, False, False, False)
    r = check_batch([p.encode() for p in pats], [s.seq for s in sam])
    assert r.n_hits == 96  # tests/fixtures/extract/log.json


def test_cfg1_example_minimal(ref_tree):
    recs = [r.seq for r in rm.parse_fastx((ref_tree / "example-minimal" / "sample.fasta").read_bytes())]
    r = check_batch([b"AAC", b"GTT"], recs)
    assert r.n_hits == 74 and sum(len(x) for x in recs) == 1795


def test_example_workflow_reads(ref_tree):
    ew = ref_tree / "example-workflow"
    pats = rm.parse_pattern_list(ew / "data" / "significant_kmers.txt", None, True, False, False, False)
    r1 = [r.seq for r in rm.parse_fastx((ew / "data" / "mutant_R1.fastq").read_bytes())]
    r2 = [r.seq for r in rm.parse_fastx((ew / "data" / "mutant_R2.fastq").read_bytes())]
    a = check_batch([p.encode() for p in pats], r1)
    b = check_batch([p.encode() for p in pats], r2)
    assert a.n_hits + b.n_hits == 36 and len(a.flagged_records()) == 9 and len(b.flagged_records()) == 15


@pytest.mark.parametrize("k", [1, 2, 3, 5, 8, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 27, 30, 31, 32, 33, 47, 63, 64, 65, 100])
def test_random_single_length(k):
    rng = np.random.default_rng(1000 + k)
    n_pat = 3 if k < 4 else 40
    pats = sorted({rand_seq(rng, k) for _ in range(n_pat)})
    recs = planted_records(rng, pats, 300, 0, 220, plant_p=0.4 if k > 3 else 0.0)
    check_batch(pats, recs)


@pytest.mark.parametrize("seed", range(6))
def test_random_mixed_lengths(seed):
    rng = np.random.default_rng(7 + seed)
    lo = [1, 4, 12, 17, 21, 31][seed]
    pats = sorted({rand_seq(rng, int(rng.integers(lo, lo + 45))) for _ in range(60)})
    # suffix/prefix-related patterns: two patterns ending at the same position
    pats = sorted(set(pats) | {p[3:] for p in pats[:10] if len(p) - 3 >= lo} | {p[:-2] for p in pats[10:20] if len(p) - 2 >= lo})
    recs = planted_records(rng, pats, 400, 0, 400, plant_p=0.5)
    check_batch(pats, recs)


def test_non_acgt_queries_and_text():
    rng = np.random.default_rng(5)
    base = [rand_seq(rng, 31) for _ in range(30)]
    pats = []
    for p in base:
        b = bytearray(p)
        b[int(rng.integers(31))] = ord("N")
        pats.append(bytes(b))
    pats += [b"MKVLAAGIVGLCAKENYVDQWERTPLSHFM", b"acgtacgtacgtacgtacgtacgtacgtacg", b"ACGTNNNNNNNNNNNNNNNNNNNNNNNACGT"]
    pats = sorted(set(pats + base[:5]))
    recs = planted_records(rng, pats, 300, 20, 300, alphabet=b"ACGTNacgt", plant_p=0.6)
    # the same windows with one byte changed to N / lower case must not hit
    broken = []
    for p in base[:10]:
        b = bytearray(b"GG" + p + b"GG")
        b[2 + 7] = ord("N")
        broken.append(bytes(b))
        broken.append((b"GG" + p + b"GG").lower())
    check_batch(pats, recs + broken)


def test_case_modes():
    rng = np.random.default_rng(11)
    pats = sorted({rand_seq(rng, 24) for _ in range(20)})
    recs = planted_records(rng, pats, 200, 10, 200, plant_p=0.6)
    recs = [r.lower() if i % 3 == 0 else (r.swapcase() if i % 3 == 1 else r) for i, r in enumerate(recs)]
    check_batch(pats, recs, case_insensitive=False)
    check_batch(pats, recs, case_insensitive=True)
    # -L style: lower-case queries against upper-case reads -> no hits unless -I
    low = sorted(p.lower() for p in pats)
    up = [r.upper() for r in recs]
    r = check_batch(low, up, case_insensitive=False)
    assert r.n_hits == 0
    r = check_batch(low, up, case_insensitive=True)
    assert r.n_hits > 0
    # patterns equal up to case share their spans under -I: ascending pattern index
    r = check_batch([b"ACGTACGTACGTAC", b"acgtacgtacgtac", b"AcGtAcGtAcGtAc"], [b"TTACGTACGTACGTACTT"], case_insensitive=True)
    assert r.hits["pattern"].tolist() == [0, 1, 2]


def test_record_boundaries_and_ragged_layout():
    rng = np.random.default_rng(3)
    p = rand_seq(rng, 31)
    # a hit may not straddle two records; hits at offset 0 and at len-k
    recs = [p[:15], p[15:], p, b"", b"", p + p, b"A", p[:-1], b"G" + p, p + b"G", b""]
    r = check_batch([p], recs)
    assert r.hits["record"].tolist() == [2, 5, 5, 8, 9]
    # odd record lengths shift every later record off the 16-byte grid
    recs = planted_records(rng, [p], 500, 0, 97, plant_p=0.7)
    check_batch([p], recs)


def test_long_record_every_alignment():
    rng = np.random.default_rng(17)
    pats = sorted({rand_seq(rng, int(k)) for k in (21, 31, 33, 40, 64, 65)})
    big = bytearray(rand_seq(rng, 1 << 20))
    pos = 1000
    for i in range(600):
        p = pats[i % len(pats)]
        big[pos:pos + len(p)] = p
        pos += 1600 + (i % 67)  # walks through every residue mod 16/32/512/2048
    check_batch(pats, [bytes(big[:300000]), bytes(big[300000:])])


def test_filter_flavours_of_the_shared_memory_filter(monkeypatch):
    """Small seed sets use 32-bit filter blocks (one LDS.32 per probe), larger ones 64-bit blocks; both are
    only filters: the hits are the oracle's either way, for unit seeds (k >= 31), window seeds and BAM4."""
    rng = np.random.default_rng(88)
    for k in (21, 27, 31, 40):
        pats = sorted({rand_seq(rng, k) for _ in range(40)})
        recs = planted_records(rng, pats, 400, 0, 300, plant_p=0.5)
        with capi.Engine(pats, max_batch_bytes=200000, max_batch_records=1000) as e:
            check_batch(pats, recs, engine=e)
        check_bam4(pats, recs)
    monkeypatch.setenv("MK_NO_FILTER32", "1")
    for k in (21, 31):
        pats = sorted({rand_seq(rng, k) for _ in range(40)})
        recs = planted_records(rng, pats, 400, 0, 300, plant_p=0.5)
        check_batch(pats, recs)
        check_bam4(pats, recs)


@pytest.mark.parametrize("k", [15, 18, 19, 22, 23, 30])
def test_window_scan_equals_ordered_scan(monkeypatch, k):
    """Strides 8 and 4 have two kernels (window seeds in the permuted packing / ordered packing): both
    must report the oracle's hits, also for long records that cross every tile boundary."""
    rng = np.random.default_rng(500 + k)
    pats = sorted({rand_seq(rng, int(x)) for x in rng.integers(k, k + 20, size=30)})
    big = bytearray(rand_seq(rng, 300000))
    pos = 100
    for i in range(400):
        p = pats[i % len(pats)]
        big[pos:pos + len(p)] = p
        pos += 700 + (i % 53)
    recs = [bytes(big[:100001]), bytes(big[100001:])] + planted_records(rng, pats, 200, 0, 200, plant_p=0.5)
    check_batch(pats, recs)
    monkeypatch.setenv("MK_NO_WIN_SCAN", "1")
    check_batch(pats, recs)


def test_bucket_sort_and_its_radix_fallback(monkeypatch):
    """The hit list is sorted by a bucket split + shared-memory sort; hits piled up in one region overflow
    a bucket and the batch falls back to the radix passes. Both give the oracle's order."""
    rng = np.random.default_rng(41)
    # (a) evenly spread hits, many of them: bucket path
    pats = sorted({rand_seq(rng, int(k)) for k in rng.integers(5, 9, size=40)})
    recs = [rand_seq(rng, int(rng.integers(50, 400))) for _ in range(3000)]
    check_batch(pats, recs)
    # (b) a poly-A island in a long record, 24 nested patterns: > 2048 hits in one bucket
    pile = [b"A" * k for k in range(1, 25)]
    text = rand_seq(rng, 600000, b"CGT") + b"A" * 3000 + rand_seq(rng, 600000, b"CGT")
    r = check_batch(pile, [text, rand_seq(rng, 5000, b"CGT")])
    assert r.n_hits > 60000
    # the same through the radix passes only
    monkeypatch.setenv("MK_NO_BUCKET_SORT", "1")
    check_batch(pats, recs)
    check_batch(pile, [text])


def test_hit_overflow_rescan():
    rng = np.random.default_rng(23)
    recs = [rand_seq(rng, 150) for _ in range(200)]
    r = check_batch([b"A", b"AC"], recs, hit_capacity=64)
    assert r.n_rescans >= 1 and r.n_hits > 64


def test_many_patterns_l2_filter(monkeypatch):
    # the first-level filter left L2-resident (what a seed set too large for shared memory gets)
    monkeypatch.setenv("MK_FILTER_MODE", "l2")
    rng = np.random.default_rng(29)
    genome = rand_seq(rng, 400000)
    starts = rng.integers(0, len(genome) - 70, size=20000)
    pats = sorted({genome[s:s + int(k)] for s, k in zip(starts, rng.integers(21, 64, size=len(starts)))})
    with capi.Engine(pats, max_batch_bytes=len(genome) + 64, max_batch_records=4) as e:
        check_batch(pats, [genome[:150000], genome[150000:]], engine=e)
        assert e.info().filter_in_smem[0] == 0
    for k in (8, 16, 31):
        pats = sorted({rand_seq(rng, k) for _ in range(50)})
        check_batch(pats, planted_records(rng, pats, 200, 0, 200, plant_p=0.5))


def pack_bam4(recs):
    """BAM sequence layout: 2 bases / byte (first in the high nibble), records byte aligned, explicit lengths."""
    nib = np.full(256, 15, dtype=np.uint8)
    for i, c in enumerate(NIB):
        nib[c] = i
    chunks, off, lens, pos = [], [0], [], 0
    for r in recs:
        codes = nib[np.frombuffer(r, dtype=np.uint8)] if r else np.zeros(0, np.uint8)
        if len(codes) % 2:
            codes = np.append(codes, 0)
        chunks.append((codes[0::2] << 4 | codes[1::2]).astype(np.uint8))
        lens.append(len(r))
        pos += len(codes)
        off.append(pos)
    packed = np.concatenate(chunks) if chunks else np.zeros(0, np.uint8)
    return packed, np.array(off, dtype=np.uint64), np.array(lens, dtype=np.uint32)


def check_bam4(pats, recs):
    ac = rm.AhoCorasick(pats, False)
    packed, off, lens = pack_bam4(recs)
    seq, aoff = pack_records(recs)
    rec, st, pat = ac.batch_hits(seq, aoff)
    with capi.Engine(pats, max_batch_bytes=int(packed.size) + 64, max_batch_records=len(recs)) as e:
        r = e.scan(packed, off, capi.MK_MODE_ALL_HITS, enc=capi.MK_ENC_BAM4, lens=lens, n_units=int(off[-1]))
        np.testing.assert_array_equal(r.hits["record"], rec)
        np.testing.assert_array_equal(r.hits["start"], st)
        np.testing.assert_array_equal(r.hits["pattern"], pat)
        p = e.scan(packed, off, capi.MK_MODE_PATTERN_SET, enc=capi.MK_ENC_BAM4, lens=lens, n_units=int(off[-1]))
        assert list(zip(p.hits["record"].tolist(), p.hits["pattern"].tolist())) == sorted(set(zip(rec.tolist(), pat.tolist())))
        f = e.scan(packed, off, capi.MK_MODE_FLAG, enc=capi.MK_ENC_BAM4, lens=lens, n_units=int(off[-1]))
        np.testing.assert_array_equal(f.flagged_records(), np.unique(rec))
        return e.info()


def test_large_pattern_set_tables_built_in_the_background():
    # >= 50 000 patterns: the seed tables of both encodings are built on host threads while CUDA starts up
    rng = np.random.default_rng(61)
    genome = rand_seq(rng, 300000)
    starts = rng.integers(0, len(genome) - 70, size=70000)
    pats = sorted({genome[s:s + int(k)] for s, k in zip(starts, rng.integers(24, 60, size=len(starts)))})
    assert len(pats) >= 50000
    recs = [genome[:100000], genome[100000:200001], genome[200001:]]
    with capi.Engine(pats, max_batch_bytes=len(genome) + 64, max_batch_records=4) as e:
        r = check_batch(pats, recs, engine=e)
        assert r.n_hits >= len(pats) - 100
    check_bam4(pats[:60000], recs)


@pytest.mark.parametrize("kmin", [16, 19, 21, 22, 23, 27, 31])
def test_bam4_encoding(kmin):
    # kmin 16: stride 4 (ordered packing); 19..30: stride 8 (word-aligned windows, masked below 23); 31: stride 16
    rng = np.random.default_rng(31 + kmin)
    pats = sorted({rand_seq(rng, int(k)) for k in rng.integers(kmin, kmin + 34, size=40)} | {b"ACGTNNACGTACGTAAGGCTNACACGTTGCAAT", b"acgtacgtacgtacgtacgtacgtacgtacgtac"})
    recs = planted_records(rng, [p for p in pats if p.isupper()], 300, 0, 180, alphabet=b"ACGTNRY", plant_p=0.6)
    check_bam4(pats, recs)


@pytest.mark.parametrize("lo", [2, 5, 12, 17, 21, 25])
@pytest.mark.parametrize("long_seeds", [True, False])
def test_dual_key_filter_mixed_lengths(monkeypatch, lo, long_seeds):
    """Large mixed-length query sets with a stride below 16 (BASELINE cfg5): patterns of >= d + 15 bases
    are keyed by a 16-base seed, the others by the q-base seed; both are probed with one L2 load."""
    monkeypatch.setenv("MK_FILTER_MODE", "l2")
    if not long_seeds:
        monkeypatch.setenv("MK_NO_LONG_SEEDS", "1")
    rng = np.random.default_rng(100 + lo)
    genome = rand_seq(rng, 120000)
    starts = rng.integers(0, len(genome) - 70, size=3000)
    pats = {genome[s:s + int(k)] for s, k in zip(starts, rng.integers(lo, lo + 44, size=len(starts)))}
    # queries that straddle the group threshold by one base, queries with N, related prefixes / suffixes
    some = sorted(pats)[:200]
    pats |= {p[1:] for p in some if len(p) - 1 >= lo} | {p[:-1] for p in some if len(p) - 1 >= lo}
    for p in some[:40]:
        if len(p) >= 8:
            b = bytearray(p)
            b[len(b) // 2] = ord("N")
            pats.add(bytes(b))
    pats = sorted(pats)
    text = bytearray(genome)
    for i, p in enumerate(q for q in pats if b"N" in q):
        text[1000 + 500 * i: 1000 + 500 * i + len(p)] = p
    text[50000:50300] = bytes(text[50000:50300]).lower()
    text = bytes(text)
    recs = [text[:33333], text[33333:90001], text[90001:]]
    with capi.Engine(pats, max_batch_bytes=len(text) + 64, max_batch_records=4) as e:
        r = check_batch(pats, recs, engine=e)
        info = e.info()
        assert info.filter_in_smem[0] == 0 and r.n_hits >= len(starts)
    with capi.Engine(pats, case_insensitive=True, max_batch_bytes=len(text) + 64, max_batch_records=4) as e:
        check_batch(pats, recs, case_insensitive=True, engine=e)
    if lo >= 12:
        info = check_bam4(pats, [r.upper() for r in recs] + [b"", b"ACG"])
        assert info.filter_in_smem[1] == 0


def test_double_buffered_slots_match_single_batch():
    rng = np.random.default_rng(37)
    pats = sorted({rand_seq(rng, 31) for _ in range(100)})
    recs = planted_records(rng, pats, 4000, 150, 150, plant_p=0.05)
    whole = oracle_hits(pats, recs)
    with capi.Engine(pats, n_slots=2, max_batch_bytes=1000 * 150, max_batch_records=1000) as e:
        got = []
        pending = []
        for b in range(4):
            part = recs[b * 1000:(b + 1) * 1000]
            seq, off = pack_records(part)
            slot = b % 2
            if len(pending) == 2:
                s0, b0 = pending.pop(0)
                r = e.wait(s0)
                got.append((b0, r.hits.copy()))
            a, o, _ = e.slot_arrays(slot)
            a[: seq.size] = seq
            o[: off.size] = off
            e.submit(slot, len(part), int(seq.size), capi.MK_ENC_ASCII, capi.MK_MODE_ALL_HITS)
            pending.append((slot, b))
        for s0, b0 in pending:
            got.append((b0, e.wait(s0).hits.copy()))
    rec = np.concatenate([h["record"] + 1000 * b for b, h in got])
    st = np.concatenate([h["start"] for b, h in got])
    pat = np.concatenate([h["pattern"] for b, h in got])
    np.testing.assert_array_equal(rec, whole[0])
    np.testing.assert_array_equal(st, whole[1])
    np.testing.assert_array_equal(pat, whole[2])


def test_errors_across_the_boundary():
    with pytest.raises(capi.MkError) as ei:
        capi.Engine([])
    assert ei.value.code == -3 and "No k-mers found" in ei.value.message
    with pytest.raises(capi.MkError) as ei:
        capi.Engine([b"ACGT", b""])
    assert ei.value.code == -2 and "Pattern is empty" in ei.value.message
    with capi.Engine([b"ACGT"], n_slots=1, max_batch_bytes=64, max_batch_records=4) as e:
        with pytest.raises(capi.MkError) as ei:
            e.scan(np.zeros(1000, np.uint8), np.array([0, 1000], dtype=np.uint64))
        assert ei.value.code == -6
        with pytest.raises(capi.MkError) as ei:
            e.wait(0)
        assert ei.value.code == -7


def test_device_resident_batches_in_flight():
    """mk_scan_device_submit: two batches that already sit in device memory, in flight on two slots."""
    import torch
    rng = np.random.default_rng(71)
    pats = sorted({rand_seq(rng, 31) for _ in range(60)})
    parts = [planted_records(rng, pats, 1500, 100, 160, plant_p=0.2) for _ in range(3)]
    with capi.Engine(pats, n_slots=2, max_batch_bytes=64, max_batch_records=4) as e:
        dev = []
        for recs in parts:
            seq, off = pack_records(recs)
            d_seq = torch.from_numpy(np.concatenate([seq, np.zeros(64, np.uint8)])).cuda()
            d_off = torch.from_numpy(off.astype(np.int64)).cuda()
            dev.append((d_seq, d_off, len(recs), int(seq.size)))
        got = []
        for i, (d_seq, d_off, n, nb) in enumerate(dev):
            if i >= 2:
                got.append(e.wait(i % 2))
            e.scan_device_submit(i % 2, d_seq.data_ptr(), d_off.data_ptr(), n, nb, capi.MK_MODE_ALL_HITS, fetch=True)
        got.append(e.wait(1))
        got.append(e.wait(0))
        with pytest.raises(capi.MkError):
            e.wait(0)  # nothing in flight any more
        for recs, r in zip(parts, got):
            rec, st, pat = oracle_hits(pats, recs)
            np.testing.assert_array_equal(r.hits["record"], rec)
            np.testing.assert_array_equal(r.hits["start"], st)
            np.testing.assert_array_equal(r.hits["pattern"], pat)
            np.testing.assert_array_equal(r.flagged_records(), np.unique(rec))


def test_fixed_length_batches_without_offsets():
    """mk_scan_host_uniform: records of one common length, offsets generated on the device (ASCII and BAM4)."""
    rng = np.random.default_rng(73)
    pats = sorted({rand_seq(rng, 31) for _ in range(50)})
    recs = planted_records(rng, pats, 3000, 150, 150, plant_p=0.2)
    seq, off = pack_records(recs)
    rec, st, pat = oracle_hits(pats, recs)
    with capi.Engine(pats, n_slots=2, max_batch_bytes=int(seq.size) + 64, max_batch_records=len(recs)) as e:
        e.scan_host_uniform_async(0, seq, len(recs), 150, capi.MK_ENC_ASCII, capi.MK_MODE_ALL_HITS)
        r = e.wait(0)
        np.testing.assert_array_equal(r.hits["record"], rec)
        np.testing.assert_array_equal(r.hits["start"], st)
        np.testing.assert_array_equal(r.hits["pattern"], pat)
        packed, poff, lens = pack_bam4(recs)
        e.scan_host_uniform_async(1, packed, len(recs), 150, capi.MK_ENC_BAM4, capi.MK_MODE_FLAG)
        f = e.wait(1)
        np.testing.assert_array_equal(f.flagged_records(), np.unique(rec))
        with pytest.raises(capi.MkError):
            e.scan_host_uniform_async(0, packed, 10, 151, capi.MK_ENC_BAM4, capi.MK_MODE_FLAG)


@pytest.mark.parametrize("kmin", [19, 20, 21, 22, 23, 26, 30])
@pytest.mark.parametrize("foreign", [b"n", b"N", b"\x00", b"acgt"])
def test_dual8_alphabet_gate_edges(monkeypatch, kmin, foreign):
    """mk_scan_dual8 (stride 8, L2-resident dual-key filter) with its alphabet gate: bytes that occur in no pattern sit
    directly in front of, directly behind and at every distance from occurrences, for every seed length q = 12..16
    and every alignment of an occurrence to the 8-base grid. The gate may only drop windows that cannot match."""
    monkeypatch.setenv("MK_FILTER_MODE", "l2")
    rng = np.random.default_rng(kmin * 7 + len(foreign))
    pats = sorted({rand_seq(rng, int(k)) for k in rng.integers(kmin, kmin + 30, size=300)} | {rand_seq(rng, kmin) for _ in range(20)})
    pieces = []
    for i in range(1500):
        p = pats[int(rng.integers(len(pats)))]
        gap = int(rng.integers(0, 40))
        x = bytes([foreign[int(rng.integers(len(foreign)))]])
        kind = i % 4
        if kind == 0:    # foreign byte right behind and right in front of occurrences
            pieces += [p, x, p, x * int(rng.integers(1, 20))]
        elif kind == 1:  # clean text of any length between occurrences: every alignment to the grid
            pieces += [p, rand_seq(rng, gap)]
        elif kind == 2:  # an occurrence spoiled by one foreign byte anywhere inside it (no hit)
            b = bytearray(p)
            b[int(rng.integers(len(b)))] = x[0]
            pieces += [bytes(b), rand_seq(rng, gap), p]
        else:            # long foreign spans (whole tiles without a live window) with occurrences at their edges
            pieces += [x * int(rng.integers(100, 5000)), p, x * int(rng.integers(1, 3)), p]
    text = b"".join(pieces)
    cuts = sorted(set(rng.integers(0, len(text), size=6).tolist()) | {0, len(text)})
    recs = [text[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
    with capi.Engine(pats, max_batch_bytes=len(text) + 64, max_batch_records=len(recs)) as e:
        r = check_batch(pats, recs, engine=e)
        assert "mk_scan_dual8" in e.scan_kernel(capi.MK_ENC_ASCII) and r.n_hits > 2000
        gate = bool(e.info().features & capi.MK_FEATURE_GATE)
        assert gate == (foreign != b"acgt" or True)  # upper-case ACGT patterns: bit 5 (and others) are constant, so a gate always exists
    if foreign == b"acgt":  # -I: the folded alphabet holds both cases, the lower-case text matches and must not be gated away
        with capi.Engine(pats, case_insensitive=True, max_batch_bytes=len(text) + 64, max_batch_records=len(recs)) as e:
            r2 = check_batch(pats, recs, case_insensitive=True, engine=e)
            assert r2.n_hits >= r.n_hits


def test_candidate_index_range_is_refused_not_wrapped():
    """candidate positions are 32-bit grid indices; a batch whose last index would not fit is refused
    with MK_ERR_CAPACITY before anything is allocated or launched (it used to be scanned at wrapped positions). The
    device buffer here is tiny: the sizes are only claimed."""
    import torch
    rng = np.random.default_rng(1)
    d = torch.zeros(4096, dtype=torch.uint8, device="cuda")
    off = torch.zeros(2, dtype=torch.int64, device="cuda")
    cases = [
        # (patterns, encoding, bases that must be refused, bases just below the limit of that geometry)
        ([rand_seq(rng, 31) for _ in range(50)], capi.MK_ENC_BAM4, (1 << 36) + 64, None),       # stride 16 BAM4: 2 candidates per vector
        ([rand_seq(rng, 27) for _ in range(50)], capi.MK_ENC_ASCII, (1 << 35) + 64, None),      # window scan, stride 8: 2 per vector
        ([rand_seq(rng, 16) for _ in range(50)], capi.MK_ENC_ASCII, (1 << 34) + 64, None),      # window scan, stride 4: 4 per vector
        ([rand_seq(rng, 12) for _ in range(50)], capi.MK_ENC_ASCII, (1 << 32) + 64, None),      # ordered scan: base positions
    ]
    for pats, enc, too_many, _ in cases:
        with capi.Engine(sorted(set(pats)), n_slots=0) as e:
            with pytest.raises(capi.MkError) as ei:
                e.scan_device(d.data_ptr(), off.data_ptr(), 1, too_many, capi.MK_MODE_FLAG, enc)
            assert ei.value.code == -6 and ("candidate index" in ei.value.message or "2^32 bases" in ei.value.message), ei.value.message
            # the engine is still usable
            small = e.scan_device(d.data_ptr(), off.data_ptr(), 1, 0, capi.MK_MODE_FLAG, enc, fetch=True)
            assert small.n_hits == 0


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 7, 9, 10, 11, 12, 13, 14])
def test_short_patterns_direct_bitmap(k):
    """mk_scan_short (strides 1 and 2, direct prefix bitmap): every pattern length below 15, dense hits (short patterns
    occur everywhere), bursts that overflow the per-warp queue (poly-A), N and lower case next to occurrences, records
    of every length, and a pattern set mixing the shortest length with much longer ones."""
    rng = np.random.default_rng(200 + k)
    pats = sorted({rand_seq(rng, k) for _ in range(6)} | {rand_seq(rng, int(x)) for x in rng.integers(k, k + 40, size=30)} | {b"A" * k})
    recs = planted_records(rng, pats, 300, 0, 400, plant_p=0.6)
    recs += [b"A" * 3000, b"", b"ACGT"[:k], rand_seq(rng, 20000), (b"acgtn" * 50) + pats[0] + b"N" + pats[-1] + b"n" + pats[1]]
    with capi.Engine(pats, max_batch_bytes=sum(map(len, recs)) + 64, max_batch_records=len(recs), hit_capacity=1 << 12) as e:
        check_batch(pats, recs, engine=e)
        assert "mk_scan_short" in e.scan_kernel(capi.MK_ENC_ASCII), e.scan_kernel(capi.MK_ENC_ASCII)
    if k >= 3:  # -I through the same path
        check_batch(pats[:12], recs[:100] + [recs[-1]], case_insensitive=True)


def test_engines_on_shared_tables_and_free_running_slots(monkeypatch):
    """mk_tables_create / mk_engine_create_shared through capi.Tables: two engines on one table object give the oracle's
    hits, also after the Python handle of the tables is closed (the engines hold their own reference); and with
    MK_FREE=1 (the slots' streams not ordered among each other) batches in flight still come back complete."""
    import torch
    rng = np.random.default_rng(91)
    pats = sorted({rand_seq(rng, int(rng.integers(31, 50))) for _ in range(300)})
    recs = planted_records(rng, pats, 3000, 100, 200, plant_p=0.2)
    want = oracle_hits(pats, recs)
    tables = capi.Tables(pats)
    e1 = capi.Engine(tables, n_slots=2, max_batch_bytes=1 << 20, max_batch_records=1 << 12)
    e2 = capi.Engine(tables, n_slots=2, max_batch_bytes=1 << 20, max_batch_records=1 << 12)
    tables.close()
    seq, off = pack_records(recs)
    for e in (e1, e2):
        r = e.scan(seq, off)
        for name, w in zip(("record", "start", "pattern"), want):
            np.testing.assert_array_equal(r.hits[name], w)
    e2.close()
    monkeypatch.setenv("MK_FREE", "1")
    d_seq = torch.from_numpy(np.concatenate([seq, np.zeros(64, np.uint8)])).cuda()
    d_off = torch.from_numpy(off.astype(np.int64)).cuda()
    for i in range(6):
        if i >= 2:
            r = e1.wait(i % 2)
            for name, w in zip(("record", "start", "pattern"), want):
                np.testing.assert_array_equal(r.hits[name], w)
        e1.scan_device_submit(i % 2, d_seq.data_ptr(), d_off.data_ptr(), len(recs), int(seq.size), capi.MK_MODE_ALL_HITS, fetch=True)
    for s in (0, 1):
        r = e1.wait(s)
        np.testing.assert_array_equal(r.hits["start"], want[1])
    e1.close()


@pytest.mark.parametrize("knobs", [{"MK_NO_WIN_SCAN": "1"}, {"MK_NO_WIN_SCAN": "1", "MK_NO_LONG_SEEDS": "1"}, {"MK_NO_WIN_SCAN": "1", "MK_NO_GATE": "1"}])
def test_l2_filter_stride_8_without_the_window_layout_of_the_first_index_pass(monkeypatch, knobs):
    """Found by scripts/gpu_fuzz.py: with MK_NO_WIN_SCAN the first index pass packs ordered seed codes, and the stride-8
    scan with the L2-resident filter (mk_scan_dual8, which takes window codes) found nothing. The builder now indexes
    again for it. Amino-acid alphabet, k = 19..35, every switch combination that reached the bug."""
    rng = np.random.default_rng(77)
    alpha = b"ACDEFGHIKLMNPQRSTVWY"
    monkeypatch.setenv("MK_FILTER_MODE", "l2")
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    for kmin, kmax, n_pat in ((25, 35, 7), (21, 31, 1), (19, 20, 300)):
        pats = sorted({rand_seq(rng, int(rng.integers(kmin, kmax + 1)), alpha) for _ in range(n_pat)})
        recs = planted_records(rng, pats, 400, 0, 300, alpha, plant_p=0.5)
        r = check_batch(pats, recs)
        assert r.n_hits > 50
