"""Checkers shared by bench.py's `configs` entries, scripts/bench_configs.py and tests/test_gpu_fullsize.py:
the device result of a workload (merkurio_b200/synth/workloads.py) against the oracle on a bounded sample,
and against direct byte comparison at full size.

TEST INFRASTRUCTURE (like everything under oracle/): the product never imports this module."""
from __future__ import annotations

import numpy as np

from . import refmodel as rm


def reverify_hits(d_seq, d_off, hits, pats, bam4: bool) -> bool:
    """Every reported hit: text[start, start + len) == pattern and the hit stays inside its record (gathers on the GPU)."""
    import torch
    if len(hits) == 0:
        return True
    lens = np.array([len(x) for x in pats], dtype=np.int64)
    width = int(lens.max())
    pm = np.zeros((len(pats), width), dtype=np.uint8)
    for i, x in enumerate(pats):
        pm[i, :len(x)] = np.frombuffer(x, dtype=np.uint8)
    pm, pl = torch.from_numpy(pm).cuda(), torch.from_numpy(lens).cuda()
    dec = torch.from_numpy(np.frombuffer(rm.NIBBLE_CHARS, dtype=np.uint8).copy()).cuda()
    col = torch.arange(width, device="cuda")[None, :]
    ok = True
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
        ok = ok and bool(((txt == pm[pid]) | ~valid).all().item())
        ok = ok and bool((st + pl[pid] <= d_off[rec + 1] - d_off[rec]).all().item())
    return ok


def reads_sample_equal(full, wl, n_check: int) -> int:
    """Read workloads (cfg3 / cfg4): the records < n_check of the full-size device result `full` against the oracle's
    Aho-Corasick scan of the same reads generated on the host. Returns the number of oracle hits compared."""
    from merkurio_b200 import capi
    n_check = min(n_check, wl.n_records)
    h_seq, h_off = wl.syn.host_reads(0, n_check, 0)
    rec, st, pat = rm.AhoCorasick(wl.pats).batch_hits(h_seq, h_off)
    hits = full.hits
    sel = hits["record"] < n_check
    if wl.mode == capi.MK_MODE_ALL_HITS:
        assert np.array_equal(hits["record"][sel], rec) and np.array_equal(hits["start"][sel], st) and np.array_equal(hits["pattern"][sel], pat), \
            "hit list differs from the oracle"
    else:
        pairs = sorted(set(zip(rec.tolist(), pat.tolist())))
        assert list(zip(hits["record"][sel].tolist(), hits["pattern"][sel].tolist())) == pairs, "pattern sets differ from the oracle"
    bits = np.unpackbits(full.flags.view(np.uint8), bitorder="little")[:n_check]
    assert np.array_equal(np.nonzero(bits)[0], np.unique(rec)), "flags differ from the oracle"
    return int(len(rec))


def genome_slice_equal(full, wl, slice_len: int = 4_000_000) -> int:
    """cfg5: the hits of the FULL run (full pattern set, the engine and filter path under test) that lie inside the
    first slice_len bases of record 0, against the oracle's Aho-Corasick automaton of the same full pattern set run
    over that slice (a 30-40 M state DFA: ~25 s and ~5 GB to build). Returns the number of hits compared."""
    sl = min(slice_len, int(wl.extra["lens"][0]))
    text = wl.d_seq[:sl].cpu().numpy()
    rec, st, pat = rm.AhoCorasick(wl.pats).batch_hits(text, np.array([0, sl], dtype=np.uint64))
    hits = full.hits
    sel = (hits["record"] == 0) & (hits["start"].astype(np.int64) + hits["len"] <= sl)
    assert np.array_equal(hits["start"][sel], st) and np.array_equal(hits["pattern"][sel], pat), "slice of the full run differs from the oracle"
    return int(len(st))


def genome_expected_found(full, wl, step: int = 1) -> int:
    """cfg5: every query that still equals the text at its sampling position must be reported there. Returns the
    number of expected (record, start, pattern) triples that are missing."""
    x = wl.extra
    pid_of = {p_: i for i, p_ in enumerate(wl.pats)}
    hits = full.hits
    key = (hits["record"].astype(np.uint64) << np.uint64(52)) | (hits["start"].astype(np.uint64) << np.uint64(20)) | hits["pattern"].astype(np.uint64)
    idx = np.nonzero(x["expected"])[0][::step]
    q = x["queries"]
    want = np.array([(int(x["chrom"][i]) << 52) | (int(x["qs"][i] - x["off"][x["chrom"][i]]) << 20) | pid_of[q[i]] for i in idx], dtype=np.uint64)
    return int((~np.isin(want, key)).sum())
