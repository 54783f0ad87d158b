"""The BASELINE configs other than cfg2 as device-resident synthetic batches,
shared by bench.py (`configs` array), scripts/bench_configs.py and tests/test_gpu_fullsize.py.

  cfg3  paired-end 2 x 150 bp, 50 M pairs = 100 M mates, 10 000 canonical 31-mers, ALL_HITS (JSON log)
  cfg4  BAM 4-bit, 50 M x 150 bp, 10 000 31-mers, PATTERN_SET (km tag, keep-only-matching)
  cfg5  3 Gbp in 24 records with hg38's proportions, 2 % in N runs, 30 % lower-case soft-masked spans;
        1 M queries of 21..63 bases sampled from it, 1 % with one base replaced by N; ALL_HITS.
        `upper_queries` (the default, SURVEY 8d: "lower-case spans must NOT match without -I") upper-cases
        the sampled queries, as a k-mer list is; False keeps them verbatim (round 1's variant: a query
        sampled from a soft-masked span is lower case and matches there).

Generators only: nothing here touches the oracle. Needs a CUDA device (the data is made on the GPU)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from .. import capi
from .. import patterns as pt
from . import Synth

HG38_MBP = [248.9, 242.2, 198.3, 190.2, 181.5, 170.8, 159.3, 145.1, 138.4, 133.8, 135.1, 133.3, 114.4, 107.0, 102.0, 90.3,
            83.3, 80.4, 58.6, 64.4, 46.7, 50.8, 156.0, 57.2]


@dataclass
class Workload:
    name: str
    d_seq: object                 # torch uint8 tensor (device), 64 bytes of zero slack after the data
    d_off: object                 # torch int64 tensor (device), n_records + 1 unit offsets
    n_records: int
    n_units: int                  # bases
    enc: int
    mode: int
    pats: List[bytes]
    algorithmic_bytes: int        # bytes the scan must read: 1 B/base (ASCII) or 0.5 B/base (BAM4)
    hit_capacity: int
    syn: Optional[Synth] = None   # read workloads: the generator (host_reads for oracle samples)
    read_len: int = 0
    extra: dict = field(default_factory=dict)

    def scan(self, eng, fetch=False):
        return eng.scan_device(self.d_seq.data_ptr(), self.d_off.data_ptr(), self.n_records, self.n_units, self.mode, self.enc, fetch=fetch)


def reads_workload(cfg: str, scale: float = 1.0) -> Workload:
    import torch
    L = 150
    n = int((100_000_000 if cfg == "cfg3" else 50_000_000) * scale) // 64 * 64
    syn = Synth(0x5EED0003 if cfg == "cfg3" else 0x5EED0004, n, L, 31, 10000)
    if cfg == "cfg3":
        pats = pt.parse_pattern_list(syn.query_list(), canonical_=True)
        enc, mode, nbytes = capi.MK_ENC_ASCII, capi.MK_MODE_ALL_HITS, n * L
    elif cfg == "cfg4":
        pats = pt.parse_pattern_list(syn.query_list())
        enc, mode, nbytes = capi.MK_ENC_BAM4, capi.MK_MODE_PATTERN_SET, n * L // 2
    else:
        raise ValueError(cfg)
    d_seq = torch.empty(nbytes + 64, dtype=torch.uint8, device="cuda")
    d_off = torch.empty(n + 1, dtype=torch.int64, device="cuda")
    d_q = torch.from_numpy(syn.queries).cuda()
    syn.device_reads(d_q.data_ptr(), 0, n, d_seq.data_ptr(), d_off.data_ptr(), 1 if enc else 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    d_seq[nbytes:] = 0
    return Workload(cfg, d_seq, d_off, n, n * L, enc, mode, pats, nbytes, max(n // 20, 1 << 20), syn, L)


def genome_workload(scale: float = 1.0, upper_queries: bool = True, n_queries: Optional[int] = None) -> Workload:
    import torch
    total = int(3_000_000_000 * scale)
    lens = np.array([int(x / sum(HG38_MBP) * total) for x in HG38_MBP], dtype=np.int64)
    total = int(lens.sum())
    off = np.zeros(25, dtype=np.int64)
    off[1:] = np.cumsum(lens)
    g = torch.Generator(device="cuda").manual_seed(0x5EED0005)
    d_seq = torch.empty(total + 64, dtype=torch.uint8, device="cuda")
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device="cuda")
    step = 1 << 28
    for s in range(0, total, step):
        e = min(total, s + step)
        d_seq[s:e] = lut[torch.randint(0, 4, (e - s,), generator=g, device="cuda", dtype=torch.uint8).long()]
    rng = np.random.default_rng(5)
    pos = n_low = n_n = 0
    while pos < total:  # spans of 2-200 kbp: 30 % lower case (soft-masked), 2 % N
        e = min(total, pos + int(rng.integers(2000, 200000)))
        kind = rng.random()
        if kind < 0.30:
            d_seq[pos:e] |= 0x20
            n_low += e - pos
        elif kind < 0.32:
            d_seq[pos:e] = 78
            n_n += e - pos
        pos = e
    d_seq[total:] = 0
    nq = n_queries if n_queries is not None else int(1_000_000 * min(1.0, max(scale, 0.02)))
    ql = rng.integers(21, 64, size=nq)
    chrom = rng.choice(24, size=nq, p=lens / lens.sum())
    qs = off[chrom] + (rng.random(nq) * (lens[chrom] - ql)).astype(np.int64)
    idx = torch.from_numpy(qs).cuda()[:, None] + torch.arange(63, device="cuda")[None, :]
    qverb = d_seq[idx.clamp_(max=total - 1)].cpu().numpy()
    del idx
    # samples from inside an N run are dropped; one or two N at a run's edge stay (they occur verbatim)
    keep = ((qverb == 78) & (np.arange(63)[None, :] < ql[:, None])).sum(axis=1) <= 2
    qverb, ql, chrom, qs = qverb[keep], ql[keep], chrom[keep], qs[keep]
    nq = int(keep.sum())
    qmat = np.where((qverb >= 97) & (qverb <= 122), qverb - 32, qverb).astype(np.uint8) if upper_queries else qverb
    queries = [qmat[i, :ql[i]].tobytes() for i in range(nq)]
    for i in rng.choice(nq, size=nq // 100, replace=False):  # 1 %: one base becomes N
        b = bytearray(queries[i])
        b[int(rng.integers(len(b)))] = 78
        queries[i] = bytes(b)
    # a query is expected at its sampling position iff it still equals the text there
    expected = np.array([qverb[i, :ql[i]].tobytes() == queries[i] for i in range(nq)], dtype=bool)
    pats = pt.parse_pattern_list(queries)
    name = "cfg5" if upper_queries else "cfg5_verbatim_case"
    return Workload(name, d_seq, torch.from_numpy(off).cuda(), 24, total, capi.MK_ENC_ASCII, capi.MK_MODE_ALL_HITS, pats, total,
                    4 * nq, None, 0,
                    {"queries": queries, "chrom": chrom, "qs": qs, "off": off, "lens": lens, "expected": expected,
                     "lower_case_frac": n_low / total, "n_frac": n_n / total, "upper_queries": upper_queries})
