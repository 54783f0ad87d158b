"""The synthetic generator produces identical data on host and device, and the device scan of it is
identical to the oracle's (BASELINE config 2 shape at reduced size)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_synth_host_equals_device_and_scan_parity():
    import torch
    from merkurio_b200 import capi, patterns as pt
    from merkurio_b200.synth import Synth
    from oracle import refmodel as rm

    n, L = 200_000, 150
    syn = Synth(0x5EED0002, n, L, 31, 1000)
    h_seq, h_off = syn.host_reads(0, n)
    d_seq = torch.zeros(n * L + 64, dtype=torch.uint8, device="cuda")
    d_off = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    d_q = torch.from_numpy(syn.queries).cuda()
    syn.device_reads(d_q.data_ptr(), 0, n, d_seq.data_ptr(), d_off.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_seq[: n * L].cpu().numpy(), h_seq)
    assert np.array_equal(d_off.cpu().numpy().astype(np.uint64), h_off)
    # a sub-range generated on its own equals the slice of the whole
    part, _ = syn.host_reads(1234, 1300)
    assert np.array_equal(part, h_seq[1234 * L: 1300 * L])

    pats = pt.parse_pattern_list(syn.query_list(), reverse_complement_=True)
    assert 1990 <= len(pats) <= 2000
    ac = rm.AhoCorasick(pats)
    rec, st, pat = ac.batch_hits(h_seq, h_off)
    frac = len(np.unique(rec)) / n
    assert 0.005 < frac < 0.03  # ~1 % of the reads hit
    with capi.Engine(pats, n_slots=0) as e:
        r = e.scan_device(d_seq.data_ptr(), d_off.data_ptr(), n, n * L, capi.MK_MODE_ALL_HITS, fetch=True)
        assert np.array_equal(r.hits["record"], rec) and np.array_equal(r.hits["start"], st) and np.array_equal(r.hits["pattern"], pat)
        f = e.scan_device(d_seq.data_ptr(), d_off.data_ptr(), n, n * L, capi.MK_MODE_FLAG, fetch=True)
        assert np.array_equal(f.flagged_records(), np.unique(rec))


def test_synth_bam4_matches_ascii():
    from merkurio_b200.synth import Synth
    from oracle import refmodel as rm
    syn = Synth(7, 5000, 150, 31, 100)
    a, _ = syn.host_reads(0, 5000, 0)
    b, _ = syn.host_reads(0, 5000, 1)
    dec = np.frombuffer(rm.NIBBLE_CHARS, dtype=np.uint8)
    unpacked = np.empty(b.size * 2, dtype=np.uint8)
    unpacked[0::2] = dec[b >> 4]
    unpacked[1::2] = dec[b & 15]
    assert np.array_equal(unpacked, a)
