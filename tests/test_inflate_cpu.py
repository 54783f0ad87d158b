"""The host's own DEFLATE decoder (merkurio_b200/host/inflate.cpp: gzip members and BGZF blocks of the inputs the
reference reads through needletail / flate2) against zlib: every block type and code shape zlib's deflate can be made
to emit, streams handed out in pieces of every size, truncation at every byte, flipped bits — through the
`merkurio records <file> cat` diagnostic, which prints the decompressed byte stream exactly as the readers get it."""
import gzip
import io
import os
import random
import struct
import subprocess
import zlib
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def exe():
    from merkurio_b200.build import build_host
    return str(build_host())


def cat(exe, path, chunk=None, env=None):
    cmd = [exe, "records", str(path), "cat"] + ([str(chunk)] if chunk else [])
    r = subprocess.run(cmd, capture_output=True, env={**os.environ, **(env or {})})
    return r.returncode, r.stdout, r.stderr


def fastq(n, seed):
    rng = random.Random(seed)
    out = []
    for i in range(n):
        s = "".join(rng.choice("ACGT") for _ in range(rng.randint(50, 160)))
        q = "".join(rng.choice("FFFFFFFF:,#") for _ in range(len(s)))
        out.append(f"@read{i} x/1\n{s}\n+\n{q}\n")
    return "".join(out).encode()


def gz(raw, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, mem=9):
    c = zlib.compressobj(level, zlib.DEFLATED, 31, mem, strategy)
    return c.compress(raw) + c.flush()


def stream_cases():
    rng = random.Random(7)
    fq = fastq(6000, 1)
    rnd = os.urandom(700_000)
    cases = {
        "fastq_level6": (fq, gz(fq)),
        "fastq_level1": (fq, gz(fq, 1)),
        "fastq_level9": (fq, gz(fq, 9)),
        "stored_blocks": (fq, gz(fq, 0)),
        "fixed_codes": (fq[:200_000], gz(fq[:200_000], 6, zlib.Z_FIXED)),
        "huffman_only": (fq[:300_000], gz(fq[:300_000], 6, zlib.Z_HUFFMAN_ONLY)),
        "run_length": (fq[:300_000], gz(fq[:300_000], 6, zlib.Z_RLE)),
        "small_blocks": (fq, gz(fq, 6, mem=1)),  # memLevel 1: a new dynamic block every ~500 symbols
        "incompressible": (rnd, gz(rnd)),
        "zeros": (b"\0" * 3_000_000, gz(b"\0" * 3_000_000)),  # distance 1, length 258
        "period_25": ((b"ACGTTGCA" * 3 + b"N") * 80_000, gz((b"ACGTTGCA" * 3 + b"N") * 80_000, 9)),
        "short_periods": (b"".join(bytes(rng.randrange(256) for _ in range(p)) * rng.randint(2, 40) for p in [1, 2, 3, 4, 5, 6, 7] * 3000),) * 2,
        "empty": (b"", gz(b"")),
        "one_byte": (b"A", gz(b"A")),
        "members": (fq + rnd + b"tail", gz(fq) + gz(rnd, 1) + gz(b"") + gz(b"tail")),
    }
    raw, _ = cases["short_periods"]
    cases["short_periods"] = (raw, gz(raw, 9))
    # Z_SYNC_FLUSH / Z_FULL_FLUSH in the middle (empty stored blocks between the others), also within the last kilobyte of
    # the stream, where the decoder works on a padded copy of the input's tail
    def flushed(data, cuts, level=6):
        c = zlib.compressobj(level, zlib.DEFLATED, 31)
        out, pos = b"", 0
        for cut, how in cuts:
            out += c.compress(data[pos:cut]) + c.flush(how)
            pos = cut
        return out + c.compress(data[pos:]) + c.flush()
    n = len(fq)
    cases["flushes"] = (fq, flushed(fq, [(n // 5, zlib.Z_SYNC_FLUSH), (n // 2, zlib.Z_FULL_FLUSH), (n // 2 + 1, zlib.Z_SYNC_FLUSH), (n - 300, zlib.Z_SYNC_FLUSH), (n - 7, zlib.Z_FULL_FLUSH), (n, zlib.Z_SYNC_FLUSH)]))
    # dynamic, fixed and stored blocks alternating in one member: raw deflate pieces of independent compressors, each ended
    # with Z_FULL_FLUSH (byte aligned, not final), behind one gzip header — the decoder's tables change kind from block
    # to block, also in the tasks of the parallel reader that start in the middle
    def mixed(parts):
        out, raw_all = b"", b""
        for i, (data, level, strategy) in enumerate(parts):
            c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
            out += c.compress(data) + (c.flush() if i + 1 == len(parts) else c.flush(zlib.Z_FULL_FLUSH))
            raw_all += data
        return raw_all, b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\x03" + out + struct.pack("<II", zlib.crc32(raw_all), len(raw_all) & 0xFFFFFFFF)
    kinds = [(6, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED), (0, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_FIXED)]
    cases["mixed_block_types"] = mixed([(fastq(rng.randint(100, 700), 100 + i),) + kinds[i % len(kinds)] for i in range(40)])
    rep = b"A" * 1_500_000
    cases["flushes_in_a_short_stream"] = (rep, flushed(rep, [(700_000, zlib.Z_SYNC_FLUSH), (1_400_000, zlib.Z_FULL_FLUSH), (1_499_999, zlib.Z_SYNC_FLUSH)]))
    # a header with every optional field
    b = io.BytesIO()
    with gzip.GzipFile(filename="some_name.fq", mode="wb", fileobj=b, mtime=5) as f:
        f.write(fq[:5000])
    cases["file_name"] = (fq[:5000], b.getvalue())
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = co.compress(fq[:7000]) + co.flush()
    hdr = bytes([0x1F, 0x8B, 8, 4 | 8 | 16 | 2, 0, 0, 0, 0, 0, 3]) + bytes([5, 0]) + b"EXTRA" + b"name\0" + b"comment\0"
    hdr += struct.pack("<H", zlib.crc32(hdr) & 0xFFFF)
    cases["all_header_fields"] = (fq[:7000], hdr + body + struct.pack("<II", zlib.crc32(fq[:7000]), 7000))
    return cases


CASES = stream_cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_streams_decode_like_zlib(exe, tmp_path, name):
    raw, comp = CASES[name]
    assert gzip.decompress(comp) == raw
    p = tmp_path / "x.gz"
    p.write_bytes(comp)
    for chunk in (None, 1, 700, 65536):  # size of the pieces the caller asks for
        if chunk == 1 and len(raw) > 400_000:
            continue
        rc, out, err = cat(exe, p, chunk)
        assert rc == 0, (name, chunk, err)
        assert out == raw, (name, chunk)
    rc, out, _ = cat(exe, p, env={"MERKURIO_ZLIB_INFLATE": "1"})  # the zlib path stays equivalent
    assert rc == 0 and out == raw


@pytest.mark.parametrize("variant", ["dynamic", "fixed", "stored", "two_members"])
def test_truncation_at_every_byte(exe, tmp_path, variant):
    """A stream cut anywhere is an error, and what comes out before it is a prefix of the data (no symbol decoded from
    behind the end of the input is ever written) — except a cut exactly between two members, which is a complete file."""
    raw = fastq({"dynamic": 40, "fixed": 12, "stored": 8, "two_members": 8}[variant], 3)
    comp = {"dynamic": gz(raw), "fixed": gz(raw, 6, zlib.Z_FIXED), "stored": gz(raw, 0), "two_members": gz(raw) + gz(raw)}[variant]
    want = raw + raw if variant == "two_members" else raw
    p = tmp_path / "cut.gz"
    step = 1 if len(comp) < 1500 else 3
    for cut in range(2, len(comp), step):
        p.write_bytes(comp[:cut])
        rc, out, err = cat(exe, p)
        if rc == 0:
            assert gzip.decompress(comp[:cut]) == out, cut  # (zlib agrees that this prefix is a whole file)
        else:
            assert b"Error while decompressing the input" in err, (cut, err)
            assert want.startswith(out), cut


def test_flipped_bits_never_pass_silently(exe, tmp_path):
    """One flipped bit: either zlib still accepts the file (the bit sits in an ignored header field) and we return the
    same bytes, or the run ends with an error (invalid code, distance too far back, CRC-32 or length mismatch)."""
    rng = random.Random(5)
    raw = fastq(300, 9)
    comp = gz(raw)
    p = tmp_path / "bad.gz"
    for _ in range(150):
        c = bytearray(comp)
        c[rng.randrange(10, len(comp))] ^= 1 << rng.randrange(8)
        p.write_bytes(bytes(c))
        rc, out, _ = cat(exe, p)
        try:
            ref = gzip.decompress(bytes(c))
        except Exception:
            ref = None
        if ref is None:
            assert rc != 0
        else:
            assert rc == 0 and out == ref


def test_crc_and_length_of_the_trailer_are_checked(exe, tmp_path):
    raw = fastq(50, 4)
    comp = bytearray(gz(raw))
    p = tmp_path / "t.gz"
    for pos in (-8, -5, -4, -1):  # CRC-32 (4 bytes), ISIZE (4 bytes)
        c = bytearray(comp)
        c[pos] ^= 0x01
        p.write_bytes(bytes(c))
        rc, out, err = cat(exe, p)
        assert rc != 0 and b"Error while decompressing the input (gzip)" in err
        assert out == raw  # the data in front of the damage is handed out first
    raw2, with_hcrc = CASES["all_header_fields"]
    c = bytearray(with_hcrc)
    c[12] ^= 0x20  # inside the extra field: the header CRC-16 no longer matches
    p.write_bytes(bytes(c))
    rc, out, err = cat(exe, p)
    assert rc != 0 and out == b""
    p.write_bytes(bytes(comp) + b"\0\0\0\0")  # bytes that are no member behind the last one
    rc, out, err = cat(exe, p)
    assert rc != 0 and out == raw


def test_large_stream_in_pieces(exe, tmp_path):
    """Larger than the decoder's window and the input buffer: every refill and slide boundary is crossed."""
    raw = fastq(40_000, 11)
    for level in (1, 6):
        p = tmp_path / f"big{level}.gz"
        p.write_bytes(gz(raw, level))
        rc, out, err = cat(exe, p)
        assert rc == 0, err
        assert out == raw


# ------------------------------------------------------------------------------------------------
# One gzip member on several threads (merkurio_b200/host/pgzip.cpp): pieces of the compressed file are decoded without
# their 32 KiB of history and stitched where the real decode arrives at their first bit. Small pieces here, so that a
# megabyte of input is dozens of tasks; the result must be what zlib returns, whatever the tasks find.
PAR = {"MERKURIO_GZIP_THREADS": "4", "MERKURIO_GZIP_PIECE_KB": "16", "MERKURIO_TIMING": "1"}


def _stats(err: bytes):
    """(pieces, used, not used, sequential blocks) from the MERKURIO_TIMING line of the parallel reader."""
    import re
    m = re.search(rb"gzip on (\d+) threads: (\d+) pieces of the file, (\d+) continued the decode where it stood, (\d+) not used; (\d+) blocks", err)
    assert m, err
    return tuple(int(x) for x in m.groups()[1:])


@pytest.mark.parametrize("name", ["fastq_level6", "fastq_level1", "fastq_level9", "stored_blocks", "fixed_codes", "huffman_only", "run_length",
                                  "small_blocks", "incompressible", "zeros", "period_25", "short_periods", "members", "flushes",
                                  "mixed_block_types"])
def test_parallel_gzip_equals_zlib(exe, tmp_path, name):
    raw, comp = CASES[name]
    p = tmp_path / "x.gz"
    p.write_bytes(comp)
    for env in (PAR, {**PAR, "MERKURIO_GZIP_PIECE_KB": "5", "MERKURIO_GZIP_THREADS": "7"}, {**PAR, "MERKURIO_GZIP_PIECE_KB": "200", "MERKURIO_GZIP_THREADS": "2"}):
        if len(comp) < 3 * 1024 * int(env["MERKURIO_GZIP_PIECE_KB"]):
            continue  # (too small for this piece size: the sequential reader takes it)
        rc, out, err = cat(exe, p, None, env)
        assert rc == 0, (name, err)
        assert out == raw, name
        pieces, used, unused, seq = _stats(err)
        assert used + unused <= pieces  # (pieces behind the end of the last member's data are never looked at)
        if name.startswith("fastq_level"):
            assert used >= 2, (pieces, used, unused, seq)  # (pieces smaller than a block find the same block start: one of them is used)
        if name == "stored_blocks":
            assert used <= 1 and seq > 0  # nothing the search accepts: the sequential decoder does the work, correctly
        if name == "fixed_codes":
            assert used == 1 and seq == 0  # one block from the first bit to the last: the first task decodes all of it


def test_parallel_gzip_pieces_line_up(exe, tmp_path):
    """Pieces larger than a block, dynamic blocks all along (what a FASTQ from gzip looks like): every piece finds a block
    start and the real decode arrives exactly there, so the sequential decoder has nothing to do."""
    raw = fastq(30_000, 31)
    p = tmp_path / "x.gz"
    p.write_bytes(gz(raw))
    rc, out, err = cat(exe, p, None, {**PAR, "MERKURIO_GZIP_PIECE_KB": "200"})
    assert rc == 0 and out == raw
    pieces, used, unused, seq = _stats(err)
    assert pieces >= 8 and used == pieces and seq == 0, (pieces, used, unused, seq)


def test_parallel_gzip_reads_in_pieces_of_any_size(exe, tmp_path):
    raw, comp = CASES["fastq_level6"]
    p = tmp_path / "x.gz"
    p.write_bytes(comp)
    for chunk in (1000, 65536, 3_000_000):
        rc, out, err = cat(exe, p, chunk, PAR)
        assert rc == 0 and out == raw, chunk


def test_parallel_gzip_errors_are_the_sequential_reader_s(exe, tmp_path):
    """Cut or damaged files: the same exit status, the same message, and a prefix of the data in front of it."""
    rng = random.Random(8)
    raw = fastq(4000, 21)
    comp = gz(raw)
    p = tmp_path / "bad.gz"
    seq_env = {"MERKURIO_GZIP_THREADS": "1"}
    for cut in [len(comp) - 1, len(comp) - 4, len(comp) - 8, len(comp) - 9, len(comp) // 2, len(comp) // 3 + 1, 70_000, 50_001]:
        p.write_bytes(comp[:cut])
        rc, out, err = cat(exe, p, None, PAR)
        rc1, out1, err1 = cat(exe, p, None, seq_env)
        assert rc != 0 and rc1 != 0
        assert [l for l in err.splitlines() if l.startswith(b"#error")] == [l for l in err1.splitlines() if l.startswith(b"#error")], cut
        assert raw.startswith(out), cut
    for _ in range(40):
        c = bytearray(comp)
        c[rng.randrange(10, len(comp))] ^= 1 << rng.randrange(8)
        p.write_bytes(bytes(c))
        rc, out, err = cat(exe, p, None, PAR)
        try:
            ref = gzip.decompress(bytes(c))
        except Exception:
            ref = None
        if ref is None:
            assert rc != 0
        else:
            assert rc == 0 and out == ref
    p.write_bytes(comp + b"\0" * 16)  # bytes that are no member behind the last one
    rc, out, err = cat(exe, p, None, PAR)
    assert rc != 0 and out == raw
    c = bytearray(comp)
    c[-6] ^= 0x40  # the CRC-32 of the trailer
    p.write_bytes(bytes(c))
    rc, out, err = cat(exe, p, None, PAR)
    assert rc != 0 and out == raw and b"Error while decompressing the input (gzip)" in err


def test_parallel_gzip_through_the_fastq_reader(exe, tmp_path):
    """The chunked FASTQ reader on top of the parallel stream hands over the same records as from the plain file."""
    raw = fastq(8000, 5)
    plain = tmp_path / "r.fastq"
    plain.write_bytes(raw)
    zipped = tmp_path / "r.fastq.gz"
    zipped.write_bytes(gz(raw))
    want = subprocess.run([exe, "records", str(plain), "chunked"], capture_output=True)
    got = subprocess.run([exe, "records", str(zipped), "chunked"], capture_output=True, env={**os.environ, **PAR})
    assert want.returncode == 0 and got.returncode == 0
    assert got.stdout == want.stdout and want.stdout.count(b"#id\t") == 8000


def test_decoder_under_address_and_undefined_behaviour_sanitizers(tmp_path):
    """tests/stub/inflate_fuzz.cpp: valid, cut and bit-flipped streams of five kinds of data through every entry point of
    the decoder (run in pieces, inflate_exact, probe_dynamic_header, run_markers) with exactly the padding behind the
    input that inflate.h promises — built with -fsanitize=address,undefined, so one byte read or written outside the
    buffers ends the run; valid streams must decode to the original bytes."""
    exe = tmp_path / "inflate_fuzz"
    cc = subprocess.run(["g++", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-std=c++17",
                         "-I", str(ROOT / "merkurio_b200" / "host"), "-o", str(exe), str(ROOT / "tests" / "stub" / "inflate_fuzz.cpp"),
                         str(ROOT / "merkurio_b200" / "host" / "inflate.cpp"), "-lz"], capture_output=True, text=True)
    if cc.returncode != 0 and "sanitize" in cc.stderr:
        pytest.skip("the sanitizer runtimes are not installed")
    assert cc.returncode == 0, cc.stderr
    r = subprocess.run([str(exe), "20", "3"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "done:" in r.stdout, (r.stdout[-500:], r.stderr[-2000:])


def test_parallel_gzip_leaves_data_that_expands_a_thousandfold_to_the_sequential_decoder(exe, tmp_path):
    """A task holds 16-bit symbols for at most 12 times its piece of the file; 40 MB of zeros in 39 KB exceed that in every
    piece, so every task gives up and the sequential decoder produces the output (bounded memory, same bytes)."""
    raw = b"\0" * 40_000_000
    p = tmp_path / "z.gz"
    p.write_bytes(gz(raw))
    rc, out, err = cat(exe, p, None, {**PAR, "MERKURIO_GZIP_PIECE_KB": "5"})
    assert rc == 0 and out == raw
    pieces, used, unused, seq = _stats(err)
    assert used == 0 and seq > 0, (pieces, used, unused, seq)
