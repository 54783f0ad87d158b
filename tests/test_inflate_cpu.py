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
