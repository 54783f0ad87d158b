"""CPU tests of the C++ host (merkurio_b200/host): query-list preprocessing, algorithm choice and
flag rules against the oracle and the reference's own helper tests (src/helpers.rs:218-568,
src/main.rs:60-293). Nothing here needs a GPU — these paths run before any engine is created."""
import gzip
import os
import random
import subprocess
import zlib
from pathlib import Path

import pytest

from oracle import refmodel as rm

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def exe():
    from merkurio_b200.build import build_host
    return str(build_host())


def run(exe, *args):
    return subprocess.run([exe, *map(str, args)], capture_output=True, text=True)


def cli_patterns(exe, *args):
    r = run(exe, "patterns", *args)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.split("\n")
    pats = [ln for ln in lines if ln and not ln.startswith("#search_algorithm")]
    algo = [ln.split("\t")[1] for ln in lines if ln.startswith("#search_algorithm")][0]
    return pats, algo


@pytest.mark.parametrize("flags,kw", [
    ((), {}),
    (("-r",), dict(reverse_complement_=True)),
    (("-c",), dict(canonical_=True)),
    (("-L",), dict(lowercase=True)),
    (("-U", "-r"), dict(uppercase=True, reverse_complement_=True)),
])
@pytest.mark.parametrize("name", ["kmers.txt", "kmers.fasta", "kmers-duplicates.txt", "kmers-messy.txt", "kmers-many-40.txt", "kmers-aa.txt"])
def test_pattern_list_matches_oracle(exe, ref_tree, name, flags, kw):
    f = ref_tree / "tests" / "data" / name
    want = rm.parse_pattern_list(f, None, kw.get("reverse_complement_", False), kw.get("canonical_", False),
                                 kw.get("lowercase", False), kw.get("uppercase", False))
    got, algo = cli_patterns(exe, "-f", f, *flags)
    assert got == want
    assert algo == ("Aho-Corasick" if rm.recommend_aho_corasick(want) else "BNDMq")


def test_pattern_list_from_sequences(exe):
    got, algo = cli_patterns(exe, "-s", "ACG", "-r")
    assert got == ["ACG", "CGT"] and algo == "BNDMq"
    got, _ = cli_patterns(exe, "-s", "ARYKMBVDHSWN*x", "acgtn", "-r")
    assert got == rm.parse_pattern_list(None, ["ARYKMBVDHSWN*x", "acgtn"], True, False, False, False)
    # palindromes collapse under -r (src/helpers.rs:124-126)
    got, _ = cli_patterns(exe, "-s", "ACGT", "AATT", "-r")
    assert got == ["AATT", "ACGT"]
    # canonical keeps only one orientation
    got, _ = cli_patterns(exe, "-s", "TTTT", "AAAA", "CCGG", "-c")
    assert got == ["AAAA", "CCGG"]


def test_algorithm_choice(exe):
    # src/helpers.rs:203-211 and src/cmd_extract.rs:165-171
    thirteen = [f"ACGT{'ACGT'[i % 4]}{'ACGT'[i // 4]}A" for i in range(13)]
    assert cli_patterns(exe, "-s", *thirteen)[1] == "BNDMq"
    assert cli_patterns(exe, "-s", *thirteen, "TTTTTTT")[1] == "Aho-Corasick"
    assert cli_patterns(exe, "-s", "A" * 64)[1] == "BNDMq"
    assert cli_patterns(exe, "-s", "A" * 65)[1] == "Aho-Corasick"
    assert cli_patterns(exe, "-s", "ACGT", "-a")[1] == "Aho-Corasick"
    assert cli_patterns(exe, "-s", "ACGT", "-I")[1] == "Aho-Corasick"
    assert cli_patterns(exe, "-s", *thirteen, "TTTTTTT", "-q", "3")[1] == "BNDMq"


def test_bndmq_errors_of_forced_q(exe):
    r = run(exe, "patterns", "-s", "ACGT", "-q", "7")
    assert r.returncode == 1 and "Invalid q-gram length: 7. Must be between 1 and pattern length." in r.stderr
    r = run(exe, "patterns", "-s", "A" * 65, "-q", "4")
    assert r.returncode == 1 and "is too large for this architecture when using BNDM (max 64)" in r.stderr
    r = run(exe, "patterns", "-s", "ACGT", "-q", "0")
    assert r.returncode == 1 and "Invalid q-gram length: 0" in r.stderr


def test_pattern_file_errors(exe, ref_tree, tmp_path):
    r = run(exe, "patterns", "-f", ref_tree / "tests" / "data" / "kmers-empty.txt")
    assert r.returncode == 1 and "No k-mers found in the file." in r.stderr and "Problem parsing pattern list." in r.stderr
    r = run(exe, "patterns", "-f", tmp_path / "missing.txt")
    assert r.returncode == 1 and "File not found." in r.stderr
    r = run(exe, "patterns", "-f", tmp_path)
    assert r.returncode == 1 and "is a directory, not a file." in r.stderr
    r = run(exe, "patterns", "-s", "")
    assert r.returncode == 1 and "No k-mers found in file or provided sequence." in r.stderr


def test_cfg1_literal_command_is_refused(exe, ref_tree):
    # BASELINE config 1 as written: log and records both on stdout (src/helpers.rs:195-197)
    em = ref_tree / "example-minimal"
    r = run(exe, "extract", "-i", em / "sample.fasta", "-f", em / "kmers.txt", "-r", "-l")
    assert r.returncode == 1
    assert r.stderr.strip() == ("Error: Cannot write log to stdout when normal output is also stdout. "
                                "Specify an output file with -o or suppress output with -S.")
    assert r.stdout == ""


def test_log_flag_conflicts(exe, ref_tree):
    fa = ref_tree / "tests" / "fixtures" / "input" / "simple.fasta"
    r = run(exe, "extract", "-i", fa, "-s", "ACG", "-l", "-j", "-o", "x")
    assert r.returncode == 1 and "both to stdout" in r.stderr
    r = run(exe, "extract", "-i", fa, "-s", "ACG", "-j")
    assert r.returncode == 1 and "Cannot write log to stdout when normal output is also stdout" in r.stderr
    sam = ref_tree / "tests" / "fixtures" / "input" / "simple.sam"
    r = run(exe, "tag", "-i", sam, "-s", "CTC", "-l")
    assert r.returncode == 1 and "Cannot write log to stdout" in r.stderr


def test_clap_grammar(exe, ref_tree):
    # src/main.rs:60-293: group exclusivity, -S rules, required arguments -> usage error, exit code 2
    fa = ref_tree / "tests" / "fixtures" / "input" / "simple.fasta"
    cases = [
        ("extract", "-i", fa, "-s", "A", "-f", "k.txt"),
        ("extract", "-i", fa, "-s", "A", "-r", "-c"),
        ("extract", "-i", fa, "-s", "A", "-I", "-L"),
        ("extract", "-i", fa, "-s", "A", "-L", "-U"),
        ("extract", "-i", fa, "-s", "A", "-q", "2", "-a"),
        ("extract", "-i", fa, "-s", "A", "-S"),                 # -S requires logging
        ("extract", "-i", fa, "-s", "A", "-S", "-l", "-o", "x"),  # -S conflicts with -o
        ("extract", "-i", fa),                                  # no k-mers
        ("extract", "-s", "A"),                                 # no input
        ("extract", "-i", fa, "-s", "A", "--bogus"),
        ("tag", "-i", "x.sam", "-s", "A", "-m", "-v"),
        ("tag", "-i", "x.sam", "-s", "A", "-p", "abc"),
        ("frobnicate",),
    ]
    for c in cases:
        r = run(exe, *c)
        assert r.returncode == 2, (c, r.returncode, r.stderr)
        assert r.stderr.startswith("error:"), c
    assert run(exe).returncode == 2
    assert run(exe, "--version").stdout.startswith("merkurio ")
    assert run(exe, "extract", "--help").returncode == 0


def test_tag_argument_checks(exe, ref_tree):
    sam = ref_tree / "tests" / "fixtures" / "input" / "simple.sam"
    r = run(exe, "tag", "-i", sam, "-s", "CTC", "-p", "0", "-o", "x.sam")
    assert r.returncode == 1 and "Number of threads must be at least 1." in r.stderr
    r = run(exe, "tag", "-i", sam, "-s", "CTC", "-t", "kmer", "-o", "x.sam")
    assert r.returncode == 1 and "Tag must be exactly two characters long." in r.stderr
    r = run(exe, "tag", "-i", ref_tree / "tests" / "data" / "sample.fasta", "-s", "CTC", "-o", "x.sam")
    assert r.returncode == 1 and "Input file must be a BAM or SAM file." in r.stderr


def test_no_cpu_fallback_in_the_cli(exe, ref_tree, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    fa = ref_tree / "tests" / "fixtures" / "input" / "simple.fasta"
    r = run(exe, "extract", "-i", fa, "-s", "ACG", "-o", tmp_path / "o")
    assert r.returncode == 1 and "no CPU path" in r.stderr


# ------------------------------------------------------------------------------------------------
# FASTQ ingest: the chunked reader (one thread per file, records indexed in place) must hand the
# matcher and the writer exactly what the line-by-line reader does, whatever the chunk size.
FASTQ_CASES = {
    "plain": b"@r1 desc\nACGT\n+\nIIII\n@r2\nGG\n+\nII\n",
    "no_final_newline": b"@r1\nACGT\n+\nIIII\n@r2\nGGA\n+\nIII",
    "plus_repeats_id": b"@r1\nACGT\n+r1\nIIII\n@r2\nGG\n+\nII\n",
    "crlf": b"@r1\r\nACGT\r\n+\r\nIIII\r\n@r2\r\nGG\r\n+\r\nII\r\n",
    "mixed_crlf": b"@r1\nACGT\r\n+\nIIII\n@r2\r\nGG\n+\nII\n",
    "blank_lines_between": b"@r1\nACGT\n+\nIIII\n\n\r\n@r2\nGG\n+\nII\n\n\n",
    "empty_sequence": b"@r1\n\n+\n\n@r2\nGG\n+\nII\n",
    "quality_starts_with_at": b"@r1\nACGT\n+\n@III\n@r2\nGG\n+\n@@\n",
    "truncated": b"@r1\nACGT\n+\nIIII\n@r2\nGG\n+\n",
    "length_mismatch": b"@r1\nACGT\n+\nIIII\n@r2\nGGA\n+\nII\n@r3\nA\n+\nI\n",
    "bad_header": b"@r1\nACGT\n+\nIIII\nr2\nGG\n+\nII\n",
    "missing_plus": b"@r1\nACGT\n-\nIIII\n",
    "only_header": b"@r1\n",
    "long_record": b"@long\n" + b"ACGTTGCA" * 5000 + b"\n+\n" + b"I" * 40000 + b"\n@r2\nGG\n+\nII\n",
    # longer than the reader's head room and than the stretch its indexer works on at a time
    "very_long_record": b"@r0\nAC\n+\nII\n@long\n" + b"ACGTTGCA" * 50000 + b"\n+\n" + b"I" * 400000 + b"\n@r2\nGG\n+\nII\n@r3\nT\n+\nI",
}


@pytest.mark.parametrize("name", sorted(FASTQ_CASES))
@pytest.mark.parametrize("gz", [False, True])
def test_chunked_fastq_reader_equals_line_reader(exe, tmp_path, name, gz):
    import gzip
    data = FASTQ_CASES[name]
    p = tmp_path / ("x.fastq.gz" if gz else "x.fastq")
    p.write_bytes(gzip.compress(data) if gz else data)
    want = subprocess.run([exe, "records", str(p), "generic"], capture_output=True)
    assert want.returncode == 0, want.stderr
    assert want.stdout.startswith(b"#id\t") or name in ("only_header", "missing_plus")
    for chunk in (4096, 5000, 1 << 20):
        got = subprocess.run([exe, "records", str(p), "chunked", str(chunk)], capture_output=True)
        assert got.returncode == 0, got.stderr
        assert got.stdout == want.stdout, (name, chunk)


def test_chunked_fastq_reader_many_records(exe, tmp_path):
    import numpy as np
    rng = np.random.default_rng(5)
    recs = []
    for i in range(20000):
        n = int(rng.integers(0, 300))
        s = bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=n).tobytes())
        recs.append(b"@read%d some text\n%s\n+\n%s\n" % (i, s, b"F" * n))
    p = tmp_path / "many.fastq"
    p.write_bytes(b"".join(recs))
    want = subprocess.run([exe, "records", str(p), "generic"], capture_output=True).stdout
    assert want.count(b"#id\t") == 20000
    for chunk in (4096, 70001, 1 << 22):
        got = subprocess.run([exe, "records", str(p), "chunked", str(chunk)], capture_output=True).stdout
        assert got == want


def test_fasta_is_left_to_the_line_reader(exe, tmp_path):
    p = tmp_path / "x.fasta"
    p.write_bytes(b">s1\nACGT\nAC\n>s2\nGG\n")
    got = subprocess.run([exe, "records", str(p), "chunked"], capture_output=True)
    assert got.stdout == b"#not-fastq\n"


def bgzf_compress(data: bytes, rng, eof_marker=True) -> bytes:
    """BGZF as bgzip / BAM writers produce it: gzip members of <= 64 KiB with a 'BC' extra field."""
    import struct
    import zlib
    out = bytearray()
    pos = 0
    chunks = []
    while pos < len(data):
        n = int(rng.integers(1, 65000))
        chunks.append(data[pos:pos + n])
        pos += n
    if eof_marker:
        chunks.append(b"")
    for ch in chunks:
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = co.compress(ch) + co.flush()
        bsize = 12 + 6 + len(body) + 8
        out += b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1)
        out += body + struct.pack("<II", zlib.crc32(ch), len(ch))
    return bytes(out)


def test_bgzf_input_is_inflated_block_parallel(exe, tmp_path):
    import numpy as np
    rng = np.random.default_rng(11)
    recs = []
    for i in range(30000):
        n = int(rng.integers(1, 400))
        s = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n).tobytes())
        recs.append(b"@r%d\n%s\n+\n%s\n" % (i, s, b"I" * n))
    data = b"".join(recs)
    plain = tmp_path / "p.fastq"
    plain.write_bytes(data)
    bg = tmp_path / "b.fastq.gz"
    bg.write_bytes(bgzf_compress(data, rng))
    want = subprocess.run([exe, "records", str(plain), "generic"], capture_output=True).stdout
    assert want.count(b"#id\t") == 30000
    for how in ("generic", "chunked"):
        got = subprocess.run([exe, "records", str(bg), how], capture_output=True)
        assert got.returncode == 0, got.stderr
        assert got.stdout == want, how
    # a flipped byte inside a block is caught by its CRC (or by inflate)
    bad = bytearray(bg.read_bytes())
    bad[len(bad) // 2] ^= 0x55
    (tmp_path / "bad.fastq.gz").write_bytes(bytes(bad))
    r = subprocess.run([exe, "records", str(tmp_path / "bad.fastq.gz"), "generic"], capture_output=True)
    assert r.returncode != 0 or b"#error" in r.stdout
    assert r.stdout != want


# ------------------------------------------------------------------------------------------------
# SAM / BAM ingest: the chunked record index must hand over what the record-by-record readers do.
def _sam_text(n=3000, seed=3):
    import numpy as np
    rng = np.random.default_rng(seed)
    out = [b"@HD\tVN:1.6\tSO:unsorted", b"@SQ\tSN:chr1\tLN:100000", b"@SQ\tSN:chr2\tLN:5000"]
    for i in range(n):
        L = int(rng.integers(0, 200))
        seq = bytes(rng.choice(np.frombuffer(b"ACGTNacgtRYK", dtype=np.uint8), size=L).tobytes()) if L else b"*"
        qual = b"F" * L if L else b"*"
        tags = b"\tNM:i:%d\tXS:Z:a,b;c" % (i % 7) + (b"\tkm:Z:AAA,CCC" if i % 40 == 0 else b"")
        out.append(b"r%d\t%d\t%s\t%d\t%d\t%s\t%s\t%d\t%d\t%s\t%s%s" % (
            i, 99 if i % 2 else 147, b"chr1" if i % 3 else b"chr2", 1 + i * 13 % 4000, i % 61, b"%dM" % L if L else b"*",
            b"=" if i % 4 else b"*", 1 + i * 7 % 4000 if i % 4 else 0, i % 300 - 150, seq, qual, tags))
        if i % 500 == 0:
            out.append(b"")  # blank lines are skipped
    return b"\n".join(out) + b"\n"


@pytest.mark.parametrize("variant", ["plain", "crlf", "no_final_newline", "truncated"])
def test_chunked_sam_reader_equals_line_reader(exe, tmp_path, variant):
    data = _sam_text()
    if variant == "crlf":
        data = data.replace(b"\n", b"\r\n")
    if variant == "no_final_newline":
        data = data.rstrip(b"\n")
    if variant == "truncated":
        cut = len(data) // 2
        data = data[:cut].rsplit(b"\t", 3)[0] + b"\n" + data[cut:]
    p = tmp_path / "x.sam"
    p.write_bytes(data)
    want = subprocess.run([exe, "alnrecords", str(p), "generic"], capture_output=True)
    assert want.returncode == 0, want.stderr
    assert (b"#error" in want.stdout) == (variant == "truncated")
    assert want.stdout.count(b"#name\t") >= (1000 if variant == "truncated" else 3000)
    for chunk in (4096, 30011, 1 << 22):
        got = subprocess.run([exe, "alnrecords", str(p), "chunked", str(chunk)], capture_output=True)
        assert got.returncode == 0, got.stderr
        assert got.stdout == want.stdout, (variant, chunk)


def _sam_to_bam(sam: bytes, rng) -> bytes:
    """Minimal SAM -> BAM encoder for test inputs (BGZF blocks of random size: records cross them)."""
    import struct
    lines = [ln for ln in sam.replace(b"\r\n", b"\n").split(b"\n") if ln]
    header = [ln for ln in lines if ln.startswith(b"@")]
    refs = [(f[1][3:], int(f[2][3:])) for f in (ln.split(b"\t") for ln in header) if f[0] == b"@SQ"]
    ref_id = {name: i for i, (name, _) in enumerate(refs)}
    text = b"\n".join(header) + b"\n"
    out = bytearray(b"BAM\x01" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(refs)))
    for name, ln in refs:
        out += struct.pack("<i", len(name) + 1) + name + b"\x00" + struct.pack("<i", ln)
    nib = {c: i for i, c in enumerate(b"=ACMGRSVTWYHKDBN")}
    for ln in lines[len(header):]:
        f = ln.split(b"\t")
        seq = b"" if f[9] == b"*" else f[9].upper()
        cigar = b"" if f[5] == b"*" else struct.pack("<I", (int(f[5][:-1]) << 4) | 0)
        rid = ref_id.get(f[2], -1)
        nrid = rid if f[6] == b"=" else ref_id.get(f[6], -1)
        codes = [nib.get(c, 15) for c in seq] + [0]
        packed = bytes((codes[i] << 4) | codes[i + 1] for i in range(0, len(seq), 2))
        qual = b"\xff" * len(seq) if f[10] == b"*" else bytes(c - 33 for c in f[10])
        aux = bytearray()
        for t in f[11:]:
            tag, typ, val = t[:2], t[3:4], t[5:]
            aux += tag + (b"C" + struct.pack("<B", int(val)) if typ == b"i" else b"Z" + val + b"\x00")
        body = struct.pack("<iiBBHHHiiii", rid, int(f[3]) - 1, len(f[0]) + 1, int(f[4]), 4680, len(cigar) // 4, int(f[1]), len(seq), nrid,
                           int(f[7]) - 1, int(f[8])) + f[0] + b"\x00" + cigar + packed + qual + bytes(aux)
        out += struct.pack("<i", len(body)) + body
    return bgzf_compress(bytes(out), rng)


def test_chunked_bam_reader_equals_record_reader(exe, ref_tree, tmp_path):
    import numpy as np
    bam = tmp_path / "x.bam"
    bam.write_bytes(_sam_to_bam(_sam_text(4000, seed=9), np.random.default_rng(4)))
    for src in [ref_tree / "tests" / "fixtures" / "input" / "simple.bam", bam]:
        want = subprocess.run([exe, "alnrecords", str(src), "generic"], capture_output=True)
        assert want.returncode == 0 and want.stdout.count(b"#name\t") > 0, want.stderr
        assert b"#error" not in want.stdout
        if src == bam:
            assert want.stdout.count(b"#name\t") == 4000 and b"\tkm:Z:AAA,CCC" in want.stdout
        for chunk in (4096, 50021, 1 << 22):
            got = subprocess.run([exe, "alnrecords", str(src), "chunked", str(chunk)], capture_output=True)
            assert got.returncode == 0, got.stderr
            assert got.stdout == want.stdout, (src.name, chunk)


@pytest.mark.parametrize("codec", ["gz", "gz_multi", "bz2", "bz2_multi", "xz", "xz_multi", "zstd", "zstd_multi"])
def test_compressed_inputs_are_recognised_by_their_magic_bytes(exe, tmp_path, codec):
    """needletail opens .gz / .bz2 / .xz / .zst by content, not by name (README.md:39; its "compression" feature,
    Cargo.toml:26, includes zstd): same records as the plain file."""
    import bz2
    import gzip
    import lzma
    import numpy as np
    if codec.startswith("zstd"):
        pa = pytest.importorskip("pyarrow")  # the only zstd encoder at hand
        zstd = lambda d: pa.compress(d, codec="zstd", asbytes=True)
    else:
        zstd = None
    rng = np.random.default_rng(13)
    recs = []
    for i in range(6000):
        n = int(rng.integers(1, 300))
        s = bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=n).tobytes())
        recs.append(b"@r%d\n%s\n+\n%s\n" % (i, s, b"I" * n))
    data = b"".join(recs)
    half = data.index(b"\n@r3000\n") + 1
    packed = {"gz": gzip.compress(data), "gz_multi": gzip.compress(data[:half]) + gzip.compress(data[half:]), "bz2": bz2.compress(data), "bz2_multi": bz2.compress(data[:half]) + bz2.compress(data[half:]),
              "xz": lzma.compress(data), "xz_multi": lzma.compress(data[:half]) + lzma.compress(data[half:]),
              "zstd": zstd and zstd(data), "zstd_multi": zstd and zstd(data[:half]) + zstd(data[half:])}[codec]
    plain = tmp_path / "p.fastq"
    plain.write_bytes(data)
    comp = tmp_path / "reads.fastq.dat"  # the name says nothing
    comp.write_bytes(packed)
    want = subprocess.run([exe, "records", str(plain), "generic"], capture_output=True).stdout
    assert want.count(b"#id\t") == 6000
    for how in ("generic", "chunked"):
        got = subprocess.run([exe, "records", str(comp), how, "100000"], capture_output=True)
        assert got.returncode == 0, got.stderr
        assert got.stdout == want, (codec, how)
    # a truncated stream is an error, not a silently shorter input
    (tmp_path / "cut.dat").write_bytes(packed[: len(packed) // 2])
    r = subprocess.run([exe, "records", str(tmp_path / "cut.dat"), "generic"], capture_output=True)
    assert r.returncode != 0 or b"#error" in r.stdout
    # FASTA through the same streams
    fa = b"".join(b">s%d\n%s\n" % (i, b"ACGTTGCA" * 20) for i in range(500))
    (tmp_path / "g.fa").write_bytes(fa)
    (tmp_path / "g.fa.z").write_bytes({"gz": gzip.compress, "gz_multi": gzip.compress, "bz2": bz2.compress, "bz2_multi": bz2.compress, "xz": lzma.compress, "xz_multi": lzma.compress,
                                       "zstd": zstd, "zstd_multi": zstd}[codec](fa))
    a = subprocess.run([exe, "records", str(tmp_path / "g.fa"), "generic"], capture_output=True).stdout
    b = subprocess.run([exe, "records", str(tmp_path / "g.fa.z"), "generic"], capture_output=True).stdout
    assert a == b and a.count(b"#id\t") == 500


def test_chunked_fastq_reader_fuzz(exe, tmp_path):
    """Random FASTQ with random damage (cut ranges, inserted bytes, CRLF islands, doubled newlines): the
    chunked reader and the line reader must print the same records and stop with the same error at the
    same record, whatever the chunk size."""
    import numpy as np
    rng = np.random.default_rng(2024)
    for trial in range(120):
        recs = []
        for i in range(int(rng.integers(1, 400))):
            n = int(rng.integers(0, 120))
            s = bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=n).tobytes())
            le = b"\r\n" if rng.random() < 0.05 else b"\n"
            plus = b"+" + (b"x%d" % i if rng.random() < 0.1 else b"")
            q = bytes(rng.choice(np.frombuffer(b"@+IF#", dtype=np.uint8), size=n).tobytes())
            recs.append(b"@q%d/%d" % (trial, i) + le + s + le + plus + le + q + le + (b"\n" if rng.random() < 0.05 else b""))
        data = bytearray(b"".join(recs))
        kind = trial % 4
        if kind == 1 and len(data) > 10:  # cut a range
            a = int(rng.integers(0, len(data) - 1))
            del data[a:a + int(rng.integers(1, 60))]
        elif kind == 2 and len(data) > 10:  # insert bytes
            a = int(rng.integers(0, len(data)))
            data[a:a] = bytes(rng.choice(np.frombuffer(b"@+\nACGT\r", dtype=np.uint8), size=int(rng.integers(1, 8))).tobytes())
        elif kind == 3:  # no final newline
            while data and data[-1] in b"\r\n":
                data.pop()
        if not data or data[0] != ord("@"):
            continue  # the chunked reader only takes files that start like FASTQ
        p = tmp_path / ("f%d.fastq" % trial)
        p.write_bytes(bytes(data))
        want = subprocess.run([exe, "records", str(p), "generic"], capture_output=True)
        assert want.returncode == 0, want.stderr
        for chunk in (4096, int(rng.integers(4097, 20000))):
            # every third file also through the SSE2 flavour of the line-break scanner
            env = dict(os.environ, MERKURIO_NO_AVX2="1") if trial % 3 == 0 and chunk == 4096 else None
            got = subprocess.run([exe, "records", str(p), "chunked", str(chunk)], capture_output=True, env=env)
            assert got.returncode == 0, got.stderr
            assert got.stdout == want.stdout, (trial, kind, chunk)


def test_chunked_fastq_reader_with_several_read_helpers(exe, tmp_path):
    """Blocks of 1 MiB and more have their line breaks — and what stands either side of each (BlockReader::kNl*) — located
    by several helper threads, a slice each; the indexer recognises ordinary records from those notes alone. Messy
    records (CRLF islands, '+id' lines, blank lines, qualities that start with '@' or '+', damage) at every slice and
    block boundary must still come out as the line reader has them."""
    import numpy as np
    rng = np.random.default_rng(77)
    for trial in range(6):
        recs = []
        for i in range(30000):
            n = int(rng.integers(0, 90))
            s = bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=n).tobytes())
            le = b"\r\n" if rng.random() < 0.01 else b"\n"
            plus = b"+" + (b"x%d" % i if rng.random() < 0.02 else b"")
            q = bytes(rng.choice(np.frombuffer(b"@+IF#", dtype=np.uint8), size=n).tobytes())
            recs.append(b"@q%d/%d" % (trial, i) + le + s + le + plus + le + q + le + (b"\n" if rng.random() < 0.01 else b""))
        data = bytearray(b"".join(recs))
        if trial % 3 == 1:
            a = int(rng.integers(len(data) // 2, len(data) - 100))
            del data[a:a + int(rng.integers(1, 60))]
        elif trial % 3 == 2:
            while data and data[-1] in b"\r\n":
                data.pop()
        p = tmp_path / ("h%d.fastq" % trial)
        p.write_bytes(bytes(data))
        want = subprocess.run([exe, "records", str(p), "generic"], capture_output=True)
        assert want.returncode == 0 and want.stdout.count(b"#id\t") > 10000, want.stderr
        for threads, chunk in ((3, 1 << 20), (5, (1 << 20) + 4097), (2, 3 << 20)):
            env = dict(os.environ, MERKURIO_READ_THREADS=str(threads))
            got = subprocess.run([exe, "records", str(p), "chunked", str(chunk)], capture_output=True, env=env)
            assert got.returncode == 0, got.stderr
            assert got.stdout == want.stdout, (trial, threads, chunk)


def test_chunked_sam_reader_fuzz(exe, tmp_path):
    """Damaged SAM text: both alignment readers stop at the same record with the same message."""
    import numpy as np
    rng = np.random.default_rng(7)
    base = _sam_text(300, seed=21)
    for trial in range(60):
        data = bytearray(base)
        for _ in range(int(rng.integers(0, 3))):
            a = int(rng.integers(60, len(data) - 1))  # past the header
            if rng.random() < 0.5:
                del data[a:a + int(rng.integers(1, 40))]
            else:
                data[a:a] = bytes(rng.choice(np.frombuffer(b"\t\n\rACGT*", dtype=np.uint8), size=int(rng.integers(1, 6))).tobytes())
        p = tmp_path / ("s%d.sam" % trial)
        p.write_bytes(bytes(data))
        want = subprocess.run([exe, "alnrecords", str(p), "generic"], capture_output=True)
        assert want.returncode == 0, want.stderr
        env = dict(os.environ, MERKURIO_NO_AVX2="1") if trial % 3 == 0 else None  # the SSE2 flavour of the separator scan
        got = subprocess.run([exe, "alnrecords", str(p), "chunked", str(int(rng.integers(4096, 9000)))], capture_output=True, env=env)
        assert got.returncode == 0, got.stderr
        assert got.stdout == want.stdout, trial


def test_truncated_gzip_is_an_error_wherever_it_is_cut(exe, tmp_path):
    """zlib's gzread reports an incomplete stream only from gzclose(): a FASTA cut anywhere, or a FASTQ cut where the
    decompressed bytes happen to end on a record boundary, must not pass for a complete, shorter input."""
    rng = random.Random(9)
    fa = b"".join(b">s%d\n%s\n" % (i, bytes(rng.choice(b"ACGT") for _ in range(120))) for i in range(3000))
    z = gzip.compress(fa)
    (tmp_path / "cut.fa.gz").write_bytes(z[: len(z) * 2 // 3])
    r = subprocess.run([exe, "records", str(tmp_path / "cut.fa.gz"), "generic"], capture_output=True)  # (the chunked reader of this command is FASTQ only)
    assert r.returncode != 0 or b"#error" in r.stdout
    # FASTQ: stored (uncompressed) deflate blocks, cut right after a block that ends on a record boundary
    rec = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, b"ACGT" * 25, b"I" * 100) for i in range(2000))
    co = zlib.compressobj(0, zlib.DEFLATED, 31)
    cut_at = rec.index(b"@r1000\n")
    first = co.compress(rec[:cut_at]) + co.flush(zlib.Z_FULL_FLUSH)  # ends on a record boundary, no trailer
    (tmp_path / "cut.fastq.gz").write_bytes(first)
    for how in ("generic", "chunked"):
        r = subprocess.run([exe, "records", str(tmp_path / "cut.fastq.gz"), how], capture_output=True)
        assert r.returncode != 0 or b"#error" in r.stdout, how
    # the complete stream of the same object is fine
    (tmp_path / "ok.fastq.gz").write_bytes(first + co.compress(rec[cut_at:]) + co.flush())
    r = subprocess.run([exe, "records", str(tmp_path / "ok.fastq.gz"), "generic"], capture_output=True)
    assert r.returncode == 0 and r.stdout.count(b"#id\t") == 2000
