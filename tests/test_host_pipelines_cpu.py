"""CPU tests of the host's ingest pipelines (reader -> packer -> driver -> writer) with a test double
of the C ABI that reports no hits (tests/stub/stub_engine.c, preloaded): with `-v` every record is
then written, so the output shows exactly what the pipelines read, how they cut batches and pieces,
and what the writers emit. Each pipeline must agree byte for byte with the record-by-record path,
whatever the chunk and batch sizes. (Matching itself is only ever tested on the GPU.)"""
import gzip
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
QUERY = "ACGTACGTACGTACGTACGTACGTACGTAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAC"


@pytest.fixture(scope="module")
def exe():
    from merkurio_b200.build import build_host
    return str(build_host())


@pytest.fixture(scope="module")
def stub(tmp_path_factory):
    out = tmp_path_factory.mktemp("stub") / "libstub_engine.so"
    subprocess.run(["gcc", "-O1", "-shared", "-fPIC", "-I", str(ROOT / "include"), "-o", str(out), str(ROOT / "tests" / "stub" / "stub_engine.c")],
                   check=True)
    return str(out)


def run(exe, stub, args, env=None):
    e = dict(os.environ, LD_PRELOAD=stub)
    e.update(env or {})
    return subprocess.run([exe, *map(str, args)], capture_output=True, env=e)


def log_body(p: Path) -> bytes:
    return b"\n".join(p.read_bytes().split(b"\n")[4:])


SIZES = [{}, {"MERKURIO_BATCH_BYTES": "20000", "MERKURIO_CHUNK_BYTES": "5000"},
         {"MERKURIO_BATCH_BYTES": "17000", "MERKURIO_CHUNK_BYTES": "4096", "MERKURIO_SLOTS": "1"},
         # batches dealt over two engines (created side by side), results consumed in batch order
         {"MERKURIO_GPUS": "2", "MERKURIO_BATCH_BYTES": "30000", "MERKURIO_CHUNK_BYTES": "9000"}]


def fasta_text(rng, crlf=False, blanks=False, final_nl=True, width=60):
    out = bytearray()
    le = b"\r\n" if crlf else b"\n"
    for i in range(40):
        n = int(rng.integers(0, 9000)) if i % 7 else int(rng.integers(40000, 90000))
        s = bytes(rng.choice(np.frombuffer(b"ACGTNacgt", dtype=np.uint8), size=n).tobytes())
        out += b">rec%d some description" % i + le
        for k in range(0, n, width):
            out += s[k:k + width] + le
            if blanks and rng.random() < 0.02:
                out += le
        if blanks and i % 5 == 0:
            out += le + le
    return bytes(out) if final_nl else bytes(out).rstrip(b"\r\n")


@pytest.mark.parametrize("kw", [{}, {"crlf": True}, {"blanks": True}, {"final_nl": False}, {"width": 7},
                                {"crlf": True, "blanks": True, "final_nl": False}], ids=lambda k: "-".join(k) or "plain")
def test_fasta_pipeline_equals_record_path(exe, stub, tmp_path, kw):
    rng = np.random.default_rng(3)
    src = tmp_path / "g.fa"
    src.write_bytes(fasta_text(rng, **kw))
    results = []
    for env in [{"MERKURIO_NO_FASTA_PIPELINE": "1"}] + SIZES:
        o, lg = tmp_path / ("o%d.fa" % len(results)), tmp_path / ("l%d.log" % len(results))
        r = run(exe, stub, ["extract", "-i", src, "-s", QUERY, "-v", "-o", o, "-l", lg], env)
        assert r.returncode == 0, r.stderr
        results.append((o.read_bytes(), log_body(lg)))
    assert results[0][0].count(b">rec") == 40 and b"records searched: 40" in results[0][1]
    for other in results[1:]:
        assert other == results[0]


def paired_fasta_text(rng, n, tag, crlf=False, width=60, long_every=0, blanks=False, final_nl=True):
    out = bytearray()
    le = b"\r\n" if crlf else b"\n"
    for i in range(n):
        m = int(rng.integers(0, 400)) if not long_every or i % long_every else int(rng.integers(30000, 70000))
        s = bytes(rng.choice(np.frombuffer(b"ACGTNacgt", dtype=np.uint8), size=m).tobytes())
        out += b">pair%d/%s" % (i, tag) + le
        for k in range(0, m, width):
            out += s[k:k + width] + le
            if blanks and rng.random() < 0.05:
                out += le
        if blanks and i % 9 == 0:
            out += le + le
    return bytes(out) if final_nl else bytes(out).rstrip(b"\r\n")


@pytest.mark.parametrize("kw", [{}, {"crlf": True}, {"width": 10 ** 6}, {"long_every": 50}, {"blanks": True, "final_nl": False},
                                {"crlf": True, "blanks": True, "final_nl": False, "long_every": 70, "width": 7}], ids=lambda k: "-".join(k) or "plain")
def test_paired_fasta_pipeline_equals_record_path(exe, stub, tmp_path, kw):
    """Two FASTA files of mates (src/cmd_extract.rs:412-418, :463-607) go through the chunked FASTA pipeline as well: the
    same two output files and the same log as the record-by-record path, whatever the chunk and batch sizes, with records
    that span batches (`long_every`) in either file."""
    rng = np.random.default_rng(5)
    a, b = tmp_path / "m_1.fa", tmp_path / "m_2.fa"
    a.write_bytes(paired_fasta_text(rng, 600, b"1", **kw))
    b.write_bytes(b"\n\n" + paired_fasta_text(rng, 600, b"2", **kw))  # (blank lines in front of the first record)
    results = []
    for env in [{"MERKURIO_NO_FASTA_PIPELINE": "1"}] + SIZES:
        d = tmp_path / ("r%d" % len(results))
        d.mkdir()
        r = run(exe, stub, ["extract", "-i", a, "-2", b, "-s", QUERY, "-v", "-o", d / "o.fa", "-l", d / "l.log"], env)
        assert r.returncode == 0, r.stderr
        results.append(((d / "o_1.fa").read_bytes(), (d / "o_2.fa").read_bytes(), log_body(d / "l.log")))
    assert results[0][0].count(b">pair") == 600 and results[0][1].count(b">pair") == 600
    assert b"records searched: 1200" in results[0][2]
    for other in results[1:]:
        assert other == results[0]


@pytest.mark.parametrize("case", ["second_shorter", "second_longer", "second_cut_gz", "first_cut_gz"])
def test_paired_fasta_pipeline_fails_like_the_record_path(exe, stub, tmp_path, case):
    """Files with different numbers of records and inputs that cannot be read to their end stop the paired FASTA pipeline
    with the messages and the output of the record-by-record path (everything in front of the error is written)."""
    rng = np.random.default_rng(6)
    t1, t2 = paired_fasta_text(rng, 3000, b"1"), paired_fasta_text(rng, 3000, b"2")
    a, b = tmp_path / "m_1.fa", tmp_path / "m_2.fa"
    if case == "second_shorter":
        t2 = t2[: t2.index(b">pair2000/")]
    elif case == "second_longer":
        t1 = t1[: t1.index(b">pair2000/")]
    if case == "second_cut_gz":
        z = gzip.compress(t2)
        b = tmp_path / "m_2.fa.gz"
        t2 = z[: len(z) // 2]
    if case == "first_cut_gz":
        z = gzip.compress(t1)
        a = tmp_path / "m_1.fa.gz"
        t1 = z[: len(z) // 2]
    a.write_bytes(t1)
    b.write_bytes(t2)
    results = []
    for env in [{"MERKURIO_NO_FASTA_PIPELINE": "1"}, {}, {"MERKURIO_BATCH_BYTES": "20000", "MERKURIO_CHUNK_BYTES": "5000"}]:
        d = tmp_path / ("r%d" % len(results))
        d.mkdir()
        r = run(exe, stub, ["extract", "-i", a, "-2", b, "-s", QUERY, "-v", "-o", d / "o.fa"], env)
        assert r.returncode != 0
        results.append((r.returncode, r.stderr, (d / "o_1.fa").read_bytes(), (d / "o_2.fa").read_bytes()))
    assert results[0][2].count(b">pair") >= 1000, results[0][1]
    for other in results[1:]:
        assert other == results[0]


@pytest.mark.parametrize("kw", [{}, {"crlf": True, "blanks": True, "final_nl": False}, {"width": 7}, {"width": 300}], ids=lambda k: "-".join(k) or "plain")
def test_fasta_packer_puts_the_same_bases_into_the_slots(exe, stub, tmp_path, kw):
    """What the FASTA packer (single file and two files of mates) copies into the engine's slots — sequence lines without
    their line breaks, whole lines in a tight loop, the rest by the general rules — hashed by the test double
    (MK_STUB_DIGEST_FILE) and compared with the hash of the records computed here. Records are not cut in these runs
    (a cut record repeats bases in its next piece); chunks are small, so runs of lines end everywhere."""
    rng = np.random.default_rng(31)
    texts = [paired_fasta_text(rng, 800, b"1", **kw), paired_fasta_text(rng, 800, b"2", **kw)]

    def fnv(b):
        h = 1469598103934665603
        for c in b:
            h = ((h ^ c) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        return h

    def records(text):
        out = []
        for block in text.replace(b"\r", b"").split(b">")[1:]:
            out.append(b"".join(block.split(b"\n")[1:]))
        return out

    files = []
    for f, t in enumerate(texts):
        p = tmp_path / ("m_%d.fa" % (f + 1))
        p.write_bytes(t)
        files.append(p)
    for paired in (False, True):
        recs = records(texts[0]) + (records(texts[1]) if paired else [])
        want = "%d %d %016x" % (len(recs), sum(map(len, recs)), sum(map(fnv, recs)) & 0xFFFFFFFFFFFFFFFF)
        args = ["extract", "-i", files[0]] + (["-2", files[1]] if paired else []) + ["-s", QUERY, "-o", tmp_path / "o.fa"]
        for env in ({}, {"MERKURIO_CHUNK_BYTES": "4096"}, {"MERKURIO_CHUNK_BYTES": "5000", "MERKURIO_GPUS": "2"}):
            dg = tmp_path / "digest.txt"
            dg.unlink(missing_ok=True)
            r = run(exe, stub, args, dict(env, MK_STUB_DIGEST_FILE=str(dg)))
            assert r.returncode == 0, r.stderr
            assert dg.read_text().strip() == want, (paired, env)


def fastq_text(rng, n, prefix, crlf=False, odd=False):
    out = bytearray()
    le = b"\r\n" if crlf else b"\n"
    for i in range(n):
        L = int(rng.integers(0, 200))
        s = bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=L).tobytes())
        plus = b"+" + (b"%s%d" % (prefix, i) if odd and i % 3 == 0 else b"")
        out += b"@%s%d d=%d" % (prefix, i, i) + le + s + le + plus + le + b"F" * L + le
        if odd and i % 11 == 0:
            out += b"\n"
    return bytes(out)


@pytest.mark.parametrize("flavour", ["plain", "crlf", "odd", "gz", "truncated"])
def test_fastq_pipeline_equals_record_path(exe, stub, tmp_path, flavour):
    rng = np.random.default_rng(9)
    d1 = fastq_text(rng, 3000, b"a", crlf=flavour == "crlf", odd=flavour == "odd")
    d2 = fastq_text(rng, 3000, b"b", crlf=flavour == "crlf", odd=flavour == "odd")
    if flavour == "truncated":
        d2 = d2[: len(d2) // 2]
    p1, p2 = tmp_path / "r1.fq", tmp_path / "r2.fq"
    p1.write_bytes(gzip.compress(d1) if flavour == "gz" else d1)
    p2.write_bytes(gzip.compress(d2) if flavour == "gz" else d2)
    # gz: also with both files through the parallel gzip reader (host/pgzip.cpp), in pieces of 8 KiB
    par_gz = [{"MERKURIO_GZIP_THREADS": "3", "MERKURIO_GZIP_PIECE_KB": "8", "MERKURIO_BATCH_BYTES": "20000", "MERKURIO_CHUNK_BYTES": "5000"}] if flavour == "gz" else []
    for paired in (False, True):
        results = []
        for env in [{"MERKURIO_NO_FASTQ_PIPELINE": "1"}] + SIZES + par_gz:
            d = tmp_path / ("out%d%d" % (paired, len(results)))
            d.mkdir()
            args = ["extract", "-i", p1, "-s", QUERY, "-v", "-o", d / "x.fastq", "-l", d / "x.log"] + (["-2", p2] if paired else [])
            r = run(exe, stub, args, env)
            files = {f.name: (log_body(f) if f.suffix == ".log" else f.read_bytes()) for f in sorted(d.iterdir())}
            results.append((r.returncode, r.stderr, files))
        assert (results[0][0] != 0) == (paired and flavour == "truncated")
        # (the output takes the input's extension: x.fq, x_1.fq / x_2.fq)
        assert sum(v.count(b"\n@a") for k, v in results[0][2].items() if k.endswith(".fq")) > 1000
        for other in results[1:]:
            assert other == results[0]


@pytest.mark.parametrize("paired", [False, True])
def test_fastq_packer_puts_the_same_bases_into_the_slots(exe, stub, tmp_path, paired):
    """The packer's copies run on a pool of threads while it goes on choosing records (CopyPool: begin / publish / finish).
    What arrives in the engine's slots — the test double sums a hash of every record's bases (MK_STUB_DIGEST_FILE) — must
    be the same with one thread, with several, and on the record-by-record path, and equal the hash computed here."""
    rng = np.random.default_rng(12)
    n = 30000
    reads = [bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), size=int(rng.integers(0, 250))).tobytes()) for _ in range(n * (2 if paired else 1))]

    def fnv(b):
        h = 1469598103934665603
        for c in b:
            h = ((h ^ c) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        return h

    want = "%d %d %016x" % (len(reads), sum(map(len, reads)), sum(map(fnv, reads)) & 0xFFFFFFFFFFFFFFFF)
    files = []
    for f in range(2 if paired else 1):
        p = tmp_path / ("r_%d.fastq" % (f + 1))
        mine = reads[f::2] if paired else reads
        p.write_bytes(b"".join(b"@r%d\n%s\n+\n%s\n" % (i, r, b"I" * len(r)) for i, r in enumerate(mine)))
        files.append(p)
    args = ["extract", "-i", files[0]] + (["-2", files[1]] if paired else []) + ["-s", QUERY, "-o", tmp_path / "o.fastq"]
    for env in ({"MERKURIO_PACK_THREADS": "1"}, {"MERKURIO_PACK_THREADS": "4"}, {"MERKURIO_PACK_THREADS": "7", "MERKURIO_BATCH_BYTES": "3000000"},
                {"MERKURIO_PACK_THREADS": "3", "MERKURIO_BATCH_BYTES": "50000", "MERKURIO_CHUNK_BYTES": "20000"}, {"MERKURIO_NO_FASTQ_PIPELINE": "1"}):
        dg = tmp_path / "digest.txt"
        dg.unlink(missing_ok=True)
        r = run(exe, stub, args, dict(env, MK_STUB_DIGEST_FILE=str(dg)))
        assert r.returncode == 0, r.stderr
        assert dg.read_text().strip() == want, env


def test_fastq_record_larger_than_a_batch_ends_the_run(exe, stub, tmp_path):
    """A read that does not fit a slot stops the FASTQ pipeline with a message that names the remedy — after the copies of
    the records in front of it have been carried out (the pool is not left running behind the exception), no hang."""
    rng = np.random.default_rng(8)
    recs = [b"@r%d\n%s\n+\n%s\n" % (i, b"ACGT" * 25, b"I" * 100) for i in range(9000)]
    big = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=300000).tobytes())
    recs.insert(7000, b"@big\n" + big + b"\n+\n" + b"I" * len(big) + b"\n")
    p = tmp_path / "r.fastq"
    p.write_bytes(b"".join(recs))
    r = run(exe, stub, ["extract", "-i", p, "-s", QUERY, "-v", "-o", tmp_path / "o.fastq"], {"MERKURIO_BATCH_BYTES": "200000", "MERKURIO_PACK_THREADS": "4"})
    assert r.returncode != 0 and b"does not fit a batch (raise MERKURIO_BATCH_MB)" in r.stderr, r.stderr


@pytest.mark.parametrize("n_reads,level", [(3000, 9), (40000, 1)])
def test_fastq_pipeline_reports_read_errors_like_the_record_path(exe, stub, tmp_path, n_reads, level):
    """A gzip stream that breaks half way (flipped bytes): both paths stop with the same status and the same
    chain of messages — the decompression error under the context of the file it came from — after writing the
    same records."""
    rng = np.random.default_rng(10)
    good = gzip.compress(fastq_text(rng, n_reads, b"a"), compresslevel=level)
    bad = bytearray(good)
    bad[len(bad) // 2] ^= 0xFF
    bad[len(bad) // 2 + 1] ^= 0x55
    pg, pb = tmp_path / "good.fq.gz", tmp_path / "bad.fq.gz"
    pg.write_bytes(good)
    pb.write_bytes(bytes(bad))
    for case, (files, ctx) in enumerate((([pb], b"Error during FASTQ/A record parsing."), ([pb, pg], b"parsing of first file"),
                                         ([pg, pb], b"parsing of second file"))):
        seen = []
        for env in [{"MERKURIO_NO_FASTQ_PIPELINE": "1"}] + SIZES:
            d = tmp_path / ("e%d_%d" % (case, len(seen)))
            d.mkdir()
            args = ["extract", "-i", files[0], "-s", QUERY, "-v", "-o", d / "x.fastq"] + (["-2", files[1]] if len(files) > 1 else [])
            r = run(exe, stub, args, env)
            seen.append((r.returncode, r.stderr, {f.name: f.read_bytes() for f in sorted(d.iterdir())}))
        # (flipped bytes either break the deflate stream or decode to garbage that no longer parses)
        assert seen[0][0] == 1 and ctx in seen[0][1] and b"Caused by:" in seen[0][1]
        if n_reads == 3000:  # (which of the two depends on how far inflate gets before it meets an invalid code)
            assert b"decompressing" in seen[0][1] or seen[0][1].count(b"record parsing") >= 1
        else:  # the records in front of the damage (less zlib's last buffer) were written, by both paths alike
            assert sum(v.count(b"\n@a") for v in seen[0][2].values()) > 500
        for other in seen[1:]:
            assert other == seen[0]


@pytest.mark.parametrize("flags", [[], ["-v"], ["-l", "@/t.log", "-j", "@/t.json"]])
def test_sam_pipeline_equals_record_path(exe, stub, tmp_path, flags):
    from tests.test_cli_cpu import _sam_text
    src = tmp_path / "in.sam"
    src.write_bytes(_sam_text(3000, seed=5))
    results = []
    for env in [{"MERKURIO_NO_ALN_PIPELINE": "1"}] + SIZES:
        d = tmp_path / ("o%d" % len(results))
        d.mkdir()
        r = run(exe, stub, ["tag", "-i", src, "-s", QUERY[:31], "-o", d / "t.sam"] + [f.replace("@", str(d)) for f in flags], env)
        assert r.returncode == 0, r.stderr
        files = {}
        for f in sorted(d.iterdir()):
            data = f.read_bytes()
            if f.suffix == ".log":
                data = log_body(f)
            elif f.suffix == ".sam":
                data = b"\n".join(ln for ln in data.split(b"\n") if not ln.startswith(b"@PG"))
            elif f.suffix == ".json":
                import json
                j = json.loads(data)
                j["meta_information"] = {k: v for k, v in j["meta_information"].items() if k not in ("timestamp", "command_line")}
                data = json.dumps(j, sort_keys=True).encode()
            files[f.name] = data
        results.append(files)
    # nothing matched: every record is kept and gets a tag (records that already had one keep the old field too)
    assert results[0]["t.sam"].count(b"\tkm:Z:") == 3000 + 75
    assert sum(1 for ln in results[0]["t.sam"].split(b"\n") if ln and not ln.startswith(b"@")) == 3000
    for other in results[1:]:
        assert other == results[0]


@pytest.mark.parametrize("damage", ["truncated", "flipped"])
def test_bam_pipeline_stops_where_the_record_path_stops(exe, stub, tmp_path, damage):
    """A BAM file that breaks off / has a corrupt BGZF block three quarters in: both paths write every record
    in front of the broken block (the reference reads block by block) and fail with the same message."""
    from tests.test_cli_cpu import _sam_text
    sam = tmp_path / "in.sam"
    sam.write_bytes(_sam_text(60000, seed=8))
    bam = tmp_path / "in.bam"
    r = run(exe, stub, ["tag", "-i", sam, "-s", QUERY[:31], "-o", bam])
    assert r.returncode == 0, r.stderr
    data = bytearray(bam.read_bytes())
    at = len(data) * 3 // 4
    if damage == "truncated":
        del data[at:]
    else:
        data[at] ^= 0xFF
        data[at + 1] ^= 0x55
    broken = tmp_path / "broken.bam"
    broken.write_bytes(bytes(data))
    seen = []
    for env in [{"MERKURIO_NO_ALN_PIPELINE": "1"}, {}, {"MERKURIO_BATCH_BYTES": "20000", "MERKURIO_CHUNK_BYTES": "5000"}]:
        out = tmp_path / ("o%d.sam" % len(seen))
        r = run(exe, stub, ["tag", "-i", broken, "-s", QUERY[:31], "-o", out], env)
        body = b"\n".join(ln for ln in out.read_bytes().split(b"\n") if not ln.startswith(b"@PG"))
        seen.append((r.returncode, r.stderr, body))
    assert seen[0][0] == 1 and b"Error during BAM record parsing: " in seen[0][1]
    n_written = sum(1 for ln in seen[0][2].split(b"\n") if ln and not ln.startswith(b"@"))
    assert 30000 < n_written < 60000
    for other in seen[1:]:
        assert other == seen[0]


# ------------------------------------------------------------------------------------------------
# The same comparisons with flagged and unflagged records mixed: the test double then flags every record that
# starts with a given base (tests/stub/stub_engine.c, MK_STUB_FLAG_FIRST) and, with logs on, reports a made-up
# hit for it. Which records those are does not depend on how a path cuts its batches, so every path must
# write, tag and log the same ones.
@pytest.mark.parametrize("extra", [[], ["-v"], ["-l", "@/x.log", "-j", "@/x.json"], ["-v", "-l", "@/x.log"]], ids=lambda e: "_".join(e).replace("@/", "") or "plain")
@pytest.mark.parametrize("flavour", ["plain", "odd"])
def test_fastq_paths_agree_on_flagged_records(exe, stub, tmp_path, flavour, extra):
    rng = np.random.default_rng(12)
    d1 = fastq_text(rng, 2500, b"a", odd=flavour == "odd")
    d2 = fastq_text(rng, 2500, b"b", odd=flavour == "odd")
    p1, p2 = tmp_path / "r1.fq", tmp_path / "r2.fq"
    p1.write_bytes(d1)
    p2.write_bytes(d2)
    # gz: also with both files through the parallel gzip reader (host/pgzip.cpp), in pieces of 8 KiB
    par_gz = [{"MERKURIO_GZIP_THREADS": "3", "MERKURIO_GZIP_PIECE_KB": "8", "MERKURIO_BATCH_BYTES": "20000", "MERKURIO_CHUNK_BYTES": "5000"}] if flavour == "gz" else []
    for paired in (False, True):
        results = []
        for env in [{"MERKURIO_NO_FASTQ_PIPELINE": "1"}] + SIZES + par_gz:
            d = tmp_path / ("out%d%d" % (paired, len(results)))
            d.mkdir()
            args = ["extract", "-i", p1, "-s", QUERY[:25], "-o", d / "x.fastq"] + [a.replace("@", str(d)) for a in extra] + (["-2", p2] if paired else [])
            r = run(exe, stub, args, dict(env, MK_STUB_FLAG_FIRST="G"))
            assert r.returncode == 0, r.stderr
            files = {}
            for f in sorted(d.iterdir()):
                data = f.read_bytes()
                if f.suffix == ".log":
                    data = log_body(f)
                elif f.suffix == ".json":
                    import json
                    j = json.loads(data)
                    j["meta_information"] = {k: v for k, v in j["meta_information"].items() if k not in ("timestamp", "command_line")}
                    data = json.dumps(j, sort_keys=True).encode()
                files[f.name] = data
            results.append(files)
        written = sum(v.count(b"\n@a") for k, v in results[0].items() if k.endswith(".fq"))
        assert 200 < written < 2300  # some records, not all of them
        for other in results[1:]:
            assert other == results[0]


@pytest.mark.parametrize("flags", [["-m"], ["-v"], [], ["-m", "-l", "@/t.log", "-j", "@/t.json"]], ids=lambda e: "_".join(e).replace("@/", "") or "keep_all")
@pytest.mark.parametrize("bam", [False, True])
def test_tag_paths_agree_on_flagged_records(exe, stub, tmp_path, flags, bam):
    from tests.test_cli_cpu import _sam_text
    src = tmp_path / "in.sam"
    src.write_bytes(_sam_text(3000, seed=6))
    if bam:
        r = run(exe, stub, ["tag", "-i", src, "-s", QUERY[:25], "-o", tmp_path / "in.bam"])
        assert r.returncode == 0, r.stderr
        src = tmp_path / "in.bam"
    results = []
    for env in [{"MERKURIO_NO_ALN_PIPELINE": "1"}] + SIZES:
        d = tmp_path / ("o%d" % len(results))
        d.mkdir()
        r = run(exe, stub, ["tag", "-i", src, "-s", QUERY[:25], "-o", d / "t.sam"] + [f.replace("@", str(d)) for f in flags], dict(env, MK_STUB_FLAG_FIRST="C"))
        assert r.returncode == 0, r.stderr
        files = {}
        for f in sorted(d.iterdir()):
            data = f.read_bytes()
            if f.suffix == ".log":
                data = log_body(f)
            elif f.suffix == ".sam":
                data = b"\n".join(ln for ln in data.split(b"\n") if not ln.startswith(b"@PG"))
            elif f.suffix == ".json":
                import json
                j = json.loads(data)
                j["meta_information"] = {k: v for k, v in j["meta_information"].items() if k not in ("timestamp", "command_line")}
                data = json.dumps(j, sort_keys=True).encode()
            files[f.name] = data
        results.append(files)
    kept = sum(1 for ln in results[0]["t.sam"].split(b"\n") if ln and not ln.startswith(b"@"))
    # (a record that came with a km tag keeps the old field; the new one, last on the line, merges both lists)
    tagged = sum(1 for ln in results[0]["t.sam"].split(b"\n") if b"\tkm:Z:" in ln and QUERY[:25].encode() in ln.rsplit(b"\tkm:Z:", 1)[1])
    if flags[:1] == ["-m"]:
        assert kept == tagged and 300 < tagged < 2000
    elif flags[:1] == ["-v"]:
        assert tagged == 0 and 1000 < kept < 2700
    else:
        assert kept == 3000 and 300 < tagged < 2000
    for other in results[1:]:
        assert other == results[0]


def test_truncated_gzip_fasta_and_write_errors_end_the_run(exe, stub, tmp_path):
    """(a) a gzip FASTA cut short is an error in the FASTA pipeline and in the record path alike (zlib's
    gzread would have reported it from gzclose() only); (b) a record writer that cannot write (/dev/full) ends the run
    with "Error writing record to output file" and a non-zero status instead of a truncated file and status 0."""
    rng = np.random.default_rng(21)
    z = gzip.compress(fasta_text(rng))
    cut = tmp_path / "cut.fa.gz"
    cut.write_bytes(z[: len(z) * 2 // 3])
    seen = []
    for env in ({}, {"MERKURIO_NO_FASTA_PIPELINE": "1"}, {"MERKURIO_BATCH_BYTES": "20000", "MERKURIO_CHUNK_BYTES": "5000"}):
        r = run(exe, stub, ["extract", "-i", cut, "-s", QUERY, "-v", "-o", tmp_path / "o.fa"], env)
        assert r.returncode != 0 and b"decompress" in r.stderr, (env, r.stderr)
        seen.append((r.stderr, (tmp_path / "o.fa").read_bytes()))
    assert seen[0][1].count(b">rec") >= 10 and seen[1] == seen[0] and seen[2] == seen[0]  # the records in front of the cut are written
    ok = tmp_path / "ok.fa.gz"
    ok.write_bytes(z)
    assert run(exe, stub, ["extract", "-i", ok, "-s", QUERY, "-v", "-o", tmp_path / "o2.fa"]).returncode == 0
    if os.path.exists("/dev/full"):
        os.symlink("/dev/full", tmp_path / "full.fa")  # (-o keeps the name: its extension is already the input's)
        for env in ({}, {"MERKURIO_NO_FASTA_PIPELINE": "1"}):
            r = run(exe, stub, ["extract", "-i", ok, "-s", QUERY, "-v", "-o", tmp_path / "full.fa"], env)
            assert r.returncode != 0 and b"Error writing record to output file" in r.stderr, (env, r.stderr)
        sam = tmp_path / "x.sam"
        sam.write_bytes(b"@HD\tVN:1.6\n" + b"".join(b"r%d\t4\t*\t0\t0\t*\t*\t0\t0\t%s\t*\n" % (i, b"ACGT" * 30) for i in range(20000)))
        os.symlink("/dev/full", tmp_path / "full.sam")
        r = run(exe, stub, ["tag", "-i", sam, "-s", QUERY, "-o", tmp_path / "full.sam"])
        assert r.returncode != 0 and b"Error writing record to output file" in r.stderr, r.stderr


def test_input_that_can_be_read_only_once(exe, stub, tmp_path):
    """a FIFO (process substitution, /dev/stdin) is read exactly once — plain and gzip FASTQ through a
    named pipe give the output of the same data in a file."""
    import threading
    rng = np.random.default_rng(33)
    data = fastq_text(rng, 3000, b"a")
    src = tmp_path / "r.fastq"
    src.write_bytes(data)
    want_o = tmp_path / "want.fastq"
    assert run(exe, stub, ["extract", "-i", src, "-s", QUERY, "-v", "-o", want_o]).returncode == 0
    for payload, name in ((data, "p.fastq"), (gzip.compress(data), "z.fastq")):
        fifo = tmp_path / name
        os.mkfifo(fifo)
        t = threading.Thread(target=lambda: open(fifo, "wb").write(payload))
        t.start()
        got_o = tmp_path / ("got_" + name)
        r = run(exe, stub, ["extract", "-i", fifo, "-s", QUERY, "-v", "-o", got_o])
        t.join()
        assert r.returncode == 0, r.stderr
        assert got_o.read_bytes() == want_o.read_bytes(), name
