"""GPU end-to-end tests of the C++ host binary (merkurio_b200/lib/merkurio): the reference's own
end-to-end tests (src/cmd_extract.rs:885-1056, src/cmd_tag.rs:1009-1132) replayed through the CLI
against the bundled goldens, plus randomised runs compared file-by-file with the oracle for the
paths no reference fixture covers (Aho-Corasick mode, -I, -c, -v, gzip input, long records)."""
import gzip
import json
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import refmodel as rm
from tests.golden_util import assert_json_equal, assert_log_equal, assert_sam_equal, json_layout

ROOT = Path(__file__).resolve().parent.parent
EXE = str(ROOT / "merkurio_b200" / "lib" / "merkurio")


def run(*args, env=None, check=True):
    e = dict(os.environ)
    if env:
        e.update(env)
    r = subprocess.run([EXE, *map(str, args)], capture_output=True, env=e)
    if check:
        assert r.returncode == 0, r.stderr.decode()
    return r


# ------------------------------------------------------------------ reference fixtures: extract
def test_extract_simple(ref_tree, tmp_path):
    fx = ref_tree / "tests" / "fixtures"
    run("extract", "-i", fx / "input" / "simple.fasta", "-r", "-s", "ACG", "-o", tmp_path / "out.fasta", "-l", tmp_path / "out.log", "-j", tmp_path / "out.json")
    assert (tmp_path / "out.fasta").read_bytes() == (fx / "extract" / "simple.extracted.fasta").read_bytes()
    assert_log_equal((tmp_path / "out.log").read_bytes(), (fx / "extract" / "simple.log").read_bytes())
    assert_json_equal((tmp_path / "out.json").read_bytes(), (fx / "extract" / "simple.json").read_bytes())
    assert json_layout((tmp_path / "out.json").read_bytes()) == json_layout((fx / "extract" / "simple.json").read_bytes())


def test_extract_simple_inverted(ref_tree, tmp_path):
    fx = ref_tree / "tests" / "fixtures"
    run("extract", "-i", fx / "input" / "simple.fasta", "-r", "-s", "ACG", "-v", "-o", tmp_path / "out.fasta", "-l", tmp_path / "out.log", "-j", tmp_path / "out.json")
    assert (tmp_path / "out.fasta").read_bytes() == (fx / "extract" / "simple-inv.extracted.fasta").read_bytes()
    assert_log_equal((tmp_path / "out.log").read_bytes(), (fx / "extract" / "simple-inv.log").read_bytes())
    assert_json_equal((tmp_path / "out.json").read_bytes(), (fx / "extract" / "simple-inv.json").read_bytes())


def test_extract_fixed_width_aa(ref_tree, tmp_path):
    fx = ref_tree / "tests" / "fixtures"
    run("extract", "-i", fx / "input" / "fixed-width.faa", "-s", "DKAT", "-o", tmp_path / "out.faa", "-l", tmp_path / "out.log", "-j", tmp_path / "out.json")
    assert (tmp_path / "out.faa").read_bytes() == (fx / "extract" / "fixed-width.extracted.faa").read_bytes()
    assert_log_equal((tmp_path / "out.log").read_bytes(), (fx / "extract" / "fixed-width.log").read_bytes())
    assert_json_equal((tmp_path / "out.json").read_bytes(), (fx / "extract" / "fixed-width.json").read_bytes())


def test_extract_paired(ref_tree, tmp_path):
    fx = ref_tree / "tests" / "fixtures"
    run("extract", "-i", fx / "input" / "paired-1.fastq", "-2", fx / "input" / "paired-2.fastq", "-s", "CTT",
        "-o", tmp_path / "paired.extracted.fastq", "-l", tmp_path / "out.log", "-j", tmp_path / "out.json")
    assert (tmp_path / "paired_1.extracted.fastq").read_bytes() == (fx / "extract" / "paired_1.extracted.fastq").read_bytes()
    assert (tmp_path / "paired_2.extracted.fastq").read_bytes() == (fx / "extract" / "paired_2.extracted.fastq").read_bytes()
    assert_log_equal((tmp_path / "out.log").read_bytes(), (fx / "extract" / "paired.log").read_bytes())
    assert_json_equal((tmp_path / "out.json").read_bytes(), (fx / "extract" / "paired.json").read_bytes())


def test_extract_cfg1_example_minimal(ref_tree, tmp_path):
    em = ref_tree / "example-minimal"
    run("extract", "-i", em / "sample.fasta", "-f", em / "kmers.txt", "-r", "-o", tmp_path / "o", "-l", tmp_path / "o.log")
    lines = (tmp_path / "o.log").read_text().split("\n")
    hits = [ln for ln in lines if ln and not ln.startswith("#")]
    assert len(hits) == 74 and "#AAC\t2" in lines and "#GTT\t2" in lines
    assert "#Total number of characters searched: 1795" in lines
    # and to stdout without logging (FLAG mode): both records
    r = run("extract", "-i", em / "sample.fasta", "-f", em / "kmers.txt", "-r")
    assert r.stdout == (tmp_path / "o.fasta").read_bytes()


def test_extract_example_workflow(ref_tree, tmp_path):
    ew = ref_tree / "example-workflow"
    run("extract", "-i", ew / "data" / "mutant_R1.fastq", "-2", ew / "data" / "mutant_R2.fastq", "-f", ew / "data" / "significant_kmers.txt",
        "-r", "-o", tmp_path / "mutant_extracted", "-l", tmp_path / "x.log", "-j", tmp_path / "x.json")
    assert (tmp_path / "mutant_extracted_1.fastq").read_bytes() == (ew / "output" / "mutant_extracted_1.fastq").read_bytes()
    assert (tmp_path / "mutant_extracted_2.fastq").read_bytes() == (ew / "output" / "mutant_extracted_2.fastq").read_bytes()
    got = json.loads((tmp_path / "x.json").read_bytes())
    want = json.loads((ew / "logs" / "mutant_extracted.stats.json").read_bytes())
    for k in ("matching_records", "pattern_hit_counts", "summary_statistics", "paired_end_reads_statistics"):
        assert got[k] == want[k], k
    # the same extraction without logs (FLAG mode, early-exit semantics)
    run("extract", "-i", ew / "data" / "mutant_R1.fastq", "-2", ew / "data" / "mutant_R2.fastq", "-f", ew / "data" / "significant_kmers.txt",
        "-r", "-o", tmp_path / "nolog")
    assert (tmp_path / "nolog_1.fastq").read_bytes() == (ew / "output" / "mutant_extracted_1.fastq").read_bytes()
    assert (tmp_path / "nolog_2.fastq").read_bytes() == (ew / "output" / "mutant_extracted_2.fastq").read_bytes()


# ------------------------------------------------------------------ reference fixtures: tag
@pytest.mark.parametrize("inp,flags,stem,sam", [
    ("simple.sam", ["-m"], "simple", "simple.extracted.sam"),
    ("simple.sam", ["-v"], "simple-inv", "simple-inv.extracted.sam"),
    ("simple.bam", [], "simple-bam", "simple.tagged.extracted.sam"),
])
def test_tag_fixtures(ref_tree, tmp_path, inp, flags, stem, sam):
    fx = ref_tree / "tests" / "fixtures"
    run("tag", "-i", fx / "input" / inp, "-o", tmp_path / "out.sam", "-s", "CTC", "-r", "-l", tmp_path / "out.log", "-j", tmp_path / "out.json", "-p", "2", *flags)
    assert_sam_equal((tmp_path / "out.sam").read_bytes(), (fx / "tag" / sam).read_bytes())
    assert_log_equal((tmp_path / "out.log").read_bytes(), (fx / "tag" / f"{stem}.log").read_bytes())
    assert_json_equal((tmp_path / "out.json").read_bytes(), (fx / "tag" / f"{stem}.json").read_bytes())
    assert json_layout((tmp_path / "out.json").read_bytes()) == json_layout((fx / "tag" / f"{stem}.json").read_bytes())


def test_tag_aho_corasick_golden(ref_tree, tmp_path):
    fx = ref_tree / "tests" / "fixtures"
    run("tag", "-i", fx / "input" / "simple.bam", "-S", "-s", "CTC", "AC", "CT", "AA", "T", "A", "C", "G", "GA", "AG", "-r", "-j", tmp_path / "log.json")
    got = json.loads((tmp_path / "log.json").read_bytes())
    want = json.loads((fx / "extract" / "log.json").read_bytes())
    assert got["matching_records"] == want["matching_records"] and len(got["matching_records"]) == 96
    assert got["pattern_hit_counts"] == want["pattern_hit_counts"]
    assert got["summary_statistics"] == want["summary_statistics"]
    assert got["meta_information"]["search_algorithm"] == "Aho-Corasick"


def test_tag_example_workflow(ref_tree, tmp_path):
    ew = ref_tree / "example-workflow"
    run("tag", "-i", ew / "output" / "mutant_extracted.sorted.sam", "-o", tmp_path / "t.sam", "-f", ew / "data" / "significant_kmers.txt", "-r")
    assert_sam_equal((tmp_path / "t.sam").read_bytes(), (ew / "output" / "mutant_extracted.sorted.tagged.sam").read_bytes())


# ------------------------------------------------------------------ randomised, against the oracle
def _rand_reads(rng, n, lo, hi, pats, plant=0.3, alphabet=b"ACGT"):
    al = np.frombuffer(alphabet, dtype=np.uint8)
    out = []
    for _ in range(n):
        L = int(rng.integers(lo, hi + 1))
        r = bytearray(rng.choice(al, size=L).tobytes())
        if L and rng.random() < plant:
            p = pats[int(rng.integers(len(pats)))]
            if len(p) <= L:
                s = int(rng.integers(0, L - len(p) + 1))
                r[s:s + len(p)] = p
        if L > 10 and rng.random() < 0.1:
            s = int(rng.integers(0, L - 3))
            r[s:s + 3] = b"NNN"
        out.append(bytes(r))
    return out


def _write_fastq(path, reads, prefix):
    with open(path, "wb") as f:
        for i, r in enumerate(reads):
            f.write(b"@%s%d some description\n%s\n+\n%s\n" % (prefix, i, r, b"I" * len(r)))


def _write_fasta(path, reads, width=60):
    with open(path, "wb") as f:
        for i, r in enumerate(reads):
            f.write(b">chr%d test\n" % i)
            for s in range(0, len(r), width):
                f.write(r[s:s + width] + b"\n")


def _compare_with_oracle(tmp_path, cli_args, oracle_args, outputs, env=None):
    """Run the CLI into tmp/cli and the oracle into tmp/ora; compare every output file."""
    (tmp_path / "cli").mkdir()
    (tmp_path / "ora").mkdir()
    run(*[a.replace("@OUT@", str(tmp_path / "cli")) if isinstance(a, str) else a for a in cli_args], env=env)
    oracle_args(tmp_path / "ora")
    for name, kind in outputs:
        a, b = (tmp_path / "cli" / name).read_bytes(), (tmp_path / "ora" / name).read_bytes()
        if kind == "raw":
            assert a == b, name
        elif kind == "log":
            assert_log_equal(a, b)
        elif kind == "json":
            assert_json_equal(a, b)
            assert json_layout(a) == json_layout(b)
        elif kind == "sam":
            assert_sam_equal(a, b)


@pytest.mark.parametrize("mode", ["ac", "bndmq", "insensitive", "canonical", "inverted"])
def test_extract_random_single(tmp_path, mode):
    rng = np.random.default_rng({"ac": 1, "bndmq": 2, "insensitive": 3, "canonical": 4, "inverted": 5}[mode])
    n_pat = 6 if mode == "bndmq" else 30
    pats = sorted({rng.choice(np.frombuffer(b"ACGT", np.uint8), size=int(rng.integers(12, 40))).tobytes() for _ in range(n_pat)})
    reads = _rand_reads(rng, 3000, 0, 200, pats)
    if mode == "insensitive":
        reads = [r.lower() if i % 2 else r for i, r in enumerate(reads)]
    fq = tmp_path / "reads.fastq"
    _write_fastq(fq, reads, b"r")
    kf = tmp_path / "k.txt"
    kf.write_bytes(b"# queries\n" + b"\n".join(pats) + b"\n")
    flags = {"ac": ["-r"], "bndmq": ["-r"], "insensitive": ["-I"], "canonical": ["-c"], "inverted": ["-r", "-v"]}[mode]
    kw = dict(reverse_complement=mode in ("ac", "bndmq", "inverted"), case_insensitive=mode == "insensitive", canonical=mode == "canonical", invert_match=mode == "inverted")

    def oracle(out):
        rm.extract_records(rm.CmdExtract(in_fastx=str(fq), kmer_file=str(kf), out_fastx=str(out / "o"), out_log=str(out / "o.log"), json_log=str(out / "o.json"), **kw))

    _compare_with_oracle(tmp_path, ["extract", "-i", fq, "-f", kf, *flags, "-o", "@OUT@/o", "-l", "@OUT@/o.log", "-j", "@OUT@/o.json"], oracle,
                         [("o.fastq", "raw"), ("o.log", "log"), ("o.json", "json")], env={"MERKURIO_BATCH_BYTES": "100000", "MERKURIO_SLOTS": "2"})
    # without logs: the same records
    (tmp_path / "nolog").mkdir()
    run("extract", "-i", fq, "-f", kf, *flags, "-o", tmp_path / "nolog" / "o")
    assert (tmp_path / "nolog" / "o.fastq").read_bytes() == (tmp_path / "ora" / "o.fastq").read_bytes()


@pytest.mark.parametrize("n_pat,gz_threads", [(4, 1), (40, 1), (40, 3)])
def test_extract_random_paired_gz(tmp_path, n_pat, gz_threads):
    """gz_threads = 3: both files through the parallel gzip reader (host/pgzip.cpp), in pieces of 8 KiB."""
    rng = np.random.default_rng(n_pat)
    pats = sorted({rng.choice(np.frombuffer(b"ACGT", np.uint8), size=31).tobytes() for _ in range(n_pat)})
    r1 = _rand_reads(rng, 2500, 150, 150, pats, plant=0.05)
    r2 = _rand_reads(rng, 2500, 150, 150, pats, plant=0.05)
    _write_fastq(tmp_path / "a_1.fastq", r1, b"p")
    _write_fastq(tmp_path / "a_2.fastq", r2, b"p")
    for n in ("a_1.fastq", "a_2.fastq"):
        (tmp_path / (n + ".gz")).write_bytes(gzip.compress((tmp_path / n).read_bytes()))
    kf = tmp_path / "k.txt"
    kf.write_bytes(b"\n".join(pats) + b"\n")

    def oracle(out):
        rm.extract_records(rm.CmdExtract(in_fastx=str(tmp_path / "a_1.fastq.gz"), in_fastq_2=str(tmp_path / "a_2.fastq.gz"), kmer_file=str(kf),
                                         reverse_complement=True, out_fastx=str(out / "x"), out_log=str(out / "x.log"), json_log=str(out / "x.json")))

    _compare_with_oracle(tmp_path, ["extract", "-i", tmp_path / "a_1.fastq.gz", "-2", tmp_path / "a_2.fastq.gz", "-f", kf, "-r", "-o", "@OUT@/x",
                                    "-l", "@OUT@/x.log", "-j", "@OUT@/x.json"], oracle,
                         [("x_1.fastq", "raw"), ("x_2.fastq", "raw"), ("x.log", "log"), ("x.json", "json")],
                         env={"MERKURIO_BATCH_BYTES": "200000", "MERKURIO_GZIP_THREADS": str(gz_threads), "MERKURIO_GZIP_PIECE_KB": "8"})


def test_extract_long_fasta_records_in_pieces(tmp_path):
    # chromosomes much longer than a batch: cut into overlapping pieces, hits on every piece boundary
    rng = np.random.default_rng(77)
    pats = sorted({rng.choice(np.frombuffer(b"ACGT", np.uint8), size=int(k)).tobytes() for k in rng.integers(21, 64, size=40)} | {b"ACGTNACGTNACGTNACGTNACGT"})
    chroms = []
    for n in (300000, 5, 0, 123457):
        c = bytearray(rng.choice(np.frombuffer(b"ACGTacgtN", np.uint8), size=n, p=[.22, .22, .22, .22, .02, .02, .02, .02, .04]).tobytes())
        for s in range(100, max(n - 100, 0), 997):
            p = pats[(s // 997) % len(pats)]
            c[s:s + len(p)] = p
        chroms.append(bytes(c))
    fa = tmp_path / "g.fasta"
    _write_fasta(fa, chroms)
    kf = tmp_path / "k.txt"
    kf.write_bytes(b"\n".join(pats) + b"\n")

    def oracle(out):
        rm.extract_records(rm.CmdExtract(in_fastx=str(fa), kmer_file=str(kf), suppress_output=True, json_log=str(out / "g.json"), out_log=str(out / "g.log")))

    _compare_with_oracle(tmp_path, ["extract", "-i", fa, "-f", kf, "-S", "-j", "@OUT@/g.json", "-l", "@OUT@/g.log"], oracle,
                         [("g.log", "log"), ("g.json", "json")], env={"MERKURIO_BATCH_BYTES": "20000", "MERKURIO_SLOTS": "2"})


@pytest.mark.parametrize("flags,kw", [(["-m"], dict(filter_matching=True)), (["-v"], dict(invert_match=True)), ([], {})])
def test_tag_random_sam(tmp_path, flags, kw):
    rng = np.random.default_rng(len(flags) + 5)
    pats = sorted({rng.choice(np.frombuffer(b"ACGT", np.uint8), size=31).tobytes() for _ in range(25)})
    reads = _rand_reads(rng, 2000, 0, 151, pats, plant=0.2)
    sam = tmp_path / "in.sam"
    with open(sam, "wb") as f:
        f.write(b"@HD\tVN:1.6\tSO:unsorted\n@SQ\tSN:1\tLN:100000\n")
        for i, r in enumerate(reads):
            seq = r if r else b"*"
            extra = b"\tkm:Z:ZZZ,AAA" if i % 50 == 0 else b""
            f.write(b"q%d\t0\t1\t%d\t60\t%dM\t*\t0\t0\t%s\t%s\tNM:i:0%s\n" % (i, i + 1, max(len(r), 1), seq, b"F" * len(r) if r else b"*", extra))
    kf = tmp_path / "k.txt"
    kf.write_bytes(b"\n".join(pats) + b"\n")

    def oracle(out):
        rm.tag_records(rm.CmdTag(in_file=str(sam), out_file=str(out / "t.sam"), kmer_file=str(kf), reverse_complement=True,
                                 out_log=str(out / "t.log"), json_log=str(out / "t.json"), **kw))

    _compare_with_oracle(tmp_path, ["tag", "-i", sam, "-o", "@OUT@/t.sam", "-f", kf, "-r", "-l", "@OUT@/t.log", "-j", "@OUT@/t.json", *flags], oracle,
                         [("t.sam", "sam"), ("t.log", "log"), ("t.json", "json")], env={"MERKURIO_BATCH_BYTES": "50000"})
    # without logs (PATTERN_SET mode): same SAM
    (tmp_path / "nolog").mkdir()
    run("tag", "-i", sam, "-o", tmp_path / "nolog" / "t.sam", "-f", kf, "-r", *flags)
    assert_sam_equal((tmp_path / "nolog" / "t.sam").read_bytes(), (tmp_path / "ora" / "t.sam").read_bytes())


def test_unequal_paired_files(ref_tree, tmp_path):
    fx = ref_tree / "tests" / "fixtures" / "input"
    one = (fx / "paired-1.fastq").read_bytes()
    short = b"\n".join(one.split(b"\n")[:4]) + b"\n"
    (tmp_path / "short.fastq").write_bytes(short)
    r = run("extract", "-i", fx / "paired-1.fastq", "-2", tmp_path / "short.fastq", "-s", "CTT", "-o", tmp_path / "o", check=False)
    assert r.returncode == 1 and b"Do the two input files contain the same number of records?" in r.stderr
    r = run("extract", "-i", tmp_path / "short.fastq", "-2", fx / "paired-2.fastq", "-s", "CTT", "-o", tmp_path / "p", check=False)
    assert r.returncode == 1 and b"The two input files have a different number of records." in r.stderr


# ------------------------------------------------------------------ BAM output (untested by the reference, src/cmd_tag.rs:1134)
@pytest.mark.parametrize("inp", ["simple.sam", "simple.bam"])
def test_tag_bam_output_round_trip(ref_tree, tmp_path, inp):
    fx = ref_tree / "tests" / "fixtures" / "input"
    run("tag", "-i", fx / inp, "-o", tmp_path / "out.bam", "-s", "CTC", "-r")
    header, recs = rm.parse_bam((tmp_path / "out.bam").read_bytes())
    res = rm.tag_records(rm.CmdTag(in_file=str(fx / inp), out_file=str(tmp_path / "ora.sam"), kmer_seq=["CTC"], reverse_complement=True))
    want = [ln for ln in (tmp_path / "ora.sam").read_bytes().split(b"\n") if ln]
    got = header + [r.line for r in recs]
    assert [ln for ln in got if not ln.startswith(b"@PG")] == [ln for ln in want if not ln.startswith(b"@PG")]
    assert any(ln.startswith(b"@PG\tID:merkurio") for ln in got)
    # BGZF framing: gzip members with the BC extra field and the 28-byte EOF marker
    raw = (tmp_path / "out.bam").read_bytes()
    assert raw[:4] == b"\x1f\x8b\x08\x04" and raw[12:14] == b"BC" and raw.endswith(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))


def test_tag_bam_output_random(tmp_path):
    rng = np.random.default_rng(99)
    pats = sorted({rng.choice(np.frombuffer(b"ACGT", np.uint8), size=25).tobytes() for _ in range(20)})
    reads = _rand_reads(rng, 3000, 1, 151, pats, plant=0.3)
    sam = tmp_path / "in.sam"
    with open(sam, "wb") as f:
        f.write(b"@HD\tVN:1.6\tSO:unsorted\n@SQ\tSN:chr1\tLN:1000000\n@SQ\tSN:chr2\tLN:5000\n")
        for i, r in enumerate(reads):
            f.write(b"q%d\t%d\tchr%d\t%d\t%d\t%dM\t%s\t%d\t%d\t%s\t%s\tNM:i:%d\tXS:A:+\tZZ:Z:hello world\tXB:B:s,-3,7,300\n" % (
                i, 99 if i % 2 else 147, 1 + i % 2, 1 + (i * 37) % 4000, i % 61, len(r), b"=" if i % 3 else b"*", (i * 11) % 4000 if i % 3 else 0,
                (i % 500) - 250, r, b"F" * len(r), (i * 7919) % 100000 - 300))
    kf = tmp_path / "k.txt"
    kf.write_bytes(b"\n".join(pats) + b"\n")
    run("tag", "-i", sam, "-o", tmp_path / "o.bam", "-f", kf, "-r", "-m", env={"MERKURIO_BATCH_BYTES": "60000"})
    rm.tag_records(rm.CmdTag(in_file=str(sam), out_file=str(tmp_path / "ora.sam"), kmer_file=str(kf), reverse_complement=True, filter_matching=True))
    header, recs = rm.parse_bam((tmp_path / "o.bam").read_bytes())
    want = [ln for ln in (tmp_path / "ora.sam").read_bytes().split(b"\n") if ln and not ln.startswith(b"@")]
    assert [r.line for r in recs] == want
    # and the BAM we wrote is a valid input of its own
    run("tag", "-i", tmp_path / "o.bam", "-o", tmp_path / "again.sam", "-f", kf, "-r", "-t", "kk")
    again = [ln for ln in (tmp_path / "again.sam").read_bytes().split(b"\n") if ln and not ln.startswith(b"@")]
    assert len(again) == len(want) and all(a.startswith(w) for a, w in zip(again, want))


@pytest.mark.parametrize("n_engines", [2, 3])
def test_several_engines_from_the_cli(ref_tree, tmp_path, n_engines):
    """The host's multi-GPU path: one engine per device, batches dealt round-robin, results merged in batch order, the
    seed tables built once and uploaded per engine (mk_tables_create / mk_engine_create_shared). On a box with fewer
    GPUs than engines the engines share device 0 (MERKURIO_DEVICES=0,0,...): the deal and the merge are the same."""
    import torch
    multi = ({"MERKURIO_GPUS": str(n_engines)} if torch.cuda.device_count() >= n_engines
             else {"MERKURIO_DEVICES": ",".join(["0"] * n_engines)})
    ew = ref_tree / "example-workflow"
    run("extract", "-i", ew / "data" / "mutant_R1.fastq", "-2", ew / "data" / "mutant_R2.fastq", "-f", ew / "data" / "significant_kmers.txt",
        "-r", "-o", tmp_path / "m", "-j", tmp_path / "m.json", env=dict(multi, MERKURIO_BATCH_BYTES="300000", MERKURIO_TIMING="1"))
    assert (tmp_path / "m_1.fastq").read_bytes() == (ew / "output" / "mutant_extracted_1.fastq").read_bytes()
    assert (tmp_path / "m_2.fastq").read_bytes() == (ew / "output" / "mutant_extracted_2.fastq").read_bytes()
    got = json.loads((tmp_path / "m.json").read_bytes())
    want = json.loads((ew / "logs" / "mutant_extracted.stats.json").read_bytes())
    assert got["matching_records"] == want["matching_records"] and got["summary_statistics"] == want["summary_statistics"]
    # FASTA pieces and SAM records dealt over two engines: same files as on one
    rng = np.random.default_rng(5)
    pats = sorted({rng.choice(np.frombuffer(b"ACGT", np.uint8), size=31).tobytes() for _ in range(20)})
    recs = _rand_reads(rng, 6, 30000, 60000, pats, plant=1.0)
    fa = tmp_path / "g.fa"
    _write_fasta(fa, recs)
    kf = tmp_path / "k.txt"
    kf.write_bytes(b"\n".join(pats) + b"\n")
    one = _outputs(tmp_path, "fa1", ["extract", "-i", fa, "-f", kf, "-r", "-o", "@OUT@/x.fa", "-l", "@OUT@/x.log"], {"MERKURIO_BATCH_BYTES": "25000"})
    two = _outputs(tmp_path, "fa2", ["extract", "-i", fa, "-f", kf, "-r", "-o", "@OUT@/x.fa", "-l", "@OUT@/x.log"],
                   dict(multi, MERKURIO_BATCH_BYTES="25000"))
    assert one[0] == two[0] == 0 and one[2]["x.fa"] == two[2]["x.fa"] and one[2]["x.fa"].count(b">chr") >= 1
    assert _strip_volatile("x.log", one[2]["x.log"]) == _strip_volatile("x.log", two[2]["x.log"])
    sam = ref_tree / "example-workflow" / "output" / "mutant_extracted.sorted.sam"
    a = _outputs(tmp_path, "t1", ["tag", "-i", sam, "-f", ew / "data" / "significant_kmers.txt", "-r", "-o", "@OUT@/t.sam"], {"MERKURIO_BATCH_BYTES": "20000"})
    b = _outputs(tmp_path, "t2", ["tag", "-i", sam, "-f", ew / "data" / "significant_kmers.txt", "-r", "-o", "@OUT@/t.sam"],
                 dict(multi, MERKURIO_BATCH_BYTES="20000"))
    assert a[0] == b[0] == 0 and _strip_volatile("t.sam", a[2]["t.sam"]) == _strip_volatile("t.sam", b[2]["t.sam"])


# ------------------------------------------------------------------ ingest pipelines vs the line-by-line readers
def _outputs(tmp_path, tag, args, env):
    d = tmp_path / tag
    d.mkdir()
    r = run(*[a.replace("@OUT@", str(d)) if isinstance(a, str) else a for a in args], env=env, check=False)
    files = {p.name: p.read_bytes() for p in sorted(d.iterdir())}
    return r.returncode, r.stderr, files


def _strip_volatile(name, data):
    if name.endswith(".log"):
        return b"\n".join(data.split(b"\n")[4:])
    if name.endswith(".json"):
        try:
            d = json.loads(data)
        except json.JSONDecodeError:
            return data  # a run that ended with an error leaves the hit objects without the closing sections
        d["meta_information"] = {k: v for k, v in d["meta_information"].items() if k not in ("timestamp", "command_line")}
        return json.dumps(d, sort_keys=True).encode()
    if name.endswith(".sam"):
        return b"\n".join(ln for ln in data.split(b"\n") if not ln.startswith(b"@PG"))
    return data


def _outputs_batched(tmp_path, runs):
    """Several (tag, args, env) runs in ONE process (`merkurio batch`): one CUDA start-up for all of them."""
    lines = []
    for tag, args, env in runs:
        d = tmp_path / tag
        d.mkdir()
        fields = ["%s=%s" % kv for kv in env.items()] + [str(a).replace("@OUT@", str(d)) for a in args]
        assert not any("\t" in x or "\n" in x for x in fields)
        lines.append("\t".join(fields))
    script = tmp_path / "batch.txt"
    script.write_text("\n".join(lines) + "\n")
    r = subprocess.run([EXE, "batch", str(script)], capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    out = []
    for i, (tag, _, _) in enumerate(runs):
        rc = int((tmp_path / ("batch.txt.%d.rc" % i)).read_text())
        err = (tmp_path / ("batch.txt.%d.err" % i)).read_bytes()
        out.append((rc, err, {p.name: p.read_bytes() for p in sorted((tmp_path / tag).iterdir())}))
    return out


def _same_both_ways(tmp_path, args, off_switch, env=None):
    """The reader -> packer -> GPU pipeline and the record-by-record path must agree on every output
    file, the exit status and the error text — also when the input breaks off half way."""
    e = dict(env or {})
    a, b = _outputs_batched(tmp_path, [("pipe", args, e), ("plain", args, dict(e, **{off_switch: "1"}))])
    assert a[0] == b[0], (a[1], b[1])
    assert a[1].replace(b"[merkurio]", b"") == b[1].replace(b"[merkurio]", b"")
    assert a[2].keys() == b[2].keys()
    for name in a[2]:
        assert _strip_volatile(name, a[2][name]) == _strip_volatile(name, b[2][name]), name
    return a


@pytest.mark.parametrize("flavour,logs", [("plain", False), ("plain", True), ("crlf", True), ("odd", True), ("gz", True),
                                          ("truncated", False), ("truncated", True), ("bad_second_file", True)])
def test_fastq_pipeline_equals_record_path(tmp_path, flavour, logs):
    rng = np.random.default_rng(77)
    pats = sorted({rng.choice(np.frombuffer(b"ACGT", np.uint8), size=int(k)).tobytes() for k in rng.integers(18, 40, size=30)})
    r1 = _rand_reads(rng, 3000, 0, 160, pats, plant=0.3)
    r2 = _rand_reads(rng, 3000, 0, 160, pats, plant=0.3)

    def fastq(reads, prefix):
        out = bytearray()
        for i, r in enumerate(reads):
            le = b"\r\n" if flavour == "crlf" else b"\n"
            plus = b"+" + (b"%s%d" % (prefix, i) if flavour == "odd" and i % 3 == 0 else b"")
            out += b"@%s%d d=%d" % (prefix, i, i) + le + r + le + plus + le + b"F" * len(r) + le
            if flavour == "odd" and i % 7 == 0:
                out += b"\n"
        if flavour == "odd":
            out = out.rstrip(b"\n")  # no final newline
        return bytes(out)

    d1, d2 = fastq(r1, b"a"), fastq(r2, b"b")
    if flavour == "truncated":
        d1 = d1[: len(d1) // 2]
    if flavour == "bad_second_file":
        d2 = d2[: len(d2) // 3] + b"garbage\n" + d2[len(d2) // 3:]
    ext = ".fastq.gz" if flavour == "gz" else ".fastq"
    p1, p2 = tmp_path / ("r1" + ext), tmp_path / ("r2" + ext)
    p1.write_bytes(gzip.compress(d1) if flavour == "gz" else d1)
    p2.write_bytes(gzip.compress(d2) if flavour == "gz" else d2)
    kf = tmp_path / "k.txt"
    kf.write_bytes(b"\n".join(pats) + b"\n")
    log_args = ["-l", "@OUT@/x.log", "-j", "@OUT@/x.json"] if logs else []
    env = {"MERKURIO_BATCH_BYTES": "70000", "MERKURIO_CHUNK_BYTES": "50000"}
    # single end
    (tmp_path / "se").mkdir()
    rc, err, files = _same_both_ways(tmp_path / "se", ["extract", "-i", p1, "-f", kf, "-r", "-o", "@OUT@/x.fastq", *log_args],
                                     "MERKURIO_NO_FASTQ_PIPELINE", env)
    assert (rc != 0) == (flavour == "truncated"), err
    assert files["x.fastq"].count(b"\n@a") > 50
    # paired, also inverted
    for extra, tag in (([], "pe"), (["-v"], "pev")):
        if extra and flavour != "plain":
            continue  # every CLI run pays seconds of CUDA start-up: the inverted variant once is enough
        (tmp_path / tag).mkdir()
        rc, err, files = _same_both_ways(tmp_path / tag, ["extract", "-i", p1, "-2", p2, "-f", kf, "-r", "-o", "@OUT@/x.fastq", *extra, *log_args],
                                         "MERKURIO_NO_FASTQ_PIPELINE", env)
        assert (rc != 0) == (flavour in ("truncated", "bad_second_file")), (tag, err)
        assert set(files) >= {"x_1.fastq", "x_2.fastq"}


@pytest.mark.parametrize("kind,flags", [("sam", ["-m"]), ("sam", ["-v"]), ("sam", []), ("sam", ["-m", "-l", "@OUT@/t.log", "-j", "@OUT@/t.json"]),
                                        ("bam", ["-m"]), ("bam", []), ("bam", ["-l", "@OUT@/t.log", "-j", "@OUT@/t.json"]),
                                        ("sam_truncated", ["-m"])])
def test_aln_pipeline_equals_record_path(tmp_path, kind, flags):
    rng = np.random.default_rng(91)
    pats = sorted({rng.choice(np.frombuffer(b"ACGT", np.uint8), size=31).tobytes() for _ in range(25)})
    reads = _rand_reads(rng, 3000, 0, 151, pats, plant=0.2, alphabet=b"ACGTNacgtRY")
    sam = tmp_path / "in.sam"
    with open(sam, "wb") as f:
        f.write(b"@HD\tVN:1.6\tSO:unsorted\n@SQ\tSN:1\tLN:100000\n@SQ\tSN:2\tLN:5000\n")
        for i, r in enumerate(reads):
            extra = b"\tkm:Z:ZZZ,AAA" if i % 50 == 0 else b""
            f.write(b"q%d\t%d\t%d\t%d\t%d\t%s\t%s\t%d\t0\t%s\t%s\tNM:i:%d\tXA:Z:x,y;%s\n" % (
                i, 99 if i % 2 else 147, 1 + i % 2, 1 + (i * 37) % 4000, i % 61, b"%dM" % len(r) if r else b"*",
                b"=" if i % 3 else b"*", (i * 11) % 4000 if i % 3 else 0, r.upper() if r else b"*", b"F" * len(r) if r else b"*", i % 5, extra))
            if i % 400 == 0:
                f.write(b"\n")
    kf = tmp_path / "k.txt"
    kf.write_bytes(b"\n".join(pats) + b"\n")
    src = sam
    if kind == "bam":
        src = tmp_path / "in.bam"
        run("tag", "-i", sam, "-o", src, "-s", "NNNNNNNNNNNNNNNNNNNNNNNNNNNNNNN", env={"MERKURIO_BATCH_BYTES": "60000"})
    if kind == "sam_truncated":
        src = tmp_path / "cut.sam"
        data = sam.read_bytes()
        src.write_bytes(data[: len(data) // 2].rsplit(b"\t", 4)[0] + b"\n" + data[len(data) // 2:])
    env = {"MERKURIO_BATCH_BYTES": "50000", "MERKURIO_CHUNK_BYTES": "40000"}
    rc, err, files = _same_both_ways(tmp_path, ["tag", "-i", src, "-o", "@OUT@/t.sam", "-f", kf, "-r", *flags], "MERKURIO_NO_ALN_PIPELINE", env)
    assert (rc != 0) == (kind == "sam_truncated"), err
    if rc == 0:
        assert files["t.sam"].count(b"\tkm:Z:") > 100


@pytest.mark.parametrize("flags", [[], ["-m"]])
def test_tag_bam_to_bam_passthrough(tmp_path, flags):
    """BAM in, BAM out: kept records are copied in their binary form with the tag appended (no SAM text round
    trip), blocks compressed on several threads. Must decode to the same records as the text round trip and
    as the oracle's SAM output."""
    rng = np.random.default_rng(97)
    pats = sorted({rng.choice(np.frombuffer(b"ACGT", np.uint8), size=31).tobytes() for _ in range(25)})
    reads = _rand_reads(rng, 4000, 0, 151, pats, plant=0.2)
    sam = tmp_path / "in.sam"
    with open(sam, "wb") as f:
        f.write(b"@HD\tVN:1.6\tSO:unsorted\n@SQ\tSN:1\tLN:100000\n@SQ\tSN:2\tLN:5000\n")
        for i, r in enumerate(reads):
            extra = b"\tkm:Z:ZZZ,AAA" if i % 50 == 0 else b""
            f.write(b"q%d\t%d\t%d\t%d\t%d\t%s\t%s\t%d\t0\t%s\t%s\tNM:i:%d\tXB:B:s,1,-2,300%s\n" % (
                i, 99 if i % 2 else 147, 1 + i % 2, 1 + (i * 37) % 4000, i % 61, b"%dM" % len(r) if r else b"*",
                b"=" if i % 3 else b"*", (i * 11) % 4000 if i % 3 else 0, r if r else b"*", b"F" * len(r) if r else b"*", i % 5, extra))
    kf = tmp_path / "k.txt"
    kf.write_bytes(b"\n".join(pats) + b"\n")
    bam = tmp_path / "in.bam"
    run("tag", "-i", sam, "-o", bam, "-s", "NNNNNNNNNNNNNNNNNNNNNNNNNNNNNNN")
    # the input BAM itself carries an (empty) km tag from that run: existing tags are merged, the old field stays
    runs = [("a", ["tag", "-i", bam, "-o", "@OUT@/o.bam", "-f", kf, "-r", "-p", "4", *flags], {"MERKURIO_BATCH_BYTES": "60000"}),
            ("b", ["tag", "-i", bam, "-o", "@OUT@/o.bam", "-f", kf, "-r", *flags], {"MERKURIO_NO_BAM_PASSTHROUGH": "1"}),
            ("c", ["tag", "-i", bam, "-o", "@OUT@/o.sam", "-f", kf, "-r", *flags], {})]
    (a, b, c) = _outputs_batched(tmp_path, runs)
    assert a[0] == b[0] == c[0] == 0, (a[1], b[1], c[1])
    ha, ra = rm.parse_bam(a[2]["o.bam"])
    hb, rb = rm.parse_bam(b[2]["o.bam"])
    assert [x for x in ha if not x.startswith(b"@PG")] == [x for x in hb if not x.startswith(b"@PG")]
    assert [r.line for r in ra] == [r.line for r in rb]
    sam_lines = [ln for ln in c[2]["o.sam"].split(b"\n") if ln and not ln.startswith(b"@")]
    assert [r.line for r in ra] == sam_lines
    assert len(ra) == (4000 if not flags else sum(1 for ln in sam_lines)) and len(ra) > 300


@pytest.mark.parametrize("logs", [False, True])
def test_fasta_pipeline_equals_record_path_and_oracle(tmp_path, logs):
    """Multi-line FASTA through the reader -> packer -> GPU pipeline with batches far smaller than the records:
    every occurrence must be reported once, at its position in the record, also when it straddles two pieces."""
    rng = np.random.default_rng(123)
    pats = sorted({rng.choice(np.frombuffer(b"ACGT", np.uint8), size=int(k)).tobytes() for k in rng.integers(21, 64, size=30)})
    recs = []
    for i in range(12):
        n = int(rng.integers(0, 3000)) if i % 3 else int(rng.integers(60000, 140000))
        r = bytearray(rng.choice(np.frombuffer(b"ACGTNacgt", np.uint8), size=n).tobytes())
        pos = 50
        while pos + 70 < n:  # an occurrence every ~1 kb: some of them cross every kind of boundary
            p = pats[int(rng.integers(len(pats)))]
            r[pos:pos + len(p)] = p
            pos += int(rng.integers(700, 1400))
        recs.append(bytes(r))
    fa = tmp_path / "g.fa"
    _write_fasta(fa, recs, width=70)
    kf = tmp_path / "k.txt"
    kf.write_bytes(b"\n".join(pats) + b"\n")
    log_args = ["-l", "@OUT@/x.log", "-j", "@OUT@/x.json"] if logs else []
    env = {"MERKURIO_BATCH_BYTES": "30000", "MERKURIO_CHUNK_BYTES": "8000"}
    args = ["extract", "-i", fa, "-f", kf, "-o", "@OUT@/x.fa", *log_args]
    rc, err, files = _same_both_ways(tmp_path, args, "MERKURIO_NO_FASTA_PIPELINE", env)
    assert rc == 0, err
    # and against the oracle
    (tmp_path / "ora").mkdir()
    rm.extract_records(rm.CmdExtract(in_fastx=str(fa), kmer_file=str(kf), out_fastx=str(tmp_path / "ora" / "x.fa"),
                                     out_log=str(tmp_path / "ora" / "x.log") if logs else None, json_log=str(tmp_path / "ora" / "x.json") if logs else None))
    assert files["x.fa"] == (tmp_path / "ora" / "x.fa").read_bytes()
    assert files["x.fa"].count(b">chr") >= 4
    if logs:
        assert_log_equal(files["x.log"], (tmp_path / "ora" / "x.log").read_bytes())
        assert_json_equal(files["x.json"], (tmp_path / "ora" / "x.json").read_bytes())


@pytest.mark.parametrize("mode", ["ac_logs", "bndmq_logs", "no_logs", "unequal"])
def test_paired_fasta_pipeline_equals_record_path_and_oracle(tmp_path, mode):
    """Two FASTA files of mates (src/cmd_extract.rs:412-418, :463-607) through the chunked FASTA pipeline: batches far
    smaller than the long records, so mates end in different batches and occurrences straddle pieces in either file. Same
    pair files, logs, exit status and error text as the record-by-record path, and the oracle's outputs."""
    rng = np.random.default_rng(321)
    n_pat = 3 if mode == "bndmq_logs" else 25
    pats = sorted({rng.choice(np.frombuffer(b"ACGT", np.uint8), size=int(k)).tobytes() for k in rng.integers(21, 50, size=n_pat)})
    files_in = []
    for f in range(2):
        recs = []
        for i in range(120):
            n = int(rng.integers(0, 500)) if (i + 7 * f) % 40 else int(rng.integers(40000, 90000))
            r = bytearray(rng.choice(np.frombuffer(b"ACGTNacgt", np.uint8), size=n).tobytes())
            pos = int(rng.integers(0, 900))
            while pos + 70 < n:
                p = pats[int(rng.integers(len(pats)))]
                r[pos:pos + len(p)] = p
                pos += int(rng.integers(700, 1400))
            recs.append(bytes(r))
        if mode == "unequal" and f == 1:
            recs = recs[:90]
        fa = tmp_path / ("m_%d.fa" % (f + 1))
        _write_fasta(fa, recs, width=70)
        files_in.append(fa)
    kf = tmp_path / "k.txt"
    kf.write_bytes(b"\n".join(pats) + b"\n")
    logs = mode != "no_logs"
    log_args = ["-l", "@OUT@/x.log", "-j", "@OUT@/x.json"] if logs else []
    env = {"MERKURIO_BATCH_BYTES": "30000", "MERKURIO_CHUNK_BYTES": "8000"}
    args = ["extract", "-i", files_in[0], "-2", files_in[1], "-f", kf, "-o", "@OUT@/x.fa", *log_args]
    rc, err, files = _same_both_ways(tmp_path, args, "MERKURIO_NO_FASTA_PIPELINE", env)
    if mode == "unequal":
        assert rc != 0 and b"Do the two input files contain the same number of records?" in err
        assert files["x_1.fa"].count(b">chr") >= 10
        return
    assert rc == 0, err
    (tmp_path / "ora").mkdir()
    rm.extract_records(rm.CmdExtract(in_fastx=str(files_in[0]), in_fastq_2=str(files_in[1]), kmer_file=str(kf), out_fastx=str(tmp_path / "ora" / "x.fa"),
                                     out_log=str(tmp_path / "ora" / "x.log") if logs else None, json_log=str(tmp_path / "ora" / "x.json") if logs else None))
    for name in ("x_1.fa", "x_2.fa"):
        assert files[name] == (tmp_path / "ora" / name).read_bytes(), name
    assert files["x_1.fa"].count(b">chr") >= 20
    if logs:
        assert_log_equal(files["x.log"], (tmp_path / "ora" / "x.log").read_bytes())
        assert_json_equal(files["x.json"], (tmp_path / "ora" / "x.json").read_bytes())
