"""CPU tests: the oracle (oracle/mk_oracle.c + oracle/refmodel.py) against every golden vector and
known-answer test the reference holds for the matching path."""
import json
import random
from pathlib import Path

import pytest

from oracle import refmodel as rm
from tests.golden_util import (assert_json_equal, assert_log_equal, assert_sam_equal, json_layout, vectors)


# ---------------------------------------------------------------- unit vectors of the reference
def test_bndmq_toy_vectors():
    # src/pattern_matching.rs:353-392
    assert rm.BNDMq(b"abc", 2).find_all(b"abcabcabc") == [0, 3, 6]
    assert rm.BNDMq(b"abc", 2).find_all(b"xabcabcabcx") == [1, 4, 7]
    assert rm.BNDMq(b"abc", 1).find_all(b"abcabcabc") == [0, 3, 6]
    assert rm.BNDMq(b"abc", 3).find_all(b"abcabcabc") == [0, 3, 6]
    assert rm.BNDMq(b"abcd", 2).find_all(b"abc") == []      # pattern longer than text
    assert rm.BNDMq(b"abc", 2).find_all(b"") == []          # empty text
    assert rm.BNDMq(b"abc", 2).find_match(b"xxabcxx") is True
    assert rm.BNDMq(b"abc", 2).find_match(b"xxabxcx") is False


def test_bndmq_errors():
    # src/pattern_matching.rs:61-78, tests :394-430
    with pytest.raises(rm.RefError, match="Invalid q-gram length: 0"):
        rm.BNDMq(b"abc", 0)
    with pytest.raises(rm.RefError, match="Invalid q-gram length: 4"):
        rm.BNDMq(b"abc", 4)
    with pytest.raises(rm.RefError, match="Pattern is empty"):
        rm.BNDMq(b"", 1)
    with pytest.raises(rm.RefError, match="too large"):
        rm.BNDMq(b"A" * 65, 4)


def test_tune_q_value():
    # src/pattern_matching.rs:213-225, test :467-482
    assert rm.tune_q_value("A" * 31) == 5
    assert [rm.tune_q_value("A" * n) for n in (1, 2, 3, 4, 8, 9, 30, 31, 55, 56, 64)] == [1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6]
    with pytest.raises(rm.RefError):
        rm.tune_q_value("A" * 65)


def test_generate_masks():
    # src/pattern_preprocessing.rs:54-84
    masks, accept = rm.generate_masks(b"abc")
    assert (masks[ord("a")], masks[ord("b")], masks[ord("c")], accept) == (0b100, 0b010, 0b001, 0b100)
    assert sum(1 for m in masks if m) == 3
    masks, accept = rm.generate_masks(b"3$$X3")
    assert (masks[ord("3")], masks[ord("$")], masks[ord("X")], accept) == (0b10001, 0b01100, 0b00010, 0b10000)
    with pytest.raises(rm.RefError):
        rm.generate_masks(b"A" * 65)


def test_bndmq_equals_naive_random():
    # the BNDMq iterator reports exactly the naive occurrence set, for every q
    rng = random.Random(1234)
    for _ in range(3000):
        sigma = rng.randint(1, 4)
        m = rng.randint(1, 64)
        pat = bytes(rng.choice(b"ACGT"[:sigma]) for _ in range(m))
        n = rng.randint(0, 200)
        text = bytearray(rng.choice(b"ACGT"[:sigma]) for _ in range(n))
        if n >= m and rng.random() < 0.7:
            for _k in range(rng.randint(1, 3)):
                s = rng.randint(0, n - m)
                text[s:s + m] = pat
        q = rng.randint(1, m)
        assert rm.BNDMq(pat, q).find_all(bytes(text)) == rm.naive_find_all(pat, bytes(text))
        assert rm.BNDMq(pat, q).find_match(bytes(text)) == bool(rm.naive_find_all(pat, bytes(text)))


def test_ac_equals_naive_random():
    rng = random.Random(99)
    for _ in range(300):
        sigma = rng.randint(1, 4)
        pats = sorted({bytes(rng.choice(b"ACGT"[:sigma]) for _ in range(rng.randint(1, 8))) for _ in range(rng.randint(1, 20))})
        text = bytes(rng.choice(b"ACGTN"[:sigma + 1]) for _ in range(rng.randint(0, 120)))
        got = rm.AhoCorasick(pats).find_overlapping_iter(text)
        want = sorted(((s + len(p), s, i) for i, p in enumerate(pats) for s in rm.naive_find_all(p, text)))
        assert [(s + len(pats[i]), s, i) for i, s in got] == want  # order: end asc, start asc


def test_ac_case_insensitive():
    ac = rm.AhoCorasick([b"ACg", b"acG", b"cg"], True)
    assert ac.find_overlapping_iter(b"xAcGx") == [(0, 1), (1, 1), (2, 2)]


# ---------------------------------------------------------------- src/helpers.rs tests
def test_read_kmers_variants(ref_tree):
    d = ref_tree / "tests" / "data"
    assert sorted(rm.read_kmers_from_file(d / "kmers.txt")) == sorted(rm.read_kmers_from_file(d / "kmers.fasta"))
    assert len(rm.read_kmers_from_file(d / "kmers.txt")) == 3
    messy = rm.read_kmers_from_file(d / "kmers-messy.txt")
    assert messy[0] == "AAAAAAAAAAAAAAAAAAAAAAAAAAAA" and messy[1] == "TTGCATGAATATTGTA"
    with pytest.raises(rm.RefError, match="No k-mers found in the file"):
        rm.read_kmers_from_file(d / "kmers-empty.txt")
    with pytest.raises(rm.RefError, match="File not found"):
        rm.read_kmers_from_file(d / "does-not-exist.txt")


def test_parse_pattern_list_rules(ref_tree):
    d = ref_tree / "tests" / "data"
    # src/helpers.rs:300-331: duplicates + -r -> 4 unique, sorted
    pats = rm.parse_pattern_list(d / "kmers-duplicates.txt", None, True, False, False, False)
    assert len(pats) == 4 and pats == sorted(pats)
    # src/helpers.rs:363-397: reverse complement and canonical of the three 32-mers
    fwd = ["AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA", "CTCCGAAGAAGTTGCTGTTCTTGATGGTTATT", "TTGCATGAATATTGTAACCACATATTACCTGT"]
    rcs = ["TTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTT", "AATAACCATCAAGAACAGCAACTTCTTCGGAG", "ACAGGTAATATGTGGTTACAATATTCATGCAA"]
    assert [rm.reverse_complement(s.encode()).decode() for s in fwd] == rcs
    assert rm.parse_pattern_list(None, fwd, True, False, False, False) == sorted(fwd + rcs)
    assert rm.parse_pattern_list(None, fwd, False, True, False, False) == sorted(min(a, b) for a, b in zip(fwd, rcs))
    assert rm.parse_pattern_list(None, ["AcGt"], False, False, True, False) == ["acgt"]
    assert rm.parse_pattern_list(None, ["AcGt"], False, False, False, True) == ["ACGT"]
    # -f wins over -s; amino acids pass through; empty -> error
    assert rm.parse_pattern_list(d / "kmers-aa.txt", ["XXX"], False, False, False, False) == sorted(rm.read_kmers_from_file(d / "kmers-aa.txt"))
    with pytest.raises(rm.RefError, match="No k-mers found in file or provided sequence"):
        rm.parse_pattern_list(None, [""], False, False, False, False)
    # IUPAC complement pairs; other bytes unchanged
    assert rm.reverse_complement(b"ARYKMBVDHSWN*x") == b"x*NWSDHBVKMRYT"


def test_recommend_aho_corasick(ref_tree):
    # src/helpers.rs:555-567
    assert rm.recommend_aho_corasick(["ACGT"] * 13) is False
    assert rm.recommend_aho_corasick(["ACGT"] * 14) is True
    assert rm.recommend_aho_corasick(["A" * 65]) is True
    assert rm.recommend_aho_corasick(["A" * 64]) is False
    many = rm.read_kmers_from_file(ref_tree / "tests" / "data" / "kmers-many-40.txt")
    assert rm.recommend_aho_corasick(many) is True


def test_path_helpers():
    # src/helpers.rs:218-280
    assert str(rm.add_suffix_to_file_prefix("dir/sample.fasta.gz", "_1")) == "dir/sample_1.fasta.gz"
    assert rm.identify_uncompressed_type("x/sample.fasta.gz") == "fasta"
    assert rm.identify_uncompressed_type("x/sample.fq.bz2") == "fq"
    assert rm.identify_uncompressed_type("x/sample.fasta.xz") == "fasta"
    assert rm.identify_uncompressed_type("x/sample.fastq") == "fastq"
    with pytest.raises(rm.RefError):
        rm.identify_uncompressed_type("x/sample")


def test_log_flag_conflicts():
    # src/helpers.rs:434-552
    rm.check_log_flag_conflict(None, None, None, False)
    rm.check_log_flag_conflict("STDOUT", None, "out.fa", False)
    rm.check_log_flag_conflict("STDOUT", None, None, True)
    rm.check_log_flag_conflict("a.log", "STDOUT", "o", False)
    with pytest.raises(rm.RefError, match="both to stdout"):
        rm.check_log_flag_conflict("STDOUT", "STDOUT", "o", False)
    with pytest.raises(rm.RefError, match="Cannot write log to stdout when normal output is also stdout"):
        rm.check_log_flag_conflict("STDOUT", None, None, False)
    with pytest.raises(rm.RefError, match="Cannot write log to stdout"):
        rm.check_log_flag_conflict(None, "STDOUT", None, False)


# ---------------------------------------------------------------- end-to-end goldens: extract
def _extract(ref_tree, tmp_path, **kw):
    args = rm.CmdExtract(out_fastx=str(tmp_path / "out"), out_log=str(tmp_path / "out.log"), json_log=str(tmp_path / "out.json"), **kw)
    res = rm.extract_records(args)
    return res, args


def test_extract_simple(ref_tree, tmp_path):
    fx = ref_tree / "tests" / "fixtures"
    res, _ = _extract(ref_tree, tmp_path, in_fastx=str(fx / "input" / "simple.fasta"), kmer_seq=["ACG"], reverse_complement=True)
    assert res.search_algorithm == "BNDMq"
    assert (tmp_path / "out.fasta").read_bytes() == (fx / "extract" / "simple.extracted.fasta").read_bytes()
    assert_log_equal((tmp_path / "out.log").read_bytes(), (fx / "extract" / "simple.log").read_bytes())
    assert_json_equal((tmp_path / "out.json").read_bytes(), (fx / "extract" / "simple.json").read_bytes())
    assert json_layout((tmp_path / "out.json").read_bytes()) == json_layout((fx / "extract" / "simple.json").read_bytes())


def test_extract_simple_inverted(ref_tree, tmp_path):
    fx = ref_tree / "tests" / "fixtures"
    _extract(ref_tree, tmp_path, in_fastx=str(fx / "input" / "simple.fasta"), kmer_seq=["ACG"], reverse_complement=True, invert_match=True)
    assert (tmp_path / "out.fasta").read_bytes() == (fx / "extract" / "simple-inv.extracted.fasta").read_bytes()
    assert_log_equal((tmp_path / "out.log").read_bytes(), (fx / "extract" / "simple-inv.log").read_bytes())
    assert_json_equal((tmp_path / "out.json").read_bytes(), (fx / "extract" / "simple-inv.json").read_bytes())


def test_extract_fixed_width_aa(ref_tree, tmp_path):
    fx = ref_tree / "tests" / "fixtures"
    _extract(ref_tree, tmp_path, in_fastx=str(fx / "input" / "fixed-width.faa"), kmer_seq=["DKAT"])
    assert (tmp_path / "out.faa").read_bytes() == (fx / "extract" / "fixed-width.extracted.faa").read_bytes()
    assert_log_equal((tmp_path / "out.log").read_bytes(), (fx / "extract" / "fixed-width.log").read_bytes())
    assert_json_equal((tmp_path / "out.json").read_bytes(), (fx / "extract" / "fixed-width.json").read_bytes())


def test_extract_paired(ref_tree, tmp_path):
    fx = ref_tree / "tests" / "fixtures"
    _extract(ref_tree, tmp_path, in_fastx=str(fx / "input" / "paired-1.fastq"), in_fastq_2=str(fx / "input" / "paired-2.fastq"), kmer_seq=["CTT"])
    assert (tmp_path / "out_1.fastq").read_bytes() == (fx / "extract" / "paired_1.extracted.fastq").read_bytes()
    assert (tmp_path / "out_2.fastq").read_bytes() == (fx / "extract" / "paired_2.extracted.fastq").read_bytes()
    assert_log_equal((tmp_path / "out.log").read_bytes(), (fx / "extract" / "paired.log").read_bytes())
    assert_json_equal((tmp_path / "out.json").read_bytes(), (fx / "extract" / "paired.json").read_bytes())


def test_extract_example_workflow(ref_tree, tmp_path):
    # example-workflow/README.md:98 — 12 480 pairs x 150 bp, 3 queries + rc, BNDMq mode
    ew = ref_tree / "example-workflow"
    args = rm.CmdExtract(in_fastx=str(ew / "data" / "mutant_R1.fastq"), in_fastq_2=str(ew / "data" / "mutant_R2.fastq"),
                         kmer_file=str(ew / "data" / "significant_kmers.txt"), reverse_complement=True,
                         out_fastx=str(tmp_path / "mutant_extracted"), out_log=str(tmp_path / "x.log"), json_log=str(tmp_path / "x.json"))
    rm.extract_records(args)
    assert (tmp_path / "mutant_extracted_1.fastq").read_bytes() == (ew / "output" / "mutant_extracted_1.fastq").read_bytes()
    assert (tmp_path / "mutant_extracted_2.fastq").read_bytes() == (ew / "output" / "mutant_extracted_2.fastq").read_bytes()
    got = json.loads((tmp_path / "x.json").read_bytes())
    want = json.loads((ew / "logs" / "mutant_extracted.stats.json").read_bytes())
    for k in ("matching_records", "pattern_hit_counts", "summary_statistics", "paired_end_reads_statistics"):
        assert got[k] == want[k], k
    assert got["meta_information"]["search_algorithm"] == want["meta_information"]["search_algorithm"] == "BNDMq"


def test_extract_cfg1_known_answer(ref_tree, tmp_path):
    # BASELINE config 1 (example-minimal). The literal command logs to stdout while records also go
    # to stdout: the reference refuses it (src/helpers.rs:195-197).
    em = ref_tree / "example-minimal"
    with pytest.raises(rm.RefError, match="Cannot write log to stdout when normal output is also stdout"):
        rm.extract_records(rm.CmdExtract(in_fastx=str(em / "sample.fasta"), kmer_file=str(em / "kmers.txt"), reverse_complement=True, out_log="STDOUT"))
    args = rm.CmdExtract(in_fastx=str(em / "sample.fasta"), kmer_file=str(em / "kmers.txt"), reverse_complement=True,
                         out_fastx=str(tmp_path / "o"), out_log=str(tmp_path / "o.log"))
    res = rm.extract_records(args)
    assert res.patterns == ["AAC", "GTT"] and res.search_algorithm == "BNDMq"
    lines = (tmp_path / "o.log").read_text().split("\n")
    hits = [ln.split("\t") for ln in lines if ln and not ln.startswith("#")]
    assert len(hits) == 74
    assert "#AAC\t2" in lines and "#GTT\t2" in lines
    assert "#Total number of characters searched: 1795" in lines
    first = [(h[2], int(h[3])) for h in hits[:3]]
    assert first == [("AAC", 48), ("AAC", 54), ("AAC", 321)]
    # both records are extracted; needletail's writer terminates the last record with a line break
    src = (em / "sample.fasta").read_bytes()
    assert (tmp_path / "o.fasta").read_bytes() == (src if src.endswith(b"\n") else src + b"\n")


# ---------------------------------------------------------------- end-to-end goldens: tag
def _tag(tmp_path, **kw):
    args = rm.CmdTag(out_file=str(tmp_path / "out.sam"), out_log=str(tmp_path / "out.log"), json_log=str(tmp_path / "out.json"),
                     kmer_seq=["CTC"], reverse_complement=True, threads=2, **kw)
    return rm.tag_records(args)


@pytest.mark.parametrize("inp,kw,stem,sam", [
    ("simple.sam", dict(filter_matching=True), "simple", "simple.extracted.sam"),
    ("simple.sam", dict(invert_match=True), "simple-inv", "simple-inv.extracted.sam"),
    ("simple.bam", dict(), "simple-bam", "simple.tagged.extracted.sam"),
])
def test_tag_fixtures(ref_tree, tmp_path, inp, kw, stem, sam):
    fx = ref_tree / "tests" / "fixtures"
    res = _tag(tmp_path, in_file=str(fx / "input" / inp), **kw)
    assert res.search_algorithm == "BNDMq"
    assert_sam_equal((tmp_path / "out.sam").read_bytes(), (fx / "tag" / sam).read_bytes())
    assert_log_equal((tmp_path / "out.log").read_bytes(), (fx / "tag" / f"{stem}.log").read_bytes())
    assert_json_equal((tmp_path / "out.json").read_bytes(), (fx / "tag" / f"{stem}.json").read_bytes())
    assert json_layout((tmp_path / "out.json").read_bytes()) == json_layout((fx / "tag" / f"{stem}.json").read_bytes())


def test_tag_aho_corasick_golden(ref_tree, tmp_path):
    # tests/fixtures/extract/log.json: the only reference vector that went through Aho-Corasick
    # (10 queries + rc = 14 patterns): pins the overlapping report order and per-hit counts.
    fx = ref_tree / "tests" / "fixtures"
    args = rm.CmdTag(in_file=str(fx / "input" / "simple.bam"), suppress_output=True, reverse_complement=True,
                     kmer_seq=["CTC", "AC", "CT", "AA", "T", "A", "C", "G", "GA", "AG"], json_log=str(tmp_path / "log.json"))
    res = rm.tag_records(args)
    assert res.search_algorithm == "Aho-Corasick" and len(res.patterns) == 14
    got = json.loads((tmp_path / "log.json").read_bytes())
    want = json.loads((fx / "extract" / "log.json").read_bytes())
    assert got["matching_records"] == want["matching_records"]
    assert len(got["matching_records"]) == 96
    assert got["pattern_hit_counts"] == want["pattern_hit_counts"]
    assert got["summary_statistics"] == want["summary_statistics"]
    assert got["meta_information"]["search_algorithm"] == "Aho-Corasick"


def test_tag_example_workflow(ref_tree, tmp_path):
    # example-workflow/README.md:259 — tag without logging (BNDMq find_match per pattern)
    ew = ref_tree / "example-workflow"
    args = rm.CmdTag(in_file=str(ew / "output" / "mutant_extracted.sorted.sam"), out_file=str(tmp_path / "t.sam"),
                     kmer_file=str(ew / "data" / "significant_kmers.txt"), reverse_complement=True)
    rm.tag_records(args)
    assert_sam_equal((tmp_path / "t.sam").read_bytes(), (ew / "output" / "mutant_extracted.sorted.tagged.sam").read_bytes())
