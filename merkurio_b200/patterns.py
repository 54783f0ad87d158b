"""Python mirror of the query-list preprocessing the C++ host performs (host/helpers.cpp), for the
benchmark and test plumbing. Follows the reference's src/helpers.rs:76-133 (parse_pattern_list):
optional case conversion, append reverse complements (-r) or map to canonical form (-c), drop
empties, byte-sort, dedup. Complement table as needletail 0.6.3: ACGT and the IUPAC pairs in both
cases, every other byte unchanged."""
from __future__ import annotations

from typing import List, Sequence

_PAIRS = ("AT", "CG", "RY", "KM", "BV", "DH", "SS", "WW")
_T = list(range(256))
for _a, _b in _PAIRS:
    for x, y in ((_a, _b), (_b, _a), (_a.lower(), _b.lower()), (_b.lower(), _a.lower())):
        _T[ord(x)] = ord(y)
_COMPLEMENT = bytes(_T)


def reverse_complement(seq: bytes) -> bytes:
    return seq.translate(_COMPLEMENT)[::-1]


def canonical(seq: bytes) -> bytes:
    rc = reverse_complement(seq)
    return rc if rc < seq else seq


def parse_pattern_list(kmers: Sequence[bytes], reverse_complement_: bool = False, canonical_: bool = False,
                       lowercase: bool = False, uppercase: bool = False) -> List[bytes]:
    pats = [bytes(k) for k in kmers]
    if lowercase:
        pats = [p.lower() for p in pats]
    elif uppercase:
        pats = [p.upper() for p in pats]
    if reverse_complement_:
        pats = pats + [reverse_complement(p) for p in pats]
    if canonical_:
        pats = [canonical(p) for p in pats]
    pats = sorted(set(p for p in pats if p))
    if not pats:
        raise ValueError("No k-mers found in file or provided sequence.")
    return pats


def recommend_aho_corasick(patterns: Sequence[bytes]) -> bool:
    """src/helpers.rs:203-211 — which algorithm the reference would pick (decides log order/counts)."""
    return len(patterns) >= 14 or max(len(p) for p in patterns) > 64
