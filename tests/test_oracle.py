"""The oracle against its pins: committed golden vectors (tiktoken 0.12.0 output) and, when the
engine is importable, tiktoken itself on fresh fuzz strings."""
import base64

import numpy as np
import pytest

from conftest import COMBOS, golden_cases
from oracle import oracle
from oracle import patterns as P
import fuzzgen


@pytest.mark.parametrize("pat,n_ranks", COMBOS)
def test_oracle_matches_golden(golden, oracle_vocabs, pat, n_ranks):
    cases = golden_cases(golden)
    ids, offs = golden["ids_%d" % pat], golden["id_offsets_%d" % pat]
    ov = oracle_vocabs[pat]
    assert ov.n_ranks == n_ranks
    for i, c in enumerate(cases):
        want = ids[int(offs[i]):int(offs[i + 1])]
        got = ov.encode(pat, c)
        assert np.array_equal(got, want), (P.PATTERN_NAMES[pat], c)


def test_oracle_batch_matches_single(golden, oracle_vocabs):
    cases = golden_cases(golden)[:600]
    from conftest import pack
    data, offs = pack(cases)
    for pat in (0, 3):
        ids, out_off, counts = oracle.encode_batch([oracle_vocabs[pat]], [pat], data, offs, nthreads=4)
        for i, c in enumerate(cases):
            assert np.array_equal(ids[int(out_off[i]):int(out_off[i + 1])], oracle_vocabs[pat].encode(pat, c))
            assert counts[i] == out_off[i + 1] - out_off[i]


def test_oracle_rejects_bad_utf8(oracle_vocabs):
    for bad in [b"\xff", b"a\x80", b"\xc3", b"\xe2\x82", b"\xed\xa0\x80", b"\xf4\x90\x80\x80", b"\xc0\xaf", b"\xe0\x80\xaf"]:
        with pytest.raises(ValueError):
            oracle_vocabs[0].encode(0, bad)
        with pytest.raises(ValueError):
            oracle.split(0, bad)


def test_oracle_split_covers_text(oracle_vocabs):
    for s in fuzzgen.fuzz_strings(5, 300):
        b = s.encode()
        for pat in range(4):
            ends = oracle.split(pat, b)
            if b:
                assert ends[-1] == len(b) and np.all(np.diff(ends.astype(np.int64)) > 0)
            else:
                assert len(ends) == 0


@pytest.mark.parametrize("pat,n_ranks", COMBOS)
def test_oracle_matches_live_tiktoken(tekken_bytes, oracle_vocabs, pat, n_ranks):
    tiktoken = pytest.importorskip("tiktoken")
    lines = tekken_bytes.splitlines()[:n_ranks]
    ranks = {base64.b64decode(l.split()[0]): i for i, l in enumerate(lines)}
    enc = tiktoken.Encoding("t", pat_str=P.PATTERNS[pat], mergeable_ranks=ranks, special_tokens={})
    strs = fuzzgen.fuzz_strings(1000 + pat, 2500) + fuzzgen.long_runs(50 + pat)
    want = enc.encode_ordinary_batch(strs, num_threads=4)
    for s, w in zip(strs, want):
        assert oracle_vocabs[pat].encode(pat, s.encode()).tolist() == w, repr(s)
