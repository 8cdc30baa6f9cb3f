"""Host table builder (csrc/vocab.cpp): parsers, all-splits pair table, piece tables."""
import base64
import json

import numpy as np
import pytest

import simlib


@pytest.fixture(scope="module")
def toks(tekken_bytes):
    return [base64.b64decode(l.split()[0]) for l in tekken_bytes.splitlines()]


@pytest.fixture(scope="module")
def sv(tekken_bytes):
    return simlib.SimVocab(tekken_bytes, 0, 3, 130072)


def test_info(sv):
    assert sv.n_ranks == 130072 and sv.max_token_len == 76
    assert sv.n_pair_entries == 269443          # valid (left,right) splits of Tekken[:130072]


def test_every_token_found_by_piece_lookup(sv, toks):
    for i in range(0, 130072, 7):
        assert sv.piece_lookup(toks[i]) == i
    for i in range(130072, 130400):
        assert sv.piece_lookup(toks[i]) == 0xFFFFFFFF      # truncated away
    assert sv.piece_lookup(b"definitely not a token \x00\x01") == 0xFFFFFFFF


def test_pair_table_is_all_splits(sv, toks):
    index = {t: i for i, t in enumerate(toks[:130072])}
    rng = np.random.default_rng(1)
    for i in rng.integers(256, 130072, size=3000):
        t = toks[int(i)]
        for k in range(1, len(t)):
            l, r = index.get(t[:k]), index.get(t[k:])
            if l is not None and r is not None:
                assert sv.pair_lookup(l, r) == int(i)
    for _ in range(2000):
        l, r = (int(x) for x in rng.integers(0, 130072, size=2))
        want = index.get(toks[l] + toks[r], 0xFFFFFFFF)
        assert sv.pair_lookup(l, r) == want


def test_tekken_json_parser(toks):
    vocab = [{"rank": i, "token_bytes": base64.b64encode(toks[i]).decode(), "token_str": None if i % 3 else "x\"\\y"}
             for i in range(2000)]
    doc = json.dumps({"config": {"pattern": "p", "vocab": "decoy"}, "vocab": vocab, "image": None}).encode()
    v = simlib.SimVocab(doc, 1, 3, 0)
    assert v.n_ranks == 2000
    for i in range(0, 2000, 13):
        assert v.piece_lookup(toks[i]) == i
    v2 = simlib.SimVocab(doc, 1, 3, 1000)
    assert v2.n_ranks == 1000


@pytest.mark.parametrize("bad", [
    b"", b"AA== 0\n", b"AA==\n", b"!!!! 0\n", b"AA== x\n",
    b"\n".join(base64.b64encode(bytes([i])) + b" " + str(i).encode() for i in range(255)),            # a byte is missing
    b"\n".join(base64.b64encode(bytes([i % 255])) + b" " + str(i).encode() for i in range(256)),      # duplicate token
])
def test_bad_rank_files_are_rejected(bad):
    with pytest.raises(ValueError):
        simlib.SimVocab(bad, 0, 0, 0)
