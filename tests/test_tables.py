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


def test_imported_blob_is_checked_by_content_not_only_by_header():
    """cfbpe_vocab_import takes a blob from another process: the check walks the tables.  A table without free slots
    would spin the device probes forever, a wild offset or id would read outside the blob, an altered token byte changes ids."""
    import struct
    ranks = b"".join(base64.b64encode(bytes([i])) + b" %d\n" % i for i in range(256))
    ranks += b"".join(base64.b64encode(t) + b" %d\n" % (256 + i) for i, t in enumerate([b"ab", b"abc", b"th", b"the", b"a very long token indeed"]))
    v = simlib.SimVocab(ranks, 0, 0)
    good = v.blob()
    assert simlib.validate_blob(good) == (0, "")
    hdr = struct.unpack_from("<6I Q 7Q 4I Q", good, 0)
    off_pair, off_short, off_long, off_tokoff, off_blob = hdr[9], hdr[10], hdr[11], hdr[12], hdr[13]
    cap_pair, cap_long = hdr[14], hdr[16]

    def broken(mut):
        b = good.copy(); mut(b); rc, msg = simlib.validate_blob(b); assert rc != 0, msg; return msg

    assert simlib.validate_blob(good[:40])[0] != 0                                        # truncated
    assert "size" in simlib.validate_blob(np.concatenate([good, np.zeros(16, np.uint8)]))[1]
    def fill_pairs(b):
        b[off_pair:off_pair + 8 * cap_pair].view(np.uint64)[:] = (1 << 42) | (2 << 21) | 3
    assert "half full" in broken(fill_pairs)
    def wild_pair(b):
        b[off_pair:off_pair + 8].view(np.uint64)[0] = (0x1FFFF0 << 42) | (2 << 21) | 3
    assert "outside the vocabulary" in broken(wild_pair)
    def wild_long(b):
        slots = b[off_long:off_long + 16 * cap_long].view(np.uint32).reshape(-1, 4)
        used = np.nonzero(slots[:, 2] != 0xFFFFFFFF)[0]
        slots[used[0], 3] = 0x7FFFFFF0
    assert "long-token" in broken(wild_long)
    def altered_byte(b):
        b[off_blob + 100] ^= 1
    assert "content hash" in broken(altered_byte)
    def bad_tokoff(b):
        b[off_tokoff + 40:off_tokoff + 44].view(np.uint32)[0] = 0xFFFFFF
    assert "token offsets" in broken(bad_tokoff)
