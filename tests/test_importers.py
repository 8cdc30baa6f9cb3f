"""HF tokenizer.json -> rank file importer, checked by a round trip: a rank file is turned into the
tokenizer.json a byte-level BPE trainer would have written for it (merges derived by replaying the merge loop), imported
back, and must come out byte for byte; files that cannot be converted exactly must be refused."""
import base64
import json

import pytest

import synth_vocab
from cfbpe import importers as I


def _bpe_split(rank, tok):
    """the two parts whose merge makes `tok` under rank-by-bytes merging restricted to ranks below tok's own"""
    parts = [bytes([b]) for b in tok]
    limit = rank[tok]
    while len(parts) > 2:
        best, bi = None, -1
        for i in range(len(parts) - 1):
            r = rank.get(parts[i] + parts[i + 1])
            if r is not None and r < limit and (best is None or r < best):
                best, bi = r, i
        assert best is not None, tok
        parts[bi:bi + 2] = [parts[bi] + parts[bi + 1]]
    return parts


def _to_hf(rank_file, pattern=None, specials=()):
    b2u = I.bytes_to_unicode()
    toks = [base64.b64decode(l.split()[0]) for l in rank_file.splitlines() if l.strip()]
    rank = {t: i for i, t in enumerate(toks)}
    enc = lambda t: "".join(b2u[b] for b in t)
    merges = []
    for t in toks:
        if len(t) > 1:
            a, b = _bpe_split(rank, t)
            merges.append(enc(a) + " " + enc(b))
    j = {"model": {"type": "BPE", "vocab": {enc(t): i for i, t in enumerate(toks)}, "merges": merges},
         "added_tokens": [{"id": len(toks) + k, "content": s, "special": True} for k, s in enumerate(specials)],
         "pre_tokenizer": {"type": "Sequence", "pretokenizers": [{"type": "Split", "pattern": {"Regex": pattern}, "behavior": "Isolated"},
                                                                 {"type": "ByteLevel", "add_prefix_space": False}]} if pattern else None}
    for k, s in enumerate(specials):
        j["model"]["vocab"][s] = len(toks) + k
    return json.dumps(j).encode(), toks


def _canonical_rank_file():
    """a small vocabulary whose ranks ARE a merge order (every token's parts rank below it): first 256 bytes, then merges"""
    import random
    rng = random.Random(3)
    toks = [bytes([b]) for b in range(256)]
    have = set(toks)
    while len(toks) < 700:
        a, b = rng.choice(toks[:400]), rng.choice(toks[:400])
        if len(a + b) <= 12 and a + b not in have:
            toks.append(a + b); have.add(a + b)
    return b"".join(base64.b64encode(t) + b" " + str(i).encode() + b"\n" for i, t in enumerate(toks))


def test_round_trip_of_a_merge_ordered_vocabulary():
    rf = _canonical_rank_file()
    llama3 = next(p for p, i in I._KNOWN_PATTERNS.items() if i == 2)
    hf, toks = _to_hf(rf, pattern=llama3, specials=("<|begin_of_text|>", "<|end_of_text|>"))
    out, meta = I.hf_tokenizer_json_to_rank_file(hf)
    assert out == rf
    assert meta["n_ranks"] == len(toks) and meta["pattern_id"] == 2
    assert meta["special_tokens"] == {"<|begin_of_text|>": len(toks), "<|end_of_text|>": len(toks) + 1}
    d = I.tokenizer_descriptor("llama3", 2, out)
    assert d == {"vocab_id": "llama3", "pattern_id": "llama3", "sha256": meta["sha256"]}


def test_files_that_cannot_be_converted_exactly_are_refused():
    rf = _canonical_rank_file()
    hf, _ = _to_hf(rf)
    j = json.loads(hf)
    for breakit in ("wordpiece", "swap_merges", "drop_byte", "gap", "not_json"):
        k = json.loads(hf)
        if breakit == "wordpiece":
            k["model"]["type"] = "WordPiece"
        elif breakit == "swap_merges":
            k["model"]["merges"][3], k["model"]["merges"][40] = k["model"]["merges"][40], k["model"]["merges"][3]
        elif breakit == "drop_byte":
            del k["model"]["vocab"][I.bytes_to_unicode()[0x41]]
        elif breakit == "gap":
            first = next(s for s, i in k["model"]["vocab"].items() if i == 300)
            k["model"]["vocab"][first] = 5000
        data = b"{" if breakit == "not_json" else json.dumps(k).encode()
        with pytest.raises(I.ImportError_):
            I.hf_tokenizer_json_to_rank_file(data)
    assert j["model"]["type"] == "BPE"


def test_rank_orders_that_are_not_a_merge_order_are_refused():
    """tiktoken-style rank files may rank a merged token below its parts (tests/synth_vocab.py builds such files on purpose);
    written as merges they would tokenize differently, so the importer must say so: here two ids are exchanged while the merge
    list stays as it was"""
    rf = _canonical_rank_file()
    hf, toks = _to_hf(rf)
    k = json.loads(hf)
    inv = {i: s for s, i in k["model"]["vocab"].items()}
    k["model"]["vocab"][inv[300]], k["model"]["vocab"][inv[650]] = 650, 300
    with pytest.raises(I.ImportError_):
        I.hf_tokenizer_json_to_rank_file(json.dumps(k).encode())
