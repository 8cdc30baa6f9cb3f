#!/usr/bin/env python3
"""Generate tests/golden/encode_golden.npz from the stand-in oracle engine (tiktoken 0.12.0
CoreBPE.encode_ordinary) in THIS container.  The reference tree holds no tokenizer, no golden
token vectors and no known-answer tests for this path ("Tests that pin the hot
path's results: none"), so these vectors are what pins oracle/bpe_oracle.c -- and through it
the CUDA path.

Cases: seeded fuzz strings (tests/fuzzgen.py), adversarial runs, slices of the benchmark corpora.
Combos: the four patterns, each with the vocabulary size its benchmark slot uses.
"""
import base64
import os
import sys

import numpy as np
import tiktoken

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "cyberfabric-core_b200"))
import fuzzgen  # noqa: E402
from oracle import patterns as P  # noqa: E402
from cfbpe import workload as W  # noqa: E402

COMBOS = [(P.PAT_CL100K, 100256), (P.PAT_O200K, 150000), (P.PAT_LLAMA3, 128000), (P.PAT_TEKKEN, 130072)]


def main():
    lines = open(os.path.join(ROOT, "vocabs", "tekken_240911.tiktoken"), "rb").read().splitlines()
    toks = [base64.b64decode(l.split()[0]) for l in lines]
    texts = fuzzgen.fuzz_strings(20260921, 1500) + fuzzgen.long_runs(20260922)
    data, offs, _ = W.make_batch(300, 8, 600, seed=77)
    for i in range(300):
        texts.append(bytes(data[int(offs[i]):int(offs[i + 1])]).decode("utf-8"))
    texts += ["", " ", "\n", "a", "'", "'s", "0", "\r\n", "hello world", "Hello World's 1234567 tests!!\n\n  x"]
    blob = "\x00".join(texts)   # NUL never appears inside the cases except as its own punctuation atom
    assert all("\x00\x00" not in t for t in texts)
    enc_texts = [t.encode("utf-8") for t in texts]
    toff = np.zeros(len(texts) + 1, dtype=np.uint64)
    toff[1:] = np.cumsum([len(b) for b in enc_texts])
    out = {"text_bytes": np.frombuffer(b"".join(enc_texts), dtype=np.uint8), "text_offsets": toff,
           "combos": np.array(COMBOS, dtype=np.uint32)}
    for pat, n in COMBOS:
        ranks = {toks[i]: i for i in range(n)}
        enc = tiktoken.Encoding("g%d" % pat, pat_str=P.PATTERNS[pat], mergeable_ranks=ranks, special_tokens={})
        res = enc.encode_ordinary_batch(texts, num_threads=8)
        ioff = np.zeros(len(texts) + 1, dtype=np.uint64)
        ioff[1:] = np.cumsum([len(r) for r in res])
        out["ids_%d" % pat] = np.concatenate([np.asarray(r, dtype=np.uint32) for r in res])
        out["id_offsets_%d" % pat] = ioff
        print(P.PATTERN_NAMES[pat], n, "cases", len(texts), "ids", int(ioff[-1]))
    path = os.path.join(ROOT, "tests", "golden", "encode_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
