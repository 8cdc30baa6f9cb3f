"""Synthetic prompt batches for BASELINE.json's five configurations.

There is no network and no text corpus on the box, and /root/reference is not on the GPU box,
so prompts are drawn from corpora assembled from what ships with the image, deterministically:

  english       CPython's own documentation strings (`pydoc_data.topics`, ~0.5 MB of technical English)
  code          CPython standard-library sources (argparse.py, typing.py, ...)
  multilingual  pseudo-sentences of real words: the vocabulary's own whole-word tokens in Cyrillic,
                Greek, Arabic, Hebrew, Devanagari, Thai, Hangul, Kana and CJK, joined the way each script
                joins words, sprinkled with that script's punctuation
  digits/ws     numbers, tables, indentation runs
  adversarial   single-character runs and long random "words" (one piece per prompt: the worst case of
                the merge loop)

Mix (by prompt): 80 % english(+code 1:4), 10 % multilingual, 5 % digits/whitespace, 5 % adversarial.
Prompts are UTF-8-boundary-snapped slices; generator = numpy PCG64 with the seed SURVEY.md names.
"""
from __future__ import annotations

import base64
import hashlib
import os
import sysconfig
import unicodedata

import numpy as np

from . import vocabs as V

_SCRIPTS = ["CYRILLIC", "GREEK", "ARABIC", "HEBREW", "DEVANAGARI", "THAI", "HANGUL", "HIRAGANA", "KATAKANA", "CJK"]
_SPACELESS = {"THAI", "HIRAGANA", "KATAKANA", "CJK"}
_PUNCT = {"CJK": ["，", "。", "、", "？"], "HIRAGANA": ["、", "。"], "KATAKANA": ["・", "。"],
          "ARABIC": ["،", "."], "default": [",", ".", ";", "?", "!"]}

_cache = {}


def _english() -> bytes:
    import pydoc_data.topics as T
    return "\n\n".join(T.topics[k] for k in sorted(T.topics)).encode("utf-8")


def _code() -> bytes:
    std = sysconfig.get_paths()["stdlib"]
    out = []
    for name in ["argparse.py", "typing.py", "dataclasses.py", "json/decoder.py", "textwrap.py", "heapq.py", "bisect.py"]:
        p = os.path.join(std, name)
        if os.path.exists(p):
            with open(p, "rb") as f:
                out.append(f.read())
    return b"\n".join(out)


def _script_of(word: str):
    s = None
    for ch in word:
        if not ch.isalpha():
            return None
        try:
            nm = unicodedata.name(ch)
        except ValueError:
            return None
        sc = next((x for x in _SCRIPTS if nm.startswith(x)), None)
        if sc is None or (s is not None and sc != s):
            return None
        s = sc
    return s


def _multilingual(seed: int, target_bytes: int = 1 << 20) -> bytes:
    words = {s: [] for s in _SCRIPTS}
    with open(V.TEKKEN_FILE, "rb") as f:
        for line in f:
            tok = base64.b64decode(line.split()[0])
            if len(tok) < 4:
                continue
            try:
                w = tok.decode("utf-8")
            except UnicodeDecodeError:
                continue
            w = w.lstrip(" ")
            sc = _script_of(w) if w else None
            if sc:
                words[sc].append(w)
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    size = 0
    scripts = [s for s in _SCRIPTS if len(words[s]) >= 20]
    while size < target_bytes:
        sc = scripts[int(rng.integers(len(scripts)))]
        ws = words[sc]
        n = int(rng.integers(4, 24))
        sep = "" if sc in _SPACELESS else " "
        sent = sep.join(ws[int(i)] for i in rng.integers(len(ws), size=n))
        p = _PUNCT.get(sc, _PUNCT["default"])
        sent += p[int(rng.integers(len(p)))] + ("\n" if rng.random() < 0.15 else " ")
        b = sent.encode("utf-8")
        out.append(b)
        size += len(b)
    return b"".join(out)


def _digits_ws(seed: int, target_bytes: int = 1 << 19) -> bytes:
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    size = 0
    while size < target_bytes:
        k = int(rng.integers(5))
        if k == 0:
            s = " ".join(str(int(x)) for x in rng.integers(0, 10 ** int(rng.integers(1, 12)), size=8)) + "\n"
        elif k == 1:
            s = " " * int(rng.integers(1, 24)) + "x = %d;\n" % int(rng.integers(1 << 30))
        elif k == 2:
            s = "\t".join("%.4f" % x for x in rng.random(6)) + "\r\n"
        elif k == 3:
            s = "\n" * int(rng.integers(1, 5)) + "  - item %d:  %s\n" % (int(rng.integers(1000)), "=" * int(rng.integers(1, 20)))
        else:
            v = [int(x) for x in rng.integers(1, 28, size=5)]
            s = "2026-%02d-%02dT%02d:%02d:%02dZ 0x%08x %d%%\n" % (v[0] % 12 + 1, v[1], v[2] % 24, v[3], v[4], int(rng.integers(1 << 31)), int(rng.integers(101)))
        b = s.encode()
        out.append(b)
        size += len(b)
    return b"".join(out)


def corpora(seed: int):
    key = ("corpora", seed)
    if key not in _cache:
        _cache[key] = {
            "english": _english(), "code": _code(),
            "multilingual": _multilingual(seed * 7919 + 1), "digits_ws": _digits_ws(seed * 7919 + 2),
        }
    return _cache[key]


def _snap_tables(buf: np.ndarray):
    """next_start[i]: first char start >= i ; prev_start[i]: last char start <= i (i in 0..n)"""
    n = len(buf)
    is_start = np.ones(n + 1, dtype=bool)
    is_start[:n] = (buf & 0xC0) != 0x80
    idx = np.arange(n + 1, dtype=np.int64)
    prev_start = np.maximum.accumulate(np.where(is_start, idx, 0))
    nxt = np.where(is_start, idx, n)
    next_start = np.minimum.accumulate(nxt[::-1])[::-1]
    return next_start, prev_start


def _slices(buf: bytes, lengths: np.ndarray, rng) -> list:
    a = np.frombuffer(buf, dtype=np.uint8)
    n = len(a)
    ns, ps = _snap_tables(a)
    starts = rng.integers(0, max(n - 1, 1), size=len(lengths))
    out = []
    for s0, ln in zip(starts, lengths):
        s = int(ns[int(s0)])
        e = int(ps[min(s + int(ln), n)])
        if e <= s:
            s = 0
            e = int(ps[min(int(ln), n)])
        out.append(a[s:e])
    return out


def _adversarial(lengths: np.ndarray, rng) -> list:
    letters = np.frombuffer(b"abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ", dtype=np.uint8)
    out = []
    for ln in lengths:
        ln = int(ln)
        k = int(rng.integers(4))
        if k == 0:      # one repeated character
            ch = b"a !\n0"[int(rng.integers(5))]
            out.append(np.full(ln, ch, dtype=np.uint8))
        elif k == 1:    # one long random "word"
            out.append(letters[rng.integers(len(letters), size=ln)])
        elif k == 2:    # short period repeats: abababab...
            per = letters[rng.integers(len(letters), size=int(rng.integers(2, 5)))]
            out.append(np.resize(per, ln))
        else:           # long lowercase word
            out.append(letters[rng.integers(26, size=ln)])
    return out


def make_batch(n_prompts: int, min_len: int, max_len: int, seed: int, mix=(0.80, 0.10, 0.05, 0.05)):
    """returns (bytes uint8, offsets uint64 n+1, meta dict).  Lengths are i.i.d. uniform integers in
    [min_len, max_len] before UTF-8 boundary snapping; meta reports the realised total."""
    rng = np.random.Generator(np.random.PCG64(seed))
    C = corpora(seed)
    lengths = rng.integers(min_len, max_len + 1, size=n_prompts)
    kind = rng.choice(4, size=n_prompts, p=list(mix))
    parts = [None] * n_prompts
    eng = C["english"] + b"\n\n" + C["code"]
    for k, src in ((0, eng), (1, C["multilingual"]), (2, C["digits_ws"])):
        idx = np.nonzero(kind == k)[0]
        if len(idx):
            for i, sl in zip(idx, _slices(src, lengths[idx], rng)):
                parts[i] = sl
    idx = np.nonzero(kind == 3)[0]
    if len(idx):
        for i, sl in zip(idx, _adversarial(lengths[idx], rng)):
            parts[i] = sl
    offs = np.zeros(n_prompts + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(p) for p in parts], dtype=np.uint64)
    data = np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint8)
    meta = {"n_prompts": n_prompts, "total_bytes": int(offs[-1]), "min_len": min_len, "max_len": max_len, "seed": seed,
            "mix": {"english+code": mix[0], "multilingual": mix[1], "digits_ws": mix[2], "adversarial": mix[3]},
            "sha256": hashlib.sha256(data.tobytes()).hexdigest()[:16]}
    return np.ascontiguousarray(data), offs, meta


# BASELINE.json configs -> concrete inputs; vocab names resolve through cfbpe.vocabs
CONFIGS = {
    1: dict(name="1x512B cl100k", n=1, min_len=512, max_len=512, seed=1, vocabs=["cl100k_base"]),
    2: dict(name="1Kx512B cl100k", n=1024, min_len=512, max_len=512, seed=2, vocabs=["cl100k_base"]),
    3: dict(name="64K mixed 8-4096B cl100k", n=65536, min_len=8, max_len=4096, seed=3, vocabs=["cl100k_base"]),
    4: dict(name="16Kx1KiB o200k", n=16384, min_len=1024, max_len=1024, seed=4, vocabs=["o200k_base"]),
    5: dict(name="256 tenants x 256 prompts, vocab = tenant mod 3", n=65536, min_len=8, max_len=4096, seed=5,
            vocabs=["cl100k_base", "o200k_base", "llama3"], tenants=256),
}


def make_config(cfg_id: int, scale: float = 1.0):
    """(bytes, offsets, vocab_ids or None, meta) of a BASELINE.json config; scale < 1 shrinks n_prompts (tests)."""
    c = CONFIGS[cfg_id]
    n = max(1, int(round(c["n"] * scale)))
    data, offs, meta = make_batch(n, c["min_len"], c["max_len"], c["seed"])
    vid = None
    if "tenants" in c:
        per = max(1, n // c["tenants"])
        tenant = np.minimum(np.arange(n) // per, c["tenants"] - 1)
        vid = (tenant % len(c["vocabs"])).astype(np.uint8)
    meta.update(config=cfg_id, name=c["name"], vocabs=c["vocabs"])
    return data, offs, vid, meta
