"""ctypes binding of the CPU oracle (oracle/bpe_oracle.c).  TEST INFRASTRUCTURE ONLY:
importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; the product package never imports this module."""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "_build", "liboracle.so")


def build(force=False):
    src = os.path.join(_DIR, "bpe_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _DIR, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.oracle_vocab_load.restype = C.c_void_p
        L.oracle_vocab_load.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32]
        L.oracle_vocab_free.argtypes = [C.c_void_p]
        L.oracle_vocab_size.restype = C.c_uint32
        L.oracle_vocab_size.argtypes = [C.c_void_p]
        L.oracle_split.restype = C.c_long
        L.oracle_split.argtypes = [C.c_int, C.c_char_p, C.c_size_t, C.c_void_p]
        L.oracle_encode.restype = C.c_long
        L.oracle_encode.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_size_t, C.c_void_p]
        L.oracle_encode_batch.restype = C.c_int
        L.oracle_encode_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib = L
    return _lib


class OracleVocab:
    def __init__(self, tiktoken_file_bytes: bytes, max_ranks: int = 0):
        self._h = lib().oracle_vocab_load(tiktoken_file_bytes, len(tiktoken_file_bytes), max_ranks)
        if not self._h:
            raise ValueError("oracle: malformed .tiktoken rank file")
        self.n_ranks = lib().oracle_vocab_size(self._h)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().oracle_vocab_free(self._h)
                self._h = None
        except Exception:      # interpreter shutdown: the module globals may be gone already
            pass

    def encode(self, pattern_id: int, data: bytes) -> np.ndarray:
        out = np.empty(max(len(data), 1), dtype=np.uint32)
        n = lib().oracle_encode(self._h, pattern_id, data, len(data), out.ctypes.data)
        if n < 0:
            raise ValueError("oracle: malformed UTF-8" if n == -1 else "oracle: error %d" % n)
        return out[:n].copy()


def split(pattern_id: int, data: bytes):
    """piece END byte offsets (regex find_iter over the text)."""
    out = np.empty(max(len(data), 1), dtype=np.uint32)
    n = lib().oracle_split(pattern_id, data, len(data), out.ctypes.data)
    if n < 0:
        raise ValueError("oracle: malformed UTF-8" if n == -1 else "oracle: error %d" % n)
    return out[:n].copy()


def encode_batch(vocabs, patterns, data: np.ndarray, offsets: np.ndarray, vocab_ids=None, nthreads=1,
                 want_ids=True):
    """vocabs: list[OracleVocab]; patterns: list[int] (indexed by vocab id).
    data: uint8 packed prompt bytes; offsets: uint64 n+1.
    returns (ids uint32, out_offsets uint64 n+1, counts uint32 n)."""
    n = len(offsets) - 1
    data = np.ascontiguousarray(data, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    vh = (C.c_void_p * len(vocabs))(*[v._h for v in vocabs])
    pt = (C.c_int * len(patterns))(*patterns)
    vid = None
    if vocab_ids is not None:
        vocab_ids = np.ascontiguousarray(vocab_ids, dtype=np.uint8)
        vid = vocab_ids.ctypes.data
    total = int(offsets[-1])
    ids = np.empty(max(total, 1), dtype=np.uint32) if want_ids else None
    out_off = np.empty(n + 1, dtype=np.uint64)
    counts = np.empty(max(n, 1), dtype=np.uint32)
    rc = lib().oracle_encode_batch(vh, pt, vid, n, data.ctypes.data, offsets.ctypes.data,
                                   ids.ctypes.data if want_ids else None, out_off.ctypes.data, counts.ctypes.data,
                                   nthreads)
    if rc != 0:
        raise ValueError("oracle: malformed UTF-8 in batch")
    nt = int(out_off[n])
    return (ids[:nt] if want_ids else None), out_off, counts[:n]
