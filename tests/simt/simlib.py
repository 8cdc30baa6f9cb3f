"""ctypes binding of the SIMT-emulator harness (test infrastructure)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import build as _build  # noqa: E402

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_build.build())
        L.sim_vocab_build.restype = C.c_void_p
        L.sim_vocab_build.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, C.c_char_p, C.c_size_t]
        L.sim_vocab_free.argtypes = [C.c_void_p]
        L.sim_vocab_info.argtypes = [C.c_void_p, C.c_void_p]
        L.sim_piece_lookup.restype = C.c_uint32
        L.sim_piece_lookup.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32]
        L.sim_pair_lookup.restype = C.c_uint32
        L.sim_pair_lookup.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.sim_split.restype = C.c_int
        L.sim_split.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.sim_encode_batch.restype = C.c_int
        L.sim_encode_batch.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.sim_decode_batch.restype = C.c_int
        L.sim_decode_batch.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.sim_plan_sub_batches.restype = C.c_int
        L.sim_plan_sub_batches.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_void_p]
        L.sim_split_fixups.restype = C.c_ulonglong
        L.sim_split_fixups.argtypes = [C.c_int]
        L.sim_vocab_blob.restype = C.c_uint64
        L.sim_vocab_blob.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.sim_validate_blob.restype = C.c_int
        L.sim_validate_blob.argtypes = [C.c_void_p, C.c_uint64, C.c_char_p, C.c_size_t]
        L.sim_dbg_counter.restype = C.c_ulonglong
        L.sim_dbg_counter.argtypes = [C.c_uint32, C.c_int]
        _lib = L
    return _lib


def validate_blob(blob: np.ndarray):
    """(rc, message) of the check cfbpe_vocab_import runs on a packed table blob"""
    err = C.create_string_buffer(256)
    b = np.ascontiguousarray(blob, dtype=np.uint8)
    rc = lib().sim_validate_blob(b.ctypes.data, b.size, err, 256)
    return rc, err.value.decode()


def split_fixups(reset=False):
    """K1 threads that stopped in S_W_U and were finished by pretok_fixup_kernel, in sim_split calls so far"""
    return int(lib().sim_split_fixups(1 if reset else 0))


def dbg_counter(i, reset=False):
    """path counters of csrc/bpe_kernels.cuh (emulator build only): which path did the kernels take"""
    return int(lib().sim_dbg_counter(i, 1 if reset else 0))


class SimVocab:
    def __init__(self, file_bytes, fmt, pattern, max_ranks=0):
        err = C.create_string_buffer(256)
        self._h = lib().sim_vocab_build(file_bytes, len(file_bytes), fmt, pattern, max_ranks, err, 256)
        if not self._h:
            raise ValueError(err.value.decode())
        info = (C.c_uint32 * 6)()
        lib().sim_vocab_info(self._h, info)
        self.n_ranks, self.pattern_id, self.max_token_len, self.n_pair_entries = info[0], info[1], info[2], info[3]
        self.table_bytes = info[4] | (info[5] << 32)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().sim_vocab_free(self._h)
            self._h = None

    def blob(self) -> np.ndarray:
        """the packed tables (what cfbpe_vocab_export hands out)"""
        n = int(lib().sim_vocab_blob(self._h, None, 0))
        out = np.empty(n, dtype=np.uint8)
        lib().sim_vocab_blob(self._h, out.ctypes.data, n)
        return out

    def piece_lookup(self, b: bytes):
        return lib().sim_piece_lookup(self._h, b, len(b))

    def pair_lookup(self, l, r):
        return lib().sim_pair_lookup(self._h, l, r)


def pack(prompts):
    """list[bytes] -> (uint8 array, uint64 offsets)"""
    offs = np.zeros(len(prompts) + 1, dtype=np.uint64)
    if prompts:
        offs[1:] = np.cumsum([len(p) for p in prompts], dtype=np.uint64)
    data = np.frombuffer(b"".join(prompts), dtype=np.uint8).copy() if prompts else np.zeros(0, np.uint8)
    return data, offs


def split(patterns, prompts, vocab_ids=None):
    """piece END offsets per prompt (relative to the prompt) from the K1 bitmask"""
    data, offs = pack(prompts)
    total = int(offs[-1])
    nw = (total + 31) // 32
    bits = np.zeros(nw + 2, dtype=np.uint32)
    pats = np.asarray(patterns, dtype=np.uint32)
    vid = None if vocab_ids is None else np.ascontiguousarray(vocab_ids, dtype=np.uint8)
    dbuf = np.concatenate([data, np.zeros(8, np.uint8)])
    rc = lib().sim_split(pats.ctypes.data, len(pats), len(prompts), dbuf.ctypes.data, offs.ctypes.data,
                         None if vid is None else vid.ctypes.data, bits.ctypes.data)
    flags = np.unpackbits(bits.view(np.uint8), bitorder="little")[:total]
    out = []
    for i in range(len(prompts)):
        a, b = int(offs[i]), int(offs[i + 1])
        starts = np.nonzero(flags[a:b])[0]
        ends = list(starts[1:]) + ([b - a] if b > a else [])
        out.append([int(e) for e in ends])
    return rc, out


def encode_batch(vocabs, prompts, vocab_ids=None, out_cap=None):
    data, offs = pack(prompts)
    total = int(offs[-1])
    cap = total + 1 if out_cap is None else out_cap
    ids = np.zeros(max(cap, 1), dtype=np.uint32)
    out_off = np.zeros(len(prompts) + 1, dtype=np.uint64)
    counts = np.zeros(max(len(prompts), 1), dtype=np.uint32)
    vh = (C.c_void_p * len(vocabs))(*[v._h for v in vocabs])
    vid = None if vocab_ids is None else np.ascontiguousarray(vocab_ids, dtype=np.uint8)
    nlong = C.c_uint64(0)
    dbuf = np.concatenate([data, np.zeros(8, np.uint8)])
    rc = lib().sim_encode_batch(vh, len(vocabs), len(prompts), dbuf.ctypes.data, offs.ctypes.data,
                                None if vid is None else vid.ctypes.data, ids.ctypes.data, cap,
                                out_off.ctypes.data, counts.ctypes.data, C.byref(nlong))
    return rc, ids, out_off, counts[:len(prompts)], nlong.value


def decode_batch(vocabs, ids, id_offsets, vocab_ids=None, out_cap=None):
    """the decode kernels on the emulator: (rc, bytes uint8, byte offsets uint64 n+1)"""
    ids = np.ascontiguousarray(ids, dtype=np.uint32)
    id_offsets = np.ascontiguousarray(id_offsets, dtype=np.uint64)
    n = len(id_offsets) - 1
    cap = int(out_cap) if out_cap is not None else int(len(ids)) * 260 + 64
    out = np.zeros(max(cap, 1), dtype=np.uint8)
    out_off = np.zeros(n + 1, dtype=np.uint64)
    vh = (C.c_void_p * len(vocabs))(*[v._h for v in vocabs])
    vid = None if vocab_ids is None else np.ascontiguousarray(vocab_ids, dtype=np.uint8)
    rc = lib().sim_decode_batch(vh, len(vocabs), n, ids.ctypes.data, id_offsets.ctypes.data,
                                None if vid is None else vid.ctypes.data, out.ctypes.data, cap, out_off.ctypes.data)
    return rc, out[:int(out_off[n])] if rc == 0 else out[:0], out_off


def plan_sub_batches(offsets, chunk, max_chunks=64):
    """csrc/subbatch.h:plan_sub_batches -> list of prompt indices cut[0..nc]"""
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    cut = np.zeros(max_chunks + 1, dtype=np.uint32)
    nc = lib().sim_plan_sub_batches(offsets.ctypes.data, len(offsets) - 1, int(chunk), max_chunks, cut.ctypes.data)
    return cut[:nc + 1].tolist()
