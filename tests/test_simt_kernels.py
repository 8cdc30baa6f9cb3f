"""The product's CUDA kernels (csrc/bpe_kernels.cuh, unchanged) executed on the CPU SIMT emulator
and compared with the oracle.  This is how kernel logic is validated in the GPU-less container;
the same comparisons run on the real device in test_gpu_parity.py."""
import numpy as np
import pytest

import fuzzgen
import simlib
from conftest import COMBOS, golden_cases
from oracle import oracle


@pytest.fixture(scope="module")
def sim_vocabs(tekken_bytes):
    return {pat: simlib.SimVocab(tekken_bytes, 0, pat, n) for pat, n in COMBOS}


def check_batch(sv, ov, pat, prompts, **kw):
    rc, ids, off, counts, nlong = simlib.encode_batch([sv], prompts, **kw)
    assert rc == 0
    for i, p in enumerate(prompts):
        want = ov.encode(pat, p)
        got = ids[int(off[i]):int(off[i + 1])]
        assert np.array_equal(got, want), (pat, p)
        assert counts[i] == len(want)
    return nlong


@pytest.mark.parametrize("pat", [0, 1, 2, 3])
def test_split_kernel_matches_oracle(pat):
    prompts = [s.encode() for s in fuzzgen.fuzz_strings(70 + pat, 1200) + fuzzgen.long_runs(9)]
    rc, ends = simlib.split([pat], prompts)
    assert rc == 0
    for p, e in zip(prompts, ends):
        assert oracle.split(pat, p).tolist() == e, (pat, p)


@pytest.mark.parametrize("pat,n_ranks", COMBOS)
def test_pipeline_matches_golden(golden, sim_vocabs, pat, n_ranks):
    cases = golden_cases(golden)
    rc, ids, off, counts, _ = simlib.encode_batch([sim_vocabs[pat]], cases)
    assert rc == 0
    gi, go = golden["ids_%d" % pat], golden["id_offsets_%d" % pat]
    assert int(off[-1]) == int(go[-1])
    assert np.array_equal(ids[:int(off[-1])], gi)
    assert np.array_equal(off, go)


@pytest.mark.parametrize("pat", [0, 3])
def test_pipeline_fuzz_and_long_pieces(sim_vocabs, oracle_vocabs, pat):
    prompts = [s.encode() for s in fuzzgen.fuzz_strings(500 + pat, 800, max_atoms=60) + fuzzgen.long_runs(13)]
    nlong = check_batch(sim_vocabs[pat], oracle_vocabs[pat], pat, prompts)
    assert nlong > 0   # the long-piece kernel was exercised


def test_empty_and_tiny_prompts(sim_vocabs, oracle_vocabs):
    prompts = [b"", b"a", b"", b"", b" ", b"\n", b"ab", b"", b"x" * 40, b""]
    check_batch(sim_vocabs[0], oracle_vocabs[0], 0, prompts)
    rc, ids, off, counts, _ = simlib.encode_batch([sim_vocabs[0]], [])
    assert rc == 0 and int(off[0]) == 0
    rc, ids, off, counts, _ = simlib.encode_batch([sim_vocabs[0]], [b"", b""])
    assert rc == 0 and off.tolist() == [0, 0, 0]


def test_multi_vocab_batch(sim_vocabs, oracle_vocabs):
    prompts = [s.encode() for s in fuzzgen.fuzz_strings(31, 300, max_atoms=40)]
    vid = np.array([i % 3 for i in range(len(prompts))], dtype=np.uint8)
    vocs = [sim_vocabs[0], sim_vocabs[1], sim_vocabs[2]]
    rc, ids, off, counts, _ = simlib.encode_batch(vocs, prompts, vocab_ids=vid)
    assert rc == 0
    for i, p in enumerate(prompts):
        pat = int(vid[i])
        assert np.array_equal(ids[int(off[i]):int(off[i + 1])], oracle_vocabs[pat].encode(pat, p)), (pat, p)


def test_bad_utf8_is_reported(sim_vocabs):
    for bad in [b"ok \xff bad", b"\xc3", b"abc\xe2\x82", b"\xed\xa0\x80", b"x\x80y"]:
        rc, *_ = simlib.encode_batch([sim_vocabs[0]], [b"fine", bad, b"also fine"])
        assert rc == -84


def test_out_cap_too_small(sim_vocabs):
    rc, ids, off, counts, _ = simlib.encode_batch([sim_vocabs[0]], [b"hello world, this is a test"], out_cap=2)
    assert rc == -28


def test_roundtrip_decode(sim_vocabs, tekken_bytes):
    """encode -> concatenate token bytes == input (size-independent property used at full size on the GPU)"""
    import base64
    toks = [base64.b64decode(l.split()[0]) for l in tekken_bytes.splitlines()[:100256]]
    prompts = [s.encode() for s in fuzzgen.fuzz_strings(99, 400)]
    rc, ids, off, counts, _ = simlib.encode_batch([sim_vocabs[0]], prompts)
    assert rc == 0
    for i, p in enumerate(prompts):
        assert b"".join(toks[t] for t in ids[int(off[i]):int(off[i + 1])]) == p


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_adversarial_rank_orders(seed):
    """random rank order over a 3-letter alphabet: merged tokens may rank below their parts, so the batched
    rounds of the long-piece kernel must cut exactly where the sequential loop would deviate"""
    import synth_vocab
    rf = synth_vocab.make_rank_file(seed, n_extra=40 + 10 * seed, max_len=4 + seed % 3)
    ov = oracle.OracleVocab(rf)
    sv = simlib.SimVocab(rf, 0, 0, 0)
    prompts = [t.encode() for t in synth_vocab.make_texts(100 + seed, 250)]
    prompts += [t.encode() for t in synth_vocab.make_texts(200 + seed, 10, max_len=3000)]   # list phase in bpe_list_kernel
    nlong = check_batch(sv, ov, 0, prompts)
    assert nlong > 50


def test_long_random_words_list_mode(sim_vocabs, oracle_vocabs):
    import random
    rng = random.Random(5)
    letters = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"
    prompts = []
    for n in [33, 40, 64, 65, 100, 257, 600, 1500, 4096, 6100]:   # 6100: too big for bpe_list_kernel, stays in K2b
        prompts.append("".join(rng.choice(letters) for _ in range(n)).encode())
        prompts.append("".join(rng.choice("etaoinshr") for _ in range(n)).encode())
        prompts.append(("xyz" * n)[:n].encode())
        prompts.append((" " * n).encode())
        prompts.append(("=" * n).encode())
    check_batch(sim_vocabs[0], oracle_vocabs[0], 0, prompts)
    check_batch(sim_vocabs[3], oracle_vocabs[3], 3, prompts)


@pytest.mark.parametrize("pat", [0, 1, 2, 3])
def test_split_exhaustive_short_strings(pat):
    """every string of <= 5 characters over 13 representative characters (402 233 prompts packed in one batch:
    chunk boundaries and sync points fall everywhere)"""
    import itertools
    alpha = ["a", "B", "中", "́", "1", " ", "\t", "\n", "'", "/", "!", "s", "l"]
    strs = ["".join(t).encode() for n in range(1, 6) for t in itertools.product(alpha, repeat=n)]
    rc, ends = simlib.split([pat], strs)
    assert rc == 0
    bad = [(p, e) for p, e in zip(strs, ends) if oracle.split(pat, p).tolist() != e]
    assert not bad, bad[:5]


@pytest.mark.parametrize("pat", [0, 1, 2, 3])
def test_split_long_runs_across_chunks(pat):
    """long letter / mark / punctuation runs with contractions sprinkled in: state-sync hand-over between threads"""
    import random
    rng = random.Random(40 + pat)
    alpha = ["a", "b", "B", "C", "'", "s", "l", "t", "中", "文", "é", "́", " ", "!", "/", "\n", "1"]
    weights = [12, 8, 6, 4, 3, 3, 3, 2, 6, 4, 3, 1, 1, 2, 1, 1, 1]
    strs = []
    for _ in range(1500):
        n = rng.randint(40, 400)
        strs.append("".join(rng.choices(alpha, weights, k=n)).encode())
    rc, ends = simlib.split([pat], strs)
    assert rc == 0
    bad = [(p, e) for p, e in zip(strs, ends) if oracle.split(pat, p).tolist() != e]
    assert not bad, bad[:3]


@pytest.mark.parametrize("pat", [0, 1, 2, 3])
def test_split_bulk_runs(pat):
    """runs of one whitespace byte and of ASCII digits are consumed in bulk by K1 (they hold no sync point): every
    run length modulo 3 and modulo 16, every alignment, every kind of neighbour"""
    import random
    rng = random.Random(70 + pat)
    glue = ["a", "B", "中", "1", "٣", " ", "\t", "\n", "\r\n", "'s", "/", "!", "x y", "", " ", "　"]
    strs = []
    for _ in range(1200):
        parts = []
        for _ in range(rng.randint(1, 6)):
            k = rng.randint(0, 5)
            n = rng.choice([1, 2, 3, 4, 5, 15, 16, 17, 31, 32, 33, 63, 64, 65, rng.randint(1, 300)])
            if k == 0: parts.append(" " * n)
            elif k == 1: parts.append("\n" * n)
            elif k == 2: parts.append("\t" * n)
            elif k == 3: parts.append("".join(rng.choice("0123456789") for _ in range(n)))
            elif k == 4: parts.append("\r" * n)
            else: parts.append("".join(rng.choice(" \n\t\r") for _ in range(n)))
            parts.append(rng.choice(glue))
        strs.append("".join(parts).encode())
    simlib.dbg_counter(3, reset=True); simlib.dbg_counter(4, reset=True)
    rc, ends = simlib.split([pat], strs)
    assert rc == 0
    bad = [(p, e) for p, e in zip(strs, ends) if oracle.split(pat, p).tolist() != e]
    assert not bad, bad[:3]
    assert simlib.dbg_counter(3) > 100 and simlib.dbg_counter(4) > 100


def test_periodic_pieces_switch_back_to_batched_rounds(sim_vocabs, oracle_vocabs):
    """periods with a ragged tail or a foreign character inside: the list phase meets long stretches of one rank
    (strictly ordered by position: one merge per round) and has to hand back to a batched round -- in the warp
    kernel (<= 256 bytes) and in bpe_list_kernel (<= 4096 parts); beyond that the global-memory list path"""
    import random
    rng = random.Random(11)
    letters = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"
    prompts = []
    for n in [60, 130, 200, 256, 257, 700, 1800, 4096, 5200]:
        for _ in range(6):
            per = "".join(rng.choice(letters) for _ in range(rng.randint(2, 5)))
            s = list((per * (n // len(per) + 1))[:n])
            for _ in range(rng.randint(0, 3)):
                s[rng.randrange(n)] = rng.choice(letters)
            prompts.append("".join(s).encode())
    for i in range(8): simlib.dbg_counter(i, reset=True)
    check_batch(sim_vocabs[0], oracle_vocabs[0], 0, prompts)
    check_batch(sim_vocabs[1], oracle_vocabs[1], 1, prompts)
    assert simlib.dbg_counter(0) > 0 and simlib.dbg_counter(1) > 0 and simlib.dbg_counter(2) > 0 and simlib.dbg_counter(5) > 0, \
        [simlib.dbg_counter(i) for i in range(8)]


@pytest.mark.parametrize("pat", [1, 3])
def test_split_cased_runs(pat):
    """cased patterns: runs of upper-case / both-sets (Lo, Lm, M) characters, where the automaton's state depends on how
    the word began -- long CJK and all-caps runs with every kind of character in front of them (lower case, contraction
    suffixes, marks after punctuation, digits, nothing), crossing chunk boundaries everywhere"""
    import random
    rng = random.Random(90 + pat)
    heads = ["", "a", "ab", "abc", "x's", "x'll", "'s", "'", "''", "!", "!!", "1", " ", "\n", "A", "Ab", "aB", "中a", "a中", "́", "'́", "''́", "é", "É"]
    runs = [lambda n: "".join(rng.choice("中文字漢") for _ in range(n)), lambda n: "".join(rng.choice("ABCDÉ") for _ in range(n)),
            lambda n: "".join(rng.choice("中文́AB") for _ in range(n)), lambda n: "".join(rng.choice("中A") for _ in range(n)),
            lambda n: "".join(rng.choice("中文ａʰ") for _ in range(n))]
    tails = ["", "a", "B", "b c", "'s", "!", " x", "1", "\n", "́a"]
    strs = []
    for _ in range(1500):
        parts = []
        for _ in range(rng.randint(1, 4)):
            parts += [rng.choice(heads), rng.choice(runs)(rng.choice([1, 2, 3, 5, 20, 21, 22, 40, 70, 150, 260])), rng.choice(tails)]
        strs.append("".join(parts).encode())
    simlib.split_fixups(reset=True)
    rc, ends = simlib.split([pat], strs)
    assert rc == 0
    bad = [(p, e) for p, e in zip(strs, ends) if oracle.split(pat, p).tolist() != e]
    assert not bad, bad[:3]
    assert simlib.split_fixups() > 50      # threads that started in S_W_U, met an upper-case letter and were finished by the fixup kernel


def test_decode_kernels_invert_encode(sim_vocabs):
    """the decode kernels (ids -> bytes) on the emulator: the inverse of the encode path on fuzz prompts, two vocabularies in one
    batch, empty sequences, more ids than one scan tile; an id outside the vocabulary and a short buffer are reported"""
    prompts = [s.encode() for s in fuzzgen.fuzz_strings(321, 400)] + [b"", b"x", ("word " * 900).encode(), b""]
    vids = np.array([i % 2 for i in range(len(prompts))], dtype=np.uint8)
    rc, ids, off, counts, _ = simlib.encode_batch([sim_vocabs[0], sim_vocabs[1]], prompts, vocab_ids=vids)
    assert rc == 0
    n_ids = int(off[len(prompts)])
    rc, out, boff = simlib.decode_batch([sim_vocabs[0], sim_vocabs[1]], ids[:n_ids], off, vocab_ids=vids)
    assert rc == 0
    data = b"".join(prompts)
    assert bytes(out) == data
    assert boff.tolist() == [0] + list(np.cumsum([len(p) for p in prompts]))
    rc, _, _ = simlib.decode_batch([sim_vocabs[0]], np.array([1, 2, 5_000_000], dtype=np.uint32), np.array([0, 3], dtype=np.uint64))
    assert rc != 0
    rc, _, need = simlib.decode_batch([sim_vocabs[0], sim_vocabs[1]], ids[:n_ids], off, vocab_ids=vids, out_cap=10)
    assert rc != 0 and int(need[len(prompts)]) == len(data)


def test_sub_batch_plan_covers_every_prompt_once():
    """csrc/subbatch.h (host code of the pipelined call): whole prompts, in order, none lost, none empty, at most max_chunks
    sub-batches, sizes ramping up to `chunk` and down again; long prompts, empty prompts, tiny and huge batches"""
    import random
    rng = random.Random(9)
    for trial in range(300):
        n = rng.choice([1, 2, 3, 17, 500, 5000])
        kind = rng.randint(0, 3)
        lens = [rng.choice([0, 1, 7, 4096]) if kind == 0 else rng.randint(0, 4096) if kind == 1 else rng.randint(0, 40) if kind == 2
                else rng.choice([0, 0, 3, 300000]) for _ in range(n)]
        offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        chunk = rng.choice([1, 64, 1000, 12 << 10, 12 << 20])
        mc = rng.choice([9, 10, 16, 64])
        cut = simlib.plan_sub_batches(offs, chunk, mc)
        assert cut[0] == 0 and cut[-1] == n and len(cut) - 1 <= mc, (trial, cut[:5], n)
        assert all(a < b for a, b in zip(cut[:-1], cut[1:])), (trial, cut[:8])
    # the shape on the bench batch: 65 536 prompts of 8..4096 bytes, 12 MiB sub-batches
    lens = np.random.default_rng(3).integers(8, 4097, size=65536)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    cut = simlib.plan_sub_batches(offs, 12 << 20)
    sizes = [int(offs[b] - offs[a]) for a, b in zip(cut[:-1], cut[1:])]
    assert sizes[0] < 2 << 20 and sizes[-1] < 2 << 20 and max(sizes) < 13 << 20 and sizes[0] < sizes[1] < sizes[2] < sizes[3]


def test_split_pattern_changes_at_every_prompt():
    """multi-tenant batches: the pattern (and its casedness, which the context automaton depends on) changes from prompt to
    prompt, with prompts so short that several of them share a 16-byte block of K1"""
    import random
    rng = random.Random(11)
    strs = fuzzgen.fuzz_strings(321, 3000, max_atoms=6) + fuzzgen.fuzz_strings(322, 600, max_atoms=30)
    prompts = [s.encode() for s in strs]
    for i in range(0, len(prompts), 7):
        prompts[i] = prompts[i][:rng.randint(0, 3)].decode("utf-8", "ignore").encode()     # 0..3-byte prompts too
    vids = [rng.randrange(4) for _ in prompts]
    rc, ends = simlib.split([0, 1, 2, 3], prompts, vocab_ids=vids)
    assert rc == 0
    bad = [(v, p, e) for v, p, e in zip(vids, prompts, ends) if oracle.split(v, p).tolist() != e]
    assert not bad, bad[:3]


def test_split_tables_are_small_and_complete():
    """the enumerations behind K1's tables (pretok_ctx.h) stay inside their fixed capacities"""
    L = simlib.lib()
    L.sim_ctx_count.restype = L.sim_prod_count.restype = __import__("ctypes").c_uint32
    assert all(0 < L.sim_ctx_count(c) <= 64 for c in (0, 1))
    assert all(0 < L.sim_prod_count(p) < 128 for p in range(4))


def test_utf8_validation_at_every_alignment():
    """K1 decodes and validates UTF-8 from 16-byte blocks in registers: every malformed sequence must be reported, and no valid
    one, wherever it falls relative to the block boundaries and to the prompt boundaries"""
    bad = [b"\xff", b"\x80", b"\xbf", b"\xc3", b"\xc3(", b"\xe2\x82", b"\xe2(\xa1", b"\xe2\x82(", b"\xed\xa0\x80", b"\xf4\x90\x80\x80",
           b"\xc0\xaf", b"\xc1\xbf", b"\xe0\x80\xaf", b"\xf0\x80\x80\xaf", b"\xf0\x9f\x98", b"\xf5\x80\x80\x80", b"\xe4\xb8\xad\x80",
           b"\xf0\x9f\x98\x80\x80"]
    good = ["é", "中", "😀", "éé中中😀😀", "　", "á", "ʰ"]
    for pad in range(0, 36):
        head = b"x" * pad
        cases_bad = [head + b + tail for b in bad for tail in (b"", b" tail text that goes on for a while")]
        cases_good = [head + g.encode() + tail for g in good for tail in (b"", b" tail text that goes on for a while")]
        for c in cases_bad:
            rc, _ = simlib.split([0], [b"ok", c, b"fine"])
            assert rc == -84, (pad, c)
            with pytest.raises(ValueError):
                oracle.split(0, c)
        rc, ends = simlib.split([3], cases_good)
        assert rc == 0, pad
        for c, e in zip(cases_good, ends):
            assert oracle.split(3, c).tolist() == e, (pad, c)
    # a character cut by the end of its prompt is malformed even when the next prompt starts with the missing bytes
    rc, _ = simlib.split([0], [b"abc\xe4\xb8", b"\xad def"])
    assert rc == -84
    rc, _ = simlib.split([0], [b"x" * 14 + b"\xe4\xb8", b"\xad def"])
    assert rc == -84


@pytest.mark.parametrize("pat", [0, 1, 2, 3])
def test_split_window_overruns(pat):
    """A walker that has met no sync point by the end of its 32-byte window goes on with the product automaton over the next
    blocks (split_extend, up to eight of them), then with the per-character walker.  Runs of 17 .. 200 bytes without a sync point
    (spaces, mixed white space, digits, digits and white space alternating) at every alignment of the run's start, followed by
    everything that can stop them: a letter, a contraction, a character of several bytes, punctuation, the end of the prompt --
    and the same runs with prompt boundaries falling inside them (every prompt restarts the automaton)."""
    import random
    rng = random.Random(900 + pat)
    runs = lambda n: [" " * n, "\t " * (n // 2), "7" * n, "0123456789" * (n // 10 + 1), " \n" * (n // 2), "\n" * n, "12 " * (n // 3), "　" * (n // 3)]
    stops = ["a", "Zebra", "'s", "'ll x", "é", "中文", "!", "?!", "", "\n", "x1", "٣"]
    strs = []
    for n in (17, 24, 31, 32, 33, 47, 48, 49, 64, 100, 159, 160, 161, 200):
        for r in runs(n):
            for lead in range(0, 33, 3):
                strs.append(("w" * lead + "." + r + rng.choice(stops) + " tail").encode())
    rc, ends = simlib.split([pat], strs)
    assert rc == 0
    bad = [(p, e) for p, e in zip(strs, ends) if oracle.split(pat, p).tolist() != e]
    assert not bad, bad[:3]
    assert simlib.dbg_counter(8) >= 0
    # prompt boundaries inside the runs: cut each string at a random byte (on a character boundary) into two prompts
    cut_strs = []
    for s in strs[::3]:
        t = s.decode()
        k = rng.randint(1, max(1, len(t) - 1))
        cut_strs += [t[:k].encode(), t[k:].encode()]
    rc, ends = simlib.split([pat], cut_strs)
    assert rc == 0
    bad = [(p, e) for p, e in zip(cut_strs, ends) if oracle.split(pat, p).tolist() != e]
    assert not bad, bad[:3]


def test_split_window_overruns_across_vocabularies():
    """the same overruns in a multi-vocabulary batch: the prompt after a boundary inside a run may use another pattern"""
    import random
    rng = random.Random(77)
    strs, vids = [], []
    for n in (20, 33, 50, 90, 170):
        for r in (" " * n, "7" * n, " \n" * (n // 2), "12 " * (n // 3)):
            for lead in (0, 5, 14, 15, 16, 27):
                t = "w" * lead + "," + r + rng.choice(["a", "'s", "é", "", "!"]) + " z"
                k = rng.randint(1, len(t) - 1)
                strs += [t[:k].encode(), t[k:].encode()]
                vids += [rng.randrange(4), rng.randrange(4)]
    rc, ends = simlib.split([0, 1, 2, 3], strs, vocab_ids=vids)
    assert rc == 0
    bad = [(v, p, e) for v, p, e in zip(vids, strs, ends) if oracle.split(v, p).tolist() != e]
    assert not bad, bad[:3]
