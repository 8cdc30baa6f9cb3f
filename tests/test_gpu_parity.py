"""Parity tests proper: the CUDA path, called through the C ABI (libcfbpe.so), against the CPU oracle
and the committed golden vectors.  Bit-exact (integer work): no tolerance anywhere."""
import base64
import os

import numpy as np
import pytest

import fuzzgen
from conftest import COMBOS, golden_cases, pack

pytestmark = pytest.mark.gpu

SLOT_NAMES = {0: "cl100k_base", 1: "o200k_base", 2: "llama3", 3: "tekken"}


@pytest.fixture(scope="module")
def plug():
    from cfbpe import plugin as P
    p = P.GpuBpeTokenizerPlugin(device=0, vocab_names=("cl100k_base", "o200k_base", "llama3", "tekken"),
                                max_batch_bytes=160 << 20, max_prompts=1 << 17)
    yield p
    p.close()


@pytest.fixture(scope="module")
def ctx():
    from cfbpe import plugin as P
    return P.SecurityContext.anonymous()


def encode(plug, ctx, name, data, offs, per_prompt=None):
    from cfbpe import plugin as P
    return plug.encode_batch(ctx, P.EncodeBatchRequest(P.VocabRef(name), data, offs, per_prompt))


def test_native_library_is_the_cuda_one(plug):
    """the product path is libcfbpe.so in-tree and nothing else"""
    from cfbpe import _native
    assert os.path.samefile(_native.SO_PATH, os.path.join(os.path.dirname(_native.__file__), "libcfbpe.so"))
    with open("/proc/self/maps") as f:
        assert any("libcfbpe.so" in line for line in f)


@pytest.mark.parametrize("pat,n_ranks", COMBOS)
def test_golden_vectors(plug, ctx, golden, pat, n_ranks):
    cases = golden_cases(golden)
    data, offs = pack(cases)
    r = encode(plug, ctx, SLOT_NAMES[pat], data, offs)
    assert np.array_equal(r.offsets, golden["id_offsets_%d" % pat])
    assert np.array_equal(r.ids, golden["ids_%d" % pat])


@pytest.mark.parametrize("pat,n_ranks", COMBOS)
def test_fuzz_against_oracle(plug, ctx, oracle_vocabs, pat, n_ranks):
    from oracle import oracle
    prompts = [s.encode() for s in fuzzgen.fuzz_strings(9000 + pat, 20000, max_atoms=48) + fuzzgen.long_runs(77)]
    data, offs = pack(prompts)
    want_ids, want_off, want_counts = oracle.encode_batch([oracle_vocabs[pat]], [pat], data, offs, nthreads=os.cpu_count())
    r = encode(plug, ctx, SLOT_NAMES[pat], data, offs)
    assert np.array_equal(r.offsets, want_off)
    assert np.array_equal(r.ids, want_ids)
    assert np.array_equal(r.counts, want_counts)


@pytest.mark.parametrize("cfg_id", [1, 2, 3, 4])
def test_benchmark_configs_full_size_against_oracle(plug, ctx, cfg_id):
    """BASELINE.json configs 1-4 at FULL size: every id, offset and count equal to the oracle's (no sampling, no scaling)"""
    from cfbpe import plugin as P
    from cfbpe import workload as W
    from oracle import oracle
    data, offs, vid, meta = W.make_config(cfg_id, 1.0)
    name = meta["vocabs"][0]
    rv = plug.resolved[name]
    ov = oracle.OracleVocab(rv.file_bytes, rv.max_ranks)
    want_ids, want_off, want_counts = oracle.encode_batch([ov], [rv.pattern_id], data, offs, nthreads=os.cpu_count())
    r = encode(plug, ctx, name, data, offs)
    assert np.array_equal(r.offsets, want_off)
    assert np.array_equal(r.ids, want_ids)
    assert np.array_equal(r.counts, want_counts)
    counts = plug.count_tokens(ctx, P.CountTokensRequest(P.VocabRef(name), data, offs))
    assert np.array_equal(counts, want_counts)


def test_multi_tenant_vocab_mix_full_size_against_oracle(plug, ctx):
    """BASELINE.json config 5 at FULL size: 256 tenants x 256 prompts, vocabulary = tenant mod 3, one batch"""
    from cfbpe import plugin as P
    from cfbpe import workload as W
    from oracle import oracle
    data, offs, vid, meta = W.make_config(5, 1.0)
    names = meta["vocabs"]
    ovs, pats = [], []
    for nm in names:
        rv = plug.resolved[nm]
        ovs.append(oracle.OracleVocab(rv.file_bytes, rv.max_ranks))
        pats.append(rv.pattern_id)
    want_ids, want_off, want_counts = oracle.encode_batch(ovs, pats, data, offs, vocab_ids=vid, nthreads=os.cpu_count())
    refs = [P.VocabRef(names[int(v)]) for v in vid]
    r = encode(plug, ctx, names[0], data, offs, per_prompt=refs)
    assert np.array_equal(r.offsets, want_off)
    assert np.array_equal(r.ids, want_ids)
    assert np.array_equal(r.counts, want_counts)


def _tiktoken_encoding(tekken_bytes, pat, n_ranks, special=None):
    tiktoken = pytest.importorskip("tiktoken")
    from oracle import patterns as PT
    lines = tekken_bytes.splitlines()[:n_ranks]
    ranks = {base64.b64decode(l.split()[0]): i for i, l in enumerate(lines)}
    return tiktoken.Encoding("live%d" % pat, pat_str=PT.PATTERNS[pat], mergeable_ranks=ranks, special_tokens=special or {})


@pytest.mark.parametrize("pat,n_ranks", COMBOS)
def test_cuda_path_directly_against_live_tiktoken(plug, ctx, tekken_bytes, pat, n_ranks):
    """the pin without the C port in between: libcfbpe.so vs tiktoken 0.12.0 `encode_ordinary_batch` on the same strings
    (fuzz strings, adversarial runs, and the prompts of BASELINE.json config 2)"""
    from cfbpe import workload as W
    enc = _tiktoken_encoding(tekken_bytes, pat, n_ranks)
    strs = fuzzgen.fuzz_strings(31000 + pat, 6000, max_atoms=48) + fuzzgen.long_runs(91 + pat)
    d2, o2, _, _ = W.make_config(2, 1.0)
    strs += [bytes(d2[int(o2[i]):int(o2[i + 1])]).decode("utf-8") for i in range(len(o2) - 1)]
    want = enc.encode_ordinary_batch(strs, num_threads=min(32, os.cpu_count() or 1))
    data, offs = pack([s.encode() for s in strs])
    r = encode(plug, ctx, SLOT_NAMES[pat], data, offs)
    want_off = np.zeros(len(strs) + 1, dtype=np.uint64)
    want_off[1:] = np.cumsum([len(w) for w in want])
    assert np.array_equal(r.offsets, want_off)
    assert np.array_equal(r.ids, np.concatenate([np.asarray(w, dtype=np.uint32) for w in want]))


def test_encode_with_special_on_device_against_tiktoken(plug, ctx, tekken_bytes):
    """special tokens over the CUDA path vs `tiktoken.Encoding.encode(text, allowed_special=...)`"""
    from cfbpe import plugin as P
    special = {"<|endoftext|>": 100257, "<|fim_prefix|>": 100258, "<|endofprompt|>": 100276}
    enc = _tiktoken_encoding(tekken_bytes, 0, 100256, special)
    hub = P.ClientHub()
    hub.register_scoped(P.TokenizerPluginClient, plug.instance.id, plug)
    svc = P.LlmGatewayTokenizerService(hub, [plug.instance])
    base = fuzzgen.fuzz_strings(77, 400, max_atoms=30)
    keys = list(special)
    texts = []
    for i, t in enumerate(base):
        k = keys[i % 3]
        texts.append([t, k + t, t + k, t[: len(t) // 2] + k + t[len(t) // 2:] + k + k, k][i % 5])
    got = svc.encode_with_special(ctx, "cl100k_base", texts, special, allowed_special="all")
    for t, g in zip(texts, got):
        assert g.tolist() == enc.encode(t, allowed_special="all"), repr(t)
    only = {"<|endoftext|>"}
    plain = [t for t in texts if "<|fim_prefix|>" not in t and "<|endofprompt|>" not in t]
    got = svc.encode_with_special(ctx, "cl100k_base", plain, special, allowed_special=only)
    for t, g in zip(plain, got):
        assert g.tolist() == enc.encode(t, allowed_special=only), repr(t)
    with pytest.raises(P.InvalidInput):          # tiktoken raises ValueError: the default disallows every special token
        svc.encode_with_special(ctx, "cl100k_base", ["a <|endoftext|> b"], special)
    # disallowed_special=() : the special token's text is encoded as ordinary text
    got = svc.encode_with_special(ctx, "cl100k_base", ["a <|endoftext|> b"], special, allowed_special=(), disallowed_special=())
    assert got[0].tolist() == enc.encode("a <|endoftext|> b", allowed_special=set(), disallowed_special=())


def test_chat_template_accounting_on_device_against_tiktoken(plug, ctx, tekken_bytes):
    """Usage.input_tokens with the provider's framing: the Llama 3 template over the CUDA path vs tiktoken on
    the rendered conversation with the template's control tokens as special tokens; 300 conversations in the fuzz alphabet"""
    from cfbpe import plugin as P
    tpl = P.CHAT_TEMPLATES["llama3-instruct"]
    enc = _tiktoken_encoding(tekken_bytes, 2, 128000, {t: 128000 + i for i, t in enumerate(tpl.special_tokens)})
    hub = P.ClientHub()
    hub.register_scoped(P.TokenizerPluginClient, plug.instance.id, plug)
    svc = P.LlmGatewayTokenizerService(hub, [plug.instance])
    texts = [t for t in fuzzgen.fuzz_strings(4711, 1200, max_atoms=24) if "<|" not in t]
    roles = ["system", "user", "assistant", "tool"]
    for c in range(300):
        conv = [{"role": roles[(c + i) % 4], "content": [{"type": "text", "text": texts[(4 * c + i) % len(texts)]}]} for i in range(1 + c % 4)]
        rendered = tpl.bos + "".join(tpl.message_prefix.format(role=m["role"]) + m["content"][0]["text"] + tpl.message_suffix for m in conv) + tpl.generation_prompt
        assert svc.count_chat_tokens(ctx, "llama3", conv, tpl).input_tokens == len(enc.encode(rendered, allowed_special="all")), rendered


def test_micro_batcher_on_device_from_64_threads(plug, ctx, oracle_vocabs):
    """SURVEY.md 8(f) item 4 on the real plugin: 64 request threads, two vocabularies, every caller gets its own counts;
    one bad request fails alone"""
    import threading
    from cfbpe import plugin as P
    mb = P.CountTokensMicroBatcher(plug, max_batch_bytes=4 << 20, max_wait_s=0.001).start()
    texts = fuzzgen.fuzz_strings(555, 64 * 24, max_atoms=40)
    models = ["cl100k_base", "tekken"]
    pat_of = {"cl100k_base": 0, "tekken": 3}
    results, errors = {}, []

    def worker(t):
        try:
            for j in range(8):
                mine = texts[(t * 8 + j) * 3:(t * 8 + j) * 3 + 3]
                m = models[(t + j) % 2]
                results[(t, j)] = (m, mine, mb.count(ctx, m, mine, timeout=60))
        except Exception as e:   # noqa: BLE001
            errors.append(e)
    th = [threading.Thread(target=worker, args=(t,)) for t in range(64)]
    for x in th:
        x.start()
    bad = None
    try:
        mb.count(ctx, "no-such-model", ["x"], timeout=60)
    except P.TokenizerError as e:
        bad = e
    for x in th:
        x.join()
    mb.stop()
    assert not errors, errors[:1]
    assert isinstance(bad, P.VocabNotFound)
    assert len(results) == 64 * 8 and mb.batches < mb.items
    for m, mine, counts in results.values():
        want = [len(oracle_vocabs[pat_of[m]].encode(pat_of[m], s.encode())) for s in mine]
        assert counts.tolist() == want


def test_full_size_config3_properties(plug, ctx, tekken_bytes):
    """BASELINE.json config 3 at full size (65 536 prompts, ~134 MB): size-independent properties --
    decode(encode(x)) == x byte for byte, offsets monotone and consistent with counts, count_tokens == encode counts,
    and a sampled subset bit-exact against the oracle."""
    from cfbpe import plugin as P
    from cfbpe import workload as W
    from oracle import oracle
    data, offs, vid, meta = W.make_config(3, 1.0)
    r = encode(plug, ctx, "cl100k_base", data, offs)
    n = len(offs) - 1
    assert len(r.offsets) == n + 1 and r.offsets[0] == 0 and int(r.offsets[-1]) == len(r.ids)
    assert np.array_equal(np.diff(r.offsets.astype(np.int64)), r.counts.astype(np.int64))
    counts = plug.count_tokens(ctx, P.CountTokensRequest(P.VocabRef("cl100k_base"), data, offs))
    assert np.array_equal(counts, r.counts)
    # round trip: token byte lengths must re-tile every prompt, and the bytes must match
    rv = plug.resolved["cl100k_base"]
    toks = [base64.b64decode(l.split()[0]) for l in tekken_bytes.splitlines()[:rv.max_ranks]]
    tlen = np.array([len(t) for t in toks], dtype=np.int64)
    assert r.ids.max() < len(toks)
    per_prompt_bytes = np.add.reduceat(tlen[r.ids], r.offsets[:-1].astype(np.int64))
    assert np.array_equal(per_prompt_bytes, np.diff(offs.astype(np.int64)))
    blob = np.frombuffer(b"".join(toks), dtype=np.uint8)
    tstart = np.concatenate([[0], np.cumsum(tlen)[:-1]])
    rng = np.random.default_rng(0)
    for i in rng.integers(0, n, size=300):
        ids = r.ids[int(r.offsets[i]):int(r.offsets[i + 1])]
        dec = b"".join(toks[t] for t in ids)
        assert dec == bytes(data[int(offs[i]):int(offs[i + 1])])
    # ... and the whole batch through the device decode path: byte for byte the input, same prompt boundaries
    dec = plug.decode_batch(ctx, P.DecodeBatchRequest(P.VocabRef("cl100k_base"), r.ids, r.offsets))
    assert np.array_equal(dec.offsets, offs) and np.array_equal(dec.bytes, data)
    ov = oracle.OracleVocab(rv.file_bytes, rv.max_ranks)
    sel = rng.integers(0, n, size=2000)
    for i in sel:
        want = ov.encode(rv.pattern_id, bytes(data[int(offs[i]):int(offs[i + 1])]))
        assert np.array_equal(r.ids[int(r.offsets[i]):int(r.offsets[i + 1])], want)


def test_edge_cases(plug, ctx):
    from cfbpe import plugin as P
    for prompts in ([], [b""], [b"", b"", b""], [b"a"], [b"", b"a", b""], [b" " * 5000], [b"\n" * 33, b"", b"x" * 4096]):
        data, offs = pack(prompts)
        r = encode(plug, ctx, "cl100k_base", data, offs)
        assert len(r.offsets) == len(prompts) + 1
        assert int(r.offsets[-1]) == len(r.ids)
        for i, p in enumerate(prompts):
            if not p:
                assert r.counts[i] == 0


def test_error_codes(plug, ctx):
    from cfbpe import plugin as P
    data, offs = pack([b"fine", b"bad \xff\xfe", b"ok"])
    with pytest.raises(P.InvalidInput):
        encode(plug, ctx, "cl100k_base", data, offs)
    with pytest.raises(P.VocabNotFound):
        encode(plug, ctx, "no-such-vocab", data, offs)
    bad_offs = np.array([0, 5, 3], dtype=np.uint64)
    with pytest.raises(P.InvalidInput):
        encode(plug, ctx, "cl100k_base", np.zeros(8, np.uint8), bad_offs)
    # ENOSPC: out buffer too small reports the needed size
    from cfbpe import _native as N
    d, o = pack([b"hello world, hello world, hello world"])
    out_ids = np.zeros(2, dtype=np.uint32)
    with pytest.raises(N.NativeError) as ei:
        plug.ctx.encode_batch(d, o, None, out_ids)
    assert ei.value.code == N.ENOSPC
    # the context still works afterwards
    r = encode(plug, ctx, "cl100k_base", *pack([b"still alive"]))
    assert len(r.ids) > 0


def test_device_resident_api_matches_host_api(plug, ctx):
    import torch
    from cfbpe import workload as W
    data, offs, vid, meta = W.make_config(2, 1.0)
    r = encode(plug, ctx, "cl100k_base", data, offs)
    dev = torch.device("cuda:0")
    d_bytes = torch.from_numpy(np.concatenate([data, np.zeros(64, np.uint8)])).to(dev)
    d_offs = torch.from_numpy(offs.astype(np.int64)).to(dev)
    n = len(offs) - 1
    d_ids = torch.zeros(len(data) + 1, dtype=torch.int32, device=dev)
    d_out_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    d_counts = torch.zeros(n, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    nt = plug.ctx.encode_batch_device(n, d_bytes.data_ptr(), int(offs[-1]), d_offs.data_ptr(), None, d_ids.data_ptr(),
                                      d_ids.numel(), d_out_off.data_ptr(), d_counts.data_ptr(), stream, sync=True)
    assert nt == len(r.ids)
    assert np.array_equal(d_ids[:nt].cpu().numpy().view(np.uint32), r.ids)
    assert np.array_equal(d_out_off.cpu().numpy().astype(np.uint64), r.offsets)


def test_vocab_export_import_roundtrip(plug, ctx):
    from cfbpe import _native as N
    blob = plug.export_vocab("tekken")
    c2 = N.Context(0, 4 << 20, 1024)
    c2.vocab_import(0, blob)
    data, offs = pack([s.encode() for s in fuzzgen.fuzz_strings(3, 500)])
    a = plug.ctx.encode_batch(data, offs, np.full(len(offs) - 1, plug._slot["tekken"], dtype=np.uint8))
    b = c2.encode_batch(data, offs)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    with pytest.raises(N.NativeError):
        c2.vocab_import(1, blob[:1000])
    c2.close()


def test_usage_count_tokens_service(plug, ctx):
    from cfbpe import plugin as P
    hub = P.ClientHub()
    hub.register_scoped(P.TokenizerPluginClient, plug.instance.id, plug)
    svc = P.LlmGatewayTokenizerService(hub, [plug.instance])
    msgs = [{"role": "user", "content": [{"type": "text", "text": "Hello there, how's it going?"},
                                         {"type": "image", "url": "x"}]},
            {"role": "assistant", "content": [{"type": "text", "text": "Fine."}]}]
    u = svc.count_tokens(ctx, "openai::gpt-4", msgs)
    ids = svc.encode(ctx, "openai::gpt-4", ["Hello there, how's it going?", "Fine."])
    assert u.input_tokens == sum(len(x) for x in ids) > 0
    assert svc.check_budget(ctx, "openai::gpt-4", msgs, u.input_tokens)
    assert not svc.check_budget(ctx, "openai::gpt-4", msgs, u.input_tokens - 1)


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6, 7, 8])
def test_adversarial_rank_orders_on_device(seed):
    """random rank order (merged tokens may rank below their parts): guard of the batched merge rounds"""
    import synth_vocab
    from cfbpe import _native as N
    from oracle import oracle
    rf = synth_vocab.make_rank_file(seed, n_extra=40 + 10 * seed, max_len=4 + seed % 3)
    ov = oracle.OracleVocab(rf)
    c = N.Context(0, 8 << 20, 1 << 14)
    c.vocab_load(0, rf, N.FORMAT_TIKTOKEN, seed % 4, 0)
    prompts = [t.encode() for t in synth_vocab.make_texts(100 + seed, 3000, max_len=700)]
    data, offs = pack(prompts)
    want_ids, want_off, _ = oracle.encode_batch([ov], [seed % 4], data, offs, nthreads=os.cpu_count())
    got_ids, got_off, _ = c.encode_batch(data, offs)
    assert np.array_equal(got_off, want_off)
    assert np.array_equal(got_ids, want_ids)
    c.close()


def test_long_pieces_on_device(plug, ctx, oracle_vocabs):
    import random
    from oracle import oracle
    rng = random.Random(5)
    letters = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"
    prompts = []
    for n in [33, 40, 64, 65, 100, 257, 600, 1500, 4096, 20000]:
        for _ in range(8):
            prompts.append("".join(rng.choice(letters) for _ in range(n)).encode())
            prompts.append("".join(rng.choice("etaoinshr") for _ in range(n)).encode())
        prompts += [("xyz" * n)[:n].encode(), (" " * n).encode(), ("=" * n).encode(), ("ab" * n)[:n].encode(), ("\n" * n).encode()]
    data, offs = pack(prompts)
    for pat in (0, 1, 3):
        want_ids, want_off, _ = oracle.encode_batch([oracle_vocabs[pat]], [pat], data, offs, nthreads=os.cpu_count())
        r = encode(plug, ctx, SLOT_NAMES[pat], data, offs)
        assert np.array_equal(r.offsets, want_off)
        assert np.array_equal(r.ids, want_ids)


@pytest.mark.parametrize("pat", [1, 3])
def test_cased_runs_and_periodic_pieces_on_device(plug, ctx, oracle_vocabs, pat):
    """what the second half of round 1 added: threads that start in the undecided states W_U / W_V and are finished by
    pretok_fixup_kernel (CJK and all-caps runs under the cased patterns), whitespace / digit runs taken in bulk, periodic
    pieces that make the list kernels hand back to batched rounds, pieces of 4097.. parts (global-memory list path)"""
    import random
    from oracle import oracle
    rng = random.Random(31 + pat)
    heads = ["", "a", "abc", "x's", "x'LL", "'", "''", "!", "1", " ", "\n", "A", "aB", "a中", "́", "''́", "é"]
    runs = [lambda n: "".join(rng.choice("中文字漢") for _ in range(n)), lambda n: "".join(rng.choice("ABCDÉ") for _ in range(n)),
            lambda n: "".join(rng.choice("中文́AB") for _ in range(n)), lambda n: "".join(rng.choice("中A") for _ in range(n)),
            lambda n: " " * n, lambda n: "\n" * n, lambda n: "".join(rng.choice("0123456789") for _ in range(n)),
            lambda n: ("".join(rng.choice("erabAB") for _ in range(rng.randint(2, 5))) * n)[:n]]
    tails = ["", "a", "B", "b c", "'s", "!", " x", "1", "\n", "́a"]
    prompts = []
    for _ in range(4000):
        parts = []
        for _ in range(rng.randint(1, 4)):
            parts += [rng.choice(heads), rng.choice(runs)(rng.choice([1, 2, 3, 5, 21, 22, 40, 70, 150, 260, 700, 1805, 5000])), rng.choice(tails)]
        prompts.append("".join(parts).encode())
    data, offs = pack(prompts)
    want_ids, want_off, want_counts = oracle.encode_batch([oracle_vocabs[pat]], [pat], data, offs, nthreads=os.cpu_count())
    r = encode(plug, ctx, SLOT_NAMES[pat], data, offs)
    assert np.array_equal(r.offsets, want_off)
    assert np.array_equal(r.ids, want_ids)
    assert np.array_equal(r.counts, want_counts)


def test_decode_batch_is_the_inverse_of_encode(plug, ctx, tekken_bytes):
    """decode = concatenation of the tokens' bytes (tiktoken decode_bytes).  Checked against
    the rank file directly (token bytes by id) and as a round trip of encode over fuzz prompts, two vocabularies in one batch"""
    from cfbpe import plugin as P, _native as N
    toks = [base64.b64decode(l.split()[0]) for l in tekken_bytes.splitlines() if l.strip()]
    prompts = [s.encode() for s in fuzzgen.fuzz_strings(4242, 3000, max_atoms=40)] + [b"", b"a", "中文".encode() * 700]
    data, offs = pack(prompts)
    vocabs = [P.VocabRef(SLOT_NAMES[i % 2]) for i in range(len(prompts))]
    enc = plug.encode_batch(ctx, P.EncodeBatchRequest(P.VocabRef(SLOT_NAMES[0]), data, offs, vocabs_per_prompt=vocabs))
    dec = plug.decode_batch(ctx, P.DecodeBatchRequest(P.VocabRef(SLOT_NAMES[0]), enc.ids, enc.offsets, vocabs_per_prompt=vocabs))
    assert np.array_equal(dec.offsets, offs)
    assert bytes(dec.bytes) == bytes(data)
    # ids straight from the rank file, not only what encode produces
    rng = np.random.default_rng(5)
    ids = rng.integers(0, 100256, size=50000, dtype=np.uint32)
    ioffs = np.array([0, 1, 1, 777, 30000, 50000], dtype=np.uint64)
    dec = plug.decode_batch(ctx, P.DecodeBatchRequest(P.VocabRef(SLOT_NAMES[0]), ids, ioffs))
    want = [b"".join(toks[int(t)] for t in ids[int(a):int(b)]) for a, b in zip(ioffs[:-1], ioffs[1:])]
    assert bytes(dec.bytes) == b"".join(want)
    assert dec.offsets.tolist() == [0] + list(np.cumsum([len(w) for w in want]))
    # errors: an id outside the vocabulary, an output buffer that is too small
    with pytest.raises(P.InvalidInput):
        plug.decode_batch(ctx, P.DecodeBatchRequest(P.VocabRef(SLOT_NAMES[0]), np.array([5, 100256], dtype=np.uint32), np.array([0, 2], dtype=np.uint64)))
    with pytest.raises(N.NativeError) as ei:
        plug.ctx.decode_batch(ids, ioffs, None, out_cap=10)
    assert ei.value.code == N.ENOSPC


def test_pipelined_host_path_matches_oracle(oracle_vocabs, tekken_bytes, monkeypatch):
    """force the pipelined (sub-batched, 3-stream) host path with tiny sub-batches: many chunk seams, chained token ranks"""
    from cfbpe import _native as N
    from oracle import oracle
    monkeypatch.setenv("CFBPE_PIPE_CHUNK_BYTES", "20000")
    monkeypatch.setenv("CFBPE_PIPE_MIN_BYTES", "1")
    c = N.Context(0, 8 << 20, 1 << 15)
    c.vocab_load(0, tekken_bytes, N.FORMAT_TIKTOKEN, 0, 100256)
    c.vocab_load(1, tekken_bytes, N.FORMAT_TIKTOKEN, 1, 150000)
    prompts = [s.encode() for s in fuzzgen.fuzz_strings(4321, 12000, max_atoms=40) + fuzzgen.long_runs(8)] + [b"", b"", b"x" * 30000, b""]
    data, offs = pack(prompts)
    vid = (np.arange(len(prompts)) % 2).astype(np.uint8)
    want_ids, want_off, want_counts = oracle.encode_batch([oracle_vocabs[0], oracle_vocabs[1]], [0, 1], data, offs, vocab_ids=vid,
                                                          nthreads=os.cpu_count())
    ids, off, counts = c.encode_batch(data, offs, vid)
    assert np.array_equal(off, want_off)
    assert np.array_equal(ids, want_ids)
    assert np.array_equal(counts, want_counts)
    assert np.array_equal(c.count_batch(data, offs, vid), want_counts)
    # ENOSPC and malformed UTF-8 through the pipelined path
    small = np.zeros(10, dtype=np.uint32)
    with pytest.raises(N.NativeError) as ei:
        c.encode_batch(data, offs, vid, small)
    assert ei.value.code == N.ENOSPC
    bad = [p for p in prompts[:3000]] + [b"\xff\xfe"] + [p for p in prompts[3000:6000]]
    d2, o2 = pack(bad)
    with pytest.raises(N.NativeError) as ei:
        c.encode_batch(d2, o2)
    assert ei.value.code == N.EILSEQ
    ids, off, counts = c.encode_batch(data, offs, vid)      # still healthy
    assert np.array_equal(ids, want_ids)
    c.close()


def test_async_device_calls_and_buffer_reuse(plug, ctx, oracle_vocabs):
    """back-to-back asynchronous device calls on one stream (what bench.py times), then sizes going down and up again:
    the workspace is reused, results must not leak from one call into the next"""
    import torch
    from oracle import oracle
    dev = torch.device("cuda:0")
    stream = torch.cuda.current_stream().cuda_stream
    batches = []
    for seed, n in [(1, 3000), (2, 40), (3, 9000), (4, 1)]:
        prompts = [s.encode() for s in fuzzgen.fuzz_strings(seed, n, max_atoms=50)]
        data, offs = pack(prompts)
        want_ids, want_off, _ = oracle.encode_batch([oracle_vocabs[0]], [0], data, offs, nthreads=os.cpu_count())
        total = int(offs[-1])
        d_bytes = torch.zeros(total + 64, dtype=torch.uint8, device=dev)
        d_bytes[:total] = torch.from_numpy(data).to(dev)
        bufs = dict(d_bytes=d_bytes, d_offs=torch.from_numpy(offs.astype(np.int64)).to(dev),
                    d_ids=torch.full((total + 1,), -1, dtype=torch.int32, device=dev),
                    d_out=torch.zeros(len(prompts) + 1, dtype=torch.int64, device=dev),
                    d_cnt=torch.zeros(max(len(prompts), 1), dtype=torch.int32, device=dev))
        batches.append((prompts, total, bufs, want_ids, want_off))
    for _ in range(2):
        for prompts, total, b, want_ids, want_off in batches:        # enqueue everything without syncing in between
            plug.ctx.encode_batch_device(len(prompts), b["d_bytes"].data_ptr(), total, b["d_offs"].data_ptr(), None, b["d_ids"].data_ptr(),
                                         b["d_ids"].numel(), b["d_out"].data_ptr(), b["d_cnt"].data_ptr(), stream, sync=False)
        plug.ctx.device_status(stream)
        for prompts, total, b, want_ids, want_off in batches:
            off = b["d_out"].cpu().numpy().astype(np.uint64)
            assert np.array_equal(off, want_off)
            assert np.array_equal(b["d_ids"][:int(off[-1])].cpu().numpy().view(np.uint32), want_ids)


def test_concurrent_calls_on_one_context(oracle_vocabs, tekken_bytes):
    """"safe to call concurrently from several host threads (internal stream pool ...)".
    (1) Two threads encode different batches at the same time on one context, both bit-exact -- with two workspaces (the calls
        run side by side) and with one (the calls queue).
    (2) Calls do not wait for each other: while one thread is inside a 134 MB call, a 64 KiB call from another thread returns
        BEFORE the big one does.  Behind one mutex (round 1) it could not: it would start only when the big call ended.
        (Equal-sized calls are timed too and printed, not asserted: a call that fills the GPU, or one that is all launch
        overhead, gains nothing from running beside its twin -- DESIGN.md section 4.)"""
    import threading, time
    from cfbpe import _native as N
    from cfbpe import workload as W
    from oracle import oracle
    batches = []
    for seed in (11, 12):
        prompts = [s.encode() for s in fuzzgen.fuzz_strings(seed, 30000, max_atoms=60)]
        data, offs = pack(prompts)
        want = oracle.encode_batch([oracle_vocabs[0]], [0], data, offs, nthreads=os.cpu_count())
        batches.append((data, offs, want))
    big_data, big_offs, _, _ = W.make_config(3, 1.0)
    for n_ws in (2, 1):
        c = N.Context(0, len(big_data) + 4096, 1 << 16, n_workspaces=n_ws)
        c.vocab_load(0, tekken_bytes, N.FORMAT_TIKTOKEN, 0, 100256)
        pinned = []
        for data, offs, _ in batches:      # pinned buffers: the copies of the two calls overlap too
            hb = c.pinned(len(data), np.uint8); hb.array[:] = data
            pinned.append((hb, c.pinned(len(data) + 1, np.uint32), c.pinned(len(offs), np.uint64), c.pinned(len(offs), np.uint32)))
        out, errs = {}, []

        def call(i):
            try:
                hb, hi, ho, hc = pinned[i]
                out[i] = c.encode_batch(hb.array, batches[i][1], None, hi.array, ho.array, hc.array)
            except Exception as e:   # noqa: BLE001
                errs.append(e)
        for i in (0, 1):
            call(i)                                   # warm-up, one after the other
        t_one, t_both = 1e9, 1e9
        for _ in range(5):
            for i in (0, 1):
                t0 = time.perf_counter(); call(i); t_one = min(t_one, time.perf_counter() - t0)
        for _ in range(5):
            th = [threading.Thread(target=call, args=(i,)) for i in (0, 1)]
            t0 = time.perf_counter()
            for t in th: t.start()
            for t in th: t.join()
            t_both = min(t_both, time.perf_counter() - t0)
        print("concurrent calls, %d workspace(s): one call %.3f ms, two at once %.3f ms (%.2fx), %d bytes each" % (n_ws, 1e3 * t_one, 1e3 * t_both, t_both / t_one, len(batches[0][0])))
        assert not errs, errs
        for i in (0, 1):
            ids, off, counts = out[i]
            assert np.array_equal(off, batches[i][2][1]) and np.array_equal(ids, batches[i][2][0]) and np.array_equal(counts, batches[i][2][2])

        # (2) a small call beside a big one
        gb = c.pinned(len(big_data), np.uint8); gb.array[:] = big_data
        g_out = (c.pinned(len(big_data) + 1, np.uint32), c.pinned(len(big_offs), np.uint64), c.pinned(len(big_offs), np.uint32))
        small_n = int(np.searchsorted(batches[0][1], 65536))        # the first ~64 KiB of batch 0
        s_offs = np.ascontiguousarray(batches[0][1][:small_n + 1])
        s_want_counts = batches[0][2][2][:small_n]
        big_ref = c.encode_batch(gb.array, big_offs, None, *[x.array for x in g_out])
        big_ref = (big_ref[0].copy(), big_ref[1].copy(), big_ref[2].copy())
        t0 = time.perf_counter(); c.count_batch(pinned[0][0].array, s_offs, None); t_small = time.perf_counter() - t0
        overtook = 0
        stamps = []
        for attempt in range(4):
            done = {}

            def big():
                try:
                    r = c.encode_batch(gb.array, big_offs, None, *[x.array for x in g_out])
                    done["t"] = time.perf_counter(); done["r"] = r
                except Exception as e:   # noqa: BLE001
                    errs.append(e)
            th = threading.Thread(target=big)
            t0 = time.perf_counter()
            th.start()
            time.sleep(0.001)                       # the big call is in flight (it takes > 5 ms)
            t_b0 = time.perf_counter()
            s_counts = c.count_batch(pinned[0][0].array, s_offs, None)
            t_b1 = time.perf_counter()
            th.join()
            assert not errs, errs
            assert np.array_equal(s_counts, s_want_counts)
            assert all(np.array_equal(x, y) for x, y in zip(done["r"], big_ref))
            stamps.append((1e3 * (t_b0 - t0), 1e3 * (t_b1 - t0), 1e3 * (done["t"] - t0)))
            overtook += t_b1 < done["t"]
        print("small call beside a 134 MB call, %d workspace(s): alone %.3f ms; (start, end) of the small call and end of the big one, ms: %s"
              % (n_ws, 1e3 * t_small, ", ".join("(%.2f, %.2f | %.2f)" % s for s in stamps)))
        if n_ws == 2:
            assert overtook >= 1, stamps          # with its own workspace the small call never waits for the big one to finish
        c.close()


def test_error_message_is_per_thread(plug, ctx):
    """the last error is kept per calling thread: a failing call on one thread does not clobber another thread's message"""
    import threading
    from cfbpe import _native as N
    msgs = {}

    def bad():
        try:
            plug.ctx.encode_batch(*pack([b"bad \xff"]))
        except N.NativeError as e:
            msgs["bad"] = str(e)
    t = threading.Thread(target=bad); t.start(); t.join()
    assert "UTF-8" in msgs["bad"]
    ids, off, counts = plug.ctx.encode_batch(*pack([b"fine"]))
    assert len(ids) > 0


@pytest.mark.skipif("__import__('torch').cuda.device_count() < 2")
@pytest.mark.parametrize("mode", ["shards", "round_robin"])
def test_multi_device_context_shards_a_batch(oracle_vocabs, tekken_bytes, mode, monkeypatch):
    """cfbpe_create(cfg: devices[], n_devices): one context over the GPUs of the box returns exactly what one device returns.
    "shards": one contiguous range of prompts a device, tables by ncclBroadcast, token totals by ncclAllGather, offsets rebased
    on the devices (batches a device cannot hold whole, or devices without peer access).  "round_robin": the sub-batches of ONE
    pipelined call go round the devices, the token-rank chain crosses NVLink peer memory (here forced onto a small batch: 40
    sub-batches of 64 KiB)."""
    import torch
    from cfbpe import _native as N
    from oracle import oracle
    ndev = min(torch.cuda.device_count(), 8)
    if mode == "round_robin":
        monkeypatch.setenv("CFBPE_PIPE_MIN_BYTES", "1")
        monkeypatch.setenv("CFBPE_PIPE_CHUNK_BYTES", str(64 << 10))
    else:
        monkeypatch.setenv("CFBPE_NO_PEER", "1")
    prompts = [s.encode() for s in fuzzgen.fuzz_strings(77, 40000, max_atoms=60) + fuzzgen.long_runs(5)] + [b"", b"x", b""]
    data, offs = pack(prompts)
    vid = (np.arange(len(prompts)) % 2).astype(np.uint8)
    want_ids, want_off, want_counts = oracle.encode_batch([oracle_vocabs[0], oracle_vocabs[3]], [0, 3], data, offs, vocab_ids=vid, nthreads=os.cpu_count())
    torch.cuda.set_device(0)
    c = N.Context(0, 64 << 20, 1 << 17, devices=list(range(ndev)))
    c.vocab_load(0, tekken_bytes, N.FORMAT_TIKTOKEN, 0, 100256)
    c.vocab_load(1, tekken_bytes, N.FORMAT_TIKTOKEN, 3, 130072)
    assert torch.cuda.current_device() == 0        # the library leaves the caller's current device alone
    for _ in range(2):
        ids, off, counts = c.encode_batch(data, offs, vid)
        assert np.array_equal(off, want_off) and np.array_equal(ids, want_ids) and np.array_equal(counts, want_counts)
        assert np.array_equal(c.count_batch(data, offs, vid), want_counts)
    small = np.zeros(10, dtype=np.uint32)
    with pytest.raises(N.NativeError) as ei:
        c.encode_batch(data, offs, vid, small)
    assert ei.value.code == N.ENOSPC
    bad = prompts[:1000] + [b"\xff\xfe"] + prompts[1000:]
    with pytest.raises(N.NativeError) as ei:
        c.encode_batch(*pack(bad))
    assert ei.value.code == N.EILSEQ
    ids, off, counts = c.encode_batch(*pack([b"one prompt only"]))        # fewer prompts than devices: one device does it
    assert np.array_equal(ids, oracle_vocabs[0].encode(0, b"one prompt only"))
    ids, off, counts = c.encode_batch(data, offs, vid)                    # and the context still works after the failures
    assert np.array_equal(ids, want_ids)
    c.close()
    assert torch.cuda.current_device() == 0
