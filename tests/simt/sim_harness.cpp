// sim_harness.cpp -- runs the product's CUDA kernels on the CPU SIMT emulator (tests/simt/cusim.h).
// TEST INFRASTRUCTURE: built by tests/simt/build.py into tests/simt/_build/libcfbpe_sim.so and
// loaded only by the non-GPU tests.  The table builder (csrc/vocab.cpp) and the kernels
// (csrc/bpe_kernels.cuh, csrc/pipeline.cuh) are the product sources, compiled unchanged.
#include "cusim.h"

#define CFBPE_LAUNCH(kernel, grid, block, stream, ...) cusim::launch((grid), (block), [&] { kernel(__VA_ARGS__); })
#define CFBPE_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...) cusim::launch((grid), (block), [&] { kernel(__VA_ARGS__); })
#define CFBPE_ZERO(ptr, bytes, stream) std::memset((ptr), 0, (bytes))
#define CFBPE_MARK(prof, idx, stream, begin) ((void)0)
#define CFBPE_FORK(main, aux, ev) ((void)0)
#define CFBPE_JOIN(main, aux, ev) ((void)0)

#include <string>
#include <vector>

#include "../../cyberfabric-core_b200/csrc/pipeline.cuh"
#include "../../cyberfabric-core_b200/csrc/pretok_ctx.h"
#include "../../cyberfabric-core_b200/csrc/subbatch.h"
#include "../../cyberfabric-core_b200/csrc/unicode_tables.h"
#include "../../cyberfabric-core_b200/csrc/vocab.h"
#include "../../include/cfbpe.h"

using namespace cfbpe;

struct SimVocab {
    std::vector<uint8_t> blob;
    TablesHeader hdr;
};

static UcTables uc_tables() {
    static uint16_t fsm[kNumPatterns * kPretokTableSize];
    static uint8_t ascii[128];
    static SplitTablesHost st;
    static bool init = false;
    if (!init) { build_pretok_tables(fsm); build_ascii_classes(ascii); build_split_tables(&st); init = true; }
    return UcTables{cfbpe_uc_stage1, cfbpe_uc_stage2, ascii, fsm, st.cls256, st.fsm16, st.ctx16, st.prod, st.prod_info, st.prod_skip, st.prod_start};
}
extern "C" __attribute__((visibility("default"))) uint32_t sim_ctx_count(uint32_t cased) { static SplitTablesHost st; build_split_tables(&st); return st.n_ctx[cased & 1]; }
extern "C" __attribute__((visibility("default"))) uint32_t sim_prod_count(uint32_t pat) { static SplitTablesHost st; build_split_tables(&st); return st.n_prod[pat & 3]; }

extern "C" {

__attribute__((visibility("default"))) void* sim_vocab_build(const uint8_t* file, size_t len, uint32_t format,
                                                             uint32_t pattern, uint32_t max_ranks, char* err, size_t errcap) {
    std::vector<std::string> toks;
    std::string e;
    int rc = (format == CFBPE_FORMAT_TEKKEN_JSON) ? parse_tekken_json(file, len, max_ranks, toks, e)
                                                  : parse_tiktoken(file, len, max_ranks, toks, e);
    SimVocab* v = nullptr;
    if (rc == 0) {
        v = new SimVocab();
        rc = build_tables(toks, pattern, v->blob, e);
        if (rc == 0) rc = validate_tables(v->blob.data(), v->blob.size(), e);
        if (rc != 0) { delete v; v = nullptr; }
        else std::memcpy(&v->hdr, v->blob.data(), sizeof(TablesHeader));
    }
    if (!v && err && errcap) { std::snprintf(err, errcap, "%s", e.c_str()); }
    return v;
}
// the packed tables of a built vocabulary (what cfbpe_vocab_export hands out) and the check cfbpe_vocab_import runs on a blob
__attribute__((visibility("default"))) uint64_t sim_vocab_blob(void* vp, uint8_t* out, uint64_t cap) {
    SimVocab* v = static_cast<SimVocab*>(vp);
    if (out && cap >= v->blob.size()) std::memcpy(out, v->blob.data(), v->blob.size());
    return v->blob.size();
}
__attribute__((visibility("default"))) int sim_validate_blob(const uint8_t* blob, uint64_t size, char* err, size_t errcap) {
    std::string e;
    const int rc = validate_tables(blob, size, e);
    if (err && errcap) std::snprintf(err, errcap, "%s", e.c_str());
    return rc;
}
__attribute__((visibility("default"))) unsigned long long sim_dbg_counter(uint32_t i, int reset) {
    unsigned long long* c = cfbpe::dbg_counters();
    const unsigned long long v = c[i & 15];
    if (reset) c[i & 15] = 0;
    return v;
}
__attribute__((visibility("default"))) void sim_vocab_free(void* v) { delete static_cast<SimVocab*>(v); }
__attribute__((visibility("default"))) void sim_vocab_info(void* vp, cfbpe_vocab_info* out) {
    SimVocab* v = static_cast<SimVocab*>(vp);
    out->n_ranks = v->hdr.n_ranks; out->pattern_id = v->hdr.pattern_id; out->max_token_len = v->hdr.max_token_len;
    out->n_pair_entries = v->hdr.n_pair_entries; out->table_bytes = v->hdr.total_bytes;
}
// host-side lookups through the packed tables (table-builder tests)
__attribute__((visibility("default"))) uint32_t sim_piece_lookup(void* vp, const uint8_t* p, uint32_t n) {
    SimVocab* v = static_cast<SimVocab*>(vp);
    std::vector<uint8_t> tmp(p, p + n); tmp.resize(n + 16);
    return piece_lookup(make_view(v->blob.data(), v->hdr), tmp.data(), n);
}
__attribute__((visibility("default"))) uint32_t sim_pair_lookup(void* vp, uint32_t l, uint32_t r) {
    SimVocab* v = static_cast<SimVocab*>(vp);
    return pair_lookup(make_view(v->blob.data(), v->hdr), l, r);
}

// K1 only: piece-start bits (n_words+2 words) for a packed batch; patterns[] per vocab id
static unsigned long long fixups_seen = 0, resumes_seen = 0;
static int legacy_split = 0;
__attribute__((visibility("default"))) void sim_use_legacy_split(int on) { legacy_split = on; }
__attribute__((visibility("default"))) unsigned long long sim_split_fixups(int reset) { const unsigned long long v = fixups_seen; if (reset) fixups_seen = 0; return v; }
__attribute__((visibility("default"))) int sim_split(const uint32_t* patterns, uint32_t n_patterns, uint32_t n_prompts,
                                                     const uint8_t* bytes, const uint64_t* offsets, const uint8_t* vocab_ids,
                                                     uint32_t* piece_bits) {
    const uint64_t total = offsets[n_prompts];
    std::vector<uint8_t> padded(bytes, bytes + total); padded.resize(total + 64);
    BatchView b{padded.data(), offsets, vocab_ids, n_prompts, total};
    VocabSet vs{};
    for (uint32_t i = 0; i < n_patterns && i < kMaxVocabs; ++i) vs.v[i].pattern_id = patterns[i];
    DeviceStatus st{};
    const uint64_t nw = n_flag_words(total);
    std::memset(piece_bits, 0, (nw + 2) * 4);
    vs.loaded_mask = n_patterns >= 32 ? 0xFFFFFFFFu : ((1u << n_patterns) - 1u);
    if (total) {
        UcTables uc = uc_tables();
        std::vector<SplitFix> fix(total / 16 + 2);
        const uint32_t fix_cap = static_cast<uint32_t>(fix.size());
        if (legacy_split) {
            const uint64_t n_chunks = (total + kSplitChunk - 1) / kSplitChunk;
            cusim::launch(static_cast<unsigned>((n_chunks + 255) / 256), 256, [&] { pretok_split_kernel(b, vs, uc, piece_bits, &st, fix.data(), fix_cap); });
        } else {
            std::vector<uint32_t> pstart(nw + 2), bprompt((total >> kPromptBlockShift) + 2);
            cusim::launch(static_cast<unsigned>((static_cast<uint64_t>(n_prompts) + 1 + 255) / 256), 256, [&] { prompt_map_kernel(b, vs, pstart.data(), bprompt.data(), &st); });
            const uint64_t n_blocks16 = (total + 15) / 16;
            const uint32_t n_tiles = static_cast<uint32_t>((n_blocks16 + kSplitWarpOwned - 1) / kSplitWarpOwned);
            const uint32_t n_tabs = vocab_ids ? kNumPatterns : 1u;
            cusim::launch(n_tiles < 17 ? 1u : 2u, kSplitCta,      // fewer warps than tiles: the persistent loop runs
                          [&] { pretok_split16_kernel(b, vs, uc, pstart.data(), bprompt.data(), piece_bits, &st, fix.data(), fix_cap, n_tabs, n_tiles); });
        }
        cusim::launch(2u, 256, [&] { pretok_fixup_kernel(b, vs, uc, piece_bits, &st, fix.data(), fix_cap); });
        fixups_seen += st.fix_n;
        resumes_seen += cfbpe::dbg_counters()[3] + cfbpe::dbg_counters()[4];
    }
    return st.bad_utf8 ? CFBPE_EILSEQ : 0;
}

// the whole path (K1..K3) on host memory
__attribute__((visibility("default"))) int sim_encode_batch(void* const* vocabs, uint32_t n_vocabs, uint32_t n_prompts,
                                                            const uint8_t* bytes, const uint64_t* offsets, const uint8_t* vocab_ids,
                                                            uint32_t* out_ids, uint64_t out_cap, uint64_t* out_offsets,
                                                            uint32_t* out_counts, uint64_t* n_long_out) {
    const uint64_t total = offsets[n_prompts];
    std::vector<uint8_t> padded(bytes, bytes + total); padded.resize(total + 64);
    BatchView b{padded.data(), offsets, vocab_ids, n_prompts, total};
    VocabSet vs{};
    for (uint32_t i = 0; i < n_vocabs && i < kMaxVocabs; ++i) {
        SimVocab* v = static_cast<SimVocab*>(vocabs[i]);
        vs.v[i] = make_view(v->blob.data(), v->hdr);
    }
    const uint64_t nw = n_flag_words(total);
    const uint32_t nt = n_scan_tiles(total);
    std::vector<uint32_t> piece_bits(nw + 2), tok_bits(nw + 2), ids(total + 1, 0xDEADBEEF), rk(total + 1), nx(total + 1), pv(total + 1);
    std::vector<uint32_t> tile_counts(nt + 1);
    std::vector<uint64_t> tile_base(nt + 1);
    std::vector<LongPiece> ll(total / 32 + 1);
    DeviceStatus st{};
    std::vector<SplitFix> fix(total / 16 + 2);
    std::vector<uint64_t> miss[3];
    MissLists ml;
    for (uint32_t c = 0; c < 3; ++c) {
        miss[c].resize(miss_list_words(total, c, 1));
        ml.list[c] = miss[c].data();
        ml.cap[c] = static_cast<uint32_t>(miss[c].size());
    }
    std::vector<uint32_t> pstart(nw + 2), bprompt((total >> kPromptBlockShift) + 2);
    std::vector<uint32_t> by_piece(total + 1, 0xDEADBEEF), extras(total + 1, 0xDEADBEEF), tile_pieces((total >> 11) + 2);
    std::vector<uint64_t> piece_base((total >> 11) + 2);
    Workspace w{piece_bits.data(), tok_bits.data(), ids.data(), LongScratch{rk.data(), nx.data(), pv.data()},
                ll.data(), static_cast<uint32_t>(ll.size()), tile_counts.data(), tile_base.data(), &st, ml, fix.data(), static_cast<uint32_t>(fix.size()),
                DenseIds{by_piece.data(), extras.data(), static_cast<uint32_t>(extras.size()), tile_pieces.data(), piece_base.data()},
                pstart.data(), bprompt.data()};
    vs.loaded_mask = n_vocabs >= 32 ? 0xFFFFFFFFu : ((1u << n_vocabs) - 1u);
    int* prof = nullptr;
    enqueue_encode(b, vs, uc_tables(), w, out_ids, out_cap, out_offsets, out_counts, 4u, 0, 0, 0, 0, 0, 0, prof);
    if (n_long_out) *n_long_out = static_cast<uint64_t>(st.n_long) + st.n_big;
    if (st.bad_utf8) return CFBPE_EILSEQ;
    if (st.long_overflow || st.miss_overflow) return CFBPE_EIO;
    if (out_ids && st.n_tokens > out_cap) return CFBPE_ENOSPC;
    return 0;
}

// the sub-batch plan of a pipelined host call (host code of the product, csrc/subbatch.h)
__attribute__((visibility("default"))) int sim_plan_sub_batches(const uint64_t* offsets, uint32_t n, uint64_t chunk, int max_chunks, uint32_t* cut) {
    return plan_sub_batches(offsets, n, offsets[n], chunk, max_chunks, cut);
}

// the decode path (ids -> bytes) on host memory; returns the number of decoded bytes in out_offsets[n_seqs]
__attribute__((visibility("default"))) int sim_decode_batch(void* const* vocabs, uint32_t n_vocabs, uint32_t n_seqs, const uint32_t* ids,
                                                            const uint64_t* id_offsets, const uint8_t* vocab_ids, uint8_t* out, uint64_t out_cap,
                                                            uint64_t* out_offsets) {
    VocabSet vs{};
    for (uint32_t i = 0; i < n_vocabs && i < kMaxVocabs; ++i) {
        SimVocab* v = static_cast<SimVocab*>(vocabs[i]);
        vs.v[i] = make_view(v->blob.data(), v->hdr);
    }
    const uint64_t n_ids = id_offsets[n_seqs];
    DecodeView d{ids, id_offsets, vocab_ids, n_seqs, n_ids};
    const uint32_t n_tiles = static_cast<uint32_t>((n_ids + kDecodeTile - 1) / kDecodeTile);
    std::vector<uint32_t> lens(n_ids + 1), sums(n_tiles + 1);
    std::vector<uint64_t> base(n_tiles + 1);
    DeviceStatus st{};
    if (n_tiles) cusim::launch(n_tiles, 256, [&] { decode_len_kernel(d, vs, lens.data(), sums.data(), &st); });
    cusim::launch(1u, n_tiles ? 1024u : 32u, [&] { tile_scan_kernel(sums.data(), n_tiles, base.data(), &st, nullptr); });
    if (st.bad_utf8) return CFBPE_EINVAL;
    if (n_tiles) cusim::launch(n_tiles, 256, [&] { decode_copy_kernel(d, vs, lens.data(), base.data(), out, out_cap); });
    cusim::launch(static_cast<unsigned>((static_cast<uint64_t>(n_seqs) + 1 + 255) / 256), 256, [&] { decode_offsets_kernel(d, lens.data(), base.data(), out_offsets, &st); });
    return st.tok_end > out_cap ? CFBPE_ENOSPC : 0;
}

}  // extern "C"
