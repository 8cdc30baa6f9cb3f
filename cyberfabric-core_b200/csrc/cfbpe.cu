// cfbpe.cu -- libcfbpe.so: device context, vocab upload and the C ABI of include/cfbpe.h.
//
// Built for sm_100a only.  There is no CPU path in this library: every entry point that
// computes runs the kernels of bpe_kernels.cuh on the device or returns an error.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/cfbpe.h"

struct ProfEvents;
#define CFBPE_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#define CFBPE_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define CFBPE_ZERO(ptr, bytes, stream) cudaMemsetAsync((ptr), 0, (bytes), (stream))
#define CFBPE_MARK(prof, idx, stream, begin) prof_mark((prof), (idx), (stream), (begin))
static inline void prof_mark(ProfEvents* p, int idx, cudaStream_t s, bool begin);
// make `aux` wait for what is queued on `main` so far / make `main` wait for `aux`; no-ops when both are the same stream
#define CFBPE_FORK(main, aux, ev) do { if ((main) != (aux)) { cudaEventRecord((ev), (main)); cudaStreamWaitEvent((aux), (ev), 0); } } while (0)
#define CFBPE_JOIN(main, aux, ev) do { if ((main) != (aux)) { cudaEventRecord((ev), (aux)); cudaStreamWaitEvent((main), (ev), 0); } } while (0)

#include "pipeline.cuh"
#include "pretok_ctx.h"
#include "subbatch.h"
#include "unicode_tables.h"
#include "vocab.h"

using namespace cfbpe;

struct ProfEvents {
    cudaEvent_t ev[CFBPE_NUM_KERNELS][2];
    cudaEvent_t h2d[2], d2h[2], total[2];
    bool launched[CFBPE_NUM_KERNELS];
};
static inline void prof_mark(ProfEvents* p, int idx, cudaStream_t s, bool begin) {
    if (!p) return;
    cudaEventRecord(p->ev[idx][begin ? 0 : 1], s);
    p->launched[idx] = true;
}

constexpr int kMaxPipeChunks = 64;
constexpr int kSideStreams = 16;
#ifndef CFBPE_FRONT_STREAMS
#define CFBPE_FRONT_STREAMS 6
#endif
constexpr int kFrontStreams = CFBPE_FRONT_STREAMS;
constexpr uint64_t kPipeChunkBytes = 12ull << 20;   // largest sub-batch of a pipelined host call (the sizes ramp up to it and down again); measured: profiles/e2e_subbatch_sizes_r01t.jsonl
constexpr uint64_t kPipeMinBytes = 4ull << 20;      // smaller calls run as one shot (a 134 MB batch sharded over 8 GPUs is 16.8 MB a rank: it must still pipeline)

struct VocabSlot {
    bool loaded = false;
    uint8_t* d_blob = nullptr;
    uint64_t* d_hot = nullptr;       // the hot slice of the pair table (tables.h kHotRanks / kHotCap)
    std::vector<uint8_t> h_blob;
    TablesHeader hdr{};
};

struct cfbpe_ctx {
    int device = 0;
    std::mutex mu;
    std::string err;
    uint64_t max_bytes = 0;
    uint32_t max_prompts = 0;
    cudaStream_t stream = nullptr;       // compute
    cudaStream_t h2d_stream = nullptr;   // pipelined host calls: uploads run ahead of the kernels ...
    cudaStream_t d2h_stream = nullptr;   // ... and downloads trail them
    uint32_t* d_dec_sums = nullptr;      // decode: bytes per tile of kDecodeTile tokens ...
    uint64_t* d_dec_base = nullptr;      // ... and their exclusive scan
    cudaStream_t aux_stream = nullptr;   // the long-piece kernel runs here, next to the short-piece kernel
    cudaStream_t aux2_stream = nullptr;  // ... and the big-piece kernel here, next to both
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join2 = nullptr;
    cudaStream_t side2[kSideStreams] = {};       // pipelined host calls: the big-piece kernel of sub-batch k
    cudaEvent_t ev_list[kMaxPipeChunks] = {};
    cudaEvent_t ev_ws = nullptr;         // recorded at the end of an asynchronous device-path call: the workspace is busy until then
    bool ws_pending = false;             // ... and whether one is outstanding
    uint64_t dev_out_cap = 0;            // out_cap of the last device-path call (cfbpe_device_status reports ENOSPC against it)
    bool dev_want_ids = false;
    cudaEvent_t ev_scan[kMaxPipeChunks] = {};
    cudaStream_t front[kFrontStreams] = {};   // front streams 1.. of a pipelined host call (0 = stream)
    cudaStream_t side[kSideStreams] = {};  // long-piece tails + emit of sub-batch k overlap the front of k+1
    cudaEvent_t ev_front[kMaxPipeChunks] = {};
    cudaEvent_t ev_h2d[kMaxPipeChunks] = {};
    cudaEvent_t ev_done[kMaxPipeChunks] = {};
    cudaEvent_t ev_chain[kMaxPipeChunks] = {};
    cudaEvent_t (*trace)[6] = nullptr;           // CFBPE_PIPE_TRACE=1: timed events per sub-batch (h2d, split, short, long, back, d2h) + [nc][0] = start   // tile_scan of sub-batch k done: the next sub-batch's scan may read tok_end
    uint64_t pipe_chunk = kPipeChunkBytes, pipe_min = kPipeMinBytes;   // CFBPE_PIPE_CHUNK_BYTES / CFBPE_PIPE_MIN_BYTES override (tests)
    DeviceStatus* d_status_arr = nullptr; // one status per sub-batch
    DeviceStatus* h_status_arr = nullptr; // pinned
    uint64_t* h_offs_stage = nullptr;     // pinned: sub-batch-local offsets
    // inputs / outputs of the host API
    uint8_t* d_bytes = nullptr;
    uint64_t* d_offsets = nullptr;
    uint8_t* d_vocab_ids = nullptr;
    uint32_t* d_out_ids = nullptr;
    uint64_t* d_out_offsets = nullptr;
    uint32_t* d_out_counts = nullptr;
    Workspace ws{};
    DeviceStatus* h_status = nullptr;  // pinned
    uint8_t* d_uc1 = nullptr;
    uint8_t* d_uc2 = nullptr;
    uint8_t* d_ascii = nullptr;
    uint16_t* d_fsm = nullptr;
    uint8_t* d_split_tables = nullptr;   // SplitTablesHost: class bytes, 16-wide transition tables, context automaton (pretok_ctx.h)
    UcTables uc{};
    VocabSlot vocabs[CFBPE_MAX_VOCABS];
    VocabSet vs{};
    int sm_count = 148;
    bool profiling = false;
    ProfEvents prof{};
    bool prof_ready = false;
    cfbpe_profile last_profile{};
};

namespace {

int fail(cfbpe_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    return code;
}
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            return fail(ctx, CFBPE_EIO, std::string(#call) + ": " + cudaGetErrorString(e_));             \
        }                                                                                                \
    } while (0)

template <typename T>
cudaError_t dmalloc(T** p, uint64_t count) { return cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T)); }

int validate_batch(cfbpe_ctx* ctx, uint32_t n, const uint64_t* offsets, const uint8_t* vocab_ids, uint64_t* total_out) {
    if (n > ctx->max_prompts) return fail(ctx, CFBPE_EINVAL, "n_prompts exceeds max_prompts of this context");
    if (!offsets) return fail(ctx, CFBPE_EINVAL, "offsets is NULL");
    if (offsets[0] != 0) return fail(ctx, CFBPE_EINVAL, "offsets[0] must be 0");
    for (uint32_t i = 0; i < n; ++i)
        if (offsets[i + 1] < offsets[i]) return fail(ctx, CFBPE_EINVAL, "offsets are not monotonic at prompt " + std::to_string(i));
    if (offsets[n] > ctx->max_bytes) return fail(ctx, CFBPE_EINVAL, "batch exceeds max_batch_bytes of this context");
    if (vocab_ids) {
        for (uint32_t i = 0; i < n; ++i)
            if (vocab_ids[i] >= CFBPE_MAX_VOCABS || !ctx->vocabs[vocab_ids[i]].loaded)
                return fail(ctx, CFBPE_ENOENT, "prompt " + std::to_string(i) + " names a vocab that is not loaded");
    } else if (!ctx->vocabs[0].loaded) {
        return fail(ctx, CFBPE_ENOENT, "vocab 0 is not loaded");
    }
    *total_out = offsets[n];
    return CFBPE_OK;
}

int install_blob(cfbpe_ctx* ctx, uint32_t vocab_id, std::vector<uint8_t>&& blob) {
    VocabSlot& v = ctx->vocabs[vocab_id];
    uint8_t* d = nullptr;
    CK(cudaMalloc(reinterpret_cast<void**>(&d), blob.size()));
    cudaError_t e = cudaMemcpy(d, blob.data(), blob.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(d); return fail(ctx, CFBPE_EIO, std::string("table upload: ") + cudaGetErrorString(e)); }
    // the hot slice: the pair entries with a low merged id, re-hashed into a table small enough for shared memory
    uint64_t* d_hot = nullptr;
    {
        TablesHeader h;
        std::memcpy(&h, blob.data(), sizeof h);
        const uint64_t* pairs = reinterpret_cast<const uint64_t*>(blob.data() + h.off_pair);
        std::vector<uint64_t> hot(kHotCap, kPairEmpty);
        uint32_t n_hot = 0;
        for (uint32_t i = 0; i < h.cap_pair; ++i) {
            const uint64_t sl = pairs[i];
            if (sl == kPairEmpty || (static_cast<uint32_t>(sl) & kIdMask) >= kHotRanks || n_hot >= kHotCap * 3 / 4) continue;
            uint32_t hh = pair_hash(static_cast<uint32_t>(sl >> (2 * kIdBits)), static_cast<uint32_t>(sl >> kIdBits) & kIdMask) & (kHotCap - 1);
            while (hot[hh] != kPairEmpty) hh = (hh + 1) & (kHotCap - 1);
            hot[hh] = sl; ++n_hot;
        }
        if (cudaMalloc(reinterpret_cast<void**>(&d_hot), kHotCap * sizeof(uint64_t)) != cudaSuccess ||
            cudaMemcpy(d_hot, hot.data(), kHotCap * sizeof(uint64_t), cudaMemcpyHostToDevice) != cudaSuccess) {
            cudaFree(d); cudaFree(d_hot); return fail(ctx, CFBPE_EIO, "hot pair table upload failed");
        }
    }
    if (v.d_blob) { cudaDeviceSynchronize(); cudaFree(v.d_blob); cudaFree(v.d_hot); }   // kernels of a device-path call on any stream may still read the old tables
    v.d_blob = d;
    v.d_hot = d_hot;
    v.h_blob = std::move(blob);
    std::memcpy(&v.hdr, v.h_blob.data(), sizeof(TablesHeader));
    v.loaded = true;
    ctx->vs.v[vocab_id] = make_view(v.d_blob, v.hdr);
    ctx->vs.v[vocab_id].hot = v.d_hot;
    ctx->vs.loaded_mask |= 1u << vocab_id;
    // slots that are not loaded alias a loaded one: a bad vocabulary id handed in by a device-path caller is reported
    // (DeviceStatus::bad_vocab -> CFBPE_ENOENT) instead of dereferencing a null table
    for (uint32_t i = 0; i < CFBPE_MAX_VOCABS; ++i) if (!ctx->vocabs[i].loaded) ctx->vs.v[i] = ctx->vs.v[vocab_id];
    return CFBPE_OK;
}

void fill_profile(cfbpe_ctx* ctx, uint64_t n_bytes) {
    cfbpe_profile& p = ctx->last_profile;
    std::memset(&p, 0, sizeof p);
    for (int k = 0; k < CFBPE_NUM_KERNELS; ++k) {
        if (!ctx->prof.launched[k]) continue;
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ctx->prof.ev[k][0], ctx->prof.ev[k][1]) == cudaSuccess) p.kernel_ms[k] = ms;
        p.kernel_launches[k] = (k == K_EMIT) ? 2 : 1;   // emit_compact + prompt_offsets
    }
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->prof.h2d[0], ctx->prof.h2d[1]) == cudaSuccess) p.h2d_ms = ms;
    if (cudaEventElapsedTime(&ms, ctx->prof.d2h[0], ctx->prof.d2h[1]) == cudaSuccess) p.d2h_ms = ms;
    if (cudaEventElapsedTime(&ms, ctx->prof.total[0], ctx->prof.total[1]) == cudaSuccess) p.total_ms = ms;
    p.n_tokens = ctx->h_status->n_tokens;
    p.n_long_pieces = static_cast<uint64_t>(ctx->h_status->n_long) + ctx->h_status->n_big;
    p.n_bytes = n_bytes;
    p.n_long_bytes = ctx->h_status->long_bytes;
    p.n_long_tokens = ctx->h_status->long_tokens;
    p.n_miss_pieces = static_cast<uint64_t>(ctx->h_status->miss_n[0]) + ctx->h_status->miss_n[1] + ctx->h_status->miss_n[2];
    p.n_list_pieces = ctx->h_status->defer_n;
    p.n_list_parts = ctx->h_status->defer_parts;
    ctx->prof_ready = true;
}


// Pipelined host call: the batch is cut into sub-batches of ~kPipeChunkBytes on prompt boundaries; each is an
// independent encode pass on its own slice of the workspace.  Uploads (h2d_stream) run ahead of the kernels
// (stream), downloads (d2h_stream) trail them; token ranks are chained on the device (DeviceStatus::tok_end), so
// ids and offsets land at their final places.  The host only waits for each sub-batch's status to learn how many
// ids to fetch.
int run_host_pipelined(cfbpe_ctx* ctx, uint32_t n, const uint8_t* bytes, const uint64_t* offsets, const uint8_t* vocab_ids,
                       uint32_t* out_ids, uint64_t out_cap, uint64_t* out_offsets, uint32_t* out_counts, bool want_ids,
                       uint64_t total) {
    // ---- cut
    uint32_t cut[kMaxPipeChunks + 1];
    const int nc = plan_sub_batches(offsets, n, total, ctx->pipe_chunk, kMaxPipeChunks, cut);
    cudaStream_t cs = ctx->stream, hs = ctx->h2d_stream, ds = ctx->d2h_stream;
    // local offsets of every sub-batch, staged in pinned memory (sub-batch k occupies [p_k + k, p_{k+1} + k])
    for (int k = 0; k < nc; ++k) {
        const uint32_t p0 = cut[k], p1 = cut[k + 1];
        const uint64_t o0 = offsets[p0];
        uint64_t* dst = ctx->h_offs_stage + p0 + k;
        for (uint32_t i = p0; i <= p1; ++i) dst[i - p0] = offsets[i] - o0;
    }
    // ---- enqueue everything that does not depend on the host knowing a result
    const bool trace = getenv("CFBPE_PIPE_TRACE") != nullptr;
    if (trace && !ctx->trace) {
        ctx->trace = new cudaEvent_t[kMaxPipeChunks + 1][6];
        for (int k = 0; k <= kMaxPipeChunks; ++k) for (int j = 0; j < 6; ++j) cudaEventCreate(&ctx->trace[k][j]);
    }
    if (trace) CK(cudaEventRecord(ctx->trace[nc][0], hs));
    const auto host_t0 = std::chrono::steady_clock::now();
    auto host_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count(); };
    double host_enq[kMaxPipeChunks] = {}, host_dl[kMaxPipeChunks] = {};
    for (int k = 0; k < nc; ++k) {
        const uint32_t p0 = cut[k], p1 = cut[k + 1], nk = p1 - p0;
        const uint64_t o0 = offsets[p0], len = offsets[p1] - o0;
        // every sub-batch lands on a 16-byte boundary of the device buffer (K1 reads 16 bytes per lane with one load)
        uint8_t* const d_sub = ctx->d_bytes + ((o0 + 15) & ~15ull) + 16ull * k;
        if (len) CK(cudaMemcpyAsync(d_sub, bytes + o0, len, cudaMemcpyHostToDevice, hs));
        CK(cudaMemcpyAsync(ctx->d_offsets + p0 + k, ctx->h_offs_stage + p0 + k, (static_cast<uint64_t>(nk) + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, hs));
        if (vocab_ids && nk) CK(cudaMemcpyAsync(ctx->d_vocab_ids + p0, vocab_ids + p0, nk, cudaMemcpyHostToDevice, hs));
        CK(cudaEventRecord(ctx->ev_h2d[k], hs));
        if (trace) CK(cudaEventRecord(ctx->trace[k][0], hs));
        const int fk = k < kFrontStreams ? k : kFrontStreams - 1;
        cudaStream_t ck = fk ? ctx->front[fk] : cs;   // the short-piece kernels of sub-batch k: priority falls with k (earlier sub-batches finish, and download, first)
        Workspace w = ctx->ws;
        const uint64_t w0 = (o0 >> 5) + 4ull * k;
        w.piece_bits += w0; w.tok_bits += w0; w.pstart_bits += w0;
        w.block_prompt += (o0 >> kPromptBlockShift) + 2ull * k;
        w.dense.by_piece += o0; w.dense.extras += o0; w.dense.extras_cap = static_cast<uint32_t>(len + 1);
        w.dense.tile_pieces += (o0 >> 11) + 2ull * k; w.dense.piece_base += (o0 >> 11) + 2ull * k;
        w.ids_by_pos += o0; w.lscratch.rank += o0; w.lscratch.aux0 += o0; w.lscratch.aux1 += o0;
        w.long_list += (o0 >> 5) + k;
        w.long_cap = static_cast<uint32_t>(len / 32 + 1);
        const uint64_t t0 = (o0 >> 13) + 2ull * k;
        w.tile_counts += t0; w.tile_base += t0;
        w.status = ctx->d_status_arr + k;
        w.miss = slice_miss(ctx->ws.miss, o0, len, static_cast<uint32_t>(k));
        w.fix_list += (o0 >> 4) + 2ull * k;
        w.fix_cap = static_cast<uint32_t>(len / 16 + 2);
        BatchView b{d_sub, ctx->d_offsets + p0 + k, vocab_ids ? ctx->d_vocab_ids + p0 : nullptr, nk, len};
        // split, long pieces and the back stage run on a top-priority stream of their own: the long-piece kernels are a latency
        // chain that uses little of the machine, so they start as early as possible and the short-piece kernels fill the rest
        cudaStream_t ss = ctx->side[k % kSideStreams];
        CK(cudaStreamWaitEvent(ss, ctx->ev_h2d[k], 0));
        enqueue_split(b, ctx->vs, ctx->uc, w, ss, static_cast<ProfEvents*>(nullptr));
        CK(cudaEventRecord(ctx->ev_scan[k], ss));
        if (trace) CK(cudaEventRecord(ctx->trace[k][1], ss));
        CK(cudaStreamWaitEvent(ck, ctx->ev_scan[k], 0));
        cudaStream_t ss2 = ctx->side2[k % kSideStreams];
        CK(cudaStreamWaitEvent(ss2, ctx->ev_scan[k], 0));
        enqueue_list(b, ctx->vs, w, static_cast<uint32_t>(ctx->sm_count * 4), ss2, static_cast<ProfEvents*>(nullptr));   // the big pieces, beside everything else
        CK(cudaEventRecord(ctx->ev_list[k], ss2));
        enqueue_long(b, ctx->vs, w, static_cast<uint32_t>(ctx->sm_count * 4), ss, static_cast<ProfEvents*>(nullptr));   // tail overlaps what follows on cs
        if (trace) CK(cudaEventRecord(ctx->trace[k][3], ss));
        enqueue_short(b, ctx->vs, w, static_cast<uint32_t>(ctx->sm_count * 4), ck, static_cast<ProfEvents*>(nullptr));
        CK(cudaEventRecord(ctx->ev_front[k], ck));
        if (trace) CK(cudaEventRecord(ctx->trace[k][2], ck));
        CK(cudaStreamWaitEvent(ss, ctx->ev_front[k], 0));
        CK(cudaStreamWaitEvent(ss, ctx->ev_list[k], 0));
        enqueue_count(b, w, ss, static_cast<ProfEvents*>(nullptr));
        if (k) CK(cudaStreamWaitEvent(ss, ctx->ev_chain[k - 1], 0));    // token ranks chain through DeviceStatus::tok_end: only the scan waits
        enqueue_scan(b, w, ss, static_cast<ProfEvents*>(nullptr), k ? &ctx->d_status_arr[k - 1].tok_end : nullptr);
        CK(cudaEventRecord(ctx->ev_chain[k], ss));
        enqueue_emit(b, w, want_ids ? ctx->d_out_ids : nullptr, ctx->max_bytes, ctx->d_out_offsets + p0 + k, ctx->d_out_counts + p0,
                     ss, static_cast<ProfEvents*>(nullptr));
        CK(cudaGetLastError());
        status_publish_kernel<<<1, 64, 0, ss>>>(ctx->d_status_arr + k, ctx->h_status_arr + k);
        CK(cudaEventRecord(ctx->ev_done[k], ss));
        if (trace) { CK(cudaEventRecord(ctx->trace[k][4], ss)); host_enq[k] = host_ms(); }
    }
    // ---- trail the kernels with the downloads
    int err = CFBPE_OK;
    uint64_t tok_total = 0;
    for (int k = 0; k < nc; ++k) {
        CK(cudaEventSynchronize(ctx->ev_done[k]));
        const DeviceStatus st = ctx->h_status_arr[k];
        const uint32_t p0 = cut[k], p1 = cut[k + 1], nk = p1 - p0;
        if ((st.long_overflow || st.miss_overflow) && !err) err = fail(ctx, CFBPE_EIO, "internal: long-piece list overflow");
        if (st.bad_vocab && !err) err = fail(ctx, CFBPE_ENOENT, "a prompt names a vocabulary that is not loaded");
        if (st.bad_utf8 && !err) err = fail(ctx, CFBPE_EILSEQ, "a prompt holds malformed UTF-8");
        const uint64_t base = st.tok_end - st.n_tokens;
        tok_total = st.tok_end;
        if (err) continue;
        if (want_ids && st.tok_end <= out_cap && st.n_tokens)
            CK(cudaMemcpyAsync(out_ids + base, ctx->d_out_ids + base, st.n_tokens * sizeof(uint32_t), cudaMemcpyDeviceToHost, ds));
        if (out_offsets) CK(cudaMemcpyAsync(out_offsets + p0, ctx->d_out_offsets + p0 + k, (static_cast<uint64_t>(nk) + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, ds));
        if (out_counts && nk) CK(cudaMemcpyAsync(out_counts + p0, ctx->d_out_counts + p0, static_cast<uint64_t>(nk) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ds));
        if (trace) { CK(cudaEventRecord(ctx->trace[k][5], ds)); host_dl[k] = host_ms(); }
    }
    CK(cudaStreamSynchronize(ds));
    CK(cudaStreamSynchronize(cs));
    for (int k = 1; k < kFrontStreams; ++k) CK(cudaStreamSynchronize(ctx->front[k]));
    for (int k = 0; k < kSideStreams; ++k) CK(cudaStreamSynchronize(ctx->side[k]));
    for (int k = 0; k < kSideStreams; ++k) CK(cudaStreamSynchronize(ctx->side2[k]));
    if (trace && !err) {
        fprintf(stderr, "pipe trace (ms since the first upload was enqueued): sub-batch bytes | h2d split short long_end back d2h\n");
        for (int k = 0; k < nc; ++k) {
            float t[6];
            for (int j = 0; j < 6; ++j) cudaEventElapsedTime(&t[j], ctx->trace[nc][0], ctx->trace[k][j]);
            fprintf(stderr, "  %2d %9llu | %6.2f %6.2f %6.2f %6.2f %6.2f %6.2f | host: enqueued %.2f download issued %.2f\n", k,
                    static_cast<unsigned long long>(offsets[cut[k + 1]] - offsets[cut[k]]), t[0], t[1], t[2], t[3], t[4], t[5], host_enq[k], host_dl[k]);
        }
    }
    if (err) return err;
    if (want_ids && tok_total > out_cap) {
        if (out_offsets) out_offsets[n] = tok_total;
        return fail(ctx, CFBPE_ENOSPC, "out_cap too small: need " + std::to_string(tok_total) + " ids");
    }
    return CFBPE_OK;
}

// shared body of encode_batch / count_batch (host buffers)
int run_host(cfbpe_ctx* ctx, uint32_t n, const uint8_t* bytes, const uint64_t* offsets, const uint8_t* vocab_ids,
             uint32_t* out_ids, uint64_t out_cap, uint64_t* out_offsets, uint32_t* out_counts, bool want_ids) {
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->err.clear();
    uint64_t total = 0;
    int rc = validate_batch(ctx, n, offsets, vocab_ids, &total);
    if (rc) return rc;
    if (total && !bytes) return fail(ctx, CFBPE_EINVAL, "bytes is NULL");
    if (want_ids && (!out_offsets || (!out_ids && out_cap))) return fail(ctx, CFBPE_EINVAL, "output pointer is NULL");
    CK(cudaSetDevice(ctx->device));
    if (ctx->ws_pending) { CK(cudaEventSynchronize(ctx->ev_ws)); ctx->ws_pending = false; }   // an asynchronous device-path call still owns the workspace
    if (!ctx->profiling && total >= ctx->pipe_min && n >= 2)
        return run_host_pipelined(ctx, n, bytes, offsets, vocab_ids, out_ids, out_cap, out_offsets, out_counts, want_ids, total);
    cudaStream_t s = ctx->stream;
    ProfEvents* prof = ctx->profiling ? &ctx->prof : nullptr;
    if (prof) { std::memset(prof->launched, 0, sizeof prof->launched); cudaEventRecord(prof->total[0], s); cudaEventRecord(prof->h2d[0], s); }
    if (total) CK(cudaMemcpyAsync(ctx->d_bytes, bytes, total, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->d_offsets, offsets, (static_cast<uint64_t>(n) + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    if (vocab_ids && n) CK(cudaMemcpyAsync(ctx->d_vocab_ids, vocab_ids, n, cudaMemcpyHostToDevice, s));
    if (prof) cudaEventRecord(prof->h2d[1], s);

    BatchView b{ctx->d_bytes, ctx->d_offsets, vocab_ids ? ctx->d_vocab_ids : nullptr, n, total};
    enqueue_encode(b, ctx->vs, ctx->uc, ctx->ws, want_ids ? ctx->d_out_ids : nullptr, ctx->max_bytes, ctx->d_out_offsets,
                   ctx->d_out_counts, static_cast<uint32_t>(ctx->sm_count * 4), s, prof ? s : ctx->aux_stream, prof ? s : ctx->aux2_stream,
                   ctx->ev_fork, ctx->ev_join, ctx->ev_join2, prof);
    CK(cudaGetLastError());
    if (prof) cudaEventRecord(prof->d2h[0], s);
    CK(cudaMemcpyAsync(ctx->h_status, ctx->ws.status, sizeof(DeviceStatus), cudaMemcpyDeviceToHost, s));
    if (out_offsets) CK(cudaMemcpyAsync(out_offsets, ctx->d_out_offsets, (static_cast<uint64_t>(n) + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    if (out_counts && n) CK(cudaMemcpyAsync(out_counts, ctx->d_out_counts, static_cast<uint64_t>(n) * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const DeviceStatus st = *ctx->h_status;
    if (st.long_overflow || st.miss_overflow) return fail(ctx, CFBPE_EIO, "internal: long-piece list overflow");
    if (st.bad_vocab) return fail(ctx, CFBPE_ENOENT, "a prompt names a vocabulary that is not loaded");
    if (st.bad_utf8) return fail(ctx, CFBPE_EILSEQ, "a prompt holds malformed UTF-8");
    if (want_ids) {
        if (st.n_tokens > out_cap) {
            if (out_offsets) out_offsets[n] = st.n_tokens;
            return fail(ctx, CFBPE_ENOSPC, "out_cap too small: need " + std::to_string(st.n_tokens) + " ids");
        }
        if (st.n_tokens) CK(cudaMemcpyAsync(out_ids, ctx->d_out_ids, st.n_tokens * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    }
    if (prof) { cudaEventRecord(prof->d2h[1], s); cudaEventRecord(prof->total[1], s); }
    CK(cudaStreamSynchronize(s));
    if (prof) fill_profile(ctx, total);
    return CFBPE_OK;
}

}  // namespace

extern "C" {

int cfbpe_abi_version(void) { return static_cast<int>(CFBPE_ABI_VERSION); }

#ifndef CFBPE_SRC_HASH
#define CFBPE_SRC_HASH "unknown"
#endif
const char* cfbpe_build_id(void) { return CFBPE_SRC_HASH; }

int cfbpe_create(const cfbpe_config* cfg, cfbpe_ctx** out) {
    if (!cfg || !out || cfg->struct_size < sizeof(cfbpe_config)) return CFBPE_EINVAL;
    *out = nullptr;
    // (a pipelined host call keeps ~20 streams busy: hosts should export CUDA_DEVICE_MAX_CONNECTIONS=32 before CUDA
    //  initialises -- INTEGRATION.md; the library does not touch the process environment)
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return CFBPE_ENODEV;
    if (cfg->device < 0 || cfg->device >= ndev) return CFBPE_ENODEV;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return CFBPE_ENODEV;
    if (prop.major != 10) return CFBPE_ENODEV;  // sm_100a SASS only
    if (cudaSetDevice(cfg->device) != cudaSuccess) return CFBPE_ENODEV;

    cfbpe_ctx* ctx = new (std::nothrow) cfbpe_ctx();
    if (!ctx) return CFBPE_ENOMEM;
    ctx->device = cfg->device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->max_bytes = cfg->max_batch_bytes ? cfg->max_batch_bytes : (256ull << 20);
    ctx->max_prompts = cfg->max_prompts ? cfg->max_prompts : (1u << 20);
    if (ctx->max_bytes >= (1ull << 32) - 4096) { delete ctx; return CFBPE_EINVAL; }   // byte positions inside a batch are 32-bit in the work lists
    const uint64_t mb = ctx->max_bytes, mp = ctx->max_prompts;
    const uint64_t nw = n_flag_words(mb) + 4 + 4 * kMaxPipeChunks;      // + per-sub-batch slack of a pipelined call
    const uint64_t nt = n_scan_tiles(mb) + 1 + 2 * kMaxPipeChunks;
    int prio_lo0 = 0, prio_hi0 = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo0, &prio_hi0);
    bool ok = cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_hi0) == cudaSuccess;   // front stream of sub-batch 0
    ok = ok && cudaFuncSetAttribute(bpe_list_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kListSmemBytes)) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(bpe_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kHotCap * 8)) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(pretok_split16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kNumPatterns * kProdTableBytes)) == cudaSuccess;
    ok = ok && dmalloc(&ctx->d_bytes, mb + 256 + 16 * (kMaxPipeChunks + 1)) == cudaSuccess;
    ok = ok && dmalloc(&ctx->d_offsets, mp + 1 + kMaxPipeChunks) == cudaSuccess;
    ok = ok && dmalloc(&ctx->d_vocab_ids, mp + 1) == cudaSuccess;
    ok = ok && dmalloc(&ctx->d_out_ids, mb + 1) == cudaSuccess;
    ok = ok && dmalloc(&ctx->d_out_offsets, mp + 1 + kMaxPipeChunks) == cudaSuccess;
    ok = ok && dmalloc(&ctx->d_out_counts, mp + 1) == cudaSuccess;
    ok = ok && dmalloc(&ctx->ws.piece_bits, nw) == cudaSuccess;
    ok = ok && dmalloc(&ctx->ws.tok_bits, nw) == cudaSuccess;
    ok = ok && dmalloc(&ctx->ws.ids_by_pos, mb + 1) == cudaSuccess;
    ok = ok && dmalloc(&ctx->ws.lscratch.rank, mb + 1) == cudaSuccess;
    ok = ok && dmalloc(&ctx->ws.lscratch.aux0, mb + 1) == cudaSuccess;
    ok = ok && dmalloc(&ctx->ws.lscratch.aux1, mb + 1) == cudaSuccess;
    ctx->ws.long_cap = static_cast<uint32_t>(mb / 32 + 1 + kMaxPipeChunks);   // a long piece holds more than 32 bytes
    ok = ok && dmalloc(&ctx->ws.long_list, ctx->ws.long_cap) == cudaSuccess;
    for (uint32_t c = 0; c < 3; ++c) {
        const uint64_t words = miss_list_words(mb, c, kMaxPipeChunks);      // 64-bit entries
        ok = ok && dmalloc(&ctx->ws.miss.list[c], words) == cudaSuccess;
        ctx->ws.miss.cap[c] = static_cast<uint32_t>(words);
    }
    ok = ok && dmalloc(&ctx->ws.dense.by_piece, mb + 1) == cudaSuccess;
    ok = ok && dmalloc(&ctx->ws.dense.extras, mb + 1) == cudaSuccess;
    ctx->ws.dense.extras_cap = static_cast<uint32_t>(mb + 1);
    ok = ok && dmalloc(&ctx->ws.dense.tile_pieces, (mb >> 11) + 2 + 2 * kMaxPipeChunks) == cudaSuccess;
    ok = ok && dmalloc(&ctx->ws.dense.piece_base, (mb >> 11) + 2 + 2 * kMaxPipeChunks) == cudaSuccess;
    ok = ok && dmalloc(&ctx->ws.pstart_bits, nw) == cudaSuccess;
    ok = ok && dmalloc(&ctx->ws.block_prompt, (mb >> kPromptBlockShift) + 2 + 2 * kMaxPipeChunks) == cudaSuccess;
    ctx->ws.fix_cap = static_cast<uint32_t>(mb / 16 + 2 + 2 * kMaxPipeChunks);
    ok = ok && dmalloc(&ctx->ws.fix_list, ctx->ws.fix_cap) == cudaSuccess;
    ok = ok && dmalloc(&ctx->d_dec_sums, mb / kDecodeTile + 2) == cudaSuccess;
    ok = ok && dmalloc(&ctx->d_dec_base, mb / kDecodeTile + 2) == cudaSuccess;
    ok = ok && dmalloc(&ctx->ws.tile_counts, nt) == cudaSuccess;
    ok = ok && dmalloc(&ctx->ws.tile_base, nt) == cudaSuccess;
    ok = ok && dmalloc(&ctx->ws.status, 1) == cudaSuccess;
    ok = ok && cudaMallocHost(reinterpret_cast<void**>(&ctx->h_status), sizeof(DeviceStatus)) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && dmalloc(&ctx->d_status_arr, kMaxPipeChunks) == cudaSuccess;
    ok = ok && cudaMallocHost(reinterpret_cast<void**>(&ctx->h_status_arr), sizeof(DeviceStatus) * kMaxPipeChunks) == cudaSuccess;
    ok = ok && cudaMallocHost(reinterpret_cast<void**>(&ctx->h_offs_stage), sizeof(uint64_t) * (mp + 1 + kMaxPipeChunks)) == cudaSuccess;
    {   // a pipelined host call gives earlier sub-batches the higher priority, so that they finish first and their downloads
        // run while the later ones compute (with equal priorities the sub-batches finished together and the downloads queued up
        // at the end: tools/pipe_trace.py)
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        const int levels = prio_lo - prio_hi + 1;
        for (int k = 1; ok && k < kFrontStreams; ++k)
            ok = cudaStreamCreateWithPriority(&ctx->front[k], cudaStreamNonBlocking, prio_hi + (k < levels ? k : levels - 1)) == cudaSuccess;
        for (int k = 0; ok && k < kSideStreams; ++k) ok = cudaStreamCreateWithPriority(&ctx->side[k], cudaStreamNonBlocking, prio_hi) == cudaSuccess;
        for (int k = 0; ok && k < kSideStreams; ++k) ok = cudaStreamCreateWithPriority(&ctx->side2[k], cudaStreamNonBlocking, prio_hi) == cudaSuccess;
    }
    {   // the long-piece kernels are latency-bound and small: their CTAs go first, the short-piece kernels fill the rest
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
#ifndef CFBPE_AUX_PRIO_LOW
#define CFBPE_AUX_PRIO_LOW 0      // A/B: 1 = the long-piece streams at the LOWEST priority (they take what the short-piece kernels leave)
#endif
        const int aux_prio = CFBPE_AUX_PRIO_LOW ? prio_lo : prio_hi;
        ok = ok && cudaStreamCreateWithPriority(&ctx->aux_stream, cudaStreamNonBlocking, aux_prio) == cudaSuccess;
        ok = ok && cudaStreamCreateWithPriority(&ctx->aux2_stream, cudaStreamNonBlocking, aux_prio) == cudaSuccess;
    }
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_ws, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_join2, cudaEventDisableTiming) == cudaSuccess;
    for (int k = 0; ok && k < kMaxPipeChunks; ++k) ok = cudaEventCreateWithFlags(&ctx->ev_list[k], cudaEventDisableTiming) == cudaSuccess;
    for (int k = 0; ok && k < kMaxPipeChunks; ++k) ok = cudaEventCreateWithFlags(&ctx->ev_scan[k], cudaEventDisableTiming) == cudaSuccess;
    for (int k = 0; ok && k < kMaxPipeChunks; ++k)
        ok = cudaEventCreateWithFlags(&ctx->ev_h2d[k], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_front[k], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_done[k], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_chain[k], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && dmalloc(&ctx->d_uc1, sizeof cfbpe_uc_stage1) == cudaSuccess;
    ok = ok && dmalloc(&ctx->d_uc2, sizeof cfbpe_uc_stage2) == cudaSuccess;
    ok = ok && cudaMemcpy(ctx->d_uc1, cfbpe_uc_stage1, sizeof cfbpe_uc_stage1, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemcpy(ctx->d_uc2, cfbpe_uc_stage2, sizeof cfbpe_uc_stage2, cudaMemcpyHostToDevice) == cudaSuccess;
    {
        std::vector<uint16_t> fsm(kNumPatterns * kPretokTableSize);
        uint8_t ascii[128];
        build_pretok_tables(fsm.data());
        build_ascii_classes(ascii);
        ok = ok && dmalloc(&ctx->d_ascii, 128) == cudaSuccess;
        ok = ok && dmalloc(&ctx->d_fsm, fsm.size()) == cudaSuccess;
        ok = ok && cudaMemcpy(ctx->d_ascii, ascii, 128, cudaMemcpyHostToDevice) == cudaSuccess;
        ok = ok && cudaMemcpy(ctx->d_fsm, fsm.data(), fsm.size() * sizeof(uint16_t), cudaMemcpyHostToDevice) == cudaSuccess;
    }
    {
        std::vector<SplitTablesHost> st(1);
        build_split_tables(st.data());
        ok = ok && dmalloc(&ctx->d_split_tables, sizeof(SplitTablesHost)) == cudaSuccess;
        ok = ok && cudaMemcpy(ctx->d_split_tables, st.data(), sizeof(SplitTablesHost), cudaMemcpyHostToDevice) == cudaSuccess;
    }
    ok = ok && cudaMemset(ctx->d_bytes, 0, mb + 256 + 16 * (kMaxPipeChunks + 1)) == cudaSuccess;
    for (int k = 0; ok && k < CFBPE_NUM_KERNELS; ++k)
        ok = cudaEventCreate(&ctx->prof.ev[k][0]) == cudaSuccess && cudaEventCreate(&ctx->prof.ev[k][1]) == cudaSuccess;
    for (int k = 0; ok && k < 2; ++k)
        ok = cudaEventCreate(&ctx->prof.h2d[k]) == cudaSuccess && cudaEventCreate(&ctx->prof.d2h[k]) == cudaSuccess &&
             cudaEventCreate(&ctx->prof.total[k]) == cudaSuccess;
    if (!ok) {
        cudaGetLastError();
        cfbpe_destroy(ctx);
        return CFBPE_ENOMEM;
    }
    ctx->uc = UcTables{ctx->d_uc1, ctx->d_uc2, ctx->d_ascii, ctx->d_fsm,
                       ctx->d_split_tables + offsetof(SplitTablesHost, cls256),
                       reinterpret_cast<const uint16_t*>(ctx->d_split_tables + offsetof(SplitTablesHost, fsm16)),
                       reinterpret_cast<const uint16_t*>(ctx->d_split_tables + offsetof(SplitTablesHost, ctx16)),
                       reinterpret_cast<const uint64_t*>(ctx->d_split_tables + offsetof(SplitTablesHost, prod)),
                       reinterpret_cast<const ProdInfo*>(ctx->d_split_tables + offsetof(SplitTablesHost, prod_info)),
                       ctx->d_split_tables + offsetof(SplitTablesHost, prod_skip),
                       ctx->d_split_tables + offsetof(SplitTablesHost, prod_start)};
    if (const char* e = std::getenv("CFBPE_PIPE_CHUNK_BYTES")) { const uint64_t v = std::strtoull(e, nullptr, 10); if (v >= 1024) ctx->pipe_chunk = v; }
    if (const char* e = std::getenv("CFBPE_PIPE_MIN_BYTES")) { const uint64_t v = std::strtoull(e, nullptr, 10); if (v >= 1) ctx->pipe_min = v; }
    *out = ctx;
    return CFBPE_OK;
}

void cfbpe_destroy(cfbpe_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();             // device-path calls may still be running on the caller's streams
    cudaFree(ctx->d_bytes); cudaFree(ctx->d_offsets); cudaFree(ctx->d_vocab_ids);
    cudaFree(ctx->d_out_ids); cudaFree(ctx->d_out_offsets); cudaFree(ctx->d_out_counts);
    cudaFree(ctx->ws.piece_bits); cudaFree(ctx->ws.tok_bits); cudaFree(ctx->ws.ids_by_pos);
    cudaFree(ctx->ws.lscratch.rank); cudaFree(ctx->ws.lscratch.aux0); cudaFree(ctx->ws.lscratch.aux1);
    for (uint32_t c = 0; c < 3; ++c) cudaFree(ctx->ws.miss.list[c]);
    cudaFree(ctx->d_dec_sums); cudaFree(ctx->d_dec_base);
    cudaFree(ctx->ws.dense.by_piece); cudaFree(ctx->ws.dense.extras); cudaFree(ctx->ws.dense.tile_pieces); cudaFree(ctx->ws.dense.piece_base);
    cudaFree(ctx->ws.fix_list); cudaFree(ctx->ws.pstart_bits); cudaFree(ctx->ws.block_prompt); cudaFree(ctx->d_split_tables);
    cudaFree(ctx->ws.long_list); cudaFree(ctx->ws.tile_counts); cudaFree(ctx->ws.tile_base); cudaFree(ctx->ws.status);
    cudaFree(ctx->d_uc1); cudaFree(ctx->d_uc2); cudaFree(ctx->d_ascii); cudaFree(ctx->d_fsm);
    if (ctx->h_status) cudaFreeHost(ctx->h_status);
    if (ctx->h_status_arr) cudaFreeHost(ctx->h_status_arr);
    if (ctx->h_offs_stage) cudaFreeHost(ctx->h_offs_stage);
    cudaFree(ctx->d_status_arr);
    for (int k = 0; k < kMaxPipeChunks; ++k) { if (ctx->ev_h2d[k]) cudaEventDestroy(ctx->ev_h2d[k]); if (ctx->ev_done[k]) cudaEventDestroy(ctx->ev_done[k]); if (ctx->ev_front[k]) cudaEventDestroy(ctx->ev_front[k]); if (ctx->ev_chain[k]) cudaEventDestroy(ctx->ev_chain[k]); }
    for (int k = 1; k < kFrontStreams; ++k) if (ctx->front[k]) cudaStreamDestroy(ctx->front[k]);
    for (int k = 0; k < kSideStreams; ++k) if (ctx->side[k]) cudaStreamDestroy(ctx->side[k]);
    for (int k = 0; k < kMaxPipeChunks; ++k) if (ctx->ev_scan[k]) cudaEventDestroy(ctx->ev_scan[k]);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    if (ctx->aux2_stream) cudaStreamDestroy(ctx->aux2_stream);
    if (ctx->ev_join2) cudaEventDestroy(ctx->ev_join2);
    for (int k = 0; k < kSideStreams; ++k) if (ctx->side2[k]) cudaStreamDestroy(ctx->side2[k]);
    for (int k = 0; k < kMaxPipeChunks; ++k) if (ctx->ev_list[k]) cudaEventDestroy(ctx->ev_list[k]);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_ws) cudaEventDestroy(ctx->ev_ws);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    for (auto& v : ctx->vocabs) { if (v.d_blob) cudaFree(v.d_blob); if (v.d_hot) cudaFree(v.d_hot); }
    for (int k = 0; k < CFBPE_NUM_KERNELS; ++k) for (int j = 0; j < 2; ++j) if (ctx->prof.ev[k][j]) cudaEventDestroy(ctx->prof.ev[k][j]);
    for (int j = 0; j < 2; ++j) {
        if (ctx->prof.h2d[j]) cudaEventDestroy(ctx->prof.h2d[j]);
        if (ctx->prof.d2h[j]) cudaEventDestroy(ctx->prof.d2h[j]);
        if (ctx->prof.total[j]) cudaEventDestroy(ctx->prof.total[j]);
    }
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* cfbpe_last_error(const cfbpe_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int cfbpe_vocab_load(cfbpe_ctx* ctx, uint32_t vocab_id, const uint8_t* ranks_file, size_t len, uint32_t format,
                     uint32_t pattern_id, uint32_t max_ranks) {
    if (!ctx) return CFBPE_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->err.clear();
    if (vocab_id >= CFBPE_MAX_VOCABS) return fail(ctx, CFBPE_EINVAL, "vocab_id out of range");
    if (!ranks_file || !len) return fail(ctx, CFBPE_EINVAL, "empty rank file");
    if (pattern_id >= CFBPE_PATTERN_COUNT) return fail(ctx, CFBPE_EINVAL, "unknown pattern id");
    std::vector<std::string> toks;
    std::string e;
    int rc;
    if (format == CFBPE_FORMAT_TIKTOKEN) rc = parse_tiktoken(ranks_file, len, max_ranks, toks, e);
    else if (format == CFBPE_FORMAT_TEKKEN_JSON) rc = parse_tekken_json(ranks_file, len, max_ranks, toks, e);
    else return fail(ctx, CFBPE_EINVAL, "unknown rank-file format");
    if (rc) return fail(ctx, rc, e);
    std::vector<uint8_t> blob;
    rc = build_tables(toks, pattern_id, blob, e);
    if (rc) return fail(ctx, rc, e);
    CK(cudaSetDevice(ctx->device));
    return install_blob(ctx, vocab_id, std::move(blob));
}

int cfbpe_vocab_get_info(const cfbpe_ctx* ctx, uint32_t vocab_id, cfbpe_vocab_info* out) {
    if (!ctx || !out || vocab_id >= CFBPE_MAX_VOCABS) return CFBPE_EINVAL;
    const VocabSlot& v = ctx->vocabs[vocab_id];
    if (!v.loaded) return CFBPE_ENOENT;
    out->n_ranks = v.hdr.n_ranks;
    out->pattern_id = v.hdr.pattern_id;
    out->max_token_len = v.hdr.max_token_len;
    out->n_pair_entries = v.hdr.n_pair_entries;
    out->table_bytes = v.hdr.total_bytes;
    return CFBPE_OK;
}

int cfbpe_vocab_export(const cfbpe_ctx* ctx, uint32_t vocab_id, uint8_t* buf, uint64_t cap, uint64_t* size) {
    if (!ctx || vocab_id >= CFBPE_MAX_VOCABS) return CFBPE_EINVAL;
    const VocabSlot& v = ctx->vocabs[vocab_id];
    if (!v.loaded) return CFBPE_ENOENT;
    if (size) *size = v.h_blob.size();
    if (!buf) return CFBPE_OK;
    if (cap < v.h_blob.size()) return CFBPE_ENOSPC;
    std::memcpy(buf, v.h_blob.data(), v.h_blob.size());
    return CFBPE_OK;
}

int cfbpe_vocab_import(cfbpe_ctx* ctx, uint32_t vocab_id, const uint8_t* buf, uint64_t size) {
    if (!ctx) return CFBPE_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->err.clear();
    if (vocab_id >= CFBPE_MAX_VOCABS || !buf) return fail(ctx, CFBPE_EINVAL, "bad argument");
    std::string e;
    int rc = validate_tables(buf, size, e);
    if (rc) return fail(ctx, rc, e);
    CK(cudaSetDevice(ctx->device));
    return install_blob(ctx, vocab_id, std::vector<uint8_t>(buf, buf + size));
}

int cfbpe_encode_batch(cfbpe_ctx* ctx, uint32_t n_prompts, const uint8_t* bytes, const uint64_t* offsets,
                       const uint8_t* vocab_ids, uint32_t* out_ids, uint64_t out_cap, uint64_t* out_offsets,
                       uint32_t* out_counts) {
    if (!ctx) return CFBPE_EINVAL;
    return run_host(ctx, n_prompts, bytes, offsets, vocab_ids, out_ids, out_cap, out_offsets, out_counts, true);
}

int cfbpe_count_batch(cfbpe_ctx* ctx, uint32_t n_prompts, const uint8_t* bytes, const uint64_t* offsets,
                      const uint8_t* vocab_ids, uint32_t* out_counts) {
    if (!ctx) return CFBPE_EINVAL;
    if (!out_counts && n_prompts) return fail(ctx, CFBPE_EINVAL, "out_counts is NULL");
    return run_host(ctx, n_prompts, bytes, offsets, vocab_ids, nullptr, 0, nullptr, out_counts, false);
}

int cfbpe_decode_batch(cfbpe_ctx* ctx, uint32_t n_seqs, const uint32_t* ids, const uint64_t* id_offsets,
                       const uint8_t* vocab_ids, uint8_t* out_bytes, uint64_t out_cap, uint64_t* out_offsets) {
    if (!ctx) return CFBPE_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->err.clear();
    if (!id_offsets || !out_offsets) return fail(ctx, CFBPE_EINVAL, "offsets pointer is NULL");
    if (n_seqs > ctx->max_prompts) return fail(ctx, CFBPE_EINVAL, "batch exceeds the limits of this context");
    if (id_offsets[0] != 0) return fail(ctx, CFBPE_EINVAL, "id_offsets[0] must be 0");
    for (uint32_t i = 0; i < n_seqs; ++i) if (id_offsets[i + 1] < id_offsets[i]) return fail(ctx, CFBPE_EINVAL, "id_offsets must not decrease");
    const uint64_t n_ids = id_offsets[n_seqs];
    if (n_ids > ctx->max_bytes) return fail(ctx, CFBPE_EINVAL, "batch exceeds the limits of this context");
    if (n_ids && !ids) return fail(ctx, CFBPE_EINVAL, "ids is NULL");
    for (uint32_t i = 0; vocab_ids && i < n_seqs; ++i)
        if (vocab_ids[i] >= kMaxVocabs || !ctx->vocabs[vocab_ids[i]].loaded) return fail(ctx, CFBPE_ENOENT, "vocab " + std::to_string(vocab_ids[i]) + " is not loaded");
    if (!vocab_ids && !ctx->vocabs[0].loaded) return fail(ctx, CFBPE_ENOENT, "vocab 0 is not loaded");
    CK(cudaSetDevice(ctx->device));
    if (ctx->ws_pending) { CK(cudaEventSynchronize(ctx->ev_ws)); ctx->ws_pending = false; }
    cudaStream_t s = ctx->stream;
    if (n_ids) CK(cudaMemcpyAsync(ctx->d_out_ids, ids, n_ids * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->d_offsets, id_offsets, (static_cast<uint64_t>(n_seqs) + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    if (vocab_ids && n_seqs) CK(cudaMemcpyAsync(ctx->d_vocab_ids, vocab_ids, n_seqs, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(ctx->ws.status, 0, sizeof(DeviceStatus), s));
    DecodeView d{ctx->d_out_ids, ctx->d_offsets, vocab_ids ? ctx->d_vocab_ids : nullptr, n_seqs, n_ids};
    const uint32_t n_tiles = static_cast<uint32_t>((n_ids + kDecodeTile - 1) / kDecodeTile);
    if (n_tiles) decode_len_kernel<<<n_tiles, 256, 0, s>>>(d, ctx->vs, ctx->ws.ids_by_pos, ctx->d_dec_sums, ctx->ws.status);
    tile_scan_kernel<<<1, n_tiles ? 1024 : 32, 0, s>>>(ctx->d_dec_sums, n_tiles, ctx->d_dec_base, ctx->ws.status, nullptr);
    if (n_tiles) decode_copy_kernel<<<n_tiles, 256, 0, s>>>(d, ctx->vs, ctx->ws.ids_by_pos, ctx->d_dec_base, ctx->d_bytes, ctx->max_bytes);
    decode_offsets_kernel<<<static_cast<unsigned>((static_cast<uint64_t>(n_seqs) + 1 + 255) / 256), 256, 0, s>>>(d, ctx->ws.ids_by_pos, ctx->d_dec_base, ctx->d_out_offsets, ctx->ws.status);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->h_status, ctx->ws.status, sizeof(DeviceStatus), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const DeviceStatus st = *ctx->h_status;
    if (st.bad_utf8) return fail(ctx, CFBPE_EINVAL, "a token id is outside its vocabulary");
    const uint64_t total = st.tok_end;
    if (total > ctx->max_bytes) return fail(ctx, CFBPE_EINVAL, "the decoded batch exceeds max_batch_bytes of this context");
    CK(cudaMemcpyAsync(out_offsets, ctx->d_out_offsets, (static_cast<uint64_t>(n_seqs) + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    if (total > out_cap || (total && !out_bytes)) {
        CK(cudaStreamSynchronize(s));
        out_offsets[n_seqs] = total;
        return fail(ctx, CFBPE_ENOSPC, "out_cap too small: need " + std::to_string(total) + " bytes");
    }
    if (total) CK(cudaMemcpyAsync(out_bytes, ctx->d_bytes, total, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return CFBPE_OK;
}

int cfbpe_encode_batch_device(cfbpe_ctx* ctx, uint32_t n_prompts, const uint8_t* d_bytes, uint64_t total_bytes,
                              const uint64_t* d_offsets, const uint8_t* d_vocab_ids, uint32_t* d_out_ids,
                              uint64_t out_cap, uint64_t* d_out_offsets, uint32_t* d_out_counts, uint64_t* n_tokens,
                              void* stream) {
    if (!ctx) return CFBPE_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->err.clear();
    if (n_prompts > ctx->max_prompts || total_bytes > ctx->max_bytes) return fail(ctx, CFBPE_EINVAL, "batch exceeds the limits of this context");
    if (!d_offsets || !d_out_offsets || (total_bytes && !d_bytes)) return fail(ctx, CFBPE_EINVAL, "device pointer is NULL");
    if (!ctx->vocabs[0].loaded && !d_vocab_ids) return fail(ctx, CFBPE_ENOENT, "vocab 0 is not loaded");
    if (!ctx->vs.loaded_mask) return fail(ctx, CFBPE_ENOENT, "no vocabulary is loaded");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // one workspace per context: a call on another stream waits (on the device) for the previous device-path call
    if (ctx->ws_pending) CK(cudaStreamWaitEvent(s, ctx->ev_ws, 0));
    ProfEvents* prof = ctx->profiling ? &ctx->prof : nullptr;
    if (prof) { std::memset(prof->launched, 0, sizeof prof->launched); cudaEventRecord(prof->total[0], s); cudaEventRecord(prof->h2d[0], s); cudaEventRecord(prof->h2d[1], s); }
    BatchView b{d_bytes, d_offsets, d_vocab_ids, n_prompts, total_bytes};
    enqueue_encode(b, ctx->vs, ctx->uc, ctx->ws, d_out_ids, out_cap, d_out_offsets, d_out_counts,
                   static_cast<uint32_t>(ctx->sm_count * 4), s, prof ? s : ctx->aux_stream, prof ? s : ctx->aux2_stream,
                   ctx->ev_fork, ctx->ev_join, ctx->ev_join2, prof);   // profiling: one stream, so that the per-kernel times do not overlap
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev_ws, s));
    ctx->ws_pending = true;
    ctx->dev_out_cap = out_cap;
    ctx->dev_want_ids = d_out_ids != nullptr;
    if (prof) { cudaEventRecord(prof->d2h[0], s); cudaEventRecord(prof->d2h[1], s); cudaEventRecord(prof->total[1], s); }
    if (n_tokens || prof) {
        CK(cudaMemcpyAsync(ctx->h_status, ctx->ws.status, sizeof(DeviceStatus), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (prof) fill_profile(ctx, total_bytes);
        const DeviceStatus st = *ctx->h_status;
        if (n_tokens) *n_tokens = st.n_tokens;
        if (st.long_overflow || st.miss_overflow) return fail(ctx, CFBPE_EIO, "internal: long-piece list overflow");
        if (st.bad_vocab) return fail(ctx, CFBPE_ENOENT, "a prompt names a vocabulary that is not loaded");
        if (st.bad_utf8) return fail(ctx, CFBPE_EILSEQ, "a prompt holds malformed UTF-8");
        if (d_out_ids && st.n_tokens > out_cap) return fail(ctx, CFBPE_ENOSPC, "out_cap too small: need " + std::to_string(st.n_tokens) + " ids");
    }
    return CFBPE_OK;
}

int cfbpe_device_status(cfbpe_ctx* ctx, void* stream) {
    if (!ctx) return CFBPE_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CK(cudaMemcpyAsync(ctx->h_status, ctx->ws.status, sizeof(DeviceStatus), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (ctx->h_status->long_overflow || ctx->h_status->miss_overflow) return fail(ctx, CFBPE_EIO, "internal: long-piece list overflow");
    if (ctx->h_status->bad_vocab) return fail(ctx, CFBPE_ENOENT, "a prompt names a vocabulary that is not loaded");
    if (ctx->h_status->bad_utf8) return fail(ctx, CFBPE_EILSEQ, "a prompt holds malformed UTF-8");
    if (ctx->dev_want_ids && ctx->h_status->n_tokens > ctx->dev_out_cap)
        return fail(ctx, CFBPE_ENOSPC, "out_cap too small: need " + std::to_string(ctx->h_status->n_tokens) + " ids");
    return CFBPE_OK;
}

void* cfbpe_host_alloc(cfbpe_ctx* ctx, size_t size) {
    if (!ctx) return nullptr;
    void* p = nullptr;
    cudaSetDevice(ctx->device);
    if (cudaMallocHost(&p, size ? size : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void cfbpe_host_free(cfbpe_ctx* ctx, void* ptr) {
    if (!ctx || !ptr) return;
    cudaSetDevice(ctx->device);
    cudaFreeHost(ptr);
}

int cfbpe_profile_enable(cfbpe_ctx* ctx, int on) {
    if (!ctx) return CFBPE_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->profiling = on != 0;
    ctx->prof_ready = false;
    return CFBPE_OK;
}
int cfbpe_profile_read(cfbpe_ctx* ctx, cfbpe_profile* out) {
    if (!ctx || !out) return CFBPE_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (!ctx->prof_ready) return CFBPE_ENOENT;
    *out = ctx->last_profile;
    return CFBPE_OK;
}

}  // extern "C"
