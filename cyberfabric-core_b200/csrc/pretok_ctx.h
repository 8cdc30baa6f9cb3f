// pretok_ctx.h -- host-built tables of the lane-per-16-bytes split kernel (pretok_lanes.cuh).
//
//   cls256    byte -> class byte: ASCII bytes their extended class (pretok_fsm.h X_*), continuation bytes X_CONT,
//             lead bytes X_LEAD (the kernel replaces those by  class | (length - 1) << 4  after decoding the character)
//   fsm16     the transition tables of pretok_fsm.h with a row stride of 16 (index = state << 4 | class)
//   ctx16     the CONTEXT automaton: what pretok_sync.cuh::sync_rule looks at -- class of the previous character, number of
//             consecutive letters (<= 3) and of run punctuation (<= 2) before the position -- is a function of the last three
//             characters, so it is itself a small automaton over classes.  An entry gives, for context c and the class x of
//             the character at the position:  the context after x | sync_rule(x, c) << 8  (the automaton state a thread may
//             assume BEFORE consuming x, kNoSync = 0xFF when the context does not determine it).
//             One table per casedness (run punctuation differs).  A lane knows, after one table lookup per character, where
//             its first sync point is, and a lane coming from the left knows -- from the same table, the same context -- where
//             the lane to its right started.
//   ctxinfo   context -> prevx | nlet << 4 | npun << 6  (for the per-character walker that takes over on long runs)
//
// Host code; the tables are a restatement of sync_rule / the update rule of the old per-character loop, built by enumeration
// (breadth first from the context of a prompt start), so they cannot drift from the predicate.
#pragma once
#include <stdint.h>
#include <string.h>

#include "pretok_fsm.h"
#include "pretok_sync.cuh"

namespace cfbpe {

enum : uint32_t { X_CONT = 12, X_LEAD = 13 };
constexpr uint32_t kFsm16Size = S_COUNT * 16;        // u16 entries per pattern
constexpr uint32_t kCtxMax = 64;                     // contexts per casedness (the enumeration finds ~25)
constexpr uint32_t kCtx16Size = kCtxMax * 16;        // u16 entries per casedness
constexpr uint32_t kCtxStart = 0;                    // the context at a prompt start: (X_EOT, 0, 0)

// ---- the PRODUCT automaton: split state x context in one table, so that a step of the walk is ONE lookup ----
// A walker is in one of: DONE (state 0: nothing left to do, every entry leads back to it), NOSYNC(c) (states 1 + c: it has
// not met a sync point yet, only the context moves), SKIPn(c) (inside a contraction that was taken whole: n characters
// to go, then START), or (q, c) with q a state of pretok_fsm.h.  ~90 states per pattern are reachable.
// Entry (two words, read with one 8-byte load), for state s and class x:
//   lo  bits 14..7  next state (so lo & PE_NEXT_MASK is the byte offset of its row: 16 classes x 8 bytes = 128 bytes a row)
//       bit 0 A_B_NOW   bits 1..3 SET_ALC | SET_LAST | SET_LBE   bit 4 the position is a sync point (hand-over in the second
//       block)   bit 5 rare: a contraction may start here, or an undecided state has to be resolved (out-of-line path)
//       bits 16..18 EMIT_ALC | EMIT_LAST | EMIT_LBE
//   hi  byte mask of the remembered positions this step sets (0xFF per position: alc | last << 8 | lbe << 16)
// Continuation bytes (class X_CONT) and the classes that never occur map every state to itself without flags.
constexpr uint32_t kProdMax = 128;                   // states per pattern (the enumeration finds < 100: checked by the tests)
constexpr uint32_t kProdRowBytes = 16 * 8;           // 16 classes x 8 bytes
constexpr uint32_t kProdTableBytes = kProdMax * kProdRowBytes;      // 16 KB per pattern
enum : uint32_t {
    PE_B_NOW = 1u << 0, PE_SET_ALC = 1u << 1, PE_SET_LAST = 1u << 2, PE_SET_LBE = 1u << 3, PE_SYNC = 1u << 4, PE_RARE = 1u << 5,
    PE_NEXT_SHIFT = 7, PE_NEXT_MASK = 0xFFu << 7, PE_EMIT_ALC = 1u << 16, PE_EMIT_LAST = 1u << 17, PE_EMIT_LBE = 1u << 18,
    PE_EMIT_ANY = 7u << 16
};
enum : uint32_t { PQ_DONE = 0xFF, PQ_NOSYNC = 0xFE, PQ_SKIP1 = 0xFD, PQ_SKIP2 = 0xFC };    // pseudo split states in ProdInfo::q
struct ProdInfo { uint8_t q, ctx; };                 // what a product state is made of

// Every array starts on a 16-byte boundary and is a multiple of 16 bytes long: K1 stages them into shared memory with TMA bulk
// copies (cp.async.bulk: 16-byte granularity on both sides).
struct SplitTablesHost {
    alignas(16) uint8_t cls256[256];
    alignas(16) uint16_t fsm16[kNumPatterns * kFsm16Size];
    alignas(16) uint16_t ctx16[2 * kCtx16Size];
    alignas(16) uint16_t ctxinfo[2 * kCtxMax];
    alignas(16) uint32_t n_ctx[4];
    alignas(16) uint64_t prod[kNumPatterns * kProdMax * 16];     // [pattern][state * 16 + class]
    alignas(16) ProdInfo prod_info[kNumPatterns * kProdMax];
    alignas(16) uint8_t prod_skip[kNumPatterns * 2 * kCtxMax];   // [pattern][n - 1][context] -> state SKIPn(context)
    alignas(16) uint8_t prod_start[16];                          // the state (S_START, kCtxStart): where a prompt begins (kNumPatterns used)
    alignas(16) uint32_t n_prod[kNumPatterns];
};
static_assert(sizeof(uint16_t) * kNumPatterns * kFsm16Size % 16 == 0 && kProdTableBytes % 16 == 0 && sizeof(ProdInfo) * kNumPatterns * kProdMax % 16 == 0,
              "table sizes are multiples of 16 bytes (bulk copies)");

inline void build_split_tables(SplitTablesHost* t) {
    memset(t, 0, sizeof *t);
    uint8_t ascii[128];
    build_ascii_classes(ascii);
    for (uint32_t b = 0; b < 256; ++b) t->cls256[b] = b < 128 ? ascii[b] : (b < 0xC0 ? X_CONT : X_LEAD);
    for (uint32_t pat = 0; pat < kNumPatterns; ++pat)
        for (uint32_t st = 0; st < S_COUNT; ++st)
            for (uint32_t x = 0; x < 16; ++x)
                t->fsm16[pat * kFsm16Size + st * 16 + x] =
                    static_cast<uint16_t>(x < X_COUNT ? fsm_detail::transition(pat, st, x) : fsm_detail::transition(pat, st, X_OTHER));
    for (uint32_t cased = 0; cased < 2; ++cased) {
        uint16_t info[kCtxMax];
        uint32_t n = 0;
        info[n++] = static_cast<uint16_t>(X_EOT);                      // (prevx = EOT, nlet = 0, npun = 0)
        for (uint32_t c = 0; c < n; ++c) {
            const uint32_t prevx = info[c] & 15u, nlet = (info[c] >> 4) & 3u, npun = (info[c] >> 6) & 3u;
            for (uint32_t x = 0; x < 16; ++x) {
                uint16_t e = static_cast<uint16_t>(c | (kNoSync << 8));   // classes that are never looked up: stay
                if (x < X_EOT) {
                    const uint32_t nl = x_is_letter(x) ? (nlet < 3 ? nlet + 1 : 3u) : 0u;
                    const uint32_t np = x_is_run_punct(x, cased != 0) ? (npun < 2 ? npun + 1 : 2u) : 0u;
                    const uint16_t next = static_cast<uint16_t>(x | (nl << 4) | (np << 6));
                    uint32_t j = 0;
                    while (j < n && info[j] != next) ++j;
                    if (j == n) { if (n < kCtxMax) info[n++] = next; else j = 0; }   // (never more than kCtxMax: checked by the tests)
                    // a prompt start (context 0) is a sync point of its own (state START); sync_rule is never asked there
                    const uint32_t s = (c == kCtxStart) ? static_cast<uint32_t>(S_START) : sync_rule(x, prevx, nlet, npun, cased != 0);
                    e = static_cast<uint16_t>(j | (s << 8));
                }
                t->ctx16[cased * kCtx16Size + c * 16 + x] = e;
            }
        }
        t->n_ctx[cased] = n;
        for (uint32_t c = 0; c < n; ++c) t->ctxinfo[cased * kCtxMax + c] = info[c];
    }
    // ---- product automaton, per pattern: breadth first from DONE, the NOSYNC states and the state of a prompt start
    for (uint32_t pat = 0; pat < kNumPatterns; ++pat) {
        const uint32_t cased = pat & 1u, nctx = t->n_ctx[cased];
        const uint16_t* ctx = t->ctx16 + cased * kCtx16Size;
        const uint16_t* fsm = t->fsm16 + pat * kFsm16Size;
        ProdInfo* info = t->prod_info + pat * kProdMax;
        uint64_t* tab = t->prod + static_cast<uint64_t>(pat) * kProdMax * 16;
        uint32_t n = 0;
        auto get = [&](uint32_t q, uint32_t c) -> uint32_t {
            for (uint32_t i = 0; i < n; ++i) if (info[i].q == q && info[i].ctx == c) return i;
            if (n >= kProdMax) return 0;                              // (never: checked by the tests through n_prod)
            info[n].q = static_cast<uint8_t>(q); info[n].ctx = static_cast<uint8_t>(c);
            return n++;
        };
        get(PQ_DONE, 0);
        for (uint32_t c = 0; c < nctx; ++c) get(PQ_NOSYNC, c);       // states 1 + c
        t->prod_start[pat] = static_cast<uint8_t>(get(S_START, kCtxStart));
        for (uint32_t i = 0; i < n; ++i) {
            const uint32_t q = info[i].q, c = info[i].ctx;
            for (uint32_t x = 0; x < 16; ++x) {
                uint64_t e = static_cast<uint64_t>(i) << PE_NEXT_SHIFT;                        // stay, no flags
                if (q != PQ_DONE && x < X_EOT) {
                    const uint32_t ce = ctx[c * 16 + x], c2 = ce & 0xFFu, sy = ce >> 8;
                    uint32_t lo = sy != kNoSync ? static_cast<uint32_t>(PE_SYNC) : 0u, hi = 0;
                    uint32_t q2 = q;
                    if (q == PQ_NOSYNC && sy != kNoSync) q2 = sy;                             // the first sync point: the state follows from the context
                    if (q2 == PQ_NOSYNC) lo |= get(PQ_NOSYNC, c2) << PE_NEXT_SHIFT;
                    else if (q2 == PQ_SKIP2) lo = get(PQ_SKIP1, c2) << PE_NEXT_SHIFT;          // (no position inside a contraction is examined:
                    else if (q2 == PQ_SKIP1) lo = get(S_START, c2) << PE_NEXT_SHIFT;           //  no hand-over there)
                    else {
                        const uint32_t a = fsm[q2 * 16 + x];
                        if (a & A_B_NOW) lo |= PE_B_NOW;
                        if (a & A_SET_ALC) { lo |= PE_SET_ALC; hi |= 0xFFu; }
                        if (a & A_SET_LAST) { lo |= PE_SET_LAST; hi |= 0xFF00u; }
                        if (a & A_SET_LBE) { lo |= PE_SET_LBE; hi |= 0xFF0000u; }
                        if (a & A_EMIT_ALC) lo |= PE_EMIT_ALC;
                        if (a & A_EMIT_LAST) lo |= PE_EMIT_LAST;
                        if (a & A_EMIT_LBE) lo |= PE_EMIT_LBE;
                        if (a & (A_CONTR | A_RESOLVE)) lo |= PE_RARE;
                        if (a & A_CONTR) {                                                    // the states a contraction leads to
                            t->prod_skip[(pat * 2 + 0) * kCtxMax + c2] = static_cast<uint8_t>(get(PQ_SKIP1, c2));
                            t->prod_skip[(pat * 2 + 1) * kCtxMax + c2] = static_cast<uint8_t>(get(PQ_SKIP2, c2));
                        }
                        lo |= ((a & A_RESOLVE) ? 0u : get(a & A_STATE_MASK, c2)) << PE_NEXT_SHIFT;   // (resolve: the rare path stops the walker)
                    }
                    e = lo | (static_cast<uint64_t>(hi) << 32);
                }
                tab[i * 16 + x] = e;
            }
        }
        t->n_prod[pat] = n;
    }
}

}  // namespace cfbpe
