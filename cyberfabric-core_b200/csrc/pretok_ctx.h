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

struct SplitTablesHost {
    uint8_t cls256[256];
    uint16_t fsm16[kNumPatterns * kFsm16Size];
    uint16_t ctx16[2 * kCtx16Size];
    uint16_t ctxinfo[2 * kCtxMax];
    uint32_t n_ctx[2];
};

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
}

}  // namespace cfbpe
