// pretok_fsm.h -- the pre-tokenizer as a table-driven character automaton.
//
// pretok.cuh states each regex alternative procedurally (one thread walks a whole match: good
// for reasoning, bad for SIMT -- ncu showed 4.3 active lanes per instruction).  Here the same
// semantics are compiled, per pattern, into ONE transition table T[state][class] -> action, so
// that all 32 lanes of a warp execute the identical instruction stream, one character per
// iteration, whatever match each lane is in.
//
// The automaton reads one character at a time and decides, at that character, whether a piece
// starts there (B_NOW).  Three decisions need unbounded look-ahead in the regex and are instead
// taken retroactively from remembered positions (at most one of each is live):
//   alc   position after the last CR/LF of the current whitespace run   (\s*[\r\n]+ ends there)
//   last  start of the last character of the current whitespace run     (\s+(?!\S) stops before it)
//   lbe   position after the last "both-sets" (Lm/Lo/M) character of a cased word whose [upper]*
//         part has not met a lower-case letter yet (where greedy [upper]* backtracks to)
// Contractions are the one place the automaton peeks ahead (<= 3 bytes), flagged by CONTR.
//
// build_pretok_table() is host code; it is the specification.  tests/ checks it against the
// oracle exhaustively over short class strings and by fuzzing, and against pretok.cuh::match_end.
#pragma once
#include <stdint.h>

namespace cfbpe {

// extended character classes (what one table column means)
enum : uint32_t {
    X_OTHER = 0, X_WS = 1, X_CRLF = 2, X_N = 3, X_LU = 4, X_LL = 5, X_LO = 6, X_M = 7,   // = C_* of pretok.cuh
    X_SPACE = 8, X_APOS = 9, X_SLASH = 10, X_EOT = 11, X_COUNT = 12
};
// states
enum : uint32_t {
    S_START = 0, S_LETTERS, S_D1, S_D2, S_D3, S_ORUN, S_OTRAIL, S_PFX_O,
    S_WS_N1S, S_WS_N1, S_WS_NMS, S_WS_NM, S_WS_C0, S_WS_C1S, S_WS_C1, S_WS_CMS, S_WS_CM,
    S_W_X0, S_W_XB0, S_W_XBU, S_W_Y,
    S_W_U,      // cased word, right after a both-sets letter, [upper] or [lower] part not known (a thread that STARTED there)
    S_W_V,      // cased word, inside a run of upper-case letters, W_X0 or W_XBU not known (likewise)
    S_COUNT
};
// action bits
enum : uint32_t {
    A_STATE_MASK = 31, A_B_NOW = 1u << 5, A_EMIT_ALC = 1u << 6, A_EMIT_LAST = 1u << 7, A_EMIT_LBE = 1u << 8,
    A_SET_ALC = 1u << 9, A_SET_LAST = 1u << 10, A_SET_LBE = 1u << 11,
    A_CONTR = 1u << 12,         // a contraction may start here: if it does, skip it and go to START
    A_CONTR_SUFFIX = 1u << 13,  // ... and it belongs to the piece that just ended (no boundary here)
    A_RESOLVE = 1u << 14        // S_W_U meets an upper-case letter / S_W_V meets the end of the word: the real state has to be found first
};
constexpr uint32_t kPretokTableSize = S_COUNT * X_COUNT;   // u16 entries per pattern
constexpr uint32_t kNumPatterns = 4;

namespace fsm_detail {
struct Traits { bool cased, contr_start, contr_suffix, slash, ws_eot; uint32_t max_digits; };
inline Traits traits(uint32_t pat) {
    Traits t;
    t.cased = (pat == 1 || pat == 3);
    t.contr_start = (pat == 0 || pat == 2);
    t.contr_suffix = (pat == 1);
    t.slash = (pat == 1 || pat == 3);
    t.ws_eot = (pat == 0);
    t.max_digits = (pat == 3) ? 1u : 3u;
    return t;
}
inline bool is_letter(uint32_t x) { return x == X_LU || x == X_LL || x == X_LO; }
inline bool is_ws(uint32_t x) { return x == X_WS || x == X_CRLF || x == X_SPACE; }
// [^\s\p{L}\p{N}] : OTHER, M, and the ASCII specials that are OTHER
inline bool is_punct(uint32_t x) { return x == X_OTHER || x == X_M || x == X_APOS || x == X_SLASH; }

// x processed as the FIRST character of a match
inline uint32_t start_with(const Traits& T, uint32_t x) {
    const uint32_t B = A_B_NOW;
    switch (x) {
    case X_WS: return B | S_WS_N1 | A_SET_LAST;
    case X_SPACE: return B | S_WS_N1S | A_SET_LAST;
    case X_CRLF: return B | S_WS_C0 | A_SET_ALC | A_SET_LAST;
    case X_N: return B | S_D1;
    case X_APOS: return B | S_PFX_O | (T.contr_start ? A_CONTR : 0u);
    case X_OTHER: case X_SLASH: return B | S_PFX_O;
    case X_M: return T.cased ? (B | S_W_XB0 | A_SET_LBE) : (B | S_PFX_O);
    case X_LU: return T.cased ? (B | S_W_X0) : (B | S_LETTERS);
    case X_LL: return T.cased ? (B | S_W_Y) : (B | S_LETTERS);
    case X_LO: return T.cased ? (B | S_W_XB0 | A_SET_LBE) : (B | S_LETTERS);
    default: return S_START;
    }
}
// a word (after an optional one-character prefix) begins with x: no boundary at x
inline uint32_t word_after_prefix(const Traits& T, uint32_t x) {
    if (!T.cased) return S_LETTERS;
    if (x == X_LU) return S_W_X0;
    if (x == X_LL) return S_W_Y;
    return S_W_XB0 | A_SET_LBE;   // Lo / M
}
inline bool starts_word(const Traits& T, uint32_t x) { return is_letter(x) || (T.cased && x == X_M); }
// a cased word ended just before x
inline uint32_t word_end(const Traits& T, uint32_t x) {
    uint32_t a = start_with(T, x);
    if (x == X_APOS && T.contr_suffix) a |= A_CONTR | A_CONTR_SUFFIX;
    return a;
}

inline uint32_t transition(uint32_t pat, uint32_t st, uint32_t x) {
    const Traits T = traits(pat);
    if (x == X_EOT) {   // end of prompt: settle what was pending, the next prompt starts clean
        uint32_t a = S_START;
        if (!T.ws_eot && (st == S_WS_C1S || st == S_WS_C1 || st == S_WS_CMS || st == S_WS_CM)) a |= A_EMIT_ALC;
        if (st == S_W_XBU) a |= A_EMIT_LBE;
        if (st == S_W_V && T.cased) a |= A_RESOLVE;      // W_XBU would emit, W_X0 would not
        return a;
    }
    switch (st) {
    case S_START: return start_with(T, x);
    case S_LETTERS: return is_letter(x) ? S_LETTERS : start_with(T, x);
    case S_D1: return (x == X_N && T.max_digits > 1) ? S_D2 : start_with(T, x);
    case S_D2: return (x == X_N && T.max_digits > 2) ? S_D3 : start_with(T, x);
    case S_D3: return start_with(T, x);
    case S_ORUN:
        if (is_punct(x)) return S_ORUN;
        if (x == X_CRLF) return S_OTRAIL;
        return start_with(T, x);
    case S_OTRAIL:
        if (x == X_CRLF || (x == X_SLASH && T.slash)) return S_OTRAIL;
        return start_with(T, x);
    case S_PFX_O:   // one punctuation character at the start of a match
        if (starts_word(T, x)) return word_after_prefix(T, x);
        if (is_punct(x)) return S_ORUN;
        if (x == X_CRLF) return S_OTRAIL;
        return start_with(T, x);
    default: break;
    }
    if (st >= S_WS_N1S && st <= S_WS_CM) {
        const bool has_crlf = st >= S_WS_C0;
        if (is_ws(x)) {
            if (x == X_CRLF) return S_WS_C0 | A_SET_ALC | A_SET_LAST;
            const bool sp = (x == X_SPACE);
            if (!has_crlf) return (sp ? S_WS_NMS : S_WS_NM) | A_SET_LAST;
            if (st == S_WS_C0) return (sp ? S_WS_C1S : S_WS_C1) | A_SET_LAST;
            return (sp ? S_WS_CMS : S_WS_CM) | A_SET_LAST;
        }
        // the run ends before x
        uint32_t retro = 0;
        bool rest_nonempty = true, last_is_space = false;
        switch (st) {
        case S_WS_N1S: last_is_space = true; break;
        case S_WS_N1: break;
        case S_WS_NMS: retro = A_EMIT_LAST; last_is_space = true; break;
        case S_WS_NM: retro = A_EMIT_LAST; break;
        case S_WS_C0: rest_nonempty = false; break;
        case S_WS_C1S: retro = A_EMIT_ALC; last_is_space = true; break;
        case S_WS_C1: retro = A_EMIT_ALC; break;
        case S_WS_CMS: retro = A_EMIT_ALC | A_EMIT_LAST; last_is_space = true; break;
        default: retro = A_EMIT_ALC | A_EMIT_LAST; break;   // S_WS_CM
        }
        if (rest_nonempty) {
            if (starts_word(T, x)) return retro | word_after_prefix(T, x);            // last ws char is the word's prefix
            if (last_is_space && is_punct(x)) return retro | S_ORUN;                  // " ?" of the punctuation alternative
        }
        return retro | start_with(T, x);
    }
    switch (st) {   // cased words
    case S_W_X0:
        if (x == X_LU) return S_W_X0;
        if (x == X_LO || x == X_M) return S_W_XB0 | A_SET_LBE;
        if (x == X_LL) return S_W_Y;
        return word_end(T, x);
    case S_W_XB0:
        if (x == X_LU) return S_W_XBU;
        if (x == X_LO || x == X_M) return S_W_XB0 | A_SET_LBE;
        if (x == X_LL) return S_W_Y;
        return word_end(T, x);
    case S_W_XBU:
        if (x == X_LU) return S_W_XBU;
        if (x == X_LO || x == X_M) return S_W_XB0 | A_SET_LBE;
        if (x == X_LL) return S_W_Y;
        return A_EMIT_LBE | word_end(T, x);
    case S_W_Y:
        if (x == X_LL || x == X_LO || x == X_M) return S_W_Y;
        if (x == X_LU) return A_B_NOW | S_W_X0;
        return word_end(T, x);
    case S_W_V:     // W_X0 and W_XBU agree until the word ends (W_XBU then gives back what greedy [upper]* took: boundary at lbe)
        if (!T.cased) return start_with(T, x);
        if (x == X_LU) return S_W_V;
        if (x == X_LO || x == X_M) return S_W_XB0 | A_SET_LBE;
        if (x == X_LL) return S_W_Y;
        return A_RESOLVE | S_W_V;
    case S_W_U:     // W_Y and W_XB0 agree on everything but an upper-case letter (boundary | none) -- and on lbe, kept as W_XB0 would
        if (!T.cased) return start_with(T, x);
        if (x == X_LO || x == X_M) return S_W_U | A_SET_LBE;
        if (x == X_LL) return S_W_Y;
        if (x == X_LU) return A_RESOLVE | S_W_U;
        return word_end(T, x);
    default: return start_with(T, x);
    }
}
}  // namespace fsm_detail

// table[pat][state * X_COUNT + class]
inline void build_pretok_tables(uint16_t* out /* kNumPatterns * kPretokTableSize */) {
    for (uint32_t pat = 0; pat < kNumPatterns; ++pat)
        for (uint32_t st = 0; st < S_COUNT; ++st)
            for (uint32_t x = 0; x < X_COUNT; ++x)
                out[pat * kPretokTableSize + st * X_COUNT + x] = static_cast<uint16_t>(fsm_detail::transition(pat, st, x));
}
// ASCII byte -> extended class
inline void build_ascii_classes(uint8_t* out /* 128 */) {
    for (uint32_t b = 0; b < 128; ++b) {
        uint32_t x = X_OTHER;
        if (b >= 'a' && b <= 'z') x = X_LL;
        else if (b >= 'A' && b <= 'Z') x = X_LU;
        else if (b >= '0' && b <= '9') x = X_N;
        else if (b == ' ') x = X_SPACE;
        else if (b == '\t' || b == 0x0B || b == 0x0C) x = X_WS;
        else if (b == '\n' || b == '\r') x = X_CRLF;
        else if (b == '\'') x = X_APOS;
        else if (b == '/') x = X_SLASH;
        out[b] = static_cast<uint8_t>(x);
    }
}
}  // namespace cfbpe
