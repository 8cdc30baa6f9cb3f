// pretok_sync.cuh -- "sync points": positions where the automaton's state follows from a few characters of
// left context, so a K1 thread may begin there knowing nothing else (and the thread coming from the left hands over).
//
//   START   whitespace (not CR/LF) right after a non-whitespace character
//           CR/LF right after a letter or digit
//           a digit right after a non-digit
//           a punctuation character (class OTHER, not an apostrophe, not a mark) right after a letter or digit
//   LETTERS (uncased patterns) the three previous characters are letters: a contraction covers at most two
//   W_Y     (cased patterns) the previous character is a lower-case letter and the two before it are letters
//   ORUN    the two previous characters are punctuation (not '/', which a newline trailer may have eaten; not a
//           mark in the cased patterns, where marks are word characters): the second of them sits in a punctuation run
//
// sync_rule() is the predicate on classes; sync_state() evaluates it from memory (used to FIND a start);
// the running thread evaluates the same predicate from the classes it has just seen (no loads).
#pragma once
#include "pretok.cuh"
#include "pretok_fsm.h"

namespace cfbpe {

constexpr uint32_t kNoSync = 0xFFu;

CFBPE_HD uint32_t ext_class(const Ch& c) {   // C_* class + code point -> X_* class
    if (c.cp == ' ') return X_SPACE;
    if (c.cp == '\'') return X_APOS;
    if (c.cp == '/') return X_SLASH;
    return c.cls;
}
CFBPE_HD bool x_is_letter(uint32_t x) { return x == X_LU || x == X_LL || x == X_LO; }
CFBPE_HD bool x_is_ws(uint32_t x) { return x == X_WS || x == X_SPACE || x == X_CRLF; }
CFBPE_HD bool x_is_run_punct(uint32_t x, bool cased) { return x == X_OTHER || x == X_APOS || (!cased && x == X_M); }

// state BEFORE the character of class x is consumed, given the class of the previous character and the number
// (saturated at 3) of consecutive letters right before it; kNoSync if the context does not determine it
CFBPE_HD uint32_t sync_rule(uint32_t x, uint32_t prevx, uint32_t nlet, uint32_t npun, bool cased) {
    const bool prev_ln = x_is_letter(prevx) || prevx == X_N;
    if (x == X_WS || x == X_SPACE) return x_is_ws(prevx) ? kNoSync : static_cast<uint32_t>(S_START);
    if (x == X_CRLF) return prev_ln ? static_cast<uint32_t>(S_START) : kNoSync;
    if (x == X_N) return prevx != X_N ? static_cast<uint32_t>(S_START) : kNoSync;
    if ((x == X_OTHER || x == X_SLASH) && prev_ln) return S_START;
    if (nlet >= 3 && (cased ? prevx == X_LL : x_is_letter(prevx))) return cased ? static_cast<uint32_t>(S_W_Y) : static_cast<uint32_t>(S_LETTERS);
    if (npun >= 2) return S_ORUN;
    return kNoSync;
}

// evaluate the rule at byte position pos (ps < pos < pe) by decoding up to three characters to the left
template <typename Txt>
CFBPE_HD uint32_t sync_state(const Txt& s, uint64_t pos, uint64_t ps, uint64_t pe, const UcTables& uc, bool cased,
                             uint32_t* prevx_out = nullptr, uint32_t* nlet_out = nullptr, uint32_t* npun_out = nullptr) {
    const uint32_t b = s[pos];
    if ((b & 0xC0) == 0x80) return kNoSync;  // inside a character
    // two ASCII special cases of the rule below, decided from raw bytes (most positions of Latin-script text end here)
    if (pos >= ps + 3) {
        const uint32_t b1 = s[pos - 1], b2 = s[pos - 2], b3 = s[pos - 3];
        const bool l1 = ((b1 | 0x20u) - 'a') < 26u, l2 = ((b2 | 0x20u) - 'a') < 26u, l3 = ((b3 | 0x20u) - 'a') < 26u;
        if (b == ' ' && b1 < 0x80 && ascii_class(b1) != C_WS && ascii_class(b1) != C_CRLF) {   // space after an ASCII non-space
            if (prevx_out) *prevx_out = b1 == '\'' ? X_APOS : (b1 == '/' ? X_SLASH : ascii_class(b1));
            if (nlet_out) *nlet_out = 0;
            if (npun_out) *npun_out = 0;
            return S_START;
        }
        if (((b | 0x20u) - 'a') < 26u && l1 && l2 && l3 && (!cased || (b1 - 'a') < 26u)) {           // fourth letter of an ASCII word
            if (prevx_out) *prevx_out = (b1 - 'a') < 26u ? X_LL : X_LU;
            if (nlet_out) *nlet_out = 3;
            if (npun_out) *npun_out = 0;
            return cased ? static_cast<uint32_t>(S_W_Y) : static_cast<uint32_t>(S_LETTERS);
        }
    }
    int bad = 0;
    const Ch cur = get_char(s, pos, pe, uc, &bad);
    if (bad) return kNoSync;
    const Ch prev = get_prev_char(s, pos, ps, pe, uc);
    const uint32_t prevx = ext_class(prev);
    uint32_t nlet = 0;
    if (x_is_letter(prevx)) {
        nlet = 1;
        uint64_t q = pos - prev.len;
        while (nlet < 3 && q > ps) {
            const Ch c = get_prev_char(s, q, ps, pe, uc);
            if (!is_letter(c.cls)) break;
            ++nlet;
            q -= c.len;
        }
    }
    uint32_t npun = 0;
    if (x_is_run_punct(prevx, cased)) {
        npun = 1;
        const uint64_t q = pos - prev.len;
        if (q > ps) {
            const Ch c = get_prev_char(s, q, ps, pe, uc);
            if (x_is_run_punct(ext_class(c), cased)) npun = 2;
        }
    }
    if (prevx_out) *prevx_out = prevx;
    if (nlet_out) *nlet_out = nlet;
    if (npun_out) *npun_out = npun;
    return sync_rule(ext_class(cur), prevx, nlet, npun, cased);
}

}  // namespace cfbpe
