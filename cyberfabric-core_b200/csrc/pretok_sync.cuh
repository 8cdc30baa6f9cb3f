// pretok_sync.cuh -- "sync points": positions where the automaton's state follows from a few characters of
// left context, so a K1 thread may begin there knowing nothing else (and the thread coming from the left hands over).
//
//   START   whitespace (not CR/LF) right after a non-whitespace character
//           CR/LF right after a letter or digit
//           a digit right after a non-digit
//           a punctuation character (class OTHER, not an apostrophe, not a mark) right after a letter or digit
//   LETTERS (uncased patterns) the three previous characters are letters: a contraction covers at most two
//   W_Y     (cased patterns) the previous character is a lower-case letter and the two before it are letters
//   W_U     (cased patterns) the previous character is a both-sets LETTER (Lo/Lm: CJK, kana, ...): the word is in its
//           [upper] or its [lower] part, which only matters when an upper-case letter follows (pretok_fsm.h S_W_U)
//   W_V     (cased patterns) the three previous characters are letters, the last one upper case: W_X0 or W_XBU, which only
//           matters when the word ends (pretok_fsm.h S_W_V)
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
    if (cased && prevx == X_LO) return S_W_U;
    if (cased && prevx == X_LU && nlet >= 3) return S_W_V;      // (three letters: not inside or right after a contraction suffix)
    if (npun >= 2) return S_ORUN;
    return kNoSync;
}

// evaluate the rule at byte position pos (ps < pos < pe) by decoding up to three characters to the left
template <typename Txt>
CFBPE_HD uint32_t sync_state(const Txt& s, uint64_t pos, uint64_t ps, uint64_t pe, const UcTables& uc, bool cased,
                             uint32_t* prevx_out = nullptr, uint32_t* nlet_out = nullptr, uint32_t* npun_out = nullptr) {
    const uint32_t b = s[pos];
    if ((b & 0xC0) == 0x80) return kNoSync;  // inside a character
    // four ASCII bytes of context: the rule from the class table alone, no decoding (almost every position of Latin-script
    // text ends here, and all lanes run the same few instructions)
    if (pos >= ps + 3) {
        const uint32_t b1 = s[pos - 1], b2 = s[pos - 2], b3 = s[pos - 3];
        if ((b | b1 | b2 | b3) < 0x80u) {
            const uint32_t x = uc.ascii_x[b], p1 = uc.ascii_x[b1], p2 = uc.ascii_x[b2], p3 = uc.ascii_x[b3];
            const uint32_t nlet = x_is_letter(p1) ? (x_is_letter(p2) ? (x_is_letter(p3) ? 3u : 2u) : 1u) : 0u;
            const uint32_t npun = x_is_run_punct(p1, cased) ? (x_is_run_punct(p2, cased) ? 2u : 1u) : 0u;
            if (prevx_out) *prevx_out = p1;
            if (nlet_out) *nlet_out = nlet;
            if (npun_out) *npun_out = npun;
            return sync_rule(x, p1, nlet, npun, cased);
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

// The automaton's real state (and lbe) before pos, for a thread in S_W_U / S_W_V that has to know: walk left to the nearest
// position whose state the class rules give outright (or the prompt start), then run the automaton forward to pos without
// emitting anything.  O(distance), and rare: an upper-case letter inside a run of CJK-like letters.
template <uint32_t kRow = X_COUNT, typename Txt>
CFBPE_HD uint32_t exact_state_before(const Txt& s, uint64_t pos, uint64_t ps, uint64_t pe, const UcTables& uc, const uint16_t* tab, bool cased,
                                     uint64_t* lbe_out = nullptr) {
    uint64_t q = pos, lbe = 0;
    uint32_t state = S_START;
    while (q > ps) {
        q -= get_prev_char(s, q, ps, pe, uc).len;
        if (q == ps) break;
        const uint32_t st = sync_state(s, q, ps, pe, uc, cased);
        if (st != kNoSync && st != S_W_U && st != S_W_V) { state = st; break; }
    }
    while (q < pos) {
        int bad = 0;
        const Ch c = get_char(s, q, pe, uc, &bad);
        const uint32_t a = tab[state * kRow + ext_class(c)];
        if (a & A_CONTR) {
            const uint32_t skip = contraction_bytes(s, q, pe);
            if (skip) { state = S_START; q += skip; continue; }
        }
        state = a & A_STATE_MASK;
        q += c.len ? c.len : 1;
        if (a & A_SET_LBE) lbe = q;
    }
    if (lbe_out) *lbe_out = lbe;
    return state;
}

// The same answer, found the short way when the look-back is simple (it nearly always is): what decides is the character in
// front of the run of upper-case letters (S_W_V) or of both-sets characters (S_W_U) that ends at pos.
//   upper-case run:  W_X0 unless a both-sets character stands in front of it (then the long way)
//   both-sets run:   W_Y if a lower-case letter stands in front of it -- unless an apostrophe one or two characters further
//                    left may make that letter a contraction suffix (then the long way); W_XB0 (lbe = pos) otherwise
template <uint32_t kRow = X_COUNT, typename Txt>
CFBPE_HD uint32_t resolve_word_state(const Txt& s, uint64_t pos, uint64_t ps, uint64_t pe, const UcTables& uc, const uint16_t* tab,
                                     uint64_t* lbe_out) {
    uint64_t q = pos;
    Ch c = get_prev_char(s, q, ps, pe, uc);
    uint32_t x = ext_class(c);
    if (x == X_LU) {
        for (;;) {
            while (q > ps && (static_cast<uint32_t>(s[q - 1]) - 'A') < 26u) --q;      // ASCII capitals: a byte at a time
            if (q == ps) { x = X_EOT; break; }
            c = get_prev_char(s, q, ps, pe, uc);
            x = ext_class(c);
            if (x != X_LU) break;
            q -= c.len;
        }
        if (x != X_LO && x != X_M) { *lbe_out = 0; return S_W_X0; }
        return exact_state_before<kRow>(s, pos, ps, pe, uc, tab, true, lbe_out);
    }
    while (x == X_LO || x == X_M) {
        q -= c.len;
        if (q == ps) { x = X_EOT; break; }
        c = get_prev_char(s, q, ps, pe, uc);
        x = ext_class(c);
    }
    *lbe_out = pos;
    if (x != X_LL) return S_W_XB0;
    uint64_t t = q - c.len;                     // start of the lower-case letter
    if (t > ps) {
        const Ch p1 = get_prev_char(s, t, ps, pe, uc);
        const uint32_t x1 = ext_class(p1);
        if (x1 == X_APOS) return exact_state_before<kRow>(s, pos, ps, pe, uc, tab, true, lbe_out);
        t -= p1.len;
        if (t > ps && x_is_letter(x1) && ext_class(get_prev_char(s, t, ps, pe, uc)) == X_APOS)
            return exact_state_before<kRow>(s, pos, ps, pe, uc, tab, true, lbe_out);
    }
    return S_W_Y;
}

}  // namespace cfbpe
