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

// Cased patterns (o200k, Tekken): inside a run of upper-case or "both-sets" (Lo/Lm/M) characters the automaton's state
// depends on how the word began -- [upper]*[lower]+ | [upper]+[lower]* with the both-sets characters in either part -- so
// the three-character rule above finds no sync point there and ONE thread used to walk a whole CJK sentence or an
// all-caps run.  This resolves the state before pos by looking back over (at most) one run of upper-case characters and
// one run of both-sets characters to the character that decides:
//     ... c0 | both-sets run | upper-case run | pos
//   c0 lower case            -> the both-sets run continued the [lower] part (W_Y); an upper-case character then starts
//                               a new word (W_X0).  (Not certain when an apostrophe sits one or two characters before c0:
//                               c0 may close a contraction suffix, after which the state is START -- no sync then.)
//   c0 anything else / none  -> the both-sets run is in the [upper] part (W_XB0, lbe = its end); upper-case characters
//                               after it give W_XBU (greedy [upper]* may have to give them back: lbe stays)
//   no both-sets run         -> upper-case characters only: W_X0 whatever came before
//   (a both-sets run that begins with a mark right after punctuation is left alone: the mark may belong to the punctuation run)
// Returns kNoSync if the previous character is neither kind, or the look-back exceeds kCasedLookBack characters.
constexpr uint32_t kCasedLookBack = 192;
template <typename Txt>
CFBPE_HD uint32_t cased_word_sync(const Txt& s, uint64_t pos, uint64_t ps, uint64_t pe, const UcTables& uc, uint64_t* lbe_out) {
    uint32_t steps = 0;
    {   // one or two letters right after an apostrophe may be a contraction suffix ('S, 'LL, 'lL ...), which the automaton
        // consumes in one step (state START after it): not a position to start from
        const Ch a1 = get_prev_char(s, pos, ps, pe, uc);
        uint64_t t = pos - a1.len;
        if (t > ps && x_is_letter(ext_class(a1))) {
            const Ch a2 = get_prev_char(s, t, ps, pe, uc);
            const uint32_t x2 = ext_class(a2);
            if (x2 == X_APOS) return kNoSync;
            t -= a2.len;
            if (t > ps && x_is_letter(x2) && ext_class(get_prev_char(s, t, ps, pe, uc)) == X_APOS) return kNoSync;
        }
    }
    uint64_t q = pos;                       // start of the upper-case run that ends at pos
    uint32_t x0 = X_EOT;                    // class of the character before it (X_EOT: the prompt starts there)
    Ch c; c.len = 0;
    while (q > ps) {
        if (++steps > kCasedLookBack) return kNoSync;
        c = get_prev_char(s, q, ps, pe, uc);
        x0 = ext_class(c);
        if (x0 != X_LU) break;
        q -= c.len; x0 = X_EOT;
    }
    const bool has_lu = q < pos;
    if (x0 != X_LO && x0 != X_M) return has_lu ? static_cast<uint32_t>(S_W_X0) : kNoSync;
    uint64_t r = q;                         // start of the both-sets run that ends at q
    uint32_t xd = X_EOT, first = x0;        // first = class of the leftmost character of the run
    Ch d; d.len = 0;
    while (r > ps) {
        if (++steps > kCasedLookBack) return kNoSync;
        d = get_prev_char(s, r, ps, pe, uc);
        xd = ext_class(d);
        if (xd != X_LO && xd != X_M) break;
        r -= d.len; first = xd; xd = X_EOT;
    }
    // a mark right after punctuation is punctuation itself when that was a run ("''M"), a word character after a single
    // prefix character ("'M"): not decided here
    if (first == X_M && (xd == X_OTHER || xd == X_APOS || xd == X_SLASH)) return kNoSync;
    bool lower_part = false;
    if (xd == X_LL) {
        uint64_t t = r - d.len;             // start of c0
        if (t > ps) {
            const Ch p1 = get_prev_char(s, t, ps, pe, uc);
            const uint32_t x1 = ext_class(p1);
            if (x1 == X_APOS) return kNoSync;
            t -= p1.len;
            if (t > ps && x_is_letter(x1)) {
                const Ch p2 = get_prev_char(s, t, ps, pe, uc);
                if (ext_class(p2) == X_APOS) return kNoSync;
            }
        }
        lower_part = true;
    }
    if (!has_lu) {
        if (lower_part) return S_W_Y;
        if (lbe_out) *lbe_out = pos;
        return S_W_XB0;
    }
    if (lower_part) return S_W_X0;
    if (lbe_out) *lbe_out = q;
    return S_W_XBU;
}

// evaluate the rule at byte position pos (ps < pos < pe) by decoding up to three characters to the left
template <typename Txt>
CFBPE_HD uint32_t sync_state(const Txt& s, uint64_t pos, uint64_t ps, uint64_t pe, const UcTables& uc, bool cased,
                             uint32_t* prevx_out = nullptr, uint32_t* nlet_out = nullptr, uint32_t* npun_out = nullptr,
                             uint64_t* lbe_out = nullptr, bool cased_scan = true) {
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
            const uint32_t st = sync_rule(x, p1, nlet, npun, cased);
            if (st != kNoSync || !cased || !cased_scan || p1 != X_LU) return st;
            return cased_word_sync(s, pos, ps, pe, uc, lbe_out);      // an all-caps run
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
    const uint32_t st = sync_rule(ext_class(cur), prevx, nlet, npun, cased);
    if (st != kNoSync || !cased || !cased_scan || !(prevx == X_LU || prevx == X_LO || prevx == X_M)) return st;
    return cased_word_sync(s, pos, ps, pe, uc, lbe_out);
}

}  // namespace cfbpe
