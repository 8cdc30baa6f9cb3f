// pretok.cuh -- the pre-tokenizer as a hand-derived scanner (device + host-inline).
//
// llm-gateway::tokenizer, row a1 of SURVEY.md section 8: "UTF-8 pre-tokenize split".  The
// reference has no implementation (modules/llm-gateway/README.md:51-52); semantics are the
// regex find_iter of the stand-in oracle (tiktoken/_educational.py:30) for the four patterns
// in tiktoken_ext/openai_public.py:89,104-112, the Llama-3 pattern and Tekken's.
//
// Unlike oracle/bpe_oracle.c (patterns as data + generic backtracking matcher) this file
// states each alternative in closed form: one forward pass over the characters of a match,
// no backtracking.  `match_end()` returns where the match that STARTS at p ends.
//
// Character classes come from the Unicode-16 two-stage table (unicode_tables.h).
#pragma once
#include <stdint.h>

#include "tables.h"

namespace cfbpe {

enum : uint32_t { C_OTHER = 0, C_WS = 1, C_CRLF = 2, C_N = 3, C_LU = 4, C_LL = 5, C_LO = 6, C_M = 7 };

struct UcTables {
    const uint8_t* stage1;   // [0x1100]
    const uint8_t* stage2;   // [nblocks*128] nibbles
    const uint8_t* ascii_x;  // [128] extended class of each ASCII byte (pretok_fsm.h)
    const uint16_t* fsm;     // [kNumPatterns * kPretokTableSize] transition tables (pretok_fsm.h)
    // tables of the lane-per-16-bytes split kernel (pretok_ctx.h)
    const uint8_t* cls256;   // [256] byte -> class byte
    const uint16_t* fsm16;   // [kNumPatterns * kFsm16Size] the transition tables with a row stride of 16
    const uint16_t* ctx16;   // [2 * kCtx16Size] the context automaton, per casedness
    const uint64_t* prod;    // [kNumPatterns * kProdMax * 16] the product automaton (split state x context)
    const struct ProdInfo* prod_info;   // [kNumPatterns * kProdMax]
    const uint8_t* prod_skip;           // [kNumPatterns * 2 * kCtxMax]
    const uint8_t* prod_start;          // [kNumPatterns]
};

// Text accessor: any type with operator[](uint64_t) -> byte.  A raw `const uint8_t*` works; K1 passes a view that serves
// the CTA's tile from shared memory (TileText in bpe_kernels.cuh).
struct Ch {
    uint32_t cp;
    uint32_t len;  // 0 = end of prompt
    uint32_t cls;
};

CFBPE_HD uint32_t ascii_class(uint32_t b) {
    if (b - 'a' < 26u) return C_LL;
    if (b - 'A' < 26u) return C_LU;
    if (b - '0' < 10u) return C_N;
    if (b == ' ' || b == '\t' || b == 0x0B || b == 0x0C) return C_WS;
    if (b == '\n' || b == '\r') return C_CRLF;
    return C_OTHER;
}

CFBPE_HD uint32_t uc_class(const UcTables& uc, uint32_t cp) {
    if (cp < 0x80) return ascii_class(cp);
    const uint32_t blk = uc.stage1[cp >> 8];
    const uint32_t byte = uc.stage2[blk * 128 + ((cp & 0xFF) >> 1)];
    return (cp & 1) ? (byte >> 4) : (byte & 0xF);
}

// Strict UTF-8 decode of the character starting at pos (< pe handled by caller: returns len 0 at pe).
// A malformed sequence is reported through *bad and consumed as ONE byte of class OTHER so that
// every scan still terminates; the batch is rejected with CFBPE_EILSEQ afterwards.
template <typename Txt>
CFBPE_HD Ch get_char(const Txt& s, uint64_t pos, uint64_t pe, const UcTables& uc, int* bad) {
    Ch c;
    if (pos >= pe) { c.cp = 0; c.len = 0; c.cls = C_OTHER; return c; }
    const uint32_t b0 = s[pos];
    if (b0 < 0x80) { c.cp = b0; c.len = 1; c.cls = ascii_class(b0); return c; }
    uint32_t need, cp;
    if (b0 >= 0xC2 && b0 <= 0xDF) { need = 1; cp = b0 & 0x1F; }
    else if (b0 >= 0xE0 && b0 <= 0xEF) { need = 2; cp = b0 & 0x0F; }
    else if (b0 >= 0xF0 && b0 <= 0xF4) { need = 3; cp = b0 & 0x07; }
    else { *bad = 1; c.cp = b0; c.len = 1; c.cls = C_OTHER; return c; }
    bool ok = pos + need < pe;
    if (ok) {
        for (uint32_t i = 1; i <= need; ++i) {
            const uint32_t b = s[pos + i];
            ok = ok && ((b & 0xC0) == 0x80);
            cp = (cp << 6) | (b & 0x3F);
        }
        if (need == 2) ok = ok && cp >= 0x800 && !(cp >= 0xD800 && cp <= 0xDFFF);
        if (need == 3) ok = ok && cp >= 0x10000 && cp <= 0x10FFFF;
    }
    if (!ok) { *bad = 1; c.cp = b0; c.len = 1; c.cls = C_OTHER; return c; }
    c.cp = cp; c.len = need + 1; c.cls = uc_class(uc, cp);
    return c;
}

// The character that ENDS at pos (pos > ps).  Steps back over at most three continuation bytes.
template <typename Txt>
CFBPE_HD Ch get_prev_char(const Txt& s, uint64_t pos, uint64_t ps, uint64_t pe, const UcTables& uc) {
    uint64_t q = pos - 1;
    uint32_t back = 0;
    while (q > ps && back < 3 && (s[q] & 0xC0) == 0x80) { --q; ++back; }
    int bad = 0;
    Ch c = get_char(s, q, pe, uc, &bad);
    if (bad || q + c.len != pos) {  // malformed neighbourhood: treat the single previous byte as OTHER
        c.cp = s[pos - 1]; c.len = 1; c.cls = C_OTHER;
    }
    return c;
}

CFBPE_HD bool is_ws(uint32_t k) { return k == C_WS || k == C_CRLF; }
CFBPE_HD bool is_letter(uint32_t k) { return k >= C_LU && k <= C_LO; }
CFBPE_HD bool is_punct(uint32_t k) { return k == C_OTHER || k == C_M; }        // [^\s\p{L}\p{N}]
CFBPE_HD bool is_prefix(uint32_t k) { return k == C_OTHER || k == C_WS || k == C_M; }  // [^\r\n\p{L}\p{N}]
CFBPE_HD bool in_upper_set(uint32_t k) { return k == C_LU || k == C_LO || k == C_M; }  // [\p{Lu}\p{Lt}\p{Lm}\p{Lo}\p{M}]
CFBPE_HD bool in_lower_set(uint32_t k) { return k == C_LL || k == C_LO || k == C_M; }  // [\p{Ll}\p{Lm}\p{Lo}\p{M}]
CFBPE_HD bool in_both_sets(uint32_t k) { return k == C_LO || k == C_M; }

// pattern traits
struct PatTraits {
    bool cased;         // o200k / tekken word alternatives
    bool contr_prefix;  // cl100k / llama3: contraction is its own alternative
    bool contr_suffix;  // o200k: optional contraction after a word
    bool slash_trailer; // o200k / tekken: [\r\n/]* after a punctuation run
    bool ws_eot;        // cl100k: \s++$
    uint32_t max_digits;
};
CFBPE_HD PatTraits pat_traits(uint32_t pat) {
    PatTraits t;
    t.cased = (pat == 1 || pat == 3);
    t.contr_prefix = (pat == 0 || pat == 2);
    t.contr_suffix = (pat == 1);
    t.slash_trailer = (pat == 1 || pat == 3);
    t.ws_eot = (pat == 0);
    t.max_digits = (pat == 3) ? 1u : 3u;
    return t;
}

// bytes of a contraction ('s 't 'm 'd 'll 've 're, any case, U+017F counts as s) at pos, else 0
template <typename Txt>
CFBPE_HD uint32_t contraction_bytes(const Txt& s, uint64_t pos, uint64_t pe) {
    if (pos + 1 >= pe || s[pos] != '\'') return 0;
    const uint32_t a = s[pos + 1] | 0x20;  // ASCII lower-case fold (only valid for letters; checked below)
    const bool a_letter = (a - 'a') < 26u;
    if (a_letter && (a == 's' || a == 'd' || a == 'm' || a == 't')) return 2;
    if (s[pos + 1] == 0xC5 && pos + 2 < pe && s[pos + 2] == 0xBF) return 3;  // 'ſ
    if (a_letter && pos + 2 < pe) {
        const uint32_t b = s[pos + 2] | 0x20;
        if ((a == 'l' && b == 'l') || (a == 'v' && b == 'e') || (a == 'r' && b == 'e')) return 3;
    }
    return 0;
}

// cased word, first form:  [upper-set]* [lower-set]+   (with the backtracking result in closed form)
CFBPE_HD uint64_t word_a(const uint8_t* __restrict__ s, uint64_t q, uint64_t pe, const UcTables& uc, int* bad) {
    uint64_t x = q, last_both_end = 0;
    Ch c = get_char(s, x, pe, uc, bad);
    while (c.len && in_upper_set(c.cls)) {
        x += c.len;
        if (in_both_sets(c.cls)) last_both_end = x;
        c = get_char(s, x, pe, uc, bad);
    }
    uint64_t y = x;
    while (c.len && in_lower_set(c.cls)) { y += c.len; c = get_char(s, y, pe, uc, bad); }
    if (y > x) return y;
    return last_both_end;  // greedy [upper-set]* gives characters back until a both-sets one can be the [lower-set]+
}
// cased word, second form: [upper-set]+ [lower-set]*
CFBPE_HD uint64_t word_b(const uint8_t* __restrict__ s, uint64_t q, uint64_t pe, const UcTables& uc, int* bad) {
    uint64_t x = q;
    Ch c = get_char(s, x, pe, uc, bad);
    while (c.len && in_upper_set(c.cls)) { x += c.len; c = get_char(s, x, pe, uc, bad); }
    if (x == q) return 0;
    while (c.len && in_lower_set(c.cls)) { x += c.len; c = get_char(s, x, pe, uc, bad); }
    return x;
}

// End of the match that starts at p (p < pe).  [p, result) is one piece.
CFBPE_HD uint64_t match_end(const uint8_t* __restrict__ s, uint64_t p, uint64_t pe, uint32_t pat, const UcTables& uc, int* bad) {
    const PatTraits T = pat_traits(pat);
    const Ch c0 = get_char(s, p, pe, uc, bad);
    // --- contraction as its own alternative
    if (T.contr_prefix && c0.cp == '\'') {
        const uint32_t cl = contraction_bytes(s, p, pe);
        if (cl) return p + cl;
    }
    // --- words, optionally led by one non-letter, non-digit, non-CR/LF character
    const bool pre = is_prefix(c0.cls);
    if (!T.cased) {
        uint64_t q = pre ? p + c0.len : p;
        Ch c = pre ? get_char(s, q, pe, uc, bad) : c0;
        if (c.len && is_letter(c.cls)) {
            do { q += c.len; c = get_char(s, q, pe, uc, bad); } while (c.len && is_letter(c.cls));
            return q;
        }
    } else {
        const bool self_word = is_letter(c0.cls) || c0.cls == C_M;
        uint64_t e = 0;
        if (pre) e = word_a(s, p + c0.len, pe, uc, bad);
        if (!e && self_word) e = word_a(s, p, pe, uc, bad);
        if (!e && pre) e = word_b(s, p + c0.len, pe, uc, bad);
        if (!e && self_word) e = word_b(s, p, pe, uc, bad);
        if (e) {
            if (T.contr_suffix) e += contraction_bytes(s, e, pe);
            return e;
        }
    }
    // --- digits
    if (c0.cls == C_N) {
        uint64_t q = p + c0.len;
        for (uint32_t cnt = 1; cnt < T.max_digits; ++cnt) {
            const Ch c = get_char(s, q, pe, uc, bad);
            if (!c.len || c.cls != C_N) break;
            q += c.len;
        }
        return q;
    }
    // --- punctuation run, optional leading space, trailing newlines (and '/')
    {
        uint64_t q = p;
        Ch c = c0;
        if (c0.cp == ' ') { q = p + 1; c = get_char(s, q, pe, uc, bad); }
        if (c.len && is_punct(c.cls)) {
            do { q += c.len; c = get_char(s, q, pe, uc, bad); } while (c.len && is_punct(c.cls));
            while (c.len && (c.cls == C_CRLF || (T.slash_trailer && c.cp == '/'))) { q += c.len; c = get_char(s, q, pe, uc, bad); }
            return q;
        }
    }
    // --- whitespace run starting at p (c0 is whitespace here)
    {
        uint64_t q = p, last_crlf_end = 0, last_start = p;
        Ch c = c0;
        while (c.len && is_ws(c.cls)) {
            if (c.cls == C_CRLF) last_crlf_end = q + c.len;
            last_start = q;
            q += c.len;
            c = get_char(s, q, pe, uc, bad);
        }
        if (q == p) return p + (c0.len ? c0.len : 1);  // unreachable for valid classes; guarantees progress
        if (q == pe) return (T.ws_eot || !last_crlf_end) ? q : last_crlf_end;
        if (last_crlf_end) return last_crlf_end;      // \s*[\r\n]+ : through the last newline of the run
        if (last_start > p) return last_start;        // \s+(?!\S) : all but the last whitespace character
        return q;                                     // \s
    }
}

}  // namespace cfbpe
