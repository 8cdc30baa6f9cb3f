/*
 * bpe_oracle.c -- CPU ORACLE for the batched BPE encode path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (cyberfabric-core_b200/csrc) never links,
 * imports or calls it.
 *
 * PARITY STATUS: "parity unpinned by the reference".  The reference tree
 * (cyberfabric/cyberfabric-core @ 2026-02-20) ships no tokenizer: modules/llm-gateway
 * is spec-only (modules/llm-gateway/README.md:51-52 "planned"), Cargo.lock pins no BPE
 * crate.  This file therefore restates the published tiktoken
 * algorithm (tiktoken 0.12.0, the stand-in oracle named in SURVEY.md section 8(c)) and is
 * pinned against golden vectors produced by that engine in this container
 * (tests/golden/, generator tools/gen_golden.py).
 *
 * What is restated, and from where:
 *   - split:  regex find_iter over the text, alternatives tried in order with
 *             backtracking (leftmost-first), tiktoken/_educational.py:30 and the
 *             pattern strings in tiktoken_ext/openai_public.py:89,104-112 (cl100k,
 *             o200k), the Llama-3 pattern, and tekken_240911.json["config"]["pattern"].
 *             The patterns are written below as data for a small generic backtracking
 *             matcher, so the code path is the regex semantics, not a hand-derived
 *             state machine (the product uses the hand-derived form; the two are
 *             compared differentially).
 *   - merge:  minimum-rank adjacent pair, leftmost on ties, until no adjacent pair
 *             concatenates to a vocabulary entry: tiktoken/_educational.py:83-116;
 *             whole-piece shortcut `if piece in ranks` as CoreBPE.encode_ordinary does
 *             (tiktoken/core.py:63-77 calls into it).
 *   - vocab:  ".tiktoken" text format, base64 token, space, decimal rank per line:
 *             tiktoken/load.py:160-172.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "unicode_ranges.h"

#if defined(__GNUC__)
#define ORACLE_API __attribute__((visibility("default")))
#else
#define ORACLE_API
#endif

/* ------------------------------------------------------------------------- */
/* Unicode features                                                          */
/* ------------------------------------------------------------------------- */
enum {
    F_WS = 1, F_CRLF = 2, F_L = 4, F_N = 8, F_LU = 16, F_LL = 32, F_LO = 64, F_M = 128,
    F_SPACE = 256, F_SLASH = 512, F_APOS = 1024
};

static uint16_t g_bmp_feat[0x10000];
static pthread_once_t g_feat_once = PTHREAD_ONCE_INIT;

static unsigned feat_slow(uint32_t cp) {
    unsigned f = 0;
    int lo = 0, hi = UC_NRANGES - 1;
    while (lo <= hi) {
        int mid = (lo + hi) >> 1;
        if (cp < uc_ranges[mid].lo) hi = mid - 1;
        else if (cp > uc_ranges[mid].hi) lo = mid + 1;
        else {
            switch (uc_ranges[mid].cls) {
            case 1: f = F_WS; break;
            case 2: f = F_WS | F_CRLF; break;
            case 3: f = F_N; break;
            case 4: f = F_L | F_LU; break;
            case 5: f = F_L | F_LL; break;
            case 6: f = F_L | F_LO; break;
            case 7: f = F_M; break;
            default: break;
            }
            break;
        }
    }
    if (cp == ' ') f |= F_SPACE;
    if (cp == '/') f |= F_SLASH;
    if (cp == '\'') f |= F_APOS;
    return f;
}

static void feat_init(void) {
    for (uint32_t c = 0; c < 0x10000; c++) g_bmp_feat[c] = (uint16_t)feat_slow(c);
}

static inline unsigned feat(uint32_t cp) {
    return cp < 0x10000 ? g_bmp_feat[cp] : feat_slow(cp);
}

/* strict UTF-8 decode (what a Rust &str / Python str.encode() can contain).
 * returns number of chars, or -1 on malformed input. cp[] and off[] need n+1 slots. */
static long utf8_decode(const uint8_t *s, size_t n, uint32_t *cp, uint32_t *off) {
    size_t i = 0;
    long k = 0;
    while (i < n) {
        uint8_t b = s[i];
        uint32_t c;
        size_t len;
        if (b < 0x80) { c = b; len = 1; }
        else if (b >= 0xC2 && b <= 0xDF) { c = b & 0x1F; len = 2; }
        else if (b >= 0xE0 && b <= 0xEF) { c = b & 0x0F; len = 3; }
        else if (b >= 0xF0 && b <= 0xF4) { c = b & 0x07; len = 4; }
        else return -1;
        if (i + len > n) return -1;
        for (size_t j = 1; j < len; j++) {
            if ((s[i + j] & 0xC0) != 0x80) return -1;
            c = (c << 6) | (s[i + j] & 0x3F);
        }
        if (len == 3 && (c < 0x800 || (c >= 0xD800 && c <= 0xDFFF))) return -1;
        if (len == 4 && (c < 0x10000 || c > 0x10FFFF)) return -1;
        cp[k] = c;
        off[k] = (uint32_t)i;
        k++;
        i += len;
    }
    off[k] = (uint32_t)n;
    return k;
}

/* ------------------------------------------------------------------------- */
/* Patterns as data + generic backtracking matcher                           */
/* ------------------------------------------------------------------------- */
enum { E_END = 0, E_CLASS, E_CONTR, E_NLA_NONWS, E_EOT };
#define INF 0x7fffffff

typedef struct {
    int kind;
    unsigned mask; /* feature mask */
    int neg;       /* class is negated */
    int min, max;  /* repetition */
    int possessive;
} elem;

#define CLS(mask, neg, mn, mx, poss) { E_CLASS, (mask), (neg), (mn), (mx), (poss) }
#define CONTR(optional) { E_CONTR, 0, 0, (optional) ? 0 : 1, 1, 0 }
#define NLA { E_NLA_NONWS, 0, 0, 0, 0, 0 }
#define EOT { E_EOT, 0, 0, 0, 0, 0 }
#define END { E_END, 0, 0, 0, 0, 0 }

#define MAX_ELEMS 6
#define MAX_ALTS 8
typedef struct { int nalts; elem alt[MAX_ALTS][MAX_ELEMS]; } pattern;

/* pattern ids are shared with include/cfbpe.h (CFBPE_PATTERN_*) */
enum { PAT_CL100K = 0, PAT_O200K = 1, PAT_LLAMA3 = 2, PAT_TEKKEN = 3, PAT_COUNT = 4 };

static const pattern g_patterns[PAT_COUNT] = {
    /* cl100k_base: tiktoken_ext/openai_public.py:89
     * '(?i:[sdmt]|ll|ve|re)|[^\r\n\p{L}\p{N}]?+\p{L}++|\p{N}{1,3}+| ?[^\s\p{L}\p{N}]++[\r\n]*+|\s++$|\s*[\r\n]|\s+(?!\S)|\s */
    { 8, {
        { CONTR(0), END },
        { CLS(F_CRLF | F_L | F_N, 1, 0, 1, 1), CLS(F_L, 0, 1, INF, 1), END },
        { CLS(F_N, 0, 1, 3, 1), END },
        { CLS(F_SPACE, 0, 0, 1, 0), CLS(F_WS | F_L | F_N, 1, 1, INF, 1), CLS(F_CRLF, 0, 0, INF, 1), END },
        { CLS(F_WS, 0, 1, INF, 1), EOT, END },
        { CLS(F_WS, 0, 0, INF, 0), CLS(F_CRLF, 0, 1, 1, 0), END },
        { CLS(F_WS, 0, 1, INF, 0), NLA, END },
        { CLS(F_WS, 0, 1, 1, 0), END },
    } },
    /* o200k_base: tiktoken_ext/openai_public.py:104-112 */
    { 7, {
        { CLS(F_CRLF | F_L | F_N, 1, 0, 1, 0), CLS(F_LU | F_LO | F_M, 0, 0, INF, 0),
          CLS(F_LL | F_LO | F_M, 0, 1, INF, 0), CONTR(1), END },
        { CLS(F_CRLF | F_L | F_N, 1, 0, 1, 0), CLS(F_LU | F_LO | F_M, 0, 1, INF, 0),
          CLS(F_LL | F_LO | F_M, 0, 0, INF, 0), CONTR(1), END },
        { CLS(F_N, 0, 1, 3, 0), END },
        { CLS(F_SPACE, 0, 0, 1, 0), CLS(F_WS | F_L | F_N, 1, 1, INF, 0), CLS(F_CRLF | F_SLASH, 0, 0, INF, 0), END },
        { CLS(F_WS, 0, 0, INF, 0), CLS(F_CRLF, 0, 1, INF, 0), END },
        { CLS(F_WS, 0, 1, INF, 0), NLA, END },
        { CLS(F_WS, 0, 1, INF, 0), END },
    } },
    /* llama3 (public Llama-3 tokenizer pattern):
     * (?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+ */
    { 7, {
        { CONTR(0), END },
        { CLS(F_CRLF | F_L | F_N, 1, 0, 1, 0), CLS(F_L, 0, 1, INF, 0), END },
        { CLS(F_N, 0, 1, 3, 0), END },
        { CLS(F_SPACE, 0, 0, 1, 0), CLS(F_WS | F_L | F_N, 1, 1, INF, 0), CLS(F_CRLF, 0, 0, INF, 0), END },
        { CLS(F_WS, 0, 0, INF, 0), CLS(F_CRLF, 0, 1, INF, 0), END },
        { CLS(F_WS, 0, 1, INF, 0), NLA, END },
        { CLS(F_WS, 0, 1, INF, 0), END },
    } },
    /* Mistral Tekken: tekken_240911.json["config"]["pattern"] (o200k form, no contractions, single \p{N}) */
    { 7, {
        { CLS(F_CRLF | F_L | F_N, 1, 0, 1, 0), CLS(F_LU | F_LO | F_M, 0, 0, INF, 0),
          CLS(F_LL | F_LO | F_M, 0, 1, INF, 0), END },
        { CLS(F_CRLF | F_L | F_N, 1, 0, 1, 0), CLS(F_LU | F_LO | F_M, 0, 1, INF, 0),
          CLS(F_LL | F_LO | F_M, 0, 0, INF, 0), END },
        { CLS(F_N, 0, 1, 1, 0), END },
        { CLS(F_SPACE, 0, 0, 1, 0), CLS(F_WS | F_L | F_N, 1, 1, INF, 0), CLS(F_CRLF | F_SLASH, 0, 0, INF, 0), END },
        { CLS(F_WS, 0, 0, INF, 0), CLS(F_CRLF, 0, 1, INF, 0), END },
        { CLS(F_WS, 0, 1, INF, 0), NLA, END },
        { CLS(F_WS, 0, 1, INF, 0), END },
    } },
};

typedef struct { const uint32_t *cp; long n; } text_t;

static inline int class_has(const elem *e, uint32_t c) {
    int in = (feat(c) & e->mask) != 0;
    return in ^ e->neg;
}

/* case-insensitive letter compare under the regex engine's simple case folding:
 * the only non-ASCII code point that folds into [sdmtlvre] is U+017F (long s). */
static inline int ci_eq(uint32_t c, char lower) {
    if (c == (uint32_t)lower || c == (uint32_t)(lower - 32)) return 1;
    return lower == 's' && c == 0x17F;
}

/* length in chars of a contraction ('s 't 'm 'd 'll 've 're) at pos, or 0 */
static long contraction_len(const text_t *t, long pos) {
    if (pos >= t->n || t->cp[pos] != '\'') return 0;
    if (pos + 1 < t->n) {
        uint32_t a = t->cp[pos + 1];
        if (ci_eq(a, 's') || ci_eq(a, 'd') || ci_eq(a, 'm') || ci_eq(a, 't')) return 2;
        if (pos + 2 < t->n) {
            uint32_t b = t->cp[pos + 2];
            if (ci_eq(a, 'l') && ci_eq(b, 'l')) return 3;
            if (ci_eq(a, 'v') && ci_eq(b, 'e')) return 3;
            if (ci_eq(a, 'r') && ci_eq(b, 'e')) return 3;
        }
    }
    return 0;
}

/* backtracking match of elements e[0..] at pos; returns end position or -1 */
static long match_seq(const text_t *t, const elem *e, long pos) {
    switch (e->kind) {
    case E_END:
        return pos;
    case E_EOT:
        return pos == t->n ? match_seq(t, e + 1, pos) : -1;
    case E_NLA_NONWS: /* (?!\S): fails iff a non-whitespace char follows */
        if (pos < t->n && !(feat(t->cp[pos]) & F_WS)) return -1;
        return match_seq(t, e + 1, pos);
    case E_CONTR: {
        long k = contraction_len(t, pos);
        if (k > 0) {
            long r = match_seq(t, e + 1, pos + k);
            if (r >= 0) return r;
        }
        if (e->min == 0) return match_seq(t, e + 1, pos); /* optional suffix */
        return -1;
    }
    case E_CLASS: {
        long k = 0;
        while (k < e->max && pos + k < t->n && class_has(e, t->cp[pos + k])) k++;
        if (k < e->min) return -1;
        if (e->possessive) return match_seq(t, e + 1, pos + k);
        for (; k >= e->min; k--) { /* greedy, give back one at a time */
            long r = match_seq(t, e + 1, pos + k);
            if (r >= 0) return r;
        }
        return -1;
    }
    }
    return -1;
}

/* one find_iter step: leftmost match at or after pos, first alternative that matches.
 * returns 1 and [*ms,*me) or 0 when no further match */
static int find_next(const pattern *p, const text_t *t, long pos, long *ms, long *me) {
    for (; pos < t->n; pos++) {
        for (int a = 0; a < p->nalts; a++) {
            long r = match_seq(t, p->alt[a], pos);
            if (r > pos) { *ms = pos; *me = r; return 1; }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Vocabulary: bytes -> rank                                                 */
/* ------------------------------------------------------------------------- */
typedef struct {
    uint32_t n_ranks;
    uint8_t *blob;      /* concatenated token bytes */
    uint32_t *tok_off;  /* n_ranks + 1 */
    uint32_t cap;       /* hash capacity (pow2) */
    uint32_t *slots;    /* rank+1, 0 = empty */
    uint32_t max_len;
} oracle_vocab;

static inline uint64_t fnv1a(const uint8_t *p, size_t n) {
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; }
    return h ^ (h >> 29);
}

#define NO_RANK 0xFFFFFFFFu

static inline uint32_t vocab_get(const oracle_vocab *v, const uint8_t *p, size_t n) {
    if (n > v->max_len) return NO_RANK;
    uint32_t m = v->cap - 1;
    uint32_t i = (uint32_t)fnv1a(p, n) & m;
    for (;;) {
        uint32_t s = v->slots[i];
        if (!s) return NO_RANK;
        uint32_t r = s - 1;
        uint32_t len = v->tok_off[r + 1] - v->tok_off[r];
        if (len == n && memcmp(v->blob + v->tok_off[r], p, n) == 0) return r;
        i = (i + 1) & m;
    }
}

static int b64val(int c) {
    if (c >= 'A' && c <= 'Z') return c - 'A';
    if (c >= 'a' && c <= 'z') return c - 'a' + 26;
    if (c >= '0' && c <= '9') return c - '0' + 52;
    if (c == '+') return 62;
    if (c == '/') return 63;
    return -1;
}

ORACLE_API void oracle_vocab_free(oracle_vocab *v) {
    if (!v) return;
    free(v->blob); free(v->tok_off); free(v->slots); free(v);
}

/* parse ".tiktoken" text (tiktoken/load.py:160-172); ranks must be 0..n-1 in file order
 * (true for every published file); keep only ranks < max_ranks (0 = all). */
ORACLE_API oracle_vocab *oracle_vocab_load(const uint8_t *file, size_t len, uint32_t max_ranks) {
    pthread_once(&g_feat_once, feat_init);
    oracle_vocab *v = calloc(1, sizeof *v);
    size_t nlines = 0;
    for (size_t i = 0; i < len; i++) nlines += file[i] == '\n';
    nlines += 1;
    v->blob = malloc(len);          /* decoded bytes never exceed the text size */
    v->tok_off = malloc((nlines + 1) * sizeof(uint32_t));
    size_t pos = 0, bo = 0;
    uint32_t n = 0;
    while (pos < len) {
        size_t eol = pos;
        while (eol < len && file[eol] != '\n') eol++;
        if (eol > pos) {
            size_t sp = pos;
            while (sp < eol && file[sp] != ' ') sp++;
            if (sp == eol) goto bad;
            /* base64 */
            uint32_t acc = 0; int bits = 0;
            v->tok_off[n] = (uint32_t)bo;
            for (size_t i = pos; i < sp; i++) {
                if (file[i] == '=') break;
                int d = b64val(file[i]);
                if (d < 0) goto bad;
                acc = (acc << 6) | (uint32_t)d; bits += 6;
                if (bits >= 8) { bits -= 8; v->blob[bo++] = (uint8_t)(acc >> bits); acc &= (1u << bits) - 1; }
            }
            unsigned long rank = strtoul((const char *)file + sp + 1, NULL, 10);
            if (rank != n) goto bad;
            size_t tl = bo - v->tok_off[n];
            if (tl == 0) goto bad;
            if (max_ranks && n >= max_ranks) { bo = v->tok_off[n]; break; }
            if (tl > v->max_len) v->max_len = (uint32_t)tl;
            n++;
        }
        pos = eol + 1;
    }
    v->tok_off[n] = (uint32_t)bo;
    v->n_ranks = n;
    v->cap = 1;
    while (v->cap < 2 * n + 16) v->cap <<= 1;
    v->slots = calloc(v->cap, sizeof(uint32_t));
    for (uint32_t r = 0; r < n; r++) {
        const uint8_t *p = v->blob + v->tok_off[r];
        size_t tl = v->tok_off[r + 1] - v->tok_off[r];
        uint32_t m = v->cap - 1, i = (uint32_t)fnv1a(p, tl) & m;
        while (v->slots[i]) {
            uint32_t q = v->slots[i] - 1;
            if (v->tok_off[q + 1] - v->tok_off[q] == tl && memcmp(v->blob + v->tok_off[q], p, tl) == 0) goto bad; /* dup */
            i = (i + 1) & m;
        }
        v->slots[i] = r + 1;
    }
    return v;
bad:
    oracle_vocab_free(v);
    return NULL;
}

ORACLE_API uint32_t oracle_vocab_size(const oracle_vocab *v) { return v->n_ranks; }

/* ------------------------------------------------------------------------- */
/* Merge loop (tiktoken/_educational.py:83-116 semantics)                    */
/* ------------------------------------------------------------------------- */
typedef struct { uint32_t *start; uint32_t *rank; size_t cap; } merge_scratch;

static void scratch_reserve(merge_scratch *s, size_t n) {
    if (s->cap >= n) return;
    s->cap = n * 2 + 64;
    s->start = realloc(s->start, s->cap * sizeof(uint32_t));
    s->rank = realloc(s->rank, s->cap * sizeof(uint32_t));
}

/* encode one piece; appends ids to out; returns count */
static size_t encode_piece(const oracle_vocab *v, const uint8_t *p, size_t n, uint32_t *out, merge_scratch *s) {
    uint32_t whole = vocab_get(v, p, n);
    if (whole != NO_RANK) { out[0] = whole; return 1; }   /* whole-piece shortcut */
    if (n == 1) { out[0] = NO_RANK; return 1; }            /* cannot happen: all 256 bytes are ranks */
    scratch_reserve(s, n + 2);
    /* parts i = [start[i], start[i+1]); rank[i] = rank of parts i,i+1 concatenated */
    size_t np = n;
    for (size_t i = 0; i < n; i++) s->start[i] = (uint32_t)i;
    s->start[n] = (uint32_t)n;
    for (size_t i = 0; i + 1 < n; i++) s->rank[i] = vocab_get(v, p + i, 2);
    s->rank[n - 1] = NO_RANK;
    for (;;) {
        uint32_t best = NO_RANK; size_t bi = 0;
        for (size_t i = 0; i + 1 < np; i++)
            if (s->rank[i] < best) { best = s->rank[i]; bi = i; }   /* strict < : leftmost on ties */
        if (best == NO_RANK) break;
        /* merge parts bi and bi+1 */
        memmove(s->start + bi + 1, s->start + bi + 2, (np - bi - 1) * sizeof(uint32_t));
        memmove(s->rank + bi + 1, s->rank + bi + 2, (np - bi - 2) * sizeof(uint32_t));
        np--;
        s->rank[bi] = (bi + 1 < np) ? vocab_get(v, p + s->start[bi], s->start[bi + 2] - s->start[bi]) : NO_RANK;
        if (bi > 0) s->rank[bi - 1] = vocab_get(v, p + s->start[bi - 1], s->start[bi + 1] - s->start[bi - 1]);
    }
    for (size_t i = 0; i < np; i++) out[i] = vocab_get(v, p + s->start[i], s->start[i + 1] - s->start[i]);
    return np;
}

/* ------------------------------------------------------------------------- */
/* Public entry points                                                       */
/* ------------------------------------------------------------------------- */
typedef struct { uint32_t *cp; uint32_t *off; size_t cap; merge_scratch ms; } work_t;

static void work_reserve(work_t *w, size_t n) {
    if (w->cap >= n + 1) return;
    w->cap = n + 1 + 256;
    w->cp = realloc(w->cp, w->cap * sizeof(uint32_t));
    w->off = realloc(w->off, w->cap * sizeof(uint32_t));
}
static void work_free(work_t *w) { free(w->cp); free(w->off); free(w->ms.start); free(w->ms.rank); }

/* split only: writes the byte offset of the END of each piece; returns number of pieces, -1 on bad UTF-8.
 * piece_ends needs room for `len` entries. */
ORACLE_API long oracle_split(int pattern_id, const uint8_t *text, size_t len, uint32_t *piece_ends) {
    pthread_once(&g_feat_once, feat_init);
    if (pattern_id < 0 || pattern_id >= PAT_COUNT) return -2;
    work_t w = {0};
    work_reserve(&w, len);
    long nc = utf8_decode(text, len, w.cp, w.off);
    if (nc < 0) { work_free(&w); return -1; }
    text_t t = { w.cp, nc };
    long pos = 0, ms, me, np = 0;
    while (find_next(&g_patterns[pattern_id], &t, pos, &ms, &me)) {
        if (ms != pos) { work_free(&w); return -3; } /* gap: cannot happen for these patterns */
        piece_ends[np++] = w.off[me];
        pos = me;
    }
    work_free(&w);
    return np;
}

static long encode_one(const oracle_vocab *v, int pattern_id, const uint8_t *text, size_t len, uint32_t *out, work_t *w) {
    work_reserve(w, len);
    long nc = utf8_decode(text, len, w->cp, w->off);
    if (nc < 0) return -1;
    text_t t = { w->cp, nc };
    long pos = 0, ms, me;
    size_t no = 0;
    while (find_next(&g_patterns[pattern_id], &t, pos, &ms, &me)) {
        no += encode_piece(v, text + w->off[ms], w->off[me] - w->off[ms], out + no, &w->ms);
        pos = me;
    }
    return (long)no;
}

/* encode_ordinary of one text; out needs room for len ids; returns count or -1 (bad UTF-8) */
ORACLE_API long oracle_encode(const oracle_vocab *v, int pattern_id, const uint8_t *text, size_t len, uint32_t *out) {
    if (pattern_id < 0 || pattern_id >= PAT_COUNT) return -2;
    work_t w = {0};
    long r = encode_one(v, pattern_id, text, len, out, &w);
    work_free(&w);
    return r;
}

/* ---- batch driver: a persistent pool of worker threads (created once, parked on a condition variable between calls),
 * a scratch id buffer that is kept across calls (a fresh 4-bytes-per-input-byte malloc was page-faulted by every call), and a
 * parallel second phase that moves every prompt's ids to their final place.  This is the CPU baseline of record of bench.py
 * (`cpu_baseline`, `--impl reference`): the port should lose to the GPU because of the hardware, not because of its driver. */
typedef struct {
    const oracle_vocab *const *vocabs; const int *patterns; const uint8_t *vocab_ids;
    uint32_t n; const uint8_t *bytes; const uint64_t *offsets;
    uint32_t *tmp; uint32_t *counts; uint32_t *out_ids; const uint64_t *out_offsets;
    volatile long next; volatile int err; int phase;
} batch_job;

static void batch_phase(batch_job *j, work_t *w) {
    for (;;) {
        long i0 = __sync_fetch_and_add(&j->next, 16);
        if (i0 >= (long)j->n) break;
        long i1 = i0 + 16 < (long)j->n ? i0 + 16 : (long)j->n;
        for (long i = i0; i < i1; i++) {
            if (j->phase == 0) {
                unsigned vid = j->vocab_ids ? j->vocab_ids[i] : 0;
                long c = encode_one(j->vocabs[vid], j->patterns[vid], j->bytes + j->offsets[i],
                                    (size_t)(j->offsets[i + 1] - j->offsets[i]), j->tmp + j->offsets[i], w);
                if (c < 0) { j->err = 1; c = 0; }
                j->counts[i] = (uint32_t)c;
            } else {
                memcpy(j->out_ids + j->out_offsets[i], j->tmp + j->offsets[i], (size_t)j->counts[i] * sizeof(uint32_t));
            }
        }
    }
}

#define POOL_MAX 1024
static struct {
    pthread_mutex_t mu; pthread_cond_t go, done;
    pthread_t th[POOL_MAX]; int n_threads;      /* workers created so far */
    batch_job *job; int want;                   /* current job and how many workers should take part */
    unsigned long gen; int running;             /* generation counter; workers still inside the current phase */
    uint32_t *tmp; uint64_t tmp_cap;            /* scratch ids, kept across calls */
    pthread_mutex_t call_mu;                    /* one batch call at a time */
} g_pool = { PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, PTHREAD_COND_INITIALIZER, {0}, 0, NULL, 0, 0, 0, NULL, 0, PTHREAD_MUTEX_INITIALIZER };

static void *pool_worker(void *arg) {
    const int id = (int)(intptr_t)arg;
    work_t w = {0};
    unsigned long seen = 0;
    pthread_mutex_lock(&g_pool.mu);
    for (;;) {
        while (g_pool.gen == seen) pthread_cond_wait(&g_pool.go, &g_pool.mu);
        seen = g_pool.gen;
        if (id >= g_pool.want) continue;        /* this phase runs on fewer threads */
        batch_job *j = g_pool.job;
        pthread_mutex_unlock(&g_pool.mu);
        batch_phase(j, &w);
        pthread_mutex_lock(&g_pool.mu);
        if (--g_pool.running == 0) pthread_cond_signal(&g_pool.done);
    }
    return NULL;
}

static void pool_run(batch_job *j, int nthreads, work_t *w) {      /* the caller is thread 0 */
    pthread_mutex_lock(&g_pool.mu);
    while (g_pool.n_threads < nthreads - 1 && g_pool.n_threads < POOL_MAX) {
        if (pthread_create(&g_pool.th[g_pool.n_threads], NULL, pool_worker, (void *)(intptr_t)g_pool.n_threads)) break;
        g_pool.n_threads++;
    }
    const int helpers = nthreads - 1 < g_pool.n_threads ? nthreads - 1 : g_pool.n_threads;
    g_pool.job = j; g_pool.want = helpers; g_pool.running = helpers; g_pool.gen++;
    pthread_cond_broadcast(&g_pool.go);
    pthread_mutex_unlock(&g_pool.mu);
    batch_phase(j, w);
    pthread_mutex_lock(&g_pool.mu);
    while (g_pool.running) pthread_cond_wait(&g_pool.done, &g_pool.mu);
    pthread_mutex_unlock(&g_pool.mu);
}

/* Batch encode over a packed prompt buffer with nthreads host threads.
 * vocabs/patterns are indexed by vocab_ids[i] (NULL vocab_ids = all 0).
 * out_ids needs room for offsets[n] ids (ids never outnumber bytes); out_offsets has n+1 entries.
 * returns 0, or -1 if any prompt held malformed UTF-8. */
ORACLE_API int oracle_encode_batch(const oracle_vocab *const *vocabs, const int *patterns, const uint8_t *vocab_ids,
                                   uint32_t n, const uint8_t *bytes, const uint64_t *offsets,
                                   uint32_t *out_ids, uint64_t *out_offsets, uint32_t *out_counts, int nthreads) {
    pthread_once(&g_feat_once, feat_init);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > POOL_MAX) nthreads = POOL_MAX;
    pthread_mutex_lock(&g_pool.call_mu);
    const uint64_t total = offsets[n];
    if (g_pool.tmp_cap < total + 1) {
        free(g_pool.tmp);
        g_pool.tmp_cap = total + 1 + total / 8;
        g_pool.tmp = malloc(g_pool.tmp_cap * sizeof(uint32_t));
    }
    batch_job job = { vocabs, patterns, vocab_ids, n, bytes, offsets, g_pool.tmp, out_counts, out_ids, out_offsets, 0, 0, 0 };
    work_t w = {0};
    pool_run(&job, nthreads, &w);
    uint64_t acc = 0;
    for (uint32_t i = 0; i < n; i++) { out_offsets[i] = acc; acc += out_counts[i]; }
    out_offsets[n] = acc;
    if (out_ids) { job.phase = 1; job.next = 0; pool_run(&job, nthreads, &w); }
    work_free(&w);
    const int err = job.err;
    pthread_mutex_unlock(&g_pool.call_mu);
    return err ? -1 : 0;
}
