// tables.h -- layout of the packed vocabulary tables (host builder <-> device kernels).
//
// One vocabulary is ONE contiguous blob (what cfbpe_vocab_export hands out and what the
// host layer broadcasts to the other GPUs): a TablesHeader followed by 256-byte aligned
// sections.  All lookups the kernels make are restated from tiktoken semantics:
//   * rank == token id, so "rank of the pair" == id of the merged token;
//   * ranks[left_bytes + right_bytes] is keyed by BYTES, therefore the pair table holds
//     every (left,right) split of every token whose halves are both tokens,
//     not only the training-time split.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CFBPE_HD __host__ __device__ __forceinline__
#else
#define CFBPE_HD inline
#endif

namespace cfbpe {

constexpr uint32_t kTablesMagic = 0x45504243u;  // "CBPE"
constexpr uint32_t kTablesVersion = 1;
constexpr uint32_t kNone = 0xFFFFFFFFu;          // "no such token / no such pair"
constexpr uint32_t kIdBits = 21;                 // vocab size < 2^21 - 1
constexpr uint32_t kIdMask = (1u << kIdBits) - 1;
constexpr uint64_t kPairEmpty = ~0ull;
constexpr uint32_t kShortMaxLen = 12;            // tokens <= 12 bytes: key stored inline
constexpr uint32_t kMetaEmpty = 0xFFFFFFFFu;

struct TablesHeader {
    uint32_t magic, version;
    uint32_t n_ranks, pattern_id, max_token_len, n_pair_entries;
    uint64_t total_bytes;
    // section offsets from blob start (bytes) and capacities (slots, powers of two)
    uint64_t off_byte2id;   // u32[256]     id of each single byte
    uint64_t off_bytepair;  // u32[65536]   merged id of raw byte pair (l<<8|r), or kNone
    uint64_t off_pair;      // u64[cap_pair] (left<<42 | right<<21 | merged), kPairEmpty = free
    uint64_t off_short;     // ShortSlot[cap_short]  tokens of <= 12 bytes, exact inline key
    uint64_t off_long;      // LongSlot[cap_long]    tokens of 13.. bytes, hash + verify
    uint64_t off_tokoff;    // u32[n_ranks+1]        start of each token in the byte blob
    uint64_t off_blob;      // u8[...]               token bytes, rank order
    uint32_t cap_pair, cap_short, cap_long, blob_bytes;
    uint64_t content_hash;  // FNV-1a of the rank list, for cross-rank consistency checks
};

struct ShortSlot { uint64_t k0; uint32_t k1; uint32_t meta; };   // meta = len<<24 | id ; kMetaEmpty = free
struct LongSlot { uint64_t hash; uint32_t meta; uint32_t blob_off; };

// device/host view: raw pointers into one blob
struct TablesView {
    const uint32_t* byte2id;
    const uint32_t* bytepair;
    const uint64_t* pair;
    const ShortSlot* shrt;
    const LongSlot* lng;
    const uint32_t* tokoff;
    const uint8_t* blob;
    uint32_t pair_mask, short_mask, long_mask;
    uint32_t n_ranks, pattern_id, max_token_len;
};

static inline TablesView make_view(const uint8_t* base, const TablesHeader& h) {
    TablesView v;
    v.byte2id = reinterpret_cast<const uint32_t*>(base + h.off_byte2id);
    v.bytepair = reinterpret_cast<const uint32_t*>(base + h.off_bytepair);
    v.pair = reinterpret_cast<const uint64_t*>(base + h.off_pair);
    v.shrt = reinterpret_cast<const ShortSlot*>(base + h.off_short);
    v.lng = reinterpret_cast<const LongSlot*>(base + h.off_long);
    v.tokoff = reinterpret_cast<const uint32_t*>(base + h.off_tokoff);
    v.blob = base + h.off_blob;
    v.pair_mask = h.cap_pair - 1;
    v.short_mask = h.cap_short - 1;
    v.long_mask = h.cap_long - 1;
    v.n_ranks = h.n_ranks;
    v.pattern_id = h.pattern_id;
    v.max_token_len = h.max_token_len;
    return v;
}

// ---- hashing (identical on host and device) -------------------------------------------
CFBPE_HD uint64_t mix64(uint64_t x) {  // splitmix64 finaliser
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}
CFBPE_HD uint32_t mix32(uint32_t h) {  // 32-bit finaliser (lowbias32)
    h ^= h >> 16; h *= 0x21f0aaadu;
    h ^= h >> 15; h *= 0x735a2d97u;
    h ^= h >> 15;
    return h;
}
CFBPE_HD uint32_t pair_hash(uint32_t left, uint32_t right) {
    return mix32(left * 0x9E3779B1u + right * 0x85EBCA77u + 0x165667B1u);
}
CFBPE_HD uint64_t pair_slot(uint32_t left, uint32_t right, uint32_t merged) {
    return (static_cast<uint64_t>(left) << (2 * kIdBits)) | (static_cast<uint64_t>(right) << kIdBits) | merged;
}
CFBPE_HD uint32_t short_hash(uint64_t k0, uint32_t k1, uint32_t len) {
    const uint32_t lo = static_cast<uint32_t>(k0), hi = static_cast<uint32_t>(k0 >> 32);
    return mix32(lo * 0x9E3779B1u + hi * 0x85EBCA77u + k1 * 0xC2B2AE3Du + len * 0x27D4EB2Fu);
}
// long tokens: first 12 bytes + last 4 bytes + length (cheap to gather); collisions are
// resolved by comparing the bytes against the blob.
CFBPE_HD uint64_t long_hash(uint64_t k0, uint32_t k1, uint32_t last4, uint32_t len) {
    return mix64(k0 ^ mix64((static_cast<uint64_t>(k1) << 32) | last4) ^ (static_cast<uint64_t>(len) << 48));
}

// ---- lookups ----------------------------------------------------------------------------
// merged id of (left,right) or kNone.  Linear probing, load <= 0.25.
CFBPE_HD uint32_t pair_lookup(const TablesView& t, uint32_t left, uint32_t right) {
    const uint64_t key = (static_cast<uint64_t>(left) << kIdBits) | right;
    uint32_t h = pair_hash(left, right) & t.pair_mask;
    for (;;) {
        const uint64_t s = t.pair[h];
        if ((s >> kIdBits) == key) return static_cast<uint32_t>(s) & kIdMask;
        if (s == kPairEmpty) return kNone;
        h = (h + 1) & t.pair_mask;
    }
}
// two independent probes with their first loads in flight together (a merge refreshes both neighbouring pairs)
CFBPE_HD void pair_lookup2(const TablesView& t, uint32_t l0, uint32_t r0, bool want0, uint32_t l1, uint32_t r1, bool want1,
                           uint32_t& out0, uint32_t& out1) {
    const uint64_t key0 = (static_cast<uint64_t>(l0) << kIdBits) | r0, key1 = (static_cast<uint64_t>(l1) << kIdBits) | r1;
    uint32_t h0 = pair_hash(l0, r0) & t.pair_mask, h1 = pair_hash(l1, r1) & t.pair_mask;
    uint64_t s0 = want0 ? t.pair[h0] : kPairEmpty;
    uint64_t s1 = want1 ? t.pair[h1] : kPairEmpty;
    out0 = kNone; out1 = kNone;
    for (;;) {
        if ((s0 >> kIdBits) == key0) { out0 = static_cast<uint32_t>(s0) & kIdMask; break; }
        if (s0 == kPairEmpty) break;
        h0 = (h0 + 1) & t.pair_mask; s0 = t.pair[h0];
    }
    for (;;) {
        if ((s1 >> kIdBits) == key1) { out1 = static_cast<uint32_t>(s1) & kIdMask; break; }
        if (s1 == kPairEmpty) break;
        h1 = (h1 + 1) & t.pair_mask; s1 = t.pair[h1];
    }
}
// id of a token of len <= 12 whose bytes are packed little-endian in (k0,k1), or kNone
CFBPE_HD uint32_t short_lookup(const TablesView& t, uint64_t k0, uint32_t k1, uint32_t len) {
    uint32_t h = short_hash(k0, k1, len) & t.short_mask;
    const uint32_t want = len << 24;
    for (;;) {
        const ShortSlot s = t.shrt[h];
        if (s.meta == kMetaEmpty) return kNone;
        if (s.k0 == k0 && s.k1 == k1 && (s.meta & 0xFF000000u) == want) return s.meta & 0x00FFFFFFu;
        h = (h + 1) & t.short_mask;
    }
}
// four bytes at p (any alignment) as a little-endian word: two aligned loads and a funnel shift on the device
// (both the text and the token blob are readable 16 bytes past their end)
CFBPE_HD uint32_t load_u32_any(const uint8_t* p) {
#if defined(__CUDA_ARCH__)
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* q = reinterpret_cast<const uint32_t*>(a & ~static_cast<uintptr_t>(3));
    const uint32_t sh = static_cast<uint32_t>(a & 3) * 8;
    return sh ? __funnelshift_r(q[0], q[1], sh) : q[0];
#else
    return static_cast<uint32_t>(p[0]) | (static_cast<uint32_t>(p[1]) << 8) | (static_cast<uint32_t>(p[2]) << 16) |
           (static_cast<uint32_t>(p[3]) << 24);
#endif
}
// len >= 4 bytes equal?  four at a time, the last word overlapping
CFBPE_HD bool bytes_equal(const uint8_t* a, const uint8_t* b, uint32_t len) {
    uint32_t i = 0;
    for (; i + 4 <= len; i += 4) if (load_u32_any(a + i) != load_u32_any(b + i)) return false;
    return i == len || load_u32_any(a + len - 4) == load_u32_any(b + len - 4);
}
// id of a token of len >= 13 given its hash and a pointer to its bytes, or kNone
CFBPE_HD uint32_t long_lookup(const TablesView& t, uint64_t hash, const uint8_t* bytes, uint32_t len) {
    uint32_t h = static_cast<uint32_t>(hash >> 17) & t.long_mask;
    for (;;) {
        const LongSlot s = t.lng[h];
        if (s.meta == kMetaEmpty) return kNone;
        if (s.hash == hash && (s.meta >> 24) == (len & 0xFF)) {
            const uint32_t id = s.meta & 0x00FFFFFFu;
            if (t.tokoff[id + 1] - t.tokoff[id] == len) {
                if (bytes_equal(t.blob + s.blob_off, bytes, len)) return id;
            }
        }
        h = (h + 1) & t.long_mask;
    }
}

// pack up to 12 bytes little-endian into (k0,k1)
CFBPE_HD void pack_key(const uint8_t* p, uint32_t len, uint64_t& k0, uint32_t& k1) {
    k0 = 0; k1 = 0;
    for (uint32_t i = 0; i < len && i < 8; ++i) k0 |= static_cast<uint64_t>(p[i]) << (8 * i);
    for (uint32_t i = 8; i < len && i < 12; ++i) k1 |= static_cast<uint32_t>(p[i]) << (8 * (i - 8));
}
CFBPE_HD uint32_t load_le32(const uint8_t* p) {
    return static_cast<uint32_t>(p[0]) | (static_cast<uint32_t>(p[1]) << 8) | (static_cast<uint32_t>(p[2]) << 16) |
           (static_cast<uint32_t>(p[3]) << 24);
}
// whole-piece lookup from raw bytes (any length); used by the long-piece kernel and host checks
CFBPE_HD uint32_t piece_lookup(const TablesView& t, const uint8_t* p, uint32_t len) {
    if (len == 0 || len > t.max_token_len) return kNone;
    uint64_t k0; uint32_t k1;
    pack_key(p, len, k0, k1);
    if (len <= kShortMaxLen) return short_lookup(t, k0, k1, len);
    return long_lookup(t, long_hash(k0, k1, load_le32(p + len - 4), len), p, len);
}

}  // namespace cfbpe
