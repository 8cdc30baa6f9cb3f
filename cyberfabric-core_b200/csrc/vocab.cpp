// vocab.cpp -- rank-file parsers and packed-table builder (see vocab.h).
#include "vocab.h"

#include <cstring>
#include <string_view>
#include <unordered_map>

#include "../../include/cfbpe.h"

namespace cfbpe {

namespace {

int b64val(int c) {
    if (c >= 'A' && c <= 'Z') return c - 'A';
    if (c >= 'a' && c <= 'z') return c - 'a' + 26;
    if (c >= '0' && c <= '9') return c - '0' + 52;
    if (c == '+') return 62;
    if (c == '/') return 63;
    return -1;
}

bool b64decode(const uint8_t* p, size_t n, std::string& out) {
    out.clear();
    uint32_t acc = 0;
    int bits = 0;
    for (size_t i = 0; i < n; ++i) {
        if (p[i] == '=') break;
        int d = b64val(p[i]);
        if (d < 0) return false;
        acc = (acc << 6) | static_cast<uint32_t>(d);
        bits += 6;
        if (bits >= 8) {
            bits -= 8;
            out.push_back(static_cast<char>((acc >> bits) & 0xFF));
            acc &= (1u << bits) - 1;
        }
    }
    return true;
}

int finish_tokens(std::vector<std::string>& tokens, std::vector<uint8_t>& seen, uint32_t max_ranks, std::string& err) {
    // ranks must be dense 0..n-1
    size_t n = tokens.size();
    if (max_ranks && n > max_ranks) { tokens.resize(max_ranks); n = max_ranks; }
    for (size_t i = 0; i < n; ++i)
        if (!seen[i] || tokens[i].empty()) { err = "rank file: rank " + std::to_string(i) + " missing"; return CFBPE_EINVAL; }
    if (n < 256) { err = "rank file: fewer than 256 ranks"; return CFBPE_EINVAL; }
    return CFBPE_OK;
}

void put_token(std::vector<std::string>& tokens, std::vector<uint8_t>& seen, uint64_t rank, std::string&& bytes) {
    if (rank >= tokens.size()) { tokens.resize(rank + 1); seen.resize(rank + 1, 0); }
    tokens[rank] = std::move(bytes);
    seen[rank] = 1;
}

constexpr uint64_t kMaxRank = (1u << kIdBits) - 2;

}  // namespace

int parse_tiktoken(const uint8_t* file, size_t len, uint32_t max_ranks, std::vector<std::string>& tokens,
                   std::string& err) {
    tokens.clear();
    std::vector<uint8_t> seen;
    size_t pos = 0, line = 0;
    std::string tb;
    while (pos < len) {
        size_t eol = pos;
        while (eol < len && file[eol] != '\n') ++eol;
        size_t end = eol;
        if (end > pos && file[end - 1] == '\r') --end;
        ++line;
        if (end > pos) {
            size_t sp = pos;
            while (sp < end && file[sp] != ' ') ++sp;
            if (sp == end || sp == pos) { err = "rank file line " + std::to_string(line) + ": expected '<base64> <rank>'"; return CFBPE_EINVAL; }
            if (!b64decode(file + pos, sp - pos, tb) || tb.empty()) { err = "rank file line " + std::to_string(line) + ": bad base64"; return CFBPE_EINVAL; }
            uint64_t rank = 0;
            size_t i = sp + 1;
            if (i == end) { err = "rank file line " + std::to_string(line) + ": missing rank"; return CFBPE_EINVAL; }
            for (; i < end; ++i) {
                if (file[i] < '0' || file[i] > '9') { err = "rank file line " + std::to_string(line) + ": bad rank"; return CFBPE_EINVAL; }
                rank = rank * 10 + (file[i] - '0');
                if (rank > kMaxRank) { err = "rank file: rank too large (limit 2^21-2)"; return CFBPE_EINVAL; }
            }
            if (!max_ranks || rank < max_ranks) {
                if (rank < seen.size() && seen[rank]) { err = "rank file: duplicate rank " + std::to_string(rank); return CFBPE_EINVAL; }
                put_token(tokens, seen, rank, std::move(tb));
            }
        }
        pos = eol + 1;
    }
    return finish_tokens(tokens, seen, max_ranks, err);
}

// Minimal scanner for tekken_*.json: walks the top-level "vocab" array and reads, per object,
// the integer "rank" and the base64 string "token_bytes".  Strings are skipped with escape handling.
int parse_tekken_json(const uint8_t* f, size_t len, uint32_t max_ranks, std::vector<std::string>& tokens,
                      std::string& err) {
    tokens.clear();
    std::vector<uint8_t> seen;
    size_t i = 0;
    auto skip_ws = [&]() { while (i < len && (f[i] == ' ' || f[i] == '\n' || f[i] == '\r' || f[i] == '\t')) ++i; };
    auto read_string = [&](size_t& s, size_t& e) -> bool {  // f[i]=='"' ; returns raw span without quotes
        if (i >= len || f[i] != '"') return false;
        s = ++i;
        while (i < len && f[i] != '"') { if (f[i] == '\\') ++i; ++i; }
        if (i >= len) return false;
        e = i++;
        return true;
    };
    // find the "vocab" key whose value is an array (top-level in the published files)
    const char* key = "\"vocab\"";
    std::string_view all(reinterpret_cast<const char*>(f), len);
    size_t kpos = 0;
    bool found = false;
    while ((kpos = all.find(key, kpos)) != std::string_view::npos) {
        i = kpos + std::strlen(key);
        skip_ws();
        if (i < len && f[i] == ':') {
            ++i; skip_ws();
            if (i < len && f[i] == '[') { ++i; found = true; break; }
        }
        kpos += 1;
    }
    if (!found) { err = "tekken json: no \"vocab\" array"; return CFBPE_EINVAL; }
    std::string tb;
    for (;;) {
        skip_ws();
        if (i < len && f[i] == ']') break;
        if (i < len && f[i] == ',') { ++i; continue; }
        if (i >= len || f[i] != '{') { err = "tekken json: expected object in vocab"; return CFBPE_EINVAL; }
        ++i;
        bool have_rank = false, have_bytes = false;
        uint64_t rank = 0;
        for (;;) {
            skip_ws();
            if (i < len && f[i] == '}') { ++i; break; }
            if (i < len && f[i] == ',') { ++i; continue; }
            size_t ks, ke;
            if (!read_string(ks, ke)) { err = "tekken json: expected key"; return CFBPE_EINVAL; }
            std::string_view k(reinterpret_cast<const char*>(f) + ks, ke - ks);
            skip_ws();
            if (i >= len || f[i] != ':') { err = "tekken json: expected ':'"; return CFBPE_EINVAL; }
            ++i; skip_ws();
            if (i < len && f[i] == '"') {
                size_t vs, ve;
                if (!read_string(vs, ve)) { err = "tekken json: bad string"; return CFBPE_EINVAL; }
                if (k == "token_bytes") {
                    if (!b64decode(f + vs, ve - vs, tb) || tb.empty()) { err = "tekken json: bad base64"; return CFBPE_EINVAL; }
                    have_bytes = true;
                }
            } else {  // number / null / true / false
                size_t vs = i;
                while (i < len && f[i] != ',' && f[i] != '}' && f[i] != ' ' && f[i] != '\n') ++i;
                if (k == "rank") {
                    rank = 0;
                    for (size_t j = vs; j < i; ++j) {
                        if (f[j] < '0' || f[j] > '9') { err = "tekken json: bad rank"; return CFBPE_EINVAL; }
                        rank = rank * 10 + (f[j] - '0');
                        if (rank > kMaxRank) { err = "tekken json: rank too large"; return CFBPE_EINVAL; }
                    }
                    have_rank = true;
                }
            }
        }
        if (!have_rank || !have_bytes) { err = "tekken json: vocab entry without rank/token_bytes"; return CFBPE_EINVAL; }
        if (!max_ranks || rank < max_ranks) {
            if (rank < seen.size() && seen[rank]) { err = "tekken json: duplicate rank"; return CFBPE_EINVAL; }
            put_token(tokens, seen, rank, std::move(tb));
        }
    }
    return finish_tokens(tokens, seen, max_ranks, err);
}

namespace {
uint32_t pow2_at_least(uint64_t x) {
    uint32_t c = 16;
    while (c < x) c <<= 1;
    return c;
}
uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }
}  // namespace

int build_tables(const std::vector<std::string>& tokens, uint32_t pattern_id, std::vector<uint8_t>& blob,
                 std::string& err) {
    const uint32_t n = static_cast<uint32_t>(tokens.size());
    if (n < 256 || n > kMaxRank) { err = "vocab size out of range"; return CFBPE_EINVAL; }
    if (pattern_id >= CFBPE_PATTERN_COUNT) { err = "unknown pattern id"; return CFBPE_EINVAL; }

    std::unordered_map<std::string_view, uint32_t> index;
    index.reserve(n * 2);
    uint32_t max_len = 0;
    uint64_t blob_bytes = 0;
    for (uint32_t r = 0; r < n; ++r) {
        const std::string& t = tokens[r];
        if (t.empty()) { err = "empty token at rank " + std::to_string(r); return CFBPE_EINVAL; }
        if (t.size() > 255) { err = "token longer than 255 bytes at rank " + std::to_string(r); return CFBPE_EINVAL; }
        if (!index.emplace(std::string_view(t), r).second) { err = "duplicate token bytes at rank " + std::to_string(r); return CFBPE_EINVAL; }
        if (t.size() > max_len) max_len = static_cast<uint32_t>(t.size());
        blob_bytes += t.size();
    }
    // every single byte must be a token (tiktoken panics otherwise)
    uint32_t byte2id[256];
    for (int b = 0; b < 256; ++b) {
        char c = static_cast<char>(b);
        auto it = index.find(std::string_view(&c, 1));
        if (it == index.end()) { err = "byte " + std::to_string(b) + " is not in the vocabulary"; return CFBPE_EINVAL; }
        byte2id[b] = it->second;
    }

    // all-splits pair list
    struct PairE { uint32_t l, r, m; };
    std::vector<PairE> pairs;
    pairs.reserve(static_cast<size_t>(n) * 5 / 2);
    uint32_t n_short = 0, n_long = 0;
    for (uint32_t id = 0; id < n; ++id) {
        const std::string& t = tokens[id];
        if (t.size() <= kShortMaxLen) ++n_short; else ++n_long;
        std::string_view sv(t);
        for (size_t k = 1; k < t.size(); ++k) {
            auto il = index.find(sv.substr(0, k));
            if (il == index.end()) continue;
            auto ir = index.find(sv.substr(k));
            if (ir == index.end()) continue;
            pairs.push_back({il->second, ir->second, id});
        }
    }

    TablesHeader h{};
    h.magic = kTablesMagic;
    h.version = kTablesVersion;
    h.n_ranks = n;
    h.pattern_id = pattern_id;
    h.max_token_len = max_len;
    h.n_pair_entries = static_cast<uint32_t>(pairs.size());
    h.cap_pair = pow2_at_least(pairs.size() * 4 + 16);   // load <= 0.25 (0.5 measured 5-20 % slower in every kernel that probes; a two-choice cuckoo table and 4-slot buckets measured no better than this)
    h.cap_short = pow2_at_least(static_cast<uint64_t>(n_short) * 2 + 16);
    h.cap_long = pow2_at_least(static_cast<uint64_t>(n_long) * 2 + 16);
    h.blob_bytes = static_cast<uint32_t>(blob_bytes);
    uint64_t off = align_up(sizeof(TablesHeader), 256);
    h.off_byte2id = off;  off = align_up(off + 256 * 4, 256);
    h.off_bytepair = off; off = align_up(off + 65536 * 4, 256);
    h.off_pair = off;     off = align_up(off + static_cast<uint64_t>(h.cap_pair) * 8, 256);
    h.off_short = off;    off = align_up(off + static_cast<uint64_t>(h.cap_short) * sizeof(ShortSlot), 256);
    h.off_long = off;     off = align_up(off + static_cast<uint64_t>(h.cap_long) * sizeof(LongSlot), 256);
    h.off_tokoff = off;   off = align_up(off + (static_cast<uint64_t>(n) + 1) * 4, 256);
    h.off_blob = off;     off = align_up(off + blob_bytes + 16, 256);   // +16: kernels may over-read a few bytes
    h.total_bytes = off;

    blob.assign(off, 0);
    uint8_t* base = blob.data();
    std::memcpy(base + h.off_byte2id, byte2id, sizeof byte2id);

    uint32_t* bytepair = reinterpret_cast<uint32_t*>(base + h.off_bytepair);
    for (uint32_t i = 0; i < 65536; ++i) bytepair[i] = kNone;
    uint64_t* pt = reinterpret_cast<uint64_t*>(base + h.off_pair);
    for (uint32_t i = 0; i < h.cap_pair; ++i) pt[i] = kPairEmpty;
    for (const PairE& p : pairs) {
        uint32_t s = pair_hash(p.l, p.r) & (h.cap_pair - 1);
        while (pt[s] != kPairEmpty) s = (s + 1) & (h.cap_pair - 1);
        pt[s] = pair_slot(p.l, p.r, p.m);
    }
    // raw byte pair table: token whose bytes are exactly (a,b)
    for (uint32_t id = 0; id < n; ++id) {
        const std::string& t = tokens[id];
        if (t.size() == 2) bytepair[(static_cast<uint8_t>(t[0]) << 8) | static_cast<uint8_t>(t[1])] = id;
    }

    ShortSlot* st = reinterpret_cast<ShortSlot*>(base + h.off_short);
    for (uint32_t i = 0; i < h.cap_short; ++i) st[i] = ShortSlot{0, 0, kMetaEmpty};
    LongSlot* lt = reinterpret_cast<LongSlot*>(base + h.off_long);
    for (uint32_t i = 0; i < h.cap_long; ++i) lt[i] = LongSlot{0, kMetaEmpty, 0};
    uint32_t* tokoff = reinterpret_cast<uint32_t*>(base + h.off_tokoff);
    uint8_t* bytes = base + h.off_blob;
    uint32_t bo = 0;
    uint64_t ch = 1469598103934665603ull;
    for (uint32_t id = 0; id < n; ++id) {
        const std::string& t = tokens[id];
        const uint32_t len = static_cast<uint32_t>(t.size());
        tokoff[id] = bo;
        std::memcpy(bytes + bo, t.data(), len);
        const uint8_t* p = bytes + bo;
        ch = (ch ^ len) * 1099511628211ull;
        for (uint32_t i = 0; i < len; ++i) ch = (ch ^ p[i]) * 1099511628211ull;
        uint64_t k0; uint32_t k1;
        pack_key(p, len, k0, k1);
        if (len <= kShortMaxLen) {
            uint32_t s = short_hash(k0, k1, len) & (h.cap_short - 1);
            while (st[s].meta != kMetaEmpty) s = (s + 1) & (h.cap_short - 1);
            st[s] = ShortSlot{k0, k1, (len << 24) | id};
        } else {
            const uint64_t hv = long_hash(k0, k1, load_le32(p + len - 4), len);
            uint32_t s = static_cast<uint32_t>(hv >> 17) & (h.cap_long - 1);
            while (lt[s].meta != kMetaEmpty) s = (s + 1) & (h.cap_long - 1);
            lt[s] = LongSlot{hv, (len << 24) | id, bo};
        }
        bo += len;
    }
    tokoff[n] = bo;
    h.content_hash = ch;
    std::memcpy(base, &h, sizeof h);
    return CFBPE_OK;
}

int validate_tables(const uint8_t* blob, uint64_t size, std::string& err) {
    if (size < sizeof(TablesHeader)) { err = "table blob too small"; return CFBPE_EINVAL; }
    TablesHeader h;
    std::memcpy(&h, blob, sizeof h);
    auto pow2 = [](uint32_t x) { return x && !(x & (x - 1)); };
    if (h.magic != kTablesMagic || h.version != kTablesVersion) { err = "table blob: bad magic/version"; return CFBPE_EINVAL; }
    if (h.total_bytes != size) { err = "table blob: size mismatch"; return CFBPE_EINVAL; }
    if (!pow2(h.cap_pair) || !pow2(h.cap_short) || !pow2(h.cap_long)) { err = "table blob: capacities"; return CFBPE_EINVAL; }
    if (h.n_ranks < 256 || h.n_ranks > kMaxRank || h.pattern_id >= CFBPE_PATTERN_COUNT || h.max_token_len > 255) { err = "table blob: header fields"; return CFBPE_EINVAL; }
    auto inside = [&](uint64_t off, uint64_t bytes) { return off % 16 == 0 && off <= size && bytes <= size - off; };
    if (!inside(h.off_byte2id, 1024) || !inside(h.off_bytepair, 65536 * 4) ||
        !inside(h.off_pair, static_cast<uint64_t>(h.cap_pair) * 8) ||
        !inside(h.off_short, static_cast<uint64_t>(h.cap_short) * sizeof(ShortSlot)) ||
        !inside(h.off_long, static_cast<uint64_t>(h.cap_long) * sizeof(LongSlot)) ||
        !inside(h.off_tokoff, (static_cast<uint64_t>(h.n_ranks) + 1) * 4) ||
        !inside(h.off_blob, static_cast<uint64_t>(h.blob_bytes) + 16)) { err = "table blob: section out of bounds"; return CFBPE_EINVAL; }
    // The CONTENT, not only the header: an imported blob comes from another process.  The device probes walk a table until
    // they meet a free slot, read token bytes at stored offsets and index arrays with stored ids -- a table without free slots
    // would spin a kernel forever, a wild offset or id would read outside the blob.
    const uint64_t* pairs = reinterpret_cast<const uint64_t*>(blob + h.off_pair);
    uint64_t free_pairs = 0;
    for (uint32_t i = 0; i < h.cap_pair; ++i) {
        const uint64_t sl = pairs[i];
        if (sl == kPairEmpty) { ++free_pairs; continue; }
        const uint32_t l = static_cast<uint32_t>(sl >> (2 * kIdBits)), r = static_cast<uint32_t>(sl >> kIdBits) & kIdMask, m = static_cast<uint32_t>(sl) & kIdMask;
        if (l >= h.n_ranks || r >= h.n_ranks || m >= h.n_ranks) { err = "table blob: pair entry names an id outside the vocabulary"; return CFBPE_EINVAL; }
    }
    if (free_pairs * 2 < h.cap_pair) { err = "table blob: pair table more than half full"; return CFBPE_EINVAL; }
    const ShortSlot* st = reinterpret_cast<const ShortSlot*>(blob + h.off_short);
    uint64_t free_short = 0;
    for (uint32_t i = 0; i < h.cap_short; ++i) {
        if (st[i].meta == kMetaEmpty) { ++free_short; continue; }
        if ((st[i].meta & 0x00FFFFFFu) >= h.n_ranks || (st[i].meta >> 24) > kShortMaxLen) { err = "table blob: short-token entry out of range"; return CFBPE_EINVAL; }
    }
    const LongSlot* lt = reinterpret_cast<const LongSlot*>(blob + h.off_long);
    uint64_t free_long = 0;
    for (uint32_t i = 0; i < h.cap_long; ++i) {
        if (lt[i].meta == kMetaEmpty) { ++free_long; continue; }
        const uint32_t len = lt[i].meta >> 24;
        if ((lt[i].meta & 0x00FFFFFFu) >= h.n_ranks || static_cast<uint64_t>(lt[i].blob_off) + len > h.blob_bytes) { err = "table blob: long-token entry out of range"; return CFBPE_EINVAL; }
    }
    if (!free_short || !free_long) { err = "table blob: a token table has no free slot"; return CFBPE_EINVAL; }
    const uint32_t* tokoff = reinterpret_cast<const uint32_t*>(blob + h.off_tokoff);
    const uint32_t* byte2id = reinterpret_cast<const uint32_t*>(blob + h.off_byte2id);
    const uint32_t* bytepair = reinterpret_cast<const uint32_t*>(blob + h.off_bytepair);
    for (uint32_t i = 0; i < h.n_ranks; ++i)
        if (tokoff[i] > tokoff[i + 1] || tokoff[i + 1] - tokoff[i] > 255) { err = "table blob: token offsets"; return CFBPE_EINVAL; }
    if (tokoff[0] != 0 || tokoff[h.n_ranks] != h.blob_bytes) { err = "table blob: token offsets do not cover the byte blob"; return CFBPE_EINVAL; }
    for (uint32_t i = 0; i < 256; ++i) if (byte2id[i] != kNone && byte2id[i] >= h.n_ranks) { err = "table blob: byte table"; return CFBPE_EINVAL; }
    for (uint32_t i = 0; i < 65536; ++i) if (bytepair[i] != kNone && bytepair[i] >= h.n_ranks) { err = "table blob: byte-pair table"; return CFBPE_EINVAL; }
    // the hash the builder left over (length, bytes) of every token in rank order: a blob that was altered after it was built fails here
    const uint8_t* bytes = blob + h.off_blob;
    uint64_t ch = 1469598103934665603ull;
    for (uint32_t id = 0; id < h.n_ranks; ++id) {
        const uint32_t len = tokoff[id + 1] - tokoff[id];
        ch = (ch ^ len) * 1099511628211ull;
        for (uint32_t i = 0; i < len; ++i) ch = (ch ^ bytes[tokoff[id] + i]) * 1099511628211ull;
    }
    if (ch != h.content_hash) { err = "table blob: content hash mismatch"; return CFBPE_EINVAL; }
    return CFBPE_OK;
}

}  // namespace cfbpe
