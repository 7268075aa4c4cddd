// hannoy_oracle.cpp — CPU restatement of hannoy's search path.  TEST INFRASTRUCTURE ONLY.
//
// This file is the *oracle* for hannoy_b200: a literal, single-file C++17 restatement of the
// reference algorithm (nnethercott/hannoy v0.1.3, Rust).  Only `tests/`, `__graft_entry__.smoke()`
// and `bench.py`'s cpu_baseline / `--impl reference` legs may load it; the product
// (`hannoy_b200/`, `libhannoy_b200.so`) never links, imports or calls anything in `oracle/`.
//
// PARITY PINNING: the reference cannot be compiled here (no Rust toolchain, no lockfile, no LMDB).
// The pieces of this oracle that the reference's own tests pin with known answers are checked
// against those vectors in tests/test_oracle_kat.py (quantizer bit patterns, SIMD==scalar integer
// vectors, OrderedFloat order, Hamming one-hot, empty index, self-query, reachability, candidates,
// by_item exclusion, golden graph topologies).  Search results on non-trivial graphs are NOT pinned by any
// reference fixture ("parity unpinned" for those) — this restatement is the arbiter there, which
// is why every function cites the reference lines it transcribes.
//
// All `file:line` citations are relative to /root/reference/.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <queue>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#if defined(__GNUC__)
#pragma GCC optimize("no-fast-math")
#endif
// Built with -ffp-contract=off: every fused multiply-add below is an explicit std::fmaf.

namespace orc {

// ------------------------------------------------------------------------------------------
// Metrics (src/distance/*.rs `fn name()`), same order as hb_metric in include/hannoy_b200.h
// ------------------------------------------------------------------------------------------
enum Metric : int {
    EUCLIDEAN = 0,     // src/distance/euclidean.rs:33  "euclidean"
    COSINE = 1,        // src/distance/cosine.rs:32     "cosine"
    MANHATTAN = 2,     // src/distance/manhattan.rs:33  "manhattan"
    HAMMING = 3,       // src/distance/hamming.rs:32    "hamming"            codec Binary
    BQ_COSINE = 4,     // binary_quantized_cosine.rs:36 "binary quantized cosine"
    BQ_EUCLIDEAN = 5,  // binary_quantized_euclidean.rs:35
    BQ_MANHATTAN = 6,  // binary_quantized_manhattan.rs:35
};
static const char* metric_name(int m) {
    static const char* names[] = {"euclidean", "cosine", "manhattan", "hamming",
                                  "binary quantized cosine", "binary quantized euclidean",
                                  "binary quantized manhattan"};
    return (m >= 0 && m < 7) ? names[m] : "?";
}
static inline bool is_binary(int m) { return m >= HAMMING; }
// size_of::<D::Header>() — NodeHeaderHamming { idx: usize } is 8 bytes (hamming.rs:21-25), all
// others are one f32.
static inline size_t header_bytes(int m) { return m == HAMMING ? 8 : 4; }

static inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

// ------------------------------------------------------------------------------------------
// src/spaces/simple_avx.rs + simple_sse.rs + simple.rs — exact summation order, in scalar code
// ------------------------------------------------------------------------------------------
// hsum256_ps_avx (simple_avx.rs:8-13): r[l]=x[l+4]+x[l]; (r0+r2)+(r1+r3)
static inline float hsum8(const float* x) {
    float r0 = x[4] + x[0], r1 = x[5] + x[1], r2 = x[6] + x[2], r3 = x[7] + x[3];
    float s0 = r0 + r2, s1 = r1 + r3;
    return s0 + s1;
}
// hsum128_ps_sse (simple_sse.rs:12-16): (x0+x2)+(x1+x3)
static inline float hsum4(const float* x) {
    float s0 = x[0] + x[2], s1 = x[1] + x[3];
    return s0 + s1;
}

// dot_similarity_avx (simple_avx.rs:67-110)
static float dot_avx(const float* a, const float* b, size_t n) {
    size_t m = n - (n % 32);
    float acc[32];
    for (int j = 0; j < 32; ++j) acc[j] = 0.0f;
    for (size_t i = 0; i < m; i += 32)
        for (int j = 0; j < 32; ++j) acc[j] = std::fmaf(a[i + j], b[i + j], acc[j]);
    float result = hsum8(acc) + hsum8(acc + 8);
    result = result + hsum8(acc + 16);
    result = result + hsum8(acc + 24);
    for (size_t i = m; i < n; ++i) {
        float p = a[i] * b[i];
        result = result + p;
    }
    return result;
}
// euclid_similarity_avx (simple_avx.rs:17-65)
static float euclid_avx(const float* a, const float* b, size_t n) {
    size_t m = n - (n % 32);
    float acc[32];
    for (int j = 0; j < 32; ++j) acc[j] = 0.0f;
    for (size_t i = 0; i < m; i += 32)
        for (int j = 0; j < 32; ++j) {
            float d = a[i + j] - b[i + j];
            acc[j] = std::fmaf(d, d, acc[j]);
        }
    float result = hsum8(acc) + hsum8(acc + 8);
    result = result + hsum8(acc + 16);
    result = result + hsum8(acc + 24);
    for (size_t i = m; i < n; ++i) {
        float d = a[i] - b[i];
        float p = d * d;  // (a - b).powi(2)
        result = result + p;
    }
    return result;
}
// dot_similarity_sse (simple_sse.rs:66-110): mul then add, NOT fused
static float dot_sse(const float* a, const float* b, size_t n) {
    size_t m = n - (n % 16);
    float acc[16];
    for (int j = 0; j < 16; ++j) acc[j] = 0.0f;
    for (size_t i = 0; i < m; i += 16)
        for (int j = 0; j < 16; ++j) {
            float p = a[i + j] * b[i + j];
            acc[j] = p + acc[j];
        }
    float result = hsum4(acc) + hsum4(acc + 4);
    result = result + hsum4(acc + 8);
    result = result + hsum4(acc + 12);
    for (size_t i = m; i < n; ++i) {
        float p = a[i] * b[i];
        result = result + p;
    }
    return result;
}
// euclid_similarity_sse (simple_sse.rs:18-64)
static float euclid_sse(const float* a, const float* b, size_t n) {
    size_t m = n - (n % 16);
    float acc[16];
    for (int j = 0; j < 16; ++j) acc[j] = 0.0f;
    for (size_t i = 0; i < m; i += 16)
        for (int j = 0; j < 16; ++j) {
            float d = a[i + j] - b[i + j];
            float p = d * d;
            acc[j] = p + acc[j];
        }
    float result = hsum4(acc) + hsum4(acc + 4);
    result = result + hsum4(acc + 8);
    result = result + hsum4(acc + 12);
    for (size_t i = m; i < n; ++i) {
        float d = a[i] - b[i];
        float p = d * d;
        result = result + p;
    }
    return result;
}
// dot_product_non_optimized / euclidean_distance_non_optimized (simple.rs:49-51,81-83)
static float dot_scalar(const float* a, const float* b, size_t n) {
    float s = 0.0f;
    for (size_t i = 0; i < n; ++i) {
        float p = a[i] * b[i];
        s = s + p;
    }
    return s;
}
static float euclid_scalar(const float* a, const float* b, size_t n) {
    float s = 0.0f;
    for (size_t i = 0; i < n; ++i) {
        float d0 = a[i] - b[i], d1 = a[i] - b[i];
        float p = d0 * d1;
        s = s + p;
    }
    return s;
}
// simple.rs:19-45 / 53-79 dispatch on an x86-64 host with avx+fma (the GPU box and this container)
static float dot_product(const float* a, const float* b, size_t n) {
    if (n >= 32) return dot_avx(a, b, n);
    if (n >= 16) return dot_sse(a, b, n);
    return dot_scalar(a, b, n);
}
static float euclidean_distance(const float* a, const float* b, size_t n) {
    if (n >= 32) return euclid_avx(a, b, n);
    if (n >= 16) return euclid_sse(a, b, n);
    return euclid_scalar(a, b, n);
}
// manhattan.rs:41-43 — strictly sequential scalar sum
static float manhattan_distance(const float* a, const float* b, size_t n) {
    float s = 0.0f;
    for (size_t i = 0; i < n; ++i) s = s + std::fabs(a[i] - b[i]);
    return s;
}

// ------------------------------------------------------------------------------------------
// src/unaligned_vector/binary.rs:80-94 and binary_quantized.rs:80-91 — f32 -> packed bits
// ------------------------------------------------------------------------------------------
static inline size_t n_words(size_t dims) { return (dims + 63) / 64; }
// Binary: bit = (x > 0.0) via the bits test of binary.rs:88-90
static void quantize_binary(const float* v, size_t dims, uint64_t* out) {
    size_t nw = n_words(dims);
    for (size_t w = 0; w < nw; ++w) {
        uint64_t word = 0;
        size_t lo = w * 64, hi = std::min(dims, lo + 64);
        for (size_t i = hi; i-- > lo;) {
            word <<= 1;
            uint32_t bits = f2u(v[i]);
            word += (uint64_t)(bits < 0x80000000u && bits > 0u);
        }
        out[w] = word;
    }
}
// BinaryQuantized: bit = is_sign_positive (binary_quantized.rs:84-87)
static void quantize_bq(const float* v, size_t dims, uint64_t* out) {
    size_t nw = n_words(dims);
    for (size_t w = 0; w < nw; ++w) {
        uint64_t word = 0;
        size_t lo = w * 64, hi = std::min(dims, lo + 64);
        for (size_t i = hi; i-- > lo;) {
            word <<= 1;
            word += (uint64_t)((f2u(v[i]) >> 31) == 0);
        }
        out[w] = word;
    }
}
static inline uint32_t xor_popcount(const uint64_t* a, const uint64_t* b, size_t nw) {
    uint32_t s = 0;
    for (size_t i = 0; i < nw; ++i) s += (uint32_t)__builtin_popcountll(a[i] ^ b[i]);
    return s;
}

// ------------------------------------------------------------------------------------------
// D::distance for the seven metrics.  `pa`,`pb` point at the row bytes, ha/hb are header norms.
// ------------------------------------------------------------------------------------------
static float distance(int metric, size_t dims, const void* pa, float ha, const void* pb, float hb) {
    switch (metric) {
        case EUCLIDEAN:  // euclidean.rs:40-42
            return euclidean_distance((const float*)pa, (const float*)pb, dims);
        case COSINE: {  // cosine.rs:40-56
            float pn = ha, qn = hb;
            float pq = dot_product((const float*)pa, (const float*)pb, dims);
            float pnqn = pn * qn;
            if (pnqn > 1.1920929e-07f /* f32::EPSILON */) {
                float c = pq / pnqn;
                c = c < -1.0f ? -1.0f : c;  // f32::clamp
                c = c > 1.0f ? 1.0f : c;
                return (1.0f - c) / 2.0f;
            }
            return 0.0f;
        }
        case MANHATTAN:  // manhattan.rs:41-43
            return manhattan_distance((const float*)pa, (const float*)pb, dims);
        case HAMMING: {  // hamming.rs:44-47
            size_t nw = n_words(dims);
            float d = (float)xor_popcount((const uint64_t*)pa, (const uint64_t*)pb, nw);
            return d / (float)(nw * 64);
        }
        case BQ_COSINE: {  // binary_quantized_cosine.rs:44-59 + simple.rs:119-131
            size_t nw = n_words(dims);
            int32_t h = (int32_t)xor_popcount((const uint64_t*)pa, (const uint64_t*)pb, nw);
            float pq = (float)((int32_t)(nw * 64) - 2 * h);
            float pnqn = ha * hb;
            if (pnqn != 0.0f) {
                float c = pq / pnqn;
                return (1.0f - c) / 2.0f;
            }
            return 0.0f;
        }
        case BQ_EUCLIDEAN:  // binary_quantized_euclidean.rs:76-83
            return (float)(xor_popcount((const uint64_t*)pa, (const uint64_t*)pb, n_words(dims)) * 4u);
        case BQ_MANHATTAN:  // binary_quantized_manhattan.rs:72-79
            return (float)(xor_popcount((const uint64_t*)pa, (const uint64_t*)pb, n_words(dims)) * 2u);
    }
    return NAN;
}
// D::new_header: the one f32 that matters to `distance` (norm) or 0
static float new_header(int metric, size_t dims, const void* row) {
    if (metric == COSINE) {  // cosine.rs:36-38,58-60
        const float* v = (const float*)row;
        return std::sqrt(dot_product(v, v, dims));
    }
    if (metric == BQ_COSINE) {  // binary_quantized_cosine.rs:40-42,61-63: sqrt(bq_dot(v,v)) = sqrt(L)
        size_t nw = n_words(dims);
        return std::sqrt((float)(int32_t)(nw * 64));
    }
    return 0.0f;
}

// ------------------------------------------------------------------------------------------
// Roaring portable serialization, as written by roaring-rs 0.10 (`serialize_into`): cookie
// 12346, no run containers on write; run containers (cookie 12347) accepted on read.
// Call sites: src/node.rs:143,164  src/metadata.rs:40,58  src/roaring.rs:20,29
// ------------------------------------------------------------------------------------------
static void put_u16(std::vector<uint8_t>& o, uint16_t v) { o.push_back(v & 0xff); o.push_back(v >> 8); }
static void put_u32le(std::vector<uint8_t>& o, uint32_t v) { for (int i = 0; i < 4; ++i) o.push_back((v >> (8 * i)) & 0xff); }
static void put_u32be(std::vector<uint8_t>& o, uint32_t v) { for (int i = 3; i >= 0; --i) o.push_back((v >> (8 * i)) & 0xff); }

static void roaring_serialize(const uint32_t* sorted, size_t n, std::vector<uint8_t>& out) {
    struct C { uint16_t key; size_t lo, hi; };
    std::vector<C> cs;
    for (size_t i = 0; i < n;) {
        uint16_t key = sorted[i] >> 16;
        size_t j = i;
        while (j < n && (sorted[j] >> 16) == key) ++j;
        cs.push_back({key, i, j});
        i = j;
    }
    put_u32le(out, 12346u);
    put_u32le(out, (uint32_t)cs.size());
    for (auto& c : cs) { put_u16(out, c.key); put_u16(out, (uint16_t)(c.hi - c.lo - 1)); }
    uint32_t off = 8 + 8 * (uint32_t)cs.size();
    for (auto& c : cs) {
        put_u32le(out, off);
        size_t card = c.hi - c.lo;
        off += card > 4096 ? 8192 : (uint32_t)(2 * card);
    }
    for (auto& c : cs) {
        size_t card = c.hi - c.lo;
        if (card > 4096) {
            std::vector<uint64_t> bm(1024, 0);
            for (size_t i = c.lo; i < c.hi; ++i) { uint16_t v = sorted[i] & 0xffff; bm[v >> 6] |= 1ull << (v & 63); }
            for (uint64_t w : bm) for (int b = 0; b < 8; ++b) out.push_back((w >> (8 * b)) & 0xff);
        } else {
            for (size_t i = c.lo; i < c.hi; ++i) put_u16(out, sorted[i] & 0xffff);
        }
    }
}
static bool roaring_deserialize(const uint8_t* p, size_t len, std::vector<uint32_t>& out) {
    auto rd16 = [&](size_t o) { return (uint16_t)(p[o] | (p[o + 1] << 8)); };
    auto rd32 = [&](size_t o) { return (uint32_t)p[o] | ((uint32_t)p[o + 1] << 8) | ((uint32_t)p[o + 2] << 16) | ((uint32_t)p[o + 3] << 24); };
    if (len < 4) return false;
    uint32_t cookie = rd32(0);
    size_t pos, n;
    std::vector<uint8_t> runflag;
    bool has_run = false;
    if ((cookie & 0xffff) == 12347) {
        has_run = true;
        n = (cookie >> 16) + 1;
        pos = 4;
        size_t rb = (n + 7) / 8;
        if (len < pos + rb) return false;
        runflag.assign(p + pos, p + pos + rb);
        pos += rb;
    } else if (cookie == 12346) {
        if (len < 8) return false;
        n = rd32(4);
        pos = 8;
    } else return false;
    if (len < pos + 4 * n) return false;
    std::vector<std::pair<uint16_t, uint32_t>> desc(n);
    for (size_t i = 0; i < n; ++i) { desc[i] = {rd16(pos), (uint32_t)rd16(pos + 2) + 1}; pos += 4; }
    if (!has_run || n >= 4) pos += 4 * n;  // offset header
    for (size_t i = 0; i < n; ++i) {
        uint32_t hi = (uint32_t)desc[i].first << 16, card = desc[i].second;
        bool is_run = has_run && (runflag[i / 8] >> (i % 8) & 1);
        if (is_run) {
            if (len < pos + 2) return false;
            size_t nr = rd16(pos); pos += 2;
            if (len < pos + 4 * nr) return false;
            for (size_t r = 0; r < nr; ++r) {
                uint32_t s = rd16(pos), l = rd16(pos + 2); pos += 4;
                for (uint32_t v = s; v <= s + l; ++v) out.push_back(hi | v);
            }
        } else if (card > 4096) {
            if (len < pos + 8192) return false;
            for (uint32_t w = 0; w < 1024; ++w) {
                uint64_t x = 0;
                for (int b = 0; b < 8; ++b) x |= (uint64_t)p[pos + 8 * w + b] << (8 * b);
                while (x) { int t = __builtin_ctzll(x); out.push_back(hi | (w * 64 + t)); x &= x - 1; }
            }
            pos += 8192;
        } else {
            if (len < pos + 2 * (size_t)card) return false;
            for (uint32_t k = 0; k < card; ++k) { out.push_back(hi | rd16(pos)); pos += 2; }
        }
    }
    return true;
}

// ------------------------------------------------------------------------------------------
// In-memory model of one hannoy index (what the reference keeps in LMDB under one `index: u16`)
// ------------------------------------------------------------------------------------------
using Scored = std::pair<uint32_t, uint32_t>;  // (OrderedFloat bits, id) — ordered_float.rs:25-29, hnsw.rs:30

struct Db {
    int metric = 0;
    uint32_t dims = 0;
    size_t row_bytes = 0;
    // pending adds (id -> row) until commit
    std::vector<uint32_t> ids;       // sorted ascending item ids; slot = rank
    std::vector<uint8_t> rows;       // n * row_bytes   (f32 native-endian or u64 words)
    std::vector<float> hdr;          // n   (Cosine / BQ-Cosine norm; else 0)
    std::vector<uint32_t> raw_ids;   // staging
    std::vector<uint8_t> raw_rows;
    bool committed = true;
    // graph, in SLOT space. layers[l].off has n+1 entries; has[l][slot] = a Links node exists
    struct Layer { std::vector<uint64_t> off; std::vector<uint32_t> nbr; std::vector<uint8_t> has; };
    std::vector<Layer> layers;
    std::vector<uint32_t> entry_points;  // slots, in metadata order (ascending id)
    uint32_t max_level = 0;
    bool has_metadata = false;
    // staging for set_links (id space)
    std::unordered_map<uint64_t, std::vector<uint32_t>> staged_links;  // key = level<<32 | id
    bool links_dirty = false;

    size_t n() const { return ids.size(); }
    const uint8_t* row(size_t s) const { return rows.data() + s * row_bytes; }
    int64_t slot_of(uint32_t id) const {
        auto it = std::lower_bound(ids.begin(), ids.end(), id);
        return (it != ids.end() && *it == id) ? (int64_t)(it - ids.begin()) : -1;
    }
    void commit() {
        if (committed) return;
        // merge staged adds (last write wins) with existing
        size_t m = raw_ids.size();
        std::vector<size_t> order(m);
        for (size_t i = 0; i < m; ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return raw_ids[a] < raw_ids[b]; });
        std::vector<uint32_t> nid; std::vector<uint8_t> nrows; std::vector<float> nh;
        size_t i = 0, j = 0, ne = ids.size();
        auto push_old = [&](size_t s) { nid.push_back(ids[s]); nrows.insert(nrows.end(), row(s), row(s) + row_bytes); nh.push_back(hdr[s]); };
        auto push_new = [&](size_t r) {
            nid.push_back(raw_ids[r]);
            const uint8_t* p = raw_rows.data() + r * row_bytes;
            nrows.insert(nrows.end(), p, p + row_bytes);
            nh.push_back(new_header(metric, dims, p));
        };
        while (i < ne || j < m) {
            if (j < m) {  // collapse duplicates among new: keep last
                while (j + 1 < m && raw_ids[order[j + 1]] == raw_ids[order[j]]) ++j;
            }
            if (j >= m) push_old(i++);
            else if (i >= ne) push_new(order[j++]);
            else if (ids[i] < raw_ids[order[j]]) push_old(i++);
            else if (ids[i] > raw_ids[order[j]]) push_new(order[j++]);
            else { push_new(order[j++]); ++i; }
        }
        ids.swap(nid); rows.swap(nrows); hdr.swap(nh);
        raw_ids.clear(); raw_rows.clear();
        committed = true;
    }
    void flush_links() {
        if (!links_dirty) return;
        commit();
        uint32_t L = 0;
        for (auto& kv : staged_links) L = std::max<uint32_t>(L, (uint32_t)(kv.first >> 32));
        L = std::max(L, max_level);
        size_t N = n();
        std::vector<std::vector<std::vector<uint32_t>>> tmp(L + 1);
        std::vector<std::vector<uint8_t>> has(L + 1, std::vector<uint8_t>(N, 0));
        for (uint32_t l = 0; l <= L; ++l) tmp[l].resize(N);
        for (auto& kv : staged_links) {
            uint32_t l = (uint32_t)(kv.first >> 32), id = (uint32_t)kv.first;
            int64_t s = slot_of(id);
            if (s < 0) continue;
            has[l][s] = 1;
            for (uint32_t t : kv.second) { int64_t ts = slot_of(t); if (ts >= 0) tmp[l][s].push_back((uint32_t)ts); }
            std::sort(tmp[l][s].begin(), tmp[l][s].end());
            tmp[l][s].erase(std::unique(tmp[l][s].begin(), tmp[l][s].end()), tmp[l][s].end());
        }
        layers.assign(L + 1, Layer());
        for (uint32_t l = 0; l <= L; ++l) {
            auto& ly = layers[l];
            ly.off.assign(N + 1, 0); ly.has = has[l];
            for (size_t s = 0; s < N; ++s) ly.off[s + 1] = ly.off[s] + tmp[l][s].size();
            ly.nbr.resize(ly.off[N]);
            for (size_t s = 0; s < N; ++s) std::copy(tmp[l][s].begin(), tmp[l][s].end(), ly.nbr.begin() + ly.off[s]);
        }
        links_dirty = false;
    }
};

// ------------------------------------------------------------------------------------------
// Search — src/reader.rs
// ------------------------------------------------------------------------------------------
struct Counters {  // per query
    uint64_t dist_upper = 0, dist_l0 = 0, exp_upper = 0, exp_l0 = 0, deg_upper = 0, deg_l0 = 0, flags = 0, pad = 0;
};
enum : uint64_t { FLAG_FALLBACK = 1, FLAG_LINEAR = 2, FLAG_CANCELLED = 8 };

// A deterministic `cancel_fn: impl Fn() -> bool` (reader.rs:112,171): true from its `after`-th call on (1-based);
// after == 0 never cancels (what by_vector / by_item pass, reader.rs:143).  One instance per query: the closure is
// shared by every visit of that query (reader.rs:731,817).
struct Cancel {
    uint64_t after = 0, calls = 0;
    bool operator()() { ++calls; return after != 0 && calls >= after; }
};

struct Scratch {
    std::vector<uint32_t> epoch;  // visited set (`path: RoaringBitmap`, reader.rs:734), exact
    uint32_t cur = 0;
    void reset(size_t n) {  // path.clear()
        if (epoch.size() != n) { epoch.assign(n, 0); cur = 0; }
        if (++cur == 0) { std::fill(epoch.begin(), epoch.end(), 0); cur = 1; }
    }
    bool insert(uint32_t s) { if (epoch[s] == cur) return false; epoch[s] = cur; return true; }  // RoaringBitmap::insert
    bool contains(uint32_t s) const { return epoch[s] == cur; }
};

struct Query { const void* row; float hdr; };

// MinMaxHeap<(OrderedFloat, ItemId)> as used by the reference only needs: push, push_pop_max,
// peek_max, peek_min, len, extend, drain_asc.  All keys are distinct (visited filter) so any
// correct implementation is order-equivalent; a max-heap + final sort is used here.
struct ResHeap {
    std::vector<Scored> h;  // max-heap on (bits, id)
    size_t len() const { return h.size(); }
    void push(Scored s) { h.push_back(s); std::push_heap(h.begin(), h.end()); }
    const Scored& peek_max() const { return h.front(); }
    void push_pop_max(Scored s) {
        if (s > h.front()) return;  // pushed then popped immediately
        std::pop_heap(h.begin(), h.end()); h.back() = s; std::push_heap(h.begin(), h.end());
    }
    Scored peek_min() const { return *std::min_element(h.begin(), h.end()); }
};

// Visitor::visit — src/reader.rs:301-369.  `cand` = optional candidates filter over slots,
// `excl` = slot removed from the candidates (nns_by_item, reader.rs:839-840) or UINT32_MAX.
// Returns true for Completion::Cancelled(res).  cancel == nullptr is the `&|| false` of the descent (reader.rs:736).
static bool visit(const Db& db, const Query& q, const std::vector<uint32_t>& eps, uint32_t level, size_t ef,
                  const std::vector<uint8_t>* cand, bool filter_all_but, uint32_t excl, Scratch& path,
                  ResHeap& res, Counters& ctr, Cancel* cancel = nullptr) {
    auto passes = [&](uint32_t s) {
        if (cand) return (*cand)[s] != 0 && s != excl;
        if (filter_all_but) return s != excl;
        return true;
    };
    // BinaryHeap<(Reverse<OrderedFloat>, ItemId)>: pops smallest bits; ties -> larger id first
    auto qless = [](const Scored& a, const Scored& b) { return a.first != b.first ? a.first > b.first : a.second < b.second; };
    std::priority_queue<Scored, std::vector<Scored>, decltype(qless)> search_queue(qless);
    res.h.clear();
    const Db::Layer& ly = db.layers[level];
    uint64_t& n_dist = level ? ctr.dist_upper : ctr.dist_l0;
    uint64_t& n_exp = level ? ctr.exp_upper : ctr.exp_l0;
    uint64_t& n_deg = level ? ctr.deg_upper : ctr.deg_l0;

    for (uint32_t ep : eps) {  // reader.rs:315-325
        float dist = distance(db.metric, db.dims, q.row, q.hdr, db.row(ep), db.hdr[ep]);
        ++n_dist;
        search_queue.push({f2u(dist), ep});
        path.insert(ep);
        if (passes(ep)) res.push({f2u(dist), ep});
    }
    while (!search_queue.empty()) {  // reader.rs:329-367
        if (cancel && (*cancel)()) return true;  // reader.rs:330-332
        float f = u2f(search_queue.top().first);
        float f_max = res.len() ? u2f(res.peek_max().first) : 3.40282347e+38f;
        if (f > f_max) break;
        uint32_t c = search_queue.top().second;
        search_queue.pop();
        ++n_exp;
        n_deg += ly.off[c + 1] - ly.off[c];
        for (uint64_t e = ly.off[c]; e < ly.off[c + 1]; ++e) {  // ascending id (roaring iteration)
            uint32_t point = ly.nbr[e];
            if (!path.insert(point)) continue;
            float dist = distance(db.metric, db.dims, q.row, q.hdr, db.row(point), db.hdr[point]);
            ++n_dist;
            if (res.len() < ef || dist < f_max) {  // stale f_max, live len — reader.rs:353
                search_queue.push({f2u(dist), point});
                if (!passes(point)) continue;
                if (res.len() == ef) res.push_pop_max({f2u(dist), point});
                else res.push({f2u(dist), point});
            }
        }
    }
    return false;
}

// drain_asc().take(count)
static void drain_asc_take(const Db& db, std::vector<Scored>& h, size_t count, uint32_t* out_ids, float* out_dist, uint32_t* out_len) {
    std::sort(h.begin(), h.end());
    size_t m = std::min(count, h.size());
    for (size_t i = 0; i < m; ++i) { out_ids[i] = db.ids[h[i].second]; out_dist[i] = u2f(h[i].first); }
    *out_len = (uint32_t)m;
}

struct Opts {
    size_t count, ef;  // opt.count, opt.ef (raw field: nns() default 100, ef_search() stores max(ef,count))
    const std::vector<uint8_t>* cand = nullptr;  // dense filter over slots (candidates ∩ items)
    const std::vector<uint32_t>* cand_ids = nullptr;  // the user's bitmap, ascending ids (may hold absent ids)
    size_t linear_below = 1000;
    float linear_below_ratio = 1.0f;
    size_t cand_in_db = 0;
    uint64_t cancel_after = 0;  // see Cancel
};

// should_linear_scan — reader.rs:622-640
static bool should_linear_scan(const Db& db, const Opts& o) {
    if (db.n() == 0 || !o.cand_ids) return false;
    bool below_threshold = (uint64_t)o.cand_in_db < (uint64_t)o.linear_below;
    bool below_ratio = ((float)o.cand_in_db / (float)db.n()) <= o.linear_below_ratio;
    return below_threshold && below_ratio;
}
// brute_force_search — reader.rs:668-711
static void brute_force(const Db& db, const Query& q, const Opts& o, uint32_t* out_ids, float* out_dist, uint32_t* out_len, Counters& ctr, Cancel& cancel) {
    std::vector<Scored> heap;  // BinaryHeap<(OrderedFloat, ItemId)> max-heap
    for (uint32_t id : *o.cand_ids) {
        if (cancel()) { ctr.flags |= FLAG_CANCELLED; break; }  // reader.rs:684-687
        int64_t s = db.slot_of(id);
        if (s < 0) continue;
        float d = distance(db.metric, db.dims, db.row(s), db.hdr[s], q.row, q.hdr);  // D::distance(&item, query)
        ++ctr.dist_l0;
        if (heap.size() >= o.count) {
            if (!heap.empty() && heap.front().first > f2u(d)) {  // peek.0 > OrderedFloat(distance)
                std::pop_heap(heap.begin(), heap.end());
                heap.back() = {f2u(d), (uint32_t)s};
                std::push_heap(heap.begin(), heap.end());
            }
        } else {
            heap.push_back({f2u(d), (uint32_t)s});
            std::push_heap(heap.begin(), heap.end());
        }
    }
    ctr.flags |= FLAG_LINEAR;
    drain_asc_take(db, heap, heap.size(), out_ids, out_dist, out_len);  // into_sorted_vec
}

// hnsw_search — reader.rs:722-800
static void hnsw_search(const Db& db, const Query& q, const Opts& o, Scratch& path, uint32_t* out_ids, float* out_dist, uint32_t* out_len, Counters& ctr, Cancel& cancel) {
    // return_if_cancelled! (reader.rs:749-764): only the interrupted visit's heap is returned
    auto cancelled = [&](ResHeap& r) { ctr.flags |= FLAG_CANCELLED; std::vector<Scored> f = r.h; drain_asc_take(db, f, o.count, out_ids, out_dist, out_len); };
    std::vector<uint32_t> eps = db.entry_points;
    ResHeap res;
    path.reset(db.n());
    for (uint32_t level = db.max_level; level >= 1; --level) {  // reader.rs:735-741
        visit(db, q, eps, level, 1, nullptr, false, UINT32_MAX, path, res, ctr);
        eps.assign(1, res.peek_min().second);
    }
    path.reset(db.n());  // path.clear()
    size_t ef = std::max(o.ef, o.count);
    if (visit(db, q, eps, 0, ef, o.cand, false, UINT32_MAX, path, res, ctr, &cancel)) { cancelled(res); return; }
    std::vector<Scored> neighbours = res.h;
    if (neighbours.size() < o.count) {  // reader.rs:771-795
        ctr.flags |= FLAG_FALLBACK;
        for (uint32_t s = 0; s < db.n(); ++s) {  // prefix_iter over Item keys = ascending id
            if (path.contains(s)) continue;
            eps.assign(1, s);
            size_t ef2 = o.ef > neighbours.size() ? o.ef - neighbours.size() : 0;  // saturating_sub
            if (visit(db, q, eps, 0, ef2, o.cand, false, UINT32_MAX, path, res, ctr, &cancel)) { cancelled(res); return; }
            neighbours.insert(neighbours.end(), res.h.begin(), res.h.end());
            if (neighbours.size() >= o.ef) break;
        }
    }
    drain_asc_take(db, neighbours, o.count, out_ids, out_dist, out_len);
}

// nns_by_vec — reader.rs:642-665
static void nns_by_vec(const Db& db, const Query& q, const Opts& o, Scratch& path, uint32_t* out_ids, float* out_dist, uint32_t* out_len, Counters& ctr) {
    *out_len = 0;
    Cancel cancel{o.cancel_after};
    if (db.n() == 0 || (o.cand_ids && o.cand_in_db == 0)) return;
    if (o.cand_ids && should_linear_scan(db, o)) { brute_force(db, q, o, out_ids, out_dist, out_len, ctr, cancel); return; }
    hnsw_search(db, q, o, path, out_ids, out_dist, out_len, ctr, cancel);
}

// nns_by_item — reader.rs:809-894.  Returns false for `None`.
static bool nns_by_item(const Db& db, uint32_t item, const Opts& o, Scratch& path, uint32_t* out_ids, float* out_dist, uint32_t* out_len, Counters& ctr) {
    *out_len = 0;
    if (db.n() == 0 || (o.cand_ids && o.cand_in_db == 0)) return false;
    int64_t is = db.slot_of(item);
    if (is < 0) return false;
    Query q{db.row(is), new_header(db.metric, db.dims, db.row(is))};
    Cancel cancel{o.cancel_after};
    if (o.cand_ids && should_linear_scan(db, o)) { brute_force(db, q, o, out_ids, out_dist, out_len, ctr, cancel); return true; }
    size_t ef = std::max(o.ef, o.count);
    path.reset(db.n());
    std::vector<uint32_t> eps(1, (uint32_t)is);
    ResHeap res;
    auto cancelled = [&](ResHeap& r) { ctr.flags |= FLAG_CANCELLED; std::vector<Scored> f = r.h; drain_asc_take(db, f, o.count, out_ids, out_dist, out_len); };
    if (visit(db, q, eps, 0, ef, o.cand, true, (uint32_t)is, path, res, ctr, &cancel)) { cancelled(res); return true; }
    std::vector<Scored> neighbours = res.h;
    if (neighbours.size() < o.count) {  // reader.rs:865-889
        ctr.flags |= FLAG_FALLBACK;
        for (uint32_t s = 0; s < db.n(); ++s) {
            if (path.contains(s)) continue;
            eps.assign(1, s);
            size_t ef2 = o.count - neighbours.size();
            if (visit(db, q, eps, 0, ef2, o.cand, true, (uint32_t)is, path, res, ctr, &cancel)) { cancelled(res); return true; }
            neighbours.insert(neighbours.end(), res.h.begin(), res.h.end());
            if (neighbours.size() >= o.count) break;
        }
    }
    drain_asc_take(db, neighbours, o.count, out_ids, out_dist, out_len);
    return true;
}

// ------------------------------------------------------------------------------------------
// Graph source — restatement of the build path (src/hnsw.rs) so that tests and benches have
// reference-like graphs.  Not part of the search parity claim; the reference Writer cannot run here.
// ------------------------------------------------------------------------------------------
struct Rng {  // splitmix64 (the reference uses rand's StdRng; bit-compat is not attempted)
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() { uint64_t z = (s += 0x9e3779b97f4a7c15ull); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); }
    double uniform() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
};

struct Builder {
    Db& db;
    size_t M, M0, efc;
    float alpha;
    size_t max_level = 0;
    std::vector<uint32_t> entry_points;
    // NodeState per (level, slot): links ArrayVec<[ScoredLink; M0]> (hnsw.rs:33-35).
    // level 0 is dense over slots; upper levels map slot -> row.
    struct LayerState {
        std::vector<int32_t> row_of;       // slot -> row (or -1) ; level 0: identity, omitted
        std::vector<Scored> links;         // rows * M0
        std::vector<uint16_t> len;         // rows
        std::vector<uint8_t> present;      // rows (level 0: by slot)
        size_t rows = 0;
    };
    std::vector<LayerState> layers;
    static constexpr size_t NLOCK = 1 << 14;
    std::vector<std::mutex> locks;

    Builder(Db& d, size_t m, size_t m0, size_t ef, float a) : db(d), M(m), M0(m0), efc(ef), alpha(a), locks(NLOCK) {}

    // get_default_probas — hnsw.rs:94-110
    std::vector<float> probas() const {
        std::vector<float> p;
        float level_factor = 1.0f / std::log((float)M + 1.1920929e-07f);
        for (int level = 0;; ++level) {
            float pr = std::exp((float)level * (-1.0f / level_factor)) * (1.0f - std::exp(-1.0f / level_factor));
            if (pr < 1e-09f) break;
            p.push_back(pr);
        }
        return p;
    }
    float dist_ss(uint32_t a, uint32_t b) const { return distance(db.metric, db.dims, db.row(a), db.hdr[a], db.row(b), db.hdr[b]); }

    int64_t row(size_t lvl, uint32_t slot) const { return lvl == 0 ? (int64_t)slot : (int64_t)layers[lvl].row_of[slot]; }
    std::mutex& lock(size_t lvl, uint32_t slot) { return locks[(slot * 0x9e3779b1u + (uint32_t)lvl * 0x85ebca6bu) & (NLOCK - 1)]; }

    // get_neighbours — hnsw.rs:428-456 (fresh build: nothing in LMDB, only self.layers)
    void get_neighbours(uint32_t slot, size_t lvl, std::vector<uint32_t>& out) {
        out.clear();
        int64_t r = row(lvl, slot);
        if (r < 0) return;
        std::lock_guard<std::mutex> g(lock(lvl, slot));
        auto& ls = layers[lvl];
        for (size_t i = 0; i < ls.len[r]; ++i) out.push_back(ls.links[r * M0 + i].second);
    }
    // walk_layer — hnsw.rs:460-518
    void walk_layer(uint32_t q, const std::vector<uint32_t>& eps, size_t lvl, size_t ef, Scratch& visited, std::vector<Scored>& out, std::vector<uint32_t>& nb) {
        auto qless = [](const Scored& a, const Scored& b) { return a.first != b.first ? a.first > b.first : a.second < b.second; };
        std::priority_queue<Scored, std::vector<Scored>, decltype(qless)> cands(qless);
        ResHeap res;
        visited.reset(db.n());
        for (uint32_t ep : eps) {
            float d = dist_ss(q, ep);
            cands.push({f2u(d), ep});
            res.push({f2u(d), ep});
            visited.insert(ep);
        }
        while (!cands.empty()) {
            float f = u2f(cands.top().first);
            float f_max = u2f(res.peek_max().first);
            if (f > f_max) break;
            uint32_t c = cands.top().second;
            cands.pop();
            get_neighbours(c, lvl, nb);
            for (uint32_t point : nb) {
                if (!visited.insert(point)) continue;
                float d = dist_ss(q, point);
                if (res.len() < ef || d < f_max) {
                    cands.push({f2u(d), point});
                    if (res.len() == ef) res.push_pop_max({f2u(d), point});
                    else res.push({f2u(d), point});
                }
            }
        }
        out = res.h;
    }
    // robust_prune — hnsw.rs:565-597
    std::vector<Scored> robust_prune(std::vector<Scored> cands, size_t level) const {
        size_t cap = level == 0 ? M0 : M;
        std::sort(cands.begin(), cands.end(), [](const Scored& a, const Scored& b) { return b < a; });
        std::vector<Scored> selected;
        selected.reserve(cap);
        while (!cands.empty()) {
            Scored c = cands.back();
            cands.pop_back();
            if (selected.size() == cap) break;
            bool ok = true;
            for (auto& s : selected) {
                float d = dist_ss(c.second, s.second);
                if (f2u(d * alpha) < c.first) { ok = false; break; }
            }
            if (ok) selected.push_back(c);
        }
        return selected;
    }
    // add_link — hnsw.rs:523-560
    void add_link(uint32_t p, Scored q, size_t lvl) {
        if (p == q.second) return;
        if (lvl >= layers.size()) return;
        int64_t r = row(lvl, p);
        if (r < 0) return;  // cannot happen: every visited node lives on this level
        auto& ls = layers[lvl];
        size_t cap = lvl == 0 ? M0 : M;
        // papaya's update_or_insert_with applies the pure update atomically per key: hold the stripe lock
        std::lock_guard<std::mutex> g(lock(lvl, p));
        ls.present[r] = 1;
        if (ls.len[r] < cap) { ls.links[r * M0 + ls.len[r]++] = q; return; }
        std::vector<Scored> cur(ls.links.begin() + r * M0, ls.links.begin() + r * M0 + ls.len[r]);
        std::vector<Scored> pruned = robust_prune(cur, lvl);  // q is NOT added when full (hnsw.rs:542-552)
        for (size_t i = 0; i < pruned.size(); ++i) ls.links[r * M0 + i] = pruned[i];
        ls.len[r] = (uint16_t)pruned.size();
    }
    // insert — hnsw.rs:291-328
    void insert(uint32_t query, size_t level, Scratch& visited) {
        std::vector<uint32_t> eps = entry_points, nb;
        std::vector<Scored> neighbours;
        for (size_t lvl = max_level; lvl >= level + 1; --lvl) {
            walk_layer(query, eps, lvl, 1, visited, neighbours, nb);
            eps.assign(1, std::min_element(neighbours.begin(), neighbours.end())->second);
        }
        for (size_t lvl = level + 1; lvl-- > 0;) {
            walk_layer(query, eps, lvl, efc, visited, neighbours, nb);
            eps.clear();
            for (auto& sn : robust_prune(neighbours, level)) {  // NB: cap from the item's top level (hnsw.rs:317)
                add_link(query, sn, lvl);
                add_link(sn.second, {sn.first, query}, lvl);
                eps.push_back(sn.second);
            }
        }
    }
    void build(uint64_t seed, int n_threads) {
        size_t N = db.n();
        db.layers.clear(); db.entry_points.clear(); db.max_level = 0; db.has_metadata = true;
        if (N == 0) return;
        // sample levels — hnsw.rs:113-119,142-149 (WeightedIndex over assign_probas)
        std::vector<float> pr = probas();
        std::vector<double> cdf(pr.size());
        double tot = 0; for (float p : pr) tot += p;
        double acc = 0; for (size_t i = 0; i < pr.size(); ++i) { acc += pr[i] / tot; cdf[i] = acc; }
        Rng rng(seed);
        std::vector<std::pair<uint32_t, size_t>> levels(N);
        size_t cur_max = 0;
        for (size_t s = 0; s < N; ++s) {
            double u = rng.uniform();
            size_t l = std::lower_bound(cdf.begin(), cdf.end(), u) - cdf.begin();
            l = std::min(l, pr.size() - 1);
            levels[s] = {(uint32_t)s, l};
            cur_max = std::max(cur_max, l);
        }
        // prepare_levels_and_entry_points — hnsw.rs:222-289 (fresh build: no old entry points)
        std::stable_sort(levels.begin(), levels.end(), [](auto& a, auto& b) { return a.second > b.second; });
        max_level = cur_max;
        layers.assign(max_level + 1, LayerState());
        for (size_t l = 0; l <= max_level; ++l) {
            auto& ls = layers[l];
            if (l == 0) ls.rows = N;
            else {
                ls.row_of.assign(N, -1);
                for (auto& il : levels) if (il.second >= l) ls.row_of[il.first] = (int32_t)ls.rows++;
            }
            ls.links.assign(ls.rows * M0, Scored{0, 0});
            ls.len.assign(ls.rows, 0);
            ls.present.assign(ls.rows, 0);
        }
        for (auto& il : levels) {
            if (il.second != max_level) break;
            entry_points.push_back(il.first);
        }
        std::sort(entry_points.begin(), entry_points.end());
        // add_in_layers_below for every item happens in insert(); a Links node exists on each level <= item level
        for (auto& il : levels) for (size_t l = 0; l <= il.second; ++l) layers[l].present[row(l, il.first)] = 1;
        // insert level groups top-down; inside a group in parallel (hnsw.rs:160-185)
        size_t g0 = 0;
        while (g0 < N) {
            size_t g1 = g0;
            while (g1 < N && levels[g1].second == levels[g0].second) ++g1;
            int nt = std::max(1, std::min<int>(n_threads, (int)(g1 - g0)));
            if (nt == 1) {
                Scratch sc;
                for (size_t i = g0; i < g1; ++i) insert(levels[i].first, levels[i].second, sc);
            } else {
                std::atomic<size_t> next(g0);
                std::vector<std::thread> th;
                for (int t = 0; t < nt; ++t) th.emplace_back([&] {
                    Scratch sc;
                    for (;;) { size_t i = next.fetch_add(1); if (i >= g1) break; insert(levels[i].first, levels[i].second, sc); }
                });
                for (auto& t : th) t.join();
            }
            g0 = g1;
        }
        // write Links nodes: RoaringBitmap::from_iter(links ids) — hnsw.rs:195-213
        db.layers.assign(max_level + 1, Db::Layer());
        for (size_t l = 0; l <= max_level; ++l) {
            auto& out = db.layers[l];
            auto& ls = layers[l];
            out.off.assign(N + 1, 0); out.has.assign(N, 0);
            std::vector<uint32_t> tmp;
            for (size_t s = 0; s < N; ++s) {
                int64_t r = row(l, (uint32_t)s);
                out.off[s + 1] = out.off[s];
                if (r < 0 || !ls.present[r]) continue;
                out.has[s] = 1;
                tmp.clear();
                for (size_t i = 0; i < ls.len[r]; ++i) tmp.push_back(ls.links[r * M0 + i].second);
                std::sort(tmp.begin(), tmp.end());
                tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
                out.nbr.insert(out.nbr.end(), tmp.begin(), tmp.end());
                out.off[s + 1] = out.nbr.size();
            }
        }
        db.entry_points = entry_points;
        db.max_level = (uint32_t)max_level;
    }
};

}  // namespace orc

// ==========================================================================================
// C interface for ctypes (tests / bench only)
// ==========================================================================================
using namespace orc;
extern "C" {

void* orc_db_new(int metric, uint32_t dims) {
    if (metric < 0 || metric > 6) return nullptr;
    Db* db = new Db();
    db->metric = metric; db->dims = dims;
    db->row_bytes = is_binary(metric) ? 8 * n_words(dims) : 4 * (size_t)dims;
    return db;
}
void orc_db_free(void* h) { delete (Db*)h; }
const char* orc_metric_name(int m) { return metric_name(m); }

// Writer::add_item equivalent: encode with the metric's codec (from_slice), header via new_header
int orc_db_add_items(void* h, const uint32_t* ids, uint64_t n, const float* vecs) {
    Db& db = *(Db*)h;
    size_t base = db.raw_ids.size();
    db.raw_ids.insert(db.raw_ids.end(), ids, ids + n);
    db.raw_rows.resize((base + n) * db.row_bytes);
    for (uint64_t i = 0; i < n; ++i) {
        uint8_t* dst = db.raw_rows.data() + (base + i) * db.row_bytes;
        const float* v = vecs + i * db.dims;
        if (db.metric == HAMMING) quantize_binary(v, db.dims, (uint64_t*)dst);
        else if (is_binary(db.metric)) quantize_bq(v, db.dims, (uint64_t*)dst);
        else std::memcpy(dst, v, db.row_bytes);
    }
    db.committed = false;
    return 0;
}
// add pre-encoded rows (e.g. binary codes generated on the GPU)
int orc_db_add_rows(void* h, const uint32_t* ids, uint64_t n, const void* rows) {
    Db& db = *(Db*)h;
    size_t base = db.raw_ids.size();
    db.raw_ids.insert(db.raw_ids.end(), ids, ids + n);
    db.raw_rows.resize((base + n) * db.row_bytes);
    std::memcpy(db.raw_rows.data() + base * db.row_bytes, rows, n * db.row_bytes);
    db.committed = false;
    return 0;
}
int orc_db_build(void* h, uint32_t M, uint32_t M0, uint32_t efc, float alpha, uint64_t seed, int n_threads) {
    Db& db = *(Db*)h;
    db.commit();
    db.staged_links.clear(); db.links_dirty = false;
    if (M0 > 65535 || M > M0) return 1;
    Builder b(db, M, M0, efc, alpha);
    b.build(seed, n_threads);
    return 0;
}
// install a graph directly (golden topologies from the reference's insta snapshots)
int orc_db_set_links(void* h, uint32_t id, uint32_t level, const uint32_t* nbrs, uint32_t n) {
    Db& db = *(Db*)h;
    db.staged_links[((uint64_t)level << 32) | id] = std::vector<uint32_t>(nbrs, nbrs + n);
    db.links_dirty = true;
    return 0;
}
// install one whole layer as CSR (offsets in slot order, neighbours as item ids) — graph cache reload
int orc_db_set_csr(void* h, uint32_t level, const uint64_t* off, const uint32_t* nbr_ids, uint64_t nnz) {
    Db& db = *(Db*)h;
    db.commit();
    size_t N = db.n();
    if (db.layers.size() <= level) db.layers.resize(level + 1);
    for (auto& ly : db.layers) if (ly.off.size() != N + 1) { ly.off.assign(N + 1, 0); ly.has.assign(N, 0); ly.nbr.clear(); }
    auto& ly = db.layers[level];
    ly.off.assign(off, off + N + 1);
    ly.nbr.resize(nnz);
    for (uint64_t i = 0; i < nnz; ++i) { int64_t s = db.slot_of(nbr_ids[i]); if (s < 0) return 1; ly.nbr[i] = (uint32_t)s; }
    for (size_t s = 0; s < N; ++s) ly.has[s] = ly.off[s + 1] > ly.off[s];
    return 0;
}
int orc_db_set_entry_points(void* h, const uint32_t* eps, uint32_t n, uint32_t max_level) {
    Db& db = *(Db*)h;
    db.commit();
    db.entry_points.clear();
    for (uint32_t i = 0; i < n; ++i) { int64_t s = db.slot_of(eps[i]); if (s < 0) return 1; db.entry_points.push_back((uint32_t)s); }
    db.max_level = max_level; db.has_metadata = true;
    return 0;
}
static void prep(Db& db) { db.commit(); db.flush_links(); while (db.layers.size() <= db.max_level) { Db::Layer l; l.off.assign(db.n() + 1, 0); l.has.assign(db.n(), 0); db.layers.push_back(l); } }

uint64_t orc_db_n_items(void* h) { Db& db = *(Db*)h; db.commit(); return db.n(); }
uint64_t orc_db_row_bytes(void* h) { return ((Db*)h)->row_bytes; }
uint32_t orc_db_max_level(void* h) { return ((Db*)h)->max_level; }
uint32_t orc_db_n_entry_points(void* h) { return (uint32_t)((Db*)h)->entry_points.size(); }
void orc_db_get_entry_points(void* h, uint32_t* out) { Db& db = *(Db*)h; for (size_t i = 0; i < db.entry_points.size(); ++i) out[i] = db.ids[db.entry_points[i]]; }
void orc_db_get_ids(void* h, uint32_t* out) { Db& db = *(Db*)h; db.commit(); std::copy(db.ids.begin(), db.ids.end(), out); }
void orc_db_get_rows(void* h, void* out) { Db& db = *(Db*)h; db.commit(); std::memcpy(out, db.rows.data(), db.rows.size()); }
void orc_db_get_headers(void* h, float* out) { Db& db = *(Db*)h; db.commit(); std::copy(db.hdr.begin(), db.hdr.end(), out); }
uint32_t orc_db_n_layers(void* h) { Db& db = *(Db*)h; prep(db); return (uint32_t)db.layers.size(); }
uint64_t orc_db_layer_nnz(void* h, uint32_t l) { Db& db = *(Db*)h; prep(db); return db.layers[l].nbr.size(); }
// CSR of one layer in slot order; neighbours as ITEM IDS (ascending)
void orc_db_get_layer(void* h, uint32_t l, uint64_t* off, uint32_t* nbr_ids) {
    Db& db = *(Db*)h; prep(db);
    auto& ly = db.layers[l];
    std::copy(ly.off.begin(), ly.off.end(), off);
    for (size_t i = 0; i < ly.nbr.size(); ++i) nbr_ids[i] = db.ids[ly.nbr[i]];
}

// --- search -------------------------------------------------------------------------------
}  // extern "C"
struct CandPrep { std::vector<uint8_t> dense; std::vector<uint32_t> ids; size_t in_db = 0; };
static void prep_cand(const Db& db, const uint32_t* cand, uint64_t n_cand, CandPrep& cp) {
    cp.ids.assign(cand, cand + n_cand);
    std::sort(cp.ids.begin(), cp.ids.end());
    cp.ids.erase(std::unique(cp.ids.begin(), cp.ids.end()), cp.ids.end());
    cp.dense.assign(db.n(), 0);
    for (uint32_t id : cp.ids) { int64_t s = db.slot_of(id); if (s >= 0) { cp.dense[s] = 1; ++cp.in_db; } }
}
template <class F> static void par_for(uint64_t n, int n_threads, F f) {
    int nt = std::max(1, (int)std::min<uint64_t>(n_threads, n));
    if (nt == 1) { Scratch sc; for (uint64_t i = 0; i < n; ++i) f(i, sc); return; }
    std::atomic<uint64_t> next(0);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back([&] { Scratch sc; for (;;) { uint64_t i = next.fetch_add(1); if (i >= n) break; f(i, sc); } });
    for (auto& t : th) t.join();
}
extern "C" {
// reader.nns(count).ef_search(..).candidates(..).linear_below(..).by_vector(q) for a batch.
// `ef` is the raw QueryBuilder.ef field.  cand==NULL <=> no candidates bitmap.
// out_ids/out_dist: nq*count, out_len: nq, counters: nq*8 u64 (nullable)
static thread_local uint64_t g_cancel_after = 0;
// the next orc_search_* call of this thread runs with `cancel_fn` = "true from the after-th call on" (0 = never); the
// FLAG_CANCELLED bit of the counters' flags word reports Completion::Cancelled per query
void orc_set_cancel_after(uint64_t after) { g_cancel_after = after; }
int orc_search_by_vector(void* h, const float* q, uint64_t nq, uint32_t count, uint32_t ef, const uint32_t* cand, uint64_t n_cand,
                         int has_cand, uint32_t linear_below, float linear_ratio, uint32_t* out_ids, float* out_dist, uint32_t* out_len,
                         uint64_t* counters, int n_threads) {
    Db& db = *(Db*)h; prep(db);
    CandPrep cp;
    Opts o; o.count = count; o.ef = ef; o.linear_below = linear_below; o.linear_below_ratio = linear_ratio;
    o.cancel_after = g_cancel_after; g_cancel_after = 0;
    if (has_cand) { prep_cand(db, cand, n_cand, cp); o.cand = &cp.dense; o.cand_ids = &cp.ids; o.cand_in_db = cp.in_db; }
    size_t nw = n_words(db.dims);
    par_for(nq, n_threads, [&](uint64_t i, Scratch& sc) {
        const float* v = q + i * db.dims;
        std::vector<uint64_t> code;
        Query qq;
        if (is_binary(db.metric)) {  // UnalignedVector::from_slice(vector) — reader.rs:140
            code.resize(nw);
            if (db.metric == HAMMING) quantize_binary(v, db.dims, code.data()); else quantize_bq(v, db.dims, code.data());
            qq.row = code.data();
        } else qq.row = v;
        qq.hdr = new_header(db.metric, db.dims, qq.row);  // reader.rs:141
        Counters c;
        nns_by_vec(db, qq, o, sc, out_ids + i * count, out_dist + i * count, out_len + i, c);
        if (counters) std::memcpy(counters + i * 8, &c, sizeof(c));
    });
    return 0;
}
// by_item for a batch; out_len[i] = UINT32_MAX encodes `None`
int orc_search_by_item(void* h, const uint32_t* items, uint64_t nq, uint32_t count, uint32_t ef, const uint32_t* cand, uint64_t n_cand,
                       int has_cand, uint32_t linear_below, float linear_ratio, uint32_t* out_ids, float* out_dist, uint32_t* out_len,
                       uint64_t* counters, int n_threads) {
    Db& db = *(Db*)h; prep(db);
    CandPrep cp;
    Opts o; o.count = count; o.ef = ef; o.linear_below = linear_below; o.linear_below_ratio = linear_ratio;
    o.cancel_after = g_cancel_after; g_cancel_after = 0;
    if (has_cand) { prep_cand(db, cand, n_cand, cp); o.cand = &cp.dense; o.cand_ids = &cp.ids; o.cand_in_db = cp.in_db; }
    par_for(nq, n_threads, [&](uint64_t i, Scratch& sc) {
        Counters c;
        bool some = nns_by_item(db, items[i], o, sc, out_ids + i * count, out_dist + i * count, out_len + i, c);
        if (!some) out_len[i] = UINT32_MAX;
        if (counters) std::memcpy(counters + i * 8, &c, sizeof(c));
    });
    return 0;
}
// exact kNN in the index metric (ground truth for recall), ties by (bits, id)
int orc_exact_knn(void* h, const float* q, uint64_t nq, uint32_t k, uint32_t* out_ids, float* out_dist, int n_threads) {
    Db& db = *(Db*)h; prep(db);
    size_t nw = n_words(db.dims);
    par_for(nq, n_threads, [&](uint64_t i, Scratch&) {
        const float* v = q + i * db.dims;
        std::vector<uint64_t> code; Query qq;
        if (is_binary(db.metric)) { code.resize(nw); if (db.metric == HAMMING) quantize_binary(v, db.dims, code.data()); else quantize_bq(v, db.dims, code.data()); qq.row = code.data(); }
        else qq.row = v;
        qq.hdr = new_header(db.metric, db.dims, qq.row);
        std::vector<Scored> heap;
        for (uint32_t s = 0; s < db.n(); ++s) {
            Scored x{f2u(distance(db.metric, db.dims, qq.row, qq.hdr, db.row(s), db.hdr[s])), s};
            if (heap.size() < k) { heap.push_back(x); std::push_heap(heap.begin(), heap.end()); }
            else if (k && x < heap.front()) { std::pop_heap(heap.begin(), heap.end()); heap.back() = x; std::push_heap(heap.begin(), heap.end()); }
        }
        std::sort(heap.begin(), heap.end());
        for (size_t j = 0; j < k; ++j) {
            out_ids[i * k + j] = j < heap.size() ? db.ids[heap[j].second] : UINT32_MAX;
            out_dist[i * k + j] = j < heap.size() ? u2f(heap[j].first) : INFINITY;
        }
    });
    return 0;
}

// --- primitives for known-answer tests ---------------------------------------------------------
float orc_dot_product(const float* a, const float* b, uint64_t n) { return dot_product(a, b, n); }
float orc_euclidean(const float* a, const float* b, uint64_t n) { return euclidean_distance(a, b, n); }
float orc_dot_scalar(const float* a, const float* b, uint64_t n) { return dot_scalar(a, b, n); }
float orc_euclid_scalar(const float* a, const float* b, uint64_t n) { return euclid_scalar(a, b, n); }
float orc_dot_sse(const float* a, const float* b, uint64_t n) { return dot_sse(a, b, n); }
float orc_euclid_sse(const float* a, const float* b, uint64_t n) { return euclid_sse(a, b, n); }
void orc_quantize(int binary_codec /*1=Binary,0=BinaryQuantized*/, const float* v, uint64_t dims, uint64_t* out) {
    if (binary_codec) quantize_binary(v, dims, out); else quantize_bq(v, dims, out);
}
// D::distance on two f32 vectors (encoded with the metric's codec first)
float orc_distance(int metric, const float* a, const float* b, uint32_t dims) {
    if (is_binary(metric)) {
        size_t nw = n_words(dims);
        std::vector<uint64_t> ca(nw), cb(nw);
        if (metric == HAMMING) { quantize_binary(a, dims, ca.data()); quantize_binary(b, dims, cb.data()); }
        else { quantize_bq(a, dims, ca.data()); quantize_bq(b, dims, cb.data()); }
        return distance(metric, dims, ca.data(), new_header(metric, dims, ca.data()), cb.data(), new_header(metric, dims, cb.data()));
    }
    return distance(metric, dims, a, new_header(metric, dims, a), b, new_header(metric, dims, b));
}
int orc_ordered_float_cmp(float a, float b) { uint32_t x = f2u(a), y = f2u(b); return x < y ? -1 : (x > y ? 1 : 0); }

static std::vector<uint8_t> g_buf;
uint64_t orc_roaring_serialize(const uint32_t* sorted, uint64_t n, uint8_t* out, uint64_t cap) {
    std::vector<uint8_t> b; roaring_serialize(sorted, n, b);
    if (out && cap >= b.size()) std::memcpy(out, b.data(), b.size());
    return b.size();
}
int64_t orc_roaring_deserialize(const uint8_t* p, uint64_t len, uint32_t* out, uint64_t cap) {
    std::vector<uint32_t> v; if (!roaring_deserialize(p, len, v)) return -1;
    if (out && cap >= v.size()) std::copy(v.begin(), v.end(), out);
    return (int64_t)v.size();
}

// --- export in the reference's on-disk encoding (what a heed cursor over the LMDB would yield) ----
// Stream layout: repeated [klen u32 LE][key][vlen u32 LE][value], keys in LMDB (bytewise) order.
// key  = [index u16 BE][mode u8][item u32 BE][layer u8]              src/key.rs:54-66
// meta = name\0 | dims u32 BE | size u32 BE | roaring(items) | entry points (native u32) | max_level u8   metadata.rs:22-46
// vers = 3 x u32 BE                                               src/version.rs:33-46
// item = [0] | header | vector bytes ;  links = [1] | roaring      src/node.rs:130-149
static void put_key(std::vector<uint8_t>& o, uint16_t index, uint8_t mode, uint32_t item, uint8_t layer) {
    put_u32le(o, 8); o.push_back(index >> 8); o.push_back(index & 0xff); o.push_back(mode); put_u32be(o, item); o.push_back(layer);
}
uint64_t orc_db_export_kv(void* h, uint16_t index, uint8_t* out, uint64_t cap) {
    Db& db = *(Db*)h; prep(db);
    std::vector<uint8_t> o;
    size_t N = db.n();
    {  // metadata
        std::vector<uint8_t> v;
        const char* nm = metric_name(db.metric);
        v.insert(v.end(), nm, nm + std::strlen(nm)); v.push_back(0);
        put_u32be(v, db.dims);
        std::vector<uint8_t> rb; roaring_serialize(db.ids.data(), N, rb);
        put_u32be(v, (uint32_t)rb.size()); v.insert(v.end(), rb.begin(), rb.end());
        for (uint32_t ep : db.entry_points) put_u32le(v, db.ids[ep]);
        v.push_back((uint8_t)db.max_level);
        put_key(o, index, 0, 0, 0); put_u32le(o, (uint32_t)v.size()); o.insert(o.end(), v.begin(), v.end());
        std::vector<uint8_t> ver; put_u32be(ver, 0); put_u32be(ver, 1); put_u32be(ver, 3);
        put_key(o, index, 0, 1, 0); put_u32le(o, 12); o.insert(o.end(), ver.begin(), ver.end());
    }
    for (size_t s = 0; s < N; ++s)  // links, ordered by (item, layer)
        for (size_t l = 0; l < db.layers.size(); ++l) {
            auto& ly = db.layers[l];
            if (!ly.has[s]) continue;
            std::vector<uint32_t> nb;
            for (uint64_t e = ly.off[s]; e < ly.off[s + 1]; ++e) nb.push_back(db.ids[ly.nbr[e]]);
            std::vector<uint8_t> v; v.push_back(1); roaring_serialize(nb.data(), nb.size(), v);
            put_key(o, index, 2, db.ids[s], (uint8_t)l); put_u32le(o, (uint32_t)v.size()); o.insert(o.end(), v.begin(), v.end());
        }
    size_t hb = header_bytes(db.metric);
    for (size_t s = 0; s < N; ++s) {  // items
        put_key(o, index, 3, db.ids[s], 0);
        put_u32le(o, (uint32_t)(1 + hb + db.row_bytes));
        o.push_back(0);
        if (hb == 8) { for (int i = 0; i < 8; ++i) o.push_back(0); }
        else { uint32_t u = f2u(db.hdr[s]); put_u32le(o, u); }
        o.insert(o.end(), db.row(s), db.row(s) + db.row_bytes);
    }
    if (out && cap >= o.size()) std::memcpy(out, o.data(), o.size());
    return o.size();
}

}  // extern "C"
