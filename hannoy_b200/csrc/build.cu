// build.cu — HNSW graph construction on the device (the reference's stated missing feature,
// README.md:23-24), i.e. `HnswBuilder::build` (src/hnsw.rs:122-216) with its `insert` (hnsw.rs:291-328),
// `walk_layer` (hnsw.rs:460-519), `robust_prune` (hnsw.rs:567-597) and `add_link` (hnsw.rs:524-562) run for a whole
// batch of items at a time:
//
//   levels    sampled on the host from the reference's distribution (hnsw.rs:94-120), items inserted group by group
//             from the top level down to level 0 (hnsw.rs:158,170-184); the items of the top level are the entry
//             points and are registered on every layer before anything is linked (hnsw.rs:268-279);
//   search    build_search_kernel (search.cu): one warp per item, the reader's visit() — greedy descent with ef = 1,
//             then the ef_construction walk of the layer being linked;
//   prune     prune_link_kernel (here): one warp per item — robust_prune over the walk's result, distances between
//             candidates recomputed by the whole warp, `OrderedFloat(d * alpha) < dist_to_query` on bit patterns;
//   link      add_link (hnsw.rs:524-562) in both directions without locks: the item's own list is written by its warp
//             (nobody else can reach an item that is not linked yet), the reverse links are posted into per-target
//             inboxes and applied by apply_reverse_kernel, one warp per touched target, sources in ascending order — so
//             the graph is a deterministic function of (items, seed, batch schedule).  A full list is re-pruned and the
//             new link dropped, exactly the reference's behaviour (hnsw.rs:541-553 prunes `links`, not `links + q`).
//
// The reference inserts the items of one level group concurrently (rayon, hnsw.rs:171-184): items in flight do not see
// each other, links land in arrival order.  A batch here is the same thing with more items in flight; the batch size
// grows with the graph (never more than 1/64 of what is already linked: 1/16 costs recall on some data, measured) so that
// early items still see each other.
// Like the reference's parallel build the result depends on scheduling; it is a valid hannoy graph, not a bit-copy of
// any particular CPU run.  Graph quality is tested as recall against the restated sequential builder.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <chrono>
#include <numeric>

#include "dist.cuh"

namespace hb {

namespace {

#define CUDA_TRYB(expr)                                                                  \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            set_error("%s failed: %s", #expr, cudaGetErrorString(_e));                   \
            return HB_ECUDA;                                                             \
        }                                                                                \
    } while (0)

constexpr uint32_t INBOX_K = 32;
constexpr int LIST_MAX = (int)FIXED_DEG_MAX;  // longest neighbour list (M0 <= 64); a lane handles entries lane and lane + 32

struct GraphDev {
    uint32_t n = 0, M = 0, M0 = 0;
    uint32_t stride0 = FIXED_DEG;     // layer-0 cells per item: 32, or 64 when 32 < M0 <= 64 (= DevIndex::nbr0_stride)
    uint32_t* nbr[MAX_LEVELS] = {};   // [n x stride(l)] neighbour slots in arrival order, UINT32_MAX padded
    float* dist[MAX_LEVELS] = {};     // [n x stride(l)] distance owner -> neighbour (ScoredLink, hnsw.rs:30)
    uint32_t* deg[MAX_LEVELS] = {};   // [n]
    // reverse links of the batch in flight: target t receives (source, distance) records in inbox[t * INBOX_K ..]
    uint32_t* inbox_cnt = nullptr;    // [n] records posted to t (may exceed INBOX_K: the excess is dropped and counted)
    unsigned long long* inbox = nullptr;  // [n x INBOX_K] (distance bits << 32 | source slot)
    uint32_t* touched = nullptr;      // targets with a non-empty inbox
    uint32_t* n_touched = nullptr;
    unsigned long long* n_dropped = nullptr;
    __host__ __device__ uint32_t stride(uint32_t l) const { return l == 0 ? stride0 : M; }
    __host__ __device__ uint32_t cap(uint32_t l) const { return l == 0 ? M0 : M; }
};

// first internal inconsistency seen by a kernel: {code, a, b, c, d}; the host turns it into an error instead of the
// device faulting on a wild pointer
__device__ unsigned int g_build_err[8];
__device__ __forceinline__ void build_fail(unsigned code, unsigned a, unsigned b, unsigned c, unsigned d) {
    if (lane_id() == 0 && atomicCAS(&g_build_err[0], 0u, code) == 0u) { g_build_err[1] = a; g_build_err[2] = b; g_build_err[3] = c; g_build_err[4] = d; }
}

struct PruneLinkParams {
    DevIndex ix;
    GraphDev g;
    const uint32_t* items = nullptr;   // batch slots
    uint32_t n_items = 0;
    uint32_t level = 0;                // layer being linked
    uint32_t own_cap = 0;              // robust_prune(neighbours, level_of_the_item, ..): M0 for level-0 items, else M (hnsw.rs:318)
    float alpha = 1.0f;
    const unsigned long long* cand = nullptr;
    const uint32_t* cand_len = nullptr;
    uint32_t efc = 0;
    uint32_t* sel_out = nullptr;       // [n_items][32]: the selected neighbours = entry points of the next layer down (hnsw.rs:323)
};

// D::distance between stored items, by the whole warp (any row kind; not the reference's summation order — the builder has
// no bit-parity contract, only the reader has).
// Distances from item `a` to up to PR_GROUP other items at once: the loads of all rows are in flight together, so a
// candidate costs one memory round trip per PR_GROUP selected points instead of one per point.
constexpr int PR_GROUP = 4;
__device__ void warp_pair_distance_group(const DevIndex& ix, uint32_t a, const uint32_t* others, int g, float (&out)[PR_GROUP]) {
    const int lane = lane_id();
    uint32_t b[PR_GROUP];
#pragma unroll
    for (int r = 0; r < PR_GROUP; ++r) b[r] = others[r < g ? r : 0];
    bool bad = a >= ix.n;
#pragma unroll
    for (int r = 0; r < PR_GROUP; ++r) bad = bad || b[r] >= ix.n;
    if (bad) {
        build_fail(1, a, b[0], (unsigned)g, 0);
#pragma unroll
        for (int r = 0; r < PR_GROUP; ++r) out[r] = 0.0f;
        return;
    }
    const uint8_t* ra = ix.rows + (size_t)a * ix.row_stride;
    const uint8_t* rb[PR_GROUP];
#pragma unroll
    for (int r = 0; r < PR_GROUP; ++r) rb[r] = ix.rows + (size_t)b[r] * ix.row_stride;
    const uint32_t words = ix.row_stride / 16;
    if (ix.kind == KIND_BIN) {
        uint32_t h[PR_GROUP] = {};
        for (uint32_t w = lane; w < words; w += 32) {
            const ulonglong2 x = __ldg(reinterpret_cast<const ulonglong2*>(ra) + w);
            ulonglong2 y[PR_GROUP];
#pragma unroll
            for (int r = 0; r < PR_GROUP; ++r) y[r] = __ldg(reinterpret_cast<const ulonglong2*>(rb[r]) + w);
#pragma unroll
            for (int r = 0; r < PR_GROUP; ++r) h[r] += __popcll(x.x ^ y[r].x) + __popcll(x.y ^ y[r].y);
        }
        __syncwarp();
        const float na = ix.metric == HB_BQ_COSINE ? __ldg(&ix.hdr[a]) : 0.0f;
#pragma unroll
        for (int r = 0; r < PR_GROUP; ++r) {
            const uint32_t hr = __reduce_add_sync(FULL, h[r]);
            out[r] = finish_bin(ix.metric, hr, ix.n_words * 64u, na, ix.metric == HB_BQ_COSINE ? __ldg(&ix.hdr[b[r]]) : 0.0f);
        }
        return;
    }
    float acc[PR_GROUP] = {};
    for (uint32_t w = lane; w < words; w += 32) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(ra) + w);
        float4 y[PR_GROUP];
#pragma unroll
        for (int r = 0; r < PR_GROUP; ++r) y[r] = __ldg(reinterpret_cast<const float4*>(rb[r]) + w);
#pragma unroll
        for (int r = 0; r < PR_GROUP; ++r) {
            if (ix.metric == HB_COSINE) {
                acc[r] = fmaf(x.x, y[r].x, acc[r]); acc[r] = fmaf(x.y, y[r].y, acc[r]); acc[r] = fmaf(x.z, y[r].z, acc[r]); acc[r] = fmaf(x.w, y[r].w, acc[r]);
            } else if (ix.metric == HB_MANHATTAN) {
                acc[r] += fabsf(x.x - y[r].x) + fabsf(x.y - y[r].y) + fabsf(x.z - y[r].z) + fabsf(x.w - y[r].w);
            } else {
                const float d0 = x.x - y[r].x, d1 = x.y - y[r].y, d2 = x.z - y[r].z, d3 = x.w - y[r].w;
                acc[r] = fmaf(d0, d0, acc[r]); acc[r] = fmaf(d1, d1, acc[r]); acc[r] = fmaf(d2, d2, acc[r]); acc[r] = fmaf(d3, d3, acc[r]);
            }
        }
    }
    __syncwarp();  // lanes leave the strided loop at different trip counts
    const float na = ix.metric == HB_COSINE ? __ldg(&ix.hdr[a]) : 0.0f;
#pragma unroll
    for (int r = 0; r < PR_GROUP; ++r) {
        float v = acc[r];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        out[r] = ix.metric == HB_COSINE ? finish_f32(HB_COSINE, v, na, __ldg(&ix.hdr[b[r]])) : v;
    }
}

// robust_prune (hnsw.rs:567-597): candidates in ascending (distance bits, slot) order in keys[0..n_c); selects at most
// `cap`; a candidate is dropped as soon as one selected point is closer to it (times alpha) than the query is.
// sel_slot / sel_dist: per-warp shared memory, LIST_MAX entries.  Returns the number selected.
__device__ int robust_prune_warp(const DevIndex& ix, const unsigned long long* keys, int n_c, int cap, float alpha, uint32_t* sel_slot,
                                 float* sel_dist) {
    int n_sel = 0;
    __syncwarp();
    for (int i = 0; i < n_c && n_sel < cap; ++i) {
        const unsigned long long k = keys[i];
        const uint32_t cslot = (uint32_t)k, dq_bits = (uint32_t)(k >> 32);
        bool ok = true;
        for (int j0 = 0; j0 < n_sel && ok; j0 += PR_GROUP) {
            const int g = min(PR_GROUP, n_sel - j0);
            float d[PR_GROUP];
            warp_pair_distance_group(ix, cslot, sel_slot + j0, g, d);
#pragma unroll
            for (int r = 0; r < PR_GROUP; ++r)
                if (r < g && __float_as_uint(d[r] * alpha) < dq_bits) ok = false;   // OrderedFloat(d * alpha) < dist_to_query
        }
        if (ok) {
            if (lane_id() == 0) { sel_slot[n_sel] = cslot; sel_dist[n_sel] = __uint_as_float(dq_bits); }
            __syncwarp();
            ++n_sel;
        }
    }
    return n_sel;
}

// Loads / stores of the mutable graph are strong (gpu scope) and carry a "memory" clobber, so neither the compiler nor
// the L1 serves them from an earlier state.
__device__ __forceinline__ uint32_t ld_cg(const uint32_t* p) { uint32_t v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ float ld_cg(const float* p) { float v; asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_cg(uint32_t* p, uint32_t v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_cg(float* p, float v) { asm volatile("st.relaxed.gpu.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }

// add_link (hnsw.rs:524-562) for node p on `level`; exactly one warp works on p at a time.  scratch: per-warp shared memory.
__device__ void add_link_node(const PruneLinkParams& P, uint32_t p, uint32_t x, float dist, unsigned long long* skeys, uint32_t* s_slot,
                                float* s_dist) {
    if (p == x) return;
    const int lane = lane_id();
    const uint32_t lvl = P.level, stride = P.g.stride(lvl), cap = P.g.cap(lvl);
    uint32_t* nb = P.g.nbr[lvl] + (size_t)p * stride;
    float* nd = P.g.dist[lvl] + (size_t)p * stride;
    const uint32_t deg = ld_cg(&P.g.deg[lvl][p]);
    if (deg > cap || p >= P.ix.n || x >= P.ix.n) { build_fail(2, p, x, deg, lvl); return; }
    if (deg < cap) {
        if (lane == 0) { st_cg(&nb[deg], x); st_cg(&nd[deg], dist); st_cg(&P.g.deg[lvl][p], deg + 1); }
        return;
    }
    // full: robust_prune(links) replaces the list, the new link is not part of it.  The <= LIST_MAX links are sorted by
    // (distance bits, slot) through shared memory: unsorted copy in skeys[LIST_MAX..), rank by counting, scatter to skeys[0..).
    __syncwarp();
    unsigned long long key[2] = {~0ull, ~0ull};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int e = lane + 32 * k;
        if (e < (int)deg) key[k] = ((unsigned long long)__float_as_uint(ld_cg(&nd[e])) << 32) | ld_cg(&nb[e]);
        skeys[LIST_MAX + e] = key[k];
    }
    __syncwarp();
    int rank[2] = {0, 0};
    for (int j = 0; j < (int)deg; ++j) {  // a list may hold a link twice (hnsw.rs:521-522 TODO): ties go by position
        const unsigned long long kj = skeys[LIST_MAX + j];
#pragma unroll
        for (int k = 0; k < 2; ++k) rank[k] += (kj < key[k]) || (kj == key[k] && j < lane + 32 * k);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k)
        if (lane + 32 * k < (int)deg) skeys[rank[k]] = key[k];
    __syncwarp();
    const int n_sel = robust_prune_warp(P.ix, skeys, (int)deg, (int)cap, P.alpha, s_slot, s_dist);
    __syncwarp();
    for (int e = lane; e < (int)stride; e += 32) {
        st_cg(&nb[e], e < n_sel ? s_slot[e] : 0xffffffffu);
        st_cg(&nd[e], e < n_sel ? s_dist[e] : 0.0f);
    }
    if (lane == 0) st_cg(&P.g.deg[lvl][p], (uint32_t)n_sel);
    __syncwarp();
}

constexpr int PL_WARPS = 4;

__global__ void __launch_bounds__(PL_WARPS * 32) prune_link_kernel(const PruneLinkParams P) {
    __shared__ uint32_t sh_slot[PL_WARPS][2][LIST_MAX];
    __shared__ float sh_dist[PL_WARPS][2][LIST_MAX];
    __shared__ unsigned long long sh_keys[PL_WARPS][2 * LIST_MAX];  // [0, LIST_MAX) sorted, [LIST_MAX, ..) unsorted
    const int wib = threadIdx.x >> 5, lane = lane_id();
    for (uint32_t i = blockIdx.x * PL_WARPS + wib; i < P.n_items; i += gridDim.x * PL_WARPS) {
    const uint32_t q = P.items[i];
    if (q >= P.ix.n || P.cand_len[i] > P.efc) { build_fail(3, q, P.cand_len[i], i, P.level); continue; }
    uint32_t* sel_slot = sh_slot[wib][0];
    float* sel_dist = sh_dist[wib][0];
    // robust_prune(neighbours, level_of_q, alpha) — hnsw.rs:318
    const int n_c = (int)P.cand_len[i];
    const int n_sel = robust_prune_warp(P.ix, P.cand + (size_t)i * P.efc, n_c, (int)P.own_cap, P.alpha, sel_slot, sel_dist);
    __syncwarp();
    // eps.push(n), hnsw.rs:323 — read by the walk one layer down, where at most M <= 32 were selected (on layer 0 nobody reads it)
    P.sel_out[(size_t)i * 32 + lane] = lane < n_sel ? sel_slot[lane] : 0xffffffffu;
    // add_link(query, (dist, n)) for every selected n — hnsw.rs:320.  q is not linked yet: its list is this warp's alone.
    for (int j = 0; j < n_sel; ++j) add_link_node(P, q, sel_slot[j], sel_dist[j], sh_keys[wib], sh_slot[wib][1], sh_dist[wib][1]);
    // add_link(n, (dist, query)) — hnsw.rs:321 — is posted to n's inbox
    for (int e = lane; e < n_sel; e += 32) {
        const uint32_t t = sel_slot[e];
        if (t != q) {
            if (t >= P.ix.n) build_fail(4, q, t, (unsigned)e, (unsigned)n_sel);
            else {
                const uint32_t pos = atomicAdd(&P.g.inbox_cnt[t], 1u);
                if (pos < INBOX_K) P.g.inbox[(size_t)t * INBOX_K + pos] = ((unsigned long long)__float_as_uint(sel_dist[e]) << 32) | q;
                else atomicAdd(P.g.n_dropped, 1ull);
                if (pos == 0) P.g.touched[atomicAdd(P.g.n_touched, 1u)] = t;
            }
        }
    }
    __syncwarp();
    }
}

// The reverse half of the batch's links: one warp per target that received any, sources applied in ascending
// (distance bits, slot) order.
__global__ void __launch_bounds__(PL_WARPS * 32) apply_reverse_kernel(const PruneLinkParams P) {
    __shared__ uint32_t sh_slot[PL_WARPS][LIST_MAX];
    __shared__ float sh_dist[PL_WARPS][LIST_MAX];
    __shared__ unsigned long long sh_keys[PL_WARPS][2 * LIST_MAX];
    __shared__ unsigned long long sh_in[PL_WARPS][INBOX_K];
    __shared__ unsigned long long sh_raw[PL_WARPS][INBOX_K];
    const int wib = threadIdx.x >> 5, lane = lane_id();
    const uint32_t n_t = *P.g.n_touched;
    for (uint32_t i = blockIdx.x * PL_WARPS + wib; i < n_t; i += gridDim.x * PL_WARPS) {
        const uint32_t t = P.g.touched[i];
        const uint32_t cnt = min(P.g.inbox_cnt[t], INBOX_K);
        unsigned long long rec = ~0ull;
        if (lane < (int)cnt) rec = P.g.inbox[(size_t)t * INBOX_K + lane];
        sh_raw[wib][lane] = rec;
        __syncwarp();
        int rank = 0;
        for (uint32_t j = 0; j < cnt; ++j) {
            const unsigned long long rj = sh_raw[wib][j];
            rank += (rj < rec) || (rj == rec && (int)j < lane);
        }
        if (lane < (int)cnt) sh_in[wib][rank] = rec;
        __syncwarp();
        for (uint32_t j = 0; j < cnt; ++j) {
            const unsigned long long r = sh_in[wib][j];
            add_link_node(P, t, (uint32_t)r, __uint_as_float((uint32_t)(r >> 32)), sh_keys[wib], sh_slot[wib], sh_dist[wib]);
            __syncwarp();
        }
        if (lane == 0) P.g.inbox_cnt[t] = 0;
        __syncwarp();
    }
}

__global__ void reset_touched_kernel(uint32_t* n_touched) { *n_touched = 0; }

__global__ void fill_u32_kernel(uint32_t* p, size_t n, uint32_t v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void fill_stride_offsets_kernel(uint32_t* off, size_t n_plus_1, uint32_t stride) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_plus_1) off[i] = (uint32_t)(i * stride);
}

struct Rng {  // splitmix64 (bit-compatibility with rand's StdRng is not attempted; the distribution is the reference's)
    uint64_t s;
    uint64_t next() {
        uint64_t z = (s += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
    }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

struct Frees {
    std::vector<void*> v;
    ~Frees() { for (void* p : v) cudaFree(p); }
    template <class T> hb_status alloc(T** out, size_t count) {
        void* p = nullptr;
        if (cudaMalloc(&p, std::max<size_t>(count * sizeof(T), 16)) != cudaSuccess) { set_error("device allocation of %zu bytes failed", count * sizeof(T)); cudaGetLastError(); return HB_ENOMEM; }
        v.push_back(p);
        *out = (T*)p;
        return HB_OK;
    }
};

}  // namespace

// HnswBuilder::build for every item of `ix` (hnsw.rs:122-216): fills ix->layers / eps / max_level on the host.
hb_status build_graph_on_device(hb_index* ix, uint32_t M, uint32_t M0, uint32_t efc, float alpha, uint64_t seed, uint32_t batch_max, int device,
                                uint64_t* stats /* [8]: batches, launches, items, max_level, reverse links dropped (inbox full), walks cut short */) {
    const auto t_start = std::chrono::steady_clock::now();
    auto ms_since = [](std::chrono::steady_clock::time_point t0) { return (uint64_t)std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count(); };
    const size_t n = ix->ids.size();
    if (M < 2 || M > 32 || M0 < 2 || M0 > FIXED_DEG_MAX || efc < 1 || efc > 4096) { set_error("hb_index_build_graph: need 2 <= M <= 32, 2 <= M0 <= 64, 1 <= ef_construction <= 4096"); return HB_EINVAL; }
    const uint32_t stride0 = M0 <= FIXED_DEG ? FIXED_DEG : FIXED_DEG_MAX;
    if (n >= 0xfffffff0ull / stride0) { set_error("too many items"); return HB_EINVAL; }
    ix->layers.clear();
    ix->eps.clear();
    ix->max_level = 0;
    if (n == 0) return HB_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device available (libhannoy_b200 has no CPU path)"); return HB_ECUDA; }
    if (device < 0 || device >= ndev) { set_error("bad device %d", device); return HB_EINVAL; }
    CUDA_TRYB(cudaSetDevice(device));

    // ---- levels: x ~ exp(1/ln M) quantised (hnsw.rs:94-120), one draw per item in id order (hnsw.rs:143-151) ----
    std::vector<double> cum;
    {
        const float level_factor = 1.0f / std::log((float)M + 1.1920929e-07f);
        double c = 0;
        for (int level = 0;; ++level) {
            float proba = std::exp((float)level * (-1.0f / level_factor)) * (1.0f - std::exp(-1.0f / level_factor));
            if (proba < 1e-09f) break;
            c += proba;
            cum.push_back(c);
        }
        for (double& v : cum) v /= c;
    }
    Rng rng{seed};
    std::vector<uint32_t> level(n);
    uint32_t L = 0;
    for (size_t i = 0; i < n; ++i) {
        double u = rng.uniform();
        uint32_t l = (uint32_t)(std::lower_bound(cum.begin(), cum.end(), u) - cum.begin());
        if (l >= cum.size()) l = (uint32_t)cum.size() - 1;
        level[i] = l;
        L = std::max(L, l);
    }
    if (L >= (uint32_t)MAX_LEVELS) { set_error("too many layers"); return HB_EINVAL; }
    // insertion order: level groups from the top down, ids ascending inside a group (levels.sort_unstable_by(b.cmp(a)), hnsw.rs:244)
    std::vector<uint32_t> order(n);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return level[a] > level[b]; });
    std::vector<uint32_t> eps;  // every item of the top level (hnsw.rs:268-279)
    for (uint32_t s : order) { if (level[s] != L) break; eps.push_back(s); }

    // ---- device state ----
    Frees fr;
    DevIndex d;
    hb_status st;
    if ((st = setup_dev_rows(ix, d, fr.v)) != HB_OK) return st;
    float* d_hdr = nullptr;
    if ((st = fr.alloc(&d_hdr, n)) != HB_OK) return st;
    CUDA_TRYB(cudaMemcpy(d_hdr, ix->host_hdr.data(), n * 4, cudaMemcpyHostToDevice));
    d.hdr = d_hdr;
    d.n_layers = L + 1;
    d.max_level = L;
    GraphDev g;
    g.n = (uint32_t)n; g.M = M; g.M0 = M0; g.stride0 = stride0;
    for (uint32_t l = 0; l <= L; ++l) {
        const size_t cells = n * (size_t)g.stride(l);
        if ((st = fr.alloc(&g.nbr[l], cells)) != HB_OK || (st = fr.alloc(&g.dist[l], cells)) != HB_OK || (st = fr.alloc(&g.deg[l], n)) != HB_OK)
            return st;
        fill_u32_kernel<<<(unsigned)((cells + 255) / 256), 256>>>(g.nbr[l], cells, 0xffffffffu);
        CUDA_TRYB(cudaMemset(g.dist[l], 0, cells * 4));
        CUDA_TRYB(cudaMemset(g.deg[l], 0, n * 4));
        g_launches += 1;
    }
    if ((st = fr.alloc(&g.inbox_cnt, n)) != HB_OK || (st = fr.alloc(&g.inbox, n * (size_t)INBOX_K)) != HB_OK ||
        (st = fr.alloc(&g.touched, n)) != HB_OK || (st = fr.alloc(&g.n_touched, 4)) != HB_OK || (st = fr.alloc(&g.n_dropped, 2)) != HB_OK)
        return st;
    CUDA_TRYB(cudaMemset(g.inbox_cnt, 0, n * 4));
    CUDA_TRYB(cudaMemset(g.n_touched, 0, 16));
    CUDA_TRYB(cudaMemset(g.n_dropped, 0, 16));
    uint32_t* d_off = nullptr;  // fixed-stride "CSR" offsets shared by the upper layers: list s = [s * M, (s + 1) * M)
    if ((st = fr.alloc(&d_off, n + 1)) != HB_OK) return st;
    fill_stride_offsets_kernel<<<(unsigned)((n + 1 + 255) / 256), 256>>>(d_off, n + 1, M);
    g_launches += 1;
    d.nbr0x = g.nbr[0];
    d.nbr0_stride = stride0;
    for (uint32_t l = 1; l <= L; ++l) { d.off[l] = d_off; d.nbr[l] = g.nbr[l]; }
    uint32_t* d_eps = nullptr;
    if ((st = fr.alloc(&d_eps, eps.size())) != HB_OK) return st;
    CUDA_TRYB(cudaMemcpy(d_eps, eps.data(), eps.size() * 4, cudaMemcpyHostToDevice));
    d.eps = d_eps;
    d.n_ep = (uint32_t)eps.size();

    // ---- search parameters (same ring / heap geometry rules as the reader's launches) ----
    BuildSearchParams bp;
    SearchParams& sp = bp.sp;
    sp.ix = d;
    sp.mode = 1;  // the query is a stored item
    sp.vis_atomic = 1;  // a mutable list can name a neighbour twice (collapsed when the lists are written back)
    sp.count = efc; sp.ef_raw = efc;
    sp.q_smem_bytes = (d.row_stride + 15) & ~15u;
    sp.defer = tunable("defer", 1);
    if (d.kind == KIND_F32_WARP) {
        uint32_t budget = (uint32_t)std::max(0, tunable("ring_bytes", 12288));
        uint32_t slots = budget / d.row_stride / ROW_GROUP * ROW_GROUP;
        slots = std::max<uint32_t>(ROW_GROUP, std::min<uint32_t>(slots, 32));
        sp.ring_slots = slots;
        sp.ring_stride = d.row_stride;
    }
    const uint32_t ef0 = std::max<uint32_t>(efc, d.n_ep);
    sp.res_cap = (ef0 + 32 + 31) & ~31u;
    sp.q_cap = (ef0 + 64 + 31) & ~31u;
    if (ix->metric == HB_BQ_COSINE) {
        // the only metric whose distances can come out negative (no clamp, binary_quantized_cosine.rs:49-58 — e.g. d(x, x)
        // when sqrt(L)^2 rounds below L): the reader re-runs such a query in its global-memory pass; here the walk simply
        // runs without dead-entry trimming and without deferred pops, in a larger queue
        sp.no_trim = 1;
        sp.q_cap = (8 * ef0 + 64 + 31) & ~31u;
    }
    sp.pass = 0;
    int bps = build_search_blocks_per_sm(sp);
    if (bps == 0 && sp.ring_slots) { sp.ring_slots = 0; sp.ring_stride = 0; bps = build_search_blocks_per_sm(sp); }  // rows too long to stage
    if (bps == 0) { set_error("hb_index_build_graph: ef_construction / dimensions too large for the shared-memory heaps"); return HB_EINVAL; }
    cudaDeviceProp prop;
    CUDA_TRYB(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * bps;
    const int n_slots = blocks * SEARCH_WARPS_PER_BLOCK;
    sp.vis_words = (uint32_t)(((n + 31) / 32 + 31) / 32 * 32);
    sp.touched_cap = (uint32_t)std::max(1024, tunable("touched_cap", 16384));
    uint32_t *d_vis = nullptr, *d_touched = nullptr;
    unsigned long long* d_wc = nullptr;
    if ((st = fr.alloc(&d_vis, (size_t)n_slots * sp.vis_words)) != HB_OK || (st = fr.alloc(&d_touched, (size_t)n_slots * sp.touched_cap)) != HB_OK ||
        (st = fr.alloc(&d_wc, 2)) != HB_OK)
        return st;
    CUDA_TRYB(cudaMemset(d_vis, 0, (size_t)n_slots * sp.vis_words * 4));
    sp.visited = d_vis; sp.touched = d_touched; sp.work_counter = d_wc;

    // ---- batches ----
    if (batch_max == 0) batch_max = 4096;
    uint32_t* d_items = nullptr;
    uint32_t *d_sel[2] = {nullptr, nullptr}, *d_cand_len = nullptr;
    unsigned long long* d_cand = nullptr;
    if ((st = fr.alloc(&d_items, n)) != HB_OK || (st = fr.alloc(&d_sel[0], (size_t)batch_max * 32)) != HB_OK ||
        (st = fr.alloc(&d_sel[1], (size_t)batch_max * 32)) != HB_OK || (st = fr.alloc(&d_cand_len, batch_max)) != HB_OK ||
        (st = fr.alloc(&d_cand, (size_t)batch_max * efc)) != HB_OK)
        return st;
    CUDA_TRYB(cudaMemcpy(d_items, order.data(), n * 4, cudaMemcpyHostToDevice));
    cudaStream_t stream = nullptr;
    {
        unsigned int zero[8] = {};
        CUDA_TRYB(cudaMemcpyToSymbol(g_build_err, zero, sizeof(zero)));
    }
    uint64_t n_batches = 0, n_launch = 0;
    const auto t_loop = std::chrono::steady_clock::now();
    const size_t inflight_div = (size_t)std::max(1, tunable("build_inflight_div", 64));
    const bool sync_each = tunable("build_sync", 0) != 0;
    const int link_blocks = tunable("build_link_blocks", 0);  // debugging aid: cap the prune/link grid  // debugging aid: a failure is reported with the launch that caused it
    size_t done = 0;
    while (done < n) {
        const uint32_t grp = level[order[done]];
        size_t grp_end = done;
        while (grp_end < n && level[order[grp_end]] == grp) ++grp_end;
        while (done < grp_end) {
            // items in flight never see each other: keep them a small fraction of what is already linked
            size_t b = std::min<size_t>({(size_t)batch_max, grp_end - done, std::max<size_t>(1, done / inflight_div)});
            PruneLinkParams pl;
            pl.ix = d; pl.g = g; pl.items = d_items + done; pl.n_items = (uint32_t)b; pl.alpha = alpha;
            pl.own_cap = grp == 0 ? M0 : M;
            pl.cand = d_cand; pl.cand_len = d_cand_len; pl.efc = efc;
            bp.n_items = (uint32_t)b;
            sp.q_slots = d_items + done;
            sp.nq = b; sp.n_work = (uint32_t)b;
            bp.efc = efc; bp.cand = d_cand; bp.cand_len = d_cand_len; bp.eps_stride = 32;
            bp.n_cut = g.n_dropped + 1;
            int cur = 0;
            for (int lvl = (int)grp; lvl >= 0; --lvl) {   // hnsw.rs:313-326
                bp.level = (uint32_t)lvl;
                bp.descend = lvl == (int)grp;
                bp.eps_in = d_sel[cur ^ 1];
                if ((st = launch_build_search(bp, blocks, stream)) != HB_OK) return st;
                if (sync_each) {
                    cudaError_t e = cudaStreamSynchronize(stream);
                    if (e != cudaSuccess) { set_error("build search failed (group %u, layer %d, batch %llu of %zu items at %zu): %s", grp, lvl, (unsigned long long)n_batches, b, done, cudaGetErrorString(e)); return HB_ECUDA; }
                }
                pl.level = (uint32_t)lvl;
                pl.sel_out = d_sel[cur];
                {
                    unsigned pl_blocks = (unsigned)((b + PL_WARPS - 1) / PL_WARPS);
                    if (link_blocks > 0) pl_blocks = std::min<unsigned>(pl_blocks, (unsigned)link_blocks);
                    prune_link_kernel<<<pl_blocks, PL_WARPS * 32, 0, stream>>>(pl);
                    // at most b * 32 targets; the kernel loops over *n_touched
                    unsigned ap_blocks = (unsigned)std::min<size_t>((b * 32 + PL_WARPS - 1) / PL_WARPS, (size_t)prop.multiProcessorCount * 16);
                    apply_reverse_kernel<<<ap_blocks, PL_WARPS * 32, 0, stream>>>(pl);
                    reset_touched_kernel<<<1, 1, 0, stream>>>(g.n_touched);
                    g_launches += 2;
                    n_launch += 2;
                }
                if (sync_each) {
                    cudaError_t e = cudaStreamSynchronize(stream);
                    if (e != cudaSuccess) { set_error("prune/link failed (group %u, layer %d, batch %llu of %zu items at %zu): %s", grp, lvl, (unsigned long long)n_batches, b, done, cudaGetErrorString(e)); return HB_ECUDA; }
                }
                g_launches += 1;
                n_launch += 2;
                cur ^= 1;
            }
            done += b;
            ++n_batches;
        }
        cudaError_t e = cudaStreamSynchronize(stream);   // one sync per level group keeps failures close to their cause
        if (e != cudaSuccess) { set_error("graph build failed on level group %u: %s", grp, cudaGetErrorString(e)); return HB_ECUDA; }
        unsigned int err[8] = {};
        CUDA_TRYB(cudaMemcpyFromSymbol(err, g_build_err, sizeof(err)));
        if (err[0]) {
            set_error("graph build: internal inconsistency %u on level group %u (%u, %u, %u, %u)", err[0], grp, err[1], err[2], err[3], err[4]);
            return HB_ECUDA;
        }
    }
    CUDA_TRYB(cudaDeviceSynchronize());
    const uint64_t loop_ms = ms_since(t_loop);

    // ---- back to the host: per-layer CSR over slots, neighbours ascending (what Links / roaring iteration give the reader) ----
    ix->layers.assign(L + 1, HostLayer());
    std::vector<uint32_t> h_nbr, h_deg;
    for (uint32_t l = 0; l <= L; ++l) {
        const uint32_t stride = g.stride(l);
        h_nbr.resize(n * (size_t)stride);
        h_deg.resize(n);
        CUDA_TRYB(cudaMemcpy(h_nbr.data(), g.nbr[l], h_nbr.size() * 4, cudaMemcpyDeviceToHost));
        CUDA_TRYB(cudaMemcpy(h_deg.data(), g.deg[l], n * 4, cudaMemcpyDeviceToHost));
        HostLayer& hl = ix->layers[l];
        hl.off.assign(n + 1, 0);
        for (size_t s = 0; s < n; ++s) {  // RoaringBitmap::from_iter(links) (hnsw.rs:203-207): ascending, duplicates collapse
            uint32_t dg = level[s] >= l ? std::min(h_deg[s], stride) : 0;
            uint32_t* row = h_nbr.data() + s * stride;
            std::sort(row, row + dg);
            dg = (uint32_t)(std::unique(row, row + dg) - row);
            h_deg[s] = dg;
            hl.off[s + 1] = hl.off[s] + dg;
        }
        hl.nbr.resize(hl.off[n]);
        for (size_t s = 0; s < n; ++s) std::copy(h_nbr.begin() + s * stride, h_nbr.begin() + s * stride + h_deg[s], hl.nbr.begin() + hl.off[s]);
    }
    ix->eps = eps;
    ix->max_level = L;
    ix->node_level = level;
    ix->have_metadata = true;
    ix->meta_distance = hb_metric_name(ix->metric);
    ix->meta_dims = ix->dims;
    if (stats) {
        unsigned long long h[2] = {};
        CUDA_TRYB(cudaMemcpy(h, g.n_dropped, 16, cudaMemcpyDeviceToHost));
        stats[0] = n_batches; stats[1] = n_launch; stats[2] = n; stats[3] = L; stats[4] = h[0]; stats[5] = h[1];
        stats[6] = loop_ms; stats[7] = ms_since(t_start);
    }
    return HB_OK;
}

}  // namespace hb
