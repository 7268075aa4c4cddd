// search.cu — the hot path: Visitor::visit / hnsw_search / nns_by_item / brute_force_search of
// src/reader.rs rewritten as one persistent sm_100a kernel, one warp per query.
//
//   * candidate queue and result set live in shared memory as sorted arrays of 64-bit keys
//     (distance bits << 32 | slot): ascending u64 order == the reference's (OrderedFloat, ItemId)
//     tuple order (src/ordered_float.rs:25-29), slots being order-isomorphic to ItemIds;
//   * the visited set (`path: RoaringBitmap`, reader.rs:734) is an exact per-warp bitset in global
//     memory, cleared through a touched list;
//   * f32 neighbour rows are gathered by the bulk async-copy engine (one `cp.async.bulk` per row, posted by
//     the lane that owns the neighbour, landing in a per-warp shared-memory ring behind mbarriers, ring.cuh)
//     so a warp keeps ring_slots x row_bytes in flight without registers; binary codes are read with
//     128-bit loads, one row per lane.  Distances are reduced with warp shuffles in the AVX lane order so
//     they are bit-identical to the reference (dist.cuh);
//   * layer-0 adjacency is read from a fixed-stride copy (one aligned 128-byte line per expansion) and the
//     line of the most likely next candidate is requested while the current rows are in flight;
//   * queries whose heaps outgrow shared memory are re-run by a second pass of the same code with
//     heaps in global memory (never on the CPU).
#include <cfloat>
#include <cstdio>

#include "dist.cuh"
#include "ring.cuh"
#include "sorted.cuh"

namespace hb {

// Optional phase timers (build with -DHB_PHASES; dev only): cycles per phase summed over all queries.
__device__ unsigned long long g_phase[16];
#ifdef HB_PHASES
#define PH_DECL long long ph_t = clock64();
#define PH_ADD(c, i) { long long ph_n = clock64(); (c).ph[i] += ph_n - ph_t; ph_t = ph_n; }
#define PH_RESET ph_t = clock64();
#else
#define PH_DECL
#define PH_ADD(c, i) {}
#define PH_RESET
#endif
enum { PH_STAGE = 0, PH_UPPER = 1, PH_ADJ = 2, PH_VIS = 3, PH_ROWS = 4, PH_HEAP = 5, PH_TAIL = 6, PH_TOTAL = 7, PH_N = 8 };

struct Ctx {
    const SearchParams& p;
    u64* res; int res_len; int res_cap;   // ascending (bits<<32 | slot)
    u64* que; int q_len; int q_cap;       // descending (bits<<32 | ~slot): next to pop is the LAST element
    uint32_t* vis; uint32_t* touched; uint32_t touched_len; bool touched_over;
    const float* qs; float qn;            // query (device layout) in shared memory, query header norm
    uint32_t excl;                        // by_item: slot removed from the candidates, else UINT32_MAX
    bool overflow;
    RowRing& ring;                        // per-warp, lives across queries (barrier phases persist)
#ifdef HB_PHASES
    long long ph[PH_N];
#endif
    uint32_t cur_dist, cur_exp, cur_deg;  // counters of the visit in progress
    u64 n_dist_up, n_exp_up, n_deg_up, n_dist_l0, n_exp_l0, n_deg_l0;
    __device__ Ctx(const SearchParams& pp, RowRing& rr) : p(pp), ring(rr) {}
};

__device__ __forceinline__ float key_dist(u64 k) { return __uint_as_float((uint32_t)(k >> 32)); }

// res.push (unconditional)
__device__ __forceinline__ void res_push(Ctx& c, u64 key) {
    if (c.res_len >= c.res_cap) { c.overflow = true; return; }
    int pos = count_lt(c.res, c.res_len, key);
    insert_at(c.res, c.res_len, pos, key);
    c.res_len++;
}
// `if res.len() == ef { push_pop_max } else { push }` — reader.rs:360-364
__device__ __forceinline__ void res_accept(Ctx& c, u64 key, int ef) {
    if (c.res_len == ef) {
        if (ef == 0) return;                      // push_pop_max on an empty heap returns the item
        if (key > c.res[c.res_len - 1]) return;   // pushed and popped straight away
        int pos = count_lt(c.res, c.res_len - 1, key);
        insert_at(c.res, c.res_len - 1, pos, key);
    } else {
        res_push(c, key);
    }
}
// An entry can never be popped again once res is full and its distance exceeds the current f_max:
// f_max never grows while res.len() >= ef, and the loop breaks at the first `f > f_max`
// (reader.rs:333-336).  Only argued for non-negative distances (bit order == numeric order).
//
// Pruned entries are also the reference's "break sentinels": popping one ends the walk before any entry
// that follows it in BIT order.  For non-negative distances every follower would end the walk itself, so
// dropping the sentinel changes nothing; a negative distance (only BinaryQuantizedCosine can produce one,
// binary_quantized_cosine.rs:49-58 has no clamp) sorts last by bits yet passes `f > f_max`, so the first
// negative distance seen in the pruning pass sends the query to pass 1, which never prunes.
__device__ __forceinline__ bool is_dead(const Ctx& c, uint32_t bits, int ef) {
    if (c.p.pass != 0) return false;
    if (c.res_len < ef || c.res_len == 0) return false;
    uint32_t mb = (uint32_t)(c.res[c.res_len - 1] >> 32);
    if ((bits | mb) & 0x80000000u) return false;
    return __uint_as_float(bits) > __uint_as_float(mb);
}
// search_queue.push — reader.rs:319,354
__device__ __forceinline__ void queue_push(Ctx& c, uint32_t bits, uint32_t slot, int ef) {
    if (is_dead(c, bits, ef)) return;
    u64 qk = ((u64)bits << 32) | (uint32_t)(~slot);
    if (c.q_len == c.q_cap) {
        u64 worst = c.que[0];
        if (qk > worst) { c.overflow = true; return; }  // would have to drop a live entry
        if (!is_dead(c, (uint32_t)(worst >> 32), ef)) c.overflow = true;
        int pos = count_gt(c.que, c.q_len, qk);
        insert_drop_front(c.que, pos, qk);
    } else {
        int pos = count_gt(c.que, c.q_len, qk);
        insert_at(c.que, c.q_len, pos, qk);
        c.q_len++;
    }
}

// ---- visited set --------------------------------------------------------------------------------------
__device__ __forceinline__ bool vis_test_and_set(Ctx& c, uint32_t s, bool valid) {
    bool fresh = false;
    if (valid) {
        uint32_t bit = 1u << (s & 31);
        uint32_t old = atomicOr(&c.vis[s >> 5], bit);
        fresh = !(old & bit);
    }
    unsigned m = __ballot_sync(FULL, fresh);
    if (m) {
        uint32_t r = __popc(m & ((1u << lane_id()) - 1));
        uint32_t at = c.touched_len + r;
        if (fresh) {
            if (at < c.p.touched_cap) c.touched[at] = s;
        }
        c.touched_len += __popc(m);
        if (c.touched_len > c.p.touched_cap) c.touched_over = true;
    }
    return fresh;
}
// path.clear() — reader.rs:743
__device__ __forceinline__ void vis_clear(Ctx& c) {
    __syncwarp();
    if (c.touched_over) {
        for (uint32_t w = lane_id(); w < c.p.vis_words; w += 32) c.vis[w] = 0;
    } else {
        for (uint32_t i = lane_id(); i < c.touched_len; i += 32) c.vis[c.touched[i] >> 5] = 0;
    }
    c.touched_len = 0;
    c.touched_over = false;
    __syncwarp();
    __threadfence_block();
}

__device__ __forceinline__ bool passes_filter(const Ctx& c, uint32_t s, bool filt) {
    if (!filt) return true;
    if (s == c.excl) return false;
    if (c.p.cand_bits) return (__ldg(&c.p.cand_bits[s >> 5]) >> (s & 31)) & 1;
    return true;
}

// ---- distances of one chunk (<= 32 rows, one per lane) -----------------------------------------------------
template <int KIND>
__device__ __forceinline__ float chunk_distances(Ctx& c, unsigned mask, uint32_t s) {
    const DevIndex& ix = c.p.ix;
    const int lane = lane_id();
    float mine = 0.0f;
    if (KIND == KIND_F32_WARP && c.ring.slots == 0) {
        // rows too long for the shared-memory ring: 4 rows at a time straight from global memory
        unsigned m = mask;
        while (m) {
            int l[4];
            uint32_t sl[4];
            const uint8_t* rowp[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (m) { l[r] = __ffs(m) - 1; m &= m - 1; } else l[r] = -1;
                sl[r] = __shfl_sync(FULL, s, l[r] < 0 ? l[0] : l[r]);
                rowp[r] = ix.rows + (size_t)sl[r] * ix.row_stride;
            }
            float raw[4];
            if (ix.metric == HB_COSINE) warp_rows_raw<4, true, false>(ix, c.qs, rowp, raw);
            else warp_rows_raw<4, false, false>(ix, c.qs, rowp, raw);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (lane == l[r]) {
                    float in = (ix.metric == HB_COSINE) ? __ldg(&ix.hdr[sl[r]]) : 0.0f;
                    mine = finish_f32(ix.metric, raw[r], c.qn, in);
                }
            }
        }
    } else if (KIND == KIND_F32_WARP) {
        // Row r (in ascending-lane order) lands in ring slot r % S.  The first S copies are posted at once by
        // the lanes that own them; slot group g is re-posted as soon as its ROW_GROUP rows were consumed.
        const int S = (int)c.ring.slots;
        const int n_live = __popc(mask);
        const bool has = (mask >> lane) & 1;
        const int rank = __popc(mask & ((1u << lane) - 1));
        const uint8_t* grow = ix.rows + (size_t)s * ix.row_stride;
        if (has && rank < S) c.ring.post(rank, grow, ix.row_stride);
        const float in = (has && ix.metric == HB_COSINE) ? __ldg(&ix.hdr[s]) : 0.0f;
        int slot0 = 0;
        for (int r0 = 0; r0 < n_live; r0 += ROW_GROUP) {
            const int g = min(ROW_GROUP, n_live - r0);
            const uint8_t* rowp[ROW_GROUP];
#pragma unroll
            for (int r = 0; r < ROW_GROUP; ++r) {
                if (r < g) c.ring.wait(slot0 + r);
                rowp[r] = c.ring.ptr + (size_t)(slot0 + (r < g ? r : 0)) * c.ring.stride;
            }
            float raw[ROW_GROUP];
            if (ix.metric == HB_COSINE) warp_rows_raw<ROW_GROUP, true, true>(ix, c.qs, rowp, raw);
            else warp_rows_raw<ROW_GROUP, false, true>(ix, c.qs, rowp, raw);
            __syncwarp();  // every lane has read the group's slots: they may be overwritten
            const int nxt = rank - r0 - S;
            if (has && nxt >= 0 && nxt < g) c.ring.post(slot0 + nxt, grow, ix.row_stride);
#pragma unroll
            for (int r = 0; r < ROW_GROUP; ++r)
                if (has && rank == r0 + r) mine = finish_f32(ix.metric, raw[r], c.qn, in);
            slot0 += ROW_GROUP;
            if (slot0 >= S) slot0 = 0;
        }
    } else if (KIND == KIND_F32_LANE) {
        if (mask >> lane & 1) {
            const float* row = reinterpret_cast<const float*>(ix.rows + (size_t)s * ix.row_stride);
            float in = (ix.metric == HB_COSINE) ? __ldg(&ix.hdr[s]) : 0.0f;
            mine = lane_distance_f32<true>(ix, c.qs, c.qn, row, in);
        }
    } else {
        if (mask >> lane & 1) {
            const uint64_t* row = reinterpret_cast<const uint64_t*>(ix.rows + (size_t)s * ix.row_stride);
            uint32_t h = lane_xor_popc(reinterpret_cast<const uint64_t*>(c.qs), row, ix.n_words);
            float in = (ix.metric == HB_BQ_COSINE) ? __ldg(&ix.hdr[s]) : 0.0f;
            mine = finish_bin(ix.metric, h, ix.n_words * 64u, c.qn, in);
        }
    }
    return mine;
}

enum ChunkMode { CH_EP, CH_NBR, CH_LINEAR };
constexpr int MERGE_TILES = 8;  // heaps of up to 256 entries are updated by one merge pass per chunk

// One chunk of <= 32 points (lane i holds the i-th, ascending).  Mirrors, for all 32 at once, the body of
// `for &ep in eps` (reader.rs:315-325), `for point in links.iter()` (reader.rs:342-366) or the
// brute-force loop (reader.rs:683-705).
template <int KIND, int MODE>
__device__ __forceinline__ void process_chunk(Ctx& c, uint32_t s, bool valid, float f_max, int ef, bool filt, int lvl01) {
    const int lane = lane_id();
    bool live;
    PH_DECL
    if (MODE == CH_NBR) live = vis_test_and_set(c, s, valid);           // `if !path.insert(point) { continue }`
    else if (MODE == CH_EP) { vis_test_and_set(c, s, valid); live = valid; }  // path.insert(ep), result ignored
    else live = valid;
    unsigned lm = __ballot_sync(FULL, live);
    if (lvl01) PH_ADD(c, PH_VIS)
    if (!lm) return;
    c.cur_dist += __popc(lm);
    float dist = chunk_distances<KIND>(c, lm, s);
    uint32_t bits = __float_as_uint(dist);
    if (lvl01) PH_ADD(c, PH_ROWS)
    if (MODE != CH_LINEAR && c.p.pass == 0 && __ballot_sync(FULL, live && (bits >> 31))) { c.overflow = true; return; }
    bool pf = live && passes_filter(c, s, filt);
    bool acc;
    if (MODE == CH_NBR) {
        // `res.len() < self.ef || dist < f_max` with live len and stale f_max (reader.rs:353): the first
        // ef - len filter-passing points in ascending order are taken unconditionally.
        unsigned pfm = __ballot_sync(FULL, pf);
        int before = __popc(pfm & ((1u << lane) - 1));
        bool fill = (c.res_len + before) < ef;
        acc = live && (fill || dist < f_max);
    } else {
        acc = live;
    }
    unsigned accm = __ballot_sync(FULL, acc);
    unsigned resm = __ballot_sync(FULL, acc && pf);
    u64 key = ((u64)bits << 32) | s;
    if (c.res_len <= 32 * MERGE_TILES && c.q_len <= 32 * MERGE_TILES) {
        // All accepted points of the chunk enter the heaps in one merge pass each (sorted.cuh merge_batch).
        // Result set: pushing the keys one by one with `if len == ef { push_pop_max } else { push }` leaves the
        // min(ef, len + m) smallest of the union when len <= ef, and the whole union when len > ef (or for entry points).
        const int m_res = __popc(resm);
        int target = (MODE == CH_EP || c.res_len > ef) ? c.res_len + m_res : min(ef, c.res_len + m_res);
        if (target > c.res_cap) { c.overflow = true; return; }
        if (m_res) c.res_len = merge_batch<false, false, MERGE_TILES>(c.res, c.res_len, acc && pf, key, target, 0u, nullptr);
        if (MODE == CH_LINEAR) return;
        // Queue: entries that can never be popped (is_dead, judged against the UPDATED result set) are not pushed,
        // and old ones — they sit at the front of the descending array — are trimmed in the same pass.
        const bool prune = c.p.pass == 0 && c.res_len >= ef && c.res_len > 0 && !(c.res[c.res_len - 1] >> 63);
        const uint32_t mb = prune ? (uint32_t)(c.res[c.res_len - 1] >> 32) : 0xffffffffu;
        const bool qhas = acc && !(prune && !(bits >> 31) && bits > mb);
        const int mq = __popc(__ballot_sync(FULL, qhas));
        if (c.q_len + mq > c.q_cap) {
            int d0 = 0;
            if (prune) {
                for (int i = lane; i < c.q_len; i += 32) d0 += (uint32_t)(c.que[i] >> 32) > mb;
                d0 = __reduce_add_sync(FULL, d0);
            }
            if (c.q_len + mq - d0 > c.q_cap) { c.overflow = true; return; }  // would have to drop a live entry
        }
        if (mq || prune)
            c.q_len = merge_batch<true, true, MERGE_TILES>(c.que, c.q_len, qhas, ((u64)bits << 32) | (uint32_t)(~s), c.q_cap, mb, nullptr);
    } else {
        for (unsigned m = resm; m; m &= m - 1) {
            u64 k = __shfl_sync(FULL, key, __ffs(m) - 1);
            if (MODE == CH_EP) res_push(c, k);       // reader.rs:322-324: unconditional
            else res_accept(c, k, ef);
        }
        if (MODE == CH_LINEAR) return;
        for (unsigned m = accm; m; m &= m - 1) {
            int src = __ffs(m) - 1;
            uint32_t b = __shfl_sync(FULL, bits, src), sl = __shfl_sync(FULL, s, src);
            queue_push(c, b, sl, ef);
        }
    }
    if (lvl01) PH_ADD(c, PH_HEAP)
}

// Visitor::visit — reader.rs:301-369.  Entry points: `eps` (n_eps slots in global memory) or `single`.
template <int KIND>
__device__ __forceinline__ void visit(Ctx& c, const uint32_t* eps, uint32_t n_eps, uint32_t single, uint32_t level, int ef, bool filt) {
    const DevIndex& ix = c.p.ix;
    const int lane = lane_id();
    const int l01 = level ? 0 : 1;
    c.res_len = 0;
    c.q_len = 0;
    c.cur_dist = c.cur_exp = c.cur_deg = 0;
    if (eps) {
        for (uint32_t base = 0; base < n_eps; base += 32) {
            bool valid = base + lane < n_eps;
            uint32_t s = valid ? __ldg(&eps[base + lane]) : 0;
            process_chunk<KIND, CH_EP>(c, s, valid, FLT_MAX, ef, filt, l01);
        }
    } else {
        process_chunk<KIND, CH_EP>(c, single, lane == 0, FLT_MAX, ef, filt, l01);
    }
    const uint32_t* off = ix.off[level];
    const uint32_t* nbr = ix.nbr[level];
    const uint32_t* nbrx = level == 0 ? ix.nbr0x : nullptr;
    uint32_t spec_cs = 0xffffffffu, spec_adj = 0xffffffffu;  // adjacency line requested ahead of its pop
    while (c.q_len > 0 && !c.overflow) {
        u64 top = c.que[c.q_len - 1];
        float f = key_dist(top);
        float f_max = c.res_len ? key_dist(c.res[c.res_len - 1]) : FLT_MAX;
        if (f > f_max) break;
        c.q_len--;
        uint32_t cs = ~(uint32_t)top;
        c.cur_exp += 1;
        if (nbrx) {
            PH_DECL
            uint32_t s = (cs == spec_cs) ? spec_adj : __ldg(&nbrx[(size_t)cs * FIXED_DEG + lane]);
            bool valid = s != 0xffffffffu;
            c.cur_deg += __popc(__ballot_sync(FULL, valid));
            PH_ADD(c, PH_ADJ)
            // the entry now on top of the queue is popped next unless one of cs's neighbours beats it
            if (c.q_len > 0) {
                spec_cs = ~(uint32_t)c.que[c.q_len - 1];
                spec_adj = __ldg(&nbrx[(size_t)spec_cs * FIXED_DEG + lane]);
            } else {
                spec_cs = 0xffffffffu;
            }
            process_chunk<KIND, CH_NBR>(c, valid ? s : 0, valid, f_max, ef, filt, l01);
        } else {
            uint32_t b = __ldg(&off[cs]), e = __ldg(&off[cs + 1]);
            c.cur_deg += e - b;
            for (uint32_t base = b; base < e; base += 32) {
                bool valid = base + lane < e;
                uint32_t s = valid ? __ldg(&nbr[base + lane]) : 0;
                process_chunk<KIND, CH_NBR>(c, s, valid, f_max, ef, filt, l01);
            }
        }
    }
    if (level) { c.n_dist_up += c.cur_dist; c.n_exp_up += c.cur_exp; c.n_deg_up += c.cur_deg; }
    else { c.n_dist_l0 += c.cur_dist; c.n_exp_l0 += c.cur_exp; c.n_deg_l0 += c.cur_deg; }
}

// ---- query staging ---------------------------------------------------------------------------------------
// UnalignedVector::from_slice + D::new_header (reader.rs:140-141), into the device row layout.
template <int KIND>
__device__ void stage_query(Ctx& c, float* qs, uint64_t qi) {
    const DevIndex& ix = c.p.ix;
    const int lane = lane_id();
    const uint32_t words16 = ix.row_stride / 16;
    uint4* q16 = reinterpret_cast<uint4*>(qs);
    if (c.p.mode & 1) {
        uint32_t slot = c.p.q_slots[qi];
        const uint4* src = reinterpret_cast<const uint4*>(ix.rows + (size_t)slot * ix.row_stride);
        for (uint32_t i = lane; i < words16; i += 32) q16[i] = __ldg(src + i);
    } else {
        for (uint32_t i = lane; i < words16; i += 32) q16[i] = make_uint4(0, 0, 0, 0);
        __syncwarp();
        const float* src = c.p.q + qi * ix.dims;
        if (KIND == KIND_F32_WARP) {
            uint32_t main = ix.dims - ix.tail;
            for (uint32_t e = lane; e < ix.dims; e += 32) {
                float v = __ldg(src + e);
                if (e < main) {
                    uint32_t blk = e >> 5, j = e & 31;
                    qs[(blk >> 2) * 128 + j * 4 + (blk & 3)] = v;
                } else {
                    qs[ix.tail_off + (e - main)] = v;
                }
            }
        } else if (KIND == KIND_F32_LANE) {
            for (uint32_t e = lane; e < ix.dims; e += 32) qs[e] = __ldg(src + e);
        } else {
            // src/unaligned_vector/binary.rs:80-94 (x > 0) and binary_quantized.rs:80-91 (sign positive)
            uint32_t* q32 = reinterpret_cast<uint32_t*>(qs);
            for (uint32_t base = 0; base < ix.dims; base += 32) {
                uint32_t e = base + lane;
                bool bit = false;
                if (e < ix.dims) {
                    uint32_t u = __float_as_uint(__ldg(src + e));
                    bit = (ix.metric == HB_HAMMING) ? (u < 0x80000000u && u > 0u) : ((u >> 31) == 0);
                }
                unsigned w = __ballot_sync(FULL, bit);
                if (lane == 0) q32[base >> 5] = w;  // little-endian halves of the u64 word
            }
        }
    }
    __syncwarp();
    // header
    float qn = 0.0f;
    if (ix.metric == HB_COSINE) {
        float dot;
        if (KIND == KIND_F32_WARP) {
            // dot(q, q) in the same lane order: the query doubles as the "row" (shared-memory reads)
            float acc = 0.0f;
            const float4* q4 = reinterpret_cast<const float4*>(qs);
            for (uint32_t ch = 0; ch < ix.n_chunks; ++ch) {
                float4 v = q4[ch * 32 + lane];
                acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
            }
            dot = warp_hsum_avx(acc);
            for (uint32_t e = 0; e < ix.tail; ++e) { float a = qs[ix.tail_off + e]; dot = __fadd_rn(dot, __fmul_rn(a, a)); }
        } else {
            dot = lane_raw_small<true, false>(qs, qs, ix.dims);
        }
        qn = __fsqrt_rn(dot);  // cosine.rs:58-60
    } else if (ix.metric == HB_BQ_COSINE) {
        qn = __fsqrt_rn((float)(int)(ix.n_words * 64u));  // sqrt(bq_dot(v, v)) = sqrt(padded length)
    }
    c.qs = qs;
    c.qn = qn;
}

// Sort res[0..n) given res[0..first) already sorted (insertion of the rest)
__device__ void sort_tail(u64* a, int first, int n) {
    for (int i = first; i < n; ++i) {
        u64 key = a[i];
        __syncwarp();
        int pos = count_lt(a, i, key);
        if (pos < i) insert_at(a, i, pos, key);
    }
}

// first item not yet in `path`, scanning forward from 32-word window `wb` (the cursor of reader.rs:772-783)
__device__ __forceinline__ uint32_t next_unseen(const Ctx& c, uint32_t& wb) {
    const uint32_t n = c.p.ix.n, n_words = (n + 31) >> 5;
    for (; wb < n_words; wb += 32) {
        uint32_t w = wb + lane_id();
        uint32_t word = 0xffffffffu;
        if (w < n_words) {
            word = __ldcg(&c.vis[w]);
            if (w == n_words - 1 && (n & 31)) word |= ~((1u << (n & 31)) - 1);
        }
        unsigned m = __ballot_sync(FULL, word != 0xffffffffu);
        if (m) {
            int src = __ffs(m) - 1;
            uint32_t wsel = __shfl_sync(FULL, word, src);
            return (wb + src) * 32 + (__ffs(~wsel) - 1);
        }
    }
    return 0xffffffffu;
}

template <int KIND>
__device__ void run_query(const SearchParams& p, uint64_t qi, int slot_idx, u64* heap, float* qs, RowRing& ring) {
    const DevIndex& ix = p.ix;
    const int lane = lane_id();
    Ctx c(p, ring);
    c.res = heap; c.res_cap = p.res_cap; c.res_len = 0;
    c.que = heap + p.res_cap; c.q_cap = p.q_cap; c.q_len = 0;
    c.vis = p.visited + (size_t)slot_idx * p.vis_words;
    c.touched = p.touched + (size_t)slot_idx * p.touched_cap;
    c.touched_len = 0; c.touched_over = false;
    c.excl = 0xffffffffu; c.overflow = false;
    c.n_dist_up = c.n_exp_up = c.n_deg_up = c.n_dist_l0 = c.n_exp_l0 = c.n_deg_l0 = 0;
    c.cur_dist = c.cur_exp = c.cur_deg = 0;
    u64 flags = p.pass ? HB_FLAG_SLOW_PATH : 0;
    const uint32_t count = p.count;
    const int ef0 = (int)max(p.ef_raw, p.count);

    if ((p.mode & 1) && p.q_slots[qi] == 0xffffffffu) {  // item absent -> Ok(None), reader.rs:826
        if (lane == 0) p.out_len[qi] = 0xffffffffu;
        return;
    }
#ifdef HB_PHASES
    for (int i = 0; i < PH_N; ++i) c.ph[i] = 0;
    const long long ph_q0 = clock64();
#endif
    PH_DECL
    stage_query<KIND>(c, qs, qi);
    PH_ADD(c, PH_STAGE)

    int n_out = 0;
    if (p.mode >= 2) {
        // brute_force_search over the candidate slots (ascending) — reader.rs:668-711
        flags |= HB_FLAG_LINEAR;
        for (uint32_t base = 0; base < p.n_cand_slots; base += 32) {
            bool valid = base + lane < p.n_cand_slots;
            uint32_t s = valid ? __ldg(&p.cand_slots[base + lane]) : 0;
            process_chunk<KIND, CH_LINEAR>(c, s, valid, FLT_MAX, (int)count, false, 1);
        }
        c.n_dist_l0 += c.cur_dist;
        n_out = c.res_len;
    } else {
        // One call site of visit() drives the whole search as a small state machine:
        //   ST_UPPER  greedy descent, ef = 1, levels max_level..1, shared visited set (reader.rs:732-741)
        //   ST_L0     the ef-bounded layer-0 walk (reader.rs:743-767) / by_item's seeded walk (reader.rs:842-862)
        //   ST_FB     exhaustive fallback, one visit per unseen item (reader.rs:771-795 / 865-889)
        enum { ST_UPPER, ST_L0, ST_FB };
        uint32_t ep_single = 0, level = 0;
        const uint32_t* eps = ix.eps;
        int st = ST_L0;
        if (p.mode & 1) {  // nns_by_item — reader.rs:836-842
            c.excl = p.q_slots[qi];
            ep_single = c.excl;
            eps = nullptr;
        } else if (ix.max_level >= 1) {
            st = ST_UPPER;
            level = ix.max_level;
        }
        const int target = (p.mode & 1) ? (int)count : (int)p.ef_raw;
        u64* const base_res = c.res;
        const int base_cap = c.res_cap;
        int acc_len = 0, first = -1, ef = 0;
        uint32_t wb = 0;
        for (;;) {
            bool filt = true;
            if (st == ST_UPPER) { ef = 1; filt = false; }
            else if (st == ST_L0) { ef = ef0; level = 0; }
            visit<KIND>(c, eps, ix.n_ep, ep_single, level, ef, filt);
            if (c.overflow) break;
            if (st == ST_UPPER) {
                bool found = c.res_len != 0;       // reference: expect("No neighbor was found")
                if (found) { ep_single = (uint32_t)c.res[0]; eps = nullptr; }  // peek_min
                __syncwarp();
                --level;
                if (level == 0 || !found) {
                    vis_clear(c);                  // path.clear(), reader.rs:743
                    PH_ADD(c, PH_UPPER)
                    st = ST_L0;
                }
                continue;
            }
            if (st == ST_L0) {
                PH_RESET
                acc_len = c.res_len;
                if (acc_len >= (int)count) break;
                flags |= HB_FLAG_FALLBACK;
                first = acc_len;
                st = ST_FB;
            } else {
                acc_len += c.res_len;
                __threadfence_block();
                if (acc_len >= target) break;
            }
            uint32_t s = next_unseen(c, wb);
            if (s == 0xffffffffu) break;
            ep_single = s;
            eps = nullptr;
            ef = (p.mode & 1) ? (int)count - acc_len : max(0, (int)p.ef_raw - acc_len);
            c.res = base_res + acc_len;
            c.res_cap = base_cap - acc_len;
        }
        c.res = base_res; c.res_cap = base_cap;
        if (first >= 0) {
            c.res_len = acc_len;
            if (!c.overflow) sort_tail(c.res, first, acc_len);
        }
        n_out = min(c.res_len, (int)count);
        vis_clear(c);
    }

    if (c.overflow) {
        // heaps too small for this query: hand it to the global-memory pass
        vis_clear(c);
        if (p.pass == 0) {
            if (lane == 0) { uint32_t at = atomicAdd(p.n_overflow, 1u); p.overflow_list[at] = (uint32_t)qi; }
            return;
        }
        n_out = 0;  // cannot happen: pass-1 capacities cover every push
        flags |= 0x100;
    }
    __syncwarp();
    for (int i = lane; i < n_out; i += 32) {
        u64 k = c.res[i];
        p.out_ids[qi * count + i] = __ldg(&ix.ids[(uint32_t)k]);
        p.out_dist[qi * count + i] = key_dist(k);
    }
    if (lane == 0) {
        p.out_len[qi] = (uint32_t)n_out;
        if (p.out_ctr) {
            uint64_t* o = p.out_ctr + qi * HB_N_CTR;
            o[HB_CTR_DIST_UPPER] = c.n_dist_up; o[HB_CTR_DIST_L0] = c.n_dist_l0;
            o[HB_CTR_EXP_UPPER] = c.n_exp_up; o[HB_CTR_EXP_L0] = c.n_exp_l0;
            o[HB_CTR_DEG_UPPER] = c.n_deg_up; o[HB_CTR_DEG_L0] = c.n_deg_l0;
            o[HB_CTR_FLAGS] = flags; o[HB_CTR_RESERVED] = 0;
        }
    }
    __syncwarp();
#ifdef HB_PHASES
    PH_ADD(c, PH_TAIL)
    c.ph[PH_TOTAL] = clock64() - ph_q0;
    if (lane == 0)
        for (int i = 0; i < PH_N; ++i) atomicAdd(&g_phase[i], (unsigned long long)c.ph[i]);
#endif
}

// Shared memory of one warp: [row ring][ring barriers][query][heaps (pass 0 only)].
template <int KIND>
__global__ void __launch_bounds__(SEARCH_WARPS_PER_BLOCK * 32, KIND == KIND_F32_WARP ? 3 : 4) hnsw_search_kernel(const __grid_constant__ SearchParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp_in_block = threadIdx.x >> 5;
    const int warps_per_block = blockDim.x >> 5;
    const int slot_idx = blockIdx.x * warps_per_block + warp_in_block;
    const size_t ring_bytes = (size_t)p.ring_slots * p.ring_stride;
    const size_t bar_bytes = ((size_t)p.ring_slots * 8 + 15) & ~(size_t)15;
    const size_t heap_bytes = p.pass == 0 ? (size_t)(p.res_cap + p.q_cap) * 8 : 0;
    const size_t per_warp = (ring_bytes + bar_bytes + p.q_smem_bytes + heap_bytes + 127) & ~(size_t)127;
    unsigned char* base = smem + per_warp * warp_in_block;
    float* qs = reinterpret_cast<float*>(base + ring_bytes + bar_bytes);
    u64* heap = p.pass == 0 ? reinterpret_cast<u64*>(base + ring_bytes + bar_bytes + p.q_smem_bytes)
                            : p.gheap + (size_t)slot_idx * (p.res_cap + p.q_cap);
    RowRing ring;
    ring.ptr = base;
    ring.data = smem_addr(base);
    ring.bars = smem_addr(base + ring_bytes);
    ring.slots = p.ring_slots;
    ring.stride = p.ring_stride;
    ring.phase = 0;
    ring.policy = l2_policy_evict_first();
    if (p.ring_slots) {
        if (lane_id() == 0) {
            for (uint32_t i = 0; i < p.ring_slots; ++i) mbar_init(ring.bars + i * 8, 1);
            mbar_fence_init();
        }
        __syncwarp();
    }
    const uint32_t n_work = p.pass == 0 ? p.n_work : *p.n_overflow;
    for (;;) {
        unsigned long long w = 0;
        if (lane_id() == 0) w = atomicAdd(p.work_counter + p.pass, 1ull);
        w = __shfl_sync(FULL, w, 0);
        if (w >= n_work) break;
        uint64_t qi = p.pass == 0 ? w : p.overflow_list[w];
        run_query<KIND>(p, qi, slot_idx, heap, qs, ring);
    }
}

unsigned long long g_launches = 0;

// dev: read and reset the phase timers (all zero unless built with -DHB_PHASES)
void read_phases(unsigned long long* out) {
    cudaMemcpyFromSymbol(out, g_phase, sizeof(unsigned long long) * PH_N);
    unsigned long long z[16] = {};
    cudaMemcpyToSymbol(g_phase, z, sizeof(z));
}

__global__ void fill_iota_kernel(uint32_t* list, uint32_t* n_out, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) list[i] = i;
    if (i == 0) *n_out = n;
}

size_t search_smem_per_warp(const SearchParams& p) {
    size_t ring_bytes = (size_t)p.ring_slots * p.ring_stride;
    size_t bar_bytes = ((size_t)p.ring_slots * 8 + 15) & ~(size_t)15;
    size_t heap_bytes = p.pass == 0 ? (size_t)(p.res_cap + p.q_cap) * 8 : 0;
    return (ring_bytes + bar_bytes + p.q_smem_bytes + heap_bytes + 127) & ~(size_t)127;
}

typedef void (*search_kernel_t)(const SearchParams);
static search_kernel_t kernel_for(int kind) {
    switch (kind) {
        case KIND_F32_WARP: return hnsw_search_kernel<KIND_F32_WARP>;
        case KIND_F32_LANE: return hnsw_search_kernel<KIND_F32_LANE>;
        default: return hnsw_search_kernel<KIND_BIN>;
    }
}
static void set_kernel_attrs() {
    static bool attr_set = false;
    if (!attr_set) {
        for (int k = 0; k < 3; ++k)
            cudaFuncSetAttribute(kernel_for(k), cudaFuncAttributeMaxDynamicSharedMemorySize, SEARCH_MAX_SMEM);
        attr_set = true;
    }
}

int search_blocks_per_sm(const SearchParams& p) {
    set_kernel_attrs();
    int nb = 0;
    size_t smem = search_smem_per_warp(p) * SEARCH_WARPS_PER_BLOCK;
    if (smem > (size_t)SEARCH_MAX_SMEM) return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel_for(p.ix.kind), SEARCH_WARPS_PER_BLOCK * 32, smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return nb;
}

// Launch pass 0 (shared-memory heaps, `fast`) then pass 1 (global-memory heaps over the overflow list,
// `slow`).  fast.res_cap == 0 means the heaps do not fit shared memory: every query takes pass 1.
hb_status launch_search(const SearchParams& fast, const SearchParams& slow, int blocks_fast, int blocks_slow, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const int wpb = SEARCH_WARPS_PER_BLOCK;
    set_kernel_attrs();
    cudaMemsetAsync(fast.work_counter, 0, 2 * sizeof(unsigned long long), stream);
    if (fast.res_cap) {
        cudaMemsetAsync(fast.n_overflow, 0, sizeof(uint32_t), stream);
        size_t smem = search_smem_per_warp(fast) * wpb;
        uint64_t need = ((uint64_t)fast.n_work + wpb - 1) / wpb;
        int blocks = (uint64_t)blocks_fast > need ? (int)need : blocks_fast;
        if (blocks < 1) blocks = 1;
        kernel_for(fast.ix.kind)<<<blocks, wpb * 32, smem, stream>>>(fast);
        ++g_launches;
    } else {
        uint32_t n = fast.n_work;
        fill_iota_kernel<<<(n + 255) / 256, 256, 0, stream>>>(fast.overflow_list, fast.n_overflow, n);
        ++g_launches;
    }
    size_t smem = search_smem_per_warp(slow) * wpb;
    kernel_for(slow.ix.kind)<<<blocks_slow, wpb * 32, smem, stream>>>(slow);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("search launch failed: %s", cudaGetErrorString(e)); return HB_ECUDA; }
    return HB_OK;
}

}  // namespace hb
