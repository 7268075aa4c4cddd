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
//     heaps in global memory (never on the CPU);
//   * a warp that finds no query left does not retire: it becomes a helper of the warps of its CTA that are still
//     walking and gathers a share of the rows of each of their expansions through its own ring (TeamShared below) —
//     the tail of a big batch, a batch smaller than the resident warps (one 10k batch split over 8 GPUs) and a
//     single query all get up to 4 rings per query instead of one.
#include <algorithm>
#include <cfloat>
#include <cstdio>
#include <cstdlib>

#include "dist.cuh"
#include "ring.cuh"
#include "sorted.cuh"

#ifndef HB_DIRECT_R
#define HB_DIRECT_R 8  // rows per group of the direct gather when a row is one 128-float chunk
#endif
#ifndef HB_MIN_BLOCKS_DIRECT
#define HB_MIN_BLOCKS_DIRECT 4
#endif
#ifndef HB_MIN_BLOCKS_BIN
#define HB_MIN_BLOCKS_BIN 6  // binary codes: one lane per row, no ring -> shared memory is not the limit, registers are (80 per thread;
#endif                       // the walk is a dependent chain per warp, so more resident warps beat a few spilled cold values: 4 -> 6 CTAs/SM = +15 % on C4s)
#ifndef HB_ADJ_PREFETCH_ALL
#define HB_ADJ_PREFETCH_ALL 1
#endif
#ifndef HB_EARLY_ROWS
#define HB_EARLY_ROWS 1  // layer-0 deferred pops: request the rows of an expansion before merging the previous chunk into the heaps
#endif
#ifndef HB_SPEC_VIS
#define HB_SPEC_VIS 1      // prefetch the visited-set words of the expected next expansion's neighbours (f32 kernels; worth ~1 % on C2 — the
                           // larger gains first measured for it were differences between two compilations of the same source, build.py)
#endif
#ifndef HB_VIS_LOAD
#define HB_VIS_LOAD 1
#endif
#ifndef HB_LEAN
#define HB_LEAN 1          // bookkeeping trimmed from the per-expansion chain: traversal counters only when asked for, the negative-distance
#endif                     // test only for the one metric that can produce one, the prune test's invariants hoisted
#ifndef HB_OPAQUE_ADDR
#define HB_OPAQUE_ADDR 0
#endif
#ifndef HB_TEAM_SHARE
#define HB_TEAM_SHARE 0
#endif
#ifndef HB_UPPER_KEEP
#define HB_UPPER_KEEP 1    // rows gathered on the upper layers are kept in L2 (evict_last: every query descends through the same few thousand
                           // nodes), layer-0 rows stream through it (evict_first): C3 11.58 -> 11.36 ms, C2 5.22 -> 5.18 ms
#endif
#ifndef HB_SPEC_DEDUPE_F32
#define HB_SPEC_DEDUPE_F32 0
#endif
#ifndef HB_SPEC_DEDUPE
#define HB_SPEC_DEDUPE 1
#endif
#ifndef HB_SPEC_VIS_BIN
#define HB_SPEC_VIS_BIN 0  // ... not in the binary kernel: +17 GB of DRAM reads per C4s launch (speculation that fails) for < 1 %
#endif
#ifndef HB_MIN_BLOCKS_F32
#define HB_MIN_BLOCKS_F32 3  // resident CTAs per SM the f32 ring kernel is compiled for (register budget): long rows, shared memory allows no more
#endif
#ifndef HB_MIN_BLOCKS_F32_SHORT
#define HB_MIN_BLOCKS_F32_SHORT 4  // ... and its second instantiation for rows short enough that four CTAs' rings fit one SM (C2: +10 %)
#endif

namespace hb {

// Kernel variants: the three row kinds of common.h, plus KIND_F32_WARP rows gathered with plain 128-bit loads instead
// of the bulk-copy ring: for rows too long to stage in shared memory (and, optionally, for short rows — the
// bulk-copy engine tops out near one copy per ~72 cycles per SM, tools/gather_bench.cu).  8 (one-chunk rows) or 4
// rows go through registers at a time.
constexpr int KIND_F32_DIRECT = 3;
__host__ __device__ constexpr bool is_f32_warp(int kind) { return kind == KIND_F32_WARP || kind == KIND_F32_DIRECT; }

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
// Optional event trace of ONE warp (build with -DHB_TRACE; dev only): (clock64 << 8 | event) records.
#ifdef HB_TRACE
#define HB_TRACE_CAP (1 << 18)
__device__ unsigned long long g_trace[HB_TRACE_CAP];
__device__ unsigned int g_trace_n;
#define TR(c, ev) { if ((c).tr && lane_id() == 0) { unsigned int tn_ = g_trace_n; if (tn_ < HB_TRACE_CAP) { g_trace[tn_] = ((unsigned long long)clock64() << 8) | (ev); g_trace_n = tn_ + 1; } } }
#else
#define TR(c, ev) {}
#endif
enum { TR_QSTART = 1, TR_POP = 2, TR_ADJ = 3, TR_VIS = 4, TR_POSTED = 5, TR_ROWWAIT = 6, TR_GROUP = 7, TR_HEAP = 8, TR_QEND = 9, TR_L0 = 10 };
enum { PH_STAGE = 0, PH_UPPER = 1, PH_ADJ = 2, PH_VIS = 3, PH_ROWS = 4, PH_HEAP = 5, PH_TAIL = 6, PH_TOTAL = 7, PH_POST = 8, PH_COLLECT = 9, PH_ACCEPT = 10, PH_DECIDE = 11, PH_N = 12 };

struct Ctx {
    const SearchParams& p;
    u64* res; int res_len; int res_cap;   // ascending (bits<<32 | slot)
    u64* que; int q_len; int q_cap;       // descending (bits<<32 | ~slot): next to pop is the LAST element
    uint32_t* vis; uint32_t* touched; uint32_t touched_len; bool touched_over;
    const float* qs; float qn;            // query (device layout) in shared memory, query header norm
    uint32_t excl;                        // by_item: slot removed from the candidates, else UINT32_MAX
    bool overflow;
    bool trim_ok;                         // this visit may trim dead queue entries (pass 0, no poll-exact cancellation, no negative distances expected)
    bool cancelled;                       // the visit in progress returned Completion::Cancelled(res)
    uint32_t polls;                       // calls of cancel_fn made by this query so far
    bool tr;                              // this warp writes the event trace (HB_TRACE builds)
    RowRing& ring;                        // per-warp, lives across queries (barrier phases persist)
    struct TeamShared* ts = nullptr;      // CTA-shared helper state (f32 ring kernel), or nullptr
    uint32_t* team = nullptr;             // per-warp: bits 0-3 helpers attached to this warp, bits 4-7 their done-barrier parities
#ifdef HB_PHASES
    long long ph[PH_N];
#endif
    uint32_t cur_dist, cur_exp, cur_deg;  // counters of the visit in progress
    u64 n_dist_up, n_exp_up, n_deg_up, n_dist_l0, n_exp_l0, n_deg_l0;
    __device__ Ctx(const SearchParams& pp, RowRing& rr) : p(pp), ring(rr) {}
};

__device__ __forceinline__ float key_dist(u64 k) { return __uint_as_float((uint32_t)(k >> 32)); }

// `cancel_fn()` — reader.rs:330.  cancel_next: would the next call return true (no call is made);
// cancel_poll: the call itself.
__device__ __forceinline__ bool cancel_flag_set(const Ctx& c) {
    return c.p.cancel_flag && *reinterpret_cast<const volatile uint32_t*>(c.p.cancel_flag) != 0;
}
__device__ __forceinline__ bool cancel_next(const Ctx& c) {
    return (c.p.cancel_after && c.polls + 1 >= c.p.cancel_after) || cancel_flag_set(c);
}
__device__ __forceinline__ bool cancel_poll(Ctx& c) {
    ++c.polls;
    return (c.p.cancel_after && c.polls >= c.p.cancel_after) || cancel_flag_set(c);
}

// An entry of the search queue can never be popped again once the result set is full and its distance exceeds the
// current f_max: f_max never grows while res.len() >= ef, and the loop breaks at the first `f > f_max`
// (reader.rs:333-336).  Such dead entries are trimmed from the queue (heaps_stage_queue).  Only argued for non-negative
// distances (bit order == numeric order).
//
// Dead entries are also the reference's "break sentinels": popping one ends the walk before any entry that follows it
// in BIT order.  For non-negative distances every follower would end the walk itself, so dropping the sentinel changes
// nothing; a negative distance (only BinaryQuantizedCosine can produce one, binary_quantized_cosine.rs:49-58 has no
// clamp) sorts last by bits yet passes `f > f_max`, so the first negative distance seen in the trimming pass sends the
// query to pass 1, which never trims.
enum ChunkMode { CH_EP, CH_NBR, CH_LINEAR };

// ---- visited set --------------------------------------------------------------------------------------
// `path.insert(point)` in two halves so that independent work can sit between the atomic and its use:
// vis_issue sends the atomicOr (returns the old word), vis_finish turns it into `fresh` and logs the touched slots.
// HB_VIS_LOAD: the word is READ (L2, bypassing L1) and only the lanes whose point is fresh set their bit afterwards, with a
// reduction that returns nothing — a lookup of an already visited point (more than half of them) then leaves its sector clean,
// where an atomicOr dirties it whatever the answer and costs a 32-byte write-back.  The bitset is this warp's alone; the
// __syncwarp after the reductions orders them before the next chunk's reads by other lanes.
// (Not in the graph builder: its mutable lists may name a neighbour twice, and only the atomic's serialisation tells the two
// lanes of one chunk apart — SearchParams::vis_atomic.)
__device__ __forceinline__ uint32_t vis_issue(Ctx& c, uint32_t s, bool valid) {
#if HB_VIS_LOAD
    if (!c.p.vis_atomic) return valid ? __ldcg(&c.vis[s >> 5]) : 0xffffffffu;
#endif
    return valid ? atomicOr(&c.vis[s >> 5], 1u << (s & 31)) : 0xffffffffu;
}
// vis_finish leaves the slot's place in the touched list in `log_at` (UINT32_MAX: nothing to log); the store itself
// (vis_log) is issued by the caller AFTER the row copies and helper jobs of the chunk were posted: an mbarrier arrive is a
// release at CTA scope, and it would otherwise wait for this global store to be acknowledged.
__device__ __forceinline__ bool vis_finish(Ctx& c, uint32_t s, bool valid, uint32_t old, uint32_t& log_at) {
    const bool fresh = valid && !((old >> (s & 31)) & 1u);
#if HB_VIS_LOAD
    if (!c.p.vis_atomic) {
        if (fresh) atomicOr(&c.vis[s >> 5], 1u << (s & 31));   // result unused: a reduction
        __syncwarp();
    }
#endif
    unsigned m = __ballot_sync(FULL, fresh);
    log_at = 0xffffffffu;
    if (m) {
        uint32_t r = __popc(m & ((1u << lane_id()) - 1));
        uint32_t at = c.touched_len + r;
        if (fresh && at < c.p.touched_cap) log_at = at;
        c.touched_len += __popc(m);
        if (c.touched_len > c.p.touched_cap) c.touched_over = true;
    }
    return fresh;
}
__device__ __forceinline__ void vis_log(Ctx& c, uint32_t s, uint32_t log_at) {
    if (log_at != 0xffffffffu) c.touched[log_at] = s;
}
// path.clear() — reader.rs:743
__device__ __forceinline__ void vis_clear(Ctx& c) {
    __syncwarp();
    if (c.touched_over) {
        for (uint32_t w = lane_id(); w < c.p.vis_words; w += 32) c.vis[w] = 0;
    } else {
        // the list reads are independent of each other: four in flight per lane instead of one
        uint32_t i = lane_id();
        for (; i + 96 < c.touched_len; i += 128) {
            const uint32_t t0 = c.touched[i], t1 = c.touched[i + 32], t2 = c.touched[i + 64], t3 = c.touched[i + 96];
            c.vis[t0 >> 5] = 0; c.vis[t1 >> 5] = 0; c.vis[t2 >> 5] = 0; c.vis[t3 >> 5] = 0;
        }
        for (; i < c.touched_len; i += 32) c.vis[c.touched[i] >> 5] = 0;
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
// Split in two so that a caller can do unrelated work (the previous expansion's heap update) while the first
// rows are on their way: rows_begin posts / prefetches, rows_finish consumes and returns the lane's distance.
struct RowsInFlight {
    unsigned mask;        // lanes that own a live row
    int n_live, rank;     // number of live rows, this lane's rank among them
    bool has;
    const uint8_t* grow;  // this lane's row in global memory
    float in;             // this lane's item header (norm)
    uint32_t g1, g2, g3;  // gather4 ring: the slots of the three live rows after this lane's (UINT32_MAX past the last), for lanes of rank % 4 == 0
};

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int KIND>
__device__ __forceinline__ void rows_begin(Ctx& c, unsigned mask, uint32_t s, RowsInFlight& rf) {
    const DevIndex& ix = c.p.ix;
    const int lane = lane_id();
    rf.mask = mask;
    rf.n_live = __popc(mask);
    rf.has = (mask >> lane) & 1;
    rf.rank = __popc(mask & ((1u << lane) - 1));
    rf.grow = ix.rows + (size_t)s * ix.row_stride;
    rf.in = 0.0f;
    if (KIND == KIND_F32_WARP && c.p.gather4) {
        // four rows per copy instruction: the lane that owns live row 4g collects the slots of rows 4g+1..4g+3
        // (next set bit above a lane: clear the bits up to it, take the lowest — __fns would be a loop)
        const unsigned m1 = mask & ~((2u << lane) - 1u);
        const int l1 = __ffs(m1) - 1;
        const unsigned m2 = m1 & (m1 - 1u);
        const int l2 = __ffs(m2) - 1;
        const int l3 = __ffs(m2 & (m2 - 1u)) - 1;
        const uint32_t s1 = __shfl_sync(FULL, s, l1 & 31), s2 = __shfl_sync(FULL, s, l2 & 31), s3 = __shfl_sync(FULL, s, l3 & 31);
        rf.g1 = l1 >= 0 ? s1 : ix.n; rf.g2 = l2 >= 0 ? s2 : ix.n; rf.g3 = l3 >= 0 ? s3 : ix.n;   // a row index past the end reads as zeros
    }
    if (!rf.has) return;
    if (KIND == KIND_F32_WARP) {
        // Row r (in ascending-lane order) lands in ring slot r % S: the first S copies are posted at once by the
        // lanes that own them.
        if (c.p.gather4) {
            if ((rf.rank & 3) == 0 && rf.rank < (int)c.ring.slots) c.ring.post_gather4(rf.rank, c.p.rows_tmap, s, rf.g1, rf.g2, rf.g3);
        } else if (rf.rank < (int)c.ring.slots) c.ring.post(rf.rank, rf.grow, ix.row_stride);
        if (ix.metric == HB_COSINE) rf.in = __ldg(&ix.hdr[s]);
    } else if (KIND == KIND_F32_DIRECT) {
        if (ix.metric == HB_COSINE) rf.in = __ldg(&ix.hdr[s]);
    } else {
        for (uint32_t o = 0; o < ix.row_stride; o += 128) prefetch_l2(rf.grow + o);
        if (ix.metric == HB_COSINE || ix.metric == HB_BQ_COSINE) rf.in = ix.hdr_uniform ? ix.hdr_value : __ldg(&ix.hdr[s]);
    }
}

// R live rows at a time straight from global memory: lane j loads float4 j of every chunk of each row (coalesced
// 512-byte lines), all R x chunks loads of a group are independent and in flight together.
template <int R>
__device__ __forceinline__ float direct_rows(Ctx& c, const RowsInFlight& rf, uint32_t s) {
    const DevIndex& ix = c.p.ix;
    float myraw = 0.0f;
    for (int r0 = 0; r0 < rf.n_live; r0 += R) {
        const int g = min(R, rf.n_live - r0);
        const uint8_t* rowp[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int owner = __fns(rf.mask, 0, r0 + (r < g ? r : 0) + 1);  // lane of the (r0 + r)-th live row
            rowp[r] = ix.rows + (size_t)__shfl_sync(FULL, s, owner) * ix.row_stride;
        }
        const float red = (ix.metric == HB_COSINE) ? warp_rows_group<R, true, false>(ix, c.qs, rowp) : warp_rows_group<R, false, false>(ix, c.qs, rowp);
        const int mr = rf.rank - r0;
        const float got = __shfl_sync(FULL, red, group_owner<R>(mr >= 0 && mr < R ? mr : 0));
        if (rf.has && mr >= 0 && mr < g) myraw = got;
    }
    return myraw;
}

template <int KIND>
__device__ __forceinline__ float rows_finish(Ctx& c, const RowsInFlight& rf, uint32_t s) {
    const DevIndex& ix = c.p.ix;
    const int lane = lane_id();
    float mine = 0.0f;
    if (KIND == KIND_F32_DIRECT) {
        float myraw = 0.0f;
        if (ix.n_chunks <= 1) myraw = direct_rows<HB_DIRECT_R>(c, rf, s);
        else myraw = direct_rows<4>(c, rf, s);
        if (rf.has) mine = finish_f32(ix.metric, myraw, c.qn, rf.in);
    } else if (KIND == KIND_F32_WARP) {
        // slot group g is re-posted as soon as its ROW_GROUP rows were consumed
        const int S = (int)c.ring.slots;
        const bool g4 = c.p.gather4 != 0;
        const uint32_t stride = c.ring.stride;
        int slot0 = 0;
        const uint8_t* gp = c.ring.ptr;            // first slot of the group being consumed
        float myraw = 0.0f;
        for (int r0 = 0; r0 < rf.n_live; r0 += ROW_GROUP) {
            const int g = min(ROW_GROUP, rf.n_live - r0);
            const uint8_t* rowp[ROW_GROUP];
            if (g4) {
#pragma unroll
                for (int r = 0; r < ROW_GROUP; r += 4)
                    if (r < g) { c.ring.wait(slot0 + r); TR(c, TR_ROWWAIT) }   // four gathered rows complete on their first slot's barrier; rows past the last read as zeros
#pragma unroll
                for (int r = 0; r < ROW_GROUP; ++r) rowp[r] = gp + (ROW_GROUP > 4 && r >= ((g + 3) & ~3) ? 0 : r) * stride;
            } else {
#pragma unroll
                for (int r = 0; r < ROW_GROUP; ++r) {
                    if (r < g) { c.ring.wait(slot0 + r); TR(c, TR_ROWWAIT) }
                    rowp[r] = gp + (r < g ? r : 0) * stride;
                }
            }
            // row r's sum comes back on lane group_owner(r); the lane that owns the row picks it up
            float red = (ix.metric == HB_COSINE) ? warp_rows_group<ROW_GROUP, true, true>(ix, c.qs, rowp) : warp_rows_group<ROW_GROUP, false, true>(ix, c.qs, rowp);
            __syncwarp();  // every lane has read the group's slots: they may be overwritten
            const int mr = rf.rank - r0;
            const int nxt = mr - S;
            if (g4) {
                if (rf.has && nxt >= 0 && nxt < ROW_GROUP && (nxt & 3) == 0) c.ring.post_gather4(slot0 + nxt, c.p.rows_tmap, s, rf.g1, rf.g2, rf.g3);
            } else if (rf.has && nxt >= 0 && nxt < g) c.ring.post(slot0 + nxt, rf.grow, ix.row_stride);
            const float got = __shfl_sync(FULL, red, group_owner<ROW_GROUP>(mr & (ROW_GROUP - 1)));
            if (rf.has && mr >= 0 && mr < g) myraw = got;
            slot0 += ROW_GROUP;
            gp += ROW_GROUP * stride;
            if (slot0 >= S) { slot0 = 0; gp = c.ring.ptr; }
            TR(c, TR_GROUP)
        }
        if (rf.has) mine = finish_f32(ix.metric, myraw, c.qn, rf.in);
    } else if (KIND == KIND_F32_LANE) {
        if (rf.has) mine = lane_distance_f32<true>(ix, c.qs, c.qn, reinterpret_cast<const float*>(rf.grow), rf.in);
    } else {
        if (rf.has) {
            uint32_t h = lane_xor_popc(reinterpret_cast<const uint64_t*>(c.qs), reinterpret_cast<const uint64_t*>(rf.grow), ix.n_words);
            mine = finish_bin(ix.metric, h, ix.n_words * 64u, c.qn, rf.in);
        }
    }
    return mine;
}

// ---- team: idle warps of a CTA gather rows for the ones still walking -----------------------------------------------------
// Work is handed out per warp; a warp that finds the counter exhausted marks itself FREE and waits on its job barrier.  A
// warp that is still walking (a "leader") claims FREE warps of its CTA (atomicCAS on owner[h]) and from then on splits
// the live rows of every chunk: it keeps the first share, posts the slots of the others into the helpers' mailboxes
// (mbarrier arrive = release), gathers its own share, then waits for the helpers' done barriers and picks their
// distances up.  A helper computes the whole D::distance of its rows with the same warp-wide routine (rows_begin /
// rows_finish through its own ring) against the leader's staged query, so every value is bit-identical to what the
// leader would have computed.  Helpers exist only once the work counter has run dry, so a leader never takes another
// query while it holds helpers: they are released when the leader itself goes idle.  The CTA retires when all four
// warps are idle.
constexpr uint32_t TEAM_BUSY = 0xfeu, TEAM_FREE = 0xffu;
struct TeamShared {
    unsigned long long job_bar[SEARCH_WARPS_PER_BLOCK];   // count 1: a job was posted for helper h
    unsigned long long done_bar[SEARCH_WARPS_PER_BLOCK];  // count 1: helper h has written its distances
    uint32_t owner[SEARCH_WARPS_PER_BLOCK];               // TEAM_BUSY | TEAM_FREE | warp index of the leader holding it
    uint32_t done_parity[SEARCH_WARPS_PER_BLOCK];         // parity the next wait on done_bar[h] must observe (handed from holder to holder)
    uint32_t n_idle;
    // a job, written by the leader for helper h: the leader's staged query, the range [lo, lo + n) of live rows (by rank) to do
    struct Job { const float* qs; float qn; uint32_t lo, n, leader; } job[SEARCH_WARPS_PER_BLOCK];
    uint32_t slot[SEARCH_WARPS_PER_BLOCK][32];   // per LEADER warp: the slots of the chunk's live rows, by rank
    float dist[SEARCH_WARPS_PER_BLOCK][32];      // per LEADER warp: their distances, by rank (each helper fills its range)
};
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {   // .release.cta
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {  // .acquire.cta; may give up after the hardware time limit
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ uint32_t ld_volatile_shared(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }

// leader: claim helpers if warps of this CTA have gone idle (at most its fair share of them)
__device__ __forceinline__ void team_update(Ctx& c) {
    TeamShared& ts = *c.ts;
    const uint32_t idle = ld_volatile_shared(&ts.n_idle);
    if (idle == 0) return;
    const int have = __popc(*c.team & 0xfu);
    const int want = max(1, (int)(idle / max(1u, (uint32_t)SEARCH_WARPS_PER_BLOCK - idle)));
    if (have >= want) return;
    uint32_t got = 0;
    if (lane_id() == 0) {
        const uint32_t me = threadIdx.x >> 5;
        int n = have;
        for (uint32_t h = 0; h < (uint32_t)SEARCH_WARPS_PER_BLOCK && n < want; ++h) {
            if (h == me || ld_volatile_shared(&ts.owner[h]) != TEAM_FREE) continue;
            if (atomicCAS(&ts.owner[h], TEAM_FREE, me) != TEAM_FREE) continue;
            __threadfence_block();  // acquire: the previous holder wrote done_parity[h] before it freed h
            got |= (1u << h) | ((ld_volatile_shared(&ts.done_parity[h]) & 1u) << (4 + h));
            ++n;
        }
    }
    got = __shfl_sync(FULL, got, 0);
    *c.team |= got;
}
// leader going idle: hand its helpers back
__device__ __forceinline__ void team_release(TeamShared& ts, uint32_t& team) {
    __syncwarp();
    if (lane_id() == 0) {
        for (unsigned m = team & 0xfu; m; m &= m - 1) {
            const int h = __ffs(m) - 1;
            ts.done_parity[h] = (team >> (4 + h)) & 1u;
            __threadfence_block();
            *reinterpret_cast<volatile uint32_t*>(&ts.owner[h]) = TEAM_FREE;
        }
    }
    team = 0;
    __syncwarp();
}

// the j-th (j = 0..2) set bit of a 4-bit helper mask
__device__ __forceinline__ int team_nth(unsigned hm, int j) {
    const unsigned m1 = hm & (hm - 1u), m2 = m1 & (m1 - 1u);
    return __ffs(j == 0 ? hm : (j == 1 ? m1 : m2)) - 1;
}

// leader: share the live rows of one chunk out over the attached helpers.  Rows are split in order of rank, in whole
// reduction groups: the leader keeps [0, per), the j-th helper takes [j * per, (j + 1) * per).  Everything is done once for
// all helpers — a lone warp pays every instruction's latency in full, so the hand-off is a handful of instructions: all live
// lanes store their slot by rank, lane j describes helper j's range and arrives on its job barrier.  team_post returns the mask
// of LANES that manage a helper in use; team_collect has those lanes wait for their helper, then every lane picks its distance.
__device__ __forceinline__ unsigned team_post(Ctx& c, unsigned lm, uint32_t s, int per) {
    TeamShared& ts = *c.ts;
    const int lane = lane_id();
    const uint32_t me = threadIdx.x >> 5;
    const int n_live = __popc(lm);
    const int rank = __popc(lm & ((1u << lane) - 1));
    if ((lm >> lane) & 1) ts.slot[me][rank] = s;
    const unsigned hm = *c.team & 0xfu;
    const int lo = (lane + 1) * per;
    const bool mine = lane < __popc(hm) && lo < n_live;      // lane j manages the j-th attached helper
    const int h = team_nth(hm, lane);
    if (mine) {
        TeamShared::Job& jb = ts.job[h];
        jb.qs = c.qs; jb.qn = c.qn; jb.lo = (uint32_t)lo; jb.n = (uint32_t)(min(n_live, lo + per) - lo); jb.leader = me;
    }
    __syncwarp();
    if (mine) mbar_arrive(smem_addr(&ts.job_bar[h]));
    return __ballot_sync(FULL, mine);
}
__device__ __forceinline__ float team_collect(Ctx& c, unsigned lm, unsigned used, int per, float d) {
    TeamShared& ts = *c.ts;
    const int lane = lane_id();
    const uint32_t me = threadIdx.x >> 5;
    const unsigned hm = *c.team & 0xfu;
    unsigned flip = 0;
    if ((used >> lane) & 1) {
        const int h = team_nth(hm, lane);
        const uint32_t bar = smem_addr(&ts.done_bar[h]);
        while (!mbar_try_wait(bar, (*c.team >> (4 + h)) & 1u)) {}
        flip = 1u << (4 + h);
    }
    __syncwarp();                               // the waiting lanes' acquire is ordered before every lane's read below
    *c.team ^= __reduce_or_sync(FULL, flip);
    const int rank = __popc(lm & ((1u << lane) - 1));
    if (((lm >> lane) & 1) && rank >= per) d = ts.dist[me][rank];
    return d;   // (the next hand-off's __syncwarp, before its arrive, keeps a helper's next write behind these reads)
}

// helper: serve jobs until every warp of the CTA is idle
template <int KIND>
__device__ __noinline__ void team_help(const SearchParams& p, TeamShared& ts, RowRing ring) {
    const int lane = lane_id();
    const uint32_t me = threadIdx.x >> 5;
    const uint32_t jbar = smem_addr(&ts.job_bar[me]), dbar = smem_addr(&ts.done_bar[me]);
    uint32_t par = 0;
    for (;;) {
        while (!mbar_try_wait(jbar, par))
            if (ld_volatile_shared(&ts.n_idle) >= (uint32_t)SEARCH_WARPS_PER_BLOCK) return;
        par ^= 1u;
        const TeamShared::Job& jb = ts.job[me];
        Ctx c(p, ring);
        c.qs = jb.qs; c.qn = jb.qn; c.tr = false;
        const uint32_t n = jb.n, lo = jb.lo, leader = jb.leader;
        const bool has = (uint32_t)lane < n;
        const uint32_t s = has ? ts.slot[leader][lo + lane] : 0u;
        RowsInFlight rf;
        rows_begin<KIND>(c, __ballot_sync(FULL, has), s, rf);
        const float d = rows_finish<KIND>(c, rf, s);
        if (has) ts.dist[leader][lo + lane] = d;
        __syncwarp();
        if (lane == 0) mbar_arrive(dbar);
    }
}

__device__ __forceinline__ u64 warp_min_u64(u64 v) {
    uint32_t hi = (uint32_t)(v >> 32);
    uint32_t mhi = __reduce_min_sync(FULL, hi);
    uint32_t lo = hi == mhi ? (uint32_t)v : 0xffffffffu;
    uint32_t mlo = __reduce_min_sync(FULL, lo);
    return ((u64)mhi << 32) | mlo;
}

// The heap update of one chunk of <= 32 points (lane i holds the i-th, ascending): what the bodies of
// `for &ep in eps` (reader.rs:319-324; CH_EP), `for point in links.iter()` (reader.rs:354-364; CH_NBR) and the
// brute-force loop (reader.rs:697-704; CH_LINEAR) do to `res` and `search_queue`, for all accepted points at once.
// It is applied in two stages (result set, then queue) so that the caller can put memory operations of the NEXT
// chunk between them.
struct ChunkUpdate {
    int mode;
    bool acc;       // this lane's point is accepted: pushed to the search queue
    bool pf;        // ... and passes the candidate filter: may enter the result set
    bool qskip;     // this lane's point was popped straight away by the caller: not pushed (it still enters `res`)
    uint32_t bits, s;
};

__device__ __forceinline__ void heaps_stage_res(Ctx& c, ChunkUpdate& u, int ef) {
    const unsigned resm = __ballot_sync(FULL, u.acc && u.pf);
    // Pushing the keys one by one with `if len == ef { push_pop_max } else { push }` leaves the min(ef, len + m)
    // smallest of the union when len <= ef, and the whole union when len > ef (or for entry points).
    const int m_res = __popc(resm);
    int target = (u.mode == CH_EP || c.res_len > ef) ? c.res_len + m_res : min(ef, c.res_len + m_res);
    if (target > c.res_cap) { c.overflow = true; return; }
    if (m_res) c.res_len = merge_insert<false>(c.res, c.res_len, u.acc && u.pf, ((u64)u.bits << 32) | u.s, target);
}
// The queue is a window [c.que, c.que + c.q_len) of its buffer with c.q_cap cells left from c.que on: entries that can
// never be popped again (is_dead, judged against the UPDATED result set) form a prefix of the descending array and are
// trimmed by moving the window's start, new dead ones are not pushed.  When the window hits the end of the buffer it is
// moved back to the start.
__device__ __forceinline__ void queue_rewind(Ctx& c) {
    const int head = (int)c.p.q_cap - c.q_cap;
    c.que -= head;
    c.q_cap += head;
}
__device__ __forceinline__ void heaps_stage_queue(Ctx& c, const ChunkUpdate& u, int ef) {
    if (u.mode == CH_LINEAR || c.overflow) return;
    const int lane = lane_id();
    const bool prune = (HB_LEAN ? c.trim_ok : (c.p.pass == 0 && !c.p.cancel_after && !c.p.no_trim)) && c.res_len >= ef && c.res_len > 0 && !(c.res[c.res_len - 1] >> 63);
    const uint32_t mb = prune ? (uint32_t)(c.res[c.res_len - 1] >> 32) : 0xffffffffu;
    if (prune && c.q_len > 0 && (uint32_t)(c.que[0] >> 32) > mb) {
        int lo = 1, hi = c.q_len;                  // first entry that is not dead (warp-uniform search)
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((uint32_t)(c.que[mid] >> 32) > mb) lo = mid + 1; else hi = mid;
        }
        c.que += lo; c.q_cap -= lo; c.q_len -= lo;
    }
    const bool qhas = u.acc && !u.qskip && !(prune && !(u.bits >> 31) && u.bits > mb);
    const int mq = __popc(__ballot_sync(FULL, qhas));
    if (c.q_len + mq > c.q_cap) {
        const int head = (int)c.p.q_cap - c.q_cap;
        if (c.q_len + mq > (int)c.p.q_cap) { c.overflow = true; return; }  // would have to drop a live entry
        __syncwarp();
        for (int base = 0; base < c.q_len; base += 32) {                     // a[i] -> a[i - head], bottom up
            const int i = base + lane;
            const u64 v = i < c.q_len ? c.que[i] : 0ull;
            __syncwarp();
            if (i < c.q_len) c.que[i - head] = v;
            __syncwarp();
        }
        queue_rewind(c);
    }
    if (mq) c.q_len = merge_insert<true>(c.que, c.q_len, qhas, ((u64)u.bits << 32) | (uint32_t)(~u.s), c.q_cap);
}

// Visitor::visit — reader.rs:301-369 — and, with `linear`, the candidate loop of brute_force_search
// (reader.rs:683-705).  Entry points: `eps` (n_eps slots in global memory) or `single`.
//
// One loop, one copy of the gather / distance / merge code.  Each iteration handles one chunk of <= 32 points:
// first the entry points (or the linear-scan candidates), then one chunk per expansion — layer 0 reads the
// fixed-stride adjacency line (the line of the most likely next pop is requested ahead), the other layers and
// irregular layer-0 graphs walk the CSR lists.
//
// The heap update of a chunk is DEFERRED into the next iteration (layer 0, pass 0): the node popped next is decided
// first — it is the smallest of (queue top, points the pending chunk accepted), and it is certain to be popped when
// its distance does not exceed a lower bound of the next f_max (the result-set entry that survives whatever the
// pending chunk evicts).  Its adjacency load and its visited-set atomics are then issued around the two merge
// stages, which hide their latency.  When the bound does not settle it, the pending update is applied first and
// the reference's order is followed literally.
template <int KIND>
__device__ __forceinline__ void visit(Ctx& c, const uint32_t* eps, uint32_t n_eps, uint32_t single, uint32_t level, int ef, bool filt, bool linear, bool can_cancel) {
    const DevIndex& ix = c.p.ix;
    const int lane = lane_id();
    const int l01 = level ? 0 : 1;
    const u64 NONE = ~0ull;
    c.res_len = 0;
    c.q_len = 0;
    queue_rewind(c);
#if HB_UPPER_KEEP
    if (KIND == KIND_F32_WARP) c.ring.policy = level ? l2_policy_evict_last() : l2_policy_evict_first();
#endif
    c.cur_dist = c.cur_exp = c.cur_deg = 0;
    const uint32_t* list = linear ? c.p.cand_slots : eps;
    const uint32_t n_first = linear ? c.p.n_cand_slots : (eps ? n_eps : 1u);
    uint32_t base0 = 0;
    const uint32_t* off = ix.off[level];
    const uint32_t* nbr = ix.nbr[level];
    const uint32_t* nbrx = level == 0 ? ix.nbr0x : nullptr;
    const uint32_t xstride = ix.nbr0_stride;                 // 32, or 64: a full first line is followed by the second one
    uint32_t cont_cs = 0;                                    // node whose fixed-stride list is being continued (csr_pos = 32, csr_end = xstride)
    const bool defer = nbrx && c.p.pass == 0 && c.p.defer && !linear && !c.p.no_trim;  // a linear scan pops nothing (reader.rs:683-705)
    const bool chk_neg = !linear && c.p.pass == 0 && !c.p.no_trim;   // trimming is only argued for non-negative distances (is_dead)
    c.trim_ok = c.p.pass == 0 && !c.p.cancel_after && !c.p.no_trim;
    uint32_t spec_cs = 0xffffffffu, spec_adj = 0xffffffffu;  // adjacency line requested ahead of its pop
    uint32_t spec_pf = 0xffffffffu;                          // node whose neighbours' visited-set words were already prefetched
    uint32_t csr_pos = 0, csr_end = 0;                       // rest of the CSR list being expanded
    float f_max = FLT_MAX;
    bool pend = false;
    bool merge_deferred = false;   // the queue half of the pending chunk's merge runs after the next chunk's rows were requested (HB_EARLY_ROWS)
    ChunkUpdate u;
    u.mode = CH_EP; u.acc = u.pf = u.qskip = false; u.bits = u.s = 0;
    for (;;) {
        int mode;
        uint32_t s = 0, old = 0xffffffffu;
        bool valid = false, vis_sent = false;
        PH_DECL
        if (base0 < n_first && !c.overflow) {
            if (pend) { heaps_stage_res(c, u, ef); heaps_stage_queue(c, u, ef); pend = false; if (c.overflow) break; }
            if (linear && can_cancel && cancel_flag_set(c)) { c.cancelled = true; break; }  // reader.rs:684-687, per chunk
            valid = base0 + lane < n_first;
            s = !valid ? 0 : (list ? __ldg(&list[base0 + lane]) : single);
            base0 += 32;
            mode = linear ? CH_LINEAR : CH_EP;
        } else {
            mode = CH_NBR;
            bool have = false;
            if (pend && defer && csr_pos >= csr_end && !(can_cancel && cancel_next(c))) {
                // ---- decide the next pop before the pending update is applied ----
                const u64 qk = u.acc ? (((u64)u.bits << 32) | (uint32_t)(~u.s)) : NONE;
                const u64 cand_new = warp_min_u64(qk);
                const u64 cand_old = c.q_len ? c.que[c.q_len - 1] : NONE;
                const u64 nx = cand_new < cand_old ? cand_new : cand_old;
                if (nx != NONE) {
                    const int m_res = __popc(__ballot_sync(FULL, u.acc && u.pf));
                    const int target = (u.mode == CH_EP || c.res_len > ef) ? c.res_len + m_res : min(ef, c.res_len + m_res);
                    const int idx = target - 1 - m_res;  // the largest old entry that survives the pending evictions
                    float bound = 0.0f;
                    bool known = false;
                    if (target == 0) { bound = FLT_MAX; known = true; }        // result set stays empty: f_max = f32::MAX
                    else if (idx >= 0) { bound = key_dist(c.res[idx]); known = true; }
                    if (known && !(key_dist(nx) > bound)) {
                        const bool from_old = cand_old < cand_new;
                        if (from_old) c.q_len--;
                        u.qskip = !from_old && qk == nx;
                        const uint32_t cs = ~(uint32_t)nx;
                        c.cur_exp += 1;
                        if (can_cancel) ++c.polls;                 // the call that precedes this pop returned false
                        TR(c, TR_POP)
                        if (l01) PH_ADD(c, PH_DECIDE)
                        uint32_t a = (cs == spec_cs) ? spec_adj : __ldg(&nbrx[(size_t)cs * xstride + lane]);
#if HB_EARLY_ROWS
                        // Neither the visited filter nor the row gather of this expansion depends on the heaps, so the two merges of
                        // the pending chunk are spread over the expansion's two memory round trips: the result-set merge runs while the
                        // visited-set atomics are out, the queue merge (below, "merge_deferred") while the rows are being copied.
                        if (xstride > FIXED_DEG && __shfl_sync(FULL, a, 31) != 0xffffffffu) { cont_cs = cs; csr_pos = FIXED_DEG; csr_end = xstride; }
                        valid = a != 0xffffffffu;
                        s = valid ? a : 0;
                        old = vis_issue(c, s, valid);
                        vis_sent = true;
                        heaps_stage_res(c, u, ef);
                        f_max = c.res_len ? key_dist(c.res[c.res_len - 1]) : FLT_MAX;  // what the reference reads at this pop
                        merge_deferred = true;
                        have = true;
#else
                        heaps_stage_res(c, u, ef);                 // ... while the adjacency line is on its way
                        if (xstride > FIXED_DEG && __shfl_sync(FULL, a, 31) != 0xffffffffu) { cont_cs = cs; csr_pos = FIXED_DEG; csr_end = xstride; }
                        valid = a != 0xffffffffu;
                        s = valid ? a : 0;
                        old = vis_issue(c, s, valid);
                        vis_sent = true;
                        heaps_stage_queue(c, u, ef);               // ... while the visited-set atomics are on their way
                        pend = false;
                        if (l01) PH_ADD(c, PH_HEAP)
                        TR(c, TR_HEAP)
                        f_max = c.res_len ? key_dist(c.res[c.res_len - 1]) : FLT_MAX;  // what the reference reads at this pop
                        have = true;
                        if (!c.overflow && c.q_len > 0) {          // the line of the pop after this one
                            spec_cs = ~(uint32_t)c.que[c.q_len - 1];
                            spec_adj = __ldg(&nbrx[(size_t)spec_cs * xstride + lane]);
                        } else {
                            spec_cs = 0xffffffffu;
                        }
#endif
                    }
                }
            }
            if (!have) {
                if (pend) {
                    heaps_stage_res(c, u, ef);
                    heaps_stage_queue(c, u, ef);
                    pend = false;
                    if (l01) PH_ADD(c, PH_HEAP)
                    TR(c, TR_HEAP)
                }
                if (c.overflow || linear) break;
                if (csr_pos >= csr_end) {
                    if (c.q_len == 0) break;
                    if (can_cancel && cancel_poll(c)) { c.cancelled = true; break; }  // reader.rs:330-332
                    u64 top = c.que[c.q_len - 1];
                    float f = key_dist(top);
                    f_max = c.res_len ? key_dist(c.res[c.res_len - 1]) : FLT_MAX;
                    if (f > f_max) break;
                    c.q_len--;
                    uint32_t cs = ~(uint32_t)top;
                    c.cur_exp += 1;
                    TR(c, TR_POP)
                    if (nbrx) {
                        s = (cs == spec_cs) ? spec_adj : __ldg(&nbrx[(size_t)cs * xstride + lane]);
                        if (xstride > FIXED_DEG && __shfl_sync(FULL, s, 31) != 0xffffffffu) { cont_cs = cs; csr_pos = FIXED_DEG; csr_end = xstride; }
                        // the entry now on top of the queue is popped next unless one of cs's neighbours beats it
                        if (c.q_len > 0) {
                            spec_cs = ~(uint32_t)c.que[c.q_len - 1];
                            spec_adj = __ldg(&nbrx[(size_t)spec_cs * xstride + lane]);
                        } else {
                            spec_cs = 0xffffffffu;
                        }
                    } else {
                        csr_pos = __ldg(&off[cs]);
                        csr_end = __ldg(&off[cs + 1]);
                        if (csr_pos >= csr_end) continue;
                        s = csr_pos + lane < csr_end ? __ldg(&nbr[csr_pos + lane]) : 0xffffffffu;
                        csr_pos += 32;
                    }
                } else if (nbrx) {   // the next line of a fixed-stride list longer than 32
                    s = __ldg(&nbrx[(size_t)cont_cs * xstride + csr_pos + lane]);
                    csr_pos += 32;
                } else {
                    s = csr_pos + lane < csr_end ? __ldg(&nbr[csr_pos + lane]) : 0xffffffffu;
                    csr_pos += 32;
                }
                valid = s != 0xffffffffu;
                if (!valid) s = 0;
            }
            if (!HB_LEAN || c.p.out_ctr) c.cur_deg += __popc(__ballot_sync(FULL, valid));
            if (l01) PH_ADD(c, PH_ADJ)
            TR(c, TR_ADJ)
        }
        // ---- visited filter ----
        bool live = valid;
        uint32_t log_at = 0xffffffffu;
        if (mode != CH_LINEAR) {
            if (!vis_sent) old = vis_issue(c, s, valid);
            bool fresh = vis_finish(c, s, valid, old, log_at);   // `path.insert(..)`
            if (mode == CH_NBR) live = fresh;                      // `if !path.insert(point) { continue }`; an ep's result is ignored
        }
        if (c.overflow) { vis_log(c, s, log_at); break; }          // (a deferred merge ran out of room)
        const unsigned lm = __ballot_sync(FULL, live);
        if (l01) PH_ADD(c, PH_VIS)
        TR(c, TR_VIS)
        // ---- gather: start ----
        unsigned own = lm, helpers_used = 0;
        int per = 32;
        RowsInFlight rf;
        if (lm) {
            if (!HB_LEAN || c.p.out_ctr) c.cur_dist += __popc(lm);
            if (KIND == KIND_F32_WARP && c.ts) {
                team_update(c);
                const int H = __popc(*c.team & 0xfu), n_live = __popc(lm);
                // worth it only when the rows do not fit this warp's ring in one go (short rows: one round trip either way)
                if (H && n_live > max(ROW_GROUP, (int)c.ring.slots)) {
                    const int share = H == 1 ? (n_live + 1) >> 1 : (H == 2 ? (n_live + 2) / 3 : (n_live + 3) >> 2);   // ceil(n_live / (H + 1))
                    per = (share + ROW_GROUP - 1) & ~(ROW_GROUP - 1);
                    helpers_used = team_post(c, lm, s, per);
                    own = __ballot_sync(FULL, live && __popc(lm & ((1u << lane) - 1)) < per);
                }
            }
            if (l01) PH_ADD(c, PH_POST)
            rows_begin<KIND>(c, own, s, rf);
            TR(c, TR_POSTED)
        }
        vis_log(c, s, log_at);
        bool bail = false;
        if (merge_deferred) {
            // ---- the queue half of the PREVIOUS chunk's heap update, while this chunk's rows are in flight ----
            if (l01) PH_ADD(c, PH_ROWS)
            heaps_stage_queue(c, u, ef);
            pend = false;
            merge_deferred = false;
            if (l01) PH_ADD(c, PH_HEAP)
            TR(c, TR_HEAP)
            if (!c.overflow && c.q_len > 0) {          // the adjacency line of the pop after this one
                spec_cs = ~(uint32_t)c.que[c.q_len - 1];
                spec_adj = __ldg(&nbrx[(size_t)spec_cs * xstride + lane]);
            } else {
                spec_cs = 0xffffffffu;
            }
            bail = c.overflow;                         // the merge ran out of room: drain the rows already posted, then leave
        }
        if (!lm) { if (bail) break; continue; }
        // ---- gather: finish, distances ----
        float dist = rows_finish<KIND>(c, rf, s);
#if HB_SPEC_VIS
        // the adjacency line requested ahead has landed by now: pull the visited-set words of ITS neighbours into L2, so that
        // the atomics of the next expansion — when it is the one expected — are answered by L2 instead of DRAM
        // (binary kernel: once per expected node — it stays the expected one for as long as freshly accepted points are popped ahead of
        // it; for the f32 kernels repeating the hint or not measures the same with reproducible builds)
        if ((HB_SPEC_VIS_BIN || KIND != KIND_BIN) && nbrx && spec_cs != 0xffffffffu && ((KIND != KIND_BIN && !HB_SPEC_DEDUPE_F32) || !HB_SPEC_DEDUPE || spec_cs != spec_pf)) {
            if (spec_adj != 0xffffffffu) prefetch_l2(&c.vis[spec_adj >> 5]);
            spec_pf = spec_cs;
        }
#endif
        if (l01) PH_ADD(c, PH_ROWS)
        if (KIND == KIND_F32_WARP && helpers_used) dist = team_collect(c, lm, helpers_used, per, dist);
#if HB_TEAM_SHARE
        // dev: three walks and one helper in the CTA: the helper is held for one chunk at a time, so that it goes round the walks
        if (KIND == KIND_F32_WARP && c.ts && (*c.team & 0xfu)) {
            const uint32_t idle = ld_volatile_shared(&c.ts->n_idle);
            if (2 * idle < (uint32_t)SEARCH_WARPS_PER_BLOCK) team_release(*c.ts, *c.team);
        }
#endif
        if (bail) break;
        const uint32_t bits = __float_as_uint(dist);
        if (l01) PH_ADD(c, PH_COLLECT)
        if ((HB_LEAN ? chk_neg : (mode != CH_LINEAR && c.p.pass == 0 && !c.p.no_trim)) && __any_sync(FULL, live && (bits >> 31))) { c.overflow = true; break; }
        // ---- which points are accepted, and which of those may enter the result set (reader.rs:322,353,355-359) ----
        const bool pf = live && passes_filter(c, s, filt);
        bool acc = live;
        if (mode == CH_NBR) {
            // `res.len() < self.ef || dist < f_max` with live len and stale f_max (reader.rs:353): the first
            // ef - len filter-passing points in ascending order are taken unconditionally.
            unsigned pfm = __ballot_sync(FULL, pf);
            int before = __popc(pfm & ((1u << lane) - 1));
            bool fill = (c.res_len + before) < ef;
            acc = live && (fill || dist < f_max);
            // an accepted point that beats the queue's current best is expanded very soon: start pulling its adjacency line
#ifndef HB_NO_ADJ_PREFETCH
#if HB_ADJ_PREFETCH_ALL
            if (nbrx && acc) prefetch_l2(nbrx + (size_t)s * xstride);   // 128 bytes per accepted point against a 3 KB row: every later pop finds its line in L2
#else
            if (nbrx && acc && (c.q_len == 0 || bits <= (uint32_t)(c.que[c.q_len - 1] >> 32))) prefetch_l2(nbrx + (size_t)s * xstride);
#endif
#endif
        }
        u.mode = mode; u.acc = acc; u.pf = pf; u.qskip = false; u.bits = bits; u.s = s;
        pend = true;
        if (l01) PH_ADD(c, PH_ACCEPT)
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
        if (is_f32_warp(KIND)) {
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
        if (is_f32_warp(KIND)) {
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
__device__ void run_query(const SearchParams& p, uint64_t qi, int slot_idx, u64* heap, float* qs, RowRing& ring, TeamShared* ts, uint32_t* team) {
    const DevIndex& ix = p.ix;
    const int lane = lane_id();
    Ctx c(p, ring);
    c.ts = ts; c.team = team;
    c.res = heap; c.res_cap = p.res_cap; c.res_len = 0;
    c.que = heap + p.res_cap; c.q_cap = p.q_cap; c.q_len = 0;
    c.vis = p.visited + (size_t)slot_idx * p.vis_words;
    c.touched = p.touched + (size_t)slot_idx * p.touched_cap;
    c.touched_len = 0; c.touched_over = false;
    c.excl = 0xffffffffu; c.overflow = false;
    c.cancelled = false; c.polls = 0;
    c.tr = false;
#ifdef HB_TRACE
    c.tr = slot_idx == (p.nq < 1000 ? 0 : 777) && p.pass == 0;
#endif
    TR(c, TR_QSTART)
    c.n_dist_up = c.n_exp_up = c.n_deg_up = c.n_dist_l0 = c.n_exp_l0 = c.n_deg_l0 = 0;
    c.cur_dist = c.cur_exp = c.cur_deg = 0;
    u64 flags = p.pass ? HB_FLAG_SLOW_PATH : 0;
    const uint32_t count = p.count;
    const int ef0 = (int)max(p.ef_raw, p.count);

    if ((p.mode & 1) && p.q_slots[qi] == 0xffffffffu) {  // item absent -> Ok(None), reader.rs:826
        if (lane == 0) p.out_len[qi] = 0xffffffffu;
        for (int i = lane; i < (int)count; i += 32) { p.out_ids[qi * count + i] = 0; p.out_dist[qi * count + i] = 0.0f; }
        if (p.out_ctr && lane < HB_N_CTR) p.out_ctr[qi * HB_N_CTR + lane] = 0;
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
    {
        // One call site of visit() drives the whole search as a small state machine:
        //   ST_UPPER  greedy descent, ef = 1, levels max_level..1, shared visited set (reader.rs:732-741)
        //   ST_L0     the ef-bounded layer-0 walk (reader.rs:743-767) / by_item's seeded walk (reader.rs:842-862)
        //   ST_FB     exhaustive fallback, one visit per unseen item (reader.rs:771-795 / 865-889)
        //   ST_LIN    brute_force_search over the candidate slots, ascending (reader.rs:668-711)
        enum { ST_UPPER, ST_L0, ST_FB, ST_LIN };
        uint32_t ep_single = 0, level = 0;
        const uint32_t* eps = ix.eps;
        int st = ST_L0;
        if (p.mode >= 2) {
            st = ST_LIN;
            flags |= HB_FLAG_LINEAR;
        } else if (p.mode & 1) {  // nns_by_item — reader.rs:836-842
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
            else if (st == ST_L0) { ef = ef0; level = 0; TR(c, TR_L0) }
            else if (st == ST_LIN) { ef = (int)count; level = 0; filt = false; }
            visit<KIND>(c, eps, ix.n_ep, ep_single, level, ef, filt, st == ST_LIN,
                        st != ST_UPPER && (p.cancel_flag != nullptr || p.cancel_after != 0));
            if (c.overflow || c.cancelled || st == ST_LIN) break;
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
        if (c.cancelled && !c.overflow) {
            // return_if_cancelled! (reader.rs:749-764,844-859): only the interrupted visit's heap, ascending, take(count)
            flags |= HB_FLAG_CANCELLED;
            c.res_cap = base_cap - (int)(c.res - base_res);
        } else {
            c.res = base_res; c.res_cap = base_cap;
            if (first >= 0) {
                c.res_len = acc_len;
                if (!c.overflow) sort_tail(c.res, first, acc_len);
            }
        }
        if (p.linear_cancelled) flags |= HB_FLAG_CANCELLED;
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
    for (int i = lane; i < (int)count; i += 32) {  // entries past out_len are zeroed: the output is a function of the inputs
        u64 k = i < n_out ? c.res[i] : 0;
        p.out_ids[qi * count + i] = i < n_out ? __ldg(&ix.ids[(uint32_t)k]) : 0u;
        p.out_dist[qi * count + i] = i < n_out ? key_dist(k) : 0.0f;
    }
    if (p.n_peers) {
        // fused all-gather: this shard's padded top-k of query qi goes straight into every peer's gather buffer
        for (int pr = 0; pr < p.n_peers; ++pr) {
            uint32_t* gi = p.peer_ids[pr] + ((size_t)p.shard_rank * p.nq + qi) * count;
            float* gd = p.peer_dist[pr] + ((size_t)p.shard_rank * p.nq + qi) * count;
            for (int i = lane; i < (int)count; i += 32) {
                u64 k = i < n_out ? c.res[i] : 0;
                gi[i] = i < n_out ? __ldg(&ix.ids[(uint32_t)k]) : 0xffffffffu;
                gd[i] = i < n_out ? key_dist(k) : __int_as_float(0x7f800000);
            }
        }
    }
    if (lane == 0) {
        p.out_len[qi] = (uint32_t)n_out | ((flags & HB_FLAG_CANCELLED) ? HB_LEN_CANCELLED : 0u);
        if (p.out_ctr) {
            uint64_t* o = p.out_ctr + qi * HB_N_CTR;
            o[HB_CTR_DIST_UPPER] = c.n_dist_up; o[HB_CTR_DIST_L0] = c.n_dist_l0;
            o[HB_CTR_EXP_UPPER] = c.n_exp_up; o[HB_CTR_EXP_L0] = c.n_exp_l0;
            o[HB_CTR_DEG_UPPER] = c.n_deg_up; o[HB_CTR_DEG_L0] = c.n_deg_l0;
            o[HB_CTR_FLAGS] = flags; o[HB_CTR_RESERVED] = 0;
        }
    }
    __syncwarp();
    TR(c, TR_QEND)
#ifdef HB_PHASES
    PH_ADD(c, PH_TAIL)
    c.ph[PH_TOTAL] = clock64() - ph_q0;
    if (lane == 0)
        for (int i = 0; i < PH_N; ++i) atomicAdd(&g_phase[i], (unsigned long long)c.ph[i]);
#endif
}

// Shared memory of one warp: [row ring][ring barriers][query][heaps (pass 0 only)].
template <int KIND, int MIN_BLOCKS>
__global__ void __launch_bounds__(SEARCH_WARPS_PER_BLOCK * 32, MIN_BLOCKS) hnsw_search_kernel(const __grid_constant__ SearchParams p) {
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
#if HB_OPAQUE_ADDR
    // dev: per-warp addresses are pure functions of threadIdx and the parameters, so under register pressure the compiler may
    // re-derive them at every use; made opaque, they are kept in registers or spilled once
    asm volatile("" : "+l"(qs), "+l"(heap), "+l"(ring.ptr), "+r"(ring.data), "+r"(ring.bars));
#endif
    if (p.ring_slots) {
        if (lane_id() == 0) {
            for (uint32_t i = 0; i < p.ring_slots; ++i) mbar_init(ring.bars + i * 8, 1);
            mbar_fence_init();
        }
        __syncwarp();
    }
    // helper state of the CTA (f32 ring kernel only)
    constexpr bool TEAM = KIND == KIND_F32_WARP;
    __shared__ TeamShared team_shared;
    TeamShared* ts = nullptr;
    uint32_t team = 0;
    if (TEAM && p.team) {
        ts = &team_shared;
        if (threadIdx.x < SEARCH_WARPS_PER_BLOCK) {
            mbar_init(smem_addr(&ts->job_bar[threadIdx.x]), 1);
            mbar_init(smem_addr(&ts->done_bar[threadIdx.x]), 1);
            ts->owner[threadIdx.x] = TEAM_BUSY;
            ts->done_parity[threadIdx.x] = 0;
            if (threadIdx.x == 0) ts->n_idle = 0;
            mbar_fence_init();
        }
        __syncthreads();
    }
    const uint32_t n_work = p.pass == 0 ? p.n_work : *p.n_overflow;
    // the first n_static queries are handed out by position — warp w of CTA b starts with query w * gridDim + b, so a
    // batch smaller than the grid spreads over all CTAs (and leaves every CTA helpers) — the rest by the counter
    unsigned long long w = (unsigned long long)warp_in_block * gridDim.x + blockIdx.x;
    bool by_position = w < p.n_static;
    for (;;) {
        if (!by_position) {
            if (lane_id() == 0) w = p.n_static + atomicAdd(p.work_counter + p.pass, 1ull);
            w = __shfl_sync(FULL, w, 0);
        }
        by_position = false;
        if (w >= n_work) break;
        uint64_t qi = p.pass == 0 ? w : p.overflow_list[w];
        run_query<KIND>(p, qi, slot_idx, heap, qs, ring, ts, &team);
    }
    if (TEAM && ts) {
        team_release(*ts, team);
        if (lane_id() == 0) {
            *reinterpret_cast<volatile uint32_t*>(&ts->owner[warp_in_block]) = TEAM_FREE;
            atomicAdd(&ts->n_idle, 1u);
        }
        __syncwarp();
        team_help<KIND>(p, *ts, ring);
    }
}

// ---- device graph builder: candidate search --------------------------------------------------------------------------
// One warp per item of the batch, the same visit() as the reader's walk: `walk_layer` (hnsw.rs:460-519) differs from
// Visitor::visit only in that every call starts from an empty visited set and has no candidates filter.
template <int KIND>
__global__ void __launch_bounds__(SEARCH_WARPS_PER_BLOCK * 32, KIND == KIND_F32_WARP ? HB_MIN_BLOCKS_F32 : (KIND == KIND_F32_DIRECT ? HB_MIN_BLOCKS_DIRECT : 4)) build_search_kernel(const __grid_constant__ BuildSearchParams bp) {
    extern __shared__ __align__(128) unsigned char smem[];
    const SearchParams& p = bp.sp;
    const int warp_in_block = threadIdx.x >> 5;
    const int warps_per_block = blockDim.x >> 5;
    const int slot_idx = blockIdx.x * warps_per_block + warp_in_block;
    const size_t ring_bytes = (size_t)p.ring_slots * p.ring_stride;
    const size_t bar_bytes = ((size_t)p.ring_slots * 8 + 15) & ~(size_t)15;
    const size_t heap_bytes = (size_t)(p.res_cap + p.q_cap) * 8;
    const size_t per_warp = (ring_bytes + bar_bytes + p.q_smem_bytes + heap_bytes + 127) & ~(size_t)127;
    unsigned char* base = smem + per_warp * warp_in_block;
    float* qs = reinterpret_cast<float*>(base + ring_bytes + bar_bytes);
    u64* heap = reinterpret_cast<u64*>(base + ring_bytes + bar_bytes + p.q_smem_bytes);
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
    const DevIndex& ix = p.ix;
    const int lane = lane_id();
    for (;;) {
        unsigned long long w = 0;
        if (lane == 0) w = atomicAdd(p.work_counter, 1ull);
        w = __shfl_sync(FULL, w, 0);
        if (w >= bp.n_items) break;
        Ctx c(p, ring);
        c.res = heap; c.res_cap = p.res_cap; c.res_len = 0;
        c.que = heap + p.res_cap; c.q_cap = p.q_cap; c.q_len = 0;
        c.vis = p.visited + (size_t)slot_idx * p.vis_words;
        c.touched = p.touched + (size_t)slot_idx * p.touched_cap;
        c.touched_len = 0; c.touched_over = false;
        c.excl = 0xffffffffu; c.overflow = false;
        c.cancelled = false; c.polls = 0;
        c.tr = false;
        c.n_dist_up = c.n_exp_up = c.n_deg_up = c.n_dist_l0 = c.n_exp_l0 = c.n_deg_l0 = 0;
        c.cur_dist = c.cur_exp = c.cur_deg = 0;
#ifdef HB_PHASES
        for (int i = 0; i < PH_N; ++i) c.ph[i] = 0;
#endif
        stage_query<KIND>(c, qs, w);  // the item's own row (sp.mode = by_item, sp.q_slots = the batch)
        const uint32_t* eps = nullptr;
        uint32_t n_eps = 0, single = 0;
        if (bp.descend) {
            eps = ix.eps; n_eps = ix.n_ep;
            for (uint32_t lvl = ix.max_level; lvl > bp.level && !c.overflow; --lvl) {  // hnsw.rs:304-309
                visit<KIND>(c, eps, n_eps, single, lvl, 1, false, false, false);
                if (c.res_len) { single = (uint32_t)c.res[0]; eps = nullptr; }
                __syncwarp();
                vis_clear(c);                                                          // walk_layer starts a fresh `visited`
            }
        } else {
            eps = bp.eps_in + (size_t)w * bp.eps_stride;
            uint32_t e = lane < (int)bp.eps_stride ? eps[lane] : 0xffffffffu;          // eps_stride <= 32
            n_eps = __popc(__ballot_sync(FULL, e != 0xffffffffu));                     // valid entries come first
        }
        int n_out = 0;
        if (eps == nullptr || n_eps > 0) {
            c.overflow = false;
            visit<KIND>(c, eps, n_eps, single, bp.level, (int)bp.efc, false, false, false);  // hnsw.rs:313-316
            n_out = min(c.res_len, (int)bp.efc);  // an overflowing walk (cannot happen with these capacities) keeps what it has
        }
        __syncwarp();
        for (int i = lane; i < n_out; i += 32) bp.cand[(size_t)w * bp.efc + i] = c.res[i];
        if (lane == 0) {
            bp.cand_len[w] = (uint32_t)n_out;
            if (c.overflow && bp.n_cut) atomicAdd(bp.n_cut, 1ull);
        }
        vis_clear(c);
        __syncwarp();
    }
}

typedef void (*build_search_kernel_t)(const BuildSearchParams);
static build_search_kernel_t build_kernel_for(int kind) {
    switch (kind) {
        case KIND_F32_WARP: return build_search_kernel<KIND_F32_WARP>;
        case KIND_F32_DIRECT: return build_search_kernel<KIND_F32_DIRECT>;
        case KIND_F32_LANE: return build_search_kernel<KIND_F32_LANE>;
        default: return build_search_kernel<KIND_BIN>;
    }
}
static int build_variant_of(const SearchParams& p) { return (p.ix.kind == KIND_F32_WARP && p.ring_slots == 0) ? KIND_F32_DIRECT : p.ix.kind; }
static void set_build_kernel_attrs() {
    static bool attr_set = false;
    if (!attr_set) {
        for (int k = 0; k < 4; ++k) cudaFuncSetAttribute(build_kernel_for(k), cudaFuncAttributeMaxDynamicSharedMemorySize, SEARCH_MAX_SMEM);
        attr_set = true;
    }
}
int build_search_blocks_per_sm(const SearchParams& p) {
    set_build_kernel_attrs();
    int nb = 0;
    SearchParams q = p;
    q.pass = 0;
    size_t smem = search_smem_per_warp(q) * SEARCH_WARPS_PER_BLOCK;
    if (smem > (size_t)SEARCH_MAX_SMEM) return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, build_kernel_for(build_variant_of(p)), SEARCH_WARPS_PER_BLOCK * 32, smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return nb;
}
hb_status launch_build_search(const BuildSearchParams& bp, int blocks, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    set_build_kernel_attrs();
    SearchParams q = bp.sp;
    q.pass = 0;
    size_t smem = search_smem_per_warp(q) * SEARCH_WARPS_PER_BLOCK;
    cudaMemsetAsync(bp.sp.work_counter, 0, sizeof(unsigned long long), stream);
    uint32_t need = (bp.n_items + SEARCH_WARPS_PER_BLOCK - 1) / SEARCH_WARPS_PER_BLOCK;
    if ((uint32_t)blocks > need) blocks = (int)need;
    if (blocks < 1) blocks = 1;
    build_kernel_for(build_variant_of(bp.sp))<<<blocks, SEARCH_WARPS_PER_BLOCK * 32, smem, stream>>>(bp);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("build search launch failed: %s", cudaGetErrorString(e)); return HB_ECUDA; }
    return HB_OK;
}

std::atomic<unsigned long long> g_launches{0};

// dev: read and reset the phase timers (all zero unless built with -DHB_PHASES)
void read_phases(unsigned long long* out) {
    cudaMemcpyFromSymbol(out, g_phase, sizeof(unsigned long long) * PH_N);
    unsigned long long z[16] = {};
    cudaMemcpyToSymbol(g_phase, z, sizeof(z));
}

// dev: copy out and reset the event trace (empty unless built with -DHB_TRACE)
uint32_t read_trace(unsigned long long* out, uint32_t cap) {
#ifdef HB_TRACE
    unsigned int n = 0;
    cudaMemcpyFromSymbol(&n, g_trace_n, sizeof(n));
    if (n > cap) n = cap;
    if (n) cudaMemcpyFromSymbol(out, g_trace, sizeof(unsigned long long) * n);
    unsigned int z = 0;
    cudaMemcpyToSymbol(g_trace_n, &z, sizeof(z));
    return n;
#else
    (void)out; (void)cap;
    return 0;
#endif
}

// ---- peer synchronisation of the fused all-gather -----------------------------------------------------------
// Runs behind the search kernel on the same stream: publish "shard `shard_rank` has delivered epoch `epoch`" in every
// peer's flag array (system-scope release after the search kernel's peer stores), then wait until every shard has
// published the same epoch here.  Bounded spin: a peer that never arrives makes the wait give up (flag 0xdead in
// my_flags[n_peers]) instead of hanging the device.
__global__ void peer_signal_wait_kernel(uint32_t* const* peer_flags, int n_peers, int shard_rank, uint32_t* my_flags, uint32_t epoch) {
    const int t = threadIdx.x;
    if (t < n_peers) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_flags[t] + shard_rank), "r"(epoch) : "memory");
    }
    __syncthreads();
    if (t < n_peers) {
        uint32_t v = 0;
        unsigned long long spins = 0;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(my_flags + t) : "memory");
            if ((int)(v - epoch) >= 0) break;
            __nanosleep(200);
        } while (++spins < 20000000ull);  // ~ 4 s
        if ((int)(v - epoch) < 0) my_flags[n_peers] = 0xdead;
    }
}

hb_status launch_peer_signal_wait(uint32_t* const* peer_flags, int n_peers, int shard_rank, uint32_t* my_flags, uint32_t epoch, void* stream) {
    peer_signal_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(peer_flags, n_peers, shard_rank, my_flags, epoch);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("peer sync launch failed: %s", cudaGetErrorString(e)); return HB_ECUDA; }
    return HB_OK;
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
// The kernel variant a call runs: the three row kinds, f32 warp rows without a ring (gathered directly), the f32 ring
// kernel compiled for one more resident CTA per SM (fewer registers) for rows whose rings leave room for it, and the binary
// kernel compiled for four CTAs per SM (128 registers, nothing spilled) for heaps so large that shared memory holds no more.
constexpr int VAR_F32_RING_SHORT = 4, VAR_BIN_WIDE = 5, N_VARIANTS = 6;
static int variant_of(const SearchParams& p) {
    if (p.ix.kind == KIND_BIN && p.bin_wide) return VAR_BIN_WIDE;
    if (p.ix.kind != KIND_F32_WARP) return p.ix.kind;
    if (p.ring_slots == 0) return KIND_F32_DIRECT;
    return p.ring_short ? VAR_F32_RING_SHORT : KIND_F32_WARP;
}
static search_kernel_t kernel_for(int variant) {
    switch (variant) {
        case KIND_F32_WARP: return hnsw_search_kernel<KIND_F32_WARP, HB_MIN_BLOCKS_F32>;
        case VAR_F32_RING_SHORT: return hnsw_search_kernel<KIND_F32_WARP, HB_MIN_BLOCKS_F32_SHORT>;
        case KIND_F32_DIRECT: return hnsw_search_kernel<KIND_F32_DIRECT, HB_MIN_BLOCKS_DIRECT>;
        case KIND_F32_LANE: return hnsw_search_kernel<KIND_F32_LANE, 4>;
        case VAR_BIN_WIDE: return hnsw_search_kernel<KIND_BIN, 4>;
        default: return hnsw_search_kernel<KIND_BIN, HB_MIN_BLOCKS_BIN>;
    }
}
static void set_kernel_attrs() {
    static bool attr_set = false;
    if (!attr_set) {
        for (int k = 0; k < N_VARIANTS; ++k)
            if (cudaFuncSetAttribute(kernel_for(k), cudaFuncAttributeMaxDynamicSharedMemorySize, SEARCH_MAX_SMEM) != cudaSuccess) {
                fprintf(stderr, "[hb] cudaFuncSetAttribute(MaxDynamicSharedMemorySize = %d) failed for search kernel %d: %s\n", SEARCH_MAX_SMEM, k, cudaGetErrorString(cudaGetLastError()));
            }
        attr_set = true;
    }
}

int search_blocks_per_sm(const SearchParams& p) {
    set_kernel_attrs();
    int nb = 0;
    size_t smem = search_smem_per_warp(p) * SEARCH_WARPS_PER_BLOCK;
    if (smem > (size_t)SEARCH_MAX_SMEM) return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel_for(variant_of(p)), SEARCH_WARPS_PER_BLOCK * 32, smem) != cudaSuccess) {
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
        // a batch smaller than the resident warps: one CTA per query while CTAs last (its other warps become helpers)
        const bool team = fast.ix.kind == KIND_F32_WARP && fast.ring_slots != 0 && fast.team;
        uint64_t need = team ? (uint64_t)fast.n_work : ((uint64_t)fast.n_work + wpb - 1) / wpb;
        int blocks = (uint64_t)blocks_fast > need ? (int)need : blocks_fast;
        if (blocks < 1) blocks = 1;
        SearchParams fp = fast;
        fp.n_static = (uint32_t)std::min<uint64_t>(fast.n_work, (uint64_t)blocks * wpb);
        static const bool dbg = getenv("HB_DEBUG_LAUNCH") != nullptr;
        if (dbg) fprintf(stderr, "[hb] search launch: kind %d, %d CTAs x %d threads, %zu B shared memory per CTA, res_cap %u q_cap %u\n", variant_of(fast), blocks, wpb * 32, smem, fast.res_cap, fast.q_cap);
        kernel_for(variant_of(fast))<<<blocks, wpb * 32, smem, stream>>>(fp);
        ++g_launches;
    } else {
        uint32_t n = fast.n_work;
        fill_iota_kernel<<<(n + 255) / 256, 256, 0, stream>>>(fast.overflow_list, fast.n_overflow, n);
        ++g_launches;
    }
    size_t smem = search_smem_per_warp(slow) * wpb;
    kernel_for(variant_of(slow))<<<blocks_slow, wpb * 32, smem, stream>>>(slow);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("search launch failed: %s", cudaGetErrorString(e)); return HB_ECUDA; }
    return HB_OK;
}

}  // namespace hb
