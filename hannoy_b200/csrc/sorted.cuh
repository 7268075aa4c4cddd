// sorted.cuh — warp-cooperative sorted arrays of 64-bit keys (shared or global memory).
#pragma once
#include "dist.cuh"

namespace hb {

typedef unsigned long long u64;

// ---- sorted arrays ----------------------------------------------------------------------------------
__device__ __forceinline__ int count_lt(const u64* a, int n, u64 key) {
    int c = 0;
    for (int i = lane_id(); i < n; i += 32) c += a[i] < key;
    return __reduce_add_sync(FULL, c);
}
__device__ __forceinline__ int count_gt(const u64* a, int n, u64 key) {
    int c = 0;
    for (int i = lane_id(); i < n; i += 32) c += a[i] > key;
    return __reduce_add_sync(FULL, c);
}
// a[pos..n) -> a[pos+1..n+1), a[pos] = key
__device__ __forceinline__ void insert_at(u64* a, int n, int pos, u64 key) {
    const int lane = lane_id();
    int hi = n;
    while (hi > pos) {
        int lo = max(pos, hi - 32);
        int i = lo + lane;
        u64 v = (i < hi) ? a[i] : 0;
        __syncwarp();
        if (i < hi) a[i + 1] = v;
        __syncwarp();
        hi = lo;
    }
    if (lane == 0) a[pos] = key;
    __syncwarp();
}
// drop a[0]: a[1..pos) -> a[0..pos-1), a[pos-1] = key   (pos >= 1)
__device__ __forceinline__ void insert_drop_front(u64* a, int pos, u64 key) {
    const int lane = lane_id();
    int lo = 1;
    while (lo < pos) {
        int hi = min(pos, lo + 32);
        int i = lo + lane;
        u64 v = (i < hi) ? a[i] : 0;
        __syncwarp();
        if (i < hi) a[i - 1] = v;
        __syncwarp();
        lo = hi;
    }
    if (lane == 0) a[pos - 1] = key;
    __syncwarp();
}


// keep the `cap` smallest keys: `if len == cap { push_pop_max } else { push }`
__device__ __forceinline__ void topk_insert(u64* a, int& len, int cap, u64 key) {
    if (len == cap) {
        if (cap == 0 || key > a[len - 1]) return;
        int pos = count_lt(a, len - 1, key);
        insert_at(a, len - 1, pos, key);
    } else {
        int pos = count_lt(a, len, key);
        insert_at(a, len, pos, key);
        len++;
    }
}

// ---- batched merge -------------------------------------------------------------------------------------
// Merge up to 32 new keys (lane L contributes `key` iff `has`) into the sorted array a[0..len) in ONE pass:
// every lane binary-searches the slot of its own key, the old entries are pulled into registers, each
// entry's displacement is the number of new keys that sort before it, then everything is written back.
// Equivalent to inserting the keys one after the other (all keys are distinct), at the cost of one insertion.
//   DESC       array is descending (search queue, next pop at the end) instead of ascending (result set)
//   drop_above (TRIM only) old entries whose distance bits exceed it are dropped: they sit at the front of a
//              descending array
//   keep       final length is capped to `keep` (the entries that sort last are dropped)
// Needs len <= 32 * MAX_TILES.  Returns the new length.
#ifdef HB_MERGE_NOINLINE
#define HB_MERGE_INLINE __noinline__
#else
#define HB_MERGE_INLINE __forceinline__
#endif
template <bool DESC, bool TRIM, int MAX_TILES>
__device__ HB_MERGE_INLINE int merge_batch(u64* a, int len, bool has, u64 key, int keep, uint32_t drop_above) {
    const int lane = lane_id();
    const unsigned hm = __ballot_sync(FULL, has);
    int pos = 0x7fffffff;
    if (has) {
        int lo = 0, hi = len;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            u64 x = a[mid];
            bool before = DESC ? (x > key) : (x < key);
            if (before) lo = mid + 1; else hi = mid;
        }
        pos = lo;
    }
    // Entries below the smallest insertion point do not move (unless the front is trimmed): their tiles are neither
    // loaded nor stored.  Dead entries form a prefix of a descending array, so a[0] tells whether there is any.
    const bool trim = TRIM && len > 0 && (uint32_t)(a[0] >> 32) > drop_above;
    const int t0 = trim ? 0 : (int)(__reduce_min_sync(FULL, (unsigned)pos) >> 5);
    u64 v[MAX_TILES];
    int sh[MAX_TILES];
    int d = 0;
#pragma unroll
    for (int t = 0; t < MAX_TILES; ++t) {
        int i = t * 32 + lane;
        v[t] = 0ull;
        sh[t] = 0;
        if (t >= t0 && t * 32 < len) {  // warp-uniform
            if (i < len) v[t] = a[i];
            if (TRIM) { if (trim) d += __popc(__ballot_sync(FULL, i < len && (uint32_t)(v[t] >> 32) > drop_above)); }
        }
    }
    if (!has) pos = 0;
    int rank = 0;
    for (unsigned m = hm; m; m &= m - 1) {
        int src = __ffs(m) - 1;
        u64 kb = __shfl_sync(FULL, key, src);
        int pb = __shfl_sync(FULL, pos, src);
        rank += DESC ? (kb > key) : (kb < key);
#pragma unroll
        for (int t = 0; t < MAX_TILES; ++t)
            if (t >= t0 && t * 32 < len) sh[t] += (pb <= t * 32 + lane);
    }
    __syncwarp();
    const int new_len = min(len + __popc(hm) - d, keep);
#pragma unroll
    for (int t = 0; t < MAX_TILES; ++t) {
        int i = t * 32 + lane;
        int j = i + sh[t] - d;
        if (t >= t0 && t * 32 < len && i < len && j >= 0 && j < new_len) a[j] = v[t];
    }
    if (has) {
        int j = pos + rank - d;
        if (j >= 0 && j < new_len) a[j] = key;
    }
    __syncwarp();
    return new_len;
}

// The same merge for arrays of any length (heaps beyond 32 * MAX_TILES entries: large ef, the global-memory pass): the
// old entries are moved tile by tile from the top down, each by the number of new keys that sort before it; tiles below
// the first insertion point are not touched.  A dead prefix (TRIM) is compacted away first, bottom up.
template <bool DESC, bool TRIM>
__device__ __noinline__ int merge_batch_large(u64* a, int len, bool has, u64 key, int keep, uint32_t drop_above) {
    const int lane = lane_id();
    if (TRIM && len > 0 && (uint32_t)(a[0] >> 32) > drop_above) {
        int d = 0;
        for (int i = lane; i < len; i += 32) d += (uint32_t)(a[i] >> 32) > drop_above;
        d = __reduce_add_sync(FULL, d);
        for (int base = d; base < len; base += 32) {  // a[i] -> a[i - d], bottom up
            const int i = base + lane;
            const u64 v = i < len ? a[i] : 0ull;
            __syncwarp();
            if (i < len) a[i - d] = v;
            __syncwarp();
        }
        len -= d;
    }
    const unsigned hm = __ballot_sync(FULL, has);
    if (!hm) return min(len, keep);
    int pos = 0x7fffffff;
    if (has) {
        int lo = 0, hi = len;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const u64 x = a[mid];
            const bool before = DESC ? (x > key) : (x < key);
            if (before) lo = mid + 1; else hi = mid;
        }
        pos = lo;
    }
    const int t0 = (int)(__reduce_min_sync(FULL, (unsigned)pos) >> 5);
    if (!has) pos = 0x7fffffff;
    int rank = 0;
    for (unsigned m = hm; m; m &= m - 1) {
        const u64 kb = __shfl_sync(FULL, key, __ffs(m) - 1);
        rank += DESC ? (kb > key) : (kb < key);
    }
    const int new_len = min(len + __popc(hm), keep);
    __syncwarp();
    for (int t = (len - 1) >> 5; t >= t0 && len > 0; --t) {
        const int i = t * 32 + lane;
        const u64 v = i < len ? a[i] : 0ull;
        int sh = 0;
        for (unsigned m = hm; m; m &= m - 1) sh += __shfl_sync(FULL, pos, __ffs(m) - 1) <= i;
        __syncwarp();  // the tile is in registers: its cells (and the ones above, already moved) may be overwritten
        if (i < len && i + sh < new_len) a[i + sh] = v;
        __syncwarp();
    }
    if (has) {
        const int j = pos + rank;
        if (j < new_len) a[j] = key;
    }
    __syncwarp();
    return new_len;
}

// merge of up to 32 keys into a sorted array of any length
template <bool DESC, bool TRIM, int MAX_TILES>
__device__ __forceinline__ int merge_any(u64* a, int len, bool has, u64 key, int keep, uint32_t drop_above) {
    if (len <= 32 * MAX_TILES) return merge_batch<DESC, TRIM, MAX_TILES>(a, len, has, key, keep, drop_above);
    return merge_batch_large<DESC, TRIM>(a, len, has, key, keep, drop_above);
}

}  // namespace hb
