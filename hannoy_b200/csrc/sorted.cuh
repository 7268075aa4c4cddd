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
// Merge up to 32 new keys (lane L contributes `key` iff `has`) into the sorted array a[0..len), in place, in ONE pass,
// for arrays of any length in shared or global memory.  Every lane binary-searches the slot of its own key and ranks it
// among the new keys: old position + rank is the key's FINAL position, and the final positions of the new keys are a
// bit mask F over the new array.  The old entries then fill the cells F leaves free, in order: cell j of the new
// array takes old entry j - |{new keys placed below j}| — one popcount per cell instead of one comparison per (cell, new
// key) pair.  Cells are moved in blocks of MERGE_BLOCK tiles of 32, from the top block down (an entry only ever moves up):
// the F words of a block come from MERGE_BLOCK independent warp reductions, then all loads of the block are issued, then
// all its stores — one shared-memory round trip per block (2 tiles: measured best on all three workloads against 1, 3, 4 and 8).  Tiles below the
// first insertion point are not touched.  Equivalent to inserting the keys one after the other (all keys are distinct).
// (A fast path for tiles that hold no new key — they move by a constant — was measured and is not worth its test: +-0 at ef = 800,
// -2 % on short heaps.)
//   DESC   array is descending (search queue, next pop at the end) instead of ascending (result set)
//   keep   final length is capped to `keep` (the entries that sort last are dropped)
// Returns the new length.
#ifndef HB_MERGE_BLOCK
#define HB_MERGE_BLOCK 2
#endif
constexpr int MERGE_BLOCK = HB_MERGE_BLOCK;
template <bool DESC>
__device__ __forceinline__ int merge_insert(u64* a, int len, bool has, u64 key, int keep) {
    const int lane = lane_id();
    const unsigned hm = __ballot_sync(FULL, has);
    const int new_len = min(len + __popc(hm), keep);
    if (!hm || new_len <= 0) return max(new_len, 0);
    unsigned fp = 0xffffffffu;                      // final position of this lane's key
    if (has) {
        int lo = 0, hi = len;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const u64 x = a[mid];
            const bool before = DESC ? (x > key) : (x < key);
            if (before) lo = mid + 1; else hi = mid;
        }
        fp = (unsigned)lo;
    }
    {
        int rank = 0;
        for (unsigned m = hm; m; m &= m - 1) {
            const u64 kb = __shfl_sync(FULL, key, __ffs(m) - 1);
            rank += DESC ? (kb > key) : (kb < key);
        }
        if (has) fp += (unsigned)rank;
    }
    const int t0 = (int)(__reduce_min_sync(FULL, fp) >> 5);
    const int t_top = (new_len - 1) >> 5;
    const unsigned lt = (1u << lane) - 1u;
    for (int tb = t_top - (t_top - t0) % MERGE_BLOCK; tb >= t0; tb -= MERGE_BLOCK) {   // blocks [tb, tb + MERGE_BLOCK), top down; the lowest starts at t0
        int below = __popc(__ballot_sync(FULL, fp < (unsigned)(tb << 5)));             // new keys placed under the block
        u64 v[MERGE_BLOCK];
        unsigned moved = 0;
#pragma unroll
        for (int i = 0; i < MERGE_BLOCK; ++i) {
            v[i] = 0ull;
            if (tb + i <= t_top) {                  // warp-uniform: only the top block can be short
                const unsigned f = __reduce_or_sync(FULL, (fp >> 5) == (unsigned)(tb + i) ? 1u << (fp & 31u) : 0u);
                const int j = ((tb + i) << 5) + lane;
                const int src = j - below - __popc(f & lt);
                below += __popc(f);
                const bool mv = j < new_len && !((f >> lane) & 1u) && src != j;
                if (mv) v[i] = a[src];
                moved |= (unsigned)mv << i;
            }
        }
        __syncwarp();                               // every lane holds its entries of the block: the cells may be overwritten
#pragma unroll
        for (int i = 0; i < MERGE_BLOCK; ++i)
            if ((moved >> i) & 1u) a[((tb + i) << 5) + lane] = v[i];
        __syncwarp();
    }
    if (fp < (unsigned)new_len) a[fp] = key;
    __syncwarp();
    return new_len;
}

}  // namespace hb
