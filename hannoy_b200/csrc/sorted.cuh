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

}  // namespace hb
