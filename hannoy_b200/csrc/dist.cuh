// dist.cuh — device distance routines.  Compiled with -fmad=false: every fused multiply-add is an
// explicit fmaf(), every separately-rounded op an explicit __f*_rn, so the results are bit-identical
// to the reference's x86-64 AVX+FMA / SSE / scalar code paths (src/spaces/*.rs, src/distance/*.rs).
#pragma once
#include <cuda_runtime.h>

#include "common.h"

namespace hb {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// hsum256_ps_avx per 8-lane group, then ((h1+h2)+h3)+h4 — src/spaces/simple_avx.rs:8-13,55-58.
// Lane j of the warp plays accumulator j/8, SIMD lane j%8.
__device__ __forceinline__ float warp_hsum_avx(float acc) {
    acc = __fadd_rn(acc, __shfl_xor_sync(FULL, acc, 4));
    acc = __fadd_rn(acc, __shfl_xor_sync(FULL, acc, 2));
    acc = __fadd_rn(acc, __shfl_xor_sync(FULL, acc, 1));
    float h1 = __shfl_sync(FULL, acc, 0), h2 = __shfl_sync(FULL, acc, 8);
    float h3 = __shfl_sync(FULL, acc, 16), h4 = __shfl_sync(FULL, acc, 24);
    return __fadd_rn(__fadd_rn(__fadd_rn(h1, h2), h3), h4);
}

// D::distance epilogue for the f32 metrics.  `raw` = euclid sum or dot product.
__device__ __forceinline__ float finish_f32(int metric, float raw, float qn, float in) {
    if (metric == HB_COSINE) {  // src/distance/cosine.rs:40-56 (p = query, q = item)
        float pnqn = __fmul_rn(qn, in);
        if (pnqn > 1.1920929e-07f) {
            float c = __fdiv_rn(raw, pnqn);
            c = c < -1.0f ? -1.0f : c;
            c = c > 1.0f ? 1.0f : c;
            return __fdiv_rn(__fsub_rn(1.0f, c), 2.0f);
        }
        return 0.0f;
    }
    return raw;
}
// D::distance epilogue for the popcount metrics.  h = popcount(u ^ v), L = padded bit length.
__device__ __forceinline__ float finish_bin(int metric, uint32_t h, uint32_t L, float qn, float in) {
    switch (metric) {
        case HB_HAMMING:  // src/distance/hamming.rs:44-47
            return __fdiv_rn((float)h, (float)L);
        case HB_BQ_COSINE: {  // binary_quantized_cosine.rs:44-59, simple.rs:119-131
            float pq = (float)((int)L - 2 * (int)h);
            float pnqn = __fmul_rn(qn, in);
            if (pnqn != 0.0f) return __fdiv_rn(__fsub_rn(1.0f, __fdiv_rn(pq, pnqn)), 2.0f);
            return 0.0f;
        }
        case HB_BQ_EUCLIDEAN:  // binary_quantized_euclidean.rs:76-83
            return (float)(h * 4u);
        default:  // HB_BQ_MANHATTAN — binary_quantized_manhattan.rs:72-79
            return (float)(h * 2u);
    }
}

// ---- KIND_F32_WARP ---------------------------------------------------------------------------------
// Raw euclid-sum / dot of up to R rows against the warp's query (in shared memory, same permuted
// layout).  Returns the final reduced value of row r in out[r] on every lane.
// ROWS_SMEM: the rows were staged in shared memory by the bulk-copy ring (ring.cuh); else they are read from
// global memory with 128-bit loads.
template <int R, bool DOT, bool ROWS_SMEM>
__device__ __forceinline__ void warp_rows_raw(const DevIndex& ix, const float* qs, const uint8_t* const (&rowp)[R], float (&out)[R]) {
    const int lane = lane_id();
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.0f;
    const float4* q4 = reinterpret_cast<const float4*>(qs);
    for (uint32_t c = 0; c < ix.n_chunks; ++c) {
        float4 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float4* p4 = reinterpret_cast<const float4*>(rowp[r]) + c * 32 + lane;
            v[r] = ROWS_SMEM ? *p4 : __ldg(p4);
        }
        float4 qv = q4[c * 32 + lane];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (DOT) {  // dot_similarity_avx: acc = fma(a, b, acc)  (simple_avx.rs:85-97)
                acc[r] = fmaf(qv.x, v[r].x, acc[r]);
                acc[r] = fmaf(qv.y, v[r].y, acc[r]);
                acc[r] = fmaf(qv.z, v[r].z, acc[r]);
                acc[r] = fmaf(qv.w, v[r].w, acc[r]);
            } else {  // euclid_similarity_avx: d = a - b; acc = fma(d, d, acc)  (simple_avx.rs:33-53)
                float d0 = __fsub_rn(qv.x, v[r].x), d1 = __fsub_rn(qv.y, v[r].y);
                float d2 = __fsub_rn(qv.z, v[r].z), d3 = __fsub_rn(qv.w, v[r].w);
                acc[r] = fmaf(d0, d0, acc[r]);
                acc[r] = fmaf(d1, d1, acc[r]);
                acc[r] = fmaf(d2, d2, acc[r]);
                acc[r] = fmaf(d3, d3, acc[r]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float res = warp_hsum_avx(acc[r]);
        // scalar tail, n % 32 elements, unfused (simple_avx.rs:59-63,104-108); every lane redundantly
        const float* rt = reinterpret_cast<const float*>(rowp[r]) + ix.tail_off;
        const float* qt = qs + ix.tail_off;
        for (uint32_t e = 0; e < ix.tail; ++e) {
            float a = qt[e], b = ROWS_SMEM ? rt[e] : __ldg(rt + e);
            if (DOT) res = __fadd_rn(res, __fmul_rn(a, b));
            else { float d = __fsub_rn(a, b); res = __fadd_rn(res, __fmul_rn(d, d)); }
        }
        out[r] = res;
    }
}

// Same arithmetic for a group of R rows (staged in shared memory by the bulk-copy ring, or read straight from global
// memory with 128-bit loads), with the R horizontal sums folded into one butterfly: at each xor step a lane keeps
// half of the rows it still holds (it sends the partial sums it drops and receives the ones it keeps), so 8 rows
// cost 4+2+1 shuffles instead of 24.  Every individual addition is one the AVX code performs (x[l+4]+x[l]; r0+r2,
// r1+r3; s0+s1; ((h1+h2)+h3)+h4 — simple_avx.rs:8-13,55-58), fp32 addition being commutative, so the sums are
// bit-identical.  Returns row r's raw sum on the lanes l with (l & 7) == group_owner<R>(r) (all four 8-lane groups).
template <int R>
__device__ __forceinline__ int group_owner(int r) { return r * (8 / R); }

template <int R, bool DOT, bool ROWS_SMEM>
__device__ __forceinline__ float warp_rows_group(const DevIndex& ix, const float* qs, const uint8_t* const (&rowp)[R]) {
    static_assert(R == 1 || R == 2 || R == 4 || R == 8, "row group must be 1, 2, 4 or 8");
    const int lane = lane_id();
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.0f;
    const float4* q4 = reinterpret_cast<const float4*>(qs);
#pragma unroll 2
    for (uint32_t c = 0; c < ix.n_chunks; ++c) {
        float4 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float4* p4 = reinterpret_cast<const float4*>(rowp[r]) + c * 32 + lane;
            v[r] = ROWS_SMEM ? *p4 : __ldg(p4);
        }
        float4 qv = q4[c * 32 + lane];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (DOT) {  // dot_similarity_avx: acc = fma(a, b, acc)  (simple_avx.rs:85-97)
                acc[r] = fmaf(qv.x, v[r].x, acc[r]);
                acc[r] = fmaf(qv.y, v[r].y, acc[r]);
                acc[r] = fmaf(qv.z, v[r].z, acc[r]);
                acc[r] = fmaf(qv.w, v[r].w, acc[r]);
            } else {  // euclid_similarity_avx: d = a - b; acc = fma(d, d, acc)  (simple_avx.rs:33-53)
                float d0 = __fsub_rn(qv.x, v[r].x), d1 = __fsub_rn(qv.y, v[r].y);
                float d2 = __fsub_rn(qv.z, v[r].z), d3 = __fsub_rn(qv.w, v[r].w);
                acc[r] = fmaf(d0, d0, acc[r]);
                acc[r] = fmaf(d1, d1, acc[r]);
                acc[r] = fmaf(d2, d2, acc[r]);
                acc[r] = fmaf(d3, d3, acc[r]);
            }
        }
    }
    // fold: after the xor-4 / xor-2 / xor-1 steps lane l holds the 8-lane-group sum of row ((l & 7) * R) >> 3
    // (constant indices only: the partial sums must stay in registers)
#define HB_FOLD(i, j, step) { const bool hi_ = lane & step; float keep_ = hi_ ? acc[j] : acc[i]; float send_ = hi_ ? acc[i] : acc[j]; \
                              acc[i] = __fadd_rn(keep_, __shfl_xor_sync(FULL, send_, step)); }
#define HB_PLAIN(step) { acc[0] = __fadd_rn(acc[0], __shfl_xor_sync(FULL, acc[0], step)); }
    if (R == 8) {
        HB_FOLD(0, R > 4 ? 4 : 0, 4) HB_FOLD(R > 1 ? 1 : 0, R > 5 ? 5 : 0, 4) HB_FOLD(R > 2 ? 2 : 0, R > 6 ? 6 : 0, 4) HB_FOLD(R > 3 ? 3 : 0, R > 7 ? 7 : 0, 4)
        HB_FOLD(0, R > 2 ? 2 : 0, 2) HB_FOLD(R > 1 ? 1 : 0, R > 3 ? 3 : 0, 2)
        HB_FOLD(0, R > 1 ? 1 : 0, 1)
    } else if (R == 4) {
        HB_FOLD(0, R > 2 ? 2 : 0, 4) HB_FOLD(R > 1 ? 1 : 0, R > 3 ? 3 : 0, 4)
        HB_FOLD(0, R > 1 ? 1 : 0, 2)
        HB_PLAIN(1)
    } else if (R == 2) {
        HB_FOLD(0, R > 1 ? 1 : 0, 4)
        HB_PLAIN(2) HB_PLAIN(1)
    } else {
        HB_PLAIN(4) HB_PLAIN(2) HB_PLAIN(1)
    }
#undef HB_FOLD
#undef HB_PLAIN
    const float k = acc[0];
    const int sub = lane & 7;
    float h1 = __shfl_sync(FULL, k, sub), h2 = __shfl_sync(FULL, k, 8 + sub);
    float h3 = __shfl_sync(FULL, k, 16 + sub), h4 = __shfl_sync(FULL, k, 24 + sub);
    float res = __fadd_rn(__fadd_rn(__fadd_rn(h1, h2), h3), h4);
    if (ix.tail) {  // scalar tail, n % 32 elements, unfused (simple_avx.rs:59-63,104-108), on the lanes that hold the row
        const int my_r = (sub * R) >> 3;
        const uint8_t* rp = rowp[0];
#pragma unroll
        for (int r = 1; r < R; ++r) if (my_r == r) rp = rowp[r];
        const float* rt = reinterpret_cast<const float*>(rp) + ix.tail_off;
        const float* qt = qs + ix.tail_off;
        for (uint32_t e = 0; e < ix.tail; ++e) {
            float a = qt[e], b = ROWS_SMEM ? rt[e] : __ldg(rt + e);
            if (DOT) res = __fadd_rn(res, __fmul_rn(a, b));
            else { float d = __fsub_rn(a, b); res = __fadd_rn(res, __fmul_rn(d, d)); }
        }
    }
    return res;
}

// ---- KIND_F32_LANE: one lane walks one row sequentially ------------------------------------------------
__device__ __forceinline__ float hsum4_dev(const float* x) {  // hsum128_ps_sse, simple_sse.rs:12-16
    return __fadd_rn(__fadd_rn(x[0], x[2]), __fadd_rn(x[1], x[3]));
}
template <bool DOT, bool ROW_GLOBAL>
__device__ float lane_raw_small(const float* a /*query*/, const float* b /*row*/, uint32_t n) {
    auto ld = [&](uint32_t i) { return ROW_GLOBAL ? __ldg(b + i) : b[i]; };
    if (n >= 16) {  // SSE path (16 <= n < 32): one 16-wide iteration, mul then add, unfused
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float x = a[j], y = ld(j);
            float p;
            if (DOT) p = __fmul_rn(x, y);
            else { float d = __fsub_rn(x, y); p = __fmul_rn(d, d); }
            acc[j] = __fadd_rn(p, 0.0f);
        }
        float res = __fadd_rn(hsum4_dev(acc), hsum4_dev(acc + 4));
        res = __fadd_rn(res, hsum4_dev(acc + 8));
        res = __fadd_rn(res, hsum4_dev(acc + 12));
        for (uint32_t i = 16; i < n; ++i) {
            float x = a[i], y = ld(i);
            if (DOT) res = __fadd_rn(res, __fmul_rn(x, y));
            else { float d = __fsub_rn(x, y); res = __fadd_rn(res, __fmul_rn(d, d)); }
        }
        return res;
    }
    float s = 0.0f;  // scalar path, simple.rs:49-51,81-83
    for (uint32_t i = 0; i < n; ++i) {
        float x = a[i], y = ld(i);
        if (DOT) s = __fadd_rn(s, __fmul_rn(x, y));
        else { float d = __fsub_rn(x, y); s = __fadd_rn(s, __fmul_rn(d, d)); }
    }
    return s;
}
template <bool ROW_GLOBAL>
__device__ float lane_manhattan(const float* a, const float* b, uint32_t n) {  // manhattan.rs:41-43
    float s = 0.0f;
    for (uint32_t i = 0; i < n; ++i) {
        float y = ROW_GLOBAL ? __ldg(b + i) : b[i];
        s = __fadd_rn(s, fabsf(__fsub_rn(a[i], y)));
    }
    return s;
}
template <bool ROW_GLOBAL>
__device__ float lane_distance_f32(const DevIndex& ix, const float* qs, float qn, const float* row, float in) {
    if (ix.metric == HB_MANHATTAN) return lane_manhattan<ROW_GLOBAL>(qs, row, ix.dims);
    if (ix.metric == HB_COSINE) return finish_f32(HB_COSINE, lane_raw_small<true, ROW_GLOBAL>(qs, row, ix.dims), qn, in);
    return lane_raw_small<false, ROW_GLOBAL>(qs, row, ix.dims);
}

// ---- KIND_BIN: one lane walks one row ----------------------------------------------------------------
__device__ __forceinline__ uint32_t lane_xor_popc(const uint64_t* qs, const uint64_t* row, uint32_t n_words) {
    uint32_t h = 0;
    uint32_t pairs = (n_words + 1) >> 1;  // rows are padded to 16 bytes with zero words
    const ulonglong2* r2 = reinterpret_cast<const ulonglong2*>(row);
    const ulonglong2* q2 = reinterpret_cast<const ulonglong2*>(qs);
    for (uint32_t i = 0; i < pairs; ++i) {
        ulonglong2 a = q2[i], b = __ldg(r2 + i);
        h += __popcll(a.x ^ b.x) + __popcll(a.y ^ b.y);
    }
    return h;
}

}  // namespace hb
