// stage.cuh — UnalignedVector::from_slice + D::new_header (reader.rs:140-141) for one raw f32 query, by one warp, into
// the device row layout (common.h RowKind): shared by the exact-scan kernels (exact.cu, exact_tc.cu).  The search
// kernel has its own copy specialised on the row kind (search.cu stage_query).
#pragma once
#include "dist.cuh"

namespace hb {

// stage one raw f32 query into the device row layout (same code path as search.cu:stage_query, by_vector)
__device__ inline void ex_stage_query(const DevIndex& ix, const float* src, float* qs, float* qn_out) {
    const int lane = lane_id();
    const uint32_t words16 = ix.row_stride / 16;
    uint4* q16 = reinterpret_cast<uint4*>(qs);
    for (uint32_t i = lane; i < words16; i += 32) q16[i] = make_uint4(0, 0, 0, 0);
    __syncwarp();
    if (ix.kind == KIND_F32_WARP) {
        uint32_t main = ix.dims - ix.tail;
        for (uint32_t e = lane; e < ix.dims; e += 32) {
            float v = __ldg(src + e);
            if (e < main) { uint32_t blk = e >> 5, j = e & 31; qs[(blk >> 2) * 128 + j * 4 + (blk & 3)] = v; }
            else qs[ix.tail_off + (e - main)] = v;
        }
    } else if (ix.kind == KIND_F32_LANE) {
        for (uint32_t e = lane; e < ix.dims; e += 32) qs[e] = __ldg(src + e);
    } else {
        uint32_t* q32 = reinterpret_cast<uint32_t*>(qs);
        for (uint32_t base = 0; base < ix.dims; base += 32) {
            uint32_t e = base + lane;
            bool bit = false;
            if (e < ix.dims) {
                uint32_t u = __float_as_uint(__ldg(src + e));
                bit = (ix.metric == HB_HAMMING) ? (u < 0x80000000u && u > 0u) : ((u >> 31) == 0);
            }
            unsigned w = __ballot_sync(FULL, bit);
            if (lane == 0) q32[base >> 5] = w;
        }
    }
    __syncwarp();
    float qn = 0.0f;
    if (ix.metric == HB_COSINE) {
        float dot;
        if (ix.kind == KIND_F32_WARP) {
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
        qn = __fsqrt_rn(dot);
    } else if (ix.metric == HB_BQ_COSINE) {
        qn = __fsqrt_rn((float)(int)(ix.n_words * 64u));
    }
    if (lane == 0) *qn_out = qn;
    __syncwarp();
}

}  // namespace hb
