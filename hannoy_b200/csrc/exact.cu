// exact.cu — exact k-NN over every item (recall ground truth) and the per-shard top-k merge.
// The exact scan uses the same distance routines as the graph walk, so its distances are bit-identical
// to the search kernel's and to the reference's (dist.cuh).  For Cosine / Euclidean rows of 32+ dimensions the scan is
// preceded by a tf32 tensor-core GEMM that prunes the items to a provably sufficient shortlist (exact_tc.cu); the
// distances that are ranked are always the bit-exact ones computed here.
#include <cfloat>

#include "dist.cuh"
#include "sorted.cuh"
#include "stage.cuh"

namespace hb {

hb_status launch_exact_knn_scan(const DevIndex& ix, const float* d_q, uint64_t nq, uint32_t k, uint32_t* d_ids, float* d_dist, void* stream);

constexpr int EX_WARPS = 8;
constexpr int EX_TQ = 8;  // queries per block

__global__ void __launch_bounds__(EX_WARPS * 32) exact_knn_kernel(const __grid_constant__ DevIndex ix, const float* __restrict__ q,
                                                                   uint64_t nq, uint32_t k, uint32_t* __restrict__ out_ids,
                                                                   float* __restrict__ out_dist) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t qstride = (ix.row_stride + 15) & ~15u;
    float* qs_all = reinterpret_cast<float*>(smem);
    float* qn_all = reinterpret_cast<float*>(smem + (size_t)qstride * EX_TQ);
    u64* lists = reinterpret_cast<u64*>(smem + (size_t)qstride * EX_TQ + 64);  // [warp][tq][k]
    int* lens = reinterpret_cast<int*>(lists + (size_t)EX_WARPS * EX_TQ * k);   // [warp][tq]
    const uint64_t q0 = (uint64_t)blockIdx.x * EX_TQ;
    const int tq = (int)min((uint64_t)EX_TQ, nq - q0);
    for (int t = warp; t < tq; t += EX_WARPS) ex_stage_query(ix, q + (q0 + t) * ix.dims, qs_all + (size_t)t * qstride / 4, qn_all + t);
    if (lane < EX_TQ) lens[warp * EX_TQ + lane] = 0;
    __syncthreads();
    u64* my = lists + (size_t)warp * EX_TQ * k;
    int* mylen = lens + warp * EX_TQ;

    if (ix.kind == KIND_F32_WARP) {
        for (uint32_t r0 = warp * 4; r0 < ix.n; r0 += EX_WARPS * 4) {
            const uint8_t* rowp[4];
            uint32_t sl[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) { sl[r] = min(r0 + r, ix.n - 1); rowp[r] = ix.rows + (size_t)sl[r] * ix.row_stride; }
            for (int t = 0; t < tq; ++t) {
                float raw[4];
                const float* qs = qs_all + (size_t)t * qstride / 4;
                if (ix.metric == HB_COSINE) warp_rows_raw<4, true, false>(ix, qs, rowp, raw);
                else warp_rows_raw<4, false, false>(ix, qs, rowp, raw);
                int len = mylen[t];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    if (r0 + r >= ix.n) break;
                    float in = (ix.metric == HB_COSINE) ? __ldg(&ix.hdr[sl[r]]) : 0.0f;
                    float d = finish_f32(ix.metric, raw[r], qn_all[t], in);
                    u64 key = ((u64)__float_as_uint(d) << 32) | sl[r];
                    topk_insert(my + (size_t)t * k, len, (int)k, key);
                }
                if (lane == 0) mylen[t] = len;
                __syncwarp();
            }
        }
    } else {
        for (uint32_t r0 = warp * 32; r0 < ix.n; r0 += EX_WARPS * 32) {
            uint32_t s = r0 + lane;
            bool valid = s < ix.n;
            for (int t = 0; t < tq; ++t) {
                const float* qs = qs_all + (size_t)t * qstride / 4;
                float d = 0.0f;
                if (valid) {
                    if (ix.kind == KIND_F32_LANE) {
                        const float* row = reinterpret_cast<const float*>(ix.rows + (size_t)s * ix.row_stride);
                        float in = (ix.metric == HB_COSINE) ? __ldg(&ix.hdr[s]) : 0.0f;
                        d = lane_distance_f32<true>(ix, qs, qn_all[t], row, in);
                    } else {
                        const uint64_t* row = reinterpret_cast<const uint64_t*>(ix.rows + (size_t)s * ix.row_stride);
                        uint32_t h = lane_xor_popc(reinterpret_cast<const uint64_t*>(qs), row, ix.n_words);
                        float in = (ix.metric == HB_BQ_COSINE) ? __ldg(&ix.hdr[s]) : 0.0f;
                        d = finish_bin(ix.metric, h, ix.n_words * 64u, qn_all[t], in);
                    }
                }
                u64 key = ((u64)__float_as_uint(d) << 32) | s;
                int len = mylen[t];
                u64* lst = my + (size_t)t * k;
                bool want = valid && (len < (int)k || key < lst[len - 1]);
                for (unsigned m = __ballot_sync(FULL, want); m; m &= m - 1) {
                    u64 kk = __shfl_sync(FULL, key, __ffs(m) - 1);
                    topk_insert(lst, len, (int)k, kk);
                }
                if (lane == 0) mylen[t] = len;
                __syncwarp();
            }
        }
    }
    __syncthreads();
    // merge the per-warp lists of each query into warp 0's list
    for (int t = warp; t < tq; t += EX_WARPS) {
        u64* dst = lists + (size_t)t * k;  // warp 0's list for query t
        int len = lens[t];
        for (int w = 1; w < EX_WARPS; ++w) {
            const u64* src = lists + ((size_t)w * EX_TQ + t) * k;
            int sl = lens[w * EX_TQ + t];
            for (int i = 0; i < sl; ++i) topk_insert(dst, len, (int)k, src[i]);
        }
        for (int i = lane; i < (int)k; i += 32) {
            bool ok = i < len;
            u64 key = ok ? dst[i] : 0;
            out_ids[(q0 + t) * k + i] = ok ? __ldg(&ix.ids[(uint32_t)key]) : 0xffffffffu;
            out_dist[(q0 + t) * k + i] = ok ? __uint_as_float((uint32_t)(key >> 32)) : INFINITY;
        }
    }
}

hb_status launch_exact_knn_tc(const DevIndex& ix, const float* d_q, uint64_t nq, uint32_t k, uint32_t* d_ids, float* d_dist, void* stream, bool* done);

// Exact k-NN: through the tensor-core shortlist (exact_tc.cu) where it applies, else the full CUDA-core scan.  Same output either way.
hb_status launch_exact_knn(const DevIndex& ix, const float* d_q, uint64_t nq, uint32_t k, uint32_t* d_ids, float* d_dist, void* stream) {
    bool done = false;
    hb_status st = launch_exact_knn_tc(ix, d_q, nq, k, d_ids, d_dist, stream, &done);
    if (st != HB_OK || done) return st;
    return launch_exact_knn_scan(ix, d_q, nq, k, d_ids, d_dist, stream);
}

hb_status launch_exact_knn_scan(const DevIndex& ix, const float* d_q, uint64_t nq, uint32_t k, uint32_t* d_ids, float* d_dist, void* stream) {
    uint32_t qstride = (ix.row_stride + 15) & ~15u;
    size_t smem = (size_t)qstride * EX_TQ + 64 + (size_t)EX_WARPS * EX_TQ * k * 8 + EX_WARPS * EX_TQ * 4 + 16;
    if (smem > 200 * 1024) { set_error("exact_knn: k=%u / dims too large for shared memory (%zu bytes)", k, smem); return HB_EINVAL; }
    cudaFuncSetAttribute(exact_knn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    unsigned blocks = (unsigned)((nq + EX_TQ - 1) / EX_TQ);
    exact_knn_kernel<<<blocks, EX_WARPS * 32, smem, (cudaStream_t)stream>>>(ix, d_q, nq, k, d_ids, d_dist);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("exact_knn launch failed: %s", cudaGetErrorString(e)); return HB_ECUDA; }
    return HB_OK;
}

// ---- per-shard top-k merge (after the all-gather of an id-sharded index) -----------------------------------
__global__ void __launch_bounds__(128) merge_topk_kernel(const uint32_t* __restrict__ ids, const float* __restrict__ dist, uint32_t n_parts,
                                                         uint64_t nq, uint32_t k, uint32_t* __restrict__ out_ids, float* __restrict__ out_dist,
                                                         uint32_t* __restrict__ out_len) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = lane_id();
    uint64_t qi = (uint64_t)blockIdx.x * 4 + warp;
    if (qi >= nq) return;
    u64* lst = reinterpret_cast<u64*>(smem) + (size_t)warp * k;
    int len = 0;
    for (uint32_t p = 0; p < n_parts; ++p) {
        const uint32_t* pi = ids + ((size_t)p * nq + qi) * k;
        const float* pd = dist + ((size_t)p * nq + qi) * k;
        for (uint32_t b = 0; b < k; b += 32) {
            uint32_t i = b + lane;
            uint32_t id = i < k ? pi[i] : 0xffffffffu;
            float d = i < k ? pd[i] : 0.0f;
            u64 key = ((u64)__float_as_uint(d) << 32) | id;
            bool want = id != 0xffffffffu && (len < (int)k || key < lst[len - 1]);
            for (unsigned m = __ballot_sync(FULL, want); m; m &= m - 1) {
                u64 kk = __shfl_sync(FULL, key, __ffs(m) - 1);
                topk_insert(lst, len, (int)k, kk);
            }
        }
    }
    for (int i = lane; i < (int)k; i += 32) {
        bool ok = i < len;
        u64 key = ok ? lst[i] : 0;
        out_ids[qi * k + i] = ok ? (uint32_t)key : 0xffffffffu;
        out_dist[qi * k + i] = ok ? __uint_as_float((uint32_t)(key >> 32)) : INFINITY;
    }
    if (out_len && lane == 0) out_len[qi] = (uint32_t)len;
}

hb_status launch_merge_topk(const uint32_t* d_ids, const float* d_dist, uint32_t n_parts, uint64_t nq, uint32_t k,
                            uint32_t* d_out_ids, float* d_out_dist, uint32_t* d_out_len, void* stream) {
    if (nq == 0 || k == 0) return HB_OK;
    size_t smem = (size_t)4 * k * 8;
    if (smem > 200 * 1024) { set_error("merge_topk: k too large"); return HB_EINVAL; }
    cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    merge_topk_kernel<<<(unsigned)((nq + 3) / 4), 128, smem, (cudaStream_t)stream>>>(d_ids, d_dist, n_parts, nq, k, d_out_ids, d_out_dist, d_out_len);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("merge_topk launch failed: %s", cudaGetErrorString(e)); return HB_ECUDA; }
    return HB_OK;
}

}  // namespace hb
