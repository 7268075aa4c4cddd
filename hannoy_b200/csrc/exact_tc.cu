// exact_tc.cu — exact k-NN with the 5th-generation tensor cores (BASELINE north_star (c): "exact-kNN as a tensor-core GEMM").
//
// The ground truth of the recall rule must stay EXACT — same ids, same distance bits as the CUDA-core scan (exact.cu), i.e.
// as brute_force_search over every item (reader.rs:668-711) — so the tensor cores do not produce the answer, they prune:
//
//   1. tc_shortlist_kernel: a tf32 GEMM (tcgen05.mma, accumulators in TMEM, operands brought by TMA into 128-byte-swizzled
//      shared-memory tiles) gives every (query, item) pair an approximate key k~ with a RIGOROUS error bound eps: tf32 keeps 11
//      significant bits of each operand, so |dot~ - dot| <= E * |q| * |x| with E = 2^-9 (+ accumulation and fp32 slack, see
//      tc_error_model).  Queries sit on the M side (one TMEM lane = one query), items stream through N, so the epilogue is
//      thread-local: a thread reads its query's 128 keys of the tile out of TMEM, keeps the k smallest UPPER bounds k~ + eps
//      seen so far (tau = the k-th), and records every item whose LOWER bound k~ - eps does not exceed tau.  Any item of the true
//      top-k passes that test whatever the order of arrival (tau only shrinks and is never below the true k-th key), so the
//      recorded candidates are a superset of the answer — typically a few hundred out of a million.
//   2. tc_rerank_kernel: the candidates are re-evaluated with the bit-exact routines of dist.cuh (the reference's AVX summation
//      order) and the top-k by (distance bits, slot) is taken — identical to the full scan by construction.
//   3. Queries whose candidate list overflowed (or whose norm is degenerate) are re-run by the full scan (exact.cu).
//
// Cosine and Euclidean over KIND_F32_WARP rows (dims >= 32); everything else keeps the CUDA-core scan.
// Roles in tc_shortlist_kernel (192 threads): warp 0 = TMA producer (one lane), warp 1 = TMEM owner + MMA issuer (one lane),
// warps 2-5 = epilogue (TMEM lane quadrant = warp % 4).  4-stage operand ring, 2 accumulator buffers of 128 columns.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

#include "dist.cuh"
#include "ring.cuh"
#include "sorted.cuh"
#include "stage.cuh"

namespace hb {

hb_status launch_exact_knn_scan(const DevIndex& ix, const float* d_q, uint64_t nq, uint32_t k, uint32_t* d_ids, float* d_dist, void* stream);

namespace {

constexpr int TC_M = 128;        // queries per CTA: UMMA M = TMEM lanes
constexpr int TC_N = 128;        // items per tile: UMMA N = TMEM columns of one accumulator buffer
constexpr int TC_KB = 32;        // floats per K block = one 128-byte swizzle row
constexpr int TC_STAGES = 4;
constexpr int TC_THREADS = 192;
constexpr int TC_KMAX = 128;     // largest k served by this path
constexpr uint32_t TC_TILE_BYTES = TC_M * TC_KB * 4;  // 16 KB; the query tile and the item tile have the same shape
constexpr uint32_t TC_CAP = 1024;  // candidates recorded per (query, item slice) before the query is handed to the full scan

struct TcParams {
    uint32_t n, nq, k, kb_count, n_slices, tiles_per_slice, n_tiles, cap;
    const float4* aux;       // per item: key = fma(dot, aux.x, aux.y); eps = fma(|q|, aux.z, aux.w) + G * |q|^2
    const float* qnorm;      // per (padded) query: |q|, or a negative value for "skip" (padding rows, degenerate queries)
    float G;
    uint32_t* cand;          // [nq_pad][n_slices][cap] item slots
    uint32_t* cand_cnt;      // [nq_pad][n_slices] recorded count (> cap: overflow)
};

// ---- PTX wrappers (tcgen05 / TMA), sm_100a ----------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {  // arrives on `bar` once every MMA issued so far has completed
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {  // this thread's TMEM lane, 32 consecutive columns
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }

// Shared-memory matrix descriptor of a K-major tile of 128-byte rows written by TMA with SWIZZLE_128B: 8-row groups of 1024
// bytes (stride byte offset), descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B.  The tile base is 1024-byte aligned;
// a K step of 8 tf32 (32 bytes) advances the start address by 32 bytes inside the swizzle row.
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t addr) {
    return (uint64_t)((addr & 0x3ffffu) >> 4) | (1ull << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor: D = f32, A = B = tf32, both K-major, N = 128 (>> 3 at bit 17), M = 128 (>> 4 at bit 24).
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

// A candidate: record it, and tighten tau (the k-th smallest upper bound seen by this query).  Rare — about k ln(n / k) + the
// population of the error band per query — so it lives out of line.
__device__ __noinline__ float tc_accept(float key, float eps, uint32_t item, uint32_t qrow, float* uheap, uint32_t* hn_s, uint32_t* cnt_s,
                                        uint32_t* cand_base, uint32_t cap, uint32_t k) {
    const uint32_t cnt = cnt_s[qrow];
    if (cnt < cap) cand_base[cnt] = item;
    cnt_s[qrow] = cnt + 1;
    const float u = key + eps;
    uint32_t hn = hn_s[qrow];
    float* col = uheap + qrow;  // uheap[i * 128 + qrow], ascending in i
    if (hn < k || u < col[(size_t)(k - 1) * TC_M]) {
        int i = (int)(hn < k ? hn : k - 1);
        while (i > 0 && col[(size_t)(i - 1) * TC_M] > u) { col[(size_t)i * TC_M] = col[(size_t)(i - 1) * TC_M]; --i; }
        col[(size_t)i * TC_M] = u;
        if (hn < k) hn_s[qrow] = ++hn;
    }
    return hn == k ? col[(size_t)(k - 1) * TC_M] : INFINITY;
}

__global__ void __launch_bounds__(TC_THREADS, 1) tc_shortlist_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x,
                                                                     const TcParams P) {
    extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
    __shared__ unsigned long long bars[2 * TC_STAGES + 4];  // full[S], empty[S], tmem_full[2], tmem_empty[2]
    __shared__ uint32_t tmem_base_s;
    __shared__ uint32_t hn_s[TC_M], cnt_s[TC_M];
    const uint32_t raw = smem_addr(tc_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* sm = tc_smem_raw + (base - raw);
    float4* aux_s = reinterpret_cast<float4*>(sm + (size_t)TC_STAGES * 2 * TC_TILE_BYTES);  // [2][128]
    float* uheap = reinterpret_cast<float*>(aux_s + 2 * TC_N);                               // [k][128]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t full0 = smem_addr(&bars[0]), empty0 = smem_addr(&bars[TC_STAGES]);
    const uint32_t tfull0 = smem_addr(&bars[2 * TC_STAGES]), tempty0 = smem_addr(&bars[2 * TC_STAGES + 2]);

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull0 + 8 * b, 1); mbar_init(tempty0 + 8 * b, 4); }
        mbar_fence_init();
    }
    if (threadIdx.x < TC_M) { hn_s[threadIdx.x] = 0; cnt_s[threadIdx.x] = 0; }
    if (warp == 1) {  // TMEM: 2 accumulator buffers x 128 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_base_s)), "r"(2 * TC_N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    const uint32_t qt = blockIdx.x, slice = blockIdx.y;
    const uint32_t t0 = slice * P.tiles_per_slice, t1 = min(P.n_tiles, t0 + P.tiles_per_slice);

    if (warp == 0) {
        if (lane == 0) {  // ---- TMA producer ----
            uint32_t it = 0;
            for (uint32_t t = t0; t < t1; ++t) {
                for (uint32_t kb = 0; kb < P.kb_count; ++kb, ++it) {
                    const uint32_t s = it % TC_STAGES, ph = (it / TC_STAGES) & 1u;
                    mbar_wait(empty0 + 8 * s, ph ^ 1u);
                    mbar_expect_tx(full0 + 8 * s, 2 * TC_TILE_BYTES);
                    const uint32_t a_dst = base + s * 2 * TC_TILE_BYTES;
                    tma_load_2d(a_dst, &tmap_q, full0 + 8 * s, (int)(kb * TC_KB), (int)(qt * TC_M));
                    tma_load_2d(a_dst + TC_TILE_BYTES, &tmap_x, full0 + 8 * s, (int)(kb * TC_KB), (int)(t * TC_N));
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---- MMA issuer ----
            uint32_t it = 0, ti = 0;
            for (uint32_t t = t0; t < t1; ++t, ++ti) {
                const uint32_t b = ti & 1u, tph = (ti >> 1) & 1u;
                mbar_wait(tempty0 + 8 * b, tph ^ 1u);  // the epilogue has drained this accumulator buffer
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + b * TC_N;
                for (uint32_t kb = 0; kb < P.kb_count; ++kb, ++it) {
                    const uint32_t s = it % TC_STAGES, ph = (it / TC_STAGES) & 1u;
                    mbar_wait(full0 + 8 * s, ph);
                    tc_fence_after();
                    const uint64_t da = tc_smem_desc(base + s * 2 * TC_TILE_BYTES), db = tc_smem_desc(base + s * 2 * TC_TILE_BYTES + TC_TILE_BYTES);
#pragma unroll
                    for (uint32_t j = 0; j < TC_KB / 8; ++j)  // 8 tf32 = 32 bytes per MMA: + 2 in units of 16 bytes
                        tc_mma_tf32(d_tmem, da + 2 * j, db + 2 * j, TC_IDESC, (kb | j) != 0u);
                    tc_commit(empty0 + 8 * s);   // the stage may be refilled once these MMAs have read it
                }
                tc_commit(tfull0 + 8 * b);       // accumulator complete
            }
        }
    } else {
        // ---- epilogue: thread = one query (TMEM lane), a tile = 128 keys of that query ----
        const uint32_t quad = (uint32_t)warp & 3u, qrow = quad * 32 + lane, te = (uint32_t)(warp - 2) * 32 + lane;
        const uint32_t qg = qt * TC_M + qrow;
        const float qn = P.qnorm[qg];
        const float cq = P.G * qn * qn;
        float tau = qn >= 0.0f ? INFINITY : -INFINITY;  // padding rows and degenerate queries record nothing
        uint32_t* cand_base = P.cand + ((size_t)qg * P.n_slices + slice) * P.cap;
        uint32_t ti = 0;
        for (uint32_t t = t0; t < t1; ++t, ++ti) {
            const uint32_t b = ti & 1u, tph = (ti >> 1) & 1u;
            const uint32_t item0 = t * TC_N;
            aux_s[b * TC_N + te] = item0 + te < P.n ? __ldg(&P.aux[item0 + te]) : make_float4(0.0f, __int_as_float(0x7fc00000), 0.0f, 0.0f);
            asm volatile("bar.sync 1, 128;" ::: "memory");
            mbar_wait(tfull0 + 8 * b, tph);
            tc_fence_after();
#pragma unroll 1
            for (uint32_t c = 0; c < TC_N / 32; ++c) {
                uint32_t v[32];
                tc_ld32(tmem_base + ((quad * 32u) << 16) + b * TC_N + c * 32, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float4 a = aux_s[b * TC_N + c * 32 + j];
                    const float key = fmaf(__uint_as_float(v[j]), a.x, a.y);   // NaN for rows past the end: never accepted
                    const float eps = fmaf(qn, a.z, a.w) + cq;
                    if (key - eps <= tau) tau = tc_accept(key, eps, item0 + c * 32 + j, qrow, uheap, hn_s, cnt_s, cand_base, P.cap, P.k);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) tc_mbar_arrive(tempty0 + 8 * b);
        }
        if (qg < P.nq) P.cand_cnt[(size_t)qg * P.n_slices + slice] = cnt_s[qrow];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * TC_N) : "memory");
}

// ---- staging: queries into the device row layout (zero-padded to a multiple of 128 rows), their norms -----------------------
__global__ void __launch_bounds__(128) tc_stage_kernel(const __grid_constant__ DevIndex ix, const float* __restrict__ q, uint32_t nq, uint32_t nq_pad,
                                                       float* __restrict__ qbuf, float* __restrict__ qnorm, uint32_t* __restrict__ fallback) {
    __shared__ float qn_tmp[4];
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t qi = blockIdx.x * 4 + warp;
    if (qi >= nq_pad) return;
    float* dst = qbuf + (size_t)qi * (ix.row_stride / 4);
    if (qi >= nq) {
        for (uint32_t i = lane; i < ix.row_stride / 4; i += 32) dst[i] = 0.0f;
        if (lane == 0) qnorm[qi] = -1.0f;
        return;
    }
    ex_stage_query(ix, q + (size_t)qi * ix.dims, dst, &qn_tmp[warp]);
    float s = 0.0f;
    for (uint32_t e = lane; e < ix.dims; e += 32) { const float v = __ldg(q + (size_t)qi * ix.dims + e); s = fmaf(v, v, s); }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    if (lane == 0) {
        const float nrm = sqrtf(s);
        const bool ok = isfinite(nrm) && nrm > 1e-3f && nrm < 1e15f;   // degenerate queries go to the full scan (cosine.rs:44-55 returns 0.0 for tiny norms)
        qnorm[qi] = ok ? nrm : -1.0f;
        fallback[qi] = ok ? 0u : 1u;
    }
}

// ---- per item: the affine map dot -> key and the error-bound coefficients ------------------------------------------------------
//   cosine:     key = -dot / |x|       (the query's own norm is a positive per-query factor: dropped), eps = E * |q|
//   euclidean:  key = |x|^2 - 2 dot    (|q|^2 is a per-query constant: dropped),                     eps = 2 E |q| |x| + G (|q|^2 + |x|^2)
__global__ void __launch_bounds__(128) tc_aux_kernel(const __grid_constant__ DevIndex ix, float E, float G, float4* __restrict__ aux, uint32_t* __restrict__ n_bad) {
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t i = blockIdx.x * 4 + warp;
    if (i >= ix.n) return;
    float4 out;
    bool bad;
    if (ix.metric == HB_COSINE) {
        const float xn = __ldg(&ix.hdr[i]);
        bad = !(isfinite(xn) && xn > 1e-3f && xn < 1e15f);
        out = make_float4(-1.0f / xn, 0.0f, E, 0.0f);
    } else {
        const float4* row = reinterpret_cast<const float4*>(ix.rows + (size_t)i * ix.row_stride);
        float s = 0.0f;
        for (uint32_t w = lane; w < ix.row_stride / 16; w += 32) { const float4 v = __ldg(row + w); s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s); }
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
        bad = !(isfinite(s) && s < 1e30f);
        out = make_float4(-2.0f, s, 2.0f * E * sqrtf(s), G * s);
    }
    if (lane == 0) {
        aux[i] = out;
        if (bad) atomicAdd(n_bad, 1u);
    }
}

// ---- re-rank: the bit-exact distance of every candidate, top-k by (distance bits, slot) ---------------------------------------
__global__ void __launch_bounds__(128) tc_rerank_kernel(const __grid_constant__ DevIndex ix, const float* __restrict__ q, uint32_t nq, uint32_t k,
                                                        const uint32_t* __restrict__ cand, const uint32_t* __restrict__ cand_cnt, uint32_t n_slices, uint32_t cap,
                                                        uint32_t* __restrict__ out_ids, float* __restrict__ out_dist, uint32_t* __restrict__ fallback) {
    extern __shared__ __align__(16) unsigned char rr_smem[];
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t qi = blockIdx.x * 4 + warp;
    if (qi >= nq) return;
    const uint32_t qstride = (ix.row_stride + 15) & ~15u;
    const size_t per_warp = ((size_t)qstride + 16 + (size_t)k * 8 + 15) & ~(size_t)15;
    unsigned char* mine = rr_smem + per_warp * warp;
    float* qs = reinterpret_cast<float*>(mine);
    float* qn_p = reinterpret_cast<float*>(mine + qstride);
    u64* lst = reinterpret_cast<u64*>(mine + qstride + 16);
    ex_stage_query(ix, q + (size_t)qi * ix.dims, qs, qn_p);
    const float qn = *qn_p;
    int len = 0;
    bool over = false;
    for (uint32_t sl = 0; sl < n_slices; ++sl) {
        const uint32_t cnt = cand_cnt[(size_t)qi * n_slices + sl];
        if (cnt > cap) { over = true; break; }
        const uint32_t* cl = cand + ((size_t)qi * n_slices + sl) * cap;
        for (uint32_t c0 = 0; c0 < cnt; c0 += 4) {
            uint32_t s[4];
            const uint8_t* rowp[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) { s[r] = __ldg(&cl[min(c0 + r, cnt - 1)]); rowp[r] = ix.rows + (size_t)s[r] * ix.row_stride; }
            float rawv[4];
            if (ix.metric == HB_COSINE) warp_rows_raw<4, true, false>(ix, qs, rowp, rawv);
            else warp_rows_raw<4, false, false>(ix, qs, rowp, rawv);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (c0 + r >= cnt) break;
                const float in = ix.metric == HB_COSINE ? __ldg(&ix.hdr[s[r]]) : 0.0f;
                const float d = finish_f32(ix.metric, rawv[r], qn, in);
                topk_insert(lst, len, (int)k, ((u64)__float_as_uint(d) << 32) | s[r]);
            }
        }
    }
    if (over && lane == 0) fallback[qi] = 1u;
    for (int i = lane; i < (int)k; i += 32) {
        const bool ok = i < len;
        const u64 key = ok ? lst[i] : 0;
        out_ids[(size_t)qi * k + i] = ok ? __ldg(&ix.ids[(uint32_t)key]) : 0xffffffffu;
        out_dist[(size_t)qi * k + i] = ok ? __uint_as_float((uint32_t)(key >> 32)) : INFINITY;
    }
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
encode_tiled_fn tensor_map_encoder() {
    static encode_tiled_fn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = (encode_tiled_fn)p;
        else cudaGetLastError();
    });
    return fn;
}
// rows x kfloats f32, `row_bytes` apart; boxes of 128 rows x 32 floats, 128-byte swizzle, out-of-range elements read as zero
bool make_tile_map(CUtensorMap* m, const void* base, uint64_t rows, uint64_t kfloats, uint64_t row_bytes) {
    encode_tiled_fn enc = tensor_map_encoder();
    if (!enc) return false;
    cuuint64_t gdim[2] = {kfloats, rows};
    cuuint64_t gstr[1] = {row_bytes};
    cuuint32_t box[2] = {(cuuint32_t)TC_KB, (cuuint32_t)TC_M};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct DevFrees {
    std::vector<void*> v;
    ~DevFrees() { for (void* p : v) cudaFree(p); }
    template <class T> bool alloc(T** out, size_t bytes) {
        void* p = nullptr;
        if (cudaMalloc(&p, std::max<size_t>(bytes, 16)) != cudaSuccess) { cudaGetLastError(); return false; }
        v.push_back(p);
        *out = (T*)p;
        return true;
    }
};

}  // namespace

bool make_row_gather_map(void* out128, const void* base, uint64_t rows, uint64_t kfloats, uint64_t row_bytes) {
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
    encode_tiled_fn enc = tensor_map_encoder();
    if (!enc || kfloats == 0 || kfloats > 256 || (row_bytes & 15u)) return false;
    alignas(64) CUtensorMap m;
    cuuint64_t gdim[2] = {kfloats, rows};
    cuuint64_t gstr[1] = {row_bytes};
    cuuint32_t box[2] = {(cuuint32_t)kfloats, 1};
    cuuint32_t estr[2] = {1, 1};
    if (enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    memcpy(out128, &m, 128);
    return true;
}

// tf32 error model behind the candidate test: both operands lose at most their 13 low mantissa bits (relative 2^-10 each, 2^-9
// for the product, to first order), the products are exact, the fp32 accumulation of K terms adds at most K * 2^-23 relative to
// sum |q_i x_i| <= |q| |x|; the bit-exact fp32 routine the final ranking uses is itself within K * 2^-24 of the real dot product.
// E bounds |dot~ - dot| / (|q| |x|), G the fp32 rounding of the squared norms and of the exact Euclidean sum.
static void tc_error_model(uint32_t kfloats, float* E, float* G) {
    *E = 1.9922e-3f /* 2^-9 * 1.02 */ + (float)kfloats * 2.0e-7f;
    *G = (float)kfloats * 2.5e-7f + 1e-6f;
}

// Exact k-NN through the tensor-core shortlist.  Returns HB_OK and sets *done = false when the path does not apply (the caller
// then runs the full scan).
hb_status launch_exact_knn_tc(const DevIndex& ix, const float* d_q, uint64_t nq, uint32_t k, uint32_t* d_ids, float* d_dist, void* stream_, bool* done) {
    *done = false;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!tunable("exact_tc", 1) || ix.kind != KIND_F32_WARP || (ix.metric != HB_COSINE && ix.metric != HB_EUCLIDEAN)) return HB_OK;
    if (k == 0 || k > (uint32_t)TC_KMAX || nq == 0 || ix.n < 4 * (uint32_t)TC_N || nq > 0x7fffff00ull || (ix.row_stride & 15u)) return HB_OK;
    if ((size_t)nq * ix.n < (size_t)tunable("exact_tc_min_pairs", 1 << 22)) return HB_OK;   // tiny problems: the scan is as fast
    const uint32_t kf = ix.row_stride / 4;
    const uint32_t nq_pad = (uint32_t)((nq + TC_M - 1) / TC_M * TC_M);
    const uint32_t n_qtiles = nq_pad / TC_M;
    const uint32_t n_tiles = (ix.n + TC_N - 1) / TC_N;
    int n_sm = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    uint32_t n_slices = std::max<uint32_t>(1, std::min<uint32_t>(n_tiles, (uint32_t)n_sm / n_qtiles));
    const uint32_t tiles_per_slice = (n_tiles + n_slices - 1) / n_slices;
    n_slices = (n_tiles + tiles_per_slice - 1) / tiles_per_slice;
    float E, G;
    tc_error_model(kf, &E, &G);
    if (ix.metric == HB_COSINE) G = 0.0f;

    DevFrees fr;
    float *qbuf = nullptr, *qnorm = nullptr;
    float4* aux = nullptr;
    uint32_t *cand = nullptr, *cand_cnt = nullptr, *fallback = nullptr, *n_bad = nullptr;
    if (!fr.alloc(&qbuf, (size_t)nq_pad * ix.row_stride) || !fr.alloc(&qnorm, (size_t)nq_pad * 4) || !fr.alloc(&aux, (size_t)ix.n * 16) ||
        !fr.alloc(&cand, (size_t)nq_pad * n_slices * TC_CAP * 4) || !fr.alloc(&cand_cnt, (size_t)nq_pad * n_slices * 4) ||
        !fr.alloc(&fallback, (size_t)nq_pad * 4) || !fr.alloc(&n_bad, 16))
        return HB_OK;  // not enough room for the shortlist buffers: the scan needs none
    cudaMemsetAsync(fallback, 0, (size_t)nq_pad * 4, stream);
    cudaMemsetAsync(n_bad, 0, 16, stream);
    cudaMemsetAsync(cand_cnt, 0, (size_t)nq_pad * n_slices * 4, stream);
    tc_stage_kernel<<<(nq_pad + 3) / 4, 128, 0, stream>>>(ix, d_q, (uint32_t)nq, nq_pad, qbuf, qnorm, fallback);
    tc_aux_kernel<<<(ix.n + 3) / 4, 128, 0, stream>>>(ix, E, G, aux, n_bad);
    g_launches += 2;
    uint32_t h_bad = 0;
    if (cudaMemcpyAsync(&h_bad, n_bad, 4, cudaMemcpyDeviceToHost, stream) != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess) {
        set_error("exact_knn (tensor-core path): staging failed: %s", cudaGetErrorString(cudaGetLastError()));
        return HB_ECUDA;
    }
    if (h_bad) return HB_OK;  // zero / non-finite items: cosine.rs:44-55 special-cases them, the scan handles that literally

    CUtensorMap tmap_q, tmap_x;
    if (!make_tile_map(&tmap_q, qbuf, nq_pad, kf, ix.row_stride) || !make_tile_map(&tmap_x, ix.rows, ix.n, kf, ix.row_stride)) return HB_OK;
    TcParams P;
    P.n = ix.n; P.nq = (uint32_t)nq; P.k = k; P.kb_count = (kf + TC_KB - 1) / TC_KB; P.n_slices = n_slices; P.tiles_per_slice = tiles_per_slice;
    P.n_tiles = n_tiles; P.cap = TC_CAP; P.aux = aux; P.qnorm = qnorm; P.G = G; P.cand = cand; P.cand_cnt = cand_cnt;
    const size_t smem = 1024 + (size_t)TC_STAGES * 2 * TC_TILE_BYTES + 2 * TC_N * sizeof(float4) + (size_t)k * TC_M * 4;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(tc_shortlist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess) { cudaGetLastError(); return HB_OK; }
        cudaFuncSetAttribute(tc_rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr = true;
    }
    tc_shortlist_kernel<<<dim3(n_qtiles, n_slices), TC_THREADS, smem, stream>>>(tmap_q, tmap_x, P);
    const uint32_t qstride = (ix.row_stride + 15) & ~15u;
    const size_t rr_per_warp = ((size_t)qstride + 16 + (size_t)k * 8 + 15) & ~(size_t)15;
    if (rr_per_warp * 4 > 200 * 1024) { set_error("exact_knn: dims too large"); return HB_EINVAL; }
    tc_rerank_kernel<<<(unsigned)((nq + 3) / 4), 128, rr_per_warp * 4, stream>>>(ix, d_q, (uint32_t)nq, k, cand, cand_cnt, n_slices, TC_CAP, d_ids, d_dist, fallback);
    g_launches += 2;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("exact_knn (tensor-core path) launch failed: %s", cudaGetErrorString(e)); return HB_ECUDA; }
    // queries handed back to the full scan
    std::vector<uint32_t> fb(nq);
    if (cudaMemcpyAsync(fb.data(), fallback, nq * 4, cudaMemcpyDeviceToHost, stream) != cudaSuccess || (e = cudaStreamSynchronize(stream)) != cudaSuccess) {
        set_error("exact_knn (tensor-core path) failed: %s", cudaGetErrorString(e != cudaSuccess ? e : cudaGetLastError()));
        return HB_ECUDA;
    }
    std::vector<uint32_t> redo;
    for (uint64_t i = 0; i < nq; ++i) if (fb[i]) redo.push_back((uint32_t)i);
    static const bool dbg = getenv("HB_DEBUG_LAUNCH") != nullptr;
    if (dbg) fprintf(stderr, "[hb] exact_knn tensor-core path: %u x %u CTAs, %u K blocks, %zu of %llu queries handed to the scan\n", n_qtiles, n_slices, P.kb_count, redo.size(), (unsigned long long)nq);
    if (!redo.empty()) {
        float* q2 = nullptr; uint32_t* i2 = nullptr; float* d2 = nullptr;
        if (!fr.alloc(&q2, redo.size() * (size_t)ix.dims * 4) || !fr.alloc(&i2, redo.size() * (size_t)k * 4) || !fr.alloc(&d2, redo.size() * (size_t)k * 4)) return HB_ENOMEM;
        for (size_t j = 0; j < redo.size(); ++j) cudaMemcpyAsync(q2 + j * ix.dims, d_q + (size_t)redo[j] * ix.dims, (size_t)ix.dims * 4, cudaMemcpyDeviceToDevice, stream);
        hb_status st = launch_exact_knn_scan(ix, q2, redo.size(), k, i2, d2, stream);
        if (st != HB_OK) return st;
        for (size_t j = 0; j < redo.size(); ++j) {
            cudaMemcpyAsync(d_ids + (size_t)redo[j] * k, i2 + j * k, (size_t)k * 4, cudaMemcpyDeviceToDevice, stream);
            cudaMemcpyAsync(d_dist + (size_t)redo[j] * k, d2 + j * k, (size_t)k * 4, cudaMemcpyDeviceToDevice, stream);
        }
        if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) { set_error("exact_knn (scan of the handed-back queries) failed: %s", cudaGetErrorString(e)); return HB_ECUDA; }
    }
    *done = true;
    return HB_OK;
}

}  // namespace hb
