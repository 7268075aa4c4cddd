// gather4_bench.cu — dev microbenchmark (GPU box), companion of gather_bench.cu: ceilings of the short-row gathers.
//   (a) TMA tile::gather4 (four rows per copy instruction, as ring.cuh post_gather4) for 512- and 128-byte rows;
//   (b) plain-load gathers of 128-byte rows (the binary codes of config 4): one lane per row (8 x LDG.128 per lane, the
//       search kernel's pattern) against eight lanes per row (one coalesced LDG.128 per lane, four rows per instruction).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/gather4_bench tools/gather4_bench.cu -lcuda && tools/gather4_bench
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

#include "../hannoy_b200/csrc/ring.cuh"

using namespace hb;

__device__ __forceinline__ uint64_t lcg(uint64_t& s) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    return s >> 20;
}

__global__ void gather4_kernel(const __grid_constant__ CUtensorMap tmap, uint64_t n_rows, uint32_t row_bytes, uint32_t slots, uint32_t iters, float* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t per_warp = ((size_t)slots * row_bytes + slots * 8 + 127) & ~(size_t)127;
    unsigned char* base = smem + per_warp * warp;
    RowRing ring;
    ring.ptr = base;
    ring.data = smem_addr(base);
    ring.bars = smem_addr(base + (size_t)slots * row_bytes);
    ring.slots = slots;
    ring.stride = row_bytes;
    ring.phase = 0;
    ring.policy = l2_policy_evict_first();
    if (lane == 0) {
        for (uint32_t i = 0; i < slots; ++i) mbar_init(ring.bars + i * 8, 1);
        mbar_fence_init();
    }
    __syncwarp();
    uint64_t state = (uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
    float acc = 0.f;
    const uint32_t groups = slots / 4;
    if (lane < groups) ring.post_gather4(lane * 4, &tmap, lcg(state) % n_rows, lcg(state) % n_rows, lcg(state) % n_rows, lcg(state) % n_rows);
    for (uint32_t it = 0; it < iters; ++it) {
        for (uint32_t g = 0; g < groups; ++g) {
            ring.wait(g * 4);
            const float4* p = reinterpret_cast<const float4*>(base + (size_t)g * 4 * row_bytes);
            for (uint32_t c = lane; c < 4 * row_bytes / 16; c += 32) { float4 v = p[c]; acc += v.x + v.y + v.z + v.w; }
            __syncwarp();
            if (lane == g && it + 1 < iters) ring.post_gather4(g * 4, &tmap, lcg(state) % n_rows, lcg(state) % n_rows, lcg(state) % n_rows, lcg(state) % n_rows);
        }
    }
    if (acc == 123.456f) *sink = acc;
}

// 128-byte rows, one lane per row: every lane reads its own row with 8 x LDG.128; `live` lanes per round (the walk has ~14)
__global__ void ldg128_lane_kernel(const uint8_t* __restrict__ rows, uint64_t n_rows, uint32_t iters, uint32_t live, float* sink) {
    const int lane = threadIdx.x & 31;
    uint64_t state = (uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345;
    uint32_t h = 0;
    for (uint32_t it = 0; it < iters; ++it) {
        const ulonglong2* p = reinterpret_cast<const ulonglong2*>(rows + (lcg(state) % n_rows) * 128);
        if ((uint32_t)lane < live) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { ulonglong2 v = __ldg(p + i); h += __popcll(v.x) + __popcll(v.y); }
        }
    }
    if (h == 0x12345678u) *sink = (float)h;
}
// 128-byte rows, eight lanes per row: one coalesced LDG.128 per lane = four rows per instruction, ceil(live / 4) instructions per round
__global__ void ldg128_group_kernel(const uint8_t* __restrict__ rows, uint64_t n_rows, uint32_t iters, uint32_t live, float* sink) {
    const int lane = threadIdx.x & 31;
    uint64_t state = (uint64_t)(blockIdx.x * blockDim.x + (threadIdx.x & ~31)) * 0x9E3779B97F4A7C15ull + 12345;
    uint32_t h = 0;
    const uint32_t rounds = (live + 3) / 4;
    for (uint32_t it = 0; it < iters; ++it) {
        ulonglong2 v[8];
#pragma unroll
        for (uint32_t r = 0; r < 8; ++r) {
            if (r < rounds) {
                uint64_t mine = 0;
                for (int g = 0; g < 4; ++g) { uint64_t x = lcg(state) % n_rows; if ((lane >> 3) == g) mine = x; }
                v[r] = __ldg(reinterpret_cast<const ulonglong2*>(rows + mine * 128) + (lane & 7));
            }
        }
#pragma unroll
        for (uint32_t r = 0; r < 8; ++r)
            if (r < rounds) h += __popcll(v[r].x) + __popcll(v[r].y);
    }
    if (h == 0x12345678u) *sink = (float)h;
}

typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                              CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    const size_t bytes = 3ull << 30;
    uint8_t* rows;
    float* sink;
    cudaMalloc(&rows, bytes);
    cudaMalloc(&sink, 4);
    cudaMemset(rows, 1, bytes);
    cudaFuncSetAttribute(gather4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    encode_fn enc = (encode_fn)fn;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    std::printf("{\"device\": \"%s\", \"sms\": %d, \"results\": [\n", prop.name, sms);
    bool first = true;
    for (uint32_t row : {512u, 128u}) {
        const uint64_t n_rows = bytes / row;
        alignas(64) CUtensorMap m;
        cuuint64_t gdim[2] = {row / 4, n_rows};
        cuuint64_t gstr[1] = {row};
        cuuint32_t box[2] = {row / 4, 1};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, rows, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { std::printf("{\"error\": \"tensor map %d\"}]}\n", (int)r); return 1; }
        for (uint32_t wps : {8u, 12u, 16u, 20u, 24u, 32u})
            for (uint32_t slots : {4u, 8u, 16u, 32u}) {
                const int wpb = 4;
                size_t per_warp = ((size_t)slots * row + slots * 8 + 127) & ~(size_t)127;
                if (per_warp * wps > 220 * 1024) continue;
                size_t smem = per_warp * wpb;
                size_t pad = (size_t)(224 * 1024) / (wps / wpb) / 128 * 128 - 1024;
                if (pad > smem) smem = pad;
                int blocks = sms * wps / wpb;
                uint32_t iters = (uint32_t)((12ull << 30) / ((uint64_t)blocks * wpb * slots * row));
                if (iters < 4) iters = 4;
                gather4_kernel<<<blocks, wpb * 32, smem>>>(m, n_rows, row, slots, 2, sink);
                cudaEventRecord(e0);
                gather4_kernel<<<blocks, wpb * 32, smem>>>(m, n_rows, row, slots, iters, sink);
                cudaEventRecord(e1);
                cudaError_t err = cudaDeviceSynchronize();
                float ms = 0;
                cudaEventElapsedTime(&ms, e0, e1);
                double gb = (double)blocks * wpb * slots * row * iters / 1e9;
                std::printf("%s {\"kind\": \"gather4\", \"row_bytes\": %u, \"warps_per_sm\": %u, \"slots\": %u, \"gbs\": %.1f, \"err\": \"%s\"}", first ? "" : ",\n", row, wps,
                            slots, gb / (ms / 1e3), err == cudaSuccess ? "" : cudaGetErrorString(err));
                first = false;
                std::fflush(stdout);
            }
    }
    const uint64_t n128 = bytes / 128;
    for (int kind = 0; kind < 2; ++kind)
        for (uint32_t live : {14u, 32u})
            for (int wps : {8, 12, 16, 16, 20, 24, 32, 48, 64}) {
                int blocks = sms * wps / 4;
                uint32_t iters = (uint32_t)((6ull << 30) / ((uint64_t)blocks * 4 * live * 128));
                auto launch = [&](uint32_t it) {
                    if (kind == 0) ldg128_lane_kernel<<<blocks, 128>>>(rows, n128, it, live, sink);
                    else ldg128_group_kernel<<<blocks, 128>>>(rows, n128, it, live, sink);
                };
                launch(2);
                cudaEventRecord(e0);
                launch(iters);
                cudaEventRecord(e1);
                cudaDeviceSynchronize();
                float ms = 0;
                cudaEventElapsedTime(&ms, e0, e1);
                double gb = (double)blocks * 4 * live * 128 * iters / 1e9;
                std::printf(",\n {\"kind\": \"%s\", \"row_bytes\": 128, \"live_rows_per_round\": %u, \"warps_per_sm\": %d, \"gbs\": %.1f}", kind == 0 ? "ldg_lane_per_row" : "ldg_8_lanes_per_row",
                            live, wps, gb / (ms / 1e3));
                std::fflush(stdout);
            }
    std::printf("\n]}\n");
    return 0;
}
