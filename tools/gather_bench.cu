// gather_bench.cu — dev microbenchmark (GPU box): what bandwidth can random row gathers reach on this part?
// Every warp keeps a ring of S bulk async copies (cp.async.bulk, same primitive as ring.cuh) of random
// R-byte rows in flight and touches the landed bytes once.  Prints GB/s for a sweep of (R, warps/SM, S).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench tools/gather_bench.cu && ./gather_bench
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

#include "../hannoy_b200/csrc/ring.cuh"

using namespace hb;

__global__ void gather_kernel(const uint8_t* __restrict__ rows, uint64_t n_rows, uint32_t row_bytes, uint32_t slots,
                              uint32_t iters, uint32_t touch_all, float* sink) {
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
    auto next = [&]() {
        state = state * 6364136223846793005ull + 1442695040888963407ull;
        return (state >> 20) % n_rows;
    };
    float acc = 0.f;
    // prologue: lane i posts slot i
    if (lane < slots) ring.post(lane, rows + next() * row_bytes, row_bytes);
    for (uint32_t it = 0; it < iters; ++it) {
        for (uint32_t s = 0; s < slots; ++s) {
            ring.wait(s);
            const float4* p = reinterpret_cast<const float4*>(base + (size_t)s * row_bytes);
            if (touch_all) {
                for (uint32_t c = lane; c < row_bytes / 16; c += 32) { float4 v = p[c]; acc += v.x + v.y + v.z + v.w; }
            } else {
                acc += p[lane % (row_bytes / 16)].x;
            }
            __syncwarp();
            if (lane == s && it + 1 < iters) ring.post(s, rows + next() * row_bytes, row_bytes);
        }
    }
    if (acc == 123.456f) *sink = acc;
}

// same gather with plain 128-bit loads, 4 rows x (row_bytes/512) loads per lane in flight
__global__ void gather_ldg_kernel(const uint8_t* __restrict__ rows, uint64_t n_rows, uint32_t row_bytes, uint32_t iters, float* sink) {
    const int lane = threadIdx.x & 31;
    uint64_t state = (uint64_t)(blockIdx.x * blockDim.x + (threadIdx.x & ~31)) * 0x9E3779B97F4A7C15ull + 12345;
    float acc = 0.f;
    for (uint32_t it = 0; it < iters; ++it) {
        const float4* p[4];
        for (int r = 0; r < 4; ++r) {
            state = state * 6364136223846793005ull + 1442695040888963407ull;
            p[r] = reinterpret_cast<const float4*>(rows + ((state >> 20) % n_rows) * row_bytes);
        }
        for (uint32_t c = lane; c < row_bytes / 16; c += 32) {
            float4 v0 = __ldg(p[0] + c), v1 = __ldg(p[1] + c), v2 = __ldg(p[2] + c), v3 = __ldg(p[3] + c);
            acc += v0.x + v1.y + v2.z + v3.w;
        }
    }
    if (acc == 123.456f) *sink = acc;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    const size_t bytes = 3ull << 30;  // 3 GiB table >> L2
    uint8_t* rows;
    float* sink;
    cudaMalloc(&rows, bytes);
    cudaMalloc(&sink, 4);
    cudaMemset(rows, 1, bytes);
    cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    std::printf("{\"device\": \"%s\", \"sms\": %d, \"results\": [\n", prop.name, sms);
    struct Cfg { uint32_t row, warps_per_sm, slots, touch; };
    std::vector<Cfg> cfgs;
    for (uint32_t touch : {1u})
        for (uint32_t row : {3072u, 512u, 128u})
            for (uint32_t wps : {4u, 8u, 12u, 16u, 24u, 32u})
                for (uint32_t slots : {2u, 4u, 8u, 16u, 32u}) {
                    size_t per_warp = ((size_t)slots * row + slots * 8 + 127) & ~(size_t)127;
                    if (per_warp * wps > 224 * 1024) continue;
                    cfgs.push_back({row, wps, slots, touch});
                }
    bool first = true;
    for (auto c : cfgs) {
        const int wpb = 4;
        size_t per_warp = ((size_t)c.slots * c.row + c.slots * 8 + 127) & ~(size_t)127;
        size_t smem = per_warp * wpb;
        int blocks = sms * c.warps_per_sm / wpb;
        uint64_t n_rows = bytes / c.row;
        uint64_t total_target = 24ull << 30;  // ~24 GiB moved per measurement
        uint32_t iters = (uint32_t)(total_target / ((uint64_t)blocks * wpb * c.slots * c.row));
        if (iters < 4) iters = 4;
        // pad smem so exactly warps_per_sm/wpb blocks are resident
        size_t pad = (size_t)(224 * 1024) / (c.warps_per_sm / wpb);
        pad = pad / 128 * 128;
        if (pad > smem) smem = pad;
        gather_kernel<<<blocks, wpb * 32, smem>>>(rows, n_rows, c.row, c.slots, 2, c.touch, sink);
        cudaEventRecord(e0);
        gather_kernel<<<blocks, wpb * 32, smem>>>(rows, n_rows, c.row, c.slots, iters, c.touch, sink);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        double gb = (double)blocks * wpb * c.slots * c.row * iters / 1e9;
        std::printf("%s {\"kind\": \"bulk\", \"row_bytes\": %u, \"warps_per_sm\": %u, \"slots\": %u, \"inflight_kb_per_sm\": %.1f, \"gbs\": %.1f, \"err\": \"%s\"}",
                    first ? "" : ",\n", c.row, c.warps_per_sm, c.slots, c.warps_per_sm * c.slots * c.row / 1024.0, gb / (ms / 1e3),
                    err == cudaSuccess ? "" : cudaGetErrorString(err));
        first = false;
        std::fflush(stdout);
    }
    for (uint32_t row : {3072u, 512u}) {
        for (int wps : {16, 32, 64}) {
            int blocks = sms * wps / 4;
            uint64_t n_rows = bytes / row;
            uint32_t iters = (uint32_t)((24ull << 30) / ((uint64_t)blocks * 4 * 4 * row));
            gather_ldg_kernel<<<blocks, 128>>>(rows, n_rows, row, 2, sink);
            cudaEventRecord(e0);
            gather_ldg_kernel<<<blocks, 128>>>(rows, n_rows, row, iters, sink);
            cudaEventRecord(e1);
            cudaDeviceSynchronize();
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            double gb = (double)blocks * 4 * 4 * row * iters / 1e9;
            std::printf(",\n {\"kind\": \"ldg\", \"row_bytes\": %u, \"warps_per_sm\": %d, \"gbs\": %.1f}", row, wps, gb / (ms / 1e3));
        }
    }
    std::printf("\n]}\n");
    return 0;
}
