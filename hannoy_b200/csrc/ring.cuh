// ring.cuh — per-warp row ring fed by the bulk async-copy engine (TMA 1-D, `cp.async.bulk`, SASS UBLKCP).
//
// A neighbour row is one contiguous, 16-byte-aligned run of `row_stride` bytes in HBM (common.h RowKind).
// Instead of pulling it through registers with dependent LDG.128 rounds, the lane that owns the neighbour
// posts ONE bulk copy global -> shared memory that completes on an mbarrier (complete_tx::bytes).  A warp
// keeps `slots` rows in flight without spending a register on them; the distance loop then reads the row
// from shared memory.  Each warp owns its ring and its barriers: no CTA-wide synchronisation anywhere.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace hb {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// make barrier initialisation visible to the async proxy before the first copy is posted
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// the same without release semantics: the arrival orders nothing this thread wrote (the waiter is the posting warp itself and
// the data it waits for comes from the copy engine, ordered by complete_tx), so it need not wait for earlier stores
__device__ __forceinline__ void mbar_expect_tx_relaxed(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.relaxed.cta.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
// rows are streamed once per visit: keep them from evicting the adjacency / visited / header lines in L2
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// ... except, optionally, the rows of the upper layers: every query descends through the same few thousand nodes
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar), "l"(policy)
        : "memory");
}

struct RowRing {
    uint32_t data = 0;        // shared-window address of slot 0
    uint32_t bars = 0;        // shared-window address of barrier 0 (8 bytes each)
    const uint8_t* ptr = nullptr;  // generic pointer to slot 0 (for the distance loop)
    uint32_t slots = 0;       // multiple of ROW_GROUP
    uint32_t stride = 0;      // bytes between slots (>= row bytes, multiple of 16)
    uint32_t phase = 0;       // bit i = parity the next wait on slot i must observe
    uint64_t policy = 0;

    __device__ __forceinline__ void post(uint32_t slot, const void* row, uint32_t bytes) const {
        uint32_t bar = bars + slot * 8;
        mbar_expect_tx_relaxed(bar, bytes);
        bulk_g2s(data + slot * stride, row, bytes, bar, policy);
    }
    // Four rows in ONE instruction (TMA tile::gather4, sm_100): rows r0..r3 of the 2-D view of the row array described by
    // `tmap` (box = one whole row) land in slots slot0..slot0+3 and complete on slot0's barrier.  A row index past the end
    // reads as zeros.  The bulk-copy engine is paced per instruction, so short rows go four times as fast this way.
    __device__ __forceinline__ void post_gather4(uint32_t slot0, const void* tmap, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) const {
        const uint32_t bar = bars + slot0 * 8;
        mbar_expect_tx_relaxed(bar, 4 * stride);
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6, %7}], [%2], %8;"
            ::"r"(data + slot0 * stride), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(0), "r"((int)r0), "r"((int)r1), "r"((int)r2), "r"((int)r3), "l"(policy)
            : "memory");
    }
    __device__ __forceinline__ void wait(uint32_t slot) {
        mbar_wait(bars + slot * 8, (phase >> slot) & 1u);
        phase ^= 1u << slot;
    }
};

}  // namespace hb
