// Shared device helpers of the polling kernel (gpp_poll3.cuh): mbarrier / 1-D TMA bulk-copy wrappers, the per-lane
// streaming selection state that folds the reference's two-pass "max votes -> masked argmin"
// (fit_road_planes.py:116-119) into one pass, and the warp reductions that go with it.
#pragma once
#include "gpp_math.cuh"

#ifndef GPP_WAIT_SLEEP_NS
#define GPP_WAIT_SLEEP_NS 200
#endif

namespace gpp {

// ------------------------------------------------------------------ mbarrier / TMA (1-D bulk) helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// Waiting warps should not steal issue slots from the working ones (the kernels are issue / register-file
// bound): after a failed try_wait the warp sleeps a little before polling again.  (A long suspend-time hint on
// try_wait itself was measured to delay the wake-up far more than it saves.)
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
#if GPP_WAIT_SLEEP_NS == 0
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "GPP_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra GPP_WAIT_%=;\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
    return;
#endif
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra GPP_DONE_%=;\n"
        "GPP_WAIT_%=:\n"
        "nanosleep.u32 %2;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra GPP_WAIT_%=;\n"
        "GPP_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(GPP_WAIT_SLEEP_NS)
        : "memory");
}
// global -> shared bulk copy executed by the TMA unit; completion is signalled on `bar` (complete_tx).
// The slot being overwritten was last READ through the generic proxy by the consumer warps; their release is
// observed by the producer through the empty mbarrier, and the proxy fence orders those generic-proxy
// accesses before this async-proxy write.
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                            uint64_t *bar) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ------------------------------------------------------------------ per-lane streaming selection state
// A plane is a *candidate* iff votes == (final) max votes and !(z_dir_check < 0); every other plane carries
// the constant 100 at its index.  Per lane we keep the running max votes M and the best candidate seen
// under that M (strict '<' in increasing plane order = first occurrence); when M grows the candidates seen
// so far become masked, so the best is reset.  The "first masked index" needed when the sentinel wins is
// found lazily in the epilogue (it is almost always among the first few planes).
template <class T>
struct LaneState {
    int M;
    T bestR;
    int bestIdx;
    __device__ __forceinline__ void reset(T highest) { M = -1; bestR = highest; bestIdx = 0; }
    __device__ __forceinline__ void update(int V, T R, bool zneg, int j, T highest) {
        const bool grow = V > M;
        const T cur = grow ? highest : bestR;
        const bool better = (V >= M) && !zneg && (R < cur);    // NaN / >= highest never wins
        bestR = better ? R : cur;
        bestIdx = better ? j : bestIdx;
        M = max(M, V);
    }
};

__device__ __forceinline__ float warp_min_first(float v, int &idx) {
    // v >= +0 and never NaN (strict '<' from FLT_MAX): the bit pattern orders like the value
    unsigned bits = __float_as_uint(v);
    unsigned mn = __reduce_min_sync(0xffffffffu, bits);
    unsigned cand = (bits == mn) ? (unsigned)idx : 0x7fffffffu;
    idx = (int)__reduce_min_sync(0xffffffffu, cand);
    return __uint_as_float(mn);
}
__device__ __forceinline__ double warp_min_first(double v, int &idx) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, v, o);
        int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        bool take = (ov < v) || (ov == v && oi < idx);
        v = take ? ov : v;
        idx = take ? oi : idx;
    }
    return v;
}

// warp-wide minimum of a non-negative, non-NaN value
__device__ __forceinline__ float warp_min_value(float v) {
    return __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(v)));
}
__device__ __forceinline__ double warp_min_value(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace gpp
