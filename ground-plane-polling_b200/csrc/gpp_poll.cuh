// The polling kernel: one warp owns one detection (or kDpw detections), its 32 lanes stride over the planes
// of the current database tile, tiles are streamed through shared memory by 1-D TMA bulk copies
// (cp.async.bulk + mbarrier, SASS UBLKCP) in a kStages-deep ring that runs continuously across detection
// groups, and the reference's two-pass "max votes -> masked argmin" (fit_road_planes.py:116-119) is folded
// into one streaming pass per lane followed by a warp reduction.
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

// ------------------------------------------------------------------ kernel arguments
template <class T>
struct PollArgs {
    const float *boxes;          // (n_det, 12)
    const float *dims;           // (n_det, 3)
    const int32_t *orient;       // (n_det)
    const float *pinv;           // (n_img, 4, 3)
    const void *planes;          // normalised DB, n_planes x T4
    int n_planes;
    int dets_per_image;          // D
    long long n_det;             // B * D
    T *keypoints;                // (n_det, 4, 3)
    T *keyplanes;                // (n_det, 4)
    T *residuals;                // (n_det)
    long long *best;             // (n_det) or nullptr
    // optional work list (VERIFIED mode, second pass): process det_list[0 .. *det_count) instead of 0 .. n_det
    const long long *det_list;
    const unsigned int *det_count;
};

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

// ------------------------------------------------------------------ the kernel
// partial result of one warp when the warps of a CTA split the planes of ONE detection (small batches)
template <class T>
struct WarpPartial {
    T r;
    int M, idx;
};

// kWarps warps per CTA, kDpw detections per warp in flight, kTile planes per smem tile, kStages ring depth.
// kSplit: small-batch variant -- the CTA works on one detection, warp w takes the rows r = w (mod kWarps) of
// every tile and the partial arg-mins are merged through shared memory (kDpw must be 1).
template <class P, int kWarps, int kDpw, int kTile, int kStages, bool kSplit = false>
// the one-detection-per-warp fp32 kernel is capped at 64 registers: four CTAs per SM (measured best once the second
// half of most hypotheses is skipped; the two-detections-per-warp variant needs 107 registers and is slower now)
#define GPP_EXACT_BOUNDS __launch_bounds__(kWarps * 32, (sizeof(typename P::T) == 4 && kDpw == 1) ? 4 : 1)
__global__ void GPP_EXACT_BOUNDS poll_kernel(const PollArgs<typename P::T> args) {
    typedef typename P::T T;
    typedef typename P::T4 T4;
    static_assert(!kSplit || kDpw == 1, "the split variant handles one detection per CTA");
    constexpr int kGroup = kSplit ? 1 : kWarps * kDpw;   // detections per CTA pass over the database
    constexpr int kRowStep = kSplit ? kWarps : 1;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    T4 *tiles = reinterpret_cast<T4 *>(smem_raw);                                  // kStages * kTile
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem_raw + sizeof(T4) * kStages * kTile);
    uint64_t *empty_bar = full_bar + kStages;
    WarpPartial<T> *partial = reinterpret_cast<WarpPartial<T> *>(empty_bar + kStages);   // [2][kWarps], kSplit only

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int N = args.n_planes;
    const int n_tiles = (N + kTile - 1) / kTile;
    const long long n_work = args.det_list ? (long long)(*args.det_count) : args.n_det;
    const long long n_groups = (n_work + kGroup - 1) / kGroup;
    const long long my_groups = (n_groups > blockIdx.x) ? (n_groups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long total_tiles = my_groups * n_tiles;
    const T4 *gplanes = reinterpret_cast<const T4 *>(args.planes);

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](long long it) {                 // producer: thread 0 only
        const int s = int(it % kStages);
        const int t = int(it % n_tiles);
        const int cnt = min(kTile, N - t * kTile);
        const uint32_t bytes = uint32_t(cnt) * uint32_t(sizeof(T4));
        mbar_arrive_expect_tx(&full_bar[s], bytes);
        tma_load_1d(tiles + size_t(s) * kTile, gplanes + size_t(t) * kTile, bytes, &full_bar[s]);
    };
    if (threadIdx.x == 0) {
        const long long pre = total_tiles < kStages ? total_tiles : kStages;
        for (long long it = 0; it < pre; ++it) issue(it);
    }

    Detection<P> det[kDpw];
    LaneState<T> st[kDpw];
    long long det_id[kDpw];
    const T highest = P::highest();
    // Once a plane with six votes is known (warp-uniform m6), a plane matters only if its residual sum does not
    // exceed the warp's best six-vote residual wbest.  The sum of the three bottom-face residuals never exceeds
    // the full sum (rounded addition of non-negative terms is monotone; a NaN / inf sum never wins), so
    // (r1 + r2) + r3 > wbest for all 32 lanes ends the hypothesis after its first half: X_t, one division and
    // three square roots are skipped for about three of four rows, and nothing that is kept changes by a bit.
    bool m6[kDpw];
    T wbest[kDpw];

    long long it = 0;
    for (long long g = blockIdx.x; g < n_groups; g += gridDim.x) {
        // ---- per-detection prologue (warp-uniform; fit_road_planes.py:66-72, :80-83)
#pragma unroll
        for (int q = 0; q < kDpw; ++q) {
            long long m = kSplit ? g : g * kGroup + (long long)warp * kDpw + q;
            long long mm = m < n_work ? m : n_work - 1;                   // tail warps redo the last one
            if (args.det_list) mm = args.det_list[mm];
            det_id[q] = m < n_work ? mm : -1;
            load_detection<P, typename ExactOf<P>::type>(det[q], args.boxes + 12 * mm, args.dims + 3 * mm, __ldg(args.orient + mm),
                                 args.pinv + 12 * (mm / args.dets_per_image));
            st[q].reset(highest);
            m6[q] = false;
            wbest[q] = highest;
        }
        // ---- stream the whole database through the ring
        for (int t = 0; t < n_tiles; ++t, ++it) {
            const int s = int(it % kStages);
            const uint32_t parity = uint32_t((it / kStages) & 1);
            mbar_wait(&full_bar[s], parity);
            const T4 *tile = tiles + size_t(s) * kTile;
            const int cnt = min(kTile, N - t * kTile);
            const int base = t * kTile;
            const int full_rows = cnt >> 5;
            int r = kSplit ? warp : 0;
#pragma unroll 1
            for (; r < full_rows; r += kRowStep) {
                const int jj = (r << 5) + lane;
                const T4 pl = tile[jj];
#pragma unroll
                for (int q = 0; q < kDpw; ++q) {
                    T X[4][3];
                    int V; T R; bool zneg;
                    if (m6[q]) {
                        T rb[3];
                        hypothesis_bottom<P>(det[q], pl.x, pl.y, pl.z, pl.w, X, rb, zneg);
                        const T S3 = P::add(P::add(rb[0], rb[1]), rb[2]);
                        if (!__any_sync(0xffffffffu, !(S3 > wbest[q]))) continue;
                        hypothesis_top<P>(det[q], pl.x, pl.y, pl.z, X, rb, V, R);
                        st[q].update(V, R, zneg, base + jj, highest);
                        wbest[q] = warp_min_value(st[q].M == 6 ? st[q].bestR : highest);
                    } else {
                        hypothesis<P>(det[q], pl.x, pl.y, pl.z, pl.w, X, V, R, zneg);
                        st[q].update(V, R, zneg, base + jj, highest);
                    }
                }
                if (((r / kRowStep) & 3) == 3) {
#pragma unroll
                    for (int q = 0; q < kDpw; ++q)
                        if (!m6[q] && __reduce_max_sync(0xffffffffu, st[q].M) == 6) {
                            m6[q] = true;
                            wbest[q] = warp_min_value(st[q].M == 6 ? st[q].bestR : highest);
                        }
                }
            }
            if ((cnt & 31) != 0 && (!kSplit || r == full_rows)) {   // ragged last row of the last tile
                const int jj = (r << 5) + lane;
                if (jj < cnt) {
                    const T4 pl = tile[jj];
#pragma unroll
                    for (int q = 0; q < kDpw; ++q) {
                        T X[4][3];
                        int V; T R; bool zneg;
                        hypothesis<P>(det[q], pl.x, pl.y, pl.z, pl.w, X, V, R, zneg);
                        st[q].update(V, R, zneg, base + jj, highest);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);            // this warp is done with the slot
            // producer duty, skewed by one tile so that warp 0 rarely waits for the slowest warp
            if (threadIdx.x == 0 && it >= 1) {
                const long long prev = it - 1;
                if (prev + kStages < total_tiles) {
                    mbar_wait(&empty_bar[prev % kStages], uint32_t((prev / kStages) & 1));
                    issue(prev + kStages);
                }
            }
            __syncwarp();
        }
        // ---- per-detection epilogue: warp reduction, lazy first-masked search, exact recompute, store
#pragma unroll
        for (int q = 0; q < kDpw; ++q) {
            int Mw = __reduce_max_sync(0xffffffffu, st[q].M);
            T r = (st[q].M == Mw) ? st[q].bestR : highest;
            int idx = st[q].bestIdx;
            r = warp_min_first(r, idx);
            if (kSplit) {
                // merge the warps' partial results (double-buffered by group parity: one barrier per group)
                WarpPartial<T> *buf = partial + ((g / gridDim.x) & 1) * kWarps;
                if (lane == 0) { buf[warp].r = r; buf[warp].M = Mw; buf[warp].idx = idx; }
                __syncthreads();
                if (warp != 0) continue;
                const int Ml = lane < kWarps ? buf[lane].M : -1;
                Mw = __reduce_max_sync(0xffffffffu, Ml);
                r = (lane < kWarps && Ml == Mw) ? buf[lane].r : highest;
                idx = lane < kWarps ? buf[lane].idx : 0;
                r = warp_min_first(r, idx);
            }
            const bool have_cand = r < highest;
            bool sentinel = false;
            if (!(r < T(100))) {
                // the constant 100 carried by masked planes may win: find the first masked plane
                int first_masked = -1;
                for (int j0 = 0; j0 < N && first_masked < 0; j0 += 32) {
                    const int j = j0 + lane;
                    bool masked = false;
                    if (j < N) {
                        const T4 pl = gplanes[j];
                        T X[4][3];
                        int V; T R; bool zneg;
                        hypothesis<P>(det[q], pl.x, pl.y, pl.z, pl.w, X, V, R, zneg);
                        masked = (V < Mw) || zneg;
                    }
                    const unsigned b = __ballot_sync(0xffffffffu, masked);
                    if (b) first_masked = j0 + __ffs(b) - 1;
                }
                if (first_masked >= 0) {
                    if (!have_cand || T(100) < r || (T(100) == r && first_masked < idx)) {
                        sentinel = true;
                        idx = first_masked;
                    }
                } else if (!have_cand) {
                    idx = 0;                                      // nothing compares below `highest`
                }
            }
            if (det_id[q] >= 0 && lane == 0) {
                // recompute the winner in the exact arithmetic of this scalar type (fit_road_planes.py:122-137)
                typedef typename ExactOf<P>::type E;
                const T4 pl = gplanes[idx];
                Detection<E> de;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    de.dl[i] = det[q].dl[i]; de.dm[i] = det[q].dm[i];
                    de.dr[i] = det[q].dr[i]; de.dt[i] = det[q].dt[i];
                }
#pragma unroll
                for (int i = 0; i < 6; ++i) de.td[i] = det[q].td[i];
                T X[4][3];
                int V; T R; bool zneg;
                hypothesis<E>(de, pl.x, pl.y, pl.z, pl.w, X, V, R, zneg);
                const T rr = sentinel ? T(100) : R;
                const long long m = det_id[q];
                T *kp = args.keypoints + 12 * m;
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int i = 0; i < 3; ++i) kp[3 * k + i] = X[k][i];
                T *kpl = args.keyplanes + 4 * m;
                kpl[0] = pl.x; kpl[1] = pl.y; kpl[2] = pl.z; kpl[3] = pl.w;
                args.residuals[m] = E::div(rr, T(6));
                if (args.best) args.best[m] = idx;
            }
        }
    }
}

}  // namespace gpp
