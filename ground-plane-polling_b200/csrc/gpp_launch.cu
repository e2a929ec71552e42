// Kernel instantiations and launch configuration of the polling kernels (gpp_poll2.cuh: packed-pair kernels of
// the FAST and VERIFIED modes; gpp_poll.cuh: scalar kernel of the EXACT and FP64 modes).
#ifdef GPP_STATS
#include <cstdio>
#endif
#include "../../include/gpp_debug.h"
#include "gpp_internal.h"

#ifndef GPP_DEFAULT_VARIANT_FAST
#define GPP_DEFAULT_VARIANT_FAST 1
#endif
#ifndef GPP_DEFAULT_VARIANT_VERIFIED
#define GPP_DEFAULT_VARIANT_VERIFIED 2
#endif

namespace gpp {

// CTA shape: 8 warps, 1024-plane tiles (32 KB fp32 pairs / 32 KB fp64), 3-stage TMA ring.
#ifndef GPP_WARPS
#define GPP_WARPS 8
#endif
constexpr int kWarps = GPP_WARPS;
#ifdef GPP_MB_OVERRIDE
#define GPP_MB(x) GPP_MB_OVERRIDE            /* experiments: resident CTAs per SM given directly */
#else
#define GPP_MB(x) ((x) * 8 / GPP_WARPS)   /* resident CTAs per SM for x CTAs of 8 warps */
#endif
#ifndef GPP_TILE
#define GPP_TILE 1024
#endif
#ifndef GPP_STAGES
#define GPP_STAGES 3
#endif
constexpr int kTile32 = GPP_TILE, kTile64 = 512;
constexpr int kStages = GPP_STAGES;
// the packed kernels run a deeper ring: with per-warp claiming (kFree) a fast warp can bank up to three tiles of
// lead over the slowest warp of its CTA (4 x 16 KB tiles + queues = 69 KB per CTA, three CTAs per SM)
#ifndef GPP_STAGES2
#define GPP_STAGES2 4
#endif
constexpr int kStages2 = GPP_STAGES2;

// pair-interleaved, padded copy of the normalised fp32 database: pair p = planes (2p, 2p+1) stored as
// {a0,a1,b0,b1,c0,c1,d0,d1}; planes past N-1 are copies of plane N-1 (same score, higher index: they can
// never win a first-occurrence arg-min and do not change max-votes)
__global__ void interleave_pairs_kernel(const float4 *__restrict__ planes, int n, int n_pairs_padded,
                                        float *__restrict__ pairs) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs_padded) return;
    const float4 a = planes[min(2 * p, n - 1)], b = planes[min(2 * p + 1, n - 1)];
    float4 *out = reinterpret_cast<float4 *>(pairs + 8 * (size_t)p);
    out[0] = make_float4(a.x, b.x, a.y, b.y);
    out[1] = make_float4(a.z, b.z, a.w, b.w);
}

int build_pairs(gpp_handle *h, cudaStream_t s) {
    const int np = h->n_pairs_padded;
    interleave_pairs_kernel<<<(np + 127) / 128, 128, 0, s>>>(h->d_planes32, h->n_planes, np,
                                                            reinterpret_cast<float *>(h->d_pairs));
    h->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(GPP_ECUDA, "interleave_pairs_kernel launch: %s", cudaGetErrorString(e));
    return GPP_OK;
}

// ---------------------------------------------------------------------------------------------------
// Rows that repeat the previous row of their image bit for bit (FilterDetections pads every image to 100
// rows with -1, layers/filter_detections.py:170-185) are polled once: mark -> compact -> poll the unique
// rows -> copy.  Identical inputs give identical outputs, so this changes nothing but the cost.
// ---------------------------------------------------------------------------------------------------
__global__ void mark_unique_kernel(const float *__restrict__ boxes, const float *__restrict__ dims,
                                   const int32_t *__restrict__ orient, long long n_det, int D,
                                   unsigned char *__restrict__ unique, long long *__restrict__ list,
                                   unsigned int *__restrict__ count) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool same = true;                                  // out-of-range threads contribute nothing
    if (m < n_det) {
        same = (m % D) != 0;
        if (same) {
            const unsigned int *b = reinterpret_cast<const unsigned int *>(boxes) + 12 * m;
            const unsigned int *d = reinterpret_cast<const unsigned int *>(dims) + 3 * m;
#pragma unroll
            for (int i = 0; i < 12; ++i) same = same && (b[i] == b[i - 12]);
#pragma unroll
            for (int i = 0; i < 3; ++i) same = same && (d[i] == d[i - 3]);
            same = same && (orient[m] == orient[m - 1]);
        }
        unique[m] = same ? 0 : 1;
    }
    // block-level compaction: one atomic per block reserves a contiguous range of the list
    __shared__ unsigned int warp_base[9];
    const unsigned int ballot = __ballot_sync(0xffffffffu, !same);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_base[warp] = __popc(ballot);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int total = 0;
        for (int w = 0; w < 8; ++w) { const unsigned int c = warp_base[w]; warp_base[w] = total; total += c; }
        warp_base[8] = total ? atomicAdd(count, total) : 0u;
    }
    __syncthreads();
    if (!same) list[warp_base[8] + warp_base[warp] + __popc(ballot & ((1u << lane) - 1u))] = m;
}

template <class T>
__global__ void copy_duplicates_kernel(const unsigned char *__restrict__ unique, long long n_det, T *keypoints,
                                       T *keyplanes, T *residuals, long long *best) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_det || unique[m]) return;
    long long r = m - 1;
    while (!unique[r]) --r;                     // row 0 of every image is unique, so this stays in the image
#pragma unroll
    for (int i = 0; i < 12; ++i) keypoints[12 * m + i] = keypoints[12 * r + i];
#pragma unroll
    for (int i = 0; i < 4; ++i) keyplanes[4 * m + i] = keyplanes[4 * r + i];
    residuals[m] = residuals[r];
    if (best) best[m] = best[r];
}

template <class K>
static int configure_kernel(K kernel, size_t smem, int *occ) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(GPP_ECUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kernel, kWarps * 32, smem);
    if (e != cudaSuccess) return set_error(GPP_ECUDA, "occupancy query: %s", cudaGetErrorString(e));
    if (*occ < 1) return set_error(GPP_ECUDA, "polling kernel does not fit on an SM");
    return GPP_OK;
}

static long long grid_for(const gpp_handle *h, long long n_groups, int occ) {
    const int per_sm = h->force_ctas_per_sm > 0 ? h->force_ctas_per_sm : occ;
    long long grid = (long long)h->sm_count * per_sm;
    if (grid > n_groups) grid = n_groups;
    return grid < 1 ? 1 : grid;
}

constexpr size_t kSmem2 = size_t(32) * kStages2 * (kTile32 / 2) + 2 * kStages2 * sizeof(uint64_t) +
                          sizeof(int) * kWarps * kVerifyQueue + 2 * kWarps * sizeof(WarpPartial<float>) +
                          2 * sizeof(unsigned int) + sizeof(float) * 20 * kWarps + 16;   // + need_until, active (kFree)
constexpr size_t kSmem64 = sizeof(double4) * kStages * kTile64 + 2 * kStages * sizeof(uint64_t) +
                           2 * kWarps * sizeof(WarpPartial<double>);
#define GPP_K_F64 poll_kernel<ExactF64, kWarps, 1, kTile64, kStages>
#define GPP_K_F64_SPLIT poll_kernel<ExactF64, kWarps, 1, kTile64, kStages, true>

// EXACT fp32 mode runs the scalar kernel (gpp_poll.cuh): ptxas 12.9 contracts a packed mul.rn.f32x2 feeding a
// packed add.rn.f32x2 into FFMA2 even with explicit .rn (checked in SASS), which would break the FMA-free
// canonical arithmetic, so the packed kernel is used for the FAST mode only.
#define GPP_K_EXACT1 poll_kernel<ExactF32, kWarps, 1, kTile32, kStages>
#define GPP_K_EXACT2 poll_kernel<ExactF32, kWarps, 2, kTile32, kStages>
#define GPP_K_EXACT_SPLIT poll_kernel<ExactF32, kWarps, 1, kTile32, kStages, true>
#define GPP_K_FAST_SPLIT poll2_kernel<PackFast, kWarps, kTile32, kStages2, GPP_MB(3), 0, true>
#define GPP_K_VERIFIED_SPLIT poll2_kernel<PackFast, kWarps, kTile32, kStages2, GPP_MB(3), 1, true>
constexpr size_t kSmem1 = sizeof(float4) * kStages * kTile32 + 2 * kStages * sizeof(uint64_t) +
                          2 * kWarps * sizeof(WarpPartial<float>);

// Resident-database kernel (gpp_poll3.cuh): one persistent CTA of kWarps3 warps per SM
// 32 warps at 64 registers (no spills once the rarely used detection constants are parked in shared memory): measured
// 4.84e11 hypotheses/s (VERIFIED, C4) against 4.77e11 with 28 warps / 72 registers and 4.49e11 with 24 / 80
#ifndef GPP_WARPS3
#define GPP_WARPS3 32
#endif
constexpr int kWarps3 = GPP_WARPS3;
typedef void (*Poll3Fn)(const PollArgs3);
static Poll3Fn poll3_variant(bool verified, bool seg) {
    if (verified) return seg ? poll3_kernel<kWarps3, 1, true> : poll3_kernel<kWarps3, 1, false>;
    return seg ? poll3_kernel<kWarps3, 0, true> : poll3_kernel<kWarps3, 0, false>;
}

static int configure_poll3(gpp_handle *h) {
    int optin = 0;
    cudaError_t e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);
    if (e != cudaSuccess) return set_error(GPP_ECUDA, "shared memory query: %s", cudaGetErrorString(e));
    cudaFuncAttributes fa;
    size_t fixed = 0;
    for (int v = 0; v < 4; ++v) {
        e = cudaFuncGetAttributes(&fa, poll3_variant(v & 1, v & 2));
        if (e != cudaSuccess) return set_error(GPP_ECUDA, "cudaFuncGetAttributes: %s", cudaGetErrorString(e));
        if (fa.sharedSizeBytes > fixed) fixed = fa.sharedSizeBytes;
    }
    const long long room = (long long)optin - (long long)fixed - (long long)smem3_bytes(kWarps3, 0);
    if (room < 0) return set_error(GPP_ECUDA, "resident polling kernel does not fit on an SM");
    h->resident_cap_rows = (int)(room / 1024);
    const int max_dyn = (int)smem3_bytes(kWarps3, h->resident_cap_rows);
    for (int v = 0; v < 4 && e == cudaSuccess; ++v)
        e = cudaFuncSetAttribute(poll3_variant(v & 1, v & 2), cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn);
    if (e != cudaSuccess) return set_error(GPP_ECUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    // scratch of segmented detections: segments are only used below 3 detections per resident warp
    const long long slots = (long long)h->sm_count * kWarps3;
    h->seg_det_cap = 3 * slots;
    h->seg_items_cap = 6 * slots + 64;
    for (int i = 0; i < gpp_handle::kSlots3 && e == cudaSuccess; ++i) {
        gpp_handle::Slot3 &w = h->slot3[i];
        e = cudaMalloc(&w.claim, 2 * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaMemset(w.claim, 0, 2 * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaMalloc(&w.partials, sizeof(SegPartial) * (size_t)h->seg_items_cap);
        if (e == cudaSuccess) e = cudaMalloc(&w.seg_arrived, sizeof(unsigned int) * (size_t)h->seg_det_cap);
        if (e == cudaSuccess) e = cudaMemset(w.seg_arrived, 0, sizeof(unsigned int) * (size_t)h->seg_det_cap);
        if (e == cudaSuccess) e = cudaMalloc(&w.seg_best, sizeof(unsigned long long) * (size_t)h->seg_det_cap);
        if (e == cudaSuccess) e = cudaMemset(w.seg_best, 0, sizeof(unsigned long long) * (size_t)h->seg_det_cap);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&w.done, cudaEventDisableTiming);
    }
    if (e != cudaSuccess) return set_error(GPP_ECUDA, "poll3 scratch allocation: %s", cudaGetErrorString(e));
    return GPP_OK;
}

void release_poll3(gpp_handle *h) {
    for (auto &w : h->slot3) {
        cudaFree(w.claim); cudaFree(w.partials); cudaFree(w.seg_arrived); cudaFree(w.seg_best);
        if (w.done) cudaEventDestroy(w.done);
        w = gpp_handle::Slot3();
    }
}

// Schedule of one call (see the header of gpp_poll3.cuh).  Segments: below three detections per resident warp every
// detection is cut into plane segments so that the work items still fill the machine about three times over (the
// last wave is then short whatever the batch size).  Residency: staging up to 216 KB per SM pays as soon as every
// warp polls a few items; a call with fewer items than that streams every row from L2 and starts at once.
static int launch_poll3(gpp_handle *h, const PollArgs<float> &a, int mode, cudaStream_t s) {
    gpp_handle::Slot3 &w = h->slot3[h->next_slot3++ % gpp_handle::kSlots3];
    cudaError_t e = cudaSuccess;
    if (w.used) e = cudaStreamWaitEvent(s, w.done, 0);          // previous user of this slot (another stream)
    if (e != cudaSuccess) return set_error(GPP_ECUDA, "poll3 slot: %s", cudaGetErrorString(e));
    PollArgs3 b;
    b.boxes = a.boxes; b.dims = a.dims; b.pinv = a.pinv; b.orient = a.orient;
    b.pairs = h->d_pairs; b.planes = h->d_planes32; b.n_planes = h->n_planes;
    b.n_pairs_padded = h->n_pairs_padded; b.dets_per_image = a.dets_per_image; b.n_det = a.n_det;
    b.keypoints = a.keypoints; b.keyplanes = a.keyplanes; b.residuals = a.residuals; b.best = a.best;
    b.claim = w.claim; b.partials = w.partials; b.seg_arrived = w.seg_arrived; b.seg_best = w.seg_best;
    const int NR = h->n_pairs_padded / 32;
    const long long slots = (long long)h->sm_count * kWarps3;
    int n_seg = 1;
    if (h->force_seg > 0) n_seg = h->force_seg;
    else if (a.n_det < 3 * slots) n_seg = (int)((3 * slots + a.n_det - 1) / a.n_det);
    if (n_seg > 16 && h->force_seg <= 0) n_seg = 16;        // measured: a single image is fastest with 16 segments
    if (n_seg > 32) n_seg = 32;
    if (n_seg > NR) n_seg = NR;
    if (a.n_det > h->seg_det_cap || a.n_det * n_seg > h->seg_items_cap) n_seg = 1;
    b.rows_per_seg = (NR + n_seg - 1) / n_seg;
    b.n_seg = (NR + b.rows_per_seg - 1) / b.rows_per_seg;
    const long long n_items = a.n_det * b.n_seg;
    int res = n_items >= 2 * slots ? h->resident_cap_rows : 0;
    if (h->force_resident >= 0) res = h->force_resident;
    if (res > h->resident_cap_rows) res = h->resident_cap_rows;
    if (res > NR) res = NR;
    b.resident_rows = res;
    long long grid = h->sm_count;
    if (grid > n_items) grid = n_items;
    const size_t smem = smem3_bytes(kWarps3, res);
    poll3_variant(mode == GPP_MODE_VERIFIED, b.n_seg > 1)<<<(unsigned)grid, kWarps3 * 32, smem, s>>>(b);
#ifdef GPP_STATS
    if (mode == GPP_MODE_VERIFIED) {
        unsigned long long st[8], zero[8] = {0};
        cudaDeviceSynchronize();
        cudaMemcpyFromSymbol(st, g_stats3, sizeof(st));
        cudaMemcpyToSymbol(g_stats3, zero, sizeof(zero));
        const double rows = double(st[0] + st[2]), items = double(st[5] ? st[5] : 1);
        fprintf(stderr, "[gpp stats3] items %llu (identical-rays %llu) seg %d resident %d: rows/item %.1f, all-six %.1f%% (past stage 1: %.1f%% "
                "of them), general %.1f%%, rows with survivors %.2f%%; exact verifications/item %.1f, flushes/item %.2f\n",
                st[5], st[6], b.n_seg, res, rows / items, 100.0 * st[0] / (rows ? rows : 1), 100.0 * st[1] / (st[0] ? st[0] : 1),
                100.0 * st[2] / (rows ? rows : 1), 100.0 * st[7] / (rows ? rows : 1), double(st[3]) / items, double(st[4]) / items);
    }
#endif
    h->launches += 1;
    w.used = true;
    e = cudaEventRecord(w.done, s);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(GPP_ECUDA, "poll3 kernel launch: %s", cudaGetErrorString(e));
    return GPP_OK;
}

// FAST kernel variants: v = 0..2 <-> __launch_bounds__(256, 2 / 3 / 4) i.e. <= 128 / 80 / 64 registers
typedef void (*Poll2Fn)(const PollArgs2<float>);
static Poll2Fn fast_variant(int v) {
    switch (v) {
        case 0: return poll2_kernel<PackFast, kWarps, kTile32, kStages2, GPP_MB(2)>;
        case 1: return poll2_kernel<PackFast, kWarps, kTile32, kStages2, GPP_MB(3)>;
        default: return poll2_kernel<PackFast, kWarps, kTile32, kStages2, GPP_MB(4)>;
    }
}
static Poll2Fn verified_variant(int v) {
    switch (v) {
        case 0: return poll2_kernel<PackFast, kWarps, kTile32, kStages2, GPP_MB(2), 1>;
        case 2: return poll2_kernel<PackFast, kWarps, kTile32, kStages2, GPP_MB(3), 1, false, true>;   // per-warp claiming (default)
        default: return poll2_kernel<PackFast, kWarps, kTile32, kStages2, GPP_MB(3), 1>;
    }
}

int configure_kernels(gpp_handle *h) {
    int rc;
    if ((rc = configure_kernel(GPP_K_EXACT1, kSmem1, &h->occ[0]))) return rc;
    if ((rc = configure_kernel(GPP_K_EXACT2, kSmem1, &h->occ[1]))) return rc;
    for (int v = 0; v < 3; ++v)
        if ((rc = configure_kernel(fast_variant(v), kSmem2, &h->occ2[v]))) return rc;
    for (int v = 0; v < 3; ++v)
        if ((rc = configure_kernel(verified_variant(v), kSmem2, &h->occ3[v]))) return rc;
    if ((rc = configure_kernel(GPP_K_F64, kSmem64, &h->occ[2]))) return rc;
    if ((rc = configure_kernel(GPP_K_EXACT_SPLIT, kSmem1, &h->occ_split[0]))) return rc;
    if ((rc = configure_kernel(GPP_K_FAST_SPLIT, kSmem2, &h->occ_split[1]))) return rc;
    if ((rc = configure_kernel(GPP_K_VERIFIED_SPLIT, kSmem2, &h->occ_split[2]))) return rc;
    if ((rc = configure_kernel(GPP_K_F64_SPLIT, kSmem64, &h->occ_split[3]))) return rc;
    if ((rc = configure_poll3(h))) return rc;
    return GPP_OK;
}

// Small batches: one detection per CTA, planes split over the warps (kSplit kernels).  The batch kernels need
// several groups of eight detections per resident CTA to fill the GPU; measured crossover on B200
// (scripts/gpu_small_batches.py, 100-detection images x 10k / 22k planes): ~50 images for VERIFIED, ~40 for FAST,
// ~22 for EXACT, i.e. about `per_sm` detections per SM.
static bool use_split(const gpp_handle *h, long long n_det, int per_sm) {
    if (h->force_split) return h->force_split > 0;
    return n_det < (long long)per_sm * h->sm_count;
}

// All work-list slots are (re)allocated together, so that a steady stream of calls never hits cudaMalloc (which
// is not stream-ordered and would stall the GPU inside the caller's timed region) after the first one.
static int grow_worklists(gpp_handle *h, long long n_det) {
    cudaError_t e = cudaDeviceSynchronize();      // earlier launches may still use the old buffers
    for (int i = 0; i < gpp_handle::kWorkSlots && e == cudaSuccess; ++i) {
        gpp_handle::WorkSlot &w = h->work[i];
        cudaFree(w.list); cudaFree(w.ulist); cudaFree(w.unique);
        w.list = w.ulist = nullptr; w.unique = nullptr; w.cap = 0; w.used = false;
        if (!w.done) e = cudaEventCreateWithFlags(&w.done, cudaEventDisableTiming);
        if (e == cudaSuccess && !w.count) e = cudaMalloc(&w.count, 4 * sizeof(unsigned int));
        if (e == cudaSuccess) e = cudaMalloc(&w.list, sizeof(long long) * (size_t)n_det);
        if (e == cudaSuccess) e = cudaMalloc(&w.ulist, sizeof(long long) * (size_t)n_det);
        if (e == cudaSuccess) e = cudaMalloc(&w.unique, (size_t)n_det);
        if (e == cudaSuccess) w.cap = n_det;
    }
    if (e != cudaSuccess) return set_error(GPP_ECUDA, "work list allocation: %s", cudaGetErrorString(e));
    return GPP_OK;
}

static int reserve_worklist(gpp_handle *h, long long n_det, cudaStream_t s, gpp_handle::WorkSlot **out) {
    if (n_det > h->work[0].cap) {
        int rc = grow_worklists(h, n_det + n_det / 4);
        if (rc) return rc;
    }
    gpp_handle::WorkSlot &w = h->work[h->next_work++ % gpp_handle::kWorkSlots];
    cudaError_t e = cudaSuccess;
    if (w.used) e = cudaStreamWaitEvent(s, w.done, 0);        // previous user of this slot
    if (e == cudaSuccess) e = cudaMemsetAsync(w.count, 0, 4 * sizeof(unsigned int), s);
    if (e != cudaSuccess) return set_error(GPP_ECUDA, "work list setup: %s", cudaGetErrorString(e));
    *out = &w;
    return GPP_OK;
}

// mark + compact the rows that have to be polled; returns the slot holding the lists
template <class T>
static int begin_unique(gpp_handle *h, const PollArgs<T> &a, cudaStream_t s, gpp_handle::WorkSlot **w) {
    int rc = reserve_worklist(h, a.n_det, s, w);
    if (rc) return rc;
    const int threads = 256;
    const long long blocks = (a.n_det + threads - 1) / threads;
    mark_unique_kernel<<<(unsigned)blocks, threads, 0, s>>>(a.boxes, a.dims, a.orient, a.n_det, a.dets_per_image,
                                                           (*w)->unique, (*w)->ulist, (*w)->count + 1);
    h->launches += 1;
    return GPP_OK;
}

template <class T>
static int end_unique(gpp_handle *h, const PollArgs<T> &a, cudaStream_t s, gpp_handle::WorkSlot *w) {
    const int threads = 256;
    const long long blocks = (a.n_det + threads - 1) / threads;
    copy_duplicates_kernel<T><<<(unsigned)blocks, threads, 0, s>>>(w->unique, a.n_det, a.keypoints, a.keyplanes,
                                                                  a.residuals, a.best);
    h->launches += 1;
    w->used = true;
    cudaError_t e = cudaEventRecord(w->done, s);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(GPP_ECUDA, "poll kernel launch: %s", cudaGetErrorString(e));
    return GPP_OK;
}

// ---------------------------------------------------------------------------------------------------
// test hook: per-hypothesis scores of one detection, computed by the very device functions the search
// loops call (hypothesis<ExactF32> / eval_pair<PackFast>), so tests can compare them with the oracle
// hypothesis by hypothesis
// ---------------------------------------------------------------------------------------------------
__global__ void scores_exact_kernel(PollArgs<float> a, int32_t *votes, float *resid, int32_t *zneg) {
    Detection<ExactF32> det;
    load_detection<ExactF32, ExactF32>(det, a.boxes, a.dims, a.orient[0], a.pinv);
    const float4 *pl = reinterpret_cast<const float4 *>(a.planes);
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < a.n_planes; j += gridDim.x * blockDim.x) {
        float X[4][3];
        int V; float R; bool z;
        hypothesis<ExactF32>(det, pl[j].x, pl[j].y, pl[j].z, pl[j].w, X, V, R, z);
        votes[j] = V; resid[j] = R; zneg[j] = z ? 1 : 0;
    }
}

template <bool kSix>
__global__ void scores_fast_kernel(PollArgs2<float> a, int32_t *votes, float *resid, int32_t *zneg, float *margin) {
    DetConst D;
    {
        Detection<ExactF32> det;
        load_detection<ExactF32, ExactF32>(det, a.boxes, a.dims, a.orient[0], a.pinv);
        for (int i = 0; i < 6; ++i) D.td[i] = det.td[i];
        fast_constants(D, det);
    }
    const ulonglong2 *pairs = reinterpret_cast<const ulonglong2 *>(a.pairs);
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; 2 * p < a.n_planes; p += gridDim.x * blockDim.x) {
        const ulonglong2 v0 = pairs[2 * p], v1 = pairs[2 * p + 1];
        PairResult h;
        eval_pair_fast<kSix, 1>(D, from_u64(v0.x), from_u64(v0.y), from_u64(v1.x), from_u64(v1.y), h);
        const f2 R = resid_sum(h);
        finalize_margin(h, R, D);
        const int V0 = votes_of(lo(h.r[0]), lo(h.r[1]), lo(h.r[2]), lo(h.r[3]), lo(h.r[4]), lo(h.r[5]));
        const int V1 = votes_of(hi(h.r[0]), hi(h.r[1]), hi(h.r[2]), hi(h.r[3]), hi(h.r[4]), hi(h.r[5]));
        // votes: fast count + 16 x the count that is possible within the margin (the VERIFIED filters' test);
        // zneg: fast z-check + 2 x "may pass the z-check within the margin"
        const f2 zhi = z_upper(h, D);
        votes[2 * p] = V0 + 16 * loose_votes(h, false); resid[2 * p] = lo(R);
        zneg[2 * p] = (lo(h.zc) < 0.0f ? 1 : 0) + (!(lo(zhi) < 0.0f) ? 2 : 0);
        if (margin) margin[2 * p] = lo(h.m);
        if (2 * p + 1 < a.n_planes) {
            votes[2 * p + 1] = V1 + 16 * loose_votes(h, true); resid[2 * p + 1] = hi(R);
            zneg[2 * p + 1] = (hi(h.zc) < 0.0f ? 1 : 0) + (!(hi(zhi) < 0.0f) ? 2 : 0);
            if (margin) margin[2 * p + 1] = hi(h.m);
        }
    }
}

// which == 3: stage 1 of the VERIFIED all-six phase -- resid = sum of the three bottom-face residuals (merged
// reciprocals, as in the kernel), margin = the bound the stage-1 test relies on: w ms + mc + 2^-20 S3
__global__ void scores_bottom_kernel(PollArgs2<float> a, int32_t *votes, float *resid, int32_t *zneg, float *margin) {
    DetConst D;
    {
        Detection<ExactF32> det;
        load_detection<ExactF32, ExactF32>(det, a.boxes, a.dims, a.orient[0], a.pinv);
        for (int i = 0; i < 6; ++i) D.td[i] = det.td[i];
        fast_constants(D, det);
    }
    const ulonglong2 *pairs = reinterpret_cast<const ulonglong2 *>(a.pairs);
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; 2 * p < a.n_planes; p += gridDim.x * blockDim.x) {
        const ulonglong2 v0 = pairs[2 * p], v1 = pairs[2 * p + 1];
        Bottom g;
        eval_bottom<true, true>(D, from_u64(v0.x), from_u64(v0.y), from_u64(v1.x), from_u64(v1.y), g);
        const f2 r1 = sub2(PackFast::sqrt(g.na), bc(D.td[1])), r2 = sub2(PackFast::sqrt(g.nb), bc(D.td[2])),
                 r3 = sub2(PackFast::sqrt(g.nc), bc(D.td[3]));
        const f2 S3 = add2(add2(abs2(r1), abs2(r2)), abs2(r3));
        const f2 m1 = fma2(S3, bc(9.5367431640625e-07f), fma2(g.w, bc(D.ms), bc(D.mc)));
        votes[2 * p] = 0; zneg[2 * p] = 0; resid[2 * p] = lo(S3);
        if (margin) margin[2 * p] = lo(m1);
        if (2 * p + 1 < a.n_planes) {
            votes[2 * p + 1] = 0; zneg[2 * p + 1] = 0; resid[2 * p + 1] = hi(S3);
            if (margin) margin[2 * p + 1] = hi(m1);
        }
    }
}

int launch_scores(gpp_handle *h, const float *d_det /*12+3+12 floats*/, const int32_t *d_orient, int which,
                  int32_t *votes, float *resid, int32_t *zneg, float *margin, cudaStream_t s) {
    const int threads = 128, blocks = 64;
    if (which == 0) {
        PollArgs<float> a = {};
        a.det_list = nullptr; a.det_count = nullptr;
        a.boxes = d_det; a.dims = d_det + 12; a.pinv = d_det + 15; a.orient = d_orient;
        a.planes = h->d_planes32; a.n_planes = h->n_planes;
        scores_exact_kernel<<<blocks, threads, 0, s>>>(a, votes, resid, zneg);
    } else {
        PollArgs2<float> b = {};
        b.boxes = d_det; b.dims = d_det + 12; b.pinv = d_det + 15; b.orient = d_orient;
        b.pairs = h->d_pairs; b.planes = h->d_planes32; b.n_planes = h->n_planes;
        b.n_pairs_padded = h->n_pairs_padded;
        if (which == 1) scores_fast_kernel<false><<<blocks, threads, 0, s>>>(b, votes, resid, zneg, margin);
        else if (which == 2) scores_fast_kernel<true><<<blocks, threads, 0, s>>>(b, votes, resid, zneg, margin);
        else scores_bottom_kernel<<<blocks, threads, 0, s>>>(b, votes, resid, zneg, margin);
    }
    h->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(GPP_ECUDA, "scores kernel launch: %s", cudaGetErrorString(e));
    return GPP_OK;
}

// ---------------------------------------------------------------------------------------------------
// Runtime audit of the VERIFIED mode (gpp_audit_set / GPP_AUDIT=n): every n-th detection of a call is polled again
// by the EXACT kernel and compared with what the VERIFIED kernel wrote; the counters live on the device.
// ---------------------------------------------------------------------------------------------------
__global__ void audit_list_kernel(long long n_det, int every, long long n_sample, long long *list, unsigned int *count) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *count = (unsigned int)n_sample;
    if (i < n_sample) list[i] = i * every;
}
__global__ void audit_compare_kernel(const long long *list, long long n_sample, const long long *best_main,
                                     const long long *best_exact, const float *res_main, const float *res_exact,
                                     unsigned long long *counts) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool differs = false;
    if (i < n_sample) {
        const long long m = list[i];
        differs = best_main[m] != best_exact[m] || __float_as_uint(res_main[m]) != __float_as_uint(res_exact[m]);
    }
    const unsigned bad = __popc(__ballot_sync(0xffffffffu, differs));
    const unsigned seen = __popc(__ballot_sync(0xffffffffu, i < n_sample));
    if ((threadIdx.x & 31) == 0) {
        if (seen) atomicAdd(counts, (unsigned long long)seen);
        if (bad) atomicAdd(counts + 1, (unsigned long long)bad);
    }
}

void release_audit(gpp_handle *h) {
    cudaFree(h->audit_out); cudaFree(h->audit_best); cudaFree(h->audit_best_main); cudaFree(h->audit_list);
    cudaFree(h->audit_count); cudaFree(h->audit_counts);
    if (h->audit_done) cudaEventDestroy(h->audit_done);
    h->audit_out = nullptr; h->audit_best = h->audit_best_main = h->audit_list = nullptr;
    h->audit_count = nullptr; h->audit_counts = nullptr; h->audit_done = nullptr; h->audit_cap = 0;
}

static int audit_reserve(gpp_handle *h, long long n_det, cudaStream_t s) {
    cudaError_t e = cudaSuccess;
    if (!h->audit_counts) {
        e = cudaMalloc(&h->audit_counts, 2 * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaMemset(h->audit_counts, 0, 2 * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaMalloc(&h->audit_count, sizeof(unsigned int));
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->audit_done, cudaEventDisableTiming);
    }
    if (e == cudaSuccess && n_det > h->audit_cap) {
        e = cudaDeviceSynchronize();
        cudaFree(h->audit_out); cudaFree(h->audit_best); cudaFree(h->audit_best_main); cudaFree(h->audit_list);
        h->audit_out = nullptr; h->audit_best = h->audit_best_main = h->audit_list = nullptr; h->audit_cap = 0;
        if (e == cudaSuccess) e = cudaMalloc(&h->audit_out, sizeof(float) * 17 * (size_t)n_det);
        if (e == cudaSuccess) e = cudaMalloc(&h->audit_best, sizeof(long long) * (size_t)n_det);
        if (e == cudaSuccess) e = cudaMalloc(&h->audit_best_main, sizeof(long long) * (size_t)n_det);
        if (e == cudaSuccess) e = cudaMalloc(&h->audit_list, sizeof(long long) * (size_t)n_det);
        if (e == cudaSuccess) h->audit_cap = n_det;
    }
    // one set of audit buffers per handle: the previous audited call (possibly on another stream) must be through
    if (e == cudaSuccess && h->audit_used) e = cudaStreamWaitEvent(s, h->audit_done, 0);
    if (e != cudaSuccess) return set_error(GPP_ECUDA, "audit buffers: %s", cudaGetErrorString(e));
    return GPP_OK;
}

static int audit_pass(gpp_handle *h, const PollArgs<float> &a, cudaStream_t s) {
    const int every = h->audit_every;
    const long long n_sample = (a.n_det + every - 1) / every;
    const int threads = 256;
    const unsigned blocks = (unsigned)((n_sample + threads - 1) / threads);
    audit_list_kernel<<<blocks, threads, 0, s>>>(a.n_det, every, n_sample, h->audit_list, h->audit_count);
    PollArgs<float> x = a;
    x.keypoints = h->audit_out;
    x.keyplanes = h->audit_out + 12 * a.n_det;
    x.residuals = h->audit_out + 16 * a.n_det;
    x.best = h->audit_best;
    x.det_list = h->audit_list;
    x.det_count = h->audit_count;
    const long long n_groups = (n_sample + kWarps - 1) / kWarps;
    GPP_K_EXACT1<<<(unsigned)grid_for(h, n_groups, h->occ[0]), kWarps * 32, kSmem1, s>>>(x);
    audit_compare_kernel<<<blocks, threads, 0, s>>>(h->audit_list, n_sample, a.best, h->audit_best, a.residuals,
                                                   x.residuals, h->audit_counts);
    h->launches += 3;
    h->audit_used = true;
    cudaError_t e = cudaEventRecord(h->audit_done, s);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(GPP_ECUDA, "audit pass: %s", cudaGetErrorString(e));
    return GPP_OK;
}

static int launch_poll_f32_main(gpp_handle *h, const PollArgs<float> &a_in, int mode, cudaStream_t s);

int launch_poll_f32(gpp_handle *h, const PollArgs<float> &a_in, int mode, cudaStream_t s) {
    if (mode != GPP_MODE_VERIFIED || h->audit_every <= 0) return launch_poll_f32_main(h, a_in, mode, s);
    int rc = audit_reserve(h, a_in.n_det, s);
    if (rc) return rc;
    PollArgs<float> a = a_in;
    if (!a.best) a.best = h->audit_best_main;
    if ((rc = launch_poll_f32_main(h, a, mode, s))) return rc;
    return audit_pass(h, a, s);
}

static int launch_poll_f32_main(gpp_handle *h, const PollArgs<float> &a_in, int mode, cudaStream_t s) {
    // the packed modes run the resident-database kernel unless a test / tuning hook asks for a ring kernel
    if ((mode == GPP_MODE_FAST || mode == GPP_MODE_VERIFIED) && h->force_variant == 0 && h->force_split == 0)
        return launch_poll3(h, a_in, mode, s);
    gpp_handle::WorkSlot *w = nullptr;
    int rc = begin_unique(h, a_in, s, &w);
    if (rc) return rc;
    PollArgs<float> a = a_in;
    a.det_list = w->ulist; a.det_count = w->count + 1;
    // grids are sized for the worst case (every row unique); the kernels read the real count on the device
    if (mode == GPP_MODE_FAST || mode == GPP_MODE_VERIFIED) {
        PollArgs2<float> b;
        b.boxes = a.boxes; b.dims = a.dims; b.pinv = a.pinv; b.orient = a.orient;
        b.pairs = h->d_pairs; b.planes = h->d_planes32; b.n_planes = h->n_planes;
        b.n_pairs_padded = h->n_pairs_padded; b.dets_per_image = a.dets_per_image; b.n_det = a.n_det;
        b.keypoints = a.keypoints; b.keyplanes = a.keyplanes; b.residuals = a.residuals; b.best = a.best;
        b.det_list = a.det_list; b.det_count = a.det_count;
        b.group_counter = w->count + 2;
        const bool split = use_split(h, a.n_det, mode == GPP_MODE_VERIFIED ? 33 : 27);
        const long long n_groups = split ? a.n_det : (a.n_det + kWarps - 1) / kWarps;
        if (mode == GPP_MODE_VERIFIED) {
            int v = GPP_DEFAULT_VARIANT_VERIFIED;
            if (h->force_variant >= 2 && h->force_variant <= 4) v = h->force_variant - 2;
            if (split) {
                GPP_K_VERIFIED_SPLIT<<<(unsigned)grid_for(h, n_groups, h->occ_split[2]), kWarps * 32, kSmem2, s>>>(b);
            } else {
                verified_variant(v)<<<(unsigned)grid_for(h, n_groups, h->occ3[v]), kWarps * 32, kSmem2, s>>>(b);
#ifdef GPP_STATS
                {
                    unsigned long long st[8], zero[8] = {0};
                    cudaDeviceSynchronize();
                    cudaMemcpyFromSymbol(st, g_stats, sizeof(st));
                    cudaMemcpyToSymbol(g_stats, zero, sizeof(zero));
                    const double rows = double(st[0] + st[1] + st[2]);
                    fprintf(stderr, "[gpp stats] dets %llu rows/det %.1f: all-six %.1f%% (pass %.2f%%), general>=4 %.1f%% <4 %.1f%% (pass %.2f%% of all rows); "
                            "exact verifications/det %.1f, flushes/det %.2f\n", st[7], rows / st[7], 100.0 * st[0] / rows,
                            100.0 * st[3] / rows, 100.0 * st[1] / rows, 100.0 * st[2] / rows, 100.0 * st[4] / rows,
                            double(st[5]) / st[7], double(st[6]) / st[7]);
                }
#endif
            }
        } else if (split) {
            GPP_K_FAST_SPLIT<<<(unsigned)grid_for(h, n_groups, h->occ_split[1]), kWarps * 32, kSmem2, s>>>(b);
        } else {
            int v = GPP_DEFAULT_VARIANT_FAST;
            if (h->force_variant >= 2 && h->force_variant <= 4) v = h->force_variant - 2;
            fast_variant(v)<<<(unsigned)grid_for(h, n_groups, h->occ2[v]), kWarps * 32, kSmem2, s>>>(b);
        }
    } else if (use_split(h, a.n_det, 15)) {
        GPP_K_EXACT_SPLIT<<<(unsigned)grid_for(h, a.n_det, h->occ_split[0]), kWarps * 32, kSmem1, s>>>(a);
    } else {
        const long long resident = (long long)h->sm_count * h->occ[0] * kWarps;
#ifdef GPP_EXACT_DPW2   /* experiment: two detections per warp for large batches (slower since r01n) */
        const bool two = a.n_det >= 4 * resident;
#else
        const bool two = false;
        (void)resident;
#endif
        if (two) {
            const long long n_groups = (a.n_det + 2 * kWarps - 1) / (2 * kWarps);
            GPP_K_EXACT2<<<(unsigned)grid_for(h, n_groups, h->occ[1]), kWarps * 32, kSmem1, s>>>(a);
        } else {
            const long long n_groups = (a.n_det + kWarps - 1) / kWarps;
            GPP_K_EXACT1<<<(unsigned)grid_for(h, n_groups, h->occ[0]), kWarps * 32, kSmem1, s>>>(a);
        }
    }
    h->launches += 1;
    return end_unique(h, a, s, w);
}

int launch_poll_f64(gpp_handle *h, const PollArgs<double> &a_in, cudaStream_t s) {
    gpp_handle::WorkSlot *w = nullptr;
    int rc = begin_unique(h, a_in, s, &w);
    if (rc) return rc;
    PollArgs<double> a = a_in;
    a.det_list = w->ulist; a.det_count = w->count + 1;
    if (use_split(h, a.n_det, 12)) {
        GPP_K_F64_SPLIT<<<(unsigned)grid_for(h, a.n_det, h->occ_split[3]), kWarps * 32, kSmem64, s>>>(a);
    } else {
        const long long n_groups = (a.n_det + kWarps - 1) / kWarps;
        GPP_K_F64<<<(unsigned)grid_for(h, n_groups, h->occ[2]), kWarps * 32, kSmem64, s>>>(a);
    }
    h->launches += 1;
    return end_unique(h, a, s, w);
}

}  // namespace gpp
