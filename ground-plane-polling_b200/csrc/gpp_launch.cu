// Launch configuration of the polling kernel (gpp_poll3.cuh), the once-per-database pair interleave, the runtime audit
// and the per-hypothesis score kernels of the test hooks.
#if defined(GPP_STATS) || defined(GPP_TIMELINE)
#include <cstdio>
#endif
#include "../../include/gpp_debug.h"
#include "gpp_internal.h"

namespace gpp {

// pair-interleaved, padded copy of the normalised fp32 database in SCAN ORDER (gpp_order.cu): position q holds plane
// scan_index[q]; pair p = positions (2p, 2p+1) stored as {a0,a1,b0,b1,c0,c1,d0,d1}.  Positions past N-1 are copies of
// position N-1 (the scans never queue / select a position >= N).
__global__ void scan_index_kernel(const int32_t *__restrict__ order, int n, int n_padded, int32_t *__restrict__ scan_index) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_padded) return;
    const int src = min(q, n - 1);
    scan_index[q] = order ? order[src] : src;
}
__global__ void interleave_pairs_kernel(const float4 *__restrict__ planes, const int32_t *__restrict__ scan_index,
                                        int n, int n_pairs_padded, float *__restrict__ pairs,
                                        float4 *__restrict__ planes_scan) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs_padded) return;
    const float4 a = planes[scan_index[2 * p]], b = planes[scan_index[2 * p + 1]];
    float4 *out = reinterpret_cast<float4 *>(pairs + 8 * (size_t)p);
    out[0] = make_float4(a.x, b.x, a.y, b.y);
    out[1] = make_float4(a.z, b.z, a.w, b.w);
    if (2 * p < n) planes_scan[2 * p] = a;               // the plain database in scan order (EXACT scan)
    if (2 * p + 1 < n) planes_scan[2 * p + 1] = b;
}

int build_pairs(gpp_handle *h, const int32_t *order, cudaStream_t s) {
    const int np = h->n_pairs_padded, n = h->n_planes;
    cudaError_t e = cudaSuccess;
    if (order) {
        // staged in the (not yet written) pair buffer: 4 bytes per plane of its 16.  `order` is pageable host memory, so
        // the call returns once it has been read.
        int32_t *tmp = reinterpret_cast<int32_t *>(h->d_pairs) + 2 * (size_t)np;      // second quarter of the buffer
        e = cudaMemcpyAsync(tmp, order, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) return set_error(GPP_ECUDA, "scan order upload: %s", cudaGetErrorString(e));
        scan_index_kernel<<<(2 * np + 127) / 128, 128, 0, s>>>(tmp, n, 2 * np, h->d_scan_index);
    } else {
        scan_index_kernel<<<(2 * np + 127) / 128, 128, 0, s>>>(nullptr, n, 2 * np, h->d_scan_index);
    }
    interleave_pairs_kernel<<<(np + 127) / 128, 128, 0, s>>>(h->d_planes32, h->d_scan_index, n, np,
                                                            reinterpret_cast<float *>(h->d_pairs), h->d_planes32_scan);
    h->launches += 2;
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(GPP_ECUDA, "interleave_pairs_kernel launch: %s", cudaGetErrorString(e));
    return GPP_OK;
}

// One persistent CTA per SM.  32 warps at 64 registers for the fp32 modes (no spills once the rarely used detection
// constants of the VERIFIED filter are parked in shared memory): measured 4.84e11 hypotheses/s (VERIFIED, C4) against
// 4.77e11 with 28 warps / 72 registers and 4.49e11 with 24 / 80; the fp64 scan needs 128 registers: 16 warps.
#ifndef GPP_WARPS3
#define GPP_WARPS3 32
#endif
constexpr int kWarps3 = GPP_WARPS3;
constexpr int kWarps64 = 16;
typedef void (*Poll3Fn)(const PollArgs3);
template <bool kSeg, bool kPose>
static Poll3Fn poll3_variant_of(int mode) {
    switch (mode) {
        case GPP_MODE_FAST: return poll3_kernel<kWarps3, kModeFast, kSeg, kPose>;
        case GPP_MODE_VERIFIED: return poll3_kernel<kWarps3, kModeVerified, kSeg, kPose>;
        case GPP_MODE_EXACT: return poll3_kernel<kWarps3, kModeExact, kSeg, kPose>;
        default: return poll3_kernel<kWarps64, kModeF64, kSeg, kPose>;
    }
}
static Poll3Fn poll3_variant(int mode, bool seg, bool pose) {
    if (seg) return pose ? poll3_variant_of<true, true>(mode) : poll3_variant_of<true, false>(mode);
    return pose ? poll3_variant_of<false, true>(mode) : poll3_variant_of<false, false>(mode);
}
static const int kAllModes[4] = {GPP_MODE_FAST, GPP_MODE_VERIFIED, GPP_MODE_EXACT, GPP_MODE_F64};
static int warps_of(int mode) { return mode == GPP_MODE_F64 ? kWarps64 : kWarps3; }

int configure_kernels(gpp_handle *h) {
    int optin = 0;
    cudaError_t e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);
    if (e != cudaSuccess) return set_error(GPP_ECUDA, "shared memory query: %s", cudaGetErrorString(e));
    cudaFuncAttributes fa;
    size_t fixed = 0;
    for (int v = 0; v < 16; ++v) {
        e = cudaFuncGetAttributes(&fa, poll3_variant(kAllModes[v >> 2], v & 1, v & 2));
        if (e != cudaSuccess) return set_error(GPP_ECUDA, "cudaFuncGetAttributes: %s", cudaGetErrorString(e));
        if (fa.sharedSizeBytes > fixed) fixed = fa.sharedSizeBytes;
    }
    const long long room = (long long)optin - (long long)fixed - (long long)smem3_bytes(kWarps3, 0);
    if (room < 0) return set_error(GPP_ECUDA, "polling kernel does not fit on an SM");
    h->resident_cap_rows = (int)(room / 1024);
    const int max_dyn = (int)smem3_bytes(kWarps3, h->resident_cap_rows);
    for (int v = 0; v < 16 && e == cudaSuccess; ++v)
        e = cudaFuncSetAttribute(poll3_variant(kAllModes[v >> 2], v & 1, v & 2), cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn);
    if (e != cudaSuccess) return set_error(GPP_ECUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    // scratch of segmented detections: segments are only used below 3 detections per resident warp
    const long long slots = (long long)h->sm_count * kWarps3;
    h->seg_det_cap = 4 * slots;
    h->seg_items_cap = 12 * slots + 64;
    for (int i = 0; i < gpp_handle::kSlots3 && e == cudaSuccess; ++i) {
        gpp_handle::Slot3 &w = h->slot3[i];
        e = cudaMalloc(&w.claim, 2 * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaMemset(w.claim, 0, 2 * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaMalloc(&w.partials, sizeof(SegPartial) * (size_t)h->seg_items_cap);
        if (e == cudaSuccess) e = cudaMalloc(&w.seg_arrived, sizeof(unsigned int) * (size_t)h->seg_det_cap);
        if (e == cudaSuccess) e = cudaMemset(w.seg_arrived, 0, sizeof(unsigned int) * (size_t)h->seg_det_cap);
        if (e == cudaSuccess) e = cudaMalloc(&w.seg_best, sizeof(unsigned long long) * (size_t)h->seg_det_cap);
        if (e == cudaSuccess) e = cudaMemset(w.seg_best, 0, sizeof(unsigned long long) * (size_t)h->seg_det_cap);
        if (e == cudaSuccess) e = cudaMalloc(&w.seg_consts, sizeof(float) * kSharedConsts * (size_t)h->seg_det_cap);
        if (e == cudaSuccess) e = cudaMalloc(&w.seg_ready, sizeof(unsigned int) * (size_t)h->seg_det_cap);
        if (e == cudaSuccess) e = cudaMemset(w.seg_ready, 0, sizeof(unsigned int) * (size_t)h->seg_det_cap);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&w.done, cudaEventDisableTiming);
    }
    if (e != cudaSuccess) return set_error(GPP_ECUDA, "polling scratch allocation: %s", cudaGetErrorString(e));
    return GPP_OK;
}

void release_poll3(gpp_handle *h) {
    for (auto &w : h->slot3) {
        cudaFree(w.claim); cudaFree(w.partials); cudaFree(w.seg_arrived); cudaFree(w.seg_best);
        cudaFree(w.seg_consts); cudaFree(w.seg_ready);
        if (w.done) cudaEventDestroy(w.done);
        w = gpp_handle::Slot3();
    }
}

// Schedule of one call (see the header of gpp_poll3.cuh).  Segments: below four detections per resident warp every
// detection is cut into plane segments so that the work items fill the machine about eight times over (the last wave
// is then short whatever the batch size, and no single item -- a detection without a six-vote plane costs three times
// the average -- is long enough to be the tail; at least 3, at most 24 unless forced).  Measured with the segment-major
// item order (kernel ms, verified): 64 x 100 x 10k: 3 segments 0.220, 4: 0.218, 6: 0.213, 8: 0.216; 128 x 100 x 10k: 1: 0.401,
// 3: 0.372, 4: 0.380; 256 x 100 x 22k: 1: 1.35, 3: 1.35, 4: 1.31; 384 x 100 x 22k: 1: 1.88, 3: 1.95; 512 x 100 x 22k: 1: 2.44,
// 2: 2.60, 3: 2.58, 4: 2.55 (with the shared constants; left whole from four detections per warp on);
// a single image: 24: 0.038, 32: 0.045.  Tried and dropped (r02): whole rows first and segments only for the last partial
// wave (64 x 100 x 10k: 0.32 ms against 0.26 ms).  Residency: staging up to 212 KB per SM pays as soon as every warp polls
// a few items; a call with fewer items than that streams every row from L2 and starts at once.
static int launch_poll3(gpp_handle *h, const FitIO &io, int mode, int det_stride, cudaStream_t s) {
    gpp_handle::Slot3 &w = h->slot3[h->next_slot3++ % gpp_handle::kSlots3];
    cudaError_t e = cudaSuccess;
    if (w.used) e = cudaStreamWaitEvent(s, w.done, 0);          // previous user of this slot (another stream)
    if (e != cudaSuccess) return set_error(GPP_ECUDA, "polling slot: %s", cudaGetErrorString(e));
    PollArgs3 b;
    b.boxes = io.boxes; b.dims = io.dims; b.pinv = io.pinv; b.orient = io.orient;
    b.pairs = h->d_pairs; b.scan_index = h->d_scan_index; b.planes = h->d_planes32; b.planes_scan = h->d_planes32_scan; b.planes64 = h->d_planes64; b.n_planes = h->n_planes;
    b.n_pairs_padded = h->n_pairs_padded; b.dets_per_image = io.D; b.n_det = io.n_det;
    b.keypoints = io.keypoints; b.keyplanes = io.keyplanes; b.residuals = io.residuals; b.best = io.best;
    b.pose_locations = io.pose_locations; b.pose_angles = io.pose_angles; b.pose_dimensions = io.pose_dimensions;
    b.pose_kitti = io.pose_locations ? io.pose_kitti : nullptr;
    b.det_stride = det_stride;
    b.claim = w.claim; b.partials = w.partials; b.seg_arrived = w.seg_arrived; b.seg_best = w.seg_best;
    b.seg_consts = w.seg_consts; b.seg_ready = w.seg_ready;
    const int warps = warps_of(mode);
    const int NR = h->n_pairs_padded / 32;
    const long long slots = (long long)h->sm_count * warps;
    const long long n_rows = (io.n_det + det_stride - 1) / det_stride;
    int n_seg = 1;
    if (h->force_seg > 0) n_seg = h->force_seg;
    else if (n_rows < 4 * slots) {
        n_seg = (int)((8 * slots + n_rows / 2) / n_rows);
        if (n_seg < 3) n_seg = 3;            // two segments were slower than one or three at every size measured
    }
    if (h->force_seg <= 0 && n_seg > 24) n_seg = 24;      // a single image: 0.038 ms with 24 segments, 0.045 ms with 32
    if (n_seg > 32) n_seg = 32;
    if (n_seg > NR) n_seg = NR;
    if (n_rows > h->seg_det_cap || n_rows * n_seg > h->seg_items_cap) n_seg = 1;
    b.n_seg = n_seg < 1 ? 1 : n_seg;
    const long long n_items = n_rows * b.n_seg;
    int res = 0;
    if (mode == GPP_MODE_FAST || mode == GPP_MODE_VERIFIED) {
        res = n_items >= 2 * slots ? h->resident_cap_rows : 0;
        if (h->force_resident >= 0) res = h->force_resident;
        if (res > h->resident_cap_rows) res = h->resident_cap_rows;
        if (res > NR) res = NR;
    }
    b.resident_rows = res;
    long long grid = h->sm_count;
    if (grid > n_items) grid = n_items;
    const size_t smem = smem3_bytes(warps, res);
    poll3_variant(mode, b.n_seg > 1, b.pose_locations != nullptr)<<<(unsigned)grid, warps * 32, smem, s>>>(b);
#ifdef GPP_TIMELINE
    {
        static unsigned long long host_tl[8192][6];
        unsigned int n = 0, zero = 0;
        cudaEvent_t e0, e1;
        cudaDeviceSynchronize();
        cudaMemcpyFromSymbol(&n, g_timeline_n, sizeof(n));
        cudaMemcpyFromSymbol(host_tl, g_timeline, sizeof(host_tl));
        cudaMemcpyToSymbol(g_timeline_n, &zero, sizeof(zero));
        if (n > 8192) n = 8192;
        unsigned long long t_min = ~0ull, t_max = 0;
        for (unsigned i = 0; i < n; ++i) { if (host_tl[i][1] < t_min) t_min = host_tl[i][1]; if (host_tl[i][3] > t_max) t_max = host_tl[i][3]; }
        fprintf(stderr, "[gpp timeline] %u records, n_seg %d, span %.2f us\n", n, b.n_seg, (t_max - t_min) * 1e-3);
        for (unsigned i = 0; i < n; ++i) {
            if (host_tl[i][5] >> 63)
                fprintf(stderr, "[tl] item %llu kind 1 same_rays 0 start %.2f a %.2f b %.2f c %.2f d %.2f\n", host_tl[i][0],
                        (host_tl[i][1] - t_min) * 1e-3, (host_tl[i][2] - t_min) * 1e-3, (host_tl[i][4] - t_min) * 1e-3,
                        ((host_tl[i][5] & ~(1ull << 63)) - t_min) * 1e-3, (host_tl[i][3] - t_min) * 1e-3);
            else
                fprintf(stderr, "[tl] item %llu kind %llu same_rays %llu start %.2f a %.2f b %.2f\n", host_tl[i][0], host_tl[i][5], host_tl[i][4],
                        (host_tl[i][1] - t_min) * 1e-3, (host_tl[i][2] - t_min) * 1e-3, (host_tl[i][3] - t_min) * 1e-3);
        }
        (void)e0; (void)e1;
    }
#endif
#ifdef GPP_STATS
    if (mode == GPP_MODE_VERIFIED) {
        unsigned long long st[10], zero[10] = {0};
        cudaDeviceSynchronize();
        cudaMemcpyFromSymbol(st, g_stats3, sizeof(st));
        cudaMemcpyToSymbol(g_stats3, zero, sizeof(zero));
        const double rows = double(st[0] + st[2]), items = double(st[5] ? st[5] : 1);
        fprintf(stderr, "[gpp stats3] items %llu (identical-rays %llu) seg %d resident %d: rows/item %.1f, all-six %.1f%% (past stage 1: %.1f%% "
                "of them), general %.1f%% (at max-votes >= 4: %.1f%% past the bottom face), rows with survivors %.2f%%; exact verifications/item %.1f, flushes/item %.2f\n",
                st[5], st[6], b.n_seg, res, rows / items, 100.0 * st[0] / (rows ? rows : 1), 100.0 * st[1] / (st[0] ? st[0] : 1),
                100.0 * st[2] / (rows ? rows : 1), 100.0 * st[8] / (st[9] ? st[9] : 1), 100.0 * st[7] / (rows ? rows : 1), double(st[3]) / items, double(st[4]) / items);
    }
#endif
    h->launches += 1;
    w.used = true;
    e = cudaEventRecord(w.done, s);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(GPP_ECUDA, "polling kernel launch: %s", cudaGetErrorString(e));
    return GPP_OK;
}

// ---------------------------------------------------------------------------------------------------
// Runtime audit of the VERIFIED mode (gpp_audit_set / GPP_AUDIT=n): every n-th detection of a call is polled again
// in the EXACT mode (the same kernel with a row stride) and compared with what the VERIFIED pass wrote; the counters
// live on the device.
// ---------------------------------------------------------------------------------------------------
__global__ void audit_compare_kernel(long long n_det, int every, const long long *best_main, const long long *best_exact,
                                     const float *res_main, const float *res_exact, unsigned long long *counts) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long m = i * every;
    const bool in = m < n_det;
    const bool differs = in && (best_main[m] != best_exact[m] || __float_as_uint(res_main[m]) != __float_as_uint(res_exact[m]));
    const unsigned bad = __popc(__ballot_sync(0xffffffffu, differs));
    const unsigned seen = __popc(__ballot_sync(0xffffffffu, in));
    if ((threadIdx.x & 31) == 0) {
        if (seen) atomicAdd(counts, (unsigned long long)seen);
        if (bad) atomicAdd(counts + 1, (unsigned long long)bad);
    }
}

void release_audit(gpp_handle *h) {
    cudaFree(h->audit_out); cudaFree(h->audit_best); cudaFree(h->audit_best_main); cudaFree(h->audit_counts);
    if (h->audit_done) cudaEventDestroy(h->audit_done);
    h->audit_out = nullptr; h->audit_best = h->audit_best_main = nullptr;
    h->audit_counts = nullptr; h->audit_done = nullptr; h->audit_cap = 0;
}

static int audit_reserve(gpp_handle *h, long long n_det, cudaStream_t s) {
    cudaError_t e = cudaSuccess;
    if (!h->audit_counts) {
        e = cudaMalloc(&h->audit_counts, 2 * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaMemset(h->audit_counts, 0, 2 * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->audit_done, cudaEventDisableTiming);
    }
    if (e == cudaSuccess && n_det > h->audit_cap) {
        e = cudaDeviceSynchronize();
        cudaFree(h->audit_out); cudaFree(h->audit_best); cudaFree(h->audit_best_main);
        h->audit_out = nullptr; h->audit_best = h->audit_best_main = nullptr; h->audit_cap = 0;
        if (e == cudaSuccess) e = cudaMalloc(&h->audit_out, sizeof(float) * 17 * (size_t)n_det);
        if (e == cudaSuccess) e = cudaMalloc(&h->audit_best, sizeof(long long) * (size_t)n_det);
        if (e == cudaSuccess) e = cudaMalloc(&h->audit_best_main, sizeof(long long) * (size_t)n_det);
        if (e == cudaSuccess) h->audit_cap = n_det;
    }
    // one set of audit buffers per handle: the previous audited call (possibly on another stream) must be through
    if (e == cudaSuccess && h->audit_used) e = cudaStreamWaitEvent(s, h->audit_done, 0);
    if (e != cudaSuccess) return set_error(GPP_ECUDA, "audit buffers: %s", cudaGetErrorString(e));
    return GPP_OK;
}

int launch_poll(gpp_handle *h, const FitIO &io_in, int mode, cudaStream_t s) {
    if (mode != GPP_MODE_VERIFIED || h->audit_every <= 0) return launch_poll3(h, io_in, mode, 1, s);
    int rc = audit_reserve(h, io_in.n_det, s);
    if (rc) return rc;
    FitIO io = io_in;
    if (!io.best) io.best = h->audit_best_main;
    if ((rc = launch_poll3(h, io, mode, 1, s))) return rc;
    FitIO x = io;
    x.keypoints = h->audit_out;
    x.keyplanes = h->audit_out + 12 * io.n_det;
    x.residuals = h->audit_out + 16 * io.n_det;
    x.best = h->audit_best;
    x.pose_locations = x.pose_angles = x.pose_dimensions = x.pose_kitti = nullptr;
    if ((rc = launch_poll3(h, x, GPP_MODE_EXACT, h->audit_every, s))) return rc;
    const long long n_sample = (io.n_det + h->audit_every - 1) / h->audit_every;
    const int threads = 256;
    audit_compare_kernel<<<(unsigned)((n_sample + threads - 1) / threads), threads, 0, s>>>(
        io.n_det, h->audit_every, io.best, h->audit_best, static_cast<const float *>(io.residuals),
        static_cast<const float *>(x.residuals), h->audit_counts);
    h->launches += 1;
    h->audit_used = true;
    cudaError_t e = cudaEventRecord(h->audit_done, s);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(GPP_ECUDA, "audit pass: %s", cudaGetErrorString(e));
    return GPP_OK;
}

// ---------------------------------------------------------------------------------------------------
// test hook: per-hypothesis scores of one detection, computed by the very device functions the search
// loops call (hypothesis<ExactF32> / eval_pair<PackFast>), so tests can compare them with the oracle
// hypothesis by hypothesis
// ---------------------------------------------------------------------------------------------------
struct ScoreArgs {
    const float *boxes, *dims, *pinv;
    const int32_t *orient;
    const u64 *pairs;
    const int32_t *scan_index;       // plane index of every position of `pairs`
    const float4 *planes;
    int n_planes;
};

__global__ void scores_exact_kernel(ScoreArgs a, int32_t *votes, float *resid, int32_t *zneg) {
    Detection<ExactF32> det;
    load_detection<ExactF32, ExactF32>(det, a.boxes, a.dims, a.orient[0], a.pinv);
    const float4 *pl = a.planes;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < a.n_planes; j += gridDim.x * blockDim.x) {
        float X[4][3];
        int V; float R; bool z;
        hypothesis<ExactF32>(det, pl[j].x, pl[j].y, pl[j].z, pl[j].w, X, V, R, z);
        votes[j] = V; resid[j] = R; zneg[j] = z ? 1 : 0;
    }
}

template <bool kSix>
__global__ void scores_fast_kernel(ScoreArgs a, int32_t *votes, float *resid, int32_t *zneg, float *margin) {
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
        const int j0 = a.scan_index[2 * p];
        votes[j0] = V0 + 16 * loose_votes(h, false); resid[j0] = lo(R);
        zneg[j0] = (lo(h.zc) < 0.0f ? 1 : 0) + (!(lo(zhi) < 0.0f) ? 2 : 0);
        if (margin) margin[j0] = lo(h.m);
        if (2 * p + 1 < a.n_planes) {
            const int j1 = a.scan_index[2 * p + 1];
            votes[j1] = V1 + 16 * loose_votes(h, true); resid[j1] = hi(R);
            zneg[j1] = (hi(h.zc) < 0.0f ? 1 : 0) + (!(hi(zhi) < 0.0f) ? 2 : 0);
            if (margin) margin[j1] = hi(h.m);
        }
    }
}

// which == 3: stage 1 of the VERIFIED all-six phase -- resid = sum of the three bottom-face residuals (merged
// reciprocals, as in the kernel), margin = the bound the stage-1 test relies on: w ms + mc + 2^-20 S3
__global__ void scores_bottom_kernel(ScoreArgs a, int32_t *votes, float *resid, int32_t *zneg, float *margin) {
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
        const int j0 = a.scan_index[2 * p];
        votes[j0] = 0; zneg[j0] = 0; resid[j0] = lo(S3);
        if (margin) margin[j0] = lo(m1);
        if (2 * p + 1 < a.n_planes) {
            const int j1 = a.scan_index[2 * p + 1];
            votes[j1] = 0; zneg[j1] = 0; resid[j1] = hi(S3);
            if (margin) margin[j1] = hi(m1);
        }
    }
}

int launch_scores(gpp_handle *h, const float *d_det /*12+3+12 floats*/, const int32_t *d_orient, int which,
                  int32_t *votes, float *resid, int32_t *zneg, float *margin, cudaStream_t s) {
    const int threads = 128, blocks = 64;
    ScoreArgs a;
    a.boxes = d_det; a.dims = d_det + 12; a.pinv = d_det + 15; a.orient = d_orient;
    a.pairs = h->d_pairs; a.scan_index = h->d_scan_index; a.planes = h->d_planes32; a.n_planes = h->n_planes;
    if (which == 0) scores_exact_kernel<<<blocks, threads, 0, s>>>(a, votes, resid, zneg);
    else if (which == 1) scores_fast_kernel<false><<<blocks, threads, 0, s>>>(a, votes, resid, zneg, margin);
    else if (which == 2) scores_fast_kernel<true><<<blocks, threads, 0, s>>>(a, votes, resid, zneg, margin);
    else scores_bottom_kernel<<<blocks, threads, 0, s>>>(a, votes, resid, zneg, margin);
    h->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(GPP_ECUDA, "scores kernel launch: %s", cudaGetErrorString(e));
    return GPP_OK;
}

}  // namespace gpp
