// The two steps that precede polling in the reference's inference graph:
//
//   decode_kernel   RegressBoxes + RegressDims   keras_retinanet_3D/layers/_misc.py:132-140, :185-186,
//                                                backend/common.py:23-84 (bbox_transform_inv, dim_transform_inv)
//   score_kernel +  FilterDetections             keras_retinanet_3D/layers/filter_detections.py:18-189 for the
//   nms_kernel                                   configuration the model uses (models/retinanet.py:415): one class,
//                                                class_specific_filter, no orientation-specific filter, NMS on
//
// Both are HBM-bound byte/float shuffling (no dense contraction): one thread per anchor with 128-bit loads and
// stores for decode and scoring; one CTA per image for sort + greedy NMS + gather.  float32 arithmetic in the
// order the reference writes it (explicit _rn intrinsics: no FMA contraction) so the results are bit-identical
// to the oracle (oracle/detect_ref.py).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gpp.h"
#include "gpp_internal.h"

namespace gpp {

struct DecodeParams {
    float box_mean[12], box_std[12], dim_mean[3], dim_std[3];
};

__global__ void __launch_bounds__(256) decode_kernel(const float4 *__restrict__ anchors, const float4 *__restrict__ regression,
                                                     const float4 *__restrict__ classification,
                                                     const float *__restrict__ regression_dim, long long n, int A,
                                                     DecodeParams p, float4 *__restrict__ boxes, float *__restrict__ dims) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 an = anchors[i % A];
    const float4 c0 = classification[2 * i], c1 = classification[2 * i + 1];
    // argmax over the 8 scores (first maximum); the x offsets of the middle / top key-points point left (-1) when
    // it falls in the first half (_misc.py:133-136)
    const float cs[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
    int am = 0;
#pragma unroll
    for (int k = 1; k < 8; ++k)
        if (cs[k] > cs[am]) am = k;
    const float sign = am < 4 ? -1.0f : 1.0f;
    const float4 r0 = regression[3 * i], r1 = regression[3 * i + 1], r2 = regression[3 * i + 2];
    const float d[12] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w};
    const float width = __fsub_rn(an.z, an.x), height = __fsub_rn(an.w, an.y);
    float t[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) t[k] = __fadd_rn(__fmul_rn(d[k], p.box_std[k]), p.box_mean[k]);
    const float mid = __fdiv_rn(__fadd_rn(an.x, an.z), 2.0f);
    float4 o0, o1, o2;
    o0.x = __fadd_rn(an.x, __fmul_rn(t[0], width));                       // x1
    o0.y = __fadd_rn(an.y, __fmul_rn(t[1], height));                      // y1
    o0.z = __fadd_rn(an.z, __fmul_rn(t[2], width));                       // x2
    o0.w = __fadd_rn(an.w, __fmul_rn(t[3], height));                      // y2
    o1.x = __fadd_rn(an.x, __fmul_rn(t[4], width));                       // xl
    o1.y = __fadd_rn(an.w, __fmul_rn(t[5], height));                      // yl
    o1.z = __fadd_rn(mid, __fmul_rn(__fmul_rn(t[6], width), sign));       // xm
    o1.w = __fadd_rn(an.w, __fmul_rn(t[7], height));                      // ym
    o2.x = __fadd_rn(an.z, __fmul_rn(t[8], width));                       // xr
    o2.y = __fadd_rn(an.w, __fmul_rn(t[9], height));                      // yr
    o2.z = __fadd_rn(mid, __fmul_rn(__fmul_rn(t[10], width), sign));      // xt
    o2.w = __fadd_rn(an.y, __fmul_rn(t[11], height));                     // yt
    boxes[3 * i] = o0; boxes[3 * i + 1] = o1; boxes[3 * i + 2] = o2;
#pragma unroll
    for (int k = 0; k < 3; ++k)
        dims[3 * i + k] = __fadd_rn(__fmul_rn(regression_dim[3 * i + k], p.dim_std[k]), p.dim_mean[k]);
}

// ---------------------------------------------------------------------------------------------------
// FilterDetections.  Candidate key: (~score bits) << 32 | anchor index -- ascending order of the key is
// descending score with ties broken by the lower anchor index (scores above the threshold are positive
// floats, whose bit patterns order like the values).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) score_kernel(const float4 *__restrict__ classification, int A, int n_img,
                                                    float score_threshold, unsigned long long *__restrict__ keys,
                                                    unsigned char *__restrict__ orient, unsigned int *__restrict__ counts,
                                                    long long key_stride) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n_img * A) return;
    const int b = (int)(i / A), a = (int)(i % A);
    const float4 c0 = classification[2 * i], c1 = classification[2 * i + 1];
    // max over the two halves, then arg-max / max over the four orientations (filter_detections.py:66-67, :117-119)
    const float s4[4] = {fmaxf(c0.x, c1.x), fmaxf(c0.y, c1.y), fmaxf(c0.z, c1.z), fmaxf(c0.w, c1.w)};
    int o = 0;
#pragma unroll
    for (int k = 1; k < 4; ++k)
        if (s4[k] > s4[o]) o = k;
    const float s = s4[o];
    orient[i] = (unsigned char)o;
    if (s > score_threshold) {
        const unsigned int slot = atomicAdd(&counts[b], 1u);
        keys[(long long)b * key_stride + slot] = ((unsigned long long)(~__float_as_uint(s)) << 32) | (unsigned int)a;
    }
}

__device__ __forceinline__ bool iou_gt(const float4 bi, const float4 bj, float thr) {
    // TF's IOUGreaterThanThreshold, in the (x1, y1, x2, y2) layout the reference passes (symmetric formula)
    const float y1i = fminf(bi.x, bi.z), x1i = fminf(bi.y, bi.w), y2i = fmaxf(bi.x, bi.z), x2i = fmaxf(bi.y, bi.w);
    const float y1j = fminf(bj.x, bj.z), x1j = fminf(bj.y, bj.w), y2j = fmaxf(bj.x, bj.z), x2j = fmaxf(bj.y, bj.w);
    const float ai = __fmul_rn(__fsub_rn(y2i, y1i), __fsub_rn(x2i, x1i));
    const float aj = __fmul_rn(__fsub_rn(y2j, y1j), __fsub_rn(x2j, x1j));
    if (ai <= 0.0f || aj <= 0.0f) return false;
    const float ih = fmaxf(__fsub_rn(fminf(y2i, y2j), fmaxf(y1i, y1j)), 0.0f);
    const float iw = fmaxf(__fsub_rn(fminf(x2i, x2j), fmaxf(x1i, x1j)), 0.0f);
    const float inter = __fmul_rn(ih, iw);
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(ai, aj), inter)) > thr;
}

constexpr int kNmsThreads = 512;
constexpr int kNmsSmemKeys = 4096;      // candidate lists up to this size are sorted in shared memory
constexpr int kMaxDet = 128;

// One CTA per image: bitonic sort of the candidate keys, greedy NMS in sorted order (a candidate is kept iff its IoU
// with every box kept so far is <= the threshold; stops at max_detections), gather + pad with -1.
__global__ void __launch_bounds__(kNmsThreads) nms_kernel(const float *__restrict__ boxes, const float *__restrict__ dims,
                                                          const unsigned char *__restrict__ orient, int A,
                                                          unsigned long long *__restrict__ keys_all, long long key_stride,
                                                          const unsigned int *__restrict__ counts, float nms_threshold,
                                                          int max_det, float *__restrict__ out_boxes,
                                                          float *__restrict__ out_dims, float *__restrict__ out_scores,
                                                          int32_t *__restrict__ out_labels, int32_t *__restrict__ out_orient) {
    __shared__ unsigned long long skeys[kNmsSmemKeys];
    __shared__ float4 sel_box[kMaxDet];
    __shared__ float4 cand_box[kNmsThreads];
    __shared__ unsigned int batch_mask[kNmsThreads / 32];
    __shared__ unsigned char batch_alive[kNmsThreads / 32];
    __shared__ unsigned long long sel_key[kMaxDet];
    __shared__ int n_sel;
    const int b = blockIdx.x;
    const int K = (int)min(counts[b], (unsigned int)A);
    int P = 1;
    while (P < K) P <<= 1;
    unsigned long long *gkeys = keys_all + (long long)b * key_stride;
    unsigned long long *keys = (P <= kNmsSmemKeys) ? skeys : gkeys;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const unsigned long long v = i < K ? gkeys[i] : ~0ull;             // padding sorts last
        if (keys == skeys) skeys[i] = v;
        else if (i >= K) gkeys[i] = v;
    }
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < P; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const unsigned long long x = keys[i], y = keys[l];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { keys[i] = y; keys[l] = x; }
                }
            }
            __syncthreads();
        }
    if (threadIdx.x == 0) n_sel = 0;
    __syncthreads();
    const float4 *boxes4 = reinterpret_cast<const float4 *>(boxes) + (long long)b * A * 3;   // first float4 = x1,y1,x2,y2
    // Greedy pass in sorted order: a candidate is kept iff no box kept before it overlaps it by more than the threshold.
    // The chain is serial in the candidates, so it is walked 16 at a time: the candidates' boxes are gathered into
    // shared memory by the whole CTA (kNmsThreads per gather); warp w tests candidate c0 + w against every box kept
    // before the batch and against its batch-mates before it (a bit mask); one thread then resolves the batch in order
    // with bit operations.  (One candidate at a time took 0.13 ms for one image with 300 candidates, r02n.)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kBatch = kNmsThreads / 32;
    for (int g0 = 0; g0 < K && n_sel < max_det; g0 += kNmsThreads) {
        const int gn = min(kNmsThreads, K - g0);
        if ((int)threadIdx.x < gn) cand_box[threadIdx.x] = boxes4[3 * (long long)(keys[g0 + threadIdx.x] & 0xffffffffu)];
        __syncthreads();
        for (int c0 = 0; c0 < gn; c0 += kBatch) {
            const int ns = n_sel;                          // CTA-uniform: written before the last barrier
            if (ns >= max_det) break;
            const int c = c0 + warp;
            const bool valid = c < gn;
            const float4 cand = cand_box[valid ? c : 0];
            bool hit = false;
            for (int sidx = lane; sidx < ns; sidx += 32) hit = hit || iou_gt(cand, sel_box[sidx], nms_threshold);
            const bool alive = valid && !__any_sync(0xffffffffu, hit);
            const bool mate = lane < warp && c0 + lane < gn;
            const unsigned mask = __ballot_sync(0xffffffffu, mate && iou_gt(cand, cand_box[mate ? c0 + lane : 0], nms_threshold));
            if (lane == 0) { batch_mask[warp] = mask; batch_alive[warp] = alive ? 1 : 0; }
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned kept = 0u;
                int n = ns;
                for (int w = 0; w < kBatch && n < max_det; ++w)
                    if (batch_alive[w] && !(batch_mask[w] & kept)) {
                        kept |= 1u << w;
                        sel_box[n] = cand_box[c0 + w];
                        sel_key[n] = keys[g0 + c0 + w];
                        ++n;
                    }
                n_sel = n;
            }
            __syncthreads();
        }
    }
    const int ns = n_sel;
    // gather (filter_detections.py:163-168) and pad with -1 (:170-177); 17 values per output row
    for (int t = threadIdx.x; t < max_det * 17; t += blockDim.x) {
        const int r = t / 17, f = t % 17;
        const long long orow = (long long)b * max_det + r;
        if (r < ns) {
            const unsigned long long key = sel_key[r];
            const long long a = (long long)b * A + (long long)(key & 0xffffffffu);
            if (f < 12) out_boxes[orow * 12 + f] = boxes[a * 12 + f];
            else if (f < 15) out_dims[orow * 3 + (f - 12)] = dims[a * 3 + (f - 12)];
            else if (f == 15) out_scores[orow] = __uint_as_float(~(unsigned int)(key >> 32));
            else { out_labels[orow] = 0; out_orient[orow] = (int32_t)orient[a]; }
        } else {
            if (f < 12) out_boxes[orow * 12 + f] = -1.0f;
            else if (f < 15) out_dims[orow * 3 + (f - 12)] = -1.0f;
            else if (f == 15) out_scores[orow] = -1.0f;
            else { out_labels[orow] = -1; out_orient[orow] = -1; }
        }
    }
}

}  // namespace gpp

using gpp::set_error;

#define GPP_CUDA(expr)                                                                               \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return set_error(GPP_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                             __FILE__, __LINE__);                                                    \
    } while (0)

namespace {
struct DevGuard {
    int prev = -1;
    explicit DevGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        cudaSetDevice(dev);
    }
    ~DevGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

const float kBoxMean[12] = {-0.0373f, -0.0165f, 0.0373f, 0.0171f, -0.0286f, -0.0478f, 0.2929f, 0.0114f, 0.0288f, -0.0589f,
                            0.2932f, -0.0007f};                                           // layers/_misc.py:115
const float kBoxStd[12] = {0.1957f, 0.1896f, 0.1957f, 0.1897f, 0.1967f, 0.2034f, 0.2046f, 0.1898f, 0.1964f, 0.2052f,
                           0.2048f, 0.1903f};                                             // layers/_misc.py:117
const float kDimMean[3] = {1.6570f, 1.7999f, 4.2907f};                                    // layers/_misc.py:168
const float kDimStd[3] = {0.2681f, 0.2243f, 0.6281f};                                     // layers/_misc.py:170
}  // namespace

extern "C" {

int gpp_decode_device(gpp_handle *h, const float *anchors, const float *regression, const float *classification,
                      const float *regression_dim, int B, int A, const float *box_mean_std, const float *dim_mean_std,
                      float *boxes, float *dimensions, void *stream) {
    if (!h || B < 0 || A < 0) return set_error(GPP_EINVAL, "gpp_decode_device: bad argument");
    if ((long long)B * A == 0) return GPP_OK;
    if (!anchors || !regression || !classification || !regression_dim || !boxes || !dimensions)
        return set_error(GPP_EINVAL, "gpp_decode_device: NULL array argument");
    DevGuard guard(h->device);
    gpp::DecodeParams p;
    for (int k = 0; k < 12; ++k) {
        p.box_mean[k] = box_mean_std ? box_mean_std[k] : kBoxMean[k];
        p.box_std[k] = box_mean_std ? box_mean_std[12 + k] : kBoxStd[k];
    }
    for (int k = 0; k < 3; ++k) {
        p.dim_mean[k] = dim_mean_std ? dim_mean_std[k] : kDimMean[k];
        p.dim_std[k] = dim_mean_std ? dim_mean_std[3 + k] : kDimStd[k];
    }
    const long long n = (long long)B * A;
    gpp::decode_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4 *>(anchors), reinterpret_cast<const float4 *>(regression),
        reinterpret_cast<const float4 *>(classification), regression_dim, n, A, p, reinterpret_cast<float4 *>(boxes),
        dimensions);
    h->launches += 1;
    GPP_CUDA(cudaGetLastError());
    return GPP_OK;
}

int gpp_filter_device(gpp_handle *h, const float *boxes, const float *dimensions, const float *classification, int B,
                      int A, float score_threshold, float nms_threshold, int max_detections, float *out_boxes,
                      float *out_dimensions, float *out_scores, int32_t *out_labels, int32_t *out_orientations,
                      void *stream) {
    if (!h || B < 0 || A < 0 || max_detections < 1 || max_detections > gpp::kMaxDet)
        return set_error(GPP_EINVAL, "gpp_filter_device: bad argument (max_detections must be 1..%d)", gpp::kMaxDet);
    if (B == 0) return GPP_OK;
    if (!out_boxes || !out_dimensions || !out_scores || !out_labels || !out_orientations || (A > 0 && (!boxes || !dimensions || !classification)))
        return set_error(GPP_EINVAL, "gpp_filter_device: NULL array argument");
    DevGuard guard(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    long long stride = 1;
    while (stride < A) stride <<= 1;                                        // room for the sort's power-of-two padding
    if (stride < 32) stride = 32;
    // scratch: keys for a chunk of images, orientation per anchor, counters
    const long long max_chunk_keys = 1ll << 24;                             // 128 MB of keys at most
    int chunk = (int)(max_chunk_keys / stride);
    if (chunk < 1) chunk = 1;
    if (chunk > B) chunk = B;
    const size_t need_keys = sizeof(unsigned long long) * (size_t)chunk * stride;
    const size_t need_orient = (size_t)chunk * (size_t)(A > 0 ? A : 1);
    if (need_keys > h->filter_keys_bytes || need_orient > h->filter_orient_bytes) {
        GPP_CUDA(cudaDeviceSynchronize());
        cudaFree(h->filter_keys); cudaFree(h->filter_orient); cudaFree(h->filter_counts);
        h->filter_keys = nullptr; h->filter_orient = nullptr; h->filter_counts = nullptr;
        h->filter_keys_bytes = h->filter_orient_bytes = 0;
        GPP_CUDA(cudaMalloc(&h->filter_keys, need_keys));
        GPP_CUDA(cudaMalloc(&h->filter_orient, need_orient));
        GPP_CUDA(cudaMalloc(&h->filter_counts, sizeof(unsigned int) * (size_t)chunk));
        h->filter_keys_bytes = need_keys;
        h->filter_orient_bytes = need_orient;
        h->filter_counts_n = chunk;
    } else if (chunk > h->filter_counts_n) {
        GPP_CUDA(cudaDeviceSynchronize());
        cudaFree(h->filter_counts);
        h->filter_counts = nullptr;
        GPP_CUDA(cudaMalloc(&h->filter_counts, sizeof(unsigned int) * (size_t)chunk));
        h->filter_counts_n = chunk;
    }
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int nb = (b0 + chunk <= B) ? chunk : (B - b0);
        GPP_CUDA(cudaMemsetAsync(h->filter_counts, 0, sizeof(unsigned int) * (size_t)nb, s));
        const long long n = (long long)nb * A;
        if (n > 0) {
            gpp::score_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(
                reinterpret_cast<const float4 *>(classification) + 2ll * b0 * A, A, nb, score_threshold, h->filter_keys,
                h->filter_orient, h->filter_counts, stride);
            h->launches += 1;
        }
        gpp::nms_kernel<<<nb, gpp::kNmsThreads, 0, s>>>(
            boxes + 12ll * b0 * A, dimensions + 3ll * b0 * A, h->filter_orient, A, h->filter_keys, stride, h->filter_counts,
            nms_threshold, max_detections, out_boxes + 12ll * b0 * max_detections, out_dimensions + 3ll * b0 * max_detections,
            out_scores + (long long)b0 * max_detections, out_labels + (long long)b0 * max_detections,
            out_orientations + (long long)b0 * max_detections);
        h->launches += 1;
        GPP_CUDA(cudaGetLastError());
    }
    return GPP_OK;
}


// host entries: plain synchronous wrappers (device buffers for the call, copies on the handle's first stream)
int gpp_decode_host(gpp_handle *h, const float *anchors, const float *regression, const float *classification,
                    const float *regression_dim, int B, int A, const float *box_mean_std, const float *dim_mean_std,
                    float *boxes, float *dimensions) {
    if (!h || B < 0 || A < 0) return set_error(GPP_EINVAL, "gpp_decode_host: bad argument");
    const size_t n = (size_t)B * A;
    if (n == 0) return GPP_OK;
    if (!anchors || !regression || !classification || !regression_dim || !boxes || !dimensions)
        return set_error(GPP_EINVAL, "gpp_decode_host: NULL array argument");
    DevGuard guard(h->device);
    cudaStream_t s = h->streams[0];
    float *d = nullptr;
    GPP_CUDA(cudaMalloc(&d, sizeof(float) * ((size_t)A * 4 + n * (12 + 8 + 3 + 12 + 3))));
    float *d_an = d, *d_reg = d_an + (size_t)A * 4, *d_cls = d_reg + n * 12, *d_rd = d_cls + n * 8, *d_box = d_rd + n * 3,
          *d_dim = d_box + n * 12;
    int rc = GPP_OK;
    cudaError_t e = cudaMemcpyAsync(d_an, anchors, sizeof(float) * (size_t)A * 4, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_reg, regression, sizeof(float) * n * 12, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_cls, classification, sizeof(float) * n * 8, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_rd, regression_dim, sizeof(float) * n * 3, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) rc = gpp_decode_device(h, d_an, d_reg, d_cls, d_rd, B, A, box_mean_std, dim_mean_std, d_box, d_dim, s);
    if (e == cudaSuccess && rc == GPP_OK) e = cudaMemcpyAsync(boxes, d_box, sizeof(float) * n * 12, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && rc == GPP_OK) e = cudaMemcpyAsync(dimensions, d_dim, sizeof(float) * n * 3, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d);
    if (rc != GPP_OK) return rc;
    if (e != cudaSuccess) return set_error(GPP_ECUDA, "gpp_decode_host: %s", cudaGetErrorString(e));
    return GPP_OK;
}

int gpp_filter_host(gpp_handle *h, const float *boxes, const float *dimensions, const float *classification, int B,
                    int A, float score_threshold, float nms_threshold, int max_detections, float *out_boxes,
                    float *out_dimensions, float *out_scores, int32_t *out_labels, int32_t *out_orientations) {
    if (!h || B < 0 || A < 0 || max_detections < 1 || max_detections > gpp::kMaxDet)
        return set_error(GPP_EINVAL, "gpp_filter_host: bad argument (max_detections must be 1..%d)", gpp::kMaxDet);
    if (B == 0) return GPP_OK;
    if (!out_boxes || !out_dimensions || !out_scores || !out_labels || !out_orientations || (A > 0 && (!boxes || !dimensions || !classification)))
        return set_error(GPP_EINVAL, "gpp_filter_host: NULL array argument");
    DevGuard guard(h->device);
    cudaStream_t s = h->streams[0];
    const size_t n = (size_t)B * A, m = (size_t)B * max_detections;
    float *d = nullptr;
    GPP_CUDA(cudaMalloc(&d, sizeof(float) * (n * (12 + 3 + 8) + m * (12 + 3 + 1 + 2) + 1)));
    float *d_box = d, *d_dim = d_box + n * 12, *d_cls = d_dim + n * 3, *o_box = d_cls + n * 8, *o_dim = o_box + m * 12,
          *o_sc = o_dim + m * 3;
    int32_t *o_lab = reinterpret_cast<int32_t *>(o_sc + m), *o_or = o_lab + m;
    int rc = GPP_OK;
    cudaError_t e = cudaSuccess;
    if (n) {
        e = cudaMemcpyAsync(d_box, boxes, sizeof(float) * n * 12, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_dim, dimensions, sizeof(float) * n * 3, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_cls, classification, sizeof(float) * n * 8, cudaMemcpyHostToDevice, s);
    }
    if (e == cudaSuccess)
        rc = gpp_filter_device(h, d_box, d_dim, d_cls, B, A, score_threshold, nms_threshold, max_detections, o_box, o_dim,
                               o_sc, o_lab, o_or, s);
    if (e == cudaSuccess && rc == GPP_OK) {
        e = cudaMemcpyAsync(out_boxes, o_box, sizeof(float) * m * 12, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(out_dimensions, o_dim, sizeof(float) * m * 3, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(out_scores, o_sc, sizeof(float) * m, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(out_labels, o_lab, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(out_orientations, o_or, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, s);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d);
    if (rc != GPP_OK) return rc;
    if (e != cudaSuccess) return set_error(GPP_ECUDA, "gpp_filter_host: %s", cudaGetErrorString(e));
    return GPP_OK;
}

}  // extern "C"
