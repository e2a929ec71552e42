// libgpp C ABI (include/gpp.h): handle, plane-database upload + on-device normalisation, the host and
// device entry points of fit_road_planes, measurement helpers.  No CPU fallback anywhere in this file.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gpp_debug.h"
#include "gpp_internal.h"

namespace {
thread_local std::string g_last_error;
}

namespace gpp {
int set_error(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
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

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// ---------------------------------------------------------------------------------------------------
// plane normalisation, fit_road_planes.py:75-77 (once per database instead of once per call)
// ---------------------------------------------------------------------------------------------------
template <class E>
__global__ void normalise_planes_kernel(const float *__restrict__ raw, int n, typename E::T4 *__restrict__ out) {
    typedef typename E::T T;
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const float4 r = reinterpret_cast<const float4 *>(raw)[j];
    T p[4] = {T(r.x), T(r.y), T(r.z), T(r.w)};
    const T dir = -gpp::tf_sign(p[1]);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = E::mul(p[i], dir);
    const T rho = E::sqrt(E::add(E::add(E::mul(p[0], p[0]), E::mul(p[1], p[1])), E::mul(p[2], p[2])));
    typename E::T4 o;
    o.x = E::div(p[0], rho);
    o.y = E::div(p[1], rho);
    o.z = E::div(p[2], rho);
    o.w = E::div(p[3], rho);
    out[j] = o;
}

// ---------------------------------------------------------------------------------------------------
extern "C" {

int gpp_version(void) { return GPP_VERSION; }
const char *gpp_last_error(void) { return g_last_error.c_str(); }
int gpp_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) {
        cudaGetLastError();          // clear the sticky "no device" error
        return 0;
    }
    return count;
}

int gpp_create(int device, gpp_handle **out) {
    if (!out) return set_error(GPP_EINVAL, "gpp_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return set_error(GPP_ENODEV, "gpp_create: no CUDA device (%s); libgpp has no CPU fallback",
                         e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count)
        return set_error(GPP_EINVAL, "gpp_create: device %d out of range [0, %d)", device, count);
    cudaDeviceProp prop;
    GPP_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return set_error(GPP_ENODEV, "gpp_create: device %d is sm_%d%d; libgpp is built for sm_100a only", device,
                         prop.major, prop.minor);
    DeviceGuard guard(device);
    if (!guard.ok) return set_error(GPP_ECUDA, "gpp_create: cudaSetDevice(%d) failed", device);
    gpp_handle *h = new (std::nothrow) gpp_handle();
    if (!h) return set_error(GPP_ENOMEM, "gpp_create: out of host memory");
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    for (int i = 0; i < gpp_handle::kStreams; ++i) {
        e = cudaStreamCreateWithFlags(&h->streams[i], cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            gpp_destroy(h);
            return set_error(GPP_ECUDA, "gpp_create: cudaStreamCreate failed: %s", cudaGetErrorString(e));
        }
    }
    e = cudaEventCreate(&h->ev_start);
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev_stop);
    if (e != cudaSuccess) {
        gpp_destroy(h);
        return set_error(GPP_ECUDA, "gpp_create: cudaEventCreate failed: %s", cudaGetErrorString(e));
    }
    int rc = gpp::configure_kernels(h);
    if (rc != GPP_OK) {
        gpp_destroy(h);
        return rc;
    }
    if (const char *env = getenv("GPP_AUDIT")) h->audit_every = atoi(env) > 0 ? atoi(env) : 0;
    *out = h;
    return GPP_OK;
}

int gpp_destroy(gpp_handle *h) {
    if (!h) return GPP_OK;
    DeviceGuard guard(h->device);
    cudaDeviceSynchronize();
    cudaFree(h->d_raw);
    cudaFree(h->d_planes32);
    cudaFree(h->d_planes32_scan);
    cudaFree(h->d_planes64);
    cudaFree(h->d_pairs);
    cudaFree(h->d_scan_index);
    cudaFree(h->filter_keys);
    cudaFree(h->filter_orient);
    cudaFree(h->filter_counts);
    gpp::release_poll3(h);
    gpp::release_audit(h);
    for (int i = 0; i < gpp_handle::kStreams; ++i) {
        h->stage[i].release();
        h->hstage[i].release();
        if (h->streams[i]) cudaStreamDestroy(h->streams[i]);
    }
    for (auto &p : h->chunk_events) {
        cudaEventDestroy(p.first);
        cudaEventDestroy(p.second);
    }
    if (h->planes_ready) cudaEventDestroy(h->planes_ready);
    if (h->ev_start) cudaEventDestroy(h->ev_start);
    if (h->ev_stop) cudaEventDestroy(h->ev_stop);
    delete h;
    return GPP_OK;
}

static int ensure_plane_capacity(gpp_handle *h, int n) {
    if (n <= h->cap_planes) return GPP_OK;
    cudaFree(h->d_raw);
    cudaFree(h->d_planes32);
    cudaFree(h->d_planes32_scan);
    cudaFree(h->d_planes64);
    cudaFree(h->d_pairs);
    cudaFree(h->d_scan_index);
    h->d_planes32_scan = nullptr;
    h->d_pairs = nullptr;
    h->d_scan_index = nullptr;
    h->d_raw = nullptr;
    h->d_planes32 = nullptr;
    h->d_planes64 = nullptr;
    h->cap_planes = 0;
    h->n_planes = 0;
    GPP_CUDA(cudaMalloc(&h->d_raw, sizeof(float) * 4 * (size_t)n));
    GPP_CUDA(cudaMalloc(&h->d_planes32, sizeof(float4) * (size_t)n));
    GPP_CUDA(cudaMalloc(&h->d_planes32_scan, sizeof(float4) * (size_t)n));
    GPP_CUDA(cudaMalloc(&h->d_planes64, sizeof(double4) * (size_t)n));
    GPP_CUDA(cudaMalloc(&h->d_pairs, 32 * (size_t)((n + 63) / 64) * 32));
    GPP_CUDA(cudaMalloc(&h->d_scan_index, sizeof(int32_t) * 64 * (size_t)((n + 63) / 64)));
    h->cap_planes = n;
    return GPP_OK;
}

// `host_rows`: the database as fed (N x 4 float32, row-major) when the host has it -- the scan order of the pair
// database is derived from it (gpp_order.cu); null = scan in index order
static int normalise_on(gpp_handle *h, int n, const float *host_rows, cudaStream_t s) {
    const int threads = 128, blocks = (n + threads - 1) / threads;
    normalise_planes_kernel<gpp::ExactF32><<<blocks, threads, 0, s>>>(h->d_raw, n, h->d_planes32);
    normalise_planes_kernel<gpp::ExactF64><<<blocks, threads, 0, s>>>(h->d_raw, n, h->d_planes64);
    h->launches += 2;
    GPP_CUDA(cudaGetLastError());
    h->n_planes = n;
    h->n_pairs_padded = ((n + 63) / 64) * 32;
    if (!host_rows) return gpp::build_pairs(h, nullptr, s);
    std::vector<int32_t> order;
    gpp::scan_order(host_rows, n, order);
    return gpp::build_pairs(h, order.data(), s);
}

// every launch that may still read the resident database has an event on record: wait for them on `s`
static int wait_for_fits(gpp_handle *h, cudaStream_t s) {
    for (auto &w : h->slot3)
        if (w.used) GPP_CUDA(cudaStreamWaitEvent(s, w.done, 0));
    if (h->audit_used) GPP_CUDA(cudaStreamWaitEvent(s, h->audit_done, 0));
    return GPP_OK;
}

int gpp_set_planes_raw(gpp_handle *h, const void *planes, int n_planes, int dtype, int order) {
    if (!h || !planes || n_planes <= 0 || dtype < 0 || dtype > 1 || order < 0 || order > 1)
        return set_error(GPP_EINVAL, "gpp_set_planes_raw: bad argument");
    const size_t esz = dtype ? sizeof(double) : sizeof(float);
    const size_t bytes = esz * 4 * (size_t)n_planes;
    // The reference's callers re-feed the same database with every image: compare with the bytes of the last upload
    // (exact, at memcmp speed) instead of converting / hashing / uploading again.
    if (h->raw_valid && h->n_planes == n_planes && h->raw_dtype == dtype && h->raw_order == order &&
        h->raw_copy.size() == bytes && memcmp(h->raw_copy.data(), planes, bytes) == 0)
        return GPP_OK;
    DeviceGuard guard(h->device);
    // the Keras feed casts to float32 (run_network.py:105); row-major N x 4
    std::vector<float> rows(4 * (size_t)n_planes);
    for (size_t j = 0; j < (size_t)n_planes; ++j)
        for (int c = 0; c < 4; ++c) {
            const size_t src = order ? (size_t)c * n_planes + j : 4 * j + c;
            rows[4 * j + c] = dtype ? (float)static_cast<const double *>(planes)[src] : static_cast<const float *>(planes)[src];
        }
    // all earlier work of this handle that may still read the old database must be done
    GPP_CUDA(cudaDeviceSynchronize());
    int rc = ensure_plane_capacity(h, n_planes);
    if (rc) return rc;
    cudaStream_t s = h->streams[0];
    GPP_CUDA(cudaMemcpyAsync(h->d_raw, rows.data(), sizeof(float) * rows.size(), cudaMemcpyHostToDevice, s));
    rc = normalise_on(h, n_planes, rows.data(), s);
    if (rc) return rc;
    GPP_CUDA(cudaStreamSynchronize(s));
    h->n_planes = n_planes;
    h->raw_copy.assign(static_cast<const unsigned char *>(planes), static_cast<const unsigned char *>(planes) + bytes);
    h->raw_dtype = dtype;
    h->raw_order = order;
    h->raw_valid = true;
    return GPP_OK;
}

int gpp_set_planes(gpp_handle *h, const float *planes, int n_planes) {
    return gpp_set_planes_raw(h, planes, n_planes, 0, 0);
}

int gpp_set_planes_device(gpp_handle *h, const float *d_planes, int n_planes, void *stream) {
    if (!h || !d_planes || n_planes <= 0) return set_error(GPP_EINVAL, "gpp_set_planes_device: bad argument");
    DeviceGuard guard(h->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (n_planes > h->cap_planes) {
        GPP_CUDA(cudaDeviceSynchronize());
    } else {
        int rc = wait_for_fits(h, s);          // fits in flight on other streams still read the buffers overwritten below
        if (rc) return rc;
    }
    int rc = ensure_plane_capacity(h, n_planes);
    if (rc) return rc;
    GPP_CUDA(cudaMemcpyAsync(h->d_raw, d_planes, sizeof(float) * 4 * (size_t)n_planes, cudaMemcpyDeviceToDevice, s));
    // The scan order is computed on the host (a k-d tree over the planes, ~1 ms): a database of 32 rows or more is read
    // back once per update, which blocks the caller until `s` has caught up.  Inside a stream capture (no host round
    // trip possible) and for small databases the pair database keeps the index order.
    std::vector<float> host_rows;
    cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &capturing) != cudaSuccess) cudaGetLastError();
    if (n_planes >= 32 * 64 && capturing == cudaStreamCaptureStatusNone) {
        host_rows.resize(4 * (size_t)n_planes);
        GPP_CUDA(cudaMemcpyAsync(host_rows.data(), h->d_raw, sizeof(float) * host_rows.size(), cudaMemcpyDeviceToHost, s));
        GPP_CUDA(cudaStreamSynchronize(s));
    }
    rc = normalise_on(h, n_planes, host_rows.empty() ? nullptr : host_rows.data(), s);
    if (rc) return rc;
    h->n_planes = n_planes;
    h->raw_valid = false;
    // later fits on other streams must see the new database
    if (!h->planes_ready) GPP_CUDA(cudaEventCreateWithFlags(&h->planes_ready, cudaEventDisableTiming));
    GPP_CUDA(cudaEventRecord(h->planes_ready, s));
    h->planes_stream = s;
    h->planes_pending = true;
    return GPP_OK;
}

int gpp_num_planes(const gpp_handle *h) { return h ? h->n_planes : 0; }

int gpp_get_normalised_planes(gpp_handle *h, float *out) {
    if (!h || !out || h->n_planes <= 0) return set_error(GPP_EINVAL, "gpp_get_normalised_planes: no planes set");
    DeviceGuard guard(h->device);
    GPP_CUDA(cudaDeviceSynchronize());
    GPP_CUDA(cudaMemcpy(out, h->d_planes32, sizeof(float4) * (size_t)h->n_planes, cudaMemcpyDeviceToHost));
    return GPP_OK;
}

// ---------------------------------------------------------------------------------------------------
// fit: device entries
// ---------------------------------------------------------------------------------------------------
static int check_fit_args(const gpp_handle *h, const void *boxes, const void *dims, const void *orient,
                          const void *pinv, int B, int D, const void *kp, const void *kpl, const void *res,
                          const char *who) {
    if (!h) return set_error(GPP_EINVAL, "%s: handle is NULL", who);
    if (B < 0 || D < 0) return set_error(GPP_EINVAL, "%s: negative size B=%d D=%d", who, B, D);
    if (h->n_planes <= 0) return set_error(GPP_EINVAL, "%s: no plane database set (call gpp_set_planes)", who);
    if ((long long)B * D > 0 && (!boxes || !dims || !orient || !pinv || !kp || !kpl || !res))
        return set_error(GPP_EINVAL, "%s: NULL array argument", who);
    return GPP_OK;
}

static bool float_mode(int mode) { return mode == GPP_MODE_EXACT || mode == GPP_MODE_FAST || mode == GPP_MODE_VERIFIED; }

static int fit_device_impl(gpp_handle *h, const gpp::FitIO &io, int mode, cudaStream_t s) {
    if (io.n_det == 0) return GPP_OK;
    DeviceGuard guard(h->device);
    if (h->planes_pending && h->planes_stream != s) GPP_CUDA(cudaStreamWaitEvent(s, h->planes_ready, 0));
    GPP_CUDA(cudaEventRecord(h->ev_start, s));
    int rc = gpp::launch_poll(h, io, mode, s);
    if (rc) return rc;
    GPP_CUDA(cudaEventRecord(h->ev_stop, s));
    h->timing_chunks = 0;
    h->timing_single = true;
    return GPP_OK;
}

int gpp_fit_device(gpp_handle *h, const float *boxes, const float *dimensions, const int32_t *orientations,
                   const float *P_inv, int B, int D, float *keypoints, float *keyplanes, float *residuals,
                   int64_t *best_index, int mode, void *stream) {
    int rc = check_fit_args(h, boxes, dimensions, orientations, P_inv, B, D, keypoints, keyplanes, residuals,
                            "gpp_fit_device");
    if (rc) return rc;
    if (!float_mode(mode))
        return set_error(GPP_EINVAL, "gpp_fit_device: mode %d (use gpp_fit_device_f64 for the FP64 mode)", mode);
    gpp::FitIO io;
    io.boxes = boxes; io.dims = dimensions; io.orient = orientations; io.pinv = P_inv;
    io.D = D; io.n_det = (long long)B * D;
    io.keypoints = keypoints; io.keyplanes = keyplanes; io.residuals = residuals;
    io.best = reinterpret_cast<long long *>(best_index);
    return fit_device_impl(h, io, mode, static_cast<cudaStream_t>(stream));
}

int gpp_fit_pose_device(gpp_handle *h, const float *boxes, const float *dimensions, const int32_t *orientations,
                        const float *P_inv, int B, int D, float *keypoints, float *keyplanes, float *residuals,
                        int64_t *best_index, float *locations, float *angles, float *dimensions_out, float *kitti,
                        int mode, void *stream) {
    int rc = check_fit_args(h, boxes, dimensions, orientations, P_inv, B, D, keypoints, keyplanes, residuals,
                            "gpp_fit_pose_device");
    if (rc) return rc;
    if (!float_mode(mode)) return set_error(GPP_EINVAL, "gpp_fit_pose_device: mode %d", mode);
    if ((long long)B * D > 0 && (!locations || !angles || !dimensions_out))
        return set_error(GPP_EINVAL, "gpp_fit_pose_device: NULL pose array");
    gpp::FitIO io;
    io.boxes = boxes; io.dims = dimensions; io.orient = orientations; io.pinv = P_inv;
    io.D = D; io.n_det = (long long)B * D;
    io.keypoints = keypoints; io.keyplanes = keyplanes; io.residuals = residuals;
    io.best = reinterpret_cast<long long *>(best_index);
    io.pose_locations = locations; io.pose_angles = angles; io.pose_dimensions = dimensions_out; io.pose_kitti = kitti;
    return fit_device_impl(h, io, mode, static_cast<cudaStream_t>(stream));
}

int gpp_fit_device_f64(gpp_handle *h, const float *boxes, const float *dimensions,
                       const int32_t *orientations, const float *P_inv, int B, int D, double *keypoints,
                       double *keyplanes, double *residuals, int64_t *best_index, void *stream) {
    int rc = check_fit_args(h, boxes, dimensions, orientations, P_inv, B, D, keypoints, keyplanes, residuals,
                            "gpp_fit_device_f64");
    if (rc) return rc;
    gpp::FitIO io;
    io.boxes = boxes; io.dims = dimensions; io.orient = orientations; io.pinv = P_inv;
    io.D = D; io.n_det = (long long)B * D;
    io.keypoints = keypoints; io.keyplanes = keyplanes; io.residuals = residuals;
    io.best = reinterpret_cast<long long *>(best_index);
    return fit_device_impl(h, io, GPP_MODE_F64, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// fit: host entries.  The image axis is cut into chunks that alternate between two streams, so the H2D
// copy of chunk k+1 and the D2H copy of chunk k-1 overlap the kernel of chunk k (fully asynchronous when
// the caller's buffers are pinned).
// ---------------------------------------------------------------------------------------------------
int gpp::Staging::reserve(size_t bytes) {
    if (bytes <= cap) return GPP_OK;
    release();
    GPP_CUDA(cudaMalloc(reinterpret_cast<void **>(&base), bytes));
    // a staged chunk leaves the device in ONE copy that spans the alignment gaps between its arrays: give those bytes
    // a defined value once (compute-sanitizer initcheck otherwise flags every such copy)
    GPP_CUDA(cudaMemset(base, 0, bytes));
    cap = bytes;
    return GPP_OK;
}

void gpp::Staging::release() {
    cudaFree(base);
    base = nullptr;
    cap = 0;
}

int gpp::HostStaging::reserve(size_t bytes) {
    if (!done) GPP_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
    if (bytes <= cap) return GPP_OK;
    if (base) cudaFreeHost(base);
    base = nullptr; cap = 0;
    GPP_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&base), bytes, cudaHostAllocDefault));
    cap = bytes;
    return GPP_OK;
}

void gpp::HostStaging::release() {
    if (base) cudaFreeHost(base);
    if (done) cudaEventDestroy(done);
    base = nullptr; cap = 0; done = nullptr;
}

// true iff `p` is ordinary (pageable) host memory, i.e. neither cudaHostAlloc'ed / registered nor managed
static bool is_pageable(const void *p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return attr.type == cudaMemoryTypeUnregistered;
}

// One chunk of a host call occupies one block: inputs first, outputs behind them, every array 256-byte aligned.  The
// device staging block and the pinned host staging block use the same layout, so that a staged chunk moves with ONE
// copy in each direction.
struct ChunkLayout {
    size_t o_boxes, o_dims, o_orient, o_pinv, in_end, o_kp, o_kpl, o_res, o_best, o_loc, o_ang, o_pdim, o_kitti, o_end;
    ChunkLayout(size_t n_det, size_t n_img, size_t esz, bool with_best, bool with_pose, bool with_kitti) {
        auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
        o_boxes = 0;
        o_dims = o_boxes + up(48 * n_det);
        o_orient = o_dims + up(12 * n_det);
        o_pinv = o_orient + up(4 * n_det);
        in_end = o_pinv + up(48 * n_img);
        o_kp = in_end;
        o_kpl = o_kp + up(esz * 12 * n_det);
        o_res = o_kpl + up(esz * 4 * n_det);
        o_best = o_res + up(esz * n_det);
        o_loc = o_best + (with_best ? up(8 * n_det) : 0);
        o_ang = o_loc + (with_pose ? up(12 * n_det) : 0);
        o_pdim = o_ang + (with_pose ? up(12 * n_det) : 0);
        o_kitti = o_pdim + (with_pose ? up(12 * n_det) : 0);
        o_end = o_kitti + (with_kitti ? up(16 * n_det) : 0);
    }
};

// optional host outputs of the fused steps after polling
struct PoseHost {
    float *locations = nullptr, *angles = nullptr, *dimensions = nullptr, *kitti = nullptr;
};

template <class T>
static int fit_host_impl(gpp_handle *h, const float *boxes, const float *dims, const int32_t *orient,
                         const float *pinv, int B, int D, T *keypoints, T *keyplanes, T *residuals,
                         int64_t *best, int mode, const PoseHost &pose = PoseHost(),
                         const std::function<void()> *while_the_gpu_works = nullptr) {
    if ((long long)B * D == 0) return GPP_OK;
    DeviceGuard guard(h->device);
    // chunk size: enough hypotheses to fill the machine a few times over, at most 65,536 detections
    const long long max_chunk_det = 65536;
    int imgs_per_chunk = (int)(max_chunk_det / (D > 0 ? D : 1));
    // (a 512-image shard of an 8-GPU call is one chunk: splitting it in two to overlap its 0.14 ms of copies with the
    // kernel was measured -- 2.88 ms against 2.85 ms, the second launch tail costs what the overlap saves)
    if (imgs_per_chunk < 1) imgs_per_chunk = 1;
    if (imgs_per_chunk > B) imgs_per_chunk = B;
    // chunk boundaries (in images); uniform chunks (quarter- and half-size chunks at both ends were measured again in
    // round 2: 19.78 ms against 19.62 ms for C4 -- the smaller launches cost what the shorter exposed copies save)
    std::vector<int> start(1, 0);
    while (start.back() < B) start.push_back(start.back() + imgs_per_chunk < B ? start.back() + imgs_per_chunk : B);
    const int n_chunks = (int)start.size() - 1;
    const int n_streams = n_chunks > 1 ? gpp_handle::kStreams : 1;
    const bool with_pose = pose.locations != nullptr, with_kitti = with_pose && pose.kitti != nullptr;
    const ChunkLayout L((size_t)imgs_per_chunk * D, (size_t)imgs_per_chunk, sizeof(T), best != nullptr, with_pose, with_kitti);
    for (int i = 0; i < n_streams; ++i) {
        int rc = h->stage[i].reserve(L.o_end);
        if (rc) return rc;
    }
    while ((int)h->chunk_events.size() < n_chunks) {
        cudaEvent_t a, b;
        GPP_CUDA(cudaEventCreate(&a));
        GPP_CUDA(cudaEventCreate(&b));
        h->chunk_events.push_back(std::make_pair(a, b));
    }
    // Staged chunks go through the pinned blocks owned by the handle: pageable caller memory (an asynchronous copy from
    // or to pageable memory blocks the host until it is done, which would serialise the chunks), and every small call
    // (a single image is 6.4 KB in and 7.6 KB out: one copy each way instead of four).
    const bool small = n_chunks == 1 && (long long)B * D <= 16384;
    const bool staged = small || is_pageable(boxes) || is_pageable(dims) || is_pageable(orient) || is_pageable(pinv) ||
                        is_pageable(keypoints) || is_pageable(keyplanes) || is_pageable(residuals) ||
                        (best && is_pageable(best)) || (with_pose && (is_pageable(pose.locations) ||
                        is_pageable(pose.angles) || is_pageable(pose.dimensions))) || (with_kitti && is_pageable(pose.kitti));
    if (staged)
        for (int i = 0; i < n_streams; ++i) {
            int rc = h->hstage[i].reserve(L.o_end);
            if (rc) return rc;
        }
    // copies the finished outputs of chunk c from its pinned block to the caller's arrays
    auto drain = [&](int c) -> int {
        const int b0 = start[c], nb = start[c + 1] - start[c];
        const size_t m0 = (size_t)b0 * D, nm = (size_t)nb * D;
        gpp::HostStaging &hs = h->hstage[c % n_streams];
        GPP_CUDA(cudaEventSynchronize(hs.done));
        memcpy(keypoints + 12 * m0, hs.base + L.o_kp, sizeof(T) * 12 * nm);
        memcpy(keyplanes + 4 * m0, hs.base + L.o_kpl, sizeof(T) * 4 * nm);
        memcpy(residuals + m0, hs.base + L.o_res, sizeof(T) * nm);
        if (best) memcpy(best + m0, hs.base + L.o_best, sizeof(long long) * nm);
        if (with_pose) {
            memcpy(pose.locations + 3 * m0, hs.base + L.o_loc, sizeof(float) * 3 * nm);
            memcpy(pose.angles + 3 * m0, hs.base + L.o_ang, sizeof(float) * 3 * nm);
            memcpy(pose.dimensions + 3 * m0, hs.base + L.o_pdim, sizeof(float) * 3 * nm);
        }
        if (with_kitti) memcpy(pose.kitti + 4 * m0, hs.base + L.o_kitti, sizeof(float) * 4 * nm);
        return GPP_OK;
    };
    for (int c = 0; c < n_chunks; ++c) {
        const int b0 = start[c], nb = start[c + 1] - start[c];
        const long long m0 = (long long)b0 * D, nm = (long long)nb * D;
        unsigned char *dev = h->stage[c % n_streams].base;
        gpp::HostStaging &hs = h->hstage[c % n_streams];
        cudaStream_t s = h->streams[c % n_streams];
        const float *src_boxes = boxes + 12 * m0, *src_dims = dims + 3 * m0, *src_pinv = pinv + 12 * (size_t)b0;
        const int32_t *src_orient = orient + m0;
        if (staged) {
            if (c >= n_streams) {                     // the block's previous chunk: outputs out, inputs long consumed
                int rc = drain(c - n_streams);
                if (rc) return rc;
            }
            memcpy(hs.base + L.o_boxes, src_boxes, sizeof(float) * 12 * nm);
            memcpy(hs.base + L.o_dims, src_dims, sizeof(float) * 3 * nm);
            memcpy(hs.base + L.o_orient, src_orient, sizeof(int32_t) * nm);
            memcpy(hs.base + L.o_pinv, src_pinv, sizeof(float) * 12 * nb);
            GPP_CUDA(cudaMemcpyAsync(dev, hs.base, L.in_end, cudaMemcpyHostToDevice, s));
        } else {
            GPP_CUDA(cudaMemcpyAsync(dev + L.o_boxes, src_boxes, sizeof(float) * 12 * nm, cudaMemcpyHostToDevice, s));
            GPP_CUDA(cudaMemcpyAsync(dev + L.o_dims, src_dims, sizeof(float) * 3 * nm, cudaMemcpyHostToDevice, s));
            GPP_CUDA(cudaMemcpyAsync(dev + L.o_orient, src_orient, sizeof(int32_t) * nm, cudaMemcpyHostToDevice, s));
            GPP_CUDA(cudaMemcpyAsync(dev + L.o_pinv, src_pinv, sizeof(float) * 12 * nb, cudaMemcpyHostToDevice, s));
        }
        gpp::FitIO io;
        io.boxes = reinterpret_cast<const float *>(dev + L.o_boxes);
        io.dims = reinterpret_cast<const float *>(dev + L.o_dims);
        io.orient = reinterpret_cast<const int32_t *>(dev + L.o_orient);
        io.pinv = reinterpret_cast<const float *>(dev + L.o_pinv);
        io.D = D; io.n_det = nm;
        io.keypoints = dev + L.o_kp; io.keyplanes = dev + L.o_kpl; io.residuals = dev + L.o_res;
        io.best = best ? reinterpret_cast<long long *>(dev + L.o_best) : nullptr;
        if (with_pose) {
            io.pose_locations = reinterpret_cast<float *>(dev + L.o_loc);
            io.pose_angles = reinterpret_cast<float *>(dev + L.o_ang);
            io.pose_dimensions = reinterpret_cast<float *>(dev + L.o_pdim);
            io.pose_kitti = with_kitti ? reinterpret_cast<float *>(dev + L.o_kitti) : nullptr;
        }
        if (h->planes_pending && h->planes_stream != s) GPP_CUDA(cudaStreamWaitEvent(s, h->planes_ready, 0));
        GPP_CUDA(cudaEventRecord(h->chunk_events[c].first, s));
        int rc = gpp::launch_poll(h, io, mode, s);
        if (rc) return rc;
        GPP_CUDA(cudaEventRecord(h->chunk_events[c].second, s));
        if (staged) {
            GPP_CUDA(cudaMemcpyAsync(hs.base + L.o_kp, dev + L.o_kp, L.o_end - L.o_kp, cudaMemcpyDeviceToHost, s));
            GPP_CUDA(cudaEventRecord(hs.done, s));
        } else {
            GPP_CUDA(cudaMemcpyAsync(keypoints + 12 * m0, dev + L.o_kp, sizeof(T) * 12 * nm, cudaMemcpyDeviceToHost, s));
            GPP_CUDA(cudaMemcpyAsync(keyplanes + 4 * m0, dev + L.o_kpl, sizeof(T) * 4 * nm, cudaMemcpyDeviceToHost, s));
            GPP_CUDA(cudaMemcpyAsync(residuals + m0, dev + L.o_res, sizeof(T) * nm, cudaMemcpyDeviceToHost, s));
            if (best)
                GPP_CUDA(cudaMemcpyAsync(best + m0, dev + L.o_best, sizeof(long long) * nm, cudaMemcpyDeviceToHost, s));
            if (with_pose) {
                GPP_CUDA(cudaMemcpyAsync(pose.locations + 3 * m0, dev + L.o_loc, sizeof(float) * 3 * nm, cudaMemcpyDeviceToHost, s));
                GPP_CUDA(cudaMemcpyAsync(pose.angles + 3 * m0, dev + L.o_ang, sizeof(float) * 3 * nm, cudaMemcpyDeviceToHost, s));
                GPP_CUDA(cudaMemcpyAsync(pose.dimensions + 3 * m0, dev + L.o_pdim, sizeof(float) * 3 * nm, cudaMemcpyDeviceToHost, s));
            }
            if (with_kitti)
                GPP_CUDA(cudaMemcpyAsync(pose.kitti + 4 * m0, dev + L.o_kitti, sizeof(float) * 4 * nm, cudaMemcpyDeviceToHost, s));
        }
    }
    if (while_the_gpu_works) (*while_the_gpu_works)();          // everything is enqueued: host work that overlaps it
    if (staged)
        for (int c = (n_chunks > n_streams ? n_chunks - n_streams : 0); c < n_chunks; ++c) {
            int rc = drain(c);
            if (rc) return rc;
        }
    for (int i = 0; i < n_streams; ++i) GPP_CUDA(cudaStreamSynchronize(h->streams[i]));
    h->timing_chunks = n_chunks;
    h->timing_single = false;
    return GPP_OK;
}

extern "C" {

int gpp_fit_host(gpp_handle *h, const float *boxes, const float *dimensions, const int32_t *orientations,
                 const float *P_inv, int B, int D, float *keypoints, float *keyplanes, float *residuals,
                 int64_t *best_index, int mode) {
    int rc = check_fit_args(h, boxes, dimensions, orientations, P_inv, B, D, keypoints, keyplanes, residuals,
                            "gpp_fit_host");
    if (rc) return rc;
    if (!float_mode(mode))
        return set_error(GPP_EINVAL, "gpp_fit_host: mode %d (use gpp_fit_host_f64 for the FP64 mode)", mode);
    return fit_host_impl<float>(h, boxes, dimensions, orientations, P_inv, B, D, keypoints, keyplanes, residuals,
                                best_index, mode);
}

int gpp_fit_pose_host(gpp_handle *h, const float *boxes, const float *dimensions, const int32_t *orientations,
                      const float *P_inv, int B, int D, float *keypoints, float *keyplanes, float *residuals,
                      int64_t *best_index, float *locations, float *angles, float *dimensions_out, float *kitti,
                      int mode) {
    int rc = check_fit_args(h, boxes, dimensions, orientations, P_inv, B, D, keypoints, keyplanes, residuals,
                            "gpp_fit_pose_host");
    if (rc) return rc;
    if (!float_mode(mode)) return set_error(GPP_EINVAL, "gpp_fit_pose_host: mode %d", mode);
    if ((long long)B * D > 0 && (!locations || !angles || !dimensions_out))
        return set_error(GPP_EINVAL, "gpp_fit_pose_host: NULL pose array");
    PoseHost pose;
    pose.locations = locations; pose.angles = angles; pose.dimensions = dimensions_out; pose.kitti = kitti;
    return fit_host_impl<float>(h, boxes, dimensions, orientations, P_inv, B, D, keypoints, keyplanes, residuals,
                                best_index, mode, pose);
}

// The reference's callers feed the plane database with every image (run_network.py:105, utils/eval.py:91), and one image
// is a 0.04 ms kernel: the "same database as last time?" comparison (17 us for the 22k database) is the largest host
// cost of such a call.  Here it runs WHILE the GPU polls against the resident database; if the bytes turn out to differ
// the database is uploaded and the call polled again (the first results are simply overwritten).  Larger calls, and
// calls whose database cannot be the resident one (other size / type / order, none resident), take the plain order.
int gpp_fit_planes_host(gpp_handle *h, const void *planes, int n_planes, int dtype, int order, const float *boxes,
                        const float *dimensions, const int32_t *orientations, const float *P_inv, int B, int D,
                        float *keypoints, float *keyplanes, float *residuals, int64_t *best_index, float *locations,
                        float *angles, float *dimensions_out, float *kitti, int mode) {
    if (!h || !planes || n_planes <= 0 || dtype < 0 || dtype > 1 || order < 0 || order > 1)
        return set_error(GPP_EINVAL, "gpp_fit_planes_host: bad plane database argument");
    if (!float_mode(mode)) return set_error(GPP_EINVAL, "gpp_fit_planes_host: mode %d", mode);
    const bool with_pose = locations || angles || dimensions_out || kitti;
    if (with_pose && (long long)B * D > 0 && (!locations || !angles || !dimensions_out))
        return set_error(GPP_EINVAL, "gpp_fit_planes_host: NULL pose array");
    PoseHost pose;
    if (with_pose) { pose.locations = locations; pose.angles = angles; pose.dimensions = dimensions_out; pose.kitti = kitti; }
    const size_t bytes = (dtype ? sizeof(double) : sizeof(float)) * 4 * (size_t)n_planes;
    const bool may_be_resident = h->raw_valid && h->n_planes == n_planes && h->raw_dtype == dtype && h->raw_order == order &&
                                 h->raw_copy.size() == bytes;
    const bool small = (long long)B * D > 0 && (long long)B * D <= 16384;
    int rc;
    if (may_be_resident && small) {
        rc = check_fit_args(h, boxes, dimensions, orientations, P_inv, B, D, keypoints, keyplanes, residuals, "gpp_fit_planes_host");
        if (rc) return rc;
        bool same = true;
        const std::function<void()> compare = [&] { same = memcmp(h->raw_copy.data(), planes, bytes) == 0; };
        rc = fit_host_impl<float>(h, boxes, dimensions, orientations, P_inv, B, D, keypoints, keyplanes, residuals, best_index,
                                  mode, pose, &compare);
        if (rc || same) return rc;
    }
    rc = gpp_set_planes_raw(h, planes, n_planes, dtype, order);
    if (rc) return rc;
    rc = check_fit_args(h, boxes, dimensions, orientations, P_inv, B, D, keypoints, keyplanes, residuals, "gpp_fit_planes_host");
    if (rc) return rc;
    return fit_host_impl<float>(h, boxes, dimensions, orientations, P_inv, B, D, keypoints, keyplanes, residuals, best_index,
                                mode, pose);
}

// One caller, several GPUs: the images are split into contiguous shards (the first B % n handles get one image more),
// every handle polls its shard from its own host thread (plain C++ threads: no interpreter lock) and writes into its
// slice of the caller's arrays.  There is no cross-device step on this path.  The worker threads are
// started once and parked between calls (starting and joining eight threads cost 0.1 ms of a 3.3 ms call); with
// `planes` the "same database as last time?" comparison and a possible upload run in the shard's thread too (one after
// the other in the caller they were another 0.15 ms).
namespace {
struct MultiPool {
    std::mutex mu;
    std::condition_variable wake, done;
    std::vector<std::thread> workers;
    std::function<void(int)> job;
    unsigned long long generation = 0;
    int n_jobs = 0, pending = 0;
    bool stop = false;
    void worker(int id) {
        unsigned long long seen = 0;
        for (;;) {
            std::function<void(int)> fn;
            {
                std::unique_lock<std::mutex> lk(mu);
                wake.wait(lk, [&] { return stop || (generation != seen && id < n_jobs); });
                if (stop) return;
                seen = generation;
                fn = job;
            }
            fn(id);
            {
                std::lock_guard<std::mutex> lk(mu);
                if (--pending == 0) done.notify_all();
            }
        }
    }
    // runs fn(1) .. fn(n - 1) on the parked threads and fn(0) on the caller's
    void run(int n, const std::function<void(int)> &fn) {
        {
            std::lock_guard<std::mutex> lk(mu);
            while ((int)workers.size() < n - 1) {
                const int id = (int)workers.size() + 1;
                workers.emplace_back([this, id] { worker(id); });
            }
            job = fn;
            n_jobs = n;
            pending = n - 1;
            ++generation;
        }
        wake.notify_all();
        fn(0);
        std::unique_lock<std::mutex> lk(mu);
        done.wait(lk, [&] { return pending == 0; });
        n_jobs = 0;
    }
    ~MultiPool() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        wake.notify_all();
        for (auto &t : workers) t.join();
    }
};
MultiPool &multi_pool() {
    static MultiPool pool;
    return pool;
}
std::mutex g_multi_call;          // one multi-device call at a time per process (the pool holds one job)
}  // namespace

static int fit_host_multi_impl(gpp_handle **handles, int n_handles, const void *planes, int n_planes, int dtype, int order,
                               const float *boxes, const float *dimensions, const int32_t *orientations,
                               const float *P_inv, int B, int D, float *keypoints, float *keyplanes, float *residuals,
                               int64_t *best_index, int mode, const char *who) {
    if (!handles || n_handles < 1) return set_error(GPP_EINVAL, "%s: no handles", who);
    if (!float_mode(mode)) return set_error(GPP_EINVAL, "%s: mode %d", who, mode);
    for (int i = 0; i < n_handles; ++i) {
        if (!handles[i]) return set_error(GPP_EINVAL, "%s: handle %d is NULL", who, i);
        for (int k = 0; k < i; ++k)
            if (handles[k] == handles[i]) return set_error(GPP_EINVAL, "%s: handle %d passed twice", who, i);
    }
    std::vector<int> rcs(n_handles, GPP_OK);
    std::vector<std::string> msgs(n_handles);
    auto shard = [&](int i) {
        if (planes) {
            rcs[i] = gpp_set_planes_raw(handles[i], planes, n_planes, dtype, order);
            if (rcs[i] != GPP_OK) { msgs[i] = g_last_error; return; }
        }
        rcs[i] = check_fit_args(handles[i], boxes, dimensions, orientations, P_inv, B, D, keypoints, keyplanes, residuals, who);
        if (rcs[i] != GPP_OK) { msgs[i] = g_last_error; return; }
        const int base = B / n_handles, extra = B % n_handles;
        const int b0 = i * base + (i < extra ? i : extra), nb = base + (i < extra ? 1 : 0);
        if (nb == 0 || D == 0) return;
        const size_t m0 = (size_t)b0 * D;
        rcs[i] = fit_host_impl<float>(handles[i], boxes + 12 * m0, dimensions + 3 * m0, orientations + m0,
                                      P_inv + 12 * (size_t)b0, nb, D, keypoints + 12 * m0, keyplanes + 4 * m0,
                                      residuals + m0, best_index ? best_index + m0 : nullptr, mode);
        if (rcs[i] != GPP_OK) msgs[i] = g_last_error;          // the error text is thread-local
    };
    if (n_handles == 1) {
        shard(0);
    } else {
        std::lock_guard<std::mutex> one(g_multi_call);
        multi_pool().run(n_handles, shard);
    }
    for (int i = 0; i < n_handles; ++i)
        if (rcs[i] != GPP_OK) return set_error(rcs[i], "%s: shard %d: %s", who, i, msgs[i].c_str());
    return GPP_OK;
}

int gpp_fit_host_multi(gpp_handle **handles, int n_handles, const float *boxes, const float *dimensions,
                       const int32_t *orientations, const float *P_inv, int B, int D, float *keypoints,
                       float *keyplanes, float *residuals, int64_t *best_index, int mode) {
    return fit_host_multi_impl(handles, n_handles, nullptr, 0, 0, 0, boxes, dimensions, orientations, P_inv, B, D, keypoints,
                               keyplanes, residuals, best_index, mode, "gpp_fit_host_multi");
}

int gpp_fit_host_multi_planes(gpp_handle **handles, int n_handles, const void *planes, int n_planes, int dtype, int order,
                              const float *boxes, const float *dimensions, const int32_t *orientations,
                              const float *P_inv, int B, int D, float *keypoints, float *keyplanes, float *residuals,
                              int64_t *best_index, int mode) {
    if (!planes || n_planes <= 0 || dtype < 0 || dtype > 1 || order < 0 || order > 1)
        return set_error(GPP_EINVAL, "gpp_fit_host_multi_planes: bad plane database argument");
    return fit_host_multi_impl(handles, n_handles, planes, n_planes, dtype, order, boxes, dimensions, orientations, P_inv, B, D,
                               keypoints, keyplanes, residuals, best_index, mode, "gpp_fit_host_multi_planes");
}

int gpp_fit_host_f64(gpp_handle *h, const float *boxes, const float *dimensions, const int32_t *orientations,
                     const float *P_inv, int B, int D, double *keypoints, double *keyplanes,
                     double *residuals, int64_t *best_index) {
    int rc = check_fit_args(h, boxes, dimensions, orientations, P_inv, B, D, keypoints, keyplanes, residuals,
                            "gpp_fit_host_f64");
    if (rc) return rc;
    return fit_host_impl<double>(h, boxes, dimensions, orientations, P_inv, B, D, keypoints, keyplanes, residuals,
                                 best_index, GPP_MODE_F64);
}

// ---------------------------------------------------------------------------------------------------
int gpp_last_kernel_ms(gpp_handle *h, float *ms) {
    if (!h || !ms) return set_error(GPP_EINVAL, "gpp_last_kernel_ms: bad argument");
    DeviceGuard guard(h->device);
    float total = 0.f;
    if (h->timing_single) {
        GPP_CUDA(cudaEventSynchronize(h->ev_stop));
        GPP_CUDA(cudaEventElapsedTime(&total, h->ev_start, h->ev_stop));
    } else {
        for (int c = 0; c < h->timing_chunks; ++c) {
            float t = 0.f;
            GPP_CUDA(cudaEventSynchronize(h->chunk_events[c].second));
            GPP_CUDA(cudaEventElapsedTime(&t, h->chunk_events[c].first, h->chunk_events[c].second));
            total += t;
        }
    }
    *ms = total;
    return GPP_OK;
}

int64_t gpp_launch_count(const gpp_handle *h) { return h ? h->launches : 0; }

int gpp_debug_scores(gpp_handle *h, const float *box12, const float *dims3, int orientation, const float *pinv12,
                     int which, int32_t *votes, float *resid, int32_t *zneg, float *margin) {
    if (!h || !box12 || !dims3 || !pinv12 || !votes || !resid || !zneg || h->n_planes <= 0 || which < 0 || which > 3)
        return set_error(GPP_EINVAL, "gpp_debug_scores: bad argument");
    DeviceGuard guard(h->device);
    const int n = h->n_planes;
    float host_det[27];
    memcpy(host_det, box12, 12 * sizeof(float));
    memcpy(host_det + 12, dims3, 3 * sizeof(float));
    memcpy(host_det + 15, pinv12, 12 * sizeof(float));
    float *d_det = nullptr, *d_res = nullptr;
    int32_t *d_int = nullptr;
    cudaError_t e = cudaMalloc(&d_det, sizeof(host_det));
    if (e == cudaSuccess) e = cudaMalloc(&d_res, sizeof(float) * 2 * (size_t)n);
    if (e == cudaSuccess) e = cudaMalloc(&d_int, sizeof(int32_t) * (2 * (size_t)n + 1));
    int rc = GPP_OK;
    if (e == cudaSuccess) {
        cudaStream_t s = h->streams[0];
        int32_t o = orientation;
        e = cudaMemcpyAsync(d_det, host_det, sizeof(host_det), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_int + 2 * (size_t)n, &o, sizeof(o), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess && margin) e = cudaMemsetAsync(d_res + n, 0, sizeof(float) * n, s);
        if (e == cudaSuccess)
            rc = gpp::launch_scores(h, d_det, d_int + 2 * (size_t)n, which, d_int, d_res, d_int + n,
                                    (margin && which > 0) ? d_res + n : nullptr, s);
        if (e == cudaSuccess && rc == GPP_OK) {
            e = cudaMemcpyAsync(votes, d_int, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaMemcpyAsync(zneg, d_int + n, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaMemcpyAsync(resid, d_res, sizeof(float) * n, cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess && margin)
                e = cudaMemcpyAsync(margin, d_res + n, sizeof(float) * n, cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        }
    }
    cudaFree(d_det); cudaFree(d_res); cudaFree(d_int);
    if (rc != GPP_OK) return rc;
    if (e != cudaSuccess) return set_error(GPP_ECUDA, "gpp_debug_scores: %s", cudaGetErrorString(e));
    return GPP_OK;
}

int gpp_debug_set_schedule(gpp_handle *h, int n_seg, int resident_rows) {
    if (!h || n_seg < 0 || n_seg > 32) return set_error(GPP_EINVAL, "gpp_debug_set_schedule: bad argument");
    h->force_seg = n_seg;
    h->force_resident = resident_rows < 0 ? -1 : resident_rows;
    return GPP_OK;
}

int gpp_audit_set(gpp_handle *h, int every) {
    if (!h || every < 0) return set_error(GPP_EINVAL, "gpp_audit_set: bad argument");
    h->audit_every = every;
    return GPP_OK;
}

int gpp_audit_counts(gpp_handle *h, int64_t *checked, int64_t *mismatches) {
    if (!h) return set_error(GPP_EINVAL, "gpp_audit_counts: handle is NULL");
    DeviceGuard guard(h->device);
    unsigned long long c[2] = {0, 0};
    if (h->audit_counts) {
        GPP_CUDA(cudaDeviceSynchronize());
        GPP_CUDA(cudaMemcpy(c, h->audit_counts, sizeof(c), cudaMemcpyDeviceToHost));
    }
    if (checked) *checked = (int64_t)c[0];
    if (mismatches) *mismatches = (int64_t)c[1];
    return GPP_OK;
}

}  // extern "C"
