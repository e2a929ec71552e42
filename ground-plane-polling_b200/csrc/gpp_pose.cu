// Stand-alone entry points of the two steps after polling (C ABI: gpp_pose_*, gpp_kitti_*): one thread per detection
// around the device functions of gpp_pose.cuh.  The polling kernel can run the same functions in its epilogue
// (gpp_fit_pose_*), which saves the launches and the round trip of the key-points.
#include "../../include/gpp.h"
#include "gpp_internal.h"
#include "gpp_pose.cuh"

namespace gpp {

__global__ void pose_kernel(const float *__restrict__ keypoints, const float *__restrict__ dims,
                            const int32_t *__restrict__ orient, long long n, float *__restrict__ locations,
                            float *__restrict__ angles, float *__restrict__ dims_out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int o = orient[i];
    if (o < 0 || o > 3) return;                                   // reference leaves such rows untouched
    float kp[12], out9[9];
#pragma unroll
    for (int k = 0; k < 12; ++k) kp[k] = keypoints[12 * i + k];
    pose_from_keypoints(kp, dims[3 * i + 1], o, out9);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        locations[3 * i + k] = out9[k];
        angles[3 * i + k] = out9[3 + k];
        dims_out[3 * i + k] = out9[6 + k];
    }
}

__global__ void kitti_kernel(const float *__restrict__ locations, const float *__restrict__ angles,
                             const float *__restrict__ dims, long long n, float *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float pose9[9], rec[4];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        pose9[k] = locations[3 * i + k];
        pose9[3 + k] = angles[3 * i + k];
        pose9[6 + k] = dims[3 * i + k];
    }
    kitti_record(pose9, rec);
#pragma unroll
    for (int k = 0; k < 4; ++k) out[4 * i + k] = rec[k];
}

}  // namespace gpp

extern "C" {

int gpp_kitti_device(gpp_handle *h, const float *locations, const float *angles, const float *dimensions, long n,
                     float *out, void *stream) {
    if (!h || n < 0) return gpp::set_error(GPP_EINVAL, "gpp_kitti_device: bad argument");
    if (n == 0) return GPP_OK;
    if (!locations || !angles || !dimensions || !out)
        return gpp::set_error(GPP_EINVAL, "gpp_kitti_device: NULL array argument");
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(h->device);
    const int threads = 128;
    gpp::kitti_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, static_cast<cudaStream_t>(stream)>>>(
        locations, angles, dimensions, n, out);
    h->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (prev >= 0) cudaSetDevice(prev);
    if (e != cudaSuccess) return gpp::set_error(GPP_ECUDA, "kitti_kernel launch: %s", cudaGetErrorString(e));
    return GPP_OK;
}

int gpp_kitti_host(gpp_handle *h, const float *locations, const float *angles, const float *dimensions, long n,
                   float *out) {
    if (!h || n < 0) return gpp::set_error(GPP_EINVAL, "gpp_kitti_host: bad argument");
    if (n == 0) return GPP_OK;
    if (!locations || !angles || !dimensions || !out)
        return gpp::set_error(GPP_EINVAL, "gpp_kitti_host: NULL array argument");
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(h->device);
    float *d = nullptr;
    cudaError_t e = cudaMalloc(&d, sizeof(float) * 13 * (size_t)n);
    int rc = GPP_OK;
    if (e == cudaSuccess) {
        cudaStream_t s = h->streams[0];
        float *d_loc = d, *d_ang = d + 3 * (size_t)n, *d_dim = d + 6 * (size_t)n, *d_out = d + 9 * (size_t)n;
        e = cudaMemcpyAsync(d_loc, locations, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_ang, angles, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_dim, dimensions, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) rc = gpp_kitti_device(h, d_loc, d_ang, d_dim, n, d_out, s);
        if (e == cudaSuccess && rc == GPP_OK) e = cudaMemcpyAsync(out, d_out, sizeof(float) * 4 * (size_t)n, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess && rc == GPP_OK) e = cudaStreamSynchronize(s);
    }
    cudaFree(d);
    if (prev >= 0) cudaSetDevice(prev);
    if (rc != GPP_OK) return rc;
    if (e != cudaSuccess) return gpp::set_error(GPP_ECUDA, "gpp_kitti_host: %s", cudaGetErrorString(e));
    return GPP_OK;
}

int gpp_pose_device(gpp_handle *h, const float *keypoints, const float *dimensions,
                    const int32_t *orientations, long n, float *locations, float *angles,
                    float *dimensions_out, void *stream) {
    if (!h || n < 0) return gpp::set_error(GPP_EINVAL, "gpp_pose_device: bad argument");
    if (n == 0) return GPP_OK;
    if (!keypoints || !dimensions || !orientations || !locations || !angles || !dimensions_out)
        return gpp::set_error(GPP_EINVAL, "gpp_pose_device: NULL array argument");
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(h->device);
    const int threads = 128;
    const long long blocks = (n + threads - 1) / threads;
    gpp::pose_kernel<<<(unsigned)blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(
        keypoints, dimensions, orientations, n, locations, angles, dimensions_out);
    h->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (prev >= 0) cudaSetDevice(prev);
    if (e != cudaSuccess) return gpp::set_error(GPP_ECUDA, "pose_kernel launch: %s", cudaGetErrorString(e));
    return GPP_OK;
}

int gpp_pose_host(gpp_handle *h, const float *keypoints, const float *dimensions, const int32_t *orientations,
                  long n, float *locations, float *angles, float *dimensions_out) {
    if (!h || n < 0) return gpp::set_error(GPP_EINVAL, "gpp_pose_host: bad argument");
    if (n == 0) return GPP_OK;
    if (!keypoints || !dimensions || !orientations || !locations || !angles || !dimensions_out)
        return gpp::set_error(GPP_EINVAL, "gpp_pose_host: NULL array argument");
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(h->device);
    float *d = nullptr;
    int32_t *d_or = nullptr;
    const size_t nf = (size_t)n * (12 + 3 + 3 + 3 + 3);
    cudaError_t e = cudaMalloc(&d, nf * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&d_or, (size_t)n * sizeof(int32_t));
    int rc = GPP_OK;
    if (e == cudaSuccess) {
        float *d_kp = d, *d_dims = d + 12 * (size_t)n, *d_loc = d_dims + 3 * (size_t)n,
              *d_ang = d_loc + 3 * (size_t)n, *d_dout = d_ang + 3 * (size_t)n;
        cudaStream_t s = h->streams[0];
        e = cudaMemcpyAsync(d_kp, keypoints, sizeof(float) * 12 * (size_t)n, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_dims, dimensions, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_or, orientations, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, s);
        // rows with an orientation outside 0..3 keep whatever the caller's output buffers hold
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_loc, locations, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_ang, angles, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_dout, dimensions_out, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) rc = gpp_pose_device(h, d_kp, d_dims, d_or, n, d_loc, d_ang, d_dout, s);
        if (e == cudaSuccess && rc == GPP_OK) {
            e = cudaMemcpyAsync(locations, d_loc, sizeof(float) * 3 * (size_t)n, cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaMemcpyAsync(angles, d_ang, sizeof(float) * 3 * (size_t)n, cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaMemcpyAsync(dimensions_out, d_dout, sizeof(float) * 3 * (size_t)n, cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        }
    }
    cudaFree(d);
    cudaFree(d_or);
    if (prev >= 0) cudaSetDevice(prev);
    if (rc != GPP_OK) return rc;
    if (e != cudaSuccess) return gpp::set_error(GPP_ECUDA, "gpp_pose_host: %s", cudaGetErrorString(e));
    return GPP_OK;
}

}  // extern "C"
