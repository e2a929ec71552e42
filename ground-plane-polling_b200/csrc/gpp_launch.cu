// Kernel instantiations and launch configuration of the polling kernels (gpp_poll2.cuh: packed-pair fp32
// modes; gpp_poll.cuh: scalar kernel, used for the FP64 verify mode).
#include "../../include/gpp.h"
#include "gpp_internal.h"

namespace gpp {

// CTA shape: 8 warps, 1024-plane tiles (32 KB fp32 pairs / 32 KB fp64), 3-stage TMA ring.
constexpr int kWarps = 8;
constexpr int kTile32 = 1024, kTile64 = 512;
constexpr int kStages = 3;

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

constexpr size_t kSmem2 = size_t(32) * kStages * (kTile32 / 2) + 2 * kStages * sizeof(uint64_t);
constexpr size_t kSmem64 = sizeof(double4) * kStages * kTile64 + 2 * kStages * sizeof(uint64_t);
#define GPP_K_EXACT poll2_kernel<PackExact, kWarps, kTile32, kStages, GPP_MINB_EXACT>
#define GPP_K_FAST poll2_kernel<PackFast, kWarps, kTile32, kStages, GPP_MINB_FAST>
#define GPP_K_F64 poll_kernel<ExactF64, kWarps, 1, kTile64, kStages>
#ifndef GPP_MINB_EXACT
#define GPP_MINB_EXACT 2
#endif
#ifndef GPP_MINB_FAST
#define GPP_MINB_FAST 2
#endif

int configure_kernels(gpp_handle *h) {
    int rc;
    if ((rc = configure_kernel(GPP_K_EXACT, kSmem2, &h->occ[0]))) return rc;
    if ((rc = configure_kernel(GPP_K_FAST, kSmem2, &h->occ[1]))) return rc;
    if ((rc = configure_kernel(GPP_K_F64, kSmem64, &h->occ[2]))) return rc;
    return GPP_OK;
}

int launch_poll_f32(gpp_handle *h, const PollArgs<float> &a, int mode, cudaStream_t s) {
    PollArgs2<float> b;
    b.boxes = a.boxes; b.dims = a.dims; b.pinv = a.pinv; b.orient = a.orient;
    b.pairs = h->d_pairs; b.planes = h->d_planes32; b.n_planes = h->n_planes;
    b.n_pairs_padded = h->n_pairs_padded; b.dets_per_image = a.dets_per_image; b.n_det = a.n_det;
    b.keypoints = a.keypoints; b.keyplanes = a.keyplanes; b.residuals = a.residuals; b.best = a.best;
    const long long n_groups = (a.n_det + kWarps - 1) / kWarps;
    if (mode == GPP_MODE_FAST) {
        GPP_K_FAST<<<(unsigned)grid_for(h, n_groups, h->occ[1]), kWarps * 32, kSmem2, s>>>(b);
    } else {
        GPP_K_EXACT<<<(unsigned)grid_for(h, n_groups, h->occ[0]), kWarps * 32, kSmem2, s>>>(b);
    }
    h->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(GPP_ECUDA, "poll2_kernel launch: %s", cudaGetErrorString(e));
    return GPP_OK;
}

int launch_poll_f64(gpp_handle *h, const PollArgs<double> &a, cudaStream_t s) {
    const long long n_groups = (a.n_det + kWarps - 1) / kWarps;
    GPP_K_F64<<<(unsigned)grid_for(h, n_groups, h->occ[2]), kWarps * 32, kSmem64, s>>>(a);
    h->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(GPP_ECUDA, "poll_kernel<f64> launch: %s", cudaGetErrorString(e));
    return GPP_OK;
}

}  // namespace gpp
