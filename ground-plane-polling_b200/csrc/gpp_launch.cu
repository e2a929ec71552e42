// Kernel instantiations and launch configuration of the polling kernel (gpp_poll.cuh).
#include "../../include/gpp.h"
#include "gpp_internal.h"

namespace gpp {

// CTA shape: 8 warps, 1024-plane tiles (16 KB fp32 / 32 KB fp64), 3-stage TMA ring.
constexpr int kWarps = 8;
constexpr int kTile32 = 1024, kTile64 = 512;
constexpr int kStages = 3;

template <class P, int kDpw, int kTile>
struct Cfg {
    static constexpr size_t smem = sizeof(typename P::T4) * kStages * kTile + 2 * kStages * sizeof(uint64_t);
    static auto kernel() { return poll_kernel<P, kWarps, kDpw, kTile, kStages>; }
    static int configure(int *occ) {
        cudaError_t e = cudaFuncSetAttribute(kernel(), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(GPP_ECUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kernel(), kWarps * 32, smem);
        if (e != cudaSuccess) return set_error(GPP_ECUDA, "occupancy query: %s", cudaGetErrorString(e));
        if (*occ < 1) return set_error(GPP_ECUDA, "polling kernel does not fit on an SM");
        return GPP_OK;
    }
    static int launch(gpp_handle *h, const PollArgs<typename P::T> &a, int occ, cudaStream_t s) {
        constexpr int kGroup = kWarps * kDpw;
        const long long n_groups = (a.n_det + kGroup - 1) / kGroup;
        int per_sm = h->force_ctas_per_sm > 0 ? h->force_ctas_per_sm : occ;
        long long grid = (long long)h->sm_count * per_sm;
        if (grid > n_groups) grid = n_groups;
        if (grid < 1) grid = 1;
        kernel()<<<(unsigned)grid, kWarps * 32, smem, s>>>(a);
        h->launches += 1;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return set_error(GPP_ECUDA, "poll_kernel launch: %s", cudaGetErrorString(e));
        return GPP_OK;
    }
};

int configure_kernels(gpp_handle *h) {
    int rc;
    if ((rc = Cfg<ExactF32, 1, kTile32>::configure(&h->occ[0][0]))) return rc;
    if ((rc = Cfg<ExactF32, 2, kTile32>::configure(&h->occ[0][1]))) return rc;
    if ((rc = Cfg<FastF32, 1, kTile32>::configure(&h->occ[1][0]))) return rc;
    if ((rc = Cfg<FastF32, 2, kTile32>::configure(&h->occ[1][1]))) return rc;
    if ((rc = Cfg<ExactF64, 1, kTile64>::configure(&h->occ[2][0]))) return rc;
    return GPP_OK;
}

// two detections per warp once every SM has more than a couple of groups to chew on
static int pick_dpw(const gpp_handle *h, long long n_det, int occ1) {
    if (h->force_dpw == 1 || h->force_dpw == 2) return h->force_dpw;
    const long long resident = (long long)h->sm_count * occ1 * kWarps;
    return n_det >= 4 * resident ? 2 : 1;
}

int launch_poll_f32(gpp_handle *h, const PollArgs<float> &a, int mode, cudaStream_t s) {
    const int mi = mode == GPP_MODE_FAST ? 1 : 0;
    const int dpw = pick_dpw(h, a.n_det, h->occ[mi][0]);
    if (mode == GPP_MODE_FAST) {
        return dpw == 2 ? Cfg<FastF32, 2, kTile32>::launch(h, a, h->occ[1][1], s)
                        : Cfg<FastF32, 1, kTile32>::launch(h, a, h->occ[1][0], s);
    }
    return dpw == 2 ? Cfg<ExactF32, 2, kTile32>::launch(h, a, h->occ[0][1], s)
                    : Cfg<ExactF32, 1, kTile32>::launch(h, a, h->occ[0][0], s);
}

int launch_poll_f64(gpp_handle *h, const PollArgs<double> &a, cudaStream_t s) {
    return Cfg<ExactF64, 1, kTile64>::launch(h, a, h->occ[2][0], s);
}

}  // namespace gpp
