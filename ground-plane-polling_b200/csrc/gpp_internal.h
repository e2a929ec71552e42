// Internal declarations shared by the libgpp translation units (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>

#include <utility>
#include <vector>

#include "gpp_poll3.cuh"

namespace gpp {

int set_error(int code, const char *fmt, ...);

// device-side staging block of one stream of the host entry point (one chunk: inputs, then outputs; see ChunkLayout)
struct Staging {
    unsigned char *base = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    void release();
};

// Pinned host staging of one chunk (inputs and outputs) for callers that pass pageable memory: an asynchronous copy
// to or from pageable memory blocks the host until it is done, which serialises the chunks of a large call; with
// the staging block the host only does plain memcpy while the GPU works on the other chunk.
struct HostStaging {
    unsigned char *base = nullptr;
    size_t cap = 0;
    cudaEvent_t done = nullptr;      // recorded after the chunk's last device-to-host copy
    int reserve(size_t bytes);
    void release();
};

// One polling call on device memory: what the reference's operator takes and returns (fit_road_planes.py:49-61,
// :139) plus the optional extensions (arg-min index; pose and KITTI record of every row).
struct FitIO {
    const float *boxes = nullptr, *dims = nullptr, *pinv = nullptr;
    const int32_t *orient = nullptr;
    int D = 0;                        // detections per image
    long long n_det = 0;              // B * D
    void *keypoints = nullptr, *keyplanes = nullptr, *residuals = nullptr;   // float, or double in the F64 mode
    long long *best = nullptr;
    float *pose_locations = nullptr, *pose_angles = nullptr, *pose_dimensions = nullptr, *pose_kitti = nullptr;
};

}  // namespace gpp

struct gpp_handle {
    static constexpr int kStreams = 2;
    int device = 0;
    int sm_count = 0;
    // plane database (device): raw upload, normalised fp32 (float4) and fp64 (double4) copies, pair-interleaved fp32
    // copy padded to 64 planes (gpp_poll2.cuh)
    float *d_raw = nullptr;
    float4 *d_planes32 = nullptr;
    float4 *d_planes32_scan = nullptr;       // d_planes32 in scan order (gpp_order.cu)
    double4 *d_planes64 = nullptr;
    unsigned long long *d_pairs = nullptr;
    int32_t *d_scan_index = nullptr;         // [2 * n_pairs_padded] plane index of every position of d_pairs (gpp_order.cu)
    int n_pairs_padded = 0;
    int n_planes = 0, cap_planes = 0;
    // bytes of the last host upload as the caller passed them (gpp_set_planes_raw compares before doing any work)
    std::vector<unsigned char> raw_copy;
    int raw_dtype = 0, raw_order = 0;
    bool raw_valid = false;
    // a device-side database update (gpp_set_planes_device) that fits on other streams have to wait for
    cudaEvent_t planes_ready = nullptr;
    cudaStream_t planes_stream = nullptr;
    bool planes_pending = false;
    // polling kernel (gpp_poll3.cuh): self-resetting counters and the scratch of segmented detections, one set per
    // slot so that launches in flight on different streams never share one
    struct Slot3 {
        unsigned long long *claim = nullptr;     // [2]
        gpp::SegPartial *partials = nullptr;     // [seg_items_cap]
        unsigned int *seg_arrived = nullptr;     // [seg_det_cap]
        unsigned long long *seg_best = nullptr;  // [seg_det_cap]
        float *seg_consts = nullptr;             // [seg_det_cap][kSharedConsts]
        unsigned int *seg_ready = nullptr;       // [seg_det_cap]
        cudaEvent_t done = nullptr;
        bool used = false;
    };
    static constexpr int kSlots3 = 4;
    Slot3 slot3[kSlots3];
    unsigned next_slot3 = 0;
    long long seg_det_cap = 0, seg_items_cap = 0;
    int resident_cap_rows = 0;               // rows of the pair database that fit into one SM's shared memory
    int force_seg = 0, force_resident = -1;  // tuning / test hook (gpp_debug_set_schedule): 0 / -1 = automatic
    // runtime audit of the VERIFIED mode (gpp_audit_set): every n-th detection re-polled in the EXACT arithmetic
    int audit_every = 0;
    long long audit_cap = 0;
    float *audit_out = nullptr;              // key-points, key-planes, residuals of the audit pass (17 floats per row)
    long long *audit_best = nullptr, *audit_best_main = nullptr;
    unsigned long long *audit_counts = nullptr;   // [0] rows checked, [1] rows that differ
    cudaEvent_t audit_done = nullptr;
    bool audit_used = false;
    // host-entry plumbing
    cudaStream_t streams[kStreams] = {nullptr, nullptr};
    gpp::Staging stage[kStreams];
    gpp::HostStaging hstage[kStreams];       // pinned
    std::vector<std::pair<cudaEvent_t, cudaEvent_t> > chunk_events;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    int timing_chunks = 0;
    bool timing_single = false;
    int64_t launches = 0;
    // FilterDetections scratch (gpp_detect.cu): candidate keys, per-anchor orientation, per-image counters
    unsigned long long *filter_keys = nullptr;
    unsigned char *filter_orient = nullptr;
    unsigned int *filter_counts = nullptr;
    size_t filter_keys_bytes = 0, filter_orient_bytes = 0;
    int filter_counts_n = 0;
};

namespace gpp {

int configure_kernels(gpp_handle *h);
void release_poll3(gpp_handle *h);
void release_audit(gpp_handle *h);
// order[position] = plane index: the order in which the pair database is stored and scanned (gpp_order.cu)
void scan_order(const float *rows, int n, std::vector<int32_t> &order);
// fills d_scan_index (from `order`, or with the identity when it is null) and d_pairs from d_planes32, on `s`
int build_pairs(gpp_handle *h, const int32_t *order, cudaStream_t s);
// one polling call (any GPP_MODE_*) on device memory, enqueued on `s`; VERIFIED calls are followed by the audit pass
// when the handle asks for it
int launch_poll(gpp_handle *h, const FitIO &io, int mode, cudaStream_t s);
int launch_scores(gpp_handle *h, const float *d_det, const int32_t *d_orient, int which, int32_t *votes,
                  float *resid, int32_t *zneg, float *margin, cudaStream_t s);

}  // namespace gpp
