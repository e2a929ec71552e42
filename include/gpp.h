/*
 * libgpp -- B200-native ground-plane polling, C ABI.
 *
 * This is the drop-in boundary for the one hot path of arangesh/Ground-Plane-Polling: the per-detection
 * search over the road-plane database.  The reference has no FFI for this path (it is a TensorFlow graph
 * built by keras_retinanet_3D/layers/fit_road_planes.py:49-139 and run inside
 * Model.predict_on_batch, keras_retinanet_3D/bin/run_network.py:110); the entry points below are what a
 * ctypes binding of that function needs, and each one cites the reference interface it replaces.
 * INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *   - plain pointers and sizes, no torch / numpy types; row-major (C order) dense arrays;
 *   - every function returns 0 on success, a GPP_E* code otherwise; gpp_last_error() gives the text
 *     (thread-local);
 *   - the caller owns every buffer; the library keeps no host pointer past the call; the device copy of
 *     the plane database is owned by the handle;
 *   - a handle is bound to one CUDA device and is not thread-safe (one handle per thread/GPU);
 *   - there is NO CPU fallback: without a usable sm_100 device gpp_create() fails.
 */
#ifndef GPP_H_
#define GPP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPP_VERSION 100 /* 0.1.0 */

enum {
    GPP_OK = 0,
    GPP_EINVAL = 1, /* bad argument (null pointer, negative size, unknown mode, no planes set) */
    GPP_ECUDA = 2,  /* CUDA runtime error, text in gpp_last_error() */
    GPP_ENODEV = 3, /* no CUDA device / not an sm_100 part */
    GPP_ENOMEM = 4
};

/* Arithmetic mode of the per-hypothesis math (the selection semantics are identical in all modes).
 *   EXACT  : IEEE fp32, round-to-nearest, no FMA contraction, correctly rounded div/sqrt, the op order of
 *            oracle/fit_road_planes_ref.py -> bit-identical to the oracle (index, key-points, residual).
 *   FAST   : fp32 with FMA contraction and MUFU reciprocal / square root for the SEARCH; the winner's
 *            key-points / residual are then recomputed with the EXACT arithmetic.  May pick a different
 *            plane only when two planes score within float rounding noise of each other.
 *   F64    : the same search in IEEE fp64 (verify mode for near-ties); only through gpp_fit_*_f64.
 *   VERIFIED: the FAST arithmetic is used only as a filter; every plane that, within a conservative bound on
 *            the fast-vs-exact deviation, could be the arg-min is re-evaluated in the EXACT arithmetic, and
 *            all selection state is kept in exact values -> same results as EXACT at close to FAST speed.
 */
enum { GPP_MODE_EXACT = 0, GPP_MODE_FAST = 1, GPP_MODE_F64 = 2, GPP_MODE_VERIFIED = 3 };

typedef struct gpp_handle gpp_handle;

int gpp_version(void);
const char *gpp_last_error(void);
/* Number of CUDA devices visible to the process (0 without a driver / GPU; never fails).  Multi-GPU callers create one
 * handle per device and shard the images: the path has no cross-device step. */
int gpp_device_count(void);

/* Create / destroy a polling context on CUDA device `device` (>= 0). */
int gpp_create(int device, gpp_handle **out);
int gpp_destroy(gpp_handle *h);

/* Upload one road-plane database: `planes` is host memory, N x 4 floats [a, b, c, d] per row, RAW (as
 * loaded from road_planes_database_*.mat, run_network.py:75, cast to float32 like the Keras feed does).
 * The sign flip and unit-normal normalisation of fit_road_planes.py:75-77 run once on the device.
 * Re-sending identical content is detected (byte comparison with the last upload) and costs no upload, because
 * every reference caller re-feeds the same database with every image (run_network.py:105,
 * preprocessing/kitti.py:220).  Replaces the `planes` input tensor of FitRoadPlanes.call
 * (fit_road_planes.py:152-163, models/retinanet.py:396). */
int gpp_set_planes(gpp_handle *h, const float *planes, int n_planes);
/* Same for the array exactly as the reference's callers hold it: `dtype` 0 = float32, 1 = float64 (what
 * scipy.io.loadmat returns, run_network.py:75); `order` 0 = row-major N x 4, 1 = column-major (Fortran order, again
 * what loadmat returns).  The cast to float32 is the one Keras applies at feed.  The library keeps a byte copy of the
 * last upload and compares before doing anything else, so that re-sending the same database with every image
 * (run_network.py:105) costs one memcmp. */
int gpp_set_planes_raw(gpp_handle *h, const void *planes, int n_planes, int dtype, int order);
/* Same, from a device pointer (stream-ordered on `stream`, no shortcut; fits of this handle still in flight on other
 * streams are waited for, and later fits on other streams wait for the update).  A database of 2048 planes or more is
 * read back to the host once per update to derive the order it is scanned in (csrc/gpp_order.cu): that update blocks the
 * caller until `stream` has caught up (inside a stream capture the index order is kept instead). */
int gpp_set_planes_device(gpp_handle *h, const float *d_planes, int n_planes, void *stream);
int gpp_num_planes(const gpp_handle *h);
/* Copy the normalised database (N x 4 floats) back to host -- the table `keyplanes` rows are taken from. */
int gpp_get_normalised_planes(gpp_handle *h, float *out);

/* fit_road_planes(boxes, dimensions, orientations, P_inv, planes) -> [keypoints, keyplanes, residuals]
 * (fit_road_planes.py:49-61) for B images x D detections against the database set with gpp_set_planes.
 *   boxes        B*D*12 floats (x1,y1,x2,y2,xl,yl,xm,ym,xr,yr,xt,yt)
 *   dimensions   B*D*3  floats (h,w,l)
 *   orientations B*D    int32  (class 0..3; -1 = padding row, processed like any other row)
 *   P_inv        B*4*3  floats (pseudo-inverse of the scaled projection matrix, run_network.py:57-58)
 *   keypoints    B*D*4*3 floats out   (X_l, X_m, X_r, X_t)
 *   keyplanes    B*D*1*4 floats out   (normalised, sign-flipped winning plane)
 *   residuals    B*D     floats out   (masked residual of the winner / 6)
 *   best_index   B*D     int64  out, may be NULL (extension: the argmin itself, fit_road_planes.py:119)
 * Host entry: pointers are host memory (pinned memory makes the copies asynchronous); the call returns
 * when the outputs are complete.  Host<->device copies are chunked and overlapped with the kernel. */
int gpp_fit_host(gpp_handle *h, const float *boxes, const float *dimensions, const int32_t *orientations,
                 const float *P_inv, int B, int D, float *keypoints, float *keyplanes, float *residuals,
                 int64_t *best_index, int mode);

/* gpp_set_planes_raw + gpp_fit_host (+ the pose / KITTI outputs of gpp_fit_pose_host; all four NULL = none) in one
 * call, for callers that feed the database with every image like the reference's (run_network.py:105, utils/eval.py:91):
 * for a small call the "same database as last time?" comparison runs on the host while the GPU already polls against
 * the resident database, and the call is polled again after an upload if the bytes turn out to differ.  Results are
 * those of the two separate calls. */
int gpp_fit_planes_host(gpp_handle *h, const void *planes, int n_planes, int dtype, int order, const float *boxes,
                        const float *dimensions, const int32_t *orientations, const float *P_inv, int B, int D,
                        float *keypoints, float *keyplanes, float *residuals, int64_t *best_index, float *locations,
                        float *angles, float *dimensions_out, float *kitti, int mode);

/* The same call over several GPUs of one box (BASELINE.json configs[3]: images sharded, database replicated, results
 * gathered into one host array): `handles` are n_handles contexts on different devices, each holding the database
 * (gpp_set_planes on every one).  Image shard i -- contiguous, the first B % n_handles shards one image longer -- is
 * polled by handles[i] from its own host thread and lands in its slice of the output arrays.  No collective. */
int gpp_fit_host_multi(gpp_handle **handles, int n_handles, const float *boxes, const float *dimensions,
                       const int32_t *orientations, const float *P_inv, int B, int D, float *keypoints,
                       float *keyplanes, float *residuals, int64_t *best_index, int mode);
/* Same, with the plane database of gpp_set_planes_raw in the same call: every handle compares it with what it holds (and
 * uploads it if it differs) from its shard's thread, so the callers' "feed the database with every batch"
 * (run_network.py:105) costs one concurrent memcmp per device instead of one after the other. */
int gpp_fit_host_multi_planes(gpp_handle **handles, int n_handles, const void *planes, int n_planes, int dtype, int order,
                              const float *boxes, const float *dimensions, const int32_t *orientations,
                              const float *P_inv, int B, int D, float *keypoints, float *keyplanes, float *residuals,
                              int64_t *best_index, int mode);

/* Device entry (the torch / DLPack path): all pointers are device memory on the handle's device; the
 * kernels are enqueued on `stream` (a cudaStream_t, NULL = legacy default stream) and the call returns
 * without synchronising. */
int gpp_fit_device(gpp_handle *h, const float *boxes, const float *dimensions, const int32_t *orientations,
                   const float *P_inv, int B, int D, float *keypoints, float *keyplanes, float *residuals,
                   int64_t *best_index, int mode, void *stream);

/* fit_road_planes + the two steps that follow it in the reference's driver, in ONE kernel launch: the polling kernel's
 * epilogue runs pose recovery (run_network.py:137-247) and, if `kitti` is not NULL, the KITTI record arithmetic
 * (:297-327) on the winner it has just recomputed -- no second launch, no round trip of the key-points.
 *   locations, angles, dimensions_out   B*D*3 floats out each (as gpp_pose_*: rows whose orientation is no class in
 *                                       0..3 get zeros and their input dimensions)
 *   kitti                               B*D*4 floats out (alpha, h, Y, r_y) or NULL
 * The results equal gpp_fit_* followed by gpp_pose_* / gpp_kitti_* bit for bit (same device functions). */
int gpp_fit_pose_host(gpp_handle *h, const float *boxes, const float *dimensions, const int32_t *orientations,
                      const float *P_inv, int B, int D, float *keypoints, float *keyplanes, float *residuals,
                      int64_t *best_index, float *locations, float *angles, float *dimensions_out, float *kitti,
                      int mode);
int gpp_fit_pose_device(gpp_handle *h, const float *boxes, const float *dimensions, const int32_t *orientations,
                        const float *P_inv, int B, int D, float *keypoints, float *keyplanes, float *residuals,
                        int64_t *best_index, float *locations, float *angles, float *dimensions_out, float *kitti,
                        int mode, void *stream);

/* FP64 verify mode: same inputs (float32, promoted exactly), search and outputs in double. */
int gpp_fit_host_f64(gpp_handle *h, const float *boxes, const float *dimensions, const int32_t *orientations,
                     const float *P_inv, int B, int D, double *keypoints, double *keyplanes,
                     double *residuals, int64_t *best_index);
int gpp_fit_device_f64(gpp_handle *h, const float *boxes, const float *dimensions,
                       const int32_t *orientations, const float *P_inv, int B, int D, double *keypoints,
                       double *keyplanes, double *residuals, int64_t *best_index, void *stream);

/* 6-DoF pose recovery from the four selected key-points -- the inline loop of run_network.py:137-247
 * (reachable branches only: orientation 1,2 use X_m,X_r,X_t; 0,3 use X_l,X_m,X_t).
 *   keypoints    n*12 floats, dimensions n*3 floats (h,w,l), orientations n int32
 *   locations    n*3 floats out, angles n*3 floats out (Rodrigues vector of [x_dir y_dir z_dir]),
 *   dimensions_out n*3 floats out (h and l overwritten by the measured key-point distances)
 * Rows whose orientation is outside 0..3 are left untouched, like the reference's np.empty_like rows. */
int gpp_pose_host(gpp_handle *h, const float *keypoints, const float *dimensions, const int32_t *orientations,
                  long n, float *locations, float *angles, float *dimensions_out);
int gpp_pose_device(gpp_handle *h, const float *keypoints, const float *dimensions,
                    const int32_t *orientations, long n, float *locations, float *angles,
                    float *dimensions_out, void *stream);

/* KITTI record of posed detections -- the per-detection arithmetic of the reference's KITTI writer
 * (run_network.py:295-330): out[i] = (alpha, h, Y, r_y) where R = Rodrigues(angles[i]), Y / h are the bottom
 * and the height of the rotated 8-corner box, r_y = angles[i][1] wrapped to [-pi, pi) and alpha the observation
 * angle.  locations / angles / dimensions: n*3 floats (the outputs of gpp_pose_*), out: n*4 floats. */
int gpp_kitti_host(gpp_handle *h, const float *locations, const float *angles, const float *dimensions, long n,
                   float *out);
int gpp_kitti_device(gpp_handle *h, const float *locations, const float *angles, const float *dimensions, long n,
                     float *out, void *stream);

/* ---- the two steps before polling in the reference's inference graph ----
 *
 * Decode: RegressBoxes + RegressDims (layers/_misc.py:132-140, :185-186; backend/common.py:23-84).
 *   anchors A*4 (x1,y1,x2,y2), regression B*A*12, classification B*A*8 (sets the sign of the middle / top key-point
 *   x offsets), regression_dim B*A*3  ->  boxes B*A*12, dimensions B*A*3.
 *   box_mean_std: 24 floats (12 means then 12 stds) or NULL for the layer defaults; dim_mean_std: 6 floats or NULL. */
int gpp_decode_host(gpp_handle *h, const float *anchors, const float *regression, const float *classification,
                    const float *regression_dim, int B, int A, const float *box_mean_std, const float *dim_mean_std,
                    float *boxes, float *dimensions);
int gpp_decode_device(gpp_handle *h, const float *anchors, const float *regression, const float *classification,
                      const float *regression_dim, int B, int A, const float *box_mean_std, const float *dim_mean_std,
                      float *boxes, float *dimensions, void *stream);

/* FilterDetections (layers/filter_detections.py:18-189) for the configuration the model is built with
 * (models/retinanet.py:415): one class, class_specific_filter, orientation taken as the arg-max, NMS on.
 * Per image: score = max over the 8 classification values, orientation = arg-max over the 4 (halves merged by max),
 * keep score > score_threshold, greedy NMS (IoU > nms_threshold suppresses; descending score, ties by lower anchor
 * index) up to max_detections (<= 128), rows sorted by score, padded with -1.
 *   out_boxes B*max*12, out_dimensions B*max*3, out_scores B*max, out_labels / out_orientations B*max int32. */
int gpp_filter_host(gpp_handle *h, const float *boxes, const float *dimensions, const float *classification, int B,
                    int A, float score_threshold, float nms_threshold, int max_detections, float *out_boxes,
                    float *out_dimensions, float *out_scores, int32_t *out_labels, int32_t *out_orientations);
int gpp_filter_device(gpp_handle *h, const float *boxes, const float *dimensions, const float *classification, int B,
                      int A, float score_threshold, float nms_threshold, int max_detections, float *out_boxes,
                      float *out_dimensions, float *out_scores, int32_t *out_labels, int32_t *out_orientations,
                      void *stream);

/* Measurement helpers (bench.py): time of the last gpp_fit_* polling kernel(s) in milliseconds, measured
 * with CUDA events on the launching stream (valid after the stream has been synchronised); number of
 * kernels launched by this handle so far. */
int gpp_last_kernel_ms(gpp_handle *h, float *ms);
int64_t gpp_launch_count(const gpp_handle *h);

#ifdef __cplusplus
}
#endif
#endif /* GPP_H_ */
