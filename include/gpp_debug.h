/*
 * libgpp -- measurement, tuning and test hooks.  NOT part of the drop-in boundary (include/gpp.h): nothing a caller of
 * fit_road_planes needs lives here.  bench.py uses gpp_microbench for the roofline denominator; the GPU tests use
 * the other hooks to force kernel variants / schedules and to compare per-hypothesis scores with the oracle.
 */
#ifndef GPP_DEBUG_H_
#define GPP_DEBUG_H_

#include "gpp.h"

#ifdef __cplusplus
extern "C" {
#endif

/* FP32 CUDA-core pipe microbenchmarks on the handle's device (the roofline denominator of this path):
 * `kind` 0 = FFMA (3 register operands), 1 = packed FFMA2 (fma.rn.f32x2), 2 = FMUL+FADD uncontracted,
 * 3 = MUFU.RCP, 4 = MUFU.RSQ, 5 = FFMA with one ALU-pipe FMNMX per FFMA (FFMAs counted),
 * 6 = sqrt.approx, 7 = packed FMUL2, 8 = FFMA with one MUFU.RCP per 4 FFMA (FFMAs counted), 9 = packed FADD2,
 * 10 = independent FMUL2 / FADD2 streams 1:1, 11 = independent FFMA2 / FADD2 streams 1:1, 12 = FFMA2 with three
 * distinct 64-bit register operands, 13 = FFMA2 with a broadcast 32-bit operand (packed kinds count two
 * results per instruction).
 * Returns operations per second (an FMA counts as ONE operation; x2 for FLOP), the duration of the best
 * repetition, and operations per SM clock (from clock64 inside the kernel).  Any out pointer may be NULL. */
int gpp_microbench(gpp_handle *h, int kind, double *ops_per_s, float *ms, double *ops_per_clk_sm);

/* Test hook: the per-hypothesis scores of ONE detection against the resident database, computed by the same
 * device functions the search loops call -- `which` 0 = EXACT arithmetic, 1 = FAST (general path), 2 = FAST
 * (all-six-votes path, merged reciprocal), 3 = stage 1 of the VERIFIED all-six path (resid = sum of the three
 * bottom-face residuals, margin = the bound its early exit relies on; votes / zneg are zero).  votes / zneg: N
 * int32, resid: N floats (host memory); margin: N floats or NULL -- the VERIFIED mode's bound on |fast - exact| of
 * the residual sum (0 for which = 0). */
int gpp_debug_scores(gpp_handle *h, const float *box12, const float *dims3, int orientation, const float *pinv12,
                     int which, int32_t *votes, float *resid, int32_t *zneg, float *margin);

/* Schedule of the polling kernel (csrc/gpp_poll3.cuh): `n_seg` plane segments per detection (0 = automatic, 1..32)
 * and, for the FAST / VERIFIED modes, `resident_rows` rows of 64 planes kept in shared memory (-1 = automatic, 0 =
 * stream everything from L2).  Tests force the segmented and the streamed paths. */
int gpp_debug_set_schedule(gpp_handle *h, int n_seg, int resident_rows);

/* The order in which the FAST / VERIFIED scans visit a database (csrc/gpp_order.cu): order[position] = plane index,
 * a permutation of 0 .. n_planes-1 (the identity below 32 rows of 64 planes).  `planes`: n_planes x 4 float32,
 * row-major, as fed.  Host code only -- needs no device. */
int gpp_debug_scan_order(const float *planes, int n_planes, int32_t *order);

/* Runtime audit of the VERIFIED mode: with `every` = n > 0, each VERIFIED call re-polls every n-th detection in the
 * EXACT arithmetic on the same stream and counts the rows whose winning index differs (0 = off, the default; the
 * environment variable GPP_AUDIT=n sets it when a handle is created).  gpp_audit_counts returns the totals since
 * the handle was created (it synchronises the device). */
int gpp_audit_set(gpp_handle *h, int every);
int gpp_audit_counts(gpp_handle *h, int64_t *checked, int64_t *mismatches);

#ifdef __cplusplus
}
#endif
#endif /* GPP_DEBUG_H_ */
