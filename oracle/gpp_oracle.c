/* ORACLE (test infrastructure, NOT product code) -- plain-C restatement of the ground-plane polling graph,
 * /root/reference/keras_retinanet_3D/layers/fit_road_planes.py:49-139, bit-identical to the numpy oracle
 * (oracle/fit_road_planes_ref.py; checked in tests/test_oracle_c.py) and fast enough to check the CUDA
 * path at 10^7..10^9 hypotheses.  Parity status: pinned at the graph level against the reference file run
 * over numpy stand-ins (tests/golden/), unpinned at the ulp level of TF's own kernels -- see the numpy
 * oracle's header.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.
 *
 * Build (oracle/Makefile): gcc -O2 -ffp-contract=off -fno-fast-math -pthread -shared -fPIC
 */
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <unistd.h>

#define THRESH 0.7

int gpp_oracle_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

#define REAL float
#define SQRT sqrtf
#define FABS fabsf
#define HIGHEST FLT_MAX
#define FN(name) name##_f32
#include "gpp_oracle_body.inc"
#undef REAL
#undef SQRT
#undef FABS
#undef HIGHEST
#undef FN

#define REAL double
#define SQRT sqrt
#define FABS fabs
#define HIGHEST DBL_MAX
#define FN(name) name##_f64
#include "gpp_oracle_body.inc"
