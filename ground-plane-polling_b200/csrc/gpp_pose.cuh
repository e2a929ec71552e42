// 6-DoF pose recovery from the four selected 3-D key-points and the KITTI record of the posed box -- the two steps right
// after polling in the reference's driver loop (/root/reference/keras_retinanet_3D/bin/run_network.py:137-247 and
// :297-327).  Device functions shared by the stand-alone kernels (gpp_pose.cu: gpp_pose_*, gpp_kitti_*) and by the fused
// epilogue of the polling kernel (gpp_poll3.cuh), so that both give the same bits.
// Only the branches the reference loop can reach are implemented (`outlier` is 2 for orientation 0/3 and 0 for 1/2,
// :147-150): orientation 1 -> :167-177, 2 -> :178-188, 0 -> :204-214, 3 -> :237-247.  float32 where the reference
// computes in float32 numpy, double for the rotation-matrix -> Rodrigues-vector conversion (cv2.Rodrigues works in
// double: closest rotation by SVD, then axis * angle).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gpp {

// Orthogonal polar factor U*Vt of a non-singular 3x3 matrix by scaled Newton iteration
// Q <- (g*Q + Q^-T / g) / 2 -- equals the SVD projection cv2.Rodrigues applies before reading the axis.
static __device__ __noinline__ void polar_rotation(double Q[9]) {
#pragma unroll 1
    for (int iter = 0; iter < 24; ++iter) {
        double C[9];   // cofactor matrix = det * Q^-T
        C[0] = Q[4] * Q[8] - Q[5] * Q[7]; C[1] = Q[5] * Q[6] - Q[3] * Q[8]; C[2] = Q[3] * Q[7] - Q[4] * Q[6];
        C[3] = Q[2] * Q[7] - Q[1] * Q[8]; C[4] = Q[0] * Q[8] - Q[2] * Q[6]; C[5] = Q[1] * Q[6] - Q[0] * Q[7];
        C[6] = Q[1] * Q[5] - Q[2] * Q[4]; C[7] = Q[2] * Q[3] - Q[0] * Q[5]; C[8] = Q[0] * Q[4] - Q[1] * Q[3];
        const double det = Q[0] * C[0] + Q[1] * C[1] + Q[2] * C[2];
        if (!(fabs(det) > 0.0) || !isfinite(det)) return;
        double nq = 0.0, nc = 0.0;
#pragma unroll
        for (int i = 0; i < 9; ++i) { C[i] /= det; nq += Q[i] * Q[i]; nc += C[i] * C[i]; }
        const double g = sqrt(sqrt(nc / nq));
        double delta = 0.0;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const double v = 0.5 * (g * Q[i] + C[i] / g);
            delta += (v - Q[i]) * (v - Q[i]);
            Q[i] = v;
        }
        if (delta < 1e-30) return;
    }
}

// cv2.Rodrigues, matrix -> vector branch.  R row-major.
static __device__ __noinline__ void rodrigues_vec(const double Rin[9], double out[3]) {
    double R[9];
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 9; ++i) { R[i] = Rin[i]; ok = ok && (R[i] > -100.0) && (R[i] < 100.0); }
    if (!ok) { out[0] = out[1] = out[2] = 0.0; return; }         // checkRange(-100, 100) failure -> zeros
    polar_rotation(R);
    double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
    const double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = (R[0] + R[4] + R[8] - 1.0) * 0.5;
    c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
    double theta = acos(c);
    if (s < 1e-5) {
        if (c > 0) {
            rx = ry = rz = 0.0;
        } else {
            double t;
            t = (R[0] + 1.0) * 0.5; rx = sqrt(fmax(t, 0.0));
            t = (R[4] + 1.0) * 0.5; ry = sqrt(fmax(t, 0.0)) * (R[1] < 0 ? -1.0 : 1.0);
            t = (R[8] + 1.0) * 0.5; rz = sqrt(fmax(t, 0.0)) * (R[2] < 0 ? -1.0 : 1.0);
            if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && ((R[5] > 0) != (ry * rz > 0))) rz = -rz;
            theta /= sqrt(rx * rx + ry * ry + rz * rz);
            rx *= theta; ry *= theta; rz *= theta;
        }
    } else {
        const double vth = theta / (2.0 * s);
        rx *= vth; ry *= vth; rz *= vth;
    }
    out[0] = rx; out[1] = ry; out[2] = rz;
}

// One detection of the pose loop.  kp = (X_l, X_m, X_r, X_t), w = the network's width, o = orientation class in 0..3.
// out9 = location (3), Rodrigues vector (3), dimensions (h measured, w kept, l measured).
static __device__ __noinline__ void pose_from_keypoints(const float kp[12], float w, int o, float out9[9]) {
    const float *Xl = kp, *Xm = kp + 3, *Xr = kp + 6, *Xt = kp + 9;
    const bool use_r = (o == 1 || o == 2);                        // outlier == 0 -> (X_m, X_r, X_t)
    const float *Xe = use_r ? Xr : Xl;
    float tm[3], em[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { tm[k] = __fsub_rn(Xt[k], Xm[k]); em[k] = __fsub_rn(Xe[k], Xm[k]); }
    const float h = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(tm[0], tm[0]), __fmul_rn(tm[1], tm[1])), __fmul_rn(tm[2], tm[2])));
    const float l = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(em[0], em[0]), __fmul_rn(em[1], em[1])), __fmul_rn(em[2], em[2])));
    // x_dir sign: o=1 (X_m-X_r)/l, o=2 (X_r-X_m)/l, o=0 (X_m-X_l)/l, o=3 (X_l-X_m)/l
    const bool x_from_m = (o == 1 || o == 0);
    float x[3], y[3], z[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float dx = x_from_m ? __fsub_rn(Xm[k], Xe[k]) : __fsub_rn(Xe[k], Xm[k]);
        x[k] = __fdiv_rn(dx, l);
        y[k] = __fdiv_rn(__fsub_rn(Xm[k], Xt[k]), h);
    }
    z[0] = __fsub_rn(__fmul_rn(x[1], y[2]), __fmul_rn(x[2], y[1]));
    z[1] = __fsub_rn(__fmul_rn(x[2], y[0]), __fmul_rn(x[0], y[2]));
    z[2] = __fsub_rn(__fmul_rn(x[0], y[1]), __fmul_rn(x[1], y[0]));
    // location: o=1 -, o=2 +, o=0 +, o=3 -
    const bool plus = (o == 2 || o == 0);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float mid = __fdiv_rn(__fadd_rn(Xm[k], Xe[k]), 2.0f);
        const float off = __fdiv_rn(__fmul_rn(z[k], w), 2.0f);
        out9[k] = plus ? __fadd_rn(mid, off) : __fsub_rn(mid, off);
    }
    const double R[9] = {(double)x[0], (double)y[0], (double)z[0], (double)x[1], (double)y[1], (double)z[1],
                         (double)x[2], (double)y[2], (double)z[2]};   // columns x_dir, y_dir, z_dir
    double rv[3];
    rodrigues_vec(R, rv);
#pragma unroll
    for (int k = 0; k < 3; ++k) out9[3 + k] = (float)rv[k];
    out9[6] = h;
    out9[7] = w;
    out9[8] = l;
}

static __device__ __forceinline__ double wrap_pi(double a) {
    const double two_pi = 6.283185307179586476925286766559;
    a = fmod(a, two_pi);
    if (a < 0) a += two_pi;                                    // python's % returns a value in [0, 2 pi)
    if (a >= 3.14159265358979323846) a -= two_pi;
    return a;
}

// KITTI record of one posed detection -- the per-detection arithmetic of the reference's KITTI writer
// (bin/run_network.py:297-327): R = Rodrigues(angles); the 8 box corners (label_prep/computeBox3D.m convention)
// rotated and translated; Y = max corner y, h = Y - min corner y; r_y = angles[1] wrapped to [-pi, pi);
// alpha = r_y + atan2(z, x) + 1.5 pi wrapped the same way.  pose9 as written by pose_from_keypoints; out = (alpha, h, Y, r_y).
static __device__ __noinline__ void kitti_record(const float pose9[9], float out[4]) {
    const float *loc = pose9, *ang = pose9 + 3, *dims = pose9 + 6;
    const double rx = ang[0], ry = ang[1], rz = ang[2];
    const double theta = sqrt(rx * rx + ry * ry + rz * rz);
    // second row of the rotation matrix (only the y coordinates of the corners are needed)
    double R10, R11, R12;
    if (theta < 2.220446049250313e-16) {
        R10 = 0.0; R11 = 1.0; R12 = 0.0;
    } else {
        const double c = cos(theta), s = sin(theta), c1 = 1.0 - c, it = 1.0 / theta;
        const double ux = rx * it, uy = ry * it, uz = rz * it;
        R10 = c1 * uy * ux + s * uz;
        R11 = c + c1 * uy * uy;
        R12 = c1 * uy * uz - s * ux;
    }
    const double h = dims[0], w = dims[1], l = dims[2];
    const double xs[8] = {l / 2, l / 2, -l / 2, -l / 2, l / 2, l / 2, -l / 2, -l / 2};
    const double ys[8] = {0, 0, 0, 0, -h, -h, -h, -h};
    const double zs[8] = {w / 2, -w / 2, -w / 2, w / 2, w / 2, -w / 2, -w / 2, w / 2};
    const double ly = loc[1];
    double ymax = -1e300, ymin = 1e300;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const double y = R10 * xs[k] + R11 * ys[k] + R12 * zs[k] + ly;
        ymax = fmax(ymax, y);
        ymin = fmin(ymin, y);
    }
    const double r_y = wrap_pi(ry);
    const double alpha = wrap_pi(r_y + atan2((double)loc[2], (double)loc[0]) + 4.71238898038468985769);
    out[0] = (float)alpha;
    out[1] = (float)(ymax - ymin);
    out[2] = (float)ymax;
    out[3] = (float)r_y;
}

}  // namespace gpp
