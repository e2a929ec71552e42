// Packed-pair polling kernel (fp32 modes).
//
// Blackwell's FP32 pipe executes packed two-wide instructions (PTX add/sub/mul/fma .f32x2 -> SASS FADD2 /
// FMUL2 / FFMA2): a packed add or multiply delivers two IEEE-rounded results per issue slot and a packed FMA
// two FMAs per slot (measured: profiles/r01a_probe_microbench_sweep.json).  The polling kernel is issue-slot
// bound, so every lane evaluates TWO planes per iteration (the pair (2p, 2p+1) of a pair-interleaved copy of
// the database) against its warp's detection; per-detection constants are plain 32-bit registers that the
// packed instructions broadcast (SASS operand form `R.F32`), abs/neg fold into operand modifiers.
//
// Same algorithm as gpp_poll.cuh (fit_road_planes.py:86-119), plus one specialisation: as soon as the warp
// has seen a plane with all six votes, max-votes is known to be 6 for good, and "votes == 6" becomes
// "max_k |r_k| <= 0.7" (NaN-ignoring max keeps the reference's NaN-votes rule), which replaces the vote
// counting by three FMNMX3.
#pragma once
#include "gpp_poll.cuh"

#ifndef GPP_M6_UNROLL
#define GPP_M6_UNROLL 1
#endif
#define GPP_PRAGMA_(x) _Pragma(#x)
#define GPP_UNROLL(n) GPP_PRAGMA_(unroll n)

namespace gpp {

typedef unsigned long long u64;

#ifdef GPP_EXP_SCALAR   /* timing experiment only: the same code on scalar instructions */
struct f2 {
    float l, h;
};
__device__ __forceinline__ f2 pk(float lo, float hi) { return f2{lo, hi}; }
__device__ __forceinline__ f2 bc(float x) { return f2{x, x}; }
__device__ __forceinline__ float lo(f2 a) { return a.l; }
__device__ __forceinline__ float hi(f2 a) { return a.h; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return f2{a.l * b.l, a.h * b.h}; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return f2{a.l + b.l, a.h + b.h}; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { return f2{a.l - b.l, a.h - b.h}; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return f2{fmaf(a.l, b.l, c.l), fmaf(a.h, b.h, c.h)}; }
__device__ __forceinline__ f2 from_u64(u64 v) {
    f2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.l), "=f"(r.h) : "l"(v));
    return r;
}
#else
// ------------------------------------------------------------------ packed f32x2 primitives
struct f2 {
    u64 v;
};
__device__ __forceinline__ f2 pk(float lo, float hi) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f2 bc(float x) { return pk(x, x); }      // broadcast operand (folds to R.F32)
__device__ __forceinline__ float lo(f2 a) {
    float l, h;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(l), "=f"(h) : "l"(a.v));
    return l;
}
__device__ __forceinline__ float hi(f2 a) {
    float l, h;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(l), "=f"(h) : "l"(a.v));
    return h;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
    f2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
    f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b) {
    f2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
    f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}
__device__ __forceinline__ f2 from_u64(u64 v) { return f2{v}; }
#endif
__device__ __forceinline__ f2 neg2(f2 a) { return pk(-lo(a), -hi(a)); }
__device__ __forceinline__ f2 abs2(f2 a) { return pk(fabsf(lo(a)), fabsf(hi(a))); }

// per-half scalar ops
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float max3f(float a, float b, float c) {      // NaN-ignoring (IEEE maxNum)
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// ------------------------------------------------------------------ packed policies
struct PackFast {
    static constexpr bool kExact = false;
    typedef FastF32 Scalar;
    static __device__ __forceinline__ f2 madd(f2 a, f2 b, f2 c) { return fma2(a, b, c); }
    static __device__ __forceinline__ f2 rcp(f2 a) { return pk(rcp_approx(lo(a)), rcp_approx(hi(a))); }
    static __device__ __forceinline__ f2 sqrt(f2 a) { return pk(sqrt_approx(lo(a)), sqrt_approx(hi(a))); }
};

// Error margin of one hypothesis (derivation: DESIGN.md section 4.1.1).  Both the fast and the exact fp32 evaluation
// deviate from the real-valued score mainly through the rounding of t_k = n.d_k (absolute <= ~7 u |d_k| for the two
// together) divided by |t_k|: a point at distance |X_k| = |d_k| |d| / |t_k| moves by <= ~12 u |d| (|d_k| i_k)^2 with
// i_k = 1/|t_k| (quadratic in depth, invariant under a rescaling of the rays); the six distances weigh the four points
// 3 : 3 : 3 : 3, X_t inherits the error of X_m plus that of q ~ 1/(perp.n).  Worst case with every error aligned:
// 144 u w + 30 u w T/|perp.n| + 24 u |q| T/|perp.n| (w = |d| max_k |d_k|^2 i_max^2), against the margin
// K u w (1 + T/|perp.n|) + 4 K u |q| T/|perp.n| with K = 128; measured on adversarial inputs (scripts/
// gpu_margin_pressure.py): |fast - exact| <= 0.11 margin.  Two regions have no first-order bound and are handled
// separately: t_k at its rounding noise (the margin then exceeds the residual sum itself, so the plane survives) and
// perp.n at its rounding noise (explicit test in eval_top).
#ifndef GPP_MARGIN_K
#define GPP_MARGIN_K 128.0f
#endif
constexpr float kMarginScale = GPP_MARGIN_K * 5.9604645e-8f;   // K * 2^-24

// per-detection constants of the packed kernels (warp-uniform scalars)
struct DetConst {
    float td[6];       // the six target lengths (exact prologue values)
    float T;           // |d_t|^2  (fast formulation only)
    float G;           // d_t . d_m (fast formulation only)
    float ms, msT;     // VERIFIED: K 2^-24 max_k |d_k|^2 and the same times T (margin scales, ray-scale invariant)
    float msq;         // VERIFIED: 4 K 2^-24 T, scale of the conditioning term of q (see eval_top)
    float fl[2], fm[2], fr[2], ft[2];   // fast formulation: rays rescaled to z = 1, (x, y) only; T, G, ms refer to these
    float mc;          // VERIFIED: 2^-20 * (sum of the six target lengths), see finalize_margin
};

// fills the fast-formulation constants of a detection from its exact rays
__device__ __forceinline__ void fast_constants(DetConst &D, const Detection<ExactF32> &det) {
    const float zl = 1.0f / det.dl[2], zm = 1.0f / det.dm[2], zr = 1.0f / det.dr[2], zt = 1.0f / det.dt[2];
    D.fl[0] = det.dl[0] * zl; D.fl[1] = det.dl[1] * zl;
    D.fm[0] = det.dm[0] * zm; D.fm[1] = det.dm[1] * zm;
    D.fr[0] = det.dr[0] * zr; D.fr[1] = det.dr[1] * zr;
    D.ft[0] = det.dt[0] * zt; D.ft[1] = det.dt[1] * zt;
    D.T = fmaf(D.ft[0], D.ft[0], fmaf(D.ft[1], D.ft[1], 1.0f));
    D.G = fmaf(D.ft[0], D.fm[0], fmaf(D.ft[1], D.fm[1], 1.0f));
    const float d0 = fmaf(D.fl[0], D.fl[0], fmaf(D.fl[1], D.fl[1], 1.0f));
    const float d1 = fmaf(D.fm[0], D.fm[0], fmaf(D.fm[1], D.fm[1], 1.0f));
    const float d2 = fmaf(D.fr[0], D.fr[0], fmaf(D.fr[1], D.fr[1], 1.0f));
    D.ms = kMarginScale * fmaxf(fmaxf(d0, d1), fmaxf(d2, D.T));
    D.msT = D.ms * D.T;
    D.msq = 4.0f * kMarginScale * D.T;
    D.mc = 9.5367431640625e-07f * (fabsf(D.td[0]) + fabsf(D.td[1]) + fabsf(D.td[2]) + fabsf(D.td[3]) + fabsf(D.td[4]) +
                                  fabsf(D.td[5]));
}

struct PairResult;


struct PairResult {
    f2 r[6];           // signed residuals dist_k - target_k (abs is applied by the consumers)
    f2 zc;             // z_dir_check
    f2 m;              // VERIFIED mode: bound on |fast - exact| of the residual sum (see eval_pair_fast)
    f2 za0, za2, zb0, zb2;   // lazy z_dir_check: the x/z components of X_l - X_m and X_r - X_m
    __device__ __forceinline__ void finish_zc() { zc = fma2(za2, zb0, neg2(mul2(za0, zb2))); }
};



__device__ __forceinline__ f2 dot3p(f2 a0, f2 a1, f2 a2, float b0, float b1, float b2, bool exact) {
    // (a0*b0 + a1*b1) + a2*b2, exact: three multiplies and two adds; fast: multiply + two FMAs
    if (exact) return add2(add2(mul2(a0, bc(b0)), mul2(a1, bc(b1))), mul2(a2, bc(b2)));
    return fma2(a2, bc(b2), fma2(a1, bc(b1), mul2(a0, bc(b0))));
}

// ---- FAST: FMA contraction, MUFU reciprocal / square root, and algebra that is exact in real arithmetic.
// All packed FP32 instructions (FFMA2 / FMUL2 / FADD2) retire 128 results/clk/SM like their scalar forms and the
// loop is register-file-bandwidth bound (measured), so what counts is the number of instructions and operands
// per hypothesis.
//  * The rays are rescaled per detection to z = 1 (X_k = d_k |d / n.d_k| does not depend on the length of d_k),
//    so every n.d_k is two FMAs and the z coordinate of X_k is the scale s_k itself.
//  * With a unit normal n every intersection point lies on the (+-) plane, n.X_k = s_k t_k = |d| sign(t_k), which
//    removes most of calc_X_t:
//      perp = d_t x (n x d_t) = n T - d_t u           (T = |d_t|^2, u = n.d_t)
//      perp.n   = T - u^2
//      perp.X_m = s_m (T t_m - u G)                   (G = d_t.d_m)
//      X_t = X_m - q n,  |X_m - X_t| = |q|
//      |X_l - X_t|^2 = |a|^2 + q (2 a.n + q),  a = X_l - X_m,  a.n = |d| (sign(t_l) - sign(t_m))   (same for X_r)
//    and a.n = 0 unless the plane separates the rays, which is tested once per warp.
// About 55 FMA-pipe results, 8-9 MUFU and ~6 ALU-pipe instructions per hypothesis (direct formulation: 92 / 10 / 17).
// The evaluation comes in two stages so that the VERIFIED all-six phase can stop after the first one:
//   eval_bottom: the three points on the plane and the squared lengths of the bottom-face edges / diagonal
//                (residuals 1, 2, 3) -- nothing that involves X_t;
//   eval_top   : X_t (u, perp.n, q) and the squared lengths of the two slanted edges (residuals 0, 4, 5), the
//                error margin and z_dir_check.
// eval_pair_fast = both stages + the square roots; every user of the fast arithmetic goes through these two.
struct Bottom {
    f2 t0, t1, t2;     // n . d_k for the rescaled rays (l, m, r)
    f2 ad;             // |d|
    f2 s1;             // depth scale of X_m
    f2 w;              // margin weight |d| / t_min^2 (kMargin only)
    f2 a[3], b[3];     // X_l - X_m, X_r - X_m
    f2 na, nb, nc;     // |X_l - X_m|^2, |X_r - X_m|^2, |X_l - X_r|^2
};

// eval_bottom comes in two steps so that a caller can let go of the plane registers in between (the resident kernel
// fetches the next row into them): eval_dots is the only part that reads the plane itself.
__device__ __forceinline__ void eval_dots(const DetConst &D, f2 n0, f2 n1, f2 n2, f2 d4, Bottom &g) {
    g.t0 = fma2(n0, bc(D.fl[0]), fma2(n1, bc(D.fl[1]), n2));
    g.t1 = fma2(n0, bc(D.fm[0]), fma2(n1, bc(D.fm[1]), n2));
    g.t2 = fma2(n0, bc(D.fr[0]), fma2(n1, bc(D.fr[1]), n2));
    g.ad = abs2(d4);
}

template <bool kMergedRcp, bool kMargin>
__device__ __forceinline__ void eval_bottom_rest(const DetConst &D, Bottom &g) {
    f2 i0, i1;
    if (kMergedRcp) {
        // 1/t_l and 1/t_m from one MUFU.RCP.  Only used once max-votes is known to be 6, where a degenerate
        // (inf/NaN) hypothesis can neither win nor change max-votes.
        const f2 inv = PackFast::rcp(mul2(g.t0, g.t1));
        i0 = mul2(inv, g.t1);
        i1 = mul2(inv, g.t0);
    } else {
        i0 = PackFast::rcp(g.t0);
        i1 = PackFast::rcp(g.t1);
    }
    const f2 s0 = mul2(g.ad, abs2(i0));
    g.s1 = mul2(g.ad, abs2(i1));
    const f2 i2 = PackFast::rcp(abs2(g.t2));
    const f2 s2 = mul2(g.ad, i2);
    if (kMargin) {
        const f2 imax = pk(max3f(fabsf(lo(i0)), fabsf(lo(i1)), lo(i2)), max3f(fabsf(hi(i0)), fabsf(hi(i1)), hi(i2)));
        g.w = mul2(mul2(imax, imax), g.ad);                              // |d| / t_min^2
    }
    const f2 nx = neg2(mul2(bc(D.fm[0]), g.s1)), ny = neg2(mul2(bc(D.fm[1]), g.s1));      // -X_m (x, y)
    g.a[0] = fma2(bc(D.fl[0]), s0, nx); g.a[1] = fma2(bc(D.fl[1]), s0, ny); g.a[2] = sub2(s0, g.s1);   // X_l - X_m
    g.b[0] = fma2(bc(D.fr[0]), s2, nx); g.b[1] = fma2(bc(D.fr[1]), s2, ny); g.b[2] = sub2(s2, g.s1);   // X_r - X_m
#define GPP_SQN(v) fma2(v[2], v[2], fma2(v[1], v[1], mul2(v[0], v[0])))
    g.na = GPP_SQN(g.a);
    g.nb = GPP_SQN(g.b);
#undef GPP_SQN
    // |X_l - X_r|^2 from the difference itself (|a|^2 + |b|^2 - 2 a.b would cancel when X_l is close to X_r)
    const f2 c0 = sub2(g.a[0], g.b[0]), c1 = sub2(g.a[1], g.b[1]), c2 = sub2(g.a[2], g.b[2]);
    g.nc = fma2(c2, c2, fma2(c1, c1, mul2(c0, c0)));
}

template <bool kMergedRcp, bool kMargin>
__device__ __forceinline__ void eval_bottom(const DetConst &D, f2 n0, f2 n1, f2 n2, f2 d4, Bottom &g) {
    eval_dots(D, n0, n1, n2, d4, g);
    eval_bottom_rest<kMergedRcp, kMargin>(D, g);
}

// kMargin: 0 = none, 1 = out.m = geometric margin + mc, 2 = geometric margin only (the caller accounts for mc).
// Sets out.r[0] (signed residual of the height) and leaves the SQUARED lengths of the slanted edges in ne / nf.
template <int kMargin, bool kLazyZ>
__device__ __forceinline__ void eval_top(const DetConst &D, f2 n0, f2 n1, f2 n2, f2 d4, const Bottom &g,
                                         PairResult &out, f2 &ne, f2 &nf) {
    const f2 u = fma2(n0, bc(D.ft[0]), fma2(n1, bc(D.ft[1]), n2));
    const f2 den = fma2(neg2(u), u, bc(D.T));
    const f2 iden = PackFast::rcp(den);
    if (kMargin) {
        const f2 f = fma2(abs2(iden), bc(D.msT), bc(D.ms));             // K u |d_k|^2 (1 + T/|perp.n|)
        out.m = kMargin == 1 ? fma2(g.w, f, bc(D.mc)) : mul2(g.w, f);
    }
    if (kLazyZ) {           // z_dir_check is only formed for the few pairs that get that far (see the callers)
        out.za0 = g.a[0]; out.za2 = g.a[2]; out.zb0 = g.b[0]; out.zb2 = g.b[2];
    } else {
        out.zc = fma2(g.a[2], g.b[0], neg2(mul2(g.a[0], g.b[2])));
    }
    const f2 q = mul2(mul2(g.s1, fma2(u, bc(-D.G), mul2(g.t1, bc(D.T)))), iden);
    if (kMargin) {
        // perp.n = T - u^2 cancels when the top ray is nearly parallel to the plane normal: its rounding error
        // (<= 8 ulp of T) is a RELATIVE error rho = 8 u T / |perp.n| of q, and |q| is not bounded by the depth of X_m then
        // (|q| <= |X_m| sqrt(T / |perp.n|)).  Found by the adversarial soak of round 2 (steep planes: the term above
        // was exceeded up to 9-fold by hypotheses with T / |perp.n| > 5000); three residuals carry the error of q.
        // Where perp.n is down at its own rounding noise (rho' = 32 rho >= 1/4: not even its sign is known) the error
        // of q has no first-order bound: the margin is infinite there and the plane is always re-evaluated exactly.
        const f2 rho = mul2(abs2(iden), bc(D.msq));
        out.m = fma2(abs2(q), rho, out.m);
        out.m = pk(lo(rho) < 0.25f ? lo(out.m) : __int_as_float(0x7f800000),
                   hi(rho) < 0.25f ? hi(out.m) : __int_as_float(0x7f800000));
    }
    // does the plane separate the rays (sign(t_l) or sign(t_r) != sign(t_m)) for any pair of this warp?
    const unsigned sd = ((__float_as_uint(lo(g.t0)) ^ __float_as_uint(lo(g.t1))) | (__float_as_uint(lo(g.t2)) ^ __float_as_uint(lo(g.t1))) |
                         (__float_as_uint(hi(g.t0)) ^ __float_as_uint(hi(g.t1))) | (__float_as_uint(hi(g.t2)) ^ __float_as_uint(hi(g.t1)))) >> 31;
    if (__any_sync(0xffffffffu, sd != 0u)) {
        const f2 cs0 = pk(copysignf(lo(d4), lo(g.t0)), copysignf(hi(d4), hi(g.t0)));
        const f2 cs1 = pk(copysignf(lo(d4), lo(g.t1)), copysignf(hi(d4), hi(g.t1)));
        const f2 cs2 = pk(copysignf(lo(d4), lo(g.t2)), copysignf(hi(d4), hi(g.t2)));
        ne = fma2(q, fma2(sub2(cs0, cs1), bc(2.0f), q), g.na);           // |X_l - X_t|^2
        nf = fma2(q, fma2(sub2(cs2, cs1), bc(2.0f), q), g.nb);           // |X_r - X_t|^2
    } else {
        ne = fma2(q, q, g.na);
        nf = fma2(q, q, g.nb);
    }
    out.r[0] = sub2(abs2(q), bc(D.td[0]));
}

template <bool kMergedRcp, int kMargin, bool kLazyZ = false>
__device__ __forceinline__ void eval_pair_fast(const DetConst &D, f2 n0, f2 n1, f2 n2, f2 d4, PairResult &out) {
    Bottom g;
    eval_bottom<kMergedRcp, kMargin != 0>(D, n0, n1, n2, d4, g);
    f2 ne, nf;
    eval_top<kMargin, kLazyZ>(D, n0, n1, n2, d4, g, out, ne, nf);
    out.r[1] = sub2(PackFast::sqrt(g.na), bc(D.td[1]));
    out.r[2] = sub2(PackFast::sqrt(g.nb), bc(D.td[2]));
    out.r[3] = sub2(PackFast::sqrt(g.nc), bc(D.td[3]));
    out.r[4] = sub2(PackFast::sqrt(abs2(ne)), bc(D.td[4]));
    out.r[5] = sub2(PackFast::sqrt(abs2(nf)), bc(D.td[5]));
}
template <bool kSix>
__device__ __forceinline__ void eval_pair(PackFast, const DetConst &D, f2 n0, f2 n1, f2 n2, f2 d4,
                                          PairResult &out) {
    eval_pair_fast<kSix, 0>(D, n0, n1, n2, d4, out);
}

// The geometric margin of eval_pair_fast covers the error of the 3-D points.  The six distances, their
// differences to the targets and the residual sum are additionally rounded at their own magnitude
// (<= R + sum of targets): 16 ulp of that are added.
__device__ __forceinline__ void finalize_margin(PairResult &h, f2 R, const DetConst &D);

// residual sum ((((|r0|+|r1|)+|r2|)+|r3|)+|r4|)+|r5| for both planes of the pair
__device__ __forceinline__ f2 resid_sum(const PairResult &h) {
    f2 s = add2(abs2(h.r[0]), abs2(h.r[1]));
    s = add2(s, abs2(h.r[2]));
    s = add2(s, abs2(h.r[3]));
    s = add2(s, abs2(h.r[4]));
    s = add2(s, abs2(h.r[5]));
    return s;
}
__device__ __forceinline__ void finalize_margin(PairResult &h, f2 R, const DetConst &) {
    h.m = fma2(R, bc(9.5367431640625e-07f), h.m);             // eval_pair_fast already added the mc term
}
__device__ __forceinline__ int votes_of(float r0, float r1, float r2, float r3, float r4, float r5) {
    const float thr = 0.7f;   // where(greater(|r|, thr), 0, 1): NaN is not greater -> a vote (:31)
    return int(!(fabsf(r0) > thr)) + int(!(fabsf(r1) > thr)) + int(!(fabsf(r2) > thr)) +
           int(!(fabsf(r3) > thr)) + int(!(fabsf(r4) > thr)) + int(!(fabsf(r5) > thr));
}
__device__ __forceinline__ float rmax_of(float r0, float r1, float r2, float r3, float r4, float r5) {
    float m = max3f(fabsf(r0), fabsf(r1), fabsf(r2));
    m = max3f(m, fabsf(r3), fabsf(r4));
    return fmaxf(m, fabsf(r5));
}

// state of a lane once max-votes is known to be 6: best = min residual over {all six votes, z-check passes}
struct LaneBest {
    float bestR;
    int bestIdx;
    __device__ __forceinline__ void update6(float rmax, float zc, float R, int j) {
        const bool better = !(rmax > 0.7f) && !(zc < 0.0f) && (R < bestR);
        bestR = better ? R : bestR;
        bestIdx = better ? j : bestIdx;
    }
};

template <class T>
struct PollArgs2 {
    const float *boxes, *dims, *pinv;
    const int32_t *orient;
    const u64 *pairs;            // pair-interleaved normalised DB: per pair {a0,a1,b0,b1,c0,c1,d0,d1}
    const float4 *planes;        // plain normalised DB (N x float4), for the epilogue
    int n_planes;                // N
    int n_pairs_padded;          // multiple of 32 (database padded to 64 planes with copies of the last)
    int dets_per_image;
    long long n_det;
    T *keypoints, *keyplanes, *residuals;
    long long *best;
    // optional work list: process det_list[0 .. *det_count) instead of 0 .. n_det (rows that repeat the
    // previous row of their image -- FilterDetections' -1 padding -- are computed once and copied)
    const long long *det_list;
    const unsigned int *det_count;
    // dynamic scheduling: detection groups are claimed from this device counter (zeroed before the launch), so
    // CTAs that drew cheap detections take more groups instead of idling at the end of the kernel
    unsigned int *group_counter;
};

// kTile planes per smem tile (multiple of 64), one detection per warp.
// exact scalar evaluation of one plane of a pair (VERIFIED mode: general path, re-evaluation, epilogue)
__device__ __forceinline__ void exact_one(const Detection<ExactF32> &de, float n0, float n1, float n2, float d4,
                                          int &V, float &R, bool &zneg) {
    float X[4][3];
    hypothesis<ExactF32>(de, n0, n1, n2, d4, X, V, R, zneg);
}

constexpr int kVerifyQueue = 96;

// development counters (only with -DGPP_STATS; printed by launch_poll_f32): rows in the all-six phase, rows in
// the general phase with Mcur >= 4 / < 4, rows that passed the cheap test (all-six, general), exact
// verifications, flushes, detections
#ifdef GPP_STATS
__device__ unsigned long long g_stats[8];
#define GPP_STAT(i, n) (st_cnt[i] += (n))
#else
#define GPP_STAT(i, n) ((void)0)
#endif

// VERIFIED: exact re-evaluation of one queued plane with the full (max-votes, residual, index) bookkeeping;
// ties break by index explicitly (the queue is not drained in index order)
__device__ __forceinline__ void verify_general(const Detection<ExactF32> &de, const float4 *__restrict__ planes, int j,
                                               LaneState<float> &st) {
    const float4 pl = planes[j];
    int V; float R; bool z;
    exact_one(de, pl.x, pl.y, pl.z, pl.w, V, R, z);
    const bool cand = !z && (R < FLT_MAX);                     // NaN / >= highest can never win
    if (V > st.M) {
        st.M = V;
        st.bestR = cand ? R : FLT_MAX;
        st.bestIdx = cand ? j : 0;
    } else if (V == st.M && cand && (R < st.bestR || (R == st.bestR && j < st.bestIdx))) {
        st.bestR = R;
        st.bestIdx = j;
    }
}
// upper bound of z_dir_check = (a x b).y with a = X_l - X_m, b = X_r - X_m: the positions move by <= m, so the
// cross product moves by <= m (|a| + |b|) (+ m^2); |a| = r[1] + target_1, |b| = r[2] + target_2
__device__ __forceinline__ f2 z_upper(const PairResult &h, const DetConst &D) {
    const f2 len = add2(add2(abs2(add2(h.r[1], bc(D.td[1]))), abs2(add2(h.r[2], bc(D.td[2])))), h.m);
    return fma2(h.m, fma2(len, bc(4.0f), bc(1.0f)), h.zc);
}
__device__ __forceinline__ int loose_votes(const PairResult &h, bool upper) {
    const float thr = 0.7f;   // votes that are possible within the margin: !(|r_k| - m > thr), NaN counts
    int v = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const float rk = upper ? hi(h.r[k]) : lo(h.r[k]);
        const float mk = upper ? hi(h.m) : lo(h.m);
        v += int(!(fabsf(rk) - mk > thr));
    }
    return v;
}

__device__ __forceinline__ int strict_votes(const PairResult &h, bool upper) {
    const float thr = 0.7f;   // votes that are certain within the margin: |r_k| + m <= thr (NaN never counts)
    int v = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const float rk = upper ? hi(h.r[k]) : lo(h.r[k]);
        const float mk = upper ? hi(h.m) : lo(h.m);
        v += int(fabsf(rk) + mk <= thr);
    }
    return v;
}

// kVerified: FAST arithmetic is only a filter -- every hypothesis that could be the arg-min within the error
// margin is re-evaluated in the EXACT arithmetic and all selection state is kept in exact values, so the
// result equals the EXACT mode's (see the header comment of the verified path below).
// kSplit: small-batch variant (one detection per CTA, warp w takes the rows r = w (mod kWarps) of every tile).
// kVMode: 0 = plain FAST search, 1 = VERIFIED.  The verified filter has two warp-uniform phases: while the
// warp's exact max-votes is below 6 it counts the votes that are possible within the margin (general phase);
// once a plane with six exact votes is known it tests max_k |r_k| - m <= 0.7 (all-six-votes phase).
// kFree (VERIFIED, not kSplit): no detection groups and no CTA barrier.  Every warp claims its detections from the
// device counter on its own and starts each one at whatever tile the CTA's ring currently delivers (the database
// is scanned in rotated order, which the explicit index tie-breaks of the exact bookkeeping allow); the ring runs
// as long as any warp needs tiles (`need_until`), warps that ran out of work keep releasing tiles until all are done.
template <class PP, int kWarps, int kTile, int kStages, int kMinBlocks, int kVMode = 0, bool kSplit = false,
          bool kFree = false>
__global__ void __launch_bounds__(kWarps * 32, kMinBlocks) poll2_kernel(const PollArgs2<float> args) {
    static_assert(!kFree || (kVMode != 0 && !kSplit), "kFree is implemented for the VERIFIED batch kernel");
    constexpr bool kVerified = kVMode != 0;
    constexpr int kTilePairs = kTile / 2;
    constexpr int kRowStep = kSplit ? kWarps : 1;
    constexpr uint32_t kPairBytes = 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ulonglong2 *tiles = reinterpret_cast<ulonglong2 *>(smem_raw);             // 2 x ulonglong2 per pair
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem_raw + size_t(kPairBytes) * kStages * kTilePairs);
    uint64_t *empty_bar = full_bar + kStages;
    // VERIFIED: per-warp queue of plane indices that survived the fast filter (at most 31 + 64 entries)
    int *queue = reinterpret_cast<int *>(empty_bar + kStages) + (threadIdx.x >> 5) * kVerifyQueue;
    WarpPartial<float> *partial = reinterpret_cast<WarpPartial<float> *>(
        reinterpret_cast<int *>(empty_bar + kStages) + kWarps * kVerifyQueue);              // [2][kWarps], kSplit only

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int N = args.n_planes;
    const int NP = args.n_pairs_padded;
    const int n_tiles = (NP + kTilePairs - 1) / kTilePairs;
    const long long n_work = args.det_list ? (long long)(*args.det_count) : args.n_det;
    const long long n_groups = kSplit ? n_work : (n_work + kWarps - 1) / kWarps;
    unsigned int *claim = reinterpret_cast<unsigned int *>(partial + 2 * kWarps);       // [2]: groups k, k+1
    float *detx = reinterpret_cast<float *>(claim + 2) + (threadIdx.x >> 5) * 20;       // this warp's exact constants
    // kFree: one past the last tile sequence number any warp has announced it needs / warps that still have work
    static_assert((size_t(32) * kStages * kTilePairs + 2 * kStages * sizeof(uint64_t) + sizeof(int) * kWarps * kVerifyQueue +
                   2 * kWarps * sizeof(WarpPartial<float>) + 2 * sizeof(unsigned int) + sizeof(float) * 20 * kWarps) % 8 == 0,
                  "need_until must be 8-byte aligned");
    unsigned long long *need_until = reinterpret_cast<unsigned long long *>(reinterpret_cast<float *>(claim + 2) + kWarps * 20);
    int *active = reinterpret_cast<int *>(need_until + 1);

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kWarps);
        }
        mbar_fence_init();
        if (kFree) {
            *need_until = 0ull;
            *active = kWarps;
        } else {
            claim[0] = atomicAdd(args.group_counter, 1u);
            claim[1] = atomicAdd(args.group_counter, 1u);
        }
    }
    __syncthreads();

    auto issue = [&](long long it) {
        const int s = int(it % kStages);
        const int t = int(it % n_tiles);
        const int cnt = min(kTilePairs, NP - t * kTilePairs);
        const uint32_t bytes = uint32_t(cnt) * kPairBytes;
        mbar_arrive_expect_tx(&full_bar[s], bytes);
        tma_load_1d(reinterpret_cast<unsigned char *>(tiles) + size_t(s) * kTilePairs * kPairBytes,
                    reinterpret_cast<const unsigned char *>(args.pairs) + size_t(t) * kTilePairs * kPairBytes, bytes,
                    &full_bar[s]);
    };
    // producer state (thread 0 only): tiles issued so far / tiles known to be needed (groups claimed so far)
    long long issued = 0, known_tiles = 0;
    auto pump = [&](long long max_index) {
        while (issued < known_tiles && issued <= max_index) {
            if (issued >= kStages)                                   // the slot's previous tile must be released
                mbar_wait(&empty_bar[issued % kStages], uint32_t(((issued / kStages) - 1) & 1));
            issue(issued);
            ++issued;
        }
    };

    auto pump_free = [&](long long max_index) {                     // kFree: bounded by the announced need instead
        const long long need = (long long)*reinterpret_cast<volatile unsigned long long *>(need_until);
        while (issued < need && issued <= max_index) {
            if (issued >= kStages)
                mbar_wait(&empty_bar[issued % kStages], uint32_t(((issued / kStages) - 1) & 1));
            issue(issued);
            ++issued;
        }
    };

    long long it = 0;
    unsigned int next_claim = 0;
    if (kFree) {
        if (lane == 0) next_claim = atomicAdd(args.group_counter, 1u);
        next_claim = __shfl_sync(0xffffffffu, next_claim, 0);
    }
    for (long long k = 0;; ++k) {
        long long w_id;
        if (kFree) {
            const unsigned int cur = next_claim;
            if ((long long)cur >= n_work) break;                         // warp-uniform: this warp retires
            if (lane == 0) {
                atomicMax(need_until, (unsigned long long)(it + n_tiles));   // announced before anything waits for it
                next_claim = atomicAdd(args.group_counter, 1u);              // claimed one detection ahead
            }
            next_claim = __shfl_sync(0xffffffffu, next_claim, 0);
            w_id = cur;
        } else {
            const unsigned int g32 = claim[k & 1], g_next = claim[(k + 1) & 1];
            if (g32 >= n_groups) break;                                  // CTA-uniform
            __syncthreads();                                             // everyone has read claim[k & 1]
            if (threadIdx.x == 0) {
                claim[k & 1] = atomicAdd(args.group_counter, 1u);        // group k + 2
                known_tiles = (k + 1 + (g_next < n_groups ? 1 : 0)) * n_tiles;
                pump(it - 1 + kStages);
            }
            const long long g = g32;
            w_id = kSplit ? g : g * kWarps + warp;
        }
        // ---- per-detection prologue (warp-uniform), exact arithmetic: fit_road_planes.py:66-72, :80-83
        long long mm = w_id < n_work ? w_id : n_work - 1;          // tail warps redo the last one
        if (args.det_list) mm = args.det_list[mm];
        const long long m = w_id < n_work ? mm : args.n_det;       // >= n_det: nothing is written
        // The exact per-detection constants are only needed by the rare exact paths (verification batches, the
        // epilogue): they are parked in this warp's shared-memory slot and re-read there (GPP_LOAD_DET) instead
        // of occupying 18 registers in the loop.
#define GPP_LOAD_DET(det)                                                                   \
    Detection<ExactF32> det;                                                                \
    _Pragma("unroll") for (int i_ = 0; i_ < 3; ++i_) {                                      \
        det.dl[i_] = detx[i_]; det.dm[i_] = detx[3 + i_]; det.dr[i_] = detx[6 + i_]; det.dt[i_] = detx[9 + i_]; \
    }                                                                                       \
    _Pragma("unroll") for (int i_ = 0; i_ < 6; ++i_) det.td[i_] = detx[12 + i_]
        DetConst D;
        {
            Detection<ExactF32> det0;
            load_detection<ExactF32, ExactF32>(det0, args.boxes + 12 * mm, args.dims + 3 * mm, __ldg(args.orient + mm),
                                               args.pinv + 12 * (mm / args.dets_per_image));
#pragma unroll
            for (int i = 0; i < 6; ++i) D.td[i] = det0.td[i];
            fast_constants(D, det0);
            __syncwarp();                            // the previous detection's readers are done
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    detx[i] = det0.dl[i]; detx[3 + i] = det0.dm[i]; detx[6 + i] = det0.dr[i]; detx[9 + i] = det0.dt[i];
                }
#pragma unroll
                for (int i = 0; i < 6; ++i) detx[12 + i] = det0.td[i];
            }
            __syncwarp();
        }

        LaneState<float> st;                 // general mode (max votes not yet known to be 6)
        st.reset(FLT_MAX);
        LaneBest b6;                         // M == 6 mode
        b6.bestR = FLT_MAX; b6.bestIdx = 0;
        bool m6 = false;
        float wbest = FLT_MAX;               // VERIFIED: warp-wide best EXACT residual so far (warp-uniform)
        float wthr = __int_as_float(0x7f800000);   // VERIFIED: (wbest + mc)(1 + 2^-18), see the all-six phase
        int qn = 0;                          // VERIFIED: survivors waiting in this warp's queue (warp-uniform)
        int Mcur = -1;                       // VERIFIED: exact max-votes so far in this warp (warp-uniform)
#ifdef GPP_STATS
        unsigned int st_cnt[8] = {0, 0, 0, 0, 0, 0, 0, 1};
#endif

        for (int tt = 0; tt < n_tiles; ++tt, ++it) {
            const int t = kFree ? int(it % n_tiles) : tt;            // kFree: wherever the ring is
            const int s = int(it % kStages);
            if (kFree && threadIdx.x == 0) pump_free(it - 1 + kStages);   // tile `it` itself may not be issued yet
            mbar_wait(&full_bar[s], uint32_t((it / kStages) & 1));
            const ulonglong2 *tile = tiles + size_t(s) * kTilePairs * 2;
            const int rows = min(kTilePairs, NP - t * kTilePairs) >> 5;
            const int base_pair = t * kTilePairs;
            int r = kSplit ? warp : 0;
            if (!kVerified && !m6) {
#pragma unroll 1
                for (; r < rows; r += kRowStep) {
                    const int p = (r << 5) + lane;
                    const ulonglong2 v0 = tile[2 * p], v1 = tile[2 * p + 1];
                    const int j = 2 * (base_pair + p);
                    {
                        PairResult h;
                        eval_pair<false>(PP(), D, from_u64(v0.x), from_u64(v0.y), from_u64(v1.x), from_u64(v1.y), h);
                        const f2 R = resid_sum(h);
                        const int V0 = votes_of(lo(h.r[0]), lo(h.r[1]), lo(h.r[2]), lo(h.r[3]), lo(h.r[4]), lo(h.r[5]));
                        const int V1 = votes_of(hi(h.r[0]), hi(h.r[1]), hi(h.r[2]), hi(h.r[3]), hi(h.r[4]), hi(h.r[5]));
                        st.update(V0, lo(R), lo(h.zc) < 0.0f, j, FLT_MAX);
                        st.update(V1, hi(R), hi(h.zc) < 0.0f, j + 1, FLT_MAX);
                    }
                    if (((r / kRowStep) & 3) == 3 && __reduce_max_sync(0xffffffffu, st.M) == 6) {
                        m6 = true;                               // warp-uniform decision
                        r += kRowStep;
                        break;
                    }
                }
                if (m6) {
                    // candidates found under a lower running max are masked from now on
                    b6.bestR = (st.M == 6) ? st.bestR : FLT_MAX;
                    b6.bestIdx = st.bestIdx;
                    wbest = __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(b6.bestR)));
                }
            }
GPP_UNROLL(GPP_M6_UNROLL)
            for (; r < rows; r += kRowStep) {
                const int p = (r << 5) + lane;
                const ulonglong2 v0 = tile[2 * p], v1 = tile[2 * p + 1];
                PairResult h;
                const int j = 2 * (base_pair + p);
                if (kVerified) {
                    // ---- filter.  A plane survives iff, within its error margin m, it could matter:
                    //   all-six phase (Mcur == 6): it has six votes, passes the z-check and scores no worse than
                    //                              the warp's best exact residual so far;
                    //   general phase (Mcur < 6):  it may have MORE votes than the exact max-votes so far, or as
                    //                              many and pass the z-check and score no worse than the best.
                    // Comparisons are written so that NaN (degenerate fast arithmetic) always survives.
                    bool trig0, trig1, urgent = false;
                    if (Mcur == 6) {
                        // the residual test comes first: once the warp's best is good, almost no pair passes it,
                        // and the vote / z-check tests (and z_dir_check itself) are skipped for the whole warp.
                        // skip iff R (1 - 2^-20) - m_geo - mc > wbest; tested as R - m_geo > wthr with the
                        // warp-uniform wthr = (wbest + mc)(1 + 2^-18), which implies it (m_geo >= 0)
                        GPP_STAT(0, 1);
                        const f2 n0 = from_u64(v0.x), n1 = from_u64(v0.y), n2 = from_u64(v1.x), d4 = from_u64(v1.y);
                        // stage 1: the bottom face only (three of the six residuals, nothing that involves X_t).  Their
                        // sum is a lower bound of the residual sum, and the part of the margin that belongs to the
                        // points on the plane (w ms <= m_geo) bounds its error: ~3 of 4 iterations end here.
                        Bottom g;
                        eval_bottom<true, true>(D, n0, n1, n2, d4, g);
                        h.r[1] = sub2(PackFast::sqrt(g.na), bc(D.td[1]));
                        h.r[2] = sub2(PackFast::sqrt(g.nb), bc(D.td[2]));
                        h.r[3] = sub2(PackFast::sqrt(g.nc), bc(D.td[3]));
                        const f2 S3 = add2(add2(abs2(h.r[1]), abs2(h.r[2])), abs2(h.r[3]));
                        {
                            const f2 Slo = fma2(neg2(g.w), bc(D.ms), S3);
                            if (!__any_sync(0xffffffffu, !(lo(Slo) > wthr) || !(hi(Slo) > wthr))) continue;
                        }
                        GPP_STAT(3, 1);
                        // stage 2: X_t, the height and the two slanted edges, the full margin
                        f2 ne, nf;
                        eval_top<2, true>(D, n0, n1, n2, d4, g, h, ne, nf);
                        h.r[4] = sub2(PackFast::sqrt(abs2(ne)), bc(D.td[4]));
                        h.r[5] = sub2(PackFast::sqrt(abs2(nf)), bc(D.td[5]));
                        const f2 R = add2(add2(add2(S3, abs2(h.r[0])), abs2(h.r[4])), abs2(h.r[5]));
                        const f2 Rlo = sub2(R, h.m);
                        trig0 = !(lo(Rlo) > wthr);
                        trig1 = !(hi(Rlo) > wthr);
                        if (!__any_sync(0xffffffffu, trig0 || trig1)) continue;
                        h.m = add2(h.m, bc(D.mc));
                        finalize_margin(h, R, D);
                        const f2 rm = pk(rmax_of(lo(h.r[0]), lo(h.r[1]), lo(h.r[2]), lo(h.r[3]), lo(h.r[4]), lo(h.r[5])),
                                         rmax_of(hi(h.r[0]), hi(h.r[1]), hi(h.r[2]), hi(h.r[3]), hi(h.r[4]), hi(h.r[5])));
                        const f2 rlo = sub2(rm, h.m);               // lower bound of max |r_k|
                        // a best residual sum above 0.7 also lets planes with fewer than six votes through the
                        // residual test: drop them before the z-check / queue work
                        trig0 = trig0 && !(lo(rlo) > 0.7f);
                        trig1 = trig1 && !(hi(rlo) > 0.7f);
                        if (!__any_sync(0xffffffffu, trig0 || trig1)) continue;
                        h.finish_zc();
                        const f2 zhi = z_upper(h, D);               // upper bound of z_dir_check
                        const f2 Rl2 = sub2(R, h.m);                // lower bound of the residual sum
                        {
                            // A plane that CERTAINLY has six votes and passes the z-check bounds the final best
                            // residual by its own upper bound R + m: lower the threshold right away instead of
                            // waiting for the next exact batch (the plane itself stays queued and is verified).
                            const f2 Rhi = fma2(R, bc(1.000001f), h.m);
                            const f2 rhi = add2(rm, h.m);
                            const f2 zlo = fma2(h.zc, bc(2.0f), neg2(zhi));
                            const float e0 = (lo(rhi) <= 0.7f && lo(zlo) > 0.0f && lo(Rhi) < FLT_MAX) ? lo(Rhi) : FLT_MAX;
                            const float e1 = (hi(rhi) <= 0.7f && hi(zlo) > 0.0f && hi(Rhi) < FLT_MAX) ? hi(Rhi) : FLT_MAX;
                            const float e = __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(fminf(e0, e1))));
                            if (e < wbest) {
                                wbest = e;
                                wthr = (wbest + D.mc) * 1.0000038f;
                            }
                        }
                        trig0 = !(lo(Rl2) > wbest) && !(lo(rlo) > 0.7f) && !(lo(zhi) < 0.0f);
                        trig1 = !(hi(Rl2) > wbest) && !(hi(rlo) > 0.7f) && !(hi(zhi) < 0.0f);
                    } else {
                        GPP_STAT(Mcur >= 4 ? 1 : 2, 1);
                        eval_pair_fast<false, 1, true>(D, from_u64(v0.x), from_u64(v0.y), from_u64(v1.x), from_u64(v1.y), h);
                        const f2 R = resid_sum(h);
                        finalize_margin(h, R, D);
                        const f2 Rlo = sub2(R, h.m);
                        if (Mcur >= 4) {
                            // cheap necessary condition first (warp-uniform skip, like the all-six phase).  With
                            // p_i = max(|r_2i|, |r_2i+1|):  six votes => max p <= thr,  >= five votes => median p <= thr
                            // (at most one residual, hence at most one pair, exceeds),  >= four votes => min p <= thr.
                            // The pair matters only if it may have MORE votes than Mcur, or as many and a residual
                            // sum no worse than the best.
                            const f2 p0 = pk(fmaxf(fabsf(lo(h.r[0])), fabsf(lo(h.r[1]))), fmaxf(fabsf(hi(h.r[0])), fabsf(hi(h.r[1]))));
                            const f2 p1 = pk(fmaxf(fabsf(lo(h.r[2])), fabsf(lo(h.r[3]))), fmaxf(fabsf(hi(h.r[2])), fabsf(hi(h.r[3]))));
                            const f2 p2 = pk(fmaxf(fabsf(lo(h.r[4])), fabsf(lo(h.r[5]))), fmaxf(fabsf(hi(h.r[4])), fabsf(hi(h.r[5]))));
                            const f2 pmax = pk(max3f(lo(p0), lo(p1), lo(p2)), max3f(hi(p0), hi(p1), hi(p2)));
                            const f2 pmin = pk(fminf(fminf(lo(p0), lo(p1)), lo(p2)), fminf(fminf(hi(p0), hi(p1)), hi(p2)));
                            const f2 pmed = pk(fmaxf(fminf(lo(p0), lo(p1)), fminf(fmaxf(lo(p0), lo(p1)), lo(p2))),
                                               fmaxf(fminf(hi(p0), hi(p1)), fminf(fmaxf(hi(p0), hi(p1)), hi(p2))));
                            const f2 more = sub2(Mcur == 5 ? pmax : pmed, h.m);      // > 0.7: cannot have more votes
                            const f2 same = sub2(Mcur == 5 ? pmed : pmin, h.m);      // > 0.7: cannot have as many
                            const bool nan0 = !(lo(R) == lo(R)), nan1 = !(hi(R) == hi(R));   // degenerate: full test
                            const bool may0 = nan0 || !(lo(more) > 0.7f) || (!(lo(same) > 0.7f) && !(lo(Rlo) > wbest));
                            const bool may1 = nan1 || !(hi(more) > 0.7f) || (!(hi(same) > 0.7f) && !(hi(Rlo) > wbest));
                            if (!__any_sync(0xffffffffu, may0 || may1)) continue;
                        }
                        GPP_STAT(4, 1);
                        h.finish_zc();
                        const f2 zhi = z_upper(h, D);
                        const int V0 = loose_votes(h, false), V1 = loose_votes(h, true);
                        const bool k0 = V0 == Mcur && !(lo(zhi) < 0.0f) && !(lo(Rlo) > wbest);
                        const bool k1 = V1 == Mcur && !(hi(zhi) < 0.0f) && !(hi(Rlo) > wbest);
                        if (Mcur >= 4 && __any_sync(0xffffffffu, k0 || k1)) {
                            // same early bound as in the all-six phase: exactly Mcur votes for certain, z-check passed
                            const f2 Rhi = fma2(R, bc(1.000001f), h.m);
                            const f2 zlo = fma2(h.zc, bc(2.0f), neg2(zhi));
                            const bool c0 = k0 && strict_votes(h, false) == Mcur && lo(zlo) > 0.0f && lo(Rhi) < FLT_MAX;
                            const bool c1 = k1 && strict_votes(h, true) == Mcur && hi(zlo) > 0.0f && hi(Rhi) < FLT_MAX;
                            const float e = __uint_as_float(__reduce_min_sync(
                                0xffffffffu, __float_as_uint(fminf(c0 ? lo(Rhi) : FLT_MAX, c1 ? hi(Rhi) : FLT_MAX))));
                            wbest = fminf(wbest, e);
                        }
                        trig0 = (V0 > Mcur) || (k0 && !(lo(Rlo) > wbest));
                        trig1 = (V1 > Mcur) || (k1 && !(hi(Rlo) > wbest));
                        urgent = (V0 > Mcur) || (V1 > Mcur);            // may raise max-votes: verify right away
                    }
                    const bool q0 = trig0 && (j < N), q1 = trig1 && (j + 1 < N);
                    const unsigned b0 = __ballot_sync(0xffffffffu, q0), b1 = __ballot_sync(0xffffffffu, q1);
                    if (b0 | b1) {
                        // ---- queue the survivors; the whole warp re-evaluates them 32 at a time (exact)
                        const unsigned below = (1u << lane) - 1u;
                        if (q0) queue[qn + __popc(b0 & below)] = j;
                        qn += __popc(b0);
                        if (q1) queue[qn + __popc(b1 & below)] = j + 1;
                        qn += __popc(b1);
                        __syncwarp();
                        const bool flush_all = __any_sync(0xffffffffu, urgent);
                        if (qn >= 32 || flush_all) {
                            GPP_STAT(6, 1);
                            GPP_STAT(5, qn);
                            GPP_LOAD_DET(det);
                            while (qn >= 32) {
                                qn -= 32;
                                verify_general(det, args.planes, queue[qn + lane], st);
                            }
                            if (flush_all && qn > 0) {
                                if (lane < qn) verify_general(det, args.planes, queue[lane], st);
                                qn = 0;
                            }
                            __syncwarp();
                            const int Mnew = __reduce_max_sync(0xffffffffu, st.M);
                            const float wnew = __uint_as_float(__reduce_min_sync(
                                0xffffffffu, __float_as_uint(st.M == Mnew ? st.bestR : FLT_MAX)));
                            wbest = (Mnew == Mcur) ? fminf(wbest, wnew) : wnew;   // early bounds stay valid at the same max-votes
                            Mcur = Mnew;
                            wthr = (wbest + D.mc) * 1.0000038f;
                        }
                    }
                } else {
                    // two stages like the VERIFIED all-six phase: the three bottom-face residuals first; their sum never
                    // exceeds the full sum (floating-point addition of non-negative terms is monotone), so leaving
                    // here decides exactly what the full test below would decide
                    const f2 n0 = from_u64(v0.x), n1 = from_u64(v0.y), n2 = from_u64(v1.x), d4 = from_u64(v1.y);
                    Bottom g;
                    eval_bottom<true, false>(D, n0, n1, n2, d4, g);
                    h.r[1] = sub2(PackFast::sqrt(g.na), bc(D.td[1]));
                    h.r[2] = sub2(PackFast::sqrt(g.nb), bc(D.td[2]));
                    h.r[3] = sub2(PackFast::sqrt(g.nc), bc(D.td[3]));
                    {
                        const f2 S3 = add2(add2(abs2(h.r[1]), abs2(h.r[2])), abs2(h.r[3]));
                        if (!__any_sync(0xffffffffu, !(lo(S3) > wbest) || !(hi(S3) > wbest))) continue;
                    }
                    f2 ne, nf;
                    eval_top<0, true>(D, n0, n1, n2, d4, g, h, ne, nf);
                    h.r[4] = sub2(PackFast::sqrt(abs2(ne)), bc(D.td[4]));
                    h.r[5] = sub2(PackFast::sqrt(abs2(nf)), bc(D.td[5]));
                    const f2 R = resid_sum(h);
                    // only a pair that scores no worse than the warp's best so far can change the result
                    if (__any_sync(0xffffffffu, !(lo(R) > wbest) || !(hi(R) > wbest))) {
                        h.finish_zc();
                        b6.update6(rmax_of(lo(h.r[0]), lo(h.r[1]), lo(h.r[2]), lo(h.r[3]), lo(h.r[4]), lo(h.r[5])),
                                   lo(h.zc), lo(R), j);
                        b6.update6(rmax_of(hi(h.r[0]), hi(h.r[1]), hi(h.r[2]), hi(h.r[3]), hi(h.r[4]), hi(h.r[5])),
                                   hi(h.zc), hi(R), j + 1);
                        wbest = __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(b6.bestR)));
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
            if (!kFree && threadIdx.x == 0) pump(it - 1 + kStages);   // skewed by one tile: rarely waits for the slowest warp
            __syncwarp();
        }

        GPP_LOAD_DET(det);                           // exact constants for the rest of this detection
        if (kVerified) {
            GPP_STAT(5, qn);
            if (lane < qn) verify_general(det, args.planes, queue[lane], st);   // the last partial batch
#ifdef GPP_STATS
            if (lane == 0)
                for (int i = 0; i < 8; ++i) atomicAdd(&g_stats[i], (unsigned long long)st_cnt[i]);
#endif
            qn = 0;
            __syncwarp();
            m6 = false;                              // the epilogue takes (max-votes, best) from `st`
        }
        // ---- epilogue: warp reduction, lazy first-masked search, exact recompute of the winner
        int Mw;
        float rbest;
        int idx;
        if (m6) {
            Mw = 6;
            rbest = b6.bestR;
            idx = b6.bestIdx;
        } else {
            Mw = __reduce_max_sync(0xffffffffu, st.M);
            rbest = (st.M == Mw) ? st.bestR : FLT_MAX;
            idx = st.bestIdx;
        }
        rbest = warp_min_first(rbest, idx);
        if (kSplit) {
            // merge the warps' partial results (double-buffered by group parity: one barrier per group)
            WarpPartial<float> *buf = partial + (k & 1) * kWarps;
            if (lane == 0) { buf[warp].r = rbest; buf[warp].M = Mw; buf[warp].idx = idx; }
            __syncthreads();
            if (warp != 0) continue;
            const int Ml = lane < kWarps ? buf[lane].M : -1;
            Mw = __reduce_max_sync(0xffffffffu, Ml);
            rbest = (lane < kWarps && Ml == Mw) ? buf[lane].r : FLT_MAX;
            idx = lane < kWarps ? buf[lane].idx : 0;
            rbest = warp_min_first(rbest, idx);
        }
        const bool have_cand = rbest < FLT_MAX;
        bool sentinel = false;
        if (!(rbest < 100.0f)) {
            int first_masked = -1;
            for (int p0 = 0; 2 * p0 < N && first_masked < 0; p0 += 32) {
                const int p = p0 + lane;                         // pair index; the padded DB covers it
                const ulonglong2 v0 = reinterpret_cast<const ulonglong2 *>(args.pairs)[2 * p];
                const ulonglong2 v1 = reinterpret_cast<const ulonglong2 *>(args.pairs)[2 * p + 1];
                int V0, V1;
                bool z0, z1;
                if (kVerified) {
                    const f2 a01 = from_u64(v0.x), b01 = from_u64(v0.y), c01 = from_u64(v1.x), d01 = from_u64(v1.y);
                    float Rx;
                    exact_one(det, lo(a01), lo(b01), lo(c01), lo(d01), V0, Rx, z0);
                    exact_one(det, hi(a01), hi(b01), hi(c01), hi(d01), V1, Rx, z1);
                } else {
                    PairResult h;
                    eval_pair<false>(PP(), D, from_u64(v0.x), from_u64(v0.y), from_u64(v1.x), from_u64(v1.y), h);
                    V0 = votes_of(lo(h.r[0]), lo(h.r[1]), lo(h.r[2]), lo(h.r[3]), lo(h.r[4]), lo(h.r[5]));
                    V1 = votes_of(hi(h.r[0]), hi(h.r[1]), hi(h.r[2]), hi(h.r[3]), hi(h.r[4]), hi(h.r[5]));
                    z0 = lo(h.zc) < 0.0f;
                    z1 = hi(h.zc) < 0.0f;
                }
                const bool mk0 = (2 * p < N) && ((V0 < Mw) || z0);
                const bool mk1 = (2 * p + 1 < N) && ((V1 < Mw) || z1);
                const unsigned b0 = __ballot_sync(0xffffffffu, mk0), b1 = __ballot_sync(0xffffffffu, mk1);
                if (b0 | b1) {
                    const int f0 = b0 ? 2 * (p0 + __ffs(b0) - 1) : 0x7fffffff;
                    const int f1 = b1 ? 2 * (p0 + __ffs(b1) - 1) + 1 : 0x7fffffff;
                    first_masked = min(f0, f1);
                }
            }
            if (first_masked >= 0) {
                if (!have_cand || 100.0f < rbest || (100.0f == rbest && first_masked < idx)) {
                    sentinel = true;
                    idx = first_masked;
                }
            } else if (!have_cand) {
                idx = 0;
            }
        }
        if (m < args.n_det && lane == 0) {
            const float4 pl = args.planes[idx];
            float X[4][3];
            int V; float R; bool zneg;
            hypothesis<ExactF32>(det, pl.x, pl.y, pl.z, pl.w, X, V, R, zneg);
            const float rr = sentinel ? 100.0f : R;
            float *kp = args.keypoints + 12 * m;
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int i = 0; i < 3; ++i) kp[3 * k + i] = X[k][i];
            float *kpl = args.keyplanes + 4 * m;
            kpl[0] = pl.x; kpl[1] = pl.y; kpl[2] = pl.z; kpl[3] = pl.w;
            args.residuals[m] = __fdiv_rn(rr, 6.0f);
            if (args.best) args.best[m] = idx;
        }
    }
    if (kFree) {
        // ---- out of work: keep releasing the tiles the other warps still stream, leave when nobody needs any
        __syncwarp();
        if (lane == 0) atomicSub(active, 1);          // every announcement of this warp precedes this
        for (;;) {
            const long long need = (long long)*reinterpret_cast<volatile unsigned long long *>(need_until);
            if (it < need) {
                const int s = int(it % kStages);
                if (threadIdx.x == 0) pump_free(it - 1 + kStages);
                mbar_wait(&full_bar[s], uint32_t((it / kStages) & 1));
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[s]);
                ++it;
                continue;
            }
            if (*reinterpret_cast<volatile int *>(active) == 0) {
                __threadfence_block();
                if (it >= (long long)*reinterpret_cast<volatile unsigned long long *>(need_until)) break;
            } else {
                __nanosleep(500);
            }
        }
    }
}

}  // namespace gpp
