// Packed-pair arithmetic of the FAST and VERIFIED modes (the kernel that drives it is gpp_poll3.cuh).
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
struct NoHook {
    __device__ __forceinline__ void operator()(const DetConst &) const {}
};
// `before_r0`: called between the plane arithmetic and the first use of a residual target (the resident kernel fetches
// the targets it has parked in shared memory there, gpp_poll3.cuh).
template <int kMargin, bool kLazyZ, class D_, class Hook>
__device__ __forceinline__ void eval_top(D_ &D, f2 n0, f2 n1, f2 n2, f2 d4, const Bottom &g,
                                         PairResult &out, f2 &ne, f2 &nf, const Hook &before_r0) {
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
    before_r0(D);
    out.r[0] = sub2(abs2(q), bc(D.td[0]));
}
template <int kMargin, bool kLazyZ>
__device__ __forceinline__ void eval_top(const DetConst &D, f2 n0, f2 n1, f2 n2, f2 d4, const Bottom &g,
                                         PairResult &out, f2 &ne, f2 &nf) {
    eval_top<kMargin, kLazyZ>(D, n0, n1, n2, d4, g, out, ne, nf, NoHook());
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

// exact scalar evaluation of one plane of a pair (VERIFIED mode: general path, re-evaluation, epilogue)
__device__ __forceinline__ void exact_one(const Detection<ExactF32> &de, float n0, float n1, float n2, float d4,
                                          int &V, float &R, bool &zneg) {
    float X[4][3];
    hypothesis<ExactF32>(de, n0, n1, n2, d4, X, V, R, zneg);
}

constexpr int kVerifyQueue = 96;

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

}  // namespace gpp
