// The polling kernel (all four arithmetic modes) -- since round 2 the only one.
//
// Round 1 streamed the database through ONE tile ring per CTA, so the eight warps of a CTA ran at the pace of their
// slowest detection (~20 % of the warp time was spent waiting for the next tile).  Here no two warps share anything:
//   * one persistent CTA per SM (32 warps at 64 registers; fp64: 16 at 128); the first `resident_rows` rows (1 row =
//     32 plane pairs = 64 planes = 1 KB) of the pair-interleaved database are staged ONCE per CTA into shared memory by
//     1-D TMA bulk copies (cp.async.bulk + mbarrier, SASS UBLKCP) -- up to 212 KB, i.e. 13.5k of the 21.6k planes of
//     the largest shipped database, all of the smaller ones;
//   * the remaining rows are read by each warp straight from L2 (the database never leaves L2), one row ahead of
//     its use (in-place register prefetch, ld.global.nc), so that the L2 latency hides behind the previous row's
//     arithmetic;
//   * every warp claims its own work items from a device counter.  An item is (detection, plane segment): large
//     batches use one segment per detection, small batches cut every detection into up to 32 segments so that a
//     single image still fills the 148 SMs; the warp that finishes the last segment of a detection merges the
//     partial (max-votes, residual, index) results and runs the epilogue;
//   * rows that repeat the previous row of their image (FilterDetections' -1 padding) are skipped when claimed; the
//     warp that finishes a detection also writes its results to the identical rows that follow it.  One launch per
//     call, no work lists, no memset: the counters are reset by the last warp / CTA that uses them.
// Modes: FAST and VERIFIED scan the pair-interleaved database with the packed arithmetic of gpp_poll2.cuh; EXACT
// (fp32) and F64 scan the plain database with the scalar hypothesis of gpp_math.cuh, one plane per lane, straight from
// L2 (they are bound by their own arithmetic: ~165 / ~300 instructions per 16 / 32 bytes).  The epilogue can run the
// two steps that follow polling in the reference's driver -- pose recovery and the KITTI record (gpp_pose.cuh) -- on
// the winner it has just recomputed (fit_road_planes.py:49-139, bin/run_network.py:137-247, :297-327).
#pragma once
#include <type_traits>

#include "gpp_poll2.cuh"
#include "gpp_pose.cuh"

#ifndef GPP_LAZY_Z6
#define GPP_LAZY_Z6 true    /* z_dir_check of the all-six stage 2 is formed only by the rows that get that far (false: at once; measured 17.51 ms against 17.35 ms on C4) */
#endif
#ifndef GPP_SEG_MAJOR
#define GPP_SEG_MAJOR 1
#endif
#ifndef GPP_GEN_EXIT
#define GPP_GEN_EXIT 1      /* general phase: rows leave after the bottom face when no plane can matter (0: measured equal at the
                             benchmark's 1.5 px key-point noise, 4 % / 6 % slower at 4 px / 10 px) */
#endif
#ifndef GPP_STAGE2
#define GPP_STAGE2 1      /* 0: experiment -- stage-1 survivors of the all-six phase go straight to the exact queue */
#endif

namespace gpp {

struct SegPartial {       // result of one plane segment of a detection (16 bytes)
    double r;
    int M, idx;
};

enum { kModeFast = 0, kModeVerified = 1, kModeExact = 2, kModeF64 = 3 };

struct PollArgs3 {
    const float *boxes, *dims, *pinv;
    const int32_t *orient;
    const u64 *pairs;            // pair-interleaved normalised fp32 DB, {a0,a1,b0,b1,c0,c1,d0,d1} per pair, in scan order: FAST / VERIFIED
    const int32_t *scan_index;   // plane index of every position of `pairs` (gpp_order.cu)
    const float4 *planes;        // plain normalised fp32 DB (index order): exact re-evaluation, epilogue
    const float4 *planes_scan;   // the same in scan order: EXACT scan
    const double4 *planes64;     // fp64 DB: F64 scan
    int n_planes, n_pairs_padded, dets_per_image;
    long long n_det;
    void *keypoints, *keyplanes, *residuals;     // float, or double in the F64 mode
    long long *best;
    // optional fused steps after polling (run_network.py:137-247, :297-327); nullptr = off.  pose_kitti needs the others.
    float *pose_locations, *pose_angles, *pose_dimensions, *pose_kitti;
    // schedule
    int det_stride;              // 1; n > 1 polls rows 0, n, 2n, ... only (runtime audit), without the repeated-row logic
    int resident_rows;           // rows of `pairs` kept in shared memory (0 = stream everything from L2)
    int n_seg;                   // plane segments per detection: a segment takes the rows seg, seg + n_seg, ... of the scan
                                 // order (rows of 64 planes for the packed scans, of 32 for the scalar ones), so every
                                 // segment starts in the sample rows that order puts first
    unsigned long long *claim;   // [0] work-item counter, [1] CTAs that have left; zero at launch, reset by the last CTA
    SegPartial *partials;        // [n_det * n_seg]                      (n_seg > 1 only)
    unsigned int *seg_arrived;   // [n_det], zero at launch, reset by the last segment of the detection
    unsigned long long *seg_best;   // [n_det], shared (max-votes, best residual) key of a detection's segments, same life cycle
    float *seg_consts;           // [n_det][kSharedConsts]: the per-detection constants, written by segment 0 (packed modes)
    unsigned int *seg_ready;     // [n_det]: seg_consts[slot] is complete; same life cycle
};

__device__ __forceinline__ Detection<ExactF32> load_det_exact(const float *detx) {
    Detection<ExactF32> det;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        det.dl[i] = detx[i]; det.dm[i] = detx[3 + i]; det.dr[i] = detx[6 + i]; det.dt[i] = detx[9 + i];
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) det.td[i] = detx[12 + i];
    return det;
}

// The detection constants that only the rare paths of the VERIFIED filter read (stage 2 of the all-six phase, the general
// phase, the threshold updates) are parked in the warp's shared-memory slot and fetched where those paths begin, so that
// the hot first stage keeps ten constants in registers instead of nineteen.
constexpr int kColdOffset = 20;      // floats; the exact constants occupy [0, 18)
// Two groups, each fetched where its first reader is (loaded earlier, the second group was parked in local memory across
// eval_top by the compiler -- four local stores and three loads per stage-2 row): A = what eval_top reads, B = the
// targets of the residuals that involve X_t and the rounding term of the margin.
__device__ __forceinline__ void store_cold(float *detx, const DetConst &D) {
    float *c = detx + kColdOffset;
    c[0] = D.ft[0]; c[1] = D.ft[1]; c[2] = D.T; c[3] = D.G;
    c[4] = D.msT; c[5] = D.msq;
    c[8] = D.td[0]; c[9] = D.td[4]; c[10] = D.td[5]; c[11] = D.mc;
}
__device__ __forceinline__ void load_cold_a(DetConst &D, const float *detx) {
    const uint32_t a = smem_u32(detx + kColdOffset);
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(D.ft[0]), "=f"(D.ft[1]), "=f"(D.T), "=f"(D.G) : "r"(a));
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+16];" : "=f"(D.msT), "=f"(D.msq) : "r"(a));
}
__device__ __forceinline__ void load_cold_b(DetConst &D, const float *detx) {
    const uint32_t a = smem_u32(detx + kColdOffset);
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+32];" : "=f"(D.td[0]), "=f"(D.td[4]), "=f"(D.td[5]), "=f"(D.mc) : "r"(a));
}
__device__ __forceinline__ void load_cold(DetConst &D, const float *detx) {
    load_cold_a(D, detx);
    load_cold_b(D, detx);
}
// eval_top hook: fetches group B between the plane arithmetic and the first residual that needs it
struct ColdB {
    const float *detx;
    __device__ __forceinline__ void operator()(DetConst &D) const { load_cold_b(D, detx); }
};

// (max-votes, best residual) as one 64-bit key that grows when the pair improves: more votes first, then a smaller
// residual (r >= +0, never NaN: the bit pattern orders like the value)
__device__ __forceinline__ unsigned long long seg_key(int M, float r) {
    return ((unsigned long long)(unsigned)(M + 1) << 32) | (unsigned long long)(0xffffffffu - __float_as_uint(r));
}

// Where a lane finds its pair of row r: shared memory for the resident rows, L2 (read-only path) for the others.
// The scan loops keep ONE copy of the lane's 32 bytes in registers: as soon as the dot products of row r are
// formed, the registers are refilled with row r + 1 (so the L2 / shared-memory latency hides behind the rest of
// row r), and the few rows that get past the first filter stage read their pair again (an L1 / shared-memory hit).
struct RowSource {
    // loop-carried state (kept as running values so that the compiler has nothing to recompute per row)
    uint32_t saddr;            // shared-memory address this lane's pair of the CURRENT row has / would have if resident
    int rows_left;             // rows of the segment not yet polled, the current one included (warp-uniform)
    int res_left;              // resident rows from the current one on (<= 0: the current row is streamed)
    // constants
    uint32_t sbase;            // shared-memory address of the resident rows
    uint32_t step_bytes;       // distance between two rows of the segment (1 KB x segments per detection)
    const unsigned char *gbase;   // global address of the pair database
    __device__ __forceinline__ void init(uint32_t sbase_, const void *pairs, int res, int r_begin, int step, int n_rows,
                                         int lane) {
        sbase = sbase_;
        gbase = static_cast<const unsigned char *>(pairs);
        saddr = sbase_ + (uint32_t(r_begin) << 10) + 32u * lane;
        step_bytes = uint32_t(step) << 10;
        rows_left = (n_rows - r_begin + step - 1) / step;
        res_left = res - r_begin;
    }
    __device__ __forceinline__ void load(bool resident, uint32_t addr, ulonglong2 &a, ulonglong2 &b) const {
        if (resident) {
            asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(a.x), "=l"(a.y) : "r"(addr));
            asm volatile("ld.shared.v2.u64 {%0, %1}, [%2+16];" : "=l"(b.x), "=l"(b.y) : "r"(addr));
        } else {
            const ulonglong2 *p = reinterpret_cast<const ulonglong2 *>(gbase + (addr - sbase));
            a = __ldg(p);
            b = __ldg(p + 1);
        }
    }
    __device__ __forceinline__ bool more() const { return rows_left > 1; }
    // (A branch-free form -- the current row again on the last row, so that the compiler cannot sink the loads behind the
    // `more()` test to the end of the row -- was measured: the loads do move up and the wait at the top of the next row
    // goes, but C4 takes 17.29 ms against 17.11 ms and the FAST mode 16.1 ms against 15.3 ms.)
    template <bool kStep>
    __device__ __forceinline__ void load_next(ulonglong2 &a, ulonglong2 &b) const {
        if (kStep) load(res_left > int(step_bytes >> 10), saddr + step_bytes, a, b);
        else load(res_left > 1, saddr + 1024u, a, b);
    }
    __device__ __forceinline__ void load_again(ulonglong2 &a, ulonglong2 &b) const { load(res_left > 0, saddr, a, b); }
    template <bool kStep>
    __device__ __forceinline__ void advance() {
        saddr += kStep ? step_bytes : 1024u;
        // with a run-time step the compiler otherwise keeps four running addresses (this row, the next one, both again as
        // global offsets) and spills one of them: one running address, the others derived where they are used
        if (kStep) asm volatile("" : "+r"(saddr));
        --rows_left;
        res_left -= kStep ? int(step_bytes >> 10) : 1;
    }
    __device__ __forceinline__ int position() const { return int((saddr - sbase) >> 4); }   // of the first plane of the lane's pair
};

// ------------------------------------------------------------------ VERIFIED: filter + exact bookkeeping of a warp
// The body of the row function is the filter of poll2_kernel<..., kVMode = 1> (see the comments there and DESIGN.md
// section 4.1): all-six phase in two stages, general phase with the order-statistic pre-test, early bounds from
// certain planes, survivors queued and re-evaluated 32 at a time in the exact arithmetic.
// development counters (-DGPP_STATS, printed by launch_poll3): [0] rows in the all-six phase, [1] of them past stage 1,
// [2] rows in the general phase, [3] exact verifications, [4] flushes, [5] work items, [6] items polled in the
// identical-rays form, [7] rows past the whole filter (something queued), [8] general-phase rows (max-votes >= 4) past
// the bottom-face test, [9] general-phase rows at max-votes >= 4
#ifdef GPP_TIMELINE   /* development only: per-item time stamps (globaltimer, ns), printed by launch_poll3 */
__device__ unsigned long long g_timeline[8192][6];
__device__ unsigned int g_timeline_n;
__device__ __forceinline__ unsigned long long gpp_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define GPP_TL(var) const unsigned long long var = gpp_now()
#else
#define GPP_TL(var) ((void)0)
#endif
#ifdef GPP_STATS
__device__ unsigned long long g_stats3[10];
#define GPP_STAT3(i, n) (stat[i] += (n))
#else
#define GPP_STAT3(i, n) ((void)0)
#endif

// Exact re-evaluation of the warp's queue, out of line: ONE copy of the exact hypothesis for the three places that drain
// the queue (the two scan loops and the end of a segment) instead of five inlined ones -- 15 KB less code for the 32
// warps of a CTA to share the instruction cache with -- and none of its registers in the scan loops' allocation.  It
// runs a few times per detection.  `n` entries starting at queue[first]; returns the lane's updated selection state.
struct LaneSel {
    int M;
    float bestR;
    int bestIdx;
};
static __device__ __noinline__ LaneSel verify_queue(const float *detx, const float4 *__restrict__ planes, const int *queue,
                                             int first, int n, int lane, int M, float bestR, int bestIdx) {
    const Detection<ExactF32> det = load_det_exact(detx);
    LaneState<float> st;
    st.M = M; st.bestR = bestR; st.bestIdx = bestIdx;
#pragma unroll 1
    for (int at = first; at < first + n; at += 32)
        if (at + lane < first + n) verify_general(det, planes, queue[at + lane], st);
    LaneSel out;
    out.M = st.M; out.bestR = st.bestR; out.bestIdx = st.bestIdx;
    return out;
}

struct VerifiedScan {
#ifdef GPP_STATS
    unsigned int stat[10];
#endif
    LaneState<float> st;   // exact selection state of this lane (max-votes, best residual, index)
    float wbest;           // warp-uniform: best EXACT residual so far at max-votes Mcur (or an early upper bound of it)
    float wthr;            // (wbest + mc)(1 + 2^-18)
    int qn;                // survivors waiting in the warp's queue (warp-uniform)
    int Mcur;              // exact max-votes so far (warp-uniform)

    __device__ __forceinline__ void begin() {
        st.reset(FLT_MAX);
        wbest = FLT_MAX;
        wthr = __int_as_float(0x7f800000);
        qn = 0;
        Mcur = -1;
#ifdef GPP_STATS
        for (int i = 0; i < 10; ++i) stat[i] = 0;
        stat[5] = 1;
#endif
    }

    // adopt a bound found by another segment of the same detection (exact values, published through seg_best)
    __device__ __forceinline__ void adopt(unsigned long long key, const float *detx) {
        const int M = int(unsigned(key >> 32)) - 1;
        const float r = __uint_as_float(0xffffffffu - unsigned(key & 0xffffffffu));
        if (M > Mcur || (M == Mcur && r < wbest)) {
            Mcur = M;
            wbest = r;
            wthr = (wbest + detx[kColdOffset + 11]) * 1.0000038f;       // mc
        }
    }

    __device__ __forceinline__ void flush(const float *detx, const float4 *__restrict__ planes, const int *queue,
                                          int lane, bool all, const DetConst &D) {
        GPP_STAT3(4, 1);
        GPP_STAT3(3, all ? qn : (qn & ~31));
        {
            // whole batches of 32 from the top of the queue; everything when a plane may raise max-votes
            const int keep = all ? 0 : (qn & 31);
            const LaneSel r = verify_queue(detx, planes, queue, keep, qn - keep, lane, st.M, st.bestR, st.bestIdx);
            st.M = r.M; st.bestR = r.bestR; st.bestIdx = r.bestIdx;
            qn = keep;
        }
        __syncwarp();
        const int Mnew = __reduce_max_sync(0xffffffffu, st.M);
        const float wnew = __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(st.M == Mnew ? st.bestR : FLT_MAX)));
        if (Mnew > Mcur) {
            Mcur = Mnew;
            wbest = wnew;
        } else if (Mnew == Mcur) {
            wbest = fminf(wbest, wnew);       // early bounds (and adopted ones) stay valid at the same max-votes
        }                                     // Mnew < Mcur: the bound was adopted from another segment, keep it
        wthr = (wbest + D.mc) * 1.0000038f;
    }

    // A row of the general phase (max-votes so far below 6).  false: no plane of the row can matter.
    __device__ __forceinline__ bool general_row(DetConst &D, const RowSource &src, Bottom &g, PairResult &h, bool &trig0,
                                                bool &trig1, bool &urgent, const float *detx) {
        GPP_STAT3(2, 1);
        load_cold(D, detx);
        eval_bottom_rest<false, true>(D, g);
        if (GPP_GEN_EXIT && Mcur >= 4) {
            // The bottom face alone: a plane has at most 3 + (bottom-face votes) votes, and its residual sum is at
            // least S3.  It can have MORE votes than Mcur only if Mcur - 2 bottom residuals may be within 0.7, and AS
            // MANY only if Mcur - 3 may be and S3 may stay below the best: order statistics of the three.  (Rows of
            // similar planes -- gpp_order.cu -- mostly fail together; detections without a six-vote plane spend their
            // whole scan here.)  m1 bounds |fast - exact| of each of the three and of their sum (stage 1).
            const f2 a1 = abs2(sub2(PackFast::sqrt(g.na), bc(D.td[1]))), a2 = abs2(sub2(PackFast::sqrt(g.nb), bc(D.td[2]))),
                     a3 = abs2(sub2(PackFast::sqrt(g.nc), bc(D.td[3])));
            const f2 S3 = add2(add2(a1, a2), a3);
            const f2 m1 = fma2(S3, bc(9.5367431640625e-07f), fma2(g.w, bc(D.ms), bc(D.mc)));
            const float md0 = fmaxf(fminf(lo(a1), lo(a2)), fminf(fmaxf(lo(a1), lo(a2)), lo(a3)));
            const float md1 = fmaxf(fminf(hi(a1), hi(a2)), fminf(fmaxf(hi(a1), hi(a2)), hi(a3)));
            float more0, more1, same0, same1;
            if (Mcur == 5) {
                more0 = max3f(lo(a1), lo(a2), lo(a3)); more1 = max3f(hi(a1), hi(a2), hi(a3));
                same0 = md0; same1 = md1;
            } else {
                more0 = md0; more1 = md1;
                same0 = fminf(fminf(lo(a1), lo(a2)), lo(a3)); same1 = fminf(fminf(hi(a1), hi(a2)), hi(a3));
            }
            const f2 Slo = sub2(S3, m1);
            const bool nan0 = !(lo(S3) == lo(S3)), nan1 = !(hi(S3) == hi(S3));      // degenerate: full test
            const bool may0 = nan0 || !(more0 - lo(m1) > 0.7f) || (!(same0 - lo(m1) > 0.7f) && !(lo(Slo) > wbest));
            const bool may1 = nan1 || !(more1 - hi(m1) > 0.7f) || (!(same1 - hi(m1) > 0.7f) && !(hi(Slo) > wbest));
            GPP_STAT3(9, 1);
            if (!__any_sync(0xffffffffu, may0 || may1)) return false;
            GPP_STAT3(8, 1);
            // The rows that go on form the three residuals again after eval_top (nine instructions for the 4 % of the
            // rows that get here): kept across it, they are what pushes the kernel past 64 registers.
            asm volatile("" : "+l"(g.na.v), "+l"(g.nb.v), "+l"(g.nc.v));
        }
        {
            ulonglong2 v0, v1;
            src.load_again(v0, v1);
            f2 ne, nf;
            eval_top<1, true>(D, from_u64(v0.x), from_u64(v0.y), from_u64(v1.x), from_u64(v1.y), g, h, ne, nf);
            h.r[1] = sub2(PackFast::sqrt(g.na), bc(D.td[1]));
            h.r[2] = sub2(PackFast::sqrt(g.nb), bc(D.td[2]));
            h.r[3] = sub2(PackFast::sqrt(g.nc), bc(D.td[3]));
            h.r[4] = sub2(PackFast::sqrt(abs2(ne)), bc(D.td[4]));
            h.r[5] = sub2(PackFast::sqrt(abs2(nf)), bc(D.td[5]));
        }
        const f2 R = resid_sum(h);
        finalize_margin(h, R, D);
        const f2 Rlo = sub2(R, h.m);
        if (Mcur >= 4) {
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
            if (!__any_sync(0xffffffffu, may0 || may1)) return false;
        }
        h.finish_zc();
        const f2 zhi = z_upper(h, D);
        const int V0 = loose_votes(h, false), V1 = loose_votes(h, true);
        const bool k0 = V0 == Mcur && !(lo(zhi) < 0.0f) && !(lo(Rlo) > wbest);
        const bool k1 = V1 == Mcur && !(hi(zhi) < 0.0f) && !(hi(Rlo) > wbest);
        if (Mcur >= 4 && __any_sync(0xffffffffu, k0 || k1)) {
            const f2 Rhi = fma2(R, bc(1.000001f), h.m);
            const f2 zlo = fma2(h.zc, bc(2.0f), neg2(zhi));
            const bool c0 = k0 && strict_votes(h, false) == Mcur && lo(zlo) > 0.0f && lo(Rhi) < FLT_MAX;
            const bool c1 = k1 && strict_votes(h, true) == Mcur && hi(zlo) > 0.0f && hi(Rhi) < FLT_MAX;
            const float e = __uint_as_float(__reduce_min_sync(
                0xffffffffu, __float_as_uint(fminf(c0 ? lo(Rhi) : FLT_MAX, c1 ? hi(Rhi) : FLT_MAX))));
            wbest = fminf(wbest, e);
        }
        urgent = (V0 > Mcur) || (V1 > Mcur);            // may raise max-votes: verify right away
        if (__any_sync(0xffffffffu, urgent)) {
            // Votes that some plane of this row CERTAINLY has: if that is more than Mcur, max-votes ends up at L or above
            // whatever the others do, and only the planes that may reach L can still matter.  (The first row of every
            // scan starts from Mcur = -1: without this all its 64 planes would be verified.)
            const int L = __reduce_max_sync(0xffffffffu, max(strict_votes(h, false), strict_votes(h, true)));
            if (L > Mcur) {
                trig0 = V0 >= L;
                trig1 = V1 >= L;
                return true;
            }
        }
        trig0 = (V0 > Mcur) || (k0 && !(lo(Rlo) > wbest));
        trig1 = (V1 > Mcur) || (k1 && !(hi(Rlo) > wbest));
        return true;
    }

    // survivors of a row -> the warp's queue (plane indices); a full batch, or a plane that may raise max-votes, is
    // verified at once
    __device__ __forceinline__ void enqueue(bool trig0, bool trig1, bool urgent, const RowSource &src, const int N,
                                            const int lane, int *queue, const float *detx,
                                            const float4 *__restrict__ planes, const int32_t *__restrict__ scan_index,
                                            const DetConst &D) {
        const int pos = src.position();
        const bool q0 = trig0 && (pos < N), q1 = trig1 && (pos + 1 < N);
        const unsigned b0 = __ballot_sync(0xffffffffu, q0), b1 = __ballot_sync(0xffffffffu, q1);
        if (b0 | b1) {
            GPP_STAT3(7, 1);
            const unsigned below = (1u << lane) - 1u;
            if (q0) queue[qn + __popc(b0 & below)] = __ldg(scan_index + pos);       // the queue holds plane indices
            qn += __popc(b0);
            if (q1) queue[qn + __popc(b1 & below)] = __ldg(scan_index + pos + 1);
            qn += __popc(b1);
            __syncwarp();
            const bool flush_all = __any_sync(0xffffffffu, urgent);
            if (qn >= 32 || flush_all) flush(detx, planes, queue, lane, flush_all, D);
        }
    }

    // The scan of a segment is two loops (the kernel below): rows are polled by row_general until the exact max-votes
    // reaches 6 -- it never falls again -- and by row_six from then on, so that the hot loop carries nothing of the
    // general phase.  In both, c0 / c1 hold this lane's pair of the current row on entry and of the next row (if any)
    // on return; `Dh` carries the hot constants only (rays l, m, r, the bottom-face targets, the margin scale).
    template <bool kStep>
    __device__ __forceinline__ void row_general(const DetConst &Dh, const RowSource &src, ulonglong2 &c0, ulonglong2 &c1,
                                                const int N, const int lane, int *queue, const float *detx,
                                                const float4 *__restrict__ planes, const int32_t *__restrict__ scan_index) {
        DetConst D = Dh;
        PairResult h;
        bool trig0, trig1, urgent = false;
        Bottom g;
        eval_dots(D, from_u64(c0.x), from_u64(c0.y), from_u64(c1.x), from_u64(c1.y), g);
        if (src.more()) src.load_next<kStep>(c0, c1);
        if (!general_row(D, src, g, h, trig0, trig1, urgent, detx)) return;
        enqueue(trig0, trig1, urgent, src, N, lane, queue, detx, planes, scan_index, D);
    }

    template <bool kStep>
    __device__ __forceinline__ void row_six(const DetConst &Dh, const RowSource &src, ulonglong2 &c0, ulonglong2 &c1,
                                            const int N, const int lane, int *queue, const float *detx,
                                            const float4 *__restrict__ planes, const int32_t *__restrict__ scan_index) {
        DetConst D = Dh;
        PairResult h;
        bool trig0, trig1;
        Bottom g;
        eval_dots(D, from_u64(c0.x), from_u64(c0.y), from_u64(c1.x), from_u64(c1.y), g);
        if (src.more()) src.load_next<kStep>(c0, c1);
        // stage 1: the bottom face only
        if (src.more()) src.load_next<kStep>(c0, c1);
        GPP_STAT3(0, 1);
        eval_bottom_rest<true, true>(D, g);
        h.r[1] = sub2(PackFast::sqrt(g.na), bc(D.td[1]));
        h.r[2] = sub2(PackFast::sqrt(g.nb), bc(D.td[2]));
        h.r[3] = sub2(PackFast::sqrt(g.nc), bc(D.td[3]));
        const f2 S3 = add2(add2(abs2(h.r[1]), abs2(h.r[2])), abs2(h.r[3]));
        const f2 Slo = fma2(neg2(g.w), bc(D.ms), S3);
        if (!__any_sync(0xffffffffu, !(lo(Slo) > wthr) || !(hi(Slo) > wthr))) return;
#if !GPP_STAGE2
        // experiment (measured r02: 1595 exact verifications per detection instead of 79, 3.8e11 instead of 4.8e11
        // hypotheses/s -- the second stage is what keeps the exact path rare)
        GPP_STAT3(1, 1);
        trig0 = !(lo(Slo) > wthr);
        trig1 = !(hi(Slo) > wthr);
#else
        // stage 2: X_t, the height and the two slanted edges, the full margin
        GPP_STAT3(1, 1);
        load_cold_a(D, detx);
        ulonglong2 v0, v1;
        src.load_again(v0, v1);
        const f2 n0 = from_u64(v0.x), n1 = from_u64(v0.y), n2 = from_u64(v1.x), d4 = from_u64(v1.y);
        f2 ne, nf;
        eval_top<2, GPP_LAZY_Z6>(D, n0, n1, n2, d4, g, h, ne, nf, ColdB{detx});
        h.r[4] = sub2(PackFast::sqrt(abs2(ne)), bc(D.td[4]));
        h.r[5] = sub2(PackFast::sqrt(abs2(nf)), bc(D.td[5]));
        const f2 R = add2(add2(add2(S3, abs2(h.r[0])), abs2(h.r[4])), abs2(h.r[5]));
        const f2 Rlo = sub2(R, h.m);
        trig0 = !(lo(Rlo) > wthr);
        trig1 = !(hi(Rlo) > wthr);
        if (!__any_sync(0xffffffffu, trig0 || trig1)) return;
        h.m = add2(h.m, bc(D.mc));
        finalize_margin(h, R, D);
        const f2 rm = pk(rmax_of(lo(h.r[0]), lo(h.r[1]), lo(h.r[2]), lo(h.r[3]), lo(h.r[4]), lo(h.r[5])),
                         rmax_of(hi(h.r[0]), hi(h.r[1]), hi(h.r[2]), hi(h.r[3]), hi(h.r[4]), hi(h.r[5])));
        const f2 rlo = sub2(rm, h.m);               // lower bound of max |r_k|
        trig0 = trig0 && !(lo(rlo) > 0.7f);
        trig1 = trig1 && !(hi(rlo) > 0.7f);
        if (!__any_sync(0xffffffffu, trig0 || trig1)) return;
        if (GPP_LAZY_Z6) h.finish_zc();
        const f2 zhi = z_upper(h, D);               // upper bound of z_dir_check
        const f2 Rl2 = sub2(R, h.m);                // lower bound of the residual sum
        {
            // a plane that CERTAINLY has six votes and passes the z-check bounds the final best residual
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
#endif
        enqueue(trig0, trig1, false, src, N, lane, queue, detx, planes, scan_index, D);
    }

    // end of the segment: the last partial batch; afterwards `st` holds the exact result of the scanned planes
    __device__ __forceinline__ void finish(const float *detx, const float4 *__restrict__ planes, const int *queue,
                                           int lane) {
        if (qn > 0) {
            GPP_STAT3(3, qn);
            const LaneSel r = verify_queue(detx, planes, queue, 0, qn, lane, st.M, st.bestR, st.bestIdx);
            st.M = r.M; st.bestR = r.bestR; st.bestIdx = r.bestIdx;
            qn = 0;
            __syncwarp();
        }
#ifdef GPP_STATS
        if (lane == 0)
            for (int i = 0; i < 10; ++i) atomicAdd(&g_stats3[i], (unsigned long long)stat[i]);
#endif
    }
    __device__ __forceinline__ void result(int &Mw, float &rbest, int &idx) const {
        Mw = __reduce_max_sync(0xffffffffu, st.M);
        rbest = (st.M == Mw) ? st.bestR : FLT_MAX;
        idx = st.bestIdx;
    }
};

// ------------------------------------------------------------------ FAST: the search in the fast arithmetic alone
struct FastScan {
    LaneState<float> st;   // general mode (max votes not yet known to be 6)
    LaneBest b6;           // once a plane with six votes has been seen
    bool m6;
    float wbest;

    __device__ __forceinline__ void begin() {
        st.reset(FLT_MAX);
        b6.bestR = FLT_MAX;
        b6.bestIdx = 0;
        m6 = false;
        wbest = FLT_MAX;
    }
    // The state holds POSITIONS of the scan order; equal residuals (duplicate planes) are settled by the plane index.
    static __device__ __forceinline__ bool index_below(const int32_t *__restrict__ scan_index, int pos, int other) {
        return __ldg(scan_index + pos) < __ldg(scan_index + other);
    }
    template <bool kStep>
    __device__ __forceinline__ void row(const DetConst &D, const RowSource &src, ulonglong2 &c0, ulonglong2 &c1,
                                        const int32_t *__restrict__ scan_index) {
        PairResult h;
        const int j = src.position();
        Bottom g;
        eval_dots(D, from_u64(c0.x), from_u64(c0.y), from_u64(c1.x), from_u64(c1.y), g);
        if (src.more()) src.load_next<kStep>(c0, c1);
        if (!m6) {
            eval_bottom_rest<false, false>(D, g);
            {
                ulonglong2 v0, v1;
                src.load_again(v0, v1);
                f2 ne, nf;
                eval_top<0, false>(D, from_u64(v0.x), from_u64(v0.y), from_u64(v1.x), from_u64(v1.y), g, h, ne, nf);
                h.r[1] = sub2(PackFast::sqrt(g.na), bc(D.td[1]));
                h.r[2] = sub2(PackFast::sqrt(g.nb), bc(D.td[2]));
                h.r[3] = sub2(PackFast::sqrt(g.nc), bc(D.td[3]));
                h.r[4] = sub2(PackFast::sqrt(abs2(ne)), bc(D.td[4]));
                h.r[5] = sub2(PackFast::sqrt(abs2(nf)), bc(D.td[5]));
            }
            const f2 R = resid_sum(h);
            const int V0 = votes_of(lo(h.r[0]), lo(h.r[1]), lo(h.r[2]), lo(h.r[3]), lo(h.r[4]), lo(h.r[5]));
            const int V1 = votes_of(hi(h.r[0]), hi(h.r[1]), hi(h.r[2]), hi(h.r[3]), hi(h.r[4]), hi(h.r[5]));
            update_general(V0, lo(R), lo(h.zc) < 0.0f, j, scan_index);
            update_general(V1, hi(R), hi(h.zc) < 0.0f, j + 1, scan_index);
            if ((src.rows_left & 3) == 0 && __reduce_max_sync(0xffffffffu, st.M) == 6) {
                m6 = true;                               // warp-uniform; candidates under a lower max are masked
                b6.bestR = (st.M == 6) ? st.bestR : FLT_MAX;
                b6.bestIdx = st.bestIdx;
                wbest = __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(b6.bestR)));
            }
            return;
        }
        // two stages like the VERIFIED all-six phase; the bottom-face sum never exceeds the full sum (rounded
        // addition of non-negative terms is monotone), so leaving here decides what the full test would decide
        eval_bottom_rest<true, false>(D, g);
        h.r[1] = sub2(PackFast::sqrt(g.na), bc(D.td[1]));
        h.r[2] = sub2(PackFast::sqrt(g.nb), bc(D.td[2]));
        h.r[3] = sub2(PackFast::sqrt(g.nc), bc(D.td[3]));
        {
            const f2 S3 = add2(add2(abs2(h.r[1]), abs2(h.r[2])), abs2(h.r[3]));
            if (!__any_sync(0xffffffffu, !(lo(S3) > wbest) || !(hi(S3) > wbest))) return;
        }
        ulonglong2 v0, v1;
        src.load_again(v0, v1);
        const f2 n0 = from_u64(v0.x), n1 = from_u64(v0.y), n2 = from_u64(v1.x), d4 = from_u64(v1.y);
        f2 ne, nf;
        eval_top<0, true>(D, n0, n1, n2, d4, g, h, ne, nf);
        h.r[4] = sub2(PackFast::sqrt(abs2(ne)), bc(D.td[4]));
        h.r[5] = sub2(PackFast::sqrt(abs2(nf)), bc(D.td[5]));
        const f2 R = resid_sum(h);
        if (__any_sync(0xffffffffu, !(lo(R) > wbest) || !(hi(R) > wbest))) {
            h.finish_zc();
            update_six(rmax_of(lo(h.r[0]), lo(h.r[1]), lo(h.r[2]), lo(h.r[3]), lo(h.r[4]), lo(h.r[5])), lo(h.zc), lo(R), j, scan_index);
            update_six(rmax_of(hi(h.r[0]), hi(h.r[1]), hi(h.r[2]), hi(h.r[3]), hi(h.r[4]), hi(h.r[5])), hi(h.zc), hi(R), j + 1, scan_index);
            wbest = __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(b6.bestR)));
        }
    }
    // LaneState::update / LaneBest::update6 for a scan that does not visit the planes in index order
    __device__ __forceinline__ void update_general(int V, float R, bool zneg, int pos, const int32_t *__restrict__ scan_index) {
        const bool grow = V > st.M;
        const float cur = grow ? FLT_MAX : st.bestR;
        bool better = (V >= st.M) && !zneg && (R < cur);
        if ((V >= st.M) && !zneg && R == cur && R < FLT_MAX) better = index_below(scan_index, pos, st.bestIdx);
        st.bestR = better ? R : cur;
        st.bestIdx = better ? pos : st.bestIdx;
        st.M = max(st.M, V);
    }
    __device__ __forceinline__ void update_six(float rmax, float zc, float R, int pos, const int32_t *__restrict__ scan_index) {
        const bool ok = !(rmax > 0.7f) && !(zc < 0.0f);
        bool better = ok && (R < b6.bestR);
        if (ok && R == b6.bestR && R < FLT_MAX) better = index_below(scan_index, pos, b6.bestIdx);
        b6.bestR = better ? R : b6.bestR;
        b6.bestIdx = better ? pos : b6.bestIdx;
    }
    // the lane's best as (max-votes, residual, PLANE INDEX)
    __device__ __forceinline__ void result(int &Mw, float &rbest, int &idx, const int32_t *__restrict__ scan_index) const {
        if (m6) {
            Mw = 6;
            rbest = b6.bestR;
            idx = __ldg(scan_index + b6.bestIdx);
        } else {
            Mw = __reduce_max_sync(0xffffffffu, st.M);
            rbest = (st.M == Mw) ? st.bestR : FLT_MAX;
            idx = __ldg(scan_index + st.bestIdx);
        }
    }
};

// rows m and n of the detection arrays hold the same bits (boxes 12 + dimensions 3 + orientation 1 words)
__device__ __forceinline__ bool same_detection(const PollArgs3 &a, long long m, long long n, int lane) {
    unsigned x = 0, y = 0;
    if (lane < 12) {
        x = __ldg(reinterpret_cast<const unsigned *>(a.boxes) + 12 * m + lane);
        y = __ldg(reinterpret_cast<const unsigned *>(a.boxes) + 12 * n + lane);
    } else if (lane < 15) {
        x = __ldg(reinterpret_cast<const unsigned *>(a.dims) + 3 * m + (lane - 12));
        y = __ldg(reinterpret_cast<const unsigned *>(a.dims) + 3 * n + (lane - 12));
    } else if (lane == 15) {
        x = (unsigned)__ldg(a.orient + m);
        y = (unsigned)__ldg(a.orient + n);
    }
    return __all_sync(0xffffffffu, x == y);
}

constexpr int kSharedConsts = 40;        // floats of a detection's constant block: [0, 32) the warp's slot, [32, 39) the hot ones
constexpr int kWarpSmem3 = kVerifyQueue * (int)sizeof(int) + kSharedConsts * (int)sizeof(float);   // queue + constants
__host__ __device__ constexpr size_t smem3_bytes(int warps, int resident_rows) {
    return size_t(resident_rows) * 1024 + size_t(warps) * kWarpSmem3 + 16;
}

// ------------------------------------------------------------------ EXACT / F64: scalar scan, one plane per lane
// The hypothesis comes in steps (gpp_math.cuh).  Once the warp's max-votes M is 5 or 6, a plane matters only if it can
// have more votes than M -- impossible at 6; at 5 all three bottom-face residuals would have to vote -- or as many and a
// residual sum that does not exceed the warp's best at M.  No partial sum of the residuals exceeds the full sum (rounded
// addition of non-negative terms is monotone; a NaN / inf sum never wins, a NaN residual votes and keeps the plane), so
// the hypothesis is dropped as soon as all 32 lanes are out: r3 > best after the points l and r (two of the four
// divisions, one of the six square roots; the diagonal is the longest edge), (r1 + r2) + r3 > best after the point m.
// Nothing that is kept changes by a bit.  The planes come in scan order (gpp_order.cu: rows of similar planes leave
// together, a sample of the whole database first); `index` maps a position back to the plane index, which settles equal
// residuals (null: index order, the F64 database).  A segment takes the rows of 32 planes first, first + step, ...
template <class T>
__device__ __forceinline__ void update_unordered(LaneState<T> &st, int V, T R, bool zneg, int j, T highest) {
    const bool grow = V > st.M;
    const T cur = grow ? highest : st.bestR;
    const bool better = (V >= st.M) && !zneg && (R < cur || (R == cur && R < highest && j < st.bestIdx));
    st.bestR = better ? R : cur;
    st.bestIdx = better ? j : st.bestIdx;
    st.M = max(st.M, V);
}

template <class P, class Q>
__device__ __forceinline__ void scalar_scan(const Detection<P> &det, const typename P::T4 *__restrict__ planes,
                                            const int32_t *__restrict__ index, const int first_row, const int row_step,
                                            const int n_planes, const int lane, const bool same_rays,
                                            LaneState<typename P::T> &st) {
    typedef typename P::T T;
    typedef typename P::T4 T4;
    // Q: P, or the same arithmetic with independent divisions / square roots paired (the EXACT mode's own scan)
    const Detection<Q> &dq = reinterpret_cast<const Detection<Q> &>(det);      // same layout
    const T highest = P::highest();
    const T thr = P::thresh();
    st.reset(highest);
    if (same_rays) {
        // FilterDetections' padding rows: the cheap form that identical rays allow (bit-identical, gpp_math.cuh).  (A
        // padded image has ONE such row, and its scan is a chain of dependent arithmetic on a nearly empty machine: 19 us of
        // a 50 us single-image call.  Fetching the plane a step ahead, unrolling by four, or moving the loop out of line
        // each cost the packed scan loops of the same kernel their spill-free register allocation.)
#pragma unroll 2
        for (int p = (first_row << 5) + lane; p < n_planes; p += row_step << 5) {
            const T4 pl = planes[p];
            T X[4][3];
            int V; T R; bool z;
            hypothesis_same_rays<Q>(dq, pl.x, pl.y, pl.z, pl.w, X, V, R, z);
            update_unordered(st, V, R, z, index ? __ldg(index + p) : p, highest);
        }
        return;
    }
    int Mw = -1;                    // warp-uniform: max-votes so far (refreshed every fourth row while below 6)
    T wbest = highest;              // best residual at Mw so far
    int it = 0;
#pragma unroll 1
    for (int p0 = first_row << 5; p0 < n_planes; p0 += row_step << 5, ++it) {
        const int p = p0 + lane;
        const bool valid = p < n_planes;                   // lanes past the end poll the last plane again and drop it
        const T4 pl = planes[valid ? p : n_planes - 1];
        T X[4][3];
        int V; T R; bool zneg;
        if (Mw >= (sizeof(T) == 8 ? 6 : 5)) {          // (the fp64 scan, in index order, is faster staged from 6 only)
            T rb[3];
            rb[2] = bottom_lr<Q>(dq, pl.x, pl.y, pl.z, pl.w, X);
            if (!__any_sync(0xffffffffu, valid && (!(rb[2] > wbest) || (Mw == 5 && !(rb[2] > thr))))) continue;
            bottom_m<Q>(dq, pl.x, pl.y, pl.z, pl.w, X, rb);
            const T S3 = P::add(P::add(rb[0], rb[1]), rb[2]);
            const bool all3 = !(rb[0] > thr) && !(rb[1] > thr) && !(rb[2] > thr);
            if (!__any_sync(0xffffffffu, valid && (!(S3 > wbest) || (Mw == 5 && all3)))) continue;
            zneg = bottom_zneg<Q>(X);
            hypothesis_top<Q>(dq, pl.x, pl.y, pl.z, X, rb, V, R);
            if (valid) update_unordered(st, V, R, zneg, index ? __ldg(index + p) : p, highest);
            Mw = __reduce_max_sync(0xffffffffu, st.M);
            wbest = warp_min_value(st.M == Mw ? st.bestR : highest);
        } else {
            hypothesis<Q>(dq, pl.x, pl.y, pl.z, pl.w, X, V, R, zneg);
            if (valid) update_unordered(st, V, R, zneg, index ? __ldg(index + p) : p, highest);
            if ((it & 3) == 3) {
                Mw = __reduce_max_sync(0xffffffffu, st.M);
                wbest = warp_min_value(st.M == Mw ? st.bestR : highest);
            }
        }
    }
}

// `run` consecutive rows of kWords words each, all equal to the words that lanes first_lane .. first_lane + kWords - 1 hold
// in `value`: written as one coalesced stream
template <int kWords, class V>
__device__ __forceinline__ void emit_rows(V *dst, long long run, V value, int first_lane, int lane) {
    const long long total = kWords * run;
    for (long long base = 0; base < total; base += 32) {
        const long long i = base + lane;
        const V v = __shfl_sync(0xffffffffu, value, first_lane + int(i % kWords));
        if (i < total) dst[i] = v;
    }
}

template <int kMode> struct ModePolicy { typedef ExactF32 type; };
template <> struct ModePolicy<kModeF64> { typedef ExactF64 type; };

// kSeg: detections are cut into plane segments (small batches); the large-batch instantiation carries none of it.
// kPose: the epilogue also runs pose recovery and the KITTI record (their double-precision code and stack frame stay
// out of the plain instantiations).
template <int kWarps, int kMode, bool kSeg, bool kPose>
__global__ void __launch_bounds__(kWarps * 32, 1) poll3_kernel(const PollArgs3 args) {
    constexpr bool kVerified = kMode == kModeVerified;
    constexpr bool kPacked = kMode == kModeFast || kMode == kModeVerified;
    typedef typename ModePolicy<kMode>::type P;       // the exact policy of the output type
    typedef typename P::T T;
    typedef typename P::T4 T4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int N = args.n_planes;
    const int NR = args.n_pairs_padded >> 5;                  // rows of 32 pairs = 64 planes
    const int res = kPacked ? args.resident_rows : 0;
    unsigned char *warp_area = smem_raw + size_t(res) * 1024 + size_t(warp) * kWarpSmem3;
    int *queue = reinterpret_cast<int *>(warp_area);
    float *detx = reinterpret_cast<float *>(warp_area + kVerifyQueue * sizeof(int));
    uint64_t *stage_bar = reinterpret_cast<uint64_t *>(smem_raw + size_t(res) * 1024 + size_t(kWarps) * kWarpSmem3);
    const T4 *planesT = kMode == kModeF64 ? reinterpret_cast<const T4 *>(args.planes64) : reinterpret_cast<const T4 *>(args.planes);

    // ---- stage the resident rows: 1-D TMA bulk copies, one mbarrier for the lot
    if (res > 0) {
        if (threadIdx.x == 0) {
            mbar_init(stage_bar, 1);
            mbar_fence_init();
            const uint32_t total = uint32_t(res) * 1024u;
            mbar_arrive_expect_tx(stage_bar, total);
            for (uint32_t off = 0; off < total; off += 32768u) {
                const uint32_t bytes = min(32768u, total - off);
                tma_load_1d(smem_raw + off, reinterpret_cast<const unsigned char *>(args.pairs) + off, bytes, stage_bar);
            }
        }
        __syncthreads();
        mbar_wait(stage_bar, 0);             // every thread: nobody leaves or reads before the copies have landed
    }

    const int n_seg = kSeg ? args.n_seg : 1;
    const int stride = args.det_stride;
    const unsigned long long n_rows = (unsigned long long)((args.n_det + stride - 1) / stride);
    const unsigned long long n_items = n_rows * (unsigned)n_seg;
    // The first item of every warp is fixed (warp w of CTA b: item w * gridDim + b, so that the items of a small call
    // spread over all SMs): a launch does not begin with 4736 atomics on one address.  The counter hands out the rest.
    unsigned long long next_claim = (unsigned long long)(warp * gridDim.x + blockIdx.x);
    next_claim = __shfl_sync(0xffffffffu, next_claim, 0);       // (keeps the compiler from specialising the first round)

    for (;;) {
        const unsigned long long item = next_claim;
        if (item >= n_items) break;                                  // warp-uniform: this warp retires
        if (lane == 0) next_claim = atomicAdd(args.claim, 1ull) + gridDim.x * kWarps;     // claimed one item ahead
        next_claim = __shfl_sync(0xffffffffu, next_claim, 0);
#if GPP_SEG_MAJOR
        // Items in segment-major order (segment 0 of every detection, then segment 1, ...; segmented launches have fewer
        // than 2^32 items): with a few rounds of items per warp, the later segments of a detection start when its first
        // one has finished, adopt its max-votes and bound before their first row and never enter the general phase.
        const int seg = kSeg ? int(unsigned(item) / unsigned(n_rows)) : 0;
        const long long slot = kSeg ? (long long)(unsigned(item) - unsigned(seg) * unsigned(n_rows)) : (long long)item;
#else
        const long long slot = kSeg ? (long long)(item / (unsigned)n_seg) : (long long)item;   // index into the scratch
        const int seg = kSeg ? int(item - (unsigned long long)slot * (unsigned)n_seg) : 0;
#endif
        const long long m = slot * stride;
        // a row that repeats the previous row of its image is written by the warp that polls that row
        if (stride == 1 && (m % args.dets_per_image) != 0 && same_detection(args, m, m - 1, lane)) continue;

        GPP_TL(tl0);
        // ---- per-detection prologue (warp-uniform), exact arithmetic: fit_road_planes.py:66-72, :80-83
        // Segments of one detection share it (packed modes): segment 0 leaves the finished constants in the detection's
        // scratch block, and a later segment that finds them there copies 160 bytes instead of redoing the rays, targets,
        // square roots and divisions -- with six segments per detection that was a tenth of all instructions of C3.  In
        // segment-major order segment 0 ran a round earlier; a segment that comes too early just computes them itself.
        Detection<P> detE;
        bool same_rays;
        DetConst D;
        if constexpr (kPacked && kSeg) {
            __syncwarp();                                                 // the previous item's readers are done
            if (seg > 0 && __ldcg(args.seg_ready + slot) != 0u) {
                __threadfence();                                          // the block was complete before the flag was set
                const float *blk = args.seg_consts + slot * kSharedConsts;
                detx[lane] = __ldcg(blk + lane);
                if (lane < kSharedConsts - 32) detx[32 + lane] = __ldcg(blk + 32 + lane);
            } else {
                load_detection<P, P>(detE, args.boxes + 12 * m, args.dims + 3 * m, __ldg(args.orient + m),
                                     args.pinv + 12 * (m / args.dets_per_image));
                DetConst Dw;
#pragma unroll
                for (int i = 0; i < 6; ++i) Dw.td[i] = detE.td[i];
                fast_constants(Dw, detE);
                if (lane == 0) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        detx[i] = detE.dl[i]; detx[3 + i] = detE.dm[i]; detx[6 + i] = detE.dr[i]; detx[9 + i] = detE.dt[i];
                    }
#pragma unroll
                    for (int i = 0; i < 6; ++i) detx[12 + i] = detE.td[i];
                    store_cold(detx, Dw);
                    detx[18] = same_ground_rays(detE) ? 1.0f : 0.0f;
                    detx[32] = Dw.fl[0]; detx[33] = Dw.fl[1]; detx[34] = Dw.fm[0]; detx[35] = Dw.fm[1];
                    detx[36] = Dw.fr[0]; detx[37] = Dw.fr[1]; detx[38] = Dw.ms;
                }
                if (seg == 0) {
                    __syncwarp();
                    float *blk = args.seg_consts + slot * kSharedConsts;
                    __stcg(blk + lane, detx[lane]);
                    if (lane < kSharedConsts - 32) __stcg(blk + 32 + lane, detx[32 + lane]);
                    __threadfence();
                    __syncwarp();
                    if (lane == 0) __stcg(args.seg_ready + slot, 1u);
                }
            }
            __syncwarp();
            // Every constant comes out of the warp's slot.  The compiler cannot know that a load from an address derived
            // from threadIdx gives the same value in all lanes, and would keep the ten hot constants in (spilled) vector
            // registers: a warp reduction of equal bit patterns returns them in uniform registers, where the scan loops
            // of the unsegmented kernel have them too.
            load_cold(D, detx);
            auto uni = [](float v) { return __uint_as_float(__reduce_or_sync(0xffffffffu, __float_as_uint(v))); };
            D.td[1] = uni(detx[13]); D.td[2] = uni(detx[14]); D.td[3] = uni(detx[15]);
            D.fl[0] = uni(detx[32]); D.fl[1] = uni(detx[33]); D.fm[0] = uni(detx[34]); D.fm[1] = uni(detx[35]);
            D.fr[0] = uni(detx[36]); D.fr[1] = uni(detx[37]); D.ms = uni(detx[38]);
            same_rays = __reduce_or_sync(0xffffffffu, __float_as_uint(detx[18])) != 0u;
        } else {
            load_detection<P, P>(detE, args.boxes + 12 * m, args.dims + 3 * m, __ldg(args.orient + m),
                                 args.pinv + 12 * (m / args.dets_per_image));
            same_rays = same_ground_rays(detE);
            if constexpr (kPacked) {
                // the exact constants are only needed by the rare exact paths and the epilogue: parked in shared memory
#pragma unroll
                for (int i = 0; i < 6; ++i) D.td[i] = detE.td[i];
                fast_constants(D, detE);
                __syncwarp();                            // the previous item's readers are done
                if (lane == 0) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        detx[i] = detE.dl[i]; detx[3 + i] = detE.dm[i]; detx[6 + i] = detE.dr[i]; detx[9 + i] = detE.dt[i];
                    }
#pragma unroll
                    for (int i = 0; i < 6; ++i) detx[12 + i] = detE.td[i];
                    store_cold(detx, D);
                }
                __syncwarp();
                // td[3] comes out of an IEEE square root (a subroutine call the compiler cannot see through): read back
                // from the warp's slot it is a load from a warp-uniform address, and joins the other hot constants in a
                // uniform register instead of being reloaded from local memory on every row
                D.td[3] = detx[15];
            }
        }

        GPP_TL(tl1);
        int Mw, idx;
        T rbest;
        if (!kPacked || (kVerified && same_rays)) {
            // EXACT / F64, and VERIFIED on FilterDetections' padding rows (all their hypotheses are within rounding
            // noise of each other, so the filter cannot drop any): every plane of the segment in the exact arithmetic
            Detection<P> det;
            if constexpr (kPacked) det = load_det_exact(detx); else det = detE;
            LaneState<T> st;
            if constexpr (kMode == kModeF64)
                scalar_scan<P, P>(det, planesT, nullptr, seg, n_seg, N, lane, same_rays, st);
            else if constexpr (kMode == kModeExact)
                scalar_scan<P, typename PairedOf<P>::type>(det, reinterpret_cast<const T4 *>(args.planes_scan),
                                                           args.scan_index, seg, n_seg, N, lane, same_rays, st);
            else
                scalar_scan<P, P>(det, reinterpret_cast<const T4 *>(args.planes_scan), args.scan_index, seg, n_seg, N, lane,
                                  same_rays, st);
            Mw = __reduce_max_sync(0xffffffffu, st.M);
            rbest = (st.M == Mw) ? st.bestR : P::highest();
            idx = st.bestIdx;
#ifdef GPP_STATS
            if (kVerified && lane == 0) { atomicAdd(&g_stats3[5], 1ull); atomicAdd(&g_stats3[6], 1ull); }
#endif
        } else if constexpr (kPacked) {
            // rows seg, seg + n_seg, ... of the pair database (scan order: gpp_order.cu)
            RowSource src;
            src.init(smem_u32(smem_raw), args.pairs, res, seg, n_seg, NR, lane);
            ulonglong2 c0, c1;
            src.load_again(c0, c1);
            float rb;
            if (kVerified) {
                VerifiedScan sc;
                sc.begin();
                if (kSeg && GPP_SEG_MAJOR) sc.adopt(__ldcg(args.seg_best + slot), detx);      // what finished segments found
                // segments of one detection share their bounds: every fourth row a warp publishes its own (one atomicMax
                // on the detection's 64-bit key) and adopts the best published so far (one uniform branch per row; the
                // round-2 profile of C3 showed the per-row form of this exchange at 40 instructions a row)
                // (exchanging on every row of the general phase, whose rows cost three of the others, was measured: C3 0.262 ms
                // against 0.245 ms; so was a form that never waits -- the key read at the top of every row, adopted at its end,
                // published by a reduction without a result: 0.248 ms, no different)
                auto exchange = [&]() {
                    if (kSeg && (src.rows_left & 3) == 1) {
                        unsigned long long key = seg_key(sc.Mcur, sc.wbest);
                        if (lane == 0) key = max(key, atomicMax(args.seg_best + slot, key));
                        sc.adopt(__shfl_sync(0xffffffffu, key, 0), detx);
                    }
                };
                for (; src.rows_left > 0 && sc.Mcur < 6; src.advance<kSeg>()) {
                    sc.row_general<kSeg>(D, src, c0, c1, N, lane, queue, detx, args.planes, args.scan_index);
                    exchange();
                }
                for (; src.rows_left > 0; src.advance<kSeg>()) {
                    sc.row_six<kSeg>(D, src, c0, c1, N, lane, queue, detx, args.planes, args.scan_index);
                    exchange();
                }
                sc.finish(detx, args.planes, queue, lane);
                sc.result(Mw, rb, idx);
            } else {
                FastScan sc;
                sc.begin();
                for (; src.rows_left > 0; src.advance<kSeg>()) sc.row<kSeg>(D, src, c0, c1, args.scan_index);
                sc.result(Mw, rb, idx, args.scan_index);
            }
            rbest = rb;
        }
        rbest = warp_min_first(rbest, idx);
        GPP_TL(tl2);
#ifdef GPP_TIMELINE
        if (lane == 0 && kSeg) {
            const unsigned k = atomicAdd(&g_timeline_n, 1u);
            if (k < 8192) { g_timeline[k][0] = item; g_timeline[k][1] = tl0; g_timeline[k][2] = tl1; g_timeline[k][3] = tl2; g_timeline[k][4] = same_rays; g_timeline[k][5] = 0; }
        }
#endif

        if (kSeg) {
            // ---- hand the partial result in; the warp that completes the detection merges and continues
            unsigned arrived = 0;
            if (lane == 0) {
                __stcg(reinterpret_cast<int4 *>(args.partials + slot * n_seg + seg),
                       make_int4(__double2loint((double)rbest), __double2hiint((double)rbest), Mw, idx));
                __threadfence();
                arrived = atomicAdd(args.seg_arrived + slot, 1u);
            }
            arrived = __shfl_sync(0xffffffffu, arrived, 0);
            if (arrived != unsigned(n_seg - 1)) continue;
            __threadfence();
            int pM = -1, pidx = 0;
            T pr = P::highest();
            if (lane < n_seg) {
                const int4 raw = __ldcg(reinterpret_cast<const int4 *>(args.partials + slot * n_seg + lane));
                pr = (T)__hiloint2double(raw.y, raw.x); pM = raw.z; pidx = raw.w;
            }
            Mw = __reduce_max_sync(0xffffffffu, pM);
            rbest = (pM == Mw) ? pr : P::highest();
            idx = pidx;
            rbest = warp_min_first(rbest, idx);
            if (lane == 0) {                          // leave the scratch as it was found
                args.seg_arrived[slot] = 0u;
                args.seg_best[slot] = 0ull;
                args.seg_ready[slot] = 0u;
            }
        }

        GPP_TL(tl3);
        // ---- epilogue: lazy first-masked search, exact recompute of the winner (fit_road_planes.py:116-137)
        Detection<P> det;
        if constexpr (kPacked) det = load_det_exact(detx); else det = detE;
        const bool have_cand = rbest < P::highest();
        bool sentinel = false;
        if (!(rbest < T(100))) {
            // the constant 100 carried by masked planes may win: find the first masked plane
            int first_masked = -1;
            if constexpr (kMode == kModeFast) {
                for (int p0 = 0; 2 * p0 < N && first_masked < 0; p0 += 32) {
                    const int p = p0 + lane;                         // planes (2p, 2p + 1) in INDEX order
                    const float4 pa = __ldg(args.planes + min(2 * p, N - 1)), pb = __ldg(args.planes + min(2 * p + 1, N - 1));
                    PairResult h;
                    eval_pair<false>(PackFast(), D, pk(pa.x, pb.x), pk(pa.y, pb.y), pk(pa.z, pb.z), pk(pa.w, pb.w), h);
                    const int V0 = votes_of(lo(h.r[0]), lo(h.r[1]), lo(h.r[2]), lo(h.r[3]), lo(h.r[4]), lo(h.r[5]));
                    const int V1 = votes_of(hi(h.r[0]), hi(h.r[1]), hi(h.r[2]), hi(h.r[3]), hi(h.r[4]), hi(h.r[5]));
                    const bool mk0 = (2 * p < N) && ((V0 < Mw) || lo(h.zc) < 0.0f);
                    const bool mk1 = (2 * p + 1 < N) && ((V1 < Mw) || hi(h.zc) < 0.0f);
                    const unsigned b0 = __ballot_sync(0xffffffffu, mk0), b1 = __ballot_sync(0xffffffffu, mk1);
                    if (b0 | b1) {
                        const int f0 = b0 ? 2 * (p0 + __ffs(b0) - 1) : 0x7fffffff;
                        const int f1 = b1 ? 2 * (p0 + __ffs(b1) - 1) + 1 : 0x7fffffff;
                        first_masked = min(f0, f1);
                    }
                }
            } else {
                for (int j0 = 0; j0 < N && first_masked < 0; j0 += 32) {
                    const int j = j0 + lane;
                    bool masked = false;
                    if (j < N) {
                        const T4 pl = planesT[j];
                        T X[4][3];
                        int V; T R; bool zneg;
                        hypothesis<P>(det, pl.x, pl.y, pl.z, pl.w, X, V, R, zneg);
                        masked = (V < Mw) || zneg;
                    }
                    const unsigned b = __ballot_sync(0xffffffffu, masked);
                    if (b) first_masked = j0 + __ffs(b) - 1;
                }
            }
            if (first_masked >= 0) {
                if (!have_cand || T(100) < rbest || (T(100) == rbest && first_masked < idx)) {
                    sentinel = true;
                    idx = first_masked;
                }
            } else if (!have_cand) {
                idx = 0;                                      // nothing compares below `highest`
            }
        }
        // the winner in the exact arithmetic (every lane computes the same values; lane k keeps output word k)
        T word = T(0);
        float pword = 0.0f;
        {
            const T4 pl = planesT[idx];
            T X[4][3];
            int V; T R; bool zneg;
            hypothesis<P>(det, pl.x, pl.y, pl.z, pl.w, X, V, R, zneg);
            const T rr = P::div(sentinel ? T(100) : R, T(6));
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int i = 0; i < 3; ++i) word = (lane == 3 * k + i) ? X[k][i] : word;
            word = (lane == 12) ? pl.x : word;
            word = (lane == 13) ? pl.y : word;
            word = (lane == 14) ? pl.z : word;
            word = (lane == 15) ? pl.w : word;
            word = (lane == 16) ? rr : word;
            if constexpr (kPose) {
                // the two steps after polling (run_network.py:137-247, :297-327) on the key-points just computed; rows
                // whose orientation is no class (padding) keep zeros and their dimensions, like the stand-alone entries
                float kp[12], pose9[9], rec[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int i = 0; i < 3; ++i) kp[3 * k + i] = (float)X[k][i];
                const int o = __ldg(args.orient + m);
#pragma unroll
                for (int i = 0; i < 6; ++i) pose9[i] = 0.0f;
#pragma unroll
                for (int i = 0; i < 3; ++i) pose9[6 + i] = __ldg(args.dims + 3 * m + i);
                if (o >= 0 && o <= 3) pose_from_keypoints(kp, pose9[7], o, pose9);
                if (args.pose_kitti) kitti_record(pose9, rec);
#pragma unroll
                for (int i = 0; i < 9; ++i) pword = (lane == i) ? pose9[i] : pword;
#pragma unroll
                for (int i = 0; i < 4; ++i) pword = (lane == 9 + i) ? rec[i] : pword;
            }
        }
        GPP_TL(tl4);
        // this row and the identical rows that follow it in the image (their claims were skipped).  The length of
        // the run is found 32 rows at a time -- lane i compares row m + 1 + i with its predecessor -- so that a
        // padded image costs a few load latencies, not one per padding row.
        const long long image_end = (m / args.dets_per_image + 1) * (long long)args.dets_per_image;
        long long run_end = m + 1;
        for (long long base = m + 1; stride == 1 && base < image_end; base += 32) {
            const long long n = base + lane;
            bool same = n < image_end;
            if (same) {
                const unsigned *bx = reinterpret_cast<const unsigned *>(args.boxes) + 12 * n;
                const unsigned *dm = reinterpret_cast<const unsigned *>(args.dims) + 3 * n;
                unsigned diff = (unsigned)(__ldg(args.orient + n) ^ __ldg(args.orient + n - 1));
#pragma unroll
                for (int i = 0; i < 12; ++i) diff |= __ldg(bx + i) ^ __ldg(bx + i - 12);     // all loads in flight at once
#pragma unroll
                for (int i = 0; i < 3; ++i) diff |= __ldg(dm + i) ^ __ldg(dm + i - 3);
                same = diff == 0u;
            }
            const unsigned b = __ballot_sync(0xffffffffu, same);
            const int lead = __ffs(~b) - 1;                   // rows of this batch that continue the run (-1: all 32)
            run_end = base + (lead < 0 ? 32 : lead);
            if (lead >= 0) break;
        }
        GPP_TL(tl5);
        T *kp_out = static_cast<T *>(args.keypoints), *kpl_out = static_cast<T *>(args.keyplanes);
        T *res_out = static_cast<T *>(args.residuals);
        if (run_end == m + 1) {
            if (lane < 12) kp_out[12 * m + lane] = word;
            else if (lane < 16) kpl_out[4 * m + (lane - 12)] = word;
            else if (lane == 16) res_out[m] = word;
            else if (lane == 17 && args.best) args.best[m] = idx;
            if constexpr (kPose) {
                if (lane < 3) args.pose_locations[3 * m + lane] = pword;
                else if (lane < 6) args.pose_angles[3 * m + (lane - 3)] = pword;
                else if (lane < 9) args.pose_dimensions[3 * m + (lane - 6)] = pword;
                else if (lane < 13 && args.pose_kitti) args.pose_kitti[4 * m + (lane - 9)] = pword;
            }
        } else {
            // a run of identical rows (a padded image: up to 99 of them): every output array as one coalesced stream, the
            // lanes spread over the rows (row by row -- four partial stores per row -- the 85 padding rows of an image
            // took 13 us, a third of a single-image call)
            const long long run = run_end - m;
            emit_rows<12>(kp_out + 12 * m, run, word, 0, lane);
            emit_rows<4>(kpl_out + 4 * m, run, word, 12, lane);
            emit_rows<1>(res_out + m, run, word, 16, lane);
            if (args.best)
                for (long long i = lane; i < run; i += 32) args.best[m + i] = idx;
            if constexpr (kPose) {
                emit_rows<3>(args.pose_locations + 3 * m, run, pword, 0, lane);
                emit_rows<3>(args.pose_angles + 3 * m, run, pword, 3, lane);
                emit_rows<3>(args.pose_dimensions + 3 * m, run, pword, 6, lane);
                if (args.pose_kitti) emit_rows<4>(args.pose_kitti + 4 * m, run, pword, 9, lane);
            }
        }
#ifdef GPP_TIMELINE
        if (lane == 0) {
            const unsigned k = atomicAdd(&g_timeline_n, 1u);
            if (k < 8192) { g_timeline[k][0] = item; g_timeline[k][1] = tl2; g_timeline[k][2] = tl3; g_timeline[k][3] = gpp_now(); g_timeline[k][4] = tl4; g_timeline[k][5] = tl5 | (1ull << 63); }
        }
#endif
    }

    // ---- the last CTA to leave resets the counters for the next launch
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long left = atomicAdd(args.claim + 1, 1ull);
        if (left == gridDim.x - 1) {
            args.claim[0] = 0ull;
            args.claim[1] = 0ull;
            __threadfence();
        }
    }
}

}  // namespace gpp
