// Frenet lattice planner hot path for B200 (sm_100a): kernels K1..K5 and their C ABI (include/frenet_b200.h).
//
// Nothing here is GEMM-shaped: the path is fp64 polynomial evaluation with store-dominated HBM
// traffic (K2, K3), a candidate x step x obstacle distance test (K4) and a tiny reduction (K5).
// Design rules used throughout (DESIGN.md has the numbers):
//   * one CTA per lattice row (scenario, target speed, horizon): the longitudinal samples, the path
//     frame (point + normal), the row's lateral coefficients and the obstacle table are produced /
//     staged once per CTA in shared memory and reused by the n_d candidates of the row; lattices with short horizons
//     (at most 32 padded samples, the reference's default) take several rows of a scenario per CTA, a warp per row
//     for the longitudinal phase (k2_short), because one such row is too little work for a CTA;
//   * inside a CTA a warp owns a candidate and its lanes own consecutive time steps (two per lane and
//     16-byte accesses for long horizons, one per lane in k2_short), so every global store is a contiguous, sector-aligned run
//     of doubles; the six fields of a candidate sit next to each other ([candidate][6][npad]);
//   * only whole 32-byte sectors are written (padding samples are stored too): a partially written
//     sector costs a DRAM read-modify-write and halves the achieved store bandwidth;
//   * K2, K3 and K4 run as ONE kernel in the pipeline: a sample is evaluated, transformed and
//     collision-tested while it is in registers and written once (48 B instead of 72 B moved);
//   * per-candidate sums (jerk integrals) and maxima are reduced with warp shuffles in a fixed
//     butterfly order: results are deterministic and sign-symmetric, so the reference's exact
//     cost ties between mirror candidates (+d / -d) survive and are broken by lowest index;
//   * collisions: per chunk of steps an fp32, lane-parallel bounding-circle cull over the obstacle
//     table, an fp32 per-lane pre-test on the survivors, and only then the reference's fp64 test -
//     conservative margins make the masks bit-identical to a brute-force fp64 sweep;
//   * the longitudinal polynomial (once per row and step) is sampled in the reference's own power
//     form with host-provided powers of t, because its values feed the next replan's state and sit
//     exactly on the low-speed threshold in converged runs.
// Compile-time knobs (FRENET_*) record measured alternatives; the defaults are the fastest on B200 (profiles/r1_history.md lists
// the alternatives that were measured and removed again).
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "frenet_b200.h"

namespace {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;
#ifndef FRENET_ROW_THREADS
#define FRENET_ROW_THREADS 256
#endif
constexpr int kRowThreads = FRENET_ROW_THREADS;   // CTA size of the per-row kernels (8 warps)
constexpr int kMaxSmem = 227 * 1024;
// Resident CTAs per SM the evaluate kernel is compiled for (register cap 65536 / (256 * n)).  Measured on B200: the
// store-bound variants are fastest at 3 (80 registers, no rematerialisation), the collision variant at 4 (it is
// latency-bound: dependent shuffle / vote / shared-memory chains want more warps).
#ifndef FRENET_K2_MIN_CTAS
#define FRENET_K2_MIN_CTAS 3
#endif
#ifndef FRENET_K2_MIN_CTAS_COLL
#define FRENET_K2_MIN_CTAS_COLL 4
#endif
// the row-table collision variant keeps the collision work out of the sampling loop: same register budget as the collision-free kernels
#ifndef FRENET_K2_MIN_CTAS_TABLE
#define FRENET_K2_MIN_CTAS_TABLE 4
#endif
// warps that share one 32-step chunk of the row table (the listed obstacles are dealt round-robin to them); 0 = as many as it takes
// to give every warp of the CTA one item.  Measured on B200 (profiles/r2_history.md): 1 is fastest - the per-item set-up (bounding box,
// obstacle listing) is not worth duplicating
#ifndef FRENET_TABLE_SPLIT
#define FRENET_TABLE_SPLIT 1
#endif
// 1: the collision variants of the long-horizon kernel reload the candidate's coefficients from shared memory in every iteration instead of
// keeping them in twelve registers across the collision test: no spills at the 64-register cap, fused kernel 6.2 -> 5.8 ms on B200
#ifndef FRENET_K2_RELOAD_COEF
#define FRENET_K2_RELOAD_COEF 1
#endif
// 1 (default): the snapshot collision stage of the long-horizon pipeline kernel works on ROW level: the n_d candidates of a row share the
// path frame and their lateral polynomials are affine in the lateral target, so every (step, obstacle) pair blocks a contiguous range
// of candidate indices.  The ranges are found once per row (fp32 with explicit error margins), certain hits become bit masks, and only
// candidates in the thin uncertain band are re-tested with the reference's fp64 expression.  0: the round-1 form (bounding-circle cull
// + fp32 pre-test + fp64 test per candidate inside the sampling loop); kept for same-box A/B runs.
#ifndef FRENET_ROW_TABLE
#define FRENET_ROW_TABLE 1
#endif
// EXPERIMENT (profiles/r2_sass_store_path.md): 1 = a candidate's [6][npad] block is staged in shared memory by the warp that owns it and
// leaves the SM as ONE bulk asynchronous copy (cp.async.bulk.global.shared::cta, SASS UBLKCP) issued by one lane, instead of six
// 16-byte STG per lane and iteration.  Needs 6 * n_pad_max doubles of staging per warp, which costs a resident CTA per SM.
#ifndef FRENET_BULK_STORE
#define FRENET_BULK_STORE 0
#endif
// 1: unit normal of the path as the normalised, rotated tangent; 0: through atan2 + sincos like the reference's code.
#ifndef FRENET_FRAME_NORMALISE
#define FRENET_FRAME_NORMALISE 0
#endif

thread_local char g_err[256] = "";

int fail(int code, const char* what, cudaError_t e = cudaSuccess) {
    if (e != cudaSuccess)
        snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    else
        snprintf(g_err, sizeof(g_err), "%s", what);
    return code;
}

// Debug build only (-DFRENET_PHASE_TIMING=1): the latest arrival of any CTA's thread 0 at each phase mark of the one-launch
// replan, in globaltimer nanoseconds (mark 0: the earliest) - scripts/phase_times.py reads them through frenet_debug_phase_times.
// 1: after the table's resolve passes only the candidates with an uncertain obstacle go through the exact-test pass (compacted
// list); 0: all of them walk through it (A/B, profiles/r2_history.md)
#ifndef FRENET_POST_COMPACT
#define FRENET_POST_COMPACT 1
#endif
#ifndef FRENET_PHASE_TIMING
#define FRENET_PHASE_TIMING 0
#endif
#if FRENET_PHASE_TIMING
__device__ unsigned long long g_phase_ts[16];
#define FRENET_PHASE(i)                                                                     \
    do {                                                                                    \
        if (threadIdx.x == 0) {                                                             \
            unsigned long long t_;                                                          \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                          \
            if ((i) == 0) atomicMin(&g_phase_ts[0], t_); else atomicMax(&g_phase_ts[(i)], t_); \
        }                                                                                   \
    } while (0)
#else
#define FRENET_PHASE(i) do { } while (0)
#endif

#define FRENET_CHECK_LAUNCH(name)                                        \
    do {                                                                 \
        cudaError_t e_ = cudaGetLastError();                             \
        if (e_ != cudaSuccess) return fail(FRENET_ERR_CUDA, name, e_);   \
    } while (0)

// ------------------------------------------------------------------------------------------------
// Polynomials
// ------------------------------------------------------------------------------------------------
// Closed-form solution of the 3x3 system QuinticPolynomial solves with np.linalg.solve
// (lib/trajectory/trajectory_1d.py:15-40).
__device__ __forceinline__ void quintic_bvp(double p0, double dp0, double ddp0, double p1, double dp1, double ddp1, double T,
                                            double a[6]) {
    const double a0 = p0, a1 = dp0, a2 = ddp0 / 2.0;
    const double T2 = T * T, T3 = T2 * T;
    const double b0 = p1 - a0 - a1 * T - a2 * T2;
    const double b1 = dp1 - a1 - 2.0 * a2 * T;
    const double b2 = ddp1 - 2.0 * a2;
    a[0] = a0;
    a[1] = a1;
    a[2] = a2;
    a[3] = (10.0 * b0 - 4.0 * T * b1 + 0.5 * T2 * b2) / T3;
    a[4] = (-15.0 * b0 + 7.0 * T * b1 - T2 * b2) / (T3 * T);
    a[5] = (6.0 * b0 - 3.0 * T * b1 + 0.5 * T2 * b2) / (T3 * T2);
}

// Closed-form solution of the 2x2 system of QuarticPolynomial (trajectory_1d.py:76-97); end acceleration 0.
__device__ __forceinline__ void quartic_bvp(double s0, double ds0, double dds0, double ds1, double T, double a[6]) {
    const double a0 = s0, a1 = ds0, a2 = dds0 / 2.0;
    const double b1 = ds1 - a1 - 2.0 * a2 * T;
    const double b2 = 0.0 - 2.0 * a2;
    const double T2 = T * T;
    a[0] = a0;
    a[1] = a1;
    a[2] = a2;
    a[3] = (3.0 * b1 - T * b2) / (3.0 * T2);
    a[4] = (T * b2 - 2.0 * b1) / (4.0 * T2 * T);
    a[5] = 0.0;
}

// Value and three derivatives (trajectory_1d.py:42-71, 99-130) by repeated synthetic division
// (Horner applied to its own intermediate sums): 14 FMAs + 2 scalings from the six coefficients
// alone, so a polynomial occupies 12 registers instead of the 30 that pre-multiplied derivative
// coefficients would need.
struct Poly {
    double a0, a1, a2, a3, a4, a5;

    __device__ __forceinline__ void load(const double* __restrict__ a) {
        a0 = a[0]; a1 = a[1]; a2 = a[2]; a3 = a[3]; a4 = a[4]; a5 = a[5];
    }
    // from shared memory, pinned in place (the compiler may not hoist it): lets the six coefficients die across code that does not
    // need them instead of occupying twelve registers
    __device__ __forceinline__ void load_shared_here(const double* a) {
        const unsigned sa = (unsigned)__cvta_generic_to_shared(a);
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a0), "=d"(a1) : "r"(sa));
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+16];" : "=d"(a2), "=d"(a3) : "r"(sa));
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+32];" : "=d"(a4), "=d"(a5) : "r"(sa));
    }
    // p(t) alone, the same FMA sequence as eval's v0 (bit-identical)
    __device__ __forceinline__ double value(double t) const {
        return fma(fma(fma(fma(fma(a5, t, a4), t, a3), t, a2), t, a1), t, a0);
    }
    // v0 = p(t), v1 = p'(t), v2 = p''(t), v3 = p'''(t)
    __device__ __forceinline__ void eval(double t, double& v0, double& v1, double& v2, double& v3) const {
        const double b4 = fma(a5, t, a4);
        const double b3 = fma(b4, t, a3);
        const double b2 = fma(b3, t, a2);
        const double b1 = fma(b2, t, a1);
        v0 = fma(b1, t, a0);
        const double c4 = fma(a5, t, b4);
        const double c3 = fma(c4, t, b3);
        const double c2 = fma(c3, t, b2);
        v1 = fma(c2, t, b1);
        const double d4 = fma(a5, t, c4);
        const double d3 = fma(d4, t, c3);
        v2 = 2.0 * fma(d3, t, c2);
        const double e4 = fma(a5, t, d4);
        v3 = 6.0 * fma(e4, t, d3);
    }
};

// The reference's own evaluation order (trajectory_1d.py:46,54,63,70): power form, every product and sum rounded
// separately, powers of t taken from the host table (C pow, like the reference's `**`).  Used for the longitudinal
// polynomial only (once per row and step): its samples become the next replan's initial state, and in converged
// runs the sampled speed sits exactly on low_speed_threshold, so how it rounds decides which lateral mode the next
// replan uses.  tp = {t^2, t^3, t^4, t^5}.
__device__ __forceinline__ void eval_power_form(const Poly& p, double t, const double* __restrict__ tp, double& v0, double& v1,
                                                double& v2, double& v3) {
    const double t2 = tp[0], t3 = tp[1], t4 = tp[2], t5 = tp[3];
    auto mul = [](double a, double b) { return __dmul_rn(a, b); };
    auto add = [](double a, double b) { return __dadd_rn(a, b); };
    v0 = add(add(add(add(add(p.a0, mul(p.a1, t)), mul(p.a2, t2)), mul(p.a3, t3)), mul(p.a4, t4)), mul(p.a5, t5));
    v1 = add(add(add(add(p.a1, mul(mul(2.0, p.a2), t)), mul(mul(3.0, p.a3), t2)), mul(mul(4.0, p.a4), t3)), mul(mul(5.0, p.a5), t4));
    v2 = add(add(add(mul(2.0, p.a2), mul(mul(6.0, p.a3), t)), mul(mul(12.0, p.a4), t2)), mul(mul(20.0, p.a5), t3));
    v3 = add(add(mul(6.0, p.a3), mul(mul(24.0, p.a4), t)), mul(mul(60.0, p.a5), t2));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

// the same over each aligned group of `lpc` lanes (16 or 8, warp-uniform): several candidates per warp in k2_short
__device__ __forceinline__ double group_sum(double v, int lpc) {
    if (lpc == 16) v += __shfl_xor_sync(kFull, v, 8);
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ double group_max(double v, int lpc) {
    if (lpc == 16) v = fmax(v, __shfl_xor_sync(kFull, v, 8));
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

// Row addressing shared by the per-row kernels (layout: include/frenet_b200.h).
constexpr int kLonFields = 4;    // s, ds, dds, ddds
constexpr int kCandFields = 6;   // d, dd, ddd, dddd, x, y

struct RowAddr {
    int b, r, iv, iT, N, npad;
    int64_t lon_off;    // first element of the row's longitudinal block [4][npad]
    int64_t cand_off;   // first element of the row's first candidate block [6][npad]; candidate id at + id * 6 * npad
    int64_t cand0;      // into per-candidate arrays
    int64_t candF;      // stride between the four cost arrays
};

__device__ __forceinline__ RowAddr row_addr(const FrenetLattice& L, int64_t br) {
    RowAddr A;
    const int R = L.n_v_cap * L.n_t;
    // (row indices are grid indices, < 2^31: 32-bit unsigned divisions - a 64-bit division is ~100 instructions, and every thread
    //  of every row CTA does this)
    const unsigned ubr = (unsigned)br;
    A.b = (int)(ubr / (unsigned)R);
    A.r = (int)(ubr - (unsigned)A.b * (unsigned)R);
    A.iv = (int)((unsigned)A.r / (unsigned)L.n_t);
    A.iT = A.r - A.iv * L.n_t;
    A.N = L.n_steps[A.iT];
    const int off = L.t_offset[A.iT];
    A.npad = L.t_offset[A.iT + 1] - off;
    const int64_t sumpad = L.t_offset[L.n_t];
    const int64_t within = (int64_t)A.iv * sumpad + off;
    A.lon_off = ((int64_t)A.b * L.row_elems + within) * kLonFields;
    A.cand_off = ((int64_t)A.b * L.cand_elems + within * L.n_d) * kCandFields;
    A.cand0 = br * L.n_d;
    A.candF = (int64_t)L.n_scen * R * L.n_d;
    return A;
}

// Per-scenario switches of a batch (both optional): `follow` overrides FrenetParams.follow scenario by scenario (agents that fell
// back to following their leading obstacle next to agents that keep their speed), `scen_mask` restricts a launch to a subset of
// the scenarios (the follow re-plan of the blocked ones); the others' outputs stay untouched.
__device__ __forceinline__ bool follow_of(const FrenetParams& P, const FrenetBuffers& B, int b) { return B.follow ? B.follow[b] != 0 : P.follow != 0; }
__device__ __forceinline__ bool skipped(const FrenetBuffers& B, int b) { return B.scen_mask != nullptr && B.scen_mask[b] == 0; }

// ------------------------------------------------------------------------------------------------
// K1: candidate generation, one thread per lattice cell (scenario, v, T, d)
// ------------------------------------------------------------------------------------------------
// One lattice cell: the row's longitudinal polynomial (written by the row's first cell) and the cell's lateral quintic.
__device__ __forceinline__ void k1_cell(const FrenetLattice& L, const FrenetParams& P, const FrenetBuffers& B, int64_t gid) {
    const int64_t C = (int64_t)L.n_v_cap * L.n_t * L.n_d;
    const int b = (int)(gid / C);
    const int c = (int)(gid % C);
    const int r = c / L.n_d, id = c % L.n_d;
    const int iv = r / L.n_t, iT = r % L.n_t;
    const int64_t br = (int64_t)b * L.n_v_cap * L.n_t + r;
    double* cl = B.coef_lon + br * 6;
    double* cq = B.coef_lat + gid * 6;
    if (skipped(B, b)) return;

    if (iv >= L.n_v[b]) {   // row beyond this scenario's longitudinal axis
        if (id == 0) {
            B.row_flags[br] = 0;
            B.row_S[br] = 0.0;
#pragma unroll
            for (int j = 0; j < 6; ++j) cl[j] = 0.0;
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) cq[j] = 0.0;
        return;
    }
    const double* st = B.state + (int64_t)b * 6;
    const double p0 = st[0], dp0 = st[1], ddp0 = st[2], s0 = st[3], ds0 = st[4], dds0 = st[5];
    const double T = L.axis_t[iT];
    const double v = L.axis_v[(int64_t)b * L.n_v_cap + iv];

    double a[6];
    if (follow_of(P, B, b)) {   // trajectory_planner.py:132-138: quintic to (st + offset, dst, ddst)
        const double* tg = B.target + ((int64_t)b * L.n_t + iT) * 3;
        quintic_bvp(s0, ds0, dds0, tg[0] + v, tg[1], tg[2], T, a);
    } else {          // :147 velocity keeping
        quartic_bvp(s0, ds0, dds0, v, T, a);
    }
    // S: arc length at the last sample (:155; signed in the Optimal variant, trajectory_planner_optimal.py:151)
    Poly pl;
    pl.load(a);
    const int k_last = L.n_steps[iT] - 1;
    double s_last, unused1, unused2, unused3;
    eval_power_form(pl, (double)k_last * P.sample_period, L.t_pow + 4 * k_last, s_last, unused1, unused2, unused3);
    const double S = (P.lowspeed_rule == FRENET_LOWSPEED_OPTIMAL) ? (s_last - s0) : fabs(s_last - s0);
    bool low = false, dropped = false;
    if (P.lowspeed_rule == FRENET_LOWSPEED_V1) {
        low = (ds0 < P.low_speed_threshold) && (S > 0.0);
    } else if (P.lowspeed_rule == FRENET_LOWSPEED_OPTIMAL) {
        low = ds0 < P.low_speed_threshold;
        dropped = low && !(S > P.s_thresh);
    }
    if (id == 0) {
        B.row_flags[br] = FRENET_ROW_EXISTS | (low ? FRENET_ROW_LOWSPEED : 0) | (dropped ? FRENET_ROW_DROPPED : 0);
        B.row_S[br] = S;
#pragma unroll
        for (int j = 0; j < 6; ++j) cl[j] = a[j];
    }
    double q[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (!dropped) quintic_bvp(p0, dp0, ddp0, L.axis_d[id], 0.0, 0.0, low ? S : T, q);   // :160 / :169
#pragma unroll
    for (int j = 0; j < 6; ++j) cq[j] = q[j];
}

__global__ void __launch_bounds__(256) k1_coefficients(FrenetLattice L, FrenetParams P, FrenetBuffers B) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (int64_t)L.n_v_cap * L.n_t * L.n_d * L.n_scen) return;
    k1_cell(L, P, B, gid);
}

// ------------------------------------------------------------------------------------------------
// ONE-LAUNCH replan (FRENET_STAGE_ONE_LAUNCH): a single planner's replan is a chain of launch latencies, not of work - K1, the
// evaluate kernel and K5 + the winner gather each run a few microseconds.  With kMono the evaluate kernels do all of it: every CTA
// first solves the boundary-value systems of its own rows (the K1 cells it is about to sample), and the CTA that finishes LAST
// (a ticket per scenario, taken after a device-wide fence) runs the argmin and the winner gather for the scenario.
// ------------------------------------------------------------------------------------------------
// Inline inputs (frenet_replan_inline): ONE planner's per-replan inputs - a few hundred bytes - travel as a kernel ARGUMENT, so a
// replan needs no host-to-device copy in front of its only kernel.  Every CTA spills them to the buffers' input arrays first
// (identical values: the duplicate stores are harmless), which is where K1 and every later launch read them.
__device__ __forceinline__ void spill_inline(const FrenetInline& in, const FrenetLattice& L, const FrenetBuffers& B) {
    if (!in.present) return;
    double* st = const_cast<double*>(B.state);
    double* sc = const_cast<double*>(B.scal);
    double* av = const_cast<double*>(L.axis_v);
    const int t = threadIdx.x;
    if (t < 6) st[t] = in.state[t];
    if (t < 3) sc[t] = in.scal[t];
    if (t < L.n_v_cap) av[t] = t < FRENET_INLINE_MAX_V ? in.axis_v[t] : 0.0;
    if (t == 0) const_cast<int32_t*>(L.n_v)[0] = in.n_v;
    if (in.present > 1 && B.target) {
        double* tg = const_cast<double*>(B.target);
        for (int i = t; i < L.n_t * 3; i += blockDim.x) tg[i] = in.target[i];
    }
    __syncthreads();
}

// Partial argmin records: each CTA reduces the candidates it produced itself (they are still in its L1 / L2), so the scenario's
// last CTA combines a few hundred records instead of sweeping every candidate (a single K5 CTA over the 9600 candidates of a dense
// lattice takes longer than the evaluate kernel).  The lexicographic (cost, index) minimum is associative: same result as K5.
constexpr int kKeys = 5;   // ctot, cd, cv, ct, masked ctot
struct Best {
    double v;
    int i;
    __device__ __forceinline__ void init() { v = CUDART_INF; i = 0x7fffffff; }
    __device__ __forceinline__ void take(double nv, int ni) {   // lexicographic (cost, index); NaN never wins
        if (nv < v || (nv == v && ni < i)) { v = nv; i = ni; }
    }
};
struct Partial {
    double v[kKeys];
    int i[kKeys];
    int nvalid, nsurv, pad;
};
static_assert(sizeof(Partial) == FRENET_PARTIAL_BYTES, "Partial record size is part of the ABI (buf->partials)");

// Block-wide combination; the result is valid in thread 0.  (Called by every thread of the CTA.)
__device__ __forceinline__ void block_best(Best (&best)[kKeys], int& nvalid, int& nsurv) {
    __shared__ double sv[kKeys][kRowThreads / 32];
    __shared__ int si[kKeys][kRowThreads / 32];
    __shared__ int sc[2][kRowThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();          // (the arrays may still be read from a previous call)
#pragma unroll
    for (int q = 0; q < kKeys; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(kFull, best[q].v, o);
            const int oi = __shfl_xor_sync(kFull, best[q].i, o);
            best[q].take(ov, oi);
        }
        if (lane == 0) { sv[q][warp] = best[q].v; si[q][warp] = best[q].i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        nvalid += __shfl_xor_sync(kFull, nvalid, o);
        nsurv += __shfl_xor_sync(kFull, nsurv, o);
    }
    if (lane == 0) { sc[0][warp] = nvalid; sc[1][warp] = nsurv; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = blockDim.x >> 5;
        for (int w = 1; w < nw; ++w) {
#pragma unroll
            for (int q = 0; q < kKeys; ++q) best[q].take(sv[q][w], si[q][w]);
            nvalid += sc[0][w];
            nsurv += sc[1][w];
        }
    }
}

template <bool kCg>
__device__ __forceinline__ void gather_one(const FrenetLattice& L, const FrenetParams& P, const FrenetBuffers& B, int b, int c, int has_xy,
                                           double* winner_out = nullptr, int32_t* steps_out = nullptr, int tid = -1, int nthreads = 0);
__device__ __forceinline__ int winner_index(const FrenetResult& res, int sel);

// The CTA's own candidates: `count` of them, the i-th at scenario-relative index first + i * stride.  Then the ticket; the holder of
// the scenario's last one combines the partial records, writes the result record and gathers the winner.
// (run > 1: the CTA's candidates come in runs of `run` consecutive indices, `stride` apart - the standalone K4's warp-major split)
__device__ __forceinline__ void mono_tail(const FrenetLattice& L, const FrenetParams& P, const FrenetBuffers& B, int b, int cta, int ctas_per_scen,
                                          int first, int count, int stride, int use_mask, int gather_sel, int has_xy, int run = 1) {
    __shared__ int sh_last, sh_winner, sh_winner2;
    const int C = L.n_v_cap * L.n_t * L.n_d;
    const int64_t base = (int64_t)b * C, candF = (int64_t)L.n_scen * C;
    Best best[kKeys];
#pragma unroll
    for (int q = 0; q < kKeys; ++q) best[q].init();
    int nvalid = 0, nsurv = 0;
    __syncthreads();          // the CTA's per-candidate results are in global memory (written by other threads of this CTA)
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        const int c = run == 1 ? first + i * stride : first + (i / run) * stride + (i % run);
        const int64_t g = base + c;
        const int fl = B.cand_flags[g];
        if (!(fl & FRENET_CAND_VALID)) continue;
        const double cd = B.cost[g], cv = B.cost[candF + g], ct = B.cost[2 * candF + g], ctot = B.cost[3 * candF + g];
        ++nvalid;
        best[0].take(ctot, c);
        best[1].take(cd, c);
        best[2].take(cv, c);
        best[3].take(ct, c);
        bool pass = true;
        if (use_mask & 1) pass = pass && (fl & FRENET_CAND_KINEMATIC_OK);
        if (use_mask & 4) pass = pass && B.coll_free[g];
        if (pass) {
            ++nsurv;
            best[4].take(ctot, c);
        }
    }
    block_best(best, nvalid, nsurv);
    FRENET_PHASE(7);
    Partial* part_rec = reinterpret_cast<Partial*>(B.partials) + (int64_t)b * ctas_per_scen;
    if (threadIdx.x == 0) {
        Partial pr;
#pragma unroll
        for (int q = 0; q < kKeys; ++q) { pr.v[q] = best[q].v; pr.i[q] = best[q].i; }
        pr.nvalid = nvalid; pr.nsurv = nsurv; pr.pad = 0;
        part_rec[cta] = pr;
    }
    __threadfence();          // this thread's stores (samples, costs, the record) are visible device-wide before the ticket is
    __syncthreads();
    if (threadIdx.x == 0) {
        const int t = atomicAdd(B.tickets + b, 1);
        sh_last = t == ctas_per_scen - 1;
        if (sh_last) B.tickets[b] = 0;      // ready for the next launch (stream order)
    }
    __syncthreads();
    FRENET_PHASE(8);
    if (!sh_last) return;
    __threadfence();
#pragma unroll
    for (int q = 0; q < kKeys; ++q) best[q].init();
    nvalid = nsurv = 0;
    for (int i = threadIdx.x; i < ctas_per_scen; i += blockDim.x) {
        const Partial* pr = part_rec + i;
#pragma unroll
        for (int q = 0; q < kKeys; ++q) best[q].take(__ldcg(&pr->v[q]), __ldcg(&pr->i[q]));
        nvalid += __ldcg(&pr->nvalid);
        nsurv += __ldcg(&pr->nsurv);
    }
    block_best(best, nvalid, nsurv);
    FRENET_PHASE(9);
    if (threadIdx.x == 0) {
        FrenetResult res;
        auto idx = [](const Best& x) { return x.i == 0x7fffffff ? -1 : x.i; };
        res.argmin_ctot = idx(best[0]);
        res.argmin_cd = idx(best[1]);
        res.argmin_cv = idx(best[2]);
        res.argmin_ct = idx(best[3]);
        res.argmin_masked = idx(best[4]);
        res.n_valid = nvalid;
        res.n_survivors = nsurv;
        res.status = nvalid == 0 ? FRENET_STATUS_EMPTY : (res.argmin_masked < 0 ? FRENET_STATUS_NO_FEASIBLE : FRENET_STATUS_OK);
        res.min_ctot = best[0].v;
        res.min_masked_ctot = best[4].v;
        res.reserved[0] = res.reserved[1] = 0.0;
        B.result[b] = res;
        sh_winner = gather_sel >= 0 ? winner_index(res, gather_sel) : -1;
        sh_winner2 = res.argmin_masked;
    }
    if (gather_sel >= 0) {
        __syncthreads();
        // With a mask in use the masked pick is gathered as well, into its own block: a replan launched with the obstacles the driver
        // will most likely present to check_paths next leaves BOTH answers behind (the planner's unmasked winner and the driver's
        // masked one).  The two copies are independent: one half of the CTA's warps takes each.
        if (use_mask && B.winner_masked && B.winner_masked_steps) {
            const int half = blockDim.x / 2;
            if ((int)threadIdx.x < half) gather_one<true>(L, P, B, b, sh_winner, has_xy, nullptr, nullptr, threadIdx.x, half);
            else gather_one<true>(L, P, B, b, sh_winner2, has_xy, B.winner_masked, B.winner_masked_steps, threadIdx.x - half, blockDim.x - half);
        } else {
            gather_one<true>(L, P, B, b, sh_winner, has_xy);
        }
    }
    FRENET_PHASE(10);
}

// ------------------------------------------------------------------------------------------------
// Reference paths (used by K3 and the fused K2)
// ------------------------------------------------------------------------------------------------
// Rounded rectangle, lib/trajectory/circle_trajectory.py:31-64.  c = descriptor block:
// l, h, r, dt, period, b[0..6] (branch boundaries evaluated on the host in the reference's order).
__device__ __forceinline__ void circle_pt(const double* __restrict__ c, double t, double& x, double& y) {
    const double l = c[0], h = c[1], r = c[2], period = c[4];
    const double* b = c + 5;
    if (t >= period) t = fmod(t, period);
    if (t >= 0.0 && t < b[0]) {
        x = t; y = 0.0;
    } else if (t >= b[0] && t < b[1]) {
        double sn, cs; sincos((t - l) / r, &sn, &cs);
        x = l + r * sn; y = r - r * cs;
    } else if (t >= b[1] && t < b[2]) {
        x = l + r; y = t - l - 2.0 * r * CUDART_PI / 4.0 + r;
    } else if (t >= b[2] && t < b[3]) {
        double sn, cs; sincos((t - b[2]) / r, &sn, &cs);
        x = l + cs * r; y = r + h + sn * r;
    } else if (t >= b[3] && t < b[4]) {
        x = l - (t - l - r * CUDART_PI - h); y = 2.0 * r + h;
    } else if (t >= b[4] && t < b[5]) {
        double sn, cs; sincos((t - b[4]) / r, &sn, &cs);
        x = -sn * r; y = r + h + r * cs;
    } else if (t >= b[5] && t < b[6]) {
        x = -r; y = r + h - (t - b[5]);
    } else {   // last arc; negative arc lengths land here too (:62-64)
        double sn, cs; sincos((t - b[6]) / r, &sn, &cs);
        x = -r * cs; y = r - r * sn;
    }
}

// Natural cubic spline over chord length, lib/trajectory/spline_trajectory.py:42-73, 146-158.
// tab rows: knot, ax, bx, cx, dx, ay, by, cy, dy.  Outside the knot range the reference returns the
// first / last knot ABSCISSA for value and derivative alike; kept.
__device__ __forceinline__ void spline_eval(const double* __restrict__ tab, int n, double t, double& x, double& y, double& gx,
                                            double& gy) {
    const double first = tab[0], last = tab[(n - 1) * 9];
    if (t < first) { x = y = gx = gy = first; return; }
    if (t >= last) { x = y = gx = gy = last; return; }
    int lo = 0, hi = n - 1;   // bisect_right(knots, t) - 1
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (tab[mid * 9] <= t) lo = mid; else hi = mid;
    }
    const double* q = tab + lo * 9;
    const double dx = t - q[0];
    const double dx2 = dx * dx, dx3 = dx2 * dx;
    x = q[1] + q[2] * dx + q[3] * dx2 + q[4] * dx3;
    y = q[5] + q[6] * dx + q[7] * dx2 + q[8] * dx3;
    gx = q[2] + 2.0 * q[3] * dx + 3.0 * q[4] * dx2;
    gy = q[6] + 2.0 * q[7] * dx + 3.0 * q[8] * dx2;
}

// Lane-centre parabola of the Duckietown runs, a*x**2 + b*x + c in the reference's operation order, no contraction
// (lib/mapper/mapper_obstacles.py:173).
__device__ __forceinline__ double parabola_y(const double* __restrict__ q, double x) {
    return __dadd_rn(__dadd_rn(__dmul_rn(q[0], __dmul_rn(x, x)), __dmul_rn(q[1], x)), q[2]);
}

// Path point and unit normal at arc length s: n = (-sin th, cos th), th = arctan2 of the path's
// first derivative (compute_ortogonal_vect, lib/tests/test_obstacle_planner.py:53-56).  The circle's
// derivative is the reference's forward chord over dt (circle_trajectory.py:66-72).
// lane: optional per-scenario parabola (a, b, c, ds) in global memory that replaces the descriptor's (batched Duckietown agents, each
// with its own lane fit); read element-wise so that the descriptor stays in the kernel's parameter space.
__device__ __forceinline__ void path_frame(int kind, const double* __restrict__ circle, const double* __restrict__ tab, int n_knots,
                                           double s, double& px, double& py, double& nx, double& ny, const double* __restrict__ lane = nullptr) {
    double gx, gy;
    if (kind == FRENET_PATH_CIRCLE) {
        circle_pt(circle, s, px, py);
        double qx, qy;
        const double h = circle[3];
        circle_pt(circle, s + h, qx, qy);
        gx = (qx - px) / h;
        gy = (qy - py) / h;
    } else if (kind == FRENET_PATH_PARABOLA) {
        // Duckietown driver: chord over ds, not divided (lib/tests/test_mapper_planner_obstacles.py:137-142)
        const double q[3] = {lane ? lane[0] : circle[0], lane ? lane[1] : circle[1], lane ? lane[2] : circle[2]};
        px = s;
        py = parabola_y(q, s);
        const double s1 = s + (lane ? lane[3] : circle[3]);
        gx = s1 - s;
        gy = parabola_y(q, s1) - py;
    } else {
        spline_eval(tab, n_knots, s, px, py, gx, gy);
    }
#if FRENET_FRAME_NORMALISE
    // (-sin th, cos th) with th = atan2(gy, gx) is the tangent rotated by 90 degrees and normalised; computing it that
    // way costs ~15 instructions instead of ~250 for atan2 + sincos and agrees to the last bit or two.  Degenerate
    // tangents (zero, non-finite) take the reference's literal route so that atan2(0, 0) = 0 still gives n = (0, 1).
    const double g2 = fma(gx, gx, gy * gy);
    if (g2 > 1e-280 && g2 < 1e280) {
        const double inv = 1.0 / sqrt(g2);
        nx = -gy * inv;
        ny = gx * inv;
        return;
    }
#endif
    const double th = atan2(gy, gx);
    double sn, cs;
    sincos(th, &sn, &cs);
    nx = -sn;
    ny = cs;
}

// Menger curvature of three consecutive samples against kappa_max, squared and division-free:
// kappa = 2 |a x b| / (|a| |b| |a + b|)  >  kappa_max   <=>   4 (a x b)^2  >  kappa_max^2 |a|^2 |b|^2 |a + b|^2.
__device__ __forceinline__ bool menger_exceeds(double x0, double y0, double x1, double y1, double x2, double y2, double k2max) {
    const double ax = x1 - x0, ay = y1 - y0;
    const double bx = x2 - x1, by = y2 - y1;
    const double cx = x2 - x0, cy = y2 - y0;
    const double cr = __dadd_rn(__dmul_rn(ax, by), -__dmul_rn(ay, bx));
    const double lhs = __dmul_rn(4.0, __dmul_rn(cr, cr));
    const double la = __dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay));
    const double lb = __dadd_rn(__dmul_rn(bx, bx), __dmul_rn(by, by));
    const double lc = __dadd_rn(__dmul_rn(cx, cx), __dmul_rn(cy, cy));
    const double rhs = __dmul_rn(k2max, __dmul_rn(__dmul_rn(la, lb), lc));
    return lhs > rhs;
}

// ------------------------------------------------------------------------------------------------
// Collision test of one 32-step chunk of one candidate (shared by K4 and the fused kernel)
// ------------------------------------------------------------------------------------------------
constexpr int kNoBlocker = 0x7fffffff;
constexpr int kCurvOkBit = 0x100;   // staged next to the candidate flags in shared memory (not part of the stored flag byte)

// x, y: this lane's point (lanes beyond the candidate's last step carry a copy of a valid point and valid = false).
// obs: obstacle table in shared memory as (x, y) pairs; obsf: the same rounded to fp32 for the cull; obstacles the
// sensor did not report are parked at +inf.
// best: lowest blocking obstacle index found so far for this candidate (warp-uniform); returns the updated value.
//
// Two stages.  (1) Cull, in fp32 on the FMA/ALU pipes: obstacles are tested lane-parallel against the bounding
// circle of the chunk's 32 points.  It is conservative: the circle radius is rounded up and the reach is padded
// by 2^-18 of the coordinate magnitude, three orders of magnitude above the fp32 rounding error of the
// operands, so no obstacle within the robot radius of any point can be dropped.  (2) The exact reference test
// (lib/tests/test_obstacle_planner.py:48) in fp64 on the few survivors, in ascending index order, so the first
// hit is the chunk's lowest blocker and the masks are bit-identical to a brute-force fp64 sweep.
__device__ __forceinline__ int collide_chunk(double x, double y, bool valid, int nvalid, const double2* __restrict__ obs,
                                             const float2* __restrict__ obsf, int O, float Rf, double radius_sq, int best,
                                             int lane) {
    const int mid = (nvalid - 1) >> 1;
    const float xf = __double2float_rn(x), yf = __double2float_rn(y);
    const float cx = __shfl_sync(kFull, xf, mid), cy = __shfl_sync(kFull, yf, mid);
    const float ex = xf - cx, ey = yf - cy;
    const float r2 = fmaf(ex, ex, ey * ey);
    const float r2max = __uint_as_float(__reduce_max_sync(kFull, __float_as_uint(r2)));   // r2 >= 0: uint order == float order
    const float reach = (Rf + __fsqrt_ru(r2max)) * 1.00001f + (fabsf(cx) + fabsf(cy)) * 3.8147e-6f;
    // Chunks whose extent does not fit fp32 (ill-conditioned low-speed candidates reach |d| ~ 1e80) are not culled at all.
    const float thr = (r2max < 1e30f) ? reach * reach : CUDART_INF_F;
    const int lim = min(O, best);   // only lower indices can still improve `best`
    for (int ob = 0; ob < lim; ob += kWarp) {
        const int o = ob + lane;
        bool near = false;
        if (o < lim) {
            const float2 q = obsf[o];
            const float ux = q.x - cx, uy = q.y - cy;
            near = !(fmaf(ux, ux, uy * uy) > thr);   // NaN counts as near
        }
        unsigned m = __ballot_sync(kFull, near);
        while (m) {
            const int j = __ffs(m) - 1;
            m &= m - 1;
            const double2 q = obs[ob + j];
            const double dx = x - q.x, dy = y - q.y;
            const bool hit = valid && (__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) <= radius_sq);
            if (__any_sync(kFull, hit)) return ob + j;   // ascending order: nothing after it can be lower
        }
    }
    return best;
}

// The same test for a lane that carries TWO consecutive samples (the fused kernel's layout: 64 steps per warp
// iteration, 16-byte stores).  One bounding circle and one cull per 64 steps; a survivor is tested against both
// of the lane's points.
__device__ __forceinline__ int collide_pair(double xa, double ya, double xb, double yb, bool valid_a, bool valid_b, int nvalid,
                                            const double2* __restrict__ obs, const float2* __restrict__ obsf, int O, float Rf,
                                            double radius_sq, int best, int lane) {
    const int mid = (nvalid - 1) >> 2;   // lane holding the middle step
    const float xf = __double2float_rn(xa), yf = __double2float_rn(ya);
    const float xg = __double2float_rn(xb), yg = __double2float_rn(yb);
    const float cx = __shfl_sync(kFull, xf, mid), cy = __shfl_sync(kFull, yf, mid);
    const float ex = xf - cx, ey = yf - cy, fx = xg - cx, fy = yg - cy;
    const float r2 = fmaxf(fmaf(ex, ex, ey * ey), fmaf(fx, fx, fy * fy));
    const float r2max = __uint_as_float(__reduce_max_sync(kFull, __float_as_uint(r2)));
    const float reach = (Rf + __fsqrt_ru(r2max)) * 1.00001f + (fabsf(cx) + fabsf(cy)) * 3.8147e-6f;
    const bool finite = r2max < 1e30f;
    const float thr = finite ? reach * reach : CUDART_INF_F;
    const float rpt = Rf * 1.00001f + (fabsf(cx) + fabsf(cy) + reach) * 3.8147e-6f;   // every point of the chunk has |coord| <= |c| + reach
    const float thr_pt = finite ? rpt * rpt : CUDART_INF_F;
    const int lim = min(O, best);
    for (int ob = 0; ob < lim; ob += kWarp) {
        const int o = ob + lane;
        bool near = false;
        if (o < lim) {
            const float2 q = obsf[o];
            const float ux = q.x - cx, uy = q.y - cy;
            near = !(fmaf(ux, ux, uy * uy) > thr);
        }
        unsigned m = __ballot_sync(kFull, near);
        while (m) {
            const int j = __ffs(m) - 1;
            m &= m - 1;
            // fp32 pre-test per lane (same conservative padding as the cull): most cull survivors touch no point at all,
            // and the fp64 reference test below runs only when some lane may be within the robot radius.
            const float2 qf = obsf[ob + j];
            const float ax = xf - qf.x, ay = yf - qf.y, bx = xg - qf.x, by = yg - qf.y;
            const bool maybe = !(fminf(fmaf(ax, ax, ay * ay), fmaf(bx, bx, by * by)) > thr_pt);   // NaN counts as maybe
            if (!__any_sync(kFull, maybe)) continue;
            const double2 q = obs[ob + j];
            const double dxa = xa - q.x, dya = ya - q.y, dxb = xb - q.x, dyb = yb - q.y;
            const bool hit = (valid_a && (__dadd_rn(__dmul_rn(dxa, dxa), __dmul_rn(dya, dya)) <= radius_sq)) ||
                             (valid_b && (__dadd_rn(__dmul_rn(dxb, dxb), __dmul_rn(dyb, dyb)) <= radius_sq));
            if (__any_sync(kFull, hit)) return ob + j;
        }
    }
    return best;
}

// ------------------------------------------------------------------------------------------------
// EXTENSION: predicted (time-indexed) obstacles.  The reference tests every step of a candidate against the
// obstacles' CURRENT positions; here step k can be tested against the obstacle's predicted position at step k.
// ------------------------------------------------------------------------------------------------
// Per (scenario, chunk of `span` steps, obstacle): bounding circle of the obstacle's predicted positions in that chunk.
// span: steps per chunk (32 for one step per lane, 64 where a lane carries two).
__global__ void __launch_bounds__(128) obstacle_chunk_bounds(const double* __restrict__ obs_t, const uint8_t* __restrict__ active, int O,
                                                             int K, int n_chunks, int span, float4* __restrict__ bounds) {
    const int b = blockIdx.x / n_chunks, chunk = blockIdx.x % n_chunks;   // (scenario on grid.x: grid.y is limited to 65535)
    for (int o = threadIdx.x; o < O; o += blockDim.x) {
        float4 out = make_float4(CUDART_INF_F, CUDART_INF_F, 0.0f, 0.0f);
        if (active == nullptr || active[(int64_t)b * O + o] != 0) {
            const double2* tbl = reinterpret_cast<const double2*>(obs_t) + ((int64_t)b * O + o) * K;
            const int k0 = chunk * span;
            const double2 c = tbl[min(k0 + span / 2, K - 1)];
            double r2 = 0.0;
            for (int k = k0; k < k0 + span; ++k) {
                const double2 q = tbl[min(k, K - 1)];
                const double dx = q.x - c.x, dy = q.y - c.y;
                r2 = fmax(r2, fma(dx, dx, dy * dy));
            }
            // .w: the obstacle's share of every cull / pre-test radius (its own extent, rounded up, plus the fp32 rounding pad of its centre)
            const float cxf = __double2float_rn(c.x), cyf = __double2float_rn(c.y), rad = __fsqrt_ru(__double2float_ru(r2));
            out = make_float4(cxf, cyf, rad, __fadd_ru(__fmul_ru(rad, 1.00001f), __fmul_ru(fabsf(cxf) + fabsf(cyf), 3.8147e-6f)));
        }
        bounds[((int64_t)b * n_chunks + chunk) * O + o] = out;
    }
}

// Collision test of one 32-step chunk against predicted obstacles.  k: this lane's step; tbl: the scenario's predicted table
// [O][K] (global memory, L2 resident: the rows of a scenario share it); cb: this chunk's obstacle bounds in shared memory.
__device__ __forceinline__ int collide_chunk_predicted(double x, double y, bool valid, int nvalid, int k, const double2* __restrict__ tbl,
                                                       int K, const float4* __restrict__ cb, int O, float Rf, double radius_sq,
                                                       int best, int lane) {
    const int mid = (nvalid - 1) >> 1;
    const float xf = __double2float_rn(x), yf = __double2float_rn(y);
    const float cx = __shfl_sync(kFull, xf, mid), cy = __shfl_sync(kFull, yf, mid);
    const float ex = xf - cx, ey = yf - cy;
    const float r2 = fmaf(ex, ex, ey * ey);
    const float r2max = __uint_as_float(__reduce_max_sync(kFull, __float_as_uint(r2)));
    const float reach = (Rf + __fsqrt_ru(r2max)) * 1.00001f + (fabsf(cx) + fabsf(cy)) * 3.8147e-6f;
    const bool wide = !(r2max < 1e30f);
    const float rpt = Rf * 1.00001f + (fabsf(cx) + fabsf(cy) + reach) * 3.8147e-6f;   // every point of the chunk has |coord| <= |c| + reach
    const int lim = min(O, best);
    const int kk = min(k, K - 1);
    for (int ob = 0; ob < lim; ob += kWarp) {
        const int o = ob + lane;
        bool near = false;
        if (o < lim) {
            const float4 q = cb[o];
            const float ux = q.x - cx, uy = q.y - cy;
            const float rr = reach + q.w;
            // obstacles the sensor did not report are parked at +inf by the bounds kernel: never near
            near = (q.x < CUDART_INF_F) && (wide || !(fmaf(ux, ux, uy * uy) > rr * rr));
        }
        unsigned m = __ballot_sync(kFull, near);
        while (m) {
            const int j = __ffs(m) - 1;
            m &= m - 1;
            // fp32 pre-test per lane against the obstacle's bounding circle over this chunk: the predicted position is
            // fetched from the table (global memory) only when some lane may be within the robot radius of it
            const float4 qc = cb[ob + j];
            const float px = xf - qc.x, py = yf - qc.y;
            const float rp = rpt + qc.w;
            if (!__any_sync(kFull, wide || !(fmaf(px, px, py * py) > rp * rp))) continue;
            const double2 q = tbl[(int64_t)(ob + j) * K + kk];
            const double dx = x - q.x, dy = y - q.y;
            const bool hit = valid && (__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) <= radius_sq);
            if (__any_sync(kFull, hit)) return ob + j;
        }
    }
    return best;
}

// The same for a lane that carries two consecutive samples (steps k0, k0 + 1): one cull per 64 steps against bounds
// computed over 64-step chunks; a survivor's two predicted positions are adjacent in the table (one 32-byte read per lane).
__device__ __forceinline__ int collide_pair_predicted(double xa, double ya, double xb, double yb, bool valid_a, bool valid_b, int nvalid,
                                                      int k0, const double2* __restrict__ tbl, int K, const float4* __restrict__ cb, int O,
                                                      float Rf, double radius_sq, int best, int lane) {
    const int mid = (nvalid - 1) >> 2;
    const float xf = __double2float_rn(xa), yf = __double2float_rn(ya);
    const float xg = __double2float_rn(xb), yg = __double2float_rn(yb);
    const float cx = __shfl_sync(kFull, xf, mid), cy = __shfl_sync(kFull, yf, mid);
    const float ex = xf - cx, ey = yf - cy, fx = xg - cx, fy = yg - cy;
    const float r2 = fmaxf(fmaf(ex, ex, ey * ey), fmaf(fx, fx, fy * fy));
    const float r2max = __uint_as_float(__reduce_max_sync(kFull, __float_as_uint(r2)));
    const float reach = (Rf + __fsqrt_ru(r2max)) * 1.00001f + (fabsf(cx) + fabsf(cy)) * 3.8147e-6f;
    const bool wide = !(r2max < 1e30f);
    const float rpt = Rf * 1.00001f + (fabsf(cx) + fabsf(cy) + reach) * 3.8147e-6f;
    const int lim = min(O, best);
    const int ka = min(k0, K - 1), kb = min(k0 + 1, K - 1);
    for (int ob = 0; ob < lim; ob += kWarp) {
        const int o = ob + lane;
        bool near = false;
        if (o < lim) {
            const float4 q = cb[o];
            const float ux = q.x - cx, uy = q.y - cy;
            const float rr = reach + q.w;
            // obstacles the sensor did not report are parked at +inf by the bounds kernel: never near
            near = (q.x < CUDART_INF_F) && (wide || !(fmaf(ux, ux, uy * uy) > rr * rr));
        }
        unsigned m = __ballot_sync(kFull, near);
        while (m) {
            const int j = __ffs(m) - 1;
            m &= m - 1;
            const float4 qc = cb[ob + j];
            const float pax = xf - qc.x, pay = yf - qc.y, pbx = xg - qc.x, pby = yg - qc.y;
            const float rp = rpt + qc.w;
            if (!__any_sync(kFull, wide || !(fminf(fmaf(pax, pax, pay * pay), fmaf(pbx, pbx, pby * pby)) > rp * rp))) continue;
            const double2* row = tbl + (int64_t)(ob + j) * K;
            const double2 qa = row[ka], qb = row[kb];
            const double dxa = xa - qa.x, dya = ya - qa.y, dxb = xb - qb.x, dyb = yb - qb.y;
            const bool hit = (valid_a && (__dadd_rn(__dmul_rn(dxa, dxa), __dmul_rn(dya, dya)) <= radius_sq)) ||
                             (valid_b && (__dadd_rn(__dmul_rn(dxb, dxb), __dmul_rn(dyb, dyb)) <= radius_sq));
            if (__any_sync(kFull, hit)) return ob + j;
        }
    }
    return best;
}

// Several candidates per warp (k2_short): the warp is split into groups of `lpc` lanes (16 or 8), one candidate per group, two
// consecutive samples per lane, the whole horizon in one go.  One bounding circle and one cull for the 64 points of all groups
// (neighbours in the lattice, so their trajectories lie close together); the exact test attributes a hit to the group it occurred in.
// Obstacles are visited in ascending order, so a group's first hit is its lowest blocker; the sweep ends when every group has one.
// base / gm: first lane and lane mask (shifted down) of this lane's group; act: the group carries a candidate.
// Returns this lane's group's lowest blocker (kNoBlocker if none).  mid: a lane holding a point near the middle.
__device__ __forceinline__ int collide_groups(double xa, double ya, double xb, double yb, bool valid_a, bool valid_b, int mid,
                                              const double2* __restrict__ obs, const float2* __restrict__ obsf, int O, float Rf,
                                              double radius_sq, bool act, int base, unsigned gm) {
    const int lane = threadIdx.x & 31;
    const float xf = __double2float_rn(xa), yf = __double2float_rn(ya);
    const float xg = __double2float_rn(xb), yg = __double2float_rn(yb);
    const float cx = __shfl_sync(kFull, xf, mid), cy = __shfl_sync(kFull, yf, mid);
    const float ex = xf - cx, ey = yf - cy, fx = xg - cx, fy = yg - cy;
    const float r2 = fmaxf(fmaf(ex, ex, ey * ey), fmaf(fx, fx, fy * fy));
    const float r2max = __uint_as_float(__reduce_max_sync(kFull, __float_as_uint(r2)));
    const float reach = (Rf + __fsqrt_ru(r2max)) * 1.00001f + (fabsf(cx) + fabsf(cy)) * 3.8147e-6f;
    const bool finite = r2max < 1e30f;
    const float thr = finite ? reach * reach : CUDART_INF_F;
    const float rpt = Rf * 1.00001f + (fabsf(cx) + fabsf(cy) + reach) * 3.8147e-6f;
    const float thr_pt = finite ? rpt * rpt : CUDART_INF_F;
    int mine = act ? kNoBlocker : -1;   // this group's lowest blocker; a group without a candidate never extends the sweep
    for (int ob = 0; ob < O; ob += kWarp) {
        const int o = ob + lane;
        bool near = false;
        if (o < O) {
            const float2 q = obsf[o];
            const float ux = q.x - cx, uy = q.y - cy;
            near = !(fmaf(ux, ux, uy * uy) > thr);
        }
        unsigned m = __ballot_sync(kFull, near);
        while (m) {
            const int j = __ffs(m) - 1;
            m &= m - 1;
            const float2 qf = obsf[ob + j];
            const float ax = xf - qf.x, ay = yf - qf.y, bx = xg - qf.x, by = yg - qf.y;
            const bool maybe = !(fminf(fmaf(ax, ax, ay * ay), fmaf(bx, bx, by * by)) > thr_pt);
            if (!__any_sync(kFull, maybe)) continue;
            const double2 q = obs[ob + j];
            const double dxa = xa - q.x, dya = ya - q.y, dxb = xb - q.x, dyb = yb - q.y;
            const bool hit = (valid_a && (__dadd_rn(__dmul_rn(dxa, dxa), __dmul_rn(dya, dya)) <= radius_sq)) ||
                             (valid_b && (__dadd_rn(__dmul_rn(dxb, dxb), __dmul_rn(dyb, dyb)) <= radius_sq));
            const unsigned hb = __ballot_sync(kFull, hit);
            if (!hb) continue;
            if (((hb >> base) & gm) && mine == kNoBlocker) mine = ob + j;
            if (__all_sync(kFull, mine != kNoBlocker)) return mine;
        }
    }
    return mine;
}

// The same against predicted obstacles: k0 is this lane's first step, cb the bounds of the only chunk (32 steps).
__device__ __forceinline__ int collide_groups_predicted(double xa, double ya, double xb, double yb, bool valid_a, bool valid_b, int mid,
                                                        int k0, const double2* __restrict__ tbl, int K, const float4* __restrict__ cb,
                                                        int O, float Rf, double radius_sq, bool act, int base, unsigned gm) {
    const int lane = threadIdx.x & 31;
    const float xf = __double2float_rn(xa), yf = __double2float_rn(ya);
    const float xg = __double2float_rn(xb), yg = __double2float_rn(yb);
    const float cx = __shfl_sync(kFull, xf, mid), cy = __shfl_sync(kFull, yf, mid);
    const float ex = xf - cx, ey = yf - cy, fx = xg - cx, fy = yg - cy;
    const float r2 = fmaxf(fmaf(ex, ex, ey * ey), fmaf(fx, fx, fy * fy));
    const float r2max = __uint_as_float(__reduce_max_sync(kFull, __float_as_uint(r2)));
    const float reach = (Rf + __fsqrt_ru(r2max)) * 1.00001f + (fabsf(cx) + fabsf(cy)) * 3.8147e-6f;
    const bool wide = !(r2max < 1e30f);
    const float rpt = Rf * 1.00001f + (fabsf(cx) + fabsf(cy) + reach) * 3.8147e-6f;
    const int ka = min(k0, K - 1), kb = min(k0 + 1, K - 1);
    int mine = act ? kNoBlocker : -1;
    for (int ob = 0; ob < O; ob += kWarp) {
        const int o = ob + lane;
        bool near = false;
        if (o < O) {
            const float4 q = cb[o];
            const float ux = q.x - cx, uy = q.y - cy;
            const float rr = reach + q.w;
            near = (q.x < CUDART_INF_F) && (wide || !(fmaf(ux, ux, uy * uy) > rr * rr));
        }
        unsigned m = __ballot_sync(kFull, near);
        while (m) {
            const int j = __ffs(m) - 1;
            m &= m - 1;
            const float4 qc = cb[ob + j];
            const float pax = xf - qc.x, pay = yf - qc.y, pbx = xg - qc.x, pby = yg - qc.y;
            const float rp = rpt + qc.w;
            if (!__any_sync(kFull, wide || !(fminf(fmaf(pax, pax, pay * pay), fmaf(pbx, pbx, pby * pby)) > rp * rp))) continue;
            const double2* row = tbl + (int64_t)(ob + j) * K;
            const double2 qa = row[ka], qb = row[kb];
            const double dxa = xa - qa.x, dya = ya - qa.y, dxb = xb - qb.x, dyb = yb - qb.y;
            const bool hit = (valid_a && (__dadd_rn(__dmul_rn(dxa, dxa), __dmul_rn(dya, dya)) <= radius_sq)) ||
                             (valid_b && (__dadd_rn(__dmul_rn(dxb, dxb), __dmul_rn(dyb, dyb)) <= radius_sq));
            const unsigned hb = __ballot_sync(kFull, hit);
            if (!hb) continue;
            if (((hb >> base) & gm) && mine == kNoBlocker) mine = ob + j;
            if (__all_sync(kFull, mine != kNoBlocker)) return mine;
        }
    }
    return mine;
}

// ------------------------------------------------------------------------------------------------
// Row-level collision table (snapshot obstacles, long-horizon pipeline kernel)
// ------------------------------------------------------------------------------------------------
// The n_d candidates of a row share the path frame (p_k, n_k), and K1's lateral quintic is affine in the lateral target D_i:
//     d_i(k) = A_k + B_k D_i           (B_k = 10 u^3 - 15 u^4 + 6 u^5 >= 0 with u = argument / horizon).
// With b = t_k . (q - p_k) and a = n_k . (q - p_k) (t_k = (n_y, -n_x), |n_k| = 1) the reference's test of candidate i against obstacle q
// at step k (lib/tests/test_obstacle_planner.py:48) reads  b^2 + (a - A_k - B_k D_i)^2 <= R^2 :  for a uniform lateral axis each
// (step, obstacle) pair blocks a contiguous RANGE of candidate indices.  A warp takes 32 consecutive steps (lane = step), lists the
// obstacles inside the oriented bounding box of the chunk's band of candidates, and for every listed obstacle each lane computes two
// index ranges in fp32: OUTER (nothing outside it can hit at this step) and INNER (everything inside it hits), separated by explicit
// error margins (below).  Per (chunk, obstacle) the warp keeps ONE outer interval - the hull of the lanes' outer ranges, conservative -
// and ONE inner interval: the union of the lanes' inner ranges when they form a connected run (consecutive steps whose ranges
// overlap or touch; then the union IS the hull), else none.  Per candidate the lowest obstacle whose inner interval contains it is
// the first blocker, unless a lower obstacle contains it only in its outer interval: those few (candidate, obstacle) pairs - about
// 1 % of the candidates on the config-5 scenes - are re-tested with the reference's fp64 expression on the stored samples.
// Masks and first blockers are therefore bit-identical to a brute-force fp64 sweep (scripts/row_collision_model.py replays the
// classification on the CPU and checks both inclusions; tests/test_gpu_parity.py compares with the standalone K4 and the oracle).
//
// Error budget of one pair (fp32 unit roundoff u = 2^-24): the inputs of b, a and d_i - a carry
//     delta = 8 u (|dx| + |dy| + |A_k| + |B_k| max|D|)  +  2e-12 M + 1e-9      (M = max_k sum_j |a_j| |u_k|^j of the row's polynomials:
// the fp64 affine model itself; rows with non-finite frames, |A| + |B| max|D| > 1e12 or M > 1e15 take the per-candidate sweep), hence
//     | b^2 + (d_i - a)^2 - R^2 |_error <= tau = 4.5 (R + delta) delta + 16 u R^2   near the boundary,
// and the candidate index of a range end carries eps = 16 u ((r + |c|) / B + max|D|) / dD + 1e-4 (approximate sqrt / reciprocal
// are covered by a relative 1e-6 on r and by eps).  Large |B_k| (low-speed rows sampled beyond their horizon, |d| up to 1e12) only
// widen the band in metres, not in candidate indices: the index error scales with delta / (B dD).
constexpr int kTableMaxD = 128;    // candidates per row the packed intervals can index; wider rows take the per-candidate sweep
constexpr int kMaybeCap = 8;       // uncertain obstacles remembered per candidate; more -> per-candidate sweep for that candidate
constexpr float kU32 = 5.9604645e-8f;
constexpr unsigned kIvlEmpty = 0x00ff00ffu;   // lo = 255, hi = 0 for both intervals

__device__ __forceinline__ int float_order(float f) {   // monotone float -> int map (for redux.sync min / max)
    const int i = __float_as_int(f);
    return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float order_float(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }
__device__ __forceinline__ float warp_min_f(float v) { return order_float(__reduce_min_sync(kFull, float_order(v))); }
__device__ __forceinline__ float warp_max_f(float v) { return order_float(__reduce_max_sync(kFull, float_order(v))); }
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

struct RowTable {
    float* mA;                  // [n_pad_max] A_k (fp32)
    float* mB;                  // [n_pad_max] B_k
    unsigned* ivl;              // [chunks][n_obs] per listed obstacle: outer lo | hi << 8 | inner lo << 16 | hi << 24
    unsigned short* id;         // [chunks][n_obs] the listed obstacles of each chunk, ascending
    int* cnt;                   // [chunks]
    int* first;                 // [n_d] lowest certain blocker per candidate (kNoBlocker if none)
    int* n_maybe;               // [n_d] uncertain obstacles below it (> kMaybeCap: overflow)
    unsigned short* maybe_id;   // [n_d][kMaybeCap]
    unsigned short* scratch;    // [warps][n_obs] a warp's private obstacle list
};

// One warp, steps [32 ch, 32 ch + 32) of the row: list the obstacles near the chunk's band, classify every (step, listed obstacle)
// pair, keep the intervals.  `sub` / `nsub`: the listed obstacles are dealt round-robin to the nsub warps working on the same chunk.
__device__ __forceinline__ void row_table_chunk(const RowTable& T, int ch, int sub, int nsub, int N, const double* fpx, const double* fpy,
                                                const double* fnx, const double* fny, const double2* __restrict__ obs, int O, double D0,
                                                double D1, int n_d, float R, float R2f, float cM, unsigned short* list, int lane) {
    const int k = ch * kWarp + lane;
    const bool valid = k < N;
    const int kk = valid ? k : N - 1;
    const double px = fpx[kk], py = fpy[kk];
    const float nxf = __double2float_rn(fnx[kk]), nyf = __double2float_rn(fny[kk]);
    const float A = T.mA[kk], Bk = T.mB[kk];
    const float D0f = __double2float_rn(D0), D1f = __double2float_rn(D1);
    const float Dmaxf = fmaxf(fabsf(D0f), fabsf(D1f)) * 1.000001f;
    const float invd = (float)max(n_d - 1, 1) / (D1f - D0f);
    // oriented bounding box of the band {p_k + n_k d : d between d_first(k) and d_last(k)} in the frame of the chunk's middle step
    const int midl = min(kWarp / 2, N - 1 - ch * kWarp);
    const double cx = __shfl_sync(kFull, px, midl), cy = __shfl_sync(kFull, py, midl);
    const float nmx = __shfl_sync(kFull, nxf, midl), nmy = __shfl_sync(kFull, nyf, midl);
    const float rx = __double2float_rn(px - cx), ry = __double2float_rn(py - cy);
    const float da = fmaf(Bk, D0f, A), db = fmaf(Bk, D1f, A);
    const float e1x = fmaf(nxf, da, rx), e1y = fmaf(nyf, da, ry), e2x = fmaf(nxf, db, rx), e2y = fmaf(nyf, db, ry);
    const float u1 = fmaf(e1x, nmy, -e1y * nmx), v1 = fmaf(e1x, nmx, e1y * nmy);   // tangent (n_y, -n_x) and normal coordinates
    const float u2 = fmaf(e2x, nmy, -e2y * nmx), v2 = fmaf(e2x, nmx, e2y * nmy);
    float umin = warp_min_f(fminf(u1, u2)), umax = warp_max_f(fmaxf(u1, u2));
    float vmin = warp_min_f(fminf(v1, v2)), vmax = warp_max_f(fmaxf(v1, v2));
    const float ext = fmaxf(fmaxf(fabsf(umin), fabsf(umax)), fmaxf(fabsf(vmin), fabsf(vmax)));
    const float m = R * 1.0001f + ext * 1e-5f + 1e-4f;   // robot radius + far more than the fp32 error of the box and of the obstacle's offset
    umin -= m; umax += m; vmin -= m; vmax += m;
    int cnt = 0;
    for (int ob = 0; ob < O; ob += kWarp) {
        const int o = ob + lane;
        bool in = false;
        if (o < O) {
            const double2 q = obs[o];   // obstacles the sensor did not report sit at +inf: every comparison below is false for them
            const float qx = __double2float_rn(q.x - cx), qy = __double2float_rn(q.y - cy);
            const float u = fmaf(qx, nmy, -qy * nmx), v = fmaf(qx, nmx, qy * nmy);
            in = (u >= umin) && (u <= umax) && (v >= vmin) && (v <= vmax);
        }
        const unsigned bal = __ballot_sync(kFull, in);
        if (in) list[cnt + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)o;
        cnt += __popc(bal);
    }
    if (sub == 0 && lane == 0) T.cnt[ch] = cnt;
    __syncwarp();
    const float SAB = fabsf(A) + fabsf(Bk) * Dmaxf;
    const bool flat = n_d == 1 || !(Bk >= 1e-6f);      // all candidates (almost) share one point at this step
    const float inv = flat ? 0.0f : rcp_approx(Bk);
    const float nd1 = (float)(n_d - 1);
    unsigned* ivl = T.ivl + ch * O;
    unsigned short* ids = T.id + ch * O;
    for (int j = sub; j < cnt; j += nsub) {
        const int o = list[j];
        const double2 q = obs[o];
        const float dxf = __double2float_rn(q.x - px), dyf = __double2float_rn(q.y - py);
        const float b = fmaf(dxf, nyf, -dyf * nxf);
        const float a = fmaf(dxf, nxf, dyf * nyf);
        const float delta = fmaf(8.0f * kU32, fabsf(dxf) + fabsf(dyf) + SAB, cM);
        const float tau = fmaf(4.5f * (R + delta), delta, 16.0f * kU32 * R2f);
        const float bb = b * b;
        const float gout = (R2f + tau) - bb;
        const bool rel = valid && (gout >= 0.0f);
        unsigned packed = kIvlEmpty;
        if (__any_sync(kFull, rel)) {
            int lo_m = 255, hi_m = -1, lo_c = 255, hi_c = -1;   // empty ranges
            if (rel) {
                const float gin = (R2f - tau) - bb;
                const float rout = sqrt_approx(gout) * 1.000001f;
                const float rin = gin > 0.0f ? sqrt_approx(gin) * 0.999999f : -1.0f;
                const float c = A - a;   // d_i - a = c + B_k D_i
                if (flat) {
                    const float spread = fabsf(Bk) * Dmaxf;   // |B_k D_i| <= spread
                    if (fmaxf(fabsf(c) - spread, 0.0f) <= rout) { lo_m = 0; hi_m = n_d - 1; }
                    if (fabsf(c) + spread <= rin) { lo_c = 0; hi_c = n_d - 1; }
                } else {
                    const float eps = fmaf(16.0f * kU32 * invd, fmaf(rout + fabsf(c), inv, Dmaxf), 1e-4f);
                    const float x_lo = fmaf((-rout - c), inv, -D0f) * invd, x_hi = fmaf((rout - c), inv, -D0f) * invd;
                    lo_m = (int)ceilf(fminf(fmaxf(x_lo - eps, 0.0f), nd1 + 1.0f));
                    hi_m = (int)floorf(fmaxf(fminf(x_hi + eps, nd1), -1.0f));
                    if (rin >= 0.0f) {
                        const float y_lo = fmaf((-rin - c), inv, -D0f) * invd, y_hi = fmaf((rin - c), inv, -D0f) * invd;
                        lo_c = (int)ceilf(fminf(fmaxf(y_lo + eps, 0.0f), nd1 + 1.0f));
                        hi_c = (int)floorf(fmaxf(fminf(y_hi - eps, nd1), -1.0f));
                    }
                }
                if (lo_m > hi_m) { lo_m = 255; hi_m = -1; }
                if (lo_c > hi_c) { lo_c = 255; hi_c = -1; }
            }
            // outer interval: hull of the lanes' outer ranges
            const int LO = __reduce_min_sync(kFull, lo_m), HI = __reduce_max_sync(kFull, hi_m);
            // inner interval: the union of the lanes' inner ranges if they form a connected run of steps
            const bool ne = lo_c <= hi_c;
            const unsigned bal = __ballot_sync(kFull, ne);
            int LOc = 255, HIc = 0;
            if (bal) {
                const unsigned run = bal >> (__ffs(bal) - 1);
                const int nlo = __shfl_down_sync(kFull, lo_c, 1), nhi = __shfl_down_sync(kFull, hi_c, 1);
                const bool gap = ne && lane < kWarp - 1 && ((bal >> (lane + 1)) & 1u) && (nlo > hi_c + 1 || lo_c > nhi + 1);
                const bool connected = ((run + 1u) & run) == 0u && !__any_sync(kFull, gap);
                const int a0 = __reduce_min_sync(kFull, lo_c), a1 = __reduce_max_sync(kFull, hi_c);
                if (connected) { LOc = a0; HIc = a1; }
            }
            if (LO <= HI) packed = (unsigned)LO | ((unsigned)HI << 8) | ((unsigned)LOc << 16) | ((unsigned)HIc << 24);
        }
        if (lane == 0) {
            ivl[j] = packed;
            ids[j] = (unsigned short)o;
        }
    }
    __syncwarp();   // the scratch list is reused by this warp's next chunk
}

// Per candidate: lowest certain blocker, and the obstacles below it that only `maybe` block.  All threads of the CTA take part:
// work item (chunk, candidate) scans that chunk's list; the candidates' results are combined with shared-memory atomics.
// Pass 1 (T.first preset to kNoBlocker): lowest obstacle whose inner interval contains the candidate.
// (a CTA that owns only every n_parts-th candidate of a split row resolves only those)
__device__ __forceinline__ void row_table_resolve_first(const RowTable& T, int n_d, int O, int nch, int n_parts = 1, int my_part = 0) {
    const int n_own = (n_d - my_part + n_parts - 1) / n_parts;
    for (int w = threadIdx.x; w < n_own * nch; w += blockDim.x) {
        const int ch = w / n_own, id = my_part + (w - ch * n_own) * n_parts;
        const unsigned* ivl = T.ivl + ch * O;
        const unsigned short* ids = T.id + ch * O;
        const int cnt = T.cnt[ch];
        for (int j = 0; j < cnt; ++j) {
            const unsigned e = ivl[j];
            if (id >= (int)((e >> 16) & 255u) && id <= (int)(e >> 24)) {
                atomicMin(T.first + id, (int)ids[j]);   // ascending within a chunk: the first match is the chunk's lowest
                break;
            }
        }
    }
}
// Pass 2 (T.n_maybe preset to 0): obstacles below it that contain the candidate only in their outer interval.  The same obstacle can
// be listed by several chunks; it is then re-tested more than once, which is harmless.
__device__ __forceinline__ void row_table_resolve_maybe(const RowTable& T, int n_d, int O, int nch, int n_parts = 1, int my_part = 0) {
    const int n_own = (n_d - my_part + n_parts - 1) / n_parts;
    for (int w = threadIdx.x; w < n_own * nch; w += blockDim.x) {
        const int ch = w / n_own, id = my_part + (w - ch * n_own) * n_parts;
        const unsigned* ivl = T.ivl + ch * O;
        const unsigned short* ids = T.id + ch * O;
        const int cnt = T.cnt[ch];
        const int first = T.first[id];
        for (int j = 0; j < cnt; ++j) {
            const int o = ids[j];
            if (o >= first) break;          // ascending within a chunk
            const unsigned e = ivl[j];
            if (id >= (int)(e & 255u) && id <= (int)((e >> 8) & 255u)) {      // (an inner interval below `first` cannot contain it)
                const int slot = atomicAdd(T.n_maybe + id, 1);
                if (slot < kMaybeCap) T.maybe_id[id * kMaybeCap + slot] = (unsigned short)o;
            }
        }
    }
}

// Stage a scenario's obstacle snapshot in shared memory.
__device__ __forceinline__ void stage_obstacles(const FrenetBuffers& B, int b, double2* obs, float2* obsf) {
    const int O = B.n_obs;
    const double2* src = reinterpret_cast<const double2*>(B.obs_xy) + (int64_t)b * O;
    for (int o = threadIdx.x; o < O; o += blockDim.x) {
        const bool on = B.obs_active == nullptr || B.obs_active[(int64_t)b * O + o] != 0;
        const double2 q = on ? src[o] : make_double2(CUDART_INF, CUDART_INF);
        obs[o] = q;
        obsf[o] = make_float2(__double2float_rn(q.x), __double2float_rn(q.y));   // inf stays inf
    }
}

// ------------------------------------------------------------------------------------------------
// K2 (+K3 +K4 fused): evaluate / cost / feasibility, one CTA per row
// ------------------------------------------------------------------------------------------------
// kXY:     also transform every sample to Cartesian (K3) while it is in registers and store x, y.
// kColl:   also run the collision test (K4) on the Cartesian point while it is in registers: 1 = against the obstacle snapshot
//          (the reference's test), 2 = against the predicted, time-indexed table (extension).
// kLimits: track max |acceleration| / |speed| for the (extension) kinematic limits; without it the
//          KINEMATIC_OK flag is set unconditionally, which is the reference's behaviour.
// The fused variants write each candidate-timestep exactly once (48 B: d, d', d'', d''', x, y) and read
// nothing back; the unfused sequence K2 -> K3 -> K4 moves 72 B for the same result.
// Shared memory (doubles): [3 * n_obs, rounded up to 4] obstacles in fp64 and fp32 (kColl) | [n_pad_max] lateral argument |
// 4 x [n_pad_max] path frame (kXY) | [n_d][6] lateral coefficients | [n_d][5] per-candidate results | [n_d] lateral axis | [n_knots][9] spline
// table (kXY).
// Every lane owns two consecutive time steps per iteration (16-byte stores, half the per-step overhead): the mapping for long
// horizons.  Lattices of at most 32 padded samples go through k2_short below instead.
constexpr bool kRowTable = FRENET_ROW_TABLE != 0;
constexpr int k2_min_ctas(int coll, bool f32 = false) {
    if (FRENET_BULK_STORE) return 2;
    // (fp32 fast path without collisions: half the store bytes leave the kernel latency-bound, a fourth resident CTA is worth 4 %;
    //  the fp64 form is store-bound and loses 5 % at 64 registers - same-box A/B in profiles/r2_history.md)
    if (f32 && coll == 0) return 4;
    return coll == 0 ? FRENET_K2_MIN_CTAS : (coll == 1 && kRowTable) ? FRENET_K2_MIN_CTAS_TABLE : FRENET_K2_MIN_CTAS_COLL;
}
// shared memory of the row table, in doubles (after the spline table)
__host__ __device__ inline size_t row_table_doubles(int npm, int n_d, int n_obs) {
    const int chunks = (npm + kWarp - 1) / kWarp;
    const size_t bytes = (size_t)npm * 8 + (size_t)chunks * n_obs * 6 + 8 + (size_t)chunks * 4 + (size_t)n_d * (8 + 2 * kMaybeCap) + 8 +
                         (size_t)(kRowThreads / kWarp) * n_obs * 2;
    return (bytes + 15) / 8;
}

// kF32 (fp32 FAST PATH, BASELINE north star "1e-3 if an fp32 fast path is offered"): the candidate blocks (d and its three derivatives, x, y) are
//          STORED as floats (24 instead of 48 bytes per candidate-timestep; FrenetLattice.lat_f32).  Everything is still computed in
//          fp64 - costs, feasibility, the collision table and its exact re-tests work on the fp64 values, so masks, argmin and costs
//          are those of the fp64 path - and the longitudinal samples, which feed the next replan, stay fp64.
// kCollOnly: the collision stage ALONE for a lattice that has been sampled already (check_paths on new obstacles): phase A and the
//          row table as usual, no sampling loop, no stores except coll_free / first_blocker - the standalone K4 result without
//          reading a single stored sample back (the exact re-tests recompute their points from the coefficients).
template <bool kXY, int kColl, bool kLimits, bool kSplit, bool kF32, bool kCollOnly = false>
__device__ __forceinline__ void k2_evaluate_body(const FrenetLattice& L, const FrenetParams& P, const FrenetPath& path, const FrenetBuffers& B, int opts, float radius_up) {
    extern __shared__ __align__(16) double sm[];
    constexpr bool kTable = kColl == 1 && kRowTable;   // snapshot collisions through the row-level table
    const int npm = L.n_pad_max;
    double2* obs = reinterpret_cast<double2*>(sm);
    float2* obsf = reinterpret_cast<float2*>(sm + 2 * B.n_obs);
    constexpr int kSpanSteps = 2 * kWarp;                          // steps covered by one warp iteration
    const int n_chunks = (npm + kSpanSteps - 1) / kSpanSteps;
    const float4* cbs = reinterpret_cast<const float4*>(sm);       // kColl == 2: [n_chunks][n_obs] per-chunk obstacle bounds
    double* su = sm + (kColl == 1 ? 4 * ((3 * B.n_obs + 3) / 4) : kColl == 2 ? 2 * n_chunks * B.n_obs : 0);   // [npm] lateral argument: t_k, or |s_k - s0| in low-speed mode
    double* fpx = su + npm;                                        // [npm] x 4 path frame (kXY): point and unit normal
    double* fpy = fpx + (kXY ? npm : 0);
    double* fnx = fpy + (kXY ? npm : 0);
    double* fny = fnx + (kXY ? npm : 0);
    double* coef = fny + (kXY ? npm : 0);                          // [n_d][6] lateral coefficients of this row (from K1)
    double* resc = coef + L.n_d * 6;                               // [n_d][4] cd, cv, ct, ctot of the row's candidates
    int* resi = reinterpret_cast<int*>(resc + L.n_d * 4);          // [n_d][2] flags, first blocker
    double* axd = resc + L.n_d * 5;                                // [n_d] lateral target axis
    double* tab = axd + L.n_d;                                     // [n_knots][9] spline table (kXY, spline path)
    RowTable RT;
    if (kTable) {
        RT.mA = reinterpret_cast<float*>(tab + ((kXY && path.kind == FRENET_PATH_SPLINE) ? 9 * path.n_knots : 0));
        RT.mB = RT.mA + npm;
        const int chunks = (npm + kWarp - 1) / kWarp;
        RT.ivl = reinterpret_cast<unsigned*>(RT.mB + npm);
        RT.cnt = reinterpret_cast<int*>(RT.ivl + chunks * B.n_obs);
        RT.first = RT.cnt + chunks;
        RT.n_maybe = RT.first + L.n_d;
        RT.maybe_id = reinterpret_cast<unsigned short*>(RT.n_maybe + L.n_d);
        RT.id = RT.maybe_id + kMaybeCap * L.n_d;
        RT.scratch = RT.id + chunks * B.n_obs + ((chunks * B.n_obs) & 1);
    }
#if FRENET_BULK_STORE
    // per-warp staging block [6][n_pad_max] behind everything else (16-byte aligned: every block before it is a whole number of doubles pairs)
    double* bulk_stage = tab + ((kXY && path.kind == FRENET_PATH_SPLINE) ? 9 * path.n_knots : 0) + (kTable ? row_table_doubles(npm, L.n_d, B.n_obs) : 0);
    bulk_stage = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(bulk_stage) + 15) & ~uintptr_t(15));
#endif
    __shared__ double red[4][kRowThreads / 32];
    __shared__ float sh_model[2];   // row table: usable (1 / 0), M
    __shared__ int sh_next;         // next candidate to hand out (dynamic distribution over the warps)
    __shared__ double sh_carry[kLimits ? kRowThreads / 32 : 1][4];   // curvature: the last pair of a warp's previous iteration
    __shared__ double sh_clong;
    __shared__ int sh_rowfeas;

    // gridDim.y > 1: a row is split over several CTAs (small batches, where one CTA per row would leave most of the GPU idle): every
    // part repeats the row's longitudinal phase and takes every parts-th round of candidates.  kSplit = false compiles the split away
    // (this kernel has no register to spare in the batched case).
#define parts (kSplit ? (int)gridDim.y : 1)
#define part (kSplit ? (int)blockIdx.y : 0)
    const int64_t br = blockIdx.x;
    const RowAddr A = row_addr(L, br);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    if (skipped(B, A.b)) return;
    const bool follow = follow_of(P, B, A.b);
    const int flags = B.row_flags[br];

    if (!(flags & FRENET_ROW_EXISTS) || (flags & FRENET_ROW_DROPPED)) {
        for (int id = tid; id < L.n_d; id += blockDim.x) {
            const int64_t c = A.cand0 + id;
            if (!kCollOnly) {
                B.cand_flags[c] = 0;
#pragma unroll
                for (int f = 0; f < 4; ++f) B.cost[f * A.candF + c] = 0.0;
            }
            if (kColl) {
                B.coll_free[c] = 0;
                B.first_blocker[c] = -1;
            }
            if (kLimits && kXY && (opts & 1)) B.curv_ok[c] = 0;
        }
        return;
    }
    const bool low = flags & FRENET_ROW_LOWSPEED;
    const double dt = P.sample_period;
    const double s0 = B.state[(int64_t)A.b * 6 + 3];
    const double* lane_q = (kXY && B.lane_params) ? B.lane_params + 4 * (int64_t)A.b : nullptr;   // per-scenario lane parabola

    // Phase 0: stage the row's lateral coefficients, the spline table and the obstacle snapshot.
    {
        const double* cq = B.coef_lat + A.cand0 * 6;
        for (int i = tid; i < L.n_d * 6; i += blockDim.x) coef[i] = cq[i];
        for (int i = tid; i < L.n_d; i += blockDim.x) axd[i] = L.axis_d[i];
        if (kColl == 1) stage_obstacles(B, A.b, obs, obsf);
        if (kColl == 2) {
            const float4* src = reinterpret_cast<const float4*>(B.obs_bounds) + (int64_t)A.b * n_chunks * B.n_obs;
            float4* dst = reinterpret_cast<float4*>(sm);
            for (int i = tid; i < n_chunks * B.n_obs; i += blockDim.x) dst[i] = src[i];
        }
        if (kXY && path.kind == FRENET_PATH_SPLINE) {
            for (int i = tid; i < path.n_knots * 9; i += blockDim.x) tab[i] = path.spline_table[i];
            if (!kTable) __syncthreads();   // phase A reads the table
        }
        if (kTable) {
            if (tid == 0) sh_next = 0;
            __syncthreads();   // phase A reads the spline table and the first / last candidate's coefficients
        }
    }
    FRENET_PHASE(3);
    // Row table: usable only if the lateral axis is a uniform ascending grid and the staged coefficients are affine in the target
    // (what K1 produces; anything else - a caller's own coefficients - takes the per-candidate sweep).
    bool table_bad = false;
    double D0 = 0.0, D1 = 0.0;
    if (kTable) {
        const int nd = L.n_d;
        D0 = axd[0];
        D1 = axd[nd - 1];
        table_bad = nd > kTableMaxD || B.n_obs > 65535 || (nd > 1 && !(D1 > D0));
        if (nd > 1 && tid < nd) {     // (n_d <= 128 < blockDim.x whenever the table is used)
            const int i = tid;
            const double span = D1 - D0, inv_n = 1.0 / (double)(nd - 1), inv_s = 1.0 / span;
            const double Di = axd[i];
            table_bad |= !(fabs(Di - fma((double)i * inv_n, span, D0)) <= 1e-9 * (fabs(D0) + fabs(D1)));
            const double wi = (Di - D0) * inv_s;
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const double lo = coef[j], hi = coef[(nd - 1) * 6 + j];
                table_bad |= !(fabs(coef[i * 6 + j] - fma(wi, hi - lo, lo)) <= 1e-12 * fmax(fabs(lo), fabs(hi)) + 1e-300);
            }
        }
    }

    // Phase A: longitudinal samples, once per row (the reference deep-copies them into each of the n_d
    // candidates; here the candidates share the row's arrays), the lateral argument and the path frame.
    {
        Poly pl;
        pl.load(B.coef_lon + br * 6);
        double* lon = B.lon + A.lon_off;
        double acc = 0.0, vmax = 0.0, amax = 0.0;
        double model_m = table_bad ? CUDART_INF : 0.0;   // row table: magnitude of the lateral polynomials' terms; +inf = not usable
        const double inv_span = (kTable && L.n_d > 1) ? 1.0 / (D1 - D0) : 0.0;
        const double dmax_abs = fmax(fabs(D0), fabs(D1));
        // Samples k in [N, npad) are padding: they are evaluated and stored like real samples so that every 32-byte
        // sector of the output is written in full (a partially written sector costs a DRAM read-modify-write, which
        // halves the achieved store bandwidth), but they take no part in costs, limits or collisions.
        for (int k = tid; k < A.npad; k += blockDim.x) {
            const double t = (double)k * dt;   // == np.arange(0, T + dt, dt)[k]
            double s, v, a, j;
            eval_power_form(pl, t, L.t_pow + 4 * k, s, v, a, j);
            if (part == 0 && !kCollOnly) {
                lon[k] = s;
                lon[A.npad + k] = v;
                lon[2 * A.npad + k] = a;
                lon[3 * A.npad + k] = j;
            }
            su[k] = low ? fabs(s - s0) : t;
            if (k < A.N) {
                acc = fma(j, j, acc);
                if (kLimits) {
                    vmax = fmax(vmax, fabs(v));
                    amax = fmax(amax, fabs(a));
                }
            }
            if (kXY) {
                double px, py, nx, ny;
                const double sp = B.path_origin ? B.path_origin[2 * A.b] + (s - B.path_origin[2 * A.b + 1]) : s;   // frenet_to_glob_planner
                path_frame(path.kind, path.circle, tab, path.n_knots, sp, px, py, nx, ny, lane_q);
                fpx[k] = px; fpy[k] = py; fnx[k] = nx; fny[k] = ny;
                if (kTable && k < A.N) {
                    // affine model of the row's lateral offsets at this step, from its first and last candidate (Horner like Poly::eval)
                    const double u = low ? fabs(s - s0) : t, ua = fabs(u);
                    const double* c0 = coef;
                    const double* c1 = coef + (L.n_d - 1) * 6;
                    const double d0 = fma(fma(fma(fma(fma(c0[5], u, c0[4]), u, c0[3]), u, c0[2]), u, c0[1]), u, c0[0]);
                    const double d1 = fma(fma(fma(fma(fma(c1[5], u, c1[4]), u, c1[3]), u, c1[2]), u, c1[1]), u, c1[0]);
                    const double m0 = fma(fma(fma(fma(fma(fabs(c0[5]), ua, fabs(c0[4])), ua, fabs(c0[3])), ua, fabs(c0[2])), ua, fabs(c0[1])), ua, fabs(c0[0]));
                    const double m1 = fma(fma(fma(fma(fma(fabs(c1[5]), ua, fabs(c1[4])), ua, fabs(c1[3])), ua, fabs(c1[2])), ua, fabs(c1[1])), ua, fabs(c1[0]));
                    const double Bk = (d1 - d0) * inv_span, Ak = d0 - Bk * D0;
                    RT.mA[k] = __double2float_rn(Ak);
                    RT.mB[k] = __double2float_rn(Bk);
                    const bool fine = (fabs(Ak) + fabs(Bk) * dmax_abs <= 1e12) && (fabs(px) + fabs(py) + fabs(nx) + fabs(ny) < 1e30);   // false for NaN
                    model_m = fine ? fmax(model_m, fmax(m0, m1)) : CUDART_INF;
                }
            }
        }
        // longitudinal jerk sum and limits: per-warp partials combined in a fixed order (deterministic)
        acc = warp_sum(acc);
        if (kLimits) {
            vmax = warp_max(vmax);
            amax = warp_max(amax);
        }
        if (kTable) model_m = warp_max(model_m);   // (fmax drops a NaN operand: m0 / m1 NaN would vanish, but then Ak / Bk are NaN too and `fine` is false)
        if (lane == 0) {
            red[0][warp] = acc;
            red[1][warp] = vmax;
            red[2][warp] = amax;
            red[3][warp] = model_m;
        }
        __syncthreads();
        if (tid == 0) {   // trajectory_planner.py:143-144, 153-154
            double jsum = 0.0, vm = 0.0, am = 0.0, mm = 0.0;
            for (int w = 0; w < nwarps; ++w) {
                jsum += red[0][w];
                vm = fmax(vm, red[1][w]);
                am = fmax(am, red[2][w]);
                mm = fmax(mm, red[3][w]);
            }
            if (kTable) {
                sh_model[0] = (mm <= 1e15) ? 1.0f : 0.0f;
                sh_model[1] = __double2float_ru(fmin(mm, 1e15));
            }
            const double T = L.axis_t[A.iT];
            double s_last, v_last, a_last, j_last;
            eval_power_form(pl, (double)(A.N - 1) * dt, L.t_pow + 4 * (A.N - 1), s_last, v_last, a_last, j_last);
            double term;
            if (follow) {
                const double e = s_last - B.target[((int64_t)A.b * L.n_t + A.iT) * 3];
                term = __dmul_rn(P.ks, __dmul_rn(e, e));
            } else {
                const double e = v_last - B.scal[(int64_t)A.b * 3 + 1];
                term = __dmul_rn(P.kdots, __dmul_rn(e, e));
            }
            sh_clong = __dadd_rn(__dadd_rn(__dmul_rn(P.kj, jsum), __dmul_rn(P.kt, T)), term);
            sh_rowfeas = kLimits ? ((vm <= P.v_max) && (am <= P.a_max)) : 1;
        }
        __syncthreads();
    }
    FRENET_PHASE(4);
    const double clong = sh_clong;
    const bool rowfeas = sh_rowfeas;
    // Row table (see row_table_chunk): built by all warps once per row, then resolved per candidate.
    const bool table_ok = kTable && sh_model[0] != 0.0f;
    const float Rf = radius_up;   // robot radius, rounded up on the host
    const int nch = (A.N + kWarp - 1) / kWarp;
    if (kTable && table_ok) {
        // The chunks' obstacle lists are dealt to the warps so that (almost) every warp gets one item; nobody waits for the table:
        // it is consulted only after the sampling loops below (the candidates are handed out dynamically, which evens out the warps
        // that built more of it).
        const float R2f = __double2float_rn(P.radius_sq);
        const float cM = fmaf(2e-12f, sh_model[1], 1e-9f);
        const int nsub = FRENET_TABLE_SPLIT > 0 ? FRENET_TABLE_SPLIT : max(1, (nwarps + nch - 1) / nch);
        for (int item = warp; item < nch * nsub; item += nwarps)
            row_table_chunk(RT, item / nsub, item % nsub, nsub, A.N, fpx, fpy, fnx, fny, obs, B.n_obs, D0, D1, L.n_d, Rf, R2f, cM,
                            RT.scratch + warp * B.n_obs, lane);
    }
    FRENET_PHASE(11);     // (warp 0 has built its share of the row table)
    // row constants of the lateral cost, fetched once (not per candidate, where a global load would stall the warp's epilogue)
    const double horizon = low ? B.row_S[br] : L.axis_t[A.iT];
    const double dd_target = B.scal[(int64_t)A.b * 3];
    // predicted mode: the scenario's table [n_obs][n_obs_steps] stays in global memory (L2 resident: every row of a scenario reads it)
    const double2* tblp = kColl == 2 ? reinterpret_cast<const double2*>(B.obs_xy_t) + (int64_t)A.b * B.n_obs * B.n_obs_steps : nullptr;
    const int npad = A.npad, N = A.N;
    // curvature limit (extension, with the kinematic limits): Menger curvature of every interior sample, tested while the
    // samples are in registers - a lane needs its left neighbour's pair, lane 0 the last pair of the previous iteration
    const bool do_curv = kLimits && kXY && (opts & 1);
    const double k2max = P.kappa_max * P.kappa_max;

    // Phase B: a warp per candidate; every lane owns TWO consecutive time steps per iteration (64 steps per warp
    // iteration, 16-byte loads from shared memory and 16-byte stores to HBM); no global loads in this loop.
    // Candidates are dealt round-robin to the warps (handed out dynamically in the row-table variant, where some warps
    // arrive late from building the table); their results are staged in shared memory and published after a barrier.
    // Candidate of hand-out slot q: id = q * parts + part.
    int slot = warp;
    const unsigned next_addr = (unsigned)__cvta_generic_to_shared(&sh_next);
    for (;;) {
        if (kCollOnly) break;
        if (kTable) {
            int t = 0;
            if (lane == 0) asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(t) : "r"(next_addr) : "memory");
            slot = __shfl_sync(kFull, t, 0);
        }
        const int id = slot * parts + part;
        if (id >= L.n_d) break;
        Poly pq;
        pq.load(coef + id * 6);
        // [6][npad] block of this candidate; `out` walks along k, the fields sit `fs` bytes apart
#if FRENET_BULK_STORE
        double* stg = bulk_stage + warp * (kCandFields * npm);
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous candidate's block has left the staging buffer
        __syncwarp();
        unsigned out = (unsigned)__cvta_generic_to_shared(stg + 2 * lane);      // shared-window address: the stores below are STS.128
#define FRENET_PUT(addr, v0, v1) asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v0), "d"(v1) : "memory")
#else
#define FRENET_PUT(addr, v0, v1)                                                                                             \
    do {                                                                                                                     \
        if (kF32) *reinterpret_cast<float2*>(addr) = make_float2(__double2float_rn(v0), __double2float_rn(v1));              \
        else *reinterpret_cast<double2*>(addr) = make_double2(v0, v1);                                                       \
    } while (0)
        const int64_t cand_elem = A.cand_off + (int64_t)id * (kCandFields * npad) + 2 * lane;
        char* out = kF32 ? reinterpret_cast<char*>(reinterpret_cast<float*>(B.lat) + cand_elem) : reinterpret_cast<char*>(B.lat + cand_elem);
#endif
#if FRENET_BULK_STORE
        const unsigned fs = (unsigned)(npad * sizeof(double));
#else
        const int64_t fs = (int64_t)npad * (kF32 ? sizeof(float) : sizeof(double));
#endif
        double acc = 0.0, amax = 0.0;
        int best = kNoBlocker;
        bool curv_bad = false;
#pragma unroll 1
        for (int kb = 0; kb < N; kb += 2 * kWarp) {
            const int k0 = kb + 2 * lane;
            const bool store = k0 < npad;           // npad is even: a pair is stored whole or not at all
            const bool valid_a = k0 < N, valid_b = k0 + 1 < N;   // real samples (the rest is padding)
            const int kk = store ? k0 : kb;         // lanes beyond the padding shadow the chunk's first pair
            const double2 u = *reinterpret_cast<const double2*>(su + kk);
            double da, va, aa, ja, db, vb, ab, jb;
#if FRENET_K2_RELOAD_COEF
            if (kColl && !kTable) pq.load_shared_here(coef + id * 6);
#endif
            pq.eval(u.x, da, va, aa, ja);
            pq.eval(u.y, db, vb, ab, jb);
            auto o = out;
            if (store) {
                FRENET_PUT(o, da, db);
                FRENET_PUT(o += fs, va, vb);
                FRENET_PUT(o += fs, aa, ab);
                FRENET_PUT(o += fs, ja, jb);
            }
            if (valid_a) {
                acc = fma(ja, ja, acc);
                if (kLimits) amax = fmax(amax, fabs(aa));
            }
            if (valid_b) {
                acc = fma(jb, jb, acc);
                if (kLimits) amax = fmax(amax, fabs(ab));
            }
            if (kXY) {
                const double2 px = *reinterpret_cast<const double2*>(fpx + kk), py = *reinterpret_cast<const double2*>(fpy + kk);
                const double2 nx = *reinterpret_cast<const double2*>(fnx + kk), ny = *reinterpret_cast<const double2*>(fny + kk);
                const double xa = __dadd_rn(px.x, __dmul_rn(nx.x, da)), xb = __dadd_rn(px.y, __dmul_rn(nx.y, db));   // pt + n * d,
                const double ya = __dadd_rn(py.x, __dmul_rn(ny.x, da)), yb = __dadd_rn(py.y, __dmul_rn(ny.y, db));   // unfused like the reference
                if (store) {
                    FRENET_PUT(o += fs, xa, xb);
                    FRENET_PUT(o += fs, ya, yb);
                }
                if (kLimits) {
                    if (do_curv) {
                        double qxa = __shfl_up_sync(kFull, xa, 1), qya = __shfl_up_sync(kFull, ya, 1);
                        double qxb = __shfl_up_sync(kFull, xb, 1), qyb = __shfl_up_sync(kFull, yb, 1);
                        if (lane == 0 && kb > 0) {   // lane 31's pair of the previous iteration, parked in shared memory
                            qxa = sh_carry[warp][0]; qya = sh_carry[warp][1]; qxb = sh_carry[warp][2]; qyb = sh_carry[warp][3];
                        }
                        // centres k0 - 1 (the neighbour's second sample) and k0 (this lane's first); interior means 1 <= k <= N - 2
                        if (k0 >= 2 && k0 <= N - 1) curv_bad |= menger_exceeds(qxa, qya, qxb, qyb, xa, ya, k2max);
                        if (k0 >= 1 && k0 <= N - 2) curv_bad |= menger_exceeds(qxb, qyb, xa, ya, xb, yb, k2max);
                        __syncwarp();
                        if (lane == kWarp - 1) { sh_carry[warp][0] = xa; sh_carry[warp][1] = ya; sh_carry[warp][2] = xb; sh_carry[warp][3] = yb; }
                        __syncwarp();
                    }
                }
                // (padding lanes carry the trajectory's continuation: they can only widen the cull circle, never hit)
                if (kColl == 1 && !kTable)
                    best = collide_pair(xa, ya, xb, yb, valid_a, valid_b, min(2 * kWarp, N - kb), obs, obsf, B.n_obs, Rf, P.radius_sq,
                                        best, lane);
                if (kColl == 2)
                    best = collide_pair_predicted(xa, ya, xb, yb, valid_a, valid_b, min(2 * kWarp, N - kb), kk, tblp, B.n_obs_steps,
                                                  cbs + (kb / (2 * kWarp)) * B.n_obs, B.n_obs, Rf, P.radius_sq, best, lane);
            }
            out += 2 * kWarp * (kF32 ? sizeof(float) : sizeof(double));
        }
#if FRENET_BULK_STORE
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the lanes' generic-proxy stores become visible to the copy engine
        __syncwarp();
        if (lane == 0) {
            double* gdst = B.lat + A.cand_off + (int64_t)id * (kCandFields * npad);
            const unsigned ssrc = (unsigned)__cvta_generic_to_shared(stg);
            const unsigned bytes = (unsigned)((kXY ? kCandFields : 4) * npad * sizeof(double));
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
#endif
        acc = warp_sum(acc);
        if (kLimits) {
            amax = warp_max(amax);
            curv_bad = __any_sync(kFull, curv_bad);
        }
        if (lane == 0) {   // :165-167 / :174-176; staged in shared memory, written out coalesced below
            const double e = axd[id] - dd_target;
            const double clat = __dadd_rn(__dadd_rn(__dmul_rn(P.kj, acc), __dmul_rn(P.kt, horizon)), __dmul_rn(P.kd, __dmul_rn(e, e)));
            resc[id * 4] = clat;
            resc[id * 4 + 1] = follow ? 0.0 : clong;
            resc[id * 4 + 2] = follow ? clong : 0.0;
            resc[id * 4 + 3] = __dadd_rn(__dmul_rn(P.klat, clat), __dmul_rn(P.klong, clong));
            const bool kin = kLimits ? (rowfeas && amax <= P.a_max) : true;
            resi[id * 2] = FRENET_CAND_VALID | (kin ? FRENET_CAND_KINEMATIC_OK : 0) | (curv_bad ? 0 : kCurvOkBit);
            resi[id * 2 + 1] = best;
        }
        if (!kTable) slot += nwarps;
    }
#if FRENET_BULK_STORE
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // shared memory must outlive the copies that read it; the
                                                                              // collision post-pass below re-reads the blocks from global memory
#endif
    if (kTable) {
        // Collision results from the row table, once every candidate has been sampled (nobody waited for the table above).
        if (table_ok && tid < L.n_d) {      // (n_d <= 128 <= blockDim.x whenever the table is used)
            RT.first[tid] = kNoBlocker;
            RT.n_maybe[tid] = 0;
        }
        __syncthreads();
        FRENET_PHASE(5);
        if (tid == 0) sh_next = 0;          // every warp has left the hand-out loop: the counter now counts the candidates left for the exact test
        // A_k / B_k are not needed any more: their storage ([2 n_pad_max] floats >= 132 shorts >= n_d) takes the list of candidates
        // that still need the exact test - about one in a hundred; everybody else's answer is final after the resolve passes
        unsigned short* todo = reinterpret_cast<unsigned short*>(RT.mA);
        int n_todo = (L.n_d - part + parts - 1) / parts;       // rows the table does not cover: every own candidate
        if (table_ok) {
            row_table_resolve_first(RT, L.n_d, B.n_obs, nch, parts, part);
            __syncthreads();
            row_table_resolve_maybe(RT, L.n_d, B.n_obs, nch, parts, part);
            __syncthreads();
#if FRENET_POST_COMPACT
            if (tid < L.n_d && tid % parts == part) {
                resi[tid * 2 + 1] = RT.first[tid];
                if (RT.n_maybe[tid] > 0) todo[atomicAdd(&sh_next, 1)] = (unsigned short)tid;
            }
            __syncthreads();
            n_todo = sh_next;
#else
            if (tid < n_todo) todo[tid] = (unsigned short)(tid * parts + part);     // A/B: every candidate walks through the post pass
            __syncthreads();
#endif
        }
        FRENET_PHASE(12);
        // The few candidates with an uncertain obstacle below their first certain blocker - and every candidate of a row the table
        // does not cover - get the reference's fp64 test on their samples, recomputed from shared memory exactly as the sampling loop
        // formed them (same FMA sequence, same unfused pt + n * d), so that nothing is read back from HBM (and the fp32 fast path,
        // whose stored samples are rounded, tests the same fp64 points).
        for (int q = warp; q < n_todo; q += nwarps) {
            const int id = table_ok ? (int)todo[q] : q * parts + part;
            int best = kNoBlocker, nm = kMaybeCap + 1;
            if (table_ok) {
                best = RT.first[id];
                nm = RT.n_maybe[id];
            }
            if (nm > 0) {
                Poly pq;
                pq.load(coef + id * 6);
                auto pair_at = [&](int kb, double& xa, double& ya, double& xb, double& yb) {
                    const int k0 = kb + 2 * lane;
                    const int kk = k0 < npad ? k0 : kb;     // lanes beyond the padding shadow the chunk's first pair (never valid)
                    const double2 u = *reinterpret_cast<const double2*>(su + kk);
                    const double da = pq.value(u.x), db = pq.value(u.y);
                    const double2 px = *reinterpret_cast<const double2*>(fpx + kk), py = *reinterpret_cast<const double2*>(fpy + kk);
                    const double2 nx = *reinterpret_cast<const double2*>(fnx + kk), ny = *reinterpret_cast<const double2*>(fny + kk);
                    xa = __dadd_rn(px.x, __dmul_rn(nx.x, da)); xb = __dadd_rn(px.y, __dmul_rn(nx.y, db));
                    ya = __dadd_rn(py.x, __dmul_rn(ny.x, da)); yb = __dadd_rn(py.y, __dmul_rn(ny.y, db));
                };
                if (nm > kMaybeCap) {   // per-candidate sweep over the obstacles below `best`
                    for (int kb = 0; kb < N; kb += 2 * kWarp) {
                        const int k0 = kb + 2 * lane;
                        double xa, ya, xb, yb;
                        pair_at(kb, xa, ya, xb, yb);
                        best = collide_pair(xa, ya, xb, yb, k0 < N, k0 + 1 < N, min(2 * kWarp, N - kb), obs, obsf, B.n_obs, Rf, P.radius_sq, best,
                                            lane);
                    }
                } else {
                    for (int j = 0; j < nm; ++j) {
                        const int o = RT.maybe_id[id * kMaybeCap + j];
                        if (o >= best) continue;        // (a lower one has been confirmed already)
                        const double2 ob = obs[o];
                        bool hit = false;
                        for (int kb = 0; kb < N; kb += 2 * kWarp) {
                            const int k0 = kb + 2 * lane;
                            double xa, ya, xb, yb;
                            pair_at(kb, xa, ya, xb, yb);
                            const double dxa = xa - ob.x, dya = ya - ob.y, dxb = xb - ob.x, dyb = yb - ob.y;
                            hit |= (k0 < N && (__dadd_rn(__dmul_rn(dxa, dxa), __dmul_rn(dya, dya)) <= P.radius_sq)) ||
                                   (k0 + 1 < N && (__dadd_rn(__dmul_rn(dxb, dxb), __dmul_rn(dyb, dyb)) <= P.radius_sq));
                        }
                        if (__any_sync(kFull, hit)) best = o;
                    }
                }
            }
            if (lane == 0) resi[id * 2 + 1] = best;
        }
    }
    __syncthreads();
    FRENET_PHASE(13);
    for (int id = tid; id < L.n_d; id += blockDim.x) {
        if (id % parts != part) continue;   // another part's candidate
        const int64_t c = A.cand0 + id;
        if (!kCollOnly) {
#pragma unroll
            for (int f = 0; f < 4; ++f) B.cost[f * A.candF + c] = resc[id * 4 + f];
            B.cand_flags[c] = (uint8_t)resi[id * 2];
            if (do_curv) B.curv_ok[c] = (resi[id * 2] & kCurvOkBit) ? 1 : 0;
        }
        if (kColl) {
            const int best = resi[id * 2 + 1];
            B.coll_free[c] = best == kNoBlocker ? 1 : 0;
            B.first_blocker[c] = best == kNoBlocker ? -1 : best;
        }
    }
}

// kMono: the one-launch replan (see mono_tail) - K1 of the row in front, K5 + gather by the scenario's last CTA behind.
template <bool kXY, int kColl, bool kLimits, bool kSplit, bool kF32 = false, bool kMono = false, bool kCollOnly = false>
__global__ void __launch_bounds__(kRowThreads, k2_min_ctas(kColl, kF32))      // (kMono too: a dense single planner is hundreds of CTAs, occupancy matters)
k2_evaluate(FrenetLattice L, FrenetParams P, FrenetPath path, FrenetBuffers B, int opts, float radius_up, int use_mask, int gather_sel, int has_xy,
            FrenetInline inl) {
    FRENET_PHASE(0);
    if (kMono && !kCollOnly) {
        spill_inline(inl, L, B);
        FRENET_PHASE(1);
        // every part of a split row solves the row's n_d cells itself (identical values: the duplicate stores are harmless)
        const int64_t cand0 = (int64_t)blockIdx.x * L.n_d;
        for (int id = threadIdx.x; id < L.n_d; id += blockDim.x) k1_cell(L, P, B, cand0 + id);
        __syncthreads();
    }
    FRENET_PHASE(2);
    k2_evaluate_body<kXY, kColl, kLimits, kSplit, kF32, kCollOnly>(L, P, path, B, opts, radius_up);
    FRENET_PHASE(6);
    if (kMono) {
        const int R = L.n_v_cap * L.n_t;
        const int b = (int)(blockIdx.x / R), r = (int)(blockIdx.x % R);
        if (skipped(B, b)) return;
        // this CTA produced the candidates part, part + parts, ... of row r
        mono_tail(L, P, B, b, r * parts + part, R * parts, r * L.n_d + part, (L.n_d - part + parts - 1) / parts, parts, use_mask, gather_sel, has_xy);
    }
}

#undef parts
#undef part
#undef FRENET_PUT

// ------------------------------------------------------------------------------------------------
// K2 (+K3 +K4 fused) for SHORT horizons: at most 32 padded samples per candidate (the reference's default lattice has 11..26)
// ------------------------------------------------------------------------------------------------
// A row of such a lattice is too little work for a CTA of its own (16 candidates x 26 steps against a path-frame phase that is a
// few hundred dependent fp64 instructions deep), so a CTA takes `G` consecutive rows of one scenario: a WARP per row for the
// longitudinal phase (lanes = time steps; the jerk sum and the limits are warp reductions), then the G x n_d candidates of the
// group two or four per warp (16 or 8 lanes each, two time steps per lane).  Same arithmetic as k2_evaluate.
// Shared memory (doubles): obstacles as in k2_evaluate | [G][n_pad_max] lateral argument | 4 x [G][n_pad_max] path frame (kXY) |
// [G n_d][6] lateral coefficients | [G n_d][5] per-candidate results | [n_d] lateral axis | [G][2] row cost and horizon |
// [G] candidate block offsets | [G][4] ints: N, npad, usable, feasible | spline table (kXY).
template <bool kXY, int kColl, bool kLimits>
__device__ __forceinline__ void k2_short_body(const FrenetLattice& L, const FrenetParams& P, const FrenetPath& path, const FrenetBuffers& B, int G, int opts) {
    extern __shared__ __align__(16) double sm[];
    const int npm = L.n_pad_max, nd = L.n_d;
    const int R = L.n_v_cap * L.n_t;
    const int groups = (R + G - 1) / G;
    const int b = blockIdx.x / groups;
    const int row0 = (blockIdx.x % groups) * G;
    const int Ga = min(G, R - row0);                               // rows of this group
    const int GC = G * nd;
    double2* obs = reinterpret_cast<double2*>(sm);
    float2* obsf = reinterpret_cast<float2*>(sm + 2 * B.n_obs);
    const float4* cbs = reinterpret_cast<const float4*>(sm);       // kColl == 2: [n_obs] bounds of the only chunk
    double* su = sm + (kColl == 1 ? 4 * ((3 * B.n_obs + 3) / 4) : kColl == 2 ? 2 * B.n_obs : 0);
    double* fpx = su + G * npm;
    double* fpy = fpx + (kXY ? G * npm : 0);
    double* fnx = fpy + (kXY ? G * npm : 0);
    double* fny = fnx + (kXY ? G * npm : 0);
    double* coef = fny + (kXY ? G * npm : 0);
    double* resc = coef + GC * 6;
    int* resi = reinterpret_cast<int*>(resc + GC * 4);
    double* axd = resc + GC * 5;
    double* rowd = axd + nd;
    int64_t* rowo = reinterpret_cast<int64_t*>(rowd + 2 * G);
    int* rowi = reinterpret_cast<int*>(rowo + G);
    double* tab = rowd + 5 * G;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int64_t br0 = (int64_t)b * R + row0;
    const int64_t cand0 = br0 * nd;                                // the group's candidates are contiguous in every per-candidate array
    const int64_t candF = (int64_t)L.n_scen * R * nd;
    const double dt = P.sample_period;
    const double s0 = B.state[(int64_t)b * 6 + 3];
    const double* lane_q = (kXY && B.lane_params) ? B.lane_params + 4 * (int64_t)b : nullptr;
    if (skipped(B, b)) return;
    const bool follow = follow_of(P, B, b);

    // Phase 0: lateral coefficients of the whole group, lateral axis, obstacles, spline table.
    {
        const double* cq = B.coef_lat + cand0 * 6;
        for (int i = tid; i < Ga * nd * 6; i += blockDim.x) coef[i] = cq[i];
        for (int i = tid; i < nd; i += blockDim.x) axd[i] = L.axis_d[i];
        if (kColl == 1) stage_obstacles(B, b, obs, obsf);
        if (kColl == 2) {
            const float4* src = reinterpret_cast<const float4*>(B.obs_bounds) + (int64_t)b * B.n_obs;   // one chunk per scenario
            float4* dst = reinterpret_cast<float4*>(sm);
            for (int i = tid; i < B.n_obs; i += blockDim.x) dst[i] = src[i];
        }
        if (kXY && path.kind == FRENET_PATH_SPLINE) {
            for (int i = tid; i < path.n_knots * 9; i += blockDim.x) tab[i] = path.spline_table[i];
            __syncthreads();
        }
    }

    FRENET_PHASE(3);
    // Phase A: a warp per row.
    for (int g = warp; g < Ga; g += nwarps) {
        const int64_t br = br0 + g;
        const int flags = B.row_flags[br];
        const RowAddr A = row_addr(L, br);
        const bool usable = (flags & FRENET_ROW_EXISTS) && !(flags & FRENET_ROW_DROPPED);
        if (lane == 0) {
            rowi[4 * g] = A.N;
            rowi[4 * g + 1] = A.npad;
            rowi[4 * g + 2] = usable ? 1 : 0;
            rowo[g] = A.cand_off;
        }
        if (!usable) continue;
        const bool low = flags & FRENET_ROW_LOWSPEED;
        Poly pl;
        pl.load(B.coef_lon + br * 6);
        double* lon = B.lon + A.lon_off;
        double acc = 0.0, vmax = 0.0, amax = 0.0;
        const int k = lane;
        if (k < A.npad) {   // padding samples are evaluated and stored like real ones (whole sectors, see k2_evaluate)
            const double t = (double)k * dt;
            double s, v, a, j;
            eval_power_form(pl, t, L.t_pow + 4 * k, s, v, a, j);
            lon[k] = s;
            lon[A.npad + k] = v;
            lon[2 * A.npad + k] = a;
            lon[3 * A.npad + k] = j;
            su[g * npm + k] = low ? fabs(s - s0) : t;
            if (k < A.N) {
                acc = fma(j, j, acc);
                if (kLimits) {
                    vmax = fmax(vmax, fabs(v));
                    amax = fmax(amax, fabs(a));
                }
            }
            if (kXY) {
                double px, py, nx, ny;
                const double sp = B.path_origin ? B.path_origin[2 * b] + (s - B.path_origin[2 * b + 1]) : s;
                path_frame(path.kind, path.circle, tab, path.n_knots, sp, px, py, nx, ny, lane_q);
                fpx[g * npm + k] = px; fpy[g * npm + k] = py; fnx[g * npm + k] = nx; fny[g * npm + k] = ny;
            }
        }
        acc = warp_sum(acc);
        if (kLimits) {
            vmax = warp_max(vmax);
            amax = warp_max(amax);
        }
        if (lane == 0) {   // trajectory_planner.py:143-144, 153-154
            const double jsum = 0.0 + acc;
            const double T = L.axis_t[A.iT];
            double s_last, v_last, a_last, j_last;
            eval_power_form(pl, (double)(A.N - 1) * dt, L.t_pow + 4 * (A.N - 1), s_last, v_last, a_last, j_last);
            double term;
            if (follow) {
                const double e = s_last - B.target[((int64_t)b * L.n_t + A.iT) * 3];
                term = __dmul_rn(P.ks, __dmul_rn(e, e));
            } else {
                const double e = v_last - B.scal[(int64_t)b * 3 + 1];
                term = __dmul_rn(P.kdots, __dmul_rn(e, e));
            }
            rowd[2 * g] = __dadd_rn(__dadd_rn(__dmul_rn(P.kj, jsum), __dmul_rn(P.kt, T)), term);
            rowd[2 * g + 1] = low ? B.row_S[br] : T;
            rowi[4 * g + 3] = kLimits ? ((vmax <= P.v_max) && (amax <= P.a_max)) : 1;
        }
    }
    __syncthreads();
    FRENET_PHASE(4);

    const double dd_target = B.scal[(int64_t)b * 3];
    const float Rf = kColl ? __double2float_ru(sqrt(P.radius_sq)) : 0.0f;
    const double2* tblp = kColl == 2 ? reinterpret_cast<const double2*>(B.obs_xy_t) + (int64_t)b * B.n_obs * B.n_obs_steps : nullptr;

    const bool do_curv = kLimits && kXY && (opts & 1);   // curvature limit, see k2_evaluate
    const double k2max = P.kappa_max * P.kappa_max;

    // Phase B: SEVERAL candidates per warp, two consecutive time steps per lane (16-byte accesses).  A horizon of at most 32 padded
    // samples fits 16 lanes, one of at most 16 fits 8: a row's candidates are taken two or four at a time, and the per-candidate
    // overhead (coefficients, reductions, cost arithmetic, collision set-up) is paid once per group of candidates.
    // Work items = (row, group of 32 / lpc candidates), numbered row by row; lane g keeps row g's first item (the host keeps G <= 32).
    const int lpc_l = (lane < Ga && rowi[4 * lane + 1] <= 16) ? 8 : 16;
    const int items_l = lane < Ga ? (nd * lpc_l + kWarp - 1) / kWarp : 0;
    int first_item = items_l;
#pragma unroll
    for (int o = 1; o < kWarp; o <<= 1) {
        const int v = __shfl_up_sync(kFull, first_item, o);
        if (lane >= o) first_item += v;
    }
    const int total_items = __shfl_sync(kFull, first_item, kWarp - 1);
    first_item -= items_l;                                  // exclusive prefix: row `lane` starts at this item
    for (int it = warp; it < total_items; it += nwarps) {
        const int g = __popc(__ballot_sync(kFull, lane < Ga && first_item <= it)) - 1;
        if (!rowi[4 * g + 2]) continue;                     // row does not exist / was dropped
        const int lpc = __shfl_sync(kFull, lpc_l, g);
        const int local = it - __shfl_sync(kFull, first_item, g);
        const int shift = lpc == 16 ? 4 : 3;
        const int sub = lane >> shift, hl = lane & (lpc - 1);
        const int per = kWarp >> shift;                     // candidates per warp: 2 or 4
        const int id_raw = local * per + sub;
        const bool act = id_raw < nd;                       // the last group of a row may be short: idle groups shadow its first candidate
        const int id = act ? id_raw : local * per;
        const int c = g * nd + id;
        const int N = rowi[4 * g], npad = rowi[4 * g + 1];
        const int k0 = 2 * hl;
        const bool store = act && k0 < npad;                // npad is even: a pair is stored whole or not at all
        const bool valid_a = act && k0 < N, valid_b = act && k0 + 1 < N;
        const int kk = k0 < npad ? k0 : 0;                  // lanes beyond the padding shadow the first pair
        Poly pq;
        pq.load(coef + c * 6);
        const double2 u = *reinterpret_cast<const double2*>(su + g * npm + kk);
        double da, va, aa, ja, db, vb, ab, jb;
        pq.eval(u.x, da, va, aa, ja);
        pq.eval(u.y, db, vb, ab, jb);
        double* out = B.lat + rowo[g] + (int64_t)id * (kCandFields * npad) + k0;
        if (store) {
            *reinterpret_cast<double2*>(out) = make_double2(da, db);
            *reinterpret_cast<double2*>(out + npad) = make_double2(va, vb);
            *reinterpret_cast<double2*>(out + 2 * npad) = make_double2(aa, ab);
            *reinterpret_cast<double2*>(out + 3 * npad) = make_double2(ja, jb);
        }
        double acc = 0.0, amax = 0.0;
        int best = kNoBlocker;
        bool curv_bad = false;
        if (valid_a) {
            acc = fma(ja, ja, acc);
            if (kLimits) amax = fmax(amax, fabs(aa));
        }
        if (valid_b) {
            acc = fma(jb, jb, acc);
            if (kLimits) amax = fmax(amax, fabs(ab));
        }
        if (kXY) {
            const int f = g * npm + kk;
            const double2 px = *reinterpret_cast<const double2*>(fpx + f), py = *reinterpret_cast<const double2*>(fpy + f);
            const double2 nx = *reinterpret_cast<const double2*>(fnx + f), ny = *reinterpret_cast<const double2*>(fny + f);
            const double xa = __dadd_rn(px.x, __dmul_rn(nx.x, da)), xb = __dadd_rn(px.y, __dmul_rn(nx.y, db));   // pt + n * d,
            const double ya = __dadd_rn(py.x, __dmul_rn(ny.x, da)), yb = __dadd_rn(py.y, __dmul_rn(ny.y, db));   // unfused like the reference
            if (store) {
                *reinterpret_cast<double2*>(out + 4 * npad) = make_double2(xa, xb);
                *reinterpret_cast<double2*>(out + 5 * npad) = make_double2(ya, yb);
            }
            const int base = sub << shift;
            const unsigned gm = lpc == 16 ? 0xffffu : 0xffu;
            if (kLimits) {
                if (do_curv) {   // the whole horizon is in the group: the left neighbour's pair is all a lane needs
                    const double qxa = __shfl_up_sync(kFull, xa, 1), qya = __shfl_up_sync(kFull, ya, 1);
                    const double qxb = __shfl_up_sync(kFull, xb, 1), qyb = __shfl_up_sync(kFull, yb, 1);
                    bool bad = false;
                    if (act && k0 >= 2 && k0 <= N - 1) bad |= menger_exceeds(qxa, qya, qxb, qyb, xa, ya, k2max);
                    if (act && k0 >= 1 && k0 <= N - 2) bad |= menger_exceeds(qxb, qyb, xa, ya, xb, yb, k2max);
                    curv_bad = ((__ballot_sync(kFull, bad) >> base) & gm) != 0;
                }
            }
            // (an idle group evaluated its row-mate's candidate again: the same points, and valid_* are false there)
            if (kColl) {
                const int mid = (N - 1) >> 2;   // group 0 always carries a candidate
                if (kColl == 1)
                    best = collide_groups(xa, ya, xb, yb, valid_a, valid_b, mid, obs, obsf, B.n_obs, Rf, P.radius_sq, act, base, gm);
                if (kColl == 2)
                    best = collide_groups_predicted(xa, ya, xb, yb, valid_a, valid_b, mid, kk, tblp, B.n_obs_steps, cbs, B.n_obs, Rf,
                                                    P.radius_sq, act, base, gm);
            }
        }
        acc = group_sum(acc, lpc);
        if (kLimits) amax = group_max(amax, lpc);
        if (hl == 0 && act) {   // :165-167 / :174-176
            const double clong = rowd[2 * g], horizon = rowd[2 * g + 1];
            const double e = axd[id] - dd_target;
            const double clat = __dadd_rn(__dadd_rn(__dmul_rn(P.kj, acc), __dmul_rn(P.kt, horizon)), __dmul_rn(P.kd, __dmul_rn(e, e)));
            resc[c * 4] = clat;
            resc[c * 4 + 1] = follow ? 0.0 : clong;
            resc[c * 4 + 2] = follow ? clong : 0.0;
            resc[c * 4 + 3] = __dadd_rn(__dmul_rn(P.klat, clat), __dmul_rn(P.klong, clong));
            const bool kin = kLimits ? (rowi[4 * g + 3] && amax <= P.a_max) : true;
            resi[c * 2] = FRENET_CAND_VALID | (kin ? FRENET_CAND_KINEMATIC_OK : 0) | (curv_bad ? 0 : kCurvOkBit);
            resi[c * 2 + 1] = best;
        }
    }
    __syncthreads();
    FRENET_PHASE(5);
    // per-candidate results of the whole group, written coalesced; candidates of rows that do not exist get the neutral record
    for (int c = tid; c < Ga * nd; c += blockDim.x) {
        const bool usable = rowi[4 * (c / nd) + 2];
        const int64_t gc = cand0 + c;
#pragma unroll
        for (int f = 0; f < 4; ++f) B.cost[f * candF + gc] = usable ? resc[c * 4 + f] : 0.0;
        B.cand_flags[gc] = usable ? (uint8_t)resi[c * 2] : 0;
        if (do_curv) B.curv_ok[gc] = (usable && (resi[c * 2] & kCurvOkBit)) ? 1 : 0;
        if (kColl) {
            const int best = usable ? resi[c * 2 + 1] : 0;
            B.coll_free[gc] = (usable && best == kNoBlocker) ? 1 : 0;
            B.first_blocker[gc] = (!usable || best == kNoBlocker) ? -1 : best;
        }
    }
}

template <bool kXY, int kColl, bool kLimits, bool kMono = false>
__global__ void __launch_bounds__(kRowThreads, kMono ? 1 : (kColl ? FRENET_K2_MIN_CTAS_COLL : FRENET_K2_MIN_CTAS))
k2_short(FrenetLattice L, FrenetParams P, FrenetPath path, FrenetBuffers B, int G, int opts, int use_mask, int gather_sel, int has_xy, FrenetInline inl) {
    const int R = L.n_v_cap * L.n_t;
    const int groups = (R + G - 1) / G;
    FRENET_PHASE(0);
    if (kMono) {
        spill_inline(inl, L, B);
        FRENET_PHASE(1);
        const int b = blockIdx.x / groups, row0 = (blockIdx.x % groups) * G;
        const int cells = min(G, R - row0) * L.n_d;
        const int64_t cand0 = ((int64_t)b * R + row0) * L.n_d;
        for (int i = threadIdx.x; i < cells; i += blockDim.x) k1_cell(L, P, B, cand0 + i);
        __syncthreads();
    }
    FRENET_PHASE(2);
    k2_short_body<kXY, kColl, kLimits>(L, P, path, B, G, opts);
    FRENET_PHASE(6);
    if (kMono) {
        const int b = blockIdx.x / groups, g = blockIdx.x % groups;
        if (skipped(B, b)) return;
        mono_tail(L, P, B, b, g, groups, g * G * L.n_d, min(G, R - g * G) * L.n_d, 1, use_mask, gather_sel, has_xy);
    }
}

// ------------------------------------------------------------------------------------------------
// K3: Frenet -> Cartesian, one CTA per row; the path frame is computed once per (row, step)
// ------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(kRowThreads) k3_cartesian(FrenetLattice L, FrenetPath path, FrenetBuffers B, int parts) {
    extern __shared__ __align__(16) double sm[];
    const int npm = L.n_pad_max;
    double* fpx = sm;                  // [n_pad_max] x 4: path point and unit normal per step
    double* fpy = fpx + npm;
    double* fnx = fpy + npm;
    double* fny = fnx + npm;
    double* tab = fny + npm;           // [n_knots][9] spline table staged in shared memory

    const int part = blockIdx.y;   // small batches: a row's candidates are split over `parts` = gridDim.y CTAs
    const int64_t br = blockIdx.x;
    const int flags = B.row_flags[br];
    if (!(flags & FRENET_ROW_EXISTS) || (flags & FRENET_ROW_DROPPED)) return;
    const RowAddr A = row_addr(L, br);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;

    if (path.kind == FRENET_PATH_SPLINE) {
        for (int i = tid; i < path.n_knots * 9; i += blockDim.x) tab[i] = path.spline_table[i];
        __syncthreads();
    }
    const double* s_row = B.lon + A.lon_off;
    const double* lane_q = B.lane_params ? B.lane_params + 4 * (int64_t)A.b : nullptr;
    const bool relative = B.path_origin != nullptr;    // Duckietown runs: path evaluated at projection + (s - s0[0])
    const double origin = relative ? B.path_origin[2 * A.b] : 0.0, s_start = relative ? B.path_origin[2 * A.b + 1] : 0.0;
    for (int k = tid; k < A.npad; k += blockDim.x) {   // padding samples included: whole sectors are written (see K2)
        double px, py, nx, ny;
        const double sp = relative ? origin + (s_row[k] - s_start) : s_row[k];
        path_frame(path.kind, path.circle, tab, path.n_knots, sp, px, py, nx, ny, lane_q);
        fpx[k] = px; fpy[k] = py; fnx[k] = nx; fny[k] = ny;
    }
    __syncthreads();
    // A warp works on TWO candidates at a time and every lane on pairs of steps: four 16-byte loads are in flight per
    // lane before the first use, which is what hides the HBM latency of this read-modify-write pass.
    const int npad = A.npad;
    constexpr int kSpan = 2 * kWarp;   // steps covered by one 16-byte access of a warp
    for (int id = warp + 2 * nwarps * part; id < L.n_d; id += 2 * nwarps * parts) {
        const bool two = id + nwarps < L.n_d;
        double* blk1 = B.lat + A.cand_off + (int64_t)id * (kCandFields * npad);
        double* blk2 = two ? blk1 + (int64_t)nwarps * (kCandFields * npad) : blk1;
        for (int kb = 0; kb < npad; kb += 2 * kSpan) {
            double2 d1[2], d2[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int k = kb + q * kSpan + 2 * lane;
                const bool in = k < npad;
                d1[q] = in ? *reinterpret_cast<const double2*>(blk1 + k) : make_double2(0.0, 0.0);
                d2[q] = (in && two) ? *reinterpret_cast<const double2*>(blk2 + k) : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int k = kb + q * kSpan + 2 * lane;
                if (k < npad) {
                    const double2 px = *reinterpret_cast<const double2*>(fpx + k), py = *reinterpret_cast<const double2*>(fpy + k);
                    const double2 nx = *reinterpret_cast<const double2*>(fnx + k), ny = *reinterpret_cast<const double2*>(fny + k);
                    // pt + n * d, unfused like the reference
                    *reinterpret_cast<double2*>(blk1 + 4 * npad + k) =
                        make_double2(__dadd_rn(px.x, __dmul_rn(nx.x, d1[q].x)), __dadd_rn(px.y, __dmul_rn(nx.y, d1[q].y)));
                    *reinterpret_cast<double2*>(blk1 + 5 * npad + k) =
                        make_double2(__dadd_rn(py.x, __dmul_rn(ny.x, d1[q].x)), __dadd_rn(py.y, __dmul_rn(ny.y, d1[q].y)));
                    if (two) {
                        *reinterpret_cast<double2*>(blk2 + 4 * npad + k) =
                            make_double2(__dadd_rn(px.x, __dmul_rn(nx.x, d2[q].x)), __dadd_rn(px.y, __dmul_rn(nx.y, d2[q].y)));
                        *reinterpret_cast<double2*>(blk2 + 5 * npad + k) =
                            make_double2(__dadd_rn(py.x, __dmul_rn(ny.x, d2[q].x)), __dadd_rn(py.y, __dmul_rn(ny.y, d2[q].y)));
                    }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256) points_to_cartesian(FrenetPath path, const double* __restrict__ s, const double* __restrict__ d,
                                                           int64_t n, double* x, double* y, int64_t out_stride) {
    extern __shared__ __align__(16) double tab[];
    if (path.kind == FRENET_PATH_SPLINE) {
        for (int i = threadIdx.x; i < path.n_knots * 9; i += blockDim.x) tab[i] = path.spline_table[i];
        __syncthreads();
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double px, py, nx, ny;
        path_frame(path.kind, path.circle, tab, path.n_knots, s[i], px, py, nx, ny);
        const double di = d[i];
        x[i * out_stride] = __dadd_rn(px, __dmul_rn(nx, di));
        y[i * out_stride] = __dadd_rn(py, __dmul_rn(ny, di));
    }
}

// EXTENSION (no reference code): discrete curvature of the Cartesian polyline at interior samples,
// kappa = 2 |a x b| / (|a| |b| |a + b|), compared squared and division-free (oracle: curvature_feasible).
__global__ void __launch_bounds__(kRowThreads) k3_curvature(FrenetLattice L, FrenetParams P, FrenetBuffers B) {
    const int64_t br = blockIdx.x;
    const RowAddr A = row_addr(L, br);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int flags = B.row_flags[br];
    if (!(flags & FRENET_ROW_EXISTS) || (flags & FRENET_ROW_DROPPED)) {
        for (int id = tid; id < L.n_d; id += blockDim.x) B.curv_ok[A.cand0 + id] = 0;
        return;
    }
    const double k2max = P.kappa_max * P.kappa_max;
    for (int id = warp; id < L.n_d; id += nwarps) {
        const double* x = B.lat + A.cand_off + (int64_t)id * (kCandFields * A.npad) + 4 * A.npad;
        const double* y = x + A.npad;
        bool bad = false;
        for (int k = 1 + lane; k + 1 < A.N; k += kWarp) {
            bad |= menger_exceeds(x[k - 1], y[k - 1], x[k], y[k], x[k + 1], y[k + 1], k2max);
        }
        bad = __any_sync(kFull, bad);
        if (lane == 0) B.curv_ok[A.cand0 + id] = bad ? 0 : 1;
    }
}

// check_offroad for both lane boundaries (lib/tests/test_mapper_planner_obstacles.py:62-70, 117-124): a candidate is off the
// road when any sample has (y - k) - a (x - h)^2 < 0 against the right boundary or > 0 against the left one.
__global__ void __launch_bounds__(kRowThreads) k3_offroad(FrenetLattice L, FrenetRoad road, FrenetBuffers B) {
    const int64_t br = blockIdx.x;
    const RowAddr A = row_addr(L, br);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int flags = B.row_flags[br];
    if (!(flags & FRENET_ROW_EXISTS) || (flags & FRENET_ROW_DROPPED)) {
        for (int id = tid; id < L.n_d; id += blockDim.x) B.road_ok[A.cand0 + id] = 0;
        return;
    }
    if (B.road_params) {   // per-scenario boundaries: [n_scen][8] = right (h, k, a, given), left (h, k, a, given)
        const double* q = B.road_params + 8 * (int64_t)A.b;
        road.right[0] = q[0]; road.right[1] = q[1]; road.right[2] = q[2]; road.has_right = q[3] != 0.0;
        road.left[0] = q[4]; road.left[1] = q[5]; road.left[2] = q[6]; road.has_left = q[7] != 0.0;
    }
    for (int id = warp; id < L.n_d; id += nwarps) {
        const double* x = B.lat + A.cand_off + (int64_t)id * (kCandFields * A.npad) + 4 * A.npad;
        const double* y = x + A.npad;
        bool off = false;
        for (int k = lane; k < A.N; k += kWarp) {
            const double xv = x[k], yv = y[k];
            if (road.has_right) {
                const double u = xv - road.right[0];
                off |= __dadd_rn(yv - road.right[1], -__dmul_rn(road.right[2], __dmul_rn(u, u))) < 0.0;
            }
            if (road.has_left) {
                const double u = xv - road.left[0];
                off |= __dadd_rn(yv - road.left[1], -__dmul_rn(road.left[2], __dmul_rn(u, u))) > 0.0;
            }
        }
        off = __any_sync(kFull, off);
        if (lane == 0) B.road_ok[A.cand0 + id] = off ? 0 : 1;
    }
}

// ------------------------------------------------------------------------------------------------
// K4: collision, one CTA per row; obstacles staged in shared memory
// ------------------------------------------------------------------------------------------------
// kMono: FRENET_STAGE_ONE_LAUNCH's second form for SHORT-horizon lattices - the sweep over the stored x, y as below, then the
// partial-argmin / ticket / gather tail of the one-launch replan (mono_tail), so that check_paths is one kernel for every lattice.
template <bool kMono>
__device__ __forceinline__ void k4_collision_row(const FrenetLattice& L, const FrenetParams& P, const FrenetBuffers& B, int parts) {
    extern __shared__ __align__(16) double sm[];
    const int O = B.n_obs;
    double2* obs = reinterpret_cast<double2*>(sm);
    float2* obsf = reinterpret_cast<float2*>(sm + 2 * O);

    const int part = blockIdx.y;   // small batches: a row's candidates are split over `parts` = gridDim.y CTAs
    const int64_t br = blockIdx.x;
    const RowAddr A = row_addr(L, br);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int flags = B.row_flags[br];
    if (!(flags & FRENET_ROW_EXISTS) || (flags & FRENET_ROW_DROPPED)) {
        for (int id = tid; id < L.n_d; id += blockDim.x) {
            B.coll_free[A.cand0 + id] = 0;
            B.first_blocker[A.cand0 + id] = -1;
        }
        return;
    }
    // Obstacles the sensor did not report are parked at infinity so the cull never selects them
    // (check_paths tests sensed obstacles only).
    stage_obstacles(B, A.b, obs, obsf);
    __syncthreads();
    const float Rf = __double2float_ru(sqrt(P.radius_sq));

    for (int id = warp + nwarps * part; id < L.n_d; id += nwarps * parts) {
        const double* xs = B.lat + A.cand_off + (int64_t)id * (kCandFields * A.npad) + 4 * A.npad;
        const double* ys = xs + A.npad;
        int best = kNoBlocker;   // lowest blocking obstacle index found so far (warp-uniform)
        // two steps per lane, two 64-step spans loaded before the first test (latency hiding)
        for (int kb = 0; kb < A.N; kb += 4 * kWarp) {
            double2 xv[2], yv[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int k0 = kb + q * 2 * kWarp;
                const int k = k0 + 2 * lane;
                const int kk = k < A.npad ? k : (k0 < A.npad ? k0 : kb);   // npad is even: a pair inside the padding is readable
                xv[q] = *reinterpret_cast<const double2*>(xs + kk);
                yv[q] = *reinterpret_cast<const double2*>(ys + kk);
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int k0 = kb + q * 2 * kWarp;
                const int k = k0 + 2 * lane;
                if (k0 < A.N)
                    best = collide_pair(xv[q].x, yv[q].x, xv[q].y, yv[q].y, k < A.N, k + 1 < A.N, min(2 * kWarp, A.N - k0), obs, obsf, O,
                                        Rf, P.radius_sq, best, lane);
            }
        }
        if (lane == 0) {
            B.coll_free[A.cand0 + id] = best == kNoBlocker ? 1 : 0;
            B.first_blocker[A.cand0 + id] = best == kNoBlocker ? -1 : best;
        }
    }
}

template <bool kMono>
__global__ void __launch_bounds__(kRowThreads) k4_collision(FrenetLattice L, FrenetParams P, FrenetBuffers B, int parts, int use_mask, int gather_sel,
                                                            int has_xy) {
    k4_collision_row<kMono>(L, P, B, parts);
    if (kMono) {
        const int R = L.n_v_cap * L.n_t, nwarps = blockDim.x >> 5, part = blockIdx.y;
        const int b = (int)(blockIdx.x / R), r = (int)(blockIdx.x % R);
        if (skipped(B, b)) return;
        // this CTA's candidates of row r: runs of nwarps consecutive indices starting at nwarps * part, nwarps * parts apart
        int count = 0;
        for (int id0 = nwarps * part; id0 < L.n_d; id0 += nwarps * parts) count += min(nwarps, L.n_d - id0);
        // (the runs are complete except possibly the last one, which mono_tail's index formula covers: i / run, i % run)
        mono_tail(L, P, B, b, r * parts + part, R * parts, r * L.n_d + nwarps * part, count, nwarps * parts, use_mask, gather_sel, has_xy, nwarps);
    }
}

__global__ void __launch_bounds__(kRowThreads) k4_collision_predicted(FrenetLattice L, FrenetParams P, FrenetBuffers B) {
    extern __shared__ __align__(16) double sm[];
    const int O = B.n_obs, K = B.n_obs_steps;
    const int n_chunks = (L.n_pad_max + kWarp - 1) / kWarp;
    float4* cb = reinterpret_cast<float4*>(sm);   // [n_chunks][O]

    const int64_t br = blockIdx.x;
    const RowAddr A = row_addr(L, br);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int flags = B.row_flags[br];
    if (!(flags & FRENET_ROW_EXISTS) || (flags & FRENET_ROW_DROPPED)) {
        for (int id = tid; id < L.n_d; id += blockDim.x) {
            B.coll_free[A.cand0 + id] = 0;
            B.first_blocker[A.cand0 + id] = -1;
        }
        return;
    }
    const float4* src = reinterpret_cast<const float4*>(B.obs_bounds) + (int64_t)A.b * n_chunks * O;
    for (int i = tid; i < n_chunks * O; i += blockDim.x) cb[i] = src[i];
    __syncthreads();
    const double2* tbl = reinterpret_cast<const double2*>(B.obs_xy_t) + (int64_t)A.b * O * K;
    const float Rf = __double2float_ru(sqrt(P.radius_sq));

    for (int id = warp; id < L.n_d; id += nwarps) {
        const double* xs = B.lat + A.cand_off + (int64_t)id * (kCandFields * A.npad) + 4 * A.npad;
        const double* ys = xs + A.npad;
        int best = kNoBlocker;
        for (int kb = 0; kb < A.N; kb += kWarp) {
            const int k = kb + lane < A.N ? kb + lane : kb;
            best = collide_chunk_predicted(xs[k], ys[k], kb + lane < A.N, min(kWarp, A.N - kb), k, tbl, K, cb + (kb / kWarp) * O, O, Rf,
                                           P.radius_sq, best, lane);
        }
        if (lane == 0) {
            B.coll_free[A.cand0 + id] = best == kNoBlocker ? 1 : 0;
            B.first_blocker[A.cand0 + id] = best == kNoBlocker ? -1 : best;
        }
    }
}

// Device-side obstacle prediction from the reference's time law (lib/sensor/dynamic_obstacle.py:43-56).
__global__ void __launch_bounds__(256) predict_obstacles(FrenetPath path, const double* __restrict__ obs_s, const double* __restrict__ obs_d,
                                                         const double* __restrict__ obs_speed, const double* __restrict__ t0,
                                                         int64_t t0_stride, int n_scen, int O, int K, double period,
                                                         double* __restrict__ out) {
    extern __shared__ __align__(16) double tab[];
    if (path.kind == FRENET_PATH_SPLINE) {
        for (int i = threadIdx.x; i < path.n_knots * 9; i += blockDim.x) tab[i] = path.spline_table[i];
        __syncthreads();
    }
    const int64_t total = (int64_t)n_scen * O * K;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % K);
        const int64_t bo = i / K;
        const int b = (int)(bo / O);
        const double t = t0[(int64_t)b * t0_stride] + (double)k * period;
        double wrapped = fmod(t * obs_speed[bo], 50.0);                  // time_offset + t * speed % 50 with Python's `%`:
        if (wrapped < 0.0) wrapped += 50.0;                              // the result takes the sign of the divisor
        const double tau = obs_s[bo] + wrapped;
        double px, py, nx, ny;
        path_frame(path.kind, path.circle, tab, path.n_knots, tau, px, py, nx, ny);
        const double d = obs_d[bo];
        out[2 * i] = __dadd_rn(px, __dmul_rn(nx, d));
        out[2 * i + 1] = __dadd_rn(py, __dmul_rn(ny, d));
    }
}

// replan_ctot's bookkeeping on the device (trajectory_planner.py:183-195, 206-212, 120), one thread per scenario.
__global__ void __launch_bounds__(128) advance_state(FrenetLattice L, FrenetBuffers B, const double* __restrict__ time, int64_t time_stride,
                                                     double lookup_period, double q, int num_sample, int follow_all, int32_t* __restrict__ status) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= L.n_scen) return;
    if (skipped(B, b)) return;
    const int n = B.winner_steps[b];
    if (n <= 0) {
        if (status) status[b] = 1;
        return;
    }
    const double* w = B.winner + (int64_t)b * FRENET_WINNER_ROWS * L.n_pad_max;
    const double t = time[(int64_t)b * time_stride];
    int k;
    if (t <= w[0]) k = 0;
    else if (t >= w[n - 1]) k = n - 1;
    else {
        k = (int)rint((t - w[0]) / lookup_period);      // Python round(): half to even
        k = k < 0 ? k + n : k;                          // (a negative index would wrap like a Python list)
        k = min(max(k, 0), n - 1);
    }
    double* st = const_cast<double*>(B.state) + (int64_t)b * 6;
    const int npm = L.n_pad_max;
    st[0] = w[1 * npm + k]; st[1] = w[2 * npm + k]; st[2] = w[3 * npm + k];     // d, dd, ddd
    st[3] = w[5 * npm + k];
    const double ds0 = w[6 * npm + k];
    st[4] = ds0;
    st[5] = w[7 * npm + k];                                                       // s, ds, dds
    const_cast<double*>(B.scal)[(int64_t)b * 3 + 2] = t;
    // np.arange(round(ds0 - q n), round(ds0 + q n + q), q): length ceil((stop - start) / step), values start + i * ((start + step) - start)
    // follow mode: the axis holds the target OFFSETS of si_interval, centred on 0 (trajectory_planner.py:119)
    const double centre = (B.follow ? B.follow[b] != 0 : follow_all != 0) ? 0.0 : ds0;
    const double lo = rint(centre - q * num_sample), hi = rint(centre + q * num_sample + q);
    int nv = (int)ceil((hi - lo) / q);
    const bool axis_overflow = nv > L.n_v_cap;      // the host path raises here; reported as status 2 (the axis is truncated)
    nv = min(max(nv, 0), L.n_v_cap);
    const double delta = (lo + q) - lo;
    double* av = const_cast<double*>(L.axis_v) + (int64_t)b * L.n_v_cap;
    for (int i = 0; i < L.n_v_cap; ++i) av[i] = lo + (double)i * delta;
    const_cast<int32_t*>(L.n_v)[b] = nv;
    if (status) status[b] = axis_overflow ? 2 : 0;
}

// ------------------------------------------------------------------------------------------------
// Robot side of the closed loop (per-agent scalar work)
// ------------------------------------------------------------------------------------------------
// Path point and first derivative as the host classes give them (circle: forward chord over dt).
__device__ __forceinline__ void path_pt_grad(int kind, const double* __restrict__ circle, const double* __restrict__ tab, int n_knots,
                                             double s, double& px, double& py, double& gx, double& gy) {
    if (kind == FRENET_PATH_CIRCLE) {
        circle_pt(circle, s, px, py);
        double qx, qy;
        const double h = circle[3];
        circle_pt(circle, s + h, qx, qy);
        gx = (qx - px) / h;
        gy = (qy - py) / h;
    } else if (kind == FRENET_PATH_PARABOLA) {   // analytic derivative (mapper_obstacles.py:174)
        px = s;
        py = parabola_y(circle, s);
        gx = 1.0;
        gy = __dadd_rn(__dmul_rn(__dmul_rn(2.0, circle[0]), s), circle[1]);
    } else {
        spline_eval(tab, n_knots, s, px, py, gx, gy);
    }
}

// trajectory.compute_curvature (circle_trajectory.py:88-117: 1/r on the four arcs; spline_trajectory.py:173-181)
__device__ __forceinline__ double path_curvature(int kind, const double* __restrict__ c, const double* __restrict__ tab, int n, double t) {
    if (kind == FRENET_PATH_CIRCLE) {
        const double r = c[2], period = c[4];
        const double* b = c + 5;
        if (t >= period) t = fmod(t, period);
        const bool straight = (t >= 0.0 && t < b[0]) || (t >= b[1] && t < b[2]) || (t >= b[3] && t < b[4]) || (t >= b[5] && t < b[6]);
        return straight ? 0.0 : 1.0 / r;
    }
    if (kind == FRENET_PATH_PARABOLA) {   // the mapper's finite-difference form, verbatim (mapper_obstacles.py:176-184)
        const double dt = 1.0 / 30.0;
        const double x0 = t - dt, x1 = t, x2 = t + dt;
        const double y0 = parabola_y(c, x0), y1 = parabola_y(c, x1), y2 = parabola_y(c, x2);
        const double t1 = atan2(sin(y1 - y0), cos(x1 - x0));
        const double t2 = atan2(sin(y2 - y1), cos(x2 - x1));
        const double ex = x1 - x0, ey = y1 - y0;
        return fabs(t2 - t1) / sqrt(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)));
    }
    const double first = tab[0], last = tab[(n - 1) * 9];
    double dx, dy, ddx, ddy;
    if (t < first) { dx = dy = ddx = ddy = first; }
    else if (t >= last) { dx = dy = ddx = ddy = last; }
    else {
        int lo = 0, hi = n - 1;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (tab[mid * 9] <= t) lo = mid; else hi = mid;
        }
        const double* q = tab + lo * 9;
        const double u = t - q[0];
        dx = q[2] + 2.0 * q[3] * u + 3.0 * q[4] * (u * u);
        dy = q[6] + 2.0 * q[7] * u + 3.0 * q[8] * (u * u);
        ddx = 2.0 * q[3] + 6.0 * q[4] * u;
        ddy = 2.0 * q[7] + 6.0 * q[8] * u;
    }
    return (ddy * dx - ddx * dy) / (dx * dx + dy * dy);
}

__global__ void __launch_bounds__(128) sim_project(FrenetPath path, const double* __restrict__ pose, double* __restrict__ estimate,
                                                   int max_iter, double damping, double* __restrict__ fpose,
                                                   double* __restrict__ curvature, int n) {
    extern __shared__ __align__(16) double tab[];
    if (path.kind == FRENET_PATH_SPLINE) {
        for (int i = threadIdx.x; i < path.n_knots * 9; i += blockDim.x) tab[i] = path.spline_table[i];
        __syncthreads();
    }
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const double x = pose[3 * a], y = pose[3 * a + 1], th = pose[3 * a + 2];
    double s = estimate[a];
    double px, py, gx, gy;
    for (int it = 0; it < max_iter; ++it) {   // do_ls, gn_transform_old.py:93-106
        path_pt_grad(path.kind, path.circle, tab, path.n_knots, s, px, py, gx, gy);
        const double ex = px - x, ey = py - y;
        const double H = __dadd_rn(__dadd_rn(__dmul_rn(gx, gx), __dmul_rn(gy, gy)), damping);
        const double bb = __dadd_rn(__dmul_rn(gx, ex), __dmul_rn(gy, ey));
        s += -bb / H;
    }
    estimate[a] = s;
    path_pt_grad(path.kind, path.circle, tab, path.n_knots, s, px, py, gx, gy);
    const double tr = atan2(gy, gx);
    double sn, cs;
    sincos(tr, &sn, &cs);
    // R^T p - R^T t with R = [[c, -s], [s, c]] (transform, :58-68)
    const double fx = __dadd_rn(__dadd_rn(__dmul_rn(cs, x), __dmul_rn(sn, y)), -__dadd_rn(__dmul_rn(cs, px), __dmul_rn(sn, py)));
    const double fy = __dadd_rn(__dadd_rn(__dmul_rn(-sn, x), __dmul_rn(cs, y)), -__dadd_rn(__dmul_rn(-sn, px), __dmul_rn(cs, py)));
    fpose[3 * a] = fx;
    fpose[3 * a + 1] = fy;
    fpose[3 * a + 2] = th - tr;
    curvature[a] = path_curvature(path.kind, path.circle, tab, path.n_knots, s);
}

__global__ void __launch_bounds__(128) sim_control_step(const double* __restrict__ fpose, const double* __restrict__ state,
                                                        const double* __restrict__ curvature, double b, double kp_s, double kp_d, double dt,
                                                        double* __restrict__ pose, double* __restrict__ control, int n) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const double* st = state + 6 * a;              // p0, dp0, ddp0, s0, ds0, dds0: what replanner returned
    const double d = fpose[3 * a + 1], t = fpose[3 * a + 2], k = curvature[a];
    // error = [0, pos_d[0]] - robot_fpose[:2]; the controller zeroes the s component (ctrl_frenet_iol.py:97-98)
    const double pe0 = 0.0 * kp_s, pe1 = (st[0] - d) * kp_d;
    (void)pe0;
    const double r0 = st[4] + 0.0, r1 = st[1] + pe1;   // feed-forward (ds, dd) + proportional term
    double sn, cs;
    sincos(t, &sn, &cs);
    const double m00 = cs * (1.0 + b * k * sn) / (1.0 - k * d), m01 = -b * sn;
    const double m10 = sn - (k * cs) / (1.0 - k * d), m11 = b * cs;
    const double det = m00 * m11 - m01 * m10;
    const double v = (m11 * r0 - m01 * r1) / det, w = (-m10 * r0 + m00 * r1) / det;
    control[2 * a] = v;
    control[2 * a + 1] = w;
    // Unicycle.step: closed-form pose increment with the reference's truncated series (unicycle.py:24-28, 52-66)
    double* p = pose + 3 * a;
    const double drho = dt * v, dth = dt * w;
    const double t2 = dth * dth, t3 = t2 * dth, t4 = t2 * t2, t5 = t4 * dth, t6 = t3 * t3, t7 = t6 * dth;
    const double Sin = 1.0 - 1.0 / 6 * t2 + 1.0 / 120 * t4 - 1.0 / 5040 * t6;
    const double Cos = 1.0 / 2 * dth - 1.0 / 24 * t3 + 1.0 / 720 * t5 - 1.0 / 40320 * t7;
    const double dx = drho * Sin, dy = drho * Cos;
    double s0, c0, s1, c1;
    sincos(p[2], &s0, &c0);
    sincos(dth, &s1, &c1);
    const double nx = __dadd_rn(__dadd_rn(__dmul_rn(c0, dx), __dmul_rn(-s0, dy)), p[0]);
    const double ny = __dadd_rn(__dadd_rn(__dmul_rn(s0, dx), __dmul_rn(c0, dy)), p[1]);
    const double nth = atan2(__dadd_rn(__dmul_rn(s0, c1), __dmul_rn(c0, s1)), __dadd_rn(__dmul_rn(c0, c1), __dmul_rn(-s0, s1)));
    const double dpt = nth - p[2];
    const double vx = (nx - p[0]) / dt, vy = (ny - p[1]) / dt, vt = atan2(sin(dpt), cos(dpt)) / dt;
    p[0] += dt * vx;     // self.p += dt * vel
    p[1] += dt * vy;
    p[2] += dt * vt;
}

// trajectory.compute_second_derivative as the host classes give it: the circle's is the DIFFERENCE of two forward-chord first
// derivatives h = 1e-5 apart, not divided by h (circle_trajectory.py:74-78); the spline's is calcdd per coordinate, which like
// calc / calcd returns the knot abscissa outside the knot range (spline_trajectory.py:75-88, 161-164).
__device__ __forceinline__ void path_second_derivative(int kind, const double* __restrict__ c, const double* __restrict__ tab, int n, double t,
                                                       double& ax, double& ay) {
    if (kind == FRENET_PATH_CIRCLE) {
        const double h = 1e-5;
        double px, py, g1x, g1y, g0x, g0y;
        path_pt_grad(kind, c, tab, n, t + h / 2, px, py, g1x, g1y);
        path_pt_grad(kind, c, tab, n, t - h / 2, px, py, g0x, g0y);
        ax = g1x - g0x;
        ay = g1y - g0y;
        return;
    }
    if (kind == FRENET_PATH_PARABOLA) {   // d2/dx2 (x, a x^2 + b x + c)
        ax = 0.0;
        ay = 2.0 * c[0];
        return;
    }
    const double first = tab[0], last = tab[(n - 1) * 9];
    if (t < first) { ax = ay = first; return; }
    if (t >= last) { ax = ay = last; return; }
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (tab[mid * 9] <= t) lo = mid; else hi = mid;
    }
    const double* q = tab + lo * 9;
    const double u = t - q[0];
    ax = 2.0 * q[3] + 6.0 * q[4] * u;
    ay = 2.0 * q[7] + 6.0 * q[8] * u;
}

// Follow-mode target triples of V1 (trajectory_planner.py:131-136) for every (scenario in follow mode, horizon), on the device:
//   time = t_initial + T;   st = s0 + tgt.compute_s(time) - d_target + delta_t * tgt.compute_ds(time)
//   dst = ds0 + tgt.compute_ds(time) - delta_t * tgt.compute_dds(time);   ddst = dds0 + tgt.compute_dds(time)
// with tgt = ObstacleTrajectory of the scenario's leading obstacle (lib/sensor/dynamic_obstacle.py:43-77): compute_s is the first
// component of the live transformer's `transform` of the obstacle's point at the time law, compute_ds / compute_dds the same of the
// PATH's first / second derivative at parameter `time` (sic).  The transformer is the robot's projection frame
// (FrenetGNTransformOld.estimatePosition, gn_transform_old.py:18-56), rebuilt here from the projection's arc length `estimate`.
__global__ void __launch_bounds__(128) follow_targets(FrenetPath path, FrenetLattice L, FrenetBuffers B, const int32_t* __restrict__ lead,
                                                      const double* __restrict__ obs_s, const double* __restrict__ obs_d,
                                                      const double* __restrict__ obs_speed, int O, const double* __restrict__ estimate,
                                                      double delta_t, double d_target, const uint8_t* __restrict__ only, double* __restrict__ target) {
    extern __shared__ __align__(16) double tab[];
    if (path.kind == FRENET_PATH_SPLINE) {
        for (int i = threadIdx.x; i < path.n_knots * 9; i += blockDim.x) tab[i] = path.spline_table[i];
        __syncthreads();
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.n_scen * L.n_t) return;
    const int b = i / L.n_t, iT = i % L.n_t;
    if ((only && !only[b]) || !(B.follow && B.follow[b])) return;
    const int o = lead[b];
    if (o < 0 || o >= O) return;
    const int64_t bo = (int64_t)b * O + o;
    const double time = __dadd_rn(B.scal[(int64_t)b * 3 + 2], L.axis_t[iT]);
    double wrapped = fmod(__dmul_rn(time, obs_speed[bo]), 50.0);
    if (wrapped < 0.0) wrapped += 50.0;
    const double tau = __dadd_rn(obs_s[bo], wrapped);
    double px, py, nx, ny;
    path_frame(path.kind, path.circle, tab, path.n_knots, tau, px, py, nx, ny);
    const double ox = __dadd_rn(px, __dmul_rn(nx, obs_d[bo])), oy = __dadd_rn(py, __dmul_rn(ny, obs_d[bo]));
    // the transformer's frame: point and heading of the path at the robot's projection
    double tx, ty, gx, gy;
    path_pt_grad(path.kind, path.circle, tab, path.n_knots, estimate[b], tx, ty, gx, gy);
    double sn, cs;
    sincos(atan2(gy, gx), &sn, &cs);
    const double anchor = __dadd_rn(__dmul_rn(cs, tx), __dmul_rn(sn, ty));           // (R^T t)[0]
    auto along = [&](double x, double y) { return __dadd_rn(__dadd_rn(__dmul_rn(cs, x), __dmul_rn(sn, y)), -anchor); };   // (R^T p - R^T t)[0]
    const double c_s = along(ox, oy);
    double qx, qy, d1x, d1y, d2x, d2y;
    path_pt_grad(path.kind, path.circle, tab, path.n_knots, time, qx, qy, d1x, d1y);
    path_second_derivative(path.kind, path.circle, tab, path.n_knots, time, d2x, d2y);
    const double c_ds = along(d1x, d1y), c_dds = along(d2x, d2y);
    const double* st = B.state + (int64_t)b * 6;
    double* out = target + ((int64_t)b * L.n_t + iT) * 3;
    out[0] = __dadd_rn(__dadd_rn(__dadd_rn(st[3], c_s), -d_target), __dmul_rn(delta_t, c_ds));
    out[1] = __dadd_rn(__dadd_rn(st[4], c_ds), -__dmul_rn(delta_t, c_dds));
    out[2] = __dadd_rn(st[5], c_dds);
}

// The all-blocked branch of the moving-obstacle driver (lib/tests/test_moving_obstacle_planner.py:92-97) for a batch: a scenario
// whose candidates ALL collide switches to follow mode behind leading_obstacles[0] - the first blocker of its first candidate -
// and is marked for the re-plan (`redo`); its longitudinal axis becomes si_interval.  Scenarios with survivors get redo = 0.
__global__ void __launch_bounds__(128) follow_select(FrenetLattice L, FrenetBuffers B, double q, int num_sample, int32_t* __restrict__ lead,
                                                     uint8_t* __restrict__ follow, uint8_t* __restrict__ redo) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= L.n_scen) return;
    const FrenetResult r = B.result[b];
    const bool blocked = r.n_valid > 0 && r.n_survivors == 0;
    redo[b] = blocked ? 1 : 0;
    if (!blocked) return;
    const int64_t C = (int64_t)L.n_v_cap * L.n_t * L.n_d;
    int first = -1;
    for (int64_t c = 0; c < C && first < 0; ++c)
        if (B.cand_flags[b * C + c] & FRENET_CAND_VALID) first = B.first_blocker[b * C + c];
    if (first < 0) {   // blocked by another filter than the obstacles: nobody to follow
        redo[b] = 0;
        return;
    }
    lead[b] = first;
    follow[b] = 1;
    const double lo = rint(-q * num_sample), hi = rint(q * num_sample + q);
    int nv = (int)ceil((hi - lo) / q);
    nv = min(max(nv, 0), L.n_v_cap);
    const double delta = (lo + q) - lo;
    double* av = const_cast<double*>(L.axis_v) + (int64_t)b * L.n_v_cap;
    for (int i = 0; i < L.n_v_cap; ++i) av[i] = lo + (double)i * delta;
    const_cast<int32_t*>(L.n_v)[b] = nv;
}

// Sampled targets of the Optimal variant (lib/optimal_frenet_frame/target.py:17-52) for a batch: the leading vehicle's quintic
// through (s0_t, s1_t) over Ts, sampled every DTs, turned into a follow / merge / stop target, then indexed at
// round((t_initial + T - 0) / delta_t) like trajectory_planner_optimal.py:130-133.  spec [n_scen][16]:
// s0[3], s1[3], Ts, DTs, mode (0 plain, 1 follow, 2 merge, 3 stop), D0, second target s0[3], s1[3] (merge).
// status[b]: 0 ok, 1 index outside the sampled target (the reference raises IndexError), 2 stop target not at rest (it asserts).
__device__ __forceinline__ void quintic_at(const double a[6], double t, double v[4]) {
    Poly p;
    p.load(a);
    const double tp[4] = {pow(t, 2.0), pow(t, 3.0), pow(t, 4.0), pow(t, 5.0)};
    eval_power_form(p, t, tp, v[0], v[1], v[2], v[3]);
}

__global__ void __launch_bounds__(128) optimal_targets(FrenetLattice L, FrenetBuffers B, const double* __restrict__ spec, double delta_t,
                                                       double* __restrict__ target, int32_t* __restrict__ status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.n_scen * L.n_t) return;
    const int b = i / L.n_t, iT = i % L.n_t;
    if (skipped(B, b)) return;
    const double* q = spec + 16 * (int64_t)b;
    const double Ts = q[6], DTs = q[7];
    const int mode = (int)q[8];
    double a[6], a2[6];
    quintic_bvp(q[0], q[1], q[2], q[3], q[4], q[5], Ts, a);
    const int n = (int)ceil((Ts + DTs) / DTs);        // len(np.arange(0, Ts + DTs, DTs))
    const int k = (int)rint((B.scal[(int64_t)b * 3 + 2] + L.axis_t[iT]) / delta_t);
    int st = 0;
    int kk = k < 0 ? k + n : k;                       // a negative index wraps like a Python list
    if (kk < 0 || kk >= n) { st = 1; kk = min(max(kk, 0), n - 1); }
    double v[4];
    quintic_at(a, (double)kk * DTs, v);
    double s = v[0], ds = v[1], dds = v[2];
    if (mode == 1) {          // get_frenet_follow_target
        bool constant_acc = true;
        double v0[4], vj[4];
        quintic_at(a, 0.0, v0);
        for (int j = 1; j < n && constant_acc; ++j) {
            quintic_at(a, (double)j * DTs, vj);
            constant_acc = vj[2] == v0[2];
        }
        s = __dadd_rn(__dadd_rn(v[0], -q[9]), __dmul_rn(DTs, v[1]));
        ds = __dadd_rn(v[1], -__dmul_rn(DTs, v[2]));
        if (!constant_acc) dds = __dadd_rn(v[2], -__dmul_rn(DTs, v[3]));
    } else if (mode == 2) {   // get_frenet_merge_target
        quintic_bvp(q[10], q[11], q[12], q[13], q[14], q[15], Ts, a2);
        double w[4];
        quintic_at(a2, (double)kk * DTs, w);
        s = __dmul_rn(0.5, __dadd_rn(v[0], w[0]));
    } else if (mode == 3) {   // get_frenet_stop_target asserts the target ends at rest
        double e[4];
        quintic_at(a, (double)(n - 1) * DTs, e);
        if (!(e[1] == 0.0 && e[2] == 0.0)) st = 2;
    }
    double* out = target + ((int64_t)b * L.n_t + iT) * 3;
    out[0] = s;
    out[1] = ds;
    out[2] = dds;
    if (status && st != 0) atomicMax(status + b, st);   // (the caller clears status first)
}

// Cone gate of the proximity sensor (lib/sensor/proximity_sensor.py:41-75), one thread per (scenario, obstacle).
__global__ void __launch_bounds__(256) sensor_gate(const double* __restrict__ pose, const double* __restrict__ obs_xy, int n_scen, int O,
                                                   double range, double aperture, uint8_t* __restrict__ active) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_scen * O) return;
    const int b = (int)(i / O);
    const double px = pose[3 * b], py = pose[3 * b + 1], th = pose[3 * b + 2];
    double sn, cs;
    sincos(th, &sn, &cs);
    const double rx = obs_xy[2 * i] - px, ry = obs_xy[2 * i + 1] - py;
    const double along = __dadd_rn(__dmul_rn(rx, cs), __dmul_rn(ry, sn));          // np.dot(diff, dir)
    bool on = !(along < 0.0 || along >= range);
    if (on) {
        const double cross = fabs(__dadd_rn(__dmul_rn(cs, ry), -__dmul_rn(sn, rx)));   // |dir x diff|
        const double dotp = __dadd_rn(__dmul_rn(cs, rx), __dmul_rn(sn, ry));           // np.dot(dir, diff)
        on = !(fabs(atan2(cross, dotp)) > aperture);
    }
    active[i] = on ? 1 : 0;
}

// Collision test of arbitrary polylines (one warp per path): the generic form of K4 for candidate lists that do not
// live in the lattice layout (e.g. `Frenet` objects built elsewhere and passed to the driver-level check_paths).
__global__ void __launch_bounds__(256) polyline_collision(const double* __restrict__ x, const double* __restrict__ y,
                                                          const int64_t* __restrict__ offsets, int n_paths,
                                                          const double* __restrict__ obs_xy, int O, double radius_sq,
                                                          uint8_t* coll_free, int32_t* first_blocker) {
    extern __shared__ __align__(16) double sm[];
    double2* obs = reinterpret_cast<double2*>(sm);
    float2* obsf = reinterpret_cast<float2*>(sm + 2 * O);
    for (int o = threadIdx.x; o < O; o += blockDim.x) {
        const double2 q = reinterpret_cast<const double2*>(obs_xy)[o];
        obs[o] = q;
        obsf[o] = make_float2(__double2float_rn(q.x), __double2float_rn(q.y));
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= n_paths) return;
    const int64_t lo = offsets[p];
    const int N = (int)(offsets[p + 1] - lo);
    const float Rf = __double2float_ru(sqrt(radius_sq));
    int best = kNoBlocker;
    for (int kb = 0; kb < N; kb += kWarp) {
        const int k = kb + lane < N ? kb + lane : kb;
        best = collide_chunk(x[lo + k], y[lo + k], kb + lane < N, min(kWarp, N - kb), obs, obsf, O, Rf, radius_sq, best, lane);
    }
    if (lane == 0) {
        coll_free[p] = best == kNoBlocker ? 1 : 0;
        first_blocker[p] = best == kNoBlocker ? -1 : best;
    }
}

// ------------------------------------------------------------------------------------------------
// K5: first-wins argmin, one CTA per scenario
// ------------------------------------------------------------------------------------------------
// The winner block of one scenario (the arrays `optimal_at_time` indexes, trajectory_planner.py:183-195), copied by all threads of the CTA.
__device__ __forceinline__ int winner_index(const FrenetResult& res, int sel) {
    return sel == FRENET_SEL_CTOT ? res.argmin_ctot
         : sel == FRENET_SEL_CD   ? res.argmin_cd
         : sel == FRENET_SEL_CV   ? res.argmin_cv
         : sel == FRENET_SEL_CT   ? res.argmin_ct
                                  : res.argmin_masked;
}

// kCg: loads bypass L1 (the one-launch replan reads what other CTAs of the SAME kernel wrote)
template <bool kCg, typename T>
__device__ __forceinline__ T ld_res(const T* p) {
    if (kCg) return __ldcg(p);
    return *p;
}

template <bool kCg>
__device__ __forceinline__ void gather_one(const FrenetLattice& L, const FrenetParams& P, const FrenetBuffers& B, int b, int c, int has_xy,
                                           double* winner_out, int32_t* steps_out, int tid, int nthreads) {
    if (tid < 0) { tid = threadIdx.x; nthreads = blockDim.x; }     // default: the whole CTA copies
    // (default: the buffers' winner block; the one-launch replan also fills buf->winner_masked through the explicit form)
    double* const wbase = winner_out ? winner_out : B.winner;
    int32_t* const steps = steps_out ? steps_out : B.winner_steps;
    if (c < 0) {
        if (tid == 0) steps[b] = 0;
        return;
    }
    const int r = c / L.n_d, id = c % L.n_d;
    const RowAddr A = row_addr(L, (int64_t)b * L.n_v_cap * L.n_t + r);
    double* w = wbase + (int64_t)b * FRENET_WINNER_ROWS * L.n_pad_max;
    const double t0 = B.scal[(int64_t)b * 3 + 2];
    const double* cand = B.lat + A.cand_off + (int64_t)id * (kCandFields * A.npad);
    const double* lon = B.lon + A.lon_off;
    if (L.lat_f32) {
        // fp32 fast path: the winner's lateral samples become the next replan's initial state, so they are re-evaluated in fp64 from
        // the candidate's coefficients (same synthetic division, same argument as the evaluate kernel); x, y come from the floats
        const float* candf = reinterpret_cast<const float*>(B.lat) + A.cand_off + (int64_t)id * (kCandFields * A.npad);
        Poly pq;
        pq.load(B.coef_lat + ((int64_t)b * L.n_v_cap * L.n_t * L.n_d + c) * 6);
        const bool low = B.row_flags[(int64_t)b * L.n_v_cap * L.n_t + r] & FRENET_ROW_LOWSPEED;
        const double s0 = B.state[(int64_t)b * 6 + 3];
        for (int k = tid; k < A.N; k += nthreads) {
            w[k] = __dadd_rn(__dmul_rn((double)k, P.sample_period), t0);   // two roundings, like the reference's arange()[k] + t_initial
            const double u = low ? fabs(lon[k] - s0) : (double)k * P.sample_period;
            double v0, v1, v2, v3;
            pq.eval(u, v0, v1, v2, v3);
            w[1 * L.n_pad_max + k] = v0; w[2 * L.n_pad_max + k] = v1; w[3 * L.n_pad_max + k] = v2; w[4 * L.n_pad_max + k] = v3;
#pragma unroll
            for (int f = 0; f < 4; ++f) w[(5 + f) * L.n_pad_max + k] = lon[f * A.npad + k];
            if (has_xy) {
                w[9 * L.n_pad_max + k] = (double)candf[4 * A.npad + k];
                w[10 * L.n_pad_max + k] = (double)candf[5 * A.npad + k];
            }
        }
    } else
    for (int k = tid; k < A.N; k += nthreads) {
        w[k] = __dadd_rn(__dmul_rn((double)k, P.sample_period), t0);   // f.t[i] += t_initial (:178-179): product and sum rounded separately
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            w[(1 + f) * L.n_pad_max + k] = ld_res<kCg>(cand + f * A.npad + k);
            w[(5 + f) * L.n_pad_max + k] = ld_res<kCg>(lon + f * A.npad + k);
        }
        if (has_xy) {
            w[9 * L.n_pad_max + k] = ld_res<kCg>(cand + 4 * A.npad + k);
            w[10 * L.n_pad_max + k] = ld_res<kCg>(cand + 5 * A.npad + k);
        }
    }
    if (tid < 4) w[11 * L.n_pad_max + tid] = ld_res<kCg>(B.cost + tid * A.candF + A.cand0 + id);
    if (tid == 0) steps[b] = A.N;
}

constexpr int kArgminThreads = 512;

// gather_sel >= 0: the CTA also copies that selection's winner block (gather_winner) once the record is known - one launch less
// in every pipeline that wants both.
template <bool kCg>
__device__ void k5_scenario(const FrenetLattice& L, const FrenetParams& P, const FrenetBuffers& B, int use_mask, int gather_sel, int has_xy, int b) {
    __shared__ int sh_winner;
    __shared__ double sv[kKeys][kArgminThreads / 32];
    __shared__ int si[kKeys][kArgminThreads / 32];
    __shared__ int scount[2][kArgminThreads / 32];
    const int C = L.n_v_cap * L.n_t * L.n_d;
    const int64_t base = (int64_t)b * C;
    const int64_t candF = (int64_t)L.n_scen * C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (skipped(B, b)) return;

    Best best[kKeys];
#pragma unroll
    for (int q = 0; q < kKeys; ++q) best[q].init();
    int nvalid = 0, nsurv = 0;
    // Every load of an iteration is issued before anything depends on one (a single planner is one CTA here: its latency is a chain
    // of memory round trips), and four candidates are in flight per thread.
    constexpr int kUnroll = 4;
    for (int c0 = tid; c0 < C; c0 += kUnroll * blockDim.x) {
        int fl[kUnroll];
        double cd[kUnroll], cv[kUnroll], ct[kUnroll], ctot[kUnroll];
        uint8_t m_curv[kUnroll], m_coll[kUnroll], m_road[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int c = c0 + u * blockDim.x;
            const bool in = c < C;
            const int64_t g = base + (in ? c : c0);
            fl[u] = in ? ld_res<kCg>(B.cand_flags + g) : 0;
            cd[u] = ld_res<kCg>(B.cost + g);
            cv[u] = ld_res<kCg>(B.cost + candF + g);
            ct[u] = ld_res<kCg>(B.cost + 2 * candF + g);
            ctot[u] = ld_res<kCg>(B.cost + 3 * candF + g);
            m_curv[u] = (use_mask & 2) ? ld_res<kCg>(B.curv_ok + g) : 1;
            m_coll[u] = (use_mask & 4) ? ld_res<kCg>(B.coll_free + g) : 1;
            m_road[u] = (use_mask & 8) ? ld_res<kCg>(B.road_ok + g) : 1;
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int c = c0 + u * blockDim.x;
            if (!(fl[u] & FRENET_CAND_VALID)) continue;
            ++nvalid;
            best[0].take(ctot[u], c);
            best[1].take(cd[u], c);
            best[2].take(cv[u], c);
            best[3].take(ct[u], c);
            bool pass = true;
            if (use_mask & 1) pass = pass && (fl[u] & FRENET_CAND_KINEMATIC_OK);
            pass = pass && m_curv[u] && m_coll[u] && m_road[u];
            if (pass) {
                ++nsurv;
                best[4].take(ctot[u], c);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < kKeys; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(kFull, best[q].v, o);
            const int oi = __shfl_xor_sync(kFull, best[q].i, o);
            best[q].take(ov, oi);
        }
        if (lane == 0) { sv[q][warp] = best[q].v; si[q][warp] = best[q].i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        nvalid += __shfl_xor_sync(kFull, nvalid, o);
        nsurv += __shfl_xor_sync(kFull, nsurv, o);
    }
    if (lane == 0) { scount[0][warp] = nvalid; scount[1][warp] = nsurv; }
    __syncthreads();
    if (tid == 0) {
        const int nw = blockDim.x >> 5;
        Best out[kKeys];
        int tv = 0, ts = 0;
        for (int q = 0; q < kKeys; ++q) {
            out[q].init();
            for (int w = 0; w < nw; ++w) out[q].take(sv[q][w], si[q][w]);
        }
        for (int w = 0; w < nw; ++w) { tv += scount[0][w]; ts += scount[1][w]; }
        FrenetResult res;
        auto idx = [](const Best& x) { return x.i == 0x7fffffff ? -1 : x.i; };
        res.argmin_ctot = idx(out[0]);
        res.argmin_cd = idx(out[1]);
        res.argmin_cv = idx(out[2]);
        res.argmin_ct = idx(out[3]);
        res.argmin_masked = idx(out[4]);
        res.n_valid = tv;
        res.n_survivors = ts;
        res.status = tv == 0 ? FRENET_STATUS_EMPTY : (res.argmin_masked < 0 ? FRENET_STATUS_NO_FEASIBLE : FRENET_STATUS_OK);
        res.min_ctot = out[0].v;
        res.min_masked_ctot = out[4].v;
        res.reserved[0] = res.reserved[1] = 0.0;
        B.result[b] = res;
        sh_winner = gather_sel >= 0 ? winner_index(res, gather_sel) : -1;
    }
    if (gather_sel >= 0) {
        __syncthreads();
        gather_one<kCg>(L, P, B, b, sh_winner, has_xy);
    }
}

__global__ void __launch_bounds__(kArgminThreads) k5_argmin(FrenetLattice L, FrenetParams P, FrenetBuffers B, int use_mask, int gather_sel,
                                                            int has_xy) {
    k5_scenario<false>(L, P, B, use_mask, gather_sel, has_xy, blockIdx.x);
}

// Winner gather: the sampled state optimal_at_time indexes (trajectory_planner.py:183-195), 11 rows + a cost row.
__global__ void __launch_bounds__(128) gather_winner(FrenetLattice L, FrenetParams P, FrenetBuffers B, int sel, int has_xy) {
    const int b = blockIdx.x;
    if (skipped(B, b)) return;
    gather_one<false>(L, P, B, b, winner_index(B.result[b], sel), has_xy);
}

// ------------------------------------------------------------------------------------------------
// Host side of the C ABI
// ------------------------------------------------------------------------------------------------
int check_lattice(const FrenetLattice* L) {
    if (!L) return fail(FRENET_ERR_BAD_ARG, "lattice is NULL");
    if (L->n_scen <= 0 || L->n_v_cap <= 0 || L->n_t <= 0 || L->n_d <= 0 || L->n_pad_max <= 0)
        return fail(FRENET_ERR_BAD_ARG, "lattice sizes must be positive");
    if (!L->axis_d || !L->axis_t || !L->n_steps || !L->t_offset || !L->t_pow || !L->axis_v || !L->n_v)
        return fail(FRENET_ERR_BAD_ARG, "lattice axis pointer is NULL");
    if (L->row_elems <= 0 || L->cand_elems != L->row_elems * L->n_d || (L->row_elems & (L->lat_f32 ? 7 : 3)))
        return fail(FRENET_ERR_BAD_ARG, "inconsistent lattice strides");
    if (L->lat_f32 != 0 && L->lat_f32 != 1) return fail(FRENET_ERR_BAD_ARG, "lat_f32 must be 0 or 1");
    if ((int64_t)L->n_scen * L->n_v_cap * L->n_t > 0x7fffffffLL)
        return fail(FRENET_ERR_BAD_ARG, "too many rows for one launch");
    return FRENET_OK;
}

template <typename K>
int set_smem(K kernel, size_t bytes, const char* name) {
    if (bytes > (size_t)kMaxSmem) return fail(FRENET_ERR_CAPACITY, name);
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return fail(FRENET_ERR_CUDA, name, e);
    }
    return FRENET_OK;
}

inline unsigned n_rows(const FrenetLattice* L) { return (unsigned)((int64_t)L->n_scen * L->n_v_cap * L->n_t); }

int check_path(const FrenetPath* path) {
    if (!path) return fail(FRENET_ERR_BAD_ARG, "path is NULL");
    if (path->kind == FRENET_PATH_SPLINE) {
        if (path->n_knots < 2 || !path->spline_table) return fail(FRENET_ERR_BAD_ARG, "spline path needs >= 2 knots and a table");
    } else if (path->kind != FRENET_PATH_CIRCLE && path->kind != FRENET_PATH_PARABOLA) {
        return fail(FRENET_ERR_BAD_ARG, "unknown path kind");
    }
    return FRENET_OK;
}

// predicted mode: chunk granularity of the obstacle bounds the fused kernel uses (it follows the kernel's lane mapping)
inline int predicted_span(const FrenetLattice* L) { return L->n_pad_max > kWarp ? 2 * kWarp : kWarp; }
inline int predicted_chunks(const FrenetLattice* L) { return (L->n_pad_max + predicted_span(L) - 1) / predicted_span(L); }

int launch_obstacle_bounds(const FrenetLattice* L, const FrenetBuffers* B, int span, cudaStream_t stream) {
    if (!B->obs_xy_t || !B->obs_bounds) return fail(FRENET_ERR_BAD_ARG, "predicted obstacle table or bounds scratch missing");
    const int n_chunks = (L->n_pad_max + span - 1) / span;
    if (n_chunks == 0 || L->n_scen == 0 || B->n_obs == 0) return FRENET_OK;
    if ((int64_t)n_chunks * L->n_scen > 0x7fffffffLL) return fail(FRENET_ERR_BAD_ARG, "obstacle bounds: too many scenarios for one launch");
    obstacle_chunk_bounds<<<(unsigned)(n_chunks * L->n_scen), 128, 0, stream>>>(B->obs_xy_t, B->obs_active, B->n_obs, B->n_obs_steps, n_chunks, span,
                                                                         reinterpret_cast<float4*>(B->obs_bounds));
    FRENET_CHECK_LAUNCH("obstacle_chunk_bounds");
    return FRENET_OK;
}

// How the evaluate kernel is launched for a shape (shared by launch_k2 and frenet_k2_launch_shape).
struct K2Shape {
    bool pair;        // long horizons: k2_evaluate, one CTA per row (split over `parts` CTAs for small batches); else k2_short
    int rows_per_cta; // k2_short: consecutive rows of a scenario per CTA
    int parts;
    size_t smem;
    unsigned grid;
};

K2Shape k2_shape(const FrenetLattice* L, bool xy, int coll, int n_knots_spline, int n_obs) {
    K2Shape S;
    const size_t spline = (size_t)9 * n_knots_spline;
    S.pair = L->n_pad_max > kWarp;   // long horizons: two steps per lane
    if (!S.pair) {
        // short horizons: G consecutive rows of a scenario per CTA - as many as fit 48 KB of shared memory while the grid still
        // fills the GPU a few times over (a single planner keeps one row per CTA: its latency is what matters)
        const int R = L->n_v_cap * L->n_t;
        const size_t fixed = (size_t)L->n_d + spline + (coll == 1 ? (size_t)4 * ((3 * n_obs + 3) / 4) : 0) + (coll == 2 ? (size_t)2 * n_obs : 0);
        auto bytes = [&](int g) { return (fixed + (size_t)g * ((size_t)L->n_pad_max * (xy ? 5 : 1) + (size_t)L->n_d * 11 + 5)) * sizeof(double); };
        int G = R < kWarp ? R : kWarp;   // (phase B keeps one row per lane)
        while (G > 1 && (bytes(G) > 48 * 1024 || (int64_t)L->n_scen * ((R + G - 1) / G) < 8 * 148)) G = (G + 1) / 2;
        S.rows_per_cta = G;
        S.parts = 1;
        S.smem = bytes(G);
        S.grid = (unsigned)((int64_t)L->n_scen * ((R + G - 1) / G));
        return S;
    }
    S.rows_per_cta = 1;
    S.smem = ((size_t)L->n_pad_max * (xy ? 5 : 1) + (size_t)L->n_d * 12 + spline + (coll == 1 ? (size_t)4 * ((3 * n_obs + 3) / 4) : 0) +
              (coll == 1 && kRowTable ? row_table_doubles(L->n_pad_max, L->n_d, n_obs) : 0) +
              (FRENET_BULK_STORE ? (size_t)(kRowThreads / kWarp) * kCandFields * L->n_pad_max + 2 : 0) +
              (coll == 2 ? (size_t)2 * predicted_chunks(L) * n_obs : 0)) * sizeof(double);
    // a single planner (or a handful) has too few rows to fill the GPU with one CTA each: split the rows' candidates over several CTAs
    // - but only as far as all CTAs are resident at once (one wave): a CTA's life is mostly the latency of its row set-up, which a
    // second wave would pay again (config 4: 120 rows x 8 parts = 960 CTAs in two waves took 34 us, x 4 in one wave less)
    S.parts = 1;
    const int64_t slots = (int64_t)k2_min_ctas(coll) * 148;
    while (S.parts < FRENET_MAX_ROW_PARTS && (int64_t)n_rows(L) * (S.parts + 1) <= slots && S.parts * (kRowThreads / kWarp) < L->n_d) ++S.parts;
    S.grid = n_rows(L);
    return S;
}

// mono != nullptr: the one-launch replan - {use_mask, gather_sel, has_xy} of the K5 + gather tail (see mono_tail)
const FrenetInline kNoInline = {};

template <bool kXY, int kColl>
int launch_k2(const FrenetLattice* L, const FrenetParams* P, const FrenetPath* path, const FrenetBuffers* B, int opts, cudaStream_t stream,
              const int* mono = nullptr, const FrenetInline* inl_in = nullptr) {
    const FrenetInline& inl = inl_in ? *inl_in : kNoInline;
    // opts bit 0: also test the curvature limit in the kernel (needs kXY; selects the limits variant)
    const bool limits = !(isinf(P->v_max) && P->v_max > 0 && isinf(P->a_max) && P->a_max > 0) || (kXY && (opts & 1));
    const bool spline = kXY && path->kind == FRENET_PATH_SPLINE;
    const K2Shape S = k2_shape(L, kXY, kColl, spline ? path->n_knots : 0, B->n_obs);
    static const FrenetPath no_path = {};
    const FrenetPath& pd = kXY ? *path : no_path;
    if (S.grid == 0) return FRENET_OK;
    // robot radius in fp32, rounded up (the culls and the row table only need a conservative value)
    float radius_up = (float)sqrt(P->radius_sq);
    while ((double)radius_up * (double)radius_up < P->radius_sq) radius_up = nextafterf(radius_up, INFINITY);
    radius_up = nextafterf(radius_up, INFINITY);
    if (!S.pair && L->lat_f32) return fail(FRENET_ERR_BAD_ARG, "fp32 fast path: lattices of at most 32 padded samples are not supported");
    if (!S.pair) {
#define FRENET_LAUNCH_SHORT(LIM, MONO)                                                                                              \
    do {                                                                                                                            \
        if (int rc = set_smem(k2_short<kXY, kColl, LIM, MONO>, S.smem, "k2: row group does not fit shared memory")) return rc;      \
        k2_short<kXY, kColl, LIM, MONO><<<S.grid, kRowThreads, S.smem, stream>>>(*L, *P, pd, *B, S.rows_per_cta, opts,              \
                                                                                 mono ? mono[0] : 0, mono ? mono[1] : -1, mono ? mono[2] : 0, inl); \
    } while (0)
        if constexpr (kColl != 2) {
            if (mono && mono[3]) return fail(FRENET_ERR_BAD_ARG, "one launch: the collision-only form needs a long-horizon lattice (more than 32 padded samples)");
            if (mono) {
                if (limits) FRENET_LAUNCH_SHORT(true, true);
                else FRENET_LAUNCH_SHORT(false, true);
                FRENET_CHECK_LAUNCH("k2_short (one launch)");
                return FRENET_OK;
            }
        }
        if (limits) FRENET_LAUNCH_SHORT(true, false);
        else FRENET_LAUNCH_SHORT(false, false);
#undef FRENET_LAUNCH_SHORT
        FRENET_CHECK_LAUNCH("k2_short");
        return FRENET_OK;
    }
    if (L->lat_f32) {
        // fp32 fast path: the fused long-horizon kernel only (the batched dense-lattice workload it is meant for)
        if (!kXY || kColl == 2) return fail(FRENET_ERR_BAD_ARG, "fp32 fast path: only the fused K2+K3(+K4 snapshot) kernel stores floats");
        if constexpr (kXY && kColl != 2) {
            const dim3 grid1(S.grid, 1);
            if (limits) {
                if (int rc = set_smem(k2_evaluate<true, kColl, true, false, true>, S.smem, "k2: row does not fit shared memory")) return rc;
                k2_evaluate<true, kColl, true, false, true><<<grid1, kRowThreads, S.smem, stream>>>(*L, *P, pd, *B, opts, radius_up, 0, -1, 0, kNoInline);
            } else {
                if (int rc = set_smem(k2_evaluate<true, kColl, false, false, true>, S.smem, "k2: row does not fit shared memory")) return rc;
                k2_evaluate<true, kColl, false, false, true><<<grid1, kRowThreads, S.smem, stream>>>(*L, *P, pd, *B, opts, radius_up, 0, -1, 0, kNoInline);
            }
            FRENET_CHECK_LAUNCH("k2_evaluate (fp32 fast path)");
        }
        return FRENET_OK;
    }
    const dim3 grid(S.grid, S.parts);
#define FRENET_LAUNCH_K2(LIM, SPLIT, MONO)                                                                                                  \
    do {                                                                                                                                    \
        if (int rc = set_smem(k2_evaluate<kXY, kColl, LIM, SPLIT, false, MONO>, S.smem, "k2: row does not fit shared memory")) return rc;   \
        k2_evaluate<kXY, kColl, LIM, SPLIT, false, MONO><<<grid, kRowThreads, S.smem, stream>>>(*L, *P, pd, *B, opts, radius_up,            \
                                                                                                mono ? mono[0] : 0, mono ? mono[1] : -1, mono ? mono[2] : 0, inl); \
    } while (0)
    if constexpr (kColl != 2) {
        if constexpr (kXY && kColl == 1 && kRowTable) {
            if (mono && mono[3]) {   // collision stage alone on an existing lattice + the K5 / gather tail
                if (int rc = set_smem(k2_evaluate<true, 1, false, true, false, true, true>, S.smem, "k2: row does not fit shared memory")) return rc;
                k2_evaluate<true, 1, false, true, false, true, true><<<grid, kRowThreads, S.smem, stream>>>(*L, *P, pd, *B, opts, radius_up, mono[0], mono[1], mono[2], inl);
                FRENET_CHECK_LAUNCH("k2_evaluate (collision only, one launch)");
                return FRENET_OK;
            }
        }
        if (mono && mono[3]) return fail(FRENET_ERR_BAD_ARG, "one launch: the collision-only form is not available in this build / mode");
        if (mono) {     // (always the split form: gridDim.y may be 1)
            if (limits) FRENET_LAUNCH_K2(true, true, true);
            else FRENET_LAUNCH_K2(false, true, true);
            FRENET_CHECK_LAUNCH("k2_evaluate (one launch)");
            return FRENET_OK;
        }
    }
    if (limits && S.parts > 1) FRENET_LAUNCH_K2(true, true, false);
    else if (limits) FRENET_LAUNCH_K2(true, false, false);
    else if (S.parts > 1) FRENET_LAUNCH_K2(false, true, false);
    else FRENET_LAUNCH_K2(false, false, false);
#undef FRENET_LAUNCH_K2
    FRENET_CHECK_LAUNCH("k2_evaluate");
    return FRENET_OK;
}

// The standalone K4 sweep over the stored x, y; mono = {use_mask, gather_sel, has_xy}: with the one-launch tail (short-horizon lattices)
int launch_k4(const FrenetLattice* L, const FrenetParams* P, const FrenetBuffers* B, const int* mono, cudaStream_t stream) {
    const size_t smem = (size_t)3 * B->n_obs * sizeof(double);
    int parts = 1;
    while (parts < FRENET_MAX_ROW_PARTS && (int64_t)n_rows(L) * parts < 4 * 148 && parts * (kRowThreads / kWarp) < L->n_d) parts *= 2;
    if (mono) {
        if (int rc = set_smem(k4_collision<true>, smem, "k4: too many obstacles for shared memory")) return rc;
        k4_collision<true><<<dim3(n_rows(L), parts), kRowThreads, smem, stream>>>(*L, *P, *B, parts, mono[0], mono[1], mono[2]);
    } else {
        if (int rc = set_smem(k4_collision<false>, smem, "k4: too many obstacles for shared memory")) return rc;
        k4_collision<false><<<dim3(n_rows(L), parts), kRowThreads, smem, stream>>>(*L, *P, *B, parts, 0, -1, 0);
    }
    FRENET_CHECK_LAUNCH("k4_collision");
    return FRENET_OK;
}

int check_k2(const FrenetLattice* L, const FrenetParams* P, const FrenetBuffers* B) {
    if (int rc = check_lattice(L)) return rc;
    if (!P || !B || !B->state || !B->scal || !B->coef_lon || !B->coef_lat || !B->row_S || !B->row_flags || !B->lon || !B->lat ||
        !B->cost || !B->cand_flags)
        return fail(FRENET_ERR_BAD_ARG, "k2: NULL buffer");
    if ((P->follow || B->follow) && !B->target) return fail(FRENET_ERR_BAD_ARG, "k2: follow mode needs target triples");
    return FRENET_OK;
}

int launch_k5(const FrenetLattice* L, const FrenetParams* P, const FrenetBuffers* B, int32_t use_mask, int gather_sel, int has_xy,
              cudaStream_t stream) {
    if (int rc = check_lattice(L)) return rc;
    if (!B || !B->cost || !B->cand_flags || !B->result) return fail(FRENET_ERR_BAD_ARG, "k5: NULL buffer");
    if ((use_mask & 2) && !B->curv_ok) return fail(FRENET_ERR_BAD_ARG, "k5: curvature mask requested but curv_ok is NULL");
    if ((use_mask & 4) && !B->coll_free) return fail(FRENET_ERR_BAD_ARG, "k5: collision mask requested but coll_free is NULL");
    if ((use_mask & 8) && !B->road_ok) return fail(FRENET_ERR_BAD_ARG, "k5: road mask requested but road_ok is NULL");
    if (gather_sel >= 0 && (!P || !B->winner || !B->winner_steps || !B->lat || !B->lon || !B->scal))
        return fail(FRENET_ERR_BAD_ARG, "k5 + gather: NULL buffer");
    // small lattices (the reference's default has 320-384 candidates) leave most of a 512-thread CTA idle: a narrower CTA lets
    // more scenarios be resident per SM (the lexicographic minimum does not depend on the block size)
    const int64_t C = (int64_t)L->n_v_cap * L->n_t * L->n_d;
    const int threads = C <= 1024 ? 128 : kArgminThreads;
    static const FrenetParams no_params = {};
    k5_argmin<<<L->n_scen, threads, 0, stream>>>(*L, P ? *P : no_params, *B, use_mask, gather_sel, has_xy);
    FRENET_CHECK_LAUNCH("k5_argmin");
    return FRENET_OK;
}

}  // namespace

extern "C" {

int frenet_abi_version(void) { return FRENET_B200_ABI_VERSION; }
const char* frenet_last_error(void) { return g_err; }

int frenet_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int frenet_k1_coefficients(const FrenetLattice* L, const FrenetParams* P, const FrenetBuffers* B, frenet_stream_t stream) {
    if (int rc = check_lattice(L)) return rc;
    if (!P || !B || !B->state || !B->coef_lon || !B->coef_lat || !B->row_S || !B->row_flags)
        return fail(FRENET_ERR_BAD_ARG, "k1: NULL buffer");
    if ((P->follow || B->follow) && !B->target) return fail(FRENET_ERR_BAD_ARG, "k1: follow mode needs target triples");
    const int64_t total = (int64_t)L->n_scen * L->n_v_cap * L->n_t * L->n_d;
    const int64_t blocks = (total + 255) / 256;
    if (blocks > 0x7fffffffLL) return fail(FRENET_ERR_BAD_ARG, "k1: too many candidates for one launch");
    k1_coefficients<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*L, *P, *B);
    FRENET_CHECK_LAUNCH("k1_coefficients");
    return FRENET_OK;
}

int frenet_k2_evaluate(const FrenetLattice* L, const FrenetParams* P, const FrenetBuffers* B, frenet_stream_t stream) {
    if (int rc = check_k2(L, P, B)) return rc;
    return launch_k2<false, 0>(L, P, nullptr, B, 0, (cudaStream_t)stream);
}

int frenet_k2_fused(const FrenetLattice* L, const FrenetParams* P, const FrenetPath* path, const FrenetBuffers* B,
                    int32_t with_collision, frenet_stream_t stream) {
    if (int rc = check_k2(L, P, B)) return rc;
    if (int rc = check_path(path)) return rc;
    const int opts = (with_collision & FRENET_FUSE_CURVATURE) ? 1 : 0;
    if (opts && !B->curv_ok) return fail(FRENET_ERR_BAD_ARG, "fused: curvature output is NULL");
    if (!(with_collision & FRENET_FUSE_COLLISION)) return launch_k2<true, 0>(L, P, path, B, opts, (cudaStream_t)stream);
    if (!B->coll_free || !B->first_blocker) return fail(FRENET_ERR_BAD_ARG, "fused: collision output is NULL");
    if (B->n_obs < 0 || (B->n_obs > 0 && !B->obs_xy)) return fail(FRENET_ERR_BAD_ARG, "fused: obstacle table missing");
    if (B->n_obs_steps > 0 && B->n_obs > 0) {   // predicted obstacles: per-chunk bounds first, then the fused kernel reads the table
        if (int rc = launch_obstacle_bounds(L, B, predicted_span(L), (cudaStream_t)stream)) return rc;
        return launch_k2<true, 2>(L, P, path, B, opts, (cudaStream_t)stream);
    }
    return launch_k2<true, 1>(L, P, path, B, opts, (cudaStream_t)stream);
}

int frenet_k3_cartesian(const FrenetLattice* L, const FrenetParams* P, const FrenetPath* path, const FrenetBuffers* B,
                        frenet_stream_t stream) {
    (void)P;
    if (int rc = check_lattice(L)) return rc;
    if (int rc = check_path(path)) return rc;
    if (!B || !B->row_flags || !B->lon || !B->lat) return fail(FRENET_ERR_BAD_ARG, "k3: NULL buffer");
    if (L->lat_f32) return fail(FRENET_ERR_BAD_ARG, "k3: the standalone kernels read fp64 candidate blocks (fp32 fast path: use frenet_plan / frenet_k2_fused)");
    const size_t smem = ((size_t)4 * L->n_pad_max + (path->kind == FRENET_PATH_SPLINE ? (size_t)9 * path->n_knots : 0)) * sizeof(double);
    if (int rc = set_smem(k3_cartesian, smem, "k3: horizon / spline table too large for shared memory")) return rc;
    int parts = 1;
    while (parts < 8 && (int64_t)n_rows(L) * parts < 4 * 148 && parts * 2 * (kRowThreads / kWarp) < L->n_d) parts *= 2;
    k3_cartesian<<<dim3(n_rows(L), parts), kRowThreads, smem, (cudaStream_t)stream>>>(*L, *path, *B, parts);
    FRENET_CHECK_LAUNCH("k3_cartesian");
    return FRENET_OK;
}

int frenet_k3_curvature(const FrenetLattice* L, const FrenetParams* P, const FrenetBuffers* B, frenet_stream_t stream) {
    if (int rc = check_lattice(L)) return rc;
    if (!P || !B || !B->row_flags || !B->lat || !B->curv_ok) return fail(FRENET_ERR_BAD_ARG, "curvature: NULL buffer");
    if (L->lat_f32) return fail(FRENET_ERR_BAD_ARG, "curvature: the standalone kernels read fp64 candidate blocks (fp32 fast path: use frenet_plan / frenet_k2_fused)");
    k3_curvature<<<n_rows(L), kRowThreads, 0, (cudaStream_t)stream>>>(*L, *P, *B);
    FRENET_CHECK_LAUNCH("k3_curvature");
    return FRENET_OK;
}

int frenet_k4_collision(const FrenetLattice* L, const FrenetParams* P, const FrenetBuffers* B, frenet_stream_t stream) {
    if (int rc = check_lattice(L)) return rc;
    if (!P || !B || !B->row_flags || !B->lat || !B->coll_free || !B->first_blocker) return fail(FRENET_ERR_BAD_ARG, "k4: NULL buffer");
    if (L->lat_f32) return fail(FRENET_ERR_BAD_ARG, "k4: the standalone kernels read fp64 candidate blocks (fp32 fast path: use frenet_plan / frenet_k2_fused)");
    if (B->n_obs < 0 || B->n_obs_steps < 0) return fail(FRENET_ERR_BAD_ARG, "k4: negative obstacle count");
    if (B->n_obs_steps > 0 && B->n_obs > 0) {   // predicted table
        const int n_chunks = (L->n_pad_max + kWarp - 1) / kWarp;
        if (int rc = launch_obstacle_bounds(L, B, kWarp, (cudaStream_t)stream)) return rc;
        const size_t smem = (size_t)n_chunks * B->n_obs * sizeof(float4);
        if (int rc = set_smem(k4_collision_predicted, smem, "k4: too many obstacles for shared memory")) return rc;
        k4_collision_predicted<<<n_rows(L), kRowThreads, smem, (cudaStream_t)stream>>>(*L, *P, *B);
        FRENET_CHECK_LAUNCH("k4_collision_predicted");
        return FRENET_OK;
    }
    if (B->n_obs > 0 && !B->obs_xy) return fail(FRENET_ERR_BAD_ARG, "k4: obstacle table missing");
    return launch_k4(L, P, B, nullptr, (cudaStream_t)stream);
}

int frenet_k3_offroad(const FrenetLattice* L, const FrenetRoad* road, const FrenetBuffers* B, frenet_stream_t stream) {
    if (int rc = check_lattice(L)) return rc;
    if (!B || (!road && !B->road_params)) return fail(FRENET_ERR_BAD_ARG, "k3_offroad: NULL argument");
    static const FrenetRoad no_road = {};
    if (!road) road = &no_road;        // per-scenario boundaries come from buf->road_params
    if (!B->lat || !B->road_ok || !B->row_flags) return fail(FRENET_ERR_BAD_ARG, "k3_offroad: lat, road_ok and row_flags are required");
    if (L->lat_f32) return fail(FRENET_ERR_BAD_ARG, "k3_offroad: the standalone kernels read fp64 candidate blocks (fp32 fast path: use frenet_plan / frenet_k2_fused)");
    if (n_rows(L) == 0) return FRENET_OK;
    k3_offroad<<<n_rows(L), kRowThreads, 0, (cudaStream_t)stream>>>(*L, *road, *B);
    FRENET_CHECK_LAUNCH("k3_offroad");
    return FRENET_OK;
}

int frenet_k5_argmin(const FrenetLattice* L, const FrenetBuffers* B, int32_t use_mask, frenet_stream_t stream) {
    return launch_k5(L, nullptr, B, use_mask, -1, 0, (cudaStream_t)stream);
}

int frenet_gather_winner(const FrenetLattice* L, const FrenetParams* P, const FrenetBuffers* B, int32_t sel, int32_t has_xy,
                         frenet_stream_t stream) {
    if (int rc = check_lattice(L)) return rc;
    if (!P || !B || !B->result || !B->winner || !B->winner_steps || !B->lon || !B->lat || !B->scal)
        return fail(FRENET_ERR_BAD_ARG, "gather: NULL buffer");
    if (sel < FRENET_SEL_CTOT || sel > FRENET_SEL_MASKED) return fail(FRENET_ERR_BAD_ARG, "gather: bad selector");
    gather_winner<<<L->n_scen, 128, 0, (cudaStream_t)stream>>>(*L, *P, *B, sel, has_xy);
    FRENET_CHECK_LAUNCH("gather_winner");
    return FRENET_OK;
}

// FRENET_STAGE_ONE_LAUNCH (both forms); inl: inline inputs of frenet_replan_inline, or NULL
static int plan_one_launch(const FrenetLattice* L, const FrenetParams* P, const FrenetPath* path, const FrenetBuffers* B, int32_t stages,
                           int32_t use_mask, int32_t gather_sel, frenet_stream_t stream, const FrenetInline* inl) {
    int rc = FRENET_OK;
    const int need = FRENET_STAGE_K1 | FRENET_STAGE_K2 | FRENET_STAGE_K5 | FRENET_STAGE_GATHER;
    if ((stages & ~FRENET_STAGE_ONE_LAUNCH) == (FRENET_STAGE_K4 | FRENET_STAGE_K5 | FRENET_STAGE_GATHER)) {
        // the collision stage alone on a lattice that is already sampled, through the row table (no stored sample is read back),
        // with the argmin + gather tail: what check_paths needs when the obstacles arrive after the replan
        if ((rc = check_k2(L, P, B))) return rc;
        if (!B->tickets || !B->partials) return fail(FRENET_ERR_BAD_ARG, "one launch: buf->tickets / buf->partials is NULL");
        if (L->lat_f32) return fail(FRENET_ERR_BAD_ARG, "one launch: not available on the fp32 fast path");
        if (!B->result || !B->winner || !B->winner_steps || !B->coll_free || !B->first_blocker)
            return fail(FRENET_ERR_BAD_ARG, "one launch: NULL result / winner / collision buffer");
        if ((use_mask & 2) || (use_mask & 8)) return fail(FRENET_ERR_BAD_ARG, "one launch: curvature / road masks are not produced by this launch");
        if (B->n_obs < 0 || (B->n_obs > 0 && !B->obs_xy)) return fail(FRENET_ERR_BAD_ARG, "one launch: obstacle table missing");
        if (B->n_obs_steps > 0 && B->n_obs > 0) return fail(FRENET_ERR_BAD_ARG, "one launch: predicted obstacles are not supported");
        if (L->n_pad_max <= kWarp) {
            // short horizons: the sweep over the stored x, y (a few KB per row) with the same tail; no path needed
            const int mono3[3] = {use_mask, gather_sel, 1};
            return launch_k4(L, P, B, mono3, (cudaStream_t)stream);
        }
        if ((rc = check_path(path))) return rc;
        const int mono[4] = {use_mask, gather_sel, 1, 1};
        return launch_k2<true, 1>(L, P, path, B, 0, (cudaStream_t)stream, mono, inl);
    }
    if ((stages & need) != need || (stages & (FRENET_STAGE_CURVATURE | FRENET_STAGE_NO_FUSE | FRENET_STAGE_OFFROAD)) ||
        ((stages & FRENET_STAGE_K4) && !(stages & FRENET_STAGE_K3)))
        return fail(FRENET_ERR_BAD_ARG, "one launch: needs K1 | K2 | K5 | GATHER (K3, K3 | K4 optional) and none of CURVATURE / NO_FUSE / OFFROAD");
    if ((rc = check_k2(L, P, B))) return rc;
    if (!B->tickets || !B->partials) return fail(FRENET_ERR_BAD_ARG, "one launch: buf->tickets / buf->partials is NULL");
    if (L->lat_f32) return fail(FRENET_ERR_BAD_ARG, "one launch: not available on the fp32 fast path");
    if (!B->result || !B->winner || !B->winner_steps) return fail(FRENET_ERR_BAD_ARG, "one launch: NULL result / winner buffer");
    if ((use_mask & 2) || (use_mask & 8)) return fail(FRENET_ERR_BAD_ARG, "one launch: curvature / road masks are not produced by this launch");
    if ((use_mask & 4) && !(stages & FRENET_STAGE_K4)) return fail(FRENET_ERR_BAD_ARG, "one launch: collision mask without K4");
    const int mono[4] = {use_mask, gather_sel, (stages & FRENET_STAGE_K3) ? 1 : 0, 0};
    if (!(stages & FRENET_STAGE_K3)) return launch_k2<false, 0>(L, P, nullptr, B, 0, (cudaStream_t)stream, mono, inl);
    if ((rc = check_path(path))) return rc;
    if (!(stages & FRENET_STAGE_K4)) return launch_k2<true, 0>(L, P, path, B, 0, (cudaStream_t)stream, mono, inl);
    if (!B->coll_free || !B->first_blocker) return fail(FRENET_ERR_BAD_ARG, "one launch: collision output is NULL");
    if (B->n_obs < 0 || (B->n_obs > 0 && !B->obs_xy)) return fail(FRENET_ERR_BAD_ARG, "one launch: obstacle table missing");
    if (B->n_obs_steps > 0 && B->n_obs > 0) return fail(FRENET_ERR_BAD_ARG, "one launch: predicted obstacles are not supported");
    return launch_k2<true, 1>(L, P, path, B, 0, (cudaStream_t)stream, mono, inl);
}

int frenet_plan(const FrenetLattice* L, const FrenetParams* P, const FrenetPath* path, const FrenetBuffers* B, int32_t stages,
                int32_t use_mask, int32_t gather_sel, frenet_stream_t stream) {
    int rc = FRENET_OK;
    if (stages & FRENET_STAGE_ONE_LAUNCH) return plan_one_launch(L, P, path, B, stages, use_mask, gather_sel, stream, nullptr);
    if ((stages & FRENET_STAGE_K1) && (rc = frenet_k1_coefficients(L, P, B, stream))) return rc;
    // K2 and K3 (and K4 when nothing sits between them) run as ONE kernel unless the caller asks for separate launches:
    // same arithmetic, same outputs, but every sample is written once instead of being re-read by K3 and K4.
    const bool fuse = (stages & FRENET_STAGE_K2) && (stages & FRENET_STAGE_K3) && !(stages & FRENET_STAGE_NO_FUSE);
    const bool fuse_k4 = fuse && (stages & FRENET_STAGE_K4);
    const bool fuse_curv = fuse && (stages & FRENET_STAGE_CURVATURE);
    if (fuse) {
        if ((rc = frenet_k2_fused(L, P, path, B, (fuse_k4 ? FRENET_FUSE_COLLISION : 0) | (fuse_curv ? FRENET_FUSE_CURVATURE : 0), stream))) return rc;
    } else {
        if ((stages & FRENET_STAGE_K2) && (rc = frenet_k2_evaluate(L, P, B, stream))) return rc;
        if ((stages & FRENET_STAGE_K3) && (rc = frenet_k3_cartesian(L, P, path, B, stream))) return rc;
    }
    if ((stages & FRENET_STAGE_CURVATURE) && !fuse_curv && (rc = frenet_k3_curvature(L, P, B, stream))) return rc;
    if ((stages & FRENET_STAGE_K4) && !fuse_k4 && (rc = frenet_k4_collision(L, P, B, stream))) return rc;
    if ((stages & FRENET_STAGE_OFFROAD) && (rc = frenet_k3_offroad(L, nullptr, B, stream))) return rc;
    // the winner block carries x, y when they exist: K3 / K4 ran in this call, or a mask computed from them is in use
    const int winner_xy = ((stages & (FRENET_STAGE_K3 | FRENET_STAGE_K4)) ||
                           (use_mask & (FRENET_MASK_CURVATURE | FRENET_MASK_COLLISION | FRENET_MASK_ROAD))) ? 1 : 0;
    if ((stages & FRENET_STAGE_K5) && (stages & FRENET_STAGE_GATHER) && !(stages & FRENET_STAGE_NO_FUSE))
        return launch_k5(L, P, B, use_mask, gather_sel, winner_xy, (cudaStream_t)stream);   // argmin and winner gather in one launch
    if ((stages & FRENET_STAGE_K5) && (rc = frenet_k5_argmin(L, B, use_mask, stream))) return rc;
    if ((stages & FRENET_STAGE_GATHER) && (rc = frenet_gather_winner(L, P, B, gather_sel, winner_xy, stream)))
        return rc;
    return rc;
}

int frenet_stream_synchronize(frenet_stream_t stream) {
    const cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e != cudaSuccess) return fail(FRENET_ERR_CUDA, "stream synchronize", e);
    return FRENET_OK;
}

int frenet_replan_inline(const FrenetLattice* L, const FrenetParams* P, const FrenetPath* path, const FrenetBuffers* B, int32_t stages,
                         int32_t use_mask, int32_t gather_sel, const double* h_state, const double* h_scal, const double* h_axis_v,
                         const int32_t* h_n_v, const double* h_target, int32_t wait, frenet_stream_t stream) {
    if (!L || !h_state || !h_scal || !h_axis_v || !h_n_v) return fail(FRENET_ERR_BAD_ARG, "replan inline: NULL host input");
    if (L->n_scen != 1) return fail(FRENET_ERR_BAD_ARG, "replan inline: one planner only (n_scen == 1)");
    if (L->n_v_cap > FRENET_INLINE_MAX_V || *h_n_v > FRENET_INLINE_MAX_V || *h_n_v < 0)
        return fail(FRENET_ERR_CAPACITY, "replan inline: longitudinal axis longer than FRENET_INLINE_MAX_V");
    if (!(stages & FRENET_STAGE_ONE_LAUNCH) || !(stages & FRENET_STAGE_K1))
        return fail(FRENET_ERR_BAD_ARG, "replan inline: needs FRENET_STAGE_ONE_LAUNCH with K1");
    FrenetInline in;
    in.present = 1;
    in.n_v = *h_n_v;
    for (int i = 0; i < 6; ++i) in.state[i] = h_state[i];
    for (int i = 0; i < 3; ++i) in.scal[i] = h_scal[i];
    for (int i = 0; i < FRENET_INLINE_MAX_V; ++i) in.axis_v[i] = i < L->n_v_cap ? h_axis_v[i] : 0.0;
    const bool follow = P && (P->follow || (B && B->follow));
    if (follow) {
        if (!h_target) return fail(FRENET_ERR_BAD_ARG, "replan inline: follow mode needs target triples");
        if (L->n_t * 3 > FRENET_INLINE_MAX_T * 3) return fail(FRENET_ERR_CAPACITY, "replan inline: more horizons than FRENET_INLINE_MAX_T");
        in.present = 2;
        for (int i = 0; i < L->n_t * 3; ++i) in.target[i] = h_target[i];
    }
    if (int rc = plan_one_launch(L, P, path, B, stages, use_mask, gather_sel, stream, &in)) return rc;
    if (wait) {
        const cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
        if (e != cudaSuccess) return fail(FRENET_ERR_CUDA, "replan inline: stream synchronize", e);
    }
    return FRENET_OK;
}

#if FRENET_PHASE_TIMING
// out[16]; reset != 0 re-arms the marks (0 -> max, the others -> 0)
int frenet_debug_phase_times(unsigned long long* out, int reset) {
    if (out) cudaMemcpyFromSymbol(out, g_phase_ts, sizeof(unsigned long long) * 16);
    if (reset) {
        unsigned long long z[16] = {};
        z[0] = ~0ull;
        cudaMemcpyToSymbol(g_phase_ts, z, sizeof(z));
    }
    return FRENET_OK;
}
#endif

int frenet_predict_obstacles(const FrenetPath* path, const double* obs_s, const double* obs_d, const double* obs_speed,
                             const double* t0, int64_t t0_stride, int32_t n_scen, int32_t n_obs, int32_t n_steps, double period,
                             double* obs_xy_t, frenet_stream_t stream) {
    if (int rc = check_path(path)) return rc;
    if (n_scen < 0 || n_obs < 0 || n_steps < 0) return fail(FRENET_ERR_BAD_ARG, "predict: negative size");
    const int64_t total = (int64_t)n_scen * n_obs * n_steps;
    if (total == 0) return FRENET_OK;
    if (!obs_s || !obs_d || !obs_speed || !t0 || !obs_xy_t) return fail(FRENET_ERR_BAD_ARG, "predict: NULL buffer");
    const size_t smem = (path->kind == FRENET_PATH_SPLINE ? (size_t)9 * path->n_knots : 0) * sizeof(double);
    if (int rc = set_smem(predict_obstacles, smem, "predict: spline table too large for shared memory")) return rc;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    predict_obstacles<<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(*path, obs_s, obs_d, obs_speed, t0, t0_stride, n_scen, n_obs,
                                                                             n_steps, period, obs_xy_t);
    FRENET_CHECK_LAUNCH("predict_obstacles");
    return FRENET_OK;
}

int frenet_advance_state(const FrenetLattice* L, const FrenetBuffers* B, const double* time, int64_t time_stride, double lookup_period,
                         double d_d_s, int32_t num_sample, int32_t follow, int32_t* status, frenet_stream_t stream) {
    if (int rc = check_lattice(L)) return rc;
    if (!B || !B->winner || !B->winner_steps || !B->state || !B->scal || !time) return fail(FRENET_ERR_BAD_ARG, "advance: NULL buffer");
    if (!(lookup_period > 0.0) || !(d_d_s > 0.0) || num_sample < 0) return fail(FRENET_ERR_BAD_ARG, "advance: bad interval parameters");
    advance_state<<<(L->n_scen + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*L, *B, time, time_stride, lookup_period, d_d_s, num_sample,
                                                                             follow, status);
    FRENET_CHECK_LAUNCH("advance_state");
    return FRENET_OK;
}

int frenet_follow_targets(const FrenetPath* path, const FrenetLattice* L, const FrenetBuffers* B, const int32_t* lead, const double* obs_s,
                          const double* obs_d, const double* obs_speed, int32_t n_obs, const double* estimate, double delta_t, double d_target,
                          const uint8_t* only, frenet_stream_t stream) {
    if (int rc = check_lattice(L)) return rc;
    if (int rc = check_path(path)) return rc;
    if (!B || !B->state || !B->scal || !B->target || !B->follow || !lead || !obs_s || !obs_d || !obs_speed || !estimate || n_obs <= 0)
        return fail(FRENET_ERR_BAD_ARG, "follow targets: NULL buffer");
    const size_t smem = (path->kind == FRENET_PATH_SPLINE ? (size_t)9 * path->n_knots : 0) * sizeof(double);
    if (int rc = set_smem(follow_targets, smem, "follow targets: spline table too large for shared memory")) return rc;
    const int64_t n = (int64_t)L->n_scen * L->n_t;
    follow_targets<<<(unsigned)((n + 127) / 128), 128, smem, (cudaStream_t)stream>>>(*path, *L, *B, lead, obs_s, obs_d, obs_speed, n_obs, estimate,
                                                                                    delta_t, d_target, only, const_cast<double*>(B->target));
    FRENET_CHECK_LAUNCH("follow_targets");
    return FRENET_OK;
}

int frenet_follow_select(const FrenetLattice* L, const FrenetBuffers* B, double d_d_s, int32_t num_sample, int32_t* lead, uint8_t* follow,
                         uint8_t* redo, frenet_stream_t stream) {
    if (int rc = check_lattice(L)) return rc;
    if (!B || !B->result || !B->cand_flags || !B->first_blocker || !lead || !follow || !redo)
        return fail(FRENET_ERR_BAD_ARG, "follow select: NULL buffer");
    if (!(d_d_s > 0.0) || num_sample < 0) return fail(FRENET_ERR_BAD_ARG, "follow select: bad interval parameters");
    follow_select<<<(L->n_scen + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*L, *B, d_d_s, num_sample, lead, follow, redo);
    FRENET_CHECK_LAUNCH("follow_select");
    return FRENET_OK;
}

int frenet_optimal_targets(const FrenetLattice* L, const FrenetBuffers* B, const double* spec, double delta_t, int32_t* status,
                           frenet_stream_t stream) {
    if (int rc = check_lattice(L)) return rc;
    if (!B || !B->scal || !B->target || !spec) return fail(FRENET_ERR_BAD_ARG, "optimal targets: NULL buffer");
    if (!(delta_t > 0.0)) return fail(FRENET_ERR_BAD_ARG, "optimal targets: delta_t must be positive");
    const int64_t n = (int64_t)L->n_scen * L->n_t;
    optimal_targets<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*L, *B, spec, delta_t, const_cast<double*>(B->target), status);
    FRENET_CHECK_LAUNCH("optimal_targets");
    return FRENET_OK;
}

int frenet_sim_project(const FrenetPath* path, const double* pose, double* estimate, int32_t max_iter, double damping, double* fpose,
                       double* curvature, int32_t n, frenet_stream_t stream) {
    if (int rc = check_path(path)) return rc;
    if (n < 0 || max_iter < 0) return fail(FRENET_ERR_BAD_ARG, "sim project: negative size");
    if (n == 0) return FRENET_OK;
    if (!pose || !estimate || !fpose || !curvature) return fail(FRENET_ERR_BAD_ARG, "sim project: NULL buffer");
    const size_t smem = (path->kind == FRENET_PATH_SPLINE ? (size_t)9 * path->n_knots : 0) * sizeof(double);
    if (int rc = set_smem(sim_project, smem, "sim project: spline table too large for shared memory")) return rc;
    sim_project<<<(n + 127) / 128, 128, smem, (cudaStream_t)stream>>>(*path, pose, estimate, max_iter, damping, fpose, curvature, n);
    FRENET_CHECK_LAUNCH("sim_project");
    return FRENET_OK;
}

int frenet_sim_control_step(const double* fpose, const double* state, const double* curvature, double b, double kp_s, double kp_d,
                            double dt, double* pose, double* control, int32_t n, frenet_stream_t stream) {
    if (n < 0) return fail(FRENET_ERR_BAD_ARG, "sim control: negative size");
    if (n == 0) return FRENET_OK;
    if (!fpose || !state || !curvature || !pose || !control) return fail(FRENET_ERR_BAD_ARG, "sim control: NULL buffer");
    sim_control_step<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(fpose, state, curvature, b, kp_s, kp_d, dt, pose, control, n);
    FRENET_CHECK_LAUNCH("sim_control_step");
    return FRENET_OK;
}

int frenet_sensor_gate(const double* robot_pose, const double* obs_xy, int32_t n_scen, int32_t n_obs, double range, double aperture,
                       uint8_t* active, frenet_stream_t stream) {
    if (n_scen < 0 || n_obs < 0) return fail(FRENET_ERR_BAD_ARG, "sensor gate: negative size");
    const int64_t total = (int64_t)n_scen * n_obs;
    if (total == 0) return FRENET_OK;
    if (!robot_pose || !obs_xy || !active) return fail(FRENET_ERR_BAD_ARG, "sensor gate: NULL buffer");
    sensor_gate<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(robot_pose, obs_xy, n_scen, n_obs, range, aperture, active);
    FRENET_CHECK_LAUNCH("sensor_gate");
    return FRENET_OK;
}

int frenet_polyline_collision(const double* x, const double* y, const int64_t* offsets, int32_t n_paths, const double* obs_xy,
                              int32_t n_obs, double radius_sq, uint8_t* coll_free, int32_t* first_blocker, frenet_stream_t stream) {
    if (n_paths < 0 || n_obs < 0) return fail(FRENET_ERR_BAD_ARG, "polyline: negative size");
    if (n_paths == 0) return FRENET_OK;
    if (!x || !y || !offsets || !coll_free || !first_blocker || (n_obs > 0 && !obs_xy))
        return fail(FRENET_ERR_BAD_ARG, "polyline: NULL buffer");
    const size_t smem = (size_t)3 * n_obs * sizeof(double);
    if (int rc = set_smem(polyline_collision, smem, "polyline: too many obstacles for shared memory")) return rc;
    polyline_collision<<<(n_paths + 7) / 8, 256, smem, (cudaStream_t)stream>>>(x, y, offsets, n_paths, obs_xy, n_obs, radius_sq,
                                                                              coll_free, first_blocker);
    FRENET_CHECK_LAUNCH("polyline_collision");
    return FRENET_OK;
}

int frenet_points_to_cartesian(const FrenetPath* path, const double* s, const double* d, int64_t n, double* x, double* y,
                               int64_t out_stride, frenet_stream_t stream) {
    if (int rc = check_path(path)) return rc;
    if (n < 0 || (n > 0 && (!s || !d || !x || !y))) return fail(FRENET_ERR_BAD_ARG, "points: NULL buffer");
    if (out_stride < 1) return fail(FRENET_ERR_BAD_ARG, "points: out_stride must be >= 1");
    if (n == 0) return FRENET_OK;
    const size_t smem = (path->kind == FRENET_PATH_SPLINE ? (size_t)9 * path->n_knots : 0) * sizeof(double);
    if (int rc = set_smem(points_to_cartesian, smem, "points: spline table too large for shared memory")) return rc;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;   // grid-stride beyond one full wave of 8 CTAs per SM
    points_to_cartesian<<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(*path, s, d, n, x, y, out_stride);
    FRENET_CHECK_LAUNCH("points_to_cartesian");
    return FRENET_OK;
}

int frenet_workspace_bytes(const FrenetLattice* L, int32_t n_obs, int32_t n_obs_steps, FrenetSizes* out) {
    if (!out) return fail(FRENET_ERR_BAD_ARG, "workspace: output is NULL");
    if (!L || L->n_scen <= 0 || L->n_v_cap <= 0 || L->n_t <= 0 || L->n_d <= 0 || L->n_pad_max <= 0 || L->row_elems <= 0 ||
        L->cand_elems != L->row_elems * L->n_d || (L->row_elems & (L->lat_f32 ? 7 : 3)) || n_obs < 0 || n_obs_steps < 0)
        return fail(FRENET_ERR_BAD_ARG, "workspace: inconsistent lattice sizes");
    memset(out, 0, sizeof(*out));
    const int64_t Bn = L->n_scen, R = (int64_t)L->n_v_cap * L->n_t, C = R * L->n_d;
    out->rows = Bn * R;
    out->candidates = Bn * C;
    out->state = Bn * 6 * 8;
    out->scal = Bn * 3 * 8;
    out->target = Bn * L->n_t * 3 * 8;
    out->axis_v = Bn * L->n_v_cap * 8;
    out->n_v = Bn * 4;
    out->axis_d = (int64_t)L->n_d * 8;
    out->axis_t = (int64_t)L->n_t * 8;
    out->n_steps = (int64_t)L->n_t * 4;
    out->t_offset = (int64_t)(L->n_t + 1) * 4;
    out->t_pow = (int64_t)L->n_pad_max * 4 * 8;
    out->obs_xy = Bn * n_obs * 2 * 8;
    out->obs_active = Bn * n_obs;
    out->obs_xy_t = Bn * n_obs * (int64_t)n_obs_steps * 2 * 8;
    out->obs_bounds = n_obs_steps > 0 ? Bn * ((L->n_pad_max + kWarp - 1) / kWarp) * n_obs * 4 * 4 : 0;
    out->coef_lon = Bn * R * 6 * 8;
    out->coef_lat = Bn * C * 6 * 8;
    out->row_S = Bn * R * 8;
    out->row_flags = Bn * R * 4;
    out->lon = Bn * L->row_elems * kLonFields * 8;
    out->lat = Bn * L->cand_elems * kCandFields * (L->lat_f32 ? 4 : 8);
    out->cost = 4 * Bn * C * 8;
    out->cand_flags = Bn * C;
    out->curv_ok = Bn * C;
    out->coll_free = Bn * C;
    out->first_blocker = Bn * C * 4;
    out->road_ok = Bn * C;
    out->result = Bn * (int64_t)sizeof(FrenetResult);
    out->winner = Bn * FRENET_WINNER_ROWS * L->n_pad_max * 8;
    out->winner_steps = Bn * 4;
    out->path_origin = Bn * 2 * 8;
    out->lane_params = Bn * 4 * 8;
    out->road_params = Bn * 8 * 8;
    out->tickets = Bn * 4;
    out->partials = Bn * R * FRENET_MAX_ROW_PARTS * FRENET_PARTIAL_BYTES;
    const int64_t* f = &out->state;
    int64_t tot = 0;
    for (; f <= &out->partials; ++f) tot += (*f + 255) / 256 * 256;   // every buffer on its own 256-byte boundary
    out->total = tot;
    return FRENET_OK;
}

int frenet_k2_launch_shape(const FrenetLattice* L, const FrenetPath* path, int32_t n_obs, int32_t n_obs_steps, int32_t with_collision,
                           int32_t* rows_per_cta, int32_t* ctas_per_row, int64_t* grid, int64_t* shared_bytes) {
    if (!L || L->n_scen <= 0 || L->n_v_cap <= 0 || L->n_t <= 0 || L->n_d <= 0 || L->n_pad_max <= 0 || n_obs < 0)
        return fail(FRENET_ERR_BAD_ARG, "launch shape: inconsistent lattice sizes");
    const bool xy = path != nullptr;
    const int coll = (xy && with_collision) ? (n_obs_steps > 0 && n_obs > 0 ? 2 : 1) : 0;
    const K2Shape S = k2_shape(L, xy, coll, (xy && path->kind == FRENET_PATH_SPLINE) ? path->n_knots : 0, n_obs);
    if (rows_per_cta) *rows_per_cta = S.rows_per_cta;
    if (ctas_per_row) *ctas_per_row = S.parts;
    if (grid) *grid = (int64_t)S.grid * S.parts;
    if (shared_bytes) *shared_bytes = (int64_t)S.smem;
    return S.smem > (size_t)kMaxSmem ? fail(FRENET_ERR_CAPACITY, "launch shape: the row does not fit shared memory") : FRENET_OK;
}

}  // extern "C"
