/*
 * frenet_b200.h - C ABI of the B200 (sm_100a) Frenet lattice planner hot path.
 *
 * The reference (pepes97/Optimal-trajectory-generation-in-the-Duckietown-environment) is pure
 * Python; it has no FFI layer.  Its boundary for this path is the Python planner API, so every
 * entry point below names the reference function it replaces (paths relative to the reference
 * root).  The Python host side (`otg_b200.planner`, `otg_b200.drivers`, `otg_b200.batch`) binds
 * these symbols with ctypes; INTEGRATION.md shows the stub a maintainer of the reference would
 * add to call them from `lib/planner/trajectory_planner.py`.
 *
 * Conventions
 *   - plain C types only; every pointer inside FrenetBuffers / FrenetLattice / FrenetPath is a
 *     DEVICE pointer owned by the caller (allocate with any CUDA allocator, e.g. torch);
 *     the structs themselves are HOST PODs read during the call.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*); no global state
 *     except the per-thread last-error string; kernels are re-entrant across streams.
 *   - return value: FRENET_OK or a negative FRENET_ERR_* code; nothing throws across the ABI.
 *     "No feasible candidate" is not an error of the call: it is reported per scenario in
 *     FrenetResult.status (the Python wrapper raises ValueError like the reference's `min([])`).
 *   - there is no CPU fallback: without a CUDA device every launch returns FRENET_ERR_CUDA.
 *
 * Indexing: a batch holds n_scen independent planner instances
 * ("scenarios").  Scenario b owns n_v[b] <= n_v_cap longitudinal targets; row r = i_v * n_t + i_T
 * owns one longitudinal polynomial, candidate c = r * n_d + i_d one lateral polynomial - the
 * order in which `generate_range_polynomials` appends (lib/planner/trajectory_planner.py:128-180).
 * Horizon i_T has n_steps[i_T] samples t_k = k * sample_period (np.arange(0, T + dt, dt)), stored
 * padded to npad(i_T) = n_steps rounded up to a multiple of 4 (8 with lat_f32) elements; t_offset[i_T] is the prefix sum
 * of the padded counts.  With  within(r) = i_v * t_offset[n_t] + t_offset[i_T] :
 *
 *   longitudinal block of row r   (4 fields s, ds, dds, ddds, each npad long, back to back):
 *        lon[ (b * row_elems  + within(r)) * 4              + f * npad + k ]
 *   candidate block of c = (r, i_d)  (6 fields d, dd, ddd, dddd, x, y, each npad long, back to back):
 *        lat[ (b * cand_elems + within(r) * n_d) * 6 + i_d * 6 * npad + f * npad + k ]
 *   per-candidate arrays:  element (b, c) at  b * n_v_cap * n_t * n_d + c
 *   per-row arrays:        element (b, r) at  b * n_v_cap * n_t + r
 *
 * Every field of a candidate is a contiguous, 32-byte aligned run of doubles (k fastest), and the six
 * fields of one candidate sit next to each other, so a warp that owns a candidate writes one compact
 * region.  The longitudinal samples are stored once per row: the n_d candidates of a row share them (the
 * reference deep-copies them into every candidate, lib/planner/trajectory_planner.py:157).
 */
#ifndef FRENET_B200_H
#define FRENET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FRENET_B200_ABI_VERSION 3
#define FRENET_PARTIAL_BYTES 72   /* one partial argmin record of the one-launch replan (buf->partials) */
#define FRENET_MAX_ROW_PARTS 8    /* a row is split over at most this many CTAs */

typedef void* frenet_stream_t; /* cudaStream_t */

enum {
    FRENET_OK = 0,
    FRENET_ERR_BAD_ARG = -1, /* null pointer, non-positive size, inconsistent layout */
    FRENET_ERR_CUDA = -2,    /* launch or runtime error, text in frenet_last_error() */
    FRENET_ERR_CAPACITY = -3 /* shared-memory demand of the requested shape exceeds 227 KB */
};

/* FrenetResult.status */
enum { FRENET_STATUS_OK = 0, FRENET_STATUS_NO_FEASIBLE = 1, FRENET_STATUS_EMPTY = 2 };

/* which lateral parameterisation rule a planner variant uses */
enum {
    FRENET_LOWSPEED_OFF = 0,    /* V1DT, V1DTObstacles: `... and False` (trajectory_planner_dt.py:156) */
    FRENET_LOWSPEED_V1 = 1,     /* ds0 < threshold and |S| > 0 (trajectory_planner.py:159) */
    FRENET_LOWSPEED_OPTIMAL = 2 /* ds0 < threshold: keep only if signed S > s_thresh (trajectory_planner_optimal.py:155-168) */
};

/* row_flags bits (K1) */
enum { FRENET_ROW_EXISTS = 1, FRENET_ROW_LOWSPEED = 2, FRENET_ROW_DROPPED = 4 };
/* cand_flags bits (K2) */
enum { FRENET_CAND_VALID = 1, FRENET_CAND_KINEMATIC_OK = 2 };

enum { FRENET_PATH_CIRCLE = 0, FRENET_PATH_SPLINE = 1, FRENET_PATH_PARABOLA = 2 };

/* bits of `use_mask` (frenet_k5_argmin, frenet_plan): which filters the masked pick honours */
enum { FRENET_MASK_KINEMATIC = 1, FRENET_MASK_CURVATURE = 2, FRENET_MASK_COLLISION = 4, FRENET_MASK_ROAD = 8 };

/* selector for frenet_gather_winner */
#define FRENET_WINNER_ROWS 12
enum { FRENET_SEL_CTOT = 0, FRENET_SEL_CD = 1, FRENET_SEL_CV = 2, FRENET_SEL_CT = 3, FRENET_SEL_MASKED = 4 };

/* Weights and thresholds: the attributes of TrajectoryPlannerParams (lib/planner/trajectory_planner.py:38-67)
 * that `generate_range_polynomials` reads, plus the collision radius of the drivers and the
 * (extension, default +inf) kinematic limits. */
typedef struct FrenetParams {
    double kj, kt, kd, ks, kdots, klat, klong;
    double sample_period;       /* delta_t; sampling_t for V1DTObstacles (trajectory_planner_obstacles.py:136) */
    double low_speed_threshold;
    double s_thresh;            /* Optimal only */
    double v_max, a_max, kappa_max; /* EXTENSION: +inf reproduces the reference (no limits) */
    double radius_sq;           /* robot_radius ** 2, lib/tests/test_obstacle_planner.py:16,48 */
    int32_t lowspeed_rule;      /* FRENET_LOWSPEED_* */
    int32_t follow;             /* 1: `s_target` given - quintic longitudinal to the target triples, cost ct */
} FrenetParams;

/* Lattice geometry shared by every scenario of a batch. */
typedef struct FrenetLattice {
    int32_t n_scen;
    int32_t n_v_cap;
    int32_t n_t;
    int32_t n_d;
    int32_t n_pad_max;       /* max_i (t_offset[i+1] - t_offset[i]) */
    int32_t lat_f32;         /* fp32 FAST PATH (0 = off): 1 = `lat` holds FLOATS - the candidate blocks d, dd, ddd, dddd, x, y are stored in
                                fp32 (24 instead of 48 bytes per candidate-timestep), same element offsets, and the padded step counts
                                are multiples of 8 (whole 32-byte sectors).  Only the fused pipeline kernel of long horizons supports it
                                (frenet_k2_fused / frenet_plan with K2 + K3 fused); all arithmetic, the longitudinal samples `lon`, the
                                costs, the collision test and the gathered winner stay fp64, so masks and argmin are those of the fp64
                                path and only the stored lateral / Cartesian samples carry fp32 rounding (north star: 1e-3 relative) */
    int64_t row_elems;       /* n_v_cap * t_offset[n_t] */
    int64_t cand_elems;      /* row_elems * n_d */
    const double* axis_d;    /* [n_d]   np.arange(*di_interval) */
    const double* axis_t;    /* [n_t]   np.arange(*t_interval) */
    const int32_t* n_steps;  /* [n_t]   len(np.arange(0, T + dt, dt)) */
    const int32_t* t_offset; /* [n_t+1] prefix sum of the step counts rounded up to 4 */
    const double* t_pow;     /* [n_pad_max][4] t_k^2 .. t_k^5 for t_k = k * sample_period, evaluated on the host with the
                                reference's own `**` (C pow): the longitudinal polynomial is sampled in the reference's
                                power form with these, so that its values - which become the next replan's initial state
                                and sit on the low-speed threshold in converged runs - round the way the reference's do */
    const double* axis_v;    /* [n_scen][n_v_cap] target speeds (velocity keeping) or target offsets (follow) */
    const int32_t* n_v;      /* [n_scen] */
} FrenetLattice;

/* Reference path descriptor (lib/trajectory/circle_trajectory.py:11-72, spline_trajectory.py:9-158).
 * FRENET_PATH_PARABOLA is the lane-centre fit of the Duckietown runs (lib/mapper/mapper_obstacles.py:169-186):
 * pt(x) = (x, a x^2 + b x + c) parameterised by x, normal from the chord over ds (compute_ortogonal_vect,
 * lib/tests/test_mapper_planner_obstacles.py:137-142); circle[0..3] = a, b, c, ds. */
typedef struct FrenetPath {
    int32_t kind;        /* FRENET_PATH_* */
    int32_t n_knots;     /* spline only */
    double circle[16];   /* circle: l, h, r, dt, period, 7 branch boundaries (host-evaluated in the reference's order);
                            parabola: a, b, c, ds */
    const double* spline_table; /* device [n_knots][9]: knot, ax, bx, cx, dx, ay, by, cy, dy */
} FrenetPath;

/* Per-scenario result record written by K5 (64 bytes). */
typedef struct FrenetResult {
    int32_t argmin_ctot;   /* min(paths, key=ctot), first minimal wins (trajectory_planner.py:101,212); -1 if none */
    int32_t argmin_cd;     /* key cd  (:99,105) */
    int32_t argmin_cv;     /* key cv  (:100,106) */
    int32_t argmin_ct;     /* key ct  (:103) */
    int32_t argmin_masked; /* key ctot over candidates that pass every enabled filter; -1 if none */
    int32_t n_valid;       /* candidates the reference would have appended */
    int32_t n_survivors;   /* of those, how many pass every enabled filter */
    int32_t status;        /* FRENET_STATUS_* */
    double min_ctot;
    double min_masked_ctot;
    double reserved[2];
} FrenetResult;

/* Device buffers.  Inputs are read-only for the kernels; outputs may be NULL when the kernel that
 * writes them is not launched. */
typedef struct FrenetBuffers {
    /* inputs */
    const double* state;       /* [n_scen][6]  p0, dp0, ddp0, s0, ds0, dds0 */
    const double* scal;        /* [n_scen][3]  dd, desired_speed, t_initial */
    const double* target;      /* [n_scen][n_t][3] (st, dst, ddst) per horizon, or NULL (params.follow == 0) */
    const double* obs_xy;      /* [n_scen][n_obs][2] obstacle snapshot, or NULL */
    const uint8_t* obs_active; /* [n_scen][n_obs] 1 = sensed (tested), or NULL = all */
    int32_t n_obs;
    int32_t n_obs_steps;       /* EXTENSION: 0 = snapshot table `obs_xy` (the reference: every step is tested against the obstacles'
                                  CURRENT positions, lib/tests/test_moving_obstacle_planner.py:89-91); K > 0 = predicted table below */
    const double* obs_xy_t;    /* [n_scen][n_obs][n_obs_steps][2] predicted positions: step k is tested against entry min(k, K-1) */
    float* obs_bounds;         /* scratch of n_scen * ceil(n_pad_max/32) * n_obs * 4 floats: per chunk of 32 (or 64) steps and obstacle the
                                  bounding circle (cx, cy, r, padded r) of its predicted positions; the collision stage fills it */
    /* K1 outputs */
    double* coef_lon;          /* [n_scen][R][6] a0..a5 (a5 = 0 for the velocity-keeping quartic) */
    double* coef_lat;          /* [n_scen][C][6] */
    double* row_S;             /* [n_scen][R] arc length travelled (low-speed horizon) */
    int32_t* row_flags;        /* [n_scen][R] FRENET_ROW_* */
    /* K2 outputs */
    double* lon;               /* [n_scen][row_elems * 4]:  per row  s, ds, dds, ddds (layout above) */
    double* lat;               /* [n_scen][cand_elems * 6]: per candidate d, dd, ddd, dddd (K2) and x, y (K3) */
    double* cost;              /* 4 arrays [4][n_scen][C]: cd, cv, ct, ctot (unset keys are 0.0 like Frenet()) */
    uint8_t* cand_flags;       /* [n_scen][C] FRENET_CAND_* */
    /* K3 outputs */
    uint8_t* curv_ok;          /* [n_scen][C] EXTENSION curvature limit, or NULL (kappa_max = +inf) */
    /* K4 outputs */
    uint8_t* coll_free;        /* [n_scen][C] 1 = no collision (check_collisions -> True) */
    int32_t* first_blocker;    /* [n_scen][C] lowest blocking obstacle index, -1 if none */
    /* K5 outputs */
    FrenetResult* result;      /* [n_scen] */
    /* winner gather */
    double* winner;            /* [n_scen][12][n_pad_max]: t, d, dd, ddd, dddd, s, ds, dds, ddds, x, y, then a row holding
                                  the winner's (cd, cv, ct, ctot) in its first four entries */
    int32_t* winner_steps;     /* [n_scen] number of valid samples, 0 if no winner */
    /* Duckietown lane runs; both optional (NULL) */
    const double* path_origin; /* [n_scen][2] (projection, planner.s0[0]): the lattice transform evaluates the path at
                                  projection + (s - s0[0]) (frenet_to_glob_planner,
                                  lib/tests/test_mapper_planner_obstacles.py:25-37); NULL: at s */
    uint8_t* road_ok;          /* [n_scen][C] 1 = stays between the lane boundaries (frenet_k3_offroad) */
    /* batched lane runs: every scenario has its own lane fit (both optional, NULL = the shared FrenetPath / FrenetRoad) */
    const double* lane_params; /* [n_scen][4] a, b, c, ds of the scenario's lane-centre parabola (path kind FRENET_PATH_PARABOLA) */
    const double* road_params; /* [n_scen][8] right boundary (h, k, a, given != 0), left boundary (h, k, a, given != 0) */
    /* per-scenario switches of a batch (ABI 2; both optional, NULL = off) */
    const uint8_t* follow;     /* [n_scen] 1 = this scenario plans in follow mode (overrides FrenetParams.follow scenario by scenario):
                                  agents that fell back to following their leading obstacle next to agents that keep their speed */
    const uint8_t* scen_mask;  /* [n_scen] 1 = process this scenario, 0 = leave its outputs untouched (K1, K2 / fused, K5, gather,
                                  frenet_advance_state): the follow re-plan of the blocked scenarios only */
    /* one-launch replan (ABI 3; optional, NULL = FRENET_STAGE_ONE_LAUNCH unavailable) */
    int32_t* tickets;          /* [n_scen] scratch, zero before the first launch (the kernel leaves it zero): completion tickets of a
                                  scenario's CTAs - the CTA that takes the last one finishes the argmin and gathers the winner */
    double* winner_masked;     /* optional [n_scen][12][n_pad_max]: FRENET_STAGE_ONE_LAUNCH also gathers the MASKED pick here (next to
                                  the selection `gather_sel` in `winner`), so one launch answers the planner and check_paths */
    int32_t* winner_masked_steps; /* [n_scen], with winner_masked */
    void* partials;            /* scratch of n_scen * R * FRENET_MAX_ROW_PARTS * FRENET_PARTIAL_BYTES bytes: every CTA's own partial
                                  argmin record (frenet_workspace_bytes reports the size as `partials`) */
} FrenetBuffers;

/* Inline inputs of frenet_replan_inline: ONE planner's per-replan inputs, passed to the kernel as an argument (no copy node). */
#define FRENET_INLINE_MAX_V 16   /* longitudinal targets */
#define FRENET_INLINE_MAX_T 24   /* horizons (follow-mode target triples) */
typedef struct FrenetInline {
    int32_t present;           /* 0: not used; 1: state, scal, axis_v, n_v; 2: + target */
    int32_t n_v;
    double state[6];
    double scal[3];
    double axis_v[FRENET_INLINE_MAX_V];
    double target[FRENET_INLINE_MAX_T * 3];
} FrenetInline;

/* Lane boundaries for frenet_k3_offroad: each a parabola in vertex form (h, k, a) =
 * get_vertex_and_focus_distance(fit) (lib/tests/test_mapper_planner_obstacles.py:105-109). */
typedef struct FrenetRoad {
    double right[3];
    double left[3];
    int32_t has_right, has_left;   /* 0: that boundary was not detected (check_paths skips the test, :117-124) */
} FrenetRoad;

/* pipeline stage bits for frenet_plan */
enum {
    FRENET_STAGE_K1 = 1, FRENET_STAGE_K2 = 2, FRENET_STAGE_K3 = 4, FRENET_STAGE_K4 = 8,
    FRENET_STAGE_K5 = 16, FRENET_STAGE_GATHER = 32, FRENET_STAGE_CURVATURE = 64,
    FRENET_STAGE_NO_FUSE = 128, /* launch K2, K3, K4 separately even when frenet_plan could fuse them */
    FRENET_STAGE_OFFROAD = 256, /* lane-boundary filter from buf->road_params (per scenario), between K4 and K5 */
    FRENET_STAGE_ONE_LAUNCH = 512 /* K1 | K2 (| K3 (| K4)) | K5 | GATHER as ONE kernel launch: every CTA of the evaluate kernel solves
                                     its own rows' boundary-value systems first, and the scenario's last CTA to finish runs the argmin
                                     and the winner gather (needs buf->tickets).  Same results as the separate launches; meant for a
                                     single planner, whose replan is a chain of launch latencies.  Not combined with the curvature /
                                     off-road stages, predicted obstacles, FRENET_STAGE_NO_FUSE or the fp32 fast path.
                                     Second form: K4 | K5 | GATHER | ONE_LAUNCH - the collision stage alone on a lattice that is already
                                     sampled (x, y present), plus the argmin / gather tail: check_paths on obstacles that arrive after
                                     the replan.  Long horizons (path required): from the row-level table, no stored sample is read
                                     back; lattices of at most 32 padded samples: the sweep over the stored x, y (path may be NULL). */
};

int frenet_abi_version(void);
const char* frenet_last_error(void);

/* Number of CUDA devices visible to the library (0 when there is no driver). */
int frenet_device_count(void);

/* K1 - candidate generation: closed-form boundary-value solutions, one thread per lattice cell.
 * Replaces QuarticPolynomial.__init__ / QuinticPolynomial.__init__ (lib/trajectory/trajectory_1d.py:15-40,
 * 76-97) as called from generate_range_polynomials (lib/planner/trajectory_planner.py:138,147,160,169). */
int frenet_k1_coefficients(const FrenetLattice* lat, const FrenetParams* prm, const FrenetBuffers* buf, frenet_stream_t stream);

/* K2 - evaluate / cost / feasibility: samples every candidate, accumulates jerk sums and cd/cv/ct/ctot.
 * Replaces the sampling and cost lines of generate_range_polynomials (trajectory_planner.py:139-176). */
int frenet_k2_evaluate(const FrenetLattice* lat, const FrenetParams* prm, const FrenetBuffers* buf, frenet_stream_t stream);

/* K2 + K3 (+ K4, + the curvature limit) as one kernel: every sample is evaluated, transformed (and collision- / curvature-tested)
 * while it is in registers and written to HBM once.  Outputs are identical to the separate launches.
 * with_collision: bit set of FRENET_FUSE_* (1 = the historical "with collision"). */
enum { FRENET_FUSE_COLLISION = 1, FRENET_FUSE_CURVATURE = 2 };
int frenet_k2_fused(const FrenetLattice* lat, const FrenetParams* prm, const FrenetPath* path, const FrenetBuffers* buf,
                    int32_t with_collision, frenet_stream_t stream);

/* K3 - Frenet -> Cartesian of every candidate against the circle / spline reference path.
 * Replaces frenet_to_glob + compute_ortogonal_vect (lib/tests/test_obstacle_planner.py:18-25, 53-56). */
int frenet_k3_cartesian(const FrenetLattice* lat, const FrenetParams* prm, const FrenetPath* path, const FrenetBuffers* buf,
                        frenet_stream_t stream);

/* EXTENSION - discrete curvature limit on the Cartesian polyline (no reference code). */
int frenet_k3_curvature(const FrenetLattice* lat, const FrenetParams* prm, const FrenetBuffers* buf, frenet_stream_t stream);

/* K4 - candidate x time-step x obstacle collision test.
 * Replaces check_paths / check_collisions (lib/tests/test_moving_obstacle_planner.py:26-52). */
int frenet_k4_collision(const FrenetLattice* lat, const FrenetParams* prm, const FrenetBuffers* buf, frenet_stream_t stream);

/* K5 - first-wins argmin per cost key, unmasked and masked.
 * Replaces min(paths, key=attrgetter(...)) (trajectory_planner.py:99-106, 212; test_obstacle_planner.py:90).
 * use_mask bits: 1 kinematic flags, 2 curv_ok, 4 coll_free. */
int frenet_k5_argmin(const FrenetLattice* lat, const FrenetBuffers* buf, int32_t use_mask, frenet_stream_t stream);

/* Copies the selected winner's sampled state into buf->winner (what optimal_at_time reads,
 * trajectory_planner.py:183-195).  sel: FRENET_SEL_*; has_xy: 0 leaves x, y rows untouched. */
int frenet_gather_winner(const FrenetLattice* lat, const FrenetParams* prm, const FrenetBuffers* buf, int32_t sel, int32_t has_xy,
                         frenet_stream_t stream);

/* Off-road filter of the Duckietown runs: check_offroad (lib/tests/test_mapper_planner_obstacles.py:62-70) for every candidate and
 * both boundaries: road_ok = no sample with (y - k) - a (x - h)^2 < 0 on the right boundary nor > 0 on the left one.
 * Needs K3 (x, y).  Honoured by K5 through FRENET_MASK_ROAD. */
int frenet_k3_offroad(const FrenetLattice* lat, const FrenetRoad* road, const FrenetBuffers* buf, frenet_stream_t stream);
/* (road may be NULL when buf->road_params holds per-scenario boundaries, which take precedence) */

/* The pipeline: launches the stages selected in `stages` in order K1, K2, K3, (curvature), K4, K5, gather;
 * K2+K3(+K4, +curvature) are fused into one launch when selected together (FRENET_STAGE_NO_FUSE keeps them apart). */
int frenet_plan(const FrenetLattice* lat, const FrenetParams* prm, const FrenetPath* path, const FrenetBuffers* buf,
                int32_t stages, int32_t use_mask, int32_t gather_sel, frenet_stream_t stream);

/* Point-wise Frenet -> Cartesian (obstacle placement, target point): (x, y) = pt(s) + n(s) * d.
 * Replaces `trajectory.compute_pt(s) + compute_ortogonal_vect(trajectory, s) * d`
 * (lib/tests/test_obstacle_planner.py:21; lib/sensor/dynamic_obstacle.py:46-56).
 * Point i is written to x[i * out_stride], y[i * out_stride] (out_stride 2 with y = x + 1 fills an [n][2] table). */
int frenet_points_to_cartesian(const FrenetPath* path, const double* s, const double* d, int64_t n, double* x, double* y,
                               int64_t out_stride, frenet_stream_t stream);

/* cudaStreamSynchronize behind the same error convention (a host that holds only the raw stream handle - e.g. a Python binding, where
 * fetching a stream OBJECT to wait on costs more than the wait - calls this after the asynchronous entry points). */
int frenet_stream_synchronize(frenet_stream_t stream);

/* frenet_plan with FRENET_STAGE_ONE_LAUNCH (first form: K1 | K2 ... ) for ONE planner (n_scen == 1) whose per-replan inputs are read
 * from HOST memory at call time and handed to the kernel as an argument: h_state[6], h_scal[3], h_axis_v[n_v_cap], *h_n_v and - in
 * follow mode - h_target[n_t][3] (NULL otherwise).  The kernel spills them into buf->state / scal / target and lat->axis_v / n_v
 * (which must be device-writable), so later launches on the same buffers (check_paths' collision stage, a gather) see them.  No
 * host-to-device copy, no graph: a replan is this one launch.  n_v_cap <= FRENET_INLINE_MAX_V, n_t <= FRENET_INLINE_MAX_T in
 * follow mode (FRENET_ERR_CAPACITY otherwise - use frenet_plan).  wait != 0: the call returns after the stream has drained (the
 * results are then in buf->result / winner).  Replaces the same reference code as frenet_plan
 * (lib/planner/trajectory_planner.py:108-181, 206-212). */
int frenet_replan_inline(const FrenetLattice* lat, const FrenetParams* prm, const FrenetPath* path, const FrenetBuffers* buf, int32_t stages,
                         int32_t use_mask, int32_t gather_sel, const double* h_state, const double* h_scal, const double* h_axis_v,
                         const int32_t* h_n_v, const double* h_target, int32_t wait, frenet_stream_t stream);

/* EXTENSION - device-side obstacle prediction: fills a predicted table [n_scen][n_obs][n_steps][2] from the moving-obstacle
 * time law of the reference (lib/sensor/dynamic_obstacle.py:43-56): at step k obstacle o of scenario b sits at arc length
 * obs_s + ((t0[b] + k * period) * obs_speed) % 50 with lateral offset obs_d.  obs_s / obs_d / obs_speed: [n_scen][n_obs];
 * t0: [n_scen] with stride t0_stride doubles (pass buf->scal + 2 and 3 to use the planners' t_initial). */
int frenet_predict_obstacles(const FrenetPath* path, const double* obs_s, const double* obs_d, const double* obs_speed,
                             const double* t0, int64_t t0_stride, int32_t n_scen, int32_t n_obs, int32_t n_steps, double period,
                             double* obs_xy_t, frenet_stream_t stream);

/* Device-side `replan_ctot` bookkeeping for a batch (lib/planner/trajectory_planner.py:206-212 with optimal_at_time :183-195
 * and the interval rule :120): for every scenario, the new initial state is the gathered winner (buf->winner) sampled at
 * `time[b * time_stride]` - index 0 / last when the time is outside the winner's window, else round-half-even of
 * (time - t[0]) / lookup_period - t_initial becomes that time, and the target-speed axis is rebuilt from the new speed
 * (round(ds0 - d_d_s * num_sample), round(ds0 + d_d_s * num_sample + d_d_s), d_d_s) into lat->axis_v / lat->n_v.
 * Scenarios without a winner keep their state and get status[b] = 1 (the reference raises there); status may be NULL.
 * The buffers written are the `state`, `scal`, `axis_v`, `n_v` inputs themselves (cast away const: they are device memory). */
int frenet_advance_state(const FrenetLattice* lat, const FrenetBuffers* buf, const double* time, int64_t time_stride,
                         double lookup_period, double d_d_s, int32_t num_sample, int32_t follow, int32_t* status, frenet_stream_t stream);
/* (follow != 0, or buf->follow[b] != 0 where that array is given: the scenario's axis is si_interval - target offsets centred on 0,
 * trajectory_planner.py:119 - instead of the target speeds around its new speed; status[b] = 2: the axis did not fit n_v_cap) */

/* Follow-mode target triples on the device, V1 semantics (lib/planner/trajectory_planner.py:131-136) with an
 * ObstacleTrajectory target (lib/sensor/dynamic_obstacle.py:43-77): for every scenario b with buf->follow[b] != 0 (and only[b] != 0
 * when `only` is given) and every horizon T, buf->target[b][iT] = (st, dst, ddst) computed from the scenario's state, t_initial and
 * its leading obstacle lead[b] (index into obs_s / obs_d / obs_speed [n_scen][n_obs]: time offset, lateral offset, speed of the
 * obstacle's time law).  `estimate` [n_scen]: arc length of the robot's projection onto the path (frenet_sim_project) - it defines the
 * live transformer frame compute_s / compute_ds / compute_dds project into. */
int frenet_follow_targets(const FrenetPath* path, const FrenetLattice* lat, const FrenetBuffers* buf, const int32_t* lead, const double* obs_s,
                          const double* obs_d, const double* obs_speed, int32_t n_obs, const double* estimate, double delta_t, double d_target,
                          const uint8_t* only, frenet_stream_t stream);

/* The all-blocked branch of the moving-obstacle driver for a batch (lib/tests/test_moving_obstacle_planner.py:92-97): scenarios whose
 * result record shows candidates but no survivor get redo[b] = 1, follow[b] = 1, lead[b] = first blocker of their first candidate
 * (`leading_obstacles[0]`) and the follow-mode axis si_interval in lat->axis_v / lat->n_v; all others redo[b] = 0.  The caller then
 * runs frenet_follow_targets(only = redo) and frenet_plan with buf->scen_mask = redo, without the collision stage and with use_mask 0
 * (the reference takes the minimum over the UNFILTERED follow lattice).  lead / follow / redo: device arrays [n_scen]. */
int frenet_follow_select(const FrenetLattice* lat, const FrenetBuffers* buf, double d_d_s, int32_t num_sample, int32_t* lead, uint8_t* follow,
                         uint8_t* redo, frenet_stream_t stream);

/* Target triples of the Optimal variant on the device (lib/optimal_frenet_frame/target.py:17-52, trajectory_planner_optimal.py:130-133):
 * spec [n_scen][16] = leading quintic s0[3], s1[3], Ts, DTs, mode (0 plain, 1 follow, 2 merge, 3 stop), D0, second target s0[3], s1[3];
 * buf->target[b][iT] = the sampled target at index round((t_initial + T) / delta_t).  status [n_scen] (may be NULL): 0 ok, 1 index
 * outside the sampled target (the reference raises IndexError), 2 stop target does not end at rest (the reference asserts). */
int frenet_optimal_targets(const FrenetLattice* lat, const FrenetBuffers* buf, const double* spec, double delta_t, int32_t* status,
                           frenet_stream_t stream);

/* Robot side of the closed loop for a batch of agents; per-agent scalar work kept on the device so
 * that a simulation step needs no host round trip.
 *
 * frenet_sim_project - FrenetGNTransformOld.estimatePosition + transform (lib/transform/gn_transform_old.py:18-76): `max_iter`
 * damped Gauss-Newton steps on the arc length estimate (in/out, [n]), then the robot pose [n][3] expressed in the Frenet
 * frame anchored there -> fpose [n][3]; curvature [n] = trajectory.compute_curvature(estimate).
 *
 * frenet_sim_control_step - FrenetIOLController.compute (lib/controller/ctrl_frenet_iol.py:63-103) with the planner's
 * output (state [n][6] = p0, dp0, ddp0, s0, ds0, dds0 as returned by replanner) followed by Unicycle.step
 * (lib/platform/unicycle.py:16-81): pose [n][3] is advanced in place, control [n][2] = (v, omega). */
int frenet_sim_project(const FrenetPath* path, const double* pose, double* estimate, int32_t max_iter, double damping, double* fpose,
                       double* curvature, int32_t n, frenet_stream_t stream);
int frenet_sim_control_step(const double* fpose, const double* state, const double* curvature, double b, double kp_s, double kp_d,
                            double dt, double* pose, double* control, int32_t n, frenet_stream_t stream);

/* Device-side cone gate of the proximity sensor: active[b][o] = 1 iff obstacle o of scenario b lies inside the cone of
 * length `range` and half-aperture `aperture` ahead of robot_pose[b] = (x, y, theta).
 * Replaces ProximitySensor.sense (lib/sensor/proximity_sensor.py:41-75) for a whole batch; obs_xy: [n_scen][n_obs][2]. */
int frenet_sensor_gate(const double* robot_pose, const double* obs_xy, int32_t n_scen, int32_t n_obs, double range, double aperture,
                       uint8_t* active, frenet_stream_t stream);

/* Generic form of K4 for polylines that are not in the lattice layout: path p owns points offsets[p] .. offsets[p+1]-1
 * of x / y; obstacles are an [n_obs][2] table (all tested).  Replaces check_collisions on one path at a time
 * (lib/tests/test_moving_obstacle_planner.py:42-52) for every path of a list at once. */
int frenet_polyline_collision(const double* x, const double* y, const int64_t* offsets, int32_t n_paths, const double* obs_xy,
                              int32_t n_obs, double radius_sq, uint8_t* coll_free, int32_t* first_blocker, frenet_stream_t stream);

/* Buffer sizing for hosts that do not use the Python engine (SURVEY section 8b "frenet_plan_workspace_bytes"): the byte size of
 * every array FrenetLattice / FrenetBuffers point to, for a lattice shape and obstacle capacity.  Only the size fields of `lat`
 * are read (n_scen, n_v_cap, n_t, n_d, n_pad_max, row_elems, cand_elems); its pointers may be NULL.  The layout inside each
 * array is documented at the top of this header.  `total` = the sum with every array rounded up to 256 bytes, i.e. the size of
 * ONE allocation that holds them all back to back in the order of the fields below. */
typedef struct FrenetSizes {
    int64_t rows, candidates;     /* n_scen * n_v_cap * n_t, times n_d */
    /* inputs */
    int64_t state, scal, target, axis_v, n_v, axis_d, axis_t, n_steps, t_offset, t_pow;
    int64_t obs_xy, obs_active, obs_xy_t, obs_bounds;
    /* outputs */
    int64_t coef_lon, coef_lat, row_S, row_flags, lon, lat, cost, cand_flags, curv_ok, coll_free, first_blocker, road_ok;
    int64_t result, winner, winner_steps;
    /* optional per-scenario lane inputs */
    int64_t path_origin, lane_params, road_params;
    /* optional scratch of FRENET_STAGE_ONE_LAUNCH */
    int64_t tickets, partials;
    int64_t total;
} FrenetSizes;
int frenet_workspace_bytes(const FrenetLattice* lat, int32_t n_obs, int32_t n_obs_steps, FrenetSizes* out);

/* How frenet_k2_evaluate / frenet_k2_fused will launch for a shape: rows of a scenario per CTA (short-horizon kernel, > 1 only for
 * lattices of at most 32 padded samples), CTAs per row (long-horizon kernel on small batches), grid size and dynamic shared
 * memory per CTA.  path NULL = K2 alone.  Returns FRENET_ERR_CAPACITY if the shape does not fit 227 KB of shared memory.
 * Any output pointer may be NULL. */
int frenet_k2_launch_shape(const FrenetLattice* lat, const FrenetPath* path, int32_t n_obs, int32_t n_obs_steps, int32_t with_collision,
                           int32_t* rows_per_cta, int32_t* ctas_per_row, int64_t* grid, int64_t* shared_bytes);

#ifdef __cplusplus
}
#endif
#endif /* FRENET_B200_H */
