/* A host WITHOUT Python or PyTorch: one replan of the reference's default lattice through the C ABI alone.
 *
 *   nvcc -o examples/replan_c_host examples/replan_c_host.c -I include \
 *        -L optimal-trajectory-generation-in-the-duckietown-environment_b200 -lfrenet_b200 \
 *        -Xlinker -rpath -Xlinker $PWD/optimal-trajectory-generation-in-the-duckietown-environment_b200
 *   ./examples/replan_c_host [p0 dp0 ddp0 s0 ds0 dds0 t_initial]
 *
 * What it does is what TrajectoryPlannerV1.replanner does (lib/planner/trajectory_planner.py:108-181, 206-212) for the default
 * parameters (trajectory_planner.py:14-67): lattice D = arange(-4, 4, 0.5), T = arange(1, 3, 0.5), target speeds
 * arange(round(ds0 - 2), round(ds0 + 3), 1), sample period 0.1, then min(paths, key=ctot).  Everything numeric happens in
 * frenet_replan_inline (ONE kernel: boundary-value systems, sampling, costs, argmin, winner gather); the host only sizes the buffers
 * (frenet_workspace_bytes), uploads the lattice's shape and reads the record + winner from mapped host memory.
 * tests/test_gpu_planner.py::test_c_host_example_matches_the_python_planner runs it against the Python planner. */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "frenet_b200.h"

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e_ = (x);                                                      \
        if (e_ != cudaSuccess) {                                                   \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));               \
            return 2;                                                              \
        }                                                                          \
    } while (0)
#define FK(x)                                                                      \
    do {                                                                           \
        if ((x) != FRENET_OK) {                                                    \
            fprintf(stderr, "%s: %s\n", #x, frenet_last_error());                  \
            return 3;                                                              \
        }                                                                          \
    } while (0)

static void* dev_bytes(size_t n) {
    void* p = NULL;
    if (n == 0) n = 8;
    if (cudaMalloc(&p, n) != cudaSuccess) return NULL;
    cudaMemset(p, 0, n);
    return p;
}

int main(int argc, char** argv) {
    double state[6] = {0.2, 0.1, 0.0, 3.0, 2.2, 0.1};   /* p0, dp0, ddp0, s0, ds0, dds0 */
    double t_initial = 0.1;
    if (argc >= 8) {
        for (int i = 0; i < 6; ++i) state[i] = atof(argv[1 + i]);
        t_initial = atof(argv[7]);
    }
    /* ---- the lattice's shape, exactly as the reference builds it on the host ---- */
    enum { ND = 16, NT = 4, NV_CAP = 6 };
    const double period = 0.1;
    double axis_d[ND], axis_t[NT];
    for (int i = 0; i < ND; ++i) axis_d[i] = -4.0 + 0.5 * i;   /* np.arange(-W, W, d_road_width)      (:94)  */
    for (int i = 0; i < NT; ++i) axis_t[i] = 1.0 + 0.5 * i;    /* np.arange(min_t, max_t, dt)         (:95)  */
    const int32_t n_steps[NT] = {11, 16, 21, 26};              /* len(np.arange(0, T + delta_t, delta_t)) (:139,148) */
    int32_t t_offset[NT + 1] = {0};
    int n_pad_max = 0;
    for (int i = 0; i < NT; ++i) {
        const int npad = (n_steps[i] + 3) / 4 * 4;             /* whole 32-byte sectors */
        t_offset[i + 1] = t_offset[i] + npad;
        if (npad > n_pad_max) n_pad_max = npad;
    }
    double* t_pow = (double*)malloc(sizeof(double) * 4 * n_pad_max);
    for (int k = 0; k < n_pad_max; ++k)
        for (int n = 0; n < 4; ++n) t_pow[4 * k + n] = pow((double)k * period, (double)(n + 2));   /* the reference's t ** n */
    double axis_v[FRENET_INLINE_MAX_V] = {0};
    const double lo = nearbyint(state[4] - 2.0), hi = nearbyint(state[4] + 2.0 + 1.0);           /* round(): half to even (:120) */
    int32_t n_v = 0;
    for (double v = lo; v < hi && n_v < NV_CAP; v += 1.0) axis_v[n_v++] = v;
    const double scal[3] = {0.0 /* dd */, 2.5 /* desired_speed */, t_initial};

    FrenetLattice L;
    memset(&L, 0, sizeof(L));
    L.n_scen = 1; L.n_v_cap = NV_CAP; L.n_t = NT; L.n_d = ND; L.n_pad_max = n_pad_max;
    L.row_elems = (int64_t)NV_CAP * t_offset[NT];
    L.cand_elems = L.row_elems * ND;
    FrenetSizes sz;
    FK(frenet_workspace_bytes(&L, /*n_obs*/ 0, /*n_obs_steps*/ 0, &sz));

    /* geometry on the device */
    double *d_axis_d = (double*)dev_bytes(sizeof(axis_d)), *d_axis_t = (double*)dev_bytes(sizeof(axis_t));
    double* d_tpow = (double*)dev_bytes(sizeof(double) * 4 * n_pad_max);
    int32_t *d_nsteps = (int32_t*)dev_bytes(sizeof(n_steps)), *d_toff = (int32_t*)dev_bytes(sizeof(t_offset));
    CK(cudaMemcpy(d_axis_d, axis_d, sizeof(axis_d), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_axis_t, axis_t, sizeof(axis_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_tpow, t_pow, sizeof(double) * 4 * n_pad_max, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_nsteps, n_steps, sizeof(n_steps), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_toff, t_offset, sizeof(t_offset), cudaMemcpyHostToDevice));
    L.axis_d = d_axis_d; L.axis_t = d_axis_t; L.n_steps = d_nsteps; L.t_offset = d_toff; L.t_pow = d_tpow;
    L.axis_v = (const double*)dev_bytes(sz.axis_v);      /* filled by the kernel from the inline argument */
    L.n_v = (const int32_t*)dev_bytes(sz.n_v);

    FrenetBuffers B;
    memset(&B, 0, sizeof(B));
    B.state = (const double*)dev_bytes(sz.state);
    B.scal = (const double*)dev_bytes(sz.scal);
    B.coef_lon = (double*)dev_bytes(sz.coef_lon);
    B.coef_lat = (double*)dev_bytes(sz.coef_lat);
    B.row_S = (double*)dev_bytes(sz.row_S);
    B.row_flags = (int32_t*)dev_bytes(sz.row_flags);
    B.lon = (double*)dev_bytes(sz.lon);
    B.lat = (double*)dev_bytes(sz.lat);
    B.cost = (double*)dev_bytes(sz.cost);
    B.cand_flags = (uint8_t*)dev_bytes(sz.cand_flags);
    B.tickets = (int32_t*)dev_bytes(sz.tickets);
    B.partials = dev_bytes(sz.partials);
    /* the results go straight to mapped host memory: the kernel's posted writes, no copy back */
    FrenetResult* result = NULL;
    double* winner = NULL;
    int32_t* winner_steps = NULL;
    CK(cudaHostAlloc((void**)&result, sz.result, cudaHostAllocMapped));
    CK(cudaHostAlloc((void**)&winner, sz.winner, cudaHostAllocMapped));
    CK(cudaHostAlloc((void**)&winner_steps, sz.winner_steps, cudaHostAllocMapped));
    B.result = result; B.winner = winner; B.winner_steps = winner_steps;

    FrenetParams P;
    memset(&P, 0, sizeof(P));
    P.kj = 0.001; P.kt = 0.01; P.kd = 0.5; P.ks = 0.5; P.kdots = 1.0; P.klat = 1.0; P.klong = 1.0;
    P.sample_period = period; P.low_speed_threshold = 2.0; P.s_thresh = 0.0;
    P.v_max = INFINITY; P.a_max = INFINITY; P.kappa_max = INFINITY; P.radius_sq = 0.49;
    P.lowspeed_rule = FRENET_LOWSPEED_V1; P.follow = 0;

    const int stages = FRENET_STAGE_K1 | FRENET_STAGE_K2 | FRENET_STAGE_K5 | FRENET_STAGE_GATHER | FRENET_STAGE_ONE_LAUNCH;
    FK(frenet_replan_inline(&L, &P, NULL, &B, stages, /*use_mask*/ 0, FRENET_SEL_CTOT, state, scal, axis_v, &n_v, NULL, /*wait*/ 1, NULL));

    const int n = winner_steps[0];
    printf("abi %d n_v %d n_valid %d argmin_ctot %d min_ctot %.17g steps %d\n", frenet_abi_version(), n_v, result[0].n_valid,
           result[0].argmin_ctot, result[0].min_ctot, n);
    /* the six numbers the NEXT replan starts from: optimal_at_time(t_initial + 0.1) = sample 1 of the winner (rows: t, d, d', d'', d''', s, s', s'', ...) */
    if (n > 1)
        printf("next %.17g %.17g %.17g %.17g %.17g %.17g\n", winner[1 * n_pad_max + 1], winner[2 * n_pad_max + 1], winner[3 * n_pad_max + 1],
               winner[5 * n_pad_max + 1], winner[6 * n_pad_max + 1], winner[7 * n_pad_max + 1]);
    return 0;
}
