"""ctypes binding of the C ABI declared in include/frenet_b200.h.

The library is the product: if `libfrenet_b200.so` is missing this module raises at import
time, and there is no other implementation to fall back to (the NumPy oracle under oracle/ is
test infrastructure and is never imported from here).
"""
from __future__ import annotations

import ctypes as C
import os

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# OTG_B200_LIB selects another build of the same library (A/B experiments with compile-time tuning knobs)
LIB_PATH = os.environ.get("OTG_B200_LIB") or os.path.join(_PKG_DIR, "libfrenet_b200.so")

ABI_VERSION = 3

OK, ERR_BAD_ARG, ERR_CUDA, ERR_CAPACITY = 0, -1, -2, -3
STATUS_OK, STATUS_NO_FEASIBLE, STATUS_EMPTY = 0, 1, 2
LOWSPEED_OFF, LOWSPEED_V1, LOWSPEED_OPTIMAL = 0, 1, 2
ROW_EXISTS, ROW_LOWSPEED, ROW_DROPPED = 1, 2, 4
CAND_VALID, CAND_KINEMATIC_OK = 1, 2
PATH_CIRCLE, PATH_SPLINE, PATH_PARABOLA = 0, 1, 2
SEL_CTOT, SEL_CD, SEL_CV, SEL_CT, SEL_MASKED = 0, 1, 2, 3, 4
STAGE_K1, STAGE_K2, STAGE_K3, STAGE_K4, STAGE_K5, STAGE_GATHER, STAGE_CURVATURE, STAGE_NO_FUSE = 1, 2, 4, 8, 16, 32, 64, 128
STAGE_OFFROAD = 256
STAGE_ONE_LAUNCH = 512
PARTIAL_BYTES, MAX_ROW_PARTS = 72, 8
INLINE_MAX_V, INLINE_MAX_T = 16, 24
FUSE_COLLISION, FUSE_CURVATURE = 1, 2
MASK_KINEMATIC, MASK_CURVATURE, MASK_COLLISION, MASK_ROAD = 1, 2, 4, 8

c_double_p = C.c_void_p   # device pointers are passed as integers (tensor.data_ptr())


class FrenetParams(C.Structure):
    _fields_ = [(k, C.c_double) for k in (
        "kj", "kt", "kd", "ks", "kdots", "klat", "klong", "sample_period", "low_speed_threshold",
        "s_thresh", "v_max", "a_max", "kappa_max", "radius_sq")] + [("lowspeed_rule", C.c_int32), ("follow", C.c_int32)]


class FrenetLattice(C.Structure):
    _fields_ = [("n_scen", C.c_int32), ("n_v_cap", C.c_int32), ("n_t", C.c_int32), ("n_d", C.c_int32),
                ("n_pad_max", C.c_int32), ("lat_f32", C.c_int32), ("row_elems", C.c_int64), ("cand_elems", C.c_int64),
                ("axis_d", C.c_void_p), ("axis_t", C.c_void_p), ("n_steps", C.c_void_p), ("t_offset", C.c_void_p),
                ("t_pow", C.c_void_p), ("axis_v", C.c_void_p), ("n_v", C.c_void_p)]


class FrenetPath(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_knots", C.c_int32), ("circle", C.c_double * 16), ("spline_table", C.c_void_p)]


class FrenetResult(C.Structure):
    _fields_ = [("argmin_ctot", C.c_int32), ("argmin_cd", C.c_int32), ("argmin_cv", C.c_int32), ("argmin_ct", C.c_int32),
                ("argmin_masked", C.c_int32), ("n_valid", C.c_int32), ("n_survivors", C.c_int32), ("status", C.c_int32),
                ("min_ctot", C.c_double), ("min_masked_ctot", C.c_double), ("reserved", C.c_double * 2)]


# numpy view of FrenetResult records
RESULT_DTYPE = [("argmin_ctot", "<i4"), ("argmin_cd", "<i4"), ("argmin_cv", "<i4"), ("argmin_ct", "<i4"),
                ("argmin_masked", "<i4"), ("n_valid", "<i4"), ("n_survivors", "<i4"), ("status", "<i4"),
                ("min_ctot", "<f8"), ("min_masked_ctot", "<f8"), ("reserved", "<f8", (2,))]
RESULT_BYTES = 64


class FrenetBuffers(C.Structure):
    _fields_ = [("state", C.c_void_p), ("scal", C.c_void_p), ("target", C.c_void_p), ("obs_xy", C.c_void_p),
                ("obs_active", C.c_void_p), ("n_obs", C.c_int32), ("n_obs_steps", C.c_int32),
                ("obs_xy_t", C.c_void_p), ("obs_bounds", C.c_void_p),
                ("coef_lon", C.c_void_p), ("coef_lat", C.c_void_p), ("row_S", C.c_void_p), ("row_flags", C.c_void_p),
                ("lon", C.c_void_p), ("lat", C.c_void_p), ("cost", C.c_void_p), ("cand_flags", C.c_void_p),
                ("curv_ok", C.c_void_p),
                ("coll_free", C.c_void_p), ("first_blocker", C.c_void_p),
                ("result", C.c_void_p),
                ("winner", C.c_void_p), ("winner_steps", C.c_void_p),
                ("path_origin", C.c_void_p), ("road_ok", C.c_void_p),
                ("lane_params", C.c_void_p), ("road_params", C.c_void_p),
                ("follow", C.c_void_p), ("scen_mask", C.c_void_p), ("tickets", C.c_void_p), ("winner_masked", C.c_void_p),
                ("winner_masked_steps", C.c_void_p), ("partials", C.c_void_p)]


class FrenetSizes(C.Structure):
    _fields_ = [(k, C.c_int64) for k in (
        "rows", "candidates",
        "state", "scal", "target", "axis_v", "n_v", "axis_d", "axis_t", "n_steps", "t_offset", "t_pow",
        "obs_xy", "obs_active", "obs_xy_t", "obs_bounds",
        "coef_lon", "coef_lat", "row_S", "row_flags", "lon", "lat", "cost", "cand_flags", "curv_ok", "coll_free", "first_blocker", "road_ok",
        "result", "winner", "winner_steps",
        "path_origin", "lane_params", "road_params", "tickets", "partials", "total")]


class FrenetRoad(C.Structure):
    _fields_ = [("right", C.c_double * 3), ("left", C.c_double * 3), ("has_right", C.c_int32), ("has_left", C.c_int32)]


assert C.sizeof(FrenetResult) == RESULT_BYTES


class FrenetError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"libfrenet_b200 error {code}: {text}")
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback for the planner hot path.")
    lib = C.CDLL(LIB_PATH)
    P = C.POINTER
    sigs = {
        "frenet_abi_version": (C.c_int, []),
        "frenet_last_error": (C.c_char_p, []),
        "frenet_device_count": (C.c_int, []),
        "frenet_k1_coefficients": (C.c_int, [P(FrenetLattice), P(FrenetParams), P(FrenetBuffers), C.c_void_p]),
        "frenet_k2_evaluate": (C.c_int, [P(FrenetLattice), P(FrenetParams), P(FrenetBuffers), C.c_void_p]),
        "frenet_k2_fused": (C.c_int, [P(FrenetLattice), P(FrenetParams), P(FrenetPath), P(FrenetBuffers), C.c_int32, C.c_void_p]),
        "frenet_k3_cartesian": (C.c_int, [P(FrenetLattice), P(FrenetParams), P(FrenetPath), P(FrenetBuffers), C.c_void_p]),
        "frenet_k3_curvature": (C.c_int, [P(FrenetLattice), P(FrenetParams), P(FrenetBuffers), C.c_void_p]),
        "frenet_k3_offroad": (C.c_int, [P(FrenetLattice), P(FrenetRoad), P(FrenetBuffers), C.c_void_p]),
        "frenet_k4_collision": (C.c_int, [P(FrenetLattice), P(FrenetParams), P(FrenetBuffers), C.c_void_p]),
        "frenet_k5_argmin": (C.c_int, [P(FrenetLattice), P(FrenetBuffers), C.c_int32, C.c_void_p]),
        "frenet_gather_winner": (C.c_int, [P(FrenetLattice), P(FrenetParams), P(FrenetBuffers), C.c_int32, C.c_int32, C.c_void_p]),
        "frenet_plan": (C.c_int, [P(FrenetLattice), P(FrenetParams), P(FrenetPath), P(FrenetBuffers), C.c_int32, C.c_int32,
                                  C.c_int32, C.c_void_p]),
        "frenet_stream_synchronize": (C.c_int, [C.c_void_p]),
        "frenet_replan_inline": (C.c_int, [P(FrenetLattice), P(FrenetParams), P(FrenetPath), P(FrenetBuffers), C.c_int32, C.c_int32,
                                           C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
        "frenet_predict_obstacles": (C.c_int, [P(FrenetPath), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                               C.c_int32, C.c_int32, C.c_double, C.c_void_p, C.c_void_p]),
        "frenet_advance_state": (C.c_int, [P(FrenetLattice), P(FrenetBuffers), C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_int32,
                                           C.c_int32, C.c_void_p, C.c_void_p]),
        "frenet_follow_targets": (C.c_int, [P(FrenetPath), P(FrenetLattice), P(FrenetBuffers), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_int32, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p]),
        "frenet_follow_select": (C.c_int, [P(FrenetLattice), P(FrenetBuffers), C.c_double, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p]),
        "frenet_optimal_targets": (C.c_int, [P(FrenetLattice), P(FrenetBuffers), C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]),
        "frenet_sim_project": (C.c_int, [P(FrenetPath), C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_int32,
                                         C.c_void_p]),
        "frenet_sim_control_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double,
                                              C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
        "frenet_sensor_gate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_void_p, C.c_void_p]),
        "frenet_polyline_collision": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_double,
                                                C.c_void_p, C.c_void_p, C.c_void_p]),
        "frenet_points_to_cartesian": (C.c_int, [P(FrenetPath), C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                                 C.c_int64, C.c_void_p]),
        "frenet_workspace_bytes": (C.c_int, [P(FrenetLattice), C.c_int32, C.c_int32, P(FrenetSizes)]),
        "frenet_k2_launch_shape": (C.c_int, [P(FrenetLattice), P(FrenetPath), C.c_int32, C.c_int32, C.c_int32, P(C.c_int32), P(C.c_int32),
                                             P(C.c_int64), P(C.c_int64)]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)    # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    if lib.frenet_abi_version() != ABI_VERSION:
        raise ImportError(f"libfrenet_b200.so has ABI {lib.frenet_abi_version()}, binding expects {ABI_VERSION}: rebuild")
    return lib, tuple(sigs)


lib, EXPORTS = _load()


def check(rc: int):
    if rc != OK:
        raise FrenetError(rc, lib.frenet_last_error().decode(errors="replace"))


def device_count() -> int:
    return int(lib.frenet_device_count())
