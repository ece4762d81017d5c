"""ctypes bindings of include/drc_b200.h (the C ABI of libdrc_b200.so).

The library is built in-tree by dyros_robot_controller_b200.build.  Loading fails loudly when the
shared object is missing: there is no Python or CPU fallback for the compute entry points.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "libdrc_b200.so"
_lib = None

DRC_OK = 0
LAYOUT_AOS, LAYOUT_SOA = 0, 1
QP_UNSOLVED, QP_SOLVED, QP_MAX_ITER, QP_PRIMAL_INFEASIBLE, QP_DUAL_INFEASIBLE, QP_NON_CONVEX, QP_SOLVED_INACCURATE = range(7)


class DrcParams(C.Structure):
    _fields_ = [
        ("alpha", C.c_double), ("slack_weight", C.c_double), ("ik_reg", C.c_double), ("moma_ik_reg", C.c_double),
        ("mani_thresh", C.c_double), ("dist_thresh", C.c_double),
        ("Kp_task", C.c_double * 6), ("Kv_task", C.c_double * 6),
        ("Kp_joint", C.c_double * 16), ("Kv_joint", C.c_double * 16),
        ("rho", C.c_double), ("sigma", C.c_double), ("osqp_alpha", C.c_double), ("eps_abs", C.c_double),
        ("eps_rel", C.c_double), ("eps_prim_inf", C.c_double), ("eps_dual_inf", C.c_double),
        ("max_iter", C.c_int), ("check_termination", C.c_int), ("scaling", C.c_int), ("adaptive_rho", C.c_int),
        ("adaptive_rho_interval", C.c_int),
        ("adaptive_rho_tolerance", C.c_double), ("gjk_tol", C.c_double), ("epa_tol", C.c_double),
        ("gjk_max_iter", C.c_int), ("epa_max_iter", C.c_int), ("pinv_threshold", C.c_double),
        ("schedule_hint", C.c_int),
        ("rollout_fused", C.c_int),
        ("rollout_warm_start", C.c_int),
    ]


# every symbol include/drc_b200.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "drc_last_error", "drc_version", "drc_device_count",
    "drc_model_create_from_urdf", "drc_model_create_from_text", "drc_model_destroy", "drc_model_dof",
    "drc_model_frame_id", "drc_model_num_frames", "drc_model_frame_name", "drc_model_joint_name", "drc_model_limits",
    "drc_model_info", "drc_model_mesh_info", "drc_model_verbose",
    "drc_ctx_create", "drc_ctx_destroy", "drc_ctx_get_params", "drc_ctx_set_params", "drc_ctx_max_batch",
    "drc_ctx_synchronize", "drc_ctx_stream",
    "drc_batch_update_state", "drc_batch_get_frame", "drc_batch_get_dynamics", "drc_batch_get_manipulability",
    "drc_batch_get_min_distance", "drc_batch_qpik", "drc_batch_qpik_step", "drc_batch_qpid", "drc_batch_qpid_step",
    "drc_batch_clik_step", "drc_batch_osf", "drc_batch_osf_step", "drc_batch_joint_torque_step",
    "drc_batch_task_space_cubic", "drc_batch_cycle_qpik_step", "drc_batch_cycle_qpid_step", "drc_batch_cycle_clik_osf_step",
    "drc_host_update_state", "drc_host_get_frame", "drc_host_get_dynamics", "drc_host_get_manipulability",
    "drc_host_get_min_distance", "drc_host_qpik", "drc_host_qpik_step", "drc_host_qpid", "drc_host_qpid_step",
    "drc_host_clik_step", "drc_host_osf", "drc_host_osf_step", "drc_host_joint_torque_step",
    "drc_host_task_space_cubic", "drc_host_cycle_qpik_step", "drc_host_cycle_qpid_step", "drc_host_cycle_clik_osf_step",
    "drc_ctx_enable_timing", "drc_ctx_last_timing", "drc_ctx_launch_count", "drc_bench_fp64_peak",
    "drc_ctx_last_trace", "drc_ctx_enable_qp_debug", "drc_host_get_qp_debug",
    "drc_model_attach_mobile_base", "drc_model_moma_info", "drc_model_base_jacobian",
    "drc_batch_moma_update_state", "drc_batch_moma_get_state", "drc_batch_moma_qpik", "drc_batch_moma_qpik_step",
    "drc_batch_moma_qpid", "drc_batch_moma_qpid_step", "drc_batch_moma_cycle_qpik_step", "drc_batch_moma_cycle_qpid_step",
    "drc_host_moma_update_state", "drc_host_moma_get_state", "drc_host_moma_qpik", "drc_host_moma_qpik_step",
    "drc_host_moma_qpid", "drc_host_moma_qpid_step", "drc_host_moma_cycle_qpik_step", "drc_host_moma_cycle_qpid_step",
    "drc_batch_rollout_qpik", "drc_host_rollout_qpik",
    "drc_mobile_create", "drc_mobile_destroy", "drc_mobile_wheel_num", "drc_mobile_synchronize", "drc_mobile_launch_count",
    "drc_batch_mobile_fk", "drc_batch_mobile_ik", "drc_host_mobile_fk", "drc_host_mobile_ik",
]


def lib_path() -> Path:
    return _LIB_PATH


def lib():
    """Load libdrc_b200.so (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise RuntimeError(
                f"{_LIB_PATH} is missing: build it with `python -m dyros_robot_controller_b200.build` "
                "(nvcc, sm_100a).  drc_b200 has no CPU fallback.")
        L = C.CDLL(str(_LIB_PATH))
        L.drc_last_error.restype = C.c_char_p
        L.drc_model_frame_name.restype = C.c_char_p
        L.drc_model_joint_name.restype = C.c_char_p
        L.drc_model_verbose.restype = C.c_char_p
        L.drc_ctx_stream.restype = C.c_void_p
        L.drc_ctx_launch_count.restype = C.c_longlong
        L.drc_mobile_launch_count.restype = C.c_longlong
        _lib = L
    return _lib


class DrcError(RuntimeError):
    def __init__(self, code: int, where: str):
        msg = lib().drc_last_error()
        super().__init__(f"{where} failed with code {code}: {msg.decode() if msg else ''}")
        self.code = code


def check(rc: int, where: str):
    if rc != DRC_OK:
        raise DrcError(rc, where)
