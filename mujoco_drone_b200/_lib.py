"""ctypes binding of libdronesim_b200.so (C ABI: include/dronesim_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is present, loading or
`dsim_create` raises.  Build the library in-tree with `python -m mujoco_drone_b200.build` (nvcc, sm_100a).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DSIM_LIB") or os.path.join(_HERE, "libdronesim_b200.so")   # DSIM_LIB: kernel-variant experiments

OK, EINVAL, ECUDA, ENOMEM, EUNSUPPORTED, ESHAPE = 0, -1, -2, -3, -4, -5
FP32, FP64 = 0, 1
ABI_VERSION = 3
TRAJ_CIRCLE, TRAJ_STEP, TRAJ_RAMP = 0, 1, 2
TILE = 32                     # envs per page (include/dronesim_b200.h: PAGED buffers)
(BUF_STATE, BUF_NUM_STEPS, BUF_OBS, BUF_REWARD, BUF_TRUNCATED, BUF_PARAMS, BUF_CONSTS, BUF_REFERENCE,
 BUF_RESET_COUNT, BUF_STATES33, BUF_EP_RETURN, BUF_STATS, BUF_SENSORDATA, BUF_GEOMETRY) = range(14)
DT_F32, DT_F64, DT_I32, DT_U8, DT_U32, DT_I64 = range(6)

EXPORTS = [
    "dsim_abi_version", "dsim_create", "dsim_destroy", "dsim_last_error", "dsim_obs_dim", "dsim_regen_params",
    "dsim_set_params", "dsim_get_params", "dsim_get_consts", "dsim_reset_all", "dsim_reset_masked", "dsim_reset_at",
    "dsim_forward", "dsim_zero_act", "dsim_step", "dsim_evaluate", "dsim_step_host", "dsim_set_inputs_ready", "dsim_set_reference", "dsim_control_reference",
    "dsim_set_state", "dsim_get_state", "dsim_compute_states", "dsim_buffer", "dsim_stats", "dsim_sync",
    "dsim_launch_count", "dsim_debug_guard_check", "dsim_kernel_info", "dsim_debug_timeline", "dsim_beta_policy", "dsim_trajectory_reference", "dsim_policy_blob_sizes", "dsim_policy_create", "dsim_policy_destroy", "dsim_policy_forward", "dsim_policy_forward_sample", "dsim_policy_error", "dsim_policy32_blob_elems", "dsim_policy32_create", "dsim_policy32_destroy", "dsim_policy32_forward",
]


class DsimConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("abi_version", C.c_int32), ("num_envs", C.c_int32), ("precision", C.c_int32),
        ("env_id_offset", C.c_int64), ("seed", C.c_uint32), ("pendulum", C.c_int32), ("frame_skip", C.c_int32),
        ("round_precision", C.c_int32), ("frequency", C.c_double), ("obs_id", C.c_int32), ("reward_id", C.c_int32),
        ("ground_contact", C.c_int32), ("per_env_reference", C.c_int32), ("auto_reset", C.c_int32),
        ("random_start_pos", C.c_int32), ("random_params", C.c_int32),
        ("reference", C.c_double * 4), ("start_pos", C.c_double * 4), ("max_distance", C.c_double),
        ("max_steps", C.c_int64), ("max_pos_offset", C.c_double),
        ("angle_sigma", C.c_double * 2), ("vel_sigma", C.c_double * 3), ("ang_vel_sigma", C.c_double * 3),
        ("pend_rp_sigma", C.c_double * 2), ("pend_vel_sigma", C.c_double * 2),
        ("param_center", C.c_double * 6), ("param_halfwidth", C.c_double * 6), ("param_difficulty", C.c_double),
    ]


class DsimError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"dronesim_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """Load the shared library; raise (never fall back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run `python -m mujoco_drone_b200.build` "
            "(needs nvcc). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, dp, i32 = C.c_void_p, C.POINTER(C.c_double), C.c_int32
    L.dsim_abi_version.restype = C.c_int
    L.dsim_create.argtypes = [C.POINTER(DsimConfig), C.c_int, C.POINTER(vp)]
    L.dsim_destroy.argtypes = [vp]
    L.dsim_destroy.restype = None
    L.dsim_last_error.argtypes = [vp]
    L.dsim_last_error.restype = C.c_char_p
    L.dsim_obs_dim.argtypes = [C.c_int, C.c_int]
    L.dsim_regen_params.argtypes = [vp, C.c_uint32, vp]
    L.dsim_set_params.argtypes = [vp, dp, vp]
    L.dsim_get_params.argtypes = [vp, dp]
    L.dsim_get_consts.argtypes = [vp, dp]
    L.dsim_reset_all.argtypes = [vp, vp]
    L.dsim_reset_masked.argtypes = [vp, vp, vp]
    L.dsim_reset_at.argtypes = [vp, C.c_int, vp]
    L.dsim_forward.argtypes = [vp, C.c_int, vp]
    L.dsim_zero_act.argtypes = [vp, vp]
    L.dsim_step.argtypes = [vp, vp, vp]
    L.dsim_evaluate.argtypes = [vp, vp, vp]
    L.dsim_step_host.argtypes = [vp, vp, vp, vp, vp, vp]
    L.dsim_set_reference.argtypes = [vp, dp]
    L.dsim_set_inputs_ready.argtypes = [vp, C.c_int]
    L.dsim_control_reference.argtypes = [vp, vp, vp]
    L.dsim_set_state.argtypes = [vp, dp, dp, dp, C.POINTER(i32), vp]
    L.dsim_get_state.argtypes = [vp, dp, dp, dp, dp, C.POINTER(i32)]
    L.dsim_compute_states.argtypes = [vp, vp]
    L.dsim_buffer.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                              C.POINTER(C.c_int64), C.POINTER(i32), C.POINTER(C.c_int64)]
    L.dsim_stats.argtypes = [vp, dp, C.c_int]
    L.dsim_sync.argtypes = [vp, vp]
    L.dsim_launch_count.argtypes = [vp]
    L.dsim_launch_count.restype = C.c_int64
    L.dsim_debug_guard_check.argtypes = [vp]
    L.dsim_debug_guard_check.restype = C.c_int64
    L.dsim_debug_timeline.argtypes = [vp, C.POINTER(C.c_uint64), C.c_int64]
    L.dsim_beta_policy.argtypes = [vp, C.c_int, C.c_int, C.c_uint32, C.c_int64, C.c_uint32, vp, C.c_int, vp, vp, vp]
    L.dsim_policy_blob_sizes.argtypes = [C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.dsim_policy_create.argtypes = [C.c_int, vp, vp, C.POINTER(vp)]
    L.dsim_policy_destroy.argtypes = [vp]
    L.dsim_policy_destroy.restype = None
    L.dsim_policy_forward.argtypes = [vp, vp, vp, vp, C.c_int, vp, vp, vp]
    L.dsim_policy_forward_sample.argtypes = [vp, vp, vp, vp, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, vp, C.c_int, vp, vp, vp, vp, vp]
    L.dsim_policy_error.argtypes = [vp]
    L.dsim_policy32_blob_elems.restype = C.c_int64
    L.dsim_policy32_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
    L.dsim_policy32_destroy.argtypes = [vp]
    L.dsim_policy32_destroy.restype = None
    L.dsim_policy32_forward.argtypes = [vp, vp, vp, vp, C.c_int, vp, vp, vp]
    L.dsim_trajectory_reference.argtypes = [vp, C.c_int, C.c_double, C.c_double, dp, dp, dp, vp]
    L.dsim_kernel_info.argtypes = [C.c_int, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    if L.dsim_abi_version() != ABI_VERSION:
        raise ImportError("libdronesim_b200.so ABI version mismatch; rebuild with `python -m mujoco_drone_b200.build`")
    _lib = L
    return L


def check(handle, rc):
    if rc != OK:
        msg = load().dsim_last_error(handle)
        raise DsimError(rc, msg.decode() if msg else "")
    return rc
