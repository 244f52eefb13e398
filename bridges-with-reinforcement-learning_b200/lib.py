"""ctypes binding of libbridges_b200.so (C ABI: include/bridges_b200.h).

The library is built in-tree by `__graft_entry__.build()` / `csrc/Makefile`.  Loading fails
loudly when it is missing: there is no CPU implementation behind this package.
"""
import ctypes as C
import os

BW_MAX_BLOCKS = 16
BW_MAX_FACES = 6
BW_MAX_VERTS = 6
BW_MAX_SHAPES = 8
BW_MAX_OBSTACLES = 8
BW_MAX_TARGETS = 4
BW_MAX_INTERFACES = 48
BW_IMG = 64
BW_ABI_VERSION = 6
BW_OK, BW_ERR_INVALID, BW_ERR_CUDA, BW_ERR_CAPACITY, BW_ERR_STATE = 0, -1, -2, -3, -4     # bw_status

LIB_PATH = os.environ.get("BRIDGES_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                                            "libbridges_b200.so")


class bw_config(C.Structure):
    _fields_ = [("num_envs", C.c_int32), ("device", C.c_int32), ("max_steps", C.c_int32), ("use_caller_stream", C.c_int32),
                ("xlim", C.c_double * 2), ("ylim", C.c_double * 2),
                ("floor_halfwidth", C.c_double), ("floor_depth", C.c_double),
                ("mu", C.c_double), ("density", C.c_double), ("tmax", C.c_double), ("amin", C.c_double),
                ("stable_tol", C.c_double), ("stream", C.c_void_p),
                ("collision_mode", C.c_int32), ("reserved0", C.c_int32), ("collision_tol", C.c_double),
                ("bounds_lo", C.c_double * 3), ("bounds_hi", C.c_double * 3)]


class bw_shape_desc(C.Structure):
    _fields_ = [("n_faces", C.c_int32), ("n_verts", C.c_int32),
                ("target_faces_mask", C.c_uint32), ("receiving_faces_mask", C.c_uint32),
                ("face_nx", C.c_double * BW_MAX_FACES), ("face_nz", C.c_double * BW_MAX_FACES),
                ("face_cx", C.c_double * BW_MAX_FACES), ("face_cz", C.c_double * BW_MAX_FACES),
                ("end0_x", C.c_double * BW_MAX_FACES), ("end0_z", C.c_double * BW_MAX_FACES),
                ("end1_x", C.c_double * BW_MAX_FACES), ("end1_z", C.c_double * BW_MAX_FACES),
                ("vert_x", C.c_double * BW_MAX_VERTS), ("vert_z", C.c_double * BW_MAX_VERTS),
                ("com_x", C.c_double), ("com_z", C.c_double), ("area", C.c_double), ("depth", C.c_double)]


class bw_action(C.Structure):
    _fields_ = [("target_block", C.c_int32), ("target_face", C.c_int32), ("shape", C.c_int32), ("face", C.c_int32),
                ("offset_x", C.c_double), ("offset_y", C.c_double), ("frozen", C.c_int32), ("reserved0", C.c_int32)]


class bw_block(C.Structure):
    _fields_ = [("x", C.c_double), ("z", C.c_double), ("c", C.c_double), ("s", C.c_double),
                ("shape", C.c_int32), ("is_static", C.c_int32)]


class bw_task(C.Structure):
    _fields_ = [("n_obstacles", C.c_int32), ("n_targets", C.c_int32), ("n_blocks", C.c_int32), ("reserved0", C.c_int32),
                ("obstacle_xz", (C.c_double * 2) * BW_MAX_OBSTACLES),
                ("target_xz", (C.c_double * 2) * BW_MAX_TARGETS),
                ("blocks", bw_block * BW_MAX_BLOCKS)]


class bw_step_out(C.Structure):
    _fields_ = [("residual", C.c_double), ("residual_unfrozen", C.c_double),
                ("distance_to_targets", C.c_double * BW_MAX_TARGETS),
                ("reward", C.c_float), ("lin_reward", C.c_float),
                ("n_blocks", C.c_int32), ("n_interfaces", C.c_int32), ("newton_iters", C.c_int32),
                ("solver_kflops", C.c_int32),
                ("stable", C.c_uint8), ("stable_unfrozen", C.c_uint8), ("collision", C.c_uint8),
                ("collision_block", C.c_uint8), ("collision_obstacle", C.c_uint8), ("collision_floor", C.c_uint8),
                ("collision_boundary", C.c_uint8), ("terminated", C.c_uint8), ("truncated", C.c_uint8),
                ("solver_status", C.c_uint8), ("error", C.c_uint8), ("n_targets_reached", C.c_uint8),
                ("lp_pivots", C.c_uint8), ("reserved1", C.c_uint8 * 3)]


class bw_obs_out(C.Structure):
    _fields_ = [("block_img_f32", C.c_void_p), ("block_img_u8", C.c_void_p), ("binary", C.c_void_p),
                ("block_bits", C.c_void_p)]


class bw_interface(C.Structure):
    _fields_ = [("body_a", C.c_int32), ("body_b", C.c_int32), ("face_a", C.c_int32), ("face_b", C.c_int32),
                ("nx", C.c_double), ("nz", C.c_double),
                ("p0x", C.c_double), ("p0z", C.c_double), ("p1x", C.c_double), ("p1z", C.c_double),
                ("fn0", C.c_double), ("ft0", C.c_double), ("fn1", C.c_double), ("ft1", C.c_double)]


class bw_transition(C.Structure):
    _fields_ = [("block_bits", C.c_uint64 * BW_IMG), ("action_bits", C.c_uint64 * BW_IMG),
                ("next_block_bits", C.c_uint64 * BW_IMG), ("action", bw_action),
                ("reward", C.c_float), ("lin_reward", C.c_float), ("env", C.c_int32), ("step", C.c_int32),
                ("n_next_candidates", C.c_int32),
                ("binary", C.c_uint8), ("next_binary", C.c_uint8), ("done", C.c_uint8), ("terminated", C.c_uint8),
                ("truncated", C.c_uint8), ("stable", C.c_uint8), ("stable_unfrozen", C.c_uint8), ("valid", C.c_uint8),
                ("reserved", C.c_uint8 * 4)]


class bw_rollout_view(C.Structure):
    _fields_ = [("cand", C.c_void_p), ("valid", C.c_void_p), ("n_cand", C.c_void_p), ("n_valid", C.c_void_p),
                ("action_bits", C.c_void_p), ("slot", C.c_void_p), ("amax", C.c_int32), ("reserved0", C.c_int32)]


# numpy views of the same layouts (for bulk transfers)
def np_dtypes():
    import numpy as np
    action = np.dtype([("target_block", "<i4"), ("target_face", "<i4"), ("shape", "<i4"), ("face", "<i4"),
                       ("offset_x", "<f8"), ("offset_y", "<f8"), ("frozen", "<i4"), ("reserved0", "<i4")])
    step_out = np.dtype([("residual", "<f8"), ("residual_unfrozen", "<f8"),
                         ("distance_to_targets", "<f8", (BW_MAX_TARGETS,)),
                         ("reward", "<f4"), ("lin_reward", "<f4"),
                         ("n_blocks", "<i4"), ("n_interfaces", "<i4"), ("newton_iters", "<i4"), ("solver_kflops", "<i4"),
                         ("stable", "u1"), ("stable_unfrozen", "u1"), ("collision", "u1"), ("collision_block", "u1"),
                         ("collision_obstacle", "u1"), ("collision_floor", "u1"), ("collision_boundary", "u1"),
                         ("terminated", "u1"), ("truncated", "u1"), ("solver_status", "u1"), ("error", "u1"),
                         ("n_targets_reached", "u1"), ("lp_pivots", "u1"), ("reserved1", "u1", (3,))])
    block = np.dtype([("x", "<f8"), ("z", "<f8"), ("c", "<f8"), ("s", "<f8"), ("shape", "<i4"), ("is_static", "<i4")])
    task = np.dtype([("n_obstacles", "<i4"), ("n_targets", "<i4"), ("n_blocks", "<i4"), ("reserved0", "<i4"),
                     ("obstacle_xz", "<f8", (BW_MAX_OBSTACLES, 2)), ("target_xz", "<f8", (BW_MAX_TARGETS, 2)),
                     ("blocks", block, (BW_MAX_BLOCKS,))])
    interface = np.dtype([("body_a", "<i4"), ("body_b", "<i4"), ("face_a", "<i4"), ("face_b", "<i4"),
                          ("nx", "<f8"), ("nz", "<f8"), ("p0x", "<f8"), ("p0z", "<f8"), ("p1x", "<f8"),
                          ("p1z", "<f8"), ("fn0", "<f8"), ("ft0", "<f8"), ("fn1", "<f8"), ("ft1", "<f8")])
    transition = np.dtype([("block_bits", "<u8", (BW_IMG,)), ("action_bits", "<u8", (BW_IMG,)),
                           ("next_block_bits", "<u8", (BW_IMG,)), ("action", action),
                           ("reward", "<f4"), ("lin_reward", "<f4"), ("env", "<i4"), ("step", "<i4"),
                           ("n_next_candidates", "<i4"),
                           ("binary", "u1"), ("next_binary", "u1"), ("done", "u1"), ("terminated", "u1"),
                           ("truncated", "u1"), ("stable", "u1"), ("stable_unfrozen", "u1"), ("valid", "u1"),
                           ("reserved", "u1", (4,))])
    assert transition.itemsize == C.sizeof(bw_transition) == 1608
    assert action.itemsize == C.sizeof(bw_action) and step_out.itemsize == C.sizeof(bw_step_out)
    assert block.itemsize == C.sizeof(bw_block) and task.itemsize == C.sizeof(bw_task)
    assert interface.itemsize == C.sizeof(bw_interface)
    return dict(action=action, step_out=step_out, block=block, task=task, interface=interface, transition=transition)


# every symbol include/bridges_b200.h declares: name -> (restype, argtypes)
_H = C.c_void_p
_P = C.c_void_p
SIGNATURES = {
    "bw_abi_version": (C.c_int, []),
    "bw_create": (C.c_int, [C.POINTER(bw_config), C.POINTER(_H)]),
    "bw_destroy": (None, [_H]),
    "bw_last_error": (C.c_char_p, [_H]),
    "bw_config_default": (None, [C.POINTER(bw_config)]),
    "bw_sync": (C.c_int, [_H]),
    "bw_load_shapes": (C.c_int, [_H, C.POINTER(bw_shape_desc), C.c_int32]),
    "bw_set_marker_shape": (C.c_int, [_H, C.POINTER(bw_shape_desc)]),
    "bw_set_task_kernel": (C.c_int, [_H, _P, C.c_int32]),
    "bw_set_mu": (C.c_int, [_H, _P]),
    "bw_reset": (C.c_int, [_H, _P, _P]),
    "bw_reset_host": (C.c_int, [_H, _P, _P]),
    "bw_reset_done": (C.c_int, [_H]),
    "bw_step": (C.c_int, [_H, _P, _P, _P, C.POINTER(bw_obs_out)]),
    "bw_evaluate": (C.c_int, [_H, _P, _P, C.POINTER(bw_obs_out)]),
    "bw_step_host": (C.c_int, [_H, _P, _P, _P, C.POINTER(bw_obs_out)]),
    "bw_set_host_transfer": (C.c_int, [_H, C.c_int32]),
    "bw_observe": (C.c_int, [_H, _P, _P, _P, _P]),
    "bw_observe_host": (C.c_int, [_H, _P, _P, _P, _P]),
    "bw_enumerate_actions": (C.c_int, [_H, _P, C.c_int32, _P, C.c_int32, C.c_int32, _P, _P, _P, _P]),
    "bw_enumerate_actions_stored": (C.c_int, [_H, _P, C.c_int32, _P, C.c_int32, C.c_int32, _P, _P, _P, _P]),
    "bw_gather_action_bits": (C.c_int, [_H, _P, C.c_int32, _P, _P, C.c_int64, _P]),
    "bw_candidate_overflow": (C.c_int, [_H, C.POINTER(C.c_int32)]),
    "bw_expand_bits": (C.c_int, [_H, _P, C.c_int64, _P]),
    "bw_select_random": (C.c_int, [_H, _P, _P, _P, C.c_int32, C.c_uint64, _P, _P]),
    "bw_rollout_configure": (C.c_int, [_H, _P, C.c_int32, _P, C.c_int32, C.c_int32, C.c_int32]),
    "bw_rollout_begin": (C.c_int, [_H, C.POINTER(bw_rollout_view)]),
    "bw_rollout_commit": (C.c_int, [_H, _P, _P, C.POINTER(bw_obs_out)]),
    "bw_rollout_gather_bits": (C.c_int, [_H, _P, _P, C.c_int64, _P]),
    "bw_rollout_random": (C.c_int, [_H, C.c_int32, C.c_uint64, _P, C.c_int64, C.c_int64]),
    "bw_unpack_transitions": (C.c_int, [_H, _P, _P, C.c_int64, _P, _P, _P, _P, _P, _P, _P, _P]),
    "bw_get_state": (C.c_int, [_H, _P, _P]),
    "bw_get_raster_bits": (C.c_int, [_H, _P, _P]),
    "bw_copy_raster_bits": (C.c_int, [_H, _P, _P]),
    "bw_contains_2d_host": (C.c_int, [_H, C.POINTER(bw_shape_desc), _P, _P, C.c_int64, _P]),
    "bw_get_forces": (C.c_int, [_H, C.c_int32, _P, _P]),
    "bw_get_target_state": (C.c_int, [_H, _P, _P, _P]),
    "bw_query_placement_host": (C.c_int, [_H, _P, _P, _P, _P, _P]),
    "bw_render_blocks_host": (C.c_int, [_H, _P, C.c_int32, _P, C.c_int32, _P, _P, _P]),
    "bw_set_static_mask": (C.c_int, [_H, _P]),
    "bw_set_timing": (C.c_int, [_H, C.c_int32]),
    "bw_last_step_kernel_ms": (C.c_int, [_H, _P]),
    "bw_fp64_peak_gflops": (C.c_int, [_H, _P]),
    "bw_debug_lp_stats": (C.c_int, [_H, _P]),
    "bw_kernel_launches": (C.c_int64, [_H]),
}

_lib = None


class BridgesError(RuntimeError):
    pass


def load():
    """dlopen the CUDA library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BridgesError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(bridges_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.bw_abi_version() != BW_ABI_VERSION:
        raise BridgesError("libbridges_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(lib, handle, rc):
    if rc != 0:
        msg = lib.bw_last_error(handle)
        raise BridgesError(f"bridges_b200 error {rc}: {msg.decode() if msg else ''}")
