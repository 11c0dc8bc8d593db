"""ctypes binding of libb200sfm.so (include/b200sfm.h).  Fails loudly when the
CUDA library has not been built -- there is no CPU fallback."""
from __future__ import annotations

import ctypes as ct
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200SFM_LIB", os.path.join(_HERE, "libb200sfm.so"))   # env override: kernel-tuning sweeps only

INTR_STRIDE = 12
NCCL_ID_BYTES = 128

c_int32, c_int64, c_double, c_void_p = ct.c_int32, ct.c_int64, ct.c_double, ct.c_void_p
P = ct.POINTER


class LMStats(ct.Structure):
    _fields_ = [
        ("iterations", c_int32), ("num_successful_steps", c_int32), ("termination", c_int32), ("usable", c_int32),
        ("initial_cost", c_double), ("final_cost", c_double),
        ("num_observations", c_int64), ("pcg_iterations", c_int64), ("kernel_launches", c_int64),
        ("ms_total", c_double), ("ms_linearize", c_double), ("n_linearize", c_int64),
        ("ms_matvec", c_double), ("n_matvec", c_int64),
        ("ms_h2d", c_double), ("ms_d2h", c_double), ("h2d_bytes", c_int64), ("d2h_bytes", c_int64),
    ]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class BAOpts(ct.Structure):
    _fields_ = [
        ("optimize_rig_poses", c_int32), ("optimize_rotations", c_int32), ("optimize_translation", c_int32),
        ("optimize_intrinsics", c_int32), ("optimize_principal_point", c_int32), ("optimize_points", c_int32),
        ("min_num_view_per_track", c_int32), ("max_num_iterations", c_int32),
        ("thres_loss_function", c_double), ("function_tolerance", c_double), ("gradient_tolerance", c_double),
        ("parameter_tolerance", c_double),
        ("pcg_max_iterations", c_int32), ("pcg_min_iterations", c_int32), ("pcg_rel_tolerance", c_double),
        ("preconditioner", c_int32), ("profile_kernels", c_int32), ("fixed_num_iterations", c_int32),
        ("design", c_int32),
    ]


class GPOpts(ct.Structure):
    _fields_ = [
        ("optimize_positions", c_int32), ("optimize_points", c_int32), ("optimize_scales", c_int32),
        ("min_num_view_per_track", c_int32), ("max_num_iterations", c_int32),
        ("max_num_line_search_step_size_iterations", c_int32),
        ("thres_loss_function", c_double), ("function_tolerance", c_double), ("gradient_tolerance", c_double),
        ("parameter_tolerance", c_double),
        ("pcg_max_iterations", c_int32), ("pcg_min_iterations", c_int32), ("pcg_rel_tolerance", c_double),
        ("preconditioner", c_int32), ("profile_kernels", c_int32), ("fixed_num_iterations", c_int32),
        ("reserved0", c_int32),
    ]


class RAOpts(ct.Structure):
    _fields_ = [
        ("max_num_l1_iterations", c_int32), ("max_num_irls_iterations", c_int32), ("weight_type", c_int32),
        ("use_weight", c_int32), ("l1_step_convergence_threshold", c_double),
        ("irls_step_convergence_threshold", c_double), ("irls_loss_parameter_sigma", c_double),
        ("l1_max_admm_iterations", c_int32), ("reserved0", c_int32), ("l1_rho", c_double),
        ("l1_absolute_tolerance", c_double), ("l1_relative_tolerance", c_double),
        ("pcg_max_iterations", c_int32), ("reserved1", c_int32), ("pcg_rel_tolerance", c_double),
    ]


class RAStats(ct.Structure):
    _fields_ = [
        ("l1_iterations", c_int32), ("irls_iterations", c_int32), ("admm_iterations", c_int32), ("usable", c_int32),
        ("num_edges", c_int64), ("pcg_iterations", c_int64), ("kernel_launches", c_int64), ("ms_total", c_double),
    ]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


# name -> (restype, argtypes); every symbol include/b200sfm.h declares
PROTOTYPES = {
    "b200sfm_version": (c_int32, []),
    "b200sfm_create": (c_int32, [c_int32, P(c_void_p)]),
    "b200sfm_nccl_unique_id": (c_int32, [c_void_p]),
    "b200sfm_create_dist": (c_int32, [c_int32, c_int32, c_int32, c_void_p, P(c_void_p)]),
    "b200sfm_destroy": (None, [c_void_p]),
    "b200sfm_last_error": (ct.c_char_p, [c_void_p]),
    "b200sfm_rank": (c_int32, [c_void_p]),
    "b200sfm_world_size": (c_int32, [c_void_p]),
    "b200sfm_cuda_stream": (c_void_p, [c_void_p]),
    "b200sfm_kernel_launches": (c_int64, [c_void_p]),
    "b200sfm_ba_default_opts": (None, [P(BAOpts)]),
    "b200sfm_ba_solve": (c_int32, [c_void_p, P(BAOpts), c_int32, c_int32, c_int64, c_int32] + [c_void_p] * 10 + [P(LMStats)]),
    "b200sfm_ba_problem_create": (c_int32, [c_void_p, c_int32, c_int32, c_int64, c_int32] + [c_void_p] * 6 + [c_int32, P(c_void_p)]),
    "b200sfm_ba_problem_create_rig": (c_int32, [c_void_p, c_int32, c_int32, c_int64, c_int32, c_int32] + [c_void_p] * 9 + [c_int32, P(c_void_p)]),
    "b200sfm_ba_problem_set_sensor_variable": (c_int32, [c_void_p, c_void_p]),
    "b200sfm_ba_problem_get_sensor_poses": (c_int32, [c_void_p, c_void_p, c_void_p]),
    "b200sfm_ba_problem_set_state": (c_int32, [c_void_p] * 5),
    "b200sfm_ba_problem_get_state": (c_int32, [c_void_p] * 5),
    "b200sfm_ba_problem_save_state": (c_int32, [c_void_p]),
    "b200sfm_ba_problem_restore_state": (c_int32, [c_void_p]),
    "b200sfm_ba_problem_solve": (c_int32, [c_void_p, P(BAOpts), P(LMStats)]),
    "b200sfm_ba_problem_cost": (c_int32, [c_void_p, P(BAOpts), P(c_double)]),
    "b200sfm_ba_problem_free": (None, [c_void_p]),
    "b200sfm_ba_problem_filter_reprojection": (c_int32, [c_void_p, c_double, c_void_p, P(c_int64)]),
    "b200sfm_ba_problem_filter_reprojection_normalized": (c_int32, [c_void_p, c_void_p, c_double, c_void_p, P(c_int64)]),
    "b200sfm_ba_problem_filter_angle": (c_int32, [c_void_p, c_void_p, c_void_p, c_double, c_void_p, P(c_int64)]),
    "b200sfm_ba_problem_filter_triangulation_angle": (c_int32, [c_void_p, c_double, c_void_p, P(c_int64)]),
    "b200sfm_ba_problem_normalize": (c_int32, [c_void_p, c_int32, c_double, c_double, c_double, P(c_double), c_void_p]),
    "b200sfm_ba_problem_undistort": (c_int32, [c_void_p, c_void_p]),
    "b200sfm_tracks_establish": (c_int32, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_double, P(c_void_p),
                                           P(c_int64), P(c_int64), P(c_int64)]),
    "b200sfm_tracks_get": (c_int32, [c_void_p] * 5),
    "b200sfm_tracks_free": (None, [c_void_p]),
    "b200sfm_gp_default_opts": (None, [P(GPOpts)]),
    "b200sfm_gp_solve": (c_int32, [c_void_p, P(GPOpts), c_int32, c_int32, c_int64] + [c_void_p] * 8 + [P(LMStats)]),
    "b200sfm_gp_problem_create": (c_int32, [c_void_p, c_int32, c_int32, c_int64] + [c_void_p] * 5 + [c_int32, P(c_void_p)]),
    "b200sfm_gp_problem_set_rig_terms": (c_int32, [c_void_p, c_void_p, c_void_p]),
    "b200sfm_gp_problem_set_rig_unknown": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p]),
    "b200sfm_gp_problem_get_rig_unknown": (c_int32, [c_void_p, c_void_p]),
    "b200sfm_gp_problem_set_state": (c_int32, [c_void_p] * 4),
    "b200sfm_gp_problem_get_state": (c_int32, [c_void_p] * 4),
    "b200sfm_gp_problem_save_state": (c_int32, [c_void_p]),
    "b200sfm_gp_problem_restore_state": (c_int32, [c_void_p]),
    "b200sfm_gp_problem_solve": (c_int32, [c_void_p, P(GPOpts), P(LMStats)]),
    "b200sfm_gp_problem_free": (None, [c_void_p]),
    "b200sfm_ra_default_opts": (None, [P(RAOpts)]),
    "b200sfm_ra_solve": (c_int32, [c_void_p, P(RAOpts), c_int32, c_int64] + [c_void_p] * 4 + [c_int32, c_void_p, P(RAStats)]),
    "b200sfm_ra_solve_rig": (c_int32, [c_void_p, P(RAOpts), c_int32, c_int32, c_int64] + [c_void_p] * 8 + [c_int32, c_void_p, P(RAStats)]),
    "b200sfm_ra_solve_gravity": (c_int32, [c_void_p, P(RAOpts), c_int32, c_int64] + [c_void_p] * 5 + [c_int32, c_void_p, P(RAStats)]),
}

_lib = None


def load() -> ct.CDLL:
    """Load libb200sfm.so and bind every prototype.  Raises if missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(glomap_b200 has no CPU fallback)")
    lib = ct.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class B200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"b200sfm status {code}: {msg}")
        self.code = code


def check(ctx, rc: int):
    if rc != 0:
        msg = load().b200sfm_last_error(ctx).decode() if ctx else ""
        raise B200Error(rc, msg)
