"""ctypes binding of libtacex_b200.so (C ABI: include/tacex_b200.h).

The library is the product; there is NO CPU fallback: if the shared object is missing or no Blackwell GPU is
present, construction fails loudly.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

TX_MAX_BLURS = 8
TX_MAX_TAPS = 64
TX_MAX_MARKERS = 256
TX_ABI_VERSION = 2

import os

LIB_PATH = Path(os.environ.get("TACEX_B200_LIB", Path(__file__).resolve().parent / "lib" / "libtacex_b200.so"))


class TxConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int),
        ("H", C.c_int), ("W", C.c_int), ("max_envs", C.c_int), ("num_bins", C.c_int),
        ("pixmm", C.c_float), ("calib_h", C.c_float), ("calib_w", C.c_float), ("contact_scale", C.c_float),
        ("gelpad_height_m", C.c_float), ("gelpad_to_cam_min_m", C.c_float),
        ("n_blurs", C.c_int),
        ("ksx", C.c_int * TX_MAX_BLURS), ("ksy", C.c_int * TX_MAX_BLURS),
        ("taps_x", (C.c_float * TX_MAX_TAPS) * TX_MAX_BLURS), ("taps_y", (C.c_float * TX_MAX_TAPS) * TX_MAX_BLURS),
        ("marker_rows", C.c_int), ("marker_cols", C.c_int), ("marker_x0", C.c_float), ("marker_y0", C.c_float),
        ("fots_lambda", C.c_double * 3), ("mm2pix", C.c_double), ("shear_max_px", C.c_double),
        ("theta_max_rad", C.c_double),
    ]


class TxShadowConfig(C.Structure):
    _fields_ = [("D", C.c_int), ("Hn", C.c_int), ("S", C.c_int), ("F", C.c_int), ("depth_0", C.c_float),
                ("height_precision", C.c_float), ("discretize_precision", C.c_float), ("step_x", C.c_float), ("step_y", C.c_float),
                ("dil", C.c_int * 4), ("ks_sx", C.c_int), ("ks_sy", C.c_int),
                ("taps_sx", C.c_float * TX_MAX_TAPS), ("taps_sy", C.c_float * TX_MAX_TAPS)]


class TxCounters(C.Structure):
    _fields_ = [("render_calls", C.c_uint64), ("frames_rendered", C.c_uint64), ("fots_calls", C.c_uint64),
                ("depth_calls", C.c_uint64), ("kernels_launched", C.c_uint64)]


class TxFemConfig(C.Structure):
    _fields_ = [("V", C.c_int), ("T", C.c_int), ("A", C.c_int), ("S", C.c_int), ("dt", C.c_double),
                ("gravity", C.c_double * 3), ("mu", C.c_double), ("lam", C.c_double), ("density", C.c_double),
                ("attach_strength", C.c_double), ("d_hat", C.c_double), ("kappa", C.c_double),
                ("newton_max_iter", C.c_int), ("velocity_tol", C.c_double), ("pcg_tol_rate", C.c_double),
                ("pcg_max_iter_ratio", C.c_int), ("ls_max_iter", C.c_int), ("substep", C.c_int),
                ("rest_volume_det", C.c_int), ("friction_mu", C.c_double), ("eps_velocity", C.c_double)]


class TxFemIndenter(C.Structure):
    _fields_ = [("type", C.c_int), ("c", C.c_double * 3), ("R", C.c_double * 9), ("h", C.c_double * 3)]


class TxFemStats(C.Structure):
    _fields_ = [("converged", C.c_int), ("newton_iters", C.c_int), ("pcg_iters", C.c_int), ("ls_halvings", C.c_int),
                ("min_dist", C.c_double), ("last_res", C.c_double), ("energy", C.c_double)]


class TxError(RuntimeError):
    pass


_lib = None

# every symbol include/tacex_b200.h declares (tests check that the library exports all of them)
EXPORTS = [
    "tx_abi_version", "tx_create", "tx_destroy", "tx_last_error", "tx_get_counters", "tx_upload_tables",
    "tx_indentation_depth", "tx_indentation_depth_frames", "tx_render", "tx_render_depth", "tx_set_camera_resolution", "tx_render_camera", "tx_set_rect_output", "tx_set_multicast_output", "tx_obs_push", "tx_obs_fill", "tx_upload_shadow_tables", "tx_render_shadow", "tx_fots_markers", "tx_set_marker_patches", "tx_marker_overlay", "tx_resize", "tx_marker_grid", "tx_step_host", "tx_debug_set_ticks", "tx_debug_set_flags",
    "tx_fem_create", "tx_fem_destroy", "tx_fem_last_error", "tx_fem_get_mass", "tx_fem_step", "tx_fem_set_markers",
    "tx_fem_markers", "tx_fem_set_marker_output", "tx_fem_set_surface", "tx_fem_heightmap", "tx_fem_attachment_aim", "tx_fem_set_indenter_mesh", "tx_fem_set_contact_surface", "tx_fem_debug_set_cycles",
]


def load() -> C.CDLL:
    """dlopen the in-tree library and declare the prototypes. Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise TxError(f"{LIB_PATH} is missing: run `python -m tacex_b200.build` (nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    vp, fp, ip, u8p = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p  # device pointers travel as integers
    lib.tx_abi_version.restype = C.c_int
    lib.tx_create.argtypes = [C.POINTER(TxConfig), C.c_int, vp, C.POINTER(C.c_void_p)]
    lib.tx_create.restype = C.c_int
    lib.tx_destroy.argtypes = [C.c_void_p]
    lib.tx_destroy.restype = None
    lib.tx_last_error.argtypes = [C.c_void_p]
    lib.tx_last_error.restype = C.c_char_p
    lib.tx_get_counters.argtypes = [C.c_void_p, C.POINTER(TxCounters)]
    lib.tx_upload_tables.argtypes = [C.c_void_p, fp, fp, fp]
    lib.tx_indentation_depth.argtypes = [C.c_void_p, fp, C.c_int, fp]
    lib.tx_indentation_depth_frames.argtypes = [C.c_void_p, fp, C.c_int, C.c_int, fp]
    lib.tx_indentation_depth_frames.restype = C.c_int
    lib.tx_render.argtypes = [C.c_void_p, fp, fp, C.c_int, fp, fp, fp, u8p]
    lib.tx_render_depth.argtypes = [C.c_void_p, fp, C.c_float, C.c_int, fp, fp, fp]
    lib.tx_set_camera_resolution.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.tx_set_camera_resolution.restype = C.c_int
    lib.tx_render_camera.argtypes = [C.c_void_p, fp, C.c_int, C.c_float, fp, C.c_int, fp, fp, fp, u8p]
    lib.tx_render_camera.restype = C.c_int
    lib.tx_upload_shadow_tables.argtypes = [C.c_void_p, C.POINTER(TxShadowConfig), C.c_void_p, C.c_void_p, C.c_void_p]
    lib.tx_upload_shadow_tables.restype = C.c_int
    lib.tx_render_shadow.argtypes = [C.c_void_p, fp, fp, C.c_int, fp, fp]
    lib.tx_render_shadow.restype = C.c_int
    lib.tx_set_rect_output.argtypes = [C.c_void_p, vp]
    lib.tx_set_multicast_output.argtypes = [C.c_void_p, fp, ip]
    lib.tx_set_multicast_output.restype = C.c_int
    lib.tx_obs_push.argtypes = [C.c_void_p, fp, ip, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), fp, ip, vp]
    lib.tx_obs_fill.argtypes = [C.c_void_p, fp, ip, ip, C.c_int, C.c_int, C.c_int, vp]
    for name in ("tx_set_rect_output", "tx_obs_push", "tx_obs_fill"):
        getattr(lib, name).restype = C.c_int
    lib.tx_render_depth.restype = C.c_int
    lib.tx_fots_markers.argtypes = [C.c_void_p, fp, fp, C.c_int, fp, ip, fp]
    lib.tx_resize.argtypes = [C.c_void_p, fp, C.c_int, C.c_int, C.c_int, fp]
    lib.tx_resize.restype = C.c_int
    lib.tx_set_marker_patches.argtypes = [C.c_void_p, C.c_void_p]
    lib.tx_set_marker_patches.restype = C.c_int
    lib.tx_marker_overlay.argtypes = [C.c_void_p, fp, C.c_int, C.c_int, fp, C.c_int, fp, u8p, u8p]
    lib.tx_marker_overlay.restype = C.c_int
    lib.tx_marker_grid.argtypes = [C.c_void_p, ip, ip]
    lib.tx_step_host.argtypes = [C.c_void_p, fp, fp, C.c_int, fp, fp, fp]
    lib.tx_debug_set_ticks.argtypes = [C.c_void_p, C.c_void_p]
    lib.tx_debug_set_ticks.restype = C.c_int
    lib.tx_debug_set_flags.argtypes = [C.c_void_p, C.c_int]
    lib.tx_debug_set_flags.restype = C.c_int
    lib.tx_fem_create.argtypes = [C.POINTER(TxFemConfig), vp, vp, vp, vp, C.c_int, vp, C.POINTER(C.c_void_p)]
    lib.tx_fem_destroy.argtypes = [C.c_void_p]
    lib.tx_fem_destroy.restype = None
    lib.tx_fem_last_error.argtypes = [C.c_void_p]
    lib.tx_fem_last_error.restype = C.c_char_p
    lib.tx_fem_get_mass.argtypes = [C.c_void_p, vp]
    lib.tx_fem_step.argtypes = [C.c_void_p, vp, vp, vp, vp, vp, vp, C.c_int, vp]
    lib.tx_fem_set_markers.argtypes = [C.c_void_p, C.c_int, vp, vp, vp, vp, C.c_double, C.c_double, C.c_double, C.c_double]
    lib.tx_fem_markers.argtypes = [C.c_void_p, vp, C.c_int, vp]
    lib.tx_fem_attachment_aim.argtypes = [C.c_void_p, vp, vp, C.c_int, C.c_int, vp]
    lib.tx_fem_attachment_aim.restype = C.c_int
    lib.tx_fem_set_indenter_mesh.argtypes = [C.c_void_p, C.c_int, vp]
    lib.tx_fem_set_indenter_mesh.restype = C.c_int
    lib.tx_fem_set_contact_surface.argtypes = [C.c_void_p, C.c_int, vp]
    lib.tx_fem_set_contact_surface.restype = C.c_int
    lib.tx_fem_set_surface.argtypes = [C.c_void_p, C.c_int, vp]
    lib.tx_fem_set_surface.restype = C.c_int
    lib.tx_fem_heightmap.argtypes = [C.c_void_p, vp, C.c_int, vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_float]
    lib.tx_fem_heightmap.restype = C.c_int
    lib.tx_fem_set_marker_output.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int]
    lib.tx_fem_set_marker_output.restype = C.c_int
    lib.tx_fem_debug_set_cycles.argtypes = [C.c_void_p, vp]
    lib.tx_fem_debug_set_cycles.restype = C.c_int
    for name in ("tx_fem_create", "tx_fem_get_mass", "tx_fem_step", "tx_fem_set_markers", "tx_fem_markers"):
        getattr(lib, name).restype = C.c_int
    for name in ("tx_get_counters", "tx_upload_tables", "tx_indentation_depth", "tx_render", "tx_render_depth", "tx_set_camera_resolution", "tx_render_camera", "tx_set_rect_output", "tx_obs_push", "tx_obs_fill", "tx_fots_markers",
                 "tx_marker_grid", "tx_step_host"):
        getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib
