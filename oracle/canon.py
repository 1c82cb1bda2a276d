"""ctypes front-end of the canonical CPU restatement (oracle/taxim_canon.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this module.
"""

from __future__ import annotations

import ctypes as C
import math
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = _HERE / "_build" / "libtaxim_canon.so"
_MAX_BLURS = 8


def build(force: bool = False) -> Path:
    """Compile the C restatement (gcc, no GPU needed)."""
    src = _HERE / "taxim_canon.c"
    if force or not _LIB.exists() or _LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "_build/libtaxim_canon.so"], check=True, capture_output=True)
    return _LIB


class _Cfg(C.Structure):
    _fields_ = [
        ("H", C.c_int), ("W", C.c_int), ("num_bins", C.c_int), ("pixmm", C.c_float),
        ("calib_h", C.c_float), ("calib_w", C.c_float), ("contact_scale", C.c_float), ("n_blurs", C.c_int),
        ("ksx", C.c_int * _MAX_BLURS), ("ksy", C.c_int * _MAX_BLURS),
        ("off_x", C.c_int * _MAX_BLURS), ("off_y", C.c_int * _MAX_BLURS),
    ]


class _FotsCfg(C.Structure):
    _fields_ = [
        ("H", C.c_int), ("W", C.c_int), ("rows", C.c_int), ("cols", C.c_int),
        ("lamb", C.c_double * 3), ("mm2pix", C.c_double), ("shear_max_px", C.c_double), ("theta_max_rad", C.c_double),
    ]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB))
        _lib.canon_taxim_render.restype = C.c_int
        _lib.canon_num_threads.restype = C.c_int
        _lib.canon_resize_bilinear.restype = C.c_int
    return _lib


def _p(a: np.ndarray | None, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class CanonTaxim:
    """Canonical Taxim optical path. ``taps`` = [(x taps, y taps)] for the pyramid levels + the final blur."""

    def __init__(self, H, W, poly_grad, background, gel_map, taps, pixmm=0.0295, num_bins=125, calib_hw=(480, 640),
                 contact_scale=0.4, gelpad_height_m=0.0045, gelpad_to_cam_min_m=0.024):
        self.H, self.W = int(H), int(W)
        self.poly = np.ascontiguousarray(poly_grad, np.float32)
        self.bg = np.ascontiguousarray(background, np.float32)
        self.gel = None if gel_map is None else np.ascontiguousarray(gel_map, np.float32)
        assert self.poly.shape == (3, num_bins, num_bins, 6) and self.bg.shape == (3, H, W)
        cfg = _Cfg()
        cfg.H, cfg.W, cfg.num_bins, cfg.pixmm = self.H, self.W, num_bins, pixmm
        cfg.calib_h, cfg.calib_w, cfg.contact_scale = calib_hw[0], calib_hw[1], contact_scale
        cfg.n_blurs = len(taps)
        flat, off = [], 0
        for l, (kx, ky) in enumerate(taps):
            kx = np.asarray(kx, np.float32)
            ky = np.asarray(ky, np.float32)
            cfg.ksx[l], cfg.ksy[l] = kx.size, ky.size
            cfg.off_x[l] = off
            off += kx.size
            cfg.off_y[l] = off
            off += ky.size
            flat += [kx, ky]
        self.taps = np.ascontiguousarray(np.concatenate(flat), np.float32)
        self.cfg = cfg
        self.gelpad_height_m, self.gelpad_to_cam_min_m = gelpad_height_m, gelpad_to_cam_min_m

    def indentation_depth(self, hm_mm: np.ndarray) -> np.ndarray:
        """Any frame shape [N][h][w]: the reference computes it on the camera-resolution height map (taxim_sim.py:115-131)."""
        hm = np.ascontiguousarray(hm_mm, np.float32)
        out = np.empty(hm.shape[0], np.float32)
        lib().canon_indentation_depth(_p(hm, C.c_float), hm.shape[0], hm.shape[1], hm.shape[2], C.c_float(self.gelpad_height_m),
                                      C.c_float(self.gelpad_to_cam_min_m), _p(out, C.c_float))
        return out

    def render(self, hm_mm: np.ndarray, press_mm: np.ndarray, want=("deformed", "mask", "idx", "rgb")) -> dict:
        hm = np.ascontiguousarray(hm_mm, np.float32)
        N = hm.shape[0]
        pr = np.ascontiguousarray(press_mm, np.float32)
        out = {}
        dg = np.empty((N, self.H, self.W), np.float32) if "deformed" in want else None
        mk = np.empty((N, self.H, self.W), np.uint8) if "mask" in want else None
        im = np.empty((N, self.H, self.W), np.int32) if "idx" in want else None
        idr = np.empty((N, self.H, self.W), np.int32) if "idx" in want else None
        rgb = np.empty((N, self.H, self.W, 3), np.float32) if "rgb" in want else None
        rc = lib().canon_taxim_render(C.byref(self.cfg), _p(self.taps, C.c_float), _p(self.poly, C.c_float),
                                      _p(self.bg, C.c_float), _p(self.gel, C.c_float), _p(hm, C.c_float),
                                      _p(pr, C.c_float), N, _p(dg, C.c_float), _p(mk, C.c_uint8), _p(im, C.c_int32),
                                      _p(idr, C.c_int32), _p(rgb, C.c_float))
        if rc != 0:
            raise MemoryError("canon_taxim_render failed")
        out.update(deformed=dg, mask=mk, idx_mag=im, idx_dir=idr, rgb=rgb)
        return out


class _ShadowCfg(C.Structure):
    _fields_ = [("D", C.c_int), ("Hn", C.c_int), ("S", C.c_int), ("F", C.c_int), ("depth_0", C.c_float),
                ("height_precision", C.c_float), ("discretize_precision", C.c_float), ("step_x", C.c_float), ("step_y", C.c_float),
                ("dil", (C.c_int * 2) * 2), ("ks_sx", C.c_int), ("ks_sy", C.c_int)]


def render_shadow(ct: "CanonTaxim", st, hm_mm: np.ndarray, press_mm: np.ndarray, want_boundary: bool = False,
                  want_shadow_img: bool = False):
    """Canonical Taxim render WITH shadows (ref: taxim_torch.py:260-346). ``st`` = tacex_b200.calib.ShadowTables."""
    hm = np.ascontiguousarray(hm_mm, np.float32)
    pr = np.ascontiguousarray(press_mm, np.float32)
    N = hm.shape[0]
    sc = _ShadowCfg()
    sc.D, sc.Hn, sc.S = st.table.shape[1], st.table.shape[2], st.table.shape[3]
    sc.F = st.fan_cos.shape[1]
    sc.depth_0, sc.height_precision, sc.discretize_precision = st.depth_0, st.height_precision, st.discretize_precision
    sc.step_x, sc.step_y = st.step_x, st.step_y
    for r in range(2):
        sc.dil[r][0], sc.dil[r][1] = st.dilate_rounds[r]
    tx = np.ascontiguousarray(st.blur_taps[0], np.float32)
    ty = np.ascontiguousarray(st.blur_taps[1], np.float32)
    sc.ks_sx, sc.ks_sy = tx.size, ty.size
    tab = np.ascontiguousarray(st.table, np.float32)
    fc_, fs_ = np.ascontiguousarray(st.fan_cos, np.float32), np.ascontiguousarray(st.fan_sin, np.float32)
    rgb = np.empty((N, ct.H, ct.W, 3), np.float32)
    bnd = np.empty((N, ct.H, ct.W), np.uint8) if want_boundary else None
    shi = np.empty((N, 3, ct.H, ct.W), np.float32) if want_shadow_img else None
    lib().canon_taxim_render_shadow.restype = C.c_int
    rc = lib().canon_taxim_render_shadow(C.byref(ct.cfg), _p(ct.taps, C.c_float), _p(ct.poly, C.c_float), _p(ct.bg, C.c_float),
                                         _p(ct.gel, C.c_float), _p(hm, C.c_float), _p(pr, C.c_float), N, C.byref(sc),
                                         _p(tab, C.c_float), _p(fc_, C.c_float), _p(fs_, C.c_float), _p(tx, C.c_float),
                                         _p(ty, C.c_float), _p(rgb, C.c_float), _p(bnd, C.c_uint8), _p(shi, C.c_float))
    if rc != 0:
        raise MemoryError("canon_taxim_render_shadow failed")
    if want_boundary or want_shadow_img:
        return {"rgb": rgb, "boundary": bnd, "shadow_img": shi}
    return rgb


class CanonFots:
    """Canonical FOTS marker motion with the per-env trajectory state the reference keeps in python lists."""

    def __init__(self, H=240, W=320, rows=9, cols=11, x0=15, y0=26, lamb=(0.00125, 0.00021, 0.00038), mm2pix=19.58):
        c = _FotsCfg()
        c.H, c.W, c.rows, c.cols = H, W, rows, cols
        c.lamb[:] = lamb
        c.mm2pix, c.shear_max_px, c.theta_max_rad = mm2pix, 10.0, 60.0 / 180.0 * math.pi
        self.cfg, self.M = c, rows * cols
        self.mx = np.empty(self.M, np.int32)
        self.my = np.empty(self.M, np.int32)
        lib().canon_fots_grid(C.byref(c), C.c_double(x0), C.c_double(y0), _p(self.mx, C.c_int32), _p(self.my, C.c_int32))
        self.traj0 = None
        self.traj_len = None

    def reset(self, N):
        self.traj0 = np.zeros((N, 4), np.float32)
        self.traj_len = np.zeros(N, np.int32)

    def step(self, deformed, mask, press, theta) -> np.ndarray:
        N = deformed.shape[0]
        if self.traj0 is None or self.traj0.shape[0] != N:
            self.reset(N)
        dg = np.ascontiguousarray(deformed, np.float32)
        mk = np.ascontiguousarray(mask, np.uint8)
        pr = np.ascontiguousarray(press, np.float32)
        th = np.ascontiguousarray(theta, np.float32)
        out = np.empty((N, 2, self.M, 2), np.float32)
        lib().canon_fots_step(C.byref(self.cfg), _p(self.mx, C.c_int32), _p(self.my, C.c_int32), _p(dg, C.c_float),
                              _p(mk, C.c_uint8), _p(pr, C.c_float), _p(th, C.c_float), N, _p(self.traj0, C.c_float),
                              _p(self.traj_len, C.c_int32), _p(out, C.c_float))
        return out


def resize_bilinear(src: np.ndarray, out_hw: tuple[int, int]) -> np.ndarray:
    """Canonical restatement of the reference's F.resize of the height map (taxim_sim.py:88-89): [N][Hi][Wi] -> [N][Ho][Wo]."""
    s = np.ascontiguousarray(src, np.float32)
    N, Hi, Wi = s.shape
    out = np.empty((N, out_hw[0], out_hw[1]), np.float32)
    rc = lib().canon_resize_bilinear(_p(s, C.c_float), N, Hi, Wi, out_hw[0], out_hw[1], _p(out, C.c_float))
    if rc != 0:
        raise ValueError("canon_resize_bilinear: unsupported scale")
    return out


def num_threads() -> int:
    return int(lib().canon_num_threads())


def use_all_threads() -> int:
    """OpenMP thread count = all host CPUs (torchrun sets OMP_NUM_THREADS=1 in the workers)."""
    import os

    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().canon_set_threads(n)
    try:
        from . import fem_canon

        fem_canon.lib().fem_set_threads(n)
    except Exception:
        pass
    return num_threads()


# ---- marker image / marker overlay (ref: fots_marker_sim.py:346-384, ball_rolling_taxim_fots.py:918-937) ------------------
def marker_image(markers_xy: np.ndarray, patch_w: np.ndarray, H: int = 240, W: int = 320) -> np.ndarray:
    """Restatement of ``FOTSMarkerSimulator.draw_markers``: (M, 2) float32 marker positions (x, y) -> (H, W) uint8. ``patch_w`` is
    ``generate_patch_array()['patch_array'][:, :, w]`` (10, 10, 12, 12) for the marker size in use."""
    import math

    canvas = np.full((H + 24, W + 24), 255, np.uint8)
    uv = np.asarray(markers_xy, np.float32).astype(np.float64) + 0.5
    for k in range(uv.shape[0]):
        u, v = uv[k, 0] + 12, uv[k, 1] + 12
        pu = math.floor((u - math.floor(u)) * 10)
        pv = math.floor((v - math.floor(v)) * 10)
        cu, cv = math.floor(u) - 6, math.floor(v) - 6
        if canvas.shape[1] - 12 > cu >= 0 and canvas.shape[0] - 12 > cv >= 0:
            canvas[cv:cv + 12, cu:cu + 12] = patch_w[pu, pv]
    return canvas[12:-12, 12:-12]


def marker_overlay(rgb: np.ndarray, marker_img: np.ndarray) -> np.ndarray:
    """The RL task's overlay arithmetic in float32: ((rgb * 255) * (marker / 255)) / 255 per channel."""
    r = np.asarray(rgb, np.float32)
    w = (marker_img.astype(np.float32) / np.float32(255.0))[..., None]
    return ((r * np.float32(255.0)) * w) / np.float32(255.0)


def rgb_to_u8(rgb: np.ndarray) -> np.ndarray:
    """round-half-even(rgb * 255) clamped to [0, 255] (the extension output of tx_marker_overlay)."""
    return np.clip(np.rint(np.asarray(rgb, np.float32) * np.float32(255.0)), 0, 255).astype(np.uint8)
