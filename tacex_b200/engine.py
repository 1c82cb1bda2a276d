"""TactileEngine -- thin PyTorch-facing owner of one ``tx_handle`` (device memory + stream plumbing only).

All arithmetic of the hot path happens inside libtacex_b200.so (hand-written sm_100a kernels); torch supplies the
device tensors and the CUDA stream. There is no fallback: without the library or without a CUDA device the
constructor raises.
"""

from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib
from .calib import TaximTables


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


class TactileEngine:
    def __init__(
        self,
        tables: TaximTables,
        max_envs: int,
        device: str | torch.device = "cuda",
        marker_rows: int = 0,
        marker_cols: int = 0,
        marker_x0: float = 15.0,
        marker_y0: float = 26.0,
        fots_lambda: tuple[float, float, float] = (0.00125, 0.00021, 0.00038),
        mm2pix: float = 19.58,
        gelpad_height_m: float = 0.0045,
        gelpad_to_cam_min_m: float = 0.024,
        use_current_stream: bool = True,
    ):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.TxError("TactileEngine needs a CUDA (sm_100a) device; there is no CPU path")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.TxError(f"TactileEngine needs a CUDA device, got {self.device}")
        self.dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", self.dev_index)
        self.tables = tables
        H, W = tables.shape
        self.H, self.W, self.max_envs = H, W, int(max_envs)
        self.M = marker_rows * marker_cols
        p = tables.params

        cfg = _lib.TxConfig()
        cfg.abi_version = _lib.TX_ABI_VERSION
        cfg.H, cfg.W, cfg.max_envs, cfg.num_bins = H, W, self.max_envs, p.num_bins
        cfg.pixmm, cfg.calib_h, cfg.calib_w, cfg.contact_scale = p.pixmm, p.calib_h, p.calib_w, p.contact_scale
        cfg.gelpad_height_m, cfg.gelpad_to_cam_min_m = gelpad_height_m, gelpad_to_cam_min_m
        taps = p.blur_taps((H, W))
        if len(taps) > _lib.TX_MAX_BLURS:
            raise _lib.TxError("too many blur levels")
        cfg.n_blurs = len(taps)
        for l, (kx, ky) in enumerate(taps):
            if kx.numel() > _lib.TX_MAX_TAPS or ky.numel() > _lib.TX_MAX_TAPS:
                raise _lib.TxError("blur kernel too large")
            cfg.ksx[l], cfg.ksy[l] = kx.numel(), ky.numel()
            for k, v in enumerate(kx.tolist()):
                cfg.taps_x[l][k] = v
            for k, v in enumerate(ky.tolist()):
                cfg.taps_y[l][k] = v
        # the 2-CTA kernel is specialised for 240 x 320 / kernel sizes 61,33,17,9,5,3,5; any other shape (the reference's RL tasks
        # render 32 x 24 / 32 x 32 tactile images) runs the arbitrary-resolution kernel of the library: no FOTS aux, no fused resize
        self.generic = (H, W) != (240, 320) or [(kx.numel(), ky.numel()) for kx, ky in taps] != [(k, k) for k in (61, 33, 17, 9, 5, 3, 5)]
        cfg.marker_rows, cfg.marker_cols = marker_rows, marker_cols
        cfg.marker_x0, cfg.marker_y0 = marker_x0, marker_y0
        cfg.fots_lambda[:] = fots_lambda
        cfg.mm2pix, cfg.shear_max_px, cfg.theta_max_rad = mm2pix, 10.0, 60.0 / 180.0 * math.pi
        self.cfg = cfg

        with torch.cuda.device(self.dev_index):
            self.stream = torch.cuda.current_stream() if use_current_stream else torch.cuda.Stream()
            h = C.c_void_p()
            rc = self.lib.tx_create(C.byref(cfg), self.dev_index, C.c_void_p(self.stream.cuda_stream), C.byref(h))
            if rc != 0:
                raise _lib.TxError(f"tx_create failed ({rc}): {self.lib.tx_last_error(None).decode()}")
            self.h = h
            poly = tables.poly_grad.contiguous().float()
            bg = tables.background.contiguous().float()
            gel = tables.gel_map.contiguous().float() if tables.gel_map is not None else None
            self._check(self.lib.tx_upload_tables(self.h, _ptr(poly), _ptr(bg), _ptr(gel)))
        if self.M:
            import numpy as np

            mx = np.empty(self.M, np.int32)
            my = np.empty(self.M, np.int32)
            self._check(self.lib.tx_marker_grid(self.h, mx.ctypes.data, my.ctypes.data))
            self.marker_x, self.marker_y = mx, my

    # -- plumbing ------------------------------------------------------------------------------------------------
    def _check(self, rc: int) -> None:
        if rc != 0:
            raise _lib.TxError(f"libtacex_b200 error {rc}: {self.lib.tx_last_error(self.h).decode()}")

    def close(self) -> None:
        if getattr(self, "h", None):
            self.lib.tx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def counters(self) -> dict:
        c = _lib.TxCounters()
        self._check(self.lib.tx_get_counters(self.h, C.byref(c)))
        return {k: int(getattr(c, k)) for k, _ in c._fields_}

    def _chk_hm(self, hm: torch.Tensor) -> int:
        if hm.device != self.device or hm.dtype != torch.float32 or not hm.is_contiguous():
            raise _lib.TxError("height map must be a contiguous float32 tensor on the engine's device")
        if hm.dim() != 3 or tuple(hm.shape[1:]) != (self.H, self.W):
            raise _lib.TxError(f"height map must have shape (N, {self.H}, {self.W})")
        return hm.shape[0]

    # -- hot path --------------------------------------------------------------------------------------------------
    def indentation_depth(self, hm: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        N = self._chk_hm(hm)
        out = torch.empty(N, device=self.device) if out is None else out
        self._check(self.lib.tx_indentation_depth(self.h, _ptr(hm), N, _ptr(out)))
        return out

    def indentation_depth_frames(self, frames: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """Indentation depth [mm] of height maps at ANY resolution (N, Hc, Wc): what the reference computes from the camera map."""
        if frames.device != self.device or frames.dtype != torch.float32 or not frames.is_contiguous() or frames.dim() != 3:
            raise _lib.TxError("frames must be a contiguous float32 (N, Hc, Wc) tensor on the engine's device")
        N = frames.shape[0]
        out = torch.empty(N, device=self.device) if out is None else out
        self._check(self.lib.tx_indentation_depth_frames(self.h, _ptr(frames), N, int(frames.shape[1] * frames.shape[2]), _ptr(out)))
        return out

    def render(
        self,
        hm: torch.Tensor,
        press: torch.Tensor | None = None,
        out: torch.Tensor | None = None,
        depth_out: torch.Tensor | None = None,
        deformed_out: torch.Tensor | None = None,
        mask_out: torch.Tensor | None = None,
    ) -> torch.Tensor:
        """RGB (N, H, W, 3) float32. ``press`` None = fused indentation depth (stored into ``depth_out``)."""
        N = self._chk_hm(hm)
        out = torch.empty((N, self.H, self.W, 3), device=self.device) if out is None else out
        self._check(self.lib.tx_render(self.h, _ptr(hm), _ptr(press), N, _ptr(out), _ptr(depth_out), _ptr(deformed_out),
                                       _ptr(mask_out)))
        return out

    def render_depth(self, depth_m: torch.Tensor, clip_max_m: float, out: torch.Tensor | None = None,
                     depth_out: torch.Tensor | None = None, height_map_out: torch.Tensor | None = None) -> torch.Tensor:
        """Fused ``_get_height_map`` + ``compute_indentation_depth`` + ``optical_simulation`` from the raw camera depth [m]."""
        N = self._chk_hm(depth_m)
        out = torch.empty((N, self.H, self.W, 3), device=self.device) if out is None else out
        self._check(self.lib.tx_render_depth(self.h, _ptr(depth_m), float(clip_max_m), N, _ptr(out), _ptr(depth_out),
                                             _ptr(height_map_out)))
        return out

    def set_camera_resolution(self, Hc: int, Wc: int) -> None:
        """Sensor camera coarser than the tactile image (ref: taxim_sim.py:88-89): precompute the bilinear resize taps."""
        self._check(self.lib.tx_set_camera_resolution(self.h, int(Hc), int(Wc)))
        self.cam_hw = (int(Hc), int(Wc))

    def render_camera(self, frames: torch.Tensor, is_depth: bool = False, clip_max_m: float = 0.0,
                      press: torch.Tensor | None = None, out: torch.Tensor | None = None, depth_out: torch.Tensor | None = None,
                      deformed_out: torch.Tensor | None = None, mask_out: torch.Tensor | None = None) -> torch.Tensor:
        """RGB (N, H, W, 3) from camera-resolution frames (N, Hc, Wc): F.resize (bilinear) fused into the load stage."""
        if frames.device != self.device or frames.dtype != torch.float32 or not frames.is_contiguous():
            raise _lib.TxError("frames must be a contiguous float32 tensor on the engine's device")
        if frames.dim() != 3 or tuple(frames.shape[1:]) != getattr(self, "cam_hw", None):
            raise _lib.TxError("frames must have the shape given to set_camera_resolution")
        N = frames.shape[0]
        out = torch.empty((N, self.H, self.W, 3), device=self.device) if out is None else out
        self._check(self.lib.tx_render_camera(self.h, _ptr(frames), int(bool(is_depth)), float(clip_max_m), _ptr(press), N,
                                              _ptr(out), _ptr(depth_out), _ptr(deformed_out), _ptr(mask_out)))
        return out

    # -- shadow branch (with_shadow=True) ------------------------------------------------------------------
    def upload_shadow_tables(self, st) -> None:
        """``st``: :class:`tacex_b200.calib.ShadowTables` (host, init time)."""
        import numpy as np

        c = _lib.TxShadowConfig()
        c.D, c.Hn, c.S = (int(v) for v in st.table.shape[1:])
        c.F = int(st.fan_cos.shape[1])
        c.depth_0, c.height_precision, c.discretize_precision = st.depth_0, st.height_precision, st.discretize_precision
        c.step_x, c.step_y = st.step_x, st.step_y
        c.dil[:] = [int(st.dilate_rounds[0][0]), int(st.dilate_rounds[0][1]), int(st.dilate_rounds[1][0]), int(st.dilate_rounds[1][1])]
        tx, ty = np.asarray(st.blur_taps[0], np.float32), np.asarray(st.blur_taps[1], np.float32)
        if tx.size > _lib.TX_MAX_TAPS or ty.size > _lib.TX_MAX_TAPS:
            raise _lib.TxError("shadow blur kernel too large")
        c.ks_sx, c.ks_sy = int(tx.size), int(ty.size)
        for k, v in enumerate(tx.tolist()):
            c.taps_sx[k] = v
        for k, v in enumerate(ty.tolist()):
            c.taps_sy[k] = v
        tab = np.ascontiguousarray(st.table, np.float32)
        fc, fs = np.ascontiguousarray(st.fan_cos, np.float32), np.ascontiguousarray(st.fan_sin, np.float32)
        self._check(self.lib.tx_upload_shadow_tables(self.h, C.byref(c), tab.ctypes.data, fc.ctypes.data, fs.ctypes.data))

    def render_shadow(self, hm: torch.Tensor, press: torch.Tensor | None, out: torch.Tensor | None = None,
                      depth_out: torch.Tensor | None = None) -> torch.Tensor:
        """RGB (N, H, W, 3) with shadows (``render_direct(with_shadow=True)``); arguments as :meth:`render`."""
        N = self._chk_hm(hm)
        out = torch.empty((N, self.H, self.W, 3), device=self.device) if out is None else out
        self._check(self.lib.tx_render_shadow(self.h, _ptr(hm), _ptr(press), N, _ptr(out), _ptr(depth_out)))
        return out

    # -- multi-GPU observation gather ----------------------------------------------------------------------
    def set_rect_output(self, rect: torch.Tensor | None) -> None:
        """int32 (N, 2, 4) device tensor that the following renders fill with the per-half non-flat rectangle, or None."""
        self._check(self.lib.tx_set_rect_output(self.h, _ptr(rect)))

    def set_multicast_output(self, mc_rgb: int = 0, mc_rect: int = 0) -> None:
        """Multicast addresses (ints; 0, 0 = off) of this rank's block of the gathered RGB / rectangle buffers: the following renders
        store their rectangles through them (fused all-gather, needs ``set_rect_output``)."""
        self._check(self.lib.tx_set_multicast_output(self.h, mc_rgb or None, mc_rect or None))

    def obs_push(self, rgb_local: torch.Tensor, rect_local: torch.Tensor, peer_rgb: list, peer_rect: list, stream,
                 mc_rgb: int = 0, mc_rect: int = 0) -> None:
        import ctypes as C

        n = len(peer_rgb)
        pr = (C.c_void_p * max(n, 1))(*[t.data_ptr() for t in peer_rgb])
        pd = (C.c_void_p * max(n, 1))(*[t.data_ptr() for t in peer_rect])
        self._check(self.lib.tx_obs_push(self.h, _ptr(rgb_local), _ptr(rect_local), rgb_local.shape[0], n, pr, pd,
                                         mc_rgb or None, mc_rect or None, C.c_void_p(stream.cuda_stream)))

    def obs_fill(self, rgb_all: torch.Tensor, rect_all: torch.Tensor, prev_rect: torch.Tensor | None, skip_lo: int, skip_hi: int,
                 stream) -> None:
        import ctypes as C

        self._check(self.lib.tx_obs_fill(self.h, _ptr(rgb_all), _ptr(rect_all), _ptr(prev_rect), rgb_all.shape[0], int(skip_lo),
                                         int(skip_hi), C.c_void_p(stream.cuda_stream)))

    def fots_markers(self, press: torch.Tensor, theta: torch.Tensor, traj0: torch.Tensor, traj_len: torch.Tensor,
                     out: torch.Tensor | None = None) -> torch.Tensor:
        N = press.shape[0]
        out = torch.empty((N, 2, self.M, 2), device=self.device) if out is None else out
        self._check(self.lib.tx_fots_markers(self.h, _ptr(press), _ptr(theta), N, _ptr(traj0), _ptr(traj_len), _ptr(out)))
        return out

    def resize(self, frames: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """Antialiased bilinear resize (N, Hi, Wi) -> (N, H, W): torchvision's ``F.resize`` of the reference for a camera finer than
        the tactile image, bit-identical to ``torch.nn.functional.interpolate(mode="bilinear", antialias=True)``."""
        if frames.device != self.device or frames.dtype != torch.float32 or not frames.is_contiguous() or frames.dim() != 3:
            raise _lib.TxError("frames must be a contiguous float32 (N, Hi, Wi) tensor on the engine's device")
        N = frames.shape[0]
        out = torch.empty((N, self.H, self.W), device=self.device) if out is None else out
        self._check(self.lib.tx_resize(self.h, _ptr(frames), N, int(frames.shape[1]), int(frames.shape[2]), _ptr(out)))
        return out

    # -- marker image / marker overlay (ref: fots_marker_sim.py:346-384, ball_rolling_taxim_fots.py:918-937) --------------
    def set_marker_patches(self, patches) -> None:
        """``patches``: uint8 (10, 10, 12, 12) = ``generate_patch_array()['patch_array'][:, :, w]`` of the reference for the marker
        size in use (w = 15 for the default marker_size = 3); a NumPy array or CPU tensor."""
        import numpy as np

        pa = np.ascontiguousarray(np.asarray(patches), np.uint8)
        if pa.shape != (10, 10, 12, 12):
            raise _lib.TxError("marker patches must have shape (10, 10, 12, 12)")
        self._check(self.lib.tx_set_marker_patches(self.h, pa.ctypes.data))

    def marker_overlay(self, markers: torch.Tensor | None, rgb: torch.Tensor | None = None, apply: bool = True,
                       rgb_out: torch.Tensor | None = None, marker_img_out: torch.Tensor | None = None,
                       rgb_u8_out: torch.Tensor | None = None) -> None:
        """One launch for all envs: marker image (``draw_markers``), RGB modulated by it (the RL task's overlay loop) and / or
        the uint8 observation. ``rgb_out`` may be ``rgb`` itself (in place)."""
        for t, dt in ((markers, torch.float32), (rgb, torch.float32), (rgb_out, torch.float32), (marker_img_out, torch.uint8), (rgb_u8_out, torch.uint8)):
            if t is not None and (t.device != self.device or t.dtype != dt or not t.is_contiguous()):
                raise _lib.TxError("marker_overlay: tensors must be contiguous, on the engine's device, float32 / uint8 as documented")
        N = (markers if markers is not None else rgb).shape[0]
        M = markers.shape[2] if markers is not None else 0
        self._check(self.lib.tx_marker_overlay(self.h, _ptr(markers), N, M, _ptr(rgb), int(bool(apply)), _ptr(rgb_out),
                                               _ptr(marker_img_out), _ptr(rgb_u8_out)))

    def set_phase_ticks(self, ticks: torch.Tensor | None) -> None:
        """Profiling: int64 device tensor (2*N, 40) receiving per-CTA phase clock stamps, or None to disable."""
        self._check(self.lib.tx_debug_set_ticks(self.h, _ptr(ticks)))

    def set_debug_flags(self, flags: int) -> None:
        self._check(self.lib.tx_debug_set_flags(self.h, int(flags)))

    def step_host(self, hm_host: torch.Tensor, rgb_host: torch.Tensor, depth_host: torch.Tensor | None = None,
                  theta_host: torch.Tensor | None = None, markers_host: torch.Tensor | None = None) -> None:
        """End-to-end call on HOST buffers (H2D + fused path + D2H inside), what bench.py's ``e2e`` times."""
        N = hm_host.shape[0]
        self._check(self.lib.tx_step_host(self.h, _ptr(hm_host), _ptr(theta_host), N, _ptr(rgb_host), _ptr(depth_host),
                                          _ptr(markers_host)))
