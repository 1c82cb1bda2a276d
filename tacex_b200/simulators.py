"""Drop-in implementations of the reference's simulator plug-in interface, backed by libtacex_b200.so.

Interface mirrored (same method names, argument meaning, return shapes, error behaviour):
  GelSightSimulator ABC          ref: source/tacex/tacex/simulation_approaches/gelsight_simulator.py:17-56
  GelSightSimulatorCfg           ref: .../gelsight_simulator_cfg.py:7-16
  TaximSimulatorCfg              ref: .../gpu_taxim/taxim_sim_cfg.py:12-36
  TaximSimulator                 ref: .../gpu_taxim/taxim_sim.py:20-138
  FOTSMarkerSimulatorCfg         ref: .../fots/fots_marker_sim_cfg.py:15-75
  FOTSMarkerSimulator            ref: .../fots/fots_marker_sim.py:25-184

When the reference package ``tacex`` is importable (inside Isaac Lab) the classes below subclass ITS
``GelSightSimulator`` so ``GelSightSensor`` accepts them unchanged; otherwise they subclass the local ABC.
Select them with ``cfg.simulation_approach_class = B200TaximSimulator``.
"""

from __future__ import annotations

import os

from abc import ABC, abstractmethod
from dataclasses import dataclass, field
from pathlib import Path
from typing import Any

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib
from .calib import TaximTables
from .engine import TactileEngine

try:  # inside Isaac Lab: plug into the reference's own ABC
    from tacex.simulation_approaches.gelsight_simulator import GelSightSimulator as _RefBase  # type: ignore
except Exception:  # headless / no Isaac Sim
    _RefBase = None


class _LocalGelSightSimulator(ABC):
    """Same contract as the reference ABC (gelsight_simulator.py:17-56)."""

    def __init__(self, sensor, cfg):
        self.cfg = cfg
        self.sensor = sensor
        self._device = self.sensor.device if self.cfg.device is None else self.cfg.device

    @abstractmethod
    def _initialize_impl(self):
        raise NotImplementedError

    def optical_simulation(self):
        raise NotImplementedError

    def marker_motion_simulation(self):
        raise NotImplementedError

    def compute_indentation_depth(self):
        raise NotImplementedError

    @abstractmethod
    def reset(self):
        raise NotImplementedError

    def _set_debug_vis_impl(self, debug_vis: bool):
        raise NotImplementedError(f"Debug visualization is not implemented for {self.__class__.__name__}.")

    def _debug_vis_callback(self, event):
        raise NotImplementedError(f"Debug visualization is not implemented for {self.__class__.__name__}.")


GelSightSimulator = _RefBase if _RefBase is not None else _LocalGelSightSimulator


# ---- cfg classes (plain dataclasses; field-for-field the reference's @configclass definitions) -------------------
@dataclass
class GelSightSimulatorCfg:
    simulation_approach_class: type | None = None
    device: str | None = "cuda"


@dataclass
class TaximSimulatorCfg(GelSightSimulatorCfg):
    simulation_approach_class: type | None = None  # None -> B200TaximSimulator
    calib_folder_path: str = ""
    device: str | None = "cuda"
    with_shadow: bool = False
    tactile_img_res: tuple = (320, 240)  # (W, H)
    gelpad_height: float = 0.0045
    gelpad_to_camera_min_distance: float = 0.024

    def __post_init__(self):
        if self.simulation_approach_class is None:
            self.simulation_approach_class = B200TaximSimulator


@dataclass
class MarkerParams:
    num_markers_col: int = 11
    num_markers_row: int = 9
    num_markers: int = 99
    x0: float = 15.0
    y0: float = 26.0
    dx: float = 26.0
    dy: float = 29.0


@dataclass
class FOTSMarkerSimulatorCfg(GelSightSimulatorCfg):
    simulation_approach_class: type | None = None  # None -> B200FOTSMarkerSimulator
    calib_folder_path: str = ""
    device: str | None = None
    with_shadow: bool = False
    tactile_img_res: tuple = (320, 240)
    lamb: list = field(default_factory=list)  # ignored, like the reference (fots_marker_sim.py:77 hard-codes it)
    ball_radius: float = 4.70 / 2
    mm_to_pixel: float = 19.58
    pyramid_kernel_size: list = field(default_factory=list)
    kernel_size: int = 0
    marker_params: MarkerParams = field(default_factory=MarkerParams)
    init_marker_pos: tuple = ([[]], [[]])
    frame_transformer_cfg: Any = None

    def __post_init__(self):
        if self.simulation_approach_class is None:
            self.simulation_approach_class = B200FOTSMarkerSimulator


def _load_tables(path: str, shape: tuple[int, int]) -> TaximTables:
    """``calib_folder_path``: a calibration folder in the reference's format, a folder of pre-baked ``tables_{W}x{H}.npz`` files
    (one per tactile resolution), or one pre-baked ``.npz``."""
    p = Path(path)
    if p.is_dir():
        baked = p / f"tables_{shape[1]}x{shape[0]}.npz"
        return TaximTables.load(baked) if baked.exists() else TaximTables.from_calib_folder(p, shape)
    t = TaximTables.load(p)
    if tuple(t.shape) != tuple(shape):
        raise RuntimeError(f"{p} holds tables for {t.shape[1]}x{t.shape[0]}, tactile_img_res asks for {shape[1]}x{shape[0]}")
    return t


def _is_cuda(dev) -> bool:
    return torch.device(dev).type == "cuda"


class B200TaximSimulator(GelSightSimulator):
    """Taxim optical simulation on the fused sm_100a kernel (drop-in for ``TaximSimulator``)."""

    cfg: TaximSimulatorCfg

    def __init__(self, sensor, cfg: TaximSimulatorCfg):
        self.sensor = sensor
        super().__init__(sensor=sensor, cfg=cfg)

    def _initialize_impl(self):
        self._device = self.sensor.device if self.cfg.device is None else self.cfg.device
        if not _is_cuda(self._device):
            raise RuntimeError("B200TaximSimulator runs on a CUDA (sm_100a) device only; there is no CPU path")
        self._num_envs = self.sensor._num_envs
        W, H = self.cfg.tactile_img_res
        if self.cfg.with_shadow and (H, W) != (240, 320):
            # the shadow branch (tx_render_shadow, bit-exact vs the canonical restatement on a B200) belongs to the 240 x 320 kernel
            raise NotImplementedError("with_shadow=True is implemented for tactile_img_res=(320, 240) only")
        self.img_res = self.cfg.tactile_img_res
        tables = _load_tables(self.cfg.calib_folder_path, (H, W))
        mcfg = getattr(self.sensor.cfg, "marker_motion_sim_cfg", None)
        rows = cols = 0
        x0 = y0 = 0.0
        mm2pix = 19.58
        # the marker model shares the optical kernel's blur pyramid only at the GelSight Mini's 320 x 240; at any other tactile
        # resolution (e.g. the RL tasks' 32 x 24) the FOTS simulator runs its own 320 x 240 engine, as the reference does
        if (H, W) == (240, 320) and mcfg is not None and hasattr(mcfg, "marker_params") and hasattr(mcfg, "mm_to_pixel"):
            rows, cols = mcfg.marker_params.num_markers_row, mcfg.marker_params.num_markers_col
            x0, y0, mm2pix = mcfg.marker_params.x0, mcfg.marker_params.y0, mcfg.mm_to_pixel
        self.engine = TactileEngine(
            tables, max_envs=self._num_envs, device=self._device, marker_rows=rows, marker_cols=cols, marker_x0=x0,
            marker_y0=y0, mm2pix=mm2pix, gelpad_height_m=self.cfg.gelpad_height,
            gelpad_to_cam_min_m=self.cfg.gelpad_to_camera_min_distance,
        )
        if self.cfg.with_shadow:
            from .calib import ShadowTables

            p = Path(self.cfg.calib_folder_path)
            baked = p / f"shadow_tables_{W}x{H}.npz" if p.is_dir() else None
            self.engine.upload_shadow_tables(ShadowTables.load(baked) if baked is not None and baked.exists()
                                             else ShadowTables.from_calib_folder(p, (H, W)))
        dev = self.engine.device
        self._indentation_depth = torch.zeros((self._num_envs,), device=dev)
        self._own_rgb = torch.zeros((self._num_envs, H, W, 3), device=dev)
        self.tactile_rgb_img = self._own_rgb
        # tactile image without indentation (the reference shows the resized background, taxim_sim.py:62-75)
        self.background_img = tables.background.movedim(0, 2).contiguous().to(dev)
        self.tactile_rgb_img[:] = self.background_img
        self._stamp = None  # (height-map identity/version, indentation-depth version) of the last fused launch
        self._sensor_depth_version = None

    # -- helpers ---------------------------------------------------------------------------------------------------
    def _camera_mode(self, hm: torch.Tensor) -> bool:
        """True when the sensor camera is coarser than the tactile image and the fused resize of the kernel's load stage
        applies (ref: taxim_sim.py:88-89 F.resize; the RL tasks / the FEM preset run 32x32 / 32x24 cameras)."""
        W, H = self.cfg.tactile_img_res
        hc, wc = int(hm.shape[1]), int(hm.shape[2])
        if (hc, wc) == (H, W):
            return False
        if not self.engine.generic and hc <= H and wc <= W and hc * wc <= 9600:
            if getattr(self.engine, "cam_hw", None) != (hc, wc):
                self.engine.set_camera_resolution(hc, wc)
            return True
        return False

    def _height_map(self) -> torch.Tensor:
        hm = self.sensor._data.output["height_map"]
        W, H = self.cfg.tactile_img_res
        if hm.device != self.engine.device:
            hm = hm.to(self.engine.device)
        if (hm.shape[1], hm.shape[2]) != (H, W):
            # camera finer than the tactile image (down-sampling): torchvision's F.resize semantics (taxim_sim.py:88-89) by the
            # library's own antialias kernel; torch only for scales beyond its 8 taps per axis
            try:
                return self.engine.resize(hm.contiguous().float())
            except _lib.TxError:
                hm = F.interpolate(hm[:, None], size=[H, W], mode="bilinear", align_corners=False, antialias=True)[:, 0]
        return hm.contiguous()

    def _render(self, press, depth_out=None):
        raw = self.sensor._data.output["height_map"]
        if not self.cfg.with_shadow and self._camera_mode(raw):
            cam = raw if raw.device == self.engine.device else raw.to(self.engine.device)
            self.engine.render_camera(cam.contiguous(), press=press, out=self._rgb_target(), depth_out=depth_out)
            return
        hm = self._height_map()
        if press is None and tuple(raw.shape[1:]) != tuple(hm.shape[1:]):
            # the indentation depth belongs to the CAMERA-resolution map (taxim_sim.py:115-131), the render to the resized one
            cam = raw if raw.device == self.engine.device else raw.to(self.engine.device)
            press = self.engine.indentation_depth_frames(cam.contiguous().float())
            if depth_out is not None:
                depth_out.copy_(press)
        if self.cfg.with_shadow:
            self.engine.render_shadow(hm, press, out=self._rgb_target(), depth_out=depth_out if press is None else None)
            return
        self.engine.render(hm, press, out=self._rgb_target(), depth_out=depth_out if press is None else None)

    def _rgb_target(self) -> torch.Tensor:
        """Render straight into the sensor-owned output tensor when it exists: ``output[:] = returned`` in the sensor
        (gelsight_sensor.py:373-375) then degenerates to a self-copy, which torch skips (same storage, same layout)."""
        out = self.sensor._data.output.get("tactile_rgb") if self.sensor._data.output else None
        t = self._own_rgb
        if (isinstance(out, torch.Tensor) and out.shape == t.shape and out.dtype == t.dtype and out.device == t.device
                and out.is_contiguous()):
            self.tactile_rgb_img = out
            return out
        return t

    def _hm_stamp(self):
        hm = self.sensor._data.output["height_map"]
        return (hm.data_ptr(), hm._version, tuple(hm.shape))

    # -- plug-in interface -------------------------------------------------------------------------------------------
    def compute_indentation_depth(self):
        """Indentation depth [mm] (ref: taxim_sim.py:115-131). Runs the FUSED kernel: the RGB frame of the same
        height map is produced in the same launch and reused by ``optical_simulation`` if nothing changed."""
        self._render(None, depth_out=self._indentation_depth)
        self._stamp = (self._hm_stamp(), self._indentation_depth._version)
        return self._indentation_depth

    def fused_update_from_depth(self, depth_m: torch.Tensor, clip_max_m: float):
        """One launch for GelSightSensor._get_height_map + compute_indentation_depth + optical_simulation, starting from the
        raw camera depth [m] (headless sensor only; the reference sensor calls the three steps separately). Fills the
        sensor's ``height_map`` output in place and returns the indentation depth."""
        hm = self.sensor._data.output["height_map"]
        W, H = self.cfg.tactile_img_res
        if self.cfg.with_shadow:
            return None  # the shadow post-pass has no fused depth entry point: the sensor takes the reference's call order
        if tuple(depth_m.shape[1:]) != (H, W) or hm.shape != depth_m.shape or not hm.is_contiguous() or hm.device != self.engine.device:
            return None
        self.engine.render_depth(depth_m.contiguous(), clip_max_m, out=self._rgb_target(), depth_out=self._indentation_depth,
                                 height_map_out=hm)
        self._stamp = (self._hm_stamp(), self._indentation_depth._version)
        return self._indentation_depth

    def optical_simulation(self):
        """Tactile RGB (num_envs, H, W, 3) float32 in [0, 1] (ref: taxim_sim.py:80-113)."""
        if self._stamp != (self._hm_stamp(), self._indentation_depth._version):
            self._render(self._indentation_depth)
            self._stamp = (self._hm_stamp(), self._indentation_depth._version)
        sd = getattr(self.sensor, "_indentation_depth", None)
        self._sensor_depth_version = sd._version if isinstance(sd, torch.Tensor) else None
        return self.tactile_rgb_img

    def reset(self):
        self._indentation_depth = torch.zeros((self._num_envs,), device=self.engine.device)
        self.tactile_rgb_img = self._own_rgb  # never clobber the sensor-owned output (it keeps the last render)
        self.tactile_rgb_img[:] = self.background_img
        self._stamp = None

    def _set_debug_vis_impl(self, debug_vis: bool):
        pass  # GUI only in the reference (omni.ui); nothing to do headless

    def _debug_vis_callback(self, event):
        pass


class B200FOTSMarkerSimulator(GelSightSimulator):
    """FOTS marker motion on the GPU (drop-in for ``FOTSMarkerSimulator``); reuses the gel deformation of the
    optical simulator's last launch instead of recomputing the blur pyramid."""

    cfg: FOTSMarkerSimulatorCfg

    def __init__(self, sensor, cfg: FOTSMarkerSimulatorCfg):
        self.sensor = sensor
        super().__init__(sensor=sensor, cfg=cfg)
        self.frame_transformer = None
        if self.cfg.frame_transformer_cfg is not None:
            from isaaclab.sensors import FrameTransformer  # type: ignore  # only inside Isaac Lab

            self.frame_transformer = FrameTransformer(self.cfg.frame_transformer_cfg)

    def _initialize_impl(self):
        self._device = self.sensor.device if self.cfg.device is None else self.cfg.device
        self._num_envs = self.sensor._num_envs
        if (self.sensor.optical_simulator is not None) and isinstance(self.sensor.optical_simulator, B200TaximSimulator):
            self._taxim: B200TaximSimulator = self.sensor.optical_simulator
        else:
            raise RuntimeError(
                "Currently FOTS simulation approach has to be used in combination with GPU-Taxim as optical-simulator."
            )
        eng = self._taxim.engine
        W, H = self.cfg.tactile_img_res
        # The marker model works on the gel deformation at ITS tactile resolution (ref: fots_marker_sim.py:121-129). When the
        # optical simulator renders at another one (the RL tasks: optical 32 x 24, markers 320 x 240, ref:
        # ball_rolling_taxim_fots.py:321,331) the deformation cannot be shared and this simulator owns a second engine.
        self._own_engine = (eng.H, eng.W) != (H, W)
        if self._own_engine:
            mp = self.cfg.marker_params
            eng = TactileEngine(
                _load_tables(self._taxim.cfg.calib_folder_path, (H, W)), max_envs=self._num_envs, device=eng.device,
                marker_rows=mp.num_markers_row, marker_cols=mp.num_markers_col, marker_x0=mp.x0, marker_y0=mp.y0,
                mm2pix=self.cfg.mm_to_pixel, gelpad_height_m=self._taxim.cfg.gelpad_height,
                gelpad_to_cam_min_m=self._taxim.cfg.gelpad_to_camera_min_distance,
            )
        if eng.M != self.cfg.marker_params.num_markers_row * self.cfg.marker_params.num_markers_col:
            raise RuntimeError("marker grid of the optical simulator's engine does not match marker_params")
        self.engine = eng
        dev = eng.device
        self._indentation_depth = torch.zeros((self._num_envs,), device=dev)
        self.init_marker_pos = torch.stack(
            (torch.from_numpy(eng.marker_x).float(), torch.from_numpy(eng.marker_y).float()), dim=-1
        )
        self.img_res = self.cfg.tactile_img_res
        self.marker_data = torch.zeros((self._num_envs, 2, eng.M, 2), device=dev)
        self.marker_data[:, 0] = self.init_marker_pos.to(dev)
        self.marker_data[:, 1] = self.init_marker_pos.to(dev)
        # per-env trajectory state: first sample of the contact episode + number of samples. The reference keeps a
        # python list per env in sensor._data.output["traj"]; only traj[0], traj[-1] and len() are ever used.
        self.traj0 = torch.zeros((self._num_envs, 4), device=dev)
        self.traj_len = torch.zeros((self._num_envs,), device=dev, dtype=torch.int32)
        self.sensor._data.output["traj"] = self.traj0
        self.theta = torch.zeros((self._num_envs,), device=dev)
        self._scratch_rgb = None
        # marker dot patches for draw_markers / the marker overlay (ref: fots_marker_sim.py:112, generate_patch_array()): the slice
        # for the reference's default marker_size = 3, produced by executing the reference (oracle/make_golden_overlay.py)
        self._marker_img = None
        self._patches_ready = False
        baked = Path(__file__).resolve().parent / "data" / "marker_patches_size3.npy"
        if (H, W) == (240, 320) and not eng.generic and baked.exists():
            import numpy as np

            eng.set_marker_patches(np.load(baked))
            self._patches_ready = True
        if self.frame_transformer is not None:
            self.frame_transformer._initialize_impl()
            self.frame_transformer._is_initialized = True

    # -- marker image / overlay (ref: fots_marker_sim.py:346-384; ball_rolling_taxim_fots.py:918-937) -----------------------
    def set_patch_array(self, patch_array_dict: dict, marker_size: float = 3) -> None:
        """Use the dot patches of the reference's ``generate_patch_array()`` for another marker size."""
        import math

        w = math.floor((marker_size - patch_array_dict["base_circle_radius"]) * patch_array_dict["super_resolution_ratio"])
        self.engine.set_marker_patches(patch_array_dict["patch_array"][:, :, w])
        self._patches_ready = True

    def draw_markers_batch(self, marker_data: torch.Tensor | None = None) -> torch.Tensor:
        """The marker image of EVERY env in one launch: (num_envs, H, W) uint8, what ``draw_markers(marker_data[i, 1])`` returns
        per env in the reference."""
        if not self._patches_ready:
            raise RuntimeError("marker patches are not set (320 x 240 only): call set_patch_array(generate_patch_array())")
        md = self.marker_data if marker_data is None else marker_data.to(self.engine.device, torch.float32).contiguous()
        if self._marker_img is None or self._marker_img.shape[0] != md.shape[0]:
            self._marker_img = torch.empty((md.shape[0], self.engine.H, self.engine.W), dtype=torch.uint8, device=self.engine.device)
        self.engine.marker_overlay(md, marker_img_out=self._marker_img)
        return self._marker_img

    def overlay_markers(self, tactile_rgb: torch.Tensor, marker_data: torch.Tensor | None = None, rgb_u8_out: torch.Tensor | None = None) -> torch.Tensor:
        """``tactile_rgb`` (num_envs, 240, 320, 3) float32 modulated IN PLACE by the marker image of every env -- the RL task's
        per-env loop (draw_markers -> tensor -> ``rgb * 255 * (marker / 255) / 255``) as one launch; optionally also the uint8
        observation."""
        if not self._patches_ready:
            raise RuntimeError("marker patches are not set (320 x 240 only): call set_patch_array(generate_patch_array())")
        md = self.marker_data if marker_data is None else marker_data.to(self.engine.device, torch.float32).contiguous()
        self.engine.marker_overlay(md, tactile_rgb, apply=True, rgb_out=tactile_rgb, rgb_u8_out=rgb_u8_out)
        return tactile_rgb

    def _relative_yaw(self) -> torch.Tensor:
        """Yaw of the indenter relative to the sensor (ref: fots_marker_sim.py:147-159 via FrameTransformer);
        headless: the recorded value in sensor._data.output['indenter_yaw'] (zeros if absent)."""
        if self.frame_transformer is not None:
            from isaaclab.utils.math import euler_xyz_from_quat  # type: ignore

            self.frame_transformer.update(dt=0.001)
            _, _, yaw = euler_xyz_from_quat(self.frame_transformer.data.target_quat_source[:, 0])
            return yaw.to(self.engine.device, torch.float32).contiguous()
        yaw = self.sensor._data.output.get("indenter_yaw")
        if yaw is None:
            return self.theta
        return yaw.to(self.engine.device, torch.float32).contiguous()

    def marker_motion_simulation(self):
        """(num_envs, 2, M, 2): [:,0] initial, [:,1] current marker (x, y) px (ref: fots_marker_sim.py:114-184)."""
        self._indentation_depth = self.sensor._indentation_depth
        press = self._indentation_depth.to(self.engine.device, torch.float32).contiguous()
        tx = self._taxim
        # The engine still holds the gel deformation of the optical simulator's last launch. It is valid for this call
        # iff the height map has not changed since and the sensor's indentation-depth buffer is the unmodified copy
        # GelSightSensor made of the simulator's buffer (gelsight_sensor.py:361-365) -- tracked with tensor versions,
        # no host synchronisation. Otherwise recompute it for exactly (height_map, sensor._indentation_depth), as the
        # reference always does (fots_marker_sim.py:128-129).
        if self._own_engine:
            return self._markers_own_engine(press)
        reuse = (
            tx._stamp is not None
            and tx._stamp == (tx._hm_stamp(), tx._indentation_depth._version)
            and tx._sensor_depth_version == self.sensor._indentation_depth._version
        )
        if not reuse:
            if self._scratch_rgb is None:
                self._scratch_rgb = torch.empty_like(tx.tactile_rgb_img)
            self.engine.render(tx._height_map(), press, out=self._scratch_rgb)
            tx._stamp = None  # the engine's recorded deformation no longer belongs to the optical simulator's frame
        self.theta = self._relative_yaw()
        self.engine.fots_markers(press, self.theta, self.traj0, self.traj_len, out=self.marker_data)
        return self.marker_data

    def _markers_own_engine(self, press: torch.Tensor) -> torch.Tensor:
        """Marker resolution != optical resolution: the height map is resized to the marker resolution (F.resize, ref:
        fots_marker_sim.py:121-122; fused into the kernel's load stage when the camera is coarser) and the deformation is
        computed by this simulator's own engine for (height map, sensor._indentation_depth), as the reference always does."""
        W, H = self.cfg.tactile_img_res
        eng = self.engine
        hm = self.sensor._data.output["height_map"]
        hm = (hm if hm.device == eng.device else hm.to(eng.device)).contiguous()
        if self._scratch_rgb is None:
            self._scratch_rgb = torch.empty((self._num_envs, H, W, 3), device=eng.device)
        hc, wc = int(hm.shape[1]), int(hm.shape[2])
        if (hc, wc) == (H, W):
            eng.render(hm, press, out=self._scratch_rgb)
        elif hc <= H and wc <= W and hc * wc <= 9600:
            if getattr(eng, "cam_hw", None) != (hc, wc):
                eng.set_camera_resolution(hc, wc)
            eng.render_camera(hm, press=press, out=self._scratch_rgb)
        else:
            hm = F.interpolate(hm[:, None], size=[H, W], mode="bilinear", align_corners=False, antialias=True)[:, 0].contiguous()
            eng.render(hm, press, out=self._scratch_rgb)
        self.theta = self._relative_yaw()
        eng.fots_markers(press, self.theta, self.traj0, self.traj_len, out=self.marker_data)
        return self.marker_data

    def reset(self):
        pass

    def _set_debug_vis_impl(self, debug_vis: bool):
        pass

    def _debug_vis_callback(self, event):
        pass


# ---- FEM based marker simulation (ManiSkill-ViTac style) ------------------------------------------------------------------
@dataclass
class ManiSkillMarkerParams:
    num_markers: int = 128
    x0: float = 0
    y0: float = 0
    dx: float = 0
    dy: float = 0


@dataclass
class ManiSkillSimulatorCfg(GelSightSimulatorCfg):
    """Field-for-field ``ManiSkillSimulatorCfg`` (ref: .../fem_based/mani_skill_sim_cfg.py:10-70)."""

    simulation_approach_class: type | None = None  # None -> B200ManiSkillSimulator
    calib_folder_path: str = ""
    device: str | None = "cuda"
    marker_interval_range: tuple = (2.0625, 2.0625)
    marker_rotation_range: float = 0.0
    marker_translation_range: tuple = (0.0, 0.0)
    marker_pos_shift_range: tuple = (0.0, 0.0)
    marker_random_noise: float = 0.0
    marker_lose_tracking_probability: float = 0.0
    normalize: bool = False
    marker_flow_size: int = 128
    camera_params: tuple = (340, 325, 160, 125, 0.0)
    tactile_img_res: tuple = (320, 240)
    marker_params: ManiSkillMarkerParams = field(default_factory=ManiSkillMarkerParams)
    # not a reference field: where the sensor camera's optical axis meets the pad, in the pad frame [m]. The reference lays its
    # marker grid out in the CAMERA frame (x in [-8.25, 16.5] mm) and reads the camera pose from the sensor's USD asset; headless
    # it has to be stated. None = the legacy layout (a full 13 x 7 grid centred on the pad, no uv mask).
    camera_origin_in_pad_frame: tuple | None = None

    def __post_init__(self):
        if self.simulation_approach_class is None:
            self.simulation_approach_class = B200ManiSkillSimulator


class B200ManiSkillSimulator(GelSightSimulator):
    """Marker flow from the FEM gel surface (drop-in for ``ManiSkillSimulator``, ref: .../fem_based/mani_skill_sim.py:20-86 and
    sim/tactile_sensor_sapienipc_modified.py:189-413): barycentric markers on the gel's top surface projected with the
    sensor camera. ``sensor.gelpad_obj`` must be a :class:`tacex_b200.fem.GelPadSim`. Unlike the reference (env 0 only, weights
    recomputed every step) every env is read out in one launch with weights computed once."""

    cfg: ManiSkillSimulatorCfg

    def __init__(self, sensor, cfg: ManiSkillSimulatorCfg):
        self.sensor = sensor
        super().__init__(sensor=sensor, cfg=cfg)

    def _initialize_impl(self):
        from .fem import GelPadSim, marker_grid_weights

        self._device = self.sensor.device if self.cfg.device is None else self.cfg.device
        self._num_envs = self.sensor._num_envs
        gel = self.sensor.gelpad_obj
        if not isinstance(gel, GelPadSim):
            raise RuntimeError("B200ManiSkillSimulator needs sensor.gelpad_obj to be a tacex_b200.fem.GelPadSim")
        self.gel = gel
        if any(abs(v) > 0 for v in (self.cfg.marker_rotation_range, *self.cfg.marker_translation_range,
                                    *self.cfg.marker_pos_shift_range, self.cfg.marker_random_noise,
                                    self.cfg.marker_lose_tracking_probability)):
            raise NotImplementedError("randomised marker grids are not implemented (all ranges are 0 in the reference presets)")
        fx, fy, cx, cy = self.cfg.camera_params[:4]
        W, H = self.cfg.tactile_img_res
        if self.cfg.camera_origin_in_pad_frame is None:
            tri, w = marker_grid_weights(gel.mesh, pitch=self.cfg.marker_interval_range[0] * 1e-3, pad_to=self.cfg.marker_flow_size)
            gel.engine.set_markers(tri, w, intrinsics=(fx, fy, cx, cy), normalize=self.cfg.normalize, img_hw=(H, W))
        else:
            # the reference's layout: its camera-frame grid (_gen_marker_grid), the markers that fall on the gel surface
            # (_gen_marker_weight), then the uv mask / padding / normalize of gen_marker_flow
            from .fem import reference_marker_grid

            origin = np.asarray(self.cfg.camera_origin_in_pad_frame, np.float64)[:2]
            pts = reference_marker_grid(self.cfg.marker_interval_range[0]) - origin
            X = np.asarray(gel.mesh.X, np.float64)
            lo, hi = X[:, :2].min(0), X[:, :2].max(0)
            on = ((pts >= lo - 1e-12) & (pts <= hi + 1e-12)).all(1)
            tri, w = marker_grid_weights(gel.mesh, pad_to=int(on.sum()), points_xy=pts[on])
            gel.engine.set_markers(tri, w, cam_t=(origin[0], origin[1], 0.0285), intrinsics=(fx, fy, cx, cy), reference_tail=True,
                                   num_markers=self.cfg.marker_flow_size, img_hw=(H, W), normalize=self.cfg.normalize)
        self.marker_data = torch.zeros((self._num_envs, 2, self.cfg.marker_flow_size, 2), device=gel.engine.device)
        self._indentation_depth = torch.zeros((self._num_envs,), device=gel.engine.device)

    def marker_motion_simulation(self):
        self.gel.engine.markers(self.gel.x, out=self.marker_data)
        return self.marker_data

    def compute_indentation_depth(self):
        """Largest downward displacement of the gel's top surface [mm]."""
        top = torch.from_numpy(np.unique(self.gel.mesh.top_tris).astype(np.int64)).to(self.gel.x.device)
        dz = (self.gel.engine.X[top, 2][None] - self.gel.x[:, top, 2]).amax(1)
        self._indentation_depth[:] = torch.clamp(dz, min=0) * 1000
        return self._indentation_depth

    def reset(self):
        pass

    def _set_debug_vis_impl(self, debug_vis: bool):
        pass

    def _debug_vis_callback(self, event):
        pass
