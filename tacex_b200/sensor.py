"""Headless stand-in for ``tacex.GelSightSensor`` -- the owner of the plug-in boundary.

The reference sensor is an Isaac Lab ``SensorBase`` that owns a ``TiledCamera`` (RTX depth render) and dispatches to
the simulator plug-ins (ref: source/tacex/tacex/gelsight_sensor.py:31-79 construction, :108-113 lazy ``data``,
:147-201 ``reset``, :203-337 ``_initialize_impl``, :342-378 ``_update_buffers_impl``, :581-593 ``_get_height_map``).
Isaac Sim is upstream and out of scope, so this class keeps the same attributes the plug-ins read
(``cfg``, ``_data.output``, ``_num_envs``, ``_device``, ``_indentation_depth``, ``camera``, ``optical_simulator``,
``marker_motion_simulator``) and the same call order, with the camera replaced by recorded depth maps
(:meth:`set_camera_depth`). ``tacex_tasks`` code that reads ``sensor.data.output["tactile_rgb"]`` /
``sensor.indentation_depth`` works unchanged against it.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Sequence

import torch

from .simulators import FOTSMarkerSimulatorCfg, GelSightSimulatorCfg, MarkerParams, TaximSimulatorCfg


@dataclass
class GelSightSensorData:
    """ref: source/tacex/tacex/gelsight_sensor_data.py:6-24"""

    position: Any = None
    orientation: Any = None
    intrinsic_matrix: Any = None
    image_resolution: tuple | None = None
    output: dict | None = None


@dataclass
class Dimensions:
    width: float = 0.0
    length: float = 0.0
    height: float = 0.0


@dataclass
class SensorCameraCfg:
    prim_path_appendix: str = "/Camera"
    update_period: float = 0
    resolution: tuple = (32, 24)  # (W, H)
    data_types: list = field(default_factory=lambda: ["depth"])
    clipping_range: tuple = (0.0, 1.0)


@dataclass
class GelSightSensorCfg:
    """Field-for-field ``GelSightSensorCfg`` (ref: source/tacex/tacex/gelsight_sensor_cfg.py:13-64)."""

    prim_path: str = "/World/envs/env_.*/gelsight"
    update_period: float = 0.0
    debug_vis: bool = False
    case_dimensions: Dimensions = field(default_factory=Dimensions)
    gelpad_dimensions: Dimensions = field(default_factory=Dimensions)
    sensor_camera_cfg: SensorCameraCfg = field(default_factory=SensorCameraCfg)
    data_types: list = field(default_factory=lambda: ["tactile_rgb", "marker_motion", "height_map", "camera_depth"])
    optical_sim_cfg: GelSightSimulatorCfg | None = None
    marker_motion_sim_cfg: GelSightSimulatorCfg | None = None
    compute_indentation_depth_class: str = "optical_sim"
    device: str = "cuda"
    num_envs: int = 1  # Isaac Lab derives this from the scene; headless it is explicit


def gelsight_mini_cfg(calib_path: str, num_envs: int, with_markers: bool = True, marker_rows: int = 9,
                      marker_cols: int = 11, device: str = "cuda") -> GelSightSensorCfg:
    """GelSight Mini preset (ref: source/tacex_assets/tacex_assets/sensors/gelsight_mini/gsmini_cfg.py:14-77)."""
    cfg = GelSightSensorCfg(
        case_dimensions=Dimensions(32 / 1000, 28 / 1000, 24 / 1000),
        gelpad_dimensions=Dimensions(20.75 / 1000, 25.25 / 1000, 4.5 / 1000),
        sensor_camera_cfg=SensorCameraCfg(resolution=(320, 240), clipping_range=(0.024, 0.029)),
        data_types=["tactile_rgb", "height_map"] + (["marker_motion"] if with_markers else []),
        optical_sim_cfg=TaximSimulatorCfg(calib_folder_path=calib_path, gelpad_height=4.5 / 1000,
                                          gelpad_to_camera_min_distance=0.024, with_shadow=False,
                                          tactile_img_res=(320, 240), device=device),
        marker_motion_sim_cfg=(
            FOTSMarkerSimulatorCfg(
                lamb=[0.00125, 0.00021, 0.00038], pyramid_kernel_size=[51, 21, 11, 5], kernel_size=5,
                marker_params=MarkerParams(num_markers_col=marker_cols, num_markers_row=marker_rows,
                                           num_markers=marker_rows * marker_cols, x0=15, y0=26, dx=26, dy=29),
                tactile_img_res=(320, 240), device=device, frame_transformer_cfg=None)
            if with_markers else None
        ),
        compute_indentation_depth_class="optical_sim",
        device=device,
        num_envs=num_envs,
    )
    return cfg


class GelSightSensor:
    cfg: GelSightSensorCfg

    def __init__(self, cfg: GelSightSensorCfg, gelpad_obj=None):
        self.cfg = cfg
        self._prim_view = None
        self.camera = None  # the RTX TiledCamera of the reference; replaced by set_camera_depth()
        self.gelpad_obj = gelpad_obj
        self._indentation_depth: torch.Tensor | None = None
        self.optical_simulator = None
        self.marker_motion_simulator = None
        self.compute_indentation_depth_func = None
        self._data = GelSightSensorData()
        self._data.output = dict.fromkeys(self.cfg.data_types, None)
        self._num_envs = int(cfg.num_envs)
        self._device = cfg.device
        self._is_outdated = True
        self._camera_depth: torch.Tensor | None = None

        if self.cfg.optical_sim_cfg is not None:
            self.optical_simulator = self.cfg.optical_sim_cfg.simulation_approach_class(sensor=self, cfg=self.cfg.optical_sim_cfg)
        if self.cfg.marker_motion_sim_cfg is not None:
            if (self.optical_simulator is not None) and (
                self.cfg.optical_sim_cfg.simulation_approach_class == self.cfg.marker_motion_sim_cfg.simulation_approach_class
            ):
                self.marker_motion_simulator = self.optical_simulator
            else:
                self.marker_motion_simulator = self.cfg.marker_motion_sim_cfg.simulation_approach_class(
                    sensor=self, cfg=self.cfg.marker_motion_sim_cfg
                )
        self._initialize_impl()

    # -- properties (same names as the reference) --------------------------------------------------------------------
    @property
    def device(self):
        return self._device

    @property
    def data(self) -> GelSightSensorData:
        self._update_outdated_buffers()
        return self._data

    @property
    def num_instances(self) -> int:
        return self._num_envs

    @property
    def indentation_depth(self):
        return self._indentation_depth

    @property
    def camera_resolution(self):
        return self.cfg.sensor_camera_cfg.resolution[0], self.cfg.sensor_camera_cfg.resolution[1]

    @property
    def tactile_image_shape(self):
        r = self.cfg.optical_sim_cfg.tactile_img_res
        return r[1], r[0], 3

    # -- recorded camera ---------------------------------------------------------------------------------------------
    def set_camera_depth(self, depth_m: torch.Tensor) -> None:
        """Recorded depth image [m], (N, H_cam, W_cam) or (N, H_cam, W_cam, 1): stands in for
        ``camera.data.output['depth']`` of the TiledCamera."""
        if depth_m.dim() == 4:
            depth_m = depth_m[..., 0]
        self._camera_depth = depth_m.to(self._device)
        self._is_outdated = True

    # -- SensorBase-like life cycle ------------------------------------------------------------------------------------
    def _initialize_impl(self):
        dev = self._device
        self._ALL_INDICES = torch.arange(self._num_envs, device=dev, dtype=torch.long)
        self._frame = torch.zeros(self._num_envs, device=dev, dtype=torch.long)
        self._indentation_depth = torch.zeros((self._num_envs,), device=dev)
        W, H = self.camera_resolution
        self._data.output["height_map"] = torch.zeros((self._num_envs, H, W), device=dev)
        if self.optical_simulator is not None:
            self.optical_simulator._initialize_impl()
        if self.marker_motion_simulator is not None and self.marker_motion_simulator is not self.optical_simulator:
            self.marker_motion_simulator._initialize_impl()
        if "tactile_rgb" in self.cfg.data_types:
            r = self.cfg.optical_sim_cfg.tactile_img_res
            self._data.output["tactile_rgb"] = torch.zeros((self._num_envs, r[1], r[0], 3), device=dev)
        if "marker_motion" in self.cfg.data_types:
            self._data.output["marker_motion"] = torch.zeros(
                (self._num_envs, 2, self.cfg.marker_motion_sim_cfg.marker_params.num_markers, 2), device=dev
            )
        if (self.cfg.compute_indentation_depth_class == "optical_sim") and (self.optical_simulator is not None):
            self.compute_indentation_depth_func = self.optical_simulator.compute_indentation_depth
        elif (self.cfg.compute_indentation_depth_class == "marker_motion_sim") and (self.marker_motion_simulator is not None):
            self.compute_indentation_depth_func = self.marker_motion_simulator.compute_indentation_depth
        else:
            self.compute_indentation_depth_func = None
        self.reset()

    def reset(self, env_ids: Sequence[int] | None = None):
        """Same semantics as the reference (gelsight_sensor.py:147-201): a partial reset recomputes ALL envs."""
        if env_ids is None:
            env_ids = self._ALL_INDICES
        self._indentation_depth[env_ids] = 0
        self._data.output["height_map"][env_ids] = 0
        if (self.optical_simulator is not None) and ("tactile_rgb" in self._data.output):
            self._data.output["tactile_rgb"][:] = self.optical_simulator.optical_simulation()
            self.optical_simulator.reset()
        if (self.marker_motion_simulator is not None) and ("marker_motion" in self._data.output):
            self._data.output["marker_motion"][:] = self.marker_motion_simulator.marker_motion_simulation()
            self._data.output["init_marker_pos"] = ([0], [0])
            self.marker_motion_simulator.reset()
        self._frame[env_ids] = 0

    def update(self, dt: float, force_recompute: bool = False):
        self._is_outdated = True
        if force_recompute:
            self._update_outdated_buffers()

    def _update_outdated_buffers(self):
        if self._is_outdated:
            self._update_buffers_impl(self._ALL_INDICES)
            self._is_outdated = False

    def _get_height_map(self):
        """depth [m] -> height map [mm]; inf -> far clipping plane (ref: gelsight_sensor.py:581-593)."""
        if self._camera_depth is not None:
            hm = self._data.output["height_map"]
            hm[:] = self._camera_depth
            hm[torch.isinf(hm)] = self.cfg.sensor_camera_cfg.clipping_range[1]
            hm *= 1000
            return hm

    def _get_camera_depth(self):
        """Normalised uint8 depth image for debug visualisation (ref: gelsight_sensor.py:557-579), same arithmetic: mm, minus
        the near plane, divided by the far plane, x 255, truncated to uint8; (N, H_cam, W_cam, 1)."""
        if self._camera_depth is not None:
            d = self._camera_depth.clone()
            lo, hi = self.cfg.sensor_camera_cfg.clipping_range
            d[torch.isinf(d)] = hi
            d *= 1000.0
            d -= lo * 1000
            d /= hi * 1000
            W, H = self.camera_resolution
            self._data.output["camera_depth"] = (d * 255).type(dtype=torch.uint8).reshape((self._num_envs, H, W, 1))
        return self._data.output.get("camera_depth")

    def _update_buffers_impl(self, env_ids):
        """ref: gelsight_sensor.py:342-378 -- same order: height map, indentation depth, RGB, markers."""
        self._frame[env_ids] += 1
        fused = None
        if (self.compute_indentation_depth_func is not None and self._camera_depth is not None
                and self.cfg.compute_indentation_depth_class == "optical_sim"
                and hasattr(self.optical_simulator, "fused_update_from_depth")):
            # the depth pre-processing of _get_height_map is fused into the kernel's load stage (tx_render_depth)
            fused = self.optical_simulator.fused_update_from_depth(self._camera_depth, self.cfg.sensor_camera_cfg.clipping_range[1])
        if fused is not None:
            self._indentation_depth[:] = fused
        elif self.compute_indentation_depth_func is not None:
            self._get_height_map()
            self._indentation_depth[:] = self.compute_indentation_depth_func()
        if "camera_depth" in self._data.output:
            self._get_camera_depth()
        if (self.optical_simulator is not None) and ("tactile_rgb" in self.cfg.data_types):
            self._data.output["tactile_rgb"][:] = self.optical_simulator.optical_simulation()
        if (self.marker_motion_simulator is not None) and ("marker_motion" in self.cfg.data_types):
            self._data.output["marker_motion"][:] = self.marker_motion_simulator.marker_motion_simulation()
