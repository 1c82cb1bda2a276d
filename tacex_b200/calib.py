"""Init-time calibration tables for the Taxim optical model (host side, runs once per sensor).

Restates the table preparation the reference does in ``TaximTorch.__init__`` and its cached getters
(ref: source/tacex/tacex/simulation_approaches/gpu_taxim/sim/taxim_torch.py:73-95 polynomial table and gel map,
:136-164 background / feature / gel-map getters, :363-412 Gaussian kernels, :414-430 initial-frame processing;
parameters: .../sim/taxim_impl.py:17-47,89-95 and ``params.json`` in the calibration folder).

The tables are what the C-ABI library consumes (``tx_upload_tables``): nothing here is on the per-step path.
Two sources are supported:

* a calibration folder in the reference's format (``polycalib.npz``, ``dataPack.npz``, ``gelmap.npy``,
  ``params.json``) -- what ``TaximSimulatorCfg.calib_folder_path`` points at;
* a pre-baked ``.npz`` written by :meth:`TaximTables.save` (used by the benchmark on machines without the
  reference's asset tree).
"""

from __future__ import annotations

import json
import math
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

EPS_KERNEL = 1e-5  # weight of the outermost tap (ref: taxim_torch.py:397)


def gaussian_kernel_size(sigma: float) -> int:
    """Odd kernel size such that the outermost tap weighs < 1e-5 (ref: taxim_torch.py:396-403)."""
    s = np.array(sigma)
    return int(np.round(np.sqrt(-2 * np.log(EPS_KERNEL * np.sqrt(2 * np.pi) * s)) * s).astype(np.int_) // 2 * 2 + 1)


def gaussian_taps(sigma: float, kernel_size: int | None = None) -> torch.Tensor:
    """Normalised float32 1-D Gaussian taps, same float32 operation sequence as the reference
    (ref: taxim_torch.py:363-367): linspace -> exp(-0.5 (x / sigma)^2) -> divide by the sum."""
    ks = gaussian_kernel_size(sigma) if kernel_size is None else int(kernel_size)
    x = torch.linspace(-(ks - 1) * 0.5, (ks - 1) * 0.5, steps=ks)
    pdf = torch.exp(-0.5 * (x / sigma).pow(2))
    return (pdf / pdf.sum()).to(torch.float32)


def _blur_separable(img: torch.Tensor, sigma_xy: tuple[float, float]) -> torch.Tensor:
    """Reflect-padded Gaussian blur of a (C, H, W) float image by direct separable correlation (float64
    accumulate). Init-time only; the reference uses an FFT product here (taxim_torch.py:382-412)."""
    kx = gaussian_taps(sigma_xy[0]).double()
    ky = gaussian_taps(sigma_xy[1]).double()
    px, py = (kx.numel() - 1) // 2, (ky.numel() - 1) // 2
    x = img.double()[:, None]  # (C,1,H,W)
    x = F.pad(x, (px, px, 0, 0), mode="reflect")
    x = F.conv2d(x, kx.view(1, 1, 1, -1))
    x = F.pad(x, (0, 0, py, py), mode="reflect")
    x = F.conv2d(x, ky.view(1, 1, -1, 1))
    return x[:, 0].float()


def _resize_bilinear_aa(img: torch.Tensor, shape: tuple[int, int]) -> torch.Tensor:
    """torchvision's default tensor resize (bilinear, antialias) as the reference calls it (taxim_torch.py:136-164)."""
    if tuple(img.shape[-2:]) == tuple(shape):
        return img.clone()
    return F.interpolate(img[None], size=list(shape), mode="bilinear", align_corners=False, antialias=True)[0]


@dataclass
class TaximParams:
    """``params.json`` of a calibration folder, scaled to a concrete image shape (ref: taxim_impl.py:17-47)."""

    pixmm: float = 0.0295
    num_bins: int = 125
    calib_w: int = 640
    calib_h: int = 480
    contact_scale: float = 0.4
    deform_pyramid_sigma_rel: tuple = ()
    deform_final_sigma_rel: tuple = ()
    initial_frame_sigma_rel: tuple = ()
    frame_mixing_percentage: float = 0.15
    diff_threshold: float = 5
    raw: dict = field(default_factory=dict)

    @classmethod
    def from_json(cls, d: dict) -> "TaximParams":
        s, q = d["simulator"], d["sensor"]
        return cls(
            pixmm=float(q["pixmm"]),
            num_bins=int(q["num_bins"]),
            calib_w=int(q["w"]),
            calib_h=int(q["h"]),
            contact_scale=float(s["contact_scale"]),
            deform_pyramid_sigma_rel=tuple(tuple(v) for v in s["deform_pyramid_sigma_rel"]),
            deform_final_sigma_rel=tuple(s["deform_final_sigma_rel"]),
            initial_frame_sigma_rel=tuple(s["initial_frame_sigma_rel"]),
            frame_mixing_percentage=float(s["frame_mixing_percentage"]),
            diff_threshold=float(s["diff_threshold"]),
            raw=d,
        )

    def pyramid_sigmas(self, shape: tuple[int, int]) -> list[tuple[float, float]]:
        """[(sigma_x, sigma_y)] per pyramid level: ``*_rel`` times W for x and H for y (taxim_impl.py:33-47)."""
        wv, hv = self.deform_pyramid_sigma_rel
        return [(a * shape[1], b * shape[0]) for a, b in zip(wv, hv)]

    def final_sigma(self, shape: tuple[int, int]) -> tuple[float, float]:
        return (self.deform_final_sigma_rel[0] * shape[1], self.deform_final_sigma_rel[1] * shape[0])

    def blur_taps(self, shape: tuple[int, int]) -> list[tuple[torch.Tensor, torch.Tensor]]:
        """(x taps, y taps) for the pyramid levels followed by the final blur."""
        sig = self.pyramid_sigmas(shape) + [self.final_sigma(shape)]
        return [(gaussian_taps(sx), gaussian_taps(sy)) for sx, sy in sig]


@dataclass
class TaximTables:
    """Everything the device library needs, for one output shape (H, W)."""

    shape: tuple[int, int]
    params: TaximParams
    poly_grad: torch.Tensor  # (3, nb, nb, 6) float32, channel order RGB
    background: torch.Tensor  # (3, H, W) float32 in [0, 1]
    gel_map: torch.Tensor | None  # (H, W) float32 or None when the gel map is flat (== 0 after the shift)
    gel_map_shift: float = 0.0

    @classmethod
    def from_calib_folder(cls, folder: str | Path, shape: tuple[int, int] = (240, 320)) -> "TaximTables":
        folder = Path(folder)
        with (folder / "params.json").open() as f:
            params = TaximParams.from_json(json.load(f))
        calib_shape = (params.calib_h, params.calib_w)

        # polynomial table; grad_b and grad_r are swapped in the file (ref: taxim_torch.py:73-80)
        data = np.load(str(folder / "polycalib.npz"))
        poly = torch.from_numpy(np.stack([data["grad_b"], data["grad_g"], data["grad_r"]], axis=0) / 255).float()

        # gel map: blur, scale to mm, shift so that its maximum is 0 (ref: taxim_torch.py:82-90)
        gel_np = np.load(str(folder / "gelmap.npy"))
        if float(gel_np.max()) == float(gel_np.min()):
            gel_map, shift = None, float(gel_np.max()) * params.pixmm
        else:
            gel = torch.from_numpy(gel_np).float()[None]
            gel = _blur_separable(gel, params.final_sigma(calib_shape))[0] * params.pixmm
            shift = gel.max().item()
            gel_map = _resize_bilinear_aa((gel - shift)[None], shape)[0].contiguous()

        # background: BGR -> RGB, blur with a large kernel, mix (ref: taxim_torch.py:92-94, 414-430)
        f0 = np.load(str(folder / "dataPack.npz"), allow_pickle=True)["f0"] / 255
        f0 = torch.from_numpy(f0).float().permute(2, 0, 1).flip(0)
        sig0 = (params.initial_frame_sigma_rel[0] * calib_shape[1], params.initial_frame_sigma_rel[1] * calib_shape[0])
        f0_blur = _blur_separable(f0, sig0)
        d_i = torch.mean(f0_blur - f0, dim=0)
        fmp = params.frame_mixing_percentage
        bg_proc = torch.where((d_i < params.diff_threshold).unsqueeze(0), fmp * f0_blur + (1 - fmp) * f0, f0)
        background = _resize_bilinear_aa(bg_proc, shape).contiguous()
        return cls(tuple(shape), params, poly.contiguous(), background, gel_map, shift)

    # -- pre-baked tables ---------------------------------------------------------------------------------------
    def save(self, path: str | Path) -> None:
        np.savez_compressed(
            str(path),
            shape=np.array(self.shape),
            params=np.array(json.dumps(self.params.raw)),
            poly_grad=self.poly_grad.numpy(),
            background=self.background.numpy(),
            gel_map=self.gel_map.numpy() if self.gel_map is not None else np.zeros((0,), np.float32),
            gel_map_shift=np.array(self.gel_map_shift),
        )

    @classmethod
    def load(cls, path: str | Path) -> "TaximTables":
        z = np.load(str(path), allow_pickle=False)
        params = TaximParams.from_json(json.loads(str(z["params"])))
        gel = torch.from_numpy(z["gel_map"]) if z["gel_map"].size else None
        return cls(
            tuple(int(v) for v in z["shape"]),
            params,
            torch.from_numpy(z["poly_grad"]),
            torch.from_numpy(z["background"]),
            gel,
            float(z["gel_map_shift"]),
        )

    @property
    def bin_widths(self) -> tuple[float, float]:
        nb = self.params.num_bins
        return 0.5 * math.pi / (nb - 1), 2 * math.pi / (nb - 1)


@dataclass
class ShadowTables:
    """Init-time data of the Taxim SHADOW branch (``with_shadow=True``; ref: taxim_torch.py:96-125, 260-346, taxim_impl.py:17-47).

    Host-side preparation only: the tables are what a device implementation (and the CPU checker) consumes. The trigonometry of
    the ray fan is evaluated HERE, once, in float32 with torch -- exactly the tensors the reference builds at init -- so that no
    per-pixel ``cos`` / ``sin`` (whose last bit is library dependent) decides which pixel a shadow sample lands in.
    """

    table: np.ndarray        # (3, D, Hn, S) float32 / 255, RGB order, rows padded with +inf (the reference's `__shadow_table_padded`)
    fan_cos: np.ndarray      # (D, F) float32: cos of direction + fan offset
    fan_sin: np.ndarray      # (D, F) float32
    depth_0: float           # 0.4 (taxim_torch.py:97)
    height_precision: float
    discretize_precision: float
    step_x: float            # shadow_step(shape)[1] -- the reference multiplies the X coordinate by the H-relative value (quirk)
    step_y: float            # shadow_step(shape)[0]
    dilate_rounds: tuple     # ((ky, kx), (ky, kx)): box kernels of the two dilation rounds of the contact mask
    blur_taps: tuple         # (x taps, y taps) of the shadow blur

    @classmethod
    def from_calib_folder(cls, folder: str | Path, shape: tuple[int, int] = (240, 320)) -> "ShadowTables":
        folder = Path(folder)
        with (folder / "params.json").open() as f:
            sim = json.load(f)["simulator"]
        data = np.load(str(folder / "shadowTable.npz"), allow_pickle=True)
        direction = torch.from_numpy(data["shadowDirections"]).float()
        fan_angle = float(sim["fan_angle"])
        n_rays = int(fan_angle * 2 / float(sim["fan_precision"]))
        fan = direction.unsqueeze(-1) + torch.linspace(-fan_angle, fan_angle, n_rays)
        # BGR -> RGB flip; the reference also "appends an empty entry for heights outside the range", but the appended array has a
        # zero-length height axis, so nothing is appended and out-of-range heights read the LAST REAL height entry (kept: quirk)
        tab = np.flip(data["shadowTable"], axis=0)
        n_max = max(len(e) for e in tab.reshape(-1))
        padded = np.array([list(e) + [np.inf] * (n_max - len(e)) for e in tab.reshape(-1)], dtype=np.float32)
        padded = (torch.from_numpy(padded).reshape(tab.shape + (-1,)) / 255).numpy()

        def rel(name):  # `*_rel` parameters: (w value * W, h value * H)  (taxim_impl.py:33-47)
            v = sim[name + "_rel"]
            return v[0] * shape[1], v[1] * shape[0]

        ks_total = np.round(np.array(rel("shadow_attachment_kernel_size")) * 2).astype(np.int_)
        first = ks_total // 2
        rounds = tuple(tuple(int(v) for v in np.flip(np.maximum(1, k))) for k in (first, ks_total - first))
        step = rel("shadow_step")
        sig = rel("shadow_blur_sigma")
        return cls(padded, torch.cos(fan).numpy(), torch.sin(fan).numpy(), 0.4, float(sim["height_precision"]),
                   float(sim["discretize_precision"]), float(step[1]), float(step[0]), rounds,
                   (gaussian_taps(sig[0]).numpy(), gaussian_taps(sig[1]).numpy()))

    # -- pre-baked tables ---------------------------------------------------------------------------------------
    def save(self, path: str | Path) -> None:
        np.savez_compressed(
            str(path), table=self.table, fan_cos=self.fan_cos, fan_sin=self.fan_sin,
            scalars=np.array([self.depth_0, self.height_precision, self.discretize_precision, self.step_x, self.step_y]),
            dilate_rounds=np.array(self.dilate_rounds), taps_x=self.blur_taps[0], taps_y=self.blur_taps[1],
        )

    @classmethod
    def load(cls, path: str | Path) -> "ShadowTables":
        z = np.load(str(path), allow_pickle=False)
        sc = [float(v) for v in z["scalars"]]
        return cls(z["table"], z["fan_cos"], z["fan_sin"], sc[0], sc[1], sc[2], sc[3], sc[4],
                   tuple(tuple(int(v) for v in r) for r in z["dilate_rounds"]), (z["taps_x"], z["taps_y"]))
