"""Runs the UNMODIFIED reference (staged by baseline/stage_ref.py under baseline/_ref/) for bench.py -- measurement
infrastructure only; the product never imports this module.

The reference's wrapper classes (TaximSimulator, FOTSMarkerSimulator) import Isaac Sim / Isaac Lab at module level and cannot be
loaded headless; what runs here is exactly what they call on the hot path:
  ``sim.Taxim(calib_folder, backend="torch", device=...)`` -> ``TaximTorch.render_direct(height_map, with_shadow=False,
  press_depth=...)`` (taxim_sim.py:92-103), the indentation-depth arithmetic of taxim_sim.py:115-131, and the per-env
  ``MarkerMotion.marker_sim`` loop of fots_marker_sim.py:128-182 on the deformation the reference recomputes for it.
"""
from __future__ import annotations

import importlib.util
import sys
import time
import types
from pathlib import Path

REF = Path(__file__).resolve().parent / "_ref"
TAXIM_PKG = REF / "gpu_taxim"
FOTS_FILE = REF / "fots" / "sim" / "marker_motion.py"
CALIB = REF / "calibs" / "640x480"


def available() -> bool:
    return (TAXIM_PKG / "sim" / "taxim_torch.py").exists() and FOTS_FILE.exists() and (CALIB / "params.json").exists()


_cache: dict = {}


def load_taxim(device: str = "cpu"):
    if ("taxim", device) in _cache:
        return _cache[("taxim", device)]
    if "torch_scatter" not in sys.modules:  # only the shadow branch uses it; absent in this image
        sys.modules["torch_scatter"] = types.ModuleType("torch_scatter")
    if str(TAXIM_PKG) not in sys.path:
        sys.path.insert(0, str(TAXIM_PKG))
    import sim  # noqa: PLC0415  (the reference's gpu_taxim/sim package, unmodified)

    tx = sim.Taxim(calib_folder=CALIB, backend="torch", device=device)
    _cache[("taxim", device)] = tx
    return tx


def load_marker_motion():
    if "mm" not in _cache:
        spec = importlib.util.spec_from_file_location("ref_marker_motion", str(FOTS_FILE))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _cache["mm"] = mod.MarkerMotion
    return _cache["mm"]


def indentation_depth(hm_mm, gelpad_height=0.0045, min_dist=0.024):
    """Arithmetic of TaximSimulator.compute_indentation_depth (taxim_sim.py:115-131) with the reference's tensor ops."""
    import torch

    height_map = hm_mm / 1000
    d = height_map.amin((1, 2)) - min_dist
    d = torch.where(d < 0, 0, d)
    return torch.where(d <= gelpad_height, (gelpad_height - d) * 1000, 0)


class RefStep:
    """One sensor update of the reference for a batch of envs: indentation depth + Taxim RGB (NHWC) + FOTS markers."""

    def __init__(self, device: str = "cpu", rows: int = 7, cols: int = 9, with_markers: bool = True):
        import numpy as np

        self.tx = load_taxim(device)
        self.device = device
        self.with_markers = with_markers
        if with_markers:
            MarkerMotion = load_marker_motion()
            bg = self.tx._TaximTorch__get_background_img_cached((240, 320)).movedim(0, 2).cpu().numpy()
            self.mm = MarkerMotion(frame0_blur=bg, mm2pix=19.58, num_markers_col=cols, num_markers_row=rows, tactile_img_width=320,
                                   tactile_img_height=240, lamb=[0.00125, 0.00021, 0.00038], x0=15, y0=26)
            self.init = np.stack((self.mm.init_marker_x_pos, self.mm.init_marker_y_pos), axis=-1).reshape(-1, 2)
        self.traj = None

    def __call__(self, hm_mm, theta):
        import numpy as np
        import torch

        N = hm_mm.shape[0]
        press = indentation_depth(hm_mm)
        rgb = self.tx.render_direct(hm_mm, with_shadow=False, press_depth=press, orig_hm_fmt=False).movedim(1, 3)  # taxim_sim.py:104-111
        out = None
        if self.with_markers:
            if self.traj is None or len(self.traj) != N:
                self.traj = [[] for _ in range(N)]
            tx = self.tx
            sh = tx._TaximTorch__get_shifted_height_map(press, hm_mm)          # fots_marker_sim.py:128-129
            dg, mask = tx._TaximTorch__compute_gel_pad_deformation(sh)
            dg = dg.max() - dg
            out = torch.zeros((N, 2, self.init.shape[0], 2))
            out[:, 0] = torch.tensor(self.init)
            for e in range(N):
                if press[e].item() > 0.0:
                    pts = torch.argwhere(mask[e])
                    mean = torch.mean(pts.float(), dim=0).cpu().numpy()
                    mean[0] = (mean[0] - 240 / 2) / 19.58
                    mean[1] = (mean[1] - 320 / 2) / 19.58
                    self.traj[e].append([mean[1], mean[0], float(theta[e])])
                    mx, my = self.mm.marker_sim(dg[e].cpu().numpy(), mask[e].cpu().numpy(), self.traj[e])
                    self.traj[e] = self.traj[e][:1]  # keep the trajectory bounded over repeated bench steps (first + current sample)
                else:
                    self.traj[e] = []
                    mx, my = self.mm.init_marker_x_pos, self.mm.init_marker_y_pos
                out[e, 1] = torch.tensor(np.stack((mx, my), axis=-1).reshape(-1, 2))
        return rgb, press, out


def time_cpu(hm_mm, batch: int, reps: int, warmup: int = 1, with_markers: bool = True):
    """frames/s of the reference on the host cores: ``reps`` sensor updates of ``batch`` envs each."""
    import os

    import torch

    # every host thread the process may use: torchrun exports OMP_NUM_THREADS=1 to its workers, which would leave the reference on ONE core
    n_cpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if torch.get_num_threads() < n_cpu:
        torch.set_num_threads(n_cpu)
    step = RefStep("cpu", with_markers=with_markers)
    th = torch.zeros(batch)
    n = hm_mm.shape[0]
    for w in range(warmup):
        step(hm_mm[:batch], th)
    t0 = time.perf_counter()
    for r in range(reps):
        lo = (r * batch) % max(n - batch + 1, 1)
        step(hm_mm[lo:lo + batch], th)
    dt = time.perf_counter() - t0
    return batch * reps / dt, torch.get_num_threads()


def time_cuda(hm_dev, batches=(256, 1024, 4096), reps: int = 3):
    """The reference's intended mode (device="cuda"): Taxim RGB only (render_direct + NHWC), CUDA-event timed."""
    import torch

    step = RefStep("cuda", with_markers=False)
    out = {}
    for b in batches:
        if b > hm_dev.shape[0]:
            continue
        try:
            x = hm_dev[:b]
            th = None
            for _ in range(2):
                step(x, th)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                rgb, _, _ = step(x, th)
                rgb = rgb.contiguous()
            e1.record()
            torch.cuda.synchronize()
            out[str(b)] = {"frames_per_s": b * reps / (e0.elapsed_time(e1) / 1e3), "ms_per_call": e0.elapsed_time(e1) / reps}
            del rgb
        except Exception as exc:  # out of memory at the largest batch is a result, not a failure of the bench
            out[str(b)] = {"error": f"{type(exc).__name__}: {str(exc)[:120]}"}
        torch.cuda.empty_cache()
    return out
