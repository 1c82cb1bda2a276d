"""Imports the UNMODIFIED reference implementation from /root/reference (read-only) -- TEST INFRASTRUCTURE ONLY.

Works only in the build container (the GPU box has no /root/reference); used by oracle/make_golden.py to generate
the committed fixtures under tests/golden/ and by the container-only tests that pin the canonical restatement
against the executed reference.

Nothing is copied: the reference modules are imported by path and called through their own entry points
(``sim.Taxim`` / ``TaximTorch.render_direct`` and the name-mangled privates the reference's FOTS wrapper itself
uses, fots_marker_sim.py:128-129; ``MarkerMotion.marker_sim``).
"""

from __future__ import annotations

import importlib.util
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF_ROOT = Path("/root/reference")
TAXIM_PKG = REF_ROOT / "source/tacex/tacex/simulation_approaches/gpu_taxim"
FOTS_FILE = REF_ROOT / "source/tacex/tacex/simulation_approaches/fots/sim/marker_motion.py"
CALIB_DIR = REF_ROOT / "source/tacex_assets/tacex_assets/data/Sensors/GelSight_Mini/calibs/640x480"


def available() -> bool:
    return TAXIM_PKG.exists() and CALIB_DIR.exists()


_taxim = None
SCATTER_CAPTURE: dict = {}


def _torch_scatter_stand_in():
    """torch_scatter is not in this image. The reference's shadow branch calls exactly one function of it,
    ``scatter_min(src [C, K], index [K], dim_size=S, out=out [C, S])`` (taxim_torch.py:330-335); its published semantics
    (out[c, index[k]] = min(out[c, index[k]], src[c, k])) are torch's own ``scatter_reduce_(..., 'amin')``."""
    m = types.ModuleType("torch_scatter")

    def scatter_min(src, index, dim=-1, out=None, dim_size=None):
        idx = index.expand_as(src) if index.dim() < src.dim() else index
        out.scatter_reduce_(-1, idx, src, reduce="amin", include_self=True)
        SCATTER_CAPTURE["shadow_img_flat"] = out.clone()  # lets the tests see the reference's intermediate shadow image
        return out, None

    m.scatter_min = scatter_min
    return m


def load_taxim(device: str = "cpu"):
    """``sim.Taxim(calib_folder=..., backend='torch')`` of the reference. torch_scatter (absent here) is only used by the
    shadow branch (taxim_torch.py:330): see _torch_scatter_stand_in."""
    global _taxim
    if _taxim is None:
        sys.modules.setdefault("torch_scatter", _torch_scatter_stand_in())
        if str(TAXIM_PKG) not in sys.path:
            sys.path.insert(0, str(TAXIM_PKG))
        import sim  # noqa: PLC0415  (the reference's gpu_taxim/sim package)

        _taxim = sim.Taxim(calib_folder=CALIB_DIR, backend="torch", device=device)
    return _taxim


def load_marker_motion_cls():
    spec = importlib.util.spec_from_file_location("ref_marker_motion", str(FOTS_FILE))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.MarkerMotion


def ref_indentation_depth(hm_mm: torch.Tensor, gelpad_height=0.0045, min_dist=0.024) -> torch.Tensor:
    """Verbatim arithmetic of TaximSimulator.compute_indentation_depth (taxim_sim.py:115-131); the class itself
    cannot be imported without Isaac Sim (omni.usd)."""
    height_map = hm_mm / 1000
    min_distance_obj = height_map.amin((1, 2))
    dist_obj_sensor_case = min_distance_obj - min_dist
    dist_obj_sensor_case = torch.where(dist_obj_sensor_case < 0, 0, dist_obj_sensor_case)
    return torch.where(dist_obj_sensor_case <= gelpad_height, (gelpad_height - dist_obj_sensor_case) * 1000, 0)


def ref_tables(tx, shape=(240, 320)) -> dict:
    return {
        "poly_grad": tx._TaximTorch__poly_grad.clone(),
        "background": tx._TaximTorch__get_background_img_cached(tuple(shape)).clone(),
        "gel_map": tx._TaximTorch__get_gel_map_cached(tuple(shape)).clone(),
        "gel_map_shift": tx._TaximTorch__gel_map_shift,
    }


def ref_render(tx, hm_mm: torch.Tensor, press_mm: torch.Tensor) -> torch.Tensor:
    """(N,H,W,3) exactly as TaximSimulator.optical_simulation returns it (taxim_sim.py:104-111)."""
    return tx.render_direct(hm_mm, with_shadow=False, press_depth=press_mm, orig_hm_fmt=False).movedim(1, 3).contiguous()


def ref_render_shadow(tx, hm_mm: torch.Tensor, press_mm: torch.Tensor) -> torch.Tensor:
    """(N,H,W,3) with_shadow=True (taxim_torch.py:260-346), otherwise as ref_render."""
    return tx.render_direct(hm_mm, with_shadow=True, press_depth=press_mm, orig_hm_fmt=False).movedim(1, 3).contiguous()


def ref_deformed_gel(tx, hm_mm: torch.Tensor, press_mm: torch.Tensor):
    sh = tx._TaximTorch__get_shifted_height_map(press_mm, hm_mm)
    return tx._TaximTorch__compute_gel_pad_deformation(sh)


def ref_normals_bins(tx, deformed: torch.Tensor):
    """grad_mag, grad_dir, idx_mag, idx_dir as __render computes them (taxim_torch.py:237-247)."""
    px = deformed / tx.sensor_params.pixmm
    mag, dr = tx._TaximTorch__generate_normals(-px)
    x_binr = 0.5 * torch.pi / (tx.sensor_params.num_bins - 1)
    y_binr = 2 * torch.pi / (tx.sensor_params.num_bins - 1)
    return mag, dr, torch.floor(mag / x_binr).long(), torch.floor((dr + torch.pi) / y_binr).long()


class RefFots:
    """The reference's per-env FOTS loop (fots_marker_sim.py:128-182) around the unmodified ``MarkerMotion``; the yaw
    that the reference reads from Isaac Lab's FrameTransformer is an input here."""

    def __init__(self, tx, rows=9, cols=11, x0=15, y0=26, mm2pix=19.58, W=320, H=240):
        MarkerMotion = load_marker_motion_cls()
        bg = tx._TaximTorch__get_background_img_cached((H, W)).movedim(0, 2).cpu().numpy()
        self.mm = MarkerMotion(frame0_blur=bg, mm2pix=mm2pix, num_markers_col=cols, num_markers_row=rows,
                               tactile_img_width=W, tactile_img_height=H, lamb=[0.00125, 0.00021, 0.00038], x0=x0, y0=y0)
        self.tx = tx
        self.init = np.stack((self.mm.init_marker_x_pos, self.mm.init_marker_y_pos), axis=-1).reshape(-1, 2)
        self.traj = None

    def step(self, hm_mm: torch.Tensor, press_mm: torch.Tensor, theta: np.ndarray) -> torch.Tensor:
        N = hm_mm.shape[0]
        if self.traj is None:
            self.traj = [[] for _ in range(N)]
        deformed_gel, contact_mask = ref_deformed_gel(self.tx, hm_mm, press_mm)
        deformed_gel = deformed_gel.max() - deformed_gel
        out = torch.zeros((N, 2, self.init.shape[0], 2))
        out[:, 0] = torch.tensor(self.init)
        for env_id in range(N):
            if press_mm[env_id].item() > 0.0:
                contact_points = torch.argwhere(contact_mask[env_id])
                mean = torch.mean(contact_points.float(), dim=0).cpu().numpy()
                mean[0] = (mean[0] - self.mm.tactile_img_height / 2) / self.mm.mm2pix
                mean[1] = (mean[1] - self.mm.tactile_img_width / 2) / self.mm.mm2pix
                self.traj[env_id].append([mean[1], mean[0], theta[env_id]])
                mx, my = self.mm.marker_sim(deformed_gel[env_id].cpu().numpy(), contact_mask[env_id].cpu().numpy(),
                                            self.traj[env_id])
            else:
                self.traj[env_id] = []
                mx, my = self.mm.init_marker_x_pos, self.mm.init_marker_y_pos
            out[env_id, 1] = torch.tensor(np.stack((mx, my), axis=-1).reshape(-1, 2))
        return out


# ---- executing METHODS of reference classes whose modules need Isaac Sim ------------------------------------------------
def ref_methods(path: Path, class_name: str, names: list[str], namespace: dict) -> dict:
    """Compiles the named methods of ``class_name`` straight from the reference file (ast; nothing is copied into this repo) and
    returns them as plain functions taking ``self`` first. The enclosing modules import omni / isaaclab at the top, so they
    cannot be imported; the method bodies themselves only need what ``namespace`` supplies."""
    import ast

    tree = ast.parse(Path(path).read_text())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == class_name)
    fns = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in names]
    ns = dict(namespace)
    exec(compile(ast.Module(body=fns, type_ignores=[]), str(path), "exec"), ns)
    return {n: ns[n] for n in names}


def ref_class(path: Path, class_name: str, names: list[str], namespace: dict, base: type = object) -> type:
    """Like :func:`ref_methods`, but returns a CLASS holding the named methods (derived from ``base``), for methods that call
    ``super()``: instantiate it with ``cls.__new__(cls)`` and set the attributes the methods read."""
    import ast

    tree = ast.parse(Path(path).read_text())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == class_name)
    fns = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in names]
    node = ast.ClassDef(name=class_name, bases=[ast.Name(id="_RefBase", ctx=ast.Load())], keywords=[], body=fns, decorator_list=[],
                        type_params=[])
    mod = ast.fix_missing_locations(ast.Module(body=[node], type_ignores=[]))
    ns = dict(namespace)
    ns["_RefBase"] = base
    exec(compile(mod, str(path), "exec"), ns)
    return ns[class_name]


class RefTaximSimulator:
    """The reference's ``TaximSimulator.optical_simulation`` / ``compute_indentation_depth`` (taxim_sim.py:80-131) EXECUTED on a
    stand-in ``self``: the real ``sim.Taxim`` instance, a dict as sensor output, the preset's cfg values."""

    def __init__(self, tx, num_envs: int, tactile_img_res=(320, 240), gelpad_height=0.0045, min_dist=0.024, with_shadow=False):
        import types

        import torchvision.transforms.functional as F

        self.m = ref_methods(REF_ROOT / "source/tacex/tacex/simulation_approaches/gpu_taxim/taxim_sim.py", "TaximSimulator",
                             ["optical_simulation", "compute_indentation_depth"], {"torch": torch, "F": F, "np": np})
        W, H = tactile_img_res
        self.me = types.SimpleNamespace(
            sensor=types.SimpleNamespace(_data=types.SimpleNamespace(output={})),
            cfg=types.SimpleNamespace(tactile_img_res=tactile_img_res, with_shadow=with_shadow, gelpad_height=gelpad_height,
                                      gelpad_to_camera_min_distance=min_dist),
            _device="cpu", _taxim=tx, _indentation_depth=torch.zeros(num_envs), tactile_rgb_img=torch.zeros((num_envs, H, W, 3)),
        )

    def step(self, height_map_mm: torch.Tensor):
        """-> (indentation depth [N], tactile RGB [N, H, W, 3]) in the reference's call order."""
        self.me.sensor._data.output["height_map"] = height_map_mm
        depth = self.m["compute_indentation_depth"](self.me).clone()
        rgb = self.m["optical_simulation"](self.me).clone()
        return depth, rgb


class RefFotsSimulatorMethod:
    """The reference's ``FOTSMarkerSimulator.marker_motion_simulation`` (fots_marker_sim.py:114-184) EXECUTED from the file on a
    stand-in ``self``. Isaac Lab's FrameTransformer / ``euler_xyz_from_quat`` (third party, absent) are replaced by a recorded
    yaw per env: the stand-in hands the method a yaw-only quaternion and the published yaw formula atan2(2 (w z + x y),
    1 - 2 (y^2 + z^2)). Used to check that ``RefFots`` above (the loop the golden markers were generated with) is that method."""

    def __init__(self, tx, num_envs: int, rows=9, cols=11, x0=15, y0=26, mm2pix=19.58, W=320, H=240):
        import types

        import torchvision.transforms.functional as F

        MarkerMotion = load_marker_motion_cls()
        bg = tx._TaximTorch__get_background_img_cached((H, W)).movedim(0, 2).cpu().numpy()
        mm = MarkerMotion(frame0_blur=bg, mm2pix=mm2pix, num_markers_col=cols, num_markers_row=rows, tactile_img_width=W,
                          tactile_img_height=H, lamb=[0.00125, 0.00021, 0.00038], x0=x0, y0=y0)

        def euler_xyz_from_quat(q):  # (K, 4) wxyz -> roll, pitch, yaw; only the yaw is used by the caller
            w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
            yaw = torch.atan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z))
            return torch.zeros_like(yaw), torch.zeros_like(yaw), yaw

        self.fn = ref_methods(REF_ROOT / "source/tacex/tacex/simulation_approaches/fots/fots_marker_sim.py", "FOTSMarkerSimulator",
                              ["marker_motion_simulation"],
                              {"torch": torch, "F": F, "np": np, "euler_xyz_from_quat": euler_xyz_from_quat})["marker_motion_simulation"]
        init = np.stack((mm.init_marker_x_pos, mm.init_marker_y_pos), axis=-1).reshape(-1, 2)
        md = torch.zeros((num_envs, 2, init.shape[0], 2))
        md[:, 0] = torch.tensor(init)
        ft = types.SimpleNamespace(update=lambda dt: None,
                                   data=types.SimpleNamespace(target_pos_source=torch.zeros((num_envs, 1, 3)),
                                                              target_quat_source=torch.zeros((num_envs, 1, 4))))
        self.me = types.SimpleNamespace(
            sensor=types.SimpleNamespace(_indentation_depth=None, _data=types.SimpleNamespace(output={"traj": [[] for _ in range(num_envs)]})),
            cfg=types.SimpleNamespace(tactile_img_res=(W, H), mm_to_pixel=mm2pix), _device="cpu", _taxim=tx, marker_motion_sim=mm,
            frame_transformer=ft, marker_data=md,
        )

    def step(self, hm_mm: torch.Tensor, press_mm: torch.Tensor, theta: np.ndarray) -> torch.Tensor:
        th = torch.as_tensor(theta, dtype=torch.float32)
        q = torch.zeros((th.shape[0], 1, 4))
        q[:, 0, 0], q[:, 0, 3] = torch.cos(th / 2), torch.sin(th / 2)
        self.me.frame_transformer.data.target_quat_source = q
        self.me.sensor._indentation_depth = press_mm
        self.me.sensor._data.output["height_map"] = hm_mm
        return self.fn(self.me).clone()
