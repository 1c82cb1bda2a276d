"""GelFemEngine -- PyTorch-facing owner of one ``tx_fem`` handle: the batched gel FEM substep and the FEM marker read-out.

Host-side mirror of the reference's UIPC glue for the gel pad only (ref: source/tacex_uipc/tacex_uipc/sim/uipc_sim.py:32-131
configuration and :250-252 ``step``; objects/uipc_object.py:442-470 constitution set-up; sim/uipc_attachments.py:118-142,
364-428 attachment of the gel to the rigid case; source/tacex/tacex/simulation_approaches/fem_based/mani_skill_sim.py and
sim/tactile_sensor_sapienipc_modified.py:189-413 marker grid / weights / projection). The indenter is a prescribed rigid
analytic body (the rigid-contact query of Isaac Sim is upstream and stubbed with recorded poses).
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from .gel_mesh import GelMesh, lame


@dataclass
class GelFemCfg:
    """Defaults = UipcSimCfg / UipcObject defaults used by the reference's GelSight Mini FEM preset."""

    youngs_modulus: float = 1e4      # 0.01 MPa (uipc_object.py:59)
    poisson_rate: float = 0.49       # (uipc_object.py:79)
    mass_density: float = 1e3        # (uipc_object.py:84)
    dt: float = 0.01                 # (uipc_sim.py)
    gravity: tuple = (0.0, 0.0, -9.8)
    attach_strength: float = 1000.0  # benchmark scene (ball_rolling_uipc.py:121); UipcIsaacAttachments default 100
    d_hat: float = 5e-4              # benchmark scene; libuipc default 0.01
    contact_resistance: float = 1e10  # 10 GPa (uipc_sim.py:110)
    newton_max_iter: int = 1024
    newton_velocity_tol: float = 0.05
    pcg_tol_rate: float = 1e-3
    line_search_max_iter: int = 8
    animator_substep: int = 1
    rest_volume_det: bool = True     # reproduce libuipc's det(Dm) elastic rest "volume" 
    friction_ratio: float = 0.5      # contact.default_friction_ratio (uipc_sim.py); 0 disables the lagged friction
    friction_eps_velocity: float = 0.01  # contact.eps_velocity [m/s]


def indenter_array(kind, centers, half, R=None, device="cuda") -> torch.Tensor:
    """Packs N indenter poses into the device layout of ``tx_fem_indenter`` (int + 15 doubles, 128 bytes each)."""
    centers = np.asarray(centers, np.float64).reshape(-1, 3)
    N = centers.shape[0]
    arr = (_lib.TxFemIndenter * N)()
    kinds = np.broadcast_to(np.asarray(kind), (N,))
    halfs = np.broadcast_to(np.asarray(half, np.float64), (N, 3))
    Rs = np.broadcast_to(np.eye(3) if R is None else np.asarray(R, np.float64), (N, 3, 3))
    for i in range(N):
        arr[i].type = int(kinds[i])
        arr[i].c[:] = centers[i].tolist()
        arr[i].R[:] = Rs[i].reshape(-1).tolist()
        arr[i].h[:] = halfs[i].tolist()
    buf = np.frombuffer(arr, dtype=np.uint8).copy()
    return torch.from_numpy(buf).to(device)


class GelFemEngine:
    def __init__(self, mesh: GelMesh, cfg: GelFemCfg | None = None, device: str | torch.device = "cuda"):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.TxError("GelFemEngine needs a CUDA (sm_100a) device; there is no CPU path")
        self.device = torch.device(device)
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        self.mesh, self.cfg = mesh, cfg or GelFemCfg()
        c = self.cfg
        lam, mu = lame(c.youngs_modulus, c.poisson_rate)
        g = _lib.TxFemConfig()
        g.V, g.T, g.A, g.S = len(mesh.X), len(mesh.tets), len(mesh.attach), len(mesh.surf)
        g.dt = c.dt
        g.gravity[:] = c.gravity
        g.mu, g.lam, g.density, g.attach_strength = mu, lam, c.mass_density, c.attach_strength
        g.d_hat, g.kappa = c.d_hat, c.contact_resistance
        g.newton_max_iter, g.velocity_tol, g.pcg_tol_rate = c.newton_max_iter, c.newton_velocity_tol, c.pcg_tol_rate
        g.pcg_max_iter_ratio, g.ls_max_iter, g.substep = 2, c.line_search_max_iter, c.animator_substep
        g.rest_volume_det = int(c.rest_volume_det)
        g.friction_mu, g.eps_velocity = c.friction_ratio, c.friction_eps_velocity
        self.g = g
        X = np.ascontiguousarray(mesh.X, np.float64)
        tets = np.ascontiguousarray(mesh.tets, np.int32)
        att = np.ascontiguousarray(mesh.attach, np.int32)
        surf = np.ascontiguousarray(mesh.surf, np.int32)
        with torch.cuda.device(idx):
            self.stream = torch.cuda.current_stream()
            h = C.c_void_p()
            rc = self.lib.tx_fem_create(C.byref(g), X.ctypes.data, tets.ctypes.data, att.ctypes.data, surf.ctypes.data, idx,
                                        C.c_void_p(self.stream.cuda_stream), C.byref(h))
            if rc != 0:
                raise _lib.TxError(f"tx_fem_create failed ({rc}): {self.lib.tx_fem_last_error(None).decode()}")
        self.h = h
        self.V, self.A, self.S, self.M = g.V, g.A, g.S, 0
        self.X = torch.from_numpy(X).to(self.device)

    def _check(self, rc):
        if rc != 0:
            raise _lib.TxError(f"libtacex_b200 fem error {rc}: {self.lib.tx_fem_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.tx_fem_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def mass(self) -> np.ndarray:
        m = np.empty(self.V)
        self._check(self.lib.tx_fem_get_mass(self.h, m.ctypes.data))
        return m

    def new_state(self, N: int):
        x = self.X[None].repeat(N, 1, 1).contiguous()
        return x, torch.zeros_like(x), x.clone()

    def rest_aim(self, N: int) -> torch.Tensor:
        return self.X[torch.from_numpy(self.mesh.attach.astype(np.int64)).to(self.device)][None].repeat(N, 1, 1).contiguous()

    def _chk_state(self, **tensors) -> None:
        for name, t in tensors.items():
            if not isinstance(t, torch.Tensor) or t.dtype != torch.float64 or t.device != self.device or not t.is_contiguous():
                raise _lib.TxError(f"{name} must be a contiguous float64 tensor on {self.device}")

    def step(self, x, v, x_prev, aim, ind_prev: torch.Tensor, ind_next: torch.Tensor, want_stats: bool = True):
        self._chk_state(x=x, v=v, x_prev=x_prev, aim=aim)
        N = x.shape[0]
        if tuple(x.shape) != (N, self.V, 3) or v.shape != x.shape or x_prev.shape != x.shape or aim.shape[0] != N:
            raise _lib.TxError("state tensors must have shape (N, V, 3), aim (N, A, 3)")
        for ind in (ind_prev, ind_next):
            if ind.dtype != torch.uint8 or ind.device != self.device or ind.numel() != N * C.sizeof(_lib.TxFemIndenter):
                raise _lib.TxError("indenter arrays must be packed tx_fem_indenter records (fem.indenter_array) on the engine's device")
        st = torch.zeros((N, C.sizeof(_lib.TxFemStats)), dtype=torch.uint8, device=self.device) if want_stats else None
        self._check(self.lib.tx_fem_step(self.h, x.data_ptr(), v.data_ptr(), x_prev.data_ptr(), aim.data_ptr(),
                                         ind_prev.data_ptr(), ind_next.data_ptr(), N, None if st is None else st.data_ptr()))
        return st

    def set_indenter_mesh(self, tri_local: np.ndarray | None) -> None:
        """Triangles (n, 3, 3) float64 [m] of the prescribed rigid mesh indenter in its own frame (indenter type 2 of
        ``indenter_array``; pose = centre + rotation per gel); ``None`` removes it. One mesh per engine."""
        if tri_local is None:
            self._check(self.lib.tx_fem_set_indenter_mesh(self.h, 0, None))
            return
        t = np.ascontiguousarray(tri_local, np.float64)
        if t.ndim != 3 or t.shape[1:] != (3, 3):
            raise _lib.TxError("tri_local must have shape (n, 3, 3)")
        self._check(self.lib.tx_fem_set_indenter_mesh(self.h, len(t), t.ctypes.data))

    def set_contact_surface(self, tris: np.ndarray | None) -> None:
        """Triangles (n, 3) int32 of the gel surface (e.g. ``mesh.top_tris``) that the VERTICES of the mesh indenter can touch: the
        second half of the vertex-face contact (a cone tip between four gel vertices is only seen this way). ``None`` switches it off."""
        if tris is None:
            self._check(self.lib.tx_fem_set_contact_surface(self.h, 0, None))
            return
        t = np.ascontiguousarray(tris, np.int32)
        if t.ndim != 2 or t.shape[1] != 3:
            raise _lib.TxError("tris must have shape (n, 3)")
        self._check(self.lib.tx_fem_set_contact_surface(self.h, len(t), t.ctypes.data))

    @staticmethod
    def decode_stats(st: torch.Tensor) -> list[dict]:
        raw = st.cpu().numpy().tobytes()
        n = st.shape[0]
        arr = (_lib.TxFemStats * n).from_buffer_copy(raw)
        return [{k: getattr(s, k) for k, _ in _lib.TxFemStats._fields_} for s in arr]

    # -- FEM marker read-out (ManiSkill-ViTac style) ---------------------------------------------------------------------
    def set_markers(self, tri: np.ndarray, weights: np.ndarray, cam_R=None, cam_t=(0.0, 0.0, 0.0285), intrinsics=(340.0, 325.0, 160.0, 125.0),
                    reference_tail: bool = False, num_markers: int = 128, img_hw=(240, 320), normalize: bool = False):
        """``reference_tail``: ``tri`` / ``weights`` hold the UNPADDED markers that lie on the surface; apply the rest of
        ``gen_marker_flow`` -- uv mask on the initial positions (Q12), compaction, padding to ``num_markers`` -- once here
        (it only depends on the rest positions) and ``normalize`` in the kernel."""
        tri = np.ascontiguousarray(tri, np.int32)
        w = np.ascontiguousarray(weights, np.float64)
        zero_all = False
        if reference_tail:
            P = (np.asarray(self.mesh.X, np.float64)[tri] * w[..., None]).sum(1)
            sel = reference_marker_tail(project_uv(P, cam_R, cam_t, intrinsics), num_markers, img_hw)
            if sel.size == 0:
                zero_all = True
                sel = np.zeros(num_markers, np.int64)
            tri, w = np.ascontiguousarray(tri[sel]), np.ascontiguousarray(w[sel])
        self._check(self.lib.tx_fem_set_marker_output(self.h, int(bool(normalize)), float(img_hw[1]), int(zero_all)))
        R = np.ascontiguousarray(np.diag([1.0, -1.0, -1.0]) if cam_R is None else cam_R, np.float64)
        t = np.ascontiguousarray(cam_t, np.float64)
        fx, fy, cx, cy = intrinsics
        self._check(self.lib.tx_fem_set_markers(self.h, len(tri), tri.ctypes.data, w.ctypes.data, R.ctypes.data, t.ctypes.data,
                                                fx, fy, cx, cy))
        self.M = len(tri)

    # -- Isaac x UIPC attachment, per-step aim positions (ref: uipc_attachments.py:388-428) -----------------------------------
    def attachment_aim(self, pose: torch.Tensor, offsets: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """aim (N, A, 3) float64 = R(quat) offsets + pos for every env: ``pose`` (N, 7) float32 (xyz + quaternion wxyz of the body
        the gel hangs on), ``offsets`` (A, 3) or (N, A, 3) float32 body-frame attachment points (:func:`compute_attachment_data`)."""
        for t in (pose, offsets):
            if t.dtype != torch.float32 or t.device != self.device or not t.is_contiguous():
                raise _lib.TxError("pose / offsets must be contiguous float32 tensors on the engine's device")
        N = pose.shape[0]
        per_env = offsets.dim() == 3
        if tuple(pose.shape) != (N, 7) or tuple(offsets.shape[-2:]) != (self.A, 3) or (per_env and offsets.shape[0] != N):
            raise _lib.TxError("pose must be (N, 7), offsets (A, 3) or (N, A, 3) with A = the mesh's attached vertices")
        out = torch.empty((N, self.A, 3), dtype=torch.float64, device=self.device) if out is None else out
        self._chk_state(aim=out)
        self._check(self.lib.tx_fem_attachment_aim(self.h, pose.data_ptr(), offsets.data_ptr(), N, int(per_env), out.data_ptr()))
        return out

    # -- gel surface -> height map (the reference's TODO at gelsight_sensor.py:594-598) ---------------------------
    def height_map(self, x: torch.Tensor, out: torch.Tensor | None = None, shape=(240, 320), pitch_m: float = 0.0295e-3 * 640 / 320,
                   origin_xy=(0.0, 0.0), cam_z_m: float = -0.024, far_mm: float = 29.0) -> torch.Tensor:
        """(N, H, W) float32 height map [mm] rasterised from the deformed top surface of every gel: the distance from the camera
        plane to the surface along the optical axis, clipped to the far plane. The camera looks up through the pad from
        ``cam_z_m`` = -24 mm (``gelpad_to_camera_min_distance`` below the pad's bottom, GelSight Mini preset): the undeformed
        surface reads 28.5 mm, a 1 mm indentation 27.5 mm -- the values the depth camera reports for the indenter there.
        Feeds ``TactileEngine.render`` directly."""
        self._chk_state(x=x)
        if not getattr(self, "_surface_set", False):
            tt = np.ascontiguousarray(self.mesh.top_tris, np.int32)
            self._check(self.lib.tx_fem_set_surface(self.h, len(tt), tt.ctypes.data))
            self._surface_set = True
        N, (H, W) = x.shape[0], shape
        out = torch.empty((N, H, W), device=self.device) if out is None else out
        if out.dtype != torch.float32 or out.device != self.device or not out.is_contiguous() or tuple(out.shape) != (N, H, W):
            raise _lib.TxError("height map output must be a contiguous float32 (N, H, W) tensor on the engine's device")
        self._check(self.lib.tx_fem_heightmap(self.h, x.data_ptr(), N, out.data_ptr(), H, W, float(pitch_m), float(origin_xy[0]),
                                              float(origin_xy[1]), float(cam_z_m), float(far_mm)))
        return out

    def markers(self, x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        self._chk_state(x=x)
        N = x.shape[0]
        out = torch.empty((N, 2, self.M, 2), device=self.device) if out is None else out
        self._check(self.lib.tx_fem_markers(self.h, x.data_ptr(), N, out.data_ptr()))
        return out


def reference_marker_grid(interval_mm: float = 2.0625, translation_mm=(0.0, 0.0), rotation_rad: float = 0.0) -> np.ndarray:
    """Marker positions [m], (K, 2), exactly as ``_gen_marker_grid`` lays them out in the CAMERA frame when all random ranges are
    zero (ref: tactile_sensor_sapienipc_modified.py:189-247; pinned against the executed function in
    tests/test_host_cpu.py): x from -ceil(8 / d) d to +ceil(16.5 / d) d, y from -ceil(6 / d) d to +ceil(6 / d) d -- the x range is
    NOT symmetric (13 x 7 = 91 points from -8.25 mm to 16.5 mm for d = 2.0625 mm); markers outside the gel surface are dropped later
    (``_gen_marker_weight``). Pass the result, shifted into the pad frame, to :func:`marker_grid_weights` as ``points_xy``."""
    import math

    d, (tx, ty) = float(interval_mm), translation_mm
    x0 = -math.ceil((8 + tx) / d) * d + tx
    x1 = math.ceil((16.5 - tx) / d) * d + tx
    y0 = -math.ceil((6 + ty) / d) * d + ty
    y1 = math.ceil((6 - ty) / d) * d + ty
    mx = np.linspace(x0, x1, round((x1 - x0) / d) + 1, True)
    my = np.linspace(y0, y1, round((y1 - y0) / d) + 1, True)
    xy = np.array(np.meshgrid(mx, my)).reshape((2, -1)).T
    rot = np.array([[math.cos(rotation_rad), -math.sin(rotation_rad)], [math.sin(rotation_rad), math.cos(rotation_rad)]])
    return (xy @ rot.T) / 1000.0


def marker_grid_weights(mesh: GelMesh, pitch=2.0625e-3, rows=7, cols=13, pad_to=128, points_xy: np.ndarray | None = None):
    """Marker grid on the gel's top surface and its barycentric weights (init time, host).

    Restates ``_gen_marker_grid`` / ``_gen_marker_weight`` of the reference's FEM marker sensor
    (tactile_sensor_sapienipc_modified.py:189-329) for the default configuration (all random ranges zero): a
    ``cols x rows`` grid with ``pitch`` spacing centred on the pad, markers outside the surface are dropped, the list
    is padded to ``pad_to`` by repeating the last marker (mani_skill_sim_cfg.py:17,54).

    Placement: the reference lays the grid out in the CAMERA frame (:func:`reference_marker_grid`: x in [-8.25, 16.5] mm) and keeps
    what falls on the gel surface; where the camera frame sits relative to the pad is fixed by the sensor's USD asset, which
    cannot be read here. The default is therefore the full 13 x 7 grid centred on the pad; ``points_xy`` (K, 2) [m, pad frame]
    overrides it, e.g. ``reference_marker_grid() - camera_origin_in_pad_frame``."""
    X = mesh.X
    tris = mesh.top_tris
    if points_xy is not None:
        pts = np.asarray(points_xy, np.float64).reshape(-1, 2)
    else:
        xs = (np.arange(cols) - (cols - 1) / 2) * pitch
        ys = (np.arange(rows) - (rows - 1) / 2) * pitch
        pts = np.array([[x, y] for x in xs for y in ys])
    out_tri, out_w = [], []
    A, B, Cc = X[tris[:, 0], :2], X[tris[:, 1], :2], X[tris[:, 2], :2]
    for p in pts:
        v0, v1, v2 = B - A, Cc - A, p - A
        d00 = (v0 * v0).sum(1); d01 = (v0 * v1).sum(1); d11 = (v1 * v1).sum(1)
        d20 = (v2 * v0).sum(1); d21 = (v2 * v1).sum(1)
        den = d00 * d11 - d01 * d01
        b1 = (d11 * d20 - d01 * d21) / den
        b2 = (d00 * d21 - d01 * d20) / den
        b0 = 1 - b1 - b2
        inside = np.where((b0 >= -1e-12) & (b1 >= -1e-12) & (b2 >= -1e-12))[0]
        if inside.size:
            k = inside[0]
            out_tri.append(tris[k]); out_w.append([b0[k], b1[k], b2[k]])
    while len(out_tri) < pad_to:
        out_tri.append(out_tri[-1]); out_w.append(out_w[-1])
    return np.asarray(out_tri[:pad_to], np.int32), np.asarray(out_w[:pad_to], np.float64)


def quat_rotate_inverse_f32(q_wxyz: np.ndarray, v: np.ndarray) -> np.ndarray:
    """Isaac Lab's ``quat_apply_inverse`` in float32 (ref call: uipc_attachments.py:322-325): v - 2 w (q x v) + 2 q x (q x v)."""
    q = np.asarray(q_wxyz, np.float32)
    v = np.asarray(v, np.float32)
    w, xyz = q[0], q[1:]
    t = np.cross(xyz, v).astype(np.float32) * np.float32(2.0)
    return (v - w * t + np.cross(xyz, t).astype(np.float32)).astype(np.float32)


def compute_attachment_data(tet_points: np.ndarray, body_pos, body_quat_wxyz, body_half_extents, sphere_radius: float = 5e-4,
                            max_dist: float = 1e-5):
    """Which gel vertices hang on the rigid body (the sensor case) and where they sit in its frame -- the init-time half of
    ``UipcIsaacAttachments`` (ref: uipc_attachments.py:247-350). The reference asks PhysX whether a sphere of ``sphere_radius``
    swept by ``max_dist`` around each vertex touches the body's collider; headless the collider is the body's BOX (centre
    ``body_pos``, orientation ``body_quat_wxyz``, ``body_half_extents``): a vertex is attached iff its distance to the box is at
    most ``sphere_radius + max_dist``. Returns (offsets (A, 3) float32 in the body frame = quat_apply_inverse(q, v - pos) as the
    reference computes them, idx (A,) int32)."""
    P = np.asarray(tet_points, np.float64)
    pos = np.asarray(body_pos, np.float64)
    q = np.asarray(body_quat_wxyz, np.float64)
    w, x, y, z = q / np.linalg.norm(q)
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    local = (P - pos) @ R  # R^T (p - pos)
    d = np.abs(local) - np.asarray(body_half_extents, np.float64)
    dist = np.linalg.norm(np.maximum(d, 0.0), axis=1) + np.minimum(d.max(1), 0.0)
    idx = np.nonzero(dist <= sphere_radius + max_dist)[0].astype(np.int32)
    offs = np.stack([quat_rotate_inverse_f32(np.asarray(body_quat_wxyz, np.float32), (P[i] - pos).astype(np.float32)) for i in idx]) \
        if idx.size else np.zeros((0, 3), np.float32)
    return offs.astype(np.float32), idx


def project_uv(P_world: np.ndarray, cam_R=None, cam_t=(0.0, 0.0, 0.0285), intrinsics=(340.0, 325.0, 160.0, 125.0)) -> np.ndarray:
    """Pinhole projection of world points (K, 3) exactly as ``fem_marker_kernel`` does it (float64, rounded to float32)."""
    R = np.diag([1.0, -1.0, -1.0]) if cam_R is None else np.asarray(cam_R, np.float64)
    pc = (np.asarray(P_world, np.float64) - np.asarray(cam_t, np.float64)) @ R
    fx, fy, cx, cy = intrinsics
    return np.stack([fx * pc[:, 0] / pc[:, 2] + cx, fy * pc[:, 1] / pc[:, 2] + cy], -1).astype(np.float32)


def reference_marker_tail(init_uv: np.ndarray, num_markers: int = 128, img_hw=(240, 320)) -> np.ndarray:
    """Which markers ``gen_marker_flow`` returns, in which order (ref: tactile_sensor_sapienipc_modified.py:382-402, with the
    preset's zero lose-tracking probability / noise): the uv mask on the INITIAL positions -- ``5 < u < H`` and ``5 < v < W``:
    u is compared with the image HEIGHT and v with the WIDTH, the reference's axes are swapped, reproduced
    --, compaction in order, then padding to ``num_markers`` by repeating the last survivor. Returns the (num_markers,) index
    array into the unmasked marker list; no survivor -> an empty array (the reference itself raises there, :405; the product returns
    an all-zero flow). More survivors than
    ``num_markers`` makes the reference draw a RANDOM subset (np.random.choice): not reproducible, raises."""
    H, W = img_hw
    uv = np.asarray(init_uv)
    keep = np.nonzero((uv[:, 0] > 5) & (uv[:, 0] < H) & (uv[:, 1] > 5) & (uv[:, 1] < W))[0]
    if keep.size == 0:
        return keep
    if keep.size >= num_markers:
        if keep.size == num_markers:
            raise ValueError("the reference returns a random permutation when exactly num_markers markers survive")
        raise ValueError("more surviving markers than num_markers: the reference draws a random subset")
    return np.concatenate([keep, np.full(num_markers - keep.size, keep[-1])])


class GelPadSim:
    """Batched stand-in for ``UipcSim`` + the gel ``UipcObject`` of every sensor (ref: uipc_sim.py:250-252 ``step`` =
    ``world.advance(); world.retrieve()``; objects/uipc_object_deformable_data.py:135-147 ``surf_nodal_pos_w``): owns the
    float64 state of N gels on the device and advances all of them with one kernel launch per step. The reference supports
    exactly one environment here ("UIPC based envs can only be run with --num_envs=1", docs/source/showcases/ball_rolling.md:23)."""

    def __init__(self, num_envs: int, mesh: GelMesh | None = None, cfg: GelFemCfg | None = None, device="cuda"):
        from .gel_mesh import box_gel

        self.mesh = mesh or box_gel()
        self.engine = GelFemEngine(self.mesh, cfg, device)
        self.num_envs = num_envs
        self.x, self.v, self.x_prev = self.engine.new_state(num_envs)
        self.aim = self.engine.rest_aim(num_envs)
        self._ind_prev = None
        self.last_stats = None

    @property
    def nodal_pos_w(self) -> torch.Tensor:
        return self.x

    @property
    def surf_nodal_pos_w(self) -> torch.Tensor:
        idx = torch.from_numpy(np.unique(self.mesh.top_tris).astype(np.int64)).to(self.x.device)
        return self.x[:, idx]

    def reset(self, env_ids=None):
        ids = slice(None) if env_ids is None else env_ids
        self.x[ids] = self.engine.X
        self.x_prev[ids] = self.engine.X
        self.v[ids] = 0

    def set_attachment_aim(self, aim: torch.Tensor):
        """Target positions of the attached (bottom-layer) vertices, e.g. transformed by the sensor-case pose
        (ref: uipc_attachments.py:364-428)."""
        self.aim.copy_(aim)

    def step(self, indenter: torch.Tensor, want_stats: bool = True):
        """``indenter``: packed ``tx_fem_indenter`` array of the poses at the END of this step (see indenter_array)."""
        prev = indenter if self._ind_prev is None else self._ind_prev
        self.last_stats = self.engine.step(self.x, self.v, self.x_prev, self.aim, prev, indenter, want_stats)
        self._ind_prev = indenter
        return self.last_stats
