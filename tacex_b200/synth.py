"""Synthetic ("recorded") contact depth maps that stand in for Isaac Sim's RTX depth camera.

The reference obtains the height map from the sensor camera's depth image of the indenter
(ref: source/tacex/tacex/gelsight_sensor.py:581-593). That upstream producer is out of scope; the
benchmark and the parity tests replace it with the analytic indenters of SURVEY.md section 8(d):

  config 0: one sphere (r = 3 mm) pressed 1 mm, centred
  config 1: spheres, random centre / press, 10 % of the envs without contact (seed 0)
  config 2: sphere / flat cylinder / 90 deg wedge / 60 deg cone, random pose + in-plane yaw (seed 1),
            plus a FOTS trajectory (first-contact pose, then shear and twist)

All generators are deterministic (``torch.Generator`` on the CPU) so the GPU box reproduces the inputs
of the committed golden fixtures without access to the reference checkout.
"""

from __future__ import annotations

import math

import torch

# GelSight Mini geometry (ref: source/tacex_assets/tacex_assets/sensors/gelsight_mini/gsmini_cfg.py:22-32,51-52)
CLIP_MIN_M = 0.024
CLIP_MAX_M = 0.029
GELPAD_HEIGHT_M = 0.0045
GEL_SURFACE_M = CLIP_MIN_M + GELPAD_HEIGHT_M  # 0.0285 m from the camera
PIXEL_PITCH_M_320 = 0.0295e-3 * 640 / 320  # 0.059 mm per pixel at 320 x 240


def _grid(H: int, W: int, pitch: float) -> tuple[torch.Tensor, torch.Tensor]:
    ys = (torch.arange(H, dtype=torch.float64) - H / 2 + 0.5) * pitch
    xs = (torch.arange(W, dtype=torch.float64) - W / 2 + 0.5) * pitch
    Y, X = torch.meshgrid(ys, xs, indexing="ij")
    return X, Y


def indenter_profile(kind: int, X: torch.Tensor, Y: torch.Tensor, size: float) -> torch.Tensor:
    """Height of the indenter surface above its lowest point [m]; +inf outside its footprint.

    kind 0 sphere (radius ``size``), 1 flat cylinder (radius ``size``), 2 wedge (90 deg edge along y, half
    width ``size``), 3 cone (60 deg apex, base radius ``size``).
    """
    rho2 = X * X + Y * Y
    inf = torch.full_like(X, float("inf"))
    if kind == 0:
        z = size - torch.sqrt(torch.clamp(size * size - rho2, min=0.0))
        return torch.where(rho2 < size * size, z, inf)
    if kind == 1:
        return torch.where(rho2 < size * size, torch.zeros_like(X), inf)
    if kind == 2:
        z = X.abs()  # 90 deg edge: slope 1 on both flanks
        return torch.where((X.abs() < size) & (Y.abs() < 2.0 * size), z, inf)
    if kind == 3:
        z = torch.sqrt(rho2) / math.tan(math.radians(30.0))  # 60 deg full apex angle
        return torch.where(rho2 < size * size, z, inf)
    raise ValueError(f"unknown indenter kind {kind}")


def indenter_mesh(kind: int, size: float, segments: int = 24):
    """Closed triangle mesh (n, 3, 3) float64 [m] of the config-2 indenters for the gel FEM path (``tx_fem_set_indenter_mesh``), in the
    indenter's own frame: lowest point at the origin, body towards +z -- the same shapes as ``indenter_profile``: kind 1 flat
    cylinder (radius ``size``, height ``size``), 2 wedge (90 deg edge along y, half width ``size``, half length 2 ``size``),
    3 cone (60 deg apex, base radius ``size``); kind 0: the sphere of radius ``size`` as an icosphere (1280 triangles by default)."""
    import numpy as np

    tris = []
    ang = np.linspace(0.0, 2.0 * np.pi, segments + 1)
    ring = lambda r, z: np.stack([r * np.cos(ang), r * np.sin(ang), np.full_like(ang, z)], 1)  # noqa: E731
    if kind == 1:
        lo, hi, c0, c1 = ring(size, 0.0), ring(size, size), np.zeros(3), np.array([0.0, 0.0, size])
        for k in range(segments):
            tris += [[c0, lo[k + 1], lo[k]], [lo[k], lo[k + 1], hi[k + 1]], [lo[k], hi[k + 1], hi[k]], [c1, hi[k], hi[k + 1]]]
    elif kind == 2:
        s, l = size, 2.0 * size
        e0, e1 = np.array([0.0, -l, 0.0]), np.array([0.0, l, 0.0])
        a0, a1, b0, b1 = np.array([-s, -l, s]), np.array([-s, l, s]), np.array([s, -l, s]), np.array([s, l, s])
        tris += [[e0, e1, a1], [e0, a1, a0], [e0, b0, b1], [e0, b1, e1], [e0, a0, b0], [e1, b1, a1], [a0, a1, b1], [a0, b1, b0]]
    elif kind == 3:
        h = size / math.tan(math.radians(30.0))
        hi, apex, c1 = ring(size, h), np.zeros(3), np.array([0.0, 0.0, h])
        for k in range(segments):
            tris += [[apex, hi[k + 1], hi[k]], [c1, hi[k], hi[k + 1]]]
    elif kind == 0:
        # sphere of radius ``size`` (lowest point at the origin) as an icosphere: ``segments`` // 8 subdivisions of the icosahedron
        # (24 -> 3 subdivisions: 1280 triangles, 642 vertices, 1920 edges) -- a "recorded triangle soup" of realistic size
        t = (1.0 + math.sqrt(5.0)) / 2.0
        v = [np.array(p, float) / math.sqrt(1 + t * t) for p in
             [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]]
        f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
             (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
        cache: dict = {}

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m_ = v[a] + v[b]
                v.append(m_ / np.linalg.norm(m_))
                cache[key] = len(v) - 1
            return cache[key]

        for _ in range(max(segments // 8, 0)):
            f = [g for a, b, c in f for g in ((a, mid(a, b), mid(c, a)), (b, mid(b, c), mid(a, b)), (c, mid(c, a), mid(b, c)), (mid(a, b), mid(b, c), mid(c, a)))]
        off = np.array([0.0, 0.0, size])
        tris = [[size * v[a] + off, size * v[b] + off, size * v[c] + off] for a, b, c in f]
    else:
        raise ValueError(f"no mesh for indenter kind {kind}")
    return np.asarray(tris, np.float64)


def depth_map(
    kind: int,
    size_m: float,
    cx_m: float,
    cy_m: float,
    yaw: float,
    press_m: float,
    H: int = 240,
    W: int = 320,
    pitch: float = PIXEL_PITCH_M_320,
    contact: bool = True,
) -> torch.Tensor:
    """Camera depth image [m] (float32, H x W) of one indenter pressed ``press_m`` into the gel."""
    if not contact:
        return torch.full((H, W), CLIP_MAX_M, dtype=torch.float32)
    X, Y = _grid(H, W, pitch)
    Xc, Yc = X - cx_m, Y - cy_m
    c, s = math.cos(yaw), math.sin(yaw)
    Xr = c * Xc + s * Yc
    Yr = -s * Xc + c * Yc
    z = indenter_profile(kind, Xr, Yr, size_m)
    d = torch.clamp(GEL_SURFACE_M - press_m + z, max=CLIP_MAX_M)
    return d.to(torch.float32)


def height_map_mm(depth_m: torch.Tensor, clip_max_m: float = CLIP_MAX_M) -> torch.Tensor:
    """GelSightSensor._get_height_map (ref: gelsight_sensor.py:581-593): inf -> clip max, metres -> mm."""
    hm = depth_m.clone()
    hm[torch.isinf(hm)] = clip_max_m
    hm *= 1000
    return hm


def config0(H: int = 240, W: int = 320) -> dict:
    d = depth_map(0, 3e-3, 0.0, 0.0, 0.0, 1e-3, H, W)[None]
    return {"depth_m": d, "kind": torch.zeros(1, dtype=torch.int64)}


def config1(n_envs: int = 256, seed: int = 0, H: int = 240, W: int = 320) -> dict:
    g = torch.Generator().manual_seed(seed)
    cx = (torch.rand(n_envs, generator=g, dtype=torch.float64) * 8 - 4) * 1e-3
    cy = (torch.rand(n_envs, generator=g, dtype=torch.float64) * 6 - 3) * 1e-3
    p = (0.2 + torch.rand(n_envs, generator=g, dtype=torch.float64) * 1.3) * 1e-3
    nocontact = torch.rand(n_envs, generator=g, dtype=torch.float64) < 0.10
    d = torch.stack(
        [
            depth_map(0, 3e-3, cx[i].item(), cy[i].item(), 0.0, p[i].item(), H, W, contact=not bool(nocontact[i]))
            for i in range(n_envs)
        ]
    )
    return {"depth_m": d, "kind": torch.zeros(n_envs, dtype=torch.int64), "contact": ~nocontact}


def config2(n_envs: int = 1024, seed: int = 1, H: int = 240, W: int = 320) -> dict:
    """Random primitive indenters + a two-sample FOTS trajectory per env.

    Returns depth maps for the FIRST contact sample (``depth_m0``, yaw ``theta0``) and for the CURRENT sample
    (``depth_m``, yaw ``theta``): the indenter is sheared by U(-0.5, 0.5) mm and twisted by U(-30, 30) deg.
    """
    g = torch.Generator().manual_seed(seed)
    r = lambda: torch.rand(n_envs, generator=g, dtype=torch.float64)  # noqa: E731
    kind = torch.randint(0, 4, (n_envs,), generator=g)
    u_size = r()
    cx = (r() * 8 - 4) * 1e-3
    cy = (r() * 6 - 3) * 1e-3
    p = (0.2 + r() * 1.3) * 1e-3
    yaw0 = (r() * 2 - 1) * math.pi
    sx = (r() - 0.5) * 1e-3
    sy = (r() - 0.5) * 1e-3
    dth = (r() * 2 - 1) * math.radians(30.0)
    nocontact = r() < 0.10
    size = torch.empty(n_envs, dtype=torch.float64)
    for i in range(n_envs):
        k = int(kind[i])
        lo, hi = {0: (1.5, 4.0), 1: (1.0, 3.0), 2: (1.0, 3.0), 3: (1.5, 4.0)}[k]
        size[i] = (lo + (hi - lo) * u_size[i]) * 1e-3
    d0, d1 = [], []
    for i in range(n_envs):
        c = not bool(nocontact[i])
        d0.append(depth_map(int(kind[i]), size[i].item(), cx[i].item(), cy[i].item(), yaw0[i].item(), p[i].item(), H, W, contact=c))
        d1.append(
            depth_map(
                int(kind[i]), size[i].item(), (cx[i] + sx[i]).item(), (cy[i] + sy[i]).item(), (yaw0[i] + dth[i]).item(),
                p[i].item(), H, W, contact=c,
            )
        )
    return {
        "depth_m0": torch.stack(d0),
        "depth_m": torch.stack(d1),
        "theta0": yaw0.to(torch.float32),
        "theta": (yaw0 + dth).to(torch.float32),
        "kind": kind,
        "contact": ~nocontact,
    }


def golden_config1(H: int = 240, W: int = 320) -> torch.Tensor:
    """Height maps [mm] of the committed ``config1_sub`` fixture: 4 config-1 spheres + 1 env without contact."""
    d = config1(4, seed=0, H=H, W=W)["depth_m"]
    d = torch.cat([d, depth_map(0, 3e-3, 0.0, 0.0, 0.0, 0.0, H, W, contact=False)[None]])
    return height_map_mm(d)


def golden_config2(H: int = 240, W: int = 320) -> dict:
    """Inputs of the committed ``config2_sub`` fixture (8 envs, two trajectory samples)."""
    c2 = config2(8, seed=1, H=H, W=W)
    return {"hm0": height_map_mm(c2["depth_m0"]), "hm1": height_map_mm(c2["depth_m"]), "theta0": c2["theta0"],
            "theta": c2["theta"], "kind": c2["kind"]}


def bench_batch(n_envs: int, seed: int = 0, H: int = 240, W: int = 320, n_unique: int = 64) -> torch.Tensor:
    """Height maps [mm] for the benchmark: ``n_unique`` config-1 style sphere presses tiled to ``n_envs``.

    Generating thousands of analytic maps on the host is slow; the benchmark tiles a pool of unique maps (every
    env still runs the full path -- there is no caching anywhere in the engine)."""
    pool = height_map_mm(config1(min(n_unique, n_envs), seed, H, W)["depth_m"])
    reps = (n_envs + pool.shape[0] - 1) // pool.shape[0]
    return pool.repeat(reps, 1, 1)[:n_envs].contiguous()


def lowres_batch(n_envs: int, H: int, W: int, seed: int = 3) -> torch.Tensor:
    """Height maps [mm] at a coarse tactile resolution (the reference's RL tasks render 32 x 24 / 32 x 32 tactile images,
    ref: tacex_tasks/.../ball_rolling_taxim_fots.py:306-321): the four indenter kinds with random pose, the pixel pitch scaled
    so that the frame covers the same gel area as 320 x 240; every fifth env has no contact."""
    g = torch.Generator().manual_seed(seed)
    r = lambda: torch.rand(n_envs, generator=g, dtype=torch.float64)  # noqa: E731
    kind = torch.randint(0, 4, (n_envs,), generator=g)
    size = (1.5 + 2.5 * r()) * 1e-3
    cx, cy = (r() * 8 - 4) * 1e-3, (r() * 6 - 3) * 1e-3
    p = (0.2 + r() * 1.3) * 1e-3
    yaw = (r() * 2 - 1) * math.pi
    pitch = PIXEL_PITCH_M_320 * 320 / W
    d = torch.stack([
        depth_map(int(kind[i]), size[i].item(), cx[i].item(), cy[i].item(), yaw[i].item(), p[i].item(), H, W, pitch,
                  contact=(i % 5 != 4))
        for i in range(n_envs)
    ])
    return height_map_mm(d)


def box_depth_map(half_m, cx_m: float, cy_m: float, R, press_m: float, H: int = 240, W: int = 320,
                  pitch: float = PIXEL_PITCH_M_320) -> torch.Tensor:
    """Camera depth image [m] of a rigid BOX (half extents ``half_m``, rotation ``R`` 3x3, world z up towards the camera...
    the camera looks at the box's lower surface) whose lowest point is pressed ``press_m`` into the gel: per pixel the entry
    point of the vertical ray into the rotated box (slab test), relative to the box's lowest point."""
    X, Y = _grid(H, W, pitch)
    R = torch.as_tensor(R, dtype=torch.float64)
    h = torch.as_tensor(half_m, dtype=torch.float64)
    # ray o + t * ez in the box frame: o' = R^T (o - c), d' = R^T ez = third row of R
    ox, oy = X - cx_m, Y - cy_m
    t_in = torch.full_like(X, -float("inf"))
    t_out = torch.full_like(X, float("inf"))
    for k in range(3):
        o_k = R[0, k] * ox + R[1, k] * oy  # (R^T o)_k with o_z = 0
        d_k = R[2, k]
        if abs(float(d_k)) < 1e-12:
            inside = o_k.abs() <= h[k]
            t_in = torch.where(inside, t_in, torch.full_like(X, float("inf")))
            continue
        ta, tb = (-h[k] - o_k) / d_k, (h[k] - o_k) / d_k
        t_in = torch.maximum(t_in, torch.minimum(ta, tb))
        t_out = torch.minimum(t_out, torch.maximum(ta, tb))
    hit = t_in <= t_out
    low = -float((R[2, :].abs() * h).sum())  # z of the lowest corner relative to the centre
    z = torch.where(hit, t_in - low, torch.full_like(X, float("inf")))
    return torch.clamp(GEL_SURFACE_M - press_m + z, max=CLIP_MAX_M).to(torch.float32)


def _rot(yaw: float, tilt_y: float = 0.0, tilt_x: float = 0.0) -> torch.Tensor:
    c, s = math.cos(yaw), math.sin(yaw)
    Rz = torch.tensor([[c, -s, 0], [s, c, 0], [0, 0, 1.0]], dtype=torch.float64)
    c, s = math.cos(tilt_y), math.sin(tilt_y)
    Ry = torch.tensor([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=torch.float64)
    c, s = math.cos(tilt_x), math.sin(tilt_x)
    Rx = torch.tensor([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=torch.float64)
    return Rz @ Ry @ Rx


def config3_box(n_envs: int, seed: int = 2, step: int = 30, H: int = 240, W: int = 320) -> dict:
    """Depth maps of BASELINE config 3: a rigid 4 x 6 x 2 mm box pressed with an EDGE (even envs) or a CORNER (odd envs) into the
    gel, 0 -> 1 mm over 30 steps (``step`` selects the sample), seed 2 for the pose jitter."""
    g = torch.Generator().manual_seed(seed)
    r = lambda: torch.rand(n_envs, generator=g, dtype=torch.float64)  # noqa: E731
    cx, cy = (r() * 2 - 1) * 6e-3, (r() * 2 - 1) * 4e-3
    yaw = (r() * 2 - 1) * math.pi
    jit = (r() * 2 - 1) * math.radians(5.0)
    half = (2e-3, 3e-3, 1e-3)
    d, Rs = [], []
    for i in range(n_envs):
        R = _rot(yaw[i].item(), math.radians(20.0) + jit[i].item(), 0.0 if i % 2 == 0 else math.radians(15.0))
        Rs.append(R)
        d.append(box_depth_map(half, cx[i].item(), cy[i].item(), R, 1e-3 * step / 30, H, W))
    return {"depth_m": torch.stack(d), "R": torch.stack(Rs), "cx": cx, "cy": cy, "half": half}


def dense_batch(n_envs: int, seed: int = 4, H: int = 240, W: int = 320, n_unique: int = 16) -> torch.Tensor:
    """Height maps [mm] of DENSE contacts (the unfavourable case for the kernel's exact-zero skipping): flat punches covering
    about a third of the frame (10 x 8 mm boxes, any yaw), large spheres (r = 12 mm, 1.2-1.5 mm press), and config-3 edge /
    corner presses of a large box; no env without contact. ``n_unique`` maps tiled to ``n_envs``."""
    g = torch.Generator().manual_seed(seed)
    k = min(n_unique, n_envs)
    u = torch.rand((k, 5), generator=g, dtype=torch.float64)
    d = []
    for i in range(k):
        cx, cy = (u[i, 0].item() * 2 - 1) * 2e-3, (u[i, 1].item() * 2 - 1) * 1.5e-3
        yaw = (u[i, 2].item() * 2 - 1) * math.pi
        p = (1.2 + 0.3 * u[i, 3].item()) * 1e-3
        if i % 3 == 0:
            d.append(box_depth_map((5e-3, 4e-3, 1e-3), cx, cy, _rot(yaw), p, H, W))
        elif i % 3 == 1:
            d.append(depth_map(0, 12e-3, cx, cy, 0.0, p, H, W))
        else:
            d.append(box_depth_map((6e-3, 5e-3, 2e-3), cx, cy, _rot(yaw, math.radians(4.0), math.radians(3.0 * (i % 2))), p, H, W))
    pool = height_map_mm(torch.stack(d))
    reps = (n_envs + k - 1) // k
    return pool.repeat(reps, 1, 1)[:n_envs].contiguous()
