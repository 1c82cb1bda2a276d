"""FEM gel surface -> sensor height map (the reference's TODO at gelsight_sensor.py:594-598):
tx_fem_heightmap against a NumPy restatement of the same rasterisation (bit for bit), its properties, and the coupled path
FEM step -> height map -> Taxim RGB against the oracle on the rasterised map."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
H, W = 240, 320
PITCH = 0.0295e-3 * 640 / 320


def raster_numpy(X, tris, cam_z=-0.024, far=29.0, H=H, W=W, pitch=PITCH):
    """Same arithmetic, same order as raster_kernel.cu (float64, no contraction); minimum over the covering triangles."""
    hm = np.full((H, W), np.float32(far))
    for t in tris:
        a, b, c = X[t[0]], X[t[1]], X[t[2]]
        x_lo, x_hi = min(a[0], b[0], c[0]), max(a[0], b[0], c[0])
        y_lo, y_hi = min(a[1], b[1], c[1]), max(a[1], b[1], c[1])
        c0, c1 = int(np.floor(x_lo / pitch + W / 2.0 - 0.5)), int(np.ceil(x_hi / pitch + W / 2.0 - 0.5))
        r0, r1 = int(np.floor(y_lo / pitch + H / 2.0 - 0.5)), int(np.ceil(y_hi / pitch + H / 2.0 - 0.5))
        c0, r0, c1, r1 = max(c0, 0), max(r0, 0), min(c1, W - 1), min(r1, H - 1)
        if c1 < c0 or r1 < r0:
            continue
        v0x, v0y, v1x, v1y = b[0] - a[0], b[1] - a[1], c[0] - a[0], c[1] - a[1]
        d00, d01, d11 = v0x * v0x + v0y * v0y, v0x * v1x + v0y * v1y, v1x * v1x + v1y * v1y
        den = d00 * d11 - d01 * d01
        if not abs(den) > 0:
            continue
        cols = np.arange(c0, c1 + 1, dtype=np.float64)
        rows = np.arange(r0, r1 + 1, dtype=np.float64)
        px = (cols - W / 2.0 + 0.5) * pitch
        py = (rows - H / 2.0 + 0.5) * pitch
        v2x, v2y = (px - a[0])[None, :], (py - a[1])[:, None]
        d20 = v2x * v0x + v2y * v0y
        d21 = v2x * v1x + v2y * v1y
        b1 = (d11 * d20 - d01 * d21) / den
        b2 = (d00 * d21 - d01 * d20) / den
        b0 = 1.0 - b1 - b2
        inside = (b0 >= -1e-9) & (b1 >= -1e-9) & (b2 >= -1e-9)
        z = b0 * a[2] + b1 * b[2] + b2 * c[2]
        d = np.clip(((z - cam_z) * 1000.0).astype(np.float32), np.float32(0), np.float32(far))
        sub = hm[r0:r1 + 1, c0:c1 + 1]
        sub[inside] = np.minimum(sub[inside], d[inside])
    return hm


def test_rasteriser_bitwise_vs_numpy_and_properties():
    from tacex_b200 import fem, gel_mesh

    m = gel_mesh.box_gel()
    eng = fem.GelFemEngine(m)
    x, v, xp = eng.new_state(3)
    rng = np.random.default_rng(2)
    X = np.asarray(m.X, np.float64)
    Xd = X.copy()
    Xd[:, 2] -= 8e-4 * np.exp(-((X[:, 0] - 2e-3) ** 2 + (X[:, 1] + 1e-3) ** 2) / (2 * (3e-3) ** 2)) * (X[:, 2] / 4.5e-3)  # a dent
    Xd[:, :2] += 5e-5 * rng.standard_normal((X.shape[0], 2))
    x[1] = torch.from_numpy(Xd).cuda()
    x[2, :, 2] -= 1e-3  # rigid 1 mm towards the camera
    hm = eng.height_map(x).cpu().numpy()
    torch.cuda.synchronize()
    tris = np.asarray(m.top_tris)
    # the pad (20.75 x 25.25 mm) covers the whole 18.9 x 14.2 mm image: undeformed = 28.5 mm everywhere
    assert np.abs(hm[0] - 28.5).max() < 1e-4 and np.array_equal(hm[0], raster_numpy(X, tris))
    assert np.array_equal(hm[1], raster_numpy(Xd, tris))
    assert 28.5 - hm[1].min() > 0.7 and np.abs(hm[2] - 27.5).max() < 1e-4
    # a coarse camera (the FEM preset's 32 x 24) and an off-centre window
    lo = eng.height_map(x, shape=(24, 32), pitch_m=PITCH * 10).cpu().numpy()
    assert np.array_equal(lo[1], raster_numpy(Xd, tris, H=24, W=32, pitch=PITCH * 10))


def test_fem_step_to_height_map_to_tactile_rgb(tables=None):
    """The coupled path the reference leaves as a TODO: gel FEM substep (sphere press) -> rasterised height map -> fused Taxim
    kernel. RGB bit-exact vs the canonical oracle on the SAME rasterised map; the indentation depth the optical model derives
    from it equals the surface's largest displacement."""
    from conftest import GOLDEN
    from oracle import canon
    from tacex_b200 import fem, gel_mesh
    from tacex_b200.calib import TaximTables
    from tacex_b200.engine import TactileEngine

    t = TaximTables.load(GOLDEN / "gsmini_tables_320x240.npz")
    m = gel_mesh.box_gel()
    sim = fem.GelPadSim(2, m)
    r = 3e-3
    for s in range(1, 13):
        sim.step(fem.indenter_array(0, [[0.0, 0.0, 4.5e-3 + r + 4e-4 - 1e-3 * s / 12], [2e-3, -1e-3, 4.5e-3 + r + 4e-4 - 1e-3 * s / 12]], (r, 0, 0)))
    hm = sim.engine.height_map(sim.x)
    eng = TactileEngine(t, max_envs=2)
    depth = torch.empty(2, device="cuda")
    rgb = eng.render(hm, None, depth_out=depth)
    torch.cuda.synchronize()
    cn = canon.CanonTaxim(H, W, t.poly_grad.numpy(), t.background.numpy(), None, t.params.blur_taps((H, W)))
    hmn = hm.cpu().numpy()
    pc = cn.indentation_depth(hmn)
    o = cn.render(hmn, pc, want=("rgb",))
    assert np.array_equal(depth.cpu().numpy(), pc) and np.array_equal(rgb.cpu().numpy(), o["rgb"])
    top = np.unique(np.asarray(m.top_tris))
    dz = (4.5e-3 - sim.x[:, top, 2].amin(1)).cpu().numpy() * 1000
    assert np.all(pc > 0.3) and np.abs(pc - dz).max() < 0.05  # mm: the press the optical model sees = the dent of the surface
