"""Row a13 tail (ref: tactile_sensor_sapienipc_modified.py:354-413 ``gen_marker_flow``): uv mask on the initial positions with the
reference's swapped axes (Q12), compaction, padding to num_markers, ``normalize``.
* container only (refbox): the reference method EXECUTED from its file (ast; nothing copied) on a stand-in ``self`` -- Isaac Lab's
  ``project_points`` (third party, absent) is replaced by its published formula (K p, divided by z), the debug-draw calls by no-ops,
  the hard-coded ``device="cuda:0"`` by the CPU -- against the host code of tacex_b200/fem.py;
* GPU: fem_marker_kernel with the tail applied, against the same host arithmetic."""
import types

import numpy as np
import pytest
import torch


def _surface_and_markers(shift_xy):
    from tacex_b200 import gel_mesh
    from tacex_b200.fem import marker_grid_weights, reference_marker_grid

    m = gel_mesh.box_gel()
    X = np.asarray(m.X, np.float64)
    pts_pad = reference_marker_grid() - np.asarray(shift_xy)
    half = np.array([20.75e-3, 25.25e-3]) / 2
    on = (np.abs(pts_pad) <= half + 1e-12).all(1)
    tri, w = marker_grid_weights(m, pad_to=int(on.sum()), points_xy=pts_pad[on])
    return m, X, tri, w


def _camera_points(P_world, cam_t=(0.0, 0.0, 0.0285)):
    return (np.asarray(P_world, np.float64) - np.asarray(cam_t)) @ np.diag([1.0, -1.0, -1.0])


@pytest.mark.refbox
@pytest.mark.parametrize("normalize", [False, True])
@pytest.mark.parametrize("intr", [(340.0, 325.0, 160.0, 125.0), (340.0, 325.0, 60.0, 125.0), (340.0, 325.0, 160.0, -400.0)])
def test_marker_flow_tail_against_the_executed_reference_method(normalize, intr):
    import ast
    from pathlib import Path

    src = Path("/root/reference/source/tacex/tacex/simulation_approaches/fem_based/sim/tactile_sensor_sapienipc_modified.py")
    if not src.exists():
        pytest.skip("reference checkout not present on this machine")
    from tacex_b200.fem import project_uv, reference_marker_tail

    m, X, tri, w = _surface_and_markers((4.125e-3, 0.0))
    rng = np.random.default_rng(3)
    Xd = X + 2e-4 * rng.standard_normal(X.shape)
    tree = ast.parse(src.read_text())
    fns = [n for c in tree.body if isinstance(c, ast.ClassDef) for n in c.body
           if isinstance(n, ast.FunctionDef) and n.name in ("gen_marker_flow", "gen_marker_uv")]
    K = np.array([[intr[0], 0, intr[2]], [0, intr[1], intr[3]], [0, 0, 1.0]], np.float32)

    def project_points(points, intrinsic):  # isaaclab.utils.math.project_points: (K p)^T, x and y divided by z
        p = points @ intrinsic.T
        return torch.cat([p[..., :2] / p[..., 2:3], p[..., 2:3]], -1)[None]

    tproxy = types.SimpleNamespace(tensor=lambda *a, device=None, **k: torch.tensor(*a, **k), float32=torch.float32)
    ns = {"np": np, "torch": tproxy, "math_utils": types.SimpleNamespace(project_points=project_points),
          "draw": types.SimpleNamespace(clear_points=lambda: None, draw_points=lambda *a: None)}
    exec(compile(ast.Module(body=fns, type_ignores=[]), str(src), "exec"), ns)
    cam_rest = torch.from_numpy(_camera_points(X).astype(np.float32))
    cam_cur = torch.from_numpy(_camera_points(Xd).astype(np.float32))
    me = types.SimpleNamespace(
        _gen_marker_grid=lambda: None, _gen_marker_weight=lambda grid: (tri.astype(np.int64), w),
        reference_surface_vertices_camera=cam_rest, get_surface_vertices_camera=lambda: cam_cur,
        transform_camera_to_world_frame=lambda t: t, camera_intrinsic=K, tactile_img_height=240, tactile_img_width=320,
        marker_lose_tracking_probability=0.0, marker_random_noise=0.0, num_markers=128, normalize=normalize)
    me.gen_marker_uv = lambda pts: ns["gen_marker_uv"](me, pts)
    P0 = (X[tri] * w[..., None]).sum(1)
    if reference_marker_tail(project_uv(P0, intrinsics=intr), 128, (240, 320)).size == 0:
        # no survivor: the reference itself fails (broadcast of an empty slice, :405); the product returns an all-zero flow
        with pytest.raises(ValueError):
            ns["gen_marker_flow"](me)
        return
    ref = ns["gen_marker_flow"](me).numpy()  # (2, 128, 2)
    # host code of the product
    P1 = (Xd[tri] * w[..., None]).sum(1)
    uv0, uv1 = project_uv(P0, intrinsics=intr), project_uv(P1, intrinsics=intr)
    sel = reference_marker_tail(uv0, 128, (240, 320))
    if sel.size == 0:
        mine = np.zeros((2, 128, 2))
    else:
        mine = np.stack([uv0[sel], uv1[sel]]).astype(np.float64)
    if normalize:
        mine = mine / 160.0 - 1.0
    assert ref.shape == mine.shape == (2, 128, 2)
    assert np.abs(ref - mine).max() <= (2e-3 if not normalize else 2e-5), np.abs(ref - mine).max()
    if intr[2] == 60.0:
        assert 0 < np.unique(sel).size < tri.shape[0]  # some markers really fell outside the mask


def test_tail_semantics_without_the_reference():
    from tacex_b200.fem import reference_marker_tail

    uv = np.array([[10, 10], [4, 50], [250, 50], [100, 330], [239, 319], [6, 6]], np.float32)
    sel = reference_marker_tail(uv, 8, (240, 320))
    assert sel.tolist() == [0, 4, 5, 5, 5, 5, 5, 5]  # u < H = 240 (!), v < W = 320 (!): Q12
    assert reference_marker_tail(np.array([[0.0, 0.0]]), 8).size == 0
    with pytest.raises(ValueError):
        reference_marker_tail(np.full((9, 2), 50.0), 8)


@pytest.mark.gpu
@pytest.mark.parametrize("normalize", [False, True])
def test_fem_marker_kernel_with_reference_tail(normalize):
    from tacex_b200 import fem
    from tacex_b200.fem import project_uv, reference_marker_tail

    intr = (340.0, 325.0, 60.0, 125.0)
    m, X, tri, w = _surface_and_markers((4.125e-3, 0.0))
    eng = fem.GelFemEngine(m)
    eng.set_markers(tri, w, intrinsics=intr, reference_tail=True, normalize=normalize)
    x, v, xp = eng.new_state(3)
    rng = np.random.default_rng(4)
    Xd = X + 2e-4 * rng.standard_normal(X.shape)
    x[1] = torch.from_numpy(Xd).cuda()
    mk = eng.markers(x).cpu().numpy()
    torch.cuda.synchronize()
    assert mk.shape == (3, 2, 128, 2)
    P0 = (X[tri] * w[..., None]).sum(1)
    uv0 = project_uv(P0, intrinsics=intr)
    sel = reference_marker_tail(uv0, 128)
    uv1 = project_uv((Xd[tri] * w[..., None]).sum(1), intrinsics=intr)
    ref = np.stack([uv0[sel], uv1[sel]]).astype(np.float64)
    if normalize:
        ref = ref / 160.0 - 1.0
    assert np.array_equal(mk[1], ref.astype(np.float32))
    assert np.array_equal(mk[0, 0], mk[0, 1]) and np.array_equal(mk[0], mk[2])
