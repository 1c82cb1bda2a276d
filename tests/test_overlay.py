"""Marker image / marker overlay (ref: fots_marker_sim.py:346-384, ball_rolling_taxim_fots.py:918-937).
CPU: the restatement (oracle/canon.py) against the fixture oracle/make_golden_overlay.py produced by executing the reference's
``generate_patch_array`` + ``draw_markers`` + the task's overlay arithmetic. GPU: tx_marker_overlay, bit for bit."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN, H, W


@pytest.fixture(scope="module")
def g():
    return dict(np.load(GOLDEN / "marker_overlay.npz"))


def _rgb(seed):
    return torch.rand((2, H, W, 3), generator=torch.Generator().manual_seed(int(seed))).numpy()


def test_marker_image_restatement_equals_the_executed_reference(g):
    from oracle import canon

    for k in range(g["markers"].shape[0]):
        img = canon.marker_image(g["markers"][k], g["patch_w15"])
        assert np.array_equal(img, g["marker_images"][k]), f"marker set {k}"
    assert (g["marker_images"] < 255).any(axis=(1, 2)).all()


def test_overlay_arithmetic_equals_the_executed_reference(g):
    from oracle import canon

    rgb = _rgb(g["rgb_seed"])
    r0, r1 = g["overlay_rows"]
    for i in range(2):
        o = canon.marker_overlay(rgb[i], g["marker_images"][i + 2])
        assert np.array_equal(o[r0:r1], g["overlay"][i])


@pytest.mark.gpu
def test_marker_overlay_kernel_bitwise(tables, g):
    from oracle import canon
    from tacex_b200.engine import TactileEngine

    mk = g["markers"]
    K, M = mk.shape[0], mk.shape[1]
    eng = TactileEngine(tables, max_envs=K)
    eng.set_marker_patches(g["patch_w15"])
    md = torch.zeros((K, 2, M, 2))
    md[:, 1] = torch.from_numpy(mk)
    md = md.cuda()
    rgb = torch.rand((K, H, W, 3), generator=torch.Generator().manual_seed(9))
    rgbd = rgb.cuda()
    img = torch.empty((K, H, W), dtype=torch.uint8, device="cuda")
    out = torch.empty_like(rgbd)
    u8 = torch.empty((K, H, W, 3), dtype=torch.uint8, device="cuda")
    k0 = eng.counters()["kernels_launched"]
    eng.marker_overlay(md, rgbd, apply=True, rgb_out=out, marker_img_out=img, rgb_u8_out=u8)
    torch.cuda.synchronize()
    assert eng.counters()["kernels_launched"] == k0 + 1
    assert np.array_equal(img.cpu().numpy(), g["marker_images"])
    ref = np.stack([canon.marker_overlay(rgb[k].numpy(), g["marker_images"][k]) for k in range(K)])
    assert np.array_equal(out.cpu().numpy(), ref)
    assert np.array_equal(u8.cpu().numpy(), canon.rgb_to_u8(ref))
    # in place, and the plain uint8 conversion without markers
    eng.marker_overlay(md, rgbd, apply=True, rgb_out=rgbd)
    eng.marker_overlay(None, out, apply=False, rgb_u8_out=u8)
    torch.cuda.synchronize()
    assert torch.equal(rgbd, out)
    assert np.array_equal(u8.cpu().numpy(), canon.rgb_to_u8(ref))


@pytest.mark.gpu
def test_plugin_overlay_replaces_the_tasks_per_env_loop(tables, g):
    """B200FOTSMarkerSimulator.overlay_markers == the RL task's loop (draw_markers per env + the overlay arithmetic)."""
    from oracle import canon
    from tacex_b200 import sensor, synth

    n = 4
    cfg = sensor.gelsight_mini_cfg(str(GOLDEN / "gsmini_tables_320x240.npz"), num_envs=n)
    s = sensor.GelSightSensor(cfg)
    s.set_camera_depth(synth.config2(n, seed=1)["depth_m"].cuda())
    s.update(0.0, force_recompute=True)
    rgb = s.data.output["tactile_rgb"].clone()
    md = s.data.output["marker_motion"].clone()
    sim = s.marker_motion_simulator
    imgs = sim.draw_markers_batch(md).cpu().numpy()
    out = sim.overlay_markers(rgb.clone(), md)
    torch.cuda.synchronize()
    for i in range(n):
        ref_img = canon.marker_image(md[i, 1].cpu().numpy(), g["patch_w15"])
        assert np.array_equal(imgs[i], ref_img)
        assert np.array_equal(out[i].cpu().numpy(), canon.marker_overlay(rgb[i].cpu().numpy(), ref_img))
