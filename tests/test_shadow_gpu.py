"""GPU parity of the Taxim shadow branch (with_shadow=True; tx_render_shadow, csrc/taxim_shadow_kernel.cu): bitwise against
the canonical restatement, which is itself pinned to the executed reference (tests/test_shadow_cpu.py). Both tests of the
former xfail-marked file passed on the driver's B200 in round 1; they are ordinary (strict) tests now."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN, H, W

pytestmark = pytest.mark.gpu


def test_shadow_branch_bitwise_vs_canonical(tables, canon_taxim):
    from oracle import canon
    from oracle import make_golden_shadow as mg
    from tacex_b200 import synth
    from tacex_b200.calib import ShadowTables
    from tacex_b200.engine import TactileEngine

    st = ShadowTables.load(GOLDEN / "gsmini_shadow_tables_320x240.npz")
    hm = torch.cat([mg.inputs(), synth.golden_config1(H, W)])  # config 0, two config-2 envs, config-1 spheres + one env without contact
    n = hm.shape[0]
    eng = TactileEngine(tables, max_envs=n)
    eng.upload_shadow_tables(st)
    depth = torch.empty(n, device="cuda")
    rgb = eng.render_shadow(hm.cuda(), None, depth_out=depth)
    torch.cuda.synchronize()
    press = canon_taxim.indentation_depth(hm.numpy())
    ref = canon.render_shadow(canon_taxim, st, hm.numpy(), press)
    assert np.array_equal(depth.cpu().numpy(), press)
    got = rgb.cpu().numpy()
    assert np.array_equal(got, ref), f"max |d| = {np.abs(got - ref).max()}"
    # explicit press == fused; the plain render is untouched by the shadow state
    rgb2 = eng.render_shadow(hm.cuda(), torch.from_numpy(press).cuda())
    assert torch.equal(rgb2, rgb)
    plain = eng.render(hm.cuda(), None)
    assert np.array_equal(plain.cpu().numpy(), canon_taxim.render(hm.numpy(), press, want=("rgb",))["rgb"])


def test_shadow_batch_larger_than_the_scratch_chunk_keeps_fots_inputs_and_rects(tables, canon_taxim):
    """tx_render_shadow renders in chunks of 256 envs: the per-env FOTS records and the gather rectangles of EVERY chunk must
    land at their batch position (a following tx_fots_markers on the same batch is valid and equals the un-chunked path)."""
    from tacex_b200 import synth
    from tacex_b200.calib import ShadowTables
    from tacex_b200.engine import TactileEngine

    n = 300
    hm = synth.bench_batch(n, seed=5, n_unique=12).cuda()
    st = ShadowTables.load(GOLDEN / "gsmini_shadow_tables_320x240.npz")
    eng = TactileEngine(tables, max_envs=n, marker_rows=7, marker_cols=9)
    eng.upload_shadow_tables(st)
    theta = torch.zeros(n, device="cuda")
    rect_s = torch.full((n, 2, 4), -7, device="cuda", dtype=torch.int32)
    rect_p = torch.full((n, 2, 4), -7, device="cuda", dtype=torch.int32)
    depth = torch.empty(n, device="cuda")
    eng.set_rect_output(rect_s)
    rgb_s = eng.render_shadow(hm, None, depth_out=depth)
    mk_s = eng.fots_markers(depth, theta, torch.zeros((n, 4), device="cuda"), torch.zeros(n, device="cuda", dtype=torch.int32))
    eng.set_rect_output(rect_p)
    eng.render(hm, None, depth_out=depth)
    mk_p = eng.fots_markers(depth, theta, torch.zeros((n, 4), device="cuda"), torch.zeros(n, device="cuda", dtype=torch.int32))
    eng.set_rect_output(None)
    torch.cuda.synchronize()
    assert torch.equal(mk_s, mk_p)
    assert torch.equal(rect_s, rect_p) and not (rect_s == -7).any()
    # chunk boundary: env 256.. of the batch equals the same frames rendered as their own batch
    rgb_tail = eng.render_shadow(hm[256:].contiguous(), None)
    assert torch.equal(rgb_s[256:], rgb_tail)


def test_plugin_routes_with_shadow_to_the_shadow_kernels(tables, canon_taxim, tmp_path):
    """B200TaximSimulator(cfg.with_shadow=True) renders through tx_render_shadow by default (no opt-in switch)."""
    import shutil

    from oracle import canon
    from tacex_b200 import sensor, synth
    from tacex_b200.calib import ShadowTables

    shutil.copy(GOLDEN / "gsmini_tables_320x240.npz", tmp_path / "tables_320x240.npz")
    shutil.copy(GOLDEN / "gsmini_shadow_tables_320x240.npz", tmp_path / "shadow_tables_320x240.npz")
    st = ShadowTables.load(GOLDEN / "gsmini_shadow_tables_320x240.npz")
    d = torch.cat([synth.config1(4, seed=0)["depth_m"], synth.depth_map(0, 3e-3, 0.0, 0.0, 0.0, 0.0, contact=False)[None]])
    hm = synth.height_map_mm(d)
    n = hm.shape[0]
    cfg = sensor.gelsight_mini_cfg(str(tmp_path), num_envs=n, with_markers=False)
    cfg.optical_sim_cfg.with_shadow = True
    s = sensor.GelSightSensor(cfg)
    s.set_camera_depth(d.cuda())
    s.update(0.0, force_recompute=True)
    torch.cuda.synchronize()
    press = canon_taxim.indentation_depth(hm.numpy())
    ref = canon.render_shadow(canon_taxim, st, hm.numpy(), press)
    assert np.array_equal(s.data.output["tactile_rgb"].cpu().numpy(), ref)
