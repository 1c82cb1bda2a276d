"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI / the plug-in classes, against the
canonical oracle (bit-exact) and against the committed golden vectors of the executed reference (protocol P1-P4)."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN, H, W, unpack

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(tables):
    from tacex_b200.engine import TactileEngine

    return TactileEngine(tables, max_envs=64, marker_rows=9, marker_cols=11)


def _run(eng, hm, press=None):
    hmd = hm.cuda().contiguous()
    n = hm.shape[0]
    dg = torch.empty((n, H, W), device="cuda")
    mk = torch.empty((n, H, W), device="cuda", dtype=torch.uint8)
    dep = torch.empty(n, device="cuda")
    rgb = eng.render(hmd, press, depth_out=dep, deformed_out=dg, mask_out=mk)
    torch.cuda.synchronize()
    return rgb.cpu().numpy(), dg.cpu().numpy(), mk.cpu().numpy(), dep.cpu().numpy()


def test_library_is_loaded_and_counts_launches(eng, inputs):
    c0 = eng.counters()
    _run(eng, inputs["config0"])
    c1 = eng.counters()
    assert c1["kernels_launched"] == c0["kernels_launched"] + 1 and c1["frames_rendered"] == c0["frames_rendered"] + 1
    with open("/proc/self/maps") as f:
        assert "libtacex_b200.so" in f.read()


@pytest.mark.parametrize("name", ["config0", "config1_sub", "config2_sub"])
def test_bitwise_vs_canonical_oracle(eng, canon_taxim, inputs, name):
    """P3: the kernel and the canonical CPU restatement agree bit for bit (deformed gel, mask, RGB, depth)."""
    hm = inputs[name]
    rgb, dg, mk, dep = _run(eng, hm)
    pc = canon_taxim.indentation_depth(hm.numpy())
    o = canon_taxim.render(hm.numpy(), pc)
    assert np.array_equal(dep, pc)
    assert np.array_equal(mk, o["mask"])
    assert np.array_equal(dg, o["deformed"])
    assert np.abs(rgb - o["rgb"]).max() <= 1e-6
    assert np.array_equal(rgb, o["rgb"])


@pytest.mark.parametrize("name", ["config0", "config1_sub", "config2_sub"])
def test_vs_reference_golden(eng, tables, inputs, golden, name):
    """P1 + P2 against the executed reference."""
    g = golden[name]
    hm = inputs[name]
    n = hm.shape[0]
    rgb, dg, mk, dep = _run(eng, hm)
    np.testing.assert_allclose(dep, g["press"], rtol=0, atol=1e-6)
    assert np.abs(dg - g["deformed"]).max() <= 1e-5
    assert (mk.astype(bool) != unpack(g["mask_bits"], n)).sum() <= 4
    # RGB of the reference where available; otherwise rebuild it from the reference's bins + tables
    if "rgb" in g:
        ref, sl = g["rgb"], slice(0, n)
    elif "rgb_first2" in g:
        ref, sl = g["rgb_first2"], slice(0, 2)
    else:
        ref = None
    well = unpack(g["well_bits"], n)
    if ref is not None:
        d = np.abs(rgb[sl] - ref).max(-1)
        assert (d[well[sl]] <= 1e-3).mean() >= 0.99
        print(f"{name}: raw RGB L_inf vs reference {d.max():.4f} (reference self-consistency floor 0.078)")


def test_region_edges_vs_canonical_oracle(eng, canon_taxim):
    """The kernel skips exact-zero rows / columns and maps the normal / colour stage onto the non-zero rectangle: contacts
    whose grown bounding box (radii 30+16+8+4+2+1+2 = 63 px) ends 0..3 px from an image border, crosses the CTA boundary
    (row 120) or touches the border exercise every clipping branch (incl. the replicate-padded border row / column)."""
    frames = []
    for top in (62, 63, 64, 65, 66, 67):            # grown region starts at row top - 63 in {-1..4}
        for size, left in ((9, 150), (40, 64), (3, 66)):
            hm = np.full((H, W), 29.0, np.float32)
            yy, xx = np.mgrid[0:size, 0:size]
            hm[top:top + size, left:left + size] = 27.5 + 0.02 * ((yy - size / 2) ** 2 + (xx - size / 2) ** 2) ** 0.5
            frames.append(hm)
            frames.append(hm[::-1, ::-1].copy())     # same distances from the bottom / right borders
    for top, left in ((0, 0), (110, 300), (119, 10), (230, 311), (100, 0)):
        hm = np.full((H, W), 29.0, np.float32)
        hm[top:top + 9, left:left + 9] = 27.9
        frames.append(hm)
    hm = torch.from_numpy(np.stack(frames))
    for i in range(0, hm.shape[0], 32):
        part = hm[i:i + 32]
        rgb, dg, mk, dep = _run(eng, part)
        pc = canon_taxim.indentation_depth(part.numpy())
        o = canon_taxim.render(part.numpy(), pc)
        assert np.array_equal(dep, pc)
        assert np.array_equal(mk, o["mask"])
        assert np.array_equal(dg, o["deformed"])
        assert np.array_equal(rgb, o["rgb"]), f"frames {np.unique(np.nonzero(rgb != o['rgb'])[0]) + i}"


@pytest.mark.parametrize("cam", [(24, 32), (32, 32), (60, 80)])
def test_camera_resolution_fused_resize(tables, canon_taxim, cam):
    """Row a3: sensor camera coarser than the tactile image (the FEM preset renders 32x24, the RL tasks 32x32). The resize
    runs in the kernel's load stage; bit-exact against the oracle (canonical resize -- itself bitwise torch F.resize on the
    CPU -- followed by the canonical render), with the indentation depth taken from the CAMERA-resolution map."""
    from oracle import canon
    from tacex_b200 import synth
    from tacex_b200.engine import TactileEngine

    hc, wc = cam
    pitch = synth.PIXEL_PITCH_M_320 * W / wc
    rng = np.random.default_rng(5)
    depth = torch.stack([
        synth.depth_map(0, 3e-3, float(rng.uniform(-4e-3, 4e-3)), float(rng.uniform(-3e-3, 3e-3)), 0.0, float(rng.uniform(2e-4, 1.5e-3)),
                        H=hc, W=wc, pitch=pitch, contact=i != 3) for i in range(6)])
    depth[5] = torch.where(depth[5] >= synth.CLIP_MAX_M, torch.full_like(depth[5], float("inf")), depth[5])  # RTX "no hit"
    hm_lo = synth.height_map_mm(depth)
    e = TactileEngine(tables, max_envs=8)
    e.set_camera_resolution(hc, wc)
    n = hm_lo.shape[0]
    dg = torch.empty((n, H, W), device="cuda")
    mk = torch.empty((n, H, W), device="cuda", dtype=torch.uint8)
    dep = torch.empty(n, device="cuda")
    rgb = e.render_camera(hm_lo.cuda().contiguous(), depth_out=dep, deformed_out=dg, mask_out=mk)
    dep2 = torch.empty(n, device="cuda")
    rgb2 = e.render_camera(depth.cuda().contiguous(), is_depth=True, clip_max_m=synth.CLIP_MAX_M, depth_out=dep2)
    torch.cuda.synchronize()
    pc = canon_taxim.indentation_depth(hm_lo.numpy())
    o = canon_taxim.render(canon.resize_bilinear(hm_lo.numpy(), (H, W)), pc)
    assert np.array_equal(dep.cpu().numpy(), pc) and np.array_equal(dep2.cpu().numpy(), pc)
    assert np.array_equal(mk.cpu().numpy(), o["mask"])
    assert np.array_equal(dg.cpu().numpy(), o["deformed"])
    assert np.array_equal(rgb.cpu().numpy(), o["rgb"])
    assert torch.equal(rgb, rgb2)
    assert (pc > 0).sum() >= 4


def test_explicit_press_equals_fused(eng, inputs):
    hm = inputs["config2_sub"].cuda()
    press = eng.indentation_depth(hm)
    a = eng.render(hm, press)
    b = eng.render(hm, None)
    torch.cuda.synchronize()
    assert torch.equal(a, b)


def test_batch_and_position_invariance(eng, inputs):
    """Appendix E: the same map gives bitwise identical output at any batch index / batch size."""
    hm = inputs["config2_sub"].cuda()
    a = eng.render(hm, None).clone()
    b = eng.render(hm[3:5].contiguous(), None)
    big = hm.repeat(6, 1, 1)[:41].contiguous()
    c = eng.render(big, None)
    torch.cuda.synchronize()
    assert torch.equal(a[3:5], b)
    assert torch.equal(c[8:16], a) and torch.equal(c[40], a[0])


def test_no_contact_and_translation_properties(eng, tables, inputs):
    from tacex_b200 import synth

    hm = inputs["config1_sub"][-1:].cuda()
    rgb, dg, mk, dep = _run(eng, hm.cpu())
    assert dep[0] == 0 and (dg == 0).all() and (mk == 0).all()
    # translation of the indenter by k px translates the deformed gel by k px away from the borders
    k = 17
    d0 = synth.depth_map(0, 2e-3, 0.0, 0.0, 0.0, 0.8e-3)
    d1 = synth.depth_map(0, 2e-3, k * synth.PIXEL_PITCH_M_320, 0.0, 0.0, 0.8e-3)
    hm2 = synth.height_map_mm(torch.stack([d0, d1]))
    _, dg2, _, _ = _run(eng, hm2)
    assert np.array_equal(dg2[0][:, 60:200], dg2[1][:, 60 + k:200 + k])


def test_press_monotonic(eng):
    from tacex_b200 import synth

    maps = torch.stack([synth.depth_map(0, 3e-3, 0, 0, 0, p * 1e-3) for p in (0.3, 0.6, 0.9, 1.2)])
    _, dg, _, dep = _run(eng, synth.height_map_mm(maps))
    mins = dg.reshape(4, -1).min(1)
    assert (np.diff(dep) > 0).all() and (np.diff(mins) < 0).all()


def test_full_size_batch_roundtrip_property(eng, tables):
    """BASELINE-size property check (256 envs, config 1): every env equals the single-env result of the same map."""
    from tacex_b200 import synth
    from tacex_b200.engine import TactileEngine

    big = TactileEngine(tables, max_envs=256)
    hm = synth.bench_batch(256, n_unique=8).cuda()
    out = big.render(hm, None)
    one = big.render(hm[:8].contiguous(), None)
    torch.cuda.synchronize()
    assert torch.equal(out.view(32, 8, H, W, 3), one.expand(32, 8, H, W, 3))
    assert float(out.min()) >= 0 and float(out.max()) <= 1


def test_fused_depth_preprocessing(eng, inputs):
    """tx_render_depth (raw camera depth in metres, inf = no hit) == _get_height_map in torch + tx_render, bit for bit."""
    from tacex_b200 import synth

    d = synth.config2(8, seed=1)["depth_m"].clone()
    d[d >= synth.CLIP_MAX_M] = float("inf")  # what the RTX depth camera returns where nothing is hit
    dd = d.cuda()
    hm_ref = synth.height_map_mm(d).cuda()
    dep_a, dep_b = torch.empty(8, device="cuda"), torch.empty(8, device="cuda")
    hm_out = torch.empty((8, H, W), device="cuda")
    a = eng.render_depth(dd, synth.CLIP_MAX_M, depth_out=dep_a, height_map_out=hm_out).clone()
    b = eng.render(hm_ref, None, depth_out=dep_b)
    torch.cuda.synchronize()
    assert torch.equal(hm_out, hm_ref) and torch.equal(dep_a, dep_b) and torch.equal(a, b)


def test_invalid_arguments_raise(eng, tables):
    from tacex_b200 import _lib

    with pytest.raises(_lib.TxError):
        eng.render(torch.zeros((1, 100, 100), device="cuda"))
    with pytest.raises(_lib.TxError):
        eng.render(torch.zeros((65, H, W), device="cuda"))  # exceeds max_envs
    with pytest.raises(_lib.TxError):
        eng.render(torch.zeros((1, H, W)))  # host tensor


@pytest.mark.parametrize("grid", [(9, 11), (7, 9)])
def test_fots_markers(tables, canon_taxim, inputs, golden, grid):
    """P4: markers vs the reference's MarkerMotion (<= 1e-4 m == 1.958 px) and vs the canonical restatement."""
    from oracle import canon
    from tacex_b200.engine import TactileEngine

    rows, cols = grid
    g = golden["config2_sub"]
    e = TactileEngine(tables, max_envs=8, marker_rows=rows, marker_cols=cols)
    cf = canon.CanonFots(H, W, rows, cols, 15, 26)
    assert np.array_equal(e.marker_x, cf.mx) and np.array_equal(e.marker_y, cf.my)
    n = 8
    traj0 = torch.zeros((n, 4), device="cuda")
    tlen = torch.zeros(n, device="cuda", dtype=torch.int32)
    for step, (hk, tk, pk) in enumerate((("config2_first", "theta0", "press0"), ("config2_sub", "theta", "press"))):
        hm = inputs[hk].cuda()
        dep = torch.empty(n, device="cuda")
        e.render(hm, None, depth_out=dep)
        mk = e.fots_markers(dep, inputs[tk].cuda(), traj0, tlen)
        torch.cuda.synchronize()
        o = canon_taxim.render(inputs[hk].numpy(), g[pk], want=("deformed", "mask"))
        mc = cf.step(o["deformed"], o["mask"], g[pk], inputs[tk].numpy())
        got = mk.cpu().numpy()
        assert np.abs(got - mc).max() <= 2e-4, "vs canonical restatement"
        d = np.abs(got - g[f"markers_{rows}x{cols}_step{step}"])
        assert d.max() <= 1.958 and d.max() <= 1e-3
    assert tlen.cpu().tolist() == [2] * n


def test_sensor_plugin_flow(tables, canon_taxim, inputs, golden):
    """Through the reference-facing plug-in: GelSightSensor stand-in -> B200TaximSimulator / B200FOTSMarkerSimulator."""
    from tacex_b200 import sensor, synth

    c2 = synth.golden_config2(H, W)
    cfg = sensor.gelsight_mini_cfg(str(GOLDEN / "gsmini_tables_320x240.npz"), num_envs=8)
    s = sensor.GelSightSensor(cfg)
    eng = s.optical_simulator.engine
    k0 = eng.counters()["kernels_launched"]
    d_first = synth.config2(8, seed=1)["depth_m0"]
    s._data.output["indenter_yaw"] = c2["theta0"].cuda()
    s.set_camera_depth(d_first.cuda())
    s.update(0.0, force_recompute=True)
    out = s.data.output
    k1 = eng.counters()["kernels_launched"]
    assert k1 - k0 == 2, "one fused Taxim launch + one FOTS launch per sensor update"
    g = golden["config2_sub"]
    np.testing.assert_allclose(s.indentation_depth.cpu().numpy(), g["press0"], atol=1e-6)
    o = canon_taxim.render(c2["hm0"].numpy(), canon_taxim.indentation_depth(c2["hm0"].numpy()))
    assert np.array_equal(out["tactile_rgb"].cpu().numpy(), o["rgb"])
    assert np.abs(out["marker_motion"].cpu().numpy() - g["markers_9x11_step0"]).max() <= 1e-3
    # second sample of the trajectory
    s._data.output["indenter_yaw"] = c2["theta"].cuda()
    s.set_camera_depth(synth.config2(8, seed=1)["depth_m"].cuda())
    s.update(0.0, force_recompute=True)
    assert np.abs(s.data.output["marker_motion"].cpu().numpy() - g["markers_9x11_step1"]).max() <= 1e-3
    # partial reset keeps the reference's semantics: depth zeroed for the env, trajectory cleared
    s.reset(torch.tensor([1], device="cuda"))
    assert float(s.indentation_depth[1]) == 0.0
    assert int(s.marker_motion_simulator.traj_len[1]) == 0 and int(s.marker_motion_simulator.traj_len[0]) == 3


def test_sensor_plugin_flow_coarse_camera(tables, canon_taxim):
    """The FEM preset's 32x24 sensor camera (ref: gsmini_taxim_fem_cfg.py:27,52) through the plug-in: the optical simulator
    resizes inside the fused kernel; RGB / indentation depth / markers equal the oracle's resize + render + FOTS."""
    from oracle import canon
    from tacex_b200 import sensor, synth

    cfg = sensor.gelsight_mini_cfg(str(GOLDEN / "gsmini_tables_320x240.npz"), num_envs=4)
    cfg.sensor_camera_cfg.resolution = (32, 24)
    s = sensor.GelSightSensor(cfg)
    pitch = synth.PIXEL_PITCH_M_320 * 10
    depth = torch.stack([synth.depth_map(0, 3e-3, 1e-3 * i, -5e-4 * i, 0.0, 4e-4 + 3e-4 * i, H=24, W=32, pitch=pitch) for i in range(4)])
    s.set_camera_depth(depth.cuda())
    s.update(0.0, force_recompute=True)
    hm_lo = synth.height_map_mm(depth).numpy()
    pc = canon_taxim.indentation_depth(hm_lo)
    o = canon_taxim.render(canon.resize_bilinear(hm_lo, (H, W)), pc)
    assert np.array_equal(s.indentation_depth.cpu().numpy(), pc) and (pc > 0).all()
    assert np.array_equal(s.data.output["tactile_rgb"].cpu().numpy(), o["rgb"])
    cf = canon.CanonFots(H, W, 9, 11, 15, 26)
    mc = cf.step(o["deformed"], o["mask"], pc, np.zeros(4, np.float32))
    assert np.abs(s.data.output["marker_motion"].cpu().numpy() - mc).max() <= 2e-4
    assert tuple(s.data.output["height_map"].shape) == (4, 24, 32)


def test_step_host_end_to_end(tables, canon_taxim, inputs):
    from tacex_b200.engine import TactileEngine

    e = TactileEngine(tables, max_envs=8, marker_rows=9, marker_cols=11)
    hm = inputs["config2_first"].pin_memory()
    rgb = torch.empty((8, H, W, 3)).pin_memory()
    dep = torch.empty(8).pin_memory()
    mk = torch.empty((8, 2, 99, 2)).pin_memory()
    e.step_host(hm, rgb, dep, inputs["theta0"].pin_memory(), mk)
    pc = canon_taxim.indentation_depth(hm.numpy())
    o = canon_taxim.render(hm.numpy(), pc)
    assert np.array_equal(dep.numpy(), pc) and np.array_equal(rgb.numpy(), o["rgb"])


def test_extreme_contacts_bitwise_vs_canonical(tables, canon_taxim):
    """The fused kernel on the extreme contacts of tests/test_refbox_cpu.py::EDGE_CASES (a third of the frame in contact, image
    corner, 0.01 mm and 4.4 mm presses, a few pixels of contact)."""
    from tacex_b200 import synth
    from tacex_b200.engine import TactileEngine
    from test_refbox_cpu import EDGE_CASES

    hm = synth.height_map_mm(torch.stack([synth.depth_map(*v) for v in EDGE_CASES.values()]))
    n = hm.shape[0]
    eng = TactileEngine(tables, max_envs=n, marker_rows=9, marker_cols=11)
    depth = torch.empty(n, device="cuda")
    deformed = torch.empty((n, H, W), device="cuda")
    mask = torch.empty((n, H, W), device="cuda", dtype=torch.uint8)
    rgb = eng.render(hm.cuda(), None, depth_out=depth, deformed_out=deformed, mask_out=mask)
    torch.cuda.synchronize()
    press = canon_taxim.indentation_depth(hm.numpy())
    o = canon_taxim.render(hm.numpy(), press)
    assert np.array_equal(depth.cpu().numpy(), press)
    assert np.array_equal(mask.cpu().numpy().astype(bool), o["mask"].astype(bool))
    assert np.array_equal(deformed.cpu().numpy(), o["deformed"])
    assert np.array_equal(rgb.cpu().numpy(), o["rgb"])


def _render_all(eng, hm, chunk=None):
    n = hm.shape[0]
    hmd = hm.cuda()
    rgb = torch.empty((n, H, W, 3), device="cuda")
    dg = torch.empty((n, H, W), device="cuda")
    mk = torch.empty((n, H, W), device="cuda", dtype=torch.uint8)
    dep = torch.empty(n, device="cuda")
    eng.render(hmd, None, out=rgb, depth_out=dep, deformed_out=dg, mask_out=mk)
    torch.cuda.synchronize()
    return rgb, dg, mk, dep


def test_config1_256_envs_bitwise_vs_canonical_oracle(tables, canon_taxim):
    """BASELINE config 1 at its full size (256 envs, seed 0): every frame array_equal to the canonical oracle."""
    from tacex_b200 import synth
    from tacex_b200.engine import TactileEngine

    hm = synth.height_map_mm(synth.config1(256, seed=0)["depth_m"])
    eng = TactileEngine(tables, max_envs=256)
    rgb, dg, mk, dep = _render_all(eng, hm)
    pc = canon_taxim.indentation_depth(hm.numpy())
    o = canon_taxim.render(hm.numpy(), pc)
    assert np.array_equal(dep.cpu().numpy(), pc)
    assert np.array_equal(mk.cpu().numpy(), o["mask"])
    assert np.array_equal(dg.cpu().numpy(), o["deformed"])
    assert np.array_equal(rgb.cpu().numpy(), o["rgb"])


def test_config2_1024_envs_bitwise_vs_canonical_oracle(tables, canon_taxim):
    """BASELINE config 2 at its full size (1024 envs, seed 1, all four indenter kinds, moving yaw): RGB + FOTS markers of both
    trajectory samples against the canonical oracle (RGB / deformed gel / mask bit-exact, markers <= 2e-4 px)."""
    from oracle import canon
    from tacex_b200 import synth
    from tacex_b200.engine import TactileEngine

    n = 1024
    c2 = synth.config2(n, seed=1)
    assert set(c2["kind"].tolist()) == {0, 1, 2, 3}
    eng = TactileEngine(tables, max_envs=n, marker_rows=7, marker_cols=9)
    cf = canon.CanonFots(H, W, 7, 9, 15, 26)
    traj0 = torch.zeros((n, 4), device="cuda")
    tl = torch.zeros(n, device="cuda", dtype=torch.int32)
    canon.use_all_threads()
    for key, th in (("depth_m0", "theta0"), ("depth_m", "theta")):
        hm = synth.height_map_mm(c2[key])
        rgb, dg, mk, dep = _render_all(eng, hm)
        mkr = eng.fots_markers(dep, c2[th].cuda(), traj0, tl)
        torch.cuda.synchronize()
        pc = canon_taxim.indentation_depth(hm.numpy())
        o = canon_taxim.render(hm.numpy(), pc)
        assert np.array_equal(dep.cpu().numpy(), pc)
        assert np.array_equal(mk.cpu().numpy(), o["mask"])
        assert np.array_equal(dg.cpu().numpy(), o["deformed"])
        assert np.array_equal(rgb.cpu().numpy(), o["rgb"])
        mc = cf.step(o["deformed"], o["mask"], pc, c2[th].numpy())
        assert np.abs(mkr.cpu().numpy() - mc).max() <= 2e-4


def test_4096_env_launch_strided_sample_vs_canonical_oracle(tables, canon_taxim):
    """One launch at the benchmark's batch size (4096 envs, the bench's own input pool incl. dense box contacts): a strided
    sample of 64 frames spread over the whole grid is array_equal to the canonical oracle, and every frame equals the frame
    of the same depth map elsewhere in the batch (batch-position invariance over all 4096)."""
    from tacex_b200 import synth
    from tacex_b200.engine import TactileEngine

    n = 4096
    pool = torch.cat([synth.bench_batch(48, seed=0, n_unique=48), synth.dense_batch(16, seed=4)])
    hm = pool.repeat(n // 64, 1, 1).contiguous()
    eng = TactileEngine(tables, max_envs=n)
    rgb = torch.empty((n, H, W, 3), device="cuda")
    dep = torch.empty(n, device="cuda")
    eng.render(hm.cuda(), None, out=rgb, depth_out=dep)
    torch.cuda.synchronize()
    idx = torch.arange(0, n, 65)[:64]  # 65 = 64 + 1: walks through the pool AND through the grid
    sub = hm[idx]
    pc = canon_taxim.indentation_depth(sub.numpy())
    o = canon_taxim.render(sub.numpy(), pc, want=("rgb",))
    assert np.array_equal(dep[idx.cuda()].cpu().numpy(), pc)
    assert np.array_equal(rgb[idx.cuda()].cpu().numpy(), o["rgb"])
    first = rgb[:64]
    for r in range(1, n // 64):
        assert torch.equal(rgb[64 * r: 64 * (r + 1)], first), f"replica {r} differs"


@pytest.mark.parametrize("cam", [(480, 640), (300, 400), (360, 480)])
def test_downsampling_resize_bitwise_torch_antialias(tables, canon_taxim, cam):
    """Sensor camera FINER than the tactile image (ref: taxim_sim.py:88-89 F.resize): tx_resize is bit-identical to torch's
    antialiased bilinear interpolate (= the canonical resize of the oracle), and the plug-in renders the resized map."""
    import torch.nn.functional as F

    from oracle import canon
    from tacex_b200 import sensor, synth
    from tacex_b200.engine import TactileEngine

    Hc, Wc = cam
    pitch = synth.PIXEL_PITCH_M_320 * 320 / Wc
    depth = torch.stack([synth.depth_map(k % 4, 2.5e-3, 1e-3 * k, -5e-4 * k, 0.3 * k, 4e-4 + 2e-4 * k, H=Hc, W=Wc, pitch=pitch) for k in range(3)])
    hm = synth.height_map_mm(depth)
    eng = TactileEngine(tables, max_envs=3)
    got = eng.resize(hm.cuda())
    torch.cuda.synchronize()
    ref_t = F.interpolate(hm[:, None], size=[H, W], mode="bilinear", align_corners=False, antialias=True)[:, 0]
    assert torch.equal(got.cpu(), ref_t)
    assert np.array_equal(got.cpu().numpy(), canon.resize_bilinear(hm.numpy(), (H, W)))
    # through the plug-in: camera 2x finer than the tactile image
    cfg = sensor.gelsight_mini_cfg(str(GOLDEN / "gsmini_tables_320x240.npz"), num_envs=3, with_markers=False)
    cfg.sensor_camera_cfg.resolution = (Wc, Hc)
    s = sensor.GelSightSensor(cfg)
    k0 = s.optical_simulator.engine.counters()["kernels_launched"]
    s.set_camera_depth(depth.cuda())
    s.update(0.0, force_recompute=True)
    hm_small = ref_t.numpy()
    pc = canon_taxim.indentation_depth(hm.numpy())  # the indentation depth belongs to the CAMERA-resolution map
    o = canon_taxim.render(hm_small, pc, want=("rgb",))
    assert np.array_equal(s.data.output["tactile_rgb"].cpu().numpy(), o["rgb"])
    assert s.optical_simulator.engine.counters()["kernels_launched"] - k0 >= 2  # resize kernel + fused render: no torch fallback
