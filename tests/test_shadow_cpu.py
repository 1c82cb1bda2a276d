"""Taxim SHADOW branch (`with_shadow=True`, ref: taxim_torch.py:260-346): the canonical CPU restatement
(oracle/taxim_canon.c::canon_taxim_render_shadow) against the EXECUTED reference.

* committed fixtures (tests/golden/shadow_sub.npz, oracle/make_golden_shadow.py): the intermediate scatter-min shadow image
  captured from inside the reference's run -- which pixels a shadow sample lands on and with which value: the whole shadow-casting
  logic (mask dilation, attachment boundary, direction / height indices incl. the reference's last-height quirk, ray fan,
  coordinate truncation, depth test, scatter-min) -- is reproduced BIT FOR BIT; the final RGB (min, blur, + background, blur,
  clip) agrees to 1e-5 wherever the 7 x 7 neighbourhood a pixel's two blurs reach holds no pixel whose gradient bin the
  reference's FFT noise decided differently;
* container only (marker refbox): the init-time tables against the reference's tensors, and fresh inputs.
The device kernels (tx_render_shadow) are compared bit for bit with this restatement in tests/test_shadow_gpu.py."""
import numpy as np
import pytest
import torch
from scipy.ndimage import binary_erosion

from conftest import GOLDEN, H, W


@pytest.fixture(scope="module")
def shadow_tables():
    from tacex_b200.calib import ShadowTables

    return ShadowTables.load(GOLDEN / "gsmini_shadow_tables_320x240.npz")


def _check(canon_taxim, st, hm, press, rgb_ref, im_ref, id_ref, sh_ref):
    from oracle import canon

    o = canon.render_shadow(canon_taxim, st, hm, press, want_boundary=True, want_shadow_img=True)
    sh = o["shadow_img"]
    assert np.array_equal(np.isfinite(sh), np.isfinite(sh_ref)), "a shadow sample landed on a different pixel"
    fin = np.isfinite(sh_ref)
    assert fin.sum() > 1000 and np.array_equal(sh[fin], sh_ref[fin])
    base = canon_taxim.render(hm, press, want=("idx",))
    agree = (base["idx_mag"] == im_ref) & (base["idx_dir"] == id_ref)
    core = np.stack([binary_erosion(a, structure=np.ones((7, 7)), border_value=1) for a in agree])
    d = np.abs(o["rgb"] - rgb_ref).max(-1)
    touched = fin.any(1)
    assert (touched & core).sum() >= 0.9 * touched.sum()  # nearly every shadowed pixel is comparable
    assert d[core].max() <= 1e-5
    return o


def test_shadow_vs_reference_golden(canon_taxim, shadow_tables):
    from oracle import make_golden_shadow as mg

    g = dict(np.load(GOLDEN / "shadow_sub.npz"))
    hm = mg.inputs().numpy()
    assert abs(float(hm.astype(np.float64).sum()) - float(g["input_sum"])) < 1e-6 * abs(float(g["input_sum"]))
    n = hm.shape[0]
    sh_ref = np.full(n * 3 * H * W, np.inf, np.float32)
    sh_ref[g["shadow_idx"]] = g["shadow_val"]
    o = _check(canon_taxim, shadow_tables, hm, g["press"], g["rgb"], g["idx_mag"], g["idx_dir"], sh_ref.reshape(n, 3, H, W))
    assert (o["boundary"].sum((1, 2)) > 0).all()  # every fixture frame is in contact: it has a shadow attachment area


def test_no_contact_frame_has_no_shadow(canon_taxim, shadow_tables):
    from oracle import canon

    hm = np.full((1, H, W), 29.0, np.float32)
    o = canon.render_shadow(canon_taxim, shadow_tables, hm, np.zeros(1, np.float32), want_boundary=True, want_shadow_img=True)
    assert o["boundary"].sum() == 0 and not np.isfinite(o["shadow_img"]).any()


@pytest.mark.refbox
def test_shadow_tables_and_fresh_inputs_vs_executed_reference(canon_taxim, shadow_tables):
    from oracle import ref_bootstrap as rb

    if not rb.available():
        pytest.skip("reference checkout not present on this machine")
    from tacex_b200 import synth
    from tacex_b200.calib import ShadowTables

    torch.set_num_threads(1)
    tx = rb.load_taxim()
    st = ShadowTables.from_calib_folder(rb.CALIB_DIR, (H, W))
    assert np.array_equal(st.table, tx._TaximTorch__shadow_table_padded.numpy())
    fan = tx._TaximTorch__shadow_direction_fan_angles
    assert np.array_equal(st.fan_cos, torch.cos(fan).numpy()) and np.array_equal(st.fan_sin, torch.sin(fan).numpy())
    assert np.array_equal(st.table, shadow_tables.table) and st.dilate_rounds == shadow_tables.dilate_rounds
    hm = synth.height_map_mm(synth.config2(6, seed=404)["depth_m"])
    press = rb.ref_indentation_depth(hm)
    rgb = rb.ref_render_shadow(tx, hm, press).numpy()
    sh = rb.SCATTER_CAPTURE["shadow_img_flat"].reshape(3, hm.shape[0], H, W).transpose(0, 1).contiguous().numpy()
    dg, _ = rb.ref_deformed_gel(tx, hm, press)
    _, _, im, idr = rb.ref_normals_bins(tx, dg)
    _check(canon_taxim, st, hm.numpy(), press.numpy(), rgb, im.numpy(), idr.numpy(), sh)


def test_shadow_kernel_source_emulated_on_the_host_matches_the_restatement(tables, canon_taxim, shadow_tables):
    """No GPU has run csrc/taxim_shadow_kernel.cu yet. Its kernels are plain per-pixel code, so the SAME SOURCE FILE is compiled
    for the host (tools/emu: a stand-in <cuda_runtime.h> maps every intrinsic to the identical IEEE-754 operation) and driven
    thread by thread on the deformed gel + mask of the canonical restatement (with which the fused GPU kernel is bit-identical).
    The result must equal canon_taxim_render_shadow bit for bit. This checks the kernels' logic (indexing, dilation, ray casting,
    integer-atomic float minimum, blur order), not the GPU execution -- that is tests/test_zz_unvalidated_gpu.py."""
    import ctypes as C
    import subprocess
    from pathlib import Path

    from oracle import canon
    from oracle import make_golden_shadow as mg
    from tacex_b200 import synth

    root = Path(__file__).resolve().parent.parent
    so = root / "tools" / "emu" / "_build" / "libshadow_emu.so"
    srcs = [root / "tools/emu/shadow_emu.cpp", root / "tools/emu/include/cuda_runtime.h", root / "tacex_b200/csrc/taxim_shadow_kernel.cu",
            root / "tacex_b200/csrc/tx_common.cuh", root / "tacex_b200/csrc/tx_kernels.h"]
    if not so.exists() or so.stat().st_mtime < max(p.stat().st_mtime for p in srcs):
        so.parent.mkdir(exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-mfma", "-fPIC", "-shared", "-w", f"-I{root / 'tools/emu/include'}",
                        "-o", str(so), str(srcs[0])], check=True, capture_output=True)
    emu = C.CDLL(str(so))
    st = shadow_tables
    hm = torch.cat([mg.inputs()[:2], synth.golden_config1(H, W)[3:5]]).numpy()  # config 0, a config-2 env, a sphere, an env without contact
    n = hm.shape[0]
    press = canon_taxim.indentation_depth(hm)
    base = canon_taxim.render(hm, press, want=("deformed", "mask"))
    ref = canon.render_shadow(canon_taxim, st, hm, press)
    # device layouts of tx_upload_tables: poly [nb][nb][3 * 6 padded to 20], background [H][W][3]
    nb = tables.params.num_bins
    poly20 = np.zeros((nb, nb, 20), np.float32)
    poly20[:, :, :18] = tables.poly_grad.numpy().transpose(1, 2, 0, 3).reshape(nb, nb, 18)
    bg_hwc = np.ascontiguousarray(tables.background.numpy().transpose(1, 2, 0))
    taps = tables.params.blur_taps((H, W))
    tfx, tfy = (np.ascontiguousarray(t.numpy(), np.float32) for t in taps[-1])
    tsx, tsy = (np.ascontiguousarray(t, np.float32) for t in st.blur_taps)
    dil = np.array([st.dilate_rounds[0][0], st.dilate_rounds[0][1], st.dilate_rounds[1][0], st.dilate_rounds[1][1]], np.int32)
    tab = np.ascontiguousarray(st.table, np.float32)
    fcos, fsin = np.ascontiguousarray(st.fan_cos, np.float32), np.ascontiguousarray(st.fan_sin, np.float32)
    deformed = np.ascontiguousarray(base["deformed"], np.float32)
    mask = np.ascontiguousarray(base["mask"], np.uint8)
    rgb = np.empty((n, H, W, 3), np.float32)
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))  # noqa: E731
    f32 = C.c_float
    emu.emu_shadow(fp(deformed), mask.ctypes.data_as(C.POINTER(C.c_uint8)), None, fp(poly20), fp(bg_hwc), fp(tab), fp(fcos), fp(fsin),
                   tab.shape[1], tab.shape[2], tab.shape[3], fcos.shape[1], nb, dil.ctypes.data_as(C.POINTER(C.c_int)),
                   f32(tables.params.pixmm), f32(tables.params.calib_h), f32(tables.params.calib_w), f32(st.depth_0),
                   f32(st.height_precision), f32(st.discretize_precision), f32(st.step_x), f32(st.step_y), fp(tsx), tsx.size, fp(tsy),
                   tsy.size, fp(tfx), tfx.size, fp(tfy), tfy.size, n, fp(rgb))
    assert np.array_equal(rgb, ref), f"max |d| = {np.abs(rgb - ref).max()}"
