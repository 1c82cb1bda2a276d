"""Taxim SHADOW branch (`with_shadow=True`, ref: taxim_torch.py:260-346): the canonical CPU restatement
(oracle/taxim_canon.c::canon_taxim_render_shadow) against the EXECUTED reference.

* committed fixtures (tests/golden/shadow_sub.npz, oracle/make_golden_shadow.py): the intermediate scatter-min shadow image
  captured from inside the reference's run -- which pixels a shadow sample lands on and with which value: the whole shadow-casting
  logic (mask dilation, attachment boundary, direction / height indices incl. the reference's last-height quirk, ray fan,
  coordinate truncation, depth test, scatter-min) -- is reproduced BIT FOR BIT; the final RGB (min, blur, + background, blur,
  clip) agrees to 1e-5 wherever the 7 x 7 neighbourhood a pixel's two blurs reach holds no pixel whose gradient bin the
  reference's FFT noise decided differently;
* container only (marker refbox): the init-time tables against the reference's tensors, and fresh inputs.
No product kernel consumes this yet: the plug-in still raises NotImplementedError for with_shadow=True."""
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
