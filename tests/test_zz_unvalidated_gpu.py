"""GPU tests that have NOT run on a device yet (written after the round's GPU budget was spent). Everything here is marked xfail
(non-strict): an XPASS in the round-end log is the confirmation, a failure cannot turn the validated suite red, and the file name
sorts last so that a fault in an unvalidated path cannot disturb the validated tests of the same pytest process.

* shadow branch (tx_render_shadow, csrc/taxim_shadow_kernel.cu) bitwise against the canonical restatement, which is itself pinned
  to the executed reference (tests/test_shadow_cpu.py); the plug-in does not route with_shadow=True there yet;
* the fused (validated) kernel on extreme contacts added late (tests/test_refbox_cpu.py::EDGE_CASES)."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN, H, W

pytestmark = [pytest.mark.gpu, pytest.mark.xfail(reason="shadow kernels not yet validated on a GPU (written after the GPU budget of round 1 was spent)", strict=False)]


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


def test_extreme_contacts_bitwise_vs_canonical(tables, canon_taxim):
    """The fused (validated) kernel on the extreme contacts of tests/test_refbox_cpu.py::EDGE_CASES -- inputs added after the
    round's GPU budget was spent, hence in this xfail-marked file; expected to XPASS."""
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
