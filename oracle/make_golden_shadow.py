"""Generates the shadow-branch fixtures by EXECUTING the unmodified reference with with_shadow=True (container only) -- TEST
INFRASTRUCTURE.

    python -m oracle.make_golden_shadow

tests/golden/gsmini_shadow_tables_320x240.npz   the init-time shadow tables (tacex_b200.calib.ShadowTables)
tests/golden/shadow_sub.npz                      3 frames (config 0 + the first two config-2 envs): the reference's RGB with
                                                 shadows, its gradient bins, and the intermediate scatter-min shadow image captured
                                                 from inside the reference's run (sparse: flat indices + values)
torch_scatter is not in the image: oracle/ref_bootstrap.py supplies scatter_min (torch's scatter_reduce 'amin').
"""

from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

from oracle import ref_bootstrap as rb
from tacex_b200 import synth
from tacex_b200.calib import ShadowTables

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"
H, W = 240, 320


def inputs() -> torch.Tensor:
    c2 = synth.golden_config2(H, W)
    return torch.cat([synth.height_map_mm(synth.config0(H, W)["depth_m"]), c2["hm1"][:2]])


def main() -> None:
    torch.set_num_threads(1)
    tx = rb.load_taxim()
    ShadowTables.from_calib_folder(rb.CALIB_DIR, (H, W)).save(OUT / "gsmini_shadow_tables_320x240.npz")
    hm = inputs()
    press = rb.ref_indentation_depth(hm)
    rgb = rb.ref_render_shadow(tx, hm, press).numpy()
    sh = rb.SCATTER_CAPTURE["shadow_img_flat"].reshape(3, hm.shape[0], H, W).transpose(0, 1).contiguous().numpy()
    dg, _ = rb.ref_deformed_gel(tx, hm, press)
    _, _, im, idr = rb.ref_normals_bins(tx, dg)
    idx = np.flatnonzero(np.isfinite(sh))
    np.savez_compressed(OUT / "shadow_sub.npz", press=press.numpy(), rgb=rgb, idx_mag=im.numpy().astype(np.uint8),
                        idx_dir=idr.numpy().astype(np.uint8), shadow_idx=idx.astype(np.int32), shadow_val=sh.reshape(-1)[idx],
                        input_sum=np.array(hm.double().sum().item()))
    for p in (OUT / "gsmini_shadow_tables_320x240.npz", OUT / "shadow_sub.npz"):
        print(f"{p.name}: {p.stat().st_size / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
