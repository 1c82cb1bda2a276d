"""Generates tests/golden/lowres_{H}x{W}.npz by EXECUTING the unmodified reference at coarse tactile resolutions
(container only) -- TEST INFRASTRUCTURE.

    python -m oracle.make_golden_lowres

The reference's RL tasks render the tactile image at 32 x 24 / 32 x 32 (ref: tacex_tasks/.../ball_rolling_taxim_fots.py:306-321,
ball_rolling_tactile_rgb.py:303-318); every blur sigma scales with the shape (taxim_impl.py:33-47). Each fixture holds the
reference's background at that shape and its outputs on `synth.lowres_batch(8, H, W)`.
"""

from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

from oracle import ref_bootstrap as rb
from tacex_b200 import synth

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"
SHAPES = [(24, 32), (32, 32)]
N = 8


def main() -> None:
    torch.set_num_threads(1)
    tx = rb.load_taxim()
    for H, W in SHAPES:
        hm = synth.lowres_batch(N, H, W)
        press = rb.ref_indentation_depth(hm)
        dg, mask = rb.ref_deformed_gel(tx, hm, press)
        mag, _, im, idr = rb.ref_normals_bins(tx, dg)
        t = rb.ref_tables(tx, (H, W))
        np.savez_compressed(
            OUT / f"lowres_{H}x{W}.npz",
            background=t["background"].numpy(), gel_map_ref=t["gel_map"].numpy(),
            press=press.numpy(), deformed=dg.numpy(), mask=mask.numpy(), idx_mag=im.numpy().astype(np.uint8),
            idx_dir=idr.numpy().astype(np.uint8), well=(mag >= 1e-3).numpy(), rgb=rb.ref_render(tx, hm, press).numpy(),
            input_sum=np.array(hm.double().sum().item()),
        )
        p = OUT / f"lowres_{H}x{W}.npz"
        print(f"{p.name}: {p.stat().st_size / 1e3:.1f} kB; press {press.numpy().round(3)}")


if __name__ == "__main__":
    main()
