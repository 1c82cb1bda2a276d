"""Golden vectors of the marker image / marker overlay, produced by EXECUTING the reference:
``generate_patch_array()`` and ``FOTSMarkerSimulator.draw_markers`` (fots_marker_sim.py:346-440, compiled from the file with ast,
nothing copied; needs OpenCV, which this image has) and the three arithmetic lines of the RL task's overlay loop
(ball_rolling_taxim_fots.py:933-934: ``frame = tactile_rgb[i] * 255 * dstack([marker / 255] * 3); tactile_rgb[i] = frame / 255``).

    python oracle/make_golden_overlay.py     # build container only -> tests/golden/marker_overlay.npz

TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ast
import math
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
FOTS = Path("/root/reference/source/tacex/tacex/simulation_approaches/fots/fots_marker_sim.py")


def load_reference():
    import copy
    import types

    import cv2

    tree = ast.parse(FOTS.read_text())
    gen = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "generate_patch_array")
    ns = {"np": np, "cv2": cv2, "math": math}
    exec(compile(ast.Module(body=[gen], type_ignores=[]), str(FOTS), "exec"), ns)
    from oracle import ref_bootstrap as rb

    draw = rb.ref_methods(FOTS, "FOTSMarkerSimulator", ["draw_markers"], {"np": np, "math": math})["draw_markers"]
    me = types.SimpleNamespace(patch_array_dict=copy.deepcopy(ns["generate_patch_array"]()))
    return me, draw


def marker_sets() -> np.ndarray:
    """(K, M, 2) float32 marker positions: the reference 11 x 9 grid displaced by smooth fields, sub-pixel offsets of every tenth,
    overlapping markers, markers at / beyond the image border."""
    xs = np.linspace(15, 305, 11).astype(int)
    ys = np.linspace(26, 214, 9).astype(int)
    gx, gy = np.meshgrid(xs, ys)
    base = np.stack([gx.ravel(), gy.ravel()], -1).astype(np.float64)
    rng = np.random.default_rng(11)
    sets = [base, base + rng.uniform(-0.999, 0.999, base.shape), base + rng.normal(0, 6.0, base.shape),
            base * 0.35 + 90 + rng.uniform(0, 1, base.shape),                       # crowded: heavy overlap, the order matters
            base * 1.12 - 20 + rng.uniform(0, 1, base.shape),                       # some markers leave the frame
            np.concatenate([np.array([[-5.5, -5.5], [-6.49, 10.2], [0.0, 0.0], [319.99, 239.99], [325.4, 100.0], [160.5, 245.49],
                                      [160.5, 245.51], [-6.5, 120.0], [-6.51, 121.0], [313.49, 5.0]]), base[:89] + 0.05])]
    return np.stack(sets).astype(np.float32)


def main():
    me, draw = load_reference()
    mk = marker_sets()
    imgs = np.stack([draw(me, mk[k]) for k in range(mk.shape[0])])
    g = torch.Generator().manual_seed(5)
    rgb = torch.rand((2, 240, 320, 3), generator=g)
    over = []
    for i in range(2):  # the task's arithmetic (ball_rolling_taxim_fots.py:933-934), vision_obs at 320 x 240: no resize
        frame = torch.tensor(imgs[i + 2]).unsqueeze(0).movedim(0, 2)
        frame = rgb[i] * 255 * torch.dstack([frame.to(torch.float32) / 255] * 3)
        over.append((frame / 255).numpy())
    out = ROOT / "tests" / "golden" / "marker_overlay.npz"
    np.savez_compressed(out, patch_w15=me.patch_array_dict["patch_array"][:, :, 15].copy(), markers=mk, marker_images=imgs,
                        rgb_seed=np.int64(5), overlay_rows=np.array([96, 128]), overlay=np.stack(over).astype(np.float32)[:, 96:128].copy())
    print(out, out.stat().st_size, "bytes; patch table sum", int(me.patch_array_dict["patch_array"][:, :, 15].sum()))


if __name__ == "__main__":
    main()
