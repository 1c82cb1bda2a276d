"""Generates tests/golden/*.npz by EXECUTING the unmodified reference (container only) -- TEST INFRASTRUCTURE.

    python -m oracle.make_golden

Inputs are the deterministic synthetic depth maps of tacex_b200/synth.py; the
fixtures hold the reference's outputs on them: indentation depth, deformed gel, contact mask, gradient bins,
RGB, FOTS markers, plus the calibration tables as the reference prepares them. The GPU box has no /root/reference,
so these files are what pins parity there.
"""

from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import torch

from oracle import ref_bootstrap as rb
from tacex_b200 import calib, synth

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"
H, W = 240, 320


def _pack(a: np.ndarray) -> np.ndarray:
    return np.packbits(a.astype(bool).reshape(a.shape[0], -1), axis=1)


def _taxim_record(tx, hm: torch.Tensor, with_rgb: bool) -> dict:
    press = rb.ref_indentation_depth(hm)
    dg, mask = rb.ref_deformed_gel(tx, hm, press)
    mag, _, im, idr = rb.ref_normals_bins(tx, dg)
    rec = {
        "press": press.numpy(),
        "deformed": dg.numpy(),
        "mask_bits": _pack(mask.numpy()),
        "idx_mag": im.numpy().astype(np.uint8),
        "idx_dir": idr.numpy().astype(np.uint8),
        "well_bits": _pack((mag >= 1e-3).numpy()),  # well-conditioned pixels of protocol P2
        "input_sum": np.array(hm.double().sum().item()),
    }
    if with_rgb:
        rec["rgb"] = rb.ref_render(tx, hm, press).numpy()
    return rec


def main() -> None:
    torch.set_num_threads(1)  # the reference's FFT noise depends on the thread count; pin it for reproducibility
    OUT.mkdir(parents=True, exist_ok=True)
    tx = rb.load_taxim()

    # --- calibration tables exactly as the reference prepares them -------------------------------------------------
    t = rb.ref_tables(tx, (H, W))
    with (rb.CALIB_DIR / "params.json").open() as f:
        raw = json.load(f)
    params = calib.TaximParams.from_json(raw)
    tables = calib.TaximTables((H, W), params, t["poly_grad"].contiguous(), t["background"].contiguous(), None,
                               float(t["gel_map_shift"]))
    tables.save(OUT / "gsmini_tables_320x240.npz")
    # the reference's own (FFT-noise) gel map and its blur taps, to pin calib.gaussian_taps
    taps = params.blur_taps((H, W))
    ref_k = {}
    for l, (sx, sy) in enumerate(params.pyramid_sigmas((H, W)) + [params.final_sigma((H, W))]):
        ks = [calib.gaussian_kernel_size(sx), calib.gaussian_kernel_size(sy)]
        k2 = tx._TaximTorch__get_gaussian_kernel2d([sx, sy], ks, torch.float, torch.device("cpu"))
        ref_k[f"k2d_{l}"] = k2.numpy()
    np.savez_compressed(OUT / "gsmini_ref_extras.npz", gel_map_ref=t["gel_map"].numpy(), **ref_k,
                        **{f"kx_{l}": a.numpy() for l, (a, _) in enumerate(taps)})

    # --- config 0 -------------------------------------------------------------------------------------------------
    hm0 = synth.height_map_mm(synth.config0(H, W)["depth_m"])
    np.savez_compressed(OUT / "config0.npz", **_taxim_record(tx, hm0, with_rgb=True))

    # --- config 1 (4 envs with the config-1 distribution + 1 env without contact) -------------------------
    hm1 = synth.golden_config1(H, W)
    np.savez_compressed(OUT / "config1_sub.npz", **_taxim_record(tx, hm1, with_rgb=False))

    # --- config 2: 8 envs, two trajectory samples, FOTS markers for the 11x9 and 9x7 grids -------------------------
    c2 = synth.golden_config2(H, W)
    hm2a, hm2b = c2["hm0"], c2["hm1"]
    rec = _taxim_record(tx, hm2b, with_rgb=False)
    rec["rgb_first2"] = rb.ref_render(tx, hm2b[:2], rb.ref_indentation_depth(hm2b[:2])).numpy()
    rec["kind"] = c2["kind"].numpy()
    rec["theta0"] = c2["theta0"].numpy()
    rec["theta"] = c2["theta"].numpy()
    rec["press0"] = rb.ref_indentation_depth(hm2a).numpy()
    for rows, cols in [(9, 11), (7, 9)]:
        rf = rb.RefFots(tx, rows=rows, cols=cols, x0=15, y0=26)
        rec[f"markers_{rows}x{cols}_step0"] = rf.step(hm2a, rb.ref_indentation_depth(hm2a), c2["theta0"].numpy()).numpy()
        rec[f"markers_{rows}x{cols}_step1"] = rf.step(hm2b, rb.ref_indentation_depth(hm2b), c2["theta"].numpy()).numpy()
    np.savez_compressed(OUT / "config2_sub.npz", **rec)

    for p in sorted(OUT.glob("*.npz")):
        print(f"{p.name}: {p.stat().st_size / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
