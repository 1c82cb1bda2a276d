import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"
H, W = 240, 320


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (B200) device")
    config.addinivalue_line("markers", "refbox: needs the read-only reference checkout (/root/reference)")


def unpack(bits: np.ndarray, n: int) -> np.ndarray:
    return np.unpackbits(bits, axis=1)[:, : H * W].reshape(n, H, W).astype(bool)


@pytest.fixture(scope="session")
def tables():
    from tacex_b200.calib import TaximTables

    return TaximTables.load(GOLDEN / "gsmini_tables_320x240.npz")


@pytest.fixture(scope="session")
def canon_taxim(tables):
    from oracle import canon

    taps = tables.params.blur_taps((H, W))
    return canon.CanonTaxim(H, W, tables.poly_grad.numpy(), tables.background.numpy(), None, taps)


@pytest.fixture(scope="session")
def golden():
    out = {}
    for name in ("config0", "config1_sub", "config2_sub", "gsmini_ref_extras"):
        out[name] = dict(np.load(GOLDEN / f"{name}.npz"))
    return out


@pytest.fixture(scope="session")
def inputs():
    """Height maps [mm] of the golden fixtures, regenerated deterministically (no reference needed)."""
    from tacex_b200 import synth

    c2 = synth.golden_config2(H, W)
    return {
        "config0": synth.height_map_mm(synth.config0(H, W)["depth_m"]),
        "config1_sub": synth.golden_config1(H, W),
        "config2_sub": c2["hm1"],
        "config2_first": c2["hm0"],
        "theta0": c2["theta0"],
        "theta": c2["theta"],
    }


def robust_rgb_check(rgb, rgb_ref_fn, idx_mag, idx_dir, g, n):
    """Protocol P2 of SURVEY.md section 8(c): on well-conditioned pixels of the reference (grad magnitude >= 1e-3)
    the bins must agree on >= 99 % of the pixels; returns the agreement and the agreeing-pixel mask."""
    well = unpack(g["well_bits"], n)
    agree = (idx_mag == g["idx_mag"]) & (idx_dir == g["idx_dir"])
    frac = agree[well].mean() if well.any() else 1.0
    return frac, agree, well
