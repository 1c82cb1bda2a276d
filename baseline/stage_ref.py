"""Stages the UNMODIFIED reference implementation of the optical / marker path under baseline/_ref/ (git-ignored, NOT
gpurun-ignored: it travels to the GPU box with the snapshot) so that ``bench.py --impl reference`` and the ``cpu_baseline`` /
``reference_cuda`` legs can time the reference's own code on the benchmark box, where /root/reference does not exist.

Nothing under baseline/_ref/ is product source, nothing there is tracked, and the product never imports it. Files are copied
byte for byte from where they lie:
  source/tacex/tacex/simulation_approaches/gpu_taxim/sim/      (TaximTorch, the reference's Taxim implementation)
  source/tacex/tacex/simulation_approaches/fots/sim/           (MarkerMotion, the reference's FOTS marker model)
  source/tacex_assets/.../GelSight_Mini/calibs/640x480/        (calibration data the reference loads at init)

    python baseline/stage_ref.py          # build container only
"""
from __future__ import annotations

import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
REF = Path("/root/reference")
PIECES = {
    "gpu_taxim/sim": REF / "source/tacex/tacex/simulation_approaches/gpu_taxim/sim",
    "fots/sim": REF / "source/tacex/tacex/simulation_approaches/fots/sim",
    "calibs/640x480": REF / "source/tacex_assets/tacex_assets/data/Sensors/GelSight_Mini/calibs/640x480",
}


def stage(force: bool = False) -> bool:
    if not REF.exists():
        return False
    dst_root = ROOT / "_ref"
    for rel, src in PIECES.items():
        dst = dst_root / rel
        if dst.exists() and not force:
            continue
        if dst.exists():
            shutil.rmtree(dst)
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__"))
    return True


if __name__ == "__main__":
    ok = stage(force="--force" in sys.argv)
    print("staged" if ok else "no reference checkout here")
