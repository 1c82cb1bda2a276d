"""In-tree build of libtacex_b200.so (hand-written CUDA for sm_100a, C ABI in include/tacex_b200.h).

    python -m tacex_b200.build            # nvcc cross-compiles without a GPU

The shared library is written next to the sources (tacex_b200/lib/) so that it travels with the tree.
"""

from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB_DIR = PKG / "lib"
LIB = LIB_DIR / "libtacex_b200.so"
SOURCES = ["taxim_kernel.cu", "taxim_generic_kernel.cu", "taxim_shadow_kernel.cu", "fots_kernel.cu", "overlay_kernel.cu", "raster_kernel.cu", "obs_gather_kernel.cu", "fem_kernel.cu", "tx_api.cu", "fem_api.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",          # float32 parity: no implicit contraction, FMAs are written explicitly
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [PKG.parent / "include" / "tacex_b200.h"]
    return any(p.stat().st_mtime > t for p in deps)


def build_lib(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    LIB_DIR.mkdir(exist_ok=True)
    objs = []
    env = dict(os.environ)
    env.pop("CC", None), env.pop("CXX", None)
    for src in SOURCES:
        s = CSRC / src
        if not s.exists():
            continue
        o = LIB_DIR / (s.stem + ".o")
        if force or not o.exists() or o.stat().st_mtime < max(p.stat().st_mtime for p in list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [s, PKG.parent / "include" / "tacex_b200.h"]):
            # float32 bit-parity needs explicit FMAs only (Taxim / FOTS); the float64 gel solver is tolerance-checked and
            # wants contraction (DFMA instead of DMUL + DADD halves the work of the FP64 pipe)
            flags = [f for f in NVCC_FLAGS if not (src.startswith("fem_") and f == "-fmad=false")]
            cmd = [nvcc_path(), *flags, "-ccbin", "/usr/bin/g++", "-I", str(PKG.parent / "include"), "-c", str(s), "-o", str(o)]
            r = subprocess.run(cmd, capture_output=True, text=True, env=env)
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}")
            (LIB_DIR / (s.stem + ".ptxas.txt")).write_text(r.stderr)
        objs.append(str(o))
    cmd = [nvcc_path(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++", "-o", str(LIB), *objs, "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    return LIB


def build_variant(name: str, defines: list[str], source: str = "taxim_kernel.cu") -> Path:
    """Experiment builds (tools/kbench.py): recompiles ONE kernel source with extra -D flags and links it with the other
    objects of the regular build into tacex_b200/lib/libtacex_b200_<name>.so (selected with TACEX_B200_LIB)."""
    build_lib()
    env = dict(os.environ)
    env.pop("CC", None), env.pop("CXX", None)
    s = CSRC / source
    o = LIB_DIR / f"{s.stem}.{name}.o"
    flags = [f for f in NVCC_FLAGS if not (source.startswith("fem_") and f == "-fmad=false")]
    cmd = [nvcc_path(), *flags, *[f"-D{d}" for d in defines], "-ccbin", "/usr/bin/g++", "-I", str(PKG.parent / "include"), "-c", str(s), "-o", str(o)]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    (LIB_DIR / f"{s.stem}.{name}.ptxas.txt").write_text(r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"nvcc failed on {source} ({name})")
    objs = [str(LIB_DIR / (Path(x).stem + ".o")) for x in SOURCES if x != source] + [str(o)]
    out = LIB_DIR / f"libtacex_b200_{name}.so"
    r = subprocess.run([nvcc_path(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++", "-o", str(out), *objs, "-lcudart"],
                       capture_output=True, text=True, env=env)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], [a[2:] for a in sys.argv[i + 2:] if a.startswith("-D")]))
        sys.exit(0)
    p = build_lib(force="--force" in sys.argv, verbose=True)
    print(p)
