"""CPU suite: host-side logic of the product (no compute calls into the CUDA library)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from conftest import GOLDEN, H, ROOT, W


def test_gaussian_taps_match_reference_kernels(tables, golden):
    """calib.gaussian_taps reproduces the reference's kernels bit-for-bit: outer(ky, kx) == its 2-D kernel
    (taxim_torch.py:363-379) and the kernel sizes are 61,33,17,9,5,3 + 5."""
    taps = tables.params.blur_taps((H, W))
    assert [len(a) for a, _ in taps] == [61, 33, 17, 9, 5, 3, 5]
    ex = golden["gsmini_ref_extras"]
    for l, (kx, ky) in enumerate(taps):
        k2 = torch.mm(ky[:, None], kx[None, :]).numpy()
        assert np.array_equal(k2, ex[f"k2d_{l}"]), l
        assert abs(float(kx.sum()) - 1) < 1e-6
        assert torch.equal(kx, kx.flip(0)), "taps must be exactly symmetric"


def test_tables_roundtrip(tables, tmp_path):
    from tacex_b200.calib import TaximTables

    p = tmp_path / "t.npz"
    tables.save(p)
    t2 = TaximTables.load(p)
    assert torch.equal(t2.poly_grad, tables.poly_grad) and torch.equal(t2.background, tables.background)
    assert t2.gel_map is None and t2.shape == (H, W)
    assert t2.params.pixmm == pytest.approx(0.0295) and t2.params.num_bins == 125


@pytest.mark.refbox
def test_calib_loader_against_reference_tables(tables):
    """Our init-time table preparation (direct separable blur) against the reference's FFT-based one."""
    from oracle import ref_bootstrap as rb

    if not rb.available():
        pytest.skip("reference checkout not present on this machine")
    from tacex_b200.calib import TaximTables

    t = TaximTables.from_calib_folder(rb.CALIB_DIR, (H, W))
    assert torch.equal(t.poly_grad, tables.poly_grad)
    assert (t.background - tables.background).abs().max() < 1e-6
    assert t.gel_map is None and abs(t.gel_map_shift - 5.0) < 1e-6


def test_header_symbols_exported():
    """Every function include/tacex_b200.h declares is exported by the built library (no compute call)."""
    from tacex_b200 import _lib, build

    hdr = (ROOT / "include" / "tacex_b200.h").read_text()
    declared = set(re.findall(r"\b(tx_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"tx_handle", "tx_config", "tx_counters", "tx_status"}
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib_path = build.build_lib()
    lib = C.CDLL(str(lib_path))
    for name in declared:
        assert hasattr(lib, name), name
    lib.tx_abi_version.restype = C.c_int
    assert lib.tx_abi_version() == _lib.TX_ABI_VERSION


def test_config_struct_layout_matches_header():
    """ctypes mirror of tx_config has the size the C compiler gives the header's struct."""
    import subprocess
    import tempfile

    from tacex_b200 import _lib

    src = '#include "tacex_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu\\n", sizeof(tx_config), sizeof(tx_counters));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        (Path(d) / "s.c").write_text(src)
        subprocess.run(["/usr/bin/gcc", "-I", str(ROOT / "include"), str(Path(d) / "s.c"), "-o", str(Path(d) / "s")], check=True)
        a, b = subprocess.run([str(Path(d) / "s")], capture_output=True, text=True, check=True).stdout.split()
    assert int(a) == C.sizeof(_lib.TxConfig) and int(b) == C.sizeof(_lib.TxCounters)


def test_no_cpu_fallback_without_gpu(tables):
    """The product fails loudly when there is no CUDA device (no silent CPU / oracle path)."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tacex_b200 import _lib, engine

    with pytest.raises(_lib.TxError):
        engine.TactileEngine(tables, max_envs=4)


def test_product_never_imports_oracle():
    for p in (ROOT / "tacex_b200").rglob("*.py"):
        txt = p.read_text()
        assert "import oracle" not in txt and "from oracle" not in txt, p
    for p in (ROOT / "tacex_b200" / "csrc").glob("*"):
        assert not re.search(r'#include\s+"[^"]*oracle', p.read_text()), p


def test_synth_shapes_and_determinism():
    from tacex_b200 import synth

    a = synth.config1(3, seed=0)["depth_m"]
    b = synth.config1(3, seed=0)["depth_m"]
    assert torch.equal(a, b) and a.shape == (3, H, W) and a.dtype == torch.float32
    assert float(a.max()) <= 0.029 + 1e-9 and float(a.min()) >= 0.0285 - 1.5e-3 - 1e-6
    c = synth.config2(4, seed=1)
    assert set(c) >= {"depth_m0", "depth_m", "theta0", "theta", "kind"}
    hm = synth.bench_batch(10, n_unique=4)
    assert hm.shape == (10, H, W) and torch.equal(hm[0], hm[4])


def test_sensor_cfg_mirrors_reference_fields():
    from tacex_b200 import sensor, simulators

    cfg = sensor.gelsight_mini_cfg(str(GOLDEN / "gsmini_tables_320x240.npz"), num_envs=2)
    assert cfg.optical_sim_cfg.simulation_approach_class is simulators.B200TaximSimulator
    assert cfg.marker_motion_sim_cfg.simulation_approach_class is simulators.B200FOTSMarkerSimulator
    assert cfg.optical_sim_cfg.tactile_img_res == (320, 240) and cfg.sensor_camera_cfg.clipping_range == (0.024, 0.029)
    assert cfg.marker_motion_sim_cfg.marker_params.num_markers == 99
    for f in ("calib_folder_path", "device", "with_shadow", "tactile_img_res", "gelpad_height", "gelpad_to_camera_min_distance"):
        assert hasattr(cfg.optical_sim_cfg, f)


@pytest.mark.refbox
def test_marker_grid_layout_against_the_executed_reference_function():
    """`_gen_marker_grid` of the reference's FEM marker sensor (tactile_sensor_sapienipc_modified.py:189-247) is a method of a
    class whose module needs Isaac Sim; the method itself only needs numpy / math, so its source is compiled from the reference
    file (ast, nothing copied) and executed on a stand-in `self` with the preset's zero random ranges."""
    import ast
    import math
    import types
    from pathlib import Path

    src = Path("/root/reference/source/tacex/tacex/simulation_approaches/fem_based/sim/tactile_sensor_sapienipc_modified.py")
    if not src.exists():
        pytest.skip("reference checkout not present on this machine")
    from tacex_b200.fem import reference_marker_grid

    tree = ast.parse(src.read_text())
    fn = next(n for c in tree.body if isinstance(c, ast.ClassDef) for n in c.body
              if isinstance(n, ast.FunctionDef) and n.name == "_gen_marker_grid")
    ns = {"np": np, "math": math}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), str(src), "exec"), ns)
    for interval in (2.0625, 1.5, 3.0):
        me = types.SimpleNamespace(marker_interval_range=(interval, interval), marker_rotation_range=0.0,
                                   marker_translation_range=(0.0, 0.0), marker_pos_shift_range=(0.0, 0.0))
        ref = ns["_gen_marker_grid"](me)
        mine = reference_marker_grid(interval)
        assert mine.shape == ref.shape and np.abs(mine - ref).max() <= 1e-15
    g = reference_marker_grid()
    assert g.shape == (91, 2) and abs(g[:, 0].min() + 8.25e-3) < 1e-12 and abs(g[:, 0].max() - 16.5e-3) < 1e-12


@pytest.mark.refbox
def test_marker_weights_against_the_executed_reference_function():
    """`_gen_marker_weight` (tactile_sensor_sapienipc_modified.py:249-329: hull test, 4 nearest face centres by ball tree, first
    containing triangle, barycentric weights) executed from the reference file on the structured gel's top surface -- usdrt (the
    Fabric mesh accessor, Isaac Sim) is replaced by a stand-in that hands out the same triangles -- against
    fem.marker_grid_weights on the same points: the interpolated marker positions on a DEFORMED surface must coincide."""
    import ast
    import importlib.util
    import types
    from pathlib import Path

    base = Path("/root/reference/source/tacex/tacex/simulation_approaches/fem_based/sim")
    if not base.exists():
        pytest.skip("reference checkout not present on this machine")
    from sklearn.neighbors import NearestNeighbors

    from tacex_b200 import gel_mesh
    from tacex_b200.fem import marker_grid_weights, reference_marker_grid

    spec = importlib.util.spec_from_file_location("ref_geometry", str(base / "utils" / "geometry.py"))
    geo = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(geo)
    tree = ast.parse((base / "tactile_sensor_sapienipc_modified.py").read_text())
    fn = next(n for c in tree.body if isinstance(c, ast.ClassDef) for n in c.body
              if isinstance(n, ast.FunctionDef) and n.name == "_gen_marker_weight")
    m = gel_mesh.box_gel()
    tris = np.asarray(m.top_tris)

    class _Attr:
        def Get(self):
            return tris.reshape(-1).tolist()

    class _Mesh:
        def __init__(self, prim):
            pass

        def GetFaceVertexIndicesAttr(self):
            return _Attr()

    usdrt = types.SimpleNamespace(UsdGeom=types.SimpleNamespace(Mesh=_Mesh))
    ns = {"np": np, "NearestNeighbors": NearestNeighbors, "in_hull": geo.in_hull, "usdrt": usdrt}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), str(base), "exec"), ns)
    # camera frame := pad frame shifted so that the reference's asymmetric grid covers the pad
    X = np.asarray(m.X, np.float64)
    shift = np.array([4.125e-3, 0.0])
    pts_cam = reference_marker_grid()
    surf_cam = X.copy()
    surf_cam[:, :2] += shift
    me = types.SimpleNamespace(init_surface_vertices_camera=torch.from_numpy(surf_cam), gelpad_obj=types.SimpleNamespace(fabric_prim=None))
    idx_ref, w_ref = ns["_gen_marker_weight"](me, pts_cam.copy())
    # same points in the pad frame through the product's host code (no padding: pad_to = number of valid markers)
    pts_pad = pts_cam - shift
    on = geo.in_hull(pts_cam, surf_cam[:, :2])
    tri, w = marker_grid_weights(m, pad_to=int(on.sum()), points_xy=pts_pad[on])
    assert idx_ref.shape[0] == tri.shape[0] == on.sum() and on.sum() >= 70  # 11 x 7 of the 13 x 7 points lie on this pad
    rng = np.random.default_rng(0)
    Xd = X + 2e-4 * rng.standard_normal(X.shape)  # a deformed surface
    p_ref = (Xd[idx_ref] * w_ref[..., None]).sum(1)
    p_me = (Xd[tri] * w[..., None]).sum(1)
    assert np.abs(p_ref - p_me).max() <= 1e-9


def test_reference_arm_under_torchrun_prints_one_json_line_from_rank_0():
    """bench.py --impl reference launched like the driver launches it for N > 1: rank 0 alone runs and prints ONE JSON line with
    the contract's keys, the other rank exits 0 without work."""
    import json
    import os
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    port = 29800 + os.getpid() % 150
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(root / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(root))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0 and d["n_gpus"] == 2
    staged = (root / "baseline" / "_ref" / "gpu_taxim" / "sim" / "taxim_torch.py").exists()
    assert d["cpu_baseline"]["kind"] == ("reference" if staged else "port") and d["cpu_baseline"]["cores"] >= 1
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the reference arm must still use every host thread it may
    n_cpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    assert d["cpu_baseline"]["cores"] == n_cpu, d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_bench_has_no_cpu_path_for_the_product_arm():
    """Without a CUDA device the product arm of bench.py must fail loudly instead of timing anything on the CPU."""
    import subprocess
    import sys
    from pathlib import Path

    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    root = Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, str(root / "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                       timeout=300, cwd=str(root))
    assert r.returncode != 0 and "CUDA" in (r.stderr + r.stdout) and not any(l.startswith("{") for l in r.stdout.splitlines())


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_indenter_meshes_are_closed_triangle_soups(kind):
    """The config-2 indenter meshes handed to tx_fem_set_indenter_mesh (sphere as an icosphere, flat cylinder, wedge, cone): no
    degenerate triangle, every edge shared by exactly two triangles (closed surface), lowest point at the origin, body towards +z."""
    from collections import Counter

    import numpy as np

    from tacex_b200 import synth

    t = synth.indenter_mesh(kind, 3e-3)
    assert t.ndim == 3 and t.shape[1:] == (3, 3) and len(t) >= 8
    n = np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0])
    assert (np.linalg.norm(n, axis=1) > 0).all()
    verts, inv = np.unique(t.reshape(-1, 3).round(12), axis=0, return_inverse=True)
    ids = inv.reshape(-1, 3)
    edges = Counter()
    for a, b, c in ids:
        for u, v in ((a, b), (b, c), (c, a)):
            edges[(min(u, v), max(u, v))] += 1
    assert set(edges.values()) == {2}
    assert abs(verts[:, 2].min()) < 1e-15 and verts[:, 2].max() > 1e-3
    assert len(verts) - len(edges) + len(t) == 2  # Euler characteristic of a sphere
