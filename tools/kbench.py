"""Kernel A/B harness: parity of the fused Taxim kernel against the canonical oracle on a mixed set of frames, then device-timed
throughput on the benchmark's sparse (config-1) and dense workloads, optionally with per-phase clocks.

    TACEX_B200_LIB=tacex_b200/lib/variant.so python tools/kbench.py [--phases] [--n 4096] [--tag name]

Test / profiling infrastructure only (imports oracle/)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import canon
from tacex_b200 import synth
from tacex_b200.calib import TaximTables
from tacex_b200.engine import TactileEngine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H, W = 240, 320


def parity_set():
    d = [synth.config0()["depth_m"], synth.config1(6, seed=11)["depth_m"], synth.config2(8, seed=3)["depth_m"]]
    # contacts at the image edges / corners / the CTA boundary, tiny and huge
    edge = [synth.depth_map(0, 3e-3, cx, cy, 0.0, p) for cx, cy, p in
            [(-9e-3, -6.5e-3, 1e-3), (9.2e-3, 6.8e-3, 1.2e-3), (0.0, -7e-3, 0.8e-3), (0.0, 7e-3, 0.8e-3), (-9.3e-3, 0.0, 1.4e-3),
             (9.3e-3, 0.2e-3, 0.3e-3), (0.0, 0.0, 0.01e-3), (2e-3, 3.4e-3, 4.4e-3), (1e-3, -3.6e-3, 1.0e-3), (0.0, 2.2e-3, 0.6e-3)]]
    d.append(torch.stack(edge))
    hm = synth.height_map_mm(torch.cat(d))
    return torch.cat([hm, synth.dense_batch(6, seed=9, n_unique=6)])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--phases", action="store_true")
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--tag", default=os.path.basename(os.environ.get("TACEX_B200_LIB", "default")))
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--no-parity", action="store_true")
    a = ap.parse_args()
    t = TaximTables.load(ROOT + "/tests/golden/gsmini_tables_320x240.npz")
    res = {"tag": a.tag}
    if not a.no_parity:
        hm = parity_set()
        n = hm.shape[0]
        eng = TactileEngine(t, max_envs=n, marker_rows=7, marker_cols=9)
        eng.set_debug_flags(a.flags)
        dg = torch.empty((n, H, W), device="cuda"); mk = torch.empty((n, H, W), device="cuda", dtype=torch.uint8)
        dep = torch.empty(n, device="cuda")
        rgb = eng.render(hm.cuda(), None, depth_out=dep, deformed_out=dg, mask_out=mk)
        mkr = eng.fots_markers(dep, torch.zeros(n, device="cuda"), torch.zeros((n, 4), device="cuda"),
                               torch.zeros(n, device="cuda", dtype=torch.int32))
        rgb2 = eng.render(hm.cuda(), None)  # without the optional outputs
        torch.cuda.synchronize()
        cn = canon.CanonTaxim(H, W, t.poly_grad.numpy(), t.background.numpy(), None, t.params.blur_taps((H, W)))
        canon.use_all_threads()
        pc = cn.indentation_depth(hm.numpy())
        o = cn.render(hm.numpy(), pc)
        mc = canon.CanonFots(H, W, 7, 9, 15, 26).step(o["deformed"], o["mask"], pc, np.zeros(n, np.float32))
        bad = [i for i in range(n) if not np.array_equal(rgb[i].cpu().numpy(), o["rgb"][i])]
        res["parity"] = {"frames": n, "depth": bool(np.array_equal(dep.cpu().numpy(), pc)),
                         "mask": bool(np.array_equal(mk.cpu().numpy(), o["mask"])),
                         "deformed": bool(np.array_equal(dg.cpu().numpy(), o["deformed"])),
                         "rgb": len(bad) == 0, "rgb_bad_frames": bad[:8], "rgb_again": bool(torch.equal(rgb, rgb2)),
                         "markers_max_px": float(np.abs(mkr.cpu().numpy() - mc).max())}
        del eng
    E = a.n
    eng = TactileEngine(t, max_envs=E, marker_rows=7, marker_cols=9)
    eng.set_debug_flags(a.flags)
    rgb = torch.empty((E, H, W, 3), device="cuda"); dep = torch.empty(E, device="cuda")
    for kind in ("sparse", "dense", "box"):
        pool = {"sparse": lambda: synth.bench_batch(E, n_unique=64), "dense": lambda: synth.dense_batch(E, n_unique=16),
                "box": lambda: synth.height_map_mm(synth.config3_box(32, step=30)["depth_m"]).repeat(E // 32 + 1, 1, 1)[:E].contiguous()}[kind]()
        hm = pool.cuda()
        sets = [hm, hm.roll(5, 0).contiguous()]
        for _ in range(3):
            eng.render(sets[0], None, out=rgb, depth_out=dep)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K = 10
        for j in range(K):
            eng.render(sets[j & 1], None, out=rgb, depth_out=dep)
        e1.record(); torch.cuda.synchronize()
        res[kind + "_fps"] = round(E * K / (e0.elapsed_time(e1) / 1e3))
        if a.phases and kind != "box":
            n = 592
            ticks = torch.zeros((2 * n, 40), dtype=torch.int64, device="cuda")
            eng.set_phase_ticks(ticks)
            eng.set_debug_flags(a.flags)
            eng.render(hm[:n].contiguous(), None, out=rgb[:n])
            torch.cuda.synchronize()
            eng.set_phase_ticks(None)
            eng.set_debug_flags(a.flags)
            tk = ticks.cpu().double()
            press = eng.indentation_depth(hm[:n].contiguous()).cpu()
            tk = tk[(press > 0).repeat_interleave(2)]
            names = {1: "tma", 2: "min", 3: "mask"}
            for L in range(7):
                names[4 + 4 * L] = f"L{L}h"; names[5 + 4 * L] = f"L{L}s"; names[6 + 4 * L] = f"L{L}v"; names[7 + 4 * L] = f"L{L}r"
            names[32] = "halo"; names[33] = "colour"
            prev = 0; ph = {}
            for k in sorted(names):
                ph[names[k]] = float((tk[:, k] - tk[:, prev]).mean()); prev = k
            agg = {"total": float((tk[:, 33] - tk[:, 0]).mean()), "tma+min": ph["tma"] + ph["min"], "mask": ph["mask"],
                   "H": sum(ph[f"L{L}h"] for L in range(7)), "sync": sum(ph[f"L{L}s"] for L in range(7)),
                   "V": sum(ph[f"L{L}v"] for L in range(7)), "reimpose": sum(ph[f"L{L}r"] for L in range(7)),
                   "halo": ph["halo"], "colour": ph["colour"]}
            res[kind + "_phases"] = {k: round(v) for k, v in agg.items()}
            res[kind + "_levels"] = {k: round(v) for k, v in ph.items()}
        del hm, sets
    import json
    print(json.dumps(res))


if __name__ == "__main__":
    main()
