#!/usr/bin/env python
"""bench.py -- tactile frames/s of the B200-native engine (contract: see the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--envs E] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one pass of the tactile hot path over a batch of E synthetic contact depth maps per GPU
(height map -> indentation depth -> Taxim RGB 320x240 -> FOTS 63-marker motion), i.e. what one
``GelSightSensor`` update does for E parallel environments. Prints ONE JSON line on rank 0.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

H, W = 240, 320
FRAME_IN_BYTES = H * W * 4            # float32 height map
FRAME_OUT_BYTES = H * W * 3 * 4       # float32 NHWC RGB
MARKER_ROWS, MARKER_COLS = 7, 9       # 63 markers (BASELINE.json north_star; the reference default 11x9 is parity-tested)
M = MARKER_ROWS * MARKER_COLS
ALGO_BYTES_PER_FRAME = FRAME_IN_BYTES + FRAME_OUT_BYTES  # 1,228,800 B
ALL_CPUS = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None  # before any NUMA binding
METRIC = "tactile frames/sec (320x240 RGB+markers) @4096 envs, 1/2/4/8 B200"


def _peaks() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self) -> dict:
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples if len(s) > 2 + i)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(self.samples)}


def bind_to_gpu_numa_node(local: int) -> str:
    """Runs this rank on the CPU cores of the NUMA node its GPU hangs off, BEFORE the pinned host buffers are allocated
    (first touch), so that the end-to-end H2D / D2H copies do not cross the socket interconnect. Best effort."""
    try:
        import torch

        pr = torch.cuda.get_device_properties(local)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(Path(f"/sys/bus/pci/devices/{bdf}/numa_node").read_text().strip())
        if node < 0:
            return "numa: single node"
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"numa node {node} ({len(cpus)} cpus)"
        return f"numa node {node} (no allowed cpu)"
    except Exception as exc:
        return f"numa: unbound ({type(exc).__name__})"


def cpu_port_fps(n_frames: int, with_markers: bool = True) -> tuple[float, int, str]:
    """Times the CPU checker (oracle/ C restatement of the reference algorithm, OpenMP over frames) on a bounded
    sample of the same workload. This is the ONLY place bench.py executes anything under oracle/."""
    import numpy as np

    from oracle import canon
    from tacex_b200 import synth
    from tacex_b200.calib import TaximTables

    t = TaximTables.load(ROOT / "tests" / "golden" / "gsmini_tables_320x240.npz")
    cn = canon.CanonTaxim(H, W, t.poly_grad.numpy(), t.background.numpy(), None, t.params.blur_taps((H, W)))
    cf = canon.CanonFots(H, W, MARKER_ROWS, MARKER_COLS, 15, 26)
    hm = synth.bench_batch(n_frames, n_unique=min(64, n_frames)).numpy()
    th = np.zeros(n_frames, np.float32)
    canon.use_all_threads()
    cn.render(hm[:8], cn.indentation_depth(hm[:8]))  # warm-up
    t0 = time.perf_counter()
    press = cn.indentation_depth(hm)
    o = cn.render(hm, press, want=("deformed", "mask", "rgb"))
    if with_markers:
        cf.step(o["deformed"], o["mask"], press, th)
    dt = time.perf_counter() - t0
    return n_frames / dt, canon.num_threads(), f"{n_frames} frames of the same workload (config-1 sphere presses), one pass"


def fem_cpu_port(n_gels: int = 64) -> tuple[float, int, str]:
    """Gel FEM substep on the host cores with the float64 CPU restatement (libuipc itself has no CPU backend)."""
    import numpy as np

    from oracle import fem_canon as fc
    from tacex_b200 import gel_mesh

    from oracle import canon as _canon

    _canon.use_all_threads()
    m = gel_mesh.box_gel()
    cf = fc.CanonFem(m)
    rng = np.random.default_rng(2)
    offs = rng.uniform(-1, 1, (n_gels, 2)) * np.array([6e-3, 8e-3])
    half = (2e-3, 3e-3, 1e-3)
    z0 = 4.5e-3 + half[2] + 4e-4
    x, v, xp = cf.new_state(n_gels)
    aim = cf.X[cf.attach][None].repeat(n_gels, 0)
    mk = lambda s: [fc.make_indenter(1, (o[0], o[1], z0 - 1e-3 * s / 30), half) for o in offs]  # noqa: E731
    cf.step(x, v, xp, aim, mk(0), mk(1))
    t0 = time.perf_counter()
    for s in (1, 2):
        cf.step(x, v, xp, aim, mk(s), mk(s + 1))
    dt = time.perf_counter() - t0
    from oracle import canon

    return 2 * n_gels / dt, canon.num_threads(), f"{n_gels} gels x 2 steps of the config-3 box press (float64 CPU restatement)"


def fem_gpu(E: int, steps: int, dev, tactile=None) -> dict:
    """Config 3 extra: batched gel FEM substep + FEM marker read-out for E gels (box indenter pressed 0 -> 1 mm in 30 steps)."""
    import numpy as np
    import torch

    from tacex_b200 import fem, gel_mesh

    m = gel_mesh.box_gel()
    eng = fem.GelFemEngine(m, device=dev)
    tri, w = fem.marker_grid_weights(m)
    eng.set_markers(tri, w)
    rng = np.random.default_rng(2)
    offs = rng.uniform(-1, 1, (E, 2)) * np.array([6e-3, 8e-3])
    half = (2e-3, 3e-3, 1e-3)
    z0 = 4.5e-3 + half[2] + 4e-4
    x, v, xp = eng.new_state(E)
    aim = eng.rest_aim(E)
    ctr = lambda s: np.concatenate([offs, np.full((E, 1), z0 - 1e-3 * s / 30)], 1)  # noqa: E731
    inds = [fem.indenter_array(1, ctr(s), half, device=dev) for s in range(2 * steps + 3)]
    mk = torch.empty((E, 2, 128, 2), device=dev)
    for s in range(2):
        eng.step(x, v, xp, aim, inds[s], inds[s + 1], want_stats=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    st = None
    for s in range(2, 2 + steps):
        st = eng.step(x, v, xp, aim, inds[s], inds[s + 1], want_stats=True)
        eng.markers(x, out=mk)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    d = eng.decode_stats(st)
    full = None
    if tactile is not None:
        # config 3 as one step: gel FEM substep + FEM marker read-out + Taxim RGB from the recorded depth maps of the same envs
        t_eng, hm, rgb, depth = tactile
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        f0.record()
        for s in range(2 + steps, 2 + 2 * steps):
            k = min(s, len(inds) - 2)
            eng.step(x, v, xp, aim, inds[k], inds[k + 1], want_stats=False)
            eng.markers(x, out=mk)
            t_eng.render(hm[:E], None, out=rgb[:E], depth_out=depth[:E])
        f1.record()
        torch.cuda.synchronize()
        fms = f0.elapsed_time(f1) / steps
        full = {"workload": f"config 3: {E} envs, gel FEM substep + FEM markers + Taxim RGB 320x240 per step",
                "frames_per_s": E / (fms / 1e3), "ms_per_step": fms}
    out = {"workload": f"{E} gels (572 verts / 2160 tets, float64): implicit-Euler IPC substep + FEM marker read-out",
            "gel_steps_per_s": E / (ms / 1e3), "ms_per_step": ms,
            "newton_iters_mean": float(np.mean([q["newton_iters"] for q in d])),
            "pcg_iters_mean": float(np.mean([q["pcg_iters"] for q in d])),
            "compulsory_bytes_per_gel_step": 315136}
    if full is not None:
        out["config3_full_step"] = full
    return out


def run_reference(args) -> None:
    """--impl reference: the reference's algorithm on the host cores. The reference's own implementation is Python
    (torch + NumPy) inside /root/reference, which does not exist on the GPU box, so the timed code is the oracle's
    C port of it (kind = "port"), with all host threads it can use (OpenMP over frames)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = 256
    vals = []
    for _ in range(max(args.warmup, 0) and 1):
        cpu_port_fps(64)
    for _ in range(max(1, min(args.steps, 5))):
        fps, cores, sample = cpu_port_fps(n)
        vals.append(fps)
    v = statistics.median(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": args.gpus, "steps": len(vals),
        "warmup": 1, "ms_per_step": 1000.0 * n / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{n}-frame bounded sample per step of: {args.envs} envs x 320x240, Taxim RGB + FOTS {M}-marker motion"},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--envs", type=int, default=4096, help="environments per GPU (weak scaling)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--obs-gather", default="fp32", choices=["fp32", "fp32-rect", "fp32-ce", "nccl", "none"],
                    help="N>1: float32 all-gather of the RGB observation into symmetric memory (falls back to NCCL). "
                         "fp32-rect = NVLink peer stores of the non-flat rectangle of every frame + local completion from the flat "
                         "image (bit-identical to gathering whole frames, about half the link bytes); fp32-ce = whole frames by "
                         "copy-engine peer copies (no SM used); fp32 = fp32-ce for 2 GPUs (one peer: the link is not the limit), "
                         "fp32-rect beyond; nccl = all_gather_into_tensor; none = observations stay sharded")
    ap.add_argument("--no-multicast", action="store_true", help="fp32-rect: one store per peer instead of NVSwitch multicast stores")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fem", action="store_true", help="skip the extra gel-FEM measurement (config 3)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    from tacex_b200 import synth
    from tacex_b200.calib import TaximTables
    from tacex_b200.engine import TactileEngine
    from tacex_b200.shard import PeerObsGather, all_gather_obs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own banner / debug output (stdout by default) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    E, K, Wm = args.envs, args.steps, max(args.warmup, 3)

    tables = TaximTables.load(ROOT / "tests" / "golden" / "gsmini_tables_320x240.npz")
    eng = TactileEngine(tables, max_envs=E, device=dev, marker_rows=MARKER_ROWS, marker_cols=MARKER_COLS)

    # ---- synthetic "recorded" depth maps: the shard of envs [rank*E, (rank+1)*E) ---------------------------------
    hm_host = synth.bench_batch(E, seed=rank, n_unique=64).pin_memory()
    theta_host = torch.zeros(E).pin_memory()
    hm = hm_host.to(dev)
    # consecutive steps see DIFFERENT depth maps for every env (the pool of unique maps shifted by a prime number of envs):
    # nothing a step produces can be carried over from the previous one, and the rectangle transport of the observation
    # gather has to restore the previous rectangles of each buffer (contacts that jump, the unfavourable case)
    hm_sets = [hm, hm.roll(37, 0).contiguous(), hm.roll(74, 0).contiguous()]
    theta = theta_host.to(dev)
    depth = torch.empty(E, device=dev)
    rgb = torch.empty((E, H, W, 3), device=dev)
    markers = torch.empty((E, 2, M, 2), device=dev)
    traj0 = torch.zeros((E, 4), device=dev)
    traj_len = torch.zeros(E, device=dev, dtype=torch.int32)
    # N > 1: the observation all-gather of step t runs on a side stream while step t+1 computes (double-buffered RGB)
    do_gather = world > 1 and args.obs_gather in ("fp32", "fp32-rect", "fp32-ce", "nccl")
    use_rects = args.obs_gather == "fp32-rect" or (args.obs_gather == "fp32" and world > 2)
    rgb_buf = [rgb, torch.empty_like(rgb)] if do_gather else [rgb]
    peer, gathered, gather_kind = None, None, "n/a"
    if do_gather:
        if args.obs_gather in ("fp32", "fp32-rect", "fp32-ce"):
            try:
                peer = PeerObsGather(rgb.shape, rgb.dtype, dev, n_slots=2, with_rects=use_rects, multicast=not args.no_multicast)
                rgb_buf = [peer.local_block(0), peer.local_block(1)]  # the kernel renders straight into the gathered buffer
                gather_kind = ("float32 all-gather of RGB, bit-identical to gathering whole frames: NVLink peer stores of every frame's "
                               "non-flat rectangle into symmetric memory ("
                               + ("NVSwitch multicast stores" if (use_rects and peer.mc_rgb[0]) else "one store per peer")
                               + ") + local completion from the flat image, overlapped with the next step"
                               if use_rects else
                               "float32 all-gather of RGB by NVLink peer copies into symmetric memory (copy engines), overlapped with the next step")
            except Exception as exc:  # symmetric memory unavailable on this box
                print(f"[bench] symmetric-memory gather unavailable ({type(exc).__name__}: {exc}); using NCCL", file=sys.stderr)
        if peer is None:
            gathered = [torch.empty((world * E, H, W, 3), device=dev) for _ in range(2)]
            gather_kind = "float32 NCCL all_gather_into_tensor of RGB, overlapped with the next step on a side stream"
    side = torch.cuda.Stream(device=dev, priority=-1) if do_gather else None
    ev_done = [torch.cuda.Event(), torch.cuda.Event()]
    ev_free = [torch.cuda.Event(), torch.cuda.Event()]
    state = {"i": 0}

    def step():
        i = (state["i"] & 1) if do_gather else 0
        state["i"] += 1
        if do_gather:
            torch.cuda.current_stream().wait_event(ev_free[i])  # the gather that read this buffer two steps ago is done
        if peer is not None and use_rects:
            eng.set_rect_output(peer.local_rects(i))
        eng.render(hm_sets[(state["i"] - 1) % 3], None, out=rgb_buf[i], depth_out=depth)
        eng.fots_markers(depth, theta, traj0, traj_len, out=markers)
        if do_gather:
            ev_done[i].record()
            with torch.cuda.stream(side):
                side.wait_event(ev_done[i])
                if peer is not None and use_rects:
                    peer.gather_rects(eng, i, side)
                elif peer is not None:
                    peer.gather(rgb_buf[i], i)
                else:
                    all_gather_obs(rgb_buf[i], gathered[i])
                ev_free[i].record()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(Wm):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    c0 = eng.counters()["kernels_launched"]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(K):
        step()
    if do_gather:
        torch.cuda.current_stream().wait_stream(side)  # the last gathers finish inside the timed region
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = eng.counters()["kernels_launched"] - c0

    # ---- dominant kernel alone (roofline.achieved): K launches of the fused Taxim kernel --------------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    eng.set_rect_output(None)
    for _ in range(K):
        eng.render(hm, None, out=rgb, depth_out=depth)
    e1.record()
    torch.cuda.synchronize()
    kern_ms = e0.elapsed_time(e1) / K

    # ---- end to end through the C ABI with HOST buffers (H2D + kernels + D2H inside the timed region) --------------
    rgb_host = torch.empty((E, H, W, 3)).pin_memory()
    depth_host = torch.empty(E).pin_memory()
    markers_host = torch.empty((E, 2, M, 2)).pin_memory()
    Ke = max(2, min(K, 5))
    eng.step_host(hm_host, rgb_host, depth_host, theta_host, markers_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        eng.step_host(hm_host, rgb_host, depth_host, theta_host, markers_host)  # synchronises the stream itself
    barrier()
    e2e_s = (time.perf_counter() - t0) / Ke
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # ---- config 3 extra: the optional gel FEM substep, measured in the same run (not part of `value`) -------------
    fem_extra = None
    if not args.no_fem and world == 1:
        fem_extra = fem_gpu(min(E, 4096), 3, dev, tactile=(eng, hm, rgb, depth))

    # ---- max over ranks ------------------------------------------------------------------------------------------
    t = torch.tensor([ms, kern_ms, e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, kern_ms, e2e_s = t.tolist()

    if rank == 0:
        peak, peak_src = _peaks()
        frames = world * E * K
        value = frames / (ms / 1e3)
        ach = ALGO_BYTES_PER_FRAME * E / (kern_ms / 1e3) / 1e9
        traffic = None
        tp = ROOT / "profiles" / "traffic.json"
        if tp.exists():
            tj = json.loads(tp.read_text())
            traffic = tj.get("dram_bytes_per_frame", 0) * E if tj.get("dram_bytes_per_frame") else None
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {
                "workload": f"{E} envs/GPU x 320x240: indentation depth + Taxim RGB + FOTS {M}-marker motion, sphere indenters "
                            f"(config-1 distribution, 10% no contact); the optional gel FEM substep (config 3) is reported separately under fem_gel_substep",
                "envs_per_gpu": E, "global_envs": world * E, "parallelism": f"dp{world} (contiguous env shards)",
                "l2_policy": f"inputs larger than L2 ({E * FRAME_IN_BYTES / 1e6:.0f} MB in + {E * FRAME_OUT_BYTES / 1e6:.0f} MB out per step vs 126 MB L2); three input sets cycle, so every env sees a different depth map in consecutive steps",
                "obs_gather": (gather_kind if do_gather else ("none (observations stay sharded)" if world > 1 else "n/a")),
            },
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                         "kernel": "taxim_fused_kernel", "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": ALGO_BYTES_PER_FRAME * E,
                         "peak_source": peak_src,
                         "note": "FP32-FMA bound (exact separable pyramid, 266 MAC/px): see DESIGN.md section 5"},
            "e2e": {"value": world * E / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": E * (FRAME_IN_BYTES + 4),
                    "d2h_bytes_per_step": E * (FRAME_OUT_BYTES + 4 + M * 16), "api": "tx_step_host (C ABI, pinned host buffers)", "host_binding": numa},
            "gpu_launches": launches,
            "clocks": sampler.summary(),
        }
        if not args.no_cpu_baseline:
            if ALL_CPUS:
                os.sched_setaffinity(0, ALL_CPUS)  # the CPU baseline uses every host thread, not just the GPU's NUMA node
            fps, cores, sample = cpu_port_fps(256)
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample}
        if fem_extra is not None:
            line["fem_gel_substep"] = fem_extra
            if not args.no_cpu_baseline:
                g, cores, sample = fem_cpu_port(64)
                line["fem_gel_substep"]["cpu_baseline"] = {"value": g, "unit": "gel-steps/s", "cores": cores, "kind": "port",
                                                           "sample": sample}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
