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


BLUR_RADII = (30, 16, 8, 4, 2, 1, 2)


def fp32_lane_ops_per_frame(hm_pool) -> float:
    """FP32 lane-operations (FMA / FADD / FMUL on one float) the exact separable pyramid NEEDS for these frames given the
    kernel's exact-zero skipping: per level, rows x columns that can be non-zero x (2R+1) taps for the horizontal pass and the
    same for the vertical pass (R adds of the symmetric pairs + R+1 multiply-adds), the region growing by R per level.
    A dense frame (gel map or full-frame contact) costs 266 x 76,800 = 20.4 M; the colour stage is not counted."""
    import numpy as np

    hm = hm_pool.numpy() if hasattr(hm_pool, "numpy") else np.asarray(hm_pool)
    tot = 0.0
    for f in hm:
        m = f.min()
        d = max(m / 1000.0 - 0.024, 0.0)
        press = (0.0045 - d) * 1000.0 if d <= 0.0045 else 0.0
        c = (f - m - press) < 0
        if press <= 0 or not c.any():
            continue
        ys, xs = np.nonzero(c)
        r0, r1, c0, c1 = ys.min(), ys.max(), xs.min(), xs.max()
        for R in BLUR_RADII:
            h, w = r1 - r0 + 1, c1 - c0 + 1
            c0, c1 = max(c0 - R, 0), min(c1 + R, W - 1)
            tot += h * (c1 - c0 + 1) * (2 * R + 1)            # horizontal pass: input rows x output columns
            r0, r1 = max(r0 - R, 0), min(r1 + R, H - 1)
            tot += (r1 - r0 + 1) * (c1 - c0 + 1) * (2 * R + 1)  # vertical pass: output rows x columns
    return tot / len(hm)


def fem_cpu_port(n_gels: int = 128, n_steps: int = 4) -> tuple[float, int, str]:
    """Gel FEM substep on the host cores with the float64 CPU restatement (libuipc itself has no CPU backend): the config-3
    box edge / corner press, steps 21..24 of 30 after an untimed run-in of the same gels."""
    import numpy as np

    from oracle import canon as _canon
    from oracle import fem_canon as fc
    from tacex_b200 import gel_mesh, synth

    _canon.use_all_threads()
    m = gel_mesh.box_gel()
    cf = fc.CanonFem(m)
    c3 = synth.config3_box(n_gels, seed=2, step=0)
    R = c3["R"].numpy()
    half = np.asarray(c3["half"])
    low = np.abs(R[:, 2, :] * half[None]).sum(1)
    x, v, xp = cf.new_state(n_gels)
    aim = cf.X[cf.attach][None].repeat(n_gels, 0)
    mk = lambda s: [fc.make_indenter(1, (c3["cx"][i].item(), c3["cy"][i].item(), 4.5e-3 + low[i] + 4e-4 - 1e-3 * s / 30), half, R[i])  # noqa: E731
                    for i in range(n_gels)]
    first = 21
    for s in range(0, first, 3):  # coarse run-in (3 press steps per solver step) so that the timed steps see real contact
        cf.step(x, v, xp, aim, mk(s), mk(s + 3))
    t0 = time.perf_counter()
    for s in range(first, first + n_steps):
        cf.step(x, v, xp, aim, mk(s), mk(s + 1))
    dt = time.perf_counter() - t0
    return (n_steps * n_gels / dt, _canon.num_threads(),
            f"{n_gels} gels x {n_steps} steps (steps {first}..{first + n_steps - 1} of 30) of the config-3 box edge / corner press "
            f"(float64 CPU restatement, OpenMP over gels)")


class Config3:
    """BASELINE config 3 inputs: a rigid 4 x 6 x 2 mm box pressed with an EDGE / a CORNER 0 -> 1 mm into the gel over 30 steps
    (seed 2 for the pose jitter) -- the SAME poses as FEM indenters (tx_fem_indenter) and as recorded depth maps of the sensor
    camera (the reference renders the tactile image from the camera's depth image of the indenter, not from the FEM state).
    ``n_unique`` poses tiled to E envs (every env still runs the full path). Timed steps = the deepest part of the press."""

    def __init__(self, E: int, dev, seed: int = 2, n_unique: int = 64, first: int = 24, last: int = 30):
        import numpy as np
        import torch

        from tacex_b200 import fem, synth

        self.E, self.dev, self.first, self.last = E, dev, first, last
        k = min(n_unique, E)
        c3 = synth.config3_box(k, seed=seed, step=0)
        R = c3["R"].numpy()
        half = np.asarray(c3["half"])
        low = np.abs(R[:, 2, :] * half[None]).sum(1)  # lowest corner below the box centre
        reps = (E + k - 1) // k
        tile = lambda a: np.concatenate([a] * reps, 0)[:E]  # noqa: E731
        self.idx = torch.arange(E, device=dev) % k
        top = 4.5e-3
        self.inds, self.hm_pool = [], {}
        for s in range(last + 1):
            ctr = np.stack([c3["cx"].numpy(), c3["cy"].numpy(), top + low + 4e-4 - 1e-3 * s / 30], 1)
            self.inds.append(fem.indenter_array(1, tile(ctr), half, tile(R), device=dev))
        for s in range(first, last + 1):
            self.hm_pool[s] = synth.height_map_mm(synth.config3_box(k, seed=seed, step=s)["depth_m"]).to(dev)

    def height_maps(self, s: int):
        return self.hm_pool[min(max(s, self.first), self.last)][self.idx].contiguous()


def fem_gpu(E: int, dev, tactile=None) -> dict:
    """Config 3: batched gel FEM substep + FEM marker read-out for E gels, and the full config-3 step (FEM substep + FEM markers
    + Taxim RGB from the recorded depth maps of the SAME box poses). Steps 0..23 of the 30-step press run untimed; the timed
    steps are 24..29 (deepest contact)."""
    import numpy as np
    import torch

    from tacex_b200 import fem, gel_mesh

    m = gel_mesh.box_gel()
    eng = fem.GelFemEngine(m, device=dev)
    tri, w = fem.marker_grid_weights(m)
    eng.set_markers(tri, w)
    c3 = Config3(E, dev)
    x, v, xp = eng.new_state(E)
    aim = eng.rest_aim(E)
    mk = torch.empty((E, 2, 128, 2), device=dev)
    for s in range(c3.first - 3):
        eng.step(x, v, xp, aim, c3.inds[s], c3.inds[s + 1], want_stats=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    st = None
    for s in range(c3.first - 3, c3.first):
        st = eng.step(x, v, xp, aim, c3.inds[s], c3.inds[s + 1], want_stats=True)
        eng.markers(x, out=mk)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    d = eng.decode_stats(st)
    out = {"workload": f"{E} gels (572 verts / 2160 tets, float64): implicit-Euler IPC substep + FEM marker read-out, config-3 box "
                       f"edge / corner press (steps {c3.first - 3}..{c3.first - 1} of 30)",
           "gel_steps_per_s": E / (ms / 1e3), "ms_per_step": ms,
           "newton_iters_mean": float(np.mean([q["newton_iters"] for q in d])),
           "pcg_iters_mean": float(np.mean([q["pcg_iters"] for q in d])),
           "compulsory_bytes_per_gel_step": 315136,
           "hbm_frac_on_compulsory_bytes": 315136 * E / (ms / 1e3) / 1e9 / _peaks()[0]}
    if tactile is not None:
        t_eng, rgb, depth = tactile
        hms = [c3.height_maps(s + 1) for s in range(c3.first, c3.last)]
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        f0.record()
        for j, s in enumerate(range(c3.first, c3.last)):
            eng.step(x, v, xp, aim, c3.inds[s], c3.inds[s + 1], want_stats=False)
            eng.markers(x, out=mk)
            t_eng.render(hms[j], None, out=rgb[:E], depth_out=depth[:E])
        f1.record()
        torch.cuda.synchronize()
        fms = f0.elapsed_time(f1) / (c3.last - c3.first)
        out["config3_full_step"] = {
            "workload": f"config 3: {E} envs, gel FEM substep + FEM markers + Taxim RGB 320x240 from the depth maps of the same box "
                        f"poses, steps {c3.first}..{c3.last - 1} of the 30-step press",
            "frames_per_s": E / (fms / 1e3), "ms_per_step": fms}
    # config-2 indenters through the FEM path: wedge / cone as prescribed triangle meshes, vertex-face contact in both directions
    from tacex_b200 import synth

    out["mesh_indenters"] = {}
    Em = min(E, 1024)
    rng = np.random.default_rng(7)
    offs = np.concatenate([rng.uniform(-1, 1, (Em, 2)) * np.array([6e-3, 8e-3]), np.zeros((Em, 1))], 1)
    eng.set_contact_surface(m.top_tris)
    for kind, name in ((2, "wedge"), (3, "cone")):
        eng.set_indenter_mesh(synth.indenter_mesh(kind, 3e-3))
        xm, vm, xpm = eng.new_state(Em)
        aim_m = eng.rest_aim(Em)
        pose = lambda sidx: fem.indenter_array(2, offs + np.array([0.0, 0.0, 4.5e-3 + 4e-4 - 1e-3 * sidx / 10]), (0, 0, 0), device=dev)  # noqa: E731
        inds = [pose(sidx) for sidx in range(11)]
        for sidx in range(7):
            eng.step(xm, vm, xpm, aim_m, inds[sidx], inds[sidx + 1], want_stats=False)
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for sidx in range(7, 10):
            stm = eng.step(xm, vm, xpm, aim_m, inds[sidx], inds[sidx + 1], want_stats=True)
        g1.record()
        torch.cuda.synchronize()
        dm = eng.decode_stats(stm)
        out["mesh_indenters"][name] = {"gel_steps_per_s": Em / (g0.elapsed_time(g1) / 3 / 1e3), "ms_per_step": g0.elapsed_time(g1) / 3, "gels": Em,
                                       "converged": float(np.mean([q["converged"] for q in dm])),
                                       "newton_iters_mean": float(np.mean([q["newton_iters"] for q in dm]))}
    out["mesh_indenters"]["workload"] = ("config-2 wedge / 60 deg cone as prescribed triangle meshes pressed 0 -> 1 mm over 10 steps (timed: steps 7..9), gel "
                                         "vertices vs indenter triangles + indenter vertices vs the gel's 240 top triangles, ACCD")
    eng.set_indenter_mesh(None)
    eng.set_contact_surface(None)
    return out


REF_BATCH = 32  # envs per reference call: the batch size at which the reference's CPU path is fastest


def cpu_reference_fps(reps: int, warmup: int = 1) -> tuple[float, int, str] | None:
    """The UNMODIFIED reference (TaximTorch.render_direct + indentation depth + the per-env MarkerMotion loop, staged under the
    git-ignored baseline/_ref/ by baseline/stage_ref.py) on all host cores, on a bounded sample of the same workload.
    None when the staged copy is absent (then the oracle's C port is timed instead)."""
    from baseline import ref_arm

    if not ref_arm.available():
        return None
    from tacex_b200 import synth

    hm = synth.bench_batch(64, n_unique=64)
    fps, threads = ref_arm.time_cpu(hm, REF_BATCH, reps, warmup=warmup, with_markers=True)
    return fps, threads, (f"{reps} sensor updates of {REF_BATCH} envs each (config-1 sphere presses): the reference's own TaximTorch.render_direct "
                          f"+ compute_indentation_depth + per-env MarkerMotion.marker_sim loop, torch CPU, {threads} threads")


def run_reference(args) -> None:
    """--impl reference: the reference's OWN implementation of the path on the host cores -- the unmodified TaximTorch /
    MarkerMotion code staged under baseline/_ref/ (kind = "reference"), with all host threads torch can use. Each step is a
    bounded sample (one sensor update of 32 envs). Falls back to the oracle's C port (kind = "port") only if the staged copy
    is missing."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K = max(1, min(args.steps, 20))
    Wm = max(1, min(args.warmup, 3))
    r = cpu_reference_fps(K, warmup=Wm)
    if r is not None:
        v, cores, sample = r
        kind, n = "reference", REF_BATCH
    else:
        for _ in range(Wm):
            cpu_port_fps(64)
        n = 256
        vals = [cpu_port_fps(n)[0] for _ in range(min(K, 5))]
        K = len(vals)
        v = statistics.median(vals)
        _, cores, sample = cpu_port_fps(8)
        kind = "port"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": args.gpus, "steps": K,
        "warmup": Wm, "ms_per_step": 1000.0 * n / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{n}-env bounded sample per step of: {args.envs} envs x 320x240, indentation depth + Taxim RGB + FOTS {M}-marker motion, "
                               f"sphere indenters (config-1 distribution)"},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


class StepRunner:
    """One rank's step loop: [gel FEM substep + FEM markers] + fused Taxim render + FOTS markers for E envs, and (N > 1) the
    all-gather of the float32 RGB observation on a side stream, overlapped with the next step (double-buffered)."""

    def __init__(self, args, eng, E, dev, world, rank, hm_sets, fem_ctx=None):
        import torch

        from tacex_b200.shard import PeerObsGather

        self.args, self.eng, self.E, self.dev, self.world, self.rank = args, eng, E, dev, world, rank
        self.hm_sets, self.fem = hm_sets, fem_ctx
        self.theta = torch.zeros(E, device=dev)
        self.depth = torch.empty(E, device=dev)
        rgb = torch.empty((E, H, W, 3), device=dev)
        self.markers = torch.empty((E, 2, M, 2), device=dev)
        self.traj0 = torch.zeros((E, 4), device=dev)
        self.traj_len = torch.zeros(E, device=dev, dtype=torch.int32)
        self.do_gather = world > 1 and args.obs_gather in ("fp32", "fp32-fused", "fp32-rect", "fp32-ce", "nccl", "u8")
        self.u8 = world > 1 and args.obs_gather == "u8"
        self.use_rects = args.obs_gather in ("fp32-rect", "fp32-fused", "fp32")
        self.fused = False
        self.rgb_buf = [rgb, torch.empty_like(rgb)] if self.do_gather else [rgb]
        self.peer, self.gathered, self.gather_kind = None, None, "n/a"
        if self.u8:
            self.peer = PeerObsGather(rgb.shape, torch.uint8, dev, n_slots=2, with_rects=False, multicast=False)
            self.rgb_buf = [rgb, rgb]  # float32 frames stay local; the uint8 copy is what travels
            self.gather_kind = ("uint8 all-gather of RGB (round(rgb * 255), one extra conversion kernel per step) by NVLink peer copies into "
                                "symmetric memory (copy engines), overlapped with the next step")
        elif self.do_gather:
            if args.obs_gather in ("fp32", "fp32-fused", "fp32-rect", "fp32-ce"):
                try:
                    self.peer = PeerObsGather(rgb.shape, rgb.dtype, dev, n_slots=2, with_rects=self.use_rects,
                                              multicast=not args.no_multicast)
                    self.rgb_buf = [self.peer.local_block(0), self.peer.local_block(1)]  # render straight into the gathered buffer
                    self.fused = self.use_rects and args.obs_gather in ("fp32", "fp32-fused") and self.peer.fused_available()
                    if args.obs_gather == "fp32" and not self.fused and world == 2:
                        # no multicast mapping: with one peer the link is not the limit and the copy engines win over the
                        # rectangle push kernel (no SM taken)
                        self.use_rects = False
                    self.gather_kind = (
                        "float32 all-gather of RGB, bit-identical to gathering whole frames, FUSED into the render kernel: its epilogue "
                        "stores every frame's non-flat rectangle through the NVSwitch multicast mapping of the symmetric buffers "
                        "(multimem.st), the flat remainder is completed locally after a cross-rank barrier"
                        if self.fused else
                        "float32 all-gather of RGB, bit-identical to gathering whole frames: NVLink peer stores of every frame's "
                        "non-flat rectangle into symmetric memory ("
                        + ("NVSwitch multicast stores" if (self.use_rects and self.peer.mc_rgb[0]) else "one store per peer")
                        + ") + local completion from the flat image, overlapped with the next step"
                        if self.use_rects else
                        "float32 all-gather of RGB by NVLink peer copies into symmetric memory (copy engines), overlapped with the next step")
                except Exception as exc:  # symmetric memory unavailable on this box
                    print(f"[bench] symmetric-memory gather unavailable ({type(exc).__name__}: {exc}); using NCCL", file=sys.stderr)
            if self.peer is None:
                self.gathered = [torch.empty((world * E, H, W, 3), device=dev) for _ in range(2)]
                self.gather_kind = "float32 NCCL all_gather_into_tensor of RGB, overlapped with the next step on a side stream"
        self.side = torch.cuda.Stream(device=dev, priority=-1) if self.do_gather else None
        self.ev_done = [torch.cuda.Event(), torch.cuda.Event()]
        self.ev_free = [torch.cuda.Event(), torch.cuda.Event()]
        self.i = 0
        self.last = (0, 0)  # (slot, input-set index) of the last step

    def step(self):
        import torch

        from tacex_b200.shard import all_gather_obs

        i = (self.i & 1) if self.do_gather else 0
        k = self.i % len(self.hm_sets)
        self.i += 1
        self.last = (i, k)
        if self.do_gather:
            torch.cuda.current_stream().wait_event(self.ev_free[i])  # the gather that read this buffer two steps ago is done
        if self.fem is not None:
            self.fem.step()
        if self.fused:
            self.peer.begin_fused(self.eng, i)  # barrier: every rank is done with the slot; multicast output on
        elif self.peer is not None and self.use_rects:
            self.eng.set_rect_output(self.peer.local_rects(i))
        self.eng.render(self.hm_sets[k], None, out=self.rgb_buf[i], depth_out=self.depth)
        self.eng.fots_markers(self.depth, self.theta, self.traj0, self.traj_len, out=self.markers)
        if self.u8:
            self.eng.marker_overlay(None, self.rgb_buf[i], apply=False, rgb_u8_out=self.peer.local_block(i))
        if self.do_gather:
            self.ev_done[i].record()
            with torch.cuda.stream(self.side):
                self.side.wait_event(self.ev_done[i])
                if self.u8:
                    self.peer.gather(self.peer.local_block(i), i)
                elif self.fused:
                    self.peer.finish_fused(self.eng, i, self.side)
                elif self.peer is not None and self.use_rects:
                    self.peer.gather_rects(self.eng, i, self.side)
                elif self.peer is not None:
                    self.peer.gather(self.rgb_buf[i], i)
                else:
                    all_gather_obs(self.rgb_buf[i], self.gathered[i])
                self.ev_free[i].record()

    def finish(self):
        import torch

        if self.do_gather:
            torch.cuda.current_stream().wait_stream(self.side)  # the last gathers finish inside the timed region
        self.eng.set_rect_output(None)
        self.eng.set_multicast_output(0, 0)

    def verify_gather(self, remote_sets_fn) -> bool | None:
        """After a step: re-render the block of ONE remote rank locally from that rank's (deterministic) inputs and compare it,
        bit for bit, with what the gather delivered into this rank's buffer."""
        import torch

        if not self.do_gather:
            return None
        torch.cuda.synchronize()
        slot, k = self.last
        r = (self.rank + 1) % self.world
        full = self.peer.bufs[slot] if self.peer is not None else self.gathered[slot]
        ref = torch.empty((self.E, H, W, 3), device=self.dev)
        self.eng.set_rect_output(None)
        self.eng.render(remote_sets_fn(r)[k], None, out=ref)
        if self.u8:
            ref8 = torch.empty((self.E, H, W, 3), device=self.dev, dtype=torch.uint8)
            self.eng.marker_overlay(None, ref, apply=False, rgb_u8_out=ref8)
            torch.cuda.synchronize()
            return bool(torch.equal(full[r * self.E:(r + 1) * self.E], ref8))
        torch.cuda.synchronize()
        own = torch.equal(full[self.rank * self.E:(self.rank + 1) * self.E], self.rgb_buf[slot])
        return bool(own and torch.equal(full[r * self.E:(r + 1) * self.E], ref))


class FemCtx:
    """Gel FEM substep + FEM marker read-out of E gels walking through the config-3 press (used by the config-4 step)."""

    def __init__(self, E, dev, c3):
        import torch

        from tacex_b200 import fem, gel_mesh

        m = gel_mesh.box_gel()
        self.eng = fem.GelFemEngine(m, device=dev)
        tri, w = fem.marker_grid_weights(m)
        self.eng.set_markers(tri, w)
        self.c3 = c3
        self.x, self.v, self.xp = self.eng.new_state(E)
        self.aim = self.eng.rest_aim(E)
        self.mk = torch.empty((E, 2, 128, 2), device=dev)
        self.s = 0

    def step(self):
        s = min(self.s, len(self.c3.inds) - 2)
        self.eng.step(self.x, self.v, self.xp, self.aim, self.c3.inds[s], self.c3.inds[s + 1], want_stats=False)
        self.eng.markers(self.x, out=self.mk)
        self.s += 1


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--envs", type=int, default=4096, help="environments per GPU (weak scaling)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--obs-gather", default="fp32", choices=["fp32", "fp32-fused", "fp32-rect", "fp32-ce", "nccl", "u8", "none"],
                    help="N>1: float32 all-gather of the RGB observation into symmetric memory (falls back to NCCL). "
                         "fp32-rect = NVLink peer stores of the non-flat rectangle of every frame + local completion from the flat "
                         "image (bit-identical to gathering whole frames, about half the link bytes); fp32-ce = whole frames by "
                         "copy-engine peer copies (no SM used); fp32 = fp32-fused when the allocation has an NVSwitch multicast mapping (the "
                         "render kernel's epilogue stores every frame's rectangle through it: compute + collective in one kernel), "
                         "else fp32-ce for 2 GPUs / fp32-rect beyond; nccl = all_gather_into_tensor; u8 = the observation converted to uint8 by one extra "
                         "kernel (tx_marker_overlay) and gathered by copy-engine peer copies: a quarter of the link bytes (reported "
                         "alongside the mandated float32 gather, SURVEY 8e); none = observations stay sharded")
    ap.add_argument("--no-multicast", action="store_true", help="fp32-rect: one store per peer instead of NVSwitch multicast stores")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fem", action="store_true", help="skip the gel-FEM measurements (config 3 / config 4)")
    ap.add_argument("--no-extras", action="store_true", help="skip value_dense / reference_cuda (the contract keys stay)")
    ap.add_argument("--config4-envs", type=int, default=1024, help="envs per GPU of the config-4 step (8192 = 8 x 1024)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    from tacex_b200 import synth
    from tacex_b200.calib import TaximTables
    from tacex_b200.engine import TactileEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local)
    # stdout carries exactly ONE JSON line: everything a library prints to file descriptor 1 while the bench runs (NCCL's version
    # banner, for one) goes to stderr; the line itself is written to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own banner / debug output (stdout by default) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    E, K, Wm = args.envs, args.steps, max(args.warmup, 3)

    tables = TaximTables.load(ROOT / "tests" / "golden" / "gsmini_tables_320x240.npz")
    eng = TactileEngine(tables, max_envs=E, device=dev, marker_rows=MARKER_ROWS, marker_cols=MARKER_COLS)

    # ---- synthetic "recorded" depth maps: the shard of envs [rank*E, (rank+1)*E) ---------------------------------
    # consecutive steps see DIFFERENT depth maps for every env (the pool of unique maps shifted by a prime number of envs):
    # nothing a step produces can be carried over from the previous one, and the rectangle transport of the observation
    # gather has to restore the previous rectangles of each buffer (contacts that jump, the unfavourable case)
    def sets_of(r: int, host=None):
        h = synth.bench_batch(E, seed=r, n_unique=64) if host is None else host
        d = h.to(dev)
        return [d, d.roll(37, 0).contiguous(), d.roll(74, 0).contiguous()]

    hm_host = synth.bench_batch(E, seed=rank, n_unique=64).pin_memory()
    theta_host = torch.zeros(E).pin_memory()
    hm_sets = sets_of(rank, hm_host)
    hm = hm_sets[0]
    run = StepRunner(args, eng, E, dev, world, rank, hm_sets)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(runner, k_steps, w_steps):
        for _ in range(w_steps):
            runner.step()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a.record()
        for _ in range(k_steps):
            runner.step()
        runner.finish()
        b.record()
        barrier()
        return a.elapsed_time(b)

    for _ in range(Wm):
        run.step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    c0 = eng.counters()["kernels_launched"]
    ms = timed(run, K, 0)
    launches = eng.counters()["kernels_launched"] - c0
    gk = gather_kind_of(world, args, run)
    gather_ok = None
    if run.do_gather:
        run.step()
        run.finish()
        barrier()
        gather_ok = run.verify_gather(lambda r: sets_of(r))
        barrier()

    # ---- dominant kernel alone (roofline.achieved): K launches of the fused Taxim kernel --------------------------
    rgb, depth = run.rgb_buf[0], run.depth
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(K):
        eng.render(hm, None, out=rgb, depth_out=depth)
    e1.record()
    torch.cuda.synchronize()
    kern_ms = e0.elapsed_time(e1) / K

    # ---- the same step on DENSE contacts (exact-zero skipping gains nothing there) and on the config-3 box presses ----
    extras = {}
    if not args.no_extras:
        for name, pool, what in (
            ("value_dense", synth.dense_batch(E, seed=4 + rank, n_unique=16),
             "dense contacts: flat punches / large spheres / large tilted boxes covering 30-45 % of the frame, no env without contact"),
            ("value_config3_box", synth.height_map_mm(synth.config3_box(64, seed=2 + rank, step=30)["depth_m"]).repeat((E + 63) // 64, 1, 1)[:E].contiguous(),
             "config-3 depth maps: 4 x 6 x 2 mm box pressed 1 mm with an edge / a corner"),
        ):
            d = pool.to(dev)
            sets = [d, d.roll(5, 0).contiguous(), d.roll(11, 0).contiguous()]
            for _ in range(3):
                eng.render(sets[0], None, out=rgb, depth_out=depth)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for j in range(K):
                eng.render(sets[j % 3], None, out=rgb, depth_out=depth)
                eng.fots_markers(depth, run.theta, run.traj0, run.traj_len, out=run.markers)
            b.record()
            torch.cuda.synchronize()
            dms = a.elapsed_time(b) / K
            extras[name] = {"value": E / (dms / 1e3), "unit": "frames/s per GPU", "ms_per_step": dms, "workload": what,
                            "fp32_lane_ops_per_frame": fp32_lane_ops_per_frame(pool[:64])}
            del d, sets

        # config 2: 1024 envs, sphere / flat cylinder / 90 deg wedge / 60 deg cone with random pose and in-plane yaw, RGB +
        # FOTS markers along a trajectory: first contact (x0, y0, theta0), then sheared by U(-0.5, 0.5) mm and twisted by U(-30, 30) deg
        E2 = min(1024, E)
        c2 = synth.config2(min(128, E2), seed=1 + rank)
        rep = (E2 + c2["depth_m"].shape[0] - 1) // c2["depth_m"].shape[0]
        tile = lambda t: t.repeat(rep, *([1] * (t.dim() - 1)))[:E2].contiguous().to(dev)  # noqa: E731
        hm_a, hm_b = tile(synth.height_map_mm(c2["depth_m0"])), tile(synth.height_map_mm(c2["depth_m"]))
        th_a, th_b = tile(c2["theta0"]), tile(c2["theta"])
        rgb2, dep2, mk2 = rgb[:E2], depth[:E2], run.markers[:E2]
        tr2, tl2 = torch.zeros((E2, 4), device=dev), torch.zeros(E2, device=dev, dtype=torch.int32)
        for j in range(4):
            eng.render(hm_b if j % 2 else hm_a, None, out=rgb2, depth_out=dep2)
            eng.fots_markers(dep2, th_b if j % 2 else th_a, tr2, tl2, out=mk2)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for j in range(2 * K):
            eng.render(hm_b if j % 2 else hm_a, None, out=rgb2, depth_out=dep2)
            eng.fots_markers(dep2, th_b if j % 2 else th_a, tr2, tl2, out=mk2)
        b.record()
        torch.cuda.synchronize()
        dms = a.elapsed_time(b) / (2 * K)
        extras["value_config2"] = {"value": E2 / (dms / 1e3), "unit": "frames/s per GPU", "ms_per_step": dms, "envs": E2,
                                   "workload": "config 2: 1024 envs, random primitive indenters (sphere / flat cylinder / 90 deg wedge / 60 deg cone), "
                                               f"random pose + yaw, RGB + FOTS {M}-marker motion along a shear / twist trajectory (steps alternate between "
                                               "the first-contact and the moved pose); 128 unique maps tiled"}
        del hm_a, hm_b

        # the reference's RL tasks render the tactile image at 32 x 32 (ball_rolling_tactile_rgb.py:303-318): the arbitrary-resolution
        # kernel, every blur sigma scaled with the shape as the reference does (any background serves a throughput figure)
        try:
            import dataclasses

            Hl = Wl = 32
            bg = torch.nn.functional.interpolate(tables.background[None], size=[Hl, Wl], mode="bilinear", antialias=True)[0].contiguous()
            tl = dataclasses.replace(tables, shape=(Hl, Wl), background=bg, gel_map=None)
            engl = TactileEngine(tl, max_envs=E, device=dev)
            hml = synth.lowres_batch(min(E, 256), Hl, Wl, seed=7 + rank)
            hml = hml.repeat((E + hml.shape[0] - 1) // hml.shape[0], 1, 1)[:E].contiguous().to(dev)
            rgbl, depl = torch.empty((E, Hl, Wl, 3), device=dev), torch.empty(E, device=dev)
            for _ in range(3):
                engl.render(hml, None, out=rgbl, depth_out=depl)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(K):
                engl.render(hml, None, out=rgbl, depth_out=depl)
            b.record()
            torch.cuda.synchronize()
            dms = a.elapsed_time(b) / K
            extras["value_lowres_32x32"] = {"value": E / (dms / 1e3), "unit": "frames/s per GPU", "ms_per_step": dms,
                                            "workload": f"{E} envs, tactile image 32 x 32 (the RL tasks' resolution): indentation depth + Taxim RGB "
                                                        "through taxim_generic_kernel (one CTA per frame)"}
            del engl, hml, rgbl
        except Exception as ex:  # an extra must never take the contract line down
            extras["value_lowres_32x32"] = {"error": repr(ex)[:200]}

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
    del rgb_host

    # ---- config 3 (one GPU: FEM substep alone + full step) and config 4 (every N: E4 envs per GPU, gel FEM substep + FEM markers
    #      + Taxim RGB + FOTS markers + observation gather; 8 x 1024 = the north star's 8192 envs) ------------------------
    fem_extra, config4 = None, None
    if not args.no_fem:
        if world == 1:
            fem_extra = fem_gpu(min(E, 4096), dev, tactile=(eng, rgb, depth))
        E4 = min(args.config4_envs, E)
        c3 = Config3(E4, dev, seed=2 + rank)
        eng4 = TactileEngine(tables, max_envs=E4, device=dev, marker_rows=MARKER_ROWS, marker_cols=MARKER_COLS)
        K4 = c3.last - c3.first  # 6 timed steps = press steps 24..29; the depth map of step s is the pose at its END (s + 1)
        maps_of = lambda c: [c.height_maps(s + 1) for s in range(c.first, c.last)]  # noqa: E731
        args4 = args
        if world > 2 and args.obs_gather == "fp32":
            # With the FEM substep in the step the SEPARATE rectangle push wins beyond two GPUs: it runs on the side stream under the
            # next step's FEM kernel, while the fused epilogue stores stall the (short) render on the link instead
            # (8 GPUs, 8 x 1024 envs: 678 k vs 642 k frames/s, profiles/r02_bench_n8_fp32.json vs r02_bench_n8_fp32_fused.json)
            import copy

            args4 = copy.copy(args)
            args4.obs_gather = "fp32-rect"
        run4 = StepRunner(args4, eng4, E4, dev, world, rank, maps_of(c3), fem_ctx=FemCtx(E4, dev, c3))
        for _ in range(c3.first - K4):   # untimed run-in of the press (FEM only), then K4 full warm-up steps
            run4.fem.step()
        ms4 = timed(run4, K4, K4)
        ok4 = None
        if run4.do_gather:
            ok4 = run4.verify_gather(lambda r: maps_of(Config3(E4, dev, seed=2 + r)))
            f4 = torch.tensor([1.0 if ok4 else 0.0], device=dev)
            dist.all_reduce(f4, op=dist.ReduceOp.MIN)
            ok4 = bool(f4.item())
        t4 = torch.tensor([ms4], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t4, op=dist.ReduceOp.MAX)
        ms4 = float(t4.item())
        config4 = {"workload": f"config 4: {world} x {E4} envs, per step: gel FEM substep (config-3 box edge / corner press, steps "
                               f"{c3.first}..{c3.last - 1} of 30) + FEM marker read-out + Taxim RGB 320x240 + FOTS {M}-marker motion"
                               + (" + observation all-gather" if run4.do_gather else ""),
                   "frames_per_s": world * E4 * K4 / (ms4 / 1e3), "ms_per_step": ms4 / K4, "global_envs": world * E4,
                   "obs_gather": run4.gather_kind if run4.do_gather else "n/a", "gather_verified": ok4}

    # ---- the reference's own CUDA path on this GPU, rank 0 of a 1-GPU run only ------------
    reference_cuda = None
    if not args.no_extras and world == 1:
        try:
            from baseline import ref_arm

            if ref_arm.available():
                del run
                torch.cuda.empty_cache()
                reference_cuda = {"what": "the UNMODIFIED reference TaximTorch.render_direct(with_shadow=False) + NHWC copy with device='cuda' "
                                          "on this GPU, same depth maps, CUDA-event timed (RGB only, no markers)",
                                  "batches": ref_arm.time_cuda(hm, (256, 1024, 4096))}
        except Exception as exc:
            reference_cuda = {"error": f"{type(exc).__name__}: {str(exc)[:200]}"}

    # ---- max over ranks ------------------------------------------------------------------------------------------
    t = torch.tensor([ms, kern_ms, e2e_s, 0.0 if gather_ok is False else 1.0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t[:3], op=dist.ReduceOp.MAX)
        dist.all_reduce(t[3:], op=dist.ReduceOp.MIN)
    ms, kern_ms, e2e_s, gok = t.tolist()

    if rank == 0:
        peak, peak_src = _peaks()
        frames = world * E * K
        value = frames / (ms / 1e3)
        ach = ALGO_BYTES_PER_FRAME * E / (kern_ms / 1e3) / 1e9
        traffic, traffic_src = None, None
        tp = ROOT / "profiles" / "traffic.json"
        if tp.exists():
            tj = json.loads(tp.read_text())
            if tj.get("dram_bytes_per_frame"):
                traffic = tj["dram_bytes_per_frame"] * E
                traffic_src = tj.get("source", "profiles/traffic.json")
        clocks = sampler.summary()
        clk = (clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0) * 1e6
        ops = fp32_lane_ops_per_frame(hm_host[:64])
        lanes = ops * E / (kern_ms / 1e3) / (148 * clk)
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {
                "workload": f"{E} envs/GPU x 320x240: indentation depth + Taxim RGB + FOTS {M}-marker motion, sphere indenters "
                            f"(config-1 distribution, 10% no contact); dense contacts under value_dense, the gel FEM substep (config 3) under "
                            f"fem_gel_substep, config 4 (FEM + RGB + markers + gather) under config4",
                "envs_per_gpu": E, "global_envs": world * E, "parallelism": f"dp{world} (contiguous env shards)",
                "l2_policy": f"inputs larger than L2 ({E * FRAME_IN_BYTES / 1e6:.0f} MB in + {E * FRAME_OUT_BYTES / 1e6:.0f} MB out per step vs 126 MB L2); three input sets cycle, so every env sees a different depth map in consecutive steps",
                "obs_gather": gk,
            },
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                         "traffic_source": traffic_src,
                         "kernel": "taxim_fused_kernel", "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": ALGO_BYTES_PER_FRAME * E,
                         "peak_source": peak_src,
                         "fp32": {"what": "second ceiling (SURVEY 8d): FP32 lane-ops the exact separable pyramid needs for these frames "
                                          "(zero-skipped regions excluded, colour stage not counted) per clock and SM",
                                  "lane_ops_per_frame": ops, "dense_lane_ops_per_frame": 266 * H * W,
                                  "achieved_lanes_per_clk_sm": lanes, "peak_lanes_per_clk_sm": 128, "measured_peak_lanes_per_clk_sm": 125,
                                  "frac": lanes / 125, "sm_clock_mhz": clk / 1e6},
                         "note": "HBM is the nominal bound; the exact pyramid (266 lane-ops/px) caps a dense evaluation at ~30% of it: see DESIGN.md section 4.1"},
            "e2e": {"value": world * E / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": E * (FRAME_IN_BYTES + 4),
                    "d2h_bytes_per_step": E * (FRAME_OUT_BYTES + 4 + M * 16), "api": "tx_step_host (C ABI, pinned host buffers)", "host_binding": numa,
                    "pcie_gbs_d2h": E * (FRAME_OUT_BYTES + 4 + M * 16) / e2e_s / 1e9},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if world > 1:
            line["gather_verified"] = (bool(gok) if gather_ok is not None else None)
        line.update(extras)
        if reference_cuda is not None:
            line["reference_cuda"] = reference_cuda
        if not args.no_cpu_baseline:
            if ALL_CPUS:
                os.sched_setaffinity(0, ALL_CPUS)  # the CPU baseline uses every host thread, not just the GPU's NUMA node
            r = cpu_reference_fps(20, warmup=1)
            if r is not None:
                fps, cores, sample = r
                line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "reference", "sample": sample}
                pf, pc, ps = cpu_port_fps(256)
                line["cpu_baseline"]["oracle_port"] = {"value": pf, "unit": "frames/s", "cores": pc, "kind": "port", "sample": ps}
            else:
                fps, cores, sample = cpu_port_fps(256)
                line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample}
        if fem_extra is not None:
            line["fem_gel_substep"] = fem_extra
            if not args.no_cpu_baseline:
                g, cores, sample = fem_cpu_port()
                line["fem_gel_substep"]["cpu_baseline"] = {"value": g, "unit": "gel-steps/s", "cores": cores, "kind": "port",
                                                           "sample": sample}
        if config4 is not None:
            line["config4"] = config4
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


def gather_kind_of(world, args, run) -> str:
    if world == 1:
        return "n/a"
    if run is not None and run.do_gather:
        return run.gather_kind
    return "none (observations stay sharded)"


if __name__ == "__main__":
    main()
