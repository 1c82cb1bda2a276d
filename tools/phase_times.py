"""Per-phase cycle counts of the fused Taxim kernel (clock64 stamps written by thread 0 of every CTA)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tacex_b200 import synth
from tacex_b200.calib import TaximTables
from tacex_b200.engine import TactileEngine
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
E = int(sys.argv[1]) if len(sys.argv) > 1 else 74
FLAGS = int(sys.argv[2]) if len(sys.argv) > 2 else 0
t = TaximTables.load(ROOT + "/tests/golden/gsmini_tables_320x240.npz")
eng = TactileEngine(t, max_envs=E)
eng.set_debug_flags(FLAGS)
KIND = sys.argv[3] if len(sys.argv) > 3 else "sparse"
hm = (synth.dense_batch(E, n_unique=16) if KIND == "dense" else synth.bench_batch(E, n_unique=64)).cuda()
rgb = torch.empty((E, 240, 320, 3), device="cuda")
for _ in range(3):
    eng.render(hm, None, out=rgb)
ticks = torch.zeros((2 * E, 40), dtype=torch.int64, device="cuda")
eng.set_phase_ticks(ticks)
eng.render(hm, None, out=rgb)
torch.cuda.synchronize()
eng.set_phase_ticks(None)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    eng.render(hm, None, out=rgb)
e1.record(); torch.cuda.synchronize()
print(f"kernel time {e0.elapsed_time(e1)/5*1e3:.1f} us for {E} frames -> {E/(e0.elapsed_time(e1)/5e3):.0f} frames/s")
tk = ticks.cpu().double()
press = eng.indentation_depth(hm).cpu()
act = (press > 0).repeat_interleave(2)
tk = tk[act]
names = {1: "tma load", 2: "min+xchg", 3: "h/mask pass"}
for L in range(7):
    names[4 + 4 * L] = f"L{L} hpass"; names[5 + 4 * L] = f"L{L} cluster.sync"; names[6 + 4 * L] = f"L{L} vpass"; names[7 + 4 * L] = f"L{L} reimpose"
names[32] = "halo+aux+sync"; names[33] = "epilogue"
prev = 0; tot = (tk[:, 33] - tk[:, 0]).mean().item()
print(f"CTAs in contact: {tk.shape[0]}, total cycles/CTA {tot:.0f}")
for k in sorted(names):
    d = (tk[:, k] - tk[:, prev]).mean().item(); prev = k
    print(f"{names[k]:18s} {d:10.0f}  {100*d/tot:5.1f}%")
