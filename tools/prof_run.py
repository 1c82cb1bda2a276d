"""Small driver for ncu captures: a few launches of the fused kernel on resident inputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tacex_b200 import synth
from tacex_b200.calib import TaximTables
from tacex_b200.engine import TactileEngine
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
E = int(sys.argv[1]) if len(sys.argv) > 1 else 592
t = TaximTables.load(ROOT + "/tests/golden/gsmini_tables_320x240.npz")
eng = TactileEngine(t, max_envs=E, marker_rows=7, marker_cols=9)
KIND = sys.argv[2] if len(sys.argv) > 2 else "sparse"
hm = (synth.dense_batch(E, n_unique=16) if KIND == "dense" else synth.bench_batch(E, n_unique=64)).cuda()
rgb = torch.empty((E, 240, 320, 3), device="cuda"); dep = torch.empty(E, device="cuda")
for _ in range(4):
    eng.render(hm, None, out=rgb, depth_out=dep)
torch.cuda.synchronize()
