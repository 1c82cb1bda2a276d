"""Timing of the batched gel FEM substep (config 3: box indenter pressed 0 -> 1 mm over 30 steps).

    python tools/fem_time.py [N] [steps] [mesh_kind]     mesh_kind 1 / 2 / 3: the config-2 cylinder / wedge / cone as a prescribed triangle mesh; 4: a 1280-triangle icosphere
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tacex_b200 import fem, gel_mesh
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
m = gel_mesh.box_gel()
eng = fem.GelFemEngine(m)
rng = np.random.default_rng(2)
offs = rng.uniform(-1, 1, (N, 2)) * np.array([6e-3, 8e-3])
half = (2e-3, 3e-3, 1e-3)
z0 = 4.5e-3 + half[2] + 4e-4
x, v, xp = eng.new_state(N); aim = eng.rest_aim(N)
ctr = lambda s: np.concatenate([offs, np.full((N, 1), z0 - 1e-3 * s / 30)], 1)
mesh_kind = int(sys.argv[3]) if len(sys.argv) > 3 else 0
if mesh_kind:
    from tacex_b200 import synth
    eng.set_indenter_mesh(synth.indenter_mesh(mesh_kind % 4, 3e-3))
    if os.environ.get("TX_TP"):  # second half of the vertex-face contact: indenter vertices against the gel's top triangles
        eng.set_contact_surface(m.top_tris)
    z0 = 4.5e-3 + 4e-4
    inds = [fem.indenter_array(2, ctr(s), (0, 0, 0)) for s in range(steps + 1)]
else:
    inds = [fem.indenter_array(1, ctr(s), half) for s in range(steps + 1)]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tot_newton = 0; tot_pcg = 0
for s in range(steps):
    e0.record(); st = eng.step(x, v, xp, aim, inds[s], inds[s + 1]); e1.record(); torch.cuda.synchronize()
    d = eng.decode_stats(st)
    nn = np.mean([q["newton_iters"] for q in d]); pc = np.mean([q["pcg_iters"] for q in d]); cv = np.mean([q["converged"] for q in d])
    print(f"step {s}: {e0.elapsed_time(e1):8.2f} ms  newton {nn:.2f}  pcg {pc:.1f}  converged {cv:.2f}  min_dist {min(q['min_dist'] for q in d):.2e}")
cyc = torch.zeros((148, 6), dtype=torch.int64, device="cuda")
eng.lib.tx_fem_debug_set_cycles(eng.h, cyc.data_ptr())
st = eng.step(x, v, xp, aim, inds[steps - 1], inds[steps]); torch.cuda.synchronize()
eng.lib.tx_fem_debug_set_cycles(eng.h, None)
c = cyc.double().mean(0).tolist()
per = N / 148
print("cycles per gel-step (mean over CTAs): " + "  ".join(f"{n} {v / per:9.0f}" for n, v in zip(["assembly", "pcg", "linesearch", "asm:tets", "asm:rows", "asm:edges"], c)))
print(f"-> {N / (e0.elapsed_time(e1) / 1e3):.0f} gel-steps/s at the last step")
