"""Small driver for ncu captures of the gel FEM substep: a few steps of 148 gels (one per SM).

    python tools/fem_prof_run.py [N] [mesh_kind]     mesh_kind 2 / 3: wedge / cone mesh with the full simplex contact (fem_step_kernel<true>)
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tacex_b200 import fem, gel_mesh
N = int(sys.argv[1]) if len(sys.argv) > 1 else 148
m = gel_mesh.box_gel()
eng = fem.GelFemEngine(m)
rng = np.random.default_rng(2)
offs = rng.uniform(-1, 1, (N, 2)) * np.array([6e-3, 8e-3])
half = (2e-3, 3e-3, 1e-3)
z0 = 4.5e-3 + half[2] + 4e-4
x, v, xp = eng.new_state(N); aim = eng.rest_aim(N)
ctr = lambda s: np.concatenate([offs, np.full((N, 1), z0 - 1e-3 * s / 30)], 1)
mk = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if mk:
    from tacex_b200 import synth
    eng.set_indenter_mesh(synth.indenter_mesh(mk, 3e-3))
    eng.set_contact_surface(m.top_tris)
    z0 = 4.5e-3 + 4e-4
    for s in range(12):
        eng.step(x, v, xp, aim, fem.indenter_array(2, ctr(s), (0, 0, 0)), fem.indenter_array(2, ctr(s + 1), (0, 0, 0)), want_stats=False)
else:
    for s in range(6):
        eng.step(x, v, xp, aim, fem.indenter_array(1, ctr(s), half), fem.indenter_array(1, ctr(s + 1), half), want_stats=False)
torch.cuda.synchronize()
