"""ctypes front-end of oracle/fem_canon.c (float64 CPU restatement of the gel FEM substep) -- TEST INFRASTRUCTURE ONLY."""

from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = _HERE / "_build" / "libfem_canon.so"


def build(force: bool = False) -> Path:
    src = _HERE / "fem_canon.c"
    if force or not _LIB.exists() or _LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "_build/libfem_canon.so"], check=True, capture_output=True)
    return _LIB


class FemCfg(C.Structure):
    _fields_ = [("V", C.c_int), ("T", C.c_int), ("A", C.c_int), ("S", C.c_int), ("dt", C.c_double),
                ("gravity", C.c_double * 3), ("mu", C.c_double), ("lam", C.c_double), ("attach_strength", C.c_double),
                ("d_hat", C.c_double), ("kappa", C.c_double), ("newton_max_iter", C.c_int), ("velocity_tol", C.c_double),
                ("pcg_tol_rate", C.c_double), ("pcg_max_iter_ratio", C.c_int), ("ls_max_iter", C.c_int),
                ("substep", C.c_int), ("friction_mu", C.c_double), ("eps_velocity", C.c_double)]


class FemIndenter(C.Structure):
    _fields_ = [("type", C.c_int), ("c", C.c_double * 3), ("R", C.c_double * 9), ("h", C.c_double * 3)]


class FemStats(C.Structure):
    _fields_ = [("converged", C.c_int), ("newton_iters", C.c_int), ("pcg_iters", C.c_int), ("ls_halvings", C.c_int),
                ("min_dist", C.c_double), ("last_res", C.c_double), ("energy", C.c_double)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB))
        _lib.fem_pcg_solve.restype = C.c_int
    return _lib


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def make_indenter(kind: int, center, half, R=None) -> FemIndenter:
    ind = FemIndenter()
    ind.type = kind
    ind.c[:] = list(center)
    ind.R[:] = list(np.eye(3).ravel() if R is None else np.asarray(R, float).ravel())
    ind.h[:] = list(half)
    return ind


class CanonFem:
    def __init__(self, mesh, youngs=1e4, poisson=0.49, density=1e3, dt=0.01, gravity=(0, 0, -9.8), attach_strength=1000.0,
                 d_hat=5e-4, kappa=1e10, newton_max_iter=1024, velocity_tol=0.05, rest_volume_det=True, substep=1,
                 friction_mu=0.5, eps_velocity=0.01):
        from tacex_b200.gel_mesh import lame

        self.mesh = mesh
        lam, mu = lame(youngs, poisson)
        g = FemCfg()
        g.V, g.T, g.A, g.S = len(mesh.X), len(mesh.tets), len(mesh.attach), len(mesh.surf)
        g.dt = dt
        g.gravity[:] = gravity
        g.mu, g.lam, g.attach_strength, g.d_hat, g.kappa = mu, lam, attach_strength, d_hat, kappa
        g.newton_max_iter, g.velocity_tol, g.pcg_tol_rate, g.pcg_max_iter_ratio, g.ls_max_iter, g.substep = (
            newton_max_iter, velocity_tol, 1e-3, 2, 8, substep)
        g.friction_mu, g.eps_velocity = friction_mu, eps_velocity
        self.cfg = g
        self.X = np.ascontiguousarray(mesh.X, np.float64)
        self.tets = np.ascontiguousarray(mesh.tets, np.int32)
        self.attach = np.ascontiguousarray(mesh.attach, np.int32)
        self.surf = np.ascontiguousarray(mesh.surf, np.int32)
        self.Dm_inv = np.empty((g.T, 9))
        self.vol = np.empty(g.T)
        self.mass = np.empty(g.V)
        lib().fem_precompute(g.V, g.T, _d(self.X), _i(self.tets), C.c_double(density), int(rest_volume_det), _d(self.Dm_inv),
                             _d(self.vol), _d(self.mass))

    def new_state(self, N: int):
        x = np.tile(self.X[None], (N, 1, 1)).copy()
        return x, np.zeros_like(x), x.copy()

    @staticmethod
    def set_indenter_mesh(tri_local) -> None:
        """Triangles (n, 3, 3) float64 of the prescribed mesh indenter (type 2) in its own frame; process-wide like the C global."""
        t = np.ascontiguousarray(tri_local, np.float64).reshape(-1, 9)
        lib().fem_set_indenter_mesh(_d(t), C.c_int(len(t)))

    def set_contact_surface(self, tris) -> None:
        """Triangles (n, 3) int32 of the gel surface the indenter mesh's VERTICES (against the triangles) and EDGES (against the
        triangles' edges) can touch; ``None`` switches both off. Process-wide like the C global."""
        if tris is None:
            lib().fem_set_contact_surface(None, C.c_int(0), None)
            return
        t = np.ascontiguousarray(tris, np.int32).reshape(-1, 3)
        lib().fem_set_contact_surface(_i(t), C.c_int(len(t)), _d(self.X))

    def step(self, x, v, x_prev, aim, ind_prev, ind_next):
        """x, v, x_prev: (N, V, 3) float64 updated in place; aim (N, A, 3); indenters: lists of FemIndenter."""
        N = x.shape[0]
        st = (FemStats * N)()
        ip = (FemIndenter * N)(*ind_prev)
        inx = (FemIndenter * N)(*ind_next)
        aim = np.ascontiguousarray(aim, np.float64)
        lib().fem_step_batch(C.byref(self.cfg), _i(self.tets), _d(self.Dm_inv), _d(self.vol), _d(self.mass), _i(self.attach),
                             _i(self.surf), _d(aim), ip, inx, N, _d(x), _d(v), _d(x_prev), st)
        return [{k: getattr(s, k) for k, _ in FemStats._fields_} for s in st]
