"""Gel-pad tetrahedral mesh and material set-up for the batched FEM substep (host side, init time).

The reference takes the gel mesh from a USD asset with precomputed ``tet_points/tet_indices`` attributes
(ref: source/tacex_uipc/tacex_uipc/objects/uipc_object.py:153-156) -- git-LFS pointers in the reference checkout -- so
the benchmark uses the structured box of SURVEY.md section 8(d) config 3: 20.75 x 25.25 x 4.5 mm (GelSight Mini gel pad,
ref: source/tacex_assets/tacex_assets/sensors/gelsight_mini/gsmini_cfg.py:22-24), 10 x 12 x 3 cells, 6 Freudenthal tets
per cell = 2160 tets, 11 x 13 x 4 = 572 vertices. Material defaults follow UipcObject (uipc_object.py:59-84, 442-470):
E = 0.01 MPa, nu = 0.49, rho = 1e3 kg/m^3; Lame parameters as ``ElasticModuli.youngs_poisson``
(ref: libuipc/src/constitution/elastic_moduli.cpp:20-27). The bottom layer is soft-attached to the sensor case
(ref: source/tacex_uipc/tacex_uipc/sim/uipc_attachments.py:118-142, strength ratio 1000 in the benchmark scene).
"""

from __future__ import annotations

import itertools
from dataclasses import dataclass

import numpy as np


@dataclass
class GelMesh:
    X: np.ndarray          # (V, 3) float64 rest positions [m], z = 0 bottom (case side), z = height top (contact side)
    tets: np.ndarray       # (T, 4) int32, positive orientation
    attach: np.ndarray     # (A,) int32 vertices soft-attached to the case (bottom layer)
    surf: np.ndarray       # (S,) int32 vertices that can touch the indenter (boundary minus the attached layer)
    top_tris: np.ndarray   # (F, 3) int32 triangles of the top surface (marker read-out)
    dims: tuple            # (nx, ny, nz) cells


def lame(youngs: float, poisson: float) -> tuple[float, float]:
    """(lambda, mu) from Young's modulus and Poisson ratio."""
    lam = youngs * poisson / ((1 + poisson) * (1 - 2 * poisson))
    mu = youngs / (2 * (1 + poisson))
    return lam, mu


def box_gel(size=(20.75e-3, 25.25e-3, 4.5e-3), cells=(10, 12, 3)) -> GelMesh:
    nx, ny, nz = cells
    xs = np.linspace(-size[0] / 2, size[0] / 2, nx + 1)
    ys = np.linspace(-size[1] / 2, size[1] / 2, ny + 1)
    zs = np.linspace(0.0, size[2], nz + 1)
    vid = lambda i, j, k: (i * (ny + 1) + j) * (nz + 1) + k  # noqa: E731
    X = np.zeros(((nx + 1) * (ny + 1) * (nz + 1), 3))
    for i, j, k in itertools.product(range(nx + 1), range(ny + 1), range(nz + 1)):
        X[vid(i, j, k)] = (xs[i], ys[j], zs[k])
    tets = []
    for i, j, k in itertools.product(range(nx), range(ny), range(nz)):
        for perm in itertools.permutations(range(3)):  # Freudenthal: 6 tets around the main diagonal
            p = [i, j, k]
            chain = [vid(*p)]
            for a in perm:
                p = list(p)
                p[a] += 1
                chain.append(vid(*p))
            t = chain
            Dm = np.stack([X[t[1]] - X[t[0]], X[t[2]] - X[t[0]], X[t[3]] - X[t[0]]], 1)
            if np.linalg.det(Dm) < 0:
                t = [t[0], t[2], t[1], t[3]]
            tets.append(t)
    tets = np.asarray(tets, np.int32)
    attach = np.asarray([vid(i, j, 0) for i in range(nx + 1) for j in range(ny + 1)], np.int32)
    boundary = set()
    for i, j, k in itertools.product(range(nx + 1), range(ny + 1), range(nz + 1)):
        if k > 0 and (i in (0, nx) or j in (0, ny) or k == nz):
            boundary.add(vid(i, j, k))
    surf = np.asarray(sorted(boundary), np.int32)
    tris = []
    for i, j in itertools.product(range(nx), range(ny)):
        a, b, c, d = vid(i, j, nz), vid(i + 1, j, nz), vid(i + 1, j + 1, nz), vid(i, j + 1, nz)
        tris += [[a, b, c], [a, c, d]]
    return GelMesh(X, tets, attach, surf, np.asarray(tris, np.int32), cells)
