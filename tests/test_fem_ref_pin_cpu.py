"""Pins the float64 FEM restatement (oracle/fem_canon.c) against the REFERENCE'S OWN SOURCE, compiled unmodified from where it
lies under /root/reference (libuipc cuda backend; oracle/Makefile target `ref` -> oracle/_ref/libuipc_sym.so, see
oracle/ref_sym.cpp): the stable Neo-Hookean energy / gradient / Hessian (sym/stable_neo_hookean_3d.inl), the IPC barrier
(sym/codim_ipc_contact.inl), and the vertex-vs-half-plane normal and frictional contact (ipc_vertex_half_plane_contact_function.h,
codim_ipc_contact_function.h) that the restatement generalises to analytic indenters: under a flat face of a box indenter the
two models must coincide. A second library (oracle/ref_dist.cpp -> oracle/_ref/libuipc_dist.so) holds the reference's closest-feature
distances (utils/distance/distance_flagged.h, point-triangle and edge-edge), its edge-edge mollifier and its additive CCD
(details/ccd.inl): they pin the mesh-indenter contact of the restatement, the CCD also on the fixtures with ground truth that muda
ships. Skipped where the libraries have not been built (they need /root/reference; the GPU box has none)."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

from oracle import fem_canon as fc

REF = Path(fc.__file__).resolve().parent / "_ref" / "libuipc_sym.so"
pytestmark = pytest.mark.skipif(not REF.exists(), reason="oracle/_ref/libuipc_sym.so not built (make -C oracle ref; needs /root/reference)")

DP = C.POINTER(C.c_double)


def _d(a):
    return a.ctypes.data_as(DP)


@pytest.fixture(scope="module")
def ref():
    return C.CDLL(str(REF))


@pytest.fixture(scope="module")
def canon():
    return fc.lib()


def _rot(rng):
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q


def test_stable_neo_hookean_matches_reference_source(ref, canon):
    rng = np.random.default_rng(0)
    mu, lam = 3355.7, 164429.5  # E = 10 kPa, nu = 0.49
    for k in range(20):
        F = (np.eye(3) + (0.05 if k < 10 else 0.6) * rng.standard_normal((3, 3))).T.reshape(-1).copy()
        E1, g1, H1 = C.c_double(), np.empty(9), np.empty(81)
        E2, g2, H2 = C.c_double(), np.empty(9), np.empty(81)
        canon.fem_snh(_d(F), C.c_double(mu), C.c_double(lam), C.byref(E1), _d(g1), _d(H1))
        ref.ref_snh(_d(F), C.c_double(mu), C.c_double(lam), C.byref(E2), _d(g2), _d(H2))
        assert abs(E1.value - E2.value) <= 1e-12 * max(abs(E2.value), 1.0)
        assert np.abs(g1 - g2).max() <= 1e-11 * np.abs(g2).max()
        assert np.abs(H1 - H2).max() <= 1e-11 * np.abs(H2).max()


def test_barrier_matches_reference_source(ref, canon):
    d_hat, kappa = 5e-4, 1e10 * 1e-4
    for d in np.geomspace(1e-7, d_hat * 0.999, 25):
        B1, dB1, ddB1 = C.c_double(), C.c_double(), C.c_double()
        B2, dB2, ddB2 = C.c_double(), C.c_double(), C.c_double()
        canon.fem_barrier(C.c_double(d * d), C.c_double(d_hat), C.c_double(kappa), C.byref(B1), C.byref(dB1), C.byref(ddB1))
        ref.ref_kappa_barrier(C.c_double(kappa), C.c_double(d * d), C.c_double(d_hat), C.c_double(0.0), C.byref(B2), C.byref(dB2),
                              C.byref(ddB2))
        for a, b in ((B1, B2), (dB1, dB2), (ddB1, ddB2)):
            assert abs(a.value - b.value) <= 1e-9 * abs(b.value) + 1e-300


def _cfg(friction_mu=0.5):
    g = fc.FemCfg()
    g.dt, g.d_hat, g.kappa, g.friction_mu, g.eps_velocity = 0.01, 5e-4, 1e10, friction_mu, 0.01
    return g


def _face_case(rng, rotated):
    """A box indenter, a vertex under its bottom face at distance d < d_hat, and the reference's half-plane (P, N) of that face."""
    Rm = _rot(rng) if rotated else np.eye(3)
    c = rng.uniform(-1e-2, 1e-2, 3)
    half = np.array([4e-3, 5e-3, 2e-3])
    ind = fc.make_indenter(1, c, half, R=Rm)
    N = -Rm[:, 2]
    P = c - half[2] * Rm[:, 2]
    d = rng.uniform(0.02, 0.95) * 5e-4
    loc = np.array([rng.uniform(-0.8, 0.8) * half[0], rng.uniform(-0.8, 0.8) * half[1], -half[2] - d])
    x = c + Rm @ loc
    return ind, N.copy(), P.copy(), x.copy(), d


@pytest.mark.parametrize("rotated", [False, True])
def test_vertex_barrier_under_a_flat_face_is_the_reference_half_plane_model(ref, canon, rotated):
    rng = np.random.default_rng(1)
    g = _cfg()
    kap = g.kappa * (g.dt * g.dt)
    canon.fem_vertex_barrier_terms.restype = C.c_int
    for _ in range(20):
        ind, N, P, x, d = _face_case(rng, rotated)
        E1, G1, H1 = C.c_double(), np.empty(3), np.empty(9)
        assert canon.fem_vertex_barrier_terms(C.byref(g), C.byref(ind), _d(x), C.byref(E1), _d(G1), _d(H1)) == 1
        E2, G2, H2 = C.c_double(), np.empty(3), np.empty(9)
        ref.ref_ph_barrier(C.c_double(kap), C.c_double(g.d_hat), C.c_double(0.0), _d(x), _d(P), _d(N), C.byref(E2), _d(G2), _d(H2))
        # the distance itself carries ~1e-16 / d of relative rounding (d down to 1e-5 m inside coordinates of 1e-2 m)
        tol = 1e-8
        assert abs(E1.value - E2.value) <= tol * abs(E2.value)
        assert np.abs(G1 - G2).max() <= tol * np.abs(G2).max()
        assert np.abs(H1 - H2).max() <= tol * np.abs(H2).max()


@pytest.mark.parametrize("rotated", [False, True])
def test_lagged_friction_under_a_flat_face_is_the_reference_half_plane_model(ref, canon, rotated):
    rng = np.random.default_rng(2)
    g = _cfg()
    kap, eps = g.kappa * g.dt * g.dt, g.eps_velocity * g.dt
    n_stick = n_slip = 0
    for k in range(40):
        ind, N, P, xp, d = _face_case(rng, rotated)
        # stick (|u| < eps) and slip (|u| > eps) displacements, mostly tangential
        scale = eps * (0.3 if k % 2 == 0 else 5.0)
        x = xp + scale * rng.standard_normal(3) * np.array([1.0, 1.0, 0.05])
        e1, e2 = np.empty(3), np.empty(3)
        ref.ref_tan_basis(_d(N), _d(e1), _d(e2))
        u2 = ((x - xp) @ e1) ** 2 + ((x - xp) @ e2) ** 2
        n_stick += u2 < eps * eps
        n_slip += u2 >= eps * eps
        E1, G1, H1 = C.c_double(), np.empty(3), np.empty(9)
        canon.fem_friction_terms(C.byref(g), C.byref(ind), C.byref(ind), _d(xp), _d(x), C.byref(E1), _d(G1), _d(H1))
        E2, G2, H2 = C.c_double(), np.empty(3), np.empty(9)
        ref.ref_ph_friction(C.c_double(kap), C.c_double(g.d_hat), C.c_double(0.0), C.c_double(g.friction_mu), C.c_double(eps),
                            _d(xp), _d(x), _d(P), _d(N), C.byref(E2), _d(G2), _d(H2))
        tol = 1e-8
        assert E2.value > 0
        assert abs(E1.value - E2.value) <= tol * abs(E2.value)
        assert np.abs(G1 - G2).max() <= tol * np.abs(G2).max()
        assert np.abs(H1 - H2).max() <= tol * np.abs(H2).max()
    assert n_stick >= 5 and n_slip >= 5


def test_lame_parameters_match_reference_frontend(ref):
    """ElasticModuli::youngs_poisson(E * MPa, nu) of the gel object (ref: tacex_uipc/objects/uipc_object.py:453-455,
    libuipc src/constitution/elastic_moduli.cpp:20-27 -> include/uipc/constitution/conversion.h EP_to_lame)."""
    from tacex_b200.gel_mesh import lame

    for E, nu in [(1e4, 0.49), (0.01e6, 0.49), (5e5, 0.3), (2e9, 0.0)]:
        lam_r, mu_r = C.c_double(), C.c_double()
        ref.ref_ep_to_lame(C.c_double(E), C.c_double(nu), C.byref(lam_r), C.byref(mu_r))
        lam, mu = lame(E, nu)
        assert abs(lam - lam_r.value) <= 1e-12 * max(abs(lam_r.value), 1.0) and abs(mu - mu_r.value) <= 1e-12 * mu_r.value


def test_per_tet_assembly_matches_reference_fem_utils_and_constitution(ref, canon):
    """One tetrahedron: the Newton matrix and right-hand side of the restatement (inertia + dt^2 V dFdx^T spd(H9) dFdx, gradient
    dt^2 V dFdx^T dE/dF) against the reference's own Ds / Dm_inv / F / dFdx (finite_element/fem_utils.cu) combined with its
    stable Neo-Hookean gradient / Hessian (sym/stable_neo_hookean_3d.inl); make_spd (utils/make_spd.h: eigenvalue clamp) via
    numpy.linalg.eigh."""
    rng = np.random.default_rng(5)
    mu, lam, dt = 3355.7, 164429.5, 0.01
    for k in range(10):
        X = np.array([[0, 0, 0], [2e-3, 0, 0], [0, 2e-3, 0], [0, 0, 1.5e-3]]) + 2e-4 * rng.standard_normal((4, 3))
        if np.linalg.det((X[1:] - X[0]).T) < 0:
            X[[1, 2]] = X[[2, 1]]
        x = X + (2e-5 if k < 5 else 4e-4) * rng.standard_normal((4, 3))
        tets = np.array([[0, 1, 2, 3]], np.int32)
        g = fc.FemCfg()
        g.V, g.T, g.A, g.S, g.dt, g.mu, g.lam = 4, 1, 0, 0, dt, mu, lam
        g.d_hat, g.kappa, g.attach_strength, g.friction_mu, g.eps_velocity = 5e-4, 1e10, 0.0, 0.0, 0.01
        Dm_inv, vol, mass = np.empty(9), np.empty(1), np.empty(4)
        canon.fem_precompute(4, 1, _d(X), fc._i(tets), C.c_double(1e3), 1, _d(Dm_inv), _d(vol), _d(mass))
        xt = x + 1e-6 * rng.standard_normal((4, 3))
        A, b, E = np.zeros((12, 12)), np.zeros(12), C.c_double()
        far = fc.make_indenter(0, (0, 0, 1.0), (1e-3, 0, 0))
        none_i, none_d = np.zeros(1, np.int32), np.zeros(3)
        canon.fem_assemble_dense(C.byref(g), fc._i(tets), _d(Dm_inv), _d(vol), _d(mass), fc._i(none_i), fc._i(none_i), _d(none_d),
                                 C.byref(far), _d(x), _d(X), _d(xt), C.c_double(1.0), _d(A), _d(b), C.byref(E))
        # reference side
        Di, vf, P = np.empty(9), np.empty(9), np.empty(108)
        ref.ref_tet(_d(np.ascontiguousarray(X.reshape(-1))), _d(np.ascontiguousarray(x.reshape(-1))), _d(Di), _d(vf), _d(P))
        P = P.reshape(9, 12)
        assert np.abs(np.sort(np.abs(Dm_inv)) - np.sort(np.abs(Di))).max() <= 1e-9 * np.abs(Di).max()  # same matrix (layout aside)
        Er, gr, Hr = C.c_double(), np.empty(9), np.empty(81)
        ref.ref_snh(_d(vf), C.c_double(mu), C.c_double(lam), C.byref(Er), _d(gr), _d(Hr))
        w, Q = np.linalg.eigh(Hr.reshape(9, 9))
        Hspd = (Q * np.maximum(w, 0.0)) @ Q.T
        M = np.repeat(mass, 3)
        s = dt * dt * vol[0]
        G_ref = M * (x - xt).reshape(-1) + s * (P.T @ gr)
        A_ref = np.diag(M) + s * (P.T @ Hspd @ P)
        assert np.abs(-b - G_ref).max() <= 1e-9 * np.abs(G_ref).max()
        assert np.abs(A - A_ref).max() <= 1e-9 * np.abs(A_ref).max()


def test_block_jacobi_inverse_matches_muda_analytic_inverse(ref, canon):
    """The 3 x 3 inverse of the block-Jacobi preconditioner: the reference calls muda::eigen::inverse, i.e. muda's AnalyticalInverse
    (fem_diag_preconditioner.cu:142-148; external/muda/src/muda/ext/eigen/inverse/analytic_inverse.h, compiled unmodified), the
    restatement its own adjugate formula (fem_canon.c::inv3). SPD diagonal blocks as the assembly produces them (mass + stiffness,
    condition numbers up to ~1e6)."""
    rng = np.random.default_rng(5)
    for k in range(200):
        q = _rot(rng)
        ev = 10.0 ** rng.uniform(-6, 0, 3) * (1.0 if k % 2 else 1e4)
        A = (q * ev) @ q.T
        A = 0.5 * (A + A.T)
        r_ref, r_me = np.empty(9), np.empty(9)
        ref.ref_inverse3(_d(np.ascontiguousarray(A.T.reshape(-1))), _d(r_ref))  # column-major in, column-major out
        canon.canon_inv3(_d(np.ascontiguousarray(A.reshape(-1))), _d(r_me))     # row-major (symmetric: the same)
        R_ref, R_me = r_ref.reshape(3, 3).T, r_me.reshape(3, 3)
        scale = np.abs(R_ref).max()
        assert np.abs(R_ref - R_me).max() <= 1e-9 * scale
        assert np.abs(R_me @ A - np.eye(3)).max() <= 1e-6


# ---- closest-feature distances (triangle-mesh indenter) ---------------------------------------------------------------------------
REF_DIST = REF.parent / "libuipc_dist.so"
# reference flag (p, t0, t1, t2) -> kind of oracle/fem_canon.c::fem_pt_distance
_KIND = {(1, 1, 1, 1): 0, (1, 1, 1, 0): 1, (1, 0, 1, 1): 2, (1, 1, 0, 1): 3, (1, 1, 0, 0): 4, (1, 0, 1, 0): 5, (1, 0, 0, 1): 6}


@pytest.mark.skipif(not REF_DIST.exists(), reason="oracle/_ref/libuipc_dist.so not built (make -C oracle ref; needs /root/reference)")
def test_point_triangle_closest_feature_and_distance_derivatives_match_reference_source(canon):
    """fem_pt_distance vs the reference's point_triangle_distance_flag + flagged distance / gradient / Hessian
    (utils/distance/distance_flagged.h compiled from where it lies): the same closest feature for points all around random, thin and
    obtuse triangles, the same squared distance, and the p-block of the 12-gradient / 12 x 12 Hessian."""
    rd = C.CDLL(str(REF_DIST))
    canon.fem_pt_distance.restype = C.c_int
    canon.fem_pt_closest.restype = C.c_int
    rng = np.random.default_rng(7)
    IP = C.POINTER(C.c_int)
    seen = set()
    n_checked = 0
    for k in range(4000):
        t = rng.standard_normal((3, 3)) * (1e-3 if k % 2 else 1.0)
        if k % 5 == 0:  # thin / obtuse triangle
            t[2] = t[0] + (t[1] - t[0]) * rng.uniform(-0.5, 1.5) + 0.05 * rng.standard_normal(3) * np.linalg.norm(t[1] - t[0])
        e = np.linalg.norm(t[1] - t[0])
        # points spread over all seven regions: barycentric combination with weights outside [0, 1] + an offset along the normal
        w = rng.uniform(-1.0, 2.0, 3)
        w /= w.sum() if abs(w.sum()) > 0.2 else 1.0
        nrm = np.cross(t[1] - t[0], t[2] - t[0])
        p = w @ t + rng.uniform(-1, 1) * e * nrm / np.linalg.norm(nrm)
        t0, t1, t2 = (np.ascontiguousarray(t[i]) for i in range(3))
        p = np.ascontiguousarray(p)
        flag = np.zeros(4, np.int32)
        D2, G2, H2 = C.c_double(), np.zeros(12), np.zeros(144)
        rd.ref_pt(_d(p), _d(t0), _d(t1), _d(t2), flag.ctypes.data_as(IP), C.byref(D2), _d(G2), _d(H2))
        D1, G1, H1 = C.c_double(), np.zeros(3), np.zeros(9)
        kind = canon.fem_pt_distance(_d(p), _d(t0), _d(t1), _d(t2), C.byref(D1), _d(G1), _d(H1))
        assert kind == _KIND[tuple(int(f) for f in flag)], (k, kind, flag)
        seen.add(kind)
        sc = max(D2.value, 1e-300)
        assert abs(D1.value - D2.value) <= 1e-10 * sc
        assert np.abs(G1 - G2[:3]).max() <= 1e-9 * max(np.abs(G2[:3]).max(), np.sqrt(sc) * 1e-3)
        Hp = H2.reshape(12, 12)[:3, :3]
        assert np.abs(H1.reshape(3, 3) - Hp).max() <= 1e-8 * 2.0, (k, kind)
        # closest point form (the triangle's vertices as unknowns): the FULL 12-gradient is 2 r (x) [1, -w0, -w1, -w2]
        D3, r3, w3 = C.c_double(), np.zeros(3), np.zeros(3)
        assert canon.fem_pt_closest(_d(p), _d(t0), _d(t1), _d(t2), C.byref(D3), _d(r3), _d(w3)) == kind
        assert abs(w3.sum() - 1.0) <= 1e-12 and abs(D3.value - D2.value) <= 1e-10 * sc
        g12 = np.concatenate([2.0 * r3, -2.0 * w3[0] * r3, -2.0 * w3[1] * r3, -2.0 * w3[2] * r3])
        assert np.abs(g12 - G2).max() <= 1e-8 * max(np.abs(G2).max(), np.sqrt(sc) * 1e-3), (k, kind, g12, G2)
        n_checked += 1
    assert seen == set(range(7)) and n_checked == 4000


@pytest.mark.skipif(not REF_DIST.exists(), reason="oracle/_ref/libuipc_dist.so not built")
def test_per_candidate_barrier_block_projection_has_the_closed_form_the_kernel_uses(canon):
    """The CUDA kernel projects each candidate's 3 x 3 barrier block analytically: max(0, B'' + B' / (2 D)) g g^T. Checked against
    make_spd (numeric eigen-decomposition) of B'' g g^T + B' H_D built from the REFERENCE's gradient / Hessian p-blocks."""
    rd = C.CDLL(str(REF_DIST))
    rng = np.random.default_rng(11)
    IP = C.POINTER(C.c_int)
    d_hat, kappa = 5e-4, 1e10 * 1e-4
    for k in range(300):
        t = rng.standard_normal((3, 3)) * 2e-3
        w = rng.uniform(-0.6, 1.6, 3)
        w /= w.sum() if abs(w.sum()) > 0.2 else 1.0
        nrm = np.cross(t[1] - t[0], t[2] - t[0])
        p = np.ascontiguousarray(w @ t + rng.uniform(0.02, 0.98) * d_hat * nrm / np.linalg.norm(nrm))
        flag = np.zeros(4, np.int32)
        D2, G2, H2 = C.c_double(), np.zeros(12), np.zeros(144)
        rd.ref_pt(_d(p), _d(np.ascontiguousarray(t[0])), _d(np.ascontiguousarray(t[1])), _d(np.ascontiguousarray(t[2])),
                  flag.ctypes.data_as(IP), C.byref(D2), _d(G2), _d(H2))
        D = D2.value
        if not (0 < D < d_hat * d_hat):
            continue
        B, dB, ddB = C.c_double(), C.c_double(), C.c_double()
        canon.fem_barrier(C.c_double(D), C.c_double(d_hat), C.c_double(kappa), C.byref(B), C.byref(dB), C.byref(ddB))
        g = G2[:3]
        H = ddB.value * np.outer(g, g) + dB.value * H2.reshape(12, 12)[:3, :3]
        wv, V = np.linalg.eigh(0.5 * (H + H.T))
        Hproj = (V * np.maximum(wv, 0)) @ V.T
        closed = max(0.0, ddB.value + dB.value / (2 * D)) * np.outer(g, g)
        assert np.abs(Hproj - closed).max() <= 1e-9 * max(np.abs(Hproj).max(), 1e-30)


MUDA_VF = Path("/root/reference/source/tacex_uipc/libuipc/external/muda/test/data/unit-tests/vertex-face")


def _vf_queries():
    """muda's vertex-face CCD fixtures (rational coordinates num / den per axis + ground truth; 8 rows per query: the vertex and the
    three face vertices at t = 0, then at t = 1)."""
    out = []
    for f in sorted(MUDA_VF.glob("*.csv")):
        rows = np.loadtxt(f, delimiter=",", dtype=np.float64)
        pts = np.stack([rows[:, 0] / rows[:, 1], rows[:, 2] / rows[:, 3], rows[:, 4] / rows[:, 5]], 1).reshape(-1, 8, 3)
        truth = rows[:, 6].reshape(-1, 8)[:, 0].astype(bool)
        out += [(q, bool(t)) for q, t in zip(pts, truth)]
    return out


@pytest.mark.skipif(not (REF_DIST.exists() and MUDA_VF.exists()), reason="needs oracle/_ref/libuipc_dist.so and the reference tree")
def test_additive_ccd_matches_reference_source_on_mudas_vertex_face_fixtures(canon):
    """fem_pt_accd vs the reference's point_triangle_ccd (utils/distance/details/ccd.inl compiled from where it lies) on the 250
    vertex-face queries muda ships with ground truth: the same hit / miss decision and time of impact for every query, no collision
    of the ground truth is missed by either (ACCD is conservative), and the same again for random sweeps against a static triangle
    (the form the mesh-indenter path of the gel solver uses)."""
    rd = C.CDLL(str(REF_DIST))
    canon.fem_pt_accd.restype = C.c_int
    qs = _vf_queries()
    assert len(qs) == 250
    n_hit = n_truth = 0

    def both(p, t0, t1, t2, dp, d0, d1, d2, horizon):
        arrs = [np.ascontiguousarray(a, np.float64) for a in (p, t0, t1, t2, dp, d0, d1, d2)]
        ta, tb = C.c_double(horizon), C.c_double(horizon)
        ha = canon.fem_pt_accd(*[_d(a) for a in arrs], C.c_double(0.1), C.c_double(0.0), 1000, C.byref(ta))
        hb = rd.ref_pt_ccd(*[_d(a) for a in arrs], C.c_double(0.1), C.c_double(0.0), 1000, C.byref(tb))
        return ha, ta.value, hb, tb.value

    for q, truth in qs:
        d0 = np.linalg.norm(np.cross(q[2] - q[1], q[3] - q[1]))
        if d0 == 0.0:
            continue  # degenerate face: the closest-feature classification divides by the squared normal in both versions
        ha, ta, hb, tb = both(q[0], q[1], q[2], q[3], q[4] - q[0], q[5] - q[1], q[6] - q[2], q[7] - q[3], 1.0)
        assert ha == hb and (ta == tb or abs(ta - tb) <= 1e-12 * max(abs(tb), 1e-300) or (np.isnan(ta) and np.isnan(tb))), (q, ta, tb)
        n_hit += ha
        n_truth += truth
        if truth:
            assert ha == 1, q  # conservative: a true collision is never reported as a miss
    assert n_truth > 20 and n_hit >= n_truth
    rng = np.random.default_rng(3)
    z = np.zeros(3)
    hits = 0
    for k in range(500):
        t = rng.standard_normal((3, 3)) * 2e-3
        p = t.mean(0) + rng.standard_normal(3) * 2e-3
        dp = (t.mean(0) - p) * rng.uniform(0.2, 2.5) + rng.standard_normal(3) * 5e-4
        ha, ta, hb, tb = both(p, t[0], t[1], t[2], dp, z, z, z, 1.1)
        assert ha == hb and abs(ta - tb) <= 1e-12 * max(abs(tb), 1e-300), (k, ta, tb)
        hits += ha
    assert 100 < hits < 500


@pytest.mark.skipif(not REF_DIST.exists(), reason="oracle/_ref/libuipc_dist.so not built")
def test_edge_edge_closest_points_mollifier_match_reference_source(canon):
    """fem_ee_closest vs the reference's edge_edge_distance_flag + flagged distance / 12-gradient (all nine EE / PE / PP cases), and
    fem_ee_mollifier vs edge_edge_mollifier / its gradient with the reference's threshold."""
    rd = C.CDLL(str(REF_DIST))
    canon.fem_ee_closest.restype = C.c_int
    canon.fem_ee_mollifier.restype = C.c_double
    IP = C.POINTER(C.c_int)
    rng = np.random.default_rng(5)
    seen = set()
    for k in range(4000):
        sc = 1e-3 if k % 2 else 1.0
        a0, a1 = rng.standard_normal(3) * sc, rng.standard_normal(3) * sc
        if k % 4 == 0:  # nearly parallel pairs
            b0 = a0 + rng.standard_normal(3) * sc * 0.3
            b1 = b0 + (a1 - a0) * rng.uniform(0.3, 1.5) + rng.standard_normal(3) * sc * (1e-4 if k % 8 else 0.05)
        else:
            b0, b1 = rng.standard_normal(3) * sc, rng.standard_normal(3) * sc
        a0, a1, b0, b1 = (np.ascontiguousarray(q) for q in (a0, a1, b0, b1))
        flag = np.zeros(4, np.int32)
        D2, G2, H2 = C.c_double(), np.zeros(12), np.zeros(144)
        rd.ref_ee(_d(a0), _d(a1), _d(b0), _d(b1), flag.ctypes.data_as(IP), C.byref(D2), _d(G2), _d(H2))
        D1, r, s_, t_ = C.c_double(), np.zeros(3), C.c_double(), C.c_double()
        F = canon.fem_ee_closest(_d(a0), _d(a1), _d(b0), _d(b1), C.byref(D1), _d(r), C.byref(s_), C.byref(t_))
        assert F == 8 * flag[0] + 4 * flag[1] + 2 * flag[2] + flag[3], (k, F, flag)
        seen.add(F)
        scl = max(D2.value, 1e-300)
        assert abs(D1.value - D2.value) <= 1e-9 * scl, (k, F, D1.value, D2.value)
        s, t = s_.value, t_.value
        g12 = np.concatenate([2 * (1 - s) * r, 2 * s * r, -2 * (1 - t) * r, -2 * t * r])
        if F == 15 and np.linalg.norm(np.cross(a1 - a0, b1 - b0)) ** 2 < 1e-6 * np.dot(a1 - a0, a1 - a0) * np.dot(b1 - b0, b1 - b0):
            continue  # interior case of nearly parallel lines: the line-line formula is ill-conditioned on both sides
        assert np.abs(g12 - G2).max() <= 1e-7 * max(np.abs(G2).max(), np.sqrt(scl) * 1e-3), (k, F, g12, G2)
        # mollifier (threshold from "rest" edges = the same edges here)
        eps, e2, gm2 = C.c_double(), C.c_double(), np.zeros(12)
        rd.ref_ee_mollifier(_d(a0), _d(a1), _d(b0), _d(b1), _d(a0), _d(a1), _d(b0), _d(b1), C.byref(eps), C.byref(e2), _d(gm2))
        de, gc = C.c_double(), np.zeros(12)
        e1 = canon.fem_ee_mollifier(_d(a0), _d(a1), _d(b0), _d(b1), C.c_double(eps.value), C.byref(de), _d(gc))
        assert abs(e1 - e2.value) <= 1e-12 and np.abs(de.value * gc - gm2).max() <= 1e-9 * max(np.abs(gm2).max(), 1e-300)
    assert len(seen) == 9, seen


MUDA_EE = Path("/root/reference/source/tacex_uipc/libuipc/external/muda/test/data/unit-tests/edge-edge")


@pytest.mark.skipif(not (REF_DIST.exists() and MUDA_EE.exists()), reason="needs oracle/_ref/libuipc_dist.so and the reference tree")
def test_edge_edge_additive_ccd_matches_reference_source_on_mudas_fixtures(canon):
    """fem_ee_accd vs the reference's edge_edge_ccd (ccd.inl) on muda's edge-edge queries (rational coordinates + ground truth) and on
    random sweeps of a moving edge against a static one: identical hit / miss and time of impact; no true collision missed."""
    rd = C.CDLL(str(REF_DIST))
    canon.fem_ee_accd.restype = C.c_int
    qs = []
    for f in sorted(MUDA_EE.glob("*.csv")):
        rows = np.loadtxt(f, delimiter=",", dtype=np.float64)
        pts = np.stack([rows[:, 0] / rows[:, 1], rows[:, 2] / rows[:, 3], rows[:, 4] / rows[:, 5]], 1).reshape(-1, 8, 3)
        qs += list(zip(pts, rows[:, 6].reshape(-1, 8)[:, 0].astype(bool)))
    assert len(qs) == 74

    def both(a0, a1, b0, b1, d0, d1, d2, d3, horizon):
        arrs = [np.ascontiguousarray(a, np.float64) for a in (a0, a1, b0, b1, d0, d1, d2, d3)]
        ta, tb = C.c_double(horizon), C.c_double(horizon)
        ha = canon.fem_ee_accd(*[_d(a) for a in arrs], C.c_double(0.1), C.c_double(0.0), 1000, C.byref(ta))
        hb = rd.ref_ee_ccd(*[_d(a) for a in arrs], C.c_double(0.1), C.c_double(0.0), 1000, C.byref(tb))
        return ha, ta.value, hb, tb.value

    n_truth = 0
    for q, truth in qs:
        if np.linalg.norm(q[1] - q[0]) == 0.0 or np.linalg.norm(q[3] - q[2]) == 0.0:
            continue  # zero-length edge: 0 / 0 in the classification on both sides
        ha, ta, hb, tb = both(q[0], q[1], q[2], q[3], q[4] - q[0], q[5] - q[1], q[6] - q[2], q[7] - q[3], 1.0)
        assert ha == hb and (ta == tb or abs(ta - tb) <= 1e-12 * max(abs(tb), 1e-300) or (np.isnan(ta) and np.isnan(tb))), (q, ta, tb)
        n_truth += bool(truth)
        if truth:
            assert ha == 1, q
    assert n_truth > 5
    rng = np.random.default_rng(4)
    z = np.zeros(3)
    hits = 0
    for k in range(500):
        a0, a1, b0, b1 = (rng.standard_normal(3) * 2e-3 for _ in range(4))
        mid = 0.5 * (b0 + b1) - 0.5 * (a0 + a1)
        d = mid * rng.uniform(0.2, 2.5) + rng.standard_normal(3) * 5e-4
        ha, ta, hb, tb = both(a0, a1, b0, b1, d, d + rng.standard_normal(3) * 2e-4, z, z, 1.1)
        assert ha == hb and abs(ta - tb) <= 1e-12 * max(abs(tb), 1e-300), (k, ta, tb)
        hits += ha
    assert 50 < hits < 500
