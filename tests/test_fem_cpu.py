"""CPU suite for the gel FEM restatement (oracle/fem_canon.c). libuipc cannot run here (solver loop UNPINNED; the per-element physics is pinned against the reference source in
test_fem_ref_pin_cpu.py), so the
restatement is pinned by known-answer tests: finite differences, numpy.linalg.eigh / solve (protocol P5)."""
import ctypes as C

import numpy as np
import pytest

from oracle import fem_canon as fc
from tacex_b200 import gel_mesh

MU, LAM = gel_mesh.lame(1e4, 0.49)[1], gel_mesh.lame(1e4, 0.49)[0]


def _snh(F):
    F = np.ascontiguousarray(F, np.float64)
    E = C.c_double()
    g = np.zeros(9)
    H = np.zeros((9, 9))
    fc.lib().fem_snh(fc._d(F), C.c_double(MU), C.c_double(LAM), C.byref(E), fc._d(g), fc._d(H))
    return E.value, g, H


def test_snh_energy_gradient_hessian_finite_differences():
    rng = np.random.default_rng(0)
    for _ in range(5):
        F = (np.eye(3) + 0.2 * rng.standard_normal((3, 3))).T.reshape(-1)  # column-major vec
        E, g, H = _snh(F)
        J = np.linalg.det(F.reshape(3, 3).T)
        ref = 0.5 * LAM * (J - 1) ** 2 - MU * (J - 1) + 0.5 * MU * ((F ** 2).sum() - 3) + MU ** 2 / LAM ** 2
        assert abs(E - ref) <= 1e-9 * abs(ref)
        h = 1e-6
        gfd = np.zeros(9)
        Hfd = np.zeros((9, 9))
        for i in range(9):
            d = np.zeros(9)
            d[i] = h
            Ep, gp, _ = _snh(F + d)
            Em, gm, _ = _snh(F - d)
            gfd[i] = (Ep - Em) / (2 * h)
            Hfd[:, i] = (gp - gm) / (2 * h)
        assert np.abs(g - gfd).max() <= 1e-5 * np.abs(g).max()
        assert np.abs(H - Hfd).max() <= 1e-5 * np.abs(H).max()
        assert np.abs(H - H.T).max() <= 1e-12 * np.abs(H).max()


def test_snh_rest_state_is_stress_free_minimum():
    E, g, H = _snh(np.eye(3).reshape(-1))
    assert np.abs(g).max() < 1e-9 and abs(E - MU ** 2 / LAM ** 2) < 1e-9


@pytest.mark.parametrize("n", [3, 9, 12])
def test_make_spd_matches_eigh(n):
    rng = np.random.default_rng(n)
    for trial in range(4):
        A = rng.standard_normal((n, n))
        A = A + A.T
        if trial == 3:
            A = A @ A.T + np.eye(n)  # already PD: must be returned unchanged (LDL^T fast path)
        w, V = np.linalg.eigh(A)
        ref = (V * np.maximum(w, 0)) @ V.T
        H = A.copy()
        fc.lib().fem_spd_project(n, fc._d(H))
        assert np.abs(H - ref).max() <= 1e-10 * np.abs(A).max()
        assert np.linalg.eigvalsh(H).min() >= -1e-10 * np.abs(A).max()
        if trial == 3:
            assert np.array_equal(H, A)


def test_barrier_derivatives_and_limits():
    dhat, kappa = 5e-4, 1e10 * 1e-4
    f = lambda D: [v.value for v in _bar(D, dhat, kappa)]  # noqa: E731

    def _bar(D, dh, k):
        B, dB, ddB = C.c_double(), C.c_double(), C.c_double()
        fc.lib().fem_barrier(C.c_double(D), C.c_double(dh), C.c_double(k), C.byref(B), C.byref(dB), C.byref(ddB))
        return B, dB, ddB

    D0 = dhat * dhat
    assert f(D0) == [0, 0, 0] and f(2 * D0) == [0, 0, 0]  # zero at and beyond d_hat (Appendix E)
    for D in (0.9 * D0, 0.5 * D0, 0.05 * D0):
        B, dB, ddB = f(D)
        h = D * 1e-5
        assert B > 0 and dB < 0 and ddB > 0
        assert abs((f(D + h)[0] - f(D - h)[0]) / (2 * h) - dB) <= 1e-6 * abs(dB)
        assert abs((f(D + h)[1] - f(D - h)[1]) / (2 * h) - ddB) <= 1e-6 * abs(ddB)
    assert f(1e-12 * D0)[0] > f(1e-6 * D0)[0] > 10 * f(0.5 * D0)[0]  # grows (logarithmically) without bound towards contact


@pytest.mark.parametrize("kind", [0, 1])
def test_indenter_sdf_gradient_and_hessian(kind):
    rng = np.random.default_rng(kind)
    th = 0.4
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]])
    ind = fc.make_indenter(kind, (1e-3, -2e-3, 5e-3), (3e-3, 2e-3, 1e-3), R)

    def sdf(x):
        d = C.c_double()
        n = np.zeros(3)
        H = np.zeros((3, 3))
        fc.lib().fem_indenter_sdf(C.byref(ind), fc._d(np.ascontiguousarray(x)), C.byref(d), fc._d(n), fc._d(H))
        return d.value, n, H

    for _ in range(20):
        x = np.array([1e-3, -2e-3, 5e-3]) + rng.standard_normal(3) * 6e-3
        d, n, H = sdf(x)
        if d <= 1e-4:
            continue
        assert abs(np.linalg.norm(n) - 1) < 1e-12
        h = 1e-7
        nfd = np.array([(sdf(x + h * e)[0] - sdf(x - h * e)[0]) / (2 * h) for e in np.eye(3)])
        assert np.abs(n - nfd).max() < 1e-5
        Hfd = np.stack([(sdf(x + h * e)[1] - sdf(x - h * e)[1]) / (2 * h) for e in np.eye(3)], 1)
        assert np.abs(H - Hfd).max() < 1e-3 * max(1.0, np.abs(H).max())


def _small():
    m = gel_mesh.box_gel(cells=(3, 3, 2))
    return m, fc.CanonFem(m, velocity_tol=1e-3)


def test_mass_and_volume_quirk():
    m = gel_mesh.box_gel()
    a = fc.CanonFem(m, rest_volume_det=True)
    b = fc.CanonFem(m, rest_volume_det=False)
    vol = 20.75e-3 * 25.25e-3 * 4.5e-3
    assert abs(a.mass.sum() - 1e3 * vol) < 1e-12  # lumped mass uses the TRUE volume
    assert abs(b.vol.sum() - vol) < 1e-15 and abs(a.vol.sum() - 6 * vol) < 1e-14  # elastic "volume" = det(Dm), quirk Q10
    assert (m.tets.max() == len(m.X) - 1) and len(m.tets) == 2160 and len(m.X) == 572


def test_pcg_matches_dense_solve():
    m, cf = _small()
    g = cf.cfg
    n = 3 * g.V
    rng = np.random.default_rng(1)
    x = cf.X + 2e-5 * rng.standard_normal(cf.X.shape)
    xp = cf.X.copy()
    xt = cf.X.copy()
    aim = cf.X[cf.attach].copy()
    ind = fc.make_indenter(0, (0, 0, 4.5e-3 + 3e-3 + 2e-4), (3e-3, 0, 0))
    A = np.zeros((n, n))
    b = np.zeros(n)
    E = C.c_double()
    args = (C.byref(g), fc._i(cf.tets), fc._d(cf.Dm_inv), fc._d(cf.vol), fc._d(cf.mass), fc._i(cf.attach), fc._i(cf.surf),
            fc._d(aim), C.byref(ind), fc._d(x), fc._d(xp), fc._d(xt), C.c_double(1.0))
    fc.lib().fem_assemble_dense(*args, fc._d(A), fc._d(b), C.byref(E))
    assert np.abs(A - A.T).max() <= 1e-9 * np.abs(A).max()
    assert np.linalg.eigvalsh(A).min() > 0  # SPD after projection
    sol = np.zeros(n)
    g.pcg_tol_rate = 1e-16
    it = fc.lib().fem_pcg_solve(*args, fc._d(sol))
    g.pcg_tol_rate = 1e-3
    ref = np.linalg.solve(A, b)
    assert it > 1 and np.abs(sol - ref).max() <= 1e-6 * np.abs(ref).max()


def test_gradient_of_total_energy_finite_differences():
    m, cf = _small()
    g = cf.cfg
    n = 3 * g.V
    rng = np.random.default_rng(2)
    x = cf.X + 1e-5 * rng.standard_normal(cf.X.shape)
    xp, xt = cf.X.copy(), cf.X + 1e-6
    aim = cf.X[cf.attach].copy()
    ind = fc.make_indenter(1, (0, 0, 4.5e-3 + 1e-3 + 3e-4), (4e-3, 4e-3, 1e-3))

    def eval_(xx):
        A = np.zeros((n, n)); b = np.zeros(n); E = C.c_double()
        fc.lib().fem_assemble_dense(C.byref(g), fc._i(cf.tets), fc._d(cf.Dm_inv), fc._d(cf.vol), fc._d(cf.mass),
                                    fc._i(cf.attach), fc._i(cf.surf), fc._d(aim), C.byref(ind),
                                    fc._d(np.ascontiguousarray(xx)), fc._d(xp), fc._d(xt), C.c_double(1.0), fc._d(A), fc._d(b),
                                    C.byref(E))
        return E.value, b

    E0, b = eval_(x)
    idx = rng.choice(n, 12, replace=False)
    for i in idx:
        d = np.zeros(n); d[i] = 1e-9
        Ep, _ = eval_((x.reshape(-1) + d).reshape(x.shape))
        Em, _ = eval_((x.reshape(-1) - d).reshape(x.shape))
        fd = (Ep - Em) / 2e-9
        assert abs(-b[i] - fd) <= 1e-4 * max(abs(fd), np.abs(b).max() * 1e-3)


def test_rigid_translation_of_aims_has_zero_elastic_energy_and_rest_is_fixed_point():
    m, cf = _small()
    x, v, xp = cf.new_state(1)
    cf.cfg.gravity[:] = (0, 0, 0)
    far = fc.make_indenter(0, (0, 0, 1.0), (1e-3, 0, 0))
    st = cf.step(x, v, xp, cf.X[cf.attach][None], [far], [far])
    assert np.abs(x[0] - cf.X).max() < 1e-12 and st[0]["converged"] == 1


def test_press_keeps_gap_positive_and_is_monotone():
    m = gel_mesh.box_gel()
    cf = fc.CanonFem(m, velocity_tol=1e-3)
    x, v, xp = cf.new_state(1)
    aim = cf.X[cf.attach][None]
    r = 3e-3
    z0 = 4.5e-3 + r + 4e-4
    prev = fc.make_indenter(0, (0, 0, z0), (r, 0, 0))
    tops = []
    for s in range(6):
        nxt = fc.make_indenter(0, (0, 0, z0 - 1e-3 * (s + 1) / 6), (r, 0, 0))
        st = cf.step(x, v, xp, aim, [prev], [nxt])
        prev = nxt
        assert st[0]["min_dist"] > 0 and np.isfinite(x).all()
        tops.append(x[0][:, 2].max() - 0)  # track
    centre = np.argmin(np.abs(cf.X[:, 0]) + np.abs(cf.X[:, 1]) - cf.X[:, 2])
    assert x[0][centre, 2] < cf.X[centre, 2] - 3e-4  # the gel under the sphere moved down by > 0.3 mm


def test_friction_terms_match_finite_differences():
    """Lagged IPC friction of a surface vertex against the prescribed indenter (ref: ipc_vertex_half_plane_frictional_contact.cu,
    codim_ipc_contact_function.h:16-128): gradient / Hessian vs finite differences of the energy in the static, the C1-clamped
    and the sliding regime; zero when the vertex was outside d_hat at the start of the step; SPD Hessian."""
    import ctypes as C

    from oracle import fem_canon as fc
    from tacex_b200 import gel_mesh

    cf = fc.CanonFem(gel_mesh.box_gel(cells=(2, 2, 1)), friction_mu=0.5, eps_velocity=0.01)
    lib = fc.lib()
    ind0 = fc.make_indenter(0, (0.0, 0.0, 0.0103), (3e-3, 0, 0))         # sphere 0.3 mm above the point below
    ind1 = fc.make_indenter(0, (2e-5, -1e-5, 0.0102), (3e-3, 0, 0))       # moved during the step
    xp = np.array([4e-4, -3e-4, 0.0070])
    eps = cf.cfg.eps_velocity * cf.cfg.dt

    def terms(x):
        E = C.c_double()
        G = np.zeros(3)
        H = np.zeros(9)
        lib.fem_friction_terms(C.byref(cf.cfg), C.byref(ind0), C.byref(ind1), fc._d(xp), fc._d(np.ascontiguousarray(x)), C.byref(E), fc._d(G), fc._d(H))
        return E.value, G, H.reshape(3, 3)

    for slip in (0.0, 0.3 * eps, 5 * eps):
        x = xp + np.array([2e-5, -1e-5, -1e-4]) + slip * np.array([0.6, 0.8, 0.0])
        E, G, H = terms(x)
        assert E > 0 and np.all(np.linalg.eigvalsh(0.5 * (H + H.T)) >= -1e-9 * abs(H).max())
        h = 1e-9
        Gn = np.array([(terms(x + h * e)[0] - terms(x - h * e)[0]) / (2 * h) for e in np.eye(3)])
        np.testing.assert_allclose(G, Gn, rtol=2e-5, atol=1e-7 * max(1.0, abs(G).max()))
        if slip != 0.0:  # at zero slip the energy is only C1: the analytic Hessian is the one-sided limit
            Hn = np.array([(terms(x + h * e)[1] - terms(x - h * e)[1]) / (2 * h) for e in np.eye(3)]).T
            np.testing.assert_allclose(H, Hn, rtol=1e-3, atol=1e-4 * abs(H).max())
    far = xp.copy()
    far[2] = 0.0060  # more than d_hat away from the indenter at the start of the step: no lagged normal force
    E = C.c_double(1.0)
    G = np.ones(3)
    lib.fem_friction_terms(C.byref(cf.cfg), C.byref(ind0), C.byref(ind1), fc._d(far), fc._d(far + 1e-4), C.byref(E), fc._d(G), None)
    assert E.value == 0.0 and not G.any()


def test_triangle_mesh_indenter_of_the_restatement():
    """Mesh indenter (type 2) of the CPU restatement: (a) a box given as 12 triangles presses the gel like the analytic box SDF
    (same barrier wherever a single face candidate is active; near the face diagonals two candidates are, as in the reference,
    which makes the mesh contact slightly stiffer), (b) a 90 deg wedge mesh converges every step and no gel surface vertex ends
    up inside it."""
    from oracle import fem_canon as fc
    from tacex_b200 import gel_mesh, synth

    m = gel_mesh.box_gel()
    cf = fc.CanonFem(m, velocity_tol=1e-3)
    h = (2e-3, 3e-3, 1e-3)
    v = np.array([[sx * h[0], sy * h[1], sz * h[2]] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)])
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    box = np.array([[v[a], v[b], v[c]] for a, b, c, d in quads] + [[v[a], v[c], v[d]] for a, b, c, d in quads])
    aim = cf.X[cf.attach][None]
    z0 = 4.5e-3 + h[2] + 4e-4
    ctr = lambda s: [1e-3, 0.5e-3, z0 - 1e-3 * s / 6]  # noqa: E731
    fc.CanonFem.set_indenter_mesh(box)
    xa, va, xpa = cf.new_state(1)
    xb, vb, xpb = cf.new_state(1)
    for s in range(6):
        sa = cf.step(xa, va, xpa, aim, [fc.make_indenter(2, ctr(s), (0, 0, 0))], [fc.make_indenter(2, ctr(s + 1), (0, 0, 0))])
        sb = cf.step(xb, vb, xpb, aim, [fc.make_indenter(1, ctr(s), h)], [fc.make_indenter(1, ctr(s + 1), h)])
        assert sa[0]["converged"] == 1 and sb[0]["converged"] == 1
    disp = np.abs(xb - cf.X).max()
    assert disp > 5e-4 and np.abs(xa - xb).max() < 0.02 * disp

    size = 3e-3
    fc.CanonFem.set_indenter_mesh(synth.indenter_mesh(2, size))
    x, vv, xp = cf.new_state(1)
    z0 = 4.5e-3 + 4e-4
    for s in range(8):
        st = cf.step(x, vv, xp, aim, [fc.make_indenter(2, [1e-3, 0.0, z0 - 1e-3 * s / 8], (0, 0, 0))],
                     [fc.make_indenter(2, [1e-3, 0.0, z0 - 1e-3 * (s + 1) / 8], (0, 0, 0))])
        assert st[0]["converged"] == 1 and st[0]["min_dist"] > 0
    p = x[0, cf.surf] - np.array([1e-3, 0.0, z0 - 1e-3])  # surface vertices in the wedge's frame
    inside = (np.abs(p[:, 0]) < size) & (np.abs(p[:, 1]) < 2 * size) & (p[:, 2] > np.abs(p[:, 0])) & (p[:, 2] < size)
    assert not inside.any() and np.abs(x - cf.X).max() > 1e-4


def test_step_driver_converges_to_the_stationary_point_an_independent_dense_newton_finds():
    """The Newton / PCG / line-search / CCD driver of the restatement (the part no reference source can pin) against an INDEPENDENT
    solver of the same incremental potential: plain Newton iterations with numpy's dense LU on the assembled system and energy
    backtracking. A sphere resting inside the barrier zone of the gel (with gravity, attachment and friction terms active): the
    step must end where the dense Newton iteration ends -- to within its own stopping rule (relative Newton tolerance 1e-3) --, and the
    gradient there must have dropped accordingly."""
    m, cf = _small()
    g = cf.cfg
    n = 3 * g.V
    ind = fc.make_indenter(0, (0.5e-3, -0.3e-3, 4.5e-3 + 3e-3 + 3e-4), (3e-3, 0, 0))  # 0.3 mm above the gel: inside d_hat = 0.5 mm
    aim = cf.X[cf.attach].copy()
    xp = cf.X.copy()
    xt = cf.X + np.array(list(g.gravity)) * g.dt * g.dt

    def eval_(xx, want_A=True):
        A = np.zeros((n, n)); b = np.zeros(n); E = C.c_double()
        fc.lib().fem_assemble_dense(C.byref(g), fc._i(cf.tets), fc._d(cf.Dm_inv), fc._d(cf.vol), fc._d(cf.mass), fc._i(cf.attach),
                                    fc._i(cf.surf), fc._d(aim), C.byref(ind), fc._d(np.ascontiguousarray(xx)), fc._d(xp), fc._d(xt),
                                    C.c_double(1.0), fc._d(A), fc._d(b), C.byref(E))
        return E.value, b, A

    x = cf.X.copy()
    for it in range(60):
        E0, b, A = eval_(x)
        dx = np.linalg.solve(A, b).reshape(x.shape)
        a = 1.0
        while True:
            E1 = eval_(x + a * dx)[0]
            if np.isfinite(E1) and E1 <= E0:
                break
            a *= 0.5
            assert a > 1e-12
        x = x + a * dx
        if np.abs(dx).max() < 5e-12:  # quadratic convergence reaches the rounding floor (~1e-12 m) after two or three iterations
            break
    assert it < 10
    old = (g.velocity_tol, g.pcg_tol_rate)
    g.velocity_tol, g.pcg_tol_rate = 1e-9, 1e-14
    try:
        xs, vs, xps = cf.new_state(1)
        st = cf.step(xs, vs, xps, aim[None], [ind], [ind])
    finally:
        g.velocity_tol, g.pcg_tol_rate = old
    assert st[0]["converged"] == 1
    disp = np.abs(x - cf.X).max()
    assert disp > 5e-6  # the barrier (and gravity) really moved the gel
    # the driver stops at |dx| <= 1e-3 |dx of the first iteration| (the reference's relative Newton tolerance, max_translation_checker.cu:25-50)
    assert np.abs(xs[0] - x).max() <= 1.5e-3 * disp, (np.abs(xs[0] - x).max(), disp)
    _, b_end, _ = eval_(xs[0])
    _, b0, _ = eval_(cf.X)
    assert np.abs(b_end).max() <= 5e-3 * np.abs(b0).max()


def test_indenter_vertices_against_gel_triangles_in_the_restatement():
    """Second half of the vertex-face contact (fem_set_contact_surface): a 60 deg cone whose tip comes down in the middle of a
    surface cell is invisible to the gel's vertices (2.1 mm apart) but not to its triangles. With the contact surface set the gel
    follows the tip and the tip never reaches the surface; the Gauss-Newton operator the PCG applies is the assembled dense one."""
    from oracle import fem_canon as fc
    from tacex_b200 import gel_mesh, synth

    m = gel_mesh.box_gel()
    cf = fc.CanonFem(m, velocity_tol=1e-3)
    fc.CanonFem.set_indenter_mesh(synth.indenter_mesh(3, 3e-3))
    aim = cf.X[cf.attach][None]
    z0 = 4.5e-3 + 4e-4
    ctr = lambda s: [1.0e-3, 0.5e-3, z0 - 1e-3 * s / 8]  # noqa: E731
    out = {}
    try:
        for on in (False, True):
            cf.set_contact_surface(m.top_tris if on else None)
            x, v, xp = cf.new_state(1)
            for s in range(8):
                st = cf.step(x, v, xp, aim, [fc.make_indenter(2, ctr(s), (0, 0, 0))], [fc.make_indenter(2, ctr(s + 1), (0, 0, 0))])
                assert st[0]["converged"] == 1 and st[0]["min_dist"] > 0
            out[on] = (np.abs(x - cf.X).max(), st[0]["min_dist"], x.copy())
        assert out[False][0] < 5e-5 and out[True][0] > 5e-4       # only the triangles see the tip
        assert out[True][1] < cf.cfg.d_hat                        # ... and hold it inside the barrier zone
        # operator consistency: A p of the matrix-free operator (with the rank-1 candidate terms) == dense assembled A p
        g = cf.cfg
        n = 3 * g.V
        x = out[True][2][0]
        ind = fc.make_indenter(2, ctr(8), (0, 0, 0))
        A = np.zeros((n, n)); b = np.zeros(n); E = C.c_double()
        xt = x.copy()
        fc.lib().fem_assemble_dense(C.byref(g), fc._i(cf.tets), fc._d(cf.Dm_inv), fc._d(cf.vol), fc._d(cf.mass), fc._i(cf.attach),
                                    fc._i(cf.surf), fc._d(aim[0].copy()), C.byref(ind), fc._d(np.ascontiguousarray(x)), fc._d(xt), fc._d(xt),
                                    C.c_double(1.0), fc._d(A), fc._d(b), C.byref(E))
        assert np.abs(A - A.T).max() <= 1e-9 * np.abs(A).max() and np.linalg.eigvalsh(A).min() > 0
        # gradient of the candidate energy by finite differences on a touched triangle's vertex
        top = np.asarray(m.top_tris)
        touched = int(top[np.argmin(np.linalg.norm(cf.X[top].mean(1)[:, :2] - np.array(ctr(8)[:2]), axis=1))][0])

        def energy(xx):
            A2 = np.zeros((n, n)); b2 = np.zeros(n); E2 = C.c_double()
            fc.lib().fem_assemble_dense(C.byref(g), fc._i(cf.tets), fc._d(cf.Dm_inv), fc._d(cf.vol), fc._d(cf.mass), fc._i(cf.attach),
                                        fc._i(cf.surf), fc._d(aim[0].copy()), C.byref(ind), fc._d(np.ascontiguousarray(xx)), fc._d(xt),
                                        fc._d(xt), C.c_double(1.0), fc._d(A2), fc._d(b2), C.byref(E2))
            return E2.value

        for a in range(3):
            d = np.zeros_like(x); d[touched, a] = 1e-9
            fd = (energy(x + d) - energy(x - d)) / 2e-9
            assert abs(-b[3 * touched + a] - fd) <= 2e-4 * max(abs(fd), np.abs(b).max() * 1e-3), (a, b[3 * touched + a], fd)
        # edge-edge: a small wedge turned by 90 degrees, its edge along x half-way between two rows of gel vertices
        th = np.pi / 2
        R = np.array([[np.cos(th), -np.sin(th), 0.0], [np.sin(th), np.cos(th), 0.0], [0.0, 0.0, 1.0]])
        fc.CanonFem.set_indenter_mesh(synth.indenter_mesh(2, 1.0e-3))
        ctr2 = lambda s: [1.0e-3, 1.05e-3, z0 - 0.9e-3 * s / 8]  # noqa: E731
        res = {}
        for on in (False, True):
            cf.set_contact_surface(m.top_tris if on else None)
            x, v, xp = cf.new_state(1)
            for s in range(8):
                st = cf.step(x, v, xp, aim, [fc.make_indenter(2, ctr2(s), (0, 0, 0), R)], [fc.make_indenter(2, ctr2(s + 1), (0, 0, 0), R)])
                assert st[0]["converged"] == 1 and st[0]["min_dist"] > 0
            res[on] = np.abs(x - cf.X).max()
        assert res[False] < 2e-4 and res[True] > 7e-4, res
    finally:
        cf.set_contact_surface(None)


def test_lagged_friction_covers_the_multi_vertex_candidate_families():
    """A cone tip resting in the middle of a surface cell touches the gel only through indenter-vertex / gel-triangle (and edge-edge)
    candidates. Dragged sideways, it must take the gel surface along by friction -- the lagged contact of a vertex is the resultant of
    ALL candidates it takes part in --, and clearly more than without friction."""
    from oracle import fem_canon as fc
    from tacex_b200 import gel_mesh, synth

    m = gel_mesh.box_gel()
    fc.CanonFem.set_indenter_mesh(synth.indenter_mesh(3, 3e-3))
    z0 = 4.5e-3 + 4e-4

    def ctr(s):  # 6 steps down (0.8 mm), then 5 steps sideways (0.1 mm each)
        return [1.0e-3 + 1e-4 * max(s - 6, 0), 0.5e-3, z0 - 0.8e-3 * min(s, 6) / 6]

    drag = {}
    try:
        for mu in (0.0, 0.5):
            cf = fc.CanonFem(m, velocity_tol=1e-3, friction_mu=mu)
            cf.set_contact_surface(m.top_tris)
            x, v, xp = cf.new_state(1)
            aim = cf.X[cf.attach][None]
            for s in range(11):
                st = cf.step(x, v, xp, aim, [fc.make_indenter(2, ctr(s), (0, 0, 0))], [fc.make_indenter(2, ctr(s + 1), (0, 0, 0))])
                assert st[0]["converged"] == 1 and st[0]["min_dist"] > 0
            top = np.unique(np.asarray(m.top_tris))
            near = top[np.linalg.norm(cf.X[top][:, :2] - np.array([1.0e-3, 0.5e-3]), axis=1) < 2.2e-3]
            drag[mu] = float((x[0, near, 0] - cf.X[near, 0]).mean())
    finally:
        cf.set_contact_surface(None)
    assert drag[0.5] > drag[0.0] + 3e-5, drag
