"""Isaac x UIPC attachment pipeline (ref: tacex_uipc/sim/uipc_attachments.py:247-428): which gel vertices
hang on the sensor case (init, host) and their per-step aim positions (device, every env in one launch)."""
import numpy as np
import pytest
import torch


def _quat(axis, ang):
    a = np.asarray(axis, np.float64) / np.linalg.norm(axis)
    return np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * a])


def _matrix_from_quat_f32(q):
    """isaaclab.utils.math.matrix_from_quat (the pytorch3d formula) in float32."""
    r, i, j, k = [np.float32(v) for v in q]
    two_s = np.float32(2.0) / (r * r + i * i + j * j + k * k)
    return np.array([[1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r)],
                     [two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r)],
                     [two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)]], np.float32)


def test_attachment_selection_and_offsets_on_the_structured_gel():
    from tacex_b200 import fem, gel_mesh

    m = gel_mesh.box_gel()
    X = np.asarray(m.X, np.float64)
    # the sensor case: a 32 x 28 x 24 mm box whose top face carries the pad (pad bottom z = 0), rotated and moved in the world
    q = _quat((0.2, -0.5, 1.0), 0.7)
    R = _matrix_from_quat_f32(q).astype(np.float64)
    pos = np.array([0.3, -0.1, 0.25])
    case_half = np.array([16e-3, 14e-3, 12e-3])
    Xw = (X + np.array([0, 0, 12e-3])) @ R.T + pos  # pad sits on the case's top face (case centre = origin of the case frame)
    offs, idx = fem.compute_attachment_data(Xw, pos, q, case_half)
    # exactly the bottom layer of the gel (within 0.5 mm of the case), i.e. the mesh's own attach list; the 1.5 mm layer above is free
    assert sorted(idx.tolist()) == sorted(np.asarray(m.attach).tolist())
    assert offs.dtype == np.float32 and np.abs(offs - (X[idx] + np.array([0, 0, 12e-3]))).max() < 2e-7
    # a case that is too small / far away attaches nothing
    assert fem.compute_attachment_data(Xw, pos + np.array([0, 0, 1.0]), q, case_half)[1].size == 0


@pytest.mark.gpu
def test_aim_positions_kernel_and_a_moving_case_drags_the_gel():
    from tacex_b200 import fem, gel_mesh

    m = gel_mesh.box_gel()
    eng = fem.GelFemEngine(m, fem.GelFemCfg(newton_velocity_tol=1e-3))
    X = np.asarray(m.X, np.float64)
    idx = np.asarray(m.attach)
    offs = X[idx].astype(np.float32)  # body frame = pad frame at rest
    N = 5
    rng = np.random.default_rng(1)
    poses = np.zeros((N, 7), np.float32)
    for e in range(N):
        poses[e, :3] = rng.uniform(-0.2, 0.2, 3)
        poses[e, 3:] = _quat(rng.standard_normal(3), rng.uniform(-1, 1)) * rng.uniform(0.5, 2.0)  # not normalised: two_s handles it
    aim = eng.attachment_aim(torch.from_numpy(poses).cuda(), torch.from_numpy(offs).cuda()).cpu().numpy()
    torch.cuda.synchronize()
    for e in range(N):
        ref = offs @ _matrix_from_quat_f32(poses[e, 3:]).T + poses[e, :3]
        assert aim.dtype == np.float64 and np.abs(aim[e] - ref).max() <= 4e-7 * max(1.0, np.abs(ref).max())
    # per-env offsets give the same
    aim2 = eng.attachment_aim(torch.from_numpy(poses).cuda(), torch.from_numpy(np.repeat(offs[None], N, 0).copy()).cuda()).cpu().numpy()
    assert np.array_equal(aim, aim2)
    # the case translates 0.2 mm per step in x: the attached bottom layer follows (soft constraint, strength 1000), the gel with it
    sim = fem.GelPadSim(2, m, fem.GelFemCfg(newton_velocity_tol=1e-3))
    far = fem.indenter_array(0, [[0, 0, 1.0]] * 2, (1e-3, 0, 0))
    pose = torch.zeros((2, 7), device="cuda")
    pose[:, 3] = 1.0
    o = torch.from_numpy(offs).cuda()
    for s in range(1, 6):
        pose[1, 0] = 2e-4 * s
        sim.set_attachment_aim(sim.engine.attachment_aim(pose, o))
        sim.step(far)
    torch.cuda.synchronize()
    dx = (sim.x[1, :, 0] - sim.x[0, :, 0]).cpu().numpy()
    assert abs(dx[idx].mean() - 1e-3) < 5e-5 and dx.min() > 6e-4  # bottom layer at the case, the top lags a little (inertia)
