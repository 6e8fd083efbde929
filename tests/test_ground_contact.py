"""Floor contact in the oracle (oracle/dsim_oracle.c: collide_floor, fwd_constraint; SURVEY.md 8 f-3, env_gen.py:14-21,97).

MuJoCo is not installable offline, so nothing here is pinned against MuJoCo itself (tests/test_mujoco_pin.py does that when it
can).  What these tests hold is what a restatement of MuJoCo's soft-contact model must satisfy whatever its details:
geometry of the contact lists, a closed-form resting penetration derived from the documented impedance / regularisation
formulas, the optimality conditions of the convex problem checked with Jacobians rebuilt independently in numpy, dissipation,
and Coulomb-like sliding friction."""
import numpy as np
import pytest
from scipy.optimize import nnls

from oracle import oracle as O

NOPEND = [1.35, 0.15, 7.5, 0.015, 0, 0]
PEND = [1.0, 0.17, 7.0, 0.01, 1.2, 0.3]
HB = 0.05


def _quat(rpy):
    q = np.zeros(4)
    O.lib().orc_rpy2quat(O._dp(O._arr(rpy)), O._dp(q))
    return q


def _rot(q):
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def _point_jacobian(m, qpos, point, body):
    """3 x nv Jacobian of a world point moving with `body` (2 core, 3 link, 4 pendulum), MuJoCo's qvel convention (linear
    velocity in world axes, angular velocity in body axes), from first principles - no oracle code."""
    R = _rot(qpos[3:7])
    o = qpos[:3]
    J = np.zeros((3, m.nv))
    J[:, :3] = np.eye(3)
    for k in range(3):
        J[:, 3 + k] = np.cross(R[:, k], point - o)
    if m.nv == 8:
        h = o + R @ np.array([0, 0, m.pos[3][2]])
        ax = R[:, 0]
        if body >= 3:
            J[:, 6] = np.cross(ax, point - h)
        if body >= 4:
            cx, sx = np.cos(qpos[7]), np.sin(qpos[7])
            ay = R @ np.array([0, cx, sx])
            J[:, 7] = np.cross(ay, point - h)
    return J


def test_ground_is_off_unless_asked_for():
    m = O.compile_model(NOPEND, pendulum=False)
    assert m.ground == 0 and m.ngeom == 14
    qpos = np.array([0, 0, 0.01, 1, 0, 0, 0.])
    a = O.forward(m, qpos, np.zeros(6), np.zeros(4), np.zeros(4))
    assert abs(a["qacc"][2] + 9.81) < 1e-12                       # falls through the floor: no contact stage
    m2 = O.compile_model(PEND, pendulum=True)
    assert m2.ngeom == 17


def test_no_contact_is_bitwise_the_free_flight_step():
    rng = np.random.default_rng(0)
    for pend, prm in ((False, NOPEND), (True, PEND)):
        a, b = O.compile_model(prm, pend), O.compile_model(prm, pend, ground=True)
        for _ in range(20):
            qpos = np.concatenate([[0, 0, 5.0], _quat(rng.normal(size=3)), rng.normal(size=a.nq - 7)])
            qvel, act, ctrl = rng.normal(size=a.nv), rng.uniform(0, 1, 4), rng.uniform(0, 1, 4)
            for x, y in zip(O.step(a, qpos, qvel, act, ctrl, 5), O.step(b, qpos, qvel, act, ctrl, 5)):
                assert np.array_equal(x, y)


def test_body_invweight0_from_first_principles():
    """mjModel.body_invweight0: mean diagonal of J M^-1 J^T at qpos0, J rebuilt here"""
    m = O.compile_model(NOPEND, pendulum=False)
    assert abs(m.invweight0[2][0] - 1 / sum(m.mass[:])) < 1e-12    # single rigid body: 1 / mass whatever the inertia
    m = O.compile_model(PEND, pendulum=True)
    qpos0 = np.array([0, 0, 0, 1, 0, 0, 0, 0, 0.])
    M = O.forward(m, qpos0, np.zeros(8), np.zeros(4), np.zeros(4))["M"]
    com = {2: np.array(m.ipos[2][:]), 3: np.array([0, 0, m.pos[3][2]]) + np.array(m.ipos[3][:]), 4: np.array([0, 0, m.pos[3][2]]) + np.array(m.ipos[4][:])}
    for b in (2, 3, 4):
        J = _point_jacobian(m, qpos0, com[b], b)
        A = J @ np.linalg.solve(M, J.T)
        assert abs(np.trace(A) / 3 - m.invweight0[b][0]) < 1e-12 * (1 + np.trace(A))


def test_contact_lists_level_drone():
    m = O.compile_model(NOPEND, pendulum=False)
    assert O.collide(m, [0, 0, HB / 3 + 1e-6, 1, 0, 0, 0]) == []
    cons = O.collide(m, [0.3, -0.2, 0.016, 1, 0, 0, 0])            # only the core box (half height 0.016667) reaches the floor
    assert len(cons) == 4
    xy = sorted((round(c["pos"][0] - 0.3, 9), round(c["pos"][1] + 0.2, 9)) for c in cons)
    assert xy == [(-0.05, -0.05), (-0.05, 0.05), (0.05, -0.05), (0.05, 0.05)]
    for c in cons:
        assert abs(c["dist"] - (0.016 - 0.016667)) < 1e-12 and abs(c["pos"][2] - c["dist"] / 2) < 1e-15 and c["body"] == 2
    # 9 mm lower the four arms and the 'front' marker (half thickness 0.0075 each) touch too: (1 + 1 + 4) x 4 lower corners
    assert len(O.collide(m, [0, 0, 0.007, 1, 0, 0, 0])) == 24
    # upside down: propeller discs (flat on the floor: 3 rim points each), then the motors, come first
    cons = O.collide(m, [0, 0, 0.0274, 0, 1, 0, 0])
    assert len(cons) == 12 and all(m.geom[c["geom"]].type == 1 for c in cons)


def test_contact_lists_against_brute_force_depth():
    """for random poses the deepest contact of every geom is the geom's analytically lowest point; geoms above the floor have none"""
    rng = np.random.default_rng(3)
    m = O.compile_model(PEND, pendulum=True)
    seen = 0
    for _ in range(300):
        qpos = np.concatenate([[0, 0, rng.uniform(0, 1.4)], _quat(rng.normal(size=3) * [1, 1, 3]), rng.normal(size=2) * 0.7])
        cons = O.collide(m, qpos)
        R = _rot(qpos[3:7])
        cx, sx, cy, sy = np.cos(qpos[7]), np.sin(qpos[7]), np.cos(qpos[8]), np.sin(qpos[8])
        Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]); Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        frames = {2: (qpos[:3], R), 3: (qpos[:3] + R @ [0, 0, m.pos[3][2]], R @ Rx), 4: (qpos[:3] + R @ [0, 0, m.pos[3][2]], R @ Rx @ Ry)}
        for gi in range(m.ngeom):
            g = m.geom[gi]
            o, Rb = frames[g.body]
            c, s = np.cos(g.yaw), np.sin(g.yaw)
            Rg = Rb @ np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
            gp = o + Rb @ np.array(g.pos[:])
            size = np.array(g.size[:])
            if g.type == 0:
                low = gp[2] - np.abs(Rg[2, :] * size).sum()
            elif g.type == 1:
                az = Rg[2, 2]
                low = gp[2] - size[1] * abs(az) - size[0] * np.sqrt(max(0.0, 1 - az * az))
            else:
                low = gp[2] - size[0]
            mine = [k["dist"] for k in cons if k["geom"] == gi]
            if low > 1e-12:
                assert mine == []
            elif low < -1e-12:
                assert mine and abs(min(mine) - low) < 1e-12 and max(mine) <= 1e-15 and len(mine) <= 4
                seen += 1
    assert seen > 200


def _impedance(r):
    x = min(abs(r) / 0.001, 1.0)
    y = 2 * x * x if x <= 0.5 else 1 - 2 * (1 - x) ** 2
    return 0.9 + 0.05 * y


def test_resting_penetration_matches_the_closed_form():
    """level drone at rest on the four lower corners of its core box.  16 pyramid rows, each with unit normal component, carry
    m g: 16 D(r) K d(r) |r| = m g with K = 1 / (dmax tc)^2, D = d / (4 tran (1 - d)) (mu = 1: diagApprox 2 tran, Rpy = 2 R)."""
    m = O.compile_model(NOPEND, pendulum=False, ground=True)
    qpos, qvel, act = np.array([0, 0, 0.03, 1, 0, 0, 0.]), np.zeros(6), np.zeros(4)
    for _ in range(400):
        qpos, qvel, act, sens = O.step(m, qpos, qvel, act, np.zeros(4))
    mass, tran = sum(m.mass[:]), m.invweight0[2][0]
    K = 1 / (0.95 * 0.02) ** 2

    def load(r):
        d = _impedance(r)
        return 16 * d / (4 * tran * (1 - d)) * K * d * r
    lo, hi = 0.0, 1e-3
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        lo, hi = (mid, hi) if load(mid) < mass * 9.81 else (lo, mid)
    assert abs((O.round_prec5(HB / 3) - qpos[2]) - lo) < 1e-12
    assert np.abs(qvel).max() < 1e-12 and np.abs(sens - [0, 0, 9.81]).max() < 1e-9


def test_solution_satisfies_the_contact_conditions():
    """For random touching states: M qacc = qfrc_smooth + qfrc_constraint, and qfrc_constraint is a NON-NEGATIVE combination of
    the pyramid edges n +- t1, n +- t2 at the listed contact points (Jacobians rebuilt in numpy): pushing only, inside the cone."""
    rng = np.random.default_rng(8)
    m = O.compile_model(PEND, pendulum=True, ground=True)
    m0 = O.compile_model(PEND, pendulum=True)
    checked = 0
    for _ in range(200):
        qpos = np.concatenate([rng.normal(size=2), [rng.uniform(0, 1.3)], _quat(rng.normal(size=3) * [1, 1, 3]), rng.normal(size=2) * 0.6])
        qvel, act, ctrl = rng.normal(size=8), rng.uniform(0, 1, 4), rng.uniform(0, 1, 4)
        cons = O.collide(m, qpos)
        f = O.forward_contact(m, qpos, qvel, act, ctrl)
        assert f["ncon"] == len(cons)
        free = O.forward(m0, qpos, qvel, act, ctrl)
        if not cons:
            assert np.array_equal(f["qacc"], free["qacc"]) and not f["qfrc_constraint"].any()
            continue
        M, qs = free["M"], free["qfrc_smooth"]
        assert np.abs(M @ f["qacc"] - qs - f["qfrc_constraint"]).max() < 1e-9 * (1 + np.abs(qs).max())
        edges = []
        for c in cons:
            J = _point_jacobian(m, qpos, c["pos"], c["body"])
            for d in ([0, 1, 1], [0, -1, 1], [-1, 0, 1], [1, 0, 1]):
                edges.append(J.T @ np.array(d, dtype=float))
        A = np.array(edges).T
        lam, res = nnls(A, f["qfrc_constraint"], maxiter=20 * A.shape[1])
        assert res < 1e-8 * (1 + np.abs(f["qfrc_constraint"]).max()), res
        checked += 1
    assert checked > 60


def test_impact_dissipates_and_penetration_stays_bounded():
    rng = np.random.default_rng(2)
    m = O.compile_model(PEND, pendulum=True, ground=True)
    for _ in range(4):
        qpos = np.concatenate([[0, 0, 1.6], _quat(rng.uniform(-1, 1, 3) * [1, 1, 3]), rng.uniform(-.5, .5, 2)])
        qvel, act = rng.normal(size=8), np.zeros(4)
        e0 = sum(O.energy(m, qpos, qvel))
        worst = 0.0
        for k in range(1500):
            qpos, qvel, act, sens = O.step(m, qpos, qvel, act, np.zeros(4))
            if k % 10 == 0:
                cons = O.collide(m, qpos)
                worst = max(worst, -min([c["dist"] for c in cons] + [0.0]))
        ke, pe = O.energy(m, qpos, qvel)
        assert ke + pe < e0 and ke < 1e-6                         # came to rest, lost energy
        assert worst < 0.08                                       # a 5.6 m/s impact at timeconst 0.02 s: centimetres, not more
        assert abs(np.linalg.norm(sens) - 9.81) < 1e-3            # the floor carries the weight


def test_friction_stops_a_sliding_drone():
    """resting level drone pushed sideways: the pyramid's edge rows remove the slip (in a convex soft-contact model faster than
    mu g would - the edge force carries a normal component, the known lift of sliding bodies), it does not slide forever, and a
    drone at rest stays at rest."""
    m = O.compile_model(NOPEND, pendulum=False, ground=True)
    qpos, qvel, act = np.array([0, 0, 0.0166, 1, 0, 0, 0.]), np.zeros(6), np.zeros(4)
    for _ in range(200):
        qpos, qvel, act, _ = O.step(m, qpos, qvel, act, np.zeros(4))
    rest = qpos.copy()
    for v0 in (0.05, 1.0):
        qp, qv, ac = rest.copy(), np.zeros(6), np.zeros(4)
        qv[0] = v0
        for _ in range(300):
            qp, qv, ac, sens = O.step(m, qp, qv, ac, np.zeros(4))
        assert np.abs(qv).max() < 1e-6 and 0 < qp[0] < v0 * v0 / (2 * 9.81) + 0.02 * v0 + 1e-3   # stopped at least as early as Coulomb friction would
        assert abs(np.linalg.norm(sens) - 9.81) < 1e-6
    qp, qv, ac = rest.copy(), np.zeros(6), np.zeros(4)
    for _ in range(100):
        qp, qv, ac, _ = O.step(m, qp, qv, ac, np.zeros(4))
    assert np.abs(qp - rest).max() < 1e-9


def test_contact_stage_against_an_autograd_restatement():
    """Third statement of the constraint stage, sharing no code with the oracle's (C, cdof Jacobians) or the kernel's (body-frame
    rows): the Lagrangian model of tests/test_lagrangian_pin.py supplies M = d2T/dqdot2 and, by forward-mode autodiff of
    `body origin + R_body(q) p_local`, the Jacobian of every contact point; body_invweight0, impedance, regulariser and
    reference acceleration are written out again from MuJoCo's formulas; the convex problem is minimised by a damped Newton
    iteration on autograd derivatives of the cost.  Only the contact LISTS are taken from the oracle (their geometry is held by
    test_contact_lists_against_brute_force_depth)."""
    import torch
    import test_lagrangian_pin as LP
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        rng = np.random.default_rng(4)
        worst, done = 0.0, 0
        while done < 12:
            pend = done % 3 != 2
            params = np.array(PEND) * rng.uniform(0.8, 1.2, size=6)
            if not pend:
                params[4:] = 0
            nq, nv = (9, 8) if pend else (7, 6)
            qpos = np.concatenate([rng.normal(size=2), [rng.uniform(0, 1.3 if pend else 0.2)], _quat(rng.normal(size=3) * [1, 1, 3]), rng.normal(size=nq - 7) * 0.6])
            qvel, act, ctrl = rng.normal(size=nv), rng.uniform(0, 1, 4), rng.uniform(0, 1, 4)
            m = O.compile_model(params, pend, 100.0, True, ground=True)
            cons = O.collide(m, qpos)
            if not cons:
                continue
            got = O.forward_contact(m, qpos, qvel, act, ctrl)
            free = O.forward(O.compile_model(params, pend, 100.0, True), qpos, qvel, act, ctrl)
            mod = LP.Model(m, pend)
            R0 = torch.tensor(LP._quat_to_mat(qpos[3:7]))
            q = torch.zeros(nv); q[0:3] = torch.tensor(qpos[0:3])
            if pend:
                q[6:8] = torch.tensor(qpos[7:9])
            qd = torch.tensor(qvel)
            M, _ = mod.mass_and_bias(q, qd, R0)

            def body_frame(qq, b, R_=R0):                    # (origin, rotation) of MuJoCo body b = 2 core, 3 link, 4 pendulum
                p, th = qq[0:3], qq[3:6]
                R = R_ @ torch.linalg.matrix_exp(LP._hat(th))
                if b == 2:
                    return p, R
                o = p + R @ mod.pos_c
                Rc = R @ LP._rx(qq[6])
                if b == 3:
                    return o, Rc
                return o + Rc @ mod.pos_d, Rc @ LP._ry(qq[7])

            def point_jac(qq, b, world, R_=R0):
                o, Rb = body_frame(qq, b, R_)
                local = (Rb.T @ (torch.tensor(world) - o)).detach()
                return torch.func.jacfwd(lambda z: body_frame(z, b, R_)[0] + body_frame(z, b, R_)[1] @ local)(qq)

            # body_invweight0 at qpos0 (identity attitude, hinges 0)
            q0, I3 = torch.zeros(nv), torch.eye(3)
            M0, _ = mod.mass_and_bias(q0, torch.zeros(nv), I3)
            tran = {}
            for k, b in enumerate([2, 3, 4] if pend else [2]):
                com0 = mod.frames(q0, I3)[0][k][1].detach().numpy()
                J0 = point_jac(q0, b, com0, I3)
                tran[b] = torch.trace(J0 @ torch.linalg.solve(M0, J0.T)) / 3
            tc, dmax, mu = max(0.02, 2 * m.timestep), 0.95, 1.0
            K, B = 1 / (dmax * tc) ** 2, 2 / (dmax * tc)
            rows, aref, D = [], [], []
            for c in cons:
                J = point_jac(q, c["body"], c["pos"])
                d = _impedance(c["dist"])
                Rr = 2 * mu * mu * (1 - d) / d * (1 + mu * mu) * tran[c["body"]]
                for e in ([0, 1, 1], [0, -1, 1], [-1, 0, 1], [1, 0, 1]):
                    r = torch.tensor(e, dtype=torch.float64) @ J
                    rows.append(r); aref.append(-B * (r @ qd) - K * d * c["dist"]); D.append(1 / Rr)
            Jm, ar, Dv = torch.stack(rows), torch.stack(aref), torch.stack(D)
            a0 = torch.tensor(free["qacc"])

            def cost(x):
                res = torch.clamp(Jm @ x - ar, max=0.0)
                return 0.5 * (x - a0) @ M @ (x - a0) + 0.5 * (Dv * res * res).sum()
            x = a0.clone()
            for _ in range(60):                              # damped Newton on the autograd gradient / generalised Hessian
                g = torch.func.grad(cost)(x)
                if g.abs().max() < 1e-11 * (1 + (M @ a0).abs().max()):
                    break
                act_rows = (Jm @ x - ar) < 0
                H = M + (Jm[act_rows].T * Dv[act_rows]) @ Jm[act_rows]
                dx = -torch.linalg.solve(H, g)
                t, c0 = 1.0, cost(x)
                while cost(x + t * dx) > c0 + 1e-4 * t * (g @ dx) and t > 1e-10:
                    t *= 0.5
                x = x + t * dx
            err = (np.abs(x.numpy() - got["qacc"]) / (1 + np.abs(got["qacc"]))).max()
            worst = max(worst, err)
            done += 1
        print("contact stage, oracle vs autograd restatement: max relative deviation of qacc", worst)
        assert worst < 1e-8, worst
    finally:
        torch.set_default_dtype(old)
