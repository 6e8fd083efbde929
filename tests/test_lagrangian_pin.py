"""THIRD, code-independent derivation of the rigid-body dynamics of the drone + 2-hinge load, used to pin
the oracle (`orc_forward`: MuJoCo-style composite-rigid-body + recursive Newton-Euler in spatial vectors) and, through it,
the CUDA kernel (body-frame Newton-Euler + Schur complement).  MuJoCo itself is not installable offline
(tests/test_mujoco_pin.py takes over the moment it is); until then the physics rests on THREE derivations that share no
code and no intermediate quantity:

  here: Lagrangian mechanics by automatic differentiation (torch FP64).  Only the KINEMATICS are written down - where each
  body's centre of mass is and how each body is oriented as a function of the generalised coordinates (env_gen.py:45,66-72
  nesting A -> B -> C(hinge x) -> D(hinge y)) - plus T = sum 1/2 m |v_com|^2 + 1/2 w^T I w and V = sum m g z.  The mass matrix
  M = d2T/dqdot2, the bias c = (d/dq dT/dqdot) qdot - dT/dq + dV/dq (Euler-Lagrange) and the generalised forces
  Q = sum J_v^T f + J_w^T tau (virtual work, Jacobians by autodiff) come out of autograd, not out of a hand derivation.

Coordinates: q = (p, theta, phi_x, phi_y) with R = R0 exp([theta]x) around the state's own orientation R0; at theta = 0,
thetadot is the BODY-frame angular velocity and thetaddot its derivative (J_r(0) = I, d/dt J_r thetadot = -1/2 thetadot x
thetadot = 0), i.e. exactly MuJoCo's free-joint qvel / qacc convention (BaseDroneEnv.py:367-370).

Tolerances: M, bias, qfrc_smooth, qacc <= 1e-9 (1 + |x|) on 1000 random states and parameter sets.
"""
import os

import numpy as np
import pytest
import torch

# default sizes keep the CPU suite short; DSIM_PIN_CASES=1000 is the full run whose result is committed under
# profiles/r02_lagrangian_pin.txt (tools/lagrangian_pin_report.py)
NCASE = int(os.environ.get("DSIM_PIN_CASES", "120"))
NOMINAL = np.array([1, 0.17, 7, 0.01, 1.2, 0.3])
G = 9.81
RHO, ETA = 1.2, 2e-5


@pytest.fixture(autouse=True)
def _fp64_default():
    """every tensor of this module is FP64; the process-wide default is restored for the other test modules"""
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    yield
    torch.set_default_dtype(old)


def _hat(v):
    z = torch.zeros((), dtype=v.dtype)
    return torch.stack([torch.stack([z, -v[2], v[1]]), torch.stack([v[2], z, -v[0]]), torch.stack([-v[1], v[0], z])])


def _quat_to_mat(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def _rx(a):
    c, s, o, z = torch.cos(a), torch.sin(a), torch.ones(()), torch.zeros(())
    return torch.stack([torch.stack([o, z, z]), torch.stack([z, c, -s]), torch.stack([z, s, c])])


def _ry(a):
    c, s, o, z = torch.cos(a), torch.sin(a), torch.ones(()), torch.zeros(())
    return torch.stack([torch.stack([c, z, s]), torch.stack([z, o, z]), torch.stack([-s, z, c])])


class Model:
    """the numbers of a compiled OrcModel that describe mass distribution and geometry (NOT its dynamics code)"""

    def __init__(self, m, pend):
        self.pend = pend
        ids = [2, 3, 4] if pend else [2]                   # B, C, D (A = body 1 is massless)
        self.mass = [m.mass[b] for b in ids]
        self.ipos = [torch.tensor(list(m.ipos[b])) for b in ids]
        self.Ri = [torch.tensor(_quat_to_mat(list(m.iquat[b]))) for b in ids]
        self.inertia = [torch.tensor(list(m.inertia[b])) for b in ids]
        self.pos_c = torch.tensor(list(m.pos[3])) if pend else None
        self.pos_d = torch.tensor(list(m.pos[4])) if pend else None
        self.sites = [torch.tensor(list(m.site_pos[k])) for k in range(4)]
        self.gear = [list(m.gear[k]) for k in range(4)]
        self.nv = 8 if pend else 6

    def frames(self, q, R0):
        """[(R_b, com_b)] for the massive bodies and (R, p) of the free-joint frame, as functions of q"""
        p, th = q[0:3], q[3:6]
        R = R0 @ torch.linalg.matrix_exp(_hat(th))
        out = [(R, p + R @ self.ipos[0])]
        if self.pend:
            o = p + R @ self.pos_c
            Rc = R @ _rx(q[6])
            out.append((Rc, o + Rc @ self.ipos[1]))
            od = o + Rc @ self.pos_d
            Rd = Rc @ _ry(q[7])
            out.append((Rd, od + Rd @ self.ipos[2]))
        return out, (R, p)

    def body_velocities(self, q, qd, R0):
        """per body: (R_b, com velocity (world), angular velocity (world)) via forward-mode derivatives of the frames"""
        def flat(qq):
            fr, _ = self.frames(qq, R0)
            return tuple(x for Rb, cb in fr for x in (Rb, cb))
        vals, dots = torch.func.jvp(flat, (q,), (qd,))
        out = []
        for k in range(len(vals) // 2):
            Rb, Rbd, cbd = vals[2 * k], dots[2 * k], dots[2 * k + 1]
            W = Rbd @ Rb.T                                  # [w]x
            w = torch.stack([W[2, 1], W[0, 2], W[1, 0]])
            out.append((Rb, cbd, w))
        return out

    def kinetic(self, q, qd, R0):
        T = torch.zeros(())
        for k, (Rb, v, w) in enumerate(self.body_velocities(q, qd, R0)):
            wl = (Rb @ self.Ri[k]).T @ w                    # angular velocity in the body's principal frame
            T = T + 0.5 * self.mass[k] * (v @ v) + 0.5 * (self.inertia[k] * wl * wl).sum()
        return T

    def potential(self, q, R0):
        fr, _ = self.frames(q, R0)
        return sum(self.mass[k] * G * c[2] for k, (_, c) in enumerate(fr))

    def mass_and_bias(self, q, qd, R0):
        Tf = lambda a, b: self.kinetic(a, b, R0)
        M = torch.func.hessian(Tf, argnums=1)(q, qd)
        dTdq = torch.func.grad(Tf, argnums=0)(q, qd)
        mixed = torch.func.jacfwd(torch.func.grad(Tf, argnums=1), argnums=0)(q, qd)      # d/dq (dT/dqdot)
        dVdq = torch.func.grad(lambda a: self.potential(a, R0))(q)
        return M, mixed @ qd - dTdq + dVdq

    def generalized_force(self, q, R0, wrenches):
        """virtual work: wrenches = [(point_fn(q) -> world point, force_world, torque_world, R_fn(q))]"""
        Q = torch.zeros(self.nv)
        for point_fn, R_fn, f, tau in wrenches:
            Jv = torch.func.jacfwd(point_fn)(q)                                          # 3 x nv
            # angular Jacobian: column j = vee(dR/dq_j R^T)
            dR = torch.func.jacfwd(R_fn)(q)                                              # 3 x 3 x nv
            Rv = R_fn(q)
            Jw = torch.stack([torch.stack([(dR[:, :, j] @ Rv.T)[2, 1], (dR[:, :, j] @ Rv.T)[0, 2], (dR[:, :, j] @ Rv.T)[1, 0]]) for j in range(self.nv)], dim=1)
            Q = Q + Jv.T @ f + Jw.T @ tau
        return Q

    def applied_wrenches(self, q, qd, R0, act):
        """actuators (env_gen.py:59-64: site transmission, gear (0,0,F, 0,0,+-F/100), force = act) and the inertia-box fluid
        model on every massive body (viscous + quadratic drag from the equivalent box, density 1.2, viscosity 2e-5)"""
        wr = []
        fr, (R, p) = self.frames(q, R0)
        Rfree = lambda qq: self.frames(qq, R0)[1][0]
        for k in range(4):
            g = self.gear[k]
            f = R.detach() @ torch.tensor(g[0:3]) * act[k]
            tau = R.detach() @ torch.tensor(g[3:6]) * act[k]
            wr.append(((lambda qq, k=k: self.frames(qq, R0)[1][1] + self.frames(qq, R0)[1][0] @ self.sites[k]), Rfree, f, tau))
        vel = self.body_velocities(q, qd, R0)
        for k, (Rb, v, w) in enumerate(vel):
            Rp = (Rb @ self.Ri[k]).detach()                 # principal frame -> world
            I, mb = self.inertia[k], self.mass[k]
            box = torch.sqrt(torch.clamp(torch.stack([I[1] + I[2] - I[0], I[0] + I[2] - I[1], I[0] + I[1] - I[2]]), min=1e-15) / mb * 6.0)
            vl, wl = Rp.T @ v.detach(), Rp.T @ w.detach()
            d = box.sum() / 3.0
            fl = -3.0 * np.pi * d * ETA * vl
            tl = -np.pi * d ** 3 * ETA * wl
            bx, by, bz = box
            fl = fl - 0.5 * RHO * torch.stack([by * bz, bx * bz, bx * by]) * vl.abs() * vl
            tl = tl - RHO / 64.0 * torch.stack([bx * (by ** 4 + bz ** 4), by * (bx ** 4 + bz ** 4), bz * (bx ** 4 + by ** 4)]) * wl.abs() * wl
            wr.append(((lambda qq, k=k: self.frames(qq, R0)[0][k][1]), (lambda qq, k=k: self.frames(qq, R0)[0][k][0]), Rp @ fl, Rp @ tl))
        return wr


def _rand_case(rng, pend):
    quat = rng.normal(size=4)
    quat /= np.linalg.norm(quat)
    qpos = np.concatenate([[0.3, -0.2, 15.0] + rng.normal(size=3), quat] + ([rng.normal(size=2) * 0.8] if pend else []))
    qvel = rng.normal(size=8 if pend else 6) * np.array([2, 2, 2, 3, 3, 3, 4, 4][:8 if pend else 6])
    act = rng.uniform(0, 1, size=4)
    params = NOMINAL * rng.uniform(0.7, 1.3, size=6)
    if not pend:
        params[4:] = 0
    return qpos, qvel, act, params


@pytest.mark.parametrize("pend,ncase", [(True, NCASE), (False, max(20, NCASE // 5))])
def test_oracle_forward_matches_lagrangian_autograd(oracle, pend, ncase):
    rng = np.random.default_rng(2026)
    worst = dict(M=0.0, bias=0.0, smooth=0.0, qacc=0.0)
    for case in range(ncase):
        qpos, qvel, act, params = _rand_case(rng, pend)
        m = oracle.compile_model(params, pend, 100.0, bool(case % 2))
        f = oracle.forward(m, qpos, qvel, act, act)
        mod = Model(m, pend)
        nv = mod.nv
        R0 = torch.tensor(_quat_to_mat(qpos[3:7]))
        q = torch.zeros(nv)
        q[0:3] = torch.tensor(qpos[0:3])
        if pend:
            q[6:8] = torch.tensor(qpos[7:9])
        qd = torch.tensor(qvel)
        M, bias = mod.mass_and_bias(q, qd, R0)
        Q = mod.generalized_force(q, R0, mod.applied_wrenches(q, qd, R0, torch.tensor(act)))
        if pend:                                           # hinge damping 0.15 (env_gen.py:23; the <freejoint> takes no defaults)
            Q = Q + torch.cat([torch.zeros(6), -0.15 * qd[6:8]])
        smooth = Q - bias
        qacc = torch.linalg.solve(M, smooth)
        for key, got, ref in (("M", f["M"], M.numpy()), ("smooth", f["qfrc_smooth"], smooth.numpy()), ("qacc", f["qacc"], qacc.numpy())):
            err = np.abs(got - ref) / (1 + np.abs(ref))
            worst[key] = max(worst[key], err.max())
        # bias alone: the same model with no fluid, no damping, no actuation -> qfrc_smooth = -bias
        if case % 10 == 0:
            import copy
            m0 = copy.copy(m)
            m0.density, m0.viscosity = 0.0, 0.0
            for k in range(8):
                m0.damping[k] = 0.0
            f0 = oracle.forward(m0, qpos, qvel, np.zeros(4), np.zeros(4))
            worst["bias"] = max(worst["bias"], (np.abs(-f0["qfrc_smooth"] - bias.numpy()) / (1 + np.abs(bias.numpy()))).max())
    assert worst["M"] <= 1e-9 and worst["bias"] <= 1e-9 and worst["smooth"] <= 1e-9 and worst["qacc"] <= 1e-9, worst
    if os.environ.get("DSIM_PIN_REPORT"):
        with open(os.environ["DSIM_PIN_REPORT"], "a") as fh:
            fh.write(f"pendulum={pend} cases={ncase} max relative deviation oracle vs Lagrangian autograd: " + ", ".join(f"{k} {v:.2e}" for k, v in worst.items()) + "\n")


def test_euler_step_with_implicit_hinge_damping_from_the_lagrangian(oracle):
    """mj_EulerSkip: qvel += h (M + h diag(B))^-1 qfrc_smooth, then positions with the NEW velocity; the
    quaternion advances by the body-frame rotation h w.  Rebuilt here from the autograd M / qfrc_smooth and compared with
    orc_step on 100 states (<= 1e-9)."""
    rng = np.random.default_rng(7)
    h = 0.01
    for case in range(max(20, NCASE // 4)):
        qpos, qvel, act, params = _rand_case(rng, True)
        ctrl = rng.uniform(0, 1, size=4)
        m = oracle.compile_model(params, True, 100.0, True)
        h = m.timestep
        mod = Model(m, True)
        R0 = torch.tensor(_quat_to_mat(qpos[3:7]))
        q = torch.zeros(8)
        q[0:3], q[6:8] = torch.tensor(qpos[0:3]), torch.tensor(qpos[7:9])
        qd = torch.tensor(qvel)
        M, bias = mod.mass_and_bias(q, qd, R0)
        Q = mod.generalized_force(q, R0, mod.applied_wrenches(q, qd, R0, torch.tensor(act))) + torch.cat([torch.zeros(6), -0.15 * qd[6:8]])
        Md = M + h * torch.diag(torch.tensor([0, 0, 0, 0, 0, 0, 0.15, 0.15]))
        v1 = (qd + h * torch.linalg.solve(Md, Q - bias)).numpy()
        p1 = qpos[0:3] + h * v1[0:3]
        w = v1[3:6]
        ang = h * np.linalg.norm(w)
        dq = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * w / np.linalg.norm(w)])
        a, b = qpos[3:7], dq
        q1 = np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                       a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])
        hinge1 = qpos[7:9] + h * v1[6:8]
        act1 = act + h * (np.clip(ctrl, 0, 1) - act) / m.tau[0]
        oqp, oqv, oact, _ = oracle.step(m, qpos, qvel, act, ctrl, 1)
        assert np.abs(oqv - v1).max() <= 1e-9 * (1 + np.abs(v1).max())
        assert np.abs(oqp[0:3] - p1).max() <= 1e-10 and np.abs(oqp[7:9] - hinge1).max() <= 1e-10
        assert min(np.abs(oqp[3:7] - q1).max(), np.abs(oqp[3:7] + q1).max()) <= 1e-10
        assert np.abs(oact - act1).max() <= 1e-12


def test_accelerometer_from_the_lagrangian(oracle):
    """mj_sensorAcc: site acceleration minus gravity in the site frame, from the explicit qacc and the
    pre-integration state.  The site's world acceleration is obtained here by differentiating its world POSITION twice
    along the autograd trajectory (second-order forward mode), not through an acceleration formula."""
    rng = np.random.default_rng(11)
    for case in range(max(20, NCASE // 4)):
        qpos, qvel, act, params = _rand_case(rng, True)
        m = oracle.compile_model(params, True, 100.0, True)
        f = oracle.forward(m, qpos, qvel, act, act)
        mod = Model(m, True)
        R0 = torch.tensor(_quat_to_mat(qpos[3:7]))
        q = torch.zeros(8)
        q[0:3], q[6:8] = torch.tensor(qpos[0:3]), torch.tensor(qpos[7:9])
        qd, qdd = torch.tensor(qvel), torch.tensor(f["qacc"])
        sense = torch.tensor(list(m.site_pos[4]))
        site = lambda qq: mod.frames(qq, R0)[1][1] + mod.frames(qq, R0)[1][0] @ sense
        # x(t) = site(q + qd t + 1/2 qdd t^2): d2x/dt2 at t = 0
        path = lambda t: site(q + qd * t + 0.5 * qdd * t * t)
        d1 = lambda t: torch.func.jvp(path, (t,), (torch.ones(()),))[1]
        acc_world = torch.func.jvp(d1, (torch.zeros(()),), (torch.ones(()),))[1]
        reading = (R0.T @ (acc_world + torch.tensor([0, 0, G]))).numpy()
        assert np.abs(reading - f["sensordata"]).max() <= 1e-9 * (1 + np.abs(reading).max()), (case, reading, f["sensordata"])
