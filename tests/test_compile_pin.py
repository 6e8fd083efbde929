"""Independent check of the model compiler (env_gen.make_drone geometry -> per-body mass, centre of mass,
inertia; what MuJoCo's compiler does with `inertiafromgeom`).  oracle/dsim_oracle.c::orc_compile and csrc/dsim_params.cuh use
the closed-form solid formulas + parallel-axis theorem.  Here the same quantities come from DIRECT INTEGRATION of the mass
density over every geom of env_gen.py:46-72 (Gauss-Legendre quadrature in the geom's own frame: exact for the polynomial
integrands of a box, Gauss x periodic trapezoid for cylinder and sphere), placed with the geom's pos / euler, summed, and
reduced to the centre of mass numerically - no inertia formula, no parallel-axis theorem.  Rounding stage off (the "%.5g" XML
stage is a property of dm_control's writer that only real MuJoCo + dm_control can pin, tests/test_mujoco_pin.py).
"""
import numpy as np

HB = 0.05                                                       # half_body_size (env_gen.py:37)


def _gauss(n, lo, hi):
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (hi - lo) * x + 0.5 * (hi + lo), 0.5 * (hi - lo) * w


def _box_points(half):
    xs = [_gauss(4, -h, h) for h in half]
    X, Y, Z = np.meshgrid(xs[0][0], xs[1][0], xs[2][0], indexing="ij")
    W = xs[0][1][:, None, None] * xs[1][1][None, :, None] * xs[2][1][None, None, :]
    return np.stack([X, Y, Z], -1).reshape(-1, 3), W.ravel()


def _cyl_points(r, half_h):
    rr, wr = _gauss(6, 0.0, r)
    zz, wz = _gauss(4, -half_h, half_h)
    th = np.arange(16) * (2 * np.pi / 16)
    R, T, Z = np.meshgrid(rr, th, zz, indexing="ij")
    W = (wr * rr)[:, None, None] * (2 * np.pi / 16) * wz[None, None, :] * np.ones_like(T)
    return np.stack([R * np.cos(T), R * np.sin(T), Z], -1).reshape(-1, 3), W.ravel()


def _sphere_points(r):
    rr, wr = _gauss(6, 0.0, r)
    cu, wu = _gauss(6, -1.0, 1.0)                               # u = cos(polar angle)
    th = np.arange(16) * (2 * np.pi / 16)
    R, U, T = np.meshgrid(rr, cu, th, indexing="ij")
    S = np.sqrt(1 - U * U)
    W = (wr * rr * rr)[:, None, None] * wu[None, :, None] * (2 * np.pi / 16) * np.ones_like(T)
    return np.stack([R * S * np.cos(T), R * S * np.sin(T), R * U], -1).reshape(-1, 3), W.ravel()


def _rz(t):
    c, s = np.cos(t), np.sin(t)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])


def _integrate(geoms):
    """geoms: [(points_local, weights, mass, R, pos)] -> (mass, com, inertia tensor about the com) by summation"""
    P, M = [], []
    for pts, w, mass, R, pos in geoms:
        P.append(pts @ R.T + pos)
        M.append(w * (mass / w.sum()))                          # uniform density: mass / volume
    P, M = np.concatenate(P), np.concatenate(M)
    m = M.sum()
    com = (M[:, None] * P).sum(0) / m
    d = P - com
    I = (M[:, None, None] * ((d * d).sum(1)[:, None, None] * np.eye(3) - d[:, :, None] * d[:, None, :])).sum(0)
    return m, com, I


def test_compiled_bodies_match_direct_integration_of_the_geoms(oracle):
    rng = np.random.default_rng(4)
    nominal = np.array([1, 0.17, 7, 0.01, 1.2, 0.3])
    for case in range(40):
        mass, arm, F, tau, L, w = nominal * rng.uniform(0.6, 1.4, size=6)
        m = oracle.compile_model([mass, arm, F, tau, L, w], True, 100.0, False)
        # ---- core body (env_gen.py:45-61): core box, 4 arms (rotated boxes), 4 motors (cylinders); front box / props have mass 0
        geoms = [(*_box_points([HB, HB, HB / 3]), 0.56 * mass, np.eye(3), np.zeros(3))]
        for i in range(4):
            th = i * np.pi / 2 - np.pi / 4
            u = np.array([np.cos(th), np.sin(th), 0])
            geoms.append((*_box_points([arm / 2, arm / 20, arm / 20]), 0.07 * mass, _rz(th), (np.sqrt(2) * HB + 0.5 * arm) * u))
            geoms.append((*_cyl_points(0.01, 0.01), 0.04 * mass, np.eye(3), (np.sqrt(2) * HB + arm) * u + np.array([0, 0, 0.015])))
            np.testing.assert_allclose(list(m.site_pos[i]), (np.sqrt(2) * HB + arm) * u, atol=1e-14)
        mB, cB, IB = _integrate(geoms)
        RB = oracle.quat2dcm(list(m.iquat[2]))
        np.testing.assert_allclose(m.mass[2], mB, rtol=1e-12)
        np.testing.assert_allclose(list(m.ipos[2]), cB, atol=1e-13)
        np.testing.assert_allclose(RB @ np.diag(list(m.inertia[2])) @ RB.T, IB, rtol=1e-10, atol=1e-13)
        # ---- link (:66-68): sphere r = 0.02, mass 0.01 at the body origin
        mC, cC, IC = _integrate([(*_sphere_points(0.02), 0.01, np.eye(3), np.zeros(3))])
        np.testing.assert_allclose(m.mass[3], mC, rtol=1e-12)
        np.testing.assert_allclose(np.diag(IC), list(m.inertia[3]), rtol=1e-10)
        assert np.abs(cC).max() < 1e-15 and list(m.pos[3]) == [0, 0, -HB / 2]
        # ---- pendulum (:69-72): pole cylinder r = 0.005, half-length L/2 at (0,0,-L/2), mass 0.2 L; weight box 0.1 cbrt(w) at (0,0,-L), mass w
        s = 0.1 * np.cbrt(w)
        mD, cD, ID = _integrate([(*_cyl_points(0.005, L / 2), 0.2 * L, np.eye(3), np.array([0, 0, -L / 2])),
                                 (*_box_points([s, s, s]), w, np.eye(3), np.array([0, 0, -L]))])
        RD = oracle.quat2dcm(list(m.iquat[4]))
        np.testing.assert_allclose(m.mass[4], mD, rtol=1e-12)
        np.testing.assert_allclose(list(m.ipos[4]), cD, atol=1e-13)
        np.testing.assert_allclose(RD @ np.diag(list(m.inertia[4])) @ RD.T, ID, rtol=1e-10, atol=1e-13)
        # ---- actuators (:62-64) and options (:82-84)
        for k in range(4):
            np.testing.assert_allclose(list(m.gear[k]), [0, 0, F, 0, 0, F / 100 * (-1) ** k], rtol=1e-15)
            assert m.tau[k] == tau
        assert m.timestep == 0.01 and m.density == 1.2 and m.viscosity == 2e-5 and list(m.gravity) == [0, 0, -9.81]
