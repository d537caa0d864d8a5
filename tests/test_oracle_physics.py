"""Pins the CPU oracle's rigid-body dynamics (the reference has no golden vectors for
gym.simulate -- SURVEY.md 8c): ABA accelerations must satisfy an independent Newton-Euler inverse
dynamics, conserve momentum in free flight, and reproduce closed-form cases."""
import numpy as np
import pytest

from isaacgymenvs_b200.assets import load_compiled
from oracle.oracle import OracleSim
from tests import rnea_np

G = (0.0, 0.0, -9.81)


def random_state(m, rng, z=2.0):
    root = np.zeros(13)
    root[:3] = [rng.normal(), rng.normal(), z]
    qt = rng.normal(size=4); root[3:7] = qt / np.linalg.norm(qt)
    root[7:13] = rng.normal(size=6)
    if m.root_fixed:
        root[7:13] = 0
    lo = np.where(m.limited[1:] > 0, m.lower[1:], -1.0); hi = np.where(m.limited[1:] > 0, m.upper[1:], 1.0)
    q = lo + (hi - lo) * rng.uniform(0.1, 0.9, size=m.ndof)
    qd = rng.normal(size=m.ndof)
    dof = np.stack([q, qd], -1)
    return root, dof


@pytest.mark.parametrize("name", ["cartpole", "ant", "humanoid", "anymal"])
def test_aba_satisfies_inverse_dynamics(name):
    m = load_compiled(name)
    dt, sub = 0.0166, 2
    sim = OracleSim(m, dt, sub, G)
    h = dt / sub
    rng = np.random.default_rng(0)
    for trial in range(5):
        root, dof = random_state(m, rng)
        tau_act = rng.normal(size=m.ndof) * 5
        qdd, ra, da = sim.forward_dynamics(root, dof, tau_act)
        root_acc = ((ra[7:10] - root[7:10]) / h, (ra[10:13] - root[10:13]) / h)
        q, qd = dof[:, 0], dof[:, 1]
        tau_req, (f0, n0) = rnea_np.inverse_dynamics(m, root, q, qd, qdd, root_acc, G)
        # forces applied under the implicit scheme: linear terms at the end of the sub-step
        qd1 = qd + h * qdd; q1 = q + h * qd1
        eff = m.effort[1:]
        applied = np.clip(tau_act, -eff, eff) - m.damping[1:] * qd1 - m.stiffness[1:] * q1 - m.armature[1:] * qdd
        scale = max(1.0, np.abs(applied).max())
        assert np.allclose(tau_req, applied, atol=1e-8 * scale, rtol=1e-8), (name, trial, tau_req - applied)
        if not m.root_fixed:   # a free root transmits no wrench
            assert np.abs(f0).max() < 1e-7 * scale and np.abs(n0).max() < 1e-7 * scale
        # integration consistency: q' = q + h*qd'
        assert np.allclose(da[:, 1], qd1, atol=1e-12) and np.allclose(da[:, 0], q1, atol=1e-12)


def test_free_fall_and_momentum():
    m = load_compiled("ant")
    dt, sub = 0.0166, 2
    rng = np.random.default_rng(1)
    # gravity off: momentum is conserved by the continuous dynamics; the generalized-coordinate
    # integrator conserves it to O(h), so the drift must be small and shrink with the step
    root_i, dof_i = random_state(m, rng, z=5.0)
    drift = []
    for scale in (1, 4):
        sim0 = OracleSim(m, dt / scale, sub, (0, 0, 0))
        root = root_i[None].copy(); dof = dof_i[None].copy()
        P0, L0 = rnea_np.momentum(m, root[0], dof[0, :, 0], dof[0, :, 1])
        for _ in range(20 * scale):
            sim0.simulate(root, dof, np.zeros((1, m.ndof)))
        P1, L1 = rnea_np.momentum(m, root[0], dof[0, :, 0], dof[0, :, 1])
        drift.append((np.abs(P1 - P0).max(), np.abs(L1 - L0).max()))
    assert drift[0][0] < 1e-2 * np.abs(P0).max() and drift[0][1] < 2e-2 * max(1.0, np.abs(L0).max())
    assert drift[1][0] < 0.5 * drift[0][0] and drift[1][1] < 0.5 * drift[0][1]
    # gravity on: COM accelerates at g => total momentum changes by M g t
    sim = OracleSim(m, dt, sub, G)
    root, dof = random_state(m, rng, z=50.0)
    root = root[None].copy(); dof = dof[None].copy()
    P0, _ = rnea_np.momentum(m, root[0], dof[0, :, 0], dof[0, :, 1])
    n = 10
    for _ in range(n):
        sim.simulate(root, dof, np.zeros((1, m.ndof)))
    P1, _ = rnea_np.momentum(m, root[0], dof[0, :, 0], dof[0, :, 1])
    assert np.allclose(P1 - P0, m.total_mass() * np.array(G) * n * dt, atol=2e-2)


def test_pendulum_period():
    """Cartpole with the cart held by a stiff drive is a physical pendulum: small-swing period
    T = 2 pi sqrt(I_pivot / (m g l)) about the hanging equilibrium."""
    m = load_compiled("cartpole")
    import copy
    m = copy.deepcopy(m)
    m.mass[1] = 1e6                     # immobile cart
    dt, sub = 0.001, 1
    sim = OracleSim(m, dt, sub, G)
    root = np.zeros((1, 13)); root[0, 6] = 1; root[0, 2] = 2
    th0 = np.pi - 0.01                  # pole hangs down at q = pi
    dof = np.array([[[0.0, 0.0], [th0, 0.0]]])
    I = m.inertia[2][0] + m.mass[2] * m.com[2][2] ** 2 + m.armature[2]
    T = 2 * np.pi * np.sqrt(I / (m.mass[2] * 9.81 * m.com[2][2]))
    qs = []
    for _ in range(int(2.2 * T / dt)):
        sim.simulate(root, dof, np.zeros((1, 2)))
        qs.append(dof[0, 1, 0] - np.pi)
    qs = np.array(qs)
    zc = np.where((qs[:-1] < 0) & (qs[1:] >= 0))[0]   # upward zero crossings, one per period
    assert len(zc) >= 2
    T_meas = (zc[1] - zc[0]) * dt
    assert abs(T_meas - T) / T < 5e-3


def test_contact_rest_and_sensor():
    """An Ant dropped from just above the ground comes to rest on it: penetration stays small, the
    net contact force balances the weight, and the settled height is near the geometric one."""
    m = load_compiled("ant")
    import copy
    m = copy.deepcopy(m)
    m.sensor_body = np.array([2, 4, 6, 8], dtype=np.int32)
    m.sensor_pos = np.zeros((4, 3)); m.sensor_quat = np.tile([0, 0, 0, 1.0], (4, 1))
    sim = OracleSim(m, 0.0166, 2, G, ground_mu=1.0)
    root = np.zeros((1, 13)); root[0, 6] = 1; root[0, 2] = 0.6
    q0 = np.where(m.lower[1:] > 0, m.lower[1:], np.where(m.upper[1:] < 0, m.upper[1:], 0.0))
    dof = np.zeros((1, 8, 2)); dof[0, :, 0] = q0
    for _ in range(240):
        out = sim.simulate(root, dof, np.zeros((1, 8)))
    W = m.total_mass() * 9.81
    Fz = out["contact_force"][0, :, 2].sum()
    assert abs(Fz - W) / W < 0.02, (Fz, W)
    assert np.abs(root[0, 7:13]).max() < 0.02 and np.abs(dof[0, :, 1]).max() < 0.05
    assert 0.05 < root[0, 2] < 0.6
    assert out["sensor"].shape == (1, 4, 6)


# ---------------------------------------------------------------------------------------------
# the free object (ShadowHand's cube): closed-form and balance checks of the two-body extension
def _hand_orc(**kw):
    from tests.hand_common import hand_setup, DT, SUBSTEPS, G as HG
    m, obj, tendons = hand_setup()
    return m, obj, OracleSim(m, DT, SUBSTEPS, HG, obj=obj, tendons=tendons, tendon_k=30.0, tendon_d=0.1, **kw), DT


def test_object_free_flight_is_ballistic_and_torque_free():
    """Away from the hand the cube is a free rigid body: parabola for the COM, constant angular momentum."""
    m, obj, orc, dt = _hand_orc()
    root = np.zeros((1, 13)); root[0, 2] = 0.5; root[0, 3:7] = m.default_root_quat
    dof = np.zeros((1, m.ndof, 2))
    o = np.zeros((1, 13)); o[0, 0:3] = [1.0, 1.0, 3.0]; o[0, 6] = 1
    o[0, 7:10] = [0.3, -0.2, 1.0]; o[0, 10:13] = [2.0, -1.0, 0.5]
    v0, p0, w0 = o[0, 7:10].copy(), o[0, 0:3].copy(), o[0, 10:13].copy()
    n = 30
    for _ in range(n):
        orc.simulate(root, dof, target=np.zeros((1, m.ndof)), obj=o)
    t = n * dt
    assert np.allclose(o[0, 7:10], v0 + np.array([0, 0, -9.81]) * t, atol=1e-9)
    # semi-implicit Euler: position error vs the parabola is g*h*t/2 with h the sub-step
    assert np.abs(o[0, 0:3] - (p0 + v0 * t + 0.5 * np.array([0, 0, -9.81]) * t * t)).max() < 9.81 * (dt / 2) * t / 2 * 1.01
    assert np.allclose(o[0, 10:13], w0, atol=1e-9)        # isotropic inertia: w itself is constant
    assert abs(np.linalg.norm(o[0, 3:7]) - 1) < 1e-12


def test_cube_rests_on_the_palm_and_forces_balance():
    """Dropped on the open hand the cube settles; the hand's net contact force equals the cube's weight (Newton III
    through the penalty contact), and the cube stays where the palm is."""
    m, obj, orc, dt = _hand_orc()
    root = np.zeros((1, 13)); root[0, 2] = 0.5; root[0, 3:7] = m.default_root_quat
    dof = np.zeros((1, m.ndof, 2))
    o = np.zeros((1, 13)); o[0, 0:3] = [0.0, -0.39, 0.56]; o[0, 6] = 1
    for _ in range(150):
        out = orc.simulate(root, dof, target=np.zeros((1, m.ndof)), obj=o)
    W = obj["mass"] * 9.81
    F = out["contact_force"][0].sum(0)                     # on the hand's bodies, world axes
    assert abs(-F[2] - W) / W < 0.03, (F, W)
    assert np.abs(o[0, 7:13]).max() < 0.02
    assert 0.5 < o[0, 2] < 0.56 and abs(o[0, 1] + 0.39) < 0.02


def test_object_rests_on_the_ground():
    m, obj, orc, dt = _hand_orc()
    root = np.zeros((1, 13)); root[0, 2] = 0.5; root[0, 3:7] = m.default_root_quat
    dof = np.zeros((1, m.ndof, 2))
    o = np.zeros((1, 13)); o[0, 0:3] = [1.0, 1.0, 0.1]; o[0, 6] = 1
    for _ in range(200):
        orc.simulate(root, dof, target=np.zeros((1, m.ndof)), obj=o)
    pen = obj["mass"] * 9.81 / 4 / (10000.0 * obj["mass"])            # four corners share the weight
    assert abs(o[0, 2] - (0.025 - pen)) < 2e-4 and np.abs(o[0, 7:13]).max() < 1e-3


def test_aba_satisfies_inverse_dynamics_position_drives():
    """The fixed-base ShadowHand articulation (24 DOF, position drives from the MJCF <position kp forcerange>, gravity off):
    the accelerations the oracle produces need exactly the PD + passive joint forces the scheme applies (independent
    Newton-Euler inverse dynamics), as long as no drive saturates."""
    from tests.hand_common import hand_setup
    m, _, _ = hand_setup()
    dt, sub = 0.01667, 2
    sim = OracleSim(m, dt, sub, G)
    h = dt / sub
    rng = np.random.default_rng(3)
    for trial in range(5):
        root = np.zeros(13); root[2] = 0.5; root[3:7] = m.default_root_quat
        lo, hi = m.lower[1:], m.upper[1:]
        q = lo + (hi - lo) * (0.3 + 0.05 * rng.uniform(-1, 1, size=m.ndof))      # inside the limits ...
        cap = np.where(m.kp[1:] > 0, 0.6 * m.effort[1:] / np.maximum(m.kp[1:], 1e-9), np.inf)
        q = np.sign(q) * np.minimum(np.abs(q), cap)                               # ... and the drives (target 0) unsaturated
        qd = 0.2 * rng.normal(size=m.ndof)
        dof = np.stack([q, qd], -1)
        qdd, ra, da = sim.forward_dynamics(root, dof, np.zeros(m.ndof))         # position targets = 0
        tau_req, _ = rnea_np.inverse_dynamics(m, root, q, qd, qdd, (np.zeros(3), np.zeros(3)), (0.0, 0.0, 0.0))
        qd1 = qd + h * qdd; q1 = q + h * qd1
        pos = m.drive_mode[1:] == 1
        pd = np.where(pos, m.kp[1:] * (0.0 - q1) - m.kd[1:] * qd1, 0.0)
        assert (np.abs(m.kp[1:] * (0.0 - (q + h * qd)) - m.kd[1:] * qd)[pos] < m.effort[1:][pos]).all()      # no saturation here
        inside = (q > lo) & (q < hi)
        assert inside.all()
        applied = pd - m.damping[1:] * qd1 - m.stiffness[1:] * q1 - m.armature[1:] * qdd
        scale = max(1.0, np.abs(applied).max())
        assert np.allclose(tau_req, applied, atol=1e-8 * scale, rtol=1e-7), (trial, np.abs(tau_req - applied).max())


def test_object_collision_exchanges_momentum():
    """A cube thrown at a free-floating Ant in zero gravity: the contact pushes both.  The block-Jacobi coupling (each body
    implicit in its OWN acceleration only) does not apply exactly opposite impulses: the total linear momentum drifts by
    about cn*h*dv (13 % of the incoming momentum at h = 4 ms for this 2 m/s hit, 7 % at 1 ms) and the drift shrinks with the
    step.  Recorded here as a known property of the scheme (DESIGN.md 7b lists the symmetric coupling as next work)."""
    m = load_compiled("ant")
    obj = dict(mass=0.5, inertia=[0.5 * 0.1 ** 2 / 6 * 4] * 3, half=[0.1, 0.1, 0.1], mu=0.5, gravity_on=0)

    def run(dt, T=0.6, coupling=0):
        sim = OracleSim(m, dt, 1, (0.0, 0.0, 0.0), obj=dict(obj, coupling=coupling))
        root = np.zeros((1, 13)); root[0, 2] = 5.0; root[0, 6] = 1
        q0 = np.where(m.lower[1:] > 0, m.lower[1:], np.where(m.upper[1:] < 0, m.upper[1:], 0.0))
        dof = np.zeros((1, m.ndof, 2)); dof[0, :, 0] = q0
        o = np.zeros((1, 13)); o[0, 0:3] = [-0.8, 0.02, 5.03]; o[0, 6] = 1; o[0, 7] = 2.0          # flying at the torso along +x
        P0, _ = rnea_np.momentum(m, root[0], dof[0, :, 0], dof[0, :, 1])
        p0 = P0 + obj["mass"] * o[0, 7:10]
        for _ in range(int(round(T / dt))):
            sim.simulate(root, dof, np.zeros((1, m.ndof)), obj=o)
        P1, _ = rnea_np.momentum(m, root[0], dof[0, :, 0], dof[0, :, 1])
        return p0, P1 + obj["mass"] * o[0, 7:10], P1, o[0, 7:10].copy()
    p0, p1, Pant, vobj = run(0.004)
    assert Pant[0] > 0.3 and vobj[0] < 1.5                       # they did collide: the Ant took momentum, the cube slowed down
    e1 = np.abs(p1 - p0).max()
    _, p2, _, _ = run(0.001)
    e2 = np.abs(p2 - p0).max()
    assert e1 < 0.15 * np.abs(p0).max() and e2 < 0.6 * e1 + 1e-6, (e1, e2)
    # the experimental Gauss-Seidel coupling (object receives the opposite of the applied link forces) does conserve it --
    # but is unstable for the cube in the hand at the task's step, so the engine does not use it
    _, p3, _, _ = run(0.004, coupling=1)
    assert np.abs(p3 - p0).max() < 1e-4
    # extra sweeps of the two block solves on the coupled implicit law stay stable and reduce the drift, slowly
    _, p4, _, _ = run(0.004, coupling=4)
    assert np.abs(p4 - p0).max() < 0.7 * e1


def test_asset_damping_and_angular_speed_clamp():
    """AssetOptions.angular_damping / linear_damping / max_angular_velocity on a free single body (the cube asset, gravity
    off): the twist decays like exp(-d t) (explicit Euler: (1 - d h)^n) and the angular speed is clamped."""
    import copy
    m = copy.deepcopy(load_compiled("cube"))
    m.angular_damping, m.linear_damping, m.max_angular_velocity = 0.5, 0.25, 64.0
    dt, sub = 0.01, 1
    sim = OracleSim(m, dt, sub, (0.0, 0.0, 0.0))
    root = np.zeros((1, 13)); root[0, 6] = 1; root[0, 2] = 5.0
    root[0, 7:10] = [1.0, -2.0, 0.5]; root[0, 10:13] = [0.0, 0.0, 3.0]          # spin about a principal axis
    dof = np.zeros((1, 0, 2))
    n = 100
    for _ in range(n):
        sim.simulate(root, dof, np.zeros((1, 0)))
    assert np.allclose(root[0, 10:13], np.array([0, 0, 3.0]) * (1 - 0.5 * dt) ** n, rtol=1e-6, atol=1e-9)
    assert np.allclose(root[0, 7:10], np.array([1.0, -2.0, 0.5]) * (1 - 0.25 * dt) ** n, rtol=1e-6)
    assert abs((1 - 0.5 * dt) ** n - np.exp(-0.5)) < 2e-3
    root[0, 10:13] = [0.0, 80.0, 60.0]                                            # |w| = 100 > 64
    sim.simulate(root, dof, np.zeros((1, 0)))
    assert abs(np.linalg.norm(root[0, 10:13]) - 64.0) < 1e-9


# ------------------------------------------------------------------------------------ self-collision (collision filter 0)
def _humanoid_self():
    import copy
    from isaacgymenvs_b200.importer.model import enable_self_collision
    m = copy.deepcopy(load_compiled("humanoid"))
    m.angular_damping, m.max_angular_velocity = 0.01, 100.0
    return enable_self_collision(m)


def _max_overlap(m, orc, root, dof):
    """deepest overlap (m) between contact spheres of links that may collide, per env"""
    from oracle import tasks_np as T
    f32 = np.float32
    cpb = np.array(m.cp_body)
    bs = orc.body_states(root, dof)
    off_p = np.asarray(m.body_pos, f32)[cpb]; off_q = np.asarray(m.body_quat, f32)[cpb]
    loc = T.quat_rotate_inverse(off_q, np.asarray(m.cp_pos, f32) - off_p)
    n, ncp = bs.shape[0], len(cpb)
    wp = bs[:, cpb, 0:3].astype(f32) + T.quat_rotate(bs[:, cpb, 3:7].astype(f32).reshape(-1, 4), np.tile(loc, (n, 1))).reshape(n, ncp, 3)
    rr = (np.asarray(m.cp_radius)[:, None] + np.asarray(m.cp_radius)[None, :]).astype(f32)
    d = np.linalg.norm(wp[:, :, None, :] - wp[:, None, :, :], axis=-1)
    return np.where(np.asarray(m.self_pairs)[None] > 0, rr[None] - d, -1.0).max(axis=(1, 2))


def test_self_collision_pair_table():
    """Candidate pairs: symmetric, never within a link or between joint neighbours (incl. through the massless links of a
    compound joint), never between spheres that already overlap at q = 0; the Humanoid's arms may hit the torso, the
    thighs each other."""
    m = _humanoid_self()
    P = np.asarray(m.self_pairs)
    assert P.shape == (len(m.cp_link),) * 2 and (P == P.T).all() and not P.diagonal().any()
    body = lambda n: m.body_names[m.cp_body[n]]
    names = {(body(a), body(b)) for a in range(P.shape[0]) for b in range(P.shape[0]) if P[a, b]}
    assert ("right_lower_arm", "torso") in names and ("right_thigh", "left_thigh") in names and ("left_foot", "right_shin") in names
    for a, b in (("right_thigh", "right_shin"), ("right_shin", "right_foot"), ("torso", "head"), ("torso", "lower_waist"),
                 ("lower_waist", "pelvis"), ("pelvis", "right_thigh"), ("torso", "right_upper_arm"), ("right_lower_arm", "right_hand")):
        assert (a, b) not in names and (b, a) not in names, (a, b)
    assert m.self_kn == 0.5 and m.self_cn == 0.5 and m.self_mu == 1.0         # dimensionless: gains per pair from m_red and h


def test_self_collision_keeps_limbs_apart():
    """A Humanoid thrown around by full-scale random torques: without link-link contact half of the sampled states have
    limbs inside each other by more than 1 cm (deepest > 10 cm); with it the share and the deepest overlap are cut by half
    (a SOFT penalty contact: its gains are bounded by the stability of the half-explicit coupling, not by the actuators'
    1000 N), and the rollout stays finite and no faster than the uncollided one."""
    res = {}
    for on in (False, True):
        m = _humanoid_self()
        m.self_collide = on
        orc = OracleSim(m, 0.0166, 2, G, threads=16)
        n = 128
        rng = np.random.default_rng(0)
        root = np.zeros((n, 13)); root[:, 2] = 1.34; root[:, 6] = 1
        dof = np.zeros((n, m.ndof, 2)); dof[..., 0] = rng.uniform(-0.1, 0.1, size=(n, m.ndof))
        gear = np.asarray(m.actuator_gear)
        hits = tot = 0; worst = 0.0
        for k in range(120):
            orc.simulate(root, dof, rng.uniform(-1, 1, size=(n, m.ndof)) * gear[None])
            if k % 10 == 9:
                dep = _max_overlap(m, orc, root, dof)
                hits += int((dep > 0.01).sum()); tot += n; worst = max(worst, float(dep.max()))
        assert np.isfinite(root).all() and np.isfinite(dof).all()
        res[on] = (hits / tot, worst, float(np.abs(dof[..., 1]).max()))
    assert res[False][0] > 0.3 and res[False][1] > 0.08, res
    assert res[True][0] < 0.6 * res[False][0] and res[True][1] < 0.7 * res[False][1], res
    assert res[True][2] < 1.5 * res[False][2], res


def test_self_collision_is_stable_under_persistent_actuation():
    """What a learner does, not what a random policy does: bang-bang actions held for 25 steps, 512 Humanoids, 600 steps with
    resets on falling.  The first gain choice (the ground contact's, tied to the actor's mass) blew joint speeds up to 1e5
    rad/s here (the block-Jacobi coupling is explicit in the partner's velocity); the per-pair gains from the reduced link
    mass and the sub-step keep the speeds where they are without link-link contact."""
    peak = {}
    for on in (False, True):
        m = _humanoid_self()
        m.self_collide = on
        orc = OracleSim(m, 0.0166, 2, G, threads=16)
        n = 256
        rng = np.random.default_rng(0)
        root = np.zeros((n, 13)); root[:, 2] = 1.34; root[:, 6] = 1
        dof = np.zeros((n, m.ndof, 2)); dof[..., 0] = rng.uniform(-0.2, 0.2, size=(n, m.ndof))
        gear = np.asarray(m.actuator_gear)
        act = rng.uniform(-1, 1, size=(n, m.ndof))
        worst = 0.0
        for k in range(400):
            if k % 25 == 0:
                flip = rng.random(n) < 0.5
                act[flip] = np.sign(rng.uniform(-1, 1, size=(int(flip.sum()), m.ndof)))
            orc.simulate(root, dof, act * gear[None])
            assert np.isfinite(dof).all() and np.isfinite(root).all(), k
            worst = max(worst, float(np.abs(dof[..., 1]).max()))
            fallen = root[:, 2] < 0.6
            root[fallen] = 0; root[fallen, 2] = 1.34; root[fallen, 6] = 1
            dof[fallen] = 0; dof[fallen, :, 0] = rng.uniform(-0.2, 0.2, size=(int(fallen.sum()), m.ndof))
        peak[on] = worst
    assert peak[True] < 1.5 * peak[False] and peak[True] < 200.0, peak


def test_self_contact_forces_are_equal_and_opposite():
    """One isolated contact between two limbs in free flight.  The two links compute their sides of the pair independently
    (block-Jacobi: each implicit in its OWN acceleration only: F_i = F0 - h G J a_i), so the forces reported for the two
    bodies are exactly opposite only in the limit h -> 0 (the gains scale with 1 / h^2 and 1 / h, so the limit is taken with kn, cn FIXED at the task step's values).  Frictionless at h = 20 us they agree to 5 %; at the task's 8.3 ms,
    with friction, each side's force is reduced by its own response (the lighter limb's more) and only the directions still
    oppose -- the same documented property as the hand-object contact."""
    m = _humanoid_self()
    m.gravity_on = False
    m0 = _humanoid_self(); m0.gravity_on = False; m0.self_mu = 0.0
    h_task, h_fine = 0.0166 / 2, 0.00002
    m0.self_kn *= (h_fine / h_task) ** 2; m0.self_cn *= h_fine / h_task          # the same kn, cn in N/m, N s/m as at the task step
    fine = OracleSim(m0, h_fine, 1, (0.0, 0.0, 0.0))
    task = OracleSim(m, 0.0166, 2, (0.0, 0.0, 0.0))
    rng = np.random.default_rng(3)
    found, asym_task = 0, []
    for trial in range(600):
        root = np.zeros((1, 13)); root[0, 2] = 5.0; root[0, 6] = 1
        dof = np.zeros((1, m.ndof, 2)); dof[0, :, 0] = rng.uniform(np.maximum(m.lower[1:], -1.5), np.minimum(m.upper[1:], 1.5))
        dep = _max_overlap(m, fine, root, dof)[0]
        if not (0.005 < dep < 0.03):
            continue
        cf = fine.simulate(root.copy(), dof.copy(), np.zeros((1, m.ndof)))["contact_force"][0]
        touched = np.where(np.linalg.norm(cf, axis=1) > 1e-9)[0]
        if len(touched) != 2:
            continue
        a, b = touched
        assert np.linalg.norm(cf[a] + cf[b]) < 0.05 * np.linalg.norm(cf[a]), (m.body_names[a], m.body_names[b], cf[a], cf[b])
        ct = task.simulate(root.copy(), dof.copy(), np.zeros((1, m.ndof)))["contact_force"][0]
        asym_task.append(np.linalg.norm(ct[a] + ct[b]) / max(np.linalg.norm(ct[a]), np.linalg.norm(ct[b]), 1e-9))
        assert np.dot(ct[a], ct[b]) <= 0                                       # pushing apart (or already separated in the 2nd sub-step)
        found += 1
        if found >= 6:
            break
    assert found >= 3
    print("block-Jacobi force asymmetry at h = 8.3 ms (with friction):", np.round(asym_task, 2))


# ---------------------------------------------------------------------------------------------
# ShadowHand objectType egg / pen: the free object as a ROUNDED box (capsule = segment + radius)
def _obj_of(name):
    from isaacgymenvs_b200.tasks.shadow_hand import object_shape
    om = load_compiled(name)
    half, rnd = object_shape(om)
    return dict(mass=float(om.mass[0]), inertia=[float(om.inertia[0][k]) for k in range(3)], half=half, round=rnd, mu=1.0, gravity_on=1,
                angular_damping=0.5), om


def test_object_shapes_from_the_reference_assets():
    pen, pm = _obj_of("pen"); egg, em = _obj_of("egg"); cube, _ = _obj_of("cube")
    assert pen["half"] == [0.0, 0.0, 0.1] and pen["round"] == 0.008                 # pen.xml:19 capsule size="0.008 0.1"
    assert egg["half"] == [0.0, 0.0, pytest.approx(0.01)] and egg["round"] == 0.03  # egg.xml:10 ellipsoid size="0.03 0.03 0.04"
    assert cube["round"] == 0.0 and cube["half"] == [0.025, 0.025, 0.025]
    # MuJoCo convention: mass = density 1000 x volume
    r, L = 0.008, 0.1
    assert pen["mass"] == pytest.approx(1000 * (np.pi * r * r * 2 * L + 4 / 3 * np.pi * r ** 3), rel=1e-9)
    assert egg["mass"] == pytest.approx(1000 * 4 / 3 * np.pi * 0.03 * 0.03 * 0.04, rel=1e-9)
    # spheroid inertia: I_xx = m (b^2 + c^2) / 5
    assert egg["inertia"][0] == pytest.approx(egg["mass"] * (0.03 ** 2 + 0.04 ** 2) / 5, rel=1e-9)


@pytest.mark.parametrize("name", ["pen", "egg"])
def test_rounded_object_rests_on_the_ground_at_its_radius(name):
    """lying and standing on the plane: the lowest point of the capsule touches, sunk by weight / (contacts x kn)"""
    from tests.hand_common import hand_setup, DT, SUBSTEPS
    m, _, _ = hand_setup()
    obj, _ = _obj_of(name)
    orc = OracleSim(m, DT, SUBSTEPS, G, obj=obj)
    n = 2
    root = np.zeros((n, 13)); root[:, 2] = 0.5; root[:, 3:7] = m.default_root_quat
    dof = np.zeros((n, m.ndof, 2))
    o = np.zeros((n, 13)); o[:, 0] = 1.0; o[:, 2] = 0.2
    o[0, 3:7] = [np.sin(np.pi / 4), 0, 0, np.cos(np.pi / 4)]          # axis horizontal
    o[1, 3:7] = [0, 0, 0, 1]                                          # axis vertical (unstable equilibrium, unperturbed)
    for _ in range(300):
        orc.simulate(root, dof, target=np.zeros((n, m.ndof)), obj=o)
    kn = 10000.0 * obj["mass"]
    sink = obj["mass"] * 9.81 / kn
    L, r = obj["half"][2], obj["round"]
    assert abs(o[0, 2] - (r - sink / 2)) < 1e-6 and abs(o[1, 2] - (L + r - sink)) < 1e-6
    assert np.abs(o[:, 7:]).max() < 1e-9


def test_pen_is_held_by_a_contact_sphere_along_its_whole_length():
    """sphere against the capsule: the closest point of the SEGMENT decides, so a sphere pressing anywhere along the pen meets
    it at distance r_sphere + r_pen (not only near the two end points); beyond the end the end cap's sphere decides"""
    import copy
    from tests.hand_common import hand_setup, DT, SUBSTEPS
    from isaacgymenvs_b200.importer import rot
    m0, _, _ = hand_setup()
    m = copy.deepcopy(m0)
    k = 0                                                          # one contact sphere of the hand, the others removed
    li, b = int(m.cp_link[k]), int(m.cp_body[k])
    for a in ("cp_link", "cp_body", "cp_pos", "cp_radius", "cp_mu"):
        setattr(m, a, np.asarray(getattr(m, a))[k:k + 1].copy())
    m.box_link = None
    obj, _ = _obj_of("pen")
    obj["gravity_on"] = 0
    orc = OracleSim(m, DT, SUBSTEPS, (0.0, 0.0, 0.0), obj=obj)
    root = np.zeros((1, 13)); root[:, 2] = 0.5; root[:, 3:7] = m.default_root_quat
    # the sphere's world centre at q = 0 from the rigid-body state of the body riding on its link (body frame = link frame o offset)
    assert int(m.body_link[b]) == li
    bs = orc.body_states(root, np.zeros((1, m.ndof, 2)))[0, b]
    Rl = rot.quat_to_mat(bs[3:7]) @ rot.quat_to_mat(np.asarray(m.body_quat[b])).T
    c = bs[:3] - Rl @ np.asarray(m.body_pos[b]) + Rl @ np.asarray(m.cp_pos[0])
    rad, L = float(m.cp_radius[0]), obj["half"][2]
    gap = rad + obj["round"] - 0.002                               # 2 mm of overlap; the pen's axis along world z, offset in x
    for along, touches in ((-0.08, True), (0.0, True), (0.06, True), (L + 0.5 * gap, True), (L + 1.01 * (rad + obj["round"]), False)):
        dof = np.zeros((1, m.ndof, 2))
        o = np.zeros((1, 13)); o[0, 3:7] = [0, 0, 0, 1]
        o[0, :3] = c + np.array([gap if along <= L else 0.0, 0.0, along])     # beyond the end: straight above the cap
        orc.simulate(root, dof, target=np.zeros((1, m.ndof)), obj=o)
        v = o[0, 7:10]
        if not touches:
            assert np.abs(v).max() == 0.0
        elif along <= L:
            assert v[0] > 1e-3 and abs(v[2]) < 0.2 * v[0], (along, v)        # pushed away from the sphere, across the axis (friction adds a little along it once it spins)
        else:
            assert v[2] > 1e-3 and abs(v[0]) < 1e-9, (along, v)               # the end cap: pushed along the axis


def test_slender_object_in_a_fast_tumble_stays_bounded():
    """The pen's inertia is 117 : 1; flicked into a tumble of hundreds of rad/s the explicitly integrated gyroscopic term diverged
    (found by the settling run of the GPU parity test: NaN after 22 steps).  The object's AssetOptions.max_angular_velocity
    (gymapi default 64 rad/s, the bound PhysX applies too) is consumed: |w| <= 64 and the flight stays ballistic."""
    from tests.hand_common import hand_setup, DT, SUBSTEPS
    m, _, _ = hand_setup()
    obj, _ = _obj_of("pen")
    obj["max_angular_velocity"] = 64.0
    orc = OracleSim(m, DT, SUBSTEPS, G, obj=obj)
    n = 4
    root = np.zeros((n, 13)); root[:, 2] = 0.5; root[:, 3:7] = m.default_root_quat
    dof = np.zeros((n, m.ndof, 2))
    o = np.zeros((n, 13)); o[:, 0] = 2.0; o[:, 2] = 50.0; o[:, 6] = 1.0
    o[:, 10:13] = [[300.0, 20.0, 150.0], [-50.0, 400.0, 30.0], [10.0, 10.0, 500.0], [60.0, 0.0, 20.0]]
    vz = []
    for k in range(120):
        orc.simulate(root, dof, target=np.zeros((n, m.ndof)), obj=o)
        vz.append(o[:, 9].copy())
        assert np.isfinite(o).all() and np.linalg.norm(o[:, 10:13], axis=1).max() <= 64.0 + 1e-9
    assert np.allclose(vz[-1], -9.81 * DT * 120, rtol=1e-9)
    assert np.allclose(np.linalg.norm(o[:, 3:7], axis=1), 1.0, atol=1e-12)
