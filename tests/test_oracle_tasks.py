"""The numpy obs/reward restatement (oracle/tasks_np.py) against the golden vectors produced by
the reference's own jit functions (tests/golden/make_golden.py).  Tolerance: 1e-6 absolute on
O(1) observation entries (libm-vs-torch transcendental differences are 1-2 ulp); potentials and
integer outputs must match exactly."""
import os
import numpy as np
import pytest
from oracle import tasks_np as T

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
f32 = np.float32


def _consts(n):
    return (np.tile(f32([1000, 0, 0]), (n, 1)), np.tile(f32([0, 0, 0, 1]), (n, 1)),
            np.tile(f32([1, 0, 0]), (n, 1)), np.tile(f32([0, 0, 1]), (n, 1)))


def test_quat_ops():
    g = np.load(os.path.join(G, "quat_ops.npz"))
    qa, qb, v = g["qa"], g["qb"], g["v"]
    assert np.allclose(T.quat_mul(qa, qb), g["quat_mul"], atol=1e-6)
    assert np.allclose(T.quat_rotate(qa, v), g["quat_rotate"], atol=2e-6)
    assert np.allclose(T.quat_rotate_inverse(qa, v), g["quat_rotate_inverse"], atol=2e-6)
    assert np.allclose(T.quat_apply(qa, v), g["quat_apply"], atol=2e-6)
    r, p, y = T.get_euler_xyz(qa)
    for a, b in ((r, g["roll"]), (p, g["pitch"]), (y, g["yaw"])):
        d = np.abs(a - b); d = np.minimum(d, 2 * np.pi - d)
        assert d.max() < 2e-6
    assert np.allclose(T.normalize_angle(v[:, 0] * 3), g["normalize_angle"], atol=1e-6)


def test_ant_obs_reward():
    g = np.load(os.path.join(G, "ant_obs_reward.npz"))
    n = g["root"].shape[0]
    targets, isr, b0, b1 = _consts(n)
    obs, pot, prev, up, head = T.ant_observations(g["root"], targets, g["potentials_in"], isr, g["dof_pos"], g["dof_vel"],
                                                  g["lower"], g["upper"], 0.2, g["sensors"], g["actions"], float(g["dt"]), 0.1, b0, b1)
    assert np.array_equal(pot, g["potentials"]) and np.array_equal(prev, g["prev_potentials"])
    d = np.abs(obs - g["obs"])
    for col in (7, 8, 9):   # yaw, roll (mod 2pi) and angle_to_target may wrap
        d[:, col] = np.minimum(d[:, col], np.abs(2 * np.pi - d[:, col]))
    assert d.max() < 5e-6, (d.max(), np.unravel_index(d.argmax(), d.shape))
    assert np.allclose(up, g["up_vec"], atol=2e-6) and np.allclose(head, g["heading_vec"], atol=2e-6)
    rew, reset = T.ant_reward(g["obs"], np.zeros(n, np.int64), g["progress"], g["actions"], 0.1, 0.5, g["potentials"],
                              g["prev_potentials"], 0.005, 0.05, 0.1, 0.31, -2.0, 1000.0)
    assert np.array_equal(reset, g["reset"])
    assert np.allclose(rew, g["rew"], atol=2e-6 * np.maximum(1, np.abs(g["rew"])).max())


def test_humanoid_obs_reward():
    g = np.load(os.path.join(G, "humanoid_obs_reward.npz"))
    n = g["root"].shape[0]
    targets, isr, b0, b1 = _consts(n)
    obs, pot, prev, up, head = T.humanoid_observations(g["root"], targets, g["potentials_in"], isr, g["dof_pos"], g["dof_vel"],
                                                       g["dof_force"], g["lower"], g["upper"], 0.1, g["sensors"], g["actions"],
                                                       float(g["dt"]), 0.01, 0.25, b0, b1)
    assert np.array_equal(pot, g["potentials"])
    d = np.abs(obs - g["obs"])
    for col in (7, 8, 9):
        d[:, col] = np.minimum(d[:, col], np.abs(2 * np.pi - d[:, col]))
    assert d.max() < 5e-6, (d.max(), np.unravel_index(d.argmax(), d.shape))
    rew, reset = T.humanoid_reward(g["obs"], np.zeros(n, np.int64), g["progress"], g["actions"], 0.1, 0.5, g["potentials"],
                                   g["prev_potentials"], 0.01, 0.05, 0.25, float(g["motor_efforts"].max()), g["motor_efforts"],
                                   0.8, -1.0, 1000.0)
    assert np.array_equal(reset, g["reset"])
    assert np.allclose(rew, g["rew"], atol=1e-5, rtol=1e-5)


def test_cartpole_reward():
    g = np.load(os.path.join(G, "cartpole_reward.npz"))
    n = g["pole_angle"].shape[0]
    rew, reset = T.cartpole_reward(g["pole_angle"], g["pole_vel"], g["cart_vel"], g["cart_pos"], 3.0,
                                   np.zeros(n, np.int64), g["progress"], 500.0)
    assert np.array_equal(reset, g["reset"]) and np.allclose(rew, g["rew"], atol=1e-6)


def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32-10."""
    assert T.philox4x32([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert T.philox4x32([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert T.philox4x32([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_anymal_terrain_restatement_matches_reference_methods():
    """oracle/tasks_np.py anymal_* against the reference's own AnymalTerrain methods
    (tests/golden/make_golden_anymal.py)."""
    g = np.load(os.path.join(G, "anymal_terrain.npz"))
    assert np.array_equal(T.wrap_to_pi(g["wrap_to_pi_in"]), g["wrap_to_pi_out"]) or np.allclose(T.wrap_to_pi(g["wrap_to_pi_in"]), g["wrap_to_pi_out"], atol=1e-6)
    assert T.wrap_to_pi(f32([-4.0]))[0] == f32(-4.0)            # the fmod quirk
    root = g["root"]
    hs = np.repeat(np.repeat(g["height_samples"], 8, 0), 8, 1)
    blv, bav, pg, cmd = T.anymal_prepare(root, g["commands_in"])
    assert np.allclose(blv, g["base_lin_vel"], atol=2e-6) and np.allclose(bav, g["base_ang_vel"], atol=2e-6)
    assert np.allclose(pg, g["projected_gravity"], atol=2e-6) and np.allclose(cmd, g["commands"], atol=2e-6)
    mh = T.anymal_get_heights(root, hs, 20, 0.1, 0.005)
    assert (mh != g["measured_heights"]).mean() < 2e-3          # index truncation at cell borders may flip on 1-ulp differences
    n = root.shape[0]
    dt = 0.02
    max_len = int(20 / dt + 0.5)
    reset = T.anymal_check_termination(g["contact_forces"], g["progress"], max_len)
    assert np.array_equal(reset, g["reset"])
    scales = dict(termination=0.0, lin_vel_xy=1.0, lin_vel_z=-4.0, ang_vel_z=0.5, ang_vel_xy=-0.05, orient=-0.0, torque=-0.00002,
                  joint_acc=-0.0005, base_height=-0.0, air_time=1.0, collision=-0.25, stumble=-0.0, action_rate=-0.01, hip=-0.0)
    rs = {k: v * dt for k, v in scales.items()}
    rew, fat, terms = T.anymal_reward(g["base_lin_vel"], g["base_ang_vel"], g["projected_gravity"], g["commands"], root, g["torques"],
                                      g["last_dof_vel"], g["dof_vel"], g["dof_pos"], g["default_dof_pos"], g["contact_forces"],
                                      g["last_actions"], g["actions"], g["feet_air_time_in"], g["reset"], np.zeros(n, bool), rs, dt)
    assert np.allclose(rew, g["rew"], atol=2e-6) and np.allclose(fat, g["feet_air_time"], atol=1e-7)
    es = np.stack([terms[k] for k in T.ANYMAL_SUM_KEYS])
    assert np.allclose(es, g["episode_sums"], atol=2e-6)
    obs = T.anymal_observations(g["base_lin_vel"], g["base_ang_vel"], g["projected_gravity"], g["commands"], g["dof_pos"], g["dof_vel"],
                                root, g["measured_heights"], g["actions"])
    assert np.allclose(obs, g["obs"], atol=2e-6)


# ---------------------------------------------------------------------------------------------
# ShadowHand: numpy restatement against the reference's own methods (tests/golden/shadow_hand.npz)
@pytest.mark.parametrize("case", ["a", "b", "c", "f", "p"])
def test_hand_step_restatement_matches_reference(case):
    from tests.hand_common import golden_case, hand_setup, DT, SUBSTEPS, G as GRAV
    from oracle.oracle import OracleSim
    gold = np.load(os.path.join(G, {"f": "shadow_hand_force.npz", "p": "shadow_hand_pen.npz"}.get(case, "shadow_hand.npz")))
    obs_types = ["full_state", "full", "full_no_vel", "openai"] if case == "a" else ["full_state"]
    m, obj, tendons = hand_setup()
    orc = OracleSim(m, DT, SUBSTEPS, GRAV, obj=obj, tendons=tendons)
    for ot in obs_types:
        st, P, actions = golden_case(gold, case, ot)
        n = actions.shape[0]
        a = T.hand_pre_physics(st, actions, P)
        out = lambda k: gold[f"{case}_out_{k}"]
        assert np.array_equal(st["root"].reshape(-1, 13)[1::3][:, 7:], out("root")[1::3][:, 7:])
        np.testing.assert_allclose(st["root"].reshape(-1, 13), out("root"), rtol=0, atol=2e-7)
        np.testing.assert_allclose(st["goal_states"], out("goal_states"), rtol=0, atol=2e-7)
        np.testing.assert_allclose(st["cur_targets"], out("cur_targets"), rtol=0, atol=3e-7)
        np.testing.assert_allclose(st["prev_targets"], out("prev_targets"), rtol=0, atol=3e-7)
        if case == "f":       # random forces on the object (:700-709): decayed, zeroed on reset, redrawn where rand < prob
            np.testing.assert_allclose(st["obj_force"], out("obj_force"), rtol=2e-6, atol=1e-8)
            np.testing.assert_allclose(st["force_prob"], out("force_prob"), rtol=2e-6)
            assert float(out("other_forces")) == 0.0 and (out("obj_force") != 0).any(1).sum() > 100
        ds = out("dof_state").reshape(n, -1, 2)
        np.testing.assert_allclose(st["dof_pos"], ds[..., 0], rtol=0, atol=2e-7)
        np.testing.assert_allclose(st["dof_vel"], ds[..., 1], rtol=0, atol=2e-7)
        # post_physics_step: progress, observations (fingertips by the oracle's forward kinematics), reward
        st["progress"] += 1
        root64 = np.ascontiguousarray(st["root"][:, 0].astype(np.float64))
        dof64 = np.ascontiguousarray(np.stack([st["dof_pos"], st["dof_vel"]], -1).astype(np.float64))
        ft = orc.body_states(root64, dof64)[:, gold["fingertips"]].astype(np.float32)
        np.testing.assert_allclose(ft, out("fingertip_state"), rtol=0, atol=1e-6)
        obs = T.hand_observations(st, a, ft, gold[f"{case}_in_sensors"], gold[f"{case}_in_dof_force"], P)
        ref_obs = out("obs") if ot == "full_state" else out(f"obs_{ot}")
        np.testing.assert_allclose(obs, ref_obs, rtol=0, atol=2e-6)
        if ot == "full_state":
            rew, cons = T.hand_reward(st, a, float(gold[f"{case}_in_cons"][0]), P)
            np.testing.assert_allclose(rew, out("rew"), rtol=2e-6, atol=2e-5)
            assert np.array_equal(st["reset"], out("reset")) and np.array_equal(st["reset_goal"], out("reset_goal"))
            assert np.array_equal(st["progress"], out("progress")) and np.array_equal(st["successes"], out("successes"))
            np.testing.assert_allclose(cons, out("cons")[0], rtol=1e-6)
            assert out("reset").sum() > 10 and out("reset_goal").sum() > 10          # the case does exercise both branches
