"""Shared set-up of the ShadowHand + cube environment for the oracle / engine tests."""
import copy
import numpy as np

from isaacgymenvs_b200.assets import load_compiled

FINGERTIPS = ["robot0:ffdistal", "robot0:mfdistal", "robot0:rfdistal", "robot0:lfdistal", "robot0:thdistal"]
RELEVANT_TENDONS = ["robot0:T_FFJ1c", "robot0:T_MFJ1c", "robot0:T_RFJ1c", "robot0:T_LFJ1c"]   # shadow_hand.py:258
DT, SUBSTEPS = 0.01667, 2                                                                       # ShadowHand.yaml sim
G = (0.0, 0.0, -9.81)


def hand_setup():
    m = copy.deepcopy(load_compiled("shadow_hand"))
    cube = load_compiled("cube")
    m.sensor_body = np.array([m.body_names.index(n) for n in FINGERTIPS], dtype=np.int32)
    m.sensor_pos = np.zeros((5, 3)); m.sensor_quat = np.tile([0, 0, 0, 1.0], (5, 1))
    obj = dict(mass=float(cube.mass[0]), inertia=[float(cube.inertia[0][k]) for k in range(3)], half=[0.025] * 3, mu=1.0,
               gravity_on=1)
    tendons = [t for t in m.tendons if t["name"] in RELEVANT_TENDONS]
    return m, obj, tendons


def object_dict(name):
    """the free object of ShadowHand objectType block / egg / pen as the engine and the oracle take it"""
    from isaacgymenvs_b200.tasks.shadow_hand import object_shape
    om = load_compiled({"block": "cube"}.get(name, name))
    half, rnd = object_shape(om)
    return dict(mass=float(om.mass[0]), inertia=[float(om.inertia[0][k]) for k in range(3)], half=half, round=rnd, mu=1.0, gravity_on=1)


def settled_states(n, steps, seed, precision="f64", threads=8, obj_name=None):
    """Contact-rich states: the cube (or egg / pen) dropped onto the hand while the fingers chase random targets (fp64 oracle)."""
    from oracle.oracle import OracleSim
    m, obj, tendons = hand_setup()
    if obj_name is not None:
        obj = object_dict(obj_name)
    rng = np.random.default_rng(seed)
    orc = OracleSim(m, DT, SUBSTEPS, G, precision=precision, obj=obj, tendons=tendons, tendon_k=30.0, tendon_d=0.1, threads=threads)
    dt_ = np.float64 if precision == "f64" else np.float32
    root = np.zeros((n, 13), dt_); root[:, 2] = 0.5; root[:, 3:7] = m.default_root_quat
    dof = np.zeros((n, m.ndof, 2), dt_)
    o = np.zeros((n, 13), dt_)
    o[:, 0:3] = np.array([0.0, -0.39, 0.56]) + rng.uniform(-1, 1, size=(n, 3)) * np.array([0.02, 0.03, 0.02])
    q = rng.normal(size=(n, 4)); o[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    lo, hi = m.lower[1:], m.upper[1:]
    tgt = (lo + (hi - lo) * rng.uniform(0.0, 0.7, size=(n, m.ndof))).astype(dt_)
    for _ in range(steps):
        orc.simulate(root, dof, target=tgt, obj=o)
    return m, obj, tendons, orc, root, dof, o, tgt


# ---------------------------------------------------------------------------------------------
# golden cases of tests/golden/shadow_hand.npz (make_golden_hand.py)
CASES = {"a": dict(relative=False, mcs=0, mavg=1.0, fall_penalty=0.0), "b": dict(relative=True, mcs=50, mavg=1.0, fall_penalty=-50.0),
         "c": dict(relative=False, mcs=0, mavg=0.3, fall_penalty=0.0),
         # tests/golden/shadow_hand_force.npz (make_golden_hand.py --force): random forces on the object, shadow_hand.py:700-709
         "f": dict(relative=False, mcs=0, mavg=1.0, fall_penalty=0.0),
         # tests/golden/shadow_hand_pen.npz (make_golden_hand.py --pen): objectType pen -- randomize_rotation_pen at reset, ignore_z_rot in the reward
         "p": dict(relative=False, mcs=0, mavg=1.0, fall_penalty=0.0)}


def force_constants(gold):
    """the float32 constants the reference forms from its config tensors (shadow_hand.py:196-198,701)"""
    import torch
    pr = torch.tensor(gold["force_prob_range"], dtype=torch.float)
    return dict(force_scale=float(gold["force_scale"]), obj_mass=float(gold["obj_mass"]),
                force_decay_factor=float(torch.pow(torch.tensor(0.99), 0.01667 / 0.08)),
                force_logp_span=float(torch.log(pr[0]) - torch.log(pr[1])), force_logp1=float(torch.log(pr[1])))


def golden_case(gold, case, obs_type="full_state"):
    """-> (st, P, actions): the numpy-restatement view of one golden case's inputs."""
    m, _, _ = hand_setup()
    g = lambda k: gold[f"{case}_in_{k}"].copy()
    n = g("reset").shape[0]
    D = m.ndof
    ds = g("dof_state").reshape(n, D, 2)
    st = dict(root=g("root").reshape(n, 3, 13), dof_pos=np.ascontiguousarray(ds[..., 0]), dof_vel=np.ascontiguousarray(ds[..., 1]),
              cur_targets=g("cur_targets"), prev_targets=g("prev_targets"), goal_states=g("goal_states"), reset=g("reset"),
              reset_goal=g("reset_goal"), progress=g("progress"), successes=g("successes"), reset_count=g("reset_count"),
              goal_reset_count=g("goal_reset_count"))
    kw = CASES[case]
    if case == "f":
        st.update(obj_force=g("obj_force"), force_prob=g("force_prob"))
    P = dict(seed=int(gold["seed"]), goal_init=g("goal_init"), object_init=g("object_init"),
             goal_displacement=np.array([-0.2, -0.06, 0.12], np.float32), reset_position_noise=0.01, reset_dof_pos_noise=0.2,
             reset_dof_vel_noise=0.05, lower=m.lower[1:].astype(np.float32), upper=m.upper[1:].astype(np.float32),
             default_pos=np.zeros(D, np.float32), default_vel=np.zeros(D, np.float32), clip_actions=1.0,
             actuated=gold["actuated"].astype(np.int64), use_relative_control=kw["relative"], dof_speed_scale=20.0, dt=0.01667,
             act_moving_average=kw["mavg"], obs_type=obs_type, vel_obs_scale=0.2, force_torque_obs_scale=10.0,
             dist_reward_scale=-10.0, rot_reward_scale=1.0, rot_eps=0.1, action_penalty_scale=-0.0002, success_tolerance=0.1,
             reach_goal_bonus=250.0, fall_dist=0.24, fall_penalty=kw["fall_penalty"], max_consecutive_successes=kw["mcs"],
             max_episode_length=600.0, av_factor=0.1)
    if case == "f":
        P.update(force_constants(gold))
    if case == "p":
        P.update(object_type="pen", success_tolerance=0.2)      # compute_hand_reward doubles it when ignore_z_rot (:758-759)
    return st, P, g("actions")
