"""GPU parity tests (run on the B200 box): the CUDA engine, called through the C ABI, against the
fp64 CPU oracle (physics) and the reference-generated golden vectors (observation / reward).

Tolerances (fp32 engine vs fp64 oracle; stated per SURVEY.md 8c):
  one control step from identical states : |dq|,|dpos|,|dquat| <= 5e-5, base velocities <= 2e-3 relative, joint
  velocities <= 1e-3 relative for 99.9 % of the DOFs and <= 4e-3 for the worst (stiff contact)
                                           to max(1, |v|) (contact-rich states amplify round-off)
  obs / reward vs the reference's functions: 1e-6 scaled by max(1,|x|) (potentials bit-exact).
"""
import copy
import os
import numpy as np
import pytest
import torch

from isaacgymenvs_b200.assets import load_compiled

pytestmark = pytest.mark.gpu
G = (0.0, 0.0, -9.81)
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _model(name):
    m = copy.deepcopy(load_compiled(name))
    if name == "ant":
        m.sensor_body = np.array([2, 4, 6, 8], dtype=np.int32)
    elif name == "humanoid":
        m.sensor_body = np.array([m.body_names.index("right_foot"), m.body_names.index("left_foot")], dtype=np.int32)
    m.sensor_pos = np.zeros((len(m.sensor_body), 3)); m.sensor_quat = np.tile([0, 0, 0, 1.0], (len(m.sensor_body), 1))
    return m


def _random_states(m, n, rng, zlo, zhi):
    root = np.zeros((n, 13))
    root[:, 0:2] = rng.normal(size=(n, 2))
    root[:, 2] = rng.uniform(zlo, zhi, size=n)
    q = rng.normal(size=(n, 4)) * np.array([0.3, 0.3, 0.3, 0.0]) + np.array([0, 0, 0, 1.0])
    root[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    root[:, 7:13] = rng.normal(size=(n, 6)) * 0.5
    if m.root_fixed:
        root[:, 7:13] = 0; root[:, 3:7] = [0, 0, 0, 1]
    lo = np.where(m.limited[1:] > 0, m.lower[1:], -1.0); hi = np.where(m.limited[1:] > 0, m.upper[1:], 1.0)
    qpos = lo + (hi - lo) * rng.uniform(-0.05, 1.05, size=(n, m.ndof))      # some beyond the limits
    qvel = rng.normal(size=(n, m.ndof))
    return root, np.stack([qpos, qvel], -1)


CASES = {"cartpole": (0.0166, 2, 1.5, 2.5, 200.0), "ant": (0.0166, 2, 0.15, 0.8, 15.0), "humanoid": (0.0166, 2, 0.6, 1.6, 60.0)}


@pytest.mark.parametrize("name", ["cartpole", "ant", "humanoid"])
def test_simulate_matches_oracle(name):
    from isaacgymenvs_b200 import engine
    from oracle.oracle import OracleSim
    m = _model(name)
    dt, sub, zlo, zhi, tscale = CASES[name]
    n = 512
    rng = np.random.default_rng(7)
    root, dof = _random_states(m, n, rng, zlo, zhi)
    tau = rng.uniform(-1, 1, size=(n, m.ndof)) * tscale
    sim = engine.Sim(m, n, dt, sub, G, ground_mu=1.0)
    for slot in (engine.T_FORCE_SENSOR, engine.T_DOF_FORCE, engine.T_NET_CONTACT):
        sim.acquire(slot)
    orc = OracleSim(m, dt, sub, G, ground_mu=1.0, threads=8)
    sim.root_state.copy_(torch.tensor(root, dtype=torch.float32))
    sim.dof_state.copy_(torch.tensor(dof.reshape(-1, 2), dtype=torch.float32))
    sim.dof_actuation.copy_(torch.tensor(tau, dtype=torch.float32))
    # start the oracle from the SAME float32-rounded state
    r64 = sim.root_state.cpu().numpy().astype(np.float64)
    d64 = sim.dof_state.cpu().numpy().astype(np.float64).reshape(n, m.ndof, 2)
    t64 = sim.dof_actuation.cpu().numpy().astype(np.float64)
    sim.simulate(); torch.cuda.synchronize()
    out = orc.simulate(r64, d64, t64)
    rg = sim.root_state.cpu().numpy().astype(np.float64)
    dg = sim.dof_state.cpu().numpy().astype(np.float64).reshape(n, m.ndof, 2)
    if not m.root_fixed:
        assert np.abs(rg[:, :7] - r64[:, :7]).max() < 2e-5
        verr = np.abs(rg[:, 7:] - r64[:, 7:]) / np.maximum(1.0, np.abs(r64[:, 7:]))
        assert verr.max() < 2e-3, verr.max()
    assert np.abs(dg[..., 0] - d64[..., 0]).max() < 5e-5
    qerr = np.abs(dg[..., 1] - d64[..., 1]) / np.maximum(1.0, np.abs(d64[..., 1]))
    # fp32 against fp64 through stiff contacts: 99.9 % of the DOFs within 1e-3, the worst (a foot pressed into the ground) within 4e-3
    assert qerr.max() < 4e-3 and np.quantile(qerr, 0.999) < 1e-3, (qerr.max(), np.quantile(qerr, 0.999))
    # derived outputs of the last sub-step
    if len(m.sensor_body):
        sg = sim.tensors[engine.T_FORCE_SENSOR].cpu().numpy().reshape(n, -1, 6)
        scale = max(1.0, np.abs(out["sensor"]).max())
        assert np.abs(sg - out["sensor"]).max() < 2e-3 * scale
    fg = sim.tensors[engine.T_DOF_FORCE].cpu().numpy().reshape(n, -1)
    assert np.abs(fg - out["dof_force"]).max() < 2e-3 * max(1.0, np.abs(out["dof_force"]).max())
    cg = sim.tensors[engine.T_NET_CONTACT].cpu().numpy().reshape(n, m.nb, 3)
    assert np.abs(cg - out["contact_force"]).max() < 2e-3 * max(1.0, np.abs(out["contact_force"]).max())
    # forward kinematics of the new state
    bs = sim.refresh_rigid_body_state(); torch.cuda.synchronize()
    bo = orc.body_states(r64, d64)
    bg = bs.cpu().numpy().reshape(n, m.nb, 13)
    assert np.abs(bg[..., :3] - bo[..., :3]).max() < 3e-5
    qd = np.minimum(np.abs(bg[..., 3:7] - bo[..., 3:7]).max(-1), np.abs(bg[..., 3:7] + bo[..., 3:7]).max(-1))
    assert qd.max() < 3e-5
    assert (np.abs(bg[..., 7:] - bo[..., 7:]) / np.maximum(1.0, np.abs(bo[..., 7:]))).max() < 2e-3
    sim.close()


@pytest.mark.parametrize("name", ["ant", "humanoid"])
def test_rollout_tracks_oracle(name):
    """30 control steps from rest on the ground under smooth random torques: the fp32 engine stays
    within 2e-3 (positions) of the fp64 oracle -- divergence is round-off, not a model difference."""
    from isaacgymenvs_b200 import engine
    from oracle.oracle import OracleSim
    m = _model(name)
    dt, sub, _, _, tscale = CASES[name]
    n = 64
    rng = np.random.default_rng(3)
    sim = engine.Sim(m, n, dt, sub, G)
    orc = OracleSim(m, dt, sub, G, threads=8)
    root = np.zeros((n, 13)); root[:, 6] = 1; root[:, 2] = 0.5 if name == "ant" else 1.34
    q0 = np.where(m.lower[1:] > 0, m.lower[1:], np.where(m.upper[1:] < 0, m.upper[1:], 0.0))
    dof = np.zeros((n, m.ndof, 2)); dof[..., 0] = q0
    sim.root_state.copy_(torch.tensor(root, dtype=torch.float32))
    sim.dof_state.copy_(torch.tensor(dof.reshape(-1, 2), dtype=torch.float32))
    r64 = sim.root_state.cpu().numpy().astype(np.float64); d64 = sim.dof_state.cpu().numpy().astype(np.float64).reshape(n, m.ndof, 2)
    amp = rng.uniform(-1, 1, size=(n, m.ndof)) * tscale * 0.3
    for k in range(30):
        tau = amp * np.sin(0.3 * k)
        sim.dof_actuation.copy_(torch.tensor(tau, dtype=torch.float32))
        sim.simulate()
        orc.simulate(r64, d64, sim.dof_actuation.cpu().numpy().astype(np.float64))
    torch.cuda.synchronize()
    rg = sim.root_state.cpu().numpy(); dg = sim.dof_state.cpu().numpy().reshape(n, m.ndof, 2)
    assert np.isfinite(rg).all() and np.isfinite(dg).all()
    assert np.median(np.abs(rg[:, :3] - r64[:, :3]).max(1)) < 2e-3
    assert np.median(np.abs(dg[..., 0] - d64[..., 0]).max(1)) < 5e-3
    sim.close()


def _make(task, n, **env_over):
    import isaacgymenvs_b200
    from isaacgymenvs_b200 import config
    cfg = config.builtin_cfg(task, {"sim_device": "cuda:0", "rl_device": "cuda:0"})
    cfg["task"]["env"].update(env_over)
    return isaacgymenvs_b200.make(seed=42, task=task, num_envs=n, sim_device="cuda:0", rl_device="cuda:0",
                                  headless=True, cfg=cfg)


@pytest.mark.parametrize("task", ["Ant", "Humanoid"])
def test_obs_reward_match_reference_golden(task):
    """control_freq_inv=0 (no simulate): the kernel's observation/reward epilogue on the golden
    inputs must reproduce what the reference's jit functions returned."""
    g = np.load(os.path.join(GOLD, f"{task.lower()}_obs_reward.npz"))
    n = g["root"].shape[0]
    env = _make(task, n, controlFrequencyInv=0)
    dev = env.device
    t = lambda a, dt=torch.float32: torch.tensor(a, dtype=dt, device=dev)
    env.root_states.copy_(t(g["root"]))
    env.dof_pos.copy_(t(g["dof_pos"])); env.dof_vel.copy_(t(g["dof_vel"]))
    env.vec_sensor_tensor.copy_(t(g["sensors"]))
    if task == "Humanoid":
        env.dof_force_tensor.copy_(t(g["dof_force"]))
    env.potentials.copy_(t(g["potentials_in"]))
    env.progress_buf.copy_(t(g["progress"] - 1, torch.long))
    env.reset_buf.zero_()
    obs, rew, reset, extras = env.step(t(g["actions"]))
    torch.cuda.synchronize()
    assert np.array_equal(env.potentials.cpu().numpy(), g["potentials"])           # bit-exact
    assert np.array_equal(env.prev_potentials.cpu().numpy(), g["prev_potentials"])
    o = obs["obs"].cpu().numpy()
    d = np.abs(o - g["obs"])
    for col in (7, 8, 9):
        d[:, col] = np.minimum(d[:, col], np.abs(2 * np.pi - d[:, col]))
    d = d / np.maximum(1.0, np.abs(g["obs"]))
    assert d.max() < 2e-6, (d.max(), np.unravel_index(d.argmax(), d.shape))
    assert np.array_equal(reset.cpu().numpy(), g["reset"])
    r = rew.cpu().numpy()
    assert (np.abs(r - g["rew"]) / np.maximum(1.0, np.abs(g["rew"]))).max() < 5e-6
    assert np.allclose(env.up_vec.cpu().numpy(), g["up_vec"], atol=2e-6)
    assert np.allclose(env.heading_vec.cpu().numpy(), g["heading_vec"], atol=2e-6)
    expect_to = (g["progress"] >= 999) & (g["reset"] != 0)
    assert np.array_equal(extras["time_outs"].cpu().numpy(), expect_to)


def test_cartpole_reward_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "cartpole_reward.npz"))
    n = g["pole_angle"].shape[0]
    env = _make("Cartpole", n, controlFrequencyInv=0)
    dev = env.device
    t = lambda a, dt=torch.float32: torch.tensor(a, dtype=dt, device=dev)
    env.dof_pos[:, 0] = t(g["cart_pos"]); env.dof_vel[:, 0] = t(g["cart_vel"])
    env.dof_pos[:, 1] = t(g["pole_angle"]); env.dof_vel[:, 1] = t(g["pole_vel"])
    env.progress_buf.copy_(t(g["progress"] - 1, torch.long)); env.reset_buf.zero_()
    obs, rew, reset, _ = env.step(torch.zeros(n, 1, device=dev))
    assert np.array_equal(reset.cpu().numpy(), g["reset"])
    assert np.allclose(rew.cpu().numpy(), g["rew"], atol=1e-6)
    o = obs["obs"].cpu().numpy()
    expect = np.clip(np.stack([g["cart_pos"], g["cart_vel"], g["pole_angle"], g["pole_vel"]], -1), -5, 5)   # clipObservations 5
    assert np.allclose(o, expect, atol=0)


def test_fused_ant_step_equals_oracle_pipeline():
    """A whole VecTask.step: first step resets every env (reset_buf starts as ones,
    vec_task.py:316) with the Philox stream the oracle restates; following steps must equal
    oracle physics + the numpy restatement of the reference's obs/reward."""
    from oracle.oracle import OracleSim
    from oracle import tasks_np as T
    n = 256
    env = _make("Ant", n)
    m = env.model
    orc = OracleSim(m, 0.0166, 2, G, ground_mu=1.0, threads=8)
    rng = np.random.default_rng(0)
    lo, hi = env.dof_limits_lower_np, env.dof_limits_upper_np
    init = np.where(lo > 0, lo, np.where(hi < 0, hi, 0)).astype(np.float32)
    f32 = np.float32
    # --- step 0: physics from the spawn pose, then reset of all envs
    a0 = rng.uniform(-1.5, 1.5, size=(n, 8)).astype(f32)
    obs, rew, reset, _ = env.step(torch.tensor(a0, device=env.device))
    torch.cuda.synchronize()
    q = env.dof_pos.cpu().numpy(); qd = env.dof_vel.cpu().numpy()
    for e in (0, 1, n - 1):
        u = T.reset_uniforms(42, e, 0, 16)
        pos = np.clip(init + (f32(0.4) * u[:8] + f32(-0.2)), lo, hi)
        vel = f32(0.2) * u[8:] + f32(-0.1)
        assert np.allclose(q[e], pos, atol=1e-7) and np.allclose(qd[e], vel, atol=1e-7)
    assert (q >= lo - 1e-6).all() and (q <= hi + 1e-6).all()
    assert np.allclose(env.root_states.cpu().numpy(), env.initial_root_states.cpu().numpy())
    assert (env.progress_buf == 0).all() and (env.reset_count == 1).all()
    assert np.allclose(obs["obs"][:, 52:60].cpu().numpy(), np.clip(a0, -1, 1))
    # --- steps 1..5 against the oracle pipeline
    for k in range(5):
        r64 = env.root_states.cpu().numpy().astype(np.float64)
        d64 = env.dof_state.cpu().numpy().astype(np.float64).reshape(n, 8, 2)
        pot_in = env.potentials.cpu().numpy().copy()
        prog_in = env.progress_buf.cpu().numpy().copy()
        a = rng.uniform(-1.5, 1.5, size=(n, 8)).astype(f32)
        ac = np.clip(a, -1, 1)
        out = orc.simulate(r64, d64, (ac * f32(15.0)).astype(np.float64))
        obs, rew, reset, _ = env.step(torch.tensor(a, device=env.device))
        torch.cuda.synchronize()
        n_ = n
        targets = np.tile(f32([1000, 0, 0]), (n_, 1)); isr = np.tile(f32([0, 0, 0, 1]), (n_, 1))
        b0 = np.tile(f32([1, 0, 0]), (n_, 1)); b1 = np.tile(f32([0, 0, 1]), (n_, 1))
        # evaluate the reference arithmetic on the ENGINE's own state (isolates the epilogue) ...
        rg = env.root_states.cpu().numpy(); qg = env.dof_pos.cpu().numpy(); vg = env.dof_vel.cpu().numpy()
        sg = env.vec_sensor_tensor.cpu().numpy()
        o_np, pot, prev, _, _ = T.ant_observations(rg, targets, pot_in, isr, qg, vg, lo, hi, 0.2, sg, ac, 0.0166, 0.1, b0, b1)
        og = obs["obs"].cpu().numpy()
        d = np.abs(og - o_np)
        for col in (7, 8, 9):
            d[:, col] = np.minimum(d[:, col], np.abs(2 * np.pi - d[:, col]))
        assert (d / np.maximum(1, np.abs(o_np))).max() < 2e-6
        assert np.array_equal(env.potentials.cpu().numpy(), pot)
        r_np, reset_np = T.ant_reward(og, np.zeros(n, np.int64), prog_in + 1, ac, 0.1, 0.5, pot, prev, 0.005, 0.05, 0.1, 0.31, -2.0, 1000.0)
        assert np.array_equal(reset.cpu().numpy(), reset_np)
        assert (np.abs(rew.cpu().numpy() - r_np) / np.maximum(1, np.abs(r_np))).max() < 1e-5
        # ... and the physics against the fp64 oracle
        assert np.abs(rg[:, :7] - r64[:, :7]).max() < 5e-5
        assert np.abs(qg - d64[..., 0]).max() < 5e-5
        assert np.abs(sg.reshape(n, 4, 6) - out["sensor"]).max() < 5e-3 * max(1.0, np.abs(out["sensor"]).max())
        if reset_np.any():
            break


def test_single_lane_and_four_lane_ant_agree():
    """The 4-lanes-per-env kernel and the one-thread-per-env kernel are the same arithmetic in a
    different order: results agree to fp32 round-off."""
    from isaacgymenvs_b200 import engine
    m = _model("ant")
    n = 256
    rng = np.random.default_rng(5)
    root, dof = _random_states(m, n, rng, 0.15, 0.8)
    tau = rng.uniform(-15, 15, size=(n, 8))
    res = []
    for single in ("0", "1"):
        os.environ["B2G_SINGLE_LANE"] = single
        sim = engine.Sim(m, n, 0.0166, 2, G)
        sim.root_state.copy_(torch.tensor(root, dtype=torch.float32)); sim.dof_state.copy_(torch.tensor(dof.reshape(-1, 2), dtype=torch.float32))
        sim.dof_actuation.copy_(torch.tensor(tau, dtype=torch.float32))
        sim.simulate(); torch.cuda.synchronize()
        res.append((sim.root_state.cpu().numpy(), sim.dof_state.cpu().numpy()))
        sim.close()
    os.environ.pop("B2G_SINGLE_LANE")
    rel = lambda a, b: (np.abs(a - b) / np.maximum(1.0, np.abs(b))).max()
    assert rel(res[0][0], res[1][0]) < 2e-3 and rel(res[0][1], res[1][1]) < 5e-3


@pytest.mark.parametrize("task,nact,nobs,pinned", [("Ant", 8, 60, False), ("Ant", 8, 60, True), ("Humanoid", 21, 108, True)])
def test_host_buffer_step_and_launch_count(task, nact, nobs, pinned):
    """b2g_task_step_host: pageable buffers go through staged copies; pinned buffers are read / written by the kernel
    itself over PCIe (no copy launches).  Either way the host sees exactly what the device buffers hold."""
    n = 256
    env = _make(task, n)
    pin = (lambda t: t.pin_memory()) if pinned else (lambda t: t)
    h_obs = pin(torch.zeros(n, nobs)); h_rew = pin(torch.zeros(n)); h_reset = pin(torch.zeros(n, dtype=torch.long))
    h_to = pin(torch.zeros(n, dtype=torch.uint8))
    g = torch.Generator().manual_seed(1)
    for k in range(5):
        h_a = pin(torch.rand(n, nact, generator=g) * 2 - 1)
        c0 = env.sim.launch_count()
        env.step_host(h_a, h_obs, h_rew, h_reset, h_to)
        assert env.sim.launch_count() == c0 + 1
        assert torch.equal(h_obs, env.obs_clipped.cpu()) and torch.equal(h_rew, env.rew_buf.cpu())
        assert torch.equal(h_reset, env.reset_buf.cpu()) and torch.equal(h_to.bool(), env.timeout_buf.cpu())
        assert torch.equal(env.actions.cpu(), h_a.clamp(-1, 1))
    assert h_obs.abs().sum() > 0 and (env.progress_buf == 4).all()


# ------------------------------------------------------------------------------------ AnymalTerrain
def _make_anymal(n, terrain=None, **over):
    import isaacgymenvs_b200
    from isaacgymenvs_b200 import config
    cfg = config.builtin_cfg("AnymalTerrain", {"sim_device": "cuda:0", "rl_device": "cuda:0"})
    e = cfg["task"]["env"]
    if terrain:
        e["terrain"].update(terrain)
    for k, v in over.items():
        if k in e["learn"]:
            e["learn"][k] = v
        elif k in e["control"]:
            e["control"][k] = v
        else:
            e[k] = v
    return isaacgymenvs_b200.make(seed=42, task="AnymalTerrain", num_envs=n, sim_device="cuda:0", rl_device="cuda:0",
                                  headless=True, cfg=cfg)


def test_anymal_terrain_epilogue_matches_reference_golden():
    """No simulate (decimation = controlFrequencyInv = 0): termination, the 13 reward terms, feet air
    time, commands, height sampling and the 188-d observation against what the reference's own
    AnymalTerrain methods returned on the same inputs (tests/golden/make_golden_anymal.py)."""
    g = np.load(os.path.join(GOLD, "anymal_terrain.npz"))
    n = g["root"].shape[0]
    hs = np.repeat(np.repeat(g["height_samples"], 8, 0), 8, 1)
    env = _make_anymal(n, terrain={"heightSamplesOverride": hs}, skipPhysics=True, addNoise=False, pushRobots=False)
    dev = env.device
    t = lambda a, dt=torch.float32: torch.tensor(a, dtype=dt, device=dev)
    env.root_states.copy_(t(g["root"]))
    env.dof_pos.copy_(t(g["dof_pos"])); env.dof_vel.copy_(t(g["dof_vel"]))
    env.contact_forces.copy_(t(g["contact_forces"]))
    env.commands.copy_(t(g["commands_in"]))
    env.last_actions.copy_(t(g["last_actions"])); env.last_dof_vel.copy_(t(g["last_dof_vel"]))
    env.torques.copy_(t(g["torques"])); env.feet_air_time.copy_(t(g["feet_air_time_in"]))
    env.progress_buf.copy_(t(g["progress"] - 1, torch.long))
    env._episode_sums.zero_()
    assert np.allclose(env.default_dof_pos[0].cpu().numpy(), g["default_dof_pos"])
    obs, rew, reset, extras = env.step(t(g["actions"]))
    torch.cuda.synchronize()
    keep = g["reset"] == 0
    assert np.array_equal(reset.cpu().numpy(), g["reset"])
    assert np.allclose(rew.cpu().numpy(), g["rew"], atol=3e-6)
    assert np.allclose(env._base_scratch[:, :9].cpu().numpy(), np.concatenate([g["base_lin_vel"], g["base_ang_vel"], g["projected_gravity"]], 1), atol=2e-6)
    assert np.allclose(env.feet_air_time.cpu().numpy()[keep], g["feet_air_time"][keep], atol=1e-7)
    assert np.allclose(env.commands.cpu().numpy()[keep], g["commands"][keep], atol=2e-6)
    assert np.allclose(env._episode_sums.cpu().numpy()[:, keep], g["episode_sums"][:, keep], atol=3e-6)
    o = obs["obs"].cpu().numpy()[keep]; go = g["obs"][keep]
    bad = np.abs(o - go) > 3e-6
    assert bad[:, :36].sum() == 0 and bad[:, 176:].sum() == 0
    assert bad[:, 36:176].mean() < 2e-3          # height samples: index truncation at a cell border may flip on 1 ulp
    assert np.allclose(env.last_actions.cpu().numpy(), np.clip(g["actions"], -1e30, 1e30))
    # envs that were reset: progress 0, reset flag kept at 1, fresh commands inside their ranges
    r = ~keep
    assert (env.progress_buf.cpu().numpy()[r] == 0).all() and r.any()
    c = env.commands.cpu().numpy()[r]
    assert (np.abs(c[:, :2]) <= 1.0).all() and (np.abs(c[:, 3]) <= 3.14).all()


def test_anymal_heightfield_contact_matches_oracle():
    """gym.simulate on a height field: engine vs fp64 oracle from identical states."""
    from isaacgymenvs_b200 import engine
    from oracle.oracle import OracleSim
    m = copy.deepcopy(load_compiled("anymal"))
    n = 256
    rng = np.random.default_rng(11)
    hf = (rng.normal(size=(40, 40)) * 12).round().astype(np.int16)
    hf = np.repeat(np.repeat(hf, 4, 0), 4, 1)                    # 160 x 160 samples of 0.1 m, blocks of 0.4 m
    hscale, vscale, org = 0.1, 0.005, (-8.0, -8.0)
    root, dof = _random_states(m, n, rng, 0.35, 0.75)
    root[:, 0:2] = rng.uniform(-5, 5, size=(n, 2))
    tau = rng.uniform(-40, 40, size=(n, 12))
    sim = engine.Sim(m, n, 0.005, 1, G, ground_mu=1.0, hfield=hf, hf_horizontal_scale=hscale, hf_vertical_scale=vscale, hf_origin=org)
    sim.acquire(engine.T_NET_CONTACT)
    orc = OracleSim(m, 0.005, 1, G, ground_mu=1.0, threads=8, hfield=hf.astype(np.float64) * vscale, hf_scale=hscale, hf_origin=org)
    sim.root_state.copy_(torch.tensor(root, dtype=torch.float32)); sim.dof_state.copy_(torch.tensor(dof.reshape(-1, 2), dtype=torch.float32))
    sim.dof_actuation.copy_(torch.tensor(tau, dtype=torch.float32))
    r64 = sim.root_state.cpu().numpy().astype(np.float64); d64 = sim.dof_state.cpu().numpy().astype(np.float64).reshape(n, 12, 2)
    for _ in range(3):
        sim.simulate()
        out = orc.simulate(r64, d64, sim.dof_actuation.cpu().numpy().astype(np.float64))
    torch.cuda.synchronize()
    rg = sim.root_state.cpu().numpy(); dg = sim.dof_state.cpu().numpy().reshape(n, 12, 2)
    assert (out["contact_force"][:, :, 2] > 1.0).any()           # contacts did happen
    assert np.abs(rg[:, :7] - r64[:, :7]).max() < 3e-4
    assert (np.abs(rg[:, 7:] - r64[:, 7:]) / np.maximum(1, np.abs(r64[:, 7:]))).max() < 5e-3
    assert np.abs(dg[..., 0] - d64[..., 0]).max() < 3e-4
    cg = sim.tensors[engine.T_NET_CONTACT].cpu().numpy().reshape(n, m.nb, 3)
    assert np.abs(cg - out["contact_force"]).max() < 5e-3 * max(1.0, np.abs(out["contact_force"]).max())
    sim.close()


def test_anymal_terrain_rollout_is_sane():
    """Zero actions on flat ground: the PD loop holds the default pose, the robot stands (base height,
    contact forces carry the weight, no termination); on the curriculum terrain with random actions the
    rollout stays finite and resets re-spawn robots on their tile."""
    env = _make_anymal(256, terrain={"terrainType": "plane"}, addNoise=False, pushRobots=False)
    z = torch.zeros(256, 12, device=env.device)
    for _ in range(150):
        obs, rew, reset, _ = env.step(z)
    torch.cuda.synchronize()
    h = env.root_states[:, 2].cpu().numpy()
    assert np.isfinite(obs["obs"].cpu().numpy()).all()
    assert (h > 0.40).all() and (h < 0.70).all(), (h.min(), h.max())
    W = env.model.total_mass() * 9.81
    fz = env.contact_forces[:, env.feet_indices, 2].sum(1).cpu().numpy()
    assert np.abs(np.median(fz) - W) / W < 0.1
    assert reset.sum().item() == 0
    assert (rew.cpu().numpy() >= 0).all()
    env2 = _make_anymal(512)
    g = torch.Generator(device=env2.device).manual_seed(0)
    nres = 0
    for _ in range(120):
        obs, rew, reset, extras = env2.step(2 * torch.rand(512, 12, device=env2.device, generator=g) - 1)
        nres += int(reset.sum().item())
    torch.cuda.synchronize()
    assert torch.isfinite(obs["obs"]).all() and torch.isfinite(env2.root_states).all()
    assert nres > 0
    lv = env2.terrain_levels.cpu().numpy()
    assert (lv >= 0).all() and (lv < 10).all()
    d = (env2.root_states[:, :2] - env2.env_origins[:, :2]).norm(dim=1).cpu().numpy()
    assert (d < 12.0).all()
    assert "rew_lin_vel_xy" in extras["episode"] and "terrain_level" in extras["episode"]


def test_generic_gym_api_path_matches_fused_step():
    """The compatibility path (isaacgym shim -> b2g_simulate) and the fused Ant step run the same
    physics: driven the way the reference's ant.py drives `gym` (forces = actions * gear ->
    set_dof_actuation_force_tensor -> simulate -> refresh), states and force sensors agree."""
    from isaacgymenvs_b200 import compat
    compat.install()
    from isaacgym import gymapi, gymtorch
    n = 128
    env = _make("Ant", n)
    g = torch.Generator(device=env.device).manual_seed(3)
    env.step(2 * torch.rand(n, 8, device=env.device, generator=g) - 1)          # resets everything
    gym = gymapi.acquire_gym()
    sp = gymapi.SimParams(); sp.dt, sp.substeps, sp.up_axis, sp.gravity, sp.use_gpu_pipeline = 0.0166, 2, gymapi.UP_AXIS_Z, gymapi.Vec3(0, 0, -9.81), True
    sim = gym.create_sim(0, -1, gymapi.SIM_PHYSX, sp)
    pp = gymapi.PlaneParams(); pp.static_friction = pp.dynamic_friction = 1.0
    gym.add_ground(sim, pp)
    ao = gymapi.AssetOptions(); ao.angular_damping = 0.0                        # as ant.py:151 sets it (the AssetOptions default is 0.5)
    asset = gym.load_asset(sim, "/no/such/checkout/assets/mjcf", "nv_ant.xml", ao)
    assert gym.get_asset_dof_count(asset) == 8 and gym.get_asset_rigid_body_count(asset) == 9
    gears = torch.tensor([p.motor_effort for p in gym.get_asset_actuator_properties(asset)], device=env.device)
    for name in [s for s in gym.get_asset_rigid_body_names(asset) if "foot" in s]:
        gym.create_asset_force_sensor(asset, gym.find_asset_rigid_body_index(asset, name), gymapi.Transform())
    pose = gymapi.Transform(); pose.p = gymapi.Vec3(0, 0, 0.44)
    for i in range(n):
        e = gym.create_env(sim, gymapi.Vec3(-5, -5, 0), gymapi.Vec3(5, 5, 5), 8)
        gym.create_actor(e, asset, pose, "ant", i, 1, 0)
    gym.prepare_sim(sim)
    root = gymtorch.wrap_tensor(gym.acquire_actor_root_state_tensor(sim))
    dof = gymtorch.wrap_tensor(gym.acquire_dof_state_tensor(sim))
    sens = gymtorch.wrap_tensor(gym.acquire_force_sensor_tensor(sim))
    assert abs(float(root[0, 2]) - 0.44) < 1e-6
    for _ in range(4):
        root.copy_(env.root_states); dof.copy_(env.dof_state)
        a = 2 * torch.rand(n, 8, device=env.device, generator=g) - 1
        gym.set_dof_actuation_force_tensor(sim, gymtorch.unwrap_tensor(a * gears * 1.0))
        gym.simulate(sim)
        gym.refresh_dof_state_tensor(sim); gym.refresh_actor_root_state_tensor(sim); gym.refresh_force_sensor_tensor(sim)
        obs, rew, reset, _ = env.step(a)
        torch.cuda.synchronize()
        keep = (reset == 0) | True        # a reset flag raised now only takes effect in the NEXT step
        # two instantiations of the same stepper (different I/O staging): equal up to fp32 contraction order
        assert torch.allclose(root[keep], env.root_states[keep], atol=1e-5, rtol=1e-4)
        assert torch.allclose(dof.view(n, 8, 2)[keep], env.dof_state.view(n, 8, 2)[keep], atol=1e-4, rtol=1e-3)
        assert torch.allclose(sens.view(n, 24), env.vec_sensor_tensor, atol=1e-2, rtol=1e-3)


# ---------------------------------------------------------------------------------------------
# ShadowHand: fixed-base 24-DOF hand + free cube (second actor) + goal marker (third actor)
def _hand_sim(n, m, obj, tendons):
    from isaacgymenvs_b200 import engine
    from tests.hand_common import DT, SUBSTEPS, G as HG
    ext = engine.pack_model_ext(m, obj=obj, actors_per_env=3, tendons=tendons, tendon_k=30.0, tendon_d=0.1)
    sim = engine.Sim(m, n, DT, SUBSTEPS, HG, ground_mu=1.0, ext=ext)
    for slot in (engine.T_FORCE_SENSOR, engine.T_DOF_FORCE, engine.T_NET_CONTACT):
        sim.acquire(slot)
    return sim


def _hand_load(sim, root, dof, o, tgt):
    n = root.shape[0]
    rs = np.zeros((n, 3, 13), np.float32); rs[:, 0] = root; rs[:, 1] = o; rs[:, 2, 6] = 1; rs[:, 2, 0:3] = [-0.2, -0.45, 0.68]
    sim.root_state.copy_(torch.tensor(rs.reshape(-1, 13)))
    sim.dof_state.copy_(torch.tensor(dof.reshape(-1, 2), dtype=torch.float32))
    sim.dof_target.copy_(torch.tensor(tgt, dtype=torch.float32))


def test_hand_object_simulate_matches_oracle():
    """One gym.simulate() from contact-rich hand+cube states (cube on the palm / between fingers / on the ground):
    engine (world-axes ABA about O, fp32, 4 lanes) against the oracle (body-coordinate ABA, fp64)."""
    from isaacgymenvs_b200 import engine
    from tests.hand_common import settled_states
    n = 512
    m, obj, tendons, orc, root, dof, o, tgt = settled_states(n, 40, 5)
    sim = _hand_sim(n, m, obj, tendons)
    _hand_load(sim, root, dof, o, tgt)
    rs = sim.root_state.cpu().numpy().astype(np.float64).reshape(n, 3, 13)
    r64 = np.ascontiguousarray(rs[:, 0]); o64 = np.ascontiguousarray(rs[:, 1])
    d64 = sim.dof_state.cpu().numpy().astype(np.float64).reshape(n, m.ndof, 2)
    t64 = sim.dof_target.cpu().numpy().astype(np.float64)
    sim.simulate(); torch.cuda.synchronize()
    out = orc.simulate(r64, d64, target=t64, obj=o64)
    rg = sim.root_state.cpu().numpy().astype(np.float64).reshape(n, 3, 13)
    dg = sim.dof_state.cpu().numpy().astype(np.float64).reshape(n, m.ndof, 2)
    assert np.array_equal(rg[:, 0], rs[:, 0]) and np.array_equal(rg[:, 2], rs[:, 2])      # fixed base, goal marker untouched
    assert np.abs(rg[:, 1, :3] - o64[:, :3]).max() < 5e-5
    qd = np.minimum(np.abs(rg[:, 1, 3:7] - o64[:, 3:7]).max(-1), np.abs(rg[:, 1, 3:7] + o64[:, 3:7]).max(-1))
    assert qd.max() < 2e-4, qd.max()
    verr = np.abs(rg[:, 1, 7:] - o64[:, 7:]) / np.maximum(1.0, np.abs(o64[:, 7:]))
    assert verr.max() < 5e-3, verr.max()
    assert np.abs(dg[..., 0] - d64[..., 0]).max() < 1e-4
    qerr = np.abs(dg[..., 1] - d64[..., 1]) / np.maximum(1.0, np.abs(d64[..., 1]))
    assert qerr.max() < 5e-3, qerr.max()
    sg = sim.tensors[engine.T_FORCE_SENSOR].cpu().numpy().reshape(n, -1, 6)
    assert np.abs(out["sensor"]).max() > 0.1                                              # the fingertips do touch the cube
    assert np.abs(sg - out["sensor"]).max() < 5e-3 * max(1.0, np.abs(out["sensor"]).max())
    fg = sim.tensors[engine.T_DOF_FORCE].cpu().numpy().reshape(n, -1)
    assert np.abs(fg - out["dof_force"]).max() < 5e-3 * max(1.0, np.abs(out["dof_force"]).max())
    # net contact force per LINK (the engine reports a link's contacts on its first body, the oracle per body)
    cg = sim.tensors[engine.T_NET_CONTACT].cpu().numpy().reshape(n, m.nb, 3)
    agg = lambda a: np.stack([a[:, np.nonzero(m.body_link == li)[0]].sum(1) for li in range(m.nl)], 1)
    assert np.abs(out["contact_force"]).max() > 0.1
    assert np.abs(agg(cg) - agg(out["contact_force"])).max() < 5e-3 * max(1.0, np.abs(out["contact_force"]).max())
    # fingertip / object / goal rows of the rigid-body state tensor (shadow_hand.py:456)
    bs = sim.refresh_rigid_body_state(); torch.cuda.synchronize()
    bg = bs.cpu().numpy().reshape(n, m.nb + 2, 13)
    bo = orc.body_states(r64, d64)
    assert np.abs(bg[:, :m.nb, :3] - bo[..., :3]).max() < 3e-5
    assert np.array_equal(bg[:, m.nb], sim.root_state.cpu().numpy().reshape(n, 3, 13)[:, 1])
    sim.close()


def test_hand_rollout_tracks_oracle():
    """60 control steps of the cube dropped on the hand: medians stay together (round-off, not a model difference)."""
    from tests.hand_common import settled_states
    n = 128
    m, obj, tendons, orc, root, dof, o, tgt = settled_states(n, 0, 11)
    sim = _hand_sim(n, m, obj, tendons)
    _hand_load(sim, root, dof, o, tgt)
    rs = sim.root_state.cpu().numpy().astype(np.float64).reshape(n, 3, 13)
    r64 = np.ascontiguousarray(rs[:, 0]); o64 = np.ascontiguousarray(rs[:, 1])
    d64 = sim.dof_state.cpu().numpy().astype(np.float64).reshape(n, m.ndof, 2)
    t64 = sim.dof_target.cpu().numpy().astype(np.float64)
    for _ in range(60):
        sim.simulate()
        orc.simulate(r64, d64, target=t64, obj=o64)
    torch.cuda.synchronize()
    rg = sim.root_state.cpu().numpy().reshape(n, 3, 13); dg = sim.dof_state.cpu().numpy().reshape(n, m.ndof, 2)
    assert np.isfinite(rg).all() and np.isfinite(dg).all()
    assert np.median(np.abs(rg[:, 1, :3] - o64[:, :3]).max(1)) < 2e-3
    assert np.median(np.abs(dg[..., 0] - d64[..., 0]).max(1)) < 5e-3
    sim.close()


def _hand_env(n, case, obs_type, **more):
    from tests.hand_common import CASES as HC
    kw = HC[case]
    return _make("ShadowHand", n, controlFrequencyInv=0, observationType=obs_type, useRelativeControl=kw["relative"], **more,
                 maxConsecutiveSuccesses=kw["mcs"], actionsMovingAverage=kw["mavg"], fallPenalty=kw["fall_penalty"],
                 resetDofVelRandomInterval=0.05)


@pytest.mark.parametrize("case,obs_type", [("a", "full_state"), ("a", "full"), ("a", "full_no_vel"), ("a", "openai"),
                                           ("b", "full_state"), ("c", "full_state")])
def test_hand_step_matches_reference_golden(case, obs_type):
    """One fused ShadowHand step without physics (controlFrequencyInv=0) from the golden inputs: reset_idx /
    reset_target_pose / targets / observations / compute_hand_reward against the reference's own methods."""
    gold = np.load(os.path.join(GOLD, "shadow_hand.npz"))
    gi = lambda k: gold[f"{case}_in_{k}"]
    go = lambda k: gold[f"{case}_out_{k}"]
    n = gi("reset").shape[0]
    env = _hand_env(n, case, obs_type)
    assert env.seed == int(gold["seed"])
    assert np.array_equal(env.actuated_dof_indices_np, gold["actuated"]) and np.array_equal(env.fingertip_handles_np, gold["fingertips"])
    dev = env.device
    t = lambda a, dt=torch.float32: torch.tensor(np.asarray(a), dtype=dt, device=dev)
    env.root_state_tensor.copy_(t(gi("root")))
    env.initial_root_states.view(n, 3, 13)[:, 1].copy_(t(gi("object_init")))
    env.initial_root_states.view(n, 3, 13)[:, 2].copy_(t(gi("goal_init")))
    env.dof_state.copy_(t(gi("dof_state")))
    env.prev_targets.copy_(t(gi("prev_targets"))); env.cur_targets.copy_(t(gi("cur_targets")))
    env.goal_states.copy_(t(gi("goal_states")))
    env.vec_sensor_tensor.copy_(t(gi("sensors"))); env.dof_force_tensor.copy_(t(gi("dof_force")))
    env.reset_buf.copy_(t(gi("reset"), torch.long)); env.reset_goal_buf.copy_(t(gi("reset_goal"), torch.long))
    env.progress_buf.copy_(t(gi("progress"), torch.long)); env.successes.copy_(t(gi("successes")))
    env._cons[0] = float(gi("cons")[0])
    env.reset_count.copy_(t(gi("reset_count"), torch.int32)); env.goal_reset_count.copy_(t(gi("goal_reset_count"), torch.int32))
    env.step(t(gi("actions")))
    torch.cuda.synchronize()
    c = lambda x: x.detach().cpu().numpy()
    np.testing.assert_allclose(c(env.root_state_tensor), go("root"), rtol=0, atol=1e-6)
    np.testing.assert_allclose(c(env.goal_states), go("goal_states"), rtol=0, atol=1e-6)
    np.testing.assert_allclose(c(env.cur_targets), go("cur_targets"), rtol=0, atol=1e-6)
    np.testing.assert_allclose(c(env.prev_targets), go("prev_targets"), rtol=0, atol=1e-6)
    np.testing.assert_allclose(c(env.dof_state), go("dof_state"), rtol=0, atol=1e-6)
    ref_obs = go("obs") if obs_type == "full_state" else go(f"obs_{obs_type}")
    obs = c(env.obs_buf).copy()
    # fingertip block: forward kinematics in fp32 (3e-5) and a quaternion's overall sign is free
    ft0 = {"full_state": 96, "full": 72}.get(obs_type)
    mask = np.ones(obs.shape[1], bool)
    if ft0 is not None:
        for f in range(5):
            sl = slice(ft0 + 13 * f + 3, ft0 + 13 * f + 7)
            sgn = np.sign((obs[:, sl] * ref_obs[:, sl]).sum(-1, keepdims=True))
            obs[:, sl] *= sgn
            mask[ft0 + 13 * f: ft0 + 13 * f + 13] = False
        err = np.abs(obs[:, ~mask] - ref_obs[:, ~mask]) / np.maximum(1.0, np.abs(ref_obs[:, ~mask]))
        assert err.max() < 1e-4, err.max()
    else:
        ft0 = {"full_no_vel": 42, "openai": 0}[obs_type]
        mask[ft0:ft0 + 15] = False
        assert np.abs(obs[:, ~mask] - ref_obs[:, ~mask]).max() < 5e-5
    np.testing.assert_allclose(obs[:, mask], ref_obs[:, mask], rtol=0, atol=2e-6)
    np.testing.assert_allclose(c(env.obs_clipped), np.clip(c(env.obs_buf), -5.0, 5.0), rtol=0, atol=0)
    if obs_type == "full_state":
        np.testing.assert_allclose(c(env.rew_buf), go("rew"), rtol=3e-6, atol=3e-5)
        assert np.array_equal(c(env.reset_buf), go("reset")) and np.array_equal(c(env.reset_goal_buf), go("reset_goal"))
        assert np.array_equal(c(env.progress_buf), go("progress")) and np.array_equal(c(env.successes), go("successes"))
        assert np.array_equal(c(env.timeout_buf), go("timeout"))
        np.testing.assert_allclose(c(env.consecutive_successes)[0], go("cons")[0], rtol=1e-6)
        assert float(env.extras["consecutive_successes"]) == pytest.approx(float(go("cons")[0]), rel=1e-6)
    env.sim.close()


def test_hand_env_runs_and_resets():
    """The full ShadowHand env (physics on): finite, cubes that fall get reset, lanes agree."""
    env = _make("ShadowHand", 512)
    torch.manual_seed(0)
    resets = 0
    for k in range(120):
        obs, rew, reset, extras = env.step(torch.rand(512, 20, device=env.device) * 2 - 1)
        resets += int(reset.sum())
    torch.cuda.synchronize()
    assert torch.isfinite(obs["obs"]).all() and torch.isfinite(rew).all() and torch.isfinite(env.root_state_tensor).all()
    assert obs["obs"].shape == (512, 211) and resets > 0
    z = env.root_state_tensor.view(512, 3, 13)[:, 1, 2]
    assert (z > 0.2).all()                       # nothing is left lying on the ground: fallen cubes were reset
    assert "consecutive_successes" in extras
    env.sim.close()


def test_generic_gym_api_path_matches_fused_hand_step():
    """The compatibility path for a three-actor env: driven the way the reference's shadow_hand.py drives `gym`
    (assets, tendon properties, fingertip sensors, three create_actor per env, position targets -> simulate), the shim's
    engine and the fused ShadowHand step produce the same hand / cube states, sensors and joint forces."""
    from isaacgymenvs_b200 import compat
    compat.install()
    from isaacgym import gymapi, gymtorch
    n = 64
    env = _make("ShadowHand", n)
    g = torch.Generator(device=env.device).manual_seed(5)
    env.step(2 * torch.rand(n, 20, device=env.device, generator=g) - 1)         # resets everything
    gym = gymapi.acquire_gym()
    sp = gymapi.SimParams(); sp.dt, sp.substeps, sp.up_axis, sp.gravity, sp.use_gpu_pipeline = 0.01667, 2, gymapi.UP_AXIS_Z, gymapi.Vec3(0, 0, -9.81), True
    sim = gym.create_sim(0, -1, gymapi.SIM_PHYSX, sp)
    gym.add_ground(sim, gymapi.PlaneParams())
    ao = gymapi.AssetOptions()
    ao.fix_base_link, ao.collapse_fixed_joints, ao.disable_gravity, ao.thickness, ao.angular_damping = True, True, True, 0.001, 0.01
    ao.default_dof_drive_mode = gymapi.DOF_MODE_NONE
    hand = gym.load_asset(sim, "/no/such/checkout/assets", "mjcf/open_ai_assets/hand/shadow_hand.xml", ao)
    assert gym.get_asset_dof_count(hand) == 24 and gym.get_asset_actuator_count(hand) == 20 and gym.get_asset_tendon_count(hand) >= 4
    tp = gym.get_asset_tendon_properties(hand)
    for i in range(gym.get_asset_tendon_count(hand)):
        if gym.get_asset_tendon_name(hand, i) in ("robot0:T_FFJ1c", "robot0:T_MFJ1c", "robot0:T_RFJ1c", "robot0:T_LFJ1c"):
            tp[i].limit_stiffness, tp[i].damping = 30, 0.1
    gym.set_asset_tendon_properties(hand, tp)
    for name in ["robot0:ffdistal", "robot0:mfdistal", "robot0:rfdistal", "robot0:lfdistal", "robot0:thdistal"]:
        gym.create_asset_force_sensor(hand, gym.find_asset_rigid_body_index(hand, name), gymapi.Transform())
    oo = gymapi.AssetOptions()
    cube = gym.load_asset(sim, "/no/such/checkout/assets", "urdf/objects/cube_multicolor.urdf", oo)
    oo.disable_gravity = True
    goal = gym.load_asset(sim, "/no/such/checkout/assets", "urdf/objects/cube_multicolor.urdf", oo)
    props = gym.get_asset_dof_properties(hand)
    hp = gymapi.Transform(); hp.p = gymapi.Vec3(0, 0, 0.5)
    op = gymapi.Transform(); op.p = gymapi.Vec3(0, -0.39, 0.6)
    gp = gymapi.Transform(); gp.p = gymapi.Vec3(-0.2, -0.45, 0.68)
    for i in range(n):
        e = gym.create_env(sim, gymapi.Vec3(-0.75, -0.75, 0), gymapi.Vec3(0.75, 0.75, 0.75), 8)
        h = gym.create_actor(e, hand, hp, "hand", i, -1, 0)
        gym.set_actor_dof_properties(e, h, props)
        assert gym.get_actor_index(e, h, gymapi.DOMAIN_SIM) == 3 * i
        o = gym.create_actor(e, cube, op, "object", i, 0, 0)
        k = gym.create_actor(e, goal, gp, "goal_object", i + n, 0, 0)
        assert gym.get_actor_index(e, o, gymapi.DOMAIN_SIM) == 3 * i + 1 and gym.get_actor_index(e, k, gymapi.DOMAIN_SIM) == 3 * i + 2
    gym.prepare_sim(sim)
    root = gymtorch.wrap_tensor(gym.acquire_actor_root_state_tensor(sim))
    dof = gymtorch.wrap_tensor(gym.acquire_dof_state_tensor(sim))
    sens = gymtorch.wrap_tensor(gym.acquire_force_sensor_tensor(sim)).view(n, 30)
    dfrc = gymtorch.wrap_tensor(gym.acquire_dof_force_tensor(sim)).view(n, 24)
    assert root.shape == (3 * n, 13) and abs(float(root[1, 2]) - 0.6) < 1e-6
    # the fixed base keeps the pose the asset file gives the hand; the shim's start pose has identity rotation like the reference's
    root.view(n, 3, 13)[:, 0].copy_(env.root_state_tensor.view(n, 3, 13)[:, 0])
    touched = 0.0
    for _ in range(30):
        pre_reset = env.reset_buf.clone()
        root.copy_(env.root_state_tensor); dof.copy_(env.dof_state)
        a = 2 * torch.rand(n, 20, device=env.device, generator=g) - 1
        env.step(a)
        gym.set_dof_position_target_tensor(sim, gymtorch.unwrap_tensor(env.cur_targets))
        gym.simulate(sim)
        torch.cuda.synchronize()
        keep = pre_reset == 0                      # envs reset inside this fused step started from another state
        r1, r2 = root.view(n, 3, 13)[keep], env.root_state_tensor.view(n, 3, 13)[keep]
        assert torch.allclose(r1[:, 1], r2[:, 1], atol=2e-5, rtol=1e-3)
        assert torch.allclose(dof.view(n, 24, 2)[keep], env.dof_state.view(n, 24, 2)[keep], atol=1e-4, rtol=2e-3)
        assert torch.allclose(sens[keep], env.vec_sensor_tensor[keep], atol=1e-2, rtol=2e-3)
        assert torch.allclose(dfrc[keep], env.dof_force_tensor[keep], atol=1e-3, rtol=2e-3)
        touched = max(touched, float(env.vec_sensor_tensor.abs().max()))
    assert touched > 0.05                          # fingertips did press on the cube at some point
    bs = gymtorch.wrap_tensor(gym.acquire_rigid_body_state_tensor(sim)).view(n, -1, 13)
    gym.refresh_rigid_body_state_tensor(sim)
    assert bs.shape[1] == env.model.nb + 2 and torch.equal(bs[:, -2], root.view(n, 3, 13)[:, 1])
    env.sim.close()


def test_hand_asymmetric_states_buffer():
    """asymmetric_observations: obs_buf in the openai layout, states_buf in the full_state layout (shadow_hand.py:457-458),
    both against the golden vectors of the same step."""
    gold = np.load(os.path.join(GOLD, "shadow_hand.npz"))
    gi = lambda k: gold[f"a_in_{k}"]
    n = gi("reset").shape[0]
    env = _hand_env(n, "a", "openai", asymmetric_observations=True)
    assert env.num_states == 211 and env.states_buf.shape == (n, 211) and env.num_obs == 42
    dev = env.device
    t = lambda a, dt=torch.float32: torch.tensor(np.asarray(a), dtype=dt, device=dev)
    env.root_state_tensor.copy_(t(gi("root")))
    env.initial_root_states.view(n, 3, 13)[:, 1].copy_(t(gi("object_init"))); env.initial_root_states.view(n, 3, 13)[:, 2].copy_(t(gi("goal_init")))
    env.dof_state.copy_(t(gi("dof_state"))); env.prev_targets.copy_(t(gi("prev_targets"))); env.cur_targets.copy_(t(gi("cur_targets")))
    env.goal_states.copy_(t(gi("goal_states"))); env.vec_sensor_tensor.copy_(t(gi("sensors"))); env.dof_force_tensor.copy_(t(gi("dof_force")))
    env.reset_buf.copy_(t(gi("reset"), torch.long)); env.reset_goal_buf.copy_(t(gi("reset_goal"), torch.long))
    env.progress_buf.copy_(t(gi("progress"), torch.long)); env.successes.copy_(t(gi("successes")))
    env.reset_count.copy_(t(gi("reset_count"), torch.int32)); env.goal_reset_count.copy_(t(gi("goal_reset_count"), torch.int32))
    obs, rew, reset, extras = env.step(t(gi("actions")))
    torch.cuda.synchronize()
    st = env.states_buf.cpu().numpy().copy(); ref = gold["a_out_obs"]
    mask = np.ones(211, bool)
    for f in range(5):
        sl = slice(96 + 13 * f + 3, 96 + 13 * f + 7)
        st[:, sl] *= np.sign((st[:, sl] * ref[:, sl]).sum(-1, keepdims=True))
        mask[96 + 13 * f: 96 + 13 * f + 13] = False
    np.testing.assert_allclose(st[:, mask], ref[:, mask], rtol=0, atol=2e-6)
    assert (np.abs(st[:, ~mask] - ref[:, ~mask]) / np.maximum(1.0, np.abs(ref[:, ~mask]))).max() < 1e-4
    ob = env.obs_buf.cpu().numpy(); ro = gold["a_out_obs_openai"]
    assert np.abs(ob[:, 15:] - ro[:, 15:]).max() < 2e-6 and np.abs(ob[:, :15] - ro[:, :15]).max() < 5e-5
    assert torch.equal(obs["states"], torch.clamp(env.states_buf, -5.0, 5.0))
    np.testing.assert_allclose(rew.cpu().numpy(), gold["a_out_rew"], rtol=3e-6, atol=3e-5)
    env.sim.close()


@pytest.mark.parametrize("task", ["Ant", "ShadowHand"])
def test_rlgames_style_driver_loop(task):
    """The calls rl_games' vec-env wrapper makes (utils/rlgames_utils.py:242-295: get_env_info, reset, step, reset_done,
    get_number_of_agents, set_train_info, get/set_env_state) and the shapes / dtypes / devices a PPO rollout buffer expects."""
    n = 64
    env = _make(task, n)
    info = {"action_space": env.action_space, "observation_space": env.observation_space}
    assert info["action_space"].shape == (env.num_acts,) and info["observation_space"].shape == (env.num_obs,)
    assert env.get_number_of_agents() == 1 and env.num_states == 0
    env.set_train_info(0); assert env.get_env_state() is None; env.set_env_state(None)
    obs = env.reset()
    assert set(obs) == {"obs"} and obs["obs"].shape == (n, env.num_obs) and obs["obs"].device.type == "cuda"
    ep_ret = torch.zeros(n, device=env.device)
    for k in range(40):
        a = 2 * torch.rand(n, env.num_acts, device=env.device) - 1
        obs, rew, dones, infos = env.step(a)
        assert obs["obs"].shape == (n, env.num_obs) and rew.shape == (n,) and dones.shape == (n,) and dones.dtype == torch.long
        assert infos["time_outs"].shape == (n,) and infos["time_outs"].dtype == torch.bool
        assert (obs["obs"].abs() <= env.clip_obs + 1e-6).all()
        ep_ret += rew
        if k == 20:
            env.reset_idx(torch.arange(0, n, 2, device=env.device))          # force a reset of half the envs
            od, ids = env.reset_done()
            assert ids.numel() >= n // 2 and od["obs"].shape == (n, env.num_obs)
    assert torch.isfinite(ep_ret).all()
    assert (env.progress_buf[0::2] <= 20).all()                               # the forced resets restarted those episodes
    env.sim.close()


def test_generic_gym_api_position_drive_anymal_stands():
    """PhysX-style position drives through the generic path (tasks/anymal.py:199-203,226-229): ANYmal with stiffness 85 /
    damping 2 holding its default joint angles on the plane settles on its feet; net contact forces carry its weight."""
    from isaacgymenvs_b200 import compat
    compat.install()
    from isaacgym import gymapi, gymtorch
    n = 32
    gym = gymapi.acquire_gym()
    sp = gymapi.SimParams(); sp.dt, sp.substeps, sp.up_axis, sp.gravity, sp.use_gpu_pipeline = 0.02, 2, gymapi.UP_AXIS_Z, gymapi.Vec3(0, 0, -9.81), True
    sim = gym.create_sim(0, -1, gymapi.SIM_PHYSX, sp)
    gym.add_ground(sim, gymapi.PlaneParams())
    ao = gymapi.AssetOptions()
    ao.default_dof_drive_mode, ao.collapse_fixed_joints, ao.replace_cylinder_with_capsule = gymapi.DOF_MODE_NONE, True, True
    ao.density, ao.angular_damping, ao.linear_damping, ao.armature, ao.thickness = 0.001, 0.0, 0.0, 0.0, 0.01
    asset = gym.load_asset(sim, "/no/such/checkout/assets", "urdf/anymal_c/urdf/anymal_minimal.urdf", ao)
    assert gym.get_asset_dof_count(asset) == 12
    names = gym.get_asset_dof_names(asset)
    default = torch.tensor([{"HAA": 0.03 if nm.startswith("L") else -0.03, "HFE": 0.4 if nm[1] == "F" else -0.4,
                             "KFE": -0.8 if nm[1] == "F" else 0.8}[nm[3:6]] for nm in names], device="cuda:0")
    props = gym.get_asset_dof_properties(asset)
    props["driveMode"][:] = gymapi.DOF_MODE_POS; props["stiffness"][:] = 85.0; props["damping"][:] = 2.0
    pose = gymapi.Transform(); pose.p = gymapi.Vec3(0, 0, 0.62)
    for i in range(n):
        e = gym.create_env(sim, gymapi.Vec3(-2, -2, 0), gymapi.Vec3(2, 2, 2), 8)
        h = gym.create_actor(e, asset, pose, "anymal", i, 1, 0)
        gym.set_actor_dof_properties(e, h, props)
    gym.prepare_sim(sim)
    root = gymtorch.wrap_tensor(gym.acquire_actor_root_state_tensor(sim))
    dof = gymtorch.wrap_tensor(gym.acquire_dof_state_tensor(sim)).view(n, 12, 2)
    cf = gymtorch.wrap_tensor(gym.acquire_net_contact_force_tensor(sim)).view(n, -1, 3)
    dof[:, :, 0] = default
    gym.set_dof_state_tensor(sim, gymtorch.unwrap_tensor(dof))
    targets = default.repeat(n, 1).contiguous()
    for _ in range(150):
        gym.set_dof_position_target_tensor(sim, gymtorch.unwrap_tensor(targets))
        gym.simulate(sim)
    torch.cuda.synchronize()
    assert torch.isfinite(root).all() and torch.isfinite(dof).all()
    assert ((root[:, 2] > 0.4) & (root[:, 2] < 0.65)).all(), root[:, 2]
    # a PD drive has steady-state error under load: ~20 N m at the knees / 85 N m per rad = 0.24 rad
    assert (dof[:, :, 0] - default).abs().max() < 0.3 and dof[:, :, 1].abs().max() < 0.2 and root[:, 7:13].abs().max() < 0.1
    up_z = 1 - 2 * (root[:, 3] ** 2 + root[:, 4] ** 2)
    assert (up_z > 0.98).all()
    W = float(sim.asset.model.total_mass()) * 9.81
    Fz = cf[:, :, 2].sum(1)
    assert ((Fz - W).abs() / W < 0.05).all(), (Fz, W)
