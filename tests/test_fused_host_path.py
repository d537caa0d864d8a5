"""The Python side of the FUSED tasks on CPU: `engine.Sim` is replaced by a stand-in with the same tensors whose
task_step() writes a recognisable pattern, so constructor logic, buffer binding, VecTask.step()'s host code (dict
observations, time-outs, extras, domain-randomisation noise) run without a GPU.  The kernels themselves are covered by
tests/test_gpu_parity.py."""
import numpy as np
import pytest
import torch


class _FakeFusedSim:
    def __init__(self, model, num_envs, dt, substeps, gravity=(0, 0, -9.81), ground_mu=1.0, device="cpu", ext=None, **kw):
        from isaacgymenvs_b200 import engine as E
        self.E, self.model, self.num_envs, self.ext, self.kw = E, model, num_envs, ext, kw
        self.device = torch.device("cpu")
        self.actors_per_env = int(ext.actors_per_env) if ext is not None else 1
        self.nd, self.nb, self.ns = model.ndof, model.nb, len(model.sensor_body)
        self.root_state = torch.zeros(num_envs * self.actors_per_env, 13); self.root_state[:, 6] = 1
        self.dof_state = torch.zeros(num_envs * max(self.nd, 1), 2)
        self.dof_actuation = torch.zeros(num_envs, max(self.nd, 1)); self.dof_target = torch.zeros_like(self.dof_actuation)
        self.tensors, self.steps, self.params, self.bound, self.last_actions = {}, 0, None, {}, None

    def acquire(self, slot):
        E, N = self.E, self.num_envs
        shape = {E.T_RIGID_BODY_STATE: (N * (self.nb + self.actors_per_env - 1), 13), E.T_FORCE_SENSOR: (N * max(self.ns, 1), 6),
                 E.T_DOF_FORCE: (N * max(self.nd, 1),), E.T_NET_CONTACT: (N * self.nb, 3)}[slot]
        return self.tensors.setdefault(slot, torch.zeros(*shape))

    def set_task(self, params, buffers):
        self.params, self.bound = params, dict(buffers)

    def _bind(self, slot, t):
        self.bound[slot] = t
        return t

    def task_step(self, actions):
        E = self.E
        self.steps += 1
        self.last_actions = actions.clone()
        self.bound[E.T_OBS].fill_(float(self.steps))
        if E.T_OBS_CLIPPED in self.bound:
            self.bound[E.T_OBS_CLIPPED].copy_(self.bound[E.T_OBS].clamp(-float(self.params.clip_obs), float(self.params.clip_obs)))
        self.bound[E.T_REW].fill_(0.5); self.bound[E.T_PROGRESS] += 1; self.bound[E.T_RESET].zero_()

    def launch_count(self):
        return self.steps

    def close(self):
        pass


@pytest.fixture
def fused_cpu(monkeypatch):
    from isaacgymenvs_b200 import engine
    from isaacgymenvs_b200.tasks.base import vec_task
    monkeypatch.setattr(engine, "Sim", _FakeFusedSim)
    # the stand-in lives on the CPU: let the host code believe the pipeline device is usable
    orig = vec_task.Env.__init__

    class _Dev(str):                  # "cpu" for torch, but not equal to "cpu" for the no-CPU-path guard of VecTask.__init__
        def __eq__(self, other):
            return False
        __hash__ = str.__hash__

    def init(self, config, rl_device, sim_device, graphics_device_id, headless):
        orig(self, config, rl_device, sim_device, graphics_device_id, headless)
        self.device = _Dev("cpu")
    monkeypatch.setattr(vec_task.Env, "__init__", init)
    yield


def _make(task, n, env=None, task_section=None):
    from isaacgymenvs_b200 import config
    from isaacgymenvs_b200.tasks import isaacgym_task_map
    cfg = config.builtin_cfg(task, {"sim_device": "cuda:0", "rl_device": "cpu"})
    cfg["task"]["env"].update(env or {})
    cfg["task"]["task"].update(task_section or {})
    cfg["task"]["env"]["numEnvs"] = n
    t = cfg["task"]; t["seed"] = 42
    return isaacgym_task_map[task](cfg=t, rl_device="cpu", sim_device="cuda:0", graphics_device_id=-1, headless=True)


@pytest.mark.parametrize("task,nobs,nact", [("Cartpole", 4, 1), ("Ant", 60, 8), ("Humanoid", 108, 21), ("ShadowHand", 211, 20),
                                            ("AnymalTerrain", 188, 12)])
def test_fused_task_host_side(fused_cpu, task, nobs, nact):
    n = 16
    env = _make(task, n)
    E = env.sim.E
    assert env.num_obs == nobs and env.num_acts == nact and env.obs_buf.shape == (n, nobs)
    p = env.sim.params
    assert p.num_obs == nobs and p.num_actions == nact and p.control_freq_inv == env.control_freq_inv and p.seed == 42
    b = env.sim.bound
    assert b[E.T_OBS].data_ptr() == env.obs_buf.data_ptr() and b[E.T_RESET].data_ptr() == env.reset_buf.data_ptr()
    assert (env.reset_buf == 1).all()                                  # vec_task.py:309: the first step resets every env
    obs, rew, reset, extras = env.step(torch.rand(n, nact) * 2 - 1)
    assert obs["obs"].shape == (n, nobs) and (obs["obs"] == 1.0).all() and (rew == 0.5).all() and "time_outs" in extras
    assert env.sim.steps == 1 and env.control_steps == 1
    if task == "ShadowHand":
        assert env.root_state_tensor.shape == (3 * n, 13) and b[E.T_GOAL_STATES].shape == (n, 13) and "consecutive_successes" in extras
        assert env.sim.ext.actors_per_env == 3 and env.sim.ext.nten == 4 and list(p.fingertip_body) == list(env.fingertip_handles_np)


def test_fused_task_observation_and_action_noise(fused_cpu):
    """task.randomize with observation / action entries: actions are perturbed before the step, the observation after
    it, the engine-bound tensor stays the one the engine writes; physical entries are refused."""
    n = 8
    dr = {"frequency": 1, "observations": {"range": [0.0, 0.5], "operation": "additive", "distribution": "gaussian"},
          "actions": {"range": [0.0, 0.1], "operation": "additive", "distribution": "gaussian"}}
    env = _make("Ant", n, task_section={"randomize": True, "randomization_params": dr})
    E = env.sim.E
    bound = env.sim.bound[E.T_OBS]
    a = torch.zeros(n, 8)
    torch.manual_seed(0)
    obs, *_ = env.step(a)
    assert env.sim.last_actions.abs().max() > 0 and env.sim.last_actions.abs().max() < 1.0        # noisy actions reached the engine
    assert (bound == 1.0).all() and (obs["obs"] - 1.0).abs().max() > 1e-3                          # noisy observation, clean engine tensor
    obs2, *_ = env.step(a)
    assert (bound == 2.0).all() and env.sim.bound[E.T_OBS].data_ptr() == bound.data_ptr() and abs(float(obs2["obs"].mean()) - 2.0) < 0.3
    with pytest.raises(NotImplementedError):
        _make("Ant", n, task_section={"randomize": True, "randomization_params": {"frequency": 1, "sim_params": {"gravity": {}}}})


def test_fused_task_physical_domain_randomisation(fused_cpu):
    """task.randomize with the reference's Ant.yaml actor_params block (cfg/task/Ant.yaml:76-101): per-env parameter tensors
    are sampled inside the configured ranges, bound to the engine, re-sampled only for envs that are flagged for reset AND
    whose randomisation counter passed `frequency`; setup_only properties are drawn once."""
    n = 64
    dr = {"frequency": 3,
          "actor_params": {"ant": {"color": True,
                                   "rigid_body_properties": {"mass": {"range": [0.5, 1.5], "operation": "scaling", "distribution": "uniform", "setup_only": True}},
                                   "dof_properties": {"damping": {"range": [0.5, 1.5], "operation": "scaling", "distribution": "uniform"},
                                                      "stiffness": {"range": [0.5, 1.5], "operation": "scaling", "distribution": "uniform"},
                                                      "lower": {"range": [0, 0.01], "operation": "additive", "distribution": "gaussian"},
                                                      "upper": {"range": [0, 0.01], "operation": "additive", "distribution": "gaussian"}}}}}
    torch.manual_seed(0)
    env = _make("Ant", n, task_section={"randomize": True, "randomization_params": dr})
    E = env.sim.E
    pr = env.physical_randomizer
    ms, dp = env.sim.bound[E.T_ENV_MASS_SCALE], env.sim.bound[E.T_ENV_DOF_PROPS]
    assert ms.shape == (n, 9) and dp.shape == (n, 8, 4) and E.T_ENV_FRICTION not in env.sim.bound
    assert (ms >= 0.5).all() and (ms <= 1.5).all() and ms.std() > 0.1
    og = pr.og_dof
    assert ((dp[..., 0] >= 0.5 * og[:, 0] - 1e-6) & (dp[..., 0] <= 1.5 * og[:, 0] + 1e-6)).all()        # damping scaled
    assert (dp[..., 2] - og[:, 2]).abs().max() < 0.06 and (dp[..., 2] - og[:, 2]).abs().max() > 1e-3     # limits shifted by N(0, 0.01)
    ms0, dp0 = ms.clone(), dp.clone()
    a = torch.zeros(n, 8)
    env.step(a)                                                        # counters 0 -> below the frequency: nothing changes
    assert torch.equal(ms, ms0) and torch.equal(dp, dp0)
    for _ in range(3):
        env.step(a)
    env.reset_buf[: n // 2] = 1                                        # half of the envs are about to reset, their counters are past 3
    env.step(a)
    changed = (dp != dp0).any(-1).any(-1)
    assert changed[: n // 2].all() and not changed[n // 2:].any()
    assert torch.equal(ms, ms0)                                        # setup_only
    assert (env.randomize_buf[: n // 2] <= 1).all() and (env.randomize_buf[n // 2:] >= 4).all()
    with pytest.raises(NotImplementedError):
        _make("Ant", n, task_section={"randomize": True, "randomization_params": {"frequency": 1, "actor_params": {"ant": {"tendon_properties": {}}}}})
