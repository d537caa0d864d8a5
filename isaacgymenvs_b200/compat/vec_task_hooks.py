"""Hook-style `VecTask` for UNMODIFIED reference task files (the compatibility path).

Same constructor, buffers and step()/reset() contracts as the reference's
`isaacgymenvs/tasks/base/vec_task.py:207-455`, but here the task's own Python hooks run
(`create_sim`, `pre_physics_step`, `post_physics_step`, `reset_idx`) against the `gym` object of
`isaacgymenvs_b200.compat.gymapi`; only `gym.simulate` is the CUDA kernel.  The fused tasks in
`isaacgymenvs_b200.tasks` do not use this class.
"""
from typing import Any, Dict, Tuple

import numpy as np
import torch

from ..tasks.base.vec_task import Env, spaces   # noqa: F401  (device parsing, spaces, clip ranges: vec_task.py:67-204)
from . import gymapi

EXISTING_SIM = None


def _create_sim_once(gym, *args, **kwargs):
    """One sim per process (vec_task.py:58-64)."""
    global EXISTING_SIM
    if EXISTING_SIM is None:
        EXISTING_SIM = gym.create_sim(*args, **kwargs)
    return EXISTING_SIM


def reset_sim_singleton():
    global EXISTING_SIM
    EXISTING_SIM = None


class VecTask(Env):
    metadata = {"render.modes": ["human", "rgb_array"], "video.frames_per_second": 24}

    def __init__(self, config, rl_device, sim_device, graphics_device_id, headless, virtual_screen_capture: bool = False,
                 force_render: bool = False):
        self.cfg = config
        super().__init__(config, rl_device, sim_device, graphics_device_id, headless)
        self.virtual_screen_capture, self.force_render = virtual_screen_capture, force_render
        self.sim_params = self._parse_sim_params(self.cfg["physics_engine"], self.cfg["sim"])
        if self.cfg["physics_engine"] == "physx":
            self.physics_engine = gymapi.SIM_PHYSX
        elif self.cfg["physics_engine"] == "flex":
            self.physics_engine = gymapi.SIM_FLEX
        else:
            raise ValueError(f"Invalid physics engine backend: {self.cfg['physics_engine']}")
        self.dt: float = self.sim_params.dt
        self.gym = gymapi.acquire_gym()
        self.first_randomization = True
        self.original_props, self.dr_randomizations = {}, {}
        self.actor_params_generator = None
        self.extern_actor_params = {i: None for i in range(self.num_envs)}
        self.last_step = self.last_rand_step = -1
        self.sim_initialized = False
        self.create_sim()                      # the task's own create_sim (vec_task.py:259-263)
        self.gym.prepare_sim(self.sim)
        self.sim_initialized = True
        self.viewer, self.enable_viewer_sync = None, True
        self.allocate_buffers()
        self.obs_dict = {}

    def _parse_sim_params(self, physics_engine, cfg):
        """cfg['sim'] -> gymapi.SimParams (vec_task.py:514-562)."""
        p = gymapi.SimParams()
        if cfg["up_axis"] not in ("z", "y"):
            raise ValueError(f"Invalid physics up-axis: {cfg['up_axis']}")
        p.dt = cfg["dt"]
        p.num_client_threads = cfg.get("num_client_threads", 0)
        p.use_gpu_pipeline = cfg["use_gpu_pipeline"]
        p.substeps = cfg.get("substeps", 2)
        p.up_axis = gymapi.UP_AXIS_Z if cfg["up_axis"] == "z" else gymapi.UP_AXIS_Y
        p.gravity = gymapi.Vec3(*cfg["gravity"])
        for k, v in cfg.get(physics_engine, {}).items():
            setattr(getattr(p, physics_engine), k, gymapi.ContactCollection(v) if k == "contact_collection" else v)
        return p

    def create_sim(self, compute_device: int, graphics_device: int, physics_engine, sim_params):
        sim = _create_sim_once(self.gym, compute_device, graphics_device, physics_engine, sim_params)
        if sim is None:
            raise RuntimeError("*** Failed to create sim")
        return sim

    def set_viewer(self):
        self.viewer, self.enable_viewer_sync = None, True

    def allocate_buffers(self):
        dev, n = self.device, self.num_envs
        self.obs_buf = torch.zeros((n, self.num_obs), device=dev, dtype=torch.float)
        self.states_buf = torch.zeros((n, self.num_states), device=dev, dtype=torch.float)
        self.rew_buf = torch.zeros(n, device=dev, dtype=torch.float)
        self.reset_buf = torch.ones(n, device=dev, dtype=torch.long)
        self.timeout_buf = torch.zeros(n, device=dev, dtype=torch.long)
        self.progress_buf = torch.zeros(n, device=dev, dtype=torch.long)
        self.randomize_buf = torch.zeros(n, device=dev, dtype=torch.long)
        self.extras = {}

    def get_state(self):
        return torch.clamp(self.states_buf, -self.clip_obs, self.clip_obs).to(self.rl_device)

    def pre_physics_step(self, actions):
        raise NotImplementedError

    def post_physics_step(self):
        raise NotImplementedError

    def step(self, actions: torch.Tensor) -> Tuple[Dict[str, torch.Tensor], torch.Tensor, torch.Tensor, Dict[str, Any]]:
        """vec_task.py:360-408, hook by hook."""
        noise = self.dr_randomizations.get("actions")
        if noise:
            actions = noise["noise_lambda"](actions)
        self.pre_physics_step(torch.clamp(actions, -self.clip_actions, self.clip_actions))
        for _ in range(self.control_freq_inv):
            self.gym.simulate(self.sim)
        if self.device == "cpu":
            self.gym.fetch_results(self.sim, True)
        self.post_physics_step()
        self.control_steps += 1
        self.timeout_buf = (self.progress_buf >= self.max_episode_length - 1) & (self.reset_buf != 0)
        noise = self.dr_randomizations.get("observations")
        if noise:
            self.obs_buf = noise["noise_lambda"](self.obs_buf)
        self.extras["time_outs"] = self.timeout_buf.to(self.rl_device)
        self.obs_dict["obs"] = torch.clamp(self.obs_buf, -self.clip_obs, self.clip_obs).to(self.rl_device)
        if self.num_states > 0:
            self.obs_dict["states"] = self.get_state()
        return self.obs_dict, self.rew_buf.to(self.rl_device), self.reset_buf.to(self.rl_device), self.extras

    def zero_actions(self) -> torch.Tensor:
        return torch.zeros([self.num_envs, self.num_actions], dtype=torch.float32, device=self.rl_device)

    def reset_idx(self, env_idx):
        pass

    def reset(self):
        self.obs_dict["obs"] = torch.clamp(self.obs_buf, -self.clip_obs, self.clip_obs).to(self.rl_device)
        if self.num_states > 0:
            self.obs_dict["states"] = self.get_state()
        return self.obs_dict

    def reset_done(self):
        done = self.reset_buf.nonzero(as_tuple=False).flatten()
        if len(done) > 0:
            self.reset_idx(done)
        return self.reset(), done

    def render(self, mode="rgb_array"):
        return None

    def apply_randomizations(self, dr_params):
        raise NotImplementedError("domain randomisation is outside the hot path (SURVEY.md 8f rank 3)")
