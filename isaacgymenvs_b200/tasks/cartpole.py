"""Cartpole with the reference class's surface (`isaacgymenvs/tasks/cartpole.py`)."""
import copy
import torch

from .. import engine
from ..assets import load_asset_file
from ..importer.model import BuildOptions
from .base.vec_task import VecTask
from .locomotion import _asset_root


class Cartpole(VecTask):
    def __init__(self, cfg, rl_device, sim_device, graphics_device_id, headless, virtual_screen_capture=False,
                 force_render=False):
        self.cfg = cfg
        self.reset_dist = cfg["env"]["resetDist"]
        self.max_push_effort = cfg["env"]["maxEffort"]
        self.max_episode_length = 500                    # cartpole.py:44
        cfg["env"]["numObservations"] = 4
        cfg["env"]["numActions"] = 1
        self.up_axis_idx = 2
        super().__init__(config=cfg, rl_device=rl_device, sim_device=sim_device, graphics_device_id=graphics_device_id,
                         headless=headless, virtual_screen_capture=virtual_screen_capture, force_render=force_render)

    def _build_model(self):
        asset_file = self.cfg["env"].get("asset", {}).get("assetFileName", "urdf/cartpole.urdf")
        model = copy.deepcopy(load_asset_file(_asset_root(), asset_file, BuildOptions(fix_base_link=True, angular_damping=0.5)))  # cartpole.py:86-88: AssetOptions defaults otherwise (angular_damping 0.5)
        self.num_dof = model.ndof
        return model

    def create_sim(self):
        sim = super().create_sim()
        sim.root_state[:, 2] = 2.0                       # cartpole.py:91-94 (z-up pose)
        self.dof_state = sim.dof_state
        self.dof_pos = self.dof_state.view(self.num_envs, self.num_dof, 2)[..., 0]
        self.dof_vel = self.dof_state.view(self.num_envs, self.num_dof, 2)[..., 1]
        return sim

    def _task_params(self):
        p = engine.CTaskParams()
        p.task = engine.TASK_CARTPOLE
        p.max_episode_length = float(self.max_episode_length)
        p.max_push_effort = float(self.max_push_effort)
        p.reset_dist = float(self.reset_dist)
        p.dt = float(self.dt)
        return p
