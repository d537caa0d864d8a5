"""developer probe: a few Humanoid steps with link-link contact on (for ncu)"""
import sys, torch
sys.path.insert(0, '.')
import isaacgymenvs_b200
from isaacgymenvs_b200 import config
cfg = config.builtin_cfg("Humanoid", {"sim_device": "cuda:0", "rl_device": "cuda:0"}); cfg["task"]["env"]["selfCollision"] = True
env = isaacgymenvs_b200.make(seed=1, task="Humanoid", num_envs=8192, sim_device="cuda:0", rl_device="cuda:0", headless=True, cfg=cfg)
g = torch.Generator(device="cuda:0"); g.manual_seed(1)
for k in range(40):
    env.step(torch.rand((8192, env.num_acts), device="cuda:0", generator=g) * 2 - 1)
torch.cuda.synchronize()
