import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import isaacgymenvs_b200
from isaacgymenvs_b200 import config
cfg = config.builtin_cfg("AnymalTerrain", {"sim_device": "cuda:0", "rl_device": "cuda:0"})
mode = sys.argv[1] if len(sys.argv) > 1 else "plane"
if mode == "plane":
    cfg["task"]["env"]["terrain"]["terrainType"] = "plane"
env = isaacgymenvs_b200.make(seed=42, task="AnymalTerrain", num_envs=128, sim_device="cuda:0", rl_device="cuda:0", headless=True, cfg=cfg)
for k in range(3):
    o, r, d, e = env.step(torch.zeros(128, 12, device="cuda:0"))
torch.cuda.synchronize()
print("ok", mode, float(env.root_states[:, 2].mean()), int(d.sum()))
