"""Long ShadowHand rollout under random actions: finiteness, speeds, reset statistics (run on a GPU box)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import isaacgymenvs_b200
from isaacgymenvs_b200 import config
n = 4096
cfg = config.builtin_cfg("ShadowHand", {"sim_device": "cuda:0", "rl_device": "cuda:0"})
env = isaacgymenvs_b200.make(seed=1, task="ShadowHand", num_envs=n, sim_device="cuda:0", rl_device="cuda:0", headless=True, cfg=cfg)
g = torch.Generator(device="cuda:0").manual_seed(0)
resets = goal = 0; vmax = wmax = qdmax = 0.0; rsum = 0.0
a = torch.zeros(n, 20, device="cuda:0")
for k in range(3000):
    a = 0.9 * a + 0.1 * (2 * torch.rand(n, 20, device="cuda:0", generator=g) - 1) * 3      # smooth random actions
    obs, rew, reset, ex = env.step(a.clamp(-1, 1))
    resets += int(reset.sum()); goal += int(env.reset_goal_buf.sum()); rsum += float(rew.mean())
    if k % 100 == 0:
        o = env.root_state_tensor.view(n, 3, 13)[:, 1]
        vmax = max(vmax, float(o[:, 7:10].norm(dim=-1).max())); wmax = max(wmax, float(o[:, 10:13].norm(dim=-1).max()))
        qdmax = max(qdmax, float(env.shadow_hand_dof_vel.abs().max()))
        assert torch.isfinite(obs["obs"]).all() and torch.isfinite(env.root_state_tensor).all() and torch.isfinite(env.dof_state).all(), k
print("steps 3000 envs", n, "resets", resets, "goal hits", goal, "mean episode length ~", 3000 * n / max(resets, 1),
      "max cube speed", round(vmax, 2), "max cube spin", round(wmax, 1), "max joint speed", round(qdmax, 1), "mean reward", round(rsum / 3000, 3),
      "consecutive_successes", float(env.consecutive_successes))
o = env.root_state_tensor.view(n, 3, 13)[:, 1]
print("cube z quantiles", torch.quantile(o[:, 2], torch.tensor([0.0, 0.1, 0.5, 0.9, 1.0], device="cuda:0")).tolist())
