"""developer probe: where does rollout(K) differ from K x step()?"""
import sys, torch
sys.path.insert(0, '.')
import isaacgymenvs_b200
from isaacgymenvs_b200 import config
def make(n, ep):
    cfg = config.builtin_cfg("Ant", {"sim_device": "cuda:0", "rl_device": "cuda:0"}); cfg["task"]["env"]["episodeLength"] = ep
    return isaacgymenvs_b200.make(seed=42, task="Ant", num_envs=n, sim_device="cuda:0", rl_device="cuda:0", headless=True, cfg=cfg)
n, K = 16384, 6
a_env, b_env = make(n, 1000), make(n, 1000)
g = torch.Generator(device="cuda:0"); g.manual_seed(7)
acts = (torch.rand((K, n, a_env.num_acts), device="cuda:0", generator=g) * 2 - 1) * 1.2
ro, rd = [], []
for k in range(K):
    od, r, d, info = a_env.step(acts[k]); ro.append(od["obs"].clone()); rd.append(d.clone())
obs, rew, done, tout = b_env.rollout(acts)
torch.cuda.synchronize()
ro = torch.stack(ro); rd = torch.stack(rd)
bad = ((obs - ro).abs() >= 2e-4).nonzero()
envs = sorted(set(bad[:, 1].tolist()))
print("envs", envs)
e = envs[0]
print("done (steps) single:", rd[:, e].tolist(), "rollout:", done[:, e].tolist())
for k in range(K):
    print(k, "single", [round(x, 4) for x in ro[k, e, :12].tolist()]); print(k, "rollo ", [round(x, 4) for x in obs[k, e, :12].tolist()])
print("idx differing at first bad step:", bad[bad[:, 1] == e][:, [0, 2]].tolist()[:40])
print("final root single", a_env.root_states[e].tolist()); print("final root rollo ", b_env.root_states[e].tolist())
