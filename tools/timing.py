"""Developer timing probe: kernel time vs control_freq_inv (0 = prologue+epilogue only)."""
import sys, os, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import isaacgymenvs_b200
from isaacgymenvs_b200 import config

def run(task, n, cfi, steps=100):
    cfg = config.builtin_cfg(task, {"sim_device": "cuda:0", "rl_device": "cuda:0"})
    cfg["task"]["env"]["controlFrequencyInv"] = cfi
    env = isaacgymenvs_b200.make(seed=42, task=task, num_envs=n, sim_device="cuda:0", rl_device="cuda:0", headless=True, cfg=cfg)
    a = [2 * torch.rand((n, env.num_actions), device="cuda:0") - 1 for _ in range(8)]
    for k in range(10):
        env.sim.task_step(a[k % 8])
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for k in range(steps):
        env.sim.task_step(a[k % 8])
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / steps * 1e3

if __name__ == "__main__":
    task = sys.argv[1] if len(sys.argv) > 1 else "Ant"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    out = {cfi: round(run(task, n, cfi), 2) for cfi in ((1, 2) if task == "AnymalTerrain" else (0, 1, 2, 4))}
    print(task, n, os.environ.get("B2G_LIB", "default"), os.environ.get("B2G_SINGLE_LANE", ""), "us per step by control_freq_inv:", out)
