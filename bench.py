#!/usr/bin/env python
"""bench.py -- env-steps/s of the fused VecTask.step() hot path (BASELINE.json metric).

    python bench.py --gpus 1 --steps 200 --warmup 5            # Ant, num_envs=16384 per GPU
    torchrun --nproc-per-node N ... bench.py --gpus N ...       # weak scaling: 16384 envs per rank
    python bench.py --impl reference ...                        # CPU port of the path (oracle/), host cores

One "step" = one VecTask.step() over all envs under random actions U(-1,1) (the README rollout
loop of the reference, README.md:39-51).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {   # name -> (task, num_envs per GPU, algorithmic bytes per env-step: SURVEY.md 8d / DESIGN.md)
    "ant": ("Ant", 16384, 673),
    "humanoid": ("Humanoid", 8192, 1161),
    "cartpole": ("Cartpole", 16384, 89),
    "anymal": ("AnymalTerrain", 4096, 2250),
    "shadow_hand": ("ShadowHand", 4096, 3640),    # BASELINE.json config 5: 32768 envs over 8 GPUs
}
METRIC = "env-steps/s at num_envs=16384 (Ant), 1/2/4/8 B200; %HBM roofline"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_env(task, n, device, rank):
    import isaacgymenvs_b200
    from isaacgymenvs_b200 import config
    cfg = config.builtin_cfg(task, {"sim_device": device, "rl_device": device})
    cfg["task"]["env_id_offset"] = rank * n
    return isaacgymenvs_b200.make(seed=42, task=task, num_envs=n, sim_device=device, rl_device=device,
                                  headless=True, cfg=cfg)


# ------------------------------------------------------------------------------------ CPU legs
def cpu_pipeline(task, n_envs, steps, threads):
    """The same control step on host cores: oracle physics (C, float32, pthreads over envs) + the
    numpy restatement of the reference's obs/reward functions.  Returns env-steps/s."""
    import copy
    from isaacgymenvs_b200.assets import load_compiled
    from oracle.oracle import OracleSim
    from oracle import tasks_np as T
    assert task == "Ant"
    m = copy.deepcopy(load_compiled("ant"))
    m.sensor_body = np.array([2, 4, 6, 8], dtype=np.int32)
    m.sensor_pos = np.zeros((4, 3)); m.sensor_quat = np.tile([0, 0, 0, 1.0], (4, 1))
    sim = OracleSim(m, 0.0166, 2, precision="f32", threads=threads)
    f32 = np.float32
    rng = np.random.default_rng(42)
    lo = np.minimum(m.lower[1:], m.upper[1:]).astype(f32); hi = np.maximum(m.lower[1:], m.upper[1:]).astype(f32)
    init = np.where(lo > 0, lo, np.where(hi < 0, hi, 0)).astype(f32)
    root = np.zeros((n_envs, 13), f32); root[:, 2] = 0.44; root[:, 6] = 1
    dof = np.zeros((n_envs, 8, 2), f32); dof[..., 0] = init
    pot = np.full(n_envs, -1000.0 / 0.0166, f32)
    targets = np.tile(f32([1000, 0, 0]), (n_envs, 1)); isr = np.tile(f32([0, 0, 0, 1]), (n_envs, 1))
    b0 = np.tile(f32([1, 0, 0]), (n_envs, 1)); b1 = np.tile(f32([0, 0, 1]), (n_envs, 1))
    progress = np.zeros(n_envs, np.int64); reset = np.zeros(n_envs, np.int64)

    def one():
        nonlocal pot, progress, reset
        a = np.clip(rng.uniform(-1, 1, size=(n_envs, 8)).astype(f32), -1, 1)
        out = sim.simulate(root, dof, a * f32(15.0))
        progress += 1
        ids = np.nonzero(reset)[0]
        if len(ids):
            dof[ids, :, 0] = np.clip(init + rng.uniform(-0.2, 0.2, size=(len(ids), 8)).astype(f32), lo, hi)
            dof[ids, :, 1] = rng.uniform(-0.1, 0.1, size=(len(ids), 8)).astype(f32)
            root[ids] = 0; root[ids, 2] = 0.44; root[ids, 6] = 1
            pot[ids] = T.potentials_from(targets[ids] - root[ids, :3], 0.0166)
            progress[ids] = 0
        obs, pot2, prev, _, _ = T.ant_observations(root, targets, pot, isr, dof[..., 0], dof[..., 1], lo, hi, 0.2,
                                                   out["sensor"].reshape(n_envs, 24), a, 0.0166, 0.1, b0, b1)
        rew, reset = T.ant_reward(obs, np.zeros(n_envs, np.int64), progress, a, 0.1, 0.5, pot2, prev, 0.005, 0.05,
                                  0.1, 0.31, -2.0, 1000.0)
        pot = pot2
    one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return n_envs * steps / dt, dt


def cpu_pipeline_hand(n_envs, steps, threads):
    """ShadowHand control step on host cores: oracle physics of hand + cube, numpy restatement of the task."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from tests.hand_common import hand_setup, DT, SUBSTEPS, G
    from oracle.oracle import OracleSim
    from oracle import tasks_np as T
    f32 = np.float32
    m, obj, tendons = hand_setup()
    sim = OracleSim(m, DT, SUBSTEPS, G, precision="f32", obj=obj, tendons=tendons, tendon_k=30.0, tendon_d=0.1, threads=threads)
    D = m.ndof
    names = list(m.dof_names)
    rng = np.random.default_rng(42)
    root = np.zeros((n_envs, 3, 13), f32); root[:, :, 6] = 1
    root[:, 0, 0:3] = [0, 0, 0.5]; root[:, 0, 3:7] = m.default_root_quat
    obj_init = np.zeros((n_envs, 13), f32); obj_init[:, 0:3] = [0, -0.39, 0.6]; obj_init[:, 6] = 1
    goal_init = obj_init.copy(); goal_init[:, 2] -= 0.04
    st = dict(root=root, dof_pos=np.zeros((n_envs, D), f32), dof_vel=np.zeros((n_envs, D), f32), cur_targets=np.zeros((n_envs, D), f32),
              prev_targets=np.zeros((n_envs, D), f32), goal_states=goal_init.copy(), reset=np.ones(n_envs, np.int64),
              reset_goal=np.ones(n_envs, np.int64), progress=np.zeros(n_envs, np.int64), successes=np.zeros(n_envs, f32),
              reset_count=np.zeros(n_envs, np.int32), goal_reset_count=np.zeros(n_envs, np.int32))
    P = dict(seed=42, goal_init=goal_init, object_init=obj_init, goal_displacement=f32([-0.2, -0.06, 0.12]), reset_position_noise=0.01,
             reset_dof_pos_noise=0.2, reset_dof_vel_noise=0.0, lower=m.lower[1:].astype(f32), upper=m.upper[1:].astype(f32),
             default_pos=np.zeros(D, f32), default_vel=np.zeros(D, f32), clip_actions=1.0,
             actuated=np.array([names.index(j) for j in m.actuator_joint]), use_relative_control=False, dof_speed_scale=20.0, dt=DT,
             act_moving_average=1.0, obs_type="full_state", vel_obs_scale=0.2, force_torque_obs_scale=10.0, dist_reward_scale=-10.0,
             rot_reward_scale=1.0, rot_eps=0.1, action_penalty_scale=-0.0002, success_tolerance=0.1, reach_goal_bonus=250.0,
             fall_dist=0.24, fall_penalty=0.0, max_consecutive_successes=0, max_episode_length=600.0, av_factor=0.1)
    cons = f32(0)
    ft_idx = m.sensor_body

    def one():
        nonlocal cons
        a = T.hand_pre_physics(st, rng.uniform(-1, 1, size=(n_envs, 20)).astype(f32), P)
        hand = np.ascontiguousarray(st["root"][:, 0]); o = np.ascontiguousarray(st["root"][:, 1])
        dof = np.ascontiguousarray(np.stack([st["dof_pos"], st["dof_vel"]], -1))
        out = sim.simulate(hand, dof, target=st["cur_targets"], obj=o)
        st["root"][:, 1] = o; st["dof_pos"][:] = dof[..., 0]; st["dof_vel"][:] = dof[..., 1]
        st["progress"] += 1
        ft = out["body_state"][:, ft_idx]
        T.hand_observations(st, a, ft, out["sensor"], out["dof_force"], P)
        _, cons = T.hand_reward(st, a, cons, P)
    one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return n_envs * steps / dt, dt


def run_reference_arm(args):
    """`--impl reference`: the reference's CPU pipeline cannot run here (closed Isaac Gym binary,
    SURVEY.md 8c), so this arm times the CPU PORT of the path (oracle/) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    task, n_full, _ = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    n_sample = min(n_full, 4096)
    # warm-up + K steps, each step a bounded sample (n_sample envs) of the workload
    pipe = (lambda n_, k_, c_: cpu_pipeline_hand(n_, k_, c_)) if task == "ShadowHand" else (lambda n_, k_, c_: cpu_pipeline(task, n_, k_, c_))
    pipe(n_sample, max(1, args.warmup), cores)
    v, secs = pipe(n_sample, args.steps, cores)
    line = {"metric": METRIC, "impl": "reference", "value": v, "unit": "env-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{task} num_envs={n_full} random actions U(-1,1)", "sample_envs": n_sample},
            "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
                             "sample": f"{n_sample} envs x {args.steps} control steps (oracle/aba_oracle.c f32 + oracle/tasks_np.py); "
                                       "the reference's own sim_device=cpu path needs the closed Isaac Gym binary"},
            "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ GPU arm
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from isaacgymenvs_b200 import distributed as D
    rank, local, world = D.rank_info()
    torch.cuda.set_device(local)
    D.init("nccl")
    device = f"cuda:{local}"
    task, n, bytes_per = WORKLOADS[args.workload]
    if args.num_envs:
        n = args.num_envs
    # Timing hygiene: inputs larger than L2.  One env set's live tensors (n * bytes_per, ~11 MB for Ant) would stay
    # L2-resident between steps, so the bench steps R independent env sets round-robin with R * n * bytes_per >= 1.5 x L2:
    # by the time a set is stepped again, everything it reads has been evicted and comes from HBM, while the kernel's
    # code stays warm, as in a real rollout loop.
    L2_BYTES = 126 * 1024 * 1024
    R = args.sets if args.sets > 0 else max(2, -(-int(1.5 * L2_BYTES) // (n * bytes_per)))
    R = min(R, 192)
    envs = [make_env(task, n, device, rank) for _ in range(R)]
    env = envs[0]
    A = env.num_actions
    gen = torch.Generator(device=device).manual_seed(42 + rank)
    ring = [2 * torch.rand((n, A), device=device, generator=gen) - 1 for _ in range(16)]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=device)   # > 126 MB L2

    def barrier():
        D.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: per-step CUDA events on the launching stream, L2 flushed between steps
    for k in range(args.warmup):
        for ev_ in envs:
            ev_.sim.task_step(ring[k % 16])
    barrier()
    sampler = ClockSampler(local); sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    l0 = sum(e_.sim.launch_count() for e_ in envs)
    barrier()
    # ---- headline: K steps round-robin over the R sets, one event pair around all of them
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for k in range(args.steps):
        envs[k % R].sim.task_step(ring[k % 16])
    t1.record()
    barrier()
    launches = sum(e_.sim.launch_count() for e_ in envs) - l0
    total_ms = t0.elapsed_time(t1)
    # ---- same K steps on ONE set with an explicit L2 flush between steps (also evicts the kernel's code): per-step events
    sink = torch.zeros(1, device=device)
    for k in range(args.steps):
        flush.zero_()                      # write 256 MB (> 126 MB L2): evicts the previous step's tensors ...
        sink += flush.sum()                # ... then read it back: the dirty lines are written out, L2 is left clean and cold
        ev[k][0].record()
        env.sim.task_step(ring[k % 16])
        ev[k][1].record()
    barrier()
    ms_each = [a.elapsed_time(b) for a, b in ev]
    flushed_ms = float(sum(ms_each))
    # ---- back-to-back (no flush) over the same K steps: what a rollout loop with a tiny policy sees
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    s0.record()
    for k in range(args.steps):
        env.sim.task_step(ring[k % 16])
    s1.record()
    barrier()
    b2b_ms = s0.elapsed_time(s1)
    # ---- end to end through the public API with HOST buffers (pinned): H2D actions, step, D2H results
    O = env.num_obs
    h_a = [r.cpu().pin_memory() for r in ring]
    h_obs = torch.zeros(n, O).pin_memory(); h_rew = torch.zeros(n).pin_memory()
    h_reset = torch.zeros(n, dtype=torch.long).pin_memory(); h_to = torch.zeros(n, dtype=torch.uint8).pin_memory()
    for k in range(max(3, args.warmup)):
        env.step_host(h_a[k % 16], h_obs, h_rew, h_reset, h_to)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        env.step_host(h_a[k % 16], h_obs, h_rew, h_reset, h_to)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    # ---- logging collective: per-env returns gathered once per rollout (north_star), off the step path
    all_returns = D.gather_returns(env.rew_buf)           # (world*n,) in global env order
    assert all_returns.numel() == world * n
    # ---- max over ranks
    total_ms, b2b_ms, e2e_ms, flushed_ms = D.max_over_ranks([total_ms, b2b_ms, e2e_ms, flushed_ms], device=device)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = world * n * args.steps / (total_ms * 1e-3)
    peak, peak_kind = measured_peak()
    kernel_ms = total_ms / args.steps
    achieved = n * bytes_per / (kernel_ms * 1e-3) / 1e9
    traffic, flop = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            prof = json.load(f)
        traffic = prof.get(args.workload)
        flop = prof.get(args.workload + "_fp32_flop_per_launch")
    except Exception:
        pass
    if flop is not None and n != WORKLOADS[args.workload][1]:
        flop = flop * n / WORKLOADS[args.workload][1]
    line = {
        "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{task} num_envs={n} per GPU, random actions U(-1,1), sim dt {env.cfg['sim']['dt']} x {env.cfg['sim']['substeps']} substeps",
                   "num_envs_total": world * n, "env_sets": R,
                   "timing": f"CUDA events on the launching stream around K steps; inputs larger than L2: {R} independent env sets of {n} envs stepped round-robin ({R * n * bytes_per / 1e6:.0f} MB of live tensors > 126 MB L2), so every step reads its state from HBM",
                   "collective": "none on the step path; one NCCL all_gather of per-env returns per rollout (logging)"},
        "l2_flushed": {"value": world * n * args.steps / (flushed_ms * 1e-3), "unit": "env-steps/s", "ms_per_step": flushed_ms / args.steps,
                       "note": "one env set, per-step events, L2 flushed between steps by writing and reading back a 256 MB buffer (evicts the kernel's code too)"},
        "back_to_back": {"value": world * n * args.steps / (b2b_ms * 1e-3), "unit": "env-steps/s", "ms_per_step": b2b_ms / args.steps,
                         "note": "same K steps on one env set (state stays L2-resident)"},
        "e2e": {"value": world * n * args.steps / (e2e_ms * 1e-3), "unit": "env-steps/s",
                "h2d_bytes_per_step": n * A * 4, "d2h_bytes_per_step": n * (O * 4 + 4 + 8 + 1), "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_kind": peak_kind, "kernel_ms": kernel_ms,
                     "algorithmic_bytes_per_env_step": bytes_per},
        "clocks": clocks,
    }
    if flop is not None:    # SURVEY 8d cross-check: the kernel is FP32-issue-bound, not HBM-bound
        tf = flop / (kernel_ms * 1e-3) / 1e12
        line["roofline"]["fp32"] = {"flop_per_launch": flop, "achieved": tf, "peak": 74.4, "unit": "TFLOP/s", "frac": tf / 74.4,
                                    "note": "FFMA x2 + FADD + FMUL thread-instructions from the ncu capture in profiles/; peak = 148 SM x 128 lanes x 2 x 1.965 GHz"}
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        ns = n                                            # the workload's own env count ...
        ks = max(5, min(200, int(15000 * cores / ns)))    # ... for a bounded number of control steps (a few seconds of wall time)
        if task == "ShadowHand":
            ks = max(3, min(50, int(1500 * cores / ns)))
        v, secs = cpu_pipeline(task, ns, ks, cores) if task == "Ant" else (cpu_pipeline_hand(ns, ks, cores) if task == "ShadowHand" else (None, 0))
        line["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
                                "sample": f"{ns} envs x {ks} control steps, {secs:.1f} s (oracle f32 physics + numpy obs/reward)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ant", choices=sorted(WORKLOADS))
    ap.add_argument("--num-envs", type=int, default=0)
    ap.add_argument("--sets", type=int, default=0, help="independent env sets stepped round-robin (0 = enough to exceed 1.5 x L2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
