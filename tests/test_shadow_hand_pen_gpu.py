"""GPU: the fused ShadowHand step with objectType pen against the reference's own pre / post_physics_step (golden case "p",
tests/golden/make_golden_hand.py --pen): reset_idx poses the pen with randomize_rotation_pen (shadow_hand.py:626-629, :810-813)
and compute_hand_reward runs with ignore_z_rot (:421, :758-759).  (Its own file, sorted last: the kernel branch it covers was
written after the round's last GPU visit -- the numpy twin of the same arithmetic is pinned on the CPU in test_oracle_tasks.py.)"""
import os
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_hand_step_with_the_pen_matches_reference_golden():
    from tests.test_gpu_parity import _hand_env
    gold = np.load(os.path.join(GOLD, "shadow_hand_pen.npz"))
    gi = lambda k: gold[f"p_in_{k}"]
    go = lambda k: gold[f"p_out_{k}"]
    n = gi("reset").shape[0]
    env = _hand_env(n, "p", "full_state", objectType="pen")
    assert env.sim.task.object_is_pen == 1 and env.sim.task.success_tolerance == np.float32(0.2)
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
    # the reset pens' orientation: randomize_rotation_pen, not randomize_rotation
    np.testing.assert_allclose(c(env.root_state_tensor), go("root"), rtol=0, atol=1e-6)
    np.testing.assert_allclose(c(env.goal_states), go("goal_states"), rtol=0, atol=1e-6)
    np.testing.assert_allclose(c(env.dof_state), go("dof_state"), rtol=0, atol=1e-6)
    # the reward with twice the success tolerance
    np.testing.assert_allclose(c(env.rew_buf), go("rew"), rtol=3e-6, atol=3e-5)
    assert np.array_equal(c(env.reset_buf), go("reset")) and np.array_equal(c(env.reset_goal_buf), go("reset_goal"))
    assert np.array_equal(c(env.successes), go("successes")) and gi("reset").sum() > 10
    # reset_done poses the flagged pens the same way (b2g_reset_flagged)
    env.reset_buf.fill_(1)
    rc = c(env.reset_count).copy()
    env.reset_done(); torch.cuda.synchronize()
    from oracle import tasks_np as T
    D = env.num_shadow_hand_dofs
    q = c(env.root_state_tensor).reshape(n, 3, 13)[:, 1, 3:7]
    for e in (0, 5, n - 1):
        r = T.hand_rand_floats(int(gold["seed"]), e, int(rc[e]), 2 * D + 7)
        want = T.randomize_rotation_pen(r[3:4], r[4:5], 0.3, np.array([[1, 0, 0]], np.float32), np.array([[0, 1, 0]], np.float32), np.array([[0, 0, 1]], np.float32))[0]
        assert np.abs(q[e] - want).max() < 1e-6
    env.sim.close()
