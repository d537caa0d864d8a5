// b2g_quad_rollout.cuh -- K control steps of Ant in ONE launch (b2g_task_rollout): the open-loop rollout the benchmark's
// metric is defined on ("env-steps/s on random-action rollouts", README.md:39-51 of the reference; SURVEY.md 7 hard part 1).
//
// It is K x VecTask.step() (vec_task.py:360-408 + ant.py:281-297) with the actions of all K steps given up front:
//   for k in range(K): obs[k], rew[k], reset[k], time_outs[k] = env.step(actions[k])
// Same device functions, same order of operations as quad_loco_kernel, so the results are the ones K single steps produce;
// what disappears is everything a step pays for being its own launch: the launch itself, the model / state tile loads (the
// joint and base state stay in registers from one step to the next), the drain of the output stores (they overlap the next
// step's physics: the staging area does not alias the sweep scratch here) -- ncu on the single-step kernel attributes
// ~40 % of its time to that once-per-step code.  Actions arrive as double-buffered bulk-async tiles, one step ahead.
//
// Whole tiles of 16 envs (64 threads per CTA); the host falls back to K single steps otherwise.
#pragma once
#include "b2g_quad_kernels.cuh"

namespace b2g {

struct RollArgs {
    const float *actions;      // (K, N, A) device
    float *obs_out;            // (K, N, O): what step() returns -- the clipped observation when a clip is configured
    float *rew_out;            // (K, N)
    long long *reset_out;      // (K, N)
    uint8_t *timeout_out;      // (K, N) or null
    int K;
    int io_f4, model_f4, stage_f4;     // float4 offsets inside dynamic shared memory
};

template <int NS, int SP>
__global__ void __launch_bounds__(64, 7) quad_rollout_kernel(const float4 *__restrict__ gqm, Buffers B, const __grid_constant__ b2g_task_params P,
                                                              int N, int substeps, const __grid_constant__ RollArgs ra) {
    constexpr int BLOCK = 64, EPB = 16, nd = 4 * NS;
    __shared__ alignas(8) uint64_t mbar, mbar2, mbarA[2];
    float4 *const park = b2g_dyn_smem;
    float4 *const qm = b2g_dyn_smem + ra.model_f4;
    float *const io = reinterpret_cast<float *>(b2g_dyn_smem + ra.io_f4);
    float *const stage = reinterpret_cast<float *>(b2g_dyn_smem + ra.stage_f4);
    const int O = P.num_obs, K = ra.K;
    const int nsens6 = O - 12 - 3 * nd;
    const int env0 = blockIdx.x * EPB;
    float *const s_root = io;
    float *const s_dof = s_root + EPB * 13;
    float *const s_actb[2] = {s_dof + EPB * nd * 2, s_dof + EPB * nd * 2 + EPB * nd};
    float *const s_sens = s_actb[1] + EPB * nd;
    // per-step output staging (never aliases the sweep scratch): obs | rew | reset(i64) | timeout(u8)
    float *const t_obs = stage;
    float *const t_rew = t_obs + EPB * O;
    long long *const t_reset = reinterpret_cast<long long *>(t_rew + EPB);
    uint8_t *const t_to = reinterpret_cast<uint8_t *>(t_reset + EPB);
    // last-step staging of the remaining state tensors, in the (then dead) sweep scratch
    float *const l_obs = reinterpret_cast<float *>(park);
    float *const l_pot = l_obs + EPB * O, *const l_ppot = l_pot + EPB, *const l_up = l_ppot + EPB, *const l_head = l_up + 3 * EPB;
    long long *const l_prog = reinterpret_cast<long long *>(l_head + 3 * EPB);
    long long *const progress_b = (long long *)B.p[B2G_T_PROGRESS];
    long long *const reset_b = (long long *)B.p[B2G_T_RESET];
    float *const pot_b = (float *)B.p[B2G_T_POTENTIALS], *const ppot_b = (float *)B.p[B2G_T_PREV_POTENTIALS];
    float *const g_obs = (float *)B.p[B2G_T_OBS];
    float *g_obsc = (float *)B.p[B2G_T_OBS_CLIPPED];
    if (g_obsc == g_obs) g_obsc = nullptr;
    float *const g_sens = (float *)B.p[B2G_T_FORCE_SENSOR], *const g_dfrc = (float *)B.p[B2G_T_DOF_FORCE];
    const int gt = blockIdx.x * BLOCK + threadIdx.x;
    const int e = gt >> 2, lane = gt & 3;
    const int el = e - env0;
    if (threadIdx.x == 0) { mbar_init(&mbar, 1); mbar_init(&mbar2, 1); mbar_init(&mbarA[0], 1); mbar_init(&mbarA[1], 1); }
    __syncthreads();
    constexpr uint32_t ab = EPB * nd * 4;
    if (threadIdx.x == 0) {
        mbar_expect_tx(&mbar, quad_model_f4(NS) * 16);
        bulk_g2s(qm, gqm, quad_model_f4(NS) * 16, &mbar);
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x == 0) {
        constexpr uint32_t rb = EPB * 13 * 4, db = EPB * nd * 8;
        mbar_expect_tx(&mbar2, rb + db);
        bulk_g2s(s_root, (const float *)B.p[B2G_T_ROOT_STATE] + (size_t)env0 * 13, rb, &mbar2);
        bulk_g2s(s_dof, (const float *)B.p[B2G_T_DOF_STATE] + (size_t)env0 * nd * 2, db, &mbar2);
        mbar_expect_tx(&mbarA[0], ab);
        bulk_g2s(s_actb[0], ra.actions + (size_t)env0 * nd, ab, &mbarA[0]);
    }
    long long progress = progress_b[e];
    bool do_reset = reset_b[e] != 0;
    float potentials = pot_b[e];
    int *const rc = (int *)B.p[B2G_T_RESET_COUNT];
    uint32_t count = (uint32_t)rc[e];
    const uint32_t count0 = count;
    mbar_wait(&mbar, 0);
    mbar_wait(&mbar2, 0);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    QLane<NS, false, SP> L = make_qlane<NS, false, SP>(qm, nullptr, park, BLOCK, lane);
    attach_env_params(L, B, e, nd);
    float *const row_root = s_root + 13 * el;
    float2 *const row_dof = reinterpret_cast<float2 *>(s_dof + 2 * nd * el);
    RootState rs; load_root(row_root, rs);
    int dofi[NS], sens[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const float4 k16 = L.LK(s, 16);
        dofi[s] = q_f2i(k16.w); sens[s] = q_f2i(k16.y);
        const float2 v = row_dof[dofi[s]];
        L.q[s] = v.x; L.qd[s] = v.y;
    }
    const int total = P.control_freq_inv * substeps;
    QOutputs o;
    o.write = true;
    o.net_contact = B.p[B2G_T_NET_CONTACT] ? (float *)B.p[B2G_T_NET_CONTACT] + (size_t)e * (q_f2i(qm[7].w) >> 8) * 3 : nullptr;
    const bool stage_out = total > 0;
    o.sensor = stage_out ? s_sens + nsens6 * el : (g_sens ? g_sens + (size_t)e * nsens6 : nullptr);
    o.dof_force = g_dfrc ? g_dfrc + (size_t)e * nd : nullptr;
    const float clipo = P.clip_obs;
    const bool clip_sep = g_obsc != nullptr;
    const uint32_t gid = (uint32_t)(e + P.env_id_offset);

#pragma unroll 1
    for (int kk = 0; kk < K; kk++) {
        const bool last = kk == K - 1;
        float *const s_act = s_actb[kk & 1];
        if (threadIdx.x == 0 && !last) {                      // next step's actions, one step ahead
            mbar_expect_tx(&mbarA[(kk + 1) & 1], ab);
            bulk_g2s(s_actb[(kk + 1) & 1], ra.actions + ((size_t)(kk + 1) * N + env0) * nd, ab, &mbarA[(kk + 1) & 1]);
        }
        mbar_wait(&mbarA[kk & 1], (uint32_t)((kk >> 1) & 1));
        float *const row_act = s_act + nd * el;
        // ---- VecTask.step :374 clamp ; pre_physics_step (ant.py:281-285)
        float a_cl[NS];
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const float a = fminf(fmaxf(row_act[dofi[s]], -P.clip_actions), P.clip_actions);
            row_act[dofi[s]] = a;
            a_cl[s] = a;
            L.act[s] = a * P.joint_gears[dofi[s]] * P.power_scale;
        }
        // ---- control_freq_inv x gym.simulate
#pragma unroll 1
        for (int k = 0; k < total; k++) L.substep(rs, k == total - 1, o);
        // ---- post_physics_step (ant.py:287-297)
        progress += 1;
        if (do_reset) {                                       // reset_idx, ant.py:252-279
#pragma unroll
            for (int s = 0; s < NS; s++) {
                const int d = dofi[s];
                const float up = reset_uniform(P.seed, gid, count, d);
                const float uv = reset_uniform(P.seed, gid, count, nd + d);
                const float pos = (P.reset_pos_noise - (-P.reset_pos_noise)) * up + (-P.reset_pos_noise);
                L.q[s] = fmaxf(fminf(P.initial_dof_pos[d] + pos, P.dof_limits_upper[d]), P.dof_limits_lower[d]);
                L.qd[s] = (P.reset_vel_noise - (-P.reset_vel_noise)) * uv + (-P.reset_vel_noise);
            }
            const float *ir = (const float *)B.p[B2G_T_INITIAL_ROOT] + 13 * (size_t)e;
            load_root(ir, rs);
            potentials = t_potential(P.target[0] - rs.rp[0], P.target[1] - rs.rp[1], P.dt);
            progress = 0;
            count += 1;
        }
        // the previous step's output stores must have read their staging tiles before these are rewritten
        if ((threadIdx.x & 31) == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncthreads();
        float *const obs = t_obs + (size_t)el * O;            // per-step output tile: the observation step() returns
        float *const obs_raw = l_obs + (size_t)el * O;        // last step only, when a separate unclipped tensor exists
        const float to_t[3] = {P.target[0] - rs.rp[0], P.target[1] - rs.rp[1], 0.f};
        const float prev_potentials = potentials;
        potentials = t_potential(to_t[0], to_t[1], P.dt);
        const float isr[4] = {-0.f, -0.f, -0.f, 1.f};
        float tq[4]; t_quat_mul(rs.rq, isr, tq);
        float ang_mine;
        {
            const float qx = tq[0], qy = tq[1], qz = tq[2], qw = tq[3];
            const float ay = (lane == 1) ? 2.0f * (qw * qx + qy * qz) : (lane == 2) ? P.target[2] - rs.rp[2] : 2.0f * (qw * qz + qx * qy);
            const float ax = (lane == 1) ? qw * qw - qx * qx - qy * qy + qz * qz : (lane == 2) ? P.target[0] - rs.rp[0] : qw * qw + qx * qx - qy * qy - qz * qz;
            const float a = atan2f(ay, ax);
            ang_mine = (lane != 2 && a < 0.f) ? a + 6.2831855f : a;
        }
        const float roll = __shfl_sync(0xffffffffu, ang_mine, (threadIdx.x & 28) | 1);
        const float walk = __shfl_sync(0xffffffffu, ang_mine, (threadIdx.x & 28) | 2);
        const float yaw = __shfl_sync(0xffffffffu, ang_mine, (threadIdx.x & 28));
        auto put = [&](int idx, float v) {
            obs[idx] = clip_sep ? fminf(fmaxf(v, -clipo), clipo) : v;
            if (last && clip_sep) obs_raw[idx] = v;
        };
        float up_proj = 0.f, heading_proj = 0.f;
        float up_vec[3], heading_vec[3];
        if (lane == 0) {
            const float nrm = fmaxf(sqrtf(to_t[0] * to_t[0] + to_t[1] * to_t[1] + 0.f), 1e-9f);
            const float td[3] = {to_t[0] / nrm, to_t[1] / nrm, 0.f / nrm};
            const float b0[3] = {1.f, 0.f, 0.f}, b1[3] = {0.f, 0.f, 1.f};
            t_quat_rotate(tq, b1, up_vec, 1.f);
            t_quat_rotate(tq, b0, heading_vec, 1.f);
            up_proj = up_vec[2];
            heading_proj = (heading_vec[0] * td[0] + heading_vec[1] * td[1]) + heading_vec[2] * td[2];
            float vloc[3], wloc[3];
            t_quat_rotate(tq, rs.rv, vloc, -1.f);
            t_quat_rotate(tq, rs.rw, wloc, -1.f);
            put(0, rs.rp[2]);
            put(1, vloc[0]); put(2, vloc[1]); put(3, vloc[2]);
            put(4, wloc[0]); put(5, wloc[1]); put(6, wloc[2]);
            put(7, yaw); put(8, roll); put(9, walk - yaw); put(10, up_proj); put(11, heading_proj);
        }
        const int o_pos = 12, o_vel = 12 + nd, o_sens = 12 + 2 * nd, o_act = o_sens + nsens6;
        float actions_cost = 0.f, electricity = 0.f, at_limit = 0.f;
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const int d = dofi[s];
            const float a = a_cl[s];
            const float ps = t_unscale(L.q[s], P.dof_limits_lower[d], P.dof_limits_upper[d]);
            const float vs = L.qd[s] * P.dof_vel_scale;
            put(o_pos + d, ps); put(o_vel + d, vs); put(o_act + d, a);
            if (sens[s] >= 0) {
                float sv[6];
                if (stage_out) {
                    const float *sp_ = s_sens + nsens6 * el + 6 * sens[s];
#pragma unroll
                    for (int c = 0; c < 6; c++) sv[c] = sp_[c];
                } else {
#pragma unroll
                    for (int c = 0; c < 6; c++) sv[c] = g_sens ? g_sens[(size_t)e * nsens6 + 6 * sens[s] + c] : 0.f;
                }
                if (stage_out || g_sens) {
#pragma unroll
                    for (int c = 0; c < 6; c++) put(o_sens + 6 * sens[s] + c, sv[c] * P.contact_force_scale);
                }
            }
            actions_cost += a * a;
            at_limit += (ps > 0.99f) ? 1.f : 0.f;
            electricity += fabsf(a * vs);
        }
        actions_cost = lane_sum<4>(actions_cost);
        electricity = lane_sum<4>(electricity);
        at_limit = lane_sum<4>(at_limit);
        // reset / time-out depend on the height and the step count only: every lane of the env knows them (the flag drives
        // the NEXT step's reset_idx on all four lanes)
        const bool timed = (float)progress >= P.max_episode_length - 1.f;
        const bool died = rs.rp[2] < P.termination_height;
        do_reset = died || timed;
        if (lane == 0) {
            const float heading_reward = (heading_proj > 0.8f) ? P.heading_weight : P.heading_weight * heading_proj / 0.8f;
            const float up_reward = (up_proj > 0.93f) ? P.up_weight : 0.f;
            const float progress_reward = potentials - prev_potentials;
            float total_r = progress_reward + P.alive_reward + up_reward + heading_reward - P.actions_cost_scale * actions_cost -
                            P.energy_cost_scale * electricity - at_limit * P.joints_at_limit_cost_scale;
            if (died) total_r = P.death_cost;
            t_rew[el] = total_r; t_reset[el] = do_reset ? 1 : 0;
            t_to[el] = (uint8_t)(timed && do_reset);                                                      // vec_task.py:394
            if (last) {
                l_pot[el] = potentials; l_ppot[el] = prev_potentials; l_prog[el] = progress;
                l_up[3 * el] = up_vec[0]; l_up[3 * el + 1] = up_vec[1]; l_up[3 * el + 2] = up_vec[2];
                l_head[3 * el] = heading_vec[0]; l_head[3 * el + 1] = heading_vec[1]; l_head[3 * el + 2] = heading_vec[2];
                if (count != count0) rc[e] = (int)count;
            }
        }
        if (last) {
#pragma unroll
            for (int s = 0; s < NS; s++) row_dof[dofi[s]] = make_float2(L.q[s], L.qd[s]);
            if (lane == 0) store_root(row_root, rs);
        }
        fence_async_smem();
        __syncthreads();
        {
            const size_t e0 = (size_t)env0, kN = (size_t)kk * N + e0;
            if (threadIdx.x == 0) {
                bulk_s2g(ra.obs_out + kN * O, t_obs, (uint32_t)(EPB * O * 4));
                bulk_s2g(ra.rew_out + kN, t_rew, EPB * 4);
                if (last) {
                    float *const g_act_out = (float *)B.p[B2G_T_ACTIONS];
                    bulk_s2g(g_obs + e0 * O, clip_sep ? l_obs : t_obs, (uint32_t)(EPB * O * 4));
                    if (clip_sep) bulk_s2g(g_obsc + e0 * O, t_obs, (uint32_t)(EPB * O * 4));
                    bulk_s2g((float *)B.p[B2G_T_ROOT_STATE] + e0 * 13, s_root, EPB * 13 * 4);
                    bulk_s2g((float *)B.p[B2G_T_DOF_STATE] + e0 * nd * 2, s_dof, (uint32_t)(EPB * nd * 8));
                    if (g_act_out) bulk_s2g(g_act_out + e0 * nd, s_act, (uint32_t)(EPB * nd * 4));
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            } else if (threadIdx.x == 32) {
                bulk_s2g(ra.reset_out + kN, t_reset, EPB * 8);
                if (ra.timeout_out) bulk_s2g(ra.timeout_out + kN, t_to, EPB);
                if (last) {
                    if (stage_out && g_sens && nsens6) bulk_s2g(g_sens + e0 * nsens6, s_sens, (uint32_t)(EPB * nsens6 * 4));
                    bulk_s2g((float *)B.p[B2G_T_REW] + e0, t_rew, EPB * 4);
                    bulk_s2g(pot_b + e0, l_pot, EPB * 4);
                    bulk_s2g(ppot_b + e0, l_ppot, EPB * 4);
                    if (B.p[B2G_T_UP_VEC]) bulk_s2g((float *)B.p[B2G_T_UP_VEC] + 3 * e0, l_up, EPB * 12);
                    if (B.p[B2G_T_HEADING_VEC]) bulk_s2g((float *)B.p[B2G_T_HEADING_VEC] + 3 * e0, l_head, EPB * 12);
                    bulk_s2g(reset_b + e0, t_reset, EPB * 8);
                    bulk_s2g(progress_b + e0, l_prog, EPB * 8);
                    if (B.p[B2G_T_TIMEOUT]) bulk_s2g((uint8_t *)B.p[B2G_T_TIMEOUT] + e0, t_to, EPB);
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
    }
    if ((threadIdx.x & 31) == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

}  // namespace b2g
