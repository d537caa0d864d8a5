// b2g_quad_kernels.cuh -- kernels on the quad sub-step (b2g_quad.cuh).  Included by b200gym.cu after Buffers /
// TileArgs / the task helpers.
//
//   quad_simulate_kernel   gym.simulate() (vec_task.py:379-382) for "4 chains on a free base" articulations
//   quad_loco_kernel       one whole VecTask.step() of Ant (vec_task.py:360-408 + ant.py:281-297), tiles by bulk copy
#pragma once
#include "b2g_quad.cuh"

namespace b2g {

// dynamic shared memory of the quad kernels: [park: quad_park_f4(NS) x BLOCK float4][tiles ...][quad model]
template <int NS, bool HF>
__device__ __forceinline__ QLane<NS, HF> make_qlane(const float4 *qm, const int16_t *hf, float4 *park_base, int block, int lane) {
    QLane<NS, HF> L;
    L.qm = qm; L.hf = hf; L.park = park_base + threadIdx.x; L.pstride = block; L.lane = lane; L.env_mu = -1.f;
    return L;
}

template <int NS, bool HF, int BLOCK>
__global__ void __launch_bounds__(BLOCK) quad_simulate_kernel(const float4 *__restrict__ gqm, const int16_t *__restrict__ hf, Buffers B, int N, int substeps) {
    float4 *const park = b2g_dyn_smem;
    float4 *const qm = b2g_dyn_smem + quad_park_f4(NS) * BLOCK;
    for (int i = threadIdx.x; i < quad_model_f4(NS); i += BLOCK) qm[i] = gqm[i];
    __syncthreads();
    const int gt = blockIdx.x * BLOCK + threadIdx.x;
    const int env = gt >> 2, lane = gt & 3;
    const bool valid = env < N;
    const int e = valid ? env : N - 1;
    constexpr int nd = 4 * NS;
    QLane<NS, HF> L = make_qlane<NS, HF>(qm, hf, park, BLOCK, lane);
    float *const root_row = (float *)B.p[B2G_T_ROOT_STATE] + 13 * (size_t)e;
    RootState rs; load_root(root_row, rs);
    float2 *const d = (float2 *)B.p[B2G_T_DOF_STATE] + (size_t)e * nd;
    const float *act = (const float *)B.p[B2G_T_DOF_ACTUATION];
    int dofi[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) {
        dofi[s] = q_f2i(L.LK(s, 16).w);
        const float2 v = d[dofi[s]];
        L.q[s] = v.x; L.qd[s] = v.y; L.act[s] = act ? act[(size_t)e * nd + dofi[s]] : 0.f;
    }
    const int cnt = q_f2i(qm[7].w), n_sens = cnt & 255, n_body = cnt >> 8;
    float *fs = (float *)B.p[B2G_T_FORCE_SENSOR], *df = (float *)B.p[B2G_T_DOF_FORCE], *nc = (float *)B.p[B2G_T_NET_CONTACT];
    QOutputs o;
    o.sensor = fs ? fs + (size_t)e * n_sens * 6 : nullptr;
    o.dof_force = df ? df + (size_t)e * nd : nullptr;
    o.net_contact = nc ? nc + (size_t)e * n_body * 3 : nullptr;
    o.write = valid;
    // `substeps` is a kernel parameter (warp-uniform by construction): the shuffles inside need no re-convergence code
#pragma unroll 1
    for (int k = 0; k < substeps; k++) L.substep(rs, k == substeps - 1, o);
    if (!valid) return;
#pragma unroll
    for (int s = 0; s < NS; s++) d[dofi[s]] = make_float2(L.q[s], L.qd[s]);
    if (lane == 0) store_root(root_row, rs);
}

// -------------------------------------------------------------------------------------------
// One whole VecTask.step() of Ant on the quad sub-step.  Same data movement as loco_step_kernel: every tensor of
// the step is ONE bulk-async (TMA) copy per block in and out; whole tiles only (the host falls back to
// loco_step_kernel otherwise).
#ifndef B2G_QUAD_MINBLOCKS
#define B2G_QUAD_MINBLOCKS(BLOCK) ((BLOCK) == 128 ? 4 : ((BLOCK) == 64 ? 7 : 14))
#endif
template <int NS, int BLOCK, bool HOSTIO>
__global__ void __launch_bounds__(BLOCK, B2G_QUAD_MINBLOCKS(BLOCK)) quad_loco_kernel(
    const float4 *__restrict__ gqm, Buffers B, const __grid_constant__ b2g_task_params P, const float *__restrict__ actions_in, int N, int substeps, TileArgs ta) {
    __shared__ alignas(8) uint64_t mbar, mbar2;
    constexpr int EPB = BLOCK / 4;
    constexpr int nd = 4 * NS;
    float4 *const park = b2g_dyn_smem;
    float4 *const qm = b2g_dyn_smem + ta.model_f4;
    float *const io = reinterpret_cast<float *>(b2g_dyn_smem + ta.io_f4);
    const int O = P.num_obs;
    const int nsens6 = O - 12 - 3 * nd;                      // 6 * nsens, from the obs layout (checked by b2g_set_task)
    const int env0 = blockIdx.x * EPB;
    // ---- in/out tile region: root | dof | act | sensors
    float *const s_root = io;
    float *const s_dof = s_root + EPB * 13;
    float *const s_act = s_dof + EPB * nd * 2;
    float *const s_sens = s_act + EPB * nd;
    long long *const progress_b = (long long *)B.p[B2G_T_PROGRESS];
    long long *const reset_b = (long long *)B.p[B2G_T_RESET];
    float *const pot_b = (float *)B.p[B2G_T_POTENTIALS], *const ppot_b = (float *)B.p[B2G_T_PREV_POTENTIALS];
    const int gt = blockIdx.x * BLOCK + threadIdx.x;
    const int e = gt >> 2, lane = gt & 3;                    // whole tiles: every env of the block exists
    const int el = e - env0;
    // ---- prologue (programmatic dependent launch: everything before griddepcontrol.wait overlaps the previous step's tail)
    if (threadIdx.x == 0) { mbar_init(&mbar, 1); mbar_init(&mbar2, 1); }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&mbar, quad_model_f4(NS) * 16);
        bulk_g2s(qm, gqm, quad_model_f4(NS) * 16, &mbar);
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x == 0) {
        constexpr uint32_t rb = EPB * 13 * 4, db = EPB * nd * 8, ab = EPB * nd * 4;
        mbar_expect_tx(&mbar2, rb + db + (HOSTIO ? 0u : ab));
        bulk_g2s(s_root, (const float *)B.p[B2G_T_ROOT_STATE] + (size_t)env0 * 13, rb, &mbar2);
        bulk_g2s(s_dof, (const float *)B.p[B2G_T_DOF_STATE] + (size_t)env0 * nd * 2, db, &mbar2);
        if (!HOSTIO) bulk_g2s(s_act, actions_in + (size_t)env0 * nd, ab, &mbar2);
    }
    if (HOSTIO) {                  // actions straight from pinned host memory
        const float4 *src = reinterpret_cast<const float4 *>(ta.h_act + (size_t)env0 * nd);
        float4 *dst = reinterpret_cast<float4 *>(s_act);
        for (int i = threadIdx.x; i < EPB * nd / 4; i += BLOCK) dst[i] = src[i];
        __syncthreads();
    }
    const long long progress_in = progress_b[e];
    const long long reset_in = reset_b[e];
    const float potentials_in = pot_b[e];
    int *const rc = (int *)B.p[B2G_T_RESET_COUNT];
    uint32_t count = 0;
    if (reset_in != 0) count = (uint32_t)rc[e];              // read here, written after the physics: no intra-warp race
    mbar_wait(&mbar, 0);
    mbar_wait(&mbar2, 0);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    QLane<NS, false> L = make_qlane<NS, false>(qm, nullptr, park, BLOCK, lane);
    float *const row_root = s_root + 13 * el;
    float2 *const row_dof = reinterpret_cast<float2 *>(s_dof + 2 * nd * el);
    float *const row_act = s_act + nd * el;
    RootState rs; load_root(row_root, rs);

    // ---- VecTask.step :374 clamp ; pre_physics_step (ant.py:281-285)
    int dofi[NS], sens[NS];
    float a_cl[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const float4 k16 = L.LK(s, 16);
        dofi[s] = q_f2i(k16.w); sens[s] = q_f2i(k16.y);
        const float2 v = row_dof[dofi[s]];
        const float a = fminf(fmaxf(row_act[dofi[s]], -P.clip_actions), P.clip_actions);
        row_act[dofi[s]] = a;                                 // the tile becomes the clamped-action output
        a_cl[s] = a;
        L.q[s] = v.x; L.qd[s] = v.y; L.act[s] = a * P.joint_gears[dofi[s]] * P.power_scale;
    }

    // ---- control_freq_inv x gym.simulate (vec_task.py:379-382); 0: the observation reads the tensors as they stand
    const int total = P.control_freq_inv * substeps;          // kernel parameters: warp-uniform trip count
    float *const g_sens = (float *)B.p[B2G_T_FORCE_SENSOR], *const g_dfrc = (float *)B.p[B2G_T_DOF_FORCE];
    QOutputs o;
    o.write = true;
    o.net_contact = B.p[B2G_T_NET_CONTACT] ? (float *)B.p[B2G_T_NET_CONTACT] + (size_t)e * (q_f2i(qm[7].w) >> 8) * 3 : nullptr;
    const bool stage_out = total > 0;
    o.sensor = stage_out ? s_sens + nsens6 * el : (g_sens ? g_sens + (size_t)e * nsens6 : nullptr);
    o.dof_force = g_dfrc ? g_dfrc + (size_t)e * nd : nullptr;
#pragma unroll 1
    for (int k = 0; k < total; k++) L.substep(rs, k == total - 1, o);

    // ---- post_physics_step (ant.py:287-297): progress, reset_idx, observations, reward
    long long progress = progress_in + 1;
    float potentials = potentials_in;
    const bool do_reset = reset_in != 0;
    if (do_reset) {                                           // reset_idx, ant.py:252-279
        const uint32_t gid = (uint32_t)(e + P.env_id_offset);
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
        if (lane == 0) rc[e] = (int)(count + 1);
    }
#pragma unroll
    for (int s = 0; s < NS; s++) row_dof[dofi[s]] = make_float2(L.q[s], L.qd[s]);
    if (lane == 0) store_root(row_root, rs);

    // the parking area is dead from here on: it becomes the output staging area
    // layout (floats unless noted): obs | obs_clipped? | rew | pot | ppot | up(3) | head(3) | reset(i64) | progress(i64) | timeout(u8)
    __syncthreads();
    float *const g_obs = (float *)B.p[B2G_T_OBS];
    float *g_obsc = (float *)B.p[B2G_T_OBS_CLIPPED];
    if (g_obsc == g_obs) g_obsc = nullptr;
    float *const t_obs = reinterpret_cast<float *>(b2g_dyn_smem);
    float *const t_obsc = t_obs + EPB * O;
    float *const t_rew = t_obsc + (g_obsc ? EPB * O : 0);
    float *const t_pot = t_rew + EPB, *const t_ppot = t_pot + EPB, *const t_up = t_ppot + EPB, *const t_head = t_up + 3 * EPB;
    long long *const t_reset = reinterpret_cast<long long *>(t_head + 3 * EPB), *const t_prog = t_reset + EPB;
    uint8_t *const t_to = reinterpret_cast<uint8_t *>(t_prog + EPB);
    float *const obs = t_obs + (size_t)el * O;
    float *const obsc = g_obsc ? t_obsc + (size_t)el * O : nullptr;

    // compute_observations (ant.py:374-408)
    LocoRootObs ro;
    loco_root_obs(P, rs.rp, rs.rq, rs.rv, rs.rw, false, ro);
    const float prev_potentials = potentials;
    potentials = ro.potentials;
    const float clipo = P.clip_obs;
    auto put = [&](int idx, float v) {
        obs[idx] = v;
        if (obsc) obsc[idx] = fminf(fmaxf(v, -clipo), clipo);
    };
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 12; c++) put(c, ro.o[c]);
    }
    // layout: ant.py:401-406  [12 | nd pos | nd vel | 6*nsens sensors | nd actions]
    const int o_pos = 12, o_vel = 12 + nd, o_sens = 12 + 2 * nd, o_act = o_sens + nsens6;
    float actions_cost = 0.f, electricity = 0.f, at_limit = 0.f;
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int d = dofi[s];
        const float a = a_cl[s];
        const float ps = t_unscale(L.q[s], P.dof_limits_lower[d], P.dof_limits_upper[d]);
        const float vs = L.qd[s] * P.dof_vel_scale;
        put(o_pos + d, ps); put(o_vel + d, vs); put(o_act + d, a);
        if (sens[s] >= 0 && o.sensor) {
#pragma unroll
            for (int c = 0; c < 6; c++) put(o_sens + 6 * sens[s] + c, o.sensor[6 * sens[s] + c] * P.contact_force_scale);
        }
        actions_cost += a * a;                                 // compute_ant_reward, ant.py:353-355
        at_limit += (ps > 0.99f) ? 1.f : 0.f;
        electricity += fabsf(a * vs);
    }
    {
        const int rsens = q_f2i(qm[3].w);
        if (lane == 0 && rsens >= 0 && o.sensor) {
#pragma unroll
            for (int c = 0; c < 6; c++) put(o_sens + 6 * rsens + c, o.sensor[6 * rsens + c] * P.contact_force_scale);
        }
    }
    actions_cost = lane_sum<4>(actions_cost);
    electricity = lane_sum<4>(electricity);
    at_limit = lane_sum<4>(at_limit);
    if (lane == 0) {
        const float heading_proj = ro.o[11], up_proj = ro.o[10], height = ro.o[0];
        const float heading_reward = (heading_proj > 0.8f) ? P.heading_weight : P.heading_weight * heading_proj / 0.8f;
        const float up_reward = (up_proj > 0.93f) ? P.up_weight : 0.f;
        const float progress_reward = potentials - prev_potentials;
        float total_r = progress_reward + P.alive_reward + up_reward + heading_reward - P.actions_cost_scale * actions_cost -
                        P.energy_cost_scale * electricity - at_limit * P.joints_at_limit_cost_scale;
        long long reset = 0;
        if (height < P.termination_height) { total_r = P.death_cost; reset = 1; }
        if ((float)progress >= P.max_episode_length - 1.f) reset = 1;
        const uint8_t tout = (uint8_t)(((float)progress >= P.max_episode_length - 1.f) && reset != 0);   // vec_task.py:394
        t_rew[el] = total_r; t_reset[el] = reset; t_prog[el] = progress; t_pot[el] = potentials; t_ppot[el] = prev_potentials;
        t_up[3 * el] = ro.up_vec[0]; t_up[3 * el + 1] = ro.up_vec[1]; t_up[3 * el + 2] = ro.up_vec[2];
        t_head[3 * el] = ro.heading_vec[0]; t_head[3 * el + 1] = ro.heading_vec[1]; t_head[3 * el + 2] = ro.heading_vec[2];
        t_to[el] = tout;
    }
    fence_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
        const size_t e0 = (size_t)env0;
        float *const g_act_out = (float *)B.p[B2G_T_ACTIONS];
        bulk_s2g((float *)B.p[B2G_T_ROOT_STATE] + e0 * 13, s_root, EPB * 13 * 4);
        bulk_s2g((float *)B.p[B2G_T_DOF_STATE] + e0 * nd * 2, s_dof, (uint32_t)(EPB * nd * 8));
        if (g_act_out) bulk_s2g(g_act_out + e0 * nd, s_act, (uint32_t)(EPB * nd * 4));
        if (stage_out && g_sens && nsens6) bulk_s2g(g_sens + e0 * nsens6, s_sens, (uint32_t)(EPB * nsens6 * 4));
        bulk_s2g(g_obs + e0 * O, t_obs, (uint32_t)(EPB * O * 4));
        if (g_obsc) bulk_s2g(g_obsc + e0 * O, t_obsc, (uint32_t)(EPB * O * 4));
        bulk_s2g((float *)B.p[B2G_T_REW] + e0, t_rew, EPB * 4);
        bulk_s2g(pot_b + e0, t_pot, EPB * 4);
        bulk_s2g(ppot_b + e0, t_ppot, EPB * 4);
        if (B.p[B2G_T_UP_VEC]) bulk_s2g((float *)B.p[B2G_T_UP_VEC] + 3 * e0, t_up, EPB * 12);
        if (B.p[B2G_T_HEADING_VEC]) bulk_s2g((float *)B.p[B2G_T_HEADING_VEC] + 3 * e0, t_head, EPB * 12);
        bulk_s2g(reset_b + e0, t_reset, EPB * 8);
        bulk_s2g(progress_b + e0, t_prog, EPB * 8);
        if (B.p[B2G_T_TIMEOUT]) bulk_s2g((uint8_t *)B.p[B2G_T_TIMEOUT] + e0, t_to, EPB);
        bulk_commit_wait();
    }
    if (HOSTIO) {                  // host copies of what VecTask.step returns (vec_task.py:402-408), straight over PCIe
        const size_t e0 = (size_t)env0;
        auto copy16 = [&](void *dst, const void *src, int bytes) {
            float4 *d = reinterpret_cast<float4 *>(dst); const float4 *sp = reinterpret_cast<const float4 *>(src);
            for (int i = threadIdx.x; i < bytes / 16; i += BLOCK) d[i] = sp[i];
        };
        if (ta.h_obs) copy16(ta.h_obs + e0 * O, g_obsc ? t_obsc : t_obs, EPB * O * 4);
        if (ta.h_rew) copy16(ta.h_rew + e0, t_rew, EPB * 4);
        if (ta.h_reset) copy16(ta.h_reset + e0, t_reset, EPB * 8);
        if (ta.h_timeout) copy16(ta.h_timeout + e0, t_to, EPB);
    }
}

}  // namespace b2g
