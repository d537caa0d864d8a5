// b2g_quad_kernels.cuh -- kernels on the quad sub-step (b2g_quad.cuh).  Included by b200gym.cu after Buffers /
// TileArgs / the task helpers.
//
//   quad_simulate_kernel   gym.simulate() (vec_task.py:379-382) for "4 chains on a free base" articulations
//   quad_loco_kernel       one whole VecTask.step() of Ant (vec_task.py:360-408 + ant.py:281-297), tiles by bulk copy
#pragma once
#include "b2g_quad.cuh"
#include "b2g_tasks.cuh"

namespace b2g {

// dynamic shared memory of the quad kernels: [park: quad_park_f4(NS) x BLOCK float4][tiles ...][quad model]
template <int NS, bool HF, int SP>
__device__ __forceinline__ QLane<NS, HF, SP> make_qlane(const float4 *qm, const int16_t *hf, float4 *park_base, int block, int lane) {
    QLane<NS, HF, SP> L;
    L.qm = qm; L.hf = hf; L.park = park_base + threadIdx.x; L.pstride = block; L.lane = lane; L.env_mu = -1.f;
    L.dr_mass = nullptr; L.dr_dof = nullptr;
    return L;
}

// per-env physical parameters (domain randomisation tensors; null = the model's own)
template <class QL>
__device__ __forceinline__ void attach_env_params(QL &L, const Buffers &B, int e, int nd, bool arrays = true) {
    const float *ms = arrays ? (const float *)B.p[B2G_T_ENV_MASS_SCALE] : nullptr;
    const float4 *dp = arrays ? (const float4 *)B.p[B2G_T_ENV_DOF_PROPS] : nullptr;
    const float *envmu = (const float *)B.p[B2G_T_ENV_FRICTION];
    if (ms) L.dr_mass = ms + (size_t)e * (nd + 1);
    if (dp) L.dr_dof = dp + (size_t)e * nd;
    if (envmu) L.env_mu = 0.5f * (envmu[e] + L.qm[18].x);             // PhysX default combine mode: the average of the two materials
}

template <int NS, bool HF, int SP, int BLOCK>
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
    QLane<NS, HF, SP> L = make_qlane<NS, HF, SP>(qm, hf, park, BLOCK, lane);
    attach_env_params(L, B, e, nd);
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
// LEAN: no per-env physical parameters, no dof-force / net-contact outputs bound (the plain Ant task): those pointers
// become compile-time nulls -- their branches and the registers they occupy across the sub-step loop disappear
template <int NS, int SP, int BLOCK, bool HOSTIO, bool LEAN = false>
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
    // read here (unconditionally: a load that depended on reset_in would stall the prologue on it), written after the
    // physics by lane 0: no intra-warp race
    const uint32_t count = (uint32_t)rc[e];
    mbar_wait(&mbar, 0);
    mbar_wait(&mbar2, 0);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    QLane<NS, false, SP> L = make_qlane<NS, false, SP>(qm, nullptr, park, BLOCK, lane);
    if (!LEAN) attach_env_params(L, B, e, nd);
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
    o.net_contact = (!LEAN && B.p[B2G_T_NET_CONTACT]) ? (float *)B.p[B2G_T_NET_CONTACT] + (size_t)e * (q_f2i(qm[7].w) >> 8) * 3 : nullptr;
    const bool stage_out = total > 0;
    o.sensor = stage_out ? s_sens + nsens6 * el : (g_sens ? g_sens + (size_t)e * nsens6 : nullptr);
    o.dof_force = (!LEAN && g_dfrc) ? g_dfrc + (size_t)e * nd : nullptr;
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

    // compute_observations (ant.py:374-408).  The three Euler / heading angles are three atan2f calls on different
    // arguments: the lanes of the env take one each (same instruction stream, different data) and hand it to lane 0.
    const float to_t[3] = {P.target[0] - rs.rp[0], P.target[1] - rs.rp[1], 0.f};
    const float prev_potentials = potentials;                  // prev_potentials_new = potentials.clone(), ant.py:390
    potentials = t_potential(to_t[0], to_t[1], P.dt);
    const float isr[4] = {-0.f, -0.f, -0.f, 1.f};
    float tq[4]; t_quat_mul(rs.rq, isr, tq);                   // compute_heading_and_up, torch_jit_utils.py:247-262
    float ang_mine;
    {
        const float qx = tq[0], qy = tq[1], qz = tq[2], qw = tq[3];
        // get_euler_xyz :175-195 (roll, yaw) and compute_rot :265-276 (walk_target_angle); lane 3 duplicates lane 0
        const float ay = (lane == 1) ? 2.0f * (qw * qx + qy * qz) : (lane == 2) ? P.target[2] - rs.rp[2] : 2.0f * (qw * qz + qx * qy);
        const float ax = (lane == 1) ? qw * qw - qx * qx - qy * qy + qz * qz : (lane == 2) ? P.target[0] - rs.rp[0] : qw * qw + qx * qx - qy * qy - qz * qz;
        const float a = atan2f(ay, ax);
        // torch.remainder(a, 2 pi) for a in [-pi, pi] (fmod leaves such an `a` unchanged); the walk angle is not wrapped
        ang_mine = (lane != 2 && a < 0.f) ? a + 6.2831855f : a;
    }
    const float roll = __shfl_sync(0xffffffffu, ang_mine, (threadIdx.x & 28) | 1);
    const float walk = __shfl_sync(0xffffffffu, ang_mine, (threadIdx.x & 28) | 2);
    const float yaw = __shfl_sync(0xffffffffu, ang_mine, (threadIdx.x & 28));
    const float clipo = P.clip_obs;
    auto put = [&](int idx, float v) {
        obs[idx] = v;
        if (obsc) obsc[idx] = fminf(fmaxf(v, -clipo), clipo);
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
        if (sens[s] >= 0) {                                    // typed pointers: shared-memory loads of the staged tile, not generic ones
            float sv[6];
            if (stage_out) {
                const float *sp_ = s_sens + nsens6 * el + 6 * sens[s];
#pragma unroll
                for (int c = 0; c < 6; c++) sv[c] = sp_[c];
            } else {                                           // no simulate: the tensor as it stands (golden-vector mode)
#pragma unroll
                for (int c = 0; c < 6; c++) sv[c] = g_sens ? g_sens[(size_t)e * nsens6 + 6 * sens[s] + c] : 0.f;
            }
            if (stage_out || g_sens) {
#pragma unroll
                for (int c = 0; c < 6; c++) put(o_sens + 6 * sens[s] + c, sv[c] * P.contact_force_scale);
            }
        }
        actions_cost += a * a;                                 // compute_ant_reward, ant.py:353-355
        at_limit += (ps > 0.99f) ? 1.f : 0.f;
        electricity += fabsf(a * vs);
    }
    {
        const int rsens = q_f2i(qm[3].w);                      // a sensor on the base itself (not Ant): through the tensor
        const float *sp_ = stage_out ? s_sens + nsens6 * el : (g_sens ? g_sens + (size_t)e * nsens6 : nullptr);
        if (lane == 0 && rsens >= 0 && sp_) {
#pragma unroll
            for (int c = 0; c < 6; c++) put(o_sens + 6 * rsens + c, sp_[6 * rsens + c] * P.contact_force_scale);
        }
    }
    actions_cost = lane_sum<4>(actions_cost);
    electricity = lane_sum<4>(electricity);
    at_limit = lane_sum<4>(at_limit);
    if (lane == 0) {
        const float height = rs.rp[2];
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
        t_up[3 * el] = up_vec[0]; t_up[3 * el + 1] = up_vec[1]; t_up[3 * el + 2] = up_vec[2];
        t_head[3 * el] = heading_vec[0]; t_head[3 * el + 1] = heading_vec[1]; t_head[3 * el + 2] = heading_vec[2];
        t_to[el] = tout;
    }
    fence_async_smem();
    __syncthreads();
    // one bulk-async (TMA) store per output tensor.  The stores are dealt to the FIRST THREAD of each warp at compile time
    // (`threadIdx.x == 32 w` branches, each a single-thread region the compiler keeps on the uniform datapath): their issue
    // -- address arithmetic + UBLKCP each -- runs in parallel instead of as one thread's serial tail
    {
        const size_t e0 = (size_t)env0;
        float *const g_act_out = (float *)B.p[B2G_T_ACTIONS];
        constexpr int NW = BLOCK / 32;
        auto issue = [&](int w) {
            int k = 0;
#define B2G_ST(COND, DST, SRC, BYTES) do { if ((k++ % NW) == w) { if (COND) bulk_s2g(DST, SRC, BYTES); } } while (0)
            B2G_ST(true, g_obs + e0 * O, t_obs, (uint32_t)(EPB * O * 4));
            B2G_ST(g_obsc != nullptr, g_obsc + e0 * O, t_obsc, (uint32_t)(EPB * O * 4));
            B2G_ST(true, (float *)B.p[B2G_T_ROOT_STATE] + e0 * 13, s_root, EPB * 13 * 4);
            B2G_ST(true, (float *)B.p[B2G_T_DOF_STATE] + e0 * nd * 2, s_dof, (uint32_t)(EPB * nd * 8));
            B2G_ST(g_act_out != nullptr, g_act_out + e0 * nd, s_act, (uint32_t)(EPB * nd * 4));
            B2G_ST(stage_out && g_sens && nsens6, g_sens + e0 * nsens6, s_sens, (uint32_t)(EPB * nsens6 * 4));
            B2G_ST(true, (float *)B.p[B2G_T_REW] + e0, t_rew, EPB * 4);
            B2G_ST(true, pot_b + e0, t_pot, EPB * 4);
            B2G_ST(true, ppot_b + e0, t_ppot, EPB * 4);
            B2G_ST(B.p[B2G_T_UP_VEC] != nullptr, (float *)B.p[B2G_T_UP_VEC] + 3 * e0, t_up, EPB * 12);
            B2G_ST(B.p[B2G_T_HEADING_VEC] != nullptr, (float *)B.p[B2G_T_HEADING_VEC] + 3 * e0, t_head, EPB * 12);
            B2G_ST(true, reset_b + e0, t_reset, EPB * 8);
            B2G_ST(true, progress_b + e0, t_prog, EPB * 8);
            if (EPB % 16 == 0) B2G_ST(B.p[B2G_T_TIMEOUT] != nullptr, (uint8_t *)B.p[B2G_T_TIMEOUT] + e0, t_to, EPB);   // bulk copies move multiples of 16 bytes
#undef B2G_ST
            bulk_commit_wait();
        };
        if (EPB % 16 != 0 && B.p[B2G_T_TIMEOUT] && threadIdx.x < EPB) ((uint8_t *)B.p[B2G_T_TIMEOUT])[e0 + threadIdx.x] = t_to[threadIdx.x];
        if (threadIdx.x == 0) issue(0);
        else if (NW > 1 && threadIdx.x == 32) issue(1);
        else if (NW > 2 && threadIdx.x == 64) issue(2);
        else if (NW > 3 && threadIdx.x == 96) issue(3);
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
        if (ta.h_timeout) {
            if (EPB % 16 == 0) copy16(ta.h_timeout + e0, t_to, EPB);
            else if (threadIdx.x < EPB) ta.h_timeout[e0 + threadIdx.x] = t_to[threadIdx.x];
        }
    }
}


// -------------------------------------------------------------------------------------------
// AnymalTerrain kernel 1 on the quad sub-step (same contract as anymal_physics_kernel, b2g_anymal.cuh): pre_physics_step
// (PD torque + gym.simulate x decimation, anymal_terrain.py:441-451) + the control_freq_inv simulates of VecTask.step
// (vec_task.py:379-382) + post_physics_step up to compute_reward (:453-475).  The joint state stays in registers across
// the 5 sub-steps; the PD law reads it there.
// DR = false: no per-env link-mass / joint-property arrays bound -- their pointers are compile-time nulls (smaller code: this kernel runs one
// warp per scheduler, so instruction fetch is exposed: 27 % of its stall cycles are "no instruction")
template <bool HF, int BLOCK, bool DR = true>
__global__ void __launch_bounds__(BLOCK) quad_anymal_physics_kernel(const float4 *__restrict__ gqm, const int16_t *__restrict__ hf,
                                                                    Buffers B, const __grid_constant__ b2g_anymal_params P,
                                                                    const float *__restrict__ actions_in, int N, int substeps, unsigned step_counter) {
    constexpr int NS = 3, nd = 12;
    __shared__ float s_part[BLOCK / 32];
    float4 *const park = b2g_dyn_smem;
    float4 *const qm = b2g_dyn_smem + quad_park_f4(NS) * BLOCK;
    for (int i = threadIdx.x; i < quad_model_f4(NS); i += BLOCK) qm[i] = gqm[i];
    __syncthreads();
    const int gt = blockIdx.x * BLOCK + threadIdx.x;
    const int env = gt >> 2, lane = gt & 3;
    const bool valid = env < N;
    const int e = valid ? env : N - 1;
    QLane<NS, HF, 0> L = make_qlane<NS, HF, 0>(qm, hf, park, BLOCK, lane);
    attach_env_params(L, B, e, nd, DR);                      // the per-env friction (terrain buckets, anymal_terrain.py:238-247) is always honoured
    RootState rs; load_root((const float *)B.p[B2G_T_ROOT_STATE] + 13 * (size_t)e, rs);
    const float2 *dofs = (const float2 *)B.p[B2G_T_DOF_STATE] + (size_t)e * nd;
    float *act_out = (float *)B.p[B2G_T_ACTIONS] + (size_t)e * nd;
    float *torq = (float *)B.p[B2G_T_TORQUES] + (size_t)e * nd;
    const float *last_a = (const float *)B.p[B2G_T_LAST_ACTIONS] + (size_t)e * nd;
    const float *last_v = (const float *)B.p[B2G_T_LAST_DOF_VEL] + (size_t)e * nd;
    int dofi[NS];
    float a_cl[NS], tq[NS] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int s = 0; s < NS; s++) {
        dofi[s] = q_f2i(L.LK(s, 16).w);
        const float2 v = dofs[dofi[s]];
        a_cl[s] = fminf(fmaxf(actions_in[(size_t)e * nd + dofi[s]], -P.clip_actions), P.clip_actions);
        if (valid) act_out[dofi[s]] = a_cl[s];
        L.q[s] = v.x; L.qd[s] = v.y; L.act[s] = 0.f;
    }
    QOutputs o;
    o.write = valid; o.sensor = nullptr; o.dof_force = nullptr;
    o.net_contact = (float *)B.p[B2G_T_NET_CONTACT] + (size_t)e * (q_f2i(qm[7].w) >> 8) * 3;
    const int total = (P.decimation + P.control_freq_inv) * substeps;
    const int pd_until = P.decimation * substeps;
#pragma unroll 1
    for (int k = 0; k < total; k++) {
        if (k < pd_until && (k % substeps) == 0) {
#pragma unroll
            for (int s = 0; s < NS; s++) {
                float t = P.kp * (P.action_scale * a_cl[s] + P.default_dof_pos[dofi[s]] - L.q[s]) - P.kd * L.qd[s];
                t = fminf(fmaxf(t, -P.torque_limit), P.torque_limit);
                L.act[s] = t; tq[s] = t;
            }
        }
        L.substep(rs, k == total - 1, o);
    }

    // ---- post_physics_step (:453-475)
    long long *progress_b = (long long *)B.p[B2G_T_PROGRESS];
    long long *reset_b = (long long *)B.p[B2G_T_RESET];
    const long long progress = progress_b[e] + 1;
    const uint32_t gid = (uint32_t)(e + P.env_id_offset);
    if (P.push_robots && P.push_interval > 0 && (step_counter % (unsigned)P.push_interval) == 0) {   // push_robots :437-439
        rs.rv[0] = t_rand_float(-1.f, 1.f, anymal_uniform(P.seed, gid, step_counter, TAG_PUSH, 0));
        rs.rv[1] = t_rand_float(-1.f, 1.f, anymal_uniform(P.seed, gid, step_counter, TAG_PUSH, 1));
    }
    float2 *dw = (float2 *)B.p[B2G_T_DOF_STATE] + (size_t)e * nd;
    float s_torque = 0.f, s_jacc = 0.f, s_arate = 0.f, s_hip = 0.f;
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int d = dofi[s];
        if (valid) { dw[d] = make_float2(L.q[s], L.qd[s]); if (total > 0 && pd_until > 0) torq[d] = tq[s]; }
        const float t = (total > 0 && pd_until > 0) ? tq[s] : torq[d], a = a_cl[s];
        s_torque += t * t;
        const float dv = last_v[d] - L.qd[s]; s_jacc += dv * dv;
        const float da = last_a[d] - a; s_arate += da * da;
        if (d % 3 == 0) s_hip += fabsf(L.q[s] - P.default_dof_pos[d]);          // dof_pos[:, [0,3,6,9]]
    }
    if (valid && lane == 0) store_root((float *)B.p[B2G_T_ROOT_STATE] + 13 * (size_t)e, rs);
    s_torque = lane_sum<4>(s_torque); s_jacc = lane_sum<4>(s_jacc); s_arate = lane_sum<4>(s_arate); s_hip = lane_sum<4>(s_hip);

    // contact-force terms: every lane looks at bodies base / knee[lane] / foot[lane]
    const float *cf = o.net_contact;
    float *fat_b = (float *)B.p[B2G_T_FEET_AIR_TIME] + (size_t)e * 4;
    float n_knee = 0.f, n_stumble = 0.f, air = 0.f;
    bool knee_hit = false;
    __syncwarp();
    {
        const int k = lane;
        const float *fk = cf + 3 * P.knee_bodies[k], *ff = cf + 3 * P.feet_bodies[k];
        const bool kc = sqrtf(fk[0] * fk[0] + fk[1] * fk[1] + fk[2] * fk[2]) > 1.f;
        knee_hit = kc;
        n_knee += kc ? 1.f : 0.f;
        n_stumble += ((sqrtf(ff[0] * ff[0] + ff[1] * ff[1]) > 5.f) && (fabsf(ff[2]) < 1.f)) ? 1.f : 0.f;
        const bool contact = ff[2] > 1.f;
        float fat = fat_b[k];
        const bool first = (fat > 0.f) && contact;
        fat += P.dt;
        air += (fat - 0.5f) * (first ? 1.f : 0.f);
        fat = contact ? 0.f : fat;
        if (valid) fat_b[k] = fat;
    }
    n_knee = lane_sum<4>(n_knee); n_stumble = lane_sum<4>(n_stumble); air = lane_sum<4>(air);
    const float any_knee = lane_sum<4>(knee_hit ? 1.f : 0.f);

    // prepare quantities (:464-471)
    float *cmd = (float *)B.p[B2G_T_COMMANDS] + (size_t)e * 4;
    const float gvec[3] = {0.f, 0.f, -1.f}, fvec[3] = {1.f, 0.f, 0.f};
    float blv[3], bav[3], pg[3], fwd[3];
    t_quat_rotate(rs.rq, rs.rv, blv, -1.f);
    t_quat_rotate(rs.rq, rs.rw, bav, -1.f);
    t_quat_rotate(rs.rq, gvec, pg, -1.f);
    t_quat_apply(rs.rq, fvec, fwd);
    const float heading = atan2f(fwd[1], fwd[0]);
    const float c0 = cmd[0], c1 = cmd[1], c3 = cmd[3];
    const float c2 = fminf(fmaxf(0.5f * t_wrap_to_pi(c3 - heading), -1.f), 1.f);

    // check_termination (:294-300)
    const float *fb = cf + 3 * P.base_body;
    bool reset = sqrtf(fb[0] * fb[0] + fb[1] * fb[1] + fb[2] * fb[2]) > 1.f;
    if (!P.allow_knee_contacts) reset = reset || (any_knee > 0.f);
    if (progress >= (long long)P.max_episode_length - 1) reset = true;

    float part = 0.f;
    if (lane == 0 && valid) {
        // compute_reward (:315-382)
        const float *R = P.rew_scales;
        const float ex = c0 - blv[0], ey = c1 - blv[1];
        const float lin_err = ex * ex + ey * ey;
        const float ang_err = (c2 - bav[2]) * (c2 - bav[2]);
        float t[13];
        t[0] = expf(-lin_err / 0.25f) * R[1];                    // lin_vel_xy
        t[1] = blv[2] * blv[2] * R[2];                           // lin_vel_z
        t[2] = expf(-ang_err / 0.25f) * R[3];                    // ang_vel_z
        t[3] = (bav[0] * bav[0] + bav[1] * bav[1]) * R[4];       // ang_vel_xy
        t[4] = (pg[0] * pg[0] + pg[1] * pg[1]) * R[5];           // orient
        t[5] = s_torque * R[6];                                  // torques
        t[6] = s_jacc * R[7];                                    // joint_acc
        t[7] = (rs.rp[2] - 0.52f) * (rs.rp[2] - 0.52f) * R[8];   // base_height
        t[8] = air * R[9] * ((sqrtf(c0 * c0 + c1 * c1) > 0.1f) ? 1.f : 0.f);   // air_time
        t[9] = n_knee * R[10];                                   // collision
        t[10] = n_stumble * R[11];                               // stumble
        t[11] = s_arate * R[12];                                 // action_rate
        t[12] = s_hip * R[13];                                   // hip
        float rew = t[0] + t[2] + t[1] + t[3] + t[4] + t[7] + t[5] + t[6] + t[9] + t[11] + t[8] + t[12] + t[10];
        rew = fmaxf(rew, 0.f);
        const uint8_t *to = (const uint8_t *)B.p[B2G_T_TIMEOUT];
        rew += R[0] * (reset ? 1.f : 0.f) * ((to && to[e]) ? 0.f : 1.f);
        ((float *)B.p[B2G_T_REW])[e] = rew;
        float *es = (float *)B.p[B2G_T_EPISODE_SUMS];
#pragma unroll
        for (int k = 0; k < 13; k++) es[(size_t)k * N + e] += t[k];
        reset_b[e] = reset ? 1 : 0;
        progress_b[e] = progress;
        cmd[2] = c2;
        float *bs = (float *)B.p[B2G_T_BASE_SCRATCH] + (size_t)e * 12;
        bs[0] = blv[0]; bs[1] = blv[1]; bs[2] = blv[2]; bs[3] = bav[0]; bs[4] = bav[1]; bs[5] = bav[2];
        bs[6] = pg[0]; bs[7] = pg[1]; bs[8] = pg[2];
        if (reset) part = c0 * c0 + c1 * c1;
    }
    // deterministic per-block partial of sum over the reset set of |commands_xy|^2
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int w = 0; w < BLOCK / 32; w++) tot += s_part[w];
        float *red = (float *)B.p[B2G_T_REDUCE_SCRATCH];
        red[blockIdx.x] = tot;
        if (blockIdx.x == 0) for (int k = 0; k < 16; k++) red[REDUCE_PARTIALS + k] = 0.f;
    }
}

}  // namespace b2g
