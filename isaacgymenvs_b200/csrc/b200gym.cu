// b200gym.cu -- kernels + C ABI (include/b200gym.h) of the B200-native environment stepper.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>

#include "../../include/b200gym.h"
#include "b2g_device.cuh"
#include "b2g_tasks.cuh"

using namespace b2g;

// ============================================================================================
// kernels
// ============================================================================================
struct Buffers {
    void *p[B2G_T_COUNT];
};

constexpr int BLOCK = 128;
#ifndef B2G_MINBLOCKS
#define B2G_MINBLOCKS 1
#endif

// shared-memory copy of the model: only the words this model uses are moved
__device__ __forceinline__ void load_model(DevModel *sm, const DevModel *__restrict__ gm) {
    const int nl = gm->nl, ncp = gm->ncp;
    const uint32_t *src = reinterpret_cast<const uint32_t *>(gm);
    uint32_t *dst = reinterpret_cast<uint32_t *>(sm);
    const int head = (int)(offsetof(DevModel, links) / 4);
    for (int i = threadIdx.x; i < head; i += blockDim.x) dst[i] = src[i];
    const int nlw = nl * (int)(sizeof(LinkC) / 4);
    for (int i = threadIdx.x; i < nlw; i += blockDim.x) dst[head + i] = src[head + i];
    const int cph = (int)(offsetof(DevModel, cps) / 4), ncw = ncp * (int)(sizeof(CpC) / 4);
    for (int i = threadIdx.x; i < ncw; i += blockDim.x) dst[cph + i] = src[cph + i];
    __syncthreads();
}

template <class Topo>
__device__ __forceinline__ void load_env(const DevModel &sm, const Buffers &B, int e, int lane, EnvState<Topo> &st) {
    const float *r = (const float *)B.p[B2G_T_ROOT_STATE] + 13 * (size_t)e;
    st.rp[0] = r[0]; st.rp[1] = r[1]; st.rp[2] = r[2];
    st.rq[0] = r[3]; st.rq[1] = r[4]; st.rq[2] = r[5]; st.rq[3] = r[6];
    st.rv[0] = r[7]; st.rv[1] = r[8]; st.rv[2] = r[9];
    st.rw[0] = r[10]; st.rw[1] = r[11]; st.rw[2] = r[12];
    const int nd = sm.nl - 1;
    const float2 *d = (const float2 *)B.p[B2G_T_DOF_STATE] + (size_t)e * nd;
#pragma unroll
    for (int s = 0; s < Topo::NS; s++) {
        const float2 v = d[sm.slot_link[s][lane] - 1];
        st.q[s] = v.x; st.qd[s] = v.y;
    }
}

template <class Topo>
__device__ __forceinline__ void store_env(const DevModel &sm, const Buffers &B, int e, int lane, const EnvState<Topo> &st) {
    const int nd = sm.nl - 1;
    float2 *d = (float2 *)B.p[B2G_T_DOF_STATE] + (size_t)e * nd;
#pragma unroll
    for (int s = 0; s < Topo::NS; s++) d[sm.slot_link[s][lane] - 1] = make_float2(st.q[s], st.qd[s]);
    if (lane == 0 && !sm.root_fixed) {
        float *r = (float *)B.p[B2G_T_ROOT_STATE] + 13 * (size_t)e;
        r[0] = st.rp[0]; r[1] = st.rp[1]; r[2] = st.rp[2];
        r[3] = st.rq[0]; r[4] = st.rq[1]; r[5] = st.rq[2]; r[6] = st.rq[3];
        r[7] = st.rv[0]; r[8] = st.rv[1]; r[9] = st.rv[2];
        r[10] = st.rw[0]; r[11] = st.rw[1]; r[12] = st.rw[2];
    }
}

// force sensors / joint forces / net contact forces of the last sub-step -> bound output tensors.
// sens[s][6] receives the sensor reading of slot s (body frame) for the observation.
template <class Topo>
__device__ __forceinline__ void store_outputs(const DevModel &sm, const Buffers &B, int e, int lane, bool valid,
                                              const StepOut<Topo> &out, float sens[][6]) {
    constexpr int NS = Topo::NS;
    const int nd = sm.nl - 1;
    float *fs = (float *)B.p[B2G_T_FORCE_SENSOR];
    float *df = (float *)B.p[B2G_T_DOF_FORCE];
    float *nc = (float *)B.p[B2G_T_NET_CONTACT];
#pragma unroll
    for (int s = -1; s < NS; s++) {
        const int i = (s < 0) ? NS : s;
        if (s < 0 && lane != 0) continue;
        const int link = (s < 0) ? 0 : sm.slot_link[s][lane];
        const LinkC &lk = sm.links[link];
        if (lk.sensor >= 0) {
            // body frame = link frame axes for every sensor body of the five assets (sensor pose
            // identity, ant.py:176-178); torque is taken about the body origin = link origin + R*body_pos
            float Fb[3], Tb[3], T[3] = {out.cfT[i][0], out.cfT[i][1], out.cfT[i][2]};
            const int b = sm.sensor_body[lk.sensor];
            float bp[3] = {sm.body_pos[b][0], sm.body_pos[b][1], sm.body_pos[b][2]}, wb[3], bxF[3];
            matvec(out.R[i], bp, wb); cross(wb, out.cfF[i], bxF);
            T[0] -= bxF[0]; T[1] -= bxF[1]; T[2] -= bxF[2];
            matTvec(out.R[i], out.cfF[i], Fb); matTvec(out.R[i], T, Tb);
            if (s >= 0) { sens[s][0] = Fb[0]; sens[s][1] = Fb[1]; sens[s][2] = Fb[2]; sens[s][3] = Tb[0]; sens[s][4] = Tb[1]; sens[s][5] = Tb[2]; }
            if (fs && valid) {
                float *o = fs + ((size_t)e * sm.nsens + lk.sensor) * 6;
                o[0] = Fb[0]; o[1] = Fb[1]; o[2] = Fb[2]; o[3] = Tb[0]; o[4] = Tb[1]; o[5] = Tb[2];
            }
        }
        if (nc && valid && sm.link_body[link] >= 0) {
            float *o = nc + ((size_t)e * sm.nb + sm.link_body[link]) * 3;
            o[0] = out.cfF[i][0]; o[1] = out.cfF[i][1]; o[2] = out.cfF[i][2];
        }
        if (s >= 0 && df && valid) df[(size_t)e * nd + link - 1] = out.dof_force[s];
    }
}

// -------------------------------------------------------------------------------------------
// gym.simulate(): physics only
template <class Topo>
__global__ void __launch_bounds__(BLOCK) simulate_kernel(const DevModel *__restrict__ gm, const int16_t *__restrict__ hf,
                                                         Buffers B, int N) {
    __shared__ DevModel sm;
    load_model(&sm, gm);
    constexpr int L = Topo::L, NS = Topo::NS;
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int env = gt / L, lane = gt % L;
    const bool valid = env < N;
    const int e = valid ? env : N - 1;
    const int nd = sm.nl - 1;
    EnvState<Topo> st;
    load_env<Topo>(sm, B, e, lane, st);
    const float *act = (const float *)B.p[B2G_T_DOF_ACTUATION];
    const float *tgt = (const float *)B.p[B2G_T_DOF_TARGET];
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int link = sm.slot_link[s][lane];
        const float *src = (sm.links[link].drive_mode == 1) ? tgt : act;
        st.act[s] = src ? src[(size_t)e * nd + link - 1] : 0.f;
    }
    Ground gr{&sm, hf};
    StepOut<Topo> out;
    for (int k = 0; k < sm.substeps; k++) substep<Topo>(&sm, gr, lane, st, out, k == sm.substeps - 1);
    float sens[NS][6];
    store_outputs<Topo>(sm, B, e, lane, valid, out, sens);
    if (valid) store_env<Topo>(sm, B, e, lane, st);
}

// -------------------------------------------------------------------------------------------
// One whole VecTask.step() of Ant / Humanoid (vec_task.py:360-408 + ant.py:281-297 / humanoid.py)
template <class Topo, bool HUM>
__global__ void __launch_bounds__(BLOCK, B2G_MINBLOCKS) loco_step_kernel(const DevModel *__restrict__ gm, const int16_t *__restrict__ hf,
                                                          Buffers B, const __grid_constant__ b2g_task_params P,
                                                          const float *__restrict__ actions_in, int N) {
    __shared__ DevModel sm;
    load_model(&sm, gm);
    constexpr int L = Topo::L, NS = Topo::NS;
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int env = gt / L, lane = gt % L;
    const bool valid = env < N;
    const int e = valid ? env : N - 1;
    const int nd = sm.nl - 1;

    EnvState<Topo> st;
    load_env<Topo>(sm, B, e, lane, st);

    // ---- VecTask.step :374 clamp ; pre_physics_step (ant.py:281-285 / humanoid.py:281-285)
    float a[NS];
    int dof[NS];
    float *act_out = (float *)B.p[B2G_T_ACTIONS];
#pragma unroll
    for (int s = 0; s < NS; s++) {
        dof[s] = sm.slot_link[s][lane] - 1;
        const float v = actions_in[(size_t)e * nd + dof[s]];
        a[s] = fminf(fmaxf(v, -P.clip_actions), P.clip_actions);
        if (valid && act_out) act_out[(size_t)e * nd + dof[s]] = a[s];
        st.act[s] = HUM ? (a[s] * P.motor_efforts[dof[s]] * P.power_scale) : (a[s] * P.joint_gears[dof[s]] * P.power_scale);
    }

    // ---- control_freq_inv x gym.simulate (vec_task.py:379-382)
    Ground gr{&sm, hf};
    StepOut<Topo> out;
    const int total = P.control_freq_inv * sm.substeps;
    for (int k = 0; k < total; k++) substep<Topo>(&sm, gr, lane, st, out, k == total - 1);

    float sens[NS][6];
#pragma unroll
    for (int s = 0; s < NS; s++)
#pragma unroll
        for (int c = 0; c < 6; c++) sens[s][c] = 0.f;
    if (total > 0) {
        store_outputs<Topo>(sm, B, e, lane, valid, out, sens);
    } else {
        // control_freq_inv == 0: no gym.simulate this step -- the observation reads the sensor / joint
        // force tensors as they stand (what refresh_*_tensor would return); used to pin the
        // observation/reward arithmetic against the reference's golden vectors
        const float *fs = (const float *)B.p[B2G_T_FORCE_SENSOR];
        const float *df = (const float *)B.p[B2G_T_DOF_FORCE];
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const int link = sm.slot_link[s][lane], sk = sm.links[link].sensor;
            if (sk >= 0 && fs)
#pragma unroll
                for (int c = 0; c < 6; c++) sens[s][c] = fs[((size_t)e * sm.nsens + sk) * 6 + c];
            out.dof_force[s] = df ? df[(size_t)e * nd + link - 1] : 0.f;
        }
    }

    // ---- post_physics_step (ant.py:287-297): progress, reset_idx, observations, reward
    long long *progress_b = (long long *)B.p[B2G_T_PROGRESS];
    long long *reset_b = (long long *)B.p[B2G_T_RESET];
    float *pot_b = (float *)B.p[B2G_T_POTENTIALS], *ppot_b = (float *)B.p[B2G_T_PREV_POTENTIALS];
    long long progress = progress_b[e] + 1;
    float potentials = pot_b[e];
    if (reset_b[e] != 0) {
        // reset_idx (ant.py:252-279 / humanoid.py:253-279)
        int *rc = (int *)B.p[B2G_T_RESET_COUNT];
        const uint32_t count = (uint32_t)rc[e];
        const uint32_t gid = (uint32_t)(e + P.env_id_offset);
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const float up = reset_uniform(P.seed, gid, count, dof[s]);
            const float uv = reset_uniform(P.seed, gid, count, nd + dof[s]);
            const float pos = (P.reset_pos_noise - (-P.reset_pos_noise)) * up + (-P.reset_pos_noise);
            const float vel = (P.reset_vel_noise - (-P.reset_vel_noise)) * uv + (-P.reset_vel_noise);
            st.q[s] = fmaxf(fminf(P.initial_dof_pos[dof[s]] + pos, P.dof_limits_upper[dof[s]]), P.dof_limits_lower[dof[s]]);
            st.qd[s] = vel;
        }
        const float *ir = (const float *)B.p[B2G_T_INITIAL_ROOT] + 13 * (size_t)e;
        st.rp[0] = ir[0]; st.rp[1] = ir[1]; st.rp[2] = ir[2];
        st.rq[0] = ir[3]; st.rq[1] = ir[4]; st.rq[2] = ir[5]; st.rq[3] = ir[6];
        st.rv[0] = ir[7]; st.rv[1] = ir[8]; st.rv[2] = ir[9];
        st.rw[0] = ir[10]; st.rw[1] = ir[11]; st.rw[2] = ir[12];
        potentials = t_potential(P.target[0] - st.rp[0], P.target[1] - st.rp[1], P.dt);
        progress = 0;
        if (valid && lane == 0) rc[e] = (int)(count + 1);
    }
    if (valid) store_env<Topo>(sm, B, e, lane, st);

    // compute_observations
    LocoRootObs ro;
    loco_root_obs(P, st.rp, st.rq, st.rv, st.rw, HUM, ro);
    const float prev_potentials = potentials;     // prev_potentials_new = potentials.clone(), ant.py:390
    potentials = ro.potentials;
    float *obs = (float *)B.p[B2G_T_OBS] + (size_t)e * P.num_obs;
    float *obsc = (float *)B.p[B2G_T_OBS_CLIPPED];
    obsc = (obsc && obsc != (float *)B.p[B2G_T_OBS]) ? obsc + (size_t)e * P.num_obs : nullptr;
    const float clipo = P.clip_obs;
    auto put = [&](int idx, float v) {
        if (!valid) return;
        obs[idx] = v;
        if (obsc) obsc[idx] = fminf(fmaxf(v, -clipo), clipo);
    };
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 12; c++) put(c, ro.o[c]);
    }
    // layout: ant.py:401-406  [12 | nd pos | nd vel | 24 sensors | nd actions]
    //    humanoid.py:407-411  [12 | nd pos | nd vel | nd dof_force | 12 sensors | nd actions]
    const int o_pos = 12, o_vel = 12 + nd, o_frc = 12 + 2 * nd;
    const int o_sens = HUM ? 12 + 3 * nd : 12 + 2 * nd;
    const int o_act = o_sens + 6 * sm.nsens;
    float actions_cost = 0.f, electricity = 0.f, at_limit = 0.f;
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int d = dof[s];
        const float ps = t_unscale(st.q[s], P.dof_limits_lower[d], P.dof_limits_upper[d]);
        const float vs = st.qd[s] * P.dof_vel_scale;
        put(o_pos + d, ps); put(o_vel + d, vs); put(o_act + d, a[s]);
        if (HUM) put(o_frc + d, out.dof_force[s] * P.contact_force_scale);
        const int sk = sm.links[sm.slot_link[s][lane]].sensor;
        if (sk >= 0) {
#pragma unroll
            for (int c = 0; c < 6; c++) put(o_sens + 6 * sk + c, sens[s][c] * P.contact_force_scale);
        }
        // compute_ant_reward (ant.py:353-355) / compute_humanoid_reward (humanoid.py:352-359)
        actions_cost += a[s] * a[s];
        if (HUM) {
            const float ratio = P.motor_efforts[d] / P.max_motor_effort;
            const float scaled = P.joints_at_limit_cost_scale * (fabsf(ps) - 0.98f) / 0.02f;
            at_limit += (fabsf(ps) > 0.98f) ? scaled * ratio : 0.f;
            electricity += fabsf(a[s] * vs) * ratio;
        } else {
            at_limit += (ps > 0.99f) ? 1.f : 0.f;
            electricity += fabsf(a[s] * vs);
        }
    }
    actions_cost = lane_sum<L>(actions_cost);
    electricity = lane_sum<L>(electricity);
    at_limit = lane_sum<L>(at_limit);

    if (valid && lane == 0) {
        const float heading_proj = ro.o[11], up_proj = ro.o[10], height = ro.o[0];
        const float heading_reward = (heading_proj > 0.8f) ? P.heading_weight : P.heading_weight * heading_proj / 0.8f;
        const float up_reward = (up_proj > 0.93f) ? P.up_weight : 0.f;
        const float progress_reward = potentials - prev_potentials;
        float total = progress_reward + P.alive_reward + up_reward + heading_reward - P.actions_cost_scale * actions_cost -
                      P.energy_cost_scale * electricity - (HUM ? at_limit : at_limit * P.joints_at_limit_cost_scale);
        long long reset = 0;                       // reset_buf was cleared by reset_idx or was already 0
        if (height < P.termination_height) { total = P.death_cost; reset = 1; }
        if ((float)progress >= P.max_episode_length - 1.f) reset = 1;
        ((float *)B.p[B2G_T_REW])[e] = total;
        reset_b[e] = reset;
        progress_b[e] = progress;
        pot_b[e] = potentials; ppot_b[e] = prev_potentials;
        float *uv = (float *)B.p[B2G_T_UP_VEC], *hv = (float *)B.p[B2G_T_HEADING_VEC];
        if (uv) { uv[3 * e] = ro.up_vec[0]; uv[3 * e + 1] = ro.up_vec[1]; uv[3 * e + 2] = ro.up_vec[2]; }
        if (hv) { hv[3 * e] = ro.heading_vec[0]; hv[3 * e + 1] = ro.heading_vec[1]; hv[3 * e + 2] = ro.heading_vec[2]; }
        // vec_task.py:394
        uint8_t *to = (uint8_t *)B.p[B2G_T_TIMEOUT];
        if (to) to[e] = (uint8_t)(((float)progress >= P.max_episode_length - 1.f) && reset != 0);
    }
}

// -------------------------------------------------------------------------------------------
// One whole VecTask.step() of Cartpole (cartpole.py:131-163)
__global__ void __launch_bounds__(BLOCK) cartpole_step_kernel(const DevModel *__restrict__ gm, Buffers B,
                                                              const __grid_constant__ b2g_task_params P,
                                                              const float *__restrict__ actions_in, int N) {
    using Topo = TopoChain2x1;
    __shared__ DevModel sm;
    load_model(&sm, gm);
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = env < N;
    const int e = valid ? env : N - 1;
    EnvState<Topo> st;
    load_env<Topo>(sm, B, e, 0, st);
    const float a = fminf(fmaxf(actions_in[e], -P.clip_actions), P.clip_actions);
    st.act[0] = a * P.max_push_effort;   // cartpole.py:159-163: effort on DOF 0 only
    st.act[1] = 0.f;
    Ground gr{&sm, nullptr};
    StepOut<Topo> out;
    const int total = P.control_freq_inv * sm.substeps;
    for (int k = 0; k < total; k++) substep<Topo>(&sm, gr, 0, st, out, false);
    long long *progress_b = (long long *)B.p[B2G_T_PROGRESS];
    long long *reset_b = (long long *)B.p[B2G_T_RESET];
    long long progress = progress_b[e] + 1;
    if (reset_b[e] != 0) {   // reset_idx, cartpole.py:144-157
        int *rc = (int *)B.p[B2G_T_RESET_COUNT];
        const uint32_t count = (uint32_t)rc[e], gid = (uint32_t)(e + P.env_id_offset);
#pragma unroll
        for (int s = 0; s < 2; s++) {
            st.q[s] = 0.2f * (reset_uniform(P.seed, gid, count, s) - 0.5f);
            st.qd[s] = 0.5f * (reset_uniform(P.seed, gid, count, 2 + s) - 0.5f);
        }
        progress = 0;
        if (valid) rc[e] = (int)(count + 1);
    }
    if (!valid) return;
    store_env<Topo>(sm, B, e, 0, st);
    float *act_out = (float *)B.p[B2G_T_ACTIONS];
    if (act_out) act_out[e] = a;
    // compute_observations, cartpole.py:131-142
    const float o[4] = {st.q[0], st.qd[0], st.q[1], st.qd[1]};
    float *obs = (float *)B.p[B2G_T_OBS] + 4 * (size_t)e;
    float *obsc = (float *)B.p[B2G_T_OBS_CLIPPED];
    obsc = (obsc && obsc != (float *)B.p[B2G_T_OBS]) ? obsc + 4 * (size_t)e : nullptr;
#pragma unroll
    for (int c = 0; c < 4; c++) { obs[c] = o[c]; if (obsc) obsc[c] = fminf(fmaxf(o[c], -P.clip_obs), P.clip_obs); }
    float rew; long long reset = 0;
    cartpole_reward(st.q[1], st.qd[1], st.qd[0], st.q[0], P.reset_dist, progress, P.max_episode_length, rew, reset);
    ((float *)B.p[B2G_T_REW])[e] = rew;
    reset_b[e] = reset; progress_b[e] = progress;
    uint8_t *to = (uint8_t *)B.p[B2G_T_TIMEOUT];
    if (to) to[e] = (uint8_t)(((float)progress >= P.max_episode_length - 1.f) && reset != 0);
}

// -------------------------------------------------------------------------------------------
// gym.refresh_rigid_body_state_tensor(): forward kinematics, one thread per env, any topology
__global__ void __launch_bounds__(BLOCK) body_state_kernel(const DevModel *__restrict__ gm, Buffers B, int N) {
    __shared__ DevModel sm;
    load_model(&sm, gm);
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    const int nl = sm.nl, nd = nl - 1;
    float R[MAX_LINKS][9], x[MAX_LINKS][3], wv[MAX_LINKS][3], lv[MAX_LINKS][3];
    const float *r = (const float *)B.p[B2G_T_ROOT_STATE] + 13 * (size_t)e;
    const float q0[4] = {r[3], r[4], r[5], r[6]};
    quat_to_mat(q0, R[0]);
    for (int c = 0; c < 3; c++) { x[0][c] = r[c]; lv[0][c] = sm.root_fixed ? 0.f : r[7 + c]; wv[0][c] = sm.root_fixed ? 0.f : r[10 + c]; }
    const float2 *d = (const float2 *)B.p[B2G_T_DOF_STATE] + (size_t)e * nd;
    for (int i = 1; i < nl; i++) {
        const LinkC &lk = sm.links[i];
        const int p = sm.link_parent[i];
        const float2 qv = d[i - 1];
        float Rt[9], ax[3] = {lk.axis[0], lk.axis[1], lk.axis[2]}, w[3], lp[3] = {lk.lpos[0], lk.lpos[1], lk.lpos[2]}, dd[3], wxd[3];
        matmul(R[p], lk.R0, Rt); matvec(Rt, ax, w);
        if (lk.jtype == 0) {
            float sn, cs; sincosf(qv.x, &sn, &cs);
            const float oc = 1.f - cs;
            for (int j = 0; j < 3; j++) {
                float col[3] = {Rt[j], Rt[3 + j], Rt[6 + j]}, wxc[3];
                cross(w, col, wxc);
                const float wd = dot3(w, col) * oc;
                R[i][j] = col[0] * cs + wxc[0] * sn + w[0] * wd;
                R[i][3 + j] = col[1] * cs + wxc[1] * sn + w[1] * wd;
                R[i][6 + j] = col[2] * cs + wxc[2] * sn + w[2] * wd;
            }
            matvec(R[p], lp, dd);
        } else {
            for (int c = 0; c < 9; c++) R[i][c] = Rt[c];
            matvec(R[p], lp, dd);
            for (int c = 0; c < 3; c++) dd[c] += w[c] * qv.x;
        }
        cross(wv[p], dd, wxd);
        for (int c = 0; c < 3; c++) {
            x[i][c] = x[p][c] + dd[c];
            lv[i][c] = lv[p][c] + wxd[c] + (lk.jtype == 1 ? w[c] * qv.y : 0.f);
            wv[i][c] = wv[p][c] + (lk.jtype == 0 ? w[c] * qv.y : 0.f);
        }
    }
    float *bs = (float *)B.p[B2G_T_RIGID_BODY_STATE] + 13 * (size_t)e * sm.nb;
    for (int b = 0; b < sm.nb; b++) {
        const int i = sm.body_link[b];
        float bp[3] = {sm.body_pos[b][0], sm.body_pos[b][1], sm.body_pos[b][2]}, wb[3], wxb[3], Rb[9], Rwb[9], q[4];
        matvec(R[i], bp, wb); cross(wv[i], wb, wxb);
        quat_to_mat(sm.body_quat[b], Rb); matmul(R[i], Rb, Rwb); mat_to_quat(Rwb, q);
        float *o = bs + 13 * b;
        for (int c = 0; c < 3; c++) { o[c] = x[i][c] + wb[c]; o[7 + c] = lv[i][c] + wxb[c]; o[10 + c] = wv[i][c]; }
        o[3] = q[0]; o[4] = q[1]; o[5] = q[2]; o[6] = q[3];
    }
}

// ============================================================================================
// host side
// ============================================================================================
enum TopoKind { TK_NONE = 0, TK_CHAIN2x1, TK_CHAIN2x4, TK_CHAIN3x4, TK_ANT1, TK_HUM1 };

struct b2g_sim {
    int device = 0;
    int num_envs = 0;
    TopoKind topo = TK_NONE;
    int lanes = 1;
    DevModel hm;                 // host copy
    DevModel *dm = nullptr;      // device copy
    int16_t *d_hf = nullptr;
    Buffers buf;
    size_t buf_bytes[B2G_T_COUNT];
    b2g_task_params task;
    bool has_task = false;
    float *d_actions_stage = nullptr;    // device staging for b2g_task_step_host
    int64_t launches = 0;
};

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }
#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(B2G_E_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)

extern "C" const char *b2g_last_error(void) { return g_err.c_str(); }
extern "C" int b2g_version(void) { return B2G_VERSION; }
extern "C" int64_t b2g_launch_count(const b2g_sim *sim) { return sim ? sim->launches : 0; }

// does the subtree pattern under the root match `L` identical chains of `ns` links?
static bool match_chains(const b2g_model *m, int L, int ns, int slot_link[MAX_SLOTS][MAX_LANES]) {
    if (m->nl != 1 + L * ns) return false;
    std::vector<int> heads;
    for (int i = 1; i < m->nl; i++) if (m->parent[i] == 0) heads.push_back(i);
    if ((int)heads.size() != L) return false;
    for (int l = 0; l < L; l++) {
        int cur = heads[l];
        for (int s = 0; s < ns; s++) {
            slot_link[s][l] = cur;
            if (s + 1 < ns) {
                int next = -1, cnt = 0;
                for (int i = 1; i < m->nl; i++) if (m->parent[i] == cur) { next = i; cnt++; }
                if (cnt != 1) return false;
                cur = next;
            } else {
                for (int i = 1; i < m->nl; i++) if (m->parent[i] == cur) return false;
            }
        }
    }
    return true;
}
template <class Topo>
static bool match_single_lane(const b2g_model *m, int slot_link[MAX_SLOTS][MAX_LANES]) {
    if (m->nl != Topo::NS + 1) return false;
    for (int s = 0; s < Topo::NS; s++) {
        slot_link[s][0] = s + 1;
        if (m->parent[s + 1] != Topo::ps(s) + 1) return false;
    }
    return true;
}

extern "C" int b2g_create(const b2g_model *m, const b2g_sim_params *sp, int32_t num_envs, int32_t device, b2g_sim **out) {
    if (!m || !sp || !out || num_envs <= 0) return fail(B2G_E_INVALID, "b2g_create: null argument or num_envs <= 0");
    if (m->nl < 1 || m->nl > MAX_LINKS || m->nl - 1 > MAX_SLOTS || m->ncp > MAX_CP || m->nsens > MAX_SENS || m->nb > MAX_LINKS)
        return fail(B2G_E_INVALID, "b2g_create: model exceeds compiled limits (links/contact points/sensors)");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(B2G_E_CUDA, std::string("b2g_create: no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(ce));
    CUDA_TRY(cudaSetDevice(device));
    b2g_sim *s = new b2g_sim();
    s->device = device; s->num_envs = num_envs;
    memset(&s->buf, 0, sizeof(s->buf)); memset(s->buf_bytes, 0, sizeof(s->buf_bytes));
    DevModel &h = s->hm;
    memset(&h, 0, sizeof(h));
    h.nl = m->nl; h.ncp = m->ncp; h.nb = m->nb; h.nsens = m->nsens;
    h.root_fixed = m->root_fixed; h.gravity_on = m->gravity_on; h.substeps = sp->substeps;
    h.h = sp->dt / (float)sp->substeps;
    for (int c = 0; c < 3; c++) h.g[c] = m->gravity_on ? sp->gravity[c] : 0.f;
    h.kn = m->contact_kn; h.cn = m->contact_cn; h.vs2 = m->contact_vs * m->contact_vs;
    // topology
    const char *force1 = getenv("B2G_SINGLE_LANE");
    bool single = force1 && force1[0] == '1';
    if (!single && match_chains(m, 4, 2, h.slot_link)) { s->topo = TK_CHAIN2x4; s->lanes = 4; }
    else if (!single && match_chains(m, 4, 3, h.slot_link)) { s->topo = TK_CHAIN3x4; s->lanes = 4; }
    else if (match_chains(m, 1, 2, h.slot_link)) { s->topo = TK_CHAIN2x1; s->lanes = 1; }
    else if (match_single_lane<TopoAnt1>(m, h.slot_link)) { s->topo = TK_ANT1; s->lanes = 1; }
    else if (match_single_lane<TopoHumanoid1>(m, h.slot_link)) { s->topo = TK_HUM1; s->lanes = 1; }
    else { delete s; return fail(B2G_E_UNSUPPORTED, "b2g_create: no compiled kernel for this articulation topology"); }
    // links
    std::vector<int> order(m->ncp);
    for (int i = 0; i < m->ncp; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return m->cp_link[a] < m->cp_link[b]; });
    for (int i = 0; i < MAX_LINKS; i++) h.link_body[i] = -1;
    for (int b = m->nb - 1; b >= 0; b--) { h.body_link[b] = m->body_link[b]; h.link_body[m->body_link[b]] = b; }
    for (int b = 0; b < m->nb; b++) {
        for (int c = 0; c < 3; c++) h.body_pos[b][c] = m->body_pos[3 * b + c];
        for (int c = 0; c < 4; c++) h.body_quat[b][c] = m->body_quat[4 * b + c];
    }
    for (int i = 0; i < m->nl; i++) h.link_parent[i] = m->parent[i];
    for (int k = 0; k < m->nsens; k++) h.sensor_body[k] = m->sensor_body[k];
    for (int i = 0; i < m->nl; i++) {
        LinkC &l = h.links[i];
        const float *q = m->lquat + 4 * i;
        float x = q[0], y = q[1], z = q[2], w = q[3], n = sqrtf(x * x + y * y + z * z + w * w);
        x /= n; y /= n; z /= n; w /= n;
        float R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                      2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                      2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)};
        memcpy(l.R0, R, sizeof(R));
        for (int c = 0; c < 3; c++) { l.lpos[c] = m->lpos[3 * i + c]; l.axis[c] = m->axis[3 * i + c]; l.com[c] = m->com[3 * i + c]; }
        for (int c = 0; c < 6; c++) l.Ic[c] = m->inertia[6 * i + c];
        l.mass = m->mass[i];
        l.armature = m->armature[i]; l.damping = m->damping[i]; l.stiffness = m->stiffness[i];
        l.lower = m->lower[i]; l.upper = m->upper[i]; l.effort = m->effort[i];
        l.kp = m->kp[i]; l.kd = m->kd[i]; l.limit_k = m->limit_k[i]; l.limit_d = m->limit_d[i];
        l.jtype = m->jtype[i]; l.limited = m->limited[i]; l.drive_mode = m->drive_mode[i];
        l.sensor = -1;
        l.cp_begin = l.cp_end = 0;
    }
    for (int k = 0; k < m->nsens; k++) h.links[m->body_link[m->sensor_body[k]]].sensor = k;
    for (int k = 0; k < m->ncp; k++) {
        int src = order[k];
        CpC &c = h.cps[k];
        for (int j = 0; j < 3; j++) c.pos[j] = m->cp_pos[3 * src + j];
        c.radius = m->cp_radius[src]; c.mu = m->cp_mu[src]; c.body = m->cp_body[src]; c.pad = 0;
        LinkC &l = h.links[m->cp_link[src]];
        if (l.cp_end == 0 && l.cp_begin == 0) l.cp_begin = k;
        l.cp_end = k + 1;
    }
    // height field
    if (sp->hf_samples) {
        h.has_hf = 1; h.hf_nx = sp->hf_nx; h.hf_ny = sp->hf_ny;
        h.hf_scale = sp->hf_horizontal_scale; h.hf_inv_scale = 1.f / sp->hf_horizontal_scale; h.hf_vscale = sp->hf_vertical_scale;
        h.hf_ox = sp->hf_origin_x; h.hf_oy = sp->hf_origin_y;
        size_t bytes = (size_t)sp->hf_nx * sp->hf_ny * sizeof(int16_t);
        CUDA_TRY(cudaMalloc(&s->d_hf, bytes));
        CUDA_TRY(cudaMemcpy(s->d_hf, sp->hf_samples, bytes, cudaMemcpyHostToDevice));
    }
    CUDA_TRY(cudaMalloc(&s->dm, sizeof(DevModel)));
    CUDA_TRY(cudaMemcpy(s->dm, &h, sizeof(DevModel), cudaMemcpyHostToDevice));
    *out = s;
    return B2G_OK;
}

extern "C" int b2g_destroy(b2g_sim *s) {
    if (!s) return B2G_OK;
    cudaSetDevice(s->device);
    if (s->dm) cudaFree(s->dm);
    if (s->d_hf) cudaFree(s->d_hf);
    if (s->d_actions_stage) cudaFree(s->d_actions_stage);
    delete s;
    return B2G_OK;
}

extern "C" int b2g_bind(b2g_sim *s, int32_t slot, void *ptr, size_t bytes) {
    if (!s || slot < 0 || slot >= B2G_T_COUNT) return fail(B2G_E_INVALID, "b2g_bind: bad slot");
    const int N = s->num_envs, nd = s->hm.nl - 1, nb = s->hm.nb, ns = s->hm.nsens;
    size_t need = 0;
    switch (slot) {
        case B2G_T_ROOT_STATE: case B2G_T_INITIAL_ROOT: need = (size_t)N * 13 * 4; break;
        case B2G_T_DOF_STATE: need = (size_t)N * nd * 8; break;
        case B2G_T_DOF_ACTUATION: case B2G_T_DOF_TARGET: case B2G_T_DOF_FORCE: need = (size_t)N * nd * 4; break;
        case B2G_T_RIGID_BODY_STATE: need = (size_t)N * nb * 13 * 4; break;
        case B2G_T_FORCE_SENSOR: need = (size_t)N * ns * 6 * 4; break;
        case B2G_T_NET_CONTACT: need = (size_t)N * nb * 3 * 4; break;
        case B2G_T_REW: case B2G_T_POTENTIALS: case B2G_T_PREV_POTENTIALS: case B2G_T_RESET_COUNT: need = (size_t)N * 4; break;
        case B2G_T_RESET: case B2G_T_PROGRESS: need = (size_t)N * 8; break;
        case B2G_T_TIMEOUT: need = (size_t)N; break;
        case B2G_T_UP_VEC: case B2G_T_HEADING_VEC: need = (size_t)N * 12; break;
        default: need = 0; break;   // ACTIONS / OBS / OBS_CLIPPED are checked against the task in b2g_set_task
    }
    if (ptr && bytes < need) return fail(B2G_E_INVALID, "b2g_bind: buffer smaller than the tensor's layout requires");
    s->buf.p[slot] = ptr; s->buf_bytes[slot] = bytes;
    return B2G_OK;
}

static int require(const b2g_sim *s, std::initializer_list<int> slots, const char *who) {
    for (int k : slots) if (!s->buf.p[k]) return fail(B2G_E_UNBOUND, std::string(who) + ": tensor slot " + std::to_string(k) + " is not bound");
    return B2G_OK;
}

extern "C" int b2g_simulate(b2g_sim *s, void *stream) {
    if (!s) return fail(B2G_E_INVALID, "b2g_simulate: null sim");
    int rc = require(s, {B2G_T_ROOT_STATE, B2G_T_DOF_STATE}, "b2g_simulate"); if (rc) return rc;
    CUDA_TRY(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int N = s->num_envs, grid = (N * s->lanes + BLOCK - 1) / BLOCK;
    switch (s->topo) {
        case TK_CHAIN2x1: simulate_kernel<TopoChain2x1><<<grid, BLOCK, 0, st>>>(s->dm, s->d_hf, s->buf, N); break;
        case TK_CHAIN2x4: simulate_kernel<TopoChain2x4><<<grid, BLOCK, 0, st>>>(s->dm, s->d_hf, s->buf, N); break;
        case TK_CHAIN3x4: simulate_kernel<TopoChain3x4><<<grid, BLOCK, 0, st>>>(s->dm, s->d_hf, s->buf, N); break;
        case TK_ANT1: simulate_kernel<TopoAnt1><<<grid, BLOCK, 0, st>>>(s->dm, s->d_hf, s->buf, N); break;
        case TK_HUM1: simulate_kernel<TopoHumanoid1><<<grid, BLOCK, 0, st>>>(s->dm, s->d_hf, s->buf, N); break;
        default: return fail(B2G_E_UNSUPPORTED, "b2g_simulate: unsupported topology");
    }
    s->launches++;
    CUDA_TRY(cudaGetLastError());
    return B2G_OK;
}

extern "C" int b2g_refresh_rigid_body_state(b2g_sim *s, void *stream) {
    if (!s) return fail(B2G_E_INVALID, "null sim");
    int rc = require(s, {B2G_T_ROOT_STATE, B2G_T_DOF_STATE, B2G_T_RIGID_BODY_STATE}, "b2g_refresh_rigid_body_state"); if (rc) return rc;
    CUDA_TRY(cudaSetDevice(s->device));
    const int N = s->num_envs;
    body_state_kernel<<<(N + BLOCK - 1) / BLOCK, BLOCK, 0, (cudaStream_t)stream>>>(s->dm, s->buf, N);
    s->launches++;
    CUDA_TRY(cudaGetLastError());
    return B2G_OK;
}

extern "C" int b2g_set_task(b2g_sim *s, const b2g_task_params *t) {
    if (!s || !t) return fail(B2G_E_INVALID, "b2g_set_task: null argument");
    const int nd = s->hm.nl - 1;
    if (t->task == B2G_TASK_CARTPOLE) {
        if (s->topo != TK_CHAIN2x1 || t->num_obs != 4 || t->num_actions != 1) return fail(B2G_E_UNSUPPORTED, "cartpole task needs the 2-DOF fixed-base chain, 4 obs, 1 action");
    } else if (t->task == B2G_TASK_ANT) {
        if ((s->topo != TK_CHAIN2x4 && s->topo != TK_ANT1) || t->num_actions != nd || t->num_obs != 12 + 3 * nd + 6 * s->hm.nsens)
            return fail(B2G_E_UNSUPPORTED, "ant task: topology / observation size mismatch");
    } else if (t->task == B2G_TASK_HUMANOID) {
        if (s->topo != TK_HUM1 || t->num_actions != nd || t->num_obs != 12 + 4 * nd + 6 * s->hm.nsens)
            return fail(B2G_E_UNSUPPORTED, "humanoid task: topology / observation size mismatch");
    } else return fail(B2G_E_UNSUPPORTED, "b2g_set_task: unknown task id");
    if (t->control_freq_inv < 0) return fail(B2G_E_INVALID, "control_freq_inv < 0");
    s->task = *t; s->has_task = true;
    return B2G_OK;
}

extern "C" int b2g_task_step(b2g_sim *s, const float *actions, void *stream) {
    if (!s || !actions) return fail(B2G_E_INVALID, "b2g_task_step: null argument");
    if (!s->has_task) return fail(B2G_E_INVALID, "b2g_task_step: call b2g_set_task first");
    int rc = require(s, {B2G_T_ROOT_STATE, B2G_T_DOF_STATE, B2G_T_OBS, B2G_T_REW, B2G_T_RESET, B2G_T_PROGRESS, B2G_T_RESET_COUNT}, "b2g_task_step");
    if (rc) return rc;
    const b2g_task_params &P = s->task;
    const size_t N = s->num_envs;
    if (s->buf_bytes[B2G_T_OBS] < N * P.num_obs * 4) return fail(B2G_E_INVALID, "OBS buffer too small");
    if (s->buf.p[B2G_T_ACTIONS] && s->buf_bytes[B2G_T_ACTIONS] < N * P.num_actions * 4) return fail(B2G_E_INVALID, "ACTIONS buffer too small");
    CUDA_TRY(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = ((int)N * s->lanes + BLOCK - 1) / BLOCK;
    if (P.task == B2G_TASK_CARTPOLE) {
        cartpole_step_kernel<<<grid, BLOCK, 0, st>>>(s->dm, s->buf, P, actions, (int)N);
    } else {
        rc = require(s, {B2G_T_POTENTIALS, B2G_T_PREV_POTENTIALS, B2G_T_INITIAL_ROOT}, "b2g_task_step"); if (rc) return rc;
        if (P.task == B2G_TASK_ANT && s->topo == TK_CHAIN2x4) loco_step_kernel<TopoChain2x4, false><<<grid, BLOCK, 0, st>>>(s->dm, s->d_hf, s->buf, P, actions, (int)N);
        else if (P.task == B2G_TASK_ANT) loco_step_kernel<TopoAnt1, false><<<grid, BLOCK, 0, st>>>(s->dm, s->d_hf, s->buf, P, actions, (int)N);
        else loco_step_kernel<TopoHumanoid1, true><<<grid, BLOCK, 0, st>>>(s->dm, s->d_hf, s->buf, P, actions, (int)N);
    }
    s->launches++;
    CUDA_TRY(cudaGetLastError());
    return B2G_OK;
}

extern "C" int b2g_task_step_host(b2g_sim *s, const float *h_actions, float *h_obs, float *h_rew, int64_t *h_reset,
                                  uint8_t *h_timeout, void *stream) {
    if (!s || !h_actions) return fail(B2G_E_INVALID, "b2g_task_step_host: null argument");
    if (!s->has_task) return fail(B2G_E_INVALID, "b2g_task_step_host: call b2g_set_task first");
    CUDA_TRY(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t N = s->num_envs, abytes = N * s->task.num_actions * 4;
    if (!s->d_actions_stage) CUDA_TRY(cudaMalloc(&s->d_actions_stage, abytes));
    CUDA_TRY(cudaMemcpyAsync(s->d_actions_stage, h_actions, abytes, cudaMemcpyHostToDevice, st));
    int rc = b2g_task_step(s, s->d_actions_stage, stream); if (rc) return rc;
    const void *obs_src = s->buf.p[B2G_T_OBS_CLIPPED] ? s->buf.p[B2G_T_OBS_CLIPPED] : s->buf.p[B2G_T_OBS];
    if (h_obs) CUDA_TRY(cudaMemcpyAsync(h_obs, obs_src, N * s->task.num_obs * 4, cudaMemcpyDeviceToHost, st));
    if (h_rew) CUDA_TRY(cudaMemcpyAsync(h_rew, s->buf.p[B2G_T_REW], N * 4, cudaMemcpyDeviceToHost, st));
    if (h_reset) CUDA_TRY(cudaMemcpyAsync(h_reset, s->buf.p[B2G_T_RESET], N * 8, cudaMemcpyDeviceToHost, st));
    if (h_timeout && s->buf.p[B2G_T_TIMEOUT]) CUDA_TRY(cudaMemcpyAsync(h_timeout, s->buf.p[B2G_T_TIMEOUT], N, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return B2G_OK;
}
