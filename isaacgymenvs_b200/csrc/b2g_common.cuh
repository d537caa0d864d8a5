// b2g_common.cuh -- definitions shared by the kernel headers: the bound-tensor table, the tile descriptor of the
// bulk-copy step kernels, root-state row helpers.
#pragma once
#include "b2g_device.cuh"
#include "../../include/b200gym.h"

using namespace b2g;

struct Buffers {
    void *p[B2G_T_COUNT];
};

extern __shared__ float4 b2g_dyn_smem[];

struct TileArgs {
    int on;          // whole tiles + bulk copies
    int io_f4;       // float4 offset (per block) of the in/out tile region inside dynamic smem
    int model_f4;    // float4 offset of the packed model
    // b2g_task_step_host with PINNED host buffers: the kernel reads its action tile from, and writes its result tiles
    // to, host memory directly (unified addressing, plain coalesced 16-byte loads / stores -- not TMA), so the step
    // needs no separate copy launches.  Null = off.
    const float *h_act;
    float *h_obs, *h_rew;
    long long *h_reset;
    uint8_t *h_timeout;
};

__device__ __forceinline__ void load_root(const float *r, RootState &rs) {
    rs.rp[0] = r[0]; rs.rp[1] = r[1]; rs.rp[2] = r[2];
    rs.rq[0] = r[3]; rs.rq[1] = r[4]; rs.rq[2] = r[5]; rs.rq[3] = r[6];
    rs.rv[0] = r[7]; rs.rv[1] = r[8]; rs.rv[2] = r[9];
    rs.rw[0] = r[10]; rs.rw[1] = r[11]; rs.rw[2] = r[12];
}
__device__ __forceinline__ void load_obj(const float *r, ObjState &ob) {
    ob.p[0] = r[0]; ob.p[1] = r[1]; ob.p[2] = r[2];
    ob.q[0] = r[3]; ob.q[1] = r[4]; ob.q[2] = r[5]; ob.q[3] = r[6];
    ob.v[0] = r[7]; ob.v[1] = r[8]; ob.v[2] = r[9];
    ob.w[0] = r[10]; ob.w[1] = r[11]; ob.w[2] = r[12];
}
__device__ __forceinline__ void store_obj(float *r, const ObjState &ob) {
    r[0] = ob.p[0]; r[1] = ob.p[1]; r[2] = ob.p[2];
    r[3] = ob.q[0]; r[4] = ob.q[1]; r[5] = ob.q[2]; r[6] = ob.q[3];
    r[7] = ob.v[0]; r[8] = ob.v[1]; r[9] = ob.v[2];
    r[10] = ob.w[0]; r[11] = ob.w[1]; r[12] = ob.w[2];
}
__device__ __forceinline__ void store_root(float *r, const RootState &rs) {
    r[0] = rs.rp[0]; r[1] = rs.rp[1]; r[2] = rs.rp[2];
    r[3] = rs.rq[0]; r[4] = rs.rq[1]; r[5] = rs.rq[2]; r[6] = rs.rq[3];
    r[7] = rs.rv[0]; r[8] = rs.rv[1]; r[9] = rs.rv[2];
    r[10] = rs.rw[0]; r[11] = rs.rw[1]; r[12] = rs.rw[2];
}

