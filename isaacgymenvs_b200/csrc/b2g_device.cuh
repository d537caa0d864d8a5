// b2g_device.cuh -- device-side articulated-body sub-step (sm_100a).
//
// Replaces the closed `gym.simulate(sim)` (call sites tasks/base/vec_task.py:379-382,
// tasks/anymal_terrain.py:448 in the reference) with the algorithm BASELINE.json's north_star names:
// reduced-coordinate Featherstone ABA + PD/effort actuators + semi-implicit Euler + penalty contact.
//
// Formulation (differs on purpose from oracle/aba_oracle.c, which is body-coordinate ABA):
//   * all spatial quantities are expressed in WORLD-ALIGNED axes about the root link's origin O,
//     so child->parent accumulation of articulated inertias is a plain sum (no 6x6 congruence
//     transforms) and contact normals need no rotation;
//   * articulated inertia = packed symmetric 6x6: A (ang-ang, 6) | B (ang-lin, 9) | C (lin-lin, 6);
//   * every DOF is a 1-DOF link; joint damping / stiffness / PD gains / limit springs and the
//     contact spring-damper-friction are integrated implicitly by augmenting the joint-space
//     diagonal and the link inertia (DESIGN.md "time stepping").
//
// Work decomposition: an environment is owned by L lanes of a warp (L = 1, 2 or 4).  Each lane
// owns whole sub-trees hanging off the root as a fixed sequence of NS "slots" whose parent slot is
// compile-time (Topo::ps), so all per-slot state lives in registers; the root is replicated on
// the L lanes and the lanes' sub-tree contributions meet in an xor-butterfly (warp shuffles).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef B2G_FAST_TRIG
#define B2G_FAST_TRIG 1
#endif

namespace b2g {

constexpr int MAX_LINKS = 32;
constexpr int MAX_CP = 64;
constexpr int MAX_SENS = 8;
constexpr int MAX_SLOTS = 24;
constexpr int MAX_LANES = 4;

// ---------------------------------------------------------------------------------------------
// model constants (global memory -> shared memory at kernel start; strides are odd so that the
// L lanes of an env, which read different links at the same time, hit different banks)
struct LinkC {
    float R0[9];          // link frame in the parent link frame at q = 0, row-major
    float lpos[3];
    float axis[3];        // joint axis, link frame
    float com[3];
    float Ic[6];          // xx yy zz xy xz yz about the COM, link axes
    float mass;
    float armature, damping, stiffness, lower, upper, effort, kp, kd, limit_k, limit_d;
    int jtype, limited, drive_mode;
    int cp_begin, cp_end; // contact spheres of this link: [begin, end) in the link-sorted cp array
    int sensor;           // force sensor attached to this link's body (-1 none)
};                        // 41 words
static_assert(sizeof(LinkC) == 41 * 4, "LinkC stride");

struct CpC {
    float pos[3];
    float radius, mu;
    int body;
    int pad;
};                        // 7 words
static_assert(sizeof(CpC) == 7 * 4, "CpC stride");

struct DevModel {
    int nl, ncp, nb, nsens;
    int root_fixed, gravity_on, substeps, has_hf;
    float h;              // sub-step length dt / substeps
    float g[3];
    float kn, cn, vs2;    // contact stiffness, damping, (slip regularisation speed)^2
    int hf_nx, hf_ny;
    float hf_inv_scale, hf_scale, hf_vscale, hf_ox, hf_oy;
    int slot_link[MAX_SLOTS][MAX_LANES];   // (slot, lane) -> link index
    int sensor_body[MAX_SENS];
    int body_link[MAX_LINKS];
    int link_body[MAX_LINKS];              // first body riding on the link (-1: massless virtual link)
    int link_parent[MAX_LINKS];
    float body_pos[MAX_LINKS][3];
    float body_quat[MAX_LINKS][4];
    LinkC links[MAX_LINKS];
    CpC cps[MAX_CP];
};

// ---------------------------------------------------------------------------------------------
// compile-time topologies: L lanes per env, NS slots per lane, ps(s) = parent slot (-1 = root)
struct TopoChain2x4 {   // Ant: 4 legs x (hip, ankle)                         nv_ant.xml:47-78
    static constexpr int L = 4, NS = 2;
    __host__ __device__ static constexpr int ps(int s) { return s - 1; }
};
struct TopoChain3x4 {   // ANYmal: 4 legs x (HAA, HFE, KFE)                   anymal_minimal.urdf
    static constexpr int L = 4, NS = 3;
    __host__ __device__ static constexpr int ps(int s) { return s - 1; }
};
struct TopoChain2x1 {   // Cartpole: slider -> cart -> pole                   cartpole.urdf:61-75
    static constexpr int L = 1, NS = 2;
    __host__ __device__ static constexpr int ps(int s) { return s - 1; }
};
struct TopoAnt1 {       // Ant on one lane (comparison / fallback)
    static constexpr int L = 1, NS = 8;
    __host__ __device__ static constexpr int ps(int s) { return (s & 1) ? s - 1 : -1; }
};
struct TopoHumanoid1 {  // Humanoid, 21 1-DOF links on one lane                nv_humanoid.xml:36-136
    static constexpr int L = 1, NS = 21;
    __host__ __device__ static constexpr int ps(int s) {
        constexpr int p[21] = {-1, 0, 1, 2, 3, 4, 5, 6, 7, 2, 9, 10, 11, 12, 13, -1, 15, 16, -1, 18, 19};
        return p[s];
    }
};

// ---------------------------------------------------------------------------------------------
// small vector helpers (all fully inlined, arrays are register-resident after unrolling)
__device__ __forceinline__ void cross(const float a[3], const float b[3], float o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ float dot3(const float a[3], const float b[3]) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
__device__ __forceinline__ void matvec(const float R[9], const float v[3], float o[3]) {
    o[0] = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
    o[1] = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
    o[2] = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
}
__device__ __forceinline__ void matTvec(const float R[9], const float v[3], float o[3]) {
    o[0] = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
    o[1] = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
    o[2] = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
}
__device__ __forceinline__ void matmul(const float A[9], const float B[9], float C[9]) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
__device__ __forceinline__ void quat_to_mat(const float q[4], float R[9]) {
    float x = q[0], y = q[1], z = q[2], w = q[3];
    float inv = rsqrtf(x * x + y * y + z * z + w * w);
    x *= inv; y *= inv; z *= inv; w *= inv;
    R[0] = 1.f - 2.f * (y * y + z * z); R[1] = 2.f * (x * y - z * w); R[2] = 2.f * (x * z + y * w);
    R[3] = 2.f * (x * y + z * w); R[4] = 1.f - 2.f * (x * x + z * z); R[5] = 2.f * (y * z - x * w);
    R[6] = 2.f * (x * z - y * w); R[7] = 2.f * (y * z + x * w); R[8] = 1.f - 2.f * (x * x + y * y);
}
__device__ __forceinline__ void mat_to_quat(const float R[9], float q[4]) {
    float t = R[0] + R[4] + R[8], s;
    if (t > 0.f) { s = sqrtf(t + 1.f) * 2.f; q[3] = 0.25f * s; q[0] = (R[7] - R[5]) / s; q[1] = (R[2] - R[6]) / s; q[2] = (R[3] - R[1]) / s; }
    else if (R[0] > R[4] && R[0] > R[8]) { s = sqrtf(1.f + R[0] - R[4] - R[8]) * 2.f; q[3] = (R[7] - R[5]) / s; q[0] = 0.25f * s; q[1] = (R[1] + R[3]) / s; q[2] = (R[2] + R[6]) / s; }
    else if (R[4] > R[8]) { s = sqrtf(1.f + R[4] - R[0] - R[8]) * 2.f; q[3] = (R[2] - R[6]) / s; q[0] = (R[1] + R[3]) / s; q[1] = 0.25f * s; q[2] = (R[5] + R[7]) / s; }
    else { s = sqrtf(1.f + R[8] - R[0] - R[4]) * 2.f; q[3] = (R[3] - R[1]) / s; q[0] = (R[2] + R[6]) / s; q[1] = (R[5] + R[7]) / s; q[2] = 0.25f * s; }
    if (q[3] < 0.f) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
}
__device__ __forceinline__ void b2g_sincos(float a, float *s, float *c) {
#if B2G_FAST_TRIG
    __sincosf(a, s, c);
#else
    sincosf(a, s, c);
#endif
}

// packed symmetric 6x6:  IA[0..5] = A (xx yy zz xy xz yz), IA[6..14] = B row-major (ang x lin),
// IA[15..20] = C (xx yy zz xy xz yz).   y = IA * (a ; l)
__device__ __forceinline__ void sym6_mul(const float IA[21], const float a[3], const float l[3], float ya[3], float yl[3]) {
    const float *A = IA, *B = IA + 6, *C = IA + 15;
    ya[0] = A[0] * a[0] + A[3] * a[1] + A[4] * a[2] + B[0] * l[0] + B[1] * l[1] + B[2] * l[2];
    ya[1] = A[3] * a[0] + A[1] * a[1] + A[5] * a[2] + B[3] * l[0] + B[4] * l[1] + B[5] * l[2];
    ya[2] = A[4] * a[0] + A[5] * a[1] + A[2] * a[2] + B[6] * l[0] + B[7] * l[1] + B[8] * l[2];
    yl[0] = B[0] * a[0] + B[3] * a[1] + B[6] * a[2] + C[0] * l[0] + C[3] * l[1] + C[4] * l[2];
    yl[1] = B[1] * a[0] + B[4] * a[1] + B[7] * a[2] + C[3] * l[0] + C[1] * l[1] + C[5] * l[2];
    yl[2] = B[2] * a[0] + B[5] * a[1] + B[8] * a[2] + C[4] * l[0] + C[5] * l[1] + C[2] * l[2];
}
// IA += s * (ja; jl)(ja; jl)^T
__device__ __forceinline__ void sym6_rank1(float IA[21], float s, const float ja[3], const float jl[3]) {
    float sa0 = s * ja[0], sa1 = s * ja[1], sa2 = s * ja[2];
    float sl0 = s * jl[0], sl1 = s * jl[1], sl2 = s * jl[2];
    IA[0] += sa0 * ja[0]; IA[1] += sa1 * ja[1]; IA[2] += sa2 * ja[2];
    IA[3] += sa0 * ja[1]; IA[4] += sa0 * ja[2]; IA[5] += sa1 * ja[2];
    IA[6] += sa0 * jl[0]; IA[7] += sa0 * jl[1]; IA[8] += sa0 * jl[2];
    IA[9] += sa1 * jl[0]; IA[10] += sa1 * jl[1]; IA[11] += sa1 * jl[2];
    IA[12] += sa2 * jl[0]; IA[13] += sa2 * jl[1]; IA[14] += sa2 * jl[2];
    IA[15] += sl0 * jl[0]; IA[16] += sl1 * jl[1]; IA[17] += sl2 * jl[2];
    IA[18] += sl0 * jl[1]; IA[19] += sl0 * jl[2]; IA[20] += sl1 * jl[2];
}

// solve the SPD system  IA * (xa; xl) = (ba; bl)  (root of a floating base) by Cholesky
__device__ __forceinline__ void sym6_solve(const float IA[21], const float ba[3], const float bl[3], float xa[3], float xl[3]) {
    float M[6][6];
    const float *A = IA, *B = IA + 6, *C = IA + 15;
    M[0][0] = A[0]; M[1][1] = A[1]; M[2][2] = A[2]; M[1][0] = A[3]; M[2][0] = A[4]; M[2][1] = A[5];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) M[3 + j][i] = B[3 * i + j];   // lower-left block = B^T
    M[3][3] = C[0]; M[4][4] = C[1]; M[5][5] = C[2]; M[4][3] = C[3]; M[5][3] = C[4]; M[5][4] = C[5];
    float Lm[6][6], dinv[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
#pragma unroll
        for (int j = 0; j <= i; j++) {
            float s = M[i][j];
#pragma unroll
            for (int k = 0; k < j; k++) s -= Lm[i][k] * Lm[j][k];
            if (i == j) { dinv[i] = rsqrtf(s); Lm[i][i] = s * dinv[i]; }
            else Lm[i][j] = s * dinv[j];
        }
    }
    float b[6] = {ba[0], ba[1], ba[2], bl[0], bl[1], bl[2]}, y[6], x[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        float s = b[i];
#pragma unroll
        for (int k = 0; k < i; k++) s -= Lm[i][k] * y[k];
        y[i] = s * dinv[i];
    }
#pragma unroll
    for (int i = 5; i >= 0; i--) {
        float s = y[i];
#pragma unroll
        for (int k = i + 1; k < 6; k++) s -= Lm[k][i] * x[k];
        x[i] = s * dinv[i];
    }
    xa[0] = x[0]; xa[1] = x[1]; xa[2] = x[2]; xl[0] = x[3]; xl[1] = x[4]; xl[2] = x[5];
}

// ---------------------------------------------------------------------------------------------
// per-env dynamic state held in registers across the sub-steps of one control step
template <class Topo>
struct EnvState {
    float rp[3];      // root position (world)
    float rq[4];      // root quaternion xyzw
    float rv[3];      // root linear velocity (world, of the root origin)
    float rw[3];      // root angular velocity (world)
    float q[Topo::NS], qd[Topo::NS];
    float act[Topo::NS];   // actuation force (effort mode) or position target (position drive)
};

// outputs of the last sub-step that the tasks read
template <class Topo>
struct StepOut {
    float cfF[Topo::NS + 1][3];   // net contact force on slot's link (index NS = root), world axes
    float cfT[Topo::NS + 1][3];   // net contact torque about the LINK origin, world axes
    float dof_force[Topo::NS];
    float R[Topo::NS + 1][9];     // link orientation at the start of the last sub-step
};

struct Ground {
    const DevModel *m;
    const int16_t *hf;
    // height and unit normal at world (x, y)
    __device__ __forceinline__ void sample(float x, float y, float &h, float n[3]) const {
        if (!m->has_hf) { h = 0.f; n[0] = 0.f; n[1] = 0.f; n[2] = 1.f; return; }
        float fx = (x - m->hf_ox) * m->hf_inv_scale, fy = (y - m->hf_oy) * m->hf_inv_scale;
        int ix = (int)floorf(fx), iy = (int)floorf(fy);
        ix = max(0, min(ix, m->hf_nx - 2)); iy = max(0, min(iy, m->hf_ny - 2));
        float tx = fminf(fmaxf(fx - ix, 0.f), 1.f), ty = fminf(fmaxf(fy - iy, 0.f), 1.f);
        const int16_t *p = hf + (size_t)ix * m->hf_ny + iy;
        float h00 = p[0] * m->hf_vscale, h01 = p[1] * m->hf_vscale;
        float h10 = p[m->hf_ny] * m->hf_vscale, h11 = p[m->hf_ny + 1] * m->hf_vscale;
        float dhx, dhy;
        if (tx + ty <= 1.f) { dhx = h10 - h00; dhy = h01 - h00; h = h00 + tx * dhx + ty * dhy; }
        else { dhx = h11 - h01; dhy = h11 - h10; h = h11 - (1.f - tx) * dhx - (1.f - ty) * dhy; }
        float gx = dhx * m->hf_inv_scale, gy = dhy * m->hf_inv_scale;
        float inv = rsqrtf(gx * gx + gy * gy + 1.f);
        n[0] = -gx * inv; n[1] = -gy * inv; n[2] = inv;
    }
};

// contact spheres of one link against the ground.  Pass A (ACCUM): adds the explicit force to the
// bias (pa, pl) and the implicit term h*J^T G J to IA.  Pass B (!ACCUM, after the accelerations
// are known): accumulates the force actually applied over the sub-step, F = F0 - h*G*(J a).
template <bool ACCUM>
__device__ __forceinline__ void link_contacts(const DevModel *m, const Ground &gr, const LinkC &lk, const float rp[3],
                                              const float R[9], const float x[3], const float vw[3], const float vl[3],
                                              float IA[21], float pa[3], float pl[3],
                                              const float aw[3], const float al[3], float F[3], float T[3],
                                              int cp_first, int cp_step) {
    const float h = m->h;
    for (int k = lk.cp_begin + cp_first; k < lk.cp_end; k += cp_step) {
        const CpC &cp = m->cps[k];
        float pc[3], lp[3] = {cp.pos[0], cp.pos[1], cp.pos[2]};
        matvec(R, lp, pc);
        pc[0] += x[0]; pc[1] += x[1]; pc[2] += x[2];          // sphere centre relative to O
        float hg, n[3];
        gr.sample(rp[0] + pc[0], rp[1] + pc[1], hg, n);
        float d = cp.radius - (rp[2] + pc[2] - hg) * n[2];
        if (d <= 0.f) continue;
        float r[3] = {pc[0] - cp.radius * n[0], pc[1] - cp.radius * n[1], pc[2] - cp.radius * n[2]};
        float wxr[3]; cross(vw, r, wxr);
        float u[3] = {vl[0] + wxr[0], vl[1] + wxr[1], vl[2] + wxr[2]};
        float gn = m->cn + h * m->kn;
        float un = dot3(u, n);
        float Fn = m->kn * d - gn * un;
        if (Fn <= 0.f) continue;
        float ut[3] = {u[0] - un * n[0], u[1] - un * n[1], u[2] - un * n[2]};
        float gam = cp.mu * Fn * rsqrtf(dot3(ut, ut) + m->vs2);
        float F0[3] = {Fn * n[0] - gam * ut[0], Fn * n[1] - gam * ut[1], Fn * n[2] - gam * ut[2]};
        if (ACCUM) {
            float rxF[3]; cross(r, F0, rxF);
            pa[0] -= rxF[0]; pa[1] -= rxF[1]; pa[2] -= rxF[2];
            pl[0] -= F0[0]; pl[1] -= F0[1]; pl[2] -= F0[2];
            // J^T G J with G = gam*1 + (gn-gam) n n^T ; rows of J: j_k = (r x e_k ; e_k)
            float hg_ = h * gam;
            const float jx[3] = {0.f, r[2], -r[1]}, jy[3] = {-r[2], 0.f, r[0]}, jz[3] = {r[1], -r[0], 0.f};
            const float ex[3] = {1.f, 0.f, 0.f}, ey[3] = {0.f, 1.f, 0.f}, ez[3] = {0.f, 0.f, 1.f};
            sym6_rank1(IA, hg_, jx, ex); sym6_rank1(IA, hg_, jy, ey); sym6_rank1(IA, hg_, jz, ez);
            float rxn[3]; cross(r, n, rxn);
            sym6_rank1(IA, h * (gn - gam), rxn, n);
        } else {
            float axr[3]; cross(aw, r, axr);
            float Ja[3] = {al[0] + axr[0], al[1] + axr[1], al[2] + axr[2]};
            float Jan = dot3(Ja, n);
            float Fk[3];
#pragma unroll
            for (int c = 0; c < 3; c++) Fk[c] = F0[c] - h * (gam * Ja[c] + (gn - gam) * Jan * n[c]);
            float rl[3] = {r[0] - x[0], r[1] - x[1], r[2] - x[2]}, t[3];
            cross(rl, Fk, t);
#pragma unroll
            for (int c = 0; c < 3; c++) { F[c] += Fk[c]; T[c] += t[c]; }
        }
    }
}

// xor-butterfly sum over the L lanes of an env
template <int L>
__device__ __forceinline__ float lane_sum(float v) {
    if (L >= 2) v += __shfl_xor_sync(0xffffffffu, v, 1);
    if (L >= 4) v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

// ---------------------------------------------------------------------------------------------
// one sub-step for the env this lane (co-)owns.  `lane` = lane index within the env (0..L-1).
// LAST selects whether the outputs of StepOut are produced (only the last sub-step's are read).
template <class Topo>
__device__ __forceinline__ void substep(const DevModel *__restrict__ m, const Ground &gr, int lane,
                                        EnvState<Topo> &st, StepOut<Topo> &out, const bool LAST) {
    constexpr int NS = Topo::NS, L = Topo::L;
    const float h = m->h;
    const LinkC &rootc = m->links[0];
    float g[3] = {m->g[0], m->g[1], m->g[2]};

    float R[NS + 1][9], x[NS + 1][3], vw[NS + 1][3], vl[NS + 1][3];
    float w[NS][3], sl[NS][3], cwv[NS][3], clv[NS][3];
    float IA[NS + 1][21], pa[NS + 1][3], pl[NS + 1][3];
    float U[NS][6], Dinv[NS], uu[NS], tau[NS], diag[NS];
    constexpr int RT = NS;   // index of the root in the per-link arrays

    // ---- root kinematics
    quat_to_mat(st.rq, R[RT]);
    x[RT][0] = x[RT][1] = x[RT][2] = 0.f;
    if (m->root_fixed) {
#pragma unroll
        for (int c = 0; c < 3; c++) { vw[RT][c] = 0.f; vl[RT][c] = 0.f; }
    } else {
#pragma unroll
        for (int c = 0; c < 3; c++) { vw[RT][c] = st.rw[c]; vl[RT][c] = st.rv[c]; }
    }

    // ---- pass 1 (root + slots, parents before children)
#pragma unroll
    for (int s = -1; s < NS; s++) {
        const int i = (s < 0) ? RT : s;
        const LinkC &lk = (s < 0) ? rootc : m->links[m->slot_link[s][lane]];
        if (s >= 0) {
            const int p = (Topo::ps(s) < 0) ? RT : Topo::ps(s);
            float Rt[9], ax[3] = {lk.axis[0], lk.axis[1], lk.axis[2]};
            matmul(R[p], lk.R0, Rt);
            matvec(Rt, ax, w[s]);
            float lp[3] = {lk.lpos[0], lk.lpos[1], lk.lpos[2]}, d[3];
            const float q = st.q[s], qd = st.qd[s];
            if (lk.jtype == 0) {
                float sn, cs; b2g_sincos(q, &sn, &cs);
                const float oc = 1.f - cs;
#pragma unroll
                for (int j = 0; j < 3; j++) {   // rotate column j of Rt about the world axis w by q
                    float col[3] = {Rt[j], Rt[3 + j], Rt[6 + j]}, wxc[3];
                    cross(w[s], col, wxc);
                    float wd = dot3(w[s], col) * oc;
                    R[i][j] = col[0] * cs + wxc[0] * sn + w[s][0] * wd;
                    R[i][3 + j] = col[1] * cs + wxc[1] * sn + w[s][1] * wd;
                    R[i][6 + j] = col[2] * cs + wxc[2] * sn + w[s][2] * wd;
                }
                matvec(R[p], lp, d);
#pragma unroll
                for (int c = 0; c < 3; c++) x[i][c] = x[p][c] + d[c];
                cross(x[i], w[s], sl[s]);                        // S = (w ; x x w)
#pragma unroll
                for (int c = 0; c < 3; c++) { vw[i][c] = vw[p][c] + w[s][c] * qd; vl[i][c] = vl[p][c] + sl[s][c] * qd; }
                // c = crm(v) (S qd):  ang = vw x w qd ; lin = vw x sl qd + vl x w qd
                float a1[3], a2[3], a3[3];
                cross(vw[i], w[s], a1); cross(vw[i], sl[s], a2); cross(vl[i], w[s], a3);
#pragma unroll
                for (int c = 0; c < 3; c++) { cwv[s][c] = a1[c] * qd; clv[s][c] = (a2[c] + a3[c]) * qd; }
            } else {
#pragma unroll
                for (int c = 0; c < 9; c++) R[i][c] = Rt[c];
                matvec(R[p], lp, d);
#pragma unroll
                for (int c = 0; c < 3; c++) { x[i][c] = x[p][c] + d[c] + w[s][c] * q; sl[s][c] = w[s][c]; }
#pragma unroll
                for (int c = 0; c < 3; c++) { vw[i][c] = vw[p][c]; vl[i][c] = vl[p][c] + w[s][c] * qd; }
                float a2[3]; cross(vw[i], w[s], a2);             // S = (0 ; w): c = (0 ; vw x w qd)
#pragma unroll
                for (int c = 0; c < 3; c++) { cwv[s][c] = 0.f; clv[s][c] = a2[c] * qd; w[s][c] = 0.f; }
            }
            // joint force: explicit part + implicit diagonal (linear terms at the end of the sub-step)
            const float qp = q + h * qd;
            float f = -lk.damping * qd - lk.stiffness * qp;
            float dg = lk.armature + h * lk.damping + h * h * lk.stiffness;
            if (lk.drive_mode == 1) {
                float pd = lk.kp * (st.act[s] - qp) - lk.kd * qd;
                pd = fminf(fmaxf(pd, -lk.effort), lk.effort);
                f += pd; dg += h * lk.kd + h * h * lk.kp;
            } else {
                f += fminf(fmaxf(st.act[s], -lk.effort), lk.effort);
            }
            if (lk.limited) {
                if (q < lk.lower) { f += lk.limit_k * (lk.lower - qp) - lk.limit_d * qd; dg += h * lk.limit_d + h * h * lk.limit_k; }
                else if (q > lk.upper) { f += lk.limit_k * (lk.upper - qp) - lk.limit_d * qd; dg += h * lk.limit_d + h * h * lk.limit_k; }
            }
            tau[s] = f; diag[s] = dg;
        }
        // spatial inertia about O, world axes, and bias force p = v x* (I v) - gravity
        // (the root's share is computed on lane 0 only; the butterfly below spreads it)
        const bool mine = (s >= 0) || (lane == 0);
        const float mass = mine ? lk.mass : 0.f;
        float cl_[3] = {lk.com[0], lk.com[1], lk.com[2]}, cw_[3];
        matvec(R[i], cl_, cw_);
#pragma unroll
        for (int c = 0; c < 3; c++) cw_[c] += x[i][c];
        float T_[9], Icw[6];
        {   // Icw = R Ic R^T
            const float *I6 = lk.Ic;
            const float Im[9] = {I6[0], I6[3], I6[4], I6[3], I6[1], I6[5], I6[4], I6[5], I6[2]};
            matmul(R[i], Im, T_);
            Icw[0] = T_[0] * R[i][0] + T_[1] * R[i][1] + T_[2] * R[i][2];
            Icw[1] = T_[3] * R[i][3] + T_[4] * R[i][4] + T_[5] * R[i][5];
            Icw[2] = T_[6] * R[i][6] + T_[7] * R[i][7] + T_[8] * R[i][8];
            Icw[3] = T_[0] * R[i][3] + T_[1] * R[i][4] + T_[2] * R[i][5];
            Icw[4] = T_[0] * R[i][6] + T_[1] * R[i][7] + T_[2] * R[i][8];
            Icw[5] = T_[3] * R[i][6] + T_[4] * R[i][7] + T_[5] * R[i][8];
        }
        const float sc = mine ? 1.f : 0.f;
        const float hm[3] = {mass * cw_[0], mass * cw_[1], mass * cw_[2]};
        const float c2 = dot3(cw_, cw_);
        float *I = IA[i];
        I[0] = sc * Icw[0] + mass * (c2 - cw_[0] * cw_[0]);
        I[1] = sc * Icw[1] + mass * (c2 - cw_[1] * cw_[1]);
        I[2] = sc * Icw[2] + mass * (c2 - cw_[2] * cw_[2]);
        I[3] = sc * Icw[3] - mass * cw_[0] * cw_[1];
        I[4] = sc * Icw[4] - mass * cw_[0] * cw_[2];
        I[5] = sc * Icw[5] - mass * cw_[1] * cw_[2];
        I[6] = 0.f; I[7] = -hm[2]; I[8] = hm[1];
        I[9] = hm[2]; I[10] = 0.f; I[11] = -hm[0];
        I[12] = -hm[1]; I[13] = hm[0]; I[14] = 0.f;
        I[15] = mass; I[16] = mass; I[17] = mass; I[18] = 0.f; I[19] = 0.f; I[20] = 0.f;
        float na[3], nf[3];
        sym6_mul(I, vw[i], vl[i], na, nf);
        float t1[3], t2[3], t3[3];
        cross(vw[i], na, t1); cross(vl[i], nf, t2); cross(vw[i], nf, t3);
        float hxg[3]; cross(hm, g, hxg);
#pragma unroll
        for (int c = 0; c < 3; c++) { pa[i][c] = t1[c] + t2[c] - hxg[c]; pl[i][c] = t3[c] - mass * g[c]; }
        // contacts (root contact spheres are dealt round-robin to the L lanes)
        float dummy[3];
        link_contacts<true>(m, gr, lk, st.rp, R[i], x[i], vw[i], vl[i], I, pa[i], pl[i], dummy, dummy, dummy, dummy,
                            (s < 0) ? lane : 0, (s < 0) ? L : 1);
    }

    // ---- pass 2: leaf -> root
#pragma unroll
    for (int s = NS - 1; s >= 0; s--) {
        const int p = (Topo::ps(s) < 0) ? RT : Topo::ps(s);
        float Ua[3], Ul[3];
        sym6_mul(IA[s], w[s], sl[s], Ua, Ul);
        const float D = dot3(w[s], Ua) + dot3(sl[s], Ul) + diag[s];
        const float di = 1.f / D;
        const float u_ = tau[s] - (dot3(w[s], pa[s]) + dot3(sl[s], pl[s]));
        Dinv[s] = di; uu[s] = u_;
        U[s][0] = Ua[0]; U[s][1] = Ua[1]; U[s][2] = Ua[2]; U[s][3] = Ul[0]; U[s][4] = Ul[1]; U[s][5] = Ul[2];
        sym6_rank1(IA[s], -di, Ua, Ul);                          // Ia = IA - U U^T / D
        float ya[3], yl[3];
        sym6_mul(IA[s], cwv[s], clv[s], ya, yl);
        const float ud = u_ * di;
#pragma unroll
        for (int c = 0; c < 3; c++) { pa[p][c] += pa[s][c] + ya[c] + Ua[c] * ud; pl[p][c] += pl[s][c] + yl[c] + Ul[c] * ud; }
#pragma unroll
        for (int c = 0; c < 21; c++) IA[p][c] += IA[s][c];
    }

    // ---- root: gather the lanes' contributions, solve the floating base
    float aw[NS + 1][3], al[NS + 1][3];
    if (L > 1) {
#pragma unroll
        for (int c = 0; c < 21; c++) IA[RT][c] = lane_sum<L>(IA[RT][c]);
#pragma unroll
        for (int c = 0; c < 3; c++) { pa[RT][c] = lane_sum<L>(pa[RT][c]); pl[RT][c] = lane_sum<L>(pl[RT][c]); }
    }
    if (m->root_fixed) {
#pragma unroll
        for (int c = 0; c < 3; c++) { aw[RT][c] = 0.f; al[RT][c] = 0.f; }
    } else {
        float ba[3] = {-pa[RT][0], -pa[RT][1], -pa[RT][2]}, bl[3] = {-pl[RT][0], -pl[RT][1], -pl[RT][2]};
        sym6_solve(IA[RT], ba, bl, aw[RT], al[RT]);
    }

    // ---- pass 3: root -> leaves, integrate the joints
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int p = (Topo::ps(s) < 0) ? RT : Topo::ps(s);
        float a_w[3], a_l[3];
#pragma unroll
        for (int c = 0; c < 3; c++) { a_w[c] = aw[p][c] + cwv[s][c]; a_l[c] = al[p][c] + clv[s][c]; }
        const float Ua_ = U[s][0] * a_w[0] + U[s][1] * a_w[1] + U[s][2] * a_w[2] + U[s][3] * a_l[0] + U[s][4] * a_l[1] + U[s][5] * a_l[2];
        const float qdd = (uu[s] - Ua_) * Dinv[s];
#pragma unroll
        for (int c = 0; c < 3; c++) { aw[s][c] = a_w[c] + w[s][c] * qdd; al[s][c] = a_l[c] + sl[s][c] * qdd; }
        if (LAST) out.dof_force[s] = tau[s] - (diag[s] - m->links[m->slot_link[s][lane]].armature) * qdd;
        st.qd[s] += h * qdd;
        st.q[s] += h * st.qd[s];
    }

    // ---- forces actually applied by the contacts over this sub-step (sensors, net contact force)
    if (LAST) {
#pragma unroll
        for (int s = -1; s < NS; s++) {
            const int i = (s < 0) ? RT : s;
            const LinkC &lk = (s < 0) ? rootc : m->links[m->slot_link[s][lane]];
            float F[3] = {0.f, 0.f, 0.f}, T[3] = {0.f, 0.f, 0.f};
            link_contacts<false>(m, gr, lk, st.rp, R[i], x[i], vw[i], vl[i], IA[i], pa[i], pl[i], aw[i], al[i], F, T,
                                 (s < 0) ? lane : 0, (s < 0) ? L : 1);
#pragma unroll
            for (int c = 0; c < 3; c++) { out.cfF[i][c] = F[c]; out.cfT[i][c] = T[c]; }
#pragma unroll
            for (int c = 0; c < 9; c++) out.R[i][c] = R[i][c];
        }
        if (L > 1) {
#pragma unroll
            for (int c = 0; c < 3; c++) { out.cfF[RT][c] = lane_sum<L>(out.cfF[RT][c]); out.cfT[RT][c] = lane_sum<L>(out.cfT[RT][c]); }
        }
    }

    // ---- root integration (classical acceleration of the origin = spatial + w x v)
    if (!m->root_fixed) {
        float wxv[3]; cross(st.rw, st.rv, wxv);
#pragma unroll
        for (int c = 0; c < 3; c++) { st.rw[c] += h * aw[RT][c]; st.rv[c] += h * (al[RT][c] + wxv[c]); }
#pragma unroll
        for (int c = 0; c < 3; c++) st.rp[c] += h * st.rv[c];
        const float wn2 = dot3(st.rw, st.rw);
        float dq[4];
        if (wn2 > 1e-24f) {
            const float wn = sqrtf(wn2);
            float sn, cs; b2g_sincos(0.5f * wn * h, &sn, &cs);
            const float k = sn / wn;
            dq[0] = st.rw[0] * k; dq[1] = st.rw[1] * k; dq[2] = st.rw[2] * k; dq[3] = cs;
        } else { dq[0] = 0.5f * h * st.rw[0]; dq[1] = 0.5f * h * st.rw[1]; dq[2] = 0.5f * h * st.rw[2]; dq[3] = 1.f; }
        const float qx = st.rq[0], qy = st.rq[1], qz = st.rq[2], qw = st.rq[3];
        float nq[4] = {dq[3] * qx + dq[0] * qw + dq[1] * qz - dq[2] * qy,
                       dq[3] * qy - dq[0] * qz + dq[1] * qw + dq[2] * qx,
                       dq[3] * qz + dq[0] * qy - dq[1] * qx + dq[2] * qw,
                       dq[3] * qw - dq[0] * qx - dq[1] * qy - dq[2] * qz};
        const float inv = rsqrtf(nq[0] * nq[0] + nq[1] * nq[1] + nq[2] * nq[2] + nq[3] * nq[3]);
#pragma unroll
        for (int c = 0; c < 4; c++) st.rq[c] = nq[c] * inv;
    }
}

}  // namespace b2g
