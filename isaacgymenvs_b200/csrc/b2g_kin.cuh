// b2g_kin.cuh -- Jacobian and joint-space mass-matrix tensors (sm_100a).
//
// Stands behind gym.acquire_jacobian_tensor / acquire_mass_matrix_tensor + refresh_jacobian_tensors /
// refresh_mass_matrix_tensors (reference call sites tasks/franka_cube_stack.py:388-392,439-440; consumer: the
// operational-space controller :600-627, which reads the end effector's rows of J and the arm's block of M).
//
// Layouts (the closed binary defines them, the call sites fix the fixed-base case):
//   J  (N, rows, 6, nc)   row = body (a fixed base has no row for its base body 0: `jacobian[:, joint_index]`),
//                         6 = world linear velocity of the body-frame origin (3) | world angular velocity (3),
//                         nc = nd joint columns, preceded for a FLOATING base by 6 base columns: world linear then
//                         world angular velocity of the root origin (the order the root-state tensor carries them);
//   M  (N, nc, nc)        composite-rigid-body inertia in the same coordinates, + joint armature on the diagonal.
//
// Formulation (differs on purpose from the oracle's body-coordinate CRBA with 6x6 Pluecker transforms): everything in
// WORLD-ALIGNED axes about the root origin O.  There a rigid body's inertia is ten additive numbers -- the inertia tensor
// about O (6), mass x COM offset (3), mass (1) -- so the composite inertia of a sub-tree is a plain sum over descendants,
// and M[i][j] = S_j . (Ic_i S_i) needs no transforms at all.  A joint's motion subspace about O is S = (sa ; sb) -- hinge (w ; x cross w),
// slide (0 ; w), base coordinates (unit vectors) -- so the twist of a body origin p is (sa x p + sb ; sa) for every column kind.
//
// Work decomposition: ONE WARP PER ENVIRONMENT, lane = link (nl <= 32) for the kinematics / inertias, lane = body for the
// body origins, lane = output column for the fills, so every global store is a contiguous run of one output row.  Lanes
// exchange through a per-warp shared-memory scratch ordered by __syncwarp; every phase is a __host__ __device__ function
// of (lane, tables, scratch), so tests/kin_host.cu runs the exact device arithmetic on the CPU against the fp64 oracle.
#pragma once
#include "b2g_device.cuh"

namespace b2g {

constexpr int KIN_MAX_CHILD = 8;

// constant tables of one articulation (built on the host by kin_build, b2g_kin_host.h)
struct alignas(16) KinModel {
    int nl, nb, nbase, nc;          // links, bodies, base columns (0 fixed / 6 floating), columns
    int rows, row0;                 // Jacobian rows, first body with a row (1 for a fixed base)
    int maxdepth, root_stride;      // tree depth; actors per env in the root-state tensor
    int parent[MAX_LINKS], depth[MAX_LINKS], slide[MAX_LINKS], body_link[MAX_LINKS];
    unsigned anc[MAX_LINKS];        // bit j set: link j (j >= 1) is link i itself or one of its ancestors
    int nchild[MAX_LINKS], child[MAX_LINKS][KIN_MAX_CHILD];
    float R0[MAX_LINKS][9], lpos[MAX_LINKS][3], axis[MAX_LINKS][3], com[MAX_LINKS][3], Ic[MAX_LINKS][6];
    float mass[MAX_LINKS], armature[MAX_LINKS], body_pos[MAX_LINKS][3];
};

// per-warp scratch; odd strides: lane-indexed accesses fall into different banks
struct KinScratch {
    float R[MAX_LINKS][9];          // link frame, world axes
    float x[MAX_LINKS][3];          // link origin relative to the root origin O, world axes
    float sa[MAX_LINKS][3];         // motion subspace S = (sa ; sb) about O: hinge (w ; x cross w), slide (0 ; w)
    float sb[MAX_LINKS][3];
    float in[MAX_LINKS][11];        // inertia about O: I (xx yy zz xy xz yz), m c (3), m -- the link's own, then (phase 3) its sub-tree's
    float nf[MAX_LINKS][7];         // Ic_i S_i: angular momentum about O (3), linear momentum (3)
    float pb[MAX_LINKS][3];         // body-frame origins relative to O
    float cb[11];                   // composite inertia of the whole articulation (base block of M)
};

// ---- phase 0: joint transform of link i in its parent's frame -> scratch (overwritten by the world frame in phase 1)
B2G_HD void kin_local(int i, const KinModel &t, KinScratch &s, const float *root, float q) {
    if (i == 0) {
        const float rq[4] = {root[3], root[4], root[5], root[6]};
        float R[9];
        // plain normalisation (not the raw rsqrt of the step kernels: this is not a hot loop)
        const float n = 1.0f / sqrtf(rq[0] * rq[0] + rq[1] * rq[1] + rq[2] * rq[2] + rq[3] * rq[3]);
        const float x = rq[0] * n, y = rq[1] * n, z = rq[2] * n, w = rq[3] * n;
        R[0] = 1.f - 2.f * (y * y + z * z); R[1] = 2.f * (x * y - z * w); R[2] = 2.f * (x * z + y * w);
        R[3] = 2.f * (x * y + z * w); R[4] = 1.f - 2.f * (x * x + z * z); R[5] = 2.f * (y * z - x * w);
        R[6] = 2.f * (x * z - y * w); R[7] = 2.f * (y * z + x * w); R[8] = 1.f - 2.f * (x * x + y * y);
#pragma unroll
        for (int c = 0; c < 9; c++) s.R[0][c] = R[c];
        s.x[0][0] = s.x[0][1] = s.x[0][2] = 0.f;
        return;
    }
    const float *R0 = t.R0[i], *ax = t.axis[i];
    if (!t.slide[i]) {
        float sn, cs; sincosf(q, &sn, &cs);
        const float oc = 1.f - cs, ux = ax[0], uy = ax[1], uz = ax[2];
        const float Rj[9] = {cs + ux * ux * oc, ux * uy * oc - uz * sn, ux * uz * oc + uy * sn,
                             uy * ux * oc + uz * sn, cs + uy * uy * oc, uy * uz * oc - ux * sn,
                             uz * ux * oc - uy * sn, uz * uy * oc + ux * sn, cs + uz * uz * oc};
        float R[9]; matmul(R0, Rj, R);
#pragma unroll
        for (int c = 0; c < 9; c++) s.R[i][c] = R[c];
#pragma unroll
        for (int c = 0; c < 3; c++) s.x[i][c] = t.lpos[i][c];
    } else {
        float d[3]; matvec(R0, ax, d);
#pragma unroll
        for (int c = 0; c < 9; c++) s.R[i][c] = R0[c];
#pragma unroll
        for (int c = 0; c < 3; c++) s.x[i][c] = t.lpos[i][c] + d[c] * q;
    }
}

// ---- phase 1 (once per tree level d = 1 .. maxdepth): compose with the parent's world frame
B2G_HD void kin_level(int i, int d, const KinModel &t, KinScratch &s) {
    if (i >= t.nl || t.depth[i] != d) return;
    const int p = t.parent[i];
    float Rp[9], Rl[9], R[9], r[3], wr[3];
#pragma unroll
    for (int c = 0; c < 9; c++) { Rp[c] = s.R[p][c]; Rl[c] = s.R[i][c]; }
#pragma unroll
    for (int c = 0; c < 3; c++) r[c] = s.x[i][c];
    matmul(Rp, Rl, R); matvec(Rp, r, wr);
#pragma unroll
    for (int c = 0; c < 9; c++) s.R[i][c] = R[c];
#pragma unroll
    for (int c = 0; c < 3; c++) s.x[i][c] = s.x[p][c] + wr[c];
}

// ---- phase 2: world joint axis, motion subspace, the link's own inertia about O
B2G_HD void kin_link(int i, const KinModel &t, KinScratch &s) {
    float R[9], x[3], w[3], sa[3] = {0.f, 0.f, 0.f}, sb[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 9; c++) R[c] = s.R[i][c];
#pragma unroll
    for (int c = 0; c < 3; c++) x[c] = s.x[i][c];
    if (i > 0) {
        const float ax[3] = {t.axis[i][0], t.axis[i][1], t.axis[i][2]};
        matvec(R, ax, w);
        if (!t.slide[i]) { cross(x, w, sb); sa[0] = w[0]; sa[1] = w[1]; sa[2] = w[2]; }
        else { sb[0] = w[0]; sb[1] = w[1]; sb[2] = w[2]; }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) { s.sa[i][c] = sa[c]; s.sb[i][c] = sb[c]; }
    // inertia about the COM in world axes: R Ic R^T, then the parallel-axis term to O
    const float *I6 = t.Ic[i];
    const float Il[9] = {I6[0], I6[3], I6[4], I6[3], I6[1], I6[5], I6[4], I6[5], I6[2]};
    float T[9], Rt[9] = {R[0], R[3], R[6], R[1], R[4], R[7], R[2], R[5], R[8]}, Iw[9];
    matmul(R, Il, T); matmul(T, Rt, Iw);
    const float lc[3] = {t.com[i][0], t.com[i][1], t.com[i][2]};
    float c[3]; matvec(R, lc, c);
#pragma unroll
    for (int k = 0; k < 3; k++) c[k] += x[k];
    const float m = t.mass[i], cc = dot3(c, c);
    s.in[i][0] = Iw[0] + m * (cc - c[0] * c[0]);
    s.in[i][1] = Iw[4] + m * (cc - c[1] * c[1]);
    s.in[i][2] = Iw[8] + m * (cc - c[2] * c[2]);
    s.in[i][3] = Iw[1] - m * c[0] * c[1];
    s.in[i][4] = Iw[2] - m * c[0] * c[2];
    s.in[i][5] = Iw[5] - m * c[1] * c[2];
    s.in[i][6] = m * c[0]; s.in[i][7] = m * c[1]; s.in[i][8] = m * c[2];
    s.in[i][9] = m;
}

// ---- phase 3 (once per tree level d = maxdepth - 1 .. 0): composite inertia of the sub-tree rooted at link i = its own + its
// children's (already complete: they are one level deeper).  In place; a fixed order of the children keeps the sums reproducible.
B2G_HD void kin_composite_level(int i, int d, const KinModel &t, KinScratch &s) {
    if (i >= t.nl || t.depth[i] != d || t.nchild[i] == 0) return;
    float a[10];
#pragma unroll
    for (int c = 0; c < 10; c++) a[c] = s.in[i][c];
    for (int q = 0; q < t.nchild[i]; q++) {
        const int k = t.child[i][q];
#pragma unroll
        for (int c = 0; c < 10; c++) a[c] += s.in[k][c];
    }
#pragma unroll
    for (int c = 0; c < 10; c++) s.in[i][c] = a[c];
}
// ---- phase 4: Ic_i S_i = (angular momentum about O ; linear momentum) of the sub-tree moving with joint i at unit rate
B2G_HD void kin_momentum(int i, const KinModel &t, KinScratch &s) {
    if (i == 0) {
#pragma unroll
        for (int c = 0; c < 10; c++) s.cb[c] = s.in[0][c];
        return;
    }
    float a[10];
#pragma unroll
    for (int c = 0; c < 10; c++) a[c] = s.in[i][c];
    const float sa[3] = {s.sa[i][0], s.sa[i][1], s.sa[i][2]}, sb[3] = {s.sb[i][0], s.sb[i][1], s.sb[i][2]}, mc[3] = {a[6], a[7], a[8]};
    // n = I_O sa + mc x sb,  f = m sb - mc x sa
    float t1[3], t2[3];
    cross(mc, sb, t1); cross(mc, sa, t2);
    s.nf[i][0] = a[0] * sa[0] + a[3] * sa[1] + a[4] * sa[2] + t1[0];
    s.nf[i][1] = a[3] * sa[0] + a[1] * sa[1] + a[5] * sa[2] + t1[1];
    s.nf[i][2] = a[4] * sa[0] + a[5] * sa[1] + a[2] * sa[2] + t1[2];
#pragma unroll
    for (int c = 0; c < 3; c++) s.nf[i][3 + c] = a[9] * sb[c] - t2[c];
}

// ---- phase 2b: origin of body b's frame relative to O
B2G_HD void kin_body(int b, const KinModel &t, KinScratch &s) {
    const int l = t.body_link[b];
    float R[9], o[3];
#pragma unroll
    for (int c = 0; c < 9; c++) R[c] = s.R[l][c];
    const float bp[3] = {t.body_pos[b][0], t.body_pos[b][1], t.body_pos[b][2]};
    matvec(R, bp, o);
#pragma unroll
    for (int c = 0; c < 3; c++) s.pb[b][c] = s.x[l][c] + o[c];
}

// ---- fills.  A lane owns one COLUMN c of both tensors for the whole env, so what belongs to the column -- the joint's world
// axis, origin, motion subspace, Ic S and ancestor mask -- is fetched once (KinCol) and every output row costs a few
// multiply-adds; the stores of a warp are one contiguous run of an output row.
struct KinCol {
    int base;               // 0..5: base column (world linear 0..2, world angular 3..5), -1: joint column
    int link, slide;        // joint column: its link
    unsigned anc;
    float sa[3], sb[3], n[3], f[3], arm;      // the column's motion subspace about O and Ic S
};
B2G_HD KinCol kin_col(int c, const KinModel &t, const KinScratch &s) {
    KinCol k;
    k.base = c < t.nbase ? c : -1;
    const int j = c < t.nbase ? 0 : c - t.nbase + 1;
    k.link = j; k.slide = t.slide[j]; k.anc = t.anc[j]; k.arm = t.armature[j];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        k.sa[a] = c < t.nbase ? ((c >= 3 && c - 3 == a) ? 1.f : 0.f) : s.sa[j][a];      // base: (0 ; e_c) linear, (e_c ; 0) angular
        k.sb[a] = c < t.nbase ? ((c < 3 && c == a) ? 1.f : 0.f) : s.sb[j][a];
        k.n[a] = s.nf[j][a]; k.f[a] = s.nf[j][3 + a];
    }
    return k;
}

B2G_HD float kin_pick(const float v[3], int a) { return a == 0 ? v[0] : (a == 1 ? v[1] : v[2]); }   // no dynamic register indexing

// the six entries (linear 3, angular 3) of body b's Jacobian block in column k.  With S = (sa ; sb) about O the twist of a point p
// riding on the joint's sub-tree is (sa x p + sb ; sa) -- hinge, slide and the six base coordinates alike (base: S = unit vectors,
// every body rides on it) -- so the lanes of a warp run one branch-free expression.
B2G_HD void kin_jac_col(int b, const KinCol &k, const KinModel &t, const KinScratch &s, float o[6]) {
    const float pb[3] = {s.pb[b][0], s.pb[b][1], s.pb[b][2]};
    const bool on = k.base >= 0 || ((t.anc[t.body_link[b]] >> k.link) & 1u);     // the joint lies between the base and this body
    float v[3]; cross(k.sa, pb, v);
#pragma unroll
    for (int c = 0; c < 3; c++) { o[c] = on ? v[c] + k.sb[c] : 0.f; o[3 + c] = on ? k.sa[c] : 0.f; }
}

// entry (row a, column k) of the mass matrix
B2G_HD float kin_mass_col(int a, const KinCol &k, const KinModel &t, const KinScratch &s) {
    const int nb = t.nbase;
    if (a < nb) {
        const int ax = a < 3 ? a : a - 3;
        if (k.base < 0) return a < 3 ? kin_pick(k.f, ax) : kin_pick(k.n, ax);           // base row against a joint column: Ic_j S_j
        const float *cb = s.cb;
        const int kx = k.base < 3 ? k.base : k.base - 3;
        if (a < 3 && k.base < 3) return ax == kx ? cb[9] : 0.f;     // m 1
        if (a >= 3 && k.base >= 3) return ax == kx ? cb[ax] : cb[2 + ax + kx];   // I_O: (0,1) -> 3, (0,2) -> 4, (1,2) -> 5
        // linear row i, angular column j: p = w x (m c) -> -[mc]x [i][j]; the (angular, linear) entry is its transpose
        const int i = a < 3 ? ax : kx, j = a < 3 ? kx : ax;
        if (i == j) return 0.f;
        const float v = cb[6 + 3 - i - j];
        return ((j - i + 3) % 3 == 1) ? v : -v;
    }
    const int i = a - nb + 1;
    if (k.base >= 0) { const int kx = k.base < 3 ? k.base : k.base - 3; return k.base < 3 ? s.nf[i][3 + kx] : s.nf[i][kx]; }
    const int j = k.link;
    // S_deeper-or-equal's Ic S against the other's S: column joint j on row link i's path -> S_j . (Ic_i S_i); row link i on column
    // joint j's path -> S_i . (Ic_j S_j); different branches of the tree -> 0.  Selected without branching: the lanes of a warp
    // (columns) fall into all three cases in the same row.
    const bool ja = (t.anc[i] >> j) & 1u, ib = (k.anc >> i) & 1u;
    float v = 0.f;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float sa = ja ? k.sa[c] : s.sa[i][c], sb = ja ? k.sb[c] : s.sb[i][c];
        const float n = ja ? s.nf[i][c] : k.n[c], f = ja ? s.nf[i][3 + c] : k.f[c];
        v += sa * n + sb * f;
    }
    if (!(ja || ib)) v = 0.f;
    return i == j ? v + k.arm : v;
}

#ifdef __CUDACC__
// One warp per env, grid-stride.  jac / mass may be null (only the other tensor is refreshed).
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) kin_tensors_kernel(const KinModel *__restrict__ gm, const float *__restrict__ g_root,
                                                                 const float *__restrict__ g_dof, float *__restrict__ jac,
                                                                 float *__restrict__ mass, int N) {
    __shared__ KinModel t;
    __shared__ KinScratch sc[WARPS];
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(gm);
        uint4 *dst = reinterpret_cast<uint4 *>(&t);
        for (int i = threadIdx.x; i < (int)(sizeof(KinModel) / 16); i += blockDim.x) dst[i] = src[i];
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    KinScratch &s = sc[warp];
    const int nl = t.nl, nd = nl - 1, nc = t.nc;
    for (int e = blockIdx.x * WARPS + warp; e < N; e += gridDim.x * WARPS) {
        const float *root = g_root + 13 * (size_t)e * t.root_stride;
        const float2 *dof = reinterpret_cast<const float2 *>(g_dof) + (size_t)e * nd;
        __syncwarp();                                        // the previous env's readers are done with the scratch
        if (lane < nl) kin_local(lane, t, s, root, lane > 0 ? dof[lane - 1].x : 0.f);
        for (int d = 1; d <= t.maxdepth; d++) { __syncwarp(); kin_level(lane, d, t, s); }
        __syncwarp();
        if (lane < nl) kin_link(lane, t, s);
        if (jac && lane < t.nb) kin_body(lane, t, s);
        __syncwarp();
        if (mass) {
            for (int d = t.maxdepth - 1; d >= 0; d--) { kin_composite_level(lane, d, t, s); __syncwarp(); }
            if (lane < nl) kin_momentum(lane, t, s);
            __syncwarp();
        }
        for (int c = lane; c < nc; c += 32) {                // one pass for nc <= 32, two for the largest floating trees
            const KinCol k = kin_col(c, t, s);
            if (mass) {
                float *M = mass + (size_t)e * nc * nc + c;
                for (int a = 0; a < nc; a++, M += nc) *M = kin_mass_col(a, k, t, s);
            }
            if (jac) {
                float *J = jac + (size_t)e * t.rows * 6 * nc + c;
                const size_t st = (size_t)nc;
                for (int b = 0; b < t.rows; b++) {
                    float o[6]; kin_jac_col(t.row0 + b, k, t, s, o);
                    J[0] = o[0]; J[st] = o[1]; J[2 * st] = o[2]; J[3 * st] = o[3]; J[4 * st] = o[4]; J[5 * st] = o[5];
                    J += 6 * st;
                }
            }
        }
    }
}
#endif

}   // namespace b2g
