// tests/kin_host.cu -- TEST INFRASTRUCTURE: runs the phases of the Jacobian / mass-matrix kernel
// (isaacgymenvs_b200/csrc/b2g_kin.cuh) on the HOST, lane by lane in the order the warp executes them, so the exact
// arithmetic of the CUDA kernel can be compared with the fp64 oracle on a machine without a GPU.  Built by
// tests/test_kin_tensors.py (nvcc -shared, host code only is called; nothing here is part of the product library).
#include <vector>
#include "../isaacgymenvs_b200/csrc/b2g_kin_host.h"

using namespace b2g;

extern "C" int kin_host_shape(const b2g_model *m, int *rows, int *nc) {
    KinModel t;
    if (kin_build(m, 1, t)) return -1;
    *rows = t.rows; *nc = t.nc;
    return 0;
}

extern "C" int kin_host_tensors(const b2g_model *m, int root_stride, int N, const float *root, const float *dof, float *jac, float *mass) {
    KinModel *tp = new KinModel; KinModel &t = *tp;
    if (kin_build(m, root_stride, t)) { delete tp; return -1; }
    KinScratch *sp = new KinScratch; KinScratch &s = *sp;
    const int nl = t.nl, nd = nl - 1, nc = t.nc;
    for (int e = 0; e < N; e++) {
        const float *r = root + 13 * (size_t)e * root_stride, *d = dof + 2 * (size_t)e * nd;
        for (int lane = 0; lane < nl; lane++) kin_local(lane, t, s, r, lane > 0 ? d[2 * (lane - 1)] : 0.f);
        for (int lv = 1; lv <= t.maxdepth; lv++) for (int lane = 0; lane < 32; lane++) kin_level(lane, lv, t, s);
        for (int lane = 0; lane < nl; lane++) kin_link(lane, t, s);
        if (jac) for (int lane = 0; lane < t.nb; lane++) kin_body(lane, t, s);
        if (mass) {
            for (int lv = t.maxdepth - 1; lv >= 0; lv--) for (int lane = 0; lane < 32; lane++) kin_composite_level(lane, lv, t, s);
            for (int lane = 0; lane < nl; lane++) kin_momentum(lane, t, s);
        }
        for (int c = 0; c < nc; c++) {
            const KinCol k = kin_col(c, t, s);
            if (mass) for (int a = 0; a < nc; a++) mass[(size_t)e * nc * nc + (size_t)a * nc + c] = kin_mass_col(a, k, t, s);
            if (jac) for (int b = 0; b < t.rows; b++) {
                float o[6]; kin_jac_col(t.row0 + b, k, t, s, o);
                for (int r = 0; r < 6; r++) jac[(size_t)e * t.rows * 6 * nc + (size_t)(b * 6 + r) * nc + c] = o[r];
            }
        }
    }
    delete tp; delete sp;
    return 0;
}
