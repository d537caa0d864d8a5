"""Triangle meshes as far as the importer needs them: mass properties of the volume a collision mesh encloses.

The reference's robot descriptions that carry no <inertial> (urdf/franka_description/robots/franka_panda_gripper.urdf) get
their link masses from the collision geometry x AssetOptions.density inside the closed importer; for a mesh that is the
volume integral below.  Contact with meshes is NOT modelled -- only their mass, centre of mass and
inertia, which is what the Jacobian / mass-matrix tensors and the dynamics of an arm need."""
import numpy as np


def load_obj(path):
    """Wavefront .obj -> (vertices (n,3) float64, triangles (m,3) int64); polygons are fanned, negative indices resolved."""
    V, F = [], []
    with open(path) as f:
        for line in f:
            if line.startswith("v "):
                V.append([float(x) for x in line.split()[1:4]])
            elif line.startswith("f "):
                idx = [int(tok.split("/")[0]) for tok in line.split()[1:]]
                idx = [i - 1 if i > 0 else len(V) + i for i in idx]
                for k in range(1, len(idx) - 1):
                    F.append([idx[0], idx[k], idx[k + 1]])
    if not V or not F:
        raise ValueError(f"{path}: no vertices / faces")
    return np.asarray(V, dtype=np.float64), np.asarray(F, dtype=np.int64)


# covariance of the canonical tetrahedron (0, e1, e2, e3) for unit density
_C0 = np.array([[2.0, 1.0, 1.0], [1.0, 2.0, 1.0], [1.0, 1.0, 2.0]]) / 120.0


def mass_properties(V, F):
    """(volume, centre of mass (3), inertia tensor about the centre of mass (3x3)) of the solid a closed triangle mesh bounds, for
    unit density: signed tetrahedra from the origin (exact for a closed, consistently oriented surface; a globally flipped
    orientation is corrected; an open surface gives the volume of its cone to the origin -- the caller is warned by the sign)."""
    a, b, c = V[F[:, 0]], V[F[:, 1]], V[F[:, 2]]
    det = np.einsum("ij,ij->i", a, np.cross(b, c))                  # 6 x signed volume of each tetrahedron
    vol = det.sum() / 6.0
    sgn = 1.0 if vol >= 0 else -1.0
    vol *= sgn; det = det * sgn
    if vol <= 1e-18:
        raise ValueError("mesh encloses no volume")
    com = (det[:, None] * (a + b + c)).sum(0) / (24.0 * vol)
    A = np.stack([a, b, c], axis=2)                                 # columns a b c
    C = np.einsum("n,nij,jk,nlk->il", det, A, _C0, A)               # covariance about the origin
    C -= vol * np.outer(com, com)                                   # ... about the centre of mass
    I = np.trace(C) * np.eye(3) - C
    return float(vol), com, I
