"""CPU sanity for the ShadowHand physics (oracle): cube dropped on the open hand."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isaacgymenvs_b200.assets import load_compiled as load_asset
from oracle.oracle import OracleSim

m = load_asset("shadow_hand"); cube = load_asset("cube")
print("links", m.nl, "bodies", m.nb, "dofs", m.ndof, "cps", len(m.cp_link), "boxes", len(m.box_link))
print("kp", np.round(m.kp, 2)); print("kd", np.round(m.kd, 3)); print("damping", np.round(m.damping, 3)); print("arm", m.armature)
print("mass", np.round(m.mass, 4))
obj = dict(mass=float(cube.mass[0]), inertia=[float(cube.inertia[0][0])] * 3, half=[0.025] * 3, mu=1.0)
prec = sys.argv[1] if len(sys.argv) > 1 else "f64"
dt_ = np.float64 if prec == "f64" else np.float32
tend = [t for t in m.tendons if t["name"] in ("robot0:T_FFJ1c", "robot0:T_MFJ1c", "robot0:T_RFJ1c", "robot0:T_LFJ1c")]
sim = OracleSim(m, 1 / 60, 2, precision=prec, obj=obj, tendons=tend, tendon_k=30.0, tendon_d=0.1)
N = 4
root = np.zeros((N, 13), dt_); root[:, 2] = 0.5; root[:, 3:7] = m.default_root_quat
dof = np.zeros((N, m.ndof, 2), dt_)
o = np.zeros((N, 13), dt_); o[:, 0:3] = [0, -0.39, 0.6]; o[:, 6] = 1
o[:, 0] += np.linspace(-0.01, 0.01, N); 
tgt = np.zeros((N, m.ndof), dt_)
for k in range(240):
    out = sim.simulate(root, dof, target=tgt, obj=o)
    if k % 20 == 0 or k == 239:
        print(k, "cube", np.round(o[0, :3], 4), "v", np.round(o[0, 7:10], 3), "w", np.round(o[0, 10:13], 2), "|q|max", np.abs(dof[0, :, 0]).max().round(3), "|qd|max", np.abs(dof[0, :, 1]).max().round(2))
print(np.round(o[:, :3], 4))
