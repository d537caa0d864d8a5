"""ctypes front-end of the CPU oracle (oracle/aba_oracle.c).  TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs; never by isaacgymenvs_b200/.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def _lib(prec):
    path = os.path.join(_HERE, f"liboracle_{prec}.so")
    if not os.path.exists(path):
        build()
    return C.CDLL(path)


class _CModel(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("nl", "ncp", "nb", "nsens", "root_fixed", "gravity_on", "substeps", "pad0")] + \
               [(n, C.c_void_p) for n in ("parent", "jtype", "limited", "drive_mode", "cp_link", "cp_body",
                                          "body_link", "sensor_body",
                                          "axis", "lpos", "lquat", "mass", "com", "inertia",
                                          "armature", "damping", "stiffness", "lower", "upper", "effort",
                                          "kp", "kd", "limit_k", "limit_d",
                                          "cp_pos", "cp_radius", "cp_mu", "body_pos", "body_quat", "hfield")] + \
               [("hf_nx", C.c_int), ("hf_ny", C.c_int), ("hf_scale", C.c_double), ("hf_ox", C.c_double),
                ("hf_oy", C.c_double), ("kn", C.c_double), ("cn", C.c_double), ("vs", C.c_double),
                ("gravity", C.c_double * 3), ("dt", C.c_double),
                ("obj_on", C.c_int), ("obj_gravity_on", C.c_int), ("nbx", C.c_int), ("obj_coupling", C.c_int),
                ("obj_mass", C.c_double), ("obj_inertia", C.c_double * 3), ("obj_half", C.c_double * 3),
                ("obj_kn", C.c_double), ("obj_cn", C.c_double), ("obj_mu", C.c_double),
                ("box_link", C.c_void_p), ("box_body", C.c_void_p), ("box_pos", C.c_void_p), ("box_quat", C.c_void_p), ("box_half", C.c_void_p),
                ("nten", C.c_int), ("pad2", C.c_int), ("ten_dof", C.c_void_p), ("ten_coef", C.c_void_p),
                ("ten_range", C.c_void_p), ("ten_k", C.c_double), ("ten_d", C.c_double),
                ("angular_damping", C.c_double), ("linear_damping", C.c_double), ("max_angular_velocity", C.c_double),
                ("obj_angular_damping", C.c_double), ("obj_linear_damping", C.c_double),
                ("self_on", C.c_int), ("pad3", C.c_int), ("self_pairs", C.c_void_p),
                ("self_kn", C.c_double), ("self_cn", C.c_double), ("self_mu", C.c_double), ("obj_round", C.c_double),
                ("obj_max_angular_velocity", C.c_double)]


def object_contact_gains(mass):
    """Penalty gains of the hand-object and object-ground contacts (the object is light, so the
    gains scale with ITS mass; critically damped for the two-body reduced mass)."""
    kn = 10000.0 * mass
    return kn, 2.0 * np.sqrt(kn * mass / 4.0)


class OracleSim:
    """One articulation model replicated over num_envs independent environments."""

    def __init__(self, model, dt, substeps, gravity=(0.0, 0.0, -9.81), ground_mu=1.0, precision="f64",
                 hfield=None, hf_scale=1.0, hf_origin=(0.0, 0.0), threads=1, obj=None, tendons=None,
                 tendon_k=0.0, tendon_d=0.0):
        self.model, self.prec = model, precision
        self.dtype = np.float64 if precision == "f64" else np.float32
        self.lib = _lib(precision)
        assert self.lib.oracle_real_size() == np.dtype(self.dtype).itemsize
        self.lib.oracle_set_threads(int(threads))
        m = model
        self._keep = {}

        def arr(name, a, dt_):
            a = np.ascontiguousarray(a, dtype=dt_)
            self._keep[name] = a
            return a.ctypes.data
        cm = _CModel()
        cm.nl, cm.ncp, cm.nb, cm.nsens = m.nl, len(m.cp_link), m.nb, len(m.sensor_body)
        cm.root_fixed, cm.gravity_on, cm.substeps = int(m.root_fixed), int(m.gravity_on), int(substeps)
        for n in ("parent", "jtype", "limited", "drive_mode", "cp_link", "cp_body", "body_link", "sensor_body"):
            setattr(cm, n, arr(n, getattr(m, n), np.int32))
        for n in ("axis", "lpos", "lquat", "mass", "com", "inertia", "armature", "damping", "stiffness", "lower",
                  "upper", "effort", "kp", "kd", "limit_k", "limit_d", "cp_pos", "cp_radius", "body_pos", "body_quat"):
            setattr(cm, n, arr(n, getattr(m, n), np.float64))
        # PhysX default friction combine mode is "average" of the two materials
        cm.cp_mu = arr("cp_mu", 0.5 * (np.asarray(m.cp_mu) + ground_mu), np.float64)
        if hfield is not None:
            cm.hfield = arr("hfield", hfield, np.float64)
            cm.hf_nx, cm.hf_ny = hfield.shape
            cm.hf_scale, cm.hf_ox, cm.hf_oy = hf_scale, hf_origin[0], hf_origin[1]
        else:
            cm.hfield = None
        cm.kn, cm.cn, cm.vs = m.contact_kn, m.contact_cn, m.contact_vs
        cm.gravity = (C.c_double * 3)(*gravity)
        cm.dt = dt
        cm.angular_damping = float(getattr(m, "angular_damping", 0.0) or 0.0)
        cm.linear_damping = float(getattr(m, "linear_damping", 0.0) or 0.0)
        cm.max_angular_velocity = float(getattr(m, "max_angular_velocity", 0.0) or 0.0)
        cm.self_on = 0
        if getattr(m, "self_collide", False):
            cm.self_on = 1
            cm.self_pairs = arr("self_pairs", m.self_pairs, np.uint8)
            cm.self_kn, cm.self_cn, cm.self_mu = float(m.self_kn), float(m.self_cn), float(m.self_mu)
        # free object (ShadowHand's cube): obj = dict(mass, inertia(3), half(3), mu, gravity_on)
        cm.obj_on = 0
        if obj is not None:
            cm.obj_on, cm.obj_gravity_on = 1, int(obj.get("gravity_on", 1))
            cm.obj_coupling = int(obj.get("coupling", 0))      # 0 = the engine's block-Jacobi; 1 = Gauss-Seidel experiment
            cm.obj_mass = float(obj["mass"])
            cm.obj_angular_damping, cm.obj_linear_damping = float(obj.get("angular_damping", 0.0)), float(obj.get("linear_damping", 0.0))
            cm.obj_inertia = (C.c_double * 3)(*obj["inertia"]); cm.obj_half = (C.c_double * 3)(*obj["half"])
            cm.obj_round = float(obj.get("round", 0.0))
            cm.obj_max_angular_velocity = float(obj.get("max_angular_velocity", 64.0))      # gymapi.AssetOptions default
            kn, cn = object_contact_gains(cm.obj_mass)
            cm.obj_kn, cm.obj_cn, cm.obj_mu = kn, cn, float(obj.get("mu", 1.0))
            bl = getattr(m, "box_link", None)
            cm.nbx = 0 if bl is None else len(bl)
            if cm.nbx:
                cm.box_link = arr("box_link", m.box_link, np.int32)
                cm.box_body = arr("box_body", m.box_body, np.int32)
                for n in ("box_pos", "box_quat", "box_half"):
                    setattr(cm, n, arr(n, getattr(m, n), np.float64))
        cm.nten = 0
        if tendons:
            dn = list(m.dof_names)
            cm.nten = len(tendons)
            ix = lambda d: d if isinstance(d, (int, np.integer)) else dn.index(d)
            cm.ten_dof = arr("ten_dof", [[ix(t["dofs"][0]), ix(t["dofs"][1])] for t in tendons], np.int32)
            cm.ten_coef = arr("ten_coef", [t["coefs"] for t in tendons], np.float64)
            cm.ten_range = arr("ten_range", [t["range"] for t in tendons], np.float64)
            cm.ten_k, cm.ten_d = tendon_k, tendon_d
        self.cm = cm
        self.nd, self.nb, self.ns = m.ndof, m.nb, len(m.sensor_body)

    def simulate(self, root, dof, tau=None, target=None, obj=None, obj_force=None):
        """In-place gym.simulate(): root (N,13), dof (N,nd,2).  Returns dict of derived outputs.  obj_force (N,3): force on the
        free object's body in its own frame, at the COM (apply_rigid_body_force_tensors LOCAL_SPACE, shadow_hand.py:708)."""
        N = root.shape[0]
        assert root.dtype == self.dtype and dof.dtype == self.dtype and root.flags.c_contiguous and dof.flags.c_contiguous
        tau = None if tau is None else np.ascontiguousarray(tau, dtype=self.dtype)
        target = None if target is None else np.ascontiguousarray(target, dtype=self.dtype)
        out = dict(body_state=np.zeros((N, self.nb, 13), self.dtype), contact_force=np.zeros((N, self.nb, 3), self.dtype),
                   sensor=np.zeros((N, max(self.ns, 1), 6), self.dtype), dof_force=np.zeros((N, max(self.nd, 1)), self.dtype))
        p = lambda a: None if a is None else C.c_void_p(a.ctypes.data)
        if obj is not None:
            assert obj.dtype == self.dtype and obj.flags.c_contiguous and obj.shape == (N, 13) and self.cm.obj_on
        if obj_force is not None:
            obj_force = np.ascontiguousarray(obj_force, dtype=self.dtype)
            assert obj is not None and obj_force.shape == (N, 3)
        self.lib.oracle_set_obj_force(p(obj_force))
        self.lib.oracle_simulate_obj(C.byref(self.cm), C.c_int(N), p(root), p(dof), p(tau), p(target),
                                     p(out["body_state"]), p(out["contact_force"]), p(out["sensor"]),
                                     p(out["dof_force"]), p(obj))
        self.lib.oracle_set_obj_force(None)
        out["sensor"] = out["sensor"][:, :self.ns]
        out["dof_force"] = out["dof_force"][:, :self.nd]
        return out

    def body_states(self, root, dof):
        N = root.shape[0]
        bs = np.zeros((N, self.nb, 13), self.dtype)
        self.lib.oracle_body_states(C.byref(self.cm), C.c_int(N), C.c_void_p(root.ctypes.data),
                                    C.c_void_p(dof.ctypes.data), C.c_void_p(bs.ctypes.data))
        return bs

    def forward_dynamics(self, root, dof, tau):
        """Single env, one sub-step: returns (qdd, root_after, dof_after)."""
        root = np.ascontiguousarray(root, self.dtype); dof = np.ascontiguousarray(dof, self.dtype)
        tau = np.ascontiguousarray(tau, self.dtype)
        qdd = np.zeros(max(self.nd, 1), self.dtype); ra = np.zeros(13, self.dtype); da = np.zeros_like(dof)
        self.lib.oracle_forward_dynamics(C.byref(self.cm), C.c_void_p(root.ctypes.data), C.c_void_p(dof.ctypes.data),
                                         C.c_void_p(tau.ctypes.data), C.c_void_p(qdd.ctypes.data),
                                         C.c_void_p(ra.ctypes.data), C.c_void_p(da.ctypes.data))
        return qdd[:self.nd], ra, da

    # ---- gym.acquire_jacobian_tensor / acquire_mass_matrix_tensor (franka_cube_stack.py:388-392)
    def jacobian_shape(self):
        """(body rows, 6, columns): a fixed base has no row for the base body and no base columns; a floating
        base has six leading columns (world linear, world angular velocity of the root origin)."""
        rows = int(self.lib.oracle_jacobian_rows(C.byref(self.cm)))
        return rows, 6, self.nd + (0 if self.cm.root_fixed else 6)

    def jacobian(self, root, dof):
        N = root.shape[0]
        root = np.ascontiguousarray(root, self.dtype); dof = np.ascontiguousarray(dof, self.dtype)
        J = np.zeros((N,) + self.jacobian_shape(), self.dtype)
        self.lib.oracle_jacobian(C.byref(self.cm), C.c_int(N), C.c_void_p(root.ctypes.data), C.c_void_p(dof.ctypes.data),
                                 C.c_void_p(J.ctypes.data))
        return J

    def mass_matrix(self, root, dof):
        N = root.shape[0]
        root = np.ascontiguousarray(root, self.dtype); dof = np.ascontiguousarray(dof, self.dtype)
        nc = self.nd + (0 if self.cm.root_fixed else 6)
        M = np.zeros((N, nc, nc), self.dtype)
        self.lib.oracle_mass_matrix(C.byref(self.cm), C.c_int(N), C.c_void_p(root.ctypes.data), C.c_void_p(dof.ctypes.data),
                                    C.c_void_p(M.ctypes.data))
        return M
