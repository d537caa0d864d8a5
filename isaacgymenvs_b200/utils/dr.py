"""Domain randomisation of the reference (`tasks/base/vec_task.py:610-840`, `utils/dr_utils.py:71-208`).

* observations / actions: the two "non-physical" entries of `task.randomization_params` -- noise lambdas on device tensors.
* actor_params: PHYSICAL properties.  The reference rewrites them actor by actor through `gym.set_actor_*_properties` in
  a Python loop over the envs being reset (vec_task.py:752-828); here they are per-env PARAMETER ARRAYS the step kernel
  reads (include/b200gym.h B2G_T_ENV_MASS_SCALE / ENV_DOF_PROPS / ENV_FRICTION), refreshed on the device for the envs
  that are about to reset and whose randomisation counter passed `frequency` -- the reference's selection rule
  (vec_task.py:631-637).  Supported: rigid_body_properties.mass, dof_properties.{damping, stiffness, lower, upper},
  rigid_shape_properties.friction; `sim_params` (gravity) and anything else raise.

A NoiseModel is built from one YAML entry

    observations: {range: [0, .002], range_correlated: [0, .001], operation: additive, distribution: gaussian,
                   schedule: linear, schedule_steps: 40000}

and the simulation frame count (the schedule ramps the noise in).  Calling it perturbs a tensor with a per-step white
part and a correlated part whose unit sample is drawn once, the first time, and kept (`params['corr']` in the reference),
drawing from torch's global generator in the reference's order, so seeded runs reproduce the reference's numbers.
"""
import operator

import torch


def _schedule(entry, frame):
    kind = entry.get("schedule")
    if kind == "linear":
        return min(frame, entry["schedule_steps"]) / entry["schedule_steps"]
    if kind == "constant":
        return 0 if frame < entry["schedule_steps"] else 1
    return 1


class NoiseModel:
    def __init__(self, entry, frame, carry=None):
        self.dist = entry["distribution"]
        if self.dist not in ("gaussian", "uniform"):
            raise ValueError(f"unsupported noise distribution {self.dist!r} (gaussian | uniform)")
        self.additive = entry["operation"] == "additive"
        self.op = operator.add if self.additive else operator.mul
        a, b = entry["range"]
        ac, bc = entry.get("range_correlated", [0., 0.])
        s = _schedule(entry, frame)
        blend = (lambda v: v * s) if self.additive else (lambda v: v * s + 1.0 * (1.0 - s))
        if self.dist == "gaussian":                  # (mean, spread): under "scaling" only the mean is blended towards 1
            self.white = (blend(a), b * s)
            self.corr = (blend(ac), bc * s)
        else:                                        # (low, high): both ends blended
            self.white = (blend(a), blend(b))
            self.corr = (blend(ac), blend(bc))
        self.unit_corr = carry                       # the correlated part's unit sample survives re-parameterisation? no: the
        #                                              reference rebuilds its dict, dropping 'corr' -- carry stays None there

    def __call__(self, tensor):
        if self.unit_corr is None:
            self.unit_corr = torch.randn_like(tensor)
        if self.dist == "gaussian":
            mu, var = self.white
            mu_c, var_c = self.corr
            corr = self.unit_corr * var_c + mu_c
            return self.op(tensor, corr + torch.randn_like(tensor) * var + mu)
        lo, hi = self.white
        lo_c, hi_c = self.corr
        corr = self.unit_corr * (hi_c - lo_c) + lo_c
        return self.op(tensor, corr + torch.rand_like(tensor) * (hi - lo) + lo)


class Randomizer:
    """Frequency gating of `apply_randomizations` (vec_task.py:619-640) for the non-physical parameters."""

    def __init__(self, dr_params):
        bad = [k for k in dr_params if k not in ("frequency", "observations", "actions", "actor_params")]
        if bad:
            raise NotImplementedError(f"domain randomisation of {bad} is not provided (observations, actions, actor_params)")
        self.params = dr_params
        self.freq = dr_params.get("frequency", 1)
        self.first = True
        self.last_rand_frame = 0
        self.models = {}

    def update(self, frame):
        """Call before a step with the simulation frame count; re-parameterises the noise when the frequency elapsed."""
        due = self.first or (frame - self.last_rand_frame) >= self.freq
        if due:
            self.last_rand_frame = frame
            for key in ("observations", "actions"):
                if key in self.params:
                    self.models[key] = NoiseModel(self.params[key], frame)
        self.first = False
        return due


def _sched(entry, frame):
    return _schedule(entry, frame)


def _sample(entry, shape, frame, device, gen=None):
    """generate_random_samples, utils/dr_utils.py:71-131 (torch on the device instead of numpy on the host)."""
    a, b = entry["range"]
    s = _sched(entry, frame)
    additive = entry["operation"] == "additive"
    blend = (lambda v: v * s) if additive else (lambda v: v * s + 1.0 * (1.0 - s))
    dist = entry["distribution"]
    if dist == "gaussian":
        mu, var = blend(a), b * s
        return torch.randn(shape, device=device, generator=gen) * var + mu
    lo, hi = blend(a), blend(b)
    if dist == "loguniform":
        import math
        return torch.exp(torch.rand(shape, device=device, generator=gen) * (math.log(hi) - math.log(lo)) + math.log(lo))
    if dist == "uniform":
        return torch.rand(shape, device=device, generator=gen) * (hi - lo) + lo
    raise ValueError(f"unsupported distribution {dist!r}")


class PhysicalRandomizer:
    """`actor_params` of one actor type as per-env parameter tensors (see the module docstring)."""
    SUPPORTED = {"rigid_body_properties": ("mass",), "dof_properties": ("damping", "stiffness", "lower", "upper"),
                 "rigid_shape_properties": ("friction",)}

    def __init__(self, actor_params, model, num_envs, device, frequency):
        if len(actor_params) != 1:
            raise NotImplementedError("actor_params: exactly one actor type per environment")
        (self.actor, props), = actor_params.items()
        self.props = {k: v for k, v in props.items() if k not in ("color", "scale")}
        for group, attrs in self.props.items():
            if group not in self.SUPPORTED:
                raise NotImplementedError(f"actor_params.{self.actor}.{group} is not provided")
            for attr in attrs:
                if attr not in self.SUPPORTED[group]:
                    raise NotImplementedError(f"actor_params.{self.actor}.{group}.{attr} is not provided")
        self.freq, self.N, self.device, self.first = frequency, num_envs, device, True
        f = lambda a: torch.tensor(a, dtype=torch.float32, device=device)
        nl, nd = model.nl, model.ndof
        import numpy as np
        lim = np.asarray(model.limited[1:]) > 0
        self.og_dof = torch.stack([f(model.damping[1:]), f(model.stiffness[1:]), f(np.where(lim, model.lower[1:], -3e38)),
                                   f(np.where(lim, model.upper[1:], 3e38))], -1)            # (nd, 4)
        self.limited = torch.tensor(lim, device=device)
        self.og_friction = float(np.asarray(model.cp_mu)[0]) if len(model.cp_mu) else 1.0
        self.mass_scale = torch.ones(num_envs, nl, device=device)
        self.dof_props = self.og_dof.unsqueeze(0).repeat(num_envs, 1, 1).contiguous()
        self.friction = torch.full((num_envs,), self.og_friction, device=device)
        self.uses = {g: g in self.props for g in self.SUPPORTED}

    def tensors(self, E):
        """slot -> tensor for the groups that are randomised."""
        out = {}
        if self.uses["rigid_body_properties"]:
            out[E.T_ENV_MASS_SCALE] = self.mass_scale
        if self.uses["dof_properties"]:
            out[E.T_ENV_DOF_PROPS] = self.dof_props
        if self.uses["rigid_shape_properties"]:
            out[E.T_ENV_FRICTION] = self.friction
        return out

    @torch.no_grad()
    def apply(self, frame, randomize_buf, reset_buf):
        """Re-sample the parameters of the envs selected by the reference's rule: all on the first call, afterwards the
        envs flagged for reset whose counter reached `frequency` (their counter restarts)."""
        if self.first:
            mask = torch.ones(self.N, dtype=torch.bool, device=self.device)
        else:
            mask = (randomize_buf >= self.freq) & (reset_buf != 0)
            randomize_buf[mask] = 0
        N = self.N
        for group, attrs in self.props.items():
            for attr, entry in attrs.items():
                if entry.get("setup_only", False) and not self.first:
                    continue
                scaling = entry["operation"] == "scaling"
                if group == "rigid_body_properties":                       # mass, per body: here per link
                    smp = _sample(entry, self.mass_scale.shape, frame, self.device)
                    if not scaling:
                        raise NotImplementedError("rigid_body_properties.mass: scaling only (the kernels take a factor)")
                    self.mass_scale.copy_(torch.where(mask[:, None], smp, self.mass_scale))
                elif group == "dof_properties":
                    col = ("damping", "stiffness", "lower", "upper").index(attr)
                    og = self.og_dof[:, col]
                    smp = _sample(entry, (N, og.shape[0]), frame, self.device)
                    new = og[None] * smp if scaling else og[None] + smp
                    if col >= 2:
                        new = torch.where(self.limited[None], new, og[None])  # unlimited joints stay unlimited
                    self.dof_props[:, :, col].copy_(torch.where(mask[:, None], new, self.dof_props[:, :, col]))
                else:                                                        # friction: one value per env (all its shapes)
                    smp = _sample(entry, (N,), frame, self.device)
                    new = self.og_friction * smp if scaling else self.og_friction + smp
                    self.friction.copy_(torch.where(mask, new, self.friction))
        self.first = False
