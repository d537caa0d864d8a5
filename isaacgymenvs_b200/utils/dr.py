"""Observation / action noise of the reference's domain randomisation (`tasks/base/vec_task.py:648-718`): the two
"non-physical" entries of `task.randomization_params`.  Physical randomisation (sim_params, actor_params) rewrites
simulator properties through per-actor gym calls and is not provided.

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
        bad = [k for k in dr_params if k not in ("frequency", "observations", "actions")]
        if bad:
            raise NotImplementedError(f"physical domain randomisation {bad} is not provided (observations / actions noise only)")
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
