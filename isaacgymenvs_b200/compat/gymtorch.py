"""`isaacgym.gymtorch`: tensors are torch tensors already (ant.py:82-95, :284-285)."""


def wrap_tensor(desc):
    return desc.tensor if hasattr(desc, "tensor") else desc


def unwrap_tensor(t):
    return t
