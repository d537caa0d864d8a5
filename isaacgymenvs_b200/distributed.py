"""One process per GPU (`torchrun`), envs sharded with no data-path collective.

The reference's multi-GPU mode is N independent replicas picked by LOCAL_RANK
(`isaacgymenvs/utils/rlgames_utils.py:89-107`, seed offset by rank `utils/utils.py:89-94`).  Here
rank r owns the contiguous block of global env ids [r*n, (r+1)*n); the reset RNG is keyed by the
GLOBAL id so a rollout does not depend on how the envs were sharded.  The only collectives are for
logging: per-env returns gathered once per rollout, and the max-over-ranks of timings.
"""
import os
import torch
import torch.distributed as dist


def rank_info():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init(backend=None):
    rank, local, world = rank_info()
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": torch.device(f"cuda:{local}")} if backend == "nccl" else {}
        dist.init_process_group(backend, **kw)
    return rank, local, world


def env_id_offset(rank, envs_per_rank):
    return rank * envs_per_rank


def owner_of(global_env_id, envs_per_rank):
    return global_env_id // envs_per_rank, global_env_id % envs_per_rank


def gather_returns(per_env: torch.Tensor) -> torch.Tensor:
    """all_gather of this rank's (n,) per-env returns -> (world*n,) in global env order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return per_env.clone()
    out = [torch.empty_like(per_env) for _ in range(dist.get_world_size())]
    dist.all_gather(out, per_env.contiguous())
    return torch.cat(out)


def max_over_ranks(values, device=None) -> list:
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.cpu()]


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
