"""world_size-2 gloo test of the N>1 host logic (no GPU): shard bookkeeping, the logging
all_gather in global env order, max-over-ranks timing, and that the reset RNG stream of a GLOBAL env
id does not depend on which rank owns it."""
import os
import socket
import numpy as np
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from isaacgymenvs_b200 import distributed as D
    from oracle import tasks_np as T
    r, l, w = D.init("gloo")
    n = 8
    off = D.env_id_offset(r, n)
    per_env = torch.arange(n, dtype=torch.float32) + 100.0 * r
    allr = D.gather_returns(per_env)
    tmax = D.max_over_ranks([1.0 + r, 5.0 - r])
    D.barrier()
    # reset stream of global env 11 (owned by rank 1 as local env 3), drawn through the owner's offset
    owner, local = D.owner_of(11, n)
    u = T.reset_uniforms(42, D.env_id_offset(owner, n) + local, 0, 4) if owner == r else None
    q.put((r, off, allr.numpy().tolist(), tmax, None if u is None else u.tolist()))
    torch.distributed.destroy_process_group()


def test_two_rank_gloo():
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    (r0, off0, g0, t0, u0), (r1, off1, g1, t1, u1) = res
    assert (off0, off1) == (0, 8)
    expect = [float(i) for i in range(8)] + [100.0 + i for i in range(8)]
    assert g0 == expect and g1 == expect                      # global env order on every rank
    assert t0 == [2.0, 5.0] and t1 == [2.0, 5.0]              # max over ranks
    from oracle import tasks_np as T
    assert u0 is None and np.allclose(u1, T.reset_uniforms(42, 11, 0, 4))   # sharding-independent stream
