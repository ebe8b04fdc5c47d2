"""The N>1 path on CPU: world_size-2 torch.distributed (gloo).  Each rank owns the i-range
nbody_plan() gives it, evaluates the force on it with the oracle standing in for the CUDA kernel,
integrates its slice and all-gathers the new positions -- the per-step exchange of the real path
(one all-gather of N/G x 3 scalars per rank, in place into the replicated position array).  The
result must equal a single-rank run bit for bit."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, steps, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import mini_nbody_b200 as nb
    import oracle_lib as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = nb.plan(n, rank=rank, world=world)
    i0, i1 = p["i_begin"], p["i_end"]
    assert (i0, i1) == nb.shard_range(n, rank, world)
    blk, lb = p["blk"], p["local_blocks"]
    b = orc.randomize(n, 42)
    dt = np.float32(0.01)
    for _ in range(steps):
        a = orc.accel_f32(b, i0, i1)                                     # this rank's i-slice against ALL j
        for c, (xk, vk) in enumerate(zip("xyz", ("vx", "vy", "vz"))):
            v = (b[vk][i0:i1].astype(np.float64) + np.float64(dt) * a[:, c].astype(np.float64)).astype(np.float32)
            b[vk][i0:i1] = v
            b[xk][i0:i1] = (b[xk][i0:i1].astype(np.float64) + v.astype(np.float64) * np.float64(dt)).astype(np.float32)
        # exchange: equal-sized padded slices, rank-major (the layout ncclAllGather writes in place)
        send = torch.zeros(lb * blk, 3)
        send[: i1 - i0] = torch.from_numpy(np.stack([b["x"][i0:i1], b["y"][i0:i1], b["z"][i0:i1]], axis=1))
        recv = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(recv, send)
        full = torch.cat(recv)[:n].numpy()
        b["x"], b["y"], b["z"] = full[:, 0], full[:, 1], full[:, 2]
    # velocities are only gathered on download
    sendv = torch.zeros(lb * blk, 3)
    sendv[: i1 - i0] = torch.from_numpy(np.stack([b["vx"][i0:i1], b["vy"][i0:i1], b["vz"][i0:i1]], axis=1))
    recvv = [torch.empty_like(sendv) for _ in range(world)]
    dist.all_gather(recvv, sendv)
    fullv = torch.cat(recvv)[:n].numpy()
    b["vx"], b["vy"], b["vz"] = fullv[:, 0], fullv[:, 1], fullv[:, 2]
    if rank == 0:
        q.put(b.view(np.float32).copy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1000, 2048])
def test_two_rank_sharded_step_equals_single_rank(built, n):
    import torch.multiprocessing as mp
    import oracle_lib as orc
    steps, world = 2, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n % 7
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = orc.run(orc.randomize(n, 42), 0.01, steps).view(np.float32)
    np.testing.assert_array_equal(got, ref)
