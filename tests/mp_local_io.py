"""Run under torchrun with one rank per GPU (tests/test_gpu_multigpu.py::test_one_process_per_gpu_sharded_io):
sharded upload (each rank reads only its slice of the host array) + nbody_download_local against the full
nbody_download of the same state and against a single-GPU run."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mini_nbody_b200 as nb  # noqa: E402
import oracle_lib as orc  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
n = 40000                                                   # not a multiple of 128 * world
b = orc.randomize(n, 7)
ids = [nb.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
poisoned = b.copy()
with nb.NBody(n, rank=rank, world=world, device=local, nccl_id=ids[0]) as h:
    i0, i1 = h.info("i_begin"), h.info("i_end")
    # a rank must not depend on the other ranks' slices of ITS copy of the host array
    for k in poisoned.dtype.names:
        poisoned[k][:i0] = np.nan; poisoned[k][i1:] = np.nan
    h.upload(poisoned)
    h.step(0.01, 3)
    full = h.download()
    mine = h.download_local()
    assert len(mine) == i1 - i0
    for k in full.dtype.names:
        assert np.array_equal(full[k][i0:i1], mine[k]), k
    assert np.isfinite(full.view(np.float32)).all()
    parts = [None] * world
    dist.all_gather_object(parts, (i0, i1))
    assert parts[0][0] == 0 and parts[-1][1] == n and all(parts[r][1] == parts[r + 1][0] for r in range(world - 1))
    if rank == 0:
        with nb.NBody(n) as h1:                             # single-GPU run of the same state on this rank's device
            h1.upload(b); h1.step(0.01, 1); s1 = h1.download()
    h.upload(poisoned); h.step(0.01, 1); sg = h.download()
    if rank == 0:
        for comps in (("x", "y", "z"), ("vx", "vy", "vz")):
            dv = np.sqrt(sum((sg[k].astype(np.float64) - s1[k].astype(np.float64)) ** 2 for k in comps))
            nv = np.sqrt(sum(s1[k].astype(np.float64) ** 2 for k in comps))
            assert (dv / np.maximum(1.0, nv)).max() <= 3e-5, comps
dist.barrier()
if rank == 0:
    print("MP_LOCAL_IO_OK world=%d" % world)
