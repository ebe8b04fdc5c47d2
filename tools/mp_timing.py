"""torchrun helper: time upload / step / download separately for both exchange modes (2+ ranks)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import mini_nbody_b200 as nb
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
ids = [nb.nccl_unique_id() if rank == 0 else None]; dist.broadcast_object_list(ids, src=0)
h = nb.NBody(n, nb.F32, rank=rank, world=world, device=lr, nccl_id=ids[0])
blobs = [None] * world; dist.all_gather_object(blobs, h.ipc_export()); h.ipc_import(blobs)
host = nb.randomizeBodies(n, 42)
for mode in (0, 1, 0, 1):
    h.set_option("exchange", mode)
    h.upload(host); h.step(0.01, 1); h.download(host)
    dist.barrier(); torch.cuda.synchronize()
    t = {}
    for name, fn in (("upload", lambda: h.upload(host)), ("step", lambda: h.step(0.01, 1)), ("download", lambda: h.download(host)),
                     ("upload2", lambda: h.upload(host)), ("step3", lambda: h.step(0.01, 3)), ("accel", lambda: h.accel())):
        t0 = time.perf_counter(); fn(); t[name] = round((time.perf_counter() - t0) * 1e3, 2)
    print(json.dumps({"rank": rank, "exchange": mode, **t, "step_ms_dev": h.last_step_ms()}), flush=True)
    dist.barrier()
h.close(); dist.destroy_process_group()
