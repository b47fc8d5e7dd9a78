"""torchrun check (N GPUs): render_scene(world=...) == the single-GPU render of the whole job.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import light_transport_b200 as lt
from light_transport_b200 import api, scenes

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
rank, world = dist.get_rank(), dist.get_world_size()
objs, lights, cam, depth = scenes.lts_scene(api)
sc = api.Scene(camera=cam, lights=lights, width=256, height=256, max_depth=8, f_distance=depth, number_of_samples=1, allocate=False)
for trav in ("compat", "lbvh"):
    img, info = lt.render_scene(sc, objs, None, spp=64, rng="philox", seed=7, traversal=trav, world=True, return_info=True)
    paths = torch.tensor([float(info["paths"])], device="cuda")
    dist.all_reduce(paths)
    if rank == 0:
        ref = lt.render_scene(sc, objs, None, spp=64, rng="philox", seed=7, traversal=trav, device=0)
        err = np.abs(img - ref).max()
        print(f"world={world} {trav}: paths {int(paths[0])} (expect {256*256*64}), max |image(N ranks) - image(1 rank)| = {err:.3e}")
        assert int(paths[0]) == 256 * 256 * 64 and err < 1e-12
# every rank gets the image with all_ranks=True (ncclAllReduce inside lt_film_reduce)
img_all = lt.render_scene(sc, objs, None, spp=64, rng="philox", seed=7, traversal="compat", world=True, all_ranks=True)
chk = torch.tensor([float(np.abs(img_all).sum())], device="cuda", dtype=torch.float64)
lo, hi = chk.clone(), chk.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
assert float(lo) == float(hi) and img_all is not None
# error path: a bad root comes back as a status code, not a hang
from light_transport_b200 import parallel, host
comm = parallel.film_comm(local)
film = torch.zeros(16, dtype=torch.float64, device="cuda")
try:
    comm.reduce(film, root=world + 3)
    raise SystemExit("expected LT_E_INVALID")
except host.LtError as e:
    assert e.code == -1
if rank == 0:
    import ctypes as C
    v = C.c_int(); host.lib().lt_nccl_version(C.byref(v))
    print(f"world={world}: all_ranks image identical on every rank; lt_film_reduce on NCCL {v.value}; bad root -> LT_E_INVALID")
dist.barrier()
dist.destroy_process_group()
