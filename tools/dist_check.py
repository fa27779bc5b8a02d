"""Run under torchrun (one rank per GPU): the spp-sharded NCCL render must equal the single-GPU render of the same
sample set (same Philox streams) up to fp32 summation order.  Rank 0 prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pkg = graft.load_package()
spec = pkg.make_scene(sys.argv[1] if len(sys.argv) > 1 else "cornell-smoke", 1)
w, h, spp, depth, seed = 96, 96, 203, 50, 17  # 203: not divisible by the world size
dr = pkg.renderer.DistributedRenderer(spec, w, h, device=local, rank=rank, world_size=world)
dr.step(spp, depth, seed, to_host=True)
torch.cuda.synchronize()
summed = dr.accum.cpu().numpy()
ok = True
msg = {}
if rank == 0:
    full, st = dr.r.render(w, h, spp, depth, seed=seed, resolve=False)
    img, _ = dr.r.render(w, h, spp, depth, seed=seed, resolve=True)
    a, b = summed[..., :3], full[..., :3]
    rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-3)
    msg = {"world": world, "counts_ok": bool(np.all(summed[..., 3] == spp)), "max_rel_diff": float(rel.max()),
           "resolved_max_abs_diff": float(np.nanmax(np.abs(dr.host.numpy()[..., :3] - img[..., :3])))}
    ok = msg["counts_ok"] and msg["max_rel_diff"] < 1e-3 and msg["resolved_max_abs_diff"] < 1e-3
    msg["ok"] = bool(ok)
    print(json.dumps(msg), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
