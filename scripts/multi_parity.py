"""torchrun --nproc-per-node N scripts/multi_parity.py: the N-GPU render through the public API (b2_render_reduce with an NCCL
reduce of the device films in the callback) against the 1-GPU render of the same scene on rank 0.
box filter: the ranks' pixel blocks are disjoint, every other rank contributes exact zeros -> films bit-identical;
gaussian filter: the 2-pixel borders of neighbouring blocks overlap, fp32 sums in another order -> last-ulp differences
(what the reference's own thread-count dependence amounts to, SURVEY.md 8c 'Determinism')."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mitsuba2_b200 import api, scenes  # noqa: E402
from mitsuba2_b200.dist import CudaArray  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))

    def reduce_film(ptr, n, stream):
        film = torch.as_tensor(CudaArray(ptr, (n,)), device="cuda")
        dist.reduce(film, dst=0, op=dist.ReduceOp.SUM)

    H = api.Host()
    out = []
    cases = [("cbox_box", lambda: scenes.cbox(H, 256, 256), 16), ("cbox_gaussian", lambda: scenes.cbox(H, 256, 256, rfilter="gaussian"), 16),
             ("veach_box", lambda: scenes.veach_mis(H, 320, 180), 16), ("grid_box", lambda: scenes.instanced_grid(H, 256, 256, grid=21, lattice=4), 8)]
    for name, make, spp in cases:
        sc = api.Scene(make())
        integ = api.PathIntegrator(max_depth=6, rr_depth=5, block_size=32)
        film = integ.render_reduce(sc, spp, reduce_film, rank=rank, world_size=world)
        torch.cuda.synchronize(); dist.barrier()
        if rank == 0:
            ref = integ.render(sc, spp)
            same = bool(np.array_equal(film.view(np.uint32), ref.view(np.uint32)))
            rel = float(np.linalg.norm(film.astype(np.float64) - ref) / np.linalg.norm(ref.astype(np.float64)))
            out.append(dict(case=name, world_size=world, bit_identical=same, rel_l2=rel, weight_sum=float(film[..., 4].sum()), expect_weight=float(ref[..., 4].sum())))
            if "box" in name.split("_")[1]:
                assert same, (name, rel)
            else:
                assert rel < 1e-6, (name, rel)
        sc.close()
    if rank == 0:
        print(json.dumps(dict(multi_gpu_parity=out)))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
