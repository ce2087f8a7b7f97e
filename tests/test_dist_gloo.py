"""world_size-2 gloo test of the multi-GPU host logic on CPU: rank-sharded renders (same
rank/world_size partition the CUDA library applies, here executed by the oracle) summed with
mitsuba2_b200.dist.reduce_film must equal the single-process film exactly (box filter: the
rank films have disjoint support)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mitsuba2_b200 import abi
    from mitsuba2_b200.dist import reduce_film
    from tests import golden_cases as G, oracle_lib as O
    sb, rp, _ = G.load("cbox_80x48_spp8", O.OracleHost())
    rpr = abi.render_params(rp.max_depth, rp.rr_depth, rp.spp, rp.samples_per_pass, rp.block_size, rp.seed, rank, world)
    part, st = O.render(sb, rpr, threads=1)
    t = torch.from_numpy(part.copy())
    reduce_film(t, dst=0)
    n = torch.tensor([st["n_samples"]], dtype=torch.int64)
    dist.reduce(n, dst=0)
    if rank == 0:
        np.savez(out_path, film=t.numpy(), n=n.numpy())
    dist.destroy_process_group()


def test_two_rank_reduce_equals_single(tmp_path):
    sys.path.insert(0, ROOT)
    from tests import golden_cases as G, oracle_lib as O
    O.lib()
    out = str(tmp_path / "film.npz")
    mp.spawn(_worker, args=(2, 29531 + os.getpid() % 1000, out), nprocs=2, join=True)
    sb, rp, _ = G.load("cbox_80x48_spp8", O.OracleHost())
    full, st = O.render(sb, rp, threads=1)
    z = np.load(out)
    assert int(z["n"][0]) == st["n_samples"]
    assert np.array_equal(z["film"].view(np.uint32), full.view(np.uint32))


def test_owned_blocks_partition():
    from mitsuba2_b200.dist import owned_blocks
    for world in (1, 2, 3, 8):
        seen = sorted(b for r in range(world) for b in owned_blocks(37, r, world))
        assert seen == list(range(37))
