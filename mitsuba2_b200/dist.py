"""The one collective of the path: sum-reduce the per-rank XYZAW films to rank 0.

Pixel blocks are sharded over ranks (spiral block b -> rank b % world, B2RenderParams.rank /
world_size); nothing is exchanged while rendering. NCCL on the GPUs (bench.py), gloo in the
CPU tests. The tensor may alias the library's device film (zero copy)."""
import torch
import torch.distributed as dist


class CudaArray:
    """Wraps a raw device pointer for torch.as_tensor (zero copy)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


def film_tensor(ptr, height, width):
    return torch.as_tensor(CudaArray(ptr, (height, width, 5)), device="cuda")


def reduce_film(film, dst=0):
    """In-place sum of the rank films onto `dst` (other ranks' buffers are unspecified after)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM)
    return film


def owned_blocks(n_blocks, rank, world):
    """Spiral block indices rendered by `rank` (the partition b2_render_device applies)."""
    return [b for b in range(n_blocks) if b % world == rank]
