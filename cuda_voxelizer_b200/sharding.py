"""Multi-GPU plumbing: one process per GPU, each owning a disjoint region of the bit table —
a z-slab in linear order, an aligned curve segment in morton order.  Voxel values depend only on the triangle
list, never on other voxels, so there is NO reduction and no data-path collective: every rank voxelizes the
triangles routed to its region and the full table is the plain concatenation of the regions in rank order.
The only exchange is that final gather (NCCL all-gather over NVLink, or rank-wise copies into one host table).

torch.distributed is used for the plumbing only; the same code runs on the gloo backend with CPU tensors, which
is how the CPU test-suite covers it (tests/test_sharding_gloo.py).
"""
import torch
import torch.distributed as dist

from . import api


def owned_region(gridsize, morton, rank=None, world=None):
    """(Region, bytes) of this rank.  Regions of all ranks tile the table in rank order and have equal size."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    return api.partition(gridsize, morton, rank, world)


def gather_table(local_table, group=None):
    """All-gather the per-rank regions into the full table on every rank (equal-sized regions)."""
    world = dist.get_world_size(group)
    out = torch.empty(world * local_table.numel(), dtype=local_table.dtype, device=local_table.device)
    dist.all_gather_into_tensor(out, local_table.contiguous(), group=group)
    return out


def gather_table_to(local_table, dst=0, group=None):
    """Gather the regions on rank ``dst`` only (returns None elsewhere)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    bufs = [torch.empty_like(local_table) for _ in range(world)] if rank == dst else None
    dist.gather(local_table.contiguous(), bufs, dst=dst, group=group)
    return torch.cat(bufs) if rank == dst else None
