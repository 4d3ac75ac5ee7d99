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
    """(Region, bytes) of this rank.  Regions of all ranks tile the table in rank order; z-slabs differ in size when the world
    size does not divide the grid (voxb200_partition cuts z at floor(G * part / n))."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    return api.partition(gridsize, morton, rank, world)


def _region_sizes(local_table, group):
    """numel of every rank's region (they differ when the world size does not divide the grid)."""
    world = dist.get_world_size(group)
    mine = torch.tensor([local_table.numel()], dtype=torch.int64, device=local_table.device)
    sizes = torch.empty(world, dtype=torch.int64, device=local_table.device)
    dist.all_gather_into_tensor(sizes, mine, group=group)
    return [int(x) for x in sizes.tolist()]


def gather_table(local_table, group=None):
    """All-gather the per-rank regions into the full table on every rank.  Equal regions: one all_gather_into_tensor; unequal
    ones are padded to the largest for the collective and cut back afterwards."""
    world = dist.get_world_size(group)
    sizes = _region_sizes(local_table, group)
    local_table = local_table.contiguous()
    if len(set(sizes)) == 1:
        out = torch.empty(world * local_table.numel(), dtype=local_table.dtype, device=local_table.device)
        dist.all_gather_into_tensor(out, local_table, group=group)
        return out
    big = max(sizes)
    padded = torch.zeros(big, dtype=local_table.dtype, device=local_table.device)
    padded[: local_table.numel()] = local_table
    out = torch.empty(world * big, dtype=local_table.dtype, device=local_table.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * big: r * big + sizes[r]] for r in range(world)])


def gather_table_to(local_table, dst=0, group=None):
    """Gather the regions on rank ``dst`` only (returns None elsewhere); regions may differ in size."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = _region_sizes(local_table, group)
    big = max(sizes)
    send = local_table.contiguous()
    if send.numel() != big:
        send = torch.zeros(big, dtype=local_table.dtype, device=local_table.device)
        send[: local_table.numel()] = local_table
    bufs = [torch.empty(big, dtype=local_table.dtype, device=local_table.device) for _ in range(world)] if rank == dst else None
    dist.gather(send, bufs, dst=dst, group=group)
    return torch.cat([bufs[r][: sizes[r]] for r in range(world)]) if rank == dst else None


def exchange_routed(send, send_counts, group=None):
    """All-to-all of routed triangles.  ``send`` holds, back to back in rank order, the triangles (9 floats each)
    this rank routed to every rank; ``send_counts[r]`` = triangles for rank r.  Returns (recv, recv_counts): the
    triangles every rank routed to this one, in source-rank order.  NCCL moves them over NVLink."""
    dev = send.device
    sc = torch.tensor([int(c) for c in send_counts], dtype=torch.int64, device=dev)
    rc = torch.empty_like(sc)
    dist.all_to_all_single(rc, sc, group=group)
    recv_counts = [int(x) for x in rc.tolist()]
    recv = torch.empty(9 * sum(recv_counts), dtype=send.dtype, device=dev)
    dist.all_to_all_single(recv, send[: 9 * int(sum(send_counts))].contiguous(),
                           output_split_sizes=[9 * c for c in recv_counts], input_split_sizes=[9 * int(c) for c in send_counts], group=group)
    return recv, recv_counts


class ShardedHostVoxelizer:
    """End-to-end multi-GPU voxelization from HOST memory with the upload itself sharded: rank r holds (any) 1/N of the
    triangle soup in pinned memory, uploads only that, routes it on the GPU to the N regions
    (voxb200_route_triangles_multi), swaps triangles with the other ranks in one all-to-all over NVLink, voxelizes the
    triangles of its own region and copies its slab of the table back to the host.  Device buffers are kept between calls."""

    def __init__(self, grid, solid=False, morton=False):
        import copy
        self.grid, self.solid, self.morton = grid, solid, morton
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        g = grid.gridsize[0]
        self.regions = [api.partition(g, morton, r, self.world)[0] for r in range(self.world)]
        self.region_bytes = api.partition(g, morton, self.rank, self.world)[1]
        self.table = torch.empty(self.region_bytes // 4, dtype=torch.int32, device="cuda")
        self.d_chunk = None
        self.send = None
        self._copy = copy.copy
        self.ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    def __call__(self, host_chunk, host_slab):
        """host_chunk: pinned float32 CPU tensor [n_local * 9]; host_slab: pinned int32 CPU tensor of this rank's region.
        Returns device milliseconds from the first H2D byte to the last D2H byte."""
        n_local = host_chunk.numel() // 9
        if self.d_chunk is None or self.d_chunk.numel() < host_chunk.numel():
            self.d_chunk = torch.empty(host_chunk.numel(), dtype=torch.float32, device="cuda")
            self.send = torch.empty(2 * host_chunk.numel() + 9 * 1024, dtype=torch.float32, device="cuda")
        st = torch.cuda.current_stream()
        self.ev[0].record(st)
        d_chunk = self.d_chunk[: host_chunk.numel()]
        d_chunk.copy_(host_chunk, non_blocking=True)
        g_local = self._copy(self.grid)
        g_local.n_triangles = n_local
        # A triangle is written once per region it overlaps, so the routed soup can outgrow the 2x buffer.  That shows on one rank
        # only; ranks must not part ways in front of the all-to-all, so they agree on it first and grow together (n_local * world
        # triangles always fit).
        try:
            counts = api.route_triangles_multi(g_local, d_chunk, self.regions, self.send, solid=self.solid, morton=self.morton, stream=st)
            overflow = 0
        except api.VoxError as e:
            if "exceed the output capacity" not in str(e):
                raise
            counts, overflow = None, 1
        flag = torch.tensor([overflow], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        if int(flag.item()):
            need = self.world * host_chunk.numel() + 9 * 1024
            if self.send.numel() < need:
                self.send = torch.empty(need, dtype=torch.float32, device="cuda")
            counts = api.route_triangles_multi(g_local, d_chunk, self.regions, self.send, solid=self.solid, morton=self.morton, stream=st)
        recv, recv_counts = exchange_routed(self.send, counts)
        g_mine = self._copy(self.grid)
        g_mine.n_triangles = sum(recv_counts)
        fn = api.voxelize_solid if self.solid else api.voxelize
        fn(g_mine, recv, table=self.table, morton=self.morton, region=self.regions[self.rank], stream=st)
        host_slab.copy_(self.table, non_blocking=True)
        self.ev[1].record(st)
        st.synchronize()
        return self.ev[0].elapsed_time(self.ev[1])
