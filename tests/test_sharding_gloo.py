"""CPU suite: the N>1 host logic on the gloo backend, world_size 2, 3 (slabs of unequal size: 21 / 21 / 22 layers of 64) and 4.  Each rank derives its region through
the product's partition logic, fills it with the ORACLE restricted to that region (the checker standing in
for the GPU kernel, which cannot run here), and the product's gather must reproduce the oracle's full table."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, solid, result_path):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "1"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import cases
        import oracle
        from cuda_voxelizer_b200 import sharding
        import cuda_voxelizer_b200 as vb
        g = 64
        v, f = cases.mesh("bunny")
        soup = oracle.soup(v, f)
        grid = vb.grid_from_verts(v, g, len(f))
        bb_min, unit = np.array(grid.bbox_min[:], np.float32), np.array(grid.unit[:], np.float32)
        region, nbytes = sharding.owned_region(g, False)
        z0, z1 = region.lo[2], region.hi[2]
        fn = oracle.solid if solid else oracle.surface
        full_local = fn(soup, bb_min, unit, g, z_range=(z0, z1))
        words = g * g // 32
        mine = torch.from_numpy(full_local[z0 * words: z1 * words].astype(np.int32))
        assert mine.numel() * 4 == nbytes
        gathered = sharding.gather_table(mine).numpy().view(np.uint32)
        want = fn(soup, bb_min, unit, g)
        ok = bool(np.array_equal(gathered, want))
        # all-to-all of routed triangles: rank r sends r+1 triangles tagged (r, dst) to every dst
        send_counts = [rank + 1] * world
        send = torch.cat([torch.full((9 * (rank + 1),), float(100 * rank + dst)) for dst in range(world)])
        recv, recv_counts = sharding.exchange_routed(send, send_counts)
        ok = ok and recv_counts == [src + 1 for src in range(world)]
        want_recv = torch.cat([torch.full((9 * (src + 1),), float(100 * src + rank)) for src in range(world)])
        ok = ok and bool(torch.equal(recv, want_recv))
        on0 = sharding.gather_table_to(mine, dst=0)
        if rank == 0:
            ok = ok and bool(np.array_equal(on0.numpy().view(np.uint32), want))
        flag = torch.tensor([1 if ok else 0])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            with open(result_path, "w") as fh:
                fh.write("ok" if int(flag) == 1 else "mismatch")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("solid", [False, True])
def test_slab_sharding_and_gather_on_gloo(tmp_path, world, solid):
    result = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(world, _free_port(), solid, result), nprocs=world, join=True)
    assert open(result).read() == "ok"
