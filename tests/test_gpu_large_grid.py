"""GPU suite (-m gpu): grids above 4096^3.  An 8192^3 table is 64 GiB — too big for the oracle's host table — so the
checks run on z-slabs: the oracle clipped to the slab (writing into a slab-sized table) against the one-shot kernels with a region,
the prepared-mesh tile path with a region (32-bit region-relative addressing while absolute word indices exceed 2^32), and the host
entry point with the sparse read-back."""
import numpy as np
import pytest

import cases
import oracle

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
G = 8192


@pytest.fixture(scope="module")
def vb():
    import cuda_voxelizer_b200 as vb
    vb.init(0)
    return vb


@pytest.fixture(scope="module")
def scene(vb):
    """The 10M-triangle sphere of config 4 (radius 1024) placed around voxel (5000, 6000, 7000) of an 8192^3 grid of unit ~1."""
    v, f = cases.mesh("icosphere:708:1024")
    mn = np.array([-5000.0, -6000.0, -7000.0], np.float32)
    mx = mn + np.float32(G)
    grid = vb.make_grid(mn, mx, G, len(f))
    cmn, cmx, unit = np.zeros(3, np.float32), np.zeros(3, np.float32), np.zeros(3, np.float32)
    oracle.lib().oracle_bbox_cube(mn, mx, cmn, cmx)
    oracle.lib().oracle_unit(cmn, cmx, G, unit)
    assert np.array_equal(cmn, np.array(grid.bbox_min[:], np.float32)) and np.array_equal(unit, np.array(grid.unit[:], np.float32))
    soup = oracle.soup(v, f)
    return {"v": v, "f": f, "soup": soup, "d": torch.from_numpy(soup).cuda(), "grid": grid, "bb_min": cmn, "unit": unit}


@pytest.mark.parametrize("z0,z1", [(6992, 7056), (8000, 8032)])
def test_slab_of_8192_matches_the_oracle(vb, scene, z0, z1):
    from cuda_voxelizer_b200 import Region
    want = oracle.surface_slab(scene["soup"], scene["bb_min"], scene["unit"], G, z0, z1)
    assert oracle.popcount(want) > 100000
    region = Region((0, 0, z0), (G, G, z1))
    assert (G * G // 32) * z0 > 2**32                      # absolute word indices do not fit 32 bits here
    one_shot = vb.voxelize(scene["grid"], scene["d"], region=region)
    assert np.array_equal(one_shot.cpu().numpy().view(np.uint32), want)
    m = vb.Mesh(scene["grid"], tris=scene["d"], region=region)
    assert m.info()["tile_schedule"] == 1 and m.info()["instances"] > 100000
    dirty = torch.full_like(one_shot, -1)
    assert torch.equal(m.voxelize(table=dirty), one_shot)
    m.close()
    hv = torch.from_numpy(np.ascontiguousarray(scene["v"])).pin_memory()
    hf = torch.from_numpy(np.ascontiguousarray(scene["f"])).pin_memory()
    host = torch.full((want.size,), -1, dtype=torch.int32).pin_memory()
    vb.voxelize_host_indexed(scene["grid"], hv, hf, host, region=region)
    assert vb.last_readback()["sparse"]
    assert np.array_equal(host.numpy().view(np.uint32), want)
