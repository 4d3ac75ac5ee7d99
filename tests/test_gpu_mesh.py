"""GPU suite (-m gpu) of the prepared-mesh interface (voxb200_mesh_*): the tile-owner surface schedule and the
direct schedule must reproduce, bit for bit, the reference goldens / the oracle / the one-shot path."""
import copy

import numpy as np
import pytest

import cases
import oracle

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def vb():
    import cuda_voxelizer_b200 as vb
    vb.init(0)
    return vb


_cache = {}


def _mesh(name):
    if name not in _cache:
        _cache.clear()
        v, f = cases.mesh(name)
        _cache[name] = (v, f, torch.from_numpy(oracle.soup(v, f)).cuda())
    return _cache[name]


# every golden case whose grid the tile schedule covers (surface, linear, G % 256 == 0), configs 2 and 4 included
TILE_CASES = [c for c in cases.GOLDEN_CASES if c[1] % 256 == 0 and c[1] <= 2048 and not c[2] and not c[3]]


@pytest.mark.parametrize("name,g,solid,morton", TILE_CASES, ids=[cases.case_key(*c) for c in TILE_CASES])
def test_tile_schedule_matches_reference_golden(vb, golden, name, g, solid, morton):
    want = golden[cases.case_key(name, g, solid, morton)]
    v, f, d_tris = _mesh(name)
    grid = vb.grid_from_verts(v, g, len(f))
    m = vb.Mesh(grid, tris=d_tris)
    info = m.info()
    assert info["tile_schedule"] == 1 and info["tiles"] > 0
    dirty = torch.full((vb.table_bytes(g) // 4,), -1, dtype=torch.int32, device="cuda")     # every byte must be written
    table = m.voxelize(table=dirty)
    torch.cuda.synchronize()
    host = table.cpu().numpy().view(np.uint32)
    assert oracle.popcount(host) == want["popcount"]
    assert "%016x" % oracle.fnv1a64(host) == want["fnv1a64"]
    # again (the handle keeps no per-call state), and from the indexed mesh
    again = m.voxelize()
    assert torch.equal(again, table)
    m.close()
    mi = vb.Mesh(grid, verts=torch.from_numpy(np.ascontiguousarray(v)).cuda(), faces=torch.from_numpy(np.ascontiguousarray(f)).cuda())
    assert torch.equal(mi.voxelize(), table)
    mi.close()


@pytest.mark.parametrize("name,g", [("bunny", 256), ("icosphere:64:128", 256), ("torus:100:50:256", 256), ("soup:small:50000:6:512", 512),
                                    ("soup:mixed:2000:2:128", 256), ("soup:sliver:4000:4:256", 256), ("soup:axis:4000:5:256", 256)])
def test_tile_schedule_equals_one_shot_path(vb, name, g):
    v, f, d_tris = _mesh(name)
    grid = vb.grid_from_verts(v, g, len(f))
    want = vb.voxelize(grid, d_tris).clone()
    m = vb.Mesh(grid, tris=d_tris)
    assert m.info()["tile_schedule"] == 1
    assert torch.equal(m.voxelize(), want)
    # ACCUMULATE: OR into the caller's content
    gen = torch.Generator(device="cuda").manual_seed(3)
    prior = torch.randint(-2 ** 31, 2 ** 31 - 1, want.shape, dtype=torch.int32, device="cuda", generator=gen)
    t = prior.clone()
    m.voxelize(table=t, accumulate=True)
    assert torch.equal(t, prior | want)
    m.close()


@pytest.mark.parametrize("seed", range(6))
def test_tile_schedule_fuzz_vs_oracle(vb, seed):
    """Seeded soups of tiny to huge triangles (all three classes: <=3, 4, bigger) at 256^3, full-table compare with the oracle."""
    from cuda_voxelizer_b200 import meshgen
    g = 256
    kind = ("mixed", "small", "sliver", "axis", "mixed", "small")[seed]
    v, f = meshgen.random_soup(3000, 100 + seed, extent=float(g), kind=kind)
    # plus triangles of about one to four voxels (the classes the tiles own), some of them straddling tile borders
    rng = np.random.default_rng(200 + seed)
    n_small = 6000
    c = rng.uniform(2.0, g - 2.0, (n_small, 1, 3))
    c[: n_small // 4, 0, 1:] = np.round(c[: n_small // 4, 0, 1:] / 16.0) * 16.0       # centred on tile faces in y and z
    tiny = (c + rng.normal(0.0, (0.3, 0.6, 0.9, 1.2, 0.5, 1.5)[seed], (n_small, 3, 3))).reshape(-1, 3)
    v = np.vstack([v, np.clip(tiny, 0.0, float(g)).astype(np.float32)])
    f = np.arange(len(v), dtype=np.int32).reshape(-1, 3)
    soup = oracle.soup(v, f)
    mn, mx, unit = oracle.voxinfo(v, g)
    want = oracle.surface(soup, mn, unit, g, False)
    grid = vb.grid_from_verts(v, g, len(f))
    m = vb.Mesh(grid, tris=torch.from_numpy(soup).cuda())
    got = m.voxelize().cpu().numpy().view(np.uint32)
    diff = np.nonzero(got ^ want)[0]
    assert len(diff) == 0, "first differing words: %s (info %s)" % (diff[:8], m.info())
    m.close()


@pytest.mark.parametrize("n_parts", [2, 4, 8])
def test_tile_schedule_over_z_slabs(vb, n_parts):
    """Multi-GPU regions: handles prepared for the z-slabs of a partition concatenate to the whole table."""
    name, g = "icosphere:64:128", 256
    v, f, d_tris = _mesh(name)
    grid = vb.grid_from_verts(v, g, len(f))
    full = vb.voxelize(grid, d_tris).clone()
    parts = []
    for p in range(n_parts):
        region, nbytes = vb.partition(g, False, p, n_parts)
        m = vb.Mesh(grid, tris=d_tris, region=region)
        assert m.info()["tile_schedule"] == 1
        t = m.voxelize()
        assert t.numel() * 4 == nbytes
        parts.append(t.clone())
        m.close()
    assert torch.equal(torch.cat(parts), full)


@pytest.mark.parametrize("name,g,solid,morton", [("bunny", 128, 0, 0), ("bunny", 128, 1, 0), ("bunny", 64, 0, 1), ("bunny", 256, 1, 0),
                                                  ("bunny", 256, 0, 1), ("icosphere:16:64", 100, 0, 0), ("icosphere:64:128", 256, 1, 1)])
def test_direct_schedule_equals_one_shot_path(vb, name, g, solid, morton):
    """Solid, morton order and grid sizes the tiles do not cover run the one-shot kernels on the handle's own data."""
    v, f, d_tris = _mesh(name)
    grid = vb.grid_from_verts(v, g, len(f))
    fn = vb.voxelize_solid if solid else vb.voxelize
    want = fn(grid, d_tris, morton=bool(morton)).clone()
    m = vb.Mesh(grid, tris=d_tris, solid=bool(solid), morton=bool(morton))
    assert m.info()["tile_schedule"] == 0
    assert torch.equal(m.voxelize(), want)
    assert torch.equal(m.voxelize(), want)
    m.close()


def test_two_meshes_voxelize_concurrently_on_two_streams(vb, golden):
    """Re-entrancy: two handles, two streams, interleaved launches, both bit-exact against the goldens."""
    a_name, b_name, g = "icosphere:64:128", "bunny", 256
    va, fa = cases.mesh(a_name)
    vb_, fb = cases.mesh(b_name)
    ta = torch.from_numpy(oracle.soup(va, fa)).cuda()
    tb = torch.from_numpy(oracle.soup(vb_, fb)).cuda()
    ga, gb = vb.grid_from_verts(va, g, len(fa)), vb.grid_from_verts(vb_, g, len(fb))
    ma, mb = vb.Mesh(ga, tris=ta), vb.Mesh(gb, tris=tb)
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    outa = [torch.empty(vb.table_bytes(g) // 4, dtype=torch.int32, device="cuda") for _ in range(4)]
    outb = [torch.empty(vb.table_bytes(g) // 4, dtype=torch.int32, device="cuda") for _ in range(4)]
    torch.cuda.synchronize()
    for k in range(4):
        ma.voxelize(table=outa[k], stream=sa)
        mb.voxelize(table=outb[k], stream=sb)
    torch.cuda.synchronize()
    wa, wb = golden[cases.case_key(a_name, g, 0, 0)], golden[cases.case_key(b_name, g, 0, 0)]
    for k in range(4):
        ha, hb = outa[k].cpu().numpy().view(np.uint32), outb[k].cpu().numpy().view(np.uint32)
        assert "%016x" % oracle.fnv1a64(ha) == wa["fnv1a64"] and "%016x" % oracle.fnv1a64(hb) == wb["fnv1a64"]
    ma.close()
    mb.close()


def test_mesh_update_rekeys_moved_vertices(vb):
    """Per-frame use: new vertex positions through voxb200_mesh_update, same handle, table equals a fresh one-shot run."""
    name, g = "icosphere:64:128", 256
    v, f, d_tris = _mesh(name)
    grid = vb.grid_from_verts(v, g, len(f))
    m = vb.Mesh(grid, tris=d_tris)
    first = m.voxelize().clone()
    v2 = (v * np.float32(0.9) + np.float32(3.0)).astype(np.float32)        # deformed frame, same grid
    t2 = torch.from_numpy(oracle.soup(v2, f)).cuda()
    m.update(tris=t2)
    assert torch.equal(m.voxelize(), vb.voxelize(grid, t2))
    m.update(tris=d_tris)
    assert torch.equal(m.voxelize(), first)
    m.close()


def test_mesh_voxelize_is_graph_capturable(vb):
    name, g = "icosphere:64:128", 256
    v, f, d_tris = _mesh(name)
    grid = vb.grid_from_verts(v, g, len(f))
    m = vb.Mesh(grid, tris=d_tris)
    want = m.voxelize().clone()
    table = torch.zeros_like(want)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        m.voxelize(table=table)
    for _ in range(3):
        table.fill_(-1)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(table, want)
    m.close()


def test_over_full_tiles_take_the_side_path(vb):
    """More small triangles in one tile than a thread block should own: they are voxelized by the row-solver path instead."""
    rng = np.random.default_rng(5)
    n, g = 60000, 256
    base = rng.uniform(100.0, 110.0, (n, 1, 3)).astype(np.float32)            # all inside one 256x16x16 tile (after scaling)
    v = (base + rng.uniform(-0.7, 0.7, (n, 3, 3)).astype(np.float32)).reshape(-1, 3)
    v = np.vstack([v, np.array([[0, 0, 0], [256, 256, 256]], np.float32)])    # pin the grid to one unit per voxel
    f = np.arange(3 * n, dtype=np.int32).reshape(-1, 3)
    soup = oracle.soup(v, f)
    mn, mx, unit = oracle.voxinfo(v, g)
    want = oracle.surface(soup, mn, unit, g, False)
    grid = vb.grid_from_verts(v, g, len(f))
    m = vb.Mesh(grid, tris=torch.from_numpy(soup).cuda())
    info = m.info()
    assert info["side_triangles"] > 0
    got = m.voxelize().cpu().numpy().view(np.uint32)
    assert np.array_equal(got, want)
    m.close()


def test_mesh_rejects_bad_arguments(vb):
    v, f, d_tris = _mesh("bunny")
    grid = vb.grid_from_verts(v, 256, len(f))
    m = vb.Mesh(grid, tris=d_tris)
    with pytest.raises(vb.VoxError):
        m.voxelize(table=torch.empty(8, dtype=torch.int32, device="cuda")[1:])       # misaligned table
    bad = copy.copy(grid)
    bad.gridsize[1] = 128
    with pytest.raises(vb.VoxError):
        vb.Mesh(bad, tris=d_tris)
    m.close()


def test_out_of_range_face_indices_are_reported_not_read(vb):
    """Indexed input with a vertex index outside [0, n_verts): every entry point that takes faces fails with EINVAL (the indices are
    checked on the device while they are consumed, and clamped so that nothing is read out of bounds) and works again afterwards."""
    import torch
    v, f = cases.mesh("icosphere:64:128")
    g = 256
    grid = vb.grid_from_verts(v, g, len(f))
    good = vb.voxelize_host_indexed(grid, v, f)[0].copy()
    for bad_value in (len(v), -1, 2**31 - 1):
        fb = f.copy()
        fb[len(fb) // 2, 1] = bad_value
        with pytest.raises(vb.VoxError):
            vb.voxelize_host_indexed(grid, v, fb)                     # tile schedule
        with pytest.raises(vb.VoxError):
            vb.voxelize_host_indexed(grid, v, fb, solid=True)         # expansion + one-shot kernels
        with pytest.raises(vb.VoxError):
            vb.upload_indexed(v, fb)
        with pytest.raises(vb.VoxError):
            vb.Mesh(grid, verts=torch.from_numpy(np.ascontiguousarray(v)).cuda(), faces=torch.from_numpy(np.ascontiguousarray(fb)).cuda())
        with pytest.raises(vb.VoxError):
            vb.Mesh(grid, verts=torch.from_numpy(np.ascontiguousarray(v)).cuda(), faces=torch.from_numpy(np.ascontiguousarray(fb)).cuda(), solid=True)
        with pytest.raises(vb.VoxError):
            vb.voxelize_host_multi(grid, np.ascontiguousarray(v), np.ascontiguousarray(fb), n_devices=1)
        with pytest.raises(vb.VoxError):
            vb.voxelize_host_multi(grid, np.ascontiguousarray(v), np.ascontiguousarray(fb), solid=True, n_devices=1)
    assert np.array_equal(vb.voxelize_host_indexed(grid, v, f)[0], good)
