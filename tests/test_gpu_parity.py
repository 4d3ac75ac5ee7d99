"""GPU suite (-m gpu): the CUDA path, called through the C ABI, against
  (1) the golden vectors generated from the reference itself (tests/golden/golden.json),
  (2) the oracle restatement on the same seeded inputs (full-table compare at small sizes),
  (3) size-independent properties at BASELINE.json's full sizes.
Bit-exact everywhere: this is integer/bit output; there is no tolerance."""
import numpy as np
import pytest

import cases
import oracle

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def vb():
    import cuda_voxelizer_b200 as vb
    vb.init(0)          # raises (no fallback) when the library or a B200 is missing
    return vb


_mesh_cache = {}


def _device_mesh(name):
    if name not in _mesh_cache:
        _mesh_cache.clear()
        v, f = cases.mesh(name)
        soup = oracle.soup(v, f)           # the reference's own 9-float layout (main.cpp:61-80)
        _mesh_cache[name] = (v, f, torch.from_numpy(soup).cuda())
    return _mesh_cache[name]


def _run(vb, name, g, solid, morton, **kw):
    v, f, d_tris = _device_mesh(name)
    grid = vb.grid_from_verts(v, g, len(f))
    fn = vb.voxelize_solid if solid else vb.voxelize
    table = fn(grid, d_tris, morton=bool(morton), **kw)
    torch.cuda.synchronize()
    return table.cpu().numpy().view(np.uint32), grid


@pytest.mark.parametrize("name,g,solid,morton", cases.GOLDEN_CASES, ids=[cases.case_key(*c) for c in cases.GOLDEN_CASES])
def test_matches_reference_golden(vb, golden, name, g, solid, morton):
    want = golden[cases.case_key(name, g, solid, morton)]
    table, grid = _run(vb, name, g, solid, morton)
    assert [float(x) for x in grid.unit] == want["unit"]
    assert table.nbytes == vb.table_bytes(g)
    assert oracle.popcount(table) == want["popcount"]
    assert "%016x" % oracle.fnv1a64(table) == want["fnv1a64"]
    assert vb.last_counters()["solid_clamped"] == 0


ORACLE_CASES = [c for c in cases.GOLDEN_CASES if c[1] <= 256]


@pytest.mark.parametrize("name,g,solid,morton", ORACLE_CASES, ids=[cases.case_key(*c) for c in ORACLE_CASES])
def test_full_table_equals_oracle(vb, name, g, solid, morton):
    table, grid = _run(vb, name, g, solid, morton)
    v, f, _ = _device_mesh(name)
    mn, mx, unit = oracle.voxinfo(v, g)
    want = (oracle.solid if solid else oracle.surface)(oracle.soup(v, f), mn, unit, g, morton)
    diff = np.nonzero(table ^ want)[0]
    assert len(diff) == 0, "first differing words: %s" % diff[:8]


@pytest.mark.parametrize("g", [8, 16, 24, 40, 100])
@pytest.mark.parametrize("solid", [0, 1])
def test_odd_and_tiny_grids(vb, g, solid):
    """Grid sizes whose rows are not whole words (and non powers of two) take the generic paths."""
    name = "icosphere:16:64"
    table, grid = _run(vb, name, g, solid, 0)
    v, f, _ = _device_mesh(name)
    mn, mx, unit = oracle.voxinfo(v, g)
    want = (oracle.solid if solid else oracle.surface)(oracle.soup(v, f), mn, unit, g, 0)
    assert np.array_equal(table, want)


@pytest.mark.parametrize("g", [257, 513])
@pytest.mark.parametrize("solid", [0, 1])
def test_box_reaches_the_last_table_bit(vb, g, solid):
    """At these sizes the reference's ceil(G^3/32.0f)*4 is a bit short of G^3 voxels: voxb200_table_bytes must hold
    the far-corner voxel, and the kernels must set it like the oracle does."""
    name = "box:10"
    table, grid = _run(vb, name, g, solid, 0)
    assert table.nbytes == vb.table_bytes(g) and table.nbytes * 8 >= g ** 3
    v, f, _ = _device_mesh(name)
    mn, mx, unit = oracle.voxinfo(v, g)
    want = (oracle.solid if solid else oracle.surface)(oracle.soup(v, f), mn, unit, g, 0)
    assert np.array_equal(table, want)
    if not solid:
        last = g ** 3 - 1
        assert (int(table[last // 32]) >> (31 - last % 32)) & 1 == 1


@pytest.mark.parametrize("name,g,solid", [("bunny", 128, 0), ("bunny", 128, 1), ("icosphere:64:128", 256, 0), ("icosphere:64:128", 256, 1)])
def test_soa4_layout_gives_same_table(vb, name, g, solid):
    v, f, d_tris = _device_mesh(name)
    grid = vb.grid_from_verts(v, g, len(f))
    fn = vb.voxelize_solid if solid else vb.voxelize
    a = fn(grid, d_tris)
    t = d_tris.view(-1, 3, 3)
    planes = torch.zeros(3, t.shape[0], 4, device="cuda")
    planes[:, :, :3] = t.permute(1, 0, 2)
    b = fn(grid, planes.contiguous().view(-1), soa4=True)
    assert torch.equal(a, b)


@pytest.mark.parametrize("solid", [0, 1])
@pytest.mark.parametrize("morton", [0, 1])
def test_accumulate_is_or_xor_into(vb, solid, morton):
    """Reference semantics: voxelize() ORs / voxelize_solid() XORs into whatever the table holds."""
    name, g = "bunny", 64
    v, f, d_tris = _device_mesh(name)
    grid = vb.grid_from_verts(v, g, len(f))
    fn = vb.voxelize_solid if solid else vb.voxelize
    clean = fn(grid, d_tris, morton=bool(morton)).clone()
    gen = torch.Generator(device="cuda").manual_seed(7)
    prior = torch.randint(-2 ** 31, 2 ** 31 - 1, clean.shape, dtype=torch.int32, device="cuda", generator=gen)
    table = prior.clone()
    fn(grid, d_tris, table=table, morton=bool(morton), accumulate=True)
    want = (prior ^ clean) if solid else (prior | clean)
    assert torch.equal(table, want)


@pytest.mark.parametrize("n_parts", [2, 4, 8])
@pytest.mark.parametrize("solid,morton", [(0, 0), (1, 0), (0, 1), (1, 1)])
def test_regions_concatenate_to_full_table(vb, n_parts, solid, morton):
    """Multi-GPU contract: disjoint regions, no reduction — concatenation == 1-GPU table."""
    name, g = "bunny", 128
    v, f, d_tris = _device_mesh(name)
    grid = vb.grid_from_verts(v, g, len(f))
    fn = vb.voxelize_solid if solid else vb.voxelize
    full = fn(grid, d_tris, morton=bool(morton)).clone()
    parts = []
    for p in range(n_parts):
        region, nbytes = vb.partition(g, morton, p, n_parts)
        t = fn(grid, d_tris, morton=bool(morton), region=region)
        assert t.numel() * 4 == nbytes
        parts.append(t.clone())
    assert torch.equal(torch.cat(parts), full)


def test_uneven_z_slabs(vb):
    name, g = "icosphere:16:64", 128
    v, f, d_tris = _device_mesh(name)
    grid = vb.grid_from_verts(v, g, len(f))
    for fn in (vb.voxelize, vb.voxelize_solid):
        full = fn(grid, d_tris).clone()
        parts = []
        for z0, z1 in ((0, 1), (1, 50), (50, 127), (127, 128)):
            r = vb.Region()
            r.lo[:] = [0, 0, z0]
            r.hi[:] = [g, g, z1]
            parts.append(fn(grid, d_tris, region=r).clone())
        assert torch.equal(torch.cat(parts), full)


@pytest.mark.parametrize("solid", [0, 1])
@pytest.mark.parametrize("name,g", [("bunny", 128), ("icosphere:64:128", 256)])
def test_routed_triangles_give_same_region(vb, name, g, solid):
    """voxb200_route_triangles: the soup routed to a slab voxelizes to the same slab bytes as the full soup, and
    the routed counts show only boundary triangles are duplicated."""
    import copy
    v, f, d_tris = _device_mesh(name)
    grid = vb.grid_from_verts(v, g, len(f))
    fn = vb.voxelize_solid if solid else vb.voxelize
    total = 0
    for p in range(4):
        region, _ = vb.partition(g, False, p, 4)
        want = fn(grid, d_tris, region=region).clone()
        buf, n = vb.route_triangles(grid, d_tris, region, solid=bool(solid))
        total += n
        g2 = copy.copy(grid)
        g2.n_triangles = n
        got = fn(g2, buf, region=region)
        assert torch.equal(got, want)
        buf.close()
    assert total >= (len(f) if not solid else 1) and total <= 2 * len(f)


@pytest.mark.parametrize("solid,morton", [(0, 0), (1, 0), (0, 1)])
def test_multi_region_routing(vb, solid, morton):
    """voxb200_route_triangles_multi: one pass routes a soup to all N regions; voxelizing segment r over region r
    reproduces region r of the full table (what every rank does after the all-to-all)."""
    import copy
    name, g, n = "icosphere:64:128", 256, 8
    v, f, d_tris = _device_mesh(name)
    grid = vb.grid_from_verts(v, g, len(f))
    fn = vb.voxelize_solid if solid else vb.voxelize
    full = fn(grid, d_tris, morton=bool(morton)).clone()
    regions = [vb.partition(g, morton, p, n)[0] for p in range(n)]
    out = torch.empty(2 * d_tris.numel(), device="cuda")
    counts = vb.route_triangles_multi(grid, d_tris, regions, out, solid=bool(solid), morton=bool(morton))
    assert len(counts) == n and sum(counts) <= 2 * len(f) and (solid or sum(counts) >= len(f))
    parts, off = [], 0
    for p in range(n):
        g2 = copy.copy(grid)
        g2.n_triangles = counts[p]
        seg = out[9 * off: 9 * (off + counts[p])] if counts[p] else torch.zeros(9, device="cuda")
        parts.append(fn(g2, seg.contiguous(), morton=bool(morton), region=regions[p]).clone())
        off += counts[p]
    assert torch.equal(torch.cat(parts), full)
    with pytest.raises(vb.VoxError):
        vb.route_triangles_multi(grid, d_tris, regions, out[:9], solid=bool(solid), morton=bool(morton))   # capacity too small


@pytest.mark.parametrize("name,g,solid", [("bunny", 64, 0), ("bunny", 128, 1), ("icosphere:16:64", 100, 0)])
def test_extract_voxels_matches_table(vb, name, g, solid):
    """voxb200_extract_voxels (device-side table consumer): ascending indices of exactly the set voxels."""
    table, grid = _run(vb, name, g, solid, 0)
    d_table = torch.from_numpy(table.view(np.int32)).cuda()
    got = vb.extract_voxels(d_table)
    bits = np.unpackbits(table.view(np.uint8).reshape(-1, 4)[:, ::-1].reshape(-1))      # bit k = voxel index k
    want = np.nonzero(bits)[0].astype(np.uint64)
    assert np.array_equal(got, want)
    assert np.array_equal(vb.extract_voxels(d_table[:1024], first_voxel=5 * 32), want[want < 1024 * 32] + 5 * 32)
    assert len(vb.extract_voxels(torch.zeros(4096, dtype=torch.int32, device="cuda"))) == 0


def test_upload_paths(vb):
    """Triangle upload (main.cpp:61-80 replaced): soup and indexed uploads, AoS and SoA4, plus the
    device bbox reduction, all lead to the same table as torch-owned memory."""
    name, g = "bunny", 128
    v, f, d_tris = _device_mesh(name)
    grid = vb.grid_from_verts(v, g, len(f))
    want = vb.voxelize(grid, d_tris).clone()
    soup = oracle.soup(v, f)
    for soa4 in (False, True):
        buf = vb.upload_soup(soup, soa4=soa4)
        assert torch.equal(vb.voxelize(grid, buf, soa4=soa4, table=torch.empty_like(want)), want)
        buf.close()
        buf, mn, mx = vb.upload_indexed(v, f, soa4=soa4)
        assert np.array_equal(mn, v.min(axis=0)) and np.array_equal(mx, v.max(axis=0))
        assert torch.equal(vb.voxelize(grid, buf, soa4=soa4, table=torch.empty_like(want)), want)
        buf.close()
    with pytest.raises(vb.VoxError):
        vb.upload_indexed(v, f + len(v))


@pytest.mark.parametrize("solid,morton", [(0, 0), (1, 0), (0, 1)])
def test_voxelize_host_end_to_end(vb, golden, solid, morton):
    name, g = "bunny", 256
    v, f, _ = _device_mesh(name)
    grid = vb.grid_from_verts(v, g, len(f))
    soup = oracle.soup(v, f)
    table, ms = vb.voxelize_host(grid, soup, solid=bool(solid), morton=bool(morton))
    want = golden[cases.case_key(name, g, solid, morton)]
    assert oracle.popcount(table) == want["popcount"] and "%016x" % oracle.fnv1a64(table) == want["fnv1a64"]
    pinned = torch.from_numpy(soup).pin_memory()
    out = torch.empty(vb.table_bytes(g) // 4, dtype=torch.int32).pin_memory()
    vb.voxelize_host(grid, pinned, out, solid=bool(solid), morton=bool(morton))
    assert np.array_equal(out.numpy().view(np.uint32), table)
    assert len(ms) == 4 and ms[3] > 0
    table2, ms2 = vb.voxelize_host_indexed(grid, v, f, solid=bool(solid), morton=bool(morton))
    assert np.array_equal(table2, table) and ms2[3] > 0


def test_empty_and_degenerate_inputs(vb):
    g = 64
    grid = vb.make_grid([0, 0, 0], [1, 1, 1], g, 0)
    t = vb.voxelize(grid, torch.zeros(9, device="cuda"))
    assert int(t.abs().sum()) == 0
    # zero-area and collinear triangles: NaN normal -> reference semantics restated by the oracle
    soup = np.array([[0.2, 0.2, 0.2] * 3, [0.1, 0.1, 0.1, 0.5, 0.5, 0.5, 0.9, 0.9, 0.9], [0.3, 0.3, 0.3, 0.3, 0.3, 0.3, 0.7, 0.3, 0.3]], np.float32)
    grid = vb.make_grid([0, 0, 0], [1, 1, 1], g, len(soup))
    got = vb.voxelize(grid, torch.from_numpy(soup).cuda()).cpu().numpy().view(np.uint32)
    want = oracle.surface(soup, np.array(grid.bbox_min[:], np.float32), np.array(grid.unit[:], np.float32), g)
    assert np.array_equal(got, want)
    # an empty mesh through every solid schedule (mark+scan at 64, row lists at 128/256) into a dirty table: all zero afterwards
    for gs in (64, 128, 256):
        grid = vb.make_grid([0, 0, 0], [1, 1, 1], gs, 0)
        dirty = torch.full((vb.table_bytes(gs) // 4,), -1, dtype=torch.int32, device="cuda")
        t = vb.voxelize_solid(grid, torch.zeros(9, device="cuda"), table=dirty)
        assert int(t.abs().sum()) == 0


@pytest.mark.parametrize("g", [32, 64, 256])
def test_tiny_and_degenerate_triangles(vb, g):
    """Triangles no larger than ~2 voxels (the branch-free <=3x3x3 path), including zero-area ones whose
    normal is NaN, points, collinear triples and triangles hugging the grid boundary."""
    rng = np.random.default_rng(123 + g)
    n = 6000
    unit = 1.0 / g
    c = rng.uniform(0.0, 1.0, (n, 1, 3))
    t = c + rng.uniform(-1.0, 1.0, (n, 3, 3)) * unit * rng.choice([0.05, 0.5, 1.0], (n, 1, 1))
    t[0:500, 1] = t[0:500, 0]                                   # two equal vertices
    t[500:1000, 1] = t[500:1000, 0]; t[500:1000, 2] = t[500:1000, 0]   # a point
    t[1000:1500, 2] = 2 * t[1000:1500, 1] - t[1000:1500, 0]     # collinear
    t[1500:2000, :, 0] = np.round(t[1500:2000, :, 0] * g) / g   # vertices exactly on voxel faces in x
    t[2000:2500, :, 2] = t[2000:2500, 0:1, 2]                   # axis-aligned in z
    soup = np.clip(t, 0.0, 1.0).reshape(n, 9).astype(np.float32)
    soup[-1] = [0, 0, 0, 1, 0, 0, 0, 1, 0]                      # keeps the bbox at the unit cube
    soup[-2] = [1, 1, 1, 0, 1, 1, 1, 0, 1]
    grid = vb.grid_from_verts(soup.reshape(-1, 3), g, n)
    bb_min, un = np.array(grid.bbox_min[:], np.float32), np.array(grid.unit[:], np.float32)
    d = torch.from_numpy(soup).cuda()
    for morton in (False, True):
        got = vb.voxelize(grid, d, morton=morton).cpu().numpy().view(np.uint32)
        want = oracle.surface(soup, bb_min, un, g, morton)
        assert np.array_equal(got, want), "morton=%s differing words %s" % (morton, np.nonzero(got ^ want)[0][:8])


@pytest.mark.parametrize("seed", range(12))
def test_fuzz_mixed_sizes_vs_oracle(vb, seed):
    """Seeded fuzz across the path boundaries: triangles from sub-voxel to ~12 voxels (27-candidate, 64-candidate and
    row-solver paths mixed inside the same warps), random orientation, random power-of-two and odd grids, both modes
    and both orders, full-table compare against the oracle."""
    rng = np.random.default_rng(1000 + seed)
    g = int(rng.choice([32, 48, 64, 100, 128, 256]))
    n = int(rng.integers(500, 6000))
    size = rng.choice([0.6, 1.5, 2.5, 3.5, 6.0, 12.0], size=(n, 1, 1)) / g
    c = rng.uniform(0.05, 0.95, (n, 1, 3))
    t = c + rng.normal(0.0, 1.0, (n, 3, 3)) * size * 0.5
    if seed % 3 == 0:
        t[: n // 4, :, seed % 3] = np.round(t[: n // 4, :, seed % 3] * g) / g       # some vertices exactly on voxel faces
    soup = np.clip(t, 0.0, 1.0).reshape(n, 9).astype(np.float32)
    soup[-1] = [0, 0, 0, 1, 0, 0, 0, 1, 0]
    soup[-2] = [1, 1, 1, 0, 1, 1, 1, 0, 1]
    grid = vb.grid_from_verts(soup.reshape(-1, 3), g, n)
    bb_min, un = np.array(grid.bbox_min[:], np.float32), np.array(grid.unit[:], np.float32)
    d = torch.from_numpy(soup).cuda()
    pow2 = (g & (g - 1)) == 0
    for morton in ((False, True) if pow2 else (False,)):
        got = vb.voxelize(grid, d, morton=morton).cpu().numpy().view(np.uint32)
        want = oracle.surface(soup, bb_min, un, g, morton)
        assert np.array_equal(got, want), "surface g=%d morton=%s: %d differing words" % (g, morton, np.count_nonzero(got ^ want))
    # solid on an arbitrary soup is geometrically meaningless but arithmetically well defined: same XOR of column runs
    got = vb.voxelize_solid(grid, d).cpu().numpy().view(np.uint32)
    before = oracle.solid_ub_events()
    want = oracle.solid(soup, bb_min, un, g)
    if oracle.solid_ub_events() == before:            # only compare where the reference itself is well defined
        assert np.array_equal(got, want), "solid g=%d: %d differing words" % (g, np.count_nonzero(got ^ want))


@pytest.mark.parametrize("g", [128, 256])
def test_solid_row_lists_with_overflowing_rows(vb, g):
    """The solid path keeps up to 8 marks per (y,z) row in a list and spills the rest into a library table that must be
    all-zero between calls.  Large overlapping triangles give rows with dozens of crossings next to rows with a few;
    repeated calls, with a scratch-dirtying call (morton order) in between, must keep giving the oracle's table."""
    v, f = cases.mesh("soup:large:300:%d:1.0" % (7 + g))
    soup = oracle.soup(v, f)
    soup[-1] = [0, 0, 0, 1, 0, 0, 0, 1, 0]
    soup[-2] = [1, 1, 1, 0, 1, 1, 1, 0, 1]
    grid = vb.grid_from_verts(soup.reshape(-1, 3), g, len(soup))
    bb_min, un = np.array(grid.bbox_min[:], np.float32), np.array(grid.unit[:], np.float32)
    before = oracle.solid_ub_events()
    want = oracle.solid(soup, bb_min, un, g)
    if oracle.solid_ub_events() != before:
        pytest.skip("the reference itself is undefined on this soup")
    d = torch.from_numpy(soup).cuda()
    # rows crossed more than 8 times exist (otherwise this test does not reach the spill path)
    first = vb.voxelize_solid(grid, d).cpu().numpy().view(np.uint32)
    assert np.array_equal(first, want), "%d differing words" % np.count_nonzero(first ^ want)
    again = vb.voxelize_solid(grid, d).cpu().numpy().view(np.uint32)
    assert np.array_equal(again, want)
    vb.voxelize_solid(grid, d, morton=True)                  # uses (and dirties) the library's scratch table
    third = vb.voxelize_solid(grid, d).cpu().numpy().view(np.uint32)
    assert np.array_equal(third, want)
    # a z-slab region takes the same path with region-relative rows
    region, nbytes = vb.partition(g, False, 1, 2)
    half = vb.voxelize_solid(grid, d, region=region).cpu().numpy().view(np.uint32)
    assert np.array_equal(half, want[len(want) // 2:])


@pytest.mark.parametrize("solid", [0, 1])
def test_cuda_graph_capture_and_replay(vb, solid):
    """voxb200_surface / voxb200_solid only enqueue work on the caller's stream (no synchronisation, no allocation once
    the library's scratch exists), so a whole voxelization can be captured into a CUDA graph and replayed — the
    per-frame use the reference README pitches.  Every replay must reproduce the directly computed table."""
    name, g = "bunny", 256
    v, f, d_tris = _device_mesh(name)
    grid = vb.grid_from_verts(v, g, len(f))
    fn = vb.voxelize_solid if solid else vb.voxelize
    want = fn(grid, d_tris).clone()                     # also the warm-up that sizes the library's scratch buffers
    torch.cuda.synchronize()
    table = torch.empty_like(want)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        fn(grid, d_tris, table=table)
    for _ in range(3):
        table.fill_(-1)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(table, want)


@pytest.mark.parametrize("name,g", [("bunny", 256), ("icosphere:64:128", 256), ("soup:mixed:3000:5:1.0", 128)])
def test_sort_triangles_keeps_the_table(vb, name, g):
    """voxb200_sort_triangles (upload-path option) permutes the soup by z-layer: same multiset of triangles, lowest-vertex z
    layers non-decreasing, and — OR / XOR being order-independent — bit-identical surface and solid tables."""
    v, f, d_tris = _device_mesh(name)
    grid = vb.grid_from_verts(v, g, len(f))
    buf = vb.sort_triangles(grid, d_tris)
    torch.cuda.synchronize()
    host = vb.download(buf.ptr, d_tris.numel() * 4).view(np.float32).reshape(-1, 9)
    orig = d_tris.cpu().numpy().reshape(-1, 9)
    key = lambda a: a[np.lexsort(a.T[::-1])]
    assert np.array_equal(key(host), key(orig)), "the sorted soup is not a permutation of the input"
    layer = np.clip(np.floor((host[:, 2::3].min(axis=1) - np.float32(grid.bbox_min[2])) * (np.float32(1.0) / np.float32(grid.unit[2]))), 0, g - 1)
    assert np.all(np.diff(layer) >= 0)
    d_sorted = torch.from_numpy(host.reshape(-1)).cuda()
    for solid in (0, 1):
        if solid and name.startswith("soup"):
            continue
        fn = vb.voxelize_solid if solid else vb.voxelize
        assert torch.equal(fn(grid, d_sorted), fn(grid, d_tris))
    buf.close()


def test_queue_unit_overflow_is_reported(vb):
    """The large-triangle queue counts its work units ((y,z) rows) in 32 bits.  A soup that queues more than 2^32 rows in
    one call must be reported (last_counters, and an error from the synchronous host entry point), not silently mangled."""
    g, n = 1024, 4600                                   # 4600 triangles spanning the whole grid: 4600 * 2^20 rows > 2^32
    tri = np.array([0, 0, 0, 1, 1, 0.9, 0.1, 1, 1], np.float32)
    soup = np.tile(tri, (n, 1))
    grid = vb.grid_from_verts(soup.reshape(-1, 3), g, n)
    vb.voxelize(grid, torch.from_numpy(soup).cuda())
    torch.cuda.synchronize()
    assert vb.last_counters()["coop_items"] == 2 ** 64 - 1
    with pytest.raises(vb.VoxError):
        vb.voxelize_host(grid, soup)
    # a normal call afterwards is unaffected
    table, _ = _run(vb, "bunny", 64, 0, 0)
    assert vb.last_counters()["coop_items"] < 2 ** 32


def test_release_and_reuse(vb):
    """voxb200_release frees the cached scratch; the next call rebuilds it and gives the same table."""
    name, g = "bunny", 64
    v, f, d_tris = _device_mesh(name)
    grid = vb.grid_from_verts(v, g, len(f))
    a = vb.voxelize(grid, d_tris).clone()
    vb.voxelize_host_indexed(grid, v, f)
    vb.release()
    assert torch.equal(vb.voxelize(grid, d_tris), a)
    assert torch.equal(vb.voxelize_solid(grid, d_tris), vb.voxelize_solid(grid, d_tris).clone())


@pytest.mark.parametrize("solid", [0, 1])
def test_unaligned_buffers_take_the_scalar_paths(vb, solid):
    """Triangle and table pointers that are only 4-byte aligned (views offset by one element): the 16-byte fast paths
    (cp.async tile fetch, 16-byte zero-fill, 16-byte scan lanes) must step aside, the table must not change."""
    name, g = "icosphere:64:128", 256
    v, f, d_tris = _device_mesh(name)
    grid = vb.grid_from_verts(v, g, len(f))
    fn = vb.voxelize_solid if solid else vb.voxelize
    want = fn(grid, d_tris).clone()
    shifted = torch.empty(d_tris.numel() + 1, device="cuda")
    shifted[1:] = d_tris.view(-1)
    guard = torch.full((want.numel() + 2,), 0x5a5a5a5a, dtype=torch.int32, device="cuda")
    got = fn(grid, shifted[1:], table=guard[1:-1])
    assert shifted[1:].data_ptr() % 16 == 4 and guard[1:-1].data_ptr() % 16 == 4
    assert torch.equal(got, want)
    assert int(guard[0]) == 0x5a5a5a5a and int(guard[-1]) == 0x5a5a5a5a          # nothing written outside the table


def test_invalid_arguments_report_einval(vb):
    from cuda_voxelizer_b200 import _lib
    grid = vb.make_grid([0, 0, 0], [1, 1, 1], 48, 1)
    tris = torch.zeros(9, device="cuda")
    with pytest.raises(vb.VoxError) as e:
        vb.voxelize(grid, tris, morton=True)          # morton needs a power-of-two grid
    assert e.value.code == _lib.EINVAL
    grid.gridsize[1] = 32
    with pytest.raises(vb.VoxError):
        vb.voxelize(grid, tris)                       # non-cubic


# ------------------------------------------------------------------ full-size properties
def test_full_size_properties_config4(vb, golden):
    """10M-triangle icosphere at 2048^3 (config 4): golden hash (checked above) plus properties that need
    no oracle: z-slab sharding reproduces the table, every set voxel lies in the sphere's shell, and
    re-voxelizing in accumulate mode is idempotent."""
    name, g = "icosphere:708:1024", 2048
    v, f, d_tris = _device_mesh(name)
    grid = vb.grid_from_verts(v, g, len(f))
    full = vb.voxelize(grid, d_tris)
    again = full.clone()
    vb.voxelize(grid, d_tris, table=again, accumulate=True)
    assert torch.equal(again, full)
    parts = []
    for p in range(8):
        region, _ = vb.partition(g, False, p, 8)
        parts.append(vb.voxelize(grid, d_tris, region=region))
    assert torch.equal(torch.cat(parts), full)
    del parts, again
    # shell property on one z-slice through the centre: set voxels are within ~2 voxels of radius 1024
    words = g * g // 32
    sl = full[1024 * words: 1025 * words].cpu().numpy().view(np.uint32)
    bits = np.unpackbits(sl.view(np.uint8).reshape(-1, 4)[:, ::-1].reshape(-1)).reshape(g, g)
    ys, xs = np.nonzero(bits)
    unit = np.float64(grid.unit[0])
    cx = (xs + 0.5) * unit + grid.bbox_min[0]
    cy = (ys + 0.5) * unit + grid.bbox_min[1]
    cz = (1024 + 0.5) * unit + grid.bbox_min[2]
    r = np.sqrt(cx ** 2 + cy ** 2 + cz ** 2)
    assert len(xs) > 5000 and np.all(np.abs(r - 1024.0) < 2.5)


def test_full_size_solid_is_scan_of_surface_columns_config3(vb):
    """Config 3 (1M-triangle icosphere, solid, 1024^3): popcount/hash are checked against the reference
    golden above; here: the solid table equals the XOR of its z-slab shards, and the filled volume is
    the sphere's to 0.01 %."""
    name, g = "icosphere:224:512", 1024
    v, f, d_tris = _device_mesh(name)
    grid = vb.grid_from_verts(v, g, len(f))
    full = vb.voxelize_solid(grid, d_tris)
    parts = [vb.voxelize_solid(grid, d_tris, region=vb.partition(g, False, p, 4)[0]) for p in range(4)]
    assert torch.equal(torch.cat(parts), full)
    pop = oracle.popcount(full.cpu().numpy().view(np.uint32))
    ideal = 4.0 / 3.0 * np.pi * 512.0 ** 3 / float(grid.unit[0]) ** 3
    assert abs(pop - ideal) / ideal < 1e-4
