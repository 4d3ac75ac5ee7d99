"""Small end-to-end run of every kernel family for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python scripts/sanitize.py"""
import os, sys, copy
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cuda_voxelizer_b200 as vb
import cases, oracle
vb.init(0)
for name, g in (("bunny", 64), ("icosphere:64:128", 256), ("soup:mixed:2000:1:64", 64)):
    v, f = cases.mesh(name)
    soup = oracle.soup(v, f)
    d = torch.from_numpy(soup).cuda()
    grid = vb.grid_from_verts(v, g, len(f))
    mn, mx, unit = oracle.voxinfo(v, g)
    for solid in (False, True):
        if solid and name.startswith("soup"):
            continue
        for morton in (False, True):
            fn = vb.voxelize_solid if solid else vb.voxelize
            got = fn(grid, d, morton=morton)
            torch.cuda.synchronize()
            want = (oracle.solid if solid else oracle.surface)(soup, mn, unit, g, morton)
            assert np.array_equal(got.cpu().numpy().view(np.uint32), want), (name, solid, morton)
    t = vb.voxelize(grid, d)
    idx = vb.extract_voxels(t)
    assert len(idx) == oracle.popcount(t.cpu().numpy().view(np.uint32))
    regions = [vb.partition(g, False, p, 4)[0] for p in range(4)]
    out = torch.empty(2 * d.numel(), device="cuda")
    counts = vb.route_triangles_multi(grid, d, regions, out)
    parts, off = [], 0
    for p in range(4):
        g2 = copy.copy(grid); g2.n_triangles = counts[p]
        seg = out[9 * off: 9 * (off + counts[p])] if counts[p] else torch.zeros(9, device="cuda")
        parts.append(vb.voxelize(g2, seg.contiguous(), region=regions[p]).clone()); off += counts[p]
    assert torch.equal(torch.cat(parts), t)
    buf, mn2, mx2 = vb.upload_indexed(v, f, soa4=True)
    assert torch.equal(vb.voxelize(grid, buf, soa4=True), t)
    table, _ = vb.voxelize_host_indexed(grid, v, f)
    assert np.array_equal(table, t.cpu().numpy().view(np.uint32))
# solid rows with more crossings than list slots (the spill path of the row-list schedule), twice in a row
v, f = cases.mesh("soup:large:300:135:1.0")
soup = oracle.soup(v, f)
soup[-1] = [0, 0, 0, 1, 0, 0, 0, 1, 0]; soup[-2] = [1, 1, 1, 0, 1, 1, 1, 0, 1]
grid = vb.grid_from_verts(soup.reshape(-1, 3), 128, len(soup))
want = oracle.solid(soup, np.array(grid.bbox_min[:], np.float32), np.array(grid.unit[:], np.float32), 128)
d = torch.from_numpy(soup).cuda()
for _ in range(2):
    assert np.array_equal(vb.voxelize_solid(grid, d).cpu().numpy().view(np.uint32), want)
assert vb.last_counters()["solid_row_lists"] == 1
print("sanitize run ok")
# round 2: prepared mesh (tile schedule) with update, binvox encoder, sparse read-back (forced: the tables here are small)
v, f = cases.mesh("icosphere:64:128")
soup = oracle.soup(v, f)
d = torch.from_numpy(soup).cuda()
grid = vb.grid_from_verts(v, 256, len(f))
ref = vb.voxelize(grid, d)
m = vb.Mesh(grid, tris=d)
assert m.info()["tile_schedule"] == 1
assert torch.equal(m.voxelize(), ref)
m.update(tris=d)
assert torch.equal(m.voxelize(), ref)
mi = vb.Mesh(grid, verts=torch.from_numpy(np.ascontiguousarray(v)).cuda(), faces=torch.from_numpy(np.ascontiguousarray(f)).cuda())
assert torch.equal(mi.voxelize(), ref)
m.close(); mi.close()
payload = vb.binvox_rle(ref, 256)
assert len(payload) % 2 == 0 and int(payload[1::2].astype(np.int64).sum()) == 256 ** 3
vb.set_readback_mode("sparse")
host = torch.full((ref.numel(),), -1, dtype=torch.int32).pin_memory()
_, info = vb.download_table(ref, host)
assert info["sparse"] and torch.equal(host, ref.cpu())
host.fill_(-1)
vb.voxelize_host_indexed(grid, torch.from_numpy(np.ascontiguousarray(v)).pin_memory(), torch.from_numpy(np.ascontiguousarray(f)).pin_memory(), host)
assert torch.equal(host, ref.cpu())
vb.set_readback_mode("auto")
print("sanitize.py: all paths ok")
