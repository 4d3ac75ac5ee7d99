"""Deterministic synthetic meshes for the BASELINE.json configs (no RNG unless a seed is asked for).

These build the *inputs* of the hot path: indexed, watertight,
outward-CCW triangle meshes at about one world unit per voxel (radius = G/2) — the scale at which
the reference's solid path is well defined.  Vertices are computed in float64 and
cast to float32 once, so shared edges are bit-identical between neighbouring triangles.
"""
import numpy as np

_PHI = (1.0 + 5.0 ** 0.5) / 2.0

_ICO_V = np.array([
    [-1, _PHI, 0], [1, _PHI, 0], [-1, -_PHI, 0], [1, -_PHI, 0],
    [0, -1, _PHI], [0, 1, _PHI], [0, -1, -_PHI], [0, 1, -_PHI],
    [_PHI, 0, -1], [_PHI, 0, 1], [-_PHI, 0, -1], [-_PHI, 0, 1]], dtype=np.float64)
_ICO_F = np.array([
    [0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11],
    [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
    [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9],
    [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)


def euler_xyz(rx, ry, rz):
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def icosphere(nu, radius=1.0, rotation=(0.3, 0.5, 0.7), center=(0.0, 0.0, 0.0)):
    """Class-I geodesic icosphere of frequency ``nu``: 20*nu^2 triangles, 10*nu^2+2 shared vertices.

    Returns (verts float32 [V,3], faces int32 [T,3]).  nu=224 -> 1,003,520 tris (config 3);
    nu=708 -> 10,025,280 tris (config 4).
    """
    nu = int(nu)
    assert nu >= 1
    ico_v = _ICO_V / np.linalg.norm(_ICO_V[0])
    ico_f = _ICO_F.copy()
    # make sure base faces are outward CCW
    c = ico_v[ico_f].mean(axis=1)
    nrm = np.cross(ico_v[ico_f[:, 1]] - ico_v[ico_f[:, 0]], ico_v[ico_f[:, 2]] - ico_v[ico_f[:, 0]])
    flip = (nrm * c).sum(axis=1) < 0
    ico_f[flip] = ico_f[flip][:, [0, 2, 1]]

    # canonical undirected edges
    edge_id = {}
    for f in ico_f:
        for a, b in ((f[0], f[1]), (f[1], f[2]), (f[2], f[0])):
            key = (min(a, b), max(a, b))
            if key not in edge_id:
                edge_id[key] = len(edge_id)
    assert len(edge_id) == 30
    n_edge_pts = nu - 1
    n_int_pts = (nu - 1) * (nu - 2) // 2
    n_verts = 12 + 30 * n_edge_pts + 20 * n_int_pts
    pos = np.empty((n_verts, 3), np.float64)
    pos[:12] = ico_v
    t = np.arange(1, nu, dtype=np.float64)[:, None]
    for (a, b), e in edge_id.items():
        pos[12 + e * n_edge_pts: 12 + (e + 1) * n_edge_pts] = ((nu - t) * ico_v[a] + t * ico_v[b]) / nu

    # lattice (a, b): weight nu-a-b on A, a on B, b on C
    aa, bb = np.meshgrid(np.arange(nu + 1), np.arange(nu + 1), indexing="ij")
    valid = (aa + bb) <= nu
    faces_out = []
    # local triangle templates over the lattice
    ua, ub = np.nonzero((aa + bb) <= nu - 1)           # upward  (a,b),(a+1,b),(a,b+1)
    da, db = np.nonzero((aa + bb) <= nu - 2)           # downward (a+1,b),(a+1,b+1),(a,b+1)
    interior = valid & (aa > 0) & (bb > 0) & ((aa + bb) < nu)
    ia, ib = np.nonzero(interior)
    assert len(ia) == n_int_pts
    for fi, (A, B, Cv) in enumerate(ico_f):
        idx = np.full((nu + 1, nu + 1), -1, np.int64)
        idx[0, 0], idx[nu, 0], idx[0, nu] = A, B, Cv
        if nu > 1:
            s = np.arange(1, nu)

            def edge_ids(p, q):
                # ids of the points at s steps from p toward q, s = 1..nu-1
                e = edge_id[(min(p, q), max(p, q))]
                tt = s if p < q else nu - s
                return 12 + e * n_edge_pts + (tt - 1)
            idx[s, 0] = edge_ids(A, B)            # b = 0: from A to B
            idx[0, s] = edge_ids(A, Cv)           # a = 0: from A to C
            idx[nu - s, s] = edge_ids(B, Cv)      # a + b = nu: from B to C
            base = 12 + 30 * n_edge_pts + fi * n_int_pts
            idx[ia, ib] = base + np.arange(n_int_pts)
            w = (nu - ia - ib)[:, None] * ico_v[A] + ia[:, None] * ico_v[B] + ib[:, None] * ico_v[Cv]
            pos[base: base + n_int_pts] = w / nu
        up = np.stack([idx[ua, ub], idx[ua + 1, ub], idx[ua, ub + 1]], axis=1)
        dn = np.stack([idx[da + 1, db], idx[da + 1, db + 1], idx[da, db + 1]], axis=1)
        faces_out.append(np.concatenate([up, dn], axis=0))
    faces = np.concatenate(faces_out, axis=0)
    assert faces.min() >= 0 and len(faces) == 20 * nu * nu
    pos /= np.linalg.norm(pos, axis=1, keepdims=True)
    R = euler_xyz(*rotation)
    pos = (pos @ R.T) * float(radius) + np.asarray(center, np.float64)
    return pos.astype(np.float32), faces.astype(np.int32)


def torus(nu, nv, R=1.0, r=0.4, rotation=(0.3, 0.5, 0.7)):
    """Watertight torus with nu x nv quads split in two: 2*nu*nv triangles, nu*nv vertices."""
    u = np.arange(nu, dtype=np.float64) * (2 * np.pi / nu)
    v = np.arange(nv, dtype=np.float64) * (2 * np.pi / nv)
    U, V = np.meshgrid(u, v, indexing="ij")
    pos = np.stack([(R + r * np.cos(V)) * np.cos(U), (R + r * np.cos(V)) * np.sin(U), r * np.sin(V)], axis=-1).reshape(-1, 3)
    i, j = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    i1, j1 = (i + 1) % nu, (j + 1) % nv
    p00, p10, p01, p11 = i * nv + j, i1 * nv + j, i * nv + j1, i1 * nv + j1
    faces = np.concatenate([np.stack([p00, p10, p11], -1).reshape(-1, 3), np.stack([p00, p11, p01], -1).reshape(-1, 3)])
    pos = pos @ euler_xyz(*rotation).T
    return pos.astype(np.float32), faces.astype(np.int32)


def random_soup(n, seed, extent=1.0, kind="mixed"):
    """Seeded triangle soups for property tests: returns (verts [3n,3] float32, faces [n,3] int32).

    kinds: 'small' (edge ~ 2% of extent), 'large' (vertices anywhere), 'sliver', 'axis' (one
    coordinate shared by all three vertices), 'mixed' (all of the above).
    """
    rng = np.random.default_rng(seed)
    kinds = ["small", "large", "sliver", "axis"] if kind == "mixed" else [kind]
    chunks = []
    per = -(-n // len(kinds))
    for k in kinds:
        c = rng.uniform(0.1, 0.9, (per, 1, 3)) * extent
        if k == "small":
            t = c + rng.normal(0, 0.02 * extent, (per, 3, 3))
        elif k == "large":
            t = rng.uniform(0, extent, (per, 3, 3))
        elif k == "sliver":
            d = rng.normal(0, 0.3 * extent, (per, 1, 3))
            t = c + d * rng.uniform(-1, 1, (per, 3, 1)) + rng.normal(0, 1e-3 * extent, (per, 3, 3))
        else:
            t = c + rng.normal(0, 0.1 * extent, (per, 3, 3))
            ax = rng.integers(0, 3, per)
            t[np.arange(per), :, ax] = t[np.arange(per), 0:1, ax]
        chunks.append(np.clip(t, 0, extent))
    tri = np.concatenate(chunks)[:n]
    verts = tri.reshape(-1, 3).astype(np.float32)
    faces = np.arange(3 * len(tri), dtype=np.int32).reshape(-1, 3)
    return verts, faces


def box(extent=1.0):
    """Axis-aligned cube [0, extent]^3 as 12 outward-CCW triangles (8 shared vertices).  Its far corner lands in voxel
    (G-1, G-1, G-1): the mesh that exercises the last bit of the table."""
    e = float(extent)
    v = np.array([[x, y, z] for z in (0.0, e) for y in (0.0, e) for x in (0.0, e)], np.float32)
    f = np.array([[0, 2, 1], [1, 2, 3], [4, 5, 6], [5, 7, 6], [0, 1, 4], [1, 5, 4],
                  [2, 6, 3], [3, 6, 7], [0, 4, 2], [2, 4, 6], [1, 3, 5], [3, 7, 5]], np.int32)
    return v, f
