"""Named meshes shared by the golden generator (tests/golden/make_golden.py) and the parity tests.

Every mesh is deterministic.  ``bunny`` is the reference's only fixture (test_models/bunny.OBJ,
5,110 triangles — SURVEY.md F1), stored parsed in tests/golden/bunny.npz because /root/reference
does not exist on the GPU box.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from cuda_voxelizer_b200 import meshgen  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def bunny():
    d = np.load(os.path.join(GOLDEN_DIR, "bunny.npz"))
    return d["verts"], d["faces"]


def mesh(name):
    """name -> (verts float32 [V,3], faces int32 [T,3])"""
    if name == "bunny":
        return bunny()
    kind, _, arg = name.partition(":")
    if kind == "icosphere":          # icosphere:<nu>:<radius>
        nu, radius = arg.split(":")
        return meshgen.icosphere(int(nu), radius=float(radius))
    if kind == "torus":              # torus:<nu>:<nv>:<G>  (R=0.35G, r=0.14G)
        nu, nv, g = arg.split(":")
        return meshgen.torus(int(nu), int(nv), R=0.35 * float(g), r=0.14 * float(g))
    if kind == "soup":               # soup:<kind>:<n>:<seed>:<extent>
        k, n, seed, extent = arg.split(":")
        return meshgen.random_soup(int(n), int(seed), extent=float(extent), kind=k)
    if kind == "box":                # box:<extent>
        return meshgen.box(float(arg))
    raise KeyError(name)


# (mesh name, gridsize, solid, morton).  Sizes marked "slow" are generated once into golden.json and
# only re-run by the GPU parity tests (hash compare), never by the CPU suite.
GOLDEN_CASES = [
    # --- the reference's own fixture: configs 1, 2, 5 and the CI smoke size (autobuild.yml:62)
    ("bunny", 64, 0, 0), ("bunny", 64, 0, 1), ("bunny", 64, 1, 0), ("bunny", 64, 1, 1),
    ("bunny", 128, 0, 0), ("bunny", 128, 1, 0),
    ("bunny", 256, 0, 0), ("bunny", 256, 0, 1), ("bunny", 256, 1, 0), ("bunny", 256, 1, 1),
    ("bunny", 512, 0, 0), ("bunny", 512, 1, 0),
    ("bunny", 1024, 0, 0), ("bunny", 1024, 0, 1), ("bunny", 1024, 1, 0), ("bunny", 1024, 1, 1),
    ("bunny", 2048, 0, 0),
    ("bunny", 4096, 0, 0),          # 8 GiB table: the largest grid the fast paths cover
    ("bunny", 2048, 1, 0), ("bunny", 4096, 1, 0),        # solid row lists with 16 and 32 lanes per row
    # --- synthetic watertight meshes at one world unit per voxel
    ("icosphere:16:64", 128, 0, 0), ("icosphere:16:64", 128, 1, 0), ("icosphere:16:64", 128, 1, 1),
    ("icosphere:64:128", 256, 0, 0), ("icosphere:64:128", 256, 0, 1), ("icosphere:64:128", 256, 1, 0),
    ("torus:100:50:256", 256, 0, 0), ("torus:100:50:256", 256, 1, 0), ("torus:100:50:256", 256, 1, 1),
    # --- seeded soups: slivers, axis-aligned, huge and tiny triangles (surface only: not watertight)
    ("soup:mixed:2000:1:64", 64, 0, 0), ("soup:mixed:2000:2:128", 128, 0, 0), ("soup:mixed:2000:2:128", 128, 0, 1),
    ("soup:large:64:3:256", 256, 0, 0), ("soup:sliver:4000:4:256", 256, 0, 0), ("soup:axis:4000:5:256", 256, 0, 0),
    ("soup:small:50000:6:512", 512, 0, 0),
    # --- BASELINE.json configs 3 and 4 at full size
    ("icosphere:224:512", 1024, 1, 0), ("icosphere:224:512", 1024, 0, 0),
    ("icosphere:708:1024", 2048, 0, 0),
]


def case_key(name, g, solid, morton):
    return "%s|%d|%s|%s" % (name, g, "solid" if solid else "surface", "morton" if morton else "linear")
