"""Minimal OBJ reader/writer (host utility for tests and the bench; the C++ CLI has its own loader).

Restates what the reference gets from trimesh2's ``TriMesh::read`` + ``need_faces`` for the bundled
fixture (main.cpp:174-175): ``v x y z`` lines parsed to float32, ``f a//n b//n c//n`` (or ``a/t/n``,
``a``) faces taking the position index, 1-based, negative = relative; polygons are fanned.
"""
import numpy as np


def read_obj(path):
    verts, faces = [], []
    with open(path, "r") as fh:
        for line in fh:
            if line.startswith("v "):
                p = line.split()
                verts.append((np.float32(p[1]), np.float32(p[2]), np.float32(p[3])))
            elif line.startswith("f "):
                idx = []
                for tok in line.split()[1:]:
                    i = int(tok.split("/")[0])
                    idx.append(i - 1 if i > 0 else len(verts) + i)
                for k in range(1, len(idx) - 1):
                    faces.append((idx[0], idx[k], idx[k + 1]))
    return np.asarray(verts, np.float32).reshape(-1, 3), np.asarray(faces, np.int32).reshape(-1, 3)


def write_obj(path, verts, faces):
    """%.9g round-trips every float32 exactly."""
    with open(path, "w") as fh:
        for v in np.asarray(verts, np.float32):
            fh.write("v %.9g %.9g %.9g\n" % (v[0], v[1], v[2]))
        for f in np.asarray(faces):
            fh.write("f %d %d %d\n" % (f[0] + 1, f[1] + 1, f[2] + 1))
