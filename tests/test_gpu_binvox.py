"""GPU suite (-m gpu) of the device-side binvox encoder (csrc/binvox.cu, voxb200_binvox_rle; reference: util_io.cpp:202-246)."""
import json
import os
import subprocess

import numpy as np
import pytest

import cases
import oracle

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
IO_DIR = os.path.join(ROOT, "tests", "golden", "io")
CLI = os.path.join(ROOT, "cuda_voxelizer_b200", "bin", "cuda_voxelizer")


@pytest.fixture(scope="module")
def vb():
    import cuda_voxelizer_b200 as vb
    vb.init(0)
    return vb


def writer_loop_payload(table, g):
    """util_io.cpp:219-245 restated with numpy: voxels visited x-major, then z, then y; a pair is flushed when the value changes or
    the count reaches 255."""
    bits = np.unpackbits(table.view(np.uint8).reshape(-1, 4)[:, ::-1].reshape(-1))[: g * g * g]       # MSB-first words -> voxel idx order
    vol = bits.reshape(g, g, g)                       # [z][y][x]
    stream = np.ascontiguousarray(vol.transpose(2, 0, 1)).reshape(-1)      # x, z, y
    edges = np.flatnonzero(np.diff(stream)) + 1
    starts = np.concatenate([[0], edges])
    lengths = np.diff(np.concatenate([starts, [stream.size]]))
    values = stream[starts]
    n_pairs = (lengths + 254) // 255
    out_val = np.repeat(values, n_pairs)
    out_cnt = np.full(out_val.shape, 255, np.int64)
    last = np.cumsum(n_pairs) - 1
    out_cnt[last] = lengths - 255 * (n_pairs - 1)
    return np.stack([out_val.astype(np.uint8), out_cnt.astype(np.uint8)], axis=1).reshape(-1)


@pytest.mark.parametrize("key", ["bunny|256|surface", "bunny|256|solid", "bunny|512|surface"])
def test_payload_matches_the_reference_writer(vb, key):
    want = json.load(open(os.path.join(IO_DIR, "binvox_large.json")))[key]
    _, g, mode = key.split("|")
    g = int(g)
    v, f = cases.mesh("bunny")
    grid = vb.grid_from_verts(v, g, len(f))
    d = torch.from_numpy(oracle.soup(v, f)).cuda()
    table = (vb.voxelize_solid if mode == "solid" else vb.voxelize)(grid, d)
    payload = vb.binvox_rle(table, g)
    assert len(payload) == want["bytes"] - want["header_bytes"]
    assert "%016x" % oracle.fnv1a64(payload) == want["payload_fnv1a64"]


@pytest.mark.parametrize("kind", ["empty", "full", "first", "last", "random_sparse", "random_dense", "slabs"])
def test_payload_matches_the_writer_loop(vb, kind):
    g = 256
    words = g * g * g // 32
    rng = np.random.default_rng(11)
    t = np.zeros(words, np.uint32)
    if kind == "full":
        t[:] = 0xFFFFFFFF
    elif kind == "first":
        t[0] = 0x80000000
    elif kind == "last":
        t[-1] = 1
    elif kind == "random_sparse":
        idx = rng.integers(0, words, 5000)
        t[idx] = rng.integers(1, 2**32, 5000, dtype=np.uint64).astype(np.uint32)
    elif kind == "random_dense":
        t = rng.integers(0, 2**32, words, dtype=np.uint64).astype(np.uint32)
    elif kind == "slabs":                       # long runs of ones: more than 2^16 pairs in a run (x-slabs are contiguous in the stream)
        vol = np.zeros((g, g, g), np.uint8)     # [z][y][x]
        vol[:, :, 3:140] = 1
        vol[17, 200, 77] = 0
        t = np.packbits(vol.reshape(-1)).view(np.uint8).reshape(-1, 4)[:, ::-1].copy().view(np.uint32).reshape(-1)
    want = writer_loop_payload(t, g)
    got = vb.binvox_rle(torch.from_numpy(t.view(np.int32)).cuda(), g)
    assert len(got) == len(want)
    assert np.array_equal(got, want)


def test_rejects_unsupported_grids(vb):
    t = torch.zeros(64 * 64 * 64 // 32, dtype=torch.int32, device="cuda")
    with pytest.raises(vb.VoxError):
        vb.binvox_rle(t, 64)


def test_cli_binvox_file_at_256(tmp_path):
    """-o binvox at a grid size the device encoder covers: the FILE equals the reference writer's, byte for byte."""
    want = json.load(open(os.path.join(IO_DIR, "binvox_large.json")))["bunny|256|surface"]
    v, f = cases.mesh("bunny")
    obj = tmp_path / "bunny.OBJ"
    with open(obj, "w") as fh:
        for p in v:
            fh.write("v %.9g %.9g %.9g\n" % tuple(p))
        for t3 in f:
            fh.write("f %d %d %d\n" % tuple(t3 + 1))
    r = subprocess.run([CLI, "-f", str(obj), "-s", "256", "-o", "binvox"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "run-length encoded on the GPU" in r.stdout
    data = np.fromfile(str(obj) + "_256.binvox", np.uint8)
    assert len(data) == want["bytes"]
    assert "%016x" % oracle.fnv1a64(data) == want["fnv1a64"]
