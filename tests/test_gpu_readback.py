"""GPU suite (-m gpu) of the table's way back to the host (csrc/readback.cu, voxb200_download_table): the sparse read-back —
non-zero words as {index, value} pairs, expanded by host threads — must leave the host table byte-identical to the device table,
whatever was in the host buffer before, and the host entry points that use it must still match the reference goldens."""
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
import oracle

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORDS = 16 << 20            # 64 MB: above the 32 MB threshold of the sparse mode


@pytest.fixture(scope="module")
def vb():
    import cuda_voxelizer_b200 as vb
    vb.init(0)
    return vb


def _dirty_pinned(words):
    return torch.full((words,), -1, dtype=torch.int32).pin_memory()


def _random_table(words, density, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    t = torch.zeros(words, dtype=torch.int32, device="cuda")
    n = int(words * density)
    if n:
        idx = torch.randint(0, words, (n,), generator=g, device="cuda")
        val = torch.randint(-2**31, 2**31 - 1, (n,), generator=g, device="cuda", dtype=torch.int64).to(torch.int32)
        t[idx] = val
    return t


@pytest.mark.parametrize("density,expect_sparse", [(0.0, True), (1e-6, True), (0.01, True), (0.1, True), (0.5, False), (1.0, False)])
def test_download_matches_device_table(vb, density, expect_sparse):
    t = _random_table(WORDS, density, seed=int(density * 1e6) + 1)
    host = _dirty_pinned(WORDS)
    _, info = vb.download_table(t, host)
    assert info["sparse"] == expect_sparse
    assert torch.equal(host, t.cpu())
    if expect_sparse:
        assert info["nonzero_words"] == int((t != 0).sum())


def test_download_edges_of_blocks_and_slices(vb):
    t = torch.zeros(WORDS, dtype=torch.int32, device="cuda")
    edges = [0, 1, 15, 16, 2047, 2048, 2049, WORDS // 64 - 1, WORDS // 64, WORDS // 2 - 1, WORDS // 2, WORDS - 17, WORDS - 16, WORDS - 1]
    for k, e in enumerate(edges):
        t[e] = k + 1
    t[4096:4096 + 64] = 7             # four full lines in a row
    host = _dirty_pinned(WORDS)
    _, info = vb.download_table(t, host)
    assert info["sparse"] and info["nonzero_words"] == len(edges) + 64
    assert torch.equal(host, t.cpu())


def test_download_into_unaligned_or_pageable_memory(vb):
    t = _random_table(WORDS, 0.01, seed=5)
    want = t.cpu().numpy()
    raw = np.full(WORDS + 16, 0xFFFFFFFF, np.uint32)
    off = (-(raw.ctypes.data // 4) % 16 + 1) % 16 or 1          # a word offset that is NOT 64-byte aligned
    host = raw[off:off + WORDS]
    assert host.ctypes.data % 64 != 0
    _, info = vb.download_table(t, host)
    assert not info["sparse"]                                  # streaming stores need the alignment: dense copy
    assert np.array_equal(host.view(np.int32), want)
    aligned = raw[(-(raw.ctypes.data // 4) % 16):][:WORDS]
    assert aligned.ctypes.data % 64 == 0
    aligned[:] = 0xFFFFFFFF
    _, info = vb.download_table(t, aligned)                   # pageable but aligned: sparse, the table is the threads' to write
    assert info["sparse"]
    assert np.array_equal(aligned.view(np.int32), want)


def test_mode_switch(vb):
    t = _random_table(WORDS, 0.01, seed=9)
    host = _dirty_pinned(WORDS)
    try:
        vb.set_readback_mode("dense")
        _, info = vb.download_table(t, host)
        assert not info["sparse"] and torch.equal(host, t.cpu())
        vb.set_readback_mode("sparse")
        small = _random_table(4096, 0.5, seed=3)
        hs = _dirty_pinned(4096)
        _, info = vb.download_table(small, hs)
        assert info["sparse"] and torch.equal(hs, small.cpu())
    finally:
        vb.set_readback_mode("auto")


def test_small_tables_forced_sparse():
    """VOXB200_READBACK=sparse makes every eligible table take the sparse mode (the initial mode is read once per process: a subprocess)."""
    code = r'''
import numpy as np, torch, sys
sys.path.insert(0, %r)
import cuda_voxelizer_b200 as vb
vb.init(0)
for words in (16, 32, 2048, 2048 + 16, 5 * 2048, 100000 * 16):
    g = torch.Generator(device="cuda").manual_seed(words)
    t = torch.randint(-5, 5, (words,), generator=g, device="cuda", dtype=torch.int64).to(torch.int32)
    t[t < 3] = 0
    host = torch.full((words,), -1, dtype=torch.int32).pin_memory()
    _, info = vb.download_table(t, host)
    assert info["sparse"], words
    assert torch.equal(host, t.cpu()), words
print("ok")
''' % ROOT
    env = dict(os.environ, VOXB200_READBACK="sparse", VOXB200_HOST_THREADS="3")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("name,g,solid", [("bunny", 1024, 0), ("bunny", 1024, 1), ("icosphere:224:512", 1024, 0)])
def test_host_entry_points_match_golden_through_the_readback(vb, golden, name, g, solid):
    want = golden[cases.case_key(name, g, solid, 0)]
    v, f = cases.mesh(name)
    grid = vb.grid_from_verts(v, g, len(f))
    hv = torch.from_numpy(np.ascontiguousarray(v)).pin_memory()
    hf = torch.from_numpy(np.ascontiguousarray(f)).pin_memory()
    out = _dirty_pinned(vb.table_bytes(g) // 4)
    for _ in range(2):
        out.fill_(-1)
        vb.voxelize_host_indexed(grid, hv, hf, out, solid=bool(solid))
    host = out.numpy().view(np.uint32)
    assert oracle.popcount(host) == want["popcount"]
    assert "%016x" % oracle.fnv1a64(host) == want["fnv1a64"]
    out.fill_(-1)
    vb.voxelize_host_multi(grid, hv, hf, out, solid=bool(solid), n_devices=1)
    assert "%016x" % oracle.fnv1a64(out.numpy().view(np.uint32)) == want["fnv1a64"]


@pytest.mark.parametrize("name,g,solid", [("bunny", 1024, 0), ("icosphere:64:128", 256, 0), ("bunny", 256, 1)])
def test_host_nonzero_words_rebuild_the_golden_table(vb, golden, name, g, solid):
    """voxb200_voxelize_host_nonzero: the {index, bits} pairs are exactly the non-zero words of the reference's table, ascending."""
    want = golden[cases.case_key(name, g, solid, 0)]
    v, f = cases.mesh(name)
    grid = vb.grid_from_verts(v, g, len(f))
    hv = torch.from_numpy(np.ascontiguousarray(v)).pin_memory()
    hf = torch.from_numpy(np.ascontiguousarray(f)).pin_memory()
    for _ in range(2):
        pairs, ms = vb.voxelize_host_nonzero(grid, hv, hf, solid=bool(solid))
    assert ms[3] > 0.0
    assert np.all(pairs[:, 1] != 0) and np.all(np.diff(pairs[:, 0].astype(np.int64)) > 0)
    table = np.zeros(vb.table_bytes(g) // 4, np.uint32)
    table[pairs[:, 0]] = pairs[:, 1]
    assert oracle.popcount(table) == want["popcount"]
    assert "%016x" % oracle.fnv1a64(table) == want["fnv1a64"]
