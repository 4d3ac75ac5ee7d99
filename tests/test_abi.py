"""CPU suite: the C-ABI library loads, exports every symbol include/voxb200.h declares, and its host
logic (grid parameters, table size, partitions, morton) agrees bit-for-bit with the oracle.  No
compute entry point is called with a GPU here; they must fail loudly without one."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import cases
import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def vb():
    import cuda_voxelizer_b200 as vb
    from cuda_voxelizer_b200 import _lib
    if not os.path.exists(_lib.SO_PATH):
        _lib.build()
    _lib.lib()
    return vb


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_header_symbols_all_exported(vb):
    from cuda_voxelizer_b200 import _lib
    header = open(os.path.join(ROOT, "include", "voxb200.h")).read()
    declared = sorted(set(re.findall(r"\b(voxb200_[a-z0-9_]+)\s*\(", header)))
    assert declared == sorted(_lib.EXPORTS)
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.SO_PATH], capture_output=True, text=True).stdout
    for name in declared:
        assert re.search(r"\bT %s\b" % name, nm), name
    # the reference's C++ entry points (main.cpp:23-24, util_cuda.h:12) with their exact mangled names
    for sym in ("_Z8voxelizeRK7voxinfoPfPjb", "_Z14voxelize_solidRK7voxinfoPfPjb", "_Z8initCudav"):
        assert sym in nm, sym


def test_grid_struct_matches_reference_voxinfo_layout(vb):
    g = vb.Grid
    assert C.sizeof(g) == 64
    assert (g.bbox_min.offset, g.bbox_max.offset, g.gridsize.offset, g.n_triangles.offset, g.unit.offset) == (0, 12, 24, 40, 48)
    if oracle.have_ref():
        lay = oracle.ref_voxinfo_layout()
        assert lay == {"sizeof": 64, "bbox": 0, "gridsize": 24, "n_triangles": 40, "unit": 48, "alignof": 8}


@pytest.mark.parametrize("name", ["bunny", "icosphere:16:64", "torus:100:50:256", "soup:mixed:2000:1:64", "soup:axis:4000:5:256"])
@pytest.mark.parametrize("g", [8, 64, 100, 1024, 2048])
def test_make_grid_bit_identical_to_oracle(vb, name, g):
    v, f = cases.mesh(name)
    grid = vb.grid_from_verts(v, g, len(f))
    mn, mx, unit = oracle.voxinfo(v, g)
    assert np.array_equal(np.array(grid.bbox_min[:], np.float32), mn)
    assert np.array_equal(np.array(grid.bbox_max[:], np.float32), mx)
    assert np.array_equal(np.array(grid.unit[:], np.float32), unit)
    assert list(grid.gridsize) == [g, g, g] and grid.n_triangles == len(f)


def test_make_grid_golden(vb, golden):
    v, f = cases.mesh("bunny")
    grid = vb.grid_from_verts(v, 64, len(f))
    want = golden["bunny|64|surface|linear"]
    assert [float(x) for x in grid.bbox_min] == want["bbox_min"] and [float(x) for x in grid.unit] == want["unit"]


@pytest.mark.parametrize("g", [1, 2, 3, 8, 31, 32, 33, 100, 256, 257, 513, 1024, 1025, 1026, 2048])
def test_table_bytes(vb, g):
    assert vb.table_bytes(g) == oracle.lib().oracle_table_bytes(g)
    assert vb.table_bytes(g) * 8 >= g ** 3                 # every voxel has a bit (the reference's own formula fails this at 257, 513, 1025, 1026)
    ref = int(_lib_raw(vb).voxb200_reference_table_bytes(g))
    assert ref == oracle.lib().oracle_reference_table_bytes(g) and vb.table_bytes(g) >= ref
    if g % 32 == 0:
        assert vb.table_bytes(g) == ref == g ** 3 // 8


def _lib_raw(vb):
    from cuda_voxelizer_b200 import _lib
    return _lib.lib()


def test_morton_encode_matches_oracle(vb):
    rng = np.random.default_rng(1)
    for x, y, z in rng.integers(0, 1 << 16, (2000, 3)):
        assert vb.morton_encode(x, y, z) == oracle.morton(x, y, z)
    assert vb.morton_encode(1, 0, 0) == 1 and vb.morton_encode(0, 1, 0) == 2 and vb.morton_encode(0, 0, 1) == 4


@pytest.mark.parametrize("n", [1, 2, 4, 8])
@pytest.mark.parametrize("g", [64, 1024, 2048])
def test_linear_partition_tiles_the_table(vb, g, n):
    total, prev_hi = 0, 0
    for p in range(n):
        r, nbytes = vb.partition(g, False, p, n)
        assert list(r.lo)[:2] == [0, 0] and list(r.hi)[:2] == [g, g]
        assert r.lo[2] == prev_hi
        prev_hi = r.hi[2]
        assert nbytes == g * g * (r.hi[2] - r.lo[2]) // 8
        total += nbytes
    assert prev_hi == g and total == vb.table_bytes(g)


@pytest.mark.parametrize("n", [1, 2, 4, 8, 16])
def test_morton_partition_is_contiguous_curve_run(vb, n):
    g = 64
    seen = np.zeros(g ** 3, bool)
    for p in range(n):
        r, nbytes = vb.partition(g, True, p, n)
        assert nbytes == g ** 3 // 8 // n
        xs, ys, zs = (np.arange(r.lo[k], r.hi[k]) for k in range(3))
        X, Y, Z = np.meshgrid(xs, ys, zs, indexing="ij")
        idx = np.array([oracle.morton(a, b, c) for a, b, c in zip(X.ravel(), Y.ravel(), Z.ravel())])
        assert idx.min() == p * g ** 3 // n and idx.max() == (p + 1) * g ** 3 // n - 1 and len(np.unique(idx)) == len(idx)
        seen[idx] = True
    assert seen.all()


def test_partition_rejects_bad_arguments(vb):
    with pytest.raises(vb.VoxError):
        vb.partition(64, False, 2, 2)
    with pytest.raises(vb.VoxError):
        vb.partition(100, False, 0, 3)      # 100^2 bits per slice is not a multiple of 32
    with pytest.raises(vb.VoxError):
        vb.partition(64, True, 0, 3)


@pytest.mark.skipif(_have_gpu(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu(vb):
    """No CPU fallback in the product: without a device every compute entry point reports ENODEVICE."""
    from cuda_voxelizer_b200 import _lib
    with pytest.raises(vb.VoxError) as e:
        vb.init(0)
    assert e.value.code == _lib.ENODEVICE
    grid = vb.make_grid([0, 0, 0], [1, 1, 1], 32, 1)
    tris = np.zeros(9, np.float32)
    table = np.zeros(vb.table_bytes(32) // 4, np.uint32)
    with pytest.raises(vb.VoxError) as e:
        vb.voxelize_host(grid, tris, table)
    assert e.value.code == _lib.ENODEVICE

    class FakeDeviceSoup:           # any pointer will do: the call must stop at the device check
        def data_ptr(self):
            return tris.ctypes.data
    with pytest.raises(vb.VoxError) as e:
        vb.sort_triangles(grid, FakeDeviceSoup(), stream=0)
    assert e.value.code == _lib.ENODEVICE


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cuda_voxelizer_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or fn == "Makefile":
                src = open(os.path.join(dirpath, fn), errors="replace").read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src and "libvoxref" not in src, fn


def test_no_contracted_packed_arithmetic_in_sass():
    """The per-triangle kernels pair additions into FADD2 (two rounded binary32 adds per instruction).  ptxas 12.9
    contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under -fmad=false, which would change results, so products
    must stay scalar: the library's SASS may hold FADD2 but never FFMA2 / FMUL2."""
    import shutil
    import subprocess
    from cuda_voxelizer_b200 import _lib
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", _lib.SO_PATH], capture_output=True, text=True, check=True).stdout
    assert " FADD2 " in sass
    assert " FFMA2 " not in sass and " FMUL2 " not in sass


def test_readback_host_expansion_rebuilds_the_table(vb):
    """The host half of the sparse read-back (csrc/readback_host.cpp) runs without a GPU: {word index, value} pairs -> table,
    every byte written, for both the AVX-512 and the portable path."""
    from cuda_voxelizer_b200 import _lib
    L = C.CDLL(_lib.SO_PATH)
    rng = np.random.default_rng(3)
    words = 16 * 5000
    for density in (0.0, 0.001, 0.05, 0.4, 1.0):
        want = np.zeros(words, np.uint32)
        idx = np.flatnonzero(rng.random(words) < density)
        want[idx] = rng.integers(1, 2**32, len(idx), dtype=np.uint64).astype(np.uint32)
        pairs = np.stack([idx.astype(np.uint32), want[idx]], axis=1).copy()
        for sym in ("_ZN4voxb21readback_expand_sliceEPjmmPKvmmb", "_ZN4voxb26readback_expand_slice_sse2EPjmmPKvmmb"):
            fn = getattr(L, sym)
            fn.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_bool]
            fn.restype = None
            raw = np.full(words + 32, 0xDEADBEEF, np.uint32)
            off = (-(raw.ctypes.data // 4)) % 16
            table = raw[off:off + words]
            # two slices, cut on a line boundary in the middle of the pairs: the first gets the full pass, the second is zero-filled
            # ahead (no pairs) and then only gets its non-zero lines
            cut = 16 * 2311
            pc = int(np.searchsorted(idx, cut))
            fn(table.ctypes.data, 0, cut, pairs.ctypes.data, 0, pc, False)
            fn(table.ctypes.data, cut, words, None, 0, 0, False)
            assert not table[cut:].any()
            fn(table.ctypes.data, cut, words, pairs.ctypes.data, pc, len(idx), True)
            assert np.array_equal(table, want), (density, sym)
            assert raw[off + words] == 0xDEADBEEF and (off == 0 or raw[off - 1] == 0xDEADBEEF)


def test_readback_host_pool_selftest(vb):
    """The read-back's worker pool — started, cancelled, rebuilt with other thread counts — and the zero-fill that runs ahead
    (a rebuilt pool once ran the previous job again and hung the call: csrc/readback.cu HostPool)."""
    from cuda_voxelizer_b200 import _lib
    code = "import ctypes,sys; L=ctypes.CDLL(%r); sys.exit(L.voxb200_selftest_host_pool())" % _lib.SO_PATH
    r = subprocess.run([os.sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
