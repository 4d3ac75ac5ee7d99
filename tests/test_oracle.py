"""CPU suite: pins the oracle (oracle/vox_oracle.c) against the golden vectors generated from the
reference itself (tests/golden/make_golden.py) and, where the reference build is present
(oracle/_ref/libvoxref.so — this container and, prebuilt, the GPU box), against it directly."""
import numpy as np
import pytest

import cases
import oracle

SMALL = [c for c in cases.GOLDEN_CASES if c[1] <= 256 and not c[0].startswith("soup:sliver")]


def _oracle_table(name, g, solid, morton):
    v, f = cases.mesh(name)
    mn, mx, unit = oracle.voxinfo(v, g)
    fn = oracle.solid if solid else oracle.surface
    return fn(oracle.soup(v, f), mn, unit, g, morton), (mn, mx, unit)


@pytest.mark.parametrize("name,g,solid,morton", SMALL, ids=[cases.case_key(*c) for c in SMALL])
def test_oracle_matches_reference_golden(golden, name, g, solid, morton):
    want = golden[cases.case_key(name, g, solid, morton)]
    table, (mn, mx, unit) = _oracle_table(name, g, solid, morton)
    assert [float(x) for x in mn] == want["bbox_min"]
    assert [float(x) for x in mx] == want["bbox_max"]
    assert [float(x) for x in unit] == want["unit"]
    assert oracle.popcount(table) == want["popcount"]
    assert "%016x" % oracle.fnv1a64(table) == want["fnv1a64"]
    assert oracle.solid_ub_events() == 0      # fixtures never enter the reference's UB territory


def test_small_tables_bit_for_bit():
    d = np.load(cases.GOLDEN_DIR + "/tables_small.npz")
    assert len(d.files) >= 4
    for key in d.files:
        name, g, kind, order = key.split("|")
        table, _ = _oracle_table(name, int(g), kind == "solid", order == "morton")
        assert np.array_equal(table, d[key]), key


@pytest.mark.skipif(not oracle.have_ref(), reason="reference build (oracle/_ref) not present")
@pytest.mark.parametrize("name,g,solid,morton", [("bunny", 64, 0, 0), ("bunny", 128, 1, 0), ("bunny", 64, 1, 1),
                                                  ("icosphere:16:64", 128, 1, 0), ("soup:mixed:2000:1:64", 64, 0, 1),
                                                  ("torus:100:50:256", 256, 0, 0),
                                                  # G = 257: the reference's table is one bit short (main.cpp:190) and its far-corner
                                                  # write lands in the word past it; the oracle's buffers hold that word
                                                  ("box:10", 257, 0, 0), ("box:10", 257, 1, 0)])
def test_oracle_equals_compiled_reference(name, g, solid, morton):
    v, f = cases.mesh(name)
    ref_table = oracle.ref_voxelize(v, f, g, solid=solid, morton=morton, threads=1)
    table, (mn, mx, unit) = _oracle_table(name, g, solid, morton)
    rmn, rmx, runit = oracle.ref_voxinfo(v, g, len(f))
    assert np.array_equal(mn, rmn) and np.array_equal(mx, rmx) and np.array_equal(unit, runit)
    assert np.array_equal(table, ref_table)


@pytest.mark.skipif(not oracle.have_ref(), reason="reference build (oracle/_ref) not present")
def test_reference_is_thread_count_independent():
    v, f = cases.mesh("bunny")
    a = oracle.ref_voxelize(v, f, 64, solid=True, threads=1)
    b = oracle.ref_voxelize(v, f, 64, solid=True, threads=4)
    assert np.array_equal(a, b)


def test_slab_restriction_composes():
    """z-range restricted runs OR/XOR together to the unrestricted table (what z-slab sharding relies on)."""
    v, f = cases.mesh("bunny")
    g = 64
    mn, mx, unit = oracle.voxinfo(v, g)
    tris = oracle.soup(v, f)
    for fn in (oracle.surface, oracle.solid):
        full = fn(tris, mn, unit, g)
        acc = np.zeros_like(full)
        for z0, z1 in ((0, 16), (16, 40), (40, 64)):
            part = fn(tris, mn, unit, g, z_range=(z0, z1))
            words = g * g // 32
            assert not part[: z0 * words].any() and not part[z1 * words:].any()
            acc |= part
        assert np.array_equal(acc, full)


def test_morton_is_bit_interleave():
    rng = np.random.default_rng(0)
    for x, y, z in rng.integers(0, 1 << 16, (200, 3)):
        want = 0
        for i in range(16):
            want |= ((int(x) >> i) & 1) << (3 * i) | ((int(y) >> i) & 1) << (3 * i + 1) | ((int(z) >> i) & 1) << (3 * i + 2)
        assert oracle.morton(x, y, z) == want


def test_layout_msb_first():
    """checkVoxel (util.h:25-38): voxel (x,y,z) is bit 31-(idx%32) of word idx/32, idx = x + G*y + G*G*z."""
    g = 64
    t = np.zeros(g * g * g // 32, np.uint32)
    idx = 5 + g * 7 + g * g * 9
    t[idx // 32] = np.uint32(1) << np.uint32(31 - idx % 32)
    assert oracle.lib().oracle_check_voxel(5, 7, 9, g, t) == 1
    assert oracle.lib().oracle_check_voxel(6, 7, 9, g, t) == 0
    assert oracle.lib().oracle_table_bytes(g) == g ** 3 // 8


def test_morton_table_is_permutation_of_linear():
    v, f = cases.mesh("bunny")
    g = 64
    mn, mx, unit = oracle.voxinfo(v, g)
    tris = oracle.soup(v, f)
    lin = oracle.surface(tris, mn, unit, g)
    mor = oracle.surface(tris, mn, unit, g, morton=True)
    bits = np.unpackbits(lin.view(np.uint8).reshape(-1, 4)[:, ::-1].reshape(-1))      # bit k = voxel idx k
    z, y, x = np.nonzero(bits.reshape(g, g, g))
    mbits = np.unpackbits(mor.view(np.uint8).reshape(-1, 4)[:, ::-1].reshape(-1))
    midx = np.array([oracle.morton(a, b, c) for a, b, c in zip(x, y, z)])
    assert mbits.sum() == len(midx) and mbits[midx].all()


def test_slab_mode_equals_the_words_of_the_full_table():
    """oracle.surface_slab (a slab-sized table with an origin: what the 8192^3 GPU tests compare against) against the full run."""
    v, f = cases.mesh("bunny")
    mn, mx, unit = oracle.voxinfo(v, 128)
    soup = oracle.soup(v, f)
    full = oracle.surface(soup, mn, unit, 128)
    w = 128 * 128 // 32
    for z0, z1 in ((0, 16), (48, 96), (112, 128)):
        assert np.array_equal(oracle.surface_slab(soup, mn, unit, 128, z0, z1), full[z0 * w: z1 * w])
    assert np.array_equal(oracle.surface(soup, mn, unit, 128), full)        # the origin is back at 0
