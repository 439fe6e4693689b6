"""camus_b200/partition.py (the Python mirror used by the multi-rank tests) against the index math of
camus_b200/csrc/layout.cuh itself, compiled for the host with g++ (tests/layout_probe.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from camus_b200 import partition as part

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("layout") / "liblayout_probe.so")
    subprocess.run(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "layout_probe.cpp")], check=True)
    L = C.CDLL(so)
    L.probe_box_index.restype = C.c_uint64
    L.probe_tile_index.restype = C.c_uint64
    L.probe_partition_chunk.restype = C.c_uint64
    L.probe_partition_chunk.argtypes = [C.c_uint64, C.c_uint]
    L.probe_local_box_count.restype = C.c_uint64
    L.probe_local_box_count.argtypes = [C.c_uint64, C.c_uint, C.c_uint]
    L.probe_box_unrank.argtypes = [C.c_uint64, C.POINTER(C.c_uint)]
    L.probe_plane_work_list.restype = C.c_uint64
    L.probe_plane_work_list.argtypes = [C.c_uint, C.c_uint, C.c_int, C.POINTER(C.c_uint)]
    return L


def test_class_scheme_mirror_is_identical(probe):
    assert [probe.probe_seg_class(s) for s in range(40)] == part.seg_class(np.arange(40)).tolist()
    for world in (4, 8):
        assert [probe.probe_class_need_mask(r, world) for r in range(world)] == [part.class_need_mask(r, world) for r in range(world)]
        for m in (1, 2, 7, 16, 33, 40, 63, 125, 200, 256):
            assert probe.probe_class_threshold(m, world) == part.class_threshold(m, world), (m, world)
        for m in (1, 5, 12, 33):
            sa, sb, sc, sd = part.all_boxes(m)
            own = part.class_owner(sa, sb, sc, sd, m, world)
            got = [probe.probe_class_owner(int(a), int(b), int(c), int(d), m, world) for a, b, c, d in zip(sa, sb, sc, sd)]
            assert got == own.tolist(), (m, world)


def test_box_and_tile_ranks(probe):
    m = 9
    sa, sb, sc, sd = part.all_boxes(m)
    out = (C.c_uint * 4)()
    for r, (a, b, c, d) in enumerate(zip(sa, sb, sc, sd)):
        assert probe.probe_box_index(int(a), int(b), int(c), int(d)) == r == part.box_index(int(a), int(b), int(c), int(d))
        probe.probe_box_unrank(r, out)
        assert list(out) == [a, b, c, d]
    # large ranks: the fp32 estimate + integer fix-up of box_unrank stays exact
    for seg in ((0, 0, 0, 4095), (4095, 4095, 4095, 4095), (17, 900, 901, 2500), (124, 124, 124, 124)):
        r = probe.probe_box_index(*seg)
        probe.probe_box_unrank(r, out)
        assert tuple(out) == seg
    t = 0
    for s2 in range(m):
        for s1 in range(s2 + 1):
            for s0 in range(s1 + 1):
                assert probe.probe_tile_index(s0, s1, s2) == t
                t += 1


def test_chunk_scheme_mirror(probe):
    for n_taxa in (9, 200, 1000):
        nb = part.num_boxes(n_taxa)
        for world in (1, 2, 3, 5, 8):
            assert probe.probe_partition_chunk(nb, world) == part.partition_chunk(nb, world)
            for r in range(world):
                assert probe.probe_local_box_count(nb, r, world) == part.local_box_count(nb, r, world)


def test_plane_work_list(probe):
    """The sliced plane builder's work list (layout.cuh plane_work_list): every tile the share builds exactly
    once, blocks of 16 x 16 x 16 segments, and slots = the tile index, or for a class-scheme share the position of
    the tile among the share's tiles in colex order (how set_box_range numbers the compact bit-plane storage)."""
    for m in (1, 5, 16, 17, 40, 125):
        n_all = m * (m + 1) * (m + 2) // 6
        for need, compact in ((0xF, False), (0x3, True), (0x9, True), (0xD, True), (0xF, True)):
            n = probe.probe_plane_work_list(m, need, int(compact), None)
            buf = (C.c_uint * (4 * max(n, 1)))()
            assert probe.probe_plane_work_list(m, need, int(compact), buf) == n
            e = np.frombuffer(buf, dtype=np.uint32)[: 4 * n].reshape(n, 4).astype(np.int64)
            s0, s1, s2, slot = e.T
            assert np.all(s0 <= s1) and np.all(s1 <= s2) and np.all(s2 < m)
            tile = s2 * (s2 + 1) * (s2 + 2) // 6 + s1 * (s1 + 1) // 2 + s0
            assert len(np.unique(tile)) == n
            wanted = np.array([t for t in range(n_all)], dtype=np.int64)
            cls = part.seg_class(np.arange(m))
            # colex enumeration of all tiles with the class filter
            a2, a1, a0 = [], [], []
            for x2 in range(m):
                for x1 in range(x2 + 1):
                    for x0 in range(x1 + 1):
                        a2.append(x2); a1.append(x1); a0.append(x0)
            keep = ((need >> cls[np.array(a0)]) & 1).astype(bool)
            assert sorted(tile.tolist()) == wanted[keep].tolist()
            if compact:
                rank_in_share = np.full(n_all, -1, dtype=np.int64)
                rank_in_share[wanted[keep]] = np.arange(int(keep.sum()))
                assert np.array_equal(slot, rank_in_share[tile])
            else:
                assert np.array_equal(slot, tile)
            # blocked order: block keys never decrease along the list
            key = ((s2 // 16) * 4096 + (s1 // 16)) * 4096 + s0 // 16
            assert np.all(np.diff(key) >= 0)


def test_sliced_row_layout_is_conflict_free(probe):
    """Shared-memory layout of the sliced plane builder (layout.cuh slice_off): every 128-bit access pattern of
    k_triple_planes_bs touches 8 different 16-byte bank groups per quarter-warp (or the same address), the rows do
    not overlap and stay inside the array. Lane maps as in kernels_prep.cuh: staging lane t copies chunk t + 64 j
    (row = chunk / Q, q = chunk % Q); compute lane t owns micro-block xp = (t >> 2) & 3, yp = t & 3, z pair t >> 4."""
    for W in (8, 12):
        Q = W // 4
        words = probe.probe_slice_array_words(W)
        off = {(r, q): probe.probe_slice_off(W, r, q) for r in range(64) for q in range(Q)}
        assert len(set(off.values())) == 64 * Q and all(o % 4 == 0 and 0 <= o and o + 4 <= words for o in off.values())

        def check(addresses):  # one warp-wide 128-bit access: 4 quarter-warps of 8 lanes
            for k in range(0, 32, 8):
                distinct = set(addresses[k:k + 8])
                groups = {(a // 4) % 8 for a in distinct}
                assert len(groups) == len(distinct), (W, addresses[k:k + 8])

        for warp in range(2):  # staging: 64 threads, Q chunks each
            for j in range(Q):
                chunks = [warp * 32 + lane + 64 * j for lane in range(32)]
                check([off[(c // Q, c % Q)] for c in chunks])
        for warp in range(2):  # compute: zq = 2 warp + (lane >> 4)
            lanes = [(lane >> 2) & 3 for lane in range(32)], [lane & 3 for lane in range(32)], [2 * warp + (lane >> 4) for lane in range(32)]
            for q in range(Q):
                for dx in range(2):
                    for dy in range(2):
                        check([off[((2 * xp + dx) * 8 + 2 * yp + dy, q)] for xp, yp, zq in zip(*lanes)])  # ab rows
                    for zz in range(2):
                        check([off[((2 * xp + dx) * 8 + 2 * zq + zz, q)] for xp, yp, zq in zip(*lanes)])  # ac rows
                for dy in range(2):
                    for zz in range(2):
                        check([off[((2 * yp + dy) * 8 + 2 * zq + zz, q)] for xp, yp, zq in zip(*lanes)])  # bc rows
