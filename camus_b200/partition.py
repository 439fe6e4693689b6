"""Host-side mirror of the cell / box index math of csrc/layout.cuh and of the multi-rank split.

The 4-taxon-set index space is cut into 8x8x8x8 boxes over DFS-ranked taxa. Two ways to deal the
boxes to the ranks (camus_b200_set_partition; layout.cuh BoxMap):
  * chunk scheme (any world): chunks of boxes, round-robin; every rank builds every triple tile;
  * class scheme (world 4 or 8, 32..256 segments): a rank owns boxes by the classes of their first
    two segments and builds only the tiles led by those classes (1/2 or 3/4 of the bit-planes).
Every rank holds all gene trees, so counting needs no exchange; the only collective is one sum of
the partial junction tables (torch.distributed all_reduce over NCCL) plus two scalars.
"""
from __future__ import annotations

import os

import numpy as np

SEG = 8


def choose(n: int, k: int) -> int:
    if n < k:
        return 0
    r = 1
    for i in range(k):
        r = r * (n - i) // (i + 1)
    return r


def num_boxes(n_taxa: int) -> int:
    m = (n_taxa + SEG - 1) // SEG
    return choose(m + 3, 4)


def box_index(sa: int, sb: int, sc: int, sd: int) -> int:
    return choose(sd + 3, 4) + choose(sc + 2, 3) + choose(sb + 1, 2) + sa


def partition_chunk(nb: int, world: int) -> int:
    """Boxes per chunk (layout.cuh partition_chunk): chunks are dealt to the ranks round-robin."""
    return max(1, min(2048, nb // (world * 64)))


def num_segments(n_taxa: int) -> int:
    return (n_taxa + SEG - 1) // SEG


def use_class_scheme(n_taxa: int, world: int) -> bool:
    """camus_b200.cu use_class_scheme, including the CAMUS_B200_PARTITION override."""
    if world not in (4, 8):
        return False
    m = num_segments(n_taxa)
    e = os.environ.get("CAMUS_B200_PARTITION")
    if e == "chunk":
        return False
    if e == "class":
        return m <= 256
    return 32 <= m <= 256


# ---- class scheme (layout.cuh seg_class .. class_threshold) ----
_PAT = np.array([0, 1, 2, 3, 3, 2, 1, 0], np.int64)
_NEED = {8: (0x3, 0x5, 0x9, 0x6, 0xA, 0xC, 0x3, 0xC), 4: (0x3, 0xC, 0xD, 0xE)}


def seg_class(s):
    return _PAT[np.asarray(s, np.int64) & 7]


def class_need_mask(rank: int, world: int) -> int:
    return _NEED[world][rank]


def class_threshold(m: int, world: int) -> int:
    sa, sb = np.meshgrid(np.arange(m), np.arange(m), indexing="ij")
    k = (m - sb) * (m - sb + 1) // 2
    valid = sa <= sb
    same = seg_class(sa) == seg_class(sb)
    D = int(k[valid & same].sum())
    O = int(k[valid & ~same].sum())
    if D == 0 or 3 * D <= O:
        return 0
    if world == 8:
        return min(85, (64 * (3 * D - O) + (3 * D) // 2) // (3 * D))
    # world 4: ranks 2 and 3 build 3/4 of the tiles, ranks 0 and 1 half: they get T/8 boxes fewer
    T = choose(m + 2, 3)
    have6, want6 = 3 * D - O, 3 * T // 4
    if have6 <= want6:
        return 0
    return min(255, (512 * (have6 - want6) + (9 * D) // 2) // (9 * D))


def _pair_rank(lo, hi):
    return np.where(lo == 0, hi - 1, np.where(lo == 1, hi + 1, 5))


def class_owner(sa, sb, sc, sd, m: int, world: int) -> np.ndarray:
    """Owner rank of boxes (sa <= sb <= sc <= sd), vectorised (layout.cuh class_owner)."""
    sa, sb, sc, sd = (np.asarray(x, np.int64) for x in (sa, sb, sc, sd))
    t = class_threshold(m, world)
    ca, cb = seg_class(sa), seg_class(sb)
    lo, hi = np.minimum(ca, cb), np.maximum(ca, cb)
    h = (sc * 7 + sd * 13) & 255
    pure = lo == hi
    if world == 8:
        k = h // max(t, 1)
        o = np.where(k < lo, k, k + 1)
        donated = _pair_rank(np.minimum(o, lo), np.maximum(o, lo))
        own_pure = np.where(h < 3 * t, donated, np.where(lo < 2, 6, 7))
        return np.where(pure, own_pure, _pair_rank(lo, hi))
    mixed = np.where(hi < 2, 0, np.where(lo >= 2, 1, np.where(lo == 0, 2, 3)))
    own_pure = np.where(h < t, np.where((lo == 0) | (lo == 2), 2, 3), np.where(lo < 2, 0, 1))
    return np.where(pure, own_pure, mixed)


def owner_of_boxes(sa, sb, sc, sd, n_taxa: int, world: int) -> np.ndarray:
    """Owner rank of boxes given by their four segments, under the scheme the library picks."""
    sa, sb, sc, sd = (np.asarray(x, np.int64) for x in (sa, sb, sc, sd))
    if use_class_scheme(n_taxa, world):
        return class_owner(sa, sb, sc, sd, num_segments(n_taxa), world)
    b = (sd + 3) * (sd + 2) * (sd + 1) * sd // 24 + (sc + 2) * (sc + 1) * sc // 6 + (sb + 1) * sb // 2 + sa
    return (b // partition_chunk(num_boxes(n_taxa), world)) % world


def all_boxes(m: int):
    """Segments of every box in colex (= box_index) order."""
    out = [(a, b, c, d) for d in range(m) for c in range(d + 1) for b in range(c + 1) for a in range(b + 1)]
    return tuple(np.array(x, np.int64) for x in zip(*out)) if out else tuple(np.zeros(0, np.int64) for _ in range(4))


def owner_of_box(b: int, nb: int, world: int) -> int:
    """Chunk scheme only."""
    return (b // partition_chunk(nb, world)) % world


def local_box_count(nb: int, rank: int, world: int) -> int:
    """Chunk scheme only."""
    c = partition_chunk(nb, world)
    full, rem = divmod(nb, c * world)
    return full * c + max(0, min(c, rem - rank * c))


def dfs_rank_of_tip(parent, child0, child1, tip_index) -> np.ndarray:
    """DFS rank (order of appearance in a preorder walk, child0 first) of every constraint tip index."""
    n = len(parent)
    root = int(np.nonzero(np.asarray(parent) < 0)[0][0])
    out = np.full(int(np.max(tip_index)) + 1, -1, np.int64)
    stack, k = [root], 0
    while stack:
        x = stack.pop()
        if child0[x] < 0:
            out[tip_index[x]] = k
            k += 1
        else:
            stack.append(int(child1[x]))
            stack.append(int(child0[x]))
    assert k == len(out) and n == 2 * k - 1
    return out


def owner_of_quartets(keys: np.ndarray, rank_of_tip: np.ndarray, world: int) -> np.ndarray:
    """Which rank's share of the boxes holds each packed quartet key (graphs/quartet.go packing)."""
    n_taxa = len(rank_of_tip)
    keys = np.asarray(keys, np.uint64)
    r = np.stack([np.asarray(rank_of_tip, np.int64)[((keys >> np.uint64(15 * j)) & np.uint64(0x7FFF)).astype(np.int64)]
                  for j in range(4)], axis=1) if len(keys) else np.zeros((0, 4), np.int64)
    seg = np.sort(r, axis=1) // SEG
    return owner_of_boxes(seg[:, 0], seg[:, 1], seg[:, 2], seg[:, 3], n_taxa, world).astype(np.int64)


def allreduce_partials(table, scalars, group=None):
    """The one exchange step of a multi-rank run: sum the partial junction table (an int64 tensor
    on the rank's device) and the (n_survivors, sum_counts) pair over all ranks, in place."""
    import torch.distributed as dist
    dist.all_reduce(table, group=group)
    dist.all_reduce(scalars, group=group)
    return table, scalars
