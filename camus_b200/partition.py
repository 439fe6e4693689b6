"""Host-side mirror of the cell / box index math of csrc/layout.cuh and of the multi-rank split.

The 4-taxon-set index space is cut into 8x8x8x8 boxes over DFS-ranked taxa; the boxes are dealt
to the ranks in chunks, round-robin (camus_b200_set_partition; layout.cuh BoxMap). Every rank
holds all gene trees, so counting needs no exchange; the only collective is one sum of the partial
junction tables (torch.distributed all_reduce over NCCL) plus two scalars.
"""
from __future__ import annotations

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


def owner_of_box(b: int, nb: int, world: int) -> int:
    return (b // partition_chunk(nb, world)) % world


def local_box_count(nb: int, rank: int, world: int) -> int:
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
    """Which rank's box range holds each packed quartet key (graphs/quartet.go packing)."""
    n_taxa = len(rank_of_tip)
    nb = num_boxes(n_taxa)
    owners = np.empty(len(keys), np.int64)
    for i, q in enumerate(keys):
        q = int(q)
        ranks = sorted(int(rank_of_tip[(q >> (15 * j)) & 0x7FFF]) for j in range(4))
        b = box_index(*(r // SEG for r in ranks))
        owners[i] = owner_of_box(b, nb, world)
    return owners


def allreduce_partials(table, scalars, group=None):
    """The one exchange step of a multi-rank run: sum the partial junction table (an int64 tensor
    on the rank's device) and the (n_survivors, sum_counts) pair over all ranks, in place."""
    import torch.distributed as dist
    dist.all_reduce(table, group=group)
    dist.all_reduce(scalars, group=group)
    return table, scalars
