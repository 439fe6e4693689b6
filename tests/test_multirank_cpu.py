"""N>1 path on CPU (gloo, world_size 2 and 3): the box-range partition of the 4-taxon-set space is
a disjoint cover, and summing per-rank partial edge tables with one all-reduce reproduces the
single-rank result. The per-rank work is done by the oracle restricted to the quartets that fall
into the rank's box range (the GPU does the same with camus_b200_set_partition)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from camus_b200 import partition as part
from camus_b200 import synth
from camus_b200.hostlib import Problem
from oracle import oracle as orc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, scheme=None):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if scheme:
        os.environ["CAMUS_B200_PARTITION"] = scheme
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        s = synth.make(26, 40, 77, drop_frac=0.1)
        prob = Problem(s.constraint, s.genes)
        o = orc.Oracle(1).load(prob)
        o.count_filter(2, 0.3, False)
        keys, counts = o.export_quartets()
        r_of_tip = part.dfs_rank_of_tip(prob.parent, prob.child0, prob.child1, prob.tip_index)
        owner = part.owner_of_quartets(keys, r_of_tip, world)
        mine = owner == rank
        o.set_quartets(keys[mine], counts[mine])
        partial = torch.from_numpy(o.quartet_totals(False, False).astype(np.int64))
        scal = torch.tensor([int(mine.sum()), int(counts[mine].sum())], dtype=torch.int64)
        part.allreduce_partials(partial, scal)
        if rank == 0:
            o.set_quartets(keys, counts)
            full = o.quartet_totals(False, False).astype(np.int64)
            q.put((bool(np.array_equal(partial.numpy(), full)), int(scal[0]) == len(keys), int(scal[1]) == int(counts.sum()),
                   int(mine.sum())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,scheme", [(2, None), (3, None), (4, "class")])
def test_partial_tables_sum_to_the_full_table(world, scheme):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, scheme)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    same, ns_ok, sum_ok, n0 = q.get(timeout=10)
    assert same and ns_ok and sum_ok and n0 > 0


def test_box_ranges_cover_the_index_space():
    for n_taxa in (4, 9, 26, 200, 1000):
        nb = part.num_boxes(n_taxa)
        for world in (1, 2, 3, 4, 8):
            counts = [part.local_box_count(nb, r, world) for r in range(world)]
            assert sum(counts) == nb
            if nb <= 30000:
                own = [part.owner_of_box(b, nb, world) for b in range(nb)]
                assert [own.count(r) for r in range(world)] == counts
            if nb >= 64 * world:
                assert max(counts) - min(counts) <= part.partition_chunk(nb, world)
    m = 5
    seen = sorted(part.box_index(a, b, c, d) for d in range(m) for c in range(d + 1) for b in range(c + 1) for a in range(b + 1))
    assert seen == list(range(part.choose(m + 3, 4)))


@pytest.mark.parametrize("world", [4, 8])
def test_class_scheme_is_a_cover_and_ranks_read_only_their_tile_classes(world):
    """Class scheme (layout.cuh class_owner): every box has one owner; the four triple tiles of a
    box are led by its first two segments, whose classes must be in the owner's tile mask (the
    rank builds no other tiles); box counts are even; half (3/4 for two ranks of 4) of the tiles
    are built per rank."""
    for m in (7, 33, 40):
        sa, sb, sc, sd = part.all_boxes(m)
        own = part.class_owner(sa, sb, sc, sd, m, world)
        assert own.min() >= 0 and own.max() < world and len(own) == part.choose(m + 3, 4)
        for r in range(world):
            need = part.class_need_mask(r, world)
            sel = own == r
            assert np.all((need >> part.seg_class(sa[sel])) & 1) and np.all((need >> part.seg_class(sb[sel])) & 1)
        if m >= 33:
            cnt = np.bincount(own, minlength=world) * world / len(own)
            if world == 8:  # every rank builds half of the tiles: even box counts
                assert cnt.max() < 1.06 and cnt.min() > 0.94
            else:  # ranks 2, 3 build 3/4 of the tiles instead of 1/2 and get T/8 boxes fewer for it
                quarter_tiles = part.choose(m + 2, 3) / 8 * world / len(own)
                assert abs(cnt[0] - cnt[1]) < 0.03 and abs(cnt[2] - cnt[3]) < 0.03
                assert abs((cnt[0] + cnt[1] - cnt[2] - cnt[3]) / 2 - quarter_tiles) < 0.04
            lead = part.seg_class(np.array([a for c in range(m) for b in range(c + 1) for a in range(b + 1)]))
            frac = [float((((part.class_need_mask(r, world) >> lead) & 1) != 0).mean()) for r in range(world)]
            assert max(frac) < (0.78 if world == 4 else 0.53)


def test_scheme_rule(monkeypatch):
    monkeypatch.delenv("CAMUS_B200_PARTITION", raising=False)
    assert part.use_class_scheme(1000, 8) and part.use_class_scheme(1000, 4) and part.use_class_scheme(256, 8)
    assert not part.use_class_scheme(1000, 2) and not part.use_class_scheme(200, 8) and not part.use_class_scheme(1000, 1)
    monkeypatch.setenv("CAMUS_B200_PARTITION", "chunk")
    assert not part.use_class_scheme(1000, 8)
    monkeypatch.setenv("CAMUS_B200_PARTITION", "class")
    assert part.use_class_scheme(40, 8) and not part.use_class_scheme(40, 3)
