"""Deterministic synthetic inputs for the configurations of BASELINE.json.

The reference ships no simulator. Recipe: labels t0000...; RNG = SplitMix64 seeded
0xCA3B200 + config id; constraint tree = uniform random rooted binary tree built by repeatedly
joining two random roots; gene tree = the constraint tree after Poisson(N/10) random rooted NNI
moves (ILS-like noise) and, for 30 % of the trees, one of 3 fixed "reticulation" SPR moves (so the
two minor topologies are asymmetric and `-q 2` keeps something). Options add per-edge support
values U(0,1) (for `-s`), random taxon deletions, and NEXUS output.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

MASK = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & MASK

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK
        return z ^ (z >> 31)

    def below(self, n: int) -> int:
        return self.next() % n

    def uniform(self) -> float:
        return (self.next() >> 11) / float(1 << 53)

    def poisson(self, lam: float) -> int:
        L, k, p = math.exp(-min(lam, 700.0)), 0, 1.0
        while True:
            p *= self.uniform()
            if p <= L:
                return k
            k += 1


class T:
    """Rooted binary tree on parent / children arrays; leaves carry a taxon id."""
    __slots__ = ("parent", "kids", "taxon", "root")

    def __init__(self):
        self.parent, self.kids, self.taxon, self.root = [], [], [], -1

    def add(self, taxon=-1):
        self.parent.append(-1)
        self.kids.append([])
        self.taxon.append(taxon)
        return len(self.parent) - 1

    def copy(self):
        t = T()
        t.parent = list(self.parent)
        t.kids = [list(k) for k in self.kids]
        t.taxon = list(self.taxon)
        t.root = self.root
        return t

    def is_ancestor(self, a, x):
        while x >= 0:
            if x == a:
                return True
            x = self.parent[x]
        return False


def random_constraint(n_taxa: int, rng: SplitMix64) -> T:
    t = T()
    roots = [t.add(i) for i in range(n_taxa)]
    while len(roots) > 1:
        i = rng.below(len(roots))
        a = roots.pop(i)
        j = rng.below(len(roots))
        b = roots.pop(j)
        p = t.add()
        t.kids[p] = [a, b]
        t.parent[a] = t.parent[b] = p
        roots.append(p)
    t.root = roots[0]
    return t


def nni(t: T, rng: SplitMix64, internal: list[int]):
    """Swap a random child of a random internal non-root node with that node's sibling."""
    for _ in range(8):
        v = internal[rng.below(len(internal))]
        p = t.parent[v]
        if p < 0:
            continue
        s = t.kids[p][0] if t.kids[p][1] == v else t.kids[p][1]
        ci = rng.below(2)
        c = t.kids[v][ci]
        t.kids[v][ci] = s
        t.parent[s] = v
        t.kids[p][t.kids[p].index(s)] = c
        t.parent[c] = p
        return


def spr(t: T, w: int, u: int):
    """Prune the subtree at w and regraft it on the edge above u (no-op if impossible)."""
    pw = t.parent[w]
    if pw < 0 or t.parent[u] < 0 or t.is_ancestor(w, u) or u == pw:
        return
    sib = t.kids[pw][0] if t.kids[pw][1] == w else t.kids[pw][1]
    if sib == u:
        return
    gp = t.parent[pw]
    # remove pw: sib takes its place
    if gp < 0:
        t.root = sib
        t.parent[sib] = -1
    else:
        t.kids[gp][t.kids[gp].index(pw)] = sib
        t.parent[sib] = gp
    # reuse pw as the new node above u
    pu = t.parent[u]
    if pu < 0:
        t.root = pw
        t.parent[pw] = -1
    else:
        t.kids[pu][t.kids[pu].index(u)] = pw
        t.parent[pw] = pu
    t.kids[pw] = [u, w]
    t.parent[u] = pw
    t.parent[w] = pw


def newick(t: T, labels: list[str], rng: SplitMix64 | None = None, support: bool = False, drop: set[int] | None = None) -> str:
    out = []
    # iterative post-order emission
    stack = [(t.root, 0)]
    pieces: dict[int, str] = {}
    while stack:
        x, st = stack.pop()
        if not t.kids[x]:
            pieces[x] = "" if (drop and t.taxon[x] in drop) else labels[t.taxon[x]]
            continue
        if st == 0:
            stack.append((x, 1))
            for c in t.kids[x]:
                stack.append((c, 0))
        else:
            parts = [pieces.pop(c) for c in t.kids[x]]
            parts = [p for p in parts if p]
            if not parts:
                pieces[x] = ""
            elif len(parts) == 1:
                pieces[x] = parts[0]  # suppress the unifurcation left by a dropped taxon
            else:
                s = "(" + ",".join(parts) + ")"
                if support and x != t.root:
                    s += "%.3f" % rng.uniform()
                pieces[x] = s
    out.append(pieces[t.root])
    return out[0] + ";"


@dataclass
class Synth:
    constraint: str
    genes: str
    fmt: str
    n_taxa: int
    n_trees: int


def make(n_taxa: int, n_trees: int, config_id: int = 0, support: bool = False, drop_frac: float = 0.0,
         nexus: bool = False, nni_rate: float | None = None) -> Synth:
    rng = SplitMix64(0xCA3B200 + config_id)
    labels = ["t%04d" % i for i in range(n_taxa)]
    ct = random_constraint(n_taxa, rng)
    internal = [i for i in range(len(ct.parent)) if ct.kids[i]]
    nodes = list(range(len(ct.parent)))
    moves = []
    while len(moves) < 3:
        w, u = nodes[rng.below(len(nodes))], nodes[rng.below(len(nodes))]
        if w != u and ct.parent[w] >= 0 and ct.parent[u] >= 0 and not ct.is_ancestor(w, u) and not ct.is_ancestor(u, w):
            moves.append((w, u))
    lam = n_taxa / 10.0 if nni_rate is None else nni_rate
    lines = []
    for _ in range(n_trees):
        g = ct.copy()
        if rng.uniform() < 0.3:
            w, u = moves[rng.below(3)]
            spr(g, w, u)
        for _ in range(rng.poisson(lam)):
            nni(g, rng, internal)
        drop = None
        if drop_frac > 0:
            drop = {i for i in range(n_taxa) if rng.uniform() < drop_frac}
            if n_taxa - len(drop) < 4:
                drop = None
        lines.append(newick(g, labels, rng, support, drop))
    cons = newick(ct, labels)
    if nexus:
        body = "#NEXUS\n\nBEGIN TREES;\n\n" + "".join("Tree g%d = %s\n" % (i + 1, l) for i, l in enumerate(lines)) + "\nEND;\n"
        return Synth(cons, body, "nexus", n_taxa, n_trees)
    return Synth(cons, "\n".join(lines) + "\n", "newick", n_taxa, n_trees)


# The configurations of BASELINE.json ("configs"), in order.
CONFIGS = {
    1: dict(n_taxa=50, n_trees=200, qmode=2, threshold=0.5, mode="max", as_set=False, min_support=0.0),
    2: dict(n_taxa=200, n_trees=1000, qmode=2, threshold=0.5, mode="max", as_set=False, min_support=0.0),
    3: dict(n_taxa=1000, n_trees=1000, qmode=2, threshold=0.5, mode="max", as_set=False, min_support=0.0),
    4: dict(n_taxa=500, n_trees=1000, qmode=2, threshold=0.5, mode="max", as_set=True, min_support=0.7,
            support=True, drop_frac=0.1, nexus=True),
    5: dict(n_taxa=300, n_trees=2000, qmode=2, threshold=0.5, mode="sym", as_set=True, min_support=0.0),
}


def make_config(cid: int, n_taxa: int | None = None, n_trees: int | None = None, drop_frac: float | None = None) -> tuple[Synth, dict]:
    c = dict(CONFIGS[cid])
    if drop_frac is not None:  # a variant of the configuration in which every tree misses that fraction of the taxa
        c["drop_frac"] = drop_frac
    if n_taxa:
        c["n_taxa"] = n_taxa
    if n_trees:
        c["n_trees"] = n_trees
    s = make(c["n_taxa"], c["n_trees"], cid, support=c.get("support", False), drop_frac=c.get("drop_frac", 0.0),
             nexus=c.get("nexus", False))
    return s, c
