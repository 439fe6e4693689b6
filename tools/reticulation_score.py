#!/usr/bin/env python
"""Per-gene-tree support of the reticulations of a level-1 network: score.ReticulationScore
(internal/score/score.go:25-72) + prep.WriteRetScoresToCSV (internal/prep/io.go:315-340) on one B200.
The reference reaches this function from its tests only; this is the same computation as a tool.

  python tools/reticulation_score.py network.nwk gene-trees.nwk [--format newick|nexus] > scores.csv
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from camus_b200.cabi import CamusB200  # noqa: E402
from camus_b200.hostlib import NetworkProblem  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("network")
ap.add_argument("gene_trees")
ap.add_argument("--format", default="newick", choices=["newick", "nexus"])
ap.add_argument("--device", type=int, default=0)
a = ap.parse_args()
net = NetworkProblem.from_files(a.network, a.gene_trees, a.format)
gpu = CamusB200(a.device)
totals, supported = gpu.reticulation_counts(net)
print(net.csv(totals, supported))
gpu.close()
