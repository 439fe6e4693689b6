#!/bin/sh
# Run in the build container (needs /root/reference). Copies the reference's test DATA files.
set -e
R=/root/reference/internal
D=$(dirname "$0")/ref
mkdir -p "$D"
cp $R/infer/testdata/constraint.nwk "$D/infer_constraint.nwk"
gzip -9 -n -c $R/infer/testdata/gene-trees.nwk > "$D/infer_gene-trees.nwk.gz"
for f in network.nwk net_q2_t05_max.nwk net_q2_t05_norm.nwk net_q2_t05_sym_a01.nwk; do cp $R/infer/testdata/$f "$D/$f"; done
cp $R/prep/testdata/constraint.nwk "$D/prep_constraint.nwk"
gzip -9 -n -c $R/prep/testdata/g100.nwk > "$D/prep_g100.nwk.gz"
for f in quartets.nex quartets.nwk unresolved.nwk; do cp $R/prep/testdata/$f "$D/$f"; done
cp $R/score/testdata/network.nwk "$D/score_network.nwk"
cp $R/score/testdata/scores.csv "$D/score_scores.csv"
chmod 644 "$D"/*
