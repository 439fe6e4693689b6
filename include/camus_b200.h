/* camus_b200.h -- C-ABI of the B200-native CAMUS hot path (libcamus_b200.so).
 *
 * The reference (jsdoublel/camus, pure Go) has no FFI; the seam is cut at the Go function bodies
 * listed below. Every exported Go signature stays as it is; these entry points
 * replace the BODIES of:
 *
 *   camus_b200_set_constraint   <- graphs.MakeTreeData tables          internal/graphs/treedata.go:27-51
 *   camus_b200_add_gene_trees   <- input of prep.processQuartets       internal/prep/preprocess.go:71-101
 *   camus_b200_count_filter     <- prep.processQuartets                internal/prep/preprocess.go:71-124
 *                                  + prep.filterQuartets               internal/prep/quartetfilter.go:76-96
 *                                  + constraint-quartet removal        internal/prep/preprocess.go:51-57
 *                                  + TotalNumQuartets / ..Unique..     internal/graphs/treedata.go:297-307
 *   camus_b200_quartet_totals   <- score.CalculateQuartetTotals        internal/score/quartettotals.go:43-61
 *                                  (+ graphs.mapQuartetsToVertices     internal/graphs/treedata.go:197-222)
 *   camus_b200_edge_penalties   <- score.CalculateEdgePenalties        internal/score/penalty.go:11-29
 *   camus_b200_export_quartets  <- the map[Quartet]uint32 itself       internal/graphs/quartet.go:11-31
 *   camus_b200_reticulation_counts <- the per-gene-tree loop of score.ReticulationScore
 *                                                                       internal/score/score.go:31-60
 *   camus_b200_add_gene_trees_newick <- readGeneTreesFile's parse loop + per-tree preprocessing  internal/prep/io.go:133-153, preprocess.go:87-100
 *   camus_b200_dp_across        <- the scoreEdgesAcross loop of scoreAddEdgeK  internal/infer/main_dp.go:199-214, :247-287
 *
 * Conventions: plain pointers + sizes, host memory, valid for the duration of the call only (cgo
 * rule: no pointer is kept after return). All functions return 0 on success, non-zero on failure
 * (message via camus_b200_last_error); the library never aborts the process. One context per
 * device; a context is not re-entrant. Calls are synchronous. There is NO CPU fallback: without a
 * usable CUDA device camus_b200_create fails.
 */
#ifndef CAMUS_B200_H
#define CAMUS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct camus_b200 camus_b200;

enum {
  CAMUS_B200_OK = 0,
  CAMUS_B200_ERR_ARG = 1,     /* bad argument / call order */
  CAMUS_B200_ERR_CUDA = 2,    /* CUDA runtime failure */
  CAMUS_B200_ERR_NOMEM = 3,   /* problem does not fit the device */
  CAMUS_B200_ERR_INPUT = 4,   /* malformed flat tree */
  CAMUS_B200_ERR_FALLBACK = 5 /* camus_b200_add_gene_trees_newick only: the text needs the caller's own parser (nothing was added) */
};

/* One handle for n_gpus GPUs of this process (1, 2, 4, 8). devices[i] = CUDA
 * ordinal, NULL = 0..n_gpus-1. With n_gpus > 1 every call below fans out: each GPU gets all gene
 * trees and 1/n_gpus of the 4-taxon-set index space, and the partial edge tables are summed on
 * devices[0] by peer reads over NVLink. Results are bit-identical for any n_gpus. One process per
 * GPU (n_gpus = 1 each + camus_b200_set_partition + a caller-side all-reduce) is the alternative. */
int camus_b200_create(camus_b200** out, int n_gpus, const int* devices);
void camus_b200_destroy(camus_b200* ctx);
/* ctx may be NULL: returns the last error of a failed camus_b200_create on this thread. */
const char* camus_b200_last_error(const camus_b200* ctx);

/* Multi-process sharding of the 4-taxon-set index space: this context counts / filters / scores
 * only partition `rank` of `world`. Every rank must be given ALL gene trees. After
 * camus_b200_count_filter and camus_b200_quartet_totals_partial the caller sums the partial
 * buffers over ranks (NCCL all-reduce on the device pointers returned below) and then calls
 * camus_b200_quartet_totals_finish. Default is (0, 1). A handle with n_gpus > 1 deals its
 * partition on to its GPUs (world * n_gpus shares); its partial buffer lives on devices[0].
 * How the boxes of the index space are dealt is the library's business (chunks round-robin, or for
 * 4 / 8 shares by segment classes so that a share builds only the bit-planes it reads; DESIGN.md 7,
 * mirrored for tests by camus_b200/partition.py); the union of the shares is always the whole space
 * and the summed results do not depend on it. */
int camus_b200_set_partition(camus_b200* ctx, int rank, int world);

/* Constraint tree in reference node-id order (preprocess.go:28-30), rooted, binary.
 * parent[i] = -1 at the root; child0/child1 = -1 at tips, child order = TreeData.Children order;
 * tip_index[i] = constraint tip index (gotree UpdateTipIndex) at tips, -1 at internal nodes. */
int camus_b200_set_constraint(camus_b200* ctx, int32_t n_nodes, int32_t n_taxa, const int32_t* parent,
                              const int32_t* child0, const int32_t* child1, const int32_t* tip_index);

/* Gene trees as CSR over nodes, flattened by the host AFTER UpdateTipIndex, the missing-taxa check
 * and CollapseLowSupport (preprocess.go:87-100). Per tree: nodes in preorder (parent before
 * child), parent[] = index local to the tree or -1 at its root, taxon[] = constraint tip index at
 * leaves, -1 at internal nodes. Multifurcations, unifurcations and missing taxa are allowed;
 * rooting does not matter. May be called several times; trees accumulate. */
int camus_b200_add_gene_trees(camus_b200* ctx, int32_t n_trees, const int64_t* node_off /*[n_trees+1]*/,
                              const int32_t* parent, const int32_t* taxon);
/* Drop all gene trees (the constraint stays). */
int camus_b200_clear_gene_trees(camus_b200* ctx);

/* Count every gene-tree quartet topology, apply the -q mode (0/1/2) and -t threshold filter,
 * remove the topologies the constraint tree displays. n_survivors = len(qCounts)
 * (preprocess.go:58), sum_counts = exact sum of surviving counts (Go truncates it to uint32,
 * treedata.go:297-303). With a partition set these are this rank's partial values. */
int camus_b200_count_filter(camus_b200* ctx, int qmode, double threshold, uint64_t* n_survivors,
                            uint64_t* sum_counts);

/* Optional: tell the next count_filter which weight (-asSet or counts) quartet_totals will ask
 * for, so its fused count+filter+score pass already produces the right table (scorers.go:23-38:
 * the Go side knows AsSet before Preprocess). A mismatch only costs a second pass. */
int camus_b200_set_score_hint(camus_b200* ctx, int as_set);

/* quartetTotals[u][w] (row-major n*n, reference node ids), 0 where !ShouldCalcEdge.
 * as_set != 0 counts every surviving quartet once (-asSet, and always in sym mode). */
int camus_b200_quartet_totals(camus_b200* ctx, int as_set, uint64_t* out /*[n*n]*/);
/* penalties[u][w] (row-major n*n), 0 where !ShouldCalcEdge. Needs only the constraint tree. */
int camus_b200_edge_penalties(camus_b200* ctx, uint64_t* out /*[n*n]*/);

/* The constraint-tree tables graphs.MakeTreeData builds on the host (internal/graphs/treedata.go:27-51),
 * so that the Go side can skip calcLCAs -- Theta(n^3), ~8e9 inner iterations at n = 1999
 * (treedata.go:130-166) -- calcDepths (:169-178) and countLeavesBelow (:181-194):
 *   lca[u * n + w]   = node id of the lowest common ancestor of u and w (TreeData.lca, symmetric, lca[u][u] = u)
 *   depth[u]         = edges from the root (TreeData.Depths)
 *   leaves_below[u]  = tips in the subtree of u (TreeData.NumLeavesBelow)
 * all in the caller's node ids. Any of the three pointers may be NULL. Needs only set_constraint. */
int camus_b200_constraint_tables(camus_b200* ctx, int32_t* lca /*[n*n]*/, int32_t* depth /*[n]*/, uint64_t* leaves_below /*[n]*/);

/* Surviving quartets in the reference's packing (quartet.go:11-31), sorted by key. Only available
 * when the whole cell table fits on the device (parity / debug; TreeData.Quartets on small
 * inputs). Call with cap = 0 to query n_out. raw != 0 exports the counts before filter/removal. */
int camus_b200_export_quartets(camus_b200* ctx, int raw, uint64_t* keys, uint32_t* counts, uint64_t cap,
                               uint64_t* n_out);

/* Parity probe for sizes whose cell table cannot be exported: raw (pre-filter) counts of n_query
 * arbitrary 4-taxon sets, read from the bit-plane table with an independent scalar kernel.
 * quads = 4 distinct constraint tip indices per query; counts = 3 per query, in the reference's
 * topology order for the taxa sorted by tip index (0b1100, 0b1010, 0b0110; quartet.go:23-25). */
int camus_b200_query_cells(camus_b200* ctx, int32_t n_query, const int32_t* quads, uint32_t* counts);

/* score.ReticulationScore (internal/score/score.go:25-72): for every gene tree and every
 * reticulation of a given level-1 network, how many emissions of the quartet enumerator
 * (gtre.Quartets(false, ..) after gtre.UnRoot(), duplicates included, score.go:38-43) are comparable
 * with the reticulation (QuartetScore != Qdiff -> totals) and how many agree with it (== Qeq ->
 * supported). The Go side keeps ConvertToNetwork, MakeTreeData, the Level1 check and
 * getReticulationNodes, flattens what QuartetScore reads about the network into one record per
 * (reticulation r, network tip t) and divides supported / totals in float64 afterwards:
 *   records[(r * n_tips + t) * 4 + 0] = node id of LCA(w_r, t)      (quartettotals.go:121)
 *                                 + 1 = node id of LCA(u_r, t)      (quartettotals.go:123)
 *                                 + 2 = depth of the first | depth of the second << 16
 *                                 + 3 = bit 0: t under w_r, 1: under v_r, 2: under wSub_r, 3: under u_r
 * with node id 0 = the root. Gene trees: CSR as in camus_b200_add_gene_trees, taxon = tip index in
 * the NETWORK tree (its #H tips included, graphs.MapIDsFromConstTree against ntw.NetTree). The call
 * is independent of the constraint / gene trees loaded into ctx; any n_ret (batches of 64 inside); gene trees
 * of at most 24576 nodes. Outputs are
 * [n_trees * n_ret], row-major by gene tree. */
int camus_b200_reticulation_counts(camus_b200* ctx, int32_t n_tips, int32_t n_ret, const uint32_t* records,
                                   int32_t n_trees, const int64_t* node_off, const int32_t* parent,
                                   const int32_t* taxon, uint64_t* totals, uint64_t* supported);

/* ---- multi-process runs, collective owned by the library (NCCL over NVLink / NVSwitch) ----
 * One process per GPU: rank 0 calls camus_b200_comm_unique_id and hands the 128 bytes to the other
 * processes by its own means; every process then calls camus_b200_comm_init (collective, blocking),
 * which also sets the partition (rank, world). From then on camus_b200_count_filter returns the
 * GLOBAL n_survivors / sum_counts and camus_b200_quartet_totals / camus_b200_run_resident the GLOBAL
 * table on every rank: the partial junction tables are summed with one ncclAllReduce(int64, sum)
 * on the library's stream, the two scalars with another. libnccl.so.2 is loaded at run time
 * (dlopen); without it comm_init fails and everything else works as before. Single-GPU handles only. */
#define CAMUS_B200_COMM_ID_BYTES 128
int camus_b200_comm_unique_id(uint8_t id[CAMUS_B200_COMM_ID_BYTES]);
int camus_b200_comm_init(camus_b200* ctx, const uint8_t id[CAMUS_B200_COMM_ID_BYTES], int rank, int world);

/* ---- split form of quartet_totals for multi-process runs with a CALLER-side reduction ---- */
int camus_b200_quartet_totals_partial(camus_b200* ctx, int as_set);
/* device pointer + element count of the partial junction table (int64 elements) to be summed */
int camus_b200_partial_buffer(camus_b200* ctx, void** dev_ptr, uint64_t* n_int64);
int camus_b200_quartet_totals_finish(camus_b200* ctx, uint64_t* out /*[n*n]*/);

/* ---- instrumentation (bench.py) ---- */
enum {
  CAMUS_B200_T_UPLOAD = 0,   /* host->device copies of the flat trees            */
  CAMUS_B200_T_PREP = 1,     /* tree walk + LCA-depth tables + triple bit-planes */
  CAMUS_B200_T_COUNT = 2,    /* quartet count kernel                             */
  CAMUS_B200_T_FILTER = 3,   /* filter + constraint removal + totals             */
  CAMUS_B200_T_SCORE = 4,    /* edge-score scatter                               */
  CAMUS_B200_T_FINAL = 5,    /* scans + mask + copy-out                          */
  CAMUS_B200_T_PENALTY = 6,
  CAMUS_B200_T_N = 8
};
/* Re-run the whole hot path (LCA tables -> bit-planes -> count -> filter -> score -> finalise)
 * from the flat gene trees ALREADY resident in HBM (uploaded by an earlier count_filter): the
 * device-resident leg of the benchmark. Results as count_filter + quartet_totals. out == NULL
 * stops after the partial junction table (multi-process: all-reduce it, then call ..._finish). */
int camus_b200_run_resident(camus_b200* ctx, int qmode, double threshold, int as_set, uint64_t* out /*[n*n]*/,
                            uint64_t* n_survivors, uint64_t* sum_counts);
/* CUDA-event milliseconds of the last run of each stage and the number of kernel launches. */
int camus_b200_stage_ms(const camus_b200* ctx, float* ms /*[CAMUS_B200_T_N]*/, uint64_t* n_launches);
/* resolutions = sum over gene trees of C(leaves,4): the unit of the throughput metric. */
int camus_b200_workload(const camus_b200* ctx, uint64_t* resolutions, uint64_t* n_cells, uint64_t* n_boxes);

/* ---- gene-tree ingest on the GPU ----
 * For newick gene-tree files: replaces the parse loop of prep.readGeneTreesFile (internal/prep/io.go:133-153),
 * the per-tree UpdateTipIndex / missing-taxa check / CollapseLowSupport(minSupp, true) of prep.processQuartets
 * (internal/prep/preprocess.go:87-100), graphs.MapIDsFromConstTree (internal/graphs/quartet.go:131-149) and the
 * flattening in front of camus_b200_add_gene_trees: the file's bytes go to the device, the flat trees are built
 * there (csrc/kernels_ingest.cuh) and appended exactly as camus_b200_add_gene_trees would append them.
 *   set_tip_labels   the constraint tree's tip names in tip-index order (ascending byte order, gotree
 *                    UpdateTipIndex) as one byte pool + offsets; call after camus_b200_set_constraint
 *   add_gene_trees_newick   text = the file: one tree per non-blank line. min_support = the -s flag (0 = no
 *                    collapse). *missing_taxa != 0: some tree lacks a constraint taxon (the reference prints its
 *                    warning once, preprocess.go:90-97).
 * Returns CAMUS_B200_ERR_FALLBACK, having added NOTHING, when a line is outside the supported dialect
 * (single-quoted labels; with -s != 0 support values that are not plain decimals of <= 15 significant
 * digits), malformed, holds a label the constraint does not have (graphs.ErrTipNameMismatch) or a label
 * twice (prep.ErrMulTree): *bad_tree = its 1-based index, camus_b200_last_error says why, and the caller runs
 * its own parser + camus_b200_add_gene_trees, which reports the reference's error text. */
int camus_b200_set_tip_labels(camus_b200* ctx, int32_t n_taxa, const char* labels, const int64_t* label_off /*[n_taxa+1]*/);
int camus_b200_add_gene_trees_newick(camus_b200* ctx, const char* text, uint64_t len, double min_support,
                                     int32_t* n_trees, int32_t* missing_taxa, int32_t* bad_tree);
/* parity / debug: the flat gene trees the context holds (as passed to, or built by, the calls above). Any of the
 * arrays may be NULL; *n_trees / *n_nodes are always set. cap_nodes = room in parent[] / taxon[]. */
int camus_b200_get_gene_trees(camus_b200* ctx, int64_t* node_off /*[n_trees+1]*/, int32_t* parent, int32_t* taxon, uint64_t cap_nodes,
                              int32_t* n_trees, uint64_t* n_nodes);

/* ---- DP offload: the across-scan of the level-1 network DP ----
 * Replaces the loop body of scoreAddEdgeK that calls scoreEdgesAcross for every u
 * (internal/infer/main_dp.go:199-214 + :247-287) and, inside it, FourWayBestSplit
 * (internal/infer/helpers.go:83-110). The DP driver stays with the caller: RunDP's post-order walk,
 * solve's k loop, cycleDP.update, scoreEdgesDown and the traceback (main_dp.go:40-176, :219-244).
 * Score type = the Scorer's S: uint64 (max) or float64 (norm, sym), passed as raw 8-byte values;
 * float results are bit-identical to Go's because every sum is formed in the reference's order.
 * A separate handle (its own stream, about 8 n^2 bytes of HBM), independent of camus_b200.
 *
 *   create        node ids = the constraint tree's preorder ids (root 0); subtree_end[x] = one past
 *                 the last id of x's subtree; depth = TreeData.Depths; edge_scores[u*n + w] =
 *                 Scorer.CalcScore(u, w) for EVERY pair (the across-scan does not call ShouldCalcEdge)
 *   set_node      DP[node] once solve(node) has returned (tips are preset to {0})
 *   set_vertex    start of solve(v): the (u, other child subtree) pairs in SubtreePostOrder order
 *   set_path_columns  cycleDP.scores[x][k] for k0 <= k < k1 and first <= x < first + count,
 *                 cols[(k - k0) * count + (x - first)]; every list holds prev_k + 1 entries when
 *                 across(prev_k) is called
 *   across        the candidate the reference's loop over SubtreePostOrder(v) ends with when it
 *                 starts from an empty best: maximum score, then the smallest CycleLength, then the
 *                 last u in visiting order; inside one u the first maximum over w in preorder.
 *                 NaN scores (norm mode: 0 / 0 on the length-3 cycle between v's children) follow
 *                 the reference's comparisons: they never win the fold above; when the candidate of
 *                 the FIRST u is NaN it is returned in first_*, and the caller takes it iff it holds
 *                 no candidate yet (bestCycleTrace == nil, main_dp.go:194). */
typedef struct camus_b200_dp camus_b200_dp;
typedef struct {
  int32_t found;         /* 0: no (u, w) pair at this vertex */
  int32_t nan_seen;
  uint64_t score_bits;   /* uint64 score, or the IEEE bits of the float64 score */
  int32_t u, w;
  int32_t cycle_length;  /* score.CycleLength(u, w) */
  int32_t split[4];      /* wPathK, uPathK, wDownK, uDownK (main_dp.go:266) */
  int32_t first_is_nan;  /* the first u's candidate has a NaN score: first_* describe it */
  int32_t first_u, first_w;
  int32_t first_split[4];
  uint64_t first_score_bits;
} camus_b200_dp_best;
int camus_b200_dp_create(camus_b200_dp** out, int device, int score_is_f64, int32_t n_nodes, const int32_t* subtree_end,
                         const int32_t* depth, const void* edge_scores /*[n*n] 8-byte scores*/);
void camus_b200_dp_destroy(camus_b200_dp* dp);
const char* camus_b200_dp_last_error(const camus_b200_dp* dp);
int camus_b200_dp_set_node(camus_b200_dp* dp, int32_t node, int32_t len, const void* scores /*[len]*/);
int camus_b200_dp_set_vertex(camus_b200_dp* dp, int32_t v, int32_t n_items, const int32_t* item_u, const int32_t* item_sub);
int camus_b200_dp_set_path_columns(camus_b200_dp* dp, int32_t k0, int32_t k1, int32_t first, int32_t count, const void* cols);
int camus_b200_dp_across(camus_b200_dp* dp, int32_t prev_k, camus_b200_dp_best* out);
/* Integer scores (max mode) only: the same step as a max-plus product, O(k) per pair instead of O(k^2). The caller
 * uploads conv_x[k] = max over i + i' = k of cycleDP.scores[x][i] + DP[x][i'] (columns as in set_path_columns; no path
 * columns are needed) and gets the same winner -- maximum score, smallest CycleLength, last u, first w -- with
 * split[] = -1: it derives the four split indices of that ONE candidate with the reference's FourWayBestSplit. */
int camus_b200_dp_set_conv_columns(camus_b200_dp* dp, int32_t k0, int32_t k1, int32_t first, int32_t count, const void* cols);
int camus_b200_dp_across_maxplus(camus_b200_dp* dp, int32_t prev_k, camus_b200_dp_best* out);
uint64_t camus_b200_dp_launches(const camus_b200_dp* dp);

#ifdef __cplusplus
}
#endif
#endif /* CAMUS_B200_H */
