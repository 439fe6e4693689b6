// Host-side constraint-tree data and flattening for the C-ABI (include/camus_b200.h).
//
// Mirrors the host half of prep.Preprocess (internal/prep/preprocess.go:26-44, 87-100) -- the
// validation and gotree mutations that happen BEFORE the hot path -- and the pieces of
// graphs.TreeData (internal/graphs/treedata.go:11-51, 234-295) that the DP, which stays on the
// host, still needs: Children, Depths, LCA/Under, NLeaves. The quartet tables themselves
// (quartetSet / quartetCounts, treedata.go:197-222) live on the device.
#pragma once
#include <array>
#include <cstdio>
#include <functional>

#include "tree.hpp"

namespace camus {

// Flat constraint tree in reference node-id order (preorder, root = 0).
struct ConstraintFlat {
  int32_t n_nodes = 0, n_taxa = 0;
  std::vector<int32_t> parent, child0, child1, tip_index;
};

// Gene trees as CSR over nodes: per-tree preorder, parent = local index or -1, taxon = constraint
// tip index at leaves, -1 at internal nodes.
struct GeneFlat {
  std::vector<int64_t> node_off{0};
  std::vector<int32_t> parent, taxon;
  int32_t n_trees() const { return (int32_t)node_off.size() - 1; }
};

// What the DP needs from graphs.TreeData.
struct TreeData {
  Tree tree;                     // the constraint tree (after RemoveSingleNodes)
  std::vector<int> id_to_node;   // id -> index into tree.nodes        (treedata.go IdToNodes)
  std::vector<int> node_to_id;   // index into tree.nodes -> id, -1 if detached
  std::vector<std::array<int, 2>> children;  // by id; {-1,-1} at tips (treedata.go:70-83)
  std::vector<int> parent;       // by id
  std::vector<int> depths;       // by id                              (treedata.go:169-178)
  std::vector<int> subtree_end;  // by id: ids in [id, subtree_end) are the subtree (preorder)
  std::vector<uint64_t> num_leaves_below;  // by id                    (treedata.go:181-194)
  int n_leaves = 0;
  std::vector<std::string> tip_names;      // by tip index (sorted labels)

  int n() const { return (int)id_to_node.size(); }
  bool tip(int id) const { return children[id][0] < 0; }
  // td.Under(n1, n2): n2 strictly under n1 (treedata.go:291-294)
  bool under(int n1, int n2) const { return n2 > n1 && n2 < subtree_end[n1]; }
  int lca(int a, int b) const {  // td.LCA (treedata.go:239-242)
    while (!(a <= b && b < subtree_end[a])) a = parent[a];
    return a;
  }
  int sibling(int id) const {  // td.Sibling (treedata.go:245-258)
    int p = parent[id];
    return children[p][0] == id ? children[p][1] : children[p][0];
  }
};

// score.CycleLength (internal/score/quartettotals.go:67-74)
inline int cycle_length(int u, int w, const TreeData& td) {
  int v = td.lca(u, w);
  int len = (td.depths[u] - td.depths[v]) + (td.depths[w] - td.depths[v]) + 1;
  if (v == u) len += 1;
  return len;
}
// score.ShouldCalcEdge (internal/score/quartettotals.go:63-65)
inline bool should_calc_edge(int u, int w, const TreeData& td) {
  return !td.under(w, u) && cycle_length(u, w, td) > 3 && u != 0 && w != 0;
}

// Host half of prep.Preprocess for the constraint tree (preprocess.go:27-39).
inline TreeData make_tree_data(Tree tre) {
  tre.remove_single_nodes();
  if (!tre.update_tip_index()) throw CamusError(Err::MulTree, "constraint tree contains duplicate labels");
  if (!tre.rooted()) throw CamusError(Err::Unrooted, "constraint tree is not rooted");
  if (!tre.is_binary()) throw CamusError(Err::NonBinary, "constraint tree is not binary");
  TreeData td;
  td.tree = std::move(tre);
  td.id_to_node = td.tree.preorder();
  int n = (int)td.id_to_node.size();
  td.node_to_id.assign(td.tree.nodes.size(), -1);
  for (int i = 0; i < n; ++i) td.node_to_id[td.id_to_node[i]] = i;
  td.children.assign(n, {-1, -1});
  td.parent.assign(n, -1);
  td.depths.assign(n, 0);
  td.subtree_end.assign(n, 0);
  td.num_leaves_below.assign(n, 0);
  for (int i = 0; i < n; ++i) {
    const Node& nd = td.tree.nodes[td.id_to_node[i]];
    if (nd.parent >= 0) {
      td.parent[i] = td.node_to_id[nd.parent];
      td.depths[i] = td.depths[td.parent[i]] + 1;
    }
    if (!nd.tip()) td.children[i] = {td.node_to_id[nd.children[0]], td.node_to_id[nd.children[1]]};
  }
  for (int i = n - 1; i >= 0; --i) {
    if (td.tip(i)) {
      td.subtree_end[i] = i + 1;
      td.num_leaves_below[i] = 1;
      ++td.n_leaves;
    } else {
      td.subtree_end[i] = td.subtree_end[td.children[i][1]];
      td.num_leaves_below[i] = td.num_leaves_below[td.children[i][0]] + td.num_leaves_below[td.children[i][1]];
    }
  }
  td.tip_names.assign(td.n_leaves, "");
  for (int i = 0; i < n; ++i) {
    const Node& nd = td.tree.nodes[td.id_to_node[i]];
    if (nd.tip()) td.tip_names[nd.tip_index] = nd.name;
  }
  return td;
}

inline ConstraintFlat flatten_constraint(const TreeData& td) {
  ConstraintFlat f;
  f.n_nodes = td.n();
  f.n_taxa = td.n_leaves;
  f.parent.assign(td.parent.begin(), td.parent.end());
  f.child0.resize(f.n_nodes);
  f.child1.resize(f.n_nodes);
  f.tip_index.resize(f.n_nodes);
  for (int i = 0; i < f.n_nodes; ++i) {
    f.child0[i] = td.children[i][0];
    f.child1[i] = td.children[i][1];
    f.tip_index[i] = td.tree.nodes[td.id_to_node[i]].tip_index;
  }
  return f;
}

// Host half of prep.processQuartets for the gene trees (preprocess.go:87-100) plus
// graphs.MapIDsFromConstTree (quartet.go:131-149): validate, collapse, map labels, flatten.
// `warn` receives the one-off missing-taxa warning (preprocess.go:93-96).
inline GeneFlat flatten_gene_trees(std::vector<Tree>& gene_trees, const TreeData& td, double min_supp,
                                   const std::function<void(const std::string&)>& warn = nullptr) {
  std::unordered_map<std::string, int> name2idx;
  for (int i = 0; i < td.n_leaves; ++i) name2idx[td.tip_names[i]] = i;
  GeneFlat g;
  bool warned = false;
  for (size_t gi = 0; gi < gene_trees.size(); ++gi) {
    Tree& gt = gene_trees[gi];
    if (!gt.update_tip_index())
      throw CamusError(Err::MulTree, "gene tree on line " + std::to_string(gi + 1) + " : contains duplicate labels");
    std::vector<int> pre = gt.preorder();
    int ntips = 0;
    for (int x : pre) ntips += gt.nodes[x].tip();
    if (ntips != td.n_leaves && !warned) {
      warned = true;
      if (warn) warn("WARNING: missing taxa detected in one or more gene trees; this may cause issues with some scoring metrics");
    }
    if (min_supp != 0) {
      gt.collapse_low_support(min_supp);
      pre = gt.preorder();
    }
    std::vector<int> local(gt.nodes.size(), -1);
    for (size_t i = 0; i < pre.size(); ++i) local[pre[i]] = (int)i;
    for (int x : pre) {
      const Node& nd = gt.nodes[x];
      g.parent.push_back(nd.parent >= 0 ? local[nd.parent] : -1);
      if (nd.tip()) {
        auto it = name2idx.find(nd.name);
        if (it == name2idx.end())
          throw CamusError(Err::TipNameMismatch,
                           "tip name mismatch! maybe the gene tree and constraint tree labels don't match?, "
                           "no tip named " + nd.name + " in the constraint tree");
        g.taxon.push_back(it->second);
      } else {
        g.taxon.push_back(-1);
      }
    }
    g.node_off.push_back((int64_t)g.parent.size());
  }
  return g;
}

// prep.percentNoSupport (preprocess.go:179-189)
inline double percent_no_support(const std::vector<Tree>& trees) {
  long total = 0, count = 0;
  for (const Tree& t : trees)
    for (int x : t.preorder()) {
      const Node& nd = t.nodes[x];
      if (nd.parent < 0 || nd.tip()) continue;
      if (nd.has_support) ++count;
      ++total;
    }
  return double(total - count) / double(total) * 100;
}

}  // namespace camus
