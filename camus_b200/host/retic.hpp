// Host side of score.ReticulationScore (internal/score/score.go:25-95): how well each gene tree
// supports each reticulation of a GIVEN level-1 network. SURVEY.md 8(f) rank 4; in the reference
// it is reachable from tests only (score_test.go), not from the CLI.
//
// Mirrors, in this order:
//   prep.ConvertToNetwork          internal/prep/io.go:173-236   extended newick tree -> u/w ids per #H label
//   prep.NetworkIsBinary/isBinary  internal/prep/preprocess.go:138-147,160-173
//   graphs.MakeTreeData(NetTree)   internal/graphs/treedata.go:27-51  (depths, leafsets, LCA; the network
//                                  tree keeps its #H tips and the unary #H nodes above every w)
//   (*Network).Level1 / illSorted  internal/graphs/net.go:103-124
//   getReticulationNodes           internal/score/score.go:75-95
// and then FLATTENS QuartetScore's tree look-ups (internal/score/quartettotals.go:107-157) into one
// 16-byte record per (reticulation, network tip), so that the device kernel (csrc/kernels_retic.cuh)
// needs no tree at all. The quartet enumeration + scoring itself is the C-ABI call
// camus_b200_reticulation_counts (CUDA); nothing here counts quartets.
//
// Node ids: positions in the preorder walk of the network tree, root = 0. The reference
// uses the ids gotree's parser assigned; only identity and "0 is the root" matter to the result
// (io.go:203,216,229 treat id 0 as "unset", quartettotals.go:119 as "outside the cycle").
#pragma once
#include <map>
#include <unordered_map>

#include "tree.hpp"

namespace camus {

constexpr int kErrNoReticulations = 11;  // prep.ErrNoReticulations (io.go:29-33)
constexpr int kErrNotLevel1 = 12;        // score.ErrNotLevel1 (score.go:15)

struct RetError : std::runtime_error {
  int code;
  RetError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

struct NetworkData {
  Tree tree;
  int n = 0, n_tips = 0;
  std::vector<int> node_of_id, id_of_node, parent, depth, size;  // by id (preorder): subtree(x) = [x, x + size[x])
  std::vector<std::vector<int>> children;                        // by id, newick order
  std::vector<int> tip_node;                                     // tip index -> id
  std::vector<std::string> tip_names;                            // by tip index
  std::vector<std::string> labels;                               // sorted as WriteRetScoresToCSV does (io.go:320-325)
  std::vector<int> u, w, v, wsub;                                // per label

  bool under_or_eq(int anc, int x) const { return x >= anc && x < anc + size[anc]; }
  bool under(int n1, int n2) const { return n1 != n2 && under_or_eq(n1, n2); }  // treedata.go:291-294
  int lca(int a, int b) const {
    while (a != b) {
      if (depth[a] >= depth[b]) a = parent[a];
      else b = parent[b];
    }
    return a;
  }
  bool in_leafset(int node, int tip) const { return under_or_eq(node, tip_node[tip]); }
};

namespace detail {
// preprocess.go:160-173 with allowUnifurcations = true
inline bool is_binary_net(const Tree& t, int x) {
  const Node& nd = t.nodes[x];
  if (nd.tip()) return true;
  size_t nneigh = nd.children.size() + 1;
  if (nneigh == 2) return is_binary_net(t, nd.children[0]);
  if (nneigh != 3) return false;
  return is_binary_net(t, nd.children[0]) && is_binary_net(t, nd.children[1]);
}
}  // namespace detail

// prep.ConvertToNetwork + graphs.MakeTreeData + Level1 + getReticulationNodes
inline NetworkData make_network_data(Tree t) {
  NetworkData nd;
  if (!t.rooted()) throw CamusError(Err::Unrooted, "network is not rooted");
  if (!(detail::is_binary_net(t, t.nodes[t.root].children[0]) && detail::is_binary_net(t, t.nodes[t.root].children[1])))
    throw CamusError(Err::NonBinary, "network is not binary");
  std::vector<int> pre = t.preorder();
  nd.n = (int)pre.size();
  nd.node_of_id = pre;
  nd.id_of_node.assign(t.nodes.size(), -1);
  for (int i = 0; i < nd.n; ++i) nd.id_of_node[pre[i]] = i;

  // io.go:181-219: PostOrder, labels containing '#'
  std::map<std::string, std::array<int, 2>> ret;
  for (int x : t.postorder()) {
    const Node& cur = t.nodes[x];
    if (cur.name.find('#') == std::string::npos) continue;
    auto& br = ret.try_emplace(cur.name, std::array<int, 2>{0, 0}).first->second;
    int v = -1;
    if (cur.tip()) {
      if (cur.parent < 0) throw CamusError(Err::InvalidFormat, "reticulation label on the root");
      for (int c : t.nodes[cur.parent].children)  // prev's neighbours other than cur and prev's parent: the last one wins
        if (c != x) v = c;
      if (br[0] != 0 || v < 0)
        throw CamusError(Err::InvalidFormat, "too many or invalid matching reticulation label " + cur.name);
      br[0] = nd.id_of_node[v];
    } else {
      for (int c : cur.children) v = c;
      if (br[1] != 0 || v < 0)
        throw CamusError(Err::InvalidFormat, "too many or invalid matching reticulation label " + cur.name);
      br[1] = nd.id_of_node[v];
    }
  }
  if (ret.empty()) throw RetError(kErrNoReticulations, "no reticulations - not a network");
  for (auto& kv : ret)
    if (kv.second[0] == 0 || kv.second[1] == 0)
      throw CamusError(Err::InvalidFormat, "label " + kv.first + " is unmatched");
  if (!t.update_tip_index()) throw CamusError(Err::MulTree, "network contains duplicate labels");

  // graphs.MakeTreeData on the network tree as parsed (treedata.go:27-51)
  nd.parent.assign(nd.n, -1);
  nd.depth.assign(nd.n, 0);
  nd.size.assign(nd.n, 1);
  nd.children.assign(nd.n, {});
  for (int i = 0; i < nd.n; ++i) {
    const Node& x = t.nodes[pre[i]];
    if (x.parent >= 0) {
      nd.parent[i] = nd.id_of_node[x.parent];
      nd.depth[i] = nd.depth[nd.parent[i]] + 1;
    }
    for (int c : x.children) nd.children[i].push_back(nd.id_of_node[c]);
    if (x.tip()) ++nd.n_tips;
  }
  for (int i = nd.n - 1; i > 0; --i) nd.size[nd.parent[i]] += nd.size[i];
  nd.tip_node.assign(nd.n_tips, -1);
  nd.tip_names.assign(nd.n_tips, "");
  for (int i = 0; i < nd.n; ++i) {
    const Node& x = t.nodes[pre[i]];
    if (x.tip()) {
      nd.tip_node[x.tip_index] = i;
      nd.tip_names[x.tip_index] = x.name;
    }
  }

  // net.go:103-124
  std::vector<std::pair<std::string, std::array<int, 2>>> rs(ret.begin(), ret.end());
  auto ill_sorted = [&](int v1, int v2, const std::array<int, 2>& r1) {
    return nd.under(v1, v2) && (nd.under(v2, r1[0]) || nd.under(v2, r1[1]));
  };
  for (size_t i = 0; i < rs.size(); ++i)
    for (size_t j = i + 1; j < rs.size(); ++j) {
      int v1 = nd.lca(rs[i].second[0], rs[i].second[1]), v2 = nd.lca(rs[j].second[0], rs[j].second[1]);
      if (v1 == v2 || ill_sorted(v1, v2, rs[i].second) || ill_sorted(v2, v1, rs[j].second))
        throw RetError(kErrNotLevel1, "network is not level-1");
    }

  // io.go:320-325 column order: by length, then bytes
  std::sort(rs.begin(), rs.end(), [](const auto& a, const auto& b) {
    if (a.first.size() != b.first.size()) return a.first.size() < b.first.size();
    return a.first < b.first;
  });
  for (auto& kv : rs) {
    int u = kv.second[0], w = kv.second[1], v = nd.lca(u, w);
    // score.go:79-88: the loop over v's neighbours overwrites the entry every time (LCA(v, w) == v
    // always holds), so wSub ends up as v's LAST neighbour = its last child.
    if (nd.children[v].empty()) throw CamusError(Err::Internal, "could not map reticulations to nodes");
    nd.labels.push_back(kv.first);
    nd.u.push_back(u);
    nd.w.push_back(w);
    nd.v.push_back(v);
    nd.wsub.push_back(nd.children[v].back());
  }
  nd.tree = std::move(t);
  return nd;
}

// One record per (reticulation r, network tip t), everything QuartetScore reads about t
// (quartettotals.go:107-157): words {lca(w,t), lca(u,t), depth(lca(w,t)) | depth(lca(u,t)) << 16,
// flags}, flags bit 0 = t under w, 1 = under v, 2 = under wSub, 3 = under u.
inline std::vector<uint32_t> reticulation_records(const NetworkData& nd) {
  const int R = (int)nd.labels.size(), T = nd.n_tips;
  std::vector<uint32_t> rec((size_t)R * T * 4);
  for (int r = 0; r < R; ++r)
    for (int t = 0; t < T; ++t) {
      int leaf = nd.tip_node[t];
      int lw = nd.lca(nd.w[r], leaf), lu = nd.lca(nd.u[r], leaf);
      uint32_t f = (nd.in_leafset(nd.w[r], t) ? 1u : 0u) | (nd.in_leafset(nd.v[r], t) ? 2u : 0u) |
                   (nd.in_leafset(nd.wsub[r], t) ? 4u : 0u) | (nd.in_leafset(nd.u[r], t) ? 8u : 0u);
      uint32_t* o = &rec[((size_t)r * T + t) * 4];
      o[0] = (uint32_t)lw;
      o[1] = (uint32_t)lu;
      o[2] = (uint32_t)nd.depth[lw] | ((uint32_t)nd.depth[lu] << 16);
      o[3] = f;
    }
  return rec;
}

struct RetGeneFlat {
  std::vector<int64_t> node_off{0};
  std::vector<int32_t> parent, taxon;  // taxon = NETWORK tip index at leaves, -1 inside
};

// score.go:32-42 per gene tree: UpdateTipIndex (ErrMulTree), MapIDsFromConstTree against the network's
// tips (ErrTipNameMismatch). UnRoot is applied by the device kernel's multiplicity rule.
inline RetGeneFlat flatten_gene_trees_for_network(std::vector<Tree>& gene_trees, const NetworkData& nd) {
  std::unordered_map<std::string, int> name2idx;
  for (int i = 0; i < nd.n_tips; ++i) name2idx[nd.tip_names[i]] = i;
  RetGeneFlat g;
  for (size_t gi = 0; gi < gene_trees.size(); ++gi) {
    Tree& gt = gene_trees[gi];
    if (!gt.update_tip_index()) throw CamusError(Err::MulTree, "gene tree contains duplicate labels");
    std::vector<int> pre = gt.preorder();
    std::vector<int> local(gt.nodes.size(), -1);
    for (size_t i = 0; i < pre.size(); ++i) local[pre[i]] = (int)i;
    for (int x : pre) {
      const Node& n = gt.nodes[x];
      g.parent.push_back(n.parent >= 0 ? local[n.parent] : -1);
      if (n.tip()) {
        auto it = name2idx.find(n.name);
        if (it == name2idx.end())
          throw CamusError(Err::TipNameMismatch, "tip name mismatch! no tip named " + n.name + " in the network");
        g.taxon.push_back(it->second);
      } else {
        g.taxon.push_back(-1);
      }
    }
    g.node_off.push_back((int64_t)g.parent.size());
  }
  return g;
}

// strconv.FormatFloat(v, 'f', -1, 64) for the values this path produces (ratios in [0, 1] and NaN)
inline std::string format_ratio(uint64_t supported, uint64_t total) {
  if (total == 0) return "NaN";
  double v = (double)supported / (double)total;
  char buf[64];
  for (int prec = 1; prec <= 17; ++prec) {
    snprintf(buf, sizeof buf, "%.*g", prec, v);
    if (strtod(buf, nullptr) == v) break;
  }
  std::string s(buf);
  // %g may switch to exponent form for small values; 'f' never does
  if (s.find('e') != std::string::npos) {
    int exp10 = atoi(s.c_str() + s.find('e') + 1);
    std::string mant = s.substr(0, s.find('e'));
    std::string digits;
    for (char c : mant)
      if (c != '.') digits += c;
    if (exp10 < 0) s = "0." + std::string((size_t)(-exp10 - 1), '0') + digits;
  }
  return s;
}

}  // namespace camus
