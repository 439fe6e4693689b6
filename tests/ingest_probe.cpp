// Test helper: the per-tree automaton of camus_b200/csrc/kernels_ingest.cuh (the code k_nwk_pass1 / k_nwk_pass2 run on the
// GPU, one thread per gene tree) compiled for the HOST with g++, so that the `-m "not gpu"` suite can check the newick ->
// flat-tree logic against the host mirror's parser without a GPU. Not a CPU fallback: nothing in the product calls it.
#include <cstring>
#include <string>
#include <vector>

#include "../camus_b200/csrc/kernels_ingest.cuh"

using namespace camus_dev;

extern "C" {
// text: the gene-tree file; labels / label_off: constraint tip names in tip-index order. Outputs sized by the caller:
// node_off[n_lines + 1], parent / taxon [cap]. Returns the number of trees, or -(1-based tree index) of the first bad tree
// with *status = its code.
long long probe_ingest(const char* text, long long len, double min_support, int n_taxa, const char* labels, const long long* label_off,
                       long long* node_off, int* parent, int* taxon, long long cap, int* status, int* missing) {
  unsigned cap_h = 16;
  while (cap_h < 2u * (unsigned)n_taxa) cap_h <<= 1;
  std::vector<unsigned long long> hs(cap_h, 0);
  std::vector<int> tips(cap_h, -1);
  for (int t = 0; t < n_taxa; ++t) {
    unsigned long long h = nwk_hash(labels, label_off[t], label_off[t + 1]);
    unsigned s = (unsigned)h & (cap_h - 1);
    while (hs[s]) s = (s + 1) & (cap_h - 1);
    hs[s] = h;
    tips[s] = t;
  }
  LabelTable lt{hs.data(), tips.data(), label_off, labels, cap_h - 1};
  std::vector<long long> lb, le;
  for (long long start = 0; start < len;) {
    const char* nl = (const char*)memchr(text + start, '\n', (size_t)(len - start));
    long long end = nl ? (long long)(nl - text) : len, b = start, e = end;
    while (b < e && (unsigned char)text[b] <= ' ') ++b;
    while (e > b && (unsigned char)text[e - 1] <= ' ') --e;
    if (e > b) {
      lb.push_back(b);
      le.push_back(e);
    }
    start = end + 1;
  }
  const int G = (int)lb.size();
  std::vector<int> stack((size_t)len + 1);
  std::vector<unsigned char> coll((size_t)len + 1, 0);
  std::vector<NwkTreeInfo> info(G);
  for (int g = 0; g < G; ++g) nwk_pass1_tree(text, lb[g], le[g], min_support, stack.data(), coll.data(), &info[g]);
  node_off[0] = 0;
  for (int g = 0; g < G; ++g) node_off[g + 1] = node_off[g] + info[g].n_nodes;
  if (node_off[G] > cap) return -1000000000LL;
  std::vector<int> n_child((size_t)node_off[G] + 1);
  std::vector<unsigned> seen((size_t)(n_taxa + 31) / 32);
  *missing = 0;
  for (int g = 0; g < G; ++g) {
    std::fill(seen.begin(), seen.end(), 0u);
    nwk_pass2_tree(text, lb[g], le[g], coll.data(), stack.data(), lt, node_off[g], parent, taxon, n_child.data(), seen.data(), &info[g]);
    if (info[g].status != kNwkOk) {
      *status = info[g].status;
      return -(long long)(g + 1);
    }
    if (info[g].n_leaves != n_taxa) *missing = 1;
  }
  *status = 0;
  return G;
}
}
