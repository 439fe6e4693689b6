// Gene-tree ingest on the GPU: newick text -> the flat trees of
// camus_b200_add_gene_trees, directly in HBM.
//
// Replaces, for newick files in the plain dialect, the parse loop of prep.readGeneTreesFile
// (internal/prep/io.go:133-153, gotree's newick parser), the per-tree UpdateTipIndex /
// CollapseLowSupport of prep.processQuartets (internal/prep/preprocess.go:87-100), the label mapping
// of graphs.MapIDsFromConstTree (internal/graphs/quartet.go:131-149) and the flattening the caller
// would do in front of the C-ABI. One warp per gene tree, lane 0 walking the text (a tree is ~8 bytes per node
// of text and the scan is a strict left-to-right automaton), two passes:
//   pass 1  matches parentheses (stack in a scratch array), marks the internal nodes that
//           CollapseLowSupport(min_support, true) contracts (numeric label after ')' below the
//           threshold, not the root) and counts the nodes that remain;
//   pass 2  assigns preorder ids (= text order of '(' and leaf labels), writes parent[] and taxon[]
//           (label -> constraint tip index through a hash table + byte compare) and the per-tree
//           facts the host used to compute: leaves, depth, all-binary, rooted, duplicate labels.
// Dialect: what camus_b200/host/tree.hpp's NewickParser accepts MINUS quoted labels and minus support
// values that are not plain decimals of <= 15 significant digits (for those double(M) / 10^f is the
// correctly rounded value strtod returns; anything else would need a full strtod). A tree outside the
// dialect, with a label the constraint does not have or with a duplicate label gets a status code; the
// caller then adds nothing and runs its own parser, which reports the reference's error.
#pragma once
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define NWK_HD __host__ __device__ __forceinline__
#else
#define NWK_HD inline  // tests/ingest_probe.cpp compiles the per-tree automaton for the host
#endif

#include <cstdint>

namespace camus_dev {

enum : int {
  kNwkOk = 0,
  kNwkSyntax = 1,       // outside the dialect or malformed
  kNwkQuoted = 2,       // single-quoted label
  kNwkNumber = 3,       // support value that is not a plain decimal
  kNwkUnknownLabel = 4, // graphs.ErrTipNameMismatch
  kNwkDuplicate = 5,    // prep.ErrMulTree
  kNwkTooDeep = 6,
};

struct LabelTable {
  const unsigned long long* hash;  // [cap], 0 = empty slot (hashes are forced non-zero)
  const int* tip;                  // [cap]
  const long long* label_off;      // [n_taxa + 1] into pool
  const char* pool;
  unsigned int mask;               // cap - 1
};

struct NwkTreeInfo {  // per tree, read back by the host
  int status;
  int n_nodes;     // nodes that remain after the collapse
  int n_leaves;
  int max_depth;   // edges from the root
  int binary;      // every internal node has 2 children (the root may have 3)
  int root_children;
  int err_pos;     // byte offset of the first problem (diagnostics)
  int pad;
};

NWK_HD bool nwk_ws(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }
NWK_HD bool nwk_stop(char c) { return c == ',' || c == '(' || c == ')' || c == ':' || c == ';' || c == '['; }

// whitespace and [comments] (NewickParser::skip_ws_comments); false = unterminated comment
NWK_HD bool nwk_skip(const char* __restrict__ t, long long& p, long long e) {
  while (p < e) {
    const char c = t[p];
    if (nwk_ws(c)) {
      ++p;
    } else if (c == '[') {
      while (p < e && t[p] != ']') ++p;
      if (p >= e) return false;
      ++p;
    } else {
      break;
    }
  }
  return true;
}
// unquoted label (NewickParser::read_label): [ls, le) with trailing whitespace trimmed
NWK_HD int nwk_label(const char* __restrict__ t, long long& p, long long e, long long& ls, long long& le) {
  if (!nwk_skip(t, p, e)) return kNwkSyntax;
  if (p < e && t[p] == '\'') return kNwkQuoted;
  ls = p;
  while (p < e && !nwk_stop(t[p])) ++p;
  le = p;
  while (le > ls && nwk_ws(t[le - 1])) --le;
  return kNwkOk;
}
// ':' length (NewickParser::read_length)
NWK_HD int nwk_length(const char* __restrict__ t, long long& p, long long e) {
  if (!nwk_skip(t, p, e)) return kNwkSyntax;
  if (p < e && t[p] == ':') {
    ++p;
    if (!nwk_skip(t, p, e)) return kNwkSyntax;
    while (p < e) {
      const char c = t[p];
      if (c == ',' || c == '(' || c == ')' || c == ';' || c == '[') break;
      ++p;
    }
    if (!nwk_skip(t, p, e)) return kNwkSyntax;
  }
  return kNwkOk;
}
// Label after ')': 0 = a name (strtod would reject it), 1 = plain decimal (value set), 2 = may be a
// number strtod accepts but outside the exact fast path
NWK_HD int nwk_number(const char* __restrict__ t, long long ls, long long le, double& v) {
  if (le <= ls) return 0;
  unsigned long long m = 0;
  int digits = 0, sig = 0, frac = 0;
  bool dot = false, plain = true;
  for (long long p = ls; p < le; ++p) {
    const char c = t[p];
    if (c >= '0' && c <= '9') {
      ++digits;
      if (sig > 0 || c != '0') ++sig;
      if (sig <= 15) m = m * 10 + (unsigned)(c - '0');
      if (dot) ++frac;
    } else if (c == '.' && !dot) {
      dot = true;
    } else {
      plain = false;
      break;
    }
  }
  if (plain && digits > 0 && sig <= 15 && frac <= 22) {
    double den = 1.0;
    for (int i = 0; i < frac; ++i) den *= 10.0;  // 10^f, f <= 22: exact
#if defined(__CUDA_ARCH__)
    v = __ddiv_rn((double)m, den);               // one correctly rounded division = strtod's result
#else
    v = (double)m / den;
#endif
    return 1;
  }
  // a name unless strtod could still accept the whole label: sign / digit / dot first, or inf... / nan... (any case)
  auto lower = [&](long long i) -> char {
    const char c = ls + i < le ? t[ls + i] : '\0';
    return (c >= 'A' && c <= 'Z') ? (char)(c + 32) : c;
  };
  const char c0 = lower(0);
  const bool maybe = (c0 >= '0' && c0 <= '9') || c0 == '+' || c0 == '-' || c0 == '.' || (c0 == 'i' && lower(1) == 'n' && lower(2) == 'f') ||
                     (c0 == 'n' && lower(1) == 'a' && lower(2) == 'n');
  return maybe ? 2 : 0;
}

NWK_HD unsigned long long nwk_hash(const char* __restrict__ t, long long ls, long long le) {
  unsigned long long h = 14695981039346656037ull;  // FNV-1a
  for (long long p = ls; p < le; ++p) {
    h ^= (unsigned char)t[p];
    h *= 1099511628211ull;
  }
  return h ? h : 1ull;
}
NWK_HD int nwk_lookup(const LabelTable& lt, const char* __restrict__ t, long long ls, long long le) {
  const unsigned long long h = nwk_hash(t, ls, le);
  for (unsigned int s = (unsigned int)h & lt.mask;; s = (s + 1) & lt.mask) {
    const unsigned long long hs = lt.hash[s];
    if (hs == 0) return -1;
    if (hs != h) continue;
    const int tip = lt.tip[s];
    const long long b = lt.label_off[tip], n = lt.label_off[tip + 1] - b;
    if (n != le - ls) continue;
    bool same = true;
    for (long long i = 0; i < n && same; ++i) same = lt.pool[b + i] == t[ls + i];
    if (same) return tip;
  }
}

// Pass 1: one thread per tree (= non-blank line [begin, end) of the file).
NWK_HD void nwk_pass1_tree(const char* __restrict__ text, long long b, long long e, double min_support, int* stack, unsigned char* collapsed,
                           NwkTreeInfo* out) {
  long long p = b;
  NwkTreeInfo ti{};
  int depth = 0, nodes = 0, leaves = 0, st = kNwkOk;
  int* stk = stack + b;  // at most one entry per '(' of this line
  auto fail = [&](int code) {
    st = code;
    ti.err_pos = (int)(p - b);
  };
  if (!nwk_skip(text, p, e) || p >= e || text[p] != '(') fail(kNwkSyntax);
  bool want_subtree = true;  // the automaton expects a subtree (after '(' or ',') or what follows a node
  while (st == kNwkOk) {
    if (want_subtree) {
      if (!nwk_skip(text, p, e)) { fail(kNwkSyntax); break; }
      if (p < e && text[p] == '(') {
        stk[depth++] = (int)(p - b);
        collapsed[p] = 0;
        ++nodes;
        ++p;
        continue;
      }
      long long ls, le;
      int rc = nwk_label(text, p, e, ls, le);
      if (rc) { fail(rc); break; }
      ++nodes;
      ++leaves;
      if ((rc = nwk_length(text, p, e))) { fail(rc); break; }
      want_subtree = false;
      continue;
    }
    if (!nwk_skip(text, p, e) || p >= e) { fail(kNwkSyntax); break; }
    const char c = text[p];
    if (c == ',') {
      if (depth == 0) { fail(kNwkSyntax); break; }
      ++p;
      want_subtree = true;
    } else if (c == ')') {
      if (depth == 0) { fail(kNwkSyntax); break; }
      ++p;
      const int open = stk[--depth];
      long long ls, le;
      int rc = nwk_label(text, p, e, ls, le);
      if (rc) { fail(rc); break; }
      if (min_support != 0) {  // CollapseLowSupport runs only with -s != 0 (preprocess.go:98-100)
        double v = 0;
        const int kind = nwk_number(text, ls, le, v);
        if (kind == 2) { fail(kNwkNumber); break; }
        if (kind == 1 && depth > 0 && v < min_support) {  // depth > 0: not the root
          collapsed[b + open] = 1;
          --nodes;
        }
      }
      if ((rc = nwk_length(text, p, e))) { fail(rc); break; }
    } else if (c == ';') {
      if (depth != 0) { fail(kNwkSyntax); break; }
      ++p;
      if (!nwk_skip(text, p, e) || p != e) fail(kNwkSyntax);  // one tree per line
      break;
    } else {
      fail(kNwkSyntax);
      break;
    }
  }
  ti.status = st;
  ti.n_nodes = st == kNwkOk ? nodes : 0;
  ti.n_leaves = leaves;
  *out = ti;
}
#if defined(__CUDACC__)
__global__ void k_nwk_pass1(int n_trees, const char* __restrict__ text, const long long* __restrict__ line_begin,
                            const long long* __restrict__ line_end, double min_support, int* stack, unsigned char* collapsed,
                            NwkTreeInfo* info) {
  // One tree per WARP, lane 0 only: 32 lanes walking 32 different texts diverge at every byte and touch 32 cache
  // lines per load (8 ms for 1000 trees, ncu: issue 10 %); a lone lane streams its text through L1.
  const int g = blockIdx.x;
  if (g >= n_trees || threadIdx.x != 0) return;
  nwk_pass1_tree(text, line_begin[g], line_end[g], min_support, stack, collapsed, info + g);
}
#endif

#if defined(__CUDACC__)
// exclusive scan of the node counts (one block; G is at most a few hundred thousand)
__global__ void k_nwk_offsets(int n_trees, const NwkTreeInfo* __restrict__ info, long long base, long long* node_off /*[n_trees+1]*/) {
  __shared__ long long part[1024];
  const int tid = threadIdx.x, per = (n_trees + blockDim.x - 1) / blockDim.x;
  const int lo = min(n_trees, tid * per), hi = min(n_trees, lo + per);
  long long s = 0;
  for (int g = lo; g < hi; ++g) s += info[g].n_nodes;
  part[tid] = s;
  __syncthreads();
  if (tid == 0) {
    long long run = base;
    for (int i = 0; i < (int)blockDim.x; ++i) {
      const long long v = part[i];
      part[i] = run;
      run += v;
    }
    node_off[n_trees] = run;
  }
  __syncthreads();
  long long run = part[tid];
  for (int g = lo; g < hi; ++g) {
    node_off[g] = run;
    run += info[g].n_nodes;
  }
}

#endif

// Pass 2: emit the flat tree. `seen`: ceil(n_taxa / 32) words per tree, zeroed.
// o = first slot of this tree in parent[] / taxon[] / n_child[], sn = its n_taxa bits of `seen`
NWK_HD void nwk_pass2_tree(const char* __restrict__ text, long long b, long long e, const unsigned char* __restrict__ collapsed, int* stack,
                           const LabelTable& lt, long long o, int* parent, int* taxon, int* n_child, unsigned int* sn, NwkTreeInfo* io) {
  NwkTreeInfo ti = *io;
  if (ti.status != kNwkOk) return;
  int* stk = stack + b;  // entries: local id of the nearest KEPT node of that open parenthesis
  long long p = b;
  int depth = 0, kept_depth = -1, next = 0, max_depth = 0, st = kNwkOk;
  nwk_skip(text, p, e);
  bool want_subtree = true;
  while (st == kNwkOk) {
    if (want_subtree) {
      nwk_skip(text, p, e);
      const int par = depth > 0 ? stk[depth - 1] : -1;
      if (text[p] == '(') {
        if (collapsed[p]) {
          stk[depth++] = par;  // its children hang off the nearest kept ancestor
        } else {
          const int id = next++;
          parent[o + id] = par;
          taxon[o + id] = -1;
          n_child[o + id] = 0;
          if (par >= 0) ++n_child[o + par];
          stk[depth++] = id;
          ++kept_depth;
          max_depth = kept_depth > max_depth ? kept_depth : max_depth;
        }
        ++p;
        continue;
      }
      long long ls, le;
      nwk_label(text, p, e, ls, le);
      const int id = next++;
      const int tip = nwk_lookup(lt, text, ls, le);
      parent[o + id] = par;
      taxon[o + id] = tip;
      n_child[o + id] = 0;
      if (par >= 0) ++n_child[o + par];
      max_depth = kept_depth + 1 > max_depth ? kept_depth + 1 : max_depth;
      if (tip < 0) {
        st = kNwkUnknownLabel;
        ti.err_pos = (int)(ls - b);
        break;
      }
      if (sn[tip >> 5] & (1u << (tip & 31))) {
        st = kNwkDuplicate;
        ti.err_pos = (int)(ls - b);
        break;
      }
      sn[tip >> 5] |= 1u << (tip & 31);
      nwk_length(text, p, e);
      want_subtree = false;
      continue;
    }
    nwk_skip(text, p, e);
    const char c = text[p];
    if (c == ',') {
      ++p;
      want_subtree = true;
    } else if (c == ')') {
      ++p;
      --depth;
      // was the node just closed kept? collapsed ones pushed their parent's id: compare with the entry below
      long long ls, le;
      nwk_label(text, p, e, ls, le);
      nwk_length(text, p, e);
      // kept nodes were counted in kept_depth when opened; recover the flag from the scratch of pass 1
      // (the open position is not on this stack any more): a kept node's id differs from its parent's
      const int me = stk[depth], par = depth > 0 ? stk[depth - 1] : -1;
      if (me != par) --kept_depth;
    } else {  // ';' (pass 1 checked the rest)
      break;
    }
  }
  ti.status = st;
  ti.max_depth = max_depth;
  if (st == kNwkOk) {
    int binary = 1;
    for (int i = 0; i < ti.n_nodes; ++i) {
      const int nc = n_child[o + i];
      if (nc != 0 && nc != 2 && !(i == 0 && nc == 3)) binary = 0;
      if ((nc == 0) != (taxon[o + i] >= 0)) st = kNwkSyntax;  // "()" style empty groups: a leaf without a label cannot get here, an internal node always has children
    }
    ti.binary = binary;
    ti.root_children = n_child[o];
    ti.status = st;
  }
  *io = ti;
}
#if defined(__CUDACC__)
__global__ void k_nwk_pass2(int n_trees, const char* __restrict__ text, const long long* __restrict__ line_begin,
                            const long long* __restrict__ line_end, const unsigned char* __restrict__ collapsed, int* stack,
                            LabelTable lt, int n_taxa, const long long* __restrict__ node_off, long long out_base, int* parent, int* taxon,
                            int* n_child, unsigned int* seen, NwkTreeInfo* info) {
  const int g = blockIdx.x;  // one tree per warp, lane 0 only (see k_nwk_pass1)
  if (g >= n_trees || threadIdx.x != 0) return;
  nwk_pass2_tree(text, line_begin[g], line_end[g], collapsed, stack, lt, node_off[g] - out_base, parent, taxon, n_child,
                 seen + (size_t)g * ((n_taxa + 31) / 32), info + g);
}
#endif

}  // namespace camus_dev
