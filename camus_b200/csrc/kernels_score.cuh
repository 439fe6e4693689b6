// Stage 3-5 kernels: filter + constraint removal (K3), edge-score scatter (K4), finalisation
// scans + ShouldCalcEdge mask (K5), penalties, and export of the surviving quartets.
#pragma once
#include <cuda_runtime.h>

#include "layout.cuh"
#include "ptx.cuh"

namespace camus_dev {

// Constraint tree on the device. Node ids are INTERNAL preorder ids (root 0, left child = id+1,
// subtree(x) = [x, x+sz[x])); ref_id maps back to the caller's ids. Taxa are DFS ranks.
struct ConstraintDev {
  int n, n_taxa;
  const int32_t* parent;
  const int32_t* depth;
  const int32_t* sz;
  const int32_t* rc;          // right child, -1 at tips
  const int32_t* nleaves;     // leaves below node
  const int32_t* firstleaf;   // DFS rank of the first leaf below node
  const int32_t* leafnode;    // [n_taxa] node id of the leaf with that DFS rank
  const int32_t* tip_of_rank; // [n_taxa] gotree tip index of that leaf
  const int32_t* ref_id;      // [n]
  const uint32_t* lcaD;       // [n_taxa][n_taxa] (depth << 16 | node) of the LCA of two leaves
  const uint8_t* seg_uniform; // [n_seg][n_seg] 1 if every leaf pair of the two 8-leaf segments has the same LCA
  int n_seg;
};

struct CellCoords {
  uint32_t a, b, c, d;
  bool valid;
};

// 256 consecutive cells of one box per CTA; thread 0 unranks the box.
__device__ __forceinline__ CellCoords decode_cell(const BoxMap& bm, int n_taxa, uint32_t* s_seg) {
  if (threadIdx.x == 0) {
    uint32_t sa, sb, sc, sd;
    box_unrank(bm.global(blockIdx.x >> 4), sa, sb, sc, sd);
    s_seg[0] = sa;
    s_seg[1] = sb;
    s_seg[2] = sc;
    s_seg[3] = sd;
  }
  __syncthreads();
  uint32_t local = ((blockIdx.x & 15u) << 8) | threadIdx.x;
  CellCoords cc;
  cc.a = s_seg[0] * 8 + (local & 7);
  cc.b = s_seg[1] * 8 + ((local >> 3) & 7);
  cc.c = s_seg[2] * 8 + ((local >> 6) & 7);
  cc.d = s_seg[3] * 8 + (local >> 9);
  cc.valid = cc.a < cc.b && cc.b < cc.c && cc.c < cc.d && cc.d < (uint32_t)n_taxa;
  return cc;
}

// Topology the constraint tree displays on DFS-ordered a<b<c<d: 0 = ab|cd, 2 = ad|bc (never ac|bd).
// g1,g2,g3 = packed LCAs of (a,b), (b,c), (c,d); the shallowest gap is the root split.
__device__ __forceinline__ int constraint_topology(uint32_t g1, uint32_t g2, uint32_t g3) {
  if (g2 < g1 && g2 < g3) return 0;
  if (g1 < g3) return g2 < g3 ? 0 : 2;
  return g1 < g2 ? 2 : 0;
}

// partner of taxon position i (0..3 = a,b,c,d) in topology t (0 ab|cd, 1 ac|bd, 2 ad|bc)
__device__ __forceinline__ int partner_of(int t, int i) { return t == 0 ? (i ^ 1) : (t == 1 ? (i ^ 2) : (3 - i)); }

// seg_uniform[A][B] (A < B): all 64 leaf pairs of the segments A and B (8 consecutive DFS ranks
// each) have the same LCA, and both segments are full. A box whose three consecutive segment pairs
// are uniform ("regular box") has one junction / one m-node / one component per (bottom taxon,
// topology), which the fused count kernel exploits. One warp per segment pair.
__global__ void k_segment_pairs(int n_taxa, int n_seg, const uint32_t* __restrict__ lcaD, uint8_t* __restrict__ out) {
  const int pair = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (pair >= n_seg * n_seg) return;
  const int A = pair / n_seg, B = pair % n_seg;
  bool ok = A < B && B * 8 + 8 <= n_taxa;
  if (ok) {
    const uint32_t ref = lcaD[(size_t)(A * 8) * n_taxa + B * 8];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int e = lane + 32 * k;
      ok = ok && lcaD[(size_t)(A * 8 + (e >> 3)) * n_taxa + B * 8 + (e & 7)] == ref;
    }
  }
  ok = __all_sync(0xFFFFFFFFu, ok);
  if (lane == 0) out[pair] = ok ? 1 : 0;
}

// prep.filterQuartets + Threshold.Keep for one 4-set (internal/prep/quartetfilter.go:67-96).
// c[] = counts of ab|cd, ac|bd, ad|bc (DFS order); ti[] = gotree tip indices of a,b,c,d.
__device__ __forceinline__ void apply_filter(uint32_t c[3], const int ti[4], int qmode, double threshold) {
  // pos[i] = rank of taxon i among the four by tip index; the taxon with pos 0 is "taxon 0" of the
  // reference's packing. rho[t] = position (0,1,2) of topology t in the reference's AllQuartets()
  // order = pos of taxon 0's partner, minus one. All loops unroll to selects (no local arrays).
  int pos[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int r = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) r += (ti[j] < ti[i]);
    pos[i] = r;
  }
  int rho[3];
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    int r = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) r += (pos[i] == 0) ? (pos[partner_of(t, i)] - 1) : 0;
    rho[t] = r;
  }
  // stable sort of the three topologies by count (insertion sort on 3 elements in Go): the keys
  // (count, rho) are distinct, so min / middle / max are well defined; no indexed local arrays.
  const unsigned long long k0 = ((unsigned long long)c[0] << 2) | (unsigned)rho[0];
  const unsigned long long k1 = ((unsigned long long)c[1] << 2) | (unsigned)rho[1];
  const unsigned long long k2 = ((unsigned long long)c[2] << 2) | (unsigned)rho[2];
  const unsigned long long kmin = min(k0, min(k1, k2)), kmax = max(k0, max(k1, k2));
  const unsigned long long kmid = k0 ^ k1 ^ k2 ^ kmin ^ kmax;
  const uint32_t lo = (uint32_t)(kmin >> 2), mid = (uint32_t)(kmid >> 2);
  const uint32_t s = lo + mid;
  const bool keep = __double2uint_rz(threshold * (double)s) < mid - lo;  // Threshold.Keep
  const unsigned long long cut = !keep ? kmid : (qmode == 2 ? kmin : 0ull);  // delete keys <= cut
  if (!keep || qmode == 2) {
    if (k0 <= cut) c[0] = 0;
    if (k1 <= cut) c[1] = 0;
    if (k2 <= cut) c[2] = 0;
  }
}

// Threshold.Keep's uint32(float64(t) * float64(sum)) for every possible sum of two counts (at most
// the number of gene trees): one table lookup per cell instead of fp64 arithmetic.
__global__ void k_threshold_table(uint32_t* __restrict__ tab, int n, double threshold) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) tab[i] = __double2uint_rz(threshold * (double)i);
}

// Same result as apply_filter, without the tip-index ranks on the common path: the reference's
// tie order only matters when the two LARGEST counts are equal and the cell fails Keep (then one
// of the two survives); a tie between the two smallest always fails Keep and deletes both.
__device__ __forceinline__ void filter_cell(uint32_t c[3], const uint32_t x[4], int qmode, double threshold,
                                            const uint32_t* __restrict__ thr_tab, const int32_t* __restrict__ tip_of_rank) {
  const uint32_t lo = min(c[0], min(c[1], c[2])), hi = max(c[0], max(c[1], c[2]));
  const uint32_t mid = c[0] + c[1] + c[2] - lo - hi;
  const bool keep = __ldg(&thr_tab[lo + mid]) < mid - lo;
  if (keep) {
    if (qmode == 2) {  // mid > lo here, so the minimum is unique
      if (c[0] == lo) c[0] = 0;
      else if (c[1] == lo) c[1] = 0;
      else c[2] = 0;
    }
  } else if (mid != hi) {
    if (c[0] != hi) c[0] = 0;
    if (c[1] != hi) c[1] = 0;
    if (c[2] != hi) c[2] = 0;
  } else {
    const int ti[4] = {tip_of_rank[x[0]], tip_of_rank[x[1]], tip_of_rank[x[2]], tip_of_rank[x[3]]};
    apply_filter(c, ti, qmode, threshold);
  }
}

// ---------------------------------------------------------------------------------------------
// K3: prep.filterQuartets (internal/prep/quartetfilter.go:67-96) applied once per 4-set, then the
// constraint-quartet removal (internal/prep/preprocess.go:51-57), then the two scalars
// len(qCounts) and sum of counts (internal/graphs/treedata.go:297-307). In place on the cells.
// Ties in the sort keep the reference's topology order 0b1100, 0b1010, 0b0110, which is defined
// on taxa sorted by TIP INDEX (quartet.go:169-177), hence the rank computation below.
// ---------------------------------------------------------------------------------------------
struct FilterParams {
  uint4* cells;
  BoxMap bm;
  int qmode;
  double threshold;
  const uint32_t* thr_tab;     // k_threshold_table
  unsigned long long* totals;  // [0] n_survivors, [1] sum_counts
};

__global__ void __launch_bounds__(256) k_filter(FilterParams prm, ConstraintDev ct) {
  __shared__ uint32_t s_seg[4];
  __shared__ unsigned long long s_red[2][8];
  CellCoords cc = decode_cell(prm.bm, ct.n_taxa, s_seg);
  size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
  uint4 v = prm.cells[idx];
  uint32_t c[3] = {v.x, v.y, v.z};
  unsigned long long ns = 0, sum = 0;
  if (!cc.valid) {
    if (c[0] | c[1] | c[2]) prm.cells[idx] = make_uint4(0, 0, 0, 0);
  } else {
    bool changed = false;
    if (prm.qmode != 0 && (c[0] | c[1] | c[2])) {
      const uint32_t x[4] = {cc.a, cc.b, cc.c, cc.d};
      filter_cell(c, x, prm.qmode, prm.threshold, prm.thr_tab, ct.tip_of_rank);
      changed = true;
    }
    if (c[0] | c[1] | c[2]) {
      uint32_t g1 = ct.lcaD[(size_t)cc.a * ct.n_taxa + cc.b];
      uint32_t g2 = ct.lcaD[(size_t)cc.b * ct.n_taxa + cc.c];
      uint32_t g3 = ct.lcaD[(size_t)cc.c * ct.n_taxa + cc.d];
      if (constraint_topology(g1, g2, g3) == 0) {
        changed |= c[0] != 0;
        c[0] = 0;
      } else {
        changed |= c[2] != 0;
        c[2] = 0;
      }
    }
    if (changed) prm.cells[idx] = make_uint4(c[0], c[1], c[2], 0);
    ns = (c[0] != 0) + (c[1] != 0) + (c[2] != 0);
    sum = (unsigned long long)c[0] + c[1] + c[2];
  }
  // block reduction -> two 64-bit atomics per CTA
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    ns += __shfl_xor_sync(0xFFFFFFFFu, ns, off);
    sum += __shfl_xor_sync(0xFFFFFFFFu, sum, off);
  }
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    s_red[0][warp] = ns;
    s_red[1][warp] = sum;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long a = 0, b = 0;
    for (int i = 0; i < 8; ++i) {
      a += s_red[0][i];
      b += s_red[1][i];
    }
    if (a) red_add_u64(&prm.totals[0], a);
    if (b) red_add_u64(&prm.totals[1], b);
  }
}

// ---------------------------------------------------------------------------------------------
// K4: edge-score scatter. Replaces graphs.mapQuartetsToVertices + the per-(u,w) QuartetScore loop
// (internal/graphs/treedata.go:197-222, internal/score/quartettotals.go:78-157) by the
// per-quartet form of SURVEY.md Appendix A: a surviving quartet q with weight c and bottom taxon
// beta adds c to every (u, w) with
//     w on the path from leaf beta up to, excluding, m(beta) = lowest ancestor shared with
//       another taxon of q,
//     u in the component, holding beta's partner x in q, of T minus the junction J of the other
//       three taxa: LEFT(J) = subtree(J+1), RIGHT(J) = subtree(rc[J]), OUT(J) = all but the two.
// Accumulated as  Z[w-point][J][component] += c at w-point = leaf(beta), -= c at w-point = m,
// i.e. two 64-bit reductions per (quartet, bottom); K5 expands Z into the n x n table.
// ---------------------------------------------------------------------------------------------
struct ScoreParams {
  const uint4* cells;
  BoxMap bm;
  int as_set;
  unsigned long long* Z;  // [n][n][3] two's-complement int64
};

// One cell's contribution to Z (see the K4 comment above). x[] = DFS ranks a<b<c<d, g1..g3 = packed
// LCAs of (a,b), (b,c), (c,d), cnt[] = surviving counts per topology.
// acc(index, value) adds a signed value to Z[index]: straight to HBM (GlobalAcc) or through the
// per-CTA shared-memory aggregator of the fused kernel (SmemAcc, kernels_count.cuh).
struct GlobalAcc {
  unsigned long long* Z;
  __device__ __forceinline__ void operator()(size_t idx, long long v) const { red_add_u64(&Z[idx], (unsigned long long)v); }
};

template <class Acc>
__device__ __forceinline__ void score_cell(const uint32_t cnt[3], const uint32_t x[4], uint32_t g1, uint32_t g2,
                                           uint32_t g3, const ConstraintDev& ct, Acc& acc, int as_set) {
  const size_t n = (size_t)ct.n;
#pragma unroll
  for (int be = 0; be < 4; ++be) {
    uint32_t h1, h2, m;
    if (be == 0) {
      h1 = g2; h2 = g3; m = g1;
    } else if (be == 1) {
      h1 = min(g1, g2); h2 = g3; m = max(g1, g2);
    } else if (be == 2) {
      h1 = g1; h2 = min(g2, g3); m = max(g2, g3);
    } else {
      h1 = g1; h2 = g2; m = g3;
    }
    const bool first = h1 > h2;  // junction is the LCA of the first two of the other three
    const uint32_t J = (first ? h1 : h2) & 0xFFFFu;
    const size_t row_leaf = ((size_t)ct.leafnode[x[be]] * n + J) * 3;
    const size_t row_m = ((size_t)(m & 0xFFFFu) * n + J) * 3;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      if (cnt[t] == 0) continue;
      int p = partner_of(t, be);
      int pos = p - (p > be ? 1 : 0);                 // position of the partner among the other three
      int comp = first ? pos : (pos + 2) % 3;         // 0 LEFT, 1 RIGHT, 2 OUT
      const long long w = as_set ? 1ll : (long long)cnt[t];
      acc(row_leaf + comp, w);
      acc(row_m + comp, -w);
    }
  }
}

__global__ void __launch_bounds__(256) k_score(ScoreParams prm, ConstraintDev ct) {
  __shared__ uint32_t s_seg[4];
  CellCoords cc = decode_cell(prm.bm, ct.n_taxa, s_seg);
  if (!cc.valid) return;
  uint4 v = prm.cells[(size_t)blockIdx.x * 256 + threadIdx.x];
  if (!(v.x | v.y | v.z)) return;
  const uint32_t cnt[3] = {v.x, v.y, v.z};
  const uint32_t x[4] = {cc.a, cc.b, cc.c, cc.d};
  const uint32_t g1 = ct.lcaD[(size_t)cc.a * ct.n_taxa + cc.b];
  const uint32_t g2 = ct.lcaD[(size_t)cc.b * ct.n_taxa + cc.c];
  const uint32_t g3 = ct.lcaD[(size_t)cc.c * ct.n_taxa + cc.d];
  GlobalAcc acc{prm.Z};
  score_cell(cnt, x, g1, g2, g3, ct, acc, prm.as_set);
}

// ---------------------------------------------------------------------------------------------
// K5a: expand Z[w][J][comp] into a difference table over u: E[w][u], u in [0, n].
//   LEFT  : +v at J+1,   -v at rc[J]
//   RIGHT : +v at rc[J], -v at J+sz[J]
//   OUT   : +v at 0,     -v at J+1,   +v at J+sz[J]
// ---------------------------------------------------------------------------------------------
__global__ void k_expand(const unsigned long long* __restrict__ Z, unsigned long long* __restrict__ E, ConstraintDev ct) {
  int J = blockIdx.x * blockDim.x + threadIdx.x;
  int w = blockIdx.y;
  if (J >= ct.n || ct.rc[J] < 0) return;
  const size_t n = (size_t)ct.n, ld = n + 1;
  const unsigned long long* z = Z + ((size_t)w * n + J) * 3;
  long long vL = (long long)z[0], vR = (long long)z[1], vO = (long long)z[2];
  if (!(vL | vR | vO)) return;
  unsigned long long* row = E + (size_t)w * ld;
  long long at1 = vL - vO, at2 = vR - vL, at3 = vO - vR;
  if (at1) atomicAdd(&row[J + 1], (unsigned long long)at1);
  if (at2) atomicAdd(&row[ct.rc[J]], (unsigned long long)at2);
  if (at3) atomicAdd(&row[J + ct.sz[J]], (unsigned long long)at3);
  if (vO) atomicAdd(&row[0], (unsigned long long)vO);
}

// Sum the privatised copies of Z into copy 0.
__global__ void k_sum_replicas(unsigned long long* __restrict__ Z, size_t zn, int rep) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= zn) return;
  unsigned long long acc = Z[i];
  for (int r = 1; r < rep; ++r) acc += Z[(size_t)r * zn + i];
  Z[i] = acc;
}

// K5b: inclusive prefix sum along u of every row (one CTA per row).
__global__ void __launch_bounds__(256) k_row_scan(unsigned long long* __restrict__ E, int n) {
  __shared__ unsigned long long s_part[256];
  const size_t ld = (size_t)n + 1;
  unsigned long long* row = E + (size_t)blockIdx.x * ld;
  int chunk = (n + 255) / 256;
  int lo = threadIdx.x * chunk, hi = min(n, lo + chunk);
  unsigned long long acc = 0;
  for (int u = lo; u < hi; ++u) acc += row[u];
  s_part[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    for (int i = 0; i < 256; ++i) {
      unsigned long long t = s_part[i];
      s_part[i] = run;
      run += t;
    }
  }
  __syncthreads();
  acc = s_part[threadIdx.x];
  for (int u = lo; u < hi; ++u) {
    acc += row[u];
    row[u] = acc;
  }
}

// K5c: exclusive prefix sum down every column (over w); row n receives the column total, so that
// the subtree sum over w' in [w, w+sz[w]) is E[w+sz[w]][u] - E[w][u].
// One CTA per 32 columns, 32 warps: warp r owns a band of rows (lane = column, 256-byte coalesced rows), sums
// it, the band totals are scanned in shared memory, then every warp rewrites its band. (One thread per column
// walking all n rows -- a chain of n dependent load / store pairs -- took 0.6 ms at n = 1999.)
__global__ void __launch_bounds__(1024) k_col_scan(unsigned long long* __restrict__ E, int n) {
  __shared__ unsigned long long s_band[32][33];
  const int lane = threadIdx.x & 31, r = threadIdx.x >> 5;
  const int u = blockIdx.x * 32 + lane;
  const size_t ld = (size_t)n + 1;
  const int band = (n + 31) / 32, w0 = r * band, w1 = min(n, w0 + band);
  unsigned long long acc = 0;
  if (u < n)
    for (int w = w0; w < w1; ++w) acc += E[(size_t)w * ld + u];
  s_band[r][lane] = acc;
  __syncthreads();
  if (r == 0) {
    unsigned long long run = 0;
    for (int i = 0; i < 32; ++i) {
      const unsigned long long t = s_band[i][lane];
      s_band[i][lane] = run;
      run += t;
    }
    if (u < n) E[(size_t)n * ld + u] = run;
  }
  __syncthreads();
  if (u >= n) return;
  acc = s_band[r][lane];
  for (int w = w0; w < w1; ++w) {
    const unsigned long long t = E[(size_t)w * ld + u];
    E[(size_t)w * ld + u] = acc;
    acc += t;
  }
}

__device__ __forceinline__ int node_lca(const ConstraintDev& ct, int u, int w) {
  if (u <= w && w < u + ct.sz[u]) return u;
  if (w <= u && u < w + ct.sz[w]) return w;
  return (int)(ct.lcaD[(size_t)ct.firstleaf[u] * ct.n_taxa + ct.firstleaf[w]] & 0xFFFFu);
}
// score.ShouldCalcEdge / CycleLength (internal/score/quartettotals.go:63-74)
__device__ __forceinline__ bool should_calc_edge(const ConstraintDev& ct, int u, int w, int& v) {
  v = node_lca(ct, u, w);
  int len = (ct.depth[u] - ct.depth[v]) + (ct.depth[w] - ct.depth[v]) + 1 + (v == u ? 1 : 0);
  bool u_under_w = u > w && u < w + ct.sz[w];
  return !u_under_w && len > 3 && u != 0 && w != 0;
}

// graphs.calcLCAs (internal/graphs/treedata.go:130-166) for every node pair, in the caller's ids.
__global__ void k_node_lca(int32_t* __restrict__ out, ConstraintDev ct) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  int u = blockIdx.y;
  if (w >= ct.n) return;
  out[(size_t)ct.ref_id[u] * ct.n + ct.ref_id[w]] = ct.ref_id[node_lca(ct, u, w)];
}

// K5d: quartetTotals[u][w] in the caller's node ids, 0 where !ShouldCalcEdge.
__global__ void k_mask_out(const unsigned long long* __restrict__ E, unsigned long long* __restrict__ out, ConstraintDev ct) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  int u = blockIdx.y;
  if (w >= ct.n) return;
  const size_t ld = (size_t)ct.n + 1;
  int v;
  unsigned long long val = 0;
  if (should_calc_edge(ct, u, w, v)) val = E[(size_t)(w + ct.sz[w]) * ld + u] - E[(size_t)w * ld + u];
  out[(size_t)ct.ref_id[u] * ct.n + ct.ref_id[w]] = val;
}

// ---------------------------------------------------------------------------------------------
// score.CalculateEdgePenalties (internal/score/penalty.go:11-86): sizes of the subtrees hanging
// off the cycle of (u, w), third elementary symmetric polynomial, uint64 wrap-around.
// ---------------------------------------------------------------------------------------------
__global__ void k_penalties(unsigned long long* __restrict__ out, ConstraintDev ct) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  int u = blockIdx.y;
  if (w >= ct.n) return;
  int v;
  unsigned long long pen = 0;
  if (should_calc_edge(ct, u, w, v)) {
    unsigned long long c1 = 0, c2 = 0, c3 = 0;
    auto add = [&](unsigned long long s) {
      c3 += s * c2;
      c2 += s * c1;
      c1 += s;
    };
    auto walk_up = [&](int start, int end) {  // collectNodesUpPath without its first element
      int cur = start, p;
      do {
        p = ct.parent[cur];
        int sib = ct.rc[p] == cur ? p + 1 : ct.rc[p];
        add((unsigned long long)ct.nleaves[sib]);
        cur = p;
      } while (p != end);
    };
    if (u == v) {
      walk_up(w, ct.parent[v]);
    } else {
      walk_up(w, v);
      add((unsigned long long)ct.nleaves[u]);
      walk_up(u, v);
    }
    add((unsigned long long)(ct.n_taxa - ct.nleaves[v]));
    pen = (unsigned long long)ct.nleaves[w] * c3;
  }
  out[(size_t)ct.ref_id[u] * ct.n + ct.ref_id[w]] = pen;
}

// ---------------------------------------------------------------------------------------------
// Parity probe: raw topology counts of arbitrary 4-taxon sets, read straight from the triple
// bit-planes with scalar loads (an independent, slow reading of the same table k_count streams).
// quads = 4 constraint tip indices per query; out = 3 counts per query in the REFERENCE's
// topology order for taxa sorted by tip index: s0s1|s2s3, s0s2|s1s3, s0s3|s1s2 (quartet.go:23-25).
// ---------------------------------------------------------------------------------------------
__global__ void k_query_cells(int n_query, const int32_t* __restrict__ quads, uint32_t* __restrict__ out,
                              const uint32_t* __restrict__ tri, int n_groups, const int32_t* __restrict__ rank_of_tip,
                              const int32_t* __restrict__ tip_of_rank) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_query) return;
  uint32_t x[4];
  for (int i = 0; i < 4; ++i) x[i] = (uint32_t)rank_of_tip[quads[4 * q + i]];
  for (int i = 0; i < 3; ++i)
    for (int j = i + 1; j < 4; ++j)
      if (x[j] < x[i]) {
        uint32_t t = x[i];
        x[i] = x[j];
        x[j] = t;
      }
  const uint32_t a = x[0], b = x[1], c = x[2], d = x[3];
  auto word = [&](uint32_t p, uint32_t u, uint32_t v, uint32_t w, int g) {  // plane p of triple u < v < w
    return tri[(tile_index(u >> 3, v >> 3, w >> 3) * n_groups + g) * kTileWords + tile_word(p, u & 7, v & 7, w & 7)];
  };
  uint32_t cnt[3] = {0, 0, 0};
  for (int g = 0; g < n_groups; ++g) {
    cnt[0] += __popc((word(0, a, b, c, g) & word(0, a, b, d, g)) | (word(2, a, c, d, g) & word(2, b, c, d, g)));
    cnt[1] += __popc((word(1, a, b, c, g) & word(0, a, c, d, g)) | (word(2, a, b, d, g) & word(1, b, c, d, g)));
    cnt[2] += __popc((word(1, a, b, d, g) & word(1, a, c, d, g)) | (word(2, a, b, c, g) & word(0, b, c, d, g)));
  }
  int ti[4] = {tip_of_rank[a], tip_of_rank[b], tip_of_rank[c], tip_of_rank[d]};
  int pos[4];
  for (int i = 0; i < 4; ++i) {
    int r = 0;
    for (int j = 0; j < 4; ++j) r += (ti[j] < ti[i]);
    pos[i] = r;
  }
  for (int t = 0; t < 3; ++t)
    for (int i = 0; i < 4; ++i)
      if (pos[i] == 0) out[3 * q + pos[partner_of(t, i)] - 1] = cnt[t];
}

// ---------------------------------------------------------------------------------------------
// Export: every non-zero (cell, topology) as a reference-packed key (graphs/quartet.go:67-109):
// taxa sorted by tip index in 15-bit fields, topology nibble with bit i set iff sorted taxon i is
// on the side NOT holding sorted taxon 0.
// ---------------------------------------------------------------------------------------------
struct ExportParams {
  const uint4* cells;
  BoxMap bm;
  unsigned long long* keys;
  uint32_t* counts;
  unsigned long long cap;
  unsigned long long* n_out;
};

__global__ void __launch_bounds__(256) k_export(ExportParams prm, ConstraintDev ct) {
  __shared__ uint32_t s_seg[4];
  CellCoords cc = decode_cell(prm.bm, ct.n_taxa, s_seg);
  if (!cc.valid) return;
  uint4 v = prm.cells[(size_t)blockIdx.x * 256 + threadIdx.x];
  if (!(v.x | v.y | v.z)) return;
  const uint32_t cnt[3] = {v.x, v.y, v.z};
  int ti[4] = {ct.tip_of_rank[cc.a], ct.tip_of_rank[cc.b], ct.tip_of_rank[cc.c], ct.tip_of_rank[cc.d]};
  int pos[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int r = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) r += (ti[j] < ti[i]);
    pos[i] = r;
  }
  unsigned long long base = 0;
  int s0 = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    base |= (unsigned long long)ti[i] << (15 * pos[i]);
    if (pos[i] == 0) s0 = i;
  }
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    if (!cnt[t]) continue;
    int p = partner_of(t, s0);
    unsigned nib = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j != s0 && j != p) nib |= 1u << pos[j];
    unsigned long long slot = atomicAdd(prm.n_out, 1ull);
    if (slot < prm.cap) {
      prm.keys[slot] = base | ((unsigned long long)nib << 60);
      prm.counts[slot] = cnt[t];
    }
  }
}

}  // namespace camus_dev
