// Reticulation support: score.ReticulationScore (internal/score/score.go:25-72) on the device.
//
// The reference walks, per gene tree, every emission of gotree's quartet enumerator (every
// internal edge x every (pair | pair) across it, duplicates included, score.go:43) and scores it
// against each reticulation with QuartetScore (internal/score/quartettotals.go:107-157). Here:
//   * a 4-taxon set is resolved once from the tree's LCA-depth table; the
//     number of times the enumerator emits it is the number of edges of the path between its two
//     cherries after tre.UnRoot() (score.go:38): the depth-sum difference, minus one when that
//     path runs through a binary root (whose two edges UnRoot merges);
//   * QuartetScore's tree look-ups are replaced by the per-(reticulation, tip) records built on
//     the host (camus_b200/host/retic.hpp reticulation_records), so no tree lives on the device.
// Everything is integer; the ratio supported / total is taken on the host in float64.
#pragma once
#include <cuda_runtime.h>

#include "layout.cuh"

namespace camus_dev {

constexpr int kRetMax = 64;        // reticulations per call (level-1 networks have few)
constexpr int kRetThreads = 256;

// R1: per gene tree (one CTA), LCA-depth table over NETWORK tip indices:
//   D[g][b][a] (a < b) = 1 + depth(LCA(a, b)), 0 = a taxon is absent from the tree (lower triangle:
//   the score kernel's threads differ in their smallest taxon, which is then the contiguous index).
// Dynamic shared memory: 8 B per node of the largest tree (parent, depth, leaf tip, leaf gap, u16 each).
__global__ void __launch_bounds__(kRetThreads) k_ret_depths(const int64_t* __restrict__ node_off, const int32_t* __restrict__ parent,
                                                            const int32_t* __restrict__ taxon, int n_tips, int max_nodes,
                                                            uint16_t* __restrict__ D, uint8_t* __restrict__ root_binary) {
  extern __shared__ uint16_t s_ret[];
  uint16_t* s_par = s_ret;
  uint16_t* s_dep = s_par + max_nodes;
  uint16_t* s_tip = s_dep + max_nodes;  // tip index of the k-th leaf in preorder
  uint16_t* s_gap = s_tip + max_nodes;  // 1 + depth(LCA(leaf k, leaf k + 1))
  __shared__ int s_nl;
  const int g = blockIdx.x, tid = threadIdx.x;
  const int64_t base = node_off[g];
  const int nn = (int)(node_off[g + 1] - base);
  for (int i = tid; i < nn; i += kRetThreads) s_par[i] = (uint16_t)parent[base + i];
  __syncthreads();
  if (tid == 0) {
    s_dep[0] = 0;
    int pos = 0, root_kids = 0;
    bool prev_leaf = false;
    for (int i = 0; i < nn; ++i) {
      if (i) {
        s_dep[i] = (uint16_t)(s_dep[s_par[i]] + 1);
        root_kids += s_par[i] == 0;
      }
      if (prev_leaf) s_gap[pos - 1] = s_dep[i];  // the node after a leaf hangs off LCA(leaf, next leaf)
      const int t = taxon[base + i];
      prev_leaf = t >= 0;
      if (t >= 0) s_tip[pos++] = (uint16_t)t;
    }
    s_nl = pos;
    root_binary[g] = root_kids == 2;
  }
  __syncthreads();
  const int nl = s_nl;
  uint16_t* Dg = D + (size_t)g * n_tips * n_tips;
  for (int p = tid; p + 1 < nl; p += kRetThreads) {
    const uint32_t a = s_tip[p];
    uint32_t m = 0xFFFFu;
    for (int q = p + 1; q < nl; ++q) {
      m = min(m, (uint32_t)s_gap[q - 1]);
      const uint32_t b = s_tip[q];
      Dg[(size_t)max(a, b) * n_tips + min(a, b)] = (uint16_t)m;
    }
  }
}

// QuartetScore (quartettotals.go:107-157) for taxa x[0] < x[1] < x[2] < x[3] (network tip
// indices) whose tree displays x[p0] x[p1] | the other two. rec = records of ONE reticulation.
// Returns 0 = Qdiff, 1 = Qneq, 2 = Qeq.
__device__ __forceinline__ int ret_quartet_score(const uint4* __restrict__ rec, const uint32_t x[4], const int pair_of[4],
                                                 uint32_t n_leaves) {
  uint4 r[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) r[i] = __ldg(rec + x[i]);
  // uniqueTaxaBelowNodeFromQ (:160-171): exactly one taxon under w
  int bi = -1, nw = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (r[i].w & 1u) {
      if (bi < 0) bi = i;
      ++nw;
    }
  if (nw != 1) return 0;
  const bool bottom_in_u = (r[bi].w & 8u) != 0;
  uint32_t node[4], dep[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t f = r[i].w;
    if (!(f & 2u)) {  // outside L(v): node 0, the root (:119)
      node[i] = 0;
      dep[i] = 0;
    } else if ((f & 4u) || bottom_in_u) {
      node[i] = r[i].x;
      dep[i] = r[i].z & 0xFFFFu;
    } else {
      node[i] = r[i].y;
      dep[i] = r[i].z >> 16;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)  // dups (:185-194)
#pragma unroll
    for (int j = i + 1; j < 4; ++j)
      if (node[i] == node[j]) return 0;
  const int neighbor = pair_of[bi];  // neighborTaxaQ (:174-182)
  uint32_t minW = n_leaves;
  int maxU = -1, best = 0;
  bool in_u = false;
#pragma unroll
  for (int i = 0; i < 4; ++i) {  // :141-155, taxa in ascending tip-index order
    const bool wsub = (r[i].w & 4u) != 0;
    if (!in_u && wsub && dep[i] < minW) {
      minW = dep[i];
      best = i;
    } else if (!wsub && (int)dep[i] > maxU) {
      in_u = true;
      maxU = (int)dep[i];
      best = i;
    }
  }
  return best == neighbor ? 2 : 1;
}

// R2: grid = (pairs c < d of network tips, gene tree). The block stages the two depth rows D[c][.]
// and D[d][.] in shared memory and its threads walk the pairs a < b below c, so a 4-set costs one
// gathered global load (D[b][a], contiguous in a) and a cheap 32-bit pair unranking. Most (4-set,
// reticulation) pairs are Qdiff because QuartetScore wants exactly one of the four taxa under w
// (quartettotals.go:160-171): under_w[t] holds that flag for all reticulations of tip t as a bit
// mask, so a thread only evaluates the reticulations that pass. Emission counts go to shared
// counters (64-bit: a block sees up to C(c, 2) sets) and once per block to the global table.
__global__ void __launch_bounds__(kRetThreads) k_ret_score(int n_tips, int n_ret, const uint4* __restrict__ rec,
                                                           const unsigned long long* __restrict__ under_w,
                                                           const uint16_t* __restrict__ D, const uint8_t* __restrict__ root_binary,
                                                           unsigned long long* __restrict__ totals,
                                                           unsigned long long* __restrict__ supported) {
  extern __shared__ uint16_t s_rows[];  // [2][n_tips]
  __shared__ unsigned long long s_tot[kRetMax], s_sup[kRetMax];
  const int tid = threadIdx.x, g = blockIdx.y;
  uint32_t c, d;
  pair_unrank(blockIdx.x, c, d);
  if (c < 2) return;
  const uint16_t* Dg = D + (size_t)g * n_tips * n_tips;
  const uint32_t cd = Dg[(size_t)d * n_tips + c];
  if (cd == 0) return;  // c or d is not in this tree
  uint16_t* rowc = s_rows;
  uint16_t* rowd = s_rows + n_tips;
  for (uint32_t i = tid; i < c; i += kRetThreads) {
    rowc[i] = Dg[(size_t)c * n_tips + i];
    rowd[i] = Dg[(size_t)d * n_tips + i];
  }
  if (tid < kRetMax) s_tot[tid] = s_sup[tid] = 0;
  __syncthreads();
  const unsigned long long wc = under_w[c], wd = under_w[d];
  const unsigned long long wcd_any = wc | wd, wcd_two = wc & wd;
  const bool rb = root_binary[g] != 0;
  const uint32_t n_pairs = c * (c - 1) / 2;
  for (uint32_t p = tid; p < n_pairs; p += kRetThreads) {
    // pair (a < b) of rank choose2(b) + a, 32-bit
    uint32_t b = (uint32_t)((1.0f + sqrtf(1.0f + 8.0f * (float)p)) * 0.5f);
    while (b * (b - 1) / 2 > p) --b;
    while ((b + 1) * b / 2 <= p) ++b;
    const uint32_t a = p - b * (b - 1) / 2;
    const unsigned long long wa = under_w[a], wb = under_w[b];
    unsigned long long todo = (wa | wb | wcd_any) & ~((wa & wb) | ((wa | wb) & wcd_any) | wcd_two);
    if (!todo) continue;
    const uint32_t ab = Dg[(size_t)b * n_tips + a], ac = rowc[a], bc = rowc[b], ad = rowd[a], bd = rowd[b];
    const uint32_t s1 = ab + cd, s2 = ac + bd, s3 = ad + bc;
    const uint32_t mx = max(s1, max(s2, s3));
    if (!(ab && ac && ad && bc && bd) || (s1 == mx) + (s2 == mx) + (s3 == mx) != 1) continue;
    const uint32_t lo = min(s1, min(s2, s3));  // the two smaller sums are equal in a tree
    // emissions = edges between the two cherries; a binary root on that path (both cross LCAs
    // are the root: depth sums 1 + 1) loses one edge to UnRoot
    const uint32_t mult = mx - lo - ((rb && lo == 2) ? 1u : 0u);
    const uint32_t x[4] = {a, b, c, d};
    int pair_of[4];
    if (s1 == mx) { pair_of[0] = 1; pair_of[1] = 0; pair_of[2] = 3; pair_of[3] = 2; }
    else if (s2 == mx) { pair_of[0] = 2; pair_of[2] = 0; pair_of[1] = 3; pair_of[3] = 1; }
    else { pair_of[0] = 3; pair_of[3] = 0; pair_of[1] = 2; pair_of[2] = 1; }
    while (todo) {
      const int rr = __ffsll((long long)todo) - 1;
      todo &= todo - 1;
      const int comp = ret_quartet_score(rec + (size_t)rr * n_tips, x, pair_of, (uint32_t)n_tips);
      if (comp) atomicAdd(&s_tot[rr], (unsigned long long)mult);
      if (comp == 2) atomicAdd(&s_sup[rr], (unsigned long long)mult);
    }
  }
  __syncthreads();
  if (tid < n_ret) {
    if (s_tot[tid]) atomicAdd(&totals[(size_t)g * n_ret + tid], s_tot[tid]);
    if (s_sup[tid]) atomicAdd(&supported[(size_t)g * n_ret + tid], s_sup[tid]);
  }
}

}  // namespace camus_dev
