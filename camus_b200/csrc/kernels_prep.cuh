// Stage 1 kernels: flat gene trees -> per-tree LCA-depth tables -> rooted-triple bit-planes.
//
// Replaces what gotree's bipartition enumerator does behind graphs.QuartetsFromTree
// (internal/graphs/quartet.go:112-123, gotree Quartets at :119): instead of enumerating
// (pair | pair) emissions per edge, every gene tree is reduced to the rooted-triple relation
// "D(a,b) > D(a,c)" on its LCA depths, stored as one bit per tree (32 trees per word), from which
// the count kernel resolves 32 trees per logical instruction.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "layout.cuh"

namespace camus_dev {

constexpr int kPlanePad = 516;  // padded plane stride (words) of the shared tile image in K1b

// ---------------------------------------------------------------------------------------------
// K0: one thread per gene tree walks its preorder arrays once.
//   depth[i] = depth[parent[i]] + 1
//   leaves in DFS order -> lr[pos] = DFS rank (in the CONSTRAINT tree) of the pos-th leaf
//   h[pos] = 1 + depth of LCA(leaf pos, leaf pos+1) = depth of the node that follows leaf pos in
//            preorder (it hangs off that LCA). 0 is reserved for "taxon absent".
// Outputs are interleaved by tree group: [group][pos][lane], lane = tree % 32.
// ---------------------------------------------------------------------------------------------
__global__ void k_tree_prepare(int n_trees, const int64_t* __restrict__ node_off, const int32_t* __restrict__ parent,
                               const int32_t* __restrict__ taxon, const int32_t* __restrict__ rank_of_tip,
                               uint16_t* __restrict__ depth_tmp, uint16_t* __restrict__ lrT, uint16_t* __restrict__ hT,
                               int npad, int32_t* __restrict__ nleaf) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_trees) return;
  int64_t base = node_off[g];
  int nn = (int)(node_off[g + 1] - base);
  size_t grp = (size_t)(g >> 5);
  int lane = g & 31;
  int pos = 0;
  bool prev_leaf = false;
  for (int i = 0; i < nn; ++i) {
    int p = parent[base + i];
    uint16_t d = p < 0 ? 0 : (uint16_t)(depth_tmp[base + p] + 1);
    depth_tmp[base + i] = d;
    if (prev_leaf) {
      hT[(grp * npad + (pos - 1)) * kLanes + lane] = d;  // = depth(LCA) + 1
      prev_leaf = false;
    }
    int t = taxon[base + i];
    if (t >= 0) {
      lrT[(grp * npad + pos) * kLanes + lane] = (uint16_t)rank_of_tip[t];
      ++pos;
      prev_leaf = true;
    }
  }
  nleaf[g] = pos;
}

// ---------------------------------------------------------------------------------------------
// K1: LCA-depth table of every tree of a group batch, in constraint DFS-rank coordinates:
//   Dt[gb][lo][hi][lane] = 1 + depth of LCA(lo, hi) in tree (grp0+gb)*32+lane, 0 if absent.
// D(p, q) for DFS positions p < q of one tree is the running minimum of h[p..q-1]. Warp = one
// start position p, lane = tree; h/lr reads are coalesced 64 B rows.
// ---------------------------------------------------------------------------------------------
// kHalf: store the value as an fp16 bit pattern (exact up to 2048) for the packed-compare K1b; the
// buffer is then pre-filled with 0xFFFF = NaN, so "absent" fails every ordered comparison.
template <bool kHalf>
__global__ void k_depth_table(int grp0, int n_trees, int npad, const int32_t* __restrict__ nleaf,
                              const uint16_t* __restrict__ lrT, const uint16_t* __restrict__ hT,
                              uint16_t* __restrict__ Dt) {
  int gb = blockIdx.y;
  size_t grp = (size_t)(grp0 + gb);
  int lane = threadIdx.x;
  int g = (int)grp * kLanes + lane;
  int p = blockIdx.x * blockDim.y + threadIdx.y;
  int nl = g < n_trees ? nleaf[g] : 0;
  if (p + 1 >= nl) return;
  const uint16_t* lr = lrT + grp * npad * kLanes + lane;
  const uint16_t* h = hT + grp * npad * kLanes + lane;
  uint16_t* D = Dt + (size_t)gb * npad * npad * kLanes + lane;
  uint32_t a = lr[(size_t)p * kLanes];
  uint32_t m = 0xFFFFu;
  for (int q = p + 1; q < nl; ++q) {
    m = min(m, (uint32_t)h[(size_t)(q - 1) * kLanes]);
    uint32_t b = lr[(size_t)q * kLanes];
    uint32_t lo = min(a, b), hi = max(a, b);
    D[((size_t)lo * npad + hi) * kLanes] = kHalf ? __half_as_ushort(__ushort2half_rn((unsigned short)m)) : (uint16_t)m;
  }
}

// ---------------------------------------------------------------------------------------------
// K1b: rooted-triple bit-planes for taxa a < b < c (DFS ranks), lane = tree.
//   plane 0 (ab|c): D(a,b) > D(a,c)      plane 1 (ac|b): D(a,c) > D(a,b)      plane 2 (bc|a): D(b,c) > D(a,b)
// each AND "all three taxa present" (the != 0 tests). A polytomy (all three equal) sets no bit.
// __ballot_sync transposes 32 per-tree predicates into one word; lanes 0..2 store the 3 planes at
// Tri[tile(a,b,c)][group][tile_word(plane, a%8, b%8, c%8)].
// ---------------------------------------------------------------------------------------------
// One CTA builds one whole tile (s0 <= s1 <= s2) of one tree group: the 3 x 64 pair-depth rows it
// needs are staged in shared memory (coalesced 64 B rows of Dt), 8 warps x 64 triples ballot into
// a shared 6 KB image of the tile, which is then written to HBM with 128-bit coalesced stores.
// Every word of every tile is written, so Tri needs no memset.
__global__ void __launch_bounds__(256) k_triple_planes(int grp0, int n_batch, int n_groups_total, int npad,
                                                       const uint16_t* __restrict__ Dt, uint32_t* __restrict__ Tri) {
  __shared__ __align__(16) uint16_t s_d[3][64][kLanes];  // [ab | ac | bc][i*8+j][lane]
  // Tile image: the 3 planes are padded to 516 words so that lanes 0..2 (one plane each) hit
  // different banks, and the dummy row the other 29 lanes write to sits 16 banks away.
  __shared__ __align__(128) uint32_t s_img[3 * kPlanePad + 4 + 8 * 64];
  uint32_t* const s_out = s_img;
  uint32_t* const s_scr = s_img + 3 * kPlanePad + 4;
  __shared__ uint32_t s_seg[3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t tile = blockIdx.x;
  if (tid == 0) {
    uint64_t r = tile;
    uint32_t s2 = unrank3(r);
    r -= choose3((uint64_t)s2 + 2);
    uint32_t s1 = unrank2(r);
    r -= choose2((uint64_t)s1 + 1);
    s_seg[0] = (uint32_t)r;
    s_seg[1] = s1;
    s_seg[2] = s2;
  }
  __syncthreads();
  const uint32_t a0 = s_seg[0] * 8, b0 = s_seg[1] * 8, c0 = s_seg[2] * 8;
  // the CTA walks all tree groups of the batch: the tile unranking and address set-up are paid once
  for (int gb = 0; gb < n_batch; ++gb) {
  if (gb) __syncthreads();  // s_d / s_out of the previous group are free again
  const uint16_t* D = Dt + (size_t)gb * npad * npad * kLanes;
  // stage the 3 x 64 pair rows: for a fixed first taxon the 8 rows (y0..y0+7) are 512 contiguous
  // bytes of Dt, so one warp-wide 128-bit load moves a whole (which, i) segment
#pragma unroll
  for (int it = 0; it < 3; ++it) {
    const int seg = it * 8 + warp;  // 0..23
    const int which = seg >> 3, i = seg & 7, j = lane >> 2;
    const uint32_t x = (which == 2 ? b0 : a0) + i;
    const uint32_t y0 = which == 0 ? b0 : c0;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (x < y0 + j) v = *reinterpret_cast<const uint4*>(D + ((size_t)x * npad + y0 + j) * kLanes + (lane & 3) * 8);
    *reinterpret_cast<uint4*>(&s_d[which][i * 8 + j][(lane & 3) * 8]) = v;
  }
  __syncthreads();
  // warp = z (third taxon of the tile); the 8 D(a,c) values are hoisted, D(b,c) once per y, so a
  // triple costs one shared load, a 3-way min, four compares, three ballots and one store.
  const int z = warp;
  uint32_t Dac[8];
#pragma unroll
  for (int x = 0; x < 8; ++x) Dac[x] = s_d[1][x * 8 + z][lane];
  // lanes 0..2 store planes 0..2 at tile_word(lane, x, y, z); the other lanes store to a per-warp
  // scratch row so that the store needs no predicate / divergence. The plane is picked with two
  // bitwise muxes on lane-constant masks instead of compare + select.
  uint32_t* out_lane = lane < 3 ? s_out + lane * kPlanePad + z * 64 : s_scr + warp * 64;
  const uint32_t m0 = lane == 0 ? 0xFFFFFFFFu : 0u, m01 = lane <= 1 ? 0xFFFFFFFFu : 0u;
#pragma unroll
  for (int y = 0; y < 8; ++y) {
    const uint32_t Dbc = s_d[2][y * 8 + z][lane];
#pragma unroll
    for (int x = 0; x < 8; ++x) {
      const uint32_t Dab = s_d[0][x * 8 + y][lane];
      const bool ok = min(Dab, min(Dac[x], Dbc)) != 0;  // all present, and a < b < c in rank order
      const uint32_t w0 = __ballot_sync(0xFFFFFFFFu, ok && Dab > Dac[x]);
      const uint32_t w1 = __ballot_sync(0xFFFFFFFFu, ok && Dac[x] > Dab);
      const uint32_t w2 = __ballot_sync(0xFFFFFFFFu, ok && Dbc > Dab);
      const uint32_t w01 = (w0 & m0) | (w1 & ~m0);
      out_lane[((y >> 1) * 4 + (x >> 1)) * 4 + (y & 1) * 2 + (x & 1)] = (w01 & m01) | (w2 & ~m01);
    }
  }
  __syncthreads();
  uint4* dst = reinterpret_cast<uint4*>(Tri + (tile * n_groups_total + (grp0 + gb)) * kTileWords);
  for (int i = tid; i < kTileWords / 4; i += 256)
    dst[i] = *reinterpret_cast<const uint4*>(s_out + (i >> 7) * kPlanePad + (i & 127) * 4);
  }
}

// ---------------------------------------------------------------------------------------------
// K1b, packed form (default): depths are fp16 (exact: the host checks depth + 1 <= 2048), absent
// pairs are NaN. One setp.gt.f16x2 compares TWO triples (y, y+1) per lane and yields two
// predicates; NaN makes every comparison false, so no presence test is needed. The integer
// kernel above spends 6 ALU-pipe instructions per triple (min3, 4 ISETP, SEL) and is bound by that
// pipe; here the compares run on the FMA pipe and a triple costs ~8 issue slots.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cmp2_ballot(uint32_t a2, uint32_t b2, uint32_t& lo, uint32_t& hi) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "setp.gt.f16x2 p|q, %2, %3;\n"
      "vote.sync.ballot.b32 %0, p, 0xffffffff;\n"
      "vote.sync.ballot.b32 %1, q, 0xffffffff;\n"
      "}\n"
      : "=r"(lo), "=r"(hi)
      : "r"(a2), "r"(b2));
}

__global__ void __launch_bounds__(256) k_triple_planes_h2(int grp0, int n_batch, int n_groups_total, int npad,
                                                          const uint16_t* __restrict__ Dt, uint32_t* __restrict__ Tri) {
  __shared__ __align__(16) uint16_t s_d[3][64][kLanes];  // [ab | ac | bc][i*8+j][lane], fp16 bits, NaN = absent
  // Tile image: the 3 planes are padded to 516 words so that lanes 0..2 (one plane each) hit
  // different banks, and the dummy row the other 29 lanes write to sits 16 banks away.
  __shared__ __align__(128) uint32_t s_img[3 * kPlanePad + 4 + 8 * 64];
  uint32_t* const s_out = s_img;
  uint32_t* const s_scr = s_img + 3 * kPlanePad + 4;
  __shared__ uint32_t s_seg[3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t tile = blockIdx.x;
  if (tid == 0) {
    uint64_t r = tile;
    uint32_t s2 = unrank3(r);
    r -= choose3((uint64_t)s2 + 2);
    uint32_t s1 = unrank2(r);
    r -= choose2((uint64_t)s1 + 1);
    s_seg[0] = (uint32_t)r;
    s_seg[1] = s1;
    s_seg[2] = s2;
  }
  __syncthreads();
  const uint32_t a0 = s_seg[0] * 8, b0 = s_seg[1] * 8, c0 = s_seg[2] * 8;
  const int z = warp;
  uint32_t* out_lane = lane < 3 ? s_out + lane * kPlanePad + z * 64 : s_scr + warp * 64;
  const uint32_t m0 = lane == 0 ? 0xFFFFFFFFu : 0u, m01 = lane <= 1 ? 0xFFFFFFFFu : 0u;
  for (int gb = 0; gb < n_batch; ++gb) {
    if (gb) __syncthreads();
    const uint16_t* D = Dt + (size_t)gb * npad * npad * kLanes;
    // stage the 3 x 64 pair rows with warp-wide 128-bit copies (8 rows = 512 contiguous bytes)
#pragma unroll
    for (int it = 0; it < 3; ++it) {
      const int seg = it * 8 + warp;  // 0..23
      const int which = seg >> 3, i = seg & 7, j = lane >> 2;
      const uint32_t x = (which == 2 ? b0 : a0) + i;
      const uint32_t y0 = which == 0 ? b0 : c0;
      uint4 v = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);  // NaN: wrong rank order
      if (x < y0 + j) v = *reinterpret_cast<const uint4*>(D + ((size_t)x * npad + y0 + j) * kLanes + (lane & 3) * 8);
      *reinterpret_cast<uint4*>(&s_d[which][i * 8 + j][(lane & 3) * 8]) = v;
    }
    __syncthreads();
    uint32_t ac2[8];
#pragma unroll
    for (int x = 0; x < 8; ++x) {
      const uint32_t h = s_d[1][x * 8 + z][lane];
      ac2[x] = h | (h << 16);
    }
#pragma unroll
    for (int yp = 0; yp < 4; ++yp) {
      const uint32_t bc2 = (uint32_t)s_d[2][(2 * yp) * 8 + z][lane] | ((uint32_t)s_d[2][(2 * yp + 1) * 8 + z][lane] << 16);
#pragma unroll
      for (int x = 0; x < 8; ++x) {
        const uint32_t ab2 = (uint32_t)s_d[0][x * 8 + 2 * yp][lane] | ((uint32_t)s_d[0][x * 8 + 2 * yp + 1][lane] << 16);
        uint32_t w0a, w0b, w1a, w1b, w2a, w2b;
        cmp2_ballot(ab2, ac2[x], w0a, w0b);  // ab|c : D(a,b) > D(a,c)
        cmp2_ballot(ac2[x], ab2, w1a, w1b);  // ac|b : D(a,c) > D(a,b)
        cmp2_ballot(bc2, ab2, w2a, w2b);     // bc|a : D(b,c) > D(a,b)
        const uint32_t sa_ = (((w0a & m0) | (w1a & ~m0)) & m01) | (w2a & ~m01);
        const uint32_t sb_ = (((w0b & m0) | (w1b & ~m0)) & m01) | (w2b & ~m01);
        // y = 2*yp and 2*yp+1: tile_word offsets ((y>>1)*4 + (x>>1))*4 + (y&1)*2 + (x&1)
        out_lane[(yp * 4 + (x >> 1)) * 4 + (x & 1)] = sa_;
        out_lane[(yp * 4 + (x >> 1)) * 4 + 2 + (x & 1)] = sb_;
      }
    }
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(Tri + (tile * n_groups_total + (grp0 + gb)) * kTileWords);
    for (int i = tid; i < kTileWords / 4; i += 256)
      dst[i] = *reinterpret_cast<const uint4*>(s_out + (i >> 7) * kPlanePad + (i & 127) * 4);
  }
}

// ---------------------------------------------------------------------------------------------
// Constraint tree: LCA of every leaf pair (DFS ranks), packed (depth << 16 | node id). Node ids
// are preorder, so "y in subtree(x)" is an interval test.
// ---------------------------------------------------------------------------------------------
__global__ void k_constraint_lca(int n_taxa, const int32_t* __restrict__ leafnode, const int32_t* __restrict__ parent,
                                 const int32_t* __restrict__ sz, const int32_t* __restrict__ depth,
                                 uint32_t* __restrict__ lcaD) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  int a = blockIdx.y;
  if (b >= n_taxa) return;
  int x = leafnode[a], y = leafnode[b];
  if (a > b) {
    int t = x;
    x = y;
    y = t;
  }
  while (!(x <= y && y < x + sz[x])) x = parent[x];
  lcaD[(size_t)a * n_taxa + b] = ((uint32_t)depth[x] << 16) | (uint32_t)x;
}

}  // namespace camus_dev
