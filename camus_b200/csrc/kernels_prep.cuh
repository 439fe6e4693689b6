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

#include <type_traits>

#include "layout.cuh"
#include "ptx.cuh"

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

// K0, staged form (default): one warp per gene tree. The parent array is staged in shared memory
// (u16 local indices), lane 0 resolves the depths there (the serial form above pays two dependent
// L2 round trips per node), then all lanes emit lr / h with a ballot prefix over the leaves.
// Dynamic shared memory: 4 bytes per node of the largest tree (the host falls back to the serial
// kernel above kTreePrepareMaxNodes).
constexpr int kTreePrepareMaxNodes = 24576;
__global__ void __launch_bounds__(32) k_tree_prepare_warp(int n_trees, const int64_t* __restrict__ node_off,
                                                          const int32_t* __restrict__ parent, const int32_t* __restrict__ taxon,
                                                          const int32_t* __restrict__ rank_of_tip, uint16_t* __restrict__ lrT,
                                                          uint16_t* __restrict__ hT, int npad, int32_t* __restrict__ nleaf,
                                                          int max_nodes, uint16_t* __restrict__ lrP, uint16_t* __restrict__ hP,
                                                          uint16_t* __restrict__ posP) {
  extern __shared__ uint16_t s_tree[];
  uint16_t* s_par = s_tree;
  uint16_t* s_dep = s_tree + max_nodes;
  const int g = blockIdx.x, lane = threadIdx.x;
  if (g >= n_trees) return;
  const int64_t base = node_off[g];
  const int nn = (int)(node_off[g + 1] - base);
  for (int i = lane; i < nn; i += 32) s_par[i] = (uint16_t)parent[base + i];  // root: 0xFFFF, never read
  __syncwarp();
  if (lane == 0) {
    s_dep[0] = 0;
    for (int i = 1; i < nn; ++i) s_dep[i] = (uint16_t)(s_dep[s_par[i]] + 1);
  }
  __syncwarp();
  const size_t row0 = (size_t)(g >> 5) * npad;
  const int gl = g & 31;
  int pos = 0;
  for (int i0 = 0; i0 < nn; i0 += 32) {
    const int i = i0 + lane;
    const int t = i < nn ? taxon[base + i] : -1;
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, t >= 0);
    if (t >= 0) {
      const int k = pos + __popc(m & ((1u << lane) - 1u));
      const uint16_t rk = (uint16_t)rank_of_tip[t];
      lrT[(row0 + k) * kLanes + gl] = rk;
      if (i + 1 < nn) hT[(row0 + k) * kLanes + gl] = s_dep[i + 1];  // node after a leaf hangs off LCA(leaf, next leaf)
      if (lrP) {  // per-tree copies for k_depth_slices_direct: [tree][position] and the inverse [tree][rank] -> position + 1
        lrP[(size_t)g * npad + k] = rk;
        hP[(size_t)g * npad + k] = i + 1 < nn ? s_dep[i + 1] : (uint16_t)0xFFFF;
        posP[(size_t)g * npad + rk] = (uint16_t)(k + 1);
      }
    }
    pos += __popc(m);
  }
  if (lane == 0) nleaf[g] = pos;
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
// tile_list: class-scheme partitions build only the tiles this rank reads (layout.cuh) and store
// them compactly, tile tile_list[i] in slot i of Tri; nullptr = all tiles, slot = tile index.
__global__ void __launch_bounds__(256) k_triple_planes(int grp0, int n_batch, int n_groups_total, int npad,
                                                       const uint16_t* __restrict__ Dt, uint32_t* __restrict__ Tri,
                                                       const uint32_t* __restrict__ tile_list) {
  __shared__ __align__(16) uint16_t s_d[3][64][kLanes];  // [ab | ac | bc][i*8+j][lane]
  // Tile image: the 3 planes are padded to 516 words so that lanes 0..2 (one plane each) hit
  // different banks, and the dummy row the other 29 lanes write to sits 16 banks away.
  __shared__ __align__(128) uint32_t s_img[3 * kPlanePad + 4 + 8 * 64];
  uint32_t* const s_out = s_img;
  uint32_t* const s_scr = s_img + 3 * kPlanePad + 4;
  __shared__ uint32_t s_seg[3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t tile = tile_list ? tile_list[blockIdx.x] : blockIdx.x;
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
  uint4* dst = reinterpret_cast<uint4*>(Tri + ((tile_list ? (uint64_t)blockIdx.x : tile) * n_groups_total + (grp0 + gb)) * kTileWords);
  for (int i = tid; i < kTileWords / 4; i += 256)
    dst[i] = *reinterpret_cast<const uint4*>(s_out + (i >> 7) * kPlanePad + (i & 127) * 4);
  }
}

// ---------------------------------------------------------------------------------------------
// K1b, packed form (default): depths are fp16 (exact: the host checks depth + 1 <= 2048), absent
// pairs are NaN, which makes every comparison false, so no presence test is needed.
//
// The transpose "lane = tree" -> "bit = tree" is NOT done with ballots here. A ballot costs one
// issue slot per 32-bit output word plus the compare, the operand loads and the store of a word
// that only one lane keeps (~2.8 slots per word in total; the kernel was bound by instruction
// issue). Instead every thread owns one (x, y) pair of the tile and two z, walks the 32 trees of
// the group two at a time (set.gt.u32.f16x2 yields a 0xFFFF mask per half) and ORs the masked
// result into its own accumulator word: 1 HSET2 + 1 LOP3 per two output bits, operands fetched
// with 128-bit shared loads (8 trees each), D(a,b) of all 32 trees held in registers across z.
// That is ~1.1 issue slots per output word.
//
// Bit order inside a word: tree 2k of the group sits in bit k, tree 2k+1 in bit 16+k. The count
// kernel only popcounts ANDs of words of the same group, so any fixed permutation is equivalent.
//
// Shared rows are 64 B (32 trees x fp16); the 16-byte chunk c of row r is stored at chunk
// c ^ ((r >> 1) & 3), so that the 8 lanes of a 128-bit load phase (8 consecutive rows) cover all
// 32 banks.
//
// kComplete (every gene tree complete, binary, with a binary root): off the diagonal exactly one
// of the three planes is set per tree, so the third is ~(w0 | w1) & (trees of the group).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t gt2_mask(uint32_t a, uint32_t b) {
  uint32_t m;
  asm("set.gt.u32.f16x2 %0, %1, %2;" : "=r"(m) : "r"(a), "r"(b));
  return m;
}

// The same comparison on the FMA pipe: depths are integers, so sat(a - b) is 1.0 where a > b and
// 0.0 elsewhere (.sat also flushes NaN to +0). HSET2 and LOP3 both issue on the ALU pipe, which
// bound the kernel (ncu: alu 70 %, fma 9 %); half of the 32 trees take this route instead.
__device__ __forceinline__ uint32_t gt2_unit(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("fma.rn.sat.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(b), "r"(0xBC00BC00u), "r"(a));  // b * -1 + a
  return d;
}
__device__ __forceinline__ uint32_t fma2(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// u16 offset of chunk c of row r. kPad: 16 B after every 8 rows, for the array whose rows are read
// at a stride of 8 by the quarter-warps of one load (ac: 4 different x, same z).
template <bool kPad>
__device__ __forceinline__ int plane_chunk_off(int r, int c) {
  return r * kLanes + ((c ^ ((r >> 1) & 3)) << 3) + (kPad ? (r >> 3) << 3 : 0);
}

template <bool kDerive>
__device__ __forceinline__ void triple_words(const uint4 (&ab)[4], const uint16_t* s_ac, int r_ac, const uint16_t* s_bc,
                                             int r_bc, uint32_t vm, uint32_t (&w)[3]) {
  w[0] = w[1] = w[2] = 0;
  // trees 16..31 of the group: fp16 accumulators 1024 + sum of 2^k, exact below 2048; the low 8
  // mantissa bits of each half are the 8 result bits
  uint32_t f[3] = {0x64006400u, 0x64006400u, 0x64006400u};
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint4 ac = *reinterpret_cast<const uint4*>(s_ac + plane_chunk_off<true>(r_ac, c));
    const uint32_t abv[4] = {ab[c].x, ab[c].y, ab[c].z, ab[c].w};
    const uint32_t acv[4] = {ac.x, ac.y, ac.z, ac.w};
    uint32_t bcv[4] = {0, 0, 0, 0};
    if (!kDerive) {
      const uint4 bc = *reinterpret_cast<const uint4*>(s_bc + plane_chunk_off<false>(r_bc, c));
      bcv[0] = bc.x; bcv[1] = bc.y; bcv[2] = bc.z; bcv[3] = bc.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (c < 2) {  // ALU pipe: mask compare + masked OR
        const uint32_t bit = 0x00010001u << (c * 4 + j);
        w[0] |= gt2_mask(abv[j], acv[j]) & bit;  // ab|c : D(a,b) > D(a,c)
        w[1] |= gt2_mask(acv[j], abv[j]) & bit;  // ac|b : D(a,c) > D(a,b)
        if (!kDerive) w[2] |= gt2_mask(bcv[j], abv[j]) & bit;  // bc|a : D(b,c) > D(a,b)
      } else {  // FMA pipe: unit compare + weighted accumulate
        const uint32_t wt = 0x3C003C00u + 0x04000400u * ((c - 2) * 4 + j);  // 2^k in both halves
        f[0] = fma2(gt2_unit(abv[j], acv[j]), wt, f[0]);
        f[1] = fma2(gt2_unit(acv[j], abv[j]), wt, f[1]);
        if (!kDerive) f[2] = fma2(gt2_unit(bcv[j], abv[j]), wt, f[2]);
      }
    }
  }
#pragma unroll
  for (int p = 0; p < (kDerive ? 2 : 3); ++p) w[p] |= (f[p] & 0x00FF00FFu) << 8;
  if (kDerive) w[2] = ~(w[0] | w[1]) & vm;
}

#ifndef CAMUS_PLANES_MINB
#define CAMUS_PLANES_MINB 4  // CTAs per SM the packed plane builder is compiled for (A/B builds: -DCAMUS_PLANES_MINB=n)
#endif
template <bool kComplete>
__global__ void __launch_bounds__(256, CAMUS_PLANES_MINB) k_triple_planes_h2(int grp0, int n_batch, int n_groups_total, int n_trees, int npad,
                                                          const uint16_t* __restrict__ Dt, uint32_t* __restrict__ Tri,
                                                          const uint32_t* __restrict__ tile_list) {
  // fp16 bits, NaN = absent. Rows: ab[x*8+y], ac[x*8+z], bc[z*8+y] (the index that varies across
  // the lanes of a load phase is the low one).
  __shared__ __align__(16) uint16_t s_d[3][64 * kLanes + 64];
  __shared__ __align__(16) uint32_t s_img[kTileWords];
  __shared__ uint32_t s_seg[3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t tile = tile_list ? tile_list[blockIdx.x] : blockIdx.x;
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
  const bool derive = kComplete && s_seg[0] < s_seg[1] && s_seg[1] < s_seg[2];  // every (x, y, z) is in rank order
  const int y = tid & 7, x = (tid >> 3) & 7, zh = tid >> 6;

  // staging: per group 3 x 64 rows; a warp copies 8 rows (512 contiguous bytes) per 128-bit access
  const uint16_t* src[3];
  uint16_t* dst[3];
#pragma unroll
  for (int it = 0; it < 3; ++it) {
    const int seg = it * 8 + warp;  // 0..23
    const int which = seg >> 3, i = seg & 7, j = lane >> 2, c = lane & 3;
    const uint32_t lo = (which == 2 ? b0 : a0) + i;
    const uint32_t hi = (which == 0 ? b0 : c0) + j;
    src[it] = lo < hi ? Dt + ((size_t)lo * npad + hi) * kLanes + c * 8 : nullptr;  // null: wrong rank order -> NaN
    const int r = which == 2 ? j * 8 + i : i * 8 + j;
    dst[it] = &s_d[which][which == 1 ? plane_chunk_off<true>(r, c) : plane_chunk_off<false>(r, c)];
  }
  const size_t group_stride = (size_t)npad * npad * kLanes;
  const uint4 nan4 = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
  uint4 pre[3];
#pragma unroll
  for (int it = 0; it < 3; ++it) pre[it] = src[it] ? __ldg(reinterpret_cast<const uint4*>(src[it])) : nan4;

  for (int gb = 0; gb < n_batch; ++gb) {
#pragma unroll
    for (int it = 0; it < 3; ++it) *reinterpret_cast<uint4*>(dst[it]) = pre[it];
    __syncthreads();
    if (gb + 1 < n_batch) {  // next group's rows travel while this one is computed
#pragma unroll
      for (int it = 0; it < 3; ++it)
        pre[it] = src[it] ? __ldg(reinterpret_cast<const uint4*>(src[it] + (size_t)(gb + 1) * group_stride)) : nan4;
    }
    const int left = n_trees - (grp0 + gb) * kLanes;  // trees in this group
    const int n_even = left >= 32 ? 16 : (left > 0 ? (left + 1) >> 1 : 0), n_odd = left >= 32 ? 16 : (left > 0 ? left >> 1 : 0);
    const uint32_t vm = ((1u << n_even) - 1u) | (((1u << n_odd) - 1u) << 16);
    uint4 ab[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) ab[c] = *reinterpret_cast<const uint4*>(s_d[0] + plane_chunk_off<false>(x * 8 + y, c));
#pragma unroll
    for (int zz = 0; zz < 2; ++zz) {
      const int z = 2 * zh + zz;
      uint32_t w[3];
      if (derive) triple_words<true>(ab, s_d[1], x * 8 + z, s_d[2], z * 8 + y, vm, w);
      else triple_words<false>(ab, s_d[1], x * 8 + z, s_d[2], z * 8 + y, vm, w);
#pragma unroll
      for (int p = 0; p < 3; ++p) s_img[tile_word(p, x, y, z)] = w[p];
    }
    __syncthreads();
    uint4* out = reinterpret_cast<uint4*>(Tri + ((tile_list ? (uint64_t)blockIdx.x : tile) * n_groups_total + (grp0 + gb)) * kTileWords);
    for (int i = tid; i < kTileWords / 4; i += 256) out[i] = reinterpret_cast<const uint4*>(s_img)[i];
  }
}

// ---------------------------------------------------------------------------------------------
// K1b, bit-sliced form (default since round 2, depth + 1 < 2048): the 32 trees of a group are compared
// AT ONCE instead of two per instruction.
//
// k_depth_slices transposes every pair row of the integer depth table (32 trees x u16, lane = tree)
// into NB bit slices: word k of Ds[group][lo][hi] holds bit k of the depth of all 32 trees (bit =
// tree % 32), word NB the "both taxa present" mask (depth != 0). That is one ballot per slice and
// pair -- 16 M rows at 1000 taxa x 1000 trees, well under a millisecond in total.
//
// k_triple_planes_bs then runs a bit-serial comparator over the slices, least significant first:
//   gt' = x_k > y_k ? 1 : (x_k == y_k ? gt : 0)      = ONE LOP3 (LUT 0xB2) for 32 trees
// so plane 0 (D(a,b) > D(a,c)) and plane 1 (the reverse) of a triple cost 2 * NB logic operations
// for 32 trees (NB = 6 for depths below 63) where the packed-fp16 builder above needs 64.
// The builder then is bound by the 6 KB it writes per (tile, group), not by instruction issue.
//
// Work split: 64 threads per (tile, group), four groups per CTA side by side (named barriers, no CTA-wide
// one). A thread owns a 2 x 2 (x, y) micro-block and two z: its four output words per (plane, z) are
// contiguous in the tile (layout.cuh tile_word), so the results leave as 128-bit streaming stores straight
// from registers -- no shared tile image. A CTA walks kPlaneTilesPerCta consecutive entries of a host-built
// work list (PlaneDesc: segments + output slot, in 16 x 16 x 16-segment blocks so that the rows a block of
// tiles shares stay in L2 while the 6 KB per (tile, group) stream out); the rows of the next tile travel by
// cp.async into the other half of a double buffer while the current tile is compared.
// Shared rows (8 words): 16-byte chunk c of row r sits at chunk (2r + c) ^ swz(r), swz from row bits 2, 4, 5,
// which makes every 128-bit access pattern of the kernel conflict-free (8 consecutive rows when staging;
// rows 2 apart and 16 / 32 apart when reading). 12-word rows (9..11 depth bits) are padded instead.
// ---------------------------------------------------------------------------------------------
// SliceRow<NB>::W (words per pair row), slice_off<W>() and SliceArray<W> (the shared row layout) live in layout.cuh.

template <int NB>
__global__ void __launch_bounds__(256) k_depth_slices(int n_batch, int npad, const uint16_t* __restrict__ Dt,
                                                      uint32_t* __restrict__ Ds) {
  constexpr int W = SliceRow<NB>::W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lo = blockIdx.x, gb = blockIdx.y;
  const size_t row0 = ((size_t)gb * npad + lo) * npad;
  for (int hi = lo + 1 + warp; hi < npad; hi += 8) {
    const uint32_t v = Dt[(row0 + hi) * kLanes + lane];
    uint32_t mine = 0;
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      const uint32_t b = __ballot_sync(0xFFFFFFFFu, (v >> k) & 1u);
      if (lane == k) mine = b;
    }
    const uint32_t pres = __ballot_sync(0xFFFFFFFFu, v != 0);
    if (lane == NB) mine = pres;
    if (lane < W) Ds[(row0 + hi) * W + lane] = mine;
  }
}

// The slices straight from the trees, without the depth table (default up to 6400 taxa, 3200 when a tree is deeper than 126): one CTA per
// (taxon a, tree group). For every tree of the group that holds a, a warp walks the leaves right and left of
// a's position with a running minimum of h (a 32-wide prefix minimum per step) and drops D(a, b) of every
// b > a into a shared [b][tree] image; the image is then sliced row by row (one ballot per depth bit) and
// written as whole rows. k_depth_table scatters 2-byte stores to 5e8 different sectors per 1000 x 1000 input
// (latency bound, 5.6 ms) and the table has to be read again by the slicer; this form writes only the slices.
// Dynamic shared memory: npad * 32 image elements -- bytes while depth + 1 fits 8 bits (NB <= 7: twice the CTAs
// per SM), u16 above.
template <int NB>
struct DepthImage {
  using T = typename std::conditional<(NB <= 8), uint8_t, uint16_t>::type;
};
template <int NB>
__global__ void __launch_bounds__(256) k_depth_slices_direct(int grp0, int n_trees, int npad, const int32_t* __restrict__ nleaf,
                                                             const uint16_t* __restrict__ lrP, const uint16_t* __restrict__ hP,
                                                             const uint16_t* __restrict__ posP, uint32_t* __restrict__ Ds) {
  constexpr int W = SliceRow<NB>::W;
  using T = typename DepthImage<NB>::T;
  extern __shared__ __align__(16) unsigned char s_dimg_raw[];
  T* const s_dimg = reinterpret_cast<T*>(s_dimg_raw);  // [b][tree of the group]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int a = blockIdx.x, gb = blockIdx.y;
  {
    uint4* z = reinterpret_cast<uint4*>(s_dimg_raw);
    for (int i = tid; i < npad * 2 * (int)sizeof(T); i += 256) z[i] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  for (int tl = warp; tl < kLanes; tl += 8) {
    const int g = (grp0 + gb) * kLanes + tl;
    if (g >= n_trees) break;
    const int pa = (int)posP[(size_t)g * npad + a] - 1;
    if (pa < 0) continue;  // a is not in this tree
    const int nl = nleaf[g];
    const uint16_t* lr = lrP + (size_t)g * npad;
    const uint16_t* h = hP + (size_t)g * npad;
    // dir 0: leaves right of a, the j-th one at position pa + 1 + j, D = min h[pa .. pa + j];
    // dir 1: leaves left of a, the j-th one at position pa - 1 - j, D = min h[pa - 1 - j .. pa - 1].
    // A lane takes four consecutive j (serial minimum), the warp scans the lane totals.
#pragma unroll
    for (int dir = 0; dir < 2; ++dir) {
      const int cnt = dir == 0 ? nl - pa - 1 : pa;
      const uint16_t* hd = dir == 0 ? h + pa : h + pa - 1;
      const uint16_t* ld = dir == 0 ? lr + pa + 1 : lr + pa - 1;
      uint32_t carry = 0xFFFFu;
      for (int base = 0; base < cnt; base += 128) {
        const int j0 = base + 4 * lane;
        uint32_t e[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) e[i] = j0 + i < cnt ? (uint32_t)(dir == 0 ? hd[j0 + i] : hd[-(j0 + i)]) : 0xFFFFu;
        e[1] = min(e[1], e[0]);
        e[2] = min(e[2], e[1]);
        e[3] = min(e[3], e[2]);
        uint32_t sc = e[3];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, sc, d);
          sc = min(sc, lane >= d ? o : 0xFFFFu);
        }
        uint32_t ex = __shfl_up_sync(0xFFFFFFFFu, sc, 1);
        ex = min(lane ? ex : 0xFFFFu, carry);
        carry = min(carry, __shfl_sync(0xFFFFFFFFu, sc, 31));
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (j0 + i < cnt) {
            const int b = dir == 0 ? ld[j0 + i] : ld[-(j0 + i)];
            if (b > a) s_dimg[b * kLanes + tl] = (T)min(e[i], ex);
          }
      }
    }
  }
  __syncthreads();
  const size_t row0 = ((size_t)gb * npad + a) * npad;
  for (int b = a + 1 + warp; b < npad; b += 8) {
    const uint32_t v = s_dimg[b * kLanes + lane];
    uint32_t mine = 0;
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      const uint32_t bits = __ballot_sync(0xFFFFFFFFu, (v >> k) & 1u);
      if (lane == k) mine = bits;
    }
    const uint32_t pres = __ballot_sync(0xFFFFFFFFu, v != 0);
    if (lane == NB) mine = pres;
    if (lane < W) Ds[(row0 + b) * W + lane] = mine;
  }
}

__device__ __forceinline__ uint32_t lop3_b2(uint32_t x, uint32_t y, uint32_t g) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xB2;" : "=r"(d) : "r"(x), "r"(y), "r"(g));  // x & ~y | ~(x ^ y) & g
  return d;
}
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
// 16-byte async copy global -> shared; bytes = 0 zero-fills (the source is then not read)
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, uint32_t bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// PlaneDesc (one entry of the host-built work list) lives in layout.cuh.
#ifndef CAMUS_PLANES_TILES
#define CAMUS_PLANES_TILES 8  // A/B builds: -DCAMUS_PLANES_TILES=n
#endif
constexpr int kPlaneTilesPerCta = CAMUS_PLANES_TILES;

// One (tile, group) by the 64 threads of a slot, rows already staged. kDerive: the third plane is the complement
// of the other two (complete binary trees, taxa of three different segments), no presence masks, bc rows unused.
template <int NB, bool kDerive>
__device__ __forceinline__ void sliced_planes_body(const uint32_t* sab, const uint32_t* sac, const uint32_t* sbc, int zq, int xp,
                                                   int yp, uint32_t vm, uint4* __restrict__ out) {
  constexpr int W = SliceRow<NB>::W, Q = W / 4;
  uint32_t ab[2][2][W];  // [dx][dy][slice]
#pragma unroll
  for (int dx = 0; dx < 2; ++dx)
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int q = 0; q < Q; ++q)
        *reinterpret_cast<uint4*>(&ab[dx][dy][4 * q]) =
            *reinterpret_cast<const uint4*>(sab + slice_off<W>((2 * xp + dx) * 8 + 2 * yp + dy, q));
#pragma unroll
  for (int zz = 0; zz < 2; ++zz) {
    const int z = 2 * zq + zz;
    uint32_t w[3][4];  // [plane][dy * 2 + dx]
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      uint32_t ac[W];
#pragma unroll
      for (int q = 0; q < Q; ++q)
        *reinterpret_cast<uint4*>(&ac[4 * q]) = *reinterpret_cast<const uint4*>(sac + slice_off<W>((2 * xp + dx) * 8 + z, q));
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        uint32_t gt = 0, lt = 0;
#pragma unroll
        for (int k = 0; k < NB; ++k) {
          gt = lop3_b2(ab[dx][dy][k], ac[k], gt);  // ab|c : D(a,b) > D(a,c)
          lt = lop3_b2(ac[k], ab[dx][dy][k], lt);  // ac|b : D(a,c) > D(a,b)
        }
        if (kDerive) {
          w[0][dy * 2 + dx] = gt;
          w[1][dy * 2 + dx] = lt;
          w[2][dy * 2 + dx] = ~(gt | lt) & vm;
        } else {
          uint32_t bc[W];
#pragma unroll
          for (int q = 0; q < Q; ++q)
            *reinterpret_cast<uint4*>(&bc[4 * q]) = *reinterpret_cast<const uint4*>(sbc + slice_off<W>((2 * yp + dy) * 8 + z, q));
          uint32_t g2 = 0;
#pragma unroll
          for (int k = 0; k < NB; ++k) g2 = lop3_b2(bc[k], ab[dx][dy][k], g2);  // bc|a : D(b,c) > D(a,b)
          const uint32_t present = ab[dx][dy][NB] & ac[NB] & bc[NB];
          w[0][dy * 2 + dx] = gt & present;
          w[1][dy * 2 + dx] = lt & present;
          w[2][dy * 2 + dx] = g2 & present;
        }
      }
    }
#pragma unroll
    for (int p = 0; p < 3; ++p) __stcs(out + ((p * 8 + z) * 4 + yp) * 4 + xp, make_uint4(w[p][0], w[p][1], w[p][2], w[p][3]));
  }
}

#ifndef CAMUS_PLANES_BS_MINB
#define CAMUS_PLANES_BS_MINB 3  // A/B builds: -DCAMUS_PLANES_BS_MINB=n
#endif
// n_batch <= 4 tree groups per launch (one per 64-thread slot). Dynamic shared memory:
// 2 buffers x 4 slots x 3 arrays x SliceArray<W>::words x 4 bytes.
template <int NB, bool kComplete>
__global__ void __launch_bounds__(256, CAMUS_PLANES_BS_MINB) k_triple_planes_bs(int grp0, int n_batch, int n_groups_total, int n_trees, int npad,
                                                                  const uint32_t* __restrict__ Ds, uint32_t* __restrict__ Tri,
                                                                  const PlaneDesc* __restrict__ list, uint32_t n_list) {
  constexpr int W = SliceRow<NB>::W, Q = W / 4, AW = SliceArray<W>::words;
  extern __shared__ __align__(16) uint32_t s_slices[];  // [buffer][slot][ab rows x*8+y | ac rows x*8+z | bc rows y*8+z][AW]
  const int tid = threadIdx.x, slot = tid >> 6, t = tid & 63;
  if (slot >= n_batch) return;  // slots only meet at their own named barrier
  const int zq = t >> 4, xp = (t >> 2) & 3, yp = t & 3;  // compute: micro-block (xp, yp), z = 2 zq + {0, 1}
  const uint32_t* const D = Ds + (size_t)slot * npad * npad * W;
  const int left = n_trees - (grp0 + slot) * kLanes;  // trees in this group
  const uint32_t vm = left >= 32 ? 0xFFFFFFFFu : (left > 0 ? (1u << left) - 1u : 0u);
  uint32_t* const Tg = Tri + (size_t)(grp0 + slot) * kTileWords;
  const uint32_t row_smem = smem_u32(s_slices) + (uint32_t)(slot * 3 * AW) * 4u;

  const uint32_t i0 = blockIdx.x * kPlaneTilesPerCta, i1 = min(n_list, i0 + kPlaneTilesPerCta);
  // staging: the 64 rows of an array are 8 runs of 8 rows (256 or 384 contiguous bytes of Ds each); consecutive
  // lanes copy consecutive 16-byte chunks (lane-per-row copies cost twice the L1 wavefronts on both sides)
  auto stage = [&](const PlaneDesc& d, int buf) {
    const uint32_t a0 = (d.s01 & 0xFFFFu) * 8, b0 = (d.s01 >> 16) * 8, c0 = d.s2 * 8;
    const bool no_bc = kComplete && a0 < b0 && b0 < c0;  // derive path: bc rows are not read
#pragma unroll
    for (int which = 0; which < 3; ++which) {
      if (which == 2 && no_bc) continue;
      const uint32_t dst = row_smem + (uint32_t)((buf * 4 * 3 + which) * AW) * 4u;
#pragma unroll
      for (int j = 0; j < Q; ++j) {
        const int c = t + 64 * j, row = c / Q, q = c % Q;  // row = first * 8 + second
        const uint32_t lo = (which == 2 ? b0 : a0) + (row >> 3);
        const uint32_t hi = (which == 0 ? b0 : c0) + (row & 7);
        const bool ok = lo < hi;  // wrong rank order = absent = zeros
        const uint32_t* src = ok ? D + ((size_t)lo * npad + hi) * W + 4 * q : D;
        cp_async16(dst + (uint32_t)slice_off<W>(row, q) * 4u, src, ok ? 16u : 0u);
      }
    }
    cp_async_commit();
  };
  PlaneDesc cur = list[i0], nxt = cur;
  if (i0 + 1 < i1) nxt = list[i0 + 1];
  stage(cur, 0);
  int buf = 0;
  for (uint32_t i = i0; i < i1; ++i) {
    cp_async_wait_all();
    named_bar_sync(1 + slot, 64);  // tile i has landed for all 64 threads, and all of them are done with tile i - 1
    PlaneDesc nn = nxt;
    if (i + 1 < i1) {
      stage(nxt, buf ^ 1);
      if (i + 2 < i1) nn = list[i + 2];
    }
    const uint32_t s0 = cur.s01 & 0xFFFFu, s1 = cur.s01 >> 16;
    const bool derive = kComplete && s0 < s1 && s1 < cur.s2;  // every (x, y, z) is in rank order
    const uint32_t* rows = s_slices + (size_t)((buf * 4 + slot) * 3) * AW;
    uint4* const out = reinterpret_cast<uint4*>(Tg + (size_t)cur.slot * n_groups_total * kTileWords);
    if (derive) sliced_planes_body<NB, true>(rows, rows + AW, rows + 2 * AW, zq, xp, yp, vm, out);
    else sliced_planes_body<NB, false>(rows, rows + AW, rows + 2 * AW, zq, xp, yp, vm, out);
    cur = nxt;
    nxt = nn;
    buf ^= 1;
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
