// K2: quartet topology counting, bit-sliced over gene trees, with K3 (filter + constraint
// removal) and K4 (edge-score scatter) fused into its epilogue.
//
// Replaces graphs.QuartetsFromTree + the per-tree set + the 64 mutex-sharded maps of
// prep.processQuartets (internal/graphs/quartet.go:112-123, internal/prep/preprocess.go:71-124):
// count[cell][topology] = number of gene trees displaying that topology on that 4-taxon set.
//
// For a < b < c < d (DFS ranks) and R[xy|z] the rooted-triple bit-planes of K1b:
//   ab|cd  <=>  (R[ab|c] & R[ab|d]) | (R[cd|a] & R[cd|b])
//   ac|bd  <=>  (R[ac|b] & R[ac|d]) | (R[bd|a] & R[bd|c])
//   ad|bc  <=>  (R[ad|b] & R[ad|c]) | (R[bc|a] & R[bc|d])
// (a clade holding exactly one pair exists iff the unrooted tree has the split; SURVEY.md
// Appendix B). One 32-bit word = 32 gene trees, so a cell x 32 trees costs 6 LOP3 + 3 POPC + 3 IADD.
//
// One CTA = one 8x8x8x8 box of cells = 4 triple tiles (abc, abd, acd, bcd) per tree group,
// streamed HBM -> shared memory by the TMA engine (cp.async.bulk, 2 KB planes, 4-stage mbarrier
// ring). 256 threads, each owning a 2x2x2x2 micro-box (16 cells, counters in registers) so a
// 128-bit shared load feeds four cells.
//
// Measured-and-rejected variants are listed in profiles/r01_count_variants.md.
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

#include "kernels_score.cuh"
#include "layout.cuh"
#include "ptx.cuh"

namespace camus_dev {

constexpr int kCountThreads = 256;
constexpr int kPlaneWords = kTileWords / 3;  // 512 words = 2 KB: one plane of one tile
#ifndef CAMUS_COUNT_OCC4
#define CAMUS_COUNT_OCC4 1  // 0 = A/B build of the round-1 shape of the complete-tree loop (3 CTAs per SM, 4-stage ring)
#endif
#ifndef CAMUS_COUNT_LANEMAP
#define CAMUS_COUNT_LANEMAP 1  // 0 = A/B build of the round-1 lane order (ta ta tb tb tc)
#endif
#ifndef CAMUS_COUNT_EXPERIMENT
#define CAMUS_COUNT_EXPERIMENT 0  // timing experiments only (wrong results): 1 = no POPC (XU pipe idle), 2 = operands loaded once per box (LSU idle),
                                  // 3 = as 2 and no barrier, no tile copies, no mbarrier wait in the group loop (the bare compute loop)
#endif
#if CAMUS_COUNT_EXPERIMENT == 1
#define CAMUS_POPC(x) (x)
#else
#define CAMUS_POPC(x) __popc(x)
#endif
#ifndef CAMUS_COUNT_NOBAR
#define CAMUS_COUNT_NOBAR 0  // A/B build: no CTA barrier in the group loop, the last warp out of a stage refills it
#endif

__device__ __forceinline__ uint32_t atom_add_shared_relaxed(uint32_t* p, uint32_t v) {
  uint32_t old;
  asm volatile("atom.relaxed.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
  return old;
}

// Counting modes (template parameter kMode):
//   kModeGeneral  : any gene trees, any number of them. Three 32-bit counters per cell, 12 planes
//                   per stage, 128 registers, 2 CTAs per SM.
//   kModeComplete : every gene tree holds every taxon and is fully resolved (host-checked in
//                   camus_b200_add_gene_trees), at most 65535 trees. Each tree then resolves each
//                   cell, so count(ad|bc) = G - count(ab|cd) - count(ac|bd): the third topology
//                   word, its POPC and the four planes only it reads are skipped (8 planes = 16 KB
//                   per stage) and the two counters share one register (16 + 16 bits).
//   kModePacked   : incomplete / polytomous trees, at most 1023 of them: the three counters share
//                   one register (10 + 10 + 10 bits), which brings the general kernel from 128 to
//                   <= 80 registers and from 2 to 3 CTAs per SM.
// The loop is bound by latency (XU = POPC pipe ~65 % busy, issue ~60 %), not by a pipe, so
// resident warps matter: complete and packed fetch the acd / bcd operands per (c, d) pair instead
// of holding them across both, and run a 3-stage ring, to fit 4 resp. 3 CTAs per SM
// (profiles/r02_count_variants.md).
enum : int { kModeGeneral = 0, kModeComplete = 1, kModePacked = 2 };

template <int kMode>
struct CountCfg {
  static constexpr bool kComplete = kMode == kModeComplete;
  static constexpr bool kPacked = kMode == kModePacked;
  // narrow = operands fetched per (c, d) pair + 3-stage ring with one CTA barrier per group
#if CAMUS_COUNT_OCC4
  static constexpr bool kNarrow = kComplete || kPacked;
#else
  static constexpr bool kNarrow = kPacked;
#endif
  static constexpr int kStages = kNarrow ? 3 : 4;
  static constexpr int kPlanes = kComplete ? 8 : 12;
  static constexpr int kStageWords = kPlanes * kPlaneWords;
  static constexpr int kStageBytes = kStageWords * 4;
  static constexpr int kSmemBytes = kStages * kStageBytes + 64;  // + mbarriers
  static constexpr int kMinBlocks = kComplete ? (kNarrow ? 4 : 3) : (kPacked ? 3 : 2);
  static constexpr int kCnt = kMode == kModeGeneral ? 3 : 1;  // counter registers per cell
};

struct CountParams {
  const uint32_t* tri;  // [tile][group][kTileWords]
  int n_groups;
  BoxMap bm;       // blockIdx.x -> global box rank
  uint4* cells;    // table mode: [local box][kBoxCells] = (c(ab|cd), c(ac|bd), c(ad|bc), 0)
  int n_trees;
  const uint32_t* tile_slot;  // class-scheme shares store only the tiles they read: tile index -> slot in tri; nullptr = identity
  const uint2* desc;          // per local box: segments + regular flag (k_box_desc); nullptr = unrank in the kernel
};

// Box descriptor: x = sa | sb << 16, y = sc | sd << 16 | regular << 31 (segments < 4096). Unranking a
// box costs thread 0 two roots, a dozen 64-bit multiplications and its fix-up loops while 255
// threads wait at the first barrier (8 % of the warp samples of the round-2 profile); the table is
// built once per (constraint, partition) and costs 8 B per box.
__global__ void k_box_desc(BoxMap bm, uint64_t n_boxes, ConstraintDev ct, int regular, uint2* out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_boxes) return;
  uint32_t sa, sb, sc, sd;
  box_unrank(bm.global(i), sa, sb, sc, sd);
  const uint8_t* U = ct.seg_uniform;
  const uint32_t S = (uint32_t)ct.n_seg;
  const bool reg = regular && sa < sb && sb < sc && sc < sd && U[sa * S + sb] && U[sb * S + sc] && U[sc * S + sd];
  out[i] = make_uint2(sa | (sb << 16), sc | (sd << 16) | (reg ? 0x80000000u : 0u));
}

// Fused mode: the counts never leave the registers. The epilogue applies K3 (filter + constraint
// removal + scalars) and K4 (edge-score scatter into Z) per cell; no cell table exists in HBM.
struct FusedParams {
  int qmode;
  double threshold;
  const uint32_t* thr_tab;     // k_threshold_table
  int as_set;
  unsigned long long* Z;       // [z_rep][n][n][3]
  unsigned long long* totals;  // [0] n_survivors, [1] sum_counts
  int z_rep;                   // number of privatised copies of Z (hot-address spreading), stride z_stride
  unsigned long long z_stride;
  int regular;                 // 1 = regular boxes take the marginal-sum epilogue (CAMUS_B200_NO_REGULAR=1 clears it)
  int debug;                   // profiling only (CAMUS_B200_DEBUG): 1 = skip the score scatter, 2 = skip the whole epilogue,
                               // 3 = as 2 and no tile traffic after the first four groups
};

__device__ __forceinline__ uint32_t pick(const uint4& v, int hi, int lo) {
  int k = hi * 2 + lo;
  return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w));
}

// Plane slot inside a stage. General layout: [role abc,abd,acd,bcd][plane 0..2]. Complete layout
// keeps only the planes ab|cd and ac|bd read: abc{0,1} abd{0,2} acd{0,2} bcd{1,2}.
template <bool kComplete>
__device__ __forceinline__ constexpr int plane_slot(int role, int plane) {
  if (!kComplete) return role * 3 + plane;
  return role == 0 ? plane : (role == 1 ? 2 + (plane >> 1) : (role == 2 ? 4 + (plane >> 1) : 6 + (plane - 1)));
}

template <bool kFused, int kMode>
__global__ void __launch_bounds__(kCountThreads, CountCfg<kMode>::kMinBlocks)
k_count(CountParams prm, FusedParams fz, ConstraintDev ct) {
  using Cfg = CountCfg<kMode>;
  constexpr bool kComplete = Cfg::kComplete, kPacked = Cfg::kPacked;
  extern __shared__ __align__(128) uint32_t smem[];
  constexpr int kCountStages = Cfg::kStages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kCountStages * Cfg::kStageWords);
  __shared__ uint32_t s_seg[5];  // four segments, [4] = regular box
  __shared__ uint32_t s_done[4];  // CAMUS_COUNT_NOBAR: warps that have finished with each ring stage
  __shared__ uint64_t s_tile[4];  // slots of the tiles abc, abd, acd, bcd in prm.tri
  // packed constraint LCAs of the box's 8 x 8 (a, b), (b, c) and (c, d) taxon pairs: everything the epilogue reads from
  // ct.lcaD. Loaded at the start (the latency hides behind the count loop); the epilogue was stalled on these loads
  // (long scoreboard = 23 % of its samples).
  __shared__ uint32_t s_lca[kFused ? 3 : 1][64];

  const int tid = threadIdx.x;
  // Lane order. An LDS.128 costs 4 shared-memory wavefronts, or 2 when the four lanes of every
  // aligned quad read at most two addresses (measured, tools/lds_probe.cu: the address must not
  // depend on lane bit 0 or on lane bit 1). With lane bits (a0 b0 a1 b1 c0) the bcd operand ignores
  // bit 0 and the acd operand bit 1: 48 wavefronts per thread and tree group instead of 56.
#if CAMUS_COUNT_LANEMAP
  const int ta = (tid & 1) | ((tid >> 1) & 2), tb = ((tid >> 1) & 1) | ((tid >> 2) & 2), tc = (tid >> 4) & 3, td = tid >> 6;
#else
  const int ta = tid & 3, tb = (tid >> 2) & 3, tc = (tid >> 4) & 3, td = tid >> 6;
#endif

  if (tid == 0) {
    uint32_t sa, sb, sc, sd;
    bool reg;
    if (prm.desc) {
      const uint2 d = __ldg(prm.desc + prm.bm.first + blockIdx.x);
      sa = d.x & 0xFFFFu;
      sb = d.x >> 16;
      sc = d.y & 0xFFFFu;
      sd = (d.y >> 16) & 0x7FFFu;
      reg = (d.y >> 31) != 0;
    } else {
      box_unrank(prm.bm.global(blockIdx.x), sa, sb, sc, sd);
      reg = false;
      if constexpr (kFused) {
        const uint8_t* U = ct.seg_uniform;
        const uint32_t S = (uint32_t)ct.n_seg;
        reg = fz.regular && sa < sb && sb < sc && sc < sd && U[sa * S + sb] && U[sb * S + sc] && U[sc * S + sd];
      }
    }
    s_seg[0] = sa;
    s_seg[1] = sb;
    s_seg[2] = sc;
    s_seg[3] = sd;
    s_seg[4] = reg;
    auto slot = [&](uint32_t s0, uint32_t s1, uint32_t s2) {
      const uint64_t t = tile_index(s0, s1, s2);
      return prm.tile_slot ? (uint64_t)__ldg(prm.tile_slot + t) : t;
    };
    s_tile[0] = slot(sa, sb, sc);
    s_tile[1] = slot(sa, sb, sd);
    s_tile[2] = slot(sa, sc, sd);
    s_tile[3] = slot(sb, sc, sd);
    for (int s = 0; s < kCountStages; ++s) {
      mbar_init(&bars[s], 1);
      s_done[s] = 0;
    }
    fence_mbar_init();
  }
  __syncthreads();

  const int G = prm.n_groups;
  const uint32_t sa = s_seg[0], sb = s_seg[1], sc = s_seg[2], sd = s_seg[3];
  const size_t gstride = (size_t)kTileWords;
  const uint32_t* src0 = prm.tri + s_tile[0] * (size_t)G * gstride;  // abc
  const uint32_t* src1 = prm.tri + s_tile[1] * (size_t)G * gstride;  // abd
  const uint32_t* src2 = prm.tri + s_tile[2] * (size_t)G * gstride;  // acd
  const uint32_t* src3 = prm.tri + s_tile[3] * (size_t)G * gstride;  // bcd

  auto issue = [&](int g) {
    const int s = g % kCountStages;
    uint32_t* dst = smem + s * Cfg::kStageWords;
    constexpr int P = kPlaneWords, PB = kPlaneWords * 4;
    const size_t o = (size_t)g * gstride;
    mbar_arrive_expect_tx(&bars[s], Cfg::kStageBytes);
    if constexpr (!kComplete) {
      bulk_g2s(dst, src0 + o, 3 * PB, &bars[s]);
      bulk_g2s(dst + 3 * P, src1 + o, 3 * PB, &bars[s]);
      bulk_g2s(dst + 6 * P, src2 + o, 3 * PB, &bars[s]);
      bulk_g2s(dst + 9 * P, src3 + o, 3 * PB, &bars[s]);
    } else {
      bulk_g2s(dst, src0 + o, 2 * PB, &bars[s]);                   // abc planes 0,1
      bulk_g2s(dst + 2 * P, src1 + o, PB, &bars[s]);               // abd plane 0
      bulk_g2s(dst + 3 * P, src1 + o + 2 * P, PB, &bars[s]);       // abd plane 2
      bulk_g2s(dst + 4 * P, src2 + o, PB, &bars[s]);               // acd plane 0
      bulk_g2s(dst + 5 * P, src2 + o + 2 * P, PB, &bars[s]);       // acd plane 2
      bulk_g2s(dst + 6 * P, src3 + o + P, 2 * PB, &bars[s]);       // bcd planes 1,2
    }
  };

  if (tid == 0)
    for (int g = 0; g < Cfg::kStages && g < G; ++g) issue(g);
  uint32_t lca_pre = 0;
  if constexpr (kFused) {
    if (tid < 192) {
      const uint32_t which = tid >> 6, u = (tid >> 3) & 7, v = tid & 7, Nt = (uint32_t)ct.n_taxa;
      const uint32_t lo = (which == 0 ? sa : (which == 1 ? sb : sc)) * 8 + u, hi = (which == 0 ? sb : (which == 1 ? sc : sd)) * 8 + v;
      lca_pre = ct.lcaD[(size_t)min(lo, Nt - 1) * Nt + min(hi, Nt - 1)];
    }
  }

  // general: cnt[c][t] = trees displaying topology t. complete: cnt[c][0] = c(ab|cd) | c(ac|bd) << 16.
  // packed: cnt[c][0] = c(ab|cd) | c(ac|bd) << 10 | c(ad|bc) << 20.
  constexpr int kCnt = Cfg::kCnt;
  uint32_t cnt[16][kCnt];
#pragma unroll
  for (int c = 0; c < 16; ++c)
#pragma unroll
    for (int t = 0; t < kCnt; ++t) cnt[c][t] = 0;

  // per-thread uint4 offsets inside a plane, (z * 4 + ty) * 4 + tx, hoisted out of the group loop
  // (made opaque to the compiler, which otherwise re-derives them from tid in every iteration)
  uint32_t o_abc = ((2 * tc) * 4 + tb) * 4 + ta;  // + 16 per kk
  uint32_t o_abd = ((2 * td) * 4 + tb) * 4 + ta;  // + 16 per ll
  uint32_t o_acd = ((2 * td) * 4 + tc) * 4 + ta;
  uint32_t o_bcd = ((2 * td) * 4 + tc) * 4 + tb;
  asm volatile("" : "+r"(o_abc), "+r"(o_abd), "+r"(o_acd), "+r"(o_bcd));
  constexpr int kPlaneU4 = kPlaneWords / 4;

  if constexpr (kFused) {
    if (G > 0) mbar_wait(&bars[0], 0u);  // the first tiles are in: by now the LCA loads have landed as well
    if (tid < 192) s_lca[tid >> 6][tid & 63] = lca_pre;
  }
  for (int g = 0; g < G; ++g) {
    const int s = g % kCountStages;
    if constexpr (CAMUS_COUNT_NOBAR || CAMUS_COUNT_EXPERIMENT == 3) {
      // nothing here: see the end of the loop body
    } else if constexpr (Cfg::kNarrow) {
      // 3-stage ring: everyone has finished group g-1, its stage is refilled with g+2
      if (g > 0) {
        __syncthreads();
        if (tid == 0 && g + 2 < G && !(kFused && fz.debug == 3)) issue(g + 2);
      }
    } else {
      // One CTA barrier per TWO groups: at even g everyone has finished groups g-2 and g-1, whose
      // stages are refilled with g+2 and g+3 (the ring holds g .. g+3 in flight or ready).
      if ((g & 1) == 0 && g > 0) {
        __syncthreads();
        if (tid == 0 && !(kFused && fz.debug == 3)) {
          if (g + 2 < G) issue(g + 2);
          if (g + 3 < G) issue(g + 3);
        }
      }
    }
    // debug 3 (timing only): the ring is filled once and re-read, i.e. the loop without tile traffic
    if (!(kFused && fz.debug == 3 && g >= kCountStages) && !(CAMUS_COUNT_EXPERIMENT == 3 && g >= kCountStages)) mbar_wait(&bars[s], (uint32_t)((g / kCountStages) & 1));

#if CAMUS_COUNT_EXPERIMENT >= 2
    const uint4* st = reinterpret_cast<const uint4*>(smem) + ((g & 1) ? 0 : 0);  // loop invariant: the loads are hoisted out of the group loop
#else
    const uint4* st = reinterpret_cast<const uint4*>(smem + s * Cfg::kStageWords);
#endif
    auto ld = [&](int role, int plane, uint32_t off) -> uint4 { return st[plane_slot<kComplete>(role, plane) * kPlaneU4 + off]; };

    uint4 abc[3][2];
#pragma unroll
    for (int p = 0; p < (kComplete ? 2 : 3); ++p)
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) abc[p][kk] = ld(0, p, o_abc + 16 * kk);

    if constexpr (Cfg::kNarrow) {
      // 64-bit halves of the acd / bcd words: the pair (c = kk, d = ll) fixes the half
      auto ld2 = [&](int role, int plane, uint32_t off, int kk) -> uint2 {
        return reinterpret_cast<const uint2*>(st + plane_slot<kComplete>(role, plane) * kPlaneU4 + off)[kk];
      };
#pragma unroll
      for (int ll = 0; ll < 2; ++ll) {
        uint4 abd[3];
#pragma unroll
        for (int p = 0; p < 3; ++p)
          if (!kComplete || p != 1) abd[p] = ld(1, p, o_abd + 16 * ll);
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          uint2 acd[3], bcd[3];
#pragma unroll
          for (int p = 0; p < 3; ++p) {
            if (!kComplete || p != 1) acd[p] = ld2(2, p, o_acd + 16 * ll, kk);
            if (!kComplete || p != 0) bcd[p] = ld2(3, p, o_bcd + 16 * ll, kk);
          }
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int c = ((ll * 2 + kk) * 2 + j) * 2 + i;
              const uint32_t t0 = (pick(abc[0][kk], j, i) & pick(abd[0], j, i)) | ((i ? acd[2].y : acd[2].x) & (j ? bcd[2].y : bcd[2].x));  // ab|cd
              const uint32_t t1 = (pick(abc[1][kk], j, i) & (i ? acd[0].y : acd[0].x)) | (pick(abd[2], j, i) & (j ? bcd[1].y : bcd[1].x));  // ac|bd
              if constexpr (kComplete) {
                cnt[c][0] += CAMUS_POPC(t0) + (CAMUS_POPC(t1) << 16);
              } else {
                const uint32_t t2 = (pick(abd[1], j, i) & (i ? acd[1].y : acd[1].x)) | (pick(abc[2][kk], j, i) & (j ? bcd[0].y : bcd[0].x));  // ad|bc
                cnt[c][0] += CAMUS_POPC(t0) + (CAMUS_POPC(t1) << 10) + (CAMUS_POPC(t2) << 20);
              }
            }
        }
      }
    } else {
#pragma unroll
    for (int ll = 0; ll < 2; ++ll) {
      uint4 abd[3], acd[3], bcd[3];
#pragma unroll
      for (int p = 0; p < 3; ++p) {
        if (!kComplete || p != 1) abd[p] = ld(1, p, o_abd + 16 * ll);
        if (!kComplete || p != 1) acd[p] = ld(2, p, o_acd + 16 * ll);
        if (!kComplete || p != 0) bcd[p] = ld(3, p, o_bcd + 16 * ll);
      }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int c = ((ll * 2 + kk) * 2 + j) * 2 + i;
            const uint32_t t0 = (pick(abc[0][kk], j, i) & pick(abd[0], j, i)) | (pick(acd[2], kk, i) & pick(bcd[2], kk, j));  // ab|cd
            const uint32_t t1 = (pick(abc[1][kk], j, i) & pick(acd[0], kk, i)) | (pick(abd[2], j, i) & pick(bcd[1], kk, j));  // ac|bd
            if constexpr (kComplete) {
              cnt[c][0] += __popc(t0) + (__popc(t1) << 16);
            } else {
              const uint32_t t2 = (pick(abd[1], j, i) & pick(acd[1], kk, i)) | (pick(abc[2][kk], j, i) & pick(bcd[0], kk, j));  // ad|bc
              cnt[c][0] += __popc(t0);
              cnt[c][1] += __popc(t1);
              cnt[c][2] += __popc(t2);
            }
          }
    }
    }
    if constexpr (CAMUS_COUNT_NOBAR) {
      // release the stage: shared loads of a warp are performed in order ahead of its atomic, and the
      // last of the 8 warps to get here refills the stage with group g + stages
      __syncwarp();
      if ((tid & 31) == 0) {
        const uint32_t old = atom_add_shared_relaxed(&s_done[s], 1u);
        if ((old & (kCountThreads / 32 - 1)) == kCountThreads / 32 - 1 && g + kCountStages < G) issue(g + kCountStages);
      }
    }
  }

  // counts of the three topologies of cell ci
  auto counts_of = [&](int ci, uint32_t c[3]) {
    if constexpr (kComplete) {
      c[0] = cnt[ci][0] & 0xFFFFu;
      c[1] = cnt[ci][0] >> 16;
      c[2] = (uint32_t)prm.n_trees - c[0] - c[1];
    } else if constexpr (kPacked) {
      c[0] = cnt[ci][0] & 0x3FFu;
      c[1] = (cnt[ci][0] >> 10) & 0x3FFu;
      c[2] = cnt[ci][0] >> 20;
    } else {
      c[0] = cnt[ci][0];
      c[1] = cnt[ci][1];
      c[2] = cnt[ci][2];
    }
  };

  if constexpr (!kFused) {
    uint4* out = prm.cells + (size_t)blockIdx.x * kBoxCells;
#pragma unroll
    for (int ll = 0; ll < 2; ++ll)
#pragma unroll
      for (int kk = 0; kk < 2; ++kk)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            uint32_t c[3];
            counts_of(((ll * 2 + kk) * 2 + j) * 2 + i, c);
            out[cell_local(2 * ta + i, 2 * tb + j, 2 * tc + kk, 2 * td + ll)] = make_uint4(c[0], c[1], c[2], 0u);
          }
  } else {
    if (fz.debug >= 2) {  // timing decomposition only: the main loop alone (results are wrong on purpose)
      uint32_t keep = 0;
#pragma unroll
      for (int c = 0; c < 16; ++c)
#pragma unroll
        for (int t = 0; t < kCnt; ++t) keep ^= cnt[c][t];
      if (keep == 0xFFFFFFFFu) fz.totals[0] = keep;  // keeps the counters alive
      return;
    }
    // ---- fused K3 + K4 on the register-resident counts ----
    // Round-1 profile: 27 % of the kernel's instructions were in this epilogue (~140 per cell, a
    // dozen branches each). It is now one branch-free pass per cell (~45 instructions):
    //   X = count of ac|bd, Y = count of the one of ab|cd / ad|bc the constraint does not display,
    //   Z = count of the one it does (never a survivor); lo <= mid <= hi their sorted values;
    //   keep = Keep(lo, mid);  X survives iff keep ? (mode 1 or X != lo) : (X == hi), Y likewise.
    // The one case this cannot decide -- Keep fails and the two largest counts tie, where the
    // reference's stable topology order picks the survivor -- is deferred: the cell goes to the
    // warp's queue with its raw counts and is resolved there by apply_filter.
    // Regular boxes (the three consecutive segment pairs have one LCA each, ConstraintDev::
    // seg_uniform; half of all boxes of a random 1000-taxon constraint): junction, m-node and
    // component of a (bottom taxon, topology) are box constants, so the box contributes 4 axes x
    // 2 topologies x 8 marginal sums of the surviving weights: 72 reductions per box.
    // Other boxes: survivors are compacted into a queue PER WARP (its eighth of the idle tile
    // ring: no shared atomics, no CTA barrier) and scored 32 at a time.
    __shared__ unsigned long long s_tot[2];
    __shared__ uint32_t s_marg[4][2][8];  // [bottom axis a..d][topology slot][coordinate]
    __shared__ uint32_t s_wq[kCountThreads / 32];  // queue fill of the warps of a regular box (deferred ties only)
    if (tid < 2) s_tot[tid] = 0;
    if (tid < 64) (&s_marg[0][0][0])[tid] = 0;
    if (tid < kCountThreads / 32) s_wq[tid] = 0;
    const bool regular = s_seg[4] != 0;
    __syncthreads();  // also: every issued bulk copy has been waited for, the ring can be reused
    constexpr int kWarps = kCountThreads / 32;
    // queue entry = cell (12 bits) | deferred-tie flag | three counts: 16 bytes, or 8 for complete
    // / packed mode (counts <= 65535), so that the 512 cells of a warp fit its eighth of even a 48 KB ring
    constexpr bool kSmallCounts = kComplete || kPacked;
    using QEntry = typename std::conditional<kSmallCounts, uint2, uint4>::type;
    static_assert(kCountStages * Cfg::kStageBytes >= kBoxCells * (int)sizeof(QEntry), "the per-warp queues must fit the ring");
    const int lane = tid & 31, warp = tid >> 5;
    QEntry* queue = reinterpret_cast<QEntry*>(smem) + warp * (kBoxCells / kWarps);  // 512 entries: every cell of the warp
    auto q_pack = [](uint32_t local, bool tie, uint32_t k0, uint32_t k1, uint32_t k2) -> QEntry {
      if constexpr (kSmallCounts) return make_uint2(local | (tie ? 0x1000u : 0u) | (k0 << 13) | (k1 << 29), (k1 >> 3) | (k2 << 13));
      else return make_uint4(local | (tie ? 0x1000u : 0u), k0, k1, k2);
    };
    auto q_unpack = [](const QEntry& q, uint32_t& local, bool& tie, uint32_t c[3]) {
      local = q.x & 0xFFFu;
      tie = (q.x & 0x1000u) != 0;
      if constexpr (kSmallCounts) {
        c[0] = (q.x >> 13) & 0xFFFFu;
        c[1] = ((q.x >> 29) | (q.y << 3)) & 0xFFFFu;
        c[2] = (q.y >> 13) & 0xFFFFu;
      } else {
        c[0] = q.y; c[1] = q.z; c[2] = q.w;
      }
    };
    const uint32_t N = (uint32_t)ct.n_taxa;
    const uint32_t a0 = sa * 8 + 2 * ta, b0 = sb * 8 + 2 * tb, c0 = sc * 8 + 2 * tc, d0 = sd * 8 + 2 * td;
    const uint32_t Gt = (uint32_t)prm.n_trees;
    const int qmode = fz.qmode;
    const uint32_t* __restrict__ thr_tab = fz.thr_tab;
    uint32_t ns = 0, sum = 0;  // per thread: at most 48 keys, 48 x trees counts (the host keeps trees < 2^24 here)
    uint32_t qn = 0;           // warp-uniform queue fill (irregular boxes)
    // packed LCAs of the box's first taxa: THE LCAs of a regular box
    const uint32_t bg1 = s_lca[0][0], bg2 = s_lca[1][0], bg3 = s_lca[2][0];
    const bool box_kill0 = constraint_topology(bg1, bg2, bg3) == 0;  // regular box: the constraint displays ab|cd (else ad|bc)

    // one cell: (wx, wy) = surviving counts of ac|bd and of the other non-constraint topology; tie = deferred
    auto decide = [&](const uint32_t c[3], bool kill0, uint32_t& wx, uint32_t& wy, bool& tie) {
      const uint32_t X = c[1], Y = kill0 ? c[2] : c[0], Z = kill0 ? c[0] : c[2];
      tie = false;
      if (qmode == 0) {
        wx = X;
        wy = Y;
        return;
      }
      const uint32_t lo = min(X, min(Y, Z)), hi = max(X, max(Y, Z));
      const uint32_t mid = (kComplete ? Gt : X + Y + Z) - lo - hi;
      const bool keep = __ldg(&thr_tab[lo + mid]) < mid - lo;  // Threshold.Keep (quartetfilter.go:67-74)
      const bool drop_lo = qmode == 2;
      const bool sx = keep ? (!drop_lo || X != lo) : (X == hi);
      const bool sy = keep ? (!drop_lo || Y != lo) : (Y == hi);
      tie = !keep && mid == hi && hi != 0;
      wx = sx && !tie ? X : 0u;
      wy = sy && !tie ? Y : 0u;
    };

    if (regular) {
      uint32_t ma[2][2] = {{0, 0}, {0, 0}}, mb[2][2] = {{0, 0}, {0, 0}}, mc[2][2] = {{0, 0}, {0, 0}}, md[2][2] = {{0, 0}, {0, 0}};
#pragma unroll
      for (int ll = 0; ll < 2; ++ll)
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              uint32_t c[3], w0, w1;
              bool tie;
              counts_of(((ll * 2 + kk) * 2 + j) * 2 + i, c);
              decide(c, box_kill0, w0, w1, tie);
              if (tie) {  // rare: resolved in the queue pass below
                const uint32_t at = atomicAdd(&s_wq[warp], 1u);
                queue[at] = q_pack(cell_local(2 * ta + i, 2 * tb + j, 2 * tc + kk, 2 * td + ll), true, c[0], c[1], c[2]);
              }
              ns += (w0 != 0) + (w1 != 0);
              sum += w0 + w1;
              if (fz.as_set) {
                w0 = w0 != 0;
                w1 = w1 != 0;
              }
              ma[0][i] += w0; mb[0][j] += w0; mc[0][kk] += w0; md[0][ll] += w0;
              ma[1][i] += w1; mb[1][j] += w1; mc[1][kk] += w1; md[1][ll] += w1;
            }
      // warp sums over the lanes that share the coordinate, one shared atomic per (warp, sum)
#if CAMUS_COUNT_LANEMAP
      const uint32_t mask_a = 0x05050505u << ((ta & 1) + 4 * (ta >> 1)), mask_b = 0x00330033u << (2 * (tb & 1) + 8 * (tb >> 1));
      const bool lead_a = (lane & 0x1A) == 0, lead_b = (lane & 0x15) == 0;
#else
      const uint32_t mask_a = 0x11111111u << ta, mask_b = 0x000F000Fu << (4 * tb);
      const bool lead_a = lane < 4, lead_b = lane == 4 * tb;
#endif
      const uint32_t mask_c = lane < 16 ? 0x0000FFFFu : 0xFFFF0000u;
#pragma unroll
      for (int sl = 0; sl < 2; ++sl)
#pragma unroll
        for (int par = 0; par < 2; ++par) {
          uint32_t v = __reduce_add_sync(mask_a, ma[sl][par]);
          if (lead_a && v) atomicAdd(&s_marg[0][sl][2 * ta + par], v);
          v = __reduce_add_sync(mask_b, mb[sl][par]);
          if (lead_b && v) atomicAdd(&s_marg[1][sl][2 * tb + par], v);
          v = __reduce_add_sync(mask_c, mc[sl][par]);
          if ((lane & 15) == 0 && v) atomicAdd(&s_marg[2][sl][2 * tc + par], v);
          v = __reduce_add_sync(0xFFFFFFFFu, md[sl][par]);
          if (lane == 0 && v) atomicAdd(&s_marg[3][sl][2 * td + par], v);
        }
      __syncwarp();
      qn = s_wq[warp];
    } else {
      // LCAs of the thread's 2x2 (a,b), (b,c), (c,d) pairs, loaded once (was: three loads per live cell)
      uint32_t g1[2][2], g2[2][2], g3[2][2];
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          g1[u][v] = s_lca[0][(2 * ta + u) * 8 + 2 * tb + v];
          g2[u][v] = s_lca[1][(2 * tb + u) * 8 + 2 * tc + v];
          g3[u][v] = s_lca[2][(2 * tc + u) * 8 + 2 * td + v];
        }
#pragma unroll
      for (int ll = 0; ll < 2; ++ll)
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              uint32_t c[3], wx, wy;
              bool tie;
              counts_of(((ll * 2 + kk) * 2 + j) * 2 + i, c);
              const bool valid = a0 + i < b0 + j && b0 + j < c0 + kk && c0 + kk < d0 + ll && d0 + ll < N;
              if (!valid) c[0] = c[1] = c[2] = 0;  // diagonal / padded cells hold garbage (and would index the Keep table with it)
              const bool kill0 = constraint_topology(g1[i][j], g2[j][kk], g3[kk][ll]) == 0;
              decide(c, kill0, wx, wy, tie);
              tie = tie && valid;
              const bool live = valid && ((wx | wy) != 0 || tie);
              const uint32_t bal = __ballot_sync(0xFFFFFFFFu, live);
              if (bal) {  // warp-uniform
                if (live) {
                  const uint32_t local = cell_local(2 * ta + i, 2 * tb + j, 2 * tc + kk, 2 * td + ll);
                  const uint32_t at = qn + __popc(bal & ((1u << lane) - 1));
                  // decided cells carry their surviving counts, deferred ties the raw ones (+ flag)
                  queue[at] = tie ? q_pack(local, true, c[0], c[1], c[2]) : q_pack(local, false, kill0 ? 0u : wy, wx, kill0 ? wy : 0u);
                }
                qn += __popc(bal);
              }
            }
      __syncwarp();
    }

    // queue pass (one copy of this code): survivors of an irregular box, deferred ties of any box
    if (qn && fz.debug != 1) {
      unsigned long long* Zrep = fz.Z + (size_t)(blockIdx.x % (unsigned)fz.z_rep) * fz.z_stride;
      GlobalAcc acc{Zrep};
#pragma unroll 1
      for (uint32_t e = lane; e < qn; e += 32) {
        uint32_t local, c[3];
        bool was_tie;
        q_unpack(queue[e], local, was_tie, c);
        const uint32_t x[4] = {sa * 8 + (local & 7), sb * 8 + ((local >> 3) & 7), sc * 8 + ((local >> 6) & 7), sd * 8 + (local >> 9)};
        const uint32_t h1 = s_lca[0][(local & 7) * 8 + ((local >> 3) & 7)];
        const uint32_t h2 = s_lca[1][((local >> 3) & 7) * 8 + ((local >> 6) & 7)];
        const uint32_t h3 = s_lca[2][((local >> 6) & 7) * 8 + (local >> 9)];
        const int kc = constraint_topology(h1, h2, h3);
        if (was_tie) {  // deferred tie: the reference's stable topology order decides (apply_filter)
          const int ti[4] = {ct.tip_of_rank[x[0]], ct.tip_of_rank[x[1]], ct.tip_of_rank[x[2]], ct.tip_of_rank[x[3]]};
          apply_filter(c, ti, qmode, fz.threshold);
        }
        if (kc == 0) c[0] = 0;
        else c[2] = 0;
        if (!regular || was_tie) {
          ns += (c[0] != 0) + (c[1] != 0) + (c[2] != 0);
          sum += c[0] + c[1] + c[2];
        }
        if (regular) {  // a deferred tie of a regular box joins the marginal sums
          const uint32_t w0 = fz.as_set ? (c[1] != 0) : c[1];
          const uint32_t cy = box_kill0 ? c[2] : c[0];
          const uint32_t w1 = fz.as_set ? (cy != 0) : cy;
          const uint32_t co[4] = {local & 7, (local >> 3) & 7, (local >> 6) & 7, local >> 9};
#pragma unroll
          for (int ax = 0; ax < 4; ++ax) {
            if (w0) atomicAdd(&s_marg[ax][0][co[ax]], w0);
            if (w1) atomicAdd(&s_marg[ax][1][co[ax]], w1);
          }
        } else {
          score_cell(c, x, h1, h2, h3, ct, acc, fz.as_set);
        }
      }
    }

    // survivor totals: warp sums -> two shared 64-bit atomics per warp -> two global ones per box
    ns = __reduce_add_sync(0xFFFFFFFFu, ns);
    const unsigned long long wsum = (unsigned long long)__reduce_add_sync(0xFFFFFFFFu, sum >> 16) * 65536ull + __reduce_add_sync(0xFFFFFFFFu, sum & 0xFFFFu);
    if (lane == 0 && ns) {
      atomicAdd(&s_tot[0], (unsigned long long)ns);
      atomicAdd(&s_tot[1], wsum);
    }
    __syncthreads();
    if (tid == 0) {
      if (s_tot[0]) red_add_u64(&fz.totals[0], s_tot[0]);
      if (s_tot[1]) red_add_u64(&fz.totals[1], s_tot[1]);
    }
    if (regular && tid < 72) {
      // threads 0..63: one (axis, slot, coordinate) sum, added at the leaf; 64..71: the (axis,
      // slot) total, subtracted at the m-node. Junction / m / component as in score_cell.
      const int be = tid < 64 ? tid >> 4 : (tid - 64) >> 1;
      const int sl = tid < 64 ? (tid >> 3) & 1 : tid & 1;
      const int t = sl == 0 ? 1 : (box_kill0 ? 2 : 0);
      uint32_t h1, h2, m;
      if (be == 0) {
        h1 = bg2; h2 = bg3; m = bg1;
      } else if (be == 1) {
        h1 = min(bg1, bg2); h2 = bg3; m = max(bg1, bg2);
      } else if (be == 2) {
        h1 = bg1; h2 = min(bg2, bg3); m = max(bg2, bg3);
      } else {
        h1 = bg1; h2 = bg2; m = bg3;
      }
      const bool first = h1 > h2;
      const uint32_t J = (first ? h1 : h2) & 0xFFFFu;
      const int p = partner_of(t, be);
      const int pos = p - (p > be ? 1 : 0);
      const int comp = first ? pos : (pos + 2) % 3;
      long long v = 0;
      size_t row;
      if (tid < 64) {
        v = s_marg[be][sl][tid & 7];
        row = (size_t)ct.leafnode[s_seg[be] * 8 + (tid & 7)];
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) v -= s_marg[be][sl][k];
        row = m & 0xFFFFu;
      }
      if (v) red_add_u64(fz.Z + (size_t)(blockIdx.x % (unsigned)fz.z_rep) * fz.z_stride + (row * (size_t)ct.n + J) * 3 + comp,
                         (unsigned long long)v);
    }
  }
}

}  // namespace camus_dev
