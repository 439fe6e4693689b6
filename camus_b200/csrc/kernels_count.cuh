// K2: quartet topology counting, bit-sliced over gene trees.
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
// streamed HBM -> shared memory by the TMA engine (cp.async.bulk, 6 KB per tile, 4-stage mbarrier
// ring). 256 threads, each owning a 2x2x2x2 micro-box (16 cells, 48 counters in registers) so a
// 128-bit shared load feeds four cells.
#pragma once
#include <cuda_runtime.h>

#include "kernels_score.cuh"
#include "layout.cuh"
#include "ptx.cuh"

// Build-time switches kept for A/B measurements (profiles/r01_count_variants.md). Both lost on
// B200 for config 2 and are off: CAMUS_POPC_SKIP (skip the third POPC behind a warp-uniform branch
// when every tree of the group resolves the cell: 2.54 -> 3.22 ms, extra instructions + spills)
// and CAMUS_RING_BARRIER (per-stage empty mbarriers instead of one CTA barrier per tree group:
// 2.54 -> 2.61 ms).
#ifndef CAMUS_POPC_SKIP
#define CAMUS_POPC_SKIP 0
#endif
#ifndef CAMUS_RING_BARRIER
#define CAMUS_RING_BARRIER 0
#endif

namespace camus_dev {

constexpr int kCountThreads = 256;
constexpr int kCountStages = 4;
constexpr int kStageWords = 4 * kTileWords;  // 6144 words
constexpr int kStageBytes = kStageWords * 4;  // 24576 B
constexpr int kCountSmemBytes = kCountStages * kStageBytes + 128;  // + full/empty mbarriers

struct CountParams {
  const uint32_t* tri;  // [tile][group][kTileWords]
  int n_groups;
  BoxMap bm;       // blockIdx.x -> global box rank
  uint4* cells;    // table mode: [local box][kBoxCells] = (c(ab|cd), c(ac|bd), c(ad|bc), 0)
  int n_trees;     // for the mask of the last (partial) tree group
};

// Fused mode: the counts never leave the registers. The epilogue applies K3 (filter + constraint
// removal + scalars) and K4 (edge-score scatter into Z) per cell; no cell table exists in HBM.
struct FusedParams {
  int qmode;
  double threshold;
  const uint32_t* thr_tab;     // k_threshold_table
  int as_set;
  unsigned long long* Z;       // [n][n][3]
  unsigned long long* totals;  // [0] n_survivors, [1] sum_counts
  int z_rep;                   // number of privatised copies of Z (hot-address spreading), stride z_stride
  unsigned long long z_stride;
  int debug;                   // profiling only (CAMUS_B200_DEBUG): 1 = skip the score scatter, 2 = skip the whole epilogue
};

__device__ __forceinline__ uint32_t pick(const uint4& v, int hi, int lo) {
  int k = hi * 2 + lo;
  return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w));
}

// kComplete: every gene tree holds every taxon and is fully resolved (host-checked in
// camus_b200_add_gene_trees). Then each tree resolves each cell, so count(ad|bc) = G - count(ab|cd)
// - count(ac|bd): the third topology word, its POPC (the XU pipe is the kernel's narrowest) and
// the four planes only it reads are skipped; the ring then carries 16 KB instead of 24 KB per stage.
template <bool kFused, bool kComplete>
__global__ void __launch_bounds__(kCountThreads, 2) k_count(CountParams prm, FusedParams fz, ConstraintDev ct) {
  extern __shared__ __align__(128) uint32_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kCountStages * kStageWords);
  __shared__ uint32_t s_seg[4];

  const int tid = threadIdx.x;
  const int ta = tid & 3, tb = (tid >> 2) & 3, tc = (tid >> 4) & 3, td = tid >> 6;

  if (tid == 0) {
    uint32_t sa, sb, sc, sd;
    box_unrank(prm.bm.global(blockIdx.x), sa, sb, sc, sd);
    s_seg[0] = sa;
    s_seg[1] = sb;
    s_seg[2] = sc;
    s_seg[3] = sd;
    for (int s = 0; s < kCountStages; ++s) {
      mbar_init(&bars[s], 1);                                     // full: one arrive.expect_tx per refill
      mbar_init(&bars[kCountStages + s], kCountThreads / 32);      // empty: one arrive per consumer warp
    }
    fence_mbar_init();
  }
  __syncthreads();

  const int G = prm.n_groups;
  const uint32_t sa = s_seg[0], sb = s_seg[1], sc = s_seg[2], sd = s_seg[3];
  const size_t gstride = (size_t)kTileWords;
  const uint32_t* src0 = prm.tri + tile_index(sa, sb, sc) * (size_t)G * gstride;  // abc
  const uint32_t* src1 = prm.tri + tile_index(sa, sb, sd) * (size_t)G * gstride;  // abd
  const uint32_t* src2 = prm.tri + tile_index(sa, sc, sd) * (size_t)G * gstride;  // acd
  const uint32_t* src3 = prm.tri + tile_index(sb, sc, sd) * (size_t)G * gstride;  // bcd

  auto issue = [&](int g) {
    int s = g % kCountStages;
    uint32_t* dst = smem + s * kStageWords;
    constexpr int P = kTileWords / 3;  // words per plane (2 KB)
    if constexpr (!kComplete) {
      mbar_arrive_expect_tx(&bars[s], kStageBytes);
      bulk_g2s(dst, src0 + (size_t)g * gstride, kTileBytes, &bars[s]);
      bulk_g2s(dst + kTileWords, src1 + (size_t)g * gstride, kTileBytes, &bars[s]);
      bulk_g2s(dst + 2 * kTileWords, src2 + (size_t)g * gstride, kTileBytes, &bars[s]);
      bulk_g2s(dst + 3 * kTileWords, src3 + (size_t)g * gstride, kTileBytes, &bars[s]);
    } else {  // planes read by ab|cd and ac|bd only: abc{0,1} abd{0,2} acd{0,2} bcd{1,2}
      mbar_arrive_expect_tx(&bars[s], 8 * P * 4);
      bulk_g2s(dst, src0 + (size_t)g * gstride, 2 * P * 4, &bars[s]);
      bulk_g2s(dst + kTileWords, src1 + (size_t)g * gstride, P * 4, &bars[s]);
      bulk_g2s(dst + kTileWords + 2 * P, src1 + (size_t)g * gstride + 2 * P, P * 4, &bars[s]);
      bulk_g2s(dst + 2 * kTileWords, src2 + (size_t)g * gstride, P * 4, &bars[s]);
      bulk_g2s(dst + 2 * kTileWords + 2 * P, src2 + (size_t)g * gstride + 2 * P, P * 4, &bars[s]);
      bulk_g2s(dst + 3 * kTileWords + P, src3 + (size_t)g * gstride + P, 2 * P * 4, &bars[s]);
    }
  };

  if (tid == 0)
    for (int g = 0; g < kCountStages - 1 && g < G; ++g) issue(g);

  uint32_t cnt[16][3];
#pragma unroll
  for (int c = 0; c < 16; ++c) cnt[c][0] = cnt[c][1] = cnt[c][2] = 0;

  // cnt[c][0], cnt[c][1] = trees displaying ab|cd, ac|bd; cnt[c][2] = trees of the group range that
  // do NOT resolve the cell (absent taxon or polytomy). ad|bc = G - the three, recovered after the
  // loop. For complete binary trees the unresolved word is 0 and its POPC (XU pipe, the kernel's
  // narrowest pipe at 16 lanes/clk/SM) is skipped by a warp-uniform branch.
  const uint32_t last_mask = (prm.n_trees & 31) ? ((1u << (prm.n_trees & 31)) - 1u) : 0xFFFFFFFFu;
  const uint32_t vm_unused_guard = last_mask;
  for (int g = 0; g < G; ++g) {
    const int s = g % kCountStages;
    // refill the stage that group g-1 used, once all 8 warps have released it (no CTA barrier:
    // warps drift independently, only thread 0 waits here)
#if CAMUS_RING_BARRIER
    if (tid == 0 && g + kCountStages - 1 < G) {
      if (g >= 1) mbar_wait(&bars[kCountStages + (g - 1) % kCountStages], (uint32_t)(((g - 1) / kCountStages) & 1));
      issue(g + kCountStages - 1);
    }
    __syncwarp();
#else
    __syncthreads();
    if (tid == 0 && g + kCountStages - 1 < G) issue(g + kCountStages - 1);
#endif
    mbar_wait(&bars[s], (uint32_t)((g / kCountStages) & 1));
#if CAMUS_POPC_SKIP
    const uint32_t vm = (g == G - 1) ? last_mask : 0xFFFFFFFFu;
#endif

    const uint4* tABC = reinterpret_cast<const uint4*>(smem + s * kStageWords);
    const uint4* tABD = tABC + kTileWords / 4;
    const uint4* tACD = tABD + kTileWords / 4;
    const uint4* tBCD = tACD + kTileWords / 4;

    uint4 abc[3][2];
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) abc[p][kk] = tABC[((p * 8 + 2 * tc + kk) * 4 + tb) * 4 + ta];

#pragma unroll
    for (int ll = 0; ll < 2; ++ll) {
      const int z = 2 * td + ll;
      uint4 abd[3], acd[3], bcd[3];
#pragma unroll
      for (int p = 0; p < 3; ++p) {
        abd[p] = tABD[((p * 8 + z) * 4 + tb) * 4 + ta];
        acd[p] = tACD[((p * 8 + z) * 4 + tc) * 4 + ta];
        bcd[p] = tBCD[((p * 8 + z) * 4 + tc) * 4 + tb];
      }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const uint32_t ABC0 = pick(abc[0][kk], j, i), ABC1 = pick(abc[1][kk], j, i), ABC2 = pick(abc[2][kk], j, i);
            const uint32_t ABD0 = pick(abd[0], j, i), ABD1 = pick(abd[1], j, i), ABD2 = pick(abd[2], j, i);
            const uint32_t ACD0 = pick(acd[0], kk, i), ACD1 = pick(acd[1], kk, i), ACD2 = pick(acd[2], kk, i);
            const uint32_t BCD0 = pick(bcd[0], kk, j), BCD1 = pick(bcd[1], kk, j), BCD2 = pick(bcd[2], kk, j);
            const uint32_t t0 = (ABC0 & ABD0) | (ACD2 & BCD2);  // ab|cd
            const uint32_t t1 = (ABC1 & ACD0) | (ABD2 & BCD1);  // ac|bd
            const int c = ((ll * 2 + kk) * 2 + j) * 2 + i;
            cnt[c][0] += __popc(t0);
            cnt[c][1] += __popc(t1);
            if constexpr (kComplete) continue;
            const uint32_t t2 = (ABD1 & ACD1) | (ABC2 & BCD0);  // ad|bc
#if CAMUS_POPC_SKIP
            const uint32_t un = vm & ~(t0 | t1 | t2);
            if (__any_sync(0xFFFFFFFFu, un != 0)) cnt[c][2] += __popc(un);
#else
            cnt[c][2] += __popc(t2);
#endif
          }
    }
#if CAMUS_RING_BARRIER
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&bars[kCountStages + s]);
#endif
  }
#if CAMUS_POPC_SKIP
  {
    const uint32_t gt = (uint32_t)prm.n_trees;
#pragma unroll
    for (int c = 0; c < 16; ++c) cnt[c][2] = gt - cnt[c][0] - cnt[c][1] - cnt[c][2];
  }
#else
  (void)vm_unused_guard;
  if constexpr (kComplete) {
    const uint32_t gt = (uint32_t)prm.n_trees;
#pragma unroll
    for (int c = 0; c < 16; ++c) cnt[c][2] = gt - cnt[c][0] - cnt[c][1];
  }
#endif

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
            const int c = ((ll * 2 + kk) * 2 + j) * 2 + i;
            out[cell_local(2 * ta + i, 2 * tb + j, 2 * tc + kk, 2 * td + ll)] = make_uint4(cnt[c][0], cnt[c][1], cnt[c][2], 0u);
          }
  } else {
    // ---- fused K3 + K4 on the register-resident counts ----
    // Phase 1 (per thread, its 16 cells): filter + constraint removal; survivors are COMPACTED into
    // a queue in the shared memory of the now idle tile ring. Phase 2: the queue is scored with
    // all lanes busy. Survivors are sparse under -q 1/2 (a few % of the cells), so scoring them in
    // place would run the whole scatter code for nearly every warp step with 1-3 lanes active.
    __shared__ unsigned long long s_red[2][kCountThreads / 32];
    __shared__ uint32_t s_qn;
    if (tid == 0) s_qn = 0;
    __syncthreads();  // also: every issued bulk copy has been waited for, the ring can be reused
    uint4* queue = reinterpret_cast<uint4*>(smem);  // up to kBoxCells entries (64 KB of the 96 KB ring)
    const uint32_t N = (uint32_t)ct.n_taxa;
    const uint32_t a0 = sa * 8 + 2 * ta, b0 = sb * 8 + 2 * tb, c0 = sc * 8 + 2 * tc, d0 = sd * 8 + 2 * td;
    const int lane = tid & 31;
    unsigned long long ns = 0, sum = 0;
#pragma unroll
    for (int ll = 0; ll < 2; ++ll)
#pragma unroll
      for (int kk = 0; kk < 2; ++kk)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int ci = ((ll * 2 + kk) * 2 + j) * 2 + i;
            const uint32_t x[4] = {a0 + i, b0 + j, c0 + kk, d0 + ll};
            uint32_t c[3] = {cnt[ci][0], cnt[ci][1], cnt[ci][2]};
            bool live = x[0] < x[1] && x[1] < x[2] && x[2] < x[3] && x[3] < N && (c[0] | c[1] | c[2]);
            if (fz.debug == 2) {
              if (live) sum += c[0] + c[1] + c[2];
              continue;
            }
            if (live && fz.qmode != 0) {
              filter_cell(c, x, fz.qmode, fz.threshold, fz.thr_tab, ct.tip_of_rank);
              live = (c[0] | c[1] | c[2]) != 0;
            }
            if (live) {
              const uint32_t g1 = ct.lcaD[(size_t)x[0] * N + x[1]];
              const uint32_t g2 = ct.lcaD[(size_t)x[1] * N + x[2]];
              const uint32_t g3 = ct.lcaD[(size_t)x[2] * N + x[3]];
              if (constraint_topology(g1, g2, g3) == 0) c[0] = 0;
              else c[2] = 0;
              live = (c[0] | c[1] | c[2]) != 0;
            }
            // warp-aggregated queue push: one shared atomic per warp step
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, live);
            if (m) {
              uint32_t base = 0;
              if (lane == 0) base = atomicAdd(&s_qn, (uint32_t)__popc(m));
              base = __shfl_sync(0xFFFFFFFFu, base, 0);
              if (live) {
                ns += (c[0] != 0) + (c[1] != 0) + (c[2] != 0);
                sum += (unsigned long long)c[0] + c[1] + c[2];
                const uint32_t local = cell_local(2 * ta + i, 2 * tb + j, 2 * tc + kk, 2 * td + ll);
                queue[base + __popc(m & ((1u << lane) - 1))] = make_uint4(local, c[0], c[1], c[2]);
              }
            }
          }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      ns += __shfl_xor_sync(0xFFFFFFFFu, ns, off);
      sum += __shfl_xor_sync(0xFFFFFFFFu, sum, off);
    }
    if (lane == 0) {
      s_red[0][tid >> 5] = ns;
      s_red[1][tid >> 5] = sum;
    }
    __syncthreads();
    if (tid == 0) {
      unsigned long long t0 = 0, t1 = 0;
#pragma unroll
      for (int w = 0; w < kCountThreads / 32; ++w) {
        t0 += s_red[0][w];
        t1 += s_red[1][w];
      }
      if (t0) red_add_u64(&fz.totals[0], t0);
      if (t1) red_add_u64(&fz.totals[1], t1);
    }
    if (fz.debug != 1) {
      const uint32_t qn = s_qn;
      GlobalAcc acc{fz.Z + (size_t)(blockIdx.x % (unsigned)fz.z_rep) * fz.z_stride};
#pragma unroll 1
      for (uint32_t e = tid; e < qn; e += kCountThreads) {
        const uint4 q = queue[e];
        const uint32_t x[4] = {sa * 8 + (q.x & 7), sb * 8 + ((q.x >> 3) & 7), sc * 8 + ((q.x >> 6) & 7), sd * 8 + (q.x >> 9)};
        const uint32_t c[3] = {q.y, q.z, q.w};
        const uint32_t g1 = ct.lcaD[(size_t)x[0] * N + x[1]];
        const uint32_t g2 = ct.lcaD[(size_t)x[1] * N + x[2]];
        const uint32_t g3 = ct.lcaD[(size_t)x[2] * N + x[3]];
        score_cell(c, x, g1, g2, g3, ct, acc, fz.as_set);
      }
    }
  }
}

}  // namespace camus_dev
