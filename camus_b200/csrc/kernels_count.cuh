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

namespace camus_dev {

constexpr int kCountThreads = 256;
constexpr int kCountStages = 4;
constexpr int kStageWords = 4 * kTileWords;  // 6144 words
constexpr int kStageBytes = kStageWords * 4;  // 24576 B
constexpr int kCountSmemBytes = kCountStages * kStageBytes + 64;

struct CountParams {
  const uint32_t* tri;  // [tile][group][kTileWords]
  int n_groups;
  uint64_t box0;   // first box rank of this launch
  uint4* cells;    // table mode: [box - box0][kBoxCells] = (c(ab|cd), c(ac|bd), c(ad|bc), 0)
};

// Fused mode: the counts never leave the registers. The epilogue applies K3 (filter + constraint
// removal + scalars) and K4 (edge-score scatter into Z) per cell; no cell table exists in HBM.
struct FusedParams {
  int qmode;
  double threshold;
  int as_set;
  unsigned long long* Z;       // [n][n][3]
  unsigned long long* totals;  // [0] n_survivors, [1] sum_counts
};

__device__ __forceinline__ uint32_t pick(const uint4& v, int hi, int lo) {
  int k = hi * 2 + lo;
  return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w));
}

template <bool kFused>
__global__ void __launch_bounds__(kCountThreads, 2) k_count(CountParams prm, FusedParams fz, ConstraintDev ct) {
  extern __shared__ __align__(128) uint32_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kCountStages * kStageWords);
  __shared__ uint32_t s_seg[4];

  const int tid = threadIdx.x;
  const int ta = tid & 3, tb = (tid >> 2) & 3, tc = (tid >> 4) & 3, td = tid >> 6;

  if (tid == 0) {
    uint32_t sa, sb, sc, sd;
    box_unrank(prm.box0 + blockIdx.x, sa, sb, sc, sd);
    s_seg[0] = sa;
    s_seg[1] = sb;
    s_seg[2] = sc;
    s_seg[3] = sd;
    for (int s = 0; s < kCountStages; ++s) mbar_init(&bars[s], 1);
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
    mbar_arrive_expect_tx(&bars[s], kStageBytes);
    bulk_g2s(dst, src0 + (size_t)g * gstride, kTileBytes, &bars[s]);
    bulk_g2s(dst + kTileWords, src1 + (size_t)g * gstride, kTileBytes, &bars[s]);
    bulk_g2s(dst + 2 * kTileWords, src2 + (size_t)g * gstride, kTileBytes, &bars[s]);
    bulk_g2s(dst + 3 * kTileWords, src3 + (size_t)g * gstride, kTileBytes, &bars[s]);
  };

  if (tid == 0)
    for (int g = 0; g < kCountStages - 1 && g < G; ++g) issue(g);

  uint32_t cnt[16][3];
#pragma unroll
  for (int c = 0; c < 16; ++c) cnt[c][0] = cnt[c][1] = cnt[c][2] = 0;

  for (int g = 0; g < G; ++g) {
    const int s = g % kCountStages;
    // everyone has finished group g-1, whose stage is the one refilled now
    __syncthreads();
    if (tid == 0 && g + kCountStages - 1 < G) issue(g + kCountStages - 1);
    mbar_wait(&bars[s], (uint32_t)((g / kCountStages) & 1));

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
            const uint32_t t2 = (ABD1 & ACD1) | (ABC2 & BCD0);  // ad|bc
            const int c = ((ll * 2 + kk) * 2 + j) * 2 + i;
            cnt[c][0] += __popc(t0);
            cnt[c][1] += __popc(t1);
            cnt[c][2] += __popc(t2);
          }
    }
  }

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
    __shared__ unsigned long long s_red[2][kCountThreads / 32];
    const uint32_t N = (uint32_t)ct.n_taxa;
    const uint32_t a0 = sa * 8 + 2 * ta, b0 = sb * 8 + 2 * tb, c0 = sc * 8 + 2 * tc, d0 = sd * 8 + 2 * td;
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
            const bool valid = x[0] < x[1] && x[1] < x[2] && x[2] < x[3] && x[3] < N;
            if (!valid || !(c[0] | c[1] | c[2])) continue;
            if (fz.qmode != 0) {
              const int ti[4] = {ct.tip_of_rank[x[0]], ct.tip_of_rank[x[1]], ct.tip_of_rank[x[2]], ct.tip_of_rank[x[3]]};
              apply_filter(c, ti, fz.qmode, fz.threshold);
              if (!(c[0] | c[1] | c[2])) continue;
            }
            const uint32_t g1 = ct.lcaD[(size_t)x[0] * N + x[1]];
            const uint32_t g2 = ct.lcaD[(size_t)x[1] * N + x[2]];
            const uint32_t g3 = ct.lcaD[(size_t)x[2] * N + x[3]];
            if (constraint_topology(g1, g2, g3) == 0) c[0] = 0;
            else c[2] = 0;
            if (!(c[0] | c[1] | c[2])) continue;
            ns += (c[0] != 0) + (c[1] != 0) + (c[2] != 0);
            sum += (unsigned long long)c[0] + c[1] + c[2];
            score_cell(c, x, g1, g2, g3, ct, fz.Z, fz.as_set);
          }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      ns += __shfl_xor_sync(0xFFFFFFFFu, ns, off);
      sum += __shfl_xor_sync(0xFFFFFFFFu, sum, off);
    }
    if ((tid & 31) == 0) {
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
  }
}

}  // namespace camus_dev
