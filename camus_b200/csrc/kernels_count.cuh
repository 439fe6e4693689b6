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

#include "kernels_score.cuh"
#include "layout.cuh"
#include "ptx.cuh"

namespace camus_dev {

constexpr int kCountThreads = 256;
constexpr int kCountStages = 4;
constexpr int kPlaneWords = kTileWords / 3;  // 512 words = 2 KB: one plane of one tile

// kComplete: every gene tree holds every taxon and is fully resolved (host-checked in
// camus_b200_add_gene_trees) and there are at most 65535 trees. Then each tree resolves each
// cell, so count(ad|bc) = G - count(ab|cd) - count(ac|bd): the third topology word, its POPC and
// the four planes only it reads are skipped (8 planes = 16 KB per stage instead of 12 = 24 KB),
// and the two remaining counters share one register (16 + 16 bits). Fewer registers and less
// shared memory let 3 CTAs share an SM instead of 2, which matters because the kernel is bound
// by latency at 16 warps/SM, not by a pipe (profiles/r01_k_count_config3_ncu_full.txt).
template <bool kComplete>
struct CountCfg {
  static constexpr int kPlanes = kComplete ? 8 : 12;
  static constexpr int kStageWords = kPlanes * kPlaneWords;
  static constexpr int kStageBytes = kStageWords * 4;
  static constexpr int kSmemBytes = kCountStages * kStageBytes + 64;  // + mbarriers
  static constexpr int kMinBlocks = kComplete ? 3 : 2;
};

struct CountParams {
  const uint32_t* tri;  // [tile][group][kTileWords]
  int n_groups;
  BoxMap bm;       // blockIdx.x -> global box rank
  uint4* cells;    // table mode: [local box][kBoxCells] = (c(ab|cd), c(ac|bd), c(ad|bc), 0)
  int n_trees;
  const uint32_t* tile_slot;  // class-scheme shares store only the tiles they read: tile index -> slot in tri; nullptr = identity
};

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

template <bool kFused, bool kComplete>
__global__ void __launch_bounds__(kCountThreads, CountCfg<kComplete>::kMinBlocks)
k_count(CountParams prm, FusedParams fz, ConstraintDev ct) {
  using Cfg = CountCfg<kComplete>;
  extern __shared__ __align__(128) uint32_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kCountStages * Cfg::kStageWords);
  __shared__ uint32_t s_seg[5];  // four segments, [4] = regular box
  __shared__ uint64_t s_tile[4];  // slots of the tiles abc, abd, acd, bcd in prm.tri

  const int tid = threadIdx.x;
  const int ta = tid & 3, tb = (tid >> 2) & 3, tc = (tid >> 4) & 3, td = tid >> 6;

  if (tid == 0) {
    uint32_t sa, sb, sc, sd;
    box_unrank(prm.bm.global(blockIdx.x), sa, sb, sc, sd);
    s_seg[0] = sa;
    s_seg[1] = sb;
    s_seg[2] = sc;
    s_seg[3] = sd;
    auto slot = [&](uint32_t s0, uint32_t s1, uint32_t s2) {
      const uint64_t t = tile_index(s0, s1, s2);
      return prm.tile_slot ? (uint64_t)__ldg(prm.tile_slot + t) : t;
    };
    s_tile[0] = slot(sa, sb, sc);
    s_tile[1] = slot(sa, sb, sd);
    s_tile[2] = slot(sa, sc, sd);
    s_tile[3] = slot(sb, sc, sd);
    if constexpr (kFused) {
      const uint8_t* U = ct.seg_uniform;
      const uint32_t S = (uint32_t)ct.n_seg;
      s_seg[4] = fz.regular && sa < sb && sb < sc && sc < sd && U[sa * S + sb] && U[sb * S + sc] && U[sc * S + sd];
    }
    for (int s = 0; s < kCountStages; ++s) mbar_init(&bars[s], 1);
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
    for (int g = 0; g < kCountStages && g < G; ++g) issue(g);

  // general: cnt[c][t] = trees displaying topology t. complete: cnt[c][0] = c(ab|cd) | c(ac|bd) << 16.
  constexpr int kCnt = kComplete ? 1 : 3;
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

  for (int g = 0; g < G; ++g) {
    const int s = g % kCountStages;
    // One CTA barrier per TWO groups: at even g everyone has finished groups g-2 and g-1, whose
    // stages are refilled with g+2 and g+3 (the ring holds g .. g+3 in flight or ready).
    if ((g & 1) == 0 && g > 0) {
      __syncthreads();
      if (tid == 0 && !(kFused && fz.debug == 3)) {
        if (g + 2 < G) issue(g + 2);
        if (g + 3 < G) issue(g + 3);
      }
    }
    // debug 3 (timing only): the ring is filled once and re-read, i.e. the loop without tile traffic
    if (!(kFused && fz.debug == 3 && g >= kCountStages)) mbar_wait(&bars[s], (uint32_t)((g / kCountStages) & 1));

    const uint4* st = reinterpret_cast<const uint4*>(smem + s * Cfg::kStageWords);
    auto ld = [&](int role, int plane, uint32_t off) -> uint4 { return st[plane_slot<kComplete>(role, plane) * kPlaneU4 + off]; };

    uint4 abc[3][2];
#pragma unroll
    for (int p = 0; p < (kComplete ? 2 : 3); ++p)
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) abc[p][kk] = ld(0, p, o_abc + 16 * kk);

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

  // counts of the three topologies of cell ci
  auto counts_of = [&](int ci, uint32_t c[3]) {
    if constexpr (kComplete) {
      c[0] = cnt[ci][0] & 0xFFFFu;
      c[1] = cnt[ci][0] >> 16;
      c[2] = (uint32_t)prm.n_trees - c[0] - c[1];
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
    // ---- fused K3 + K4 on the register-resident counts ----
    // Phase 1 (per thread, its 16 cells): filter + constraint removal; survivors are COMPACTED into
    // a queue in the shared memory of the now idle tile ring. Phase 2: the queue is scored with
    // all lanes busy. Survivors are sparse under -q 1/2 (a few % of the cells).
    //
    // Regular boxes (the three consecutive segment pairs have one LCA each, ConstraintDev::
    // seg_uniform; half of all boxes of a random 1000-taxon constraint) skip the queue: junction,
    // m-node and component of a (bottom taxon, topology) are box constants, so the box's whole
    // contribution is 4 axes x 2 topologies x 8 marginal sums of the surviving counts (+ their
    // totals, subtracted at the m-nodes): 72 reductions per box instead of 8 per survivor.
    __shared__ unsigned long long s_red[2][kCountThreads / 32];
    __shared__ uint32_t s_qn;
    __shared__ uint32_t s_marg[4][2][8];  // [bottom axis a..d][topology slot][coordinate]
    if (tid == 0) s_qn = 0;
    if (tid < 64) (&s_marg[0][0][0])[tid] = 0;
    const bool regular = s_seg[4] != 0;
    __syncthreads();  // also: every issued bulk copy has been waited for, the ring can be reused
    static_assert(kCountStages * Cfg::kStageBytes >= kBoxCells * 16, "survivor queue must fit the ring");
    uint4* queue = reinterpret_cast<uint4*>(smem);  // up to kBoxCells entries (64 KB)
    const uint32_t N = (uint32_t)ct.n_taxa;
    const uint32_t a0 = sa * 8 + 2 * ta, b0 = sb * 8 + 2 * tb, c0 = sc * 8 + 2 * tc, d0 = sd * 8 + 2 * td;
    const int lane = tid & 31;
    unsigned long long ns = 0, sum = 0;
    // packed LCAs of the box's first taxa: THE LCAs of a regular box
    const uint32_t bg1 = ct.lcaD[(size_t)(sa * 8) * N + sb * 8], bg2 = ct.lcaD[(size_t)(sb * 8) * N + sc * 8],
                   bg3 = ct.lcaD[(size_t)(sc * 8) * N + sd * 8];
    const bool kill0 = constraint_topology(bg1, bg2, bg3) == 0;  // the constraint displays ab|cd here (else ad|bc)
    if (regular) {
      // slot 0 = ac|bd, slot 1 = the one of ab|cd / ad|bc the constraint does not display
      uint32_t ma[2][2] = {{0, 0}, {0, 0}}, mb[2][2] = {{0, 0}, {0, 0}}, mc[2][2] = {{0, 0}, {0, 0}}, md[2][2] = {{0, 0}, {0, 0}};
#pragma unroll
      for (int ll = 0; ll < 2; ++ll)
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const uint32_t x[4] = {a0 + i, b0 + j, c0 + kk, d0 + ll};
              uint32_t c[3];
              counts_of(((ll * 2 + kk) * 2 + j) * 2 + i, c);
              if (fz.debug >= 2) {
                sum += c[0] + c[1] + c[2];
                continue;
              }
              if (fz.qmode != 0 && (c[0] | c[1] | c[2])) filter_cell(c, x, fz.qmode, fz.threshold, fz.thr_tab, ct.tip_of_rank);
              uint32_t w0 = c[1], w1 = kill0 ? c[2] : c[0];
              ns += (w0 != 0) + (w1 != 0);
              sum += (unsigned long long)w0 + w1;
              if (fz.as_set) {
                w0 = w0 != 0;
                w1 = w1 != 0;
              }
              ma[0][i] += w0; mb[0][j] += w0; mc[0][kk] += w0; md[0][ll] += w0;
              ma[1][i] += w1; mb[1][j] += w1; mc[1][kk] += w1; md[1][ll] += w1;
            }
      if (fz.debug == 0) {
        // warp sums over the lanes that share the coordinate, one shared atomic per (warp, sum)
        const uint32_t mask_a = 0x11111111u << ta, mask_b = 0x000F000Fu << (4 * tb);
        const uint32_t mask_c = lane < 16 ? 0x0000FFFFu : 0xFFFF0000u;
#pragma unroll
        for (int sl = 0; sl < 2; ++sl)
#pragma unroll
          for (int par = 0; par < 2; ++par) {
            uint32_t v = __reduce_add_sync(mask_a, ma[sl][par]);
            if (lane < 4 && v) atomicAdd(&s_marg[0][sl][2 * ta + par], v);
            v = __reduce_add_sync(mask_b, mb[sl][par]);
            if (lane == 4 * tb && v) atomicAdd(&s_marg[1][sl][2 * tb + par], v);
            v = __reduce_add_sync(mask_c, mc[sl][par]);
            if ((lane & 15) == 0 && v) atomicAdd(&s_marg[2][sl][2 * tc + par], v);
            v = __reduce_add_sync(0xFFFFFFFFu, md[sl][par]);
            if (lane == 0 && v) atomicAdd(&s_marg[3][sl][2 * td + par], v);
          }
      }
    } else {
#pragma unroll
    for (int ll = 0; ll < 2; ++ll)
#pragma unroll
      for (int kk = 0; kk < 2; ++kk)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const uint32_t x[4] = {a0 + i, b0 + j, c0 + kk, d0 + ll};
            uint32_t c[3];
            counts_of(((ll * 2 + kk) * 2 + j) * 2 + i, c);
            bool live = x[0] < x[1] && x[1] < x[2] && x[2] < x[3] && x[3] < N && (c[0] | c[1] | c[2]);
            if (fz.debug >= 2) {
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
    if (regular) {
      if (fz.debug == 0 && tid < 72) {
        // threads 0..63: one (axis, slot, coordinate) sum, added at the leaf; 64..71: the (axis,
        // slot) total, subtracted at the m-node. Junction / m / component as in score_cell.
        const int be = tid < 64 ? tid >> 4 : (tid - 64) >> 1;
        const int sl = tid < 64 ? (tid >> 3) & 1 : tid & 1;
        const int t = sl == 0 ? 1 : (kill0 ? 2 : 0);
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
    } else if (fz.debug != 1) {
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
