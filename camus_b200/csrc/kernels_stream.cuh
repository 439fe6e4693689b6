// K2+K3+K4, streaming form (default): persistent CTAs walk the box list, the tile ring never
// drains between boxes, and there is no CTA barrier in the count loop.
//
// Same arithmetic as k_count (kernels_count.cuh: bit-sliced counting over rooted-triple planes,
// replacing graphs.QuartetsFromTree + prep.processQuartets, internal/graphs/quartet.go:112-123,
// internal/prep/preprocess.go:71-124; epilogue = prep.filterQuartets quartetfilter.go:67-96,
// constraint removal preprocess.go:51-57, edge-score scatter quartettotals.go:43-157). What
// changed is the schedule, driven by the round-1 profile (profiles/r01_k_count_config3_ncu_full.txt:
// barrier and scoreboard stalls, 27 % of the instructions in the per-box prologue / epilogue):
//
//  * grid = resident CTAs only (SMs x CTAs/SM); a CTA takes every grid-th box and describes its
//    boxes in a small shared ring a few boxes ahead of the warps;
//  * the (box, tree group) pairs of a CTA form ONE stream of positions p = k * G + g. Position p
//    lives in ring stage p % S; when the last of the 8 warps is done with a stage (a shared
//    counter, acq_rel) that warp's lane 0 refills it with position p + S, which may already belong
//    to the next box: tiles keep arriving during the epilogue and the next box starts hot;
//  * warps drift up to S positions apart; the only CTA barrier is one per box, in the epilogue;
//  * epilogue on registers: the filter works on (lo, mid, hi) without touching the counts, Keep's
//    truncated product comes from a table staged in shared memory, survivor totals stay in
//    registers until the kernel ends (8 global reductions per CTA instead of 2 per box), survivors
//    of irregular boxes are compacted per WARP (no shared atomics, no barrier) and flushed in
//    batches, the LCAs of a thread's 2x2 taxon pairs are loaded once per box.
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

#include "kernels_count.cuh"
#include "kernels_score.cuh"
#include "layout.cuh"
#include "ptx.cuh"

namespace camus_dev {

constexpr int kBoxRing = 8;        // box descriptors kept in shared memory (>= lookahead 4 + 3)
constexpr int kWarpQueue = 64;     // survivor queue entries per warp
constexpr int kThrTabMax = 2048;   // Keep table entries staged in shared memory (else read from global)

struct BoxInfo {
  unsigned long long tile[4];  // word offsets of the tiles abc, abd, acd, bcd of group 0 in tri
  uint32_t seg[4];
  uint32_t regular, valid, zrep, pad;
};

// counting mode: 0 = general (3 counters per cell), 1 = complete trees (2 topologies, 16 + 16 bits)
template <int kMode>
struct StreamCfg {
  static constexpr bool kComplete = kMode == 1;
  static constexpr int kPlanes = kComplete ? 8 : 12;
  static constexpr int kStages = 4;
  static constexpr int kStageWords = kPlanes * kPlaneWords;
  static constexpr int kStageBytes = kStageWords * 4;
  static constexpr int kMinBlocks = kComplete ? 3 : 2;
  static constexpr int kCnt = kComplete ? 1 : 3;
  // shared memory: ring | full barriers | done counters | box ring | marginal sums (3 slots) |
  // Keep table (u16) | per-warp survivor queues (uint4)
  static constexpr int kOffBars = kStages * kStageBytes;
  static constexpr int kOffDone = kOffBars + kStages * 8;  // S counters, then the CTA's two survivor totals (u64)
  static constexpr int kOffBoxes = kOffDone + 32;
  static constexpr int kOffMarg = kOffBoxes + kBoxRing * (int)sizeof(BoxInfo);
  static constexpr int kOffThr = kOffMarg + 3 * 64 * 4;
  static constexpr int kOffQueue = kOffThr + kThrTabMax * 2;
  static constexpr int kQueueBytes = (kCountThreads / 32) * kWarpQueue * (kComplete ? 8 : 16);
  static constexpr int kSmemBytes = kOffQueue + kQueueBytes;
};

struct StreamParams {
  const uint32_t* tri;  // [tile slot][group][kTileWords]
  int n_groups, n_trees;
  BoxMap bm;                   // local box index -> global box rank
  unsigned long long n_boxes;  // local boxes of this launch
  unsigned int* work;          // next unclaimed local box (zeroed before the launch); nullptr = static round-robin
  const uint32_t* tile_slot;   // class-scheme shares: tile index -> slot; nullptr = identity
};

__device__ __forceinline__ uint32_t atom_add_acq_rel_shared(uint32_t* p, uint32_t v) {
  uint32_t old;
#if defined(CAMUS_STREAM_RELAXED)  // A/B build: no fence around the counter
  asm volatile("atom.relaxed.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
#else
  asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
#endif
  return old;
}

// Which topologies of one cell survive prep.filterQuartets (quartetfilter.go:76-96): bit t of the
// result. c[] = counts of ab|cd, ac|bd, ad|bc; thr(s) = uint32(float64(t) * float64(s)).
// keep  (Threshold.Keep, :67-74): mode 2 drops the (unique) minimum, mode 1 nothing;
// !keep: only the maximum stays; a tie of the two largest is broken in the reference's topology
// order (apply_filter). Zero counts are absent keys and never "survive".
// the rare tie of the two largest counts under a failed Keep: the reference's stable topology order decides
__device__ __noinline__ uint32_t filter_tie_break(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t x0, uint32_t x1, uint32_t x2,
                                                  uint32_t x3, int qmode, double threshold, const int32_t* __restrict__ tip_of_rank) {
  uint32_t cc[3] = {c0, c1, c2};
  const int ti[4] = {tip_of_rank[x0], tip_of_rank[x1], tip_of_rank[x2], tip_of_rank[x3]};
  apply_filter(cc, ti, qmode, threshold);
  return (cc[0] != 0 ? 1u : 0u) | (cc[1] != 0 ? 2u : 0u) | (cc[2] != 0 ? 4u : 0u);
}

template <class Thr>
__device__ __forceinline__ uint32_t filter_survivors(const uint32_t c[3], const uint32_t x[4], int qmode, double threshold,
                                                     Thr&& thr, const int32_t* __restrict__ tip_of_rank) {
  const uint32_t nz = (c[0] != 0 ? 1u : 0u) | (c[1] != 0 ? 2u : 0u) | (c[2] != 0 ? 4u : 0u);
  if (qmode == 0) return nz;
  const uint32_t lo = min(c[0], min(c[1], c[2])), hi = max(c[0], max(c[1], c[2]));
  const uint32_t mid = c[0] + c[1] + c[2] - lo - hi;
  if (thr(lo + mid) < mid - lo) {  // keep
    if (qmode != 2) return nz;
    return nz & ((c[0] != lo ? 1u : 0u) | (c[1] != lo ? 2u : 0u) | (c[2] != lo ? 4u : 0u));
  }
  if (mid != hi) return nz & ((c[0] == hi ? 1u : 0u) | (c[1] == hi ? 2u : 0u) | (c[2] == hi ? 4u : 0u));
  if (hi == 0) return 0;
  return filter_tie_break(c[0], c[1], c[2], x[0], x[1], x[2], x[3], qmode, threshold, tip_of_rank);
}

// Score the survivors a warp has queued (irregular boxes): 32 at a time, all lanes busy.
template <bool kComplete>
__device__ __noinline__ void flush_warp_queue(const void* queue_, uint32_t qn, uint32_t sa, uint32_t sb, uint32_t sc, uint32_t sd,
                                              ConstraintDev ct, unsigned long long* Zrep, int as_set) {
  using QEntry = typename std::conditional<kComplete, uint2, uint4>::type;
  const QEntry* queue = reinterpret_cast<const QEntry*>(queue_);
  const uint32_t N = (uint32_t)ct.n_taxa;
  GlobalAcc acc{Zrep};
  __syncwarp();
  for (uint32_t e = threadIdx.x & 31; e < qn; e += 32) {
    const QEntry q = queue[e];
    uint32_t local, c[3];
    if constexpr (kComplete) {
      local = q.x & 0xFFFu;
      c[0] = (q.x >> 12) & 0xFFFFu;  // entry = local (12 bits) | c0 | c1 | c2 (16 bits each)
      c[1] = ((q.x >> 28) | (q.y << 4)) & 0xFFFFu;
      c[2] = (q.y >> 12) & 0xFFFFu;
    } else {
      local = q.x;
      c[0] = q.y; c[1] = q.z; c[2] = q.w;
    }
    const uint32_t x[4] = {sa * 8 + (local & 7), sb * 8 + ((local >> 3) & 7), sc * 8 + ((local >> 6) & 7), sd * 8 + (local >> 9)};
    const uint32_t h1 = ct.lcaD[(size_t)x[0] * N + x[1]];
    const uint32_t h2 = ct.lcaD[(size_t)x[1] * N + x[2]];
    const uint32_t h3 = ct.lcaD[(size_t)x[2] * N + x[3]];
    score_cell(c, x, h1, h2, h3, ct, acc, as_set);
  }
  __syncwarp();
}

template <int kMode>
__global__ void __launch_bounds__(kCountThreads, StreamCfg<kMode>::kMinBlocks)
k_count_stream(StreamParams prm, FusedParams fz, ConstraintDev ct) {
  using Cfg = StreamCfg<kMode>;
  constexpr bool kComplete = Cfg::kComplete;
  constexpr int S = Cfg::kStages;
  constexpr int kWarps = kCountThreads / 32;
  extern __shared__ __align__(128) uint32_t smem[];
  unsigned char* sm8 = reinterpret_cast<unsigned char*>(smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(sm8 + Cfg::kOffBars);
  uint32_t* done = reinterpret_cast<uint32_t*>(sm8 + Cfg::kOffDone);
  BoxInfo* boxes = reinterpret_cast<BoxInfo*>(sm8 + Cfg::kOffBoxes);
  uint32_t* marg = reinterpret_cast<uint32_t*>(sm8 + Cfg::kOffMarg);  // [3][axis 4][slot 2][coordinate 8]
  uint16_t* s_thr = reinterpret_cast<uint16_t*>(sm8 + Cfg::kOffThr);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ta = tid & 3, tb = (tid >> 2) & 3, tc = (tid >> 4) & 3, td = tid >> 6;
  const uint32_t G = (uint32_t)prm.n_groups;
  const uint32_t LA = (S + G - 1) / G;  // boxes a refill can reach ahead of the box being counted
  const uint32_t N = (uint32_t)ct.n_taxa;
  const bool thr_smem = prm.n_trees + 2 <= kThrTabMax;

  // Boxes are claimed from a global counter, in order: the CTAs in flight always work on ~grid
  // consecutive boxes, which share triple tiles in L2 (a static round-robin lets the CTAs drift
  // apart over thousands of boxes).
  auto claim = [&](uint32_t k) -> uint32_t {
    return prm.work ? atomicAdd(prm.work, 1u) : blockIdx.x + k * gridDim.x;
  };
  auto describe = [&](uint32_t k, uint32_t i) {  // thread 0: this CTA's k-th box is local box i
    BoxInfo& b = boxes[k % kBoxRing];
    if (i >= prm.n_boxes) {
      b.valid = 0;
      return;
    }
    uint32_t sa, sb, sc, sd;
    box_unrank(prm.bm.global(i), sa, sb, sc, sd);
    b.seg[0] = sa; b.seg[1] = sb; b.seg[2] = sc; b.seg[3] = sd;
    auto slot = [&](uint32_t s0, uint32_t s1, uint32_t s2) -> unsigned long long {
      const uint64_t t = tile_index(s0, s1, s2);
      return (prm.tile_slot ? (unsigned long long)__ldg(prm.tile_slot + t) : t) * G * kTileWords;
    };
    b.tile[0] = slot(sa, sb, sc);
    b.tile[1] = slot(sa, sb, sd);
    b.tile[2] = slot(sa, sc, sd);
    b.tile[3] = slot(sb, sc, sd);
    const uint8_t* U = ct.seg_uniform;
    const uint32_t M = (uint32_t)ct.n_seg;
    b.regular = fz.regular && sa < sb && sb < sc && sc < sd && U[sa * M + sb] && U[sb * M + sc] && U[sc * M + sd];
    b.zrep = (uint32_t)(i % (unsigned)fz.z_rep);
    b.valid = 1;
  };
  auto issue = [&](uint32_t k, uint32_t g, int s) {  // one thread: group g of box k -> stage s, if the box exists
    while (g >= G) {
      g -= G;
      ++k;
    }
    const BoxInfo& b = boxes[k % kBoxRing];
    if (!b.valid) return;
    uint32_t* dst = smem + s * Cfg::kStageWords;
    constexpr int P = kPlaneWords, PB = kPlaneWords * 4;
    const size_t o = (size_t)g * kTileWords;
    const uint32_t *src0 = prm.tri + b.tile[0] + o, *src1 = prm.tri + b.tile[1] + o, *src2 = prm.tri + b.tile[2] + o,
                   *src3 = prm.tri + b.tile[3] + o;
    mbar_arrive_expect_tx(&full[s], Cfg::kStageBytes);
    if constexpr (!kComplete) {
      bulk_g2s(dst, src0, 3 * PB, &full[s]);
      bulk_g2s(dst + 3 * P, src1, 3 * PB, &full[s]);
      bulk_g2s(dst + 6 * P, src2, 3 * PB, &full[s]);
      bulk_g2s(dst + 9 * P, src3, 3 * PB, &full[s]);
    } else {
      bulk_g2s(dst, src0, 2 * PB, &full[s]);              // abc planes 0,1
      bulk_g2s(dst + 2 * P, src1, PB, &full[s]);          // abd plane 0
      bulk_g2s(dst + 3 * P, src1 + 2 * P, PB, &full[s]);  // abd plane 2
      bulk_g2s(dst + 4 * P, src2, PB, &full[s]);          // acd plane 0
      bulk_g2s(dst + 5 * P, src2 + 2 * P, PB, &full[s]);  // acd plane 2
      bulk_g2s(dst + 6 * P, src3 + P, 2 * PB, &full[s]);  // bcd planes 1,2
    }
  };

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      done[s] = 0;
    }
    reinterpret_cast<unsigned long long*>(done + 4)[0] = 0;
    reinterpret_cast<unsigned long long*>(done + 4)[1] = 0;
    fence_mbar_init();
    for (uint32_t k = 0; k <= LA; ++k) describe(k, claim(k));
    for (int s = 0; s < S; ++s) issue(0, (uint32_t)s, s);
  }
  for (int i = tid; i < 3 * 64; i += kCountThreads) marg[i] = 0;
  if (thr_smem)
    for (int i = tid; i < prm.n_trees + 2; i += kCountThreads) s_thr[i] = (uint16_t)__ldg(&fz.thr_tab[i]);
  __syncthreads();
  auto thr = [&](uint32_t s) -> uint32_t { return thr_smem ? (uint32_t)s_thr[s] : __ldg(&fz.thr_tab[s]); };

  uint32_t o_abc = ((2 * tc) * 4 + tb) * 4 + ta;  // + 16 per kk
  uint32_t o_abd = ((2 * td) * 4 + tb) * 4 + ta;  // + 16 per ll
  uint32_t o_acd = ((2 * td) * 4 + tc) * 4 + ta;
  uint32_t o_bcd = ((2 * td) * 4 + tc) * 4 + tb;
  asm volatile("" : "+r"(o_abc), "+r"(o_abd), "+r"(o_acd), "+r"(o_bcd));
  constexpr int kPlaneU4 = kPlaneWords / 4;

  // survivors of all boxes of this CTA: [0] number, [1] sum of counts (behind the S done counters)
  unsigned long long* s_tot = reinterpret_cast<unsigned long long*>(done + 4);
  uint32_t p = 0;  // stream position of the next (box, group)

  for (uint32_t k = 0;; ++k) {
    const BoxInfo& bx = boxes[k % kBoxRing];
    if (!bx.valid) break;
    // claim the box LA + 1 ahead now, describe it after the count loop (the counter's round trip
    // stays off the critical path of warp 0)
    uint32_t claimed = 0;
    if (tid == 0) claimed = claim(k + LA + 1);

    uint32_t cnt[16][Cfg::kCnt];
#pragma unroll
    for (int c = 0; c < 16; ++c)
#pragma unroll
      for (int t = 0; t < Cfg::kCnt; ++t) cnt[c][t] = 0;

#pragma unroll 1
    for (uint32_t g = 0; g < G; ++g, ++p) {
      const int s = p % S;
      mbar_wait(&full[s], (p / S) & 1u);
      const uint4* st = reinterpret_cast<const uint4*>(smem + s * Cfg::kStageWords);
      auto ld = [&](int role, int plane, uint32_t off) -> uint4 { return st[plane_slot<kComplete>(role, plane) * kPlaneU4 + off]; };

      uint4 abc[3][2];
#pragma unroll
      for (int pl = 0; pl < (kComplete ? 2 : 3); ++pl)
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) abc[pl][kk] = ld(0, pl, o_abc + 16 * kk);
#pragma unroll
      for (int ll = 0; ll < 2; ++ll) {
        uint4 abd[3], acd[3], bcd[3];
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) {
          if (!kComplete || pl != 1) abd[pl] = ld(1, pl, o_abd + 16 * ll);
          if (!kComplete || pl != 1) acd[pl] = ld(2, pl, o_acd + 16 * ll);
          if (!kComplete || pl != 0) bcd[pl] = ld(3, pl, o_bcd + 16 * ll);
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
      // release the stage; the last warp to do so refills it with position p + S
      __syncwarp();
      if (lane == 0) {
        const uint32_t old = atom_add_acq_rel_shared(&done[s], 1u);
        if ((old & (kWarps - 1)) == kWarps - 1) issue(k, g + S, s);
      }
    }

    if (tid == 0) describe(k + LA + 1, claimed);

    // ------------------------------- epilogue of box k, on registers -------------------------------
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
    const uint32_t sa = bx.seg[0], sb = bx.seg[1], sc = bx.seg[2], sd = bx.seg[3];
    const uint32_t a0 = sa * 8 + 2 * ta, b0 = sb * 8 + 2 * tb, c0 = sc * 8 + 2 * tc, d0 = sd * 8 + 2 * td;
    unsigned long long* Zrep = fz.Z + (size_t)bx.zrep * fz.z_stride;
    uint32_t ns = 0, sum = 0;  // per thread and box: at most 48 keys, 48 x trees counts
    uint32_t* mg = marg + (k % 3) * 64;
    if (bx.regular) {
      // One junction / m-node / component per (bottom taxon, topology) for the whole box: its
      // contribution is 4 axes x 2 topologies x 8 marginal sums of the surviving weights.
      const uint32_t bg1 = ct.lcaD[(size_t)(sa * 8) * N + sb * 8], bg2 = ct.lcaD[(size_t)(sb * 8) * N + sc * 8],
                     bg3 = ct.lcaD[(size_t)(sc * 8) * N + sd * 8];
      const bool kill0 = constraint_topology(bg1, bg2, bg3) == 0;  // the constraint displays ab|cd here (else ad|bc)
      const uint32_t bit1 = kill0 ? 4u : 1u;                       // slot 1 = the one of ab|cd / ad|bc it does not display
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
              const uint32_t m = filter_survivors(c, x, fz.qmode, fz.threshold, thr, ct.tip_of_rank);
              uint32_t w0 = (m & 2u) ? c[1] : 0u, w1 = (m & bit1) ? (kill0 ? c[2] : c[0]) : 0u;
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
      const uint32_t mask_a = 0x11111111u << ta, mask_b = 0x000F000Fu << (4 * tb);
      const uint32_t mask_c = lane < 16 ? 0x0000FFFFu : 0xFFFF0000u;
#pragma unroll
      for (int sl = 0; sl < 2; ++sl)
#pragma unroll
        for (int par = 0; par < 2; ++par) {
          uint32_t v = __reduce_add_sync(mask_a, ma[sl][par]);
          if (lane < 4 && v) atomicAdd(&mg[(0 * 2 + sl) * 8 + 2 * ta + par], v);
          v = __reduce_add_sync(mask_b, mb[sl][par]);
          if (lane == 4 * tb && v) atomicAdd(&mg[(1 * 2 + sl) * 8 + 2 * tb + par], v);
          v = __reduce_add_sync(mask_c, mc[sl][par]);
          if ((lane & 15) == 0 && v) atomicAdd(&mg[(2 * 2 + sl) * 8 + 2 * tc + par], v);
          v = __reduce_add_sync(0xFFFFFFFFu, md[sl][par]);
          if (lane == 0 && v) atomicAdd(&mg[(3 * 2 + sl) * 8 + 2 * td + par], v);
        }
      __syncthreads();  // the one CTA barrier per box
      if (tid < 72) {
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
        const int pp = partner_of(t, be);
        const int pos = pp - (pp > be ? 1 : 0);
        const int comp = first ? pos : (pos + 2) % 3;
        long long v = 0;
        size_t row;
        if (tid < 64) {
          v = mg[(be * 2 + sl) * 8 + (tid & 7)];
          row = (size_t)ct.leafnode[bx.seg[be] * 8 + (tid & 7)];
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) v -= mg[(be * 2 + sl) * 8 + q];
          row = m & 0xFFFFu;
        }
        if (v) red_add_u64(Zrep + (row * (size_t)ct.n + J) * 3 + comp, (unsigned long long)v);
      }
    } else {
      // Irregular box: filter + constraint removal per cell, survivors compacted into this WARP's
      // queue and scored 32 at a time. The LCAs of the thread's 2x2 (a,b), (b,c), (c,d) pairs
      // are loaded once.
      uint32_t g1[2][2], g2[2][2], g3[2][2];  // [first][second] taxon of the pair
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const uint32_t xa = min(a0 + u, N - 1), xb = min(b0 + u, N - 1), xc = min(c0 + u, N - 1);
          const uint32_t yb = min(b0 + v, N - 1), yc = min(c0 + v, N - 1), yd = min(d0 + v, N - 1);
          g1[u][v] = ct.lcaD[(size_t)xa * N + yb];
          g2[u][v] = ct.lcaD[(size_t)xb * N + yc];
          g3[u][v] = ct.lcaD[(size_t)xc * N + yd];
        }
      using QEntry = typename std::conditional<kComplete, uint2, uint4>::type;
      QEntry* queue = reinterpret_cast<QEntry*>(sm8 + Cfg::kOffQueue) + warp * kWarpQueue;
      uint32_t qn = 0;  // warp-uniform
      auto flush = [&]() {
        flush_warp_queue<kComplete>(queue, qn, sa, sb, sc, sd, ct, Zrep, fz.as_set);
        qn = 0;
      };
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
              if (live) {
                uint32_t m = filter_survivors(c, x, fz.qmode, fz.threshold, thr, ct.tip_of_rank);
                m &= constraint_topology(g1[i][j], g2[j][kk], g3[kk][ll]) == 0 ? 6u : 3u;
                if (!(m & 1u)) c[0] = 0;
                if (!(m & 2u)) c[1] = 0;
                if (!(m & 4u)) c[2] = 0;
                live = m != 0;
              }
              const uint32_t bal = __ballot_sync(0xFFFFFFFFu, live);
              if (bal) {
                if (live) {
                  ns += (c[0] != 0) + (c[1] != 0) + (c[2] != 0);
                  sum += c[0] + c[1] + c[2];
                  const uint32_t local = cell_local(2 * ta + i, 2 * tb + j, 2 * tc + kk, 2 * td + ll);
                  const uint32_t at = qn + __popc(bal & ((1u << lane) - 1));
                  if constexpr (kComplete)
                    queue[at] = make_uint2(local | (c[0] << 12) | (c[1] << 28), (c[1] >> 4) | (c[2] << 12));
                  else
                    queue[at] = make_uint4(local, c[0], c[1], c[2]);
                }
                qn += __popc(bal);
                if (qn > kWarpQueue - 32) flush();
              }
            }
      if (qn) flush();
      __syncthreads();  // the one CTA barrier per box (keeps the warps within one box of each other)
    }
    if (tid < 64) marg[((k + 2) % 3) * 64 + tid] = 0;  // slot of box k - 1, flushed before anyone reached this barrier
    ns = __reduce_add_sync(0xFFFFFFFFu, ns);
    const unsigned long long wsum = __reduce_add_sync(0xFFFFFFFFu, sum >> 16) * 65536ull + __reduce_add_sync(0xFFFFFFFFu, sum & 0xFFFFu);
    if (lane == 0 && ns) {
      atomicAdd(&s_tot[0], (unsigned long long)ns);
      atomicAdd(&s_tot[1], wsum);
    }
  }
  __syncthreads();
  if (tid == 0) {
    if (s_tot[0]) red_add_u64(&fz.totals[0], s_tot[0]);
    if (s_tot[1]) red_add_u64(&fz.totals[1], s_tot[1]);
  }
}

}  // namespace camus_dev
