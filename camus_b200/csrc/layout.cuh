// Index math shared by host and device code: how 4-taxon sets ("cells"), their 8x8x8x8 boxes
// and the 8x8x8 triple tiles are laid out in HBM. See DESIGN.md "Data layout in HBM".
//
// Taxa are renumbered by their rank in a depth-first walk of the CONSTRAINT tree (DFS rank), not
// by gotree tip index: with that order the constraint tree never displays ac|bd on a<b<c<d, the
// LCA of two leaves is a range minimum, and subtrees are id intervals. The reference's packing
// (graphs/quartet.go:67-109, taxa sorted by tip index) is produced only at export time.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

#if defined(__CUDACC__)
#define CAMUS_HD __host__ __device__ __forceinline__
#else
#define CAMUS_HD inline
#endif

namespace camus_dev {

constexpr int kSeg = 8;                    // taxa per segment
constexpr int kTileWords = 3 * 8 * 8 * 8;  // one triple tile: 3 planes x 512 triples (u32 bit-planes)
constexpr int kTileBytes = kTileWords * 4;  // 6144
constexpr int kBoxCells = 8 * 8 * 8 * 8;   // 4096 cells per box
constexpr int kLanes = 32;                 // gene trees per bit-plane word ("tree group")

CAMUS_HD uint64_t choose2(uint64_t x) { return x < 2 ? 0 : x * (x - 1) / 2; }
CAMUS_HD uint64_t choose3(uint64_t x) { return x < 3 ? 0 : x * (x - 1) / 2 * (x - 2) / 3; }
CAMUS_HD uint64_t choose4(uint64_t x) { return x < 4 ? 0 : x * (x - 1) / 2 * (x - 2) / 3 * (x - 3) / 4; }

// tile (s0 <= s1 <= s2) and box (sa <= sb <= sc <= sd) ranks: colex order over multisets
CAMUS_HD uint64_t tile_index(uint32_t s0, uint32_t s1, uint32_t s2) {
  return choose3((uint64_t)s2 + 2) + choose2((uint64_t)s1 + 1) + s0;
}
CAMUS_HD uint64_t box_index(uint32_t sa, uint32_t sb, uint32_t sc, uint32_t sd) {
  return choose4((uint64_t)sd + 3) + choose3((uint64_t)sc + 2) + choose2((uint64_t)sb + 1) + sa;
}
CAMUS_HD uint64_t num_tiles(uint32_t m) { return choose3((uint64_t)m + 2); }
CAMUS_HD uint64_t num_boxes(uint32_t m) { return choose4((uint64_t)m + 3); }

// largest s with choose_k(s + k - 1) <= r
CAMUS_HD uint32_t unrank4(uint64_t r) {
  // fp32 estimate + exact integer fix-up (fp64 is slow on B200 and the fix-up loops are exact)
  float x = sqrtf(sqrtf(24.0f * (float)r));
  int64_t s = (int64_t)x - 2;
  if (s < 0) s = 0;
  while (choose4((uint64_t)s + 4) <= r) ++s;
  while (s > 0 && choose4((uint64_t)s + 3) > r) --s;
  return (uint32_t)s;
}
CAMUS_HD uint32_t unrank3(uint64_t r) {
  float x = cbrtf(6.0f * (float)r);
  int64_t s = (int64_t)x - 2;
  if (s < 0) s = 0;
  while (choose3((uint64_t)s + 3) <= r) ++s;
  while (s > 0 && choose3((uint64_t)s + 2) > r) --s;
  return (uint32_t)s;
}
CAMUS_HD uint32_t unrank2(uint64_t r) {
  float x = sqrtf(2.0f * (float)r);
  int64_t s = (int64_t)x - 2;
  if (s < 0) s = 0;
  while (choose2((uint64_t)s + 2) <= r) ++s;
  while (s > 0 && choose2((uint64_t)s + 1) > r) --s;
  return (uint32_t)s;
}
CAMUS_HD void box_unrank(uint64_t r, uint32_t& sa, uint32_t& sb, uint32_t& sc, uint32_t& sd) {
  sd = unrank4(r);
  r -= choose4((uint64_t)sd + 3);
  sc = unrank3(r);
  r -= choose3((uint64_t)sc + 2);
  sb = unrank2(r);
  r -= choose2((uint64_t)sb + 1);
  sa = (uint32_t)r;
}
// pair (a < b) <-> rank choose2(b) + a
CAMUS_HD void pair_unrank(uint64_t r, uint32_t& a, uint32_t& b) {
  float x = sqrtf(2.0f * (float)r);
  int64_t s = (int64_t)x - 1;
  if (s < 1) s = 1;
  while (choose2((uint64_t)s + 1) <= r) ++s;
  while (s > 1 && choose2((uint64_t)s) > r) --s;
  b = (uint32_t)s;
  a = (uint32_t)(r - choose2((uint64_t)s));
}

// Multi-rank split of the box index space, two schemes.
//
// Chunk scheme (any world): boxes are dealt to ranks in chunks, round-robin (rank r owns chunks
// r, r+world, ...), which balances the ranks (score-scatter cost varies along the index space)
// while neighbouring CTAs still share triple tiles in L2. Every rank reads every triple tile.
//
// Class scheme (world 4 or 8, see class_owner below): ranks own boxes by the CLASSES of their
// first two segments, so that a rank only reads -- and therefore only builds -- the triple tiles
// whose first segment falls in 2 (or 3) of the 4 classes. The local -> global box map is then an
// explicit list (`list`, colex order) built on the host.
//
// Local box index l of a rank <-> global box rank. Mirrored by camus_b200/partition.py.
struct BoxMap {
  uint64_t first;  // local index of blockIdx 0 of this launch
  uint64_t chunk;  // boxes per chunk
  uint32_t rank, world;
  const uint32_t* list;  // class scheme: global box rank of every local box; nullptr = chunk scheme
#if defined(__CUDACC__)
  __device__ __forceinline__ uint64_t global(uint64_t i) const {
    uint64_t l = first + i;
    if (list) return list[l];
    return ((l / chunk) * world + rank) * chunk + l % chunk;
  }
#endif
};

// ---- class scheme ----
// A box (sa <= sb <= sc <= sd) reads the tiles (sa,sb,sc), (sa,sb,sd), (sa,sc,sd), (sb,sc,sd):
// their FIRST segments are sa and sb only. Segments are coloured with 4 classes (a reflected
// pattern, so that every class sees the same mix of low and high segments); a rank that owns the
// boxes of the unordered class pair {i, j} needs the tiles led by classes i and j = half of the
// table. The 6 mixed pairs weigh 2, the 4 pure pairs (i, i) weigh 1:
//   world 8: ranks 0..5 own the mixed pairs (0,1) (0,2) (0,3) (1,2) (1,3) (2,3), rank 6 owns
//            (0,0) + (1,1), rank 7 owns (2,2) + (3,3): every rank builds 1/2 of the tiles;
//   world 4: rank 0 owns everything inside {0,1}, rank 1 everything inside {2,3} (1/2 of the
//            tiles), rank 2 owns (0,2) (0,3), rank 3 owns (1,2) (1,3) (3/4 of the tiles).
// The pure pairs are slightly heavier than half a mixed pair (they hold the sa == sb boxes); a
// fraction t/256 of them, picked by a hash of (sc, sd), is donated to ranks that need the class
// anyway (class_threshold evens the box counts out).
CAMUS_HD uint32_t seg_class(uint32_t s) {
  uint32_t r = s & 7u;
  return r < 4 ? r : 7 - r;
}
CAMUS_HD int class_pair_rank(uint32_t lo, uint32_t hi) { return lo == 0 ? (int)hi - 1 : (lo == 1 ? (int)hi + 1 : 5); }
CAMUS_HD uint32_t class_hash(uint32_t sc, uint32_t sd) { return (sc * 7u + sd * 13u) & 255u; }
CAMUS_HD int class_owner(uint32_t ca, uint32_t cb, uint32_t h, int world, uint32_t t) {
  const uint32_t lo = ca < cb ? ca : cb, hi = ca < cb ? cb : ca;
  if (world == 8) {
    if (lo != hi) return class_pair_rank(lo, hi);
    if (h < 3 * t) {
      const uint32_t k = h / t, o = k < lo ? k : k + 1;  // k-th of the three other classes
      return class_pair_rank(o < lo ? o : lo, o < lo ? lo : o);
    }
    return lo < 2 ? 6 : 7;
  }
  // world == 4
  if (lo != hi) return hi < 2 ? 0 : (lo >= 2 ? 1 : (lo == 0 ? 2 : 3));
  if (h < t) return (lo == 0 || lo == 2) ? 2 : 3;
  return lo < 2 ? 0 : 1;
}
// classes (bit mask) whose tiles rank `rank` reads
CAMUS_HD uint32_t class_need_mask(int rank, int world) {
  if (world == 8) {
    const uint32_t m[8] = {0x3, 0x5, 0x9, 0x6, 0xA, 0xC, 0x3, 0xC};
    return m[rank & 7];
  }
  const uint32_t m[4] = {0x3, 0xC, 0xD, 0xE};
  return m[rank & 3];
}
// donation threshold (of 256) for m segments. World 8: every rank builds half of the tiles, so
// the box counts are evened out. World 4: ranks 2 and 3 build 3/4 of the tiles against 1/2 for
// ranks 0 and 1; with the bit-sliced plane builder a tile costs about half of what a box costs to
// count (per tree group at 1000 taxa: 1.2 ns vs 2.45 ns with the complete-tree kernels; the fp16
// builder of round 1 cost 3.2 ns and the split then was T/4), so ranks 2 and 3 are given T/8 boxes
// FEWER than ranks 0 and 1, T = number of tiles.
inline uint32_t class_threshold(uint32_t m, int world) {
  uint64_t D = 0, O = 0;  // boxes of pure / mixed class pairs
  for (uint32_t sa = 0; sa < m; ++sa)
    for (uint32_t sb = sa; sb < m; ++sb) {
      uint64_t k = (uint64_t)(m - sb) * (m - sb + 1) / 2;
      (seg_class(sa) == seg_class(sb) ? D : O) += k;
    }
  if (D == 0 || 3 * D <= O) return 0;
  if (world == 8) {
    uint64_t t = (64 * (3 * D - O) + (3 * D) / 2) / (3 * D);
    return (uint32_t)(t > 85 ? 85 : t);
  }
  // world 4: boxes(rank 0) - boxes(rank 2) = D/2 - O/6 - (t/256) * 3D/4, wanted = T/8
  const uint64_t T = num_tiles(m);
  const uint64_t have6 = 3 * D - O, want6 = 3 * T / 4;  // both sides times 6
  if (have6 <= want6) return 0;
  uint64_t t = (256 * 2 * (have6 - want6) + (9 * D) / 2) / (9 * D);
  return (uint32_t)(t > 255 ? 255 : t);
}
CAMUS_HD uint64_t partition_chunk(uint64_t nb, uint32_t world) {
  uint64_t c = nb / ((uint64_t)world * 64);
  return c < 1 ? 1 : (c > 2048 ? 2048 : c);
}
CAMUS_HD uint64_t local_box_count(uint64_t nb, uint64_t chunk, uint32_t rank, uint32_t world) {
  uint64_t round = chunk * world, full = nb / round, rem = nb % round;
  uint64_t lo = (uint64_t)rank * chunk;
  uint64_t extra = rem > lo ? (rem - lo < chunk ? rem - lo : chunk) : 0;
  return full * chunk + extra;
}

// Word offset inside a triple tile for plane p and local triple (x, y, z), x/y/z in 0..7 being
// the positions of the tile's three segments. The (x pair, y pair) 2x2 micro-block is contiguous
// so that one 128-bit shared load fetches what a count thread needs (kernels_count.cuh).
CAMUS_HD uint32_t tile_word(uint32_t p, uint32_t x, uint32_t y, uint32_t z) {
  return (((p * 8 + z) * 4 + (y >> 1)) * 4 + (x >> 1)) * 4 + (y & 1) * 2 + (x & 1);
}
// planes of triple (p < q < r): 0 = pq|r, 1 = pr|q, 2 = qr|p

// Bit-sliced pair rows of the sliced plane builder (kernels_prep.cuh): NB depth-bit slices + one presence word,
// padded to 128 bits.
template <int NB>
struct SliceRow {
  static constexpr int W = NB < 8 ? 8 : 12;  // words per pair row in HBM (NB slices + presence, padded to 128 bits)
};
// word offset of 16-byte chunk q of row r inside one 64-row shared array. 8-word rows: chunk (2r + q) ^ swz(r) with
// swz from row bits 2, 4, 5 -- conflict-free for 8 consecutive rows (staging), rows 2 apart and rows 16 / 32 apart
// (reads), checked by tests/test_layout_mirror.py. 12-word rows: padded by 4 words every 16 rows.
template <int W>
CAMUS_HD int slice_off(int r, int q) {
  if (W == 8) return (((2 * r + q) ^ (((r >> 2) & 1) | (((r >> 4) & 1) << 1) | (((r >> 5) & 1) << 2))) << 2);
  return r * 12 + (r >> 4) * 4 + 4 * q;
}
template <int W>
struct SliceArray {
  static constexpr int words = W == 8 ? 64 * 8 : 64 * 12 + 16;
};

// One entry of the sliced plane builder's work list (kernels_prep.cuh k_triple_planes_bs).
struct PlaneDesc {
  uint32_t s01;   // s0 | s1 << 16
  uint32_t s2;
  uint32_t slot;  // where the tile lives in the bit-plane table (tile index, or the compact slot of a class-scheme share)
  uint32_t pad;
};
// The work list: the tiles (s0 <= s1 <= s2) of m segments whose first segment's class is in `need` (0xF = all),
// in blocks of kPlaneBlock^3 segments -- a block of tiles shares 3 x 128 x 128 pair rows, which stay in L2 while
// its tiles stream out -- colex inside a block. compact: slots number the listed tiles in colex order (the order
// camus_b200.cu set_box_range stores a class-scheme share's tiles in); else slot = tile index.
constexpr uint32_t kPlaneBlock = 16;
inline std::vector<PlaneDesc> plane_work_list(uint32_t m, uint32_t need, bool compact) {
  std::vector<PlaneDesc> list;
  list.reserve((size_t)num_tiles(m));
  std::vector<uint32_t> slot_of;
  if (compact) {
    slot_of.assign((size_t)num_tiles(m), 0xFFFFFFFFu);
    uint32_t idx = 0, next = 0;
    for (uint32_t s2 = 0; s2 < m; ++s2)
      for (uint32_t s1 = 0; s1 <= s2; ++s1)
        for (uint32_t s0 = 0; s0 <= s1; ++s0, ++idx)
          if ((need >> seg_class(s0)) & 1u) slot_of[idx] = next++;
  }
  const uint32_t B = kPlaneBlock, nblk = (m + B - 1) / B;
  auto upto = [](uint32_t a, uint32_t b) { return a < b ? a : b; };
  for (uint32_t B2 = 0; B2 < nblk; ++B2)
    for (uint32_t B1 = 0; B1 <= B2; ++B1)
      for (uint32_t B0 = 0; B0 <= B1; ++B0)
        for (uint32_t s2 = B2 * B; s2 < upto(m, (B2 + 1) * B); ++s2)
          for (uint32_t s1 = B1 * B; s1 < upto(s2 + 1, (B1 + 1) * B); ++s1)
            for (uint32_t s0 = B0 * B; s0 < upto(s1 + 1, (B0 + 1) * B); ++s0) {
              if (!((need >> seg_class(s0)) & 1u)) continue;
              const uint64_t idx = tile_index(s0, s1, s2);
              list.push_back(PlaneDesc{s0 | (s1 << 16), s2, compact ? slot_of[idx] : (uint32_t)idx, 0u});
            }
  return list;
}

// cell inside a box, a fastest
CAMUS_HD uint32_t cell_local(uint32_t al, uint32_t bl, uint32_t cl, uint32_t dl) {
  return ((dl * 8 + cl) * 8 + bl) * 8 + al;
}

}  // namespace camus_dev
