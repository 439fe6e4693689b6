// DP offload: the across-scan of the level-1 network DP on the GPU.
//
// Reference: infer/main_dp.go:178-216 (scoreAddEdgeK) visits every u of one child subtree of v in
// post-order and, for each, every w of the other child subtree in preorder (scoreEdgesAcross,
// main_dp.go:247-287); per (u, w) it runs FourWayBestSplit (infer/helpers.go:83-110) over the four
// lists cycleDP.scores[w], cycleDP.scores[u], DP[w], DP[u] -- O(k^2) additions per pair and step, which
// is where the DP's time goes (float score modes at 1000 taxa: minutes on the host).
//
// Here one thread evaluates one (u, w) pair LITERALLY: the same sums in the same order, the same
// first-maximum tie-breaks (BestSplit, helpers.go:51-72: `cur > best || !found` scanning the left index
// upwards), the final score as edge + cs_w + cs_u + dp_w + dp_u left to right (main_dp.go:268), so
// float (norm / sym) scores are bit-identical to Go's. The only shortcut: c1[i] (the best split of
// cs_w / cs_u into i edges) is computed just for the i the last BestSplit can reach.
//
// Reduction, equal to the reference's sequential acceptance when no score is NaN:
//   inside one u : first maximum over w in preorder (`score > bestScore || traceback == nil`);
//   across the u : maximum score, then the smallest cycle length, then the LAST u in visiting order
//                  (`curScore == bestScore && cycleLen <= bestCycleLen`, main_dp.go:206).
// NaN scores are real in norm mode (0 / 0 for the length-3 cycle between the two children of v,
// which scoreEdgesAcross does not exclude) and the reference's comparisons give them a definite fate:
// inside one u a NaN is kept only when it belongs to the FIRST w (then nothing replaces it), across
// the u a NaN candidate is kept only when nothing was accepted before it (then nothing replaces it
// either). So the fold above runs over the non-NaN candidates, and item 0's candidate is returned
// separately when it is NaN: the caller takes it iff it has no candidate of its own (scoreEdgesDown).
//
// The DP driver (post-order over v, the k loop, cycleDP.update, traceback) stays with the caller:
// camus_b200/host/dp.hpp here, unchanged Go in the drop-in.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

#include "camus_b200.h"

namespace {

// Device buffers come from the stream-ordered pool (cudaMallocAsync / cudaFreeAsync on the null stream, release
// threshold raised once): a DP handle is created and destroyed per inference, and plain cudaFree -- a device-wide
// synchronisation plus an unmap -- took 0.5 - 1 s per handle inside a process that holds tens of GB.
inline void pool_keep_memory() {
  static bool done = false;
  if (done) return;
  done = true;
  int dev = 0;
  cudaMemPool_t pool;
  if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  cudaGetLastError();
}
template <class T>
struct Buf {
  T* p = nullptr;
  size_t n = 0;
  ~Buf() { release(); }
  void release() {
    if (p && cudaFreeAsync(p, nullptr) != cudaSuccess) {
      cudaGetLastError();
      cudaFree(p);
    }
    p = nullptr;
    n = 0;
  }
  cudaError_t alloc(size_t count) {
    if (count <= n && p) return cudaSuccess;
    if (p) cudaDeviceSynchronize();  // regrowth (rare: the tables are pre-sized): work in flight may still read the old block
    release();
    if (count == 0) return cudaSuccess;
    pool_keep_memory();
    cudaError_t e = cudaMallocAsync(&p, count * sizeof(T), nullptr);
    if (e == cudaSuccess) e = cudaStreamSynchronize(nullptr);  // the handle's own stream does not wait on the null stream
    if (e == cudaSuccess) n = count;
    else p = nullptr;
    return e;
  }
};

// pinned host buffers are parked at destroy and handed to the next handle (cudaMallocHost / cudaFreeHost cost milliseconds)
struct PinnedCache {
  std::mutex mu;
  std::vector<std::pair<void*, size_t>> free_list;
  void* take(size_t bytes, size_t* got) {
    {
      std::lock_guard<std::mutex> lk(mu);
      for (size_t i = 0; i < free_list.size(); ++i)
        if (free_list[i].second >= bytes) {
          void* p = free_list[i].first;
          *got = free_list[i].second;
          free_list.erase(free_list.begin() + i);
          return p;
        }
    }
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    *got = bytes;
    return p;
  }
  void give(void* p, size_t bytes) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(mu);
    if (free_list.size() < 8) free_list.push_back({p, bytes});
    else cudaFreeHost(p);
  }
};
inline PinnedCache& pinned_cache() {
  static PinnedCache c;
  return c;
}

struct ItemBest {
  unsigned long long score;  // raw 8 bytes (u64 or f64)
  int ok, u, w, len;
  int idx[4];
};

struct StepOut {
  ItemBest best;   // fold over the items whose score is not NaN
  int item;
  int nan_seen;    // statistic: some pair scored NaN
  ItemBest first;  // item 0 when its score is NaN (first.ok), see the header comment
};

template <class S>
__device__ __forceinline__ S as_score(unsigned long long raw) {
  if constexpr (sizeof(S) == 8 && !std::is_integral<S>::value) return __longlong_as_double((long long)raw);
  else return (S)raw;
}
template <class S>
__device__ __forceinline__ unsigned long long raw_of(S v) {
  if constexpr (sizeof(S) == 8 && !std::is_integral<S>::value) return (unsigned long long)__double_as_longlong(v);
  else return (unsigned long long)v;
}
template <class S>
__device__ __forceinline__ bool is_nan(S v) {
  if constexpr (std::is_integral<S>::value) return false;
  else return v != v;
}

// pending DP[node] lists -> node-major table
__global__ void k_dp_scatter(int n_pending, const int* __restrict__ desc /*[3*n_pending]: node, len, offset*/,
                             const unsigned long long* __restrict__ blob, unsigned long long* dp, int* dplen, int dcap) {
  const int p = blockIdx.x;
  if (p >= n_pending) return;
  const int node = desc[3 * p], len = desc[3 * p + 1], off = desc[3 * p + 2];
  for (int j = threadIdx.x; j < len; j += blockDim.x) dp[(size_t)node * dcap + j] = blob[off + j];
  if (threadIdx.x == 0) dplen[node] = len;
}

__global__ void k_dp_repack(const unsigned long long* __restrict__ src, unsigned long long* dst, int n, int old_cap, int new_cap) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * old_cap) return;
  dst[(i / old_cap) * new_cap + i % old_cap] = src[i];
}

// path columns [k0, k1) of nodes [first, first + count), as uploaded -> node-major table
__global__ void k_cs_scatter(const unsigned long long* __restrict__ cols, unsigned long long* cs, int kcap, int k0, int nk, int first, int count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nk * count) return;
  const int jj = i / count, x = i % count;
  cs[(size_t)(first + x) * kcap + k0 + jj] = cols[i];
}

constexpr int kAcrossThreads = 128;

// One block per (item, chunk of 128 consecutive w): one thread evaluates one (u, w) pair, so the work
// of a heavy u (long DP[u]: (K_u + K_w) * k additions per pair) is spread over many blocks.
// cs: [n][kcap], dp: [n][dcap] (node-major: a thread walks its own rows).
struct ChunkBest {
  unsigned long long score;
  int ok, w, stuck;  // stuck: the chunk holds the item's FIRST w and its score is NaN
  int idx[4];
};

template <class S>
__global__ void __launch_bounds__(kAcrossThreads)
k_dp_across(int n_items, int n_blocks, const int* __restrict__ items /*[2*n_items]: u, sub*/, const int* __restrict__ blk_item,
            const int* __restrict__ item_blk0 /*[n_items+1]*/, const int* __restrict__ sub_end, const int* __restrict__ depth, int v, int k, int n,
            int kcap, int dcap, const unsigned long long* __restrict__ cs_raw, const unsigned long long* __restrict__ dp_raw,
            const int* __restrict__ dplen, const unsigned long long* __restrict__ edge_raw, ChunkBest* chunk_best, StepOut* out) {
  extern __shared__ unsigned long long sh_raw[];
  S* B = reinterpret_cast<S*>(sh_raw);  // cs_u[0..k]
  S* D = B + (k + 1);                   // dp_u[0..Ld)
  __shared__ unsigned long long r_score[kAcrossThreads];
  __shared__ int r_w[kAcrossThreads], r_ok[kAcrossThreads];
  __shared__ int r_idx[kAcrossThreads][4];
  __shared__ int s_stuck;
  const S* cs = reinterpret_cast<const S*>(cs_raw);
  const S* dp = reinterpret_cast<const S*>(dp_raw);
  const S* edge = reinterpret_cast<const S*>(edge_raw);
  const int tid = threadIdx.x;
  const int it = blk_item[blockIdx.x];
  const int chunk = blockIdx.x - item_blk0[it];
  const int u = items[2 * it], sub = items[2 * it + 1];
  const int end = sub_end[sub];
  const int Ld = dplen[u];
  for (int j = tid; j <= k; j += kAcrossThreads) B[j] = cs[(size_t)u * kcap + j];
  for (int j = tid; j < Ld; j += kAcrossThreads) D[j] = dp[(size_t)u * dcap + j];
  if (tid == 0) s_stuck = 0;
  __syncthreads();

  bool ok = false;
  S best{};
  int bidx[4] = {0, 0, 0, 0};
  const int w = sub + chunk * kAcrossThreads + tid;
  if (w < end) {
    const S* A = cs + (size_t)w * kcap;
    const S* C = dp + (size_t)w * dcap;
    const int Lc = dplen[w];
    // FourWayBestSplit (helpers.go:83-110): lists {cs_w, cs_u, dp_w, dp_u}, both cs lists hold k + 1 entries
    const int m2 = min(k, Lc + Ld - 2) + 1;  // entries SolveAllSplits(dp_w, dp_u) produces
    bool found = false;
    S top{};
    for (int i = k - (m2 - 1); i <= k; ++i) {
      S c1 = A[0] + B[i];  // BestSplit(cs_w, cs_u, i): left index 0..i
      int a1 = 0;
      for (int j = 1; j <= i; ++j) {
        const S cur = A[j] + B[i - j];
        if (cur > c1) {
          c1 = cur;
          a1 = j;
        }
      }
      const int i2 = k - i;  // BestSplit(dp_w, dp_u, i2)
      const int lo2 = max(0, i2 - (Ld - 1)), hi2 = min(i2, Lc - 1);
      S c2 = C[lo2] + D[i2 - lo2];
      int cc = lo2;
      for (int j = lo2 + 1; j <= hi2; ++j) {
        const S cur = C[j] + D[i2 - j];
        if (cur > c2) {
          c2 = cur;
          cc = j;
        }
      }
      const S cur = c1 + c2;
      if (!found || cur > top) {
        top = cur;
        bidx[0] = a1;
        bidx[1] = i - a1;
        bidx[2] = cc;
        bidx[3] = i2 - cc;
        found = true;
      }
    }
    S score = edge[(size_t)u * n + w];  // main_dp.go:268, left to right
    score = score + A[bidx[0]];
    score = score + B[bidx[1]];
    score = score + C[bidx[2]];
    score = score + D[bidx[3]];
    best = score;
    ok = true;
    if (is_nan(score)) {
      atomicOr(&out->nan_seen, 1);
      // `score > bestScore` is false for a NaN: only the item's first w keeps one (and is then never replaced)
      if (w == sub) s_stuck = 1;
      else ok = false;
    }
  }
  r_score[tid] = raw_of<S>(best);
  r_w[tid] = w;
  r_ok[tid] = ok;
  for (int q = 0; q < 4; ++q) r_idx[tid][q] = bidx[q];
  __syncthreads();
  // first maximum over w: larger score, or equal score and smaller w; a stuck first w wins outright
  for (int s = kAcrossThreads / 2; s > 0; s >>= 1) {
    if (tid < s && r_ok[tid + s] && !s_stuck) {
      const S a = as_score<S>(r_score[tid]), b = as_score<S>(r_score[tid + s]);
      if (!r_ok[tid] || b > a || (b == a && r_w[tid + s] < r_w[tid])) {
        r_score[tid] = r_score[tid + s];
        r_w[tid] = r_w[tid + s];
        r_ok[tid] = 1;
        for (int q = 0; q < 4; ++q) r_idx[tid][q] = r_idx[tid + s][q];
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    ChunkBest cb;
    cb.score = r_score[0];
    cb.ok = r_ok[0];
    cb.w = r_w[0];
    cb.stuck = s_stuck;
    for (int q = 0; q < 4; ++q) cb.idx[q] = r_idx[0][q];
    chunk_best[blockIdx.x] = cb;
  }
}

// Tiled form of k_dp_across: one block = up to kTileU consecutive items that scan the SAME other subtree x one
// chunk of kTileW consecutive w. The cs / DP rows of the chunk's w (read k times per pair) are staged once
// in shared memory, transposed so that a warp's lanes read consecutive words, and shared by the kTileU items:
// the one-thread-per-pair kernel above re-reads them from L1 / L2 with a 32-way scattered access per load,
// which is what bounded it (0.5 ms per step at the root of a 1000-taxon constraint).
constexpr int kTileW = 64, kTileU = 8, kTileLd = kTileW + 1, kTabInts = 4 + kTileU;

template <class S>
__global__ void __launch_bounds__(kTileW* kTileU)
k_dp_across_tiled(const int* __restrict__ items, const int* __restrict__ blk_tab /*[kTabInts*n_blocks]: items in the tile, chunk, offset into perm, other subtree, then the item ids*/,
                  const int* __restrict__ perm /*the other subtree's nodes, longest DP list first*/, const int* __restrict__ item_cb0,
                  const int* __restrict__ sub_end, int k, int n, int kcap, int dcap, int max_len,
                  const unsigned long long* __restrict__ cs_raw, const unsigned long long* __restrict__ dp_raw, const int* __restrict__ dplen,
                  const unsigned long long* __restrict__ edge_raw, ChunkBest* chunk_best, StepOut* out) {
  extern __shared__ unsigned long long sh_raw[];
  S* As = reinterpret_cast<S*>(sh_raw);         // [k + 1][kTileLd]
  S* Cs = As + (size_t)(k + 1) * kTileLd;       // [max_len][kTileLd]
  S* Bs = Cs + (size_t)max_len * kTileLd;       // [kTileU][k + 1]
  S* Ds = Bs + (size_t)kTileU * (k + 1);        // [kTileU][max_len]
  unsigned long long* r_score = reinterpret_cast<unsigned long long*>(Ds + (size_t)kTileU * max_len);  // [kTileU][kTileW]
  int* r_w = reinterpret_cast<int*>(r_score + kTileU * kTileW);
  int* r_ok = r_w + kTileU * kTileW;
  int* r_idx = r_ok + kTileU * kTileW;  // [kTileU][kTileW][4]
  __shared__ int s_ld[kTileU], s_stuck[kTileU];
  const S* cs = reinterpret_cast<const S*>(cs_raw);
  const S* dp = reinterpret_cast<const S*>(dp_raw);
  const S* edge = reinterpret_cast<const S*>(edge_raw);
  const int wt = threadIdx.x, ui = threadIdx.y, tid = ui * kTileW + wt, nthr = kTileW * kTileU;
  // the tile's items need not be consecutive: they are grouped by the length of DP[u], so that the rows of a block
  // finish together (ncu of the consecutive form: barrier was the top stall, the light rows waiting for the heavy one)
  const int* tab = blk_tab + (size_t)kTabInts * blockIdx.x;
  const int n_it = tab[0], chunk = tab[1];
  const int sub = tab[3], end = sub_end[sub];
  const int* tile_items = tab + 4;
  // The w of a chunk come from a list sorted by the length of DP[w]: the work of a pair grows with that length, and
  // 32 consecutive preorder ids mix leaves with their heavy ancestors (ncu: 5x more warp instructions than useful ones)
  const int* pw = perm + tab[2] + chunk * kTileW;
  const int n_w = min(kTileW, (end - sub) - chunk * kTileW);  // valid entries of pw
  // stage: the rows are node-major in HBM, so consecutive threads read consecutive j of one row (coalesced;
  // reading them column-wise cost 4x the bytes in 32-byte sectors and bounded the kernel); the transposed
  // shared rows are padded to kTileLd so that those stores spread over the banks
  for (int idx = tid; idx < (k + 1) * kTileW; idx += nthr) {
    const int r = idx / (k + 1), j = idx % (k + 1);
    As[j * kTileLd + r] = r < n_w ? cs[(size_t)pw[r] * kcap + j] : S{};
  }
  for (int idx = tid; idx < max_len * kTileW; idx += nthr) {
    const int r = idx / max_len, j = idx % max_len;
    Cs[j * kTileLd + r] = r < n_w ? dp[(size_t)pw[r] * dcap + j] : S{};
  }
  for (int idx = tid; idx < kTileU * (k + 1); idx += nthr) {
    const int q = idx / (k + 1), j = idx % (k + 1);
    Bs[idx] = q < n_it ? cs[(size_t)items[2 * tile_items[q]] * kcap + j] : S{};
  }
  for (int idx = tid; idx < kTileU * max_len; idx += nthr) {
    const int q = idx / max_len, j = idx % max_len;
    Ds[idx] = q < n_it ? dp[(size_t)items[2 * tile_items[q]] * dcap + j] : S{};
  }
  if (tid < kTileU) {
    s_ld[tid] = tid < n_it ? dplen[items[2 * tile_items[tid]]] : 1;
    s_stuck[tid] = 0;
  }
  __syncthreads();

  const int w = wt < n_w ? pw[wt] : 0x7FFFFFFF;
  bool ok = false;
  S best{};
  int bidx[4] = {0, 0, 0, 0};
  if (ui < n_it && wt < n_w) {
    const int u = items[2 * tile_items[ui]];
    const S* A = As + wt;  // A[j] = A[j * kTileW]
    const S* C = Cs + wt;
    const S* B = Bs + (size_t)ui * (k + 1);
    const S* D = Ds + (size_t)ui * max_len;
    const int Lc = dplen[w], Ld = s_ld[ui];
    const int m2 = min(k, Lc + Ld - 2) + 1;
    bool found = false;
    S top{};
    for (int i = k - (m2 - 1); i <= k; ++i) {
      S c1 = A[0] + B[i];
      int a1 = 0;
      for (int j = 1; j <= i; ++j) {
        const S cur = A[j * kTileLd] + B[i - j];
        if (cur > c1) {
          c1 = cur;
          a1 = j;
        }
      }
      const int i2 = k - i;
      const int lo2 = max(0, i2 - (Ld - 1)), hi2 = min(i2, Lc - 1);
      S c2 = C[lo2 * kTileLd] + D[i2 - lo2];
      int cc = lo2;
      for (int j = lo2 + 1; j <= hi2; ++j) {
        const S cur = C[j * kTileLd] + D[i2 - j];
        if (cur > c2) {
          c2 = cur;
          cc = j;
        }
      }
      const S cur = c1 + c2;
      if (!found || cur > top) {
        top = cur;
        bidx[0] = a1;
        bidx[1] = i - a1;
        bidx[2] = cc;
        bidx[3] = i2 - cc;
        found = true;
      }
    }
    S score = edge[(size_t)u * n + w];  // main_dp.go:268, left to right
    score = score + A[bidx[0] * kTileLd];
    score = score + B[bidx[1]];
    score = score + C[bidx[2] * kTileLd];
    score = score + D[bidx[3]];
    best = score;
    ok = true;
    if (is_nan(score)) {
      atomicOr(&out->nan_seen, 1);
      if (w == sub) s_stuck[ui] = 1 + wt;  // the item's first w in preorder: its NaN stays (any lane: the w are sorted by DP length)
      else ok = false;
    }
  }
  r_score[tid] = raw_of<S>(best);
  r_w[tid] = w;
  r_ok[tid] = ok;
  for (int q = 0; q < 4; ++q) r_idx[tid * 4 + q] = bidx[q];
  __syncthreads();
  for (int s = kTileW / 2; s > 0; s >>= 1) {
    if (wt < s && r_ok[tid + s] && !s_stuck[ui]) {
      const S a = as_score<S>(r_score[tid]), b = as_score<S>(r_score[tid + s]);
      if (!r_ok[tid] || b > a || (b == a && r_w[tid + s] < r_w[tid])) {
        r_score[tid] = r_score[tid + s];
        r_w[tid] = r_w[tid + s];
        r_ok[tid] = 1;
        for (int q = 0; q < 4; ++q) r_idx[tid * 4 + q] = r_idx[(tid + s) * 4 + q];
      }
    }
    __syncthreads();
  }
  if (wt == 0 && ui < n_it) {
    const int src = s_stuck[ui] ? ui * kTileW + s_stuck[ui] - 1 : tid;  // a stuck first w wins outright (nothing was merged)
    ChunkBest cb;
    cb.score = r_score[src];
    cb.ok = r_ok[src];
    cb.w = r_w[src];
    cb.stuck = s_stuck[ui] != 0;
    for (int q = 0; q < 4; ++q) cb.idx[q] = r_idx[src * 4 + q];
    chunk_best[item_cb0[tile_items[ui]] + chunk] = cb;
  }
}

// Integer scores only (max mode): the across-scan as a max-plus product. FourWayBestSplit maximises
// cs_w[a] + cs_u[b] + dp_w[c] + dp_u[d] over a + b + c + d = k; max-plus convolution is associative and exact on
// integers, so the same maximum is max_j conv_w[j] + conv_u[k - j] with the per-node conv_x = cs_x (*) dp_x the caller
// uploads (camus_b200_dp_set_conv_columns): O(k) per pair instead of O(k^2). Only the VALUE comes from this form; the
// caller re-derives the split indices of the one candidate a step ends with by the reference's own FourWayBestSplit
// (split = -1 in the result), so tie-breaks are untouched. Same tiling, same reduction as k_dp_across_tiled.
__global__ void __launch_bounds__(kTileW* kTileU)
k_dp_across_maxplus(const int* __restrict__ items, const int* __restrict__ blk_tab, const int* __restrict__ perm, const int* __restrict__ item_cb0,
                    const int* __restrict__ sub_end, int k, int n, int kcap, const unsigned long long* __restrict__ cv,
                    const unsigned long long* __restrict__ edge, ChunkBest* chunk_best) {
  extern __shared__ unsigned long long sh_raw[];
  unsigned long long* As = sh_raw;                       // [k + 1][kTileLd]: conv of the chunk's w
  unsigned long long* Bs = As + (size_t)(k + 1) * kTileLd;  // [kTileU][k + 1]: conv of the tile's u
  unsigned long long* r_score = Bs + (size_t)kTileU * (k + 1);
  int* r_w = reinterpret_cast<int*>(r_score + kTileU * kTileW);
  int* r_ok = r_w + kTileU * kTileW;
  const int wt = threadIdx.x, ui = threadIdx.y, tid = ui * kTileW + wt, nthr = kTileW * kTileU;
  const int* tab = blk_tab + (size_t)kTabInts * blockIdx.x;
  const int n_it = tab[0], chunk = tab[1];
  const int sub = tab[3], end = sub_end[sub];
  const int* tile_items = tab + 4;
  const int* pw = perm + tab[2] + chunk * kTileW;
  const int n_w = min(kTileW, (end - sub) - chunk * kTileW);
  for (int idx = tid; idx < (k + 1) * kTileW; idx += nthr) {
    const int r = idx / (k + 1), j = idx % (k + 1);
    As[j * kTileLd + r] = r < n_w ? cv[(size_t)pw[r] * kcap + j] : 0ull;
  }
  for (int idx = tid; idx < kTileU * (k + 1); idx += nthr) {
    const int q = idx / (k + 1), j = idx % (k + 1);
    Bs[idx] = q < n_it ? cv[(size_t)items[2 * tile_items[q]] * kcap + j] : 0ull;
  }
  __syncthreads();
  const int w = wt < n_w ? pw[wt] : 0x7FFFFFFF;
  bool ok = false;
  unsigned long long best = 0;
  if (ui < n_it && wt < n_w) {
    const int u = items[2 * tile_items[ui]];
    const unsigned long long* A = As + wt;
    const unsigned long long* B = Bs + (size_t)ui * (k + 1);
    unsigned long long val = 0;
    for (int j = 0; j <= k; ++j) val = max(val, A[j * kTileLd] + B[k - j]);
    best = edge[(size_t)u * n + w] + val;
    ok = true;
  }
  r_score[tid] = best;
  r_w[tid] = w;
  r_ok[tid] = ok;
  __syncthreads();
  for (int s = kTileW / 2; s > 0; s >>= 1) {
    if (wt < s && r_ok[tid + s]) {
      const unsigned long long a = r_score[tid], b = r_score[tid + s];
      if (!r_ok[tid] || b > a || (b == a && r_w[tid + s] < r_w[tid])) {
        r_score[tid] = b;
        r_w[tid] = r_w[tid + s];
        r_ok[tid] = 1;
      }
    }
    __syncthreads();
  }
  if (wt == 0 && ui < n_it) {
    ChunkBest cb;
    cb.score = r_score[tid];
    cb.ok = r_ok[tid];
    cb.w = r_w[tid];
    cb.stuck = 0;
    for (int q = 0; q < 4; ++q) cb.idx[q] = -1;  // the caller derives the split indices of the winner
    chunk_best[item_cb0[tile_items[ui]] + chunk] = cb;
  }
}

// Fold of one step (one block): per item its chunks in w order (first maximum; a stuck chunk 0 ends
// it); across the items max score, then min cycle length, then the last item; NaN candidates never
// win, item 0's is reported separately.
constexpr int kFoldThreads = 1024;
template <class S>
__global__ void __launch_bounds__(kFoldThreads)
k_dp_fold(int n_items, const int* __restrict__ items, const int* __restrict__ item_blk0, const int* __restrict__ depth, int v,
          const ChunkBest* __restrict__ chunk_best, StepOut* out) {
  const int tid = threadIdx.x;
  int mine = -1;
  ItemBest mb{};
  for (int i = tid; i < n_items; i += kFoldThreads) {
    const int b0 = item_blk0[i], b1 = item_blk0[i + 1];
    ChunkBest c = chunk_best[b0];
    for (int b = b0 + 1; b < b1 && !c.stuck; ++b) {  // chunks in any w order: first maximum = larger score, then smaller w
      const ChunkBest d = chunk_best[b];
      if (d.stuck) {
        c = d;  // the item's first w scored NaN: nothing replaces it
      } else if (d.ok) {
        const S a = as_score<S>(c.score), e = as_score<S>(d.score);
        if (!c.ok || e > a || (e == a && d.w < c.w)) c = d;
      }
    }
    if (!c.ok) continue;
    ItemBest ib;
    ib.score = c.score;
    ib.ok = 1;
    ib.u = items[2 * i];
    ib.w = c.w;
    ib.len = (depth[ib.u] - depth[v]) + (depth[ib.w] - depth[v]) + 1;  // CycleLength, u not above w
    for (int q = 0; q < 4; ++q) ib.idx[q] = c.idx[q];
    if (is_nan(as_score<S>(ib.score))) {
      if (i == 0) out->first = ib;
      continue;
    }
    bool take = mine < 0;
    if (!take) {
      const S a = as_score<S>(mb.score), b = as_score<S>(ib.score);
      take = b > a || (b == a && ib.len <= mb.len);  // i increases: equal length -> the later item
    }
    if (take) {
      mb = ib;
      mine = i;
    }
  }
  __shared__ ItemBest f_best[kFoldThreads];
  __shared__ int f_item[kFoldThreads];
  f_best[tid] = mb;
  f_item[tid] = mine;
  __syncthreads();
  for (int s = kFoldThreads / 2; s > 0; s >>= 1) {
    if (tid < s && f_item[tid + s] >= 0) {
      bool take = f_item[tid] < 0;
      if (!take) {
        const S a = as_score<S>(f_best[tid].score), b = as_score<S>(f_best[tid + s].score);
        take = b > a || (b == a && (f_best[tid + s].len < f_best[tid].len ||
                                    (f_best[tid + s].len == f_best[tid].len && f_item[tid + s] > f_item[tid])));
      }
      if (take) {
        f_best[tid] = f_best[tid + s];
        f_item[tid] = f_item[tid + s];
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    out->best = f_best[0];
    out->item = f_item[0];
    if (f_item[0] < 0) out->best.ok = 0;
  }
}

}  // namespace

struct camus_b200_dp {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool f64 = false;
  int n = 0;
  int kcap = 0, dcap = 0;
  Buf<int> d_end, d_depth, d_dplen, d_items, d_desc, d_blk_item, d_item_blk0;
  Buf<ChunkBest> d_chunk_best;
  Buf<int> d_tile_tab, d_item_cb0, d_perm;
  std::vector<int> h_dplen;  // host copy of the DP list lengths (set_node)  // tiled kernel: (first item, items, chunk) per block; first chunk slot per item
  int n_tile_blocks = 0, max_len = 1;
  bool tiled_ok = false;  // the opt-in shared-memory size was granted
  int n_blocks = 0;
  std::vector<int> h_end;
  Buf<unsigned long long> d_edge, d_cs, d_cv, d_dp, d_blob, d_cols;  // d_cv: conv columns (max-plus form), same shape as d_cs
  Buf<StepOut> d_out;
  Buf<unsigned int> d_counter;
  StepOut* h_out = nullptr;  // pinned
  size_t h_out_bytes = 0;
  unsigned long long* h_stage = nullptr;  // pinned staging of the path columns (copied asynchronously)
  size_t stage_cap = 0;
  bool stage_busy = false;
  int n_items = 0, cur_v = -1;
  std::vector<int> pend_desc;
  std::vector<unsigned long long> pend_blob;
  std::string err;
  uint64_t launches = 0;
  int fail(int code, const std::string& m) {
    err = m;
    return code;
  }
};

static thread_local std::string g_dp_create_err;

#define DCK(call)                                                                                      \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess) return dp->fail(CAMUS_B200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

extern "C" {

int camus_b200_dp_create(camus_b200_dp** out, int device, int score_is_f64, int32_t n_nodes, const int32_t* subtree_end,
                         const int32_t* depth, const void* edge_scores) {
  if (!out || !subtree_end || !depth || !edge_scores || n_nodes < 1) {
    g_dp_create_err = "camus_b200_dp_create: null argument or no nodes";
    return CAMUS_B200_ERR_ARG;
  }
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
    g_dp_create_err = "camus_b200_dp_create: no such CUDA device (there is no CPU fallback)";
    return CAMUS_B200_ERR_CUDA;
  }
  auto* dp = new camus_b200_dp();
  dp->device = device;
  dp->f64 = score_is_f64 != 0;
  dp->n = n_nodes;
  auto bail = [&](const char* what, cudaError_t e) {
    g_dp_create_err = std::string("camus_b200_dp_create: ") + what + ": " + cudaGetErrorString(e);
    camus_b200_dp_destroy(dp);
    return CAMUS_B200_ERR_CUDA;
  };
  cudaError_t e;
  if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
  if ((e = cudaStreamCreateWithFlags(&dp->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
  dp->h_out = static_cast<StepOut*>(pinned_cache().take(256, &dp->h_out_bytes));
  if (!dp->h_out) return bail("pinned result", cudaErrorMemoryAllocation);
  dp->stage_cap = (size_t)n_nodes * 128;  // path columns of 128 steps at once; grows on demand
  {
    size_t got = 0;
    dp->h_stage = static_cast<unsigned long long*>(pinned_cache().take(dp->stage_cap * 8, &got));
    if (!dp->h_stage) return bail("pinned staging", cudaErrorMemoryAllocation);
    dp->stage_cap = got / 8;
  }
  const size_t n = (size_t)n_nodes;
  dp->kcap = 128;  // doubled on demand (k_dp_repack)
  dp->dcap = 64;
  if ((e = dp->d_end.alloc(n)) != cudaSuccess || (e = dp->d_depth.alloc(n)) != cudaSuccess || (e = dp->d_dplen.alloc(n)) != cudaSuccess ||
      (e = dp->d_edge.alloc(n * n)) != cudaSuccess || (e = dp->d_cs.alloc(n * dp->kcap)) != cudaSuccess ||
      (e = dp->d_cv.alloc(dp->f64 ? 1 : n * dp->kcap)) != cudaSuccess ||
      (e = dp->d_dp.alloc(n * dp->dcap)) != cudaSuccess ||
      (e = dp->d_items.alloc(2 * n)) != cudaSuccess || (e = dp->d_out.alloc(1)) != cudaSuccess || (e = dp->d_counter.alloc(1)) != cudaSuccess)
    return bail("cudaMalloc", e);
  dp->h_end.assign(subtree_end, subtree_end + n);
  dp->h_dplen.assign(n, 1);
  std::vector<int> ones(n, 1);  // tips: DP[tip] = {0} (main_dp.go:98-100)
  if ((e = cudaMemcpyAsync(dp->d_end.p, subtree_end, n * 4, cudaMemcpyHostToDevice, dp->stream)) != cudaSuccess ||
      (e = cudaMemcpyAsync(dp->d_depth.p, depth, n * 4, cudaMemcpyHostToDevice, dp->stream)) != cudaSuccess ||
      (e = cudaMemcpyAsync(dp->d_dplen.p, ones.data(), n * 4, cudaMemcpyHostToDevice, dp->stream)) != cudaSuccess ||
      (e = cudaMemcpyAsync(dp->d_edge.p, edge_scores, n * n * 8, cudaMemcpyHostToDevice, dp->stream)) != cudaSuccess ||
      (e = cudaMemsetAsync(dp->d_dp.p, 0, n * dp->dcap * 8, dp->stream)) != cudaSuccess ||
      (e = cudaMemsetAsync(dp->d_cs.p, 0, n * dp->kcap * 8, dp->stream)) != cudaSuccess ||
      (e = cudaMemsetAsync(dp->d_cv.p, 0, (dp->f64 ? 1 : n * dp->kcap) * 8, dp->stream)) != cudaSuccess ||
      (e = cudaMemsetAsync(dp->d_counter.p, 0, 4, dp->stream)) != cudaSuccess || (e = cudaStreamSynchronize(dp->stream)) != cudaSuccess)
    return bail("upload", e);
  // Every per-vertex table is sized once for the largest vertex of THIS tree (binary, preorder ids: children of v are
  // v + 1 and subtree_end[v + 1]): cudaMalloc / cudaFree in the middle of the DP cost milliseconds each in a process
  // that already holds tens of GB (0.3 - 0.5 s per DP, measured inside bench.py).
  {
    size_t max_b128 = 1, max_cb64 = 1, max_tiles = 1;
    for (int v = 0; v < n_nodes; ++v) {
      if (subtree_end[v] - v < 3) continue;
      const int c0 = v + 1, c1 = subtree_end[c0];
      if (c1 >= subtree_end[v]) continue;
      const size_t s0 = (size_t)(subtree_end[c0] - c0), s1 = (size_t)(subtree_end[c1] - c1);
      max_b128 = std::max(max_b128, s0 * ((s1 + kAcrossThreads - 1) / kAcrossThreads) + s1 * ((s0 + kAcrossThreads - 1) / kAcrossThreads));
      max_cb64 = std::max(max_cb64, s0 * ((s1 + kTileW - 1) / kTileW) + s1 * ((s0 + kTileW - 1) / kTileW));
      max_tiles = std::max(max_tiles, ((s0 + kTileU - 1) / kTileU + 1) * ((s1 + kTileW - 1) / kTileW) + ((s1 + kTileU - 1) / kTileU + 1) * ((s0 + kTileW - 1) / kTileW));
    }
    if ((e = dp->d_blk_item.alloc(max_b128)) != cudaSuccess || (e = dp->d_item_blk0.alloc(n + 1)) != cudaSuccess ||
        (e = dp->d_item_cb0.alloc(n + 1)) != cudaSuccess || (e = dp->d_tile_tab.alloc(kTabInts * max_tiles)) != cudaSuccess ||
        (e = dp->d_perm.alloc(n)) != cudaSuccess || (e = dp->d_chunk_best.alloc(std::max(max_b128, max_cb64))) != cudaSuccess ||
        (e = dp->d_cols.alloc(n * 128)) != cudaSuccess || (e = dp->d_desc.alloc(3 * n)) != cudaSuccess || (e = dp->d_blob.alloc(16 * n)) != cudaSuccess)
      return bail("cudaMalloc (per-vertex tables)", e);
  }
  dp->tiled_ok = cudaFuncSetAttribute(k_dp_across_tiled<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) == cudaSuccess &&
                 cudaFuncSetAttribute(k_dp_across_maxplus, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) == cudaSuccess &&
                 cudaFuncSetAttribute(k_dp_across_tiled<unsigned long long>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) == cudaSuccess;
  cudaGetLastError();
  *out = dp;
  return CAMUS_B200_OK;
}

void camus_b200_dp_destroy(camus_b200_dp* dp) {
  if (!dp) return;
  cudaSetDevice(dp->device);
  if (dp->stream) {
    cudaStreamSynchronize(dp->stream);  // the buffers go back to the pool in null-stream order, which this stream does not join
    cudaStreamDestroy(dp->stream);
  }
  pinned_cache().give(dp->h_out, dp->h_out_bytes);
  pinned_cache().give(dp->h_stage, dp->stage_cap * 8);
  delete dp;
}

const char* camus_b200_dp_last_error(const camus_b200_dp* dp) { return dp ? dp->err.c_str() : g_dp_create_err.c_str(); }

int camus_b200_dp_set_node(camus_b200_dp* dp, int32_t node, int32_t len, const void* scores) {
  if (!dp) return CAMUS_B200_ERR_ARG;
  if (node < 0 || node >= dp->n || len < 1 || !scores) return dp->fail(CAMUS_B200_ERR_ARG, "dp_set_node: node out of range or empty list");
  dp->max_len = std::max(dp->max_len, (int)len);
  dp->h_dplen[node] = len;
  dp->pend_desc.push_back(node);
  dp->pend_desc.push_back(len);
  dp->pend_desc.push_back((int)dp->pend_blob.size());
  const auto* v = static_cast<const unsigned long long*>(scores);
  dp->pend_blob.insert(dp->pend_blob.end(), v, v + len);
  return CAMUS_B200_OK;
}

static int flush_nodes(camus_b200_dp* dp) {
  if (dp->pend_desc.empty()) return CAMUS_B200_OK;
  int need = dp->dcap;
  for (size_t i = 0; i < dp->pend_desc.size(); i += 3) need = std::max(need, dp->pend_desc[i + 1]);
  if (need > dp->dcap) {
    int cap = dp->dcap;
    while (cap < need) cap *= 2;
    Buf<unsigned long long> bigger;
    DCK(bigger.alloc((size_t)dp->n * cap));
    DCK(cudaMemsetAsync(bigger.p, 0, (size_t)dp->n * cap * 8, dp->stream));
    const size_t tot = (size_t)dp->n * dp->dcap;
    k_dp_repack<<<(unsigned)((tot + 255) / 256), 256, 0, dp->stream>>>(dp->d_dp.p, bigger.p, dp->n, dp->dcap, cap);
    DCK(cudaGetLastError());
    DCK(cudaStreamSynchronize(dp->stream));
    std::swap(dp->d_dp.p, bigger.p);
    std::swap(dp->d_dp.n, bigger.n);
    dp->dcap = cap;
    ++dp->launches;
  }
  const int np = (int)(dp->pend_desc.size() / 3);
  DCK(dp->d_desc.alloc(dp->pend_desc.size()));
  DCK(dp->d_blob.alloc(dp->pend_blob.size()));
  DCK(cudaMemcpyAsync(dp->d_desc.p, dp->pend_desc.data(), dp->pend_desc.size() * 4, cudaMemcpyHostToDevice, dp->stream));
  DCK(cudaMemcpyAsync(dp->d_blob.p, dp->pend_blob.data(), dp->pend_blob.size() * 8, cudaMemcpyHostToDevice, dp->stream));
  k_dp_scatter<<<np, 32, 0, dp->stream>>>(np, dp->d_desc.p, dp->d_blob.p, dp->d_dp.p, dp->d_dplen.p, dp->dcap);
  DCK(cudaGetLastError());
  DCK(cudaStreamSynchronize(dp->stream));  // the host vectors are reused
  ++dp->launches;
  dp->pend_desc.clear();
  dp->pend_blob.clear();
  return CAMUS_B200_OK;
}

int camus_b200_dp_set_vertex(camus_b200_dp* dp, int32_t v, int32_t n_items, const int32_t* item_u, const int32_t* item_sub) {
  if (!dp) return CAMUS_B200_ERR_ARG;
  if (v < 0 || v >= dp->n || n_items < 0 || n_items > dp->n || (n_items && (!item_u || !item_sub)))
    return dp->fail(CAMUS_B200_ERR_ARG, "dp_set_vertex: bad vertex or item list");
  cudaSetDevice(dp->device);
  std::vector<int> it(2 * (size_t)n_items), blk0((size_t)n_items + 1, 0), blk_item;
  for (int i = 0; i < n_items; ++i) {
    if (item_u[i] < 0 || item_u[i] >= dp->n || item_sub[i] < 0 || item_sub[i] >= dp->n) return dp->fail(CAMUS_B200_ERR_ARG, "dp_set_vertex: node out of range");
    it[2 * i] = item_u[i];
    it[2 * i + 1] = item_sub[i];
    const int nw = dp->h_end[item_sub[i]] - item_sub[i];  // nodes of the other subtree
    const int chunks = std::max(1, (nw + kAcrossThreads - 1) / kAcrossThreads);
    blk0[i + 1] = blk0[i] + chunks;
    blk_item.insert(blk_item.end(), chunks, i);
  }
  dp->n_blocks = blk0[n_items];
  // tiled kernel: consecutive items with the same other subtree, kTileU at a time, chunks of kTileW
  std::vector<int> cb0((size_t)n_items + 1, 0), tab;
  for (int i = 0; i < n_items; ++i) {
    const int nw = dp->h_end[item_sub[i]] - item_sub[i];
    cb0[i + 1] = cb0[i] + std::max(1, (nw + kTileW - 1) / kTileW);
  }
  // per distinct other subtree: its nodes sorted by the length of their DP list, longest first (stable)
  std::vector<int> perm;
  std::vector<std::pair<int, int>> perm_of;  // (sub, offset into perm)
  auto perm_offset = [&](int sub) {
    for (auto& po : perm_of)
      if (po.first == sub) return po.second;
    const int off = (int)perm.size(), e = dp->h_end[sub];
    for (int x = sub; x < e; ++x) perm.push_back(x);
    std::stable_sort(perm.begin() + off, perm.end(), [&](int a, int b) { return dp->h_dplen[a] > dp->h_dplen[b]; });
    perm_of.push_back({sub, off});
    return off;
  };
  for (int i = 0; i < n_items;) {  // runs of items with the same other subtree
    int j = i + 1;
    while (j < n_items && item_sub[j] == item_sub[i]) ++j;
    std::vector<int> ord(j - i);
    for (int q = 0; q < j - i; ++q) ord[q] = i + q;
    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return dp->h_dplen[item_u[a]] > dp->h_dplen[item_u[b]]; });
    const int chunks = cb0[i + 1] - cb0[i], poff = perm_offset(item_sub[i]);
    for (size_t t0 = 0; t0 < ord.size(); t0 += kTileU) {
      const int n_it = (int)std::min<size_t>(kTileU, ord.size() - t0);
      for (int c = 0; c < chunks; ++c) {
        tab.push_back(n_it);
        tab.push_back(c);
        tab.push_back(poff);
        tab.push_back(item_sub[i]);
        for (int q = 0; q < kTileU; ++q) tab.push_back(q < n_it ? ord[t0 + q] : ord[t0]);
      }
    }
    i = j;
  }
  dp->n_tile_blocks = (int)(tab.size() / kTabInts);
  if (n_items) {
    DCK(dp->d_tile_tab.alloc(tab.size()));
    DCK(dp->d_item_cb0.alloc(cb0.size()));
    DCK(dp->d_perm.alloc(std::max<size_t>(perm.size(), 1)));
    DCK(cudaMemcpyAsync(dp->d_perm.p, perm.data(), perm.size() * 4, cudaMemcpyHostToDevice, dp->stream));
    DCK(cudaMemcpyAsync(dp->d_tile_tab.p, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, dp->stream));
    DCK(cudaMemcpyAsync(dp->d_item_cb0.p, cb0.data(), cb0.size() * 4, cudaMemcpyHostToDevice, dp->stream));
    DCK(dp->d_chunk_best.alloc((size_t)std::max(cb0[n_items], blk0[n_items])));
  }
  if (n_items) {
    DCK(dp->d_blk_item.alloc(blk_item.size()));
    DCK(dp->d_item_blk0.alloc(blk0.size()));
    DCK(cudaMemcpyAsync(dp->d_items.p, it.data(), it.size() * 4, cudaMemcpyHostToDevice, dp->stream));
    DCK(cudaMemcpyAsync(dp->d_blk_item.p, blk_item.data(), blk_item.size() * 4, cudaMemcpyHostToDevice, dp->stream));
    DCK(cudaMemcpyAsync(dp->d_item_blk0.p, blk0.data(), blk0.size() * 4, cudaMemcpyHostToDevice, dp->stream));
  }
  DCK(cudaStreamSynchronize(dp->stream));
  dp->n_items = n_items;
  dp->cur_v = v;
  return CAMUS_B200_OK;
}

// columns [k0, k1) of nodes [first, first + count) into the cycleDP table (conv = false) or the max-plus table (conv = true)
static int push_columns(camus_b200_dp* dp, bool conv, int32_t k0, int32_t k1, int32_t first, int32_t count, const void* cols, const char* who) {
  if (!dp) return CAMUS_B200_ERR_ARG;
  if (k0 < 0 || k1 < k0 || first < 0 || count < 0 || first + count > dp->n || (!cols && k1 > k0 && count))
    return dp->fail(CAMUS_B200_ERR_ARG, std::string(who) + ": bad range");
  if (conv && dp->f64) return dp->fail(CAMUS_B200_ERR_ARG, std::string(who) + ": the max-plus form needs integer scores (float sums are order dependent)");
  cudaSetDevice(dp->device);
  if (k1 > dp->kcap) {  // both tables share the capacity
    int cap = dp->kcap;
    while (cap < k1) cap *= 2;
    for (int t = 0; t < (dp->f64 ? 1 : 2); ++t) {
      Buf<unsigned long long>& tab = t == 0 ? dp->d_cs : dp->d_cv;
      Buf<unsigned long long> bigger;
      DCK(bigger.alloc((size_t)dp->n * cap));
      DCK(cudaMemsetAsync(bigger.p, 0, (size_t)dp->n * cap * 8, dp->stream));
      const size_t tot = (size_t)dp->n * dp->kcap;
      k_dp_repack<<<(unsigned)((tot + 255) / 256), 256, 0, dp->stream>>>(tab.p, bigger.p, dp->n, dp->kcap, cap);
      DCK(cudaGetLastError());
      DCK(cudaStreamSynchronize(dp->stream));
      std::swap(tab.p, bigger.p);
      std::swap(tab.n, bigger.n);
      ++dp->launches;
    }
    dp->kcap = cap;
  }
  const size_t tot = (size_t)(k1 - k0) * count;
  if (tot) {
    // through a pinned staging buffer, asynchronously: the across call synchronises once per step
    if (dp->stage_busy) {
      DCK(cudaStreamSynchronize(dp->stream));
      dp->stage_busy = false;
    }
    if (tot > dp->stage_cap) {
      pinned_cache().give(dp->h_stage, dp->stage_cap * 8);
      dp->h_stage = nullptr;
      dp->stage_cap = 0;
      size_t got = 0;
      dp->h_stage = static_cast<unsigned long long*>(pinned_cache().take(tot * 2 * 8, &got));
      if (!dp->h_stage) return dp->fail(CAMUS_B200_ERR_CUDA, "pinned staging buffer");
      dp->stage_cap = got / 8;
    }
    memcpy(dp->h_stage, cols, tot * 8);
    DCK(dp->d_cols.alloc(tot));
    DCK(cudaMemcpyAsync(dp->d_cols.p, dp->h_stage, tot * 8, cudaMemcpyHostToDevice, dp->stream));
    k_cs_scatter<<<(unsigned)((tot + 255) / 256), 256, 0, dp->stream>>>(dp->d_cols.p, conv ? dp->d_cv.p : dp->d_cs.p, dp->kcap, k0, k1 - k0, first, count);
    DCK(cudaGetLastError());
    ++dp->launches;
    dp->stage_busy = true;
  }
  return CAMUS_B200_OK;
}

int camus_b200_dp_set_path_columns(camus_b200_dp* dp, int32_t k0, int32_t k1, int32_t first, int32_t count, const void* cols) {
  return push_columns(dp, false, k0, k1, first, count, cols, "dp_set_path_columns");
}
int camus_b200_dp_set_conv_columns(camus_b200_dp* dp, int32_t k0, int32_t k1, int32_t first, int32_t count, const void* cols) {
  return push_columns(dp, true, k0, k1, first, count, cols, "dp_set_conv_columns");
}

// fold + read-back shared by the two across forms
static int finish_step(camus_b200_dp* dp, const int* cb0, camus_b200_dp_best* out) {
  if (dp->f64)
    k_dp_fold<double><<<1, kFoldThreads, 0, dp->stream>>>(dp->n_items, dp->d_items.p, cb0, dp->d_depth.p, dp->cur_v, dp->d_chunk_best.p, dp->d_out.p);
  else
    k_dp_fold<unsigned long long><<<1, kFoldThreads, 0, dp->stream>>>(dp->n_items, dp->d_items.p, cb0, dp->d_depth.p, dp->cur_v, dp->d_chunk_best.p, dp->d_out.p);
  DCK(cudaGetLastError());
  dp->launches += 2;
  DCK(cudaMemcpyAsync(dp->h_out, dp->d_out.p, sizeof(StepOut), cudaMemcpyDeviceToHost, dp->stream));
  DCK(cudaStreamSynchronize(dp->stream));
  dp->stage_busy = false;
  const StepOut& r = *dp->h_out;
  out->found = r.best.ok;
  out->nan_seen = r.nan_seen;
  out->first_is_nan = r.first.ok;
  out->first_u = r.first.u;
  out->first_w = r.first.w;
  out->first_score_bits = r.first.score;
  for (int q = 0; q < 4; ++q) out->first_split[q] = r.first.idx[q];
  out->score_bits = r.best.score;
  out->u = r.best.u;
  out->w = r.best.w;
  out->cycle_length = r.best.len;
  for (int q = 0; q < 4; ++q) out->split[q] = r.best.idx[q];
  return CAMUS_B200_OK;
}

int camus_b200_dp_across_maxplus(camus_b200_dp* dp, int32_t prev_k, camus_b200_dp_best* out) {
  if (!dp || !out) return CAMUS_B200_ERR_ARG;
  if (dp->f64) return dp->fail(CAMUS_B200_ERR_ARG, "dp_across_maxplus: integer scores only");
  if (dp->cur_v < 0) return dp->fail(CAMUS_B200_ERR_ARG, "dp_across_maxplus: set the vertex first");
  if (prev_k < 0 || prev_k >= dp->kcap) return dp->fail(CAMUS_B200_ERR_ARG, "dp_across_maxplus: conv columns 0..prev_k have not been set");
  cudaSetDevice(dp->device);
  memset(out, 0, sizeof *out);
  if (dp->n_items == 0) return CAMUS_B200_OK;
  const size_t shm = ((size_t)(prev_k + 1) * (kTileLd + kTileU)) * 8 + (size_t)kTileU * kTileW * (8 + 4 + 4);
  if (!dp->tiled_ok || shm > 200 * 1024) return dp->fail(CAMUS_B200_ERR_NOMEM, "dp_across_maxplus: step too large for the shared-memory tile; use camus_b200_dp_across");
  DCK(cudaMemsetAsync(dp->d_out.p, 0, sizeof(StepOut), dp->stream));
  k_dp_across_maxplus<<<dp->n_tile_blocks, dim3(kTileW, kTileU), shm, dp->stream>>>(dp->d_items.p, dp->d_tile_tab.p, dp->d_perm.p, dp->d_item_cb0.p,
                                                                                   dp->d_end.p, prev_k, dp->n, dp->kcap, dp->d_cv.p, dp->d_edge.p,
                                                                                   dp->d_chunk_best.p);
  DCK(cudaGetLastError());
  return finish_step(dp, dp->d_item_cb0.p, out);
}

int camus_b200_dp_across(camus_b200_dp* dp, int32_t prev_k, camus_b200_dp_best* out) {
  if (!dp || !out) return CAMUS_B200_ERR_ARG;
  if (dp->cur_v < 0) return dp->fail(CAMUS_B200_ERR_ARG, "dp_across: set the vertex first");
  if (prev_k < 0 || prev_k >= dp->kcap) return dp->fail(CAMUS_B200_ERR_ARG, "dp_across: path columns 0..prev_k have not been set");
  cudaSetDevice(dp->device);
  memset(out, 0, sizeof *out);
  if (dp->n_items == 0) return CAMUS_B200_OK;
  if (int rc = flush_nodes(dp)) return rc;
  DCK(cudaMemsetAsync(dp->d_out.p, 0, sizeof(StepOut), dp->stream));
  static const bool no_tiled = getenv("CAMUS_DP_NO_TILED") && getenv("CAMUS_DP_NO_TILED")[0] == '1';  // A/B, tests
  const int ml = std::min(dp->max_len, dp->dcap);
  const size_t shm_tiled = ((size_t)(prev_k + 1 + ml) * (kTileLd + kTileU)) * 8 + (size_t)kTileU * kTileW * (8 + 4 + 4 + 16);
  const bool tiled = dp->tiled_ok && !no_tiled && shm_tiled <= 200 * 1024;
  const int* cb0 = tiled ? dp->d_item_cb0.p : dp->d_item_blk0.p;
  if (tiled) {
    const dim3 blk(kTileW, kTileU);
    if (dp->f64)
      k_dp_across_tiled<double><<<dp->n_tile_blocks, blk, shm_tiled, dp->stream>>>(dp->d_items.p, dp->d_tile_tab.p, dp->d_perm.p, dp->d_item_cb0.p, dp->d_end.p, prev_k, dp->n,
                                                                                   dp->kcap, dp->dcap, ml, dp->d_cs.p, dp->d_dp.p, dp->d_dplen.p, dp->d_edge.p,
                                                                                   dp->d_chunk_best.p, dp->d_out.p);
    else
      k_dp_across_tiled<unsigned long long><<<dp->n_tile_blocks, blk, shm_tiled, dp->stream>>>(dp->d_items.p, dp->d_tile_tab.p, dp->d_perm.p, dp->d_item_cb0.p, dp->d_end.p,
                                                                                               prev_k, dp->n, dp->kcap, dp->dcap, ml, dp->d_cs.p, dp->d_dp.p,
                                                                                               dp->d_dplen.p, dp->d_edge.p, dp->d_chunk_best.p, dp->d_out.p);
  } else {
    const size_t shm = (size_t)(prev_k + 1 + dp->dcap) * 8;
    if (dp->f64)
      k_dp_across<double><<<dp->n_blocks, kAcrossThreads, shm, dp->stream>>>(dp->n_items, dp->n_blocks, dp->d_items.p, dp->d_blk_item.p, dp->d_item_blk0.p,
                                                                             dp->d_end.p, dp->d_depth.p, dp->cur_v, prev_k, dp->n, dp->kcap, dp->dcap,
                                                                             dp->d_cs.p, dp->d_dp.p, dp->d_dplen.p, dp->d_edge.p, dp->d_chunk_best.p,
                                                                             dp->d_out.p);
    else
      k_dp_across<unsigned long long><<<dp->n_blocks, kAcrossThreads, shm, dp->stream>>>(dp->n_items, dp->n_blocks, dp->d_items.p, dp->d_blk_item.p,
                                                                                         dp->d_item_blk0.p, dp->d_end.p, dp->d_depth.p, dp->cur_v, prev_k,
                                                                                         dp->n, dp->kcap, dp->dcap, dp->d_cs.p, dp->d_dp.p, dp->d_dplen.p,
                                                                                         dp->d_edge.p, dp->d_chunk_best.p, dp->d_out.p);
  }
  DCK(cudaGetLastError());
  return finish_step(dp, cb0, out);
}

uint64_t camus_b200_dp_launches(const camus_b200_dp* dp) { return dp ? dp->launches : 0; }

}  // extern "C"
