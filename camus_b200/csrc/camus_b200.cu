// libcamus_b200.so: context management + the C-ABI of include/camus_b200.h over the sm_100a
// kernels in kernels_*.cuh. Host code here only validates and re-indexes the flat trees, plans
// memory, launches kernels and copies results; every quartet is counted, filtered and scored on
// the GPU. There is no CPU fallback.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "camus_b200.h"
#include "kernels_count.cuh"
#include "kernels_ingest.cuh"
#include "kernels_prep.cuh"
#include "kernels_retic.cuh"
#include "kernels_score.cuh"

using namespace camus_dev;

namespace {

thread_local std::string g_create_err;

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  cudaError_t alloc(size_t count) {
    if (count <= n && p) return cudaSuccess;
    release();
    if (count == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc(&p, count * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
  size_t bytes() const { return n * sizeof(T); }
};

}  // namespace

constexpr int kDirectImageMaxBytes = 200 * 1024;  // shared [taxon][32 trees] image of k_depth_slices_direct

struct camus_b200 {
  // n_gpus > 1: this handle only fans out to one single-device context per GPU (subs); the 4-set
  // index space is dealt to them exactly as to the ranks of a multi-process run (BoxMap), and
  // their partial junction tables are summed on subs[0] over NVLink peer reads.
  std::vector<camus_b200*> subs;
  int user_rank = 0, user_world = 1;  // partition of the whole handle (multi-process on top)
  bool peer_access = false;
  bool reduced = false;                // subs[0]'s junction table already holds the sum over all devices
  DevBuf<unsigned long long> d_stage;  // on subs[0]: landing buffer when peer access is unavailable

  int device = 0;
  std::string err;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  int rank = 0, world = 1;

  // ---- constraint tree ----
  int n = 0, N = 0, npad = 0, M = 0;
  std::vector<int32_t> int_of_ref, rank_of_tip_h;
  DevBuf<int32_t> d_parent, d_depth, d_sz, d_rc, d_nleaves, d_firstleaf, d_leafnode, d_tip_of_rank, d_ref_id, d_rank_of_tip;
  DevBuf<uint32_t> d_lcaD;
  DevBuf<uint8_t> d_seg_uniform;
  ConstraintDev ct{};

  // ---- gene trees (host staging until the next count) ----
  std::vector<int64_t> h_node_off{0};
  std::vector<int32_t> h_parent, h_taxon;
  uint64_t resolutions = 0;
  int G = 0, G32 = 0;
  bool genes_dirty = true;  // triple planes need (re)building

  // ---- device state ----
  DevBuf<int64_t> d_node_off;
  DevBuf<int32_t> d_gparent, d_gtaxon, d_nleaf;
  DevBuf<uint16_t> d_depth_tmp, d_lrT, d_hT, d_Dt;
  DevBuf<uint32_t> d_Ds;  // bit-sliced depth rows of one batch (k_depth_slices)
  DevBuf<uint16_t> d_lrP, d_hP, d_posP;  // per-tree leaf order, h and taxon -> position (k_depth_slices_direct)
  DevBuf<PlaneDesc> d_plane_list;  // work list of the sliced plane builder (make_plane_list)
  uint32_t n_plane_list = 0;
  int plane_list_key[4] = {-1, -1, -1, -1};  // segments, rank, world, partial
  DevBuf<uint32_t> d_tri;
  DevBuf<uint4> d_cells;
  DevBuf<unsigned long long> d_Z, d_E, d_out, d_totals, d_keys, d_nout;
  DevBuf<uint32_t> d_counts, d_thr_tab;
  uint64_t n_local = 0, chunk = 1;  // this rank's share of the boxes (BoxMap)
  // class-scheme partition (layout.cuh): explicit local box list + the tiles this rank builds
  bool class_scheme = false;
  DevBuf<uint32_t> d_box_list, d_tile_list, d_tile_slot;  // tile_slot: tile index -> compact slot (0xFFFFFFFF = not stored)
  bool tri_compact = false;        // the bit-planes hold only this share's tiles, in tile_list order ...
  int tri_key[3] = {-1, -1, -1};   // ... of this (segments, rank, world)
  uint64_t n_tiles_local = 0;
  int list_key[3] = {-1, -1, -1};  // (segments, rank, world) the lists were built for
  DevBuf<uint2> d_box_desc;        // k_box_desc: segments + regular flag of every local box (fused count kernel)
  uint64_t desc_key[6] = {0, 0, 0, 0, 0, 0};  // (constraint generation, segments, rank, world, scheme, regular) it was built for
  uint64_t constraint_gen = 0;     // bumped by set_constraint
  uint32_t tri_mask = 0;           // segment classes whose tiles the bit-planes hold (0xF = every tile)
  bool want_all_tiles = false;     // camus_b200_query_cells reads arbitrary tiles
  enum CellState { NONE, RAW, FILTERED } cell_state = NONE;
  int last_qmode = 0;
  double last_threshold = 0;
  bool z_valid = false;
  int z_as_set = -1;
  bool counted = false;      // count_filter has run for the current trees / partition
  bool use_fused = true;     // count + filter + score in one kernel (no cell table); CAMUS_B200_UNFUSED=1 disables
  int as_set_hint = 0;       // weight the fused pass of count_filter scores with
  int max_depth = 0;         // deepest node over all gene trees
  bool all_complete = true;  // every gene tree so far holds all taxa and is fully resolved
  bool all_rooted = true;    // ... and none of them has a trifurcating (unrooted) root
  bool allow_half = true;      // CAMUS_B200_NO_HALF=1 forces the integer plane builder (A/B, tests)
  bool allow_complete = true;  // CAMUS_B200_NO_COMPLETE=1 forces the general kernel (A/B, tests)
  bool allow_sliced = true;    // CAMUS_B200_NO_SLICED=1 forces the packed-fp16 plane builder (A/B, tests)
  bool allow_direct = true;    // CAMUS_B200_NO_DIRECT=1: slices from the depth table instead of straight from the trees (A/B, tests)
  bool allow_regular = true;   // CAMUS_B200_NO_REGULAR=1 sends every box through the survivor queue (A/B, tests)

  // ---- library-owned collective (camus_b200_comm_init) ----
  ncclComm_t comm = nullptr;
  bool z_reduced = false;             // the junction table already holds the sum over all ranks
  DevBuf<unsigned long long> d_scal;  // the two survivor scalars, all-reduced in place
  std::vector<int32_t> h_depth, h_nleaves, h_ref_of_int;  // constraint tables in internal ids (constraint_tables)

  // ---- gene-tree ingest on the GPU (camus_b200_add_gene_trees_newick) ----
  DevBuf<unsigned long long> d_lab_hash;
  DevBuf<int> d_lab_tip;
  DevBuf<long long> d_lab_off;
  DevBuf<char> d_lab_pool;
  unsigned int lab_mask = 0;
  uint64_t lab_gen = 0;  // constraint generation the label table belongs to (0 = none)
  DevBuf<char> nwk_text;  // scratch of add_gene_trees_newick, kept between calls
  DevBuf<long long> nwk_lb, nwk_le, nwk_off;
  DevBuf<int> nwk_stack, nwk_par, nwk_tax, nwk_nchild;
  DevBuf<unsigned char> nwk_coll;
  DevBuf<unsigned int> nwk_seen;
  DevBuf<NwkTreeInfo> nwk_info;
  // ---- inputs beyond the bit-plane budget: the 8 class-scheme shares run one after the other on this GPU ----
  int z_rep_used = 1;        // private copies of Z the running pass sequence uses
  int passes = 1;            // 8 while such a run is in progress or was the last one
  int passes_env = 0;        // CAMUS_B200_PASSES at create: 8 / 1 force the mode on / off (tests, A/B)
  bool force_class = false;  // passes mode: class scheme whatever the number of segments (<= 256)
  bool allow_packed = true;  // CAMUS_B200_NO_PACKED=1 keeps the 3-register general kernel below 1024 trees (A/B, tests)

  float ms[CAMUS_B200_T_N] = {0};
  uint64_t launches = 0;

  int fail(int code, const std::string& m) {
    err = m;
    return code;
  }
};

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return ctx->fail(CAMUS_B200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
  } while (0)

#define CK_LAUNCH()              \
  do {                           \
    ++ctx->launches;             \
    CK(cudaGetLastError());      \
  } while (0)

namespace {

struct StageTimer {
  camus_b200* ctx;
  int stage;
  StageTimer(camus_b200* c, int s) : ctx(c), stage(s) { cudaEventRecord(ctx->ev[0], ctx->stream); }
  int stop() {
    cudaEventRecord(ctx->ev[1], ctx->stream);
    cudaError_t e = cudaEventSynchronize(ctx->ev[1]);
    if (e != cudaSuccess) return ctx->fail(CAMUS_B200_ERR_CUDA, std::string("stage sync: ") + cudaGetErrorString(e));
    float t = 0;
    cudaEventElapsedTime(&t, ctx->ev[0], ctx->ev[1]);
    ctx->ms[stage] += t;
    return 0;
  }
};

template <class T>
int upload(camus_b200* ctx, DevBuf<T>& buf, const T* src, size_t count) {
  CK(buf.alloc(std::max<size_t>(count, 1)));
  if (count) CK(cudaMemcpyAsync(buf.p, src, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}

// n_gpus > 1: run f(sub, index) on every device at once (the single-device calls block on their
// own stream), first failure wins.
template <class F>
int fan_out(camus_b200* ctx, F f) {
  const int n = (int)ctx->subs.size();
  std::vector<int> rc(n, 0);
  std::vector<std::thread> th;
  for (int i = 1; i < n; ++i) th.emplace_back([&, i] { rc[i] = f(ctx->subs[i], i); });
  rc[0] = f(ctx->subs[0], 0);
  for (auto& t : th) t.join();
  for (int i = 0; i < n; ++i)
    if (rc[i]) return ctx->fail(rc[i], "device " + std::to_string(ctx->subs[i]->device) + ": " + ctx->subs[i]->err);
  return 0;
}
int reduce_partials(camus_b200* ctx);
void nccl_api_destroy(ncclComm_t comm);

int build_planes(camus_b200* ctx, bool do_upload);
bool planes_cover(camus_b200* ctx);
int wanted_passes(camus_b200* ctx);
int run_passes(camus_b200* ctx, int passes, int qmode, double threshold, int as_set, bool do_upload, uint64_t* ns, uint64_t* sum);
int set_box_range(camus_b200* ctx);
int run_count(camus_b200* ctx);
int run_filter(camus_b200* ctx, int qmode, double threshold, uint64_t* ns, uint64_t* sum);

uint64_t cells_budget_boxes(size_t free_bytes) {
  size_t reserve = (size_t)3 << 30;
  size_t budget = free_bytes > reserve ? free_bytes - reserve : free_bytes / 2;
  budget = std::min<size_t>(budget, (size_t)48 << 30);
  return std::max<uint64_t>(1, budget / (kBoxCells * sizeof(uint4)));
}

}  // namespace

extern "C" {

static int create_one(camus_b200** out, int dev) {
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    g_create_err = std::string("camus_b200_create: no usable CUDA device (") +
                   (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") + "); there is no CPU fallback";
    return CAMUS_B200_ERR_CUDA;
  }
  if (dev < 0 || dev >= count) {
    g_create_err = "camus_b200_create: device ordinal out of range";
    return CAMUS_B200_ERR_ARG;
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, dev);
  if (prop.major != 10) {
    g_create_err = "camus_b200_create: device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                   ", this library is built for sm_100a only";
    return CAMUS_B200_ERR_CUDA;
  }
  camus_b200* ctx = new camus_b200();
  ctx->device = dev;
  if (cudaSetDevice(dev) != cudaSuccess || cudaStreamCreate(&ctx->stream) != cudaSuccess ||
      cudaEventCreate(&ctx->ev[0]) != cudaSuccess || cudaEventCreate(&ctx->ev[1]) != cudaSuccess) {
    g_create_err = std::string("camus_b200_create: ") + cudaGetErrorString(cudaGetLastError());
    delete ctx;
    return CAMUS_B200_ERR_CUDA;
  }
  {
    cudaError_t a = cudaSuccess;
    auto set = [&](auto kernel, int bytes) {
      cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
      if (a == cudaSuccess) a = e;
    };
    set(k_count<false, kModeGeneral>, CountCfg<kModeGeneral>::kSmemBytes);
    set(k_count<true, kModeGeneral>, CountCfg<kModeGeneral>::kSmemBytes);
    set(k_count<false, kModeComplete>, CountCfg<kModeComplete>::kSmemBytes);
    set(k_count<true, kModeComplete>, CountCfg<kModeComplete>::kSmemBytes + 64 * 1024);  // + room for CAMUS_B200_COUNT_PAD_SMEM
    set(k_count<true, kModePacked>, CountCfg<kModePacked>::kSmemBytes);
    // the sliced plane builders: set once here -- cudaFuncSetAttribute between launches drains the stream
    // (measured: 16 calls per step cost 10 ms of idle GPU at 1000 taxa, 16 ms of a 104 ms step at 500)
#define CAMUS_SET_SLICED(nb)                                                                    \
  set(k_triple_planes_bs<nb, true>, 2 * 4 * 3 * SliceArray<SliceRow<nb>::W>::words * 4);        \
  set(k_triple_planes_bs<nb, false>, 2 * 4 * 3 * SliceArray<SliceRow<nb>::W>::words * 4);       \
  set(k_depth_slices_direct<nb>, kDirectImageMaxBytes);
    CAMUS_SET_SLICED(5)
    CAMUS_SET_SLICED(6)
    CAMUS_SET_SLICED(7)
    CAMUS_SET_SLICED(9)
    CAMUS_SET_SLICED(11)
#undef CAMUS_SET_SLICED
    if (a != cudaSuccess) {
      g_create_err = std::string("camus_b200_create: count / plane kernels do not fit this device (") + cudaGetErrorString(a) + ")";
      camus_b200_destroy(ctx);
      return CAMUS_B200_ERR_CUDA;
    }
  }
  if (const char* e = getenv("CAMUS_B200_NO_COMPLETE")) ctx->allow_complete = !(e[0] == '1');
  if (const char* e = getenv("CAMUS_B200_NO_REGULAR")) ctx->allow_regular = !(e[0] == '1');
  if (const char* e = getenv("CAMUS_B200_NO_HALF")) ctx->allow_half = !(e[0] == '1');
  if (const char* e = getenv("CAMUS_B200_NO_SLICED")) ctx->allow_sliced = !(e[0] == '1');
  if (const char* e = getenv("CAMUS_B200_NO_DIRECT")) ctx->allow_direct = !(e[0] == '1');
  if (const char* e = getenv("CAMUS_B200_UNFUSED")) ctx->use_fused = !(e[0] == '1');
  if (const char* e = getenv("CAMUS_B200_NO_PACKED")) ctx->allow_packed = !(e[0] == '1');
  if (const char* e = getenv("CAMUS_B200_PASSES")) ctx->passes_env = atoi(e);
  *out = ctx;
  return CAMUS_B200_OK;
}

int camus_b200_create(camus_b200** out, int n_gpus, const int* devices) {
  if (!out) return CAMUS_B200_ERR_ARG;
  *out = nullptr;
  if (n_gpus < 1 || n_gpus > 16) {
    g_create_err = "camus_b200_create: n_gpus must be in 1..16 (GPUs of one process)";
    return CAMUS_B200_ERR_ARG;
  }
  if (n_gpus == 1) return create_one(out, devices ? devices[0] : 0);
  for (int i = 0; i < n_gpus; ++i)
    for (int j = i + 1; devices && j < n_gpus; ++j)
      if (devices[i] == devices[j]) {
        g_create_err = "camus_b200_create: device listed twice";
        return CAMUS_B200_ERR_ARG;
      }
  camus_b200* ctx = new camus_b200();
  for (int i = 0; i < n_gpus; ++i) {
    camus_b200* sub = nullptr;
    int rc = create_one(&sub, devices ? devices[i] : i);
    if (rc) {
      camus_b200_destroy(ctx);
      return rc;
    }
    ctx->subs.push_back(sub);
    camus_b200_set_partition(sub, i, n_gpus);
  }
  // subs[0] sums the partial tables by reading its peers' memory directly
  ctx->peer_access = true;
  cudaSetDevice(ctx->subs[0]->device);
  for (int i = 1; i < n_gpus; ++i) {
    int can = 0;
    cudaDeviceCanAccessPeer(&can, ctx->subs[0]->device, ctx->subs[i]->device);
    if (can) {
      cudaError_t e = cudaDeviceEnablePeerAccess(ctx->subs[i]->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can = 0;
      cudaGetLastError();
    }
    if (!can) ctx->peer_access = false;
  }
  *out = ctx;
  return CAMUS_B200_OK;
}

void camus_b200_destroy(camus_b200* ctx) {
  if (!ctx) return;
  if (!ctx->subs.empty()) {
    cudaSetDevice(ctx->subs[0]->device);
    ctx->d_stage.release();
    for (camus_b200* s : ctx->subs) camus_b200_destroy(s);
    delete ctx;
    return;
  }
  cudaSetDevice(ctx->device);
  if (ctx->comm) nccl_api_destroy(ctx->comm);
  if (ctx->ev[0]) cudaEventDestroy(ctx->ev[0]);
  if (ctx->ev[1]) cudaEventDestroy(ctx->ev[1]);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* camus_b200_last_error(const camus_b200* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int camus_b200_set_partition(camus_b200* ctx, int rank, int world) {
  if (!ctx || world < 1 || rank < 0 || rank >= world) return ctx ? ctx->fail(CAMUS_B200_ERR_ARG, "bad partition") : CAMUS_B200_ERR_ARG;
  if (!ctx->subs.empty()) {  // this process's share is dealt on to its devices
    const int n = (int)ctx->subs.size();
    ctx->user_rank = rank;
    ctx->user_world = world;
    ctx->reduced = false;
    for (int i = 0; i < n; ++i) camus_b200_set_partition(ctx->subs[i], rank * n + i, world * n);
    return 0;
  }
  ctx->rank = rank;
  ctx->world = world;
  ctx->cell_state = camus_b200::NONE;
  ctx->z_valid = false;
  ctx->z_reduced = false;
  ctx->counted = false;
  return 0;
}

int camus_b200_set_constraint(camus_b200* ctx, int32_t n_nodes, int32_t n_taxa, const int32_t* parent,
                              const int32_t* child0, const int32_t* child1, const int32_t* tip_index) {
  if (!ctx) return CAMUS_B200_ERR_ARG;
  if (!parent || !child0 || !child1 || !tip_index || n_taxa < 2 || n_nodes != 2 * n_taxa - 1)
    return ctx->fail(CAMUS_B200_ERR_ARG, "set_constraint: need a rooted binary tree (n_nodes = 2*n_taxa - 1)");
  if (n_taxa > 32767) return ctx->fail(CAMUS_B200_ERR_ARG, "set_constraint: at most 32767 taxa (15-bit packing, quartet.go:17)");
  if (!ctx->subs.empty()) {
    ctx->reduced = false;
    int rc = fan_out(ctx, [&](camus_b200* s, int) { return camus_b200_set_constraint(s, n_nodes, n_taxa, parent, child0, child1, tip_index); });
    ctx->n = rc ? 0 : n_nodes;
    ctx->N = rc ? 0 : n_taxa;
    return rc;
  }
  CK(cudaSetDevice(ctx->device));
  const int n = n_nodes, N = n_taxa;
  int root = -1;
  for (int i = 0; i < n; ++i) {
    if (parent[i] < 0) {
      if (root >= 0) return ctx->fail(CAMUS_B200_ERR_INPUT, "set_constraint: more than one root");
      root = i;
    }
    if (parent[i] >= n) return ctx->fail(CAMUS_B200_ERR_INPUT, "set_constraint: parent index out of range");
    bool tip = child0[i] < 0;
    if (tip != (child1[i] < 0) || (tip && (tip_index[i] < 0 || tip_index[i] >= N)) || (!tip && (child0[i] >= n || child1[i] >= n)))
      return ctx->fail(CAMUS_B200_ERR_INPUT, "set_constraint: node is neither a tip nor binary");
  }
  if (root < 0) return ctx->fail(CAMUS_B200_ERR_INPUT, "set_constraint: no root");
  // internal preorder ids (left child = id + 1), DFS ranks of the leaves
  std::vector<int32_t> ref_of_int, par(n), depth(n), sz(n, 1), rc(n, -1), nleaves(n, 0), firstleaf(n, 0);
  std::vector<int32_t> leafnode(N), tip_of_rank(N);
  // built in locals and moved into the context only when the whole tree has been accepted
  std::vector<int32_t> int_of_ref(n, -1), rank_of_tip_h(N, -1);
  ref_of_int.reserve(n);
  {
    std::vector<int> st{root};
    while (!st.empty()) {
      int x = st.back();
      st.pop_back();
      if (int_of_ref[x] >= 0 || (int)ref_of_int.size() >= n)
        return ctx->fail(CAMUS_B200_ERR_INPUT, "set_constraint: child pointers do not form a tree");
      int_of_ref[x] = (int)ref_of_int.size();
      ref_of_int.push_back(x);
      if (child0[x] >= 0) {
        if (parent[child0[x]] != x || parent[child1[x]] != x)
          return ctx->fail(CAMUS_B200_ERR_INPUT, "set_constraint: parent/child arrays disagree");
        st.push_back(child1[x]);
        st.push_back(child0[x]);
      }
    }
    if ((int)ref_of_int.size() != n) return ctx->fail(CAMUS_B200_ERR_INPUT, "set_constraint: unreachable nodes");
  }
  int nl = 0;
  for (int i = 0; i < n; ++i) {
    int r = ref_of_int[i];
    par[i] = parent[r] < 0 ? -1 : int_of_ref[parent[r]];
    if (par[i] >= i) return ctx->fail(CAMUS_B200_ERR_INPUT, "set_constraint: parent/child arrays disagree");
    depth[i] = par[i] < 0 ? 0 : depth[par[i]] + 1;
    if (child0[r] >= 0) {
      rc[i] = int_of_ref[child1[r]];
    } else {
      if (nl >= N || rank_of_tip_h[tip_index[r]] >= 0) return ctx->fail(CAMUS_B200_ERR_INPUT, "set_constraint: tip indices are not a permutation");
      leafnode[nl] = i;
      tip_of_rank[nl] = tip_index[r];
      rank_of_tip_h[tip_index[r]] = nl;
      nleaves[i] = 1;
      firstleaf[i] = nl;
      ++nl;
    }
  }
  if (nl != N) return ctx->fail(CAMUS_B200_ERR_INPUT, "set_constraint: leaf count mismatch");
  for (int i = n - 1; i > 0; --i) {
    sz[par[i]] += sz[i];
    nleaves[par[i]] += nleaves[i];
  }
  for (int i = n - 1; i >= 0; --i)
    if (rc[i] >= 0) firstleaf[i] = firstleaf[i + 1];
  ctx->int_of_ref = std::move(int_of_ref);
  ctx->rank_of_tip_h = std::move(rank_of_tip_h);
  ctx->h_depth = depth;
  ctx->h_nleaves = nleaves;
  ctx->h_ref_of_int = ref_of_int;
  ctx->n = n;
  ctx->N = N;
  ++ctx->constraint_gen;
  ctx->npad = (N + kSeg - 1) / kSeg * kSeg;
  ctx->M = ctx->npad / kSeg;
  if (upload(ctx, ctx->d_parent, par.data(), n) || upload(ctx, ctx->d_depth, depth.data(), n) ||
      upload(ctx, ctx->d_sz, sz.data(), n) || upload(ctx, ctx->d_rc, rc.data(), n) ||
      upload(ctx, ctx->d_nleaves, nleaves.data(), n) || upload(ctx, ctx->d_firstleaf, firstleaf.data(), n) ||
      upload(ctx, ctx->d_leafnode, leafnode.data(), N) || upload(ctx, ctx->d_tip_of_rank, tip_of_rank.data(), N) ||
      upload(ctx, ctx->d_ref_id, ref_of_int.data(), n) || upload(ctx, ctx->d_rank_of_tip, ctx->rank_of_tip_h.data(), N))
    return CAMUS_B200_ERR_CUDA;
  CK(ctx->d_lcaD.alloc((size_t)N * N));
  CK(ctx->d_seg_uniform.alloc((size_t)ctx->M * ctx->M));
  ctx->ct = ConstraintDev{n, N, ctx->d_parent.p, ctx->d_depth.p, ctx->d_sz.p, ctx->d_rc.p, ctx->d_nleaves.p,
                          ctx->d_firstleaf.p, ctx->d_leafnode.p, ctx->d_tip_of_rank.p, ctx->d_ref_id.p, ctx->d_lcaD.p,
                          ctx->d_seg_uniform.p, ctx->M};
  dim3 grid((N + 127) / 128, N);
  k_constraint_lca<<<grid, 128, 0, ctx->stream>>>(N, ctx->d_leafnode.p, ctx->d_parent.p, ctx->d_sz.p, ctx->d_depth.p, ctx->d_lcaD.p);
  CK_LAUNCH();
  k_segment_pairs<<<(ctx->M * ctx->M + 7) / 8, 256, 0, ctx->stream>>>(N, ctx->M, ctx->d_lcaD.p, ctx->d_seg_uniform.p);
  CK_LAUNCH();
  CK(cudaStreamSynchronize(ctx->stream));
  // taxon numbering changed: gene trees given for an earlier constraint are dropped
  ctx->h_node_off.assign(1, 0);
  ctx->h_parent.clear();
  ctx->h_taxon.clear();
  ctx->resolutions = 0;
  ctx->all_complete = true;
  ctx->all_rooted = true;
  ctx->max_depth = 0;
  ctx->G = 0;
  ctx->genes_dirty = true;
  ctx->cell_state = camus_b200::NONE;
  ctx->z_valid = false;
  ctx->counted = false;
  return 0;
}

int camus_b200_clear_gene_trees(camus_b200* ctx) {
  if (!ctx) return CAMUS_B200_ERR_ARG;
  if (!ctx->subs.empty()) {
    ctx->reduced = false;
    for (camus_b200* s : ctx->subs) camus_b200_clear_gene_trees(s);
    return 0;
  }
  ctx->h_node_off.assign(1, 0);
  ctx->h_parent.clear();
  ctx->h_taxon.clear();
  ctx->resolutions = 0;
  ctx->all_complete = true;
  ctx->all_rooted = true;
  ctx->max_depth = 0;
  ctx->G = 0;
  ctx->genes_dirty = true;
  ctx->cell_state = camus_b200::NONE;
  ctx->z_valid = false;
  ctx->counted = false;
  return 0;
}

namespace {
// what validating a batch of flat gene trees establishes (a handle over several GPUs validates once)
struct BatchFacts {
  uint64_t res = 0;
  int max_depth = 0;
  bool all_complete = true, all_rooted = true;
};
int add_gene_trees_impl(camus_b200* ctx, int32_t n_trees, const int64_t* node_off, const int32_t* parent, const int32_t* taxon,
                        const BatchFacts* known, BatchFacts* learned);
}  // namespace

int camus_b200_add_gene_trees(camus_b200* ctx, int32_t n_trees, const int64_t* node_off, const int32_t* parent,
                              const int32_t* taxon) {
  if (!ctx) return CAMUS_B200_ERR_ARG;
  if (ctx->n == 0) return ctx->fail(CAMUS_B200_ERR_ARG, "add_gene_trees: set the constraint tree first");
  if (!ctx->subs.empty()) {
    ctx->reduced = false;
    // validated once (on the first device's context), appended everywhere
    BatchFacts facts;
    if (int rc = add_gene_trees_impl(ctx->subs[0], n_trees, node_off, parent, taxon, nullptr, &facts)) return ctx->fail(rc, ctx->subs[0]->err);
    return fan_out(ctx, [&](camus_b200* s, int i) { return i == 0 ? 0 : add_gene_trees_impl(s, n_trees, node_off, parent, taxon, &facts, nullptr); });
  }
  return add_gene_trees_impl(ctx, n_trees, node_off, parent, taxon, nullptr, nullptr);
}

namespace {
int add_gene_trees_impl(camus_b200* ctx, int32_t n_trees, const int64_t* node_off, const int32_t* parent, const int32_t* taxon,
                        const BatchFacts* known, BatchFacts* learned) {
  if (ctx->n == 0) return ctx->fail(CAMUS_B200_ERR_ARG, "add_gene_trees: set the constraint tree first");
  if (n_trees < 0 || !node_off || (n_trees > 0 && (!parent || !taxon))) return ctx->fail(CAMUS_B200_ERR_ARG, "add_gene_trees: null argument");
  // Validation is per tree and independent: large batches are checked on several host threads
  // (16 MB of nodes at 1000 x 1000 took ~10 ms on one). The first failing tree (lowest index) wins.
  struct Part {
    uint64_t res = 0;
    int max_depth = 0, bad_tree = -1;
    bool all_complete = true, all_rooted = true;
    std::string msg;
  };
  const int N = ctx->N;
  auto validate = [&](int g_lo, int g_hi, Part& pt) {
    std::vector<int> stamp(N, -1);
    std::vector<char> has_child;
    std::vector<int> n_child, depth_h;
    auto bad = [&](int g, const std::string& m) {
      pt.bad_tree = g;
      pt.msg = m;
    };
    for (int g = g_lo; g < g_hi; ++g) {
      int64_t b = node_off[g], e = node_off[g + 1];
      int64_t nn = e - b;
      if (nn <= 0 || nn > 65535) return bad(g, "add_gene_trees: tree " + std::to_string(g) + " has " + std::to_string(nn) + " nodes (1..65535 supported)");
      has_child.assign(nn, 0);
      n_child.assign(nn, 0);
      depth_h.assign(nn, 0);
      uint64_t nl = 0;
      for (int64_t i = 0; i < nn; ++i) {
        int p = parent[b + i];
        if ((i == 0) != (p < 0) || p >= i)
          return bad(g, "add_gene_trees: tree " + std::to_string(g) + " is not in preorder (parent before child, single root)");
        if (p >= 0) {
          has_child[p] = 1;
          ++n_child[p];
          depth_h[i] = depth_h[p] + 1;
          if (depth_h[i] >= 65535)
            return bad(g, "add_gene_trees: tree " + std::to_string(g) + " is deeper than 65534 edges (16-bit depth tables)");
          if (depth_h[i] > pt.max_depth) pt.max_depth = depth_h[i];
        }
        int t = taxon[b + i];
        if (t >= N) return bad(g, "add_gene_trees: taxon index out of range");
        if (t >= 0) {
          if (stamp[t] == g) return bad(g, "add_gene_trees: tree " + std::to_string(g) + " contains a taxon twice");
          stamp[t] = g;
          ++nl;
        }
      }
      for (int64_t i = 0; i < nn; ++i)
        if ((taxon[b + i] >= 0) == (has_child[i] != 0))
          return bad(g, "add_gene_trees: tree " + std::to_string(g) + ": taxon set on an internal node or missing on a leaf");
      pt.res += choose4(nl);
      // complete + fully resolved: all taxa present, every internal node binary (the root may be a
      // trifurcation, i.e. the unrooted form of a binary tree)
      bool complete = nl == (uint64_t)N;
      for (int64_t i = 0; i < nn && complete; ++i)
        if (n_child[i] != 0 && n_child[i] != 2 && !(i == 0 && n_child[i] == 3)) complete = false;
      if (!complete) pt.all_complete = false;
      if (n_child[0] != 2) pt.all_rooted = false;
    }
  };
  BatchFacts facts;
  if (known) facts = *known;
  const int64_t total_nodes = n_trees ? node_off[n_trees] - node_off[0] : 0;
  int n_thr = known ? 0 : 1;
  if (!known && total_nodes > (1 << 18)) n_thr = (int)std::max(1u, std::min({std::thread::hardware_concurrency(), 16u, (unsigned)n_trees}));
  std::vector<Part> parts(n_thr);
  if (!known) {
    std::vector<std::thread> th;
    for (int k = 1; k < n_thr; ++k)
      th.emplace_back([&, k] { validate((int)((int64_t)n_trees * k / n_thr), (int)((int64_t)n_trees * (k + 1) / n_thr), parts[k]); });
    validate(0, (int)((int64_t)n_trees / n_thr), parts[0]);
    for (auto& t : th) t.join();
  }
  uint64_t res = 0;
  for (const Part& pt : parts)  // chunks are in tree order: the first bad chunk holds the lowest failing tree
    if (pt.bad_tree >= 0) return ctx->fail(CAMUS_B200_ERR_INPUT, pt.msg);
  for (const Part& pt : parts) {
    facts.res += pt.res;
    facts.max_depth = std::max(facts.max_depth, pt.max_depth);
    if (!pt.all_complete) facts.all_complete = false;
    if (!pt.all_rooted) facts.all_rooted = false;
  }
  res = facts.res;
  ctx->max_depth = std::max(ctx->max_depth, facts.max_depth);
  if (!facts.all_complete) ctx->all_complete = false;
  if (!facts.all_rooted) ctx->all_rooted = false;
  if (learned) *learned = facts;
  int64_t base = ctx->h_node_off.back() - node_off[0];
  for (int g = 0; g < n_trees; ++g) ctx->h_node_off.push_back(node_off[g + 1] + base);
  size_t cnt = n_trees ? (size_t)(node_off[n_trees] - node_off[0]) : 0;
  ctx->h_parent.insert(ctx->h_parent.end(), parent + (n_trees ? node_off[0] : 0), parent + (n_trees ? node_off[0] : 0) + cnt);
  ctx->h_taxon.insert(ctx->h_taxon.end(), taxon + (n_trees ? node_off[0] : 0), taxon + (n_trees ? node_off[0] : 0) + cnt);
  ctx->resolutions += res;
  ctx->G = (int)ctx->h_node_off.size() - 1;
  ctx->genes_dirty = true;
  ctx->cell_state = camus_b200::NONE;
  ctx->z_valid = false;
  ctx->counted = false;
  return 0;
}
}  // namespace

}  // extern "C"

namespace {

// Work list of k_triple_planes_bs: the tiles this context builds (all of them, or the tiles led by the classes
// of its share), with their output slot, in blocks of 16 x 16 x 16 segments: a block of tiles shares
// 3 x 128 x 128 pair rows (6 MB for four tree groups), which stay in L2 while its 4096 tiles stream out.
int make_plane_list(camus_b200* ctx, bool partial) {
  const int key[4] = {ctx->M, partial ? ctx->rank : 0, partial ? ctx->world : 1, partial ? 1 : 0};
  if (!memcmp(key, ctx->plane_list_key, sizeof key)) return 0;
  const std::vector<PlaneDesc> list = plane_work_list((uint32_t)ctx->M, partial ? class_need_mask(ctx->rank, ctx->world) : 0xFu, partial);
  CK(ctx->d_plane_list.alloc(std::max<size_t>(list.size(), 1)));
  if (!list.empty()) CK(cudaMemcpyAsync(ctx->d_plane_list.p, list.data(), list.size() * sizeof(PlaneDesc), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));  // the host vector goes away
  ctx->n_plane_list = (uint32_t)list.size();
  memcpy(ctx->plane_list_key, key, sizeof key);
  return 0;
}

// K0 + K1 + K1b: gene trees -> rooted-triple bit-planes in HBM.
int build_planes(camus_b200* ctx, bool do_upload = true) {
  const int G = ctx->G, npad = ctx->npad;
  ctx->G32 = (G + kLanes - 1) / kLanes;
  const int G32 = std::max(ctx->G32, 1);
  if (int rc = set_box_range(ctx)) return rc;
  // class-scheme partitions read, and therefore build, only the tiles led by their classes
  const bool partial = ctx->class_scheme && !ctx->want_all_tiles;
  const uint32_t* tile_list = partial ? ctx->d_tile_list.p : nullptr;
  if (do_upload) {
    StageTimer t(ctx, CAMUS_B200_T_UPLOAD);
    if (upload(ctx, ctx->d_node_off, ctx->h_node_off.data(), ctx->h_node_off.size()) ||
        upload(ctx, ctx->d_gparent, ctx->h_parent.data(), ctx->h_parent.size()) ||
        upload(ctx, ctx->d_gtaxon, ctx->h_taxon.data(), ctx->h_taxon.size()))
      return CAMUS_B200_ERR_CUDA;
    if (t.stop()) return CAMUS_B200_ERR_CUDA;
  }
  StageTimer t(ctx, CAMUS_B200_T_PREP);
  size_t per_group = (size_t)npad * kLanes;
  CK(ctx->d_depth_tmp.alloc(std::max<size_t>(ctx->h_parent.size(), 1)));
  CK(ctx->d_nleaf.alloc(std::max(G, 1)));
  CK(ctx->d_lrT.alloc(per_group * G32));
  CK(ctx->d_hT.alloc(per_group * G32));
  CK(cudaMemsetAsync(ctx->d_lrT.p, 0, ctx->d_lrT.bytes(), ctx->stream));
  CK(cudaMemsetAsync(ctx->d_hT.p, 0, ctx->d_hT.bytes(), ctx->stream));
  size_t tri_words = (size_t)(partial ? ctx->n_tiles_local : num_tiles(ctx->M)) * G32 * kTileWords;
  if (ctx->d_tri.n < tri_words) {  // only query the driver when the table has to grow
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    if (tri_words * 4 + ((size_t)4 << 30) > free_b + ctx->d_tri.bytes())
      return ctx->fail(CAMUS_B200_ERR_NOMEM, "triple bit-planes need " + std::to_string(tri_words * 4 >> 20) +
                                                 " MiB, device has " + std::to_string(free_b >> 20) + " MiB free");
  }
  CK(ctx->d_tri.alloc(tri_words));
  if (G == 0) CK(cudaMemsetAsync(ctx->d_tri.p, 0, tri_words * 4, ctx->stream));
  if (G > 0) {
    int max_nodes = 1;
    for (int g = 0; g < G; ++g) max_nodes = std::max<int>(max_nodes, (int)(ctx->h_node_off[g + 1] - ctx->h_node_off[g]));
    static const bool serial_prepare = getenv("CAMUS_B200_SERIAL_PREPARE") && getenv("CAMUS_B200_SERIAL_PREPARE")[0] == '1';  // A/B, tests
    const bool warp_prepare = max_nodes <= kTreePrepareMaxNodes && !serial_prepare;
    // bit-sliced plane builder (default): NB = bits of the largest stored value (depth + 1), up to 11
    int need_bits = 1;
    while ((ctx->max_depth + 1) >> need_bits) ++need_bits;
    const bool sliced_path = need_bits <= 11 && ctx->allow_sliced && ctx->allow_half;
    // slices straight from the trees: one [taxon][32 trees] u16 image per CTA in shared memory
    const size_t image_bytes = (size_t)npad * kLanes * (need_bits <= 7 ? 1 : 2);  // DepthImage<NB>::T, NB = 5, 6, 7 | 9, 11
    const bool direct_path = sliced_path && ctx->allow_direct && warp_prepare && image_bytes <= (size_t)kDirectImageMaxBytes;
    if (direct_path) {
      CK(ctx->d_lrP.alloc((size_t)G * npad));
      CK(ctx->d_hP.alloc((size_t)G * npad));
      CK(ctx->d_posP.alloc((size_t)G * npad));
      CK(cudaMemsetAsync(ctx->d_posP.p, 0, (size_t)G * npad * 2, ctx->stream));
    }
    if (warp_prepare) {
      const size_t smem = (size_t)max_nodes * 4;
      if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_tree_prepare_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_tree_prepare_warp<<<G, 32, smem, ctx->stream>>>(G, ctx->d_node_off.p, ctx->d_gparent.p, ctx->d_gtaxon.p,
                                                         ctx->d_rank_of_tip.p, ctx->d_lrT.p, ctx->d_hT.p, npad, ctx->d_nleaf.p,
                                                         max_nodes, direct_path ? ctx->d_lrP.p : nullptr, ctx->d_hP.p, ctx->d_posP.p);
    } else {
      k_tree_prepare<<<(G + 127) / 128, 128, 0, ctx->stream>>>(G, ctx->d_node_off.p, ctx->d_gparent.p, ctx->d_gtaxon.p,
                                                              ctx->d_rank_of_tip.p, ctx->d_depth_tmp.p, ctx->d_lrT.p,
                                                              ctx->d_hT.p, npad, ctx->d_nleaf.p);
    }
    CK_LAUNCH();
    size_t dt_group = (size_t)npad * npad * kLanes;  // u16 elements
    const unsigned n_tiles = partial ? (unsigned)ctx->n_tiles_local : (unsigned)num_tiles(ctx->M);
    // packed fp16 compares need exact depths: depth + 1 <= 2048 (checked in add_gene_trees)
    const bool half_path = ctx->max_depth + 1 <= 2048 && ctx->allow_half;
    const int NB = need_bits <= 5 ? 5 : need_bits <= 6 ? 6 : need_bits <= 7 ? 7 : need_bits <= 9 ? 9 : 11;
    const size_t slice_words = NB < 8 ? 8 : 12;
    // Tree groups go through the depth table in batches of <= 256 MB (measured: running K1 of the next batch
    // on a second stream under K1b of this one gains nothing; 1 GB batches make K1's scattered 2-byte stores
    // miss L2 and double its time).
    const size_t dt_cap = (size_t)256 << 20;
    int gb_max = (int)std::max<size_t>(1, std::min<size_t>(ctx->G32, dt_cap / (dt_group * 2)));
    if (const char* e = getenv("CAMUS_B200_PREP_BATCH")) gb_max = std::max(1, std::min(ctx->G32, atoi(e)));  // tuning
    if (!direct_path) CK(ctx->d_Dt.alloc(dt_group * gb_max));
    if (sliced_path) {
      CK(ctx->d_Ds.alloc((size_t)npad * npad * slice_words * gb_max));
      if (int rc = make_plane_list(ctx, partial)) return rc;
    }
    // rooted triples are all resolved only under a binary root (a trifurcating root resolves
    // every quartet but leaves the triples across its three subtrees unresolved)
    const bool complete = ctx->all_complete && ctx->all_rooted && ctx->allow_complete;
    uint16_t* dt = ctx->d_Dt.p;
    for (int g0 = 0; g0 < ctx->G32; g0 += gb_max) {
      int gb = std::min(gb_max, ctx->G32 - g0);
      // integer depths: 0 = absent; fp16 depths: 0xFFFF = NaN = absent
      const bool fp16_table = half_path && !sliced_path;
      if (!direct_path) {
        CK(cudaMemsetAsync(dt, fp16_table ? 0xFF : 0, dt_group * gb * 2, ctx->stream));
        dim3 blk(kLanes, 8);
        dim3 g1((npad + 7) / 8, gb);
        if (fp16_table)
          k_depth_table<true><<<g1, blk, 0, ctx->stream>>>(g0, G, npad, ctx->d_nleaf.p, ctx->d_lrT.p, ctx->d_hT.p, dt);
        else
          k_depth_table<false><<<g1, blk, 0, ctx->stream>>>(g0, G, npad, ctx->d_nleaf.p, ctx->d_lrT.p, ctx->d_hT.p, dt);
        CK_LAUNCH();
      }
      if (n_tiles == 0) continue;  // a share without boxes
      if (sliced_path) {
        const unsigned n_cta = (ctx->n_plane_list + kPlaneTilesPerCta - 1) / kPlaneTilesPerCta;
        const size_t group_words = (size_t)npad * npad * slice_words;
#define CAMUS_SLICED_PLANES(nb, cmpl)                                                                                          \
  {                                                                                                                            \
    const int smem = 2 * 4 * 3 * SliceArray<SliceRow<nb>::W>::words * 4; /* attribute set in camus_b200_create */             \
    for (int q0 = 0; q0 < gb; q0 += 4)                                                                                         \
      k_triple_planes_bs<nb, cmpl><<<n_cta, 256, smem, ctx->stream>>>(g0 + q0, std::min(4, gb - q0), G32, G, npad,             \
                                                                     ctx->d_Ds.p + (size_t)q0 * group_words, ctx->d_tri.p,    \
                                                                     ctx->d_plane_list.p, ctx->n_plane_list);                 \
  }
#define CAMUS_SLICED(nb)                                                                                                        \
  case nb:                                                                                                                      \
    if (direct_path) {                                                                                                          \
      const int smem_d = npad * kLanes * (int)sizeof(DepthImage<nb>::T);                                                        \
      k_depth_slices_direct<nb><<<dim3(npad, gb), 256, smem_d, ctx->stream>>>(g0, G, npad, ctx->d_nleaf.p, ctx->d_lrP.p,        \
                                                                              ctx->d_hP.p, ctx->d_posP.p, ctx->d_Ds.p);        \
    } else                                                                                                                      \
      k_depth_slices<nb><<<dim3(npad, gb), 256, 0, ctx->stream>>>(gb, npad, dt, ctx->d_Ds.p);                                  \
    if (complete) CAMUS_SLICED_PLANES(nb, true) else CAMUS_SLICED_PLANES(nb, false)             \
    break;
        switch (NB) {
          CAMUS_SLICED(5)
          CAMUS_SLICED(6)
          CAMUS_SLICED(7)
          CAMUS_SLICED(9)
          CAMUS_SLICED(11)
        }
#undef CAMUS_SLICED
#undef CAMUS_SLICED_PLANES
      } else if (half_path && complete)
        k_triple_planes_h2<true><<<n_tiles, 256, 0, ctx->stream>>>(g0, gb, G32, G, npad, dt, ctx->d_tri.p, tile_list);
      else if (half_path)
        k_triple_planes_h2<false><<<n_tiles, 256, 0, ctx->stream>>>(g0, gb, G32, G, npad, dt, ctx->d_tri.p, tile_list);
      else
        k_triple_planes<<<n_tiles, 256, 0, ctx->stream>>>(g0, gb, G32, npad, dt, ctx->d_tri.p, tile_list);
      CK_LAUNCH();
    }
  }
  if (t.stop()) return CAMUS_B200_ERR_CUDA;
  ctx->genes_dirty = false;
  ctx->tri_mask = partial ? class_need_mask(ctx->rank, ctx->world) : 0xFu;
  ctx->tri_compact = partial;
  ctx->tri_key[0] = ctx->M;
  ctx->tri_key[1] = ctx->rank;
  ctx->tri_key[2] = ctx->world;
  return 0;
}

// do the bit-planes in HBM hold every tile the current partition reads?
bool planes_cover(camus_b200* ctx) {
  if (set_box_range(ctx)) return false;
  // compact storage is laid out for one (segments, rank, world): any change of those needs a rebuild
  if (ctx->tri_compact && (!ctx->class_scheme || ctx->tri_key[0] != ctx->M || ctx->tri_key[1] != ctx->rank || ctx->tri_key[2] != ctx->world))
    return false;
  const uint32_t need = ctx->class_scheme ? class_need_mask(ctx->rank, ctx->world) : 0xFu;
  return (need & ~ctx->tri_mask) == 0;
}

const uint32_t* tile_slots(const camus_b200* ctx) { return ctx->tri_compact ? ctx->d_tile_slot.p : nullptr; }

// the complete-tree kernel packs two 16-bit counters per register
bool use_complete(const camus_b200* ctx) { return ctx->all_complete && ctx->allow_complete && ctx->G > 0 && ctx->G <= 65535; }

int make_thr_table(camus_b200* ctx, double threshold) {
  int nt = ctx->G + 2;  // the sum of the two smallest counts never exceeds the number of trees
  CK(ctx->d_thr_tab.alloc(nt));
  k_threshold_table<<<(nt + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_thr_tab.p, nt, threshold);
  CK_LAUNCH();
  return 0;
}

// Which split of the box index space (layout.cuh): the class scheme for 4 or 8 shares of a
// constraint with 32..256 segments, else chunks. CAMUS_B200_PARTITION=chunk|class overrides
// (class still needs 4 or 8 shares). camus_b200/partition.py mirrors the rule.
bool use_class_scheme(const camus_b200* ctx) {
  if (ctx->world != 4 && ctx->world != 8) return false;
  if (ctx->force_class) return ctx->M <= 256;
  if (const char* e = getenv("CAMUS_B200_PARTITION")) {
    if (!strcmp(e, "chunk")) return false;
    if (!strcmp(e, "class")) return ctx->M <= 256;
  }
  return ctx->M >= 32 && ctx->M <= 256;
}

BoxMap box_map(const camus_b200* ctx, uint64_t b0) {
  return BoxMap{b0, ctx->chunk, (uint32_t)ctx->rank, (uint32_t)ctx->world, ctx->class_scheme ? ctx->d_box_list.p : nullptr};
}

int set_box_range(camus_b200* ctx) {
  const uint64_t nb = num_boxes(ctx->M);
  if (!use_class_scheme(ctx)) {
    ctx->class_scheme = false;
    ctx->chunk = partition_chunk(nb, (uint32_t)ctx->world);
    ctx->n_local = local_box_count(nb, ctx->chunk, (uint32_t)ctx->rank, (uint32_t)ctx->world);
    return 0;
  }
  ctx->class_scheme = true;
  ctx->chunk = 1;
  if (ctx->list_key[0] == ctx->M && ctx->list_key[1] == ctx->rank && ctx->list_key[2] == ctx->world) return 0;
  // host-built lists, colex order (= the order of the unpartitioned kernel, filtered)
  const uint32_t S = (uint32_t)ctx->M;
  const int world = ctx->world, rank = ctx->rank;
  const uint32_t thr = class_threshold(S, world), need = class_need_mask(rank, world);
  std::vector<uint8_t> cls(S);
  for (uint32_t s = 0; s < S; ++s) cls[s] = (uint8_t)seg_class(s);
  std::vector<uint32_t> boxes, tiles, slots((size_t)num_tiles(S), 0xFFFFFFFFu);
  boxes.reserve((size_t)(nb / world + nb / (8 * (uint64_t)world) + 64));
  uint32_t idx = 0;
  for (uint32_t sd = 0; sd < S; ++sd)
    for (uint32_t sc = 0; sc <= sd; ++sc) {
      const uint32_t h = class_hash(sc, sd);
      bool allow[4][4];
      for (uint32_t i = 0; i < 4; ++i)
        for (uint32_t j = 0; j < 4; ++j) allow[i][j] = class_owner(i, j, h, world, thr) == rank;
      for (uint32_t sb = 0; sb <= sc; ++sb) {
        const bool* row = allow[cls[sb]];
        for (uint32_t sa = 0; sa <= sb; ++sa, ++idx)
          if (row[cls[sa]]) boxes.push_back(idx);
      }
    }
  idx = 0;
  for (uint32_t s2 = 0; s2 < S; ++s2)
    for (uint32_t s1 = 0; s1 <= s2; ++s1)
      for (uint32_t s0 = 0; s0 <= s1; ++s0, ++idx)
        if ((need >> cls[s0]) & 1u) {
          slots[idx] = (uint32_t)tiles.size();
          tiles.push_back(idx);
        }
  CK(ctx->d_box_list.alloc(std::max<size_t>(boxes.size(), 1)));
  CK(ctx->d_tile_list.alloc(std::max<size_t>(tiles.size(), 1)));
  CK(ctx->d_tile_slot.alloc(slots.size()));
  CK(cudaMemcpyAsync(ctx->d_tile_slot.p, slots.data(), slots.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (!boxes.empty()) CK(cudaMemcpyAsync(ctx->d_box_list.p, boxes.data(), boxes.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (!tiles.empty()) CK(cudaMemcpyAsync(ctx->d_tile_list.p, tiles.data(), tiles.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));  // the host vectors go away
  ctx->n_local = boxes.size();
  ctx->n_tiles_local = tiles.size();
  ctx->list_key[0] = ctx->M;
  ctx->list_key[1] = ctx->rank;
  ctx->list_key[2] = ctx->world;
  return 0;
}

// Fused K2+K3+K4 over this rank's box range: counts stay in registers, survivors are scored
// straight into the junction table Z. No cell table.
// pass_flags (passes mode): bit 0 = add to the junction table and the scalars of the earlier passes,
// bit 1 = more passes follow (the private copies of Z are summed after the last one only)
int run_fused(camus_b200* ctx, int qmode, double threshold, int as_set, uint64_t* ns, uint64_t* sum, int pass_flags = 0) {
  if (int rc = set_box_range(ctx)) return rc;
  uint64_t mine = ctx->n_local;
  size_t zn = (size_t)ctx->n * ctx->n * 3;
  static const int rep_env = getenv("CAMUS_B200_ZREP") ? atoi(getenv("CAMUS_B200_ZREP")) : 0;
  // Privatised copies of Z spread the hot (w-point, junction) entries over several L2 lines. The
  // boxes in flight at any time differ mostly in their last segment and add to the same entries
  // for their other three, so one copy per segment (up to 128) is what removes the collisions
  // (measured at 125 segments: 1 copy 2098 ms, 8: 1066, 32: 974, 64: 962, 128: 956).
  int rep = 8;
  while (rep < ctx->M && rep < 128) rep <<= 1;
  if (rep_env > 0) rep = rep_env;
  while (rep > 1 && zn * rep * 8 > ((size_t)16 << 30)) rep >>= 1;
  for (;;) {  // shrink if the bit-planes left too little memory
    cudaError_t e = ctx->d_Z.alloc(zn * rep);
    if (e == cudaSuccess) break;
    cudaGetLastError();
    if (rep == 1) CK(e);
    rep >>= 1;
  }
  CK(ctx->d_totals.alloc(2));
  StageTimer t(ctx, CAMUS_B200_T_COUNT);
  if (make_thr_table(ctx, threshold)) return CAMUS_B200_ERR_CUDA;
  if (!(pass_flags & 1)) {
    CK(cudaMemsetAsync(ctx->d_Z.p, 0, zn * rep * 8, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_totals.p, 0, 16, ctx->stream));
    ctx->z_rep_used = rep;
  } else {
    rep = ctx->z_rep_used;  // the copies the first pass zeroed
  }
  static const int dbg = getenv("CAMUS_B200_DEBUG") ? atoi(getenv("CAMUS_B200_DEBUG")) : 0;
  FusedParams fz{qmode, threshold, ctx->d_thr_tab.p, as_set, ctx->d_Z.p, ctx->d_totals.p, rep, (unsigned long long)zn, ctx->allow_regular ? 1 : 0, dbg};
  // box descriptors (segments + regular flag), rebuilt only when the constraint or the partition changed
  static const bool no_desc = getenv("CAMUS_B200_NO_BOXDESC") && getenv("CAMUS_B200_NO_BOXDESC")[0] == '1';  // A/B
  const uint2* desc = nullptr;
  if (!no_desc && mine > 0 && ctx->M < 4096) {
    const uint64_t key[6] = {ctx->constraint_gen, (uint64_t)ctx->M, (uint64_t)ctx->rank, (uint64_t)ctx->world,
                             ctx->class_scheme ? 1u : 0u, ctx->allow_regular ? 1u : 0u};
    bool have = ctx->d_box_desc.p != nullptr && memcmp(key, ctx->desc_key, sizeof key) == 0;
    if (!have) {
      cudaError_t e = ctx->d_box_desc.alloc(mine);
      if (e == cudaSuccess) {
        k_box_desc<<<(unsigned)((mine + 255) / 256), 256, 0, ctx->stream>>>(box_map(ctx, 0), mine, ctx->ct, ctx->allow_regular ? 1 : 0, ctx->d_box_desc.p);
        CK_LAUNCH();
        memcpy(ctx->desc_key, key, sizeof key);
        have = true;
      } else {
        cudaGetLastError();  // no room: the kernel unranks its box itself
        memset(ctx->desc_key, 0, sizeof key);
      }
    }
    if (have) desc = ctx->d_box_desc.p;
  }
  const bool complete = use_complete(ctx);
  // incomplete / polytomous trees: up to 1023 of them fit three 10-bit counters in one register
  const bool packed = !complete && ctx->allow_packed && ctx->G <= 1023;
  const uint64_t max_grid = 1u << 30;
  for (uint64_t b0 = 0; b0 < mine; b0 += max_grid) {
    uint64_t nbx = std::min(max_grid, mine - b0);
    CountParams prm{ctx->d_tri.p, std::max(ctx->G32, 1), box_map(ctx, b0), nullptr, ctx->G, tile_slots(ctx), desc};
    static const int pad_smem = getenv("CAMUS_B200_COUNT_PAD_SMEM") ? atoi(getenv("CAMUS_B200_COUNT_PAD_SMEM")) : 0;  // experiments: fewer CTAs per SM
    if (complete)
      k_count<true, kModeComplete><<<(unsigned)nbx, kCountThreads, CountCfg<kModeComplete>::kSmemBytes + pad_smem, ctx->stream>>>(prm, fz, ctx->ct);
    else if (packed)
      k_count<true, kModePacked><<<(unsigned)nbx, kCountThreads, CountCfg<kModePacked>::kSmemBytes, ctx->stream>>>(prm, fz, ctx->ct);
    else
      k_count<true, kModeGeneral><<<(unsigned)nbx, kCountThreads, CountCfg<kModeGeneral>::kSmemBytes, ctx->stream>>>(prm, fz, ctx->ct);
    CK_LAUNCH();
  }
  if (rep > 1 && !(pass_flags & 2)) {
    k_sum_replicas<<<(unsigned)((zn + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_Z.p, zn, rep);
    CK_LAUNCH();
  }
  unsigned long long h[2] = {0, 0};
  CK(cudaMemcpyAsync(h, ctx->d_totals.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
  if (t.stop()) return CAMUS_B200_ERR_CUDA;
  if (ns) *ns = h[0];
  if (sum) *sum = h[1];
  ctx->last_qmode = qmode;
  ctx->last_threshold = threshold;
  ctx->z_valid = true;
  ctx->z_reduced = false;
  ctx->z_as_set = as_set;
  ctx->counted = true;
  return 0;
}

// Inputs beyond the bit-plane budget (12 * tiles * ceil(G/32) * 512 bytes): the 8 shares of the class
// scheme (layout.cuh) are run one after the other on this GPU. A share reads, and so builds, only the
// tiles led by 2 of the 4 segment classes = half of the table; the junction table and the two scalars
// accumulate over the passes (integer sums: the result is the one-pass result). Chosen automatically
// when the whole table does not fit beside the junction tables; CAMUS_B200_PASSES=8 / =1 force it.
int wanted_passes(camus_b200* ctx) {
  const int env = ctx->passes_env;
  if (!ctx->use_fused || ctx->world != 1 || !ctx->subs.empty() || ctx->comm || ctx->M > 256 || ctx->want_all_tiles) return 1;
  if (env == 8 || env == 1) return env;
  const size_t full = (size_t)num_tiles(ctx->M) * std::max((ctx->G + kLanes - 1) / kLanes, 1) * kTileBytes;
  const size_t zn = (size_t)ctx->n * ctx->n * 3 * 8;
  const size_t z = std::min<size_t>(zn * 128, (size_t)16 << 30);
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return 1;
  const size_t avail = free_b + ctx->d_tri.bytes() + ctx->d_Z.bytes();
  return full + z + ((size_t)4 << 30) > avail ? 8 : 1;
}

int run_passes(camus_b200* ctx, int passes, int qmode, double threshold, int as_set, bool do_upload, uint64_t* ns, uint64_t* sum) {
  ctx->passes = passes;
  int rc = 0;
  for (int p = 0; p < passes && !rc; ++p) {
    ctx->rank = p;
    ctx->world = passes;
    ctx->force_class = true;
    rc = build_planes(ctx, do_upload && p == 0);
    if (!rc) rc = run_fused(ctx, qmode, threshold, as_set, ns, sum, (p > 0 ? 1 : 0) | (p + 1 < passes ? 2 : 0));
  }
  ctx->rank = 0;
  ctx->world = 1;
  ctx->force_class = false;
  return rc;
}

// K2 (table mode) over this rank's box range (whole range resident in d_cells).
int run_count(camus_b200* ctx) {
  // the bit-planes may belong to one share of a passes run (or to another partition): the table pipeline needs this partition's
  if (!planes_cover(ctx))
    if (int rc = build_planes(ctx, false)) return rc;
  if (int rc = set_box_range(ctx)) return rc;
  uint64_t mine = ctx->n_local;
  if (ctx->d_cells.n < mine * kBoxCells) {
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    if (mine > cells_budget_boxes(free_b + ctx->d_cells.bytes()))
      return ctx->fail(CAMUS_B200_ERR_NOMEM, "cell table of " + std::to_string(mine * kBoxCells * 16 >> 20) +
                                                 " MiB does not fit on the device (only the fused path scales to this size)");
  }
  CK(ctx->d_cells.alloc(std::max<uint64_t>(mine, 1) * kBoxCells));
  StageTimer t(ctx, CAMUS_B200_T_COUNT);
  const uint64_t max_grid = 1u << 30;
  for (uint64_t b0 = 0; b0 < mine; b0 += max_grid) {
    uint64_t nbx = std::min(max_grid, mine - b0);
    CountParams prm{ctx->d_tri.p, std::max(ctx->G32, 1), box_map(ctx, b0), ctx->d_cells.p + b0 * kBoxCells, ctx->G, tile_slots(ctx), nullptr};
    if (use_complete(ctx))
      k_count<false, kModeComplete><<<(unsigned)nbx, kCountThreads, CountCfg<kModeComplete>::kSmemBytes, ctx->stream>>>(prm, FusedParams{}, ctx->ct);
    else
      k_count<false, kModeGeneral><<<(unsigned)nbx, kCountThreads, CountCfg<kModeGeneral>::kSmemBytes, ctx->stream>>>(prm, FusedParams{}, ctx->ct);
    CK_LAUNCH();
  }
  if (t.stop()) return CAMUS_B200_ERR_CUDA;
  ctx->cell_state = camus_b200::RAW;
  return 0;
}

int run_filter(camus_b200* ctx, int qmode, double threshold, uint64_t* ns, uint64_t* sum) {
  uint64_t mine = ctx->n_local;
  CK(ctx->d_totals.alloc(2));
  StageTimer t(ctx, CAMUS_B200_T_FILTER);
  if (make_thr_table(ctx, threshold)) return CAMUS_B200_ERR_CUDA;
  CK(cudaMemsetAsync(ctx->d_totals.p, 0, 16, ctx->stream));
  const uint64_t max_boxes = 1u << 26;
  for (uint64_t b0 = 0; b0 < mine; b0 += max_boxes) {
    uint64_t nbx = std::min(max_boxes, mine - b0);
    FilterParams prm{ctx->d_cells.p + b0 * kBoxCells, box_map(ctx, b0), qmode, threshold, ctx->d_thr_tab.p, ctx->d_totals.p};
    k_filter<<<(unsigned)(nbx * 16), 256, 0, ctx->stream>>>(prm, ctx->ct);
    CK_LAUNCH();
  }
  unsigned long long h[2] = {0, 0};
  CK(cudaMemcpyAsync(h, ctx->d_totals.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
  if (t.stop()) return CAMUS_B200_ERR_CUDA;
  if (ns) *ns = h[0];
  if (sum) *sum = h[1];
  ctx->cell_state = camus_b200::FILTERED;
  ctx->last_qmode = qmode;
  ctx->last_threshold = threshold;
  ctx->counted = true;
  return 0;
}

int ensure_filtered(camus_b200* ctx) {
  if (ctx->cell_state == camus_b200::FILTERED) return 0;
  if (!ctx->counted) return ctx->fail(CAMUS_B200_ERR_ARG, "call camus_b200_count_filter first");
  int rc = run_count(ctx);
  if (rc) return rc;
  return run_filter(ctx, ctx->last_qmode, ctx->last_threshold, nullptr, nullptr);
}

int run_score(camus_b200* ctx, int as_set) {
  if (!ctx->counted) return ctx->fail(CAMUS_B200_ERR_ARG, "call camus_b200_count_filter first");
  if (ctx->z_valid && ctx->z_as_set == as_set) return 0;
  if (ctx->use_fused) {
    const int passes = wanted_passes(ctx);
    if (passes > 1) return run_passes(ctx, passes, ctx->last_qmode, ctx->last_threshold, as_set, false, nullptr, nullptr);
    if (!planes_cover(ctx))  // the last run went through the passes: its bit-planes are one share's
      if (int rc = build_planes(ctx, false)) return rc;
    return run_fused(ctx, ctx->last_qmode, ctx->last_threshold, as_set, nullptr, nullptr);
  }
  int rc = ensure_filtered(ctx);
  if (rc) return rc;
  size_t zn = (size_t)ctx->n * ctx->n * 3;
  CK(ctx->d_Z.alloc(zn));
  StageTimer t(ctx, CAMUS_B200_T_SCORE);
  CK(cudaMemsetAsync(ctx->d_Z.p, 0, zn * 8, ctx->stream));
  uint64_t mine = ctx->n_local;
  const uint64_t max_boxes = 1u << 26;
  for (uint64_t b0 = 0; b0 < mine; b0 += max_boxes) {
    uint64_t nbx = std::min(max_boxes, mine - b0);
    ScoreParams prm{ctx->d_cells.p + b0 * kBoxCells, box_map(ctx, b0), as_set, ctx->d_Z.p};
    k_score<<<(unsigned)(nbx * 16), 256, 0, ctx->stream>>>(prm, ctx->ct);
    CK_LAUNCH();
  }
  if (t.stop()) return CAMUS_B200_ERR_CUDA;
  ctx->z_valid = true;
  ctx->z_reduced = false;
  ctx->z_as_set = as_set;
  return 0;
}

// ---- NCCL, loaded at run time: the library has no link-time dependency on it ----
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
  bool ok() const { return lib && GetUniqueId && CommInitRank && CommDestroy && AllReduce && GetErrorString; }
};
NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* env = getenv("CAMUS_B200_NCCL_LIB");
    for (const char* name : {env ? env : "libnccl.so.2", "libnccl.so.2", "libnccl.so"}) {
      api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (!api.lib) {
      api.err = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "unknown error");
      return;
    }
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.lib, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.lib, "ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.lib, "ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(api.lib, "ncclAllReduce"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.lib, "ncclGetErrorString"));
    if (!api.ok()) api.err = "libnccl.so.2 lacks a required symbol";
  });
  return api;
}

void nccl_api_destroy(ncclComm_t comm) {
  NcclApi& api = nccl_api();
  if (api.ok()) api.CommDestroy(comm);
}

// Sum the partial junction tables of all ranks in place (one all-reduce over NVLink).
int comm_reduce_table(camus_b200* ctx) {
  if (!ctx->comm || ctx->z_reduced) return 0;
  if (!ctx->z_valid) return ctx->fail(CAMUS_B200_ERR_ARG, "no partial table to reduce");
  NcclApi& api = nccl_api();
  const size_t zn = (size_t)ctx->n * ctx->n * 3;
  StageTimer t(ctx, CAMUS_B200_T_FINAL);
  ncclResult_t r = api.AllReduce(ctx->d_Z.p, ctx->d_Z.p, zn, ncclInt64, ncclSum, ctx->comm, ctx->stream);
  if (r != ncclSuccess) return ctx->fail(CAMUS_B200_ERR_CUDA, std::string("ncclAllReduce: ") + api.GetErrorString(r));
  if (t.stop()) return CAMUS_B200_ERR_CUDA;
  ctx->z_reduced = true;
  return 0;
}
int comm_reduce_scalars(camus_b200* ctx, uint64_t* ns, uint64_t* sum) {
  if (!ctx->comm) return 0;
  NcclApi& api = nccl_api();
  unsigned long long h[2] = {ns ? *ns : 0, sum ? *sum : 0};
  CK(ctx->d_scal.alloc(2));
  CK(cudaMemcpyAsync(ctx->d_scal.p, h, 16, cudaMemcpyHostToDevice, ctx->stream));
  ncclResult_t r = api.AllReduce(ctx->d_scal.p, ctx->d_scal.p, 2, ncclUint64, ncclSum, ctx->comm, ctx->stream);
  if (r != ncclSuccess) return ctx->fail(CAMUS_B200_ERR_CUDA, std::string("ncclAllReduce: ") + api.GetErrorString(r));
  CK(cudaMemcpyAsync(h, ctx->d_scal.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (ns) *ns = h[0];
  if (sum) *sum = h[1];
  return 0;
}

// d_out -> caller memory. Measured at 32 MB (n = 1999): a direct copy into the caller's pageable
// buffer 2.3 ms; through a pinned staging buffer + memcpy 12.5 ms. So: direct.
int copy_out(camus_b200* ctx, uint64_t* out, size_t count) {
  CK(cudaMemcpyAsync(out, ctx->d_out.p, count * 8, cudaMemcpyDeviceToHost, ctx->stream));
  return 0;
}

int run_finish(camus_b200* ctx, uint64_t* out) {
  if (!ctx->z_valid) return ctx->fail(CAMUS_B200_ERR_ARG, "quartet_totals_finish: no partial table (call quartet_totals_partial)");
  const int n = ctx->n;
  size_t en = (size_t)(n + 1) * (n + 1);
  CK(ctx->d_E.alloc(en));
  CK(ctx->d_out.alloc((size_t)n * n));
  StageTimer t(ctx, CAMUS_B200_T_FINAL);
  CK(cudaMemsetAsync(ctx->d_E.p, 0, en * 8, ctx->stream));
  k_expand<<<dim3((n + 127) / 128, n), 128, 0, ctx->stream>>>(ctx->d_Z.p, ctx->d_E.p, ctx->ct);
  CK_LAUNCH();
  k_row_scan<<<n, 256, 0, ctx->stream>>>(ctx->d_E.p, n);
  CK_LAUNCH();
  k_col_scan<<<(n + 31) / 32, 1024, 0, ctx->stream>>>(ctx->d_E.p, n);
  CK_LAUNCH();
  k_mask_out<<<dim3((n + 127) / 128, n), 128, 0, ctx->stream>>>(ctx->d_E.p, ctx->d_out.p, ctx->ct);
  CK_LAUNCH();
  if (int rc = copy_out(ctx, out, (size_t)n * n)) return rc;
  if (t.stop()) return CAMUS_B200_ERR_CUDA;
  return 0;
}

// Sum the devices' partial junction tables into subs[0]'s (copy 0 of its Z). With peer access one
// kernel on device 0 reads the other tables straight over NVLink; otherwise they are copied in.
__global__ void k_add_peers(unsigned long long* __restrict__ z, const unsigned long long* const* __restrict__ peers, int n_peers,
                            size_t zn) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= zn) return;
  unsigned long long acc = z[i];
  for (int p = 0; p < n_peers; ++p) acc += peers[p][i];
  z[i] = acc;
}

int reduce_partials(camus_b200* wrap) {
  camus_b200* ctx = wrap->subs[0];  // CK reports into the leader's context
  const size_t zn = (size_t)ctx->n * ctx->n * 3;
  for (camus_b200* s : wrap->subs)
    if (!s->z_valid) return wrap->fail(CAMUS_B200_ERR_ARG, "reduce_partials: a device has no partial table");
  int rc = [&]() -> int {
    CK(cudaSetDevice(ctx->device));
    StageTimer t(ctx, CAMUS_B200_T_FINAL);
    const int np = (int)wrap->subs.size() - 1;
    if (wrap->peer_access) {
      std::vector<const unsigned long long*> h(np);
      for (int i = 0; i < np; ++i) h[i] = wrap->subs[i + 1]->d_Z.p;
      CK(wrap->d_stage.alloc(np));
      CK(cudaMemcpyAsync(wrap->d_stage.p, h.data(), np * sizeof(void*), cudaMemcpyHostToDevice, ctx->stream));
      k_add_peers<<<(unsigned)((zn + 255) / 256), 256, 0, ctx->stream>>>(
          ctx->d_Z.p, reinterpret_cast<const unsigned long long* const*>(wrap->d_stage.p), np, zn);
      CK_LAUNCH();
      CK(cudaStreamSynchronize(ctx->stream));  // h must outlive the copy
    } else {
      CK(wrap->d_stage.alloc(zn + 1));
      for (int i = 0; i < np; ++i) {
        const unsigned long long* one = wrap->d_stage.p;
        CK(cudaMemcpyPeerAsync(wrap->d_stage.p, ctx->device, wrap->subs[i + 1]->d_Z.p, wrap->subs[i + 1]->device, zn * 8, ctx->stream));
        CK(cudaMemcpyAsync(wrap->d_stage.p + zn, &one, sizeof(void*), cudaMemcpyHostToDevice, ctx->stream));
        k_add_peers<<<(unsigned)((zn + 255) / 256), 256, 0, ctx->stream>>>(
            ctx->d_Z.p, reinterpret_cast<const unsigned long long* const*>(wrap->d_stage.p + zn), 1, zn);
        CK_LAUNCH();
        CK(cudaStreamSynchronize(ctx->stream));
      }
    }
    if (t.stop()) return CAMUS_B200_ERR_CUDA;
    return 0;
  }();
  if (rc) return wrap->fail(rc, ctx->err);
  wrap->reduced = true;
  return 0;
}

}  // namespace

extern "C" {

int camus_b200_count_filter(camus_b200* ctx, int qmode, double threshold, uint64_t* n_survivors, uint64_t* sum_counts) {
  if (!ctx) return CAMUS_B200_ERR_ARG;
  if (ctx->n == 0) return ctx->fail(CAMUS_B200_ERR_ARG, "count_filter: set the constraint tree first");
  if (qmode < 0 || qmode > 2 || !(threshold >= 0 && threshold <= 1))
    return ctx->fail(CAMUS_B200_ERR_ARG, "count_filter: quartet mode in [0,2] and threshold in [0,1] (quartetfilter.go:41-61)");
  if (ctx->G >= (1 << 24)) return ctx->fail(CAMUS_B200_ERR_ARG, "count_filter: at most 16 777 215 gene trees (32-bit per-thread survivor sums)");
  if (!ctx->subs.empty()) {
    std::vector<uint64_t> ns(ctx->subs.size(), 0), sc(ctx->subs.size(), 0);
    ctx->reduced = false;
    int rc = fan_out(ctx, [&](camus_b200* s, int i) { return camus_b200_count_filter(s, qmode, threshold, &ns[i], &sc[i]); });
    if (rc) return rc;
    uint64_t a = 0, b = 0;
    for (size_t i = 0; i < ns.size(); ++i) {
      a += ns[i];
      b += sc[i];
    }
    if (n_survivors) *n_survivors = a;
    if (sum_counts) *sum_counts = b;
    return 0;
  }
  CK(cudaSetDevice(ctx->device));
  std::memset(ctx->ms, 0, sizeof(ctx->ms));
  ctx->launches = 0;
  ctx->want_all_tiles = false;  // a camus_b200_query_cells probe asks again
  int rc;
  const int passes = wanted_passes(ctx);
  if (passes > 1) {
    ctx->cell_state = camus_b200::NONE;
    ctx->z_valid = false;
    uint64_t pns = 0, psc = 0;
    if ((rc = run_passes(ctx, passes, qmode, threshold, ctx->as_set_hint, ctx->genes_dirty, &pns, &psc))) return rc;
    if (n_survivors) *n_survivors = pns;
    if (sum_counts) *sum_counts = psc;
    return 0;
  }
  ctx->passes = 1;
  if (ctx->genes_dirty) {
    if ((rc = build_planes(ctx, true))) return rc;
  } else if (!planes_cover(ctx) && (rc = build_planes(ctx, false))) {  // the partition changed under resident trees
    return rc;
  }
  ctx->cell_state = camus_b200::NONE;
  ctx->z_valid = false;
  uint64_t ns = 0, sc = 0;
  if (ctx->use_fused) {
    if ((rc = run_fused(ctx, qmode, threshold, ctx->as_set_hint, &ns, &sc))) return rc;
  } else {
    if ((rc = run_count(ctx))) return rc;
    if ((rc = run_filter(ctx, qmode, threshold, &ns, &sc))) return rc;
  }
  if ((rc = comm_reduce_scalars(ctx, &ns, &sc))) return rc;  // with a communicator: the global values
  if (n_survivors) *n_survivors = ns;
  if (sum_counts) *sum_counts = sc;
  return 0;
}

int camus_b200_set_tip_labels(camus_b200* ctx, int32_t n_taxa, const char* labels, const int64_t* label_off) {
  if (!ctx) return CAMUS_B200_ERR_ARG;
  if (!ctx->subs.empty()) return fan_out(ctx, [&](camus_b200* s, int) { return camus_b200_set_tip_labels(s, n_taxa, labels, label_off); });
  if (ctx->n == 0 || n_taxa != ctx->N) return ctx->fail(CAMUS_B200_ERR_ARG, "set_tip_labels: set the constraint tree first; one label per constraint taxon");
  if (!labels || !label_off || label_off[0] != 0) return ctx->fail(CAMUS_B200_ERR_ARG, "set_tip_labels: null argument");
  CK(cudaSetDevice(ctx->device));
  unsigned int cap = 16;
  while (cap < 2u * (unsigned)n_taxa) cap <<= 1;
  std::vector<unsigned long long> hs(cap, 0);
  std::vector<int> tips(cap, -1);
  for (int t = 0; t < n_taxa; ++t) {
    const int64_t b = label_off[t], e = label_off[t + 1];
    if (e < b) return ctx->fail(CAMUS_B200_ERR_ARG, "set_tip_labels: offsets must not decrease");
    unsigned long long h = 14695981039346656037ull;  // FNV-1a, as nwk_hash
    for (int64_t i = b; i < e; ++i) {
      h ^= (unsigned char)labels[i];
      h *= 1099511628211ull;
    }
    if (!h) h = 1;
    unsigned int s = (unsigned int)h & (cap - 1);
    while (hs[s]) {
      if (hs[s] == h) {
        const int o = tips[s];
        if (label_off[o + 1] - label_off[o] == e - b && std::memcmp(labels + label_off[o], labels + b, (size_t)(e - b)) == 0)
          return ctx->fail(CAMUS_B200_ERR_INPUT, "set_tip_labels: the same label twice");
      }
      s = (s + 1) & (cap - 1);
    }
    hs[s] = h;
    tips[s] = t;
  }
  std::vector<long long> off(label_off, label_off + n_taxa + 1);
  if (upload(ctx, ctx->d_lab_hash, hs.data(), hs.size()) || upload(ctx, ctx->d_lab_tip, tips.data(), tips.size()) ||
      upload(ctx, ctx->d_lab_off, off.data(), off.size()) || upload(ctx, ctx->d_lab_pool, labels, (size_t)std::max<int64_t>(label_off[n_taxa], 1)))
    return CAMUS_B200_ERR_CUDA;
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->lab_mask = cap - 1;
  ctx->lab_gen = ctx->constraint_gen;
  return 0;
}

int camus_b200_add_gene_trees_newick(camus_b200* ctx, const char* text, uint64_t len, double min_support, int32_t* n_trees_out,
                                     int32_t* missing_taxa, int32_t* bad_tree) {
  if (!ctx) return CAMUS_B200_ERR_ARG;
  if (n_trees_out) *n_trees_out = 0;
  if (missing_taxa) *missing_taxa = 0;
  if (bad_tree) *bad_tree = 0;
  if (!ctx->subs.empty()) {
    ctx->reduced = false;
    std::vector<int32_t> nt(ctx->subs.size(), 0), mt(ctx->subs.size(), 0), bt(ctx->subs.size(), 0);
    int rc = fan_out(ctx, [&](camus_b200* s, int i) { return camus_b200_add_gene_trees_newick(s, text, len, min_support, &nt[i], &mt[i], &bt[i]); });
    if (n_trees_out) *n_trees_out = nt[0];
    if (missing_taxa) *missing_taxa = mt[0];
    if (bad_tree) *bad_tree = bt[0];
    return rc;
  }
  if (ctx->n == 0 || ctx->lab_gen != ctx->constraint_gen)
    return ctx->fail(CAMUS_B200_ERR_ARG, "add_gene_trees_newick: call camus_b200_set_constraint and camus_b200_set_tip_labels first");
  if (!text || len == 0 || len > ((uint64_t)1 << 31) - 64) return ctx->fail(CAMUS_B200_ERR_ARG, "add_gene_trees_newick: 1 byte .. 2 GiB of text");
  if (!(min_support == min_support)) return ctx->fail(CAMUS_B200_ERR_ARG, "add_gene_trees_newick: min_support is NaN");
  CK(cudaSetDevice(ctx->device));
  // non-blank lines (io.go:138-152: one tree per line), trimmed; finding them is a memchr walk
  std::vector<long long> lb, le;
  for (uint64_t start = 0; start < len;) {
    const char* nl = (const char*)memchr(text + start, '\n', len - start);
    uint64_t end = nl ? (uint64_t)(nl - text) : len, b = start, e = end;
    while (b < e && (unsigned char)text[b] <= ' ') ++b;
    while (e > b && (unsigned char)text[e - 1] <= ' ') --e;
    if (e > b) {
      lb.push_back((long long)b);
      le.push_back((long long)e);
    }
    start = end + 1;
  }
  const int G = (int)lb.size();
  if (G == 0) return ctx->fail(CAMUS_B200_ERR_FALLBACK, "add_gene_trees_newick: no tree in the text");
  StageTimer t(ctx, CAMUS_B200_T_UPLOAD);
  // scratch lives in the context: cudaMalloc / cudaFree per call cost more than the two kernels
  DevBuf<char>& d_text = ctx->nwk_text;
  DevBuf<long long>&d_lb = ctx->nwk_lb, &d_le = ctx->nwk_le, &d_off = ctx->nwk_off;
  DevBuf<int>&d_stack = ctx->nwk_stack, &d_par = ctx->nwk_par, &d_tax = ctx->nwk_tax, &d_nchild = ctx->nwk_nchild;
  DevBuf<unsigned char>& d_coll = ctx->nwk_coll;
  DevBuf<unsigned int>& d_seen = ctx->nwk_seen;
  DevBuf<NwkTreeInfo>& d_info = ctx->nwk_info;
  if (upload(ctx, d_text, text, (size_t)len) || upload(ctx, d_lb, lb.data(), lb.size()) || upload(ctx, d_le, le.data(), le.size())) return CAMUS_B200_ERR_CUDA;
  CK(d_stack.alloc((size_t)len));
  CK(d_coll.alloc((size_t)len));
  CK(d_info.alloc(G));
  CK(d_off.alloc((size_t)G + 1));
  const size_t words = (size_t)(ctx->N + 31) / 32;
  CK(d_seen.alloc((size_t)G * words));
  CK(cudaMemsetAsync(d_seen.p, 0, (size_t)G * words * 4, ctx->stream));
  const int tb = 32, gb = G;  // one warp per tree
  k_nwk_pass1<<<gb, tb, 0, ctx->stream>>>(G, d_text.p, d_lb.p, d_le.p, min_support, d_stack.p, d_coll.p, d_info.p);
  CK_LAUNCH();
  const long long base = (long long)ctx->h_node_off.back();
  k_nwk_offsets<<<1, 1024, 0, ctx->stream>>>(G, d_info.p, base, d_off.p);
  CK_LAUNCH();
  long long total_end = 0;
  CK(cudaMemcpyAsync(&total_end, d_off.p + G, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  const size_t cnt = (size_t)(total_end - base);
  CK(d_par.alloc(std::max<size_t>(cnt, 1)));
  CK(d_tax.alloc(std::max<size_t>(cnt, 1)));
  CK(d_nchild.alloc(std::max<size_t>(cnt, 1)));
  LabelTable lt{ctx->d_lab_hash.p, ctx->d_lab_tip.p, ctx->d_lab_off.p, ctx->d_lab_pool.p, ctx->lab_mask};
  k_nwk_pass2<<<gb, tb, 0, ctx->stream>>>(G, d_text.p, d_lb.p, d_le.p, d_coll.p, d_stack.p, lt, ctx->N, d_off.p, base, d_par.p, d_tax.p, d_nchild.p,
                                         d_seen.p, d_info.p);
  CK_LAUNCH();
  std::vector<NwkTreeInfo> info(G);
  std::vector<long long> off((size_t)G + 1);
  CK(cudaMemcpyAsync(info.data(), d_info.p, (size_t)G * sizeof(NwkTreeInfo), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(off.data(), d_off.p, ((size_t)G + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  static const char* why[] = {"ok", "not in the plain newick dialect or malformed", "single-quoted label", "support value that is not a plain decimal",
                              "label that is not in the constraint tree (ErrTipNameMismatch)", "label twice in one tree (ErrMulTree)", "too deep"};
  uint64_t res = 0;
  int max_depth = 0;
  bool all_complete = true, all_rooted = true, missing = false;
  for (int g = 0; g < G; ++g) {
    const NwkTreeInfo& ti = info[g];
    int st = ti.status;
    if (st == kNwkOk && (ti.n_nodes <= 0 || ti.n_nodes > 65535 || ti.max_depth >= 65535)) st = kNwkTooDeep;
    if (st != kNwkOk) {
      if (bad_tree) *bad_tree = g + 1;
      return ctx->fail(CAMUS_B200_ERR_FALLBACK, "add_gene_trees_newick: gene tree " + std::to_string(g + 1) + ", byte " + std::to_string(ti.err_pos) + ": " +
                                                    why[st < 7 ? st : 1] + "; nothing was added, parse it on the host");
    }
    res += choose4((uint64_t)ti.n_leaves);
    max_depth = std::max(max_depth, ti.max_depth);
    if (ti.n_leaves != ctx->N) missing = true;
    if (ti.n_leaves != ctx->N || !ti.binary) all_complete = false;
    if (ti.root_children != 2) all_rooted = false;
  }
  const size_t old = ctx->h_parent.size();
  ctx->h_parent.resize(old + cnt);
  ctx->h_taxon.resize(old + cnt);
  if (cnt) {
    CK(cudaMemcpyAsync(ctx->h_parent.data() + old, d_par.p, cnt * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_taxon.data() + old, d_tax.p, cnt * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  if (t.stop()) return CAMUS_B200_ERR_CUDA;
  for (int g = 0; g < G; ++g) ctx->h_node_off.push_back((int64_t)off[g + 1]);
  ctx->resolutions += res;
  ctx->max_depth = std::max(ctx->max_depth, max_depth);
  if (!all_complete) ctx->all_complete = false;
  if (!all_rooted) ctx->all_rooted = false;
  ctx->G = (int)ctx->h_node_off.size() - 1;
  ctx->genes_dirty = true;
  ctx->cell_state = camus_b200::NONE;
  ctx->z_valid = false;
  ctx->counted = false;
  if (n_trees_out) *n_trees_out = G;
  if (missing_taxa) *missing_taxa = missing ? 1 : 0;
  return 0;
}

int camus_b200_get_gene_trees(camus_b200* ctx, int64_t* node_off, int32_t* parent, int32_t* taxon, uint64_t cap_nodes, int32_t* n_trees,
                              uint64_t* n_nodes) {
  if (!ctx) return CAMUS_B200_ERR_ARG;
  if (!ctx->subs.empty()) return camus_b200_get_gene_trees(ctx->subs[0], node_off, parent, taxon, cap_nodes, n_trees, n_nodes);
  const uint64_t nn = ctx->h_parent.size();
  if (n_trees) *n_trees = (int32_t)ctx->h_node_off.size() - 1;
  if (n_nodes) *n_nodes = nn;
  if (node_off) std::copy(ctx->h_node_off.begin(), ctx->h_node_off.end(), node_off);
  if ((parent || taxon) && cap_nodes < nn) return ctx->fail(CAMUS_B200_ERR_ARG, "get_gene_trees: cap_nodes is smaller than the number of nodes");
  if (parent) std::copy(ctx->h_parent.begin(), ctx->h_parent.end(), parent);
  if (taxon) std::copy(ctx->h_taxon.begin(), ctx->h_taxon.end(), taxon);
  return 0;
}

int camus_b200_set_score_hint(camus_b200* ctx, int as_set) {
  if (!ctx) return CAMUS_B200_ERR_ARG;
  for (camus_b200* s : ctx->subs) camus_b200_set_score_hint(s, as_set);
  ctx->as_set_hint = as_set != 0;
  return 0;
}

int camus_b200_run_resident(camus_b200* ctx, int qmode, double threshold, int as_set, uint64_t* out,
                            uint64_t* n_survivors, uint64_t* sum_counts) {
  if (!ctx) return CAMUS_B200_ERR_ARG;
  if (!ctx->subs.empty()) {
    std::vector<uint64_t> ns(ctx->subs.size(), 0), sc(ctx->subs.size(), 0);
    ctx->reduced = false;
    int rc = fan_out(ctx, [&](camus_b200* s, int i) { return camus_b200_run_resident(s, qmode, threshold, as_set, nullptr, &ns[i], &sc[i]); });
    if (rc) return rc;
    uint64_t a = 0, b = 0;
    for (size_t i = 0; i < ns.size(); ++i) {
      a += ns[i];
      b += sc[i];
    }
    if (n_survivors) *n_survivors = a;
    if (sum_counts) *sum_counts = b;
    if ((rc = reduce_partials(ctx))) return rc;
    if (!out) return 0;
    if ((rc = camus_b200_quartet_totals_finish(ctx->subs[0], out))) return ctx->fail(rc, ctx->subs[0]->err);
    return 0;
  }
  if (ctx->genes_dirty || ctx->n == 0)
    return ctx->fail(CAMUS_B200_ERR_ARG, "run_resident: the gene trees are not on the device yet (call camus_b200_count_filter once)");
  if (qmode < 0 || qmode > 2 || !(threshold >= 0 && threshold <= 1)) return ctx->fail(CAMUS_B200_ERR_ARG, "run_resident: bad filter options");
  CK(cudaSetDevice(ctx->device));
  std::memset(ctx->ms, 0, sizeof(ctx->ms));
  ctx->launches = 0;
  ctx->want_all_tiles = false;
  int rc;
  const int passes = wanted_passes(ctx);
  uint64_t ns = 0, sc = 0;
  if (passes > 1) {
    ctx->cell_state = camus_b200::NONE;
    ctx->z_valid = false;
    if ((rc = run_passes(ctx, passes, qmode, threshold, as_set != 0, false, &ns, &sc))) return rc;
    if (n_survivors) *n_survivors = ns;
    if (sum_counts) *sum_counts = sc;
    if (!out) return 0;
    return run_finish(ctx, out);
  }
  ctx->passes = 1;
  if ((rc = build_planes(ctx, false))) return rc;
  ctx->cell_state = camus_b200::NONE;
  ctx->z_valid = false;
  if (ctx->use_fused) {
    if ((rc = run_fused(ctx, qmode, threshold, as_set != 0, &ns, &sc))) return rc;
  } else {
    if ((rc = run_count(ctx))) return rc;
    if ((rc = run_filter(ctx, qmode, threshold, &ns, &sc))) return rc;
    if ((rc = run_score(ctx, as_set != 0))) return rc;
  }
  if (ctx->comm) {  // library-owned collective: global scalars and table on every rank
    if ((rc = comm_reduce_scalars(ctx, &ns, &sc))) return rc;
    if ((rc = comm_reduce_table(ctx))) return rc;
  }
  if (n_survivors) *n_survivors = ns;
  if (sum_counts) *sum_counts = sc;
  if (!out) return 0;  // caller-side reduction: all-reduce the partial table, then call ..._finish
  return run_finish(ctx, out);
}

int camus_b200_quartet_totals_partial(camus_b200* ctx, int as_set) {
  if (!ctx) return CAMUS_B200_ERR_ARG;
  if (!ctx->subs.empty()) {
    // The sum over the devices lives in subs[0]'s table. It is formed ONCE per scoring pass: a second
    // call with the same weight (or a finish after a caller-side all-reduce) must not add the peers again.
    for (camus_b200* s : ctx->subs)
      if (!(s->z_valid && s->z_as_set == (as_set != 0))) ctx->reduced = false;  // someone rescans: all partial again
    if (ctx->reduced) return 0;
    int rc = fan_out(ctx, [&](camus_b200* s, int) { return camus_b200_quartet_totals_partial(s, as_set); });
    return rc ? rc : reduce_partials(ctx);
  }
  CK(cudaSetDevice(ctx->device));
  if (int rc = run_score(ctx, as_set != 0)) return rc;
  return comm_reduce_table(ctx);
}

int camus_b200_partial_buffer(camus_b200* ctx, void** dev_ptr, uint64_t* n_int64) {
  if (!ctx || !dev_ptr || !n_int64) return CAMUS_B200_ERR_ARG;
  if (!ctx->subs.empty()) {
    int rc = camus_b200_partial_buffer(ctx->subs[0], dev_ptr, n_int64);
    return rc ? ctx->fail(rc, ctx->subs[0]->err) : 0;
  }
  if (!ctx->z_valid) return ctx->fail(CAMUS_B200_ERR_ARG, "partial_buffer: call quartet_totals_partial first");
  *dev_ptr = ctx->d_Z.p;
  *n_int64 = (uint64_t)ctx->n * ctx->n * 3;
  return 0;
}

int camus_b200_quartet_totals_finish(camus_b200* ctx, uint64_t* out) {
  if (!ctx || !out) return CAMUS_B200_ERR_ARG;
  if (!ctx->subs.empty()) {
    if (!ctx->reduced) {  // e.g. straight after a fused count_filter: the devices hold partial tables
      if (int rc = reduce_partials(ctx)) return rc;
    }
    int rc = camus_b200_quartet_totals_finish(ctx->subs[0], out);
    return rc ? ctx->fail(rc, ctx->subs[0]->err) : 0;
  }
  CK(cudaSetDevice(ctx->device));
  return run_finish(ctx, out);
}

int camus_b200_quartet_totals(camus_b200* ctx, int as_set, uint64_t* out) {
  if (!ctx || !out) return CAMUS_B200_ERR_ARG;
  if (!ctx->subs.empty()) {
    int rc = camus_b200_quartet_totals_partial(ctx, as_set);
    return rc ? rc : camus_b200_quartet_totals_finish(ctx, out);
  }
  CK(cudaSetDevice(ctx->device));
  int rc = run_score(ctx, as_set != 0);
  if (rc) return rc;
  if ((rc = comm_reduce_table(ctx))) return rc;
  return run_finish(ctx, out);
}

int camus_b200_edge_penalties(camus_b200* ctx, uint64_t* out) {
  if (!ctx || !out) return CAMUS_B200_ERR_ARG;
  if (ctx->n == 0) return ctx->fail(CAMUS_B200_ERR_ARG, "edge_penalties: set the constraint tree first");
  if (!ctx->subs.empty()) {
    int rc = camus_b200_edge_penalties(ctx->subs[0], out);
    return rc ? ctx->fail(rc, ctx->subs[0]->err) : 0;
  }
  CK(cudaSetDevice(ctx->device));
  const int n = ctx->n;
  CK(ctx->d_out.alloc((size_t)n * n));
  StageTimer t(ctx, CAMUS_B200_T_PENALTY);
  k_penalties<<<dim3((n + 127) / 128, n), 128, 0, ctx->stream>>>(ctx->d_out.p, ctx->ct);
  CK_LAUNCH();
  if (int rc = copy_out(ctx, out, (size_t)n * n)) return rc;
  if (t.stop()) return CAMUS_B200_ERR_CUDA;
  return 0;
}

int camus_b200_constraint_tables(camus_b200* ctx, int32_t* lca, int32_t* depth, uint64_t* leaves_below) {
  if (!ctx) return CAMUS_B200_ERR_ARG;
  if (ctx->n == 0) return ctx->fail(CAMUS_B200_ERR_ARG, "constraint_tables: set the constraint tree first");
  if (!ctx->subs.empty()) {
    int rc = camus_b200_constraint_tables(ctx->subs[0], lca, depth, leaves_below);
    return rc ? ctx->fail(rc, ctx->subs[0]->err) : 0;
  }
  const int n = ctx->n;
  for (int i = 0; i < n; ++i) {  // internal preorder id i <-> caller's id
    const int r = ctx->h_ref_of_int[i];
    if (depth) depth[r] = ctx->h_depth[i];
    if (leaves_below) leaves_below[r] = (uint64_t)ctx->h_nleaves[i];
  }
  if (!lca) return 0;
  CK(cudaSetDevice(ctx->device));
  DevBuf<int32_t> d_lca;
  CK(d_lca.alloc((size_t)n * n));
  k_node_lca<<<dim3((n + 127) / 128, n), 128, 0, ctx->stream>>>(d_lca.p, ctx->ct);
  CK_LAUNCH();
  CK(cudaMemcpyAsync(lca, d_lca.p, (size_t)n * n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int camus_b200_comm_unique_id(uint8_t id[CAMUS_B200_COMM_ID_BYTES]) {
  static_assert(sizeof(ncclUniqueId) == CAMUS_B200_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
  NcclApi& api = nccl_api();
  if (!id || !api.ok()) {
    g_create_err = "camus_b200_comm_unique_id: " + (id ? api.err : std::string("null argument"));
    return CAMUS_B200_ERR_ARG;
  }
  ncclUniqueId u;
  ncclResult_t r = api.GetUniqueId(&u);
  if (r != ncclSuccess) {
    g_create_err = std::string("ncclGetUniqueId: ") + api.GetErrorString(r);
    return CAMUS_B200_ERR_CUDA;
  }
  std::memcpy(id, &u, sizeof(u));
  return 0;
}

int camus_b200_comm_init(camus_b200* ctx, const uint8_t id[CAMUS_B200_COMM_ID_BYTES], int rank, int world) {
  if (!ctx || !id) return CAMUS_B200_ERR_ARG;
  if (!ctx->subs.empty()) return ctx->fail(CAMUS_B200_ERR_ARG, "comm_init: one GPU per process (create the handle with n_gpus = 1)");
  if (world < 1 || rank < 0 || rank >= world) return ctx->fail(CAMUS_B200_ERR_ARG, "comm_init: bad rank / world");
  NcclApi& api = nccl_api();
  if (!api.ok()) return ctx->fail(CAMUS_B200_ERR_CUDA, "comm_init: " + api.err);
  CK(cudaSetDevice(ctx->device));
  if (ctx->comm) {
    api.CommDestroy(ctx->comm);
    ctx->comm = nullptr;
  }
  ncclUniqueId u;
  std::memcpy(&u, id, sizeof(u));
  ncclResult_t r = api.CommInitRank(&ctx->comm, world, u, rank);
  if (r != ncclSuccess) {
    ctx->comm = nullptr;
    return ctx->fail(CAMUS_B200_ERR_CUDA, std::string("ncclCommInitRank: ") + api.GetErrorString(r));
  }
  return camus_b200_set_partition(ctx, rank, world);
}

int camus_b200_export_quartets(camus_b200* ctx, int raw, uint64_t* keys, uint32_t* counts, uint64_t cap, uint64_t* n_out) {
  if (!ctx || !n_out) return CAMUS_B200_ERR_ARG;
  if (!ctx->subs.empty()) {
    // every device exports its share; the shares are disjoint, the merged list is sorted by key
    std::vector<std::vector<uint64_t>> k(ctx->subs.size());
    std::vector<std::vector<uint32_t>> c(ctx->subs.size());
    std::vector<uint64_t> cnt(ctx->subs.size(), 0);
    const bool fetch = keys && counts && cap;
    int rc = fan_out(ctx, [&](camus_b200* s, int i) {
      int r = camus_b200_export_quartets(s, raw, nullptr, nullptr, 0, &cnt[i]);
      if (r || !fetch || !cnt[i]) return r;
      k[i].resize(cnt[i]);
      c[i].resize(cnt[i]);
      return camus_b200_export_quartets(s, raw, k[i].data(), c[i].data(), cnt[i], &cnt[i]);
    });
    if (rc) return rc;
    uint64_t total = 0;
    for (uint64_t x : cnt) total += x;
    *n_out = total;
    if (fetch) {
      std::vector<std::pair<uint64_t, uint32_t>> all;
      all.reserve(total);
      for (size_t i = 0; i < k.size(); ++i)
        for (size_t j = 0; j < k[i].size(); ++j) all.emplace_back(k[i][j], c[i][j]);
      std::sort(all.begin(), all.end());
      for (uint64_t i = 0; i < std::min<uint64_t>(total, cap); ++i) {
        keys[i] = all[i].first;
        counts[i] = all[i].second;
      }
    }
    return 0;
  }
  CK(cudaSetDevice(ctx->device));
  if (!ctx->counted) return ctx->fail(CAMUS_B200_ERR_ARG, "export_quartets: call camus_b200_count_filter first");
  int rc;
  if (raw) {
    if (ctx->cell_state != camus_b200::RAW && (rc = run_count(ctx))) return rc;
  } else if ((rc = ensure_filtered(ctx))) {  // the table pipeline (K2 -> K3); Z, if any, stays valid
    return rc;
  }
  uint64_t mine = ctx->n_local;
  CK(ctx->d_nout.alloc(1));
  CK(ctx->d_keys.alloc(std::max<uint64_t>(cap, 1)));
  CK(ctx->d_counts.alloc(std::max<uint64_t>(cap, 1)));
  CK(cudaMemsetAsync(ctx->d_nout.p, 0, 8, ctx->stream));
  const uint64_t max_boxes = 1u << 26;
  for (uint64_t b0 = 0; b0 < mine; b0 += max_boxes) {
    uint64_t nbx = std::min(max_boxes, mine - b0);
    ExportParams prm{ctx->d_cells.p + b0 * kBoxCells, box_map(ctx, b0), ctx->d_keys.p, ctx->d_counts.p, cap, ctx->d_nout.p};
    k_export<<<(unsigned)(nbx * 16), 256, 0, ctx->stream>>>(prm, ctx->ct);
    CK_LAUNCH();
  }
  unsigned long long total = 0;
  CK(cudaMemcpyAsync(&total, ctx->d_nout.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *n_out = total;
  uint64_t got = std::min<uint64_t>(total, cap);
  if (got && keys && counts) {
    std::vector<unsigned long long> k(got);
    std::vector<uint32_t> c(got);
    CK(cudaMemcpy(k.data(), ctx->d_keys.p, got * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(c.data(), ctx->d_counts.p, got * 4, cudaMemcpyDeviceToHost));
    std::vector<uint32_t> order(got);
    for (uint64_t i = 0; i < got; ++i) order[i] = (uint32_t)i;
    std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return k[x] < k[y]; });
    for (uint64_t i = 0; i < got; ++i) {
      keys[i] = k[order[i]];
      counts[i] = c[order[i]];
    }
  }
  return 0;
}

int camus_b200_query_cells(camus_b200* ctx, int32_t n_query, const int32_t* quads, uint32_t* counts) {
  if (!ctx || n_query < 0 || (n_query && (!quads || !counts))) return CAMUS_B200_ERR_ARG;
  if (!ctx->subs.empty()) {  // every device holds all bit-planes
    int rc = camus_b200_query_cells(ctx->subs[0], n_query, quads, counts);
    return rc ? ctx->fail(rc, ctx->subs[0]->err) : 0;
  }
  if (ctx->genes_dirty || ctx->n == 0) return ctx->fail(CAMUS_B200_ERR_ARG, "query_cells: call camus_b200_count_filter first");
  for (int i = 0; i < 4 * n_query; ++i)
    if (quads[i] < 0 || quads[i] >= ctx->N) return ctx->fail(CAMUS_B200_ERR_ARG, "query_cells: tip index out of range");
  for (int q = 0; q < n_query; ++q)
    for (int i = 0; i < 4; ++i)
      for (int j = i + 1; j < 4; ++j)
        if (quads[4 * q + i] == quads[4 * q + j]) return ctx->fail(CAMUS_B200_ERR_ARG, "query_cells: repeated taxon in a 4-set");
  if (n_query == 0) return 0;
  CK(cudaSetDevice(ctx->device));
  if (ctx->tri_mask != 0xFu) {  // a class-scheme partition built only its own tiles
    ctx->want_all_tiles = true;
    if (int rc = build_planes(ctx, false)) return rc;
  }
  DevBuf<int32_t> dq;
  DevBuf<uint32_t> dc;
  CK(dq.alloc((size_t)4 * n_query));
  CK(dc.alloc((size_t)3 * n_query));
  CK(cudaMemcpyAsync(dq.p, quads, (size_t)16 * n_query, cudaMemcpyHostToDevice, ctx->stream));
  k_query_cells<<<(n_query + 127) / 128, 128, 0, ctx->stream>>>(n_query, dq.p, dc.p, ctx->d_tri.p, std::max(ctx->G32, 1),
                                                                ctx->d_rank_of_tip.p, ctx->d_tip_of_rank.p);
  CK_LAUNCH();
  CK(cudaMemcpyAsync(counts, dc.p, (size_t)12 * n_query, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

static int reticulation_counts_batch(camus_b200* ctx, int32_t n_tips, int32_t n_ret, const uint32_t* records,
                                   int32_t n_trees, const int64_t* node_off, const int32_t* parent,
                                   const int32_t* taxon, uint64_t* totals, uint64_t* supported) {
  if (!ctx) return CAMUS_B200_ERR_ARG;
  if (!ctx->subs.empty()) {
    int rc = camus_b200_reticulation_counts(ctx->subs[0], n_tips, n_ret, records, n_trees, node_off, parent, taxon, totals, supported);
    return rc ? ctx->fail(rc, ctx->subs[0]->err) : 0;
  }
  if (n_tips < 1 || n_tips > 32767 || n_ret < 1 || n_ret > kRetMax || n_trees < 0 || !records || !node_off || !totals || !supported ||
      (n_trees > 0 && (!parent || !taxon)))
    return ctx->fail(CAMUS_B200_ERR_ARG, "reticulation_counts: bad argument (1..32767 tips, 1..64 reticulations)");
  // the flat trees are validated like camus_b200_add_gene_trees does
  int max_nodes = 1;
  std::vector<int> stamp(n_tips, -1), depth_h;
  std::vector<char> has_child;
  for (int g = 0; g < n_trees; ++g) {
    const int64_t b = node_off[g], nn = node_off[g + 1] - b;
    if (nn <= 0 || nn > 24576)
      return ctx->fail(CAMUS_B200_ERR_INPUT, "reticulation_counts: tree " + std::to_string(g) + " has " + std::to_string(nn) + " nodes (1..24576 supported)");
    max_nodes = std::max<int>(max_nodes, (int)nn);
    has_child.assign(nn, 0);
    depth_h.assign(nn, 0);
    for (int64_t i = 0; i < nn; ++i) {
      const int p = parent[b + i];
      if ((i == 0) != (p < 0) || p >= i)
        return ctx->fail(CAMUS_B200_ERR_INPUT, "reticulation_counts: tree " + std::to_string(g) + " is not in preorder");
      if (p >= 0) {
        has_child[p] = 1;
        depth_h[i] = depth_h[p] + 1;
      }
      const int t = taxon[b + i];
      if (t >= n_tips) return ctx->fail(CAMUS_B200_ERR_INPUT, "reticulation_counts: taxon index out of range");
      if (t >= 0) {
        if (stamp[t] == g) return ctx->fail(CAMUS_B200_ERR_INPUT, "reticulation_counts: tree " + std::to_string(g) + " contains a taxon twice");
        stamp[t] = g;
      }
    }
    for (int64_t i = 0; i < nn; ++i)
      if ((taxon[b + i] >= 0) == (has_child[i] != 0))
        return ctx->fail(CAMUS_B200_ERR_INPUT, "reticulation_counts: tree " + std::to_string(g) + ": taxon set on an internal node or missing on a leaf");
  }
  const size_t n_out = (size_t)n_trees * n_ret;
  if (n_trees == 0) return 0;
  CK(cudaSetDevice(ctx->device));
  const uint64_t n_sets = choose4((uint64_t)n_tips);
  if (n_sets == 0) {
    std::memset(totals, 0, n_out * 8);
    std::memset(supported, 0, n_out * 8);
    return 0;
  }
  const size_t smem_rows = (size_t)n_tips * 4;  // two u16 depth rows per block
  if (smem_rows > 200 * 1024) return ctx->fail(CAMUS_B200_ERR_NOMEM, "reticulation_counts: network too large for the row-staged kernel");
  const size_t total_nodes = (size_t)(node_off[n_trees] - node_off[0]);
  DevBuf<int64_t> d_off;
  DevBuf<int32_t> d_par, d_tax;
  DevBuf<uint32_t> d_rec;
  DevBuf<uint16_t> d_D;
  DevBuf<uint8_t> d_rb;
  DevBuf<unsigned long long> d_tot, d_sup, d_uw;
  // per tip: bit r = the tip is under w of reticulation r (bit 0 of the record's flag word)
  std::vector<unsigned long long> under_w(n_tips, 0);
  for (int r = 0; r < n_ret; ++r)
    for (int t = 0; t < n_tips; ++t)
      if (records[((size_t)r * n_tips + t) * 4 + 3] & 1u) under_w[t] |= 1ull << r;
  // depth tables of a batch of trees: <= 1 GiB, <= 65535 trees (grid.y)
  const size_t per_tree = (size_t)n_tips * n_tips;
  const int batch = (int)std::max<size_t>(1, std::min<size_t>({(size_t)n_trees, (size_t)65535, ((size_t)1 << 29) / per_tree}));
  std::vector<int64_t> off0(node_off, node_off + n_trees + 1);
  for (auto& o : off0) o -= node_off[0];
  CK(d_off.alloc(n_trees + 1));
  CK(d_par.alloc(total_nodes));
  CK(d_tax.alloc(total_nodes));
  CK(d_rec.alloc((size_t)n_ret * n_tips * 4));
  CK(d_D.alloc(per_tree * batch));
  CK(d_rb.alloc(n_trees));
  CK(d_tot.alloc(n_out));
  CK(d_sup.alloc(n_out));
  CK(d_uw.alloc(n_tips));
  CK(cudaMemcpyAsync(d_uw.p, under_w.data(), (size_t)n_tips * 8, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_off.p, off0.data(), (n_trees + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_par.p, parent + node_off[0], total_nodes * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_tax.p, taxon + node_off[0], total_nodes * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(d_rec.p, records, (size_t)n_ret * n_tips * 16, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync(d_tot.p, 0, n_out * 8, ctx->stream));
  CK(cudaMemsetAsync(d_sup.p, 0, n_out * 8, ctx->stream));
  const size_t smem = (size_t)max_nodes * 8;
  if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_ret_depths, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (smem_rows > 48 * 1024) CK(cudaFuncSetAttribute(k_ret_score, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows));
  for (int g0 = 0; g0 < n_trees; g0 += batch) {
    const int nb = std::min(batch, n_trees - g0);
    CK(cudaMemsetAsync(d_D.p, 0, per_tree * nb * 2, ctx->stream));
    k_ret_depths<<<nb, kRetThreads, smem, ctx->stream>>>(d_off.p + g0, d_par.p, d_tax.p, n_tips, max_nodes, d_D.p, d_rb.p + g0);
    CK_LAUNCH();
    k_ret_score<<<dim3((unsigned)choose2((uint64_t)n_tips), nb), kRetThreads, smem_rows, ctx->stream>>>(
        n_tips, n_ret, reinterpret_cast<const uint4*>(d_rec.p), d_uw.p, d_D.p, d_rb.p + g0,
        d_tot.p + (size_t)g0 * n_ret, d_sup.p + (size_t)g0 * n_ret);
    CK_LAUNCH();
  }
  CK(cudaMemcpyAsync(totals, d_tot.p, n_out * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(supported, d_sup.p, n_out * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// More than 64 reticulations (score.ReticulationScore has no limit): batches of 64 on this side of the ABI;
// the per-reticulation records and outputs are independent.
int camus_b200_reticulation_counts(camus_b200* ctx, int32_t n_tips, int32_t n_ret, const uint32_t* records, int32_t n_trees,
                                   const int64_t* node_off, const int32_t* parent, const int32_t* taxon, uint64_t* totals,
                                   uint64_t* supported) {
  if (!ctx) return CAMUS_B200_ERR_ARG;
  if (n_ret <= kRetMax || !records || !totals || !supported || n_trees < 0 || n_tips < 1)
    return reticulation_counts_batch(ctx, n_tips, n_ret, records, n_trees, node_off, parent, taxon, totals, supported);
  std::vector<uint64_t> t((size_t)n_trees * kRetMax), s((size_t)n_trees * kRetMax);
  for (int r0 = 0; r0 < n_ret; r0 += kRetMax) {
    const int rc_n = std::min<int>(kRetMax, n_ret - r0);
    int rc = reticulation_counts_batch(ctx, n_tips, rc_n, records + (size_t)r0 * n_tips * 4, n_trees, node_off, parent, taxon, t.data(), s.data());
    if (rc) return rc;
    for (int g = 0; g < n_trees; ++g)
      for (int r = 0; r < rc_n; ++r) {
        totals[(size_t)g * n_ret + r0 + r] = t[(size_t)g * rc_n + r];
        supported[(size_t)g * n_ret + r0 + r] = s[(size_t)g * rc_n + r];
      }
  }
  return 0;
}

int camus_b200_stage_ms(const camus_b200* ctx, float* ms, uint64_t* n_launches) {
  if (!ctx) return CAMUS_B200_ERR_ARG;
  if (!ctx->subs.empty()) {  // slowest device per stage, launches of all
    float m[CAMUS_B200_T_N] = {0};
    uint64_t l = 0;
    for (const camus_b200* s : ctx->subs) {
      for (int i = 0; i < CAMUS_B200_T_N; ++i) m[i] = std::max(m[i], s->ms[i]);
      l += s->launches;
    }
    if (ms) std::memcpy(ms, m, sizeof(m));
    if (n_launches) *n_launches = l;
    return 0;
  }
  if (ms) std::memcpy(ms, ctx->ms, sizeof(ctx->ms));
  if (n_launches) *n_launches = ctx->launches;
  return 0;
}

int camus_b200_workload(const camus_b200* ctx, uint64_t* resolutions, uint64_t* n_cells, uint64_t* n_boxes) {
  if (!ctx) return CAMUS_B200_ERR_ARG;
  if (!ctx->subs.empty()) return camus_b200_workload(ctx->subs[0], resolutions, n_cells, n_boxes);
  if (resolutions) *resolutions = ctx->resolutions;
  if (n_cells) *n_cells = choose4((uint64_t)ctx->N);
  if (n_boxes) *n_boxes = ctx->n ? num_boxes(ctx->M) : 0;
  return 0;
}

}  // extern "C"
