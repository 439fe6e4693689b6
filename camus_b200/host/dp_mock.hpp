// CPU stand-in for the camus_b200_dp_* entry points (include/camus_b200.h, "DP offload"), used ONLY by the
// `-m "not gpu"` tests: it lets them drive the host side of the offload (camus_b200/host/dp.hpp: column
// uploads, the NaN rule, one materialise per step, thresholds) without a GPU. Like the kernels of
// camus_b200/csrc/dp_across.cu it evaluates every (u, w) pair literally and then reduces with the CLOSED FORM
// of the reference's sequential acceptance (first maximum over w; across the u maximum score, smallest cycle
// length, last u; NaN candidates never win, the first u's NaN candidate is reported separately) -- so the tests
// check that closed form against the literal loop of main_dp.go:199-214, on the CPU. Not linked into any
// product path (the product's accelerator is libcamus_b200.so, which has no CPU fallback).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "dp.hpp"

namespace camus {
namespace dp_mock {

struct Handle {
  bool f64 = false;
  int n = 0;
  std::vector<int32_t> end, depth;
  std::vector<uint64_t> edge;                // n * n raw scores
  std::vector<std::vector<uint64_t>> dp;     // DP[x]
  std::vector<std::vector<uint64_t>> cs;     // cs[x][k]
  std::vector<std::vector<uint64_t>> cv;     // conv[x][k] (max-plus form)
  std::vector<int32_t> item_u, item_sub;
  int v = -1;
  uint64_t calls = 0;
  std::string err;
};

template <class S>
inline S as(uint64_t r) {
  S v;
  std::memcpy(&v, &r, 8);
  return v;
}
template <class S>
inline uint64_t raw(S v) {
  uint64_t r;
  std::memcpy(&r, &v, 8);
  return r;
}
template <class S>
inline bool nan_of(S) { return false; }
template <>
inline bool nan_of<double>(double v) { return v != v; }

struct Cand {
  bool ok = false;
  uint64_t score = 0;
  int u = 0, w = 0, len = 0, idx[4] = {0, 0, 0, 0};
};

template <class S>
inline void pair_eval(const Handle& h, int u, int w, int k, Cand& c) {
  const auto &A = h.cs[w], &B = h.cs[u], &C = h.dp[w], &D = h.dp[u];
  const int Lc = (int)C.size(), Ld = (int)D.size();
  const int m2 = std::min(k, Lc + Ld - 2) + 1;
  bool found = false;
  S top{};
  int ti[4] = {0, 0, 0, 0};
  for (int i = k - (m2 - 1); i <= k; ++i) {
    S c1 = as<S>(A[0]) + as<S>(B[i]);
    int a1 = 0;
    for (int j = 1; j <= i; ++j) {
      const S cur = as<S>(A[j]) + as<S>(B[i - j]);
      if (cur > c1) {
        c1 = cur;
        a1 = j;
      }
    }
    const int i2 = k - i, lo2 = std::max(0, i2 - (Ld - 1)), hi2 = std::min(i2, Lc - 1);
    S c2 = as<S>(C[lo2]) + as<S>(D[i2 - lo2]);
    int cc = lo2;
    for (int j = lo2 + 1; j <= hi2; ++j) {
      const S cur = as<S>(C[j]) + as<S>(D[i2 - j]);
      if (cur > c2) {
        c2 = cur;
        cc = j;
      }
    }
    const S cur = c1 + c2;
    if (!found || cur > top) {
      top = cur;
      ti[0] = a1;
      ti[1] = i - a1;
      ti[2] = cc;
      ti[3] = i2 - cc;
      found = true;
    }
  }
  S score = as<S>(h.edge[(size_t)u * h.n + w]);
  score = score + as<S>(A[ti[0]]);
  score = score + as<S>(B[ti[1]]);
  score = score + as<S>(C[ti[2]]);
  score = score + as<S>(D[ti[3]]);
  c.ok = true;
  c.score = raw(score);
  c.u = u;
  c.w = w;
  for (int q = 0; q < 4; ++q) c.idx[q] = ti[q];
}

template <class S>
inline void across(Handle& h, int k, DpAccelBest& out) {
  std::memset(&out, 0, sizeof out);
  Cand best;
  for (size_t it = 0; it < h.item_u.size(); ++it) {
    const int u = h.item_u[it], sub = h.item_sub[it];
    // per item, in an ORDER-FREE form (the kernels visit w sorted by DP length): first maximum = larger score, then smaller w;
    // NaN scores never enter, except the one of the item's first w, which wins outright
    Cand ib;
    bool stuck = false;
    for (int w = h.end[sub] - 1; w >= sub && !stuck; --w) {  // deliberately NOT preorder
      Cand c;
      pair_eval<S>(h, u, w, k, c);
      if (nan_of(as<S>(c.score))) {
        out.nan_seen = 1;
        if (w != sub) continue;
        ib = c;
        stuck = true;
        break;
      }
      if (!ib.ok || as<S>(c.score) > as<S>(ib.score) || (as<S>(c.score) == as<S>(ib.score) && c.w < ib.w)) ib = c;
    }
    if (!ib.ok) continue;
    ib.len = (h.depth[u] - h.depth[h.v]) + (h.depth[ib.w] - h.depth[h.v]) + 1;
    if (nan_of(as<S>(ib.score))) {
      if (it == 0) {
        out.first_is_nan = 1;
        out.first_u = ib.u;
        out.first_w = ib.w;
        out.first_score_bits = ib.score;
        for (int q = 0; q < 4; ++q) out.first_split[q] = ib.idx[q];
      }
      continue;
    }
    if (!best.ok || as<S>(ib.score) > as<S>(best.score) || (as<S>(ib.score) == as<S>(best.score) && ib.len <= best.len)) best = ib;
  }
  out.found = best.ok;
  out.score_bits = best.score;
  out.u = best.u;
  out.w = best.w;
  out.cycle_length = best.len;
  for (int q = 0; q < 4; ++q) out.split[q] = best.idx[q];
}

inline int create(void** out, int, int is_f64, int32_t n, const int32_t* end, const int32_t* depth, const void* edge) {
  auto* h = new Handle();
  h->f64 = is_f64 != 0;
  h->n = n;
  h->end.assign(end, end + n);
  h->depth.assign(depth, depth + n);
  h->edge.assign(static_cast<const uint64_t*>(edge), static_cast<const uint64_t*>(edge) + (size_t)n * n);
  h->dp.assign(n, std::vector<uint64_t>(1, 0));
  h->cs.assign(n, {});
  h->cv.assign(n, {});
  *out = h;
  return 0;
}
inline void destroy(void* p) { delete static_cast<Handle*>(p); }
inline const char* last_error(const void* p) { return p ? static_cast<const Handle*>(p)->err.c_str() : "mock"; }
inline int set_node(void* p, int32_t node, int32_t len, const void* scores) {
  auto* h = static_cast<Handle*>(p);
  h->dp[node].assign(static_cast<const uint64_t*>(scores), static_cast<const uint64_t*>(scores) + len);
  return 0;
}
inline int set_vertex(void* p, int32_t v, int32_t n_items, const int32_t* iu, const int32_t* is) {
  auto* h = static_cast<Handle*>(p);
  h->v = v;
  h->item_u.assign(iu, iu + n_items);
  h->item_sub.assign(is, is + n_items);
  return 0;
}
inline int set_path_columns(void* p, int32_t k0, int32_t k1, int32_t first, int32_t count, const void* cols) {
  auto* h = static_cast<Handle*>(p);
  const auto* c = static_cast<const uint64_t*>(cols);
  for (int j = k0; j < k1; ++j)
    for (int x = 0; x < count; ++x) {
      auto& row = h->cs[first + x];
      if ((int)row.size() <= j) row.resize(j + 1, 0);
      row[j] = c[(size_t)(j - k0) * count + x];
    }
  return 0;
}
inline int set_conv_columns(void* p, int32_t k0, int32_t k1, int32_t first, int32_t count, const void* cols) {
  auto* h = static_cast<Handle*>(p);
  const auto* c = static_cast<const uint64_t*>(cols);
  for (int j = k0; j < k1; ++j)
    for (int x = 0; x < count; ++x) {
      auto& row = h->cv[first + x];
      if ((int)row.size() <= j) row.resize(j + 1, 0);
      row[j] = c[(size_t)(j - k0) * count + x];
    }
  return 0;
}
// max-plus form: value only, split = -1; the same order-free reduction as across()
inline int across_maxplus_entry(void* p, int32_t k, void* best) {
  auto* h = static_cast<Handle*>(p);
  ++h->calls;
  auto& out = *static_cast<DpAccelBest*>(best);
  std::memset(&out, 0, sizeof out);
  Cand top;
  for (size_t it = 0; it < h->item_u.size(); ++it) {
    const int u = h->item_u[it], sub = h->item_sub[it];
    Cand ib;
    for (int w = h->end[sub] - 1; w >= sub; --w) {
      uint64_t val = 0;
      for (int j = 0; j <= k; ++j) val = std::max(val, h->cv[w][j] + h->cv[u][k - j]);
      const uint64_t score = h->edge[(size_t)u * h->n + w] + val;
      if (!ib.ok || score > ib.score || (score == ib.score && w < ib.w)) {
        ib.ok = true;
        ib.score = score;
        ib.u = u;
        ib.w = w;
      }
    }
    if (!ib.ok) continue;
    ib.len = (h->depth[u] - h->depth[h->v]) + (h->depth[ib.w] - h->depth[h->v]) + 1;
    if (!top.ok || ib.score > top.score || (ib.score == top.score && ib.len <= top.len)) top = ib;
  }
  out.found = top.ok;
  out.score_bits = top.score;
  out.u = top.u;
  out.w = top.w;
  out.cycle_length = top.len;
  for (int q = 0; q < 4; ++q) out.split[q] = -1;
  return 0;
}
inline int across_entry(void* p, int32_t prev_k, void* best) {
  auto* h = static_cast<Handle*>(p);
  ++h->calls;
  if (h->f64) across<double>(*h, prev_k, *static_cast<DpAccelBest*>(best));
  else across<uint64_t>(*h, prev_k, *static_cast<DpAccelBest*>(best));
  return 0;
}
inline uint64_t launches(const void* p) { return static_cast<const Handle*>(p)->calls; }

inline const DpAccelApi* api() {
  static const DpAccelApi a{create, destroy, last_error, set_node, set_vertex, set_path_columns, across_entry, launches,
                            set_conv_columns, across_maxplus_entry};
  return &a;
}

}  // namespace dp_mock
}  // namespace camus
