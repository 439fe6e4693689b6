// ORACLE -- TEST INFRASTRUCTURE ONLY (included by camus_oracle.cpp; see the header there).
//
// Dense mode: the same results as the oracle's literal / fast modes at sizes where a
// map[Quartet]uint32 (internal/prep/preprocess.go:64-79) cannot be held: config 2 in full
// (200 x 1000: 6.5e10 resolutions), 1000 taxa x 32..64 trees (4.1e10 cells), config 5.
// No hash map, no cell table: 4-taxon sets are walked in blocks (a, b, c) x 16 consecutive d,
// the three pairings are resolved per gene tree with the four-point test on LCA depths
// (s1 = D(a,b)+D(c,d), s2 = D(a,c)+D(b,d), s3 = D(a,d)+D(b,c); a pairing
// is displayed iff its sum is the strict maximum), vectorised over d; each cell then goes through
//   prep.filterQuartets / Threshold.Keep        internal/prep/quartetfilter.go:67-96
//   constraint-quartet removal                  internal/prep/preprocess.go:51-57
//   TotalNumQuartets / TotalNumUniqueQuartets   internal/graphs/treedata.go:297-307
// and the survivors are scattered into the edge table with scatter_quartet (the form that is
// property-tested against the literal per-(u,w) QuartetScore loop, quartettotals.go:78-157).
// Taxa are constraint tip indices throughout (a < b < c < d by tip index), so the three pairings
// ab|cd, ac|bd, ad|bc are the reference's topologies 0b1100, 0b1010, 0b0110 (quartet.go:23-25).
//
// dense_literal_at() is the literal per-(u,w) loop for ONE LCA vertex v at any size: it builds
// td.Quartets(v) (treedata.go:197-222: quartets with >= 3 taxa under v) from dense counts and
// runs quartets_total (quartettotals.go:78-93) with the line-by-line quartet_score on it.
#pragma once

typedef int16_t v16 __attribute__((vector_size(32)));
typedef int16_t v16u __attribute__((vector_size(32), aligned(2)));

struct Dense {
  int N = 0, NB = 0, G = 0;  // taxa, blocks of 16 taxa, gene trees
  bool any_absent = false;
  std::vector<int16_t> D;   // [x][yb][g][16]: depth of LCA(x, 16*yb + lane) in tree g (0 if either is absent)
  std::vector<int16_t> P;   // [yb][g][16]: -1 if taxon 16*yb + lane is in tree g, else 0
  std::vector<int16_t> CT;  // [x][y], stride 16*NB: depth of LCA(x, y) in the constraint tree
  std::vector<int32_t> LT;  // [x][y], same stride: node id of that LCA
  size_t at(int x, int yb, int g) const { return (((size_t)x * NB + yb) * G + g) * 16; }
  const int16_t* pair(int x, int y) const { return &D[at(x, y >> 4, 0) + (y & 15)]; }  // stride 16 over g
  const int16_t* present(int x) const { return &P[((size_t)(x >> 4) * G) * 16 + (x & 15)]; }  // stride 16 over g
};

// counts of the pairings ab|cd, ac|bd, ad|bc for d = the 16 taxa of block db, over all trees.
// pab/pac/pbc: per-tree scalars, stride 16; vad/vbd/vcd: per-tree vectors; pm: per-tree scalar
// presence of a, b and c (nullptr: every tree holds every taxon), vp: per-tree presence of d.
__attribute__((target_clones("avx2", "default"))) void dense_count_run(const int16_t* pab, const int16_t* pac,
                                                                       const int16_t* pbc, const int16_t* vad,
                                                                       const int16_t* vbd, const int16_t* vcd,
                                                                       const int16_t* pm, const int16_t* vp, int G,
                                                                       int16_t out[3][16]) {
  v16 a0 = {0}, a1 = {0}, a2 = {0};
  if (!pm) {
    for (int g = 0; g < G; g++) {
      const v16 s1 = *(const v16u*)(vcd + 16 * g) + pab[16 * g];
      const v16 s2 = *(const v16u*)(vbd + 16 * g) + pac[16 * g];
      const v16 s3 = *(const v16u*)(vad + 16 * g) + pbc[16 * g];
      a0 += (s1 > s2) & (s1 > s3);
      a1 += (s2 > s1) & (s2 > s3);
      a2 += (s3 > s1) & (s3 > s2);
    }
  } else {
    for (int g = 0; g < G; g++) {
      if (!pm[g]) continue;
      const v16 ok = *(const v16u*)(vp + 16 * g);
      const v16 s1 = *(const v16u*)(vcd + 16 * g) + pab[16 * g];
      const v16 s2 = *(const v16u*)(vbd + 16 * g) + pac[16 * g];
      const v16 s3 = *(const v16u*)(vad + 16 * g) + pbc[16 * g];
      a0 += (s1 > s2) & (s1 > s3) & ok;
      a1 += (s2 > s1) & (s2 > s3) & ok;
      a2 += (s3 > s1) & (s3 > s2) & ok;
    }
  }
  for (int l = 0; l < 16; l++) {
    out[0][l] = (int16_t)-a0[l];
    out[1][l] = (int16_t)-a1[l];
    out[2][l] = (int16_t)-a2[l];
  }
}

// LCA-depth tables of every gene tree (running minimum over the DFS leaf order, as
// enumerate_quartets_fast does) and of the constraint tree.
std::shared_ptr<Dense> build_dense(const Ctx& c, int nthreads) {
  auto dn = std::make_shared<Dense>();
  Dense& d = *dn;
  d.N = c.n_taxa;
  d.NB = (c.n_taxa + 15) / 16;
  d.G = (int)c.genes.size();
  if (d.G > 32767) return nullptr;
  d.D.assign((size_t)d.N * d.NB * std::max(d.G, 1) * 16, 0);
  d.P.assign((size_t)d.NB * std::max(d.G, 1) * 16, 0);
  std::atomic<bool> absent{false}, too_deep{false};
  parallel_for(d.G, nthreads, [&](int g, int) {
    const FlatTree& t = c.genes[g];
    std::vector<int> depth(t.n, 0), order;
    std::vector<int> st{t.root};
    while (!st.empty()) {
      int x = st.back();
      st.pop_back();
      order.push_back(x);
      for (auto it = t.children[x].rbegin(); it != t.children[x].rend(); ++it) {
        depth[*it] = depth[x] + 1;
        st.push_back(*it);
      }
    }
    std::vector<int> leaves, pos(t.n, -1);
    for (int i = 0; i < (int)order.size(); i++) {
      pos[order[i]] = i;
      if (t.children[order[i]].empty()) leaves.push_back(order[i]);
    }
    const int m = (int)leaves.size();
    if (m != d.N) absent = true;
    for (int x : leaves) d.P[((size_t)(t.taxon[x] >> 4) * d.G + g) * 16 + (t.taxon[x] & 15)] = -1;
    std::vector<int> gap(std::max(m - 1, 0));
    for (int k = 0; k + 1 < m; k++) {
      gap[k] = depth[t.parent[order[pos[leaves[k]] + 1]]];
      if (gap[k] > 8000) too_deep = true;  // sums of two depths must fit int16
    }
    for (int p = 0; p < m; p++) {
      int mn = 1 << 30;
      const int tp = t.taxon[leaves[p]];
      for (int q = p + 1; q < m; q++) {
        mn = std::min(mn, gap[q - 1]);
        const int tq = t.taxon[leaves[q]];
        d.D[d.at(tp, tq >> 4, g) + (tq & 15)] = (int16_t)mn;
        d.D[d.at(tq, tp >> 4, g) + (tp & 15)] = (int16_t)mn;
      }
    }
  });
  if (too_deep) return nullptr;
  d.any_absent = absent;
  const int stride = d.NB * 16;
  d.CT.assign((size_t)d.N * stride, 0);
  d.LT.assign((size_t)d.N * stride, 0);
  for (int x = 0; x < d.N; x++)
    for (int y = 0; y < d.N; y++) {
      const int l = c.lca[c.tip_to_node[x]][c.tip_to_node[y]];
      d.LT[(size_t)x * stride + y] = l;
      if (x != y) d.CT[(size_t)x * stride + y] = (int16_t)c.depth[l];
    }
  return dn;
}

// prep.filterQuartets (quartetfilter.go:76-96) on the three topology counts of ONE 4-taxon set,
// c[i] = count of topology i in the reference's AllQuartets() order (0 = key absent from the
// map). The reference visits every key of the map once and re-reads the counts after earlier
// deletions; visiting the surviving keys of this 4-set in order 0, 1, 2 is one such schedule
// (the result does not depend on the schedule for thresholds below 1: a second visit finds
// (0, mid, hi) or (0, 0, hi) and deletes nothing; tests/test_oracle_dense.py checks all six).
inline void filter_cell_literal(uint32_t c[3], int mode, double thresh, const int order[3]) {
  for (int k = 0; k < 3; k++) {
    if (c[order[k]] == 0) continue;  // not in the map (any more)
    int idx[3] = {0, 1, 2};          // slices.SortFunc on 3 elements: insertion sort, stable
    for (int i = 1; i < 3; i++)
      for (int j = i; j > 0 && c[idx[j]] < c[idx[j - 1]]; j--) std::swap(idx[j], idx[j - 1]);
    if (!threshold_keep(thresh, c[0], c[1], c[2])) {
      c[idx[0]] = 0;
      c[idx[1]] = 0;
      continue;
    }
    if (mode == 2) c[idx[0]] = 0;
  }
}

inline uint64_t mix_key(uint64_t key, uint32_t count) {  // order-independent digests of (key, count) sets
  uint64_t z = key + 0x9E3779B97F4A7C15ull * (uint64_t)(count + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

struct DenseStats {
  uint64_t n_survivors = 0, sum_counts = 0, n_raw_keys = 0, sum_raw = 0, hash_raw = 0, hash_surv = 0, resolutions = 0, cells = 0;
};

// topology (0, 1, 2) the constraint tree displays on a < b < c < d (four-point on its own depths;
// the constraint is binary, so the maximum is strict)
inline int constraint_topology(const Dense& d, int a, int b, int c, int e) {
  const int stride = d.NB * 16;
  const int16_t* T = d.CT.data();
  const int s1 = T[(size_t)a * stride + b] + T[(size_t)c * stride + e];
  const int s2 = T[(size_t)a * stride + c] + T[(size_t)b * stride + e];
  const int s3 = T[(size_t)a * stride + e] + T[(size_t)b * stride + c];
  return s1 > s2 && s1 > s3 ? 0 : (s2 > s3 ? 1 : 2);
}

// The whole hot path, dense: counts -> filter -> removal -> scalars (+ digests) -> edge table.
// out == nullptr skips the scatter (counting / filtering only).
int dense_run(Ctx& c, int qmode, double thresh, bool as_set, int nthreads, uint64_t* out, DenseStats* stats, bool digests) {
  if (!c.dense) c.dense = build_dense(c, nthreads);
  if (!c.dense) {
    c.err = "dense mode: more than 32767 gene trees or a tree deeper than 8000";
    return 1;
  }
  if (out && !ids_are_preorder(c)) {
    c.err = "dense totals need preorder node ids";
    return 1;
  }
  const Dense& d = *c.dense;
  const int N = d.N, G = d.G, n = c.n;
  nthreads = std::max(1, nthreads);
  static const uint8_t kTopos[3] = {kTopo1, kTopo2, kTopo3};
  static const int kOrder[3] = {0, 1, 2};
  ScatterAux aux(c);
  std::vector<std::vector<int64_t>> diff(nthreads);
  std::vector<DenseStats> st(nthreads);
  // Tasks = pairs of 4-taxon tiles (A <= B). Inside a task the block of 16 candidate d is the
  // outer loop, so the 8 per-tree blocks of the tile taxa stay in L2 while the blocks of c stream
  // through once and serve up to 16 (a, b) pairs.
  constexpr int kTile = 4;
  const int NT = (N + kTile - 1) / kTile;
  std::vector<std::pair<int, int>> tasks;
  for (int tb = 0; tb < NT; tb++)
    for (int ta = 0; ta <= tb; ta++) tasks.emplace_back(ta, tb);  // small tb first: most (c, d) above them
  parallel_for((int)tasks.size(), nthreads, [&](int ti, int th) {
    const int a_lo = tasks[ti].first * kTile, b_lo = tasks[ti].second * kTile;
    const int a_hi = std::min(a_lo + kTile, N), b_hi = std::min(b_lo + kTile, N);
    DenseStats& s = st[th];
    if (out && diff[th].empty()) diff[th].assign((size_t)n * (n + 1), 0);
    std::vector<int16_t> pm;
    for (int db = (b_lo + 2) >> 4; db < d.NB; db++) {
      const int c_hi = std::min(N - 1, db * 16 + 15);  // c < d <= 16 * db + 15
      for (int cc = b_lo + 1; cc < c_hi; cc++)
        for (int a = a_lo; a < a_hi; a++)
          for (int b = std::max(b_lo, a + 1); b < b_hi && b < cc; b++) {
            const int16_t* pmask = nullptr;
            if (d.any_absent) {
              pm.resize(G);
              const int16_t *pa = d.present(a), *pb = d.present(b), *pc = d.present(cc);
              for (int g = 0; g < G; g++) pm[g] = pa[16 * g] & pb[16 * g] & pc[16 * g];
              pmask = pm.data();
            }
            int16_t cnt[3][16];
            dense_count_run(d.pair(a, b), d.pair(a, cc), d.pair(b, cc), &d.D[d.at(a, db, 0)], &d.D[d.at(b, db, 0)],
                            &d.D[d.at(cc, db, 0)], pmask, &d.P[((size_t)db * G) * 16], G, cnt);
            // per cell: raw statistics, then the cells where something besides the constraint's own
            // topology was counted (only those can leave survivors) go through filter + removal
            const int stride = d.NB * 16;
            const int16_t *cta = &d.CT[(size_t)a * stride + db * 16], *ctb = &d.CT[(size_t)b * stride + db * 16],
                          *ctc = &d.CT[(size_t)cc * stride + db * 16];
            const int ctab = d.CT[(size_t)a * stride + b], ctac = d.CT[(size_t)a * stride + cc], ctbc = d.CT[(size_t)b * stride + cc];
            const int l_lo = std::max(0, cc + 1 - db * 16), l_hi = std::min(16, N - db * 16);
            s.cells += (uint64_t)std::max(0, l_hi - l_lo);
            for (int l = l_lo; l < l_hi; l++) {
              uint32_t k[3] = {(uint32_t)cnt[0][l], (uint32_t)cnt[1][l], (uint32_t)cnt[2][l]};
              const uint32_t tot = k[0] + k[1] + k[2];
              if (!tot) continue;
              const int nz = (k[0] != 0) + (k[1] != 0) + (k[2] != 0);
              s.n_raw_keys += nz;
              s.sum_raw += tot;
              const int e = db * 16 + l;
              const int16_t ids[4] = {(int16_t)a, (int16_t)b, (int16_t)cc, (int16_t)e};
              if (digests)
                for (int t = 0; t < 3; t++)
                  if (k[t]) s.hash_raw += mix_key(make_quartet(ids, kTopos[t]), k[t]);
              // topology the constraint displays (four-point on its own depths; binary => strict maximum)
              const int s1 = ctab + ctc[l], s2 = ctac + ctb[l], s3 = ctbc + cta[l];
              const int ct = s1 > s2 && s1 > s3 ? 0 : (s2 > s3 ? 1 : 2);
              if (k[ct] == tot) continue;
              // a single key (0, 0, hi) fails Keep (0 < 0) and loses its two absent siblings: unchanged
              if (qmode != 0 && nz > 1) filter_cell_literal(k, qmode, thresh, kOrder);
              k[ct] = 0;
              if (!(k[0] | k[1] | k[2])) continue;
              const int leaf[4] = {c.tip_to_node[a], c.tip_to_node[b], c.tip_to_node[cc], c.tip_to_node[e]};
              int Lm[4][4];
              if (out) {
                const int32_t* LT = d.LT.data();
                Lm[0][1] = Lm[1][0] = LT[(size_t)a * stride + b];
                Lm[0][2] = Lm[2][0] = LT[(size_t)a * stride + cc];
                Lm[1][2] = Lm[2][1] = LT[(size_t)b * stride + cc];
                Lm[0][3] = Lm[3][0] = LT[(size_t)a * stride + e];
                Lm[1][3] = Lm[3][1] = LT[(size_t)b * stride + e];
                Lm[2][3] = Lm[3][2] = LT[(size_t)cc * stride + e];
              }
              for (int t = 0; t < 3; t++)
                if (k[t]) {
                  s.n_survivors++;
                  s.sum_counts += k[t];
                  if (digests) s.hash_surv += mix_key(make_quartet(ids, kTopos[t]), k[t]);
                  if (out) scatter_core(c, aux, leaf, Lm, kTopos[t], as_set ? 1 : (int64_t)k[t], diff[th].data());
                }
            }
          }
    }
  });
  DenseStats tot;
  for (const DenseStats& s : st) {
    tot.n_survivors += s.n_survivors;
    tot.sum_counts += s.sum_counts;
    tot.n_raw_keys += s.n_raw_keys;
    tot.sum_raw += s.sum_raw;
    tot.hash_raw += s.hash_raw;
    tot.hash_surv += s.hash_surv;
    tot.cells += s.cells;
  }
  for (const FlatTree& t : c.genes) {
    uint64_t m = 0;
    for (int i = 0; i < t.n; i++) m += t.taxon[i] >= 0;
    tot.resolutions += m < 4 ? 0 : m * (m - 1) / 2 * (m - 2) / 3 * (m - 3) / 4;
  }
  if (stats) *stats = tot;
  if (out) {
    std::vector<int64_t> sum((size_t)n * (n + 1), 0);
    for (auto& dv : diff)
      if (!dv.empty())
        for (size_t i = 0; i < sum.size(); i++) sum[i] += dv[i];
    scatter_finish(c, sum, out);
  }
  return 0;
}

// Literal quartetsTotal (quartettotals.go:78-93) for pairs (u, w) that share the LCA vertex v:
// td.Quartets(v) = surviving quartets with >= 3 taxa under v (treedata.go:197-222) is built from
// dense counts of exactly those 4-taxon sets, then every pair runs the per-quartet quartet_score.
int dense_literal_at(Ctx& c, int v, int qmode, double thresh, bool as_set, int nthreads, int n_pairs, const int32_t* pairs,
                     uint64_t* out, uint64_t* n_quartets_at_v) {
  if (!c.dense) c.dense = build_dense(c, nthreads);
  if (!c.dense) {
    c.err = "dense mode: more than 32767 gene trees or a tree deeper than 8000";
    return 1;
  }
  const Dense& d = *c.dense;
  const int N = d.N, G = d.G;
  for (int i = 0; i < n_pairs; i++)
    if (c.lca[pairs[2 * i]][pairs[2 * i + 1]] != v) {
      c.err = "dense_literal_at: a pair does not have v as its LCA";
      return 1;
    }
  std::vector<int> L;  // tip indices under v, ascending
  for (int t = 0; t < N; t++)
    if (c.in_leafset(v, t)) L.push_back(t);
  const int k = (int)L.size();
  std::vector<char> inL(N, 0);
  for (int t : L) inL[t] = 1;
  struct Tri {
    int p, q, r;
  };
  std::vector<Tri> tris;
  for (int i = 0; i < k; i++)
    for (int j = i + 1; j < k; j++)
      for (int l = j + 1; l < k; l++) tris.push_back({L[i], L[j], L[l]});
  nthreads = std::max(1, nthreads);
  std::vector<std::vector<std::pair<Quartet, uint32_t>>> found(nthreads);
  static const uint8_t kTopos[3] = {kTopo1, kTopo2, kTopo3};
  static const int kOrder[3] = {0, 1, 2};
  parallel_for((int)tris.size(), nthreads, [&](int ti, int th) {
    const int p = tris[ti].p, q = tris[ti].q, r = tris[ti].r;  // p < q < r, all under v
    std::vector<int16_t> pm;
    const int16_t* pmask = nullptr;
    if (d.any_absent) {
      pm.resize(G);
      const int16_t *pa = d.present(p), *pb = d.present(q), *pc = d.present(r);
      for (int g = 0; g < G; g++) pm[g] = pa[16 * g] & pb[16 * g] & pc[16 * g];
      pmask = pm.data();
    }
    for (int db = 0; db < d.NB; db++) {
      int16_t cnt[3][16];  // pairings pq|ro, pr|qo, qr|po
      dense_count_run(d.pair(p, q), d.pair(p, r), d.pair(q, r), &d.D[d.at(p, db, 0)], &d.D[d.at(q, db, 0)],
                      &d.D[d.at(r, db, 0)], pmask, &d.P[((size_t)db * G) * 16], G, cnt);
      for (int l = 0; l < 16; l++) {
        const int o = db * 16 + l;
        if (o >= N || o == p || o == q || o == r) continue;
        if (inL[o] && o < r) continue;  // a 4-set inside L(v) is taken from its three smallest taxa
        // sorted taxa and the pairing -> topology map: pairing i joins o with x_i = (r, q, p)[i]
        int s[4] = {p, q, r, o};
        std::sort(s, s + 4);
        const int partner_of_o[3] = {r, q, p};
        uint32_t kk[3] = {0, 0, 0};
        for (int i = 0; i < 3; i++) {
          // topology index = which sorted taxon is paired with s[0]: s[1] -> 0, s[2] -> 1, s[3] -> 2
          int mate;  // the taxon sharing a side with s[0]
          if (s[0] == o) mate = partner_of_o[i];
          else if (s[0] == partner_of_o[i]) mate = o;
          else {  // s[0] is one of the two others: its mate is the other one
            int rest[2], nr = 0;
            const int all[3] = {p, q, r};
            for (int x : all)
              if (x != partner_of_o[i]) rest[nr++] = x;
            mate = rest[0] == s[0] ? rest[1] : rest[0];
          }
          const int topo = mate == s[1] ? 0 : (mate == s[2] ? 1 : 2);
          kk[topo] = (uint32_t)cnt[i][l];
        }
        if (!(kk[0] | kk[1] | kk[2])) continue;
        if (qmode != 0) filter_cell_literal(kk, qmode, thresh, kOrder);
        kk[constraint_topology(d, s[0], s[1], s[2], s[3])] = 0;
        const int16_t ids[4] = {(int16_t)s[0], (int16_t)s[1], (int16_t)s[2], (int16_t)s[3]};
        for (int t = 0; t < 3; t++)
          if (kk[t]) found[th].emplace_back(make_quartet(ids, kTopos[t]), kk[t]);
      }
    }
  });
  std::vector<std::pair<Quartet, uint32_t>> qv;
  for (auto& f : found) qv.insert(qv.end(), f.begin(), f.end());
  if (n_quartets_at_v) *n_quartets_at_v = qv.size();
  parallel_for(n_pairs, nthreads, [&](int i, int) {
    const int u = pairs[2 * i], w = pairs[2 * i + 1];
    uint64_t total = 0;
    if (should_calc_edge(c, u, w)) {  // CalculateQuartetTotals only visits these (quartettotals.go:52-55)
      const int wSub = get_w_subtree(c, u, w, v);
      for (const auto& kv : qv)
        if (quartet_score(c, kv.first, u, w, v, wSub) == Qeq) total += as_set ? 1 : (uint64_t)kv.second;
    }
    out[i] = total;
  });
  return 0;
}
