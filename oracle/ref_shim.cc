// oracle/ref_shim.cc -- TEST INFRASTRUCTURE ONLY.
//
// Thin C-ABI wrapper that instantiates the UNMODIFIED reference templates for Key=int from the sources
// where they lie under /root/reference (nothing is copied into this repo). Built by oracle/Makefile into
// oracle/_ref/libppr_ref.so (git-ignored, travels to the GPU box with the snapshot).
//
// One TU only: include/mccompletepathv2.h and include/internal/kendall.h define non-inline globals
// (mccompletepathv2.h:32-34), so a second TU would violate the ODR.
//
// The graph handed to the reference is always built the same way -- keys 0..n-1 inserted in ascending
// order, successor vectors in CSR order -- so the unordered_map iteration order is reproducible;
// ref_iteration_order() exposes it, because the canonical dense id of the restatement/CUDA path is the
// position in that order (SURVEY.md 8b "Relabel rule").
#include <chrono>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

// A second key type with a different std::hash: the same graph iterates in another order, so partitions, summation
// order and every boundary tie may fall differently -- the reference's self-noise band (SURVEY.md 7-1, App. B rmat_probe).
struct AltKey {
  int v;
  bool operator==(const AltKey& o) const { return v == o.v; }
};
namespace std {
template <>
struct hash<AltKey> {
  size_t operator()(const AltKey& k) const {
    unsigned long long x = (unsigned long long)(unsigned int)k.v * 0x9E3779B97F4A7C15ull;
    return (size_t)(x ^ (x >> 29));
  }
};
}  // namespace std

#include <grank.h>             // /root/reference/include/grank.h
#include <grankMulti.h>        // /root/reference/header-only/grankMulti.h
#include <mccompletepathv2.h>  // /root/reference/include/mccompletepathv2.h
#include <pprSingleSource.h>   // /root/reference/include/internal/pprSingleSource.h
#include <kendall.h>           // /root/reference/include/internal/kendall.h

typedef std::unordered_map<int, std::vector<int>> graph_t;
typedef std::unordered_map<int, std::unordered_map<int, double>> result_t;

static graph_t build_graph(const int64_t* row_ptr, const int32_t* col, int32_t n) {
  graph_t g;
  for (int32_t v = 0; v < n; v++) {
    std::vector<int>& s = g[v];
    s.assign(col + row_ptr[v], col + row_ptr[v + 1]);
  }
  return g;
}

// flatten: per key v (original key space), up to `cap` (id, score) pairs in map iteration order, pad id=-1
static void flatten(const result_t& res, int32_t n, uint32_t cap, int32_t* out_ids, double* out_scores, uint32_t* out_cnt) {
  for (int32_t v = 0; v < n; v++) {
    uint32_t c = 0;
    auto it = res.find(v);
    if (it != res.end())
      for (const auto& kv : it->second) {
        if (c < cap) { out_ids[(size_t)v * cap + c] = kv.first; out_scores[(size_t)v * cap + c] = kv.second; }
        c++;
      }
    out_cnt[v] = c;
    for (uint32_t i = c; i < cap; i++) { out_ids[(size_t)v * cap + i] = -1; out_scores[(size_t)v * cap + i] = 0.0; }
  }
}

extern "C" {

int ref_hardware_concurrency(void) { return (int)std::thread::hardware_concurrency(); }

// order[i] = key visited i-th when iterating the map
int ref_iteration_order(const int64_t* row_ptr, const int32_t* col, int32_t n, int32_t* order) {
  graph_t g = build_graph(row_ptr, col, n);
  int32_t i = 0;
  for (const auto& kv : g) order[i++] = kv.first;
  return 0;
}

// colour[key] = 0 if key in partitions.first else 1 (pprInternal.h:29-99)
int ref_find_partitions(const int64_t* row_ptr, const int32_t* col, int32_t n, uint8_t* colour) {
  graph_t g = build_graph(row_ptr, col, n);
  auto parts = ppr::pprInternal::findPartitions<int>(g);
  for (int32_t v = 0; v < n; v++) colour[v] = parts.first.count(v) ? 0 : (parts.second.count(v) ? 1 : 255);
  return 0;
}

// grank.h:42-150. out_* sized n*K. seconds = steady_clock around the call as src/main.cc:36-39
int ref_grank(const int64_t* row_ptr, const int32_t* col, int32_t n, uint32_t K, uint32_t L, uint32_t iterations,
              double damping, double tolerance, int32_t* out_ids, double* out_scores, uint32_t* out_cnt, double* seconds) {
  graph_t g = build_graph(row_ptr, col, n);
  auto t0 = std::chrono::steady_clock::now();
  result_t res = ppr::grank<int>(g, K, L, iterations, damping, tolerance);
  auto t1 = std::chrono::steady_clock::now();
  if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
  if (out_ids) flatten(res, n, K, out_ids, out_scores, out_cnt);
  return 0;
}

// grank.h:42-150 instantiated for AltKey (same graph, same parameters, different hash order)
int ref_grank_althash(const int64_t* row_ptr, const int32_t* col, int32_t n, uint32_t K, uint32_t L, uint32_t iterations,
                      double damping, double tolerance, int32_t* out_ids, double* out_scores, uint32_t* out_cnt) {
  std::unordered_map<AltKey, std::vector<AltKey>> g;
  for (int32_t v = 0; v < n; v++) {
    std::vector<AltKey>& s = g[AltKey{v}];
    for (int64_t i = row_ptr[v]; i < row_ptr[v + 1]; i++) s.push_back(AltKey{col[i]});
  }
  auto res = ppr::grank<AltKey>(g, K, L, iterations, damping, tolerance);
  for (int32_t v = 0; v < n; v++) {
    uint32_t c = 0;
    auto it = res.find(AltKey{v});
    if (it != res.end())
      for (const auto& kv : it->second) {
        if (c < K) { out_ids[(size_t)v * K + c] = kv.first.v; out_scores[(size_t)v * K + c] = kv.second; }
        c++;
      }
    out_cnt[v] = c;
    for (uint32_t i = c; i < K; i++) { out_ids[(size_t)v * K + i] = -1; out_scores[(size_t)v * K + i] = 0.0; }
  }
  return 0;
}

// header-only/grankMulti.h:289-436
int ref_grankMulti(const int64_t* row_ptr, const int32_t* col, int32_t n, uint32_t K, uint32_t L, uint32_t iterations,
                   double damping, double tolerance, uint32_t nThreads, int32_t* out_ids, double* out_scores,
                   uint32_t* out_cnt, double* seconds) {
  graph_t g = build_graph(row_ptr, col, n);
  auto t0 = std::chrono::steady_clock::now();
  result_t res = ppr::grankMulti<int>(g, K, L, iterations, damping, tolerance, nThreads);
  auto t1 = std::chrono::steady_clock::now();
  if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
  if (out_ids) flatten(res, n, K, out_ids, out_scores, out_cnt);
  return 0;
}

// mccompletepathv2.h:182-258 (random_device-seeded: not reproducible run to run)
int ref_mccompletepathv2(const int64_t* row_ptr, const int32_t* col, int32_t n, uint32_t K, uint32_t L, uint32_t iterations,
                         double damping, int32_t* out_ids, double* out_scores, uint32_t* out_cnt, double* seconds) {
  graph_t g = build_graph(row_ptr, col, n);
  auto t0 = std::chrono::steady_clock::now();
  result_t res = ppr::mccompletepathv2<int>(g, K, L, iterations, damping);
  auto t1 = std::chrono::steady_clock::now();
  if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
  if (out_ids) flatten(res, n, K, out_ids, out_scores, out_cnt);
  return 0;
}

// pprSingleSource.h:28-75 -> dense vector out[n]
int ref_ppr_single_source(const int64_t* row_ptr, const int32_t* col, int32_t n, uint32_t iterations, double damping,
                          double tolerance, int32_t source, double* out) {
  graph_t g = build_graph(row_ptr, col, n);
  auto res = ppr::pprInternal::pprSingleSource<int>(g, iterations, damping, tolerance, source);
  std::memset(out, 0, sizeof(double) * (size_t)n);
  for (const auto& kv : res) out[kv.first] = kv.second;
  return 0;
}

// several sources against one graph build (the evaluator pattern, benchmarkAlgorithm.h:91)
int ref_ppr_multi_source(const int64_t* row_ptr, const int32_t* col, int32_t n, uint32_t iterations, double damping,
                         double tolerance, const int32_t* sources, int32_t nsources, double* out /* nsources*n */) {
  graph_t g = build_graph(row_ptr, col, n);
  for (int32_t i = 0; i < nsources; i++) {
    auto res = ppr::pprInternal::pprSingleSource<int>(g, iterations, damping, tolerance, sources[i]);
    double* o = out + (size_t)i * n;
    std::memset(o, 0, sizeof(double) * (size_t)n);
    for (const auto& kv : res) o[kv.first] = kv.second;
  }
  return 0;
}

// keepTop (pprInternal.h:109-137) on one map given as parallel arrays; returns kept count, arrays rewritten
int ref_keep_top(uint32_t L, int32_t* ids, double* scores, int32_t cnt) {
  std::unordered_map<int, double> m;
  for (int32_t i = 0; i < cnt; i++) m[ids[i]] = scores[i];
  ppr::pprInternal::keepTop<int>(L, m);
  int32_t c = 0;
  for (const auto& kv : m) { ids[c] = kv.first; scores[c] = kv.second; c++; }
  return c;
}

// norm1 (pprInternal.h:147-165)
double ref_norm1(const int32_t* ids1, const double* s1, int32_t c1, const int32_t* ids2, const double* s2, int32_t c2) {
  std::unordered_map<int, double> a, b;
  for (int32_t i = 0; i < c1; i++) a[ids1[i]] = s1[i];
  for (int32_t i = 0; i < c2; i++) b[ids2[i]] = s2[i];
  return ppr::pprInternal::norm1<int>(a, b);
}

}  // extern "C"

// kendall.h:22-180 and pprInternal.h:174-186 (the quality evaluator's two statistics)
extern "C" double ref_kendall(const double* x, const double* y, int32_t n) {
  return kendallCorrelation(std::vector<double>(x, x + n), std::vector<double>(y, y + n));
}

extern "C" double ref_jaccard(const int32_t* a, int32_t na, const int32_t* b, int32_t nb) {
  return ppr::pprInternal::jaccard<int>(std::unordered_set<int>(a, a + na), std::unordered_set<int>(b, b + nb));
}
