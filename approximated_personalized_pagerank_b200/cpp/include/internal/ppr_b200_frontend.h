// ppr_b200_frontend.h -- host front-end shared by the drop-in headers grank.h / grankMulti.h / mccompletepathv2.h.
//
// The reference's public API is three C++11 function templates over
//     std::unordered_map<Key, std::vector<Key>>  ->  std::unordered_map<Key, std::unordered_map<Key, double>>
// (/root/reference/include/grank.h:42-48, header-only/grankMulti.h:289-296, include/mccompletepathv2.h:182-187).
// The templates stay in headers (Key is the caller's type); everything below the relabel step is the C-ABI of
// libppr_b200.so (include/pprb200.h). Steps:
//   1. parameter checks: same text, same order, cerr + exit(EXIT_FAILURE), before the graph is touched
//      (grank.h:51-55; test/grankTest.cc:22-28 fire them on an empty graph);
//   2. relabel: dense id = position of the key in the CALLER'S MAP ITERATION ORDER. This is parity-critical: the
//      reference's findPartitions takes component roots and builds predecessor lists in that order
//      (pprInternal.h:34-64), and the canonical tie-break is (score desc, dense id asc);
//   3. CSR with the successor-vector order and multiplicity preserved (multi-edges and self-loops count,
//      test/grankTest.cc:60,79). A successor that is not a key is an error here (the reference dereferences
//      end(): undefined behaviour, pprInternal.h:76; README.md:68-74 requires sinks to be keys);
//   4. one extern "C" call (device work, no CPU fallback: failure prints pprb200_last_error() and exits);
//   5. materialise the baskets as maps keyed by copies of the caller's keys (README.md:61-66), in parallel.
#ifndef PPR_B200_FRONTEND_H
#define PPR_B200_FRONTEND_H

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <thread>
#include <unordered_map>
#include <utility>
#include <vector>

#include <pprb200.h>

namespace ppr {
namespace b200 {

inline void die(const char* msg) {
  std::cerr << msg << std::endl;
  std::exit(EXIT_FAILURE);
}

// grank.h:51-55 / grankMulti.h:299-304 / mccompletepathv2.h:190-194
inline void checkParameters(size_t K, size_t L, size_t iterations, double damping) {
  if (K == 0) die("K must be positive");
  if (L == 0) die("L must be positive");
  if (K > L) die("K must be <= L");
  if (iterations == 0) die("iterations must be positive");
  if (damping < 0 || damping > 1) die("damping must be [0,1]");
}

template <typename Key>
struct DenseGraph {
  std::vector<const Key*> keyOf;  // dense id -> caller's key (pointers into the caller's map, alive for the call)
  std::vector<int64_t> rowPtr;
  std::vector<int32_t> col;
};

inline size_t hostThreads() {
  const unsigned hc = std::thread::hardware_concurrency();
  return hc ? hc : 1;
}

// runs fn(begin, end) over [0, n) on up to `threads` std::threads (inline when the range is small)
template <typename Fn>
inline void parallelRanges(size_t n, size_t threads, size_t grain, Fn fn) {
  const size_t parts = std::min(threads ? threads : 1, (n + grain - 1) / (grain ? grain : 1));
  if (parts <= 1) { fn((size_t)0, n); return; }
  std::vector<std::thread> pool;
  for (size_t t = 1; t < parts; t++) pool.emplace_back(fn, n * t / parts, n * (t + 1) / parts);
  fn((size_t)0, n / parts);
  for (auto& th : pool) th.join();
}

template <typename Key>
DenseGraph<Key> relabel(const std::unordered_map<Key, std::vector<Key>>& graph, size_t threads = hostThreads()) {
  DenseGraph<Key> g;
  const size_t n = graph.size();
  if (n > (size_t)1 << 30) die("graphs with more than 2^30 nodes are not supported");
  g.keyOf.reserve(n);
  g.rowPtr.reserve(n + 1);
  std::unordered_map<Key, int32_t> idOf;
  idOf.reserve(n);
  std::vector<const std::vector<Key>*> succOf;
  succOf.reserve(n);
  g.rowPtr.push_back(0);
  for (const auto& kv : graph) {  // map iteration order defines the dense ids
    idOf.emplace(kv.first, (int32_t)g.keyOf.size());
    g.keyOf.push_back(&kv.first);
    succOf.push_back(&kv.second);
    g.rowPtr.push_back(g.rowPtr.back() + (int64_t)kv.second.size());
  }
  g.col.resize((size_t)g.rowPtr.back());
  // the successor lookups (one hash find per edge) only read idOf: node ranges in parallel
  std::vector<char> missing(1, 0);
  parallelRanges(n, threads, 4096, [&](size_t begin, size_t end) {
    for (size_t v = begin; v < end; v++) {
      int64_t o = g.rowPtr[v];
      for (const Key& s : *succOf[v]) {
        const auto it = idOf.find(s);
        if (it == idOf.end()) { missing[0] = 1; return; }
        g.col[(size_t)o++] = it->second;
      }
    }
  });
  if (missing[0]) die("successor is not a key of the graph: nodes without edges must be mapped to an empty vector");
  return g;
}

template <typename Key>
std::unordered_map<Key, std::unordered_map<Key, double>> materialise(const DenseGraph<Key>& g, size_t K,
                                                                     const std::vector<int32_t>& ids,
                                                                     const std::vector<double>& scores,
                                                                     const std::vector<uint32_t>& cnt, size_t nThreads) {
  const size_t n = g.keyOf.size();
  std::vector<std::unordered_map<Key, double>> inner(n);
  auto fill = [&](size_t begin, size_t end) {
    for (size_t v = begin; v < end; v++) {
      std::unordered_map<Key, double>& m = inner[v];
      m.reserve(cnt[v]);
      for (uint32_t i = 0; i < cnt[v]; i++) m.emplace(*g.keyOf[(size_t)ids[v * K + i]], scores[v * K + i]);
    }
  };
  if (nThreads <= 1 || n < 4096) {
    fill(0, n);
  } else {
    std::vector<std::thread> pool;
    const size_t step = (n + nThreads - 1) / nThreads;
    for (size_t t = 0; t < nThreads; t++) {
      const size_t b = t * step, e = b + step < n ? b + step : n;
      if (b < e) pool.emplace_back(fill, b, e);
    }
    for (auto& th : pool) th.join();
  }
  std::unordered_map<Key, std::unordered_map<Key, double>> out;
  out.reserve(n);
  for (size_t v = 0; v < n; v++) out.emplace(*g.keyOf[v], std::move(inner[v]));
  return out;
}

// ---- flat result view ---------------------------------------------------------------------------
// At a million nodes and more, building n * K hash-map entries on the host costs more than the kernels. FlatBaskets
// is the same result without the maps: node v (dense id = position in the caller's map iteration order) has
// cnt[v] <= K entries ids[v*K + i] / scores[v*K + i], sorted by (score descending, dense id ascending); key(v) maps a
// dense id back to the caller's key. toMaps() gives the reference's return type.
template <typename Key>
struct FlatBaskets {
  size_t K = 0;
  std::vector<const Key*> keyOf;  // pointers into the caller's graph: valid while the graph is
  std::vector<int32_t> ids;
  std::vector<double> scores;
  std::vector<uint32_t> cnt;
  size_t size() const { return keyOf.size(); }
  const Key& key(size_t denseId) const { return *keyOf[denseId]; }
  std::unordered_map<Key, std::unordered_map<Key, double>> toMaps(size_t threads = hostThreads()) const {
    DenseGraph<Key> g;
    g.keyOf = keyOf;
    return materialise(g, K, ids, scores, cnt, threads);
  }
};

template <typename Key>
FlatBaskets<Key> grankFlat(const std::unordered_map<Key, std::vector<Key>>& graph, size_t K, size_t L, size_t iterations,
                           double damping, double tolerance) {
  checkParameters(K, L, iterations, damping);
  FlatBaskets<Key> out;
  out.K = K;
  if (graph.empty()) return out;
  if (K > 0xffffffffu || L > 0xffffffffu || iterations > 0xffffffffu) die("K, L and iterations must fit 32 bits");
  DenseGraph<Key> g = relabel(graph);
  const size_t n = g.keyOf.size();
  out.ids.resize(n * K);
  out.scores.resize(n * K);
  out.cnt.resize(n);
  const int rc = pprb200_grank(g.rowPtr.data(), g.col.data(), (int32_t)n, NULL, (uint32_t)K, (uint32_t)L, (uint32_t)iterations,
                               damping, tolerance, 0, out.ids.data(), out.scores.data(), out.cnt.data(), NULL);
  if (rc != PPRB200_OK) die(pprb200_last_error());
  out.keyOf.swap(g.keyOf);
  return out;
}

template <typename Key>
FlatBaskets<Key> mccompletepathv2Flat(const std::unordered_map<Key, std::vector<Key>>& graph, size_t K, size_t L,
                                      size_t iterations, double damping) {
  checkParameters(K, L, iterations, damping);
  FlatBaskets<Key> out;
  out.K = K;
  if (graph.empty()) return out;
  if (K > 0xffffffffu || L > 0xffffffffu || iterations > 0xffffffffu) die("K, L and iterations must fit 32 bits");
  DenseGraph<Key> g = relabel(graph);
  const size_t n = g.keyOf.size();
  out.ids.resize(n * K);
  out.scores.resize(n * K);
  out.cnt.resize(n);
  uint64_t seed = PPRB200_DEFAULT_MC_SEED;
  if (const char* e = std::getenv("PPRB200_MC_SEED")) seed = std::strtoull(e, NULL, 0);
  const int rc = pprb200_mccompletepathv2(g.rowPtr.data(), g.col.data(), (int32_t)n, (uint32_t)K, (uint32_t)L, (uint32_t)iterations,
                                          damping, seed, PPRB200_DEFAULT_MC_ROUNDS, 0, out.ids.data(), out.scores.data(),
                                          out.cnt.data(), NULL);
  if (rc != PPRB200_OK) die(pprb200_last_error());
  out.keyOf.swap(g.keyOf);
  return out;
}

template <typename Key>
std::unordered_map<Key, std::unordered_map<Key, double>> runGrank(const std::unordered_map<Key, std::vector<Key>>& graph,
                                                                  size_t K, size_t L, size_t iterations, double damping,
                                                                  double tolerance, size_t hostThreadsForMaps) {
  if (graph.empty()) return std::unordered_map<Key, std::unordered_map<Key, double>>();  // test/grankTest.cc:31-36
  if (K > 0xffffffffu || L > 0xffffffffu || iterations > 0xffffffffu) die("K, L and iterations must fit 32 bits");
  DenseGraph<Key> g = relabel(graph);
  const size_t n = g.keyOf.size();
  std::vector<int32_t> ids(n * K);
  std::vector<double> scores(n * K);
  std::vector<uint32_t> cnt(n);
  // colour = NULL: the library runs the reference's findPartitions on the dense graph (pprInternal.h:29-99)
  const int rc = pprb200_grank(g.rowPtr.data(), g.col.data(), (int32_t)n, NULL, (uint32_t)K, (uint32_t)L,
                               (uint32_t)iterations, damping, tolerance, 0, ids.data(), scores.data(), cnt.data(), NULL);
  if (rc != PPRB200_OK) die(pprb200_last_error());
  return materialise(g, K, ids, scores, cnt, hostThreadsForMaps);
}

template <typename Key>
std::unordered_map<Key, std::unordered_map<Key, double>> runMc(const std::unordered_map<Key, std::vector<Key>>& graph, size_t K,
                                                               size_t L, size_t iterations, double damping) {
  if (graph.empty()) return std::unordered_map<Key, std::unordered_map<Key, double>>();
  if (K > 0xffffffffu || L > 0xffffffffu || iterations > 0xffffffffu) die("K, L and iterations must fit 32 bits");
  DenseGraph<Key> g = relabel(graph);
  const size_t n = g.keyOf.size();
  std::vector<int32_t> ids(n * K);
  std::vector<double> scores(n * K);
  std::vector<uint32_t> cnt(n);
  uint64_t seed = PPRB200_DEFAULT_MC_SEED;
  if (const char* e = std::getenv("PPRB200_MC_SEED")) seed = std::strtoull(e, NULL, 0);
  const int rc = pprb200_mccompletepathv2(g.rowPtr.data(), g.col.data(), (int32_t)n, (uint32_t)K, (uint32_t)L,
                                          (uint32_t)iterations, damping, seed, PPRB200_DEFAULT_MC_ROUNDS, 0, ids.data(),
                                          scores.data(), cnt.data(), NULL);
  if (rc != PPRB200_OK) die(pprb200_last_error());
  return materialise(g, K, ids, scores, cnt, hostThreads());
}

}  // namespace b200
}  // namespace ppr

#endif
