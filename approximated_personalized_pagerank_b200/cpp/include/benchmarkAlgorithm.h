// benchmarkAlgorithm.h -- drop-in replacement of /root/reference/include/benchmarkAlgorithm.h: same
// namespace, template signature, statistic names, sampling rule and error behaviour. The reference runs one exact
// power iteration (pprSingleSource, 100 iterations / damping 0.85 / tolerance 0.0001, benchmarkAlgorithm.h:91) per sampled
// node on the host; here all sampled nodes advance together on the GPU (pprb200_ppr_exact), the five statistics are then
// computed on the host.
#ifndef BENCHMARKALGORITHM_H
#define BENCHMARKALGORITHM_H

#include <algorithm>
#include <cmath>
#include <iostream>
#include <random>
#include <string>
#include <unordered_map>
#include <vector>

#include <internal/ppr_b200_frontend.h>

namespace ppr {
namespace b200 {

// Kendall tau-b as the reference computes it (include/internal/kendall.h:22-180): pairs tied in x or in y count for
// neither side, the denominator is sqrt((pairs - tiedX) * (pairs - tiedY)); a zero denominator gives 1 when as many pairs
// are tied in x as in y and 0 otherwise. Pairs are counted directly (the vectors are baskets: a few hundred entries).
inline double kendallTauB(const std::vector<double>& x, const std::vector<double>& y) {
  const size_t n = x.size() < y.size() ? x.size() : y.size();
  unsigned long long tiedX = 0, tiedY = 0;
  long long balance = 0;  // concordant - discordant
  for (size_t i = 0; i < n; i++)
    for (size_t j = i + 1; j < n; j++) {
      const int sx = (x[i] > x[j]) - (x[i] < x[j]);
      const int sy = (y[i] > y[j]) - (y[i] < y[j]);
      tiedX += sx == 0;
      tiedY += sy == 0;
      balance += sx * sy;
    }
  const unsigned long long pairs = n ? (unsigned long long)n * (n - 1) / 2 : 0;
  const long double den = std::sqrt((long double)(pairs - tiedX) * (long double)(pairs - tiedY));
  if (den == 0.0L) return tiedX == tiedY ? 1.0 : 0.0;
  return (double)((long double)balance / den);
}

}  // namespace b200

/**
 * Compares the provided top-K baskets with exact Personalized PageRank for `testNodes` randomly chosen source nodes
 * (reference include/benchmarkAlgorithm.h:51-153).
 * @param ppr       source node -> basket of scores produced by grank / grankMulti / mccompletepathv2.
 * @param graph     the graph those scores were computed on.
 * @param testNodes number of sampled nodes (= exact PPR computations).
 * @param strict    skip nodes without out-edges when sampling.
 * @return "jaccard average", "jaccard min", "kendall average", "kendall min", "average map size"; all -1 when no node
 *         could be sampled.
 */
template <typename Key>
std::unordered_map<std::string, double> benchmarkAlgorithm(const std::unordered_map<Key, std::unordered_map<Key, double>>& ppr,
                                                           const std::unordered_map<Key, std::vector<Key>>& graph, size_t testNodes,
                                                           bool strict) {
  if (testNodes == 0) b200::die("testNodes must be positive");
  std::unordered_map<std::string, double> result;
  std::vector<const Key*> nodes;
  for (const auto& kv : ppr) {
    const auto it = graph.find(kv.first);
    if (it == graph.end()) {
      std::cerr << "node " << kv.first << " in the provided map is not part of the provided graph" << std::endl;
      std::exit(EXIT_FAILURE);
    }
    if (!strict || !it->second.empty()) nodes.push_back(&kv.first);
  }
  std::random_device rd;
  std::mt19937 gen(rd());
  std::shuffle(nodes.begin(), nodes.end(), gen);
  const size_t samples = std::min(nodes.size(), testNodes);
  if (samples == 0) {
    result["jaccard average"] = -1;
    result["jaccard min"] = -1;
    result["kendall average"] = -1;
    result["kendall min"] = -1;
    result["average map size"] = -1;
    return result;
  }

  b200::DenseGraph<Key> g = b200::relabel(graph);
  const size_t n = g.keyOf.size();
  std::unordered_map<Key, int32_t> idOf;
  idOf.reserve(n);
  for (size_t v = 0; v < n; v++) idOf.emplace(*g.keyOf[v], (int32_t)v);
  std::vector<int32_t> sources(samples);
  for (size_t i = 0; i < samples; i++) sources[i] = idOf.find(*nodes[i])->second;
  std::vector<double> exact(samples * n);
  const int rc = pprb200_ppr_exact(g.rowPtr.data(), g.col.data(), (int32_t)n, sources.data(), (uint32_t)samples, 100, 0.85, 0.0001,
                                   exact.data(), NULL, NULL);
  if (rc != PPRB200_OK) b200::die(pprb200_last_error());

  double jaccardAverage = 0, jaccardMin = 1.0, kendallAverage = 0, kendallMin = 1.0, averageMapSize = 0;
  std::vector<int32_t> reached;
  std::vector<char> inBasket(n, 0);
  for (size_t i = 0; i < samples; i++) {
    const std::unordered_map<Key, double>& basket = ppr.find(*nodes[i])->second;
    const double* row = exact.data() + i * n;
    // the exact top-|basket| among the nodes the power iteration reached
    reached.clear();
    for (size_t v = 0; v < n; v++)
      if (row[v] > 0) reached.push_back((int32_t)v);
    const size_t keep = std::min(basket.size(), reached.size());
    if (keep < reached.size())
      std::nth_element(reached.begin(), reached.begin() + (long)keep, reached.end(),
                       [&](int32_t a, int32_t b) { return row[a] > row[b] || (row[a] == row[b] && a < b); });
    std::vector<double> ours, theirs;
    ours.reserve(basket.size());
    theirs.reserve(basket.size());
    std::vector<int32_t> ids;
    ids.reserve(basket.size());
    for (const auto& e : basket) {
      const auto it = idOf.find(e.first);
      const int32_t id = it == idOf.end() ? -1 : it->second;
      ids.push_back(id);
      if (id >= 0) inBasket[(size_t)id] = 1;
      ours.push_back(e.second);
      theirs.push_back(id >= 0 ? row[id] : 0.0);
    }
    size_t common = 0;
    for (size_t k = 0; k < keep; k++) common += inBasket[(size_t)reached[k]];
    for (const int32_t id : ids)
      if (id >= 0) inBasket[(size_t)id] = 0;
    const size_t uni = basket.size() + keep - common;
    const double jac = uni == 0 ? 1.0 : (double)common / (double)uni;
    const double ken = b200::kendallTauB(ours, theirs);
    jaccardAverage += jac;
    jaccardMin = std::min(jaccardMin, jac);
    kendallAverage += ken;
    kendallMin = std::min(kendallMin, ken);
    averageMapSize += (double)basket.size();
  }
  result["jaccard average"] = jaccardAverage / (double)samples;
  result["jaccard min"] = jaccardMin;
  result["kendall average"] = kendallAverage / (double)samples;
  result["kendall min"] = kendallMin;
  result["average map size"] = averageMapSize / (double)samples;
  return result;
}

}  // namespace ppr

#endif
