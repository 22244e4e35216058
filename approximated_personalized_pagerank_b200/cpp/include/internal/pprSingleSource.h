// pprSingleSource.h -- drop-in replacement of /root/reference/include/internal/pprSingleSource.h: exact
// Personalized PageRank of one source by power iteration, same namespace, signature, parameter checks and messages
// (pprSingleSource.h:37-39). The iteration runs on the GPU (pprb200_ppr_exact with a batch of one; use
// benchmarkAlgorithm.h or the C-ABI directly to advance many sources together).
#ifndef PPRSINGLESOURCE_H
#define PPRSINGLESOURCE_H

#include <unordered_map>
#include <vector>

#include <internal/ppr_b200_frontend.h>

namespace ppr {
namespace pprInternal {

/**
 * @param graph      node -> successors (nodes without edges must be keys).
 * @param iterations max number of iterations.
 * @param damping    damping factor in [0,1].
 * @param tolerance  stop once the norm-1 change of an iteration is below it; negative = never.
 * @param source     node whose personalized pagerank is computed.
 * @return           score of every node the iteration reached (the source is always a key).
 */
template <typename Key>
std::unordered_map<Key, double> pprSingleSource(const std::unordered_map<Key, std::vector<Key>>& graph, size_t iterations,
                                                double damping, double tolerance, Key source) {
  if (iterations == 0) b200::die("iterations must be positive");
  if (damping < 0 || damping > 1) b200::die("damping must be [0,1]");
  if (graph.find(source) == graph.end()) b200::die("source node not part of the graph");
  if (iterations > 0xffffffffu) b200::die("iterations must fit 32 bits");
  b200::DenseGraph<Key> g = b200::relabel(graph);
  const size_t n = g.keyOf.size();
  int32_t src = -1;
  for (size_t v = 0; v < n && src < 0; v++)
    if (*g.keyOf[v] == source) src = (int32_t)v;
  std::vector<double> scores(n);
  const int rc = pprb200_ppr_exact(g.rowPtr.data(), g.col.data(), (int32_t)n, &src, 1, (uint32_t)iterations, damping, tolerance,
                                   scores.data(), NULL, NULL);
  if (rc != PPRB200_OK) b200::die(pprb200_last_error());
  std::unordered_map<Key, double> out;
  for (size_t v = 0; v < n; v++)
    if (scores[v] != 0.0 || (int32_t)v == src) out.emplace(*g.keyOf[v], scores[v]);
  return out;
}

}  // namespace pprInternal
}  // namespace ppr

#endif
