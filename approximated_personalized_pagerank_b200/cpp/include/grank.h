// grank.h -- drop-in replacement of /root/reference/include/grank.h (and header-only/grank.h): same namespace,
// template signature, parameter meaning and error behaviour; the work runs on a B200 through libppr_b200.so.
#ifndef GRANK_H
#define GRANK_H

#include <unordered_map>
#include <vector>

#include <internal/ppr_b200_frontend.h>

namespace ppr
{
  /**
   * Approximated Personalized Pagerank for all nodes in the graph (GRank, reference include/grank.h:42-150).
   * @param graph      node -> successors; nodes without edges must be keys mapped to an empty vector.
   * @param K          entries returned per source (top-K), K <= L.
   * @param L          entries kept per source during the computation (top-L).
   * @param iterations max number of iterations (one of the two BFS partitions is updated per iteration).
   * @param damping    damping factor in [0,1].
   * @param tolerance  stop once the max norm-1 change of both partitions is below it; negative = never.
   * @return           for every node its top-K basket.
   */
  template<typename Key>
  std::unordered_map<Key, std::unordered_map<Key, double>> grank(const std::unordered_map<Key, std::vector<Key>>& graph,
  size_t K, size_t L, size_t iterations, double damping, double tolerance)
  {
    b200::checkParameters(K, L, iterations, damping);
    return b200::runGrank(graph, K, L, iterations, damping, tolerance, b200::hostThreads());
  }
}

#endif
