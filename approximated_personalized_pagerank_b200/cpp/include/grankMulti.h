// grankMulti.h -- drop-in replacement of /root/reference/header-only/grankMulti.h:289-436. The reference fans the
// source nodes out over nThreads std::threads; here the data parallelism over sources happens on the GPU, so
// nThreads is validated like the reference does (:304) and otherwise only sizes the host threads that build the
// result maps. Results equal ppr::grank's (test/grankMultiThreadTest.cc:384-576).
#ifndef GRANKMULTI_H
#define GRANKMULTI_H

#include <unordered_map>
#include <vector>

#include <internal/ppr_b200_frontend.h>

namespace ppr
{
  template<typename Key>
  std::unordered_map<Key, std::unordered_map<Key, double>> grankMulti(const std::unordered_map<Key, std::vector<Key>>& graph,
  size_t K, size_t L, size_t iterations, double damping, double tolerance, size_t nThreads)
  {
    b200::checkParameters(K, L, iterations, damping);
    if(nThreads == 0) b200::die("nThreads must be positive");
    return b200::runGrank(graph, K, L, iterations, damping, tolerance, nThreads);
  }
}

#endif
