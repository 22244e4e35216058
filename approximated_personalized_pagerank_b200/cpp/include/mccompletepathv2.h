// mccompletepathv2.h -- drop-in replacement of /root/reference/include/mccompletepathv2.h:182-258 (and the
// header-only copy). Same signature and score scale (expected visits per walk = PPR / (1 - damping)). Unlike the
// reference (global random_device-seeded mt19937, :32-34) the result is deterministic: Philox streams keyed
// (source, walk) with the seed PPRB200_DEFAULT_MC_SEED, overridable through the environment variable
// PPRB200_MC_SEED; the template is re-entrant and safe to include in several translation units.
#ifndef MCCOMPLETEPATHV2_H
#define MCCOMPLETEPATHV2_H

#include <unordered_map>
#include <vector>

#include <internal/ppr_b200_frontend.h>

namespace ppr
{
  /**
   * @param iterations number of Monte-Carlo walks per node in the worst case (R); floor(R * damping) walks are run.
   */
  template<typename Key>
  std::unordered_map<Key, std::unordered_map<Key, double>> mccompletepathv2(const std::unordered_map<Key, std::vector<Key>>& graph,
  size_t K, size_t L, size_t iterations, double damping)
  {
    b200::checkParameters(K, L, iterations, damping);
    return b200::runMc(graph, K, L, iterations, damping);
  }
}

#endif
