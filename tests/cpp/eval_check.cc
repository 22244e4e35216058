// eval_check.cc -- one user program against the reference's evaluator API (include/benchmarkAlgorithm.h:51-153), two
// builds: tests/cpp/eval_check_b200 (our headers + libppr_b200.so) and oracle/_ref/eval_check_ref (the unmodified
// reference headers). With testNodes >= the number of candidate nodes every node is sampled, so the random shuffle does
// not matter and the two builds must report the same statistics (up to arbitrary ties inside the exact top-K).
#include <cstdio>
#include <string>
#include <unordered_map>
#include <vector>

#include <grank.h>
#include <mccompletepathv2.h>
#include <benchmarkAlgorithm.h>

static unsigned lcg(unsigned& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

int main() {
  std::unordered_map<int, std::vector<int>> g;
  unsigned s = 99u;
  const int n = 400;
  for (int i = 0; i < n; i++) g[i];
  for (int e = 0; e < 10 * n; e++) {
    const int a = (int)(lcg(s) % n), b = (int)((unsigned long long)(lcg(s) % n) * (lcg(s) % n) / n);
    if (a % 9 != 0) g[a].push_back(b);
  }
  const auto res = ppr::grank(g, 30, 60, 30, 0.85, 0.0001);
  const auto st = ppr::benchmarkAlgorithm(res, g, 100000, true);
  const char* names[5] = {"jaccard average", "jaccard min", "kendall average", "kendall min", "average map size"};
  for (int i = 0; i < 5; i++) std::printf("%s = %.12f\n", names[i], st.at(names[i]));
  std::unordered_map<int, std::vector<int>> sinks;
  for (int i = 0; i < 5; i++) sinks[i];
  const auto none = ppr::benchmarkAlgorithm(ppr::grank(sinks, 1, 1, 1, 0.85, 0.1), sinks, 10, true);
  std::printf("no samples = %.1f\n", none.at("jaccard average"));
  return 0;
}
