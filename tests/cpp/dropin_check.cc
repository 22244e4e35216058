// dropin_check.cc -- one user program, two builds:
//   tests/cpp/dropin_check_b200   compiled against approximated_personalized_pagerank_b200/cpp/include (+ libppr_b200.so)
//   oracle/_ref/dropin_check_ref  compiled against the UNMODIFIED reference headers under /root/reference
// It is written only against the reference's public API (README.md:36-40) -- #include <grank.h>, <grankMulti.h>,
// <mccompletepathv2.h>; ppr::grank / ppr::grankMulti / ppr::mccompletepathv2 over unordered_map -- so that it
// compiling and giving the same answers under both include paths IS the drop-in claim.
//
//   dropin_check <case>      prints the result maps in a canonical text form (keys sorted, %.17g)
//   dropin_check death <n>   calls an API with a bad parameter: must print the reference's message and exit(1)
//
// The cases re-express the reference's own gtest cases (test/grankTest.cc, test/grankMultiThreadTest.cc,
// test/mccompletepathv2Test.cc; cited per case); tests/test_dropin.py diffs the two builds' output.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include <grank.h>
#include <grankMulti.h>
#include <mccompletepathv2.h>

using std::string;
using std::unordered_map;
using std::vector;

template <typename Key>
static void dump(const char* name, const unordered_map<Key, unordered_map<Key, double>>& res) {
  std::map<Key, std::map<Key, double>> sorted;
  for (const auto& kv : res) sorted[kv.first] = std::map<Key, double>(kv.second.begin(), kv.second.end());
  std::cout << "case " << name << " nodes " << sorted.size() << "\n";
  for (const auto& kv : sorted) {
    std::cout << kv.first << ":";
    for (const auto& e : kv.second) {
      char buf[64];
      snprintf(buf, sizeof(buf), "%.17g", e.second);
      std::cout << " " << e.first << "=" << buf;
    }
    std::cout << "\n";
  }
}

// deterministic LCG so both builds see the same "random" graph (the reference's tests use an unseeded engine)
static unsigned lcg(unsigned& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

static unordered_map<int, vector<int>> ring(int n) {
  unordered_map<int, vector<int>> g;
  for (int i = 0; i < n; i++) g[i].push_back((i + 1) % n);
  return g;
}

static int run_case(const string& c) {
  if (c == "readme_ring") {  // README.md:105-115 == BASELINE configs[0]
    dump("readme_ring", ppr::grank(ring(100), 50, 100, 30, 0.85, 0.001));
  } else if (c == "empty") {  // test/grankTest.cc:31-36
    unordered_map<int, vector<int>> g;
    dump("empty_grank", ppr::grank(g, 10, 30, 100, 0.85, 0.0001));
    dump("empty_multi", ppr::grankMulti(g, 10, 30, 100, 0.85, 0.0001, 4));
    dump("empty_mc", ppr::mccompletepathv2(g, 10, 30, 100, 0.85));
  } else if (c == "no_edges") {  // test/grankTest.cc:38-50, mccompletepathv2Test.cc:38-50
    unordered_map<int, vector<int>> g;
    for (int i = 0; i < 10; i++) g[i];
    dump("no_edges_grank", ppr::grank(g, 10, 30, 100, 0.85, 0.0001));
    dump("no_edges_mc", ppr::mccompletepathv2(g, 10, 30, 100, 0.85));
  } else if (c == "self_loop") {  // test/grankTest.cc:70-84
    unordered_map<int, vector<int>> g;
    g[0].push_back(0);
    dump("self_loop", ppr::grank(g, 10, 30, 100, 0.85, 0.0001));
  } else if (c == "ring6") {  // test/grankTest.cc:107-152
    dump("ring6_k10_l30", ppr::grank(ring(6), 10, 30, 100, 0.85, 0.0001));
    dump("ring6_k3_l4", ppr::grank(ring(6), 3, 4, 100, 0.85, 0.0001));
    dump("ring6_k3_l3", ppr::grank(ring(6), 3, 3, 100, 0.85, 0.0001));
  } else if (c == "star") {  // test/grankTest.cc:154-182
    unordered_map<int, vector<int>> g;
    g[0];
    for (int i = 1; i < 6; i++) g[i].push_back(0);
    dump("star", ppr::grank(g, 10, 30, 100, 0.85, 0.0001));
    g[0].push_back(0);
    dump("star_selfloop", ppr::grank(g, 10, 30, 100, 0.85, 0.0001));
  } else if (c == "ring100") {  // test/grankTest.cc:184-283
    dump("ring100_k10_l10", ppr::grank(ring(100), 10, 10, 100, 0.85, 0.0001));
    dump("ring100_k10_l20", ppr::grank(ring(100), 10, 20, 100, 0.85, 0.0001));
    dump("ring100_k100", ppr::grank(ring(100), 100, 100, 100, 0.85, -1));
    dump("ring100_k200", ppr::grank(ring(100), 200, 200, 100, 0.85, -1));
  } else if (c == "random_full") {  // test/grankTest.cc:343-361: 100 nodes, 5000 multi-edges, K = L = N (no truncation)
    unordered_map<int, vector<int>> g;
    for (int i = 0; i < 100; i++) g[i];
    unsigned s = 12345;
    for (int i = 0; i < 5000; i++) { int a = lcg(s) % 100, b = lcg(s) % 100; g[a].push_back(b); }
    dump("random_full", ppr::grank(g, 100, 100, 100, 0.85, -1));
  } else if (c == "multi_equals_single") {  // test/grankMultiThreadTest.cc:384-576 incl. positive tolerances
    unordered_map<int, vector<int>> g;
    for (int i = 0; i < 80; i++) g[i];
    unsigned s = 777;
    for (int i = 0; i < 640; i++) { int a = lcg(s) % 80, b = lcg(s) % 80; g[a].push_back(b); }
    const double tols[4] = {0.01, 0.0005, 0.00001, 0.001};
    for (int t = 0; t < 4; t++) {
      auto single = ppr::grank(g, 80, 80, 60, 0.85, tols[t]);
      dump("single", single);
      const size_t threads[2] = {4, 1};
      for (int j = 0; j < 2; j++) {
        auto multi = ppr::grankMulti(g, 80, 80, 60, 0.85, tols[t], threads[j]);
        double worst = 0;
        bool same_keys = multi.size() == single.size();
        for (auto& kv : single) {
          same_keys = same_keys && multi[kv.first].size() == kv.second.size();
          for (auto& e : kv.second) worst = std::max(worst, std::abs(multi[kv.first][e.first] - e.second));
        }
        std::cout << "multi" << threads[j] << " same_keys " << same_keys << " within_1e-4 " << (worst < 1e-4) << "\n";
      }
    }
  } else if (c == "string_keys") {  // README.md:41-66: any hashable Key; keys are copied into the result
    unordered_map<string, vector<string>> g;
    const char* names[5] = {"alpha", "beta", "gamma", "delta", "sink"};
    g["alpha"] = {"beta", "gamma", "gamma"};
    g["beta"] = {"gamma", "alpha"};
    g["gamma"] = {"delta", "sink", "gamma"};
    g["delta"] = {"alpha"};
    g["sink"] = {};
    (void)names;
    dump("string_keys", ppr::grank(g, 5, 5, 80, 0.85, -1));
    dump("string_keys_multi", ppr::grankMulti(g, 5, 5, 80, 0.85, -1, 3));
  } else if (c == "long_keys") {  // non-contiguous 64-bit keys
    unordered_map<long, vector<long>> g;
    unsigned s = 99;
    vector<long> keys;
    for (int i = 0; i < 120; i++) keys.push_back(1000003L * i * i + 17L * i + 5000000000L);
    for (long k : keys) g[k];
    for (int i = 0; i < 700; i++) { long a = keys[lcg(s) % 120], b = keys[lcg(s) % 120]; g[a].push_back(b); }
    dump("long_keys", ppr::grank(g, 120, 120, 40, 0.85, -1));
  } else if (c == "mc_structural") {  // test/mccompletepathv2Test.cc:154-219 (values are walk statistics: check, don't print)
    unordered_map<int, vector<int>> g;
    g[0];
    for (int i = 1; i < 6; i++) g[i].push_back(0);
    auto res = ppr::mccompletepathv2(g, 10, 30, 100, 0.85);
    bool ok = res[0].size() == 1 && res[0][0] == 1.0;
    for (int i = 1; i < 6; i++) ok = ok && res[i].size() == 2 && std::abs(res[i][0] - 0.85) < 10e-5 && res[i][i] == 1.0;
    unordered_map<int, vector<int>> h;
    for (int i = 1; i < 6; i++) { h[i]; h[0].push_back(i); }
    auto res2 = ppr::mccompletepathv2(h, 10, 30, 1000, 0.85);
    for (int i = 1; i < 6; i++) ok = ok && res2[i].size() == 1 && std::abs(res2[0][i] - 0.85 / 5) < 0.05;
    auto res3 = ppr::mccompletepathv2(ring(6), 10, 30, 1000, 0.85);  // :107-152 weak monotonicity
    for (int i = 0; i < 6; i++) {
      ok = ok && res3[i].size() == 6;
      for (int u = 0; u < 5; u++) ok = ok && res3[i][(i + u) % 6] >= res3[i][(i + u + 1) % 6];
    }
    std::cout << "case mc_structural " << (ok ? "OK" : "FAILED") << "\n";
    return ok ? 0 : 2;
  } else {
    std::cerr << "unknown case " << c << "\n";
    return 64;
  }
  return 0;
}

static int run_death(int which) {  // test/grankTest.cc:20-29, grankMultiThreadTest.cc death tests, mccompletepathv2Test.cc:20-29
  unordered_map<int, vector<int>> g;  // the checks fire before the graph is looked at
  switch (which) {
    case 0: ppr::grank(g, 0, 1, 1, 0.85, 0.1); break;            // K must be positive
    case 1: ppr::grank(g, 1, 0, 1, 0.85, 0.1); break;            // L must be positive
    case 2: ppr::grank(g, 2, 1, 1, 0.85, 0.1); break;            // K must be <= L
    case 3: ppr::grank(g, 1, 1, 0, 0.85, 0.1); break;            // iterations must be positive
    case 4: ppr::grank(g, 1, 1, 1, 1.5, 0.1); break;             // damping must be [0,1]
    case 5: ppr::grank(g, 1, 1, 1, -0.1, 0.1); break;
    case 6: ppr::grankMulti(g, 1, 1, 1, 0.85, 0.1, 0); break;    // nThreads must be positive
    case 7: ppr::grankMulti(g, 0, 1, 1, 0.85, 0.1, 2); break;
    case 8: ppr::mccompletepathv2(g, 0, 1, 1, 0.85); break;
    case 9: ppr::mccompletepathv2(g, 1, 1, 0, 0.85); break;
    case 10: ppr::mccompletepathv2(g, 3, 2, 1, 0.85); break;
    case 11: ppr::mccompletepathv2(g, 1, 1, 1, 2.0); break;
    default: return 64;
  }
  std::cout << "survived\n";
  return 0;
}

int main(int argc, char** argv) {
  if (argc == 3 && !strcmp(argv[1], "death")) return run_death(atoi(argv[2]));
  if (argc != 2) { std::cerr << "usage: dropin_check <case> | death <n>\n"; return 64; }
  return run_case(argv[1]);
}
