// eval_check.cc -- one user program against the reference's evaluator API (include/benchmarkAlgorithm.h:51-153,
// include/internal/pprSingleSource.h:28-75), two builds: tests/cpp/eval_check_b200 (our headers + libppr_b200.so) and
// oracle/_ref/eval_check_ref (the unmodified reference headers). The cases re-express the reference's own
// test/benchmarkAlgorithmTest.cc:21-160 and test/internal/pprSingleSourceTest.cc:13-193 (cited per case); wherever
// benchmarkAlgorithm samples, testNodes >= the number of candidates, so every node is evaluated and the random shuffle
// drops out: the two builds must print the same statistics (up to arbitrary ties inside the exact top-K).
//
//   eval_check            prints "name = value" lines
//   eval_check death <n>  calls an API with a bad parameter: must print the reference's message and exit(1)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include <grank.h>
#include <mccompletepathv2.h>
#include <benchmarkAlgorithm.h>
#include <internal/pprSingleSource.h>

typedef std::unordered_map<int, std::vector<int>> Graph;
typedef std::unordered_map<int, std::unordered_map<int, double>> Baskets;

static unsigned lcg(unsigned& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

static void show(const char* tag, const std::unordered_map<std::string, double>& st) {
  const char* names[5] = {"jaccard average", "jaccard min", "kendall average", "kendall min", "average map size"};
  for (int i = 0; i < 5; i++) std::printf("%s / %s = %.12f\n", tag, names[i], st.at(names[i]));
}

static Graph complete(int n) {
  Graph g;
  for (int i = 0; i < n; i++)
    for (int u = 0; u < n; u++) g[i].push_back(u);
  return g;
}

static Graph randomGraph(int n, int edges, unsigned seed) {
  Graph g;
  for (int i = 0; i < n; i++) g[i];
  for (int e = 0; e < edges; e++) { const int a = (int)(lcg(seed) % n), b = (int)(lcg(seed) % n); g[a].push_back(b); }
  return g;
}

static int run_death(int which) {
  Graph graph;
  switch (which) {
    case 0: {  // benchmarkAlgorithmTest.cc:21-27
      Baskets gr = ppr::grank(graph, 1, 3, 42, 0.5, 0.0001);
      ppr::benchmarkAlgorithm(gr, graph, 0, false);
      break;
    }
    case 1: {  // benchmarkAlgorithmTest.cc:28-31
      Baskets gr;
      gr[5];
      ppr::benchmarkAlgorithm(gr, graph, 10, false);
      break;
    }
    case 2: ppr::pprInternal::pprSingleSource(graph, 0, 0.85, 0.001, 0); break;   // pprSingleSourceTest.cc:16
    case 3: ppr::pprInternal::pprSingleSource(graph, 1, 1.85, 0.001, 0); break;   // :17
    case 4: ppr::pprInternal::pprSingleSource(graph, 1, -1.85, 0.001, 0); break;  // :18
    case 5: ppr::pprInternal::pprSingleSource(graph, 1, 0.85, 0.001, 0); break;   // :19
    default: return 64;
  }
  return 0;  // not reached when the check fires
}

int main(int argc, char** argv) {
  if (argc == 3 && !std::strcmp(argv[1], "death")) return run_death(std::atoi(argv[2]));
  using ppr::pprInternal::pprSingleSource;
  {  // GRank baskets on a skewed random graph, every non-sink node evaluated
    Graph g;
    unsigned s = 99u;
    const int n = 400;
    for (int i = 0; i < n; i++) g[i];
    for (int e = 0; e < 10 * n; e++) {
      const int a = (int)(lcg(s) % n), b = (int)((unsigned long long)(lcg(s) % n) * (lcg(s) % n) / n);
      if (a % 9 != 0) g[a].push_back(b);
    }
    show("grank random400", ppr::benchmarkAlgorithm(ppr::grank(g, 30, 60, 30, 0.85, 0.0001), g, 100000, true));
  }
  {  // benchmarkAlgorithmTest.cc:33-40 empty map; :42-52 strict leaves nothing to sample
    Graph graph;
    Baskets gr;
    std::printf("empty map = %.1f\n", ppr::benchmarkAlgorithm(gr, graph, 400, false).at("kendall min"));
    graph[0];
    graph[1];
    std::printf("strict no samples = %.1f\n", ppr::benchmarkAlgorithm(ppr::grank(graph, 1, 3, 42, 0.5, 0.0001), graph, 50, true).at("jaccard average"));
  }
  {  // :54-64 no edges: every statistic is 1
    Graph graph;
    for (int i = 0; i < 100; i++) graph[i];
    show("no edges", ppr::benchmarkAlgorithm(ppr::grank(graph, 1, 3, 42, 0.5, 0.0001), graph, 1000, false));
  }
  {  // :66-83 exact PPR against itself on the complete graph; :147-160 half of every basket is foreign
    Graph graph = complete(100);
    Baskets pr;
    for (int i = 0; i < 100; i++) pr[i] = pprSingleSource(graph, 100, 0.85, 0.0001, i);
    show("complete100 exact", ppr::benchmarkAlgorithm(pr, graph, 1000, false));
    Baskets other;  // :104-117 a basket that shares nothing with the exact top
    for (int i = 0; i < 100; i++) other[i][i + 1] = -1.0;
    show("complete100 disjoint", ppr::benchmarkAlgorithm(other, graph, 1000, false));
    for (int i = 0; i < 100; i++)
      for (int u = 1; u <= 100; u++) { graph[-u]; pr[i][-u] = 1.0; }
    const auto half = ppr::benchmarkAlgorithm(pr, graph, 1000, false);
    std::printf("complete100 half / jaccard average = %.12f\ncomplete100 half / jaccard min = %.12f\n", half.at("jaccard average"), half.at("jaccard min"));
  }
  {  // :85-102 and :119-145 random multigraph: exact vs itself, then with every score negated (Kendall -1)
    const Graph graph = randomGraph(100, 4000, 7u);
    Baskets pr;
    for (int i = 0; i < 100; i++) pr[i] = pprSingleSource(graph, 100, 0.85, 0.0001, i);
    show("random100 exact", ppr::benchmarkAlgorithm(pr, graph, 1000, false));
    for (auto& kv : pr)
      for (auto& e : kv.second) e.second *= -1;
    show("random100 negated", ppr::benchmarkAlgorithm(pr, graph, 1000, false));
  }
  {  // pprSingleSourceTest.cc:22-54 single node, no edges, isolated source; :56-110 the origin scores highest
    Graph one;
    one[0];
    std::printf("single node = %.17g\n", pprSingleSource(one, 100, 0.85, 0.001, 0)[0]);
    Graph iso;
    for (int i = 1; i < 5; i++)
      for (int u = 1; u < 5; u++) iso[i].push_back(u);
    iso[0];
    const auto r0 = pprSingleSource(iso, 100, 0.85, 0.001, 0);
    std::printf("isolated source = %.17g size %zu\n", r0.at(0), r0.size());
    const Graph rg = randomGraph(60, 600, 3u);
    int origin_highest = 0;
    for (int i = 0; i < 60; i++) {
      const auto r = pprSingleSource(rg, 100, 0.85, 0.0001, i);
      bool top = true;
      for (const auto& e : r) top = top && (e.first == i || e.second <= r.at(i));
      origin_highest += top;
    }
    std::printf("origin highest = %d of 60\n", origin_highest);
    const Graph ring = [] { Graph g; for (int i = 0; i < 10; i++) g[i].push_back((i + 1) % 10); return g; }();
    const auto rr = pprSingleSource(ring, 100, 0.85, -1.0, 0);  // :112-193 scores decrease along the ring
    for (int i = 0; i < 10; i++) std::printf("ring10[%d] = %.12f\n", i, rr.at(i));
  }
  return 0;
}
