// demo_main.cc -- the reference's demo program (/root/reference/src/main.cc:30-112) against the drop-in headers: load a
// CSV edge list, run grankMulti / grank / mccompletepathv2 with the demo's parameters, print wall time and the five
// quality statistics of benchmarkAlgorithm. Because it only uses the reference's public API it also builds against the
// reference's own include directories.
//
//   g++ -std=c++11 -O2 -I approximated_personalized_pagerank_b200/cpp/include -I include examples/demo_main.cc
//       -L approximated_personalized_pagerank_b200 -lppr_b200 -Wl,-rpath,$PWD/approximated_personalized_pagerank_b200 -lpthread -o demo
//   ./demo edges.csv
#include <chrono>
#include <fstream>
#include <iostream>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include <grank.h>
#include <grankMulti.h>
#include <mccompletepathv2.h>
#include <benchmarkAlgorithm.h>

typedef std::unordered_map<int, std::vector<int>> Graph;

// "node1,node2" per line; a repeated edge is kept once, every target is a key even without out-edges (main.cc:78-112)
static Graph loadEdgeList(const std::string& fname) {
  Graph graph;
  std::unordered_map<int, std::unordered_set<int>> seen;
  std::ifstream in(fname.c_str());
  if (!in) { std::cerr << "cannot open " << fname << std::endl; std::exit(EXIT_FAILURE); }
  std::string line;
  size_t edges = 0;
  while (std::getline(in, line)) {
    const size_t comma = line.find(',');
    if (comma == std::string::npos) continue;
    const int from = std::stoi(line.substr(0, comma)), to = std::stoi(line.substr(comma + 1));  // stoi stops at '\r'
    graph[to];
    if (seen[from].insert(to).second) { graph[from].push_back(to); edges++; }
  }
  std::cout << "nodes: " << graph.size() << " edges: " << edges << std::endl;
  return graph;
}

template <typename Fn>
static void timed(const char* what, const Graph& graph, Fn run) {
  const auto t0 = std::chrono::steady_clock::now();
  const auto baskets = run();
  const auto t1 = std::chrono::steady_clock::now();
  std::cout << what << " run-time = " << std::chrono::duration_cast<std::chrono::milliseconds>(t1 - t0).count() << " ms" << std::endl;
  const auto stats = ppr::benchmarkAlgorithm(baskets, graph, 200, true);
  std::cout << "-------" << std::endl;
  for (const auto& kv : stats) std::cout << kv.first << "     " << kv.second << std::endl;
  std::cout << "-------" << std::endl;
}

int main(int argc, char** argv) {
  const Graph graph = loadEdgeList(argc > 1 ? argv[1] : "example.txt");
  timed("grank multi", graph, [&]() { return ppr::grankMulti(graph, 50, 100, 30, 0.85, 0.0001, 4); });
  timed("grank", graph, [&]() { return ppr::grank(graph, 50, 100, 30, 0.85, 0.0001); });
  timed("mc", graph, [&]() { return ppr::mccompletepathv2<int>(graph, 50, 200, 1000, 0.85); });
  return 0;
}
