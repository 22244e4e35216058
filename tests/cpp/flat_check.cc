// flat_check.cc -- the flat result view of the drop-in front-end (ppr::b200::grankFlat / mccompletepathv2Flat,
// SURVEY.md 8-f1) must hold exactly what the map API returns. Built against OUR headers only (the reference has no
// such view); needs a GPU. Prints "flat_check OK" and exits 0.
#include <cstdio>
#include <string>
#include <unordered_map>
#include <vector>

#include <grank.h>
#include <mccompletepathv2.h>

static unsigned lcg(unsigned& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <typename Key>
static bool same(const ppr::b200::FlatBaskets<Key>& flat, const std::unordered_map<Key, std::unordered_map<Key, double>>& maps) {
  if (flat.size() != maps.size()) return false;
  for (size_t v = 0; v < flat.size(); v++) {
    const auto it = maps.find(flat.key(v));
    if (it == maps.end() || it->second.size() != flat.cnt[v]) return false;
    for (uint32_t i = 0; i < flat.cnt[v]; i++) {
      const auto e = it->second.find(flat.key((size_t)flat.ids[v * flat.K + i]));
      if (e == it->second.end() || e->second != flat.scores[v * flat.K + i]) return false;
      if (i && flat.scores[v * flat.K + i] > flat.scores[v * flat.K + i - 1]) return false;  // score descending
    }
  }
  return flat.toMaps() == maps;
}

int main() {
  std::unordered_map<std::string, std::vector<std::string>> g;
  unsigned s = 12345u;
  const int n = 3000;
  for (int i = 0; i < n; i++) g["node" + std::to_string(i)];
  for (int e = 0; e < 12 * n; e++) {
    const int a = (int)(lcg(s) % n), b = (int)((lcg(s) % n) * (unsigned long long)(lcg(s) % n) / n);  // skewed targets
    if (a % 7 != 0) g["node" + std::to_string(a)].push_back("node" + std::to_string(b));
  }
  bool ok = same(ppr::b200::grankFlat(g, 20, 60, 12, 0.85, 1e-4), ppr::grank(g, 20, 60, 12, 0.85, 1e-4));
  ok = same(ppr::b200::mccompletepathv2Flat(g, 10, 40, 200, 0.85), ppr::mccompletepathv2(g, 10, 40, 200, 0.85)) && ok;
  std::unordered_map<int, std::vector<int>> empty;
  ok = ppr::b200::grankFlat(empty, 1, 1, 1, 0.5, 0.1).size() == 0 && ok;
  std::printf("flat_check %s\n", ok ? "OK" : "FAILED");
  return ok ? 0 : 1;
}
