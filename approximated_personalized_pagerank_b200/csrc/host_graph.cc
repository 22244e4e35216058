// host_graph.cc -- host-side (no GPU) pieces of libppr_b200.so:
//   * error string plumbing,
//   * pprb200_find_partitions: the reference's BFS 2-colouring (pprInternal.h:29-99) on dense ids,
//   * the synthetic workload generators of BASELINE.json (R-MAT, Barabasi-Albert).
#include "ppr_internal.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

namespace pprb200 {

static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int validate_csr(const int64_t* row_ptr, const int32_t* col, int32_t n) {
  if (n < 0) return fail(PPRB200_ERR_GRAPH, "n must be >= 0");
  if (n == 0) return PPRB200_OK;
  if (!row_ptr) return fail(PPRB200_ERR_GRAPH, "row_ptr is NULL");
  if (row_ptr[0] != 0) return fail(PPRB200_ERR_GRAPH, "row_ptr[0] must be 0");
  for (int32_t v = 0; v < n; v++)
    if (row_ptr[v + 1] < row_ptr[v]) return fail(PPRB200_ERR_GRAPH, "row_ptr not monotone at node %d", v);
  const int64_t e = row_ptr[n];
  if (e > 0 && !col) return fail(PPRB200_ERR_GRAPH, "col is NULL");
  if ((uint32_t)n > (1u << 30)) return fail(PPRB200_ERR_GRAPH, "n > 2^30 dense ids not supported");
  for (int64_t i = 0; i < e; i++)
    if (col[i] < 0 || col[i] >= n)
      return fail(PPRB200_ERR_GRAPH, "successor %d at edge %lld is not a node of the graph (every sink must be a key, README.md:68-74)",
                  col[i], (long long)i);
  return PPRB200_OK;
}

// Reference semantics (pprInternal.h:29-99): predecessor lists are filled by scanning the map in
// iteration order (= ascending dense id here) -> ascending source id with multiplicity; roots are taken in
// iteration order and go to `first`; a popped node pushes its unvisited successors (vector order), then
// its unvisited predecessors, all coloured opposite to itself. The queue is FIFO.
int find_partitions(const int64_t* row_ptr, const int32_t* col, int32_t n, uint8_t* colour) {
  if (n == 0) return PPRB200_OK;
  const int64_t e = row_ptr[n];
  std::vector<int64_t> prow((size_t)n + 1, 0);
  std::vector<int32_t> pcol((size_t)e);
  for (int64_t i = 0; i < e; i++) prow[(size_t)col[i] + 1]++;
  for (int32_t v = 0; v < n; v++) prow[(size_t)v + 1] += prow[v];
  {
    std::vector<int64_t> cursor(prow.begin(), prow.end() - 1);
    for (int32_t u = 0; u < n; u++)
      for (int64_t i = row_ptr[u]; i < row_ptr[u + 1]; i++) pcol[(size_t)cursor[col[i]]++] = u;
  }
  std::vector<uint8_t> seen((size_t)n, 0);
  std::vector<int32_t> fifo((size_t)n);
  for (int32_t root = 0; root < n; root++) {
    if (seen[root]) continue;
    size_t head = 0, tail = 0;
    seen[root] = 1;
    colour[root] = 0;
    fifo[tail++] = root;
    while (head < tail) {
      const int32_t x = fifo[head++];
      const uint8_t other = colour[x] ^ 1u;
      for (int64_t i = row_ptr[x]; i < row_ptr[x + 1]; i++) {
        const int32_t s = col[i];
        if (!seen[s]) { seen[s] = 1; colour[s] = other; fifo[tail++] = s; }
      }
      for (int64_t i = prow[x]; i < prow[(size_t)x + 1]; i++) {
        const int32_t p = pcol[(size_t)i];
        if (!seen[p]) { seen[p] = 1; colour[p] = other; fifo[tail++] = p; }
      }
    }
  }
  return PPRB200_OK;
}

static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

}  // namespace pprb200

using namespace pprb200;

extern "C" {

const char* pprb200_version(void) { return "pprb200 0.1 (sm_100a)"; }
const char* pprb200_last_error(void) { return g_err; }

int pprb200_find_partitions(const int64_t* row_ptr, const int32_t* col, int32_t n, uint8_t* colour) {
  int rc = validate_csr(row_ptr, col, n);
  if (rc) return rc;
  if (n > 0 && !colour) return fail(PPRB200_ERR_PARAM, "colour is NULL");
  return find_partitions(row_ptr, col, n, colour);
}

// Edge i, level l draws u = splitmix64(seed ^ splitmix64(i*scale + l)) / 2^64 (counter based: any edge can be
// regenerated independently, so the generator is trivially parallel and reproducible in numpy).
int pprb200_gen_rmat(uint32_t scale, uint32_t edge_factor, uint64_t seed, double a, double b, double c,
                     int64_t* row_ptr, int32_t* col) {
  if (scale == 0 || scale > 30 || edge_factor == 0) return fail(PPRB200_ERR_PARAM, "rmat: scale in 1..30, edge_factor > 0");
  if (!row_ptr || !col) return fail(PPRB200_ERR_PARAM, "rmat: NULL output");
  const int64_t n = (int64_t)1 << scale;
  const int64_t e = n * (int64_t)edge_factor;
  std::vector<int32_t> src((size_t)e), dst((size_t)e);
  const double ab = a + b, abc = a + b + c;
  unsigned nt = std::max(1u, std::min(64u, std::thread::hardware_concurrency()));
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < nt; t++)
    pool.emplace_back([&, t]() {
      const int64_t lo = e * t / nt, hi = e * (t + 1) / nt;
      for (int64_t i = lo; i < hi; i++) {
        uint32_t s = 0, d = 0;
        for (uint32_t l = 0; l < scale; l++) {
          const uint64_t r = splitmix64(seed ^ splitmix64((uint64_t)i * scale + l));
          const double u = (double)(r >> 11) * 0x1p-53;
          const uint32_t rb = u >= ab, cb = (u >= a && u < ab) || u >= abc;
          s = (s << 1) | rb;
          d = (d << 1) | cb;
        }
        src[(size_t)i] = (int32_t)s;
        dst[(size_t)i] = (int32_t)d;
      }
    });
  for (auto& th : pool) th.join();
  std::memset(row_ptr, 0, sizeof(int64_t) * (size_t)(n + 1));
  for (int64_t i = 0; i < e; i++) row_ptr[(size_t)src[(size_t)i] + 1]++;
  for (int64_t v = 0; v < n; v++) row_ptr[v + 1] += row_ptr[v];
  std::vector<int64_t> cursor(row_ptr, row_ptr + n);
  for (int64_t i = 0; i < e; i++) col[(size_t)cursor[(size_t)src[(size_t)i]]++] = dst[(size_t)i];
  return PPRB200_OK;
}

// Nodes 0..m form a clique; node v > m attaches to m distinct earlier nodes drawn proportionally to degree
// (uniform draw from the endpoint list, redraw on duplicates). Every undirected edge is stored in both
// directions, adjacency in creation order.
int pprb200_gen_ba(int32_t n, uint32_t m, uint64_t seed, int64_t* row_ptr, int32_t* col, int64_t* n_edges) {
  if (m == 0 || n <= (int32_t)m) return fail(PPRB200_ERR_PARAM, "ba: need n > m > 0");
  const int64_t und = (int64_t)m * (m + 1) / 2 + (int64_t)(n - (int64_t)m - 1) * m;
  if (n_edges) *n_edges = 2 * und;
  if (!col || !row_ptr) return PPRB200_OK;
  std::vector<int32_t> ends;
  ends.reserve((size_t)(2 * und));
  std::vector<int32_t> eu((size_t)und), ev((size_t)und);
  int64_t ne = 0;
  for (uint32_t i = 0; i <= m; i++)
    for (uint32_t j = i + 1; j <= m; j++) {
      eu[(size_t)ne] = (int32_t)i; ev[(size_t)ne] = (int32_t)j; ne++;
      ends.push_back((int32_t)i); ends.push_back((int32_t)j);
    }
  uint64_t ctr = 0;
  std::vector<int32_t> picked(m);
  for (int32_t v = (int32_t)m + 1; v < n; v++) {
    const size_t pool = ends.size();
    for (uint32_t k = 0; k < m; k++) {
      for (;;) {
        const uint64_t r = splitmix64(seed ^ splitmix64(ctr++));
        const int32_t t = ends[(size_t)(((unsigned __int128)r * pool) >> 64)];
        bool dup = false;
        for (uint32_t q = 0; q < k; q++) dup |= picked[q] == t;
        if (!dup) { picked[k] = t; break; }
      }
    }
    for (uint32_t k = 0; k < m; k++) {
      eu[(size_t)ne] = v; ev[(size_t)ne] = picked[k]; ne++;
      ends.push_back(v); ends.push_back(picked[k]);
    }
  }
  std::memset(row_ptr, 0, sizeof(int64_t) * ((size_t)n + 1));
  for (int64_t i = 0; i < ne; i++) { row_ptr[(size_t)eu[(size_t)i] + 1]++; row_ptr[(size_t)ev[(size_t)i] + 1]++; }
  for (int32_t v = 0; v < n; v++) row_ptr[(size_t)v + 1] += row_ptr[v];
  std::vector<int64_t> cursor(row_ptr, row_ptr + n);
  for (int64_t i = 0; i < ne; i++) {
    col[(size_t)cursor[(size_t)eu[(size_t)i]]++] = ev[(size_t)i];
    col[(size_t)cursor[(size_t)ev[(size_t)i]]++] = eu[(size_t)i];
  }
  return PPRB200_OK;
}

}  // extern "C"
