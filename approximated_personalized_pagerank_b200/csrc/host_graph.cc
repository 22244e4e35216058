// host_graph.cc -- host-side (no GPU) pieces of libppr_b200.so:
//   * error string plumbing,
//   * pprb200_find_partitions: the reference's BFS 2-colouring (pprInternal.h:29-99) on dense ids,
//   * the synthetic workload generators of BASELINE.json (R-MAT, Barabasi-Albert).
#include "ppr_internal.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <mutex>
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
  std::vector<int64_t> bad((size_t)host_threads(), -1);  // first offending edge of each range
  host_parallel_for(e, 1 << 16, [&](int t, int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; i++)
      if ((uint32_t)col[i] >= (uint32_t)n) { bad[(size_t)t] = i; break; }
  });
  for (const int64_t i : bad)
    if (i >= 0)
      return fail(PPRB200_ERR_GRAPH, "successor %d at edge %lld is not a node of the graph (every sink must be a key, README.md:68-74)",
                  col[i], (long long)i);
  return PPRB200_OK;
}

// ---- worker pool ------------------------------------------------------------------------------------------------
namespace {
struct HostPool {
  std::mutex m;
  std::condition_variable cv_go, cv_done;
  std::vector<std::thread> workers;
  const std::function<void(int)>* fn = nullptr;
  int parts = 0, next = 0, running = 0;
  unsigned long long gen = 0;
  std::mutex call;  // one parallel region at a time

  explicit HostPool(int n) {
    for (int i = 1; i < n; i++) {
      workers.emplace_back([this]() {
        unsigned long long seen = 0;
        std::unique_lock<std::mutex> lk(m);
        for (;;) {
          cv_go.wait(lk, [&] { return gen != seen; });
          seen = gen;
          while (next < parts) {
            const int part = next++;
            running++;
            lk.unlock();
            (*fn)(part);
            lk.lock();
            running--;
          }
          if (running == 0) cv_done.notify_all();
        }
      });
      workers.back().detach();
    }
  }
  void run(int nparts, const std::function<void(int)>& f) {
    std::lock_guard<std::mutex> serial(call);
    std::unique_lock<std::mutex> lk(m);
    fn = &f;
    parts = nparts;
    next = 0;
    gen++;
    cv_go.notify_all();
    while (next < parts) {
      const int part = next++;
      running++;
      lk.unlock();
      f(part);
      lk.lock();
      running--;
    }
    cv_done.wait(lk, [&] { return running == 0 && next >= parts; });
    parts = 0;
    fn = nullptr;
  }
};

int configured_threads() {
  int t = (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
  if (const char* e = getenv("PPRB200_HOST_THREADS")) t = std::max(1, std::min(64, atoi(e)));
  return t;
}
HostPool& pool() {
  static HostPool* p = new HostPool(configured_threads());  // never destroyed: the workers are detached
  return *p;
}
}  // namespace

int host_threads() {
  static const int t = configured_threads();
  return t;
}

void host_parallel(int parts, const std::function<void(int)>& fn) {
  if (parts <= 1 || host_threads() == 1) {
    for (int i = 0; i < parts; i++) fn(i);
    return;
  }
  pool().run(parts, fn);
}

void host_parallel_for(int64_t n, int64_t grain, const std::function<void(int, int64_t, int64_t)>& fn) {
  if (n <= 0) return;
  int parts = (int)std::min<int64_t>(host_threads(), (n + grain - 1) / grain);
  if (parts <= 1) { fn(0, 0, n); return; }
  host_parallel(parts, [&](int t) { fn(t, n * t / parts, n * (t + 1) / parts); });
}

void host_indegree(const int32_t* col, int64_t e, int32_t n, uint32_t* indeg) {
  std::memset(indeg, 0, sizeof(uint32_t) * (size_t)n);
  const int64_t grain = 1 << 16;
  int parts = (int)std::min<int64_t>(host_threads(), (e + grain - 1) / grain);
  // private histograms as long as they stay under 256 MB in total, else fewer threads
  while (parts > 1 && (int64_t)parts * n * 4 > (256ll << 20)) parts--;
  if (parts <= 1) {
    for (int64_t i = 0; i < e; i++) indeg[(size_t)col[i]]++;
    return;
  }
  std::vector<std::vector<uint32_t>> h((size_t)parts);
  host_parallel(parts, [&](int t) {
    h[(size_t)t].assign((size_t)n, 0u);
    uint32_t* mine = h[(size_t)t].data();
    for (int64_t i = e * t / parts, hi = e * (t + 1) / parts; i < hi; i++) mine[(size_t)col[i]]++;
  });
  host_parallel_for(n, 1 << 14, [&](int, int64_t lo, int64_t hi) {
    for (int64_t v = lo; v < hi; v++) {
      uint32_t sacc = 0;
      for (int t = 0; t < parts; t++) sacc += h[(size_t)t][(size_t)v];
      indeg[(size_t)v] = sacc;
    }
  });
}

// CSR transpose: predecessor lists by ascending source id, multiplicity kept (= the order pprInternal.h:34-43 produces);
// edges whose source has skip_source[u] != 0 are left out (nullptr: none)
void host_transpose(const int64_t* row_ptr, const int32_t* col, int32_t n, std::vector<int64_t>& prow, std::vector<int32_t>& pcol,
                    const uint8_t* skip_source) {
  const int64_t e = row_ptr[n];
  auto skipped = [&](int32_t u) { return skip_source != nullptr && skip_source[(size_t)u] != 0; };
  // predecessor lists (CSR transpose); node ranges balanced by edge count, private cursors per range
  prow.assign((size_t)n + 1, 0);
  {
    int64_t e_eff = e;  // edges that are actually transposed
    if (skip_source != nullptr) {
      std::vector<int64_t> part_sum((size_t)host_threads(), 0);
      host_parallel_for(n, 1 << 16, [&](int t, int64_t lo, int64_t hi) {
        int64_t acc = 0;
        for (int64_t u = lo; u < hi; u++)
          if (!skip_source[(size_t)u]) acc += row_ptr[u + 1] - row_ptr[u];
        part_sum[(size_t)t] = acc;
      });
      e_eff = 0;
      for (const int64_t x : part_sum) e_eff += x;
    }
    // private histograms cost parts * n words: only worth it when there are many more edges than nodes per thread
    pcol.resize((size_t)std::max<int64_t>(e_eff, 1));
    int parts = (int)std::min<int64_t>(host_threads(), (e_eff + (1 << 16) - 1) / (1 << 16));
    if (e_eff < (int64_t)n) parts = 1;
    while (parts > 1 && (int64_t)parts * n * 4 > (256ll << 20)) parts--;
    if (parts <= 1) {
      for (int32_t u = 0; u < n; u++)
        if (!skipped(u))
          for (int64_t i = row_ptr[u]; i < row_ptr[u + 1]; i++) prow[(size_t)col[i] + 1]++;
      for (int32_t v = 0; v < n; v++) prow[(size_t)v + 1] += prow[v];
      std::vector<int64_t> cursor(prow.begin(), prow.end() - 1);
      for (int32_t u = 0; u < n; u++)
        if (!skipped(u))
          for (int64_t i = row_ptr[u]; i < row_ptr[u + 1]; i++) pcol[(size_t)cursor[col[i]]++] = u;
    } else {
      std::vector<int32_t> cut((size_t)parts + 1, 0);  // node ranges with ~e/parts edges each
      for (int t = 1; t < parts; t++)
        cut[(size_t)t] = (int32_t)(std::lower_bound(row_ptr, row_ptr + n + 1, e * t / parts) - row_ptr);
      cut[(size_t)parts] = n;
      for (int t = 1; t <= parts; t++) cut[(size_t)t] = std::max(cut[(size_t)t], cut[(size_t)t - 1]);
      std::vector<std::vector<uint32_t>> h((size_t)parts);
      host_parallel(parts, [&](int t) {
        h[(size_t)t].assign((size_t)n, 0u);
        uint32_t* mine = h[(size_t)t].data();
        for (int32_t u = cut[(size_t)t]; u < cut[(size_t)t + 1]; u++)
          if (!skipped(u))
            for (int64_t i = row_ptr[u]; i < row_ptr[u + 1]; i++) mine[(size_t)col[i]]++;
      });
      host_parallel_for(n, 1 << 14, [&](int, int64_t lo, int64_t hi) {
        for (int64_t v = lo; v < hi; v++) {
          uint32_t sacc = 0;
          for (int t = 0; t < parts; t++) sacc += h[(size_t)t][(size_t)v];
          prow[(size_t)v + 1] = sacc;
        }
      });
      for (int32_t v = 0; v < n; v++) prow[(size_t)v + 1] += prow[v];
      host_parallel_for(n, 1 << 14, [&](int, int64_t lo, int64_t hi) {  // counts -> first write offsets
        for (int64_t v = lo; v < hi; v++) {
          uint32_t run = 0;
          for (int t = 0; t < parts; t++) { const uint32_t c = h[(size_t)t][(size_t)v]; h[(size_t)t][(size_t)v] = run; run += c; }
        }
      });
      host_parallel(parts, [&](int t) {
        uint32_t* mine = h[(size_t)t].data();
        for (int32_t u = cut[(size_t)t]; u < cut[(size_t)t + 1]; u++)
          if (!skipped(u))
            for (int64_t i = row_ptr[u]; i < row_ptr[u + 1]; i++) {
              const int32_t s = col[i];
              pcol[(size_t)(prow[(size_t)s] + mine[(size_t)s]++)] = u;
            }
      });
    }
  }
}

// Reference semantics (pprInternal.h:29-99): roots are taken in map-iteration order (= ascending dense id here) and go
// to `first`; a popped node colours its unvisited successors and predecessors opposite to itself; the queue is FIFO.
// Hence colour(v) = parity of the undirected BFS distance from the root of v's component (the smallest dense id in
// it), whatever the visiting order inside a level: the colouring below is level-synchronous, and the
// levels of large components are expanded by all host threads. oracle/ppr_oracle.c keeps the literal FIFO version;
// tests/test_host_logic.py compares the two.
int find_partitions(const int64_t* row_ptr, const int32_t* col, int32_t n, uint8_t* colour, const ComponentFn* device_component) {
  if (n == 0) return PPRB200_OK;
  const bool timing = getenv("PPRB200_HOST_TIMING") != nullptr;
  auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_begin = now();
  const int64_t e = row_ptr[n];
  const int T = host_threads();
  std::vector<uint8_t> seen((size_t)n, 0);  // phase A: 0 unseen, 1 seen, 2 current frontier, 3 next frontier
  std::vector<int32_t> frontier, next;
  std::vector<std::vector<int32_t>> local((size_t)T);
  int32_t first_root = 0;
  uint8_t other = 1;

  // ---- phase A: the first non-trivial component WITHOUT predecessor lists -------------------------------------------
  // On power-law graphs one component holds nearly every edge, and building the transpose (random scatter of E entries)
  // costs several times more than the BFS itself. In-edges are followed bottom-up instead: an unseen node whose successor
  // is in the current frontier is a predecessor of the frontier, hence in the next level. Each level streams the out-edges
  // of the still-unseen nodes (parallel over node ranges, no atomics: a thread writes only its own nodes). The scanned edges
  // are budgeted at 4 E; a component that needs more levels than that (a long chain) continues in phase B.
  double t_a = t_begin;
  const bool phase_a = e >= (1 << 17) || device_component != nullptr;
  if (phase_a) {
    if (row_ptr[1] == row_ptr[0]) {  // node 0 has no out-edges: in-degrees tell isolated roots from sink roots
      std::vector<uint32_t> indeg((size_t)n);
      host_indegree(col, e, n, indeg.data());
      while (first_root < n && row_ptr[first_root + 1] == row_ptr[first_root] && indeg[(size_t)first_root] == 0) {
        seen[(size_t)first_root] = 1;  // isolated nodes are components of their own (pprInternal.h:58-64: root -> first)
        colour[first_root++] = 0;
      }
    }
    if (first_root < n && device_component && (*device_component)(first_root, seen.data(), colour)) {
      first_root++;  // (the session entry points: this component -- nearly every edge -- was levelled on the device, plan_device.cuh)
    } else if (first_root < n) {
      seen[(size_t)first_root] = 2;
      colour[first_root] = 0;
      frontier.assign(1, first_root);
      long long budget = 4 * (long long)e;
      std::vector<long long> scanned((size_t)T, 0);
      while (!frontier.empty() && budget > 0) {
        const int64_t fs = (int64_t)frontier.size();
        const int fparts = (int)std::max<int64_t>(1, std::min<int64_t>(T, (fs + 511) / 512));
        for (int t = 0; t < T; t++) local[(size_t)t].clear();
        // top-down over the out-edges of the frontier
        host_parallel(fparts, [&](int t) {
          std::vector<int32_t>& out = local[(size_t)t];
          for (int64_t k = fs * t / fparts, hi = fs * (t + 1) / fparts; k < hi; k++) {
            const int32_t x = frontier[(size_t)k];
            for (int64_t i = row_ptr[x]; i < row_ptr[x + 1]; i++) {
              const int32_t sx = col[i];
              if (__atomic_load_n(&seen[(size_t)sx], __ATOMIC_RELAXED) != 0) continue;
              if (__atomic_exchange_n(&seen[(size_t)sx], (uint8_t)3, __ATOMIC_RELAXED) != 0) continue;
              colour[sx] = other;
              out.push_back(sx);
            }
          }
        });
        // bottom-up: unseen nodes with a successor in the current frontier
        const int nparts = (int)std::max<int64_t>(1, std::min<int64_t>(T, ((int64_t)n + 4095) / 4096));
        host_parallel(nparts, [&](int t) {
          std::vector<int32_t>& out = local[(size_t)t];
          long long cnt = 0;
          for (int64_t v = (int64_t)n * t / nparts, hi = (int64_t)n * (t + 1) / nparts; v < hi; v++) {
            if (seen[(size_t)v] != 0) continue;
            for (int64_t i = row_ptr[v]; i < row_ptr[v + 1]; i++) {
              cnt++;
              if (seen[(size_t)col[i]] == 2) {
                seen[(size_t)v] = 3;
                colour[v] = other;
                out.push_back((int32_t)v);
                break;
              }
            }
          }
          scanned[(size_t)t] = cnt;
        });
        for (int t = 0; t < nparts; t++) budget -= scanned[(size_t)t];
        for (const int32_t x : frontier) seen[(size_t)x] = 1;
        next.clear();
        for (int t = 0; t < T; t++) next.insert(next.end(), local[(size_t)t].begin(), local[(size_t)t].end());
        for (const int32_t x : next) seen[(size_t)x] = 2;
        frontier.swap(next);
        other ^= 1u;
      }
      for (const int32_t x : frontier) seen[(size_t)x] = 1;  // (budget ran out mid-component: phase B goes on from this frontier)
      first_root++;
    }
    t_a = now();
  }

  // ---- phase B: whatever is left, with predecessor lists restricted to edges that leave unseen nodes ------------------
  // (an unseen node has only unseen predecessors, or predecessors in the frontier phase A handed over)
  // (an unseen node that has in-edges has them from unseen nodes, which then have out-edges: looking at out-degrees is enough)
  bool work_left = !frontier.empty() || !phase_a;
  for (int32_t v = first_root; v < n && !work_left; v++) work_left = !seen[(size_t)v] && row_ptr[v + 1] > row_ptr[v];
  if (!work_left) {
    for (int32_t v = first_root; v < n; v++)
      if (!seen[(size_t)v]) colour[v] = 0;  // isolated nodes
    if (timing) fprintf(stderr, "[pprb200] find_partitions: transpose-free bfs %.2f ms (%d host threads)\n", now() - t_begin, T);
    return PPRB200_OK;
  }
  // Sparse variant: phase A usually leaves isolated nodes and a handful of tiny components. Their in-edges are then a short
  // list of (target, source) pairs, sorted by target -- no n-sized predecessor arrays, no pass over the seen nodes' edges.
  if (phase_a && frontier.empty()) {
    const int parts = (int)std::max<int64_t>(1, std::min<int64_t>(T, ((int64_t)n + (1 << 16) - 1) >> 16));
    std::vector<std::vector<int32_t>> mine((size_t)parts);
    std::vector<int64_t> edges_of((size_t)parts, 0);
    host_parallel(parts, [&](int t) {
      int64_t cnt = 0;
      for (int64_t v = std::max<int64_t>(first_root, (int64_t)n * t / parts), hi = (int64_t)n * (t + 1) / parts; v < hi; v++)
        if (!seen[(size_t)v] && row_ptr[v + 1] > row_ptr[v]) { mine[(size_t)t].push_back((int32_t)v); cnt += row_ptr[v + 1] - row_ptr[v]; }
      edges_of[(size_t)t] = cnt;
    });
    int64_t e_unseen = 0;
    for (int t = 0; t < parts; t++) e_unseen += edges_of[(size_t)t];
    if (e_unseen * 8 <= e) {
      std::vector<std::pair<int32_t, int32_t>> pred;  // (target, source), both unseen
      pred.reserve((size_t)e_unseen);
      std::vector<uint8_t> has_pred((size_t)n, 0);
      for (int t = 0; t < parts; t++)
        for (const int32_t u : mine[(size_t)t])
          for (int64_t i = row_ptr[u]; i < row_ptr[u + 1]; i++) { pred.emplace_back(col[i], u); has_pred[(size_t)col[i]] = 1; }
      std::sort(pred.begin(), pred.end());
      // isolated nodes are components of their own; what is left are the candidate roots, in ascending id order
      for (int t = 0; t < parts; t++) mine[(size_t)t].clear();
      host_parallel(parts, [&](int t) {
        for (int64_t v = std::max<int64_t>(first_root, (int64_t)n * t / parts), hi = (int64_t)n * (t + 1) / parts; v < hi; v++)
          if (!seen[(size_t)v]) {
            if (row_ptr[v + 1] == row_ptr[v] && !has_pred[(size_t)v]) { seen[(size_t)v] = 1; colour[v] = 0; }
            else mine[(size_t)t].push_back((int32_t)v);
          }
      });
      for (int t = 0; t < parts; t++)
        for (const int32_t root : mine[(size_t)t]) {
          if (seen[(size_t)root]) continue;
          seen[(size_t)root] = 1;
          colour[root] = 0;
          frontier.assign(1, root);
          other = 1;
          while (!frontier.empty()) {
            next.clear();
            for (const int32_t x : frontier) {
              for (int64_t i = row_ptr[x]; i < row_ptr[x + 1]; i++) {
                const int32_t sx = col[i];
                if (!seen[(size_t)sx]) { seen[(size_t)sx] = 1; colour[sx] = other; next.push_back(sx); }
              }
              for (auto it = std::lower_bound(pred.begin(), pred.end(), std::make_pair(x, (int32_t)0)); it != pred.end() && it->first == x; ++it) {
                const int32_t px = it->second;
                if (!seen[(size_t)px]) { seen[(size_t)px] = 1; colour[px] = other; next.push_back(px); }
              }
            }
            frontier.swap(next);
            other ^= 1u;
          }
        }
      if (timing)
        fprintf(stderr, "[pprb200] find_partitions: first component %.2f ms, %lld edges among the rest: sparse bfs %.2f ms (%d host threads)\n",
                t_a - t_begin, (long long)e_unseen, now() - t_a, T);
      return PPRB200_OK;
    }
  }
  std::vector<int64_t> prow;
  std::vector<int32_t> pcol;
  host_transpose(row_ptr, col, n, prow, pcol, phase_a ? seen.data() : nullptr);
  const double t_transposed = now();
  auto run_component = [&]() {  // level-synchronous BFS from `frontier`; the next level gets colour `other`
    while (!frontier.empty()) {
      next.clear();
      if (frontier.size() < 2048 || T == 1) {
        for (const int32_t x : frontier) {
          for (int64_t i = row_ptr[x]; i < row_ptr[x + 1]; i++) {
            const int32_t sx = col[i];
            if (!seen[(size_t)sx]) { seen[(size_t)sx] = 1; colour[sx] = other; next.push_back(sx); }
          }
          for (int64_t i = prow[(size_t)x]; i < prow[(size_t)x + 1]; i++) {
            const int32_t p = pcol[(size_t)i];
            if (!seen[(size_t)p]) { seen[(size_t)p] = 1; colour[p] = other; next.push_back(p); }
          }
        }
      } else {
        const int64_t fs = (int64_t)frontier.size();
        const int parts = (int)std::min<int64_t>(T, (fs + 511) / 512);
        host_parallel(parts, [&](int t) {
          std::vector<int32_t>& out = local[(size_t)t];
          out.clear();
          auto visit = [&](int32_t sx) {
            if (__atomic_load_n(&seen[(size_t)sx], __ATOMIC_RELAXED)) return;
            if (__atomic_exchange_n(&seen[(size_t)sx], (uint8_t)1, __ATOMIC_RELAXED)) return;
            colour[sx] = other;
            out.push_back(sx);
          };
          for (int64_t k = fs * t / parts, hi = fs * (t + 1) / parts; k < hi; k++) {
            const int32_t x = frontier[(size_t)k];
            for (int64_t i = row_ptr[x]; i < row_ptr[x + 1]; i++) visit(col[i]);
            for (int64_t i = prow[(size_t)x]; i < prow[(size_t)x + 1]; i++) visit(pcol[(size_t)i]);
          }
        });
        for (int t = 0; t < parts; t++) next.insert(next.end(), local[(size_t)t].begin(), local[(size_t)t].end());
      }
      frontier.swap(next);
      other ^= 1u;
    }
  };
  run_component();  // the component phase A handed over, if any
  for (int32_t root = first_root; root < n; root++) {
    if (seen[(size_t)root]) continue;
    seen[(size_t)root] = 1;
    colour[root] = 0;
    if (row_ptr[root + 1] == row_ptr[root] && prow[(size_t)root + 1] == prow[(size_t)root]) continue;  // isolated node
    frontier.assign(1, root);
    other = 1;
    run_component();
  }
  if (timing)
    fprintf(stderr, "[pprb200] find_partitions: transpose-free bfs %.2f ms, restricted transpose %.2f ms, bfs %.2f ms (%d host threads)\n",
            t_a - t_begin, t_transposed - t_a, now() - t_transposed, T);
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
