/*
 * oracle/ppr_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into or called by the product path).
 *
 * CPU restatement, on dense int32 ids + CSR, of the reference's two hot paths:
 *   GRank            /root/reference/include/grank.h:42-150 (+ header-only/grankMulti.h:289-436, same arithmetic)
 *   findPartitions   /root/reference/include/internal/pprInternal.h:29-99
 *   keepTop          /root/reference/include/internal/pprInternal.h:109-137
 *   norm1            /root/reference/include/internal/pprInternal.h:147-165
 *   pprSingleSource  /root/reference/include/internal/pprSingleSource.h:28-75
 *   MC walk/combine  /root/reference/include/mccompletepathv2.h:115-165, 211-256 (north-star semantics, see below)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * Canonical conventions (shared bit-for-bit with the CUDA path; DESIGN.md "Canonical semantics"):
 *   - dense id = position of the key in the caller's unordered_map iteration order;
 *   - per key, contributions are accumulated in successor-vector order with fma(val, factor, acc)
 *     (grank.h:107-115 under the reference's -O3 -march=native build contracts to vfmadd);
 *   - keepTop ties (reference: arbitrary, nth_element) are broken (score desc, dense id asc);
 *   - nodes with out-degree > hub_threshold (0 = never) accumulate order-free in 2^-62 fixed point;
 *   - norm1 is summed in 2^-61 fixed point (order-free); maxDiff is the max of those integers.
 *
 * Pinned against the reference itself (oracle/_ref/libppr_ref.so built from the unmodified headers)
 * and the golden vectors under tests/golden/ by tests/test_oracle_vs_reference.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GRANK_HUB_SCALE 0x1p62
#define GRANK_HUB_INV 0x1p-62
#define MC_HUB_SCALE 0x1p59
#define MC_HUB_INV 0x1p-59
#define NORM_SCALE 0x1p61
#define NORM_INV 0x1p-61

typedef struct oracle_stats {
  uint32_t iterations_run;
  uint32_t pad;
  uint64_t node_iterations;         /* sum over executed iterations of |active partition| (grank.h:96) */
  uint64_t nonsink_node_iterations; /* same, nodes with out-degree > 0 only */
  uint64_t edge_reads;              /* successor baskets visited */
  uint64_t merged_entries;          /* basket entries merged (grank.h:114-115 executions) */
  uint64_t candidates;              /* distinct keys before keepTop, summed */
  uint64_t truncations;             /* keepTop calls that dropped something */
  uint64_t boundary_ties;           /* truncations where score[L-1] == score[L] */
  uint64_t algorithmic_bytes;       /* SURVEY.md 8(d) formula, non-sink node-iterations */
  uint64_t walk_steps;              /* MC: executions of the hop (mccompletepathv2.h:149) */
  uint64_t walks;                   /* MC: walks started */
  double max_diff[2];
} oracle_stats;

typedef struct { double s; int32_t id; } cand_t;

/* probe hook: when set, oracle_grank stores every node's distinct-candidate count of its latest update */
static int32_t* g_ncand_out = NULL;
void oracle_set_ncand_out(int32_t* p) { g_ncand_out = p; }

static int cand_cmp(const void* a, const void* b) {
  const cand_t* x = (const cand_t*)a; const cand_t* y = (const cand_t*)b;
  if (x->s > y->s) return -1;
  if (x->s < y->s) return 1;
  return (x->id > y->id) - (x->id < y->id);
}

static int num_threads(int req) {
#ifdef _OPENMP
  if (req <= 0) return omp_get_max_threads();
  return req;
#else
  (void)req; return 1;
#endif
}

/* ---------------------------------------------------------------------------------------------
 * findPartitions (pprInternal.h:29-99) on dense ids. Roots in dense order -> colour 0 ("first");
 * a popped node colours its unvisited successors (vector order) then unvisited predecessors
 * (ascending dense id with multiplicity = the order pprInternal.h:34-43 builds them) the opposite colour.
 * ------------------------------------------------------------------------------------------- */
int oracle_find_partitions(const int64_t* row_ptr, const int32_t* col, int32_t n, uint8_t* colour) {
  if (n == 0) return 0;
  int64_t e = row_ptr[n];
  int64_t* prow = (int64_t*)calloc((size_t)n + 1, sizeof(int64_t));
  int32_t* pcol = (int32_t*)malloc(sizeof(int32_t) * (size_t)(e ? e : 1));
  int64_t* fill = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
  int32_t* queue = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
  uint8_t* visited = (uint8_t*)calloc((size_t)n, 1);
  if (!prow || !pcol || !fill || !queue || !visited) return -1;
  for (int64_t i = 0; i < e; i++) {
    if (col[i] < 0 || col[i] >= n) return -3;
    prow[col[i] + 1]++;
  }
  for (int32_t v = 0; v < n; v++) prow[v + 1] += prow[v];
  for (int32_t v = 0; v < n; v++) fill[v] = prow[v];
  for (int32_t u = 0; u < n; u++)
    for (int64_t i = row_ptr[u]; i < row_ptr[u + 1]; i++) pcol[fill[col[i]]++] = u;
  for (int32_t r = 0; r < n; r++) {
    if (visited[r]) continue;
    int64_t head = 0, tail = 0;
    visited[r] = 1; colour[r] = 0; queue[tail++] = r;
    while (head < tail) {
      int32_t x = queue[head++];
      uint8_t c = (uint8_t)(colour[x] ^ 1);
      for (int64_t i = row_ptr[x]; i < row_ptr[x + 1]; i++) {
        int32_t s = col[i];
        if (!visited[s]) { visited[s] = 1; colour[s] = c; queue[tail++] = s; }
      }
      for (int64_t i = prow[x]; i < prow[x + 1]; i++) {
        int32_t p = pcol[i];
        if (!visited[p]) { visited[p] = 1; colour[p] = c; queue[tail++] = p; }
      }
    }
  }
  free(prow); free(pcol); free(fill); free(queue); free(visited);
  return 0;
}

/* per-thread scratch: dense accumulator with generation stamps */
typedef struct {
  double* acc; int64_t* facc; int32_t* stamp; int32_t* touched; cand_t* cands; int32_t gen; int32_t n;
} scratch_t;

static int scratch_init(scratch_t* s, int32_t n) {
  s->n = n; s->gen = 0;
  s->acc = (double*)malloc(sizeof(double) * (size_t)n);
  s->facc = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
  s->stamp = (int32_t*)calloc((size_t)n, sizeof(int32_t));
  s->touched = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
  s->cands = (cand_t*)malloc(sizeof(cand_t) * (size_t)n);
  return (s->acc && s->facc && s->stamp && s->touched && s->cands) ? 0 : -1;
}
static void scratch_free(scratch_t* s) { free(s->acc); free(s->facc); free(s->stamp); free(s->touched); free(s->cands); }

/* keepTop (pprInternal.h:109-137) with the canonical tie-break; cands sorted (score desc, id asc) on return.
 * returns kept count; *tie set if a boundary tie was cut. */
static int32_t keep_top(cand_t* c, int32_t cnt, uint32_t L, int* truncated, int* tie) {
  qsort(c, (size_t)cnt, sizeof(cand_t), cand_cmp);
  *truncated = 0; *tie = 0;
  if ((uint32_t)cnt > L) {
    *truncated = 1;
    if (L > 0 && c[L - 1].s == c[L].s) *tie = 1;
    return (int32_t)L;
  }
  return cnt;
}

/* norm1 (pprInternal.h:147-165) in 2^-61 fixed point. newb sorted arbitrary; uses scratch stamps. */
static int64_t norm1_fixed(const int32_t* nid, const double* nsc, int32_t ncnt,
                           const int32_t* oid, const double* osc, int32_t ocnt, scratch_t* s) {
  int64_t sum = 0;
  s->gen++;
  for (int32_t i = 0; i < ocnt; i++) { s->stamp[oid[i]] = s->gen; s->acc[oid[i]] = osc[i]; }
  for (int32_t i = 0; i < ncnt; i++) {
    double o = (s->stamp[nid[i]] == s->gen) ? s->acc[nid[i]] : 0.0;
    sum += llrint(fabs(nsc[i] - o) * NORM_SCALE);
  }
  s->gen++;
  for (int32_t i = 0; i < ncnt; i++) s->stamp[nid[i]] = s->gen;
  for (int32_t i = 0; i < ocnt; i++)
    if (s->stamp[oid[i]] != s->gen) sum += llrint(osc[i] * NORM_SCALE);
  return sum;
}

/* ---------------------------------------------------------------------------------------------
 * GRank (grank.h:42-150). colour[v]==0 <=> v in partitions.first. out_* are [n*K], sorted
 * (score desc, id asc), padded with id -1 / score 0.
 * ------------------------------------------------------------------------------------------- */
int oracle_grank(const int64_t* row_ptr, const int32_t* col, int32_t n, const uint8_t* colour,
                 uint32_t K, uint32_t L, uint32_t iterations, double damping, double tolerance,
                 uint32_t hub_threshold, int32_t* out_ids, double* out_scores, uint32_t* out_cnt,
                 oracle_stats* st, int nthreads) {
  if (K == 0 || L == 0 || K > L || iterations == 0 || damping < 0 || damping > 1) return -1; /* grank.h:51-55 */
  oracle_stats local; memset(&local, 0, sizeof(local));
  if (n == 0) { if (st) *st = local; return 0; }
  const double one_minus_d = 1.0 - damping;
  const int T = num_threads(nthreads);
  size_t slots = (size_t)n * L;
  int32_t* cid = (int32_t*)malloc(slots * sizeof(int32_t));
  double* csc = (double*)malloc(slots * sizeof(double));
  int32_t* ccnt = (int32_t*)calloc((size_t)n, sizeof(int32_t));
  int32_t* nid = (int32_t*)malloc(slots * sizeof(int32_t));
  double* nsc = (double*)malloc(slots * sizeof(double));
  int32_t* ncnt = (int32_t*)calloc((size_t)n, sizeof(int32_t));
  scratch_t* scr = (scratch_t*)calloc((size_t)T, sizeof(scratch_t));
  if (!cid || !csc || !ccnt || !nid || !nsc || !ncnt || !scr) return -2;
  for (int t = 0; t < T; t++) if (scratch_init(&scr[t], n)) return -2;

  uint64_t truncs = 0, ties = 0;
  /* init, grank.h:64-83 */
#pragma omp parallel for num_threads(T) schedule(dynamic, 64) reduction(+ : truncs, ties)
  for (int32_t v = 0; v < n; v++) {
#ifdef _OPENMP
    scratch_t* s = &scr[omp_get_thread_num()];
#else
    scratch_t* s = &scr[0];
#endif
    int64_t b = row_ptr[v], e = row_ptr[v + 1];
    double factor = damping / (double)(uint64_t)(e - b);
    s->gen++;
    int32_t nt = 0;
    s->stamp[v] = s->gen; s->acc[v] = one_minus_d; s->touched[nt++] = v; /* grank.h:76 */
    for (int64_t i = b; i < e; i++) {                                     /* grank.h:79-80 */
      int32_t k = col[i];
      if (s->stamp[k] != s->gen) { s->stamp[k] = s->gen; s->acc[k] = 0.0; s->touched[nt++] = k; }
      s->acc[k] += factor;
    }
    for (int32_t i = 0; i < nt; i++) { s->cands[i].id = s->touched[i]; s->cands[i].s = s->acc[s->touched[i]]; }
    int tr, ti;
    int32_t kept = keep_top(s->cands, nt, L, &tr, &ti);                   /* grank.h:82 */
    truncs += (uint64_t)tr; ties += (uint64_t)ti;
    for (int32_t i = 0; i < kept; i++) { cid[(size_t)v * L + i] = s->cands[i].id; csc[(size_t)v * L + i] = s->cands[i].s; }
    ccnt[v] = kept;
  }

  int64_t tolfix_dummy = 0; (void)tolfix_dummy;
  double maxDiff[2] = {tolerance, tolerance};                             /* grank.h:90 */
  uint8_t active = 0;                                                     /* partitions.first is processed first */
  uint32_t it = 0;
  uint64_t node_it = 0, ns_node_it = 0, edge_reads = 0, merged = 0, ncand = 0, abytes = 0;
  for (; it < iterations && (maxDiff[0] > maxDiff[1] ? maxDiff[0] : maxDiff[1]) >= tolerance; it++) { /* grank.h:92 */
    int64_t maxfix = 0;
#pragma omp parallel for num_threads(T) schedule(dynamic, 16) reduction(+ : truncs, ties, node_it, ns_node_it, edge_reads, merged, ncand, abytes) reduction(max : maxfix)
    for (int32_t v = 0; v < n; v++) {
      if (colour[v] != active) continue;                                  /* grank.h:96 */
      node_it++;
      int64_t b = row_ptr[v], e = row_ptr[v + 1];
      if (e == b) continue; /* sink: new map == old map == {v:1-d}, diff 0 */
      ns_node_it++;
#ifdef _OPENMP
      scratch_t* s = &scr[omp_get_thread_num()];
#else
      scratch_t* s = &scr[0];
#endif
      uint64_t deg = (uint64_t)(e - b);
      double factor = damping / (double)deg;                              /* grank.h:105 */
      int hub = hub_threshold != 0 && deg > hub_threshold;
      s->gen++;
      int32_t nt = 0;
      s->stamp[v] = s->gen; s->touched[nt++] = v;                         /* grank.h:101 */
      if (hub) s->facc[v] = llrint(one_minus_d * GRANK_HUB_SCALE); else s->acc[v] = one_minus_d;
      uint64_t m = 0;
      for (int64_t i = b; i < e; i++) {                                   /* grank.h:107 */
        int32_t su = col[i];
        const int32_t* bid = cid + (size_t)su * L; const double* bsc = csc + (size_t)su * L;
        int32_t bc = ccnt[su];
        m += (uint64_t)bc;
        for (int32_t j = 0; j < bc; j++) {                                /* grank.h:114-115 */
          int32_t k = bid[j];
          if (s->stamp[k] != s->gen) { s->stamp[k] = s->gen; s->acc[k] = 0.0; s->facc[k] = 0; s->touched[nt++] = k; }
          if (hub) s->facc[k] += llrint((bsc[j] * factor) * GRANK_HUB_SCALE);
          else s->acc[k] = fma(bsc[j], factor, s->acc[k]);
        }
      }
      for (int32_t i = 0; i < nt; i++) {
        int32_t k = s->touched[i];
        s->cands[i].id = k;
        s->cands[i].s = hub ? (double)s->facc[k] * GRANK_HUB_INV : s->acc[k];
      }
      int tr, ti;
      int32_t kept = keep_top(s->cands, nt, L, &tr, &ti);                 /* grank.h:119 */
      truncs += (uint64_t)tr; ties += (uint64_t)ti;
      int32_t* oi = nid + (size_t)v * L; double* os = nsc + (size_t)v * L;
      for (int32_t i = 0; i < kept; i++) { oi[i] = s->cands[i].id; os[i] = s->cands[i].s; }
      ncnt[v] = kept;
      int64_t df = norm1_fixed(oi, os, kept, cid + (size_t)v * L, csc + (size_t)v * L, ccnt[v], s); /* grank.h:123 */
      if (df > maxfix) maxfix = df;
      edge_reads += deg; merged += m; ncand += (uint64_t)nt;
      if (g_ncand_out) g_ncand_out[v] = nt;
      abytes += 12 * m + 12 * (uint64_t)ccnt[v] + 12 * (uint64_t)kept + 4 + 4 * deg + 16;
    }
    /* grank.h:125-137: results of the active partition become current, the other one is carried */
#pragma omp parallel for num_threads(T) schedule(static)
    for (int32_t v = 0; v < n; v++) {
      if (colour[v] != active || row_ptr[v + 1] == row_ptr[v]) continue;
      memcpy(cid + (size_t)v * L, nid + (size_t)v * L, sizeof(int32_t) * (size_t)ncnt[v]);
      memcpy(csc + (size_t)v * L, nsc + (size_t)v * L, sizeof(double) * (size_t)ncnt[v]);
      ccnt[v] = ncnt[v];
    }
    maxDiff[0] = (double)maxfix * NORM_INV;                               /* grank.h:94,123 */
    active ^= 1;                                                          /* grank.h:129 */
    { double t = maxDiff[0]; maxDiff[0] = maxDiff[1]; maxDiff[1] = t; }   /* grank.h:140 */
  }

  /* final keepTop(K), grank.h:143-147: baskets are already sorted canonically */
  for (int32_t v = 0; v < n; v++) {
    uint32_t c = (uint32_t)ccnt[v];
    if (c > K) { truncs++; if (csc[(size_t)v * L + K - 1] == csc[(size_t)v * L + K]) ties++; c = K; }
    for (uint32_t i = 0; i < K; i++) {
      out_ids[(size_t)v * K + i] = i < c ? cid[(size_t)v * L + i] : -1;
      out_scores[(size_t)v * K + i] = i < c ? csc[(size_t)v * L + i] : 0.0;
    }
    out_cnt[v] = c;
  }
  local.iterations_run = it;
  local.node_iterations = node_it; local.nonsink_node_iterations = ns_node_it;
  local.edge_reads = edge_reads; local.merged_entries = merged; local.candidates = ncand;
  local.truncations = truncs; local.boundary_ties = ties; local.algorithmic_bytes = abytes;
  local.max_diff[0] = maxDiff[0]; local.max_diff[1] = maxDiff[1];
  if (st) *st = local;
  for (int t = 0; t < T; t++) scratch_free(&scr[t]);
  free(scr); free(cid); free(csc); free(ccnt); free(nid); free(nsc); free(ncnt);
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * pprSingleSource (pprSingleSource.h:28-75), dense vectors. The reference accumulates into a hash
 * map in map-iteration order, so its values differ from this one in the last bits only; used as the
 * "exact power-iteration PPR" of the MC L1-error criterion and the GRank == PPR known-answer tests.
 * out[n] receives the score vector (0 for unreached nodes).
 * ------------------------------------------------------------------------------------------- */
int oracle_ppr_single_source(const int64_t* row_ptr, const int32_t* col, int32_t n, uint32_t iterations,
                             double damping, double tolerance, int32_t source, double* out) {
  if (iterations == 0 || damping < 0 || damping > 1 || source < 0 || source >= n) return -1;
  double* cur = (double*)calloc((size_t)n, sizeof(double));
  double* nxt = (double*)calloc((size_t)n, sizeof(double));
  if (!cur || !nxt) return -2;
  cur[source] = 1.0;
  double diff = tolerance;
  for (uint32_t i = 0; i < iterations && diff >= tolerance; i++) {
    memset(nxt, 0, sizeof(double) * (size_t)n);
    nxt[source] = 1.0 - damping;
    for (int32_t u = 0; u < n; u++) {
      if (cur[u] == 0.0) continue;
      int64_t b = row_ptr[u], e = row_ptr[u + 1];
      if (e == b) continue;
      double factor = damping / (double)(uint64_t)(e - b);
      for (int64_t j = b; j < e; j++) nxt[col[j]] += cur[u] * factor;
    }
    diff = 0;
    for (int32_t u = 0; u < n; u++) diff += fabs(cur[u] - nxt[u]);
    double* t = cur; cur = nxt; nxt = t;
  }
  memcpy(out, cur, sizeof(double) * (size_t)n);
  free(cur); free(nxt);
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al., SC'11) -- the counter-based generator of the north-star MC design.
 * key = (source dense id, walk index); counter = (step/2, 0, seed_lo, seed_hi); step parity picks
 * words (0,1) or (2,3): word A chooses the successor (mulhi(A, outdeg)), word B is the teleport coin
 * (continue iff B < floor(damping * 2^32), saturated to 2^32-1).
 * ------------------------------------------------------------------------------------------- */
static inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}

uint32_t oracle_mc_coin_threshold(double damping) {
  double t = floor(damping * 4294967296.0);
  if (t >= 4294967295.0) return 0xFFFFFFFFu;
  if (t <= 0) return 0;
  return (uint32_t)t;
}

#define MC_MAX_STEPS 4096u /* hard cap per walk (P[len>4096] = d^4096; reference loops forever at d=1 on a cycle) */

/* walk phase for one source (mccompletepathv2.h:115-165 with the north-star changes: uniformly random
 * successor from Philox instead of the shared rotating index, exact visit counts instead of the
 * first-come cap). counts accumulate into s->facc (dense), touched list in s->touched. */
static int32_t mc_walk_source(const int64_t* row_ptr, const int32_t* col, int32_t src, uint64_t R, uint64_t W,
                              uint32_t thresh, uint64_t seed, scratch_t* s, uint64_t* steps) {
  s->gen++;
  int32_t nt = 0;
  s->stamp[src] = s->gen; s->facc[src] = (int64_t)R; s->touched[nt++] = src; /* :124 */
  uint64_t st = 0;
  for (uint64_t w = 0; w < W; w++) {                                         /* :134 */
    int32_t cur = src;
    uint32_t rnd[4];
    for (uint32_t step = 0; step < MC_MAX_STEPS; step++) {
      int64_t b = row_ptr[cur], e = row_ptr[cur + 1];
      if (e == b) break;                                                     /* :144-145 */
      if ((step & 1) == 0)
        philox4x32_10(step >> 1, 0, (uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)src, (uint32_t)w, rnd);
      uint32_t a = rnd[(step & 1) * 2], c = rnd[(step & 1) * 2 + 1];
      uint64_t deg = (uint64_t)(e - b);
      cur = col[b + (int64_t)(((uint64_t)a * deg) >> 32)];                   /* :149, random successor */
      if (s->stamp[cur] != s->gen) { s->stamp[cur] = s->gen; s->facc[cur] = 0; s->touched[nt++] = cur; }
      s->facc[cur]++;                                                        /* :152-153 without the cap */
      st++;
      if (!(c < thresh)) break;                                              /* :155 */
    }
  }
  *steps += st;
  return nt;
}

/* MCCompletePathV2, north-star semantics: independent Philox walks from every non-sink
 * node, exact counts -> top-L -> /R; then `rounds` Jacobi rounds of the combine identity
 * (mccompletepathv2.h:211-250) over all nodes; final top-K (:252-256). */
int oracle_mccompletepathv2(const int64_t* row_ptr, const int32_t* col, int32_t n, uint32_t K, uint32_t L,
                            uint32_t R, double damping, uint64_t seed, uint32_t rounds, uint32_t hub_threshold,
                            int32_t* out_ids, double* out_scores, uint32_t* out_cnt, oracle_stats* st, int nthreads) {
  if (K == 0 || L == 0 || K > L || R == 0 || damping < 0 || damping > 1) return -1; /* :190-194 */
  oracle_stats local; memset(&local, 0, sizeof(local));
  if (n == 0) { if (st) *st = local; return 0; }
  const int T = num_threads(nthreads);
  const uint64_t W = (uint64_t)((double)R * damping);                       /* :132 */
  const uint32_t thresh = oracle_mc_coin_threshold(damping);
  size_t slots = (size_t)n * L;
  int32_t* cid = (int32_t*)malloc(slots * sizeof(int32_t));
  double* csc = (double*)malloc(slots * sizeof(double));
  int32_t* ccnt = (int32_t*)calloc((size_t)n, sizeof(int32_t));
  int32_t* nid = (int32_t*)malloc(slots * sizeof(int32_t));
  double* nsc = (double*)malloc(slots * sizeof(double));
  int32_t* ncnt = (int32_t*)calloc((size_t)n, sizeof(int32_t));
  scratch_t* scr = (scratch_t*)calloc((size_t)T, sizeof(scratch_t));
  if (!cid || !csc || !ccnt || !nid || !nsc || !ncnt || !scr) return -2;
  for (int t = 0; t < T; t++) if (scratch_init(&scr[t], n)) return -2;
  uint64_t steps = 0, walks = 0, truncs = 0, ties = 0, merged = 0;
#pragma omp parallel for num_threads(T) schedule(dynamic, 16) reduction(+ : steps, walks, truncs, ties)
  for (int32_t v = 0; v < n; v++) {
#ifdef _OPENMP
    scratch_t* s = &scr[omp_get_thread_num()];
#else
    scratch_t* s = &scr[0];
#endif
    if (row_ptr[v + 1] == row_ptr[v]) {                                      /* :162-163 */
      cid[(size_t)v * L] = v; csc[(size_t)v * L] = 1.0; ccnt[v] = 1; continue;
    }
    uint64_t stp = 0;
    int32_t nt = mc_walk_source(row_ptr, col, v, R, W, thresh, seed, s, &stp);
    steps += stp; walks += W;
    /* top-L on integer counts (count desc, id asc), then /R (:159-160) */
    for (int32_t i = 0; i < nt; i++) { s->cands[i].id = s->touched[i]; s->cands[i].s = (double)s->facc[s->touched[i]]; }
    int tr, ti;
    int32_t kept = keep_top(s->cands, nt, L, &tr, &ti);
    truncs += (uint64_t)tr; ties += (uint64_t)ti;
    for (int32_t i = 0; i < kept; i++) { cid[(size_t)v * L + i] = s->cands[i].id; csc[(size_t)v * L + i] = s->cands[i].s / (double)R; }
    ccnt[v] = kept;
  }
  for (uint32_t r = 0; r < rounds; r++) {
#pragma omp parallel for num_threads(T) schedule(dynamic, 16) reduction(+ : truncs, ties, merged)
    for (int32_t v = 0; v < n; v++) {
#ifdef _OPENMP
      scratch_t* s = &scr[omp_get_thread_num()];
#else
      scratch_t* s = &scr[0];
#endif
      int64_t b = row_ptr[v], e = row_ptr[v + 1];
      if (e == b) { nid[(size_t)v * L] = v; nsc[(size_t)v * L] = 1.0; ncnt[v] = 1; continue; } /* f=1: {v:1} */
      uint64_t deg = (uint64_t)(e - b);
      double factor = damping / (double)deg;                                 /* :214 */
      int hub = hub_threshold != 0 && deg > hub_threshold;
      s->gen++;
      int32_t nt = 0;
      s->stamp[v] = s->gen; s->touched[nt++] = v;
      if (hub) s->facc[v] = llrint(1.0 * MC_HUB_SCALE); else s->acc[v] = 1.0 / factor; /* :226 */
      for (int64_t i = b; i < e; i++) {                                      /* :228 */
        int32_t su = col[i];
        const int32_t* bid = cid + (size_t)su * L; const double* bsc = csc + (size_t)su * L;
        int32_t bc = ccnt[su];
        merged += (uint64_t)bc;
        for (int32_t j = 0; j < bc; j++) {                                   /* :240-241 */
          int32_t k = bid[j];
          if (s->stamp[k] != s->gen) { s->stamp[k] = s->gen; s->acc[k] = 0.0; s->facc[k] = 0; s->touched[nt++] = k; }
          if (hub) s->facc[k] += llrint((bsc[j] * factor) * MC_HUB_SCALE);
          else s->acc[k] += bsc[j];
        }
      }
      for (int32_t i = 0; i < nt; i++) {
        int32_t k = s->touched[i];
        s->cands[i].id = k;
        s->cands[i].s = hub ? (double)s->facc[k] * MC_HUB_INV : s->acc[k];
      }
      int tr, ti;
      int32_t kept = keep_top(s->cands, nt, L, &tr, &ti);                    /* :243 */
      truncs += (uint64_t)tr; ties += (uint64_t)ti;
      for (int32_t i = 0; i < kept; i++) {
        nid[(size_t)v * L + i] = s->cands[i].id;
        nsc[(size_t)v * L + i] = hub ? s->cands[i].s : s->cands[i].s * factor; /* :246-247 */
      }
      ncnt[v] = kept;
    }
    { int32_t* t = cid; cid = nid; nid = t; }
    { double* t = csc; csc = nsc; nsc = t; }
    { int32_t* t = ccnt; ccnt = ncnt; ncnt = t; }
  }
  for (int32_t v = 0; v < n; v++) {                                          /* :252-256 */
    /* after a combine round the basket is sorted by pre-scale score; scaling by f>0 is monotone but can
       merge distinct doubles, so re-sort canonically before cutting to K */
    int32_t c = ccnt[v];
    cand_t* tmp = scr[0].cands;
    for (int32_t i = 0; i < c; i++) { tmp[i].id = cid[(size_t)v * L + i]; tmp[i].s = csc[(size_t)v * L + i]; }
    qsort(tmp, (size_t)c, sizeof(cand_t), cand_cmp);
    if ((uint32_t)c > K) { truncs++; if (tmp[K - 1].s == tmp[K].s) ties++; c = (int32_t)K; }
    for (uint32_t i = 0; i < K; i++) {
      out_ids[(size_t)v * K + i] = (int32_t)i < c ? tmp[i].id : -1;
      out_scores[(size_t)v * K + i] = (int32_t)i < c ? tmp[i].s : 0.0;
    }
    out_cnt[v] = (uint32_t)c;
  }
  local.walk_steps = steps; local.walks = walks; local.truncations = truncs; local.boundary_ties = ties;
  local.merged_entries = merged; local.iterations_run = rounds;
  if (st) *st = local;
  for (int t = 0; t < T; t++) scratch_free(&scr[t]);
  free(scr); free(cid); free(csc); free(ccnt); free(nid); free(nsc); free(ncnt);
  return 0;
}
