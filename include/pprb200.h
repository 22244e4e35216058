/*
 * pprb200.h -- C-ABI of libppr_b200.so, the B200 (sm_100a) implementation of the two approximate
 * all-sources Personalized-PageRank hot paths of fruttasecca/approximated_personalized_pagerank.
 *
 * This is the drop-in boundary. The reference has no FFI of its own: its public API is
 * three C++ function templates over std::unordered_map. The template headers shipped in
 * approximated_personalized_pagerank_b200/cpp/include/{grank.h,grankMulti.h,mccompletepathv2.h} keep those
 * signatures, relabel the caller's map to dense int32 ids (dense id = position in the map's iteration
 * order) + CSR, and call the entry points below; INTEGRATION.md shows the binding.
 *
 *   pprb200_grank              replaces ppr::grank            /root/reference/include/grank.h:42-150
 *                              and      ppr::grankMulti       /root/reference/header-only/grankMulti.h:289-436
 *   pprb200_mccompletepathv2   replaces ppr::mccompletepathv2 /root/reference/include/mccompletepathv2.h:182-258
 *   pprb200_find_partitions    replaces pprInternal::findPartitions /root/reference/include/internal/pprInternal.h:29-99
 *
 * Conventions: plain pointers and sizes only; the caller owns every host buffer; the library owns all
 * device memory. Every function returns PPRB200_OK (0) or a negative error code, and
 * pprb200_last_error() then describes the failure (thread-local string). There is no CPU fallback:
 * compute entry points fail with PPRB200_ERR_CUDA when no sm_100 device is usable.
 *
 * Graph input: CSR over dense ids 0..n-1; row_ptr has n+1 entries; col keeps the successor-vector order
 * and multiplicity of the caller's graph (multi-edges and self-loops are legal and counted, reference
 * test/grankTest.cc:60,79). Sinks are rows of length 0.
 *
 * Result layout ("baskets"): out_ids[v*K + i], out_scores[v*K + i] for i < out_cnt[v], sorted by
 * (score descending, dense id ascending); unused slots hold id -1 / score 0.
 */
#ifndef PPRB200_H
#define PPRB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PPRB200_OK 0
#define PPRB200_ERR_PARAM (-1)  /* K==0, L==0, K>L, iterations==0, damping outside [0,1] (grank.h:51-55) */
#define PPRB200_ERR_ALLOC (-2)
#define PPRB200_ERR_GRAPH (-3)  /* malformed CSR / successor outside 0..n-1 (reference: UB, pprInternal.h:76) */
#define PPRB200_ERR_CUDA (-4)   /* no device, launch or runtime failure */
#define PPRB200_ERR_STATE (-5)  /* session used in the wrong order */

/* default out-degree above which a node is accumulated order-free in fixed point */
#define PPRB200_DEFAULT_HUB_THRESHOLD 12u
#define PPRB200_DEFAULT_MC_ROUNDS 3u
#define PPRB200_DEFAULT_MC_SEED 0x5eed5eed5eed5eedull

typedef struct pprb200_stats {
  uint32_t iterations_run;          /* GRank iterations executed / MC combine rounds executed */
  uint32_t n_gpus;
  uint64_t node_iterations;         /* sum over executed iterations of |active partition| (grank.h:96), sinks included */
  uint64_t nonsink_node_iterations; /* same, out-degree > 0 only (the nodes the kernels touch) */
  uint64_t edge_reads;              /* successor baskets merged */
  uint64_t merged_entries;          /* basket entries merged (grank.h:114-115 executions) */
  uint64_t candidates;              /* distinct keys before keepTop, summed over node-iterations */
  uint64_t truncations;             /* keepTop calls that dropped entries */
  uint64_t boundary_ties;           /* of those, cuts through a run of equal scores */
  uint64_t algorithmic_bytes;       /* 12*sum|B_s| + 12|B_v| + 12|B'_v| + 4 + 4*outdeg + 16 per non-sink node-iteration */
  uint64_t walk_steps;              /* MC: hops executed (mccompletepathv2.h:149) */
  uint64_t walks;                   /* MC: walks started */
  uint64_t overflow_requeues;       /* node-iterations that had to be retried with a larger hash table */
  uint64_t walk_algorithmic_bytes;  /* MC walk phase, SURVEY.md 8(d): 12 B per hop + 12*|basket| + 4 per source */
  double max_diff[2];               /* final maxDiff pair (grank.h:90,140) */
  double prep_ms;                   /* host: partition + storage order + CSR encode */
  double h2d_ms;
  double kernel_ms;                 /* device time of init + iterations + final top-K (CUDA events) */
  double d2h_ms;
  double total_ms;
} pprb200_stats;

const char* pprb200_version(void);
const char* pprb200_last_error(void);

/* number of usable sm_100 devices (0 when none; never fails) */
int pprb200_device_count(void);

/* ---- host-side preprocessing (no GPU needed) ------------------------------------------------------ */

/* BFS 2-colouring exactly as pprInternal.h:29-99 produces it when the map is iterated in dense order:
 * colour[v] = 0 for partitions.first, 1 for partitions.second. */
int pprb200_find_partitions(const int64_t* row_ptr, const int32_t* col, int32_t n, uint8_t* colour);
/* The same colouring the way the session / one-shot entry points compute it on graphs of a million edges and more: the
 * first non-trivial component (on power-law graphs nearly everything) is levelled by a BFS on the device over the uploaded
 * CSR (csrc/plan_device.cuh), the rest on the host. Identical output; needs an sm_100 device. */
int pprb200_find_partitions_device(const int64_t* row_ptr, const int32_t* col, int32_t n, uint8_t* colour);

/* ---- one-shot entry points with HOST buffers (what the template headers call) ---------------------- */

/* colour may be NULL (computed with pprb200_find_partitions). hub_threshold 0 = library default,
 * UINT32_MAX = never use the fixed-point hub path. stats may be NULL. */
int pprb200_grank(const int64_t* row_ptr, const int32_t* col, int32_t n, const uint8_t* colour,
                  uint32_t K, uint32_t L, uint32_t iterations, double damping, double tolerance,
                  uint32_t hub_threshold,
                  int32_t* out_ids, double* out_scores, uint32_t* out_cnt, pprb200_stats* stats);

/* R = the reference's `iterations` argument (walks per node in the worst case); rounds = Jacobi combine
 * rounds (0 = pure Monte-Carlo baskets). */
int pprb200_mccompletepathv2(const int64_t* row_ptr, const int32_t* col, int32_t n,
                             uint32_t K, uint32_t L, uint32_t R, double damping, uint64_t seed,
                             uint32_t rounds, uint32_t hub_threshold,
                             int32_t* out_ids, double* out_scores, uint32_t* out_cnt, pprb200_stats* stats);

/* ---- session API: graph and baskets stay resident in HBM (bench `value`, multi-GPU sharding) -------- */

typedef struct pprb200_session pprb200_session;

/* Preprocess + upload. rank/world shard the source nodes (world=1: everything). stream is a
 * cudaStream_t passed as void* (NULL = legacy default stream); all later work of the session is
 * enqueued on it. max_L bounds the L of later runs (sizes the basket arrays). */
int pprb200_session_create(const int64_t* row_ptr, const int32_t* col, int32_t n, const uint8_t* colour,
                           uint32_t max_L, uint32_t hub_threshold, int32_t rank, int32_t world, void* stream,
                           pprb200_session** out);
void pprb200_session_destroy(pprb200_session* s);

/* Enqueue a whole GRank run (init + iterations with the device-side convergence flag + final top-K);
 * asynchronous with respect to the host. */
int pprb200_session_grank(pprb200_session* s, uint32_t K, uint32_t L, uint32_t iterations, double damping,
                          double tolerance);
/* Enqueue a whole MCCompletePathV2 run (walks + combine rounds + final top-K). */
int pprb200_session_mc(pprb200_session* s, uint32_t K, uint32_t L, uint32_t R, double damping, uint64_t seed,
                       uint32_t rounds);
/* Synchronise and copy the last run's baskets to host buffers ([n*K], [n*K], [n]); any may be NULL. */
int pprb200_session_fetch(pprb200_session* s, int32_t* out_ids, double* out_scores, uint32_t* out_cnt);
/* Synchronise and read the device-side counters of the last run. */
int pprb200_session_stats(pprb200_session* s, pprb200_stats* stats);
/* Device time of the dominant kernel family of the last run, measured with CUDA events on the session
 * stream: which = 0 merge (GRank iterations / MC combine), 1 MC walks. Returns launches and total ms. */
int pprb200_session_kernel_time(pprb200_session* s, int which, uint32_t* launches, double* total_ms);

/* ---- multi-GPU: one process (rank) per GPU, sources sharded, baskets pushed to the peers over NVLink ---------
 * Every rank creates a session with its (rank, world) on its own device, exports PPRB200_IPC_BYTES of CUDA IPC
 * handles, the host plumbing all-gathers them (torch.distributed / MPI / a file -- not this library's business),
 * and every rank attaches the world*PPRB200_IPC_BYTES blob (rank-major). All ranks must then enqueue the same
 * runs in the same order; the kernels exchange baskets and the convergence flag themselves. The caller keeps
 * every session alive until all ranks have synchronised their last run (peers write into this rank's memory). */
#define PPRB200_IPC_BYTES 192
#define PPRB200_MAX_WORLD 8
int pprb200_session_ipc_export(pprb200_session* s, void* out /* PPRB200_IPC_BYTES */);
int pprb200_session_ipc_attach(pprb200_session* s, const void* all_handles /* world * PPRB200_IPC_BYTES */);
/* The same wiring for the ranks of one job that live in ONE process (a session per device): plain peer pointers, no IPC.
 * all[r] = the unattached session of rank r, r = 0..world-1. (The one-shot entry points do this themselves: PPR_NUM_GPUS.) */
int pprb200_session_attach_local(pprb200_session** all, int32_t world);
/* owner[v] = rank that updates node v in a world-rank session (-1 for sinks: nobody). Host only. colour may be
 * NULL (MC sessions: one class). */
int pprb200_shard_owner(const int64_t* row_ptr, const int32_t* col, int32_t n, const uint8_t* colour, uint32_t hub_threshold,
                        int32_t world, int32_t* owner);

/* Device memory freed by destroyed sessions stays cached in the stream-ordered pool for the next call; this returns it to the
 * driver (e.g. before another process needs the GPU). */
int pprb200_release_cached_memory(void);

/* Kernels the last run enqueued on the session stream (bench.py's gpu_launches). */
int pprb200_session_launches(pprb200_session* s, uint64_t* launches);

/* Debug / CPU tests: the host front half of a session alone -- colouring (colour == NULL), storage order, rank labels, CSR
 * encode, work items of rank `rank` of `world` -- without touching a device. pos_of[n] (storage position, -1 for sinks),
 * rank_of[n], row_off[n+1] (M+1 used), enc[E] (column words), item_*[item_cap]; summary[16] = M, n_items, chunk, mid_deg,
 * range_begin[2], range_end[2], item_begin[2][2], item_end[2][2]. Every output pointer may be NULL. */
int pprb200_debug_host_plan(const int64_t* row_ptr, const int32_t* col, int32_t n, const uint8_t* colour, uint32_t hub_threshold,
                            int32_t rank, int32_t world, int32_t* pos_of, int32_t* rank_of, int64_t* row_off, uint32_t* enc,
                            int32_t* item_pos, int64_t* item_off, int32_t* item_len, int32_t item_cap, int32_t* summary);

/* Debug / CPU tests: the multi-GPU push plan of `world` ranks, by NODE (host only). owner[v] = rank that updates v (-1: sink);
 * need[v] bit r = rank r owns a predecessor of v, i.e. reads v's basket during the iterations -- the ranks v's owner stores
 * it into (everybody else receives the final basket once, before the top-K). 0 for sinks: they have no basket. */
int pprb200_debug_need_mask(const int64_t* row_ptr, const int32_t* col, int32_t n, const uint8_t* colour, uint32_t hub_threshold,
                            int32_t world, int32_t* owner, uint8_t* need);

/* Debug: per-CTA phase cycle counters of the order-free merge kernels (sessions created with PPRB200_PROF=1 in the
 * environment; out[2][8 * sm_count][8], cleared by the call) and merge_dense_kernel's bookkeeping of the last run
 * (out[8]: nodes finished, nodes that ran pass 2, nodes without a lower bound of the cut, successors of the nodes handed to
 * merge_par_kernel after reading them, nodes selected straight from the tables (candidates > CMAX), tail table full, of those
 * finished by pass-2 rounds, nodes handed over unread because their old basket was below PPRB200_MIN_OLD entries). */
int pprb200_debug_prof(pprb200_session* s, unsigned long long* out, int* n_ctas);
int pprb200_debug_counters(pprb200_session* s, unsigned long long* out);
/* Debug / known-answer tests: out[i][0..3] = Philox4x32-10 with counter ctr_key[i][0..3] and key ctr_key[i][4..5], computed on the
 * device by the generator the walk kernel uses (csrc/mc_walk.cuh; replaces the reference's random_device-seeded mt19937,
 * /root/reference/include/mccompletepathv2.h:32-34). */
int pprb200_debug_philox(const uint32_t* ctr_key, uint32_t* out, int32_t nblocks);

/* Quality-evaluator yardstick: exact Personalized PageRank by power iteration for a batch of
 * sources -- replaces ppr::pprInternal::pprSingleSource (/root/reference/include/internal/pprSingleSource.h:28-75)
 * as include/benchmarkAlgorithm.h:91 calls it, once per sampled node. out_scores[i*n + v] = score of node v for
 * sources[i] (0 for unreached nodes); out_iterations[i] (nullable) = iterations that source ran (it stops on its own
 * once its norm-1 change drops below `tolerance`; negative = never); kernel_ms (nullable) = device time. */
int pprb200_ppr_exact(const int64_t* row_ptr, const int32_t* col, int32_t n, const int32_t* sources, uint32_t n_sources,
                      uint32_t iterations, double damping, double tolerance, double* out_scores, uint32_t* out_iterations,
                      double* kernel_ms);


/* ---- synthetic workloads of BASELINE.json (host only; used by bench.py and the tests) --------------- */

/* R-MAT (a,b,c,1-a-b-c), 2^scale nodes, edge_factor*2^scale directed edges, duplicates and self-loops
 * kept, no vertex permutation, adjacency in edge-generation order. row_ptr[2^scale+1], col[E]. */
int pprb200_gen_rmat(uint32_t scale, uint32_t edge_factor, uint64_t seed, double a, double b, double c,
                     int64_t* row_ptr, int32_t* col);
/* Barabasi-Albert, m attachments per new node, symmetrised. Call with col == NULL to get the directed
 * edge count in *n_edges, then again with buffers. */
int pprb200_gen_ba(int32_t n, uint32_t m, uint64_t seed, int64_t* row_ptr, int32_t* col, int64_t* n_edges);

#ifdef __cplusplus
}
#endif
#endif /* PPRB200_H */
