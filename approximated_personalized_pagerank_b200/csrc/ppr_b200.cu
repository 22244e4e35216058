// ppr_b200.cu -- session management, kernel orchestration and the C-ABI of libppr_b200.so (include/pprb200.h).
//
// One GRank run (include/grank.h:42-150) is enqueued as
//   init cascade (both colours)  ->  [ merge cascade(colour it&1) -> iter_end ] x iterations  ->  final top-K
// with no host synchronisation in between: the convergence test of grank.h:92 runs on the device
// (iter_end clears RunState::active, later launches return immediately).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <numeric>
#include <vector>

#include "mc_walk.cuh"
#include "ppr_exact.cuh"
#include "merge_par.cuh"
#include "merge_dense.cuh"
#include "plan_device.cuh"
#include "ppr_internal.h"

namespace pprb200 {

#define CUDA_TRY(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return fail(PPRB200_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ------------------------------------------------------------------------------------------------
// small control kernels
// ------------------------------------------------------------------------------------------------
__global__ void state_reset_kernel(RunState* st) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const unsigned long long seq = st->barrier_seq;  // survives across runs: all ranks count executed barriers in lockstep
    memset(st, 0, sizeof(RunState));
    st->barrier_seq = seq;
    st->active = 1;
    st->m_prev = -1;
    st->m_last = -1;
  }
}

// empty pool tables: the GSlot array at the start of each table gets key -1 / accumulator 0
__global__ void pool_init_kernel(unsigned char* pool, size_t tbl_bytes, unsigned int cap, int n_tables) {
  const size_t total = (size_t)cap * (size_t)n_tables;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    GSlot* g = reinterpret_cast<GSlot*>(pool + (i / cap) * tbl_bytes) + (i % cap);
    g->key = KEY_EMPTY;
    g->pad = 0;
    g->acc = 0ull;
  }
}

// debug: the device-side Philox4x32-10 on caller-chosen (counter, key) blocks (known-answer tests)
__global__ void philox_probe_kernel(const uint32_t* ctr_key, uint32_t* out, int nblocks) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nblocks) return;
  const uint32_t* c = ctr_key + (size_t)i * 6;
  uint32_t r[4];
  philox4x32_10(c[0], c[1], c[2], c[3], c[4], c[5], r);
  for (int j = 0; j < 4; j++) out[(size_t)i * 4 + j] = r[j];
}

// end of an init / combine phase: clear cascade counters (and optionally the work statistics)
__global__ void phase_end_kernel(RunState* st, int clear_stats, int flip_both, PeerDev peers, int barrier) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    // multi-GPU: every rank's pushes of this phase have landed before anybody reads them (or overwrites their source)
    if (barrier && peers.world > 1 && !st->peer_timeout) cross_gpu_barrier(peers, ++st->barrier_seq, 0, &st->peer_timeout);
    if (st->peer_timeout) st->active = 0;  // a dead peer: the remaining launches return at once, the host reports the error
    for (int i = 0; i < 12; i++) st->work[i] = 0;
    for (int i = 0; i < 8; i++) st->qcount[i] = 0;
    if (clear_stats) {
      st->node_iters = st->edge_reads = st->merged = st->cands = st->abytes = st->requeues = 0;
    }
    if (flip_both) { st->slot[0] ^= 1; st->slot[1] ^= 1; st->iter++; }
  }
}

// Multi-GPU, after the last iteration: during the run a basket only went to the ranks that read it (PeerDev::need); now
// every rank that was left out receives the final one, so that each rank holds the complete result for its final top-K.
// One warp per position this rank owns.
__global__ void __launch_bounds__(256) final_push_kernel(const RunState* st, PeerDev pd, const unsigned char* owner, int M, int pos_split, int Lp) {
  const int lane = threadIdx.x & 31;
  const long long stride = ((long long)gridDim.x * blockDim.x) >> 5;
  const size_t sb = slot_bytes(Lp);
  const int n16 = (int)(sb >> 4);
  const unsigned int all = (1u << pd.world) - 1u;
  for (long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < M; p += stride) {
    if (owner[p] != (unsigned char)pd.rank) continue;
    const unsigned int missing = all & ~((unsigned int)pd.need[p] | (1u << pd.rank));
    if (!missing) continue;
    const int b = st->slot[p >= pos_split ? 1 : 0];
    const int4* src = reinterpret_cast<const int4*>(pd.buf[pd.rank][b] + (size_t)p * sb);
    for (int i = lane; i < n16; i += 32) {
      const int4 v = __ldcg(src + i);
      for (int r = 0; r < pd.world; r++)
        if ((missing >> r) & 1u) __stcg(reinterpret_cast<int4*>(pd.buf[r][b] + (size_t)p * sb) + i, v);
    }
  }
}

// end of GRank iteration `it` (0-based) on `colour`: grank.h:129-140 + the loop test of :92
__global__ void iter_end_kernel(RunState* st, int colour, double tolerance, PeerDev peers) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    for (int i = 0; i < 12; i++) st->work[i] = 0;
    for (int i = 0; i < 8; i++) st->qcount[i] = 0;
    if (!st->active) return;
    // multi-GPU: barrier + allreduce(max) of the iteration's max-diff in one mailbox round; every rank takes the
    // same decision below, so converged runs skip the remaining barriers consistently
    if (peers.world > 1) st->cur_max = cross_gpu_barrier(peers, ++st->barrier_seq, st->cur_max, &st->peer_timeout);
    if (st->peer_timeout) { st->active = 0; return; }
    st->slot[colour] ^= 1;
    st->iter++;
    st->m_prev = st->m_last;
    st->m_last = st->cur_max;
    st->cur_max = 0;
    // maxDiff = {tolerance, tolerance} initially (grank.h:90): the test can only fail once both slots hold
    // measured values, i.e. from the second executed iteration on.
    if (st->iter >= 2) {
      const long long m = st->m_prev > st->m_last ? st->m_prev : st->m_last;
      if (!((double)m * NORM_INV >= tolerance)) st->active = 0;
    } else {
      const double m1 = (double)st->m_last * NORM_INV;
      const double mx = m1 > tolerance ? m1 : tolerance;
      if (!(mx >= tolerance)) st->active = 0;
    }
  }
}

// final keepTop(K) (grank.h:143-147, mccompletepathv2.h:252-256): one warp per dense id; output sorted
// (score desc, id asc) by rank counting over the <= L basket entries.
__global__ void final_topk_kernel(const int* __restrict__ pos_of, const int* __restrict__ dense_of,
                                  const unsigned char* __restrict__ colour,
                                  unsigned char* buf0, unsigned char* buf1, const RunState* st, int n, int Lp, int K,
                                  double sink_score, int* __restrict__ out_ids, double* __restrict__ out_scores,
                                  unsigned int* __restrict__ out_cnt, unsigned long long* trunc_ties) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = lane_id();
  const int w = threadIdx.x >> 5;
  const int warps = blockDim.x >> 5;
  double* s_sc = reinterpret_cast<double*>(smem) + (size_t)w * Lp;
  int* s_id = reinterpret_cast<int*>(smem + (size_t)warps * Lp * 8) + (size_t)w * Lp;
  unsigned long long truncs = 0, ties = 0;
  for (int v = blockIdx.x * warps + w; v < n; v += gridDim.x * warps) {
    int* oi = out_ids + (size_t)v * K;
    double* os = out_scores + (size_t)v * K;
    const int p = pos_of[v];
    if (p < 0) {  // sink
      for (int i = lane; i < K; i += 32) { oi[i] = i == 0 ? v : KEY_EMPTY; os[i] = i == 0 ? sink_score : 0.0; }
      if (lane == 0) out_cnt[v] = 1;
      continue;
    }
    const unsigned char* slot = (st->slot[colour[v]] ? buf1 : buf0) + (size_t)p * slot_bytes(Lp);
    const int* ids = reinterpret_cast<const int*>(slot);
    const double* sc = reinterpret_cast<const double*>(slot + (size_t)Lp * 4);
    int cnt = 0;
    for (int e = lane; e < Lp; e += 32) {
      const int id = ids[e];
      s_id[e] = id >= 0 ? dense_of[id] : -1;
      s_sc[e] = id >= 0 ? sc[score_index(e, Lp)] : 0.0;
      cnt += id >= 0;
    }
    cnt = warp_sum_int(cnt);
    __syncwarp();
    const int kept = cnt < K ? cnt : K;
    bool tie = false;
    for (int e = lane; e < cnt; e += 32) {
      const double x = s_sc[e];
      const int id = s_id[e];
      int rank = 0;
      for (int j = 0; j < cnt; j++) {
        const double y = s_sc[j];
        rank += (y > x) || (y == x && s_id[j] < id);
      }
      if (rank < K) { oi[rank] = id; os[rank] = x; }
      if (rank == K - 1 && cnt > K) {
        // boundary tie <=> some other entry with the same score ranks K
        for (int j = 0; j < cnt; j++) tie |= (s_sc[j] == x && s_id[j] > id);
      }
    }
    for (int i = kept + lane; i < K; i += 32) { oi[i] = KEY_EMPTY; os[i] = 0.0; }
    if (lane == 0) out_cnt[v] = (unsigned int)kept;
    tie = __any_sync(FULL, tie);
    if (cnt > K) { truncs += (lane == 0); ties += (lane == 0 && tie); }
    __syncwarp();
  }
  if (lane == 0 && truncs) { atomicAdd(&trunc_ties[0], truncs); atomicAdd(&trunc_ties[1], ties); }
}

}  // namespace pprb200

using namespace pprb200;

// ------------------------------------------------------------------------------------------------
// session
// ------------------------------------------------------------------------------------------------
struct StageCfg {
  int cap;      // 0 = global workspace
  int warps;    // warps per CTA
  int grid;     // CTAs
  size_t smem;  // dynamic shared memory per CTA
  int limit;
};

struct pprb200_session {
  int32_t n = 0;
  int32_t M = 0;  // non-sink nodes
  int64_t E = 0;
  uint32_t max_L = 0, hub_threshold = 0;
  int rank = 0, world = 1;
  int device = 0;                   // CUDA device the session lives on
  bool ipc = false;                 // basket buffers / mailbox are plain cudaMalloc allocations (exported through CUDA IPC)
  PeerDev peers;                    // device view of the peer mappings (world 1: zeroed)
  int* d_seq_list = nullptr;        // world > 1: own positions of the exact-order class (range_begin/end index it)
  unsigned char* d_need = nullptr;  // world > 1, [M]: ranks that read the basket at a position (PeerDev::need)
  unsigned char* d_owner = nullptr; // world > 1, [M]: rank that owns the position
  int pos_split = 0;                // first position of colour 1
  Mailbox* d_mbox = nullptr;        // [2][MAX_WORLD] barrier mailboxes of this rank, written by the peers
  void* ipc_opened[MAX_WORLD][3];   // peer mappings to close
  bool attached = false;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  // storage ranges [colour]: exact-order class (out-degree <= hub_threshold), stored first
  int range_begin[2] = {0, 0}, range_end[2] = {0, 0};
  int range_split[2] = {0, 0};  // exact-order nodes from here on have out-degree <= SEQ_SMALL_DEG
  // order-free class (out-degree > hub_threshold): work items [colour][0 = mid (128-thread CTAs), 1 = big (512)]
  int item_begin[2][2] = {{0, 0}, {0, 0}}, item_end[2][2] = {{0, 0}, {0, 0}};
  int n_items = 0;
  int chunk = 4096, mid_deg = 64;
  int hub_items[2] = {0, 0};          // leading items of the big list that are chunks of split hubs (out-degree > chunk)
  bool use_dense = true;              // merge_dense_kernel for single-item nodes (PPRB200_DENSE=0: merge_par only)
  int dense_threads = 512;            // CTA size of the big instantiation (PPRB200_DENSE_THREADS): 512 threads at 128 registers
                                      // beat 1024 at 64 with spills by 5 % on R-MAT-22 (profiles/r2/sweeps.txt)
  int* d_item_pos = nullptr;
  long long* d_item_off = nullptr;
  int* d_item_len = nullptr;
  unsigned char* d_pool = nullptr;
  unsigned int tbl_cap[2] = {0, 0};   // [0 = mid, 1 = big] slots per pool table
  int tbl_count_cls[2] = {0, 0};      // tables per class
  int tbl_first[2] = {0, 0};          // first index into tbl_inuse / tbl_count
  size_t pool_off[2] = {0, 0};        // byte offset of the class's tables in d_pool
  unsigned int* d_tbl_inuse = nullptr;
  unsigned int* d_tbl_count = nullptr;
  unsigned int* d_node_tbl = nullptr;
  unsigned int* d_node_done = nullptr;
  int32_t max_deg_seq = 0, max_deg_par = 0;
  unsigned long long* d_prof = nullptr;  // [2][sm_count*3][8] phase cycles of merge_par (debug, PPRB200_PROF=1)
  int32_t colour_count[2] = {0, 0};  // all nodes (sinks included) per colour
  int32_t max_deg = 0;
  // device
  long long* d_row_off = nullptr;
  unsigned long long* d_rowdeg = nullptr;  // [M] offset << ROWDEG_SHIFT | out-degree (MC walk kernel)
  uint32_t* d_col = nullptr;
  int* d_label = nullptr;
  int* d_pos_of = nullptr;
  int* d_dense_of = nullptr;
  unsigned char* d_colour = nullptr;
  unsigned char* d_buf[2] = {nullptr, nullptr};
  size_t buf_bytes = 0;
  unsigned int* d_queue[4] = {nullptr, nullptr, nullptr, nullptr};
  unsigned int* d_fb_queue = nullptr;  // [n_items] items merge_dense hands over to merge_par
  // hub teams
  int n_team_nodes[2] = {0, 0}, team_first[2] = {0, 0}, team_count[2] = {0, 0}, team_item_begin[2] = {0, 0}, team_item_end[2] = {0, 0};
  int* d_team_item_pos = nullptr;
  long long* d_team_item_off = nullptr;
  int* d_team_item_len = nullptr;
  int* d_team_item_team = nullptr;
  TeamInfo* d_teams = nullptr;
  TeamHeader* d_team_hdr = nullptr;
  unsigned char* d_stage = nullptr;
  size_t stage_bytes = 0;
  int* d_ncand = nullptr;
  RunState* d_state = nullptr;
  unsigned long long* d_final_stats = nullptr;
  unsigned char* d_ws = nullptr;
  size_t ws_bytes = 0;
  int* d_out_ids = nullptr;
  double* d_out_scores = nullptr;
  unsigned int* d_out_cnt = nullptr;
  size_t out_capacity = 0;  // n*K entries allocated
  // last run
  int last_mode = -1;
  uint32_t last_K = 0, last_L = 0, last_iterations = 0;
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
  std::vector<cudaEvent_t> ev_merge;  // pairs (begin,end) per iteration
  uint32_t merge_launches = 0;
  cudaEvent_t ev_walk[2] = {nullptr, nullptr};
  // the three node classes of a colour (big, mid, exact-order cascade) are launched on three streams so that the tail of
  // one persistent kernel overlaps the others; `cur` is the stream the launch helpers enqueue on
  cudaStream_t aux[2] = {nullptr, nullptr};
  cudaStream_t cur = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
  bool overlap = true;
  unsigned long long* d_walk_ws = nullptr;  // fallback visit-count tables of the MC walk phase
  size_t walk_ws_bytes = 0;
  uint64_t launch_count = 0;  // kernels enqueued by the last run
  double prep_ms = 0, h2d_ms = 0;
};

// L as the kernels see it: keepTop(L) is the identity once L >= n (a basket holds at most n distinct keys)
static uint32_t effective_L(uint32_t L, int32_t n) { return std::min<uint32_t>(L, (uint32_t)std::max<int32_t>(n, 1)); }

static int device_ok() {
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) {
    cudaGetLastError();
    return fail(PPRB200_ERR_CUDA, "no CUDA device available (libppr_b200 has no CPU fallback)");
  }
  int dev = 0;
  cudaGetDevice(&dev);
  static bool checked[64] = {false};  // (attribute queries cost a millisecond or two: ask once per device)
  if (checked[dev & 63]) return PPRB200_OK;
  int major = 0, minor = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess)
    return fail(PPRB200_ERR_CUDA, "cudaDeviceGetAttribute failed");
  if (major != 10) return fail(PPRB200_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", dev, major, minor);
  checked[dev & 63] = true;
  return PPRB200_OK;
}

// Device memory comes from the stream-ordered default pool with an unlimited release threshold: after the first call of
// a process, allocating and freeing a session's buffers costs microseconds instead of cudaMalloc/cudaFree round trips
// (the host-buffer entry points build and drop a session per call). Buffers that peers map through CUDA IPC must be
// plain cudaMalloc allocations (`ipc` = true).
static cudaStream_t g_alloc_stream = nullptr;

static void pool_setup_once() {
  static bool done[64] = {false};  // per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (done[dev & 63]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  cudaGetLastError();
  done[dev & 63] = true;
}

template <typename T>
static int dev_alloc(T** p, size_t count, bool ipc = false) {
  *p = nullptr;
  if (count == 0) count = 1;
  cudaError_t e = ipc ? cudaMalloc((void**)p, count * sizeof(T)) : cudaMallocAsync((void**)p, count * sizeof(T), g_alloc_stream);
  if (e != cudaSuccess) return fail(PPRB200_ERR_ALLOC, "device allocation of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
  return PPRB200_OK;
}

static void dev_free(void* p, bool ipc = false) {
  if (!p) return;
  if (ipc) cudaFree(p);
  else cudaFreeAsync(p, g_alloc_stream);
}

static void session_free(pprb200_session* s) {
  if (!s) return;
  for (int r = 0; r < MAX_WORLD; r++)
    for (int i = 0; i < 3; i++)
      if (s->ipc_opened[r][i]) cudaIpcCloseMemHandle(s->ipc_opened[r][i]);
  g_alloc_stream = s->stream;
  const bool ipc = s->ipc;
  dev_free(s->d_mbox, ipc);
  dev_free(s->d_buf[0], ipc); dev_free(s->d_buf[1], ipc);
  void* plain[] = {s->d_need, s->d_owner, s->d_rowdeg, s->d_seq_list, s->d_row_off, s->d_col, s->d_label, s->d_pos_of, s->d_dense_of, s->d_colour, s->d_queue[0], s->d_queue[1],
                   s->d_queue[2], s->d_queue[3], s->d_ncand, s->d_state, s->d_final_stats, s->d_ws, s->d_out_ids, s->d_out_scores,
                   s->d_out_cnt, s->d_item_pos, s->d_item_off, s->d_item_len, s->d_pool, s->d_walk_ws, s->d_prof, s->d_tbl_inuse,
                   s->d_tbl_count, s->d_node_tbl, s->d_node_done, s->d_fb_queue, s->d_team_item_pos, s->d_team_item_off,
                   s->d_team_item_len, s->d_team_item_team, s->d_teams, s->d_team_hdr, s->d_stage};
  for (void* q : plain) dev_free(q);
  for (int i = 0; i < 2; i++) if (s->ev_walk[i]) cudaEventDestroy(s->ev_walk[i]);
  for (int i = 0; i < 2; i++) { if (s->aux[i]) cudaStreamDestroy(s->aux[i]); if (s->ev_join[i]) cudaEventDestroy(s->ev_join[i]); }
  if (s->ev_fork) cudaEventDestroy(s->ev_fork);
  if (s->ev_begin) cudaEventDestroy(s->ev_begin);
  if (s->ev_end) cudaEventDestroy(s->ev_end);
  for (auto e : s->ev_merge) cudaEventDestroy(e);
  delete s;
}

// Largest out-degree of the mid class (128-thread CTAs, H = TCAP = 2048). The larger the label space, the larger the
// share of tail labels per basket and the earlier the 2048-slot tail table fills up (measured on R-MAT 16..22,
// profiles/r1/sweeps.txt): 64 up to 256 K nodes, 48 up to 2 M, 32 above.
static bool dense_enabled() {
  const char* e = getenv("PPRB200_DENSE");
  return !e || atoi(e) != 0;
}

static int default_mid_deg(int32_t n) {
  if (const char* e = getenv("PPRB200_MID_DEG")) return std::min(PAR_MID_MAX, std::max(1, atoi(e)));
  // merge_dense_kernel: the 512-thread instantiation (two CTAs per SM) up to 128 successors
  if (dense_enabled()) return PAR_MID_MAX;
  return n <= (1 << 18) ? 64 : (n <= (1 << 21) ? 48 : 32);
}

// Successors per work item of the big class. A node above it is split into chunks that meet in a global table (single
// pass, tail labels spill to L2) instead of taking the two-pass shared-memory scheme, which costs far more per entry
// than it gains in balance: n/64 clamped to [2048, 32768] (measured: R-MAT-16 2048, -18 4096, -20 16384, -22 32768).
// With several GPUs every rank holds 1/world of the work, so the longest single-CTA item must shrink with it or it
// becomes the tail of every iteration (R-MAT-22 on 8 GPUs: 99 M node-iterations/s with 4096, 92 M with 32768).
static int default_chunk(int32_t n, int world) {
  if (const char* e = getenv("PPRB200_CHUNK")) return std::min(1 << 20, std::max(32, atoi(e)));
  // merge_dense_kernel: a whole hub on one CTA (the largest R-MAT-22 hub, 16 M entries, is ~8 ms of a >= 10 ms
  // iteration) beats its chunks meeting in an L2-resident table by a wide margin (profiles/r2/launches_r22_v2.txt)
  // (several GPUs: the same. Measured on 2 B200s, R-MAT-22: the 23 hubs above 32768 successors, split into chunks for
  // balance, cost ~10 ms per iteration and rank on merge_par -- a fifth of the iteration -- and 1.29x was all two GPUs gave)
  if (dense_enabled()) return 1 << 20;
  int c = 2048;
  while (c < 32768 && (long long)c * 64 * std::max(world, 1) < (long long)n) c <<= 1;
  return c;
}

// Stable counting sort of the nodes 0..n-1 on a small key, on all host threads: every thread counts a contiguous id range
// into its own histogram, an exclusive scan in (key, thread) order turns the counts into write offsets, and the same ranges
// scatter. Ties keep ascending id order whatever the thread count (ranges are contiguous and scanned in order).
// key(v) < K, or SORT_SKIP to leave the node out. start[k] = first output index of key k (start[K] = total).
constexpr uint32_t SORT_SKIP = 0xFFFFFFFFu;
template <typename KeyFn>
static void parallel_counting_sort(int32_t n, size_t K, KeyFn key, std::vector<int32_t>& start, std::vector<int32_t>& out) {
  const int parts = (int)std::max<int64_t>(1, std::min<int64_t>(host_threads(), ((int64_t)n + (1 << 16) - 1) >> 16));
  std::vector<std::vector<int32_t>> hist((size_t)parts);
  host_parallel(parts, [&](int t) {
    std::vector<int32_t>& h = hist[(size_t)t];
    h.assign(K, 0);
    for (int64_t v = (int64_t)n * t / parts, hi = (int64_t)n * (t + 1) / parts; v < hi; v++) {
      const uint32_t k = key((int32_t)v);
      if (k != SORT_SKIP) h[k]++;
    }
  });
  start.assign(K + 1, 0);
  int32_t run = 0;
  for (size_t k = 0; k < K; k++) {
    start[k] = run;
    for (int t = 0; t < parts; t++) { const int32_t c = hist[(size_t)t][k]; hist[(size_t)t][k] = run; run += c; }
  }
  start[K] = run;
  out.assign((size_t)run, 0);
  host_parallel(parts, [&](int t) {
    std::vector<int32_t>& h = hist[(size_t)t];
    for (int64_t v = (int64_t)n * t / parts, hi = (int64_t)n * (t + 1) / parts; v < hi; v++) {
      const uint32_t k = key((int32_t)v);
      if (k != SORT_SKIP) out[(size_t)h[k]++] = (int32_t)v;
    }
  });
}

// degrees up to SORT_DEG_CAP - 2 get a counting-sort key of their own; the few nodes above share the first key of their
// group and are put in order by a comparison sort afterwards (power-law graphs: thousands of nodes, not millions)
constexpr int64_t SORT_DEG_CAP = 4096;

// storage positions of the non-sink nodes: colour-major, then class (0 exact-order, 1 mid, 2 big), then out-degree
// descending (ties by dense id)
// owner_of_pos (world > 1): longest-processing-time-first over each (colour, class) list, work = out-degree, so that
// every iteration's merged entries are balanced up to the single largest node.
static void storage_order(const int64_t* row_ptr, int32_t n, const uint8_t* colour, uint32_t hub_threshold, int mid_deg,
                          std::vector<int32_t>& order, int cls_begin[2][3], int cls_end[2][3], int world = 1,
                          std::vector<int32_t>* owner_of_pos = nullptr) {
  // key = (colour, class, out-degree descending): slot 0 of a group = "SORT_DEG_CAP - 1 successors or more"
  const size_t per_group = (size_t)SORT_DEG_CAP;
  auto key_of = [&](int32_t v) -> uint32_t {
    const int64_t d = row_ptr[v + 1] - row_ptr[v];
    if (d <= 0) return SORT_SKIP;
    const int cls = (uint64_t)d <= (uint64_t)hub_threshold ? 0 : (d <= mid_deg ? 1 : 2);
    const int64_t dd = std::min<int64_t>(d, SORT_DEG_CAP - 1);
    return (uint32_t)(((size_t)colour[v] * 3 + (size_t)cls) * per_group + (size_t)(SORT_DEG_CAP - 1 - dd));
  };
  std::vector<int32_t> start;
  parallel_counting_sort(n, 6 * per_group, key_of, start, order);
  for (int c = 0; c < 2; c++)
    for (int cls = 0; cls < 3; cls++) {
      const size_t g = (size_t)c * 3 + (size_t)cls;
      cls_begin[c][cls] = start[g * per_group];
      cls_end[c][cls] = start[(g + 1) * per_group];
      // the group's large nodes (all under its first key, in id order): out-degree descending, ties by id
      std::stable_sort(order.begin() + start[g * per_group], order.begin() + start[g * per_group + 1], [&](int32_t a, int32_t b) {
        return row_ptr[a + 1] - row_ptr[a] > row_ptr[b + 1] - row_ptr[b];
      });
    }
  if (owner_of_pos) {
    owner_of_pos->assign(order.size(), 0);
    if (world > 1)
      for (int c = 0; c < 2; c++) {
        std::vector<long long> load((size_t)world, 0);  // carried across the classes of one colour: they run back to back
        for (int cls = 2; cls >= 0; cls--)
          for (int p = cls_begin[c][cls]; p < cls_end[c][cls]; p++) {
            int best = 0;
            for (int r = 1; r < world; r++)
              if (load[r] < load[best]) best = r;
            (*owner_of_pos)[(size_t)p] = best;
            load[best] += row_ptr[order[(size_t)p] + 1] - row_ptr[order[(size_t)p]];
          }
      }
  }
}

// pprb200_debug_host_plan: the host front half alone (no device is touched), copied out for the CPU tests
struct HostPlanOut {
  int32_t* pos_of;      // [n] storage position or -1 (sink)
  int32_t* rank_of;     // [n] rank label
  int64_t* row_off;     // [n+1] (M+1 used)
  uint32_t* enc;        // [E] column words in storage order
  int32_t* item_pos;    // [item_cap]
  int64_t* item_off;
  int32_t* item_len;
  int32_t item_cap;
  int32_t* summary;     // [16] M, n_items, chunk, mid_deg, range_begin[2], range_end[2], item_begin[2][2], item_end[2][2]
};

// The host front half, split in two so that a multi-GPU run plans once:
//   HostPlan  rank-independent: colouring, storage order (+ owner of every position), rank labels, CSR in storage order
//   RankPlan  what one rank of `world` works on: its exact-order positions and its order-free work items
struct HostPlan {
  int32_t n = 0, M = 0;
  int64_t E = 0;
  int world = 1;
  uint32_t hub_threshold = 0;
  int chunk = 4096, mid_deg = 64;
  int32_t colour_count[2] = {0, 0};
  int32_t max_deg = 0, max_deg_seq = 0;
  int cls_begin[2][3], cls_end[2][3];
  std::vector<uint8_t> colour;
  std::vector<int32_t> order, owner_of_pos, dense_of, rank_of, pos_of, label;
  std::vector<long long> row_off;
  std::vector<unsigned long long> rowdeg;
  std::vector<uint8_t> owner8, need_mask;  // world > 1, per position: owning rank; ranks that read the basket (bit r: rank r owns a predecessor)
  std::vector<uint32_t> enc;     // column words in storage order -- on the host (small graphs, pprb200_debug_host_plan) ...
  uint32_t* d_enc = nullptr;     // ... or produced on the device (plan_device.cuh): then `enc` stays empty
  int enc_device = 0;
  double prep_ms = 0;
  HostPlan() = default;
  HostPlan(const HostPlan&) = delete;
  HostPlan& operator=(const HostPlan&) = delete;
  ~HostPlan() {
    if (d_enc) {
      int cur = 0;
      cudaGetDevice(&cur);
      cudaSetDevice(enc_device);
      cudaFree(d_enc);
      cudaSetDevice(cur);
    }
  }
};

// The caller's CSR on the device while the plan is made: the colouring BFS and the column-word encode read it there
// (plan_device.cuh). Allocated with plain cudaMalloc/cudaFree on the legacy stream: a plan is made once per call.
struct RawCsrDev {
  long long* row_ptr = nullptr;
  int* col = nullptr;
  int32_t n = 0;
  int64_t E = 0;
  int device = 0;
  bool ok = false;
  ~RawCsrDev() { release(); }
  void release() {
    if (row_ptr) cudaFreeAsync(row_ptr, 0);
    if (col) cudaFreeAsync(col, 0);
    row_ptr = nullptr; col = nullptr; ok = false;
  }
  bool upload(const int64_t* h_row_ptr, const int32_t* h_col, int32_t n_, int64_t E_) {
    n = n_; E = E_;
    cudaGetDevice(&device);
    if (cudaMallocAsync((void**)&row_ptr, ((size_t)n + 1) * sizeof(long long), 0) != cudaSuccess ||
        cudaMallocAsync((void**)&col, (size_t)std::max<int64_t>(E, 1) * sizeof(int), 0) != cudaSuccess ||
        cudaMemcpyAsync(row_ptr, h_row_ptr, ((size_t)n + 1) * sizeof(long long), cudaMemcpyHostToDevice, 0) != cudaSuccess ||
        cudaMemcpyAsync(col, h_col, (size_t)E * sizeof(int), cudaMemcpyHostToDevice, 0) != cudaSuccess) {
      cudaGetLastError();
      release();
      return false;
    }
    ok = true;
    return true;
  }
};

// find_partitions' device_component: level-synchronous BFS of root's component, one launch per level
static bool device_component(const RawCsrDev& G, int32_t root, uint8_t* seen, uint8_t* colour) {
  const int32_t n = G.n;
  unsigned char* d_level = nullptr;
  unsigned int* d_changed = nullptr;
  if (cudaMallocAsync((void**)&d_level, (size_t)n, 0) != cudaSuccess || cudaMallocAsync((void**)&d_changed, sizeof(unsigned int), 0) != cudaSuccess) {
    cudaGetLastError();
    if (d_level) cudaFreeAsync(d_level, 0);
    return false;
  }
  int sm = 148;
  cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, G.device);
  cudaMemsetAsync(d_level, BFS_UNSEEN, (size_t)n, 0);
  cudaMemsetAsync(d_level + root, 0, 1, 0);
  int cur = 0;
  bool ok = true;
  for (; cur < BFS_MAX_LEVEL; cur++) {
    unsigned int changed = 0;
    cudaMemsetAsync(d_changed, 0, sizeof(unsigned int), 0);
    bfs_level_kernel<<<sm * 8, 256, 0, 0>>>(G.row_ptr, G.col, n, d_level, (unsigned char)cur, d_changed);
    if (cudaMemcpyAsync(&changed, d_changed, sizeof(unsigned int), cudaMemcpyDeviceToHost, 0) != cudaSuccess || cudaStreamSynchronize(0) != cudaSuccess) { ok = false; break; }
    if (!changed) break;
  }
  if (cur >= BFS_MAX_LEVEL) ok = false;  // a long chain: the host's frontier-based BFS does this one
  if (ok) {
    std::vector<unsigned char> level((size_t)n);
    ok = cudaMemcpyAsync(level.data(), d_level, (size_t)n, cudaMemcpyDeviceToHost, 0) == cudaSuccess && cudaStreamSynchronize(0) == cudaSuccess;
    if (ok)
      host_parallel_for(n, 1 << 16, [&](int, int64_t lo, int64_t hi) {
        for (int64_t v = lo; v < hi; v++)
          if (level[(size_t)v] != BFS_UNSEEN) { seen[v] = 1; colour[v] = (uint8_t)(level[(size_t)v] & 1u); }
      });
  }
  cudaGetLastError();
  cudaFreeAsync(d_level, 0);
  cudaFreeAsync(d_changed, 0);
  return ok;
}

constexpr int SEQ_SMALL_DEG = 3;  // exact-order nodes up to this out-degree fit 512-slot tables: twice the warps per SM

struct RankPlan {
  int range_begin[2] = {0, 0}, range_end[2] = {0, 0};
  int range_split[2] = {0, 0};  // [range_split, range_end): out-degree <= SEQ_SMALL_DEG (the class is stored by out-degree descending)
  std::vector<int> seq_list;  // world > 1: this rank's positions of the exact-order class, colour-major
  int item_begin[2][2] = {{0, 0}, {0, 0}}, item_end[2][2] = {{0, 0}, {0, 0}};
  int hub_items[2] = {0, 0};
  int32_t max_deg_par = 0;
  std::vector<int> item_pos;
  std::vector<long long> item_off;
  std::vector<int> item_len;
  // hub teams (merge_dense.cuh): the leading n_team_nodes[c] big items of colour c are also cut into chunks
  int n_team_nodes[2] = {0, 0};
  int team_first[2] = {0, 0}, team_count[2] = {0, 0};          // into `teams`
  int team_item_begin[2] = {0, 0}, team_item_end[2] = {0, 0};  // into the team item arrays
  std::vector<TeamInfo> teams;
  std::vector<int> team_item_pos, team_item_len, team_item_team;
  std::vector<long long> team_item_off;
};

// out-degree above which a big-class node is worked on by a team of CTAs, and the successors per member
static int team_min_deg() {
  if (const char* e = getenv("PPRB200_TEAM_DEG")) return std::max(PAR_MID_MAX + 1, atoi(e));
  return 16384;
}
static int team_chunk_len() {
  if (const char* e = getenv("PPRB200_TEAM_CHUNK")) return std::max(32, atoi(e));
  return 8192;
}

// use_device: the edge-sized parts of the plan (colouring BFS of the large component, column words) run on the current
// device when the graph is large enough to pay for the upload; pprb200_debug_host_plan (no device) keeps everything here
static int build_host_plan(const int64_t* row_ptr, const int32_t* col, int32_t n, const uint8_t* colour_in, uint32_t hub_threshold,
                           int32_t world, bool need_colour, HostPlan& H, bool use_device = false) {
  const double t0 = now_ms();
  int rc;
  H.n = n;
  H.world = world;
  H.hub_threshold = hub_threshold == 0 ? PPRB200_DEFAULT_HUB_THRESHOLD : hub_threshold;
  H.colour.assign((size_t)n, 0);
  std::vector<uint8_t>& colour = H.colour;
  RawCsrDev G;
  if (use_device && row_ptr[n] >= (1ll << 20) && !getenv("PPRB200_HOST_PLAN")) G.upload(row_ptr, col, n, row_ptr[n]);
  if (colour_in) std::memcpy(colour.data(), colour_in, (size_t)n);
  else if (need_colour) {  // (MC-only session: one class)
    const ComponentFn on_device = [&](int32_t root, uint8_t* seen, uint8_t* col_out) { return device_component(G, root, seen, col_out); };
    if ((rc = find_partitions(row_ptr, col, n, colour.data(), G.ok ? &on_device : nullptr))) return rc;
  }
  for (int32_t v = 0; v < n; v++) {
    if (colour[v] > 1) return fail(PPRB200_ERR_PARAM, "colour[%d] = %d is not 0/1", v, colour[v]);
    H.colour_count[colour[v]]++;
  }
  const double t_col = now_ms();
  // storage order: colour-major; inside a colour the exact-order class first, then the order-free class
  // (mid, big); every class by out-degree descending (ties by dense id) -- big nodes first for load balance
  H.chunk = default_chunk(n, world);
  H.mid_deg = default_mid_deg(n);
  storage_order(row_ptr, n, colour.data(), H.hub_threshold, H.mid_deg, H.order, H.cls_begin, H.cls_end, world, &H.owner_of_pos);
  const std::vector<int32_t>& order = H.order;
  const double t_ord = now_ms();
  // rank labels: in-degree descending, ties by dense id (the keys stored in the baskets) -- a counting sort
  H.rank_of.assign((size_t)n, 0);
  {
    std::vector<uint32_t> indeg((size_t)n);
    host_indegree(col, row_ptr[n], n, indeg.data());
    std::vector<int32_t> start;
    parallel_counting_sort(n, (size_t)SORT_DEG_CAP, [&](int32_t v) -> uint32_t {
      return (uint32_t)(SORT_DEG_CAP - 1 - (int64_t)std::min<uint32_t>(indeg[(size_t)v], (uint32_t)(SORT_DEG_CAP - 1)));
    }, start, H.dense_of);
    std::stable_sort(H.dense_of.begin(), H.dense_of.begin() + start[1], [&](int32_t a, int32_t b) { return indeg[(size_t)a] > indeg[(size_t)b]; });
    host_parallel_for(n, 1 << 16, [&](int, int64_t lo, int64_t hi) {
      for (int64_t r = lo; r < hi; r++) H.rank_of[(size_t)H.dense_of[(size_t)r]] = (int32_t)r;
    });
  }
  const double t_rank = now_ms();
  const int32_t M = (int32_t)order.size();
  H.M = M;
  H.pos_of.assign((size_t)n, -1);
  H.row_off.assign((size_t)M + 1, 0);
  H.rowdeg.assign((size_t)std::max(M, 1), 0ull);
  H.label.assign((size_t)std::max(M, 1), 0);
  {
    // offsets in storage order: per-range sums, a scan over the ranges, then every range writes its part
    const int parts = (int)std::max<int64_t>(1, std::min<int64_t>(host_threads(), ((int64_t)M + (1 << 16) - 1) >> 16));
    std::vector<long long> base((size_t)parts + 1, 0);
    std::vector<long long> part_max((size_t)parts, 0);
    host_parallel(parts, [&](int t) {
      long long sum = 0, mx = 0;
      for (int64_t p = (int64_t)M * t / parts, hi = (int64_t)M * (t + 1) / parts; p < hi; p++) {
        const int64_t d = row_ptr[order[(size_t)p] + 1] - row_ptr[order[(size_t)p]];
        sum += d;
        mx = std::max<long long>(mx, d);
      }
      base[(size_t)t + 1] = sum;
      part_max[(size_t)t] = mx;
    });
    long long max_d = 0;
    for (int t = 0; t < parts; t++) { base[(size_t)t + 1] += base[(size_t)t]; max_d = std::max(max_d, part_max[(size_t)t]); }
    H.max_deg = (int32_t)std::min<long long>(max_d, INT32_MAX);
    if ((unsigned long long)max_d >> ROWDEG_SHIFT || (unsigned long long)base[(size_t)parts] >> (64 - ROWDEG_SHIFT))
      return fail(PPRB200_ERR_GRAPH, "out-degree %lld / edge count %lld exceed the packed row word (2^%d successors per node, 2^%d edges)",
                  max_d, base[(size_t)parts], ROWDEG_SHIFT, 64 - ROWDEG_SHIFT);
    host_parallel(parts, [&](int t) {
      long long off = base[(size_t)t];
      for (int64_t p = (int64_t)M * t / parts, hi = (int64_t)M * (t + 1) / parts; p < hi; p++) {
        const int32_t v = order[(size_t)p];
        const int64_t d = row_ptr[v + 1] - row_ptr[v];
        H.pos_of[(size_t)v] = (int32_t)p;
        H.row_off[(size_t)p] = off;
        H.rowdeg[(size_t)p] = ((unsigned long long)off << ROWDEG_SHIFT) | (unsigned long long)d;
        H.label[(size_t)p] = H.rank_of[(size_t)v];
        off += d;
      }
    });
    H.row_off[(size_t)M] = base[(size_t)parts];
  }
  const int64_t E = H.row_off[M];
  H.E = E;
  for (int c = 0; c < 2; c++)  // (a class is stored by out-degree descending: its first node has the largest)
    if (H.cls_end[c][0] > H.cls_begin[c][0]) {
      const int p = H.cls_begin[c][0];
      H.max_deg_seq = std::max<int32_t>(H.max_deg_seq, (int32_t)std::min<long long>(H.row_off[(size_t)p + 1] - H.row_off[p], INT32_MAX));
    }
  const double t_off = now_ms();
  // column words: one lookup table (word of every node), then a gather -- on the device over the uploaded CSR, or here in
  // parallel over edge-balanced position ranges
  std::vector<uint32_t> word_of((size_t)n);
  host_parallel_for(n, 1 << 15, [&](int, int64_t lo, int64_t hi) {
    for (int64_t v = lo; v < hi; v++)
      word_of[(size_t)v] = H.pos_of[(size_t)v] < 0 ? (COL_SINK | (uint32_t)H.rank_of[(size_t)v])
                                                    : ((uint32_t)H.pos_of[(size_t)v] | ((uint32_t)colour[(size_t)v] << COL_COLOUR_SHIFT));
  });
  bool encoded = false;
  if (G.ok && M > 0 && E > 0) {
    int* d_order = nullptr;
    long long* d_off = nullptr;
    unsigned int* d_word = nullptr;
    H.enc_device = G.device;
    int sm = 148;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, G.device);
    if (cudaMalloc((void**)&H.d_enc, (size_t)E * sizeof(uint32_t)) == cudaSuccess &&
        cudaMallocAsync((void**)&d_order, (size_t)M * sizeof(int), 0) == cudaSuccess &&
        cudaMallocAsync((void**)&d_off, ((size_t)M + 1) * sizeof(long long), 0) == cudaSuccess &&
        cudaMallocAsync((void**)&d_word, (size_t)n * sizeof(unsigned int), 0) == cudaSuccess &&
        cudaMemcpyAsync(d_order, order.data(), (size_t)M * sizeof(int), cudaMemcpyHostToDevice, 0) == cudaSuccess &&
        cudaMemcpyAsync(d_off, H.row_off.data(), ((size_t)M + 1) * sizeof(long long), cudaMemcpyHostToDevice, 0) == cudaSuccess &&
        cudaMemcpyAsync(d_word, word_of.data(), (size_t)n * sizeof(unsigned int), cudaMemcpyHostToDevice, 0) == cudaSuccess) {
      encode_kernel<<<sm * 8, 256, 0, 0>>>(G.row_ptr, G.col, d_order, d_off, d_word, M, H.d_enc);
      encoded = cudaGetLastError() == cudaSuccess && cudaStreamSynchronize(0) == cudaSuccess;
      if (encoded && world > 1) {  // who reads whose basket (publish_slot)
        unsigned char* d_owner = nullptr;
        unsigned int* d_need = nullptr;
        const size_t words = ((size_t)M + 3) / 4;
        H.owner8.resize((size_t)M);
        for (int32_t p2 = 0; p2 < M; p2++) H.owner8[(size_t)p2] = (uint8_t)H.owner_of_pos[(size_t)p2];
        H.need_mask.assign(words * 4, 0);
        bool ok = cudaMallocAsync((void**)&d_owner, (size_t)M, 0) == cudaSuccess && cudaMallocAsync((void**)&d_need, words * 4, 0) == cudaSuccess &&
                  cudaMemcpyAsync(d_owner, H.owner8.data(), (size_t)M, cudaMemcpyHostToDevice, 0) == cudaSuccess &&
                  cudaMemsetAsync(d_need, 0, words * 4, 0) == cudaSuccess;
        if (ok) {
          need_mask_kernel<<<sm * 8, 256, 0, 0>>>(d_off, H.d_enc, d_owner, M, d_need);
          ok = cudaMemcpyAsync(H.need_mask.data(), d_need, words * 4, cudaMemcpyDeviceToHost, 0) == cudaSuccess && cudaStreamSynchronize(0) == cudaSuccess;
        }
        cudaGetLastError();
        if (d_owner) cudaFreeAsync(d_owner, 0);
        if (d_need) cudaFreeAsync(d_need, 0);
        if (!ok) H.need_mask.clear();  // (computed on the host below)
      }
    }
    cudaGetLastError();
    if (d_order) cudaFreeAsync(d_order, 0);
    if (d_off) cudaFreeAsync(d_off, 0);
    if (d_word) cudaFreeAsync(d_word, 0);
    if (!encoded && H.d_enc) { cudaFree(H.d_enc); H.d_enc = nullptr; }
  }
  G.release();
  if (!encoded) {
    H.enc.assign((size_t)std::max<int64_t>(E, 1), 0u);
    const int parts = (int)std::max<int64_t>(1, std::min<int64_t>(host_threads(), E / (1 << 15)));
    host_parallel(parts, [&](int t) {
      const int32_t p_lo = (int32_t)(std::lower_bound(H.row_off.begin(), H.row_off.end(), (long long)(E * t / parts)) - H.row_off.begin());
      const int32_t p_hi = t + 1 == parts ? M : (int32_t)(std::lower_bound(H.row_off.begin(), H.row_off.end(), (long long)(E * (t + 1) / parts)) - H.row_off.begin());
      for (int32_t p = p_lo; p < p_hi; p++) {
        const int32_t v = order[(size_t)p];
        long long o = H.row_off[(size_t)p];
        for (int64_t i = row_ptr[v]; i < row_ptr[v + 1]; i++) H.enc[(size_t)o++] = word_of[(size_t)col[i]];
      }
    });
  }
  if (world > 1 && M > 0 && H.need_mask.empty()) {
    H.owner8.resize((size_t)M);
    for (int32_t p2 = 0; p2 < M; p2++) H.owner8[(size_t)p2] = (uint8_t)H.owner_of_pos[(size_t)p2];
    if (!H.enc.empty()) {
      H.need_mask.assign(((size_t)M + 3) / 4 * 4, 0);
      host_parallel_for(M, 1 << 14, [&](int, int64_t lo, int64_t hi) {
        for (int64_t p2 = lo; p2 < hi; p2++) {
          const uint8_t bit = (uint8_t)(1u << H.owner8[(size_t)p2]);
          for (long long i = H.row_off[(size_t)p2]; i < H.row_off[(size_t)p2 + 1]; i++) {
            const uint32_t w = H.enc[(size_t)i];
            if (!(w & COL_SINK) && !(H.need_mask[w & COL_POS_MASK] & bit)) __atomic_fetch_or(&H.need_mask[w & COL_POS_MASK], bit, __ATOMIC_RELAXED);
          }
        }
      });
    }  // (else: the column words exist on the device only and the mask kernel failed -- everybody gets everything)
  }
  H.prep_ms = now_ms() - t0;
  if (getenv("PPRB200_HOST_TIMING"))
    fprintf(stderr, "[pprb200] host plan %.2f ms: colour %.2f, storage order %.2f, rank labels %.2f, offsets %.2f, encode %.2f (%d threads)\n",
            H.prep_ms, t_col - t0, t_ord - t_col, t_rank - t_ord, t_off - t_rank, now_ms() - t_off, host_threads());
  return PPRB200_OK;
}

static void build_rank_plan(const HostPlan& H, int32_t rank, RankPlan& R) {
  const int world = H.world;
  for (int c = 0; c < 2; c++) {
    if (world == 1) {
      R.range_begin[c] = H.cls_begin[c][0];
      R.range_end[c] = H.cls_end[c][0];
      R.range_split[c] = R.range_end[c];
      for (int p = R.range_begin[c]; p < R.range_end[c]; p++)
        if (H.row_off[(size_t)p + 1] - H.row_off[p] <= SEQ_SMALL_DEG) { R.range_split[c] = p; break; }
    } else {
      R.range_begin[c] = (int)R.seq_list.size();
      R.range_split[c] = -1;
      for (int p = H.cls_begin[c][0]; p < H.cls_end[c][0]; p++)
        if (H.owner_of_pos[(size_t)p] == rank) {
          if (R.range_split[c] < 0 && H.row_off[(size_t)p + 1] - H.row_off[p] <= SEQ_SMALL_DEG) R.range_split[c] = (int)R.seq_list.size();
          R.seq_list.push_back(p);
        }
      R.range_end[c] = (int)R.seq_list.size();
      if (R.range_split[c] < 0) R.range_split[c] = R.range_end[c];
    }
  }
  const bool teams_on = dense_enabled() && !getenv("PPRB200_NO_TEAMS");
  const long long tdeg = team_min_deg(), tchunk = team_chunk_len();
  for (int c = 0; c < 2; c++) {
    R.team_first[c] = (int)R.teams.size();
    R.team_item_begin[c] = (int)R.team_item_pos.size();
    for (int cls = 1; cls < 3; cls++) {
      R.item_begin[c][cls - 1] = (int)R.item_pos.size();
      for (int p = H.cls_begin[c][cls]; p < H.cls_end[c][cls]; p++) {
        if (H.owner_of_pos[(size_t)p] != rank) continue;  // multi-GPU: somebody else's node
        const long long d = H.row_off[(size_t)p + 1] - H.row_off[p];
        if (d > R.max_deg_par) R.max_deg_par = (int32_t)std::min<long long>(d, INT32_MAX);
        if (teams_on && cls == 2 && d > tdeg && d <= H.chunk && R.n_team_nodes[c] == (int)R.item_pos.size() - R.item_begin[c][1]) {
          // (the class is stored by out-degree descending: team hubs are its leading items)
          const int nch = (int)std::min<long long>(TEAM_MAX_CHUNKS, (d + tchunk - 1) / tchunk);
          if (nch >= 2) {
            const long long per = (d + nch - 1) / nch;
            TeamInfo ti;
            ti.regular_item = (int)R.item_pos.size();
            ti.nchunks = 0;
            for (long long o = 0; o < d; o += per) {
              R.team_item_pos.push_back(p);
              R.team_item_off.push_back(H.row_off[p] + o);
              R.team_item_len.push_back((int)std::min<long long>(per, d - o));
              R.team_item_team.push_back((int)R.teams.size());
              ti.nchunks++;
            }
            R.teams.push_back(ti);
            R.n_team_nodes[c]++;
          }
        }
        if (cls == 2 && d > H.chunk) R.hub_items[c] += (int)((d + H.chunk - 1) / H.chunk);
        for (long long o = 0; o < d; o += H.chunk) {
          R.item_pos.push_back(p);
          R.item_off.push_back(H.row_off[p] + o);
          R.item_len.push_back((int)std::min<long long>(H.chunk, d - o));
        }
      }
      R.item_end[c][cls - 1] = (int)R.item_pos.size();
    }
    R.team_count[c] = (int)R.teams.size() - R.team_first[c];
    R.team_item_end[c] = (int)R.team_item_pos.size();
  }
}

// allocate + upload one rank's session on the CURRENT device
// `src`: a session of the same plan on ANOTHER device of this process (peer access enabled): the rank-independent arrays
// -- CSR, labels, maps, 0.33 GB on R-MAT-22 -- are then copied from there over NVLink instead of crossing PCIe once per GPU
static int session_from_plan(const HostPlan& H, const RankPlan& R, uint32_t max_L, int32_t rank, void* stream, bool ipc,
                             pprb200_session** out, const pprb200_session* src = nullptr) {
  int rc;
  *out = nullptr;
  const double t1 = now_ms();
  pprb200_session* s = new pprb200_session();
  std::memset(&s->peers, 0, sizeof(s->peers));
  std::memset(s->ipc_opened, 0, sizeof(s->ipc_opened));
  pool_setup_once();
  g_alloc_stream = (cudaStream_t)stream;
  const int32_t n = H.n, M = H.M;
  const int64_t E = H.E;
  const int world = H.world;
  s->n = n; s->M = M; s->E = E;
  s->max_L = max_L;
  s->hub_threshold = H.hub_threshold;
  s->rank = rank;
  s->world = world;
  s->ipc = ipc;
  s->stream = (cudaStream_t)stream;
  cudaGetDevice(&s->device);
  cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, s->device);
  s->chunk = H.chunk; s->mid_deg = H.mid_deg;
  s->max_deg = H.max_deg; s->max_deg_seq = H.max_deg_seq; s->max_deg_par = R.max_deg_par;
  for (int c = 0; c < 2; c++) {
    s->colour_count[c] = H.colour_count[c];
    s->range_begin[c] = R.range_begin[c]; s->range_end[c] = R.range_end[c]; s->range_split[c] = R.range_split[c];
    s->hub_items[c] = R.hub_items[c];
    for (int k = 0; k < 2; k++) { s->item_begin[c][k] = R.item_begin[c][k]; s->item_end[c][k] = R.item_end[c][k]; }
  }
  s->n_items = (int)R.item_pos.size();
  s->prep_ms = H.prep_ms;
  // a basket never holds more than n keys: slots are sized for min(max_L, n) (runs clamp L the same way), so that the
  // reference's grank(graph, graph.size(), 2 * graph.size(), ...) pattern (test/grankTest.cc:261-283) costs nothing extra
  const int Lp = roundup4((int)effective_L(max_L, n));
  s->buf_bytes = (size_t)std::max(M, 1) * slot_bytes(Lp);
  if ((rc = dev_alloc(&s->d_row_off, (size_t)M + 1)) || (rc = dev_alloc(&s->d_col, (size_t)E)) || (rc = dev_alloc(&s->d_rowdeg, (size_t)std::max(M, 1))) ||
      (rc = dev_alloc(&s->d_label, (size_t)M)) || (rc = dev_alloc(&s->d_pos_of, (size_t)n)) || (rc = dev_alloc(&s->d_dense_of, (size_t)n)) ||
      (rc = dev_alloc(&s->d_colour, (size_t)n)) || (rc = dev_alloc(&s->d_buf[0], s->buf_bytes, ipc)) ||
      (rc = dev_alloc(&s->d_buf[1], s->buf_bytes, ipc)) || (rc = dev_alloc(&s->d_queue[0], (size_t)M)) ||
      (rc = dev_alloc(&s->d_queue[1], (size_t)M)) || (rc = dev_alloc(&s->d_queue[2], (size_t)M)) || (rc = dev_alloc(&s->d_queue[3], (size_t)M)) ||
      (rc = dev_alloc(&s->d_ncand, (size_t)M)) || (rc = dev_alloc(&s->d_state, 1)) ||
      (rc = dev_alloc(&s->d_final_stats, 2))) {
    session_free(s);
    return rc;
  }
  cudaStream_t st = s->stream;
#define UP(dst, src, bytes)                                                                          \
  do {                                                                                               \
    cudaError_t _e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);                   \
    if (_e != cudaSuccess) { session_free(s); return fail(PPRB200_ERR_CUDA, "H2D copy failed: %s", cudaGetErrorString(_e)); } \
  } while (0)
#define SHARED(field, host, bytes)                                                                                      \
  do {                                                                                                                  \
    if (src) {                                                                                                          \
      cudaError_t _e = cudaMemcpyPeerAsync(s->field, s->device, src->field, src->device, bytes, st);                    \
      if (_e != cudaSuccess) { session_free(s); return fail(PPRB200_ERR_CUDA, "peer copy failed: %s", cudaGetErrorString(_e)); } \
    } else UP(s->field, host, bytes);                                                                                   \
  } while (0)
  SHARED(d_row_off, H.row_off.data(), ((size_t)M + 1) * sizeof(long long));
  if (M) SHARED(d_rowdeg, H.rowdeg.data(), (size_t)M * sizeof(unsigned long long));
  if (E && !src && H.d_enc) {  // column words made on the device (build_host_plan)
    cudaError_t _e = cudaMemcpyPeerAsync(s->d_col, s->device, H.d_enc, H.enc_device, (size_t)E * sizeof(uint32_t), st);
    if (_e != cudaSuccess) { session_free(s); return fail(PPRB200_ERR_CUDA, "device copy of the column words failed: %s", cudaGetErrorString(_e)); }
  } else if (E) SHARED(d_col, H.enc.data(), (size_t)E * sizeof(uint32_t));
  if (M) SHARED(d_label, H.label.data(), (size_t)M * sizeof(int));
  if (n) SHARED(d_dense_of, H.dense_of.data(), (size_t)n * sizeof(int));
  if (n) SHARED(d_pos_of, H.pos_of.data(), (size_t)n * sizeof(int));
  if (n) SHARED(d_colour, H.colour.data(), (size_t)n);
#undef SHARED
  s->pos_split = H.cls_begin[1][0];
  if (world > 1) {
    if ((rc = dev_alloc(&s->d_seq_list, R.seq_list.size()))) { session_free(s); return rc; }
    if (!R.seq_list.empty()) UP(s->d_seq_list, R.seq_list.data(), R.seq_list.size() * sizeof(int));
    if (M > 0 && !H.need_mask.empty() && !getenv("PPRB200_NO_NEED_MASK")) {
      if ((rc = dev_alloc(&s->d_need, H.need_mask.size())) || (rc = dev_alloc(&s->d_owner, (size_t)M))) { session_free(s); return rc; }
      UP(s->d_need, H.need_mask.data(), H.need_mask.size());
      UP(s->d_owner, H.owner8.data(), (size_t)M);
    }
  }
  if (s->n_items > 0) {
    // global-table pool of the order-free path: one table per CTA that can be in flight, sized for the worst case of
    // its class (big: the largest hub; mid: out-degree <= mid_deg) so that an acquired table can never overflow
    const int Lpm = Lp;
    auto pow2cap = [&](unsigned long long deg) {
      const unsigned long long worst = std::min<unsigned long long>(2ull * (deg * Lpm + 2ull), 2ull * ((unsigned long long)n + 1ull));
      unsigned int cap = 1024;
      while ((unsigned long long)cap < worst) cap <<= 1;
      return cap;
    };
    const unsigned long long per_slot = sizeof(GSlot) + sizeof(unsigned int) + 4 + 2;  // slot, list entry, compact score (cap/2 x 8), label (cap/2 x 4)
    s->tbl_cap[1] = pow2cap((unsigned long long)s->max_deg_par);
    s->tbl_cap[0] = std::min(s->tbl_cap[1], pow2cap((unsigned long long)s->mid_deg));
    s->tbl_count_cls[1] = s->sm_count + 8;      // >= CTAs in flight of the big instantiation
    s->tbl_count_cls[0] = s->sm_count * 3 + 8;  // ... of the mid instantiation
    // merge_dense_kernel in front: merge_par only sees the init pass (multiplicities: small tables), split hubs and the
    // hand-overs (a CTA that finds no table free waits for one): 68 GB of pool become 46
    if (dense_enabled()) s->tbl_count_cls[1] = 104;
    s->tbl_first[1] = 0;
    s->tbl_first[0] = s->tbl_count_cls[1];
    s->pool_off[1] = 0;
    s->pool_off[0] = (size_t)s->tbl_count_cls[1] * s->tbl_cap[1] * per_slot;
    const size_t pool_bytes = s->pool_off[0] + (size_t)s->tbl_count_cls[0] * s->tbl_cap[0] * per_slot;
    const int n_tables = s->tbl_count_cls[0] + s->tbl_count_cls[1];
    if ((rc = dev_alloc(&s->d_item_pos, (size_t)s->n_items)) || (rc = dev_alloc(&s->d_item_off, (size_t)s->n_items)) ||
        (rc = dev_alloc(&s->d_item_len, (size_t)s->n_items)) || (rc = dev_alloc(&s->d_pool, pool_bytes)) ||
        (rc = dev_alloc(&s->d_tbl_inuse, (size_t)n_tables)) || (rc = dev_alloc(&s->d_tbl_count, (size_t)n_tables)) ||
        (rc = dev_alloc(&s->d_node_tbl, (size_t)M)) || (rc = dev_alloc(&s->d_node_done, (size_t)M)) ||
        (rc = dev_alloc(&s->d_fb_queue, (size_t)s->n_items))) {
      session_free(s);
      return rc;
    }
    for (int c = 0; c < 2; c++) {
      s->n_team_nodes[c] = R.n_team_nodes[c]; s->team_first[c] = R.team_first[c]; s->team_count[c] = R.team_count[c];
      s->team_item_begin[c] = R.team_item_begin[c]; s->team_item_end[c] = R.team_item_end[c];
    }
    if (!R.teams.empty()) {
      const size_t nti = R.team_item_pos.size(), nt = R.teams.size();
      // (sized for the 512-thread instantiation: the larger pass-2 allowance per member)
      s->stage_bytes = (team_stage_bytes<8192, 16384, 4096, 512>() + 255) & ~(size_t)255;
      if ((rc = dev_alloc(&s->d_team_item_pos, nti)) || (rc = dev_alloc(&s->d_team_item_off, nti)) || (rc = dev_alloc(&s->d_team_item_len, nti)) ||
          (rc = dev_alloc(&s->d_team_item_team, nti)) || (rc = dev_alloc(&s->d_teams, nt)) || (rc = dev_alloc(&s->d_team_hdr, nt)) ||
          (rc = dev_alloc(&s->d_stage, nt * s->stage_bytes))) {
        session_free(s);
        return rc;
      }
      UP(s->d_team_item_pos, R.team_item_pos.data(), nti * sizeof(int));
      UP(s->d_team_item_off, R.team_item_off.data(), nti * sizeof(long long));
      UP(s->d_team_item_len, R.team_item_len.data(), nti * sizeof(int));
      UP(s->d_team_item_team, R.team_item_team.data(), nti * sizeof(int));
      UP(s->d_teams, R.teams.data(), nt * sizeof(TeamInfo));
      cudaMemsetAsync(s->d_team_hdr, 0, nt * sizeof(TeamHeader), st);
      cudaMemsetAsync(s->d_stage, 0, nt * s->stage_bytes, st);  // (the finishing member of a team leaves its staging area clean)
    }
    UP(s->d_item_pos, R.item_pos.data(), (size_t)s->n_items * sizeof(int));
    UP(s->d_item_off, R.item_off.data(), (size_t)s->n_items * sizeof(long long));
    UP(s->d_item_len, R.item_len.data(), (size_t)s->n_items * sizeof(int));
    cudaMemsetAsync(s->d_tbl_inuse, 0, (size_t)n_tables * sizeof(unsigned int), st);
    cudaMemsetAsync(s->d_tbl_count, 0, (size_t)n_tables * sizeof(unsigned int), st);
    cudaMemsetAsync(s->d_node_tbl, 0, (size_t)M * sizeof(unsigned int), st);
    cudaMemsetAsync(s->d_node_done, 0, (size_t)M * sizeof(unsigned int), st);
    for (int cls = 0; cls < 2; cls++)  // pool tables start empty: keys -1, accumulators 0
      pool_init_kernel<<<s->sm_count * 8, 256, 0, st>>>(s->d_pool + s->pool_off[cls], (size_t)s->tbl_cap[cls] * per_slot, s->tbl_cap[cls],
                                                        s->tbl_count_cls[cls]);
  }
#undef UP
  cudaError_t e = cudaStreamSynchronize(st);  // (the plan's vectors may go out of scope)
  if (e != cudaSuccess) { session_free(s); return fail(PPRB200_ERR_CUDA, "upload failed: %s", cudaGetErrorString(e)); }
  cudaMemsetAsync(s->d_ncand, 0, (size_t)std::max(M, 1) * sizeof(int), st);
  if ((rc = dev_alloc(&s->d_mbox, (size_t)2 * MAX_WORLD, ipc))) { session_free(s); return rc; }
  cudaMemsetAsync(s->d_mbox, 0, sizeof(Mailbox) * 2 * MAX_WORLD, st);
  cudaMemsetAsync(s->d_state, 0, sizeof(RunState), st);
  cudaEventCreate(&s->ev_begin);
  cudaEventCreate(&s->ev_end);
  cudaEventCreate(&s->ev_walk[0]);
  cudaEventCreate(&s->ev_walk[1]);
  for (int i = 0; i < 2; i++) {
    cudaStreamCreateWithFlags(&s->aux[i], cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&s->ev_join[i], cudaEventDisableTiming);
  }
  cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming);
  s->cur = s->stream;
  if (const char* ev = getenv("PPRB200_OVERLAP")) s->overlap = atoi(ev) != 0;
  s->use_dense = dense_enabled();
  if (const char* ev = getenv("PPRB200_DENSE_THREADS")) s->dense_threads = atoi(ev) == 1024 ? 1024 : 512;
  if (getenv("PPRB200_PROF")) {
    if ((rc = dev_alloc(&s->d_prof, (size_t)2 * s->sm_count * 8 * 8))) { session_free(s); return rc; }
    cudaMemsetAsync(s->d_prof, 0, (size_t)2 * s->sm_count * 8 * 8 * sizeof(unsigned long long), st);
  }
  s->h2d_ms = now_ms() - t1;
  if (getenv("PPRB200_HOST_TIMING")) fprintf(stderr, "[pprb200] rank %d: device setup + H2D %.2f ms\n", rank, s->h2d_ms);
  *out = s;
  return PPRB200_OK;
}

static int session_create_impl(const int64_t* row_ptr, const int32_t* col, int32_t n, const uint8_t* colour_in,
                               uint32_t max_L, uint32_t hub_threshold, int32_t rank, int32_t world, void* stream,
                               pprb200_session** out, bool need_colour = true, const HostPlanOut* plan_out = nullptr) {
  if (!out) return fail(PPRB200_ERR_PARAM, "out is NULL");
  *out = nullptr;
  if (max_L == 0) return fail(PPRB200_ERR_PARAM, "L must be positive");
  if (world < 1 || world > MAX_WORLD || rank < 0 || rank >= world) return fail(PPRB200_ERR_PARAM, "rank %d / world %d: need 0 <= rank < world <= %d", rank, world, MAX_WORLD);
  int rc = validate_csr(row_ptr, col, n);
  if (rc) return rc;
  if (!plan_out && (rc = device_ok())) return rc;
  HostPlan H;
  if ((rc = build_host_plan(row_ptr, col, n, colour_in, hub_threshold, world, need_colour, H, /*use_device=*/plan_out == nullptr))) return rc;
  RankPlan R;
  build_rank_plan(H, rank, R);
  if (plan_out) {
    const HostPlanOut& o = *plan_out;
    const int32_t M = H.M;
    if (o.pos_of) std::memcpy(o.pos_of, H.pos_of.data(), sizeof(int32_t) * (size_t)n);
    if (o.rank_of) std::memcpy(o.rank_of, H.rank_of.data(), sizeof(int32_t) * (size_t)n);
    if (o.row_off) for (int32_t p = 0; p <= M; p++) o.row_off[p] = H.row_off[(size_t)p];
    if (o.enc && H.E) std::memcpy(o.enc, H.enc.data(), sizeof(uint32_t) * (size_t)H.E);
    const int32_t n_items = (int32_t)R.item_pos.size();
    const int32_t ni = std::min<int32_t>(n_items, o.item_cap);
    for (int32_t i = 0; i < ni; i++) {
      if (o.item_pos) o.item_pos[i] = R.item_pos[(size_t)i];
      if (o.item_off) o.item_off[i] = R.item_off[(size_t)i];
      if (o.item_len) o.item_len[i] = R.item_len[(size_t)i];
    }
    if (o.summary) {
      int32_t* q = o.summary;
      q[0] = M; q[1] = n_items; q[2] = H.chunk; q[3] = H.mid_deg;
      for (int c = 0; c < 2; c++) { q[4 + c] = R.range_begin[c]; q[6 + c] = R.range_end[c]; }
      for (int c = 0; c < 2; c++)
        for (int k = 0; k < 2; k++) { q[8 + 2 * c + k] = R.item_begin[c][k]; q[12 + 2 * c + k] = R.item_end[c][k]; }
    }
    return PPRB200_OK;
  }
  return session_from_plan(H, R, max_L, rank, stream, /*ipc=*/world > 1, out);
}

// ---- stage configuration -----------------------------------------------------------------------
static int stage_limit(int cap, int Lp) {
  const int guard = std::max(Lp, 32) + 1;
  return std::min(cap * 3 / 4, cap - guard);
}

static unsigned int next_pow2(unsigned long long x) {
  unsigned long long p = 1;
  while (p < x) p <<= 1;
  return (unsigned int)std::min<unsigned long long>(p, 1ull << 31);
}

template <int CAP, int WARPS, typename IdxT>
static cudaError_t launch_stage(pprb200_session* s, const MergeParams& P, int grid, size_t smem, unsigned char* ws,
                                unsigned int ws_cap, int ws_identity) {
  static bool configured[64] = {false};  // cudaFuncSetAttribute is per device
  if (!configured[s->device & 63]) {
    cudaError_t e = cudaFuncSetAttribute(merge_seq_kernel<CAP, WARPS, IdxT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured[s->device & 63] = true;
  }
  merge_seq_kernel<CAP, WARPS, IdxT><<<grid, WARPS * 32, smem, s->cur>>>(P, ws, ws_cap, ws_identity);
  s->launch_count++;
  return cudaGetLastError();
}

// Enqueue the table-size cascade for one colour (or, MC, for everything): 1024 -> 2048 -> 4096 -> 16384 -> global.
static int enqueue_cascade(pprb200_session* s, MergeParams P, int range_begin, int range_end, int L, int range_split = -1) {
  const int Lp = roundup4(L);
  P.Lp = Lp;
  P.L = L;
  // the low-degree tail of the range on 512-slot tables (27 warps per SM instead of 14), if those can never overflow there
  if (range_split < range_begin || range_split > range_end || stage_limit(512, Lp) < SEQ_SMALL_DEG * Lp + 1 || getenv("PPRB200_NO_SMALL_STAGE")) range_split = range_end;
  const int caps[4] = {1024, 2048, 4096, 16384};
  const int warps[4] = {14, 7, 3, 1};
  bool have_source = false;  // false: next stage reads the range; true: reads queue `qsrc`
  int qsrc = -1;
  int work_idx = 0;
  for (int i = 0; i < 4; i++) {
    const int limit = stage_limit(caps[i], Lp);
    if (limit < 8) continue;
    MergeParams Q = P;
    Q.limit = limit;
    Q.work_idx = work_idx++;
    const bool first = !have_source;
    if (first) { Q.range_begin = range_begin; Q.range_end = range_split; Q.queue_in = nullptr; Q.queue_in_idx = -1; }
    else { Q.queue_in = s->d_queue[qsrc]; Q.queue_in_idx = qsrc; }
    const int qdst = qsrc + 1;
    Q.queue_out = s->d_queue[qdst];
    Q.queue_out_idx = qdst;
    const size_t smem = (size_t)warps[i] * ((size_t)caps[i] * 14 + 1040);
    cudaError_t e = cudaSuccess;
    if (first && range_split < range_end) {
      MergeParams T = Q;
      T.range_begin = range_split; T.range_end = range_end;
      T.limit = stage_limit(512, Lp);
      T.work_idx = 5;
      e = launch_stage<512, 27, unsigned short>(s, T, s->sm_count, (size_t)27 * ((size_t)512 * 14 + 1040), nullptr, 0, 0);
      if (e != cudaSuccess) return fail(PPRB200_ERR_CUDA, "merge stage (512-slot tables) launch failed: %s", cudaGetErrorString(e));
    }
    if (first && range_split == range_begin) { have_source = true; qsrc = qdst; continue; }  // (nothing but low-degree nodes)
    if (i == 0) e = launch_stage<1024, 14, unsigned short>(s, Q, s->sm_count, smem, nullptr, 0, 0);
    else if (i == 1) e = launch_stage<2048, 7, unsigned short>(s, Q, s->sm_count, smem, nullptr, 0, 0);
    else if (i == 2) e = launch_stage<4096, 3, unsigned short>(s, Q, s->sm_count, smem, nullptr, 0, 0);
    else e = launch_stage<16384, 1, unsigned short>(s, Q, s->sm_count, smem, nullptr, 0, 0);
    if (e != cudaSuccess) return fail(PPRB200_ERR_CUDA, "merge stage %d launch failed: %s", i, cudaGetErrorString(e));
    have_source = true;
    qsrc = qdst;
  }
  // final stage: table in the global workspace, sized for the worst case of this run -> cannot overflow
  {
    MergeParams Q = P;
    Q.limit = 0x7fffffff;
    Q.work_idx = work_idx++;
    if (!have_source) { Q.range_begin = range_begin; Q.range_end = range_end; Q.queue_in = nullptr; Q.queue_in_idx = -1; }
    else { Q.queue_in = s->d_queue[qsrc]; Q.queue_in_idx = qsrc; }
    Q.queue_out = nullptr;
    Q.queue_out_idx = 0;
    const unsigned long long bound = std::min<unsigned long long>((unsigned long long)s->max_deg_seq * (unsigned long long)Lp + 2ull + 32ull,
                                                                  (unsigned long long)s->n + 1ull);
    unsigned int cap;
    int identity;
    if ((unsigned long long)s->n <= 2 * bound) { cap = next_pow2((unsigned long long)std::max(s->n, 64)); identity = 1; }
    else { cap = next_pow2(2 * bound); identity = 0; }
    const size_t region = (size_t)cap * 16;  // vals 8 + keys 4 + list 4
    const size_t budget = (size_t)8 << 30;
    int gw = (int)std::max<size_t>(1, std::min<size_t>((size_t)s->sm_count * 8, budget / region));
    const int warps_g = 4;
    int grid = std::max(1, gw / warps_g);
    const size_t need = (size_t)grid * warps_g * region;
    if (need > s->ws_bytes) {
      g_alloc_stream = s->cur;  // stream-ordered with the launch below
      dev_free(s->d_ws);
      s->d_ws = nullptr;
      s->ws_bytes = 0;
      int rc = dev_alloc(&s->d_ws, need);
      g_alloc_stream = s->stream;
      if (rc) return rc;
      s->ws_bytes = need;
    }
    cudaError_t e = launch_stage<0, 4, unsigned int>(s, Q, grid, (size_t)warps_g * 1040, s->d_ws, cap, identity);
    if (e != cudaSuccess) return fail(PPRB200_ERR_CUDA, "global merge stage launch failed: %s", cudaGetErrorString(e));
  }
  return PPRB200_OK;
}

template <int H, int TCAP, int CMAX, int COLCAP, int R, int THREADS>
static cudaError_t launch_par(pprb200_session* s, const ParParams& P, int grid) {
  const size_t smem = par_smem_bytes<H, TCAP, CMAX, COLCAP, R>() + 8 + par_queue_bytes(THREADS);
  static bool configured[64] = {false};
  if (!configured[s->device & 63]) {
    cudaError_t e = cudaFuncSetAttribute(merge_par_kernel<H, TCAP, CMAX, COLCAP, R, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured[s->device & 63] = true;
  }
  merge_par_kernel<H, TCAP, CMAX, COLCAP, R, THREADS><<<grid, THREADS, smem, s->cur>>>(P);
  s->launch_count++;
  return cudaGetLastError();
}

template <int H, int R, int TCAP, int CMAX, int COLCAP, int THREADS, int MINB, bool TEAMS = false>
static cudaError_t launch_dense(pprb200_session* s, const DenseParams& P, int per_sm) {
  const size_t smem = dense_smem_bytes<H, R, TCAP, CMAX, COLCAP>();
  static bool configured[64] = {false};
  if (!configured[s->device & 63]) {
    cudaError_t e = cudaFuncSetAttribute(merge_dense_kernel<H, R, TCAP, CMAX, COLCAP, THREADS, MINB, TEAMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured[s->device & 63] = true;
  }
  const int work = P.n_items + (TEAMS ? P.n_team_items : 0);
  if (work <= 0) return cudaSuccess;
  merge_dense_kernel<H, R, TCAP, CMAX, COLCAP, THREADS, MINB, TEAMS><<<std::min(s->sm_count * per_sm, work), THREADS, smem, s->cur>>>(P);
  s->launch_count++;
  return cudaGetLastError();
}

// Enqueue the order-free path for the nodes of colour c above the hub threshold.
//   init (grank.h:64-83): merge_par_kernel on both classes (multiplicities, one pass over the column words).
//   iterations / MC combine rounds: merge_dense_kernel -- 1024-thread CTAs (1 per SM, H = 8192, R = 16384) for the big
//   class, 512-thread CTAs (2 per SM, H = 4096, R = 8192) for the mid class -- with merge_par_kernel behind it for the chunks of split
//   hubs (launched alongside) and for whatever the dense kernels hand over through the device queue (launched after).
static int enqueue_par(pprb200_session* s, const MergeParams& M, int c, int L) {
  ParParams P;
  std::memset(&P, 0, sizeof(P));
  P.M = M;
  P.M.Lp = roundup4(L);
  P.M.L = L;
  P.M.colour = c;
  P.chunk = s->chunk;
  P.node_tbl = s->d_node_tbl;
  P.node_done = s->d_node_done;
  P.n_ids = s->n;
  P.use_sketch = 1;  // two-pass merge for single-item nodes of the big class (PPRB200_SKETCH=0: single pass, for A/B runs)
  if (const char* e = getenv("PPRB200_SKETCH")) P.use_sketch = atoi(e) != 0;
  const bool dense = s->use_dense && !M.init_mode;
  auto par_class = [&](int cls) {  // pool tables of a class
    P.pool = s->d_pool + s->pool_off[cls];
    P.capmax = s->tbl_cap[cls];
    P.tbl_bytes = (size_t)s->tbl_cap[cls] * (sizeof(GSlot) + sizeof(unsigned int) + 4 + 2);
    P.n_tables = s->tbl_count_cls[cls];
    P.tbl_inuse = s->d_tbl_inuse + s->tbl_first[cls];
    P.tbl_count = s->d_tbl_count + s->tbl_first[cls];
  };
  cudaStream_t side = s->overlap ? s->aux[0] : s->stream;
  for (int cls = 1; cls >= 0; cls--) {
    const int b = s->item_begin[c][cls], e = s->item_end[c][cls];
    if (e == b) continue;
    const int n_team = (dense && cls == 1) ? s->n_team_nodes[c] : 0;  // leading items that a team of CTAs works on (chunk items below)
    const int n_hub = (dense && cls == 1) ? std::min(s->hub_items[c], e - b) : 0;
    if (!dense || n_hub > 0) {
      const int nb = dense ? n_hub : e - b;
      P.item_pos = s->d_item_pos + b;
      P.item_begin = s->d_item_off + b;
      P.item_len = s->d_item_len + b;
      P.n_items = nb;
      P.item_queue = nullptr;
      P.work_idx = 6 + cls;  // (0..4: the exact-order cascade)
      par_class(cls);
      P.prof = (s->d_prof && !dense) ? s->d_prof + (size_t)cls * s->sm_count * 8 * 8 : nullptr;
      s->cur = (dense || cls == 0) ? side : s->stream;
      cudaError_t err = cls == 1 ? launch_par<8192, 2048, 3072, PAR_CHUNK_MAX, 8192, 512>(s, P, std::min(s->sm_count, nb))
                                 : launch_par<2048, 2048, 2048, PAR_MID_MAX, 0, 128>(s, P, std::min(s->sm_count * 3, nb));
      if (err != cudaSuccess) return fail(PPRB200_ERR_CUDA, "merge_par launch failed: %s", cudaGetErrorString(err));
    }
    if (dense && e - b > n_hub) {
      const int skip = std::max(n_hub, n_team);  // (hubs split for merge_par and team hubs exclude each other: see build_rank_plan)
      DenseParams D;
      std::memset(&D, 0, sizeof(D));
      D.M = P.M;
      D.item_pos = s->d_item_pos + b + skip;
      D.item_begin = s->d_item_off + b + skip;
      D.item_len = s->d_item_len + b + skip;
      D.n_items = e - b - skip;
      D.item_base = b + skip;
      if (n_team > 0) {
        const int tb = s->team_item_begin[c];
        D.team_item_pos = s->d_team_item_pos + tb;
        D.team_item_begin = s->d_team_item_off + tb;
        D.team_item_len = s->d_team_item_len + tb;
        D.team_item_team = s->d_team_item_team + tb;
        D.n_team_items = s->team_item_end[c] - tb;
        D.teams = s->d_teams;
        D.team_hdr = s->d_team_hdr;
        D.stage = s->d_stage;
        D.stage_bytes = s->stage_bytes;
        cudaMemsetAsync(s->d_team_hdr + s->team_first[c], 0, (size_t)s->team_count[c] * sizeof(TeamHeader), cls == 0 ? side : s->stream);
      }
      D.chunk = s->chunk;
      D.work_idx = 8 + cls;
      D.fb_queue = s->d_fb_queue;
      D.fb_idx = 4;
      {
        // (1: every node with an old basket at all stays on merge_dense_kernel -- the exact-candidate bound covers baskets
        // that are not full yet, R-MAT-22 job 1 844 -> 1 781 ms; PPRB200_MIN_OLD=0 restores the hand-over of not-full baskets)
        const char* mo = getenv("PPRB200_MIN_OLD");
        const int min_old_env = mo ? atoi(mo) : 1;
        D.min_old = min_old_env >= 1 ? std::min(min_old_env, L) : L;
        const char* tl = getenv("PPRB200_TAIL_LIMIT");  // tests: force the pass-2 rounds on small graphs
        D.tail_limit = tl ? atoi(tl) : 0;
      }
      D.prof = s->d_prof ? s->d_prof + (size_t)cls * s->sm_count * 8 * 8 : nullptr;
      s->cur = cls == 0 ? side : s->stream;
      cudaError_t err;
      if (cls == 1) {
        err = cudaSuccess;
        if (D.n_team_items > 0) {  // the hubs' chunk items first, on the instantiation that knows about teams
          DenseParams T = D;
          T.n_items = 0;
          err = s->dense_threads == 512 ? launch_dense<8192, 16384, 4096, 2048, 1024, 512, 1, true>(s, T, 1)
                                        : launch_dense<8192, 16384, 4096, 2048, 1024, 1024, 1, true>(s, T, 1);
          D.n_team_items = 0;
          D.work_idx = 11;  // (its own work counter)
        }
        if (err == cudaSuccess)
          err = s->dense_threads == 512 ? launch_dense<8192, 16384, 4096, 2048, 1024, 512, 1>(s, D, 1)
                                        : launch_dense<8192, 16384, 4096, 2048, 1024, 1024, 1>(s, D, 1);
      } else {
        static const int cfg = getenv("PPRB200_MID_CFG") ? atoi(getenv("PPRB200_MID_CFG")) : 0;  // A/B hook
        if (cfg == 3) err = launch_dense<2048, 8192, 1024, 512, PAR_MID_MAX, 256, 3>(s, D, 3);
        else if (cfg == 5) err = launch_dense<1024, 4096, 512, 512, PAR_MID_MAX, 128, 5>(s, D, 5);
        else if (cfg == 1) err = launch_dense<4096, 8192, 2048, 1024, PAR_MID_MAX, 512, 2>(s, D, 2);
        else err = launch_dense<4096, 8192, 2048, 1024, PAR_MID_MAX, 256, 2>(s, D, 2);  // measured best on R-MAT-22 (profiles/r2/sweeps.txt)
      }
      if (err != cudaSuccess) return fail(PPRB200_ERR_CUDA, "merge_dense launch failed: %s", cudaGetErrorString(err));
    }
  }
  if (dense && s->item_end[c][1] > s->item_begin[c][0]) {
    // whatever the dense kernels handed over (first updates of low-degree nodes, a few overflows): merge_par from the queue, on
    // its 512-thread two-pass instantiation (measured: the 128-thread one is slower on these, profiles/r2/sweeps.txt)
    if (s->overlap) {
      cudaEventRecord(s->ev_join[0], side);
      cudaStreamWaitEvent(s->stream, s->ev_join[0], 0);
    }
    P.item_pos = s->d_item_pos;
    P.item_begin = s->d_item_off;
    P.item_len = s->d_item_len;
    P.n_items = 0;
    P.prof = nullptr;
    P.item_queue = s->d_fb_queue;
    P.queue_idx = 4;
    P.work_idx = 10;
    par_class(1);
    s->cur = s->stream;
    cudaError_t err = launch_par<8192, 2048, 3072, PAR_CHUNK_MAX, 8192, 512>(s, P, std::min(s->sm_count, s->tbl_count_cls[1]));
    if (err != cudaSuccess) return fail(PPRB200_ERR_CUDA, "merge_par (hand-over queue) launch failed: %s", cudaGetErrorString(err));
  }
  s->cur = s->stream;
  return PPRB200_OK;
}

// One colour's worth of merge work: big class on the session stream, mid class and the exact-order cascade on the two
// auxiliary streams (fork / join through events), so that the three persistent kernels share the SMs as CTAs retire.
static int enqueue_colour(pprb200_session* s, const MergeParams& Q, int c, int L) {
  int rc;
  if (s->overlap) {
    cudaEventRecord(s->ev_fork, s->stream);
    cudaStreamWaitEvent(s->aux[0], s->ev_fork, 0);
    cudaStreamWaitEvent(s->aux[1], s->ev_fork, 0);
  }
  if ((rc = enqueue_par(s, Q, c, L))) return rc;
  s->cur = s->overlap ? s->aux[1] : s->stream;
  if (s->range_end[c] > s->range_begin[c] && (rc = enqueue_cascade(s, Q, s->range_begin[c], s->range_end[c], L, s->range_split[c]))) return rc;
  s->cur = s->stream;
  if (s->overlap) {
    for (int i = 0; i < 2; i++) {
      cudaEventRecord(s->ev_join[i], s->aux[i]);
      cudaStreamWaitEvent(s->stream, s->ev_join[i], 0);
    }
  }
  return PPRB200_OK;
}

static int ensure_outputs(pprb200_session* s, uint32_t K) {
  const size_t need = (size_t)std::max(s->n, 1) * K;
  if (need > s->out_capacity) {
    dev_free(s->d_out_ids); dev_free(s->d_out_scores);
    s->d_out_ids = nullptr; s->d_out_scores = nullptr; s->out_capacity = 0;
    int rc;
    if ((rc = dev_alloc(&s->d_out_ids, need)) || (rc = dev_alloc(&s->d_out_scores, need))) return rc;
    s->out_capacity = need;
  }
  if (!s->d_out_cnt) { int rc = dev_alloc(&s->d_out_cnt, (size_t)std::max(s->n, 1)); if (rc) return rc; }
  return PPRB200_OK;
}

static int enqueue_final(pprb200_session* s, int L, uint32_t K, double sink_score) {
  if (s->world > 1 && s->peers.need && s->peers.push_mode != 3 && s->M > 0) {
    // complete every rank's copy of the result (see final_push_kernel), then wait until everybody's pushes have landed
    final_push_kernel<<<s->sm_count * 8, 256, 0, s->stream>>>(s->d_state, s->peers, s->d_owner, s->M, s->pos_split, roundup4(L));
    phase_end_kernel<<<1, 1, 0, s->stream>>>(s->d_state, 0, 0, s->peers, 1);
    s->launch_count += 2;
  }
  const int Lp = roundup4(L);
  const int warps = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)(200 * 1024) / ((size_t)Lp * 12)));
  const size_t smem = (size_t)warps * Lp * 12;
  static size_t configured[64] = {0};
  if (smem > 48 * 1024 && smem > configured[s->device & 63]) {
    cudaError_t e = cudaFuncSetAttribute(final_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(PPRB200_ERR_CUDA, "final_topk smem %zu: %s", smem, cudaGetErrorString(e));
    configured[s->device & 63] = smem;
  }
  cudaMemsetAsync(s->d_final_stats, 0, 2 * sizeof(unsigned long long), s->stream);
  const int grid = std::max(1, std::min((s->n + warps - 1) / warps, s->sm_count * 8));
  final_topk_kernel<<<grid, warps * 32, smem, s->stream>>>(s->d_pos_of, s->d_dense_of, s->d_colour, s->d_buf[0], s->d_buf[1], s->d_state,
                                                           s->n, Lp, (int)K, sink_score, s->d_out_ids, s->d_out_scores,
                                                           s->d_out_cnt, s->d_final_stats);
  s->launch_count++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(PPRB200_ERR_CUDA, "final_topk launch failed: %s", cudaGetErrorString(e));
  return PPRB200_OK;
}

static int check_params(uint32_t K, uint32_t L, uint32_t iterations, double damping) {
  if (K == 0) return fail(PPRB200_ERR_PARAM, "K must be positive");
  if (L == 0) return fail(PPRB200_ERR_PARAM, "L must be positive");
  if (K > L) return fail(PPRB200_ERR_PARAM, "K must be <= L");
  if (iterations == 0) return fail(PPRB200_ERR_PARAM, "iterations must be positive");
  if (!(damping >= 0 && damping <= 1)) return fail(PPRB200_ERR_PARAM, "damping must be [0,1]");
  return PPRB200_OK;
}

static int session_grank_impl(pprb200_session* s, uint32_t K, uint32_t L, uint32_t iterations, double damping,
                              double tolerance) {
  int rc = check_params(K, L, iterations, damping);
  if (rc) return rc;
  if (L > s->max_L) return fail(PPRB200_ERR_PARAM, "L=%u exceeds the session's max_L=%u", L, s->max_L);
  L = effective_L(L, s->n);
  if ((size_t)roundup4((int)L) * 12 > 200 * 1024) return fail(PPRB200_ERR_PARAM, "min(L, n)=%u is above this build's limit of 17064", L);
  g_alloc_stream = s->stream;
  if ((rc = ensure_outputs(s, K))) return rc;
  if (s->world > 1 && !s->attached) return fail(PPRB200_ERR_STATE, "world=%d session: call pprb200_session_ipc_attach before running", s->world);
  s->last_mode = MODE_GRANK;
  s->last_K = K; s->last_L = L; s->last_iterations = iterations;
  s->merge_launches = 0;

  cudaStream_t st = s->stream;
  cudaEventRecord(s->ev_begin, st);
  state_reset_kernel<<<1, 1, 0, st>>>(s->d_state);
  s->launch_count = 1;
  if (s->world > 1) {  // nobody overwrites a buffer a slower peer is still reading for its previous run's top-K
    phase_end_kernel<<<1, 1, 0, st>>>(s->d_state, 0, 0, s->peers, 1);
    s->launch_count++;
  }
  MergeParams P;
  std::memset(&P, 0, sizeof(P));
  P.g.row_off = s->d_row_off; P.g.col = s->d_col; P.g.label = s->d_label; P.g.dense_of = s->d_dense_of;
  P.buf[0] = s->d_buf[0]; P.buf[1] = s->d_buf[1];
  P.st = s->d_state;
  P.mode = MODE_GRANK;
  P.damping = damping;
  P.self_grank = 1.0 - damping;
  P.ncand = s->d_ncand;
  P.peers = s->peers;
  P.work_list = s->d_seq_list;
  P.n_ids = s->n;
  if (s->M > 0) {
    // init (grank.h:64-83)
    cudaMemsetAsync(s->d_ncand, 0, (size_t)s->M * sizeof(int), st);
    for (int c = 0; c < 2; c++) {
      MergeParams Q = P;
      Q.init_mode = 1; Q.do_norm = 0; Q.colour = c;
      if ((rc = enqueue_colour(s, Q, c, (int)L))) return rc;
      phase_end_kernel<<<1, 1, 0, st>>>(s->d_state, 1, 0, s->peers, c == 1);
      s->launch_count++;
    }
  }
  // The loop is enqueued without a host round trip per iteration (the convergence test runs on the device and turns
  // the remaining launches into no-ops). A caller that passes a huge safety bound together with a tolerance should not
  // pay for millions of such no-ops: after every ENQUEUE_WINDOW iterations the host looks at RunState::active -- the
  // same value on every rank, it comes out of the barrier's max-reduction -- and stops enqueueing once it is clear.
  constexpr uint32_t ENQUEUE_WINDOW = 64;
  uint32_t enqueued = 0;
  for (uint32_t it = 0; it < iterations; it++) {
    const int c = (int)(it & 1);  // partitions.first on even iterations (grank.h:96,129)
    while (s->ev_merge.size() < 2 * (size_t)(it + 1)) { cudaEvent_t e; cudaEventCreate(&e); s->ev_merge.push_back(e); }
    cudaEventRecord(s->ev_merge[2 * it], st);
    {
      MergeParams Q = P;
      Q.init_mode = 0; Q.do_norm = 1; Q.colour = c;
      if ((rc = enqueue_colour(s, Q, c, (int)L))) return rc;
    }
    cudaEventRecord(s->ev_merge[2 * it + 1], st);
    iter_end_kernel<<<1, 1, 0, st>>>(s->d_state, c, tolerance, s->peers);
    s->launch_count++;
    enqueued = it + 1;
    if (tolerance >= 0 && enqueued % ENQUEUE_WINDOW == 0 && enqueued < iterations) {
      int active = 1;
      cudaError_t e = cudaMemcpyAsync(&active, &s->d_state->active, sizeof(int), cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) return fail(PPRB200_ERR_CUDA, "run failed: %s", cudaGetErrorString(e));
      if (!active) break;
    }
  }
  s->merge_launches = enqueued;
  if ((rc = enqueue_final(s, (int)L, K, 1.0 - damping))) return rc;
  cudaEventRecord(s->ev_end, st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(PPRB200_ERR_CUDA, "enqueue failed: %s", cudaGetErrorString(e));
  return PPRB200_OK;
}


// ---- MCCompletePathV2 (mccompletepathv2.h:182-258, north-star semantics) ---------------------------------------
template <bool GLOBAL, int THREADS>
static cudaError_t launch_walk_t(pprb200_session* s, const WalkParams& P, int grid, size_t smem) {
  static size_t configured[64] = {0};
  if (smem > 48 * 1024 && smem > configured[s->device & 63]) {
    cudaError_t e = cudaFuncSetAttribute(mc_walk_kernel<GLOBAL, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaFuncSetAttribute(mc_walk_kernel<GLOBAL, THREADS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    configured[s->device & 63] = smem;
  }
  mc_walk_kernel<GLOBAL, THREADS><<<grid, THREADS, smem, s->stream>>>(P);
  s->launch_count++;
  return cudaGetLastError();
}

// threads per source: a walk is sequential, so a CTA is busy for at least its longest walk (~ln(W)/(1-d) hops); with
// fewer threads per source the average thread works about as long as that and little of the CTA idles in the tail
template <bool GLOBAL>
static cudaError_t launch_walk(pprb200_session* s, const WalkParams& P, int grid, size_t smem, int threads) {
  return threads == 128 ? launch_walk_t<GLOBAL, 128>(s, P, grid, smem) : launch_walk_t<GLOBAL, 256>(s, P, grid, smem);
}

static uint32_t mc_coin_threshold(double damping) {
  const double t = std::floor(damping * 4294967296.0);
  if (t >= 4294967295.0) return 0xFFFFFFFFu;
  if (t <= 0) return 0u;
  return (uint32_t)t;
}

// MC scores are visits per walk: at most ~1/(1-d) (a walk's expected length), and MC_MAX_STEPS in the worst case. The
// order-free accumulator of the combine rounds is 2^-59 fixed point in 64 bits, i.e. it holds sums below 16. Up to this
// damping (1/(1-d) = 8) that leaves a factor of two; above it every node takes the exact-order fp64 path instead.
constexpr double MC_ORDER_FREE_MAX_DAMPING = 0.875;

static int session_mc_impl(pprb200_session* s, uint32_t K, uint32_t L, uint32_t R, double damping, uint64_t seed, uint32_t rounds) {
  int rc = check_params(K, L, R, damping);
  if (rc) return rc;
  if (damping > MC_ORDER_FREE_MAX_DAMPING && rounds > 0)
    for (int c = 0; c < 2; c++)
      if (s->item_end[c][1] > s->item_begin[c][0])
        return fail(PPRB200_ERR_PARAM,
                    "damping %g > %g with order-free nodes in the session: the 2^-59 fixed-point accumulator holds at most 16 visits per "
                    "walk; create the session with hub_threshold = UINT32_MAX (pprb200_mccompletepathv2 does so by itself)",
                    damping, MC_ORDER_FREE_MAX_DAMPING);
  if (L > s->max_L) return fail(PPRB200_ERR_PARAM, "L=%u exceeds the session's max_L=%u", L, s->max_L);
  L = effective_L(L, s->n);
  if ((size_t)roundup4((int)L) * 12 > 200 * 1024) return fail(PPRB200_ERR_PARAM, "min(L, n)=%u is above this build's limit of 17064", L);
  g_alloc_stream = s->stream;
  if ((rc = ensure_outputs(s, K))) return rc;
  if (s->world > 1 && !s->attached) return fail(PPRB200_ERR_STATE, "world=%d session: call pprb200_session_ipc_attach before running", s->world);
  s->last_mode = MODE_MC;
  s->last_K = K; s->last_L = L; s->last_iterations = rounds;
  while (s->ev_merge.size() < 2 * (size_t)std::max<uint32_t>(rounds, 1)) { cudaEvent_t e; cudaEventCreate(&e); s->ev_merge.push_back(e); }
  s->merge_launches = 0;
  const int Lp = roundup4((int)L);
  const unsigned long long W = (unsigned long long)((double)R * damping);  // mccompletepathv2.h:132

  cudaStream_t st = s->stream;
  cudaEventRecord(s->ev_begin, st);
  state_reset_kernel<<<1, 1, 0, st>>>(s->d_state);
  s->launch_count = 1;
  if (s->world > 1) {
    phase_end_kernel<<<1, 1, 0, st>>>(s->d_state, 0, 0, s->peers, 1);
    s->launch_count++;
  }
  cudaEventRecord(s->ev_walk[0], st);
  if (s->M > 0) {
    WalkParams P;
    std::memset(&P, 0, sizeof(P));
    P.g.row_off = s->d_row_off; P.g.col = s->d_col; P.g.label = s->d_label; P.g.dense_of = s->d_dense_of;
    P.buf[0] = s->d_buf[0]; P.buf[1] = s->d_buf[1];
    P.st = s->d_state;
    P.M = s->M; P.src_begin = 0; P.src_end = s->M; P.n_ids = s->n;
    P.colour = s->d_colour;
    P.peers = s->peers;
    P.rowdeg = s->d_rowdeg;
    P.Lp = Lp; P.L = (int)L; P.R = R; P.W = W; P.thresh = mc_coin_threshold(damping); P.seed = seed;
    // shared-memory table sized for the expected number of distinct visited nodes (<= hops + 1)
    const double len = damping >= 1.0 ? (double)MC_MAX_STEPS : std::min<double>((double)MC_MAX_STEPS, 1.0 / (1.0 - damping));
    const double expect = std::min<double>((double)s->n + 1.0, 1.0 + (double)W * len * 0.8);
    // capacities: powers of two and the 1.5x steps between them (the table is indexed by mulhi, not by a mask)
    const unsigned int caps[] = {1024u, 1536u, 2048u, 3072u, 4096u, 6144u, 8192u, 12288u, 16384u};
    unsigned int tcap = 16384u;
    for (const unsigned int c : caps)
      if ((double)c * 0.75 >= expect) { tcap = c; break; }
    int walk_threads = 256;  // measured on R-MAT-20, R=1000: 256 threads 19.8 G hops/s, 128 threads 12.3 G hops/s (fewer walks in flight)
    if (const char* e = getenv("PPRB200_WALK_THREADS")) walk_threads = atoi(e) == 256 ? 256 : 128;
    if (const char* e = getenv("PPRB200_WALK_TCAP")) {  // test hook: force the fallback path
      const unsigned int want = (unsigned int)std::max(1024, atoi(e));
      tcap = 16384u;
      for (const unsigned int c : caps)
        if (c >= want) { tcap = c; break; }
    }
    P.tcap = tcap; P.limit = tcap * 3 / 4 - 1;
    P.work_idx = 0; P.queue_in = nullptr; P.queue_in_idx = -1; P.queue_out = s->d_queue[0]; P.queue_out_idx = 0;
    const size_t smem = ((sizeof(WalkShared) + 15) & ~(size_t)15) + (size_t)tcap * (sizeof(WalkSlot) + sizeof(unsigned short));
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(walk_threads == 128 ? 16 : 8, (size_t)(227 * 1024) / (smem + 1024)));
    cudaError_t e = launch_walk<false>(s, P, std::max(1, std::min(s->M, s->sm_count * per_sm)), smem, walk_threads);
    if (e != cudaSuccess) return fail(PPRB200_ERR_CUDA, "mc_walk launch failed: %s", cudaGetErrorString(e));
    // fallback: sources that visited more distinct nodes than the shared table admits
    const unsigned long long worst = std::min<unsigned long long>((unsigned long long)s->n + 1ull, W * (unsigned long long)MC_MAX_STEPS + 1ull);
    if (worst > P.limit) {
      unsigned long long cap = 2048;
      while (cap < 2 * worst) cap <<= 1;
      const size_t per = (size_t)cap * (sizeof(WalkSlot) + sizeof(unsigned int));  // table + first-touch list
      const int grid = (int)std::max<size_t>(1, std::min<size_t>((size_t)s->sm_count, ((size_t)2 << 30) / per));
      if ((size_t)grid * per > s->walk_ws_bytes) {
        dev_free(s->d_walk_ws);
        s->d_walk_ws = nullptr; s->walk_ws_bytes = 0;
        if ((rc = dev_alloc(&s->d_walk_ws, (size_t)grid * per / sizeof(unsigned long long) + 1))) return rc;
        s->walk_ws_bytes = (size_t)grid * per;
      }
      WalkParams Q = P;
      Q.tcap = (unsigned int)cap; Q.limit = 0xffffffffu;
      Q.work_idx = 1; Q.queue_in = s->d_queue[0]; Q.queue_in_idx = 0; Q.queue_out = nullptr; Q.queue_out_idx = 0;
      Q.ws = s->d_walk_ws;
      e = launch_walk<true>(s, Q, grid, (sizeof(WalkShared) + 15) & ~(size_t)15, walk_threads);
      if (e != cudaSuccess) return fail(PPRB200_ERR_CUDA, "mc_walk fallback launch failed: %s", cudaGetErrorString(e));
    }
    phase_end_kernel<<<1, 1, 0, st>>>(s->d_state, 1, 0, s->peers, 1);
    s->launch_count++;
  }
  cudaEventRecord(s->ev_walk[1], st);

  // combine rounds (mccompletepathv2.h:211-250 as Jacobi sweeps over all nodes)
  MergeParams P;
  std::memset(&P, 0, sizeof(P));
  P.g.row_off = s->d_row_off; P.g.col = s->d_col; P.g.label = s->d_label; P.g.dense_of = s->d_dense_of;
  P.buf[0] = s->d_buf[0]; P.buf[1] = s->d_buf[1];
  P.st = s->d_state;
  P.mode = MODE_MC;
  P.damping = damping;
  P.self_grank = 1.0 - damping;
  P.ncand = s->d_ncand;
  P.peers = s->peers;
  P.work_list = s->d_seq_list;
  P.n_ids = s->n;
  if (s->M > 0 && rounds > 0) cudaMemsetAsync(s->d_ncand, 0, (size_t)s->M * sizeof(int), st);
  for (uint32_t r = 0; r < rounds; r++) {
    cudaEventRecord(s->ev_merge[2 * r], st);
    if (s->M > 0) {
      for (int c = 0; c < 2; c++) {
        MergeParams Q = P;
        Q.init_mode = 0; Q.do_norm = 0; Q.colour = c;
        if ((rc = enqueue_colour(s, Q, c, (int)L))) return rc;
        // both colours read the same (old) buffer: counters are cleared between the two cascades, slots flip at the end
        phase_end_kernel<<<1, 1, 0, st>>>(s->d_state, 0, c == 1 ? 1 : 0, s->peers, c == 1);
        s->launch_count++;
      }
    } else {
      phase_end_kernel<<<1, 1, 0, st>>>(s->d_state, 0, 1, s->peers, 0);
      s->launch_count++;
    }
    cudaEventRecord(s->ev_merge[2 * r + 1], st);
  }
  s->merge_launches = rounds;
  if ((rc = enqueue_final(s, (int)L, K, 1.0))) return rc;
  cudaEventRecord(s->ev_end, st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(PPRB200_ERR_CUDA, "enqueue failed: %s", cudaGetErrorString(e));
  return PPRB200_OK;
}

static int session_stats_impl(pprb200_session* s, pprb200_stats* out, bool with_final = true) {
  if (!out) return fail(PPRB200_ERR_PARAM, "stats is NULL");
  std::memset(out, 0, sizeof(*out));
  if (s->last_mode < 0) return fail(PPRB200_ERR_STATE, "no run has been enqueued on this session");
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  RunState h;
  unsigned long long fin[2];
  CUDA_TRY(cudaMemcpy(&h, s->d_state, sizeof(h), cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(fin, s->d_final_stats, sizeof(fin), cudaMemcpyDeviceToHost));
  if (h.peer_timeout) return fail(PPRB200_ERR_CUDA, "peer barrier timed out: a rank of this %d-GPU run died or never attached; results are invalid", s->world);
  out->iterations_run = (uint32_t)h.iter;
  out->n_gpus = (uint32_t)s->world;
  uint64_t ni = 0;
  if (s->last_mode == MODE_GRANK)
    for (int i = 0; i < h.iter; i++) ni += (uint64_t)s->colour_count[i & 1];
  else
    ni = (uint64_t)h.iter * (uint64_t)s->n;
  out->node_iterations = ni;
  out->nonsink_node_iterations = h.node_iters;
  out->edge_reads = h.edge_reads;
  out->merged_entries = h.merged;
  out->candidates = h.cands;
  out->truncations = h.truncs + (with_final ? fin[0] : 0);  // (every rank runs the final top-K over all nodes: counted once)
  out->boundary_ties = h.ties + (with_final ? fin[1] : 0);
  out->algorithmic_bytes = h.abytes;
  out->walk_steps = h.walk_steps;
  out->walks = h.walks;
  out->overflow_requeues = h.requeues;
  out->walk_algorithmic_bytes = h.walk_bytes;
  // maxDiff pair as grank.h leaves it: [0] = older, [1] = latest (after the swap of :140)
  out->max_diff[0] = h.m_prev < 0 ? 0.0 : (double)h.m_prev * NORM_INV;
  out->max_diff[1] = h.m_last < 0 ? 0.0 : (double)h.m_last * NORM_INV;
  float ms = 0;
  if (cudaEventElapsedTime(&ms, s->ev_begin, s->ev_end) == cudaSuccess) out->kernel_ms = ms;
  out->prep_ms = s->prep_ms;
  out->h2d_ms = s->h2d_ms;
  return PPRB200_OK;
}

static int session_fetch_impl(pprb200_session* s, int32_t* out_ids, double* out_scores, uint32_t* out_cnt) {
  if (s->last_mode < 0) return fail(PPRB200_ERR_STATE, "no run has been enqueued on this session");
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  if (s->world > 1) {
    int timed_out = 0;
    CUDA_TRY(cudaMemcpy(&timed_out, reinterpret_cast<const unsigned char*>(s->d_state) + offsetof(RunState, peer_timeout), sizeof(int), cudaMemcpyDeviceToHost));
    if (timed_out) return fail(PPRB200_ERR_CUDA, "peer barrier timed out: a rank of this %d-GPU run died or never attached; results are invalid", s->world);
  }
  const size_t cnt = (size_t)s->n * s->last_K;
  if (out_ids && cnt) CUDA_TRY(cudaMemcpyAsync(out_ids, s->d_out_ids, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, s->stream));
  if (out_scores && cnt) CUDA_TRY(cudaMemcpyAsync(out_scores, s->d_out_scores, cnt * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  if (out_cnt && s->n) CUDA_TRY(cudaMemcpyAsync(out_cnt, s->d_out_cnt, (size_t)s->n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  return PPRB200_OK;
}

// CUDA loads a kernel's code onto a device at its first launch there, and that load can wait for the device to go idle. In a
// one-process multi-GPU run the first kernels of rank 0 spin in the mailbox barrier until rank 1 -- enqueued by the same host
// thread, later -- arrives: a first-use load behind such a kernel would never return. So every kernel this library can launch
// is loaded on the current device up front (cudaFuncGetAttributes does that), before anything that waits for a peer is enqueued.
template <typename F>
static void preload(F* kernel) {
  cudaFuncAttributes attr;
  cudaFuncGetAttributes(&attr, reinterpret_cast<const void*>(kernel));
}
static void preload_kernels() {
  preload(state_reset_kernel); preload(pool_init_kernel); preload(phase_end_kernel); preload(iter_end_kernel); preload(final_topk_kernel); preload(final_push_kernel);
  preload(bfs_level_kernel); preload(encode_kernel); preload(need_mask_kernel);
  preload(merge_seq_kernel<512, 27, unsigned short>); preload(merge_seq_kernel<1024, 14, unsigned short>);
  preload(merge_seq_kernel<2048, 7, unsigned short>); preload(merge_seq_kernel<4096, 3, unsigned short>);
  preload(merge_seq_kernel<16384, 1, unsigned short>); preload(merge_seq_kernel<0, 4, unsigned int>);
  preload(merge_par_kernel<8192, 2048, 3072, PAR_CHUNK_MAX, 8192, 512>); preload(merge_par_kernel<2048, 2048, 2048, PAR_MID_MAX, 0, 128>);
  preload(merge_dense_kernel<8192, 16384, 4096, 2048, 1024, 512, 1>); preload(merge_dense_kernel<8192, 16384, 4096, 2048, 1024, 1024, 1>);
  preload(merge_dense_kernel<8192, 16384, 4096, 2048, 1024, 512, 1, true>); preload(merge_dense_kernel<8192, 16384, 4096, 2048, 1024, 1024, 1, true>);
  preload(merge_dense_kernel<2048, 8192, 1024, 512, PAR_MID_MAX, 256, 3>); preload(merge_dense_kernel<1024, 4096, 512, 512, PAR_MID_MAX, 128, 5>);
  preload(merge_dense_kernel<4096, 8192, 2048, 1024, PAR_MID_MAX, 512, 2>); preload(merge_dense_kernel<4096, 8192, 2048, 1024, PAR_MID_MAX, 256, 2>);
  preload(mc_walk_kernel<false, 128>); preload(mc_walk_kernel<false, 256>); preload(mc_walk_kernel<true, 128>); preload(mc_walk_kernel<true, 256>);
  cudaGetLastError();
}

static std::mutex g_api_mutex;  // one run at a time per process (SURVEY.md 8b re-entrancy)

template <int BT>
static void launch_exact_iter(const ExactParams& P, int grid, int sms, cudaStream_t st) {
  ppr_exact_iter_kernel<BT><<<grid, 256, 0, st>>>(P);
  if (P.n_heavy > 0) ppr_exact_heavy_kernel<BT><<<std::min(P.n_heavy, sms * 4), 256, 0, st>>>(P);
}

extern "C" {

int pprb200_device_count(void) {
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess) { cudaGetLastError(); return 0; }
  int ok = 0;
  for (int d = 0; d < cnt; d++) {
    cudaDeviceProp pr;
    if (cudaGetDeviceProperties(&pr, d) == cudaSuccess && pr.major == 10) ok++;
  }
  return ok;
}

int pprb200_session_create(const int64_t* row_ptr, const int32_t* col, int32_t n, const uint8_t* colour, uint32_t max_L,
                           uint32_t hub_threshold, int32_t rank, int32_t world, void* stream, pprb200_session** out) {
  std::lock_guard<std::mutex> lk(g_api_mutex);
  return session_create_impl(row_ptr, col, n, colour, max_L, hub_threshold, rank, world, stream, out);
}

void pprb200_session_destroy(pprb200_session* s) {
  std::lock_guard<std::mutex> lk(g_api_mutex);
  if (s) { cudaStreamSynchronize(s->stream); session_free(s); }
}

int pprb200_session_grank(pprb200_session* s, uint32_t K, uint32_t L, uint32_t iterations, double damping, double tolerance) {
  if (!s) return fail(PPRB200_ERR_PARAM, "session is NULL");
  std::lock_guard<std::mutex> lk(g_api_mutex);
  return session_grank_impl(s, K, L, iterations, damping, tolerance);
}

int pprb200_session_mc(pprb200_session* s, uint32_t K, uint32_t L, uint32_t R, double damping, uint64_t seed, uint32_t rounds) {
  if (!s) return fail(PPRB200_ERR_PARAM, "session is NULL");
  std::lock_guard<std::mutex> lk(g_api_mutex);
  return session_mc_impl(s, K, L, R, damping, seed, rounds);
}

int pprb200_session_fetch(pprb200_session* s, int32_t* out_ids, double* out_scores, uint32_t* out_cnt) {
  if (!s) return fail(PPRB200_ERR_PARAM, "session is NULL");
  std::lock_guard<std::mutex> lk(g_api_mutex);
  return session_fetch_impl(s, out_ids, out_scores, out_cnt);
}

int pprb200_session_stats(pprb200_session* s, pprb200_stats* stats) {
  if (!s) return fail(PPRB200_ERR_PARAM, "session is NULL");
  std::lock_guard<std::mutex> lk(g_api_mutex);
  return session_stats_impl(s, stats);
}

int pprb200_session_kernel_time(pprb200_session* s, int which, uint32_t* launches, double* total_ms) {
  if (!s) return fail(PPRB200_ERR_PARAM, "session is NULL");
  std::lock_guard<std::mutex> lk(g_api_mutex);
  if (s->last_mode < 0) return fail(PPRB200_ERR_STATE, "no run has been enqueued on this session");
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  if (which == 1) {
    float ms = 0;
    if (s->last_mode != MODE_MC || cudaEventElapsedTime(&ms, s->ev_walk[0], s->ev_walk[1]) != cudaSuccess) ms = 0;
    if (launches) *launches = s->last_mode == MODE_MC ? 1u : 0u;
    if (total_ms) *total_ms = ms;
    return PPRB200_OK;
  }
  if (which != 0) return fail(PPRB200_ERR_PARAM, "which=%d unknown", which);
  RunState h;
  CUDA_TRY(cudaMemcpy(&h, s->d_state, sizeof(h), cudaMemcpyDeviceToHost));
  double tot = 0;
  uint32_t cnt = 0;
  for (int i = 0; i < h.iter && (size_t)(2 * i + 1) < s->ev_merge.size(); i++) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, s->ev_merge[2 * i], s->ev_merge[2 * i + 1]) == cudaSuccess) { tot += ms; cnt++; }
  }
  if (launches) *launches = cnt;
  if (total_ms) *total_ms = tot;
  return PPRB200_OK;
}

// debug: copy the merge_par phase cycle counters (PPRB200_PROF=1) and clear them; out[2 * sm*3 * 8]
int pprb200_debug_prof(pprb200_session* s, unsigned long long* out, int* n_ctas) {
  if (!s || !s->d_prof) return fail(PPRB200_ERR_STATE, "profiling counters not enabled (PPRB200_PROF=1)");
  cudaStreamSynchronize(s->stream);
  const size_t cnt = (size_t)2 * s->sm_count * 8 * 8;
  cudaMemcpy(out, s->d_prof, cnt * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  cudaMemset(s->d_prof, 0, cnt * sizeof(unsigned long long));
  if (n_ctas) *n_ctas = s->sm_count * 8;
  return PPRB200_OK;
}

static int peer_push_mode() {
  const char* e = getenv("PPRB200_PUSH_MODE");
  return e ? atoi(e) : 0;
}
// SM clocks a cross-GPU barrier waits before it declares a peer dead (PPRB200_PEER_TIMEOUT_MS, default 30 s at ~2 GHz)
static long long peer_timeout_cycles() {
  double ms = 30000.0;
  if (const char* e = getenv("PPRB200_PEER_TIMEOUT_MS")) ms = std::max(1.0, atof(e));
  return (long long)(ms * 2.0e6);
}

// hand the stream-ordered allocator's cached blocks of the current device back to the driver (sessions keep freed memory in the
// pool for the next call: another process that wants the GPU's memory asks for this first)
int pprb200_release_cached_memory(void) {
  std::lock_guard<std::mutex> lk(g_api_mutex);
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return PPRB200_OK; }
  cudaDeviceSynchronize();
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
  cudaGetLastError();
  return PPRB200_OK;
}

int pprb200_find_partitions_device(const int64_t* row_ptr, const int32_t* col, int32_t n, uint8_t* colour) {
  int rc = validate_csr(row_ptr, col, n);
  if (rc) return rc;
  if (n > 0 && !colour) return fail(PPRB200_ERR_PARAM, "colour is NULL");
  if ((rc = device_ok())) return rc;
  if (n == 0) return PPRB200_OK;
  std::lock_guard<std::mutex> lk(g_api_mutex);
  RawCsrDev G;
  if (!G.upload(row_ptr, col, n, row_ptr[n])) return fail(PPRB200_ERR_CUDA, "upload of the graph failed");
  const ComponentFn on_device = [&](int32_t root, uint8_t* seen, uint8_t* col_out) { return device_component(G, root, seen, col_out); };
  return find_partitions(row_ptr, col, n, colour, &on_device);
}

// debug: out[i][4] = Philox4x32-10(counter = ctr_key[i][0..3], key = ctr_key[i][4..5]) computed by the device function of mc_walk.cuh
int pprb200_debug_philox(const uint32_t* ctr_key, uint32_t* out, int32_t nblocks) {
  if (!ctr_key || !out || nblocks <= 0) return fail(PPRB200_ERR_PARAM, "bad argument");
  int rc = device_ok();
  if (rc) return rc;
  uint32_t *d_in = nullptr, *d_out = nullptr;
  CUDA_TRY(cudaMalloc((void**)&d_in, (size_t)nblocks * 24));
  CUDA_TRY(cudaMalloc((void**)&d_out, (size_t)nblocks * 16));
  CUDA_TRY(cudaMemcpy(d_in, ctr_key, (size_t)nblocks * 24, cudaMemcpyHostToDevice));
  philox_probe_kernel<<<(nblocks + 127) / 128, 128>>>(d_in, d_out, nblocks);
  cudaError_t e = cudaMemcpy(out, d_out, (size_t)nblocks * 16, cudaMemcpyDeviceToHost);
  cudaFree(d_in);
  cudaFree(d_out);
  if (e != cudaSuccess) return fail(PPRB200_ERR_CUDA, "philox probe failed: %s", cudaGetErrorString(e));
  return PPRB200_OK;
}

// debug: merge_dense_kernel bookkeeping of the last run (RunState::dbg), out[8]
int pprb200_debug_counters(pprb200_session* s, unsigned long long* out) {
  if (!s || !out) return fail(PPRB200_ERR_PARAM, "NULL argument");
  cudaStreamSynchronize(s->stream);
  RunState h;
  CUDA_TRY(cudaMemcpy(&h, s->d_state, sizeof(h), cudaMemcpyDeviceToHost));
  for (int i = 0; i < 8; i++) out[i] = h.dbg[i];
  return PPRB200_OK;
}

// ---- multi-GPU wiring: CUDA IPC handles of the two basket buffers and the mailbox ----------------------------------
int pprb200_session_ipc_export(pprb200_session* s, void* out) {
  if (!s || !out) return fail(PPRB200_ERR_PARAM, "NULL argument");
  std::lock_guard<std::mutex> lk(g_api_mutex);
  cudaIpcMemHandle_t h[3];
  CUDA_TRY(cudaStreamSynchronize(s->stream));  // the mailbox is zeroed before any peer can post into it
  CUDA_TRY(cudaIpcGetMemHandle(&h[0], s->d_buf[0]));
  CUDA_TRY(cudaIpcGetMemHandle(&h[1], s->d_buf[1]));
  CUDA_TRY(cudaIpcGetMemHandle(&h[2], s->d_mbox));
  static_assert(sizeof(h) == PPRB200_IPC_BYTES, "PPRB200_IPC_BYTES");
  std::memcpy(out, h, sizeof(h));
  return PPRB200_OK;
}

int pprb200_session_ipc_attach(pprb200_session* s, const void* all_handles) {
  if (!s || !all_handles) return fail(PPRB200_ERR_PARAM, "NULL argument");
  std::lock_guard<std::mutex> lk(g_api_mutex);
  if (s->attached) return fail(PPRB200_ERR_STATE, "peers already attached");
  PeerDev pd;
  std::memset(&pd, 0, sizeof(pd));
  pd.world = s->world;
  pd.rank = s->rank;
  pd.timeout_cycles = peer_timeout_cycles();
  pd.push_mode = peer_push_mode();
  pd.need = s->d_need;
  for (int r = 0; r < s->world; r++) {
    if (r == s->rank) {
      pd.buf[r][0] = s->d_buf[0]; pd.buf[r][1] = s->d_buf[1]; pd.mbox[r] = s->d_mbox;
      continue;
    }
    cudaIpcMemHandle_t h[3];
    std::memcpy(h, (const unsigned char*)all_handles + (size_t)r * PPRB200_IPC_BYTES, sizeof(h));
    for (int i = 0; i < 3; i++) {
      void* ptr = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&ptr, h[i], cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) return fail(PPRB200_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d, object %d) failed: %s", r, i, cudaGetErrorString(e));
      s->ipc_opened[r][i] = ptr;
    }
    pd.buf[r][0] = (unsigned char*)s->ipc_opened[r][0];
    pd.buf[r][1] = (unsigned char*)s->ipc_opened[r][1];
    pd.mbox[r] = (Mailbox*)s->ipc_opened[r][2];
  }
  s->peers = pd;
  s->attached = true;
  return PPRB200_OK;
}

// same wiring for sessions that live in ONE process (one per device, or -- tests -- several on one device): plain peer
// pointers instead of IPC handles. all[r] = the session of rank r; every session gets the same view.
int pprb200_session_attach_local(pprb200_session** all, int32_t world) {
  if (!all || world < 1 || world > MAX_WORLD) return fail(PPRB200_ERR_PARAM, "bad argument");
  std::lock_guard<std::mutex> lk(g_api_mutex);
  for (int r = 0; r < world; r++)
    if (!all[r] || all[r]->world != world || all[r]->rank != r || all[r]->attached)
      return fail(PPRB200_ERR_STATE, "session %d is not an unattached rank %d of %d", r, r, world);
  int dev0 = 0;
  cudaGetDevice(&dev0);
  for (int a = 0; a < world; a++)
    for (int b2 = 0; b2 < world; b2++) {
      if (all[a]->device == all[b2]->device) continue;
      cudaSetDevice(all[a]->device);
      const cudaError_t e = cudaDeviceEnablePeerAccess(all[b2]->device, 0);
      cudaGetLastError();
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        cudaSetDevice(dev0);
        return fail(PPRB200_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", all[a]->device, all[b2]->device, cudaGetErrorString(e));
      }
    }
  for (int r = 0; r < world; r++) {
    PeerDev pd;
    std::memset(&pd, 0, sizeof(pd));
    pd.world = world;
    pd.rank = r;
    pd.timeout_cycles = peer_timeout_cycles();
    pd.push_mode = peer_push_mode();
    pd.need = all[r]->d_need;
    for (int q = 0; q < world; q++) { pd.buf[q][0] = all[q]->d_buf[0]; pd.buf[q][1] = all[q]->d_buf[1]; pd.mbox[q] = all[q]->d_mbox; }
    cudaSetDevice(all[r]->device);
    preload_kernels();                      // (see preload_kernels: no first-use code load behind a kernel that waits for a peer)
    cudaStreamSynchronize(all[r]->stream);  // mailbox zeroed before anybody posts
    all[r]->peers = pd;
    all[r]->attached = true;
  }
  cudaSetDevice(dev0);
  return PPRB200_OK;
}

// ---- quality evaluator yardstick: batched exact PPR (pprSingleSource.h:28-75) -----------------

int pprb200_ppr_exact(const int64_t* row_ptr, const int32_t* col, int32_t n, const int32_t* sources, uint32_t n_sources,
                      uint32_t iterations, double damping, double tolerance, double* out_scores, uint32_t* out_iterations,
                      double* kernel_ms) {
  // pprSingleSource.h:37-39, before the graph is touched
  if (iterations == 0) return fail(PPRB200_ERR_PARAM, "iterations must be positive");
  if (!(damping >= 0 && damping <= 1)) return fail(PPRB200_ERR_PARAM, "damping must be [0,1]");
  int rc = validate_csr(row_ptr, col, n);
  if (rc) return rc;
  if (n_sources == 0) return PPRB200_OK;
  if (!sources || !out_scores) return fail(PPRB200_ERR_PARAM, "NULL argument");
  for (uint32_t i = 0; i < n_sources; i++)
    if (sources[i] < 0 || sources[i] >= n) return fail(PPRB200_ERR_PARAM, "source node not part of the graph");
  if ((rc = device_ok())) return rc;
  std::lock_guard<std::mutex> lk(g_api_mutex);
  pool_setup_once();
  cudaStream_t st = nullptr;
  g_alloc_stream = st;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);

  std::vector<int64_t> prow;
  std::vector<int32_t> pcol;
  host_transpose(row_ptr, col, n, prow, pcol);
  std::vector<double> factor((size_t)n, 0.0);
  host_parallel_for(n, 1 << 15, [&](int, int64_t lo, int64_t hi) {
    for (int64_t u = lo; u < hi; u++) {
      const int64_t d = row_ptr[u + 1] - row_ptr[u];
      if (d > 0) factor[(size_t)u] = damping / (double)(uint64_t)d;  // pprSingleSource.h:57
    }
  });
  const int64_t E = row_ptr[n];
  std::vector<int> heavy;  // nodes whose predecessor list gets a CTA of its own
  int heavy_threshold = EXACT_HEAVY;
  if (const char* e = getenv("PPRB200_EXACT_HEAVY")) heavy_threshold = std::max(32, atoi(e));
  for (int32_t v = 0; v < n; v++)
    if (prow[(size_t)v + 1] - prow[(size_t)v] > heavy_threshold) heavy.push_back(v);
  int* d_heavy = nullptr;
  long long* d_prow = nullptr;
  int* d_pcol = nullptr;
  double* d_factor = nullptr;
  double* d_buf[2] = {nullptr, nullptr};
  int* d_src = nullptr;
  int* d_active = nullptr;
  int* d_parity = nullptr;
  long long* d_diff = nullptr;
  unsigned int* d_iters = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  auto cleanup = [&]() {
    dev_free(d_prow); dev_free(d_pcol); dev_free(d_factor); dev_free(d_buf[0]); dev_free(d_buf[1]); dev_free(d_src);
    dev_free(d_active); dev_free(d_parity); dev_free(d_diff); dev_free(d_iters); dev_free(d_heavy);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
  };
  // sources advance in batches of up to 256 (8 scores per lane), bounded so that the two score arrays fit in 16 GB
  const uint32_t by_mem = (uint32_t)std::max<uint64_t>(1, ((uint64_t)8 << 30) / ((uint64_t)std::max(n, 1) * 8));
  const uint32_t batch_max = std::min<uint32_t>(256u, std::max<uint32_t>(1u, by_mem));
  const uint32_t Bcap = std::min<uint32_t>(batch_max, n_sources);
  if ((rc = dev_alloc(&d_prow, (size_t)n + 1)) || (rc = dev_alloc(&d_pcol, (size_t)std::max<int64_t>(E, 1))) ||
      (rc = dev_alloc(&d_factor, (size_t)n)) || (rc = dev_alloc(&d_buf[0], (size_t)n * Bcap)) || (rc = dev_alloc(&d_buf[1], (size_t)n * Bcap)) ||
      (rc = dev_alloc(&d_src, Bcap)) || (rc = dev_alloc(&d_active, Bcap)) || (rc = dev_alloc(&d_parity, 1)) ||
      (rc = dev_alloc(&d_diff, Bcap)) || (rc = dev_alloc(&d_iters, Bcap)) || (rc = dev_alloc(&d_heavy, heavy.size()))) {
    cleanup();
    return rc;
  }
  if (!heavy.empty()) cudaMemcpyAsync(d_heavy, heavy.data(), heavy.size() * sizeof(int), cudaMemcpyHostToDevice, st);
  static_assert(sizeof(long long) == sizeof(int64_t), "row offsets");
  cudaMemcpyAsync(d_prow, prow.data(), ((size_t)n + 1) * sizeof(long long), cudaMemcpyHostToDevice, st);
  if (E) cudaMemcpyAsync(d_pcol, pcol.data(), (size_t)E * sizeof(int), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d_factor, factor.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st);
  cudaEventCreate(&ev0);
  cudaEventCreate(&ev1);
  double total_ms = 0;
  std::vector<double> stage;
  std::vector<int> ones;
  for (uint32_t b0 = 0; b0 < n_sources; b0 += Bcap) {
    const int B = (int)std::min<uint32_t>(Bcap, n_sources - b0);
    ones.assign((size_t)B, 1);
    cudaMemsetAsync(d_buf[0], 0, (size_t)n * B * sizeof(double), st);
    cudaMemsetAsync(d_parity, 0, sizeof(int), st);
    cudaMemsetAsync(d_diff, 0, (size_t)B * sizeof(long long), st);
    cudaMemsetAsync(d_iters, 0, (size_t)B * sizeof(unsigned int), st);
    cudaMemcpyAsync(d_src, sources + b0, (size_t)B * sizeof(int), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_active, ones.data(), (size_t)B * sizeof(int), cudaMemcpyHostToDevice, st);
    ExactParams P;
    P.prow = d_prow; P.pcol = d_pcol; P.factor = d_factor; P.buf[0] = d_buf[0]; P.buf[1] = d_buf[1]; P.parity = d_parity;
    P.source = d_src; P.active = d_active; P.diff = d_diff; P.iters = d_iters; P.n = n; P.B = B;
    P.heavy = d_heavy; P.n_heavy = (int)heavy.size(); P.heavy_threshold = heavy_threshold;
    P.teleport = 1.0 - damping; P.tolerance = tolerance;
    cudaEventRecord(ev0, st);
    ppr_exact_init_kernel<<<(B + 255) / 256, 256, 0, st>>>(d_buf[0], d_src, B);
    const int grid = std::max(1, std::min((n + 7) / 8, sms * 8));
    const int bt = (B + 31) / 32;
    for (uint32_t it = 0; it < iterations; it++) {
      if (bt <= 1) launch_exact_iter<1>(P, grid, sms, st);
      else if (bt <= 2) launch_exact_iter<2>(P, grid, sms, st);
      else if (bt <= 4) launch_exact_iter<4>(P, grid, sms, st);
      else launch_exact_iter<8>(P, grid, sms, st);
      ppr_exact_step_kernel<<<1, 256, 0, st>>>(P);
    }
    cudaEventRecord(ev1, st);
    int parity = 0;
    cudaMemcpyAsync(&parity, d_parity, sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { cleanup(); return fail(PPRB200_ERR_CUDA, "exact PPR failed: %s", cudaGetErrorString(e)); }
    float ms = 0;
    cudaEventElapsedTime(&ms, ev0, ev1);
    total_ms += ms;
    // [n][B] on the device -> out_scores[source][n]
    stage.resize((size_t)n * B);
    cudaMemcpy(stage.data(), d_buf[parity], (size_t)n * B * sizeof(double), cudaMemcpyDeviceToHost);
    host_parallel_for(B, 1, [&](int, int64_t lo, int64_t hi) {
      for (int64_t b = lo; b < hi; b++) {
        double* dst = out_scores + (size_t)(b0 + b) * (size_t)n;
        for (int32_t v = 0; v < n; v++) dst[v] = stage[(size_t)v * B + (size_t)b];
      }
    });
    if (out_iterations) cudaMemcpy(out_iterations + b0, d_iters, (size_t)B * sizeof(unsigned int), cudaMemcpyDeviceToHost);
  }
  if (kernel_ms) *kernel_ms = total_ms;
  cleanup();
  return PPRB200_OK;
}

// debug / CPU tests: the host front half of a session (colouring, storage order, rank labels, CSR encode, work items)
// without touching a device. Every output pointer may be NULL.
int pprb200_debug_host_plan(const int64_t* row_ptr, const int32_t* col, int32_t n, const uint8_t* colour, uint32_t hub_threshold,
                            int32_t rank, int32_t world, int32_t* pos_of, int32_t* rank_of, int64_t* row_off, uint32_t* enc,
                            int32_t* item_pos, int64_t* item_off, int32_t* item_len, int32_t item_cap, int32_t* summary) {
  HostPlanOut o;
  o.pos_of = pos_of; o.rank_of = rank_of; o.row_off = row_off; o.enc = enc; o.item_pos = item_pos; o.item_off = item_off;
  o.item_len = item_len; o.item_cap = item_cap; o.summary = summary;
  pprb200_session* unused = nullptr;
  std::lock_guard<std::mutex> lk(g_api_mutex);
  return session_create_impl(row_ptr, col, n, colour, 1, hub_threshold, rank, world, nullptr, &unused, true, &o);
}

// debug / CPU tests: owner and need mask of every node (the plan's per-position arrays mapped back to nodes); host only
int pprb200_debug_need_mask(const int64_t* row_ptr, const int32_t* col, int32_t n, const uint8_t* colour, uint32_t hub_threshold,
                            int32_t world, int32_t* owner, uint8_t* need) {
  if (world < 1 || world > MAX_WORLD) return fail(PPRB200_ERR_PARAM, "world must be 1..%d", MAX_WORLD);
  int rc = validate_csr(row_ptr, col, n);
  if (rc) return rc;
  if (n > 0 && (!owner || !need)) return fail(PPRB200_ERR_PARAM, "NULL argument");
  std::lock_guard<std::mutex> lk(g_api_mutex);
  HostPlan H;
  if ((rc = build_host_plan(row_ptr, col, n, colour, hub_threshold, world, /*need_colour=*/true, H, /*use_device=*/false))) return rc;
  for (int32_t v = 0; v < n; v++) {
    const int32_t p = H.pos_of[(size_t)v];
    owner[v] = p < 0 ? -1 : H.owner_of_pos[(size_t)p];
    need[v] = (p < 0 || H.need_mask.empty()) ? (uint8_t)0 : H.need_mask[(size_t)p];
  }
  return PPRB200_OK;
}

// which rank of `world` updates node v (-1: sinks, nobody), exactly as the sessions shard the work
int pprb200_shard_owner(const int64_t* row_ptr, const int32_t* col, int32_t n, const uint8_t* colour_in, uint32_t hub_threshold,
                        int32_t world, int32_t* owner) {
  if (world < 1 || world > MAX_WORLD) return fail(PPRB200_ERR_PARAM, "world must be 1..%d", MAX_WORLD);
  int rc = validate_csr(row_ptr, col, n);
  if (rc) return rc;
  if (!owner) return fail(PPRB200_ERR_PARAM, "owner is NULL");
  std::vector<uint8_t> colour((size_t)n, 0);
  if (colour_in) std::memcpy(colour.data(), colour_in, (size_t)n);
  std::vector<int32_t> order;
  int cls_begin[2][3], cls_end[2][3];
  std::vector<int32_t> owner_of_pos;
  storage_order(row_ptr, n, colour.data(), hub_threshold == 0 ? PPRB200_DEFAULT_HUB_THRESHOLD : hub_threshold, default_mid_deg(n), order,
                cls_begin, cls_end, world, &owner_of_pos);
  for (int32_t v = 0; v < n; v++) owner[v] = -1;
  for (size_t p = 0; p < order.size(); p++) owner[order[p]] = owner_of_pos[p];
  return PPRB200_OK;
}

int pprb200_session_launches(pprb200_session* s, uint64_t* launches) {
  if (!s || !launches) return fail(PPRB200_ERR_PARAM, "NULL argument");
  *launches = s->launch_count;
  return PPRB200_OK;
}

// ---- the one-shot entry points (what the template headers call) --------------------------------------------------
// GPUs of a one-shot call: PPR_NUM_GPUS from the environment, else every usable device as far as the graph
// gives each of them work (one GPU per 8 M edges: below that the per-iteration exchange and the 8 uploads cost more than they
// save), at most MAX_WORLD. Multi-GPU runs live in ONE process: a session per device sharing one host plan, peer access
// between the devices (cudaDeviceEnablePeerAccess; no IPC), the same kernels and mailbox barriers as the process-per-GPU path.
static int oneshot_world(int64_t n_edges) {
  const int avail = std::min(pprb200_device_count(), (int)MAX_WORLD);
  if (avail <= 1) return 1;
  if (const char* e = getenv("PPR_NUM_GPUS")) return std::max(1, std::min(avail, atoi(e)));
  return (int)std::max<int64_t>(1, std::min<int64_t>(avail, n_edges / (8ll << 20)));
}

struct OneShot {
  int mode;  // MODE_GRANK / MODE_MC
  uint32_t K, L, iterations;  // iterations = R for MC
  double damping, tolerance;
  uint64_t seed;
  uint32_t rounds;
};

static int run_oneshot(const int64_t* row_ptr, const int32_t* col, int32_t n, const uint8_t* colour, uint32_t hub_threshold,
                       const OneShot& job, int32_t* out_ids, double* out_scores, uint32_t* out_cnt, pprb200_stats* stats) {
  const double t0 = now_ms();
  int rc = validate_csr(row_ptr, col, n);
  if (rc) return rc;
  if ((rc = device_ok())) return rc;
  const int world = oneshot_world(row_ptr[n]);
  int dev0 = 0;
  cudaGetDevice(&dev0);
  std::vector<int> devs;  // sm_100 devices, the current one first
  {
    int cnt = 0;
    cudaGetDeviceCount(&cnt);
    devs.push_back(dev0);
    for (int d = 0; d < cnt && (int)devs.size() < world; d++) {
      int major = 0;
      if (d != dev0 && cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) devs.push_back(d);
    }
  }
  std::vector<pprb200_session*> ss((size_t)world, nullptr);
  std::vector<cudaStream_t> streams((size_t)world, nullptr);
  auto cleanup = [&]() {
    for (int r = 0; r < world; r++) {
      cudaSetDevice(devs[(size_t)r]);
      if (ss[(size_t)r]) { cudaStreamSynchronize(ss[(size_t)r]->stream); session_free(ss[(size_t)r]); }
      if (streams[(size_t)r]) cudaStreamDestroy(streams[(size_t)r]);
    }
    cudaSetDevice(dev0);
  };
  double t_plan = 0, t_up = 0, t_enq = 0, t_run = 0, t_d2h = 0;
  {
    HostPlan H;
    if ((rc = build_host_plan(row_ptr, col, n, colour, hub_threshold, world, job.mode == MODE_GRANK, H, /*use_device=*/true))) return rc;
    t_plan = now_ms();
    if (world > 1) {
      for (int a = 0; a < world && !rc; a++) {
        cudaSetDevice(devs[(size_t)a]);
        for (int b2 = 0; b2 < world; b2++) {
          if (a == b2) continue;
          int can = 0;
          cudaDeviceCanAccessPeer(&can, devs[(size_t)a], devs[(size_t)b2]);
          if (!can) { rc = fail(PPRB200_ERR_CUDA, "device %d cannot access device %d: set PPR_NUM_GPUS=1", devs[(size_t)a], devs[(size_t)b2]); break; }
          const cudaError_t e = cudaDeviceEnablePeerAccess(devs[(size_t)b2], 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { rc = fail(PPRB200_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", devs[(size_t)a], devs[(size_t)b2], cudaGetErrorString(e)); break; }
          cudaGetLastError();
        }
      }
      if (rc) { cudaSetDevice(dev0); return rc; }
    }
    for (int r = 0; r < world; r++) {
      cudaSetDevice(devs[(size_t)r]);
      if (world > 1) cudaStreamCreateWithFlags(&streams[(size_t)r], cudaStreamNonBlocking);
      RankPlan R;
      build_rank_plan(H, r, R);
      if ((rc = session_from_plan(H, R, job.L, r, streams[(size_t)r], /*plain cudaMalloc for peer-visible buffers*/ world > 1, &ss[(size_t)r],
                                  r > 0 ? ss[0] : nullptr))) { cleanup(); return rc; }
    }
    t_up = now_ms();
  }
  if (world > 1) {
    for (int r = 0; r < world; r++) {
      PeerDev pd;
      std::memset(&pd, 0, sizeof(pd));
      pd.world = world;
      pd.rank = r;
      pd.timeout_cycles = peer_timeout_cycles();
      pd.push_mode = peer_push_mode();
      pd.need = ss[(size_t)r]->d_need;
      for (int q = 0; q < world; q++) { pd.buf[q][0] = ss[(size_t)q]->d_buf[0]; pd.buf[q][1] = ss[(size_t)q]->d_buf[1]; pd.mbox[q] = ss[(size_t)q]->d_mbox; }
      ss[(size_t)r]->peers = pd;
      ss[(size_t)r]->attached = true;
    }
    for (int r = 0; r < world; r++) {
      cudaSetDevice(devs[(size_t)r]);
      preload_kernels();
      cudaStreamSynchronize(ss[(size_t)r]->stream);  // mailboxes zeroed everywhere before anybody posts
    }
  }
  for (int r = 0; r < world && !rc; r++) {
    cudaSetDevice(devs[(size_t)r]);
    rc = job.mode == MODE_GRANK ? session_grank_impl(ss[(size_t)r], job.K, job.L, job.iterations, job.damping, job.tolerance)
                                : session_mc_impl(ss[(size_t)r], job.K, job.L, job.iterations, job.damping, job.seed, job.rounds);
  }
  t_enq = now_ms();
  if (!rc) {
    for (int r = 0; r < world; r++) { cudaSetDevice(devs[(size_t)r]); cudaStreamSynchronize(ss[(size_t)r]->stream); }
    t_run = now_ms();
    cudaSetDevice(devs[0]);
    rc = session_fetch_impl(ss[0], out_ids, out_scores, out_cnt);  // every rank holds every basket: rank 0's final top-K is the result
    t_d2h = now_ms() - t_run;
  }
  if (!rc && stats) {
    pprb200_stats acc;
    std::memset(&acc, 0, sizeof(acc));
    for (int r = 0; r < world && !rc; r++) {
      cudaSetDevice(devs[(size_t)r]);
      pprb200_stats st;
      rc = session_stats_impl(ss[(size_t)r], &st, /*with_final=*/r == 0);
      if (rc) break;
      if (r == 0) acc = st;
      else {
        acc.nonsink_node_iterations += st.nonsink_node_iterations; acc.edge_reads += st.edge_reads; acc.merged_entries += st.merged_entries;
        acc.candidates += st.candidates; acc.truncations += st.truncations; acc.boundary_ties += st.boundary_ties;
        acc.algorithmic_bytes += st.algorithmic_bytes; acc.walk_steps += st.walk_steps; acc.walks += st.walks;
        acc.overflow_requeues += st.overflow_requeues; acc.walk_algorithmic_bytes += st.walk_algorithmic_bytes;
        acc.kernel_ms = std::max(acc.kernel_ms, st.kernel_ms);
        acc.h2d_ms += st.h2d_ms;
      }
    }
    acc.n_gpus = (uint32_t)world;
    acc.d2h_ms = t_d2h;
    *stats = acc;
  }
  const double t_f0 = now_ms();
  cleanup();
  if (getenv("PPRB200_HOST_TIMING"))
    fprintf(stderr, "[pprb200] one-shot call on %d GPU(s): plan %.2f ms, upload %.2f, enqueue %.2f, wait %.2f, fetch %.2f, free %.2f\n", world,
            t_plan - t0, t_up - t_plan, t_enq - t_up, t_run - t_enq, t_d2h, now_ms() - t_f0);
  if (!rc && stats) stats->total_ms = now_ms() - t0;
  return rc;
}

int pprb200_grank(const int64_t* row_ptr, const int32_t* col, int32_t n, const uint8_t* colour, uint32_t K, uint32_t L,
                  uint32_t iterations, double damping, double tolerance, uint32_t hub_threshold, int32_t* out_ids,
                  double* out_scores, uint32_t* out_cnt, pprb200_stats* stats) {
  int rc = check_params(K, L, iterations, damping);  // before touching the graph (test/grankTest.cc:22-28)
  if (rc) return rc;
  if (stats) std::memset(stats, 0, sizeof(*stats));
  if (n == 0) return PPRB200_OK;  // empty graph -> empty result (test/grankTest.cc:31-36)
  OneShot job;
  job.mode = MODE_GRANK; job.K = K; job.L = L; job.iterations = iterations; job.damping = damping; job.tolerance = tolerance;
  job.seed = 0; job.rounds = 0;
  std::lock_guard<std::mutex> lk(g_api_mutex);
  return run_oneshot(row_ptr, col, n, colour, hub_threshold, job, out_ids, out_scores, out_cnt, stats);
}

int pprb200_mccompletepathv2(const int64_t* row_ptr, const int32_t* col, int32_t n, uint32_t K, uint32_t L, uint32_t R,
                             double damping, uint64_t seed, uint32_t rounds, uint32_t hub_threshold, int32_t* out_ids,
                             double* out_scores, uint32_t* out_cnt, pprb200_stats* stats) {
  int rc = check_params(K, L, R, damping);  // before touching the graph (mccompletepathv2.h:190-194)
  if (rc) return rc;
  if (stats) std::memset(stats, 0, sizeof(*stats));
  if (n == 0) return PPRB200_OK;
  OneShot job;
  job.mode = MODE_MC; job.K = K; job.L = L; job.iterations = R; job.damping = damping; job.tolerance = 0.0;
  job.seed = seed; job.rounds = rounds;
  if (damping > MC_ORDER_FREE_MAX_DAMPING) hub_threshold = UINT32_MAX;  // (see MC_ORDER_FREE_MAX_DAMPING)
  std::lock_guard<std::mutex> lk(g_api_mutex);
  return run_oneshot(row_ptr, col, n, nullptr, hub_threshold, job, out_ids, out_scores, out_cnt, stats);
}

}  // extern "C"
