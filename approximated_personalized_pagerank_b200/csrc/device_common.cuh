// device_common.cuh -- data layout and warp-level building blocks shared by the GRank / MC kernels.
//
// HBM layout (DESIGN.md "Data layout"):
//   * Non-sink nodes live at "storage positions" p = 0..M-1 (colour-major, then class, then out-degree
//     descending). Sinks own no storage: their basket is the constant {v: 1-d} (GRank) / {v: 1} (MC).
//   * row_off[M+1] (int64) / col[E'] (uint32): CSR in storage order. A col word is
//        sink successor      : 0x80000000 | rank label of the sink
//        non-sink successor  : storage position | colour << 30
//   * Basket entries carry "rank labels": nodes numbered by in-degree descending (ties by dense id), so that the
//     most popular keys are the smallest integers (merge_par.cuh indexes a dense accumulator with them).
//     label[M] = rank label of the node at position p; dense_of[n] maps a label back to the dense (external) id,
//     which the canonical tie-break (score desc, dense id asc) and the final output use.
//   * baskets: two buffers of M slots; a slot is Lp = roundup4(L) entries stored as
//        int32 ids[Lp] | double scoreA[Lp/2] | double scoreB[Lp/2]
//     where entry e = 4g + r keeps its score in scoreA[2g + r] (r < 2) or scoreB[2g + r - 2] (r >= 2), so
//     lane g fetches four entries with three fully coalesced 16-byte loads. Unused entries have id -1.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pprb200 {

constexpr int KEY_EMPTY = -1;
constexpr uint32_t COL_SINK = 0x80000000u;
constexpr uint32_t COL_COLOUR_SHIFT = 30;
constexpr uint32_t COL_POS_MASK = 0x3fffffffu;
constexpr unsigned FULL = 0xffffffffu;

constexpr double NORM_SCALE = 0x1p61;   // norm1 fixed point (order-free sum, oracle/ppr_oracle.c)
constexpr double NORM_INV = 0x1p-61;
constexpr double GRANK_HUB_SCALE = 0x1p62, GRANK_HUB_INV = 0x1p-62;
constexpr double MC_HUB_SCALE = 0x1p59, MC_HUB_INV = 0x1p-59;

enum : int { MODE_GRANK = 0, MODE_MC = 1 };

// device-resident control block of one run (one per session)
struct RunState {
  int slot[2];               // buffer holding the current baskets of colour c
  int active;                // 1 while iterating; cleared by the convergence test (grank.h:92)
  int iter;                  // iterations executed so far
  long long m_prev, m_last;  // maxDiff of the two most recent iterations, 2^-61 fixed point
  long long cur_max;         // running max of the iteration in flight
  unsigned int work[12];     // work-fetch counters, one per kernel stage
  unsigned int qcount[8];    // overflow queue lengths (0-3: exact-order cascade, 4: merge_dense -> merge_par hand-over)
  unsigned long long node_iters, edge_reads, merged, cands, truncs, ties, abytes, requeues;
  unsigned long long walk_steps, walks, walk_bytes;
  unsigned int ws_next;      // bump allocator for the global-table workspace
  int peer_timeout;          // set when a cross-GPU barrier gave up waiting
  unsigned long long barrier_seq;  // executed cross-GPU barriers (kept across runs)
  // merge_dense_kernel bookkeeping (pprb200_debug_counters): nodes finished, nodes that ran pass 2, nodes with tau = 0,
  // hand-overs by reason (untrusted contribution, candidates > CMAX, tail table full), split-hub items passed on,
  // nodes handed over unread because their old basket was not full
  unsigned long long dbg[8];
};

// ------------------------------------------------------------------------------------------------
// Multi-GPU: one process per GPU; every GPU holds the whole CSR and both basket buffers and owns
// a share of each (colour, class) list chosen longest-processing-time-first (storage_order). A node's new basket is
// written locally and PUSHED into the same slot of every peer's buffer with plain stores through NVLink peer
// mappings (cudaIpc) by the warp / CTA that produced it -- the allgather of the reference design is fused into the
// merge / walk kernels' epilogue and overlaps the rest of the grid's work. Iterations are separated by a mailbox
// barrier (cross_gpu_barrier) that also carries the max-diff of the convergence test.
// ------------------------------------------------------------------------------------------------
constexpr int MAX_WORLD = 8;

struct Mailbox {  // one per (parity, sender rank), written by the sender, polled by the owner
  unsigned long long seq;
  long long value;
};

struct PeerDev {
  int world, rank;
  unsigned char* buf[MAX_WORLD][2];  // peer basket buffers (entry [rank] = the local ones)
  Mailbox* mbox[MAX_WORLD];          // peer mailboxes: mbox[r][parity * MAX_WORLD + sender]
  long long timeout_cycles;          // a barrier gives up after this many SM clocks (default ~30 s; PPRB200_PEER_TIMEOUT_MS)
  int push_mode;                     // A/B hook (PPRB200_PUSH_MODE): 0 = every 16 bytes to all peers in turn, 1 = peer by peer,
                                     // 2 = nothing is pushed (WRONG results: timing experiments only), 3 = to every peer (no need masks)
  const unsigned char* need;         // [M] bit r: rank r owns a predecessor of the node at this position, i.e. reads its basket
                                     // during the iterations (nullptr: everybody gets everything)
};

// copy the freshly written slot of position p to every peer; `lane`/`nlanes` = the calling warp or CTA.
// The caller must have made the local slot visible to all calling threads (__syncwarp / __syncthreads).
// Only peers that READ the basket get it (pd.need): on 8 GPUs a node of in-degree 4 has its predecessors on ~3 of them, and
// pushing to all 7 made the NVLink stores the bottleneck of an iteration (profiles/r2/sweeps.txt); whoever was left out
// receives the final baskets once, after the last iteration (final_push_kernel).
__device__ __forceinline__ void publish_slot(const PeerDev& pd, int which_buf, size_t slot_off, size_t bytes, int lane, int nlanes, int p) {
  if (pd.world <= 1) return;
  const int4* src = reinterpret_cast<const int4*>(pd.buf[pd.rank][which_buf] + slot_off);
  const int n16 = (int)(bytes >> 4);
  if (pd.push_mode == 2) return;
  unsigned int mask = (pd.need && pd.push_mode != 3 && p >= 0) ? (unsigned int)pd.need[p] : 0xffu;  // (p < 0: to everybody)
  mask &= ~(1u << pd.rank) & ((1u << pd.world) - 1u);
  if (mask == 0u) return;
  if (pd.push_mode == 1) {
    for (int r = 0; r < pd.world; r++) {
      if (!((mask >> r) & 1u)) continue;
      int4* dst = reinterpret_cast<int4*>(pd.buf[r][which_buf] + slot_off);
      for (int i = lane; i < n16; i += nlanes) __stcg(dst + i, __ldcg(src + i));
    }
    return;
  }
  for (int i = lane; i < n16; i += nlanes) {
    const int4 v = __ldcg(src + i);
    for (int r = 0; r < pd.world; r++)
      if ((mask >> r) & 1u) __stcg(reinterpret_cast<int4*>(pd.buf[r][which_buf] + slot_off) + i, v);
  }
}

// One thread of one CTA: tell every peer that this rank finished step `seq` (with `value`), wait for all of them,
// return the max of the values. Relies on stream order: the kernels that pushed this step's slots have completed,
// so their peer stores are performed before the flag store below is issued.
__device__ inline long long cross_gpu_barrier(const PeerDev& pd, unsigned long long seq, long long value, int* timed_out) {
  if (pd.world <= 1) return value;
  const int par = (int)(seq & 1ull);
  __threadfence_system();
  for (int r = 0; r < pd.world; r++) {
    Mailbox* m = pd.mbox[r] + par * MAX_WORLD + pd.rank;
    *reinterpret_cast<volatile long long*>(&m->value) = value;
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long*>(&m->seq) = seq;
  }
  __threadfence_system();
  long long best = value;
  const long long t0 = clock64();
  for (int r = 0; r < pd.world; r++) {
    const Mailbox* m = pd.mbox[pd.rank] + par * MAX_WORLD + r;
    while (*reinterpret_cast<const volatile unsigned long long*>(&m->seq) != seq) {
      if (clock64() - t0 > pd.timeout_cycles) { *timed_out = 1; return best; }  // a peer died: the host reports it (session_stats / fetch)
      __nanosleep(200);
    }
    __threadfence_system();
    const long long v = *reinterpret_cast<const volatile long long*>(&m->value);
    best = v > best ? v : best;
  }
  return best;
}

struct GraphDev {
  const long long* row_off;  // [M+1]
  const uint32_t* col;       // [E']
  const int* label;          // [M] rank label of the node stored at position p
  const int* dense_of;       // [n] rank label -> dense (external) id, for the canonical tie-break and the output
};

__host__ __device__ inline int roundup4(int x) { return (x + 3) & ~3; }
__host__ __device__ inline size_t slot_bytes(int Lp) { return (size_t)Lp * 12; }

__device__ __forceinline__ int score_index(int e, int Lp) {
  const int g = e >> 2, r = e & 3;
  return (r < 2) ? (2 * g + r) : ((Lp >> 1) + 2 * g + r - 2);
}

__device__ __forceinline__ uint32_t hash_key(int k) {
  uint32_t x = (uint32_t)k * 0x9E3779B1u;
  return x ^ (x >> 15);
}

__device__ __forceinline__ long long fix_norm(double x) { return __double2ll_rn(x * NORM_SCALE); }

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ unsigned long long warp_sum_ull(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ unsigned long long warp_max_ull(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { unsigned long long t = __shfl_xor_sync(FULL, v, o); v = t > v ? t : v; }
  return v;
}
__device__ __forceinline__ unsigned long long warp_min_ull(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { unsigned long long t = __shfl_xor_sync(FULL, v, o); v = t < v ? t : v; }
  return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------
// Warp-level radix select (keepTop, pprInternal.h:109-137, with the canonical tie-break).
// Finds the k-th largest of the 64-bit keys key(i), i in [0,n) with pred(i); 8-bit digits from the highest
// differing bit down; stops as soon as the boundary bucket is taken whole.
// Returns the threshold t: every in-play key > t is selected; if !tie all keys == (t's known prefix) bucket
// are selected too (selected <=> key >= t); if tie, `krem` of the keys == t must still be chosen.
// hist: 256 uint32 of shared memory private to the warp.
// ------------------------------------------------------------------------------------------------
// From the third pass on the keys still in play are probed first (min/max): a run of equal keys across the cut -- a
// quarter of all truncations on power-law graphs -- ends the search at once instead of walking the remaining digits,
// and keys that share more leading bits than one digit skip those positions. *ntied = keys equal to the threshold.
template <typename KeyFn, typename PredFn>
__device__ unsigned long long warp_radix_select(int n, int k, KeyFn key, PredFn pred, unsigned int* hist,
                                                bool* tie, int* krem, int* ntied = nullptr) {
  const int lane = lane_id();
  unsigned long long lo = ~0ull, hi = 0ull;
  for (int i = lane; i < n; i += 32)
    if (pred(i)) { const unsigned long long b = key(i); lo = b < lo ? b : lo; hi = b > hi ? b : hi; }
  lo = warp_min_ull(lo);
  hi = warp_max_ull(hi);
  *tie = false;
  *krem = 0;
  if (lo == hi) {  // every in-play key identical
    int cnt = 0;
    for (int i = lane; i < n; i += 32) cnt += pred(i) ? 1 : 0;
    cnt = warp_sum_int(cnt);
    if (cnt > k) { *tie = true; *krem = k; if (ntied) *ntied = cnt; }
    return lo;
  }
  const int top = 63 - __clzll((long long)(lo ^ hi));
  unsigned long long known = (top == 63) ? 0ull : ~((2ull << top) - 1ull);
  unsigned long long prefix = hi & known;
  int shift = top - 7 > 0 ? top - 7 : 0;
  int in_bucket = 0;
  for (int pass = 0;; pass++) {
    if (pass >= 2) {
      unsigned long long l2 = ~0ull, h2 = 0ull;
      for (int i = lane; i < n; i += 32)
        if (pred(i)) {
          const unsigned long long b = key(i);
          if ((b & known) == prefix) { l2 = b < l2 ? b : l2; h2 = b > h2 ? b : h2; }
        }
      l2 = warp_min_ull(l2);
      h2 = warp_max_ull(h2);
      if (l2 == h2) { *tie = true; *krem = k; if (ntied) *ntied = in_bucket; return l2; }
      const int t2 = 63 - __clzll((long long)(l2 ^ h2));
      known = ~((2ull << t2) - 1ull);
      prefix = h2 & known;
      shift = t2 - 7 > 0 ? t2 - 7 : 0;
    }
    for (int i = lane; i < 256; i += 32) hist[i] = 0;
    __syncwarp();
    for (int i = lane; i < n; i += 32)
      if (pred(i)) {
        const unsigned long long b = key(i);
        if ((b & known) == prefix) atomicAdd(&hist[(unsigned)(b >> shift) & 0xffu], 1u);
      }
    __syncwarp();
    // lane owns bins [8*lane, 8*lane+8); walk from the top bin down
    unsigned int mine[8];
    unsigned int local = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) { mine[j] = hist[lane * 8 + j]; local += mine[j]; }
    // above = in-play keys in bins of higher lanes
    unsigned int incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int t = __shfl_down_sync(FULL, incl, o);
      if (lane + o < 32) incl += t;
    }
    const unsigned int above = incl - local;
    const bool here = above < (unsigned)k && incl >= (unsigned)k;
    const unsigned ball = __ballot_sync(FULL, here);
    const int owner = __ffs(ball) - 1;  // exactly one lane
    int digit = 0;
    unsigned int cnt_above = 0, cnt_d = 0;
    if (lane == owner) {
      unsigned int run = above;
#pragma unroll
      for (int j = 7; j >= 0; j--) {
        if (run < (unsigned)k && run + mine[j] >= (unsigned)k) { digit = lane * 8 + j; cnt_above = run; cnt_d = mine[j]; }
        run += mine[j];
      }
    }
    digit = __shfl_sync(FULL, digit, owner);
    cnt_above = __shfl_sync(FULL, cnt_above, owner);
    cnt_d = __shfl_sync(FULL, cnt_d, owner);
    k -= (int)cnt_above;
    prefix |= (unsigned long long)digit << shift;
    known |= 0xffull << shift;
    __syncwarp();
    if ((int)cnt_d == k) return prefix;  // whole bucket selected; unknown low bits of the threshold are 0
    in_bucket = (int)cnt_d;
    if (shift == 0) { *tie = true; *krem = k; if (ntied) *ntied = in_bucket; return prefix; }
    shift = shift - 8 > 0 ? shift - 8 : 0;
  }
}

struct Threshold {
  unsigned long long bits;  // selected <=> score bits > bits, or == bits and id <= id_max
  int id_max;
};

}  // namespace pprb200
