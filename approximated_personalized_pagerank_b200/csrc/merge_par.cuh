// merge_par.cuh -- the order-free merge step for nodes above the hub threshold: one CTA per work chunk.
//
// Same operator as merge_seq.cuh (grank.h:96-126, mccompletepathv2.h:211-250) but every contribution
// x * d/outdeg is rounded once to 2^-62 (MC: 2^-59) fixed point and summed as an integer, so the result does
// not depend on the order of the additions: all warps of a CTA -- and, for nodes split into several chunks,
// several CTAs -- accumulate concurrently with atomics and still produce bit-identical sums
// (oracle/ppr_oracle.c applies the same rounding). |difference to the reference's fma chain| <= outdeg * 2^-62.
//
// Basket keys are "rank labels": nodes numbered by in-degree descending. On power-law graphs 93-98 % of all
// merged entries carry one of the ~12 K most popular keys (profiles/README.md), so the accumulator is
//   1. DENSE: a direct-mapped shared-memory array of H fixed-point words for labels < H -- no hashing, no key
//      compare, no insertion: one ATOMS.ADD.32 on the low word, the carry into the high word;
//   2. TAIL: a shared-memory open-addressing table (TCAP slots) for the other labels, first come first served;
//   3. GLOBAL: when the tail outgrows its table, or the node is split into several chunks, a table from a small
//      L2-resident pool; chunk CTAs flush their shared accumulators into it and the CTA that completes the last
//      chunk of the node selects the top-L.
#pragma once
#include "merge_seq.cuh"

namespace pprb200 {

struct GSlot {
  int key;
  int pad;
  unsigned long long acc;
};

struct ParParams {
  MergeParams M;               // graph, buffers, mode, colour, init_mode, do_norm ...
  const int* item_pos;         // work items: node position,
  const long long* item_begin; //             first successor (absolute offset into col),
  const int* item_len;         //             number of successors
  int n_items;
  int chunk;                   // successors per chunk (nchunks = ceil(deg / chunk))
  int work_idx;
  // global table pool
  unsigned char* pool;
  size_t tbl_bytes;
  unsigned int capmax;         // slots per pool table (power of two)
  int n_tables;
  unsigned int* tbl_inuse;     // [n_tables]
  unsigned int* tbl_count;     // [n_tables] distinct keys inserted
  unsigned int* node_tbl;      // [M] 0 none, 1 being acquired, else table index + 2
  unsigned int* node_done;     // [M] chunks completed
  int n_ids;                   // label space (identity hashing when the table covers it)
  int use_sketch;              // two-pass merge with the tail sketch for single-item nodes
  const unsigned int* item_queue;  // nullable: item indices handed over by merge_dense_kernel (its overflow / split hubs),
  int queue_idx;                   //   st->qcount[queue_idx] of them; null: items 0 .. n_items-1
  unsigned long long* prof;    // optional [gridDim.x * 8] phase cycle counters (PPRB200_PROF=1)
};

struct ParShared {
  int tcount;           // occupied slots of the tail table
  int spilled;          // an entry found the tail table closed and no global table was bound
  int table;            // global table index of the current node (-1 none)
  unsigned int gmask;
  int gidentity;
  int is_last;
  int out_pos;
  int ncand;
  unsigned int item;
  unsigned long long red_a[64];
  int sel_digit, sel_above, sel_cnt;
  unsigned int hist[256];
};

// CTA-wide reductions: warp shuffles, one shared-memory word per warp, then every warp folds the (<= 32) partial
// results with shuffles again -- no serial loop over the warps. All threads must call; `scratch` holds >= 32 words.
__device__ __forceinline__ long long block_reduce_sum_ll(long long v, unsigned long long* scratch) {
  v = warp_sum_ll(v);
  const int nw = (int)(blockDim.x >> 5);
  if (lane_id() == 0) scratch[threadIdx.x >> 5] = (unsigned long long)v;
  __syncthreads();
  long long r = warp_sum_ll(lane_id() < nw ? (long long)scratch[lane_id()] : 0ll);
  __syncthreads();
  return r;
}
// two reductions behind one pair of barriers: min of `a`, sum of `b`
__device__ __forceinline__ void block_reduce_min_sum(unsigned long long& a, long long& b, unsigned long long* scratch) {
  a = warp_min_ull(a);
  b = warp_sum_ll(b);
  const int nw = (int)(blockDim.x >> 5);
  if (lane_id() == 0) { scratch[threadIdx.x >> 5] = a; scratch[32 + (threadIdx.x >> 5)] = (unsigned long long)b; }
  __syncthreads();
  a = warp_min_ull(lane_id() < nw ? scratch[lane_id()] : ~0ull);
  b = warp_sum_ll(lane_id() < nw ? (long long)scratch[32 + lane_id()] : 0ll);
  __syncthreads();
}
__device__ __forceinline__ void block_reduce_min_max(unsigned long long& a, unsigned long long& b, unsigned long long* scratch) {
  a = warp_min_ull(a);
  b = warp_max_ull(b);
  const int nw = (int)(blockDim.x >> 5);
  if (lane_id() == 0) { scratch[threadIdx.x >> 5] = a; scratch[32 + (threadIdx.x >> 5)] = b; }
  __syncthreads();
  a = warp_min_ull(lane_id() < nw ? scratch[lane_id()] : ~0ull);
  b = warp_max_ull(lane_id() < nw ? scratch[32 + lane_id()] : 0ull);
  __syncthreads();
}
__device__ __forceinline__ void block_reduce_sum2(long long& a, long long& b, unsigned long long* scratch) {
  a = warp_sum_ll(a);
  b = warp_sum_ll(b);
  const int nw = (int)(blockDim.x >> 5);
  if (lane_id() == 0) { scratch[threadIdx.x >> 5] = (unsigned long long)a; scratch[32 + (threadIdx.x >> 5)] = (unsigned long long)b; }
  __syncthreads();
  a = warp_sum_ll(lane_id() < nw ? (long long)scratch[lane_id()] : 0ll);
  b = warp_sum_ll(lane_id() < nw ? (long long)scratch[32 + lane_id()] : 0ll);
  __syncthreads();
}

// CTA-wide version of warp_radix_select (device_common.cuh); same contract (incl. the min/max probe from pass 3 on).
template <typename KeyFn, typename PredFn>
__device__ unsigned long long block_radix_select(int n, int k, KeyFn key, PredFn pred, ParShared* S, bool* tie, int* krem,
                                                 int* ntied = nullptr) {
  const int tid = threadIdx.x, T = blockDim.x;
  unsigned long long lo = ~0ull, hi = 0ull;
  long long cnt = 0;
  for (int i0 = tid; i0 < n; i0 += 4 * T) {  // four independent loads in flight per thread (the keys may live in L2)
    unsigned long long b[4];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; u++) { const int i = i0 + u * T; ok[u] = i < n && pred(i); b[u] = ok[u] ? key(i) : 0ull; }
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (ok[u]) { lo = b[u] < lo ? b[u] : lo; hi = b[u] > hi ? b[u] : hi; cnt++; }
  }
  block_reduce_min_max(lo, hi, S->red_a);
  *tie = false;
  *krem = 0;
  if (lo == hi) {
    cnt = block_reduce_sum_ll(cnt, S->red_a);
    if (cnt > k) { *tie = true; *krem = k; if (ntied) *ntied = (int)cnt; }
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
      for (int i0 = tid; i0 < n; i0 += 4 * T) {
        unsigned long long b[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { const int i = i0 + u * T; ok[u] = i < n && pred(i); b[u] = ok[u] ? key(i) : 0ull; }
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (ok[u] && (b[u] & known) == prefix) { l2 = b[u] < l2 ? b[u] : l2; h2 = b[u] > h2 ? b[u] : h2; }
      }
      block_reduce_min_max(l2, h2, S->red_a);
      if (l2 == h2) { *tie = true; *krem = k; if (ntied) *ntied = in_bucket; return l2; }
      const int t2 = 63 - __clzll((long long)(l2 ^ h2));
      known = ~((2ull << t2) - 1ull);
      prefix = h2 & known;
      shift = t2 - 7 > 0 ? t2 - 7 : 0;
    }
    for (int i = tid; i < 256; i += T) S->hist[i] = 0;
    __syncthreads();
    for (int i0 = tid; i0 < n; i0 += 4 * T) {
      unsigned long long b[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; u++) { const int i = i0 + u * T; ok[u] = i < n && pred(i); b[u] = ok[u] ? key(i) : 0ull; }
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (ok[u] && (b[u] & known) == prefix) atomicAdd(&S->hist[(unsigned)(b[u] >> shift) & 0xffu], 1u);
    }
    __syncthreads();
    if (tid < 32) {
      const int lane = tid;
      unsigned int mine[8];
      unsigned int local = 0;
#pragma unroll
      for (int j = 0; j < 8; j++) { mine[j] = S->hist[lane * 8 + j]; local += mine[j]; }
      unsigned int incl = local;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_down_sync(FULL, incl, o);
        if (lane + o < 32) incl += t;
      }
      const unsigned int above = incl - local;
      if (above < (unsigned)k && incl >= (unsigned)k) {
        unsigned int run = above;
#pragma unroll
        for (int j = 7; j >= 0; j--) {
          if (run < (unsigned)k && run + mine[j] >= (unsigned)k) { S->sel_digit = lane * 8 + j; S->sel_above = (int)run; S->sel_cnt = (int)mine[j]; }
          run += mine[j];
        }
      }
    }
    __syncthreads();
    const int digit = S->sel_digit, cnt_above = S->sel_above, cnt_d = S->sel_cnt;
    k -= cnt_above;
    prefix |= (unsigned long long)digit << shift;
    known |= 0xffull << shift;
    __syncthreads();
    if (cnt_d == k) return prefix;
    in_bucket = cnt_d;
    if (shift == 0) { *tie = true; *krem = k; if (ntied) *ntied = in_bucket; return prefix; }
    shift = shift - 8 > 0 ? shift - 8 : 0;
  }
}

// Cut through a run of `ntied` (<= 32) candidates whose key equals `tb`: the `krem` smallest dense ids stay. The tied
// elements are gathered into shared memory and ranked by warp 0 with shuffles -- one scan instead of a second radix
// select with a dense_of lookup per pass. Returns id_max. All threads of the CTA must call it.
template <typename KeyFn, typename PredFn, typename DenseFn>
__device__ int block_small_tie_cut(int n, int krem, unsigned long long tb, KeyFn key, PredFn pred, DenseFn dense, ParShared* S) {
  const int tid = threadIdx.x, T = blockDim.x;
  __syncthreads();
  if (tid == 0) S->sel_cnt = 0;
  __syncthreads();
  for (int i = tid; i < n; i += T)
    if (pred(i) && key(i) == tb) { const int pos = atomicAdd(&S->sel_cnt, 1); if (pos < 32) S->hist[pos] = (unsigned int)dense(i); }
  __syncthreads();
  if (tid < 32) {
    const int m = S->sel_cnt < 32 ? S->sel_cnt : 32;
    const int myd = tid < m ? (int)S->hist[tid] : 0x7fffffff;
    int rank = 0;
    for (int j = 0; j < m; j++) rank += __shfl_sync(FULL, myd, j) < myd;
    const unsigned who = __ballot_sync(FULL, tid < m && rank == krem - 1);
    const int cut = __shfl_sync(FULL, myd, __ffs(who) - 1);
    if (tid == 0) S->sel_digit = cut;
  }
  __syncthreads();
  const int r = S->sel_digit;
  __syncthreads();
  return r;
}

// fixed-point word -> score. init mode: the word is a multiplicity m and the score is `base + mult + ... + mult`
// (m additions, grank.h:79-80); otherwise score = word * 2^-62 (GRank) / 2^-59 (MC).
__device__ __forceinline__ double par_score(unsigned long long acc, bool init_mode, double inv, double mult, double base) {
  if (!init_mode) return (double)(long long)acc * inv;
  double a = base;
  for (unsigned long long r = 0; r < acc; r++) a += mult;
  return a;
}

// accumulate x on key k of a global table; returns true when this call inserted the key (slot in *slot): the caller
// appends the slot to the table's first-touch list (glist_append: one counter atomic per warp, not per key)
__device__ __forceinline__ bool gtable_upsert(GSlot* slots, unsigned int mask, int identity, int k, unsigned long long x,
                                              unsigned int* slot) {
  unsigned int h = (identity ? (unsigned int)k : hash_key(k)) & mask;
  bool inserted = false;
  for (;;) {
    const int cur = *reinterpret_cast<volatile int*>(&slots[h].key);
    if (cur == k) break;
    if (cur == KEY_EMPTY) {
      const int old = atomicCAS(&slots[h].key, KEY_EMPTY, k);
      if (old == KEY_EMPTY) { inserted = true; break; }
      if (old == k) break;
    }
    h = (h + 1) & mask;
  }
  if (x) atomicAdd(&slots[h].acc, x);
  *slot = h;
  return inserted;
}

// all 32 lanes call
__device__ __forceinline__ void glist_append(bool inserted, unsigned int slot, unsigned int* glist, unsigned int* gcount) {
  const unsigned m = __ballot_sync(FULL, inserted);
  if (!m) return;
  const int lane = lane_id(), leader = __ffs(m) - 1;
  unsigned int base = 0;
  if (lane == leader) base = atomicAdd(gcount, (unsigned int)__popc(m));
  base = __shfl_sync(FULL, base, leader);
  if (inserted) glist[base + __popc(m & ((1u << lane) - 1u))] = slot;
}

// any thread on its own (divergent callers)
__device__ __forceinline__ void gtable_add(GSlot* slots, unsigned int* glist, unsigned int* gcount, unsigned int mask,
                                           int identity, int k, unsigned long long x) {
  unsigned int h;
  if (gtable_upsert(slots, mask, identity, k, x, &h)) glist[atomicAdd(gcount, 1u)] = h;
}

__device__ __forceinline__ int gtable_find(const GSlot* slots, unsigned int mask, int identity, int k) {
  unsigned int h = (identity ? (unsigned int)k : hash_key(k)) & mask;
  for (;;) {
    const int cur = slots[h].key;
    if (cur == k) return (int)h;
    if (cur == KEY_EMPTY) return -1;
    h = (h + 1) & mask;
  }
}

__device__ __forceinline__ void fixed_add_shared(uint2* word, unsigned long long x) {
  const unsigned int xlo = (unsigned int)x, xhi = (unsigned int)(x >> 32);
  const unsigned int old = atomicAdd(&word->x, xlo);
  const unsigned int hi = xhi + ((old + xlo) < old ? 1u : 0u);
  if (hi) atomicAdd(&word->y, hi);
}

constexpr int PAR_CHUNK_MAX = 1024;  // column words of the big class are staged in shared memory in tiles of this many successors
constexpr int PAR_PRE = 128;         // old-basket tail labels that get an exact accumulator of their own during pass 1
constexpr int PAR_MID_MAX = 128;     // largest out-degree the mid class may be configured for

template <int H, int TCAP, int CMAX, int COLCAP, int R>
constexpr size_t par_smem_bytes() {
  return (size_t)H * 8 + (size_t)TCAP * 14 + (size_t)H / 8 + (size_t)CMAX * 12 + (size_t)COLCAP * 4 +
         (R > 0 ? (size_t)R * 8 + (size_t)R / 8 + (size_t)R + (size_t)PAR_PRE * 16 : 0) + sizeof(ParShared);
}
constexpr size_t par_queue_bytes(int threads) { return (size_t)(threads / 32) * 64 * 12; }

// unconditional fetch of lane g's four entries (three independent 16-byte loads; unused entries hold id -1 and
// whatever score bytes were there -- never looked at)
__device__ __forceinline__ void load_frag_all(const unsigned char* slot, int Lp, int g, BasketFrag* f) {
  const int4* ids = reinterpret_cast<const int4*>(slot);
  const double2* sc = reinterpret_cast<const double2*>(slot + (size_t)Lp * 4);
  f->id = __ldg(ids + g);
  f->sa = __ldg(sc + g);
  f->sb = __ldg(sc + (Lp >> 2) + g);
}

// H dense labels, TCAP tail slots (TLIMIT distinct tail keys admitted; concurrent inserts may overshoot by < THREADS),
// CMAX compact candidates, COLCAP staged column words per tile, R sketch buckets (0: single-pass only), THREADS threads.
//
// Two-pass merge of a single-item node (R > 0): most of a hub's candidates are tail labels that receive one or two tiny
// contributions and can never reach the top-L, yet hashing them costs far more than the dense labels that decide the
// result. Pass 1 adds dense labels exactly and tail labels into a count-min style sketch (bucket = hash(label), same
// fixed point): a bucket sum bounds every label in it from above. The L-th largest dense score tau bounds the final
// threshold from below (adding candidates can only raise it), so a tail label whose bucket sums to less than tau is out
// -- strictly, ties included. Pass 2 re-reads the baskets and accumulates exactly only the tail labels of surviving
// buckets. Everything stays in shared memory; the result is the same set of sums the single pass produces for every
// candidate that can be selected.
template <int H, int TCAP, int CMAX, int COLCAP, int R, int THREADS>
__global__ void __launch_bounds__(THREADS, (THREADS <= 128 ? 3 : 1)) merge_par_kernel(ParParams P) {
  constexpr int TLIMIT = TCAP * 13 / 16 - THREADS;
  constexpr int NW = THREADS / 32;
  static_assert(H + TCAP <= 65536, "candidate references are 16 bit");
  extern __shared__ __align__(16) unsigned char smem[];
  const MergeParams& M = P.M;
  RunState* st = M.st;
  if (!st->active) return;
  unsigned char* sp = smem;
  uint2* s_dense = reinterpret_cast<uint2*>(sp); sp += (size_t)H * 8;
  uint2* t_acc = reinterpret_cast<uint2*>(sp); sp += (size_t)TCAP * 8;
  static_assert(((size_t)COLCAP * 4) % 8 == 0 && ((size_t)R / 8) % 4 == 0, "shared-memory carve-up alignment");
  unsigned long long* c_bits = reinterpret_cast<unsigned long long*>(sp); sp += (size_t)CMAX * 8;  // compact candidates: score bits
  int* t_keys = reinterpret_cast<int*>(sp); sp += (size_t)TCAP * 4;
  int* c_id = reinterpret_cast<int*>(sp); sp += (size_t)CMAX * 4;                                   //                     labels
  uint32_t* s_col = reinterpret_cast<uint32_t*>(sp); sp += (size_t)COLCAP * 4;
  uint2* s_sk = reinterpret_cast<uint2*>(sp); sp += (size_t)R * 8;                     // sketch buckets (fixed point)
  unsigned int* s_alive = reinterpret_cast<unsigned int*>(sp); sp += (size_t)R / 8;    // buckets that may hold a top-L label
  uint2* s_pacc = reinterpret_cast<uint2*>(sp); sp += (R > 0 ? (size_t)PAR_PRE * 8 : 0);   // exact sums of old-basket tail labels
  int* s_pkey = reinterpret_cast<int*>(sp); sp += (R > 0 ? (size_t)PAR_PRE * 4 : 0);       //   their labels (-1: slot unused)
  unsigned char* s_ptouch = sp; sp += (R > 0 ? (size_t)PAR_PRE * 4 : 0);                   //   touched flags (padded)
  unsigned char* s_pslot = sp; sp += (size_t)R;                                            // bucket -> slot + 1 (0: none)
  unsigned int* s_zbits = reinterpret_cast<unsigned int*>(sp); sp += (size_t)H / 8;  // dense labels touched with a 0 word
  unsigned short* t_list = reinterpret_cast<unsigned short*>(sp); sp += (size_t)TCAP * 2;
  ParShared* S = reinterpret_cast<ParShared*>(sp); sp += (sizeof(ParShared) + 7) & ~(size_t)7;
  unsigned long long* q_val = reinterpret_cast<unsigned long long*>(sp); sp += (size_t)NW * 64 * 8;  // per-warp slow-entry queues
  int* q_key = reinterpret_cast<int*>(sp);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int Lp = M.Lp, groups = Lp >> 2, L = M.L;
  const bool init_mode = M.init_mode != 0;
  const double scale = (M.mode == MODE_GRANK) ? GRANK_HUB_SCALE : MC_HUB_SCALE;
  const double inv = (M.mode == MODE_GRANK) ? GRANK_HUB_INV : MC_HUB_INV;
  const int* __restrict__ dense_of = M.g.dense_of;

  for (int i = tid; i < H; i += THREADS) s_dense[i] = make_uint2(0u, 0u);
  for (int i = tid; i < H / 32; i += THREADS) s_zbits[i] = 0u;
  for (int i = tid; i < TCAP; i += THREADS) { t_keys[i] = KEY_EMPTY; t_acc[i] = make_uint2(0u, 0u); }
  for (int i = tid; i < R; i += THREADS) s_pslot[i] = 0;
  if (tid == 0) { S->tcount = 0; S->spilled = 0; S->table = -1; }
  __syncthreads();

  const int read_slot0 = st->slot[0], read_slot1 = st->slot[1];  // (two scalars: a dynamically indexed array would live in local memory)

  unsigned long long s_merged = 0, s_edges = 0, s_cands = 0, s_truncs = 0, s_ties = 0, s_bytes = 0, s_nodes = 0, s_requeue = 0;

  // tail-table accumulate; false when the key is absent and the table is closed (caller spills)
  auto tail_add = [&](int k, unsigned long long x) -> bool {
    unsigned int h = hash_key(k) & (TCAP - 1);
    volatile int* keys = t_keys;
    for (;;) {
      const int cur = keys[h];
      if (cur == k) break;
      if (cur == KEY_EMPTY) {
        if (*reinterpret_cast<volatile int*>(&S->tcount) >= TLIMIT) return false;
        const int old = atomicCAS(&t_keys[h], KEY_EMPTY, k);
        if (old == KEY_EMPTY) { const int pos = atomicAdd(&S->tcount, 1); t_list[pos] = (unsigned short)h; break; }
        if (old == k) break;
      }
      h = (h + 1) & (TCAP - 1);
    }
    if (x) fixed_add_shared(&t_acc[h], x);
    return true;
  };
  auto dense_touched = [&](int i, const uint2& a) -> bool { return ((a.x | a.y) != 0u) || ((s_zbits[i >> 5] >> (i & 31)) & 1u); };

  unsigned long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long t_last = clock64();
#define PROF_MARK(i) do { if (P.prof && tid == 0) { const long long t_now = clock64(); pc[i] += (unsigned long long)(t_now - t_last); t_last = t_now; } } while (0)
  const unsigned int n_work = P.item_queue ? st->qcount[P.queue_idx] : (unsigned)P.n_items;
  for (;;) {
    if (tid == 0) {
      const unsigned int idx = atomicAdd(&st->work[P.work_idx], 1u);
      S->item = idx >= n_work ? 0xffffffffu : (P.item_queue ? P.item_queue[idx] : idx);
    }
    __syncthreads();
    const unsigned int item = S->item;
    if (item == 0xffffffffu) break;
    PROF_MARK(0);
    const int p = P.item_pos[item];
    const long long cb = P.item_begin[item];
    const int clen = P.item_len[item];
    const long long rb = M.g.row_off[p], re = M.g.row_off[p + 1];
    const long long deg = re - rb;
    const int nchunks = (int)((deg + P.chunk - 1) / P.chunk);
    const int self_id = M.g.label[p];
    const double f = M.damping / (double)(unsigned long long)deg;
    const double mult = f;  // the hub path pre-scales by f = d/outdeg in both modes (oracle: llrint((x * f) * scale))
    // (x * f) * 2^s == x * (f * 2^s) exactly: scaling by a power of two commutes with the rounding of the product
    const double fscale = f * scale;
    const double self0 = (M.mode == MODE_GRANK) ? M.self_grank : 1.0;
    const int write_slot = init_mode ? st->slot[M.colour] : (st->slot[M.colour] ^ 1);

    // lazily bound global table of this node (acquired by the first CTA that needs it)
    GSlot* gslots = nullptr;
    unsigned int* glist = nullptr;
    unsigned int* gcount = nullptr;
    unsigned long long* gbits = nullptr;
    int* cids = nullptr;
    auto bind_table = [&]() {
      if (tid == 0 && S->table < 0) {
        unsigned int v = atomicCAS(&P.node_tbl[p], 0u, 1u);
        if (v == 0u) {
          int i = (int)((blockIdx.x * 2u) % (unsigned)P.n_tables);
          while (atomicCAS(&P.tbl_inuse[i], 0u, 1u) != 0u) i = (i + 1) % P.n_tables;
          __threadfence();
          atomicExch(&P.node_tbl[p], (unsigned)i + 2u);
          v = (unsigned)i + 2u;
        } else {
          while (v == 1u) v = atomicAdd(&P.node_tbl[p], 0u);
        }
        S->table = (int)v - 2;
        // per-node capacity: never overflows (worst case deg*Lp+1 distinct keys, or the whole label space)
        const unsigned long long bound = (unsigned long long)deg * (unsigned long long)(init_mode ? 1 : Lp) + 2ull;
        const unsigned long long want = 2ull * bound;
        unsigned int cap = P.capmax;
        if (want < (unsigned long long)P.capmax) { cap = 1024u; while ((unsigned long long)cap < want) cap <<= 1; }
        S->gmask = cap - 1u;
        S->gidentity = cap > (unsigned)P.n_ids ? 1 : 0;
      }
      __syncthreads();
      unsigned char* base = P.pool + (size_t)S->table * P.tbl_bytes;
      gslots = reinterpret_cast<GSlot*>(base);
      glist = reinterpret_cast<unsigned int*>(base + (size_t)P.capmax * sizeof(GSlot));
      gbits = reinterpret_cast<unsigned long long*>(base + (size_t)P.capmax * (sizeof(GSlot) + 4));
      cids = reinterpret_cast<int*>(base + (size_t)P.capmax * (sizeof(GSlot) + 4 + 4));
      gcount = &P.tbl_count[S->table];
    };
    if (nchunks > 1) bind_table();

    auto put_self = [&]() {  // first chunk owns the self term (grank.h:101 / mccompletepathv2.h:226)
      if (tid == 0 && cb == rb) {
        const unsigned long long x = init_mode ? 0ull : (unsigned long long)__double2ll_rn(self0 * scale);
        if ((unsigned)self_id < (unsigned)H) {
          if (x) s_dense[self_id] = make_uint2((unsigned int)x, (unsigned int)(x >> 32));
          else s_zbits[self_id >> 5] |= 1u << (self_id & 31);
        } else {
          const unsigned int h = hash_key(self_id) & (TCAP - 1);
          t_keys[h] = self_id;
          t_acc[h] = make_uint2((unsigned int)x, (unsigned int)(x >> 32));
          t_list[0] = (unsigned short)h;
          S->tcount = 1;
        }
      }
    };
    put_self();
    __syncthreads();

    // ---- accumulate ----
    unsigned long long merged = 0;
    auto slow_contribute = [&](int k, unsigned long long x, bool spill_ok) {
      if ((unsigned)k < (unsigned)H) {  // dense label with a zero word: remember that it was touched
        atomicOr(&s_zbits[k >> 5], 1u << (k & 31));
        return;
      }
      if (!tail_add(k, x)) {
        if (spill_ok) gtable_add(gslots, glist, gcount, S->gmask, S->gidentity, k, x);
        else S->spilled = 1;
      }
    };
    // pass 0: single pass (dense + tail table, spilling to the node's global table when `spill_ok`)
    // pass 1: dense labels exactly, tail labels into the sketch          (R > 0, single-item nodes)
    // pass 2: tail labels of surviving buckets exactly into the tail table
    auto accumulate = [&](int pass, bool spill_ok) {
      merged = 0;
      bool tail_seen = false;
      // warp w merges successors w, w+NW, ... of a tile with the next two baskets already in flight
      auto slot_of = [&](uint32_t cc) -> const unsigned char* {
        return M.buf[((cc >> COL_COLOUR_SHIFT) & 1u) ? read_slot1 : read_slot0] + (size_t)(cc & COL_POS_MASK) * slot_bytes(Lp);
      };
      // Entries that miss the fast paths (tail labels, zero products) are rare but slow (hash probe, CAS); taken
      // inline they would stall the whole warp behind two or three lanes. They are parked in a per-warp queue and
      // drained 32 at a time with every lane busy.
      int qn = 0;
      int* qk = q_key + w * 64;
      unsigned long long* qv = q_val + w * 64;
      auto drain = [&](int cnt) {  // the last `cnt` (<= 32) queued entries, one per lane
        __syncwarp();
        bool ins = false;
        unsigned int slot = 0;
        if (lane < cnt) {
          const int k = qk[qn - cnt + lane];
          const unsigned long long x = qv[qn - cnt + lane];
          if ((unsigned)k < (unsigned)H) atomicOr(&s_zbits[k >> 5], 1u << (k & 31));  // dense label, zero product: touched
          else if (!tail_add(k, x)) {
            if (spill_ok) ins = gtable_upsert(gslots, S->gmask, S->gidentity, k, x, &slot);
            else S->spilled = 1;
          }
        }
        if (spill_ok) glist_append(ins, slot, glist, gcount);
        qn -= cnt;
        __syncwarp();
      };
      auto contribute = [&](int k, unsigned long long xf) {  // called by all 32 lanes; k < 0: nothing
        const bool dense = (unsigned)k < (unsigned)H;
        bool slow;
        if (pass == 2) {
          slow = false;
          if (k >= H) {
            const unsigned int b = hash_key(k) & (unsigned)(R > 0 ? R - 1 : 0);
            const int ps = s_pslot[b];
            slow = ((s_alive[b >> 5] >> (b & 31)) & 1u) && !(ps && s_pkey[ps - 1] == k);  // pre-slot labels are exact already
          }
        } else {
          const bool fast = dense && xf != 0ull;
          if (fast) fixed_add_shared(&s_dense[k], xf);
          slow = !fast && k >= 0;
        }
        const unsigned m = __ballot_sync(FULL, slow);
        if (m) {
          if (slow) { const int pos = qn + __popc(m & ((1u << lane) - 1u)); qk[pos] = k; qv[pos] = xf; }
          qn += __popc(m);
          if (qn >= 32) drain(32);
        }
      };
      for (int t0 = 0; t0 < clen; t0 += COLCAP) {
        const int tlen = clen - t0 < COLCAP ? clen - t0 : COLCAP;
        __syncthreads();
        // stage the tile's column words (coalesced) so that the basket prefetch below never waits on them
        for (int j = tid; j < tlen; j += THREADS) s_col[j] = M.g.col[cb + t0 + j];
        __syncthreads();
        if (init_mode) {
          // grank.h:79-80: every occurrence of a successor adds `factor`; here: multiplicity += 1
          for (int j = tid; j < tlen; j += THREADS) {
            const uint32_t c = s_col[j];
            const int k = (c & COL_SINK) ? (int)(c & ~COL_SINK) : M.g.label[c & COL_POS_MASK];
            if ((unsigned)k < (unsigned)H) fixed_add_shared(&s_dense[k], 1ull);
            else slow_contribute(k, 1ull, spill_ok);
            merged++;
          }
          continue;
        }
        auto fetch = [&](int j, BasketFrag* fr) {
          fr->id = make_int4(-1, -1, -1, -1);
          fr->sa = fr->sb = make_double2(0.0, 0.0);
          if (j < tlen) {
            const uint32_t cc = s_col[j];
            if (!(cc & COL_SINK) && lane < groups) {
              // pass 2 decides on the labels alone: the scores (2/3 of the bytes) are fetched only by lanes that hold a survivor
              if (pass == 2) fr->id = __ldg(reinterpret_cast<const int4*>(slot_of(cc)) + lane);
              else load_frag_all(slot_of(cc), Lp, lane, fr);
            }
          }
        };
        BasketFrag f0, f1;
        fetch(w, &f0);
        fetch(w + NW, &f1);
        for (int j = w; j < tlen; j += NW) {
          BasketFrag f2;
          fetch(j + 2 * NW, &f2);
          const uint32_t c = s_col[j];
          if (c & COL_SINK) {
            const double x = (M.mode == MODE_GRANK) ? M.self_grank : 1.0;
            const int k = (int)(c & ~COL_SINK);
            const unsigned long long xf = (unsigned long long)__double2ll_rn(x * fscale);
            if (pass == 1) {
              if (lane == 0) {
                if (k < H) {
                  if (xf) fixed_add_shared(&s_dense[k], xf);
                  else atomicOr(&s_zbits[k >> 5], 1u << (k & 31));
                } else {
                  tail_seen = true;
                  const unsigned int b = hash_key(k) & (unsigned)(R > 0 ? R - 1 : 0);
                  const int ps = s_pslot[b];
                  if (ps && s_pkey[ps - 1] == k) { if (xf) fixed_add_shared(&s_pacc[ps - 1], xf); s_ptouch[ps - 1] = 1; }
                  else if (xf) fixed_add_shared(&s_sk[b], xf);
                }
              }
            } else {
              contribute(lane == 0 ? k : -1, xf);
            }
            merged += (lane == 0);
          } else {
            for (int g0 = 0; g0 < groups; g0 += 32) {
              BasketFrag fr;
              if (g0 == 0) fr = f0;
              else {
                fr.id = make_int4(-1, -1, -1, -1);
                fr.sa = fr.sb = make_double2(0.0, 0.0);
                if (g0 + lane < groups) {
                  if (pass == 2) fr.id = __ldg(reinterpret_cast<const int4*>(slot_of(c)) + g0 + lane);
                  else load_frag_all(slot_of(c), Lp, g0 + lane, &fr);
                }
              }
              const int ids[4] = {fr.id.x, fr.id.y, fr.id.z, fr.id.w};
              double xs[4] = {fr.sa.x, fr.sa.y, fr.sb.x, fr.sb.y};
              if (pass == 1) {
                // lean pass: no queue, no votes -- every entry is one or two shared-memory atomics on the dense word, on the
                // exact accumulator of an old-basket label, or on its sketch bucket
#pragma unroll
                for (int e = 0; e < 4; e++) {
                  const int k = ids[e];
                  if (k < 0) continue;
                  merged++;
                  const unsigned long long xf = (unsigned long long)__double2ll_rn(xs[e] * fscale);
                  if (k < H) {
                    if (xf) fixed_add_shared(&s_dense[k], xf);
                    else atomicOr(&s_zbits[k >> 5], 1u << (k & 31));
                  } else {
                    tail_seen = true;
                    const unsigned int b = hash_key(k) & (unsigned)(R > 0 ? R - 1 : 0);
                    const int ps = s_pslot[b];
                    if (ps && s_pkey[ps - 1] == k) {
                      if (xf) fixed_add_shared(&s_pacc[ps - 1], xf);
                      s_ptouch[ps - 1] = 1;
                    } else if (xf) {
                      fixed_add_shared(&s_sk[b], xf);
                    }
                  }
                }
              } else if (pass == 2) {
                // most baskets hold no surviving tail label: one vote skips them
                bool mine = false;
#pragma unroll
                for (int e = 0; e < 4; e++) {
                  const int k = ids[e];
                  if (k >= H) { const unsigned int b = hash_key(k) & (unsigned)(R > 0 ? R - 1 : 0); mine |= (s_alive[b >> 5] >> (b & 31)) & 1u; }
                  merged += (k >= 0);
                }
                if (__any_sync(FULL, mine)) {
                  if (mine) {
                    const double2* sc = reinterpret_cast<const double2*>(slot_of(c) + (size_t)Lp * 4);
                    const double2 sa = __ldg(sc + g0 + lane), sb = __ldg(sc + (Lp >> 2) + g0 + lane);
                    xs[0] = sa.x; xs[1] = sa.y; xs[2] = sb.x; xs[3] = sb.y;
                  }
#pragma unroll
                  for (int e = 0; e < 4; e++) contribute(ids[e], (unsigned long long)__double2ll_rn(xs[e] * fscale));
                }
              } else {
#pragma unroll
                for (int e = 0; e < 4; e++) {
                  contribute(ids[e], (unsigned long long)__double2ll_rn(xs[e] * fscale));
                  merged += (ids[e] >= 0);
                }
              }
            }
          }
          f0 = f1;
          f1 = f2;
        }
      }
      if (qn > 0) drain(qn);
      return tail_seen;
    };
    auto reset_shared = [&]() {  // drop the partial sums of this item
      __syncthreads();
      for (int i = tid; i < H; i += THREADS) s_dense[i] = make_uint2(0u, 0u);
      for (int i = tid; i < H / 32; i += THREADS) s_zbits[i] = 0u;
      const int dirty = S->tcount;
      for (int i = tid; i < dirty; i += THREADS) {
        const int s = t_list[i];
        t_keys[s] = KEY_EMPTY;
        t_acc[s] = make_uint2(0u, 0u);
      }
      __syncthreads();
      if (tid == 0) { S->tcount = 0; S->spilled = 0; s_requeue++; }
    };
    // compaction of the candidates held in shared memory into (score bits, label) arrays; `from_dense` / `from_tail`
    // select the sources, positions continue from S->ncand
    const double base_self = init_mode ? M.self_grank : 0.0;
    bool dropped_local = false;  // this thread left a candidate out of the compact arrays (below the lower bound)
    auto compact = [&](bool from_dense, bool from_tail, int tail_from, unsigned long long theta, bool from_pre = false) {
      if (from_dense)
        for (int i0 = 0; i0 < H; i0 += THREADS) {
          const int i = i0 + tid;
          const uint2 a = s_dense[i];
          bool ok = dense_touched(i, a);
          unsigned long long bits = 0ull;
          if (ok) {
            bits = (unsigned long long)__double_as_longlong(
                par_score(((unsigned long long)a.y << 32) | a.x, init_mode, inv, mult, (init_mode && i == self_id) ? base_self : 0.0));
            ok = bits >= theta;  // below the lower bound of the cut: cannot be kept
            dropped_local |= !ok;
          }
          const unsigned m = __ballot_sync(FULL, ok);
          if (m) {
            int basep = 0;
            if (lane == (int)(__ffs(m) - 1)) basep = atomicAdd(&S->ncand, __popc(m));
            basep = __shfl_sync(FULL, basep, __ffs(m) - 1);
            if (ok) {
              const int pos = basep + __popc(m & ((1u << lane) - 1u));
              if (pos < CMAX) { c_bits[pos] = bits; c_id[pos] = i; }
            }
          }
        }
      if (from_pre && tid < 32 * ((PAR_PRE + 31) / 32)) {  // the first PAR_PRE/32 warps, whole warps
        bool ok = tid < PAR_PRE && s_pkey[tid] >= 0 && s_ptouch[tid];
        unsigned long long bits = 0ull;
        if (ok) {
          bits = (unsigned long long)__double_as_longlong(par_score(((unsigned long long)s_pacc[tid].y << 32) | s_pacc[tid].x, false, inv, mult, 0.0));
          ok = bits >= theta;
          dropped_local |= !ok;
        }
        const unsigned m = __ballot_sync(FULL, ok);
        if (m) {
          int basep = 0;
          if (lane == (int)(__ffs(m) - 1)) basep = atomicAdd(&S->ncand, __popc(m));
          basep = __shfl_sync(FULL, basep, __ffs(m) - 1);
          if (ok) {
            const int pos = basep + __popc(m & ((1u << lane) - 1u));
            if (pos < CMAX) { c_bits[pos] = bits; c_id[pos] = s_pkey[tid]; }
          }
        }
      }
      if (from_tail) {
        const int nt0 = S->tcount;
        for (int i0 = tail_from; i0 < nt0; i0 += THREADS) {
          const int i = i0 + tid;
          bool ok = i < nt0;
          unsigned long long bits = 0ull;
          int id = 0;
          if (ok) {
            const int sl = t_list[i];
            id = t_keys[sl];
            bits = (unsigned long long)__double_as_longlong(par_score(((unsigned long long)t_acc[sl].y << 32) | t_acc[sl].x, init_mode, inv,
                                                                      mult, (init_mode && id == self_id) ? base_self : 0.0));
            ok = bits >= theta;
            dropped_local |= !ok;
          }
          const unsigned m = __ballot_sync(FULL, ok);
          if (m) {
            int basep = 0;
            if (lane == (int)(__ffs(m) - 1)) basep = atomicAdd(&S->ncand, __popc(m));
            basep = __shfl_sync(FULL, basep, __ffs(m) - 1);
            if (ok) {
              const int pos = basep + __popc(m & ((1u << lane) - 1u));
              if (pos < CMAX) { c_bits[pos] = bits; c_id[pos] = id; }
            }
          }
        }
      }
    };
    // Lower bound of the cut from the previous basket (warm start): when the old basket is full, the new L-th largest
    // score is almost always above half the old one. Candidates below theta are dropped while compacting, which is
    // exact as long as at least L candidates remain (the L-th largest of a subset bounds the cut from below);
    // otherwise theta is relaxed and the compaction repeated.
    double theta0 = 0.0;
    if (!init_mode) {  // (MC combine rounds too: the node's previous basket is on the same score scale)
      const unsigned char* old = M.buf[write_slot ^ 1] + (size_t)p * slot_bytes(Lp);
      const int* oid = reinterpret_cast<const int*>(old);
      const double* osc = reinterpret_cast<const double*>(old + (size_t)Lp * 4);
      unsigned long long mn = ~0ull;
      long long cntv = 0;
      for (int i = tid; i < Lp; i += THREADS)
        if (oid[i] >= 0) { const unsigned long long b = (unsigned long long)__double_as_longlong(osc[score_index(i, Lp)]); mn = b < mn ? b : mn; cntv++; }
      block_reduce_min_sum(mn, cntv, S->red_a);
      if (cntv >= L) theta0 = __longlong_as_double((long long)mn);
    }
    // compaction with the warm-start filter; returns the candidate count (S->ncand), positions start at S->ncand = 0
    double theta_used = 0.0;  // lower bound the compact arrays were filtered with
    auto compact_filtered = [&](bool from_dense, bool from_tail, bool from_pre = false) -> int {
      double th = theta0 * 0.5;
      for (int attempt = 0;; attempt++) {
        theta_used = th;
        dropped_local = false;
        __syncthreads();
        if (tid == 0) S->ncand = 0;
        __syncthreads();
        compact(from_dense, from_tail, 0, (unsigned long long)__double_as_longlong(th), from_pre);
        __syncthreads();
        const int c = S->ncand;
        if (c >= L || th == 0.0) return c;
        th = attempt == 0 ? th * 0.125 : 0.0;
      }
    };
    // a node expected to outgrow the shared-memory structures binds its global table up front
    // per-node memory of the previous update (M.ncand[p]): NEEDS_GLOBAL = outgrew shared memory -> bind a global table
    // up front. Single-item nodes of the big class take the two-pass scheme (P.use_sketch; measured faster than the
    // single pass from R-MAT-16 to R-MAT-22, the more so the smaller the share of the label space the dense range covers).
    constexpr int NEEDS_GLOBAL = 0x3fffffff;
    const int hint = M.ncand[p];
    const bool two_pass = R > 0 && P.use_sketch && nchunks == 1 && !init_mode && hint != NEEDS_GLOBAL;
    if (!two_pass && S->table < 0 && hint == NEEDS_GLOBAL) bind_table();
    PROF_MARK(1);
    int n = 0;
    int my_bucket = -1;  // sketch bucket whose exact pre-slot this thread owns (two-pass only)
    if (two_pass) {
      // The old basket is the best predictor of the new one: its tail labels get exact accumulators of their own for
      // pass 1 (one per sketch bucket, first come first served), so that tau (below) is the L-th largest of (dense labels
      // + old-basket labels) -- close to the final cut -- and pass 2 only has to pick up the few new entrants.
      for (int i = tid; i < R; i += THREADS) s_sk[i] = make_uint2(0u, 0u);
      if (tid < PAR_PRE) {
        const int* old_ids = reinterpret_cast<const int*>(M.buf[write_slot ^ 1] + (size_t)p * slot_bytes(Lp));
        const int k = tid < Lp ? old_ids[tid] : -1;
        s_pkey[tid] = -1;
        s_pacc[tid] = make_uint2(0u, 0u);
        s_ptouch[tid] = 0;
        // (the node's own label keeps its place in the tail table, where put_self started it; two old labels in one
        // bucket: the last writer owns it, the other stays in the sketch)
        if (k >= H && k != self_id) { my_bucket = (int)(hash_key(k) & (unsigned)(R > 0 ? R - 1 : 0)); s_pslot[my_bucket] = (unsigned char)(tid + 1); }
      }
      __syncthreads();
      if (tid < PAR_PRE && my_bucket >= 0) {
        if (s_pslot[my_bucket] == (unsigned char)(tid + 1)) {
          const int* old_ids = reinterpret_cast<const int*>(M.buf[write_slot ^ 1] + (size_t)p * slot_bytes(Lp));
          s_pkey[tid] = old_ids[tid];
        } else {
          my_bucket = -1;
        }
      }
      __syncthreads();
      const int tail_seen = __syncthreads_or(accumulate(1, false) ? 1 : 0);
      PROF_MARK(2);
      n = compact_filtered(true, false, true);
      __syncthreads();
      if (n > CMAX) {
        if (tid == 0) S->spilled = 1;
      } else {
        unsigned long long tau = 0ull;
        if (tail_seen) {
          // tau = L-th largest exact score so far (0 when there are not more than L candidates)
          if (n > L) {
            bool tie;
            int krem;
            auto keyfn = [&](int i) { return c_bits[i]; };
            auto all_ = [](int) { return true; };
            tau = block_radix_select(n, L, keyfn, all_, S, &tie, &krem);
          }
          for (int i = tid; i < R / 32; i += THREADS) s_alive[i] = 0u;
          __syncthreads();
          int any_alive = 0;
          for (int i = tid; i < R; i += THREADS) {
            const uint2 a = s_sk[i];
            if ((a.x | a.y) == 0u && tau != 0ull) continue;
            const unsigned long long bits =
                (unsigned long long)__double_as_longlong(par_score(((unsigned long long)a.y << 32) | a.x, false, inv, mult, 0.0));
            if (tau == 0ull || bits >= tau) { atomicOr(&s_alive[i >> 5], 1u << (i & 31)); any_alive = 1; }
          }
          if (tid == 0 && self_id >= H) {  // contributions to the node's own (tail) label went to the sketch: always exact in pass 2
            const unsigned int b = hash_key(self_id) & (unsigned)(R > 0 ? R - 1 : 0);
            atomicOr(&s_alive[b >> 5], 1u << (b & 31));
            any_alive = 1;
          }
          any_alive = __syncthreads_or(any_alive);
          PROF_MARK(3);
          if (any_alive) {
            accumulate(2, false);
            __syncthreads();
          }
        }
        if (!S->spilled) {
          compact(false, true, 0, tau);  // the tail table: the node itself (if a tail label) and the survivors, all exact now
          __syncthreads();
          n = S->ncand;
          if (n > CMAX && tid == 0) S->spilled = 1;
        }
      }
      __syncthreads();
      PROF_MARK(4);
    } else {
      accumulate(0, S->table >= 0);
      __syncthreads();
      // single-chunk node that stayed in shared memory: compact the candidates (score bits, label)
      if (S->table < 0 && !S->spilled) {
        n = compact_filtered(true, true);
        if (n > CMAX && tid == 0) S->spilled = 1;  // more candidates than the compact arrays hold: take the global path
        __syncthreads();
      }
    }
    PROF_MARK(2);
    if (S->spilled) {
      // outgrew shared memory: drop the partial sums, bind a global table and run the item again in a single pass
      reset_shared();
      bind_table();
      put_self();
      __syncthreads();
      accumulate(0, true);
      __syncthreads();
      PROF_MARK(3);
    }

    const int nt = S->tcount;  // occupied tail slots
    const bool use_global = S->table >= 0;
    int kept = 0, old_cnt = 0;
    bool finalize = !use_global;  // single chunk, nothing spilled: finish from shared memory
    if (use_global) {
      // flush the shared accumulators into the node's global table (and leave them clean)
      for (int i = tid; i < H; i += THREADS) {
        const uint2 a = s_dense[i];
        bool ins = false;
        unsigned int slot = 0;
        if (dense_touched(i, a)) {
          ins = gtable_upsert(gslots, S->gmask, S->gidentity, i, ((unsigned long long)a.y << 32) | a.x, &slot);
          s_dense[i] = make_uint2(0u, 0u);
        }
        glist_append(ins, slot, glist, gcount);
      }
      for (int i0 = 0; i0 < nt; i0 += THREADS) {
        const int i = i0 + tid;
        bool ins = false;
        unsigned int slot = 0;
        if (i < nt) {
          const int s = t_list[i];
          ins = gtable_upsert(gslots, S->gmask, S->gidentity, t_keys[s], ((unsigned long long)t_acc[s].y << 32) | t_acc[s].x, &slot);
        }
        glist_append(ins, slot, glist, gcount);
      }
      __threadfence();
      __syncthreads();
      for (int i = tid; i < H / 32; i += THREADS) s_zbits[i] = 0u;
      if (tid == 0) {
        const unsigned int done = atomicAdd(&P.node_done[p], 1u) + 1u;
        S->is_last = done == (unsigned)nchunks;
        if (S->is_last) __threadfence();
      }
      __syncthreads();
      finalize = S->is_last != 0;
      PROF_MARK(4);
    }

    if (finalize) {
      // candidates: compact arrays (score bits, label) -- in shared memory, or gathered once from the global table
      const unsigned long long* kb = c_bits;
      const int* ki = c_id;
      if (use_global) {
        n = (int)*reinterpret_cast<volatile unsigned int*>(gcount);
        for (int i = tid; i < n; i += THREADS) {
          const GSlot g = gslots[glist[i]];
          cids[i] = g.key;
          gbits[i] = (unsigned long long)__double_as_longlong(
              par_score(g.acc, init_mode, inv, mult, (init_mode && g.key == self_id) ? base_self : 0.0));
        }
        __syncthreads();
        kb = gbits;
        ki = cids;
      }
      Threshold th;
      // n <= L: every candidate that reached the filter's bound is kept and nothing below it is (norm1 asks about those)
      th.bits = use_global ? 0ull : (unsigned long long)__double_as_longlong(theta_used);
      th.id_max = 0x7fffffff;
      kept = n;
      auto all = [](int) { return true; };
      if (n > L) {
        kept = L;
        bool tie;
        int krem;
        auto keyfn = [&](int i) { return kb[i]; };
        int ntied = 0;
        th.bits = block_radix_select(n, L, keyfn, all, S, &tie, &krem, &ntied);
        if (tie) {
          const unsigned long long tb = th.bits;
          if (ntied <= 32) {
            auto densefn = [&](int i) { return dense_of[ki[i]]; };
            th.id_max = block_small_tie_cut(n, krem, tb, keyfn, all, densefn, S);
          } else {
            auto idkey = [&](int i) { return (unsigned long long)(0x7fffffff - dense_of[ki[i]]); };
            auto tied = [&](int i) { return kb[i] == tb; };
            bool tie2;
            int krem2;
            const unsigned long long tid_key = block_radix_select(n, krem, idkey, tied, S, &tie2, &krem2);
            th.id_max = 0x7fffffff - (int)tid_key;
          }
          s_ties += (tid == 0);
        }
        s_truncs += (tid == 0);
      } else if (n == L && !use_global) {
        if (__syncthreads_or(dropped_local ? 1 : 0)) s_truncs += (tid == 0);  // the filter already cut the rest away
      }
      auto selected = [&](unsigned long long bits, int label) -> bool {
        return bits > th.bits || (bits == th.bits && (th.id_max == 0x7fffffff || dense_of[label] <= th.id_max));
      };
      PROF_MARK(5);
      // ---- write B'_v ----
      unsigned char* out = M.buf[write_slot] + (size_t)p * slot_bytes(Lp);
      int* out_ids = reinterpret_cast<int*>(out);
      double* out_sc = reinterpret_cast<double*>(out + (size_t)Lp * 4);
      if (tid == 0) S->out_pos = 0;
      __syncthreads();
      long long dsum = 0;
      for (int i = tid; i < n; i += THREADS) {
        const unsigned long long bits = kb[i];
        const int id = ki[i];
        if (selected(bits, id)) {
          const int pos = atomicAdd(&S->out_pos, 1);
          const double v = __longlong_as_double((long long)bits);
          out_ids[pos] = id;
          out_sc[score_index(pos, Lp)] = v;  // hub path: no post-scale (already multiplied by f)
          dsum += fix_norm(v);
        }
      }
      for (int i = kept + tid; i < Lp; i += THREADS) out_ids[i] = KEY_EMPTY;
      // ---- norm1 against the old basket (pprInternal.h:147-165) ----
      if (M.do_norm) {
        const unsigned char* old = M.buf[write_slot ^ 1] + (size_t)p * slot_bytes(Lp);
        for (int g = tid; g < groups; g += THREADS) {
          BasketFrag fr;
          load_frag(old, Lp, g, &fr);
          const int ids[4] = {fr.id.x, fr.id.y, fr.id.z, fr.id.w};
          const double xs[4] = {fr.sa.x, fr.sa.y, fr.sb.x, fr.sb.y};
#pragma unroll
          for (int e = 0; e < 4; e++) {
            if (ids[e] >= 0) {
              old_cnt++;
              bool found = false;
              double nv_ = 0.0;
              if (use_global) {
                const int s = gtable_find(gslots, S->gmask, S->gidentity, ids[e]);
                if (s >= 0) { found = true; nv_ = par_score(gslots[s].acc, false, inv, mult, 0.0); }
              } else if ((unsigned)ids[e] < (unsigned)H) {
                const uint2 a = s_dense[ids[e]];
                found = dense_touched(ids[e], a);
                nv_ = par_score(((unsigned long long)a.y << 32) | a.x, false, inv, mult, 0.0);
              } else if (R > 0 && two_pass && s_pslot[hash_key(ids[e]) & (unsigned)(R > 0 ? R - 1 : 0)] &&
                         s_pkey[s_pslot[hash_key(ids[e]) & (unsigned)(R > 0 ? R - 1 : 0)] - 1] == ids[e]) {
                const int ps = s_pslot[hash_key(ids[e]) & (unsigned)(R > 0 ? R - 1 : 0)] - 1;
                found = s_ptouch[ps] != 0;
                nv_ = par_score(((unsigned long long)s_pacc[ps].y << 32) | s_pacc[ps].x, false, inv, mult, 0.0);
              } else {
                for (unsigned int h = hash_key(ids[e]) & (TCAP - 1);; h = (h + 1) & (TCAP - 1)) {
                  const int cur = t_keys[h];
                  if (cur == ids[e]) { found = true; nv_ = par_score(((unsigned long long)t_acc[h].y << 32) | t_acc[h].x, false, inv, mult, 0.0); break; }
                  if (cur == KEY_EMPTY) break;
                }
              }
              const bool in_new = found && selected((unsigned long long)__double_as_longlong(nv_), ids[e]);
              if (in_new) dsum += fix_norm(fabs(nv_ - xs[e])) - fix_norm(nv_);
              else dsum += fix_norm(xs[e]);
            }
          }
        }
        long long oc = old_cnt;
        block_reduce_sum2(dsum, oc, S->red_a);
        old_cnt = (int)oc;
        if (tid == 0 && dsum > 0) atomicMax(&st->cur_max, dsum);
      }
      __syncthreads();
      publish_slot(M.peers, write_slot, (size_t)p * slot_bytes(Lp), slot_bytes(Lp), tid, THREADS, p);
      if (use_global) {
        // leave the pool table clean and hand it back
        for (int i = tid; i < n; i += THREADS) {
          const unsigned int h = glist[i];
          gslots[h].key = KEY_EMPTY;
          gslots[h].acc = 0ull;
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
          *gcount = 0u;
          P.node_done[p] = 0u;
          P.node_tbl[p] = 0u;
          __threadfence();
          atomicExch(&P.tbl_inuse[S->table], 0u);
        }
      } else {
        for (int i = tid; i < H; i += THREADS) s_dense[i] = make_uint2(0u, 0u);
        for (int i = tid; i < H / 32; i += THREADS) s_zbits[i] = 0u;
      }
      if (tid == 0) {
        M.ncand[p] = (use_global && !init_mode) ? NEEDS_GLOBAL : 0;  // a node that needed its global table binds it up front next time
        s_cands += (unsigned long long)n;
        s_nodes += 1;
        s_bytes += 12ull * (unsigned long long)old_cnt + 12ull * (unsigned long long)kept + 4ull + 16ull;
      }
    }
    PROF_MARK(6);
    if (R > 0 && my_bucket >= 0) s_pslot[my_bucket] = 0;  // leave the bucket map clean for the next node
    // ---- reset the tail table through its list ----
    for (int i = tid; i < nt; i += THREADS) {
      const int s = t_list[i];
      t_keys[s] = KEY_EMPTY;
      t_acc[s] = make_uint2(0u, 0u);
    }
    merged = (unsigned long long)block_reduce_sum_ll((long long)merged, S->red_a);
    if (tid == 0) {
      s_merged += merged;
      s_edges += (unsigned long long)clen;
      s_bytes += 12ull * merged + 4ull * (unsigned long long)clen;
      S->tcount = 0;
      S->spilled = 0;
      S->table = -1;
    }
    __syncthreads();
    PROF_MARK(7);
  }
  if (P.prof && tid == 0)
    for (int i = 0; i < 8; i++) atomicAdd(&P.prof[(size_t)blockIdx.x * 8 + i], pc[i]);
  if (tid == 0) {
    if (s_nodes) atomicAdd(&st->node_iters, s_nodes);
    if (s_edges) atomicAdd(&st->edge_reads, s_edges);
    if (s_merged) atomicAdd(&st->merged, s_merged);
    if (s_cands) atomicAdd(&st->cands, s_cands);
    if (s_truncs) atomicAdd(&st->truncs, s_truncs);
    if (s_ties) atomicAdd(&st->ties, s_ties);
    if (s_bytes) atomicAdd(&st->abytes, s_bytes);
    if (s_requeue) atomicAdd(&st->requeues, s_requeue);
  }
}

}  // namespace pprb200
