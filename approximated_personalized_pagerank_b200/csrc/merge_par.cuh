// merge_par.cuh -- the order-free merge step for nodes above the hub threshold: one CTA per work chunk.
//
// Same operator as merge_seq.cuh (grank.h:96-126, mccompletepathv2.h:211-250) but every contribution
// x * d/outdeg is rounded once to 2^-62 (MC: 2^-59) fixed point and summed as an integer, so the result does
// not depend on the order of the additions: all warps of a CTA -- and, for nodes split into several chunks,
// several CTAs -- accumulate concurrently with atomics and still produce bit-identical sums
// (oracle/ppr_oracle.c applies the same rounding). |difference to the reference's fma chain| <= outdeg * 2^-62.
//
// Accumulator hierarchy:
//   1. shared memory open-addressing table of the CTA, PAR_CAP slots: int32 key + 2 x uint32 fixed-point
//      words (ATOMS.ADD.32 on the low word, carry into the high word) + list of occupied slots;
//      filled first come first served up to PAR_LIMIT distinct keys;
//   2. when the node has more candidates than that, or is split into several chunks, a global (L2-resident)
//      table taken from a small pool: entries that miss the full shared table are added there directly
//      (CAS on the key, RED.ADD.64 on the value) and at the end of the chunk the shared table is flushed into it.
//      The CTA that completes the last chunk of a node selects the top-L from the global table.
#pragma once
#include "merge_seq.cuh"

namespace pprb200 {

// Two instantiations: <16384 slots, 512 threads> (1 CTA/SM) for big nodes and chunks of hubs, and
// <4096 slots, 128 threads> (3 CTAs/SM) for mid-degree nodes. PAR_LIMIT distinct keys are admitted to the shared
// table (concurrent inserts may overshoot by < THREADS).

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
  unsigned int capmax;         // slots per pool table (power of two >= 2(n+1))
  int n_tables;
  unsigned int* tbl_inuse;     // [n_tables]
  unsigned int* tbl_count;     // [n_tables] distinct keys inserted
  unsigned int* node_tbl;      // [M] 0 none, 1 being acquired, else table index + 2
  unsigned int* node_done;     // [M] chunks completed
  int n_ids;                   // dense id space (identity hashing when the table covers it)
};

struct ParShared {
  int count;            // occupied slots of the shared table
  int spilled;          // some entry went to the global table
  int table;            // global table index of the current node (-1 none)
  unsigned int gmask;
  int gidentity;
  int is_last;
  int out_pos;
  unsigned int item;
  unsigned long long red_a[16], red_b[16];
  unsigned long long sel_prefix;
  int sel_digit, sel_above, sel_cnt;
  unsigned int hist[256];
};

__device__ __forceinline__ unsigned long long block_reduce_min_ull(unsigned long long v, unsigned long long* scratch) {
  v = warp_min_ull(v);
  if (lane_id() == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  unsigned long long r = scratch[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); i++) r = scratch[i] < r ? scratch[i] : r;
  __syncthreads();
  return r;
}
__device__ __forceinline__ unsigned long long block_reduce_max_ull(unsigned long long v, unsigned long long* scratch) {
  v = warp_max_ull(v);
  if (lane_id() == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  unsigned long long r = scratch[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); i++) r = scratch[i] > r ? scratch[i] : r;
  __syncthreads();
  return r;
}
__device__ __forceinline__ long long block_reduce_sum_ll(long long v, unsigned long long* scratch) {
  v = warp_sum_ll(v);
  if (lane_id() == 0) scratch[threadIdx.x >> 5] = (unsigned long long)v;
  __syncthreads();
  long long r = 0;
  for (int i = 0; i < (int)(blockDim.x >> 5); i++) r += (long long)scratch[i];
  __syncthreads();
  return r;
}

// CTA-wide version of warp_radix_select (device_common.cuh); same contract.
template <typename KeyFn, typename PredFn>
__device__ unsigned long long block_radix_select(int n, int k, KeyFn key, PredFn pred, ParShared* S, bool* tie, int* krem) {
  const int tid = threadIdx.x, T = blockDim.x;
  unsigned long long lo = ~0ull, hi = 0ull;
  long long cnt = 0;
  for (int i = tid; i < n; i += T)
    if (pred(i)) { const unsigned long long b = key(i); lo = b < lo ? b : lo; hi = b > hi ? b : hi; cnt++; }
  lo = block_reduce_min_ull(lo, S->red_a);
  hi = block_reduce_max_ull(hi, S->red_a);
  *tie = false;
  *krem = 0;
  if (lo == hi) {
    cnt = block_reduce_sum_ll(cnt, S->red_a);
    if (cnt > k) { *tie = true; *krem = k; }
    return lo;
  }
  const int top = 63 - __clzll((long long)(lo ^ hi));
  unsigned long long known = (top == 63) ? 0ull : ~((2ull << top) - 1ull);
  unsigned long long prefix = hi & known;
  int shift = top - 7 > 0 ? top - 7 : 0;
  for (;;) {
    for (int i = tid; i < 256; i += T) S->hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += T)
      if (pred(i)) {
        const unsigned long long b = key(i);
        if ((b & known) == prefix) atomicAdd(&S->hist[(unsigned)(b >> shift) & 0xffu], 1u);
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
    if (shift == 0) { *tie = true; *krem = k; return prefix; }
    shift = shift - 8 > 0 ? shift - 8 : 0;
  }
}

// fixed-point word -> score. init mode: the word is a multiplicity m and the score is `base + mult + ... + mult`
// (m additions, grank.h:79-80); otherwise score = word * 2^-62 (GRank) / 2^-59 (MC).
__device__ __forceinline__ double par_score(unsigned long long acc, bool init_mode, double inv, double mult, double base) {
  if (!init_mode) return (double)(long long)acc * inv;
  double a = base;
  for (unsigned long long r = 0; r < acc; r++) a += mult;
  return a;
}

__device__ __forceinline__ void gtable_add(GSlot* slots, unsigned int* glist, unsigned int* gcount, unsigned int mask,
                                           int identity, int k, unsigned long long x) {
  unsigned int h = (identity ? (unsigned int)k : hash_key(k)) & mask;
  for (;;) {
    const int cur = *reinterpret_cast<volatile int*>(&slots[h].key);
    if (cur == k) break;
    if (cur == KEY_EMPTY) {
      const int old = atomicCAS(&slots[h].key, KEY_EMPTY, k);
      if (old == KEY_EMPTY) { const unsigned int pos = atomicAdd(gcount, 1u); glist[pos] = h; break; }
      if (old == k) break;
    }
    h = (h + 1) & mask;
  }
  if (x) atomicAdd(&slots[h].acc, x);
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

template <int PAR_CAP, int PAR_THREADS>
__global__ void __launch_bounds__(PAR_THREADS) merge_par_kernel(ParParams P) {
  constexpr int PAR_LIMIT = PAR_CAP * 13 / 16 - PAR_THREADS;
  extern __shared__ __align__(16) unsigned char smem[];
  const MergeParams& M = P.M;
  RunState* st = M.st;
  if (!st->active) return;
  int* s_keys = reinterpret_cast<int*>(smem);
  uint2* s_acc = reinterpret_cast<uint2*>(smem + (size_t)PAR_CAP * 4);
  unsigned short* s_list = reinterpret_cast<unsigned short*>(smem + (size_t)PAR_CAP * 12);
  ParShared* S = reinterpret_cast<ParShared*>(smem + (size_t)PAR_CAP * 14);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  constexpr int NW = PAR_THREADS / 32;
  const int Lp = M.Lp, groups = Lp >> 2, L = M.L;
  const bool init_mode = M.init_mode != 0;
  const double scale = (M.mode == MODE_GRANK) ? GRANK_HUB_SCALE : MC_HUB_SCALE;
  const double inv = (M.mode == MODE_GRANK) ? GRANK_HUB_INV : MC_HUB_INV;

  for (int i = tid; i < PAR_CAP; i += PAR_THREADS) { s_keys[i] = KEY_EMPTY; s_acc[i] = make_uint2(0u, 0u); }
  if (tid == 0) { S->count = 0; S->spilled = 0; S->table = -1; }
  __syncthreads();

  int read_slot[2];
  read_slot[0] = st->slot[0];
  read_slot[1] = st->slot[1];

  unsigned long long s_merged = 0, s_edges = 0, s_cands = 0, s_truncs = 0, s_ties = 0, s_bytes = 0, s_nodes = 0;

  // shared-table accumulate; returns false when the key is absent and the table is closed (caller spills)
  auto smem_add = [&](int k, unsigned long long x) -> bool {
    unsigned int h = hash_key(k) & (PAR_CAP - 1);
    volatile int* keys = s_keys;
    for (;;) {
      const int cur = keys[h];
      if (cur == k) break;
      if (cur == KEY_EMPTY) {
        if (*reinterpret_cast<volatile int*>(&S->count) >= PAR_LIMIT) return false;
        const int old = atomicCAS(&s_keys[h], KEY_EMPTY, k);
        if (old == KEY_EMPTY) { const int pos = atomicAdd(&S->count, 1); s_list[pos] = (unsigned short)h; break; }
        if (old == k) break;
      }
      h = (h + 1) & (PAR_CAP - 1);
    }
    const unsigned int xlo = (unsigned int)x, xhi = (unsigned int)(x >> 32);
    const unsigned int old = atomicAdd(&s_acc[h].x, xlo);
    const unsigned int carry = (old + xlo) < old ? 1u : 0u;
    if (xhi + carry) atomicAdd(&s_acc[h].y, xhi + carry);
    return true;
  };

  for (;;) {
    if (tid == 0) S->item = atomicAdd(&st->work[P.work_idx], 1u);
    __syncthreads();
    const unsigned int item = S->item;
    if (item >= (unsigned)P.n_items) break;
    const int p = P.item_pos[item];
    const long long cb = P.item_begin[item];
    const int clen = P.item_len[item];
    const long long rb = M.g.row_off[p], re = M.g.row_off[p + 1];
    const long long deg = re - rb;
    const int nchunks = (int)((deg + P.chunk - 1) / P.chunk);
    const int self_id = M.g.label[p];
    const double f = M.damping / (double)(unsigned long long)deg;
    const double mult = f;  // the hub path pre-scales by f = d/outdeg in both modes (oracle: llrint((x * f) * scale))
    const double self0 = (M.mode == MODE_GRANK) ? M.self_grank : 1.0;
    const int write_slot = init_mode ? st->slot[M.colour] : (st->slot[M.colour] ^ 1);

    // lazily bound global table of this node (acquired by the first CTA that needs it)
    GSlot* gslots = nullptr;
    unsigned int* glist = nullptr;
    unsigned int* gcount = nullptr;
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
        // per-node capacity: never overflows (worst case deg*Lp+1 distinct keys, or the whole id space)
        unsigned long long bound = (unsigned long long)deg * (unsigned long long)(init_mode ? 1 : Lp) + 2ull;
        unsigned long long want = 2ull * bound;
        unsigned int cap = P.capmax;
        if (want < (unsigned long long)P.capmax) { cap = 1024u; while ((unsigned long long)cap < want) cap <<= 1; }
        S->gmask = cap - 1u;
        S->gidentity = cap > (unsigned)P.n_ids ? 1 : 0;
      }
      __syncthreads();
      unsigned char* base = P.pool + (size_t)S->table * P.tbl_bytes;
      gslots = reinterpret_cast<GSlot*>(base);
      glist = reinterpret_cast<unsigned int*>(base + (size_t)P.capmax * sizeof(GSlot));
      gcount = &P.tbl_count[S->table];
    };
    if (nchunks > 1) bind_table();

    if (tid == 0 && cb == rb) {  // first chunk owns the self term (grank.h:101 / mccompletepathv2.h:226)
      const unsigned long long x = init_mode ? 0ull : (unsigned long long)__double2ll_rn(self0 * scale);
      const unsigned int h = hash_key(self_id) & (PAR_CAP - 1);
      s_keys[h] = self_id;
      s_acc[h] = make_uint2((unsigned int)x, (unsigned int)(x >> 32));
      s_list[0] = (unsigned short)h;
      S->count = 1;
    }
    __syncthreads();

    // ---- accumulate: warp w takes runs of 32 successors of the chunk; a contribution whose key is absent
    // from the (closed) shared table goes to the global table when one is bound, else the pass is repeated ----
    unsigned long long merged = 0;
    auto contribute = [&](int k, unsigned long long xf, bool spill_ok) {
      if (!smem_add(k, xf)) {
        if (spill_ok) gtable_add(gslots, glist, gcount, S->gmask, S->gidentity, k, xf);
        else S->spilled = 1;
      }
    };
    auto accumulate = [&](bool spill_ok) {
      merged = 0;
      if (init_mode) {
        // grank.h:79-80: every occurrence of a successor adds `factor`; here: multiplicity += 1
        for (int j = tid; j < clen; j += PAR_THREADS) {
          const uint32_t c = M.g.col[cb + j];
          const int k = (c & COL_SINK) ? (int)(c & ~COL_SINK) : M.g.label[c & COL_POS_MASK];
          contribute(k, 1ull, spill_ok);
          merged++;
        }
        return;
      }
      for (int j = w; j < clen; j += NW) {  // warp w merges successors w, w+NW, ...
        const uint32_t c = M.g.col[cb + j];
        if (c & COL_SINK) {
          if (lane == 0) {
            const double x = (M.mode == MODE_GRANK) ? M.self_grank : 1.0;
            contribute((int)(c & ~COL_SINK), (unsigned long long)__double2ll_rn((x * mult) * scale), spill_ok);
            merged++;
          }
        } else {
          const unsigned int sp = c & COL_POS_MASK;
          const int sc = (int)((c >> COL_COLOUR_SHIFT) & 1u);
          const unsigned char* slot = M.buf[read_slot[sc]] + (size_t)sp * slot_bytes(Lp);
          for (int g = lane; g < groups; g += 32) {
            BasketFrag fr;
            load_frag(slot, Lp, g, &fr);
            const int ids[4] = {fr.id.x, fr.id.y, fr.id.z, fr.id.w};
            const double xs[4] = {fr.sa.x, fr.sa.y, fr.sb.x, fr.sb.y};
#pragma unroll
            for (int e = 0; e < 4; e++) {
              if (ids[e] >= 0) {
                contribute(ids[e], (unsigned long long)__double2ll_rn((xs[e] * mult) * scale), spill_ok);
                merged++;
              }
            }
          }
        }
      }
    };
    // a node expected to outgrow the shared table binds its global table up front
    if (S->table < 0 && M.ncand[p] > PAR_LIMIT - PAR_THREADS) bind_table();
    accumulate(S->table >= 0);
    __syncthreads();
    if (S->spilled) {
      // mispredicted: drop the partial sums, bind a table and run the chunk again with spilling enabled
      const int dirty = S->count;
      __syncthreads();
      for (int i = tid; i < dirty; i += PAR_THREADS) {
        const int s = s_list[i];
        s_keys[s] = KEY_EMPTY;
        s_acc[s] = make_uint2(0u, 0u);
      }
      __syncthreads();
      if (tid == 0) { S->count = 0; S->spilled = 0; }
      bind_table();
      if (tid == 0 && cb == rb) {
        const unsigned long long x = init_mode ? 0ull : (unsigned long long)__double2ll_rn(self0 * scale);
        const unsigned int h = hash_key(self_id) & (PAR_CAP - 1);
        s_keys[h] = self_id;
        s_acc[h] = make_uint2((unsigned int)x, (unsigned int)(x >> 32));
        s_list[0] = (unsigned short)h;
        S->count = 1;
      }
      __syncthreads();
      accumulate(true);
      __syncthreads();
    }

    const int ns = S->count;  // occupied shared slots
    const bool use_global = S->table >= 0;
    int n = ns;
    int kept = 0, old_cnt = 0;
    bool finalize = !use_global;  // single chunk, nothing spilled: finish from shared memory
    if (use_global) {
      // flush the shared table into the node's global table
      for (int i = tid; i < ns; i += PAR_THREADS) {
        const int s = s_list[i];
        const unsigned long long a = ((unsigned long long)s_acc[s].y << 32) | s_acc[s].x;
        gtable_add(gslots, glist, gcount, S->gmask, S->gidentity, s_keys[s], a);
      }
      __threadfence();
      __syncthreads();
      if (tid == 0) {
        const unsigned int done = atomicAdd(&P.node_done[p], 1u) + 1u;
        S->is_last = done == (unsigned)nchunks;
        if (S->is_last) __threadfence();
      }
      __syncthreads();
      finalize = S->is_last != 0;
      if (finalize) n = (int)*reinterpret_cast<volatile unsigned int*>(gcount);
    }

    if (finalize) {
      const double base_self = init_mode ? M.self_grank : 0.0;
      // candidate accessors
      auto cand_id = [&](int i) -> int { return use_global ? gslots[glist[i]].key : s_keys[s_list[i]]; };
      auto cand_val = [&](int i) -> double {
        unsigned long long a;
        int id;
        if (use_global) { const GSlot g = gslots[glist[i]]; a = g.acc; id = g.key; }
        else { const int s = s_list[i]; a = ((unsigned long long)s_acc[s].y << 32) | s_acc[s].x; id = s_keys[s]; }
        return par_score(a, init_mode, inv, mult, (init_mode && id == self_id) ? base_self : 0.0);
      };
      Threshold th;
      th.bits = 0ull;
      th.id_max = 0x7fffffff;
      kept = n;
      if (n > L) {
        kept = L;
        bool tie;
        int krem;
        auto keyfn = [&](int i) { return (unsigned long long)__double_as_longlong(cand_val(i)); };
        auto all = [&](int) { return true; };
        th.bits = block_radix_select(n, L, keyfn, all, S, &tie, &krem);
        if (tie) {
          const unsigned long long tb = th.bits;
          auto idkey = [&](int i) { return (unsigned long long)(0x7fffffff - cand_id(i)); };
          auto tied = [&](int i) { return (unsigned long long)__double_as_longlong(cand_val(i)) == tb; };
          bool tie2;
          int krem2;
          const unsigned long long tid_key = block_radix_select(n, krem, idkey, tied, S, &tie2, &krem2);
          th.id_max = 0x7fffffff - (int)tid_key;
          s_ties += (tid == 0);
        }
        s_truncs += (tid == 0);
      }
      // ---- write B'_v ----
      unsigned char* out = M.buf[write_slot] + (size_t)p * slot_bytes(Lp);
      int* out_ids = reinterpret_cast<int*>(out);
      double* out_sc = reinterpret_cast<double*>(out + (size_t)Lp * 4);
      if (tid == 0) S->out_pos = 0;
      __syncthreads();
      long long dsum = 0;
      for (int i = tid; i < n; i += PAR_THREADS) {
        const int id = cand_id(i);
        const double v = cand_val(i);
        if (is_selected(th, (unsigned long long)__double_as_longlong(v), id)) {
          const int pos = atomicAdd(&S->out_pos, 1);
          out_ids[pos] = id;
          out_sc[score_index(pos, Lp)] = v;  // hub path: no post-scale (already multiplied by f)
          dsum += fix_norm(v);
        }
      }
      for (int i = kept + tid; i < Lp; i += PAR_THREADS) out_ids[i] = KEY_EMPTY;
      // ---- norm1 against the old basket (pprInternal.h:147-165) ----
      if (M.do_norm) {
        const unsigned char* old = M.buf[write_slot ^ 1] + (size_t)p * slot_bytes(Lp);
        for (int g = tid; g < groups; g += PAR_THREADS) {
          BasketFrag fr;
          load_frag(old, Lp, g, &fr);
          const int ids[4] = {fr.id.x, fr.id.y, fr.id.z, fr.id.w};
          const double xs[4] = {fr.sa.x, fr.sa.y, fr.sb.x, fr.sb.y};
#pragma unroll
          for (int e = 0; e < 4; e++) {
            if (ids[e] >= 0) {
              old_cnt++;
              bool found = false;
              double nv = 0.0;
              if (use_global) {
                const int s = gtable_find(gslots, S->gmask, S->gidentity, ids[e]);
                if (s >= 0) { found = true; nv = par_score(gslots[s].acc, false, inv, mult, 0.0); }
              } else {
                for (unsigned int h = hash_key(ids[e]) & (PAR_CAP - 1);; h = (h + 1) & (PAR_CAP - 1)) {
                  const int cur = s_keys[h];
                  if (cur == ids[e]) { found = true; nv = par_score(((unsigned long long)s_acc[h].y << 32) | s_acc[h].x, false, inv, mult, 0.0); break; }
                  if (cur == KEY_EMPTY) break;
                }
              }
              const bool in_new = found && is_selected(th, (unsigned long long)__double_as_longlong(nv), ids[e]);
              if (in_new) dsum += fix_norm(fabs(nv - xs[e])) - fix_norm(nv);
              else dsum += fix_norm(xs[e]);
            }
          }
        }
        dsum = block_reduce_sum_ll(dsum, S->red_a);
        old_cnt = (int)block_reduce_sum_ll(old_cnt, S->red_a);
        if (tid == 0 && dsum > 0) atomicMax(&st->cur_max, dsum);
      }
      __syncthreads();
      if (use_global) {
        // leave the pool table clean and hand it back
        for (int i = tid; i < n; i += PAR_THREADS) {
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
      }
      if (tid == 0) {
        M.ncand[p] = n;
        s_cands += (unsigned long long)n;
        s_nodes += 1;
        s_bytes += 12ull * (unsigned long long)old_cnt + 12ull * (unsigned long long)kept + 4ull + 16ull;
      }
    }
    // ---- reset the shared table through its list ----
    for (int i = tid; i < ns; i += PAR_THREADS) {
      const int s = s_list[i];
      s_keys[s] = KEY_EMPTY;
      s_acc[s] = make_uint2(0u, 0u);
    }
    merged = (unsigned long long)block_reduce_sum_ll((long long)merged, S->red_a);
    if (tid == 0) {
      s_merged += merged;
      s_edges += (unsigned long long)clen;
      s_bytes += 12ull * merged + 4ull * (unsigned long long)clen;
      S->count = 0;
      S->spilled = 0;
      S->table = -1;
    }
    __syncthreads();
  }
  if (tid == 0) {
    if (s_nodes) atomicAdd(&st->node_iters, s_nodes);
    if (s_edges) atomicAdd(&st->edge_reads, s_edges);
    if (s_merged) atomicAdd(&st->merged, s_merged);
    if (s_cands) atomicAdd(&st->cands, s_cands);
    if (s_truncs) atomicAdd(&st->truncs, s_truncs);
    if (s_ties) atomicAdd(&st->ties, s_ties);
    if (s_bytes) atomicAdd(&st->abytes, s_bytes);
  }
}

}  // namespace pprb200
