// mc_walk.cuh -- the Monte-Carlo "complete path" walk phase (M0) of MCCompletePathV2.
//
// Reproduces include/mccompletepathv2.h:115-165 (walkNode) for every non-sink source with the north-star changes
// (SURVEY.md 8a-12/13): the successor of each hop is drawn uniformly from a counter-based generator instead of the
// reference's shared rotating index, every visit is counted exactly (no first-come cap) and the top-L is taken
// afterwards. One thread per walk: Philox4x32-10 keyed (source dense id, walk index), counter = (hop/2, 0, seed);
// hop parity picks words (0,1) or (2,3): word A chooses the successor (mulhi(A, outdeg)), word B is the teleport
// coin (continue iff B < floor(d * 2^32)). The first edge is always taken, W = (size_t)(R * d) walks per source,
// counts are divided by R (:132,:142-160). oracle/ppr_oracle.c:mc_walk_source is the CPU twin, bit for bit:
// visit counts are integers, so neither the thread that ran a walk nor the GPU that ran a source matters.
//
// A CTA owns one source at a time: visit counts live in a shared-memory open-addressing table keyed by the CSR
// column word of the visited node (position | colour, or sink | label -- unique per node); threads fetch walk
// indices from a shared counter so that long walks do not idle the rest of the warp for long. A source whose set of
// visited nodes outgrows the table is queued for the fallback instantiation (table in a global workspace).
#pragma once
#include "merge_par.cuh"

namespace pprb200 {

constexpr unsigned int MC_MAX_STEPS = 4096u;  // hard cap per walk (oracle: same); P[len > 4096] = d^4096
constexpr uint32_t WALK_EMPTY = 0xffffffffu;  // never a valid column word (sink | label 0x7fffffff is out of range)

struct WalkParams {
  GraphDev g;
  unsigned char* buf[2];
  RunState* st;
  int M;                      // sources = non-sink storage positions 0..M-1 (this rank: [src_begin, src_end))
  int n_ids;                  // number of nodes
  int src_begin, src_end;
  const unsigned char* colour;  // [n] colour by dense id (a node's own column word carries it)
  int Lp, L;
  unsigned int R;
  unsigned long long W;
  uint32_t thresh;
  unsigned long long seed;
  unsigned int tcap;          // table slots (power of two)
  unsigned int limit;         // max distinct visited nodes admitted by this launch
  int work_idx;
  const unsigned int* queue_in;
  int queue_in_idx;           // -1: the range [src_begin, src_end)
  unsigned int* queue_out;    // nullable (fallback launch never overflows)
  int queue_out_idx;
  unsigned long long* ws;     // fallback: gridDim.x x (table of tcap 64-bit slots + first-touch list of tcap 32-bit slot indices)
  PeerDev peers;              // multi-GPU: source index i of this rank maps to position src_begin + i*world + rank
  const unsigned long long* rowdeg;  // [M] packed row word: first column offset << ROWDEG_SHIFT | out-degree (one 8-byte gather per hop)
};

constexpr int ROWDEG_SHIFT = 26;  // out-degree < 2^26, column offsets < 2^38 (validated when the session is created)

struct WalkSlot {
  uint32_t key;
  uint32_t count;
};

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ uint32_t hash_word(uint32_t w) {
  uint32_t x = w * 0x9E3779B1u;
  return x ^ (x >> 15);
}

constexpr int WALK_CH = 1024;  // bins of the visit-count histogram that finds the top-L cut in one pass

struct WalkShared {
  ParShared P;               // scratch of block_radix_select
  unsigned int next_walk;
  unsigned int ndistinct;
  int overflow;
  unsigned int item;
  int cut_count, cut_above, cut_ties;
  unsigned int chist[WALK_CH];
};

template <bool GLOBAL, int THREADS>
__global__ void __launch_bounds__(THREADS, (THREADS <= 128 ? 6 : 3)) mc_walk_kernel(WalkParams P) {
  extern __shared__ __align__(16) unsigned char smem[];
  RunState* st = P.st;
  WalkShared* S = reinterpret_cast<WalkShared*>(smem);
  // The occupied slots of the table are tracked in a first-touch list, so that selecting, writing and clearing cost
  // O(distinct visited nodes), not O(capacity). Shared-memory tables: 16-bit list behind the table; fallback tables
  // (global workspace, sized for the worst case): 32-bit list behind the table. The capacity need not be a power of two.
  unsigned char* gbase = GLOBAL ? reinterpret_cast<unsigned char*>(P.ws) + (size_t)blockIdx.x * P.tcap * (sizeof(WalkSlot) + sizeof(unsigned int)) : nullptr;
  WalkSlot* tbl = GLOBAL ? reinterpret_cast<WalkSlot*>(gbase)
                         : reinterpret_cast<WalkSlot*>(smem + ((sizeof(WalkShared) + 15) & ~(size_t)15));
  unsigned int* glist = GLOBAL ? reinterpret_cast<unsigned int*>(gbase + (size_t)P.tcap * sizeof(WalkSlot)) : nullptr;
  unsigned short* slist = GLOBAL ? nullptr : reinterpret_cast<unsigned short*>(reinterpret_cast<unsigned char*>(tbl) + (size_t)P.tcap * sizeof(WalkSlot));
  auto list_set = [&](unsigned int i, unsigned int h) { if (GLOBAL) glist[i] = h; else slist[i] = (unsigned short)h; };
  auto list_get = [&](unsigned int i) -> unsigned int { return GLOBAL ? glist[i] : (unsigned int)slist[i]; };
  for (unsigned int i = threadIdx.x; i < P.tcap; i += THREADS) { tbl[i].key = WALK_EMPTY; tbl[i].count = 0u; }
  __syncthreads();
  const int tid = threadIdx.x;
  const unsigned int cap = P.tcap;
  auto slot0 = [&](uint32_t word) -> unsigned int { return __umulhi(hash_word(word), cap); };
  const int Lp = P.Lp, L = P.L;
  const double inv_r_den = (double)P.R;

  unsigned int total;
  if (P.queue_in_idx >= 0) total = st->qcount[P.queue_in_idx];
  else {
    const int len = P.src_end - P.src_begin, w = P.peers.world > 1 ? P.peers.world : 1, r = P.peers.world > 1 ? P.peers.rank : 0;
    total = len > r ? (unsigned int)((len - r + w - 1) / w) : 0u;
  }
  const int stride = P.peers.world > 1 ? P.peers.world : 1, offset = P.peers.world > 1 ? P.peers.rank : 0;
  const int write_slot = st->slot[0];
  unsigned long long tot_steps = 0, tot_walks = 0, tot_bytes = 0, tot_truncs = 0, tot_ties = 0, tot_requeue = 0;

  for (;;) {
    if (tid == 0) S->item = atomicAdd(&st->work[P.work_idx], 1u);
    __syncthreads();
    const unsigned int item = S->item;
    if (item >= total) break;
    const int p = (P.queue_in_idx >= 0) ? (int)P.queue_in[item] : P.src_begin + (int)item * stride + offset;
    const int self_label = P.g.label[p];
    const uint32_t src_dense = (uint32_t)P.g.dense_of[self_label];
    const uint32_t self_word = (uint32_t)p | ((uint32_t)P.colour[src_dense] << COL_COLOUR_SHIFT);

    if (tid == 0) {  // res[src] = R (mccompletepathv2.h:124); the table is empty here
      S->next_walk = 0u; S->ndistinct = 1u; S->overflow = 0;
      const unsigned int h = slot0(self_word);
      tbl[h].key = self_word;
      tbl[h].count = P.R;
      list_set(0u, h);
    }
    __syncthreads();

    // ---- walks ----
    // Every lane carries one walk; the warp advances all of them in lock step, two hops per iteration (one Philox block
    // yields the successor choice and the teleport coin of an even and an odd hop), and lanes whose walk has ended are
    // refilled at the top of the next iteration with one shared-memory atomic per warp. The round-1 loop let every thread
    // run its own walk-fetch / Philox / probe sequence: 12.6 of 32 lanes active per issued instruction (profiles/r1).
    unsigned long long steps = 0;
    {
      const int lane = tid & 31;
      bool active = false;
      uint32_t pos = (uint32_t)p, w = 0u;
      unsigned int step = 0u;
      for (;;) {
        const unsigned need = __ballot_sync(FULL, !active);
        if (need) {
          unsigned int base = 0u;
          const int leader = __ffs(need) - 1;
          if (lane == leader) base = atomicAdd(&S->next_walk, (unsigned int)__popc(need));
          base = __shfl_sync(FULL, base, leader);
          if (!active) {
            w = base + (unsigned int)__popc(need & ((1u << lane) - 1u));
            if ((unsigned long long)w < P.W && !*reinterpret_cast<volatile int*>(&S->overflow)) { active = true; pos = (uint32_t)p; step = 0u; }
          }
        }
        if (!__any_sync(FULL, active)) break;
        uint32_t rnd[4];
        philox4x32_10(step >> 1, 0u, (uint32_t)P.seed, (uint32_t)(P.seed >> 32), src_dense, w, rnd);  // (step is even here)
#pragma unroll
        for (int half = 0; half < 2; half++) {
          if (active) {
            const unsigned long long rw = __ldg(P.rowdeg + pos);
            const unsigned long long deg = rw & ((1ull << ROWDEG_SHIFT) - 1ull);
            const uint32_t a = rnd[2 * half], c = rnd[2 * half + 1];
            const uint32_t word = __ldg(P.g.col + (rw >> ROWDEG_SHIFT) + (((unsigned long long)a * deg) >> 32));  // :149, random successor
            // count the visit (:152-153 without the cap)
            unsigned int h = slot0(word);
            for (;;) {
              const uint32_t cur = *reinterpret_cast<volatile uint32_t*>(&tbl[h].key);
              if (cur == word) break;
              if (cur == WALK_EMPTY) {
                const uint32_t old = atomicCAS(&tbl[h].key, WALK_EMPTY, word);
                if (old == WALK_EMPTY) {
                  const unsigned int np = atomicAdd(&S->ndistinct, 1u);
                  list_set(np, h);
                  if (np + 1u > P.limit) S->overflow = 1;
                  break;
                }
                if (old == word) break;
              }
              h = h + 1u == cap ? 0u : h + 1u;
            }
            atomicAdd(&tbl[h].count, 1u);
            steps++;
            step++;
            // :144-145 the walk stops on reaching a sink; :155 the teleport coin; hard cap; a full table ends the source
            if ((word & COL_SINK) || !(c < P.thresh) || step >= MC_MAX_STEPS || *reinterpret_cast<volatile int*>(&S->overflow)) active = false;
            else pos = word & COL_POS_MASK;
          }
        }
      }
    }
    __syncthreads();
    if (S->overflow) {
      const unsigned int nd = S->ndistinct;
      for (unsigned int i = tid; i < nd; i += THREADS) { const unsigned int h = list_get(i); tbl[h].key = WALK_EMPTY; tbl[h].count = 0u; }
      if (tid == 0) {
        const unsigned int q = atomicAdd(&st->qcount[P.queue_out_idx], 1u);
        P.queue_out[q] = (unsigned int)p;
        tot_requeue++;
      }
      __syncthreads();
      continue;
    }

    // ---- top-L on (count desc, dense id asc), scores = count / R (:159-160) ----
    // Visit counts are small integers: one pass over the first-touch list fills a histogram of min(count, WALK_CH - 1),
    // a warp finds the cut. (A cut inside the last bin -- at least L nodes visited >= WALK_CH - 1 times -- takes the radix
    // select.) Ties on the cut count are broken by dense id with a radix select over the tied entries only.
    const int n = (int)S->ndistinct;
    const int nscan = n;
    auto slot_at = [&](int i) -> unsigned int { return list_get((unsigned int)i); };
    auto word_label = [&](uint32_t wd) -> int { return (wd & COL_SINK) ? (int)(wd & ~COL_SINK) : P.g.label[wd & COL_POS_MASK]; };
    Threshold th;
    th.bits = 0ull;
    th.id_max = 0x7fffffff;
    int kept = n;
    if (n > L) {
      kept = L;
      bool tie = false;
      int krem = 0, ntied = 0;
      for (int i = tid; i < WALK_CH; i += THREADS) S->chist[i] = 0u;
      __syncthreads();
      {
        // most nodes are visited once or a few times: those bins are counted in registers (thousands of shared-memory
        // atomics on one address would serialise), the rest with atomics
        int small[4] = {0, 0, 0, 0};
        for (int i = tid; i < n; i += THREADS) {
          const unsigned int c = tbl[slot_at(i)].count;
          if (c >= 1u && c <= 4u) {
#pragma unroll
            for (int j = 0; j < 4; j++) small[j] += (c == (unsigned)(j + 1));
          } else {
            atomicAdd(&S->chist[c < (unsigned)(WALK_CH - 1) ? c : (unsigned)(WALK_CH - 1)], 1u);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int t = warp_sum_int(small[j]);
          if ((tid & 31) == 0 && t) atomicAdd(&S->chist[j + 1], (unsigned int)t);
        }
      }
      __syncthreads();
      if (tid < 32) {
        constexpr int PER = WALK_CH / 32;
        unsigned int local = 0;
        for (int j = 0; j < PER; j++) local += S->chist[tid * PER + j];
        unsigned int incl = local;  // entries in the bins of this lane and of all higher lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned int t = __shfl_down_sync(FULL, incl, o);
          if (tid + o < 32) incl += t;
        }
        const unsigned int above = incl - local;
        if (above < (unsigned)L && incl >= (unsigned)L) {
          unsigned int run = above;
          for (int j = PER - 1; j >= 0; j--) {
            const unsigned int c = S->chist[tid * PER + j];
            if (run < (unsigned)L && run + c >= (unsigned)L) { S->cut_count = tid * PER + j; S->cut_above = (int)run; S->cut_ties = (int)c; }
            run += c;
          }
        }
      }
      __syncthreads();
      if (S->cut_count < WALK_CH - 1) {
        th.bits = (unsigned long long)S->cut_count;
        krem = L - S->cut_above;
        ntied = S->cut_ties;
        tie = ntied > krem;
      } else {
        auto occupied = [&](int) { return true; };
        auto keyfn = [&](int i) { return (unsigned long long)tbl[slot_at(i)].count; };
        th.bits = block_radix_select(nscan, L, keyfn, occupied, &S->P, &tie, &krem, &ntied);
      }
      __syncthreads();
      if (tie) {
        const unsigned long long tb = th.bits;
        auto idkey = [&](int i) { return (unsigned long long)(0x7fffffff - P.g.dense_of[word_label(tbl[slot_at(i)].key)]); };
        auto tied = [&](int i) { return (unsigned long long)tbl[slot_at(i)].count == tb; };
        bool tie2;
        int krem2;
        const unsigned long long tid_key = block_radix_select(nscan, krem, idkey, tied, &S->P, &tie2, &krem2);
        th.id_max = 0x7fffffff - (int)tid_key;
        tot_ties += (tid == 0);
      }
      tot_truncs += (tid == 0);
    }
    unsigned char* out = P.buf[write_slot] + (size_t)p * slot_bytes(Lp);
    int* out_ids = reinterpret_cast<int*>(out);
    double* out_sc = reinterpret_cast<double*>(out + (size_t)Lp * 4);
    if (tid == 0) S->P.out_pos = 0;
    __syncthreads();
    for (int i = tid; i < nscan; i += THREADS) {
      const WalkSlot ts = tbl[slot_at(i)];
      const uint32_t wd = ts.key;
      const unsigned long long cnt = ts.count;
      bool sel = cnt > th.bits;
      int label = -1;
      if (!sel && cnt == th.bits) {
        label = word_label(wd);
        sel = th.id_max == 0x7fffffff || P.g.dense_of[label] <= th.id_max;
      }
      if (sel) {
        if (label < 0) label = word_label(wd);
        const int pos = atomicAdd(&S->P.out_pos, 1);
        out_ids[pos] = label;
        out_sc[score_index(pos, Lp)] = __ddiv_rn((double)cnt, inv_r_den);
      }
    }
    for (int i = kept + tid; i < Lp; i += THREADS) out_ids[i] = KEY_EMPTY;
    __syncthreads();
    // leave the table empty for the next source
    for (int i = tid; i < n; i += THREADS) { const unsigned int h = slot_at(i); tbl[h].key = WALK_EMPTY; tbl[h].count = 0u; }
    // (to every peer, whatever the need masks say: walk sources are dealt out round-robin, not to the ranks that own the
    // positions in the combine rounds -- the owner reads this basket as its old one, and with rounds = 0 it is the result)
    publish_slot(P.peers, write_slot, (size_t)p * slot_bytes(Lp), slot_bytes(Lp), tid, THREADS, -1);
    steps = (unsigned long long)block_reduce_sum_ll((long long)steps, S->P.red_a);
    if (tid == 0) {
      tot_steps += steps;
      tot_walks += P.W;
      tot_bytes += 12ull * steps + 12ull * (unsigned long long)kept + 4ull;
    }
    __syncthreads();
  }
  if (tid == 0) {
    if (tot_steps) atomicAdd(&st->walk_steps, tot_steps);
    if (tot_walks) atomicAdd(&st->walks, tot_walks);
    if (tot_bytes) atomicAdd(&st->walk_bytes, tot_bytes);
    if (tot_truncs) atomicAdd(&st->truncs, tot_truncs);
    if (tot_ties) atomicAdd(&st->ties, tot_ties);
    if (tot_requeue) atomicAdd(&st->requeues, tot_requeue);
  }
}

}  // namespace pprb200
