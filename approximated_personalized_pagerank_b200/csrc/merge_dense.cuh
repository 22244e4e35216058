// merge_dense.cuh -- the fast path of the order-free merge step: one CTA per node, everything in shared memory.
//
// Same operator and the same sums as merge_par.cuh (grank.h:96-126, mccompletepathv2.h:211-250 with every product
// rounded once to 2^-62 / 2^-59 fixed point, oracle/ppr_oracle.c hub mode), restructured around what the round-1
// profile showed (profiles/r2/): the accumulate loops were bound by DIVERGENCE, not by memory or atomics -- with
// 10-20 % tail labels per basket nearly every warp step executed the dense path, the sketch path and the
// old-basket path one after the other -- and by ~37 % per-node overhead (two selects, three table sweeps).
//
//   * Pass 1 is ONE branch-free instruction stream for every entry -- cvt, address select, ATOMS.ADD low word, carry
//     into the high word -- over three kinds of accumulators: rank label k < H owns a 64-bit fixed-point word; a
//     tail label of the node's OLD basket ("pre" label: the best predictor of the new basket) owns the word of its
//     home slot in the tail hash table; every other label shares a 32-bit sketch bucket that holds an UPPER bound of
//     the sum (each product rounded up to 2^-30 / 2^-27): one atomic, no carry, and twice the buckets per byte.
//   * Lanes walk the node's successor baskets as one stream of 16-byte quads (4 entries): 32 consecutive quads per
//     warp step whatever the basket length, so all 32 lanes work (L = 100 gives 25 quads per basket: a warp per
//     basket would idle 7 lanes) and the next step's three LDG.128 per lane are in flight while this one is added.
//   * The cut is bounded from below by tau = the old basket's smallest score (or 3/4, 1/2, 1/16 of it), validated by
//     counting: if at least L exact candidates (dense + pre labels) reach tau, no label below tau can be kept. A sketch
//     bucket below tau holds no label that can be kept (strictly, ties included), so pass 2 -- needed only when some
//     bucket survives -- re-reads the label words (L2-hot) and accumulates exactly the tail labels of the surviving
//     buckets in the hash table. One radix select per node.
//   * Anything that does not fit (more candidates / surviving tail labels than the shared arrays hold, a product
//     that rounds to zero -- such a label is a candidate with score 0 and leaves no trace in a sum --, a sketch
//     bucket that wraps, nodes split into several chunks) is handed to merge_par_kernel through a device queue:
//     same sums, slower path.
#pragma once
#include "merge_par.cuh"

namespace pprb200 {

// HUB TEAMS. A hub with tens of thousands of successors is the critical path of its iteration when one CTA owns it (the
// largest R-MAT-22 hub: 16 M entries, two passes, ~15 ms -- more than the whole iteration takes on 8 GPUs). Its successor
// list is cut into chunks that consecutive work items hand to different CTAs (a team). Every member runs pass 1 on its
// chunk into its own shared-memory tables and adds what it found into the team's staging area in global memory (L2): the
// tables are position-compatible -- dense words by label, sketch buckets by hash, pre words by home slot (pre labels are
// placed deterministically for teams) -- so the staging area is simply their sum. The member that flushes last loads the
// totals and carries on exactly as the single owner would; if pass 2 is needed it publishes the alive bitmap, every member
// runs pass 2 on its chunk and appends its exact (label, sum) pairs to the staging area, and the last member merges them.
// Members wait for each other in flight (spinning on the team header), which cannot deadlock: work items are handed out in
// order, a team's chunks are consecutive, and a team has at most TEAM_MAX_CHUNKS (< CTAs in the grid) of them.
constexpr int TEAM_MAX_CHUNKS = 48;

struct TeamInfo {
  int regular_item;  // global index of the hub's undivided work item (what is handed to merge_par_kernel if the team gives up)
  int nchunks;
};

struct TeamHeader {
  unsigned int done1;       // members that flushed pass 1
  unsigned int done2;       // members that appended their pass-2 sums
  unsigned int decision;    // 0 pending, 1 no pass 2, 2 pass 2 (alive bitmap published), 3 team gives up
  unsigned int tail_count;  // (label, sum) pairs appended by pass 2
  unsigned int bad;         // bit 0: an untrusted contribution / wrapped bucket, bit 1: a member's tail table overflowed
  unsigned int edges;       // successors read by the members (statistics)
  unsigned long long merged;  // basket entries merged by the members (statistics)
};

struct TeamTailEntry {
  int key;
  int pad;
  unsigned long long acc;
};

template <int H, int R, int TCAP, int THREADS>
constexpr size_t team_stage_bytes() {
  return (size_t)H * 8 + (size_t)TCAP * 8 + (size_t)R * 4 + (size_t)R / 8 + (size_t)TEAM_MAX_CHUNKS * (size_t)(TCAP * 13 / 16 - THREADS) * sizeof(TeamTailEntry);
}

struct DenseParams {
  MergeParams M;
  const int* item_pos;          // work items of this launch: node position,
  const long long* item_begin;  //   first successor (absolute offset into col),
  const int* item_len;          //   number of successors
  int n_items;
  int item_base;                // global index of item 0 (queue entries are global item indices)
  int chunk;                    // nodes above this out-degree are split into several items: not handled here
  int work_idx;
  unsigned int* fb_queue;       // items handed to merge_par_kernel
  int fb_idx;                   // its length: st->qcount[fb_idx]
  int min_old;                  // nodes whose old basket holds fewer entries than this are handed over unread (L: only full baskets stay)
  int tail_limit;               // distinct tail labels the table admits (<= the instantiation's TLIMIT; tests lower it to force the rounds)
  unsigned long long* prof;     // optional [gridDim.x * 8] phase cycle counters (PPRB200_PROF=1)
  // hub teams (below): work indices [0, n_team_items) are chunks of team hubs, the rest the items above
  const int* team_item_pos;
  const long long* team_item_begin;
  const int* team_item_len;
  const int* team_item_team;
  int n_team_items;
  const TeamInfo* teams;
  TeamHeader* team_hdr;
  unsigned char* stage;         // team t's staging area: stage + t * stage_bytes
  size_t stage_bytes;
};

struct DenseShared {
  ParShared P;   // scratch of the block reductions / radix select
  int ncol;      // non-sink column words staged for the current tile
  int bail;      // hand the node to the general kernel
  int lvl[4];    // exact candidates >= theta0, 3/4 theta0, 1/2 theta0, 1/16 theta0
  int team_last;      // this CTA flushed last: it finishes the team's hub
  int team_decision;  // what the waiting members read from the team header
  unsigned int team_base;  // where this member's pass-2 pairs go in the staging area
  unsigned int team_edges;          // the whole team's statistics (finishing member)
  unsigned long long team_merged;
  unsigned long long cut_bits;  // rank-count select: the L-th largest score,
  int cut_gt, cut_eq;           //   candidates above it / equal to it
};

template <int H, int R, int TCAP, int CMAX, int COLCAP>
constexpr size_t dense_smem_bytes() {
  return (size_t)H * 8 + (size_t)R * 4 + (size_t)CMAX * 12 + (size_t)TCAP * 14 + (size_t)COLCAP * 4 + (size_t)R / 8 + sizeof(DenseShared) + 16;
}

// TEAMS: the instantiation that serves hub teams' chunk items (launched with those alone); without it the team code is
// compiled out -- it costs the 64-register instantiations ~6 % in spills (profiles/r2/sweeps.txt)
template <int H, int R, int TCAP, int CMAX, int COLCAP, int THREADS, int MINB, bool TEAMS = false>
__global__ void __launch_bounds__(THREADS, MINB) merge_dense_kernel(DenseParams P) {
  constexpr int NW = THREADS / 32;
  constexpr int TLIMIT = TCAP * 13 / 16 - THREADS;  // distinct tail labels admitted (concurrent inserts overshoot by < THREADS)
  static_assert(TLIMIT >= 128, "tail table too small for this CTA size");
  static_assert((R & (R - 1)) == 0 && (TCAP & (TCAP - 1)) == 0 && R % 32 == 0 && H % 32 == 0, "power-of-two tables");
  static_assert(H + TCAP <= 65536, "tail list entries are 16 bit");
  constexpr int RBITS = __builtin_ctz((unsigned)R);
  // shared memory, as 32-bit words from the start of the dynamic segment (accumulator addresses are word indices):
  //   dense words [0, 2H) | tail sums [2H, 2H + 2 TCAP) | sketch [.., + R) | everything else
  constexpr unsigned W_TACC = 2u * H, W_SK = W_TACC + 2u * TCAP;
  extern __shared__ __align__(16) unsigned char smem[];
  const MergeParams& M = P.M;
  RunState* st = M.st;
  if (!st->active) return;
  unsigned int* const sw = reinterpret_cast<unsigned int*>(smem);
  uint2* acc = reinterpret_cast<uint2*>(smem);                                  // dense labels
  uint2* t_acc = reinterpret_cast<uint2*>(sw + W_TACC);                         // tail table: sums
  unsigned int* sk = sw + W_SK;                                                 // sketch buckets (upper bounds, 2^-30 / 2^-27)
  unsigned char* sp = reinterpret_cast<unsigned char*>(sk + R);
  unsigned long long* c_bits = reinterpret_cast<unsigned long long*>(sp); sp += (size_t)CMAX * 8;   // candidates: score bits
  int* c_id = reinterpret_cast<int*>(sp); sp += (size_t)CMAX * 4;                                   // candidates: labels
  int* t_keys = reinterpret_cast<int*>(sp); sp += (size_t)TCAP * 4;          // tail table: label (inserted by pass 2) or ~label (pre)
  uint32_t* s_col = reinterpret_cast<uint32_t*>(sp); sp += (size_t)COLCAP * 4;                      // staged column words
  unsigned int* s_alive = reinterpret_cast<unsigned int*>(sp); sp += (size_t)R / 8;                 // buckets that may hold a kept label
  unsigned short* t_list = reinterpret_cast<unsigned short*>(sp); sp += (size_t)TCAP * 2;           // occupied tail slots
  sp = reinterpret_cast<unsigned char*>(((uintptr_t)sp + 7) & ~(uintptr_t)7);
  DenseShared* DS = reinterpret_cast<DenseShared*>(sp);
  ParShared* S = &DS->P;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int Lp = M.Lp, groups = Lp >> 2, L = M.L;
  const double scale = (M.mode == MODE_GRANK) ? GRANK_HUB_SCALE : MC_HUB_SCALE;
  const double inv = (M.mode == MODE_GRANK) ? GRANK_HUB_INV : MC_HUB_INV;
  const double sk_inv = inv * 4294967296.0;  // one sketch unit
  const int* __restrict__ dense_of = M.g.dense_of;
  const size_t slotb = slot_bytes(Lp);

  for (int i = tid; i < H; i += THREADS) acc[i] = make_uint2(0u, 0u);
  for (int i = tid; i < R; i += THREADS) sk[i] = 0u;
  for (int i = tid; i < TCAP; i += THREADS) { t_keys[i] = KEY_EMPTY; t_acc[i] = make_uint2(0u, 0u); }
  if (tid == 0) { S->tcount = 0; S->spilled = 0; }
  __syncthreads();

  const unsigned char* const rbuf0 = M.buf[st->slot[0]];  // current baskets of colour 0 / 1
  const unsigned char* const rbuf1 = M.buf[st->slot[1]];
  const int write_slot = st->slot[M.colour] ^ 1;
  unsigned long long s_merged = 0, s_edges = 0, s_cands = 0, s_truncs = 0, s_ties = 0, s_bytes = 0, s_nodes = 0, s_requeue = 0;

  const int tlimit = P.tail_limit > 0 && P.tail_limit < TLIMIT ? P.tail_limit : TLIMIT;
  auto home_of = [](unsigned int hk) -> unsigned int { return hk & (unsigned)(TCAP - 1); };
  auto bucket_of = [](unsigned int hk) -> unsigned int { return hk >> (32 - RBITS); };
  // exact accumulate of a tail label (pass 2); false when the key is absent and the table is closed
  auto tail_add = [&](int k, unsigned long long x) -> bool {
    unsigned int h = home_of(hash_key(k));
    volatile int* keys = t_keys;
    for (;;) {
      const int cur = keys[h];
      if (cur == k) break;
      if (cur == KEY_EMPTY) {
        if (*reinterpret_cast<volatile int*>(&S->tcount) >= tlimit) return false;
        const int old = atomicCAS(&t_keys[h], KEY_EMPTY, k);
        if (old == KEY_EMPTY) { const int pos = atomicAdd(&S->tcount, 1); t_list[pos] = (unsigned short)h; break; }
        if (old == k) break;
      }
      h = (h + 1) & (TCAP - 1);
    }
    fixed_add_shared(&t_acc[h], x);
    return true;
  };
  // one contribution (pass 1, any thread): dense word, pre word or sketch bucket. Returns true when it must not be
  // trusted (a product that rounds to zero, a sketch bucket that wrapped): the node is handed over.
  auto contribute = [&](int k, unsigned long long xf) -> bool {
    const unsigned int hk = hash_key(k);
    const unsigned int hm = home_of(hk);
    const bool dense = (unsigned)k < (unsigned)H;
    const bool exact = dense || t_keys[hm] == ~k;
    const unsigned int wi = dense ? 2u * (unsigned)k : (exact ? W_TACC + 2u * hm : W_SK + bucket_of(hk));
    const unsigned int xlo = (unsigned int)xf, xhi = (unsigned int)(xf >> 32);
    const unsigned int lo = exact ? xlo : xhi + 1u;  // sketch: the product rounded UP to one unit
    const unsigned int old = atomicAdd(sw + wi, lo);
    const unsigned int carry = (old + lo) < old ? 1u : 0u;
    const unsigned int hi = exact ? xhi + carry : 0u;
    if (hi) atomicAdd(sw + wi + 1u, hi);
    return xf == 0ull || (!exact && carry);
  };
  auto alive = [&](unsigned int hk) -> bool {
    const unsigned int b = bucket_of(hk);
    return (s_alive[b >> 5] >> (b & 31)) & 1u;
  };
  // pass 2: does label k (any value) need its exact sum from this sweep? tail label, not pre, bucket alive
  auto wanted = [&](int k) -> bool {
    if (k < H) return false;
    const unsigned int hk = hash_key(k);
    return t_keys[home_of(hk)] != ~k && alive(hk);
  };
  auto word_score = [&](uint2 a) -> double { return (double)(long long)(((unsigned long long)a.y << 32) | a.x) * inv; };
  // append to the compact candidate arrays (all 32 lanes call)
  auto append = [&](bool ok, unsigned long long bits, int id) {
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
  };

  unsigned long long dbg[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // (thread 0)
  unsigned long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long t_last = clock64();
#define DPROF_MARK(i) do { if (P.prof && tid == 0) { const long long t_now = clock64(); pc[i] += (unsigned long long)(t_now - t_last); t_last = t_now; } } while (0)

  for (;;) {
    __syncthreads();
    if (tid == 0) { S->item = atomicAdd(&st->work[P.work_idx], 1u); DS->bail = 0; S->ncand = 0; DS->lvl[0] = DS->lvl[1] = DS->lvl[2] = DS->lvl[3] = 0; }
    __syncthreads();
    const unsigned int item = S->item;
    if (item >= (unsigned)(P.n_team_items + P.n_items)) break;
    DPROF_MARK(0);
    const int team = (TEAMS && item < (unsigned)P.n_team_items) ? P.team_item_team[item] : -1;
    const unsigned int ritem = item - (unsigned)P.n_team_items;  // index into the undivided items (team < 0)
    const int p = team >= 0 ? P.team_item_pos[item] : P.item_pos[ritem];
    const long long cb = team >= 0 ? P.team_item_begin[item] : P.item_begin[ritem];
    const int clen = team >= 0 ? P.team_item_len[item] : P.item_len[ritem];
    const long long deg = M.g.row_off[p + 1] - M.g.row_off[p];
    const bool first_chunk = cb == M.g.row_off[p];
    const int team_n = team >= 0 ? P.teams[team].nchunks : 1;
    TeamHeader* const hdr = team >= 0 ? P.team_hdr + team : nullptr;
    unsigned char* const stg = team >= 0 ? P.stage + (size_t)team * P.stage_bytes : nullptr;
    unsigned long long* const g_dense = reinterpret_cast<unsigned long long*>(stg);
    unsigned long long* const g_pre = g_dense + H;
    unsigned int* const g_sk = reinterpret_cast<unsigned int*>(g_pre + TCAP);
    unsigned int* const g_alive = g_sk + R;
    TeamTailEntry* const g_tail = reinterpret_cast<TeamTailEntry*>(g_alive + R / 32);
    // the work item handed to merge_par_kernel when this node cannot be finished here (a team: its undivided item)
    const unsigned int handover_item = team >= 0 ? (unsigned int)P.teams[team].regular_item : (unsigned int)P.item_base + ritem;
    if (team < 0 && deg > (long long)P.chunk) {  // a chunk of a hub split for merge_par_kernel: the general kernel's job
      if (tid == 0) P.fb_queue[atomicAdd(&st->qcount[P.fb_idx], 1u)] = handover_item;
      continue;
    }
    const int self_id = M.g.label[p];
    const double f = M.damping / (double)(unsigned long long)deg;
    const double fscale = f * scale;  // (x * f) * 2^s == x * (f * 2^s): scaling by a power of two commutes with the rounding
    const double self0 = (M.mode == MODE_GRANK) ? M.self_grank : 1.0;
    const unsigned long long xself = (unsigned long long)__double2ll_rn(self0 * scale);
    const unsigned long long xsink = (unsigned long long)__double2ll_rn(((M.mode == MODE_GRANK) ? M.self_grank : 1.0) * fscale);
    const unsigned char* old = M.buf[write_slot ^ 1] + (size_t)p * slotb;

    // The old basket: its smallest score (when full) is the first guess of the new cut; its tail labels -- and the node's
    // own label -- become "pre" labels: exact accumulators at the home slots of the tail table (first come first served;
    // a label that finds its home taken stays an ordinary tail label).
    double theta0 = 0.0;
    {
      const int* oid = reinterpret_cast<const int*>(old);
      const double* osc = reinterpret_cast<const double*>(old + (size_t)Lp * 4);
      unsigned long long mn = ~0ull;
      long long cntv = 0;
      int* const owner = reinterpret_cast<int*>(c_bits);  // (teams) home slot -> smallest old-basket index that wants it
      if (team >= 0) {
        static_assert((size_t)CMAX * 8 >= (size_t)TCAP * 4, "the candidate array doubles as the owner array");
        for (int i = tid; i < TCAP; i += THREADS) owner[i] = 0x7fffffff;
        __syncthreads();
        for (int i = tid; i < Lp + 1; i += THREADS) {
          const int k = i < Lp ? oid[i] : self_id;
          if (k >= H) atomicMin(&owner[home_of(hash_key(k))], i);
        }
        __syncthreads();
      }
      for (int i = tid; i < Lp + 1; i += THREADS) {
        const int k = i < Lp ? oid[i] : self_id;
        if (i < Lp && k >= 0) { const unsigned long long b = (unsigned long long)__double_as_longlong(osc[score_index(i, Lp)]); mn = b < mn ? b : mn; cntv++; }
        if (k >= H) {
          const unsigned int hm = home_of(hash_key(k));
          if (team >= 0) {  // every member of a team must place the same labels in the same slots: lowest index wins
            if (owner[hm] == i) { t_keys[hm] = ~k; t_list[atomicAdd(&S->tcount, 1)] = (unsigned short)hm; }
          } else if (atomicCAS(&t_keys[hm], KEY_EMPTY, ~k) == KEY_EMPTY) {
            t_list[atomicAdd(&S->tcount, 1)] = (unsigned short)hm;
          }
        }
      }
      block_reduce_min_sum(mn, cntv, S->red_a);
      // (a basket that is not full yet still gives a first guess: the level counts below validate whatever is used)
      if (cntv >= (long long)P.min_old) theta0 = __longlong_as_double((long long)mn);
    }
    const int npre = S->tcount;  // (block_reduce_min_sum ends with a barrier)
    if (theta0 == 0.0) {
      // the old basket is not full yet (the first update of a low-degree node): nothing bounds the cut, every label would be
      // a candidate -- the general kernel's job; hand the node over before reading a single successor basket
      __syncthreads();
      for (int i = tid; i < npre; i += THREADS) { const int s2 = t_list[i]; t_keys[s2] = KEY_EMPTY; }
      if (tid == 0) {
        if (first_chunk) { P.fb_queue[atomicAdd(&st->qcount[P.fb_idx], 1u)] = handover_item; s_requeue++; dbg[7]++; }  // (a team: once)
        S->tcount = 0;
      }
      continue;
    }
    bool bad = false;            // this thread saw a contribution that must not be trusted
    if (tid == 0 && first_chunk) {  // the self term (grank.h:101 / mccompletepathv2.h:226)
      bad |= contribute(self_id, xself);
    }

    // One sweep over the node's successor baskets.
    //   pass 1: every entry into its dense word, pre word or sketch bucket
    //   pass 2: tail labels of surviving buckets exactly into the tail table (label words only; scores fetched on a hit)
    unsigned int merged = 0;
    auto sweep = [&](const int pass) {
      const int sj = (NW * 32) / groups, sg = (NW * 32) - sj * groups;  // one warp-grid step in (basket, quad) coordinates
      for (int t0 = 0; t0 < clen; t0 += COLCAP) {
        const int tlen = clen - t0 < COLCAP ? clen - t0 : COLCAP;
        __syncthreads();
        if (tid == 0) DS->ncol = 0;
        __syncthreads();
        // stage the tile's non-sink column words; a sink's basket is the constant {s: 1-d} / {s: 1}: added right here
        for (int j0 = 0; j0 < tlen; j0 += THREADS) {
          const int j = j0 + tid;
          const uint32_t c = j < tlen ? M.g.col[cb + t0 + j] : COL_SINK;
          const bool ns = !(c & COL_SINK);
          if (j < tlen && !ns) {
            const int k = (int)(c & ~COL_SINK);
            if (pass == 1) {
              merged++;
              bad |= contribute(k, xsink);
            } else if (wanted(k)) {
              if (!tail_add(k, xsink)) S->spilled = 1;
            }
          }
          const unsigned m = __ballot_sync(FULL, ns);
          if (m) {
            int base = 0;
            if (lane == (int)(__ffs(m) - 1)) base = atomicAdd(&DS->ncol, __popc(m));
            base = __shfl_sync(FULL, base, __ffs(m) - 1);
            if (ns) s_col[base + __popc(m & ((1u << lane) - 1u))] = c;
          }
        }
        __syncthreads();
        const int total = DS->ncol * groups;  // quads of this tile
        int q = w * 32 + lane;
        int j = q / groups, g = q - j * groups;
        auto base_of = [&](int jj) -> const unsigned char* {
          const uint32_t cc = s_col[jj];
          return (((cc >> COL_COLOUR_SHIFT) & 1u) ? rbuf1 : rbuf0) + (size_t)(cc & COL_POS_MASK) * slotb;
        };
        if (pass == 1) {
          int4 id0 = make_int4(-1, -1, -1, -1);
          double2 a0 = make_double2(0.0, 0.0), b0 = a0;
          if (q < total) {
            const unsigned char* bp = base_of(j) + (size_t)g * 16;
            id0 = __ldg(reinterpret_cast<const int4*>(bp));
            a0 = __ldg(reinterpret_cast<const double2*>(bp + (size_t)Lp * 4));
            b0 = __ldg(reinterpret_cast<const double2*>(bp + (size_t)Lp * 8));
          }
          for (int q0 = w * 32; q0 < total; q0 += NW * 32) {
            // next step's quad
            q += NW * 32; j += sj; g += sg;
            if (g >= groups) { g -= groups; j++; }
            int4 id1 = make_int4(-1, -1, -1, -1);
            double2 a1 = make_double2(0.0, 0.0), b1 = a1;
            if (q < total) {
              const unsigned char* bp = base_of(j) + (size_t)g * 16;
              id1 = __ldg(reinterpret_cast<const int4*>(bp));
              a1 = __ldg(reinterpret_cast<const double2*>(bp + (size_t)Lp * 4));
              b1 = __ldg(reinterpret_cast<const double2*>(bp + (size_t)Lp * 8));
            }
            const int ids[4] = {id0.x, id0.y, id0.z, id0.w};
            const double xs[4] = {a0.x, a0.y, b0.x, b0.y};
            // (invalid entries: id -1, arbitrary score bytes -- computed on, never added)
            unsigned int wi[4], lo[4], xhi[4], oldv[4];
            bool ex[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
              const int k = ids[e];
              const unsigned long long xf = (unsigned long long)__double2ll_rn(xs[e] * fscale);
              const unsigned int hk = hash_key(k);
              const unsigned int hm = home_of(hk);
              const bool dense = (unsigned)k < (unsigned)H;
              ex[e] = dense || t_keys[hm] == ~k;
              wi[e] = dense ? 2u * (unsigned)k : (ex[e] ? W_TACC + 2u * hm : W_SK + bucket_of(hk));
              xhi[e] = (unsigned int)(xf >> 32);
              lo[e] = ex[e] ? (unsigned int)xf : xhi[e] + 1u;
              bad |= k >= 0 && xf == 0ull;
            }
#pragma unroll
            for (int e = 0; e < 4; e++) oldv[e] = ids[e] >= 0 ? atomicAdd(sw + wi[e], lo[e]) : 0u;
#pragma unroll
            for (int e = 0; e < 4; e++) {
              const unsigned int carry = (oldv[e] + lo[e]) < oldv[e] ? 1u : 0u;
              const unsigned int hi = ex[e] ? xhi[e] + carry : 0u;
              if (ids[e] >= 0) {
                if (hi) atomicAdd(sw + wi[e] + 1u, hi);
                bad |= !ex[e] && carry;
              }
            }
            merged += (ids[0] >= 0) + (ids[1] >= 0) + (ids[2] >= 0) + (ids[3] >= 0);
            id0 = id1; a0 = a1; b0 = b1;
          }
        } else {
          int4 id0 = make_int4(-1, -1, -1, -1);
          if (q < total) id0 = __ldg(reinterpret_cast<const int4*>(base_of(j) + (size_t)g * 16));
          for (int q0 = w * 32; q0 < total; q0 += NW * 32) {
            const int jc = j, gc = g;
            q += NW * 32; j += sj; g += sg;
            if (g >= groups) { g -= groups; j++; }
            int4 id1 = make_int4(-1, -1, -1, -1);
            if (q < total) id1 = __ldg(reinterpret_cast<const int4*>(base_of(j) + (size_t)g * 16));
            const int ids[4] = {id0.x, id0.y, id0.z, id0.w};
            bool hit[4];
            bool mine = false;
#pragma unroll
            for (int e = 0; e < 4; e++) { hit[e] = wanted(ids[e]); mine |= hit[e]; }
            if (__any_sync(FULL, mine)) {  // most quads hold no surviving tail label: one vote skips them
              if (mine) {
                const unsigned char* bp = base_of(jc) + (size_t)gc * 16;
                const double2 sa = __ldg(reinterpret_cast<const double2*>(bp + (size_t)Lp * 4));
                const double2 sb = __ldg(reinterpret_cast<const double2*>(bp + (size_t)Lp * 8));
                const double xs[4] = {sa.x, sa.y, sb.x, sb.y};
#pragma unroll
                for (int e = 0; e < 4; e++)
                  if (hit[e] && !tail_add(ids[e], (unsigned long long)__double2ll_rn(xs[e] * fscale))) S->spilled = 1;
              }
            }
            id0 = id1;
          }
        }
      }
    };

    sweep(1);
    DPROF_MARK(1);
    bool bail = __syncthreads_or(bad ? 1 : 0) != 0;  // (also: pass 1 complete)
    bool team_member = false;  // a team member that did not flush last: it only serves pass 2
    if (team >= 0) {
      // add this chunk's sums into the team's staging area; whoever arrives last takes the totals
      bool wrapped = false;
      for (int i = tid; i < H; i += THREADS) {
        const uint2 a = acc[i];
        if ((a.x | a.y) != 0u) atomicAdd(&g_dense[i], ((unsigned long long)a.y << 32) | a.x);
      }
      for (int i = tid; i < npre; i += THREADS) {
        const int sl = t_list[i];
        const uint2 a = t_acc[sl];
        if ((a.x | a.y) != 0u) atomicAdd(&g_pre[sl], ((unsigned long long)a.y << 32) | a.x);
      }
      for (int i = tid; i < R; i += THREADS) {
        const unsigned int v = sk[i];
        if (v) { const unsigned int o = atomicAdd(&g_sk[i], v); wrapped |= o + v < o; }
      }
      if (__syncthreads_or(wrapped ? 1 : 0)) bail = true;
      const unsigned long long mg_chunk = (unsigned long long)block_reduce_sum_ll((long long)merged, S->red_a);
      __threadfence();
      __syncthreads();
      if (tid == 0) {
        if (bail) atomicOr(&hdr->bad, 1u);
        atomicAdd(&hdr->merged, mg_chunk);
        atomicAdd(&hdr->edges, (unsigned int)clen);
        __threadfence();
        DS->team_last = atomicAdd(&hdr->done1, 1u) == (unsigned)team_n - 1u;
        if (DS->team_last) {
          __threadfence();
          DS->team_merged = *reinterpret_cast<volatile unsigned long long*>(&hdr->merged);
          DS->team_edges = *reinterpret_cast<volatile unsigned int*>(&hdr->edges);
        }
      }
      __syncthreads();
      if (DS->team_last) {
        __threadfence();
        for (int i = tid; i < H; i += THREADS) {
          const unsigned long long v = __ldcg(&g_dense[i]);
          acc[i] = make_uint2((unsigned int)v, (unsigned int)(v >> 32));
          if (v) g_dense[i] = 0ull;  // (the staging area is left clean for the next iteration)
        }
        for (int i = tid; i < npre; i += THREADS) {
          const int sl = t_list[i];
          const unsigned long long v = __ldcg(&g_pre[sl]);
          t_acc[sl] = make_uint2((unsigned int)v, (unsigned int)(v >> 32));
          if (v) g_pre[sl] = 0ull;
        }
        for (int i = tid; i < R; i += THREADS) {
          const unsigned int v = __ldcg(&g_sk[i]);
          sk[i] = v;
          if (v) g_sk[i] = 0u;
        }
        if (*reinterpret_cast<volatile unsigned int*>(&hdr->bad) & 1u) bail = true;
      } else {
        team_member = true;
        uint4* z = reinterpret_cast<uint4*>(acc);
        for (int i = tid; i < H / 2; i += THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
        uint4* zs = reinterpret_cast<uint4*>(sk);
        for (int i = tid; i < R / 4; i += THREADS) zs[i] = make_uint4(0u, 0u, 0u, 0u);
        for (int i = tid; i < npre; i += THREADS) t_acc[t_list[i]] = make_uint2(0u, 0u);  // (the pre KEYS stay: pass 2 skips those labels)
      }
      __syncthreads();
    }
    if (team_member) {
      // wait for the finishing member's verdict; serve pass 2 if it asks for it
      if (tid == 0) {
        unsigned int d;
        while ((d = *reinterpret_cast<volatile unsigned int*>(&hdr->decision)) == 0u) __nanosleep(100);
        DS->team_decision = (int)d;
      }
      __syncthreads();
      if (DS->team_decision == 2) {
        __threadfence();
        for (int i = tid; i < R / 32; i += THREADS) s_alive[i] = __ldcg(&g_alive[i]);
        __syncthreads();
        sweep(2);
        __syncthreads();
        const int nt1 = S->tcount;
        if (tid == 0) {
          DS->team_base = atomicAdd(&hdr->tail_count, (unsigned int)(nt1 - npre));
          if (S->spilled) atomicOr(&hdr->bad, 2u);
        }
        __syncthreads();
        for (int i = npre + tid; i < nt1; i += THREADS) {
          const int sl = t_list[i];
          TeamTailEntry e;
          e.key = t_keys[sl];
          e.pad = 0;
          e.acc = ((unsigned long long)t_acc[sl].y << 32) | t_acc[sl].x;
          g_tail[DS->team_base + (unsigned int)(i - npre)] = e;
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) atomicAdd(&hdr->done2, 1u);
      }
      // leave the tables clean; this chunk's share of the statistics
      __syncthreads();
      {
        const int ntl = S->tcount;
        for (int i = tid; i < ntl; i += THREADS) {
          const int sl = t_list[i];
          t_keys[sl] = KEY_EMPTY;
          t_acc[sl] = make_uint2(0u, 0u);
        }
      }
      if (tid == 0) { S->tcount = 0; S->spilled = 0; }  // (this chunk's statistics went into the team header: the finishing member books them)
      continue;
    }

    if (team >= 0 && bail) {  // (finishing member) an untrusted contribution somewhere in the team: everybody stops here
      if (tid == 0) { __threadfence(); *reinterpret_cast<volatile unsigned int*>(&hdr->decision) = 3u; }
    }
    // ---- exact candidates (dense + pre labels) -> compact (score bits, label) arrays ----
    // scan 1 counts how many reach theta0, 3/4, 1/2, 1/16 of it: tau = the highest level that L of them reach is a lower
    // bound of the cut (0 when none is); scan 2 compacts the candidates >= tau.
    bool dropped_local = false;  // this thread saw a candidate that is not in the compact arrays
    bool wide = false;           // more candidates than the compact arrays hold: select / write address the tables themselves
    bool rounds_used = false;    // pass 2 ran slice by slice: its labels live in the compact arrays only
    double tau = 0.0;
    int n = 0;
    if (!bail) {
      if (theta0 > 0.0) {
        const unsigned long long l0 = (unsigned long long)__double_as_longlong(theta0), l1 = (unsigned long long)__double_as_longlong(theta0 * 0.75),
                                 l2 = (unsigned long long)__double_as_longlong(theta0 * 0.5), l3 = (unsigned long long)__double_as_longlong(theta0 * 0.0625);
        int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
        // (a fixed-point word is compared as an integer first: words clearly below the lowest level -- all the empty ones, and
        // the thousands of light labels -- never pay the 64-bit int -> double conversion; 2048 covers its rounding)
        const long long floor3 = (long long)(theta0 * 0.0625 * scale) - 2048;
        auto tally = [&](uint2 a) {
          const long long sfix = (long long)(((unsigned long long)a.y << 32) | a.x);
          if (sfix < floor3 || sfix == 0) return;
          const unsigned long long bits = (unsigned long long)__double_as_longlong((double)sfix * inv);
          c0 += bits >= l0; c1 += bits >= l1; c2 += bits >= l2; c3 += bits >= l3;
        };
        for (int i = tid; i < H; i += THREADS) tally(acc[i]);
        for (int i = tid; i < npre; i += THREADS) tally(t_acc[t_list[i]]);
        c0 = warp_sum_int(c0); c1 = warp_sum_int(c1); c2 = warp_sum_int(c2); c3 = warp_sum_int(c3);
        if (lane == 0) {
          if (c0) atomicAdd(&DS->lvl[0], c0);
          if (c1) atomicAdd(&DS->lvl[1], c1);
          if (c2) atomicAdd(&DS->lvl[2], c2);
          if (c3) atomicAdd(&DS->lvl[3], c3);
        }
        __syncthreads();
        tau = DS->lvl[0] >= L ? theta0 : (DS->lvl[1] >= L ? theta0 * 0.75 : (DS->lvl[2] >= L ? theta0 * 0.5 : (DS->lvl[3] >= L ? theta0 * 0.0625 : 0.0)));
      }
      if (tau == 0.0 && team < 0) {
        // No guess held -- a basket in its first sweeps, whose old content says little about the new cut. The L-th largest
        // EXACT candidate bounds the cut from below just as well (L labels reach it), at the price of one more select:
        // without it every non-empty bucket survives, and these nodes overflow the tail table or the candidate arrays.
        long long nz = 0;
        for (int i = tid; i < H; i += THREADS) { const uint2 a = acc[i]; nz += (a.x | a.y) != 0u; }
        for (int i = tid; i < npre; i += THREADS) { const uint2 a = t_acc[t_list[i]]; nz += (a.x | a.y) != 0u; }
        nz = block_reduce_sum_ll(nz, S->red_a);
        if (nz > (long long)L) {
          auto xbits = [&](int i) -> unsigned long long {
            const uint2 a = i < H ? acc[i] : t_acc[t_list[i - H]];
            return (unsigned long long)__double_as_longlong(word_score(a));
          };
          auto xok = [&](int i) -> bool { const uint2 a = i < H ? acc[i] : t_acc[t_list[i - H]]; return (a.x | a.y) != 0u; };
          bool tie1;
          int krem1;
          tau = __longlong_as_double((long long)block_radix_select(H + npre, L, xbits, xok, S, &tie1, &krem1));
        }
      }
      const unsigned long long tb0 = (unsigned long long)__double_as_longlong(tau);
      const long long floor_tau = (long long)(tau * scale) - 2048;
      for (int i0 = 0; i0 < H; i0 += THREADS) {
        const int i = i0 + tid;
        const uint2 a = i < H ? acc[i] : make_uint2(0u, 0u);
        const long long sfix = (long long)(((unsigned long long)a.y << 32) | a.x);
        const bool nz = sfix != 0;
        unsigned long long bits = 0ull;
        bool ok = false;
        if (nz && sfix >= floor_tau) {
          bits = (unsigned long long)__double_as_longlong((double)sfix * inv);
          ok = bits >= tb0;
        }
        dropped_local |= nz && !ok;
        append(ok, bits, i);
      }
      for (int i0 = 0; i0 < npre; i0 += THREADS) {
        const int i = i0 + tid;
        bool ok = false;
        unsigned long long bits = 0ull;
        int id = 0;
        if (i < npre) {
          const int sl = t_list[i];
          const uint2 a = t_acc[sl];
          id = ~t_keys[sl];
          const bool nz = (a.x | a.y) != 0u;
          bits = (unsigned long long)__double_as_longlong(word_score(a));
          ok = nz && bits >= tb0;
          dropped_local |= nz && !ok;
        }
        append(ok, bits, id);
      }
      __syncthreads();
      n = S->ncand;
      if (n > CMAX) { wide = true; dbg[4]++; }  // (a long run of tied scores at the cut, typically: selected straight from the tables)
      dbg[2] += tau == 0.0;
    }
    const unsigned long long thb = (unsigned long long)__double_as_longlong(tau);  // what the compact arrays were filtered with
    DPROF_MARK(2);

    // ---- sketch buckets that can hold a kept label: upper bound >= tau (tau > 0 only with >= L candidates above it) ----
    if (!bail) {
      // smallest bucket value whose upper bound reaches tau (tau = 0: every non-zero bucket)
      unsigned int sk_floor;
      {
        const double qf = tau / sk_inv;  // (division by a power of two: exact)
        sk_floor = qf >= 4294967295.0 ? 0xffffffffu : (unsigned int)ceil(qf);
        if (sk_floor == 0u) sk_floor = 1u;
      }
      int any = 0;
      for (int i0 = 0; i0 < R; i0 += THREADS) {
        const int i = i0 + tid;
        const unsigned int a = i < R ? sk[i] : 0u;
        const bool al = a >= sk_floor;
        dropped_local |= a != 0u && !al;
        const unsigned m = __ballot_sync(FULL, al);
        if (lane == 0 && i < R) s_alive[i >> 5] = m;
        any |= m != 0u;
      }
      any = __syncthreads_or(any);
      DPROF_MARK(3);
      dbg[1] += any != 0;
      if (team >= 0) {  // (finishing member) tell the team: done, or pass 2 with this alive bitmap
        if (any) {
          for (int i = tid; i < R / 32; i += THREADS) g_alive[i] = s_alive[i];
          __threadfence();
        }
        __syncthreads();
        if (tid == 0) { __threadfence(); *reinterpret_cast<volatile unsigned int*>(&hdr->decision) = any ? 2u : 1u; }
      }
      if (any) {
        // Pass 2, and its second chance. More surviving tail labels than the table admits (hubs of a few thousand successors,
        // ~600 per iteration on R-MAT-22) used to send the node to merge_par_kernel: three full passes on the slow path, and
        // the critical path of a multi-GPU iteration. Instead the sweep is repeated in PASS2_ROUNDS rounds, each over the
        // buckets of one slice of the hash range: a round's labels are summed exactly, those that reach the bound move to the
        // compact candidate arrays, and the table is emptied for the next slice.
        constexpr int PASS2_ROUNDS = 4, ROUND_SHIFT = RBITS - 2;
        int nrounds = 1;
        for (int attempt = 0;; attempt++) {
          for (int r = 0; r < nrounds; r++) {
            if (nrounds > 1) {
              for (int i0 = 0; i0 < R; i0 += THREADS) {  // this round's slice of the alive bitmap
                const int i = i0 + tid;
                const bool al = i < R && sk[i] >= sk_floor && (i >> ROUND_SHIFT) == r;
                const unsigned m = __ballot_sync(FULL, al);
                if (lane == 0 && i < R) s_alive[i >> 5] = m;
              }
              __syncthreads();
            }
            // (the node's own label is normally a pre label; if it lost its home slot its self term went to the sketch too)
            if (tid == 0 && wanted(self_id) && !tail_add(self_id, xself)) S->spilled = 1;
            sweep(2);
            __syncthreads();
            if (team >= 0) {
              // the other members' exact (label, sum) pairs: wait for all of them, then fold them into this table
              if (tid == 0) {
                while (*reinterpret_cast<volatile unsigned int*>(&hdr->done2) < (unsigned)team_n - 1u) __nanosleep(100);
                __threadfence();
                DS->team_base = *reinterpret_cast<volatile unsigned int*>(&hdr->tail_count);
                if (*reinterpret_cast<volatile unsigned int*>(&hdr->bad) & 2u) S->spilled = 1;
              }
              __syncthreads();
              const unsigned int npairs = DS->team_base;
              for (unsigned int i = tid; i < npairs; i += THREADS) {
                if (!tail_add(__ldcg(&g_tail[i].key), __ldcg(&g_tail[i].acc))) S->spilled = 1;
              }
              __syncthreads();
            }
            if (S->spilled) break;
            if (nrounds > 1) {
              const int nt0 = S->tcount;
              for (int i0 = npre; i0 < nt0; i0 += THREADS) {  // this round's labels: candidates that reach the bound, then out
                const int i = i0 + tid;
                bool ok = i < nt0;
                unsigned long long bits = 0ull;
                int id = 0;
                if (ok) {
                  const int sl = t_list[i];
                  id = t_keys[sl];
                  bits = (unsigned long long)__double_as_longlong(word_score(t_acc[sl]));
                  ok = bits >= thb;
                  dropped_local |= !ok;
                  t_keys[sl] = KEY_EMPTY;
                  t_acc[sl] = make_uint2(0u, 0u);
                }
                append(ok, bits, id);
              }
              __syncthreads();
              if (tid == 0) S->tcount = npre;
              __syncthreads();
            }
          }
          if (!S->spilled || attempt == 1 || team >= 0 || wide) break;
          dbg[5]++;
          // second chance: drop what the first attempt put into the table and go round by round
          {
            const int nt0 = S->tcount;
            for (int i = npre + tid; i < nt0; i += THREADS) {
              const int sl = t_list[i];
              t_keys[sl] = KEY_EMPTY;
              t_acc[sl] = make_uint2(0u, 0u);
            }
          }
          __syncthreads();
          if (tid == 0) { S->tcount = npre; S->spilled = 0; }
          __syncthreads();
          nrounds = PASS2_ROUNDS;
        }
        rounds_used = nrounds > 1;
        if (S->spilled) { bail = true; if (!rounds_used) dbg[5]++; }
        if (!bail && !rounds_used) {
          const int nt0 = S->tcount;
          for (int i0 = npre; i0 < nt0; i0 += THREADS) {
            const int i = i0 + tid;
            bool ok = i < nt0;
            unsigned long long bits = 0ull;
            int id = 0;
            if (ok) {
              const int sl = t_list[i];
              id = t_keys[sl];
              bits = (unsigned long long)__double_as_longlong(word_score(t_acc[sl]));
              ok = bits >= thb;
              dropped_local |= !ok;
            }
            append(ok, bits, id);
          }
          __syncthreads();
          n = S->ncand;
          if (n > CMAX && !wide) { wide = true; dbg[4]++; }
        } else if (!bail) {
          n = S->ncand;
          if (n > CMAX) bail = true;  // (the round's labels left the table: no wide select from it)
          else dbg[6]++;
        }
      }
      DPROF_MARK(4);
    }

    int kept = 0, old_cnt = 0;
    if (!bail) {
      // ---- keepTop(L) with the canonical tie-break (pprInternal.h:109-137) ----
      // candidates: the compact arrays, or -- wide -- the tables themselves: dense words [0, H), then the tail table's list
      const int ntl = S->tcount;
      const int nidx = wide ? H + ntl : n;
      auto wbits = [&](int i) -> unsigned long long {
        const uint2 a = i < H ? acc[i] : t_acc[t_list[i - H]];
        return (unsigned long long)__double_as_longlong(word_score(a));
      };
      auto kbits = [&](int i) -> unsigned long long { return wide ? wbits(i) : c_bits[i]; };
      auto kid = [&](int i) -> int {
        if (!wide) return c_id[i];
        if (i < H) return i;
        const int tk = t_keys[t_list[i - H]];
        return tk < 0 ? ~tk : tk;
      };
      auto kok = [&](int i) -> bool {
        if (!wide) return true;
        const unsigned long long b = wbits(i);
        return b != 0ull && b >= thb;
      };
      Threshold thr;
      thr.bits = thb;  // n <= L: everything that reached the bound is kept, nothing below it is
      thr.id_max = 0x7fffffff;
      kept = n;
      const int dropped_any = __syncthreads_or(dropped_local ? 1 : 0);
      if (n > L) {
        kept = L;
        bool tie;
        int krem;
        int ntied = 0;
        if (!wide && n <= 160) {
          // few candidates (the bound leaves little more than L): every thread ranks its candidate against all of them --
          // two barriers instead of the radix select's dozen. The value with (above < L <= above + equal) is the cut.
          for (int i = tid; i < n; i += THREADS) {
            const unsigned long long b = c_bits[i];
            int gt = 0, eq = 0;
#pragma unroll 4
            for (int j = 0; j < n; j++) { const unsigned long long x = c_bits[j]; gt += x > b; eq += x == b; }
            if (gt < L && gt + eq >= L) { DS->cut_bits = b; DS->cut_gt = gt; DS->cut_eq = eq; }  // (same values from every writer)
          }
          __syncthreads();
          thr.bits = DS->cut_bits;
          ntied = DS->cut_eq;
          krem = L - DS->cut_gt;
          tie = DS->cut_gt + DS->cut_eq > L;
          __syncthreads();
        } else {
          thr.bits = block_radix_select(nidx, L, kbits, kok, S, &tie, &krem, &ntied);
        }
        if (tie) {
          const unsigned long long tb = thr.bits;
          if (ntied <= 32) {
            auto densefn = [&](int i) { return dense_of[kid(i)]; };
            thr.id_max = block_small_tie_cut(nidx, krem, tb, kbits, kok, densefn, S);
          } else {
            auto idkey = [&](int i) { return (unsigned long long)(0x7fffffff - dense_of[kid(i)]); };
            auto tied = [&](int i) { return kok(i) && kbits(i) == tb; };
            bool tie2;
            int krem2;
            const unsigned long long tid_key = block_radix_select(nidx, krem, idkey, tied, S, &tie2, &krem2);
            thr.id_max = 0x7fffffff - (int)tid_key;
          }
          s_ties += (tid == 0);
        }
        s_truncs += (tid == 0);
      } else if (dropped_any) {
        s_truncs += (tid == 0);  // exactly L candidates reached the bound and others did not: a cut without a boundary tie
      }
      auto selected = [&](unsigned long long bits, int label) -> bool {
        return bits > thr.bits || (bits == thr.bits && (thr.id_max == 0x7fffffff || dense_of[label] <= thr.id_max));
      };
      DPROF_MARK(5);
      // ---- write B'_v ----
      unsigned char* out = M.buf[write_slot] + (size_t)p * slotb;
      int* out_ids = reinterpret_cast<int*>(out);
      double* out_sc = reinterpret_cast<double*>(out + (size_t)Lp * 4);
      if (tid == 0) S->out_pos = 0;
      __syncthreads();
      long long dsum = 0;
      for (int i = tid; i < nidx; i += THREADS) {
        if (!kok(i)) continue;
        const unsigned long long bits = kbits(i);
        const int id = kid(i);
        if (selected(bits, id)) {
          const int pos = atomicAdd(&S->out_pos, 1);
          const double v = __longlong_as_double((long long)bits);
          out_ids[pos] = id;
          out_sc[score_index(pos, Lp)] = v;  // already multiplied by f = d/outdeg
          dsum += fix_norm(v);
        }
      }
      for (int i = kept + tid; i < Lp; i += THREADS) out_ids[i] = KEY_EMPTY;
      // ---- norm1 against the old basket (pprInternal.h:147-165) ----
      if (M.do_norm) {
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
              uint2 a = make_uint2(0u, 0u);
              if ((unsigned)ids[e] < (unsigned)H) {
                a = acc[ids[e]];
                found = (a.x | a.y) != 0u;
              } else {
                // pre label: exact word at (or, its home taken, not at) the home slot; otherwise in the table iff its
                // bucket survived -- if not, its new score is below the cut
                for (unsigned int h = home_of(hash_key(ids[e]));; h = (h + 1) & (TCAP - 1)) {
                  const int cur = t_keys[h];
                  if (cur == ids[e] || cur == ~ids[e]) { a = t_acc[h]; found = (a.x | a.y) != 0u; break; }
                  if (cur == KEY_EMPTY) break;
                }
              }
              double nv_ = word_score(a);
              if (!found && rounds_used && ids[e] >= H) {  // (an old label without a home slot: among the rounds' candidates, or below the cut)
                for (int j = 0; j < n; j++)
                  if (c_id[j] == ids[e]) { nv_ = __longlong_as_double((long long)c_bits[j]); found = true; break; }
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
      publish_slot(M.peers, write_slot, (size_t)p * slotb, slotb, tid, THREADS, p);
    }
    DPROF_MARK(6);
    // ---- leave the tables clean ----
    __syncthreads();
    {
      uint4* z = reinterpret_cast<uint4*>(acc);  // 16-byte stores: dense words, then the sketch
      for (int i = tid; i < H / 2; i += THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
      uint4* zs = reinterpret_cast<uint4*>(sk);
      for (int i = tid; i < R / 4; i += THREADS) zs[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    {
      const int nt = S->tcount;
      for (int i = tid; i < nt; i += THREADS) {
        const int s = t_list[i];
        t_keys[s] = KEY_EMPTY;
        t_acc[s] = make_uint2(0u, 0u);
      }
    }
    if (bail) {
      // partial sums dropped; merge_par_kernel runs the node from scratch (same sums, global table)
      if (tid == 0) { P.fb_queue[atomicAdd(&st->qcount[P.fb_idx], 1u)] = handover_item; s_requeue++; dbg[3] += (unsigned long long)clen; }
    } else {
      unsigned long long mg = (unsigned long long)block_reduce_sum_ll((long long)merged, S->red_a);
      if (tid == 0) {
        unsigned long long ne = (unsigned long long)clen;
        if (team >= 0) { mg = DS->team_merged; ne = DS->team_edges; }  // the whole team's work
        M.ncand[p] = 0;
        s_merged += mg;
        s_edges += ne;
        s_cands += (unsigned long long)n;
        s_nodes += 1;
        s_bytes += 12ull * mg + 4ull * ne + 12ull * (unsigned long long)old_cnt + 12ull * (unsigned long long)kept + 4ull + 16ull;
      }
    }
    __syncthreads();
    if (tid == 0) { S->tcount = 0; S->spilled = 0; }
    DPROF_MARK(7);
  }
#undef DPROF_MARK
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
    dbg[0] = s_nodes;
    for (int i = 0; i < 8; i++) if (dbg[i]) atomicAdd(&st->dbg[i], dbg[i]);
  }
}

}  // namespace pprb200
