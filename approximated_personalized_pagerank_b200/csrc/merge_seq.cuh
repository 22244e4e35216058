// merge_seq.cuh -- the exact-order merge step (G1), one warp per node.
//
// Reproduces include/grank.h:96-126 (+ keepTop pprInternal.h:109-137, norm1 :147-165) and the MC combine
// step include/mccompletepathv2.h:211-250 for one node v:
//     acc = {v: self}; for s in succ(v) in vector order: for (k, x) in B_s: acc[k] = fma(x, mult, acc[k])
//     keepTop(L) (score desc, dense id asc); B'_v = post * acc; diff = norm1(B'_v, B_v)
// GRank: self = 1-d, mult = d/outdeg, post = 1.   MC: self = 1/f, mult = 1, post = f = d/outdeg.
//
// Bit-exactness: the warp walks the successors strictly in order; the <= L entries of one successor
// basket have distinct keys, so the lanes update distinct table slots in parallel without atomics and
// every key sees its contributions in successor order -- the same fma chain as the reference.
//
// The accumulator is an open-addressing (linear probing) hash table private to the warp: int32 keys,
// fp64 values, plus the list of occupied slots in first-touch order. The table lives in shared memory
// (CAP = 1024 / 4096 / 16384 slots) or, for the rare node whose candidate set outgrows that, in a global
// (L2-resident) workspace; a node that overflows its table is aborted and queued for the next larger one.
#pragma once
#include "device_common.cuh"

namespace pprb200 {

struct MergeParams {
  GraphDev g;
  unsigned char* buf[2];     // basket buffers
  RunState* st;
  int Lp, L;
  int mode;                  // MODE_GRANK / MODE_MC
  double damping;
  double self_grank;         // 1 - d
  int colour;                // colour processed by this launch (selects the write slot)
  int all_colours_same_slot; // MC combine: every node flips together
  // work source: either a contiguous storage range or a queue of positions
  int range_begin, range_end;
  const unsigned int* queue_in;    // nullable
  int queue_in_idx;                // index into st->qcount (length of queue_in), -1 = range
  unsigned int* queue_out;         // nodes that overflowed this launch's table
  int queue_out_idx;
  int work_idx;                    // index into st->work
  int limit;                       // max distinct candidates this launch may hold (<= CAP - Lp - 1)
  int* ncand;                      // [M] distinct-candidate count of the node's latest update (class prediction)
  int do_norm;                     // GRank: 1; MC: 0
  PeerDev peers;                   // multi-GPU peers that receive every basket this launch writes
  const int* work_list;            // multi-GPU: this rank's positions of the range (range = indices into the list)
  int n_ids;                       // number of nodes (dense ids are < n_ids)
  int init_mode;                   // GRank init (grank.h:64-83): every successor s contributes {s: +factor}; writes the current slot
};

template <typename IdxT>
struct WarpTable {
  int* keys;
  double* vals;
  IdxT* list;
  unsigned int mask;
  int identity_hash;  // table at least as large as the id space: slot = id (never collides)
};

template <typename IdxT>
__device__ __forceinline__ int table_find_or_insert(const WarpTable<IdxT>& t, int k, bool* is_new) {
  unsigned int h = (t.identity_hash ? (unsigned int)k : hash_key(k)) & t.mask;
  volatile int* keys = t.keys;
  for (;;) {
    const int cur = keys[h];
    if (cur == k) { *is_new = false; return (int)h; }
    if (cur == KEY_EMPTY) {
      const int old = atomicCAS(&t.keys[h], KEY_EMPTY, k);
      if (old == KEY_EMPTY) { *is_new = true; return (int)h; }
      if (old == k) { *is_new = false; return (int)h; }
    }
    h = (h + 1) & t.mask;
  }
}

template <typename IdxT>
__device__ __forceinline__ int table_find(const WarpTable<IdxT>& t, int k) {
  unsigned int h = (t.identity_hash ? (unsigned int)k : hash_key(k)) & t.mask;
  for (;;) {
    const int cur = t.keys[h];
    if (cur == k) return (int)h;
    if (cur == KEY_EMPTY) return -1;
    h = (h + 1) & t.mask;
  }
}

// one successor basket held by a warp: lane g < Lp/4 owns entries 4g..4g+3 (rounds of 32 lanes for Lp > 128)
struct BasketFrag {
  int4 id;
  double2 sa, sb;
};

__device__ __forceinline__ void load_frag(const unsigned char* slot, int Lp, int g, BasketFrag* f) {
  const int4* ids = reinterpret_cast<const int4*>(slot);
  const double2* sc = reinterpret_cast<const double2*>(slot + (size_t)Lp * 4);
  f->id = __ldg(ids + g);
  // scores are only fetched for valid entries (baskets are filled front to back)
  if (f->id.x >= 0) f->sa = __ldg(sc + g); else f->sa = make_double2(0.0, 0.0);
  if (f->id.z >= 0) f->sb = __ldg(sc + (Lp >> 2) + g); else f->sb = make_double2(0.0, 0.0);
}

// Processes node at storage position p. Returns false when the table overflowed (node must be requeued).
// Per-warp statistics are accumulated into the reference arguments.
template <typename IdxT>
__device__ bool merge_node_seq(const MergeParams& P, const WarpTable<IdxT>& T, unsigned int* hist, int* s_count,
                               int p, int write_slot, int read_slot0, int read_slot1, unsigned long long& st_merged,
                               unsigned long long& st_edges, unsigned long long& st_cands,
                               unsigned long long& st_truncs, unsigned long long& st_ties,
                               unsigned long long& st_bytes, long long& warp_maxdiff) {
  const int lane = lane_id();
  const int Lp = P.Lp;
  const int groups = Lp >> 2;
  const long long rb = P.g.row_off[p], re = P.g.row_off[p + 1];
  const long long deg = re - rb;
  const int self_id = P.g.label[p];
  const double f = P.damping / (double)(unsigned long long)deg;  // grank.h:105 / mccompletepathv2.h:214
  const double mult = (P.mode == MODE_GRANK) ? f : 1.0;
  const double post = (P.mode == MODE_GRANK) ? 1.0 : f;
  const double self0 = (P.mode == MODE_GRANK) ? P.self_grank : 1.0 / f;  // grank.h:101 / mccompletepathv2.h:226

  if (lane == 0) {
    bool nw;
    const int s = table_find_or_insert(T, self_id, &nw);
    T.vals[s] = self0;
    T.list[0] = (IdxT)s;
  }
  int cnt = 1;  // distinct keys so far (warp-uniform); list[] holds their slots in first-touch order
  __syncwarp();

  unsigned long long merged = 0;
  bool overflow = false;
  // one accumulate step, executed by all 32 lanes (k < 0: nothing to add): acc[k] = fma(x, mult, acc[k]); new keys are
  // appended to the first-touch list with one ballot instead of a contended shared-memory counter
  auto add_entry = [&](int k, double x) {
    bool nw = false;
    int s = 0;
    if (k >= 0) {
      s = table_find_or_insert(T, k, &nw);
      const double a = nw ? 0.0 : T.vals[s];
      T.vals[s] = fma(x, mult, a);  // grank.h:115 (mult = 1: exactly acc + x, mccompletepathv2.h:241)
    }
    const unsigned m = __ballot_sync(FULL, nw);
    if (nw) T.list[cnt + __popc(m & ((1u << lane) - 1u))] = (IdxT)s;
    cnt += __popc(m);
  };
  for (long long eb = rb; eb < re && !overflow; eb += 32) {
    const int chunk = (int)((re - eb) < 32 ? (re - eb) : 32);
    const uint32_t mycol = (lane < chunk) ? P.g.col[eb + lane] : 0u;
    if (P.init_mode) {
      // grank.h:79-80: scores[v][succ] += factor, once per occurrence. All increments are equal, so only the
      // multiplicity matters: lanes holding the same successor elect a leader that adds `factor` m times.
      int k = -2 - lane;
      if (lane < chunk) k = (mycol & COL_SINK) ? (int)(mycol & ~COL_SINK) : P.g.label[mycol & COL_POS_MASK];
      const unsigned peers = __match_any_sync(FULL, k);
      bool nw = false;
      int s = 0;
      if (lane < chunk && (int)(__ffs(peers) - 1) == lane) {
        s = table_find_or_insert(T, k, &nw);
        double a = nw ? 0.0 : T.vals[s];
        for (int r = __popc(peers); r > 0; r--) a += mult;
        T.vals[s] = a;
      }
      const unsigned m = __ballot_sync(FULL, nw);
      if (nw) T.list[cnt + __popc(m & ((1u << lane) - 1u))] = (IdxT)s;
      cnt += __popc(m);
      merged += (lane < chunk);
      __syncwarp();
      if (cnt > P.limit) { overflow = true; }
      continue;
    }
    // successor baskets are fetched two steps ahead of the one being merged (lane g holds entries 4g..4g+3)
    auto fetch = [&](int j, BasketFrag* fr) {
      fr->id = make_int4(-1, -1, -1, -1);
      fr->sa = fr->sb = make_double2(0.0, 0.0);
      const uint32_t c = __shfl_sync(FULL, mycol, j < chunk ? j : 0);
      if (j < chunk && !(c & COL_SINK) && lane < groups) {
        const unsigned char* slot = P.buf[((c >> COL_COLOUR_SHIFT) & 1u) ? read_slot1 : read_slot0] + (size_t)(c & COL_POS_MASK) * slot_bytes(Lp);
        const int4* ids = reinterpret_cast<const int4*>(slot);
        const double2* sc = reinterpret_cast<const double2*>(slot + (size_t)Lp * 4);
        fr->id = __ldg(ids + lane);
        fr->sa = __ldg(sc + lane);
        fr->sb = __ldg(sc + (Lp >> 2) + lane);
      }
    };
    BasketFrag f0, f1;
    fetch(0, &f0);
    fetch(1, &f1);
    for (int j = 0; j < chunk; j++) {
      BasketFrag f2;
      fetch(j + 2, &f2);
      const uint32_t c = __shfl_sync(FULL, mycol, j);
      if (c & COL_SINK) {
        // sink successor: its basket is the constant {s: 1-d} (GRank) / {s: 1} (MC)
        add_entry(lane == 0 ? (int)(c & ~COL_SINK) : -1, (P.mode == MODE_GRANK) ? P.self_grank : 1.0);
        merged += (lane == 0);
      } else {
        const unsigned char* slot = P.buf[((c >> COL_COLOUR_SHIFT) & 1u) ? read_slot1 : read_slot0] + (size_t)(c & COL_POS_MASK) * slot_bytes(Lp);
        for (int g0 = 0; g0 < groups; g0 += 32) {
          BasketFrag fr;
          if (g0 == 0) fr = f0;
          else {
            fr.id = make_int4(-1, -1, -1, -1);
            fr.sa = fr.sb = make_double2(0.0, 0.0);
            if (g0 + lane < groups) load_frag(slot, Lp, g0 + lane, &fr);
          }
          const int ids[4] = {fr.id.x, fr.id.y, fr.id.z, fr.id.w};
          const double xs[4] = {fr.sa.x, fr.sa.y, fr.sb.x, fr.sb.y};
#pragma unroll
          for (int e = 0; e < 4; e++) {
            add_entry(ids[e], xs[e]);
            merged += (ids[e] >= 0);
          }
        }
      }
      __syncwarp();
      if (cnt > P.limit) { overflow = true; break; }
      f0 = f1;
      f1 = f2;
    }
  }
  const int n = cnt;
  __syncwarp();
  if (overflow) {
    for (int i = lane; i < n; i += 32) T.keys[T.list[i]] = KEY_EMPTY;
    __syncwarp();
    return false;
  }

  // ---- keepTop(L) ----
  const int L = P.L;
  Threshold th;
  th.bits = 0ull;
  th.id_max = 0x7fffffff;
  int kept = n;
  if (n > L) {
    kept = L;
    bool tie;
    int krem;
    auto keyfn = [&](int i) { return (unsigned long long)__double_as_longlong(T.vals[T.list[i]]); };
    auto all = [&](int) { return true; };
    int ntied = 0;
    th.bits = warp_radix_select(n, L, keyfn, all, hist, &tie, &krem, &ntied);
    if (tie) {
      const unsigned long long tb = th.bits;
      if (ntied <= 32) {
        // few tied candidates (the usual case): gather their dense ids and rank them with shuffles
        int* sd = reinterpret_cast<int*>(hist);
        int off = 0;
        for (int i0 = 0; i0 < n; i0 += 32) {
          const int i = i0 + lane;
          const bool in = i < n && (unsigned long long)__double_as_longlong(T.vals[T.list[i]]) == tb;
          const unsigned m = __ballot_sync(FULL, in);
          if (in) sd[off + __popc(m & ((1u << lane) - 1u))] = P.g.dense_of[T.keys[T.list[i]]];
          off += __popc(m);
        }
        __syncwarp();
        const int myd = lane < off ? sd[lane] : 0x7fffffff;
        int rank = 0;
        for (int j = 0; j < off; j++) rank += __shfl_sync(FULL, myd, j) < myd;
        const unsigned who = __ballot_sync(FULL, lane < off && rank == krem - 1);
        th.id_max = __shfl_sync(FULL, myd, __ffs(who) - 1);
        __syncwarp();
      } else {
        auto idkey = [&](int i) { return (unsigned long long)(0x7fffffff - P.g.dense_of[T.keys[T.list[i]]]); };
        auto tied = [&](int i) { return (unsigned long long)__double_as_longlong(T.vals[T.list[i]]) == tb; };
        bool tie2;
        int krem2;
        const unsigned long long tid = warp_radix_select(n, krem, idkey, tied, hist, &tie2, &krem2);
        th.id_max = 0x7fffffff - (int)tid;
      }
      st_ties += (lane == 0);
    }
    st_truncs += (lane == 0);
  }

  // ties are cut by dense id: only entries whose score equals the threshold need the label -> dense lookup
  auto selected = [&](unsigned long long bits, int label) -> bool {
    return bits > th.bits || (bits == th.bits && (th.id_max == 0x7fffffff || P.g.dense_of[label] <= th.id_max));
  };
  // ---- write B'_v, norm1 vs B_v ----
  unsigned char* out = P.buf[write_slot] + (size_t)p * slot_bytes(Lp);
  int* out_ids = reinterpret_cast<int*>(out);
  double* out_sc = reinterpret_cast<double*>(out + (size_t)Lp * 4);
  long long dsum = 0;
  int base = 0;
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + lane;
    bool sel = false;
    int id = 0;
    double v = 0.0;
    if (i < n) {
      const int s = T.list[i];
      id = T.keys[s];
      v = T.vals[s];
      sel = selected((unsigned long long)__double_as_longlong(v), id);
    }
    const unsigned m = __ballot_sync(FULL, sel);
    if (sel) {
      const int pos = base + __popc(m & ((1u << lane) - 1u));
      const double w = v * post;  // mccompletepathv2.h:246-247
      out_ids[pos] = id;
      out_sc[score_index(pos, Lp)] = w;
      dsum += fix_norm(w);
    }
    base += __popc(m);
  }
  for (int i = kept + lane; i < Lp; i += 32) out_ids[i] = KEY_EMPTY;

  int old_cnt = 0;
  if (P.do_norm) {
    // norm1 (pprInternal.h:147-165): sum_{k in new}|new_k - old_k| + sum_{k in old \ new} old_k
    const unsigned char* old = P.buf[write_slot ^ 1] + (size_t)p * slot_bytes(Lp);
    for (int g = lane; g < groups; g += 32) {
      BasketFrag fr;
      load_frag(old, Lp, g, &fr);
      const int ids[4] = {fr.id.x, fr.id.y, fr.id.z, fr.id.w};
      const double xs[4] = {fr.sa.x, fr.sa.y, fr.sb.x, fr.sb.y};
#pragma unroll
      for (int e = 0; e < 4; e++) {
        if (ids[e] >= 0) {
          old_cnt++;
          const int s = table_find(T, ids[e]);
          bool in_new = false;
          double nv = 0.0;
          if (s >= 0) {
            nv = T.vals[s];
            in_new = selected((unsigned long long)__double_as_longlong(nv), ids[e]);
          }
          if (in_new) dsum += fix_norm(fabs(nv - xs[e])) - fix_norm(nv);
          else dsum += fix_norm(xs[e]);
        }
      }
    }
    dsum = warp_sum_ll(dsum);
    old_cnt = warp_sum_int(old_cnt);
    if (dsum > warp_maxdiff) warp_maxdiff = dsum;
  }
  __syncwarp();
  publish_slot(P.peers, write_slot, (size_t)p * slot_bytes(Lp), slot_bytes(Lp), lane, 32, p);
  // reset the table through the occupied-slot list
  for (int i = lane; i < n; i += 32) T.keys[T.list[i]] = KEY_EMPTY;
  if (lane == 0) {
    P.ncand[p] = n;
    st_edges += (unsigned long long)deg;
    st_cands += (unsigned long long)n;
  }
  merged = warp_sum_ull(merged);
  if (lane == 0) {
    st_merged += merged;
    st_bytes += 12ull * merged + 12ull * (unsigned long long)old_cnt + 12ull * (unsigned long long)kept + 4ull +
                4ull * (unsigned long long)deg + 16ull;
  }
  __syncwarp();
  return true;
}

// Persistent kernel: WARPS warps per CTA, each with a private CAP-slot table in shared memory
// (CAP == 0: table in the global workspace `ws`, ws_cap slots per warp).
template <int CAP, int WARPS, typename IdxT>
__global__ void __launch_bounds__(WARPS * 32, 1) merge_seq_kernel(MergeParams P, unsigned char* ws, unsigned int ws_cap,
                                                               int ws_identity) {
  extern __shared__ __align__(16) unsigned char smem[];
  RunState* st = P.st;
  if (!st->active) return;
  const int lane = lane_id();
  const int w = threadIdx.x >> 5;

  WarpTable<IdxT> T;
  unsigned int* hist;
  int* s_count;
  unsigned int cap;
  if (CAP > 0) {
    cap = CAP;
    unsigned char* base = smem + (size_t)w * ((size_t)CAP * (12 + sizeof(IdxT)) + 1024 + 16);
    T.vals = reinterpret_cast<double*>(base);
    T.keys = reinterpret_cast<int*>(base + (size_t)CAP * 8);
    T.list = reinterpret_cast<IdxT*>(base + (size_t)CAP * 12);
    hist = reinterpret_cast<unsigned int*>(base + (size_t)CAP * (12 + sizeof(IdxT)));
    s_count = reinterpret_cast<int*>(base + (size_t)CAP * (12 + sizeof(IdxT)) + 1024);
    T.identity_hash = 0;
  } else {
    cap = ws_cap;
    const size_t gw = (size_t)blockIdx.x * WARPS + w;
    unsigned char* base = ws + gw * ((size_t)ws_cap * (12 + sizeof(IdxT)));
    T.vals = reinterpret_cast<double*>(base);
    T.keys = reinterpret_cast<int*>(base + (size_t)ws_cap * 8);
    T.list = reinterpret_cast<IdxT*>(base + (size_t)ws_cap * 12);
    hist = reinterpret_cast<unsigned int*>(smem + (size_t)w * (1024 + 16));
    s_count = reinterpret_cast<int*>(smem + (size_t)w * (1024 + 16) + 1024);
    T.identity_hash = ws_identity;
  }
  T.mask = cap - 1;
  bool table_clean = false;  // cleared lazily: a warp that never gets work never touches its table

  const int write_slot = P.init_mode ? st->slot[P.colour] : (st->slot[P.colour] ^ 1);
  const int read_slot0 = st->slot[0], read_slot1 = st->slot[1];

  unsigned int total;
  if (P.queue_in_idx >= 0) total = st->qcount[P.queue_in_idx];
  else total = (unsigned int)(P.range_end - P.range_begin);

  unsigned long long s_merged = 0, s_edges = 0, s_cands = 0, s_truncs = 0, s_ties = 0, s_bytes = 0, s_nodes = 0, s_requeue = 0;
  long long maxdiff = 0;
  // nodes are fetched in small batches: thousands of tiny nodes hammering one global counter would serialise there
  // (only when there are plenty of nodes per warp: with few nodes a batch would serialise them on a handful of warps)
  const unsigned int batch = (P.queue_in_idx >= 0 || CAP == 0 || total < 16u * gridDim.x * WARPS) ? 1u : 4u;
  for (;;) {
    unsigned int idx0 = 0;
    if (lane == 0) idx0 = atomicAdd(&st->work[P.work_idx], batch);
    idx0 = __shfl_sync(FULL, idx0, 0);
    if (idx0 >= total) break;
    if (!table_clean) {
      for (unsigned int i = lane; i < cap; i += 32) T.keys[i] = KEY_EMPTY;
      __syncwarp();
      table_clean = true;
    }
    for (unsigned int idx = idx0; idx < idx0 + batch && idx < total; idx++) {
      const int p = (P.queue_in_idx >= 0) ? (int)P.queue_in[idx] : (P.work_list ? P.work_list[P.range_begin + (int)idx] : P.range_begin + (int)idx);
      bool ok = false;
      // class prediction: a node whose previous candidate count already exceeds this table goes straight on
      const bool skip = P.queue_out != nullptr && P.ncand[p] > P.limit;
      if (!skip)
        ok = merge_node_seq<IdxT>(P, T, hist, s_count, p, write_slot, read_slot0, read_slot1, s_merged, s_edges, s_cands, s_truncs,
                                  s_ties, s_bytes, maxdiff);
      if (!ok) {
        if (lane == 0) {
          const unsigned int q = atomicAdd(&st->qcount[P.queue_out_idx], 1u);
          P.queue_out[q] = (unsigned int)p;
          s_requeue += skip ? 0 : 1;
        }
      } else {
        s_nodes += (lane == 0);
      }
    }
  }
  if (lane == 0) {
    if (s_nodes) atomicAdd(&st->node_iters, s_nodes);
    if (s_edges) atomicAdd(&st->edge_reads, s_edges);
    if (s_merged) atomicAdd(&st->merged, s_merged);
    if (s_cands) atomicAdd(&st->cands, s_cands);
    if (s_truncs) atomicAdd(&st->truncs, s_truncs);
    if (s_ties) atomicAdd(&st->ties, s_ties);
    if (s_bytes) atomicAdd(&st->abytes, s_bytes);
    if (s_requeue) atomicAdd(&st->requeues, s_requeue);
    if (maxdiff > 0) atomicMax(&st->cur_max, maxdiff);
  }
}

}  // namespace pprb200
