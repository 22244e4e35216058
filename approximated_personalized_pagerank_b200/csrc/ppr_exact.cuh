// ppr_exact.cuh -- batched exact Personalized PageRank by power iteration: the yardstick of the quality evaluator
// (reference include/internal/pprSingleSource.h:28-75, called once per sampled node by
// include/benchmarkAlgorithm.h:91).
//
// The reference pushes score along out-edges into a hash map, one source at a time. Here B sources advance together:
// X[n][B] (node-major, the B scores of a node are contiguous) and one iteration is a PULL over the transposed CSR,
//     Y[v][b] = (1-d) * [v == source_b] + sum_{u in pred(v)} X[u][b] * d / outdeg(u),
// one warp per node v, lanes over b: every gather of a predecessor row is a fully coalesced B*8-byte read, every
// sum has a fixed order (predecessors by ascending id) and there are no atomics on the scores. The norm-1 change of
// each source (pprSingleSource.h:66) is summed in 2^-61 fixed point (order-free); a source whose change drops below
// the tolerance is frozen at that iteration, as the reference's loop condition does (:47), while the others go on.
// Values differ from the reference's in the last bits only (summation order).
#pragma once
#include "device_common.cuh"

namespace pprb200 {

struct ExactParams {
  const long long* prow;   // [n+1] transposed CSR
  const int* pcol;         // [E] predecessors
  const double* factor;    // [n] d / outdeg(u) (unused for sinks: they are nobody's predecessor)
  double* buf[2];          // two [n][B] score arrays; *parity says which one holds the current scores
  int* parity;             // flipped by the step kernel after every iteration that still had an active source
  const int* source;       // [B]
  int* active;             // [B] 1 while the source iterates
  long long* diff;         // [B] fixed-point norm-1 change of this iteration
  unsigned int* iters;     // [B] iterations executed
  int n, B;
  const int* heavy;        // nodes with more than heavy_threshold predecessors: one CTA each (ppr_exact_heavy_kernel)
  int n_heavy, heavy_threshold;
  double teleport;         // 1 - d
  double tolerance;
};

constexpr int EXACT_HEAVY = 64;   // in-degree above which a node gets a CTA of its own (power-law in-degrees: one warp
                                   // would walk tens of thousands of predecessor rows alone)

// one warp's share of a predecessor range [pb, pe): chunks of 32 predecessors (lane-loaded, broadcast by shuffle), four
// rows in flight; acc[j] += X[u][lane + 32 j] * factor[u] in predecessor order
template <int BT>
__device__ __forceinline__ void exact_gather(const ExactParams& P, const double* __restrict__ X, long long pb, long long pe,
                                             long long first_chunk, long long chunk_stride, const bool (&act)[BT], double (&acc)[BT]) {
  const int lane = threadIdx.x & 31;
  const int B = P.B;
  for (long long e0 = pb + first_chunk * 32; e0 < pe; e0 += chunk_stride * 32) {
    const int cnt = (int)(pe - e0 < 32 ? pe - e0 : 32);
    int u_l = 0;
    double f_l = 0.0;
    if (lane < cnt) { u_l = __ldg(P.pcol + e0 + lane); f_l = __ldg(P.factor + u_l); }
    for (int i = 0; i < cnt; i += 4) {
      int u[4];
      double f[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {  // lanes beyond cnt hold u = 0, f = 0: a row of node 0 times zero adds nothing
        u[k] = __shfl_sync(FULL, u_l, (i + k) & 31);
        f[k] = __shfl_sync(FULL, f_l, (i + k) & 31);
        if (i + k >= cnt) { u[k] = 0; f[k] = 0.0; }
      }
      double x[4][BT];
#pragma unroll
      for (int k = 0; k < 4; k++)
#pragma unroll
        for (int j = 0; j < BT; j++) x[k][j] = act[j] ? __ldg(X + (size_t)u[k] * B + lane + 32 * j) : 0.0;
#pragma unroll
      for (int k = 0; k < 4; k++)
#pragma unroll
        for (int j = 0; j < BT; j++) acc[j] = fma(x[k][j], f[k], acc[j]);
    }
  }
}

// grid-stride over nodes, one warp per node; BT = ceil(B / 32) scores per lane
template <int BT>
__global__ void __launch_bounds__(256) ppr_exact_iter_kernel(ExactParams P) {
  const int lane = threadIdx.x & 31;
  const int warp = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  const int nwarps = (int)((gridDim.x * (unsigned)blockDim.x) >> 5);
  const int B = P.B;
  int src[BT];
  bool act[BT];
  long long dsum[BT];
#pragma unroll
  for (int j = 0; j < BT; j++) {
    const int b = lane + 32 * j;
    src[j] = b < B ? P.source[b] : -1;
    act[j] = b < B && P.active[b] != 0;
    dsum[j] = 0;
  }
  bool any = false;
#pragma unroll
  for (int j = 0; j < BT; j++) any |= act[j];
  if (!__any_sync(FULL, any)) return;  // every source of this batch has converged (the step kernel stops flipping too)
  const int par = *P.parity;
  const double* __restrict__ X = P.buf[par];
  double* __restrict__ Y = P.buf[par ^ 1];
  for (int v = warp; v < P.n; v += nwarps) {
    double acc[BT];
#pragma unroll
    for (int j = 0; j < BT; j++) acc[j] = 0.0;
    const long long pb = P.prow[v], pe = P.prow[v + 1];
    if (pe - pb > P.heavy_threshold) continue;  // ppr_exact_heavy_kernel
    exact_gather<BT>(P, X, pb, pe, 0, 1, act, acc);
#pragma unroll
    for (int j = 0; j < BT; j++) {
      const int b = lane + 32 * j;
      if (b < B) {
        const double old = X[(size_t)v * B + b];
        double val = old;
        if (act[j]) {
          val = acc[j] + (src[j] == v ? P.teleport : 0.0);
          dsum[j] += fix_norm(fabs(val - old));
        }
        Y[(size_t)v * B + b] = val;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < BT; j++)
    if (act[j] && dsum[j]) atomicAdd(reinterpret_cast<unsigned long long*>(P.diff + lane + 32 * j), (unsigned long long)dsum[j]);
}

// one CTA (8 warps) per heavy node: warp w takes predecessor chunks w, w + 8, ...; the eight partial rows are summed in
// warp order (a fixed order: results do not depend on scheduling)
template <int BT>
__global__ void __launch_bounds__(256) ppr_exact_heavy_kernel(ExactParams P) {
  __shared__ double part[8][BT * 32];
  __shared__ long long bdiff[BT * 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int B = P.B;
  bool act[BT];
  bool any = false;
#pragma unroll
  for (int j = 0; j < BT; j++) { const int b = lane + 32 * j; act[j] = b < B && P.active[b] != 0; any |= act[j]; }
  if (!__syncthreads_or(any ? 1 : 0)) return;
  const int par = *P.parity;
  const double* __restrict__ X = P.buf[par];
  double* __restrict__ Y = P.buf[par ^ 1];
  for (int t = threadIdx.x; t < BT * 32; t += 256) bdiff[t] = 0;
  for (int h = blockIdx.x; h < P.n_heavy; h += gridDim.x) {
    const int v = P.heavy[h];
    double acc[BT];
#pragma unroll
    for (int j = 0; j < BT; j++) acc[j] = 0.0;
    exact_gather<BT>(P, X, P.prow[v], P.prow[v + 1], w, 8, act, acc);
#pragma unroll
    for (int j = 0; j < BT; j++) part[w][lane + 32 * j] = acc[j];
    __syncthreads();
    for (int b = threadIdx.x; b < B; b += 256) {
      const double old = X[(size_t)v * B + b];
      double val = old;
      if (P.active[b]) {
        double sacc = 0.0;
        for (int k = 0; k < 8; k++) sacc += part[k][b];
        val = sacc + (P.source[b] == v ? P.teleport : 0.0);
        bdiff[b] += fix_norm(fabs(val - old));
      }
      Y[(size_t)v * B + b] = val;
    }
    __syncthreads();
  }
  for (int b = threadIdx.x; b < B; b += 256)
    if (bdiff[b]) atomicAdd(reinterpret_cast<unsigned long long*>(P.diff + b), (unsigned long long)bdiff[b]);
}

// after every iteration: pprSingleSource.h:47 `i < iterations && diff >= tolerance`. One block.
__global__ void ppr_exact_step_kernel(ExactParams P) {
  int was_active = 0;
  for (int b = threadIdx.x; b < P.B; b += blockDim.x) {
    if (P.active[b]) {
      was_active = 1;
      P.iters[b] += 1u;
      const double diff = (double)P.diff[b] * NORM_INV;
      if (!(diff >= P.tolerance)) P.active[b] = 0;
    }
    P.diff[b] = 0;
  }
  was_active = __syncthreads_or(was_active);
  if (threadIdx.x == 0 && was_active) *P.parity ^= 1;
}

__global__ void ppr_exact_init_kernel(double* x, const int* source, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) x[(size_t)source[b] * B + b] = 1.0;  // pprSingleSource.h:43
}

}  // namespace pprb200
