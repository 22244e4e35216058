// plan_device.cuh -- the two parts of the front half that are graph traversals over every edge, on
// the device. The caller's CSR is uploaded once as it is; what comes back is a byte per node (BFS level) -- the encoded
// column words never exist on the host at all.
//
//   * bfs_level_kernel: one level of the 2-colouring BFS (pprInternal.h:29-99) of the first non-trivial component.
//     colour(v) = parity of the undirected BFS distance from the component's root (host_graph.cc: find_partitions), so
//     the level array is all that is needed. Direction-free and transpose-free like the host version: a frontier node
//     marks its unseen successors (top-down over out-edges), and an unseen node that has a successor in the frontier is a
//     predecessor of the frontier, hence in the next level (bottom-up over its own out-edges). Eight lanes per node.
//   * encode_kernel: column words in storage order, enc[row_off[p] + i] = word_of[col[row_ptr[order[p]] + i]]
//     (one warp per storage position; the 16 MB word table is L2-resident).
#pragma once
#include <cstdint>

namespace pprb200 {

constexpr unsigned char BFS_UNSEEN = 0xFF;
constexpr int BFS_MAX_LEVEL = 250;  // levels are bytes: a component deeper than this is coloured by the host

__global__ void __launch_bounds__(256) bfs_level_kernel(const long long* __restrict__ row_ptr, const int* __restrict__ col, int n,
                                                         unsigned char* level, unsigned char cur, unsigned int* changed) {
  const int lane8 = threadIdx.x & 7;
  const unsigned gmask = 0xFFu << (threadIdx.x & 24);
  const long long stride = ((long long)gridDim.x * blockDim.x) >> 3;
  bool any = false;
  for (long long v = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3; v < n; v += stride) {
    const unsigned char lv = level[v];
    if (lv != cur && lv != BFS_UNSEEN) continue;
    const long long rb = row_ptr[v], re = row_ptr[v + 1];
    if (lv == cur) {
      for (long long i = rb + lane8; i < re; i += 8) {
        const int s = col[i];
        if (level[s] == BFS_UNSEEN) { level[s] = (unsigned char)(cur + 1); any = true; }
      }
    } else {
      bool found = false;
      for (long long i0 = rb; i0 < re && !found; i0 += 8) {
        const long long i = i0 + lane8;
        const bool hit = i < re && level[col[i]] == cur;
        found = __ballot_sync(gmask, hit) != 0u;
      }
      if (found && lane8 == 0) { level[v] = (unsigned char)(cur + 1); any = true; }
    }
  }
  if (any) *changed = 1u;
}

__global__ void __launch_bounds__(256) encode_kernel(const long long* __restrict__ row_ptr, const int* __restrict__ col,
                                                      const int* __restrict__ order, const long long* __restrict__ row_off,
                                                      const unsigned int* __restrict__ word_of, int M, unsigned int* __restrict__ enc) {
  const int lane = threadIdx.x & 31;
  const long long stride = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < M; p += stride) {
    const int v = order[p];
    const long long rb = row_ptr[v], d = row_ptr[v + 1] - rb, o = row_off[p];
    for (long long i = lane; i < d; i += 32) enc[o + i] = word_of[col[rb + i]];
  }
}

// need[q] |= 1 << owner[p] for every stored successor q of position p: the ranks that read q's basket (multi-GPU pushes)
__global__ void __launch_bounds__(256) need_mask_kernel(const long long* __restrict__ row_off, const unsigned int* __restrict__ enc,
                                                         const unsigned char* __restrict__ owner, int M, unsigned int* need32) {
  const int lane = threadIdx.x & 31;
  const long long stride = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < M; p += stride) {
    const unsigned int bit = 1u << owner[p];
    const long long rb = row_off[p], re = row_off[p + 1];
    for (long long i = rb + lane; i < re; i += 32) {
      const unsigned int w = enc[i];
      if (w & 0x80000000u) continue;  // a sink: no basket
      const unsigned int q = w & 0x3fffffffu;
      const unsigned int m = bit << ((q & 3u) * 8u);
      if (!(need32[q >> 2] & m)) atomicOr(&need32[q >> 2], m);
    }
  }
}

}  // namespace pprb200
