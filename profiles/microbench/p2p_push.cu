// p2p_push.cu -- how fast can the SMs of one B200 push 1 200-byte basket slots into a peer's memory over NVLink?
// (publish_slot is issued by the warp / CTA that produced the slot.) Variants:
//   st16    the publish_slot pattern: lanes copy the slot with 16-byte loads (local) and 16-byte stores (peer)
//   st16x3  the same with the slot's three rounds of loads issued before the stores
//   bulk    the slot staged in shared memory, one lane issues cp.async.bulk shared -> peer global (TMA engine)
// each with 14 and 32 warps per SM; slots are visited in a scattered order like LPT-sharded positions.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o p2p_push p2p_push.cu ; run on a box with >= 2 GPUs
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
constexpr int SLOT = 1200, N16 = SLOT / 16;

__device__ __forceinline__ size_t slot_of(unsigned i, unsigned nslots) { return (size_t)((i * 2654435761u) % nslots); }

template <int MODE>
__global__ void push_kernel(const unsigned char* src, unsigned char* dst0, unsigned char* dst1, int ndst, unsigned nslots, unsigned* counter) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned char* stage = smem + (size_t)w * 1280;
  for (;;) {
    unsigned i = 0;
    if (lane == 0) i = atomicAdd(counter, 1u);
    i = __shfl_sync(0xffffffffu, i, 0);
    if (i >= nslots) break;
    const size_t off = slot_of(i, nslots) * SLOT;
    const int4* s = reinterpret_cast<const int4*>(src + off);
    if (MODE == 0) {
      for (int k = lane; k < N16; k += 32) {
        const int4 v = __ldcg(s + k);
        __stcg(reinterpret_cast<int4*>(dst0 + off) + k, v);
        if (ndst > 1) __stcg(reinterpret_cast<int4*>(dst1 + off) + k, v);
      }
    } else if (MODE == 1) {
      int4 v[3];
#pragma unroll
      for (int r = 0; r < 3; r++) if (lane + 32 * r < N16) v[r] = __ldcg(s + lane + 32 * r);
#pragma unroll
      for (int r = 0; r < 3; r++) if (lane + 32 * r < N16) {
        __stcg(reinterpret_cast<int4*>(dst0 + off) + lane + 32 * r, v[r]);
        if (ndst > 1) __stcg(reinterpret_cast<int4*>(dst1 + off) + lane + 32 * r, v[r]);
      }
    } else {
      // wait until the previous bulk copy has finished READING the staging buffer, refill it, make it visible to the async proxy
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
      for (int k = lane; k < N16; k += 32) reinterpret_cast<int4*>(stage)[k] = __ldcg(s + k);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        const unsigned sa = (unsigned)__cvta_generic_to_shared(stage);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst0 + off), "r"(sa), "r"(SLOT) : "memory");
        if (ndst > 1) asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst1 + off), "r"(sa), "r"(SLOT) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
  }
  if (MODE == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int MODE>
static void run(const char* name, int warps, const unsigned char* src, unsigned char* d0, unsigned char* d1, int ndst, unsigned nslots, unsigned* counter, int sm) {
  const size_t smem = MODE == 2 ? (size_t)warps * 1280 : 0;
  if (smem > 48 * 1024) CK(cudaFuncSetAttribute(push_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaMemset(counter, 0, 4));
    CK(cudaEventRecord(a));
    push_kernel<MODE><<<sm, warps * 32, smem>>>(src, d0, d1, ndst, nslots, counter);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (rep && ms < best) best = ms;
  }
  printf("%-7s %2d warps/SM, %d destination(s): %8.3f ms  %7.1f GB/s pushed\n", name, warps, ndst, best, (double)nslots * SLOT * ndst / best / 1e6);
}

int main() {
  int cnt = 0; CK(cudaGetDeviceCount(&cnt));
  if (cnt < 2) { printf("needs 2 GPUs\n"); return 0; }
  const unsigned nslots = 2000000;  // 2.4 GB
  unsigned char *src, *dst_local, *dst_peer; unsigned* counter;
  CK(cudaSetDevice(1)); CK(cudaMalloc(&dst_peer, (size_t)nslots * SLOT)); CK(cudaMemset(dst_peer, 0, (size_t)nslots * SLOT));
  CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0));
  CK(cudaMalloc(&src, (size_t)nslots * SLOT)); CK(cudaMemset(src, 1, (size_t)nslots * SLOT));
  CK(cudaMalloc(&dst_local, (size_t)nslots * SLOT)); CK(cudaMalloc(&counter, 4));
  int sm; CK(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0));
  printf("local destination (HBM, for scale):\n");
  for (int warps : {14, 32}) { run<0>("st16", warps, src, dst_local, nullptr, 1, nslots, counter, sm); run<2>("bulk", warps, src, dst_local, nullptr, 1, nslots, counter, sm); }
  printf("peer destination (NVLink):\n");
  for (int warps : {14, 32}) {
    run<0>("st16", warps, src, dst_peer, nullptr, 1, nslots, counter, sm);
    run<1>("st16x3", warps, src, dst_peer, nullptr, 1, nslots, counter, sm);
    run<2>("bulk", warps, src, dst_peer, nullptr, 1, nslots, counter, sm);
  }
  // verify the bulk variant delivered the bytes
  unsigned char h[16]; CK(cudaMemcpy(h, dst_peer + (size_t)12345 * SLOT + 1184, 16, cudaMemcpyDeviceToHost));
  printf("check byte %d (want 1)\n", h[15]);
  return 0;
}
