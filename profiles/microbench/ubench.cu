// profiles/microbench/ubench.cu -- design-time microbenchmarks (B200): which primitive can carry the merge?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o ubench ubench.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__);exit(1);} }while(0)

__device__ __forceinline__ uint32_t hash32(uint32_t x){ x^=x>>16; x*=0x7feb352dU; x^=x>>15; x*=0x846ca68bU; x^=x>>16; return x; }

constexpr int TAB = 8192;  // entries
// mode 0: ATOMS.ADD.32 ; 1: 2x ATOMS.32 with carry (fixed-point 64) ; 2: atomicAdd u64 (CAS spin) ; 3: plain LDS.64/DADD/STS.64 ; 4: LDS.32 key probe + plain RMW
template<int MODE>
__global__ void smem_rmw(uint64_t* out, int iters){
  extern __shared__ unsigned char smraw[];
  uint32_t* t32=(uint32_t*)smraw; unsigned long long* t64=(unsigned long long*)smraw; double* td=(double*)smraw;
  for(int i=threadIdx.x;i<TAB*2;i+=blockDim.x) t32[i]=0;
  __syncthreads();
  uint32_t x = hash32(blockIdx.x*blockDim.x+threadIdx.x+1);
  for(int it=0; it<iters; it++){
    #pragma unroll 4
    for(int u=0;u<4;u++){
      x = x*1664525u+1013904223u;
      uint32_t h = (x>>8)&(TAB-1);
      if(MODE==0){ atomicAdd(&t32[h], x); }
      else if(MODE==1){ uint32_t lo=x, hi=x&0xff; uint32_t old=atomicAdd(&t32[2*h],lo); uint32_t carry=(old+lo)<old; atomicAdd(&t32[2*h+1],hi+carry); }
      else if(MODE==2){ atomicAdd(&t64[h], (unsigned long long)x); }
      else if(MODE==3){ double v=td[h]; v=fma((double)x,1e-9,v); td[h]=v; }
    }
    if(MODE==3) __syncwarp();
  }
  __syncthreads();
  if(threadIdx.x==0) out[blockIdx.x]=t64[0]+t64[17];
}

// global REDG.ADD.64 over a table of `n` u64 (random addresses)
__global__ void gred(unsigned long long* t, uint32_t mask, int iters){
  uint32_t x = hash32(blockIdx.x*blockDim.x+threadIdx.x+1);
  for(int it=0; it<iters; it++){ x=x*1664525u+1013904223u; atomicAdd(&t[(x>>4)&mask], (unsigned long long)x); }
}
// global hash-style: CAS on key then RED on value
__global__ void gcasred(int* keys, unsigned long long* vals, uint32_t mask, int iters){
  uint32_t x = hash32(blockIdx.x*blockDim.x+threadIdx.x+1);
  for(int it=0; it<iters; it++){ x=x*1664525u+1013904223u; uint32_t h=(x>>4)&mask; int k=(int)h;
    int cur=keys[h]; if(cur!=k){ atomicCAS(&keys[h],-1,k);} atomicAdd(&vals[h],(unsigned long long)x); }
}

// random 1200-byte slot gathers: each warp reads `per_warp` random slots with 3x LDG.128 per lane (25 lanes), DEPTH slots in flight
template<int DEPTH>
__global__ void gather_ldg(const int4* base, uint32_t nslots, int per_warp, uint64_t* out){
  int warp=(blockIdx.x*blockDim.x+threadIdx.x)>>5, lane=threadIdx.x&31;
  uint32_t x=hash32(warp+1); uint64_t acc=0;
  for(int i=0;i<per_warp;i+=DEPTH){
    int4 a[DEPTH],b[DEPTH],c[DEPTH];
    #pragma unroll
    for(int d=0;d<DEPTH;d++){ x=x*1664525u+1013904223u; uint32_t s=(uint32_t)(((uint64_t)x*nslots)>>32); const int4* p=base+(size_t)s*75;
      if(lane<25){ a[d]=__ldg(p+lane); b[d]=__ldg(p+25+lane); c[d]=__ldg(p+50+lane);} else {a[d]=b[d]=c[d]=make_int4(0,0,0,0);} }
    #pragma unroll
    for(int d=0;d<DEPTH;d++) acc+= (unsigned)(a[d].x^b[d].y^c[d].z);
  }
  if(acc==0x1234567) out[0]=acc;
}
// same via cp.async.bulk (TMA 1D bulk copy) into a smem ring, mbarrier completion
template<int STAGES>
__global__ void gather_bulk(const int4* base, uint32_t nslots, int per_warp, uint64_t* out){
  extern __shared__ __align__(128) unsigned char sm[];
  int wib=threadIdx.x>>5, lane=threadIdx.x&31; int nw=blockDim.x>>5;
  unsigned char* ring = sm + (size_t)wib*STAGES*1216;
  uint64_t* bars = (uint64_t*)(sm + (size_t)nw*STAGES*1216) + wib*STAGES;
  uint32_t x=hash32(blockIdx.x*nw+wib+1); uint64_t acc=0;
  if(lane==0){ for(int s=0;s<STAGES;s++){ uint32_t ba=(uint32_t)__cvta_generic_to_shared(&bars[s]); asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;"::"r"(ba)); } }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  auto issue=[&](int s){ x=x*1664525u+1013904223u; uint32_t sl=(uint32_t)(((uint64_t)x*nslots)>>32); const int4* p=base+(size_t)sl*75;
     uint32_t ba=(uint32_t)__cvta_generic_to_shared(&bars[s]); uint32_t da=(uint32_t)__cvta_generic_to_shared(ring+s*1216);
     asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"::"r"(ba),"r"(1200):"memory");
     asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"::"r"(da),"l"(p),"r"(1200),"r"(ba):"memory"); };
  if(lane==0) for(int s=0;s<STAGES;s++) issue(s);
  for(int i=0;i<per_warp;i++){
    int s=i%STAGES; uint32_t ph=(i/STAGES)&1; uint32_t ba=(uint32_t)__cvta_generic_to_shared(&bars[s]);
    asm volatile("{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT;\nDONE:\n}"::"r"(ba),"r"(ph):"memory");
    const int4* q=(const int4*)(ring+s*1216);
    if(lane<25){ int4 a=q[lane],b=q[25+lane],c=q[50+lane]; acc+=(unsigned)(a.x^b.y^c.z);} 
    __syncwarp();
    if(lane==0 && i+STAGES<per_warp) issue(s);
  }
  if(acc==0x1234567) out[0]=acc;
}

template<class F> float timeit(F f,int reps=3){ cudaEvent_t a,b; cudaEventCreate(&a);cudaEventCreate(&b); f(); CK(cudaDeviceSynchronize()); float best=1e30f; for(int r=0;r<reps;r++){ cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms,a,b); if(ms<best)best=ms;} return best; }

int main(){
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr,0)); int sms=pr.multiProcessorCount; double ghz=pr.clockRate*1e-6;
  printf("device %s sms %d clock %.3f GHz\n",pr.name,sms,ghz);
  uint64_t* out; CK(cudaMalloc(&out,1<<20));
  // ---- smem RMW
  const char* names[]={"ATOMS.ADD.32","2xATOMS.32 carry","atomicAdd u64 (CAS spin)","plain LDS64/DFMA/STS64"};
  for(int warps: {4,8,16,32}){
    int iters=2000; int threads=warps*32; size_t smem=TAB*8;
    auto run=[&](int mode){ float ms=0; 
      if(mode==0){ CK(cudaFuncSetAttribute(smem_rmw<0>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem)); ms=timeit([&]{smem_rmw<0><<<sms,threads,smem>>>(out,iters);}); }
      if(mode==1){ CK(cudaFuncSetAttribute(smem_rmw<1>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem)); ms=timeit([&]{smem_rmw<1><<<sms,threads,smem>>>(out,iters);}); }
      if(mode==2){ CK(cudaFuncSetAttribute(smem_rmw<2>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem)); ms=timeit([&]{smem_rmw<2><<<sms,threads,smem>>>(out,iters);}); }
      if(mode==3){ CK(cudaFuncSetAttribute(smem_rmw<3>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem)); ms=timeit([&]{smem_rmw<3><<<sms,threads,smem>>>(out,iters);}); }
      double ops=(double)sms*threads*iters*4; printf("smem %-26s warps/SM %2d: %.1f Gops/s  = %.3f ops/clk/SM (at %.2f GHz nominal)\n",names[mode],warps,ops/ms*1e-6,ops/ms*1e-6/sms/ghz,ghz); };
    for(int m=0;m<4;m++) run(m);
  }
  // ---- global RED
  for(size_t mb: {32,256,4096}){
    size_t n=mb*1024*1024/8; unsigned long long* t; CK(cudaMalloc(&t,n*8)); CK(cudaMemset(t,0,n*8));
    int iters=256; int blocks=sms*8, threads=256;
    float ms=timeit([&]{gred<<<blocks,threads>>>(t,(uint32_t)(n-1),iters);});
    double ops=(double)blocks*threads*iters; printf("global REDG.ADD.64 table %5zu MB: %.1f Gops/s\n",mb,ops/ms*1e-6);
    int* keys; CK(cudaMalloc(&keys,n*4)); CK(cudaMemset(keys,0xff,n*4));
    ms=timeit([&]{gcasred<<<blocks,threads>>>(keys,t,(uint32_t)(n-1),iters);});
    printf("global key-check+RED     table %5zu MB: %.1f Gops/s\n",mb,ops/ms*1e-6);
    cudaFree(t); cudaFree(keys);
  }
  // ---- random 1200B gathers from a 6 GB array
  { uint32_t nslots=5000000; int4* base; CK(cudaMalloc(&base,(size_t)nslots*1200)); CK(cudaMemset(base,1,(size_t)nslots*1200));
    for(int wps: {8,16,32,64}){
      int per_warp=512; int threads=256; int blocks=sms*wps/8;
      auto rep=[&](const char* nm,float ms){ double bytes=(double)blocks*(threads/32)*per_warp*1200; printf("gather %-22s warps/SM %2d: %.0f GB/s\n",nm,wps,bytes/ms*1e-6); };
      rep("LDG.128 depth1",timeit([&]{gather_ldg<1><<<blocks,threads>>>(base,nslots,per_warp,out);}));
      rep("LDG.128 depth2",timeit([&]{gather_ldg<2><<<blocks,threads>>>(base,nslots,per_warp,out);}));
      rep("LDG.128 depth4",timeit([&]{gather_ldg<4><<<blocks,threads>>>(base,nslots,per_warp,out);}));
      if(wps<=32){
        { size_t sm=(size_t)8*2*1216+8*2*8; CK(cudaFuncSetAttribute(gather_bulk<2>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)sm)); rep("bulk(TMA) 2 stages",timeit([&]{gather_bulk<2><<<blocks,threads,sm>>>(base,nslots,per_warp,out);})); }
        { size_t sm=(size_t)8*4*1216+8*4*8; CK(cudaFuncSetAttribute(gather_bulk<4>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)sm)); rep("bulk(TMA) 4 stages",timeit([&]{gather_bulk<4><<<blocks,threads,sm>>>(base,nslots,per_warp,out);})); }
        { size_t sm=(size_t)8*8*1216+8*8*8; CK(cudaFuncSetAttribute(gather_bulk<8>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)sm)); rep("bulk(TMA) 8 stages",timeit([&]{gather_bulk<8><<<blocks,threads,sm>>>(base,nslots,per_warp,out);})); }
      }
    }
    cudaFree(base);
  }
  printf("done\n"); return 0;
}
