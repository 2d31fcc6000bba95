// Microbenchmark: what does one random 32-byte table probe cost in HBM traffic and time on B200?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o randaccess randaccess.cu
// run  : ./randaccess [table_GiB] [limit 0|32|64|128]   (under ncu: --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("err %s line %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
__device__ __forceinline__ uint64_t mix(uint64_t x){ x^=x>>31; x*=0x9E3779B97F4A7C15ULL; x^=x>>29; x*=0xBF58476D1CE4E5B9ULL; x^=x>>32; return x; }

template<int MODE>
__global__ void __launch_bounds__(256) k(uint64_t* t, uint64_t nb_mask, uint64_t per_thread, uint64_t* sink){
  uint64_t id = blockIdx.x*(uint64_t)blockDim.x+threadIdx.x; uint64_t acc=0;
  for(uint64_t i=0;i<per_thread;i++){
    uint64_t h = mix(id*per_thread+i+1);
    uint64_t* p = t + (h & nb_mask)*4;
    if(MODE==0){ uint64_t a,b,c,d; asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(a),"=l"(b),"=l"(c),"=l"(d):"l"(p)); acc+=a^b^c^d; }
    if(MODE==1){ uint64_t a,b,c,d; asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];":"=l"(a),"=l"(b):"l"(p)); asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];":"=l"(c),"=l"(d):"l"(p+2)); acc+=a^b^c^d; }
    if(MODE==2){ uint64_t a; asm volatile("ld.global.cg.u64 %0, [%1];":"=l"(a):"l"(p)); acc+=a; }
    if(MODE==3){ atomicAdd((unsigned long long*)p, 1ULL); }                               // RED only
    if(MODE==4){ uint64_t a,b,c,d; asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(a),"=l"(b),"=l"(c),"=l"(d):"l"(p)); acc+=a^b^c^d; atomicAdd((unsigned long long*)p+(acc&3), 1ULL); } // probe + RED
    if(MODE==5){ uint64_t a,b,c,d; asm volatile("ld.global.cg.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(a),"=l"(b),"=l"(c),"=l"(d):"l"(p)); acc+=a^b^c^d; }
    if(MODE==6){ uint64_t a,b,c,d; asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(a),"=l"(b),"=l"(c),"=l"(d):"l"(p)); acc+=a^b^c^d; }
    if(MODE==7){ uint64_t a,b,c,d; asm volatile("ld.relaxed.gpu.global.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(a),"=l"(b),"=l"(c),"=l"(d):"l"(p)); acc+=a^b^c^d; }
  }
  if(acc==0x1234567) *sink=acc;
}
int main(int argc,char**argv){
  double gib = argc>1? atof(argv[1]) : 8; int limit = argc>2? atoi(argv[2]) : 0;
  if(limit) CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, limit));
  size_t lim=0; cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity); printf("L2 fetch granularity limit = %zu\n", lim);
  uint64_t nb = 1; while((nb*2)*32 <= (uint64_t)(gib*(1ull<<30))) nb*=2;
  uint64_t* t; CK(cudaMalloc(&t, nb*32)); CK(cudaMemset(t,0,nb*32)); uint64_t* sink; CK(cudaMalloc(&sink,8));
  const uint64_t per=64; int blocks=148*8*4; uint64_t total = (uint64_t)blocks*256*per;
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  #define RUN(M,name) { k<M><<<blocks,256>>>(t,nb-1,per,sink); CK(cudaDeviceSynchronize()); cudaEventRecord(a); k<M><<<blocks,256>>>(t,nb-1,per,sink); cudaEventRecord(b); CK(cudaDeviceSynchronize()); float ms; cudaEventElapsedTime(&ms,a,b); printf("%-28s %8.3f ms  %7.2f G probes/s  (%.0f GB/s at 32 B/probe)\n", name, ms, total/ms/1e6, total*32.0/ms/1e6); }
  printf("table %.1f GiB, %llu probes per launch\n", nb*32.0/(1ull<<30), (unsigned long long)total);
  RUN(0,"ld.cg.v4.u64 (256-bit)"); RUN(1,"2 x ld.cg.v2.u64"); RUN(2,"ld.cg.u64 (8 B)"); RUN(3,"red.add.u64 only"); RUN(4,"256-bit probe + red"); RUN(5,"ld.cg.L2::64B.v4.u64"); RUN(6,"ld.nc.L1::no_allocate.v4"); RUN(7,"ld.relaxed.gpu.v4.u64");
  return 0;
}
