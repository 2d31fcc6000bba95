// Microbenchmark: floors of the primitive operations the binned counting path is made of, on B200.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2ops l2ops.cu
//   run  : ./l2ops [table_MiB]         (16 = L2 resident slice, 8192 = HBM sized)
// Every variant performs N "key operations"; reported as G ops/s and cycles per op per SM at 1.965 GHz x 148 SMs.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("err %s line %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
__device__ __forceinline__ uint64_t mix(uint64_t x){ x^=x>>31; x*=0x9E3779B97F4A7C15ULL; x^=x>>29; x*=0xBF58476D1CE4E5B9ULL; x^=x>>32; return x; }
__device__ __forceinline__ void ld256(const uint64_t* p, uint64_t w[4]){ asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(w[0]),"=l"(w[1]),"=l"(w[2]),"=l"(w[3]):"l"(p)); }

// keys: precomputed random bucket indices (coalesced 8-byte loads, like the bin stream)
template<int MODE>
__global__ void __launch_bounds__(256) k(uint64_t* t, uint64_t nb_mask, const uint64_t* __restrict__ keys, uint64_t n, uint64_t* sink){
  uint64_t acc=0;
  const uint64_t stride=(uint64_t)gridDim.x*1024;
  for(uint64_t base=(uint64_t)blockIdx.x*1024; base<n; base+=stride){
    uint64_t h[4]; uint64_t w[4][4];
    #pragma unroll
    for(int u=0;u<4;u++){ uint64_t i=base+u*256+threadIdx.x; h[u]= i<n ? __ldcs(keys+i) : 0; }
    if(MODE==0){ // probe only, 4 in flight
      #pragma unroll
      for(int u=0;u<4;u++) ld256(t+(h[u]&nb_mask)*4, w[u]);
      #pragma unroll
      for(int u=0;u<4;u++) acc+=w[u][0]^w[u][1]^w[u][2]^w[u][3];
    }
    if(MODE==1){ // RED only
      #pragma unroll
      for(int u=0;u<4;u++) atomicAdd((unsigned long long*)(t+(h[u]&nb_mask)*4+((h[u]>>40)&3)), 1ULL);
    }
    if(MODE==2){ // probe then RED into the probed bucket, one key after the other (dependent)
      #pragma unroll
      for(int u=0;u<4;u++){ ld256(t+(h[u]&nb_mask)*4, w[u]); uint64_t s=(w[u][0]^w[u][1]^w[u][2]^w[u][3]^(h[u]>>40))&3; atomicAdd((unsigned long long*)(t+(h[u]&nb_mask)*4+s), 1ULL); }
    }
    if(MODE==3){ // 4 probes in flight, then 4 REDs (k_bin_count's shape)
      #pragma unroll
      for(int u=0;u<4;u++) ld256(t+(h[u]&nb_mask)*4, w[u]);
      #pragma unroll
      for(int u=0;u<4;u++){ uint64_t s=(w[u][0]^w[u][1]^w[u][2]^w[u][3]^(h[u]>>40))&3; atomicAdd((unsigned long long*)(t+(h[u]&nb_mask)*4+s), 1ULL); }
    }
    if(MODE==4){ // ATOM with return only
      #pragma unroll
      for(int u=0;u<4;u++) acc+=atomicAdd((unsigned long long*)(t+(h[u]&nb_mask)*4+((h[u]>>40)&3)), 1ULL);
    }
    if(MODE==5){ // CAS only
      #pragma unroll
      for(int u=0;u<4;u++) acc+=atomicCAS((unsigned long long*)(t+(h[u]&nb_mask)*4+((h[u]>>40)&3)), 0ULL, h[u]|1);
    }
    if(MODE==6){ // plain 8-byte store, divergent
      #pragma unroll
      for(int u=0;u<4;u++) t[(h[u]&nb_mask)*4+((h[u]>>40)&3)] = h[u];
    }
    if(MODE==7){ // probe + plain store (no atomic)
      #pragma unroll
      for(int u=0;u<4;u++) ld256(t+(h[u]&nb_mask)*4, w[u]);
      #pragma unroll
      for(int u=0;u<4;u++){ uint64_t s=(w[u][0]^w[u][1]^w[u][2]^w[u][3]^(h[u]>>40))&3; t[(h[u]&nb_mask)*4+s]=w[u][0]+1; }
    }
    if(MODE==8){ // 8-byte probe (one slot) + RED
      uint64_t v[4];
      #pragma unroll
      for(int u=0;u<4;u++) v[u]=__ldcg(t+(h[u]&nb_mask)*4+((h[u]>>40)&3));
      #pragma unroll
      for(int u=0;u<4;u++) atomicAdd((unsigned long long*)(t+(h[u]&nb_mask)*4+((h[u]>>40)&3)), 1ULL+(v[u]&0));
    }
    if(MODE==10){ // 16 B probe + RED
      uint64_t v[4][2];
      #pragma unroll
      for(int u=0;u<4;u++){ const uint64_t* p=t+(h[u]&nb_mask)*4+((h[u]>>40)&2); asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];":"=l"(v[u][0]),"=l"(v[u][1]):"l"(p)); }
      #pragma unroll
      for(int u=0;u<4;u++) atomicAdd((unsigned long long*)(t+(h[u]&nb_mask)*4+((h[u]>>40)&2)+((v[u][0]^v[u][1])&1)), 1ULL);
    }
    if(MODE==11){ // two 8 B probes (neighbouring slots) + RED
      uint64_t v[4][2];
      #pragma unroll
      for(int u=0;u<4;u++){ const uint64_t* p=t+(h[u]&nb_mask)*4+((h[u]>>40)&2); v[u][0]=__ldcg(p); v[u][1]=__ldcg(p+1); }
      #pragma unroll
      for(int u=0;u<4;u++) atomicAdd((unsigned long long*)(t+(h[u]&nb_mask)*4+((h[u]>>40)&2)+((v[u][0]^v[u][1])&1)), 1ULL);
    }
    if(MODE==12){ // 8 B probe + CAS
      uint64_t v[4];
      #pragma unroll
      for(int u=0;u<4;u++) v[u]=__ldcg(t+(h[u]&nb_mask)*4+((h[u]>>40)&3));
      #pragma unroll
      for(int u=0;u<4;u++) acc+=atomicCAS((unsigned long long*)(t+(h[u]&nb_mask)*4+((h[u]>>40)&3)), v[u], v[u]+1);
    }
    if(MODE==13){ // 32 B probe + RED to an unrelated bucket
      #pragma unroll
      for(int u=0;u<4;u++) ld256(t+(h[u]&nb_mask)*4, w[u]);
      #pragma unroll
      for(int u=0;u<4;u++){ uint64_t s=(w[u][0]^w[u][1]^w[u][2]^w[u][3]^(h[u]>>40))&3; atomicAdd((unsigned long long*)(t+((h[u]>>20)&nb_mask)*4+s), 1ULL); }
    }
    if(MODE==14){ // 32 B probe as two 16 B loads + RED
      #pragma unroll
      for(int u=0;u<4;u++){ const uint64_t* p=t+(h[u]&nb_mask)*4; asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];":"=l"(w[u][0]),"=l"(w[u][1]):"l"(p)); asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];":"=l"(w[u][2]),"=l"(w[u][3]):"l"(p+2)); }
      #pragma unroll
      for(int u=0;u<4;u++){ uint64_t s=(w[u][0]^w[u][1]^w[u][2]^w[u][3]^(h[u]>>40))&3; atomicAdd((unsigned long long*)(t+(h[u]&nb_mask)*4+s), 1ULL); }
    }
    if(MODE==15){ // 32 B probe with L1 allocation (ld.ca) + RED
      #pragma unroll
      for(int u=0;u<4;u++){ const uint64_t* p=t+(h[u]&nb_mask)*4; asm volatile("ld.global.ca.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(w[u][0]),"=l"(w[u][1]),"=l"(w[u][2]),"=l"(w[u][3]):"l"(p)); }
      #pragma unroll
      for(int u=0;u<4;u++){ uint64_t s=(w[u][0]^w[u][1]^w[u][2]^w[u][3]^(h[u]>>40))&3; atomicAdd((unsigned long long*)(t+(h[u]&nb_mask)*4+s), 1ULL); }
    }
    if(MODE==9){ // 32-bit RED (u32 counters)
      #pragma unroll
      for(int u=0;u<4;u++) atomicAdd((unsigned int*)(t+(h[u]&nb_mask)*4)+((h[u]>>40)&7), 1u);
    }
  }
  if(acc==0x1234567) *sink=acc;
}

// shared-memory ranking as in k_bin_scatter: 16 keys per thread, P bins
template<int MODE>
__global__ void __launch_bounds__(256) ks(const uint64_t* __restrict__ keys, uint64_t n, int pbits, uint64_t* out, unsigned long long* cursors, uint64_t cap, uint64_t* sink){
  extern __shared__ uint32_t s_hist[];
  const uint32_t P=1u<<pbits; uint64_t acc=0;
  unsigned long long* s_base=(unsigned long long*)(s_hist+P);
  for(uint64_t t0=(uint64_t)blockIdx.x*4096; t0<n; t0+=(uint64_t)gridDim.x*4096){
    for(uint32_t i=threadIdx.x;i<P;i+=256) s_hist[i]=0;
    __syncthreads();
    uint64_t h[16]; uint32_t rk[16];
    #pragma unroll
    for(int u=0;u<16;u++){ uint64_t i=t0+u*256+threadIdx.x; h[u]= i<n ? __ldcs(keys+i) : 0; }
    if(MODE==0 || MODE>=2){
      #pragma unroll
      for(int u=0;u<16;u++) rk[u]=atomicAdd(&s_hist[(h[u]>>(64-pbits))&(P-1)],1u);
    } else {
      #pragma unroll
      for(int u=0;u<16;u++) rk[u]=u;
    }
    __syncthreads();
    if(MODE==2){
      for(uint32_t i=threadIdx.x;i<P;i+=256){ uint32_t c=s_hist[i]; s_base[i]= c? atomicAdd(&cursors[i],(unsigned long long)c):0ULL; }
      __syncthreads();
      #pragma unroll
      for(int u=0;u<16;u++){ uint32_t b=(h[u]>>(64-pbits))&(P-1); uint64_t pos=(s_base[b]+rk[u]); if(pos<cap) out[b*cap+pos]=h[u]; }
    }
    if(MODE==3){ // full, but the output wraps inside a small (L2 resident) window per bin
      for(uint32_t i=threadIdx.x;i<P;i+=256){ uint32_t c=s_hist[i]; s_base[i]= c? atomicAdd(&cursors[i],(unsigned long long)c):0ULL; }
      __syncthreads();
      const uint64_t win = (cap < 4096 ? cap : 4096);
      #pragma unroll
      for(int u=0;u<16;u++){ uint32_t b=(h[u]>>(64-pbits))&(P-1); uint64_t pos=(s_base[b]+rk[u]) % win; out[b*cap+pos]=h[u]; }
    }
    if(MODE==4){ // full-sector stores: every 4th key of a thread writes an aligned 32-byte sector at its bin's frontier
      for(uint32_t i=threadIdx.x;i<P;i+=256){ uint32_t c=s_hist[i]; s_base[i]= c? atomicAdd(&cursors[i],(unsigned long long)((c+3)&~3u)):0ULL; }
      __syncthreads();
      #pragma unroll
      for(int u=0;u<16;u+=4){ uint32_t b=(h[u]>>(64-pbits))&(P-1); uint64_t pos=(s_base[b]+(rk[u]&~3u)) ; if(pos+4<=cap){ uint64_t* q=out+b*cap+pos; asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};"::"l"(q),"l"(h[u]),"l"(h[u+1]),"l"(h[u+2]),"l"(h[u+3]):"memory"); } }
    }
    if(MODE==1){ // divergent store only: position = hashed pseudo-cursor (no ranking)
      #pragma unroll
      for(int u=0;u<16;u++){ uint32_t b=(h[u]>>(64-pbits))&(P-1); uint64_t pos=((t0>>12)*2+ (rk[u]&1)) % cap; out[b*cap+pos]=h[u]; }
    }
    #pragma unroll
    for(int u=0;u<16;u++) acc+=rk[u];
    __syncthreads();
  }
  if(acc==0x1234567) *sink=acc;
}

__global__ void gen(uint64_t* keys, uint64_t n, uint64_t seed){ for(uint64_t i=blockIdx.x*(uint64_t)blockDim.x+threadIdx.x;i<n;i+=(uint64_t)gridDim.x*blockDim.x) keys[i]=mix(i+seed); }

int main(int argc,char**argv){
  double mib = argc>1? atof(argv[1]) : 16;
  uint64_t nb=1; while((nb*2)*32 <= (uint64_t)(mib*(1ull<<20))) nb*=2;
  uint64_t* t; CK(cudaMalloc(&t, nb*32)); CK(cudaMemset(t,0,nb*32)); uint64_t* sink; CK(cudaMalloc(&sink,8));
  const uint64_t n = 1ull<<28; uint64_t* keys; CK(cudaMalloc(&keys,n*8)); gen<<<148*8,256>>>(keys,n,12345); CK(cudaDeviceSynchronize());
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  printf("table %.1f MiB (%llu buckets), %llu keys per launch\n", nb*32.0/(1<<20), (unsigned long long)nb, (unsigned long long)n);
  #define RUN(M,name,blocks) { k<M><<<blocks,256>>>(t,nb-1,keys,n,sink); CK(cudaDeviceSynchronize()); cudaEventRecord(a); k<M><<<blocks,256>>>(t,nb-1,keys,n,sink); cudaEventRecord(b); CK(cudaDeviceSynchronize()); float ms; cudaEventElapsedTime(&ms,a,b); printf("%-44s grid %5d %8.3f ms %7.1f G/s %6.2f cyc/op/SM\n", name, blocks, ms, n/ms/1e6, ms*1e-3*148*1.965e9/n); }
  for(int per=8; per<=8; per+=4){
    RUN(0,"probe LDG.256 x4",148*per); RUN(1,"RED.64",148*per); RUN(2,"probe->RED dependent",148*per); RUN(3,"4 probes then 4 REDs",148*per);
    RUN(4,"ATOM.64 with return",148*per); RUN(5,"CAS.64",148*per); RUN(6,"plain 8 B store",148*per); RUN(7,"4 probes then 4 stores",148*per); RUN(8,"8 B probe + RED",148*per); RUN(9,"RED.32",148*per);
    RUN(10,"16 B probe + RED",148*per); RUN(11,"2 x 8 B probe + RED",148*per); RUN(12,"8 B probe + CAS",148*per); RUN(13,"32 B probe + RED elsewhere",148*per); RUN(14,"2 x 16 B probe + RED",148*per); RUN(15,"32 B probe (ld.ca) + RED",148*per);
  }
  // scatter pieces
  uint64_t cap = (n >> 6); uint64_t* out; CK(cudaMalloc(&out, (n + (n>>2))*8)); unsigned long long* cur; CK(cudaMalloc(&cur, 8192*8));
  #define RUNS(M,name,pb) { size_t sm=(size_t)(1u<<pb)*12; uint64_t capb=(n+(n>>3))>>pb; CK(cudaMemset(cur,0,8192*8)); ks<M><<<148*3,256,sm>>>(keys,n,pb,out,cur,capb,sink); CK(cudaDeviceSynchronize()); CK(cudaMemset(cur,0,8192*8)); cudaEventRecord(a); ks<M><<<148*3,256,sm>>>(keys,n,pb,out,cur,capb,sink); cudaEventRecord(b); CK(cudaDeviceSynchronize()); float ms; cudaEventElapsedTime(&ms,a,b); printf("%-44s P=2^%-2d %8.3f ms %7.1f G/s %6.2f cyc/op/SM\n", name, pb, ms, n/ms/1e6, ms*1e-3*148*1.965e9/n); }
  (void)cap;
  for(int pb=6; pb<=12; pb+=3){ RUNS(0,"scatter: load + smem rank only",pb); RUNS(2,"scatter: full (rank, reserve, store)",pb); RUNS(3,"scatter: full, output window L2 resident",pb); RUNS(4,"scatter: one aligned 32 B store per 4 keys",pb); }
  return 0;
}
