// Microbenchmark: ceilings of the random-access primitives of the tracking kernel
// on one B200 (build: nvcc -arch=sm_100a -O3 -o random_access random_access.cu).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
struct __align__(32) V8 { uint32_t w[8]; };
__device__ __forceinline__ uint64_t mix(uint64_t x){ x^=x>>32; x*=0xD6E8FEB86659FD93ull; x^=x>>32; x*=0xD6E8FEB86659FD93ull; x^=x>>32; return x; }
__device__ __forceinline__ V8 ld32(const void* a){ V8 v; asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v.w[0]),"=r"(v.w[1]),"=r"(v.w[2]),"=r"(v.w[3]),"=r"(v.w[4]),"=r"(v.w[5]),"=r"(v.w[6]),"=r"(v.w[7]) : "l"(a)); return v; }
// mode 0: random 32B gather; 1: random atomicAdd (4B, return used); 2: random 4B store;
// 3: dependent pair (32B gather -> address of a second 32B gather); 4: streaming 32B load+store
// window: addresses of element i are confined to [w0, w0+window) with w0 following i (region locality)
__global__ void __launch_bounds__(1024,1) k(int mode, const V8* __restrict__ src, V8* __restrict__ dst, uint32_t* __restrict__ cnt,
   int64_t n, int64_t n_elems, int64_t window, uint32_t* __restrict__ sink, int unroll) {
  uint32_t acc = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int64_t w0 = (i / window) * window; if (w0 + window > n_elems) w0 = n_elems - window;
    const int64_t j = w0 + (int64_t)(mix((uint64_t)i) % (uint64_t)window);
    if (mode == 0) { V8 v = ld32(src + j); acc += v.w[0] ^ v.w[7]; }
    else if (mode == 1) { acc += atomicAdd(cnt + j, 1u); }
    else if (mode == 2) { cnt[j] = (uint32_t)i; }
    else if (mode == 3) { V8 v = ld32(src + j); int64_t j2 = w0 + (int64_t)(mix((uint64_t)(v.w[0] + i)) % (uint64_t)window); V8 u = ld32(src + j2); acc += u.w[3]; }
    else { V8 v = ld32(src + i); v.w[0] += 1; dst[i] = v; }
  }
  if (acc == 0x12345678u) sink[0] = acc;
}
int main(){
  const int64_t n = 13500000;            // particles per snapshot in the bench
  const int64_t n_elems = n;             // 432 MB of 32 B records / 54 MB of counters
  V8 *src, *dst; uint32_t *cnt, *sink;
  cudaMalloc(&src, n_elems * 32); cudaMalloc(&dst, n_elems * 32); cudaMalloc(&cnt, n_elems * 4); cudaMalloc(&sink, 4);
  cudaMemset(src, 1, n_elems * 32); cudaMemset(cnt, 0, n_elems * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[] = {"gather32B", "atomicAdd4B", "store4B", "dependent2x32B", "stream32B_rw"};
  int64_t windows[] = {13500, 150000, 1000000, n};
  for (int mode = 0; mode < 5; ++mode) for (int wi = 0; wi < 4; ++wi) {
    if (mode == 4 && wi > 0) continue;
    float best = 1e9;
    for (int rep = 0; rep < 5; ++rep) {
      cudaMemset(dst, rep, 64 << 20);    // disturb L2 a little
      cudaEventRecord(e0);
      k<<<148, 1024>>>(mode, src, dst, cnt, n, n_elems, windows[wi], sink, 1);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("%-16s window %9lld : %.3f ms  (%.1f G ops/s)\n", names[mode], (long long)windows[wi], best, n / best / 1e6);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
