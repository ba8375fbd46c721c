// Microbenchmark of the mbarrier hand-off costs the CIPS / ray-SIREN pipelines are built from (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mbar_bench mbar_bench.cu && ./mbar_bench
// Prints SM clocks per iteration.  One CTA (the costs in question are SM-local); every variant runs 2000 iterations after 200 of warm-up.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* b) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p));
  return p != 0;
}
// wait flavours: 0 try_wait + 20 us hint, 1 try_wait (default limit), 2 test_wait spin, 3 test_wait + nanosleep(32) back-off
template <int W>
__device__ __forceinline__ bool probe(uint64_t* b, uint32_t par) {
  uint32_t ok;
  if (W == 0)
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(par), "r"(20000u) : "memory");
  else if (W == 1)
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
  else
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
  return ok != 0;
}
// LANES: 0 = every lane probes, 1 = lane 0 probes + __syncwarp
template <int W, int LANES>
__device__ __forceinline__ void wait(uint64_t* b, uint32_t par) {
  if (LANES == 0) {
    while (!probe<W>(b, par)) { if (W == 3) __nanosleep(32); }
  } else {
    if ((threadIdx.x & 31) == 0) while (!probe<W>(b, par)) { if (W == 3) __nanosleep(32); }
    __syncwarp();
  }
}

constexpr int kIters = 2000, kWarm = 200, kMaxD = 16;
struct Sm { uint64_t full[kMaxD], empty[kMaxD], a, b; };

// T1: one warp, arrive then wait on its own barrier (the cost of a satisfied wait + an arrive)
template <int W, int LANES, int COMMIT>
__global__ void t_self(long long* out) {
  __shared__ Sm s;
  if (threadIdx.x == 0) { mbar_init(&s.a, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (threadIdx.x >= 32) return;
  long long t0 = 0;
  for (int i = 0; i < kIters + kWarm; ++i) {
    if (i == kWarm) t0 = clock64();
    if (elect_one()) { if (COMMIT) tc_commit(&s.a); else mbar_arrive(&s.a); }
    __syncwarp();
    wait<W, LANES>(&s.a, i & 1);
  }
  if (threadIdx.x == 0) out[0] = (clock64() - t0) / kIters;
}
// T2: two warps ping-pong (round trip = two hand-offs)
template <int W, int LANES, int COMMIT>
__global__ void t_pingpong(long long* out) {
  __shared__ Sm s;
  if (threadIdx.x == 0) { mbar_init(&s.a, 1); mbar_init(&s.b, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  if (warp >= 2) return;
  long long t0 = 0;
  for (int i = 0; i < kIters + kWarm; ++i) {
    if (i == kWarm) t0 = clock64();
    if (warp == 0) {
      if (elect_one()) { if (COMMIT) tc_commit(&s.a); else mbar_arrive(&s.a); }
      __syncwarp();
      wait<W, LANES>(&s.b, i & 1);
    } else {
      wait<W, LANES>(&s.a, i & 1);
      if (elect_one()) { if (COMMIT) tc_commit(&s.b); else mbar_arrive(&s.b); }
      __syncwarp();
    }
  }
  if (threadIdx.x == 0) out[0] = (clock64() - t0) / kIters;
}
// T3: the CIPS weight ring without work: producer warp, two consumer warps that both observe every fill (empty count 2), D stages;
// POLLERS extra warps sit in a wait on a barrier that never completes until the end (the 16 epilogue warps)
template <int W, int LANES>
__global__ void t_ring(long long* out, int D, int pollers) {
  __shared__ Sm s;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxD; ++i) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], 2); }
    mbar_init(&s.a, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  long long t0 = 0;
  if (warp == 0) {
    uint32_t st = 0, ph = 0;
    for (int i = 0; i < kIters + kWarm; ++i) {
      wait<W, LANES>(&s.empty[st], ph ^ 1);
      if (elect_one()) mbar_arrive(&s.full[st]);
      __syncwarp();
      if (++st == (uint32_t)D) { st = 0; ph ^= 1; }
    }
  } else if (warp <= 2) {
    uint32_t st = 0, ph = 0;
    for (int i = 0; i < kIters + kWarm; ++i) {
      if (i == kWarm) t0 = clock64();
      wait<W, LANES>(&s.full[st], ph);
      if (elect_one()) mbar_arrive(&s.empty[st]);
      __syncwarp();
      if (++st == (uint32_t)D) { st = 0; ph ^= 1; }
    }
    if (threadIdx.x == 32) { out[0] = (clock64() - t0) / kIters; mbar_arrive(&s.a); }
  } else if (warp < 3 + pollers) {
    wait<W, LANES>(&s.a, 0);
  }
}

template <typename F> static long long run(F launch) {
  long long* d; cudaMalloc(&d, 8); cudaMemset(d, 0, 8);
  launch(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return -1; }
  long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); cudaFree(d);
  return h;
}
#define SELF(W, L, C) printf("self      wait=%d lanes=%s %s: %lld clk/iter\n", W, L ? "lane0" : "all  ", C ? "commit" : "arrive", run([](long long* d) { t_self<W, L, C><<<1, 64>>>(d); }))
#define PP(W, L, C) printf("pingpong  wait=%d lanes=%s %s: %lld clk/round trip (2 hand-offs)\n", W, L ? "lane0" : "all  ", C ? "commit" : "arrive", run([](long long* d) { t_pingpong<W, L, C><<<1, 64>>>(d); }))
#define RING(W, L, D, P) printf("ring      wait=%d lanes=%s D=%2d pollers=%2d: %lld clk/step\n", W, L ? "lane0" : "all  ", D, P, run([](long long* d) { t_ring<W, L><<<1, 32 * (3 + P)>>>(d, D, P); }))
int main() {
  printf("wait flavours: 0 = try_wait + 20 us hint, 1 = try_wait, 2 = test_wait spin, 3 = test_wait + nanosleep(32)\n");
  SELF(0, 0, 0); SELF(0, 1, 0); SELF(2, 0, 0); SELF(2, 1, 0); SELF(0, 0, 1); SELF(2, 0, 1);
  PP(0, 0, 0); PP(0, 1, 0); PP(1, 0, 0); PP(1, 1, 0); PP(2, 0, 0); PP(2, 1, 0); PP(3, 0, 0); PP(3, 1, 0); PP(0, 0, 1); PP(2, 0, 1); PP(2, 1, 1);
  RING(0, 0, 1, 0); RING(0, 0, 2, 0); RING(0, 0, 3, 0); RING(0, 0, 5, 0); RING(0, 0, 10, 0); RING(0, 0, 16, 0);
  RING(0, 0, 5, 16); RING(0, 1, 5, 0); RING(0, 1, 5, 16);
  RING(2, 0, 5, 0); RING(2, 1, 5, 0); RING(2, 0, 5, 16); RING(2, 1, 5, 16); RING(2, 1, 1, 0); RING(2, 1, 10, 0);
  RING(3, 1, 5, 0); RING(3, 1, 5, 16);
  return 0;
}
