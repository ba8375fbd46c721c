// Microbenchmark (sm_100a), CTA pair: what slows the tensor pipe while an epilogue runs beside it.  The leader's issuer streams
// 4 x (M = 256, N = 256, K = 16) SS-form MMAs + one multicast commit per step (the CIPS pair kernel's step, 512 clk of tensor work)
// into TMEM columns 0..255, while the 16 "epilogue" warps of BOTH CTAs loop over one of:
//   tcgen05.ld of columns 256..511 (x16 or x32 per instruction), 16-byte st.shared stores (the A-operand stores), st.global stores,
//   or plain FMAs -- each with WORK dependent FMAs between two memory operations (0 = saturating, ~150 = the kernel's ratio).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../cips-3d_b200/csrc -o contention_bench contention_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "c3d_common.cuh"
using namespace c3d;

constexpr int kIters = 600, kWarm = 60;
struct __align__(1024) Sm {
  uint8_t a[16384];             // this CTA's 128 rows x 64 K, fp16
  uint8_t b[4][16384];          // 4 stages of this CTA's half of B (128 rows x 64 K)
  uint8_t x[65536];             // scratch the "epilogue" stores into (layout of the A operand: thread = row, 16 B slots)
  uint64_t bar[4];
  uint32_t tmem;
  volatile int stop;
};

// EPI: 0 none, 1 tcgen05.ld x16, 2 tcgen05.ld x32, 3 st.shared 2 x 16 B, 4 ld x16 + st.shared (the kernel's mix), 5 FMAs only, 6 st.global 2 x 16 B
template <int EPI, int WORK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(640, 1) t_cont(long long* out, float4* gscratch) {
  extern __shared__ uint8_t raw[];
  Sm& s = *reinterpret_cast<Sm*>(raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < (int)sizeof(Sm) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(&s)[i] = 0;
  __syncthreads();
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&s.bar[i], 1); fence_mbar_init(); }
  if (warp == 2) tmem_alloc_cg2<512>(&s.tmem);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = s.tmem;
  if (rank == 0 && warp == 1) {
    const uint32_t idesc = umma_idesc_f16(256, 256);
    const uint32_t dhi = umma_desc_hi(128);
    const uint32_t a_lo0 = umma_desc_lo(smem_u32(s.a), 128 * 16);
    const uint32_t b_lo0 = umma_desc_lo(smem_u32(s.b[0]), 128 * 16);
    long long t0 = 0;
    for (int i = 0; i < kIters + kWarm; ++i) {
      if (i == kWarm) t0 = clock64();
      if (i >= 3) mbar_wait(&s.bar[(i - 3) & 3], ((i - 3) >> 2) & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int m = 0; m < 4; ++m) umma_ss_w_cg2(tmem, a_lo0 + m * 256u, b_lo0 + (i & 3) * (16384u >> 4) + m * 256u, dhi, idesc, 1);
        tc_commit_cg2_mc(&s.bar[i & 3], 3);
      }
      __syncwarp();
    }
    for (int i = kIters + kWarm - 3; i < kIters + kWarm; ++i) mbar_wait(&s.bar[i & 3], (i >> 2) & 1);
    const long long t1 = clock64();
    if (lane == 0) {
      out[0] = (t1 - t0) / kIters;
      s.stop = 1;
      // stop the peer's epilogue warps too
      asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, 1;\n\tst.shared::cluster.u32 [ra], %1;\n\t}" ::"r"(smem_u32((const void*)&s.stop)), "r"(1) : "memory");
    }
  } else if (warp >= 4 && EPI != 0) {
    const int row = (warp & 3) * 32 + lane, wg = (warp - 4) >> 2;
    const uint32_t tcol = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256u + (uint32_t)wg * 32u;
    uint8_t* xp = s.x + (size_t)(wg * 4) * 2048 + row * 16;
    float4* gp = gscratch + ((size_t)blockIdx.x * 640 + threadIdx.x) * 2;
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0x3f800000u + i;
    float acc = 1.f;
    long long n = 0;
    while (!s.stop) {
      if (EPI == 1 || EPI == 4) { tmem_ld16(tcol, reinterpret_cast<uint32_t(&)[16]>(v)); tmem_ld16(tcol + 16, reinterpret_cast<uint32_t(&)[16]>(v[16])); tc_wait_ld(); }
      if (EPI == 2) { tmem_ld32(tcol, v); tc_wait_ld(); }
#pragma unroll 4
      for (int k = 0; k < WORK; ++k) acc = fmaf(acc, 1.0000001f, __uint_as_float(v[k & 31]) * 1e-9f);
      if (EPI == 3 || EPI == 4) {
        *reinterpret_cast<uint4*>(xp) = make_uint4(v[0], v[1], v[2], __float_as_uint(acc));
        *reinterpret_cast<uint4*>(xp + 2048) = make_uint4(v[4], v[5], v[6], v[7]);
        *reinterpret_cast<uint4*>(xp + 4096) = make_uint4(v[8], v[9], v[10], v[11]);
        *reinterpret_cast<uint4*>(xp + 6144) = make_uint4(v[12], v[13], v[14], v[15]);
      }
      if (EPI == 6) { gp[0] = make_float4(acc, 1.f, 2.f, 3.f); gp[1] = make_float4(acc, 4.f, 5.f, 6.f); }
      if (EPI == 7 || EPI == 8) {      // the hand-over of the kernel's epilogue: A-operand stores, proxy fence (every thread / lane 0 only), tcgen05 fence
        *reinterpret_cast<uint4*>(xp) = make_uint4(v[0], v[1], v[2], __float_as_uint(acc));
        *reinterpret_cast<uint4*>(xp + 2048) = make_uint4(v[4], v[5], v[6], v[7]);
        if (EPI == 7 || lane == 0) fence_proxy_async();
        tc_fence_before();
        __syncwarp();
      }
      if (EPI == 9) { tmem_ld16(tcol, reinterpret_cast<uint32_t(&)[16]>(v)); tc_wait_ld(); tc_fence_before(); __syncwarp(); tc_fence_after(); }
      ++n;
    }
    if (acc == 123.f) out[3] = n;      // keep the loop alive
    if (rank == 0 && warp == 4 && lane == 0) out[1] = n;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_cg2<512>(tmem);
}


// The CIPS pair kernel's weight ring around the same MMA step: warp 0 of both CTAs streams this CTA's 16 KB half of a big tile per step
// (LOADS: 0 = arrive without loading, 1 = one 16 KB bulk copy, 2 = four 4 KB bulk copies), the peer's warp 1 relays its fills to the
// leader's full barrier (count 2), the leader's commit releases the stage in both CTAs.  NST stages.
struct __align__(1024) SmR {
  uint8_t a[16384];
  uint8_t b[8][16384];
  uint64_t full[8], empty[8];
  uint32_t tmem;
};
template <int LOADS, int NST>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) t_ringmma(long long* out, const uint8_t* w, int wtiles) {
  extern __shared__ uint8_t raw[];
  SmR& s = *reinterpret_cast<SmR*>(raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < (int)sizeof(SmR) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(&s)[i] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) { mbar_init(&s.full[i], rank == 0 ? 2 : 1); mbar_init(&s.empty[i], 1); }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_cg2<512>(&s.tmem);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = s.tmem;
  const int total = kIters + kWarm;
  uint32_t st = 0, ph = 0;
  if (warp == 0) {
    for (int i = 0; i < total; ++i) {
      mbar_wait(&s.empty[st], ph ^ 1);
      if (elect_one()) {
        if (LOADS == 0) mbar_arrive(&s.full[st]);
        else {
          const uint8_t* src = w + ((size_t)((2 * i + rank) % wtiles)) * 16384;
          mbar_arrive_expect_tx(&s.full[st], 16384);
          if (LOADS == 1) bulk_g2s(s.b[st], src, 16384, &s.full[st]);
          else for (int q = 0; q < 4; ++q) bulk_g2s(s.b[st] + q * 4096, src + q * 4096, 4096, &s.full[st]);
        }
      }
      __syncwarp();
      if (++st == NST) { st = 0; ph ^= 1; }
    }
  } else if (warp == 1 && rank == 1) {
    for (int i = 0; i < total; ++i) {
      mbar_wait(&s.full[st], ph);
      if (elect_one()) mbar_arrive_cluster(&s.full[st], 0);
      __syncwarp();
      if (++st == NST) { st = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_f16(256, 256);
    const uint32_t dhi = umma_desc_hi(128);
    const uint32_t a_lo0 = umma_desc_lo(smem_u32(s.a), 128 * 16);
    const uint32_t b_lo0 = umma_desc_lo(smem_u32(s.b[0]), 128 * 16);
    long long t0 = 0;
    for (int i = 0; i < total; ++i) {
      if (i == kWarm) t0 = clock64();
      mbar_wait(&s.full[st], ph);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int m = 0; m < 4; ++m) umma_ss_w_cg2(tmem, a_lo0 + m * 256u, b_lo0 + st * (16384u >> 4) + m * 256u, dhi, idesc, 1);
        tc_commit_cg2_mc(&s.empty[st], 3);
      }
      __syncwarp();
      if (++st == NST) { st = 0; ph ^= 1; }
    }
    // drain: the last NST commits
    for (int k = 0; k < NST; ++k) { mbar_wait(&s.empty[st], ph ^ 1); if (++st == NST) { st = 0; ph ^= 1; } }
    if (lane == 0) out[0] = (clock64() - t0) / kIters;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_cg2<512>(tmem);
}
template <typename K> static void run_ring(const char* name, K kern, int blocks) {
  long long* d; cudaMalloc(&d, 32); cudaMemset(d, 0, 32);
  uint8_t* w; const int wtiles = 576; cudaMalloc(&w, (size_t)wtiles * 16384); cudaMemset(w, 0, (size_t)wtiles * 16384);
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmR) + 1024);
  kern<<<blocks, 128, sizeof(SmR) + 1024>>>(d, w, wtiles);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); exit(1); }
  long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost); cudaFree(d); cudaFree(w);
  printf("%-52s MMA step %4lld clk (floor 512)\n", name, h[0]);
}
#define RUNR(LOADS, NST, BLOCKS) run_ring("ring: loads=" #LOADS " stages=" #NST " CTAs=" #BLOCKS, t_ringmma<LOADS, NST>, BLOCKS)
template <typename K> static void run(const char* name, K kern) {
  long long* d; cudaMalloc(&d, 32); cudaMemset(d, 0, 32);
  float4* g; cudaMalloc(&g, 2 * 640 * 2 * sizeof(float4));
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Sm) + 1024);
  kern<<<2, 640, sizeof(Sm) + 1024>>>(d, g);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); exit(1); }
  long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost); cudaFree(d); cudaFree(g);
  printf("%-44s MMA step %4lld clk (floor 512)   epilogue loop iterations per warp: %lld\n", name, h[0], h[1]);
}
#define RUN(EPI, WORK) run("epilogue=" #EPI " work=" #WORK, t_cont<EPI, WORK>)
int main() {
  printf("epilogue: 0 none, 1 tcgen05.ld 2 x x16, 2 tcgen05.ld x32, 3 st.shared 4 x 16 B, 4 = 1 + 3, 5 FMAs only, 6 st.global 2 x 16 B; work = FMAs between memory operations\n");
  RUN(0, 0);
  RUN(5, 150);
  RUN(1, 0); RUN(1, 150); RUN(1, 600);
  RUN(2, 0); RUN(2, 150); RUN(2, 600);
  RUN(3, 0); RUN(3, 150); RUN(3, 600);
  RUN(4, 0); RUN(4, 150); RUN(4, 600);
  RUN(6, 150);
  printf("7 = 2 x st.shared + fence.proxy.async by every thread + tcgen05.fence, 8 = the same with the proxy fence by lane 0 only, 9 = tcgen05.ld + tcgen05 fences\n");
  RUN(7, 0); RUN(7, 150); RUN(7, 600); RUN(8, 0); RUN(8, 150); RUN(9, 0); RUN(9, 150);
  printf("weight ring around the MMA step (no epilogue); CTAs = 2: one pair alone, 148: the whole chip streams 9.4 MB from L2\n");
  RUNR(0, 5, 2); RUNR(1, 5, 2); RUNR(2, 5, 2); RUNR(1, 3, 2); RUNR(1, 8, 2);
  RUNR(0, 5, 148); RUNR(1, 5, 148); RUNR(2, 5, 148); RUNR(1, 3, 148); RUNR(1, 8, 148);
  return 0;
}
