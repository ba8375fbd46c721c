// Microbenchmark (sm_100a), CTA pair (tcgen05 cta_group::2, M = 256): sustained issue of NM MMAs + one multicast commit per step from
// one issuer thread, or from two issuer threads (own TMEM columns, own barriers) at once -- what the CIPS pair kernel's issuers do.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../cips-3d_b200/csrc -o issue_bench_cg2 issue_bench_cg2.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "c3d_common.cuh"
using namespace c3d;

constexpr int kIters = 1000, kWarm = 100;
struct __align__(1024) Sm {
  uint8_t a[16384];             // this CTA's 128 rows x 64 K, fp16
  uint8_t b[2][4][16384];       // per issuer: 4 stages of this CTA's half of B (N/2 rows x 64 K)
  uint64_t bar[2][4];
  uint64_t extra[2][4];         // second commit target per step (the acc_ready commits of the kernel)
  uint32_t tmem;
};

// NI issuer warps (leader CTA), each: issue NM MMAs (N = NN), commit (multicast to both CTAs if MC) to bar[i % 4] (+ NX more commits to other
// barriers), wait for the commit of 3 steps ago.  FLAGS & 4: no MMAs.
template <int NM, int NN, int NI, int MC, int NX, int FLAGS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) t_issue(long long* out) {
  extern __shared__ uint8_t raw[];
  Sm& s = *reinterpret_cast<Sm*>(raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < (int)sizeof(Sm) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(&s)[i] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 0; w < 2; ++w) for (int i = 0; i < 4; ++i) { mbar_init(&s.bar[w][i], 1); mbar_init(&s.extra[w][i], 1); }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_cg2<512>(&s.tmem);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = s.tmem;
  if (rank == 0 && warp < NI) {
    const uint32_t idesc = umma_idesc_f16(256, NN);
    const uint32_t dhi = umma_desc_hi(128);
    const uint32_t a_lo0 = umma_desc_lo(smem_u32(s.a), 128 * 16);
    const uint32_t b_lo0 = umma_desc_lo(smem_u32(s.b[warp][0]), (NN / 2) * 16);
    const uint32_t d = tmem + warp * 256;
    uint64_t* bar = s.bar[warp];
    long long t0 = 0, issue_clk = 0;
    for (int i = 0; i < kIters + kWarm; ++i) {
      if (i == kWarm) { t0 = clock64(); issue_clk = 0; }
      if (i >= 3) mbar_wait(&bar[(i - 3) & 3], ((i - 3) >> 2) & 1);
      tc_fence_after();
      if (elect_one()) {
        const long long c0 = clock64();
        if (!(FLAGS & 4)) {
#pragma unroll
          for (int m = 0; m < NM; ++m)
            umma_ss_w_cg2(d, a_lo0 + (m & 3) * 256u, b_lo0 + (i & 3) * (16384u >> 4) + (m & 3) * (uint32_t)NN, dhi, idesc, 1);
        }
        if (MC) tc_commit_cg2_mc(&bar[i & 3], 3); else tc_commit_cg2_mc(&bar[i & 3], 1);
#pragma unroll
        for (int x = 0; x < NX; ++x) tc_commit_cg2_mc(&s.extra[warp][x], 3);
        issue_clk += clock64() - c0;
      }
      __syncwarp();
    }
    for (int i = kIters + kWarm - 3; i < kIters + kWarm; ++i) mbar_wait(&bar[i & 3], (i >> 2) & 1);
    const long long t1 = clock64();
    issue_clk = __shfl_sync(0xffffffffu, issue_clk, 0);
    if ((threadIdx.x & 31) == 0) { out[2 * warp] = (t1 - t0) / kIters; out[2 * warp + 1] = issue_clk / kIters; }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_cg2<512>(tmem);
}

template <typename K> static void run(const char* name, K kern, int ni) {
  long long* d; cudaMalloc(&d, 32); cudaMemset(d, 0, 32);
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Sm) + 1024);
  kern<<<2, 128, sizeof(Sm) + 1024>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); exit(1); }
  long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost); cudaFree(d);
  printf("%-52s issuer 0: %4lld clk/step (instr %3lld)", name, h[0], h[1]);
  if (ni > 1) printf("   issuer 1: %4lld clk/step (instr %3lld)", h[2], h[3]);
  printf("\n");
}
#define RUN(NM, NN, NI, MC, NX, FLAGS) run("NM=" #NM " N=" #NN " issuers=" #NI " mc=" #MC " extra=" #NX " flags=" #FLAGS, t_issue<NM, NN, NI, MC, NX, FLAGS>, NI)
int main() {
  printf("CTA pair (cta_group::2, M = 256).  Tensor floor: N=128 -> 64 clk per MMA, N=256 -> 128 clk per MMA.  mc=1: commit multicast to both CTAs\n");
  RUN(4, 128, 1, 1, 0, 4); RUN(4, 128, 1, 0, 0, 4); RUN(4, 128, 2, 1, 0, 4);
  RUN(1, 128, 1, 1, 0, 0); RUN(2, 128, 1, 1, 0, 0); RUN(4, 128, 1, 1, 0, 0); RUN(4, 128, 1, 0, 0, 0); RUN(8, 128, 1, 1, 0, 0);
  RUN(4, 128, 1, 1, 1, 0); RUN(4, 128, 1, 1, 2, 0);
  RUN(2, 256, 1, 1, 0, 0); RUN(4, 256, 1, 1, 0, 0); RUN(8, 256, 1, 1, 0, 0);
  RUN(4, 128, 2, 1, 0, 0); RUN(8, 128, 2, 1, 0, 0); RUN(4, 256, 2, 1, 0, 0); RUN(4, 128, 2, 1, 1, 0);
  return 0;
}
