// Microbenchmark (sm_100a): what one MMA-issuer step of the CIPS kernel costs -- tcgen05.mma issue, commit, fences, indexed
// constant loads -- and how fast one / two issuer threads can feed the tensor pipe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../cips-3d_b200/csrc -o issue_bench issue_bench.cu
// One CTA, 1 issuer warp (+1).  Operands are whatever is in shared memory (zeros); M = 128, K = 16 per MMA, no swizzle.
#include <cstdio>
#include <cuda_runtime.h>
#include "c3d_common.cuh"
using namespace c3d;

constexpr int kIters = 1000, kWarm = 100;
struct __align__(1024) Sm {
  uint8_t a[16384];          // 128 x 64 fp16
  uint8_t b[4][32768];       // up to 256 x 64 fp16, 4 "stages"
  uint64_t bar[8];
  uint32_t tmem;
};
struct Tab { uint32_t e[32]; };

// MODE 0: issue NM MMAs (N = NN), commit, wait -- serial (latency of a step with its own completion)
// MODE 1: issue NM MMAs, commit to bar[i % 4], wait for the commit of 3 steps ago -- sustained feed from one thread
// FLAGS: 1 = tcgen05.fence::after_thread_sync every step, 2 = indexed constant-table load feeding the descriptors, 4 = no MMAs (commit only)
template <int NM, int NN, int MODE, int FLAGS>
__global__ void __launch_bounds__(128, 1) t_issue(long long* out, const Tab tab) {
  extern __shared__ uint8_t raw[];
  Sm& s = *reinterpret_cast<Sm*>(raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (int)sizeof(Sm) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(&s)[i] = 0;
  __syncthreads();
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&s.bar[i], 1); fence_mbar_init(); }
  if (warp == 2) tmem_alloc<512>(&s.tmem);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s.tmem;
  if (warp == 0) {
    const uint32_t idesc = umma_idesc_f16(128, NN);
    const uint32_t dhi = umma_desc_hi(128);
    const uint32_t a_lo0 = umma_desc_lo(smem_u32(s.a), 128 * 16);
    const uint32_t b_lo0 = umma_desc_lo(smem_u32(s.b[0]), NN * 16);
    long long t0 = 0, issue_clk = 0;
    for (int i = 0; i < kIters + kWarm; ++i) {
      if (i == kWarm) { t0 = clock64(); issue_clk = 0; }
      uint32_t kc = 0;
      if (FLAGS & 2) kc = tab.e[i & 31] & 3u;
      if (FLAGS & 1) tc_fence_after();
      if (MODE == 1 && i >= 3) mbar_wait(&s.bar[(i - 3) & 3], ((i - 3) >> 2) & 1);
      if (elect_one()) {
        const long long c0 = clock64();
        if (!(FLAGS & 4)) {
#pragma unroll
          for (int m = 0; m < NM; ++m)
            umma_ss_w(tmem, a_lo0 + ((kc + (m & 3)) & 3) * 256u, b_lo0 + ((i & 3) * 32768u >> 4) + (m & 3) * (uint32_t)(NN * 2), dhi, idesc, 1);
        }
        tc_commit(&s.bar[MODE == 1 ? (i & 3) : 0]);
        issue_clk += clock64() - c0;
      }
      __syncwarp();
      if (MODE == 0) mbar_wait(&s.bar[0], i & 1);
    }
    if (MODE == 1) for (int i = kIters + kWarm - 3; i < kIters + kWarm; ++i) mbar_wait(&s.bar[i & 3], (i >> 2) & 1);
    const long long t1 = clock64();
    issue_clk = __shfl_sync(0xffffffffu, issue_clk, 0);       // the elected lane is lane 0 on a converged warp
    if (threadIdx.x == 0) { out[0] = (t1 - t0) / kIters; out[1] = issue_clk / kIters; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem);
}

template <typename K> static void run(const char* name, K kern) {
  long long* d; cudaMalloc(&d, 16); cudaMemset(d, 0, 16);
  Tab tab; for (int i = 0; i < 32; ++i) tab.e[i] = (uint32_t)(i * 7);
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Sm) + 1024);
  kern<<<1, 128, sizeof(Sm) + 1024>>>(d, tab);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); exit(1); }
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); cudaFree(d);
  printf("%-64s %5lld clk/step   (issue + commit instructions: %4lld clk)\n", name, h[0], h[1]);
}
#define RUN(NM, NN, MODE, FLAGS) run("NM=" #NM " N=" #NN " mode=" #MODE " flags=" #FLAGS, t_issue<NM, NN, MODE, FLAGS>)
int main() {
  printf("mode 0: issue, commit, wait for it;  mode 1: issue, commit, wait for the commit of 3 steps ago (sustained feed)\n");
  printf("flags: 1 = tcgen05.fence::after every step, 2 = indexed constant load feeds the descriptors, 4 = commit only\n");
  printf("tensor floor: N=128 -> 64 clk per MMA, N=256 -> 128 clk per MMA\n");
  RUN(4, 128, 0, 4); RUN(4, 128, 0, 5); RUN(4, 128, 0, 6);
  RUN(1, 128, 0, 0); RUN(4, 128, 0, 0); RUN(8, 128, 0, 0); RUN(4, 256, 0, 0);
  RUN(4, 128, 1, 4); RUN(1, 128, 1, 0); RUN(2, 128, 1, 0); RUN(4, 128, 1, 0); RUN(4, 128, 1, 1); RUN(4, 128, 1, 3); RUN(8, 128, 1, 0);
  RUN(2, 256, 1, 0); RUN(4, 256, 1, 0);
  return 0;
}
