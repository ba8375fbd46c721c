// Self-test of the tcgen05 building block used by the fused kernels:
//   D[128,N] (fp32, TMEM) = A[128,K] (fp16) * B[N,K]^T (fp16)
// with B always in shared memory (K-major, no-swizzle canonical layout) and A either in shared
// memory (SS form) or in tensor memory (TS form, two fp16 per 32-bit column, row = lane).
// Exercises: TMEM alloc/dealloc, descriptor encoding, tcgen05.mma/commit, mbarrier wait,
// tcgen05.st/ld round trips, the generic->async proxy fence.
#include "c3d_common.cuh"

namespace c3d {

__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                     int N, int K, int a_in_tmem) {
  C3D_DYN_SMEM_ALIGNED(uint8_t, smem, 1024);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  __half* sA = reinterpret_cast<__half*>(smem);                 // 128*K halfs
  __half* sB = reinterpret_cast<__half*>(smem + 128 * K * 2);   // N*K halfs
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tmem_d = tmem;            // columns [0,N)
  const uint32_t tmem_a = tmem + 256;      // columns [256, 256+K/2)

  // ---- stage operands
  const uint32_t lbo_a = 128 * 16, lbo_b = (uint32_t)N * 16, sbo = 128;
  for (int i = tid; i < N * K; i += 128) {
    int n = i / K, k = i % K;
    sB[((n % 8) * 16 + (n / 8) * sbo + (k / 8) * lbo_b) / 2 + (k % 8)] = __float2half_rn(B[i]);
  }
  if (!a_in_tmem) {
    for (int i = tid; i < 128 * K; i += 128) {
      int r = i / K, k = i % K;
      sA[((r % 8) * 16 + (r / 8) * sbo + (k / 8) * lbo_a) / 2 + (k % 8)] = __float2half_rn(A[i]);
    }
  } else {
    // thread = row (lane of TMEM); warp w owns lanes 32w..32w+31
    const uint32_t taddr = tmem_a + ((uint32_t)(warp * 32) << 16);
    for (int k0 = 0; k0 < K; k0 += 16) {
      uint32_t v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        v[j] = pack_f16(A[(size_t)tid * K + k0 + 2 * j], A[(size_t)tid * K + k0 + 2 * j + 1]);
      tmem_st8(taddr + k0 / 2, v);
    }
    tc_wait_st();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // ---- single-thread MMA issue
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_f16(128, N);
    for (int k0 = 0; k0 < K; k0 += 16) {
      uint64_t bdesc = umma_desc_kmajor(smem_u32(sB) + (k0 / 8) * lbo_b, lbo_b, sbo);
      if (a_in_tmem) {
        umma_ts(tmem_d, tmem_a + k0 / 2, bdesc, idesc, k0 > 0);
      } else {
        uint64_t adesc = umma_desc_kmajor(smem_u32(sA) + (k0 / 8) * lbo_a, lbo_a, sbo);
        umma_ss(tmem_d, adesc, bdesc, idesc, k0 > 0);
      }
    }
    tc_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();

  // ---- epilogue: TMEM -> registers -> global
  const uint32_t trow = tmem_d + ((uint32_t)(warp * 32) << 16);
  for (int n0 = 0; n0 < N; n0 += 16) {
    uint32_t v[16];
    tmem_ld16(trow + n0, v);
    tc_wait_ld();
#pragma unroll
    for (int j = 0; j < 16; ++j) D[(size_t)tid * N + n0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

}  // namespace c3d

extern "C" int c3d_selftest_umma(const float* a, const float* b, float* d, int32_t n, int32_t k,
                                 int32_t a_in_tmem, void* stream) {
  C3D_CHECK_ARG(a && b && d, "selftest_umma: null pointer");
  C3D_CHECK_ARG(n >= 16 && n <= 256 && n % 16 == 0, "selftest_umma: N must be a multiple of 16 in [16,256]");
  C3D_CHECK_ARG(k >= 16 && k <= 256 && k % 16 == 0, "selftest_umma: K must be a multiple of 16 in [16,256]");
  size_t smem = (size_t)(128 + n) * k * 2;
  C3D_CUDA(cudaFuncSetAttribute(c3d::umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  C3D_LAUNCH(c3d::umma_selftest_kernel, 1, 128, smem, (cudaStream_t)stream, a, b, d, n, k, a_in_tmem);
  C3D_LAUNCH_CHECK();
  return C3D_OK;
}

// ---- cta_group::2 building block (CTA pair), in isolation:  D[256,N] = A[256,K] * B[N,K]^T
// Cluster of two CTAs; CTA r stages rows [128r, 128r+128) of A and rows [N/2*r, N/2*(r+1)) of B in ITS shared memory
// (same offsets in both CTAs), the leader (rank 0) issues tcgen05.mma.cta_group::2 (M = 256) and commits with multicast
// to the barrier of both CTAs; each CTA reads its 128 x N block of D from its own TMEM.  Exercises exactly what the
// CTA-pair CIPS kernel relies on: tcgen05.alloc/dealloc.cta_group::2, the operand partitioning of the 2-CTA MMA,
// commit.cta_group::2.multicast, a remote mbarrier arrive + cluster-scope wait (peer -> leader "operands staged").
namespace c3d {

__global__ void __launch_bounds__(128, 1)
umma_selftest_pair_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int N, int K) {
  C3D_DYN_SMEM_ALIGNED(uint8_t, smem, 1024);
  const int tid = threadIdx.x, warp = tid / 32;
  const uint32_t rank = cluster_ctarank();
  const int NH = N / 2;
  __half* sA = reinterpret_cast<__half*>(smem);                 // 128*K halfs
  __half* sB = reinterpret_cast<__half*>(smem + 128 * K * 2);   // NH*K halfs
  // barriers in dynamic shared memory: the same CTA-relative offset in both CTAs (multicast / remote arrives)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)(128 + NH) * K * 2);
  uint64_t& bar_done = bars[0];      // MMAs complete (multicast commit)
  uint64_t& bar_staged = bars[1];    // leader only: the peer has staged its operands
  uint32_t& tmem_base_s = *reinterpret_cast<uint32_t*>(&bars[2]);
  if (tid == 0) {
    mbar_init(&bar_done, 1);
    mbar_init(&bar_staged, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc_cg2<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t lbo_a = 128 * 16, lbo_b = (uint32_t)NH * 16, sbo = 128;
  for (int i = tid; i < NH * K; i += 128) {
    const int n = i / K, k = i % K;
    sB[((n % 8) * 16 + (n / 8) * sbo + (k / 8) * lbo_b) / 2 + (k % 8)] = __float2half_rn(B[(size_t)(rank * NH + n) * K + k]);
  }
  for (int i = tid; i < 128 * K; i += 128) {
    const int r = i / K, k = i % K;
    sA[((r % 8) * 16 + (r / 8) * sbo + (k / 8) * lbo_a) / 2 + (k % 8)] = __float2half_rn(A[(size_t)(rank * 128 + r) * K + k]);
  }
  fence_proxy_async_all();
  tc_fence_before();
  __syncthreads();
  if (rank == 1 && tid == 0) mbar_arrive_cluster(&bar_staged, 0);   // peer -> leader: my halves are in place
  if (rank == 0 && tid == 0) {
    mbar_wait_cluster(&bar_staged, 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_f16(256, N);
    const uint32_t dhi = umma_desc_hi(sbo);
    for (int k0 = 0; k0 < K; k0 += 16)
      umma_ss_w_cg2(tmem, umma_desc_lo(smem_u32(sA) + (k0 / 8) * lbo_a, lbo_a), umma_desc_lo(smem_u32(sB) + (k0 / 8) * lbo_b, lbo_b),
                    dhi, idesc, k0 > 0);
    tc_commit_cg2_mc(&bar_done, 3);
  }
  mbar_wait(&bar_done, 0);
  tc_fence_after();
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
  for (int n0 = 0; n0 < N; n0 += 16) {
    uint32_t v[16];
    tmem_ld16(trow + n0, v);
    tc_wait_ld();
#pragma unroll
    for (int j = 0; j < 16; ++j) D[(size_t)(rank * 128 + tid) * N + n0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_cg2<512>(tmem);
}

}  // namespace c3d

extern "C" int c3d_selftest_umma_pair(const float* a, const float* b, float* d, int32_t n, int32_t k, void* stream) {
  C3D_CHECK_ARG(a && b && d, "selftest_umma_pair: null pointer");
  C3D_CHECK_ARG(n >= 32 && n <= 256 && n % 32 == 0, "selftest_umma_pair: N must be a multiple of 32 in [32,256]");
  C3D_CHECK_ARG(k >= 16 && k <= 256 && k % 16 == 0, "selftest_umma_pair: K must be a multiple of 16 in [16,256]");
  const size_t smem = (size_t)(128 + n / 2) * k * 2 + 32;
  auto kern = c3d::umma_selftest_pair_kernel;
  C3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  c3d_count_launch();
#ifdef C3D_EMU
  C3D_CUDA(C3D_LAUNCH_CLUSTER(kern, 2, 128, smem, (cudaStream_t)stream, 2, a, b, d, n, k));
#else
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  C3D_CUDA(cudaLaunchKernelEx(&cfg, kern, a, b, d, n, k));
#endif
  return C3D_OK;
}
