// Per-point linear layer of the NeRF training graph on tcgen05:   Y[P, N] = (X[P, K] * s) . W[N, K]^T / s  (+ bias)
// with fp32-equivalent products (both operands split into fp16 hi + lo, three MMA passes, fp32 accumulation in TMEM) --
// the GEMM of FiLMLayer.forward (exp/comm/models/film_layer.py:78-107: `self.linear(x)`) and, with W transposed, its
// data gradient dX = dZ . W.  torch runs both as fp32 SIMT sgemm (TF32 is off by default and would miss the 1e-3 bar,
// SURVEY 7.2): 42 % of a config-3 train step (profiles/r02b_train_profiles.md).  K, N in {32, 64, 128}.
//
// HBM-bound by construction (4 (K + N) bytes per point against 6 K N tensor-core MACs): the structure is the simplest one
// that keeps the tensor pipe out of the way -- 256 threads per 128-row tile: thread = (point row = TMEM lane, half): the eight
// warps split their rows' fp32 values straight into the TMEM A operand (tcgen05.st, hi and lo halves; half h owns the h-th half
// of K), one elected thread issues the 3 K/16 MMAs against the weight blob resident in shared memory, and every thread drains
// its half of the accumulator row to global memory.  Two CTAs per SM (256 TMEM columns and <= 64 KB of shared memory each)
// overlap one tile's loads with the other's MMAs and stores.  (The first version had one thread per row: 0.44 of the HBM peak,
// too few loads in flight -- profiles/r02h_ncu_plin_summary.md.)
//
// `scale` (device scalar, may be null = 1): operands are multiplied by it before the fp16 split and the result divided by it,
// so that small-magnitude inputs (gradients) stay in fp16's normal range; the caller passes 1024 / max|X|.
#include <atomic>

#include "c3d_common.cuh"

namespace c3d {
namespace plin {

constexpr int kRows = 128;
constexpr int kThreads = 256;       // two threads per row
constexpr float kWScale = 256.f, kWInv = 1.f / 256.f;       // weights are stored x 2^8 so that their lo parts stay normal

struct KArgs {
  const float* x;        // (rows, K)
  const float* bias;     // (N) or null
  const float* scale;    // device scalar or null
  float* y;              // (rows, N)
  const uint8_t* wblob;  // hi | lo fp16 UMMA-B blobs (prep kernel)
  long long rows;
  int tiles;
};

template <int N, int K>
struct Smem {
  alignas(1024) uint8_t w[2 * N * K * 2];
  alignas(8) uint64_t w_full;
  uint64_t d_ready;
  uint32_t tmem_base;
};

// fp32 (N, K) weights -> scaled fp16 hi / lo blobs in the UMMA K-major no-swizzle layout:
//   element (n, k) at (n % 8) * 16 + (n / 8) * 128 + (k / 8) * (N * 16) + (k % 8) * 2
__global__ void plin_prep_kernel(const float* __restrict__ w, int N, int K, int transposed, uint8_t* __restrict__ blob) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  __half* hi = reinterpret_cast<__half*>(blob);
  __half* lo = reinterpret_cast<__half*>(blob + (size_t)N * K * 2);
  for (int i = tid; i < N * K; i += nth) {
    const int n = i / K, k = i % K;
    // transposed: the caller's matrix is (K, N) row-major and this GEMM contracts over its FIRST index (dX = dZ . W)
    const float v = (transposed ? w[(size_t)k * N + n] : w[(size_t)n * K + k]) * kWScale;
    const __half h = __float2half_rn(v);
    const int e = ((n % 8) * 16 + (n / 8) * 128 + (k / 8) * (N * 16)) / 2 + (k % 8);
    hi[e] = h;
    lo[e] = __float2half_rn(v - __half2float(h));
  }
}

template <int N, int K>
__global__ void __launch_bounds__(kThreads, 2) plin_kernel(const KArgs a) {
  C3D_DYN_SMEM(uint8_t, smem_raw);
  Smem<N, K>& s = *reinterpret_cast<Smem<N, K>*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5;
  const int q = warp & 3, half = warp >> 2;      // TMEM lane quarter (fixed by warp % 4), K / N half
  const int r_in_tile = q * 32 + (threadIdx.x & 31);
  constexpr int kWBytes = 2 * N * K * 2;
  constexpr int KH = K / 2, NH = N / 2;          // this thread's share of a row (both multiples of 16)
  if (threadIdx.x == 0) {
    mbar_init(&s.w_full, 1);
    mbar_init(&s.d_ready, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<256>(&s.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s.tmem_base;
  if (threadIdx.x == 0) {      // weights: one bulk load per CTA, resident for the whole kernel
    mbar_arrive_expect_tx(&s.w_full, kWBytes);
    constexpr int kChunk = kWBytes < 16384 ? kWBytes : 16384;
    for (int off = 0; off < kWBytes; off += kChunk) bulk_g2s(s.w + off, a.wblob + off, kChunk, &s.w_full);
  }
  const float sc = a.scale ? __ldg(a.scale) : 1.f;
  const float inv = kWInv / sc;
  const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
  const uint32_t a_hi = tmem + lane_sel, a_lo = a_hi + 64, dcol = a_hi + 128;
  const uint32_t dhi = umma_desc_hi(128);
  const uint32_t b_hi = umma_desc_lo(smem_u32(s.w), N * 16), b_lo = umma_desc_lo(smem_u32(s.w) + N * K * 2, N * 16);
  constexpr uint32_t idesc = umma_idesc_f16(kRows, N);
  constexpr uint32_t kstep = (2u * N * 16u) >> 4;
  uint32_t dpar = 0;
  bool w_seen = false;

  for (int tile = blockIdx.x; tile < a.tiles; tile += gridDim.x) {
    const long long row = (long long)tile * kRows + r_in_tile;
    const bool row_ok = row < a.rows;
    // ---- this thread's half row: fp32 -> (hi, lo) fp16 pairs -> TMEM A operand, 16 values per step
    const float4* xr = reinterpret_cast<const float4*>(a.x + (size_t)(row_ok ? row : 0) * K + half * KH);
#pragma unroll
    for (int c = 0; c < KH / 16; ++c) {
      float4 v4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v4[j] = row_ok ? __ldg(xr + c * 4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        split_f16(v4[j].x * sc, v4[j].y * sc, hi[2 * j], lo[2 * j]);
        split_f16(v4[j].z * sc, v4[j].w * sc, hi[2 * j + 1], lo[2 * j + 1]);
      }
      tmem_st8(a_hi + (uint32_t)(half * (KH / 2) + c * 8), hi);
      tmem_st8(a_lo + (uint32_t)(half * (KH / 2) + c * 8), lo);
    }
    tc_wait_st();
    tc_fence_before();
    __syncthreads();
    // ---- three-pass split product, one elected thread of warp 0
    if (warp == 0) {
      if (!w_seen) {
        mbar_wait(&s.w_full, 0);
        w_seen = true;
      }
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d = tmem + 128, ah = tmem, al = tmem + 64;
#pragma unroll
        for (int k = 0; k < K / 16; ++k) umma_ts_w(d, ah + 8 * k, b_hi + kstep * k, dhi, idesc, k != 0);
#pragma unroll
        for (int k = 0; k < K / 16; ++k) umma_ts_w(d, al + 8 * k, b_hi + kstep * k, dhi, idesc, 1);
#pragma unroll
        for (int k = 0; k < K / 16; ++k) umma_ts_w(d, ah + 8 * k, b_lo + kstep * k, dhi, idesc, 1);
        tc_commit(&s.d_ready);
      }
      __syncwarp();
    }
    mbar_wait(&s.d_ready, dpar);
    dpar ^= 1;
    tc_fence_after();
    // ---- accumulator row -> global
    float4* yr = reinterpret_cast<float4*>(a.y + (size_t)(row_ok ? row : 0) * N);
#pragma unroll
    for (int n0 = half * NH; n0 < half * NH + NH; n0 += 16) {
      uint32_t acc[16];
      tmem_ld16(dcol + (uint32_t)n0, acc);
      tc_wait_ld();
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4 o;
          o.x = __uint_as_float(acc[4 * j]) * inv;
          o.y = __uint_as_float(acc[4 * j + 1]) * inv;
          o.z = __uint_as_float(acc[4 * j + 2]) * inv;
          o.w = __uint_as_float(acc[4 * j + 3]) * inv;
          if (a.bias) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + n0) + j);
            o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
          }
          yr[n0 / 4 + j] = o;
        }
      }
    }
    tc_fence_before();
    __syncthreads();      // every thread has drained D and the next tile may overwrite A
    tc_fence_after();
  }
  if (warp == 0 && !w_seen) mbar_wait(&s.w_full, 0);      // never leave a bulk copy in flight
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

template <int N, int K>
static int launch(const KArgs& ka, int grid, cudaStream_t st) {
  const size_t smem = sizeof(Smem<N, K>) + 1024;
  static std::atomic<unsigned long long> attr_set{0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(attr_set.load() >> (dev & 63) & 1ull)) {
    auto kern0 = plin_kernel<N, K>;
    C3D_CUDA(cudaFuncSetAttribute(kern0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set.fetch_or(1ull << (dev & 63));
  }
  auto kern = plin_kernel<N, K>;
  C3D_LAUNCH(kern, grid, kThreads, smem, st, ka);
  C3D_LAUNCH_CHECK();
  return C3D_OK;
}

}  // namespace plin
}  // namespace c3d

using namespace c3d;
using namespace c3d::plin;

extern "C" size_t c3d_points_linear_workspace_bytes(int32_t n, int32_t k) { return (size_t)2 * n * k * 2 + 256; }

// y (rows, n) = x (rows, k) . w^T (+ bias): w is (n, k) row-major, or -- transposed != 0 -- (k, n) row-major and the product
// contracts over its first index (the data gradient of the same layer).  n, k in {32, 64, 128}; x, y 16-byte aligned.
extern "C" int c3d_points_linear(const float* x, const float* w, const float* bias, const float* scale, float* y, int64_t rows,
                                 int32_t k, int32_t n, int32_t transposed, void* workspace, size_t workspace_bytes, void* stream) {
  C3D_CHECK_ARG(x && w && y, "points_linear: null pointer");
  C3D_CHECK_ARG((k == 32 || k == 64 || k == 128) && (n == 32 || n == 64 || n == 128), "points_linear: k, n must be 32, 64 or 128 (got %d, %d)", k, n);
  C3D_CHECK_ARG(rows >= 0 && rows < (1LL << 38), "points_linear: bad row count");
  C3D_CHECK_ARG(((uintptr_t)x & 15u) == 0 && ((uintptr_t)y & 15u) == 0 && (!bias || ((uintptr_t)bias & 15u) == 0), "points_linear: x, y, bias must be 16-byte aligned");
  if (workspace_bytes < c3d_points_linear_workspace_bytes(n, k) || !workspace) {
    c3d_set_error("points_linear: workspace too small");
    return C3D_EWORKSPACE;
  }
  if (rows == 0) return C3D_OK;
  int dev = 0;
  cudaGetDevice(&dev);
  if (!c3d_device_supported(dev)) {
    c3d_set_error("points_linear: device %d is not sm_100 (tcgen05 required)", dev);
    return C3D_EARCH;
  }
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* blob = (uint8_t*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  C3D_LAUNCH(plin_prep_kernel, 16, 256, 0, st, w, n, k, transposed, blob);
  C3D_LAUNCH_CHECK();
  KArgs ka;
  ka.x = x; ka.bias = bias; ka.scale = scale; ka.y = y; ka.wblob = blob; ka.rows = rows;
  const long long tiles = (rows + kRows - 1) / kRows;
  C3D_CHECK_ARG(tiles < 2147483647LL, "points_linear: too many rows");
  ka.tiles = (int)tiles;
  const int sms = c3d_device_sm_count(dev);
  int grid = (int)(tiles < 2LL * sms ? tiles : 2LL * sms);
#define C3D_PLIN_CASE(NN, KK) if (n == NN && k == KK) return launch<NN, KK>(ka, grid, st)
  C3D_PLIN_CASE(128, 128); C3D_PLIN_CASE(64, 128); C3D_PLIN_CASE(32, 128);
  C3D_PLIN_CASE(128, 64); C3D_PLIN_CASE(64, 64); C3D_PLIN_CASE(32, 64);
  C3D_PLIN_CASE(128, 32); C3D_PLIN_CASE(64, 32); C3D_PLIN_CASE(32, 32);
#undef C3D_PLIN_CASE
  return C3D_EINVAL;
}
