// FiLM + sine of the training graph (HBM-bound): y = sin(gain * z + bias), z = (batch, points, channels) the linear layer's
// output, gain / bias = (batch, channels) per image -- exp/comm/.../film_layer.py:94-107 (`torch.sin(gain * self.linear(x) + bias)`)
// and piGAN_lib/siren/siren.py:83-94 (`torch.sin(freq * x + phase_shift)`).
// The fused inference kernels never materialise z; this is for the autograd graph of the NeRF branch (configs 1-3), where torch
// runs mul, add, sin forward (three passes, two saved intermediates) and cos, mul, mul, two broadcast-reductions backward.
// Here: one pass forward (8 B / element), one pass backward (read z and dy, write dz: 12 B / element) with the per-image
// reductions for dgain / dbias folded into it; only z is saved.  Arithmetic order follows torch's (mul then add, precise
// sinf / cosf), so forward values equal torch-CUDA's.
#include "c3d_common.cuh"

namespace c3d {
namespace film {

constexpr int kThreads = 256;
constexpr int kIters = 128;      // rows per chunk = kIters * (kThreads / (channels / 4))

// grid (blocks, batch).  channels / 4 divides the block size (a power of two), so a thread keeps ONE channel quad for the whole
// grid-stride loop: gain / bias live in registers and the loop has no integer division.
__global__ void __launch_bounds__(kThreads) film_sin_fwd_kernel(const float4* __restrict__ z, const float4* __restrict__ gain,
                                                                const float4* __restrict__ bias, float4* __restrict__ y,
                                                                long long per_img, int quads) {
  const long long b = blockIdx.y, stride = (long long)gridDim.x * blockDim.x;
  const int q = threadIdx.x & (quads - 1);
  const float4 g = __ldg(gain + b * quads + q), s = __ldg(bias + b * quads + q);
  const float4* zi = z + b * per_img;
  float4* yi = y + b * per_img;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_img; i += stride) {
    const float4 v = __ldcs(zi + i);
    float4 o;
    o.x = sinf(__fadd_rn(__fmul_rn(g.x, v.x), s.x));
    o.y = sinf(__fadd_rn(__fmul_rn(g.y, v.y), s.y));
    o.z = sinf(__fadd_rn(__fmul_rn(g.z, v.z), s.z));
    o.w = sinf(__fadd_rn(__fmul_rn(g.w, v.w), s.w));
    yi[i] = o;                   // read again by the next linear layer: default caching
  }
}

// grid (chunks, batch).  Thread (rg, q): channel quad q, rows rg, rg + rpi, ... of the chunk (rpi = kThreads / quads rows per
// iteration: a warp reads whole rows, 512 contiguous bytes for 128 channels).  a = dy * cos(arg); dz = a * gain; the per-thread sums
// of a (-> dbias) and a * z (-> dgain) are reduced over the chunk's row groups in shared memory and written as one partial row
// per chunk: part[(b * chunks + chunk) * 2 + {0, 1}][channels].
__global__ void __launch_bounds__(kThreads) film_sin_bwd_kernel(const float4* __restrict__ z, const float4* __restrict__ gain,
                                                                const float4* __restrict__ bias, const float4* __restrict__ dy,
                                                                float4* __restrict__ dz, float4* __restrict__ part,
                                                                long long points, int quads) {
  __shared__ float4 red[2][kThreads];
  const int rpi = kThreads / quads, q = threadIdx.x % quads, rg = threadIdx.x / quads;
  const long long b = blockIdx.y, row0 = (long long)blockIdx.x * rpi * kIters;
  const float4 g = __ldg(gain + b * quads + q), s = __ldg(bias + b * quads + q);
  float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sz = sa;
  for (int it = 0; it < kIters; ++it) {
    const long long p = row0 + (long long)it * rpi + rg;
    if (p >= points) break;
    const long long i = (b * points + p) * quads + q;
    const float4 v = __ldcs(z + i), d = __ldcs(dy + i);
    float4 a, o;
    a.x = __fmul_rn(d.x, cosf(__fadd_rn(__fmul_rn(g.x, v.x), s.x)));
    a.y = __fmul_rn(d.y, cosf(__fadd_rn(__fmul_rn(g.y, v.y), s.y)));
    a.z = __fmul_rn(d.z, cosf(__fadd_rn(__fmul_rn(g.z, v.z), s.z)));
    a.w = __fmul_rn(d.w, cosf(__fadd_rn(__fmul_rn(g.w, v.w), s.w)));
    o.x = __fmul_rn(a.x, g.x); o.y = __fmul_rn(a.y, g.y); o.z = __fmul_rn(a.z, g.z); o.w = __fmul_rn(a.w, g.w);
    dz[i] = o;
    sa.x += a.x; sa.y += a.y; sa.z += a.z; sa.w += a.w;
    sz.x = fmaf(a.x, v.x, sz.x); sz.y = fmaf(a.y, v.y, sz.y); sz.z = fmaf(a.z, v.z, sz.z); sz.w = fmaf(a.w, v.w, sz.w);
  }
  red[0][threadIdx.x] = sa;
  red[1][threadIdx.x] = sz;
  __syncthreads();
  if (rg == 0) {                 // threads 0 .. quads-1: sum the row groups in a fixed order (deterministic)
    for (int r = 1; r < rpi; ++r) {
      const float4 x = red[0][r * quads + q], w = red[1][r * quads + q];
      sa.x += x.x; sa.y += x.y; sa.z += x.z; sa.w += x.w;
      sz.x += w.x; sz.y += w.y; sz.z += w.z; sz.w += w.w;
    }
    const long long o = (b * gridDim.x + blockIdx.x) * 2;
    part[o * quads + q] = sa;
    part[(o + 1) * quads + q] = sz;
  }
}

// one thread per (image, channel): fixed-order sum of the chunk partials
__global__ void __launch_bounds__(kThreads) film_sin_bwd_finish_kernel(const float* __restrict__ part, float* __restrict__ dgain,
                                                                       float* __restrict__ dbias, int batch, int channels, int chunks) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * channels) return;
  const int b = i / channels, c = i % channels;
  float sa = 0.f, sz = 0.f;
  for (int k = 0; k < chunks; ++k) {
    const long long o = ((long long)b * chunks + k) * 2;
    sa += part[o * channels + c];
    sz += part[(o + 1) * channels + c];
  }
  dbias[i] = sa;
  dgain[i] = sz;
}

static int chunks_of(long long points, int channels) {
  const long long rows = (long long)(kThreads / (channels / 4)) * kIters;
  return (int)((points + rows - 1) / rows);
}

}  // namespace film
}  // namespace c3d

using namespace c3d;
using namespace c3d::film;

static bool film_shape_ok(int32_t batch, int64_t points, int32_t channels) {
  return batch >= 0 && batch <= 65535 && points >= 0 && channels >= 4 && channels % 4 == 0 && kThreads % (channels / 4) == 0;
}

extern "C" int c3d_film_sin_fwd(const float* z, const float* gain, const float* bias, float* y, int32_t batch, int64_t points,
                                int32_t channels, void* stream) {
  C3D_CHECK_ARG(film_shape_ok(batch, points, channels), "film_sin: channels must be a multiple of 4 with 256 %% (channels / 4) == 0 "
                "(64, 128, 256, ...), got batch %d points %lld channels %d", batch, (long long)points, channels);
  if (batch == 0 || points == 0) return C3D_OK;
  C3D_CHECK_ARG(z && gain && bias && y, "film_sin: null pointer");
  C3D_CHECK_ARG((((uintptr_t)z | (uintptr_t)gain | (uintptr_t)bias | (uintptr_t)y) & 15u) == 0, "film_sin: pointers must be 16-byte aligned");
  int dev = 0;
  cudaGetDevice(&dev);
  const int sms = c3d_device_sm_count(dev), quads = channels / 4;
  const long long per_img = (long long)points * quads;
  long long blocks = (per_img + kThreads - 1) / kThreads;
  const long long cap = ((long long)sms * 16 + batch - 1) / batch;      // ~16 blocks per SM over the whole grid
  if (blocks > cap) blocks = cap;
  C3D_LAUNCH(film_sin_fwd_kernel, dim3((unsigned)blocks, (unsigned)batch), kThreads, 0, (cudaStream_t)stream, (const float4*)z,
             (const float4*)gain, (const float4*)bias, (float4*)y, per_img, quads);
  C3D_LAUNCH_CHECK();
  return C3D_OK;
}

extern "C" size_t c3d_film_sin_bwd_workspace_bytes(int32_t batch, int64_t points, int32_t channels) {
  if (!film_shape_ok(batch, points, channels) || batch == 0 || points == 0) return 0;
  return (size_t)batch * chunks_of(points, channels) * 2 * channels * sizeof(float);
}

extern "C" int c3d_film_sin_bwd(const float* z, const float* gain, const float* bias, const float* dy, float* dz, float* dgain,
                                float* dbias, int32_t batch, int64_t points, int32_t channels, void* workspace,
                                size_t workspace_bytes, void* stream) {
  C3D_CHECK_ARG(film_shape_ok(batch, points, channels), "film_sin_bwd: channels must be a multiple of 4 with 256 %% (channels / 4) == 0, "
                "got batch %d points %lld channels %d", batch, (long long)points, channels);
  if (batch == 0) return C3D_OK;
  C3D_CHECK_ARG(dgain && dbias, "film_sin_bwd: null dgain / dbias");
  cudaStream_t st = (cudaStream_t)stream;
  if (points == 0) {
    C3D_CUDA(cudaMemsetAsync(dgain, 0, (size_t)batch * channels * sizeof(float), st));
    C3D_CUDA(cudaMemsetAsync(dbias, 0, (size_t)batch * channels * sizeof(float), st));
    return C3D_OK;
  }
  C3D_CHECK_ARG(z && gain && bias && dy && dz, "film_sin_bwd: null pointer");
  C3D_CHECK_ARG((((uintptr_t)z | (uintptr_t)gain | (uintptr_t)bias | (uintptr_t)dy | (uintptr_t)dz | (uintptr_t)workspace) & 15u) == 0,
                "film_sin_bwd: pointers must be 16-byte aligned");
  const size_t need = c3d_film_sin_bwd_workspace_bytes(batch, points, channels);
  if (workspace_bytes < need || !workspace) {
    c3d_set_error("film_sin_bwd: workspace too small (%zu < %zu)", workspace_bytes, need);
    return C3D_EWORKSPACE;
  }
  const int chunks = chunks_of(points, channels);
  C3D_LAUNCH(film_sin_bwd_kernel, dim3(chunks, batch), kThreads, 0, st, (const float4*)z, (const float4*)gain, (const float4*)bias,
             (const float4*)dy, (float4*)dz, (float4*)workspace, (long long)points, channels / 4);
  C3D_LAUNCH_CHECK();
  C3D_LAUNCH(film_sin_bwd_finish_kernel, c3d_div_up(batch * channels, kThreads), kThreads, 0, st, (const float*)workspace, dgain, dbias,
             batch, channels, chunks);
  C3D_LAUNCH_CHECK();
  return C3D_OK;
}
