// Discriminator element-wise / stencil ops (HBM-bound): fused bias + leaky-ReLU and upfirdn2d.
// Semantics follow exp/comm/op/fused_bias_act_kernel.cu:19-50 and
// exp/comm/op/upfirdn2d_kernel.cu:17-50 (generic form) of the reference; the implementation
// is new: 128-bit vectorised streaming for bias_act, a shared-memory tile for the FIR.
#include "c3d_common.cuh"

namespace c3d {

__device__ __forceinline__ float bias_act_one(float x, float ref, int act, int grad, float alpha) {
  // switch (act*10+grad) of fused_bias_act_kernel.cu:34-46
  if (act == 3) {
    if (grad == 0) return x > 0.f ? x : x * alpha;
    if (grad == 1) return ref > 0.f ? x : x * alpha;
    return 0.f;
  }
  return grad == 2 ? 0.f : x;  // act == 1 (linear)
}

// VEC = 4: size_x % 4 == 0 and 16B-aligned, and either step_b % 4 == 0 (a float4 never straddles a bias boundary: NCHW planes)
// or step_b == 1 with size_b % 4 == 0 (channels-last: the four lanes take four consecutive bias entries)
template <int VEC>
__global__ void __launch_bounds__(256) bias_act_kernel(const float* __restrict__ x,
                                                       const float* __restrict__ bias,
                                                       const float* __restrict__ ref,
                                                       float* __restrict__ y, long long n_vec,
                                                       int step_b, int size_b, int act, int grad,
                                                       float alpha, float scale) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n_vec; i += stride) {
    if (VEC == 4) {
      float4 v = __ldcs(reinterpret_cast<const float4*>(x) + i);
      float4 r = ref ? __ldcs(reinterpret_cast<const float4*>(ref) + i) : make_float4(0, 0, 0, 0);
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (bias) {
        if (step_b == 1) {
          b4 = __ldg(reinterpret_cast<const float4*>(bias + (i * 4) % size_b));
        } else {
          const float b = __ldg(bias + ((i * 4) / step_b) % size_b);
          b4 = make_float4(b, b, b, b);
        }
      }
      float4 o;
      o.x = bias_act_one(v.x + b4.x, r.x, act, grad, alpha) * scale;
      o.y = bias_act_one(v.y + b4.y, r.y, act, grad, alpha) * scale;
      o.z = bias_act_one(v.z + b4.z, r.z, act, grad, alpha) * scale;
      o.w = bias_act_one(v.w + b4.w, r.w, act, grad, alpha) * scale;
      __stcs(reinterpret_cast<float4*>(y) + i, o);
    } else {
      float b = bias ? __ldg(bias + (i / step_b) % size_b) : 0.f;
      float r = ref ? ref[i] : 0.f;
      y[i] = bias_act_one(x[i] + b, r, act, grad, alpha) * scale;
    }
  }
}

// upfirdn2d: one CTA per (plane, 32x32 output tile); the input footprint of the tile is staged
// in shared memory once and every tap is read from there.
constexpr int kTileW = 32, kTileH = 32, kMaxK = 8;
__global__ void __launch_bounds__(256) upfirdn2d_kernel(
    const float* __restrict__ x, const float* __restrict__ kernel, float* __restrict__ y, int in_h,
    int in_w, int out_h, int out_w, int kh, int kw, int up_x, int up_y, int down_x, int down_y,
    int pad_x0, int pad_y0, int tiles_x) {
  C3D_DYN_SMEM(float, sm);
  __shared__ float sk[kMaxK * kMaxK];
  const int plane = blockIdx.y;
  const int tx0 = (blockIdx.x % tiles_x) * kTileW, ty0 = (blockIdx.x / tiles_x) * kTileH;
  if (threadIdx.x < kh * kw) {
    int ky = threadIdx.x / kw, kx = threadIdx.x % kw;
    sk[threadIdx.x] = kernel[(kh - 1 - ky) * kw + (kw - 1 - kx)];  // flipped kernel
  }
  // footprint in the (virtual) upsampled+padded image: rows ty0*down_y .. +(kTileH-1)*down_y+kh-1
  const int fw = (kTileW - 1) * down_x + kw, fh = (kTileH - 1) * down_y + kh;
  const float* xp = x + (size_t)plane * in_h * in_w;
  for (int i = threadIdx.x; i < fw * fh; i += blockDim.x) {
    int fy = i / fw, fx = i % fw;
    int uy = ty0 * down_y + fy - pad_y0, ux = tx0 * down_x + fx - pad_x0;  // upsampled coords
    float v = 0.f;
    if (uy >= 0 && ux >= 0 && uy % up_y == 0 && ux % up_x == 0) {
      int iy = uy / up_y, ix = ux / up_x;
      if (iy < in_h && ix < in_w) v = __ldg(xp + (size_t)iy * in_w + ix);
    }
    sm[i] = v;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < kTileW * kTileH; o += blockDim.x) {
    int oy = o / kTileW, ox = o % kTileW;
    int gy = ty0 + oy, gx = tx0 + ox;
    if (gy >= out_h || gx >= out_w) continue;
    float acc = 0.f;
    for (int ky = 0; ky < kh; ++ky)
      for (int kx = 0; kx < kw; ++kx)
        acc = fmaf(sm[(oy * down_y + ky) * fw + ox * down_x + kx], sk[ky * kw + kx], acc);
    y[((size_t)plane * out_h + gy) * out_w + gx] = acc;
  }
}

// Fast path for up = down = 1 (every Blur of the discriminator): one CTA = 64 x 16 outputs of one plane.
// The (16+kh-1) x (64+kw-1) input footprint is staged with coalesced row reads; each thread produces 4
// vertically adjacent outputs from a sliding window (kh+3 shared-memory rows x kw taps).
constexpr int kBW = 64, kBH = 16, kBPitch = kBW + kMaxK;   // pitch 72 floats
// KH/KW > 0: compile-time kernel extent (fully unrolled taps, weights in registers); 0: run-time extent.
template <int KH, int KW>
__global__ void __launch_bounds__(256) blur_tile_kernel(const float* __restrict__ x, const float* __restrict__ kernel,
                                                        float* __restrict__ y, int in_h, int in_w, int out_h,
                                                        int out_w, int kh_rt, int kw_rt, int pad_x0, int pad_y0) {
  const int kh = KH > 0 ? KH : kh_rt, kw = KW > 0 ? KW : kw_rt;
  __shared__ float sm[(kBH + kMaxK - 1) * kBPitch];
  __shared__ float sk[kMaxK * kMaxK];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;      // 64 x 4 threads
  const int ox0 = blockIdx.x * kBW, oy0 = blockIdx.y * kBH;
  const float* xp = x + (size_t)blockIdx.z * in_h * in_w;
  if (threadIdx.x < kh * kw) {
    const int ky = threadIdx.x / kw, kx = threadIdx.x % kw;
    sk[threadIdx.x] = kernel[(kh - 1 - ky) * kw + (kw - 1 - kx)];   // flipped kernel (correlation form)
  }
  const int fh = kBH + kh - 1, fw = kBW + kw - 1;
  for (int r = ty; r < fh; r += 4) {
    const int iy = oy0 + r - pad_y0;
    const bool row_ok = iy >= 0 && iy < in_h;
    for (int c = tx; c < fw; c += 64) {
      const int ix = ox0 + c - pad_x0;
      sm[r * kBPitch + c] = (row_ok && ix >= 0 && ix < in_w) ? __ldg(xp + (size_t)iy * in_w + ix) : 0.f;
    }
  }
  __syncthreads();
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const int r0 = ty * 4;
  if (KH > 0) {
    float wreg[(KH > 0 ? KH : 1) * (KW > 0 ? KW : 1)];
#pragma unroll
    for (int i = 0; i < KH * KW; ++i) wreg[i] = sk[i];
#pragma unroll
    for (int r = 0; r < KH + 3; ++r) {
#pragma unroll
      for (int kx = 0; kx < KW; ++kx) {
        const float v = sm[(r0 + r) * kBPitch + tx + kx];
#pragma unroll
        for (int o = 0; o < 4; ++o)
          if (r - o >= 0 && r - o < KH) acc[o] = fmaf(v, wreg[(r - o) * KW + kx], acc[o]);   // resolved at compile time
      }
    }
  } else {
    for (int r = 0; r < kh + 3; ++r) {
      for (int kx = 0; kx < kw; ++kx) {
        const float v = sm[(r0 + r) * kBPitch + tx + kx];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          const int ky = r - o;
          if (ky >= 0 && ky < kh) acc[o] = fmaf(v, sk[ky * kw + kx], acc[o]);
        }
      }
    }
  }
  const int gx = ox0 + tx;
  if (gx < out_w) {
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const int gy = oy0 + r0 + o;
      if (gy < out_h) y[((size_t)blockIdx.z * out_h + gy) * out_w + gx] = acc[o];
    }
  }
}

// ---- Blur, streaming form (C3D_BLUR_TMA=1; opt-in until timed on hardware).
// A persistent CTA walks strips of kStripH output rows of one plane.  The strip's input rows (kStripH + 3 of them,
// each a contiguous, 16-byte aligned run of W floats) are staged into shared memory by the TMA engine -- one
// cp.async.bulk per row into a padded row (zero margins = the FIR's zero padding in x), rows above / below the plane
// are zero-filled by the threads -- double buffered behind two mbarriers, so the next strip streams in while the
// current one is filtered.  Filtering: lane = output column (conflict-free shared reads, fully coalesced stores),
// 8 output rows per thread from 11 input rows (44 LDS + 128 FMA per 8 outputs), taps in registers.  Columns beyond
// the last multiple of 32 (the output width is odd: W +- 1) go through a small per-element tail.
// Bytes: reads H*W (+ 3 halo rows per strip, L2 hits), writes OH*OW per plane.
constexpr int kStripH = 32, kMargin = 4, kRowPad = 16, kBlurThreads = 256;

struct BlurArgs {
  const float* x;
  const float* kernel;
  float* y;
  int planes, in_h, in_w, out_h, out_w, pad_x0, pad_y0, strips_per_plane;
  long long total_strips;
};

__global__ void __launch_bounds__(kBlurThreads) blur_tma_kernel(const BlurArgs a) {
  C3D_DYN_SMEM(float, sm);
  __shared__ uint64_t full[2];
  const int pitch = a.in_w + kRowPad;                 // floats per staged row: [4 zeros][W data][12 zeros]
  const int rows = kStripH + 3;
  auto buf_of = [&](int b) { return sm + (size_t)b * rows * pitch; };
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float w[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) w[i] = __ldg(a.kernel + (3 - i / 4) * 4 + (3 - i % 4));    // flipped taps (correlation form)
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    fence_mbar_init();
  }
  for (int i = tid; i < 2 * rows * pitch; i += kBlurThreads) sm[i] = 0.f;               // margins stay zero for good
  fence_proxy_async();
  __syncthreads();

  // stage strip `s` into buffer b: TMA for the rows inside the plane (thread 0), zero-fill for the others (all threads)
  auto stage = [&](long long s, int b) {
    const int plane = (int)(s / a.strips_per_plane), oy0 = (int)(s % a.strips_per_plane) * kStripH;
    const float* xp = a.x + (size_t)plane * a.in_h * a.in_w;
    const int iy0 = oy0 - a.pad_y0;
    int n_in = 0;
    for (int r = 0; r < rows; ++r) n_in += (iy0 + r >= 0 && iy0 + r < a.in_h) ? 1 : 0;
    for (int r = 0; r < rows; ++r) {
      const int iy = iy0 + r;
      if (iy >= 0 && iy < a.in_h) continue;
      for (int c = tid; c < a.in_w; c += kBlurThreads) buf_of(b)[(size_t)r * pitch + kMargin + c] = 0.f;
    }
    fence_proxy_async();      // generic writes (zero rows, and the previous consumers' reads) before the async writes
    __syncthreads();
    if (tid == 0) {
      mbar_arrive_expect_tx(&full[b], (uint32_t)(n_in * a.in_w * 4));
      for (int r = 0; r < rows; ++r) {
        const int iy = iy0 + r;
        if (iy >= 0 && iy < a.in_h)
          bulk_g2s(buf_of(b) + (size_t)r * pitch + kMargin, xp + (size_t)iy * a.in_w, (uint32_t)(a.in_w * 4), &full[b]);
      }
    }
  };

  const long long first = blockIdx.x, step = gridDim.x;
  if (first < a.total_strips) stage(first, 0);
  uint32_t par = 0;      // bit b = parity of buffer b's next completion
  int b = 0;
  for (long long s = first; s < a.total_strips; s += step, b ^= 1) {
    if (s + step < a.total_strips) stage(s + step, b ^ 1);      // prefetch the next strip into the other buffer
    mbar_wait(&full[b], (par >> b) & 1u);
    par ^= 1u << b;
    const int plane = (int)(s / a.strips_per_plane), oy0 = (int)(s % a.strips_per_plane) * kStripH;
    const int nrow = a.out_h - oy0 < kStripH ? a.out_h - oy0 : kStripH;
    float* yp = a.y + ((size_t)plane * a.out_h + oy0) * a.out_w;
    const float* sb = buf_of(b) + kMargin - a.pad_x0;      // sb[r*pitch + ox + kx] = input (iy0 + r, ox - pad_x0 + kx)
    const int q = a.out_w / 32;
    // ---- main part: warp items (column chunk c, row block rb of 8 rows), lane = column
    for (int item = warp; item < q * (kStripH / 8); item += kBlurThreads / 32) {
      const int c = item % q, rb = item / q;
      if (rb * 8 >= nrow) continue;
      const int ox = c * 32 + lane;
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const float* sp = sb + (size_t)(rb * 8) * pitch + ox;
#pragma unroll
      for (int r = 0; r < 11; ++r) {
        const float v0 = sp[r * pitch], v1 = sp[r * pitch + 1], v2 = sp[r * pitch + 2], v3 = sp[r * pitch + 3];
#pragma unroll
        for (int o = 0; o < 8; ++o) {
          const int ky = r - o;
          if (ky >= 0 && ky < 4) {      // resolved at compile time
            acc[o] = fmaf(v0, w[ky * 4 + 0], acc[o]);
            acc[o] = fmaf(v1, w[ky * 4 + 1], acc[o]);
            acc[o] = fmaf(v2, w[ky * 4 + 2], acc[o]);
            acc[o] = fmaf(v3, w[ky * 4 + 3], acc[o]);
          }
        }
      }
#pragma unroll
      for (int o = 0; o < 8; ++o)
        if (rb * 8 + o < nrow) __stcs(yp + (size_t)(rb * 8 + o) * a.out_w + ox, acc[o]);
    }
    // ---- tail columns [32 q, out_w): one output per thread and step
    const int rem = a.out_w - q * 32;
    for (int i = tid; i < rem * nrow; i += kBlurThreads) {
      const int oy = i / rem, ox = q * 32 + i % rem;
      const float* sp = sb + (size_t)oy * pitch + ox;
      float acc = 0.f;
#pragma unroll
      for (int ky = 0; ky < 4; ++ky)
#pragma unroll
        for (int kx = 0; kx < 4; ++kx) acc = fmaf(sp[ky * pitch + kx], w[ky * 4 + kx], acc);
      yp[(size_t)oy * a.out_w + ox] = acc;
    }
    __syncthreads();      // everyone is done reading this buffer before it is staged again (next iteration's prefetch)
  }
}


// ---- Blur, register-streaming form (default for the 4x4 FIR with up = down = 1: every Blur of the discriminator and its
// backward).  No shared memory, no barriers: a thread owns ONE output column and marches down a strip of kStreamRows output
// rows; the 4 x 4 window lives in registers (four rotating accumulators, one per output row in flight).  Per loop trip a
// thread issues 16 independent 4-byte loads (4 input rows x 4 taps; 3 of the 4 taps of a row are L1 hits on the neighbour
// lanes' lines) before it touches any of them, so a resident SM keeps ~128 KB of loads in flight -- the bytes-in-flight an
// HBM-bound kernel needs and the tile / TMA forms (staged through shared memory behind barriers) never had: they measured
// 15-32 % of the HBM peak (profiles/r02a_first_run.md).  Warp loads and stores are 128-byte coalesced rows (streaming
// stores).  Bytes: reads H*W (+ 3 halo rows per strip: L2 hits), writes OH*OW per plane.
constexpr int kStreamRows = 32;
// SEP: the taps are an outer product ky (x) kx (every Blur of the discriminator: make_kernel([1,3,3,1])): 4 + 4 FMAs per output
// instead of 16 -- the generic form is issue-bound (26 instructions per 8 bytes of traffic: measured 0.40-0.54 of the HBM peak,
// profiles/r02c), the separable one needs 16.  Decided per launch, inside the kernel, from the 16 taps every thread holds anyway.
template <bool SEP>
__device__ __forceinline__ void blur_stream_strip(const float* __restrict__ xp, float* __restrict__ yp, const float (&kf)[4][4],
                                                  const float (&kyv)[4], const bool (&cok)[4], int iy0, int in_h, int in_w,
                                                  int out_w, int nrow, int tap) {
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int rb = 0; rb < nrow + 3; rb += 4) {          // four input rows per trip: rows rb .. rb + 3 of the strip's footprint
    float v[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int iy = iy0 + rb + j;
      const bool rok = iy >= 0 && iy < in_h && rb + j < nrow + 3;
      const float* rp = xp + (size_t)(rok ? iy : 0) * in_w;
#pragma unroll
      for (int k = 0; k < 4; ++k) v[j][k] = (rok && cok[k]) ? __ldg(rp + k * tap) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // input row r = rb + j feeds outputs o = r - ky (ky = 0..3); output o lives in acc[o & 3] = acc[(j - ky) & 3]
      if (SEP) {
        float h = v[j][0] * kf[0][0];
#pragma unroll
        for (int k = 1; k < 4; ++k) h = fmaf(v[j][k], kf[0][k], h);
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) acc[(j - ky) & 3] = fmaf(h, kyv[ky], acc[(j - ky) & 3]);
      } else {
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
          float a = acc[(j - ky) & 3];
#pragma unroll
          for (int k = 0; k < 4; ++k) a = fmaf(v[j][k], kf[ky][k], a);
          acc[(j - ky) & 3] = a;
        }
      }
      const int o = rb + j - 3;          // complete after its fourth row
      if (o >= 0 && o < nrow) __stcs(yp + (size_t)o * out_w, acc[(j + 1) & 3]);
      acc[(j + 1) & 3] = 0.f;
    }
  }
}

// `tap` = distance in elements between horizontally adjacent pixels: 1 for NCHW planes; C for channels-last tensors, where a
// "plane" is a whole image of in_w = W * C / out_w = OW * C interleaved elements and pad_x0 counts PIXELS (c3d_blur_nhwc).
__global__ void __launch_bounds__(256) blur_stream_kernel(const float* __restrict__ x, const float* __restrict__ kernel,
                                                          float* __restrict__ y, int in_h, int in_w, int out_h, int out_w,
                                                          int pad_x0, int pad_y0, int strips_per_plane, int tap) {
  const int ox = blockIdx.y * blockDim.x + threadIdx.x;
  if (ox >= out_w) return;
  const long long plane = blockIdx.x / strips_per_plane;
  const int oy0 = (int)(blockIdx.x % strips_per_plane) * kStreamRows;
  const int nrow = out_h - oy0 < kStreamRows ? out_h - oy0 : kStreamRows;
  float kf[4][4];                      // flipped taps (correlation form), as upfirdn2d_kernel.cu:52-139
#pragma unroll
  for (int i = 0; i < 16; ++i) kf[i / 4][i % 4] = __ldg(kernel + (3 - i / 4) * 4 + (3 - i % 4));
  float kyv[4] = {1.f, 0.f, 0.f, 0.f};
  bool sep = kf[0][0] != 0.f;
  if (sep) {
#pragma unroll
    for (int i = 1; i < 4; ++i) kyv[i] = kf[i][0] / kf[0][0];
#pragma unroll
    for (int i = 1; i < 4; ++i)
#pragma unroll
      for (int j = 1; j < 4; ++j) sep = sep && fabsf(kyv[i] * kf[0][j] - kf[i][j]) <= 1e-6f * fabsf(kf[i][j]);
  }
  const int ix0 = ox - pad_x0 * tap;
  bool cok[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) cok[k] = ix0 + k * tap >= 0 && ix0 + k * tap < in_w;
  const float* xp = x + (size_t)plane * in_h * in_w + ix0;
  float* yp = y + ((size_t)plane * out_h + oy0) * out_w + ox;
  const int iy0 = oy0 - pad_y0;
  if (sep) blur_stream_strip<true>(xp, yp, kf, kyv, cok, iy0, in_h, in_w, out_w, nrow, tap);     // block-uniform branch
  else blur_stream_strip<false>(xp, yp, kf, kyv, cok, iy0, in_h, in_w, out_w, nrow, tap);
}

}  // namespace c3d

using namespace c3d;

extern "C" int c3d_bias_act(const float* x, const float* bias, const float* ref, float* y,
                            int64_t size_x, int32_t step_b, int32_t size_b, int32_t act,
                            int32_t grad, float alpha, float scale, void* stream) {
  C3D_CHECK_ARG(x && y, "bias_act: null x/y");
  C3D_CHECK_ARG(act == 1 || act == 3, "bias_act: act must be 1 (linear) or 3 (lrelu), got %d", act);
  C3D_CHECK_ARG(grad >= 0 && grad <= 2, "bias_act: grad must be 0..2");
  C3D_CHECK_ARG(!bias || (step_b > 0 && size_b > 0), "bias_act: bad bias geometry");
  C3D_CHECK_ARG(grad != 1 || ref, "bias_act: grad=1 needs ref");
  if (size_x == 0) return C3D_OK;
  cudaStream_t st = (cudaStream_t)stream;
  bool vec = size_x % 4 == 0 && (!bias || step_b % 4 == 0 || (step_b == 1 && size_b % 4 == 0 && (uintptr_t)bias % 16 == 0)) &&
             ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0) && (!ref || (uintptr_t)ref % 16 == 0);
  int dev = 0;
  cudaGetDevice(&dev);
  const int sms = c3d_device_sm_count(dev);
  if (vec) {
    long long nv = size_x / 4;
    int grid = (int)(nv / 256 + 1 < (long long)sms * 16 ? nv / 256 + 1 : (long long)sms * 16);
    C3D_LAUNCH(bias_act_kernel<4>, grid, 256, 0, st, x, bias, ref, y, nv, step_b, size_b, act, grad, alpha, scale);
  } else {
    int grid = (int)(size_x / 256 + 1 < (long long)sms * 16 ? size_x / 256 + 1 : (long long)sms * 16);
    C3D_LAUNCH(bias_act_kernel<1>, grid, 256, 0, st, x, bias, ref, y, size_x, step_b, size_b, act, grad, alpha, scale);
  }
  C3D_LAUNCH_CHECK();
  return C3D_OK;
}

static constexpr auto blur44 = blur_tile_kernel<4, 4>;
static constexpr auto blur00 = blur_tile_kernel<0, 0>;

extern "C" int c3d_upfirdn2d(const float* x, const float* kernel, float* y, int32_t planes,
                             int32_t in_h, int32_t in_w, int32_t kh, int32_t kw, int32_t up_x,
                             int32_t up_y, int32_t down_x, int32_t down_y, int32_t pad_x0,
                             int32_t pad_x1, int32_t pad_y0, int32_t pad_y1, void* stream) {
  C3D_CHECK_ARG(x && kernel && y, "upfirdn2d: null pointer");
  C3D_CHECK_ARG(kh >= 1 && kw >= 1 && kh <= kMaxK && kw <= kMaxK, "upfirdn2d: kernel up to 8x8");
  C3D_CHECK_ARG(up_x >= 1 && up_y >= 1 && down_x >= 1 && down_y >= 1 && down_x <= 4 && down_y <= 4,
                "upfirdn2d: bad up/down factors");
  int out_h = (in_h * up_y + pad_y0 + pad_y1 - kh) / down_y + 1;
  int out_w = (in_w * up_x + pad_x0 + pad_x1 - kw) / down_x + 1;
  C3D_CHECK_ARG(out_h > 0 && out_w > 0, "upfirdn2d: empty output");
  if (planes == 0) return C3D_OK;
  int tiles_x = c3d_div_up(out_w, kTileW), tiles_y = c3d_div_up(out_h, kTileH);
  int fw = (kTileW - 1) * down_x + kw, fh = (kTileH - 1) * down_y + kh;
  size_t smem = (size_t)fw * fh * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  const int blur_impl = c3d_options().blur_impl;          // 0 tile | 1 tma | 2 stream (default)
  if (up_x == 1 && up_y == 1 && down_x == 1 && down_y == 1 && kh == 4 && kw == 4 && blur_impl == 2) {
    const int strips = c3d_div_up(out_h, kStreamRows);
    const int nw = c3d_div_up(out_w, 32);
    int bw = nw <= 8 ? nw : 8;                                   // warps per block: a divisor of the column-warp count if there is one
    for (int d = 8; d >= 2 && nw > 8; --d)
      if (nw % d == 0) { bw = d; break; }
    const int col_blocks = c3d_div_up(out_w, bw * 32);
    const long long max_planes = 2147483647LL / strips;
    for (long long p0 = 0; p0 < planes; p0 += max_planes) {
      const long long np = planes - p0 < max_planes ? planes - p0 : max_planes;
      dim3 grid((unsigned)(np * strips), col_blocks);
      C3D_LAUNCH(blur_stream_kernel, grid, bw * 32, 0, st, x + (size_t)p0 * in_h * in_w, kernel, y + (size_t)p0 * out_h * out_w,
                 in_h, in_w, out_h, out_w, pad_x0, pad_y0, strips, 1);
      C3D_LAUNCH_CHECK();
    }
    return C3D_OK;
  }
  if (up_x == 1 && up_y == 1 && down_x == 1 && down_y == 1 && kh == 4 && kw == 4 && blur_impl == 1 && in_w % 4 == 0 && ((uintptr_t)x & 15u) == 0 && pad_x0 >= 0 && pad_x0 <= kMargin &&
      pad_x1 >= 0 && pad_x1 <= 4 && pad_y0 >= 0 && pad_y1 >= 0 && in_w <= 1024) {
    BlurArgs ba;
    ba.x = x; ba.kernel = kernel; ba.y = y;
    ba.planes = planes; ba.in_h = in_h; ba.in_w = in_w; ba.out_h = out_h; ba.out_w = out_w;
    ba.pad_x0 = pad_x0; ba.pad_y0 = pad_y0;
    ba.strips_per_plane = c3d_div_up(out_h, kStripH);
    ba.total_strips = (long long)planes * ba.strips_per_plane;
    const size_t bsm = (size_t)2 * (kStripH + 3) * (in_w + kRowPad) * sizeof(float);
    int dev = 0;
    cudaGetDevice(&dev);
    const int sms = c3d_device_sm_count(dev);
    int per_sm = (int)((200 * 1024) / (bsm + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    long long grid = (long long)sms * per_sm;
    if (grid > ba.total_strips) grid = ba.total_strips;
    if (bsm > 48 * 1024) C3D_CUDA(cudaFuncSetAttribute(blur_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsm));
    C3D_LAUNCH(blur_tma_kernel, (int)grid, kBlurThreads, bsm, st, ba);
    C3D_LAUNCH_CHECK();
    return C3D_OK;
  }
  if (up_x == 1 && up_y == 1 && down_x == 1 && down_y == 1) {     // the discriminator's Blur
    for (int p0 = 0; p0 < planes; p0 += 65535) {
      int np = planes - p0 < 65535 ? planes - p0 : 65535;
      dim3 grid(c3d_div_up(out_w, kBW), c3d_div_up(out_h, kBH), np);
      if (kh == 4 && kw == 4)
        C3D_LAUNCH(blur44, grid, 256, 0, st, x + (size_t)p0 * in_h * in_w, kernel, y + (size_t)p0 * out_h * out_w,
                                                     in_h, in_w, out_h, out_w, kh, kw, pad_x0, pad_y0);
      else
        C3D_LAUNCH(blur00, grid, 256, 0, st, x + (size_t)p0 * in_h * in_w, kernel, y + (size_t)p0 * out_h * out_w,
                                                     in_h, in_w, out_h, out_w, kh, kw, pad_x0, pad_y0);
      C3D_LAUNCH_CHECK();
    }
    return C3D_OK;
  }
  if (smem > 48 * 1024)
    C3D_CUDA(cudaFuncSetAttribute(upfirdn2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int p0 = 0; p0 < planes; p0 += 65535) {  // gridDim.y limit
    int np = planes - p0 < 65535 ? planes - p0 : 65535;
    dim3 grid(tiles_x * tiles_y, np);
    C3D_LAUNCH(upfirdn2d_kernel, grid, 256, smem, st, x + (size_t)p0 * in_h * in_w, kernel,
                                              y + (size_t)p0 * out_h * out_w, in_h, in_w, out_h, out_w,
                                              kh, kw, up_x, up_y, down_x, down_y, pad_x0, pad_y0, tiles_x);
    C3D_LAUNCH_CHECK();
  }
  return C3D_OK;
}

// The discriminator's 4 x 4 Blur (up = down = 1) on a channels-last tensor: x is (N, H, W, C) in memory, y (N, OH, OW, C) with
// OH = H + pad_y0 + pad_y1 - 3, OW = W + pad_x0 + pad_x1 - 3 (pads in pixels, may be negative).  Same register-streaming
// kernel: an image is one "plane" of W * C interleaved elements whose horizontal taps are C elements apart.  Not part of the
// reference's native boundary (its upfirdn2d op sees (N*C, H, W, 1) planes): it exists so that the whole discriminator can stay
// in the layout cuDNN's tensor-core convolutions want (no NCHW <-> NHWC transposes around every conv).
extern "C" int c3d_blur_nhwc(const float* x, const float* kernel, float* y, int32_t n, int32_t in_h, int32_t in_w,
                             int32_t channels, int32_t pad_x0, int32_t pad_x1, int32_t pad_y0, int32_t pad_y1, void* stream) {
  C3D_CHECK_ARG(x && kernel && y, "blur_nhwc: null pointer");
  C3D_CHECK_ARG(n >= 0 && in_h >= 1 && in_w >= 1 && channels >= 1, "blur_nhwc: bad sizes");
  const int out_h = in_h + pad_y0 + pad_y1 - 3, out_w = in_w + pad_x0 + pad_x1 - 3;
  C3D_CHECK_ARG(out_h > 0 && out_w > 0, "blur_nhwc: empty output");
  C3D_CHECK_ARG((long long)in_w * channels < 2147483647LL && (long long)out_w * channels < 2147483647LL, "blur_nhwc: row too long");
  if (n == 0) return C3D_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int strips = c3d_div_up(out_h, kStreamRows);
  const int ew = out_w * channels;                               // elements per output row
  const int col_blocks = c3d_div_up(ew, 256);
  C3D_CHECK_ARG(col_blocks <= 65535, "blur_nhwc: row too long for one launch");
  dim3 grid((unsigned)((long long)n * strips), col_blocks);
  C3D_LAUNCH(blur_stream_kernel, grid, 256, 0, st, x, kernel, y, in_h, in_w * channels, out_h, ew, pad_x0, pad_y0, strips, channels);
  C3D_LAUNCH_CHECK();
  return C3D_OK;
}
