// Volume integration of the training graph (HBM-bound): fancy_integration of the sorted samples of a ray --
// exp/pigan/pigan_utils.py:212-273 (called from exp/cips3d/models/generator.py:1744-1752 and, under no_grad, from
// get_fine_points_and_direction, generator_nerf_inr.py:537-598).  The fused inference kernel composites in registers; this
// is the same arithmetic as a differentiable op for the autograd graph of the NeRF branch (configs 1-3) and step (B) of the
// ray-SIREN backward plan: d(loss)/d(pixels_fea) -> d(loss)/d(sigma_i), d(loss)/d(feature_i).
// torch runs ~16 elementwise / scan launches forward (each a pass over (rays, samples) plus one over (rays, samples, C)) and
// about twice that backward, with the (rays, samples, C) product weights * rgbs materialised and saved.  Here: one pass forward
// (read (C + 2) floats per sample, write C per ray), one pass backward (read the same, write C + 1 per sample); nothing but the
// inputs is saved -- the backward recomputes alphas and transmittances in registers.
//
// Merged form (c3d_integrate_merge_*): the samples arrive as the two unsorted halves the renderer produces -- fine (rays, S, C + 1)
// with depths z_fine, coarse (rays, S, C + 1) with depths z -- and the op does torch.cat + torch.sort + torch.gather
// (generator.py:1733-1738) in registers: a stable rank sort of the 2S depths by warp shuffles; colour rows are then read from, and
// gradients written to, their SOURCE rows, so the cat / gather copies and the scatter of their backward never touch HBM.
//
// One warp per ray: lane i owns sample i (samples <= 32) for the per-sample scalars and channel(s) lane, lane + 32, ... for
// the colour sums.  Products / sums run in the sequential order of torch.cumprod and of a left-to-right sum.
#include "c3d_common.cuh"

namespace c3d {
namespace integ {

constexpr int kThreads = 256;
constexpr int kMaxKC = 4;        // channels per lane: dim_rgb <= 128

struct Args {
  const float* rgb_sigma;   // (rays, samples, channels + 1), sigma last;  merged form: the fine half (rays, samples / 2, channels + 1)
  const float* z;           // (rays, samples) sorted depths;              merged form: fine depths (rays, samples / 2), unsorted
  const float* rgb_sigma2;  // merged form: the coarse half, else null
  const float* z2;          // merged form: coarse depths
  const float* noise;       // (rays, samples) already scaled, indexed by SORTED position, or null
  const float* d_fea;       // (rays, channels)                       [bwd]
  float* fea;               // (rays, channels)                       [fwd]
  float* weights;           // (rays, samples) or null                [fwd]
  float* z_sorted;          // (rays, samples) or null                [fwd, merged form]
  float* d_rgb_sigma;       // like rgb_sigma                         [bwd]
  float* d_rgb_sigma2;      // like rgb_sigma2                        [bwd, merged form]
  long long rays;
  int samples, channels, softplus, last_back, white_back;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct Sample {      // per-lane scalars of the sample at sorted position `lane`
  float sig, delta, ex, alpha, t, trans, w, z;
  int src;           // index of that sample in cat([fine, coarse]) (= lane when the input is already sorted)
};

// row of source sample `src` (an index into cat([first, second])) in a tensor pair laid out like (rgb_sigma, rgb_sigma2)
template <class T>
__device__ __forceinline__ T* src_row(T* first, T* second, long long ray, int src, int samples, int rs) {
  if (!second) return first + (ray * samples + src) * rs;
  const int half = samples >> 1;
  return src < half ? first + (ray * half + src) * rs : second + (ray * half + (src - half)) * rs;
}

// alphas_i = 1 - exp(-delta_i * act(sigma_i + noise_i)); T_i = prod_{j<i} (1 - alpha_j + 1e-10); w_i = alpha_i * T_i
__device__ __forceinline__ Sample ray_scalars(const Args& a, long long ray, int lane) {
  const int T = a.samples, rs = a.channels + 1;
  Sample s;
  const bool on = lane < T;
  float zu, su = 0.f;                                              // this lane's UNSORTED sample
  if (a.rgb_sigma2) {
    const int half = T >> 1;
    zu = !on ? 3.0e38f : lane < half ? __ldg(a.z + ray * half + lane) : __ldg(a.z2 + ray * half + (lane - half));
  } else {
    zu = on ? __ldg(a.z + ray * T + lane) : 3.0e38f;
  }
  if (on) su = __ldg(src_row(a.rgb_sigma, a.rgb_sigma2, ray, lane, T, rs) + a.channels);
  s.src = lane;
  s.z = zu;
  s.sig = su;
  if (a.rgb_sigma2) {                                              // stable rank sort (torch.sort of the concatenated depths)
    int rank = 0;
    for (int j = 0; j < T; ++j) {
      const float zj = __shfl_sync(0xffffffffu, zu, j);
      rank += (zj < zu || (zj == zu && j < lane)) ? 1 : 0;
    }
    if (!on) rank = lane;                                          // idle lanes keep their own position
    int src = lane;
    for (int j = 0; j < T; ++j) {                                  // invert the permutation: who lands on position `lane`
      const int rj = __shfl_sync(0xffffffffu, rank, j);
      if (rj == lane) src = j;
    }
    s.src = src;
    s.z = __shfl_sync(0xffffffffu, zu, src);
    s.sig = __shfl_sync(0xffffffffu, su, src);
  }
  const float zn = __shfl_down_sync(0xffffffffu, s.z, 1);
  s.delta = lane < T - 1 ? __fsub_rn(zn, s.z) : 1e10f;
  if (a.noise && on) s.sig = __fadd_rn(s.sig, __ldg(a.noise + ray * T + lane));
  float act;
  if (a.softplus) act = s.sig > 20.f ? s.sig : log1pf(expf(s.sig));     // F.softplus: beta 1, threshold 20
  else act = fmaxf(s.sig, 0.f);
  s.ex = on ? expf(__fmul_rn(-s.delta, act)) : 1.f;
  s.alpha = __fsub_rn(1.f, s.ex);
  s.t = __fadd_rn(__fsub_rn(1.f, s.alpha), 1e-10f);
  float run = 1.f;
  s.trans = 1.f;
  for (int i = 0; i < T; ++i) {                   // torch.cumprod([1, t_0, ..., t_{T-1}])[:-1], same multiplication order
    const float ti = __shfl_sync(0xffffffffu, s.t, i);
    if (lane == i) s.trans = run;
    run = __fmul_rn(run, ti);
  }
  s.w = on ? __fmul_rn(s.alpha, s.trans) : 0.f;
  return s;
}

template <int KC>
__global__ void __launch_bounds__(kThreads) integrate_fwd_kernel(Args a) {
  const int lane = threadIdx.x & 31, T = a.samples, C = a.channels, rs = C + 1;
  const long long warps = (long long)gridDim.x * (kThreads / 32);
  for (long long ray = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); ray < a.rays; ray += warps) {
    Sample s = ray_scalars(a, ray, lane);
    const float wsum = warp_sum(s.w);
    float w = s.w;
    if (a.last_back && lane == T - 1) w = __fadd_rn(w, __fsub_rn(1.f, wsum));
    if (a.weights && lane < T) a.weights[ray * T + lane] = w;
    if (a.z_sorted && lane < T) a.z_sorted[ray * T + lane] = s.z;
    float acc[KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) acc[k] = 0.f;
    for (int i0 = 0; i0 < T; i0 += 8) {
      float col[8][KC];
#pragma unroll
      for (int j = 0; j < 8; ++j) {            // loads of eight samples in flight before the dependent FMAs
        const int src = __shfl_sync(0xffffffffu, s.src, (i0 + j) & 31);
        const float* row = src_row(a.rgb_sigma, a.rgb_sigma2, ray, src, T, rs);
#pragma unroll
        for (int k = 0; k < KC; ++k) {
          const int c = lane + 32 * k;
          col[j][k] = (i0 + j < T && c < C) ? __ldcs(row + c) : 0.f;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float wi = __shfl_sync(0xffffffffu, w, (i0 + j) & 31);
#pragma unroll
        for (int k = 0; k < KC; ++k) acc[k] = fmaf(i0 + j < T ? wi : 0.f, col[j][k], acc[k]);
      }
    }
    const float back = a.white_back ? __fsub_rn(1.f, wsum) : 0.f;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const int c = lane + 32 * k;
      if (c < C) a.fea[ray * C + c] = a.white_back ? __fadd_rn(acc[k], back) : acc[k];
    }
  }
}

// g = d(loss)/d(fea).  d_i = <g, c_i>;  e_i = d(loss)/d(w_i) = d_i - [last_back] d_{T-1} - [white_back] sum(g);
// d(alpha_i) = e_i T_i - (sum_{k>i} e_k w_k) / t_i   (the cumprod backward of torch for non-zero inputs);
// d(sigma_i) = d(alpha_i) * exp(-delta_i s_i) * delta_i * act'(sigma_i);  d(c_i) = w'_i g.
template <int KC>
__global__ void __launch_bounds__(kThreads) integrate_bwd_kernel(Args a) {
  const int lane = threadIdx.x & 31, T = a.samples, C = a.channels, rs = C + 1;
  const long long warps = (long long)gridDim.x * (kThreads / 32);
  for (long long ray = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); ray < a.rays; ray += warps) {
    Sample s = ray_scalars(a, ray, lane);
    const float wsum = warp_sum(s.w);
    float wp = s.w;                                              // w' (after last_back)
    if (a.last_back && lane == T - 1) wp = __fadd_rn(wp, __fsub_rn(1.f, wsum));
    float g[KC], gsum = 0.f;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const int c = lane + 32 * k;
      g[k] = c < C ? __ldg(a.d_fea + ray * C + c) : 0.f;
      gsum += g[k];
    }
    gsum = warp_sum(gsum);
    float d = 0.f;
    for (int i0 = 0; i0 < T; i0 += 8) {
      float part[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        part[j] = 0.f;
        const float wi = __shfl_sync(0xffffffffu, wp, (i0 + j) & 31);
        const int src = __shfl_sync(0xffffffffu, s.src, (i0 + j) & 31);
        const float* row = src_row(a.rgb_sigma, a.rgb_sigma2, ray, src, T, rs);
        float* out = src_row(a.d_rgb_sigma, a.d_rgb_sigma2, ray, src, T, rs);
#pragma unroll
        for (int k = 0; k < KC; ++k) {
          const int c = lane + 32 * k;
          if (i0 + j < T && c < C) {
            part[j] = fmaf(g[k], __ldcs(row + c), part[j]);
            __stcs(out + c, __fmul_rn(wi, g[k]));
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float di = warp_sum(part[j]);
        if (lane == i0 + j) d = di;
      }
    }
    const float dlast = __shfl_sync(0xffffffffu, d, T - 1);
    float e = d;
    if (a.last_back) e -= dlast;
    if (a.white_back) e -= gsum;
    if (lane >= T) e = 0.f;
    const float ew = e * s.w;
    float run = 0.f, suffix = 0.f;
    for (int i = T - 1; i >= 0; --i) {            // exclusive suffix sum of e_k w_k
      const float v = __shfl_sync(0xffffffffu, ew, i);
      if (lane == i) suffix = run;
      run += v;
    }
    const float dalpha = e * s.trans - suffix / s.t;
    float dact = dalpha * s.ex * s.delta;          // alpha = 1 - exp(-delta * act)
    float dsig;
    if (a.softplus) {
      const float zx = expf(s.sig);
      dsig = s.sig > 20.f ? dact : dact * zx / (zx + 1.f);
    } else {
      dsig = s.sig > 0.f ? dact : 0.f;             // relu'(0) = 0 as in torch
    }
    if (lane < T) src_row(a.d_rgb_sigma, a.d_rgb_sigma2, ray, s.src, T, rs)[C] = dsig;
  }
}

static int check(const Args& a, const char* what) {
  C3D_CHECK_ARG(a.rays >= 0 && a.samples >= 1 && a.samples <= 32, "%s: samples per ray must be 1..32, got %d (rays %lld)", what,
                a.samples, a.rays);
  C3D_CHECK_ARG(a.channels >= 1 && a.channels <= 32 * kMaxKC, "%s: channels must be 1..%d, got %d", what, 32 * kMaxKC, a.channels);
  return C3D_OK;
}

static int grid_for(long long rays) {
  int dev = 0;
  cudaGetDevice(&dev);
  const long long want = (rays + kThreads / 32 - 1) / (kThreads / 32), cap = (long long)c3d_device_sm_count(dev) * 8;
  return (int)(want < cap ? want : cap);           // 8 resident blocks of 8 warps per SM
}

}  // namespace integ
}  // namespace c3d

using namespace c3d;
using namespace c3d::integ;

#define C3D_INTEG_DISPATCH(kernel, a, st)                                              \
  do {                                                                                 \
    const int kc_ = (a.channels + 31) / 32, grid_ = grid_for(a.rays);                  \
    if (kc_ == 1) C3D_LAUNCH(kernel<1>, grid_, kThreads, 0, st, a);                    \
    else if (kc_ == 2) C3D_LAUNCH(kernel<2>, grid_, kThreads, 0, st, a);               \
    else C3D_LAUNCH(kernel<4>, grid_, kThreads, 0, st, a);                             \
    C3D_LAUNCH_CHECK();                                                                \
  } while (0)

static Args make_args(const float* rs, const float* z, const float* rs2, const float* z2, const float* noise, int64_t rays,
                      int32_t samples, int32_t channels, int32_t softplus, int32_t last_back, int32_t white_back) {
  Args a{};
  a.rgb_sigma = rs, a.z = z, a.rgb_sigma2 = rs2, a.z2 = z2, a.noise = noise;
  a.rays = (long long)rays, a.samples = samples, a.channels = channels;
  a.softplus = softplus, a.last_back = last_back, a.white_back = white_back;
  return a;
}

extern "C" int c3d_integrate_fwd(const float* rgb_sigma, const float* z, const float* noise, float* fea, float* weights,
                                 int64_t rays, int32_t samples, int32_t channels, int32_t softplus, int32_t last_back,
                                 int32_t white_back, void* stream) {
  Args a = make_args(rgb_sigma, z, nullptr, nullptr, noise, rays, samples, channels, softplus, last_back, white_back);
  a.fea = fea, a.weights = weights;
  if (int e = check(a, "integrate_fwd")) return e;
  if (rays == 0) return C3D_OK;
  C3D_CHECK_ARG(rgb_sigma && z && fea, "integrate_fwd: null pointer");
  C3D_INTEG_DISPATCH(integrate_fwd_kernel, a, (cudaStream_t)stream);
  return C3D_OK;
}

extern "C" int c3d_integrate_bwd(const float* rgb_sigma, const float* z, const float* noise, const float* d_fea,
                                 float* d_rgb_sigma, int64_t rays, int32_t samples, int32_t channels, int32_t softplus,
                                 int32_t last_back, int32_t white_back, void* stream) {
  Args a = make_args(rgb_sigma, z, nullptr, nullptr, noise, rays, samples, channels, softplus, last_back, white_back);
  a.d_fea = d_fea, a.d_rgb_sigma = d_rgb_sigma;
  if (int e = check(a, "integrate_bwd")) return e;
  if (rays == 0) return C3D_OK;
  C3D_CHECK_ARG(rgb_sigma && z && d_fea && d_rgb_sigma, "integrate_bwd: null pointer");
  C3D_INTEG_DISPATCH(integrate_bwd_kernel, a, (cudaStream_t)stream);
  return C3D_OK;
}

extern "C" int c3d_integrate_merge_fwd(const float* fine, const float* z_fine, const float* coarse, const float* z_coarse,
                                       const float* noise, float* fea, float* weights, float* z_sorted, int64_t rays,
                                       int32_t samples_each, int32_t channels, int32_t softplus, int32_t last_back,
                                       int32_t white_back, void* stream) {
  C3D_CHECK_ARG(samples_each >= 1 && samples_each <= 16, "integrate_merge_fwd: samples per half must be 1..16, got %d", samples_each);
  Args a = make_args(fine, z_fine, coarse, z_coarse, noise, rays, 2 * samples_each, channels, softplus, last_back, white_back);
  a.fea = fea, a.weights = weights, a.z_sorted = z_sorted;
  if (int e = check(a, "integrate_merge_fwd")) return e;
  if (rays == 0) return C3D_OK;
  C3D_CHECK_ARG(fine && z_fine && coarse && z_coarse && fea, "integrate_merge_fwd: null pointer");
  C3D_INTEG_DISPATCH(integrate_fwd_kernel, a, (cudaStream_t)stream);
  return C3D_OK;
}

extern "C" int c3d_integrate_merge_bwd(const float* fine, const float* z_fine, const float* coarse, const float* z_coarse,
                                       const float* noise, const float* d_fea, float* d_fine, float* d_coarse, int64_t rays,
                                       int32_t samples_each, int32_t channels, int32_t softplus, int32_t last_back,
                                       int32_t white_back, void* stream) {
  C3D_CHECK_ARG(samples_each >= 1 && samples_each <= 16, "integrate_merge_bwd: samples per half must be 1..16, got %d", samples_each);
  Args a = make_args(fine, z_fine, coarse, z_coarse, noise, rays, 2 * samples_each, channels, softplus, last_back, white_back);
  a.d_fea = d_fea, a.d_rgb_sigma = d_fine, a.d_rgb_sigma2 = d_coarse;
  if (int e = check(a, "integrate_merge_bwd")) return e;
  if (rays == 0) return C3D_OK;
  C3D_CHECK_ARG(fine && z_fine && coarse && z_coarse && d_fea && d_fine && d_coarse, "integrate_merge_bwd: null pointer");
  C3D_INTEG_DISPATCH(integrate_bwd_kernel, a, (cudaStream_t)stream);
  return C3D_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// sample_pdf, exp/pigan/pigan_utils.py:164-209 (inverse-CDF resampling of the coarse weights; called under no_grad from
// get_fine_points_and_direction, generator_nerf_inr.py:570-581), as a standalone op with the reference function's boundary:
// bins (rays, n + 1), weights (rays, n), u (rays, k) -> samples (rays, k).  One warp per ray: lane j owns weight j (n <= 32):
//   pdf = (w + eps) / sum(w + eps);  cdf = [0, cumsum(pdf)]  (left-to-right, as torch.cumsum);
//   i = searchsorted(cdf, u) (first cdf[i] >= u), below = max(i - 1, 0), above = min(i, n);
//   denom = cdf[above] - cdf[below], denom < eps -> 1;  sample = bins[below] + (u - cdf[below]) / denom * (bins[above] - bins[below]).
// Same rounding sequence as ray_math.cuh sample_pdf_ray (the fused renderer's form, validated on hardware against the reference).
namespace c3d {
namespace integ {

__global__ void __launch_bounds__(kThreads) sample_pdf_kernel(const float* __restrict__ bins, const float* __restrict__ weights,
                                                              const float* __restrict__ u, float* __restrict__ samples,
                                                              long long rays, int n, int k, float eps) {
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (kThreads / 32);
  for (long long ray = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); ray < rays; ray += warps) {
    const float w = lane < n ? __fadd_rn(__ldg(weights + ray * n + lane), eps) : 0.f;
    float sum = 0.f;
    for (int j = 0; j < n; ++j) sum = __fadd_rn(sum, __shfl_sync(0xffffffffu, w, j));
    const float pdf = __fdiv_rn(w, sum);
    // lane j holds cdf[j] (j = 0..n): exclusive running sum, built left to right; lane n's value is the total
    float run = 0.f, cdf = 0.f;
    for (int j = 0; j <= n && j < 32; ++j) {
      if (lane == j) cdf = run;
      run = __fadd_rn(run, __shfl_sync(0xffffffffu, pdf, j));
    }
    const float cdf_n = run;                       // cdf[n] when n == 32 (no lane left to hold it)
    const float bin = lane <= n && lane < 32 ? __ldg(bins + ray * (n + 1) + lane) : 0.f;
    const float bin_n = __ldg(bins + ray * (n + 1) + n);
    for (int k0 = 0; k0 < k; k0 += 32) {
      const int kk = k0 + lane;
      const float uk = kk < k ? __ldg(u + ray * k + kk) : 0.f;
      int i = 0;                                   // first i in [0, n] with cdf[i] >= u, else n + 1
      for (int j = 0; j <= n; ++j) {
        const float cj = j < 32 ? __shfl_sync(0xffffffffu, cdf, j) : cdf_n;
        i += cj < uk ? 1 : 0;                      // cdf is non-decreasing: the count of entries below u is that index
      }
      const int below = max(i - 1, 0), above = min(i, n);
      const float cb0 = __shfl_sync(0xffffffffu, cdf, below & 31), ca0 = __shfl_sync(0xffffffffu, cdf, above & 31);
      const float bb0 = __shfl_sync(0xffffffffu, bin, below & 31), ba0 = __shfl_sync(0xffffffffu, bin, above & 31);
      const float cb = below < 32 ? cb0 : cdf_n, ca = above < 32 ? ca0 : cdf_n;
      const float bb = below < 32 ? bb0 : bin_n, ba = above < 32 ? ba0 : bin_n;
      float denom = __fsub_rn(ca, cb);
      if (denom < eps) denom = 1.f;
      if (kk < k) samples[ray * k + kk] = __fadd_rn(bb, __fmul_rn(__fdiv_rn(__fsub_rn(uk, cb), denom), __fsub_rn(ba, bb)));
    }
  }
}

}  // namespace integ
}  // namespace c3d

extern "C" int c3d_sample_pdf(const float* bins, const float* weights, const float* u, float* samples, int64_t rays,
                              int32_t n_weights, int32_t n_importance, float eps, void* stream) {
  C3D_CHECK_ARG(rays >= 0 && n_weights >= 1 && n_weights <= 32 && n_importance >= 0,
                "sample_pdf: 1..32 weights per ray, got %d (rays %lld, samples %d)", n_weights, (long long)rays, n_importance);
  if (rays == 0 || n_importance == 0) return C3D_OK;
  C3D_CHECK_ARG(bins && weights && u && samples, "sample_pdf: null pointer");
  C3D_LAUNCH(sample_pdf_kernel, grid_for(rays), kThreads, 0, (cudaStream_t)stream, bins, weights, u, samples, (long long)rays,
             n_weights, n_importance, eps);
  C3D_LAUNCH_CHECK();
  return C3D_OK;
}
