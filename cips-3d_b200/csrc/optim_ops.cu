// Optimiser tail of the training step as two HBM-bound multi-tensor kernels:
//   exp/cips3d/scripts/train.py:417-438 (D) / :468-491 (G):
//       clip_grad_norm_(params, grad_clip) -> Adam.step() [-> zero_grad] -> EMA.update(state_dict)
//   exp/comm/comm_model_utils.py:99-121 EMA.update:  target = target * decay + source * (1 - decay)
// The reference runs this as ~8 elementwise torch launches PER PARAMETER TENSOR (119 tensors in G, 150 in D);
// here one launch reads every gradient once for the global L2 norm and one launch does
//   g' = g * clip;  m = lerp(m, g', 1-b1);  v = b2 v + (1-b2) g'^2;  p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps);
//   ema = ema * decay + p * (1 - decay)                       (torch.optim.Adam single-tensor formulas, amsgrad off)
// for all tensors: 20 B read + 16 B written per parameter (the second read of g comes from L2: G's gradients are
// 45 MB, the L2 holds 126 MB).  The tensor table travels as a kernel argument (<= 32 KB on sm_70+), no copies.
#include "c3d_common.cuh"

namespace c3d {
namespace opt {

constexpr int kMaxTensors = C3D_OPT_MAX_TENSORS;     // per launch; the host loops over groups
constexpr int kChunk = 4096;                          // elements per work unit (16 KB per stream)
constexpr int kThreads = 256;

struct Table {
  C3dOptTensor t[kMaxTensors];
  int chunk_start[kMaxTensors + 1];    // prefix sum of ceil(n / kChunk)
  int n_tensors;
};

// chunk index -> tensor (binary search on the prefix table; <= 8 steps)
__device__ __forceinline__ int find_tensor(const Table& tb, int chunk) {
  int lo = 0, hi = tb.n_tensors;     // invariant: chunk_start[lo] <= chunk < chunk_start[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (tb.chunk_start[mid] <= chunk) lo = mid;
    else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// partial[blockIdx.x] = sum of g^2 over the chunks this CTA walks (fixed assignment -> deterministic)
__global__ void __launch_bounds__(kThreads) grad_sq_partial_kernel(const __grid_constant__ Table tb, float* __restrict__ partial) {
  __shared__ float red[kThreads / 32];
  float acc = 0.f;
  const int total = tb.chunk_start[tb.n_tensors];
  for (int c = blockIdx.x; c < total; c += gridDim.x) {
    const int ti = find_tensor(tb, c);
    const C3dOptTensor& T = tb.t[ti];
    const long long e0 = (long long)(c - tb.chunk_start[ti]) * kChunk;
    const long long e1 = e0 + kChunk < T.n ? e0 + kChunk : T.n;
    const float* g = T.grad + e0;
    const int n = (int)(e1 - e0);
    if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
      const int n4 = n >> 2;
      for (int i = threadIdx.x; i < n4; i += kThreads) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
        acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
      }
      for (int i = (n4 << 2) + threadIdx.x; i < n; i += kThreads) acc = fmaf(g[i], g[i], acc);
    } else {
      for (int i = threadIdx.x; i < n; i += kThreads) acc = fmaf(g[i], g[i], acc);
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) s += red[w];
    partial[blockIdx.x] = s;
  }
}

// norm_out[0] = sqrt(norm_out[0]^2 (if accumulate) + sum partial);  norm_out[1] = min(1, max_norm / (norm + 1e-6))
// (torch.nn.utils.clip_grad_norm_: clip_coef = max_norm / (total_norm + 1e-6), clamped to 1)
__global__ void grad_norm_finish_kernel(const float* __restrict__ partial, int n_partial, float max_norm, int accumulate,
                                        float* __restrict__ norm_out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = accumulate ? (double)norm_out[0] * (double)norm_out[0] : 0.0;     // <= ~1200 terms, one thread
    for (int i = 0; i < n_partial; ++i) s += (double)partial[i];
    const float nrm = (float)sqrt(s);
    norm_out[0] = nrm;
    const float coef = max_norm / (nrm + 1e-6f);
    norm_out[1] = max_norm > 0.f ? fminf(coef, 1.f) : 1.f;
  }
}

// Scalars are prepared on the host the way torch prepares them: Python doubles (1 - beta, lr / bias_correction1,
// sqrt(bias_correction2), 1 - decay) rounded to fp32 when they meet the fp32 tensors.
struct Hyper {
  float beta2, eps;
  float w1, one_minus_w1;     // lerp weight 1 - beta1 and its complement
  float omb2;                 // 1 - beta2
  float neg_step_size;        // -lr / (1 - beta1^step)
  float bc2_sqrt;             // sqrt(1 - beta2^step)
  float ema_decay, omd;       // decay (< 0: no EMA update this step), 1 - decay
  int zero_grad;              // write 0 into the gradient after use (optimizer.zero_grad(set_to_none=False))
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float* ema, const Hyper& h, float clip) {
  g = __fmul_rn(g, clip);
  // exp_avg.lerp_(grad, 1 - beta1): ATen lerp = |w| < 0.5 ? a + w (b - a) : b - (b - a)(1 - w)
  const float d = __fsub_rn(g, m);
  m = h.w1 < 0.5f ? __fadd_rn(m, __fmul_rn(h.w1, d)) : __fsub_rn(g, __fmul_rn(d, h.one_minus_w1));
  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2): self + value * t1 * t2
  v = __fadd_rn(__fmul_rn(v, h.beta2), __fmul_rn(__fmul_rn(h.omb2, g), g));
  // denom = sqrt(v) / sqrt(bc2) + eps;  param.addcdiv_(exp_avg, denom, value = -step_size): self + value * t1 / t2
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), h.bc2_sqrt), h.eps);
  p = __fadd_rn(p, __fdiv_rn(__fmul_rn(h.neg_step_size, m), denom));
  if (ema) *ema = __fadd_rn(__fmul_rn(*ema, h.ema_decay), __fmul_rn(p, h.omd));
}

__global__ void __launch_bounds__(kThreads) adam_ema_kernel(const __grid_constant__ Table tb, const Hyper h,
                                                            const float* __restrict__ clip_coef) {
  const float clip = clip_coef ? __ldg(clip_coef) : 1.f;
  const int total = tb.chunk_start[tb.n_tensors];
  for (int c = blockIdx.x; c < total; c += gridDim.x) {
    const int ti = find_tensor(tb, c);
    const C3dOptTensor& T = tb.t[ti];
    const long long e0 = (long long)(c - tb.chunk_start[ti]) * kChunk;
    const long long e1 = e0 + kChunk < T.n ? e0 + kChunk : T.n;
    const int n = (int)(e1 - e0);
    float* p = T.param + e0;
    float* g = T.grad + e0;
    float* m = T.exp_avg + e0;
    float* v = T.exp_avg_sq + e0;
    float* e = (T.ema && h.ema_decay >= 0.f) ? T.ema + e0 : nullptr;
    const uintptr_t al = reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                         reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(e);
    int done = 0;
    if ((al & 15u) == 0) {
      const int n4 = n >> 2;
      for (int i = threadIdx.x; i < n4; i += kThreads) {
        float4 P = reinterpret_cast<float4*>(p)[i], M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
        const float4 Gv = reinterpret_cast<const float4*>(g)[i];
        float4 E = e ? reinterpret_cast<float4*>(e)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        adam_one(P.x, Gv.x, M.x, V.x, e ? &E.x : nullptr, h, clip);
        adam_one(P.y, Gv.y, M.y, V.y, e ? &E.y : nullptr, h, clip);
        adam_one(P.z, Gv.z, M.z, V.z, e ? &E.z : nullptr, h, clip);
        adam_one(P.w, Gv.w, M.w, V.w, e ? &E.w : nullptr, h, clip);
        reinterpret_cast<float4*>(p)[i] = P;
        reinterpret_cast<float4*>(m)[i] = M;
        reinterpret_cast<float4*>(v)[i] = V;
        if (e) reinterpret_cast<float4*>(e)[i] = E;
        if (h.zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      done = n4 << 2;
    }
    for (int i = done + threadIdx.x; i < n; i += kThreads) {
      float P = p[i], M = m[i], V = v[i], E = e ? e[i] : 0.f;
      adam_one(P, g[i], M, V, e ? &E : nullptr, h, clip);
      p[i] = P; m[i] = M; v[i] = V;
      if (e) e[i] = E;
      if (h.zero_grad) g[i] = 0.f;
    }
  }
}

// plain EMA of tensors that have no optimiser state (buffers; or every tensor before Adam is fused in)
__global__ void __launch_bounds__(kThreads) ema_kernel(const __grid_constant__ Table tb, float decay, float omd) {
  const int total = tb.chunk_start[tb.n_tensors];
  for (int c = blockIdx.x; c < total; c += gridDim.x) {
    const int ti = find_tensor(tb, c);
    const C3dOptTensor& T = tb.t[ti];
    const long long e0 = (long long)(c - tb.chunk_start[ti]) * kChunk;
    const long long e1 = e0 + kChunk < T.n ? e0 + kChunk : T.n;
    const int n = (int)(e1 - e0);
    const float* p = T.param + e0;
    float* e = T.ema + e0;
    for (int i = threadIdx.x; i < n; i += kThreads) e[i] = __fadd_rn(__fmul_rn(e[i], decay), __fmul_rn(p[i], omd));
  }
}

}  // namespace opt
}  // namespace c3d

using namespace c3d;
using namespace c3d::opt;

static int fill_table(Table& tb, const C3dOptTensor* t, int n, bool need_grad, bool need_state, bool need_ema) {
  tb.n_tensors = n;
  int chunks = 0;
  for (int i = 0; i < n; ++i) {
    C3D_CHECK_ARG(t[i].n >= 0, "optim: tensor %d has negative size", i);
    C3D_CHECK_ARG(!need_grad || t[i].grad, "optim: tensor %d has no gradient", i);
    C3D_CHECK_ARG(!need_state || (t[i].param && t[i].exp_avg && t[i].exp_avg_sq), "optim: tensor %d lacks param/exp_avg/exp_avg_sq", i);
    C3D_CHECK_ARG(!need_ema || (t[i].param && t[i].ema), "optim: tensor %d lacks param/ema", i);
    tb.t[i] = t[i];
    tb.chunk_start[i] = chunks;
    const long long c = (t[i].n + kChunk - 1) / kChunk;
    C3D_CHECK_ARG(chunks + c < (1ll << 30), "optim: too many elements in one call");
    chunks += (int)c;
  }
  tb.chunk_start[n] = chunks;
  return C3D_OK;
}

static int opt_grid(int chunks) {
  int dev = 0;
  cudaGetDevice(&dev);
  const int cap = c3d_device_sm_count(dev) * 8;     // 8 resident CTAs of 256 threads per SM
  return chunks < cap ? (chunks > 0 ? chunks : 1) : cap;
}

extern "C" size_t c3d_optim_workspace_bytes(void) { return (size_t)148 * 8 * 4 * 2 + 4096; }

extern "C" int c3d_grad_norm(const C3dOptTensor* tensors, int32_t n_tensors, float max_norm, float* norm_out,
                             void* workspace, size_t workspace_bytes, void* stream) {
  C3D_CHECK_ARG(tensors || n_tensors == 0, "grad_norm: null tensor table");
  C3D_CHECK_ARG(norm_out && workspace, "grad_norm: null norm_out / workspace");
  C3D_CHECK_ARG(workspace_bytes >= c3d_optim_workspace_bytes(), "grad_norm: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = (float*)workspace;
  int first = 1;
  for (int i0 = 0; i0 < n_tensors || first; i0 += kMaxTensors) {
    const int n = n_tensors - i0 < kMaxTensors ? n_tensors - i0 : kMaxTensors;
    Table tb;
    if (int e = fill_table(tb, tensors + i0, n > 0 ? n : 0, true, false, false)) return e;
    int grid = opt_grid(tb.chunk_start[tb.n_tensors]);
    if ((size_t)grid * 4 > workspace_bytes) grid = (int)(workspace_bytes / 4);
    C3D_LAUNCH(grad_sq_partial_kernel, grid, kThreads, 0, st, tb, partial);
    C3D_LAUNCH_CHECK();
    C3D_LAUNCH(grad_norm_finish_kernel, 1, 32, 0, st, partial, grid, max_norm, first ? 0 : 1, norm_out);
    C3D_LAUNCH_CHECK();
    first = 0;
  }
  return C3D_OK;
}

extern "C" int c3d_adam_ema_step(const C3dOptTensor* tensors, int32_t n_tensors, const C3dAdamParams* hp,
                                 const float* clip_coef, void* stream) {
  C3D_CHECK_ARG(tensors || n_tensors == 0, "adam_ema_step: null tensor table");
  C3D_CHECK_ARG(hp, "adam_ema_step: null hyper-parameters");
  C3D_CHECK_ARG(hp->step >= 1, "adam_ema_step: step must be >= 1 (it is the value AFTER the increment, as in torch)");
  C3D_CHECK_ARG(hp->beta1 >= 0. && hp->beta1 < 1. && hp->beta2 >= 0. && hp->beta2 < 1., "adam_ema_step: betas must be in [0,1)");
  Hyper h;
  h.beta2 = (float)hp->beta2; h.eps = (float)hp->eps;
  h.w1 = (float)(1.0 - hp->beta1);
  h.one_minus_w1 = (float)(1.0 - (1.0 - hp->beta1));
  h.omb2 = (float)(1.0 - hp->beta2);
  h.neg_step_size = (float)(-(hp->lr / (1.0 - pow(hp->beta1, (double)hp->step))));
  h.bc2_sqrt = (float)sqrt(1.0 - pow(hp->beta2, (double)hp->step));
  h.ema_decay = (float)hp->ema_decay;
  h.omd = (float)(1.0 - hp->ema_decay);
  h.zero_grad = hp->zero_grad;
  cudaStream_t st = (cudaStream_t)stream;
  for (int i0 = 0; i0 < n_tensors; i0 += kMaxTensors) {
    const int n = n_tensors - i0 < kMaxTensors ? n_tensors - i0 : kMaxTensors;
    Table tb;
    if (int e = fill_table(tb, tensors + i0, n, true, true, false)) return e;
    if (tb.chunk_start[n] == 0) continue;
    C3D_LAUNCH(adam_ema_kernel, opt_grid(tb.chunk_start[n]), kThreads, 0, st, tb, h, clip_coef);
    C3D_LAUNCH_CHECK();
  }
  return C3D_OK;
}

extern "C" int c3d_ema_update(const C3dOptTensor* tensors, int32_t n_tensors, double decay, void* stream) {
  C3D_CHECK_ARG(tensors || n_tensors == 0, "ema_update: null tensor table");
  cudaStream_t st = (cudaStream_t)stream;
  for (int i0 = 0; i0 < n_tensors; i0 += kMaxTensors) {
    const int n = n_tensors - i0 < kMaxTensors ? n_tensors - i0 : kMaxTensors;
    Table tb;
    if (int e = fill_table(tb, tensors + i0, n, false, false, true)) return e;
    if (tb.chunk_start[n] == 0) continue;
    C3D_LAUNCH(ema_kernel, opt_grid(tb.chunk_start[n]), kThreads, 0, st, tb, (float)decay, (float)(1.0 - decay));
    C3D_LAUNCH_CHECK();
  }
  return C3D_OK;
}

// ------------------------------------------------------------------------------------------------
// Style prep of the CIPS MLP in one launch (exp/comm/models/mod_conv_fc.py:452-496, the per-image vectors the fused kernel
// needs): for every layer l and image b
//     s1p[l][b][k]   = modulation_l(w_b)[k] + 1 = sum_j Wm_l[k][j] w_b[j] + bm_l[k] + 1                       (k < in_l)
//     demod[l][b][n] = rsqrt(sum_k (W_l[k][n] s1p[l][b][k])^2 + eps)                                            (n < 512)
// The module computes this with ~7 torch launches per layer (126 per forward, 5 % of the step with the other glue);
// here: grid (layers, B), one CTA per (layer, image).  C3D_STYLE_PREP=fused, opt-in until it has run on hardware.
namespace c3d {
namespace sprep {
constexpr int kH = 512;
__global__ void __launch_bounds__(512) style_prep_kernel(const C3dStylePrep a) {
  __shared__ float wv[kH];
  __shared__ float s1[kH];
  const int l = blockIdx.x, b = blockIdx.y;
  const int in_l = a.in_dim[l];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* w = a.style[l] + (size_t)b * a.style_dim;
  for (int j = tid; j < a.style_dim; j += blockDim.x) wv[j] = w[j];
  __syncthreads();
  // s1p: one warp per output k, lanes over the style dimension (coalesced rows of the Linear weight (in_l, style_dim))
  for (int k = warp; k < in_l; k += blockDim.x / 32) {
    const float* row = a.mod_w[l] + (size_t)k * a.style_dim;
    float acc = 0.f;
    for (int j = lane; j < a.style_dim; j += 32) acc = fmaf(row[j], wv[j], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      const float v = acc + a.mod_b[l][k] + 1.f;
      s1[k] = v;
      a.s1p[l][(size_t)b * in_l + k] = v;
    }
  }
  __syncthreads();
  // demod: thread per output column n, loop over k (coalesced rows of W (in_l, 512))
  for (int n = tid; n < kH; n += blockDim.x) {
    const float* W = a.w[l];
    float acc = 0.f;
    for (int k = 0; k < in_l; ++k) {
      const float t = W[(size_t)k * kH + n] * s1[k];
      acc = fmaf(t, t, acc);
    }
    a.demod[l][(size_t)b * kH + n] = rsqrtf(acc + a.eps);
  }
}
}  // namespace sprep
}  // namespace c3d

extern "C" int c3d_cips_style_prep(const C3dStylePrep* a, int32_t batch, void* stream) {
  C3D_CHECK_ARG(a, "style_prep: null argument");
  C3D_CHECK_ARG(a->n_layers >= 1 && a->n_layers <= C3D_CIPS_MAX_LAYERS && a->style_dim >= 1 && a->style_dim <= 512, "style_prep: bad sizes");
  for (int l = 0; l < a->n_layers; ++l) {
    C3D_CHECK_ARG(a->style[l] && a->mod_w[l] && a->mod_b[l] && a->w[l] && a->s1p[l] && a->demod[l], "style_prep: null pointer for layer %d", l);
    C3D_CHECK_ARG(a->in_dim[l] >= 1 && a->in_dim[l] <= 512, "style_prep: in_dim of layer %d", l);
  }
  if (batch == 0) return C3D_OK;
  C3D_LAUNCH(c3d::sprep::style_prep_kernel, dim3(a->n_layers, batch), 512, 0, (cudaStream_t)stream, *a);
  C3D_LAUNCH_CHECK();
  return C3D_OK;
}
