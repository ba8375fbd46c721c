// Unfused fp32-FMA pipeline (C3D_IMPL_SIMT): the on-device cross-check for the fused
// tcgen05 kernels.  Same arithmetic, every intermediate materialised in a caller-provided
// workspace.  Not the product path; it exists so that the tensor-core kernels can be compared
// with an independent fp32 evaluation on the GPU at sizes the CPU oracle cannot reach.
#include "c3d_common.cuh"
#include "ray_math.cuh"

namespace c3d {

// ------------------------------------------------------------------ generic tiled GEMM
enum Epi { EPI_BIAS = 0, EPI_FILM_SIN = 1, EPI_DEMOD_LRELU = 2 };

struct GemmArgs {
  const float* A; int lda;          // (M,K) row-major
  const float* a_scale;             // (B,K) per-image scale of A columns or null
  const float* W; int w_sk, w_sn;   // element (k,n) at W[k*w_sk + n*w_sn]
  float* C; int ldc;                // (M,N)
  const float* bias;                // (N) or null
  const float* g; const float* b;   // (B,N): gamma/beta (FILM) or demod in g (DEMOD)
  const float* res; int ldres;      // residual added after activation, or null
  int M, N, K, rows_per_img;
};

template <int EPI>
__global__ void __launch_bounds__(256) gemm_epi_kernel(GemmArgs a) {
  constexpr int TM = 64, TN = 64, TK = 16;
  __shared__ float As[TK][TM + 4];
  __shared__ float Ws[TK][TN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  const int tx = tid % 16, ty = tid / 16;  // 16x16 threads, 4x4 outputs each
  float acc[4][4] = {};
  for (int k0 = 0; k0 < a.K; k0 += TK) {
    for (int i = tid; i < TM * TK; i += 256) {
      int r = i / TK, c = i % TK;
      int gm = m0 + r, gk = k0 + c;
      float v = 0.f;
      if (gm < a.M && gk < a.K) {
        v = a.A[(size_t)gm * a.lda + gk];
        if (a.a_scale) v *= a.a_scale[(size_t)(gm / a.rows_per_img) * a.K + gk];
      }
      As[c][r] = v;
    }
    for (int i = tid; i < TN * TK; i += 256) {
      int c, n;
      if (a.w_sn == 1) { n = i % TN; c = i / TN; } else { c = i % TK; n = i / TK; }
      int gk = k0 + c, gn = n0 + n;
      Ws[c][n] = (gk < a.K && gn < a.N) ? a.W[(size_t)gk * a.w_sk + (size_t)gn * a.w_sn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float av[4], wv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) wv[j] = Ws[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int gm = m0 + ty * 4 + i;
    if (gm >= a.M) continue;
    int img = gm / a.rows_per_img;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
      if (gn >= a.N) continue;
      float v = acc[i][j];
      if (EPI == EPI_BIAS) {
        if (a.bias) v += a.bias[gn];
      } else if (EPI == EPI_FILM_SIN) {
        v += a.bias[gn];
        v = sinf(a.g[(size_t)img * a.N + gn] * v + a.b[(size_t)img * a.N + gn]);
      } else {
        v *= a.g[(size_t)img * a.N + gn];
        v = v > 0.f ? v : 0.2f * v;
        if (a.res) v += a.res[(size_t)gm * a.ldres + gn];
      }
      a.C[(size_t)gm * a.ldc + gn] = v;
    }
  }
}

template <int EPI>
static int launch_gemm(const GemmArgs& a, cudaStream_t st) {
  dim3 grid(c3d_div_up(a.M, 64), c3d_div_up(a.N, 64));
  C3D_LAUNCH(gemm_epi_kernel<EPI>, grid, 256, 0, st, a);
  C3D_LAUNCH_CHECK();
  return C3D_OK;
}

// out[row, col0 + n] (+)= X[row,:] . W[n,:] + b[n], N <= 4 (sigma head, ToRGB); final tanh opt.
__global__ void skinny_linear_kernel(const float* __restrict__ X, int ldx, int K,
                                     const float* __restrict__ W, const float* __restrict__ b,
                                     int N, float* out, int ldo, int col0, int M, int accumulate,
                                     int do_tanh) {
  int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= M) return;
  float acc[4] = {0, 0, 0, 0};
  const float* x = X + (size_t)row * ldx;
  for (int k = 0; k < K; ++k) {
    float xv = x[k];
    for (int n = 0; n < N; ++n) acc[n] = fmaf(xv, __ldg(W + (size_t)n * K + k), acc[n]);
  }
  for (int n = 0; n < N; ++n) {
    float v = acc[n] + b[n];
    float* o = out + (size_t)row * ldo + col0 + n;
    if (accumulate) v += *o;
    if (do_tanh) v = tanhf(v);
    *o = v;
  }
}

// ------------------------------------------------------------------ ray kernels (unfused)
// First FiLM layer on the fly from ray parameters: h0 = sin(g0*(W0 (p*2/0.24) + b0) + beta0)
// (UniformBoxWarp nerf_network.py:39-45, FiLMLayer film_layer.py:78-107).
__device__ __forceinline__ float film0(const float* W0, const float* b0, const float* g0,
                                       const float* be0, int img, int j, float px, float py,
                                       float pz) {
  const float sc = 2.f / 0.24f;
  float x = px * sc, y = py * sc, z = pz * sc;
  float pre = W0[j * 3 + 0] * x + W0[j * 3 + 1] * y + W0[j * 3 + 2] * z + b0[j];
  return sinf(g0[img * 128 + j] * pre + be0[img * 128 + j]);
}

// coarse pass: one thread per (image, ray, sample, unit-group) -> z, h0
__global__ void coarse_points_h0_kernel(C3dRayParams p, C3dSirenWeights w, C3dRayIO io, int b0,
                                        int nb, float* __restrict__ zbuf, float* __restrict__ h0) {
  const int S = p.num_steps, N = p.n_rays, R = p.img_size;
  long long pt = (long long)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;  // warp per point
  long long total = (long long)nb * N * S;
  if (pt >= total) return;
  int lane = threadIdx.x % 32;
  int s = (int)(pt % S);
  long long rr = pt / S;
  int n = (int)(rr % N), bl = (int)(rr / N), b = b0 + bl;
  int ray = io.ray_idx ? io.ray_idx[n] : p.ray_offset + n;
  const float* M = io.cam2world + (size_t)b * 16;
  RayFrame f = make_ray_frame(M, ray, R, p.z_cam);
  float u = io.jitter_u[((size_t)b * R * R + ray) * S + s];
  float z, px, py, pz;
  coarse_sample(f, M, p.ray_start, p.ray_end, S, s, u, z, px, py, pz);
  if (lane == 0) zbuf[pt] = z;
  for (int j = lane; j < 128; j += 32)
    h0[(size_t)pt * 128 + j] = film0(w.w0, w.b0, w.gamma0, w.beta0, b, j, px, py, pz);
}

// importance resampling: one thread per ray -> fine z (unsorted), then h0 of the fine points
__global__ void fine_z_kernel(C3dRayParams p, C3dRayIO io, int b0, int nb,
                              const float* __restrict__ zbuf, const float* __restrict__ coarse33,
                              float* __restrict__ fzbuf) {
  const int S = p.num_steps, N = p.n_rays;
  long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= (long long)nb * N) return;
  int n = (int)(r % N), b = b0 + (int)(r / N);
  const float* z = zbuf + r * S;
  const float* sig = coarse33 + r * S * kOutC + kFeat;
  const float* nz = io.noise_c ? io.noise_c + ((size_t)b * N + n) * S : nullptr;
  float w[kMaxS], zz[kMaxS], u[kMaxS], fz[kMaxS];
  for (int i = 0; i < S; ++i) {
    zz[i] = z[i];
    u[i] = io.pdf_u[((size_t)b * N + n) * S + i];
  }
  const float ns = p.noise_std;
  integrate_weights(
      S, p.clamp_mode, [&](int i) { return zz[i]; }, [&](int i) { return sig[(size_t)i * kOutC]; },
      [&](int i) { return nz ? __fmul_rn(nz[i], ns) : 0.f; }, [&](int i, float v) { w[i] = v; });
  sample_pdf_ray(S, w, zz, u, fz);
  for (int i = 0; i < S; ++i) fzbuf[r * S + i] = fz[i];
}

__global__ void fine_points_h0_kernel(C3dRayParams p, C3dSirenWeights w, C3dRayIO io, int b0,
                                      int nb, const float* __restrict__ fzbuf,
                                      float* __restrict__ h0) {
  const int S = p.num_steps, N = p.n_rays, R = p.img_size;
  long long pt = (long long)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  if (pt >= (long long)nb * N * S) return;
  int lane = threadIdx.x % 32;
  long long rr = pt / S;
  int n = (int)(rr % N), b = b0 + (int)(rr / N);
  int ray = io.ray_idx ? io.ray_idx[n] : p.ray_offset + n;
  const float* M = io.cam2world + (size_t)b * 16;
  RayFrame f = make_ray_frame(M, ray, R, p.z_cam);
  float px, py, pz;
  fine_sample(f, fzbuf[pt], px, py, pz);
  for (int j = lane; j < 128; j += 32)
    h0[(size_t)pt * 128 + j] = film0(w.w0, w.b0, w.gamma0, w.beta0, b, j, px, py, pz);
}

// merge + final integration: one thread per ray
__global__ void composite_kernel(C3dRayParams p, C3dRayIO io, int b0, int nb,
                                 const float* __restrict__ zbuf, const float* __restrict__ coarse33,
                                 const float* __restrict__ fzbuf, const float* __restrict__ fine33) {
  const int S = p.num_steps, N = p.n_rays;
  const int nS = p.hierarchical ? 2 * S : S;
  long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= (long long)nb * N) return;
  int n = (int)(r % N), b = b0 + (int)(r / N);
  size_t ro = (size_t)b * N + n;  // output ray slot
  float key[kMaxNS];
  int idx[kMaxNS];
  if (p.hierarchical) {
    for (int i = 0; i < S; ++i) { key[i] = fzbuf[r * S + i]; idx[i] = i; }          // fine first
    for (int i = 0; i < S; ++i) { key[S + i] = zbuf[r * S + i]; idx[S + i] = S + i; }
    sort_keys(nS, key, idx);
  } else {
    for (int i = 0; i < S; ++i) { key[i] = zbuf[r * S + i]; idx[i] = S + i; }
  }
  auto row = [&](int i) -> const float* {
    int id = idx[i];
    return id < S ? fine33 + ((size_t)r * S + id) * kOutC : coarse33 + ((size_t)r * S + (id - S)) * kOutC;
  };
  const float* nz = io.noise_f ? io.noise_f + ro * nS : nullptr;
  const float ns = p.noise_std;
  float w[kMaxNS];
  float wsum = integrate_weights(
      nS, p.clamp_mode, [&](int i) { return key[i]; }, [&](int i) { return row(i)[kFeat]; },
      [&](int i) { return nz ? __fmul_rn(nz[i], ns) : 0.f; }, [&](int i, float v) { w[i] = v; });
  if (p.last_back) w[nS - 1] += 1.f - wsum;
  float depth = 0.f;
  float acc[kFeat];
  for (int c = 0; c < kFeat; ++c) acc[c] = 0.f;
  for (int i = 0; i < nS; ++i) {
    const float* f = row(i);
    for (int c = 0; c < kFeat; ++c) acc[c] = fmaf(w[i], f[c], acc[c]);
    depth = fmaf(w[i], key[i], depth);
  }
  for (int c = 0; c < kFeat; ++c) io.pixels_fea[ro * kFeat + c] = acc[c] + (p.white_back ? 1.f - wsum : 0.f);
  if (io.depth) io.depth[ro] = depth;
  if (io.weights) for (int i = 0; i < nS; ++i) io.weights[ro * nS + i] = w[i];
  if (io.dbg_all_z) for (int i = 0; i < nS; ++i) io.dbg_all_z[ro * nS + i] = key[i];
}

__global__ void copy_rows_kernel(const float* src, float* dst, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

// h0 -> [feature32, sigma] for P points of images [b0, b0+nb)
static int siren_tail(const C3dSirenWeights& w, long long P, int rows_per_img, int b0,
                      const float* h0, float* h1, float* h2, float* out33, cudaStream_t st) {
  GemmArgs g{};
  g.rows_per_img = rows_per_img;
  g.M = (int)P;
  // network.1
  g.A = h0; g.lda = 128; g.K = 128; g.N = 128; g.W = w.w1; g.w_sk = 1; g.w_sn = 128;
  g.C = h1; g.ldc = 128; g.bias = w.b1; g.g = w.gamma1 + (size_t)b0 * 128; g.b = w.beta1 + (size_t)b0 * 128;
  if (int e = launch_gemm<EPI_FILM_SIN>(g, st)) return e;
  // sigma head -> out33[:, 32]
  C3D_LAUNCH(skinny_linear_kernel, c3d_div_up(P, 256), 256, 0, st, h1, 128, 128, w.w_sigma, w.b_sigma, 1, out33,
                                                          kOutC, kFeat, (int)P, 0, 0);
  C3D_LAUNCH_CHECK();
  // color_layer_sine
  g.A = h1; g.N = 64; g.W = w.wc; g.w_sn = 128; g.C = h2; g.ldc = 64; g.bias = w.bc;
  g.g = w.gammac + (size_t)b0 * 64; g.b = w.betac + (size_t)b0 * 64;
  if (int e = launch_gemm<EPI_FILM_SIN>(g, st)) return e;
  // color_layer_linear -> out33[:, 0:32]
  g.A = h2; g.lda = 64; g.K = 64; g.N = 32; g.W = w.wl; g.w_sn = 64; g.C = out33; g.ldc = kOutC;
  g.bias = w.bl; g.g = g.b = nullptr;
  return launch_gemm<EPI_BIAS>(g, st);
}

}  // namespace c3d

using namespace c3d;

static long long simt_chunk_images(const C3dRayParams* p) {
  const long long budget_pts = 1ll << 21;  // ~3 GB of intermediates
  long long per_img = (long long)p->n_rays * p->num_steps;
  long long nb = budget_pts / (per_img > 0 ? per_img : 1);
  if (nb < 1) nb = 1;
  if (nb > p->batch) nb = p->batch;
  return nb;
}

size_t c3d_ray_siren_simt_workspace_bytes(const C3dRayParams* p) {
  long long P = simt_chunk_images(p) * p->n_rays * p->num_steps;
  // z, fz, h0(128), h1(128), h2(64), coarse33, fine33
  return (size_t)P * (1 + 1 + 128 + 128 + 64 + 33 + 33) * sizeof(float);
}

int c3d_ray_siren_fwd_simt(const C3dRayParams* p, const C3dSirenWeights* w, const C3dRayIO* io,
                           void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (workspace_bytes < c3d_ray_siren_simt_workspace_bytes(p)) {
    c3d_set_error("ray_siren(simt): workspace too small");
    return C3D_EWORKSPACE;
  }
  const int S = p->num_steps, N = p->n_rays;
  const long long cb = simt_chunk_images(p);
  const long long Pmax = cb * N * S;
  float* zbuf = (float*)workspace;
  float* fzbuf = zbuf + Pmax;
  float* h0 = fzbuf + Pmax;
  float* h1 = h0 + Pmax * 128;
  float* h2 = h1 + Pmax * 128;
  float* c33 = h2 + Pmax * 64;
  float* f33 = c33 + Pmax * 33;
  for (int b0 = 0; b0 < p->batch; b0 += (int)cb) {
    int nb = (int)((p->batch - b0) < cb ? (p->batch - b0) : cb);
    long long P = (long long)nb * N * S, NR = (long long)nb * N;
    C3D_LAUNCH(coarse_points_h0_kernel, c3d_div_up(P, 8), 256, 0, st, *p, *w, *io, b0, nb, zbuf, h0);
    C3D_LAUNCH_CHECK();
    if (int e = siren_tail(*w, P, N * S, b0, h0, h1, h2, c33, st)) return e;
    if (io->dbg_coarse) {
      C3D_LAUNCH(copy_rows_kernel, c3d_div_up(P * 33, 256), 256, 0, st, c33, io->dbg_coarse + (size_t)b0 * N * S * 33, P * 33);
      C3D_LAUNCH_CHECK();
    }
    if (p->hierarchical) {
      C3D_LAUNCH(fine_z_kernel, c3d_div_up(NR, 128), 128, 0, st, *p, *io, b0, nb, zbuf, c33, fzbuf);
      C3D_LAUNCH_CHECK();
      C3D_LAUNCH(fine_points_h0_kernel, c3d_div_up(P, 8), 256, 0, st, *p, *w, *io, b0, nb, fzbuf, h0);
      C3D_LAUNCH_CHECK();
      if (int e = siren_tail(*w, P, N * S, b0, h0, h1, h2, f33, st)) return e;
      if (io->dbg_fine) {
        C3D_LAUNCH(copy_rows_kernel, c3d_div_up(P * 33, 256), 256, 0, st, f33, io->dbg_fine + (size_t)b0 * N * S * 33, P * 33);
        C3D_LAUNCH_CHECK();
      }
    }
    C3D_LAUNCH(composite_kernel, c3d_div_up(NR, 128), 128, 0, st, *p, *io, b0, nb, zbuf, c33, fzbuf, f33);
    C3D_LAUNCH_CHECK();
  }
  return C3D_OK;
}

// ------------------------------------------------------------------ CIPS (unfused)
size_t c3d_cips_simt_workspace_bytes(const C3dCipsParams* p) {
  return (size_t)p->batch * p->n_pix * p->hidden * 3 * sizeof(float);  // x_a, x_b, x_block
}

int c3d_cips_fwd_simt(const C3dCipsParams* p, const C3dCipsWeights* w, const float* x, float* rgb,
                      float* hidden_out, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (workspace_bytes < c3d_cips_simt_workspace_bytes(p)) {
    c3d_set_error("cips(simt): workspace too small");
    return C3D_EWORKSPACE;
  }
  const long long M = (long long)p->batch * p->n_pix;
  const int H = p->hidden;
  float* buf[3] = {(float*)workspace, (float*)workspace + M * H, (float*)workspace + 2 * M * H};
  const float* cur = x;
  int cur_dim = p->in_dim;
  int cur_buf = -1;
  bool rgb_started = false;
  for (int blk = 0; blk < p->n_blocks; ++blk) {
    // pick two scratch buffers different from the one holding the block input
    int t1 = (cur_buf + 1 + 3) % 3, t2 = (cur_buf + 2 + 3) % 3;
    if (cur_buf < 0) { t1 = 0; t2 = 1; }
    GemmArgs g{};
    g.rows_per_img = p->n_pix; g.M = (int)M; g.N = H; g.w_sk = H; g.w_sn = 1;
    // mod1 + lrelu
    g.A = cur; g.lda = cur_dim; g.K = cur_dim; g.a_scale = w->style1p[2 * blk]; g.W = w->w[2 * blk];
    g.g = w->demod[2 * blk]; g.C = buf[t1]; g.ldc = H;
    if (int e = launch_gemm<EPI_DEMOD_LRELU>(g, st)) return e;
    // mod2 + lrelu (+ residual)
    g.A = buf[t1]; g.lda = H; g.K = H; g.a_scale = w->style1p[2 * blk + 1]; g.W = w->w[2 * blk + 1];
    g.g = w->demod[2 * blk + 1]; g.C = buf[t2];
    if (blk >= p->skip_from && cur_dim == H) { g.res = cur; g.ldres = H; }
    if (int e = launch_gemm<EPI_DEMOD_LRELU>(g, st)) return e;
    cur = buf[t2]; cur_dim = H; cur_buf = t2;
    if (blk >= p->rgb_from) {
      bool last = blk == p->n_blocks - 1;
      C3D_LAUNCH(skinny_linear_kernel, c3d_div_up(M, 256), 256, 0, st, cur, H, H, w->rgb_w[blk], w->rgb_b[blk], 3, rgb, 3, 0,
                                                              (int)M, rgb_started ? 1 : 0, last ? 1 : 0);
      C3D_LAUNCH_CHECK();
      rgb_started = true;
    }
  }
  if (!rgb_started) C3D_CUDA(cudaMemsetAsync(rgb, 0, (size_t)M * 3 * sizeof(float), st));  // tanh(0)
  if (hidden_out) {
    C3D_LAUNCH(copy_rows_kernel, c3d_div_up(M * H, 256), 256, 0, st, cur, hidden_out, M * H);
    C3D_LAUNCH_CHECK();
  }
  return C3D_OK;
}
