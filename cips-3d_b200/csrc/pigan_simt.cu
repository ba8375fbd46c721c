// pi-GAN renderer: ImplicitGenerator3d.forward / staged_forward with a SPATIALSIRENBASELINE or
// TALLSIREN field (piGAN_lib/generators/generators.py:26-96,110-204; piGAN_lib/siren/siren.py:97-215).
// First implementation: fp32-FMA kernels, one launch per layer (activations in a caller-provided workspace) -- the same
// building blocks as the SIMT cross-check of the CIPS-3D renderer.  The hidden width (256) x 8 layers does not fit the
// shared-memory-resident weights of ray_siren_tc.cu; a streamed-weight tcgen05 version is the follow-up.
//   per point:  x = p * (2/0.24 if gridwarp);  h = sin(f_i (W_i h + b_i) + ph_i), i = 0..7;  sigma = W_s h + b_s;
//               c = sin(f_8 (W_c [dir, h] + b_c) + ph_8);  rgb = sigmoid(W_l c + b_l)
//   per ray:    hierarchical resampling, merge, compositing of (rgb, sigma) exactly as the CIPS-3D renderer (ray_math.cuh).
#include "c3d_common.cuh"
#include "ray_math.cuh"

namespace c3d {
namespace pg {

constexpr int kLd = 3 + 256;     // activation row: [ray direction (3) | hidden (<= 256)]

// ---- tiled GEMM + FiLM-sin epilogue:  C[m, n] = sin(f[img, n] * (sum_k A[m, k] W[n, k] + b[n]) + ph[img, n])
struct FilmGemmArgs {
  const float* A; int lda;
  const float* W; int ldw;        // W (N, K) row-major (torch Linear layout)
  const float* bias;
  const float* f; const float* ph;   // (B, N)
  float* C; int ldc;
  int M, N, K, rows_per_img, img0;
};

__global__ void __launch_bounds__(256) film_gemm_kernel(const FilmGemmArgs a) {
  constexpr int TM = 64, TN = 64, TK = 16;
  __shared__ float As[TK][TM + 4];
  __shared__ float Ws[TK][TN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  const int tx = tid % 16, ty = tid / 16;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < a.K; k0 += TK) {
    for (int i = tid; i < TM * TK; i += 256) {
      const int r = i / TK, c = i % TK;
      const int gm = m0 + r, gk = k0 + c;
      As[c][r] = (gm < a.M && gk < a.K) ? a.A[(size_t)gm * a.lda + gk] : 0.f;
    }
    for (int i = tid; i < TN * TK; i += 256) {
      const int c = i % TK, n = i / TK;
      const int gk = k0 + c, gn = n0 + n;
      Ws[c][n] = (gk < a.K && gn < a.N) ? a.W[(size_t)gn * a.ldw + gk] : 0.f;
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
    const int gm = m0 + ty * 4 + i;
    if (gm >= a.M) continue;
    const int img = a.img0 + gm / a.rows_per_img;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= a.N) continue;
      const float v = acc[i][j] + a.bias[gn];
      a.C[(size_t)gm * a.ldc + gn] = sinf(fmaf(a.f[(size_t)img * a.N + gn], v, a.ph[(size_t)img * a.N + gn]));
    }
  }
}

// ---- sample points (coarse: jittered depths; fine: given depths) -> activation rows [dir | h0], depths
// one warp per point; lanes over the hidden units of layer 0
__global__ void points_h0_kernel(const C3dRayParams p, const C3dPiganWeights w, const C3dRayIO io, int lock_view, int b0, int nb,
                                 const float* __restrict__ fine_z, float* __restrict__ zbuf, float* __restrict__ actA,
                                 float* __restrict__ actB) {
  const int S = p.num_steps, N = p.n_rays, R = p.img_size, H = w.hidden;
  const long long pt = (long long)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  if (pt >= (long long)nb * N * S) return;
  const int lane = threadIdx.x % 32;
  const int s = (int)(pt % S);
  const long long rr = pt / S;
  const int n = (int)(rr % N), b = b0 + (int)(rr / N);
  const int ray = io.ray_idx ? io.ray_idx[n] : p.ray_offset + n;
  const float* M = io.cam2world + (size_t)b * 16;
  const RayFrame f = make_ray_frame(M, ray, R, p.z_cam);
  float z, px, py, pz;
  if (fine_z) {
    z = fine_z[pt];
    fine_sample(f, z, px, py, pz);
  } else {
    coarse_sample(f, M, p.ray_start, p.ray_end, S, s, io.jitter_u[((size_t)b * R * R + ray) * S + s], z, px, py, pz);
    if (lane == 0) zbuf[pt] = z;
  }
  if (lane < 3) {
    const float d = lock_view ? (lane == 2 ? -1.f : 0.f) : (lane == 0 ? f.dwx : (lane == 1 ? f.dwy : f.dwz));
    actA[(size_t)pt * kLd + lane] = d;      // the colour layer reads [dir | h] from whichever buffer holds the last layer
    actB[(size_t)pt * kLd + lane] = d;
  }
  const float sc = w.gridwarp ? 2.f / 0.24f : 1.f;
  const float x = px * sc, y = py * sc, zz = pz * sc;
  const float* W0 = w.w[0];
  for (int j = lane; j < H; j += 32) {
    const float pre = W0[j * 3 + 0] * x + W0[j * 3 + 1] * y + W0[j * 3 + 2] * zz + w.b[0][j];
    actA[(size_t)pt * kLd + 3 + j] = sinf(fmaf(w.freq[0][(size_t)b * H + j], pre, w.phase[0][(size_t)b * H + j]));
  }
}

// ---- heads: sigma = W_s h + b_s (from the last hidden layer), rgb = sigmoid(W_l c + b_l) (from the colour layer)
__global__ void heads_kernel(const float* __restrict__ h, const float* __restrict__ c, int ld, int H, const float* __restrict__ ws,
                             const float* __restrict__ bs, const float* __restrict__ wl, const float* __restrict__ bl,
                             float* __restrict__ out4, long long P) {
  const long long pt = (long long)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;   // warp per point
  if (pt >= P) return;
  const int lane = threadIdx.x % 32;
  const float* hr = h + (size_t)pt * ld;
  const float* cr = c + (size_t)pt * ld;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (int k = lane; k < H; k += 32) {
    const float cv = cr[k];
    a0 = fmaf(cv, __ldg(wl + k), a0);
    a1 = fmaf(cv, __ldg(wl + H + k), a1);
    a2 = fmaf(cv, __ldg(wl + 2 * H + k), a2);
    a3 = fmaf(hr[k], __ldg(ws + k), a3);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    a3 += __shfl_xor_sync(0xffffffffu, a3, o);
  }
  if (lane == 0) {
    float* o = out4 + (size_t)pt * 4;
    o[0] = 1.f / (1.f + expf(-(a0 + bl[0])));
    o[1] = 1.f / (1.f + expf(-(a1 + bl[1])));
    o[2] = 1.f / (1.f + expf(-(a2 + bl[2])));
    o[3] = a3 + bs[0];
  }
}

// ---- importance resampling: one thread per ray (generators.py:58-72)
__global__ void fine_z_kernel(const C3dRayParams p, const C3dRayIO io, int b0, int nb, const float* __restrict__ zbuf,
                              const float* __restrict__ coarse4, float* __restrict__ fzbuf) {
  const int S = p.num_steps, N = p.n_rays;
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= (long long)nb * N) return;
  const int n = (int)(r % N), b = b0 + (int)(r / N);
  const float* sig = coarse4 + r * S * 4 + 3;
  const float* nz = io.noise_c ? io.noise_c + ((size_t)b * N + n) * S : nullptr;
  float w[kMaxS], zz[kMaxS], u[kMaxS], fz[kMaxS];
  for (int i = 0; i < S; ++i) {
    zz[i] = zbuf[r * S + i];
    u[i] = io.pdf_u[((size_t)b * N + n) * S + i];
  }
  const float ns = p.noise_std;
  integrate_weights(
      S, p.clamp_mode, [&](int i) { return zz[i]; }, [&](int i) { return sig[(size_t)i * 4]; },
      [&](int i) { return nz ? __fmul_rn(nz[i], ns) : 0.f; }, [&](int i, float v) { w[i] = v; });
  sample_pdf_ray(S, w, zz, u, fz);
  for (int i = 0; i < S; ++i) fzbuf[r * S + i] = fz[i];
}

// ---- merge + final integration of (rgb, sigma): one thread per ray (generators.py:77-93)
__global__ void composite_kernel(const C3dRayParams p, const C3dRayIO io, int b0, int nb, const float* __restrict__ zbuf,
                                 const float* __restrict__ coarse4, const float* __restrict__ fzbuf, const float* __restrict__ fine4) {
  const int S = p.num_steps, N = p.n_rays;
  const int nS = p.hierarchical ? 2 * S : S;
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= (long long)nb * N) return;
  const int n = (int)(r % N), b = b0 + (int)(r / N);
  const size_t ro = (size_t)b * N + n;
  float key[kMaxNS];
  int idx[kMaxNS];
  if (p.hierarchical) {
    for (int i = 0; i < S; ++i) { key[i] = fzbuf[r * S + i]; idx[i] = i; }
    for (int i = 0; i < S; ++i) { key[S + i] = zbuf[r * S + i]; idx[S + i] = S + i; }
    sort_keys(nS, key, idx);
  } else {
    for (int i = 0; i < S; ++i) { key[i] = zbuf[r * S + i]; idx[i] = S + i; }
  }
  auto row = [&](int i) -> const float* {
    const int id = idx[i];
    return id < S ? fine4 + ((size_t)r * S + id) * 4 : coarse4 + ((size_t)r * S + (id - S)) * 4;
  };
  const float* nz = io.noise_f ? io.noise_f + ro * nS : nullptr;
  const float ns = p.noise_std;
  float w[kMaxNS];
  const float wsum = integrate_weights(
      nS, p.clamp_mode, [&](int i) { return key[i]; }, [&](int i) { return row(i)[3]; },
      [&](int i) { return nz ? __fmul_rn(nz[i], ns) : 0.f; }, [&](int i, float v) { w[i] = v; });
  if (p.last_back) w[nS - 1] += 1.f - wsum;
  float depth = 0.f, acc[3] = {0.f, 0.f, 0.f};
  for (int i = 0; i < nS; ++i) {
    const float* f = row(i);
    for (int c = 0; c < 3; ++c) acc[c] = fmaf(w[i], f[c], acc[c]);
    depth = fmaf(w[i], key[i], depth);
  }
  for (int c = 0; c < 3; ++c) io.pixels_fea[ro * 3 + c] = acc[c] + (p.white_back ? 1.f - wsum : 0.f);
  if (io.depth) io.depth[ro] = depth;
  if (io.weights) for (int i = 0; i < nS; ++i) io.weights[ro * nS + i] = w[i];
  if (io.dbg_all_z) for (int i = 0; i < nS; ++i) io.dbg_all_z[ro * nS + i] = key[i];
}

__global__ void copy_kernel(const float* src, float* dst, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

}  // namespace pg
}  // namespace c3d

using namespace c3d;
using namespace c3d::pg;

static long long pg_chunk_images(const C3dRayParams* p) {
  const long long budget_pts = 1ll << 20;     // ~2.2 GB of intermediates
  const long long per_img = (long long)p->n_rays * p->num_steps;
  long long nb = budget_pts / (per_img > 0 ? per_img : 1);
  if (nb < 1) nb = 1;
  if (nb > p->batch) nb = p->batch;
  return nb;
}

// tcgen05 form (pigan_tc.cu): C3D_PIGAN_IMPL=tc, emulation-verified, opt-in until it has run on hardware
size_t c3d_pigan_tc_workspace_bytes(const C3dRayParams* p);
bool c3d_pigan_tc_supported(const C3dRayParams* p, const C3dPiganWeights* w);
int c3d_pigan_render_fwd_tc(const C3dRayParams* p, const C3dPiganWeights* w, const C3dRayIO* io, int lock_view, void* workspace,
                            size_t workspace_bytes, cudaStream_t st);
static bool pg_want_tc() { return c3d_options().pigan_tc != 0; }

extern "C" size_t c3d_pigan_workspace_bytes(const C3dRayParams* p) {
  if (!p) return 0;
  const long long P = pg_chunk_images(p) * p->n_rays * p->num_steps;
  const size_t simt = (size_t)P * (1 + 1 + 2 * kLd + 4 + 4) * sizeof(float);      // z, fine z, two activation buffers, coarse4, fine4
  const size_t tc = c3d_pigan_tc_workspace_bytes(p);
  return simt > tc ? simt : tc;
}

// evaluate the field on the P points whose [dir | h0] rows are in actA; result rows (rgb, sigma) -> out4
static int pg_field(const C3dPiganWeights& w, long long P, int rows_per_img, int b0, float* actA, float* actB, float* out4,
                    cudaStream_t st) {
  const int H = w.hidden;
  float* cur = actA;
  float* nxt = actB;
  FilmGemmArgs g;
  g.M = (int)P; g.rows_per_img = rows_per_img; g.img0 = b0;
  for (int l = 1; l < w.n_layers; ++l) {
    g.A = cur + 3; g.lda = kLd; g.W = w.w[l]; g.ldw = H; g.bias = w.b[l]; g.f = w.freq[l]; g.ph = w.phase[l];
    g.C = nxt + 3; g.ldc = kLd; g.N = H; g.K = H;
    C3D_LAUNCH(film_gemm_kernel, dim3(c3d_div_up(P, 64), c3d_div_up(H, 64)), 256, 0, st, g);
    C3D_LAUNCH_CHECK();
    float* t = cur; cur = nxt; nxt = t;
  }
  // colour layer: input [dir | h] (K = H + 3), FiLM parameters of slot n_layers
  g.A = cur; g.lda = kLd; g.W = w.wc; g.ldw = H + 3; g.bias = w.bc; g.f = w.freq[w.n_layers]; g.ph = w.phase[w.n_layers];
  g.C = nxt + 3; g.ldc = kLd; g.N = H; g.K = H + 3;
  C3D_LAUNCH(film_gemm_kernel, dim3(c3d_div_up(P, 64), c3d_div_up(H, 64)), 256, 0, st, g);
  C3D_LAUNCH_CHECK();
  C3D_LAUNCH(heads_kernel, c3d_div_up(P, 8), 256, 0, st, cur + 3, nxt + 3, kLd, H, w.w_sigma, w.b_sigma, w.wl, w.bl, out4, P);
  C3D_LAUNCH_CHECK();
  return C3D_OK;
}

extern "C" int c3d_pigan_render_fwd(const C3dRayParams* p, const C3dPiganWeights* w, const C3dRayIO* io, int32_t lock_view,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  C3D_CHECK_ARG(p && w && io, "pigan_render: null struct pointer");
  C3D_CHECK_ARG(p->batch >= 0 && p->img_size >= 1 && p->n_rays >= 0, "pigan_render: bad sizes");
  C3D_CHECK_ARG(p->num_steps >= 3 && p->num_steps <= kMaxS, "pigan_render: num_steps must be in [3,32], got %d", p->num_steps);
  C3D_CHECK_ARG((long long)p->n_rays <= (long long)p->img_size * p->img_size, "pigan_render: n_rays > img_size^2");
  C3D_CHECK_ARG(io->ray_idx || (long long)p->ray_offset + p->n_rays <= (long long)p->img_size * p->img_size,
                "pigan_render: ray_offset + n_rays exceeds the image");
  C3D_CHECK_ARG(p->clamp_mode == 0 || p->clamp_mode == 1, "pigan_render: clamp_mode must be 0 (relu) or 1 (softplus)");
  C3D_CHECK_ARG(io->cam2world && io->jitter_u && io->pixels_fea, "pigan_render: null cam2world/jitter_u/pixels_fea");
  C3D_CHECK_ARG(!p->hierarchical || io->pdf_u, "pigan_render: hierarchical sampling needs pdf_u");
  C3D_CHECK_ARG(w->n_layers >= 1 && w->n_layers <= C3D_PIGAN_MAX_LAYERS && w->hidden >= 1 && w->hidden <= 256,
                "pigan_render: n_layers in [1,8], hidden in [1,256]");
  for (int l = 0; l < w->n_layers; ++l)
    C3D_CHECK_ARG(w->w[l] && w->b[l] && w->freq[l] && w->phase[l], "pigan_render: null weights / FiLM parameters of layer %d", l);
  C3D_CHECK_ARG(w->freq[w->n_layers] && w->phase[w->n_layers] && w->wc && w->bc && w->wl && w->bl && w->w_sigma && w->b_sigma,
                "pigan_render: null head weights");
  if (p->batch == 0 || p->n_rays == 0) return C3D_OK;
  C3D_CHECK_ARG(workspace, "pigan_render: null workspace");
  if (workspace_bytes < c3d_pigan_workspace_bytes(p)) {
    c3d_set_error("pigan_render: workspace too small");
    return C3D_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (pg_want_tc() && c3d_pigan_tc_supported(p, w)) return c3d_pigan_render_fwd_tc(p, w, io, lock_view, workspace, workspace_bytes, st);
  const int S = p->num_steps, N = p->n_rays;
  const long long cb = pg_chunk_images(p), Pmax = cb * N * S;
  float* zbuf = (float*)workspace;
  float* fzbuf = zbuf + Pmax;
  float* actA = fzbuf + Pmax;
  float* actB = actA + Pmax * kLd;
  float* c4 = actB + Pmax * kLd;
  float* f4 = c4 + Pmax * 4;
  for (int b0 = 0; b0 < p->batch; b0 += (int)cb) {
    const int nb = (int)((p->batch - b0) < cb ? (p->batch - b0) : cb);
    const long long P = (long long)nb * N * S, NR = (long long)nb * N;
    C3D_LAUNCH(points_h0_kernel, c3d_div_up(P, 8), 256, 0, st, *p, *w, *io, lock_view, b0, nb, (const float*)nullptr, zbuf, actA, actB);
    C3D_LAUNCH_CHECK();
    if (int e = pg_field(*w, P, N * S, b0, actA, actB, c4, st)) return e;
    if (io->dbg_coarse) {
      C3D_LAUNCH(copy_kernel, c3d_div_up(P * 4, 256), 256, 0, st, c4, io->dbg_coarse + (size_t)b0 * N * S * 4, P * 4);
      C3D_LAUNCH_CHECK();
    }
    if (p->hierarchical) {
      C3D_LAUNCH(pg::fine_z_kernel, c3d_div_up(NR, 128), 128, 0, st, *p, *io, b0, nb, zbuf, c4, fzbuf);
      C3D_LAUNCH_CHECK();
      C3D_LAUNCH(points_h0_kernel, c3d_div_up(P, 8), 256, 0, st, *p, *w, *io, lock_view, b0, nb, (const float*)fzbuf, zbuf, actA, actB);
      C3D_LAUNCH_CHECK();
      if (int e = pg_field(*w, P, N * S, b0, actA, actB, f4, st)) return e;
      if (io->dbg_fine) {
        C3D_LAUNCH(copy_kernel, c3d_div_up(P * 4, 256), 256, 0, st, f4, io->dbg_fine + (size_t)b0 * N * S * 4, P * 4);
        C3D_LAUNCH_CHECK();
      }
    }
    C3D_LAUNCH(pg::composite_kernel, c3d_div_up(NR, 128), 128, 0, st, *p, *io, b0, nb, zbuf, c4, fzbuf, f4);
    C3D_LAUNCH_CHECK();
  }
  return C3D_OK;
}
