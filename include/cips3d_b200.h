/*
 * cips3d_b200 -- C-ABI of the B200-native CIPS-3D hot path (libcips3d_b200.so).
 *
 * The reference has no FFI for this path: its boundary is the Python nn.Module surface
 * (exp/cips3d/models/generator.py:1158, exp/cips3d/models/discriminator.py:588) plus two
 * pybind11 ops used by the discriminator.  Every entry point below names the reference
 * code it replaces.  Conventions (all entry points):
 *   - plain C: device pointers + sizes, no torch types; the CALLER owns every buffer
 *     (outputs and workspaces are allocated by the caller, e.g. with torch.empty);
 *   - all tensors are fp32, contiguous, row-major, in the layouts documented per argument;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*), nothing synchronises,
 *     no hidden allocation, no global state -> safe to call from several host threads;
 *   - returns 0 on success, a negative C3D_E* code otherwise; c3d_last_error() returns a
 *     thread-local message for the last failure on the calling thread.
 */
#ifndef CIPS3D_B200_H_
#define CIPS3D_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define C3D_OK 0
#define C3D_EINVAL (-1)   /* bad argument (null pointer, unsupported size)            */
#define C3D_ECUDA (-2)    /* a CUDA runtime call / kernel launch failed               */
#define C3D_EARCH (-3)    /* device is not sm_100 (the tensor-core path needs tcgen05) */
#define C3D_EWORKSPACE (-4) /* workspace too small                                    */

/* Which kernel generation executes the GEMM-shaped parts.  Both are sm_100a CUDA.
 *   C3D_IMPL_TC   : fused tcgen05/TMEM kernels (the product path, default)
 *   C3D_IMPL_SIMT : unfused fp32-FMA kernels; kept as an on-device cross-check only */
#define C3D_IMPL_TC 0
#define C3D_IMPL_SIMT 1

int c3d_version(void);
const char* c3d_last_error(void);
/* 1 if device `dev` is compute capability 10.x */
int c3d_device_supported(int dev);
/* number of kernels this library has launched in this process (monotonic; bench.py's
 * gpu_launches is the difference across the timed region) */
unsigned long long c3d_launch_count(void);

/* ------------------------------------------------------------------------------------
 * Fused volumetric renderer: rays -> FiLM-SIREN -> hierarchical resampling -> composite.
 * Replaces (reference, paths relative to the checkout):
 *   exp/comm/comm_utils.py:365-412 get_initial_rays_trig, :416-438 perturb_points,
 *   :584-679 transform_sampled_points (point/direction transform part),
 *   exp/cips3d/models/generator.py:260-317 NeRFNetwork.forward_with_frequencies_phase_shifts,
 *   exp/comm/models/film_layer.py:78-107 FiLMLayer.forward,
 *   exp/dev/nerf_inr/models/generator_nerf_inr.py:537-598 get_fine_points_and_direction,
 *   exp/pigan/pigan_utils.py:164-209 sample_pdf, :212-273 fancy_integration,
 *   exp/cips3d/models/generator.py:1733-1738 (merge/sort of coarse+fine samples).
 * Random numbers are NOT generated here: the caller draws them with torch in the
 * reference's order (hard part 4) and passes the tensors.
 * ---------------------------------------------------------------------------------- */
typedef struct C3dRayParams {
  int32_t batch;        /* B images                                                     */
  int32_t img_size;     /* R: image is R x R, ray index = h*R + w (comm_utils.py:392)   */
  int32_t num_steps;    /* S coarse samples per ray, 2 <= S <= 32                       */
  int32_t n_rays;       /* N rays rendered per image (N <= R*R)                         */
  int32_t ray_offset;   /* when ray_idx == NULL: local ray i is global ray ray_offset+i */
  int32_t hierarchical; /* 1: S more importance samples, merged and sorted (2S total)   */
  int32_t clamp_mode;   /* 0 relu, 1 softplus (pigan_utils.py:248-253)                  */
  int32_t white_back;   /* pigan_utils.py:265-266                                       */
  int32_t last_back;    /* pigan_utils.py:259-260                                       */
  int32_t impl;         /* C3D_IMPL_*                                                   */
  float z_cam;          /* -1/tan(fov*pi/360) (comm_utils.py:396)                       */
  float ray_start, ray_end;
  float noise_std;      /* nerf_noise; scales noise_c / noise_f (pigan_utils.py:246)    */
} C3dRayParams;

typedef struct C3dSirenWeights { /* NeRFNetwork parameters, torch Linear layout (out,in) */
  const float* w0; const float* b0;         /* network.0.linear            (128,3) (128) */
  const float* w1; const float* b1;         /* network.1.linear          (128,128) (128) */
  const float* w_sigma; const float* b_sigma; /* final_layer               (1,128) (1)   */
  const float* wc; const float* bc;         /* color_layer_sine.linear    (64,128) (64)  */
  const float* wl; const float* bl;         /* color_layer_linear.0        (32,64) (32)  */
  /* per-image FiLM parameters, gamma = 15*gain_fc(w)+30, beta = bias_fc(w)
   * (film_layer.py:59,89-91); computed by the caller (B x style GEMV, not hot).      */
  const float* gamma0; const float* beta0;  /* (B,128) */
  const float* gamma1; const float* beta1;  /* (B,128) */
  const float* gammac; const float* betac;  /* (B,64)  */
} C3dSirenWeights;

typedef struct C3dRayIO {
  const float* cam2world;  /* (B,4,4) row-major, create_cam2world_matrix comm_utils.py:538 */
  const int32_t* ray_idx;  /* (N) global ray indices or NULL (gather_points comm_utils.py:264) */
  const float* jitter_u;   /* (B,R*R,S) U[0,1), indexed by GLOBAL ray (comm_utils.py:432)  */
  const float* noise_c;    /* (B,N,S)  N(0,1) or NULL -> 0  (coarse-pass integration)      */
  const float* pdf_u;      /* (B*N,S)  U[0,1)  (pigan_utils.py:192); NULL iff !hierarchical */
  const float* noise_f;    /* (B,N,nS) N(0,1) or NULL -> 0, nS = hierarchical ? 2S : S     */
  float* pixels_fea;       /* out (B,N,32)  integrated feature (rgb_final)                 */
  float* depth;            /* out (B,N) or NULL                                            */
  float* weights;          /* out (B,N,nS) or NULL                                         */
  float* dbg_coarse;       /* out (B,N,S,33) [feature32, sigma] or NULL (parity tests)     */
  float* dbg_fine;         /* out (B,N,S,33) or NULL                                       */
  float* dbg_all_z;        /* out (B,N,nS) sorted sample depths or NULL                    */
} C3dRayIO;

size_t c3d_ray_siren_workspace_bytes(const C3dRayParams* p);
int c3d_ray_siren_fwd(const C3dRayParams* p, const C3dSirenWeights* w, const C3dRayIO* io,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Fused per-pixel CIPS synthesis MLP.
 * Replaces exp/cips3d/models/generator.py:1107-1154 CIPSNet.forward (+ SinBlock :949-974,
 * ToRGB :983-1006) and exp/comm/models/mod_conv_fc.py:452-496 SinStyleMod.forward_bmm,
 * using  y = ((x * (s+1)) @ W) * d,  d[b,o] = rsqrt(sum_i (W[i,o](s[b,i]+1))^2 + 1e-8)
 * so one W per layer serves every image.  s+1 and d are per-image vectors computed by
 * the caller (B x 512 GEMMs, not hot).
 * ---------------------------------------------------------------------------------- */
#define C3D_CIPS_MAX_LAYERS 18
typedef struct C3dCipsParams {
  int32_t batch;        /* B                                                    */
  int32_t n_pix;        /* N pixels per image                                   */
  int32_t in_dim;       /* 32                                                   */
  int32_t hidden;       /* 512                                                  */
  int32_t n_blocks;     /* blocks that run (9; fewer if img_size stops early)   */
  int32_t skip_from;    /* block idx >= skip_from adds the residual (4)         */
  int32_t rgb_from;     /* block idx >= rgb_from accumulates ToRGB (3)          */
  int32_t impl;         /* C3D_IMPL_*                                           */
} C3dCipsParams;

typedef struct C3dCipsWeights {
  const float* w[C3D_CIPS_MAX_LAYERS];      /* SinStyleMod.weight[0], (in,out) row-major   */
  const float* style1p[C3D_CIPS_MAX_LAYERS]; /* (B,in)  modulation(w_inr) + 1               */
  const float* demod[C3D_CIPS_MAX_LAYERS];  /* (B,out) rsqrt(sum_i (W s1p)^2 + eps)        */
  const float* rgb_w[C3D_CIPS_MAX_LAYERS / 2]; /* ToRGB.linear.weight (3,hidden) per block  */
  const float* rgb_b[C3D_CIPS_MAX_LAYERS / 2]; /* ToRGB.linear.bias (3)                     */
} C3dCipsWeights;

size_t c3d_cips_workspace_bytes(const C3dCipsParams* p);
/* x (B,N,in_dim) -> rgb (B,N,3) = tanh(sum of ToRGB skips); hidden_out (B,N,hidden) or NULL.
 * With hidden_out == NULL (the image alone: what GeneratorNerfINR.forward asks for) the tensor-core kernel keeps the skip connections'
 * stream between layers as fp16 -- the very values the next layer's MMAs read -- instead of fp32; the image's error against an fp64
 * evaluation is unchanged by that (tests), it halves the L2 traffic of the residual layers.  C3D_CIPS_RES16=0 keeps fp32 always. */
int c3d_cips_fwd(const C3dCipsParams* p, const C3dCipsWeights* w, const float* x, float* rgb,
                 float* hidden_out, void* workspace, size_t workspace_bytes, void* stream);

/* Per-image vectors of c3d_cips_fwd for all layers in one launch (the module computes them with ~7 torch launches per layer):
 * s1p[l] (B,in_l) = modulation_l(style_l) + 1, demod[l] (B,hidden) = rsqrt(sum_k (W_l[k][n] s1p[l][b][k])^2 + eps)
 * (SinStyleMod, exp/comm/models/mod_conv_fc.py:452-496).  hidden = 512. */
typedef struct C3dStylePrep {
  const float* style[C3D_CIPS_MAX_LAYERS];  /* (B, style_dim) style vector of layer l                       */
  const float* mod_w[C3D_CIPS_MAX_LAYERS];  /* modulation.weight (in_l, style_dim)                          */
  const float* mod_b[C3D_CIPS_MAX_LAYERS];  /* modulation.bias (in_l)                                       */
  const float* w[C3D_CIPS_MAX_LAYERS];      /* SinStyleMod.weight[0] (in_l, hidden)                         */
  float* s1p[C3D_CIPS_MAX_LAYERS];          /* out (B, in_l)                                                */
  float* demod[C3D_CIPS_MAX_LAYERS];        /* out (B, hidden)                                              */
  int32_t in_dim[C3D_CIPS_MAX_LAYERS];
  int32_t n_layers, style_dim;
  float eps;                                /* 1e-8                                                         */
} C3dStylePrep;
int c3d_cips_style_prep(const C3dStylePrep* a, int32_t batch, void* stream);

/* Training forward and the backward CHAIN of the fused CIPS MLP (first part).
 *   c3d_cips_fwd_train: c3d_cips_fwd that also writes every layer's output y_l (after LeakyReLU / residual) as the fp16
 *       values the next layer consumed: acts_f16 (n_layers, B, N, hidden), and the sign bits of z_l for the layers that add a
 *       residual (where y_l - y_{l-2} of two fp16 stashes cannot recover them): zsign_u16 (n_layers, B, N, hidden/16).
 *   c3d_cips_bwd: given g = dL/d(rgb before tanh) (B,N,3), PRE-MULTIPLIED by a scale S the caller chooses so that fp16
 *       gradients neither underflow nor overflow, runs dZ_l = dL/dy_l * lrelu'(z_l), dX_l = dZ_l W''_l^T for l = L-1..0 on
 *       chip and writes dz_f16 (n_layers, B, N, hidden) = S dZ_l and dx (B,N,in_dim) = S dL/dx (or NULL).
 *   The weight gradients are plain K = pixels GEMMs over the two stashes, dW''_l[b] = X_l[b]^T dZ_l[b] (X_0 = x, X_l = y_{l-1}),
 *   left to the library; cips3d_b200/ops.py folds them into dW, d style1p, d demod and the ToRGB gradients. */
int c3d_cips_fwd_train(const C3dCipsParams* p, const C3dCipsWeights* w, const float* x, float* rgb, void* acts_f16,
                       void* zsign_u16, void* workspace, size_t workspace_bytes, void* stream);
size_t c3d_cips_bwd_workspace_bytes(const C3dCipsParams* p);
int c3d_cips_bwd(const C3dCipsParams* p, const C3dCipsWeights* w, const void* acts_f16, const void* zsign_u16,
                 const float* g_rgb_pre, void* dz_f16, float* dx, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * FiLM + sine of the NeRF branch's AUTOGRAD graph (the torch training path of configs 1-3):
 *   y = sin(gain * z + bias)     exp/cips3d/models/film_layer.py:94-107, piGAN_lib/siren/siren.py:83-94
 * z, y, dy, dz: (batch, points, channels) fp32 contiguous; gain, bias, dgain, dbias: (batch, channels).
 * Forward is one pass; backward is one pass that also produces the per-image reductions
 *   dz = dy * cos(gain * z + bias) * gain,  dbias = sum_p dy * cos(.),  dgain = sum_p dy * cos(.) * z
 * (deterministic: fixed-order partial sums through `workspace`).  channels: a multiple of 4 with 256 % (channels / 4) == 0.
 * All pointers 16-byte aligned.  Only z has to be kept for the backward.
 * ---------------------------------------------------------------------------------- */
int c3d_film_sin_fwd(const float* z, const float* gain, const float* bias, float* y, int32_t batch, int64_t points,
                     int32_t channels, void* stream);
size_t c3d_film_sin_bwd_workspace_bytes(int32_t batch, int64_t points, int32_t channels);
int c3d_film_sin_bwd(const float* z, const float* gain, const float* bias, const float* dy, float* dz, float* dgain,
                     float* dbias, int32_t batch, int64_t points, int32_t channels, void* workspace, size_t workspace_bytes,
                     void* stream);

/* ------------------------------------------------------------------------------------
 * Volume integration of the NeRF branch's AUTOGRAD graph (SURVEY section 8(a) R14-R15 as a differentiable op; section 8(f)
 * rank 1, step (B) of the ray-SIREN backward): fancy_integration, exp/pigan/pigan_utils.py:212-273, as called from
 * exp/cips3d/models/generator.py:1744-1752 and generator_nerf_inr.py:537-598 (coarse pass, no_grad).
 *   rgb_sigma (rays, samples, channels + 1) fp32, sigma in the last column, samples SORTED by depth; z (rays, samples);
 *   noise (rays, samples) already multiplied by nerf_noise, or NULL; softplus: 0 = relu clamp, 1 = softplus clamp.
 *   fwd: fea (rays, channels) = sum_i w_i c_i (+ 1 - sum w with white_back); weights (rays, samples) or NULL (w after last_back).
 *   bwd: d_rgb_sigma (rays, samples, channels + 1) from d_fea (rays, channels); nothing but the inputs is kept for it.
 * samples <= 32, channels <= 128.  z and noise get no gradient (the reference draws the fine depths under no_grad).
 * ---------------------------------------------------------------------------------- */
int c3d_integrate_fwd(const float* rgb_sigma, const float* z, const float* noise, float* fea, float* weights, int64_t rays,
                      int32_t samples, int32_t channels, int32_t softplus, int32_t last_back, int32_t white_back, void* stream);
int c3d_integrate_bwd(const float* rgb_sigma, const float* z, const float* noise, const float* d_fea, float* d_rgb_sigma,
                      int64_t rays, int32_t samples, int32_t channels, int32_t softplus, int32_t last_back, int32_t white_back,
                      void* stream);
/* Merged form: exp/cips3d/models/generator.py:1733-1752 -- torch.cat([fine, coarse]) + torch.sort of the depths + torch.gather +
 * fancy_integration -- without the concatenated / gathered copies.  fine, coarse: (rays, samples_each, channels + 1) as the field
 * produced them (unsorted relative to each other); z_fine, z_coarse: (rays, samples_each).  The 2 * samples_each depths of a ray are
 * sorted in registers (stable: on equal depths the fine sample, first in cat order, comes first); noise (rays, 2 * samples_each)
 * is indexed by SORTED position as in the reference; weights / z_sorted (rays, 2 * samples_each, sorted order) may be NULL.
 * bwd writes d_fine / d_coarse at the source rows (every element is written).  samples_each <= 16, channels <= 128.          */
int c3d_integrate_merge_fwd(const float* fine, const float* z_fine, const float* coarse, const float* z_coarse, const float* noise,
                            float* fea, float* weights, float* z_sorted, int64_t rays, int32_t samples_each, int32_t channels,
                            int32_t softplus, int32_t last_back, int32_t white_back, void* stream);
int c3d_integrate_merge_bwd(const float* fine, const float* z_fine, const float* coarse, const float* z_coarse, const float* noise,
                            const float* d_fea, float* d_fine, float* d_coarse, int64_t rays, int32_t samples_each, int32_t channels,
                            int32_t softplus, int32_t last_back, int32_t white_back, void* stream);

/* sample_pdf, exp/pigan/pigan_utils.py:164-209 (called under no_grad from generator_nerf_inr.py:570-581), for given uniforms:
 * bins (rays, n_weights + 1), weights (rays, n_weights), u (rays, n_importance) -> samples (rays, n_importance);
 * pdf = (w + eps) / sum, cdf = [0, cumsum(pdf)], searchsorted (left), linear inverse CDF with denom < eps -> 1.
 * n_weights <= 32.  The caller draws u (torch.rand, or linspace in the reference's det mode).                              */
int c3d_sample_pdf(const float* bins, const float* weights, const float* u, float* samples, int64_t rays, int32_t n_weights,
                   int32_t n_importance, float eps, void* stream);

/* ------------------------------------------------------------------------------------
 * Image export for the inference paths: generator output (batch, channels, height, width)
 * fp32 -> (batch, height, width, channels) uint8 on the device, bit-identical to what the reference's scripts hand to
 * PIL, so an inference batch leaves the GPU as 1 byte per sample.  mode:
 *   C3D_U8_SAVE_IMAGE    torchvision save_image(img, path, normalize=True, value_range=(lo, hi)) on ONE image per file --
 *                        exp/cips3d/scripts/gen_images.py:64, sample_images.py:73:
 *                        y = (clamp(x, lo, hi) - lo) / max(hi - lo, 1e-5); u8 = trunc(clamp(y * 255 + 0.5, 0, 255)).
 *                        The division is a true fp32 division (torch CPU); torch CUDA multiplies by the fp32 reciprocal,
 *                        identical for power-of-two ranges such as the reference's (-1, 1).
 *   C3D_U8_TENSOR_TO_PIL exp/cips3d/models/st_web.py:44-46: y = x * 0.5 + 0.5; u8 = trunc(clamp(y * 255 + 0.5, 0, 255))
 *   C3D_U8_TO_PIL        exp/comm/comm_utils.py:21-24 + torchvision to_pil_image: u8 = trunc(((x + 1) * 0.5) * 255);
 *                        inputs outside [-1, 1] saturate here (numpy's float -> uint8 cast is undefined there)
 * channels_last != 0: img is already (batch, height, width, channels) in memory -- the layout the CIPS kernel writes, of
 * which the generator returns an NCHW view -- and the conversion is a flat stream.
 * lo / hi are read by mode 0 only.  NaN -> 0.  channels 1..4.  Empty batches / images are a no-op.
 * ---------------------------------------------------------------------------------- */
enum { C3D_U8_SAVE_IMAGE = 0, C3D_U8_TENSOR_TO_PIL = 1, C3D_U8_TO_PIL = 2 };
int c3d_image_to_u8(const float* img, uint8_t* out, int32_t batch, int32_t channels, int32_t height, int32_t width,
                    int32_t channels_last, int32_t mode, double lo, double hi, void* stream);

/* ------------------------------------------------------------------------------------
 * Discriminator ops.  Same semantics as the reference's two pybind11 modules:
 *   fused.fused_bias_act(input, bias, refer, act, grad, alpha, scale)
 *       exp/comm/op/fused_bias_act.cpp:11-20, fused_bias_act_kernel.cu:19-50
 *   upfirdn2d.upfirdn2d(input(N,H,W,1), kernel, up_x, up_y, down_x, down_y, pad_x0..pad_y1)
 *       exp/comm/op/upfirdn2d.cpp:12-22, upfirdn2d_kernel.cu:52-139
 * ---------------------------------------------------------------------------------- */
/* y[i] = f(x[i] + b[(i / step_b) % size_b]) * scale;  act: 1 linear, 3 leaky-relu;
 * grad: 0 forward, 1 first derivative gated on ref (= saved output), 2 -> 0.
 * bias / ref may be NULL. */
int c3d_bias_act(const float* x, const float* bias, const float* ref, float* y, int64_t size_x,
                 int32_t step_b, int32_t size_b, int32_t act, int32_t grad, float alpha,
                 float scale, void* stream);

/* x: (planes,in_h,in_w) -> y: (planes,out_h,out_w),
 * out = (in*up + pad0 + pad1 - k) / down + 1; zero-upsample, pad (negative = crop),
 * correlate with the flipped kernel (kh,kw), decimate. */
int c3d_upfirdn2d(const float* x, const float* kernel, float* y, int32_t planes, int32_t in_h,
                  int32_t in_w, int32_t kh, int32_t kw, int32_t up_x, int32_t up_y,
                  int32_t down_x, int32_t down_y, int32_t pad_x0, int32_t pad_x1, int32_t pad_y0,
                  int32_t pad_y1, void* stream);

/* ------------------------------------------------------------------------------------
 * pi-GAN renderer: rays -> FiLM-SIREN (hidden 256 x 8, view-dependent colour) ->
 * hierarchical resampling -> compositing of (rgb, sigma).  Replaces
 *   piGAN_lib/generators/generators.py:26-96 ImplicitGenerator3d.forward, :110-204 staged_forward,
 *   piGAN_lib/siren/siren.py:133-152 TALLSIREN / :196-215 SPATIALSIRENBASELINE .forward_with_frequencies_phase_shifts,
 *   piGAN_lib/generators/volumetric_rendering.py (rays, perturbation, camera transform, sample_pdf, fancy_integration).
 * Uses C3dRayParams / C3dRayIO with 4-channel samples: io->pixels_fea is (B,N,3) = integrated rgb in [0,1] (the caller
 * applies *2-1), dbg_coarse / dbg_fine are (B,N,S,4) = [rgb(3), sigma].  Random draws come from the caller as above.
 * ---------------------------------------------------------------------------------- */
#define C3D_PIGAN_MAX_LAYERS 8
typedef struct C3dPiganWeights {
  const float* w[C3D_PIGAN_MAX_LAYERS];        /* network.<i>.layer.weight (hidden, in_i), in_0 = 3         */
  const float* b[C3D_PIGAN_MAX_LAYERS];        /* network.<i>.layer.bias (hidden)                           */
  const float* freq[C3D_PIGAN_MAX_LAYERS + 1];  /* (B,hidden) per FiLM layer: frequencies*15+30 (siren.py:134, 197); slot n_layers = colour layer */
  const float* phase[C3D_PIGAN_MAX_LAYERS + 1]; /* (B,hidden) phase shifts                                   */
  const float* w_sigma; const float* b_sigma;  /* final_layer (1,hidden) (1)                                */
  const float* wc; const float* bc;            /* color_layer_sine.layer (hidden, hidden+3) (hidden): input = [ray direction, h] */
  const float* wl; const float* bl;            /* color_layer_linear.0 (3,hidden) (3); sigmoid applied here */
  int32_t n_layers;                            /* 8                                                         */
  int32_t hidden;                              /* 256                                                       */
  int32_t gridwarp;                            /* 1: SPATIALSIRENBASELINE's UniformBoxWarp(0.24), 0: TALLSIREN */
} C3dPiganWeights;

size_t c3d_pigan_workspace_bytes(const C3dRayParams* p);
/* lock_view: ray directions fed to the colour layer are (0,0,-1) (generators.py:43-45) */
int c3d_pigan_render_fwd(const C3dRayParams* p, const C3dPiganWeights* w, const C3dRayIO* io, int32_t lock_view,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Optimiser tail of the training step, multi-tensor, HBM-bound.
 * Replaces, per optimiser (exp/cips3d/scripts/train.py:417-438 for D, :468-491 for G):
 *   torch.nn.utils.clip_grad_norm_(params, grad_clip)         -> c3d_grad_norm
 *   torch.optim.Adam(..., weight_decay=0).step() [+ zero_grad] -> c3d_adam_ema_step
 *   comm_model_utils.EMA.update (exp/comm/comm_model_utils.py:99-121, target = target*decay + source*(1-decay))
 *                                                              -> fused into c3d_adam_ema_step; c3d_ema_update for
 *                                                                 state_dict entries without optimiser state
 * The tensor table is a HOST array of device pointers (it travels as a kernel argument).
 * ---------------------------------------------------------------------------------- */
#define C3D_OPT_MAX_TENSORS 160   /* per launch; longer tables are processed in groups */
typedef struct C3dOptTensor {
  float* param;       /* (n) parameter                                                   */
  float* grad;        /* (n) gradient (written only when zero_grad != 0)                  */
  float* exp_avg;     /* (n) Adam first moment  (torch state 'exp_avg')                   */
  float* exp_avg_sq;  /* (n) Adam second moment (torch state 'exp_avg_sq')                */
  float* ema;         /* (n) EMA copy of the parameter or NULL                            */
  int64_t n;
} C3dOptTensor;

typedef struct C3dAdamParams {   /* doubles: torch holds them as Python floats */
  double lr, beta1, beta2, eps;
  double ema_decay;   /* < 0: skip the EMA update this step (itr < start_itr)            */
  int64_t step;       /* the step count AFTER this update's increment (>= 1)             */
  int32_t zero_grad;  /* 1: also write zeros into the gradients                          */
} C3dAdamParams;

size_t c3d_optim_workspace_bytes(void);
/* norm_out[0] = || all gradients ||_2 ; norm_out[1] = min(1, max_norm / (norm + 1e-6))  (1 if max_norm <= 0).
 * norm_out: 2 device floats.  Deterministic (fixed summation order for a given device). */
int c3d_grad_norm(const C3dOptTensor* tensors, int32_t n_tensors, float max_norm, float* norm_out,
                  void* workspace, size_t workspace_bytes, void* stream);
/* clip_coef: device pointer to the clip coefficient (norm_out + 1) or NULL (no clipping). */
int c3d_adam_ema_step(const C3dOptTensor* tensors, int32_t n_tensors, const C3dAdamParams* hp,
                      const float* clip_coef, void* stream);
/* ema = ema * decay + param * (1 - decay) for every tensor (param, ema, n used). */
int c3d_ema_update(const C3dOptTensor* tensors, int32_t n_tensors, double decay, void* stream);

/* ------------------------------------------------------------------------------------
 * Self-test of the tcgen05 building block (UMMA descriptors, TMEM round trip):
 * D[128,N] = A[128,K] * B[N,K]^T with fp16 operands / fp32 accumulate, A and B given as
 * fp32 row-major (K contiguous) and converted on device.  Used by tests/test_umma_gpu.py.
 * ---------------------------------------------------------------------------------- */
int c3d_selftest_umma(const float* a, const float* b, float* d, int32_t n, int32_t k,
                      int32_t a_in_tmem, void* stream);

/* Same for the CTA-pair form (tcgen05 cta_group::2, cluster of two CTAs): D[256,N] = A[256,K] * B[N,K]^T, N a
 * multiple of 32.  Isolates what the CTA-pair CIPS kernel (C3D_CIPS_PAIR=1) relies on. */
int c3d_selftest_umma_pair(const float* a, const float* b, float* d, int32_t n, int32_t k, void* stream);

/* ------------------------------------------------------------------------------------
 * Per-point linear layer of the NeRF TRAINING graph on tcgen05: replaces the `self.linear(x)` GEMM of
 * FiLMLayer.forward (exp/comm/models/film_layer.py:78-107) and of color_layer_linear (exp/cips3d/models/generator.py:236-243)
 * and, with transposed != 0, their data gradient -- fp32-equivalent products (fp16 hi/lo split, three MMA passes), where
 * torch runs fp32 SIMT sgemm.
 *   y (rows, n) = x (rows, k) . w^T (+ bias)       w: (n, k) row-major         [transposed == 0: the layer forward]
 *   y (rows, n) = x (rows, k) . w    (+ bias)       w: (k, n) row-major         [transposed != 0: dX = dZ . W with W as stored]
 * k, n in {32, 64, 128}; x, y, bias 16-byte aligned fp32; scale: device scalar or NULL -- operands are multiplied by it before
 * the fp16 split and the result divided by it (pass 1024 / max|x| for small-magnitude gradients).  workspace: at least
 * c3d_points_linear_workspace_bytes(n, k) bytes of device memory (the prepared weight blob).  Returns 0 or a negative C3D_E* code.
 * ---------------------------------------------------------------------------------- */
size_t c3d_points_linear_workspace_bytes(int32_t n, int32_t k);
int c3d_points_linear(const float* x, const float* w, const float* bias, const float* scale, float* y, int64_t rows, int32_t k,
                      int32_t n, int32_t transposed, void* workspace, size_t workspace_bytes, void* stream);

/* The discriminator's 4 x 4 Blur (upfirdn2d with up = down = 1; exp/cips3d/models/discriminator.py:67-82) on a CHANNELS-LAST
 * tensor: x (n, in_h, in_w, channels) in memory -> y (n, in_h + pad_y0 + pad_y1 - 3, in_w + pad_x0 + pad_x1 - 3, channels);
 * kernel: 4 x 4 taps (device), applied as c3d_upfirdn2d applies them (flipped: true convolution); pads in pixels, negative =
 * crop.  An extension beside the reference's native boundary (whose op sees (N*C, H, W, 1) planes): it lets the discriminator
 * stay in the layout the tensor-core convolutions use.  Returns 0 or a negative C3D_E* code. */
int c3d_blur_nhwc(const float* x, const float* kernel, float* y, int32_t n, int32_t in_h, int32_t in_w, int32_t channels,
                  int32_t pad_x0, int32_t pad_x1, int32_t pad_y0, int32_t pad_y1, void* stream);

/* ------------------------------------------------------------------------------------
 * Introspection for the CPU-side protocol test (host only, no GPU work): the order in which
 * c3d_cips_fwd's kernel issues the 32 (K64 x N128) weight tiles of a 512x512 layer and the 4
 * tiles of the padded input layer.  Entry = kc | nc << 4 | need << 8 | rdy << 12 (see
 * csrc/cips_tc.cu, KArgs::order_full).  Returns the depth of the weight ring (> 0) or a
 * negative C3D_E* code.  tests/test_cips_protocol_cpu.py replays the kernel's mbarrier
 * protocol on it (no parity aliasing, no in-place overwrite of a live operand, no deadlock).
 * ---------------------------------------------------------------------------------- */
int c3d_debug_cips_tile_order(uint16_t* order_full32, uint16_t* order_in4);

/* Which form of the fused renderer kernel the last c3d_ray_siren_fwd call on this process launched:
 * 0 block-wide ray math (default), 1 C3D_RAY_MATH=warp, 2 C3D_RAY_MATH=fold; -1 before the first call.
 * The opt-in forms fall back (S too large, per-point debug outputs requested); tests use this to make sure they
 * exercised the form they name. */
int c3d_debug_ray_math_mode(void);

/* Kernel-variant switches (A/B hooks of this build, not part of the drop-in contract): C3D_CIPS_CLUSTER, C3D_CIPS_PAIR,
 * C3D_BLUR, C3D_PIGAN_IMPL, C3D_PIGAN_PAIR, C3D_RAY_MATH are read from the environment ONCE, at the first call into the
 * library -- never on a launch path.  c3d_reload_options() re-reads them (host layer / tests: the Python binding calls it
 * only when one of the variables changed).  Not safe against concurrent launches on other threads. */
void c3d_reload_options(void);

/* Diagnostic (no GPU work): how many thread-block clusters of `cl` CTAs of the CIPS kernel (pair != 0: its cta_group::2
 * instantiation) can be resident at once on the current device -- a persistent kernel whose grid exceeds that runs in waves.
 * Negative on error. */
int c3d_debug_cips_max_clusters(int cl, int pair);

#ifdef __cplusplus
}
#endif
#endif /* CIPS3D_B200_H_ */
