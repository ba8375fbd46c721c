// Backward of the fused CIPS MLP (first part): the gradient CHAIN on tcgen05.
// Emulation-verified, not yet run on hardware; simple protocol first (one issuer, layer-level hand-offs), to be tuned on a GPU.
//
// Forward (cips_tc.cu, CIPSNet.forward exp/cips3d/models/generator.py:1107-1154): per image, layers l = 0..L-1
//     z_l = x_l W''_l ,   y_l = lrelu_0.2(z_l) (+ y_{l-2} on the second layer of a skip block) ,   x_{l+1} = y_l ,
//     rgb_pre = sum_blk y_{2 blk + 1} Wrgb_blk^T + b ,  rgb = tanh(rgb_pre),        W''[k][n] = s1p[k] W[k][n] d[n].
// Given g = dL/d rgb_pre (B,N,3) and the activation stash of the training forward (c3d_cips_fwd_train: y_l as fp16), one
// persistent CTA per SM walks 128-pixel tiles and runs the chain on-chip, last layer first:
//     G_l   = dL/dy_l = dX_{l+1}  (+ g Wrgb_blk on a block's second layer)  (+ G_{l+2} when block blk+1 is a skip block)
//     dZ_l  = G_l * lrelu'(z_l)          (sign of z_l from the stash: y_l - residual)
//     dX_l  = dZ_l W''_l^T               tcgen05: A = dZ_l (fp16, shared memory, 128 x 512), B = 16 KB tiles of W''_l^T streamed
//                                        through a 5-stage ring, fp32 accumulator = all 512 TMEM columns
// dZ_l is also written to HBM (fp16): the weight gradients dW''_l = X_l^T dZ_l are K = pixels GEMMs over the two stashes and
// are left to the library (plain GEMMs; the host side folds them into dW, d s1p, d demod).  Gradients are carried with a
// caller-chosen scale S (g is pre-multiplied) so that fp16 operands neither underflow nor overflow; dZ and dX come out scaled.
#include <atomic>

#include "c3d_common.cuh"

namespace c3d {
namespace cipsb {

constexpr int kH = 512, kTileM = 128, kKC = 64, kNC = 128;
constexpr int kWTileBytes = kKC * kNC * 2;      // 16 KB: (N = 128 rows of the layer's INPUT index) x (K = 64 of its output index)
constexpr int kStages = 5;
constexpr int kXBytes = kTileM * kH * 2;
constexpr int kLBO = kTileM * 16, kSBO = 128;
constexpr int kNumEpiWarps = 16, kThreads = 32 * (4 + kNumEpiWarps);
constexpr int kMaxLayers = C3D_CIPS_MAX_LAYERS;

struct Smem {
  alignas(1024) uint8_t x[kXBytes];               // A operand: dZ_l (fp16, UMMA K-major no-swizzle layout)
  alignas(1024) uint8_t w[kStages][kWTileBytes];
  float4 rgbw[kH];                                // ToRGB weights of the current block: (w0, w1, w2, 0) per hidden unit
  alignas(8) uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t a_ready, d_ready;
  uint32_t tmem_base;
};

struct KArgs {
  const __half* acts;          // (L, B, N, 512) stash of the training forward
  const uint16_t* zsign;       // (L, B, N, 32) sign bits of z_l, valid for the layers that add a residual
  size_t layer_stride;         // B * N * 512
  const float* g;              // (B, N, 3) dL/d rgb_pre, pre-multiplied by the scale
  __half* dz;                  // out (L, B, N, 512) scaled dZ_l
  float* dx;                   // out (B, N, in_dim) scaled dL/dx or null
  const __half* wtiles;        // per image: layers L-1 .. 0, each nblk(l) x 8 tiles in issue order
  size_t img_tile_stride;
  const float4* rgbw;          // (n_blocks, 512)
  float4* skipg;               // (gridDim.x, 128, 128) float4 scratch: gradient that bypasses a skip block
  int B, N, in_dim, n_layers, skip_from, rgb_from, tiles_per_img, total_tiles;
  int layer_tile_off[kMaxLayers + 1];     // indexed by the layer number l (tiles of layer l start here)
};

__global__ void __launch_bounds__(kThreads, 1) cips_bwd_tc_kernel(const KArgs a) {
  C3D_DYN_SMEM(uint8_t, smem_raw);
  Smem& s = *reinterpret_cast<Smem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&s.full[i], 1);
      mbar_init(&s.empty[i], 1);
    }
    mbar_init(&s.a_ready, kNumEpiWarps);
    mbar_init(&s.d_ready, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(&s.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s.tmem_base;
  const int iters = (a.total_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int L = a.n_layers;
  auto nblk = [&](int l) { return l == 0 ? 1 : kH / kNC; };       // 128-wide blocks of the layer's input dimension

  if (warp < 4) {
    reg_dec<56>();
    if (warp == 0) {
      // ---------------------------------------------------------- weight producer: layers L-1 .. 0 of the tile's image
      uint32_t stage = 0, phase = 0;
      for (int it = 0; it < iters; ++it) {
        const int tile = it * (int)gridDim.x + (int)blockIdx.x;
        const int img = tile < a.total_tiles ? tile / a.tiles_per_img : 0;
        for (int l = L - 1; l >= 0; --l) {
          const uint8_t* src = reinterpret_cast<const uint8_t*>(a.wtiles) +
                               ((size_t)img * a.img_tile_stride + (size_t)a.layer_tile_off[l]) * kWTileBytes;
          const int ntiles = nblk(l) * (kH / kKC);
          for (int t = 0; t < ntiles; ++t) {
            mbar_wait(&s.empty[stage], phase ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(&s.full[stage], kWTileBytes);
              bulk_g2s(s.w[stage], src + (size_t)t * kWTileBytes, kWTileBytes, &s.full[stage]);
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------- MMA issuer
      const uint32_t idesc = umma_idesc_f16(kTileM, kNC);
      const uint32_t dhi = umma_desc_hi(kSBO);
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(s.x), kLBO);
      const uint32_t b_lo0 = umma_desc_lo(smem_u32(s.w[0]), kLBO);
      constexpr uint32_t kStepK16 = (2 * kLBO) >> 4;
      constexpr uint32_t kStepStage = kWTileBytes >> 4;
      uint32_t stage = 0, phase = 0, apar = 0;
      for (int it = 0; it < iters; ++it)
        for (int l = L - 1; l >= 0; --l) {
          mbar_wait(&s.a_ready, apar);
          apar ^= 1;
          tc_fence_after();
          const int nb = nblk(l);
#pragma unroll 1
          for (int nc = 0; nc < nb; ++nc)
#pragma unroll 1
            for (int kc = 0; kc < kH / kKC; ++kc) {
              mbar_wait(&s.full[stage], phase);
              tc_fence_after();
              if (elect_one()) {
                const uint32_t a_lo = a_lo0 + kc * (kStepK16 * (kKC / 16));
                const uint32_t b_lo = b_lo0 + stage * kStepStage;
                const uint32_t d = tmem + nc * kNC;
#pragma unroll
                for (int k = 0; k < kKC / 16; ++k) umma_ss_w(d, a_lo + k * kStepK16, b_lo + k * kStepK16, dhi, idesc, (kc | k) != 0);
                tc_commit(&s.empty[stage]);
                if (nc == nb - 1 && kc == kH / kKC - 1) tc_commit(&s.d_ready);
              }
              __syncwarp();
              if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    }
  } else {
    // ------------------------------------------------------------ workers: thread = row x 128 columns (4 chunks of 32)
    reg_inc<104>();
    const int ew = warp - 4;
    const int wg = ew >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    float4* skipg = a.skipg + (size_t)blockIdx.x * (kH / 4) * kTileM;
    uint32_t dpar = 0;
    for (int it = 0; it < iters; ++it) {
      const int tile = it * (int)gridDim.x + (int)blockIdx.x;
      const bool tile_ok = tile < a.total_tiles;
      const int img = tile_ok ? tile / a.tiles_per_img : 0;
      const int pix = tile_ok ? (tile % a.tiles_per_img) * kTileM + row : a.N;
      const bool row_ok = tile_ok && pix < a.N;
      const size_t prow = (size_t)img * a.N + pix;
      float g0 = 0.f, g1 = 0.f, g2 = 0.f;
      if (row_ok) { g0 = a.g[prow * 3]; g1 = a.g[prow * 3 + 1]; g2 = a.g[prow * 3 + 2]; }
      for (int l = L - 1; l >= 0; --l) {
        const int blk = l >> 1;
        const bool second = (l & 1) != 0;
        const bool inject_rgb = second && blk >= a.rgb_from;
        const bool skip_here = second && blk >= a.skip_from && blk >= 1;                 // y_l = lrelu(z_l) + y_{l-2}
        const bool skip_above = second && (blk + 1) * 2 < L && (blk + 1) >= a.skip_from;  // y_l also feeds y_{l+2} directly
        if (inject_rgb) {       // stage this block's ToRGB weights (the previous users are past their barrier)
          s.rgbw[(int)threadIdx.x - 128] = __ldg(a.rgbw + (size_t)blk * kH + ((int)threadIdx.x - 128));
          named_bar_sync_c<1, kNumEpiWarps * 32>();
        }
        if (l < L - 1) {
          mbar_wait(&s.d_ready, dpar);
          dpar ^= 1;
          tc_fence_after();
        }
        const __half* yl = a.acts + (size_t)l * a.layer_stride + prow * kH;
        const uint16_t* zs = skip_here ? a.zsign + ((size_t)l * a.layer_stride + prow * kH) / 16 : nullptr;
        __half* dzl = a.dz + (size_t)l * a.layer_stride + prow * kH;
#pragma unroll 1
        for (int j = 0; j < 4; ++j)
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            const int col = j * 128 + wg * 32 + h * 16;
            float G[16];
            if (l < L - 1) {
              uint32_t acc[16];
              tmem_ld16(trow + (uint32_t)col, acc);
              tc_wait_ld();
#pragma unroll
              for (int i = 0; i < 16; ++i) G[i] = __uint_as_float(acc[i]);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) G[i] = 0.f;
            }
            if (inject_rgb) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float4 w4 = s.rgbw[col + i];
                G[i] = fmaf(g0, w4.x, fmaf(g1, w4.y, fmaf(g2, w4.z, G[i])));
              }
            }
            float4* sp = skipg + (size_t)(col / 4) * kTileM + row;
            if (skip_above) {
#pragma unroll
              for (int gq = 0; gq < 4; ++gq) {
                const float4 v = sp[gq * kTileM];
                G[4 * gq] += v.x; G[4 * gq + 1] += v.y; G[4 * gq + 2] += v.z; G[4 * gq + 3] += v.w;
              }
            }
            if (skip_here) {
#pragma unroll
              for (int gq = 0; gq < 4; ++gq) sp[gq * kTileM] = make_float4(G[4 * gq], G[4 * gq + 1], G[4 * gq + 2], G[4 * gq + 3]);
            }
            // lrelu'(z_l): sign of z_l = sign of (y_l - residual) from the stash
            uint32_t pk[8];
            if (row_ok) {
              const uint4 ya = reinterpret_cast<const uint4*>(yl + col)[0], yb = reinterpret_cast<const uint4*>(yl + col)[1];
              const uint32_t yw[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
              const uint32_t sb = zs ? (uint32_t)zs[col / 16] : 0u;      // residual layers: stored sign bits of z_l
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float2 yy = __half22float2(*reinterpret_cast<const __half2*>(&yw[i]));
                const bool p0 = zs ? (sb >> (2 * i) & 1u) != 0 : yy.x > 0.f;
                const bool p1 = zs ? (sb >> (2 * i + 1) & 1u) != 0 : yy.y > 0.f;
                pk[i] = pack_f16(G[2 * i] * (p0 ? 1.f : 0.2f), G[2 * i + 1] * (p1 ? 1.f : 0.2f));
              }
              reinterpret_cast<uint4*>(dzl + col)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              reinterpret_cast<uint4*>(dzl + col)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) pk[i] = 0u;
            }
            uint8_t* xp = s.x + (size_t)(col / 8) * kLBO + row * 16;
            *reinterpret_cast<uint4*>(xp) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(xp + kLBO) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.a_ready);
      }
      // ---- dL/dx = dX_0 (first in_dim columns of accumulator block 0)
      mbar_wait(&s.d_ready, dpar);
      dpar ^= 1;
      tc_fence_after();
      if (wg == 0) {
#pragma unroll 1
        for (int c0 = 0; c0 < a.in_dim; c0 += 16) {
          uint32_t acc[16];
          tmem_ld16(trow + (uint32_t)c0, acc);
          tc_wait_ld();
          if (a.dx && row_ok) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (c0 + i < a.in_dim) a.dx[prow * a.in_dim + c0 + i] = __uint_as_float(acc[i]);
          }
        }
      }
      tc_fence_before();
      named_bar_sync_c<1, kNumEpiWarps * 32>();      // accumulator and rgbw are free for the next tile
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem);
}

// W''_l^T tiles: block = one 16 KB tile of one image; tile t of layer l is (nc, kc) = (t / 8, t % 8) and holds
// B[i][n] = W''[nc*128 + i][kc*64 + n] = s1p[.] W[.][.] d[.] at byte (i%8)*16 + (i/8)*128 + (n/8)*2048 + (n%8)*2
struct PrepArgs {
  const float* w[kMaxLayers];
  const float* s1p[kMaxLayers];
  const float* demod[kMaxLayers];
  int layer_tile_off[kMaxLayers + 1];
  int in_dim0, n_layers;
};
__global__ void cips_bwd_prep_weights_kernel(const PrepArgs pa, __half* __restrict__ out) {
  const int gt = blockIdx.x, b = blockIdx.y;
  int l = 0;
  while (l + 1 < pa.n_layers && gt >= pa.layer_tile_off[l + 1]) ++l;
  const int t = gt - pa.layer_tile_off[l];
  const int nc = t / (kH / kKC), kc = t % (kH / kKC);
  const int in_dim = l == 0 ? pa.in_dim0 : kH;
  const float* W = pa.w[l];
  const float* sv = pa.s1p[l] + (size_t)b * in_dim;
  const float* dv = pa.demod[l] + (size_t)b * kH + kc * kKC;
  __half* o = out + ((size_t)b * pa.layer_tile_off[pa.n_layers] + gt) * (kWTileBytes / 2);
  for (int e = threadIdx.x; e < kKC * kNC; e += blockDim.x) {
    const int i = e / kKC, n = e % kKC;       // n fastest: coalesced reads of a W row
    const int gi = nc * kNC + i;
    const float v = gi < in_dim ? (sv[gi] * W[(size_t)gi * kH + kc * kKC + n]) * dv[n] : 0.f;
    o[((i % 8) * 16 + (i / 8) * 128 + (n / 8) * kLBO) / 2 + (n % 8)] = __float2half_rn(v);
  }
}
__global__ void cips_bwd_prep_rgbw_kernel(C3dCipsWeights w, int n_blocks, int rgb_from, float4* rgbw) {
  const int blk = blockIdx.x;
  if (blk < n_blocks && blk >= rgb_from)
    for (int n = threadIdx.x; n < kH; n += blockDim.x)
      rgbw[(size_t)blk * kH + n] = make_float4(w.rgb_w[blk][n], w.rgb_w[blk][kH + n], w.rgb_w[blk][2 * kH + n], 0.f);
}

}  // namespace cipsb
}  // namespace c3d

using namespace c3d;
using namespace c3d::cipsb;

struct BwdWs {
  size_t wtiles, rgbw, skipg, total;
  int tiles_per_img;
};
static BwdWs bwd_ws_layout(const C3dCipsParams* p) {
  BwdWs o;
  const int L = 2 * p->n_blocks;
  o.tiles_per_img = (1 + (L - 1) * (kH / kNC)) * (kH / kKC);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t r = off; off += (bytes + 255) / 256 * 256; return r; };
  o.wtiles = take((size_t)(p->batch > 0 ? p->batch : 1) * o.tiles_per_img * kWTileBytes);
  o.rgbw = take((size_t)p->n_blocks * kH * 16);
  o.skipg = take((size_t)160 * (kH / 4) * kTileM * 16);
  o.total = off;
  return o;
}

extern "C" size_t c3d_cips_bwd_workspace_bytes(const C3dCipsParams* p) { return p ? bwd_ws_layout(p).total : 0; }

extern "C" int c3d_cips_bwd(const C3dCipsParams* p, const C3dCipsWeights* w, const void* acts_f16, const void* zsign_u16,
                            const float* g_rgb_pre, void* dz_f16, float* dx, void* workspace, size_t workspace_bytes, void* stream) {
  C3D_CHECK_ARG(p && w && acts_f16 && zsign_u16 && g_rgb_pre && dz_f16, "cips_bwd: null pointer");
  C3D_CHECK_ARG(p->hidden == kH, "cips_bwd: hidden must be 512, got %d", p->hidden);
  C3D_CHECK_ARG(p->in_dim >= 1 && p->in_dim <= 64, "cips_bwd: in_dim must be <= 64, got %d", p->in_dim);
  C3D_CHECK_ARG(p->n_blocks >= 1 && 2 * p->n_blocks <= kMaxLayers && p->skip_from >= 1, "cips_bwd: bad block counts");
  const int L = 2 * p->n_blocks;
  for (int l = 0; l < L; ++l) C3D_CHECK_ARG(w->w[l] && w->style1p[l] && w->demod[l], "cips_bwd: null weight/style/demod for layer %d", l);
  for (int b = p->rgb_from; b < p->n_blocks; ++b) C3D_CHECK_ARG(b < 0 || w->rgb_w[b], "cips_bwd: null ToRGB weights for block %d", b);
  if (p->batch == 0 || p->n_pix == 0) return C3D_OK;
  const BwdWs ws = bwd_ws_layout(p);
  C3D_CHECK_ARG(workspace, "cips_bwd: null workspace");
  if (workspace_bytes < ws.total) {
    c3d_set_error("cips_bwd: workspace too small (%zu < %zu)", workspace_bytes, ws.total);
    return C3D_EWORKSPACE;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  if (!c3d_device_supported(dev)) {
    c3d_set_error("cips_bwd: device %d is not sm_100 (tcgen05 required)", dev);
    return C3D_EARCH;
  }
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* base = (uint8_t*)workspace;
  KArgs ka = {};
  ka.acts = (const __half*)acts_f16;
  ka.zsign = (const uint16_t*)zsign_u16;
  ka.layer_stride = (size_t)p->batch * p->n_pix * kH;
  ka.g = g_rgb_pre;
  ka.dz = (__half*)dz_f16;
  ka.dx = dx;
  ka.wtiles = (const __half*)(base + ws.wtiles);
  ka.img_tile_stride = (size_t)ws.tiles_per_img;
  ka.rgbw = (const float4*)(base + ws.rgbw);
  ka.skipg = (float4*)(base + ws.skipg);
  ka.B = p->batch; ka.N = p->n_pix; ka.in_dim = p->in_dim; ka.n_layers = L;
  ka.skip_from = p->skip_from; ka.rgb_from = p->rgb_from;
  ka.tiles_per_img = (p->n_pix + kTileM - 1) / kTileM;
  ka.total_tiles = p->batch * ka.tiles_per_img;
  PrepArgs pa = {};
  int off = 0;
  for (int l = 0; l < L; ++l) {
    ka.layer_tile_off[l] = pa.layer_tile_off[l] = off;
    off += (l == 0 ? 1 : kH / kNC) * (kH / kKC);
    pa.w[l] = w->w[l]; pa.s1p[l] = w->style1p[l]; pa.demod[l] = w->demod[l];
  }
  ka.layer_tile_off[L] = pa.layer_tile_off[L] = off;
  pa.in_dim0 = p->in_dim; pa.n_layers = L;
  C3D_LAUNCH(cips_bwd_prep_weights_kernel, dim3(off, p->batch), 256, 0, st, pa, (__half*)(base + ws.wtiles));
  C3D_LAUNCH_CHECK();
  C3D_LAUNCH(cips_bwd_prep_rgbw_kernel, p->n_blocks, 256, 0, st, *w, p->n_blocks, p->rgb_from, (float4*)(base + ws.rgbw));
  C3D_LAUNCH_CHECK();
  const size_t smem = sizeof(Smem) + 1024;
  static std::atomic<unsigned long long> attr_set{0};
  if (!(attr_set.load() >> (dev & 63) & 1ull)) {
    C3D_CUDA(cudaFuncSetAttribute(cips_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set.fetch_or(1ull << (dev & 63));
  }
  const int sms = c3d_device_sm_count(dev);
  int grid = ka.total_tiles < sms ? ka.total_tiles : sms;
  if (grid > 160) grid = 160;
  C3D_LAUNCH(cips_bwd_tc_kernel, grid, kThreads, smem, st, ka);
  C3D_LAUNCH_CHECK();
  return C3D_OK;
}
