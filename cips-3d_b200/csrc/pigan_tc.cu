// pi-GAN renderer on tcgen05 (C3D_PIGAN_IMPL=tc; emulation-verified, not yet run on hardware): rays -> 8-layer 256-wide
// FiLM-SIREN -> sigma head + view-dependent colour -> importance resampling -> second pass -> merge -> rgb compositing, one
// kernel, nothing per-sample written to HBM (piGAN_lib/generators/generators.py:26-96, siren/siren.py:133-152,196-215).
//
// Work unit: a ray group of G = floor(128 / S) whole rays = one 128-row MMA block.  A persistent CTA carries ONE group at a
// time (the 256-wide layers need all of tensor memory: A operand 2 x 128 columns (fp16 hi / lo, two K values per column,
// written by the epilogue with tcgen05.st) + fp32 accumulator 256 columns); with K = N = 256 a layer is 6.1 k tensor-pipe
// clocks against ~1.5 k of epilogue, so one slot already keeps the tensor pipe busy most of the time.
//   * weights (8 layers x 256x256, fp16 hi + lo = 2 MB, scaled by 2^8 so the lo parts stay normal) do not fit shared
//     memory: they stream as 32 KB tiles (N 256 x K 64) through a 4-stage ring filled by bulk async copies, hi and lo tile
//     of each K chunk in turn; every layer is the split-precision sum A_hi*B_hi + A_lo*B_hi + A_hi*B_lo (as in
//     ray_siren_tc.cu: FiLM gains ~30 amplify the pre-activation error of a single fp16 pass beyond the 1e-3 bar);
//   * layer 0 (3 -> 256) is three FMAs per unit in the epilogue threads against per-image folded constants (gain, box
//     warp, bias, phase) -- no MMA phase, and every B operand is image independent (needed for the CTA-pair form);
//   * PAIR (C3D_PIGAN_PAIR=1): the two CTAs of a cluster form a tcgen05 cta_group::2 pair -- M = 256 = two ray groups,
//     each CTA streams half of every weight tile (N rows 128 r .. +127), the leader issues, the protocol is the CIPS pair
//     kernel's (relayed fills, workers of both CTAs report to the leader's a_ready, multicast commits);
//   * the sigma head and the colour linear (256 -> 3) are dot products in the epilogues of layer 7 / the colour layer
//     (partial sums per 64-column thread, reduced through shared memory), the ray direction's 3 input columns of the colour
//     layer are added there too (no K = 259 MMA);
//   * per-ray math (resampling, stable rank sort, transmittance scans, compositing) is one warp per ray with shuffles.
// warp 0: weight producer; warp 1: MMA issuer; warp 2: TMEM allocation; warps 4..19: 16 worker warps (row = TMEM lane,
// 64 columns each).
#include <atomic>

#include "c3d_common.cuh"
#include "ray_math.cuh"

namespace c3d {
namespace pgt {

constexpr int kRows = 128, kH = 256, kLayers = 8;          // hidden width, FiLM layers before the heads
constexpr int kKC = 64;                                    // K per streamed weight tile
constexpr int kTileBytes = kH * kKC * 2;                   // 32 KB: B tile (N = 256) x (K = 64) fp16
constexpr int kStages = 4;
constexpr int kRingBytes = kStages * kTileBytes;
template <bool PAIR> struct Ring {                         // PAIR: each CTA holds N/2 = 128 rows of every tile
  static constexpr int kStageBytes = PAIR ? kTileBytes / 2 : kTileBytes;
  static constexpr int kNS = PAIR ? 2 * kStages : kStages;
  static constexpr int kLBO_B = PAIR ? (kH / 2) * 16 : kH * 16;
};
constexpr int kTilesPerLayer = 2 * (kH / kKC);             // (hi, lo) x 4 K chunks
constexpr int kStreamLayers = kLayers;                     // layers 1..7 and the colour layer
constexpr int kLBO = kH * 16;                              // K-direction core-matrix stride of a 256-row operand
constexpr float kWScale = 256.f, kWInv = 1.f / 256.f;
constexpr int kThreads = 640, kWorkers = 512;

struct ImgConsts {                    // per image, written by the prep kernel
  float4 w0f[kH];                     // layer 0 folded: (f0*s*W0[j][0..2], f0*b0[j] + ph0[j])
  float2 film[kLayers][kH];           // layers 1..7 and colour (index 7): (f / 256, f * b + ph)
  float4 wdir[kH];                    // f_colour * Wc[:, 0:3]  (the ray direction's input columns)
};

template <bool PAIR>
struct SmemT {
  static constexpr int NS = Ring<PAIR>::kNS;
  alignas(1024) uint8_t ring[kRingBytes];
  alignas(128) float4 w0f[kH];
  float2 film[kLayers][kH];
  float4 wdir[kH];
  float4 wl4[kH];                     // (Wl[0][k], Wl[1][k], Wl[2][k], Wsigma[k])
  float4 samp[2][kRows];              // [0] fine pass, [1] coarse pass: (r, g, b, sigma) per sample
  float z_c[kRows], z_f[kRows];
  float part_sig[4][kRows];
  float4 part_rgb[4][kRows];
  float skey[2 * kRows];
  int sidx[2 * kRows];
  alignas(8) uint64_t full[NS];
  uint64_t empty[NS];
  uint64_t peer_full[NS];             // PAIR, leader only: the peer's half of stage s has landed
  uint64_t a_ready, d_ready;
  uint32_t tmem_base;
};

struct KArgs {
  C3dRayParams p;
  C3dRayIO io;
  const uint8_t* wtiles;        // kStreamLayers * kTilesPerLayer tiles of kTileBytes, stream order
  const ImgConsts* consts;      // (B)
  const float4* wl4;            // (256)
  const float* b_sigma;         // (1) final_layer bias
  const float* bl;              // (3) color_layer_linear bias
  int lock_view;
  int G, groups_per_img, total_groups;
};

__device__ __forceinline__ void store_a16(uint32_t a_hi_col, uint32_t a_lo_col, const float (&v)[16]) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) split_f16(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
  tmem_st8(a_hi_col, hi);
  tmem_st8(a_lo_col, lo);
}
__device__ __forceinline__ float sample_alpha(float delta, float sigma, float noise, int clamp_mode) {
  const float sn = __fadd_rn(sigma, noise);
  const float act = clamp_mode == 1 ? softplus_f32(sn) : fmaxf(sn, 0.f);
  return __fsub_rn(1.f, expf(__fmul_rn(-delta, act)));
}

template <bool PAIR>
__global__ void __launch_bounds__(kThreads, 1) pigan_tc_kernel(const KArgs a) {
  using Smem = SmemT<PAIR>;
  using RC = Ring<PAIR>;
  constexpr int NS = RC::kNS;
  C3D_DYN_SMEM(uint8_t, smem_raw);
  Smem& s = *reinterpret_cast<Smem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = PAIR ? cluster_ctarank() : 0;
  const bool leader = !PAIR || crank == 0;
  const C3dRayParams& p = a.p;
  const int S = p.num_steps, G = a.G;
  const bool hier = p.hierarchical != 0;
  const int nS = hier ? 2 * S : S;
  const int passes = hier ? 2 : 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(&s.full[i], 1);
      mbar_init(&s.empty[i], 1);
      mbar_init(&s.peer_full[i], 1);
    }
    mbar_init(&s.a_ready, PAIR ? 2 * (kWorkers / 32) : kWorkers / 32);     // PAIR: the leader's barrier collects both CTAs
    mbar_init(&s.d_ready, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    if (PAIR) tmem_alloc_cg2<512>(&s.tmem_base);
    else tmem_alloc<512>(&s.tmem_base);
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = s.tmem_base;
  // group walked in iteration `it`: PAIR -> the cluster takes two consecutive groups
  const int grid_units = PAIR ? (int)gridDim.x / 2 : (int)gridDim.x;
  const int total_units = PAIR ? (a.total_groups + 1) / 2 : a.total_groups;
  const int iters = (total_units + grid_units - 1) / grid_units;
  auto group_of = [&](int it) {
    return PAIR ? (it * grid_units + (int)blockIdx.x / 2) * 2 + (int)crank : it * (int)gridDim.x + (int)blockIdx.x;
  };

  if (warp < 4) {
    reg_dec<56>();
    if (warp == 0) {
      // ---------------------------------------------------------- weight producer: the same 64 tiles for every group and pass
      uint32_t stage = 0, phase = 0;
      const int n_tiles = kStreamLayers * kTilesPerLayer;
      for (int it = 0; it < iters; ++it)
        for (int pass = 0; pass < passes; ++pass)
          for (int t = 0; t < n_tiles; ++t) {
            mbar_wait(&s.empty[stage], phase ^ 1);
            if (elect_one()) {       // PAIR: this CTA's half (N rows 128*rank .. +127) of the tile
              mbar_arrive_expect_tx(&s.full[stage], RC::kStageBytes);
              bulk_g2s(s.ring + stage * RC::kStageBytes, a.wtiles + (size_t)t * kTileBytes + crank * RC::kStageBytes, RC::kStageBytes,
                       &s.full[stage]);
            }
            __syncwarp();
            if (++stage == NS) { stage = 0; phase ^= 1; }
          }
    } else if (PAIR && !leader) {
      // ---------------------------------------------------------- peer CTA of a pair: warp 1 relays its fills to the leader
      if (warp == 1) {
        uint32_t stage = 0, phase = 0;
        const int n_tiles = kStreamLayers * kTilesPerLayer;
        for (int it = 0; it < iters; ++it)
          for (int pass = 0; pass < passes; ++pass)
#pragma unroll 1
            for (int t = 0; t < n_tiles; ++t) {
              mbar_wait(&s.full[stage], phase);
              if (elect_one()) mbar_arrive_cluster(&s.peer_full[stage], 0);
              __syncwarp();
              if (++stage == NS) { stage = 0; phase ^= 1; }
            }
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------- MMA issuer (converged warp, one elected lane issues)
      const uint32_t idesc = umma_idesc_f16(PAIR ? 2 * kRows : kRows, kH);
      const uint32_t dhi = umma_desc_hi(128);
      const uint32_t a_hi = tmem, a_lo = tmem + 128, d = tmem + 256;
      const uint32_t ring_lo0 = umma_desc_lo(smem_u32(s.ring), RC::kLBO_B);
      constexpr uint32_t kStepK16 = (2 * RC::kLBO_B) >> 4;      // one K = 16 step = 2 core-matrix columns of the B tile
      constexpr uint32_t kStepStage = RC::kStageBytes >> 4;
      auto mma = [&](uint32_t a_col, uint32_t b_lo, uint32_t acc) {
        if (PAIR) umma_ts_w_cg2(d, a_col, b_lo, dhi, idesc, acc);
        else umma_ts_w(d, a_col, b_lo, dhi, idesc, acc);
      };
      auto wait_tile = [&](uint32_t stage, uint32_t phase) {
        mbar_wait(&s.full[stage], phase);
        if (PAIR) mbar_wait_cluster(&s.peer_full[stage], phase);
        tc_fence_after();
      };
      auto release = [&](uint64_t* bar) {
        if (PAIR) tc_commit_cg2_mc(bar, 3);
        else tc_commit(bar);
      };
      uint32_t stage = 0, phase = 0, apar = 0;
      for (int it = 0; it < iters; ++it)
        for (int pass = 0; pass < passes; ++pass)
          for (int l = 0; l < kStreamLayers; ++l) {
            if (PAIR) mbar_wait_cluster(&s.a_ready, apar);
            else mbar_wait(&s.a_ready, apar);
            apar ^= 1;
            tc_fence_after();
#pragma unroll 1
            for (int kc = 0; kc < kH / kKC; ++kc) {
              // hi tile: A_hi * B_hi + A_lo * B_hi
              wait_tile(stage, phase);
              if (elect_one()) {
                const uint32_t b = ring_lo0 + stage * kStepStage;
#pragma unroll
                for (int k = 0; k < kKC / 16; ++k) mma(a_hi + (uint32_t)(kc * 4 + k) * 8, b + k * kStepK16, (kc | k) != 0);
#pragma unroll
                for (int k = 0; k < kKC / 16; ++k) mma(a_lo + (uint32_t)(kc * 4 + k) * 8, b + k * kStepK16, 1);
                release(&s.empty[stage]);
              }
              __syncwarp();
              if (++stage == NS) { stage = 0; phase ^= 1; }
              // lo tile: A_hi * B_lo
              wait_tile(stage, phase);
              if (elect_one()) {
                const uint32_t b = ring_lo0 + stage * kStepStage;
#pragma unroll
                for (int k = 0; k < kKC / 16; ++k) mma(a_hi + (uint32_t)(kc * 4 + k) * 8, b + k * kStepK16, 1);
                release(&s.empty[stage]);
                if (kc == kH / kKC - 1) release(&s.d_ready);
              }
              __syncwarp();
              if (++stage == NS) { stage = 0; phase ^= 1; }
            }
          }
    }
  } else {
    // ------------------------------------------------------------ workers: epilogues + per-ray math
    reg_inc<104>();
    const int ew = warp - 4;              // 0..15
    const int cq = ew >> 2;               // column quarter: columns [64 cq, 64 cq + 64)
    const int q = warp & 3;               // TMEM lane quarter
    const int row = q * 32 + lane;
    const int wtid = (int)threadIdx.x - 128;      // 0..511
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const uint32_t a_hi = tmem + lane_sel, a_lo = a_hi + 128, dcol = a_hi + 256 + (uint32_t)(cq * 64);
    auto wsync = [&]() { named_bar_sync_c<1, kWorkers>(); };
    uint32_t dpar = 0;
    auto signal_a = [&]() {
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(&s.a_ready, 0);
        else mbar_arrive(&s.a_ready);
      }
    };
    auto wait_d = [&]() {
      mbar_wait(&s.d_ready, dpar);
      dpar ^= 1;
      tc_fence_after();
    };
    const int g_row = row / S, s_row = row - g_row * S;
    const bool row_in_group = g_row < G;
    for (int i = wtid; i < kH; i += kWorkers) s.wl4[i] = __ldg(a.wl4 + i);
    int cur_img = -1;

    for (int it = 0; it < iters; ++it) {
      const int grp = group_of(it);
      const bool grp_ok = grp < a.total_groups;
      const int img = grp_ok ? grp / a.groups_per_img : 0;
      const int ray0 = grp_ok ? (grp % a.groups_per_img) * G : 0;
      const int n_valid = grp_ok ? min(G, p.n_rays - ray0) : 0;
      const bool pt_ok = row_in_group && g_row < n_valid;
      const int nloc = ray0 + g_row;
      const size_t ro_row = (size_t)img * p.n_rays + nloc;
      const float* M = a.io.cam2world + (size_t)img * 16;
      if (img != cur_img) {      // (CTA-uniform) stage the image's folded constants
        const uint4* src = reinterpret_cast<const uint4*>(&a.consts[img]);
        uint4* dst = reinterpret_cast<uint4*>(s.w0f);    // w0f, film, wdir are contiguous in both structs
        for (int i = wtid; i < (int)(sizeof(ImgConsts) / 16); i += kWorkers) dst[i] = __ldg(src + i);
        cur_img = img;
        wsync();
      }
      RayFrame fr;
      int gray = 0;
      float dx = 0.f, dy = 0.f, dz = -1.f;
      if (pt_ok) {
        gray = a.io.ray_idx ? a.io.ray_idx[nloc] : p.ray_offset + nloc;
        fr = make_ray_frame(M, gray, p.img_size, p.z_cam);
        if (!a.lock_view) { dx = fr.dwx; dy = fr.dwy; dz = fr.dwz; }
      }

      for (int pass = 0; pass < passes; ++pass) {
        // ---------------- layer 0 in the epilogue threads: h0 = sin(f0 (W0 (p s) + b0) + ph0), folded per image -> A
        float px = 0.f, py = 0.f, pz = 0.f;
        if (pt_ok) {
          if (pass == 0) {
            const float u = a.io.jitter_u[((size_t)img * p.img_size * p.img_size + gray) * S + s_row];
            float z;
            coarse_sample(fr, M, p.ray_start, p.ray_end, S, s_row, u, z, px, py, pz);
            if (cq == 0) s.z_c[row] = z;
          } else {
            fine_sample(fr, s.z_f[row], px, py, pz);
          }
        }
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float4 w4 = s.w0f[cq * 64 + c * 16 + j];
            v[j] = __sinf(fmaf(w4.x, px, fmaf(w4.y, py, fmaf(w4.z, pz, w4.w))));
          }
          store_a16(a_hi + (uint32_t)(cq * 32 + c * 8), a_lo + (uint32_t)(cq * 32 + c * 8), v);
        }
        signal_a();
        // ---------------- E1..E7: FiLM + sin -> A; E7 also accumulates the sigma head
        float psig = 0.f;
#pragma unroll 1
        for (int l = 0; l < kLayers - 1; ++l) {
          wait_d();
          const float2* fl = s.film[l] + cq * 64;
          const bool last_hidden = l == kLayers - 2;
          uint32_t acc[16];
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            tmem_ld16(dcol + (uint32_t)(c * 16), acc);
            tc_wait_ld();
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float2 g2 = fl[c * 16 + j];
              v[j] = __sinf(fmaf(__uint_as_float(acc[j]), g2.x, g2.y));
            }
            if (last_hidden) {
#pragma unroll
              for (int j = 0; j < 16; ++j) psig = fmaf(v[j], s.wl4[cq * 64 + c * 16 + j].w, psig);
            }
            store_a16(a_hi + (uint32_t)(cq * 32 + c * 8), a_lo + (uint32_t)(cq * 32 + c * 8), v);
          }
          signal_a();
        }
        // ---------------- colour layer: sin(f (Wc [d, h] + bc) + ph), then the colour linear as partial dot products
        wait_d();
        {
          const float2* fl = s.film[kLayers - 1] + cq * 64;
          float r0 = 0.f, r1 = 0.f, r2 = 0.f;
          uint32_t acc[16];
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            tmem_ld16(dcol + (uint32_t)(c * 16), acc);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int col = cq * 64 + c * 16 + j;
              const float2 g2 = fl[c * 16 + j];
              const float4 wd = s.wdir[col];
              const float dirt = fmaf(dx, wd.x, fmaf(dy, wd.y, dz * wd.z));
              const float v = __sinf(fmaf(__uint_as_float(acc[j]), g2.x, g2.y) + dirt);
              const float4 w4 = s.wl4[col];
              r0 = fmaf(v, w4.x, r0);
              r1 = fmaf(v, w4.y, r1);
              r2 = fmaf(v, w4.z, r2);
            }
          }
          s.part_rgb[cq][row] = make_float4(r0, r1, r2, 0.f);
          s.part_sig[cq][row] = psig;
        }
        tc_fence_before();
        wsync();
        if (cq == 0) {      // heads: rgb = sigmoid(Wl c + bl), sigma = Ws h + bs
          const float4 p0 = s.part_rgb[0][row], p1 = s.part_rgb[1][row], p2 = s.part_rgb[2][row], p3 = s.part_rgb[3][row];
          float4 o;
          o.x = 1.f / (1.f + expf(-(((p0.x + p1.x) + (p2.x + p3.x)) + __ldg(a.bl))));
          o.y = 1.f / (1.f + expf(-(((p0.y + p1.y) + (p2.y + p3.y)) + __ldg(a.bl + 1))));
          o.z = 1.f / (1.f + expf(-(((p0.z + p1.z) + (p2.z + p3.z)) + __ldg(a.bl + 2))));
          o.w = ((s.part_sig[0][row] + s.part_sig[1][row]) + (s.part_sig[2][row] + s.part_sig[3][row])) + __ldg(a.b_sigma);
          s.samp[pass == 0 ? 1 : 0][row] = o;
          float* dbg = pass == 0 ? a.io.dbg_coarse : a.io.dbg_fine;
          if (dbg && pt_ok) reinterpret_cast<float4*>(dbg)[ro_row * S + s_row] = o;
        }
        wsync();
        // ---------------- importance resampling: warp per ray, two rays per warp (S <= 16), lane e = coarse sample e
        if (pass == 0 && hier) {
          const unsigned fullm = 0xffffffffu;
          const int sub = lane >> 4, e = lane & 15, ns = S - 2;
          for (int g0 = 0; g0 < n_valid; g0 += 32) {
            const int g = g0 + ew + 16 * sub;
            const bool act = g < n_valid && e < S;
            const int r0 = g * S;
            const size_t ro = (size_t)img * p.n_rays + ray0 + g;
            const float z = act ? s.z_c[r0 + e] : 0.f;
            const float zn = __shfl_down_sync(fullm, z, 1, 16);
            float alpha = 0.f, f = 1.f;
            if (act) {
              const float delta = e + 1 < S ? __fsub_rn(zn, z) : 1e10f;
              const float nz = a.io.noise_c ? __fmul_rn(a.io.noise_c[ro * S + e], p.noise_std) : 0.f;
              alpha = sample_alpha(delta, s.samp[1][r0 + e].w, nz, p.clamp_mode);
              f = __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f);
            }
            float P = f;
#pragma unroll
            for (int dd = 1; dd < 16; dd <<= 1) {
              const float t = __shfl_up_sync(fullm, P, dd, 16);
              if (e >= dd) P = __fmul_rn(P, t);
            }
            float T = __shfl_up_sync(fullm, P, 1, 16);
            if (e == 0) T = 1.f;
            const float w = __fmul_rn(alpha, T);
            const float wn = __shfl_down_sync(fullm, w, 1, 16);
            const float wt = e <= S - 3 ? __fadd_rn(__fadd_rn(wn, 1e-5f), 1e-5f) : 0.f;
            float sum = wt;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(fullm, sum, o, 16);
            const float pdf = __fmul_rn(wt, __fdividef(1.f, sum));
            float C = pdf;
#pragma unroll
            for (int dd = 1; dd < 16; dd <<= 1) {
              const float t = __shfl_up_sync(fullm, C, dd, 16);
              if (e >= dd) C = __fadd_rn(C, t);
            }
            float cdf = __shfl_up_sync(fullm, C, 1, 16);
            if (e == 0) cdf = 0.f;
            const float uk = act ? a.io.pdf_u[ro * S + e] : 0.f;
            int i = 0;
            for (int j = 0; j <= ns; ++j) i += __shfl_sync(fullm, cdf, j, 16) < uk ? 1 : 0;
            const int below = max(i - 1, 0), above = min(i, ns);
            const float cb = __shfl_sync(fullm, cdf, below, 16), ca = __shfl_sync(fullm, cdf, above, 16);
            const float bb = 0.5f * __fadd_rn(__shfl_sync(fullm, z, below, 16), __shfl_sync(fullm, z, below + 1, 16));
            const float ba = 0.5f * __fadd_rn(__shfl_sync(fullm, z, above, 16), __shfl_sync(fullm, z, above + 1, 16));
            float denom = __fsub_rn(ca, cb);
            if (denom < 1e-5f) denom = 1.f;
            if (act) s.z_f[r0 + e] = __fadd_rn(bb, __fmul_rn(__fdiv_rn(__fsub_rn(uk, cb), denom), __fsub_rn(ba, bb)));
          }
          wsync();
        }
      }
      // ---------------- merge + compositing: warp per ray, lane e = element e of cat([fine, coarse])
      {
        const unsigned fullm = 0xffffffffu;
        for (int g = ew; g < n_valid; g += kWorkers / 32) {
          const int rc0 = g * S, base = g * nS, e = lane;
          const bool act = e < nS;
          const size_t ro = (size_t)img * p.n_rays + ray0 + g;
          const float k = act ? (hier ? (e < S ? s.z_f[rc0 + e] : s.z_c[rc0 + e - S]) : s.z_c[rc0 + e]) : 3.0e38f;
          int rank = 0;
          for (int j = 0; j < nS; ++j) {
            const float kj = __shfl_sync(fullm, k, j);
            rank += (kj < k || (kj == k && j < e)) ? 1 : 0;
          }
          if (act) {
            s.skey[base + rank] = k;
            s.sidx[base + rank] = hier ? e : S + e;
          }
          __syncwarp();
          const float ks = act ? s.skey[base + e] : 0.f;
          const int src = act ? s.sidx[base + e] : S;
          const float4 sv = act ? (src < S ? s.samp[0][rc0 + src] : s.samp[1][rc0 + src - S]) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float kn = __shfl_down_sync(fullm, ks, 1);
          float alpha = 0.f, f = 1.f;
          if (act) {
            const float delta = e + 1 < nS ? __fsub_rn(kn, ks) : 1e10f;
            const float nz = a.io.noise_f ? __fmul_rn(a.io.noise_f[ro * nS + e], p.noise_std) : 0.f;
            alpha = sample_alpha(delta, sv.w, nz, p.clamp_mode);
            f = __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f);
          }
          float P = f;
#pragma unroll
          for (int dd = 1; dd < 32; dd <<= 1) {
            const float t = __shfl_up_sync(fullm, P, dd);
            if (e >= dd) P = __fmul_rn(P, t);
          }
          float T = __shfl_up_sync(fullm, P, 1);
          if (e == 0) T = 1.f;
          float w = __fmul_rn(alpha, T);
          float wsum = w;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(fullm, wsum, o);
          if (p.last_back && e == nS - 1) w += 1.f - wsum;
          float c0 = w * sv.x, c1 = w * sv.y, c2 = w * sv.z, dd = act ? __fmul_rn(w, ks) : 0.f;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            c0 += __shfl_xor_sync(fullm, c0, o);
            c1 += __shfl_xor_sync(fullm, c1, o);
            c2 += __shfl_xor_sync(fullm, c2, o);
            dd += __shfl_xor_sync(fullm, dd, o);
          }
          if (lane < 3) {
            const float c = lane == 0 ? c0 : (lane == 1 ? c1 : c2);
            a.io.pixels_fea[ro * 3 + lane] = c + (p.white_back ? 1.f - wsum : 0.f);
          }
          if (lane == 0 && a.io.depth) a.io.depth[ro] = dd;
          if (act && a.io.weights) a.io.weights[ro * nS + e] = w;
          if (act && a.io.dbg_all_z) a.io.dbg_all_z[ro * nS + e] = ks;
          __syncwarp();
        }
        wsync();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  if (warp == 2) {
    if (PAIR) tmem_dealloc_cg2<512>(tmem);
    else tmem_dealloc<512>(tmem);
  }
}

// ---- prep: weights -> scaled fp16 (hi, lo) UMMA-B tiles in stream order; per-image folded constants
// tile index = (l * 4 + kc) * 2 + (0 hi | 1 lo); element (n, k) of a tile at (n%8)*16 + (n/8)*128 + (k/8)*LBO + (k%8)*2;
// PAIR: the tile is two 16 KB halves (N rows 0..127 | 128..255), each in the layout of a 128-row operand
template <bool PAIR>
__global__ void pigan_prep_weights_kernel(const C3dPiganWeights w, uint8_t* __restrict__ tiles, float4* __restrict__ wl4) {
  const int t = blockIdx.x;                       // (l, kc) pair
  const int l = t / (kH / kKC), kc = t % (kH / kKC);
  const bool colour = l == kLayers - 1;
  const float* W = colour ? w.wc : w.w[l + 1];
  const int ldw = colour ? kH + 3 : kH, k_off = colour ? 3 : 0;
  __half* hi = reinterpret_cast<__half*>(tiles + (size_t)(t * 2) * kTileBytes);
  __half* lo = reinterpret_cast<__half*>(tiles + (size_t)(t * 2 + 1) * kTileBytes);
  for (int i = threadIdx.x; i < kH * kKC; i += blockDim.x) {
    const int n = i / kKC, k = i % kKC;
    const float v = W[(size_t)n * ldw + k_off + kc * kKC + k] * kWScale;
    const __half h = __float2half_rn(v);
    const int nh = PAIR ? n % (kH / 2) : n;
    const int e = ((PAIR ? (n / (kH / 2)) * (kTileBytes / 2) : 0) + (nh % 8) * 16 + (nh / 8) * 128 + (k / 8) * Ring<PAIR>::kLBO_B) / 2 + (k % 8);
    hi[e] = h;
    lo[e] = __float2half_rn(v - __half2float(h));
  }
  if (t == 0)
    for (int k = threadIdx.x; k < kH; k += blockDim.x)
      wl4[k] = make_float4(w.wl[k], w.wl[kH + k], w.wl[2 * kH + k], w.w_sigma[k]);
}

__global__ void pigan_prep_consts_kernel(const C3dPiganWeights w, int B, ImgConsts* __restrict__ consts) {
  const int b = blockIdx.x;
  if (b >= B) return;
  ImgConsts& ic = consts[b];
  const float sc = w.gridwarp ? 2.f / 0.24f : 1.f;
  for (int j = threadIdx.x; j < kH; j += blockDim.x) {
    const float f0 = w.freq[0][(size_t)b * kH + j];
    ic.w0f[j] = make_float4(f0 * sc * w.w[0][j * 3], f0 * sc * w.w[0][j * 3 + 1], f0 * sc * w.w[0][j * 3 + 2],
                            fmaf(f0, w.b[0][j], w.phase[0][(size_t)b * kH + j]));
    for (int l = 1; l <= kLayers; ++l) {      // layers 1..7 and the colour layer (slot kLayers)
      const float f = w.freq[l][(size_t)b * kH + j];
      const float bias = l < kLayers ? w.b[l][j] : w.bc[j];
      ic.film[l - 1][j] = make_float2(f * kWInv, fmaf(f, bias, w.phase[l][(size_t)b * kH + j]));
    }
    const float fc = w.freq[kLayers][(size_t)b * kH + j];
    ic.wdir[j] = make_float4(fc * w.wc[(size_t)j * (kH + 3)], fc * w.wc[(size_t)j * (kH + 3) + 1], fc * w.wc[(size_t)j * (kH + 3) + 2], 0.f);
  }
}

}  // namespace pgt
}  // namespace c3d

using namespace c3d;
using namespace c3d::pgt;

static_assert(offsetof(SmemT<false>, film) - offsetof(SmemT<false>, w0f) == offsetof(ImgConsts, film) &&
                  offsetof(SmemT<false>, wdir) - offsetof(SmemT<false>, w0f) == offsetof(ImgConsts, wdir),
              "the staged image constants are copied as one block");

struct PgWs {
  size_t tiles, wl4, consts, total;
};
static PgWs pg_ws_layout(const C3dRayParams* p) {
  PgWs o;
  size_t off = 0;
  auto take = [&](size_t b) { size_t r = off; off += (b + 255) / 256 * 256; return r; };
  o.tiles = take((size_t)kStreamLayers * kTilesPerLayer * kTileBytes);
  o.wl4 = take((size_t)kH * 16);
  o.consts = take((size_t)(p->batch > 0 ? p->batch : 1) * sizeof(ImgConsts));
  o.total = off;
  return o;
}
size_t c3d_pigan_tc_workspace_bytes(const C3dRayParams* p) { return pg_ws_layout(p).total; }

bool c3d_pigan_tc_supported(const C3dRayParams* p, const C3dPiganWeights* w) {
  const int nS = p->num_steps * (p->hierarchical ? 2 : 1);
  return w->hidden == kH && w->n_layers == kLayers && nS <= 32 && (!p->hierarchical || p->num_steps <= 16);
}

int c3d_pigan_render_fwd_tc(const C3dRayParams* p, const C3dPiganWeights* w, const C3dRayIO* io, int lock_view, void* workspace,
                            size_t workspace_bytes, cudaStream_t st) {
  const PgWs ws = pg_ws_layout(p);
  if (workspace_bytes < ws.total) {
    c3d_set_error("pigan_render(tc): workspace too small (%zu < %zu)", workspace_bytes, ws.total);
    return C3D_EWORKSPACE;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  if (!c3d_device_supported(dev)) {
    c3d_set_error("pigan_render(tc): device %d is not sm_100 (tcgen05 required)", dev);
    return C3D_EARCH;
  }
  uint8_t* base = (uint8_t*)workspace;
  const int sms = c3d_device_sm_count(dev);
  // C3D_PIGAN_PAIR=1: tcgen05 CTA pairs (two ray groups per weight stream).  Opt-in until timed on hardware.
  const bool pair = c3d_options().pigan_pair != 0 && sms >= 2;
  if (pair) C3D_LAUNCH(pigan_prep_weights_kernel<true>, kStreamLayers * (kH / kKC), 256, 0, st, *w, base + ws.tiles, (float4*)(base + ws.wl4));
  else C3D_LAUNCH(pigan_prep_weights_kernel<false>, kStreamLayers * (kH / kKC), 256, 0, st, *w, base + ws.tiles, (float4*)(base + ws.wl4));
  C3D_LAUNCH_CHECK();
  C3D_LAUNCH(pigan_prep_consts_kernel, p->batch, 256, 0, st, *w, p->batch, (ImgConsts*)(base + ws.consts));
  C3D_LAUNCH_CHECK();
  KArgs ka = {};
  ka.p = *p;
  ka.io = *io;
  ka.wtiles = base + ws.tiles;
  ka.consts = (const ImgConsts*)(base + ws.consts);
  ka.wl4 = (const float4*)(base + ws.wl4);
  ka.lock_view = lock_view;
  ka.G = kRows / p->num_steps;
  ka.groups_per_img = (p->n_rays + ka.G - 1) / ka.G;
  ka.total_groups = p->batch * ka.groups_per_img;
  ka.b_sigma = w->b_sigma;
  ka.bl = w->bl;
  if (pair) {
    const size_t smem = sizeof(SmemT<true>) + 1024;
    auto kern = pigan_tc_kernel<true>;
    C3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = (ka.total_groups + 1) / 2 * 2;
    if (grid > sms / 2 * 2) grid = sms / 2 * 2;
    c3d_count_launch();
#ifdef C3D_EMU
    C3D_CUDA(C3D_LAUNCH_CLUSTER(kern, grid, kThreads, smem, st, 2, ka));
#else
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    C3D_CUDA(cudaLaunchKernelEx(&cfg, kern, ka));
#endif
    return C3D_OK;
  }
  const size_t smem = sizeof(SmemT<false>) + 1024;
  C3D_CUDA(cudaFuncSetAttribute(pigan_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = ka.total_groups;
  if (grid > sms) grid = sms;
  if (grid < 1) grid = 1;
  C3D_LAUNCH(pigan_tc_kernel<false>, grid, kThreads, smem, st, ka);
  C3D_LAUNCH_CHECK();
  return C3D_OK;
}
