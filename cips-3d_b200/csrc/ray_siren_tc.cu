// Fused volumetric renderer on tcgen05 (C3D_IMPL_TC): rays -> FiLM-SIREN -> importance
// resampling -> second FiLM-SIREN pass -> merge/sort -> front-to-back compositing, one kernel,
// nothing per-sample ever written to HBM.
//
// Work unit: a ray group (RG) of G = floor(128 / S) whole rays = one 128-row MMA block (rows =
// sample points, ray-major).  A persistent CTA owns two independent TMEM "slots"
// (A operand 128 cols + fp32 accumulator 128 cols each); each slot carries one RG through
//     coarse:  MMA K=16 (layer 0: [x,y,z,1] x per-image folded FiLM weights) -> sin -> MMA 128x128
//              -> FiLM+sin -> MMA 128x(64+sigma) -> FiLM+sin -> MMA 64x32 -> features/sigma to shared memory
//     sample:  per-ray weights -> pdf/cdf -> inverse-CDF -> fine depths     (ray_math.cuh)
//     fine:    the same MLP on the resampled points
//     merge:   per-ray stable sort of the 2S depths, compositing weights
//     output:  sum_i w_i * feature_i  -> (B,N,32)
// while the other slot is in a different phase, so tensor pipe, MUFU (sin) and FMA pipes overlap.
//
// Numerics: hidden activations are sin() outputs in [-1,1]; both MMA operands are split into
// fp16 (hi, lo) pairs and every layer is three tcgen05.mma passes (hi*hi + lo*hi + hi*lo), which
// restores ~fp32 products (hard part 2: single-pass TF32/BF16 misses 1e-3).
// The A operand never touches shared memory: epilogue threads write it straight into TMEM
// (tcgen05.st, two fp16 per 32-bit column, row = lane) and the MMAs use the A-from-TMEM form.
// Weights (fp16 hi/lo, scaled by 2^8 to keep the lo parts normal) sit in shared memory for the
// whole kernel (112 KB, loaded once per CTA by bulk async copies).
//
// Round 2 measured five other scheduling forms of this kernel on B200 (one blocking MMA issuer per slot; K-chunked staged MMA issue;
// an "epilogue turn" that anti-phases the slots; an initial stagger; merge + compositing deferred into the next group's MMA waits
// with jitter / pdf_u prefetch) -- all correct, none faster on the same box (7.18-7.25 ms for this form against 7.41-8.50 ms:
// profiles/r02de_ray_scheduling.md, r02i / r02j_ray_variants.txt; the sources of those forms are commits 64f9cf8 .. fe7aa8d of this
// file).  The finding behind it: TS-form MMAs and a running epilogue's tcgen05.ld/st slow each other down, so tensor time and
// epilogue time add whatever the phase of the two slots; the single polling issuer below happens to keep the slots out of phase.
#include <atomic>
#include <string.h>

#include "c3d_common.cuh"
#include "ray_math.cuh"

namespace c3d {
namespace rtc {

constexpr int kRows = 128;
constexpr float kWScale = 256.f, kWInv = 1.f / 256.f;
constexpr int kN2 = 80;                       // color_layer_sine (64) + sigma head (1) padded to a multiple of 16
// shared-memory weight blob (bytes); every matrix is UMMA-B K-major no-swizzle:
//   element (n,k) at (n%8)*16 + (n/8)*128 + (k/8)*(N*16) + (k%8)*2
constexpr int kW1Bytes = 128 * 128 * 2;       // per hi / lo
constexpr int kW2Bytes = kN2 * 128 * 2;
constexpr int kW3Bytes = 32 * 64 * 2;
constexpr int kOffW1h = 0, kOffW1l = kW1Bytes, kOffW2h = 2 * kW1Bytes, kOffW2l = kOffW2h + kW2Bytes,
              kOffW3h = kOffW2l + kW2Bytes, kOffW3l = kOffW3h + kW3Bytes, kWBlobBytes = kOffW3l + kW3Bytes;
static_assert(kWBlobBytes == 114688, "weight blob is 112 KB");

// per-image epilogue constants (prep kernel): FiLM folded with the linear bias and weight scale
constexpr int kW0Bytes = 128 * 16 * 2;   // layer-0 B operand (N=128, K=16) per hi / lo
struct ImgConsts {
  // layer 0 as a K=16 MMA: A = [x, y, z, 1, 0...], B[j] = [g0*s*W0[j][0..2], g0*b0[j] + beta0[j], 0...]
  // (s = 2/0.24, UniformBoxWarp), fp16 hi then lo, UMMA K-major layout, unscaled
  uint8_t w0[2 * kW0Bytes];
  float2 l1[128];      // (g1/256, g1*b1 + beta1)
  float2 lc[64];       // (gc/256, gc*bc + betac)
};

struct SlotMem {
  float feat[2][kRows][33];   // [0] fine pass, [1] coarse pass features (padded rows: conflict-free both ways)
  float z_c[kRows], sig_c[kRows], z_f[kRows], sig_f[kRows];
  float wc[kRows];            // coarse compositing weights
  float cdf[kRows];           // per ray: S-1 cdf entries (stride S)
  float fbuf[2 * kRows];      // 1 - alpha + 1e-10 per (sorted) sample
  float abuf[2 * kRows];      // alpha per (sorted) sample
  float skey[2 * kRows];      // sorted depths, ray g at [g*nS, (g+1)*nS)
  int sidx[2 * kRows];        // source of each sorted sample: < S fine, >= S coarse
  float w_all[2 * kRows];     // final compositing weights (sorted order)
  int frow[2 * kRows];        // row of each sorted sample in feat[0..1] viewed as [2*kRows][33]
  float wsum[16];             // per ray: sum of weights (before last_back)
  alignas(128) uint8_t w0[2 * kW0Bytes];  // image constants of the group's image (ImgConsts)
  float2 l1[128];
  float2 lc[64];
};

struct Smem {
  alignas(1024) uint8_t w[kWBlobBytes];
  SlotMem slot[2];
  alignas(8) uint64_t w_full;
  uint64_t a_ready[2];
  uint64_t d_ready[2];
  uint32_t tmem_base;
};

struct KArgs {
  C3dRayParams p;
  C3dRayIO io;
  const uint8_t* wblob;       // kWBlobBytes, prepped
  const ImgConsts* consts;    // (B)
  const float* b_sigma;       // (1) final_layer bias
  const float* bl;            // (32) color_layer_linear bias
  int G;                      // rays per group
  int groups_per_img, total_groups;
  const float* w_sigma;       // (128) final_layer weight, fp32 (MATH = 2 only: sigma head in the E1 epilogue)
};

__device__ __forceinline__ float fast_sin(float x) { return __sinf(x); }

#ifdef C3D_TRACE   // debug build: block 0 stamps the phases of iterations 3 and 4.  One fixed slot per (iteration parity, warp,
// stamp index): a plain fire-and-forget store -- round 2's first trace used an atomic counter whose ~500 clk round trip sat on
// the critical path of every stamp (issuer wake-up -> atomic -> MMA issue) and inflated each phase by about 1 k clk.
__device__ unsigned long long g_rtrace[2 * 20 * 64];
__device__ __forceinline__ void rtrace(int it, uint32_t tag, uint32_t a0, int& n) {
  if (blockIdx.x == 0 && (it == 3 || it == 4)) {
    const int w = threadIdx.x >> 5;
    if (n < 64) g_rtrace[((it & 1) * 20 + w) * 64 + n] = ((unsigned long long)tag << 56) | ((unsigned long long)(a0 & 0xFFFF) << 40) | (clock64() & 0xFFFFFFFFFFull);
    ++n;
  } else if (it != 3 && it != 4) {
    n = 0;
  }
}
#define RTRACE(it, tag, a0) rtrace(it, tag, a0, tr_n)
#else
#define RTRACE(it, tag, a0)
#endif

// ---- write 16 consecutive K values (fp32) of this thread's row into the TMEM A operand (hi, lo)
__device__ __forceinline__ void store_a16(uint32_t a_hi_col, uint32_t a_lo_col, const float (&v)[16]) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) split_f16(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
  tmem_st8(a_hi_col, hi);
  tmem_st8(a_lo_col, lo);
}

// ---- three-pass split-precision MMA:  D = A_hi*B_hi + A_lo*B_hi + A_hi*B_lo   (A in TMEM).
// b_*_lo: low descriptor words of the hi / lo weight matrices (umma_desc_lo); K-step = 2 core columns.
template <int N, int K>
__device__ __forceinline__ void mma_split3(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi_lo,
                                           uint32_t b_lo_lo, uint32_t dhi) {
  constexpr uint32_t idesc = umma_idesc_f16(kRows, N);
  constexpr uint32_t kstep = (2u * N * 16u) >> 4;
#pragma unroll
  for (int k = 0; k < K / 16; ++k) umma_ts_w(d_tmem, a_hi + 8 * k, b_hi_lo + kstep * k, dhi, idesc, k != 0);
#pragma unroll
  for (int k = 0; k < K / 16; ++k) umma_ts_w(d_tmem, a_lo + 8 * k, b_hi_lo + kstep * k, dhi, idesc, 1);
#pragma unroll
  for (int k = 0; k < K / 16; ++k) umma_ts_w(d_tmem, a_hi + 8 * k, b_lo_lo + kstep * k, dhi, idesc, 1);
}

// s1.14 fixed point through the float adder: x + 1.5*2^9 has ulp 2^-14 for |x| <= 1, so the low 16 bits of its pattern
// relative to 0x44400000 are round(x * 2^14) as an int16 (RN-even, like the adder)
__device__ __forceinline__ uint32_t pack_q14(float x0, float x1) {
  const uint32_t b0 = __float_as_uint(__fadd_rn(x0, 768.f)), b1 = __float_as_uint(__fadd_rn(x1, 768.f));
  return (b0 & 0xFFFFu) | (b1 << 16);
}
__device__ __forceinline__ float unpack_q14_lo(uint32_t u) {
  return __fsub_rn(__uint_as_float(0x44400000u + (uint32_t)((int32_t)(u << 16) >> 16)), 768.f);
}
__device__ __forceinline__ float unpack_q14_hi(uint32_t u) {
  return __fsub_rn(__uint_as_float(0x44400000u + (uint32_t)((int32_t)u >> 16)), 768.f);
}

// alpha of one sample (pigan_utils.py:246-251): 1 - exp(-delta * clamp(sigma + noise))
__device__ __forceinline__ float sample_alpha(float delta, float sigma, float noise, int clamp_mode) {
  const float sn = __fadd_rn(sigma, noise);
  const float act = clamp_mode == 1 ? softplus_f32(sn) : fmaxf(sn, 0.f);
  return __fsub_rn(1.f, expf(__fmul_rn(-delta, act)));
}

// MATH = 0: per-ray math by all 256 threads of the slot on shared-memory arrays, block barriers between the steps (the
//           kernel measured in round 1).
// MATH = 1: per-ray math by ONE WARP PER RAY with shuffles (lane = sample): prefix products / sums are warp scans, the
//           stable rank sort is nS shuffles, compositing shuffles the weights -- the ~10 slot-wide barriers and the
//           serial shared-memory loops of MATH 0 (28 % of a ray group's time in the round-1 trace) disappear.  Needs
//           2S <= 32 (hierarchical: two rays share a warp in the resampling step).  C3D_RAY_MATH=warp; emulation-verified,
//           not yet timed on hardware.
// MATH = 2: MATH 1 plus the two heads taken off the tensor-core chain.  sigma = final_layer(h1) is a 128-term fp32 dot
//           product accumulated in the layer-1 epilogue while the sines are still in registers (two threads per row, 64
//           columns each), so the colour MMA shrinks from N = 80 to N = 64.  color_layer_linear (64 -> 32) is linear, so it
//           commutes with compositing: the 64 colour sines of every sample stay in shared memory (s1.14 fixed point) and
//           the per-ray warp composites them and applies Wl once per RAY instead of once per SAMPLE -- the fourth MMA phase
//           and its epilogue (a d_ready round trip per pass on the serial per-slot chain) are gone.  The per-point debug
//           outputs do not exist in this form (the host falls back to MATH 1).  C3D_RAY_MATH=fold; emulation-verified,
//           not yet timed on hardware.
template <int MATH>
__global__ void __launch_bounds__(640, 1) ray_siren_tc_kernel(const KArgs a) {
  C3D_DYN_SMEM(uint8_t, smem_raw);
  Smem& s = *reinterpret_cast<Smem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const C3dRayParams& p = a.p;
  const int S = p.num_steps, G = a.G;
  const bool hier = p.hierarchical != 0;
  const int nS = hier ? 2 * S : S;
  constexpr bool WARP = MATH >= 1, FOLD = MATH == 2;
  constexpr int kPassPhases = FOLD ? 3 : 4;  // per pass: layer 0, layer 1, colour(+sigma), [colour linear]
  constexpr int kNc = FOLD ? 64 : kN2;       // N of the colour-sine MMA
  const int mma_phases = hier ? 2 * kPassPhases : kPassPhases;

  if (threadIdx.x == 0) {
    mbar_init(&s.w_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s.a_ready[i], 8);
      mbar_init(&s.d_ready[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(&s.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s.tmem_base;
  const int iters = (a.total_groups + 2 * (int)gridDim.x - 1) / (2 * (int)gridDim.x);

  if (warp < 4) {
    reg_dec<56>();  // 128*32 + 512*112 == 640*96: setmaxnreg only recycles this CTA's own registers
    if (warp == 0) {
      if (lane == 0) {   // weights: one 112 KB bulk load per CTA
        mbar_arrive_expect_tx(&s.w_full, kWBlobBytes);
        for (int off = 0; off < kWBlobBytes; off += 16384) bulk_g2s(s.w + off, a.wblob + off, 16384, &s.w_full);
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------- MMA issuer: the whole warp stays converged
      // (operands live in uniform registers, no per-MMA lane loop); one elected lane issues for whichever slot
      // is ready.
      mbar_wait(&s.w_full, 0);
      const uint32_t wb = smem_u32(s.w);
      const uint32_t dhi = umma_desc_hi(128);
      const uint32_t w1h = umma_desc_lo(wb + kOffW1h, 128 * 16), w1l = umma_desc_lo(wb + kOffW1l, 128 * 16);
      const uint32_t w2h = umma_desc_lo(wb + kOffW2h, kNc * 16), w2l = umma_desc_lo(wb + kOffW2l, kNc * 16);
      const uint32_t w3h = umma_desc_lo(wb + kOffW3h, 32 * 16), w3l = umma_desc_lo(wb + kOffW3l, 32 * 16);
      uint32_t par[2] = {0, 0};
      int done[2] = {0, 0};
#ifdef C3D_TRACE
      int tr_n = 0;
#endif
      const int total = iters * mma_phases;
      uint32_t idle = 0;
      unsigned long long idle_t0 = 0;
      while (done[0] < total || done[1] < total) {
        if ((++idle & 0xFFFFu) == 0 && (idle_t0 == 0 ? (idle_t0 = c3d_globaltimer(), false)
                                                      : c3d_globaltimer() - idle_t0 > C3D_WATCHDOG_NS)) {
          if (lane == 0) printf("c3d watchdog: ray MMA issuer starved (block %d, done %d/%d of %d)\n", (int)blockIdx.x, done[0], done[1], total);
          __trap();
        }
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
          if (done[sl] < total && __all_sync(0xffffffffu, mbar_test(&s.a_ready[sl], par[sl]))) {
            par[sl] ^= 1;
            tc_fence_after();
            const int layer = FOLD ? done[sl] % 3 : (done[sl] & 3);
            if (lane == 0) RTRACE(done[sl] / mma_phases, 1, (uint32_t)(sl << 15 | ((done[sl] / mma_phases) & 1) << 8 | (done[sl] % mma_phases)));
            if (elect_one()) {
              uint32_t a_hi = tmem + (uint32_t)(sl * 256), a_lo = a_hi + 64;
              const uint32_t d = a_hi + 128;
              // opaque per-iteration copies: keeps the compiler from hoisting ~60 loop-invariant descriptor
              // words out of the loop (they would spill and cost an LDL per MMA)
              const uint32_t w0h = umma_desc_lo(smem_u32(s.slot[sl].w0), 128 * 16);
              uint32_t bh = layer == 0 ? w0h : (layer == 1 ? w1h : (layer == 2 ? w2h : w3h));
              uint32_t bl = layer == 0 ? w0h + (kW0Bytes >> 4) : (layer == 1 ? w1l : (layer == 2 ? w2l : w3l));
              asm volatile("" : "+r"(bh), "+r"(bl), "+r"(a_hi), "+r"(a_lo));
              if (layer == 0) mma_split3<128, 16>(d, a_hi, a_lo, bh, bl, dhi);
              else if (layer == 1) mma_split3<128, 128>(d, a_hi, a_lo, bh, bl, dhi);
              else if (layer == 2) mma_split3<kNc, 128>(d, a_hi, a_lo, bh, bl, dhi);
              else if (!FOLD) mma_split3<32, 64>(d, a_hi, a_lo, bh, bl, dhi);
              tc_commit(&s.d_ready[sl]);
            }
            __syncwarp();
            ++done[sl];
            idle = 0;
            idle_t0 = 0;
          }
        }
        if (idle) __nanosleep(64);   // nothing ready: yield the issue port to the workers on this scheduler
      }
    }
  } else {
    // ------------------------------------------------------------ slot workers (epilogue + ray math)
    reg_inc<104>();
    const int ew = warp - 4;             // 0..15
    const int sl = (ew >> 2) & 1;        // slot
    const int half = ew >> 3;            // column half
    const int q = warp & 3;
    const int row = q * 32 + lane;       // point row of the group
    const int stid = half * 128 + row;   // 0..255 within the slot
    const int tw = q + 4 * half;         // warp 0..7 within the slot's team
    SlotMem& sm = s.slot[sl];
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const uint32_t a_hi = tmem + (uint32_t)(sl * 256) + lane_sel, a_lo = a_hi + 64, dcol = a_hi + 128;
    const int bar_id = 1 + sl;
    auto slot_sync = [&]() { named_bar_sync_n<256>(bar_id); };
    uint32_t dpar = 0;
    auto signal_a = [&]() {
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.a_ready[sl]);
    };
    auto wait_d = [&]() {
      mbar_wait(&s.d_ready[sl], dpar);
      dpar ^= 1;
      tc_fence_after();
    };
#ifdef C3D_TRACE
    int tr_it = 0, tr_ph = 0, tr_n = 0;
    // trace word: slot << 15 | team warp << 12 | iteration parity << 8 | phase counter (lane 0 of every worker warp stamps)
    auto stamp = [&](uint32_t tag) { if (lane == 0) RTRACE(tr_it, tag, (uint32_t)(sl << 15 | tw << 12 | (tr_it & 1) << 8 | tr_ph)); ++tr_ph; };
#else
    auto stamp = [](uint32_t) {};
#endif
    const int g_row = row / S, s_row = row - g_row * S;   // ray within the group, sample index
    const bool row_in_group = g_row < G;
    const int g_el = stid / nS, e_el = stid - g_el * nS;  // (ray, element) view used by the merge phases
    // sigma of a coarse / fine row: MATH 2 keeps the two half-row partial sums of the layer-1 epilogue (see E1)
    const float bsig = FOLD ? __ldg(a.b_sigma) : 0.f;
    auto sig_at = [&](bool fine, int r) -> float {
      if (FOLD) {
        const float* ps = fine ? sm.w_all : sm.fbuf;
        return __fadd_rn(__fadd_rn(ps[r], ps[kRows + r]), bsig);
      }
      return (fine ? sm.sig_f : sm.sig_c)[r];
    };
    int cur_img = -1;
    if (FOLD && stid < 128) sm.abuf[stid] = __ldg(a.w_sigma + stid);   // visible after the first image-constants barrier
    if (FOLD) mbar_wait(&s.w_full, 0);   // the workers read Wl^T from the bulk-loaded blob themselves: observe its barrier
#ifdef C3D_RAY_STAGGER_NS   // start slot 1 half a pass late so its MMA phases fall into slot 0's worker phases
    if (sl == 1) __nanosleep(C3D_RAY_STAGGER_NS);
#endif

    for (int it = 0; it < iters; ++it) {
#ifdef C3D_TRACE
      tr_it = it; tr_ph = 0;
#endif
      const int grp = (it * (int)gridDim.x + (int)blockIdx.x) * 2 + sl;
      const bool grp_ok = grp < a.total_groups;
      const int img = grp_ok ? grp / a.groups_per_img : 0;
      const int ray0 = grp_ok ? (grp % a.groups_per_img) * G : 0;          // first local ray of the group
      const int n_valid = grp_ok ? min(G, p.n_rays - ray0) : 0;            // rays of this group that exist
      const bool pt_ok = row_in_group && g_row < n_valid;
      const int nloc = ray0 + g_row;                                       // local ray index (output slot)
      const size_t ro_row = (size_t)img * p.n_rays + nloc;
      const float* M = a.io.cam2world + (size_t)img * 16;
      if (img != cur_img) {    // (slot-uniform) stage the image's folded FiLM constants
        const ImgConsts& ic = a.consts[img];
        reinterpret_cast<uint4*>(sm.w0)[stid] = __ldg(reinterpret_cast<const uint4*>(ic.w0) + stid);           // 8 KB
        reinterpret_cast<uint4*>(sm.w0)[stid + 256] = __ldg(reinterpret_cast<const uint4*>(ic.w0) + stid + 256);
        if (stid < 128) sm.l1[stid] = __ldg(&ic.l1[stid]);
        else if (stid < 192) sm.lc[stid - 128] = __ldg(&ic.lc[stid - 128]);
        cur_img = img;
        fence_proxy_async();      // w0 is read by the tensor core (async proxy)
        slot_sync();
      }
      RayFrame fr;
      int gray = 0;
      if (pt_ok) {
        gray = a.io.ray_idx ? a.io.ray_idx[nloc] : p.ray_offset + nloc;
        fr = make_ray_frame(M, gray, p.img_size, p.z_cam);
      }

      for (int pass = 0; pass < (hier ? 2 : 1); ++pass) {
        // ---------------- L0: point position -> A operand [x, y, z, 1] (K = 16) for the layer-0 MMA
        float px = 0.f, py = 0.f, pz = 0.f;
        if (pt_ok) {
          if (pass == 0) {
            const float u = a.io.jitter_u[((size_t)img * p.img_size * p.img_size + gray) * S + s_row];
            float z;
            coarse_sample(fr, M, p.ray_start, p.ray_end, S, s_row, u, z, px, py, pz);
            if (half == 0) sm.z_c[row] = z;
          } else {
            fine_sample(fr, sm.z_f[row], px, py, pz);
          }
        }
        if (half == 0) {
          float v[16] = {px, py, pz, 1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          store_a16(a_hi, a_lo, v);
        }
        stamp(2);
        signal_a();
        // ---------------- E0: D(128) = g0*(W0 p*s + b0) + beta0  ->  sin  ->  A (h0)
        wait_d();
        stamp(3);
        {
          uint32_t accA[16], accB[16];
          auto e0 = [&](const uint32_t (&acc)[16], int c) {
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fast_sin(__uint_as_float(acc[j]));
            store_a16(a_hi + (uint32_t)(half * 32 + c / 2), a_lo + (uint32_t)(half * 32 + c / 2), v);
          };
          tmem_ld16(dcol + (uint32_t)(half * 64), accA);
          tc_wait_ld();
          tmem_ld16(dcol + (uint32_t)(half * 64 + 16), accB);
          e0(accA, 0);
          tc_wait_ld();
          tmem_ld16(dcol + (uint32_t)(half * 64 + 32), accA);
          e0(accB, 16);
          tc_wait_ld();
          tmem_ld16(dcol + (uint32_t)(half * 64 + 48), accB);
          e0(accA, 32);
          tc_wait_ld();
          e0(accB, 48);
        }
        stamp(4);
        signal_a();
        // ---------------- E1: D(128) -> FiLM+sin -> A (h1); TMEM loads double-buffered
        wait_d();
        stamp(5);
        {
          uint32_t accA[16], accB[16];
          float psig = 0.f;       // MATH 2: this thread's 64 columns of final_layer (fp32, before the fp16 split)
          auto e1 = [&](const uint32_t (&acc)[16], int c) {
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float2 g2 = sm.l1[half * 64 + c + j];
              v[j] = fast_sin(fmaf(__uint_as_float(acc[j]), g2.x, g2.y));
              if (FOLD) psig = fmaf(v[j], sm.abuf[half * 64 + c + j], psig);
            }
            store_a16(a_hi + (uint32_t)(half * 32 + c / 2), a_lo + (uint32_t)(half * 32 + c / 2), v);
          };
          tmem_ld16(dcol + (uint32_t)(half * 64), accA);
          tc_wait_ld();
          tmem_ld16(dcol + (uint32_t)(half * 64 + 16), accB);
          e1(accA, 0);
          tc_wait_ld();
          tmem_ld16(dcol + (uint32_t)(half * 64 + 32), accA);
          e1(accB, 16);
          tc_wait_ld();
          tmem_ld16(dcol + (uint32_t)(half * 64 + 48), accB);
          e1(accA, 32);
          tc_wait_ld();
          e1(accB, 48);
          // partial sums of the two column halves, per pass; combined where sigma is used, i.e. after the slot-wide barrier
          // that ends the pass (fbuf / w_all are free in the warp-math forms)
          if (FOLD) (pass == 0 ? sm.fbuf : sm.w_all)[half * kRows + row] = psig;
        }
        stamp(6);
        signal_a();
        // ---------------- E2: D(80): cols 0..63 -> FiLM+sin -> A (h2, K=64); col 64 -> sigma
        wait_d();
        stamp(7);
        float sigma = 0.f;
        if (FOLD) {
          // colour sines stay in shared memory as s1.14 fixed point (the feat area viewed as [2*kRows][33] pairs:
          // |error| <= 2^-15, eight times tighter than fp16 near +-1 where sines spend their time); the 64 -> 32
          // color_layer_linear is applied once per ray to the composited sines (linear in the features)
          uint32_t accA[16], accB[16];
          tmem_ld16(dcol + (uint32_t)(half * 32), accA);
          tmem_ld16(dcol + (uint32_t)(half * 32 + 16), accB);
          tc_wait_ld();
          uint32_t* crow = reinterpret_cast<uint32_t*>(&sm.feat[0][0][0]) + ((pass == 0 ? kRows : 0) + row) * 33 + half * 16;
          auto e2 = [&](const uint32_t (&acc)[16], int c) {
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              const float2 g0 = sm.lc[half * 32 + c + j], g1 = sm.lc[half * 32 + c + j + 1];
              crow[(c + j) >> 1] = pack_q14(fast_sin(fmaf(__uint_as_float(acc[j]), g0.x, g0.y)),
                                            fast_sin(fmaf(__uint_as_float(acc[j + 1]), g1.x, g1.y)));
            }
          };
          e2(accA, 0);
          e2(accB, 16);
          stamp(8);
          stamp(9);
        } else {
          uint32_t accS[16];
          uint32_t accA[16], accB[16];
          auto e2 = [&](const uint32_t (&acc)[16], int c) {
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float2 g2 = sm.lc[half * 32 + c + j];
              v[j] = fast_sin(fmaf(__uint_as_float(acc[j]), g2.x, g2.y));
            }
            store_a16(a_hi + (uint32_t)(half * 16 + c / 2), a_lo + (uint32_t)(half * 16 + c / 2), v);
          };
          tmem_ld16(dcol + (uint32_t)(half * 32), accA);
          tmem_ld16(dcol + (uint32_t)(half * 32 + 16), accB);
          if (half == 1) tmem_ld16(dcol + 64u, accS);
          tc_wait_ld();
          if (half == 1) {
            sigma = fmaf(__uint_as_float(accS[0]), kWInv, __ldg(a.b_sigma));
            (pass == 0 ? sm.sig_c : sm.sig_f)[row] = sigma;
          }
          e2(accA, 0);
          e2(accB, 16);
        }
        if (!FOLD) {
        stamp(8);
        signal_a();
        // ---------------- E3: D(32) -> + bias -> features to shared memory
        wait_d();
        stamp(9);
        {
          uint32_t acc[16];
          tmem_ld16(dcol + (uint32_t)(half * 16), acc);
          tc_wait_ld();
          float(*feat)[33] = pass == 0 ? sm.feat[1] : sm.feat[0];
          float* dbg = pass == 0 ? a.io.dbg_coarse : a.io.dbg_fine;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float f = fmaf(__uint_as_float(acc[j]), kWInv, __ldg(a.bl + half * 16 + j));
            feat[row][half * 16 + j] = f;
            if (dbg && pt_ok) dbg[(ro_row * S + s_row) * kOutC + half * 16 + j] = f;
          }
          if (dbg && pt_ok && half == 1) dbg[(ro_row * S + s_row) * kOutC + kFeat] = sigma;
        }
        }  // !FOLD
        tc_fence_before();
        slot_sync();
        stamp(10);
        // ---------------- importance resampling, all threads (generator_nerf_inr.py:537-598, pigan_utils.py:164-209)
        if (WARP && pass == 0 && hier) {
          // ---- warp per ray, two rays per warp (lanes 0-15 / 16-31; S <= 16): lane e = coarse sample e
          const unsigned full = 0xffffffffu;
          const int sub = lane >> 4, e = lane & 15, ns = S - 2;
          for (int g0 = 0; g0 < n_valid; g0 += 16) {
            const int g = g0 + tw + 8 * sub;
            const bool act = g < n_valid && e < S;
            const int r0 = g * S;
            const size_t ro = (size_t)img * p.n_rays + ray0 + g;
            const float z = act ? sm.z_c[r0 + e] : 0.f;
            const float zn = __shfl_down_sync(full, z, 1, 16);
            float alpha = 0.f, f = 1.f;
            if (act) {
              const float delta = e + 1 < S ? __fsub_rn(zn, z) : 1e10f;
              const float nz = a.io.noise_c ? __fmul_rn(a.io.noise_c[ro * S + e], p.noise_std) : 0.f;
              alpha = sample_alpha(delta, sig_at(false, r0 + e), nz, p.clamp_mode);
              f = __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f);
            }
            float P = f;                                  // inclusive product scan over the ray's lanes
#pragma unroll
            for (int d = 1; d < 16; d <<= 1) {
              const float t = __shfl_up_sync(full, P, d, 16);
              if (e >= d) P = __fmul_rn(P, t);
            }
            float T = __shfl_up_sync(full, P, 1, 16);
            if (e == 0) T = 1.f;
            const float w = __fmul_rn(alpha, T);          // coarse compositing weight (pigan_utils.py:256-257)
            const float wn = __shfl_down_sync(full, w, 1, 16);
            const float wt = e <= S - 3 ? __fadd_rn(__fadd_rn(wn, 1e-5f), 1e-5f) : 0.f;   // (w + 1e-5)[1:-1] + 1e-5
            float sum = wt;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(full, sum, o, 16);
            const float pdf = __fmul_rn(wt, __fdividef(1.f, sum));
            float C = pdf;                                // cdf_j = sum_{i<j} pdf_i, j = 0..S-2
#pragma unroll
            for (int d = 1; d < 16; d <<= 1) {
              const float t = __shfl_up_sync(full, C, d, 16);
              if (e >= d) C = __fadd_rn(C, t);
            }
            float cdf = __shfl_up_sync(full, C, 1, 16);
            if (e == 0) cdf = 0.f;
            const float uk = act ? a.io.pdf_u[ro * S + e] : 0.f;
            int i = 0;                                    // searchsorted(cdf, u, right=False): the cdf is non-decreasing
            for (int j = 0; j <= ns; ++j) i += __shfl_sync(full, cdf, j, 16) < uk ? 1 : 0;
            const int below = max(i - 1, 0), above = min(i, ns);
            const float cb = __shfl_sync(full, cdf, below, 16), ca = __shfl_sync(full, cdf, above, 16);
            const float bb = 0.5f * __fadd_rn(__shfl_sync(full, z, below, 16), __shfl_sync(full, z, below + 1, 16));
            const float ba = 0.5f * __fadd_rn(__shfl_sync(full, z, above, 16), __shfl_sync(full, z, above + 1, 16));
            float denom = __fsub_rn(ca, cb);
            if (denom < 1e-5f) denom = 1.f;
            if (act) sm.z_f[r0 + e] = __fadd_rn(bb, __fmul_rn(__fdiv_rn(__fsub_rn(uk, cb), denom), __fsub_rn(ba, bb)));
          }
          slot_sync();
        } else if (pass == 0 && hier) {
          const int r0 = g_row * S;
          float alpha = 0.f;
          if (half == 0 && pt_ok) {   // A: alpha_i and (1 - alpha_i + 1e-10)
            const float delta = s_row + 1 < S ? __fsub_rn(sm.z_c[row + 1], sm.z_c[row]) : 1e10f;
            const float nz = a.io.noise_c ? __fmul_rn(a.io.noise_c[ro_row * S + s_row], p.noise_std) : 0.f;
            alpha = sample_alpha(delta, sm.sig_c[row], nz, p.clamp_mode);
            sm.fbuf[row] = __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f);
          }
          slot_sync();
          if (half == 0 && pt_ok) {   // B: T_i = prod_{j<i} f_j (sequential order), w_i = alpha_i * T_i
            float T = 1.f;
            for (int j = 0; j < s_row; ++j) T = __fmul_rn(T, sm.fbuf[r0 + j]);
            sm.wc[row] = __fmul_rn(alpha, T);
          }
          slot_sync();
          if (half == 0 && pt_ok && s_row <= S - 2) {   // C: cdf_j, j = 0..S-2, over weights (w+1e-5)[1:-1]+1e-5
            float sum = 0.f;
            for (int j = 0; j < S - 2; ++j) sum += __fadd_rn(__fadd_rn(sm.wc[r0 + j + 1], 1e-5f), 1e-5f);
            float c = 0.f;
            const float inv = __fdividef(1.f, sum);      // pdf_j = wt_j / sum (1-2 ulp; cdf only feeds a 2e-4-conditioned inverse)
            for (int j = 0; j < s_row; ++j)
              c = __fadd_rn(c, __fmul_rn(__fadd_rn(__fadd_rn(sm.wc[r0 + j + 1], 1e-5f), 1e-5f), inv));
            sm.cdf[r0 + s_row] = c;
          }
          slot_sync();
          if (half == 0 && pt_ok) {   // D: inverse CDF for u_k, k = s_row
            const int ns = S - 2;
            const float uk = a.io.pdf_u[ro_row * S + s_row];
            int i = 0;
            while (i <= ns && sm.cdf[r0 + i] < uk) ++i;         // searchsorted(cdf, u, right=False)
            const int below = max(i - 1, 0), above = min(i, ns);
            const float cb = sm.cdf[r0 + below], ca = sm.cdf[r0 + above];
            const float bb = 0.5f * __fadd_rn(sm.z_c[r0 + below], sm.z_c[r0 + below + 1]);
            const float ba = 0.5f * __fadd_rn(sm.z_c[r0 + above], sm.z_c[r0 + above + 1]);
            float denom = __fsub_rn(ca, cb);
            if (denom < 1e-5f) denom = 1.f;
            sm.z_f[row] = __fadd_rn(bb, __fmul_rn(__fdiv_rn(__fsub_rn(uk, cb), denom), __fsub_rn(ba, bb)));
          }
          slot_sync();
        }
      }
      stamp(11);
      // ---------------- merge (stable rank sort of the nS depths of each ray), generator.py:1733-1738
      if (WARP) {
        // ---- warp per ray: lane e = element e of cat([fine, coarse]) for the sort, lane = channel for the compositing
        const unsigned full = 0xffffffffu;
        const float* featf = &sm.feat[0][0][0];
        for (int g = tw; g < n_valid; g += 8) {          // warp-uniform
          const int rc0 = g * S, base = g * nS, e = lane;
          const bool act = e < nS;
          const size_t ro = (size_t)img * p.n_rays + ray0 + g;
          const float k = act ? (hier ? (e < S ? sm.z_f[rc0 + e] : sm.z_c[rc0 + e - S]) : sm.z_c[rc0 + e]) : 3.0e38f;
          int rank = 0;                                   // stable ascending rank (torch.sort over cat([fine, coarse]))
          for (int j = 0; j < nS; ++j) {
            const float kj = __shfl_sync(full, k, j);
            rank += (kj < k || (kj == k && j < e)) ? 1 : 0;
          }
          if (act) {
            sm.skey[base + rank] = k;
            sm.sidx[base + rank] = hier ? e : S + e;
          }
          __syncwarp();
          const float ks = act ? sm.skey[base + e] : 0.f;
          const int src = act ? sm.sidx[base + e] : S;
          const float kn = __shfl_down_sync(full, ks, 1);
          float alpha = 0.f, f = 1.f;
          if (act) {
            const float sg = src < S ? sig_at(true, rc0 + src) : sig_at(false, rc0 + src - S);
            const float delta = e + 1 < nS ? __fsub_rn(kn, ks) : 1e10f;
            const float nz = a.io.noise_f ? __fmul_rn(a.io.noise_f[ro * nS + e], p.noise_std) : 0.f;
            alpha = sample_alpha(delta, sg, nz, p.clamp_mode);
            f = __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f);
          }
          float P = f;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const float t = __shfl_up_sync(full, P, d);
            if (e >= d) P = __fmul_rn(P, t);
          }
          float T = __shfl_up_sync(full, P, 1);
          if (e == 0) T = 1.f;
          float w = __fmul_rn(alpha, T);
          float wsum = w;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(full, wsum, o);
          if (p.last_back && e == nS - 1) w += 1.f - wsum;
          const int frow = src < S ? rc0 + src : kRows + rc0 + src - S;
          float acc = 0.f;
          if (FOLD) {
            // composite the 64 colour sines (lane = sine pair), then color_layer_linear once per ray:
            //   sum_i w_i (Wl c_i + bl) = Wl (sum_i w_i c_i) + bl sum_i w_i
            const uint32_t* cq = reinterpret_cast<const uint32_t*>(featf);
            float a0 = 0.f, a1 = 0.f;
#pragma unroll 4
            for (int i = 0; i < nS; ++i) {
              const float wi = __shfl_sync(full, w, i);
              const uint32_t u = cq[__shfl_sync(full, frow, i) * 33 + lane];
              a0 = fmaf(wi, unpack_q14_lo(u), a0);
              a1 = fmaf(wi, unpack_q14_hi(u), a1);
            }
            const float* wlt = reinterpret_cast<const float*>(s.w + kOffW3h);   // fp32 Wl^T [64][32] (prep, fold mode)
            acc = __ldg(a.bl + lane) * (p.last_back ? 1.f : wsum);
#pragma unroll 8
            for (int k2 = 0; k2 < 32; ++k2) {
              acc = fmaf(wlt[(2 * k2) * 32 + lane], __shfl_sync(full, a0, k2), acc);
              acc = fmaf(wlt[(2 * k2 + 1) * 32 + lane], __shfl_sync(full, a1, k2), acc);
            }
          } else {
#pragma unroll 4
            for (int i = 0; i < nS; ++i)
              acc = fmaf(__shfl_sync(full, w, i), featf[__shfl_sync(full, frow, i) * 33 + lane], acc);
          }
          if (p.white_back) acc += 1.f - wsum;
          a.io.pixels_fea[ro * kFeat + lane] = acc;
          if (act && a.io.weights) a.io.weights[ro * nS + e] = w;
          if (act && a.io.dbg_all_z) a.io.dbg_all_z[ro * nS + e] = ks;
          if (a.io.depth) {
            float d = act ? __fmul_rn(w, ks) : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(full, d, o);
            if (lane == 0) a.io.depth[ro] = d;
          }
          __syncwarp();
        }
        slot_sync();
        stamp(12);
        stamp(13);
        continue;
      }
      const bool el_ok = g_el < n_valid;
      const int rc0 = g_el * S;           // first row of ray g_el
      const int base = g_el * nS;
      const size_t ro_el = (size_t)img * p.n_rays + ray0 + g_el;
      if (el_ok) {
        auto key_of = [&](int e) { return hier ? (e < S ? sm.z_f[rc0 + e] : sm.z_c[rc0 + e - S]) : sm.z_c[rc0 + e]; };
        const float k = key_of(e_el);
        int rank = 0;
        for (int e = 0; e < nS; ++e) {
          const float ke = key_of(e);
          rank += (ke < k || (ke == k && e < e_el)) ? 1 : 0;
        }
        const int src = hier ? e_el : S + e_el;
        sm.skey[base + rank] = k;
        sm.sidx[base + rank] = src;
        sm.frow[base + rank] = src < S ? rc0 + src : kRows + rc0 + src - S;
      }
      slot_sync();
      float alpha_el = 0.f;
      if (el_ok) {       // alpha of sorted position e_el (pigan_utils.py:241-251)
        const int src = sm.sidx[base + e_el];
        const float sg = src < S ? sm.sig_f[rc0 + src] : sm.sig_c[rc0 + src - S];
        const float delta = e_el + 1 < nS ? __fsub_rn(sm.skey[base + e_el + 1], sm.skey[base + e_el]) : 1e10f;
        const float nz = a.io.noise_f ? __fmul_rn(a.io.noise_f[ro_el * nS + e_el], p.noise_std) : 0.f;
        alpha_el = sample_alpha(delta, sg, nz, p.clamp_mode);
        sm.fbuf[base + e_el] = __fadd_rn(__fsub_rn(1.f, alpha_el), 1e-10f);
      }
      slot_sync();
      if (el_ok) {       // transmittance in the reference's cumprod order, weight
        float T = 1.f;
        for (int j = 0; j < e_el; ++j) T = __fmul_rn(T, sm.fbuf[base + j]);
        sm.w_all[base + e_el] = __fmul_rn(alpha_el, T);
      }
      slot_sync();
      stamp(12);
      // ---------------- composite: pixels_fea[ray][c] = sum_i w_i * feature_i[c]  (pigan_utils.py:255-266)
      if (el_ok && e_el == 0) {      // per ray: weight sum in the reference's order, last_back folded into the last weight
        float wsum = 0.f;
        for (int i = 0; i < nS; ++i) wsum += sm.w_all[base + i];
        sm.wsum[g_el] = wsum;
        if (p.last_back) sm.w_all[base + nS - 1] += 1.f - wsum;
      }
      slot_sync();
      {
        const int c = stid & 31;
        const float* featf = &sm.feat[0][0][0];
        for (int g = stid >> 5; g < n_valid; g += 8) {      // warp = ray, lane = channel
          const int b0 = g * nS;
          float acc = 0.f;
#pragma unroll 4
          for (int i = 0; i < nS; ++i) acc = fmaf(sm.w_all[b0 + i], featf[sm.frow[b0 + i] * 33 + c], acc);
          if (p.white_back) acc += 1.f - sm.wsum[g];
          const size_t ro = (size_t)img * p.n_rays + ray0 + g;
          a.io.pixels_fea[ro * kFeat + c] = acc;
          if (c < nS && (a.io.weights || a.io.dbg_all_z)) {     // lanes cover the nS <= 32 fast case, loop otherwise
            for (int i = c; i < nS; i += 32) {
              if (a.io.weights) a.io.weights[ro * nS + i] = sm.w_all[b0 + i];
              if (a.io.dbg_all_z) a.io.dbg_all_z[ro * nS + i] = sm.skey[b0 + i];
            }
          }
          if (c == 0 && a.io.depth) {
            float depth = 0.f;
            for (int i = 0; i < nS; ++i) depth = fmaf(sm.w_all[b0 + i], sm.skey[b0 + i], depth);
            a.io.depth[ro] = depth;
          }
        }
      }
      slot_sync();
      stamp(13);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem);
}

// ---- prep: fp32 weights -> scaled fp16 (hi, lo) UMMA-B blobs; per-image folded FiLM constants
__device__ __forceinline__ void put_split(uint8_t* blob, int off_hi, int off_lo, int N, int n, int k, float v) {
  const __half h = __float2half_rn(v);
  const __half l = __float2half_rn(v - __half2float(h));
  const int e = ((n % 8) * 16 + (n / 8) * 128 + (k / 8) * (N * 16)) / 2 + (k % 8);
  reinterpret_cast<__half*>(blob + off_hi)[e] = h;
  reinterpret_cast<__half*>(blob + off_lo)[e] = l;
}

// fold != 0 (MATH 2): the colour-sine matrix alone as N = 64, and fp32 Wl^T [64][32] where the W3 tiles would be
__global__ void ray_prep_kernel(C3dSirenWeights w, int B, uint8_t* blob, ImgConsts* consts, int fold) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int i = tid; i < 128 * 128; i += nth) put_split(blob, kOffW1h, kOffW1l, 128, i / 128, i % 128, w.w1[i] * kWScale);
  if (fold) {
    for (int i = tid; i < 64 * 128; i += nth) put_split(blob, kOffW2h, kOffW2l, 64, i / 128, i % 128, w.wc[i] * kWScale);
    for (int i = tid; i < 32 * 64; i += nth) reinterpret_cast<float*>(blob + kOffW3h)[(i % 64) * 32 + i / 64] = w.wl[i];
  } else {
    for (int i = tid; i < kN2 * 128; i += nth) {
      const int n = i / 128, k = i % 128;
      const float v = n < 64 ? w.wc[n * 128 + k] : (n == 64 ? w.w_sigma[k] : 0.f);
      put_split(blob, kOffW2h, kOffW2l, kN2, n, k, v * kWScale);
    }
    for (int i = tid; i < 32 * 64; i += nth) put_split(blob, kOffW3h, kOffW3l, 32, i / 64, i % 64, w.wl[i] * kWScale);
  }
  const float sc = 2.f / 0.24f;
  for (int i = tid; i < B * 128; i += nth) {
    const int b = i / 128, j = i % 128;
    const float g0 = w.gamma0[i], g1 = w.gamma1[i];
    const float c4[4] = {g0 * sc * w.w0[j * 3], g0 * sc * w.w0[j * 3 + 1], g0 * sc * w.w0[j * 3 + 2],
                         fmaf(g0, w.b0[j], w.beta0[i])};
    for (int k = 0; k < 16; ++k) put_split(consts[b].w0, 0, kW0Bytes, 128, j, k, k < 4 ? c4[k] : 0.f);
    consts[b].l1[j] = make_float2(g1 * kWInv, fmaf(g1, w.b1[j], w.beta1[i]));
  }
  for (int i = tid; i < B * 64; i += nth) {
    const int b = i / 64, j = i % 64;
    const float gc = w.gammac[i];
    consts[b].lc[j] = make_float2(gc * kWInv, fmaf(gc, w.bc[j], w.betac[i]));
  }
}

}  // namespace rtc
}  // namespace c3d

using namespace c3d;
using namespace c3d::rtc;

static std::atomic<int> g_ray_math_mode{-1};
extern "C" int c3d_debug_ray_math_mode(void) { return g_ray_math_mode.load(); }

struct RayWs {
  size_t blob, consts, total;
};
static RayWs ray_ws_layout(const C3dRayParams* p) {
  RayWs o;
  size_t off = 0;
  auto take = [&](size_t b) { size_t r = off; off += (b + 255) / 256 * 256; return r; };
  o.blob = take(kWBlobBytes);
  o.consts = take((size_t)(p->batch > 0 ? p->batch : 1) * sizeof(ImgConsts));
  o.total = off;
  return o;
}
size_t c3d_ray_siren_tc_workspace_bytes(const C3dRayParams* p) { return ray_ws_layout(p).total; }

#ifdef C3D_TRACE
extern "C" int c3d_debug_ray_trace(unsigned long long* out, int cap) {
  static unsigned long long host[2 * 20 * 64];
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(host, c3d::rtc::g_rtrace, sizeof(host));
  int n = 0;
  for (int i = 0; i < 2 * 20 * 64 && n < cap; ++i)
    if (host[i]) out[n++] = host[i];
  memset(host, 0, sizeof(host));
  cudaMemcpyToSymbol(c3d::rtc::g_rtrace, host, sizeof(host));
  return n;
}
#endif

int c3d_ray_siren_fwd_tc(const C3dRayParams* p, const C3dSirenWeights* w, const C3dRayIO* io, void* workspace,
                         size_t workspace_bytes, cudaStream_t st) {
  const RayWs ws = ray_ws_layout(p);
  if (workspace_bytes < ws.total) {
    c3d_set_error("ray_siren(tc): workspace too small (%zu < %zu)", workspace_bytes, ws.total);
    return C3D_EWORKSPACE;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  if (!c3d_device_supported(dev)) {
    c3d_set_error("ray_siren(tc): device %d is not sm_100 (tcgen05 required)", dev);
    return C3D_EARCH;
  }
  const int sms = c3d_device_sm_count(dev);
  uint8_t* base = (uint8_t*)workspace;
  // C3D_RAY_MATH=warp: warp-per-ray math (needs 2S <= 32 samples per warp); =fold: warp math + the sigma head in the E1
  // epilogue and color_layer_linear applied after compositing (3 MMA phases per pass instead of 4; not with the
  // per-point debug outputs); default: the round-1 block-wide form
  const int rm = c3d_options().ray_math;
  const bool warp_ok = p->num_steps * (p->hierarchical ? 2 : 1) <= 32 && (!p->hierarchical || p->num_steps <= 16);
  const bool fold_math = rm == 2 && warp_ok && !io->dbg_coarse && !io->dbg_fine;
  const bool warp_math = rm >= 1 && warp_ok;
  C3D_LAUNCH(ray_prep_kernel, 64, 256, 0, st, *w, p->batch, base + ws.blob, (ImgConsts*)(base + ws.consts), fold_math ? 1 : 0);
  C3D_LAUNCH_CHECK();
  KArgs ka = {};
  ka.p = *p;
  ka.io = *io;
  ka.wblob = base + ws.blob;
  ka.consts = (const ImgConsts*)(base + ws.consts);
  ka.bl = w->bl;
  ka.G = kRows / p->num_steps;
  ka.groups_per_img = (p->n_rays + ka.G - 1) / ka.G;
  ka.total_groups = p->batch * ka.groups_per_img;
  ka.b_sigma = w->b_sigma;
  ka.w_sigma = w->w_sigma;
  const size_t smem = sizeof(Smem) + 1024;
  static std::atomic<unsigned long long> attr_set{0};     // per device, once
  if (!(attr_set.load() >> (dev & 63) & 1ull)) {
    C3D_CUDA(cudaFuncSetAttribute(ray_siren_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    C3D_CUDA(cudaFuncSetAttribute(ray_siren_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    C3D_CUDA(cudaFuncSetAttribute(ray_siren_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set.fetch_or(1ull << (dev & 63));
  }
  int grid = (ka.total_groups + 1) / 2;
  if (grid > sms) grid = sms;
  if (grid < 1) grid = 1;
  g_ray_math_mode.store(fold_math ? 2 : (warp_math ? 1 : 0));
  if (fold_math) C3D_LAUNCH(ray_siren_tc_kernel<2>, grid, 640, smem, st, ka);
  else if (warp_math) C3D_LAUNCH(ray_siren_tc_kernel<1>, grid, 640, smem, st, ka);
  else C3D_LAUNCH(ray_siren_tc_kernel<0>, grid, 640, smem, st, ka);
  C3D_LAUNCH_CHECK();
  return C3D_OK;
}
