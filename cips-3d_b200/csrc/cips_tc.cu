// Fused per-pixel CIPS synthesis MLP on tcgen05 (C3D_IMPL_TC).
//
// Two forms of one kernel (template parameter PAIR).  The default wherever an image has an even number of tiles is the CTA PAIR
// (tcgen05 cta_group::2: M = 256 x N = 256 MMAs issued by one thread of the leader CTA, each CTA streaming its half of the weights;
// see the PAIR branch and DESIGN.md 4.2b for the measurements that led there); the single-CTA form described next serves odd tile
// counts and the training forward (activation stash).  Both share the tile walk, the staircase and the two-phase epilogue.
//
// One persistent CTA per SM walks 128-pixel tiles (all pixels of a tile belong to one image).
// For a tile, the whole 18-layer chain  x -> [mod-linear 512x512 + demod + LeakyReLU]x2 (+skip)
// -> ToRGB ... -> tanh  runs on-chip:
//   * activations live in shared memory as the fp16 A operand (128 x 512, UMMA K-major
//     no-swizzle layout, 128 KB) and are overwritten in place by each layer's epilogue;
//   * the fp32 accumulator (128 x 512) fills all 512 TMEM columns;
//   * weights are pre-tiled fp16 blobs (64 K x 128 N, 16 KB) streamed through a 5-stage
//     shared-memory ring by bulk async copies (TMA engine, UBLKCP), optionally multicast to
//     every CTA of a thread-block cluster so that L2 sees each byte once per cluster;
//   * warp 0 = weight producer, warps 1 and 3 = tcgen05.mma issuers (each owns two of the four 128-column
//     accumulator blocks; one elected lane issues), warps 4..19 = epilogue (TMEM -> registers: LeakyReLU,
//     residual, ToRGB, fp16 pack -> shared memory).
//   * software pipeline between layers, in both directions ("staircase"): layer l+1's MMAs start as soon
//     as the epilogue of layer l has produced the K-chunk they read and drained the accumulator block they
//     overwrite (epi_done[j]); the epilogue of layer l starts on column chunk j as soon as accumulator
//     block j is complete and no MMA of layer l still reads A-operand chunk j, which it overwrites in
//     place (acc_ready[j]).  The tile order inside a layer is chosen to make both happen early.
//   * both issuers observe every fill of the weight ring and a stage is recycled only when both have
//     released it (an issuer that skipped the phases of tiles it does not own could alias a whole ring
//     revolution on the parity wait).
// Per-image modulation (mod_conv_fc.py:452-496): the prep kernel builds, once per forward and per image,
// fp16 tiles of  W''[k][n] = s1p[k] * W[k][n] * d[n]  (the reference's modulated + demodulated weight);
// all 512 pixel tiles of an image stream the same 9.3 MB, so the epilogue needs no per-column constants
// (they would have to come from shared memory, whose bandwidth the MMA operand fetch already saturates).  Only HBM traffic: x (128 B/pixel) in, rgb (12 B/pixel) out; the residual stream
// of the skip blocks goes through an L2-resident per-CTA scratch (fp32, 256 KB; fp16 when only the image is returned: ResT below).
#include <atomic>
#include <utility>
#include <string.h>

#include "c3d_common.cuh"

namespace c3d {
namespace cips {

constexpr int kH = 512;            // hidden width
constexpr int kTileM = 128;        // pixels per tile
constexpr int kKC = 64;            // K per weight tile
constexpr int kNC = 128;           // N per weight tile
constexpr int kWTileBytes = kKC * kNC * 2;              // 16 KB
#ifndef C3D_CIPS_STAGES
#define C3D_CIPS_STAGES 5
#endif
constexpr int kStages = C3D_CIPS_STAGES;      // depth of the weight ring (5 x 16 KB is what fits beside the 128 KB activation tile)
// CTA-pair variant (PAIR, tcgen05 cta_group::2): an MMA is M = 256 (128 pixels in each CTA of the pair) x N = 256, and each CTA
// holds -- and streams from L2 -- the 128 N rows of its half: a ring stage is again one 16 KB (kc, nc) tile, CTA r holding
// (kc, 2h + r) of the "big tile" (kc, h).  Per CTA and layer that is half the weight traffic and half the A-operand reads of the
// single-CTA kernel, and 16 issue steps instead of 32 (see the PAIR branch of the kernel).
template <bool PAIR> struct RingCfg {
  static constexpr int kStagesP = kStages;
  static constexpr int kStageBytes = kWTileBytes;
  static constexpr int kLBO_B = kNC * 16;     // K-direction core-matrix stride of the B tile (128 rows)
};
constexpr int kXBytes = kTileM * kH * 2;                // 128 KB
constexpr int kLBO = kTileM * 16;                       // 2048: K-direction core-matrix stride (A and B tiles have 128 rows)
constexpr int kSBO = 128;                               // M/N-direction stride between 8-row groups
constexpr int kNumEpiWarps = 16;
constexpr int kThreads = 32 * (4 + kNumEpiWarps);       // 640
constexpr int kMaxLayers = C3D_CIPS_MAX_LAYERS;

template <bool PAIR>
struct SmemT {
  static constexpr int NS = RingCfg<PAIR>::kStagesP;
  alignas(1024) uint8_t x[kXBytes];
  alignas(1024) uint8_t w[kStages][kWTileBytes];
  union {                                // ToRGB weights of the current block / per-tile rgb partial sums
    float4 rgbw[kH];
    float rgb_part[4][kTileM][4];
  };
  alignas(8) uint64_t full[NS];
  uint64_t empty[NS];
  uint64_t epi_done[4];
  uint64_t acc_ready[4];   // accumulator block j complete AND A-operand chunk j no longer read by this layer's MMAs
  uint32_t tmem_base;
  uint32_t pad_;
};
using Smem = SmemT<false>;

struct KArgs {
  const float* x;            // (B,N,in_dim)
  float* rgb;                // (B,N,3)
  float* hidden_out;         // (B,N,512) or null
  const __half* wtiles;      // prepped weights: image-major, then layer-major, tiles in stream order (16 KB each)
  size_t img_tile_stride;    // tiles per image
  const float4* rgbw;        // (n_blocks,512) float4 (w0,w1,w2,0) ; valid for blocks >= rgb_from
  const float* rgbb;         // (3) summed ToRGB biases
  float4* resid;             // (gridDim.x, 128, 128) float4 scratch
  int B, N, in_dim, n_layers, skip_from, rgb_from, tiles_per_img, total_tiles;
  int layer_tile_off[kMaxLayers + 1];   // offset (in tiles) of each layer's first weight tile
  int layer_kc[kMaxLayers];             // number of K chunks of each layer (1 for the padded input layer)
  // Issue order of a full layer's 32 weight tiles ("staircase"): entry = kc | nc << 4 | need << 8 | rdy << 12.
  // `need` is the epilogue chunk of the previous layer that must be complete (input K-chunk written, accumulator
  // columns drained); bit j of `rdy` says the tile's owner commits acc_ready[j] after it.  Tiles are stored in
  // this order so the producer streams linearly.
  uint16_t order_full[32];
  uint16_t order_in[4];
  // training (DUMP instantiation only): every layer's output y_l (after LeakyReLU and residual) as the fp16 values the next
  // layer consumed, (n_layers, B, N, 512) row-major, for the backward pass (cips_bwd_tc.cu).  Appended: the offsets of the
  // fields above are those of the round-1 kernel.
  __half* acts;
  size_t acts_layer_stride;      // elements between consecutive layers = B * N * 512
  // sign bits of z_l for the layers that add a residual (there y_l - y_{l-2} cannot recover the sign of a small lrelu(z_l)
  // from two fp16-rounded stashes): (n_layers, B, N, 32) uint16, bit i of word c = (z_l[16 c + i] > 0)
  uint16_t* zsign;
  // A/B knob (C3D_CIPS_STAGGER_NS): CTA i starts i * stagger_ns late, so that the 148 CTAs -- identical work on the same image's 9.4 MB
  // of weights -- do not all stream the same tile at the same moment
  int stagger_ns;
  // -DC3D_CIPS_ABLATE builds only (tools/build_ablate_lib.sh; timing experiments, results are garbage): bit 0 = epilogue without its
  // TMEM reads / math / stores, bit 1 = issuers commit without issuing MMAs, bit 2 = producer signals the stages without loading,
  // bit 3 (with bit 1, single CTA) = plain mbarrier arrives instead of tcgen05.commit, bit 4 = no epilogue warps at all (the issuers do
  // not wait for them): the weight ring alone; epilogue parts: bit 5 = no TMEM loads, bit 6 = no A-operand stores, bit 7 = no residual
  // loads, bit 8 = no ToRGB, bit 9 = no residual stores
  int ablate;
};

template <int CL>
__device__ __forceinline__ void commit_stage_free(uint64_t* bar) {
  if (CL == 1) {
    tc_commit(bar);
  } else {
    tc_commit_mc(bar, (uint16_t)((1u << CL) - 1));
  }
}
// non-owner issuer: "I have observed this fill" -> one plain arrive on the stage's empty barrier of every CTA
// that multicasts into it
template <int CL>
__device__ __forceinline__ void observe_stage_free(uint64_t* bar) {
  if (CL == 1) {
    mbar_arrive(bar);
  } else {
#pragma unroll
    for (uint32_t r = 0; r < (uint32_t)CL; ++r) mbar_arrive_cluster(bar, r);
  }
}
template <int CL>
__device__ __forceinline__ void load_w_tile(void* dst, const uint8_t* src, uint64_t* bar, uint32_t rank) {
  if (CL == 1) {
    bulk_g2s(dst, src, kWTileBytes, bar);
  } else {
    constexpr uint32_t slice = kWTileBytes / CL;
    bulk_g2s_mc((uint8_t*)dst + rank * slice, src + rank * slice, slice, bar, (uint16_t)((1u << CL) - 1));
  }
}

__device__ __forceinline__ float lrelu02(float v) { return fmaxf(v, 0.2f * v); }

// Who probes an mbarrier while a warp waits.  0 (the product): every lane; 1: lane 0 probes, the warp joins it at a __syncwarp;
// 2: as 1 and the 16 epilogue warps wait for their accumulator block behind ONE probing lane and a named barrier.  Measured
// (profiles/r02u_cips_poll.txt, r02v_mbar_bench.txt): a hand-off to a single probing lane is 2-3x slower than to a full warp
// (ping-pong 695 vs 271 clk per round trip; the whole kernel 23.6 vs 11.8 ms), and 16 extra polling warps cost the ring nothing.
#ifndef C3D_CIPS_POLL
#define C3D_CIPS_POLL 0
#endif
__device__ __forceinline__ void wait_warp(uint64_t* bar, uint32_t parity, int lane) {
#if C3D_CIPS_POLL == 0
  mbar_wait(bar, parity);
#else
  if (lane == 0) mbar_wait(bar, parity);
  __syncwarp();
#endif
}
__device__ __forceinline__ void wait_warp_cluster(uint64_t* bar, uint32_t parity, int lane) {
#if C3D_CIPS_POLL == 0
  mbar_wait_cluster(bar, parity);
#else
  if (lane == 0) mbar_wait_cluster(bar, parity);
  __syncwarp();
#endif
}
// epilogue: all 16 warps wait for the same barrier
__device__ __forceinline__ void wait_epilogue(uint64_t* bar, uint32_t parity, int warp, int lane) {
#if C3D_CIPS_POLL == 2
  if (warp == 4 && lane == 0) mbar_wait(bar, parity);
  named_bar_sync_c<2, kNumEpiWarps * 32>();
#else
  wait_warp(bar, parity, lane);
#endif
}
__device__ __forceinline__ bool test_warp(uint64_t* bar, uint32_t parity, int lane) {
#if C3D_CIPS_POLL == 0
  return __all_sync(0xffffffffu, mbar_test(bar, parity)) != 0;
#elif C3D_CIPS_POLL == 2
  return false;      // every warp must reach the named barrier of wait_epilogue
#else
  return __shfl_sync(0xffffffffu, lane == 0 ? (int)mbar_test(bar, parity) : 0, 0) != 0;
#endif
}

#ifdef C3D_TRACE   // debug build: blocks 0 and 1 stamp the pipeline events of tile iteration 1 (steady state).  One fixed slot per
// (block, warp, stamp index): plain stores, no atomics (an atomic per stamp costs ~500 clk on the stamping warp's critical path).
constexpr int kTraceCap = 3072;
__device__ unsigned long long g_trace[2 * 20 * kTraceCap];
__device__ __forceinline__ void trace_ev(int it, uint32_t tag, uint32_t a0, int& n) {
  if (blockIdx.x < 2 && it == 1) {
    if (n < kTraceCap)
      g_trace[((size_t)blockIdx.x * 20 + (threadIdx.x >> 5)) * kTraceCap + n] =
          ((unsigned long long)tag << 56) | ((unsigned long long)(a0 & 0xFFFF) << 40) | (clock64() & 0xFFFFFFFFFFull);
    ++n;
  }
}
#define TRACE(it, tag, a0) trace_ev(it, tag, a0, tr_n)
#define TRACE_DECL int tr_n = 0; (void)tr_n
// -DC3D_TRACE=1 ("light"): the per-tile stamps of the producer / issuers only at tiles 0, 18 (first of the tail the previous layer's
// last epilogue chunk unlocks), 22 and 31, and the epilogue stamps by one warp only -- a stamp costs ~170 clk on the stamping warp, and
// three per tile make the issuers the bottleneck of the traced CTA (profiles/r02r_cips_trace_pair.txt)
#if C3D_TRACE + 0 == 1
#define TRACE_TILE(t) ((t) == 0 || (t) == 6 || (t) == 10 || (t) == 15 || (t) == 18 || (t) == 22 || (t) == 31)
#define TRACE_EPI_WARP(w) ((w) == 4)
#else
#define TRACE_TILE(t) true
#define TRACE_EPI_WARP(w) true
#endif
#else
#define TRACE(it, tag, a0)
#define TRACE_DECL
#define TRACE_TILE(t) false
#define TRACE_EPI_WARP(w) false
#endif

// The residual stream y_b (block output, added to the output of block b + 1) between its two layers lives in a per-CTA scratch that
// stays in L2.  fp32: 512 KB of L2 traffic per CTA and residual layer (256 KB in, 256 KB out) beside the 256 KB of weights -- those
// layers are L2-bandwidth-bound (a residual layer took ~21 k clk, a plain one ~13 k: profiles/r02ac_cips_light_pair_l8.txt).
// R16 keeps it as fp16 -- the very bits the next layer's A operand holds, so the skip connection adds what the MMAs saw -- and halves
// that.  Image error vs the fp64 oracle is unchanged by it (2.8-3.3e-4 against 2.5-3.3e-4 on the CPU emulation), the error of the
// 512-wide hidden state grows by a quarter (up to 1.2e-3): the launcher uses R16 only when no hidden output / activation stash is asked for.
template <bool R16> struct ResT;
template <> struct ResT<false> { using Vec = float4; static constexpr int kVecs = 4; };     // vectors per 16-column slice
template <> struct ResT<true> { using Vec = uint4; static constexpr int kVecs = 2; };       // 8 halves each

struct EpiFlags {
  bool add_res, keep_res, do_rgb, last;
  bool no_x = false;      // -DC3D_CIPS_ABLATE timing experiments only: skip the A-operand stores
};

// One thread, 16 accumulator columns of its row (the per-image scales are already inside the weights), in two phases:
//   epi16_act : what the NEXT LAYER'S MMAs wait for -- a = fp16(y) into the A operand -- and nothing else;
//   epi16_tail: everything else (residual store, ToRGB, the last layer's outputs, the training stash), after the chunk has been
//               handed to the issuer.  It recomputes y from the same registers (a max, a multiply and an add per value), so what
//               it stores is bit-identical to what epi16_act packed.
// Why two phases: fence.proxy.async (needed between the A-operand stores and the hand-over) is MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC in
// SASS and waits for the thread's outstanding GLOBAL stores as well; with the residual stores in front of it a chunk of a
// residual + ToRGB layer took 4-5 k clk against 1.3 k for a plain one (profiles/r02aa_cips_light_pair_l8.txt), and the epilogue chain
// is the critical path of a layer.  Removing the residual stores alone: -1.4 ms of 8.9, the ToRGB FMAs: -0.75 ms
// (profiles/r02ab_cips_pair_residual.txt).
//   first layer of a block : a = lrelu(acc)
//   second layer of a block: y = lrelu(acc) (+ residual); ToRGB += y.Wrgb; a = y
//   rwp: ToRGB weights of column c; xp: the thread's 16-byte slot of K-group c/8 in the A operand; rp: residual scratch (float4
//   index c/4, this row).
__device__ __forceinline__ void add_residual(float (&y)[16], const float4 (&rs)[4]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    y[4 * g + 0] += rs[g].x; y[4 * g + 1] += rs[g].y; y[4 * g + 2] += rs[g].z; y[4 * g + 3] += rs[g].w;
  }
}
__device__ __forceinline__ void add_residual(float (&y)[16], const uint4 (&rs)[2]) {
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const uint32_t w[4] = {rs[g].x, rs[g].y, rs[g].z, rs[g].w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 v = __half22float2(*reinterpret_cast<const __half2*>(&w[q]));
      y[8 * g + 2 * q] += v.x;
      y[8 * g + 2 * q + 1] += v.y;
    }
  }
}
__device__ __forceinline__ void store_residual(float4* rp, const float (&y)[16]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) rp[g * kTileM] = make_float4(y[4 * g], y[4 * g + 1], y[4 * g + 2], y[4 * g + 3]);
}
__device__ __forceinline__ void store_residual(uint4* rp, const float (&y)[16]) {
#pragma unroll
  for (int g = 0; g < 2; ++g)
    rp[g * kTileM] = make_uint4(pack_f16(y[8 * g], y[8 * g + 1]), pack_f16(y[8 * g + 2], y[8 * g + 3]), pack_f16(y[8 * g + 4], y[8 * g + 5]),
                                pack_f16(y[8 * g + 6], y[8 * g + 7]));
}
template <bool SECOND, typename RV, int NV>
__device__ __forceinline__ void epi16_y(const uint32_t (&acc)[16], const RV (&rs)[NV], const EpiFlags f, float (&y)[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) y[i] = lrelu02(__uint_as_float(acc[i]));
  if (SECOND && f.add_res) add_residual(y, rs);
}
template <bool SECOND, typename RV, int NV>
__device__ __forceinline__ void epi16_act(const uint32_t (&acc)[16], const RV (&rs)[NV], uint8_t* xp, const EpiFlags f) {
  if (SECOND && f.last) return;      // the last layer feeds no MMA
  float y[16];
  epi16_y<SECOND>(acc, rs, f, y);
  uint32_t pk[8];
#pragma unroll
  for (int g = 0; g < 8; ++g) pk[g] = pack_f16(y[2 * g], y[2 * g + 1]);
#ifdef C3D_CIPS_ABLATE
  if (!f.no_x)
#endif
  {
    *reinterpret_cast<uint4*>(xp) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    *reinterpret_cast<uint4*>(xp + kLBO) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  }
}
template <bool SECOND, bool DUMP, typename RV, int NV>
__device__ __forceinline__ void epi16_tail(const uint32_t (&acc)[16], const RV (&rs)[NV], const float4* __restrict__ rwp, RV* rp,
                                           const EpiFlags f, float& rgb0, float& rgb1, float& rgb2, float* hid_out,
                                           uint4* dump = nullptr, uint16_t* zsign = nullptr) {
  if (!SECOND && !DUMP) return;
  float y[16];
  epi16_y<SECOND>(acc, rs, f, y);
  if (DUMP && SECOND && zsign && f.add_res) {
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) m |= (__uint_as_float(acc[i]) > 0.f ? 1u : 0u) << i;
    *zsign = (uint16_t)m;
  }
  if (SECOND) {
    if (f.keep_res) store_residual(rp, y);
    if (f.do_rgb) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float4 w4 = rwp[i];
        rgb0 = fmaf(y[i], w4.x, rgb0);
        rgb1 = fmaf(y[i], w4.y, rgb1);
        rgb2 = fmaf(y[i], w4.z, rgb2);
      }
    }
    if (f.last && hid_out) {
#pragma unroll
      for (int g = 0; g < 4; ++g)
        reinterpret_cast<float4*>(hid_out)[g] = make_float4(y[4 * g], y[4 * g + 1], y[4 * g + 2], y[4 * g + 3]);
    }
  }
  if (DUMP && dump) {
    uint32_t pk[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) pk[g] = pack_f16(y[2 * g], y[2 * g + 1]);
    dump[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    dump[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  }
}

// ---- PAIR kernel: the issue order of a layer as COMPILE-TIME constants.
// Entry = kc | h << 4 | need << 8 | acc_ready commit mask << 12 (build_pair_order on the host derives the same tables from the dependency
// rules and lays the weights out in that order; c3d_cips_fwd_tc refuses to launch if the two ever disagree).  The issuer's step
// is the kernel's critical instruction stream: it shares its scheduler with four epilogue warps, and at ~110 instructions a step (table
// look-ups, decoding, generic loops) it stretched from ~500 to 750-900 clk whenever the epilogue was busy (light trace
// profiles/r02ae_cips_light_pair_l8.txt; in isolation neither the epilogue's TMEM loads, its shared / global stores, its proxy fences
// nor the weight stream slow the tensor pipe by a single clock: profiles/r02ai_contention_bench.txt).  With every field a template
// argument a step is a wait, an election, four MMAs and a commit.
constexpr uint16_t pair_entry(int kc, int h, int rdy) {
  return (uint16_t)(kc | (h << 4) | ((kc / 2 > 2 * h + 1 ? kc / 2 : 2 * h + 1) << 8) | (rdy << 12));
}
constexpr uint16_t kPairFull[16] = {
    pair_entry(0, 0, 0), pair_entry(1, 0, 0), pair_entry(2, 0, 0), pair_entry(3, 0, 0), pair_entry(4, 0, 0), pair_entry(5, 0, 0),
    pair_entry(6, 0, 0), pair_entry(7, 0, 0), pair_entry(0, 1, 0), pair_entry(1, 1, 1), pair_entry(2, 1, 0), pair_entry(3, 1, 2),
    pair_entry(4, 1, 0), pair_entry(5, 1, 0), pair_entry(6, 1, 0), pair_entry(7, 1, 4 | 8)};
constexpr uint16_t kPairIn[2] = {pair_entry(0, 0, 2), pair_entry(0, 1, 1 | 4 | 8)};

template <bool PAIR>
struct PairIssuer {      // state of the leader's issuer warp
  SmemT<PAIR>* s;
  uint32_t a_lo0, b_lo0, tmem, stage, phase, epar;
  int lane, it, l, abl;
#ifdef C3D_TRACE
  int tr_n;
#endif
};
// One step: big tile E; PREV_NEED = the epilogue chunks of the previous layer already waited for in this layer (-1: none).
template <uint16_t E, int PREV_NEED, int IDX, bool PAIR>
__device__ __forceinline__ void pair_step(PairIssuer<PAIR>& c) {
  constexpr uint32_t kc = E & 15u, h = (E >> 4) & 15u, rdy = (E >> 12) & 15u;
  constexpr int need = (E >> 8) & 15;
  constexpr int NS = SmemT<PAIR>::NS;
  constexpr uint32_t idesc = umma_idesc_f16(2 * kTileM, 2 * kNC);
  constexpr uint32_t dhi = umma_desc_hi(kSBO);
  constexpr uint32_t kStepK16 = (2 * kLBO) >> 4;       // one K=16 MMA step  = 2 core-matrix columns
  constexpr uint32_t kStepStage = kWTileBytes >> 4;
  SmemT<PAIR>& s = *c.s;
#ifdef C3D_TRACE
  int& tr_n = c.tr_n;
#endif
  if (c.lane == 0 && TRACE_TILE(IDX)) TRACE(c.it, 1, (uint32_t)(c.l << 8 | IDX));          // tile reached
  if (need > PREV_NEED && !(c.abl & 16)) {
#pragma unroll
    for (int j = PREV_NEED + 1; j <= need; ++j) wait_warp_cluster(&s.epi_done[j], c.epar, c.lane);     // half of the arrivals come from the peer CTA
  }
  if (c.lane == 0 && TRACE_TILE(IDX)) TRACE(c.it, 3, (uint32_t)(c.l << 8 | IDX));          // epilogue dependency satisfied
  wait_warp(&s.full[c.stage], c.phase, c.lane);
  if (c.lane == 0 && TRACE_TILE(IDX)) TRACE(c.it, 5, (uint32_t)(c.l << 8 | IDX));          // weight tile landed (both halves)
  tc_fence_after();
  if (elect_one()) {
    const uint32_t a_lo = c.a_lo0 + kc * (kStepK16 * (kKC / 16));
    const uint32_t b_lo = c.b_lo0 + c.stage * kStepStage;
    const uint32_t d = c.tmem + h * (2 * kNC);
    if (!(c.abl & 2)) {
      umma_ss_w_cg2(d, a_lo, b_lo, dhi, idesc, kc != 0);
      umma_ss_w_cg2(d, a_lo + kStepK16, b_lo + kStepK16, dhi, idesc, 1);
      umma_ss_w_cg2(d, a_lo + 2 * kStepK16, b_lo + 2 * kStepK16, dhi, idesc, 1);
      umma_ss_w_cg2(d, a_lo + 3 * kStepK16, b_lo + 3 * kStepK16, dhi, idesc, 1);
    }
    tc_commit_cg2_mc(&s.empty[c.stage], 3);      // stage free in BOTH CTAs
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (rdy & (1u << j)) tc_commit_cg2_mc(&s.acc_ready[j], 3);      // S_j complete: accumulator block j done, A chunk j no longer read
  }
  __syncwarp();
  if (++c.stage == (uint32_t)NS) { c.stage = 0; c.phase ^= 1; }
}
template <bool PAIR, size_t... I>
__device__ __forceinline__ void pair_layer_full(PairIssuer<PAIR>& c, std::index_sequence<I...>) {
  (pair_step<kPairFull[I], (I == 0 ? -1 : (int)((kPairFull[I == 0 ? 0 : I - 1] >> 8) & 15)), (int)I, PAIR>(c), ...);
}
template <bool PAIR>
__device__ __forceinline__ void pair_layer_in(PairIssuer<PAIR>& c) {
  pair_step<kPairIn[0], -1, 0, PAIR>(c);
  pair_step<kPairIn[1], (int)((kPairIn[0] >> 8) & 15), 1, PAIR>(c);
}

// PAIR (CL == 2 only): the two CTAs of a cluster form a tcgen05 CTA pair.  Only the leader (cluster rank 0) issues MMAs
// (cta_group::2, M = 256: the leader's 128 pixels and the peer's 128 pixels against the same 256 weight columns); each CTA
// streams ITS half of every big weight tile into its own ring; the peer relays "my half landed" to the leader's full barrier;
// both CTAs' epilogue warps report to the leader's epi_done barriers; the leader's commits are multicast to the
// empty / acc_ready barriers of both CTAs.  The epilogue (4 chunks of 128 columns a layer) is the single-CTA kernel's.
template <int CL, bool PAIR = false, bool DUMP = false, bool R16 = false>
__global__ void __launch_bounds__(kThreads, 1) cips_tc_kernel(const KArgs a) {
  using ResVec = typename ResT<R16>::Vec;
  constexpr int kResVecs = ResT<R16>::kVecs, kResCols = 16 / kResVecs;
  static_assert(!PAIR || CL == 2, "a CTA pair is a cluster of two");
  using SM = SmemT<PAIR>;
  using RC = RingCfg<PAIR>;
  constexpr int NS = RC::kStagesP;
  C3D_DYN_SMEM(uint8_t, smem_raw);
  // identical offset in every CTA of a cluster (multicast lands at the same CTA-relative address)
  SM& s = *reinterpret_cast<SM*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0;
  const bool leader = !PAIR || crank == 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) {
      // PAIR, leader: a fill of stage i is complete when the leader's own half has landed (its producer's arrive + bytes) AND the
      // peer has relayed "my half landed" with one remote arrive on this SAME barrier -- the issuer waits on one CTA-local barrier.
      mbar_init(&s.full[i], (PAIR && leader) ? 2 : 1);
#ifdef C3D_INJECT_RING_RACE           // test-only (tests/test_emu_cpu.py): re-creates the round-1 parity-aliasing race
      mbar_init(&s.empty[i], CL);
#else
      // both issuers release every stage (see the issuer loop); PAIR: one issuer, whose multicast commit releases the stage in both CTAs
      mbar_init(&s.empty[i], PAIR ? 1 : 2 * CL);
#endif
    }
    for (int i = 0; i < 4; ++i) mbar_init(&s.epi_done[i], PAIR ? 2 * kNumEpiWarps : kNumEpiWarps);
    for (int i = 0; i < 4; ++i) mbar_init(&s.acc_ready[i], PAIR ? 1 : 2);   // one commit per issuer warp, layer and chunk (PAIR: one issuer)
    fence_mbar_init();
  }
  if (warp == 2) {
    if (PAIR) tmem_alloc_cg2<512>(&s.tmem_base);
    else tmem_alloc<512>(&s.tmem_base);
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = s.tmem_base;

  // tile walked in iteration `it`: PAIR -> the cluster takes two consecutive tiles (same image: tiles_per_img is even)
  const int grid_units = PAIR ? (int)gridDim.x / 2 : (int)gridDim.x;
  const int total_units = PAIR ? a.total_tiles / 2 : a.total_tiles;
  const int iters = (total_units + grid_units - 1) / grid_units;
  auto tile_of = [&](int it) {
    return PAIR ? (it * grid_units + (int)blockIdx.x / 2) * 2 + (int)crank : it * (int)gridDim.x + (int)blockIdx.x;
  };
  const int L = a.n_layers;
#ifdef C3D_CIPS_ABLATE
  const int abl = a.ablate;
#else
  constexpr int abl = 0;
#endif
  if (a.stagger_ns > 0) {        // whole CTA, before any pipeline role starts
    const unsigned long long t_end = c3d_globaltimer() + (unsigned long long)a.stagger_ns * (blockIdx.x / (CL > 1 ? CL : 1));
    while (c3d_globaltimer() < t_end) __nanosleep(100);
    __syncthreads();
  }

  if (warp < 4 && PAIR) {
    // ================================================================ CTA pair: ONE issuer, N = 256 MMAs.
    // What the measurements of round 2 say about the issue side (profiles/r02s..r02y): an issuer warp cannot turn a ring step
    // (wait for the fill, issue, commit) around in less than ~350 clk, whatever the step contains, and the tensor pipe needs
    // 64 clk per N = 128 MMA -- so with 4 MMAs a step (256 clk of work) the ISSUE LOOP sets the pace, not the tensor pipe
    // (first pair form, two issuers sharing one ring and observing every fill: 10.6 ms; one ring per issuer: 9.25 ms;
    // 4.1 ms of either is there with all the work removed).  An N = 256 MMA is 128 clk: 4 of them are 512 clk of work per step,
    // one issuer thread keeps the pipe full (microbenchmark: 513 clk per step sustained), the layer is 16 steps, and the A operand
    // is read from shared memory twice per layer instead of four times (the MMA operand reads + the weight stream + the epilogue's
    // stores are ~144 B/clk of shared-memory traffic at the tensor floor with N = 128, ~112 B/clk with N = 256; the port gives 128).
    // Roles: warp 0 streams this CTA's half of every big tile (both CTAs), warp 1 issues (leader) or relays "my half landed" to the
    // leader's full barrier (peer), warp 2 owns the TMEM allocation, warp 3 idles.
    reg_dec<56>();
    TRACE_DECL;
    uint32_t stage = 0, phase = 0;
    if (warp == 0) {
      // ---------------------------------------------------------- weight producer
      for (int it = 0; it < iters; ++it) {
        const int tile = tile_of(it);
        const int img = tile < a.total_tiles ? tile / a.tiles_per_img : 0;     // dummy tiles stream image 0
        for (int l = 0; l < L; ++l) {
          const uint8_t* src = reinterpret_cast<const uint8_t*>(a.wtiles) +
                               ((size_t)img * a.img_tile_stride + (size_t)a.layer_tile_off[l] + crank) * kWTileBytes;
          const int n_big = a.layer_kc[l] * 2;
#pragma unroll 1
          for (int i = 0; i < n_big; ++i) {
            wait_warp(&s.empty[stage], phase ^ 1, lane);
            if (lane == 0 && TRACE_TILE(i)) TRACE(it, 10, (uint32_t)(l << 8 | i));                  // producer: stage free, load issued
            if (abl & 4) {
              if (elect_one()) mbar_arrive(&s.full[stage]);
            } else if (elect_one()) {      // this CTA's half of big tile i: stream tile 2 i + rank
              mbar_arrive_expect_tx(&s.full[stage], kWTileBytes);
              bulk_g2s(s.w[stage], src + (size_t)(2 * i) * kWTileBytes, kWTileBytes, &s.full[stage]);
            }
            __syncwarp();
            if (++stage == NS) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else if (warp == 1 && !leader) {
      // ---------------------------------------------------------- peer CTA: relay "my half of this fill has landed" -- the second
      // arrival of the leader's full barrier (one remote arrive per fill, in ring order)
      for (int it = 0; it < iters; ++it)
        for (int l = 0; l < L; ++l) {
          const int n_big = a.layer_kc[l] * 2;
#pragma unroll 1
          for (int i = 0; i < n_big; ++i) {
            wait_warp(&s.full[stage], phase, lane);
            if (lane == 0 && TRACE_TILE(i)) TRACE(it, 11, (uint32_t)(l << 8 | i));
            if (elect_one()) mbar_arrive_cluster(&s.full[stage], 0);
            __syncwarp();
            if (++stage == NS) { stage = 0; phase ^= 1; }
          }
        }
    } else if (warp == 1) {
      // ---------------------------------------------------------- MMA issuer (leader CTA; whole warp converged, one lane issues):
      // the layer's steps are unrolled with their tile, dependency and commit mask as template arguments (pair_step above)
      PairIssuer<PAIR> c;
      c.s = &s;
      c.a_lo0 = umma_desc_lo(smem_u32(s.x), kLBO);
      c.b_lo0 = umma_desc_lo(smem_u32(s.w[0]), RC::kLBO_B);
      c.tmem = tmem;
      c.stage = 0; c.phase = 0;
      c.lane = lane; c.abl = abl;
#ifdef C3D_TRACE
      c.tr_n = 0;
#endif
      for (int it = 0; it < iters; ++it) {
        c.it = it;
#pragma unroll 1
        for (int l = 0; l < L; ++l) {
          c.l = l;
          c.epar = (uint32_t)(it * L + l) & 1u;   // phase of the epilogue that feeds layer l (staging for l = 0)
          if (a.layer_kc[l] * 2 == 16) pair_layer_full<PAIR>(c, std::make_index_sequence<16>{});
          else pair_layer_in<PAIR>(c);
        }
      }
    }
  } else if (warp < 4) {
    // ================================================================ single CTA (and multicast clusters): one ring, two issuers
    reg_dec<56>();
    if (warp == 0) {
      // ---------------------------------------------------------- weight producer (whole warp converged, one lane issues)
      TRACE_DECL;
      uint32_t stage = 0, phase = 0;
      for (int it = 0; it < iters; ++it) {
        const int tile = tile_of(it);
        const int img = tile < a.total_tiles ? tile / a.tiles_per_img : 0;     // dummy tiles stream image 0
        for (int l = 0; l < L; ++l) {
          const uint8_t* src = reinterpret_cast<const uint8_t*>(a.wtiles) +
                               ((size_t)img * a.img_tile_stride + (size_t)a.layer_tile_off[l]) * kWTileBytes;
          const int ntiles = a.layer_kc[l] * 4;
          for (int t = 0; t < ntiles; ++t) {
            wait_warp(&s.empty[stage], phase ^ 1, lane);
            if (lane == 0 && TRACE_TILE(t)) TRACE(it, 10, (uint32_t)(l << 8 | t));                  // producer: stage free, load issued
            if (abl & 4) {
              if (elect_one()) mbar_arrive(&s.full[stage]);
            } else if (elect_one()) {
              mbar_arrive_expect_tx(&s.full[stage], kWTileBytes);
              load_w_tile<CL>(s.w[stage], src + (size_t)t * kWTileBytes, &s.full[stage], crank);
            }
            __syncwarp();
            if (++stage == NS) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else if (warp == 1 || warp == 3) {
      // ---------------------------------------------------------- MMA issuers (whole warp converged, one lane issues).
      // Two issuer warps share the weight ring: issuer i owns accumulator column blocks nc = 2i, 2i+1 (disjoint
      // TMEM columns, so the two instruction streams never touch the same accumulator) -- one thread alone
      // cannot issue 4 MMAs + bookkeeping inside the 256 clk a tile occupies the tensor pipe.
      TRACE_DECL;
      const uint32_t me = warp == 1 ? 0u : 1u;
      const uint32_t idesc = umma_idesc_f16(kTileM, kNC);
      const uint32_t dhi = umma_desc_hi(kSBO);
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(s.x), kLBO);
      const uint32_t b_lo0 = umma_desc_lo(smem_u32(s.w[0]), RC::kLBO_B);
      constexpr uint32_t kStepK16 = (2 * kLBO) >> 4;       // one K=16 MMA step  = 2 core-matrix columns
      constexpr uint32_t kStepK16B = (2 * RC::kLBO_B) >> 4;
      constexpr uint32_t kStepStage = RC::kStageBytes >> 4;
      uint32_t stage = 0, phase = 0;
      for (int it = 0; it < iters; ++it) {
        for (int l = 0; l < L; ++l) {
          const uint32_t epar = (uint32_t)(it * L + l) & 1u;   // phase of the epilogue that feeds layer l (staging for l = 0)
          const int ntiles = a.layer_kc[l] * 4;
          const bool full_layer = ntiles == 32;
          const uint16_t* order = full_layer ? a.order_full : a.order_in;
          int waited = -1;
#pragma unroll 1
          for (int t = 0; t < ntiles; ++t) {
            const uint32_t e = order[t];
            const uint32_t kc = e & 15u, nc = (e >> 4) & 15u;
            const bool mine = (nc >> 1) == me;
            if (mine) {
              const int need = (int)((e >> 8) & 15u);
              if (lane == 0 && TRACE_TILE(t)) TRACE(it, 1 + me, (uint32_t)(l << 8 | t));          // tile reached
              if (need > waited && !(abl & 16)) {
                for (int j = waited + 1; j <= need; ++j) {
                  wait_warp(&s.epi_done[j], epar, lane);
                }
                waited = need;
              }
              if (lane == 0 && TRACE_TILE(t)) TRACE(it, 3 + me, (uint32_t)(l << 8 | t));          // epilogue dependency satisfied
            }
            // BOTH issuers observe every fill of every stage, in ring order, and a stage is released only when
            // both have (empty count 2): an issuer that skipped the phases of tiles it does not own could see
            // full[stage] one whole revolution stale (parity aliasing) whenever the epilogue outruns the refill.
#ifdef C3D_INJECT_RING_RACE
            if (mine)
#endif
            wait_warp(&s.full[stage], phase, lane);
            if (mine) {
              if (lane == 0 && TRACE_TILE(t)) TRACE(it, 5 + me, (uint32_t)(l << 8 | t));          // weight tile landed
              tc_fence_after();
              if (elect_one()) {
                const uint32_t a_lo = a_lo0 + kc * (kStepK16 * (kKC / 16));
                const uint32_t b_lo = b_lo0 + stage * kStepStage;
                const uint32_t d = tmem + nc * kNC;
                if ((abl & 10) == 10 && CL == 1) {      // no MMAs in flight: plain arrives stand in for the commits
                  mbar_arrive(&s.empty[stage]);
                  const uint32_t rdy8 = e >> 12;
                  for (int j = 0; j < 4; ++j)
                    if (rdy8 & (1u << j)) mbar_arrive(&s.acc_ready[j]);
                } else if (abl & 2) {
                  commit_stage_free<CL>(&s.empty[stage]);
                } else {
                  umma_ss_w(d, a_lo, b_lo, dhi, idesc, kc != 0);
                  umma_ss_w(d, a_lo + kStepK16, b_lo + kStepK16, dhi, idesc, 1);
                  umma_ss_w(d, a_lo + 2 * kStepK16, b_lo + 2 * kStepK16, dhi, idesc, 1);
                  umma_ss_w(d, a_lo + 3 * kStepK16, b_lo + 3 * kStepK16, dhi, idesc, 1);
                  commit_stage_free<CL>(&s.empty[stage]);
                }
                const uint32_t rdy = ((abl & 10) == 10 && CL == 1) ? 0u : e >> 12;    // chunks j for which this is the issuer's last tile of S_j
                if (rdy) {
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    if (rdy & (1u << j)) {
                      tc_commit(&s.acc_ready[j]);
                    }
                }
              }
            }
#ifndef C3D_INJECT_RING_RACE
            else if (elect_one()) {
              observe_stage_free<CL>(&s.empty[stage]);
            }
#endif
            __syncwarp();
            if (++stage == NS) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps
    TRACE_DECL;
    if (lane == 0) TRACE(1, 20, 0);      // trace builds: how long setmaxnreg.inc waits for the control warps' registers
    reg_inc<104>();
    if (lane == 0) TRACE(1, 21, 0);
    const int ew = warp - 4;
    const int wg = ew >> 2;            // column group 0..3
    const int q = warp & 3;            // TMEM lane quarter
    const int row = q * 32 + lane;     // row of the tile this thread owns
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    ResVec* resid = reinterpret_cast<ResVec*>(a.resid) + (size_t)blockIdx.x * (kH / kResCols) * kTileM;
    for (int it = 0; it < ((abl & 16) ? 0 : iters); ++it) {
      const int tile = tile_of(it);
      const bool tile_ok = tile < a.total_tiles;
      const int img = tile_ok ? tile / a.tiles_per_img : 0;
      const int pix = tile_ok ? (tile % a.tiles_per_img) * kTileM + row : a.N;
      const bool row_ok = tile_ok && pix < a.N;
      float rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f;
      // ---- e = 0: stage the input tile (K padded to 64) as layer 0's A operand
      {
        if (wg < 2) {   // 64 columns: wg 0 -> k 0..31, wg 1 -> k 32..63
          const int k0 = wg * 32;
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int k = k0 + 2 * j;
            float v0 = 0.f, v1 = 0.f;
            if (row_ok && k < a.in_dim) {
              const float* xp = a.x + ((size_t)img * a.N + pix) * a.in_dim;
              v0 = xp[k];
              if (k + 1 < a.in_dim) v1 = xp[k + 1];
            }
            pk[j] = pack_f16(v0, v1);
          }
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4*>(s.x + (size_t)(k0 / 8 + g) * kLBO + row * 16) =
                make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
        }
        fence_proxy_async();      // the A operand in THIS CTA's shared memory is what the tensor core (async proxy) reads; PAIR too
        tc_fence_before();
        __syncwarp();
        if (lane == 0)
          for (int j = 0; j < 4; ++j) {
            if (PAIR) mbar_arrive_cluster(&s.epi_done[j], 0);      // the leader's barrier collects both CTAs
            else mbar_arrive(&s.epi_done[j]);
          }
      }
      // ---- e = l + 1: epilogue of layer l
      for (int l = 0; l < L; ++l) {
        const int blk = l >> 1;
        const bool second = (l & 1) != 0;
        EpiFlags f;
        f.last = l == L - 1;
        f.add_res = second && blk >= a.skip_from && blk >= 1;      // block input dim == 512 for blk >= 1
        f.keep_res = second && (blk + 1 >= a.skip_from) && !f.last;  // the next block adds this output
        f.do_rgb = second && blk >= a.rgb_from;
        if (abl & 64) f.no_x = true;
        if (abl & 128) f.add_res = false;
        if (abl & 512) f.keep_res = false;
        if (abl & 256) f.do_rgb = false;
        if (f.do_rgb) {   // ToRGB weights of this block -> shared memory (overlaps the MMAs)
          s.rgbw[(int)threadIdx.x - 128] = __ldg(a.rgbw + (size_t)blk * kH + ((int)threadIdx.x - 128));
          named_bar_sync_c<1, kNumEpiWarps * 32>();
        }
        float* hid = (f.last && a.hidden_out && row_ok) ? a.hidden_out + ((size_t)img * a.N + pix) * kH : nullptr;
        // the 8 x 16-column slices this thread owns (chunk j, halves 0/1); every address advances by a constant per chunk
        uint32_t accA[16] = {}, accB[16] = {};
        ResVec rsA[kResVecs], rsB[kResVecs];
        const int cw = wg * 32;
        uint32_t tcol = trow + (uint32_t)cw;                    // TMEM column of slice (j, 0)
        ResVec* rp = resid + (size_t)(cw / kResCols) * kTileM + row;   // residual slot of slice (j, 0), g = 0
        uint8_t* xp = s.x + (size_t)(cw / 8) * kLBO + row * 16;
        const float4* rwp = s.rgbw + cw;
        float* hp = hid ? hid + cw : nullptr;
        // DUMP: this row's 16-byte units of layer l's output in the activation stash (two per 16-column slice)
        uint4* dp = (DUMP && a.acts && row_ok)
                        ? reinterpret_cast<uint4*>(a.acts + (size_t)l * a.acts_layer_stride + ((size_t)img * a.N + pix) * kH + cw)
                        : nullptr;
        uint16_t* zp = (DUMP && a.zsign && row_ok) ? a.zsign + ((size_t)l * a.acts_layer_stride + ((size_t)img * a.N + pix) * kH + cw) / 16
                                                   : nullptr;
        auto load_res = [&](ResVec (&rs)[kResVecs], const ResVec* p) {
#pragma unroll
          for (int g = 0; g < kResVecs; ++g) rs[g] = p[g * kTileM];
        };
        constexpr int kResSlice = kResVecs * kTileM, kResChunk = (kNC / kResCols) * kTileM;   // pointer steps: 16 / 128 columns
        // chunk j starts as soon as accumulator block j is complete and no MMA of this layer still reads
        // A-operand chunk j (acc_ready[j]); the rest of the layer's MMAs run underneath.
        const uint32_t apar = (uint32_t)(it * L + l) & 1u;
        bool have = false;                                      // the chunk's accumulators are already on their way to registers
        if (f.add_res) { load_res(rsA, rp); load_res(rsB, rp + kResSlice); }
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
          if (abl & 1) {
            wait_epilogue(&s.acc_ready[j], apar, warp, lane);
            tc_fence_after();
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0 && !f.last) {
              if (PAIR) mbar_arrive_cluster(&s.epi_done[j], 0);
              else mbar_arrive(&s.epi_done[j]);
            }
            continue;
          }
          if (!have) {
            wait_epilogue(&s.acc_ready[j], apar, warp, lane);
            tc_fence_after();
            if (!(abl & 32)) { tmem_ld16(tcol, accA); tmem_ld16(tcol + 16, accB); }
          }
          if (lane == 0 && TRACE_EPI_WARP(warp)) TRACE(it, 8, (uint32_t)(l << 8 | j));                 // accumulator block j complete
          tc_wait_ld();
          // ---- phase 1: the next layer's A operand, then hand the chunk over
          if (second) { epi16_act<true>(accA, rsA, xp, f); epi16_act<true>(accB, rsB, xp + 2 * kLBO, f); }
          else { epi16_act<false>(accA, rsA, xp, f); epi16_act<false>(accB, rsB, xp + 2 * kLBO, f); }
          fence_proxy_async();      // CTA scope: the A operand in THIS CTA's shared memory is what the tensor core (async proxy) reads
          tc_fence_before();
          __syncwarp();
          if (lane == 0 && !f.last) {
            if (PAIR) mbar_arrive_cluster(&s.epi_done[j], 0);
            else mbar_arrive(&s.epi_done[j]);
          }
          if (lane == 0 && TRACE_EPI_WARP(warp)) TRACE(it, 9, (uint32_t)(l << 8 | (warp - 4) << 2 | j));   // warp handed chunk j over
          // ---- phase 2: residual store, ToRGB, outputs -- under the MMAs the hand-over released
          if (second) {
            epi16_tail<true, DUMP>(accA, rsA, rwp, rp, f, rgb0, rgb1, rgb2, hp, dp, zp);
            epi16_tail<true, DUMP>(accB, rsB, rwp + 16, rp + kResSlice, f, rgb0, rgb1, rgb2, hp ? hp + 16 : nullptr, dp ? dp + 2 : nullptr, zp ? zp + 1 : nullptr);
          } else if (DUMP) {
            epi16_tail<false, DUMP>(accA, rsA, rwp, rp, f, rgb0, rgb1, rgb2, hp, dp);
            epi16_tail<false, DUMP>(accB, rsB, rwp + 16, rp + kResSlice, f, rgb0, rgb1, rgb2, hp ? hp + 16 : nullptr, dp ? dp + 2 : nullptr);
          }
          have = false;
          if (j < 3) {      // the next chunk: its residual a whole chunk ahead, its accumulators if they are complete already
            if (f.add_res) { load_res(rsA, rp + kResChunk); load_res(rsB, rp + kResChunk + kResSlice); }
            have = test_warp(&s.acc_ready[j + 1], apar, lane);
            if (have) {
              tc_fence_after();
              if (!(abl & 32)) { tmem_ld16(tcol + 128, accA); tmem_ld16(tcol + 144, accB); }
            }
          }
          tcol += 128; rp += kResChunk; xp += 16 * kLBO; rwp += 128;
          if (hp) hp += 128;
          if (dp) dp += 16;
          if (zp) zp += 8;
        }
      }
      named_bar_sync_c<1, kNumEpiWarps * 32>();   // rgbw (aliased) no longer read
      s.rgb_part[wg][row][0] = rgb0;
      s.rgb_part[wg][row][1] = rgb1;
      s.rgb_part[wg][row][2] = rgb2;
      named_bar_sync_c<1, kNumEpiWarps * 32>();
      if (wg == 0 && row_ok) {
        float o[3];
#pragma unroll
        for (int c = 0; c < 3; ++c)
          o[c] = tanhf(((s.rgb_part[0][row][c] + s.rgb_part[1][row][c]) + (s.rgb_part[2][row][c] + s.rgb_part[3][row][c])) + a.rgbb[c]);
        float* op = a.rgb + ((size_t)img * a.N + pix) * 3;
        op[0] = o[0]; op[1] = o[1]; op[2] = o[2];
      }
      named_bar_sync_c<1, kNumEpiWarps * 32>();
    }
  }
  // ------------------------------------------------------------ teardown
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();
  if (warp == 2) {
    if (PAIR) tmem_dealloc_cg2<512>(tmem);
    else tmem_dealloc<512>(tmem);
  }
}

// fp32 (in,out) weights -> fp16 UMMA-B tiles, one launch for all layers.  Block = one 16 KB tile in STREAM
// order: tile t of layer l is (kc,nc) = order[t]; it holds B[n][k] = W[kc*64+k][nc*128+n] at byte
// (n%8)*16 + (n/8)*128 + (k/8)*2048 + (k%8)*2.  Rows k >= in_dim are zero (padded input layer).
struct PrepArgs {
  const float* w[kMaxLayers];
  const float* s1p[kMaxLayers];     // (B,in_l)
  const float* demod[kMaxLayers];   // (B,512)
  int layer_tile_off[kMaxLayers + 1];
  int in_dim0, n_layers, B;
  uint16_t order_full[32];
  uint16_t order_in[4];
};
// grid = (tiles per image, B): block = one 16 KB tile of one image, in STREAM order.  (The PAIR kernel streams the same tiles in ITS
// order -- see the host's pair_stream_order: the two halves of a big tile are two ordinary tiles, adjacent.)
__global__ void cips_prep_weights_kernel(const PrepArgs pa, __half* __restrict__ out) {
  const int gt = blockIdx.x, b = blockIdx.y;
  int l = 0;
  while (l + 1 < pa.n_layers && gt >= pa.layer_tile_off[l + 1]) ++l;
  const int t = gt - pa.layer_tile_off[l];
  const uint32_t e = (pa.layer_tile_off[l + 1] - pa.layer_tile_off[l]) == 32 ? pa.order_full[t] : pa.order_in[t];
  const int kc = e & 15, nc = (e >> 4) & 15;
  const int in_dim = l == 0 ? pa.in_dim0 : kH;
  const float* W = pa.w[l];
  const float* sv = pa.s1p[l] + (size_t)b * in_dim;
  const float* dv = pa.demod[l] + (size_t)b * kH + nc * kNC;
  __half* o = out + ((size_t)b * pa.layer_tile_off[pa.n_layers] + gt) * (kWTileBytes / 2);
  for (int i = threadIdx.x; i < kKC * kNC; i += blockDim.x) {
    const int k = i / kNC, n = i % kNC;    // n fastest -> coalesced reads of W rows
    const int gk = kc * kKC + k;
    const float v = gk < in_dim ? (sv[gk] * W[(size_t)gk * kH + nc * kNC + n]) * dv[n] : 0.f;
    o[((n % 8) * 16 + (n / 8) * 128 + (k / 8) * kLBO) / 2 + (k % 8)] = __float2half_rn(v);
  }
}

// packed ToRGB weights + summed biases
__global__ void cips_prep_consts_kernel(C3dCipsWeights w, int n_blocks, int rgb_from, float4* rgbw, float* rgbb) {
  const int blk = blockIdx.x;
  if (blk < n_blocks && blk >= rgb_from)
    for (int n = threadIdx.x; n < kH; n += blockDim.x)
      rgbw[(size_t)blk * kH + n] = make_float4(w.rgb_w[blk][n], w.rgb_w[blk][kH + n], w.rgb_w[blk][2 * kH + n], 0.f);
  if (blk == 0 && threadIdx.x < 3) {
    float sacc = 0.f;
    for (int b = rgb_from; b < n_blocks; ++b) sacc += w.rgb_b[b][threadIdx.x];   // same order as the reference's skip chain
    rgbb[threadIdx.x] = sacc;
  }
}

}  // namespace cips
}  // namespace c3d

using namespace c3d;
using namespace c3d::cips;

static int cips_grid(const C3dCipsParams* p, int* cl_out, bool* pair_out) {
  int dev = 0;
  cudaGetDevice(&dev);
  const int sms = c3d_device_sm_count(dev);
  int cl = c3d_options().cips_cluster;
  const int tiles_per_img = (p->n_pix + kTileM - 1) / kTileM;
  // tcgen05 CTA pairs (cta_group::2) wherever an image has an even number of tiles: 8.4 ms against the single-CTA kernel's 10.9 ms at
  // B = 16, r256 (profiles/r02ac_cips_epilogue.txt).  C3D_CIPS_PAIR=0 selects the single-CTA kernel.
  bool pair = c3d_options().cips_pair != 0 && cl == 1;      // an explicit C3D_CIPS_CLUSTER=2|4 asks for the multicast (non-pair) form
  if (pair && (tiles_per_img % 2 || sms < 2)) pair = false;     // a pair works on two tiles of ONE image
  if (pair) cl = 2;
  *pair_out = pair;
  // the CTAs of a cluster share one multicast weight stream: they must all work on the same image at the same
  // time, which holds iff clusters never straddle an image boundary (found by the CPU emulation, tools/emu)
  if (tiles_per_img % cl || sms < cl) cl = 1;
  long long total = (long long)p->batch * tiles_per_img;
  int grid = (int)(total < sms ? total : sms);
  grid = (grid + cl - 1) / cl * cl;
  if (grid > sms) grid = sms / cl * cl;
  *cl_out = cl;
  return grid;
}

struct CipsWs {
  size_t wtiles, rgbw, rgbb, resid, total;
};
static CipsWs cips_ws_layout(const C3dCipsParams* p) {
  CipsWs o;
  const int L = 2 * p->n_blocks;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t r = off; off += (bytes + 255) / 256 * 256; return r; };
  o.wtiles = take((size_t)(p->batch > 0 ? p->batch : 1) * (4 + (L - 1) * 32) * kWTileBytes);   // per image
  o.rgbw = take((size_t)p->n_blocks * kH * 16);
  o.rgbb = take(16);
  o.resid = take((size_t)160 * (kH / 4) * kTileM * 16);
  o.total = off;
  return o;
}

size_t c3d_cips_tc_workspace_bytes(const C3dCipsParams* p) { return cips_ws_layout(p).total; }

#ifdef C3D_TRACE
// out: records of 2 words: [block << 8 | warp, stamp]; returns the number of records
extern "C" int c3d_debug_cips_trace(unsigned long long* out, int cap) {
  static unsigned long long host[2 * 20 * c3d::cips::kTraceCap];
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(host, c3d::cips::g_trace, sizeof(host));
  int n = 0;
  for (int w = 0; w < 40; ++w)
    for (int i = 0; i < c3d::cips::kTraceCap && 2 * n + 1 < cap; ++i)
      if (host[(size_t)w * c3d::cips::kTraceCap + i]) {
        out[2 * n] = (unsigned long long)((w / 20) << 8 | (w % 20));
        out[2 * n + 1] = host[(size_t)w * c3d::cips::kTraceCap + i];
        ++n;
      }
  memset(host, 0, sizeof(host));
  cudaMemcpyToSymbol(c3d::cips::g_trace, host, sizeof(host));
  return n;
}
#endif

template <int CL, bool PAIR = false, bool DUMP = false, bool R16 = false>
static int launch_cips(const KArgs& ka, int grid, cudaStream_t st) {
  const size_t smem = sizeof(SmemT<PAIR>) + 1024;
  static std::atomic<unsigned long long> attr_set{0};     // per device, once
  int dev = 0;
  cudaGetDevice(&dev);
  auto kern = cips_tc_kernel<CL, PAIR, DUMP, R16>;
  if (!(attr_set.load() >> (dev & 63) & 1ull)) {
    C3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set.fetch_or(1ull << (dev & 63));
  }
#ifdef C3D_EMU
  c3d_count_launch();
  C3D_CUDA(C3D_LAUNCH_CLUSTER(kern, grid, kThreads, smem, st, CL, ka));
  return C3D_OK;
#else
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  c3d_count_launch();
  C3D_CUDA(cudaLaunchKernelEx(&cfg, kern, ka));
  return C3D_OK;
#endif
}

// diagnostic: how many clusters of `cl` CTAs of this kernel can be resident at once on the current device (a persistent
// kernel whose grid exceeds that runs in WAVES).  pair != 0: the cta_group::2 instantiation.
extern "C" int c3d_debug_cips_max_clusters(int cl, int pair) {
#ifdef C3D_EMU
  return 1 << 20;
#else
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148 / cl * cl);
  cfg.blockDim = dim3(kThreads);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = -1;
  cudaError_t e;
  if (pair) {
    cfg.dynamicSmemBytes = sizeof(SmemT<true>) + 1024;
    cudaFuncSetAttribute(cips_tc_kernel<2, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes);
    e = cudaOccupancyMaxActiveClusters(&n, cips_tc_kernel<2, true, false>, &cfg);
  } else if (cl == 2) {
    cfg.dynamicSmemBytes = sizeof(SmemT<false>) + 1024;
    cudaFuncSetAttribute(cips_tc_kernel<2, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes);
    e = cudaOccupancyMaxActiveClusters(&n, cips_tc_kernel<2, false, false>, &cfg);
  } else {
    cfg.dynamicSmemBytes = sizeof(SmemT<false>) + 1024;
    cudaFuncSetAttribute(cips_tc_kernel<1, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes);
    e = cudaOccupancyMaxActiveClusters(&n, cips_tc_kernel<1, false, false>, &cfg);
  }
  if (e != cudaSuccess) { c3d_set_error("cudaOccupancyMaxActiveClusters: %s", cudaGetErrorString(e)); return -1; }
  return n;
#endif
}

// Issue order of the weight tiles of one layer (see KArgs::order_full).  Host-only; also exported for the
// CPU protocol test (c3d_debug_cips_tile_order).
static void build_tile_order(uint16_t* order_full, uint16_t* order_in) {
  int n = 0;
  // head: group j (everything epilogue chunk j of the previous layer unlocks) = new K-chunks 2j, 2j+1 for the
  // accumulator blocks <= j, plus the new accumulator block j for the K-chunks seen so far.
  // tail (everything chunk 3 unlocks) is ordered so that accumulator block j completes -- and A-operand chunk j
  // (kc = 2j, 2j+1) is read for the last time -- as early as possible:
  //   (6,0)(7,0)(0,3)(1,3) | (6,1)(7,1)(2,3)(3,3) | (6,2)(7,2)(4,3)(5,3) | (6,3)(7,3)
  auto put = [&](int kc, int nc, int need) { order_full[n++] = (uint16_t)(kc | (nc << 4) | (need << 8)); };
  for (int j = 0; j < 3; ++j)
    for (int kc = 0; kc < 2 * (j + 1); ++kc)
      for (int nc = 0; nc <= j; ++nc)
        if (kc >= 2 * j || nc == j) put(kc, nc, j);
  for (int j = 0; j < 4; ++j) {
    put(6, j, 3);
    put(7, j, 3);
    if (j < 3) { put(2 * j, 3, 3); put(2 * j + 1, 3, 3); }
  }
  // commit mask (bits 12..15): tile is its owner's last one in S_j = {nc == j or kc/2 == j}; an issuer without a
  // tile in S_j commits acc_ready[j] with its first tile
  auto mark = [&](uint16_t* ord, int cnt) {
    for (int me = 0; me < 2; ++me)
      for (int j = 0; j < 4; ++j) {
        int at = -1, first = -1;
        for (int i = 0; i < cnt; ++i) {
          const int kc = ord[i] & 15, nc = (ord[i] >> 4) & 15;
          if ((nc >> 1) != me) continue;
          if (first < 0) first = i;
          if (nc == j || (kc >> 1) == j) at = i;
        }
        if (at < 0) at = first;
        ord[at] |= (uint16_t)(1u << (12 + j));
      }
  };
  for (int nc = 0; nc < 4; ++nc) order_in[nc] = (uint16_t)(0 | (nc << 4) | (nc << 8));   // padded input layer: K = 64
  mark(order_full, 32);
  mark(order_in, 4);
}

// PAIR kernel (N = 256 MMAs, one issuer): issue order of the big tiles (kc, h), h = accumulator columns 256 h .. 256 h + 255.
// Big tile (kc, h) overwrites / accumulates into accumulator blocks 2h and 2h+1 and reads A-operand chunk kc / 2, so it needs the
// previous layer's epilogue chunks up to max(kc / 2, 2h + 1).  Head: (0..3, 0) after chunk 1, (4, 0)(5, 0) after chunk 2; tail, after
// chunk 3: (6, 0)(7, 0) then (0..7, 1).  acc_ready[j] (block j complete, A chunk j no longer read) is committed with the last tile of
// S_j = {h == j / 2 or kc / 2 == j}: (1, 1), (3, 1), (7, 1), (7, 1).  stream_*: the same order as (kc, nc) tiles, the leader's half
// (nc = 2h) then the peer's (nc = 2h + 1) -- what the prep kernel writes and the two producers read.
static void build_pair_order(uint16_t* big_full, uint16_t* big_in, uint16_t* stream_full, uint16_t* stream_in) {
  int n = 0;
  auto put = [&](int kc, int h) {
    const int need = kc / 2 > 2 * h + 1 ? kc / 2 : 2 * h + 1;
    big_full[n++] = (uint16_t)(kc | (h << 4) | (need << 8));
  };
  for (int kc = 0; kc < 8; ++kc) put(kc, 0);
  for (int kc = 0; kc < 8; ++kc) put(kc, 1);
  big_in[0] = (uint16_t)(0 | (0 << 4) | (1 << 8));
  big_in[1] = (uint16_t)(0 | (1 << 4) | (3 << 8));
  auto mark = [&](uint16_t* ord, int cnt) {
    for (int j = 0; j < 4; ++j) {
      int at = -1;
      for (int i = 0; i < cnt; ++i) {
        const int kc = ord[i] & 15, h = (ord[i] >> 4) & 15;
        if (h == j / 2 || (kc >> 1) == j) at = i;
      }
      if (at < 0) at = 0;
      ord[at] |= (uint16_t)(1u << (12 + j));
    }
  };
  mark(big_full, 16);
  mark(big_in, 2);
  for (int i = 0; i < 16; ++i)
    for (int r = 0; r < 2; ++r) stream_full[2 * i + r] = (uint16_t)((big_full[i] & 15) | ((2 * ((big_full[i] >> 4) & 15) + r) << 4));
  for (int i = 0; i < 2; ++i)
    for (int r = 0; r < 2; ++r) stream_in[2 * i + r] = (uint16_t)((big_in[i] & 15) | ((2 * ((big_in[i] >> 4) & 15) + r) << 4));
}

extern "C" int c3d_debug_cips_tile_order(uint16_t* order_full32, uint16_t* order_in4) {
  if (!order_full32 || !order_in4) return C3D_EINVAL;
  build_tile_order(order_full32, order_in4);
  return kStages;
}

int c3d_cips_fwd_tc(const C3dCipsParams* p, const C3dCipsWeights* w, const float* x, float* rgb, float* hidden_out,
                    void* workspace, size_t workspace_bytes, cudaStream_t st, void* acts_f16, void* zsign_u16) {
  C3D_CHECK_ARG(p->hidden == kH, "cips(tc): hidden must be 512, got %d", p->hidden);
  C3D_CHECK_ARG(p->in_dim >= 1 && p->in_dim <= kKC, "cips(tc): in_dim must be <= 64, got %d", p->in_dim);
  C3D_CHECK_ARG(p->skip_from >= 1, "cips(tc): skip_from must be >= 1 (block 0 changes width)");
  const CipsWs ws = cips_ws_layout(p);
  if (workspace_bytes < ws.total) {
    c3d_set_error("cips(tc): workspace too small (%zu < %zu)", workspace_bytes, ws.total);
    return C3D_EWORKSPACE;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  if (!c3d_device_supported(dev)) {
    c3d_set_error("cips(tc): device %d is not sm_100 (tcgen05 required)", dev);
    return C3D_EARCH;
  }
  uint8_t* base = (uint8_t*)workspace;
  const int L = 2 * p->n_blocks;
  KArgs ka = {};
  ka.x = x; ka.rgb = rgb; ka.hidden_out = hidden_out;
  ka.acts = (__half*)acts_f16;
  ka.zsign = (uint16_t*)zsign_u16;
  ka.stagger_ns = c3d_options().cips_stagger_ns;
  ka.ablate = 0;
#ifdef C3D_CIPS_ABLATE
  if (const char* e = getenv("C3D_CIPS_ABLATE")) ka.ablate = atoi(e);
#endif
  ka.acts_layer_stride = (size_t)p->batch * p->n_pix * kH;
  ka.wtiles = (const __half*)(base + ws.wtiles);
  ka.rgbw = (const float4*)(base + ws.rgbw);
  ka.rgbb = (const float*)(base + ws.rgbb);
  ka.resid = (float4*)(base + ws.resid);
  ka.B = p->batch; ka.N = p->n_pix; ka.in_dim = p->in_dim; ka.n_layers = L;
  ka.skip_from = p->skip_from; ka.rgb_from = p->rgb_from;
  ka.tiles_per_img = (p->n_pix + kTileM - 1) / kTileM;
  ka.total_tiles = p->batch * ka.tiles_per_img;
  int off = 0;
  for (int l = 0; l < L; ++l) {
    ka.layer_tile_off[l] = off;
    ka.layer_kc[l] = l == 0 ? 1 : kH / kKC;
    off += ka.layer_kc[l] * 4;
  }
  ka.layer_tile_off[L] = off;
  ka.img_tile_stride = (size_t)off;
  build_tile_order(ka.order_full, ka.order_in);
  uint16_t big_full[16], big_in[2], pair_full[32], pair_in[4];      // PAIR kernel: big-tile order; the same as (kc, nc) tiles for the prep kernel
  build_pair_order(big_full, big_in, pair_full, pair_in);
  for (int i = 0; i < 16; ++i)      // the kernel's compile-time order (kPairFull / kPairIn) must be the order the weights are prepared in
    if (big_full[i] != kPairFull[i] || (i < 2 && big_in[i] != kPairIn[i])) {
      c3d_set_error("cips(tc): pair tile order mismatch at %d (host 0x%x, kernel 0x%x)", i, big_full[i], kPairFull[i]);
      return C3D_EINVAL;
    }
  int cl = 1;
  bool pair = false;
  const int grid = cips_grid(p, &cl, &pair);
  if (acts_f16) {       // the training forward is the single-CTA kernel: its weight tiles must be laid out for it
    pair = false;
    cl = 1;
  }
  // ---- prep: weights -> fp16 tiles (one launch), per-image epilogue vectors
  {
    PrepArgs pa = {};
    for (int l = 0; l < L; ++l) {
      pa.w[l] = w->w[l];
      pa.s1p[l] = w->style1p[l];
      pa.demod[l] = w->demod[l];
    }
    for (int l = 0; l <= L; ++l) pa.layer_tile_off[l] = ka.layer_tile_off[l];
    pa.in_dim0 = p->in_dim;
    pa.n_layers = L;
    pa.B = p->batch;
    for (int i = 0; i < 32; ++i) pa.order_full[i] = pair ? pair_full[i] : ka.order_full[i];
    for (int i = 0; i < 4; ++i) pa.order_in[i] = pair ? pair_in[i] : ka.order_in[i];
    C3D_LAUNCH(cips_prep_weights_kernel, dim3(ka.layer_tile_off[L], p->batch), 256, 0, st, pa, (__half*)(base + ws.wtiles));
    C3D_LAUNCH_CHECK();
  }
  C3D_LAUNCH(cips_prep_consts_kernel, p->n_blocks, 256, 0, st, *w, p->n_blocks, p->rgb_from, (float4*)(base + ws.rgbw),
             (float*)(base + ws.rgbb));
  C3D_LAUNCH_CHECK();
  if (acts_f16) {       // training forward: single-CTA kernel with the activation stash
    int g1 = ka.total_tiles < c3d_device_sm_count(dev) ? ka.total_tiles : c3d_device_sm_count(dev);
    return launch_cips<1, false, true>(ka, g1 < 1 ? 1 : g1, st);
  }
  // fp16 residual stream (see ResT): only when the caller takes the image alone
  const bool r16 = c3d_options().cips_res16 != 0 && !hidden_out;
  if (pair) return r16 ? launch_cips<2, true, false, true>(ka, grid, st) : launch_cips<2, true>(ka, grid, st);
  if (cl == 1) return r16 ? launch_cips<1, false, false, true>(ka, grid, st) : launch_cips<1>(ka, grid, st);
  if (cl == 2) return launch_cips<2>(ka, grid, st);
  return launch_cips<4>(ka, grid, st);
}
