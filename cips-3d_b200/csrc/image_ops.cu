// Image export (HBM-bound byte work): generator output (B, C, H, W) fp32 in [-1, 1] -> (B, H, W, C) uint8, the layout
// PIL / numpy take, in one pass on the device -- so the device->host copy of an inference batch moves 1 byte per
// sample instead of 4 and no fp32 intermediate (normalised copy, permuted copy) is ever written.
// Three conversions exist in the reference's inference paths; each is reproduced bit for bit (same fp32 operations
// in the same order, round-to-nearest, no FMA contraction; truncating float -> uint8 cast):
//   mode 0  torchvision save_image(img, path, normalize=True, value_range=(lo, hi))
//           exp/cips3d/scripts/gen_images.py:64, sample_images.py:73  (torchvision/utils.py make_grid.norm_ip, save_image)
//             y = (clamp(x, lo, hi) - lo) / max(hi - lo, 1e-5);  u8 = trunc(clamp(y * 255 + 0.5, 0, 255))
//   mode 1  tensor_to_PIL, exp/cips3d/models/st_web.py:44-46
//             y = x * 0.5 + 0.5;                                  u8 = trunc(clamp(y * 255 + 0.5, 0, 255))
//   mode 2  comm_utils.to_pil, exp/comm/comm_utils.py:21-24 (torchvision to_pil_image: (pic * 255).astype(uint8))
//             y = (x + 1) * 0.5;                                  u8 = trunc(y * 255)   (saturated; see header)
// Algorithmic bytes per pixel: 4*C read + C written (15 for RGB).
#include "c3d_common.cuh"

namespace c3d {

struct U8Args {
  const float* x;
  uint8_t* y;
  long long hw;        // pixels per plane
  int batch;
  float lo, hi, range; // range = max(hi - lo, 1e-5) (mode 0); its reciprocal when it is a power of two (kSavePow2)
};

// conversions: the three modes of the C-ABI, plus mode 0 specialised for a power-of-two range (the reference's (-1, 1)):
// x / 2^k and x * 2^-k are the same correctly rounded value, so the IEEE division sequence (~12 instructions and a
// slow-path call per element -- it would make this stream issue-bound) becomes one FMUL
enum { kSave = 0, kTensorToPil = 1, kToPil = 2, kSavePow2 = 3 };

template <int MODE>
__device__ __forceinline__ uint32_t to_u8(float x, float lo, float hi, float range) {
  float v;
  if (MODE == kSave || MODE == kSavePow2) {
    const float d = __fsub_rn(fminf(fmaxf(x, lo), hi), lo);
    v = __fadd_rn(__fmul_rn(MODE == kSave ? __fdiv_rn(d, range) : __fmul_rn(d, range), 255.f), 0.5f);
  } else if (MODE == kTensorToPil) {
    v = __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(x, 0.5f), 0.5f), 255.f), 0.5f);
  } else {
    v = __fmul_rn(__fmul_rn(__fadd_rn(x, 1.f), 0.5f), 255.f);
  }
  // clamp_(0, 255), then the truncating cast as a round-toward-zero add of 2^23: the integer lands in the low mantissa
  // byte (a full-rate FADD.RZ instead of a quarter-rate F2I); NaN -> 0 (fmaxf returns the non-NaN operand).
  // Only byte 0 of the result is meaningful.
  return __float_as_uint(__fadd_rz(fminf(fmaxf(v, 0.f), 255.f), 8388608.f));
}
__device__ __forceinline__ uint32_t pack4(uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3) {   // byte 0 of each
  return __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410);
}

// C = 3, hw % 4 == 0, 16 B-aligned planes: a thread converts 4 consecutive pixels -- three float4 loads (one per
// plane, each warp-wide request 512 contiguous bytes) and 12 output bytes as three 32-bit stores (a warp writes 384
// contiguous bytes).
template <int MODE>
__global__ void __launch_bounds__(256) image_u8_rgb4_kernel(const U8Args a) {
  const long long quads = a.hw >> 2, total = quads * a.batch;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long b = i / quads, q = i - b * quads;
    const float4* src = reinterpret_cast<const float4*>(a.x + b * 3 * a.hw) + q;
    const float4 r = __ldcs(src), g = __ldcs(src + quads), bl = __ldcs(src + 2 * quads);
    const float rr[4] = {r.x, r.y, r.z, r.w}, gg[4] = {g.x, g.y, g.z, g.w}, bb[4] = {bl.x, bl.y, bl.z, bl.w};
    uint32_t by[12];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      by[3 * p] = to_u8<MODE>(rr[p], a.lo, a.hi, a.range);
      by[3 * p + 1] = to_u8<MODE>(gg[p], a.lo, a.hi, a.range);
      by[3 * p + 2] = to_u8<MODE>(bb[p], a.lo, a.hi, a.range);
    }
    uint32_t* dst = reinterpret_cast<uint32_t*>(a.y + (b * a.hw + 4 * q) * 3);
#pragma unroll
    for (int w = 0; w < 3; ++w) __stcs(dst + w, pack4(by[4 * w], by[4 * w + 1], by[4 * w + 2], by[4 * w + 3]));
  }
}

// channels-last input (the generator's own output layout: CIPS writes (B, H*W, 3)): source and destination have the
// same element order, so the conversion is flat -- one float4 in, four bytes out
template <int MODE>
__global__ void __launch_bounds__(256) image_u8_flat4_kernel(const U8Args a, long long n4) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(a.x) + i);
    __stcs(reinterpret_cast<uint32_t*>(a.y) + i,
           pack4(to_u8<MODE>(v.x, a.lo, a.hi, a.range), to_u8<MODE>(v.y, a.lo, a.hi, a.range),
                 to_u8<MODE>(v.z, a.lo, a.hi, a.range), to_u8<MODE>(v.w, a.lo, a.hi, a.range)));
  }
}

// any C <= 4, any size / alignment: one thread per pixel (channels-first) or per element (flat)
template <int MODE>
__global__ void __launch_bounds__(256) image_u8_generic_kernel(const U8Args a, int channels, int flat) {
  const long long total = a.hw * a.batch * (flat ? channels : 1);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    if (flat) {
      a.y[i] = (uint8_t)to_u8<MODE>(__ldcs(a.x + i), a.lo, a.hi, a.range);
      continue;
    }
    const long long b = i / a.hw, px = i - b * a.hw;
    for (int c = 0; c < channels; ++c)
      a.y[i * channels + c] = (uint8_t)to_u8<MODE>(__ldcs(a.x + (b * channels + c) * a.hw + px), a.lo, a.hi, a.range);
  }
}

template <int MODE>
static int launch_u8(const U8Args& a, int channels, int channels_last, int sms, cudaStream_t st) {
  const bool aligned = ((uintptr_t)a.x & 15u) == 0 && ((uintptr_t)a.y & 3u) == 0;
  const long long n = a.hw * a.batch * channels;
  const bool flat4 = channels_last && n % 4 == 0 && aligned;
  const bool rgb4 = !channels_last && channels == 3 && a.hw % 4 == 0 && aligned;
  const long long items = flat4 ? n >> 2 : (rgb4 ? (a.hw >> 2) * a.batch : (channels_last ? n : a.hw * a.batch));
  long long blocks = (items + 255) / 256;
  if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;     // grid-stride over a whole number of waves
  if (flat4) C3D_LAUNCH(image_u8_flat4_kernel<MODE>, (int)blocks, 256, 0, st, a, n >> 2);
  else if (rgb4) C3D_LAUNCH(image_u8_rgb4_kernel<MODE>, (int)blocks, 256, 0, st, a);
  else C3D_LAUNCH(image_u8_generic_kernel<MODE>, (int)blocks, 256, 0, st, a, channels, channels_last ? 1 : 0);
  C3D_LAUNCH_CHECK();
  return C3D_OK;
}

}  // namespace c3d

using namespace c3d;

extern "C" int c3d_image_to_u8(const float* img, uint8_t* out, int32_t batch, int32_t channels, int32_t height,
                               int32_t width, int32_t channels_last, int32_t mode, double lo, double hi, void* stream) {
  C3D_CHECK_ARG(batch >= 0 && height >= 0 && width >= 0, "image_to_u8: negative size");
  C3D_CHECK_ARG(channels >= 1 && channels <= 4, "image_to_u8: 1..4 channels (PIL modes L, LA, RGB, RGBA), got %d", channels);
  C3D_CHECK_ARG(mode >= 0 && mode <= 2, "image_to_u8: mode must be 0 (save_image), 1 (tensor_to_PIL) or 2 (to_pil), got %d", mode);
  const long long hw = (long long)height * width;
  if (batch == 0 || hw == 0) return C3D_OK;
  C3D_CHECK_ARG(img && out, "image_to_u8: null pointer");
  int dev = 0;
  cudaGetDevice(&dev);
  const int sms = c3d_device_sm_count(dev);
  U8Args a;
  a.x = img;
  a.y = out;
  a.hw = hw;
  a.batch = batch;
  // torch turns the Python scalars low, high and max(high - low, 1e-5) into fp32 one by one (norm_ip's clamp_, sub_, div_)
  a.lo = (float)lo;
  a.hi = (float)hi;
  a.range = (float)(hi - lo > 1e-5 ? hi - lo : 1e-5);
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == C3D_U8_TENSOR_TO_PIL) return launch_u8<kTensorToPil>(a, channels, channels_last, sms, st);
  if (mode == C3D_U8_TO_PIL) return launch_u8<kToPil>(a, channels, channels_last, sms, st);
  int e = 0;
  if (frexpf(a.range, &e) == 0.5f && e > -100 && e < 100) {     // power of two: divide by multiplying with the exact reciprocal
    a.range = ldexpf(1.f, 1 - e);
    return launch_u8<kSavePow2>(a, channels, channels_last, sms, st);
  }
  return launch_u8<kSave>(a, channels, channels_last, sms, st);
}
