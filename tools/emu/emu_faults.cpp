// Deliberately broken micro-kernels: the emulator must REPORT each of them (tests/test_emu_cpu.py::test_emu_reports_*).
// Compiled only into libcips3d_b200_emu.so.  TEST INFRASTRUCTURE ONLY.
#include "../../cips-3d_b200/csrc/c3d_common.cuh"

namespace c3d {
namespace faults {

// shared-memory layout of every case: [operands ...][barriers at +48 KB]
struct Sm {
  alignas(1024) uint8_t a[128 * 64 * 2];      // A operand 128 x 64 fp16 (K-major canonical layout)
  alignas(1024) uint8_t b[128 * 64 * 2];      // B operand 128 x 64
  alignas(8) uint64_t bar[4];
  uint32_t tmem_base;
};

__device__ inline void fill_operands(Sm& s) {
  for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {
    reinterpret_cast<__half*>(s.a)[i] = __float2half_rn(0.01f * (float)(i % 97));
    reinterpret_cast<__half*>(s.b)[i] = __float2half_rn(0.02f * (float)(i % 89));
  }
}

// 0: correct reference case (must pass)        1: nobody arrives on the barrier the CTA waits on (deadlock)
// 3: warp 1 reads TMEM lanes 0..31 (lane-quarter violation)
// 4: the A operand is overwritten right after the MMA was issued, before it can have executed (operand race)
// 5: tcgen05.ld of the accumulator without waiting for the commit (read of in-flight columns)
// 6: tcgen05.ld beyond the allocated columns
__global__ void __launch_bounds__(128, 1) fault_kernel(int which, float* out) {
  C3D_DYN_SMEM(uint8_t, smem_raw);
  Sm& s = *reinterpret_cast<Sm*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&s.bar[0], 1);
    mbar_init(&s.bar[1], 4);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<128>(&s.tmem_base);
  fill_operands(s);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s.tmem_base;
  if (which == 1) {
    mbar_wait(&s.bar[1], 0);                  // count 4, nobody arrives
    return;
  }
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_f16(128, 128), dhi = umma_desc_hi(128);
    for (int k = 0; k < 4; ++k)
      umma_ss_w(tmem, umma_desc_lo(smem_u32(s.a) + k * 2 * 2048, 2048), umma_desc_lo(smem_u32(s.b) + k * 2 * 2048, 2048), dhi, idesc, k > 0);
    if (which == 4) reinterpret_cast<__half*>(s.a)[3] = __float2half_rn(123.f);      // too early: the MMAs may not have run yet
    tc_commit(&s.bar[0]);
  }
  if (which != 5) {
    mbar_wait(&s.bar[0], 0);
    tc_fence_after();
  }
  uint32_t v[16];
  const uint32_t lane_q = which == 3 ? 0u : (uint32_t)(warp * 32);
  const uint32_t col = which == 6 ? 120u : 0u;
  tmem_ld16(tmem + (lane_q << 16) + col, v);
  tc_wait_ld();
  if (out) out[threadIdx.x] = __uint_as_float(v[0]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<128>(tmem);
}

}  // namespace faults
}  // namespace c3d

extern "C" int c3d_emu_fault_case(int which, float* out128) {
  const size_t smem = sizeof(c3d::faults::Sm) + 1024;
  C3D_LAUNCH(c3d::faults::fault_kernel, 1, 128, smem, (cudaStream_t)0, which, out128);
  C3D_LAUNCH_CHECK();
  return C3D_OK;
}
