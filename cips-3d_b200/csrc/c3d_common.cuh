// Shared device/host helpers for libcips3d_b200 (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cips3d_b200.h"

// Every kernel launch and every PTX instruction of this library goes through the macros / wrappers of this
// header, so that a -DC3D_EMU build (g++, tools/emu/c3d_emu.h: functional CPU emulation used by the CPU test
// suite) can substitute them.  The product build never defines C3D_EMU.
#ifdef C3D_EMU
#include "c3d_emu.h"
#else
#define C3D_DYN_SMEM(type, name) extern __shared__ type name[]
#define C3D_DYN_SMEM_ALIGNED(type, name, al) extern __shared__ __align__(al) type name[]
#define C3D_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

// ---------------------------------------------------------------- host-side error plumbing
void c3d_set_error(const char* fmt, ...);
#define C3D_CHECK_ARG(cond, ...)        \
  do {                                  \
    if (!(cond)) {                      \
      c3d_set_error(__VA_ARGS__);       \
      return C3D_EINVAL;                \
    }                                   \
  } while (0)
#define C3D_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      c3d_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return C3D_ECUDA;                                                                  \
    }                                                                                    \
  } while (0)
void c3d_count_launch();
#define C3D_LAUNCH_CHECK()         \
  do {                             \
    c3d_count_launch();            \
    C3D_CUDA(cudaGetLastError());  \
  } while (0)

int c3d_device_sm_count(int dev);   // cached (api.cu)
// Kernel-variant selection (A/B switches of the build, not part of the drop-in contract).  Read from the environment ONCE,
// at the first use (or when the host layer calls c3d_reload_options()), never on a launch path.
#ifndef C3D_RAY_MATH_DEFAULT
#define C3D_RAY_MATH_DEFAULT 0
#endif
struct C3dOptions {
  int cips_cluster;   // C3D_CIPS_CLUSTER: 1 | 2 | 4   weight multicast clusters of the CIPS kernel (default 1)
  int cips_stagger_ns; // C3D_CIPS_STAGGER_NS: CTA i of the CIPS kernel starts i * this many ns late (default 0)
  int cips_res16;      // C3D_CIPS_RES16: fp16 residual stream in the CIPS kernel when only the image is asked for (default 1)
  int cips_pair;      // C3D_CIPS_PAIR:    0 | 1       tcgen05 cta_group::2 CTA pairs (default 1 since round 2's N = 256 / one-issuer form)
  int blur_impl;      // C3D_BLUR:         0 tile | 1 tma | 2 stream     4x4 FIR fast path (default: stream)
  int pigan_tc;       // C3D_PIGAN_IMPL:   simt -> 0 | tc -> 1           pi-GAN renderer (default tc: 9.4x faster, r02a)
  int pigan_pair;     // C3D_PIGAN_PAIR:   0 | 1 (default 0: measured 4 % slower, r02a)
  int ray_math;       // C3D_RAY_MATH:     block -> 0 | warp -> 1 | fold -> 2   per-ray math form of the renderer
};
const C3dOptions& c3d_options();
static inline int c3d_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------- PTX wrappers (device)
#if defined(__CUDACC__) || defined(C3D_EMU)
namespace c3d {
#ifdef C3D_EMU
#include "c3d_emu_ptx.h"
#else

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  // generic-proxy writes (st.shared) -> visible to the async proxy (UMMA / bulk copies)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
// The suspend-time hint lets a waiting warp sleep in hardware until the phase completes (or the hint
// expires) instead of re-issuing the probe: spinning waiters otherwise steal issue slots from the warps
// that share their scheduler (ncu: ~25 % of all executed instructions were wait-loop probes).
#ifndef C3D_SUSPEND_HINT_NS
#define C3D_SUSPEND_HINT_NS 20000
#endif
constexpr uint32_t kSuspendHintNs = C3D_SUSPEND_HINT_NS > 0 ? C3D_SUSPEND_HINT_NS : 0;
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
#if C3D_SUSPEND_HINT_NS > 0
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
#elif C3D_SUSPEND_HINT_NS == 0
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
#else      /* experiment builds: pure spin */
      "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
#endif
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(kSuspendHintNs)
      : "memory");
  return ok != 0;
}
// non-blocking probe (never suspends)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ unsigned long long c3d_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---- bulk async copy global -> shared (TMA engine, non-tensor form; SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---- tcgen05 / TMEM
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // one full warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// MMA completion -> mbarrier arrive (implies fence::before_thread_sync)
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T     (single-thread issue)
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_ss_w(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_ts_w(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// TMEM <-> registers, 32 lanes x 32 bit, N consecutive columns per thread (thread = lane)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
        "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
        "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
      "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]),
      "r"(v[15])
      : "memory");
}


// ---- named barriers, register budgets, clusters
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
template <int kId, int kThreads>
__device__ __forceinline__ void named_bar_sync_c() {
  asm volatile("bar.sync %0, %1;" ::"n"(kId), "n"(kThreads) : "memory");
}
template <int kThreads>
__device__ __forceinline__ void named_bar_sync_n(int id) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kThreads) : "memory");
}
template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// MMA completion -> arrive on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
// plain arrive on the barrier at this offset in CTA `rank` of the cluster.  Default semantics (release at CTA scope): the
// `.release.cluster` form compiles to MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR in front of the arrive, ~1.3 k clk per call
// (profiles/r02q_cips_trace_pair.txt: the relay warp of the CTA-pair kernel managed one arrive per 1.3 k clk and paced the whole
// pipeline).  What the remote waiter consumes here is never generic-proxy global data: it is shared memory filled by the
// TMA engine or written before a fence.proxy.async, and TMEM -- ordered by the fences the callers already issue.
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
      "r"(rank)
      : "memory");
}
// bulk copy multicast to the same offset (data and barrier) of every CTA in `mask`
__device__ __forceinline__ void bulk_g2s_mc(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
// ---- tcgen05 cta_group::2 (CTA pair: M = 256 across the two CTAs of a cluster, each CTA holds half of B's N rows)
template <int kCols>
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t* smem_result) {  // one full warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr) {  // one full warp in each CTA
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// issued by ONE thread of the leader CTA (cluster rank 0); descriptors are CTA-relative and apply to both CTAs
__device__ __forceinline__ void umma_ss_w_cg2(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from tensor memory (each CTA of the pair reads ITS OWN TMEM columns a_tmem), B from shared memory as above
__device__ __forceinline__ void umma_ts_w_cg2(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %4, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of this thread's earlier cta_group::2 MMAs -> arrive on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void tc_commit_cg2_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
// wait on a LOCAL barrier whose arrivals come from another CTA of the cluster.  Same default (CTA-scope acquire) form as the
// local wait, matching mbar_arrive_cluster above: the cluster-scope acquire costs ~350 clk per successful wait
// (profiles/r02o_cips_trace_pair.txt) and pairs with nothing once the arrive side is CTA-scope.
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(kSuspendHintNs)
      : "memory");
  return ok != 0;
}
// generic-proxy writes -> visible to the async proxy, all state spaces (operands another CTA's MMA will read)
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
#endif  // !C3D_EMU  (end of the PTX section; everything below is plain C++ shared with the emulated build)

// Spin with a watchdog: a protocol bug traps (-> CUDA error at the caller) instead of hanging the GPU.
// Time-based (globaltimer), so slow instrumented replays (ncu source counters) do not trip it.
#ifndef C3D_WATCHDOG_NS
#define C3D_WATCHDOG_NS 20000000000ull   /* 20 s */
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  unsigned long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 1023u) == 0) {
      const unsigned long long now = c3d_globaltimer();
      if (t0 == 0) t0 = now;
      else if (now - t0 > C3D_WATCHDOG_NS) {
        printf("c3d watchdog: mbarrier wait timed out (block %d thread %d bar 0x%x parity %u)\n", (int)blockIdx.x,
               (int)threadIdx.x, smem_u32(bar), parity);
        __trap();
      }
    }
  }
}


__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  unsigned long long t0 = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if ((++spins & 1023u) == 0) {
      const unsigned long long now = c3d_globaltimer();
      if (t0 == 0) t0 = now;
      else if (now - t0 > C3D_WATCHDOG_NS) {
        printf("c3d watchdog: cluster mbarrier wait timed out (block %d thread %d bar 0x%x parity %u)\n", (int)blockIdx.x,
               (int)threadIdx.x, smem_u32(bar), parity);
        __trap();
      }
    }
  }
}

// UMMA shared-memory matrix descriptor, K-major, no swizzle ("interleaved" canonical layout,
// cute/arch/mma_sm100_desc.hpp SmemDescriptor + cute/atom/mma_traits_sm100.hpp:273-303):
//   element (row r, k) lives at  (r%8)*16 + (r/8)*SBO + (k/8)*LBO + (k%8)*2   [bytes, 16-bit]
// i.e. 8x(16 B) core matrices of 128 contiguous bytes.
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_addr, uint32_t lbo_bytes,
                                                     uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version 1 (Blackwell)
  // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0)
  return d;
}
// Instruction descriptor for kind::f16 (cute UMMA::InstrDescriptor): fp16 A/B, fp32 D,
// both operands K-major, dense, no negate.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
  return (1u << 4)                      // c_format = F32
         | (0u << 7) | (0u << 10)       // a_format = b_format = F16
         | (0u << 15) | (0u << 16)      // a_major = b_major = K
         | ((uint32_t)(N >> 3) << 17)   // n_dim
         | ((uint32_t)(M >> 4) << 24);  // m_dim
}
// Descriptor words for the hot issue loops: the 64-bit K-major/no-swizzle descriptor is
//   lo = (addr >> 4) | (LBO >> 4) << 16,   hi = (SBO >> 4) | 1 << 14 (version)
// so stepping along K or to another ring stage is a 32-bit add on `lo`.
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr >> 4) & 0x3FFF) | (((lbo_bytes >> 4) & 0x3FFF) << 16);
}
__device__ __forceinline__ constexpr uint32_t umma_desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFF) | (1u << 14); }
// fp32 -> (hi, lo) fp16 split: hi = rn(x), lo = rn(x - hi); x ~= hi + lo to ~22 bits.
__device__ __forceinline__ void split_f16(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  __half2 h = __floats2half2_rn(x0, x1);
  float2 hf = __half22float2(h);
  __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = *reinterpret_cast<uint32_t*>(&l);
}
__device__ __forceinline__ uint32_t pack_f16(float x0, float x1) {
  __half2 h = __floats2half2_rn(x0, x1);
  return *reinterpret_cast<uint32_t*>(&h);
}

}  // namespace c3d
#endif  // __CUDACC__ || C3D_EMU
