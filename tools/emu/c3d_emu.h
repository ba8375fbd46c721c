// c3d_emu.h -- functional CPU emulation of the CUDA / PTX subset the cips3d_b200 kernels use.
//
// TEST INFRASTRUCTURE ONLY.  Compiling csrc/*.cu with  g++ -x c++ -DC3D_EMU -I tools/emu  produces
// libcips3d_b200_emu.so with the same C-ABI as the product library, in which "device pointers" are host
// pointers and every kernel runs on the CPU:
//   * every CUDA thread is a fiber (own stack, cooperative switches); CTAs of a launch run one cluster at a
//     time (persistent kernels do not communicate across clusters);
//   * mbarrier, named barriers, __syncthreads, warp collectives and cluster barriers are emulated with their
//     PTX semantics (phase parity, expect_tx / complete_tx, arrival counts, over-arrival = error);
//   * tensor memory is a 128-lane x 512-column array per CTA; tcgen05.ld/st enforce the "warp w touches lanes
//     32*(w%4)..+31" rule and the allocation bounds;
//   * tcgen05.mma decodes the real 64-bit shared-memory descriptors and the 32-bit instruction descriptor the
//     kernels build (K-major, no-swizzle canonical layout), reads the operands from the emulated shared /
//     tensor memory and accumulates in fp32.  MMAs and bulk copies are ASYNCHRONOUS: they sit in per-thread
//     queues and execute later (policy: eager / lazy / random, seeded), tcgen05.commit arrives only after the
//     issuing thread's earlier MMAs ran.  An operand that changes between issue and execution, a tcgen05.ld
//     of columns an in-flight MMA writes, or a tcgen05.st to columns an in-flight MMA reads is reported as a
//     race;
//   * a deadlock (no runnable fiber, nothing in flight) aborts the launch with a dump of who waits on what.
// Not modelled: timing, memory-model fences (they are no-ops), register budgets, swizzled layouts.
// The emulator is validated by running the hardware-validated kernels of this repo against the oracle
// (tests/test_emu_cpu.py); new kernels are then developed against it before they see a GPU.
#pragma once
#ifndef C3D_EMU
#error "c3d_emu.h is only for -DC3D_EMU builds"
#endif
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>

#include <algorithm>
#include <cmath>
#include <deque>
#include <functional>
#include <string>
#include <unordered_map>
#include <vector>

namespace emu {

// ------------------------------------------------------------------------------------------------ config
struct Config {
  int async_mode = 2;         // 0 eager (at issue), 1 lazy (only when nothing else can run), 2 random
  uint64_t seed = 1;
  int preempt_permille = 20;  // chance that an emulated primitive yields to another fiber
  int sms = 2;                // what c3d_device_sm_count() reports
  long long max_steps = 400000000ll;
  int verbose = 0;
};
inline Config& config() {
  static Config c = [] {
    Config c;
    if (const char* e = getenv("C3D_EMU_ASYNC")) c.async_mode = atoi(e);
    if (const char* e = getenv("C3D_EMU_SEED")) c.seed = strtoull(e, nullptr, 10);
    if (const char* e = getenv("C3D_EMU_PREEMPT")) c.preempt_permille = atoi(e);
    if (const char* e = getenv("C3D_EMU_SMS")) c.sms = atoi(e);
    if (const char* e = getenv("C3D_EMU_VERBOSE")) c.verbose = atoi(e);
    return c;
  }();
  return c;
}

struct Rng {
  uint64_t s = 0x9E3779B97F4A7C15ull;
  void seed(uint64_t v) { s = v * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull; next(); }
  uint64_t next() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
  uint32_t below(uint32_t n) { return (uint32_t)((next() >> 11) % n); }
};

// ------------------------------------------------------------------------------------------------ fibers
extern "C" void c3d_emu_switch(void** save_sp, void* load_sp);
#ifdef C3D_EMU_IMPL
asm(R"(
.text
.globl c3d_emu_switch
.type c3d_emu_switch,@function
c3d_emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size c3d_emu_switch,.-c3d_emu_switch
)");
#endif

constexpr size_t kStackBytes = 128 * 1024;
constexpr int kTmemLanes = 128, kTmemCols = 512;

struct Cta;
// A converged warp executes mbarrier.try_wait / test_wait as ONE instruction: every lane sees the same outcome.
// The fibers of a warp run independently, so the emulation makes the outcome of "the k-th wait of this warp on
// barrier b" common to its lanes: once one lane has seen it succeed, it has succeeded for all of them (otherwise
// a slow lane could miss a whole phase that its siblings observed -- impossible in hardware).
struct WaitRec {
  uint64_t wait_cnt[32] = {};
  uint64_t passed_upto = 0;
  uint64_t test_cnt[32] = {};
  uint64_t test_base = 0;          // index of test_results.front()
  std::deque<uint8_t> test_results;
};
struct Warp {
  uint32_t arrived = 0, exited = 0, gen = 0;
  uint32_t vals[32];
  uint32_t shfl[32];
  uint32_t result = 0;
  int nthreads = 32;
  std::unordered_map<void*, WaitRec> waits;
};
typedef bool (*CondFn)(void*, uint64_t);
struct Fiber {
  void* sp = nullptr;
  uint8_t* stack = nullptr;
  Cta* cta = nullptr;
  Warp* warp = nullptr;
  int tid = 0, lane = 0, wid = 0;
  uint3 thread_idx;
  bool done = false;
  CondFn cond = nullptr;
  void* cond_a = nullptr;
  uint64_t cond_b = 0;
  const char* wait_what = "";
};

struct MBar {      // lives in the 8 bytes of the kernel's uint64_t mbarrier
  int32_t tx;
  int16_t pending;
  uint8_t init;
  uint8_t phase;
};
static_assert(sizeof(MBar) == 8, "mbarrier state must fit the 8-byte object");

struct AsyncOp {
  std::function<void()> run;
  // hazard bookkeeping for MMAs (ranges in TMEM columns; -1 = none)
  int d_col0 = -1, d_col1 = -1, a_col0 = -1, a_col1 = -1;
  bool both_ctas = false;
  const char* what = "";
};
struct AsyncQueue {
  int owner_tid;
  std::deque<AsyncOp> q;
};

struct NamedBar {
  int arrived = 0;
  uint32_t gen = 0;
};

struct Cta {
  uint3 block_idx;
  int rank = 0;                  // rank in cluster
  uint8_t* smem = nullptr;       // dynamic shared memory window (1024-aligned)
  size_t smem_bytes = 0;
  std::vector<uint32_t> tmem;    // [lane][col]
  int tmem_allocated = 0;        // bump allocator (columns)
  std::vector<AsyncQueue> queues;   // per issuing thread (tensor pipe) + one for bulk copies (owner -1)
  NamedBar bars[16];
  int live_threads = 0;
  std::vector<Warp> warps;
  Cta() : tmem((size_t)kTmemLanes * kTmemCols, 0xDEADBEEFu) {}
  uint32_t& T(int lane, int col) { return tmem[(size_t)lane * kTmemCols + col]; }
};

struct Launch {
  dim3 grid, block;
  int cluster = 1;
  std::function<void()> body;
  std::vector<Cta*> ctas;        // CTAs of the cluster being executed
  std::vector<Fiber> fibers;
  void* sched_sp = nullptr;
  bool failed = false;
  std::string error;
  long long steps = 0;
  NamedBar cluster_bar;
  Rng rng;
  const char* name = "";
};

struct Global {
  Launch* launch = nullptr;
  Fiber* cur = nullptr;
  std::vector<std::pair<void*, uint32_t>> foreign_smem;   // static __shared__ objects seen by smem_u32
  std::vector<uint8_t*> stack_pool;
  std::vector<uint8_t*> smem_pool;
  unsigned long long timer = 0;
  int last_error = 0;
};
#ifdef C3D_EMU_IMPL
Global g_state;
#else
extern Global g_state;
#endif
inline Global& G() { return g_state; }
inline Fiber* cur() { return G().cur; }
inline Launch& L() { return *G().launch; }

struct Failure {};
[[noreturn]] inline void fail(const char* fmt, ...);
inline void to_scheduler() { Fiber* f = cur(); c3d_emu_switch(&f->sp, L().sched_sp); }
inline void yield() { to_scheduler(); }
inline void preempt_point() {
  Launch& l = L();
  if (config().preempt_permille > 0 && (int)l.rng.below(1000) < config().preempt_permille) yield();
}
inline void wait_until(CondFn fn, void* a, uint64_t b, const char* what) {
  if (fn(a, b)) { preempt_point(); return; }
  Fiber* f = cur();
  f->cond = fn; f->cond_a = a; f->cond_b = b; f->wait_what = what;
  to_scheduler();
}

#include <stdarg.h>
[[noreturn]] inline void fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  Launch& l = L();
  Fiber* f = cur();
  char where[128] = "";
  if (f) snprintf(where, sizeof(where), " [kernel %s block %u thread %d]", l.name, f->cta->block_idx.x, f->tid);
  if (!l.failed) { l.failed = true; l.error = std::string(buf) + where; }
  if (f) {
    f->done = true;
    for (;;) to_scheduler();
  }
  throw Failure();     // called from an async op / wait condition on the scheduler stack
}

// ------------------------------------------------------------------------------------------------ addresses
// shared::cluster address = rank << 24 | offset into the CTA's dynamic shared window (+ 0x100 so 0 is invalid)
constexpr uint32_t kSmemBias = 0x400;
inline uint32_t smem_addr_of(const void* p) {
  Launch& l = L();
  for (Cta* c : l.ctas) {
    if ((const uint8_t*)p >= c->smem && (const uint8_t*)p < c->smem + c->smem_bytes)
      return ((uint32_t)c->rank << 24) | (uint32_t)((const uint8_t*)p - c->smem + kSmemBias);
  }
  // a static __shared__ object: hand out a fake address in the 0x00F00000 region
  auto& fs = G().foreign_smem;
  for (auto& e : fs) if (e.first == p) return e.second;
  uint32_t a = 0x00F00000u + (uint32_t)fs.size() * 16u;
  fs.push_back({const_cast<void*>(p), a});
  return a;
}
inline void* smem_ptr_of(uint32_t addr, size_t bytes, const char* what) {
  Launch& l = L();
  const uint32_t off = addr & 0xFFFFFFu;
  if (off >= 0x00F00000u) {
    for (auto& e : G().foreign_smem) if (e.second == off) return e.first;
    fail("%s: unknown static shared address 0x%x", what, addr);
  }
  const uint32_t rank = addr >> 24;
  if (rank >= l.ctas.size()) fail("%s: shared::cluster address 0x%x names CTA rank %u of a %zu-CTA cluster", what, addr, rank, l.ctas.size());
  Cta* c = l.ctas[rank];
  if (off < kSmemBias || off - kSmemBias + bytes > c->smem_bytes) fail("%s: shared address 0x%x (+%zu) out of the CTA's %zu-byte window", what, addr, bytes, c->smem_bytes);
  return c->smem + (off - kSmemBias);
}
// the 14-bit (>>4) address field of an MMA descriptor: offset inside the ISSUING CTA's window (or `cta`)
inline uint8_t* desc_ptr(Cta* c, uint32_t addr18, size_t span, const char* what) {
  const uint32_t off = addr18 & 0x3FFFFu;
  if (off < kSmemBias || off - kSmemBias + span > c->smem_bytes) fail("%s: descriptor address 0x%x (+%zu) outside shared memory", what, off, span);
  return c->smem + (off - kSmemBias);
}

// ------------------------------------------------------------------------------------------------ mbarrier
inline MBar* mb(void* p) { return reinterpret_cast<MBar*>(p); }
inline void mbar_check(MBar* b) {
  if (b->pending < 0) fail("mbarrier over-arrived (pending %d, init %d)", (int)b->pending, (int)b->init);
  if (b->pending == 0 && b->tx == 0) { b->phase ^= 1; b->pending = b->init; }
}
inline void mbar_do_init(void* bar, uint32_t count) {
  if (count == 0 || count > 255) fail("mbarrier.init with count %u", count);
  MBar* b = mb(bar); b->tx = 0; b->pending = (int16_t)count; b->init = (uint8_t)count; b->phase = 0;
}
inline void mbar_do_arrive(void* bar) { MBar* b = mb(bar); if (b->init == 0) fail("arrive on an uninitialised mbarrier"); b->pending -= 1; mbar_check(b); }
inline void mbar_do_expect_tx(void* bar, uint32_t bytes) { MBar* b = mb(bar); b->tx += (int32_t)bytes; }
inline void mbar_do_complete_tx(void* bar, uint32_t bytes) { MBar* b = mb(bar); b->tx -= (int32_t)bytes; mbar_check(b); }
inline bool mbar_phase_done(void* bar, uint64_t parity) { return (mb(bar)->phase & 1u) != (uint32_t)parity; }
struct WarpWait { WaitRec* rec; void* bar; uint32_t parity; uint64_t k; };
inline bool warp_wait_ok(void* a, uint64_t) {
  WarpWait* w = (WarpWait*)a;
  if (w->k <= w->rec->passed_upto) return true;
  if (mbar_phase_done(w->bar, w->parity)) { w->rec->passed_upto = w->k; return true; }
  return false;
}
inline void mbar_warp_wait(void* bar, uint32_t parity) {
  Fiber* f = cur();
  WaitRec& r = f->warp->waits[bar];
  WarpWait w{&r, bar, parity, ++r.wait_cnt[f->lane]};
  wait_until(warp_wait_ok, &w, 0, "mbarrier.try_wait");    // `w` lives on this fiber's stack while it is blocked
}
inline bool mbar_warp_test(void* bar, uint32_t parity) {
  Fiber* f = cur();
  WaitRec& r = f->warp->waits[bar];
  const uint64_t k = r.test_cnt[f->lane]++;
  if (k < r.test_base) return mbar_phase_done(bar, parity);          // record already dropped
  if (k - r.test_base < r.test_results.size()) return r.test_results[k - r.test_base] != 0;
  const bool ok = mbar_phase_done(bar, parity);
  r.test_results.push_back(ok ? 1 : 0);
  if (r.test_results.size() > 4096) { r.test_results.pop_front(); r.test_base++; }
  return ok;
}

// ------------------------------------------------------------------------------------------------ async engine
inline AsyncQueue& queue_for(Cta* c, int tid) {
  for (auto& q : c->queues) if (q.owner_tid == tid) return q;
  c->queues.push_back(AsyncQueue{tid, {}});
  return c->queues.back();
}
inline void submit_async(Cta* c, int tid, AsyncOp&& op) {
  if (config().async_mode == 0) { op.run(); return; }
  queue_for(c, tid).q.push_back(std::move(op));
}
// run one queued op (front of a random non-empty queue); false if nothing is in flight
inline bool run_one_async() {
  Launch& l = L();
  int n = 0;
  for (Cta* c : l.ctas) for (auto& q : c->queues) if (!q.q.empty()) ++n;
  if (!n) return false;
  int pick = (int)l.rng.below((uint32_t)n);
  for (Cta* c : l.ctas)
    for (auto& q : c->queues)
      if (!q.q.empty() && pick-- == 0) {
        AsyncOp op = std::move(q.q.front());
        q.q.pop_front();
        Fiber* saved = G().cur;
        G().cur = nullptr;        // async ops run on the scheduler stack
        try { op.run(); } catch (Failure&) {}
        G().cur = saved;
        return true;
      }
  return false;
}

// ------------------------------------------------------------------------------------------------ scheduler
inline void fiber_entry();
inline uint8_t* take_stack() {
  auto& pool = G().stack_pool;
  if (!pool.empty()) { uint8_t* s = pool.back(); pool.pop_back(); return s; }
  void* p = mmap(nullptr, kStackBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  if (p == MAP_FAILED) { perror("c3d_emu: mmap stack"); abort(); }
  return (uint8_t*)p;
}
inline void init_fiber_stack(Fiber& f) {
  uint64_t* top = reinterpret_cast<uint64_t*>(f.stack + kStackBytes);
  // layout popped by c3d_emu_switch: r15 r14 r13 r12 rbx rbp, then `ret` to fiber_entry with rsp % 16 == 8
  top -= 1; *top = 0;                                  // fake return address of fiber_entry
  top -= 1; *top = (uint64_t)(void (*)())&fiber_entry;
  for (int i = 0; i < 6; ++i) { top -= 1; *top = 0; }
  f.sp = top;
}
inline void thread_exit_bookkeeping(Fiber* f) {
  f->done = true;
  f->cta->live_threads -= 1;
  f->warp->exited |= 1u << f->lane;
}
inline void fiber_entry() {
  Launch& l = L();
  l.body();
  Fiber* f = cur();
  thread_exit_bookkeeping(f);
  for (;;) to_scheduler();
}

inline std::string describe_deadlock() {
  Launch& l = L();
  std::string s = "deadlock: no runnable thread and nothing in flight;";
  struct Grp { const char* what; void* a; uint64_t b; int cta; int first_tid, last_tid, n; };
  std::vector<Grp> grps;
  int blocked = 0;
  for (Fiber& f : l.fibers) {
    if (f.done) continue;
    ++blocked;
    bool found = false;
    void* ca = f.cond_a;
    uint64_t cb = f.cond_b;
    if (f.cond == warp_wait_ok) { ca = ((WarpWait*)f.cond_a)->bar; cb = ((WarpWait*)f.cond_a)->parity; }
    for (Grp& g : grps)
      if (g.what == f.wait_what && g.a == ca && g.b == cb && g.cta == f.cta->rank) { g.last_tid = f.tid; g.n++; found = true; break; }
    if (!found) grps.push_back(Grp{f.wait_what, ca, cb, f.cta->rank, f.tid, f.tid, 1});
  }
  for (Grp& g : grps) {
    char b[256];
    long off = -1;
    for (Cta* c : l.ctas) if ((uint8_t*)g.a >= c->smem && (uint8_t*)g.a < c->smem + c->smem_bytes) off = (long)((uint8_t*)g.a - c->smem);
    if (off >= 0 && strstr(g.what, "mbarrier")) {
      MBar* m = (MBar*)g.a;
      snprintf(b, sizeof(b), "\n  cta %d threads %d..%d (%d): %s parity %llu on smem+%ld {phase %d pending %d/%d tx %d}", g.cta, g.first_tid, g.last_tid, g.n,
               g.what, (unsigned long long)g.b, off, (int)m->phase, (int)m->pending, (int)m->init, (int)m->tx);
    } else {
      snprintf(b, sizeof(b), "\n  cta %d threads %d..%d (%d): %s arg %llu", g.cta, g.first_tid, g.last_tid, g.n, g.what, (unsigned long long)g.b);
    }
    s += b;
  }
  char b[64];
  snprintf(b, sizeof(b), "\n  (%d threads blocked)", blocked);
  return s + b;
}

// run the fibers of the current cluster to completion
inline void run_cluster() {
  Launch& l = L();
  const int n = (int)l.fibers.size();
  int live = n;
  const int mode = config().async_mode;
  unsigned passes = 0;
  while (live > 0 && !l.failed) {
    bool progress = false;
    const int start = (int)l.rng.below((uint32_t)n);
    for (int k = 0; k < n && !l.failed; ++k) {
      Fiber& f = l.fibers[(start + k) % n];
      if (f.done) continue;
      if (f.cond) {
        bool ok = false;
        try { ok = f.cond(f.cond_a, f.cond_b); } catch (Failure&) {}
        if (!ok) continue;
        f.cond = nullptr;
      }
      G().cur = &f;
      c3d_emu_switch(&l.sched_sp, f.sp);
      G().cur = nullptr;
      progress = true;
      if (f.done) --live;
      if (++l.steps > config().max_steps) { l.failed = true; l.error = "step budget exhausted (livelock?)"; break; }
      if (mode == 2 && l.rng.below(4) == 0) run_one_async();
    }
    // lazy mode: asynchronous work completes only when nobody can run -- or, so that polling loops
    // (test_wait + nanosleep: always "runnable") cannot starve it, every 8th pass over the fibers
    if (mode == 1 && (++passes & 7) == 0 && run_one_async()) progress = true;
    if (l.failed) break;
    // recount (threads may have been marked done by fail())
    live = 0;
    for (Fiber& f : l.fibers) if (!f.done) ++live;
    if (!progress) {
      if (run_one_async()) continue;
      if (live > 0) { l.failed = true; l.error = describe_deadlock(); }
    }
  }
  // drain whatever is still in flight (e.g. a commit nobody waits for)
  if (!l.failed) while (run_one_async()) {}
}

template <class F>
inline int launch(const char* name, dim3 grid, dim3 block, size_t dyn_smem, int cluster, F&& body) {
  Global& g = G();
  if (g.launch) { fprintf(stderr, "c3d_emu: nested launch\n"); abort(); }
  Launch l;
  l.grid = grid; l.block = block; l.cluster = cluster < 1 ? 1 : cluster;
  l.body = std::function<void()>(body);
  l.name = name;
  l.rng.seed(config().seed);
  g.launch = &l;
  if (config().verbose) fprintf(stderr, "c3d_emu: launch %s grid (%u,%u,%u) block %u cluster %d smem %zu\n", name, grid.x, grid.y, grid.z, block.x, l.cluster, dyn_smem);
  const int nthreads = (int)(block.x * block.y * block.z);
  const long long nblocks = (long long)grid.x * grid.y * grid.z;
  if (grid.x % l.cluster) { fprintf(stderr, "c3d_emu: grid.x %% cluster != 0\n"); abort(); }
#if defined(__SANITIZE_ADDRESS__)
  const size_t smem_alloc = dyn_smem ? dyn_smem : 1;      // exact: the first byte past the request is an ASan redzone
#else
  const size_t smem_alloc = ((dyn_smem + 1023) / 1024 + 1) * 1024;
#endif
  std::vector<Cta> ctas(l.cluster);
  for (int r = 0; r < l.cluster; ++r) {
    uint8_t* raw;
    if (posix_memalign((void**)&raw, 1024, smem_alloc)) abort();
    ctas[r].smem = raw;
    ctas[r].smem_bytes = dyn_smem;
    ctas[r].rank = r;
    ctas[r].warps.resize((nthreads + 31) / 32);
    l.ctas.push_back(&ctas[r]);
  }
  l.fibers.resize((size_t)nthreads * l.cluster);
  for (Fiber& f : l.fibers) f.stack = take_stack();
  int rc = 0;
  for (long long b0 = 0; b0 < nblocks && !l.failed; b0 += l.cluster) {
    g.foreign_smem.clear();
    l.cluster_bar = NamedBar();
    for (int r = 0; r < l.cluster; ++r) {
      Cta& c = ctas[r];
      const long long b = b0 + r;
      c.block_idx = make_uint3((unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((long long)grid.x * grid.y)));
      memset(c.smem, 0xCD, dyn_smem);
      std::fill(c.tmem.begin(), c.tmem.end(), 0xDEADBEEFu);
      c.tmem_allocated = 0;
      c.queues.clear();
      for (auto& nb : c.bars) nb = NamedBar();
      c.live_threads = nthreads;
      for (size_t w = 0; w < c.warps.size(); ++w) {
        c.warps[w] = Warp();
        const int rem = nthreads - (int)w * 32;
        c.warps[w].nthreads = rem < 32 ? rem : 32;
        if (rem < 32) c.warps[w].exited = ~((1u << rem) - 1u);   // lanes that do not exist
      }
      for (int t = 0; t < nthreads; ++t) {
        Fiber& f = l.fibers[(size_t)r * nthreads + t];
        uint8_t* st = f.stack;
        f = Fiber();
        f.stack = st;
        f.cta = &c;
        f.tid = t; f.lane = t & 31; f.wid = t >> 5;
        f.warp = &c.warps[f.wid];
        f.thread_idx = make_uint3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
        init_fiber_stack(f);
      }
    }
    run_cluster();
  }
  if (l.failed) {
    fprintf(stderr, "c3d_emu: kernel %s FAILED: %s\n", name, l.error.c_str());
    rc = 1;
    g.last_error = 1;
  }
  for (Fiber& f : l.fibers) g.stack_pool.push_back(f.stack);
  for (int r = 0; r < l.cluster; ++r) free(ctas[r].smem);
  g.launch = nullptr;
  g.cur = nullptr;
  return rc;
}

// ------------------------------------------------------------------------------------------------ block / warp sync
inline bool gen_changed(void* a, uint64_t gen) { return *(volatile uint32_t*)a != (uint32_t)gen; }
inline void named_barrier(NamedBar& nb, int expected, const char* what) {
  const uint32_t gen = nb.gen;
  if (++nb.arrived == expected) { nb.arrived = 0; nb.gen++; preempt_point(); return; }
  if (nb.arrived > expected) fail("%s: %d arrivals, %d expected", what, nb.arrived, expected);
  wait_until(gen_changed, &nb.gen, gen, what);
}
inline void syncthreads() {
  Fiber* f = cur();
  // exited threads do not participate (CUDA semantics since Volta: counted as arrived)
  NamedBar& nb = f->cta->bars[0];
  const uint32_t gen = nb.gen;
  nb.arrived++;
  struct H { static bool ok(void* a, uint64_t gen) {
      Cta* c = (Cta*)a; NamedBar& nb = c->bars[0];
      if (nb.gen != (uint32_t)gen) return true;
      if (nb.arrived >= c->live_threads) { nb.arrived = 0; nb.gen++; return true; }
      return false; } };
  (void)gen;
  wait_until(H::ok, f->cta, gen, "__syncthreads");
}
inline void bar_sync(int id, int nthreads) {
  if (id < 1 || id > 15) fail("bar.sync id %d", id);
  named_barrier(cur()->cta->bars[id], nthreads, "bar.sync");
}
inline void cluster_sync() {
  Launch& l = L();
  int total = 0;
  for (Cta* c : l.ctas) total += c->live_threads;
  named_barrier(l.cluster_bar, total, "barrier.cluster");
}
// generic full-warp collective: every live lane contributes `v`; `combine` runs once when all arrived
template <class Combine>
inline uint32_t warp_collective(uint32_t mask, uint32_t v, Combine combine, const char* what) {
  Fiber* f = cur();
  Warp* w = f->warp;
  const uint32_t expect = mask & ~w->exited;
  if (!(expect >> f->lane & 1u)) fail("%s: calling lane %d not in mask 0x%x", what, f->lane, mask);
  const uint32_t gen = w->gen;
  w->vals[f->lane] = v;
  w->arrived |= 1u << f->lane;
  if ((w->arrived & expect) == expect) {
    w->result = combine(w->vals, expect);
    w->arrived = 0;
    w->gen++;
    preempt_point();
    return w->result;
  }
  wait_until(gen_changed, &w->gen, gen, what);
  return w->result;
}
inline void syncwarp(uint32_t mask = 0xffffffffu) {
  warp_collective(mask, 0, [](const uint32_t*, uint32_t) { return 0u; }, "__syncwarp");
}
inline int all_sync(uint32_t mask, int pred) {
  return (int)warp_collective(mask, pred ? 1u : 0u, [](const uint32_t* v, uint32_t m) {
    uint32_t r = 1; for (int i = 0; i < 32; ++i) if (m >> i & 1u) r &= v[i]; return r; }, "__all_sync");
}
inline int any_sync(uint32_t mask, int pred) {
  return (int)warp_collective(mask, pred ? 1u : 0u, [](const uint32_t* v, uint32_t m) {
    uint32_t r = 0; for (int i = 0; i < 32; ++i) if (m >> i & 1u) r |= v[i]; return r; }, "__any_sync");
}
inline uint32_t ballot_sync(uint32_t mask, int pred) {
  return warp_collective(mask, pred ? 1u : 0u, [](const uint32_t* v, uint32_t m) {
    uint32_t r = 0; for (int i = 0; i < 32; ++i) if ((m >> i & 1u) && v[i]) r |= 1u << i; return r; }, "__ballot_sync");
}
// shuffles: publish into a dedicated per-warp table, rendezvous, read, rendezvous (so nobody republishes early)
inline uint32_t shfl_idx(uint32_t mask, uint32_t v, int src) {
  Fiber* f = cur();
  Warp* w = f->warp;
  w->shfl[f->lane] = v;
  syncwarp(mask);
  const uint32_t r = w->shfl[src & 31];
  syncwarp(mask);
  return r;
}

// ------------------------------------------------------------------------------------------------ TMEM + MMA
inline void tmem_check_warp(uint32_t taddr, int ncols, const char* what) {
  Fiber* f = cur();
  const int lane0 = (int)(taddr >> 16), col = (int)(taddr & 0xFFFF);
  if (lane0 != 32 * (f->wid & 3)) fail("%s: warp %d may only access TMEM lanes %d..%d, address names lane %d", what, f->wid, 32 * (f->wid & 3), 32 * (f->wid & 3) + 31, lane0);
  if (col < 0 || col + ncols > f->cta->tmem_allocated) fail("%s: columns %d..%d outside the allocation (%d columns)", what, col, col + ncols - 1, f->cta->tmem_allocated);
}
inline void tmem_hazard(Cta* c, int col0, int col1, bool is_store, const char* what) {
  for (Cta* oc : L().ctas)
  for (auto& q : oc->queues)
    for (auto& op : q.q) {
      if (oc != c && !op.both_ctas) continue;     // a cta_group::2 MMA touches the same columns in both CTAs
      if (op.d_col0 >= 0 && col0 < op.d_col1 && op.d_col0 < col1)
        fail("race: %s of TMEM columns %d..%d while a queued %s still writes columns %d..%d", what, col0, col1 - 1, op.what, op.d_col0, op.d_col1 - 1);
      if (is_store && op.a_col0 >= 0 && col0 < op.a_col1 && op.a_col0 < col1)
        fail("race: %s of TMEM columns %d..%d while a queued %s still reads them as its A operand (%d..%d)", what, col0, col1 - 1, op.what, op.a_col0, op.a_col1 - 1);
    }
}
template <int N>
inline void tmem_ld(uint32_t taddr, uint32_t* v) {
  tmem_check_warp(taddr, N, "tcgen05.ld");
  Fiber* f = cur();
  const int lane = (int)(taddr >> 16) + f->lane, col = (int)(taddr & 0xFFFF);
  tmem_hazard(f->cta, col, col + N, false, "tcgen05.ld");
  for (int j = 0; j < N; ++j) v[j] = f->cta->T(lane, col + j);
  preempt_point();
}
template <int N>
inline void tmem_st(uint32_t taddr, const uint32_t* v) {
  tmem_check_warp(taddr, N, "tcgen05.st");
  Fiber* f = cur();
  const int lane = (int)(taddr >> 16) + f->lane, col = (int)(taddr & 0xFFFF);
  tmem_hazard(f->cta, col, col + N, true, "tcgen05.st");
  for (int j = 0; j < N; ++j) f->cta->T(lane, col + j) = v[j];
  preempt_point();
}
inline void tmem_do_alloc(uint32_t* smem_result, int cols) {
  Fiber* f = cur();
  if (cols < 32 || cols > 512 || (cols & (cols - 1))) fail("tcgen05.alloc: %d columns (power of two in [32,512] required)", cols);
  if (f->lane == 0) {
    Cta* c = f->cta;
    if (c->tmem_allocated + cols > kTmemCols) fail("tcgen05.alloc: out of tensor memory (%d + %d)", c->tmem_allocated, cols);
    *smem_result = (uint32_t)c->tmem_allocated;
    c->tmem_allocated += cols;
  }
  syncwarp();
}
inline void tmem_do_dealloc(uint32_t taddr, int cols) {
  Fiber* f = cur();
  syncwarp();
  if (f->lane == 0) {
    Cta* c = f->cta;
    for (auto& q : c->queues) if (!q.q.empty() && q.owner_tid >= 0) fail("tcgen05.dealloc with MMAs still in flight");
    (void)taddr;
    c->tmem_allocated -= cols;
    if (c->tmem_allocated < 0) fail("tcgen05.dealloc: more columns freed than allocated");
  }
}

struct SmemDesc { uint32_t addr, lbo, sbo; };
inline SmemDesc decode_desc(uint64_t d, const char* what) {
  SmemDesc s;
  s.addr = (uint32_t)(d & 0x3FFF) << 4;
  s.lbo = (uint32_t)((d >> 16) & 0x3FFF) << 4;
  s.sbo = (uint32_t)((d >> 32) & 0x3FFF) << 4;
  const uint32_t version = (uint32_t)((d >> 46) & 3), layout = (uint32_t)((d >> 61) & 7), base_off = (uint32_t)((d >> 49) & 7);
  if (version != 1) fail("%s: descriptor version %u (sm_100 needs 1)", what, version);
  if (layout != 0) fail("%s: swizzled layout %u is not emulated", what, layout);
  if (base_off != 0) fail("%s: base_offset %u", what, base_off);
  if (d & (1ull << 52)) fail("%s: lbo_mode set", what);
  return s;
}
struct InstrDesc { int M, N; };
inline InstrDesc decode_idesc(uint32_t id, int cta_group) {
  if ((id >> 4 & 3u) != 1u) fail("tcgen05.mma: c_format must be F32");
  if ((id >> 7 & 7u) != 0u || (id >> 10 & 7u) != 0u) fail("tcgen05.mma: only F16 A/B formats are emulated (idesc 0x%x)", id);
  if ((id >> 15 & 1u) || (id >> 16 & 1u)) fail("tcgen05.mma: only K-major operands are emulated");
  if (id & 0xFu) fail("tcgen05.mma: sparsity / saturate bits set");
  if ((id >> 13 & 3u)) fail("tcgen05.mma: negate bits set");
  InstrDesc r;
  r.N = (int)(id >> 17 & 0x3Fu) << 3;
  r.M = (int)(id >> 24 & 0x1Fu) << 4;
  const int m_ok = cta_group == 1 ? 128 : 256;
  if (r.M != m_ok) fail("tcgen05.mma: M = %d not emulated for cta_group::%d (need %d)", r.M, cta_group, m_ok);
  const int nstep = 16;
  if (r.N < nstep || r.N > 256 || r.N % nstep) fail("tcgen05.mma: N = %d invalid for M = %d", r.N, r.M);
  return r;
}
inline float h2f(uint16_t h) { __half_raw r; r.x = h; return __half2float(__half(r)); }
inline uint16_t canon_elem(const uint8_t* base, const SmemDesc& d, int r, int k) {
  const size_t off = (size_t)(r % 8) * 16 + (size_t)(r / 8) * d.sbo + (size_t)(k / 8) * d.lbo + (size_t)(k % 8) * 2;
  uint16_t v;
  memcpy(&v, base + off, 2);
  return v;
}
inline size_t canon_span(const SmemDesc& d, int rows) { return (size_t)((rows - 1) / 8) * d.sbo + (size_t)d.lbo + 128; }
inline uint64_t hash_operand(const uint8_t* base, const SmemDesc& d, int rows) {
  uint64_t h = 1469598103934665603ull;
  for (int r = 0; r < rows; ++r)
    for (int k = 0; k < 16; ++k) h = (h ^ canon_elem(base, d, r, k)) * 1099511628211ull;
  return h;
}
// D[128 x N] (+)= A[128 x 16] * B[N x 16]^T for one CTA (cta_group::1).  a_tmem_col < 0: A from shared memory.
inline void mma_execute_cg1(Cta* c, int d_col, bool a_from_tmem, int a_col, uint8_t* a_base, SmemDesc ad, uint8_t* b_base,
                            SmemDesc bd, int N, bool accumulate) {
  std::vector<float> B((size_t)N * 16);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < 16; ++k) B[(size_t)n * 16 + k] = h2f(canon_elem(b_base, bd, n, k));
  for (int r = 0; r < 128; ++r) {
    float a[16];
    for (int k = 0; k < 16; ++k) {
      if (a_from_tmem) {
        const uint32_t w = c->T(r, a_col + k / 2);
        a[k] = h2f((uint16_t)((k & 1) ? (w >> 16) : (w & 0xFFFF)));
      } else {
        a[k] = h2f(canon_elem(a_base, ad, r, k));
      }
    }
    for (int n = 0; n < N; ++n) {
      uint32_t& dw = c->T(r, d_col + n);
      float acc = 0.f;
      if (accumulate) memcpy(&acc, &dw, 4);
      const float* b = &B[(size_t)n * 16];
      for (int k = 0; k < 16; ++k) acc += a[k] * b[k];
      memcpy(&dw, &acc, 4);
    }
  }
}
inline void mma_issue_cg1(uint32_t d_tmem, bool a_from_tmem, uint64_t a_desc_or_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  Fiber* f = cur();
  Cta* c = f->cta;
  const InstrDesc id = decode_idesc(idesc, 1);
  if (d_tmem >> 16) fail("tcgen05.mma: D address names lane %u (must be 0 for M = 128)", d_tmem >> 16);
  const int d_col = (int)(d_tmem & 0xFFFF);
  if (d_col + id.N > c->tmem_allocated) fail("tcgen05.mma: D columns %d..%d outside the allocation", d_col, d_col + id.N - 1);
  SmemDesc ad{}, bd = decode_desc(b_desc, "tcgen05.mma B");
  uint8_t* a_base = nullptr;
  int a_col = -1;
  if (a_from_tmem) {
    if ((uint32_t)a_desc_or_tmem >> 16) fail("tcgen05.mma: A (TMEM) address names lane %u", (uint32_t)a_desc_or_tmem >> 16);
    a_col = (int)(a_desc_or_tmem & 0xFFFF);
    if (a_col + 8 > c->tmem_allocated) fail("tcgen05.mma: A columns outside the allocation");
    if (a_col < d_col + id.N && d_col < a_col + 8) fail("tcgen05.mma: A (TMEM cols %d..%d) overlaps D (%d..%d)", a_col, a_col + 7, d_col, d_col + id.N - 1);
  } else {
    ad = decode_desc(a_desc_or_tmem, "tcgen05.mma A");
    a_base = desc_ptr(c, ad.addr, canon_span(ad, 128), "tcgen05.mma A");
  }
  uint8_t* b_base = desc_ptr(c, bd.addr, canon_span(bd, id.N), "tcgen05.mma B");
  const uint64_t ha = a_from_tmem ? 0 : hash_operand(a_base, ad, 128);
  const uint64_t hb = hash_operand(b_base, bd, id.N);
  const int N = id.N;
  const bool acc = accumulate != 0;
  AsyncOp op;
  op.what = "tcgen05.mma";
  op.d_col0 = d_col; op.d_col1 = d_col + N;
  if (a_from_tmem) { op.a_col0 = a_col; op.a_col1 = a_col + 8; }
  op.run = [=]() {
    if (!a_from_tmem && hash_operand(a_base, ad, 128) != ha) fail("race: the shared-memory A operand of a tcgen05.mma changed between issue and execution");
    if (hash_operand(b_base, bd, N) != hb) fail("race: the shared-memory B operand of a tcgen05.mma changed between issue and execution");
    mma_execute_cg1(c, d_col, a_from_tmem, a_col, a_base, ad, b_base, bd, N, acc);
  };
  submit_async(c, f->tid, std::move(op));
  preempt_point();
}
// cta_group::2: D_v[128 x N] (+)= A_v[128 x 16] * [B_0; B_1][N x 16]^T in each CTA v of the pair; CTA v holds rows
// [v*N/2, (v+1)*N/2) of B at the descriptor's address in ITS shared memory (cute MMA_Traits<SM100_MMA_F16BF16_2x1SM_SS>:
// ALayout / BLayout / CLayout split M, N and M across the two CTAs).
inline void mma_issue_cg2(uint32_t d_tmem, bool a_from_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  Fiber* f = cur();
  Launch& l = L();
  if (l.ctas.size() != 2) fail("tcgen05.mma.cta_group::2 needs a 2-CTA cluster (have %zu)", l.ctas.size());
  if (f->cta->rank != 0) fail("tcgen05.mma.cta_group::2 issued by CTA rank %d (only the leader, rank 0, may issue)", f->cta->rank);
  const InstrDesc id = decode_idesc(idesc, 2);
  if (d_tmem >> 16) fail("tcgen05.mma: D address names lane %u", d_tmem >> 16);
  const int d_col = (int)(d_tmem & 0xFFFF), N = id.N, NH = id.N / 2;
  const SmemDesc ad = a_from_tmem ? SmemDesc{} : decode_desc(a_desc, "tcgen05.mma A"), bd = decode_desc(b_desc, "tcgen05.mma B");
  const int a_col = a_from_tmem ? (int)(a_desc & 0xFFFF) : -1;
  if (a_from_tmem) {
    if ((uint32_t)a_desc >> 16) fail("tcgen05.mma: A (TMEM) address names lane %u", (uint32_t)a_desc >> 16);
    if (a_col < d_col + N && d_col < a_col + 8) fail("tcgen05.mma: A (TMEM cols %d..%d) overlaps D (%d..%d)", a_col, a_col + 7, d_col, d_col + N - 1);
  }
  uint8_t* a_base[2];
  uint8_t* b_base[2];
  uint64_t ha[2], hb[2];
  for (int v = 0; v < 2; ++v) {
    Cta* c = l.ctas[v];
    if (d_col + N > c->tmem_allocated) fail("tcgen05.mma: D columns %d..%d outside CTA %d's allocation", d_col, d_col + N - 1, v);
    a_base[v] = a_from_tmem ? nullptr : desc_ptr(c, ad.addr, canon_span(ad, 128), "tcgen05.mma A");
    if (a_from_tmem && a_col + 8 > c->tmem_allocated) fail("tcgen05.mma: A columns outside CTA %d's allocation", v);
    b_base[v] = desc_ptr(c, bd.addr, canon_span(bd, NH), "tcgen05.mma B");
    ha[v] = a_from_tmem ? 0 : hash_operand(a_base[v], ad, 128);
    hb[v] = hash_operand(b_base[v], bd, NH);
  }
  Cta* c0 = l.ctas[0];
  Cta* c1 = l.ctas[1];
  const bool acc = accumulate != 0;
  AsyncOp op;
  op.what = "tcgen05.mma.cta_group::2";
  op.d_col0 = d_col; op.d_col1 = d_col + N;
  if (a_from_tmem) { op.a_col0 = a_col; op.a_col1 = a_col + 8; }
  op.both_ctas = true;
  op.run = [=]() {
    Cta* cs[2] = {c0, c1};
    for (int v = 0; v < 2; ++v) {
      if (!a_from_tmem && hash_operand(a_base[v], ad, 128) != ha[v]) fail("race: CTA %d's A operand of a cta_group::2 MMA changed between issue and execution", v);
      if (hash_operand(b_base[v], bd, NH) != hb[v]) fail("race: CTA %d's half of the B operand of a cta_group::2 MMA changed between issue and execution", v);
    }
    std::vector<float> B((size_t)N * 16);
    for (int n = 0; n < N; ++n)
      for (int k = 0; k < 16; ++k) B[(size_t)n * 16 + k] = h2f(canon_elem(b_base[n / NH], bd, n % NH, k));
    for (int v = 0; v < 2; ++v)
      for (int r = 0; r < 128; ++r) {
        float a[16];
        for (int k = 0; k < 16; ++k) {
          if (a_from_tmem) {
            const uint32_t w = cs[v]->T(r, a_col + k / 2);
            a[k] = h2f((uint16_t)((k & 1) ? (w >> 16) : (w & 0xFFFF)));
          } else {
            a[k] = h2f(canon_elem(a_base[v], ad, r, k));
          }
        }
        for (int n = 0; n < N; ++n) {
          uint32_t& dw = cs[v]->T(r, d_col + n);
          float accv = 0.f;
          if (acc) memcpy(&accv, &dw, 4);
          const float* b = &B[(size_t)n * 16];
          for (int k = 0; k < 16; ++k) accv += a[k] * b[k];
          memcpy(&dw, &accv, 4);
        }
      }
  };
  submit_async(f->cta, f->tid, std::move(op));
  preempt_point();
}
// tcgen05.commit: the arrive happens once every MMA this thread issued before has executed
inline void commit_arrive(std::vector<void*> bars) {
  Fiber* f = cur();
  AsyncOp op;
  op.what = "tcgen05.commit";
  op.run = [bars]() { for (void* b : bars) mbar_do_arrive(b); };
  submit_async(f->cta, f->tid, std::move(op));
  preempt_point();
}
// bulk copy global -> shared (+ complete_tx); dsts/bars: one per destination CTA
inline void bulk_copy(std::vector<std::pair<void*, void*>> dst_bar, const void* src, uint32_t bytes) {
  Fiber* f = cur();
  if ((bytes & 15u) || ((uintptr_t)src & 15u)) fail("cp.async.bulk: size %u / source %p not 16-byte aligned", bytes, src);
  for (auto& db : dst_bar) if ((uintptr_t)db.first & 15u) fail("cp.async.bulk: destination not 16-byte aligned");
  AsyncOp op;
  op.what = "cp.async.bulk";
  op.run = [dst_bar, src, bytes]() {
    for (auto& db : dst_bar) { memcpy(db.first, src, bytes); mbar_do_complete_tx(db.second, bytes); }
  };
  submit_async(f->cta, -1 - (int)(L().rng.below(4)), std::move(op));   // 4 independent copy queues: completion order is not issue order
  preempt_point();
}

}  // namespace emu

// ================================================================================================ CUDA surface
#undef __shared__
#define __shared__ static
#undef __launch_bounds__
#undef __grid_constant__
#define __grid_constant__
#define __launch_bounds__(...)
#define C3D_DYN_SMEM_ALIGNED(type, name, al) C3D_DYN_SMEM(type, name)
#define C3D_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::cur()->cta->smem)
#define threadIdx (emu::cur()->thread_idx)
#define blockIdx (emu::cur()->cta->block_idx)
#define blockDim (emu::L().block)
#define gridDim (emu::L().grid)
using std::max;
using std::min;

inline void __syncthreads() { emu::syncthreads(); }
inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::syncwarp(mask); }
inline int __all_sync(unsigned mask, int pred) { return emu::all_sync(mask, pred); }
inline int __any_sync(unsigned mask, int pred) { return emu::any_sync(mask, pred); }
inline unsigned __ballot_sync(unsigned mask, int pred) { return emu::ballot_sync(mask, pred); }
// warp shuffles with the CUDA `width` semantics (segments of `width` lanes; out-of-segment sources return the own value)
template <class T> inline T c3d_emu_shfl_from(unsigned mask, T v, int src_lane) {
  static_assert(sizeof(T) == 4, "32-bit shuffles only");
  uint32_t u; memcpy(&u, &v, 4); u = emu::shfl_idx(mask, u, src_lane); memcpy(&v, &u, 4); return v;
}
template <class T> inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
  const int lane = emu::cur()->lane;
  return c3d_emu_shfl_from(mask, v, (lane & ~(width - 1)) | (src & (width - 1)));
}
template <class T> inline T __shfl_xor_sync(unsigned mask, T v, int x, int width = 32) {
  const int lane = emu::cur()->lane, t = lane ^ x;
  return c3d_emu_shfl_from(mask, v, (t & ~(width - 1)) == (lane & ~(width - 1)) ? t : lane);
}
template <class T> inline T __shfl_down_sync(unsigned mask, T v, int d, int width = 32) {
  const int lane = emu::cur()->lane;
  return c3d_emu_shfl_from(mask, v, (lane & (width - 1)) + d < width ? lane + d : lane);
}
template <class T> inline T __shfl_up_sync(unsigned mask, T v, int d, int width = 32) {
  const int lane = emu::cur()->lane;
  return c3d_emu_shfl_from(mask, v, (lane & (width - 1)) >= d ? lane - d : lane);
}
inline void __nanosleep(unsigned) { emu::yield(); }
// atomics: fibers are cooperative (switches only at preempt points), so plain read-modify-writes are atomic; a preempt point
// after each one lets the schedules explore the interleavings around it
inline int atomicCAS(int* p, int cmp, int val) { const int old = *p; if (old == cmp) *p = val; emu::preempt_point(); return old; }
inline int atomicExch(int* p, int val) { const int old = *p; *p = val; emu::preempt_point(); return old; }
inline int atomicAdd(int* p, int v) { const int old = *p; *p = old + v; emu::preempt_point(); return old; }
inline void __trap() { emu::fail("__trap()"); }
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldcs(const T* p) { return *p; }
template <class T> inline void __stcs(T* p, T v) { *p = v; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline unsigned __float2uint_rz(float a) { return a != a || a <= 0.f ? 0u : (a >= 4294967040.f ? 0xFFFFFFFFu : (unsigned)a); }
inline float __fadd_rz(float a, float b) {      // exact in double for the magnitudes used here, then truncate toward zero
  const double d = (double)a + (double)b;
  float f = (float)d;
  if (fabs((double)f) > fabs(d)) f = nextafterf(f, 0.f);
  return f;
}
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
  const unsigned long long v = ((unsigned long long)y << 32) | x;
  unsigned r = 0;
  for (int i = 0; i < 4; ++i) r |= (unsigned)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xFF) << (8 * i);
  return r;
}
inline float __fsqrt_rn(float a) { return sqrtf(a); }
inline float __fdividef(float a, float b) { return a / b; }
inline float __sinf(float a) { return sinf(a); }
inline float __cosf(float a) { return cosf(a); }
inline float __expf(float a) { return expf(a); }
inline float rsqrtf(float a) { return 1.f / sqrtf(a); }
inline float __frcp_rn(float a) { return 1.f / a; }
inline long long clock64() { return (long long)++emu::G().timer; }

// ---- host runtime stubs (the emulated library never links libcudart)
#define C3D_EMU_STUB static inline
C3D_EMU_STUB cudaError_t emu_cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
C3D_EMU_STUB cudaError_t emu_cudaGetLastError() { const int e = emu::G().last_error; emu::G().last_error = 0; return e ? cudaErrorLaunchFailure : cudaSuccess; }
C3D_EMU_STUB const char* emu_cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated kernel failed (see stderr)"; }
C3D_EMU_STUB cudaError_t emu_cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int) {
  *v = a == cudaDevAttrComputeCapabilityMajor ? 10 : (a == cudaDevAttrComputeCapabilityMinor ? 0 : emu::config().sms);
  return cudaSuccess;
}
template <class K> C3D_EMU_STUB cudaError_t emu_cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
C3D_EMU_STUB cudaError_t emu_cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
C3D_EMU_STUB cudaError_t emu_cudaDeviceSynchronize() { return cudaSuccess; }
#define cudaGetDevice emu_cudaGetDevice
#define cudaGetLastError emu_cudaGetLastError
#define cudaGetErrorString emu_cudaGetErrorString
#define cudaDeviceGetAttribute emu_cudaDeviceGetAttribute
#define cudaFuncSetAttribute emu_cudaFuncSetAttribute
#define cudaMemsetAsync emu_cudaMemsetAsync
#define cudaDeviceSynchronize emu_cudaDeviceSynchronize

// kernel<<<grid, block, smem, stream>>>(args...)
#define C3D_LAUNCH(kernel, grid, block, smem, stream, ...) \
  emu::launch(#kernel, dim3(grid), dim3(block), (size_t)(smem), 1, [&]() { kernel(__VA_ARGS__); })
// cudaLaunchKernelEx with a cluster dimension
#define C3D_LAUNCH_CLUSTER(kernel, grid, block, smem, stream, cluster, ...) \
  (emu::launch(#kernel, dim3(grid), dim3(block), (size_t)(smem), (cluster), [&]() { kernel(__VA_ARGS__); }) ? (emu::G().last_error = 0, cudaErrorLaunchFailure) : cudaSuccess)
