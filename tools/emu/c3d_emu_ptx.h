// c3d_emu_ptx.h -- emulated twins of the PTX wrappers of csrc/c3d_common.cuh (same names and signatures).
// Included INSIDE namespace c3d by c3d_common.cuh in -DC3D_EMU builds.  TEST INFRASTRUCTURE ONLY.

inline uint32_t smem_u32(const void* p) { return emu::smem_addr_of(p); }
inline uint32_t lane_id() { return (uint32_t)emu::cur()->lane; }
inline bool elect_one() { return emu::cur()->lane == 0; }     // converged warp: the lowest lane is elected

// ---- mbarrier
inline void mbar_init(uint64_t* bar, uint32_t count) { emu::mbar_do_init(bar, count); }
inline void fence_mbar_init() {}
inline void fence_proxy_async() {}
inline void mbar_arrive(uint64_t* bar) { emu::mbar_do_arrive(bar); emu::preempt_point(); }
inline void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  emu::mbar_do_expect_tx(bar, bytes);
  emu::mbar_do_arrive(bar);
  emu::preempt_point();
}
constexpr uint32_t kSuspendHintNs = 0;
inline bool mbar_try_wait(uint64_t* bar, uint32_t parity) {     // blocks the fiber until the phase completes
  if (parity > 1) emu::fail("mbarrier.try_wait: parity %u", parity);
  emu::mbar_warp_wait(bar, parity);
  return true;
}
inline bool mbar_test(uint64_t* bar, uint32_t parity) {
  if (parity > 1) emu::fail("mbarrier.test_wait: parity %u", parity);
  emu::preempt_point();
  return emu::mbar_warp_test(bar, parity);
}
inline unsigned long long c3d_globaltimer() { return emu::G().timer += 1000; }

// ---- bulk async copy
inline void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  (void)emu::smem_addr_of(smem_dst);    // must be a shared-memory address
  emu::bulk_copy({{smem_dst, (void*)bar}}, gsrc, bytes);
}
inline void bulk_g2s_mc(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  const uint32_t d = emu::smem_addr_of(smem_dst), b = emu::smem_addr_of(bar);
  std::vector<std::pair<void*, void*>> db;
  for (uint32_t r = 0; r < 16; ++r)
    if (mask >> r & 1u)
      db.push_back({emu::smem_ptr_of((d & 0xFFFFFFu) | (r << 24), bytes, "cp.async.bulk.multicast dst"),
                    emu::smem_ptr_of((b & 0xFFFFFFu) | (r << 24), 8, "cp.async.bulk.multicast mbarrier")});
  emu::bulk_copy(db, gsrc, bytes);
}

// ---- tcgen05 / TMEM
template <int kCols>
inline void tmem_alloc(uint32_t* smem_result) { emu::tmem_do_alloc(smem_result, kCols); }
template <int kCols>
inline void tmem_dealloc(uint32_t taddr) { emu::tmem_do_dealloc(taddr, kCols); }
inline void tc_fence_before() {}
inline void tc_fence_after() {}
inline void tc_wait_ld() {}
inline void tc_wait_st() {}
inline void tc_commit(uint64_t* bar) { emu::commit_arrive({(void*)bar}); }
inline void tc_commit_mc(uint64_t* bar, uint16_t mask) {
  const uint32_t b = emu::smem_addr_of(bar);
  std::vector<void*> bars;
  for (uint32_t r = 0; r < 16; ++r)
    if (mask >> r & 1u) bars.push_back(emu::smem_ptr_of((b & 0xFFFFFFu) | (r << 24), 8, "tcgen05.commit.multicast"));
  emu::commit_arrive(bars);
}
inline void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  const uint32_t b = emu::smem_addr_of(bar);
  emu::mbar_do_arrive(emu::smem_ptr_of((b & 0xFFFFFFu) | (rank << 24), 8, "mbarrier.arrive.shared::cluster"));
  emu::preempt_point();
}
inline void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  emu::mma_issue_cg1(d_tmem, false, a_desc, b_desc, idesc, accumulate);
}
inline void umma_ss_w(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate) {
  emu::mma_issue_cg1(d_tmem, false, (uint64_t)a_lo | (uint64_t)desc_hi << 32, (uint64_t)b_lo | (uint64_t)desc_hi << 32, idesc, accumulate);
}
inline void umma_ts_w(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate) {
  emu::mma_issue_cg1(d_tmem, true, a_tmem, (uint64_t)b_lo | (uint64_t)desc_hi << 32, idesc, accumulate);
}
inline void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  emu::mma_issue_cg1(d_tmem, true, a_tmem, b_desc, idesc, accumulate);
}
inline void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) { emu::tmem_ld<8>(taddr, v); }
inline void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) { emu::tmem_ld<16>(taddr, v); }
inline void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) { emu::tmem_ld<32>(taddr, v); }
inline void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) { emu::tmem_st<8>(taddr, v); }
inline void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) { emu::tmem_st<16>(taddr, v); }

// ---- named barriers, register budgets, clusters
inline void named_bar_sync(int id, int nthreads) { emu::bar_sync(id, nthreads); }
template <int kId, int kThreads>
inline void named_bar_sync_c() { emu::bar_sync(kId, kThreads); }
template <int kThreads>
inline void named_bar_sync_n(int id) { emu::bar_sync(id, kThreads); }
template <int N>
inline void reg_dec() { emu::syncwarp(); }
template <int N>
inline void reg_inc() { emu::syncwarp(); }
inline uint32_t cluster_ctarank() { return (uint32_t)emu::cur()->cta->rank; }
inline void cluster_sync_all() { emu::cluster_sync(); }

// ---- tcgen05 cta_group::2
template <int kCols>
inline void tmem_alloc_cg2(uint32_t* smem_result) { emu::tmem_do_alloc(smem_result, kCols); }
template <int kCols>
inline void tmem_dealloc_cg2(uint32_t taddr) { emu::tmem_do_dealloc(taddr, kCols); }
inline void umma_ss_w_cg2(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate) {
  emu::mma_issue_cg2(d_tmem, false, (uint64_t)a_lo | (uint64_t)desc_hi << 32, (uint64_t)b_lo | (uint64_t)desc_hi << 32, idesc, accumulate);
}
inline void umma_ts_w_cg2(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate) {
  emu::mma_issue_cg2(d_tmem, true, a_tmem, (uint64_t)b_lo | (uint64_t)desc_hi << 32, idesc, accumulate);
}
inline void tc_commit_cg2_mc(uint64_t* bar, uint16_t mask) {
  if (emu::cur()->cta->rank != 0) emu::fail("tcgen05.commit.cta_group::2 issued by CTA rank %d", emu::cur()->cta->rank);
  tc_commit_mc(bar, mask);
}
inline bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) { return mbar_try_wait(bar, parity); }
inline void fence_proxy_async_all() {}
