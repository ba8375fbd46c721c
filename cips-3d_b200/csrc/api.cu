// extern "C" entry points of libcips3d_b200.so (see include/cips3d_b200.h) + error plumbing.
#include <stdarg.h>
#include <string.h>

#include "c3d_common.cuh"

#include <atomic>
static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};
void c3d_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" unsigned long long c3d_launch_count(void) { return g_launches.load(); }

void c3d_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// implemented in simt_pipeline.cu / ray_siren_tc.cu / cips_tc.cu
size_t c3d_ray_siren_simt_workspace_bytes(const C3dRayParams* p);
int c3d_ray_siren_fwd_simt(const C3dRayParams*, const C3dSirenWeights*, const C3dRayIO*, void*, size_t, cudaStream_t);
size_t c3d_cips_simt_workspace_bytes(const C3dCipsParams* p);
int c3d_cips_fwd_simt(const C3dCipsParams*, const C3dCipsWeights*, const float*, float*, float*, void*, size_t, cudaStream_t);
size_t c3d_ray_siren_tc_workspace_bytes(const C3dRayParams* p);
int c3d_ray_siren_fwd_tc(const C3dRayParams*, const C3dSirenWeights*, const C3dRayIO*, void*, size_t, cudaStream_t);
size_t c3d_cips_tc_workspace_bytes(const C3dCipsParams* p);
int c3d_cips_fwd_tc(const C3dCipsParams*, const C3dCipsWeights*, const float*, float*, float*, void*, size_t, cudaStream_t, void* acts_f16, void* zsign_u16);

// ---- kernel-variant options: one process-wide snapshot of the environment
#include <mutex>
#include <stdlib.h>
static C3dOptions g_opts;
static std::once_flag g_opts_once;
static void load_options() {
  C3dOptions o;
  auto geti = [](const char* n, int dflt) { const char* e = getenv(n); return e && *e ? atoi(e) : dflt; };
  o.cips_cluster = geti("C3D_CIPS_CLUSTER", 1);
  if (o.cips_cluster != 1 && o.cips_cluster != 2 && o.cips_cluster != 4) o.cips_cluster = 1;
  o.cips_pair = geti("C3D_CIPS_PAIR", 1) != 0;
  o.cips_stagger_ns = geti("C3D_CIPS_STAGGER_NS", 0);
  o.cips_res16 = geti("C3D_CIPS_RES16", 1);
  const char* b = getenv("C3D_BLUR");
  o.blur_impl = !b || !*b ? 2 : (b[0] == 't' && b[1] == 'i' ? 0 : (b[0] == 't' ? 1 : 2));      // tile | tma | stream
  if (geti("C3D_BLUR_TMA", 0) != 0) o.blur_impl = 1;                                             // round-1 spelling
  const char* pg = getenv("C3D_PIGAN_IMPL");
  o.pigan_tc = !(pg && (pg[0] == 's' || pg[0] == 'S'));
  o.pigan_pair = geti("C3D_PIGAN_PAIR", 0) != 0;
  const char* rm = getenv("C3D_RAY_MATH");
  o.ray_math = !rm || !*rm ? C3D_RAY_MATH_DEFAULT : (rm[0] == 'f' ? 2 : (rm[0] == 'w' ? 1 : 0));
  g_opts = o;
}
const C3dOptions& c3d_options() {
  std::call_once(g_opts_once, load_options);
  return g_opts;
}
// host layer / tests: re-read the environment (not thread-safe against concurrent launches; an A/B hook, not an API)
extern "C" void c3d_reload_options(void) {
  std::call_once(g_opts_once, [] {});
  load_options();
}

extern "C" int c3d_version(void) { return 100; }
extern "C" const char* c3d_last_error(void) { return g_err; }

// cudaGetDeviceProperties costs milliseconds: query each device once (attributes only) and cache.
static std::atomic<int> g_dev_cc[64];     // 0 = unknown, else major*10+minor
static std::atomic<int> g_dev_sms[64];
static void query_device(int dev) {
  int major = 0, minor = 0, sms = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  g_dev_sms[dev].store(sms);
  g_dev_cc[dev].store(major * 10 + minor + 1000);
}
extern "C" int c3d_device_supported(int dev) {
  if (dev < 0 || dev >= 64) return 0;
  if (g_dev_cc[dev].load() == 0) query_device(dev);
  return (g_dev_cc[dev].load() - 1000) / 10 == 10 ? 1 : 0;
}
int c3d_device_sm_count(int dev) {
#ifdef C3D_EMU
  return emu::config().sms;      // tests change the emulated SM count between calls
#endif
  if (dev < 0 || dev >= 64) return 148;
  if (g_dev_cc[dev].load() == 0) query_device(dev);
  return g_dev_sms[dev].load();
}

static int check_ray_args(const C3dRayParams* p, const C3dSirenWeights* w, const C3dRayIO* io) {
  C3D_CHECK_ARG(p && w && io, "ray_siren: null struct pointer");
  C3D_CHECK_ARG(p->batch >= 0 && p->img_size >= 1 && p->n_rays >= 0, "ray_siren: bad sizes");
  C3D_CHECK_ARG(p->num_steps >= 3 && p->num_steps <= 32, "ray_siren: num_steps must be in [3,32], got %d", p->num_steps);
  C3D_CHECK_ARG((long long)p->n_rays <= (long long)p->img_size * p->img_size, "ray_siren: n_rays > img_size^2");
  C3D_CHECK_ARG(io->ray_idx || (long long)p->ray_offset + p->n_rays <= (long long)p->img_size * p->img_size,
                "ray_siren: ray_offset + n_rays exceeds the image");
  C3D_CHECK_ARG(p->clamp_mode == 0 || p->clamp_mode == 1, "ray_siren: clamp_mode must be 0 (relu) or 1 (softplus)");
  C3D_CHECK_ARG(io->cam2world && io->jitter_u && io->pixels_fea, "ray_siren: null cam2world/jitter_u/pixels_fea");
  C3D_CHECK_ARG(!p->hierarchical || io->pdf_u, "ray_siren: hierarchical sampling needs pdf_u");
  C3D_CHECK_ARG(w->w0 && w->b0 && w->w1 && w->b1 && w->w_sigma && w->b_sigma && w->wc && w->bc && w->wl && w->bl,
                "ray_siren: null weight pointer");
  C3D_CHECK_ARG(w->gamma0 && w->beta0 && w->gamma1 && w->beta1 && w->gammac && w->betac, "ray_siren: null FiLM pointer");
  return C3D_OK;
}

extern "C" size_t c3d_ray_siren_workspace_bytes(const C3dRayParams* p) {
  if (!p) return 0;
  return p->impl == C3D_IMPL_SIMT ? c3d_ray_siren_simt_workspace_bytes(p) : c3d_ray_siren_tc_workspace_bytes(p);
}

extern "C" int c3d_ray_siren_fwd(const C3dRayParams* p, const C3dSirenWeights* w, const C3dRayIO* io,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  if (int e = check_ray_args(p, w, io)) return e;
  if (p->batch == 0 || p->n_rays == 0) return C3D_OK;
  C3D_CHECK_ARG(workspace || c3d_ray_siren_workspace_bytes(p) == 0, "ray_siren: null workspace");
  if (p->impl == C3D_IMPL_SIMT) return c3d_ray_siren_fwd_simt(p, w, io, workspace, workspace_bytes, (cudaStream_t)stream);
  if (p->impl == C3D_IMPL_TC) return c3d_ray_siren_fwd_tc(p, w, io, workspace, workspace_bytes, (cudaStream_t)stream);
  c3d_set_error("ray_siren: unknown impl %d", p->impl);
  return C3D_EINVAL;
}

static int check_cips_args(const C3dCipsParams* p, const C3dCipsWeights* w, const float* x, float* rgb) {
  C3D_CHECK_ARG(p && w && x && rgb, "cips: null pointer");
  C3D_CHECK_ARG(p->n_blocks >= 1 && 2 * p->n_blocks <= C3D_CIPS_MAX_LAYERS, "cips: n_blocks must be in [1,9]");
  C3D_CHECK_ARG(p->in_dim >= 1 && p->hidden >= 1 && p->batch >= 0 && p->n_pix >= 0, "cips: bad sizes");
  for (int l = 0; l < 2 * p->n_blocks; ++l)
    C3D_CHECK_ARG(w->w[l] && w->style1p[l] && w->demod[l], "cips: null weight/style/demod for layer %d", l);
  for (int b = p->rgb_from; b < p->n_blocks; ++b)
    C3D_CHECK_ARG(b < 0 || (w->rgb_w[b] && w->rgb_b[b]), "cips: null ToRGB weights for block %d", b);
  return C3D_OK;
}

extern "C" size_t c3d_cips_workspace_bytes(const C3dCipsParams* p) {
  if (!p) return 0;
  return p->impl == C3D_IMPL_SIMT ? c3d_cips_simt_workspace_bytes(p) : c3d_cips_tc_workspace_bytes(p);
}

extern "C" int c3d_cips_fwd(const C3dCipsParams* p, const C3dCipsWeights* w, const float* x, float* rgb,
                            float* hidden_out, void* workspace, size_t workspace_bytes, void* stream) {
  if (int e = check_cips_args(p, w, x, rgb)) return e;
  if (p->batch == 0 || p->n_pix == 0) return C3D_OK;
  C3D_CHECK_ARG(workspace || c3d_cips_workspace_bytes(p) == 0, "cips: null workspace");
  if (p->impl == C3D_IMPL_SIMT) return c3d_cips_fwd_simt(p, w, x, rgb, hidden_out, workspace, workspace_bytes, (cudaStream_t)stream);
  if (p->impl == C3D_IMPL_TC) return c3d_cips_fwd_tc(p, w, x, rgb, hidden_out, workspace, workspace_bytes, (cudaStream_t)stream, nullptr, nullptr);
  c3d_set_error("cips: unknown impl %d", p->impl);
  return C3D_EINVAL;
}

// training forward: c3d_cips_fwd + the activation stash the backward pass needs (see include/cips3d_b200.h)
extern "C" int c3d_cips_fwd_train(const C3dCipsParams* p, const C3dCipsWeights* w, const float* x, float* rgb, void* acts_f16,
                                  void* zsign_u16, void* workspace, size_t workspace_bytes, void* stream) {
  if (int e = check_cips_args(p, w, x, rgb)) return e;
  C3D_CHECK_ARG(acts_f16 && zsign_u16, "cips_fwd_train: null activation / sign stash");
  C3D_CHECK_ARG(p->impl == C3D_IMPL_TC, "cips_fwd_train: tensor-core kernels only");
  if (p->batch == 0 || p->n_pix == 0) return C3D_OK;
  C3D_CHECK_ARG(workspace, "cips_fwd_train: null workspace");
  return c3d_cips_fwd_tc(p, w, x, rgb, nullptr, workspace, workspace_bytes, (cudaStream_t)stream, acts_f16, zsign_u16);
}
