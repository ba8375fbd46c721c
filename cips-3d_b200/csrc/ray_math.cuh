// Per-ray arithmetic of the volumetric renderer, shared by the fused tcgen05 kernel and the
// unfused SIMT cross-check path.  Each function cites the reference lines it restates.
#pragma once
#include "c3d_common.cuh"

namespace c3d {

constexpr int kMaxS = 32;       // coarse samples per ray
constexpr int kMaxNS = 2 * kMaxS;
constexpr int kFeat = 32;       // rgb_dim of the shipping NeRF (ffhq_exp.yaml:55)
constexpr int kOutC = kFeat + 1;  // [feature(32), sigma]

// torch.linspace(start, end, steps)[idx] in fp32 (ATen RangeFactoriesKernel: symmetric form)
__device__ __forceinline__ float linspace_f32(float start, float end, int steps, int idx) {
  float step = __fdiv_rn(end - start, (float)(steps - 1));
  return idx < steps / 2 ? __fadd_rn(start, __fmul_rn(step, (float)idx))
                         : __fsub_rn(end, __fmul_rn(step, (float)(steps - idx - 1)));
}

// Camera-space unit direction of global ray `ray` = h*R + w.
// get_initial_rays_trig, exp/comm/comm_utils.py:392-398.
__device__ __forceinline__ void ray_dir_cam(int ray, int R, float z_cam, float& dx, float& dy,
                                            float& dz) {
  int h = ray / R, w = ray - h * R;
  float x = linspace_f32(-1.f, 1.f, R, w);
  float y = linspace_f32(1.f, -1.f, R, h);
  float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z_cam, z_cam)));
  dx = __fdiv_rn(x, nrm);
  dy = __fdiv_rn(y, nrm);
  dz = __fdiv_rn(z_cam, nrm);
}

struct RayFrame {  // per-ray constants in world space
  float ox, oy, oz;     // camera origin  (cam2world[:, :3, 3], comm_utils.py:666-677)
  float dwx, dwy, dwz;  // world direction R*d_cam (comm_utils.py:658-664)
  float dcx, dcy, dcz;  // camera-space direction
};

__device__ __forceinline__ RayFrame make_ray_frame(const float* __restrict__ M /*16*/, int ray,
                                                   int R, float z_cam) {
  RayFrame f;
  ray_dir_cam(ray, R, z_cam, f.dcx, f.dcy, f.dcz);
  f.dwx = M[0] * f.dcx + M[1] * f.dcy + M[2] * f.dcz;
  f.dwy = M[4] * f.dcx + M[5] * f.dcy + M[6] * f.dcz;
  f.dwz = M[8] * f.dcx + M[9] * f.dcy + M[10] * f.dcz;
  f.ox = M[3];
  f.oy = M[7];
  f.oz = M[11];
  return f;
}

// Coarse sample s of a ray: jittered depth and world position.
// perturb_points comm_utils.py:430-437 + transform_sampled_points :649-656.
__device__ __forceinline__ void coarse_sample(const RayFrame& f, const float* __restrict__ M,
                                              float ray_start, float ray_end, int S, int s,
                                              float u, float& z, float& px, float& py,
                                              float& pz) {
  float z0 = linspace_f32(ray_start, ray_end, S, 0);
  float z1 = linspace_f32(ray_start, ray_end, S, 1);
  float zs = linspace_f32(ray_start, ray_end, S, s);
  float off = __fmul_rn(u - 0.5f, z1 - z0);
  z = __fadd_rn(zs, off);
  float cx = __fadd_rn(__fmul_rn(f.dcx, zs), __fmul_rn(off, f.dcx));
  float cy = __fadd_rn(__fmul_rn(f.dcy, zs), __fmul_rn(off, f.dcy));
  float cz = __fadd_rn(__fmul_rn(f.dcz, zs), __fmul_rn(off, f.dcz));
  px = M[0] * cx + M[1] * cy + M[2] * cz + M[3];
  py = M[4] * cx + M[5] * cy + M[6] * cz + M[7];
  pz = M[8] * cx + M[9] * cy + M[10] * cz + M[11];
}

// Fine sample: origin + dir * z  (generator_nerf_inr.py:586-588)
__device__ __forceinline__ void fine_sample(const RayFrame& f, float z, float& px, float& py,
                                            float& pz) {
  px = __fadd_rn(f.ox, __fmul_rn(f.dwx, z));
  py = __fadd_rn(f.oy, __fmul_rn(f.dwy, z));
  pz = __fadd_rn(f.oz, __fmul_rn(f.dwz, z));
}

__device__ __forceinline__ float softplus_f32(float x) {  // F.softplus, beta 1, threshold 20
  return x > 20.f ? x : log1pf(expf(x));
}

// fancy_integration weights, exp/pigan/pigan_utils.py:241-257.
//   delta_i = z_{i+1}-z_i (last 1e10); alpha = 1-exp(-delta*clamp(sigma+noise));
//   T_i = prod_{j<i}(1-alpha_j+1e-10); w_i = alpha_i*T_i.  Returns sum_i w_i.
template <typename ZF, typename SF, typename NF, typename WF>
__device__ __forceinline__ float integrate_weights(int n, int clamp_mode, ZF z_at, SF sigma_at,
                                                   NF noise_at, WF w_out) {
  float T = 1.f, wsum = 0.f;
  float zc = z_at(0);
  for (int i = 0; i < n; ++i) {
    float zn = (i + 1 < n) ? z_at(i + 1) : 0.f;
    float delta = (i + 1 < n) ? __fsub_rn(zn, zc) : 1e10f;
    float s = __fadd_rn(sigma_at(i), noise_at(i));
    float a = clamp_mode == 1 ? softplus_f32(s) : fmaxf(s, 0.f);
    float alpha = __fsub_rn(1.f, expf(__fmul_rn(-delta, a)));
    float w = __fmul_rn(alpha, T);
    w_out(i, w);
    wsum += w;
    T = __fmul_rn(T, __fadd_rn(__fsub_rn(1.f, alpha), 1e-10f));
    zc = zn;
  }
  return wsum;
}

// sample_pdf for one ray, exp/pigan/pigan_utils.py:164-209 called from
// generator_nerf_inr.py:570-581 with weights = (w + 1e-5)[1:-1], bins = midpoints of z.
// w: S coarse weights, z: S coarse depths, u: S uniforms -> fz: S fine depths (unsorted).
__device__ inline void sample_pdf_ray(int S, const float* w, const float* z, const float* u,
                                      float* fz) {
  const int ns = S - 2;
  float cdf[kMaxS];  // ns+1 entries
  float sum = 0.f;
  for (int j = 0; j < ns; ++j) sum += __fadd_rn(__fadd_rn(w[j + 1], 1e-5f), 1e-5f);
  cdf[0] = 0.f;
  for (int j = 0; j < ns; ++j) {
    float pdf = __fdiv_rn(__fadd_rn(__fadd_rn(w[j + 1], 1e-5f), 1e-5f), sum);
    cdf[j + 1] = __fadd_rn(cdf[j], pdf);
  }
  for (int k = 0; k < S; ++k) {
    float uk = u[k];
    int i = 0;  // searchsorted(cdf, u, right=False): first i with cdf[i] >= u
    while (i <= ns && cdf[i] < uk) ++i;
    int below = max(i - 1, 0), above = min(i, ns);
    float cb = cdf[below], ca = cdf[above];
    float bb = 0.5f * __fadd_rn(z[below], z[below + 1]);
    float ba = 0.5f * __fadd_rn(z[above], z[above + 1]);
    float denom = __fsub_rn(ca, cb);
    if (denom < 1e-5f) denom = 1.f;
    fz[k] = __fadd_rn(bb, __fmul_rn(__fdiv_rn(__fsub_rn(uk, cb), denom), __fsub_rn(ba, bb)));
  }
}

// Stable ascending sort of keys with an index payload (n <= 64), insertion sort.
// generator.py:1735: torch.sort over cat([fine, coarse]).
__device__ inline void sort_keys(int n, float* key, int* idx) {
  for (int i = 1; i < n; ++i) {
    float k = key[i];
    int v = idx[i];
    int j = i - 1;
    while (j >= 0 && key[j] > k) {
      key[j + 1] = key[j];
      idx[j + 1] = idx[j];
      --j;
    }
    key[j + 1] = k;
    idx[j + 1] = v;
  }
}

}  // namespace c3d
