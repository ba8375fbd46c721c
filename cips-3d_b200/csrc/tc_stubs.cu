// Placeholder until the tcgen05 kernels land: the TC entry points fail loudly.
#include "c3d_common.cuh"
size_t c3d_ray_siren_tc_workspace_bytes(const C3dRayParams*) { return 0; }
int c3d_ray_siren_fwd_tc(const C3dRayParams*, const C3dSirenWeights*, const C3dRayIO*, void*, size_t, cudaStream_t) {
  c3d_set_error("ray_siren: tcgen05 kernel not built yet");
  return C3D_EINVAL;
}
size_t c3d_cips_tc_workspace_bytes(const C3dCipsParams*) { return 0; }
int c3d_cips_fwd_tc(const C3dCipsParams*, const C3dCipsWeights*, const float*, float*, float*, void*, size_t, cudaStream_t) {
  c3d_set_error("cips: tcgen05 kernel not built yet");
  return C3D_EINVAL;
}
