// render_core.cuh -- internal launchers of the device-driven inference rounds (row f-3), shared between raymarch.cu,
// nerfnet.cu and render.cu.  `ctl` is the int32[32] device control block documented in raymarch.cu.
#pragma once
#include "common.cuh"

namespace lnrf {

constexpr int kRenderCtlInts = 32;

int render_begin_launch(int32_t* ctl, uint32_t n_rays, uint32_t max_steps, uint32_t row_budget, uint32_t step_cap, int32_t* rays_alive, float* rays_t,
                        const float* nears,
                        float* weights_sum, float* depth, float* image, float* weights_edit_sum, float* depth_edit, int32_t* ray_steps,
                        uint8_t* ray_flags, const int32_t* nstep_seq, uint32_t nstep_len, cudaStream_t st);
// one march over the rays of the current round (grids sized by the ray capacity; geometry read from ctl)
int march_infer_dev_launch(bool distill, const int32_t* ctl, uint32_t n_rays_cap, const int32_t* rays_alive, const float* rays_t,
                           const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t C,
                           uint32_t H, const uint8_t* grid, const uint8_t* edit_grid, const float* fars, float* xyzs, float* dirs,
                           float* deltas, uint8_t* edit_occ, const float* noises, bool first, const float* occ_box, cudaStream_t st);
int composite_infer_dev_launch(bool distill, const int32_t* ctl, uint32_t n_rays_cap, float T_thresh, int32_t* rays_alive, float* rays_t,
                               const float* sigmas, const float* rgbs, const float* deltas, float* weights_sum, float* weights_edit_sum,
                               float* depth, float* depth_edit, const uint8_t* edit_occ, float* image, cudaStream_t st);
// compact rounds of the fast schedule (raymarch.cu: k_march_infer_compact): samples back to back, ctl[15] = rows of the round
size_t march_compact_scratch_bytes(uint32_t n_rays_cap);
int march_infer_compact_dev_launch(bool distill, int32_t* ctl, uint32_t n_rays_cap, const int32_t* rays_alive, const float* rays_t,
                                   const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t C,
                                   uint32_t H, const uint8_t* grid, const uint8_t* edit_grid, const float* fars, float* xyzs, float* dirs,
                                   float* deltas, uint8_t* edit_occ, void* scratch_m, const float* occ_box, cudaStream_t st);
int composite_infer_compact_dev_launch(bool distill, const int32_t* ctl, uint32_t n_rays_cap, float T_thresh, int32_t* rays_alive, float* rays_t,
                                       const float* sigmas, const float* rgbs, const float* deltas, float* weights_sum,
                                       float* weights_edit_sum, float* depth, float* depth_edit, const uint8_t* edit_occ, float* image,
                                       void* scratch_m, cudaStream_t st);
// compacts rays_alive -> out and publishes the next round in ctl
int compact_alive_dev_launch(int32_t* ctl, uint32_t n_rays_cap, const int32_t* rays_alive, int32_t* out, void* scratch,
                             size_t scratch_bytes, cudaStream_t st);
// inference network on ctl[kCtlRows] samples (capacity M_cap rows)
int nerf_forward_dev_launch(const void* enc_f16, const float* dirs, const void* w_sigma_f16, const void* w_color_f16, uint32_t M_cap,
                            const int32_t* M_dev, uint32_t ns, uint32_t nc, float density_scale, float* sigmas, float* rgbs,
                            cudaStream_t st);

}  // namespace lnrf
