// render.cu -- device-driven inference rounds (row f-3 of SURVEY.md section 8): the loop of NeRFRenderer.run_cuda /
// run_cuda_distill (nerf/renderer.py:335-387, 425-470)
//     while step < max_steps and n_alive > 0:
//         n_step = clamp(n_rays // n_alive, 1, 8)
//         march_rays -> network -> composite_rays -> rays_alive = rays_alive[rays_alive >= 0]; step += n_step
// without a device->host synchronisation per round.  The round geometry (n_alive, n_step, padded row count) lives in a
// device control block that the compaction kernel of round r rewrites for round r+1; every kernel of a round reads it.
// n_alive * n_step never exceeds n_rays, so one set of sample buffers of n_rays + 128 rows serves every round.  The host
// queues K rounds per call and looks at the control block only between calls (rounds after the last live one are no-ops).
#include "render_core.cuh"

using namespace lnrf;

// rows the sample buffers hold beyond the 128-row padding (the row budget of a round: n_alive * n_step never exceeds it)
static uint32_t sample_row_budget(const lnrf_render_desc* d) {
    return d->sample_rows > d->n_rays + 128u ? d->sample_rows - 128u : d->n_rays;
}

extern "C" {

// [alive-ray compaction look-back words | compact rounds: marcher look-back words + (offset, count) per alive ray]
size_t lnrf_render_scratch_bytes(uint32_t n_rays) { return lnrf_compact_alive_scratch_bytes(n_rays) + march_compact_scratch_bytes(n_rays); }

int lnrf_render_begin(const lnrf_render_desc* d, lnrf_stream_t stream) {
    LNRF_REQUIRE(d && d->ctl, "render_begin: null descriptor / control block");
    return render_begin_launch(d->ctl, d->n_rays, d->max_steps, sample_row_budget(d), d->sample_rows > d->n_rays + 128u && d->samples_per_round ? d->samples_per_round : 8u, d->rays_alive[0], d->rays_t, d->nears, d->weights_sum, d->depth, d->image,
                               d->weights_edit_sum, d->depth_edit, d->ray_steps, d->ray_flags, d->nstep_seq, d->nstep_len,
                               reinterpret_cast<cudaStream_t>(stream));
}

int lnrf_render_rounds(const lnrf_render_desc* d, uint32_t first_round, uint32_t n_rounds, lnrf_stream_t stream) {
    LNRF_REQUIRE(d && d->ctl, "render_rounds: null descriptor / control block");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool distill = d->edit_bitfield != nullptr;
    LNRF_REQUIRE(!distill || (d->edit_occ && d->weights_edit_sum && d->depth_edit), "render_rounds: distillation needs edit_occ / weights_edit_sum / depth_edit");
    LNRF_REQUIRE(d->rays_alive[0] && d->rays_alive[1] && d->enc_f16 && d->sigmas && d->rgbs, "render_rounds: null buffer");
    const uint32_t cap = d->n_rays;                  // ray capacity (march groups, composite, compaction)
    const uint32_t row_cap = sample_row_budget(d);   // sample-row capacity (encoder, network)
    // fast schedule (more than 8 samples per ray per round, own row budget, no prescribed sequence): rounds >= 1 are COMPACT -- a
    // ray's samples sit right behind the previous ray's (raymarch.cu k_march_infer_compact), the round's row count is ctl[15]
    const bool compact_rounds = d->sample_rows > d->n_rays + 128u && d->samples_per_round > 8u && !d->nstep_seq;
    const size_t ca_bytes = lnrf_compact_alive_scratch_bytes(cap);
    LNRF_REQUIRE(!compact_rounds || (d->scratch && d->scratch_bytes >= ca_bytes + march_compact_scratch_bytes(cap)),
                 "render_rounds: scratch too small for compact rounds (lnrf_render_scratch_bytes)");
    void* scratch_m = compact_rounds ? static_cast<void*>(static_cast<uint8_t*>(d->scratch) + ca_bytes) : nullptr;
    for (uint32_t r = first_round; r < first_round + n_rounds; r++) {
        int32_t* cur = d->rays_alive[r & 1u];
        int32_t* nxt = d->rays_alive[(r + 1u) & 1u];
        const bool compact = compact_rounds && r > 0;
        const int32_t* rows_dev = d->ctl + (compact ? 15 : 3);  // kCtlCompactRows : kCtlRows
        const float* noises = (r == 0 && d->first_round_noises) ? d->first_round_noises : nullptr;  // perturb only on the first round
        if (compact) {
            if (int e = march_infer_compact_dev_launch(distill, d->ctl, cap, cur, d->rays_t, d->rays_o, d->rays_d, d->bound, d->dt_gamma,
                                                       d->max_steps, d->cascade, d->grid_size, d->density_bitfield, d->edit_bitfield, d->fars,
                                                       d->xyzs, d->dirs, d->deltas, d->edit_occ, scratch_m, d->occupied_box, st))
                return e;
        } else if (int e = march_infer_dev_launch(distill, d->ctl, cap, cur, d->rays_t, d->rays_o, d->rays_d, d->bound, d->dt_gamma, d->max_steps,
                                                  d->cascade, d->grid_size, d->density_bitfield, d->edit_bitfield, d->fars, d->xyzs, d->dirs,
                                                  d->deltas, d->edit_occ, noises, /*first=*/r == 0, d->occupied_box, st))
            return e;
        if (int e = lnrf_grid_encode_forward_world(d->xyzs, d->bound, d->embeddings_f16, d->offsets_host, d->enc_f16, row_cap + 128u, rows_dev,
                                                   d->num_levels, d->level_scale_log2, d->base_resolution, d->gridtype, d->align_corners,
                                                   d->interpolation, LNRF_F16, stream))
            return e;
        if (int e = nerf_forward_dev_launch(d->enc_f16, d->dirs, d->w_sigma_f16, d->w_color_f16, row_cap + 128u, rows_dev, d->num_layers_sigma,
                                            d->num_layers_color, d->density_scale, d->sigmas, d->rgbs, st))
            return e;
        if (compact) {
            if (int e = composite_infer_compact_dev_launch(distill, d->ctl, cap, d->T_thresh, cur, d->rays_t, d->sigmas, d->rgbs, d->deltas,
                                                           d->weights_sum, d->weights_edit_sum, d->depth, d->depth_edit, d->edit_occ, d->image,
                                                           scratch_m, st))
                return e;
        } else if (int e = composite_infer_dev_launch(distill, d->ctl, cap, d->T_thresh, cur, d->rays_t, d->sigmas, d->rgbs, d->deltas, d->weights_sum,
                                                      d->weights_edit_sum, d->depth, d->depth_edit, d->edit_occ, d->image, st))
            return e;
        if (int e = compact_alive_dev_launch(d->ctl, cap, cur, nxt, d->scratch, d->scratch_bytes, st)) return e;
    }
    return LNRF_OK;
}

}  // extern "C"
