// Launch descriptor of the fused front-end kernel (frontend_kernel.cu): packed RGB32 -> I420 planes + adaptive-quant
// statistics + the four lowres planes in one pass.
#pragma once
#include <cuda.h>
#include "csp_kernels.h"

namespace xv {

struct FrontendJob {
    alignas(64) CUtensorMap tmap;          // packed source as (w pixels of 4 bytes, h rows, n frames); filled by launch_frontend
    // stage 1 output: tight I420 planes (device), frame f at + f * dst_frame_bytes
    uint8_t *dst_y, *dst_u, *dst_v; int y_stride, c_stride; size_t dst_frame_bytes;
    // lowres: four padded planes per frame, frame f at + f * lowres_frame_bytes
    uint8_t *lowres; size_t lowres_frame_bytes; int lw, lh, lstride, lplane_bytes, lorigin;
    // adaptive quant: per-MB arrays (frame f at + f * mb_frame_stride elements), frame sums (6 per frame)
    float *qp_offset, *qp_offset_aq; uint16_t *inv_qscale; unsigned long long *stats; size_t mb_frame_stride;
    int aq_on, aq_mode; float strength;    // strength: aq-mode 1 (already * 1.0397)
    const float *log2_lut; const uint8_t *exp2_lut;
    int w, h, mb_w, mb_h, luma_h, flip;
    RgbKernelCoef k;
};

bool frontend_eligible(const void *src, long long src_stride, size_t src_frame_bytes, int w, int h, int n_frames,
                       const void *dst_y, int y_stride, const void *dst_u, const void *dst_v, int c_stride, size_t dst_frame_bytes);
int launch_frontend(cudaStream_t st, FrontendJob &job, const uint8_t *src, long long src_stride, size_t src_frame_bytes, int n_frames);

} // namespace xv
