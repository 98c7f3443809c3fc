// Launch descriptors for the colour-space kernels (host <-> csp_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace xv {

// Fixed-point RGB->YUV coefficients, 20 fractional bits (csp.c:252-297 evaluated).
struct RgbCoef {
    uint32_t y_r, y_g, y_b, y_add;
    uint32_t u_r, u_g, u_b, u_add;
    uint32_t v_r, v_g, v_b, v_add;
};

// The same coefficients re-encoded for the kernel's dot-product form (launch_rgb_to_420 fills it).
struct RgbKernelCoef {
    uint32_t y_bg_hi, y_rx_hi, y_lo, y_add16;
    uint32_t u_add, u_b, u_r_neg, u_g_neg;
    uint32_t v_add, v_r, v_g_neg, v_b_neg;
};

struct RgbJob {
    const uint8_t *src; ptrdiff_t src_stride;     // negative stride = bottom-up DIB
    uint8_t *dst_y, *dst_u, *dst_v;               // NV12: dst_u = interleaved plane, dst_v unused
    int y_stride, u_stride, v_stride;
    int w, h;
    size_t src_frame_bytes, dst_frame_bytes;
    RgbCoef c;
    RgbKernelCoef k;
};

struct PackedJob {
    const uint8_t *src; ptrdiff_t src_stride;
    uint8_t *dst_y, *dst_u, *dst_v;
    int y_stride, u_stride, v_stride;
    int w, h;
    size_t src_frame_bytes, dst_frame_bytes;
};

struct PlaneOp {
    const uint8_t *src; ptrdiff_t src_stride;
    uint8_t *dst; int dst_stride;
    int w, h;          // destination size in bytes x rows
    int op;            // 0 copy, 1 subsamplev2, 2 subsamplehv2
    int nthreads;      // filled by launch_planes
};

struct PlanesJob {
    PlaneOp p[3];
    int n;
    size_t src_frame_bytes, dst_frame_bytes;
};

int launch_rgb_to_420(cudaStream_t st, const RgbJob &job, int bpp, bool nv12, bool vec, int n_frames);
int launch_packed422(cudaStream_t st, const PackedJob &job, bool uyvy, int mode, bool vec, int n_frames);
int launch_planes(cudaStream_t st, PlanesJob &job, bool vec, int n_frames);

// lowres_kernels.cu
struct LowresJob {
    const uint8_t *y; int y_stride; int w, h;      // tight luma (display size)
    uint8_t *dst;                                  // 4 padded planes per frame
    int luma_w, luma_h;                            // mod-16 size
    int lw, lh, lstride, lplane_bytes, lorigin;
    size_t src_frame_bytes, dst_frame_bytes;
};
int launch_lowres_init(cudaStream_t st, const LowresJob &job, int n_frames);

struct LumaPadJob {
    const uint8_t *y; int y_stride; int w, h;
    uint8_t *dst; int dst_stride; int luma_w, luma_h;
    size_t src_frame_bytes, dst_frame_bytes;
};
int launch_luma_pad(cudaStream_t st, const LumaPadJob &job, int n_frames);

struct ChromaPadJob {
    const uint8_t *u, *v; int c_stride; int w, h;     // planar 4:2:0 chroma, (w/2) x (h/2) each; w, h = luma size
    uint8_t *dst; int dst_stride; int luma_w, luma_h;  // interleaved plane, luma_w bytes x luma_h/2 rows
    size_t src_frame_bytes, dst_frame_bytes;
};
int launch_chroma_nv12_pad(cudaStream_t st, const ChromaPadJob &job, int n_frames);

// hpel_kernels.cu (the job descriptor is in hpel_kernel.cuh, shared with the CPU lockstep simulation)
struct HpelJob;
int launch_hpel(cudaStream_t st, HpelJob &job, int n_frames);

} // namespace xv
