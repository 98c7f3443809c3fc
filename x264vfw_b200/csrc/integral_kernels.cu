// the integral image behind the half-pel planes -- launch and C ABI.  Kernel: integral_kernel.cuh (the same
// source runs on the CPU in lockstep, tests/sim/integral_sim.cpp).
#include "integral_kernel.cuh"
#include "../../include/x264vfw_cuda.h"


using namespace xv;

extern "C" int x264vfw_cuda_integral_init(x264vfw_cuda_ctx *ctx, uint16_t *sum8, uint16_t *sum4, const uint8_t *plane,
                                          int stride, int rows, size_t plane_bytes, size_t sum_elems, int n_frames)
{
    if (!ctx || !sum8 || !plane) { set_error("null argument"); return -1; }
    if (stride < 16 || rows < 8 || (stride & 3) || ((uintptr_t)plane & 3) || (plane_bytes & 3) || ((uintptr_t)sum8 & 7) || ((uintptr_t)sum4 & 7) || (sum_elems & 3)) {
        set_error("integral_init: plane 4-byte aligned, sums 8-byte aligned, stride a multiple of 4, at least 16 x 8");
        return -1;
    }
    if (n_frames <= 0) return 0;
    if (n_frames > 65535) { set_error("at most 65535 planes per launch"); return -1; }
    Ctx *c = (Ctx *)ctx;
    XV_CUDA_OK(cudaSetDevice(c->device));
    IntegralJob j = {plane, sum8, sum4, stride, rows, plane_bytes, sum_elems};
    // output rows 0 .. rows-8 (sum8) / rows-4 (sum4): strips of IT_ROWS rows
    const int out_rows = sum4 ? rows - 3 : rows - 7;
    const int strips = (out_rows + IT_ROWS - 1) / IT_ROWS;
    dim3 grid((stride + 127) / 128, (strips + 3) / 4, n_frames);
    if (sum4) integral_kernel<true><<<grid, 128, 0, c->stream>>>(j);
    else      integral_kernel<false><<<grid, 128, 0, c->stream>>>(j);
    XV_LAUNCH_CHECK();
    return 0;
}
