// CPU lockstep run of the hpel warp program (x264vfw_b200/csrc/hpel_kernel.cuh) -- test infrastructure,
// see warp_sim.h.  Built by tests/test_hpel_sim.py with g++; mirrors hpel_kernels.cu:launch_hpel.
#include "warp_sim.h"
#include "../../x264vfw_b200/csrc/hpel_kernel.cuh"

extern "C" int sim_hpel(uint8_t *dst, const uint8_t *src, int src_stride, int w, int h, int stride,
                        size_t plane_bytes, int rows_per_strip, size_t sfb, size_t dfb, int n_frames)
{
    xv::HpelJob job;
    job.src = src; job.src_stride = src_stride; job.w = w; job.h = h;
    job.dst = dst; job.stride = stride; job.plane_bytes = plane_bytes;
    job.rows_per_strip = rows_per_strip;
    job.src_frame_bytes = sfb; job.dst_frame_bytes = dfb;
    const long long units = xv::hpel_plan(job, n_frames);
    for (int f = 0; f < n_frames; f++)
        for (long long u = 0; u < units; u++)
            xv::sim_run_warp([&](int lane) {
                if (job.aligned) xv::hpel_unit_any<true>(job, (int)u, f, lane);
                else xv::hpel_unit_any<false>(job, (int)u, f, lane);
            });
    return (int)units;
}
