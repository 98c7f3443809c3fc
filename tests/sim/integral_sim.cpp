// CPU lockstep run of the integral-image kernel (x264vfw_b200/csrc/integral_kernel.cuh) -- test infrastructure, see warp_sim.h.
// The kernel's warps are independent (a lane talks to its warp only, through shuffles), so every warp of the grid is run as 32
// OS threads in lockstep.  Mirrored from the CUDA headers: __shfl_down_sync / __shfl_sync (full mask, width 32: a source lane past
// the warp returns the caller's own value), __funnelshift_r, __dp4a (unsigned), __ldg.
#define XV_INTEGRAL_SIM 1
#include "warp_sim.h"

#define __global__
#define __device__
#define __forceinline__ inline
#define __launch_bounds__(...)
struct uint2 { uint32_t x, y; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
static thread_local dim3 blockIdx, threadIdx;
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) { s &= 31; return s ? (lo >> s) | (hi << (32 - s)) : lo; }
static inline uint32_t __dp4a(uint32_t a, uint32_t b, uint32_t c)
{
    for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 0xff) * ((b >> (8 * i)) & 0xff);
    return c;
}
static inline uint32_t __shfl_down_sync(unsigned, uint32_t v, int d) { return xv::sim_exchange(v, xv::t_lane + d <= 31 ? xv::t_lane + d : xv::t_lane); }
static inline uint32_t __shfl_sync(unsigned, uint32_t v, int src) { return xv::sim_exchange(v, src); }

#include "../../x264vfw_b200/csrc/integral_kernel.cuh"

// mirrors x264vfw_cuda_integral_init's launch: grid (stride / 128, strips / 4, planes), 4 warps per block
extern "C" int sim_integral(uint16_t *sum8, uint16_t *sum4, const uint8_t *plane, int stride, int rows, size_t plane_bytes, size_t sum_elems, int n)
{
    xv::IntegralJob j = {plane, sum8, sum4, stride, rows, plane_bytes, sum_elems};
    const int out_rows = sum4 ? rows - 3 : rows - 7;
    const int strips = (out_rows + IT_ROWS - 1) / IT_ROWS;
    const unsigned gx = (stride + 127) / 128, gy = (strips + 3) / 4;
    int warps = 0;
    for (unsigned bz = 0; bz < (unsigned)n; bz++) for (unsigned by = 0; by < gy; by++) for (unsigned bx = 0; bx < gx; bx++)
        for (unsigned wp = 0; wp < 4; wp++, warps++)
            xv::sim_run_warp([&](int lane) {
                blockIdx = dim3(bx, by, bz); threadIdx = dim3(32 * wp + lane, 0, 0);
                if (sum4) xv::integral_kernel<true>(j); else xv::integral_kernel<false>(j);
            });
    return warps;
}
