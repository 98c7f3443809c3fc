// CPU run of the decoder-side conversion kernels (x264vfw_b200/csrc/decode_kernel.cuh): the very source the sm_100a build
// compiles -- kernels, host-side tables and dispatch -- compiled by g++ and executed thread by thread.  These kernels exchange
// nothing between threads (no shuffles, no shared memory, no barriers), so a grid is a plain loop nest.  TEST INFRASTRUCTURE:
// the CPU suite holds the kernel source against the checker before it goes to the GPU box; nothing in the product includes this
// file and it is far too slow to be a fallback.
//
// Mirrored from the CUDA headers / PTX ISA: __byte_perm (prmt default mode), __ldg, dp2a.{lo,hi}.s32.u32, cvt.pack.sat.u8.s32.b32
// (d = c << 16 | sat(a) << 8 | sat(b)), vector types.
#define XV_DECODE_SIM 1
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stddef.h>
#include <algorithm>
#include <functional>

#define __global__
#define __device__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define XV_LAUNCH_CHECK() do { } while (0)

struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
struct int4 { int x, y, z, w; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
static thread_local dim3 blockIdx, threadIdx, blockDim, gridDim;

using std::max;
using std::min;
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t s)
{
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
}

namespace xv {
struct Ctx;
static char g_sim_error[512];
static inline void set_error(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_sim_error, sizeof(g_sim_error), fmt, ap); va_end(ap); }
static inline uint32_t ldg_stream32(const void *p) { if ((uintptr_t)p & 3) abort(); return *(const uint32_t *)p; }      // the device loads need
static inline uint2 ldg_stream64(const void *p) { if ((uintptr_t)p & 7) abort(); return *(const uint2 *)p; }            // natural alignment
static inline uint4 ldg_stream128(const void *p) { if ((uintptr_t)p & 15) abort(); return *(const uint4 *)p; }
static inline int dp2a_lo_su(int a, uint32_t b, int c) { return c + (int)(int16_t)(a & 0xffff) * (int)(b & 0xff) + (int)(int16_t)(a >> 16) * (int)((b >> 8) & 0xff); }
static inline int dp2a_hi_su(int a, uint32_t b, int c) { return c + (int)(int16_t)(a & 0xffff) * (int)((b >> 16) & 0xff) + (int)(int16_t)(a >> 16) * (int)(b >> 24); }
static inline uint32_t pack_sat(int a, int b, uint32_t c)
{
    const uint32_t sa = a < 0 ? 0 : a > 255 ? 255 : a, sb = b < 0 ? 0 : b > 255 ? 255 : b;
    return (c << 16) | (sa << 8) | sb;
}
}

#define DEC_STREAM int
#define DEC_LAUNCH(grid, block, st, arg, ...) xv_sim_launch(grid, block, [&] { __VA_ARGS__(arg); })
static inline void xv_sim_launch(dim3 grid, dim3 block, const std::function<void()> &thread)
{
    gridDim = grid; blockDim = block;
    for (unsigned bz = 0; bz < grid.z; bz++) for (unsigned by = 0; by < grid.y; by++) for (unsigned bx = 0; bx < grid.x; bx++)
        for (unsigned ty = 0; ty < block.y; ty++) for (unsigned tx = 0; tx < block.x; tx++) {
            blockIdx = dim3(bx, by, bz); threadIdx = dim3(tx, ty, 0);
            thread();
        }
}

#include "../../x264vfw_b200/csrc/decode_kernel.cuh"

using namespace xv;

// One picture through dec_configure + dec_launch, all buffers in host memory.  Returns 0 / -1 (configure refused; sim_last_error()).
extern "C" int sim_dec_convert(int i_out_csp, int w, int h, int i_src_chroma, int i_avcol_spc, int b_fullrange,
                               uint8_t *dst, const uint8_t *y, const uint8_t *u, const uint8_t *v, int ys, int us, int vs,
                               size_t src_frame_bytes, size_t dst_frame_bytes, int n_frames)
{
    Dec d;
    DecTables t;
    if (dec_configure(d, t, i_out_csp, w, h, i_src_chroma, i_avcol_spc, b_fullrange) < 0) return -1;
    d.d_rows = t.rows.data(); d.d_cols = t.cols.data(); d.d_vtaps = t.vtaps.data();
    const uint8_t *src[3] = {y, u, v};
    const int ss[3] = {ys, us, vs};
    return dec_launch(&d, 0, dst, dst_frame_bytes, src, ss, src_frame_bytes, n_frames);
}
extern "C" const char *sim_last_error(void) { return g_sim_error; }
