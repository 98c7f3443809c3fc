// Exact integer RGB -> YUV arithmetic of the reference (csp.c:252-388), shared by the colour-space kernels and the
// fused front-end kernel.
#pragma once
#include "common.cuh"
#include "csp_kernels.h"

namespace xv {

// kernel-side coefficient forms (exact re-encodings of csp.c:252-297, see luma16 / chroma_quad)
static inline RgbKernelCoef make_rgb_kernel_coef(const RgbCoef &c)
{
    RgbKernelCoef k;
    const uint32_t yb = c.y_b << 4, yg = c.y_g << 4, yr = c.y_r << 4;
    k.y_bg_hi = (yb >> 8) | ((yg >> 8) << 16);
    k.y_rx_hi = (yr >> 8);
    k.y_lo = (yb & 0xff) | ((yg & 0xff) << 8) | ((yr & 0xff) << 16);
    k.y_add16 = c.y_add << 4;
    k.u_add = c.u_add; k.u_b = c.u_b; k.u_r_neg = 0u - c.u_r; k.u_g_neg = 0u - c.u_g;
    k.v_add = c.v_add; k.v_r = c.v_r; k.v_g_neg = 0u - c.v_g; k.v_b_neg = 0u - c.v_b;
    return k;
}

#ifdef __CUDACC__
// The reference arithmetic is Y = (Y_ADD + Y_R*r + Y_G*g + Y_B*b) >> 20 in uint32 (csp.c:334-337).
// The kernel is integer-pipe bound if done with per-channel extraction + IMAD, so the three
// products are evaluated with the integer dot-product unit instead, exactly:
//   16*coef = hi*256 + lo  (hi < 2^16, lo < 2^8)  =>  16*sum = ((dp2a(hi, px)) << 8) + dp4a(lo, px) + 16*Y_ADD
// and Y is then the top byte of the 32-bit result (16 * 2^20 = 2^24).  All intermediate values
// stay below 2^32 (max 16*(524288 + 1048576*255) = 4,286,578,688), so nothing wraps.
__device__ __forceinline__ uint32_t luma16(const RgbKernelCoef &k, uint32_t p)
{
    // p = B | G<<8 | R<<16 | X<<24; the X byte meets a zero coefficient
    const uint32_t t = __dp2a_hi(k.y_rx_hi, p, __dp2a_lo(k.y_bg_hi, p, 0u));
    return (t << 8) + __dp4a(p, k.y_lo, k.y_add16);
}

// U/V from the sums over the 2x2 quad (csp.c:371-379).  Channel sums are formed two at a
// time in 16-bit lanes (B|R and G|X); the subtractions are additions of the negated
// coefficients mod 2^32, which is the same uint32 arithmetic as the reference.
__device__ __forceinline__ void chroma_quad(const RgbKernelCoef &k, uint32_t p00, uint32_t p01,
                                            uint32_t p10, uint32_t p11, uint32_t &u, uint32_t &v)
{
    const uint32_t br = (p00 & 0x00ff00ffu) + (p01 & 0x00ff00ffu) + (p10 & 0x00ff00ffu) + (p11 & 0x00ff00ffu);
    const uint32_t gx = __byte_perm(p00, 0, 0x4341) + __byte_perm(p01, 0, 0x4341) +
                        __byte_perm(p10, 0, 0x4341) + __byte_perm(p11, 0, 0x4341);
    const uint32_t cb = br & 0xffffu, cr = br >> 16, cg = gx & 0xffffu;
    u = ((k.u_add + k.u_b * cb + k.u_r_neg * cr + k.u_g_neg * cg) >> 22) & 0xff;
    v = ((k.v_add + k.v_r * cr + k.v_g_neg * cg + k.v_b_neg * cb) >> 22) & 0xff;
}

#endif

} // namespace xv
