// Device helpers shared by the lookahead kernels: scalar per-thread 8x8 block metrics and
// motion compensation on the four half-pel phase planes ([x264] common/pixel.c, common/mc.c).
#pragma once
#include "common.cuh"
#include "la_kernels.h"

namespace xv {

#define LA_COST_MAX (1 << 28)
#define LA_LOWRES_COST_MASK ((1 << 14) - 1)
#define LA_LOWRES_COST_SHIFT 14

__device__ __forceinline__ int clip3i(int v, int lo, int hi) { return min(max(v, lo), hi); }
__device__ __forceinline__ int clip_px(int v) { return min(max(v, 0), 255); }
__device__ __forceinline__ int median3i(int a, int b, int c) { return max(min(a, b), min(max(a, b), c)); }
__device__ __forceinline__ int mv_x(int packed) { return (int)(short)(packed & 0xffff); }
__device__ __forceinline__ int mv_y(int packed) { return packed >> 16; }
__device__ __forceinline__ int mv_pack(int x, int y) { return (x & 0xffff) | (y << 16); }

// 8 bytes from an arbitrarily aligned address: three aligned words + funnel shifts
__device__ __forceinline__ uint2 load8u(const uint8_t *p)
{
    const uintptr_t a = (uintptr_t)p;
    const uint32_t *q = (const uint32_t *)(a & ~(uintptr_t)3);
    const unsigned sh = (unsigned)(a & 3) * 8;
    uint32_t w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2);
    return make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
}

__device__ __forceinline__ int weight_px_dev(const WeightDev &w, int p)
{
    // [x264] mc_weight: opscale / opscale_noden
    if (w.denom >= 1) return clip_px(((p * w.scale + (1 << (w.denom - 1))) >> w.denom) + w.offset);
    return clip_px(p * w.scale + w.offset);
}
__device__ __forceinline__ uint32_t weight_word(const WeightDev &w, uint32_t v)
{
    return (uint32_t)weight_px_dev(w, v & 0xff) | ((uint32_t)weight_px_dev(w, (v >> 8) & 0xff) << 8) |
           ((uint32_t)weight_px_dev(w, (v >> 16) & 0xff) << 16) | ((uint32_t)weight_px_dev(w, v >> 24) << 24);
}

__constant__ const uint8_t c_hpel_ref0[16] = {0, 1, 1, 1, 0, 1, 1, 1, 2, 3, 3, 3, 0, 1, 1, 1};
__constant__ const uint8_t c_hpel_ref1[16] = {0, 0, 1, 0, 2, 2, 3, 2, 2, 2, 3, 2, 2, 2, 3, 2};

// One row (8 px) of [x264] get_ref at quarter-pel (mvx,mvy), row r of the block whose
// full-pel origin is `pel` (byte offset into every plane).
__device__ __forceinline__ uint2 get_ref_row(const uint8_t *const planes[4], int stride, int pel, int mvx, int mvy,
                                             int r, const WeightDev &w)
{
    const int qidx = ((mvy & 3) << 2) + (mvx & 3);
    const int off = pel + ((mvy >> 2) + r) * stride + (mvx >> 2);
    uint2 a = load8u(planes[c_hpel_ref0[qidx]] + off + ((mvy & 3) == 3) * stride);
    if (qidx & 5) {
        uint2 b = load8u(planes[c_hpel_ref1[qidx]] + off + ((mvx & 3) == 3));
        a.x = avg4(a.x, b.x);
        a.y = avg4(a.y, b.y);
    }
    if (w.on) { a.x = weight_word(w, a.x); a.y = weight_word(w, a.y); }
    return a;
}

// Same, for the four phase planes of one frame laid out plane_stride bytes apart (avoids a
// dynamically indexed pointer array, which would force the caller's state into local memory).
__device__ __forceinline__ uint2 get_ref_row_ps(const uint8_t *plane0, int plane_stride, int stride, int pel, int mvx, int mvy,
                                                int r, const WeightDev &w)
{
    const int qidx = ((mvy & 3) << 2) + (mvx & 3);
    const int off = pel + ((mvy >> 2) + r) * stride + (mvx >> 2);
    uint2 a = load8u(plane0 + (size_t)c_hpel_ref0[qidx] * plane_stride + off + ((mvy & 3) == 3) * stride);
    if (qidx & 5) {
        uint2 b = load8u(plane0 + (size_t)c_hpel_ref1[qidx] * plane_stride + off + ((mvx & 3) == 3));
        a.x = avg4(a.x, b.x);
        a.y = avg4(a.y, b.y);
    }
    if (w.on) { a.x = weight_word(w, a.x); a.y = weight_word(w, a.y); }
    return a;
}

// ---- scalar block metrics on 8 rows held as uint2 (thread-per-MB kernels) -------------------
__device__ __forceinline__ int sad8x8_rows(const uint2 a[8], const uint2 b[8])
{
    int s = 0;
#pragma unroll
    for (int r = 0; r < 8; r++) s += __vsadu4(a[r].x, b[r].x) + __vsadu4(a[r].y, b[r].y);
    return s;
}

__device__ __forceinline__ int px_of(uint2 v, int i) { return (int)(((i < 4 ? v.x : v.y) >> (8 * (i & 3))) & 0xff); }

// ---- SATD building blocks -------------------------------------------------------------------------------------
// Row transform of a 4x4 block WITHOUT extracting bytes: coefficient k of the row of differences a - b is
//   t_k = dp4a(a, H_k) + dp4a(b, -H_k)      (IDP.4A, unsigned pixels x signed Hadamard row)
// two dot products per coefficient instead of 8 byte extractions, 4 subtractions and 8 butterfly adds per row.
__device__ __forceinline__ int dp4a_us(uint32_t a, int b, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ void hadamard_row4(uint32_t a, uint32_t b, int t[4])
{
    t[0] = dp4a_us(a, 0x01010101, dp4a_us(b, (int)0xffffffffu, 0));      // + + + +
    t[1] = dp4a_us(a, (int)0xffff0101u, dp4a_us(b, 0x0101ffff, 0));      // + + - -
    t[2] = dp4a_us(a, 0x01ffff01, dp4a_us(b, (int)0xff0101ffu, 0));      // + - - +
    t[3] = dp4a_us(a, (int)0xff01ff01u, dp4a_us(b, 0x01ff01ff, 0));      // + - + -
}
// Column transform + sum of magnitudes of one coefficient column (v0..v3 = that coefficient of the 4 rows), HALVED:
// |s+t| + |s-t| = 2 max(|s|, |t|), so the last butterfly stage and the final >> 1 of x264_pixel_satd_8x4 fold into
// two maxima (exact in integers: every pair sum is even).
__device__ __forceinline__ int hadamard_col4_half(int v0, int v1, int v2, int v3)
{
    const int s01 = v0 + v1, e01 = v0 - v1, s23 = v2 + v3, e23 = v2 - v3;
    return max(abs(s01), abs(s23)) + max(abs(e01), abs(e23));
}
// one 4x4 block given its four rows as packed words: SATD contribution already halved
__device__ __forceinline__ int satd4x4_half(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                            uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3)
{
    int t0[4], t1[4], t2[4], t3[4];
    hadamard_row4(a0, b0, t0); hadamard_row4(a1, b1, t1); hadamard_row4(a2, b2, t2); hadamard_row4(a3, b3, t3);
    int sum = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) sum += hadamard_col4_half(t0[k], t1[k], t2[k], t3[k]);
    return sum;
}

// [x264] x264_pixel_satd_8x4 on rows r0..r0+3
__device__ __forceinline__ int satd8x4_rows(const uint2 a[8], const uint2 b[8], int r0)
{
    return satd4x4_half(a[r0].x, a[r0 + 1].x, a[r0 + 2].x, a[r0 + 3].x, b[r0].x, b[r0 + 1].x, b[r0 + 2].x, b[r0 + 3].x) +
           satd4x4_half(a[r0].y, a[r0 + 1].y, a[r0 + 2].y, a[r0 + 3].y, b[r0].y, b[r0 + 1].y, b[r0 + 2].y, b[r0 + 3].y);
}
__device__ __forceinline__ int satd8x8_rows(const uint2 a[8], const uint2 b[8])
{
    return satd8x4_rows(a, b, 0) + satd8x4_rows(a, b, 4);
}
__device__ __forceinline__ int mbcmp_rows(int satd, const uint2 a[8], const uint2 b[8])
{
    return satd ? satd8x8_rows(a, b) : sad8x8_rows(a, b);
}

// LPS = 4: a lane holds two rows.  Rows 0-3 live in lane parts 0,1 and rows 4-7 in parts 2,3:
// each lane pair forms one 8x4 (horizontal butterflies in the lane, vertical ones across the
// pair with packed 16-bit shuffles), the two halves add.
__device__ __forceinline__ int satd_rows4(uint2 fe0, uint2 fe1, uint2 a0, uint2 a1, bool odd)
{
    int s[8], d[8];
#pragma unroll
    for (int half = 0; half < 2; half++) {
        int t0[4], t1[4];
        hadamard_row4(half ? fe0.y : fe0.x, half ? a0.y : a0.x, t0);
        hadamard_row4(half ? fe1.y : fe1.x, half ? a1.y : a1.x, t1);
#pragma unroll
        for (int k = 0; k < 4; k++) { s[half * 4 + k] = t0[k] + t1[k]; d[half * 4 + k] = t0[k] - t1[k]; }
    }
    int sum = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int mine = (s[k] & 0xffff) | (d[k] << 16);
        const int other = __shfl_xor_sync(0xffffffffu, mine, 1);
        const int so = (int)(short)(other & 0xffff), dd = other >> 16;
        sum += odd ? abs(so - s[k]) + abs(dd - d[k]) : abs(s[k] + so) + abs(d[k] + dd);
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);     // 8x4 total in both lanes of the pair
    sum >>= 1;
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);     // + the other 8x4
    return sum;
}

// [x264] slicetype_mb_cost: spel limits of the lowres MB (the fpel limits are spel >> 2)
__device__ __forceinline__ void mv_limits(int mb_x, int mb_y, int mb_w, int mb_h, int mv_range2,
                                          int &min_x, int &max_x, int &min_y, int &max_y)
{
    min_x = max(4 * (-8 * mb_x - 12), -mv_range2);
    max_x = min(4 * (8 * (mb_w - mb_x - 1) + 12), mv_range2 - 1);
    min_y = max(4 * (-8 * mb_y - 12), -mv_range2);
    max_y = min(4 * (8 * (mb_h - mb_y - 1) + 12), mv_range2 - 1);
}

} // namespace xv
