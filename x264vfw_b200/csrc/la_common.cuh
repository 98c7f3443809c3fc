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

// [x264] x264_pixel_satd_8x4 on rows r0..r0+3
__device__ __forceinline__ int satd8x4_rows(const uint2 a[8], const uint2 b[8], int r0)
{
    int sum = 0;
#pragma unroll
    for (int blk = 0; blk < 2; blk++) {
        int t[4][4];
#pragma unroll
        for (int y = 0; y < 4; y++) {
            uint32_t wa = blk ? a[r0 + y].y : a[r0 + y].x, wb = blk ? b[r0 + y].y : b[r0 + y].x;
            int d0 = (int)(wa & 0xff) - (int)(wb & 0xff), d1 = (int)((wa >> 8) & 0xff) - (int)((wb >> 8) & 0xff);
            int d2 = (int)((wa >> 16) & 0xff) - (int)((wb >> 16) & 0xff), d3 = (int)(wa >> 24) - (int)(wb >> 24);
            int s01 = d0 + d1, e01 = d0 - d1, s23 = d2 + d3, e23 = d2 - d3;
            t[y][0] = s01 + s23; t[y][1] = s01 - s23; t[y][2] = e01 + e23; t[y][3] = e01 - e23;
        }
#pragma unroll
        for (int x = 0; x < 4; x++) {
            int s01 = t[0][x] + t[1][x], e01 = t[0][x] - t[1][x], s23 = t[2][x] + t[3][x], e23 = t[2][x] - t[3][x];
            sum += abs(s01 + s23) + abs(s01 - s23) + abs(e01 + e23) + abs(e01 - e23);
        }
    }
    return sum >> 1;
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
        const uint32_t f0 = half ? fe0.y : fe0.x, f1 = half ? fe1.y : fe1.x;
        const uint32_t r0 = half ? a0.y : a0.x, r1 = half ? a1.y : a1.x;
        int t0[4], t1[4];
        {
            int e0 = (int)(f0 & 0xff) - (int)(r0 & 0xff), e1 = (int)((f0 >> 8) & 0xff) - (int)((r0 >> 8) & 0xff);
            int e2 = (int)((f0 >> 16) & 0xff) - (int)((r0 >> 16) & 0xff), e3 = (int)(f0 >> 24) - (int)(r0 >> 24);
            int s01 = e0 + e1, d01 = e0 - e1, s23 = e2 + e3, d23 = e2 - e3;
            t0[0] = s01 + s23; t0[1] = s01 - s23; t0[2] = d01 + d23; t0[3] = d01 - d23;
        }
        {
            int e0 = (int)(f1 & 0xff) - (int)(r1 & 0xff), e1 = (int)((f1 >> 8) & 0xff) - (int)((r1 >> 8) & 0xff);
            int e2 = (int)((f1 >> 16) & 0xff) - (int)((r1 >> 16) & 0xff), e3 = (int)(f1 >> 24) - (int)(r1 >> 24);
            int s01 = e0 + e1, d01 = e0 - e1, s23 = e2 + e3, d23 = e2 - e3;
            t1[0] = s01 + s23; t1[1] = s01 - s23; t1[2] = d01 + d23; t1[3] = d01 - d23;
        }
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
