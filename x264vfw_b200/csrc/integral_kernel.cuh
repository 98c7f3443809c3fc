// the integral image [x264] x264_frame_filter builds behind the half-pel planes when the encoder
// searches exhaustively (me esa / tesa): common/mc.c integral_init8h + integral_init8v, and 4h + 4v for the 4x4 plane.
//
// Upstream fills it in place, row by row: a horizontal running sum added to the row above, then, 8 rows later, a vertical
// difference -- all in uint16 that wraps.  What is left, and what the exhaustive search reads, is a box sum:
//   sum8[y][x] = sum of the 8x8 pixels whose top-left is (x, y)   (mod 2^16),   sum4 likewise for 4x4.
// The device computes that directly, one WARP per 128 columns x IT_ROWS output rows and nothing shared between warps: a lane
// owns 4 columns, fetches its 4 pixels as one word and gets the 7 pixels to its right from the next two lanes by shuffle;
// the horizontal 4- and 8-sums are DP4A of byte-shifted words with 0x01010101; the last 8 (4) rows of those sums live in
// registers as a ring of packed 16-bit pairs, and the vertical sums are running sums over that ring (packed 32-bit adds are
// exact here: every box sum fits its 16-bit half).  A row of results leaves as one 8-byte store per lane.
// Bytes: 1 read + 2 (+2) written per pixel.
#pragma once
#ifndef XV_INTEGRAL_SIM
#include "common.cuh"
#endif
#include <stdint.h>
#include <stddef.h>

namespace xv {

#define IT_ROWS 64

struct IntegralJob {
    const uint8_t *plane; uint16_t *sum8, *sum4;
    int stride, rows;
    size_t plane_bytes, sum_elems;
};

__device__ __forceinline__ uint32_t it_word(const uint8_t *row, int cx, int stride)
{
    if (cx + 3 < stride) return __ldg((const uint32_t *)(row + cx));
    uint32_t v = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) if (cx + q < stride) v |= (uint32_t)__ldg(row + cx + q) << (8 * q);
    return v;
}

__device__ __forceinline__ void it_store4(uint16_t *dst, int x, int limit, uint32_t p0, uint32_t p1)
{
    if (x + 3 <= limit) { *(uint2 *)dst = make_uint2(p0, p1); return; }
    if (x <= limit) dst[0] = (uint16_t)p0;
    if (x + 1 <= limit) dst[1] = (uint16_t)(p0 >> 16);
    if (x + 2 <= limit) dst[2] = (uint16_t)p1;
}

template <bool SUM4>
__global__ void __launch_bounds__(128) integral_kernel(const IntegralJob j)
{
    const int lane = threadIdx.x & 31;
    const int strip = blockIdx.y * 4 + (threadIdx.x >> 5);
    const int y0 = strip * IT_ROWS;
    if (y0 + (SUM4 ? 4 : 8) > j.rows) return;                  // no complete window starts in this strip
    const int xb = blockIdx.x * 128, x = xb + 4 * lane;
    const uint8_t *P = j.plane + (size_t)blockIdx.z * j.plane_bytes;
    uint16_t *S8 = j.sum8 + (size_t)blockIdx.z * j.sum_elems;
    uint16_t *S4 = SUM4 ? j.sum4 + (size_t)blockIdx.z * j.sum_elems : nullptr;
    uint32_t r8[8][2], r4[4][2];                                // rings of packed horizontal sums (16 bits per column)
#pragma unroll
    for (int i = 0; i < 8; i++) r8[i][0] = r8[i][1] = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) r4[i][0] = r4[i][1] = 0;
    uint32_t run8[2] = {0, 0}, run4[2] = {0, 0};
    const int y_end = min(y0 + IT_ROWS + 7, j.rows);           // pixel rows y0 .. y0 + IT_ROWS + 6
    for (int base = 0; y0 + base < y_end; base += 8) {
        // all 8 rows of this trip are requested before the first result is stored (loads do not move above stores by themselves)
        uint32_t wv[8], ev[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int y = min(y0 + base + k, j.rows - 1);
            const uint8_t *row = P + (size_t)y * j.stride;
            wv[k] = it_word(row, x, j.stride);
            ev[k] = lane < 2 ? it_word(row, xb + 128 + 4 * lane, j.stride) : 0;     // the 8 pixels right of the warp's segment
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int y = y0 + base + k;
            if (y >= y_end) continue;                          // uniform; `continue` keeps the 8-row body unrolled with static ring slots
            const uint32_t w0 = wv[k];
            uint32_t w1 = __shfl_down_sync(0xffffffffu, w0, 1), w2 = __shfl_down_sync(0xffffffffu, w0, 2);
            const uint32_t e0 = __shfl_sync(0xffffffffu, ev[k], 0), e1 = __shfl_sync(0xffffffffu, ev[k], 1);
            if (lane == 31) { w1 = e0; w2 = e1; }
            if (lane == 30) w2 = e0;
            uint32_t h8p[2] = {0, 0}, h4p[2] = {0, 0};
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t a = __funnelshift_r(w0, w1, 8 * i), b = __funnelshift_r(w1, w2, 8 * i);   // pixels x+i .. x+i+3, x+i+4 .. x+i+7
                const uint32_t h4 = __dp4a(a, 0x01010101u, 0u);
                const uint32_t h8 = __dp4a(b, 0x01010101u, h4);
                h4p[i >> 1] |= h4 << (16 * (i & 1));
                h8p[i >> 1] |= h8 << (16 * (i & 1));
            }
            // vertical running sums over the last 8 / 4 rows; the slot being replaced holds the row that leaves the window
            run8[0] += h8p[0] - r8[k][0]; run8[1] += h8p[1] - r8[k][1];
            r8[k][0] = h8p[0]; r8[k][1] = h8p[1];
            if (SUM4) {
                run4[0] += h4p[0] - r4[k & 3][0]; run4[1] += h4p[1] - r4[k & 3][1];
                r4[k & 3][0] = h4p[0]; r4[k & 3][1] = h4p[1];
            }
            const int kk = base + k;
            if (kk >= 7) it_store4(S8 + (size_t)(y - 7) * j.stride + x, x, j.stride - 9, run8[0], run8[1]);
            if (SUM4 && kk >= 3 && y - 3 < y0 + IT_ROWS) it_store4(S4 + (size_t)(y - 3) * j.stride + x, x, j.stride - 5, run4[0], run4[1]);
        }
    }
}

} // namespace xv
