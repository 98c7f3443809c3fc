// the integral image [x264] x264_frame_filter builds behind the half-pel planes when the encoder
// searches exhaustively (me esa / tesa): common/mc.c integral_init8h + integral_init8v, and 4h + 4v for the 4x4 plane.
//
// Upstream fills it in place, row by row: a horizontal running sum added to the row above, then, 8 rows later, a vertical
// difference -- all in uint16 that wraps.  What is left, and what the exhaustive search reads, is a box sum:
//   sum8[y][x] = sum of the 8x8 pixels whose top-left is (x, y)   (mod 2^16),   sum4 likewise for 4x4.
// The device computes that directly.  A block owns 128 columns x IT_ROWS output rows: every pixel row of the strip passes
// through shared memory once (coalesced 32-bit loads), a thread forms the horizontal 4- and 8-sums of its column from it
// and keeps the last 8 of them in registers as a ring, from which the vertical sums fall out as running sums.
// Bytes: 1 read + 2 (+2) written per pixel -- HBM-bound by bytes; one barrier per pixel row bounds it in practice.
#include "common.cuh"
#include "../../include/x264vfw_cuda.h"

namespace xv {

#define IT_COLS 128
#define IT_ROWS 64

struct IntegralJob {
    const uint8_t *plane; uint16_t *sum8, *sum4;
    int stride, rows;
    size_t plane_bytes, sum_elems;
};

template <bool SUM4>
__global__ void __launch_bounds__(IT_COLS) integral_kernel(const IntegralJob j)
{
    __shared__ __align__(16) uint8_t srow[2][IT_COLS + 16];
    const int tx = threadIdx.x;
    const int x = blockIdx.x * IT_COLS + tx;
    const int y0 = blockIdx.y * IT_ROWS;
    const uint8_t *P = j.plane + (size_t)blockIdx.z * j.plane_bytes;
    uint16_t *S8 = j.sum8 + (size_t)blockIdx.z * j.sum_elems;
    uint16_t *S4 = SUM4 ? j.sum4 + (size_t)blockIdx.z * j.sum_elems : nullptr;
    const int x_base = blockIdx.x * IT_COLS;
    int ring8[8], ring4[4];
#pragma unroll
    for (int i = 0; i < 8; i++) ring8[i] = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) ring4[i] = 0;
    int run8 = 0, run4 = 0;
    const int y_end = min(y0 + IT_ROWS + 7, j.rows);          // pixel rows this strip needs: y0 .. y0 + IT_ROWS + 6
    for (int y = y0; y < y_end; y++) {
        // stage pixel row y, columns x_base .. x_base + 135 (zeros past the row's end: those sums are never written)
        uint8_t *s = srow[(y - y0) & 1];
        const uint8_t *row = P + (size_t)y * j.stride;
        if (tx < (IT_COLS + 16) / 4) {
            const int cx = x_base + 4 * tx;
            uint32_t v = 0;
            if (cx + 3 < j.stride) v = __ldg((const uint32_t *)(row + cx));     // stride and plane are 4-byte aligned (checked by the launcher)
            else
                for (int q = 0; q < 4; q++) if (cx + q < j.stride) v |= (uint32_t)__ldg(row + cx + q) << (8 * q);
            ((uint32_t *)s)[tx] = v;
        }
        __syncthreads();                                       // one barrier per row: the other buffer is still being read
        const int h4 = s[tx] + s[tx + 1] + s[tx + 2] + s[tx + 3];
        const int h8 = h4 + s[tx + 4] + s[tx + 5] + s[tx + 6] + s[tx + 7];
        const int k = y - y0;
        // vertical running sums over the last 8 / 4 rows
        run8 += h8 - ring8[k & 7]; ring8[k & 7] = h8;
        if (SUM4) { run4 += h4 - ring4[k & 3]; ring4[k & 3] = h4; }
        if (k >= 7 && x <= j.stride - 9) S8[(size_t)(y - 7) * j.stride + x] = (uint16_t)run8;
        if (SUM4 && k >= 3 && y - 3 < y0 + IT_ROWS && x <= j.stride - 5) S4[(size_t)(y - 3) * j.stride + x] = (uint16_t)run4;
    }
}

} // namespace xv

using namespace xv;

extern "C" int x264vfw_cuda_integral_init(x264vfw_cuda_ctx *ctx, uint16_t *sum8, uint16_t *sum4, const uint8_t *plane,
                                          int stride, int rows, size_t plane_bytes, size_t sum_elems, int n_frames)
{
    if (!ctx || !sum8 || !plane) { set_error("null argument"); return -1; }
    if (stride < 16 || rows < 8 || (stride & 3) || ((uintptr_t)plane & 3) || (plane_bytes & 3)) { set_error("integral_init: plane must be 4-byte aligned, stride a multiple of 4, at least 16 x 8"); return -1; }
    if (n_frames <= 0) return 0;
    if (n_frames > 65535) { set_error("at most 65535 planes per launch"); return -1; }
    Ctx *c = (Ctx *)ctx;
    XV_CUDA_OK(cudaSetDevice(c->device));
    IntegralJob j = {plane, sum8, sum4, stride, rows, plane_bytes, sum_elems};
    // output rows 0 .. rows-8 (sum8) / rows-4 (sum4): strips of IT_ROWS rows
    const int out_rows = sum4 ? rows - 3 : rows - 7;
    dim3 grid((stride + IT_COLS - 1) / IT_COLS, (out_rows + IT_ROWS - 1) / IT_ROWS, n_frames);
    if (sum4) integral_kernel<true><<<grid, IT_COLS, 0, c->stream>>>(j);
    else      integral_kernel<false><<<grid, IT_COLS, 0, c->stream>>>(j);
    XV_LAUNCH_CHECK();
    return 0;
}
