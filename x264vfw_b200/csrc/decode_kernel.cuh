// the decoder-side output conversion of x264vfw_decompress (codec.c:2258-2292): kernels, the host-side tables
// they are handed, and the dispatch -- everything but the CUDA runtime calls (decode_kernels.cu).  tests/sim/decode_sim.cpp
// compiles this same file with g++ and runs it on the CPU against the checker.
//
// The reference hands every decoded yuv420p picture to libswscale's sws_scale() with the context of
// x264vfw_init_sws_context (codec.c:2075-2152).  What that context computes for a same-size picture
// is (DESIGN.md 4.5 has the derivation and how it is pinned against libswscale 9.1.100):
//   luma untouched; chroma kept at half horizontal resolution (SWS_FULL_CHR_H_INT never reaches the
//   context, codec.c:2097 vs :2110) and interpolated vertically by a 4-tap bicubic with 12-bit
//   coefficients; rows 0..h-3 go through libswscale's 16-bit "accurate rounding" SIMD writers,
//   rows h-2 and h-1 (and every row of UYVY) through its table-driven C writers; I420 / YV12 / NV12
//   are plane copies.
//
// Device formulation: one thread owns 8 pixels x the row pair (2k-1, 2k), whose chroma windows coincide.  The four chroma lines a row needs are
// loaded as 32-bit words (4 chroma samples each), transposed with PRMT so that one register holds
// the 4 vertical taps of one chroma column, and the filter is two DP2A (s16 coefficient pair x u8
// sample pair) per sample; pixels are packed with saturating I2IP (cvt.pack.sat.u8.s32) and leave
// as 128-bit stores.  HBM traffic per pixel is 1.5 bytes in + 2..4 bytes out; the kernel is bounded
// by instruction issue, like the RGB->I420 direction.
#pragma once
#ifndef XV_DECODE_SIM
#include "common.cuh"
#define DEC_STREAM cudaStream_t
#define DEC_LAUNCH(grid, block, st, arg, ...) __VA_ARGS__<<<grid, block, 0, st>>>(arg)
#endif
#include "../../include/x264vfw_cuda.h"
#include <vector>
#include <algorithm>
#include <string.h>
#include <stdlib.h>

namespace xv {

enum { DEC_BGRA = 0, DEC_BGR = 1, DEC_YUYV = 2, DEC_UYVY = 3 };
// resident blocks per SM the register allocation aims at: measured best per format on B200 (profiles/README.md R2.6)
#define DEC_BLOCKS_PER_SM(FMT) ((FMT) == DEC_BGRA || (FMT) == DEC_BGR ? 6 : 8)

struct DecRow {             // one output row of the vertical chroma filter
    int pos;                // first of the 4 chroma lines
    int c01, c23;           // coefficient pairs (s16 | s16 << 16) as the writer of this row sees them
    int c_writer;           // 1: libswscale's C writer (last two rows, UYVY), 0: its SIMD writer
};

struct DecConst {
    // SIMD writers: Yv = (y * yc + ykf) >> 13, chroma deltas = ((s >> 9) * coeff + coeff0) >> 16 with coeff0 = (rounder - 1024) * coeff
    int yc, ykf, vr, ub, vg, ug, vr0, ub0, vg0, ug0;
    // C writers: component = clip8((idx * cy + bias) >> 16), idx = Y + ((C * cxx) >> 16) - (cxx >> 9) ...
    int cy, bias, crv, cbu, cgu, cgv, crv9, cbu9, cgu9, cgv9;
    // 4:4:4 pictures (libswscale's full-chroma C writer): X = (y << 9) * fy + fy0 + ((u - 128) << 9) * fu.. + ..., >> 22
    int fy, fy0, fvr, fvg, fug, fub;
};

struct DecJob {
    const uint8_t *y, *u, *v;
    int ys, us, vs;
    uint8_t *dst;
    long long dst_stride;   // negative for a bottom-up DIB (x264vfw_picture_vflip, codec.c:510-527)
    int w, h;
    int v422;               // decoder picture is 4:2:2: every luma row has its own chroma line (no vertical filter)
    size_t src_frame_bytes, dst_frame_bytes;
    const DecRow *rows;
    DecConst k;
};

#ifndef XV_DECODE_SIM
__device__ __forceinline__ int dp2a_lo_su(int a, uint32_t b, int c)
{
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_su(int a, uint32_t b, int c)
{
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// (sat_u8(a) << 8 | sat_u8(b)) | c << 16
__device__ __forceinline__ uint32_t pack_sat(int a, int b, uint32_t c)
{
    uint32_t d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
#endif  // the CPU run of this source (tests/sim/decode_sim.cpp) supplies these three from the PTX ISA's definitions
#define DEC_ST128(p, v) (*(p) = (v))      // __stcs measured the same (profiles/README.md R2.6)
__device__ __forceinline__ int clip8(int v) { return min(max(v, 0), 255); }

// 4 bytes of a chroma line starting at column c0 (zeros past the line's end)
template <bool VEC>
__device__ __forceinline__ uint32_t load_c4(const uint8_t *line, int c0, int cw)
{
    if (VEC) return __ldg((const uint32_t *)(line + c0));
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 4; i++)
        if (c0 + i < cw) r |= (uint32_t)__ldg(line + c0 + i) << (8 * i);
    return r;
}

// One output row of 8 pixels: filter the 4 chroma columns, convert, store.
template <int FMT, bool VEC, bool CW>
__device__ __forceinline__ void dec_row_impl(const DecJob &j, const DecRow t, const uint32_t (&uc)[4], const uint32_t (&vc)[4],
                                             const uint32_t (&yw)[2], uint8_t *o, int npx)
{
    const DecConst &K = j.k;
    uint32_t px[8];                             // BGRA/BGR: one word per pixel (B | G<<8 | R<<16 | 255<<24); 4:2:2: one per pixel pair
#pragma unroll
    for (int c = 0; c < 4; c++) {
        // vertical filter: sum of sample * coefficient (the reference's 15-bit intermediates are sample << 7)
        const int su = dp2a_hi_su(t.c23, uc[c], dp2a_lo_su(t.c01, uc[c], 0));
        const int sv = dp2a_hi_su(t.c23, vc[c], dp2a_lo_su(t.c01, vc[c], 0));
        const uint32_t ywc = yw[c >> 1];
        if (FMT == DEC_YUYV || FMT == DEC_UYVY) {
            const int ya = (ywc >> (16 * (c & 1))) & 0xff, yb = (ywc >> (16 * (c & 1) + 8)) & 0xff;
            int Uo, Vo;
            if (CW) { Uo = (su + 2048) >> 12; Vo = (sv + 2048) >> 12; }               // (acc + (1 << 18)) >> 19
            else            { Uo = ((su >> 9) + 4) >> 3; Vo = ((sv >> 9) + 4) >> 3; } // psrad 16, +rounder, psraw 3
            px[c] = FMT == DEC_YUYV ? pack_sat(Uo, ya, pack_sat(Vo, yb, 0)) : pack_sat(ya, Uo, pack_sat(yb, Vo, 0));
        } else if (CW) {
            const int ya = (ywc >> (16 * (c & 1))) & 0xff, yb = (ywc >> (16 * (c & 1) + 8)) & 0xff;
            const int Uc = clip8((su + 2048) >> 12), Vc = clip8((sv + 2048) >> 12);
            const int dr = ((Vc * K.crv) >> 16) - K.crv9;
            const int db = ((Uc * K.cbu) >> 16) - K.cbu9;
            const int dg = ((Uc * K.cgu) >> 16) - K.cgu9 + ((Vc * K.cgv) >> 16) - K.cgv9;
            px[2 * c]     = pack_sat(((ya + dg) * K.cy + K.bias) >> 16, ((ya + db) * K.cy + K.bias) >> 16,
                                     pack_sat(255, ((ya + dr) * K.cy + K.bias) >> 16, 0));
            px[2 * c + 1] = pack_sat(((yb + dg) * K.cy + K.bias) >> 16, ((yb + db) * K.cy + K.bias) >> 16,
                                     pack_sat(255, ((yb + dr) * K.cy + K.bias) >> 16, 0));
        } else {
            // 16-bit SIMD writer: chroma = (s >> 9) + 4 - (128 << 3); delta = chroma * coeff >> 16 (constants folded);
            // luma = ((y << 3) + 4 - y_offset) * y_coeff >> 16, as one DP2A on the packed bytes and one shift
            const int uq = su >> 9, vq = sv >> 9;
            const int ub = (uq * K.ub + K.ub0) >> 16, vr = (vq * K.vr + K.vr0) >> 16;
            const int g = ((uq * K.ug + K.ug0) >> 16) + ((vq * K.vg + K.vg0) >> 16);
            const int y0v = ((c & 1) ? dp2a_hi_su(K.yc, ywc, K.ykf) : dp2a_lo_su(K.yc, ywc, K.ykf)) >> 13;
            const int y1v = ((c & 1) ? dp2a_hi_su(K.yc << 16, ywc, K.ykf) : dp2a_lo_su(K.yc << 16, ywc, K.ykf)) >> 13;
            px[2 * c]     = pack_sat(y0v + g, y0v + ub, pack_sat(255, y0v + vr, 0));
            px[2 * c + 1] = pack_sat(y1v + g, y1v + ub, pack_sat(255, y1v + vr, 0));
        }
    }
    if (FMT == DEC_BGRA) {
        if (VEC) {
            DEC_ST128((uint4 *)o, make_uint4(px[0], px[1], px[2], px[3]));
            DEC_ST128((uint4 *)o + 1, make_uint4(px[4], px[5], px[6], px[7]));
        } else {
#pragma unroll
            for (int q = 0; q < 8; q++)
                if (q < npx) *(uint32_t *)(o + 4 * q) = px[q];          // DIB rows are 4-byte aligned by construction
        }
    } else if (FMT == DEC_BGR) {
        if (VEC) {
            uint32_t wd[6];
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const uint32_t p0 = px[4 * q], p1 = px[4 * q + 1], p2 = px[4 * q + 2], p3 = px[4 * q + 3];
                wd[3 * q]     = __byte_perm(p0, p1, 0x4210);          // B0 G0 R0 B1
                wd[3 * q + 1] = __byte_perm(p1, p2, 0x5421);          // G1 R1 B2 G2
                wd[3 * q + 2] = __byte_perm(p2, p3, 0x6542);          // R2 B3 G3 R3
            }
#pragma unroll
            for (int q = 0; q < 6; q++) ((uint32_t *)o)[q] = wd[q];
        } else {
#pragma unroll
            for (int q = 0; q < 8; q++)
                if (q < npx) { o[3 * q] = px[q]; o[3 * q + 1] = px[q] >> 8; o[3 * q + 2] = px[q] >> 16; }
        }
    } else {
        if (VEC)
            DEC_ST128((uint4 *)o, make_uint4(px[0], px[1], px[2], px[3]));
        else {
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (2 * q < npx) *(uint32_t *)(o + 4 * q) = px[q];
        }
    }
}

template <int FMT, bool VEC>
__device__ __forceinline__ void dec_row(const DecJob &j, const DecRow t, const uint32_t (&uc)[4], const uint32_t (&vc)[4],
                                        const uint32_t (&yw)[2], uint8_t *o, int npx)
{
    if (FMT == DEC_UYVY || t.c_writer) dec_row_impl<FMT, VEC, true>(j, t, uc, vc, yw, o, npx);
    else                               dec_row_impl<FMT, VEC, false>(j, t, uc, vc, yw, o, npx);
}

// 8 luma bytes of a row (zeros past the picture's right edge)
template <bool VEC>
__device__ __forceinline__ void dec_luma8(const uint8_t *yrow, int npx, uint32_t (&yw)[2])
{
    if (VEC) {
        const uint2 t2 = ldg_stream64(yrow);
        yw[0] = t2.x; yw[1] = t2.y;
    } else {
        yw[0] = yw[1] = 0;
#pragma unroll
        for (int q = 0; q < 8; q++)
            if (q < npx) yw[q >> 2] |= (uint32_t)__ldg(yrow + q) << (8 * (q & 3));
    }
}

// The 4 chroma lines starting at `pos`, 4 columns from c0, transposed: word c = the 4 vertical taps of column c0 + c.
template <bool VEC>
__device__ __forceinline__ void dec_window(const uint8_t *plane, int stride, int pos, int c0, int cw, uint32_t (&col)[4])
{
    uint32_t l[4];
    const uint8_t *p = plane + (ptrdiff_t)pos * stride;
#pragma unroll
    for (int q = 0; q < 4; q++, p += stride) l[q] = load_c4<VEC>(p, c0, cw);
    const uint32_t a0 = __byte_perm(l[0], l[1], 0x5140), a1 = __byte_perm(l[0], l[1], 0x7362);
    const uint32_t a2 = __byte_perm(l[2], l[3], 0x5140), a3 = __byte_perm(l[2], l[3], 0x7362);
    col[0] = __byte_perm(a0, a2, 0x5410); col[1] = __byte_perm(a0, a2, 0x7632);
    col[2] = __byte_perm(a1, a3, 0x5410); col[3] = __byte_perm(a1, a3, 0x7632);
}

// Thread = 8 pixels x the row pair (2k-1, 2k): in libswscale's filter those two rows read the same 4 chroma lines
// (rows -1 and h do not exist).
template <int FMT, bool VEC>
__global__ void __launch_bounds__(256, DEC_BLOCKS_PER_SM(FMT)) dec_packed_kernel(const __grid_constant__ DecJob j)
{
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * 8;
    const int k = blockIdx.y * 8 + threadIdx.y;
    const int ra = 2 * k - 1, rb = 2 * k;
    if (x0 >= j.w || ra >= j.h) return;
    const int cw = j.w >> 1, c0 = x0 >> 1;
    const int npx = min(8, j.w - x0);
    const size_t fo = (size_t)blockIdx.z * j.src_frame_bytes;
    const uint8_t *Y = j.y + fo + x0, *U = j.u + fo, *V = j.v + fo;
    constexpr int BPP2 = FMT == DEC_BGRA ? 8 : FMT == DEC_BGR ? 6 : 4;       // bytes per pixel pair
    uint8_t *D = j.dst + (size_t)blockIdx.z * j.dst_frame_bytes + (size_t)c0 * BPP2;

    // both rows of the pair read chroma lines pos..pos+3, pos = clamp(k - 2, 0, h/2 - 4) (checked against the filter
    // table when the context is opened), so every load of this thread can be issued before anything is computed
    // (4:2:2 pictures: the pair's two chroma lines 2k-1, 2k lie in the window that starts at clamp(2k - 1, 0, h - 4))
    const int pos = j.v422 ? min(max(2 * k - 1, 0), j.h - 4) : min(max(k - 2, 0), (j.h >> 1) - 4);
    const int4 a4 = __ldg((const int4 *)(j.rows + max(ra, 0))), b4 = __ldg((const int4 *)(j.rows + min(rb, j.h - 1)));
    const DecRow ta = {a4.x, a4.y, a4.z, a4.w}, tb = {b4.x, b4.y, b4.z, b4.w};
    uint32_t uc[4], vc[4], ya[2], yb[2];
    dec_luma8<VEC>(Y + (ptrdiff_t)max(ra, 0) * j.ys, npx, ya);             // before any store: a load after a store waits for it
    dec_luma8<VEC>(Y + (ptrdiff_t)min(rb, j.h - 1) * j.ys, npx, yb);
    dec_window<VEC>(U, j.us, pos, c0, cw, uc);
    dec_window<VEC>(V, j.vs, pos, c0, cw, vc);
    if (ra >= 0)
        dec_row<FMT, VEC>(j, ta, uc, vc, ya, D + (ptrdiff_t)ra * j.dst_stride, npx);
    if (rb < j.h)
        dec_row<FMT, VEC>(j, tb, uc, vc, yb, D + (ptrdiff_t)rb * j.dst_stride, npx);
}

// 4:4:4 decoder pictures -> RGB: libswscale turns full chroma interpolation on by itself and every row goes through its C writer
// yuv2rgb_full_1_c / yuv2rgb_write_full: per pixel, 32-bit integer arithmetic that wraps, clip to 30 bits, >> 22.
template <bool BGRA, bool VEC>
__global__ void __launch_bounds__(256) dec_444_kernel(const __grid_constant__ DecJob j)
{
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * 8, r = blockIdx.y * 8 + threadIdx.y;       // thread = 8 pixels of a row
    if (x0 >= j.w || r >= j.h) return;
    const size_t fo = (size_t)blockIdx.z * j.src_frame_bytes;
    const uint8_t *py = j.y + fo + (ptrdiff_t)r * j.ys + x0, *pu = j.u + fo + (ptrdiff_t)r * j.us + x0, *pv = j.v + fo + (ptrdiff_t)r * j.vs + x0;
    const int npx = min(8, j.w - x0);
    uint32_t yw[2] = {0, 0}, uw[2] = {0, 0}, vw[2] = {0, 0};
    if (VEC) {
        const uint2 a = ldg_stream64(py), b = ldg_stream64(pu), c = ldg_stream64(pv);
        yw[0] = a.x; yw[1] = a.y; uw[0] = b.x; uw[1] = b.y; vw[0] = c.x; vw[1] = c.y;
    } else
        for (int q = 0; q < npx; q++) {
            yw[q >> 2] |= (uint32_t)__ldg(py + q) << (8 * (q & 3)); uw[q >> 2] |= (uint32_t)__ldg(pu + q) << (8 * (q & 3));
            vw[q >> 2] |= (uint32_t)__ldg(pv + q) << (8 * (q & 3));
        }
    const DecConst &K = j.k;
    uint32_t px[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const int sh = 8 * (q & 3);
        const int Y = (int)(((yw[q >> 2] >> sh) & 0xff) << 9) * K.fy + K.fy0;
        const int U = ((int)((uw[q >> 2] >> sh) & 0xff) - 128) << 9, V = ((int)((vw[q >> 2] >> sh) & 0xff) - 128) << 9;
        // unsigned wrap-around like the C writer's (unsigned) products; a sum past 2^31 turns negative and clips to 0
        const int R = (int)((unsigned)Y + (unsigned)V * (unsigned)K.fvr);
        const int G = (int)((unsigned)Y + (unsigned)V * (unsigned)K.fvg + (unsigned)U * (unsigned)K.fug);
        const int B = (int)((unsigned)Y + (unsigned)U * (unsigned)K.fub);
        const int lim = (1 << 30) - 1;
        px[q] = (uint32_t)(min(max(B, 0), lim) >> 22) | ((uint32_t)(min(max(G, 0), lim) >> 22) << 8) |
                ((uint32_t)(min(max(R, 0), lim) >> 22) << 16) | 0xff000000u;
    }
    uint8_t *o = j.dst + (size_t)blockIdx.z * j.dst_frame_bytes + (ptrdiff_t)r * j.dst_stride + (size_t)x0 * (BGRA ? 4 : 3);
    if (BGRA) {
        if (VEC) { ((uint4 *)o)[0] = make_uint4(px[0], px[1], px[2], px[3]); ((uint4 *)o)[1] = make_uint4(px[4], px[5], px[6], px[7]); }
        else for (int q = 0; q < npx; q++) *(uint32_t *)(o + 4 * q) = px[q];
    } else if (VEC) {
#pragma unroll
        for (int q = 0; q < 2; q++) {
            ((uint32_t *)o)[3 * q]     = __byte_perm(px[4 * q], px[4 * q + 1], 0x4210);
            ((uint32_t *)o)[3 * q + 1] = __byte_perm(px[4 * q + 1], px[4 * q + 2], 0x5421);
            ((uint32_t *)o)[3 * q + 2] = __byte_perm(px[4 * q + 2], px[4 * q + 3], 0x6542);
        }
    } else
        for (int q = 0; q < npx; q++) { o[3 * q] = px[q]; o[3 * q + 1] = px[q] >> 8; o[3 * q + 2] = px[q] >> 16; }
}

// I420 / YV12 / NV12 targets: plane copies (libswscale's planarCopyWrapper / planarToNv12Wrapper); YV12 arrives
// here with the destination U/V pointers already swapped (codec.c:2263-2274).
struct DecPlanarJob {
    const uint8_t *y, *u, *v;
    int ys, us, vs;
    uint8_t *dy, *du, *dv;      // dv == nullptr: NV12 (du rows hold U,V interleaved)
    int w, h, cw;               // cw: chroma plane width (w / 2, or w for 4:4:4); chroma rows follow from the grid: rows h .. gridDim.y - 1
    size_t src_frame_bytes, dst_frame_bytes;
};

template <bool VEC>
__global__ void __launch_bounds__(256) dec_planar_kernel(const DecPlanarJob j)
{
    const int xb = (blockIdx.x * 256 + threadIdx.x) * 16;
    const int row = blockIdx.y;
    const size_t so = (size_t)blockIdx.z * j.src_frame_bytes, dof = (size_t)blockIdx.z * j.dst_frame_bytes;
    const int cw = j.cw;
    if (row < j.h) {
        if (xb >= j.w) return;
        const uint8_t *s = j.y + so + (ptrdiff_t)row * j.ys + xb;
        uint8_t *d = j.dy + dof + (size_t)row * j.w + xb;
        if (VEC) DEC_ST128((uint4 *)d, ldg_stream128(s));
        else for (int q = 0; q < 16 && xb + q < j.w; q++) d[q] = __ldg(s + q);
        return;
    }
    const int r = row - j.h;
    const uint8_t *su = j.u + so + (ptrdiff_t)r * j.us, *sv = j.v + so + (ptrdiff_t)r * j.vs;
    if (!j.dv) {                                            // NV12: 16 output bytes = 8 U + 8 V
        if (xb >= j.w) return;
        uint8_t *d = j.du + dof + (size_t)r * j.w + xb;
        if (VEC) {
            const uint2 a = ldg_stream64(su + (xb >> 1)), b = ldg_stream64(sv + (xb >> 1));
            DEC_ST128((uint4 *)d, make_uint4(__byte_perm(a.x, b.x, 0x5140), __byte_perm(a.x, b.x, 0x7362),
                                             __byte_perm(a.y, b.y, 0x5140), __byte_perm(a.y, b.y, 0x7362)));
        } else
            for (int q = 0; q < 8 && (xb >> 1) + q < cw; q++) { d[2 * q] = __ldg(su + (xb >> 1) + q); d[2 * q + 1] = __ldg(sv + (xb >> 1) + q); }
        return;
    }
    if (xb >= cw) return;
    uint8_t *du = j.du + dof + (size_t)r * cw + xb, *dv = j.dv + dof + (size_t)r * cw + xb;
    if (VEC) { DEC_ST128((uint4 *)du, ldg_stream128(su + xb)); DEC_ST128((uint4 *)dv, ldg_stream128(sv + xb)); }
    else for (int q = 0; q < 16 && xb + q < cw; q++) { du[q] = __ldg(su + xb + q); dv[q] = __ldg(sv + xb + q); }
}

// YUV output with ANOTHER chroma resolution than the decoder picture (4:2:0 -> YV16 / YV24, 4:2:2 -> I420 / YV12 / NV12 / YV24,
// 4:4:4 -> I420 / YV12 / NV12 / YV16): libswscale's general scaler on a chroma plane -- horizontal bicubic (4 taps for 2x up, 8 for
// 2:1 down, 14-bit) into 15-bit intermediates, c15 = min((sum tap * sample) >> 7, 32767), then the vertical filter (4 / 8 taps,
// 12-bit) and the 8-bit plane writer, out = clip8((sum tap * c15 + (64 << 12)) >> 19).  Where a direction is not scaled its table
// holds one tap of 1.0 (16384 / 4096), which makes the same expression the identity.  Thread = one output sample of both
// planes; rare paths, kept simple (zero taps are skipped, which also keeps every load inside the plane).
struct DecTap8 { int pos; int c[4]; int pad[3]; };          // first sample + 8 coefficients as s16 pairs

struct DecUpJob {
    const uint8_t *u, *v; int us, vs;
    uint8_t *du, *dv;               // first output sample of each plane
    int ostride, ostep;             // bytes between output rows / samples (NV12: w, 2; planar: chroma width, 1)
    int ocw, och;                   // output chroma plane
    const DecTap8 *cols, *rows;     // horizontal / vertical taps per output column / row
    size_t src_frame_bytes, dst_frame_bytes;
};

__device__ __forceinline__ int dec_tap(const int4 &a, const int4 &b, int i)
{
    // a = {pos, c01, c23, c45}, b.x = c67
    const int w = i < 2 ? a.y : i < 4 ? a.z : i < 6 ? a.w : b.x;
    return (i & 1) ? w >> 16 : (int)(short)(w & 0xffff);
}

// 15-bit intermediate of one chroma line at the output column whose taps are (ha, hb): hScale8To15
__device__ __forceinline__ int dec_hscale(const uint8_t *line, const int4 &ha, const int4 &hb)
{
    int a = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int c = dec_tap(ha, hb, i);
        if (c) a += (int)__ldg(line + ha.x + i) * c;
    }
    return min(a >> 7, 32767);
}

__global__ void __launch_bounds__(256) dec_resample_kernel(const DecUpJob j)
{
    const int x = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y * 8 + threadIdx.y;
    if (x >= j.ocw || r >= j.och) return;
    const int4 ha = __ldg((const int4 *)(j.cols + x)), hb = __ldg((const int4 *)(j.cols + x) + 1);
    const int4 va = __ldg((const int4 *)(j.rows + r)), vb = __ldg((const int4 *)(j.rows + r) + 1);
    const size_t so = (size_t)blockIdx.z * j.src_frame_bytes, dof = (size_t)blockIdx.z * j.dst_frame_bytes;
#pragma unroll
    for (int c = 0; c < 2; c++) {
        const uint8_t *sp = (c ? j.v : j.u) + so;
        const int st = c ? j.vs : j.us;
        int acc = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int vcoef = dec_tap(va, vb, q);
            if (vcoef) acc += dec_hscale(sp + (ptrdiff_t)(va.x + q) * st, ha, hb) * vcoef;
        }
        ((c ? j.dv : j.du) + dof)[(size_t)r * j.ostride + (size_t)x * j.ostep] = (uint8_t)clip8((acc + (64 << 12)) >> 19);
    }
}

// 4:4:4 picture -> YUY2 / UYVY: chroma down-sampled horizontally as above, then libswscale's single-line packed writers:
// yuv2yuyv422_1 SIMD in rows 0..h-3 (c15 >> 7), the C writer in the last two rows and in every row of UYVY ((c15 + 64) >> 7).
// Thread = one pixel pair.
struct Dec444PackedJob {
    const uint8_t *y, *u, *v; int ys, us, vs;
    uint8_t *dst;
    int w, h, uyvy;
    const DecTap8 *cols;
    size_t src_frame_bytes, dst_frame_bytes;
};

__global__ void __launch_bounds__(256) dec_444_packed_kernel(const Dec444PackedJob j)
{
    const int x = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y * 8 + threadIdx.y;
    if (2 * x >= j.w || r >= j.h) return;
    const int4 ha = __ldg((const int4 *)(j.cols + x)), hb = __ldg((const int4 *)(j.cols + x) + 1);
    const size_t so = (size_t)blockIdx.z * j.src_frame_bytes;
    const int cu = dec_hscale(j.u + so + (ptrdiff_t)r * j.us, ha, hb), cv = dec_hscale(j.v + so + (ptrdiff_t)r * j.vs, ha, hb);
    const bool c_writer = j.uyvy || r >= j.h - 2;
    const int U = clip8(c_writer ? (cu + 64) >> 7 : cu >> 7), V = clip8(c_writer ? (cv + 64) >> 7 : cv >> 7);
    const uint8_t *yp = j.y + so + (ptrdiff_t)r * j.ys + 2 * x;
    const int y0 = __ldg(yp), y1 = __ldg(yp + 1);
    const uint32_t px = j.uyvy ? (uint32_t)U | (y0 << 8) | (V << 16) | (y1 << 24) : (uint32_t)y0 | (U << 8) | (y1 << 16) | (V << 24);
    *(uint32_t *)(j.dst + (size_t)blockIdx.z * j.dst_frame_bytes + (size_t)r * 2 * j.w + 4 * (size_t)x) = px;
}

// ---- host side: the context x264vfw_init_sws_context builds, as numbers --------------------------------------

// The vertical chroma filter libswscale's initFilter() [libswscale/utils.c] yields for this context: bicubic
// (B = 0, C = 0.6), src_n -> 2 * src_n, both chroma sitings 128, coefficients normalised to 1 << 12, at most 4 taps
// after near-zero taps are dropped, out-of-picture taps folded onto the edge line.
static bool bicubic_2x_filter(int src_n, std::vector<DecRow> &rows, bool c_writer_everywhere, int one, int align, bool kernel_pos_rule);
static bool vertical_chroma_filter(int src_n, std::vector<DecRow> &rows, bool c_writer_everywhere)
{
    return bicubic_2x_filter(src_n, rows, c_writer_everywhere, 1 << 12, 2, true);
}

static bool bicubic_2x_filter(int src_n, std::vector<DecRow> &rows, bool c_writer_everywhere, int one, int align, bool kernel_pos_rule)
{
    const int dst_n = 2 * src_n;
    if (src_n < 6) return false;                               // below that the tap positions stop following the kernel's rule
    const int taps = src_n - 2 < 5 ? src_n - 2 : 5;            // 1 + sizeFactor(bicubic), capped by the source height
    const long long inc = (((long long)src_n << 16) + (dst_n >> 1)) / dst_n;        // 1 << 15
    const long long Cq = (long long)(0.6 * (1 << 24));
    std::vector<long long> f((size_t)dst_n * taps);
    std::vector<int> pos(dst_n);
    long long at = ((128 * inc) >> 7) - ((128 * 0x10000LL) >> 7);
    for (int i = 0; i < dst_n; i++, at += 2 * inc) {
        int xx = (int)((at - (long long)(taps - 2) * (1LL << 16)) / (1 << 17));
        pos[i] = xx;
        for (int t = 0; t < taps; t++, xx++) {
            const long long d = llabs((long long)xx * (1 << 17) - at) << 13;
            long long c = 0;
            if (d < (1LL << 31)) {
                const long long dd = (d * d) >> 30, ddd = (dd * d) >> 30;
                c = d < (1LL << 30)
                        ? (12 * (1 << 24) - 6 * Cq) * ddd + (-18 * (1 << 24) + 6 * Cq) * dd + 6LL * (1 << 24) * (1LL << 30)
                        : -6 * Cq * ddd + 30 * Cq * dd - 48 * Cq * d + 24 * Cq * (1LL << 30);
            }
            f[(size_t)i * taps + t] = c;
        }
    }
    // drop near-zero leading taps (keeping positions monotonic), measure the longest remaining support
    const double cut = 0.002 * 18014398509481984.0;      // SWS_MAX_REDUCE_CUTOFF * 2^54
    int support = 0;
    for (int i = dst_n - 1; i >= 0; i--) {
        long long *fi = &f[(size_t)i * taps];
        long long acc = 0;
        for (int t = 0; t < taps; t++) {
            acc += llabs(fi[0]);
            if ((double)acc > cut || (i < dst_n - 1 && pos[i] >= pos[i + 1])) break;
            memmove(fi, fi + 1, (taps - 1) * sizeof(*fi));
            fi[taps - 1] = 0;
            pos[i]++;
        }
        int n = taps;
        acc = 0;
        for (int t = taps - 1; t > 0; t--) {
            acc += llabs(fi[t]);
            if ((double)acc > cut) break;
            n--;
        }
        support = n > support ? n : support;
    }
    const int size = (support + align - 1) & ~(align - 1);   // filterAlign of the x86 build: 2 vertical, 4 horizontal
    if (size != 4) return false;                          // the kernel reads 4 lines per row
    rows.resize(dst_n);
    for (int i = 0; i < dst_n; i++) {
        long long t[4] = {0, 0, 0, 0};
        for (int q = 0; q < size && q < taps; q++) t[q] = f[(size_t)i * taps + q];
        if (pos[i] < 0) {
            for (int q = 1; q < size; q++) { const int to = q + pos[i] > 0 ? q + pos[i] : 0; t[to] += t[q]; t[q] = 0; }
            pos[i] = 0;
        }
        if (pos[i] + size > src_n) {
            const int shift = pos[i] + size - src_n;       // src_n >= 5 > taps kept
            long long acc = 0;
            for (int q = size - 1; q >= 0; q--) if (pos[i] + q >= src_n) { acc += t[q]; t[q] = 0; }
            for (int q = size - 1; q >= 0; q--) t[q] = q < shift ? 0 : t[q - shift];
            pos[i] -= shift;
            t[src_n - 1 - pos[i]] += acc;
        }
        long long sum = 0, err = 0;
        for (int q = 0; q < 4; q++) sum += t[q];
        sum = (sum + one / 2) / one;
        if (!sum) sum = 1;
        int c[4];
        for (int q = 0; q < 4; q++) {
            const long long v = t[q] + err;
            const long long iv = v >= 0 ? (v + (sum >> 1)) / sum : (v - (sum >> 1)) / sum;
            c[q] = (int)iv;
            err = v - iv * sum;
        }
        const bool cwr = c_writer_everywhere || i >= dst_n - 2;    // libswscale leaves SIMD for the last two lines
        if (!cwr)          // ff_updateMMXDitherTables packs f[q] + f[q+1] * 65536 into ONE int: a negative f[q] borrows
            for (int q = 0; q < 4; q += 2) if (c[q] < 0) c[q + 1] = (int16_t)(c[q + 1] - 1);
        if (kernel_pos_rule && pos[i] != std::min(std::max(((i + 1) >> 1) - 2, 0), src_n - 4)) return false;   // dec_packed_kernel derives it
        rows[i].pos = pos[i];
        rows[i].c01 = (c[0] & 0xffff) | (int)((uint32_t)c[1] << 16);
        rows[i].c23 = (c[2] & 0xffff) | (int)((uint32_t)c[3] << 16);
        rows[i].c_writer = cwr;
    }
    return true;
}

// Tap tables of dec_resample_kernel: libswscale's initFilter() for the bicubic cases this context meets -- 2x up-sampling (1 + 4
// taps before the cut), 2:1 down-sampling (1 + 4 * src / dst = 9 taps before the cut, distances scaled by dst / src, unit halved) --
// or one tap of 1.0 when the sizes agree.  At most 8 taps survive the cut; out-of-plane taps are folded onto the edge sample.
static bool resample_taps(int src_n, int dst_n, int one, int align, std::vector<DecTap8> &out)
{
    out.assign(dst_n, DecTap8{0, {0, 0, 0, 0}, {0, 0, 0}});
    auto put = [&](int i, int k, int c) { out[i].c[k >> 1] |= (k & 1) ? (int)((unsigned)(c & 0xffff) << 16) : (c & 0xffff); };
    if (src_n == dst_n) {
        for (int i = 0; i < dst_n; i++) { out[i].pos = i; put(i, 0, one); }
        return true;
    }
    const bool down = src_n == 2 * dst_n;
    if (!down && dst_n != 2 * src_n) return false;
    if (down ? src_n < 12 : src_n < 6) return false;            // below that libswscale cuts its tap count to the plane: not covered
    const long long inc = (((long long)src_n << 16) + (dst_n >> 1)) / dst_n;
    const long long unit = 1LL << (down ? 53 : 54);              // fone = 1 << (54 - av_log2(src / dst))
    const int taps = std::min(down ? 9 : 5, src_n - 2);           // capped by the plane
    const long long Cq = (long long)(0.6 * (1 << 24));
    std::vector<long long> f((size_t)dst_n * taps);
    std::vector<int> pos(dst_n);
    long long at = ((128 * inc) >> 7) - ((128 * 0x10000LL) >> 7);
    for (int i = 0; i < dst_n; i++, at += 2 * inc) {
        int xx = (int)((at - (long long)(taps - 2) * (1LL << 16)) / (1 << 17));
        pos[i] = xx;
        for (int t = 0; t < taps; t++, xx++) {
            long long d = llabs((long long)xx * (1 << 17) - at) << 13;
            if (down) d = d * dst_n / src_n;
            long long c = 0;
            if (d < (1LL << 31)) {
                const long long dd = (d * d) >> 30, ddd = (dd * d) >> 30;
                c = d < (1LL << 30)
                        ? (12 * (1 << 24) - 6 * Cq) * ddd + (-18 * (1 << 24) + 6 * Cq) * dd + 6LL * (1 << 24) * (1LL << 30)
                        : -6 * Cq * ddd + 30 * Cq * dd - 48 * Cq * d + 24 * Cq * (1LL << 30);
            }
            f[(size_t)i * taps + t] = c / ((1LL << 54) / unit);
        }
    }
    const double cut = 0.002 * (double)unit;
    int support = 0;
    for (int i = dst_n - 1; i >= 0; i--) {
        long long *fi = &f[(size_t)i * taps];
        long long acc = 0;
        for (int t = 0; t < taps; t++) {
            acc += llabs(fi[0]);
            if ((double)acc > cut || (i < dst_n - 1 && pos[i] >= pos[i + 1])) break;
            memmove(fi, fi + 1, (taps - 1) * sizeof(*fi));
            fi[taps - 1] = 0;
            pos[i]++;
        }
        int n = taps;
        acc = 0;
        for (int t = taps - 1; t > 0; t--) {
            acc += llabs(fi[t]);
            if ((double)acc > cut) break;
            n--;
        }
        support = n > support ? n : support;
    }
    const int size = (support + align - 1) & ~(align - 1);
    if (size > 8 || size > src_n) return false;
    for (int i = 0; i < dst_n; i++) {
        long long t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int q = 0; q < size && q < taps; q++) t[q] = f[(size_t)i * taps + q];
        if (pos[i] < 0) {
            for (int q = 1; q < size; q++) { const int to = q + pos[i] > 0 ? q + pos[i] : 0; t[to] += t[q]; t[q] = 0; }
            pos[i] = 0;
        }
        if (pos[i] + size > src_n) {
            const int shift = pos[i] + size - src_n;
            long long acc = 0;
            for (int q = size - 1; q >= 0; q--) if (pos[i] + q >= src_n) { acc += t[q]; t[q] = 0; }
            for (int q = size - 1; q >= 0; q--) t[q] = q < shift ? 0 : t[q - shift];
            pos[i] -= shift;
            t[src_n - 1 - pos[i]] += acc;
        }
        long long sum = 0, err = 0;
        for (int q = 0; q < size; q++) sum += t[q];
        sum = (sum + one / 2) / one;
        if (!sum) sum = 1;
        out[i].pos = pos[i];
        for (int q = 0; q < size; q++) {
            const long long v = t[q] + err;
            const long long iv = v >= 0 ? (v + (sum >> 1)) / sum : (v - (sum >> 1)) / sum;
            put(i, q, (int)iv);
            err = v - iv * sum;
        }
    }
    return true;
}

static int to_int16(long long f)
{
    long long r = (f + (1 << 15)) >> 16;
    return (int)(r < -0x7FFF ? -0x7FFF : r > 0x7FFF ? 0x7FFF : r);
}

// ff_yuv2rgb_c_init_tables [libswscale/yuv2rgb.c] for brightness 0, contrast = saturation = 1.0 (codec.c:2141-2144)
static void colour_constants(DecConst &k, int avcol_spc, int fullrange, int rounder)
{
    // sws_getCoefficients(): {crv, cbu, cgu, cgv}; the switch of codec.c:2114-2140
    static const int coeffs[5][4] = {
        {104597, 132201, 25675, 53279},     // ITU601 / SMPTE170M / default
        {117489, 138438, 13975, 34925},     // ITU709
        {104448, 132798, 24759, 53109},     // FCC
        {117579, 136230, 16907, 35559},     // SMPTE240M
        {110013, 140363, 12277, 42626},     // BT2020
    };
    int row = 0;
    switch (avcol_spc) {
    case 1: row = 1; break;                 // AVCOL_SPC_BT709
    case 4: row = 2; break;                 // AVCOL_SPC_FCC
    case 7: row = 3; break;                 // AVCOL_SPC_SMPTE240M
    case 9: case 10: row = 4; break;        // AVCOL_SPC_BT2020_NCL / _CL
    default: row = 0;                       // BT470BG, SMPTE170M, anything else -> SWS_CS_DEFAULT
    }
    long long crv = coeffs[row][0], cbu = coeffs[row][1], cgu = -coeffs[row][2], cgv = -coeffs[row][3];
    long long cy = 1 << 16, oy = 0;
    if (!fullrange) { cy = (cy * 255) / 219; oy = 16 << 16; }
    else { crv = crv * 224 / 255; cbu = cbu * 224 / 255; cgu = cgu * 224 / 255; cgv = cgv * 224 / 255; }
    const int y_coeff = to_int16(cy << 13), y_off = to_int16(oy << 3);
    // ((y << 3) + 4 - y_off) * y_coeff >> 16 == (y * y_coeff + floor((4 - y_off) * y_coeff / 8)) >> 13 for integer y
    // rounder: the +4 (0.5 in 1/8 units) the vertical-filter writers add; the single-line writers of 4:2:2 pictures do not
    const int yk = (rounder - y_off) * y_coeff;
    k.yc = y_coeff; k.ykf = yk >= 0 ? yk / 8 : -((-yk + 7) / 8);
    k.vr = to_int16(crv * 8192); k.ub = to_int16(cbu * 8192); k.vg = to_int16(cgv * 8192); k.ug = to_int16(cgu * 8192);
    const int c0 = rounder - 1024;
    k.vr0 = c0 * k.vr; k.ub0 = c0 * k.ub; k.vg0 = c0 * k.vg; k.ug0 = c0 * k.ug;
    crv = (crv * 65536 + 0x8000) / cy; cbu = (cbu * 65536 + 0x8000) / cy;
    cgu = (cgu * 65536 + 0x8000) / cy; cgv = (cgv * 65536 + 0x8000) / cy;
    k.cy = (int)cy;
    k.bias = (int)((fullrange ? 384 : 326) * cy - (384LL << 16) - oy + 0x8000);
    k.crv = (int)crv; k.cbu = (int)cbu; k.cgu = (int)cgu; k.cgv = (int)cgv;
    k.crv9 = (int)(crv >> 9); k.cbu9 = (int)(cbu >> 9); k.cgu9 = (int)(cgu >> 9); k.cgv9 = (int)(cgv >> 9);
    // yuv2rgb_write_full: Y = ((y << 9) - y_offset) * y_coeff + (1 << 21), y_offset = to_int16(oy << 9); chroma terms (c - 128) << 9 times
    // the same 13-bit coefficients
    k.fy = y_coeff; k.fy0 = (1 << 21) - to_int16(oy << 9) * y_coeff;
    k.fvr = k.vr; k.fvg = k.vg; k.fug = k.ug; k.fub = k.ub;
}

struct Dec {
    Ctx *ctx = nullptr;
    int csp, flip, w, h, v422, v444;
    int up = 0;                 // 1: planar output with another chroma resolution (dec_resample_kernel), 2: 4:4:4 -> YUY2 / UYVY
    DecConst k;
    DecRow *d_rows = nullptr;
    DecTap8 *d_cols = nullptr, *d_vtaps = nullptr;
    // staging of the host-buffer entry
    uint8_t *d_src = nullptr, *d_dst = nullptr;
    size_t src_bytes = 0, dst_bytes = 0;
};

struct DecTables { std::vector<DecRow> rows; std::vector<DecTap8> cols, vtaps; };

// x264vfw_picture_get_size (codec.c:505-508) for the covered formats
static inline int64_t dec_picture_size(int i_out_csp, int w, int h)
{
    switch (i_out_csp & X264VFW_CUDA_CSP_MASK) {
    case X264VFW_CUDA_CSP_I420: case X264VFW_CUDA_CSP_YV12: case X264VFW_CUDA_CSP_NV12:
        return (int64_t)w * h + 2 * (int64_t)(w / 2) * (h / 2);
    case X264VFW_CUDA_CSP_YV16: case X264VFW_CUDA_CSP_YUYV: case X264VFW_CUDA_CSP_UYVY: return (int64_t)w * 2 * h;
    case X264VFW_CUDA_CSP_YV24: return (int64_t)w * 3 * h;
    case X264VFW_CUDA_CSP_BGR:  return (int64_t)((w * 3 + 3) & ~3) * h;
    case X264VFW_CUDA_CSP_BGRA: return (int64_t)w * 4 * h;
    default: return -1;
    }
}

// Everything x264vfw_init_sws_context decides, without a device: which path serves (picture format, output csp), the colour
// constants and the filter tables.  0, or -1 with the reason in the thread's error string.
static int dec_configure(Dec &d, DecTables &t, int i_out_csp, int w, int h, int i_src_chroma, int i_avcol_spc, int b_fullrange)
{
    const int csp = i_out_csp & X264VFW_CUDA_CSP_MASK, flip = (i_out_csp & X264VFW_CUDA_CSP_VFLIP) != 0;
    if (w <= 0 || h <= 0 || (w & 1) || (h & 1)) { set_error("width/height must be positive and even (codec.c:1950-1954)"); return -1; }
    if (i_src_chroma < 1 || i_src_chroma > 3) { set_error("decoder pictures must be 4:2:0 (1), 4:2:2 (2) or 4:4:4 (3)"); return -1; }
    if (dec_picture_size(csp, w, h) < 0) { set_error("output csp %d is not covered", csp); return -1; }
    const bool v422 = i_src_chroma == 2, v444 = i_src_chroma == 3;
    const bool out420 = csp == X264VFW_CUDA_CSP_I420 || csp == X264VFW_CUDA_CSP_YV12 || csp == X264VFW_CUDA_CSP_NV12;
    const int out_chroma = out420 ? 1 : csp == X264VFW_CUDA_CSP_YV16 || csp == X264VFW_CUDA_CSP_YUYV || csp == X264VFW_CUDA_CSP_UYVY ? 2 :
                           csp == X264VFW_CUDA_CSP_YV24 ? 3 : 0;
    const bool planar_out = out420 || csp == X264VFW_CUDA_CSP_YV16 || csp == X264VFW_CUDA_CSP_YV24;
    // 1: planar output with another chroma resolution than the picture; 2: 4:4:4 -> YUY2 / UYVY (libswscale resamples the chroma)
    const int up = planar_out && out_chroma != i_src_chroma ? 1 : v444 && out_chroma == 2 ? 2 : 0;
    const bool rgb = csp == X264VFW_CUDA_CSP_BGR || csp == X264VFW_CUDA_CSP_BGRA;
    if (flip && !rgb) { set_error("only RGB output can be bottom-up (codec.c:510-527)"); return -1; }
    const bool planar = planar_out || v444;               // no packed-writer row table: plane copies, resampled planes, the per-pixel 4:4:4 writer
    d.csp = csp; d.flip = flip; d.w = w; d.h = h; d.v422 = v422; d.v444 = v444; d.up = up;
    colour_constants(d.k, i_avcol_spc, b_fullrange != 0, v422 ? 0 : 4);
    if (up) {
        const int cw = v444 ? w : w / 2, ch = v422 || v444 ? h : h / 2;
        const int ocw = csp == X264VFW_CUDA_CSP_YV24 ? w : w / 2, och = up == 2 || !out420 ? h : h / 2;
        if (!resample_taps(cw, ocw, 1 << 14, 4, t.cols) || !resample_taps(ch, och, 1 << 12, 2, t.vtaps)) {
            set_error("picture too small for libswscale's chroma filters (12 samples to halve, 6 to double)");
            return -1;
        }
    } else if (!planar) {
        if (v422 && h >= 12) {
            // no vertical filter (libswscale's yuv2packed1 writers; plain interleave for 4:2:2 output): one tap of 1.0 on the
            // row's own chroma line, addressed inside the 4-line window the kernel loads for the row pair
            t.rows.resize(h);
            for (int r = 0; r < h; r++) {
                const int k = (r + 1) >> 1, pos = std::min(std::max(2 * k - 1, 0), h - 4), tp = r - pos;
                t.rows[r].pos = pos;
                t.rows[r].c01 = tp == 0 ? 4096 : tp == 1 ? (4096 << 16) : 0;
                t.rows[r].c23 = tp == 2 ? 4096 : tp == 3 ? (4096 << 16) : 0;
                t.rows[r].c_writer = csp == X264VFW_CUDA_CSP_UYVY || r >= h - 2;
            }
        } else if (v422 || !vertical_chroma_filter(h / 2, t.rows, csp == X264VFW_CUDA_CSP_UYVY)) {
            set_error("pictures below 12 rows are not covered");
            return -1;
        }
    }
    return 0;
}

static inline bool al(const void *p, size_t a) { return ((uintptr_t)p & (a - 1)) == 0; }
static inline bool als(long long v, long long a) { return (v & (a - 1)) == 0; }

static int dec_launch(Dec *d, DEC_STREAM st, uint8_t *dst, size_t dfb, const uint8_t *const src[3], const int ss[3], size_t sfb, int n)
{
    const int w = d->w, h = d->h, cw = d->v444 ? w : w / 2, ch = d->v422 || d->v444 ? h : h / 2;
    if (n <= 0) return 0;
    if (n > 65535) { set_error("at most 65535 pictures per launch"); return -1; }
    if (d->up == 2) {
        Dec444PackedJob pj;
        pj.y = src[0]; pj.u = src[1]; pj.v = src[2]; pj.ys = ss[0]; pj.us = ss[1]; pj.vs = ss[2];
        pj.dst = dst; pj.w = w; pj.h = h; pj.uyvy = d->csp == X264VFW_CUDA_CSP_UYVY; pj.cols = d->d_cols;
        pj.src_frame_bytes = sfb; pj.dst_frame_bytes = dfb;
        if (!al(dst, 4) || !als((long long)dfb, 4)) { set_error("output picture must be 4-byte aligned"); return -1; }
        DEC_LAUNCH(dim3((w / 2 + 31) / 32, (h + 7) / 8, n), dim3(32, 8), st, pj, dec_444_packed_kernel);
        XV_LAUNCH_CHECK();
        return 0;
    }
    if (d->up) {
        // luma: plane copy (the planar kernel with no chroma rows in its grid); chroma: the resampling kernel
        const bool o420 = d->csp == X264VFW_CUDA_CSP_I420 || d->csp == X264VFW_CUDA_CSP_YV12 || d->csp == X264VFW_CUDA_CSP_NV12;
        const int ocw = d->csp == X264VFW_CUDA_CSP_YV24 ? w : w / 2, och = o420 ? h / 2 : h;
        DecPlanarJob pj;
        pj.y = src[0]; pj.u = src[1]; pj.v = src[2]; pj.ys = ss[0]; pj.us = ss[1]; pj.vs = ss[2];
        pj.w = w; pj.h = h; pj.cw = cw; pj.src_frame_bytes = sfb; pj.dst_frame_bytes = dfb;
        pj.dy = dst; pj.du = pj.dv = nullptr;
        const bool vec = als(w, 16) && al(dst, 16) && als((long long)dfb, 16) && als((long long)sfb, 16) && al(src[0], 16) && als(ss[0], 16);
        dim3 grid((w + 4095) / 4096, h, n);
        if (vec) DEC_LAUNCH(grid, dim3(256), st, pj, dec_planar_kernel<true>);
        else     DEC_LAUNCH(grid, dim3(256), st, pj, dec_planar_kernel<false>);
        XV_LAUNCH_CHECK();
        DecUpJob uj;
        uj.u = src[1]; uj.v = src[2]; uj.us = ss[1]; uj.vs = ss[2];
        uint8_t *p1 = dst + (size_t)w * h, *p2 = p1 + (size_t)ocw * och;
        if (d->csp == X264VFW_CUDA_CSP_NV12) { uj.du = p1; uj.dv = p1 + 1; uj.ostride = w; uj.ostep = 2; }
        else {
            uj.ostride = ocw; uj.ostep = 1;
            if (d->csp == X264VFW_CUDA_CSP_I420) { uj.du = p1; uj.dv = p2; } else { uj.du = p2; uj.dv = p1; }   // YV12 / YV16 / YV24: codec.c:2263-2274
        }
        uj.ocw = ocw; uj.och = och; uj.cols = d->d_cols; uj.rows = d->d_vtaps;
        uj.src_frame_bytes = sfb; uj.dst_frame_bytes = dfb;
        DEC_LAUNCH(dim3((ocw + 31) / 32, (och + 7) / 8, n), dim3(32, 8), st, uj, dec_resample_kernel);
        XV_LAUNCH_CHECK();
        return 0;
    }
    if (d->csp == X264VFW_CUDA_CSP_I420 || d->csp == X264VFW_CUDA_CSP_YV12 || d->csp == X264VFW_CUDA_CSP_NV12 || d->csp == X264VFW_CUDA_CSP_YV16 ||
        d->csp == X264VFW_CUDA_CSP_YV24) {
        DecPlanarJob j;
        j.y = src[0]; j.u = src[1]; j.v = src[2]; j.ys = ss[0]; j.us = ss[1]; j.vs = ss[2];
        j.w = w; j.h = h; j.cw = cw; j.src_frame_bytes = sfb; j.dst_frame_bytes = dfb;
        j.dy = dst;
        uint8_t *p1 = dst + (size_t)w * h, *p2 = p1 + (size_t)cw * ch;             // x264vfw_picture_fill, codec.c:425-439,469-480
        if (d->csp == X264VFW_CUDA_CSP_NV12) { j.du = p1; j.dv = nullptr; }
        else if (d->csp != X264VFW_CUDA_CSP_I420) { j.du = p2; j.dv = p1; }             // YV12 / YV16 / YV24: codec.c:2263-2274
        else { j.du = p1; j.dv = p2; }
        const bool vec = als(w, 32) && al(dst, 16) && als((long long)dfb, 16) && als((long long)sfb, 16) &&
                         al(src[0], 16) && al(src[1], 16) && al(src[2], 16) && als(ss[0], 16) && als(ss[1], 16) && als(ss[2], 16) &&
                         als((long long)cw * ch, 16);
        dim3 grid((w + 4095) / 4096, h + ch, n);
        if (vec) DEC_LAUNCH(grid, dim3(256), st, j, dec_planar_kernel<true>);
        else     DEC_LAUNCH(grid, dim3(256), st, j, dec_planar_kernel<false>);
        XV_LAUNCH_CHECK();
        return 0;
    }
    DecJob j;
    j.y = src[0]; j.u = src[1]; j.v = src[2]; j.ys = ss[0]; j.us = ss[1]; j.vs = ss[2];
    j.w = w; j.h = h; j.v422 = d->v422; j.src_frame_bytes = sfb; j.dst_frame_bytes = dfb;
    j.rows = d->d_rows; j.k = d->k;
    long long stride = d->csp == X264VFW_CUDA_CSP_BGR ? ((w * 3 + 3) & ~3) : d->csp == X264VFW_CUDA_CSP_BGRA ? w * 4 : w * 2;
    j.dst = dst; j.dst_stride = stride;
    if (d->flip) { j.dst = dst + stride * (h - 1); j.dst_stride = -stride; }       // codec.c:515-518
    if (d->v444) {
        const bool bgra = d->csp == X264VFW_CUDA_CSP_BGRA;
        const size_t a = bgra ? 16 : 4;
        const bool v4 = als(w, 8) && al(dst, a) && als(stride, a) && als((long long)dfb, a) && als((long long)sfb, 8) &&
                        al(src[0], 8) && al(src[1], 8) && al(src[2], 8) && als(ss[0], 8) && als(ss[1], 8) && als(ss[2], 8);
        if (!al(dst, 4) || !als((long long)dfb, 4)) { set_error("output picture must be 4-byte aligned"); return -1; }
        dim3 blk(32, 8), grd((w + 255) / 256, (h + 7) / 8, n);
        if (bgra) { if (v4) DEC_LAUNCH(grd, blk, st, j, dec_444_kernel<true, true>); else DEC_LAUNCH(grd, blk, st, j, dec_444_kernel<true, false>); }
        else      { if (v4) DEC_LAUNCH(grd, blk, st, j, dec_444_kernel<false, true>); else DEC_LAUNCH(grd, blk, st, j, dec_444_kernel<false, false>); }
        XV_LAUNCH_CHECK();
        return 0;
    }
    const size_t da = d->csp == X264VFW_CUDA_CSP_BGR ? 4 : 16;
    const bool vec = als(w, 8) && al(dst, da) && als(stride, da) && als((long long)dfb, da) && als((long long)sfb, 8) &&
                     al(src[0], 8) && als(ss[0], 8) && al(src[1], 4) && al(src[2], 4) && als(ss[1], 4) && als(ss[2], 4);
    if (!al(dst, 4) || !als((long long)dfb, 4)) { set_error("output picture must be 4-byte aligned"); return -1; }
    dim3 block(32, 8), grid((w + 255) / 256, (h / 2 + 1 + 7) / 8, n);
#define DEC_GO(F) do { if (vec) DEC_LAUNCH(grid, block, st, j, dec_packed_kernel<F, true>); \
                       else     DEC_LAUNCH(grid, block, st, j, dec_packed_kernel<F, false>); } while (0)
    switch (d->csp) {
    case X264VFW_CUDA_CSP_BGRA: DEC_GO(DEC_BGRA); break;
    case X264VFW_CUDA_CSP_BGR:  DEC_GO(DEC_BGR); break;
    case X264VFW_CUDA_CSP_YUYV: DEC_GO(DEC_YUYV); break;
    default:                    DEC_GO(DEC_UYVY); break;
    }
#undef DEC_GO
    XV_LAUNCH_CHECK();
    return 0;
}

} // namespace xv
