/*
 * csp_oracle.c -- scalar restatement of the reference colour-space converters.
 * TEST INFRASTRUCTURE (see oracle.h).  Written as table-driven plain loops rather than the
 * reference's macro families; each function cites the reference lines it follows.
 */
#include "oracle.h"
#include <string.h>

enum { IN_I420 = 1, IN_YV12, IN_YV16, IN_YV24, IN_NV12, IN_YUYV, IN_UYVY, IN_BGR, IN_BGRA };   /* csp.h:33-43 */
enum { OUT_I420 = 2, OUT_NV12 = 4, OUT_I422 = 6, OUT_I444 = 0xc, OUT_BGR = 0xe, OUT_BGRA = 0xf };
#define VFLIP 0x1000

uint64_t orc_fnv1a64(const uint8_t *p, size_t n)
{
    uint64_t h = 0xcbf29ce484222325ull;
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 0x100000001b3ull; }
    return h;
}

void orc_lcg_fill(uint8_t *p, size_t n, int w, int h)
{
    uint32_t s = 0x264u + 31u * (uint32_t)w + (uint32_t)h;
    for (size_t i = 0; i < n; i++) { s = s * 1664525u + 1013904223u; p[i] = (uint8_t)(s >> 24); }
}

/* A source plane seen top-down: flipping = start at the last row, walk backwards
 * (csp.c:75-91 for planes, :166-170 / :310-314 for packed inputs). */
typedef struct { const uint8_t *p; ptrdiff_t stride; } view_t;
static view_t view(const uint8_t *p, int stride, int rows, int flip)
{
    view_t v = { p, stride };
    if (flip) { v.p = p + (ptrdiff_t)(rows - 1) * stride; v.stride = -stride; }
    return v;
}

/* csp.c:28-37 */
static void rows_copy(uint8_t *d, int ds, view_t s, int bytes, int rows)
{
    for (int y = 0; y < rows; y++) memcpy(d + (size_t)y * ds, s.p + y * s.stride, bytes);
}
/* csp.c:39-55: vertical 2:1, rounded mean */
static void rows_v2(uint8_t *d, int ds, view_t s, int w, int rows)
{
    for (int y = 0; y < rows; y++) {
        const uint8_t *a = s.p + (2 * y) * s.stride, *b = a + s.stride;
        for (int x = 0; x < w; x++) d[(size_t)y * ds + x] = (uint8_t)((a[x] + b[x] + 1) >> 1);
    }
}
/* csp.c:57-73: 2x2 box, +2 >> 2 */
static void rows_hv2(uint8_t *d, int ds, view_t s, int w, int rows)
{
    for (int y = 0; y < rows; y++) {
        const uint8_t *a = s.p + (2 * y) * s.stride, *b = a + s.stride;
        for (int x = 0; x < w; x++)
            d[(size_t)y * ds + x] = (uint8_t)((a[2 * x] + a[2 * x + 1] + b[2 * x] + b[2 * x + 1] + 2) >> 2);
    }
}

/* csp.c:252-297.  FIX(f) = (uint32_t)(f * 2^20 + 0.5), evaluated in double like the C
 * preprocessor-constant expressions of the reference. */
void orc_rgb_coefficients(int colmatrix, int fullrange, uint32_t o[12])
{
    const double kb = colmatrix == 1 ? 0.0722 : 0.114, kr = colmatrix == 1 ? 0.2126 : 0.299;
    const double kg = 1.0 - kb - kr, sb = 1.0 - kb, sr = 1.0 - kr;
    const double one = 1048576.0;
    const double ky = fullrange ? 1.0 : 1.0 * 219.0 / 255.0;
    const double ku = fullrange ? (0.5 / sb) : (0.5 / sb) * 224.0 / 255.0;
    const double kv = fullrange ? (0.5 / sr) : (0.5 / sr) * 224.0 / 255.0;
    const double ay = fullrange ? 0.0 : 16.0;
    const int bias = fullrange ? -1 : 0;
    o[0] = (uint32_t)(kr * ky * one + 0.5); o[1] = (uint32_t)(kg * ky * one + 0.5); o[2] = (uint32_t)(kb * ky * one + 0.5);
    o[3] = (uint32_t)(ay * one + 524288 + 0.5);
    o[4] = (uint32_t)(kr * ku * one + 0.5); o[5] = (uint32_t)(kg * ku * one + 0.5); o[6] = (uint32_t)(sb * ku * one + 0.5);
    o[7] = (uint32_t)((128.0 * one + 524288) * 4 + bias + 0.5);
    o[8] = (uint32_t)(sr * kv * one + 0.5); o[9] = (uint32_t)(kg * kv * one + 0.5); o[10] = (uint32_t)(kb * kv * one + 0.5);
    o[11] = (uint32_t)((128.0 * one + 524288) * 4 + bias + 0.5);
}

/* csp.c:299-388 (RGB_TO_I420).  bpp = 3 (bgr) or 4 (bgra); byte order B,G,R[,X]. */
static void rgb_to_planes(const uint32_t c[12], view_t s, int bpp, int w, int h,
                          uint8_t *Y, int ys, uint8_t *U, int us, int ustep, uint8_t *V, int vs, int vstep)
{
    for (int y = 0; y < h; y += 2) {
        const uint8_t *r0 = s.p + y * s.stride, *r1 = r0 + s.stride;
        for (int x = 0; x < w; x += 2) {
            uint32_t sb = 0, sg = 0, sr = 0;
            for (int dx = 0; dx < 2; dx++)
                for (int dy = 0; dy < 2; dy++) {
                    const uint8_t *px = (dy ? r1 : r0) + (x + dx) * bpp;
                    uint32_t b = px[0], g = px[1], r = px[2];
                    sb += b; sg += g; sr += r;
                    Y[(size_t)(y + dy) * ys + x + dx] = (uint8_t)((c[3] + c[0] * r + c[1] * g + c[2] * b) >> 20);
                }
            U[(size_t)(y >> 1) * us + (x >> 1) * ustep] = (uint8_t)((c[7] + c[6] * sb - c[4] * sr - c[5] * sg) >> 22);
            V[(size_t)(y >> 1) * vs + (x >> 1) * vstep] = (uint8_t)((c[11] + c[8] * sr - c[9] * sg - c[10] * sb) >> 22);
        }
    }
}

/* csp.c:155-250 (YYUV_TO_I420 / YYUV_TO_I422).  yoff: byte of Y0 in the 4-byte group. */
static void packed422_to_planes(view_t s, int uyvy, int to420, int w, int h,
                                uint8_t *Y, int ys, uint8_t *U, int us, uint8_t *V, int vs)
{
    const int y0 = uyvy ? 1 : 0, y1 = uyvy ? 3 : 2, uo = uyvy ? 0 : 1, vo = uyvy ? 2 : 3;
    for (int y = 0; y < h; y++) {
        const uint8_t *r = s.p + y * s.stride;
        for (int x = 0; x < w; x += 2) {
            const uint8_t *g = r + 2 * x;
            Y[(size_t)y * ys + x] = g[y0];
            Y[(size_t)y * ys + x + 1] = g[y1];
            if (!to420) {
                U[(size_t)y * us + (x >> 1)] = g[uo];
                V[(size_t)y * vs + (x >> 1)] = g[vo];
            } else if (!(y & 1)) {
                const uint8_t *g2 = g + s.stride;
                U[(size_t)(y >> 1) * us + (x >> 1)] = (uint8_t)((g[uo] + g2[uo] + 1) >> 1);
                V[(size_t)(y >> 1) * vs + (x >> 1)] = (uint8_t)((g[vo] + g2[vo] + 1) >> 1);
            }
        }
    }
}

int orc_csp_convert(int out_csp, int colmatrix, int fullrange,
                    orc_image_t *dst, const orc_image_t *src, int w, int h)
{
    const int in = src->i_csp & 0xff, flip = !!(src->i_csp & VFLIP);
    uint32_t c[12];
    switch (out_csp) {
    case OUT_I420:                                                   /* csp.c:447-488 */
        switch (in) {
        case IN_I420: case IN_YV12: case IN_YV16: case IN_YV24: {
            const int swap = in != IN_I420;                          /* csp.c:409-414 */
            rows_copy(dst->plane[0], dst->i_stride[0], view(src->plane[0], src->i_stride[0], h, flip), w, h);
            for (int k = 1; k <= 2; k++) {
                uint8_t *d = dst->plane[swap ? 3 - k : k]; int ds = dst->i_stride[swap ? 3 - k : k];
                if (in == IN_YV16)      rows_v2(d, ds, view(src->plane[k], src->i_stride[k], h, flip), w / 2, h / 2);
                else if (in == IN_YV24) rows_hv2(d, ds, view(src->plane[k], src->i_stride[k], h, flip), w / 2, h / 2);
                else                    rows_copy(d, ds, view(src->plane[k], src->i_stride[k], h / 2, flip), w / 2, h / 2);
            }
            return 0;
        }
        case IN_YUYV: case IN_UYVY:                                  /* csp.c:422-423 */
            packed422_to_planes(view(src->plane[0], src->i_stride[0], h, flip), in == IN_UYVY, 1, w, h,
                                dst->plane[0], dst->i_stride[0], dst->plane[1], dst->i_stride[1], dst->plane[2], dst->i_stride[2]);
            return 0;
        case IN_BGR: case IN_BGRA:                                   /* csp.c:428-435, 456-487 */
            orc_rgb_coefficients(colmatrix, fullrange, c);
            rgb_to_planes(c, view(src->plane[0], src->i_stride[0], h, flip), in == IN_BGRA ? 4 : 3, w, h,
                          dst->plane[0], dst->i_stride[0], dst->plane[1], dst->i_stride[1], 1, dst->plane[2], dst->i_stride[2], 1);
            return 0;
        }
        return -1;
    case OUT_NV12:                                                   /* csp.c:490-492, 420 */
        if (in != IN_NV12) return -1;
        rows_copy(dst->plane[0], dst->i_stride[0], view(src->plane[0], src->i_stride[0], h, flip), w, h);
        rows_copy(dst->plane[1], dst->i_stride[1], view(src->plane[1], src->i_stride[1], h / 2, flip), w, h / 2);
        return 0;
    case OUT_I422:                                                   /* csp.c:494-499 */
        if (in == IN_YV16) {
            rows_copy(dst->plane[0], dst->i_stride[0], view(src->plane[0], src->i_stride[0], h, flip), w, h);
            rows_copy(dst->plane[2], dst->i_stride[2], view(src->plane[1], src->i_stride[1], h, flip), w / 2, h);
            rows_copy(dst->plane[1], dst->i_stride[1], view(src->plane[2], src->i_stride[2], h, flip), w / 2, h);
            return 0;
        }
        if (in == IN_YUYV || in == IN_UYVY) {                        /* csp.c:425-426 */
            packed422_to_planes(view(src->plane[0], src->i_stride[0], h, flip), in == IN_UYVY, 0, w, h,
                                dst->plane[0], dst->i_stride[0], dst->plane[1], dst->i_stride[1], dst->plane[2], dst->i_stride[2]);
            return 0;
        }
        return -1;
    case OUT_I444:                                                   /* csp.c:501-504 */
        if (in != IN_YV24) return -1;
        rows_copy(dst->plane[0], dst->i_stride[0], view(src->plane[0], src->i_stride[0], h, flip), w, h);
        rows_copy(dst->plane[2], dst->i_stride[2], view(src->plane[1], src->i_stride[1], h, flip), w, h);
        rows_copy(dst->plane[1], dst->i_stride[1], view(src->plane[2], src->i_stride[2], h, flip), w, h);
        return 0;
    case OUT_BGR:                                                    /* csp.c:506-508, 437 */
        if (in != IN_BGR) return -1;
        rows_copy(dst->plane[0], dst->i_stride[0], view(src->plane[0], src->i_stride[0], h, flip), w * 3, h);
        return 0;
    case OUT_BGRA:                                                   /* csp.c:510-512, 438 */
        if (in != IN_BGRA) return -1;
        rows_copy(dst->plane[0], dst->i_stride[0], view(src->plane[0], src->i_stride[0], h, flip), w * 4, h);
        return 0;
    }
    return -1;
}

/* Extension (no reference path): bgr/bgra -> I420 arithmetic of csp.c:299-388
 * with U/V written interleaved, i.e. what [x264] x264_frame_copy_picture makes of the
 * reference's I420 result. */
int orc_ext_rgb_to_nv12(int colmatrix, int fullrange, orc_image_t *dst, const orc_image_t *src, int w, int h)
{
    const int in = src->i_csp & 0xff, flip = !!(src->i_csp & VFLIP);
    uint32_t c[12];
    if (in != IN_BGR && in != IN_BGRA) return -1;
    orc_rgb_coefficients(colmatrix, fullrange, c);
    rgb_to_planes(c, view(src->plane[0], src->i_stride[0], h, flip), in == IN_BGRA ? 4 : 3, w, h,
                  dst->plane[0], dst->i_stride[0], dst->plane[1], dst->i_stride[1], 2, dst->plane[1] + 1, dst->i_stride[1], 2);
    return 0;
}

/* Extension (no reference path): yuyv/uyvy -> the reference's I422 samples (csp.c:209-250),
 * each chroma sample replicated to two horizontal positions. */
int orc_ext_422_to_i444(orc_image_t *dst, const orc_image_t *src, int w, int h)
{
    const int in = src->i_csp & 0xff, flip = !!(src->i_csp & VFLIP);
    if (in != IN_YUYV && in != IN_UYVY) return -1;
    view_t s = view(src->plane[0], src->i_stride[0], h, flip);
    const int uyvy = in == IN_UYVY;
    for (int y = 0; y < h; y++) {
        const uint8_t *r = s.p + y * s.stride;
        for (int x = 0; x < w; x += 2) {
            const uint8_t *g = r + 2 * x;
            dst->plane[0][(size_t)y * dst->i_stride[0] + x] = g[uyvy ? 1 : 0];
            dst->plane[0][(size_t)y * dst->i_stride[0] + x + 1] = g[uyvy ? 3 : 2];
            dst->plane[1][(size_t)y * dst->i_stride[1] + x] = dst->plane[1][(size_t)y * dst->i_stride[1] + x + 1] = g[uyvy ? 0 : 1];
            dst->plane[2][(size_t)y * dst->i_stride[2] + x] = dst->plane[2][(size_t)y * dst->i_stride[2] + x + 1] = g[uyvy ? 2 : 3];
        }
    }
    return 0;
}
