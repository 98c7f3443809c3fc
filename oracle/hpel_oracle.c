/*
 * hpel_oracle.c -- scalar restatement of libx264's half-pel reference planes.
 * TEST INFRASTRUCTURE (see oracle.h).  Upstream libx264 is not in the reference tree, so this is NOT pinned
 * against libx264; it IS pinned against the H.264 standard's fractional sample interpolation as libavcodec's
 * decoder evaluates it (tests/test_h264_pins.py, fixtures tests/golden/h264_pins.json: all quarter-sample
 * phases, vectors up to 22 samples outside the picture, random and 0/255 content) -- an encoder's reference
 * planes have to give the decoder's prediction.  What that leaves unpinned: the layout (stride, 32-sample
 * border) and the planes' bytes further than 22 samples outside the frame.
 * It follows upstream's published C algorithm by function name:
 *   [x264] common/frame.c  x264_frame_expand_border (plane_expand_border), x264_frame_expand_border_filtered
 *   [x264] common/mc.c     hpel_filter (C version, 8-bit: pad = 0), x264_frame_filter (progressive)
 * It goes through upstream's literal sequence (materialised 32-pixel border, filter over the frame plus
 * 8 pixels, second border pass that starts 4 / 8 pixels outside the frame) so that the CUDA kernel,
 * which folds all of it into clamped addressing, is checked against the real order of operations.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

#define PADH 32
#define PADV 32

void orc_hpel_geometry(int w, int h, int g[3])
{
    g[0] = (w + 2 * PADH + 63) & ~63;          /* stride                       */
    g[1] = g[0] * (h + 2 * PADV);              /* bytes of one padded plane    */
    g[2] = PADV * g[0] + PADH;                 /* offset of pixel (0,0)        */
}

/* [x264] common/frame.c plane_expand_border, b_chroma = 0, 8-bit */
static void plane_expand_border(uint8_t *pix, int stride, int width, int height, int padh, int padv,
                                int pad_top, int pad_bottom)
{
    for (int y = 0; y < height; y++) {
        memset(pix + (ptrdiff_t)y * stride - padh, pix[(ptrdiff_t)y * stride], padh);
        memset(pix + (ptrdiff_t)y * stride + width, pix[(ptrdiff_t)y * stride + width - 1], padh);
    }
    if (pad_top)
        for (int y = 0; y < padv; y++)
            memcpy(pix - (ptrdiff_t)(y + 1) * stride - padh, pix - padh, width + 2 * padh);
    if (pad_bottom)
        for (int y = 0; y < padv; y++)
            memcpy(pix + (ptrdiff_t)(height + y) * stride - padh, pix + (ptrdiff_t)(height - 1) * stride - padh, width + 2 * padh);
}

static inline uint8_t clip_pixel(int x) { return (x & ~255) ? (uint8_t)((-x) >> 31) : (uint8_t)x; }

#define TAPFILTER(pix, d) ((pix)[x - 2 * (d)] + (pix)[x + 3 * (d)] - 5 * ((pix)[x - (d)] + (pix)[x + 2 * (d)]) + 20 * ((pix)[x] + (pix)[x + (d)]))

/* [x264] common/mc.c hpel_filter */
static void hpel_filter(uint8_t *dsth, uint8_t *dstv, uint8_t *dstc, const uint8_t *src,
                        ptrdiff_t stride, int width, int height, int16_t *buf)
{
    for (int y = 0; y < height; y++) {
        for (int x = -2; x < width + 3; x++) {
            int v = TAPFILTER(src, stride);
            dstv[x] = clip_pixel((v + 16) >> 5);
            buf[x + 2] = (int16_t)v;
        }
        for (int x = 0; x < width; x++)
            dstc[x] = clip_pixel((TAPFILTER(buf + 2, 1) + 512) >> 10);
        for (int x = 0; x < width; x++)
            dsth[x] = clip_pixel((TAPFILTER(src, 1) + 16) >> 5);
        dsth += stride; dstv += stride; dstc += stride; src += stride;
    }
}

/* The reference planes of one reconstructed frame: dst = 4 padded planes (orc_hpel_geometry),
 * [0] = the frame with x264_frame_expand_border's 32-pixel border, [1..3] = filtered H, V, C after
 * x264_frame_expand_border_filtered.  src = tight w x h plane (w, h = 16*mb_w, 16*mb_h upstream). */
void orc_hpel_planes(uint8_t *dst4planes, const uint8_t *src, int src_stride, int w, int h)
{
    int g[3];
    orc_hpel_geometry(w, h, g);
    const int stride = g[0];
    uint8_t *p[4];
    for (int i = 0; i < 4; i++) p[i] = dst4planes + (size_t)i * g[1] + g[2];
    memset(dst4planes, 0xA5, (size_t)4 * g[1]);             /* anything not written below shows up */

    /* x264_frame_expand_border (whole frame, progressive) */
    for (int y = 0; y < h; y++) memcpy(p[0] + (ptrdiff_t)y * stride, src + (size_t)y * src_stride, w);
    plane_expand_border(p[0], stride, w, h, PADH, PADV, 1, 1);

    /* x264_frame_filter( h, frame, mb_y = 0, b_end = 1 ): start = -8, height = lines + 8 */
    const int start = -8, height = h + 8;
    const ptrdiff_t offs = (ptrdiff_t)start * stride - 8;
    int16_t *buf = (int16_t *)malloc(sizeof(int16_t) * (size_t)(w + 16 + 5 + 16));
    hpel_filter(p[1] + offs, p[2] + offs, p[3] + offs, p[0] + offs, stride, w + 16, height - start, buf);
    free(buf);

    /* x264_frame_expand_border_filtered( h, frame, mb_y = 0, b_end = 1 ) */
    for (int i = 1; i < 4; i++)
        plane_expand_border(p[i] - (ptrdiff_t)8 * stride - 4, stride, w + 8, h + 16, PADH - 4, PADV - 8, 1, 1);
}
