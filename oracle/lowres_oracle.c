/*
 * lowres_oracle.c -- scalar restatement of libx264's half-resolution plane construction.
 * TEST INFRASTRUCTURE (see oracle.h).  PARITY UNPINNED: upstream libx264 is not in the
 * reference tree; this follows upstream's published C algorithm by function name:
 *   [x264] common/frame.c  x264_frame_expand_border_mod16, x264_frame_expand_border_lowres
 *   [x264] common/mc.c     x264_frame_init_lowres, frame_init_lowres_core
 * It deliberately goes through the same intermediate steps as upstream (materialised
 * mod-16 plane, duplicated row/column, separate border pass) so that the CUDA kernel, which
 * folds all of them into clamped addressing, is checked against the literal sequence.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

#define PAD 32

void orc_lowres_geometry(int w, int h, int g[10])
{
    int mb_w = (w + 15) >> 4, mb_h = (h + 15) >> 4;
    g[0] = mb_w; g[1] = mb_h; g[2] = 16 * mb_w; g[3] = 16 * mb_h;
    g[4] = (g[2] + 1 + 63) & ~63;
    g[5] = g[2] / 2; g[6] = g[3] / 2;
    g[7] = (g[5] + 2 * PAD + 63) & ~63;
    g[8] = g[7] * (g[6] + 2 * PAD);
    g[9] = PAD * g[7] + PAD;
}

/* step 1: [x264] x264_frame_copy_picture (luma) + x264_frame_expand_border_mod16 + the
 * duplicated column/row at the start of x264_frame_init_lowres.  dst: stride x (luma_h+1). */
void orc_luma_pad(uint8_t *dst, int dst_stride, const uint8_t *y, int y_stride, int w, int h)
{
    int g[10];
    orc_lowres_geometry(w, h, g);
    const int lw = g[2], lh = g[3];
    for (int r = 0; r < h; r++) {
        memcpy(dst + (size_t)r * dst_stride, y + (size_t)r * y_stride, w);
        if (lw > w) memset(dst + (size_t)r * dst_stride + w, dst[(size_t)r * dst_stride + w - 1], lw - w);
    }
    for (int r = h; r < lh; r++) memcpy(dst + (size_t)r * dst_stride, dst + (size_t)(h - 1) * dst_stride, lw);
    for (int r = 0; r < lh; r++) dst[(size_t)r * dst_stride + lw] = dst[(size_t)r * dst_stride + lw - 1];
    memcpy(dst + (size_t)lh * dst_stride, dst + (size_t)(lh - 1) * dst_stride, lw + 1);
}

/* [x264] common/frame.c: x264_frame_copy_picture, 4:2:0 chroma (common/mc.c: plane_copy_interleave)
 * followed by x264_frame_expand_border_mod16 for the chroma plane: the last U/V pair is repeated to
 * the right, the last row downwards.  dst: dst_stride x (luma_h/2), luma_w bytes per row used. */
void orc_chroma_nv12_pad(uint8_t *dst, int dst_stride, const uint8_t *u, const uint8_t *v, int c_stride, int w, int h)
{
    int g[10];
    orc_lowres_geometry(w, h, g);
    const int lw = g[2], lh = g[3];
    const int cw = w / 2, ch = h / 2;
    for (int r = 0; r < ch; r++) {
        uint8_t *d = dst + (size_t)r * dst_stride;
        for (int x = 0; x < cw; x++) { d[2 * x] = u[(size_t)r * c_stride + x]; d[2 * x + 1] = v[(size_t)r * c_stride + x]; }
        for (int x = cw; x < lw / 2; x++) { d[2 * x] = d[2 * cw - 2]; d[2 * x + 1] = d[2 * cw - 1]; }
    }
    for (int r = ch; r < lh / 2; r++) memcpy(dst + (size_t)r * dst_stride, dst + (size_t)(ch - 1) * dst_stride, lw);
}

static inline int filt(int a, int b, int c, int d) { return (((a + b + 1) >> 1) + ((c + d + 1) >> 1) + 1) >> 1; }

/* step 2+3: frame_init_lowres_core over the padded luma, then 32-px edge replication.
 * dst: 4 consecutive padded planes (geometry of orc_lowres_geometry). */
void orc_lowres_init(uint8_t *dst, const uint8_t *y, int y_stride, int w, int h)
{
    int g[10];
    orc_lowres_geometry(w, h, g);
    const int ls = g[4], lw = g[5], lh = g[6], st = g[7], pb = g[8], org = g[9];
    uint8_t *luma = malloc((size_t)ls * (g[3] + 1));
    orc_luma_pad(luma, ls, y, y_stride, w, h);
    uint8_t *p0 = dst + org, *ph = p0 + pb, *pv = ph + pb, *pc = pv + pb;
    for (int r = 0; r < lh; r++) {
        const uint8_t *s0 = luma + (size_t)(2 * r) * ls, *s1 = s0 + ls, *s2 = s1 + ls;
        for (int x = 0; x < lw; x++) {
            p0[(size_t)r * st + x] = filt(s0[2 * x], s1[2 * x], s0[2 * x + 1], s1[2 * x + 1]);
            ph[(size_t)r * st + x] = filt(s0[2 * x + 1], s1[2 * x + 1], s0[2 * x + 2], s1[2 * x + 2]);
            pv[(size_t)r * st + x] = filt(s1[2 * x], s2[2 * x], s1[2 * x + 1], s2[2 * x + 1]);
            pc[(size_t)r * st + x] = filt(s1[2 * x + 1], s2[2 * x + 1], s1[2 * x + 2], s2[2 * x + 2]);
        }
    }
    free(luma);
    for (int k = 0; k < 4; k++) {
        uint8_t *p = dst + (size_t)k * pb + org;
        for (int r = 0; r < lh; r++) {
            memset(p + (size_t)r * st - PAD, p[(size_t)r * st], PAD);
            memset(p + (size_t)r * st + lw, p[(size_t)r * st + lw - 1], PAD);
        }
        for (int r = 1; r <= PAD; r++) {
            memcpy(p - (size_t)r * st - PAD, p - PAD, lw + 2 * PAD);
            memcpy(p + (size_t)(lh - 1 + r) * st - PAD, p + (size_t)(lh - 1) * st - PAD, lw + 2 * PAD);
        }
    }
}
