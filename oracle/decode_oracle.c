/*
 * decode_oracle.c -- CPU restatement of the reference's DECODER-side colour conversion
 * (SURVEY.md 8(f) row 4).  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * The reference does not contain this arithmetic: x264vfw_decompress (codec.c:2258-2292) hands the
 * decoded AVFrame to libswscale's sws_scale(), with a context built by x264vfw_init_sws_context
 * (codec.c:2075-2152).  libswscale is a third-party dependency that is absent from /root/reference
 * and is not version-pinned by it (Makefile: -lswscale).  What is restated here is the PUBLISHED
 * behaviour of libswscale 9.1.100 (FFmpeg 8.0 line) on x86-64 for exactly the context the reference
 * builds, and it is PINNED: checked byte for byte against that library (the copy bundled in this
 * image's opencv wheel, opencv_python_headless.libs/libswscale-*.so.9.1.100) called the way the
 * reference calls it -- see tests/golden/make_decode_golden.py (fixtures) and
 * tests/test_decode_oracle.py (live comparison when the wheel is importable).
 *
 * What the reference's context amounts to (all same-size, 8-bit, yuv420p in):
 *   flags = SWS_BICUBIC | SWS_FULL_CHR_H_INP | SWS_ACCURATE_RND      (codec.c:2081-2082,2097)
 *   SWS_FULL_CHR_H_INT is OR-ed into the local variable only AFTER it was handed to the context
 *   (codec.c:2097 then :2110-2111), so the context never sees it: RGB output keeps chroma at half
 *   horizontal resolution (each chroma sample feeds two pixels), and because SWS_ACCURATE_RND
 *   disables libswscale's unscaled yuv2rgb shortcut, the GENERAL scaler runs:
 *     - luma: identity;
 *     - chroma, horizontal: identity;  chroma, vertical: 2x bicubic upscale, 4 taps, 12-bit
 *       coefficients from initFilter() [libswscale/utils.c], siting 128/128 (centre);
 *     - output rows 0..h-3: the x86 MMXEXT "accurate rounding" writers
 *       yuv2rgb32_X_ar / yuv2bgr24_X_ar / yuv2yuyv422_X_ar [libswscale/x86/swscale_template.c],
 *       16-bit pmulhw arithmetic, with the coefficient-pair packing of ff_updateMMXDitherTables
 *       [libswscale/x86/swscale.c] (f[i] + f[i+1]*65536 as ONE int: a negative f[i] borrows 1 from
 *       f[i+1]);
 *     - output rows h-2, h-1: the portable C writers (yuv2rgb_X_c_template / yuv2422_X_c_template
 *       [libswscale/output.c] over the tables of ff_yuv2rgb_c_init_tables [libswscale/yuv2rgb.c]) --
 *       libswscale switches to them for the last two lines so the SIMD code cannot overrun;
 *     - UYVY has no _ar writer: every row takes the C writer;
 *     - I420 / YV12 / NV12 targets are plane copies (planarCopyWrapper / planarToNv12Wrapper
 *       [libswscale/swscale_unscaled.c]); YV12 swaps the destination U/V pointers (codec.c:2263-2274).
 *   colour matrix: sws_getCoefficients(colorspace) (codec.c:2113-2140), range = decoder's, kept
 *   (codec.c:2091-2095), brightness 0, contrast = saturation = 1<<16 (codec.c:2141-2144).
 *
 * Not restated: the SIMD writers store whole groups of 8 pixels, so for widths that are not a
 * multiple of 8 libswscale writes up to 7 pixels past the end of each of rows 0..h-3 (into the next
 * row, which is then overwritten; the reference's output DIB has no padding).  The oracle writes
 * exactly width pixels per row; tests compare the width x height region.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

/* ---- initFilter() of libswscale/utils.c, specialised to what this context asks of it:
 *      bicubic (B=0, C=0.6), upscale (xInc <= 1<<16), no src/dst filter vectors. ------------- */
static int64_t rounded_div(int64_t a, int64_t b)
{
    return a >= 0 ? (a + (b >> 1)) / b : (a - (b >> 1)) / b;
}

/* initFilter() for the two bicubic cases this context meets: 2x upscale (xInc <= 1<<16: 1 + 4 taps before the cut) and 2:1
 * downscale (xInc > 1<<16: 1 + 4 * srcW / dstW = 9 taps before the cut, distances scaled by dstW / srcW, fone halved).
 * Fills coef[dst_n][8] (zero padded) and pos[dst_n]; returns the filter size (<= 8) or -1. */
int orc_sws_filter(int src_n, int dst_n, int filter_align, int one, int src_pos, int dst_pos, int16_t (*coef)[8], int *pos)
{
    const int x_inc = (int)((((int64_t)src_n << 16) + (dst_n >> 1)) / dst_n);
    const int down = x_inc > (1 << 16);
    int lg = 0;                                             /* av_log2(srcW / dstW), capped at 8 */
    for (int q = src_n / dst_n; q > 1; q >>= 1) lg++;
    const int64_t fone = (int64_t)1 << (54 - (lg < 8 ? lg : 8));
    int size = down ? 1 + (4 * src_n + dst_n - 1) / dst_n : 1 + 4;
    if (size > src_n - 2) size = src_n - 2;
    if (size < 1) size = 1;
    if (llabs((long long)x_inc - 0x10000) < 10 || size > 9) return -1;                     /* other initFilter branches */

    int64_t *f = calloc((size_t)dst_n * size, sizeof(*f));
    if (!f) return -1;
    const int64_t B = 0, Cc = (int64_t)(0.6 * (1 << 24));
    int64_t x_dst_in_src = (((int64_t)dst_pos * x_inc) >> 7) - (((int64_t)src_pos * 0x10000LL) >> 7);
    for (int i = 0; i < dst_n; i++) {
        int xx = (int)((x_dst_in_src - (int64_t)(size - 2) * (1LL << 16)) / (1 << 17));
        pos[i] = xx;
        for (int j = 0; j < size; j++) {
            int64_t d = llabs((int64_t)xx * (1 << 17) - x_dst_in_src) << 13;
            int64_t c;
            if (down) d = d * dst_n / src_n;
            if (d >= 1LL << 31)
                c = 0;
            else {
                int64_t dd = (d * d) >> 30, ddd = (dd * d) >> 30;
                if (d < 1LL << 30)
                    c = (12 * (1 << 24) - 9 * B - 6 * Cc) * ddd + (-18 * (1 << 24) + 12 * B + 6 * Cc) * dd +
                        (6 * (1 << 24) - 2 * B) * (1LL << 30);
                else
                    c = (-B - 6 * Cc) * ddd + (6 * B + 30 * Cc) * dd + (-12 * B - 48 * Cc) * d +
                        (8 * B + 24 * Cc) * (1LL << 30);
            }
            f[(size_t)i * size + j] = c / ((1LL << 54) / fone);
            xx++;
        }
        x_dst_in_src += 2LL * x_inc;
    }

    /* shrink: drop near-zero taps on the left (shifting), count them on the right */
    const double cut = 0.002 * (double)fone;               /* SWS_MAX_REDUCE_CUTOFF */
    int min_size = 0;
    for (int i = dst_n - 1; i >= 0; i--) {
        int64_t *fi = f + (size_t)i * size;
        int mn = size;
        int64_t acc = 0;
        for (int j = 0; j < size; j++) {
            acc += llabs(fi[0]);
            if ((double)acc > cut) break;
            if (i < dst_n - 1 && pos[i] >= pos[i + 1]) break;   /* keep positions monotonic */
            memmove(fi, fi + 1, (size - 1) * sizeof(*fi));
            fi[size - 1] = 0;
            pos[i]++;
        }
        acc = 0;
        for (int j = size - 1; j > 0; j--) {
            acc += llabs(fi[j]);
            if ((double)acc > cut) break;
            mn--;
        }
        if (mn > min_size) min_size = mn;
    }
    int out_size = (min_size + (filter_align - 1)) & ~(filter_align - 1);
    if (out_size > 8 || out_size > src_n) { free(f); return -1; }

    for (int i = 0; i < dst_n; i++) {
        int64_t t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int j = 0; j < out_size && j < size; j++) t[j] = f[(size_t)i * size + j];
        /* borders: fold taps that fall outside [0, src_n) onto the edge sample */
        if (pos[i] < 0) {
            for (int j = 1; j < out_size; j++) {
                int left = j + pos[i] > 0 ? j + pos[i] : 0;
                t[left] += t[j];
                t[j] = 0;
            }
            pos[i] = 0;
        }
        if (pos[i] + out_size > src_n) {
            int shift = pos[i] + (out_size - src_n < 0 ? out_size - src_n : 0);
            int64_t acc = 0;
            for (int j = out_size - 1; j >= 0; j--)
                if (pos[i] + j >= src_n) { acc += t[j]; t[j] = 0; }
            for (int j = out_size - 1; j >= 0; j--)
                t[j] = j < shift ? 0 : t[j - shift];
            pos[i] -= shift;
            t[src_n - 1 - pos[i]] += acc;
        }
        /* normalise to `one` with error feedback */
        int64_t sum = 0, err = 0;
        for (int j = 0; j < out_size; j++) sum += t[j];
        sum = (sum + one / 2) / one;
        if (!sum) sum = 1;
        for (int j = 0; j < 8; j++) coef[i][j] = 0;
        for (int j = 0; j < out_size; j++) {
            int64_t v = t[j] + err;
            int64_t iv = rounded_div(v, sum);
            coef[i][j] = (int16_t)iv;
            err = v - iv * sum;
        }
    }
    free(f);
    return out_size;
}

/* the 2x upscale with at most 4 taps, as the packed writers use it: coef[dst_n][4] */
int orc_sws_bicubic_filter(int src_n, int dst_n, int filter_align, int one, int src_pos, int dst_pos,
                           int16_t (*coef)[4], int *pos)
{
    if (dst_n != 2 * src_n) return -1;
    int16_t (*c8)[8] = malloc(sizeof(int16_t[8]) * dst_n);
    if (!c8) return -1;
    int n = orc_sws_filter(src_n, dst_n, filter_align, one, src_pos, dst_pos, c8, pos);
    if (n > 4) n = -1;
    if (n > 0) for (int i = 0; i < dst_n; i++) for (int j = 0; j < 4; j++) coef[i][j] = c8[i][j];
    free(c8);
    return n;
}

/* ---- colour constants: ff_yuv2rgb_c_init_tables() of libswscale/yuv2rgb.c ------------------ */
typedef struct {
    /* SIMD writers (16-bit) */
    int y_coeff, vr_coeff, ub_coeff, vg_coeff, ug_coeff, y_offset;
    /* C writers (tables, evaluated arithmetically) */
    int64_t cy, oy, crv, cbu, cgu, cgv;
    int yoffs;
} orc_yuv2rgb_t;

static int round_to_int16(int64_t f)
{
    int r = (int)((f + (1 << 15)) >> 16);
    return r < -0x7FFF ? -0x7FFF : r > 0x7FFF ? 0x7FFF : r;
}

/* sws_getCoefficients(): ff_yuv2rgb_coeffs[] rows {crv, cbu, cgu, cgv} by SWS_CS_* */
static const int k_sws_coeffs[11][4] = {
    {117489, 138438, 13975, 34925}, /* 0 (no sequence_display_extension) */
    {117489, 138438, 13975, 34925}, /* 1 ITU-R Rec. 709 */
    {104597, 132201, 25675, 53279}, /* 2 unspecified */
    {104597, 132201, 25675, 53279}, /* 3 reserved */
    {104448, 132798, 24759, 53109}, /* 4 FCC */
    {104597, 132201, 25675, 53279}, /* 5 ITU-R Rec. 624-4 System B, G (= ITU601, DEFAULT) */
    {104597, 132201, 25675, 53279}, /* 6 SMPTE 170M */
    {117579, 136230, 16907, 35559}, /* 7 SMPTE 240M */
    {0, 0, 0, 0},                   /* 8 YCgCo */
    {110013, 140363, 12277, 42626}, /* 9 Bt-2020-NCL */
    {110013, 140363, 12277, 42626}, /* 10 Bt-2020-CL */
};

/* codec.c:2113-2140: AVCOL_SPC_* of the decoder context -> SWS_CS_* */
static int sws_cs_of_avcol_spc(int spc)
{
    switch (spc) {
    case 1:  return 1;   /* AVCOL_SPC_BT709      -> SWS_CS_ITU709 */
    case 4:  return 4;   /* AVCOL_SPC_FCC        -> SWS_CS_FCC */
    case 5:  return 5;   /* AVCOL_SPC_BT470BG    -> SWS_CS_ITU601 */
    case 6:  return 6;   /* AVCOL_SPC_SMPTE170M  -> SWS_CS_SMPTE170M */
    case 7:  return 7;   /* AVCOL_SPC_SMPTE240M  -> SWS_CS_SMPTE240M */
    case 9:  case 10: return 9;   /* AVCOL_SPC_BT2020_NCL/_CL -> SWS_CS_BT2020 */
    default: return 5;   /* SWS_CS_DEFAULT */
    }
}

static void yuv2rgb_setup(orc_yuv2rgb_t *k, int avcol_spc, int fullrange)
{
    const int *inv = k_sws_coeffs[sws_cs_of_avcol_spc(avcol_spc)];
    int64_t crv = inv[0], cbu = inv[1], cgu = -inv[2], cgv = -inv[3];
    int64_t cy = 1 << 16, oy = 0;
    const int64_t contrast = 1 << 16, saturation = 1 << 16, brightness = 0;
    if (!fullrange) {
        cy = (cy * 255) / 219;
        oy = 16 << 16;
    } else {
        crv = (crv * 224) / 255; cbu = (cbu * 224) / 255;
        cgu = (cgu * 224) / 255; cgv = (cgv * 224) / 255;
    }
    cy  = (cy * contrast) >> 16;
    crv = (crv * contrast * saturation) >> 32;
    cbu = (cbu * contrast * saturation) >> 32;
    cgu = (cgu * contrast * saturation) >> 32;
    cgv = (cgv * contrast * saturation) >> 32;
    oy -= 256 * brightness;

    k->y_coeff  = round_to_int16(cy * (1 << 13));
    k->vr_coeff = round_to_int16(crv * (1 << 13));
    k->ub_coeff = round_to_int16(cbu * (1 << 13));
    k->vg_coeff = round_to_int16(cgv * (1 << 13));
    k->ug_coeff = round_to_int16(cgu * (1 << 13));
    k->y_offset = round_to_int16(oy * (1 << 3));

    /* "scale coefficients by cy" */
    int64_t d = cy > 1 ? cy : 1;
    k->crv = (crv * (1 << 16) + 0x8000) / d;
    k->cbu = (cbu * (1 << 16) + 0x8000) / d;
    k->cgu = (cgu * (1 << 16) + 0x8000) / d;
    k->cgv = (cgv * (1 << 16) + 0x8000) / d;
    k->cy = cy; k->oy = oy;
    k->yoffs = fullrange ? 384 : 326;
}

static inline int clip_u8(int64_t v) { return v < 0 ? 0 : v > 255 ? 255 : (int)v; }
static inline int wrap16(int v) { return (int16_t)v; }

/* one entry of the y_table the C writers index: value at table position yoffs + k */
static inline int c_table(const orc_yuv2rgb_t *k, int idx)
{
    return clip_u8((((int64_t)k->yoffs + idx) * k->cy - (384LL << 16) - k->oy + 0x8000) >> 16);
}

/* dst formats: the reference's own output csp codes (csp.h:30-44), as get_csp() returns them for
 * the output header (codec.c:1994-1998) */
enum { F_I420 = 1, F_YV12 = 2, F_YV16 = 3, F_YV24 = 4, F_NV12 = 5, F_YUYV = 6, F_UYVY = 7, F_BGR = 8, F_BGRA = 9, F_VFLIP = 0x1000 };

/* src_chroma: 1 = the decoder delivers yuv420p, 2 = yuv422p.  A 4:2:2 picture has a chroma line per luma line, so the
 * vertical filter degenerates to one tap of 1.0 and libswscale picks its single-line writers (yuv2packed1): the SIMD ones
 * (rows 0..h-3) then add no rounder, (sample << 7) >> 4 instead of (sum >> 16) + 4; the C ones (last two rows) take
 * (sample << 7 + 64) >> 7 = the sample.  YUY2 / UYVY / YV16 targets are pure (de)interleaves (yuv422pToYuy2Wrapper,
 * yuv422pToUyvyWrapper, planarCopyWrapper).  Checked against libswscale like the 4:2:0 case. */
int orc_decode_convert_src(int src_chroma, int out_csp, uint8_t *dst, const uint8_t *const src[3], const int src_stride[3],
                           int w, int h, int avcol_spc, int fullrange)
{
    const int fmt = out_csp & 0xff;
    const int flip = (out_csp & F_VFLIP) != 0;
    if (w <= 0 || h <= 0 || (w & 1) || (h & 1)) return -1;        /* codec.c:1950-1954 */
    if (src_chroma < 1 || src_chroma > 3) return -1;
    const int v422 = src_chroma == 2, v444 = src_chroma == 3;
    const int cw = v444 ? w : w / 2, ch = v422 || v444 ? h : h / 2;
    const int out420 = fmt == F_I420 || fmt == F_YV12 || fmt == F_NV12;
    const int out_chroma = out420 ? 1 : fmt == F_YV16 || fmt == F_YUYV || fmt == F_UYVY ? 2 : fmt == F_YV24 ? 3 : 0;
    const int planar_out = out420 || fmt == F_YV16 || fmt == F_YV24;
    if (planar_out && out_chroma != src_chroma) {
        /* planar output with another chroma resolution than the picture (4:2:0 -> YV16 / YV24, 4:2:2 -> I420 / YV12 / NV12 / YV24,
         * 4:4:4 -> I420 / YV12 / NV12 / YV16): libswscale's general scaler on the chroma planes -- horizontal bicubic (initFilter:
         * 4 taps for 2x up, 8 taps for 2:1 down; 14-bit coefficients, filterAlign 4) into 15-bit intermediates
         *   c15 = min((sum tap * sample) >> 7, 32767)                [hScale8To15],
         * then the vertical filter (4 / 8 taps, 12-bit, filterAlign 2; identity when the heights agree) and the 8-bit plane writer
         *   out = clip8((sum tap * c15 + (64 << 12)) >> 19)          [yuv2planeX_8 / yuv2plane1_8 / yuv2nv12cX, flat dither 64]
         * in every row (the SIMD plane writers are bit-exact with the C ones); luma is copied.  YV12 / YV16 / YV24: planes swapped. */
        if (flip) return -1;
        const int ocw = fmt == F_YV24 ? w : w / 2, och = out420 ? h / 2 : h;          /* output chroma plane */
        const int hs = ocw != cw, vs = och != ch;
        int16_t (*hc)[8] = malloc(sizeof(int16_t[8]) * ocw), (*vc)[8] = malloc(sizeof(int16_t[8]) * och);
        int *hp = malloc(sizeof(int) * ocw), *vp = malloc(sizeof(int) * och);
        int ok = hc && vc && hp && vp, hn = 1, vn = 1;
        if (ok && hs) ok = (hn = orc_sws_filter(cw, ocw, 4, 1 << 14, 128, 128, hc, hp)) > 0;
        if (ok && vs) ok = (vn = orc_sws_filter(ch, och, 2, 1 << 12, 128, 128, vc, vp)) > 0;
        /* below these sizes initFilter cuts its tap count to the picture: not restated */
        if (ok && ((hs && (ocw > cw ? cw < 6 : cw < 12)) || (vs && (och > ch ? ch < 5 : ch < 12)))) ok = 0;
        if (!ok) { free(hc); free(vc); free(hp); free(vp); return -1; }
        uint8_t *py = dst, *p1 = dst + (size_t)w * h, *p2 = p1 + (size_t)ocw * och;
        for (int r = 0; r < h; r++) memcpy(py + (size_t)r * w, src[0] + (ptrdiff_t)r * src_stride[0], w);
        for (int c = 0; c < 2; c++) {
            uint8_t *o; int ostep, ostride;
            if (fmt == F_NV12) { o = p1 + c; ostep = 2; ostride = w; }
            else { o = (fmt == F_I420) == (c == 0) ? p1 : p2; ostep = 1; ostride = ocw; }   /* YV12 / YV16 / YV24: V first (codec.c:2263-2274) */
            const uint8_t *sp = src[1 + c];
            const int st = src_stride[1 + c];
            for (int r = 0; r < och; r++)
                for (int x = 0; x < ocw; x++) {
                    int64_t acc = 0;
                    for (int j = 0; j < (vs ? vn : 1); j++) {
                        const uint8_t *line = sp + (ptrdiff_t)(vs ? vp[r] + j : r) * st;
                        int c15;
                        if (hs) {
                            int a = 0;
                            for (int i = 0; i < hn; i++) a += line[hp[x] + i] * hc[x][i];
                            c15 = a >> 7;
                            if (c15 > 32767) c15 = 32767;
                        } else
                            c15 = line[x] << 7;
                        acc += (int64_t)c15 * (vs ? vc[r][j] : 4096);
                    }
                    o[(size_t)r * ostride + (size_t)x * ostep] = clip_u8((acc + (64 << 12)) >> 19);
                }
        }
        free(hc); free(vc); free(hp); free(vp);
        return 0;
    }
    if (v444 && out_chroma == 2) {
        /* 4:4:4 picture -> YUY2 / UYVY: chroma down-sampled horizontally as above (8 taps), then libswscale's single-line packed
         * writers: yuv2yuyv422_1 SIMD in rows 0..h-3 (c15 >> 7), the C writer in the last two rows and in every row of UYVY
         * ((c15 + 64) >> 7); luma passes through. */
        if (flip || w < 12) return -1;                              /* below 12 chroma samples initFilter cuts its tap count to the plane: not restated */
        const int ocw = w / 2;
        int16_t (*hc)[8] = malloc(sizeof(int16_t[8]) * ocw);
        int *hp = malloc(sizeof(int) * ocw);
        int hn = hc && hp ? orc_sws_filter(cw, ocw, 4, 1 << 14, 128, 128, hc, hp) : -1;
        if (hn <= 0) { free(hc); free(hp); return -1; }
        for (int r = 0; r < h; r++) {
            const int c_writer = fmt == F_UYVY || r >= h - 2;
            uint8_t *o = dst + (size_t)r * 2 * w;
            for (int x = 0; x < ocw; x++) {
                int cv[2];
                for (int c = 0; c < 2; c++) {
                    const uint8_t *line = src[1 + c] + (ptrdiff_t)r * src_stride[1 + c];
                    int a = 0;
                    for (int i = 0; i < hn; i++) a += line[hp[x] + i] * hc[x][i];
                    int c15 = a >> 7;
                    if (c15 > 32767) c15 = 32767;
                    cv[c] = clip_u8(c_writer ? (c15 + 64) >> 7 : c15 >> 7);
                }
                const int y0 = src[0][(ptrdiff_t)r * src_stride[0] + 2 * x], y1 = src[0][(ptrdiff_t)r * src_stride[0] + 2 * x + 1];
                if (fmt == F_YUYV) { o[4 * x] = y0; o[4 * x + 1] = cv[0]; o[4 * x + 2] = y1; o[4 * x + 3] = cv[1]; }
                else               { o[4 * x] = cv[0]; o[4 * x + 1] = y0; o[4 * x + 2] = cv[1]; o[4 * x + 3] = y1; }
            }
        }
        free(hc); free(hp);
        return 0;
    }

    /* x264vfw_picture_fill (codec.c:419-503) geometry of the output DIB */
    if (fmt == F_I420 || fmt == F_YV12 || fmt == F_NV12 || fmt == F_YV16 || fmt == F_YV24) {
        if (flip) return -1;                                      /* x264vfw_picture_vflip: RGB only (codec.c:510-527) */
        uint8_t *py = dst, *p1 = dst + (size_t)w * h, *p2 = p1 + (size_t)cw * ch;
        for (int r = 0; r < h; r++) memcpy(py + (size_t)r * w, src[0] + (ptrdiff_t)r * src_stride[0], w);
        if (fmt == F_NV12) {
            for (int r = 0; r < ch; r++)
                for (int x = 0; x < cw; x++) {
                    p1[(size_t)r * w + 2 * x]     = src[1][(ptrdiff_t)r * src_stride[1] + x];
                    p1[(size_t)r * w + 2 * x + 1] = src[2][(ptrdiff_t)r * src_stride[2] + x];
                }
        } else {
            uint8_t *pu = fmt != F_I420 ? p2 : p1, *pv = fmt != F_I420 ? p1 : p2;   /* YV12 / YV16: codec.c:2263-2274 */
            for (int r = 0; r < ch; r++) {
                memcpy(pu + (size_t)r * cw, src[1] + (ptrdiff_t)r * src_stride[1], cw);
                memcpy(pv + (size_t)r * cw, src[2] + (ptrdiff_t)r * src_stride[2], cw);
            }
        }
        return 0;
    }
    if (fmt != F_BGR && fmt != F_BGRA && fmt != F_YUYV && fmt != F_UYVY) return -1;
    if (flip && fmt != F_BGR && fmt != F_BGRA) return -1;
    if (v444) {
        /* 4:4:4 pictures: libswscale switches full chroma interpolation on by itself and EVERY row goes through the portable
         * writer yuv2rgb_full_1_c -> yuv2rgb_write_full [libswscale/output.c]: per pixel, on the 15-bit intermediates * 4,
         *   Y = ((y << 9) - y_offset) * y_coeff + (1 << 21),  R = Y + V * v2r,  G = Y + V * v2g + U * u2g,  B = Y + U * u2b
         * in 32-bit arithmetic that wraps (the products are unsigned in the source), all three clipped to 30 bits when any of
         * them has one of the top two bits set (== clipping each on its own), >> 22.  Coefficients: the 13-bit ones of the SIMD
         * writers; y_offset = round16(oy << 9). */
        orc_yuv2rgb_t k;
        yuv2rgb_setup(&k, avcol_spc, fullrange);
        const int y_off9 = round_to_int16(k.oy * (1 << 9));
        const int bpp = fmt == F_BGRA ? 4 : 3;
        ptrdiff_t stride = fmt == F_BGR ? ((w * 3 + 3) & ~3) : w * 4;
        if (flip) { dst += stride * (h - 1); stride = -stride; }
        for (int r = 0; r < h; r++)
            for (int x = 0; x < w; x++) {
                const int32_t Y = (int32_t)((uint32_t)(((src[0][(ptrdiff_t)r * src_stride[0] + x] << 9) - y_off9) * k.y_coeff) + (1u << 21));
                const int32_t U = (src[1][(ptrdiff_t)r * src_stride[1] + x] - 128) << 9, V = (src[2][(ptrdiff_t)r * src_stride[2] + x] - 128) << 9;
                int32_t R = (int32_t)((uint32_t)Y + (uint32_t)V * (uint32_t)k.vr_coeff);
                int32_t G = (int32_t)((uint32_t)Y + (uint32_t)V * (uint32_t)k.vg_coeff + (uint32_t)U * (uint32_t)k.ug_coeff);
                int32_t B = (int32_t)((uint32_t)Y + (uint32_t)U * (uint32_t)k.ub_coeff);
                if ((R | G | B) & 0xC0000000) {
                    R = R < 0 ? 0 : R > 0x3FFFFFFF ? 0x3FFFFFFF : R;
                    G = G < 0 ? 0 : G > 0x3FFFFFFF ? 0x3FFFFFFF : G;
                    B = B < 0 ? 0 : B > 0x3FFFFFFF ? 0x3FFFFFFF : B;
                }
                uint8_t *q = dst + r * stride + (size_t)x * bpp;
                q[0] = B >> 22; q[1] = G >> 22; q[2] = R >> 22;
                if (bpp == 4) q[3] = 255;
            }
        return 0;
    }
    if (h < 10) return -1;                       /* below this initFilter degenerates (fewer taps); not restated */

    ptrdiff_t stride = fmt == F_BGR ? ((w * 3 + 3) & ~3) : fmt == F_BGRA ? w * 4 : w * 2;
    if (flip) { dst += stride * (h - 1); stride = -stride; }      /* codec.c:515-518 */

    int16_t (*coef)[4] = malloc(sizeof(int16_t[4]) * h);
    int *pos = malloc(sizeof(int) * h);
    if (!coef || !pos) { free(coef); free(pos); return -1; }
    /* vertical chroma filter: filterAlign 2 (x86), one = 1<<12, both sitings 128 */
    if (v422) {
        for (int r = 0; r < h; r++) { pos[r] = r < h - 3 ? r : h - 4; for (int j = 0; j < 4; j++) coef[r][j] = j == r - pos[r] ? 4096 : 0; }
    } else {
        int n = orc_sws_bicubic_filter(ch, h, 2, 1 << 12, 128, 128, coef, pos);
        if (n != 4) { free(coef); free(pos); return -1; }
    }
    const int rnd = v422 ? 0 : 4;                                 /* the rounder of the vertical-filter SIMD writers */

    orc_yuv2rgb_t k;
    yuv2rgb_setup(&k, avcol_spc, fullrange);

    for (int r = 0; r < h; r++) {
        const int c_writer = r >= h - 2 || fmt == F_UYVY;
        int f[4] = {coef[r][0], coef[r][1], coef[r][2], coef[r][3]};
        if (!c_writer)                                            /* the packed-pair borrow */
            for (int j = 0; j < 4; j += 2)
                if (f[j] < 0) f[j + 1] = wrap16(f[j + 1] - 1);
        const uint8_t *yrow = src[0] + (ptrdiff_t)r * src_stride[0];
        const uint8_t *ul[4], *vl[4];
        for (int j = 0; j < 4; j++) {
            ul[j] = src[1] + (ptrdiff_t)(pos[r] + j) * src_stride[1];
            vl[j] = src[2] + (ptrdiff_t)(pos[r] + j) * src_stride[2];
        }
        uint8_t *o = dst + r * stride;
        for (int x = 0; x < cw; x++) {
            /* vertical filter on the 15-bit intermediates (sample << 7) */
            int64_t au = 0, av = 0;
            for (int j = 0; j < 4; j++) {
                au += (int64_t)(ul[j][x] << 7) * f[j];
                av += (int64_t)(vl[j][x] << 7) * f[j];
            }
            const int y0 = yrow[2 * x], y1 = yrow[2 * x + 1];
            if (fmt == F_YUYV || fmt == F_UYVY) {
                int U, V;
                if (c_writer) {
                    U = clip_u8((au + (1 << 18)) >> 19);
                    V = clip_u8((av + (1 << 18)) >> 19);
                } else {                                          /* psrad 16, packssdw, paddw rounder, psraw 3, packuswb */
                    int u16 = wrap16((int)(au >> 16 < -32768 ? -32768 : au >> 16 > 32767 ? 32767 : au >> 16) + rnd);
                    int v16 = wrap16((int)(av >> 16 < -32768 ? -32768 : av >> 16 > 32767 ? 32767 : av >> 16) + rnd);
                    U = clip_u8(u16 >> 3);
                    V = clip_u8(v16 >> 3);
                }
                if (fmt == F_YUYV) { o[4 * x] = y0; o[4 * x + 1] = U; o[4 * x + 2] = y1; o[4 * x + 3] = V; }
                else               { o[4 * x] = U; o[4 * x + 1] = y0; o[4 * x + 2] = V; o[4 * x + 3] = y1; }
                continue;
            }
            int Bv[2], Gv[2], Rv[2];
            if (c_writer) {
                const int U = clip_u8((au + (1 << 18)) >> 19), V = clip_u8((av + (1 << 18)) >> 19);
                const int dr = (int)(((int64_t)V * k.crv) >> 16) - (int)(k.crv >> 9);
                const int db = (int)(((int64_t)U * k.cbu) >> 16) - (int)(k.cbu >> 9);
                const int dg = (int)(((int64_t)U * k.cgu) >> 16) - (int)(k.cgu >> 9) +
                               (int)(((int64_t)V * k.cgv) >> 16) - (int)(k.cgv >> 9);
                for (int p = 0; p < 2; p++) {
                    const int Y = p ? y1 : y0;
                    Bv[p] = c_table(&k, Y + db); Gv[p] = c_table(&k, Y + dg); Rv[p] = c_table(&k, Y + dr);
                }
            } else {
                int64_t su = au >> 16, sv = av >> 16;
                su = su < -32768 ? -32768 : su > 32767 ? 32767 : su;
                sv = sv < -32768 ? -32768 : sv > 32767 ? 32767 : sv;
                const int ud = wrap16(wrap16((int)su + rnd) - 0x400), vd = wrap16(wrap16((int)sv + rnd) - 0x400);
                const int ug = (ud * k.ug_coeff) >> 16, vg = (vd * k.vg_coeff) >> 16;
                const int ub = (ud * k.ub_coeff) >> 16, vr = (vd * k.vr_coeff) >> 16;
                const int g = wrap16(ug + vg);
                for (int p = 0; p < 2; p++) {
                    const int y16 = wrap16(((p ? y1 : y0) << 3) + rnd);
                    const int yv = (wrap16(y16 - k.y_offset) * k.y_coeff) >> 16;
                    Bv[p] = clip_u8(wrap16(yv + ub)); Gv[p] = clip_u8(wrap16(yv + g)); Rv[p] = clip_u8(wrap16(yv + vr));
                }
            }
            const int bpp = fmt == F_BGRA ? 4 : 3;
            for (int p = 0; p < 2; p++) {
                uint8_t *q = o + (size_t)(2 * x + p) * bpp;
                q[0] = Bv[p]; q[1] = Gv[p]; q[2] = Rv[p];
                if (bpp == 4) q[3] = 255;
            }
        }
    }
    free(coef); free(pos);
    return 0;
}

int orc_decode_convert(int out_csp, uint8_t *dst, const uint8_t *const src[3], const int src_stride[3],
                       int w, int h, int avcol_spc, int fullrange)
{
    return orc_decode_convert_src(1, out_csp, dst, src, src_stride, w, h, avcol_spc, fullrange);
}

/* size of the output DIB: x264vfw_picture_get_size (codec.c:505-508) */
int64_t orc_decode_picture_size(int out_csp, int w, int h)
{
    switch (out_csp & 0xff) {
    case F_I420: case F_YV12: case F_NV12: return (int64_t)w * h + 2 * (int64_t)(w / 2) * (h / 2);
    case F_YV16: case F_YUYV: case F_UYVY: return (int64_t)w * 2 * h;
    case F_YV24: return (int64_t)w * 3 * h;
    case F_BGR:  return (int64_t)((w * 3 + 3) & ~3) * h;
    case F_BGRA: return (int64_t)w * 4 * h;
    default: return -1;
    }
}
