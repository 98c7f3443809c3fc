/*
 * x264vfw_harness.c -- plain-C host side above the C ABI, mirroring the call order of the
 * reference's codec session (codec.c:1381-1876) for the hot path only:
 *
 *   harness_compress_begin  ~ x264vfw_compress_begin (codec.c:1381): csp from the BITMAPINFOHEADER
 *                             (get_csp :187-231, choose_output_csp :269-302), x264vfw_csp_init (:1672),
 *                             conv_pic allocation (:1673), encoder/lookahead open (:1623)
 *   harness_compress        ~ x264vfw_compress (codec.c:1728): img_fill (:1767), convert (:1774),
 *                             x264_encoder_encode (:1693) -> here: lookahead put + decisions
 *   harness_compress_end    ~ x264vfw_compress_end (codec.c:1838): flush (:1848-1854), close (:1857)
 *
 *   harness_decompress_query / _begin / harness_decompress / _end
 *                           ~ x264vfw_decompress_query (codec.c:1930), _begin (:1982), x264vfw_decompress (:2154,
 *                             the part after avcodec_receive_frame: picture_fill, U/V swap, vflip, lazy context,
 *                             sws_scale, :2243-2295), _end (:2298); SURVEY 8(f) row 4
 *
 * It is a Linux stand-in for the Win32 caller, not a re-implementation of the wrapper: no ICM
 * messages, no registry, no muxers.  Build: make -C host.  Usage: see main() below.
 */
#include "../include/x264vfw_cuda.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int biWidth, biHeight, biBitCount;
    unsigned biCompression;              /* 0 = BI_RGB, else FOURCC */
} harness_bih;

typedef struct {
    x264vfw_cuda_la *la;
    x264vfw_cuda_csp_function_t csp;     /* codec->csp */
    x264vfw_cuda_image_t conv_pic;       /* codec->conv_pic.img */
    uint8_t *conv_buf;
    int i_csp, i_x264_csp, width, height, mb_count;
    int b_encoder_error;                 /* x264vfw.h:193 */
    float *qp, *qp_aq;
} harness_codec;

#define FOURCC(a, b, c, d) ((unsigned)(a) | ((unsigned)(b) << 8) | ((unsigned)(c) << 16) | ((unsigned)(d) << 24))

static int get_csp(const harness_bih *h)            /* codec.c:187-231 */
{
    switch (h->biCompression) {
    case FOURCC('I', '4', '2', '0'): case FOURCC('I', 'Y', 'U', 'V'): return X264VFW_CUDA_CSP_I420;
    case FOURCC('Y', 'V', '1', '2'): return X264VFW_CUDA_CSP_YV12;
    case FOURCC('Y', 'V', '1', '6'): return X264VFW_CUDA_CSP_YV16;
    case FOURCC('Y', 'V', '2', '4'): return X264VFW_CUDA_CSP_YV24;
    case FOURCC('N', 'V', '1', '2'): return X264VFW_CUDA_CSP_NV12;
    case FOURCC('Y', 'U', 'Y', 'V'): case FOURCC('Y', 'U', 'Y', '2'): return X264VFW_CUDA_CSP_YUYV;
    case FOURCC('U', 'Y', 'V', 'Y'): case FOURCC('H', 'D', 'Y', 'C'): return X264VFW_CUDA_CSP_UYVY;
    case 0: {
        int flip = h->biHeight < 0 ? 0 : X264VFW_CUDA_CSP_VFLIP;
        if (h->biBitCount == 24) return X264VFW_CUDA_CSP_BGR | flip;
        if (h->biBitCount == 32) return X264VFW_CUDA_CSP_BGRA | flip;
    }
    }
    return X264VFW_CUDA_CSP_NONE;
}

static int choose_output_csp(int i_csp, int keep)    /* codec.c:269-302 */
{
    switch (i_csp & X264VFW_CUDA_CSP_MASK) {
    case X264VFW_CUDA_CSP_YV16: return keep ? X264VFW_CUDA_OUT_I422 : X264VFW_CUDA_OUT_I420;
    case X264VFW_CUDA_CSP_YV24: return keep ? X264VFW_CUDA_OUT_I444 : X264VFW_CUDA_OUT_I420;
    case X264VFW_CUDA_CSP_NV12: return X264VFW_CUDA_OUT_NV12;
    case X264VFW_CUDA_CSP_YUYV: case X264VFW_CUDA_CSP_UYVY: return keep ? X264VFW_CUDA_OUT_I422 : X264VFW_CUDA_OUT_I420;
    case X264VFW_CUDA_CSP_BGR: return keep ? X264VFW_CUDA_OUT_BGR : X264VFW_CUDA_OUT_I420;
    case X264VFW_CUDA_CSP_BGRA: return keep ? X264VFW_CUDA_OUT_BGRA : X264VFW_CUDA_OUT_I420;
    default: return X264VFW_CUDA_OUT_I420;
    }
}

int harness_compress_begin(harness_codec *c, const harness_bih *in, const char *preset, int keep_input_csp)
{
    memset(c, 0, sizeof(*c));
    c->width = in->biWidth; c->height = abs(in->biHeight);
    if (c->width <= 0 || c->height <= 0 || (c->width & 1) || (c->height & 1)) return -1;       /* codec.c:639 */
    c->i_csp = get_csp(in);
    if (c->i_csp == X264VFW_CUDA_CSP_NONE) return -1;
    c->i_x264_csp = choose_output_csp(c->i_csp, keep_input_csp);                                /* codec.c:1472 */
    /* colour matrix / range defaults for YUV targets: undef -> BT.601, TV range (codec.c:1571-1577) */
    const int colmatrix = 2, fullrange = 0;
    x264vfw_cuda_la_params p;
    if (x264vfw_cuda_la_params_preset(&p, preset, c->width, c->height) < 0) return -1;          /* codec.c:1463 */
    p.chroma_format = c->i_x264_csp == X264VFW_CUDA_OUT_I444 ? 3 : c->i_x264_csp == X264VFW_CUDA_OUT_I422 ? 2 : 1;
    /* the session converts on the device and mirrors conv_pic back for the CPU encoder */
    if (x264vfw_cuda_la_open(&c->la, &p, -1, c->i_csp, c->i_x264_csp, colmatrix, fullrange, 0) < 0) return -1;   /* codec.c:1623 */
    x264vfw_cuda_csp_init(&c->csp, c->i_x264_csp, colmatrix, fullrange);                        /* codec.c:1672 */
    int64_t n = x264vfw_cuda_picture_layout(&c->conv_pic, NULL, c->i_x264_csp, c->width, c->height);   /* codec.c:1673 */
    if (n < 0) return -1;
    c->conv_buf = malloc((size_t)n);
    x264vfw_cuda_picture_layout(&c->conv_pic, c->conv_buf, c->i_x264_csp, c->width, c->height);
    c->mb_count = ((c->width + 15) >> 4) * ((c->height + 15) >> 4);
    c->qp = malloc(sizeof(float) * c->mb_count);
    c->qp_aq = malloc(sizeof(float) * c->mb_count);
    return 0;
}

static void drain(harness_codec *c, FILE *out)
{
    x264vfw_cuda_la_decision d;
    while (x264vfw_cuda_la_get_decision(c->la, &d, c->qp, c->qp_aq) == 1) {
        /* here the reference would call x264_encoder_encode with pic_in.i_type = d.i_type and
         * pic_in.prop.quant_offsets = qp - qp_aq */
        double s = 0;
        for (int i = 0; i < d.mb_count; i++) s += c->qp[i];
        if (out) fprintf(out, "frame %d type %d key %d cost %d mean_qp_offset %.4f\n", d.i_frame, d.i_type, d.b_keyframe, d.i_cost_est, s / d.mb_count);
    }
}

int harness_compress(harness_codec *c, uint8_t *lpInput, FILE *out)
{
    if (c->b_encoder_error) return -1;
    x264vfw_cuda_image_t pic;
    if (x264vfw_cuda_img_fill(&pic, lpInput, c->i_csp, c->width, c->height) < 0) { c->b_encoder_error = 1; return -1; }   /* codec.c:1767 */
    /* codec.c:1774 + :1693 in one call: convert on the device, keep the planes there for the
     * lookahead, and return them in conv_pic for the CPU encoder */
    if (x264vfw_cuda_la_put_frame(c->la, &pic, 0, &c->conv_pic) < 0) { c->b_encoder_error = 1; return -1; }              /* codec.c:1776-1778 */
    drain(c, out);
    return 0;
}

int harness_compress_end(harness_codec *c, FILE *out)
{
    if (c->la && !c->b_encoder_error) { x264vfw_cuda_la_flush(c->la); drain(c, out); }          /* codec.c:1842-1856 */
    x264vfw_cuda_la_close(c->la);                                                               /* codec.c:1857 */
    free(c->conv_buf); free(c->qp); free(c->qp_aq);                                             /* codec.c:1872 */
    memset(c, 0, sizeof(*c));
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Decompress side.  The H.264 decode (libavcodec, codec.c:2160-2238) is not part
 * of the path: the harness is handed the decoded picture (AVFrame data[] / linesize[]).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    x264vfw_cuda_ctx *ctx;
    x264vfw_cuda_dec *dec;               /* codec->sws */
    int out_csp, width, height;          /* get_csp() of the output header, VFLIP bit included */
    int src_chroma;                      /* 1: the decoder delivers yuv420p, 2: yuv422p (decoder_context->pix_fmt, codec.c:2092) */
    int colorspace, fullrange;           /* decoder_context->colorspace / color_range (codec.c:2091, :2114) */
} harness_decoder;

#define HARNESS_ICERR_OK         0
#define HARNESS_ICERR_BADFORMAT (-2)
#define HARNESS_ICERR_ERROR     (-100)

int harness_decompress_query(const harness_bih *in, const harness_bih *out, unsigned out_size_image)
{
    /* codec.c:1948-1977 (the fourcc check of the INPUT header, :1945, is the caller's: the harness has no bitstream) */
    if (in->biWidth <= 0 || in->biHeight <= 0) return HARNESS_ICERR_BADFORMAT;
    if (in->biWidth % 2 || in->biHeight % 2) return HARNESS_ICERR_BADFORMAT;
    if (!out) return HARNESS_ICERR_OK;
    if (in->biWidth != out->biWidth || in->biHeight != abs(out->biHeight)) return HARNESS_ICERR_BADFORMAT;
    int i_csp = get_csp(out);
    if (i_csp == X264VFW_CUDA_CSP_NONE) return HARNESS_ICERR_BADFORMAT;
    int64_t size = x264vfw_cuda_dec_picture_size(i_csp, in->biWidth, in->biHeight);
    if (size < 0) return HARNESS_ICERR_BADFORMAT;      /* csp_to_pix_fmt == NONE there; here also the YV16 / YV24 outputs this library leaves out */
    if (out_size_image != 0 && out_size_image < (uint64_t)size) return HARNESS_ICERR_BADFORMAT;
    return HARNESS_ICERR_OK;
}

int harness_decompress_begin(harness_decoder *d, const harness_bih *in, const harness_bih *out, int src_chroma, int colorspace, int fullrange)
{
    memset(d, 0, sizeof(*d));
    if (harness_decompress_query(in, out, 0) != HARNESS_ICERR_OK) return HARNESS_ICERR_BADFORMAT;    /* codec.c:1988 */
    d->out_csp = get_csp(out);                                                                      /* codec.c:1994-1998 */
    d->width = in->biWidth; d->height = in->biHeight;
    d->src_chroma = src_chroma; d->colorspace = colorspace; d->fullrange = fullrange;
    if (x264vfw_cuda_ctx_create(&d->ctx, -1) < 0) return HARNESS_ICERR_ERROR;
    return HARNESS_ICERR_OK;
}

/* lpOutput: the application's DIB (icd->lpOutput), x264vfw_cuda_dec_picture_size() bytes */
int harness_decompress(harness_decoder *d, const uint8_t *const data[3], const int linesize[3], uint8_t *lpOutput)
{
    if (!d->dec &&                                                                                  /* codec.c:2282-2290 */
        x264vfw_cuda_dec_open(&d->dec, d->ctx, d->out_csp, d->width, d->height, d->src_chroma, d->colorspace, d->fullrange) < 0)
        return HARNESS_ICERR_ERROR;
    /* picture_fill, the YV12 pointer swap and the bottom-up flip (codec.c:2258-2280) follow from out_csp inside */
    if (x264vfw_cuda_dec_convert(d->dec, lpOutput, data, linesize) < 0) return HARNESS_ICERR_ERROR; /* codec.c:2292 */
    return HARNESS_ICERR_OK;
}

int harness_decompress_end(harness_decoder *d)
{
    x264vfw_cuda_dec_close(d->dec);                                                                 /* codec.c:2306 */
    x264vfw_cuda_ctx_destroy(d->ctx);
    memset(d, 0, sizeof(*d));
    return HARNESS_ICERR_OK;
}

/* ------------------------------------------------------------------------------------------
 * Several sessions at once, ONE HOST THREAD PER STREAM -- the reference runs the path on
 * whichever application thread sends ICM_COMPRESS (codec.c:1728), one call in flight per
 * CODEC, different CODECs concurrently.  bench.py and the tests drive this through ctypes so
 * that the per-frame loop is native code (Python threads would serialise on the interpreter
 * lock and measure that instead).
 * ---------------------------------------------------------------------------------------- */
#include <pthread.h>

typedef struct harness_stream {
    x264vfw_cuda_la *la;                 /* open session (x264vfw_cuda_la_open)                    */
    const uint8_t *const *frames;        /* clip: n_frames packed frames, host (pinned) or device   */
    int n_frames, on_device;
    int in_csp, out_csp, width, height;
    uint8_t *const *conv;                /* n_conv host buffers receiving conv_pic, or NULL         */
    int n_conv;
    long pos;                            /* frames fed so far (ping-pong playback position)         */
    long decided;                        /* decisions drained so far                                */
    double checksum;                     /* over the drained qp offsets (keeps the copies honest)   */
    int mb_count, error;
    float *qp, *qp_aq;                   /* scratch, allocated on first use                         */
    int count;                           /* frames to feed in this call (set by the runner)         */
    /* optional decision log (parity tests): log_cap records; qp arrays are kept as FNV-1a-64 hashes */
    struct harness_logrec *log; int log_cap, log_n;
} harness_stream;

typedef struct harness_logrec {
    x264vfw_cuda_la_decision d;
    uint64_t qp_fnv, qp_aq_fnv;
} harness_logrec;

static uint64_t fnv1a64(const void *p, size_t n)
{
    const uint8_t *b = (const uint8_t *)p;
    uint64_t h = 0xcbf29ce484222325ULL;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 0x100000001b3ULL; }
    return h;
}

static void stream_drain(harness_stream *s)
{
    x264vfw_cuda_la_decision d;
    while (x264vfw_cuda_la_get_decision(s->la, &d, s->qp, s->qp_aq) == 1) {
        s->checksum += s->qp[d.i_frame % s->mb_count] + d.i_type;
        if (s->log && s->log_n < s->log_cap) {
            harness_logrec *r = &s->log[s->log_n++];
            r->d = d;
            r->qp_fnv = fnv1a64(s->qp, sizeof(float) * s->mb_count);
            r->qp_aq_fnv = fnv1a64(s->qp_aq, sizeof(float) * s->mb_count);
        }
        s->decided++;
    }
}

/* the library's error string is per thread: keep the first failure of a stream thread where the caller can read it */
static char g_harness_error[512];
static void keep_error(void)
{
    if (!g_harness_error[0]) { strncpy(g_harness_error, x264vfw_cuda_last_error(), sizeof(g_harness_error) - 1); }
}
const char *harness_last_error(void) { return g_harness_error; }

static void *stream_thread(void *arg)
{
    harness_stream *s = (harness_stream *)arg;
    if (!s->qp) {
        s->mb_count = ((s->width + 15) >> 4) * ((s->height + 15) >> 4);
        s->qp = malloc(sizeof(float) * s->mb_count);
        s->qp_aq = malloc(sizeof(float) * s->mb_count);
    }
    for (int i = 0; i < s->count && !s->error; i++) {
        /* ping-pong playback: the clip loops without a hard cut at the wrap-around */
        const int n = s->n_frames;
        int k = n > 1 ? (int)(s->pos % (2 * n - 2)) : 0;
        if (k >= n) k = 2 * n - 2 - k;
        x264vfw_cuda_image_t pic, conv_pic, *cp = NULL;
        if (x264vfw_cuda_img_fill(&pic, (uint8_t *)s->frames[k], s->in_csp, s->width, s->height) < 0) { keep_error(); s->error = 1; break; }
        if (s->conv) {
            x264vfw_cuda_picture_layout(&conv_pic, s->conv[s->pos % s->n_conv], s->out_csp, s->width, s->height);
            cp = &conv_pic;
        }
        if (x264vfw_cuda_la_put_frame(s->la, &pic, s->on_device, cp) < 0) { keep_error(); s->error = 1; break; }
        stream_drain(s);
        s->pos++;
    }
    return NULL;
}

/* Feed `count` frames to each of the n streams concurrently; returns 0, or -1 if any failed. */
int harness_run_streams(harness_stream *streams, int n, int count)
{
    pthread_t th[256];
    if (n > 256) return -1;
    for (int i = 0; i < n; i++) { streams[i].count = count; if (pthread_create(&th[i], NULL, stream_thread, &streams[i])) return -1; }
    int rc = 0;
    for (int i = 0; i < n; i++) { pthread_join(th[i], NULL); if (streams[i].error) rc = -1; }
    return rc;
}

/* End of stream for every session (x264vfw_cuda_la_flush), decisions drained like in harness_run_streams. */
int harness_flush_streams(harness_stream *streams, int n)
{
    int rc = 0;
    for (int i = 0; i < n; i++) {
        harness_stream *s = &streams[i];
        if (!s->qp) continue;
        if (x264vfw_cuda_la_flush(s->la) < 0) { keep_error(); s->error = 1; rc = -1; continue; }
        stream_drain(s);
    }
    return rc;
}

void harness_stream_free(harness_stream *s) { free(s->qp); free(s->qp_aq); s->qp = s->qp_aq = NULL; }

#ifndef HARNESS_NO_MAIN
/* usage: x264vfw_harness <w> <h> <frames> [preset]   -- encodes a synthetic bottom-up RGB32 clip */
int main(int argc, char **argv)
{
    const int w = argc > 1 ? atoi(argv[1]) : 320, h = argc > 2 ? atoi(argv[2]) : 192, n = argc > 3 ? atoi(argv[3]) : 30;
    const char *preset = argc > 4 ? argv[4] : "medium";
    harness_bih bih = {w, h, 32, 0};
    harness_codec codec;
    if (harness_compress_begin(&codec, &bih, preset, 0) < 0) { fprintf(stderr, "begin failed: %s\n", x264vfw_cuda_last_error()); return 1; }
    uint8_t *buf = malloc((size_t)w * h * 4);
    unsigned s = 0x264;
    for (int f = 0; f < n; f++) {
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                s = s * 1664525u + 1013904223u;
                uint8_t *px = buf + ((size_t)y * w + x) * 4;
                int v = ((x + 2 * f) >> 3 ^ (y + f) >> 3) & 1 ? 180 : 60;
                px[0] = (uint8_t)(v + (s >> 30)); px[1] = (uint8_t)(v / 2 + (x & 63)); px[2] = (uint8_t)(255 - v); px[3] = 0;
            }
        if (harness_compress(&codec, buf, stdout) < 0) { fprintf(stderr, "compress failed: %s\n", x264vfw_cuda_last_error()); return 1; }
    }
    harness_compress_end(&codec, stdout);
    free(buf);
    return 0;
}
#endif
