/*
 * x264vfw_cuda.h -- C ABI of the B200 (sm_100a) front end for the x264vfw encoder path.
 *
 * Plain C, opaque handles, negative-on-error ints, caller-owned buffers, no exceptions.
 * Thread model: one thread per handle at a time (same as one CODEC in the reference);
 * different handles may run concurrently.
 *
 * Every entry point names the reference interface it replaces.  Paths are relative to
 * the x264vfw tree ("reference"); "[x264]" marks upstream libx264 functions, which the
 * reference only reaches through x264_encoder_encode (codec.c:1693) -- libx264 is not
 * vendored by the reference (Makefile:21-23,109).
 *
 * There is no CPU fallback behind this ABI: without a CUDA device every call fails.
 */
#ifndef X264VFW_CUDA_H
#define X264VFW_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------
 * Colour-space ids.  Input ids = csp.h:30-44 (same values).  Encoder-side ids = the
 * public libx264 X264_CSP_* values csp.c:445-513 switches on.
 * ---------------------------------------------------------------------------------- */
#define X264VFW_CUDA_CSP_MASK   0x00ff
#define X264VFW_CUDA_CSP_NONE   0x0000
#define X264VFW_CUDA_CSP_I420   0x0001
#define X264VFW_CUDA_CSP_YV12   0x0002
#define X264VFW_CUDA_CSP_YV16   0x0003
#define X264VFW_CUDA_CSP_YV24   0x0004
#define X264VFW_CUDA_CSP_NV12   0x0005
#define X264VFW_CUDA_CSP_YUYV   0x0006
#define X264VFW_CUDA_CSP_UYVY   0x0007
#define X264VFW_CUDA_CSP_BGR    0x0008
#define X264VFW_CUDA_CSP_BGRA   0x0009
#define X264VFW_CUDA_CSP_MAX    0x000a
#define X264VFW_CUDA_CSP_VFLIP  0x1000

#define X264VFW_CUDA_OUT_I420   0x0002  /* X264_CSP_I420 */
#define X264VFW_CUDA_OUT_NV12   0x0004  /* X264_CSP_NV12 */
#define X264VFW_CUDA_OUT_I422   0x0006  /* X264_CSP_I422 */
#define X264VFW_CUDA_OUT_I444   0x000c  /* X264_CSP_I444 */
#define X264VFW_CUDA_OUT_BGR    0x000e  /* X264_CSP_BGR  */
#define X264VFW_CUDA_OUT_BGRA   0x000f  /* X264_CSP_BGRA */

/* Flags for the two conversions BASELINE.json names that csp.c does NOT register
 * (csp.c:490-504 -> convert_fail).  They are only reachable through the explicit
 * entry points below, never through x264vfw_cuda_csp_init(), which keeps the reference
 * behaviour (-1).  Definitions: DESIGN.md "Extensions". */
#define X264VFW_CUDA_EXT_NONE        0
#define X264VFW_CUDA_EXT_RGB_TO_NV12 1  /* bgr/bgra -> I420 arithmetic, U/V interleaved as
                                           [x264] x264_frame_copy_picture would do next      */
#define X264VFW_CUDA_EXT_422_TO_I444 2  /* yuyv/uyvy -> I422 samples, chroma replicated x2 horizontally */

/* Same layout as libx264's public x264_image_t (what csp.h:46 takes), so a reference
 * maintainer can pass &pic.img / &codec->conv_pic.img straight through. */
typedef struct x264vfw_cuda_image_t
{
    int      i_csp;        /* X264VFW_CUDA_CSP_* | VFLIP for sources; ignored for destinations */
    int      i_plane;
    int      i_stride[4];
    uint8_t *plane[4];
} x264vfw_cuda_image_t;

/* ---- B1: the csp function table (csp.h:46-53, csp.c:440-514, call site codec.c:1774) -- */
typedef int (*x264vfw_cuda_csp_t)( x264vfw_cuda_image_t *dst, x264vfw_cuda_image_t *src,
                                   int i_width, int i_height );
typedef struct
{
    x264vfw_cuda_csp_t convert[X264VFW_CUDA_CSP_MAX];
} x264vfw_cuda_csp_function_t;

/* Drop-in for x264vfw_csp_init (csp.c:440).  Fills the table with GPU-backed converters
 * taking HOST pointers exactly like the reference ones: each call copies the borrowed
 * source (icc->lpInput, codec.c:1767) to the device, launches one kernel, and copies the
 * planes back into dst (codec->conv_pic.img) before returning 0.  Unregistered pairs
 * return -1 like convert_fail (csp.c:93-97); device errors also return -1 and leave a
 * message in x264vfw_cuda_last_error().  Staging buffers are per calling thread. */
void x264vfw_cuda_csp_init( x264vfw_cuda_csp_function_t *pf, int i_x264_csp,
                            int i_colmatrix, int b_fullrange );

/* ---- explicit-context API (same kernels, no hidden per-thread state) ---------------- */
typedef struct x264vfw_cuda_ctx x264vfw_cuda_ctx;

/* device < 0: current CUDA device.  Returns 0 or -1. */
int  x264vfw_cuda_ctx_create( x264vfw_cuda_ctx **pctx, int device );
void x264vfw_cuda_ctx_destroy( x264vfw_cuda_ctx *ctx );
/* cudaStream_t of the context as an opaque pointer (for event timing by the caller). */
void *x264vfw_cuda_ctx_stream( x264vfw_cuda_ctx *ctx );
int  x264vfw_cuda_ctx_sync( x264vfw_cuda_ctx *ctx );

/* One frame, HOST buffers, explicit context: semantics of csp.convert[] (codec.c:1774)
 * with the table selection of csp.c:445-513 folded into the arguments. */
int x264vfw_cuda_csp_convert( x264vfw_cuda_ctx *ctx, int i_x264_csp, int i_colmatrix,
                              int b_fullrange, int i_ext,
                              x264vfw_cuda_image_t *dst, x264vfw_cuda_image_t *src,
                              int i_width, int i_height );

/* n_frames frames, DEVICE buffers, one launch, asynchronous on the context's stream.
 * Frame f reads src planes at +f*src_frame_bytes and writes dst planes at
 * +f*dst_frame_bytes.  This is the entry the roofline numbers are measured on. */
int x264vfw_cuda_csp_convert_batch( x264vfw_cuda_ctx *ctx, int i_x264_csp, int i_colmatrix,
                                    int b_fullrange, int i_ext,
                                    const x264vfw_cuda_image_t *dst_dev,
                                    const x264vfw_cuda_image_t *src_dev,
                                    int i_width, int i_height,
                                    size_t src_frame_bytes, size_t dst_frame_bytes,
                                    int n_frames );

/* Buffer geometry helpers: x264vfw_img_fill (codec.c:304-379) for sources, and the tight
 * plane layout [x264] x264_picture_alloc gives conv_pic (codec.c:1673) for destinations.
 * Return the total byte size, or -1 for an unknown csp.  ptr may be NULL (sizes only). */
int64_t x264vfw_cuda_img_fill( x264vfw_cuda_image_t *img, uint8_t *ptr, int i_csp,
                               int i_width, int i_height );
int64_t x264vfw_cuda_picture_layout( x264vfw_cuda_image_t *img, uint8_t *ptr, int i_x264_csp,
                                     int i_width, int i_height );

/* ---- B3-shaped stage entry points on DEVICE buffers ([x264] function tables) ---------
 * Geometry of the lookahead's half-resolution ("lowres") planes. */
typedef struct x264vfw_cuda_lowres_geom
{
    int mb_w, mb_h;          /* (w+15)>>4, (h+15)>>4                                     */
    int luma_w, luma_h;      /* 16*mb_w, 16*mb_h  (frame padded to mod 16)               */
    int luma_stride;         /* bytes per row of the padded luma plane, >= luma_w+1      */
    int lw, lh;              /* lowres size: luma_w/2, luma_h/2                          */
    int lstride;             /* bytes per lowres row incl. 2*32 px padding               */
    int lplane_bytes;        /* lstride * (lh + 2*32)                                    */
    int lorigin;             /* offset of pixel (0,0) inside a padded lowres plane       */
} x264vfw_cuda_lowres_geom;

void x264vfw_cuda_lowres_geometry( x264vfw_cuda_lowres_geom *g, int i_width, int i_height );

/* [x264] x264_frame_copy_picture (luma part) + x264_frame_expand_border_mod16:
 * tight w*h luma -> luma_stride x (luma_h+1) plane with the last column/row replicated
 * out to the mod-16 size and one extra column/row (the duplicate x264_frame_init_lowres
 * makes).  n_frames frames per launch. */
int x264vfw_cuda_luma_pad( x264vfw_cuda_ctx *ctx, uint8_t *dst_dev, const uint8_t *y_dev,
                           int y_stride, int i_width, int i_height,
                           size_t src_frame_bytes, size_t dst_frame_bytes, int n_frames );

/* [x264] x264_frame_init_lowres = frame_init_lowres_core + x264_frame_expand_border_lowres:
 * writes the four padded half-pel phase planes (0,H,V,C), consecutive in dst_dev
 * (4*lplane_bytes per frame).  src_dev is the TIGHT w*h luma; the mod-16 replication is
 * folded into addressing (no separate padded plane needed). */
int x264vfw_cuda_lowres_init( x264vfw_cuda_ctx *ctx, uint8_t *dst_dev, const uint8_t *y_dev,
                              int y_stride, int i_width, int i_height,
                              size_t src_frame_bytes, size_t dst_frame_bytes, int n_frames );

const char *x264vfw_cuda_last_error( void );
/* "x264vfw_cuda <version> sm_100a"; also proves the library loaded. */
const char *x264vfw_cuda_version( void );
/* Number of kernels this library has launched in the calling process (all handles). */
uint64_t x264vfw_cuda_launch_count( void );

#ifdef __cplusplus
}
#endif
#endif /* X264VFW_CUDA_H */
