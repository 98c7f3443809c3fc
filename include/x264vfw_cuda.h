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

/* [x264] x264_frame_copy_picture (chroma part of a 4:2:0 frame: planar U, V interleaved into the
 * encoder's internal NV12 plane) + x264_frame_expand_border_mod16 for that plane: (w/2)x(h/2) U
 * and V -> dst_stride x (luma_h/2) bytes, the last U/V pair and the last row replicated out to
 * the mod-16 size.  Together with _luma_pad this is the internal frame libx264 builds from
 * conv_pic (codec.c:1673,1774 produce it; entered via codec.c:1693).  n_frames frames per launch. */
int x264vfw_cuda_chroma_nv12_pad( x264vfw_cuda_ctx *ctx, uint8_t *dst_dev, int dst_stride,
                                  const uint8_t *u_dev, const uint8_t *v_dev, int c_stride,
                                  int i_width, int i_height,
                                  size_t src_frame_bytes, size_t dst_frame_bytes, int n_frames );

/* [x264] x264_frame_init_lowres = frame_init_lowres_core + x264_frame_expand_border_lowres:
 * writes the four padded half-pel phase planes (0,H,V,C), consecutive in dst_dev
 * (4*lplane_bytes per frame).  src_dev is the TIGHT w*h luma; the mod-16 replication is
 * folded into addressing (no separate padded plane needed). */
int x264vfw_cuda_lowres_init( x264vfw_cuda_ctx *ctx, uint8_t *dst_dev, const uint8_t *y_dev,
                              int y_stride, int i_width, int i_height,
                              size_t src_frame_bytes, size_t dst_frame_bytes, int n_frames );

/* Fused front end of a frame: ONE kernel (TMA-staged tiles of the packed rows) that replaces, for packed RGB32
 * sources and a 4:2:0 target, the sequence csp.convert[] (codec.c:1774, RGB_TO_I420 csp.c:299-388) ->
 * [x264] x264_adaptive_quant_frame (aq-mode 1: f_qp_offset_aq, i_inv_qscale_factor, i_pixel_sum / i_pixel_ssd before
 * mean removal) -> x264_frame_init_lowres incl. border, which x264_encoder_encode runs on every input picture
 * (codec.c:1693).  The lookahead session uses it for every eligible frame; this entry runs it on n_frames
 * device-resident frames in one launch (roofline probe, parity tests).  src_dev: X264VFW_CUDA_CSP_BGRA (| VFLIP)
 * image, rows 16-byte aligned; dst_dev: tight I420 planes; lowres_dev: 4 * lplane_bytes per frame;
 * qp_offset_aq_dev / inv_qscale_dev: mb_w * mb_h per frame; stats_dev: 6 uint64 per frame (sum[3], ssd[3]).
 * Needs i_width % 16 == 0; returns -1 (with a message) when the frame is not eligible. */
int x264vfw_cuda_frontend_batch( x264vfw_cuda_ctx *ctx, int i_colmatrix, int b_fullrange,
                                 const x264vfw_cuda_image_t *dst_dev, const x264vfw_cuda_image_t *src_dev,
                                 uint8_t *lowres_dev, float *qp_offset_aq_dev, uint16_t *inv_qscale_dev,
                                 unsigned long long *stats_dev, float aq_strength,
                                 int i_width, int i_height, size_t src_frame_bytes, size_t dst_frame_bytes, int n_frames );

/* ---- half-pel reference planes of a reconstructed frame ------------------
 * [x264] x264_frame_expand_border + x264_frame_filter (x264_mc_functions_t.hpel_filter) +
 * x264_frame_expand_border_filtered for a progressive 8-bit luma plane: what libx264 runs on every
 * reconstructed reference frame inside x264_encoder_encode (codec.c:1693) before the next frame's
 * motion search can read frame->filtered[0][0..3].  Geometry of one padded plane: */
typedef struct x264vfw_cuda_hpel_geom
{
    int stride;              /* bytes per row: (w + 2*32) rounded up to 64                */
    int plane_bytes;         /* stride * (h + 2*32)                                       */
    int origin;              /* offset of pixel (0,0) inside a padded plane               */
} x264vfw_cuda_hpel_geom;

void x264vfw_cuda_hpel_geometry( x264vfw_cuda_hpel_geom *g, int i_width, int i_height );

/* src_dev: tight i_width x i_height plane (upstream: 16*mb_w x 16*mb_h; i_width % 8 == 0 required).
 * dst_dev: four padded planes per frame, consecutive (4*plane_bytes): [0] the frame with its 32-pixel
 * replicated border, [1] H, [2] V, [3] centre half-pel planes, each complete including the border
 * (the filter is evaluated 8 pixels beyond the frame and replicated from 4 columns / 8 rows outside,
 * exactly as upstream).  Bit-exact to upstream's C hpel_filter.  n_frames frames per launch. */
int x264vfw_cuda_hpel_filter( x264vfw_cuda_ctx *ctx, uint8_t *dst_dev, const uint8_t *src_dev,
                              int src_stride, int i_width, int i_height,
                              size_t src_frame_bytes, size_t dst_frame_bytes, int n_frames );

/* ------------------------------------------------------------------------------------
 * B2: the lookahead session.  Replaces what happens between x264_encoder_encode receiving
 * a picture (codec.c:1693) and the frame type / per-MB qp offsets being known:
 * [x264] x264_adaptive_quant_frame, x264_frame_init_lowres, x264_lookahead_put_frame,
 * x264_slicetype_decide / x264_slicetype_analyse (scenecut, b-adapt 1/2, mb-tree) with
 * slicetype_frame_cost on the GPU.  The decisions are handed to the CPU encoder through
 * x264_picture_t.i_type and .prop.quant_offsets.
 * ---------------------------------------------------------------------------------- */
#define X264VFW_CUDA_TYPE_AUTO     0   /* X264_TYPE_* of x264.h */
#define X264VFW_CUDA_TYPE_IDR      1
#define X264VFW_CUDA_TYPE_I        2
#define X264VFW_CUDA_TYPE_P        3
#define X264VFW_CUDA_TYPE_BREF     4
#define X264VFW_CUDA_TYPE_B        5
#define X264VFW_CUDA_BFRAME_MAX    16  /* X264_BFRAME_MAX */
#define X264VFW_CUDA_LOOKAHEAD_MAX 250 /* X264_LOOKAHEAD_MAX */

/* The x264_param_t fields the lookahead reads, after x264_param_default_preset /
 * x264_param_parse / level resolution (codec.c:1463,1349,1584). */
typedef struct x264vfw_cuda_la_params
{
    int   width, height;
    int   chroma_format;       /* 1 = 4:2:0, 2 = 4:2:2, 3 = 4:4:4 (of the ENCODER csp)            */
    int   bframes;             /* i_bframe                                                       */
    int   b_adapt;             /* i_bframe_adaptive: 0 none, 1 fast, 2 trellis                   */
    int   b_pyramid;           /* i_bframe_pyramid: 0 none, 1 strict, 2 normal                   */
    int   b_bias;              /* i_bframe_bias                                                  */
    int   rc_lookahead;        /* rc.i_lookahead                                                 */
    int   b_mbtree;            /* rc.b_mb_tree                                                   */
    int   scenecut;            /* i_scenecut_threshold                                           */
    int   keyint_max, keyint_min;
    int   open_gop;
    int   weightp;             /* analyse.i_weighted_pred                                        */
    int   weightb;             /* analyse.b_weighted_bipred                                      */
    int   subme;               /* analyse.i_subpel_refine (selects the lookahead's search mode)  */
    int   me_method;           /* analyse.i_me_method                                            */
    int   me_range;            /* analyse.i_me_range                                             */
    int   mv_range;            /* analyse.i_mv_range after level resolution                      */
    int   aq_mode;             /* rc.i_aq_mode: 0 off, 1 variance, 2 auto-variance, 3 biased     */
    float aq_strength;
    float qcompress;
    int   frame_reference;     /* i_frame_reference                                              */
    int   lookahead_threads;   /* i_lookahead_threads: band split of the MB scan (pin for parity) */
    int   fps_num, fps_den;
    int   b_psy;
} x264vfw_cuda_la_params;

/* x264 defaults + the preset deltas documented at config.c:1460-1498.  Returns 0 / -1. */
int x264vfw_cuda_la_params_preset( x264vfw_cuda_la_params *p, const char *preset, int width, int height );
/* [x264] x264_param_apply_tune reduced to the lookahead's fields; call after _params_preset
 * (codec.c:1463 applies preset and tuning together).  film/none, animation, grain, stillimage,
 * psnr, ssim, fastdecode, zerolatency, touhou.  Returns 0, -1 for an unknown name. */
int x264vfw_cuda_la_params_tune( x264vfw_cuda_la_params *p, const char *tune );

typedef struct x264vfw_cuda_la x264vfw_cuda_la;

/* i_in_csp: X264VFW_CUDA_CSP_* (| VFLIP) of the frames passed to put_frame, converted on the
 * device to i_x264_csp first (stage 1), or X264VFW_CUDA_CSP_NONE when put_frame receives
 * planar frames already in the encoder csp.  keep_frames != 0 keeps every frame addressable
 * for x264vfw_cuda_la_read (tests); otherwise frames are recycled as the window slides.
 * device < 0: current device. */
int  x264vfw_cuda_la_open( x264vfw_cuda_la **pla, const x264vfw_cuda_la_params *params, int device,
                           int i_in_csp, int i_x264_csp, int i_colmatrix, int b_fullrange,
                           int keep_frames );
void x264vfw_cuda_la_close( x264vfw_cuda_la *la );

/* One input frame in display order == one x264_encoder_encode( pic_in ) call.
 * src describes the frame like x264vfw_img_fill does; src_on_device selects host (0) or device
 * (X264VFW_CUDA_SRC_DEVICE, X264VFW_CUDA_SRC_RESIDENT) pointers.  conv_pic (may be NULL) receives
 * the converted planes in HOST memory -- what codec->conv_pic holds for the CPU encoder.  Both
 * buffers are only borrowed for the call: the source has been read and conv_pic is complete when
 * it returns (for X264VFW_CUDA_SRC_DEVICE the call waits for the device-side conversion that reads
 * it).  The one exception is X264VFW_CUDA_SRC_RESIDENT: the caller promises that the device buffer
 * stays unmodified until x264vfw_cuda_la_flush / _close (a clip that lives in HBM), and the call
 * returns without waiting for its reader.  Returns the number of decided frames waiting in the
 * output queue, or -1.
 * Like libx264 behind sync-lookahead, the session works on queued frames in a thread of its
 * own while the caller moves the next frame: the frames that become decided because of frame n
 * are published, deterministically, by the call for frame n + 2 (and all of them by _flush);
 * the sequence of decisions never depends on timing.  X264VFW_CUDA_ASYNC=0 decides inside the
 * call instead. */
#define X264VFW_CUDA_SRC_HOST     0
#define X264VFW_CUDA_SRC_DEVICE   1
#define X264VFW_CUDA_SRC_RESIDENT 2
int x264vfw_cuda_la_put_frame( x264vfw_cuda_la *la, const x264vfw_cuda_image_t *src, int src_on_device,
                               x264vfw_cuda_image_t *conv_pic );
/* End of stream (x264_encoder_encode with pic_in == NULL, codec.c:1755-1758,1848). */
int x264vfw_cuda_la_flush( x264vfw_cuda_la *la );

typedef struct x264vfw_cuda_la_decision
{
    int i_frame;        /* display index                                                        */
    int i_type;         /* X264VFW_CUDA_TYPE_IDR/I/P/BREF/B -> x264_picture_t.i_type            */
    int b_keyframe;     /* -> pic_out.b_keyframe (codec.c:1824)                                 */
    int i_bframes;      /* on the non-B of a mini-GOP: number of B-frames before it             */
    int i_cost_est;     /* [x264] i_cost_est of the chosen (p0,p1,b); -1 when not computed      */
    int i_cost_est_aq;
    int i_intra_mbs;
    int mb_count;
} x264vfw_cuda_la_decision;

/* Pops the next decided frame in CODED order.  qp_offset / qp_offset_aq (may be NULL) receive
 * mb_count floats each: x264_frame_t.f_qp_offset (mb-tree) and .f_qp_offset_aq, the values
 * x264_picture_t.prop.quant_offsets is built from.  Returns 1, 0 when empty, -1 on error. */
int x264vfw_cuda_la_get_decision( x264vfw_cuda_la *la, x264vfw_cuda_la_decision *d,
                                  float *qp_offset, float *qp_offset_aq );

/* ---- B3-shaped white-box entry points (parity tests; frames addressed by display index) --
 * [x264] slicetype_frame_cost( frames, p0, p1, b ): returns the score, -1 on error. */
int x264vfw_cuda_la_frame_cost( x264vfw_cuda_la *la, int p0, int p1, int b );
/* [x264] macroblock_tree over explicit frames/types (frame_idx[0..num_frames]). */
int x264vfw_cuda_la_mbtree( x264vfw_cuda_la *la, const int *frame_idx, const int *types, int num_frames, int b_intra );
#define X264VFW_CUDA_LA_LOWRES        1  /* 4 padded planes, uint8                               */
#define X264VFW_CUDA_LA_INTRA_COST    2  /* uint16[mb]                                           */
#define X264VFW_CUDA_LA_INV_QSCALE    3  /* uint16[mb]                                           */
#define X264VFW_CUDA_LA_PROPAGATE     4  /* int32[mb] (unsaturated shadow of i_propagate_cost)   */
#define X264VFW_CUDA_LA_QP_OFFSET     5  /* float[mb]                                            */
#define X264VFW_CUDA_LA_QP_OFFSET_AQ  6  /* float[mb]                                            */
#define X264VFW_CUDA_LA_MVS           7  /* int16[mb][2]; a = list, b = distance (>= 1)          */
#define X264VFW_CUDA_LA_MV_COSTS      8  /* int32[mb];    a = list, b = distance                 */
#define X264VFW_CUDA_LA_LOWRES_COSTS  9  /* uint16[mb];   a = b-p0, b = p1-b                     */
#define X264VFW_CUDA_LA_COST_EST      10 /* int32[3]: cost_est, cost_est_aq, intra_mbs; a,b as 9 */
#define X264VFW_CUDA_LA_PIXEL_STATS   11 /* uint64[6]: sum[3], ssd[3] (after mean removal)       */
#define X264VFW_CUDA_LA_WEIGHT        12 /* int32[4]: scale, denom, offset, enabled              */
#define X264VFW_CUDA_LA_CONV_PLANES   13 /* converted planes of the LAST put frame (tight)       */
#define X264VFW_CUDA_LA_ROW_SATDS     14 /* int32[mb_h]: [x264] i_row_satds[a][b] (AQ-weighted)  */
/* Copies the selected array of frame `frame` into dst (capacity dst_bytes).  Returns the
 * number of bytes written or -1. */
int64_t x264vfw_cuda_la_read( x264vfw_cuda_la *la, int frame, int what, int a, int b, void *dst, size_t dst_bytes );
/* [0] slicetype_frame_cost evaluations launched, [1] MB searches (MBs x lists), [2] kernel
 * launches, [3] host<->device synchronisations, [4] host microseconds spent enqueueing frame
 * preparation, [5] in the decision logic (including [6]), [6] blocked in synchronisations,
 * [7] frames put */
void x264vfw_cuda_la_counters( x264vfw_cuda_la *la, uint64_t out[8] );

/* Per-kernel-class device time of this session, measured with CUDA events on the session's
 * stream around every launch: [0] csp [1] aq [2] lowres [3] intra [4] motion search
 * (ordered part: verification wavefront) [5] cost selection [6] weights [7] mb-tree
 * [8] motion search (speculative parallel passes) [9] fused front end (csp + adaptive quant + lowres in one kernel)
 * [10..15] reserved.  Returns the totals
 * accumulated so far in ms[16] / count[16] (may be NULL); enable = 1/0 switches measurement
 * on/off and resets the totals, -1 only reads. */
int x264vfw_cuda_la_profile( x264vfw_cuda_la *la, int enable, double ms[16], uint64_t count[16] );

/* ------------------------------------------------------------------------------------
 * B3: upstream libx264's own function-table shapes on DEVICE pointers.  libx264 is not part of the
 * reference tree (Makefile:21-23,109; entered at codec.c:1693); these mirror the C functions behind
 * x264_mc_functions_t / x264_pixel_function_t ([x264] common/mc.c, common/pixel.c) argument for argument, so that
 * a patch that points those table entries at the GPU is mechanical.  Differences are only what a device call
 * needs: the context (stream) in front, device pointers, a float instead of float*, and batches (offset arrays)
 * where upstream calls the function once per block.  All calls are asynchronous on the context's stream.
 * ---------------------------------------------------------------------------------- */
/* mc.frame_init_lowres_core( src0, dst0, dsth, dstv, dstc, src_stride, dst_stride, width, height ): width x height
 * lowres pixels, reads src rows 0 .. 2*height and columns 0 .. 2*width (the caller provides the duplicated last
 * row / column, as upstream's x264_frame_init_lowres does); no border. */
int x264vfw_cuda_frame_init_lowres_core( x264vfw_cuda_ctx *ctx, const uint8_t *src0, uint8_t *dst0, uint8_t *dsth,
                                         uint8_t *dstv, uint8_t *dstc, intptr_t src_stride, intptr_t dst_stride,
                                         int width, int height );
/* mc.mbtree_propagate_cost( dst, propagate_in, intra_costs, inter_costs, inv_qscales, fps_factor, len ) */
int x264vfw_cuda_mbtree_propagate_cost( x264vfw_cuda_ctx *ctx, int16_t *dst, const uint16_t *propagate_in,
                                        const uint16_t *intra_costs, const uint16_t *inter_costs,
                                        const uint16_t *inv_qscales, float fps_factor, int len );
/* mc.mbtree_propagate_list( h, ref_costs, mvs, propagate_amount, lowres_costs, bipred_weight, mb_y, len, list ):
 * one MB row; mvs / propagate_amount / lowres_costs point at that row, ref_costs at the whole frame; h->mb.i_mb_width /
 * i_mb_height travel as the last two arguments.  Saturating 16-bit adds (32767) like upstream. */
int x264vfw_cuda_mbtree_propagate_list( x264vfw_cuda_ctx *ctx, uint16_t *ref_costs, const int16_t (*mvs)[2],
                                        const int16_t *propagate_amount, const uint16_t *lowres_costs,
                                        int bipred_weight, int mb_y, int len, int list, int mb_width, int mb_height );
/* pixf.sad[PIXEL_8x8] / pixf.satd[PIXEL_8x8]( pix1, stride1, pix2, stride2 ) for n block pairs: pair i compares the
 * blocks at pix1 + off1[i] and pix2 + off2[i]; scores[i] receives the result. */
int x264vfw_cuda_pixel_cmp_8x8( x264vfw_cuda_ctx *ctx, int b_satd, const uint8_t *pix1, intptr_t stride1,
                                const uint8_t *pix2, intptr_t stride2, const int *off1, const int *off2,
                                int *scores, int n );
/* pixf.sad_x3 / sad_x4[PIXEL_8x8]( fenc, pix0..pix3, stride, scores ) for n source blocks: block i at
 * fenc + off_fenc[i] against n_ref (3 or 4) candidates at ref + off_ref[i * n_ref + k]; scores[i * n_ref + k]. */
int x264vfw_cuda_pixel_sad_xn_8x8( x264vfw_cuda_ctx *ctx, int n_ref, const uint8_t *fenc, intptr_t fenc_stride,
                                   const int *off_fenc, const uint8_t *ref, intptr_t ref_stride,
                                   const int *off_ref, int *scores, int n );
/* pixf.intra_mbcmp_x3_8x8c( fenc, fdec, res ) for every 8x8 block of a plane: predict_8x8c_{dc,h,v} from the block's
 * neighbours inside the plane (row above, column to the left must be valid memory: the lookahead's padded lowres
 * plane), scored with SAD or SATD; res[3 * (mx + my * mb_w) + {0,1,2}] = {DC, H, V}. */
int x264vfw_cuda_intra_mbcmp_x3_8x8c( x264vfw_cuda_ctx *ctx, int b_satd, const uint8_t *plane, int stride,
                                      int mb_w, int mb_h, int *res );

/* [x264] encoder/slicetype-cl.c hook names: what a libx264 patched at its HAVE_OPENCL sites (slicetype_frame_cost,
 * x264_slicetype_analyse) would call, here on a COST-ENGINE session: x264vfw_cuda_la_open with keep_frames = 1 and
 * params.rc_lookahead = X264VFW_CUDA_LOOKAHEAD_MAX, so that the session never decides anything itself and libx264's own
 * decision code keeps the control flow.  Frames are addressed by display index = order of the lowres_init calls.
 *   x264_opencl_lowres_init( h, fenc, lambda )                        -> _opencl_lowres_init: returns the frame's index
 *   x264_opencl_motionsearch( h, frames, b, ref, b_islist1, lambda, w ) -> _opencl_motionsearch: enqueues the search
 *   x264_opencl_finalize_cost( h, lambda, frames, p0, p1, b, dsf )    -> _opencl_finalize_cost: slicetype_frame_cost(p0,p1,b)
 *                                                                        incl. the weight analysis; cost_out = {i_cost_est,
 *                                                                        i_cost_est_aq, i_intra_mbs}; returns the score
 *   x264_opencl_flush( h )                                            -> _opencl_flush
 *   x264_opencl_slicetype_prep( h, frames, num_frames, lambda )       -> _opencl_slicetype_prep( first, num_frames ): batches
 *                                                                        every search of the window up to bframes away
 *   x264_opencl_slicetype_end( h )                                    -> _opencl_slicetype_end */
int x264vfw_cuda_opencl_lowres_init( x264vfw_cuda_la *la, const x264vfw_cuda_image_t *src, int src_on_device );
int x264vfw_cuda_opencl_motionsearch( x264vfw_cuda_la *la, int b, int ref, int b_islist1 );
int x264vfw_cuda_opencl_finalize_cost( x264vfw_cuda_la *la, int p0, int p1, int b, int cost_out[3] );
int x264vfw_cuda_opencl_flush( x264vfw_cuda_la *la );
int x264vfw_cuda_opencl_slicetype_prep( x264vfw_cuda_la *la, int first, int num_frames );
int x264vfw_cuda_opencl_slicetype_end( x264vfw_cuda_la *la );

/* Work counters of the search kernels and the mb-tree, counted on the device while x264vfw_cuda_la_profile
 * is enabled (reset when it is switched on): [0] MBs whose speculative result the ordered verification kept,
 * [1] MBs it searched again in order, [2..5] MBs searched by parallel pass 0..3, [6] SAD 8x8 evaluations,
 * [7] SATD 8x8 evaluations (all search kernels), [8] mb-tree steps run by tree_chain_kernel, [9] mb-tree walks,
 * [10] searches launched speculatively, [11] searches launched on demand, [12] on-demand launches ([10]..[14] count
 * since the session was opened), [13] searches the decision logic actually asked for (upstream's count), [14] frames
 * put.  Returns 0 / -1. */
int x264vfw_cuda_la_stats( x264vfw_cuda_la *la, uint64_t out[16] );

/* ---- SURVEY 8(f) row 3, remainder: the encoder-side weight analysis and the integral image -------------------------
 * [x264] x264_weights_analyse( h, fenc, ref, 0 ) (encoder/slicetype.c; reached from x264_encoder_encode, codec.c:1693, when a P
 * frame is about to be coded with weightp >= 1): explicit weights {on, scale, denom, offset} for luma, U and V of `fenc`
 * against its nearest reference.  Luma is scored on the lowres planes with the reference compensated by the lookahead's list-0
 * vectors of that distance, chroma at full resolution on NV12 planes; the (scale, offset) window grows with subme.  All
 * pointers are DEVICE pointers; the two lowres buffers are the 4 padded planes x264vfw_cuda_lowres_init lays out (what
 * x264vfw_cuda_la_read(LA_LOWRES) returns), the chroma planes what x264vfw_cuda_chroma_nv12_pad writes (mod-16 padded NV12,
 * 4:2:0), the statistics [x264] i_pixel_sum / i_pixel_ssd (LA_PIXEL_STATS).  lowres_mvs == NULL is upstream's 0x7FFF
 * sentinel: the lookahead never searched that (frame, distance), the reference is used uncompensated. */
typedef struct x264vfw_cuda_weights_in
{
    int width, height;                 /* display size */
    const uint8_t  *fenc_lowres, *ref_lowres;
    const int16_t  *lowres_mvs;        /* int16[mb][2], quarter-pel lowres vectors of (fenc, list 0, fenc - ref), or NULL */
    const uint16_t *intra_cost;        /* uint16[mb]: fenc->i_intra_cost */
    const uint8_t  *fenc_uv, *ref_uv;
    int uv_stride;
    uint64_t fenc_sum[3], fenc_ssd[3], ref_sum[3], ref_ssd[3];
    int subme;                         /* analyse.i_subpel_refine: the search window and sad / satd */
    int weightp;                       /* analyse.i_weighted_pred; -1 (X264_WEIGHTP_FAKE) also reports cost_delta */
} x264vfw_cuda_weights_in;
int x264vfw_cuda_weights_analyse( x264vfw_cuda_ctx *ctx, const x264vfw_cuda_weights_in *in, int32_t out[3][4], float *cost_delta );
/* The same on a lookahead session opened with keep_frames: lowres planes, vectors, intra costs and statistics of display
 * indices `fenc` and `ref` come from the session, the caller adds the two chroma planes. */
int x264vfw_cuda_la_weights_analyse( x264vfw_cuda_la *la, int fenc, int ref, const uint8_t *fenc_uv_dev, const uint8_t *ref_uv_dev,
                                     int uv_stride, int32_t out[3][4], float *cost_delta );

/* [x264] the integral image x264_frame_filter builds behind the half-pel planes for the exhaustive searches (me esa / tesa;
 * common/mc.c integral_init8h + 8v, and 4h + 4v for the 4x4 plane of --partitions p4x4): for every position of a PADDED
 * plane (plane_dev = its top-left corner, `rows` rows of `stride` bytes, e.g. plane 0 of x264vfw_cuda_hpel_filter's output)
 * sum8[y*stride + x] = sum of the 8x8 pixels whose top-left is (x, y), modulo 2^16 like upstream's uint16 arithmetic; sum4
 * (may be NULL) the same for 4x4.  Written for y <= rows-8 (rows-4), x <= stride-9 (stride-5) -- everything the search can
 * address; the remaining entries are left untouched.  n_frames planes per launch (+f*plane_bytes / +f*sum_elems). */
int x264vfw_cuda_integral_init( x264vfw_cuda_ctx *ctx, uint16_t *sum8_dev, uint16_t *sum4_dev, const uint8_t *plane_dev,
                                int stride, int rows, size_t plane_bytes, size_t sum_elems, int n_frames );

/* ---- B1b: decoder-side output conversion ---------------------------------------------
 * Replaces the libswscale pair of the reference's decompress path: x264vfw_init_sws_context (codec.c:2075-2152,
 * called lazily at codec.c:2282-2290) becomes x264vfw_cuda_dec_open, sws_scale (codec.c:2292) becomes
 * x264vfw_cuda_dec_convert, sws_freeContext (codec.c:2306) becomes x264vfw_cuda_dec_close.  Source: one decoded picture
 * (data[]/linesize[] of the AVFrame), i_src_chroma = 1 for AV_PIX_FMT_YUV420P / YUVJ420P, 2 for YUV422P / YUVJ422P (High
 * 4:2:2 streams, what "keep input colorspace" makes of YUY2 / UYVY input), 3 for YUV444P / YUVJ444P (High 4:4:4 Predictive,
 * what it makes of YV24 input).  Destination: the output DIB
 * laid out by x264vfw_picture_fill (codec.c:419-503), i_out_csp = what get_csp() returns for the OUTPUT header
 * (codec.c:1994-1998): X264VFW_CUDA_CSP_{I420,YV12,YV16,YV24,NV12,YUYV,UYVY,BGR,BGRA}, | X264VFW_CUDA_CSP_VFLIP for bottom-up
 * RGB (x264vfw_picture_vflip, codec.c:510-527); the U/V swap of codec.c:2263-2274 (YV12 / YV16 / YV24) is applied inside.
 * i_avcol_spc: decoder_context->colorspace (AVCOL_SPC_*, the switch of codec.c:2114-2140); b_fullrange: color_range ==
 * AVCOL_RANGE_JPEG or a YUVJ pixel format (codec.c:2091-2095).  Results are byte-identical to libswscale 9.1.100 (x86-64)
 * driven that way -- including that SWS_FULL_CHR_H_INT never reaches the context (codec.c:2097 vs :2110), so RGB output
 * shares one chroma sample per pixel pair -- except that exactly width pixels per row are written (libswscale's SIMD
 * writers store groups of 8).  4:2:2 pictures have no vertical chroma filter: YUY2 / UYVY / YV16 are plain (de)interleaves,
 * RGB uses libswscale's single-line writers; for 4:4:4 pictures libswscale interpolates nothing and converts pixel by pixel
 * (its full-chroma C writer), YV24 is a plane copy.  A YUV output whose chroma resolution differs from the picture's (4:2:0 ->
 * YV16 / YV24, 4:2:2 -> I420 / YV12 / NV12 / YV24, 4:4:4 -> I420 / YV12 / NV12 / YV16 / YUY2 / UYVY) runs libswscale's bicubic chroma
 * scaler (4 taps for 2x up, 8 taps for 2:1 down).  Every (picture format, output csp) pair of 8-bit 4:2:0 / 4:2:2 / 4:4:4 pictures
 * is covered.  Not covered (open returns -1): pictures below 12 rows; chroma planes below 12 (6) samples in a direction that is
 * halved (doubled). */
typedef struct x264vfw_cuda_dec x264vfw_cuda_dec;
int  x264vfw_cuda_dec_open( x264vfw_cuda_dec **pdec, x264vfw_cuda_ctx *ctx, int i_out_csp, int i_width, int i_height,
                            int i_src_chroma, int i_avcol_spc, int b_fullrange );
void x264vfw_cuda_dec_close( x264vfw_cuda_dec *dec );
/* x264vfw_picture_get_size (codec.c:505-508) for the covered formats, -1 otherwise. */
int64_t x264vfw_cuda_dec_picture_size( int i_out_csp, int i_width, int i_height );
/* One picture, HOST buffers (the sws_scale call): copies the planes up, converts, copies the DIB back, returns when
 * dst_host holds x264vfw_cuda_dec_picture_size() bytes.  0 / -1. */
int x264vfw_cuda_dec_convert( x264vfw_cuda_dec *dec, uint8_t *dst_host, const uint8_t *const src_host[3],
                              const int src_stride[3] );
/* n_frames pictures, DEVICE buffers, one launch (two where the chroma resolution changes), asynchronous on the context's stream: picture f reads its planes at
 * src_dev[i] + f*src_frame_bytes and writes its DIB at dst_dev + f*dst_frame_bytes.  The entry the roofline number of
 * this stage is measured on. */
int x264vfw_cuda_dec_convert_batch( x264vfw_cuda_dec *dec, uint8_t *dst_dev, size_t dst_frame_bytes,
                                    const uint8_t *const src_dev[3], const int src_stride[3], size_t src_frame_bytes,
                                    int n_frames );

/* Host-only views of the tables those kernels are handed (no device needed; the CPU test suite holds them against the checker):
 * the tap table of a chroma plane resampled from src_n to dst_n samples (dst_n = 2 src_n, src_n = 2 dst_n or equal; one = 1 << 14
 * with align 4 horizontally, 1 << 12 with align 2 vertically) as pos[dst_n] + coef[dst_n][8]; and the per-row table of the
 * packed writers for a 4:2:0 picture with chroma_rows chroma lines: pos[2n], coef[2n][4] as the writer of that row sees them
 * (libswscale's SIMD writers pack coefficient pairs into one int, so a negative even tap borrows 1 from the odd one), and which
 * rows take the C writer.  0 / -1 (sizes the kernels do not cover). */
int x264vfw_cuda_dec_filter_taps( int src_n, int dst_n, int one, int align, int32_t *pos, int16_t *coef );
int x264vfw_cuda_dec_packed_rows( int chroma_rows, int b_uyvy, int32_t *pos, int16_t *coef, int32_t *c_writer );

const char *x264vfw_cuda_last_error( void );
/* "x264vfw_cuda <version> sm_100a"; also proves the library loaded. */
const char *x264vfw_cuda_version( void );
/* Number of kernels this library has launched in the calling process (all handles). */
uint64_t x264vfw_cuda_launch_count( void );

#ifdef __cplusplus
}
#endif
#endif /* X264VFW_CUDA_H */
