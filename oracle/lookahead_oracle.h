/*
 * lookahead_oracle.h -- CPU restatement of libx264's lookahead cost engine.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  PARITY UNPINNED: libx264 is an external,
 * un-vendored, un-pinned dependency of the reference (reference Makefile:21-23,109); the
 * only call site is x264_encoder_encode at codec.c:1693.  Every function cites the upstream
 * libx264 function it follows as "[x264] path: function".
 */
#ifndef X264VFW_LOOKAHEAD_ORACLE_H
#define X264VFW_LOOKAHEAD_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_BFRAME_MAX 16
#define ORC_LOOKAHEAD_MAX 250

/* frame types: public X264_TYPE_* values of x264.h */
#define ORC_TYPE_AUTO 0
#define ORC_TYPE_IDR 1
#define ORC_TYPE_I 2
#define ORC_TYPE_P 3
#define ORC_TYPE_BREF 4
#define ORC_TYPE_B 5
#define ORC_TYPE_KEYFRAME 6
#define ORC_WEIGHTP_FAKE (-1)   /* X264_WEIGHTP_FAKE of x264.h */

/* Plain-int parameter block (x264_param_t subset the lookahead reads).  Field order is
 * shared with x264vfw_cuda_la_params in include/x264vfw_cuda.h. */
typedef struct orc_la_params {
    int   width, height;
    int   chroma_format;       /* 1 = 4:2:0, 2 = 4:2:2, 3 = 4:4:4 (AQ chroma energy)          */
    int   bframes;             /* i_bframe                                                   */
    int   b_adapt;             /* i_bframe_adaptive: 0 none, 1 fast, 2 trellis               */
    int   b_pyramid;           /* i_bframe_pyramid: 0 none, 1 strict, 2 normal               */
    int   b_bias;              /* i_bframe_bias                                              */
    int   rc_lookahead;        /* rc.i_lookahead                                             */
    int   b_mbtree;            /* rc.b_mb_tree                                               */
    int   scenecut;            /* i_scenecut_threshold                                       */
    int   keyint_max, keyint_min;
    int   open_gop;
    int   weightp;             /* analyse.i_weighted_pred                                    */
    int   weightb;             /* analyse.b_weighted_bipred                                  */
    int   subme;               /* analyse.i_subpel_refine of the ENCODER (selects lookahead mode) */
    int   me_method;           /* analyse.i_me_method (0 dia, 1 hex, ...)                    */
    int   me_range;            /* analyse.i_me_range                                         */
    int   mv_range;            /* analyse.i_mv_range after level resolution (512 for level>=3.1) */
    int   aq_mode;             /* rc.i_aq_mode: 0 off, 1 variance, 2 auto-variance, 3 auto-variance biased                            */
    float aq_strength;
    float qcompress;
    int   frame_reference;     /* i_frame_reference                                          */
    int   lookahead_threads;   /* i_lookahead_threads (band split of the MB scan)            */
    int   fps_num, fps_den;
    int   b_psy;               /* analyse.b_psy (with mbtree: analyse past keyint)           */
} orc_la_params;

void orc_la_params_preset(orc_la_params *p, const char *preset, int width, int height);

typedef struct orc_la orc_la;

orc_la *orc_la_open(const orc_la_params *p);
void    orc_la_close(orc_la *la);
/* Feed one frame in display order (tight planar input in the encoder csp; u/v may be NULL
 * for luma-only experiments -> chroma energy 0).  Returns number of decided frames now
 * waiting in the output queue. */
int orc_la_put_frame(orc_la *la, const uint8_t *y, int y_stride, const uint8_t *u, const uint8_t *v, int c_stride);
/* End of stream: decide everything still queued ([x264] lookahead flush). */
int orc_la_flush(orc_la *la);

typedef struct orc_la_decision {
    int i_frame;        /* display index                                                     */
    int i_type;         /* ORC_TYPE_IDR/I/P/BREF/B                                           */
    int b_keyframe;
    int i_bframes;      /* for the non-B of a mini-GOP: number of B-frames before it         */
    int i_cost_est;     /* slicetype_frame_cost of the chosen (p0,p1,b); -1 if not computed  */
    int i_cost_est_aq;
    int i_intra_mbs;
    int mb_count;
} orc_la_decision;

/* Pop the next decided frame (coded order).  qp_offset / qp_offset_aq receive mb_count floats
 * each when non-NULL (x264_frame_t.f_qp_offset / f_qp_offset_aq).  Returns 1, or 0 if empty. */
int orc_la_get_decision(orc_la *la, orc_la_decision *d, float *qp_offset, float *qp_offset_aq);

/* ---- white-box access for kernel-level parity tests ------------------------------------- */
/* All frames ever fed stay addressable by display index until orc_la_close (memory grows
 * with clip length; the oracle is for short clips). */
int  orc_la_mb_count(orc_la *la);
/* run slicetype_frame_cost(p0,p1,b) on explicit display indices (frames[] = identity map) */
int  orc_la_frame_cost(orc_la *la, int p0, int p1, int b);
const uint8_t  *orc_la_lowres_planes(orc_la *la, int frame);                 /* 4 padded planes */
const uint16_t *orc_la_intra_cost(orc_la *la, int frame);
const uint16_t *orc_la_inv_qscale(orc_la *la, int frame);
const uint16_t *orc_la_propagate_cost(orc_la *la, int frame);
const float    *orc_la_qp_offset(orc_la *la, int frame, int aq);
const int16_t  *orc_la_mvs(orc_la *la, int frame, int list, int dist);       /* [mb][2], dist>=1 */
const int      *orc_la_mv_costs(orc_la *la, int frame, int list, int dist);
const uint16_t *orc_la_lowres_costs(orc_la *la, int frame, int d0, int d1);
const int      *orc_la_row_satds(orc_la *la, int frame, int d0, int d1);     /* [mb_h] */
int  orc_la_cost_est(orc_la *la, int frame, int d0, int d1, int aq);
int  orc_la_intra_mbs(orc_la *la, int frame, int d0);
void orc_la_pixel_stats(orc_la *la, int frame, uint64_t sum[3], uint64_t ssd[3]);
/* last weight found by the lookahead weight analysis for frame (scale, denom, offset, enabled) */
void orc_la_weight(orc_la *la, int frame, int out[4]);
const uint16_t *orc_cost_mv_table(int mv_range, int *half_len);
/* run one mb-tree pass over explicit display indices with explicit types (for kernel parity) */
void orc_la_mbtree(orc_la *la, const int *frame_idx, const int *types, int num_frames, int b_intra);
/* [x264] x264_weights_analyse(h, fenc, ref, 0), the encoder-side weight analysis of P frame `fenc`
 * against display index `ref` (luma on the lowres planes compensated by the lookahead's vectors, chroma at full
 * resolution on NV12 planes padded to mod 16).  out[plane] = {on, scale, denom, offset}; -1 when not 4:2:0. */
int orc_la_weights_full(orc_la *la, int fenc, int ref, const uint8_t *fenc_uv, const uint8_t *ref_uv, int uv_stride,
                        int out[3][4], float *cost_delta);
/* [x264] integral_init8h/8v (+4h/4v): upstream's recurrences on a padded plane; see the definition for the valid domain */
void orc_integral_init(uint16_t *sum8, uint16_t *sum4, const uint8_t *plane, int stride, int rows);
/* counters: number of MB motion searches / SAD / SATD evaluations performed so far */
void orc_la_counters(orc_la *la, uint64_t out[4]);

/* Test hooks on two restated [x264] pieces that a standard H.264 DECODER must reproduce, so that they can be
 * pinned against one (tests/golden/make_h264_pins.py, tests/test_h264_pins.py):
 *  - get_ref ([x264] common/mc.c): the 8x8 block at quarter-sample vector (mvx, mvy) from four half-pel planes
 *    (full, H, V, centre), each pointer already at the block's integer position;
 *  - the intra predictors of the lookahead ([x264] common/predict.c): kind 0..3 = predict_8x8c_{dc,h,v,p},
 *    kind 13..18 = predict_8x8_{ddl,ddr,vr,hd,vl,hu} after predict_8x8_filter; src = the block's top-left sample
 *    inside a plane that holds its neighbours (row above incl. 8 samples to the right, column to the left). */
void orc_test_get_ref_8x8(uint8_t dst[64], const uint8_t *p0, const uint8_t *p1, const uint8_t *p2, const uint8_t *p3,
                          int stride, int mvx, int mvy);
void orc_test_intra_pred_8x8(uint8_t dst[64], int kind, const uint8_t *src, int stride);
/* get_ref followed by the explicit weight (mc_weight: ((p * scale + 2^(denom-1)) >> denom) + offset, clipped),
 * the decoder's explicit weighted sample prediction (H.264 8.4.2.3) */
void orc_test_get_ref_8x8_weighted(uint8_t dst[64], const uint8_t *p0, const uint8_t *p1, const uint8_t *p2, const uint8_t *p3,
                                   int stride, int mvx, int mvy, int scale, int denom, int offset);
/* the bidirectional average of the lookahead: weight of the list-0 block for frames (p0, b, p1) in display order
 * (weightb = implicit weights, H.264 8.4.2.3.1/2) and pixel_avg with that weight */
/* a whole w x h picture (multiples of 8) predicted with one vector from four padded planes (32-sample border,
 * plane p at planes + p * plane_bytes): get_ref per 8x8 block; scale < 0 = unweighted */
void orc_test_predict_picture(uint8_t *dst, const uint8_t *planes, int stride, size_t plane_bytes, int w, int h,
                              int mvx, int mvy, int scale, int denom, int offset);
int orc_test_bipred_weight(int p0, int p1, int b, int weightb);
/* the two block metrics of the search ([x264] common/pixel.c: sad 8x8, satd 8x8), strides 8 */
int orc_test_sad_8x8(const uint8_t a[64], const uint8_t b[64]);
int orc_test_satd_8x8(const uint8_t a[64], const uint8_t b[64]);
void orc_test_pixel_avg_8x8(uint8_t dst[64], const uint8_t a[64], const uint8_t b[64], int weight);

#ifdef __cplusplus
}
#endif
#endif
