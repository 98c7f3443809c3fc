/* Minimal <x264.h> stand-in: just the public libx264 types/constants that the
 * reference's csp.c and its headers name.  libx264 itself is NOT in the reference
 * tree (reference Makefile:21-23,109), so this declares only the public-API shapes. */
#ifndef ORACLE_SHIM_X264_H
#define ORACLE_SHIM_X264_H
#include <stdint.h>
typedef struct x264_t x264_t;
typedef struct { int i_csp; int i_plane; int i_stride[4]; uint8_t *plane[4]; } x264_image_t;
typedef struct { int i_type; int i_qpplus1; int i_pic_struct; int b_keyframe; int64_t i_pts; int64_t i_dts;
                 void *param; x264_image_t img; } x264_picture_t;
typedef struct { int dummy; } x264_param_t;
typedef struct { int i_ref_idc; int i_type; int b_long_startcode; int i_first_mb; int i_last_mb;
                 int i_payload; uint8_t *p_payload; int i_padding; } x264_nal_t;
#define X264_CSP_I420 0x0002
#define X264_CSP_NV12 0x0004
#define X264_CSP_I422 0x0006
#define X264_CSP_I444 0x000c
#define X264_CSP_BGR  0x000e
#define X264_CSP_BGRA 0x000f
#endif
