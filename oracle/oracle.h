/*
 * oracle.h -- CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product (x264vfw_b200/, libx264vfw_cuda.so) never does.
 *
 * Stage 1 (csp_oracle.c) restates the reference's csp.c and is PINNED: it is checked
 * byte-for-byte against the unmodified reference object (oracle/_ref/libref_csp.so, built
 * from /root/reference/csp.c by oracle/Makefile) and against the golden hashes in
 * tests/golden/csp_golden.json.
 *
 * Stage 2 (lookahead_oracle.c) restates upstream libx264's lookahead, which the reference
 * only links against (Makefile:21-23,109) and does not vendor or version-pin:
 * PARITY UNPINNED -- no reference source, binary, test or golden vector exists for it.
 */
#ifndef X264VFW_ORACLE_H
#define X264VFW_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_image_t {
    int      i_csp;       /* input: X264VFW_CSP_* (csp.h:30-44) | 0x1000 vflip */
    int      i_plane;
    int      i_stride[4];
    uint8_t *plane[4];
} orc_image_t;

/* csp.c:440-514 dispatch + the selected converter.  Returns 0, or -1 where the reference
 * table holds convert_fail.  out_csp uses the public X264_CSP_* values. */
int orc_csp_convert(int out_csp, int colmatrix, int fullrange,
                    orc_image_t *dst, const orc_image_t *src, int w, int h);
/* The 12 fixed-point coefficients csp.c:252-297 evaluates to, in the order
 * y_r y_g y_b y_add u_r u_g u_b u_add v_r v_g v_b v_add. */
void orc_rgb_coefficients(int colmatrix, int fullrange, uint32_t out[12]);

/* The two BASELINE.json conversions csp.c does not register; definitions in DESIGN.md. */
int orc_ext_rgb_to_nv12(int colmatrix, int fullrange, orc_image_t *dst, const orc_image_t *src, int w, int h);
int orc_ext_422_to_i444(orc_image_t *dst, const orc_image_t *src, int w, int h);

uint64_t orc_fnv1a64(const uint8_t *p, size_t n);
/* SURVEY.md A.4 byte generator: s = 0x264+31w+h; s = s*1664525+1013904223; byte = s>>24 */
void orc_lcg_fill(uint8_t *p, size_t n, int w, int h);

/* ---- stage 2a: lowres planes (lowres_oracle.c).  g[10] = mb_w mb_h luma_w luma_h
 * luma_stride lw lh lstride lplane_bytes lorigin (same geometry as the CUDA library). */
void orc_lowres_geometry(int w, int h, int g[10]);
void orc_luma_pad(uint8_t *dst, int dst_stride, const uint8_t *y, int y_stride, int w, int h);
void orc_chroma_nv12_pad(uint8_t *dst, int dst_stride, const uint8_t *u, const uint8_t *v, int c_stride, int w, int h);
void orc_lowres_init(uint8_t *dst4planes, const uint8_t *y, int y_stride, int w, int h);

/* ---- next row f3: half-pel reference planes (hpel_oracle.c).  g[3] = stride plane_bytes origin. */
void orc_hpel_geometry(int w, int h, int g[3]);
void orc_hpel_planes(uint8_t *dst4planes, const uint8_t *src, int src_stride, int w, int h);

#ifdef __cplusplus
}
#endif
#endif
