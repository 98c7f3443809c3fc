/*
 * lookahead_oracle.c -- scalar CPU restatement of libx264's lookahead:
 * adaptive-quant statistics, lowres planes, intra SATD, lowres motion search (SAD/SATD),
 * slicetype_frame_cost, weight analysis, frame-type decision and mb-tree.
 *
 * TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see lookahead_oracle.h): written from the
 * published upstream algorithm; function-by-function citations are "[x264] file: function".
 * Four pieces are pinned all the same, against libavcodec's H.264 decoder, because the standard fixes them
 * (tests/test_h264_pins.py): get_ref_8x8 (which planes a quarter-sample position averages), weight_px (explicit
 * weighted prediction), pixel_avg_8x8 with la_bipred_weight (implicit bidirectional weights) and the ten intra
 * predictors (pred_8x8c_*, filter_edges, pred_8x8_mode).  Search order, costs, decisions and mb-tree are not.
 * The 8-bit, progressive, non-VBV, single-pass paths are restated (the only ones the
 * reference's presets reach through codec.c:1693 for the BASELINE configs).
 */
#include "lookahead_oracle.h"
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define PAD 32
#define COST_MAX (1 << 28)
#define COST_MAX64 (1ULL << 60)
#define LOWRES_COST_MASK ((1 << 14) - 1)
#define LOWRES_COST_SHIFT 14
#define MBTREE_PRECISION 0.5f
#define IS_B(t) ((t) == ORC_TYPE_B || (t) == ORC_TYPE_BREF)
#define IS_I(t) ((t) == ORC_TYPE_I || (t) == ORC_TYPE_IDR || (t) == ORC_TYPE_KEYFRAME)
#define AUTO_OR_I(t) ((t) == ORC_TYPE_AUTO || IS_I(t))
#define AUTO_OR_B(t) ((t) == ORC_TYPE_AUTO || IS_B(t))
#define MIN(a, b) ((a) < (b) ? (a) : (b))
#define MAX(a, b) ((a) > (b) ? (a) : (b))

static inline int clip3(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }
static inline float clip3f(float v, float lo, float hi) { return v < lo ? lo : v > hi ? hi : v; }
static inline int clip_pixel(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }
static inline int median3(int a, int b, int c)
{
    int t = (a - b) & ((a - b) >> 31); a -= t; b += t;        /* [x264] common/common.h: x264_median */
    b -= (b - c) & ((b - c) >> 31);
    b += (a - b) & ((a - b) >> 31);
    return b;
}

typedef struct { int on, scale, denom, offset; } weight_t;

typedef struct frame_t {
    int i_frame, i_type, i_forced_type, b_scenecut, b_keyframe, i_bframes;
    float f_duration;
    uint8_t *lowres_buf;            /* 4 padded planes */
    uint8_t *lowres[4];             /* -> pixel (0,0) of each plane */
    uint64_t pixel_sum[3], pixel_ssd[3];
    uint16_t *intra_cost, *inv_qscale, *propagate_cost;
    float *qp_offset, *qp_offset_aq;
    int16_t (*mvs[2][ORC_BFRAME_MAX + 1])[2];
    int *mv_costs[2][ORC_BFRAME_MAX + 1];
    int mvs_searched[2][ORC_BFRAME_MAX + 1];       /* == (lowres_mvs[l][d][0][0] != 0x7FFF) */
    uint16_t *lowres_costs[ORC_BFRAME_MAX + 2][ORC_BFRAME_MAX + 2];
    int cost_est[ORC_BFRAME_MAX + 2][ORC_BFRAME_MAX + 2];
    int cost_est_aq[ORC_BFRAME_MAX + 2][ORC_BFRAME_MAX + 2];
    int intra_mbs[ORC_BFRAME_MAX + 2];
    int *row_satds[ORC_BFRAME_MAX + 2][ORC_BFRAME_MAX + 2];
    float weighted_cost_delta[ORC_BFRAME_MAX + 2];
    weight_t weight;                /* fenc->weight[0][0] after the last lookahead analysis */
    int b_intra_calculated;
    int rc_d0, rc_d1;               /* (b-p0, p1-b) precomputed for ratecontrol at decide time */
} frame_t;

struct orc_la {
    orc_la_params p;
    int mb_w, mb_h, mb_count;
    int luma_w, luma_h, lw, lh, lstride, lplane, lorigin;
    uint16_t *cost_mv;              /* centred */
    int cost_mv_half;
    /* lowres analysis context ([x264] lowres_context_init) */
    int la_me_method, la_subpel_refine, la_satd;
    /* all frames by display index */
    frame_t **all; int n_all, cap_all;
    /* [x264] lookahead->next, last_nonb, i_last_keyframe */
    frame_t *next[ORC_LOOKAHEAD_MAX + ORC_BFRAME_MAX + 8]; int n_next;
    frame_t *last_nonb;
    int i_last_keyframe;
    int slicetype_length;
    /* output queue (coded order) */
    frame_t **outq; int n_out, cap_out, out_head;
    uint8_t *weight_buf;            /* h->mb.p_weight_buf[0]: one padded lowres plane */
    int16_t *scratch_amount;
    uint64_t n_search, n_sad, n_satd, n_mbcost;
    /* per-call band limits ([x264] h->i_threadslice_start/end) */
    int slice_start, slice_end;
    /* per-MB mv limits ([x264] h->mb.mv_min_spel etc.) */
    int mv_min_spel[2], mv_max_spel[2], mv_limit_fpel[2][2];
};

/* ------------------------------------------------------------------------------------------
 * tables ([x264] common/tables.c).  x264_log2_lut[i] = log2(1+i/128) to 5 decimals,
 * x264_exp2_lut[i] = round(256*(2^(i/64)-1)): regenerated from their defining formulas.
 * ---------------------------------------------------------------------------------------- */
static float g_log2_lut[128];
static uint8_t g_exp2_lut[64];
static int g_tables_ready;
static void init_tables(void)
{
    if (g_tables_ready) return;
    for (int i = 0; i < 128; i++) g_log2_lut[i] = (float)(round(log2(1.0 + i / 128.0) * 100000.0) / 100000.0);
    for (int i = 0; i < 64; i++) g_exp2_lut[i] = (uint8_t)lround(256.0 * (pow(2.0, i / 64.0) - 1.0));
    g_tables_ready = 1;
}
static inline float x264_log2(uint32_t x)
{
    int lz = __builtin_clz(x);                                          /* [x264] common/base.h: x264_log2 */
    return g_log2_lut[(x << lz >> 24) & 0x7f] + (float)(31 - lz);
}
static inline int x264_exp2fix8(float x)
{
    int i = x * (-64.f / 6.f) + 512.5f;                                 /* [x264] common/base.h: x264_exp2fix8 */
    if (i < 0) return 0;
    if (i > 1023) return 0xffff;
    return (g_exp2_lut[i & 63] + 256) << (i >> 6) >> 8;
}

/* [x264] encoder/analyse.c: x264_analyse_init_costs / init_costs for X264_LOOKAHEAD_QP
 * (lambda 1).  Index range +-2*4*mv_range. */
#include <pthread.h>
static pthread_mutex_t g_tab_lock = PTHREAD_MUTEX_INITIALIZER;
const uint16_t *orc_cost_mv_table(int mv_range, int *half_len)
{
    static uint16_t *tab; static int tab_range;
    pthread_mutex_lock(&g_tab_lock);
    if (!tab || tab_range != mv_range) {
        /* a previous table may still be in use by an open session: leak it rather than free */
        int n = 2 * 4 * mv_range;
        tab = malloc((2 * n + 1) * sizeof(uint16_t));
        for (int i = 0; i <= n; i++) {
            float lg = i == 0 ? 0.718f : log2f((float)(i + 1)) * 2.0f + 1.718f;
            int c = (int)(1 * lg + .5f);
            tab[n - i] = tab[n + i] = (uint16_t)MIN(c, 65535);
        }
        tab_range = mv_range;
    }
    pthread_mutex_unlock(&g_tab_lock);
    if (half_len) *half_len = 2 * 4 * mv_range;
    return tab;
}

/* ------------------------------------------------------------------------------------------
 * pixel metrics ([x264] common/pixel.c)
 * ---------------------------------------------------------------------------------------- */
static int sad_8x8(const uint8_t *a, int sa, const uint8_t *b, int sb)
{
    int s = 0;
    for (int y = 0; y < 8; y++, a += sa, b += sb)
        for (int x = 0; x < 8; x++) s += abs(a[x] - b[x]);
    return s;
}
/* x264_pixel_satd_8x4: two 4x4 Hadamards, sum of |coeff|, halved once per 8x4 */
static int satd_8x4(const uint8_t *a, int sa, const uint8_t *b, int sb)
{
    int sum = 0;
    for (int blk = 0; blk < 2; blk++) {
        int d[4][4], t[4][4];
        for (int y = 0; y < 4; y++)
            for (int x = 0; x < 4; x++) d[y][x] = a[y * sa + 4 * blk + x] - b[y * sb + 4 * blk + x];
        for (int y = 0; y < 4; y++) {
            int s01 = d[y][0] + d[y][1], d01 = d[y][0] - d[y][1], s23 = d[y][2] + d[y][3], d23 = d[y][2] - d[y][3];
            t[y][0] = s01 + s23; t[y][1] = s01 - s23; t[y][2] = d01 + d23; t[y][3] = d01 - d23;
        }
        for (int x = 0; x < 4; x++) {
            int s01 = t[0][x] + t[1][x], d01 = t[0][x] - t[1][x], s23 = t[2][x] + t[3][x], d23 = t[2][x] - t[3][x];
            sum += abs(s01 + s23) + abs(s01 - s23) + abs(d01 + d23) + abs(d01 - d23);
        }
    }
    return sum >> 1;
}
static int satd_8x8(const uint8_t *a, int sa, const uint8_t *b, int sb)
{
    return satd_8x4(a, sa, b, sb) + satd_8x4(a + 4 * sa, sa, b + 4 * sb, sb);
}
static inline int mbcmp(orc_la *la, const uint8_t *a, int sa, const uint8_t *b, int sb)
{
    if (la->la_satd) { la->n_satd++; return satd_8x8(a, sa, b, sb); }    /* [x264] encoder.c: mbcmp_init */
    la->n_sad++;
    return sad_8x8(a, sa, b, sb);
}
static inline int fpelcmp(orc_la *la, const uint8_t *a, int sa, const uint8_t *b, int sb)
{
    la->n_sad++;
    return sad_8x8(a, sa, b, sb);
}

/* ------------------------------------------------------------------------------------------
 * motion compensation ([x264] common/mc.c: get_ref, pixel_avg, mc_weight)
 * ---------------------------------------------------------------------------------------- */
static const uint8_t hpel_ref0[16] = {0, 1, 1, 1, 0, 1, 1, 1, 2, 3, 3, 3, 0, 1, 1, 1};
static const uint8_t hpel_ref1[16] = {0, 0, 1, 0, 2, 2, 3, 2, 2, 2, 3, 2, 2, 2, 3, 2};

static inline int weight_px(const weight_t *w, int p)
{
    if (w->denom >= 1) return clip_pixel(((p * w->scale + (1 << (w->denom - 1))) >> w->denom) + w->offset);
    return clip_pixel(p * w->scale + w->offset);
}

/* always materialises the 8x8 block into dst (stride 8); same pixels get_ref would expose */
static void get_ref_8x8(uint8_t *dst, uint8_t *const planes[4], int stride, int mvx, int mvy, const weight_t *w)
{
    int qpel_idx = ((mvy & 3) << 2) + (mvx & 3);
    int offset = (mvy >> 2) * stride + (mvx >> 2);
    const uint8_t *s1 = planes[hpel_ref0[qpel_idx]] + offset + ((mvy & 3) == 3) * stride;
    if (qpel_idx & 5) {
        const uint8_t *s2 = planes[hpel_ref1[qpel_idx]] + offset + ((mvx & 3) == 3);
        for (int y = 0; y < 8; y++)
            for (int x = 0; x < 8; x++) dst[y * 8 + x] = (s1[y * stride + x] + s2[y * stride + x] + 1) >> 1;
    } else {
        for (int y = 0; y < 8; y++) memcpy(dst + y * 8, s1 + y * stride, 8);
    }
    if (w && w->on)
        for (int i = 0; i < 64; i++) dst[i] = weight_px(w, dst[i]);
}

/* [x264] encoder/slicetype.c: the lookahead's own distance scale (slicetype_frame_cost, macroblock_tree_propagate)
 * and the weight of the list-0 block in a bidirectional average (weightb: implicit weights of H.264 8.4.2.3.1) */
static inline int la_dist_scale_factor(int p0, int p1, int b) { return (((b - p0) << 8) + ((p1 - p0) >> 1)) / (p1 - p0); }
static inline int la_bipred_weight(int weightb, int dist_scale_factor) { return weightb ? 64 - (dist_scale_factor >> 2) : 32; }

static void pixel_avg_8x8(uint8_t *dst, const uint8_t *a, int sa, const uint8_t *b, int sb, int weight)
{
    if (weight == 32) {
        for (int y = 0; y < 8; y++)
            for (int x = 0; x < 8; x++) dst[y * 8 + x] = (a[y * sa + x] + b[y * sb + x] + 1) >> 1;
    } else {
        for (int y = 0; y < 8; y++)
            for (int x = 0; x < 8; x++)
                dst[y * 8 + x] = clip_pixel((a[y * sa + x] * weight + b[y * sb + x] * (64 - weight) + 32) >> 6);
    }
}

/* ------------------------------------------------------------------------------------------
 * intra prediction ([x264] common/predict.c) on a scratch block with FDEC-like neighbours.
 * nb layout: top[-1..15] and left[0..7]; see intra_cost().
 * ---------------------------------------------------------------------------------------- */
typedef struct { int tl; int top[16]; int left[8]; } nbr_t;

static void pred_8x8c_dc(uint8_t *d, const nbr_t *n)
{
    int s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int i = 0; i < 4; i++) { s0 += n->top[i]; s1 += n->top[i + 4]; s2 += n->left[i]; s3 += n->left[i + 4]; }
    int dc[4] = {(s0 + s2 + 4) >> 3, (s1 + 2) >> 2, (s3 + 2) >> 2, (s1 + s3 + 4) >> 3};
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) d[y * 8 + x] = dc[(y >> 2) * 2 + (x >> 2)];
}
static void pred_8x8c_h(uint8_t *d, const nbr_t *n) { for (int y = 0; y < 8; y++) memset(d + y * 8, n->left[y], 8); }
static void pred_8x8c_v(uint8_t *d, const nbr_t *n) { for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) d[y * 8 + x] = n->top[x]; }
static void pred_8x8c_p(uint8_t *d, const nbr_t *n)
{
    int H = 0, V = 0;
    for (int i = 0; i < 4; i++) {
        int tm = 2 - i < 0 ? n->tl : n->top[2 - i];
        int lm = 2 - i < 0 ? n->tl : n->left[2 - i];
        H += (i + 1) * (n->top[4 + i] - tm);
        V += (i + 1) * (n->left[4 + i] - lm);
    }
    int a = 16 * (n->left[7] + n->top[7]);
    int b = (17 * H + 16) >> 5, c = (17 * V + 16) >> 5;
    int i00 = a - 3 * b - 3 * c + 16;
    for (int y = 0; y < 8; y++, i00 += c) {
        int pix = i00;
        for (int x = 0; x < 8; x++, pix += b) d[y * 8 + x] = clip_pixel(pix >> 5);
    }
}

/* [x264] predict_8x8_filter with all neighbours: ft[-1..15] (ft[-1] = corner), fl[0..7] */
typedef struct { int c; int t[16]; int l[8]; } edge_t;
static void filter_edges(edge_t *e, const nbr_t *n)
{
    e->c = (n->top[0] + 2 * n->tl + n->left[0] + 2) >> 2;
    e->t[0] = (n->tl + 2 * n->top[0] + n->top[1] + 2) >> 2;
    for (int x = 1; x < 15; x++) e->t[x] = (n->top[x - 1] + 2 * n->top[x] + n->top[x + 1] + 2) >> 2;
    e->t[15] = (n->top[14] + 3 * n->top[15] + 2) >> 2;
    e->l[0] = (n->tl + 2 * n->left[0] + n->left[1] + 2) >> 2;
    for (int y = 1; y < 7; y++) e->l[y] = (n->left[y - 1] + 2 * n->left[y] + n->left[y + 1] + 2) >> 2;
    e->l[7] = (n->left[6] + 3 * n->left[7] + 2) >> 2;
}
#define T(i) ((i) < 0 ? e->c : e->t[i])
#define L(i) ((i) < 0 ? e->c : e->l[i])
static void pred_8x8_mode(uint8_t *d, const edge_t *e, int mode)
{
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) {
            int v;
            switch (mode) {
            case 3: /* DDL */
                v = (x == 7 && y == 7) ? (T(14) + 3 * T(15) + 2) >> 2 : (T(x + y) + 2 * T(x + y + 1) + T(x + y + 2) + 2) >> 2;
                break;
            case 4: /* DDR */
                if (x > y) v = (T(x - y - 2) + 2 * T(x - y - 1) + T(x - y) + 2) >> 2;
                else if (x < y) v = (L(y - x - 2) + 2 * L(y - x - 1) + L(y - x) + 2) >> 2;
                else v = (T(0) + 2 * e->c + L(0) + 2) >> 2;
                break;
            case 5: { /* VR */
                int z = 2 * x - y, i = x - (y >> 1);
                if (z >= 0 && !(z & 1)) v = (T(i - 1) + T(i) + 1) >> 1;
                else if (z >= 0) v = (T(i - 2) + 2 * T(i - 1) + T(i) + 2) >> 2;
                else if (z == -1) v = (L(0) + 2 * e->c + T(0) + 2) >> 2;
                else v = (L(y - 2 * x - 1) + 2 * L(y - 2 * x - 2) + L(y - 2 * x - 3) + 2) >> 2;
                break;
            }
            case 6: { /* HD */
                int z = 2 * y - x, i = y - (x >> 1);
                if (z >= 0 && !(z & 1)) v = (L(i - 1) + L(i) + 1) >> 1;
                else if (z >= 0) v = (L(i - 2) + 2 * L(i - 1) + L(i) + 2) >> 2;
                else if (z == -1) v = (L(0) + 2 * e->c + T(0) + 2) >> 2;
                else v = (T(x - 2 * y - 1) + 2 * T(x - 2 * y - 2) + T(x - 2 * y - 3) + 2) >> 2;
                break;
            }
            case 7: { /* VL */
                int i = x + (y >> 1);
                v = !(y & 1) ? (T(i) + T(i + 1) + 1) >> 1 : (T(i) + 2 * T(i + 1) + T(i + 2) + 2) >> 2;
                break;
            }
            default: { /* 8: HU */
                int z = x + 2 * y, i = y + (x >> 1);
                if (z > 13) v = L(7);
                else if (z == 13) v = (L(6) + 3 * L(7) + 2) >> 2;
                else if (!(z & 1)) v = (L(i) + L(i + 1) + 1) >> 1;
                else v = (L(i) + 2 * L(i + 1) + L(i + 2) + 2) >> 2;
                break;
            }
            }
            d[y * 8 + x] = (uint8_t)v;
        }
}
#undef T
#undef L

/* intra part of [x264] encoder/slicetype.c: slicetype_mb_cost (lowres_intra_mb) */
static int intra_cost_mb(orc_la *la, const uint8_t *src, int stride)
{
    nbr_t n;
    uint8_t pred[64];
    n.tl = src[-stride - 1];
    for (int i = 0; i < 16; i++) n.top[i] = src[-stride + i];
    for (int i = 0; i < 8; i++) n.left[i] = src[i * stride - 1];
    int best = COST_MAX, c;
    pred_8x8c_dc(pred, &n); c = mbcmp(la, pred, 8, src, stride); best = MIN(best, c);
    pred_8x8c_h(pred, &n);  c = mbcmp(la, pred, 8, src, stride); best = MIN(best, c);
    pred_8x8c_v(pred, &n);  c = mbcmp(la, pred, 8, src, stride); best = MIN(best, c);
    if (la->p.subme > 1) {
        edge_t e;
        pred_8x8c_p(pred, &n); c = mbcmp(la, src, stride, pred, 8); best = MIN(best, c);
        filter_edges(&e, &n);
        for (int m = 3; m < 9; m++) { pred_8x8_mode(pred, &e, m); c = mbcmp(la, src, stride, pred, 8); best = MIN(best, c); }
    }
    return best + 5 * 1 /* intra_penalty = 5*lambda */ + 4 /* lowres_penalty */;
}

/* ------------------------------------------------------------------------------------------
 * motion search ([x264] encoder/me.c: x264_me_search_ref + refine_subpel, lowres subset)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const uint8_t *fenc; int fenc_stride;
    uint8_t *fref[4];               /* pointers at this MB's pixel offset */
    const uint8_t *fref_w;          /* weighted plane (or fref[0]) at this MB's offset */
    int stride;
    weight_t weight;
    int mvp[2];
    int mv[2];
    int cost;
} me_t;

static const int8_t hex2[8][2] = {{-1, -2}, {-2, 0}, {-1, 2}, {1, 2}, {2, 0}, {1, -2}, {-1, -2}, {-2, 0}};
static const uint8_t mod6m1[8] = {5, 0, 1, 2, 3, 4, 5, 0};
static const int8_t square1[9][2] = {{0, 0}, {0, -1}, {0, 1}, {-1, 0}, {1, 0}, {-1, -1}, {-1, 1}, {1, -1}, {1, 1}};

#define COSTMV(dx, dy) (la->cost_mv[(dx) - m->mvp[0]] + la->cost_mv[(dy) - m->mvp[1]])

static inline int cost_fpel(orc_la *la, me_t *m, int mx, int my)
{
    return fpelcmp(la, m->fenc, m->fenc_stride, m->fref_w + my * m->stride + mx, m->stride) + COSTMV(mx * 4, my * 4);
}
static inline int cost_qpel_sad(orc_la *la, me_t *m, int mx, int my)
{
    uint8_t pix[64];
    get_ref_8x8(pix, m->fref, m->stride, mx, my, &m->weight);
    return fpelcmp(la, m->fenc, m->fenc_stride, pix, 8) + COSTMV(mx, my);
}
static inline int cost_qpel_satd(orc_la *la, me_t *m, int mx, int my)
{
    uint8_t pix[64];
    get_ref_8x8(pix, m->fref, m->stride, mx, my, &m->weight);
    return mbcmp(la, m->fenc, m->fenc_stride, pix, 8) + COSTMV(mx, my);
}
static inline int mv_in_range(orc_la *la, int mx, int my)
{
    return mx >= la->mv_limit_fpel[0][0] && mx <= la->mv_limit_fpel[1][0] &&
           my >= la->mv_limit_fpel[0][1] && my <= la->mv_limit_fpel[1][1];
}

static void refine_subpel(orc_la *la, me_t *m, int hpel_iters, int qpel_iters)
{
    int bmx = m->mv[0], bmy = m->mv[1], bcost = m->cost;
    if (hpel_iters) {
        if (la->la_subpel_refine < 3) {
            int mx = clip3(m->mvp[0], la->mv_min_spel[0] + 2, la->mv_max_spel[0] - 2);
            int my = clip3(m->mvp[1], la->mv_min_spel[1] + 2, la->mv_max_spel[1] - 2);
            if ((mx - bmx) | (my - bmy)) {
                int c = cost_qpel_sad(la, m, mx, my);
                if (c < bcost) { bcost = c; bmx = mx; bmy = my; }
            }
        }
        for (int i = hpel_iters; i > 0; i--) {
            int omx = bmx, omy = bmy, best = 0;
            static const int8_t d[4][2] = {{0, -2}, {0, 2}, {-2, 0}, {2, 0}};
            for (int k = 0; k < 4; k++) {
                int c = cost_qpel_sad(la, m, omx + d[k][0], omy + d[k][1]);
                if (c < bcost) { bcost = c; best = k + 1; }
            }
            if (!best) break;
            bmx = omx + d[best - 1][0]; bmy = omy + d[best - 1][1];
        }
    }
    if (la->la_satd) {                       /* mbcmp != fpelcmp: re-score the best with SATD */
        bcost = cost_qpel_satd(la, m, bmx, bmy);
    }
    {
        int bdir = -1;
        for (int i = qpel_iters; i > 0; i--) {
            if (bmy <= la->mv_min_spel[1] || bmy >= la->mv_max_spel[1] || bmx <= la->mv_min_spel[0] || bmx >= la->mv_max_spel[0])
                break;
            int odir = bdir, omx = bmx, omy = bmy;
            static const int8_t d[4][2] = {{0, -1}, {0, 1}, {-1, 0}, {1, 0}};
            for (int k = 0; k < 4; k++) {
                if ((k ^ 1) == odir) continue;
                int c = cost_qpel_satd(la, m, omx + d[k][0], omy + d[k][1]);
                if (c < bcost) { bcost = c; bmx = omx + d[k][0]; bmy = omy + d[k][1]; bdir = k; }
            }
            if (bmx == omx && bmy == omy) break;
        }
    }
    m->cost = bcost; m->mv[0] = bmx; m->mv[1] = bmy;
}

static void me_search(orc_la *la, me_t *m, int16_t (*mvc)[2], int i_mvc)
{
    int bmx, bmy, bcost = COST_MAX, bpred_cost = COST_MAX;
    int bpred_mx = 0, bpred_my = 0, pmv_nonzero, pm_fx = 0, pm_fy = 0;
    const int mv_x_min = la->mv_limit_fpel[0][0], mv_y_min = la->mv_limit_fpel[0][1];
    const int mv_x_max = la->mv_limit_fpel[1][0], mv_y_max = la->mv_limit_fpel[1][1];
    la->n_search++;

    if (la->la_subpel_refine >= 3) {
        bpred_mx = clip3(m->mvp[0], mv_x_min * 4, mv_x_max * 4);
        bpred_my = clip3(m->mvp[1], mv_y_min * 4, mv_y_max * 4);
        const int pmx = bpred_mx, pmy = bpred_my;
        pmv_nonzero = (pmx | pmy) != 0;
        bpred_cost = cost_qpel_sad(la, m, bpred_mx, bpred_my);
        int pmv_cost = bpred_cost;
        if (i_mvc > 0) {
            /* x264_predictor_clip: drop zero and ==pmv candidates, clip the rest to 4*fpel limits */
            int cand[5][2], n = 0;
            for (int i = 0; i < i_mvc; i++) {
                int mx = mvc[i][0], my = mvc[i][1];
                if ((!mx && !my) || (mx == pmx && my == pmy)) continue;
                cand[n][0] = clip3(mx, mv_x_min * 4, mv_x_max * 4);
                cand[n][1] = clip3(my, mv_y_min * 4, mv_y_max * 4);
                n++;
            }
            if (n > 0) {
                int best = 0;          /* index 0 = pmv itself; strict < keeps the earliest on ties */
                for (int i = 0; i < n; i++) {
                    int c = cost_qpel_sad(la, m, cand[i][0], cand[i][1]);
                    if (((c << 4) + i + 1) < ((bpred_cost << 4) + best)) { bpred_cost = c; best = i + 1; }
                }
                if (best) { bpred_mx = cand[best - 1][0]; bpred_my = cand[best - 1][1]; }
            }
        }
        bmx = (bpred_mx + 2) >> 2; bmy = (bpred_my + 2) >> 2;
        if ((bpred_mx & 3) | (bpred_my & 3)) { bcost = cost_fpel(la, m, bmx, bmy); }
        else bcost = bpred_cost;
        if (pmv_nonzero) {
            if (bmx | bmy) { int c = cost_fpel(la, m, 0, 0); if (c < bcost) { bcost = c; bmx = 0; bmy = 0; } }
        } else {
            if (pmv_cost < bcost) { bcost = pmv_cost; bmx = 0; bmy = 0; }
        }
    } else {
        /* subme < 3: predictors rounded to fullpel; the rounded MVP is scored WITHOUT its mv
         * cost (upstream's deliberate bias), which is added back at the end if it wins. */
        bmx = pm_fx = clip3((m->mvp[0] + 2) >> 2, mv_x_min, mv_x_max);
        bmy = pm_fy = clip3((m->mvp[1] + 2) >> 2, mv_y_min, mv_y_max);
        pmv_nonzero = (pm_fx | pm_fy) != 0;
        bcost = fpelcmp(la, m->fenc, m->fenc_stride, m->fref_w + bmy * m->stride + bmx, m->stride);
        if (i_mvc > 0) {
            /* x264_predictor_roundclip */
            int cand[5][2], n = 0;
            for (int i = 0; i < i_mvc; i++) {
                int mx = (mvc[i][0] + 2) >> 2, my = (mvc[i][1] + 2) >> 2;
                if ((!mx && !my) || (mx == pm_fx && my == pm_fy)) continue;
                cand[n][0] = clip3(mx, mv_x_min, mv_x_max);
                cand[n][1] = clip3(my, mv_y_min, mv_y_max);
                n++;
            }
            int best = 0;
            for (int i = 0; i < n; i++) {
                int c = cost_fpel(la, m, cand[i][0], cand[i][1]);
                if (((c << 4) + i + 1) < ((bcost << 4) + best)) { bcost = c; best = i + 1; }
            }
            if (best) { bmx = cand[best - 1][0]; bmy = cand[best - 1][1]; }
        }
        if (pmv_nonzero) { int c = cost_fpel(la, m, 0, 0); if (c < bcost) { bcost = c; bmx = 0; bmy = 0; } }
    }

    if (la->la_me_method == 0) {
        /* X264_ME_DIA: diamond, radius 1 */
        bcost <<= 4;
        for (int i = la->p.me_range; ; ) {
            int c[4];
            c[0] = cost_fpel(la, m, bmx, bmy - 1); c[1] = cost_fpel(la, m, bmx, bmy + 1);
            c[2] = cost_fpel(la, m, bmx - 1, bmy); c[3] = cost_fpel(la, m, bmx + 1, bmy);
            if (((c[0] << 4) + 1) < bcost) bcost = (c[0] << 4) + 1;
            if (((c[1] << 4) + 3) < bcost) bcost = (c[1] << 4) + 3;
            if (((c[2] << 4) + 4) < bcost) bcost = (c[2] << 4) + 4;
            if (((c[3] << 4) + 12) < bcost) bcost = (c[3] << 4) + 12;
            if (!(bcost & 15)) break;
            bmx -= (int32_t)((uint32_t)bcost << 28) >> 30;
            bmy -= (int32_t)((uint32_t)bcost << 30) >> 30;
            bcost &= ~15;
            if (!(--i && mv_in_range(la, bmx, bmy))) break;
        }
        bcost >>= 4;
    } else {
        /* X264_ME_HEX: hexagon radius 2, then square refine */
        int c[8];
        c[0] = cost_fpel(la, m, bmx - 2, bmy);     c[1] = cost_fpel(la, m, bmx - 1, bmy + 2);
        c[2] = cost_fpel(la, m, bmx + 1, bmy + 2); c[4] = cost_fpel(la, m, bmx + 2, bmy);
        c[5] = cost_fpel(la, m, bmx + 1, bmy - 2); c[6] = cost_fpel(la, m, bmx - 1, bmy - 2);
        bcost <<= 3;
        if (((c[0] << 3) + 2) < bcost) bcost = (c[0] << 3) + 2;
        if (((c[1] << 3) + 3) < bcost) bcost = (c[1] << 3) + 3;
        if (((c[2] << 3) + 4) < bcost) bcost = (c[2] << 3) + 4;
        if (((c[4] << 3) + 5) < bcost) bcost = (c[4] << 3) + 5;
        if (((c[5] << 3) + 6) < bcost) bcost = (c[5] << 3) + 6;
        if (((c[6] << 3) + 7) < bcost) bcost = (c[6] << 3) + 7;
        if (bcost & 7) {
            int dir = (bcost & 7) - 2;
            bmx += hex2[dir + 1][0]; bmy += hex2[dir + 1][1];
            for (int i = (la->p.me_range >> 1) - 1; i > 0 && mv_in_range(la, bmx, bmy); i--) {
                int c0 = cost_fpel(la, m, bmx + hex2[dir + 0][0], bmy + hex2[dir + 0][1]);
                int c1 = cost_fpel(la, m, bmx + hex2[dir + 1][0], bmy + hex2[dir + 1][1]);
                int c2 = cost_fpel(la, m, bmx + hex2[dir + 2][0], bmy + hex2[dir + 2][1]);
                bcost &= ~7;
                if (((c0 << 3) + 1) < bcost) bcost = (c0 << 3) + 1;
                if (((c1 << 3) + 2) < bcost) bcost = (c1 << 3) + 2;
                if (((c2 << 3) + 3) < bcost) bcost = (c2 << 3) + 3;
                if (!(bcost & 7)) break;
                dir += (bcost & 7) - 2;
                dir = mod6m1[dir + 1];
                bmx += hex2[dir + 1][0]; bmy += hex2[dir + 1][1];
            }
        }
        bcost >>= 3;
        bcost <<= 4;
        for (int k = 1; k <= 8; k++) {
            int cc = cost_fpel(la, m, bmx + square1[k][0], bmy + square1[k][1]);
            if (((cc << 4) + k) < bcost) bcost = (cc << 4) + k;
        }
        bmx += square1[bcost & 15][0]; bmy += square1[bcost & 15][1];
        bcost >>= 4;
    }

    if (la->la_subpel_refine < 3) {
        m->cost = bcost;
        if (bmx == pm_fx && bmy == pm_fy) m->cost += COSTMV(bmx * 4, bmy * 4);   /* "compute the real cost" */
        m->mv[0] = bmx * 4; m->mv[1] = bmy * 4;
    } else {
        if (bpred_cost < bcost) { m->mv[0] = bpred_mx; m->mv[1] = bpred_my; m->cost = bpred_cost; }
        else { m->mv[0] = bmx * 4; m->mv[1] = bmy * 4; m->cost = bcost; }
    }
    /* subpel_iterations[subme]: subme 2 -> hpel 1, qpel 0; subme 4 -> hpel 1, qpel 1 */
    if (la->la_subpel_refine >= 4) refine_subpel(la, m, 1, 1);
    else refine_subpel(la, m, 1, 0);
}

/* ------------------------------------------------------------------------------------------
 * [x264] encoder/slicetype.c: slicetype_mb_cost
 * ---------------------------------------------------------------------------------------- */
enum { OUT_COST_EST, OUT_COST_EST_AQ, OUT_INTRA_MBS, OUT_N };

typedef struct {
    frame_t **frames; int p0, p1, b, dist_scale_factor; int do_search[2]; weight_t w;
    int out_inter[OUT_N], out_intra[OUT_N];
    int *row_inter, *row_intra;
} slice_ctx;

static void mb_cost(orc_la *la, slice_ctx *s, int mb_x, int mb_y)
{
    frame_t *fref0 = s->frames[s->p0], *fref1 = s->frames[s->p1], *fenc = s->frames[s->b];
    const int p0 = s->p0, p1 = s->p1, b = s->b;
    const int b_bidir = b < p1;
    const int mb_stride = la->mb_w, mb_xy = mb_x + mb_y * mb_stride;
    const int stride = la->lstride;
    const int pel = 8 * (mb_x + mb_y * stride);
    const int bipred_weight = la_bipred_weight(la->p.weightb, s->dist_scale_factor);
    int16_t (*fenc_mvs[2])[2] = {b != p0 ? &fenc->mvs[0][b - p0 - 1][mb_xy] : NULL, b != p1 ? &fenc->mvs[1][p1 - b - 1][mb_xy] : NULL};
    int *fenc_costs[2] = {b != p0 ? &fenc->mv_costs[0][b - p0 - 1][mb_xy] : NULL, b != p1 ? &fenc->mv_costs[1][p1 - b - 1][mb_xy] : NULL};
    const int b_frame_score_mb = (mb_x > 0 && mb_x < la->mb_w - 1 && mb_y > 0 && mb_y < la->mb_h - 1) || la->mb_w <= 2 || la->mb_h <= 2;
    const uint8_t *fenc_px = fenc->lowres[0] + pel;
    uint8_t pix1[64], pix2[64];
    me_t m[2];
    int bcost = COST_MAX, list_used = 0;
    const int lowres_penalty = 4;
    la->n_mbcost++;

    if (p0 == p1) goto lowres_intra_mb;

    {
        int mv_range = 2 * la->p.mv_range;
        la->mv_min_spel[0] = MAX(4 * (-8 * mb_x - 12), -mv_range);
        la->mv_max_spel[0] = MIN(4 * (8 * (la->mb_w - mb_x - 1) + 12), mv_range - 1);
        la->mv_limit_fpel[0][0] = la->mv_min_spel[0] >> 2;
        la->mv_limit_fpel[1][0] = la->mv_max_spel[0] >> 2;
        if (mb_x >= la->mb_w - 2) {
            la->mv_min_spel[1] = MAX(4 * (-8 * mb_y - 12), -mv_range);
            la->mv_max_spel[1] = MIN(4 * (8 * (la->mb_h - mb_y - 1) + 12), mv_range - 1);
            la->mv_limit_fpel[0][1] = la->mv_min_spel[1] >> 2;
            la->mv_limit_fpel[1][1] = la->mv_max_spel[1] >> 2;
        }
    }
#define CLIP_MV(mv) do { (mv)[0] = clip3((mv)[0], la->mv_min_spel[0], la->mv_max_spel[0]); \
                         (mv)[1] = clip3((mv)[1], la->mv_min_spel[1], la->mv_max_spel[1]); } while (0)
#define TRY_BIDIR(mv0, mv1, penalty) do { \
        int i_cost; \
        if (la->p.subme <= 1) { \
            int h1 = (((mv0)[0] & 2) >> 1) + ((mv0)[1] & 2), h2 = (((mv1)[0] & 2) >> 1) + ((mv1)[1] & 2); \
            const uint8_t *s1 = m[0].fref[h1] + ((mv0)[0] >> 2) + ((mv0)[1] >> 2) * stride; \
            const uint8_t *s2 = m[1].fref[h2] + ((mv1)[0] >> 2) + ((mv1)[1] >> 2) * stride; \
            pixel_avg_8x8(pix1, s1, stride, s2, stride, bipred_weight); \
        } else { \
            uint8_t r1[64]; \
            get_ref_8x8(r1, m[0].fref, stride, (mv0)[0], (mv0)[1], &s->w); \
            get_ref_8x8(pix2, m[1].fref, stride, (mv1)[0], (mv1)[1], &s->w); \
            pixel_avg_8x8(pix1, r1, 8, pix2, 8, bipred_weight); \
        } \
        i_cost = (penalty) * 1 + mbcmp(la, fenc_px, stride, pix1, 8); \
        if (i_cost < bcost) { bcost = i_cost; list_used = 3; } \
    } while (0)

    memset(m, 0, sizeof(m));
    m[0].fenc = fenc_px; m[0].fenc_stride = stride; m[0].stride = stride;
    m[0].weight = s->w;
    for (int k = 0; k < 4; k++) m[0].fref[k] = fref0->lowres[k] + pel;
    m[0].fref_w = m[0].fref[0];
    if (s->w.on) m[0].fref_w = la->weight_buf + la->lorigin + pel;

    if (b_bidir) {
        int dmv[2][2];
        m[1].fenc = fenc_px; m[1].fenc_stride = stride; m[1].stride = stride;
        m[1].weight.on = 0;
        for (int k = 0; k < 4; k++) m[1].fref[k] = fref1->lowres[k] + pel;
        m[1].fref_w = m[1].fref[0];
        if (fref1->mvs_searched[0][p1 - p0 - 1]) {
            int16_t *mvr = fref1->mvs[0][p1 - p0 - 1][mb_xy];
            dmv[0][0] = (mvr[0] * s->dist_scale_factor + 128) >> 8;
            dmv[0][1] = (mvr[1] * s->dist_scale_factor + 128) >> 8;
            dmv[1][0] = dmv[0][0] - mvr[0];
            dmv[1][1] = dmv[0][1] - mvr[1];
            CLIP_MV(dmv[0]);
            CLIP_MV(dmv[1]);
            if (la->p.subme <= 1) { dmv[0][0] &= ~1; dmv[0][1] &= ~1; dmv[1][0] &= ~1; dmv[1][1] &= ~1; }
        } else
            dmv[0][0] = dmv[0][1] = dmv[1][0] = dmv[1][1] = 0;
        TRY_BIDIR(dmv[0], dmv[1], 0);
        if (dmv[0][0] | dmv[0][1] | dmv[1][0] | dmv[1][1]) {
            pixel_avg_8x8(pix1, m[0].fref[0], stride, m[1].fref[0], stride, bipred_weight);
            int i_cost = mbcmp(la, fenc_px, stride, pix1, 8);
            if (i_cost < bcost) { bcost = i_cost; list_used = 3; }
        }
    }

    for (int l = 0; l < 1 + b_bidir; l++) {
        if (s->do_search[l]) {
            int i_mvc = 0;
            int16_t (*fenc_mv)[2] = fenc_mvs[l];
            int16_t mvc[4][2];
            memset(mvc, 0, sizeof(mvc));
#define MVC(mv) do { mvc[i_mvc][0] = (mv)[0]; mvc[i_mvc][1] = (mv)[1]; i_mvc++; } while (0)
            if (mb_x < la->mb_w - 1) MVC(fenc_mv[1]);
            if (mb_y < la->slice_end - 1) {
                MVC(fenc_mv[mb_stride]);
                if (mb_x > 0) MVC(fenc_mv[mb_stride - 1]);
                if (mb_x < la->mb_w - 1) MVC(fenc_mv[mb_stride + 1]);
            }
#undef MVC
            if (i_mvc <= 1) { m[l].mvp[0] = mvc[0][0]; m[l].mvp[1] = mvc[0][1]; }
            else { m[l].mvp[0] = median3(mvc[0][0], mvc[1][0], mvc[2][0]); m[l].mvp[1] = median3(mvc[0][1], mvc[1][1], mvc[2][1]); }

            int skip = 0;
            if (!(m[l].mvp[0] | m[l].mvp[1])) {
                m[l].cost = mbcmp(la, fenc_px, stride, m[l].fref[0], stride);
                if (m[l].cost < 64) { m[l].mv[0] = m[l].mv[1] = 0; skip = 1; }
            }
            if (!skip) {
                me_search(la, &m[l], mvc, i_mvc);
                m[l].cost -= la->cost_mv[0];
                if (m[l].mv[0] | m[l].mv[1]) m[l].cost += 5 * 1;
            }
            (*fenc_mvs[l])[0] = (int16_t)m[l].mv[0]; (*fenc_mvs[l])[1] = (int16_t)m[l].mv[1];
            *fenc_costs[l] = m[l].cost;
        } else {
            m[l].mv[0] = (*fenc_mvs[l])[0]; m[l].mv[1] = (*fenc_mvs[l])[1];
            m[l].cost = *fenc_costs[l];
        }
        if (m[l].cost < bcost) { bcost = m[l].cost; list_used = l + 1; }
    }

    if (b_bidir && (m[0].mv[0] | m[0].mv[1] | m[1].mv[0] | m[1].mv[1]))
        TRY_BIDIR(m[0].mv, m[1].mv, 5);

lowres_intra_mb:
    if (!fenc->b_intra_calculated) {
        int icost = intra_cost_mb(la, fenc_px, stride);
        fenc->intra_cost[mb_xy] = (uint16_t)icost;
        int icost_aq = icost;
        if (la->p.aq_mode) icost_aq = (icost_aq * fenc->inv_qscale[mb_xy] + 128) >> 8;
        s->row_intra[mb_y] += icost_aq;
        if (b_frame_score_mb) { s->out_intra[OUT_COST_EST] += icost; s->out_intra[OUT_COST_EST_AQ] += icost_aq; }
    }
    bcost += lowres_penalty;

    if (!b_bidir) {
        int icost = fenc->intra_cost[mb_xy];
        int b_intra = icost < bcost;
        if (b_intra) { bcost = icost; list_used = 0; }
        if (b_frame_score_mb) s->out_inter[OUT_INTRA_MBS] += b_intra;
    }
    if (p0 != p1) {
        int bcost_aq = bcost;
        if (la->p.aq_mode) bcost_aq = (bcost_aq * fenc->inv_qscale[mb_xy] + 128) >> 8;
        s->row_inter[mb_y] += bcost_aq;
        if (b_frame_score_mb) { s->out_inter[OUT_COST_EST] += bcost; s->out_inter[OUT_COST_EST_AQ] += bcost_aq; }
    }
    fenc->lowres_costs[b - p0][p1 - b][mb_xy] = (uint16_t)(MIN(bcost, LOWRES_COST_MASK) + (list_used << LOWRES_COST_SHIFT));
}

/* ------------------------------------------------------------------------------------------
 * weights ([x264] encoder/slicetype.c: x264_weights_analyse with b_lookahead=1)
 * ---------------------------------------------------------------------------------------- */
static int frame_cost(orc_la *la, frame_t **frames, int p0, int p1, int b);

static int ue_size(unsigned v) { v += 1; int n = 0; while (v >> (n + 1)) n++; return 2 * n + 1; }
static int se_size(int v) { int t = 1 - v * 2; if (t < 0) t = v * 2; int n = 0; while (t >> (n + 1)) n++; return 2 * n + 1; }

static unsigned weight_cost_luma(orc_la *la, frame_t *fenc, const uint8_t *src, const weight_t *w)
{
    unsigned cost = 0;
    const int stride = la->lstride;
    int i_mb = 0;
    for (int y = 0; y < la->lh; y += 8)
        for (int x = 0; x < la->lw; x += 8, i_mb++) {
            const uint8_t *s = src + y * stride + x, *f = fenc->lowres[0] + y * stride + x;
            int cmp;
            if (w) {
                uint8_t buf[64];
                for (int yy = 0; yy < 8; yy++) for (int xx = 0; xx < 8; xx++) buf[yy * 8 + xx] = weight_px(w, s[yy * stride + xx]);
                cmp = mbcmp(la, buf, 8, f, stride);
            } else
                cmp = mbcmp(la, s, stride, f, stride);
            cost += MIN(cmp, fenc->intra_cost[i_mb]);
        }
    if (w) {
        /* weight_slice_header_cost: lambda(=1) * numslices(=1) * (10 + denom + 2*(scale+offset bits)) */
        int denom_cost = ue_size(w->denom) * 2;
        cost += 1 * 1 * (10 + denom_cost + 2 * (se_size(w->scale) + se_size(w->offset)));
    }
    return cost;
}

static void weights_analyse(orc_la *la, frame_t *fenc, frame_t *ref)
{
    const float epsilon = 1.f / 128.f;
    weight_t *wt = &fenc->weight;
    wt->on = 0; wt->scale = 1; wt->denom = 0; wt->offset = 0;
    int zero_bias = !ref->pixel_ssd[0];
    float fenc_var = fenc->pixel_ssd[0] + zero_bias;
    float ref_var = ref->pixel_ssd[0] + zero_bias;
    float guess_scale = sqrtf(fenc_var / ref_var);
    float fenc_mean = (float)(fenc->pixel_sum[0] + zero_bias) / (la->luma_h * la->luma_w) / 1;
    float ref_mean = (float)(ref->pixel_sum[0] + zero_bias) / (la->luma_h * la->luma_w) / 1;

    if (fabsf(ref_mean - fenc_mean) < 0.5f && fabsf(1.f - guess_scale) < epsilon) return;

    /* x264_weight_get_h264( round(guess_scale*128), 0 ) */
    int scale = (int)round(guess_scale * 128), denom = 7;
    while (denom > 0 && scale > 127) { denom--; scale >>= 1; }
    scale = MIN(scale, 127);
    int found = 0, mindenom = denom, minscale = scale, minoff = 0;

    if (!fenc->b_intra_calculated) { frame_t *one[1] = {fenc}; frame_cost(la, one, 0, 0, 0); }
    const uint8_t *mcbuf = ref->lowres[0];       /* weight_cost_init_luma: no MVs yet in the lookahead */
    unsigned origscore, minscore;
    origscore = minscore = weight_cost_luma(la, fenc, mcbuf, NULL);
    if (!minscore) return;

    {
        int cur_scale = minscale;
        int cur_offset = fenc_mean - ref_mean * cur_scale / (1 << mindenom) + 0.5f * 1;
        if (cur_offset < -128 || cur_offset > 127) {
            cur_offset = clip3(cur_offset, -128, 127);
            cur_scale = clip3f((1 << mindenom) * (fenc_mean - cur_offset) / ref_mean + 0.5f, 0, 127);
        }
        int i_off = clip3(cur_offset, -128, 127);
        weight_t w = {1, cur_scale, mindenom, i_off};
        unsigned sc = weight_cost_luma(la, fenc, mcbuf, &w);
        if (sc < minscore) { minscore = sc; minscale = cur_scale; minoff = i_off; found = 1; }
    }
    while (mindenom > 0 && !(minscale & 1)) { mindenom--; minscale >>= 1; }
    if (!found || (minscale == 1 << mindenom && minoff == 0) || (float)minscore / origscore > 0.998f) return;
    wt->on = 1; wt->scale = minscale; wt->denom = mindenom; wt->offset = minoff;
    /* X264_WEIGHTP_FAKE: the weight is only used inside the lookahead; what it gained is kept for
     * macroblock_tree_finish ([x264] x264_weights_analyse: f_weighted_cost_delta[i_delta_index]) */
    if (la->p.weightp == ORC_WEIGHTP_FAKE)
        fenc->weighted_cost_delta[fenc->i_frame - ref->i_frame - 1] = (float)minscore / origscore;

    /* x264_weight_scale_plane over the whole padded lowres[0] of the reference */
    const uint8_t *src = ref->lowres_buf;
    for (int i = 0; i < la->lplane; i++) la->weight_buf[i] = weight_px(wt, src[i]);
}

/* ------------------------------------------------------------------------------------------
 * [x264] encoder/slicetype.c: slicetype_slice_cost + slicetype_frame_cost
 * ---------------------------------------------------------------------------------------- */
static int frame_cost(orc_la *la, frame_t **frames, int p0, int p1, int b)
{
    frame_t *fenc = frames[b];
    int score;
    if (fenc->cost_est[b - p0][p1 - b] >= 0) return fenc->cost_est[b - p0][p1 - b];

    slice_ctx s;
    memset(&s, 0, sizeof(s));
    s.frames = frames; s.p0 = p0; s.p1 = p1; s.b = b;
    s.dist_scale_factor = 128;
    s.do_search[0] = b != p0 && !fenc->mvs_searched[0][b - p0 - 1];
    s.do_search[1] = b != p1 && !fenc->mvs_searched[1][p1 - b - 1];
    if (s.do_search[0]) {
        if (la->p.weightp && b == p1) { weights_analyse(la, fenc, frames[p0]); s.w = fenc->weight; }
        fenc->mvs_searched[0][b - p0 - 1] = 1;
    }
    if (s.do_search[1]) fenc->mvs_searched[1][p1 - b - 1] = 1;
    if (p1 != p0) s.dist_scale_factor = la_dist_scale_factor(p0, p1, b);

    int *row_inter = calloc(la->mb_h, sizeof(int)), *row_intra = calloc(la->mb_h, sizeof(int));
    s.row_inter = row_inter; s.row_intra = row_intra;
    const int T = MAX(1, la->p.lookahead_threads);
    const int do_edges = la->p.b_mbtree || la->mb_w <= 2 || la->mb_h <= 2;
    for (int i = 0; i < T; i++) {
        la->slice_start = (la->mb_h * i + T / 2) / T;
        la->slice_end = (la->mb_h * (i + 1) + T / 2) / T;
        int start_y = MIN(la->slice_end - 1, la->mb_h - 2 + do_edges);
        int end_y = MAX(la->slice_start, 1 - do_edges);
        int start_x = la->mb_w - 2 + do_edges, end_x = 1 - do_edges;
        for (int y = start_y; y >= end_y; y--)
            for (int x = start_x; x >= end_x; x--) mb_cost(la, &s, x, y);
    }

    if (b == p1) fenc->intra_mbs[b - p0] = s.out_inter[OUT_INTRA_MBS];
    if (!fenc->b_intra_calculated) {
        fenc->cost_est[0][0] = s.out_intra[OUT_COST_EST];
        fenc->cost_est_aq[0][0] = s.out_intra[OUT_COST_EST_AQ];
        memcpy(fenc->row_satds[0][0], row_intra, la->mb_h * sizeof(int));
    }
    if (p0 != p1) {
        /* when p0==p1 the two accumulators alias [0][0] in upstream: inter sums are zero and
         * the intra sums above are kept only if intra was just calculated */
        fenc->cost_est[b - p0][p1 - b] = s.out_inter[OUT_COST_EST];
        fenc->cost_est_aq[b - p0][p1 - b] = s.out_inter[OUT_COST_EST_AQ];
        memcpy(fenc->row_satds[b - p0][p1 - b], row_inter, la->mb_h * sizeof(int));
    } else if (fenc->b_intra_calculated) {
        fenc->cost_est[0][0] = 0; fenc->cost_est_aq[0][0] = 0;   /* unreachable via the memo check */
    }
    free(row_inter); free(row_intra);

    score = fenc->cost_est[b - p0][p1 - b];
    if (b != p1) score = (int)((uint64_t)score * 100 / (120 + la->p.b_bias));
    else fenc->b_intra_calculated = 1;
    fenc->cost_est[b - p0][p1 - b] = score;
    return score;
}

/* ------------------------------------------------------------------------------------------
 * [x264] encoder/ratecontrol.c: x264_adaptive_quant_frame (aq-mode 0/1), ac_energy_mb.
 * Works on the tight planes with clamped addressing == the mod-16 replicated frame
 * ([x264] x264_frame_expand_border_mod16).
 * ---------------------------------------------------------------------------------------- */
/* x^(1/8).  [x264] x264_adaptive_quant_frame calls powf(x, 0.125f), whose last bit depends on the C
 * library; three correctly rounded square roots are the same function up to that bit and are
 * reproducible everywhere (CPU checker and GPU agree exactly). */
static inline float pow_1_8(float x) { return sqrtf(sqrtf(sqrtf(x))); }

static uint32_t block_var(const uint8_t *p, int stride, int pw, int ph, int x0, int y0, int bw, int bh, int shift,
                          uint64_t *fsum, uint64_t *fssd)
{
    uint32_t sum = 0, ssd = 0;
    for (int y = 0; y < bh; y++) {
        const uint8_t *r = p + (size_t)MIN(y0 + y, ph - 1) * stride;
        for (int x = 0; x < bw; x++) { uint32_t v = r[MIN(x0 + x, pw - 1)]; sum += v; ssd += v * v; }
    }
    *fsum += sum; *fssd += ssd;
    return ssd - (uint32_t)(((uint64_t)sum * sum) >> shift);
}

static void adaptive_quant_frame(orc_la *la, frame_t *f, const uint8_t *y, int ys, const uint8_t *u, const uint8_t *v, int cs)
{
    const orc_la_params *p = &la->p;
    const int w = p->width, h = p->height;
    const int cf = p->chroma_format;
    const int cw = cf == 3 ? w : w / 2, ch = cf == 1 ? h / 2 : h;
    const int cbw = cf == 3 ? 16 : 8, cbh = cf == 1 ? 8 : 16;
    const int cshift = cf == 3 ? 8 : cf == 2 ? 7 : 6;
    for (int i = 0; i < 3; i++) f->pixel_sum[i] = f->pixel_ssd[i] = 0;
    const int aq_on = p->aq_mode != 0 && p->aq_strength != 0;
    if (!aq_on) {
        for (int i = 0; i < la->mb_count; i++) { f->qp_offset[i] = f->qp_offset_aq[i] = 0; f->inv_qscale[i] = 256; }
        if (!p->weightp) return;
    }
    const int autovar = aq_on && (p->aq_mode == 2 || p->aq_mode == 3);   /* X264_AQ_AUTOVARIANCE, _BIASED */
    float strength = p->aq_strength * 1.0397f;
    float avg_adj = 0.f, avg_adj_pow2 = 0.f, bias_strength = 0.f;
    for (int my = 0; my < la->mb_h; my++)
        for (int mx = 0; mx < la->mb_w; mx++) {
            uint32_t energy = block_var(y, ys, w, h, 16 * mx, 16 * my, 16, 16, 8, &f->pixel_sum[0], &f->pixel_ssd[0]);
            if (u && v) {
                energy += block_var(u, cs, cw, ch, cbw * mx, cbh * my, cbw, cbh, cshift, &f->pixel_sum[1], &f->pixel_ssd[1]);
                energy += block_var(v, cs, cw, ch, cbw * mx, cbh * my, cbw, cbh, cshift, &f->pixel_sum[2], &f->pixel_ssd[2]);
            }
            int xy = mx + my * la->mb_w;
            if (autovar) {
                /* first loop of the auto-variance modes: qp_adj = (energy + 1)^(1/8), summed in MB order */
                float qp_adj = pow_1_8((float)energy * 1.f + 1);
                f->qp_offset[xy] = qp_adj;
                avg_adj += qp_adj;
                avg_adj_pow2 += qp_adj * qp_adj;
            } else if (aq_on) {
                float qp_adj = strength * (x264_log2(MAX(energy, 1)) - (14.427f + 2 * 0));
                f->qp_offset[xy] = f->qp_offset_aq[xy] = qp_adj;
                f->inv_qscale[xy] = (uint16_t)x264_exp2fix8(qp_adj);
            }
        }
    if (autovar) {
        avg_adj /= la->mb_count;
        avg_adj_pow2 /= la->mb_count;
        strength = p->aq_strength * avg_adj;
        avg_adj = avg_adj - 0.5f * (avg_adj_pow2 - 14.f) / avg_adj;
        bias_strength = p->aq_strength;
        for (int xy = 0; xy < la->mb_count; xy++) {
            float qp_adj = f->qp_offset[xy];
            if (p->aq_mode == 3) qp_adj = strength * (qp_adj - avg_adj) + bias_strength * (1.f - 14.f / (qp_adj * qp_adj));
            else qp_adj = strength * (qp_adj - avg_adj);
            f->qp_offset[xy] = f->qp_offset_aq[xy] = qp_adj;
            f->inv_qscale[xy] = (uint16_t)x264_exp2fix8(qp_adj);
        }
    }
    for (int i = 0; i < 3; i++) {
        uint64_t ssd = f->pixel_ssd[i], sum = f->pixel_sum[i];
        int pw = 16 * la->mb_w >> (i && cf != 3), ph = 16 * la->mb_h >> (i && cf == 1);
        f->pixel_ssd[i] = ssd - (sum * sum + (uint64_t)pw * ph / 2) / ((uint64_t)pw * ph);
    }
}

/* ------------------------------------------------------------------------------------------
 * frames
 * ---------------------------------------------------------------------------------------- */
static frame_t *frame_new(orc_la *la)
{
    frame_t *f = calloc(1, sizeof(*f));
    const int n = la->mb_count, B = la->p.bframes;
    f->lowres_buf = malloc((size_t)4 * la->lplane);
    for (int k = 0; k < 4; k++) f->lowres[k] = f->lowres_buf + (size_t)k * la->lplane + la->lorigin;
    f->intra_cost = calloc(n, 2); f->inv_qscale = calloc(n, 2); f->propagate_cost = calloc(n, 2);
    f->qp_offset = calloc(n, 4); f->qp_offset_aq = calloc(n, 4);
    for (int l = 0; l < 2; l++)
        for (int d = 0; d <= B; d++) { f->mvs[l][d] = calloc(n, 4); f->mv_costs[l][d] = calloc(n, 4); }
    for (int a = 0; a < B + 2; a++)
        for (int b = 0; b < B + 2; b++) {
            f->lowres_costs[a][b] = calloc(n, 2);
            f->row_satds[a][b] = calloc(la->mb_h, 4);
            f->cost_est[a][b] = f->cost_est_aq[a][b] = -1;
        }
    f->b_scenecut = 1;
    f->i_type = f->i_forced_type = ORC_TYPE_AUTO;
    f->weight.scale = 1;
    f->rc_d0 = f->rc_d1 = -1;
    return f;
}
static void frame_free(orc_la *la, frame_t *f)
{
    const int B = la->p.bframes;
    free(f->lowres_buf); free(f->intra_cost); free(f->inv_qscale); free(f->propagate_cost);
    free(f->qp_offset); free(f->qp_offset_aq);
    for (int l = 0; l < 2; l++) for (int d = 0; d <= B; d++) { free(f->mvs[l][d]); free(f->mv_costs[l][d]); }
    for (int a = 0; a < B + 2; a++) for (int b = 0; b < B + 2; b++) { free(f->lowres_costs[a][b]); free(f->row_satds[a][b]); }
    free(f);
}

/* ------------------------------------------------------------------------------------------
 * mb-tree ([x264] encoder/slicetype.c: macroblock_tree*, common/mc.c: mbtree_propagate_*)
 * ---------------------------------------------------------------------------------------- */
static inline float clip_duration(float f) { return clip3f(f, 0.01f, 1.00f); }

static void mbtree_finish(orc_la *la, frame_t *frame, float average_duration, int ref0_distance)
{
    int fps_factor = (int)round(clip_duration(average_duration) / clip_duration(frame->f_duration) * 256 / MBTREE_PRECISION);
    float weightdelta = 0.0;
    if (ref0_distance && frame->weighted_cost_delta[ref0_distance - 1] > 0)
        weightdelta = (1.0 - frame->weighted_cost_delta[ref0_distance - 1]);
    float strength = 5.0f * (1.0f - la->p.qcompress);
    for (int i = 0; i < la->mb_count; i++) {
        int intra_cost = (frame->intra_cost[i] * frame->inv_qscale[i] + 128) >> 8;
        if (intra_cost) {
            int propagate_cost = (frame->propagate_cost[i] * fps_factor + 128) >> 8;
            float log2_ratio = x264_log2(intra_cost + propagate_cost) - x264_log2(intra_cost) + weightdelta;
            frame->qp_offset[i] = frame->qp_offset_aq[i] - strength * log2_ratio;
        }
    }
}

#define CLIP_ADD(s, x) do { int t_ = (s) + (x); (s) = (uint16_t)MIN(t_, (1 << 15) - 1); } while (0)

static void mbtree_propagate(orc_la *la, frame_t **frames, float average_duration, int p0, int p1, int b, int referenced)
{
    uint16_t *ref_costs[2] = {frames[p0]->propagate_cost, frames[p1]->propagate_cost};
    int dist_scale_factor = la_dist_scale_factor(p0, p1, b);
    int bipred_weight = la_bipred_weight(la->p.weightb, dist_scale_factor);
    int16_t (*mvs[2])[2] = {b != p0 ? frames[b]->mvs[0][b - p0 - 1] : NULL, b != p1 ? frames[b]->mvs[1][p1 - b - 1] : NULL};
    int bipred_weights[2] = {bipred_weight, 64 - bipred_weight};
    int16_t *buf = la->scratch_amount;
    uint16_t *propagate_cost = frames[b]->propagate_cost;
    uint16_t *lowres_costs = frames[b]->lowres_costs[b - p0][p1 - b];
    const unsigned width = la->mb_w, height = la->mb_h, stride = la->mb_w;
    float fps_factor = clip_duration(frames[b]->f_duration) / (clip_duration(average_duration) * 256.0f) * MBTREE_PRECISION;

    if (!referenced) memset(frames[b]->propagate_cost, 0, la->mb_w * sizeof(uint16_t));

    for (int mb_y = 0; mb_y < la->mb_h; mb_y++) {
        int mb_index = mb_y * la->mb_w;
        /* mbtree_propagate_cost */
        for (int i = 0; i < la->mb_w; i++) {
            int intra_cost = frames[b]->intra_cost[mb_index + i];
            int inter_cost = MIN(intra_cost, lowres_costs[mb_index + i] & LOWRES_COST_MASK);
            float propagate_intra = intra_cost * frames[b]->inv_qscale[mb_index + i];
            float propagate_amount = propagate_cost[i] + propagate_intra * fps_factor;
            float propagate_num = intra_cost - inter_cost;
            float propagate_denom = intra_cost;
            buf[i] = (int16_t)MIN((int)(propagate_amount * propagate_num / propagate_denom + 0.5f), 32767);
        }
        if (referenced) propagate_cost += la->mb_w;
        /* mbtree_propagate_list, list 0 then list 1 */
        for (int list = 0; list < (b != p1 ? 2 : 1); list++) {
            for (int i = 0; i < la->mb_w; i++) {
                int lists_used = lowres_costs[mb_index + i] >> LOWRES_COST_SHIFT;
                if (!(lists_used & (1 << list))) continue;
                int listamount = buf[i];
                if (lists_used == 3) listamount = (listamount * bipred_weights[list] + 32) >> 6;
                int16_t *mv = mvs[list][mb_index + i];
                if (!(mv[0] | mv[1])) { CLIP_ADD(ref_costs[list][mb_y * stride + i], listamount); continue; }
                int x = mv[0], y = mv[1];
                unsigned mbx = (x >> 5) + i, mby = (y >> 5) + mb_y;
                unsigned idx0 = mbx + mby * stride, idx2 = idx0 + stride;
                x &= 31; y &= 31;
                int w0 = (32 - y) * (32 - x), w1 = (32 - y) * x, w2 = y * (32 - x), w3 = y * x;
                w0 = (w0 * listamount + 512) >> 10; w1 = (w1 * listamount + 512) >> 10;
                w2 = (w2 * listamount + 512) >> 10; w3 = (w3 * listamount + 512) >> 10;
                if (mbx < width - 1 && mby < height - 1) {
                    CLIP_ADD(ref_costs[list][idx0 + 0], w0); CLIP_ADD(ref_costs[list][idx0 + 1], w1);
                    CLIP_ADD(ref_costs[list][idx2 + 0], w2); CLIP_ADD(ref_costs[list][idx2 + 1], w3);
                } else {
                    if (mby < height) {
                        if (mbx < width) CLIP_ADD(ref_costs[list][idx0 + 0], w0);
                        if (mbx + 1 < width) CLIP_ADD(ref_costs[list][idx0 + 1], w1);
                    }
                    if (mby + 1 < height) {
                        if (mbx < width) CLIP_ADD(ref_costs[list][idx2 + 0], w2);
                        if (mbx + 1 < width) CLIP_ADD(ref_costs[list][idx2 + 1], w3);
                    }
                }
            }
        }
    }
}

static void macroblock_tree(orc_la *la, frame_t **frames, int num_frames, int b_intra)
{
    int idx = !b_intra;
    int last_nonb, cur_nonb = 1, bframes = 0;
    float total_duration = 0.0;
    for (int j = 0; j <= num_frames; j++) total_duration += frames[j]->f_duration;
    float average_duration = total_duration / (num_frames + 1);
    int i = num_frames;

    if (b_intra) frame_cost(la, frames, 0, 0, 0);
    while (i > 0 && IS_B(frames[i]->i_type)) i--;
    last_nonb = i;
    /* rc.i_lookahead == 0 (lookaheadless mb-tree) is not restated */
    if (last_nonb < idx) return;
    memset(frames[last_nonb]->propagate_cost, 0, la->mb_count * sizeof(uint16_t));

    while (i-- > idx) {
        cur_nonb = i;
        while (IS_B(frames[cur_nonb]->i_type) && cur_nonb > 0) cur_nonb--;
        if (cur_nonb < idx) break;
        frame_cost(la, frames, cur_nonb, last_nonb, last_nonb);
        memset(frames[cur_nonb]->propagate_cost, 0, la->mb_count * sizeof(uint16_t));
        bframes = last_nonb - cur_nonb - 1;
        if (la->p.b_pyramid && bframes > 1) {
            int middle = (bframes + 1) / 2 + cur_nonb;
            frame_cost(la, frames, cur_nonb, last_nonb, middle);
            memset(frames[middle]->propagate_cost, 0, la->mb_count * sizeof(uint16_t));
            while (i > cur_nonb) {
                int p0 = i > middle ? middle : cur_nonb;
                int p1 = i < middle ? middle : last_nonb;
                if (i != middle) {
                    frame_cost(la, frames, p0, p1, i);
                    mbtree_propagate(la, frames, average_duration, p0, p1, i, 0);
                }
                i--;
            }
            mbtree_propagate(la, frames, average_duration, cur_nonb, last_nonb, middle, 1);
        } else {
            while (i > cur_nonb) {
                frame_cost(la, frames, cur_nonb, last_nonb, i);
                mbtree_propagate(la, frames, average_duration, cur_nonb, last_nonb, i, 0);
                i--;
            }
        }
        mbtree_propagate(la, frames, average_duration, cur_nonb, last_nonb, last_nonb, 1);
        last_nonb = cur_nonb;
    }
    mbtree_finish(la, frames[last_nonb], average_duration, last_nonb);
    if (la->p.b_pyramid && bframes > 1)
        mbtree_finish(la, frames[last_nonb + (bframes + 1) / 2], average_duration, 0);
}

/* ------------------------------------------------------------------------------------------
 * frame-type decision ([x264] encoder/slicetype.c)
 * ---------------------------------------------------------------------------------------- */
static uint64_t path_cost(orc_la *la, frame_t **frames, const char *path, uint64_t threshold)
{
    uint64_t cost = 0;
    int loc = 1, cur_nonb = 0;
    path--;     /* the 1st path element is really the 2nd frame */
    while (path[loc]) {
        int next_nonb = loc;
        while (path[next_nonb] == 'B') next_nonb++;
        if (path[next_nonb] == 'P') cost += frame_cost(la, frames, cur_nonb, next_nonb, next_nonb);
        else cost += frame_cost(la, frames, next_nonb, next_nonb, next_nonb);
        if (cost > threshold) break;
        if (la->p.b_pyramid && next_nonb - cur_nonb > 2) {
            int middle = cur_nonb + (next_nonb - cur_nonb) / 2;
            cost += frame_cost(la, frames, cur_nonb, next_nonb, middle);
            for (int next_b = loc; next_b < middle && cost < threshold; next_b++)
                cost += frame_cost(la, frames, cur_nonb, middle, next_b);
            for (int next_b = middle + 1; next_b < next_nonb && cost < threshold; next_b++)
                cost += frame_cost(la, frames, middle, next_nonb, next_b);
        } else
            for (int next_b = loc; next_b < next_nonb && cost < threshold; next_b++)
                cost += frame_cost(la, frames, cur_nonb, next_nonb, next_b);
        loc = next_nonb + 1;
        cur_nonb = next_nonb;
    }
    return cost;
}

static void slicetype_path(orc_la *la, frame_t **frames, int length, char (*best_paths)[ORC_LOOKAHEAD_MAX + 1])
{
    char paths[2][ORC_LOOKAHEAD_MAX + 1];
    int num_paths = MIN(la->p.bframes + 1, length);
    uint64_t best_cost = COST_MAX64;
    int best_possible = 0, idx = 0;
    for (int path = 0; path < num_paths; path++) {
        int len = length - (path + 1);
        memcpy(paths[idx], best_paths[len % (ORC_BFRAME_MAX + 1)], len);
        memset(paths[idx] + len, 'B', path);
        strcpy(paths[idx] + len + path, "P");
        int possible = 1;
        for (int i = 1; i <= length; i++) {
            int t = frames[i]->i_type;
            if (t == ORC_TYPE_AUTO) continue;
            if (IS_B(t)) possible = possible && (i < len || i == length || paths[idx][i - 1] == 'B');
            else {
                possible = possible && (i < len || paths[idx][i - 1] != 'B');
                paths[idx][i - 1] = IS_I(t) ? 'I' : 'P';
            }
        }
        if (possible || !best_possible) {
            if (possible && !best_possible) best_cost = COST_MAX64;
            uint64_t cost = path_cost(la, frames, paths[idx], best_cost);
            if (cost < best_cost) { best_cost = cost; best_possible = possible; idx ^= 1; }
        }
    }
    memcpy(best_paths[length % (ORC_BFRAME_MAX + 1)], paths[idx ^ 1], length);
}

static int scenecut_internal(orc_la *la, frame_t **frames, int p0, int p1)
{
    frame_t *frame = frames[p1];
    frame_cost(la, frames, p0, p1, p1);
    int icost = frame->cost_est[0][0];
    int pcost = frame->cost_est[p1 - p0][0];
    float f_bias;
    int i_gop_size = frame->i_frame - la->i_last_keyframe;
    float f_thresh_max = la->p.scenecut / 100.0;
    float f_thresh_min = f_thresh_max * 0.25;
    if (la->p.keyint_min == la->p.keyint_max) f_thresh_min = f_thresh_max;
    if (i_gop_size <= la->p.keyint_min / 4) f_bias = f_thresh_min / 4;
    else if (i_gop_size <= la->p.keyint_min) f_bias = f_thresh_min * i_gop_size / la->p.keyint_min;
    else f_bias = f_thresh_min + (f_thresh_max - f_thresh_min) * (i_gop_size - la->p.keyint_min) / (la->p.keyint_max - la->p.keyint_min);
    return pcost >= (1.0 - f_bias) * icost;
}

static int scenecut(orc_la *la, frame_t **frames, int p0, int p1, int real_scenecut, int num_frames, int i_max_search)
{
    if (real_scenecut && la->p.bframes) {
        int origmaxp1 = p0 + 1;
        if (la->p.b_adapt == 2) origmaxp1 += la->p.bframes;
        else origmaxp1++;
        int maxp1 = MIN(origmaxp1, num_frames);
        for (int curp1 = p1; curp1 <= maxp1; curp1++)
            if (!scenecut_internal(la, frames, p0, curp1))
                for (int i = curp1; i > p0; i--) frames[i]->b_scenecut = 0;
        for (int curp0 = p0; curp0 <= maxp1; curp0++)
            if (origmaxp1 > i_max_search || (curp0 < maxp1 && scenecut_internal(la, frames, curp0, maxp1)))
                frames[curp0]->b_scenecut = 0;
    }
    if (!frames[p1]->b_scenecut) return 0;
    return scenecut_internal(la, frames, p0, p1);
}

/* [x264] x264_slicetype_analyse */
static void slicetype_analyse(orc_la *la, int intra_minigop)
{
    const orc_la_params *p = &la->p;
    frame_t *frames[ORC_LOOKAHEAD_MAX + 3] = {NULL};
    int num_frames, orig_num_frames, keyint_limit, framecnt;
    int i_max_search = MIN(la->n_next, ORC_LOOKAHEAD_MAX);
    i_max_search = MIN(i_max_search, la->slicetype_length + 1 - intra_minigop);     /* b_deterministic */
    int keyframe = !!intra_minigop;

    if (!la->last_nonb) return;
    frames[0] = la->last_nonb;
    for (framecnt = 0; framecnt < i_max_search; framecnt++) frames[framecnt + 1] = la->next[framecnt];

    if (!framecnt) {
        if (p->b_mbtree) macroblock_tree(la, frames, 0, keyframe);
        return;
    }
    keyint_limit = p->keyint_max - frames[0]->i_frame + la->i_last_keyframe - 1;
    orig_num_frames = num_frames = MIN(framecnt, keyint_limit);
    if (p->b_psy && p->b_mbtree) num_frames = framecnt;
    else if (p->open_gop && num_frames < framecnt) num_frames++;
    else if (num_frames == 0) { frames[1]->i_type = ORC_TYPE_I; return; }

    if (AUTO_OR_I(frames[1]->i_type) && p->scenecut && scenecut(la, frames, 0, 1, 1, orig_num_frames, i_max_search)) {
        if (frames[1]->i_type == ORC_TYPE_AUTO) frames[1]->i_type = ORC_TYPE_I;
        return;
    }
    for (int j = 1; j <= num_frames; j++)
        if (frames[j]->i_type == ORC_TYPE_KEYFRAME) frames[j]->i_type = p->open_gop ? ORC_TYPE_I : ORC_TYPE_IDR;
    for (int j = 2; j <= num_frames; j++)
        if (frames[j]->i_type == ORC_TYPE_IDR && AUTO_OR_B(frames[j - 1]->i_type)) frames[j - 1]->i_type = ORC_TYPE_P;

    int num_analysed_frames = num_frames;
    int reset_start;
    if (p->bframes) {
        if (p->b_adapt == 2) {
            if (num_frames > 1) {
                char best_paths[ORC_BFRAME_MAX + 1][ORC_LOOKAHEAD_MAX + 1];
                memset(best_paths, 0, sizeof(best_paths));
                strcpy(best_paths[1], "P");
                int best_path_index = num_frames % (ORC_BFRAME_MAX + 1);
                for (int j = 2; j <= num_frames; j++) slicetype_path(la, frames, j, best_paths);
                for (int j = 1; j < num_frames; j++) {
                    if (best_paths[best_path_index][j - 1] != 'B') {
                        if (AUTO_OR_B(frames[j]->i_type)) frames[j]->i_type = ORC_TYPE_P;
                    } else {
                        if (frames[j]->i_type == ORC_TYPE_AUTO) frames[j]->i_type = ORC_TYPE_B;
                    }
                }
            }
        } else if (p->b_adapt == 1) {
            int last_nonb = 0, num_bframes = p->bframes;
            char path[ORC_LOOKAHEAD_MAX + 1];
            for (int j = 1; j < num_frames; j++) {
                if (j - 1 > 0 && IS_B(frames[j - 1]->i_type)) num_bframes--;
                else { last_nonb = j - 1; num_bframes = p->bframes; }
                if (!num_bframes) {
                    if (AUTO_OR_B(frames[j]->i_type)) frames[j]->i_type = ORC_TYPE_P;
                    continue;
                }
                if (frames[j]->i_type != ORC_TYPE_AUTO) continue;
                if (IS_B(frames[j + 1]->i_type)) { frames[j]->i_type = ORC_TYPE_P; continue; }
                int bframes = j - last_nonb - 1;
                memset(path, 'B', bframes);
                strcpy(path + bframes, "PP");
                uint64_t cost_p = path_cost(la, frames + last_nonb, path, COST_MAX64);
                strcpy(path + bframes, "BP");
                uint64_t cost_b = path_cost(la, frames + last_nonb, path, cost_p);
                frames[j]->i_type = cost_b < cost_p ? ORC_TYPE_B : ORC_TYPE_P;
            }
        } else {
            int num_bframes = p->bframes;
            for (int j = 1; j < num_frames; j++) {
                if (!num_bframes) {
                    if (AUTO_OR_B(frames[j]->i_type)) frames[j]->i_type = ORC_TYPE_P;
                } else if (frames[j]->i_type == ORC_TYPE_AUTO) {
                    if (IS_B(frames[j + 1]->i_type)) frames[j]->i_type = ORC_TYPE_P;
                    else frames[j]->i_type = ORC_TYPE_B;
                }
                if (IS_B(frames[j]->i_type)) num_bframes--;
                else num_bframes = p->bframes;
            }
        }
        if (AUTO_OR_B(frames[num_frames]->i_type)) frames[num_frames]->i_type = ORC_TYPE_P;

        int num_bframes = 0;
        while (num_bframes < num_frames && IS_B(frames[num_bframes + 1]->i_type)) num_bframes++;
        for (int j = 1; j < num_bframes + 1; j++) {
            if (frames[j]->i_forced_type == ORC_TYPE_AUTO && AUTO_OR_I(frames[j + 1]->i_forced_type) &&
                p->scenecut && scenecut(la, frames, j, j + 1, 0, orig_num_frames, i_max_search)) {
                frames[j]->i_type = ORC_TYPE_P;
                num_analysed_frames = j;
                break;
            }
        }
        reset_start = keyframe ? 1 : MIN(num_bframes + 2, num_analysed_frames + 1);
    } else {
        for (int j = 1; j <= num_frames; j++)
            if (AUTO_OR_B(frames[j]->i_type)) frames[j]->i_type = ORC_TYPE_P;
        reset_start = !keyframe + 1;
    }

    if (p->b_mbtree) macroblock_tree(la, frames, MIN(num_frames, p->keyint_max), keyframe);

    /* enforce keyframe limit */
    {
        int last_keyframe = la->i_last_keyframe, last_possible = 0;
        for (int j = 1; j <= num_frames; j++) {
            frame_t *frm = frames[j];
            int keyframe_dist = frm->i_frame - last_keyframe;
            if (AUTO_OR_I(frm->i_forced_type)) {
                if (p->open_gop || !IS_B(frames[j - 1]->i_forced_type)) last_possible = j;
            }
            if (keyframe_dist >= p->keyint_max) {
                if (last_possible != 0 && last_possible != j) {
                    j = last_possible;
                    frm = frames[j];
                    keyframe_dist = frm->i_frame - last_keyframe;
                }
                last_possible = 0;
                if (frm->i_type != ORC_TYPE_IDR) frm->i_type = p->open_gop ? ORC_TYPE_I : ORC_TYPE_IDR;
            }
            if (frm->i_type == ORC_TYPE_I && keyframe_dist >= p->keyint_min) {
                if (p->open_gop) last_keyframe = frm->i_frame;
                else if (frm->i_forced_type != ORC_TYPE_I) frm->i_type = ORC_TYPE_IDR;
            }
            if (frm->i_type == ORC_TYPE_IDR) {
                last_keyframe = frm->i_frame;
                if (j > 1 && IS_B(frames[j - 1]->i_type)) frames[j - 1]->i_type = ORC_TYPE_P;
            }
        }
    }
    for (int j = reset_start; j <= num_frames; j++) frames[j]->i_type = frames[j]->i_forced_type;
}

static void out_push(orc_la *la, frame_t *f)
{
    if (la->n_out == la->cap_out) { la->cap_out = la->cap_out ? 2 * la->cap_out : 64; la->outq = realloc(la->outq, la->cap_out * sizeof(frame_t *)); }
    la->outq[la->n_out++] = f;
}

/* [x264] x264_slicetype_decide + lookahead_slicetype_decide (shift to the output queue and
 * the keyframe re-analysis) */
static void slicetype_decide_and_shift(orc_la *la)
{
    const orc_la_params *p = &la->p;
    frame_t *frames[ORC_BFRAME_MAX + 3];
    frame_t *frm;
    int bframes, brefs;
    if (!la->n_next) return;

    for (int i = 0; i < la->n_next; i++)
        la->next[i]->f_duration = (float)((double)2 * p->fps_den / ((double)p->fps_num * 2));

    if ((p->bframes && p->b_adapt) || p->scenecut || p->b_mbtree) slicetype_analyse(la, 0);

    for (bframes = 0, brefs = 0;; bframes++) {
        frm = la->next[bframes];
        if (frm->i_type == ORC_TYPE_BREF && p->b_pyramid < 2 && brefs == p->b_pyramid) frm->i_type = ORC_TYPE_B;
        else if (frm->i_type == ORC_TYPE_BREF && p->b_pyramid == 2 && brefs && p->frame_reference <= (brefs + 3)) frm->i_type = ORC_TYPE_B;
        if (frm->i_type == ORC_TYPE_KEYFRAME) frm->i_type = p->open_gop ? ORC_TYPE_I : ORC_TYPE_IDR;
        if (frm->i_frame - la->i_last_keyframe >= p->keyint_max) {
            if (frm->i_type == ORC_TYPE_AUTO || frm->i_type == ORC_TYPE_I)
                frm->i_type = p->open_gop && la->i_last_keyframe >= 0 ? ORC_TYPE_I : ORC_TYPE_IDR;
            int warn = frm->i_type != ORC_TYPE_IDR;
            if (warn && p->open_gop) warn &= frm->i_type != ORC_TYPE_I;
            if (warn) frm->i_type = p->open_gop && la->i_last_keyframe >= 0 ? ORC_TYPE_I : ORC_TYPE_IDR;
        }
        if (frm->i_type == ORC_TYPE_I && frm->i_frame - la->i_last_keyframe >= p->keyint_min) {
            if (p->open_gop) { la->i_last_keyframe = frm->i_frame; frm->b_keyframe = 1; }
            else frm->i_type = ORC_TYPE_IDR;
        }
        if (frm->i_type == ORC_TYPE_IDR) {
            la->i_last_keyframe = frm->i_frame;
            frm->b_keyframe = 1;
            if (bframes > 0) { bframes--; la->next[bframes]->i_type = ORC_TYPE_P; }
        }
        if (bframes == p->bframes || bframes + 1 >= la->n_next) {
            if (frm->i_type == ORC_TYPE_AUTO || IS_B(frm->i_type)) frm->i_type = ORC_TYPE_P;
        }
        if (frm->i_type == ORC_TYPE_BREF) brefs++;
        if (frm->i_type == ORC_TYPE_AUTO) frm->i_type = ORC_TYPE_B;
        else if (!IS_B(frm->i_type)) break;
    }
    la->next[bframes]->i_bframes = bframes;
    if (p->b_pyramid && bframes > 1 && !brefs) { la->next[(bframes - 1) / 2]->i_type = ORC_TYPE_BREF; brefs++; }

    /* precompute the frame cost ratecontrol will ask for (rc method != CQP) */
    {
        int p0, p1, b;
        p1 = b = bframes + 1;
        frames[0] = la->last_nonb;
        memcpy(&frames[1], la->next, (bframes + 1) * sizeof(frame_t *));
        if (IS_I(la->next[bframes]->i_type)) p0 = bframes + 1;
        else p0 = 0;
        frame_cost(la, frames, p0, p1, b);
        la->next[bframes]->rc_d0 = b - p0; la->next[bframes]->rc_d1 = p1 - b;
    }
    /* the full-resolution x264_weights_analyse(...,0) for P frames stays on the CPU encoder */

    /* shift to coded order: non-B first, then BREF, then B */
    out_push(la, la->next[bframes]);
    for (int i = 0; i < bframes; i++) if (la->next[i]->i_type == ORC_TYPE_BREF) out_push(la, la->next[i]);
    for (int i = 0; i < bframes; i++) if (la->next[i]->i_type != ORC_TYPE_BREF) out_push(la, la->next[i]);

    /* lookahead_update_last_nonb + lookahead_shift */
    la->last_nonb = la->next[bframes];
    int shift = bframes + 1;
    memmove(la->next, la->next + shift, (la->n_next - shift) * sizeof(frame_t *));
    la->n_next -= shift;

    /* b_analyse_keyframe: mb-tree (or vbv) lookahead re-analyses after an I-frame */
    if (p->b_mbtree && IS_I(la->last_nonb->i_type)) slicetype_analyse(la, shift);
}

/* ------------------------------------------------------------------------------------------
 * public API
 * ---------------------------------------------------------------------------------------- */
void orc_la_params_preset(orc_la_params *p, const char *preset, int width, int height)
{
    /* x264 defaults (x264_param_default) + the preset deltas x264vfw documents at
     * config.c:1460-1498; rate control is the wrapper's default CRF (config.c:109-111). */
    memset(p, 0, sizeof(*p));
    p->width = width; p->height = height; p->chroma_format = 1;
    p->bframes = 3; p->b_adapt = 1; p->b_pyramid = 2; p->b_bias = 0;
    p->rc_lookahead = 40; p->b_mbtree = 1; p->scenecut = 40;
    p->keyint_max = 250; p->keyint_min = 25; p->open_gop = 0;
    p->weightp = 2; p->weightb = 1; p->subme = 7; p->me_method = 1; p->me_range = 16; p->mv_range = 512;
    p->aq_mode = 1; p->aq_strength = 1.0f; p->qcompress = 0.6f; p->frame_reference = 3;
    p->lookahead_threads = 1; p->fps_num = 25; p->fps_den = 1; p->b_psy = 1;
    if (!strcmp(preset, "ultrafast")) {
        p->frame_reference = 1; p->scenecut = 0; p->bframes = 0; p->b_adapt = 0; p->me_method = 0; p->subme = 0;
        p->aq_mode = 0; p->b_mbtree = 0; p->rc_lookahead = 0; p->weightp = 0; p->weightb = 0;
    } else if (!strcmp(preset, "superfast")) {
        p->me_method = 0; p->subme = 1; p->frame_reference = 1; p->b_mbtree = 0; p->rc_lookahead = 0; p->weightp = 1;
    } else if (!strcmp(preset, "veryfast")) {
        p->subme = 2; p->frame_reference = 1; p->weightp = 1; p->rc_lookahead = 10;
    } else if (!strcmp(preset, "faster")) {
        p->frame_reference = 2; p->subme = 4; p->weightp = 1; p->rc_lookahead = 20;
    } else if (!strcmp(preset, "fast")) {
        p->frame_reference = 2; p->subme = 6; p->weightp = 1; p->rc_lookahead = 30;
    } else if (!strcmp(preset, "slow")) {
        p->subme = 8; p->frame_reference = 5; p->rc_lookahead = 50;
    } else if (!strcmp(preset, "slower")) {
        p->me_method = 2; p->subme = 9; p->frame_reference = 8; p->b_adapt = 2; p->rc_lookahead = 60;
    } else if (!strcmp(preset, "veryslow")) {
        p->me_method = 2; p->subme = 10; p->me_range = 24; p->frame_reference = 16; p->b_adapt = 2; p->bframes = 8; p->rc_lookahead = 60;
    } else if (!strcmp(preset, "placebo")) {
        p->me_method = 4; p->subme = 11; p->me_range = 24; p->frame_reference = 16; p->b_adapt = 2; p->bframes = 16; p->rc_lookahead = 60;
    }
}

orc_la *orc_la_open(const orc_la_params *p)
{
    init_tables();
    orc_la *la = calloc(1, sizeof(*la));
    la->p = *p;
    if (la->p.bframes > ORC_BFRAME_MAX) la->p.bframes = ORC_BFRAME_MAX;
    if (la->p.rc_lookahead > ORC_LOOKAHEAD_MAX) la->p.rc_lookahead = ORC_LOOKAHEAD_MAX;
    if (la->p.keyint_min <= 0) la->p.keyint_min = MIN(la->p.keyint_max / 10, la->p.fps_num / MAX(1, la->p.fps_den));
    /* [x264] encoder/encoder.c validate_parameters: weightp off + mb-tree + psy => X264_WEIGHTP_FAKE (-1):
     * the lookahead still analyses (and uses) luma weights and records f_weighted_cost_delta */
    if (!la->p.weightp && la->p.b_mbtree && la->p.b_psy) la->p.weightp = ORC_WEIGHTP_FAKE;
    int g[10];
    orc_lowres_geometry(p->width, p->height, g);
    la->mb_w = g[0]; la->mb_h = g[1]; la->mb_count = g[0] * g[1];
    la->luma_w = g[2]; la->luma_h = g[3]; la->lw = g[5]; la->lh = g[6];
    la->lstride = g[7]; la->lplane = g[8]; la->lorigin = g[9];
    la->cost_mv = (uint16_t *)orc_cost_mv_table(p->mv_range, &la->cost_mv_half) + la->cost_mv_half;
    /* lowres_context_init */
    if (p->subme > 1) { la->la_me_method = MIN(1, p->me_method); la->la_subpel_refine = 4; }
    else { la->la_me_method = 0; la->la_subpel_refine = 2; }
    la->la_satd = p->subme > 1;
    la->slicetype_length = MAX(la->p.bframes, la->p.rc_lookahead);
    la->i_last_keyframe = -la->p.keyint_max;
    la->weight_buf = malloc(la->lplane);
    la->scratch_amount = malloc(la->mb_w * sizeof(int16_t));
    return la;
}

void orc_la_close(orc_la *la)
{
    if (!la) return;
    for (int i = 0; i < la->n_all; i++) frame_free(la, la->all[i]);
    free(la->all); free(la->outq); free(la->weight_buf); free(la->scratch_amount);
    free(la);
}

int orc_la_put_frame(orc_la *la, const uint8_t *y, int ys, const uint8_t *u, const uint8_t *v, int cs)
{
    frame_t *f = frame_new(la);
    f->i_frame = la->n_all;
    if (la->n_all == la->cap_all) { la->cap_all = la->cap_all ? 2 * la->cap_all : 64; la->all = realloc(la->all, la->cap_all * sizeof(frame_t *)); }
    la->all[la->n_all++] = f;
    adaptive_quant_frame(la, f, y, ys, u, v, cs);
    orc_lowres_init(f->lowres_buf, y, ys, la->p.width, la->p.height);
    la->next[la->n_next++] = f;
    while (la->n_next > la->slicetype_length) slicetype_decide_and_shift(la);
    return la->n_out - la->out_head;
}

int orc_la_flush(orc_la *la)
{
    while (la->n_next) slicetype_decide_and_shift(la);
    return la->n_out - la->out_head;
}

int orc_la_get_decision(orc_la *la, orc_la_decision *d, float *qp_offset, float *qp_offset_aq)
{
    if (la->out_head >= la->n_out) return 0;
    frame_t *f = la->outq[la->out_head++];
    d->i_frame = f->i_frame; d->i_type = f->i_type; d->b_keyframe = f->b_keyframe; d->i_bframes = f->i_bframes;
    d->mb_count = la->mb_count;
    d->i_cost_est = d->i_cost_est_aq = d->i_intra_mbs = -1;
    if (f->rc_d0 >= 0) {
        d->i_cost_est = f->cost_est[f->rc_d0][f->rc_d1];
        d->i_cost_est_aq = f->cost_est_aq[f->rc_d0][f->rc_d1];
        d->i_intra_mbs = f->intra_mbs[f->rc_d0];
    }
    if (qp_offset) memcpy(qp_offset, f->qp_offset, la->mb_count * sizeof(float));
    if (qp_offset_aq) memcpy(qp_offset_aq, f->qp_offset_aq, la->mb_count * sizeof(float));
    return 1;
}

int orc_la_mb_count(orc_la *la) { return la->mb_count; }
int orc_la_frame_cost(orc_la *la, int p0, int p1, int b)
{
    if (p0 < 0 || p1 >= la->n_all || b < p0 || b > p1 || b - p0 > la->p.bframes + 1 || p1 - b > la->p.bframes + 1) return -1;
    return frame_cost(la, la->all, p0, p1, b);
}
const uint8_t *orc_la_lowres_planes(orc_la *la, int f) { return la->all[f]->lowres_buf; }
const uint16_t *orc_la_intra_cost(orc_la *la, int f) { return la->all[f]->intra_cost; }
const uint16_t *orc_la_inv_qscale(orc_la *la, int f) { return la->all[f]->inv_qscale; }
const uint16_t *orc_la_propagate_cost(orc_la *la, int f) { return la->all[f]->propagate_cost; }
const float *orc_la_qp_offset(orc_la *la, int f, int aq) { return aq ? la->all[f]->qp_offset_aq : la->all[f]->qp_offset; }
const int16_t *orc_la_mvs(orc_la *la, int f, int list, int dist) { return (const int16_t *)la->all[f]->mvs[list][dist - 1]; }
const int *orc_la_mv_costs(orc_la *la, int f, int list, int dist) { return la->all[f]->mv_costs[list][dist - 1]; }
const uint16_t *orc_la_lowres_costs(orc_la *la, int f, int d0, int d1) { return la->all[f]->lowres_costs[d0][d1]; }
const int *orc_la_row_satds(orc_la *la, int f, int d0, int d1) { return la->all[f]->row_satds[d0][d1]; }
int orc_la_cost_est(orc_la *la, int f, int d0, int d1, int aq) { return aq ? la->all[f]->cost_est_aq[d0][d1] : la->all[f]->cost_est[d0][d1]; }
int orc_la_intra_mbs(orc_la *la, int f, int d0) { return la->all[f]->intra_mbs[d0]; }
void orc_la_pixel_stats(orc_la *la, int f, uint64_t sum[3], uint64_t ssd[3])
{
    for (int i = 0; i < 3; i++) { sum[i] = la->all[f]->pixel_sum[i]; ssd[i] = la->all[f]->pixel_ssd[i]; }
}
void orc_la_weight(orc_la *la, int f, int out[4])
{
    weight_t *w = &la->all[f]->weight;
    out[0] = w->scale; out[1] = w->denom; out[2] = w->offset; out[3] = w->on;
}
void orc_la_mbtree(orc_la *la, const int *frame_idx, const int *types, int num_frames, int b_intra)
{
    frame_t *frames[ORC_LOOKAHEAD_MAX + 3];
    for (int i = 0; i <= num_frames; i++) {
        frames[i] = la->all[frame_idx[i]];
        frames[i]->i_type = types[i];
        frames[i]->f_duration = (float)((double)la->p.fps_den / la->p.fps_num);
    }
    macroblock_tree(la, frames, num_frames, b_intra);
}
/* ------------------------------------------------------------------------------------------
 * SURVEY 8(f) row 3, remainder: [x264] x264_weights_analyse( h, fenc, ref, 0 ) -- the ENCODER-side explicit
 * weight analysis of a P frame against its nearest reference (called from x264_encoder_encode via
 * reference_build_list / weighted_reference_duplicate when weightp >= 1), and the helpers it uses in
 * [x264] encoder/slicetype.c: weight_cost_init_luma, weight_cost_luma, weight_cost_init_chroma,
 * weight_cost_chroma, weight_slice_header_cost, x264_weight_get_h264; [x264] common/mc.c: mc_chroma;
 * [x264] common/pixel.c: pixel_asd8.  Restated from upstream as the rest of this file (PARITY UNPINNED).
 *
 * What differs from the lookahead's call (b_lookahead = 1, weights_analyse above):
 *   - luma is still scored on the LOWRES planes, but the reference is first motion compensated with the
 *     lookahead's own list-0 vectors for that distance (when that search was run), 8x8 by 8x8;
 *   - both chroma planes are scored at full resolution on the deinterleaved NV12 planes, the reference
 *     compensated by mc_chroma with the SAME lowres vector of the MB (upstream passes the lowres quarter-pel
 *     vector where a full-resolution one is expected, i.e. half the displacement: kept), cost = |sum of block
 *     differences| (asd8: only the DC matters), header cost with 4 x lambda;
 *   - scale and offset are searched in a window that grows with subme (weight_check_distance);
 *   - the offset guess is truncated (no + 0.5), chroma shares one denominator.
 * fenc_uv / ref_uv: NV12 chroma planes of the frames padded to mod 16 (8*mb_w pairs x 8*mb_h rows, 4:2:0);
 * references past the edge replicate the edge pair, which is what upstream's x264_frame_expand_border_chroma
 * leaves there.  out[plane] = {on, scale, denom, offset}.
 * ---------------------------------------------------------------------------------------- */
static inline int uv_at(const uint8_t *uv, int stride, int cw, int ch, int x, int y, int c)
{
    x = x < 0 ? 0 : x >= cw ? cw - 1 : x;
    y = y < 0 ? 0 : y >= ch ? ch - 1 : y;
    return uv[(ptrdiff_t)y * stride + 2 * x + c];
}

/* [x264] mc_chroma for one 8x8 block of component c; (bx, by) = block origin in chroma pixels */
static void mc_chroma_8x8(uint8_t dst[64], const uint8_t *uv, int stride, int cw, int ch, int bx, int by, int mvx, int mvy, int c)
{
    const int d8x = mvx & 7, d8y = mvy & 7;
    const int cA = (8 - d8x) * (8 - d8y), cB = d8x * (8 - d8y), cC = (8 - d8x) * d8y, cD = d8x * d8y;
    const int x0 = bx + (mvx >> 3), y0 = by + (mvy >> 3);
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++)
            dst[y * 8 + x] = (cA * uv_at(uv, stride, cw, ch, x0 + x, y0 + y, c) + cB * uv_at(uv, stride, cw, ch, x0 + x + 1, y0 + y, c) +
                              cC * uv_at(uv, stride, cw, ch, x0 + x, y0 + y + 1, c) + cD * uv_at(uv, stride, cw, ch, x0 + x + 1, y0 + y + 1, c) + 32) >> 6;
}

static int weight_header_cost(const weight_t *w, int b_chroma)
{
    const int lambda = b_chroma ? 4 : 1;                      /* x264_lambda_tab[X264_LOOKAHEAD_QP] = 1 */
    const int denom_cost = ue_size(w->denom) * (2 - b_chroma);
    return lambda * 1 * (10 + denom_cost + 2 * (se_size(w->scale) + se_size(w->offset)));     /* numslices = 1 */
}

typedef struct {
    orc_la *la; frame_t *fenc, *ref; int searched, dist;
    const uint8_t *fenc_uv, *ref_uv; int uv_stride;
} wfull_t;

static unsigned wfull_cost(wfull_t *c, int plane, const weight_t *w)
{
    orc_la *la = c->la;
    unsigned cost = 0;
    int i_mb = 0;
    if (!plane) {
        const int stride = la->lstride;
        for (int y = 0; y < la->lh; y += 8)
            for (int x = 0; x < la->lw; x += 8, i_mb++) {
                uint8_t buf[64];
                if (c->searched) {                              /* weight_cost_init_luma */
                    const int16_t *mv = c->fenc->mvs[0][c->dist][i_mb];
                    get_ref_8x8(buf, c->ref->lowres, stride, mv[0] + (x << 2), mv[1] + (y << 2), NULL);
                } else
                    for (int yy = 0; yy < 8; yy++) memcpy(buf + yy * 8, c->ref->lowres[0] + (y + yy) * stride + x, 8);
                if (w) for (int i = 0; i < 64; i++) buf[i] = weight_px(w, buf[i]);
                int cmp = mbcmp(la, buf, 8, c->fenc->lowres[0] + y * stride + x, stride);
                cost += MIN(cmp, c->fenc->intra_cost[i_mb]);
            }
        if (w) cost += weight_header_cost(w, 0);
        return cost;
    }
    const int cw = 8 * la->mb_w, ch = 8 * la->mb_h, comp = plane - 1;
    for (int y = 0; y < ch; y += 8)
        for (int x = 0; x < cw; x += 8, i_mb++) {
            uint8_t buf[64];
            if (c->searched) {                                  /* weight_cost_init_chroma */
                const int16_t *mv = c->fenc->mvs[0][c->dist][i_mb];
                mc_chroma_8x8(buf, c->ref_uv, c->uv_stride, cw, ch, x, y, mv[0], mv[1], comp);    /* 2*mvy >> v_shift, 4:2:0 */
            } else
                for (int i = 0; i < 64; i++) buf[i] = c->ref_uv[(ptrdiff_t)(y + (i >> 3)) * c->uv_stride + 2 * (x + (i & 7)) + comp];
            int sum = 0;                                        /* pixel_asd8 */
            for (int i = 0; i < 64; i++) {
                int r = w ? weight_px(w, buf[i]) : buf[i];
                sum += r - c->fenc_uv[(ptrdiff_t)(y + (i >> 3)) * c->uv_stride + 2 * (x + (i & 7)) + comp];
            }
            cost += abs(sum);
        }
    if (w) cost += weight_header_cost(w, 1);
    return cost;
}

int orc_la_weights_full(orc_la *la, int f_enc, int f_ref, const uint8_t *fenc_uv, const uint8_t *ref_uv, int uv_stride,
                        int out[3][4], float *cost_delta)
{
    static const uint8_t weight_check_distance[][2] = {{0, 0}, {0, 0}, {0, 1}, {0, 1}, {0, 1}, {0, 1}, {0, 1}, {1, 1}, {1, 1}, {2, 1}, {2, 1}, {4, 2}};
    if (la->p.chroma_format != 1) return -1;                    /* 4:2:0 only here */
    frame_t *fenc = la->all[f_enc], *ref = la->all[f_ref];
    const int dist = fenc->i_frame - ref->i_frame - 1;
    if (dist < 0 || dist > ORC_BFRAME_MAX) return -1;
    const float epsilon = 1.f / 128.f;
    weight_t weights[3] = {{0, 1, 0, 0}, {0, 1, 0, 0}, {0, 1, 0, 0}};
    float guess_scale[3], fenc_mean[3], ref_mean[3];
    const int dims[2] = {la->luma_h * la->luma_w, (la->luma_h / 2) * (la->luma_w / 2)};
    for (int plane = 0; plane < 3; plane++) {
        int zero_bias = !ref->pixel_ssd[plane];
        float fenc_var = fenc->pixel_ssd[plane] + zero_bias;
        float ref_var = ref->pixel_ssd[plane] + zero_bias;
        guess_scale[plane] = sqrtf(fenc_var / ref_var);
        fenc_mean[plane] = (float)(fenc->pixel_sum[plane] + zero_bias) / dims[!!plane] / 1;
        ref_mean[plane] = (float)(ref->pixel_sum[plane] + zero_bias) / dims[!!plane] / 1;
    }
    int chroma_denom = 7;
    while (chroma_denom > 0) {                                   /* make sure both chroma scales fit */
        float thresh = 127.f / (1 << chroma_denom);
        if (guess_scale[1] < thresh && guess_scale[2] < thresh) break;
        chroma_denom--;
    }
    wfull_t c = {la, fenc, ref, fenc->mvs_searched[0][dist], dist, fenc_uv, ref_uv, uv_stride};
    const int subme = la->p.subme < 0 ? 0 : la->p.subme > 11 ? 11 : la->p.subme;
    if (cost_delta) *cost_delta = 0;

    for (int plane = 0; plane < 3 && !(plane && !weights[0].on); plane++) {
        if (fabsf(ref_mean[plane] - fenc_mean[plane]) < 0.5f && fabsf(1.f - guess_scale[plane]) < epsilon) {
            weights[plane] = (weight_t){0, 1, 0, 0};
            continue;
        }
        if (plane) {
            weights[plane].denom = chroma_denom;
            weights[plane].scale = clip3((int)round(guess_scale[plane] * (1 << chroma_denom)), 0, 255);
            if (weights[plane].scale > 127) { weights[1].on = weights[2].on = 0; break; }
        } else {
            int scale = (int)round(guess_scale[0] * 128), denom = 7;     /* x264_weight_get_h264 */
            while (denom > 0 && scale > 127) { denom--; scale >>= 1; }
            weights[0].scale = MIN(scale, 127); weights[0].denom = denom; weights[0].offset = 0;
        }
        int found = 0, mindenom = weights[plane].denom, minscale = weights[plane].scale, minoff = 0;
        if (!plane && !fenc->b_intra_calculated) { frame_t *one[1] = {fenc}; frame_cost(la, one, 0, 0, 0); }
        unsigned origscore, minscore;
        origscore = minscore = wfull_cost(&c, plane, NULL);
        if (!minscore) continue;

        const int scale_dist = weight_check_distance[subme][0], offset_dist = weight_check_distance[subme][1];
        const int start_scale = clip3(minscale - scale_dist, 0, 127), end_scale = clip3(minscale + scale_dist, 0, 127);
        for (int i_scale = start_scale; i_scale <= end_scale; i_scale++) {
            int cur_scale = i_scale;
            int cur_offset = fenc_mean[plane] - ref_mean[plane] * cur_scale / (1 << mindenom) + 0.5f * 0;
            if (cur_offset < -128 || cur_offset > 127) {
                cur_offset = clip3(cur_offset, -128, 127);
                cur_scale = clip3f((1 << mindenom) * (fenc_mean[plane] - cur_offset) / ref_mean[plane] + 0.5f, 0, 127);
            }
            const int start_offset = clip3(cur_offset - offset_dist, -128, 127), end_offset = clip3(cur_offset + offset_dist, -128, 127);
            for (int i_off = start_offset; i_off <= end_offset; i_off++) {
                weight_t w = {1, cur_scale, mindenom, i_off};
                unsigned sc = wfull_cost(&c, plane, &w);
                if (sc < minscore) { minscore = sc; minscale = cur_scale; minoff = i_off; found = 1; }
                /* don't check any more offsets if the previous one had a lower cost than the current one */
                if (minoff == start_offset && i_off != start_offset) break;
            }
        }
        if (!plane)
            while (mindenom > 0 && !(minscale & 1)) { mindenom--; minscale >>= 1; }
        if (!found || (minscale == 1 << mindenom && minoff == 0) || (float)minscore / origscore > 0.998f) {
            weights[plane] = (weight_t){0, 1, 0, 0};
            continue;
        }
        weights[plane] = (weight_t){1, minscale, mindenom, minoff};
        if (la->p.weightp == ORC_WEIGHTP_FAKE && !plane && cost_delta) *cost_delta = (float)minscore / origscore;
    }
    /* optimise and unify the chroma denominator */
    if (weights[1].on || weights[2].on) {
        int denom = weights[1].on ? weights[1].denom : weights[2].denom;
        const int both = weights[1].on && weights[2].on;
        while ((!both && denom == 7) ||
               (denom > 0 && !(weights[1].on && (weights[1].scale & 1)) && !(weights[2].on && (weights[2].scale & 1)))) {
            denom--;
            for (int i = 1; i <= 2; i++)
                if (weights[i].on) { weights[i].scale >>= 1; weights[i].denom = denom; }
        }
    }
    for (int i = 0; i < 3; i++) {
        if (!weights[i].on) weights[i] = (weight_t){0, 1, 0, 0};      /* weightfn == NULL: the fields are not looked at */
        out[i][0] = weights[i].on; out[i][1] = weights[i].scale; out[i][2] = weights[i].denom; out[i][3] = weights[i].offset;
    }
    return 0;
}

/* test hook: a whole chroma picture compensated by mc_chroma with one vector (luma quarter-sample units, what mc_chroma
 * takes), 8x8 block by 8x8 block -- pinned against the H.264 decoder's chroma prediction in tests/test_h264_pins.py */
void orc_test_mc_chroma_picture(uint8_t *out_u, uint8_t *out_v, const uint8_t *uv, int stride, int cw, int ch, int mvx, int mvy)
{
    for (int by = 0; by < ch; by += 8)
        for (int bx = 0; bx < cw; bx += 8)
            for (int c = 0; c < 2; c++) {
                uint8_t blk[64];
                mc_chroma_8x8(blk, uv, stride, cw, ch, bx, by, mvx, mvy, c);
                uint8_t *o = c ? out_v : out_u;
                for (int y = 0; y < 8 && by + y < ch; y++)
                    for (int x = 0; x < 8 && bx + x < cw; x++) o[(by + y) * cw + bx + x] = blk[y * 8 + x];
            }
}

/* test hook: one score of the analysis above */
unsigned orc_test_weights_full_cost(orc_la *la, int f_enc, int f_ref, const uint8_t *fenc_uv, const uint8_t *ref_uv, int uv_stride,
                                    int plane, int weighted, int scale, int denom, int offset)
{
    frame_t *fenc = la->all[f_enc], *ref = la->all[f_ref];
    const int dist = fenc->i_frame - ref->i_frame - 1;
    wfull_t c = {la, fenc, ref, fenc->mvs_searched[0][dist], dist, fenc_uv, ref_uv, uv_stride};
    weight_t w = {1, scale, denom, offset};
    if (!plane && !fenc->b_intra_calculated) { frame_t *one[1] = {fenc}; frame_cost(la, one, 0, 0, 0); }
    return wfull_cost(&c, plane, weighted ? &w : NULL);
}

/* ------------------------------------------------------------------------------------------
 * [x264] the integral image x264_frame_filter builds after the half-pel planes when the encoder searches
 * exhaustively (me esa / tesa; [x264] common/mc.c integral_init8h / integral_init8v, and 4h / 4v for the
 * 4x4 plane of --partitions p4x4).  Upstream fills it row by row in place (a horizontal running sum added
 * to the row above, then a vertical difference 8 rows later), in uint16 arithmetic that wraps.  What the
 * exhaustive search reads afterwards is, for every position whose 8x8 (4x4) window lies inside the padded
 * plane:  sum8[y][x] = sum of the 64 pixels of the window with top-left (x, y)   (mod 2^16)
 * This function runs upstream's row recurrences (running horizontal sum + row above, vertical difference) on a padded plane (pad = the 32-pixel border of the
 * half-pel planes; plane points at the top-left corner of the PADDED plane, rows = height + 64) so that
 * the device's direct box sums can be compared against them over exactly that domain.
 * sum8 / sum4: rows x stride uint16 (sum4 may be NULL).  Valid output rows: 0 .. rows-9 (rows-5 for sum4),
 * columns 0 .. stride-9.
 * ---------------------------------------------------------------------------------------- */
void orc_integral_init(uint16_t *sum8, uint16_t *sum4, const uint8_t *plane, int stride, int rows)
{
    /* running buffers laid out like upstream's: row r+1 of `acc` = horizontal sums of pixel row r + row r of acc */
    uint16_t *acc8 = calloc((size_t)(rows + 1) * stride, sizeof(uint16_t));
    uint16_t *acc4 = sum4 ? calloc((size_t)(rows + 1) * stride, sizeof(uint16_t)) : NULL;
    for (int y = 0; y < rows; y++) {
        const uint8_t *pix = plane + (size_t)y * stride;
        uint16_t *s = acc8 + (size_t)(y + 1) * stride;
        int v = pix[0] + pix[1] + pix[2] + pix[3] + pix[4] + pix[5] + pix[6] + pix[7];        /* integral_init8h */
        for (int x = 0; x < stride - 8; x++) { s[x] = (uint16_t)(v + s[x - stride]); v += pix[x + 8] - pix[x]; }
        if (acc4) {
            s = acc4 + (size_t)(y + 1) * stride;
            v = pix[0] + pix[1] + pix[2] + pix[3];                                            /* integral_init4h */
            for (int x = 0; x < stride - 4; x++) { s[x] = (uint16_t)(v + s[x - stride]); v += pix[x + 4] - pix[x]; }
        }
    }
    for (int y = 0; y + 8 <= rows; y++)                                                       /* integral_init8v */
        for (int x = 0; x < stride - 8; x++)
            sum8[(size_t)y * stride + x] = (uint16_t)(acc8[(size_t)(y + 8) * stride + x] - acc8[(size_t)y * stride + x]);
    if (sum4)
        for (int y = 0; y + 4 <= rows; y++)                                                   /* integral_init4v, 4x4 plane */
            for (int x = 0; x < stride - 4; x++)
                sum4[(size_t)y * stride + x] = (uint16_t)(acc4[(size_t)(y + 4) * stride + x] - acc4[(size_t)y * stride + x]);
    free(acc8); free(acc4);
}

void orc_la_counters(orc_la *la, uint64_t out[4]) { out[0] = la->n_mbcost; out[1] = la->n_search; out[2] = la->n_sad; out[3] = la->n_satd; }

/* ---- test hooks (see lookahead_oracle.h) ---- */
void orc_test_get_ref_8x8(uint8_t dst[64], const uint8_t *p0, const uint8_t *p1, const uint8_t *p2, const uint8_t *p3,
                          int stride, int mvx, int mvy)
{
    uint8_t *const planes[4] = {(uint8_t *)p0, (uint8_t *)p1, (uint8_t *)p2, (uint8_t *)p3};
    get_ref_8x8(dst, planes, stride, mvx, mvy, NULL);
}

void orc_test_intra_pred_8x8(uint8_t dst[64], int kind, const uint8_t *src, int stride)
{
    nbr_t n;
    n.tl = src[-stride - 1];
    for (int i = 0; i < 16; i++) n.top[i] = src[-stride + i];
    for (int i = 0; i < 8; i++) n.left[i] = src[i * stride - 1];
    if (kind == 0) pred_8x8c_dc(dst, &n);
    else if (kind == 1) pred_8x8c_h(dst, &n);
    else if (kind == 2) pred_8x8c_v(dst, &n);
    else if (kind == 3) pred_8x8c_p(dst, &n);
    else {
        edge_t e;
        filter_edges(&e, &n);
        pred_8x8_mode(dst, &e, kind - 10);
    }
}

void orc_test_get_ref_8x8_weighted(uint8_t dst[64], const uint8_t *p0, const uint8_t *p1, const uint8_t *p2, const uint8_t *p3,
                                   int stride, int mvx, int mvy, int scale, int denom, int offset)
{
    uint8_t *const planes[4] = {(uint8_t *)p0, (uint8_t *)p1, (uint8_t *)p2, (uint8_t *)p3};
    const weight_t w = {1, scale, denom, offset};
    get_ref_8x8(dst, planes, stride, mvx, mvy, &w);
}

int orc_test_bipred_weight(int p0, int p1, int b, int weightb)
{
    return la_bipred_weight(weightb, la_dist_scale_factor(p0, p1, b));
}

void orc_test_pixel_avg_8x8(uint8_t dst[64], const uint8_t a[64], const uint8_t b[64], int weight)
{
    pixel_avg_8x8(dst, a, 8, b, 8, weight);
}

void orc_test_predict_picture(uint8_t *dst, const uint8_t *planes, int stride, size_t plane_bytes, int w, int h,
                              int mvx, int mvy, int scale, int denom, int offset)
{
    const weight_t wt = {1, scale, denom, offset};
    uint8_t blk[64];
    for (int by = 0; by < h; by += 8)
        for (int bx = 0; bx < w; bx += 8) {
            uint8_t *pl[4];
            for (int p = 0; p < 4; p++) pl[p] = (uint8_t *)planes + p * plane_bytes + (size_t)(by + PAD) * stride + bx + PAD;
            get_ref_8x8(blk, pl, stride, mvx, mvy, scale < 0 ? NULL : &wt);
            for (int y = 0; y < 8; y++) memcpy(dst + (size_t)(by + y) * w + bx, blk + 8 * y, 8);
        }
}

int orc_test_sad_8x8(const uint8_t a[64], const uint8_t b[64]) { return sad_8x8(a, 8, b, 8); }
int orc_test_satd_8x8(const uint8_t a[64], const uint8_t b[64]) { return satd_8x8(a, 8, b, 8); }
