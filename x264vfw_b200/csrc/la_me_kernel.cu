// Lowres motion search: the serial core of the lookahead, as a row-pipelined wavefront.
//
// Restates, for one (frame, reference, list) pair, the search that [x264]
// encoder/slicetype.c: slicetype_mb_cost runs per 8x8 lowres MB in reverse raster order:
//   median MV predictor from the 4 already-visited neighbours (right, below, below-left,
//   below-right), the mv0 SATD<64 early skip, [x264] encoder/me.c: x264_me_search_ref
//   (sub-pel or full-pel predictor candidates, zero check, hexagon + square refine or
//   diamond, on SAD) and refine_subpel (half-pel diamond on SAD, SATD re-score, quarter-pel
//   diamond on SATD).  Ties break exactly as upstream (strict <, candidate order).
//
// Mapping to the GPU: MB (x,y) depends on (x+1,y) and on row y+1 up to x-1, so every row is
// a 2-MB-skewed pipeline stage.  One warp owns one MB row; it walks x downwards and spins on
// the progress counter of the row below (acquire/release through L2).  Rows are handed out
// by an atomic ticket so a waiting warp only ever waits on warps that already started.
// Inside an MB the warp evaluates up to 8 candidate positions at once: 4 lanes per
// candidate, 2 block rows per lane, __vsadu4 for SAD and shuffle butterflies for the 4x4
// Hadamard of SATD -- all on the integer pipe (this is not a dense contraction).
// Several searches (both lists of a B evaluation) share one launch via blockIdx.y.
#include "la_common.cuh"

namespace xv {

#define BIG_COST 0x3fffffff

__device__ __forceinline__ int ld_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_relaxed(const int *p)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

struct WarpMb {
    // per-lane view of the current MB
    uint2 fe0, fe1;              // fenc rows 2*rp and 2*rp+1
    int rp, grp, lane;
    int stride, pel;
    const uint8_t *fref[4];
    const uint8_t *fref_w;
    WeightDev w;
    const uint16_t *cost_mv;
    int mvp_x, mvp_y;
    int satd;
};

__device__ __forceinline__ int group_sum(int v)
{
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}
__device__ __forceinline__ int warp_min(int v)
{
    v = min(v, __shfl_xor_sync(0xffffffffu, v, 4));
    v = min(v, __shfl_xor_sync(0xffffffffu, v, 8));
    v = min(v, __shfl_xor_sync(0xffffffffu, v, 16));
    return v;
}

__device__ __forceinline__ int mvcost(const WarpMb &m, int qx, int qy)
{
    return (int)__ldg(m.cost_mv + (qx - m.mvp_x)) + (int)__ldg(m.cost_mv + (qy - m.mvp_y));
}

__device__ __forceinline__ int rows_sad(const WarpMb &m, uint2 a0, uint2 a1)
{
    return __vsadu4(a0.x, m.fe0.x) + __vsadu4(a0.y, m.fe0.y) + __vsadu4(a1.x, m.fe1.x) + __vsadu4(a1.y, m.fe1.y);
}

// SATD of an 8x8 block spread over the 4 lanes of a group (2 rows per lane).
// Rows 0-3 live in lanes rp 0,1 and rows 4-7 in lanes rp 2,3: each lane pair forms one
// 8x4 ([x264] x264_pixel_satd_8x4: two 4x4 Hadamards, sum |coef|, >>1), the two halves add.
__device__ __forceinline__ int rows_satd(const WarpMb &m, uint2 a0, uint2 a1)
{
    int s[8], d[8];
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const uint32_t f0 = half ? m.fe0.y : m.fe0.x, f1 = half ? m.fe1.y : m.fe1.x;
        const uint32_t r0 = half ? a0.y : a0.x, r1 = half ? a1.y : a1.x;
        int t0[4], t1[4];
        {
            int e0 = (int)(f0 & 0xff) - (int)(r0 & 0xff), e1 = (int)((f0 >> 8) & 0xff) - (int)((r0 >> 8) & 0xff);
            int e2 = (int)((f0 >> 16) & 0xff) - (int)((r0 >> 16) & 0xff), e3 = (int)(f0 >> 24) - (int)(r0 >> 24);
            int s01 = e0 + e1, d01 = e0 - e1, s23 = e2 + e3, d23 = e2 - e3;
            t0[0] = s01 + s23; t0[1] = s01 - s23; t0[2] = d01 + d23; t0[3] = d01 - d23;
        }
        {
            int e0 = (int)(f1 & 0xff) - (int)(r1 & 0xff), e1 = (int)((f1 >> 8) & 0xff) - (int)((r1 >> 8) & 0xff);
            int e2 = (int)((f1 >> 16) & 0xff) - (int)((r1 >> 16) & 0xff), e3 = (int)(f1 >> 24) - (int)(r1 >> 24);
            int s01 = e0 + e1, d01 = e0 - e1, s23 = e2 + e3, d23 = e2 - e3;
            t1[0] = s01 + s23; t1[1] = s01 - s23; t1[2] = d01 + d23; t1[3] = d01 - d23;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) { s[half * 4 + k] = t0[k] + t1[k]; d[half * 4 + k] = t0[k] - t1[k]; }
    }
    // vertical second stage across the lane pair (xor 1); values fit int16 -> pack 2 per shuffle
    int sum = 0;
    const bool odd = m.rp & 1;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int mine = (s[k] & 0xffff) | (d[k] << 16);
        const int other = __shfl_xor_sync(0xffffffffu, mine, 1);
        const int so = (int)(short)(other & 0xffff), dd = other >> 16;
        sum += odd ? abs(so - s[k]) + abs(dd - d[k]) : abs(s[k] + so) + abs(d[k] + dd);
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);     // 8x4 total in both lanes of the pair
    sum >>= 1;
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);     // + the other 8x4
    return sum;
}

// full-pel candidate on the (possibly weighted) plane 0: SAD + mv cost
__device__ __forceinline__ int cand_fpel(const WarpMb &m, int mx, int my, bool active)
{
    int c = 0;
    if (active) {
        const uint8_t *p = m.fref_w + m.pel + (my + 2 * m.rp) * m.stride + mx;
        c = rows_sad(m, load8u(p), load8u(p + m.stride));
    }
    c = group_sum(c);
    return active ? c + mvcost(m, mx * 4, my * 4) : BIG_COST;
}
// quarter-pel candidate through get_ref: SAD or SATD + mv cost
__device__ __forceinline__ int cand_qpel(const WarpMb &m, int qx, int qy, bool active, bool use_satd)
{
    uint2 a0 = make_uint2(0, 0), a1 = a0;
    if (active) {
        a0 = get_ref_row(m.fref, m.stride, m.pel, qx, qy, 2 * m.rp, m.w);
        a1 = get_ref_row(m.fref, m.stride, m.pel, qx, qy, 2 * m.rp + 1, m.w);
    }
    int c;
    if (use_satd) c = rows_satd(m, a0, a1);
    else c = group_sum(rows_sad(m, a0, a1));
    return active ? c + mvcost(m, qx, qy) : BIG_COST;
}

__constant__ const signed char c_hex2[8][2] = {{-1, -2}, {-2, 0}, {-1, 2}, {1, 2}, {2, 0}, {1, -2}, {-1, -2}, {-2, 0}};
__constant__ const unsigned char c_mod6m1[8] = {5, 0, 1, 2, 3, 4, 5, 0};
__constant__ const signed char c_square1[9][2] = {{0, 0}, {0, -1}, {0, 1}, {-1, 0}, {1, 0}, {-1, -1}, {-1, 1}, {1, -1}, {1, 1}};
// first hexagon in upstream's evaluation order, with the direction code it is packed with
__constant__ const signed char c_hex_first[6][3] = {{-2, 0, 2}, {-1, 2, 3}, {1, 2, 4}, {2, 0, 5}, {1, -2, 6}, {-1, -2, 7}};

struct MeResult { int mvx, mvy, cost; };

__device__ MeResult me_search_mb(WarpMb &m, const MeParams &P, const int mvc[4][2], int i_mvc,
                                 int min_sx, int max_sx, int min_sy, int max_sy)
{
    const int mv_x_min = min_sx >> 2, mv_x_max = max_sx >> 2, mv_y_min = min_sy >> 2, mv_y_max = max_sy >> 2;
    const int grp = m.grp;
    int bmx, bmy, bcost, bpred_cost = LA_COST_MAX, bpred_mx = 0, bpred_my = 0;
    int pm_fx = 0, pm_fy = 0;

    if (P.subpel_refine >= 3) {
        const int pmx = clip3i(m.mvp_x, mv_x_min * 4, mv_x_max * 4), pmy = clip3i(m.mvp_y, mv_y_min * 4, mv_y_max * 4);
        // slot 0 = clipped mvp, slots 1..n = surviving clipped candidates (x264_predictor_clip)
        int cx = pmx, cy = pmy, n = 0;
        bool active = grp == 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (i < i_mvc) {
                const int mx = mvc[i][0], my = mvc[i][1];
                if ((mx | my) && !(mx == pmx && my == pmy)) {
                    n++;
                    if (grp == n) { cx = clip3i(mx, mv_x_min * 4, mv_x_max * 4); cy = clip3i(my, mv_y_min * 4, mv_y_max * 4); active = true; }
                }
            }
        }
        const int c = cand_qpel(m, cx, cy, active, false);
        const int packed = warp_min(active ? (c << 4) + grp : 0x7fffffff);
        const int pmv_cost = __shfl_sync(0xffffffffu, c, 0);
        const int best = packed & 15;
        bpred_cost = packed >> 4;
        bpred_mx = __shfl_sync(0xffffffffu, cx, best * 4);
        bpred_my = __shfl_sync(0xffffffffu, cy, best * 4);
        bmx = (bpred_mx + 2) >> 2; bmy = (bpred_my + 2) >> 2;
        const bool subpel = ((bpred_mx | bpred_my) & 3) != 0;
        const bool pmv_nz = (pmx | pmy) != 0;
        const bool need_zero = pmv_nz && (bmx | bmy);
        // slot 0: rounded best predictor (only if it was sub-pel), slot 1: the zero vector
        int fc = BIG_COST;
        if (subpel || need_zero) {
            const bool a = (grp == 0 && subpel) || (grp == 1 && need_zero);
            fc = cand_fpel(m, grp == 0 ? bmx : 0, grp == 0 ? bmy : 0, a);
        }
        const int c_round = __shfl_sync(0xffffffffu, fc, 0), c_zero = __shfl_sync(0xffffffffu, fc, 4);
        bcost = subpel ? c_round : bpred_cost;
        if (pmv_nz) { if (need_zero && c_zero < bcost) { bcost = c_zero; bmx = 0; bmy = 0; } }
        else if (pmv_cost < bcost) { bcost = pmv_cost; bmx = 0; bmy = 0; }
    } else {
        // subme < 3: full-pel predictors; the rounded mvp is scored without its mv cost
        bmx = pm_fx = clip3i((m.mvp_x + 2) >> 2, mv_x_min, mv_x_max);
        bmy = pm_fy = clip3i((m.mvp_y + 2) >> 2, mv_y_min, mv_y_max);
        int cx = bmx, cy = bmy, n = 0;
        bool active = grp == 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (i < i_mvc) {
                const int mx = (mvc[i][0] + 2) >> 2, my = (mvc[i][1] + 2) >> 2;
                if ((mx | my) && !(mx == pm_fx && my == pm_fy)) {
                    n++;
                    if (grp == n) { cx = clip3i(mx, mv_x_min, mv_x_max); cy = clip3i(my, mv_y_min, mv_y_max); active = true; }
                }
            }
        }
        const bool pmv_nz = (pm_fx | pm_fy) != 0;
        const bool zslot = pmv_nz && grp == 7;           // the zero vector rides along in slot 7
        if (zslot) { cx = 0; cy = 0; active = true; }
        int c = cand_fpel(m, cx, cy, active);
        if (grp == 0) c -= mvcost(m, cx * 4, cy * 4);
        const int c_zero = __shfl_sync(0xffffffffu, c, 28);
        const int packed = warp_min((active && !zslot) ? (c << 4) + grp : 0x7fffffff);
        const int best = packed & 15;
        bcost = packed >> 4;
        bmx = __shfl_sync(0xffffffffu, cx, best * 4);
        bmy = __shfl_sync(0xffffffffu, cy, best * 4);
        if (pmv_nz && c_zero < bcost) { bcost = c_zero; bmx = 0; bmy = 0; }
    }

    if (!P.me_hex) {
        // X264_ME_DIA
        int i = P.me_range;
        while (true) {
            const int dx = grp == 2 ? -1 : grp == 3 ? 1 : 0, dy = grp == 0 ? -1 : grp == 1 ? 1 : 0;
            const int c = cand_fpel(m, bmx + dx, bmy + dy, grp < 4);
            const int packed = warp_min(grp < 4 ? (c << 4) + grp + 1 : 0x7fffffff);
            if ((packed >> 4) >= bcost) break;
            bcost = packed >> 4;
            const int k = (packed & 15) - 1;
            bmx += k == 2 ? -1 : k == 3 ? 1 : 0;
            bmy += k == 0 ? -1 : k == 1 ? 1 : 0;
            if (!(--i && bmx >= mv_x_min && bmx <= mv_x_max && bmy >= mv_y_min && bmy <= mv_y_max)) break;
        }
    } else {
        // X264_ME_HEX: hexagon radius 2 ...
        {
            const int k = min(grp, 5);
            const int c = cand_fpel(m, bmx + c_hex_first[k][0], bmy + c_hex_first[k][1], grp < 6);
            const int packed = warp_min(grp < 6 ? (c << 3) + c_hex_first[k][2] : 0x7fffffff);
            if (packed < (bcost << 3)) {
                bcost = packed >> 3;
                int dir = (packed & 7) - 2;
                bmx += c_hex2[dir + 1][0]; bmy += c_hex2[dir + 1][1];
                // ... half hexagons while improving and in range
                for (int i = (P.me_range >> 1) - 1; i > 0 && bmx >= mv_x_min && bmx <= mv_x_max && bmy >= mv_y_min && bmy <= mv_y_max; i--) {
                    const int kk = min(grp, 2);
                    const int cc = cand_fpel(m, bmx + c_hex2[dir + kk][0], bmy + c_hex2[dir + kk][1], grp < 3);
                    const int pk = warp_min(grp < 3 ? (cc << 3) + kk + 1 : 0x7fffffff);
                    if (pk >= (bcost << 3)) break;
                    bcost = pk >> 3;
                    dir += (pk & 7) - 2;
                    dir = c_mod6m1[dir + 1];
                    bmx += c_hex2[dir + 1][0]; bmy += c_hex2[dir + 1][1];
                }
            }
        }
        // ... then square refine
        {
            const int c = cand_fpel(m, bmx + c_square1[grp + 1][0], bmy + c_square1[grp + 1][1], true);
            const int packed = warp_min((c << 4) + grp + 1);
            if (packed < (bcost << 4)) {
                bcost = packed >> 4;
                bmx += c_square1[packed & 15][0]; bmy += c_square1[packed & 15][1];
            }
        }
    }

    int mvx, mvy, cost;
    if (P.subpel_refine < 3) {
        cost = bcost;
        if (bmx == pm_fx && bmy == pm_fy) cost += mvcost(m, bmx * 4, bmy * 4);
        mvx = bmx * 4; mvy = bmy * 4;
    } else if (bpred_cost < bcost) { mvx = bpred_mx; mvy = bpred_my; cost = bpred_cost; }
    else { mvx = bmx * 4; mvy = bmy * 4; cost = bcost; }

    // ---- refine_subpel: hpel_iters = 1; qpel_iters = 1 for subme 4, 0 for subme 2 ----
    bmx = mvx; bmy = mvy; bcost = cost;
    if (P.subpel_refine < 3) {
        const int mx = clip3i(m.mvp_x, min_sx + 2, max_sx - 2), my = clip3i(m.mvp_y, min_sy + 2, max_sy - 2);
        if ((mx - bmx) | (my - bmy)) {
            const int c = __shfl_sync(0xffffffffu, cand_qpel(m, mx, my, grp == 0, false), 0);
            if (c < bcost) { bcost = c; bmx = mx; bmy = my; }
        }
    }
    {
        const int dx = grp == 2 ? -2 : grp == 3 ? 2 : 0, dy = grp == 0 ? -2 : grp == 1 ? 2 : 0;
        const int c = cand_qpel(m, bmx + dx, bmy + dy, grp < 4, false);
        const int packed = warp_min(grp < 4 ? (c << 4) + grp + 1 : 0x7fffffff);
        if ((packed >> 4) < bcost) {
            bcost = packed >> 4;
            const int k = (packed & 15) - 1;
            bmx += k == 2 ? -2 : k == 3 ? 2 : 0;
            bmy += k == 0 ? -2 : k == 1 ? 2 : 0;
        }
    }
    if (P.satd) {
        // slot 0 re-scores the half-pel winner with SATD; slots 1-4 are the quarter-pel diamond
        const bool do_qpel = P.subpel_refine >= 4 && !(bmy <= min_sy || bmy >= max_sy || bmx <= min_sx || bmx >= max_sx);
        const int dx = grp == 3 ? -1 : grp == 4 ? 1 : 0, dy = grp == 1 ? -1 : grp == 2 ? 1 : 0;
        const bool active = grp == 0 || (do_qpel && grp < 5);
        const int c = cand_qpel(m, bmx + dx, bmy + dy, active, true);
        bcost = __shfl_sync(0xffffffffu, c, 0);
        if (do_qpel) {
            const int packed = warp_min((grp >= 1 && grp < 5) ? (c << 4) + grp : 0x7fffffff);
            if ((packed >> 4) < bcost) {
                bcost = packed >> 4;
                const int k = packed & 15;
                bmx += k == 3 ? -1 : k == 4 ? 1 : 0;
                bmy += k == 1 ? -1 : k == 2 ? 1 : 0;
            }
        }
    }
    MeResult r = {bmx, bmy, bcost};
    return r;
}

__global__ void __launch_bounds__(32)
me_wavefront_kernel(LaGeom g, MeParams P)
{
    const MeJob &job = P.job[blockIdx.y];
    const int lane = threadIdx.x;
    int ticket = 0;
    if (lane == 0) ticket = atomicAdd(job.sync, 1);
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
    const int mb_y = g.mb_h - 1 - ticket;                 // bottom rows start first
    if (mb_y < 0) return;
    const int T = max(1, P.bands);
    int slice_start = 0, slice_end = g.mb_h;
    for (int i = 0; i < T; i++) {
        const int s = (g.mb_h * i + T / 2) / T, e = (g.mb_h * (i + 1) + T / 2) / T;
        if (mb_y >= s && mb_y < e) { slice_start = s; slice_end = e; }
    }
    int *progress = job.sync + 1;
    const int start_y = min(slice_end - 1, g.mb_h - 2 + P.do_edges), end_y = max(slice_start, 1 - P.do_edges);
    const int start_x = g.mb_w - 2 + P.do_edges, end_x = 1 - P.do_edges;
    if (mb_y > start_y || mb_y < end_y) {                  // row not scanned (edges without do_edges)
        if (lane == 0) st_release(progress + mb_y, -1);
        return;
    }
    const bool has_below = mb_y < slice_end - 1;
    const bool below_scanned = has_below && (mb_y + 1 <= start_y);

    WarpMb m;
    m.lane = lane; m.grp = lane >> 2; m.rp = lane & 3;
    m.stride = g.lstride;
#pragma unroll
    for (int k = 0; k < 4; k++) m.fref[k] = job.fref[k];
    m.fref_w = job.fref_w; m.w = job.w; m.cost_mv = P.cost_mv; m.satd = P.satd;

    int right_mv = 0;          // packed mv of (x+1, y); zero before the first MB like the zeroed array
    for (int mb_x = start_x; mb_x >= end_x; mb_x--) {
        const int mb_xy = mb_x + mb_y * g.mb_w;
        m.pel = 8 * (mb_x + mb_y * g.lstride);
        {
            const uint8_t *f = job.fenc + m.pel + 2 * m.rp * g.lstride;
            m.fe0 = load8u(f); m.fe1 = load8u(f + g.lstride);
        }
        // ---- reverse-order MV prediction ----
        int mvc[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
        int i_mvc = 0;
        if (mb_x < g.mb_w - 1) { mvc[i_mvc][0] = mv_x(right_mv); mvc[i_mvc][1] = mv_y(right_mv); i_mvc++; }
        if (has_below) {
            if (below_scanned) {
                const int need = max(mb_x - 1, end_x);     // row below must have finished this x
                while (ld_acquire(progress + mb_y + 1) > need) __nanosleep(64);
            }
            const int *below = job.mvs + mb_xy + g.mb_w;
            int v = ld_relaxed(below);
            mvc[i_mvc][0] = mv_x(v); mvc[i_mvc][1] = mv_y(v); i_mvc++;
            if (mb_x > 0) { v = ld_relaxed(below - 1); mvc[i_mvc][0] = mv_x(v); mvc[i_mvc][1] = mv_y(v); i_mvc++; }
            if (mb_x < g.mb_w - 1) { v = ld_relaxed(below + 1); mvc[i_mvc][0] = mv_x(v); mvc[i_mvc][1] = mv_y(v); i_mvc++; }
        }
        if (i_mvc <= 1) { m.mvp_x = mvc[0][0]; m.mvp_y = mvc[0][1]; }
        else { m.mvp_x = median3i(mvc[0][0], mvc[1][0], mvc[2][0]); m.mvp_y = median3i(mvc[0][1], mvc[1][1], mvc[2][1]); }

        int min_sx, max_sx, min_sy, max_sy;
        mv_limits(mb_x, mb_y, g.mb_w, g.mb_h, P.mv_range2, min_sx, max_sx, min_sy, max_sy);

        int out_mv = 0, out_cost = 0;
        bool skip = false;
        if (!(m.mvp_x | m.mvp_y)) {
            // fast skip: mbcmp at mv 0 on the UNWEIGHTED plane 0
            const uint8_t *p = job.fref[0] + m.pel + 2 * m.rp * g.lstride;
            const uint2 a0 = load8u(p), a1 = load8u(p + g.lstride);
            const int c = P.satd ? rows_satd(m, a0, a1) : group_sum(rows_sad(m, a0, a1));
            const int c0 = __shfl_sync(0xffffffffu, c, 0);
            if (c0 < 64) { skip = true; out_mv = 0; out_cost = c0; }
        }
        if (!skip) {
            MeResult r = me_search_mb(m, P, mvc, i_mvc, min_sx, max_sx, min_sy, max_sy);
            int cost = r.cost - (int)__ldg(P.cost_mv);      // remove mvcost from skip mbs
            if (r.mvx | r.mvy) cost += 5;
            out_mv = mv_pack(r.mvx, r.mvy); out_cost = cost;
        }
        right_mv = out_mv;
        if (lane == 0) {
            job.mvs[mb_xy] = out_mv;
            job.mv_costs[mb_xy] = out_cost;
            st_release(progress + mb_y, mb_x);
        }
    }
}

int launch_me(cudaStream_t st, const LaGeom &g, const MeParams &p)
{
    if (p.njobs <= 0) return 0;
    dim3 grid(g.mb_h, p.njobs);
    me_wavefront_kernel<<<grid, 32, 0, st>>>(g, p);
    XV_LAUNCH_CHECK();
    return 0;
}

} // namespace xv
