// Lowres motion search, PLAIN WAVEFRONT: the whole search in dependency order, one warp per MB row.
// This was the round's first design and is now the fallback (X264VFW_CUDA_ME_VARIANT=0); sessions
// use the speculative passes + verification wavefront of la_me2_kernel.cu, which
// produce the same MVs and costs bit for bit.
//
// Restates, for one (frame, reference, list) pair, the search that [x264]
// encoder/slicetype.c: slicetype_mb_cost runs per 8x8 lowres MB in reverse raster order:
//   median MV predictor from the 4 already-visited neighbours (right, below, below-left,
//   below-right), the mv0 SATD<64 early skip, [x264] encoder/me.c: x264_me_search_ref
//   (sub-pel or full-pel predictor candidates, zero check, hexagon + square refine or
//   diamond, on SAD) and refine_subpel (half-pel diamond on SAD, SATD re-score, quarter-pel
//   diamond on SATD).  Ties break exactly as upstream (strict <, candidate order).
//
// Mapping to the GPU
//   * MB (x,y) depends on (x+1,y) and on row y+1 up to x-1, so every row is a 2-MB-skewed
//     pipeline stage.  One warp owns one MB row and walks x downwards.  Rows are handed out by
//     an atomic ticket (bottom row first), so a waiting warp only ever waits on warps that
//     already started -- no co-residency assumption, no deadlock.
//   * Rows talk through one 8-byte record per MB {mv, epoch}: a single 64-bit store/load is
//     atomic, so no fence and no separate progress flag.  The one NEW record an MB needs
//     (below-left of the next MB) is prefetched while the current MB is being searched, which
//     takes the L2 round trip off the critical path; so are the next MB's source pixels.
//   * Inside an MB the warp evaluates up to 8 candidate positions at once: 4 lanes per
//     candidate, 2 block rows per lane, __vsadu4 for SAD and shuffle butterflies for the 4x4
//     Hadamard of SATD -- integer pipe only (this is not a dense contraction).
//   * Reference pixels are staged in shared memory: a 64x48 window of the (weighted) plane 0
//     around the predictor for all full-pel rounds (hexagon/square/diamond), and a 20x12
//     window of all four half-pel phase planes around the full-pel winner for the sub-pel
//     rounds.  A candidate that leaves the full-pel window falls back to global loads, so the
//     staging never changes a result.
//   * Several searches (both lists of a B evaluation, or all searches a newly arrived frame
//     enables) share one launch via blockIdx.y.
#include "la_common.cuh"

namespace xv {

#define BIG_COST 0x3fffffff
#define WIN_W 64
#define WIN_H 48
#define WIN_R 20
#define SUB_W 20
#define SUB_H 12

struct __align__(16) MeSmem {
    uint8_t win[WIN_H * WIN_W];
    uint8_t sub[4][SUB_H * SUB_W];
};

__device__ __forceinline__ uint2 ld_rec(const int2 *p)
{
    uint2 v;
    asm volatile("ld.relaxed.gpu.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_rec(int2 *p, int mv, int epoch)
{
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1,%2};" :: "l"(p), "r"(mv), "r"(epoch) : "memory");
}

struct WarpMb {
    uint2 fe0, fe1;              // fenc rows 2*rp and 2*rp+1
    int rp, grp;
    int stride, pel;
    int px, py;                  // plane coordinates of the MB origin
    const uint8_t *fref0; int plane_stride;      // four phase planes, plane_stride bytes apart
    const uint8_t *fref_w;
    WeightDev w;
    const uint16_t *cost_mv;
    int mvp_x, mvp_y;
    int satd;
    // shared-memory staging
    const uint8_t *win; int wx0, wy0;           // full-pel window origin, plane coordinates
    const uint8_t *sub; int sx0, sy0;           // sub-pel windows origin, plane coordinates
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

// 8 bytes at an arbitrary byte offset of a shared-memory window (same 3-word scheme as load8u)
__device__ __forceinline__ uint2 lds8u(const uint8_t *base, int off)
{
    const uint32_t *q = (const uint32_t *)(base + (off & ~3));
    const unsigned sh = (unsigned)(off & 3) * 8;
    const uint32_t w0 = q[0], w1 = q[1], w2 = q[2];
    return make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
}

// full-pel candidate on the (possibly weighted) plane 0: SAD + mv cost
__device__ __forceinline__ int cand_fpel(const WarpMb &m, int mx, int my, bool active)
{
    int c = 0, mvc = 0;
    if (active) {
        mvc = mvcost(m, mx * 4, my * 4);
        const int X = m.px + mx, Y = m.py + my + 2 * m.rp;
        const int wx = X - m.wx0, wy = m.py + my - m.wy0;
        uint2 a0, a1;
        if (m.win && wx >= 0 && wx + 12 <= WIN_W && wy >= 0 && wy + 8 <= WIN_H) {
            const int off = (Y - m.wy0) * WIN_W + wx;
            a0 = lds8u(m.win, off); a1 = lds8u(m.win, off + WIN_W);
        } else {
            const uint8_t *p = m.fref_w + m.pel + (my + 2 * m.rp) * m.stride + mx;
            a0 = load8u(p); a1 = load8u(p + m.stride);
        }
        c = rows_sad(m, a0, a1);
    }
    c = group_sum(c);
    return active ? c + mvc : BIG_COST;
}

// One row (8 px) of get_ref out of the staged sub-pel windows (must cover the position)
__device__ __forceinline__ uint2 get_ref_row_sub(const WarpMb &m, int mvx, int mvy, int r)
{
    const int qidx = ((mvy & 3) << 2) + (mvx & 3);
    const int X = m.px + (mvx >> 2) - m.sx0, Y = m.py + (mvy >> 2) + r - m.sy0;
    uint2 a = lds8u(m.sub + c_hpel_ref0[qidx] * (SUB_H * SUB_W), (Y + ((mvy & 3) == 3)) * SUB_W + X);
    if (qidx & 5) {
        const uint2 b = lds8u(m.sub + c_hpel_ref1[qidx] * (SUB_H * SUB_W), Y * SUB_W + X + ((mvx & 3) == 3));
        a.x = avg4(a.x, b.x);
        a.y = avg4(a.y, b.y);
    }
    if (m.w.on) { a.x = weight_word(m.w, a.x); a.y = weight_word(m.w, a.y); }
    return a;
}

// quarter-pel candidate through get_ref: SAD or SATD + mv cost
template <bool SUBWIN>
__device__ __forceinline__ int cand_qpel(const WarpMb &m, int qx, int qy, bool active, bool use_satd)
{
    uint2 a0 = make_uint2(0, 0), a1 = a0;
    int mvc = 0;
    if (active) {
        mvc = mvcost(m, qx, qy);
        if (SUBWIN) {
            a0 = get_ref_row_sub(m, qx, qy, 2 * m.rp);
            a1 = get_ref_row_sub(m, qx, qy, 2 * m.rp + 1);
        } else {
            a0 = get_ref_row_ps(m.fref0, m.plane_stride, m.stride, m.pel, qx, qy, 2 * m.rp, m.w);
            a1 = get_ref_row_ps(m.fref0, m.plane_stride, m.stride, m.pel, qx, qy, 2 * m.rp + 1, m.w);
        }
    }
    int c;
    if (use_satd) c = rows_satd(m, a0, a1);
    else c = group_sum(rows_sad(m, a0, a1));
    return active ? c + mvc : BIG_COST;
}

__constant__ const signed char c_hex2[8][2] = {{-1, -2}, {-2, 0}, {-1, 2}, {1, 2}, {2, 0}, {1, -2}, {-1, -2}, {-2, 0}};
__constant__ const unsigned char c_mod6m1[8] = {5, 0, 1, 2, 3, 4, 5, 0};
__constant__ const signed char c_square1[9][2] = {{0, 0}, {0, -1}, {0, 1}, {-1, 0}, {1, 0}, {-1, -1}, {-1, 1}, {1, -1}, {1, 1}};
// first hexagon in upstream's evaluation order, with the direction code it is packed with
__constant__ const signed char c_hex_first[6][3] = {{-2, 0, 2}, {-1, 2, 3}, {1, 2, 4}, {2, 0, 5}, {1, -2, 6}, {-1, -2, 7}};

struct MeResult { int mvx, mvy, cost; };

// Asynchronous global->shared copies (LDGSTS): no staging registers, no separate store
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.wait_all;" ::: "memory");
}

// stage the 64x48 full-pel window of fref_w: 6 x 16-byte async copies per lane, issued early
__device__ __forceinline__ void win_issue(const WarpMb &m, MeSmem &sm, const LaGeom &g, int lane, int cx, int cy, int &wx0, int &wy0)
{
    wx0 = (m.px + cx - WIN_R) & ~15;
    wx0 = min(max(wx0, -32), g.lstride - 32 - WIN_W);
    wy0 = min(max(m.py + cy - WIN_R, -32), g.lh + 32 - WIN_H);
    const uint8_t *base = m.fref_w + wy0 * m.stride + wx0;
    __syncwarp();                                        // previous readers of the window are done
#pragma unroll
    for (int k = 0; k < 6; k++) {
        const int c = lane + 32 * k;
        cp_async16(sm.win + (c >> 2) * WIN_W + (c & 3) * 16, base + (c >> 2) * m.stride + (c & 3) * 16);
    }
}
__device__ __forceinline__ void win_commit()
{
    cp_async_wait_all();
    __syncwarp();
}
// stage the four 20x12 sub-pel windows around full-pel position (fx,fy) (MB-relative)
__device__ __forceinline__ void sub_load(WarpMb &m, MeSmem &sm, int lane, int fx, int fy)
{
    m.sx0 = (m.px + fx - 1) & ~3;
    m.sy0 = m.py + fy - 1;
    __syncwarp();
    // 4 planes x 12 rows x 5 words = 240 words: lane -> plane (lane>>3), 8 lanes x 8 rounds cover 60 words
    const int pl = lane >> 3, l8 = lane & 7;
    const uint8_t *src = m.fref0 + (size_t)pl * m.plane_stride + m.sy0 * m.stride + m.sx0;
    uint8_t *dst = sm.sub[pl];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int idx = l8 + 8 * k;                      // 0..63, valid below 60
        if (idx < SUB_H * (SUB_W / 4)) {
            const int row = idx / (SUB_W / 4), wd = idx - row * (SUB_W / 4);
            cp_async4(dst + row * SUB_W + 4 * wd, src + row * m.stride + 4 * wd);
        }
    }
    cp_async_wait_all();
    __syncwarp();
    m.sub = &sm.sub[0][0];
}

__device__ __forceinline__ MeResult me_search_mb(WarpMb &m, MeSmem &sm, const LaGeom &g, const MeParams &P, const int mvc[4][2], int i_mvc,
                                 int min_sx, int max_sx, int min_sy, int max_sy, int lane)
{
    const int mv_x_min = min_sx >> 2, mv_x_max = max_sx >> 2, mv_y_min = min_sy >> 2, mv_y_max = max_sy >> 2;
    const int grp = m.grp;
    int bmx, bmy, bcost, bpred_cost = LA_COST_MAX, bpred_mx = 0, bpred_my = 0;
    int pm_fx = 0, pm_fy = 0;
    m.win = nullptr;

    if (P.subpel_refine >= 3) {
        const int pmx = clip3i(m.mvp_x, mv_x_min * 4, mv_x_max * 4), pmy = clip3i(m.mvp_y, mv_y_min * 4, mv_y_max * 4);
        int wx0, wy0;
        win_issue(m, sm, g, lane, (pmx + 2) >> 2, (pmy + 2) >> 2, wx0, wy0);
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
        const int c = cand_qpel<false>(m, cx, cy, active, false);
        win_commit();
        m.win = sm.win; m.wx0 = wx0; m.wy0 = wy0;
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
        int wx0, wy0;
        win_issue(m, sm, g, lane, bmx, bmy, wx0, wy0);
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
        win_commit();
        m.win = sm.win; m.wx0 = wx0; m.wy0 = wy0;
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
            const int c = __shfl_sync(0xffffffffu, cand_qpel<false>(m, mx, my, grp == 0, false), 0);
            if (c < bcost) { bcost = c; bmx = mx; bmy = my; }
        }
    }
    sub_load(m, sm, lane, bmx >> 2, bmy >> 2);
    {
        const int dx = grp == 2 ? -2 : grp == 3 ? 2 : 0, dy = grp == 0 ? -2 : grp == 1 ? 2 : 0;
        const int c = cand_qpel<true>(m, bmx + dx, bmy + dy, grp < 4, false);
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
        const int c = cand_qpel<true>(m, bmx + dx, bmy + dy, active, true);
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

#define ME_WARPS 2      // independent rows per block (lifts the 32-blocks-per-SM residency cap)
__global__ void __launch_bounds__(32 * ME_WARPS, 32 / ME_WARPS)
me_wavefront_kernel(LaGeom g, MeParams P)
{
    __shared__ MeSmem sm_all[ME_WARPS];
    MeSmem &sm = sm_all[threadIdx.x >> 5];
    const MeJob &job = P.job[blockIdx.y];
    const int lane = threadIdx.x & 31;
    // Rows are handed out bottom-first by an atomic ticket.  A warp that finishes its row takes
    // the next ticket, so a search keeps only P.rows_in_flight warps resident instead of one per
    // row: rows near the top could not start for ~2*(mb_h-1-y) MB steps anyway and would only
    // hold warp slots while waiting.
  for (;;) {
    int ticket = 0;
    if (lane == 0) ticket = atomicAdd(job.ticket, 1);
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
    const int mb_y = g.mb_h - 1 - ticket;                 // bottom rows start first
    if (mb_y < 0) return;
    const int T = max(1, P.bands);
    int slice_start = 0, slice_end = g.mb_h;
    for (int i = 0; i < T; i++) {
        const int s = (g.mb_h * i + T / 2) / T, e = (g.mb_h * (i + 1) + T / 2) / T;
        if (mb_y >= s && mb_y < e) { slice_start = s; slice_end = e; }
    }
    const int start_y = min(slice_end - 1, g.mb_h - 2 + P.do_edges), end_y = max(slice_start, 1 - P.do_edges);
    const int start_x = g.mb_w - 2 + P.do_edges, end_x = 1 - P.do_edges;
    if (mb_y > start_y || mb_y < end_y) continue;          // row not scanned (edges without do_edges)
    const bool has_below = mb_y < slice_end - 1;
    const bool below_scanned = has_below && (mb_y + 1 <= start_y);
    const int epoch = P.epoch;
    const int2 *below = job.rec + (mb_y + 1) * g.mb_w;
    int2 *mine = job.rec + mb_y * g.mb_w;

    WarpMb m;
    m.grp = lane >> 2; m.rp = lane & 3;
    m.stride = g.lstride;
    m.fref0 = job.fref[0]; m.plane_stride = g.lplane;      // the phase planes of a frame are contiguous
    m.fref_w = job.fref_w; m.w = job.w; m.cost_mv = P.cost_mv; m.satd = P.satd;
    m.win = nullptr; m.sub = nullptr;

    // MV of a row-below MB: zero when that position is never scanned, else wait for its record
    auto below_mv = [&](int x, uint2 pre, bool have_pre) -> int {
        if (!below_scanned || x < end_x || x > start_x) return 0;
        uint2 r = have_pre ? pre : ld_rec(below + x);
        unsigned ns = 32;
        while ((int)r.y != epoch) { __nanosleep(ns); if (ns < 512) ns <<= 1; r = ld_rec(below + x); }
        return (int)r.x;
    };

    int right_mv = 0;          // packed mv of (x+1, y); zero before the first MB like the zeroed array
    int b_m1 = 0, b_0 = 0, b_p1 = 0;                        // below-left, below, below-right of the current MB
    if (has_below) {
        const uint2 z = make_uint2(0, 0);
        if (start_x + 1 < g.mb_w) b_p1 = below_mv(start_x + 1, z, false);
        b_0 = below_mv(start_x, z, false);
        if (start_x > 0) b_m1 = below_mv(start_x - 1, z, false);
    }
    // source pixels of the first MB
    uint2 nfe0, nfe1, nr0, nr1;
    {
        const int pel = 8 * (start_x + mb_y * g.lstride) + 2 * m.rp * g.lstride;
        nfe0 = load8u(job.fenc + pel); nfe1 = load8u(job.fenc + pel + g.lstride);
        nr0 = load8u(job.fref[0] + pel); nr1 = load8u(job.fref[0] + pel + g.lstride);
    }

    for (int mb_x = start_x; mb_x >= end_x; mb_x--) {
        const int mb_xy = mb_x + mb_y * g.mb_w;
        m.pel = 8 * (mb_x + mb_y * g.lstride);
        m.px = 8 * mb_x; m.py = 8 * mb_y;
        m.fe0 = nfe0; m.fe1 = nfe1;
        const uint2 r0 = nr0, r1 = nr1;
        // ---- prefetch for the next MB: its pixels, and the one new record it needs ----
        uint2 pre = make_uint2(0, 0);
        const bool want_pre = has_below && below_scanned && mb_x - 2 >= end_x;
        if (mb_x > end_x) {
            const int pel = m.pel - 8 + 2 * m.rp * g.lstride;
            nfe0 = load8u(job.fenc + pel); nfe1 = load8u(job.fenc + pel + g.lstride);
            nr0 = load8u(job.fref[0] + pel); nr1 = load8u(job.fref[0] + pel + g.lstride);
            if (want_pre) pre = ld_rec(below + mb_x - 2);
        }
        // ---- reverse-order MV prediction: candidates in upstream's order (right, below,
        // below-left, below-right), compacted without dynamic indexing ----
        const bool has_r = mb_x < g.mb_w - 1, has_bl = has_below && mb_x > 0, has_br = has_below && has_r;
        int c0, c1, c2, c3, i_mvc;
        if (has_below) {
            c0 = has_r ? right_mv : b_0;
            c1 = has_r ? b_0 : (has_bl ? b_m1 : 0);
            c2 = has_r ? (has_bl ? b_m1 : b_p1) : 0;
            c3 = (has_r && has_bl) ? b_p1 : 0;
            i_mvc = (int)has_r + 1 + (int)has_bl + (int)has_br;
        } else {
            c0 = has_r ? right_mv : 0; c1 = c2 = c3 = 0;
            i_mvc = (int)has_r;
        }
        const int mvc[4][2] = {{mv_x(c0), mv_y(c0)}, {mv_x(c1), mv_y(c1)}, {mv_x(c2), mv_y(c2)}, {mv_x(c3), mv_y(c3)}};
        if (i_mvc <= 1) { m.mvp_x = mvc[0][0]; m.mvp_y = mvc[0][1]; }
        else { m.mvp_x = median3i(mvc[0][0], mvc[1][0], mvc[2][0]); m.mvp_y = median3i(mvc[0][1], mvc[1][1], mvc[2][1]); }

        int min_sx, max_sx, min_sy, max_sy;
        mv_limits(mb_x, mb_y, g.mb_w, g.mb_h, P.mv_range2, min_sx, max_sx, min_sy, max_sy);

        int out_mv = 0, out_cost = 0;
        bool skip = false;
        if (!(m.mvp_x | m.mvp_y)) {
            // fast skip: mbcmp at mv 0 on the UNWEIGHTED plane 0
            const int c = P.satd ? rows_satd(m, r0, r1) : group_sum(rows_sad(m, r0, r1));
            const int c0 = __shfl_sync(0xffffffffu, c, 0);
            if (c0 < 64) { skip = true; out_mv = 0; out_cost = c0; }
        }
        if (!skip) {
            MeResult r = me_search_mb(m, sm, g, P, mvc, i_mvc, min_sx, max_sx, min_sy, max_sy, lane);
            int cost = r.cost - (int)__ldg(P.cost_mv);      // remove mvcost from skip mbs
            if (r.mvx | r.mvy) cost += 5;
            out_mv = mv_pack(r.mvx, r.mvy); out_cost = cost;
        }
        right_mv = out_mv;
        if (lane == 0) {
            job.mvs[mb_xy] = out_mv;
            job.mv_costs[mb_xy] = out_cost;
            st_rec(mine + mb_x, out_mv, epoch);
        }
        // ---- slide the below-row window; resolve the prefetched record ----
        b_p1 = b_0; b_0 = b_m1;
        b_m1 = (has_below && mb_x - 2 >= 0) ? below_mv(mb_x - 2, pre, want_pre) : 0;
    }
  }
}

static dim3 wavefront_grid(const LaGeom &g, const MeParams &p)
{
    const int rows = p.rows_in_flight > 0 && p.rows_in_flight < g.mb_h ? p.rows_in_flight : g.mb_h;
    return dim3((rows + ME_WARPS - 1) / ME_WARPS, p.njobs);
}

// plain wavefront: the whole search in dependency order
int launch_me(cudaStream_t st, const LaGeom &g, const MeParams &p)
{
    if (p.njobs <= 0) return 0;
    me_wavefront_kernel<<<wavefront_grid(g, p), 32 * ME_WARPS, 0, st>>>(g, p);
    XV_LAUNCH_CHECK();
    return 0;
}

} // namespace xv
