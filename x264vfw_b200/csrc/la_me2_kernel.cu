// Lowres motion search, parallel part: speculative per-MB searches, several MBs per warp.
//
// The sequential definition ([x264] slicetype_mb_cost scanning MBs in reverse raster order,
// each MB predicted from its right / below / below-left / below-right neighbours) makes one
// search a chain of mb_w + 2(mb_h-1) dependent MB searches.  But an MB's result is a pure
// function of its pixels and of those four neighbour MVs, and motion fields are smooth: here
// every MB is searched AT ONCE from assumed neighbour MVs (the same search of the previous
// frame as a first guess, then the results of the previous pass), and a cheap verification
// wavefront (me_verify_kernel below) keeps a result only if the inputs it assumed
// are the final MVs of its neighbours -- otherwise it re-runs that MB's search in order.  The
// output is therefore exactly the sequential scan's (bit-exact incl. tie-breaks); the guess
// only decides how much of the work happens off the critical path.
//
// Mapping: a group of 8 lanes owns one MB and spends ONE lane per candidate position (a lane
// scores a whole 8x8 candidate: 8 rows of __vsadu4, or an in-register 8x8 SATD), so the
// control / addressing / reduction instruction stream is shared by the 4 MBs of a warp.
// The warp never de-converges: groups that are idle, skipped or done with a loop early walk
// through the same rounds with their candidates predicated off (see FULL below) -- per-group
// lane masks would let the groups run one after the other (measured: 10.8 active lanes per
// instruction), which defeats the purpose.
#include "la_common.cuh"

namespace xv {

#define BIG_COST 0x3fffffff
// Full-pel window staged in shared memory: the block itself + WIN_R pixels around the search centre, x origin
// aligned down to 16 bytes (cp.async chunks): 2 * WIN_R + 8 + 15 <= WIN_W, 2 * WIN_R + 8 <= WIN_H.  Candidates that
// leave the window (the hexagon may walk up to me_range + 1 = 17 pixels) are scored from global memory instead --
// same result, rarer than 1 in 100 on the bench clip -- so the window radius is a pure occupancy knob.
#ifndef WIN_R
#define WIN_R 12
#endif
#define WIN_W ((2 * WIN_R + 8 + 15 + 15) & ~15)
#define WIN_H (2 * WIN_R + 8)
#ifndef PASS_BLOCKS_PER_SM
#define PASS_BLOCKS_PER_SM 12
#endif
#define SUB_W 32
#define SUB_H 12

// The sub-pel windows reuse the full-pel window's storage (the full-pel rounds are over by
// then): 2.5 KB per MB in flight, so shared memory allows 20 warps per SM.
struct __align__(16) GroupSmem {
    union {
        uint8_t win[WIN_H * WIN_W];       // full-pel window of the (weighted) plane 0
        uint8_t sub[4][SUB_H * SUB_W];    // sub-pel windows of the four phase planes
    };
};

__device__ __forceinline__ uint2 ld_rec2(const int2 *p)
{
    uint2 v;
    asm volatile("ld.relaxed.gpu.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_rec2(int2 *p, int mv, int epoch)
{
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1,%2};" :: "l"(p), "r"(mv), "r"(epoch) : "memory");
}
__device__ __forceinline__ void cpa16(void *smem, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cpa8(void *smem, const void *gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cpa_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int LPS>
struct Mb {
    static constexpr int GL = 8 * LPS;      // lanes of a group
    static constexpr int RPL = 8 / LPS;     // block rows scored by one lane
    uint2 fe[RPL];                          // this lane's rows of the source MB
    int gl, slot, r0;                       // lane within group, candidate slot, first block row
    int stride, pel, px, py;
    const uint8_t *fref0; int plane_stride;
    const uint8_t *fref_w;
    WeightDev w;
    const uint16_t *cost_mv;
    int mvp_x, mvp_y;
    const uint8_t *win; int wx0, wy0;
    const uint8_t *sub; int sx0, sy0;
};

#define FULL 0xffffffffu
// Every shuffle / warp barrier below is executed by the WHOLE warp at a warp-uniform point;
// what differs per group is data and predicates only.  (Per-group lane masks would let the
// groups run de-converged, i.e. one after the other -- measured: 10.8 active lanes per
// instruction -- which defeats the purpose of sharing the instruction stream.)
template <int LPS> __device__ __forceinline__ int g_bcast(int v, int slot)
{
    return __shfl_sync(FULL, v, slot * LPS, 8 * LPS);
}
template <int LPS> __device__ __forceinline__ int g_min(int v)
{
    v = min(v, __shfl_xor_sync(FULL, v, LPS));
    v = min(v, __shfl_xor_sync(FULL, v, 2 * LPS));
    v = min(v, __shfl_xor_sync(FULL, v, 4 * LPS));
    return v;
}
template <int LPS> __device__ __forceinline__ int part_sum(int v)
{
    if (LPS >= 2) v += __shfl_xor_sync(FULL, v, 1);
    if (LPS >= 4) v += __shfl_xor_sync(FULL, v, 2);
    return v;
}
template <int LPS> __device__ __forceinline__ int mvcost2(const Mb<LPS> &m, int qx, int qy)
{
    return (int)__ldg(m.cost_mv + (qx - m.mvp_x)) + (int)__ldg(m.cost_mv + (qy - m.mvp_y));
}

// ---- SATD of this lane's rows ([x264] x264_pixel_satd_8x4 per 4 rows: two 4x4 Hadamards) ----
__device__ __forceinline__ int satd_8x4_regs(const uint2 *a, const uint2 *b)
{
    // dp4a row transform + max-folded column transform (la_common.cuh)
    return satd4x4_half(a[0].x, a[1].x, a[2].x, a[3].x, b[0].x, b[1].x, b[2].x, b[3].x) +
           satd4x4_half(a[0].y, a[1].y, a[2].y, a[3].y, b[0].y, b[1].y, b[2].y, b[3].y);
}
template <int LPS> __device__ __forceinline__ int satd_rows(const Mb<LPS> &m, const uint2 *a)
{
    if (LPS == 4) return satd_rows4(m.fe[0], m.fe[1 % Mb<LPS>::RPL], a[0], a[1 % Mb<LPS>::RPL], m.gl & 1);
    if (LPS == 1) {
        // two 8x4 halves through ONE copy of the butterfly code (rows 4-7 are moved down for
        // the second trip): code size matters more than 16 register moves here
        uint2 f[4], r[4];
#pragma unroll
        for (int y = 0; y < 4; y++) { f[y] = m.fe[y]; r[y] = a[y]; }
        int sum = 0;
#pragma unroll 1
        for (int half = 0; half < 2; half++) {
            sum += satd_8x4_regs(f, r);
#pragma unroll
            for (int y = 0; y < 4; y++) { f[y] = m.fe[4 * (LPS == 1) + y]; r[y] = a[4 * (LPS == 1) + y]; }
        }
        return sum;
    }
    return part_sum<LPS>(satd_8x4_regs(m.fe, a));
}
template <int LPS> __device__ __forceinline__ int sad_rows(const Mb<LPS> &m, const uint2 *a)
{
    int c = 0;
#pragma unroll
    for (int r = 0; r < Mb<LPS>::RPL; r++) c += __vsadu4(a[r].x, m.fe[r].x) + __vsadu4(a[r].y, m.fe[r].y);
    return c;
}

// weighted references are rare (fades): keep their per-pixel arithmetic out of line
__device__ __noinline__ uint32_t weight_word_call(int scale, int denom, int offset, uint32_t v)
{
    const WeightDev w = {1, scale, denom, offset};
    return weight_word(w, v);
}

// ---- candidates ----------------------------------------------------------------------------
// full-pel candidate on the (possibly weighted) plane 0: SAD + mv cost
template <int LPS>
__device__ __forceinline__ int cand_fpel(const Mb<LPS> &m, int mx, int my, bool active)
{
    constexpr int RPL = Mb<LPS>::RPL;
    int c = 0, mvc = 0;
    if (active) {
        mvc = mvcost2(m, mx * 4, my * 4);
        const int wx = m.px + mx - m.wx0, wy = m.py + my - m.wy0;
        if (m.win && wx >= 0 && wx + 12 <= WIN_W && wy >= 0 && wy + 8 <= WIN_H) {
            const int off = (wy + m.r0) * WIN_W + wx;
            const uint32_t *q = (const uint32_t *)(m.win + (off & ~3));
            const unsigned sh = (unsigned)(off & 3) * 8;
#pragma unroll
            for (int r = 0; r < RPL; r++) {
                const uint32_t w0 = q[0], w1 = q[1], w2 = q[2];
                c += __vsadu4(__funnelshift_r(w0, w1, sh), m.fe[r].x) + __vsadu4(__funnelshift_r(w1, w2, sh), m.fe[r].y);
                q += WIN_W / 4;
            }
        } else {
            const uint8_t *p = m.fref_w + m.pel + (my + m.r0) * m.stride + mx;
#pragma unroll
            for (int r = 0; r < RPL; r++) {
                const uint2 a = load8u(p);
                c += __vsadu4(a.x, m.fe[r].x) + __vsadu4(a.y, m.fe[r].y);
                p += m.stride;
            }
        }
    }
    c = part_sum<LPS>(c);
    return active ? c + mvc : BIG_COST;
}

// quarter-pel candidate through get_ref: SAD or SATD + mv cost
template <int LPS, bool SUBWIN>
__device__ __forceinline__ int cand_qpel(const Mb<LPS> &m, int qx, int qy, bool active, bool use_satd)
{
    constexpr int RPL = Mb<LPS>::RPL;
    uint2 a[RPL];
#pragma unroll
    for (int r = 0; r < RPL; r++) a[r] = make_uint2(0, 0);
    int mvc = 0;
    if (active) {
        mvc = mvcost2(m, qx, qy);
        const int qidx = ((qy & 3) << 2) + (qx & 3);
        const bool two = (qidx & 5) != 0;
        if (SUBWIN) {
            const int X = m.px + (qx >> 2) - m.sx0, Y = m.py + (qy >> 2) + m.r0 - m.sy0;
            const uint8_t *p0 = m.sub + c_hpel_ref0[qidx] * (SUB_H * SUB_W);
            const uint8_t *p1 = m.sub + c_hpel_ref1[qidx] * (SUB_H * SUB_W);
            const int o0 = (Y + ((qy & 3) == 3)) * SUB_W + X, o1 = Y * SUB_W + X + ((qx & 3) == 3);
            const uint32_t *q0 = (const uint32_t *)(p0 + (o0 & ~3)), *q1 = (const uint32_t *)(p1 + (o1 & ~3));
            const unsigned s0 = (unsigned)(o0 & 3) * 8, s1 = (unsigned)(o1 & 3) * 8;
#pragma unroll
            for (int r = 0; r < RPL; r++) {
                uint2 v = make_uint2(__funnelshift_r(q0[0], q0[1], s0), __funnelshift_r(q0[1], q0[2], s0));
                if (two) {
                    v.x = avg4(v.x, __funnelshift_r(q1[0], q1[1], s1));
                    v.y = avg4(v.y, __funnelshift_r(q1[1], q1[2], s1));
                }
                a[r] = v;
                q0 += SUB_W / 4; q1 += SUB_W / 4;
            }
        } else {
            const int off = m.pel + ((qy >> 2) + m.r0) * m.stride + (qx >> 2);
            const uint8_t *p0 = m.fref0 + (size_t)c_hpel_ref0[qidx] * m.plane_stride + off + ((qy & 3) == 3) * m.stride;
            const uint8_t *p1 = m.fref0 + (size_t)c_hpel_ref1[qidx] * m.plane_stride + off + ((qx & 3) == 3);
#pragma unroll
            for (int r = 0; r < RPL; r++) {
                uint2 v = load8u(p0);
                if (two) {
                    const uint2 b = load8u(p1);
                    v.x = avg4(v.x, b.x); v.y = avg4(v.y, b.y);
                }
                a[r] = v;
                p0 += m.stride; p1 += m.stride;
            }
        }
        if (m.w.on) {
#pragma unroll
            for (int r = 0; r < RPL; r++) { a[r].x = weight_word_call(m.w.scale, m.w.denom, m.w.offset, a[r].x); a[r].y = weight_word_call(m.w.scale, m.w.denom, m.w.offset, a[r].y); }
        }
    }
    int c;
    if (use_satd) c = satd_rows(m, a);
    else c = part_sum<LPS>(sad_rows(m, a));
    return active ? c + mvc : BIG_COST;
}

__constant__ const signed char c2_hex2[8][2] = {{-1, -2}, {-2, 0}, {-1, 2}, {1, 2}, {2, 0}, {1, -2}, {-1, -2}, {-2, 0}};
__constant__ const unsigned char c2_mod6m1[8] = {5, 0, 1, 2, 3, 4, 5, 0};
__constant__ const signed char c2_square1[9][2] = {{0, 0}, {0, -1}, {0, 1}, {-1, 0}, {1, 0}, {-1, -1}, {-1, 1}, {1, -1}, {1, 1}};
__constant__ const signed char c2_hex_first[6][3] = {{-2, 0, 2}, {-1, 2, 3}, {1, 2, 4}, {2, 0, 5}, {1, -2, 6}, {-1, -2, 7}};

struct MeResult2 { int mvx, mvy, cost; };

// stage the 64x40 full-pel window of fref_w around (cx,cy) (MB-relative full-pel): 160 chunks of 16 B
template <int LPS>
__device__ __forceinline__ void win_issue(const Mb<LPS> &m, GroupSmem &sm, const LaGeom &g, bool on, int cx, int cy, int &wx0, int &wy0)
{
    constexpr int GL = Mb<LPS>::GL;
    wx0 = (m.px + cx - WIN_R) & ~15;
    wx0 = min(max(wx0, -32), g.lstride - 32 - WIN_W);
    wy0 = min(max(m.py + cy - WIN_R, -32), g.lh + 32 - WIN_H);
    const uint8_t *base = m.fref_w + wy0 * m.stride + wx0;
    __syncwarp();                                        // previous readers of the window are done
    if (on) {
        // the window is WIN_H rows of WIN_W / 16 chunks of 16 bytes; lane gl copies chunks gl, gl + GL, ...
        constexpr int CPR = WIN_W / 16, NCH = WIN_H * CPR;
#pragma unroll
        for (int k = 0; k < (NCH + GL - 1) / GL; k++) {
            const int c = m.gl + k * GL;
            if (NCH % GL == 0 || c < NCH) {
                const int row = c / CPR, col = c - row * CPR;
                cpa16(sm.win + row * WIN_W + col * 16, base + row * m.stride + col * 16);
            }
        }
    }
}
__device__ __forceinline__ void win_commit()
{
    cpa_wait_all();
    __syncwarp();
}
// stage the four 32x12 sub-pel windows around full-pel position (fx,fy) (MB-relative): 96 chunks of 16 B
template <int LPS>
__device__ __forceinline__ void sub_load(Mb<LPS> &m, GroupSmem &sm, bool on, int fx, int fy)
{
    constexpr int GL = Mb<LPS>::GL;
    m.sx0 = (m.px + fx - 1) & ~15;
    m.sy0 = m.py + fy - 1;
    m.win = nullptr;                                     // its storage is about to be overwritten
    __syncwarp();
    if (on) {
        // per plane 24 chunks = 12 rows x 2 halves: lane gl copies (row, half) = (gl/2 + (GL/2) j, gl%2)
        const uint8_t *src = m.fref0 + m.sy0 * m.stride + m.sx0 + (m.gl >> 1) * m.stride + (m.gl & 1) * 16;
        uint8_t *dst = &sm.sub[0][0] + (m.gl >> 1) * SUB_W + (m.gl & 1) * 16;
        const int sstep = (GL / 2) * m.stride;
#pragma unroll
        for (int pl = 0; pl < 4; pl++) {
            const uint8_t *s2 = src; uint8_t *d2 = dst;
#pragma unroll
            for (int j = 0; j < (24 + GL - 1) / GL; j++) {
                if (GL * j + m.gl < 24) cpa16(d2, s2);
                s2 += sstep; d2 += (GL / 2) * SUB_W;
            }
            src += m.plane_stride; dst += SUB_H * SUB_W;
        }
    }
    cpa_wait_all();
    __syncwarp();
    m.sub = &sm.sub[0][0];
}

// `on`: this group's MB is being searched (group-uniform).
//
// The full-pel part is ONE loop around ONE candidate evaluation: every group carries its own
// state (which round of x264_me_search_ref it is in) and supplies that round's candidates, so
// groups need not be in the same round -- a group that leaves the hexagon early goes on to
// the square refine while its neighbours iterate -- and the code stays small (the first
// version inlined the evaluation at seven call sites: 129 KB of SASS and 23 % of the stall
// samples waiting for instructions).  The sub-pel part is one loop of two rounds.
enum { S_PRED, S_RZ, S_HEX1, S_HEXIT, S_SQUARE, S_DIA, S_DONE };

// n_sad / n_satd: 8x8 block metrics evaluated by the warp (counted only when P.stats is set; every lane holds the
// same totals, the caller adds them once per warp)
template <int LPS> __device__ __forceinline__ int count_cands(const Mb<LPS> &m, bool a)
{
    return __popc(__ballot_sync(FULL, a && (m.gl % LPS) == 0));
}
template <int LPS, bool QPRED>
__device__ __forceinline__ MeResult2 me_search_mb2(Mb<LPS> &m, GroupSmem &sm, const LaGeom &g, const MeParams &P, const bool on,
                                                   const int mvc[4][2], int i_mvc, int min_sx, int max_sx, int min_sy, int max_sy,
                                                   int &n_sad, int &n_satd)
{
    const int mv_x_min = min_sx >> 2, mv_x_max = max_sx >> 2, mv_y_min = min_sy >> 2, mv_y_max = max_sy >> 2;
    const int grp = m.slot;
    int bmx = 0, bmy = 0, bcost = 0, bpred_cost = LA_COST_MAX, bpred_mx = 0, bpred_my = 0;
    int pm_fx = 0, pm_fy = 0;
    int st = S_DONE;
    // S_PRED (subme < 3): this lane's full-pel predictor candidate
    int pcx = 0, pcy = 0; bool pact = false, pzslot = false;
    // S_RZ (subme >= 3): what the predictor round found
    bool subpel = false, pmv_nz = false, need_zero = false; int pmv_cost = 0;
    m.win = nullptr;

    if (QPRED) {
        const int pmx = clip3i(m.mvp_x, mv_x_min * 4, mv_x_max * 4), pmy = clip3i(m.mvp_y, mv_y_min * 4, mv_y_max * 4);
        int wx0, wy0;
        win_issue(m, sm, g, on, (pmx + 2) >> 2, (pmy + 2) >> 2, wx0, wy0);
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
        active = active && on;
        if (P.stats) n_sad += count_cands(m, active);
        const int c = cand_qpel<LPS, false>(m, cx, cy, active, false);
        win_commit();
        m.win = sm.win; m.wx0 = wx0; m.wy0 = wy0;
        const int packed = g_min<LPS>(active ? (c << 4) + grp : 0x7fffffff);
        pmv_cost = g_bcast<LPS>(c, 0);
        const int best = packed & 7;
        bpred_cost = packed >> 4;
        bpred_mx = g_bcast<LPS>(cx, best);
        bpred_my = g_bcast<LPS>(cy, best);
        bmx = (bpred_mx + 2) >> 2; bmy = (bpred_my + 2) >> 2;
        subpel = ((bpred_mx | bpred_my) & 3) != 0;
        pmv_nz = (pmx | pmy) != 0;
        need_zero = pmv_nz && (bmx | bmy);
        if (on) {
            if (subpel || need_zero) st = S_RZ;              // slot 0: rounded best predictor, slot 1: the zero vector
            else {
                bcost = bpred_cost;
                if (!pmv_nz && pmv_cost < bcost) { bcost = pmv_cost; bmx = 0; bmy = 0; }
                st = P.me_hex ? S_HEX1 : S_DIA;
            }
        }
    } else {
        // subme < 3: full-pel predictors; the rounded mvp is scored without its mv cost
        bmx = pm_fx = clip3i((m.mvp_x + 2) >> 2, mv_x_min, mv_x_max);
        bmy = pm_fy = clip3i((m.mvp_y + 2) >> 2, mv_y_min, mv_y_max);
        int wx0, wy0;
        win_issue(m, sm, g, on, bmx, bmy, wx0, wy0);
        pcx = bmx; pcy = bmy;
        int n = 0;
        pact = grp == 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (i < i_mvc) {
                const int mx = (mvc[i][0] + 2) >> 2, my = (mvc[i][1] + 2) >> 2;
                if ((mx | my) && !(mx == pm_fx && my == pm_fy)) {
                    n++;
                    if (grp == n) { pcx = clip3i(mx, mv_x_min, mv_x_max); pcy = clip3i(my, mv_y_min, mv_y_max); pact = true; }
                }
            }
        }
        pmv_nz = (pm_fx | pm_fy) != 0;
        pzslot = pmv_nz && grp == 7;                     // the zero vector rides along in slot 7
        if (pzslot) { pcx = 0; pcy = 0; pact = true; }
        win_commit();
        m.win = sm.win; m.wx0 = wx0; m.wy0 = wy0;
        if (on) st = S_PRED;
    }

    // ---- full-pel rounds: predictors (subme < 3), rounded predictor / zero (subme >= 3),
    //      hexagon + half hexagons + square refine (X264_ME_HEX) or diamond (X264_ME_DIA) ----
    int dir = 0, iters = P.me_range;
    while (__any_sync(FULL, st != S_DONE)) {
        int mx = 0, my = 0, tag = 0; bool a = false;
        if (st == S_HEX1) { const int k = min(grp, 5); mx = bmx + c2_hex_first[k][0]; my = bmy + c2_hex_first[k][1]; tag = c2_hex_first[k][2]; a = grp < 6; }
        else if (st == S_HEXIT) { const int kk = min(grp, 2); mx = bmx + c2_hex2[dir + kk][0]; my = bmy + c2_hex2[dir + kk][1]; tag = kk + 1; a = grp < 3; }
        else if (st == S_SQUARE) { mx = bmx + c2_square1[grp + 1][0]; my = bmy + c2_square1[grp + 1][1]; tag = grp + 1; a = true; }
        else if (st == S_DIA) { mx = bmx + (grp == 2 ? -1 : grp == 3 ? 1 : 0); my = bmy + (grp == 0 ? -1 : grp == 1 ? 1 : 0); tag = grp + 1; a = grp < 4; }
        else if (QPRED && st == S_RZ) { mx = grp == 0 ? bmx : 0; my = grp == 0 ? bmy : 0; a = (grp == 0 && subpel) || (grp == 1 && need_zero); }
        else if (!QPRED && st == S_PRED) { mx = pcx; my = pcy; tag = grp; a = pact; }
        if (P.stats) n_sad += count_cands(m, a);
        int c = cand_fpel(m, mx, my, a);
        if (!QPRED && st == S_PRED) {
            if (grp == 0) c -= mvcost2(m, mx * 4, my * 4);
            a = a && !pzslot;
        }
        const int packed = g_min<LPS>(a ? (c << 4) + tag : 0x7fffffff);
        const int c_s0 = g_bcast<LPS>(c, 0), c_s1 = g_bcast<LPS>(c, 1);
        int c_s7 = 0, px = 0, py = 0;
        if (!QPRED) { c_s7 = g_bcast<LPS>(c, 7); px = g_bcast<LPS>(pcx, packed & 7); py = g_bcast<LPS>(pcy, packed & 7); }
        if (st == S_HEX1) {
            st = S_SQUARE;
            if (packed < (bcost << 4)) {
                bcost = packed >> 4;
                dir = (packed & 15) - 2;
                bmx += c2_hex2[dir + 1][0]; bmy += c2_hex2[dir + 1][1];
                iters = (P.me_range >> 1) - 1;
                if (iters > 0 && bmx >= mv_x_min && bmx <= mv_x_max && bmy >= mv_y_min && bmy <= mv_y_max) st = S_HEXIT;
            }
        } else if (st == S_HEXIT) {
            if (packed >= (bcost << 4)) st = S_SQUARE;
            else {
                bcost = packed >> 4;
                dir += (packed & 15) - 2;
                dir = c2_mod6m1[dir + 1];
                bmx += c2_hex2[dir + 1][0]; bmy += c2_hex2[dir + 1][1];
                iters--;
                if (!(iters > 0 && bmx >= mv_x_min && bmx <= mv_x_max && bmy >= mv_y_min && bmy <= mv_y_max)) st = S_SQUARE;
            }
        } else if (st == S_SQUARE) {
            if (packed < (bcost << 4)) {
                bcost = packed >> 4;
                bmx += c2_square1[packed & 15][0]; bmy += c2_square1[packed & 15][1];
            }
            st = S_DONE;
        } else if (st == S_DIA) {
            if ((packed >> 4) >= bcost) st = S_DONE;
            else {
                bcost = packed >> 4;
                const int k = (packed & 15) - 1;
                bmx += k == 2 ? -1 : k == 3 ? 1 : 0;
                bmy += k == 0 ? -1 : k == 1 ? 1 : 0;
                if (!(--iters && bmx >= mv_x_min && bmx <= mv_x_max && bmy >= mv_y_min && bmy <= mv_y_max)) st = S_DONE;
            }
        } else if (QPRED && st == S_RZ) {
            bcost = subpel ? c_s0 : bpred_cost;
            if (pmv_nz) { if (need_zero && c_s1 < bcost) { bcost = c_s1; bmx = 0; bmy = 0; } }
            else if (pmv_cost < bcost) { bcost = pmv_cost; bmx = 0; bmy = 0; }
            st = P.me_hex ? S_HEX1 : S_DIA;
            iters = P.me_range;
        } else if (!QPRED && st == S_PRED) {
            bcost = packed >> 4;
            bmx = px; bmy = py;
            if (pmv_nz && c_s7 < bcost) { bcost = c_s7; bmx = 0; bmy = 0; }
            st = P.me_hex ? S_HEX1 : S_DIA;
            iters = P.me_range;
        }
    }

    int mvx, mvy, cost;
    if (!QPRED) {
        cost = bcost;
        if (on && bmx == pm_fx && bmy == pm_fy) cost += mvcost2(m, bmx * 4, bmy * 4);
        mvx = bmx * 4; mvy = bmy * 4;
    } else if (bpred_cost < bcost) { mvx = bpred_mx; mvy = bpred_my; cost = bpred_cost; }
    else { mvx = bmx * 4; mvy = bmy * 4; cost = bcost; }
    if (!on) { mvx = 0; mvy = 0; cost = 0; }

    // ---- refine_subpel: hpel_iters = 1; qpel_iters = 1 for subme 4, 0 for subme 2 ----
    bmx = mvx; bmy = mvy; bcost = cost;
    if (!QPRED) {
        const int mx = clip3i(m.mvp_x, min_sx + 2, max_sx - 2), my = clip3i(m.mvp_y, min_sy + 2, max_sy - 2);
        const bool want = on && (((mx - bmx) | (my - bmy)) != 0);
        if (__any_sync(FULL, want)) {
            if (P.stats) n_sad += count_cands(m, want && grp == 0);
            const int c = g_bcast<LPS>(cand_qpel<LPS, false>(m, mx, my, want && grp == 0, false), 0);
            if (want && c < bcost) { bcost = c; bmx = mx; bmy = my; }
        }
    }
    sub_load(m, sm, on, bmx >> 2, bmy >> 2);
    // round 0: half-pel diamond on SAD (slots 0-3); round 1 (mbcmp = SATD only): slot 0 re-scores
    // the half-pel winner with SATD, slots 1-4 are the quarter-pel diamond
    const int nrounds = P.satd ? 2 : 1;
    for (int round = 0; round < nrounds; round++) {
        int dx, dy; bool a, do_qpel = false;
        if (round == 0) {
            dx = grp == 2 ? -2 : grp == 3 ? 2 : 0; dy = grp == 0 ? -2 : grp == 1 ? 2 : 0;
            a = on && grp < 4;
        } else {
            do_qpel = on && P.subpel_refine >= 4 && !(bmy <= min_sy || bmy >= max_sy || bmx <= min_sx || bmx >= max_sx);
            dx = grp == 3 ? -1 : grp == 4 ? 1 : 0; dy = grp == 1 ? -1 : grp == 2 ? 1 : 0;
            a = on && (grp == 0 || (do_qpel && grp < 5));
        }
        if (P.stats) { if (round == 1) n_satd += count_cands(m, a); else n_sad += count_cands(m, a); }
        const int c = cand_qpel<LPS, true>(m, bmx + dx, bmy + dy, a, round == 1);
        const int c_s0 = g_bcast<LPS>(c, 0);
        if (round == 0) {
            const int packed = g_min<LPS>(a ? (c << 4) + grp + 1 : 0x7fffffff);
            if (on && (packed >> 4) < bcost) {
                bcost = packed >> 4;
                const int k = (packed & 15) - 1;
                bmx += k == 2 ? -2 : k == 3 ? 2 : 0;
                bmy += k == 0 ? -2 : k == 1 ? 2 : 0;
            }
        } else {
            const int packed = g_min<LPS>((do_qpel && grp >= 1 && grp < 5) ? (c << 4) + grp : 0x7fffffff);
            if (on) bcost = c_s0;
            if (do_qpel && (packed >> 4) < bcost) {
                bcost = packed >> 4;
                const int k = packed & 15;
                bmx += k == 3 ? -1 : k == 4 ? 1 : 0;
                bmy += k == 1 ? -1 : k == 2 ? 1 : 0;
            }
        }
    }
    MeResult2 r = {bmx, bmy, bcost};
    return r;
}

// ------------------------------------------------------------------------------------------
// Parallel pass: every MB of the frame is searched at once from ASSUMED neighbour MVs (a guess
// field for the first pass, the previous pass's results afterwards).  A group of 8*LPS lanes
// owns one MB, NG horizontally adjacent MBs share a warp.  Each MB records the four inputs it
// used next to its result; the verification wavefront (me_verify_kernel below)
// accepts a result only if those inputs equal the final MVs of the neighbours, and re-runs the
// search otherwise, so the output is exactly the sequential reverse-raster scan's.
//   pass 0      : inputs from job.guess (or zeros), all MBs
//   pass 1, 2.. : inputs from the current results; only MBs whose inputs changed are re-run
// ------------------------------------------------------------------------------------------
#define PASS_WARPS 2
// a guess field of another temporal distance, rescaled (any value is a legal guess)
__device__ __forceinline__ int scale_mv(int mv, int num, int den, int lim)
{
    const float f = (float)num / (float)den;
    const int sx = __float2int_rn((float)mv_x(mv) * f), sy = __float2int_rn((float)mv_y(mv) * f);
    return mv_pack(clip3i(sx, -lim, lim - 1), clip3i(sy, -lim, lim - 1));     // stay inside the mv cost table
}
// where an MB sits in the scan: which neighbours exist, whether it is scanned at all
struct MbPos { int mb_x, mb_y, mb_xy; bool act, has_below, has_r, has_bl, has_br; };
__device__ __forceinline__ MbPos mb_pos(const LaGeom &g, const MeParams &P, int mb_x, int mb_y, bool in_range)
{
    const int T = max(1, P.bands);
    const int start_x = g.mb_w - 2 + P.do_edges, end_x = 1 - P.do_edges;
    int slice_start = 0, slice_end = g.mb_h;
#pragma unroll 1
    for (int i = 0; i < T; i++) {
        const int s = (g.mb_h * i + T / 2) / T, e = (g.mb_h * (i + 1) + T / 2) / T;
        if (mb_y >= s && mb_y < e) { slice_start = s; slice_end = e; }
    }
    const int start_y = min(slice_end - 1, g.mb_h - 2 + P.do_edges), end_y = max(slice_start, 1 - P.do_edges);
    MbPos q;
    q.mb_x = mb_x; q.mb_y = mb_y; q.mb_xy = mb_x + mb_y * g.mb_w;
    q.act = in_range && mb_y <= start_y && mb_y >= end_y && mb_x <= start_x && mb_x >= end_x;
    q.has_below = mb_y < slice_end - 1;
    q.has_r = mb_x < g.mb_w - 1; q.has_bl = q.has_below && mb_x > 0; q.has_br = q.has_below && q.has_r;
    return q;
}
// the four inputs (right, below, below-left, below-right; zero where the neighbour does not exist)
__device__ __forceinline__ int4 mb_inputs(const LaGeom &g, const MbPos &q, const int *src)
{
    int4 in = make_int4(0, 0, 0, 0);
    if (q.act && src) {
        if (q.has_r) in.x = __ldcg(src + q.mb_xy + 1);
        if (q.has_below) in.y = __ldcg(src + q.mb_xy + g.mb_w);
        if (q.has_bl) in.z = __ldcg(src + q.mb_xy + g.mb_w - 1);
        if (q.has_br) in.w = __ldcg(src + q.mb_xy + g.mb_w + 1);
    }
    return in;
}

// One MB per group, searched from the inputs `in` (warp-uniform call; `need` is per group).
template <int LPS, bool QPRED>
__device__ __forceinline__ void pass_search_store(const LaGeom &g, const MeParams &P, const MeJob &job, GroupSmem &sm, int lane,
                                                  const MbPos &q, const int4 in, const bool need, int pass)
{
    constexpr int GL = Mb<LPS>::GL, RPL = Mb<LPS>::RPL;
    const int mb_x = q.mb_x, mb_y = q.mb_y;
    Mb<LPS> m;
    m.gl = lane % GL; m.slot = m.gl / LPS; m.r0 = (m.gl % LPS) * RPL;
    m.stride = g.lstride;
    m.fref0 = job.fref[0]; m.plane_stride = g.lplane;
    m.fref_w = job.fref_w; m.w = job.w; m.cost_mv = P.cost_mv;
    m.win = nullptr; m.sub = nullptr;
    m.wx0 = m.wy0 = m.sx0 = m.sy0 = 0;
    m.pel = 8 * (mb_x + mb_y * g.lstride);
    m.px = 8 * mb_x; m.py = 8 * mb_y;
#pragma unroll
    for (int r = 0; r < RPL; r++) m.fe[r] = make_uint2(0, 0);
    if (need) {
#pragma unroll
        for (int r = 0; r < RPL; r++) m.fe[r] = __ldg((const uint2 *)(job.fenc + m.pel + (m.r0 + r) * g.lstride));
    }

    // ---- reverse-order MV prediction, as in the sequential scan ----
    int c0, c1, c2, c3, i_mvc;
    if (q.has_below) {
        c0 = q.has_r ? in.x : in.y;
        c1 = q.has_r ? in.y : (q.has_bl ? in.z : 0);
        c2 = q.has_r ? (q.has_bl ? in.z : in.w) : 0;
        c3 = (q.has_r && q.has_bl) ? in.w : 0;
        i_mvc = (int)q.has_r + 1 + (int)q.has_bl + (int)q.has_br;
    } else {
        c0 = q.has_r ? in.x : 0; c1 = c2 = c3 = 0;
        i_mvc = (int)q.has_r;
    }
    if (!need) { c0 = c1 = c2 = c3 = 0; i_mvc = 0; }
    const int mvc[4][2] = {{mv_x(c0), mv_y(c0)}, {mv_x(c1), mv_y(c1)}, {mv_x(c2), mv_y(c2)}, {mv_x(c3), mv_y(c3)}};
    if (i_mvc <= 1) { m.mvp_x = mvc[0][0]; m.mvp_y = mvc[0][1]; }
    else { m.mvp_x = median3i(mvc[0][0], mvc[1][0], mvc[2][0]); m.mvp_y = median3i(mvc[0][1], mvc[1][1], mvc[2][1]); }

    int min_sx, max_sx, min_sy, max_sy;
    mv_limits(mb_x, mb_y, g.mb_w, g.mb_h, P.mv_range2, min_sx, max_sx, min_sy, max_sy);

    int out_mv = 0, out_cost = 0;
    int n_sad = 0, n_satd = 0;
    bool skip = false;
    const bool ztest = need && !(m.mvp_x | m.mvp_y);
    if (__any_sync(FULL, ztest)) {
        if (P.stats) { const int n = __popc(__ballot_sync(FULL, ztest && m.gl == 0)); if (P.satd) n_satd += n; else n_sad += n; }
        // fast skip: mbcmp at mv 0 on the UNWEIGHTED plane 0 (every lane scores its rows)
        uint2 a[RPL];
#pragma unroll
        for (int r = 0; r < RPL; r++) a[r] = make_uint2(0, 0);
        if (ztest) {
#pragma unroll
            for (int r = 0; r < RPL; r++) a[r] = __ldg((const uint2 *)(job.fref[0] + m.pel + (m.r0 + r) * g.lstride));
        }
        const int c = P.satd ? satd_rows(m, a) : part_sum<LPS>(sad_rows(m, a));
        if (ztest && c < 64) { skip = true; out_mv = 0; out_cost = c; }
    }
    const bool on = need && !skip;
    if (__any_sync(FULL, on)) {
        MeResult2 r = me_search_mb2<LPS, QPRED>(m, sm, g, P, on, mvc, i_mvc, min_sx, max_sx, min_sy, max_sy, n_sad, n_satd);
        if (on) {
            int cost = r.cost - (int)__ldg(P.cost_mv);      // remove mvcost from skip mbs
            if (r.mvx | r.mvy) cost += 5;
            out_mv = mv_pack(r.mvx, r.mvy); out_cost = cost;
        }
    }
    if (need && m.gl == 0) {
        job.mvs[q.mb_xy] = out_mv;
        job.mv_costs[q.mb_xy] = out_cost;
        job.assumed[q.mb_xy] = in;
    }
    if (P.stats) {     // one set of atomics per warp
        const int n_mb = __popc(__ballot_sync(FULL, need && m.gl == 0));
        if (lane == 0) { atomicAdd(P.stats + 2 + min(pass, 3), n_mb); atomicAdd(P.stats + 6, n_sad); atomicAdd(P.stats + 7, n_satd); }
    }
}

// 4 horizontally adjacent MBs per warp.  pass 0: every MB, inputs from the guess field; later
// passes: inputs from the current results, a warp leaves at once unless one of its MBs saw its
// inputs change.  (Collecting the MBs to re-search per 256-MB tile and searching them with a
// small grid was tried: same frame rate within noise, longer when they cluster, so not kept.)
template <int LPS, bool QPRED>
__global__ void __launch_bounds__(32 * PASS_WARPS, PASS_BLOCKS_PER_SM)
me_pass_kernel(LaGeom g, MeParams P, int pass)
{
    constexpr int GL = Mb<LPS>::GL, NG = 32 / GL;
    __shared__ GroupSmem sm_all[PASS_WARPS][NG];
    const MeJob &job = P.job[blockIdx.y];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gi = lane / GL;
    const int start_x = g.mb_w - 2 + P.do_edges, end_x = 1 - P.do_edges;
    const int ncols = start_x - end_x + 1;
    const int chunks = (ncols + NG - 1) / NG;
    const int wid = blockIdx.x * PASS_WARPS + warp;
    if (wid >= chunks * g.mb_h) return;
    const int mb_y = wid / chunks, k = (wid - mb_y * chunks) * NG + gi;
    const MbPos q = mb_pos(g, P, start_x - k, mb_y, k < ncols);
    int4 in = mb_inputs(g, q, pass == 0 ? job.guess : job.mvs);
    if (pass == 0 && job.guess_num != job.guess_den) {
        in.x = scale_mv(in.x, job.guess_num, job.guess_den, P.mv_range2); in.y = scale_mv(in.y, job.guess_num, job.guess_den, P.mv_range2);
        in.z = scale_mv(in.z, job.guess_num, job.guess_den, P.mv_range2); in.w = scale_mv(in.w, job.guess_num, job.guess_den, P.mv_range2);
    }
    bool need = q.act;
    if (pass > 0 && q.act) {
        const int4 a = __ldcg(job.assumed + q.mb_xy);
        need = a.x != in.x || a.y != in.y || a.z != in.z || a.w != in.w;
    }
    if (!__any_sync(FULL, need)) return;
    pass_search_store<LPS, QPRED>(g, P, job, sm_all[warp][gi], lane, q, in, need, pass);
}

int launch_me_pass(cudaStream_t st, const LaGeom &g, const MeParams &p, int pass)
{
    const int start_x = g.mb_w - 2 + p.do_edges, end_x = 1 - p.do_edges;
    const int ncols = start_x - end_x + 1;
    const int chunks = (ncols + 3) / 4;
    const int warps = chunks * g.mb_h;
    const dim3 grid((warps + PASS_WARPS - 1) / PASS_WARPS, p.njobs);
    if (p.subpel_refine >= 3) me_pass_kernel<1, true><<<grid, 32 * PASS_WARPS, 0, st>>>(g, p, pass);
    else me_pass_kernel<1, false><<<grid, 32 * PASS_WARPS, 0, st>>>(g, p, pass);
    XV_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------
// Verification wavefront: the exact, ordered half of the speculative search (see
// la_me2_kernel.cu).  Every MB already holds a result G and the four neighbour MVs A it was
// computed from.  Rows are walked in dependency order exactly like me_wavefront_kernel, but an
// MB whose assumed inputs equal the FINAL MVs of its neighbours is simply kept -- and runs of
// such MBs retire 32 at a time: lane i looks at MB x-i, polls the one new record it needs
// from the row below, takes its right neighbour's tentative G from lane i-1, and the leading
// run of matching lanes is final by induction.  Only a mismatching MB pays for a search (the
// whole warp, 8 candidates x 4 lanes: the same search code as the passes, instantiated with
// 4 lanes per candidate), with its true inputs.
// ------------------------------------------------------------------------------------------
// A block owns a BAND of VERIFY_ROWS consecutive rows, one warp per row: inside a band a row
// hands its MVs to the row above through shared memory (~100 cycles instead of an L2 round
// trip of ~1500 on the chain), only the top row of a band publishes to global memory for the
// band above.  Bands are handed out bottom-first by an atomic ticket, so a block only ever
// waits on bands that already started (no co-residency assumption between blocks).
#ifndef VERIFY_ROWS
#define VERIFY_ROWS 16
#endif
__device__ __forceinline__ uint2 lds_rec(const int2 *p)
{
    uint2 v;
    asm volatile("ld.volatile.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ void sts_rec(int2 *p, int mv, int epoch)
{
    asm volatile("st.volatile.shared.v2.u32 [%0], {%1,%2};" :: "r"((unsigned)__cvta_generic_to_shared(p)), "r"(mv), "r"(epoch) : "memory");
}

// Register budget: the band's 16 warps hold their registers for the whole walk although most of the time they only
// poll; at the compiler's own choice (123 registers) one band fills an SM's register file and nothing else -- no
// block of the parallel passes of another stream -- can run beside it.
#ifndef VERIFY_MAXREG
#define VERIFY_MAXREG 128
#endif
// RELAX = row relaxation: the same walk along every row, but WITHOUT waiting for the row below -- each row takes the
// row below as it currently is (the previous pass's results, possibly already updated by that row's own warp; whatever
// it reads is what it records as "assumed"), all rows run at once.  One such launch resolves every dependency chain
// that runs along a row (the right neighbour is the first predictor candidate, so most chains do), which a Jacobi
// pass advances by only one MB.  The exact, ordered launch afterwards finds far fewer MBs to search again.
template <bool QPRED, bool RELAX>
__global__ void __maxnreg__(VERIFY_MAXREG)
me_verify_kernel(LaGeom g, MeParams P, int relax_pass)
{
    extern __shared__ __align__(16) uint8_t vsm[];
    const int lane = threadIdx.x & 31, wrow = threadIdx.x >> 5;
    GroupSmem &sm = ((GroupSmem *)vsm)[wrow];
    int2 *srec = (int2 *)(vsm + VERIFY_ROWS * sizeof(GroupSmem));      // [VERIFY_ROWS][mb_w] {mv, epoch}
    __shared__ int s_ticket;
    const MeJob &job = P.job[blockIdx.y];
    const unsigned FULLM = 0xffffffffu;
    const int nbands = (g.mb_h + VERIFY_ROWS - 1) / VERIFY_ROWS;
  for (int round = 0;; round++) {
    if (RELAX && round) return;                            // relaxation: one band per block, no hand-over
    __syncthreads();                                       // everybody is done with the previous band's records
    if (!RELAX && threadIdx.x == 0) s_ticket = atomicAdd(job.ticket, 1);
    for (int i = lane; i < g.mb_w; i += 32) srec[wrow * g.mb_w + i] = make_int2(0, 0);   // epoch is never 0
    __syncthreads();
    const int ticket = RELAX ? (int)blockIdx.x : s_ticket;
    if (ticket >= nbands) return;
    const int mb_y = g.mb_h - 1 - (ticket * VERIFY_ROWS + wrow);   // bottom band first, warp 0 = its bottom row
    if (mb_y < 0) continue;
    const int T = max(1, P.bands);
    int slice_start = 0, slice_end = g.mb_h;
#pragma unroll 1
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
    const bool below_local = wrow > 0;                     // the row below belongs to this block
    const int2 *below = below_local ? srec + (wrow - 1) * g.mb_w : job.rec + (mb_y + 1) * g.mb_w;
    int2 *mine = job.rec + mb_y * g.mb_w;
    int2 *mine_s = srec + wrow * g.mb_w;
    const bool top = wrow == VERIFY_ROWS - 1;              // read by another block: publish to global memory too
    const int row0 = mb_y * g.mb_w;
    const int *below_mvs = job.mvs + (mb_y + 1) * g.mb_w;  // RELAX: the row below as it is right now
    auto ld_below = [&](int p) -> uint2 {
        if (RELAX) return make_uint2((unsigned)__ldcg(below_mvs + p), (unsigned)epoch);
        return below_local ? lds_rec(below + p) : ld_rec2(below + p);
    };
    auto publish = [&](int xcol, int mv) { if (RELAX) return; sts_rec(mine_s + xcol, mv, epoch); if (top) st_rec2(mine + xcol, mv, epoch); };

    Mb<4> m;                                               // the whole warp on one MB: 8 candidates x 4 lanes
    m.gl = lane; m.slot = lane >> 2; m.r0 = (lane & 3) * 2;
    m.stride = g.lstride;
    m.fref0 = job.fref[0]; m.plane_stride = g.lplane;
    m.fref_w = job.fref_w; m.w = job.w; m.cost_mv = P.cost_mv;
    m.win = nullptr; m.sub = nullptr;
    m.wx0 = m.wy0 = m.sx0 = m.sy0 = 0;

    // below-row value at column p that needs no record: position absent or never scanned
    auto below_fixed = [&](int p) -> bool { return !below_scanned || p < end_x || p > start_x || p < 0 || p >= g.mb_w; };
    auto below_wait = [&](int p) -> int {
        if (below_fixed(p)) return 0;
        uint2 r = ld_below(p);
        unsigned ns = 32;
        while ((int)r.y != epoch) { __nanosleep(ns); if (ns < 256) ns <<= 1; r = ld_below(p); }
        return (int)r.x;
    };

    int x = start_x;                // next MB of this row
    int right_mv = 0;               // final MV of (x+1, y); zero before the first MB
    int c0 = 0, c1 = 0;             // final MVs of the row below at columns x and x+1
    if (has_below) { c1 = below_wait(start_x + 1); c0 = below_wait(start_x); }
    int n_hit = 0, n_miss = 0, n_sad = 0, n_satd = 0;
    unsigned ns = 32;
    while (x >= end_x) {
        // ---- load a chunk: lane i examines MB xi = x - i ----
        const int xi = x - lane;
        const bool valid = xi >= end_x;
        int bl = 0; bool rdy = true;                       // row below at column xi-1
        if (valid && has_below && !below_fixed(xi - 1)) {
            const uint2 r = ld_below(xi - 1);
            rdy = (int)r.y == epoch; bl = (int)r.x;
        }
        int4 A = make_int4(0, 0, 0, 0); int G = 0;
        if (valid) { A = __ldcg(job.assumed + row0 + xi); G = __ldcg(job.mvs + row0 + xi); }
        const bool has_r = xi < g.mb_w - 1;
        const bool pollable = valid && has_below && !below_fixed(xi - 1);
        // ---- walk the chunk: runs of kept MBs retire together, a mismatch is searched in place ----
        int p = 0;                                         // lanes below p are final
        int rm = right_mv;                                 // final MV to the right of lane p
        uint2 nfe0 = make_uint2(0, 0), nfe1 = nfe0; int nfe_x = -1;   // prefetched source rows of MB nfe_x
        for (;;) {
            const int bl1 = __shfl_up_sync(FULLM, bl, 1), bl2 = __shfl_up_sync(FULLM, bl, 2);
            const int rd1 = __shfl_up_sync(FULLM, (int)rdy, 1), rd2 = __shfl_up_sync(FULLM, (int)rdy, 2);
            const int b0 = lane >= 1 ? bl1 : c0;
            const int bp1 = lane >= 2 ? bl2 : (lane == 1 ? c0 : c1);
            const bool cond = valid && rdy && (lane < 1 || rd1) && (lane < 2 || rd2);
            const int in_b = has_below ? b0 : 0;
            const int in_bl = (has_below && xi > 0) ? bl : 0;
            const int in_br = (has_below && has_r) ? bp1 : 0;
            const bool below_ok = cond && A.y == in_b && A.z == in_bl && A.w == in_br;
            const int Gr = __shfl_up_sync(FULLM, G, 1);
            const int in_r = has_r ? (lane == p ? rm : Gr) : 0;
            const bool hit = lane >= p && below_ok && A.x == in_r && !P.force_miss;
            const unsigned hb = __ballot_sync(FULLM, hit) >> p;
            int n = __ffs(~hb) - 1;
            if (n < 0) n = 32;
            if (n > 0) {
                if (lane >= p && lane < p + n) publish(xi, G);
                rm = __shfl_sync(FULLM, G, p + n - 1);
                p += n; n_hit += n;
                if (p >= 32) break;
            }
            if (!__shfl_sync(FULLM, (int)cond, p)) break;  // end of row, or the row below is not there yet
            // ---- mismatch at lane p: search MB x-p now, from its true inputs ----
            n_miss++;
            const int mb_x = x - p, mb_xy = row0 + mb_x;
            const int b_m1 = __shfl_sync(FULLM, bl, p), b_0 = __shfl_sync(FULLM, b0, p), b_p1 = __shfl_sync(FULLM, bp1, p);
            m.pel = 8 * (mb_x + mb_y * g.lstride);
            m.px = 8 * mb_x; m.py = 8 * mb_y;
            if (nfe_x == mb_x) { m.fe[0] = nfe0; m.fe[1] = nfe1; }
            else {
                const int pel = m.pel + m.r0 * g.lstride;
                m.fe[0] = load8u(job.fenc + pel); m.fe[1] = load8u(job.fenc + pel + g.lstride);
            }
            // while this MB is searched: the lanes still waiting for the row below poll again, and
            // the next MB's source rows are fetched (it usually mismatches too where this one does)
            uint2 pre = make_uint2(0, 0);
            const bool repoll = lane > p && pollable && !rdy;
            if (repoll) pre = ld_below(xi - 1);
            if (mb_x - 1 >= end_x) {
                const int pel = m.pel - 8 + m.r0 * g.lstride;
                nfe0 = load8u(job.fenc + pel); nfe1 = load8u(job.fenc + pel + g.lstride);
                nfe_x = mb_x - 1;
            }
            const bool mhas_r = mb_x < g.mb_w - 1, has_bl = has_below && mb_x > 0, has_br = has_below && mhas_r;
            int k0, k1, k2, k3, i_mvc;
            if (has_below) {
                k0 = mhas_r ? rm : b_0;
                k1 = mhas_r ? b_0 : (has_bl ? b_m1 : 0);
                k2 = mhas_r ? (has_bl ? b_m1 : b_p1) : 0;
                k3 = (mhas_r && has_bl) ? b_p1 : 0;
                i_mvc = (int)mhas_r + 1 + (int)has_bl + (int)has_br;
            } else {
                k0 = mhas_r ? rm : 0; k1 = k2 = k3 = 0;
                i_mvc = (int)mhas_r;
            }
            const int mvc[4][2] = {{mv_x(k0), mv_y(k0)}, {mv_x(k1), mv_y(k1)}, {mv_x(k2), mv_y(k2)}, {mv_x(k3), mv_y(k3)}};
            if (i_mvc <= 1) { m.mvp_x = mvc[0][0]; m.mvp_y = mvc[0][1]; }
            else { m.mvp_x = median3i(mvc[0][0], mvc[1][0], mvc[2][0]); m.mvp_y = median3i(mvc[0][1], mvc[1][1], mvc[2][1]); }
            int min_sx, max_sx, min_sy, max_sy;
            mv_limits(mb_x, mb_y, g.mb_w, g.mb_h, P.mv_range2, min_sx, max_sx, min_sy, max_sy);
            int out_mv = 0, out_cost = 0;
            bool skip = false;
            if (!(m.mvp_x | m.mvp_y)) {
                const int pel = m.pel + m.r0 * g.lstride;
                const uint2 a[2] = {load8u(job.fref[0] + pel), load8u(job.fref[0] + pel + g.lstride)};
                const int cz = P.satd ? satd_rows(m, a) : part_sum<4>(sad_rows(m, a));
                if (P.satd) n_satd++; else n_sad++;
                if (cz < 64) { skip = true; out_mv = 0; out_cost = cz; }
            }
            if (!skip) {
                MeResult2 r = me_search_mb2<4, QPRED>(m, sm, g, P, true, mvc, i_mvc, min_sx, max_sx, min_sy, max_sy, n_sad, n_satd);
                int cost = r.cost - (int)__ldg(P.cost_mv);      // remove mvcost from skip mbs
                if (r.mvx | r.mvy) cost += 5;
                out_mv = mv_pack(r.mvx, r.mvy); out_cost = cost;
            }
            if (lane == 0) {
                job.mvs[mb_xy] = out_mv;
                job.mv_costs[mb_xy] = out_cost;
                // a relaxation pass records what it searched from, for the ordered launch to check
                if (RELAX) job.assumed[mb_xy] = make_int4(mhas_r ? rm : 0, has_below ? b_0 : 0, has_bl ? b_m1 : 0, has_br ? b_p1 : 0);
                publish(mb_x, out_mv);
            }
            if (lane == p) G = out_mv;                     // the next lane's right neighbour
            if (repoll && (int)pre.y == epoch) { rdy = true; bl = (int)pre.x; }
            rm = out_mv;
            p += 1;
            if (p >= 32) break;
        }
        if (p > 0) {
            const int nc0 = __shfl_sync(FULLM, bl, p - 1);
            const int nc1 = p >= 2 ? __shfl_sync(FULLM, bl, p - 2) : c0;
            c0 = nc0; c1 = nc1;
            right_mv = rm;
            x -= p; ns = 32;
        } else { __nanosleep(below_local ? 20 : ns); if (ns < 256) ns <<= 1; }
    }
    if (P.stats && lane == 0) {
        if (RELAX) atomicAdd(P.stats + 2 + min(relax_pass, 3), n_miss);       // counted with the parallel passes
        else { atomicAdd(P.stats, n_hit); atomicAdd(P.stats + 1, n_miss); }
        atomicAdd(P.stats + 6, n_sad); atomicAdd(P.stats + 7, n_satd);
    }
  }
}


// relax_pass < 0: the exact ordered verification; >= 1: a row-relaxation pass (counted as parallel pass `relax_pass`)
int launch_me_verify(cudaStream_t st, const LaGeom &g, const MeParams &p, int relax_pass)
{
    if (p.njobs <= 0) return 0;
    // every band of a search resident at once: the row pipeline is up to mb_w/2 rows deep
    const dim3 grid((g.mb_h + VERIFY_ROWS - 1) / VERIFY_ROWS, p.njobs);
    const size_t smem = VERIFY_ROWS * sizeof(GroupSmem) + (size_t)VERIFY_ROWS * g.mb_w * sizeof(int2);
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(me_verify_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        cudaFuncSetAttribute(me_verify_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        cudaFuncSetAttribute(me_verify_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        cudaFuncSetAttribute(me_verify_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        attr_done = true;
    }
    if (smem > 160 * 1024) { set_error("frame too wide for the verification kernel (%d MBs)", g.mb_w); return -1; }
    const bool q = p.subpel_refine >= 3;
    if (relax_pass >= 1) {
        if (q) me_verify_kernel<true, true><<<grid, 32 * VERIFY_ROWS, smem, st>>>(g, p, relax_pass);
        else me_verify_kernel<false, true><<<grid, 32 * VERIFY_ROWS, smem, st>>>(g, p, relax_pass);
    } else {
        if (q) me_verify_kernel<true, false><<<grid, 32 * VERIFY_ROWS, smem, st>>>(g, p, 0);
        else me_verify_kernel<false, false><<<grid, 32 * VERIFY_ROWS, smem, st>>>(g, p, 0);
    }
    XV_LAUNCH_CHECK();
    return 0;
}

} // namespace xv
