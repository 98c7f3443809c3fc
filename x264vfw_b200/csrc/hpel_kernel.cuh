// Half-pel reference planes of a reconstructed frame.
//
// Replaces, in one pass over the tight w x h plane:
//   [x264] common/frame.c x264_frame_expand_border           -- 32-pixel edge replication of the frame
//   [x264] common/mc.c    hpel_filter via x264_frame_filter  -- 6-tap (1,-5,20,20,-5,1) H, V and centre planes,
//                                                               computed 8 pixels beyond the frame
//   [x264] common/frame.c x264_frame_expand_border_filtered  -- edge replication of the three filtered planes,
//                                                               starting 4 columns / 8 rows outside the frame
// Every step is "clamp a coordinate":
//   P0(x,y) = S(clamp x, clamp y)                    S = the frame, clamped to [0,w-1] x [0,h-1]
//   Pi(x,y) = Fi(clamp(x,-4,w+3), clamp(y,-8,h+7))   Fi = the filter evaluated on P0
// so the kernel walks the domain [-4,w+4) x [-8,h+8) once, and the units that own its first / last
// column or row also write the replicated border.  Read the frame once, write four padded planes once.
//
// Mapping.  A lane owns 8 pixels (two 32-bit words) of a strip of rows, a warp 30 such lanes (240 pixels);
// lanes 0 and 31 carry the neighbouring words (halo) so that everything horizontal is a shuffle, and the
// shuffle's "no such lane -> own value" is exactly the clamped neighbour at the frame's left / right edge.
// Tiles cover [0,w); the two words just outside ([-8,0) and [w,w+8)) are the halo lanes of the first /
// last tile, which also own the replicated border.  The six source rows of the vertical filter slide
// through registers (four packed 16-bit pairs per row; two rows per loop trip, see the note at the loop).  Arithmetic:
//   V  : packed 16-bit lanes, biased by 2576 = 80*32 + 16 so that lanes never go negative (no borrow between
//        lanes) and the bias supplies the rounding term; >>5, per-lane add/min/relu (DPX) gives clip().
//   H  : two dp4a per pixel on byte windows cut out of (left, own, own, right) words with PRMT.
//   C  : three dp2a per pixel on pairs of the biased 16-bit vertical sums (own + neighbours' by shuffle),
//        32-bit accumulation as upstream's C code (the 16-bit trick of upstream's asm can overflow).
//   clip + pack of H and C: cvt.pack.sat.u8.s32.
// About 23 issued instructions per pixel of an interior tile (182 per lane and row, 124 of them the filter) for
// 5 bytes of traffic per pixel.  Specialisations: ALIGNED (64-bit loads / cp.async vs byte gathers, decided on the
// host by hpel_plan) and EDGE (tiles that hold the frame's first or last column do the border work; the others run
// a loop without it, hpel_unit_any).
// Pinned beyond the checker: the four planes, read at every quarter-sample position, reproduce the motion-
// compensated pictures of an independent H.264 decoder (tests/golden/h264_pins.json; the lockstep run in the CPU suite
// for this source on the CPU, tests/test_next_hpel_gpu.py for the device).
//
// This header is compiled by nvcc (hpel_kernels.cu) and, with the lockstep warp shim in tests/sim/, by g++:
// the CPU suite runs this very code against the CPU checker.  The shim is test infrastructure; the product has no
// CPU path (the launcher lives in hpel_kernels.cu only).
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace xv {

struct HpelJob {
    const uint8_t *src; int src_stride; int w, h;     // tight plane, w % 8 == 0
    uint8_t *dst;                                     // 4 padded planes per frame: P0, H, V, C
    int stride; size_t plane_bytes;                   // stride >= w + 64, plane_bytes = stride * (h + 64)
    int rows_per_strip;
    int ntiles, nstrips;                              // ceil(w/8 / 30), ceil((h+16) / rows_per_strip)
    int h_m1, h_p7;                                   // h - 1, h + 7: compared every trip, read straight from the parameter bank
    int aligned;                                      // every frame's plane and the stride are 8-byte aligned: picks the instantiation
    // filter constants, read as parameter-bank operands (as literals each is rebuilt by a move in every row)
    uint32_t k_h_lo, k_h_hi, k_c_m2, k_c_0, k_c_p2, k_clip;
    size_t src_frame_bytes, dst_frame_bytes;
};

#define HPEL_PAD   32
#define HPEL_TILE  30          // 8-pixel words per warp that lie inside the frame (lanes 1..30)

// launch plan: fills ntiles / rows_per_strip / nstrips; returns the number of warps (units) per frame
static inline long long hpel_plan(HpelJob &job, int n_frames)
{
    const int nw8 = job.w >> 3;
    job.ntiles = (nw8 + HPEL_TILE - 1) / HPEL_TILE;
    // enough warps for every SM even with one frame: strips of 12 rows; 24 once a batch fills the GPU
    if (job.rows_per_strip <= 0)
        job.rows_per_strip = (long long)job.ntiles * ((job.h + 16 + 23) / 24) * n_frames >= 148 * 24 ? 24 : 12;
    job.nstrips = (job.h + 16 + job.rows_per_strip - 1) / job.rows_per_strip;
    job.h_m1 = job.h - 1; job.h_p7 = job.h + 7;
    job.k_h_lo = 0x1414FB01u; job.k_h_hi = 0x000001FBu;       // H: bytes x-2..x+1 and x+2..x+5 times (1,-5,20,20) (-5,1,0,0)
    job.k_c_m2 = 0xFB01u; job.k_c_0 = 0x1414u; job.k_c_p2 = 0x01FBu;   // centre: 16-bit pairs times (1,-5) (20,20) (-5,1)
    job.k_clip = 0xFFB0FFB0u;                                 // packed -80: removes the bias of the vertical sums after >>5
    job.aligned = ((((uintptr_t)job.src) | (uintptr_t)(uint32_t)job.src_stride | (n_frames > 1 ? (uintptr_t)job.src_frame_bytes : 0)) & 7) == 0;
    return (long long)job.ntiles * job.nstrips;
}

#ifndef XV_HPEL_HOST_ONLY
struct alignas(8) HpelWord { uint32_t x, y; };         // 8 pixels: x = the first four

// A lane's 8 pixels of one row.  Aligned planes (the normal case): every lane issues one 64-bit load; the lanes
// whose word lies left / right of the frame read the frame's first / last word instead and hpel_fix_word turns
// it into the replicated edge pixel when the row is consumed (two trips later) -- nothing depends on the load
// when it is issued.  Unaligned planes: eight clamped byte loads.
XV_DEVICE HpelWord hpel_load_word(const uint8_t *rowc, int fx, int cfx, int w, bool aligned)   // rowc = row + cfx
{
    HpelWord r;
    if (aligned) { xv_ld_u64(rowc, r.x, r.y); return r; }
    uint32_t b[8];
#pragma unroll
    for (int i = 0; i < 8; i++) b[i] = xv_ld_u8(rowc + (min(max(fx + i, 0), w - 1) - cfx));
    r.x = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
    r.y = b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24);
    return r;
}

XV_DEVICE HpelWord hpel_fix_word(HpelWord wd, int side)      // side: -1 left of the frame, +1 right of it, 0 inside
{
    if (side < 0) wd.x = wd.y = (wd.x & 0xFFu) * 0x01010101u;
    if (side > 0) wd.x = wd.y = (wd.y >> 24) * 0x01010101u;
    return wd;
}

#define HPEL_DIST  3            // rows in flight per lane
#define HPEL_RING  4            // ring slots (power of two > HPEL_DIST)
#define HPEL_UNROLL 2           // rows per loop trip (divides HPEL_RING)

// request a lane's 8 pixels of one row into its ring slot (one cp.async group per call)
XV_DEVICE void hpel_fetch_row(xv_saddr slot, const uint8_t *rowc, int fx, int cfx, int w, bool aligned)
{
    if (aligned) xv_cp_async8(slot, rowc);
    else { const HpelWord wd = hpel_load_word(rowc, fx, cfx, w, false); xv_sts_u64(slot, wd.x, wd.y); }
    xv_cp_async_commit();
}

XV_DEVICE void hpel_store4(uint8_t *d, size_t pb, const HpelWord v[4])
{
    // a chain of three 64-bit adds (opaque: otherwise 2*pb and 3*pb are built with wide multiplies, 10 instructions)
    xv_st_u64(d, v[0].x, v[0].y); d = xv_opaque(d + pb);
    xv_st_u64(d, v[1].x, v[1].y); d = xv_opaque(d + pb);
    xv_st_u64(d, v[2].x, v[2].y); d = xv_opaque(d + pb);
    xv_st_u64(d, v[3].x, v[3].y);
}

XV_DEVICE uint32_t hpel_clip_v(uint32_t v, const HpelJob &k)   // packed (V16 + 2576): ((V16+16)>>5) clipped to [0,255]
{
    return xv_addmin_relu_s16x2((v >> 5) & 0x07FF07FFu, k.k_clip, 0x00FF00FFu);
}

XV_DEVICE int hpel_tap_h(uint32_t win_m2, uint32_t win_p2, const HpelJob &k)   // bytes x-2..x+1 and x+2..x+5 (the last two unused)
{
    return xv_dp4a_us(win_m2, k.k_h_lo, xv_dp4a_us(win_p2, k.k_h_hi, 16)) >> 5;
}

XV_DEVICE int hpel_tap_c(uint32_t qm2, uint32_t q0, uint32_t qp2, const HpelJob &k)   // pairs (V[x-2],V[x-1]) (V[x],V[x+1]) (V[x+2],V[x+3])
{
    return xv_dp2a_lo(qm2, k.k_c_m2, xv_dp2a_lo(q0, k.k_c_0, xv_dp2a_lo(qp2, k.k_c_p2, 512 - 32 * 2576))) >> 10;
}

// one warp: tile `unit % ntiles` of strip `unit / ntiles` of frame `frame`.  ALIGNED (= job.aligned, decided by
// hpel_plan) selects 64-bit loads / cp.async; otherwise bytes are gathered.  EDGE = the tile holds the frame's first
// or last column (hpel_unit_any decides): interior tiles run a loop without any of the border work.
template <bool ALIGNED, bool EDGE>
XV_DEVICE void hpel_unit(const HpelJob &job, int unit, int frame, int lane, HpelWord (*ring)[32])
{
    const int tile = unit % job.ntiles, strip = unit / job.ntiles;
    const int w = job.w, h = job.h;
    const int nw8 = w >> 3;                                   // 8-pixel words inside the frame
    const int wj = tile * HPEL_TILE + lane - 1;               // this lane's word; -1 and nw8 = the words just outside
    const int fx = 8 * wj;                                    // frame x of its first pixel
    const uint8_t *S = job.src + (size_t)frame * job.src_frame_bytes;
    const int ss = job.src_stride;
    const int right_lane = (w >> 3) - tile * HPEL_TILE + 1;   // lane of the word just right of the frame
    // who stores: lanes 1..30 up to and including the word just right of the frame; lane 0 only as the word
    // just left of the frame; lane 31 only as the word just right of it.  The rest of the border: columns 0..23
    // (lanes 1..3 of the first tile) and w+40..w+63 (lanes 4..6 of the last tile) repeat the first / last
    // filtered pixel of the row.
    // Facts that never change, kept in two registers (recomputing them from the kernel parameters and the lane
    // number cost two dozen issue slots per row).  `flags` is warp-uniform (branches on it hold shuffles):
    // 2 = first tile of the row, 4 = last tile of the row.  `lflags` is per lane:
    // 1 = stores its own word, 2 = stores a border word, 4 = owns the word just left of the frame, 8 = the word
    // just right of it, 16 = lane > 3.
    const bool lt = tile == 0, rt = right_lane >= 1 && right_lane <= 31;
    const uint32_t flags = xv_opaque_u32((lt ? 2u : 0u) | (rt ? 4u : 0u));
    const uint32_t lflags = xv_opaque_u32(((lane >= 1 && lane <= HPEL_TILE && wj <= nw8) || wj == -1 || wj == nw8 ? 1u : 0u) |
                                          (lane >= 1 && lane <= 6 && (lane <= 3 ? lt : rt) ? 2u : 0u) |
                                          (wj == -1 ? 4u : 0u) | (wj == nw8 ? 8u : 0u) | (lane > 3 ? 16u : 0u));
    constexpr bool aligned = ALIGNED;
    const int cfx = min(max(fx, 0), w - 8);                   // column actually loaded (aligned planes)
    const int side = !aligned ? 0 : fx < 0 ? -1 : fx >= w ? 1 : 0;
    const int fy0 = strip * job.rows_per_strip - 8;

    uint8_t *D = xv_opaque(job.dst + (size_t)frame * job.dst_frame_bytes);
    const uint32_t own_off = (uint32_t)(fx + HPEL_PAD);
    const uint32_t edge_off = lane <= 3 ? 8u * (lane - 1) : (uint32_t)(w + HPEL_PAD + 8) + 8u * (lane - 4);
    const bool left_tile = EDGE && (flags & 2u) != 0, right_tile = EDGE && (flags & 4u) != 0, edge_tile = EDGE;
    const bool store_lane = (lflags & 1u) != 0, edge_lane = EDGE && (lflags & 2u) != 0;
    const bool left_word = (lflags & 4u) != 0, right_word = (lflags & 8u) != 0;
    const size_t pb = job.plane_bytes;

    // sliding window: row fy-2+k of the (clamped) frame lives in s[k], widened to 16-bit pairs
    uint32_t s[5 + HPEL_UNROLL][4];
    {
        HpelWord first[5];                                    // all five requests go out before the first is used
        if (aligned) {
#pragma unroll
            for (int k = 0; k < 5; k++) xv_ld_u64(S + ((size_t)min(max(fy0 - 2 + k, 0), h - 1) * ss + cfx), first[k].x, first[k].y);
        } else {
#pragma unroll
            for (int k = 0; k < 5; k++) first[k] = hpel_load_word(S + ((size_t)min(max(fy0 - 2 + k, 0), h - 1) * ss + cfx), fx, cfx, w, false);
        }
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const HpelWord wd = hpel_fix_word(first[k], side);
            s[k][0] = xv_prmt(wd.x, 0u, 0x4140); s[k][1] = xv_prmt(wd.x, 0u, 0x4342);
            s[k][2] = xv_prmt(wd.y, 0u, 0x4140); s[k][3] = xv_prmt(wd.y, 0u, 0x4342);
        }
    }

    // The next HPEL_DIST rows are always in flight, through a per-lane ring in shared memory filled by cp.async:
    // the copy has no destination register, so nothing in the loop waits on it until the row is due (a register
    // double buffer needs either an unrolled loop or a copy of a register that is still being filled, and that
    // copy waits for the load -- it held 25 % of the stall samples).
    const xv_saddr ring0 = xv_opaque_saddr(xv_saddr_of(&ring[0][lane]));   // slot k of this lane = ring0 + 256 k; opaque: otherwise
                                                                         // it is rematerialised (two S2R + five more) every trip
#pragma unroll
    for (int k = 0; k < HPEL_DIST; k++)
        hpel_fetch_row(ring0 + 256u * k, S + ((size_t)min(max(fy0 + 3 + k, 0), h - 1) * ss + cfx), fx, cfx, w, aligned);
    // running store address of row fy (own word; the border word is edge_delta away)
    uint8_t *dp = D + ((size_t)(fy0 + HPEL_PAD) * job.stride + own_off);
    // running load address: row clamp(fy+3+HPEL_DIST) of the frame, this lane's column; it moves down while inside the frame
    const uint8_t *rp = S + ((size_t)min(max(fy0 + 3 + HPEL_DIST, 0), h - 1) * ss + cfx);
    const ptrdiff_t edge_delta = (ptrdiff_t)edge_off - (ptrdiff_t)own_off;

    // Two rows per trip (HPEL_UNROLL): the window rotates by renaming inside a trip and moves by register copies
    // between trips (10 copies per row instead of 20); the loop body stays ~7 KB on the interior path -- unrolled
    // by six (no copies at all) it was 55 KB, beyond the 32 KB L1.5 instruction cache, and a quarter of the stall
    // samples were "no instruction".  The trip count is computed up front: an EXIT inside the loop, even
    // predicated off, waits for every load in flight.  An odd row count runs one row too many: it is computed
    // from clamped (valid) source rows and not stored.
#pragma unroll 1
    for (int fy = fy0, i = 0, left = min(job.rows_per_strip, h + 8 - fy0); left > 0;
         left -= HPEL_UNROLL, fy += HPEL_UNROLL, i = (i + HPEL_UNROLL) & (HPEL_RING - 1)) {   // i = ring slot of row fy+3
#pragma unroll
      for (int u = 0; u < HPEL_UNROLL; u++) {             // row fy+u: window s[u] .. s[u+5]
        const bool live = u == 0 || left > u;             // warp-uniform
        {   // request row fy+u+3+DIST, take delivery of row fy+u+3
            hpel_fetch_row(ring0 + 256u * ((i + u + HPEL_DIST) & (HPEL_RING - 1)), rp, fx, cfx, w, aligned);
            if ((unsigned)(fy + u + 3 + HPEL_DIST) < (unsigned)job.h_m1) rp += ss;
            xv_cp_async_wait<HPEL_DIST>();                                                       // row fy+u+3 has landed
            HpelWord wd;
            xv_lds_u64(ring0 + 256u * ((i + u) & (HPEL_RING - 1)), wd.x, wd.y);
            if (edge_tile) wd = hpel_fix_word(wd, side);
            uint32_t *n = s[u + 5];
            n[0] = xv_prmt(wd.x, 0u, 0x4140); n[1] = xv_prmt(wd.x, 0u, 0x4342);
            n[2] = xv_prmt(wd.y, 0u, 0x4140); n[3] = xv_prmt(wd.y, 0u, 0x4342);
        }
        // ---- vertical 6-tap on packed pairs, lanes biased by 2576 --------------------------------
        uint32_t v[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t t = s[u][q] + s[u + 5][q] + 0x0A100A10u;
            t += 20u * (s[u + 2][q] + s[u + 3][q]);
            t -= 5u * (s[u + 1][q] + s[u + 4][q]);
            v[q] = t;
        }
        HpelWord out[4];
        // V plane
        out[2].x = xv_prmt(hpel_clip_v(v[0], job), hpel_clip_v(v[1], job), 0x6420);
        out[2].y = xv_prmt(hpel_clip_v(v[2], job), hpel_clip_v(v[3], job), 0x6420);
        {   // P0 and H plane: byte windows of the 16-byte span (L, w0, w1, R), span offset 4 = own pixel 0
            const uint32_t w0 = xv_prmt(s[u + 2][0], s[u + 2][1], 0x6420);
            const uint32_t w1 = xv_prmt(s[u + 2][2], s[u + 2][3], 0x6420);
            const uint32_t L = xv_shfl_up1(w1), R = xv_shfl_down1(w0);
            const uint32_t o2 = xv_prmt(L, w0, 0x5432), o3 = xv_prmt(L, w0, 0x6543), o5 = xv_prmt(w0, w1, 0x4321);
            const uint32_t o6 = xv_prmt(w0, w1, 0x5432), o7 = xv_prmt(w0, w1, 0x6543), o9 = xv_prmt(w1, R, 0x4321);
            const uint32_t o10 = xv_prmt(w1, R, 0x5432), o11 = xv_prmt(w1, R, 0x6543), o13 = R >> 8;
            out[0].x = w0; out[0].y = w1;
            out[1].x = xv_pack_sat_u8(hpel_tap_h(o2, o6, job), hpel_tap_h(o3, o7, job), hpel_tap_h(w0, w1, job), hpel_tap_h(o5, o9, job));
            out[1].y = xv_pack_sat_u8(hpel_tap_h(o6, o10, job), hpel_tap_h(o7, o11, job), hpel_tap_h(w1, R, job), hpel_tap_h(o9, o13, job));
        }
        {   // C plane: horizontal 6-tap over the vertical sums (pairs q[k] = (V[k], V[k+1]), k = -2..9)
            const uint32_t Lv3 = xv_shfl_up1(v[3]), R0 = xv_shfl_down1(v[0]), R1 = xv_shfl_down1(v[1]);
            const uint32_t qm1 = xv_prmt(Lv3, v[0], 0x5432), q1 = xv_prmt(v[0], v[1], 0x5432);
            const uint32_t q3 = xv_prmt(v[1], v[2], 0x5432), q5 = xv_prmt(v[2], v[3], 0x5432);
            const uint32_t q7 = xv_prmt(v[3], R0, 0x5432), q9 = xv_prmt(R0, R1, 0x5432);
            out[3].x = xv_pack_sat_u8(hpel_tap_c(Lv3, v[0], v[1], job), hpel_tap_c(qm1, q1, q3, job),
                                      hpel_tap_c(v[0], v[1], v[2], job), hpel_tap_c(q1, q3, q5, job));
            out[3].y = xv_pack_sat_u8(hpel_tap_c(v[1], v[2], v[3], job), hpel_tap_c(q3, q5, q7, job),
                                      hpel_tap_c(v[2], v[3], R0, job), hpel_tap_c(q5, q7, q9, job));
        }
        // ---- the words just outside the frame: 4 filtered pixels next to the frame, the other 4 already
        //      border (= the outermost filtered pixel); the rest of the border goes to the edge lanes -----
        HpelWord e[4];
        if (left_tile) {                                  // warp-uniform
#pragma unroll
            for (int p = 0; p < 4; p++) {
                const uint32_t rep = (out[p].y & 0xFFu) * 0x01010101u;
                if (left_word) out[p].x = rep;
                e[p].x = e[p].y = xv_shfl_idx(rep, 0);
            }
        }
        if (right_tile) {                                 // warp-uniform
#pragma unroll
            for (int p = 0; p < 4; p++) {
                const uint32_t rep = (out[p].x >> 24) * 0x01010101u;
                if (right_word) out[p].y = rep;
                const uint32_t br = xv_shfl_idx(rep, right_lane & 31);
                if (lflags & 16u) e[p].x = e[p].y = br;
            }
        }
        if (live) {
            if (store_lane) hpel_store4(dp, pb, out);
            if (edge_lane) hpel_store4(dp + edge_delta, pb, e);
            if (fy + u == -8 || fy + u == job.h_p7) {     // top / bottom border: 24 more copies of this row
                const int rb = fy + u == -8 ? 0 : h + HPEL_PAD + 8;
#pragma unroll 1
                for (int r = rb; r < rb + 24; r++) {
                    const size_t rr = (size_t)r * job.stride;
                    if (store_lane) hpel_store4(D + (rr + own_off), pb, out);
                    if (edge_lane) hpel_store4(D + (rr + edge_off), pb, e);
                }
            }
        }
        dp += job.stride;
      }
#pragma unroll
        for (int k = 0; k < 5; k++) {
#pragma unroll
            for (int q = 0; q < 4; q++) s[k][q] = s[k + HPEL_UNROLL][q];
        }
    }
}
template <bool ALIGNED>
XV_DEVICE void hpel_unit_any(const HpelJob &job, int unit, int frame, int lane)
{
    XV_SHARED HpelWord ring[HPEL_RING][32];               // the prefetch ring of this warp (one warp per block)
    const int tile = unit % job.ntiles;
    const int right_lane = (job.w >> 3) - tile * HPEL_TILE + 1;
    if (tile == 0 || (right_lane >= 1 && right_lane <= 31)) hpel_unit<ALIGNED, true>(job, unit, frame, lane, ring);
    else hpel_unit<ALIGNED, false>(job, unit, frame, lane, ring);
}
#endif // XV_HPEL_HOST_ONLY

} // namespace xv
