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
// Mapping.  A warp owns 30 32-bit words (120 pixels) of a strip of rows; lanes 0 and 31 carry the
// neighbouring words (halo) so that everything horizontal is a shuffle.  The six source rows of the
// vertical filter slide through registers (two packed 16-bit pairs per row, loop unrolled by six so
// the rotation is register renaming).  Arithmetic per 4 pixels:
//   V  : packed 16-bit lanes, biased by 2576 = 80*32 + 16 so that lanes never go negative (no borrow between
//        lanes) and the bias supplies the rounding term; >>5, per-lane add/min/relu (DPX) gives clip().
//   H  : two dp4a per pixel on byte windows cut out of (left, own, right) words with PRMT.
//   C  : three dp2a per pixel on pairs of the biased 16-bit vertical sums (own + neighbours' by shuffle),
//        32-bit accumulation as upstream's C code (the 16-bit trick of upstream's asm can overflow).
//   clip + pack of H and C: cvt.pack.sat.u8.s32.
//
// This header is compiled by nvcc (hpel_kernels.cu) and, with the lockstep warp shim in tests/sim/, by g++:
// the CPU suite runs this very code against the CPU checker.  The shim is test infrastructure; the product has no
// CPU path (the launcher lives in hpel_kernels.cu only).
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace xv {

struct HpelJob {
    const uint8_t *src; int src_stride; int w, h;     // tight plane, w % 4 == 0
    uint8_t *dst;                                     // 4 padded planes per frame: P0, H, V, C
    int stride; size_t plane_bytes;                   // stride >= w + 64, plane_bytes = stride * (h + 64)
    int rows_per_strip;                               // multiple of 6
    int ntiles, nstrips;                              // ceil((w+8)/4 / 30), ceil((h+16) / rows_per_strip)
    size_t src_frame_bytes, dst_frame_bytes;
};

#define HPEL_PAD   32
#define HPEL_TILE  30          // words per warp that are stored (lanes 1..30)

// launch plan: fills ntiles / rows_per_strip / nstrips; returns the number of warps (units) per frame
static inline long long hpel_plan(HpelJob &job, int n_frames)
{
    const int nwords = (job.w + 8) >> 2;
    job.ntiles = (nwords + HPEL_TILE - 1) / HPEL_TILE;
    // enough warps for every SM even with one frame: strips of 12 rows; 24 once a batch fills the GPU
    if (job.rows_per_strip <= 0)
        job.rows_per_strip = (long long)job.ntiles * ((job.h + 16 + 23) / 24) * n_frames >= 148 * 32 ? 24 : 12;
    job.nstrips = (job.h + 16 + job.rows_per_strip - 1) / job.rows_per_strip;
    return (long long)job.ntiles * job.nstrips;
}

#ifndef XV_HPEL_HOST_ONLY
XV_DEVICE uint32_t hpel_load_word(const uint8_t *row, int fx, int w, bool direct)
{
    if (direct) return xv_ld_u32(row + fx);
    const int x0 = min(max(fx, 0), w - 1), x1 = min(max(fx + 1, 0), w - 1);
    const int x2 = min(max(fx + 2, 0), w - 1), x3 = min(max(fx + 3, 0), w - 1);
    return (uint32_t)xv_ld_u8(row + x0) | ((uint32_t)xv_ld_u8(row + x1) << 8) |
           ((uint32_t)xv_ld_u8(row + x2) << 16) | ((uint32_t)xv_ld_u8(row + x3) << 24);
}

XV_DEVICE void hpel_store4(uint8_t *D, uint32_t o, uint32_t pb, const uint32_t v[4])
{
    uint8_t *d = D + o;
    xv_st_u32(d, v[0]); d += pb;
    xv_st_u32(d, v[1]); d += pb;
    xv_st_u32(d, v[2]); d += pb;
    xv_st_u32(d, v[3]);
}

// one warp: tile `unit % ntiles` of strip `unit / ntiles` of frame `frame`
XV_DEVICE void hpel_unit(const HpelJob &job, int unit, int frame, int lane)
{
    const int tile = unit % job.ntiles, strip = unit / job.ntiles;
    const int w = job.w, h = job.h;
    const int nwords = (w + 8) >> 2;                          // words of the filtered domain [-4, w+4)
    const int wj = tile * HPEL_TILE + lane - 1;               // this lane's word (lanes 0, 31: halo)
    const int fx = 4 * wj - 4;                                // frame x of its first pixel
    const uint8_t *S = job.src + (size_t)frame * job.src_frame_bytes;
    const int ss = job.src_stride;
    const bool direct = fx >= 0 && fx + 3 < w && ((((uintptr_t)S) | (uintptr_t)(uint32_t)ss) & 3) == 0;
    const int fy0 = strip * job.rows_per_strip - 8;

    uint8_t *D = xv_opaque(job.dst + (size_t)frame * job.dst_frame_bytes);
    const bool store_lane = lane >= 1 && lane <= HPEL_TILE && wj < nwords;
    const int last_lane = nwords - tile * HPEL_TILE;          // lane that holds the last word of the row
    const bool left_tile = tile == 0, right_tile = last_lane >= 1 && last_lane <= HPEL_TILE;
    const uint32_t pb = (uint32_t)job.plane_bytes, own_off = (uint32_t)(fx + HPEL_PAD);
    // lanes 0..6 of the first tile write the left border (columns 0..27), lanes 7..13 of the last tile the
    // right border (columns w+36 .. w+63): the first / last pixel of the row, replicated
    const bool edge_tile = left_tile || right_tile;
    const bool edge_lane = lane < 7 ? left_tile : (lane < 14 && right_tile);
    const uint32_t edge_off = lane < 7 ? 4u * lane : (uint32_t)(w + HPEL_PAD + 4) + 4u * (lane - 7);

    // sliding window: row fy-2+k of the (clamped) frame lives in lo/hi[(j+k)%6], widened to 16-bit pairs
    uint32_t lo[6], hi[6];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const int sy = min(max(fy0 - 2 + k, 0), h - 1);
        const uint32_t word = hpel_load_word(S + (size_t)sy * ss, fx, w, direct);
        lo[k] = xv_prmt(word, 0u, 0x4140); hi[k] = xv_prmt(word, 0u, 0x4342);
    }

#pragma unroll 1
    for (int base = 0; base < job.rows_per_strip; base += 6) {
#pragma unroll
        for (int j = 0; j < 6; j++) {
            const int fy = fy0 + base + j;
            if (fy >= h + 8) return;                          // warp-uniform
            {
                const int sy = min(max(fy + 3, 0), h - 1);
                const uint32_t word = hpel_load_word(S + (size_t)sy * ss, fx, w, direct);
                lo[(j + 5) % 6] = xv_prmt(word, 0u, 0x4140); hi[(j + 5) % 6] = xv_prmt(word, 0u, 0x4342);
            }
            // ---- vertical 6-tap on packed pairs, lanes biased by 2576 --------------------------------
            uint32_t vlo = lo[j % 6] + lo[(j + 5) % 6] + 0x0A100A10u;
            uint32_t vhi = hi[j % 6] + hi[(j + 5) % 6] + 0x0A100A10u;
            vlo += 20u * (lo[(j + 2) % 6] + lo[(j + 3) % 6]);
            vhi += 20u * (hi[(j + 2) % 6] + hi[(j + 3) % 6]);
            vlo -= 5u * (lo[(j + 1) % 6] + lo[(j + 4) % 6]);
            vhi -= 5u * (hi[(j + 1) % 6] + hi[(j + 4) % 6]);
            uint32_t out[4];
            {   // V plane: ((v+16)>>5) + 80 per lane, then max(min(. - 80, 255), 0)
                const uint32_t rlo = xv_addmin_relu_s16x2((vlo >> 5) & 0x07FF07FFu, 0xFFB0FFB0u, 0x00FF00FFu);
                const uint32_t rhi = xv_addmin_relu_s16x2((vhi >> 5) & 0x07FF07FFu, 0xFFB0FFB0u, 0x00FF00FFu);
                out[2] = xv_prmt(rlo, rhi, 0x6420);
            }
            {   // P0 and H plane: byte windows of (left, own, right)
                const uint32_t raw = xv_prmt(lo[(j + 2) % 6], hi[(j + 2) % 6], 0x6420);
                const uint32_t L = xv_shfl_up1(raw), R = xv_shfl_down1(raw);
                const uint32_t a0 = xv_prmt(L, raw, 0x5432), a1 = xv_prmt(L, raw, 0x6543), a3 = xv_prmt(raw, R, 0x4321);
                const uint32_t b0 = xv_prmt(raw, R, 0x5432), b1 = xv_prmt(raw, R, 0x6543), b3 = R >> 8;
                const int h0 = xv_dp4a_us(a0, 0x1414FB01u, xv_dp4a_us(b0, 0x000001FBu, 16)) >> 5;
                const int h1 = xv_dp4a_us(a1, 0x1414FB01u, xv_dp4a_us(b1, 0x000001FBu, 16)) >> 5;
                const int h2 = xv_dp4a_us(raw, 0x1414FB01u, xv_dp4a_us(R, 0x000001FBu, 16)) >> 5;
                const int h3 = xv_dp4a_us(a3, 0x1414FB01u, xv_dp4a_us(b3, 0x000001FBu, 16)) >> 5;
                out[0] = raw;
                out[1] = xv_pack_sat_u8(h0, h1, h2, h3);
            }
            {   // C plane: horizontal 6-tap over the vertical sums; the bias contributes 32 * 2576
                const uint32_t Lhi = xv_shfl_up1(vhi), Rlo = xv_shfl_down1(vlo), Rhi = xv_shfl_down1(vhi);
                const uint32_t qm1 = xv_prmt(Lhi, vlo, 0x5432), q1 = xv_prmt(vlo, vhi, 0x5432);
                const uint32_t q3 = xv_prmt(vhi, Rlo, 0x5432), q5 = xv_prmt(Rlo, Rhi, 0x5432);
                const int cb = 512 - 32 * 2576;
                const int c0 = xv_dp2a_lo(Lhi, 0xFB01u, xv_dp2a_lo(vlo, 0x1414u, xv_dp2a_lo(vhi, 0x01FBu, cb))) >> 10;
                const int c1 = xv_dp2a_lo(qm1, 0xFB01u, xv_dp2a_lo(q1, 0x1414u, xv_dp2a_lo(q3, 0x01FBu, cb))) >> 10;
                const int c2 = xv_dp2a_lo(vlo, 0xFB01u, xv_dp2a_lo(vhi, 0x1414u, xv_dp2a_lo(Rlo, 0x01FBu, cb))) >> 10;
                const int c3 = xv_dp2a_lo(q1, 0xFB01u, xv_dp2a_lo(q3, 0x1414u, xv_dp2a_lo(q5, 0x01FBu, cb))) >> 10;
                out[3] = xv_pack_sat_u8(c0, c1, c2, c3);
            }
            // ---- stores: own word, plus the replicated border where this unit owns an edge ------------
            // (32-bit offsets from the frame's base: four planes of a frame stay below 4 GB)
            uint32_t e[4] = {0, 0, 0, 0};
            if (edge_tile) {                                  // warp-uniform
#pragma unroll
                for (int p = 0; p < 4; p++) {
                    const uint32_t bl = xv_shfl_idx(out[p], 1), br = xv_shfl_idx(out[p], last_lane & 31);
                    e[p] = (lane < 7 ? bl & 0xFFu : br >> 24) * 0x01010101u;
                }
            }
            const uint32_t ro = (uint32_t)(fy + HPEL_PAD) * (uint32_t)job.stride;
            if (store_lane) hpel_store4(D, ro + own_off, pb, out);
            if (edge_lane) hpel_store4(D, ro + edge_off, pb, e);
            if (fy == -8 || fy == h + 7) {                    // top / bottom border: 24 more copies of this row
                const int rb = fy == -8 ? 0 : h + HPEL_PAD + 8;
#pragma unroll 1
                for (int r = rb; r < rb + 24; r++) {
                    const uint32_t rr = (uint32_t)r * (uint32_t)job.stride;
                    if (store_lane) hpel_store4(D, rr + own_off, pb, out);
                    if (edge_lane) hpel_store4(D, rr + edge_off, pb, e);
                }
            }
        }
    }
}
#endif // XV_HPEL_HOST_ONLY

} // namespace xv
