// Half-resolution ("lowres") plane construction for the lookahead.
//
// Replaces, in one pass over the tight luma plane:
//   [x264] x264_frame_expand_border_mod16 (luma)      -- replicate to 16*mb_w x 16*mb_h
//   [x264] x264_frame_init_lowres: duplicate last row/column, frame_init_lowres_core
//          (FILTER(a,b,c,d) = (((a+b+1)>>1)+((c+d+1)>>1)+1)>>1, four phase planes 0/H/V/C)
//   [x264] x264_frame_expand_border_lowres            -- 32-pixel edge replication
// (reached in the reference only through codec.c:1693).
//
// The mod-16 replication and the duplicated row/column are both "clamp the source
// coordinate", so they are folded into addressing: P(x,y) = Y(min(x,w-1), min(y,h-1)).
// The 32-pixel border is "clamp the lowres coordinate", so the grid simply runs over the
// padded output domain and border threads compute the clamped pixel.  No intermediate
// padded luma plane, no separate border pass: read luma once, write 4 planes once.
#include "common.cuh"
#include "csp_kernels.h"

namespace xv {

#define LOWRES_PAD 32

// gather even / odd bytes of 16 bytes (+1 trailing byte) and form the two horizontal phases
__device__ __forceinline__ void hphase(const uint4 v, uint32_t tail, uint2 &p0, uint2 &ph)
{
    uint32_t e_lo = __byte_perm(v.x, v.y, 0x6420), e_hi = __byte_perm(v.z, v.w, 0x6420);
    uint32_t o_lo = __byte_perm(v.x, v.y, 0x7531), o_hi = __byte_perm(v.z, v.w, 0x7531);
    uint32_t n_lo = __funnelshift_r(e_lo, e_hi, 8), n_hi = __funnelshift_r(e_hi, tail, 8);
    p0 = make_uint2(avg4(e_lo, o_lo), avg4(e_hi, o_hi));
    ph = make_uint2(avg4(o_lo, n_lo), avg4(o_hi, n_hi));
}

__device__ __forceinline__ uint4 avg16(uint4 a, uint4 b)
{
    return make_uint4(avg4(a.x, b.x), avg4(a.y, b.y), avg4(a.z, b.z), avg4(a.w, b.w));
}

__device__ __forceinline__ uint32_t filt(uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    return (((a + b + 1) >> 1) + ((c + d + 1) >> 1) + 1) >> 1;
}

__global__ void __launch_bounds__(128)
lowres_init_kernel(LowresJob job)
{
    const int pchunk = blockIdx.x * blockDim.x + threadIdx.x;       // 8-px chunk of the padded row
    const int prow = blockIdx.y * blockDim.y + threadIdx.y;          // padded row
    if (pchunk >= ((job.lw + 2 * LOWRES_PAD) >> 3) || prow >= job.lh + 2 * LOWRES_PAD) return;
    const size_t f = blockIdx.z;
    const uint8_t *Y = job.y + f * job.src_frame_bytes;
    const int ys = job.y_stride, w = job.w, h = job.h;

    const int ly = min(max(prow - LOWRES_PAD, 0), job.lh - 1);   // clamped lowres row
    const int lx = (pchunk << 3) - LOWRES_PAD;                    // first lowres x of this chunk
    const int r0 = min(2 * ly, h - 1), r1 = min(2 * ly + 1, h - 1), r2 = min(2 * ly + 2, h - 1);
    const uint8_t *s0 = Y + (size_t)r0 * ys, *s1 = Y + (size_t)r1 * ys, *s2 = Y + (size_t)r2 * ys;

    uint2 o0, oh, ov, oc;
    const int sx = 2 * lx;
    if (lx >= 0 && lx + 8 <= job.lw && sx + 16 <= w && ((ys | (int)(uintptr_t)Y) & 15) == 0) {
        // interior: 16 source bytes per row + the 17th column (clamped: it is the duplicated
        // last column at the right frame edge)
        const uint4 a = *(const uint4 *)(s0 + sx), b = *(const uint4 *)(s1 + sx), c = *(const uint4 *)(s2 + sx);
        const int tx = min(sx + 16, w - 1);
        const uint32_t ta = s0[tx], tb = s1[tx], tc = s2[tx];
        hphase(avg16(a, b), (ta + tb + 1) >> 1, o0, oh);
        hphase(avg16(b, c), (tb + tc + 1) >> 1, ov, oc);
    } else if (lx + 8 <= 0 || lx >= job.lw) {
        // left/right padding: the whole chunk replicates the first/last lowres column
        const int x = lx < 0 ? 0 : job.lw - 1;
        const int c0 = min(2 * x, w - 1), c1 = min(2 * x + 1, w - 1), c2 = min(2 * x + 2, w - 1);
        const uint32_t a0 = s0[c0], a1 = s0[c1], a2 = s0[c2], b0 = s1[c0], b1 = s1[c1], b2 = s1[c2];
        const uint32_t d0 = s2[c0], d1 = s2[c1], d2 = s2[c2];
        const uint32_t v0 = filt(a0, b0, a1, b1) * 0x01010101u, vh = filt(a1, b1, a2, b2) * 0x01010101u;
        const uint32_t vv = filt(b0, d0, b1, d1) * 0x01010101u, vc = filt(b1, d1, b2, d2) * 0x01010101u;
        o0 = make_uint2(v0, v0); oh = make_uint2(vh, vh); ov = make_uint2(vv, vv); oc = make_uint2(vc, vc);
    } else {
        // odd geometry (width not a multiple of 16, unaligned plane): per-pixel, clamped
        uint32_t r[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
        for (int i = 0; i < 8; i++) {
            const int x = min(max(lx + i, 0), job.lw - 1);
            const int c0 = min(2 * x, w - 1), c1 = min(2 * x + 1, w - 1), c2 = min(2 * x + 2, w - 1);
            r[0][i >> 2] |= filt(s0[c0], s1[c0], s0[c1], s1[c1]) << (8 * (i & 3));
            r[1][i >> 2] |= filt(s0[c1], s1[c1], s0[c2], s1[c2]) << (8 * (i & 3));
            r[2][i >> 2] |= filt(s1[c0], s2[c0], s1[c1], s2[c1]) << (8 * (i & 3));
            r[3][i >> 2] |= filt(s1[c1], s2[c1], s1[c2], s2[c2]) << (8 * (i & 3));
        }
        o0 = make_uint2(r[0][0], r[0][1]); oh = make_uint2(r[1][0], r[1][1]);
        ov = make_uint2(r[2][0], r[2][1]); oc = make_uint2(r[3][0], r[3][1]);
    }
    uint8_t *d = job.dst + f * job.dst_frame_bytes + (size_t)prow * job.lstride + (pchunk << 3);
    *(uint2 *)(d) = o0;
    *(uint2 *)(d + (size_t)job.lplane_bytes) = oh;
    *(uint2 *)(d + 2 * (size_t)job.lplane_bytes) = ov;
    *(uint2 *)(d + 3 * (size_t)job.lplane_bytes) = oc;
}

int launch_lowres_init(cudaStream_t st, const LowresJob &job, int n_frames)
{
    const int chunks = (job.lw + 2 * LOWRES_PAD) >> 3, rows = job.lh + 2 * LOWRES_PAD;
    if (chunks <= 0 || rows <= 0 || n_frames <= 0) return 0;
    // block = (chunks of one row rounded to a warp multiple, up to 128) x (rows to reach 128 threads)
    const int bx = ((chunks < 128 ? chunks : 128) + 31) & ~31, by = bx >= 128 ? 1 : 128 / bx;
    dim3 block(bx, by), grid((unsigned)((chunks + bx - 1) / bx), (unsigned)((rows + by - 1) / by), (unsigned)n_frames);
    lowres_init_kernel<<<grid, block, 0, st>>>(job);
    XV_LAUNCH_CHECK();
    return 0;
}

// [x264] x264_frame_copy_picture (luma) + x264_frame_expand_border_mod16 + the extra
// duplicated column/row of x264_frame_init_lowres, as one clamped copy.
__global__ void __launch_bounds__(256)
luma_pad_kernel(LumaPadJob job)
{
    const int chunks = (job.luma_w + 1 + 15) >> 4;
    const int rows = job.luma_h + 1;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= chunks * rows) return;
    const int row = idx / chunks;
    const int x = (idx - row * chunks) << 4;
    const size_t f = blockIdx.y;
    const uint8_t *s = job.y + f * job.src_frame_bytes + (size_t)min(row, job.h - 1) * job.y_stride;
    uint8_t *d = job.dst + f * job.dst_frame_bytes + (size_t)row * job.dst_stride + x;
    const int n = min(16, job.luma_w + 1 - x);
    if (n == 16 && x + 16 <= job.w && ((job.y_stride | job.dst_stride | (int)(uintptr_t)job.y | (int)(uintptr_t)job.dst) & 15) == 0) {
        *(uint4 *)d = ldg_stream128(s + x);
    } else {
        for (int i = 0; i < n; i++) d[i] = s[min(x + i, job.w - 1)];
    }
}

int launch_luma_pad(cudaStream_t st, const LumaPadJob &job, int n_frames)
{
    const long long total = (long long)((job.luma_w + 1 + 15) >> 4) * (job.luma_h + 1);
    if (total <= 0 || n_frames <= 0) return 0;
    dim3 grid((unsigned)((total + 255) / 256), (unsigned)n_frames);
    luma_pad_kernel<<<grid, 256, 0, st>>>(job);
    XV_LAUNCH_CHECK();
    return 0;
}

// [x264] x264_frame_copy_picture, chroma of a 4:2:0 frame (plane_copy_interleave: planar U, V ->
// the encoder's interleaved NV12 plane) + x264_frame_expand_border_mod16 for that plane (the last
// U/V PAIR replicated to the right, the last row downwards): one clamped gather, 4 pairs per thread.
__global__ void __launch_bounds__(256)
chroma_nv12_pad_kernel(ChromaPadJob job)
{
    const int chunks = (job.luma_w / 2 + 3) >> 2;          // 4 UV pairs (8 bytes) per thread
    const int rows = job.luma_h / 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= chunks * rows) return;
    const int row = idx / chunks;
    const int p0 = (idx - row * chunks) << 2;               // first pair of this thread
    const size_t f = blockIdx.y;
    const int cw = job.w / 2, ch = job.h / 2;
    const int srow = min(row, ch - 1);
    const uint8_t *u = job.u + f * job.src_frame_bytes + (size_t)srow * job.c_stride;
    const uint8_t *v = job.v + f * job.src_frame_bytes + (size_t)srow * job.c_stride;
    uint8_t *d = job.dst + f * job.dst_frame_bytes + (size_t)row * job.dst_stride + 2 * p0;
    const int npairs = min(4, job.luma_w / 2 - p0);
    if (npairs == 4 && p0 + 4 <= cw && (((uintptr_t)u | (uintptr_t)v | (uintptr_t)job.c_stride) & 3) == 0 && (p0 & 3) == 0 &&
        (((uintptr_t)d | (uintptr_t)job.dst_stride) & 7) == 0) {
        const uint32_t uu = ldg_stream32(u + p0), vv = ldg_stream32(v + p0);
        *(uint2 *)d = make_uint2(__byte_perm(uu, vv, 0x5140), __byte_perm(uu, vv, 0x7362));
    } else {
        for (int i = 0; i < npairs; i++) {
            const int sp = min(p0 + i, cw - 1);
            d[2 * i] = u[sp]; d[2 * i + 1] = v[sp];
        }
    }
}

int launch_chroma_nv12_pad(cudaStream_t st, const ChromaPadJob &job, int n_frames)
{
    const long long total = (long long)((job.luma_w / 2 + 3) >> 2) * (job.luma_h / 2);
    if (total <= 0 || n_frames <= 0) return 0;
    dim3 grid((unsigned)((total + 255) / 256), (unsigned)n_frames);
    chroma_nv12_pad_kernel<<<grid, 256, 0, st>>>(job);
    XV_LAUNCH_CHECK();
    return 0;
}

} // namespace xv
