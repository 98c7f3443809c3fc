// Colour-space conversion kernels (stage 1 of the hot path).
//
// Replaces the scalar converters of the reference's csp.c behind the same function table:
//   RGB_TO_I420   csp.c:299-388  (bgr/bgra -> I420, 4 matrix x range variants, csp.c:428-435)
//   YYUV_TO_I420  csp.c:155-207  (yuyv/uyvy -> I420, chroma = rounded mean of the row pair)
//   YYUV_TO_I422  csp.c:209-250  (yuyv/uyvy -> I422, pure de-interleave)
//   YUV_TO_YUV    csp.c:99-128   (plane copy / U-V swap / v2 / hv2 subsample, +vflip 75-91)
//   NV_TO_NV, RGB_TO_RGB csp.c:130-153, 390-407 (row copies)
//
// All of these are pure streaming byte kernels: read the source once with 128/64/32-bit
// coalesced loads (lane-contiguous, L1::no_allocate), do the exact integer arithmetic of
// the reference, write every destination byte once.  The bottom-up DIB flip is a negative
// source stride folded into the addressing on the host side (no separate flip pass).
// grid.y indexes the frame so one launch converts a whole batch.
#include "common.cuh"
#include "csp_kernels.h"
#include "rgb_math.cuh"

namespace xv {

// ------------------------------------------------------------------------------------------
// RGB -> 4:2:0.  One thread = 4 pixels x 2 rows (one 16-byte BGRA load per row, or 12 bytes
// of BGR24), producing 2x4 Y, 2 U, 2 V.
// ------------------------------------------------------------------------------------------
template <int BPP, bool VEC>
__device__ __forceinline__ void load_px4(const uint8_t *p, int npx, uint32_t px[4])
{
    if (VEC && npx == 4) {
        if (BPP == 4) {
            uint4 v = ldg_stream128(p);
            px[0] = v.x; px[1] = v.y; px[2] = v.z; px[3] = v.w;
        } else {
            // three 32-bit loads at stride 12: each touches every sector of the warp's 384-byte span, so
            // these go THROUGH L1 (the second and third hit there) instead of fetching the span from L2
            // three times
            const uint32_t *q = (const uint32_t *)p;
            uint32_t w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2);
            // byte 3 of each word is junk: it only ever meets a zero coefficient / an unused lane
            px[0] = w0;
            px[1] = __byte_perm(w0, w1, 0x3543);
            px[2] = __byte_perm(w1, w2, 0x3432);
            px[3] = w2 >> 8;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            px[i] = 0;
            if (i < npx)
                px[i] = (uint32_t)p[i * BPP] | ((uint32_t)p[i * BPP + 1] << 8) | ((uint32_t)p[i * BPP + 2] << 16);
        }
    }
}

// Fast path: width % 4 == 0 and every pointer/stride aligned (checked on the host).  One thread
// converts a 4-pixel column chunk of TWO row pairs (4 source rows): four independent 128-bit
// loads in flight per thread, address arithmetic shared between the pairs.
template <int BPP, bool NV12>
__global__ void __launch_bounds__(256)
rgb_to_420_fast_kernel(RgbJob job)
{
    const int chunk = blockIdx.x * blockDim.x + threadIdx.x;
    const int pair0 = 2 * (blockIdx.y * blockDim.y + threadIdx.y);
    const int npair = job.h >> 1;
    if (chunk >= (job.w >> 2) || pair0 >= npair) return;
    const int x = chunk << 2;
    const size_t f = blockIdx.z;
    const bool two = pair0 + 1 < npair;

    const uint8_t *s = job.src + f * job.src_frame_bytes + (ptrdiff_t)(2 * pair0) * job.src_stride + x * BPP;
    uint32_t r[4][4];
    load_px4<BPP, true>(s, 4, r[0]);
    load_px4<BPP, true>(s + job.src_stride, 4, r[1]);
    if (two) {
        load_px4<BPP, true>(s + 2 * job.src_stride, 4, r[2]);
        load_px4<BPP, true>(s + 3 * job.src_stride, 4, r[3]);
    }
    uint8_t *dy = job.dst_y + f * job.dst_frame_bytes + (size_t)(2 * pair0) * job.y_stride + x;
    uint8_t *du = job.dst_u + f * job.dst_frame_bytes + (size_t)pair0 * job.u_stride;
    uint8_t *dv = job.dst_v + f * job.dst_frame_bytes + (size_t)pair0 * job.v_stride;
#pragma unroll
    for (int p = 0; p < 2; p++) {
        if (p == 1 && !two) break;
        const uint32_t *t = r[2 * p], *b = r[2 * p + 1];
        const uint32_t yt = __byte_perm(__byte_perm(luma16(job.k, t[0]), luma16(job.k, t[1]), 0x0073),
                                        __byte_perm(luma16(job.k, t[2]), luma16(job.k, t[3]), 0x0073), 0x5410);
        const uint32_t yb = __byte_perm(__byte_perm(luma16(job.k, b[0]), luma16(job.k, b[1]), 0x0073),
                                        __byte_perm(luma16(job.k, b[2]), luma16(job.k, b[3]), 0x0073), 0x5410);
        uint32_t u0, v0, u1, v1;
        chroma_quad(job.k, t[0], b[0], t[1], b[1], u0, v0);
        chroma_quad(job.k, t[2], b[2], t[3], b[3], u1, v1);
        *(uint32_t *)(dy + (size_t)(2 * p) * job.y_stride) = yt;
        *(uint32_t *)(dy + (size_t)(2 * p + 1) * job.y_stride) = yb;
        if (NV12) {
            *(uint32_t *)(du + (size_t)p * job.u_stride + x) = u0 | (v0 << 8) | (u1 << 16) | (v1 << 24);
        } else {
            *(uint16_t *)(du + (size_t)p * job.u_stride + (x >> 1)) = (uint16_t)(u0 | (u1 << 8));
            *(uint16_t *)(dv + (size_t)p * job.v_stride + (x >> 1)) = (uint16_t)(v0 | (v1 << 8));
        }
    }
}

// General path: any even width, any alignment (byte loads/stores).  Same arithmetic.
template <int BPP, bool NV12>
__global__ void __launch_bounds__(256)
rgb_to_420_kernel(RgbJob job)
{
    const int chunk = blockIdx.x * blockDim.x + threadIdx.x;
    const int pair = blockIdx.y * blockDim.y + threadIdx.y;
    const int nchunk = (job.w + 3) >> 2;
    if (chunk >= nchunk || pair >= (job.h >> 1)) return;
    const int x = chunk << 2;
    const int npx = min(4, job.w - x);
    const size_t f = blockIdx.z;

    const uint8_t *s0 = job.src + f * job.src_frame_bytes + (ptrdiff_t)(2 * pair) * job.src_stride + x * BPP;
    const uint8_t *s1 = s0 + job.src_stride;
    uint32_t t[4], b[4];
    load_px4<BPP, false>(s0, npx, t);
    load_px4<BPP, false>(s1, npx, b);
    const uint32_t yt = __byte_perm(__byte_perm(luma16(job.k, t[0]), luma16(job.k, t[1]), 0x0073),
                                    __byte_perm(luma16(job.k, t[2]), luma16(job.k, t[3]), 0x0073), 0x5410);
    const uint32_t yb = __byte_perm(__byte_perm(luma16(job.k, b[0]), luma16(job.k, b[1]), 0x0073),
                                    __byte_perm(luma16(job.k, b[2]), luma16(job.k, b[3]), 0x0073), 0x5410);
    uint32_t u0, v0, u1, v1;
    chroma_quad(job.k, t[0], b[0], t[1], b[1], u0, v0);
    chroma_quad(job.k, t[2], b[2], t[3], b[3], u1, v1);

    uint8_t *dy = job.dst_y + f * job.dst_frame_bytes + (size_t)(2 * pair) * job.y_stride + x;
    uint8_t *du = job.dst_u + f * job.dst_frame_bytes + (size_t)pair * job.u_stride;
    uint8_t *dv = job.dst_v + f * job.dst_frame_bytes + (size_t)pair * job.v_stride;
    for (int i = 0; i < npx; i++) {
        dy[i] = (uint8_t)(yt >> (8 * i));
        dy[job.y_stride + i] = (uint8_t)(yb >> (8 * i));
    }
    for (int q = 0; q < (npx >> 1); q++) {
        uint32_t u = q ? u1 : u0, v = q ? v1 : v0;
        if (NV12) { du[x + 2 * q] = (uint8_t)u; du[x + 2 * q + 1] = (uint8_t)v; }
        else      { du[(x >> 1) + q] = (uint8_t)u; dv[(x >> 1) + q] = (uint8_t)v; }
    }
}

int launch_rgb_to_420(cudaStream_t st, const RgbJob &job_in, int bpp, bool nv12, bool vec, int n_frames)
{
    RgbJob job = job_in;
    const int nchunk = (job.w + 3) >> 2, npair = job.h >> 1;
    if (nchunk <= 0 || npair <= 0 || n_frames <= 0) return 0;
    job.k = make_rgb_kernel_coef(job.c);
    const dim3 block(128, 2);
    if (vec && (job.w & 3) == 0) {
        const dim3 grid((unsigned)(((job.w >> 2) + 127) / 128), (unsigned)((npair + 3) / 4), (unsigned)n_frames);
        if (bpp == 4) { if (nv12) rgb_to_420_fast_kernel<4, true><<<grid, block, 0, st>>>(job); else rgb_to_420_fast_kernel<4, false><<<grid, block, 0, st>>>(job); }
        else          { if (nv12) rgb_to_420_fast_kernel<3, true><<<grid, block, 0, st>>>(job); else rgb_to_420_fast_kernel<3, false><<<grid, block, 0, st>>>(job); }
    } else {
        const dim3 grid((unsigned)((nchunk + 127) / 128), (unsigned)((npair + 1) / 2), (unsigned)n_frames);
        if (bpp == 4) { if (nv12) rgb_to_420_kernel<4, true><<<grid, block, 0, st>>>(job); else rgb_to_420_kernel<4, false><<<grid, block, 0, st>>>(job); }
        else          { if (nv12) rgb_to_420_kernel<3, true><<<grid, block, 0, st>>>(job); else rgb_to_420_kernel<3, false><<<grid, block, 0, st>>>(job); }
    }
    XV_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------
// Packed 4:2:2 (YUYV / UYVY) -> planar.  One thread = 8 pixels (16 source bytes) of a row pair.
//   MODE 0: I420 (chroma = (top+bottom+1)>>1, csp.c:185-186)
//   MODE 1: I422 (copy, csp.c:239-240)
//   MODE 2: I444 extension (I422 samples, each chroma sample written twice)
// ------------------------------------------------------------------------------------------
template <bool UYVY>
__device__ __forceinline__ void split422(uint4 v, uint2 &y, uint32_t &u, uint32_t &vv)
{
    // YUYV bytes: Y0 U0 Y1 V0 | Y2 U1 Y3 V1 ...   UYVY: U0 Y0 V0 Y1 | ...
    if (!UYVY) {
        y.x = __byte_perm(v.x, v.y, 0x6420);
        y.y = __byte_perm(v.z, v.w, 0x6420);
        uint32_t c0 = __byte_perm(v.x, v.y, 0x7531);  // U0 V0 U1 V1
        uint32_t c1 = __byte_perm(v.z, v.w, 0x7531);  // U2 V2 U3 V3
        u  = __byte_perm(c0, c1, 0x6420);
        vv = __byte_perm(c0, c1, 0x7531);
    } else {
        y.x = __byte_perm(v.x, v.y, 0x7531);
        y.y = __byte_perm(v.z, v.w, 0x7531);
        uint32_t c0 = __byte_perm(v.x, v.y, 0x6420);
        uint32_t c1 = __byte_perm(v.z, v.w, 0x6420);
        u  = __byte_perm(c0, c1, 0x6420);
        vv = __byte_perm(c0, c1, 0x7531);
    }
}

template <bool UYVY>
__device__ __forceinline__ uint4 load_422_row(const uint8_t *p, int npx, bool vec)
{
    if (vec && npx == 8) return ldg_stream128(p);
    uint32_t w[4] = {0, 0, 0, 0};
    for (int i = 0; i < 2 * npx; i++) w[i >> 2] |= (uint32_t)p[i] << (8 * (i & 3));
    return make_uint4(w[0], w[1], w[2], w[3]);
}

template <bool UYVY, int MODE, bool VEC>
__global__ void __launch_bounds__(256)
packed422_kernel(PackedJob job)
{
    // every mode works on a ROW PAIR per thread (height is even): two independent 128-bit loads in
    // flight, address arithmetic shared; only MODE 0 combines the two rows
    const int nchunk = (job.w + 7) >> 3;
    const int nrow = job.h >> 1;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nchunk * nrow) return;
    const int row = idx / nchunk;
    const int chunk = idx - row * nchunk;
    const int x = chunk << 3;
    const int npx = min(8, job.w - x);
    const size_t f = blockIdx.y;
    const bool vec = VEC;

    const int srow = 2 * row;
    const uint8_t *s0 = job.src + f * job.src_frame_bytes + (ptrdiff_t)srow * job.src_stride + 2 * x;
    uint2 y0, y1;
    uint32_t u, v, ub, vb;
    const uint4 in0 = load_422_row<UYVY>(s0, npx, vec), in1 = load_422_row<UYVY>(s0 + job.src_stride, npx, vec);
    split422<UYVY>(in0, y0, u, v);
    split422<UYVY>(in1, y1, ub, vb);
    if (MODE == 0) {
        u = avg4(u, ub);
        v = avg4(v, vb);
    }

    const int crow = MODE == 0 ? row : srow;               // first chroma row written by this thread
    uint8_t *dy = job.dst_y + f * job.dst_frame_bytes + (size_t)srow * job.y_stride + x;
    uint8_t *du = job.dst_u + f * job.dst_frame_bytes + (size_t)crow * job.u_stride;
    uint8_t *dv = job.dst_v + f * job.dst_frame_bytes + (size_t)crow * job.v_stride;
    if (vec && npx == 8) {
        *(uint2 *)dy = y0;
        *(uint2 *)(dy + job.y_stride) = y1;
        if (MODE == 2) {
            uint2 uu, v2;
            uu.x = __byte_perm(u, 0, 0x1100); uu.y = __byte_perm(u, 0, 0x3322);
            v2.x = __byte_perm(v, 0, 0x1100); v2.y = __byte_perm(v, 0, 0x3322);
            *(uint2 *)(du + x) = uu;
            *(uint2 *)(dv + x) = v2;
            uu.x = __byte_perm(ub, 0, 0x1100); uu.y = __byte_perm(ub, 0, 0x3322);
            v2.x = __byte_perm(vb, 0, 0x1100); v2.y = __byte_perm(vb, 0, 0x3322);
            *(uint2 *)(du + job.u_stride + x) = uu;
            *(uint2 *)(dv + job.v_stride + x) = v2;
        } else {
            *(uint32_t *)(du + (x >> 1)) = u;
            *(uint32_t *)(dv + (x >> 1)) = v;
            if (MODE == 1) {
                *(uint32_t *)(du + job.u_stride + (x >> 1)) = ub;
                *(uint32_t *)(dv + job.v_stride + (x >> 1)) = vb;
            }
        }
    } else {
        for (int i = 0; i < npx; i++) {
            dy[i] = (uint8_t)((i < 4 ? y0.x : y0.y) >> (8 * (i & 3)));
            dy[job.y_stride + i] = (uint8_t)((i < 4 ? y1.x : y1.y) >> (8 * (i & 3)));
        }
        for (int i = 0; i < (npx >> 1); i++) {
            const uint8_t u0 = (uint8_t)(u >> (8 * i)), v0 = (uint8_t)(v >> (8 * i));
            const uint8_t u1 = (uint8_t)(ub >> (8 * i)), v1 = (uint8_t)(vb >> (8 * i));
            if (MODE == 2) {
                du[x + 2 * i] = du[x + 2 * i + 1] = u0; dv[x + 2 * i] = dv[x + 2 * i + 1] = v0;
                du[job.u_stride + x + 2 * i] = du[job.u_stride + x + 2 * i + 1] = u1; dv[job.v_stride + x + 2 * i] = dv[job.v_stride + x + 2 * i + 1] = v1;
            } else {
                du[(x >> 1) + i] = u0; dv[(x >> 1) + i] = v0;
                if (MODE == 1) { du[job.u_stride + (x >> 1) + i] = u1; dv[job.v_stride + (x >> 1) + i] = v1; }
            }
        }
    }
}

int launch_packed422(cudaStream_t st, const PackedJob &job, bool uyvy, int mode, bool vec, int n_frames)
{
    const int nchunk = (job.w + 7) >> 3;
    const int nrow = job.h >> 1;                        // a row pair per thread in every mode
    const long long total = (long long)nchunk * nrow;
    if (total <= 0 || n_frames <= 0) return 0;
    dim3 grid((unsigned)((total + 255) / 256), (unsigned)n_frames);
#define XV_P(U, M, V) packed422_kernel<U, M, V><<<grid, 256, 0, st>>>(job)
#define XV_PM(U, V) do { if (mode == 0) XV_P(U, 0, V); else if (mode == 1) XV_P(U, 1, V); else XV_P(U, 2, V); } while (0)
    if (uyvy) { if (vec) XV_PM(true, true); else XV_PM(true, false); }
    else      { if (vec) XV_PM(false, true); else XV_PM(false, false); }
#undef XV_PM
#undef XV_P
    XV_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------
// Planar helpers: copy / vertical 2:1 / 2x2 subsample of up to 3 planes in one launch
// (csp.c:28-91).  One thread = 16 destination bytes of one row of one plane.
//   op 0: copy                       dst[x] = s[x]
//   op 1: subsamplev2  (csp.c:39-55) dst[x] = (s[x] + s[x+stride] + 1) >> 1
//   op 2: subsamplehv2 (csp.c:57-73) dst[x] = (s[2x]+s[2x+1]+s[2x+stride]+s[2x+1+stride]+2) >> 2
// ------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(256)
planes_kernel(PlanesJob job)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t f = blockIdx.y;
    int pl = 0, local = idx;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        if (pl == i && i < job.n && local >= job.p[i].nthreads) { local -= job.p[i].nthreads; pl = i + 1; }
    }
    if (pl >= job.n) return;
    const PlaneOp &p = job.p[pl];
    const int nchunk = (p.w + 15) >> 4;
    const int row = local / nchunk;
    const int x = (local - row * nchunk) << 4;
    const int n = min(16, p.w - x);
    uint8_t *d = p.dst + f * job.dst_frame_bytes + (size_t)row * p.dst_stride + x;
    if (p.op == 0) {
        const uint8_t *s = p.src + f * job.src_frame_bytes + (ptrdiff_t)row * p.src_stride + x;
        if (VEC && n == 16) *(uint4 *)d = ldg_stream128(s);
        else for (int i = 0; i < n; i++) d[i] = s[i];
    } else if (p.op == 1) {
        const uint8_t *s = p.src + f * job.src_frame_bytes + (ptrdiff_t)(2 * row) * p.src_stride + x;
        if (VEC && n == 16) {
            uint4 a = ldg_stream128(s), b = ldg_stream128(s + p.src_stride);
            *(uint4 *)d = make_uint4(avg4(a.x, b.x), avg4(a.y, b.y), avg4(a.z, b.z), avg4(a.w, b.w));
        } else for (int i = 0; i < n; i++) d[i] = (uint8_t)((s[i] + s[i + p.src_stride] + 1) >> 1);
    } else {
        const uint8_t *s = p.src + f * job.src_frame_bytes + (ptrdiff_t)(2 * row) * p.src_stride + 2 * x;
        if (VEC && n == 16) {
            uint32_t o[4];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                uint4 a = ldg_stream128(s + 16 * h), b = ldg_stream128(s + 16 * h + p.src_stride);
                uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    uint32_t r = 0;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        uint32_t wa = aw[2 * k + (j >> 1)], wb = bw[2 * k + (j >> 1)];
                        int sh = 16 * (j & 1);
                        uint32_t sum = ((wa >> sh) & 0xff) + ((wa >> (sh + 8)) & 0xff) +
                                       ((wb >> sh) & 0xff) + ((wb >> (sh + 8)) & 0xff) + 2;
                        r |= (sum >> 2) << (8 * j);
                    }
                    o[2 * h + k] = r;
                }
            }
            *(uint4 *)d = make_uint4(o[0], o[1], o[2], o[3]);
        } else for (int i = 0; i < n; i++)
            d[i] = (uint8_t)((s[2 * i] + s[2 * i + 1] + s[2 * i + p.src_stride] + s[2 * i + 1 + p.src_stride] + 2) >> 2);
    }
}

int launch_planes(cudaStream_t st, PlanesJob &job, bool vec, int n_frames)
{
    long long total = 0;
    for (int i = 0; i < job.n; i++) {
        job.p[i].nthreads = ((job.p[i].w + 15) >> 4) * job.p[i].h;
        total += job.p[i].nthreads;
    }
    if (total <= 0 || n_frames <= 0) return 0;
    dim3 grid((unsigned)((total + 255) / 256), (unsigned)n_frames);
    if (vec) planes_kernel<true><<<grid, 256, 0, st>>>(job);
    else     planes_kernel<false><<<grid, 256, 0, st>>>(job);
    XV_LAUNCH_CHECK();
    return 0;
}

} // namespace xv
