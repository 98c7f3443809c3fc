// SURVEY 8(b) boundary B3: upstream libx264's own function-table shapes on DEVICE pointers, so that a patch to a
// real libx264 (which the reference only links, Makefile:21-23,109; entered at codec.c:1693) is mechanical:
//   x264_mc_functions_t    .frame_init_lowres_core, .mbtree_propagate_cost, .mbtree_propagate_list
//   x264_pixel_function_t  .sad[PIXEL_8x8], .satd[PIXEL_8x8], .sad_x3 / .sad_x4[PIXEL_8x8], .intra_mbcmp_x3_8x8c
// Same argument meaning and order as upstream's C functions (common/mc.c, common/pixel.c); what changes is only
// what a device call needs: a context (stream) in front, device pointers, and batches where upstream calls the
// function once per block (a per-block launch would be all launch latency).  The session (la_host.cu) does not
// go through these: it uses the fused / speculative kernels; these exist for the function-table patch and are
// parity-tested one by one through the C ABI (tests/test_b3_gpu.py).
#include "la_common.cuh"
#include "../../include/x264vfw_cuda.h"

namespace xv {

// ---- [x264] common/mc.c: frame_init_lowres_core -----------------------------------------------------------------
// dst0/h/v/c(x, y) = FILTER of src rows 2y, 2y+1, 2y+2 and columns 2x .. 2x+2; no border, no clamping: like
// upstream, the caller provides the duplicated last row / column.  Thread = 4 output pixels of the four planes.
__global__ void __launch_bounds__(256)
lowres_core_kernel(const uint8_t *src0, uint8_t *dst0, uint8_t *dsth, uint8_t *dstv, uint8_t *dstc,
                   long long src_stride, long long dst_stride, int width, int height)
{
    const int chunks = (width + 3) >> 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= chunks * height) return;
    const int y = idx / chunks, x = (idx - y * chunks) << 2;
    const uint8_t *s0 = src0 + (long long)(2 * y) * src_stride + 2 * x, *s1 = s0 + src_stride, *s2 = s1 + src_stride;
    const int n = min(4, width - x);
    uint32_t o[4] = {0, 0, 0, 0};
    for (int i = 0; i < n; i++) {
        const uint32_t a0 = s0[2 * i], a1 = s0[2 * i + 1], a2 = s0[2 * i + 2], b0 = s1[2 * i], b1 = s1[2 * i + 1], b2 = s1[2 * i + 2];
        const uint32_t c0 = s2[2 * i], c1 = s2[2 * i + 1], c2 = s2[2 * i + 2];
#define B3_FILTER(a, b, c, d) (((((a) + (b) + 1) >> 1) + (((c) + (d) + 1) >> 1) + 1) >> 1)
        o[0] |= B3_FILTER(a0, b0, a1, b1) << (8 * i);
        o[1] |= B3_FILTER(a1, b1, a2, b2) << (8 * i);
        o[2] |= B3_FILTER(b0, c0, b1, c1) << (8 * i);
        o[3] |= B3_FILTER(b1, c1, b2, c2) << (8 * i);
#undef B3_FILTER
    }
    uint8_t *d[4] = {dst0, dsth, dstv, dstc};
    for (int k = 0; k < 4; k++)
        for (int i = 0; i < n; i++) d[k][(long long)y * dst_stride + x + i] = (uint8_t)(o[k] >> (8 * i));
}

// ---- [x264] common/mc.c: mbtree_propagate_cost -------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mbtree_cost_kernel(int16_t *dst, const uint16_t *propagate_in, const uint16_t *intra_costs, const uint16_t *inter_costs,
                   const uint16_t *inv_qscales, float fps, int len)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    const int intra_cost = intra_costs[i];
    const int inter_cost = min(intra_cost, inter_costs[i] & LA_LOWRES_COST_MASK);
    const float propagate_intra = (float)(intra_cost * (int)inv_qscales[i]);
    const float propagate_amount = (float)propagate_in[i] + propagate_intra * fps;
    const float propagate_num = (float)(intra_cost - inter_cost);
    const float propagate_denom = (float)intra_cost;
    dst[i] = (int16_t)min((int)(propagate_amount * propagate_num / propagate_denom + 0.5f), 32767);
}

// ---- [x264] common/mc.c: mbtree_propagate_list ----------------------------------------------------------------
// ref_costs is upstream's uint16 array with per-add saturation at 32767; all addends are >= 0, so the result does
// not depend on the order of the adds: a 16-bit saturating add built on a 32-bit compare-and-swap.
__device__ __forceinline__ void sat_add_u16(uint16_t *base, unsigned idx, int v)
{
    if (!v) return;
    unsigned int *word = (unsigned int *)((uintptr_t)(base + idx) & ~(uintptr_t)3);
    const unsigned sh = ((uintptr_t)(base + idx) & 2) ? 16 : 0;
    unsigned int old = *word, assumed;
    do {
        assumed = old;
        const unsigned cur = (assumed >> sh) & 0xffffu;
        const unsigned nv = min(cur + (unsigned)v, 32767u);
        old = atomicCAS(word, assumed, (assumed & ~(0xffffu << sh)) | (nv << sh));
    } while (old != assumed);
}

__global__ void __launch_bounds__(128)
mbtree_list_kernel(uint16_t *ref_costs, const int16_t *mvs, const int16_t *propagate_amount, const uint16_t *lowres_costs,
                   int bipred_weight, int mb_y, int len, int list, int mb_width, int mb_height)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    const unsigned stride = mb_width, width = mb_width, height = mb_height;
    const int lists_used = lowres_costs[i] >> LA_LOWRES_COST_SHIFT;
    if (!(lists_used & (1 << list))) return;
    int listamount = propagate_amount[i];
    if (lists_used == 3) listamount = (listamount * bipred_weight + 32) >> 6;      // upstream: "Apply bipred weighting"
    int x = mvs[2 * i], y = mvs[2 * i + 1];
    if (!(x | y)) { sat_add_u16(ref_costs, mb_y * stride + i, listamount); return; }
    const unsigned mbx = (unsigned)((x >> 5) + i), mby = (unsigned)((y >> 5) + mb_y);
    const unsigned idx0 = mbx + mby * stride, idx2 = idx0 + stride;
    x &= 31; y &= 31;
    const int w0 = ((32 - y) * (32 - x) * listamount + 512) >> 10, w1 = ((32 - y) * x * listamount + 512) >> 10;
    const int w2 = (y * (32 - x) * listamount + 512) >> 10, w3 = (y * x * listamount + 512) >> 10;
    if (mbx < width - 1 && mby < height - 1) {
        sat_add_u16(ref_costs, idx0, w0); sat_add_u16(ref_costs, idx0 + 1u, w1);
        sat_add_u16(ref_costs, idx2, w2); sat_add_u16(ref_costs, idx2 + 1u, w3);
    } else {
        if (mby < height) {
            if (mbx < width) sat_add_u16(ref_costs, idx0, w0);
            if (mbx + 1 < width) sat_add_u16(ref_costs, idx0 + 1u, w1);
        }
        if (mby + 1 < height) {
            if (mbx < width) sat_add_u16(ref_costs, idx2, w2);
            if (mbx + 1 < width) sat_add_u16(ref_costs, idx2 + 1u, w3);
        }
    }
}

// ---- [x264] common/pixel.c: sad / satd 8x8, sad_x3 / sad_x4 8x8 (batched: one block pair per thread) ------------
__device__ __forceinline__ void load_block8(const uint8_t *p, long long stride, uint2 r[8])
{
#pragma unroll
    for (int y = 0; y < 8; y++) r[y] = load8u(p + y * stride);
}

__global__ void __launch_bounds__(128)
cmp8x8_kernel(int satd, const uint8_t *pix1, long long stride1, const uint8_t *pix2, long long stride2,
              const int *off1, const int *off2, int *scores, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint2 a[8], b[8];
    load_block8(pix1 + off1[i], stride1, a);
    load_block8(pix2 + off2[i], stride2, b);
    scores[i] = satd ? satd8x8_rows(a, b) : sad8x8_rows(a, b);
}

__global__ void __launch_bounds__(128)
sad_xn_8x8_kernel(int nref, const uint8_t *fenc, long long fenc_stride, const int *off_fenc, const uint8_t *ref, long long ref_stride,
                  const int *off_ref, int *scores, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint2 a[8], b[8];
    load_block8(fenc + off_fenc[i], fenc_stride, a);
    for (int k = 0; k < nref; k++) {
        load_block8(ref + off_ref[i * nref + k], ref_stride, b);
        scores[i * nref + k] = sad8x8_rows(a, b);
    }
}

// ---- [x264] common/pixel.c: intra_{sad,satd}_x3_8x8c = predict_8x8c_{dc,h,v} scored against the block ---------------
// One thread per 8x8 block of a plane whose neighbours (row above, column to the left) are valid memory, as in
// the lookahead (padded lowres plane).  res[mb][3] = {DC, H, V} (upstream's order for 8x8c).
__global__ void __launch_bounds__(64)
intra_x3_8x8c_kernel(int satd, const uint8_t *plane, int stride, int mb_w, int mb_h, int *res)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= mb_w * mb_h) return;
    const int mx = idx % mb_w, my = idx / mb_w;
    const uint8_t *src = plane + 8 * (mx + (long long)my * stride);
    uint2 s[8], pr[8];
    int left[8], top[8];
    load_block8(src, stride, s);
    const uint2 t = load8u(src - stride);
#pragma unroll
    for (int i = 0; i < 8; i++) { left[i] = src[(long long)i * stride - 1]; top[i] = px_of(t, i); }
    int s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) { s0 += top[i]; s1 += top[i + 4]; s2 += left[i]; s3 += left[i + 4]; }
    const uint32_t d0 = ((s0 + s2 + 4) >> 3) * 0x01010101u, d1 = ((s1 + 2) >> 2) * 0x01010101u;
    const uint32_t d2 = ((s3 + 2) >> 2) * 0x01010101u, d3 = ((s1 + s3 + 4) >> 3) * 0x01010101u;
#pragma unroll
    for (int y = 0; y < 8; y++) pr[y] = y < 4 ? make_uint2(d0, d1) : make_uint2(d2, d3);
    res[3 * idx + 0] = mbcmp_rows(satd, s, pr);
#pragma unroll
    for (int y = 0; y < 8; y++) pr[y] = make_uint2((uint32_t)left[y] * 0x01010101u, (uint32_t)left[y] * 0x01010101u);
    res[3 * idx + 1] = mbcmp_rows(satd, s, pr);
#pragma unroll
    for (int y = 0; y < 8; y++) pr[y] = t;
    res[3 * idx + 2] = mbcmp_rows(satd, s, pr);
}

} // namespace xv

using namespace xv;

#define B3_CTX()                                                       \
    if (!ctx) { set_error("null context"); return -1; }                \
    const Ctx *c = (const Ctx *)ctx;                           \
    XV_CUDA_OK(cudaSetDevice(c->device))

extern "C" {

int x264vfw_cuda_frame_init_lowres_core(x264vfw_cuda_ctx *ctx, const uint8_t *src0, uint8_t *dst0, uint8_t *dsth, uint8_t *dstv, uint8_t *dstc,
                                        intptr_t src_stride, intptr_t dst_stride, int width, int height)
{
    B3_CTX();
    if (!src0 || !dst0 || !dsth || !dstv || !dstc || width <= 0 || height <= 0) { set_error("frame_init_lowres_core: bad argument"); return -1; }
    const long long total = (long long)((width + 3) >> 2) * height;
    lowres_core_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(src0, dst0, dsth, dstv, dstc, src_stride, dst_stride, width, height);
    XV_LAUNCH_CHECK();
    return 0;
}

int x264vfw_cuda_mbtree_propagate_cost(x264vfw_cuda_ctx *ctx, int16_t *dst, const uint16_t *propagate_in, const uint16_t *intra_costs,
                                       const uint16_t *inter_costs, const uint16_t *inv_qscales, float fps_factor, int len)
{
    B3_CTX();
    if (len <= 0) return 0;
    mbtree_cost_kernel<<<(len + 255) / 256, 256, 0, c->stream>>>(dst, propagate_in, intra_costs, inter_costs, inv_qscales, fps_factor, len);
    XV_LAUNCH_CHECK();
    return 0;
}

int x264vfw_cuda_mbtree_propagate_list(x264vfw_cuda_ctx *ctx, uint16_t *ref_costs, const int16_t (*mvs)[2], const int16_t *propagate_amount,
                                       const uint16_t *lowres_costs, int bipred_weight, int mb_y, int len, int list, int mb_width, int mb_height)
{
    B3_CTX();
    if (len <= 0) return 0;
    if (((uintptr_t)ref_costs & 3)) { set_error("mbtree_propagate_list: ref_costs must be 4-byte aligned"); return -1; }
    mbtree_list_kernel<<<(len + 127) / 128, 128, 0, c->stream>>>(ref_costs, (const int16_t *)mvs, propagate_amount, lowres_costs, bipred_weight,
                                                               mb_y, len, list, mb_width, mb_height);
    XV_LAUNCH_CHECK();
    return 0;
}

int x264vfw_cuda_pixel_cmp_8x8(x264vfw_cuda_ctx *ctx, int b_satd, const uint8_t *pix1, intptr_t stride1, const uint8_t *pix2, intptr_t stride2,
                               const int *off1, const int *off2, int *scores, int n)
{
    B3_CTX();
    if (n <= 0) return 0;
    cmp8x8_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(b_satd, pix1, stride1, pix2, stride2, off1, off2, scores, n);
    XV_LAUNCH_CHECK();
    return 0;
}

int x264vfw_cuda_pixel_sad_xn_8x8(x264vfw_cuda_ctx *ctx, int n_ref, const uint8_t *fenc, intptr_t fenc_stride, const int *off_fenc,
                                  const uint8_t *ref, intptr_t ref_stride, const int *off_ref, int *scores, int n)
{
    B3_CTX();
    if (n_ref != 3 && n_ref != 4) { set_error("sad_xn: n_ref must be 3 or 4"); return -1; }
    if (n <= 0) return 0;
    sad_xn_8x8_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(n_ref, fenc, fenc_stride, off_fenc, ref, ref_stride, off_ref, scores, n);
    XV_LAUNCH_CHECK();
    return 0;
}

int x264vfw_cuda_intra_mbcmp_x3_8x8c(x264vfw_cuda_ctx *ctx, int b_satd, const uint8_t *plane, int stride, int mb_w, int mb_h, int *res)
{
    B3_CTX();
    if (mb_w <= 0 || mb_h <= 0) return 0;
    intra_x3_8x8c_kernel<<<(mb_w * mb_h + 63) / 64, 64, 0, c->stream>>>(b_satd, plane, stride, mb_w, mb_h, res);
    XV_LAUNCH_CHECK();
    return 0;
}

} // extern "C"
