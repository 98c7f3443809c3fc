// C ABI of the colour-space / lowres stages (include/x264vfw_cuda.h).
//
// Host-side mirror of the reference's dispatch: x264vfw_csp_init (csp.c:440-514) picks a
// converter from (encoder csp, colmatrix==1, fullrange); x264vfw_img_fill (codec.c:304-379)
// describes the borrowed input buffer; conv_pic comes from [x264] x264_picture_alloc
// (codec.c:1673).  Here the same decisions select a CUDA kernel + launch descriptor.
#include "common.cuh"
#include <mutex>
#include <math.h>
#include "csp_kernels.h"
#include "frontend.h"
#include "rgb_math.cuh"
#define XV_DEVICE static inline      /* host view of hpel_kernel.cuh: the job descriptor only */
#define XV_HPEL_HOST_ONLY
#include "hpel_kernel.cuh"
#include "../../include/x264vfw_cuda.h"
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>

namespace xv {

static thread_local char t_err[512] = "";
std::atomic<uint64_t> g_launch_count{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

// csp.c:252-297 evaluated; index = (colmatrix==1)*2 + fullrange
static const RgbCoef k_rgb_coef[4] = {
    /* 601 tv */ {269262, 528618, 102662, 17301504, 155423, 305128, 460551, 538968064, 460551, 385654, 74897, 538968064},
    /* 601 pc */ {313524, 615514, 119538, 524288, 176932, 347356, 524288, 538968063, 524288, 439026, 85262, 538968063},
    /* 709 tv */ {191455, 644067, 65019, 17301504, 105533, 355018, 460551, 538968064, 460551, 418321, 42230, 538968064},
    /* 709 pc */ {222927, 749942, 75707, 524288, 120138, 404150, 524288, 538968063, 524288, 476214, 48074, 538968063},
};

static inline bool al(const void *p, size_t a) { return ((uintptr_t)p & (a - 1)) == 0; }
static inline bool als(long long v, long long a) { return (v & (a - 1)) == 0; }

// Number of rows of source plane i for an input csp (x264vfw_img_fill geometry).
static int src_plane_rows(int csp, int i, int h)
{
    switch (csp) {
    case X264VFW_CUDA_CSP_I420: case X264VFW_CUDA_CSP_YV12: return i ? ((h + 1) & ~1) / 2 : ((h + 1) & ~1);
    case X264VFW_CUDA_CSP_NV12: return i ? ((h + 1) & ~1) / 2 : ((h + 1) & ~1);
    default: return h;
    }
}

static int dst_plane_rows(int out_csp, int i, int h)
{
    switch (out_csp) {
    case X264VFW_CUDA_OUT_I420: case X264VFW_CUDA_OUT_NV12: return i ? h / 2 : h;
    default: return h;
    }
}

// Select + launch.  All pointers are DEVICE pointers.  Returns 0, or -1 for an unsupported
// pair (== convert_fail, csp.c:93-97) / launch failure.
static int convert_device(cudaStream_t st, int out_csp, int colmatrix, int fullrange, int ext,
                          const x264vfw_cuda_image_t *dst, const x264vfw_cuda_image_t *src,
                          int w, int h, size_t sfb, size_t dfb, int n_frames)
{
    const int in = src->i_csp & X264VFW_CUDA_CSP_MASK;
    const bool flip = (src->i_csp & X264VFW_CUDA_CSP_VFLIP) != 0;
    if (w <= 0 || h <= 0 || (w & 1) || (h & 1)) { set_error("width/height must be positive and even (codec.c:639)"); return -1; }

    // ---- RGB -> I420 (csp.c:456-487) and the NV12 extension -------------------------------
    if ((in == X264VFW_CUDA_CSP_BGR || in == X264VFW_CUDA_CSP_BGRA) &&
        (out_csp == X264VFW_CUDA_OUT_I420 || (out_csp == X264VFW_CUDA_OUT_NV12 && ext == X264VFW_CUDA_EXT_RGB_TO_NV12))) {
        const int bpp = in == X264VFW_CUDA_CSP_BGRA ? 4 : 3;
        const bool nv12 = out_csp == X264VFW_CUDA_OUT_NV12;
        RgbJob j;
        j.src_stride = src->i_stride[0];
        j.src = src->plane[0];
        if (flip) { j.src += (ptrdiff_t)(h - 1) * j.src_stride; j.src_stride = -j.src_stride; }   // csp.c:310-314
        j.dst_y = dst->plane[0]; j.y_stride = dst->i_stride[0];
        j.dst_u = dst->plane[1]; j.u_stride = dst->i_stride[1];
        j.dst_v = nv12 ? nullptr : dst->plane[2]; j.v_stride = nv12 ? 0 : dst->i_stride[2];
        j.w = w; j.h = h; j.src_frame_bytes = sfb; j.dst_frame_bytes = dfb;
        j.c = k_rgb_coef[(colmatrix == 1 ? 2 : 0) + (fullrange ? 1 : 0)];
        const size_t sa = bpp == 4 ? 16 : 4;
        bool vec = al(src->plane[0], sa) && als(src->i_stride[0], sa) && als((long long)sfb, sa) &&
                   al(j.dst_y, 4) && als(j.y_stride, 4) && als((long long)dfb, 4) &&
                   (nv12 ? (al(j.dst_u, 4) && als(j.u_stride, 4))
                         : (al(j.dst_u, 2) && al(j.dst_v, 2) && als(j.u_stride, 2) && als(j.v_stride, 2)));
        return launch_rgb_to_420(st, j, bpp, nv12, vec, n_frames);
    }

    // ---- packed 4:2:2 (csp.c:454-455, 497-498) and the I444 extension ---------------------
    if (in == X264VFW_CUDA_CSP_YUYV || in == X264VFW_CUDA_CSP_UYVY) {
        int mode;
        if (out_csp == X264VFW_CUDA_OUT_I420) mode = 0;
        else if (out_csp == X264VFW_CUDA_OUT_I422) mode = 1;
        else if (out_csp == X264VFW_CUDA_OUT_I444 && ext == X264VFW_CUDA_EXT_422_TO_I444) mode = 2;
        else return -1;
        PackedJob j;
        j.src_stride = src->i_stride[0];
        j.src = src->plane[0];
        if (flip) { j.src += (ptrdiff_t)(h - 1) * j.src_stride; j.src_stride = -j.src_stride; }   // csp.c:166-170
        j.dst_y = dst->plane[0]; j.dst_u = dst->plane[1]; j.dst_v = dst->plane[2];
        j.y_stride = dst->i_stride[0]; j.u_stride = dst->i_stride[1]; j.v_stride = dst->i_stride[2];
        j.w = w; j.h = h; j.src_frame_bytes = sfb; j.dst_frame_bytes = dfb;
        const size_t ca = mode == 2 ? 8 : 4;
        bool vec = al(src->plane[0], 16) && als(src->i_stride[0], 16) && als((long long)sfb, 16) &&
                   al(j.dst_y, 8) && als(j.y_stride, 8) && als((long long)dfb, 8) &&
                   al(j.dst_u, ca) && al(j.dst_v, ca) && als(j.u_stride, ca) && als(j.v_stride, ca);
        return launch_packed422(st, j, in == X264VFW_CUDA_CSP_UYVY, mode, vec, n_frames);
    }

    // ---- planar / row-copy family (csp.c:409-420, 437-438) --------------------------------
    PlanesJob pj;
    memset(&pj, 0, sizeof(pj));
    pj.src_frame_bytes = sfb; pj.dst_frame_bytes = dfb;
    auto add = [&](int sp, int dp, int pw, int ph, int op) {
        PlaneOp &p = pj.p[pj.n++];
        const int src_rows = op == 0 ? ph : 2 * ph;
        p.src = src->plane[sp]; p.src_stride = src->i_stride[sp];
        if (flip) { p.src += (ptrdiff_t)(src_rows - 1) * p.src_stride; p.src_stride = -p.src_stride; }   // csp.c:75-91
        p.dst = dst->plane[dp]; p.dst_stride = dst->i_stride[dp];
        p.w = pw; p.h = ph; p.op = op;
    };
    bool ok = false;
    if (out_csp == X264VFW_CUDA_OUT_I420) {
        ok = true;
        switch (in) {
        case X264VFW_CUDA_CSP_I420: add(0, 0, w, h, 0); add(1, 1, w >> 1, h >> 1, 0); add(2, 2, w >> 1, h >> 1, 0); break;
        case X264VFW_CUDA_CSP_YV12: add(0, 0, w, h, 0); add(1, 2, w >> 1, h >> 1, 0); add(2, 1, w >> 1, h >> 1, 0); break;
        case X264VFW_CUDA_CSP_YV16: add(0, 0, w, h, 0); add(1, 2, w >> 1, h >> 1, 1); add(2, 1, w >> 1, h >> 1, 1); break;
        case X264VFW_CUDA_CSP_YV24: add(0, 0, w, h, 0); add(1, 2, w >> 1, h >> 1, 2); add(2, 1, w >> 1, h >> 1, 2); break;
        default: ok = false;
        }
    } else if (out_csp == X264VFW_CUDA_OUT_NV12 && in == X264VFW_CUDA_CSP_NV12) {
        ok = true; add(0, 0, w, h, 0); add(1, 1, w, h >> 1, 0);
    } else if (out_csp == X264VFW_CUDA_OUT_I422 && in == X264VFW_CUDA_CSP_YV16) {
        ok = true; add(0, 0, w, h, 0); add(1, 2, w >> 1, h, 0); add(2, 1, w >> 1, h, 0);
    } else if (out_csp == X264VFW_CUDA_OUT_I444 && in == X264VFW_CUDA_CSP_YV24) {
        ok = true; add(0, 0, w, h, 0); add(1, 2, w, h, 0); add(2, 1, w, h, 0);
    } else if (out_csp == X264VFW_CUDA_OUT_BGR && in == X264VFW_CUDA_CSP_BGR) {
        ok = true; add(0, 0, w * 3, h, 0);
    } else if (out_csp == X264VFW_CUDA_OUT_BGRA && in == X264VFW_CUDA_CSP_BGRA) {
        ok = true; add(0, 0, w * 4, h, 0);
    }
    if (!ok) { set_error("unsupported colour-space pair in=%d out=%d (csp.c:443-444 convert_fail)", in, out_csp); return -1; }
    bool vec = als((long long)sfb, 16) && als((long long)dfb, 16);
    for (int i = 0; i < pj.n; i++)
        vec = vec && al(pj.p[i].src, 16) && al(pj.p[i].dst, 16) && als(pj.p[i].src_stride, 16) && als(pj.p[i].dst_stride, 16);
    return launch_planes(st, pj, vec, n_frames);
}

int convert_device_public(cudaStream_t st, int out_csp, int colmatrix, int fullrange, int ext,
                          const x264vfw_cuda_image_t *dst, const x264vfw_cuda_image_t *src,
                          int w, int h, size_t sfb, size_t dfb, int n_frames)
{
    return convert_device(st, out_csp, colmatrix, fullrange, ext, dst, src, w, h, sfb, dfb, n_frames);
}

static int ensure(uint8_t **p, size_t *cap, size_t need, bool pinned)
{
    if (*cap >= need) return 0;
    if (*p) { if (pinned) cudaFreeHost(*p); else cudaFree(*p); *p = nullptr; *cap = 0; }
    need = (need + (1 << 20) - 1) & ~(size_t)((1 << 20) - 1);
    XV_CUDA_OK(pinned ? cudaMallocHost((void **)p, need) : cudaMalloc((void **)p, need));
    *cap = need;
    return 0;
}

// Host-pointer conversion through staging buffers: H2D per source plane, one kernel,
// D2H per destination plane.  Strides are preserved so the kernel sees the same geometry.
static int convert_host(Ctx *ctx, int out_csp, int colmatrix, int fullrange, int ext,
                        x264vfw_cuda_image_t *dst, x264vfw_cuda_image_t *src, int w, int h)
{
    if (!dst || !src) { set_error("null image"); return -1; }
    const int in = src->i_csp & X264VFW_CUDA_CSP_MASK;
    if (in <= 0 || in >= X264VFW_CUDA_CSP_MAX) { set_error("bad input csp %d", src->i_csp); return -1; }
    if (w <= 0 || h <= 0 || (w & 1) || (h & 1)) { set_error("width/height must be positive and even"); return -1; }
    x264vfw_cuda_image_t sgeo, dgeo;
    if (x264vfw_cuda_img_fill(&sgeo, nullptr, in, w, h) < 0) return -1;
    if (x264vfw_cuda_picture_layout(&dgeo, nullptr, out_csp, w, h) < 0) { set_error("bad encoder csp %d", out_csp); return -1; }
    for (int i = 0; i < sgeo.i_plane; i++) if (src->i_stride[i] <= 0) { set_error("bad source stride"); return -1; }

    XV_CUDA_OK(cudaSetDevice(ctx->device));
    size_t soff[4], doff[4], stot = 0, dtot = 0;
    for (int i = 0; i < sgeo.i_plane; i++) { soff[i] = stot; stot += ((size_t)src->i_stride[i] * src_plane_rows(in, i, h) + 255) & ~(size_t)255; }
    for (int i = 0; i < dgeo.i_plane; i++) { doff[i] = dtot; dtot += ((size_t)dst->i_stride[i] * dst_plane_rows(out_csp, i, h) + 255) & ~(size_t)255; }
    if (ensure(&ctx->d_src, &ctx->d_src_bytes, stot, false) || ensure(&ctx->d_dst, &ctx->d_dst_bytes, dtot, false)) return -1;

    x264vfw_cuda_image_t ds = *src, dd = *dst;
    for (int i = 0; i < sgeo.i_plane; i++) {
        ds.plane[i] = ctx->d_src + soff[i];
        XV_CUDA_OK(cudaMemcpyAsync(ds.plane[i], src->plane[i], (size_t)src->i_stride[i] * src_plane_rows(in, i, h),
                                   cudaMemcpyHostToDevice, ctx->stream));
    }
    for (int i = 0; i < dgeo.i_plane; i++) dd.plane[i] = ctx->d_dst + doff[i];
    if (convert_device(ctx->stream, out_csp, colmatrix, fullrange, ext, &dd, &ds, w, h, 0, 0, 1) < 0) return -1;
    for (int i = 0; i < dgeo.i_plane; i++) {
        // copy back only the bytes the reference converter writes (row width), keeping the
        // caller's stride padding untouched
        const int rows = dst_plane_rows(out_csp, i, h);
        const size_t roww = (size_t)dgeo.i_stride[i];
        XV_CUDA_OK(cudaMemcpy2DAsync(dst->plane[i], dst->i_stride[i], dd.plane[i], dst->i_stride[i],
                                     roww, rows, cudaMemcpyDeviceToHost, ctx->stream));
    }
    XV_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

static thread_local Ctx *t_ctx = nullptr;
static Ctx *thread_ctx()
{
    if (!t_ctx) {
        x264vfw_cuda_ctx *c = nullptr;
        if (x264vfw_cuda_ctx_create(&c, -1) < 0) return nullptr;
        t_ctx = (Ctx *)c;
    }
    return t_ctx;
}

// One C function per table slot variant: the reference's converters carry their variant in
// the function identity (csp.c:428-435), not in an argument.
template <int OUT, int MAT, int FULL>
static int table_entry(x264vfw_cuda_image_t *dst, x264vfw_cuda_image_t *src, int w, int h)
{
    Ctx *c = thread_ctx();
    if (!c) return -1;
    return convert_host(c, OUT, MAT, FULL, X264VFW_CUDA_EXT_NONE, dst, src, w, h);
}
static int table_fail(x264vfw_cuda_image_t *, x264vfw_cuda_image_t *, int, int) { return -1; }   // csp.c:93-97

const RgbCoef &rgb_coef(int colmatrix, int fullrange) { return k_rgb_coef[(colmatrix == 1 ? 2 : 0) + (fullrange ? 1 : 0)]; }

// x264_log2_lut / x264_exp2_lut on the device, one copy per device, for the context-level front end
// (a lookahead session has its own)
int frontend_luts(int device, const float **log2_lut, const uint8_t **exp2_lut)
{
    static std::mutex mu;
    static float *d_l2[64] = {nullptr};
    static uint8_t *d_e2[64] = {nullptr};
    std::lock_guard<std::mutex> lk(mu);
    if (device < 0 || device >= 64) return -1;
    if (!d_l2[device]) {
        float l2[128]; uint8_t e2[64];
        for (int i = 0; i < 128; i++) l2[i] = (float)(round(log2(1.0 + i / 128.0) * 100000.0) / 100000.0);
        for (int i = 0; i < 64; i++) e2[i] = (uint8_t)lround(256.0 * (pow(2.0, i / 64.0) - 1.0));
        XV_CUDA_OK(cudaMalloc((void **)&d_l2[device], sizeof(l2)));
        XV_CUDA_OK(cudaMalloc((void **)&d_e2[device], 64));
        XV_CUDA_OK(cudaMemcpy(d_l2[device], l2, sizeof(l2), cudaMemcpyHostToDevice));
        XV_CUDA_OK(cudaMemcpy(d_e2[device], e2, 64, cudaMemcpyHostToDevice));
    }
    *log2_lut = d_l2[device]; *exp2_lut = d_e2[device];
    return 0;
}

} // namespace xv

using namespace xv;

extern "C" {

const char *x264vfw_cuda_last_error(void) { return t_err; }
const char *x264vfw_cuda_version(void) { return "x264vfw_cuda 0.1 sm_100a"; }
uint64_t x264vfw_cuda_launch_count(void) { return g_launch_count.load(); }

int x264vfw_cuda_ctx_create(x264vfw_cuda_ctx **pctx, int device)
{
    if (!pctx) return -1;
    *pctx = nullptr;
    int ndev = 0;
    XV_CUDA_OK(cudaGetDeviceCount(&ndev));
    if (ndev <= 0) { set_error("no CUDA device: this library has no CPU fallback"); return -1; }
    if (device < 0) XV_CUDA_OK(cudaGetDevice(&device));
    XV_CUDA_OK(cudaSetDevice(device));
    Ctx *c = new Ctx();
    c->device = device;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        set_error("cudaStreamCreate failed"); delete c; return -1;
    }
    *pctx = (x264vfw_cuda_ctx *)c;
    return 0;
}

void x264vfw_cuda_ctx_destroy(x264vfw_cuda_ctx *h)
{
    Ctx *c = (Ctx *)h;
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    if (c->d_src) cudaFree(c->d_src);
    if (c->d_dst) cudaFree(c->d_dst);
    cudaStreamDestroy(c->stream);
    if (t_ctx == c) t_ctx = nullptr;
    delete c;
}

void *x264vfw_cuda_ctx_stream(x264vfw_cuda_ctx *h) { return h ? (void *)((Ctx *)h)->stream : nullptr; }

int x264vfw_cuda_ctx_sync(x264vfw_cuda_ctx *h)
{
    if (!h) return -1;
    XV_CUDA_OK(cudaStreamSynchronize(((Ctx *)h)->stream));
    return 0;
}

int64_t x264vfw_cuda_img_fill(x264vfw_cuda_image_t *img, uint8_t *ptr, int i_csp, int width, int height)
{
    // codec.c:304-379
    if (!img) return -1;
    const int csp = i_csp & X264VFW_CUDA_CSP_MASK;
    memset(img, 0, sizeof(*img));
    img->i_csp = i_csp;
    int64_t rows[3] = {0, 0, 0};
    switch (csp) {
    case X264VFW_CUDA_CSP_I420: case X264VFW_CUDA_CSP_YV12:
        height = (height + 1) & ~1; width = (width + 1) & ~1;
        img->i_plane = 3; img->i_stride[0] = width; img->i_stride[1] = img->i_stride[2] = width / 2;
        rows[0] = height; rows[1] = rows[2] = height / 2; break;
    case X264VFW_CUDA_CSP_YV16:
        width = (width + 1) & ~1;
        img->i_plane = 3; img->i_stride[0] = width; img->i_stride[1] = img->i_stride[2] = width / 2;
        rows[0] = rows[1] = rows[2] = height; break;
    case X264VFW_CUDA_CSP_YV24:
        img->i_plane = 3; img->i_stride[0] = img->i_stride[1] = img->i_stride[2] = width;
        rows[0] = rows[1] = rows[2] = height; break;
    case X264VFW_CUDA_CSP_NV12:
        height = (height + 1) & ~1; width = (width + 1) & ~1;
        img->i_plane = 2; img->i_stride[0] = img->i_stride[1] = width;
        rows[0] = height; rows[1] = height / 2; break;
    case X264VFW_CUDA_CSP_YUYV: case X264VFW_CUDA_CSP_UYVY:
        width = (width + 1) & ~1;
        img->i_plane = 1; img->i_stride[0] = 2 * width; rows[0] = height; break;
    case X264VFW_CUDA_CSP_BGR:
        img->i_plane = 1; img->i_stride[0] = (3 * width + 3) & ~3; rows[0] = height; break;
    case X264VFW_CUDA_CSP_BGRA:
        img->i_plane = 1; img->i_stride[0] = 4 * width; rows[0] = height; break;
    default:
        set_error("img_fill: unknown csp %d", i_csp);
        return -1;
    }
    int64_t off = 0;
    for (int i = 0; i < img->i_plane; i++) {
        img->plane[i] = ptr ? ptr + off : nullptr;
        off += (int64_t)img->i_stride[i] * rows[i];
    }
    return off;
}

int64_t x264vfw_cuda_picture_layout(x264vfw_cuda_image_t *img, uint8_t *ptr, int out_csp, int width, int height)
{
    // [x264] x264_picture_alloc: one contiguous buffer, tight strides
    if (!img) return -1;
    memset(img, 0, sizeof(*img));
    img->i_csp = out_csp;
    int64_t rows[3] = {height, height, height};
    switch (out_csp) {
    case X264VFW_CUDA_OUT_I420: img->i_plane = 3; img->i_stride[0] = width; img->i_stride[1] = img->i_stride[2] = width / 2; rows[1] = rows[2] = height / 2; break;
    case X264VFW_CUDA_OUT_NV12: img->i_plane = 2; img->i_stride[0] = img->i_stride[1] = width; rows[1] = height / 2; break;
    case X264VFW_CUDA_OUT_I422: img->i_plane = 3; img->i_stride[0] = width; img->i_stride[1] = img->i_stride[2] = width / 2; break;
    case X264VFW_CUDA_OUT_I444: img->i_plane = 3; img->i_stride[0] = img->i_stride[1] = img->i_stride[2] = width; break;
    case X264VFW_CUDA_OUT_BGR:  img->i_plane = 1; img->i_stride[0] = 3 * width; break;
    case X264VFW_CUDA_OUT_BGRA: img->i_plane = 1; img->i_stride[0] = 4 * width; break;
    default: return -1;
    }
    int64_t off = 0;
    for (int i = 0; i < img->i_plane; i++) {
        img->plane[i] = ptr ? ptr + off : nullptr;
        off += (int64_t)img->i_stride[i] * rows[i];
    }
    return off;
}

int x264vfw_cuda_csp_convert(x264vfw_cuda_ctx *ctx, int out_csp, int colmatrix, int fullrange, int ext,
                             x264vfw_cuda_image_t *dst, x264vfw_cuda_image_t *src, int w, int h)
{
    if (!ctx) { set_error("null context"); return -1; }
    return convert_host((Ctx *)ctx, out_csp, colmatrix, fullrange, ext, dst, src, w, h);
}

int x264vfw_cuda_csp_convert_batch(x264vfw_cuda_ctx *ctx, int out_csp, int colmatrix, int fullrange, int ext,
                                   const x264vfw_cuda_image_t *dst, const x264vfw_cuda_image_t *src,
                                   int w, int h, size_t sfb, size_t dfb, int n_frames)
{
    if (!ctx || !dst || !src) { set_error("null argument"); return -1; }
    Ctx *c = (Ctx *)ctx;
    XV_CUDA_OK(cudaSetDevice(c->device));
    return convert_device(c->stream, out_csp, colmatrix, fullrange, ext, dst, src, w, h, sfb, dfb, n_frames);
}

void x264vfw_cuda_csp_init(x264vfw_cuda_csp_function_t *pf, int out_csp, int colmatrix, int fullrange)
{
    // csp.c:440-514
    for (int i = 0; i < X264VFW_CUDA_CSP_MAX; i++) pf->convert[i] = table_fail;
    switch (out_csp) {
    case X264VFW_CUDA_OUT_I420: {
        x264vfw_cuda_csp_t fn;
        if (colmatrix == 1) fn = fullrange ? table_entry<X264VFW_CUDA_OUT_I420, 1, 1> : table_entry<X264VFW_CUDA_OUT_I420, 1, 0>;
        else                fn = fullrange ? table_entry<X264VFW_CUDA_OUT_I420, 0, 1> : table_entry<X264VFW_CUDA_OUT_I420, 0, 0>;
        pf->convert[X264VFW_CUDA_CSP_I420] = fn; pf->convert[X264VFW_CUDA_CSP_YV12] = fn;
        pf->convert[X264VFW_CUDA_CSP_YV16] = fn; pf->convert[X264VFW_CUDA_CSP_YV24] = fn;
        pf->convert[X264VFW_CUDA_CSP_YUYV] = fn; pf->convert[X264VFW_CUDA_CSP_UYVY] = fn;
        pf->convert[X264VFW_CUDA_CSP_BGR] = fn;  pf->convert[X264VFW_CUDA_CSP_BGRA] = fn;
        break;
    }
    case X264VFW_CUDA_OUT_NV12:
        pf->convert[X264VFW_CUDA_CSP_NV12] = table_entry<X264VFW_CUDA_OUT_NV12, 0, 0>; break;
    case X264VFW_CUDA_OUT_I422:
        pf->convert[X264VFW_CUDA_CSP_YV16] = pf->convert[X264VFW_CUDA_CSP_YUYV] =
        pf->convert[X264VFW_CUDA_CSP_UYVY] = table_entry<X264VFW_CUDA_OUT_I422, 0, 0>; break;
    case X264VFW_CUDA_OUT_I444:
        pf->convert[X264VFW_CUDA_CSP_YV24] = table_entry<X264VFW_CUDA_OUT_I444, 0, 0>; break;
    case X264VFW_CUDA_OUT_BGR:
        pf->convert[X264VFW_CUDA_CSP_BGR] = table_entry<X264VFW_CUDA_OUT_BGR, 0, 0>; break;
    case X264VFW_CUDA_OUT_BGRA:
        pf->convert[X264VFW_CUDA_CSP_BGRA] = table_entry<X264VFW_CUDA_OUT_BGRA, 0, 0>; break;
    }
}

void x264vfw_cuda_lowres_geometry(x264vfw_cuda_lowres_geom *g, int w, int h)
{
    g->mb_w = (w + 15) >> 4; g->mb_h = (h + 15) >> 4;
    g->luma_w = 16 * g->mb_w; g->luma_h = 16 * g->mb_h;
    g->luma_stride = (g->luma_w + 1 + 63) & ~63;
    g->lw = g->luma_w / 2; g->lh = g->luma_h / 2;
    g->lstride = (g->lw + 64 + 63) & ~63;
    g->lplane_bytes = g->lstride * (g->lh + 64);
    g->lorigin = 32 * g->lstride + 32;
}

int x264vfw_cuda_lowres_init(x264vfw_cuda_ctx *ctx, uint8_t *dst, const uint8_t *y, int y_stride, int w, int h,
                             size_t sfb, size_t dfb, int n_frames)
{
    if (!ctx || !dst || !y) { set_error("null argument"); return -1; }
    Ctx *c = (Ctx *)ctx;
    XV_CUDA_OK(cudaSetDevice(c->device));
    x264vfw_cuda_lowres_geom g;
    x264vfw_cuda_lowres_geometry(&g, w, h);
    if (!al(dst, 8) || !als((long long)dfb, 8)) { set_error("lowres dst must be 8-byte aligned"); return -1; }
    LowresJob j;
    j.y = y; j.y_stride = y_stride; j.w = w; j.h = h; j.dst = dst;
    j.luma_w = g.luma_w; j.luma_h = g.luma_h; j.lw = g.lw; j.lh = g.lh;
    j.lstride = g.lstride; j.lplane_bytes = g.lplane_bytes; j.lorigin = g.lorigin;
    j.src_frame_bytes = sfb; j.dst_frame_bytes = dfb;
    return launch_lowres_init(c->stream, j, n_frames);
}

int x264vfw_cuda_luma_pad(x264vfw_cuda_ctx *ctx, uint8_t *dst, const uint8_t *y, int y_stride, int w, int h,
                          size_t sfb, size_t dfb, int n_frames)
{
    if (!ctx || !dst || !y) { set_error("null argument"); return -1; }
    Ctx *c = (Ctx *)ctx;
    XV_CUDA_OK(cudaSetDevice(c->device));
    x264vfw_cuda_lowres_geom g;
    x264vfw_cuda_lowres_geometry(&g, w, h);
    LumaPadJob j;
    j.y = y; j.y_stride = y_stride; j.w = w; j.h = h; j.dst = dst; j.dst_stride = g.luma_stride;
    j.luma_w = g.luma_w; j.luma_h = g.luma_h; j.src_frame_bytes = sfb; j.dst_frame_bytes = dfb;
    return launch_luma_pad(c->stream, j, n_frames);
}

int x264vfw_cuda_chroma_nv12_pad(x264vfw_cuda_ctx *ctx, uint8_t *dst, int dst_stride, const uint8_t *u, const uint8_t *v,
                                 int c_stride, int w, int h, size_t sfb, size_t dfb, int n_frames)
{
    if (!ctx || !dst || !u || !v) { set_error("null argument"); return -1; }
    if (w <= 0 || h <= 0 || (w & 1) || (h & 1)) { set_error("width/height must be positive and even"); return -1; }
    Ctx *c = (Ctx *)ctx;
    XV_CUDA_OK(cudaSetDevice(c->device));
    x264vfw_cuda_lowres_geom g;
    x264vfw_cuda_lowres_geometry(&g, w, h);
    if (dst_stride < g.luma_w) { set_error("chroma plane stride %d < %d", dst_stride, g.luma_w); return -1; }
    ChromaPadJob j;
    j.u = u; j.v = v; j.c_stride = c_stride; j.w = w; j.h = h; j.dst = dst; j.dst_stride = dst_stride;
    j.luma_w = g.luma_w; j.luma_h = g.luma_h; j.src_frame_bytes = sfb; j.dst_frame_bytes = dfb;
    return launch_chroma_nv12_pad(c->stream, j, n_frames);
}

int x264vfw_cuda_frontend_batch(x264vfw_cuda_ctx *ctx, int colmatrix, int fullrange, const x264vfw_cuda_image_t *dst, const x264vfw_cuda_image_t *src,
                                uint8_t *lowres, float *qp_offset_aq, uint16_t *inv_qscale, unsigned long long *stats, float aq_strength,
                                int w, int h, size_t sfb, size_t dfb, int n_frames)
{
    if (!ctx || !dst || !src || !lowres || !qp_offset_aq || !inv_qscale || !stats) { set_error("null argument"); return -1; }
    Ctx *c = (Ctx *)ctx;
    XV_CUDA_OK(cudaSetDevice(c->device));
    if ((src->i_csp & X264VFW_CUDA_CSP_MASK) != X264VFW_CUDA_CSP_BGRA) { set_error("fused front end: BGRA sources only"); return -1; }
    if (!frontend_eligible(src->plane[0], src->i_stride[0], sfb, w, h, n_frames, dst->plane[0], dst->i_stride[0], dst->plane[1], dst->plane[2],
                           dst->i_stride[1], dfb) || dst->i_stride[1] != dst->i_stride[2]) {
        set_error("fused front end: needs width %% 16 == 0, 16-byte aligned packed rows, 4-byte aligned luma rows");
        return -1;
    }
    x264vfw_cuda_lowres_geom g;
    x264vfw_cuda_lowres_geometry(&g, w, h);
    FrontendJob j;
    memset(&j, 0, sizeof(j));
    j.dst_y = dst->plane[0]; j.dst_u = dst->plane[1]; j.dst_v = dst->plane[2]; j.y_stride = dst->i_stride[0]; j.c_stride = dst->i_stride[1];
    j.dst_frame_bytes = dfb;
    j.lowres = lowres; j.lowres_frame_bytes = (size_t)4 * g.lplane_bytes; j.lw = g.lw; j.lh = g.lh; j.lstride = g.lstride;
    j.lplane_bytes = g.lplane_bytes; j.lorigin = g.lorigin;
    // f_qp_offset == f_qp_offset_aq at this point of [x264] x264_adaptive_quant_frame: one array serves both
    j.qp_offset = qp_offset_aq; j.qp_offset_aq = qp_offset_aq; j.inv_qscale = inv_qscale; j.stats = stats;
    j.mb_frame_stride = (size_t)g.mb_w * g.mb_h;
    j.aq_on = aq_strength != 0.f; j.aq_mode = 1; j.strength = aq_strength * 1.0397f;
    if (frontend_luts(c->device, &j.log2_lut, &j.exp2_lut) < 0) return -1;
    j.w = w; j.h = h; j.mb_w = g.mb_w; j.mb_h = g.mb_h; j.luma_h = g.luma_h; j.flip = (src->i_csp & X264VFW_CUDA_CSP_VFLIP) != 0;
    j.k = make_rgb_kernel_coef(rgb_coef(colmatrix, fullrange));
    XV_CUDA_OK(cudaMemsetAsync(stats, 0, (size_t)n_frames * 6 * sizeof(unsigned long long), c->stream));
    return launch_frontend(c->stream, j, src->plane[0], src->i_stride[0], sfb, n_frames);
}

void x264vfw_cuda_hpel_geometry(x264vfw_cuda_hpel_geom *g, int w, int h)
{
    g->stride = (w + 64 + 63) & ~63;
    g->plane_bytes = g->stride * (h + 64);
    g->origin = 32 * g->stride + 32;
}

int x264vfw_cuda_hpel_filter(x264vfw_cuda_ctx *ctx, uint8_t *dst, const uint8_t *src, int src_stride, int w, int h,
                             size_t sfb, size_t dfb, int n_frames)
{
    if (!ctx || !dst || !src) { set_error("null argument"); return -1; }
    if (w <= 0 || h <= 0 || (w & 7)) { set_error("hpel: width must be a positive multiple of 8 (upstream: 16*mb_w)"); return -1; }
    Ctx *c = (Ctx *)ctx;
    XV_CUDA_OK(cudaSetDevice(c->device));
    x264vfw_cuda_hpel_geom g;
    x264vfw_cuda_hpel_geometry(&g, w, h);
    if (!al(dst, 8) || !als((long long)dfb, 8)) { set_error("hpel dst must be 8-byte aligned"); return -1; }
    if (n_frames > 1 && dfb < 4 * (size_t)g.plane_bytes) { set_error("hpel: dst_frame_bytes %zu < %zu", dfb, 4 * (size_t)g.plane_bytes); return -1; }
    HpelJob j;
    j.src = src; j.src_stride = src_stride; j.w = w; j.h = h;
    j.dst = dst; j.stride = g.stride; j.plane_bytes = (size_t)g.plane_bytes;
    j.rows_per_strip = 0;
    j.src_frame_bytes = sfb; j.dst_frame_bytes = dfb;
    return launch_hpel(c->stream, j, n_frames);
}

} // extern "C"
