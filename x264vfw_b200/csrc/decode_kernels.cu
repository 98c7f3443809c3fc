// the decoder-side output conversion of x264vfw_decompress (codec.c:2258-2292) -- the CUDA runtime side and the
// C ABI.  Kernels, tables and dispatch: decode_kernel.cuh.
#include "decode_kernel.cuh"


using namespace xv;

extern "C" {

// Host-only views of the filter tables (no device needed): what the kernels are handed, for the CPU test suite.
int x264vfw_cuda_dec_filter_taps(int src_n, int dst_n, int one, int align, int32_t *pos, int16_t *coef)
{
    std::vector<DecTap8> t;
    if (!pos || !coef || src_n <= 0 || dst_n <= 0 || !resample_taps(src_n, dst_n, one, align, t)) return -1;
    for (int i = 0; i < dst_n; i++) {
        pos[i] = t[i].pos;
        for (int k = 0; k < 8; k++) coef[8 * i + k] = (int16_t)((k & 1) ? t[i].c[k >> 1] >> 16 : (t[i].c[k >> 1] & 0xffff));
    }
    return 0;
}

int x264vfw_cuda_dec_packed_rows(int chroma_rows, int b_uyvy, int32_t *pos, int16_t *coef, int32_t *c_writer)
{
    std::vector<DecRow> rows;
    if (!pos || !coef || !c_writer || chroma_rows <= 0 || !vertical_chroma_filter(chroma_rows, rows, b_uyvy != 0)) return -1;
    for (int i = 0; i < 2 * chroma_rows; i++) {
        pos[i] = rows[i].pos; c_writer[i] = rows[i].c_writer;
        coef[4 * i] = (int16_t)(rows[i].c01 & 0xffff); coef[4 * i + 1] = (int16_t)(rows[i].c01 >> 16);
        coef[4 * i + 2] = (int16_t)(rows[i].c23 & 0xffff); coef[4 * i + 3] = (int16_t)(rows[i].c23 >> 16);
    }
    return 0;
}

int64_t x264vfw_cuda_dec_picture_size(int i_out_csp, int w, int h) { return dec_picture_size(i_out_csp, w, h); }

int x264vfw_cuda_dec_open(x264vfw_cuda_dec **pdec, x264vfw_cuda_ctx *ctx, int i_out_csp, int w, int h,
                          int i_src_chroma, int i_avcol_spc, int b_fullrange)
{
    if (!pdec || !ctx) { set_error("null argument"); return -1; }
    *pdec = nullptr;
    Dec *d = new Dec;
    DecTables t;
    if (dec_configure(*d, t, i_out_csp, w, h, i_src_chroma, i_avcol_spc, b_fullrange) < 0) { delete d; return -1; }
    Ctx *c = (Ctx *)ctx;
    d->ctx = c;
    auto upload = [](void **dev, const void *host, size_t bytes) {
        return !bytes || (cudaMalloc(dev, bytes) == cudaSuccess && cudaMemcpy(*dev, host, bytes, cudaMemcpyHostToDevice) == cudaSuccess);
    };
    if (cudaSetDevice(c->device) != cudaSuccess ||
        !upload((void **)&d->d_rows, t.rows.data(), t.rows.size() * sizeof(DecRow)) ||
        !upload((void **)&d->d_cols, t.cols.data(), t.cols.size() * sizeof(DecTap8)) ||
        !upload((void **)&d->d_vtaps, t.vtaps.data(), t.vtaps.size() * sizeof(DecTap8))) {
        set_error("filter table upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        if (d->d_rows) cudaFree(d->d_rows);
        if (d->d_cols) cudaFree(d->d_cols);
        if (d->d_vtaps) cudaFree(d->d_vtaps);
        delete d;
        return -1;
    }
    *pdec = (x264vfw_cuda_dec *)d;
    return 0;
}

void x264vfw_cuda_dec_close(x264vfw_cuda_dec *dec)
{
    Dec *d = (Dec *)dec;
    if (!d) return;
    cudaSetDevice(d->ctx->device);
    cudaStreamSynchronize(d->ctx->stream);
    if (d->d_rows) cudaFree(d->d_rows);
    if (d->d_cols) cudaFree(d->d_cols);
    if (d->d_vtaps) cudaFree(d->d_vtaps);
    if (d->d_src) cudaFree(d->d_src);
    if (d->d_dst) cudaFree(d->d_dst);
    delete d;
}

int x264vfw_cuda_dec_convert_batch(x264vfw_cuda_dec *dec, uint8_t *dst_dev, size_t dst_frame_bytes,
                                   const uint8_t *const src_dev[3], const int src_stride[3], size_t src_frame_bytes,
                                   int n_frames)
{
    Dec *d = (Dec *)dec;
    if (!d || !dst_dev || !src_dev || !src_stride || !src_dev[0] || !src_dev[1] || !src_dev[2]) { set_error("null argument"); return -1; }
    XV_CUDA_OK(cudaSetDevice(d->ctx->device));
    return dec_launch(d, d->ctx->stream, dst_dev, dst_frame_bytes, src_dev, src_stride, src_frame_bytes, n_frames);
}

int x264vfw_cuda_dec_convert(x264vfw_cuda_dec *dec, uint8_t *dst_host, const uint8_t *const src_host[3], const int src_stride[3])
{
    Dec *d = (Dec *)dec;
    if (!d || !dst_host || !src_host || !src_stride || !src_host[0] || !src_host[1] || !src_host[2]) { set_error("null argument"); return -1; }
    for (int i = 0; i < 3; i++) if (src_stride[i] < (i && !d->v444 ? d->w / 2 : d->w)) { set_error("source stride below the row width"); return -1; }
    XV_CUDA_OK(cudaSetDevice(d->ctx->device));
    cudaStream_t st = d->ctx->stream;
    const int w = d->w, h = d->h, cw = d->v444 ? w : w / 2, ch = d->v422 || d->v444 ? h : h / 2;
    // device staging: tight planes with 16-byte aligned rows; the picture is gathered by 2-D copies so that the
    // decoder's linesize padding never crosses the bus
    const int ys = (w + 15) & ~15, cs = (cw + 15) & ~15;
    const size_t sneed = (size_t)ys * h + 2 * (size_t)cs * ch, dneed = (size_t)x264vfw_cuda_dec_picture_size(d->csp, w, h);
    if (d->src_bytes < sneed) {
        if (d->d_src) cudaFree(d->d_src);
        d->d_src = nullptr; d->src_bytes = 0;
        XV_CUDA_OK(cudaMalloc((void **)&d->d_src, sneed));
        d->src_bytes = sneed;
    }
    if (d->dst_bytes < dneed) {
        if (d->d_dst) cudaFree(d->d_dst);
        d->d_dst = nullptr; d->dst_bytes = 0;
        XV_CUDA_OK(cudaMalloc((void **)&d->d_dst, dneed));
        XV_CUDA_OK(cudaMemsetAsync(d->d_dst, 0, dneed, st));      // BGR24 row padding is never written: keep it defined
        d->dst_bytes = dneed;
    }
    uint8_t *py = d->d_src, *pu = py + (size_t)ys * h, *pv = pu + (size_t)cs * ch;
    XV_CUDA_OK(cudaMemcpy2DAsync(py, ys, src_host[0], src_stride[0], w, h, cudaMemcpyHostToDevice, st));
    XV_CUDA_OK(cudaMemcpy2DAsync(pu, cs, src_host[1], src_stride[1], cw, ch, cudaMemcpyHostToDevice, st));
    XV_CUDA_OK(cudaMemcpy2DAsync(pv, cs, src_host[2], src_stride[2], cw, ch, cudaMemcpyHostToDevice, st));
    const uint8_t *sp[3] = {py, pu, pv};
    const int ss[3] = {ys, cs, cs};
    if (dec_launch(d, st, d->d_dst, 0, sp, ss, 0, 1) < 0) return -1;
    XV_CUDA_OK(cudaMemcpyAsync(dst_host, d->d_dst, dneed, cudaMemcpyDeviceToHost, st));
    XV_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

} // extern "C"
