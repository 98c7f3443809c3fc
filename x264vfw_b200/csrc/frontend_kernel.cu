// Fused front end of a frame: ONE pass over the packed RGB32 rows produces everything the lookahead needs.
//
//   stage 1   RGB_TO_I420 (csp.c:299-388)                               -> Y, U, V planes (codec->conv_pic)
//   [x264]    x264_adaptive_quant_frame / ac_energy_mb                 -> per-MB qp offsets, inv_qscale, frame sums
//   [x264]    x264_frame_expand_border_mod16 + x264_frame_init_lowres  -> the four half-resolution phase planes
//             + x264_frame_expand_border_lowres                           incl. their 32-pixel replicated border
//
// Before, these were three launches per frame that read the luma plane twice more out of L2/HBM.  Here a CTA owns
// a 128 x 32 luma tile (8 x 2 macroblocks, 64 x 16 lowres pixels):
//   * the tile's packed source rows (+ one halo row and column for the lowres filter) arrive as ONE TMA 3-D box
//     (cp.async.bulk.tensor, 132 x 33 pixels of 4 bytes, frame = 3rd coordinate) completing on an mbarrier; the
//     bottom-up DIB flip is only a different box origin and a reversed row index in shared memory;
//   * every thread converts 4 pixels x 2 rows out of shared memory with the exact dp2a/dp4a re-encoding of the
//     reference arithmetic (rgb_math.cuh), stores Y/U/V, keeps Y in a shared luma tile and accumulates the
//     macroblock sums (replicated rows below the picture count with their multiplicity);
//   * the lowres phases come from the shared luma tile with nested byte-wise rounding averages
//     (FILTER(a,b,c,d) = (((a+b+1)>>1)+((c+d+1)>>1)+1)>>1), clamped at the picture edge (== the mod-16 replication
//     and the duplicated last row/column of upstream); tiles on a frame edge also write the replicated border.
// TMA zero-fills out-of-picture elements; they are never read (all indices are clamped to the picture).
//
// Eligibility (checked by the launcher): BGRA input, I420 output, width % 16 == 0, 16-byte aligned source rows.
// Everything else keeps the three separate kernels.  Algorithmic bytes at 1080p: 8,294,400 in + 3,110,400 planes
// + 4 x 522,240 lowres = 13,493,760 B per frame; border bytes are written but not credited.
#include "rgb_math.cuh"
#include "frontend.h"

namespace xv {

#define FE_LW 128
#define FE_LH 32
#define FE_BOXW (FE_LW + 4)
#define FE_BOXH (FE_LH + 1)
#define FE_YS 144
#define FE_PAD 32

struct __align__(128) FeSmem {
    uint32_t rgb[FE_BOXH][FE_BOXW];      // TMA destination: packed pixels of the tile + halo
    uint8_t y[FE_BOXH][FE_YS];           // luma of the tile + halo row / column
    uint8_t lr[4][FE_LH / 2][FE_LW / 2]; // the tile's lowres pixels (border replication reads them back)
    unsigned int mb[16][6];              // per-MB sums: Y, U, V sum then Y, U, V sum of squares
    unsigned long long mbar;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// lowres phases of 4 output pixels from 8 (vertically averaged) luma bytes + the 9th
__device__ __forceinline__ void hphase4(uint32_t lo, uint32_t hi, uint32_t tail, uint32_t &p0, uint32_t &ph)
{
    const uint32_t e = __byte_perm(lo, hi, 0x6420), o = __byte_perm(lo, hi, 0x7531);
    const uint32_t n = (e >> 8) | (tail << 24);
    p0 = avg4(e, o);
    ph = avg4(o, n);
}

// conversion of the thread's two row pairs (4 pixels wide each, same macroblock).  EDGE: the tile crosses the right or
// bottom picture edge (bounds checks, replicated rows count with their multiplicity); interior tiles skip all of that.
template <bool FLIP, bool EDGE>
__device__ __forceinline__ void fe_convert(const FrontendJob &job, FeSmem &sm, int tid, int x0, int y0, uint8_t *Y, uint8_t *U, uint8_t *V)
{
    unsigned s[6] = {0, 0, 0, 0, 0, 0};
    const int c = tid & 31, p0 = 2 * (tid >> 5);
    const int x = x0 + 4 * c;
    // byte offsets inside the frame fit 32 bits
    uint8_t *py = Y + (unsigned)((y0 + 2 * p0) * job.y_stride + x);
    uint8_t *pu = U + (unsigned)(((y0 >> 1) + p0) * job.c_stride + (x >> 1));
    uint8_t *pv = V + (unsigned)(((y0 >> 1) + p0) * job.c_stride + (x >> 1));
#pragma unroll
    for (int it = 0; it < 2; it++) {
        const int p = p0 + it, y = y0 + 2 * p;
        if (!EDGE || (x < job.w && y < job.h)) {
            const int ra = FLIP ? FE_LH - 2 * p : 2 * p, rb = FLIP ? ra - 1 : ra + 1;
            const uint4 t = *(const uint4 *)&sm.rgb[ra][4 * c], b = *(const uint4 *)&sm.rgb[rb][4 * c];
            const uint32_t yt = __byte_perm(__byte_perm(luma16(job.k, t.x), luma16(job.k, t.y), 0x0073),
                                            __byte_perm(luma16(job.k, t.z), luma16(job.k, t.w), 0x0073), 0x5410);
            const uint32_t yb = __byte_perm(__byte_perm(luma16(job.k, b.x), luma16(job.k, b.y), 0x0073),
                                            __byte_perm(luma16(job.k, b.z), luma16(job.k, b.w), 0x0073), 0x5410);
            uint32_t u0, v0, u1, v1;
            chroma_quad(job.k, t.x, b.x, t.y, b.y, u0, v0);
            chroma_quad(job.k, t.z, b.z, t.w, b.w, u1, v1);
            *(uint32_t *)(py + it * 2 * job.y_stride) = yt;
            *(uint32_t *)(py + (it * 2 + 1) * job.y_stride) = yb;
            *(uint16_t *)(pu + it * job.c_stride) = (uint16_t)(u0 | (u1 << 8));
            *(uint16_t *)(pv + it * job.c_stride) = (uint16_t)(v0 | (v1 << 8));
            *(uint32_t *)&sm.y[2 * p][4 * c] = yt;
            *(uint32_t *)&sm.y[2 * p + 1][4 * c] = yb;
            // macroblock sums; the rows below the picture replicate its last row ([x264] expand_border_mod16):
            // that row counts 1 + (luma_h - h) times, the last chroma row 1 + (8 mb_h - h/2) times
            unsigned my = 1u, mc = 1u;
            if (EDGE && y + 2 == job.h) { my += (unsigned)(job.luma_h - job.h); mc += (unsigned)(8 * job.mb_h - job.h / 2); }
            s[0] += __dp4a(yt, 0x01010101u, 0u) + my * __dp4a(yb, 0x01010101u, 0u);
            s[3] += __dp4a(yt, yt, 0u) + my * __dp4a(yb, yb, 0u);
            s[1] += mc * (u0 + u1); s[4] += mc * (u0 * u0 + u1 * u1);
            s[2] += mc * (v0 + v1); s[5] += mc * (v0 * v0 + v1 * v1);
        }
    }
    // the 4 lanes of a macroblock column meet by shuffle, one of them adds to the MB's accumulators
#pragma unroll
    for (int k = 0; k < 6; k++) {
        s[k] += __shfl_xor_sync(0xffffffffu, s[k], 1);
        s[k] += __shfl_xor_sync(0xffffffffu, s[k], 2);
    }
    if ((c & 3) == 0) {
        unsigned int *acc = sm.mb[(tid >> 7) * 8 + (c >> 2)];
#pragma unroll
        for (int k = 0; k < 6; k++) if (s[k]) atomicAdd(acc + k, s[k]);
    }
}

template <bool FLIP>
__global__ void __launch_bounds__(256, 8)
frontend_kernel(const __grid_constant__ FrontendJob job)
{
    __shared__ FeSmem sm;
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * FE_LW, y0 = blockIdx.y * FE_LH;
    const size_t f = blockIdx.z;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&sm.mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 96) (&sm.mb[0][0])[tid] = 0;
    __syncthreads();
    if (tid == 0) {
        // tile rows y0 .. y0+32 are source rows h-1-y0-32 .. h-1-y0 of a bottom-up DIB (csp.c:310-314)
        const int row0 = FLIP ? job.h - 1 - y0 - FE_LH : y0;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&sm.mbar)), "r"((uint32_t)sizeof(sm.rgb)) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     :: "r"(smem_u32(&sm.rgb[0][0])), "l"((uint64_t)&job.tmap), "r"(x0), "r"(row0), "r"((int)f), "r"(smem_u32(&sm.mbar)) : "memory");
    }
    {   // wait for the box (phase 0 of the barrier)
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(smem_u32(&sm.mbar)) : "memory");
    }
    auto srow = [&](int j) { return FLIP ? FE_LH - j : j; };      // shared-memory row of tile row j
    const int rmax = min(FE_LH, job.h - 1 - y0);                   // last tile row / column inside the picture
    const int cmax = min(FE_LW, job.w - 1 - x0);

    // ---- conversion: thread = 4 pixels x a row pair, two such items per thread ----
    uint8_t *Y = job.dst_y + f * job.dst_frame_bytes, *U = job.dst_u + f * job.dst_frame_bytes, *V = job.dst_v + f * job.dst_frame_bytes;
    if (x0 + FE_LW <= job.w && y0 + FE_LH <= job.h) fe_convert<FLIP, false>(job, sm, tid, x0, y0, Y, U, V);
    else fe_convert<FLIP, true>(job, sm, tid, x0, y0, Y, U, V);
    // halo: luma of tile column 128 (rows 0..32) and of tile row 32 (columns 0..127), where inside the picture
    for (int i = tid; i < FE_BOXH + FE_LW; i += 256) {
        const int j = i < FE_BOXH ? i : FE_LH, cx = i < FE_BOXH ? FE_LW : i - FE_BOXH;
        if (j <= rmax && cx <= cmax) sm.y[j][cx] = (uint8_t)(luma16(job.k, sm.rgb[srow(j)][cx]) >> 24);
    }
    __syncthreads();

    // ---- adaptive quantisation of the tile's 16 macroblocks + frame sums ([x264] x264_adaptive_quant_frame) ----
    if (tid < 32) {
        unsigned long long st[6] = {0, 0, 0, 0, 0, 0};
        const int mbx = (x0 >> 4) + (tid & 7), mby = (y0 >> 4) + ((tid >> 3) & 1);
        if (tid < 16 && mbx < job.mb_w && mby < job.mb_h) {
            const unsigned int *a = sm.mb[tid];
            uint32_t energy = 0;
#pragma unroll
            for (int pl = 0; pl < 3; pl++) {
                const uint32_t sum = a[pl], ssd = a[3 + pl];
                st[pl] = sum; st[3 + pl] = ssd;
                energy += ssd - (uint32_t)(((unsigned long long)sum * sum) >> (pl ? 6 : 8));
            }
            const size_t idx = f * job.mb_frame_stride + (size_t)mby * job.mb_w + mbx;
            if (job.aq_on && job.aq_mode >= 2) {
                job.qp_offset[idx] = sqrtf(sqrtf(sqrtf((float)energy * 1.f + 1)));     // aq_auto_kernel finishes the frame
            } else if (job.aq_on) {
                const uint32_t e1 = max(energy, 1u);
                const int lz = __clz(e1);
                const float qp_adj = job.strength * ((job.log2_lut[(e1 << lz >> 24) & 0x7f] + (float)(31 - lz)) - (14.427f + 2 * 0));
                job.qp_offset[idx] = qp_adj;
                job.qp_offset_aq[idx] = qp_adj;
                int i = (int)(qp_adj * (-64.f / 6.f) + 512.5f);
                job.inv_qscale[idx] = (uint16_t)(i < 0 ? 0 : i > 1023 ? 0xffff : (job.exp2_lut[i & 63] + 256) << (i >> 6) >> 8);
            } else {
                job.qp_offset[idx] = 0.f; job.qp_offset_aq[idx] = 0.f; job.inv_qscale[idx] = 256;
            }
        }
#pragma unroll
        for (int i = 0; i < 6; i++) {
            unsigned long long t = st[i];
#pragma unroll
            for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (tid == 0 && t) atomicAdd(job.stats + f * 6 + i, t);
        }
    }

    // ---- lowres: thread = 4 pixels of one lowres row, four phase planes ----
    const int lx0 = x0 >> 1, ly0 = y0 >> 1;
    uint8_t *L = job.lowres + f * job.lowres_frame_bytes + job.lorigin;
    {
        const int ch = tid & 15, r = tid >> 4;
        const int lx = lx0 + 4 * ch, ly = ly0 + r;
        if (lx < job.lw && ly < job.lh) {
            const int ya = min(2 * r, rmax), yb = min(2 * r + 1, rmax), yc = min(2 * r + 2, rmax);
            const int xc = 8 * ch, tx = min(xc + 8, cmax);
            const uint2 A = *(const uint2 *)&sm.y[ya][xc], B = *(const uint2 *)&sm.y[yb][xc], C = *(const uint2 *)&sm.y[yc][xc];
            const uint32_t ta = sm.y[ya][tx], tb = sm.y[yb][tx], tc = sm.y[yc][tx];
            uint32_t o[4];
            hphase4(avg4(A.x, B.x), avg4(A.y, B.y), (ta + tb + 1) >> 1, o[0], o[1]);
            hphase4(avg4(B.x, C.x), avg4(B.y, C.y), (tb + tc + 1) >> 1, o[2], o[3]);
            uint8_t *d = L + (size_t)ly * job.lstride + lx;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                *(uint32_t *)(d + (size_t)k * job.lplane_bytes) = o[k];
                *(uint32_t *)&sm.lr[k][r][4 * ch] = o[k];
            }
        }
    }

    // ---- 32-pixel replicated border ([x264] x264_frame_expand_border_lowres), tiles on a frame edge only ----
    const bool left = x0 == 0, right = lx0 + FE_LW / 2 >= job.lw, top = y0 == 0, bottom = ly0 + FE_LH / 2 >= job.lh;
    if (!(left || right || top || bottom)) return;
    __syncthreads();
    const int nrow = min(FE_LH / 2, job.lh - ly0), ncol = min(FE_LW / 2, job.lw - lx0);
    // padded coordinates: pixel (lx, ly) sits at row ly + 32, column lx + 32 of a plane
    uint8_t *P = job.lowres + f * job.lowres_frame_bytes;
    auto rep8 = [](uint8_t v) { const uint32_t w = v * 0x01010101u; return make_uint2(w, w); };
    if (left || right) {
        // rows of the tile: 32 bytes of the first / last column, 8 bytes per item
        for (int i = tid; i < 2 * 4 * nrow * 4; i += 256) {
            const int q = i & 3, r = (i >> 2) % nrow, k = ((i >> 2) / nrow) & 3, side = (i >> 2) / nrow >> 2;
            if (side == 0 ? !left : !right) continue;
            const uint8_t v = sm.lr[k][r][side ? ncol - 1 : 0];
            const int pc = side ? FE_PAD + job.lw + 8 * q : 8 * q;
            *(uint2 *)(P + (size_t)k * job.lplane_bytes + (size_t)(ly0 + r + FE_PAD) * job.lstride + pc) = rep8(v);
        }
    }
    if (top || bottom) {
        // 32 rows above / below: the tile's columns plus, on a corner tile, the 32 border columns beside them
        const int c8 = ncol >> 3;                                  // 8-byte chunks of the tile's columns
        const int per_row = c8 + (left ? 4 : 0) + (right ? 4 : 0);
        for (int i = tid; i < 2 * 4 * FE_PAD * per_row; i += 256) {
            const int cc = i % per_row, pr = (i / per_row) % FE_PAD, k = (i / per_row / FE_PAD) & 3, side = i / per_row / FE_PAD >> 2;
            if (side == 0 ? !top : !bottom) continue;
            const int r = side ? nrow - 1 : 0;
            int pc; uint2 v;
            if (cc < c8) { pc = FE_PAD + lx0 + 8 * cc; v = *(const uint2 *)&sm.lr[k][r][8 * cc]; }
            else if (left && cc < c8 + 4) { pc = 8 * (cc - c8); v = rep8(sm.lr[k][r][0]); }
            else { pc = FE_PAD + job.lw + 8 * (cc - c8 - (left ? 4 : 0)); v = rep8(sm.lr[k][r][ncol - 1]); }
            const int prow = side ? FE_PAD + job.lh + pr : pr;
            *(uint2 *)(P + (size_t)k * job.lplane_bytes + (size_t)prow * job.lstride + pc) = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled()
{
    // the driver entry point is resolved through the runtime: the library does not link libcuda
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

bool frontend_eligible(const void *src, long long src_stride, size_t src_frame_bytes, int w, int h, int n_frames,
                       const void *dst_y, int y_stride, const void *dst_u, const void *dst_v, int c_stride, size_t dst_frame_bytes)
{
    if ((w & 15) || (h & 1) || w < 16 || h < 2) return false;
    if (((uintptr_t)src & 15) || (src_stride & 15) || src_stride < 4LL * w) return false;
    if (n_frames > 1 && ((src_frame_bytes & 15) || src_frame_bytes < (size_t)src_stride * h)) return false;
    if (((uintptr_t)dst_y & 3) || (y_stride & 3) || ((uintptr_t)dst_u & 1) || ((uintptr_t)dst_v & 1) || (c_stride & 1) || (dst_frame_bytes & 3)) return false;
    return encode_tiled() != nullptr;
}

// src: first byte of the packed BGRA buffer (top row in memory), src_stride > 0; flip = bottom-up DIB.
int launch_frontend(cudaStream_t st, FrontendJob &job, const uint8_t *src, long long src_stride, size_t src_frame_bytes, int n_frames)
{
    EncodeTiledFn enc = encode_tiled();
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return -1; }
    const cuuint64_t dims[3] = {(cuuint64_t)job.w, (cuuint64_t)job.h, (cuuint64_t)n_frames};
    const cuuint64_t fstride = n_frames > 1 ? (cuuint64_t)src_frame_bytes : (cuuint64_t)src_stride * job.h;
    const cuuint64_t strides[2] = {(cuuint64_t)src_stride, fstride};
    const cuuint32_t box[3] = {FE_BOXW, FE_BOXH, 1}, estr[3] = {1, 1, 1};
    const CUresult rc = enc(&job.tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void *)src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)rc); return -1; }
    const dim3 grid((unsigned)((job.w + FE_LW - 1) / FE_LW), (unsigned)((job.luma_h + FE_LH - 1) / FE_LH), (unsigned)n_frames);
    if (job.flip) frontend_kernel<true><<<grid, 256, 0, st>>>(job);
    else frontend_kernel<false><<<grid, 256, 0, st>>>(job);
    XV_LAUNCH_CHECK();
    return 0;
}

} // namespace xv
