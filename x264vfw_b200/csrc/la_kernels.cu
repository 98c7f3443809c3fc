// Data-parallel lookahead kernels: adaptive-quant statistics, intra cost, per-MB cost
// selection + frame accumulators, weight scoring, mb-tree propagation.
//
// These restate, per macroblock, the upstream libx264 functions the reference reaches only
// through x264_encoder_encode (reference codec.c:1693):
//   [x264] encoder/ratecontrol.c  x264_adaptive_quant_frame, ac_energy_mb
//   [x264] encoder/slicetype.c    slicetype_mb_cost (intra part, bidir part, selection),
//                                 slicetype_frame_cost accumulators, weight_cost_luma,
//                                 macroblock_tree_propagate, macroblock_tree_finish
//   [x264] common/mc.c            mbtree_propagate_cost, mbtree_propagate_list, mc_weight
//   [x264] common/predict.c       predict_8x8c_{dc,h,v,p}, predict_8x8_filter, predict_8x8_{ddl..hu}
// Every MB is independent in these kernels (the serial part of the lookahead, the motion
// search with its spatial MV predictors, lives in la_me_kernel.cu), so each is one thread per
// MB with warp-shuffle + atomic reductions for the frame sums.  Integer results are bit-exact;
// float code is compiled with --fmad=false and keeps the reference's operation order.
#include "la_common.cuh"

namespace xv {

__device__ __forceinline__ int warp_sum(int v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned long long warp_sum64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------
// Adaptive quantisation statistics.  One thread per 16x16 MB; a warp reads 512 contiguous
// bytes per luma row.  Clamped addressing == the mod-16 replicated frame.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float dev_log2(const float *lut, uint32_t x)
{
    int lz = __clz(x);
    return lut[(x << lz >> 24) & 0x7f] + (float)(31 - lz);
}
__device__ __forceinline__ int dev_exp2fix8(const uint8_t *lut, float x)
{
    int i = (int)(x * (-64.f / 6.f) + 512.5f);
    if (i < 0) return 0;
    if (i > 1023) return 0xffff;
    return (lut[i & 63] + 256) << (i >> 6) >> 8;
}

// sum and sum of squares of rows [ya, yb) of the bw x bh block at (x0, y0); samples are `step`
// bytes apart (2 = one component of an interleaved NV12 chroma plane)
__device__ __forceinline__ void block_sums_dev(const uint8_t *p, int stride, int pw, int ph, int x0, int y0,
                                               int bw, int ya, int yb, uint32_t &sum, uint32_t &ssd, int step = 1)
{
    sum = 0; ssd = 0;
    const bool fast = (x0 + bw <= pw) && (((uintptr_t)p | (unsigned)stride) & 15) == 0 && (x0 & 15) == 0 && (bw & 7) == 0;
    // interleaved pair: p points at the component's first byte (U: even, V: odd address)
    const bool fast2 = step == 2 && (x0 + bw <= pw) && ((((uintptr_t)p & ~(uintptr_t)1) | (unsigned)stride) & 15) == 0 && (x0 & 7) == 0 && bw == 8;
    const unsigned sh2 = ((uintptr_t)p & 1) * 8;
    for (int y = ya; y < yb; y++) {
        const uint8_t *r = p + (size_t)min(y0 + y, ph - 1) * stride;
        if (step == 1 && fast) {
            for (int x = 0; x < bw; x += 8) {
                uint2 v = *(const uint2 *)(r + x0 + x);
                sum = __dp4a(v.x, 0x01010101u, sum); sum = __dp4a(v.y, 0x01010101u, sum);
                ssd = __dp4a(v.x, v.x, ssd); ssd = __dp4a(v.y, v.y, ssd);
            }
        } else if (fast2) {
            const uint4 v = *(const uint4 *)((const uint8_t *)((uintptr_t)r & ~(uintptr_t)1) + 2 * x0);
            const uint32_t a = (v.x >> sh2) & 0x00ff00ffu, b = (v.y >> sh2) & 0x00ff00ffu, c = (v.z >> sh2) & 0x00ff00ffu, d = (v.w >> sh2) & 0x00ff00ffu;
            sum = __dp4a(a, 0x01010101u, sum); sum = __dp4a(b, 0x01010101u, sum); sum = __dp4a(c, 0x01010101u, sum); sum = __dp4a(d, 0x01010101u, sum);
            ssd = __dp4a(a, a, ssd); ssd = __dp4a(b, b, ssd); ssd = __dp4a(c, c, ssd); ssd = __dp4a(d, d, ssd);
        } else {
            for (int x = 0; x < bw; x++) { uint32_t v = r[min(x0 + x, pw - 1) * step]; sum += v; ssd += v * v; }
        }
    }
}

// A block = 32 consecutive MBs x AQ_PARTS row slices: warp `part` sums rows
// [part*bh/AQ_PARTS, ...) of every plane block of its 32 MBs (512 contiguous bytes per luma
// row), the slices meet in shared memory (integer sums: order-free), warp 0 finishes the MB.
#define AQ_PARTS 8
__global__ void __launch_bounds__(32 * AQ_PARTS)
aq_kernel(LaGeom g, AqJob job)
{
    __shared__ uint32_t sh[3][2][AQ_PARTS][32];
    const int lane = threadIdx.x & 31, part = threadIdx.x >> 5;
    const int idx = blockIdx.x * 32 + lane;
    const int w = g.width, h = g.height, cf = job.chroma_format;
    const int cw = cf == 3 ? w : w / 2, ch = cf == 1 ? h / 2 : h;
    const int cbw = cf == 3 ? 16 : 8, cbh = cf == 1 ? 8 : 16, cshift = cf == 3 ? 8 : cf == 2 ? 7 : 6;
    uint32_t s[3] = {0, 0, 0}, q[3] = {0, 0, 0};
    if (idx < g.mb_count) {
        const int mx = idx % g.mb_w, my = idx / g.mb_w;
        block_sums_dev(job.y, job.y_stride, w, h, 16 * mx, 16 * my, 16, part * 16 / AQ_PARTS, (part + 1) * 16 / AQ_PARTS, s[0], q[0]);
        if (cf) {
            block_sums_dev(job.u, job.c_stride, cw, ch, cbw * mx, cbh * my, cbw, part * cbh / AQ_PARTS, (part + 1) * cbh / AQ_PARTS, s[1], q[1], job.c_step);
            block_sums_dev(job.v, job.c_stride, cw, ch, cbw * mx, cbh * my, cbw, part * cbh / AQ_PARTS, (part + 1) * cbh / AQ_PARTS, s[2], q[2], job.c_step);
        }
    }
#pragma unroll
    for (int pl = 0; pl < 3; pl++) { sh[pl][0][part][lane] = s[pl]; sh[pl][1][part][lane] = q[pl]; }
    __syncthreads();
    if (part) return;
    unsigned long long st[6] = {0, 0, 0, 0, 0, 0};
    if (idx < g.mb_count) {
        uint32_t energy = 0;
#pragma unroll
        for (int pl = 0; pl < 3; pl++) {
            uint32_t sum = 0, ssd = 0;
#pragma unroll
            for (int k = 0; k < AQ_PARTS; k++) { sum += sh[pl][0][k][lane]; ssd += sh[pl][1][k][lane]; }
            st[pl] = sum; st[3 + pl] = ssd;
            if (pl == 0 || cf) energy += ssd - (uint32_t)(((unsigned long long)sum * sum) >> (pl ? cshift : 8));
        }
        if (job.aq_on && job.aq_mode >= 2) {
            // auto-variance modes, first loop: (energy + 1)^(1/8) as three correctly rounded square
            // roots (see the CPU checker); aq_auto_kernel finishes the frame
            job.qp_offset[idx] = sqrtf(sqrtf(sqrtf((float)energy * 1.f + 1)));
        } else if (job.aq_on) {
            float qp_adj = job.strength * (dev_log2(job.log2_lut, max(energy, 1u)) - (14.427f + 2 * 0));
            job.qp_offset[idx] = qp_adj;
            job.qp_offset_aq[idx] = qp_adj;
            job.inv_qscale[idx] = (uint16_t)dev_exp2fix8(job.exp2_lut, qp_adj);
        } else {
            job.qp_offset[idx] = 0.f; job.qp_offset_aq[idx] = 0.f; job.inv_qscale[idx] = 256;
        }
    }
#pragma unroll
    for (int i = 0; i < 6; i++) {
        unsigned long long t = warp_sum64(st[i]);
        if (lane == 0 && t) atomicAdd(job.stats + i, t);
    }
}

// [x264] x264_adaptive_quant_frame, X264_AQ_AUTOVARIANCE / _BIASED: the frame averages are float
// sums taken in MB order, so ONE thread adds them in that order (out of shared memory, 8192 MBs
// at a time); the second loop over the MBs is elementwise and uses the whole block.
#define AQ_AUTO_CHUNK 8192
__global__ void __launch_bounds__(1024)
aq_auto_kernel(LaGeom g, AqJob job)
{
    __shared__ float buf[AQ_AUTO_CHUNK];
    __shared__ float s_avg, s_strength;
    float avg_adj = 0.f, avg_adj_pow2 = 0.f;
    for (int base = 0; base < g.mb_count; base += AQ_AUTO_CHUNK) {
        const int n = min(AQ_AUTO_CHUNK, g.mb_count - base);
        for (int i = threadIdx.x; i < n; i += blockDim.x) buf[i] = job.qp_offset[base + i];
        __syncthreads();
        if (threadIdx.x == 0)
            for (int i = 0; i < n; i++) { const float q = buf[i]; avg_adj += q; avg_adj_pow2 += q * q; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        avg_adj /= g.mb_count;
        avg_adj_pow2 /= g.mb_count;
        s_strength = job.aq_strength * avg_adj;
        s_avg = avg_adj - 0.5f * (avg_adj_pow2 - 14.f) / avg_adj;
    }
    __syncthreads();
    const float strength = s_strength, avg = s_avg, bias_strength = job.aq_strength;
    for (int i = threadIdx.x; i < g.mb_count; i += blockDim.x) {
        float qp_adj = job.qp_offset[i];
        if (job.aq_mode == 3) qp_adj = strength * (qp_adj - avg) + bias_strength * (1.f - 14.f / (qp_adj * qp_adj));
        else qp_adj = strength * (qp_adj - avg);
        job.qp_offset[i] = qp_adj;
        job.qp_offset_aq[i] = qp_adj;
        job.inv_qscale[i] = (uint16_t)dev_exp2fix8(job.exp2_lut, qp_adj);
    }
}

int launch_aq_auto(cudaStream_t st, const LaGeom &g, const AqJob &job);
int launch_aq(cudaStream_t st, const LaGeom &g, const AqJob &job)
{
    aq_kernel<<<(g.mb_count + 31) / 32, 32 * AQ_PARTS, 0, st>>>(g, job);
    XV_LAUNCH_CHECK();
    if (job.aq_on && job.aq_mode >= 2) return launch_aq_auto(st, g, job);
    return 0;
}

// second half of the auto-variance modes on its own (the fused front end produces the first half)
int launch_aq_auto(cudaStream_t st, const LaGeom &g, const AqJob &job)
{
    aq_auto_kernel<<<1, 1024, 0, st>>>(g, job);
    XV_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------
// Intra cost.  One thread per lowres 8x8 MB.
// ------------------------------------------------------------------------------------------
struct Nbr { int tl; int top[16]; int left[8]; };
struct Edge { int c; int t[16]; int l[8]; };

__device__ __forceinline__ uint32_t splat4(int v) { return (uint32_t)v * 0x01010101u; }

__device__ __forceinline__ void filter_edges_dev(Edge &e, const Nbr &n)
{
    e.c = (n.top[0] + 2 * n.tl + n.left[0] + 2) >> 2;
    e.t[0] = (n.tl + 2 * n.top[0] + n.top[1] + 2) >> 2;
#pragma unroll
    for (int x = 1; x < 15; x++) e.t[x] = (n.top[x - 1] + 2 * n.top[x] + n.top[x + 1] + 2) >> 2;
    e.t[15] = (n.top[14] + 3 * n.top[15] + 2) >> 2;
    e.l[0] = (n.tl + 2 * n.left[0] + n.left[1] + 2) >> 2;
#pragma unroll
    for (int y = 1; y < 7; y++) e.l[y] = (n.left[y - 1] + 2 * n.left[y] + n.left[y + 1] + 2) >> 2;
    e.l[7] = (n.left[6] + 3 * n.left[7] + 2) >> 2;
}

// index -1 is the (filtered) corner; the clamps only keep never-selected branches in bounds
#define ET(i) ((i) < 0 ? e.c : e.t[(i) < 0 ? 0 : (i) > 15 ? 15 : (i)])
#define EL(i) ((i) < 0 ? e.c : e.l[(i) < 0 ? 0 : (i) > 7 ? 7 : (i)])
template <int MODE>
__device__ __forceinline__ int pred8x8_px(const Edge &e, int x, int y)
{
    if (MODE == 3) return (x == 7 && y == 7) ? (ET(14) + 3 * ET(15) + 2) >> 2 : (ET(x + y) + 2 * ET(x + y + 1) + ET(x + y + 2) + 2) >> 2;
    if (MODE == 4) {
        if (x > y) return (ET(x - y - 2) + 2 * ET(x - y - 1) + ET(x - y) + 2) >> 2;
        if (x < y) return (EL(y - x - 2) + 2 * EL(y - x - 1) + EL(y - x) + 2) >> 2;
        return (ET(0) + 2 * e.c + EL(0) + 2) >> 2;
    }
    if (MODE == 5) {
        const int z = 2 * x - y, i = x - (y >> 1);
        if (z >= 0 && !(z & 1)) return (ET(i - 1) + ET(i) + 1) >> 1;
        if (z >= 0) return (ET(i - 2) + 2 * ET(i - 1) + ET(i) + 2) >> 2;
        if (z == -1) return (EL(0) + 2 * e.c + ET(0) + 2) >> 2;
        return (EL(y - 2 * x - 1) + 2 * EL(y - 2 * x - 2) + EL(y - 2 * x - 3) + 2) >> 2;
    }
    if (MODE == 6) {
        const int z = 2 * y - x, i = y - (x >> 1);
        if (z >= 0 && !(z & 1)) return (EL(i - 1) + EL(i) + 1) >> 1;
        if (z >= 0) return (EL(i - 2) + 2 * EL(i - 1) + EL(i) + 2) >> 2;
        if (z == -1) return (EL(0) + 2 * e.c + ET(0) + 2) >> 2;
        return (ET(x - 2 * y - 1) + 2 * ET(x - 2 * y - 2) + ET(x - 2 * y - 3) + 2) >> 2;
    }
    if (MODE == 7) {
        const int i = x + (y >> 1);
        return !(y & 1) ? (ET(i) + ET(i + 1) + 1) >> 1 : (ET(i) + 2 * ET(i + 1) + ET(i + 2) + 2) >> 2;
    }
    {   // 8: horizontal-up
        const int z = x + 2 * y, i = y + (x >> 1);
        if (z > 13) return EL(7);
        if (z == 13) return (EL(6) + 3 * EL(7) + 2) >> 2;
        if (!(z & 1)) return (EL(i) + EL(i + 1) + 1) >> 1;
        return (EL(i) + 2 * EL(i + 1) + EL(i + 2) + 2) >> 2;
    }
}
#undef ET
#undef EL

template <int MODE>
__device__ __forceinline__ void pred8x8_build(const Edge &e, uint2 pr[8])
{
#pragma unroll
    for (int y = 0; y < 8; y++) {
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int x = 0; x < 4; x++) {
            lo |= (uint32_t)pred8x8_px<MODE>(e, x, y) << (8 * x);
            hi |= (uint32_t)pred8x8_px<MODE>(e, x + 4, y) << (8 * x);
        }
        pr[y] = make_uint2(lo, hi);
    }
}

__global__ void __launch_bounds__(64)
intra_kernel(LaGeom g, IntraJob job, int do_edges)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= g.mb_count) return;
    const int mx = idx % g.mb_w, my = idx / g.mb_w;
    if (!do_edges && (mx == 0 || my == 0 || mx == g.mb_w - 1 || my == g.mb_h - 1)) return;
    const int stride = g.lstride;
    const uint8_t *src = job.plane0 + 8 * (mx + my * stride);
    uint2 s[8];
    Nbr n;
#pragma unroll
    for (int y = 0; y < 8; y++) { s[y] = load8u(src + y * stride); n.left[y] = src[y * stride - 1]; }
    n.tl = src[-stride - 1];
    {
        uint2 t0 = load8u(src - stride), t1 = load8u(src - stride + 8);
#pragma unroll
        for (int i = 0; i < 8; i++) { n.top[i] = px_of(t0, i); n.top[8 + i] = px_of(t1, i); }
    }
    Edge e;
    if (job.full) filter_edges_dev(e, n);
    // One loop over the modes around ONE copy of the block metric: the SATD butterflies are
    // most of this kernel's code, and ten inlined copies made it 111 KB of SASS (instruction
    // cache pressure on everything that runs beside it).
    int best = 1 << 30;
    const int nmodes = job.full ? 10 : 3;
#pragma unroll 1
    for (int mode = 0; mode < nmodes; mode++) {
        uint2 pr[8];
        switch (mode) {
        case 0: {   // predict_8x8c_dc
            int s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) { s0 += n.top[i]; s1 += n.top[i + 4]; s2 += n.left[i]; s3 += n.left[i + 4]; }
            uint32_t d0 = splat4((s0 + s2 + 4) >> 3), d1 = splat4((s1 + 2) >> 2), d2 = splat4((s3 + 2) >> 2), d3 = splat4((s1 + s3 + 4) >> 3);
#pragma unroll
            for (int y = 0; y < 8; y++) pr[y] = y < 4 ? make_uint2(d0, d1) : make_uint2(d2, d3);
            break;
        }
        case 1:     // predict_8x8c_h
#pragma unroll
            for (int y = 0; y < 8; y++) pr[y] = make_uint2(splat4(n.left[y]), splat4(n.left[y]));
            break;
        case 2: {   // predict_8x8c_v
            uint2 t = load8u(src - stride);
#pragma unroll
            for (int y = 0; y < 8; y++) pr[y] = t;
            break;
        }
        case 3: {   // predict_8x8c_p
            int H = 0, V = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                int tm = 2 - i < 0 ? n.tl : n.top[2 - i < 0 ? 0 : 2 - i];
                int lm = 2 - i < 0 ? n.tl : n.left[2 - i < 0 ? 0 : 2 - i];
                H += (i + 1) * (n.top[4 + i] - tm);
                V += (i + 1) * (n.left[4 + i] - lm);
            }
            int a = 16 * (n.left[7] + n.top[7]);
            int b = (17 * H + 16) >> 5, c = (17 * V + 16) >> 5;
            int i00 = a - 3 * b - 3 * c + 16;
#pragma unroll
            for (int y = 0; y < 8; y++) {
                uint32_t lo = 0, hi = 0;
#pragma unroll
                for (int x = 0; x < 4; x++) {
                    lo |= (uint32_t)clip_px((i00 + c * y + b * x) >> 5) << (8 * x);
                    hi |= (uint32_t)clip_px((i00 + c * y + b * (x + 4)) >> 5) << (8 * x);
                }
                pr[y] = make_uint2(lo, hi);
            }
            break;
        }
        case 4: pred8x8_build<3>(e, pr); break;
        case 5: pred8x8_build<4>(e, pr); break;
        case 6: pred8x8_build<5>(e, pr); break;
        case 7: pred8x8_build<6>(e, pr); break;
        case 8: pred8x8_build<7>(e, pr); break;
        default: pred8x8_build<8>(e, pr); break;
        }
        best = min(best, mbcmp_rows(job.satd, s, pr));
    }
    job.intra_cost[idx] = (uint16_t)(best + 5 + 4);      // + intra_penalty (5*lambda) + lowres_penalty
}

int launch_intra(cudaStream_t st, const LaGeom &g, const IntraJob &job, int do_edges)
{
    intra_kernel<<<(g.mb_count + 63) / 64, 64, 0, st>>>(g, job, do_edges);
    XV_LAUNCH_CHECK();
    return 0;
}

// sums for i_cost_est[0][0] / i_cost_est_aq[0][0] (+ intra row SATDs)
__global__ void __launch_bounds__(128)
intra_sum_kernel(LaGeom g, IntraSumJob job, int do_edges)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int c = 0, caq = 0;
    if (idx < g.mb_count) {
        const int mx = idx % g.mb_w, my = idx / g.mb_w;
        const bool edge = mx == 0 || my == 0 || mx == g.mb_w - 1 || my == g.mb_h - 1;
        const bool tiny = g.mb_w <= 2 || g.mb_h <= 2;
        if (do_edges || !edge) {
            int ic = job.intra_cost[idx], icaq = ic;
            if (job.aq_on) icaq = (icaq * job.inv_qscale[idx] + 128) >> 8;
            if (job.row_satd) atomicAdd(job.row_satd + my, icaq);
            if (!edge || tiny) { c = ic; caq = icaq; }
        }
    }
    c = warp_sum(c); caq = warp_sum(caq);
    if ((threadIdx.x & 31) == 0) { if (c) atomicAdd(job.result, c); if (caq) atomicAdd(job.result + 1, caq); }
}

int launch_intra_sum(cudaStream_t st, const LaGeom &g, const IntraSumJob &job)
{
    intra_sum_kernel<<<(g.mb_count + 127) / 128, 128, 0, st>>>(g, job, 1);
    XV_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------
// Per-MB cost selection ([x264] slicetype_mb_cost minus the searches).  One thread per MB.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void fetch_ref8(uint2 out[8], const uint8_t *const planes[4], int stride, int pel,
                                           int mvx, int mvy, const WeightDev &w)
{
#pragma unroll
    for (int r = 0; r < 8; r++) out[r] = get_ref_row(planes, stride, pel, mvx, mvy, r, w);
}

__device__ __forceinline__ uint32_t avg_weighted4(uint32_t a, uint32_t b, int wgt)
{
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int pa = (a >> (8 * i)) & 0xff, pb = (b >> (8 * i)) & 0xff;
        r |= (uint32_t)clip_px((pa * wgt + pb * (64 - wgt) + 32) >> 6) << (8 * i);
    }
    return r;
}

__device__ __forceinline__ int bidir_cost(const uint2 fenc[8], uint2 a[8], const uint2 b[8], int wgt, int satd)
{
#pragma unroll
    for (int r = 0; r < 8; r++) {
        if (wgt == 32) { a[r].x = avg4(a[r].x, b[r].x); a[r].y = avg4(a[r].y, b[r].y); }
        else { a[r].x = avg_weighted4(a[r].x, b[r].x, wgt); a[r].y = avg_weighted4(a[r].y, b[r].y, wgt); }
    }
    return mbcmp_rows(satd, fenc, a);
}

__global__ void __launch_bounds__(64)
finalize_kernel(LaGeom g, FinalizeJob job)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int acc_cost = 0, acc_aq = 0, acc_intra = 0;
    if (idx < g.mb_count) {
        const int mx = idx % g.mb_w, my = idx / g.mb_w;
        const bool edge = mx == 0 || my == 0 || mx == g.mb_w - 1 || my == g.mb_h - 1;
        const bool tiny = g.mb_w <= 2 || g.mb_h <= 2;
        if (job.do_edges || !edge) {
            const int stride = g.lstride, pel = 8 * (mx + my * stride);
            int bcost = LA_COST_MAX, list_used = 0;
            int mv0 = 0, mv1 = 0;
            if (job.b_bidir) {
                const WeightDev w0 = {0, 1, 0, 0};
                uint2 fenc[8], ra[8], rb[8];
#pragma unroll
                for (int r = 0; r < 8; r++) fenc[r] = load8u(job.fenc + pel + r * stride);
                int min_x, max_x, min_y, max_y;
                mv_limits(mx, my, g.mb_w, g.mb_h, job.mv_range2, min_x, max_x, min_y, max_y);
                int d0x = 0, d0y = 0, d1x = 0, d1y = 0;
                if (job.ref1_mvs) {
                    const int mvr = job.ref1_mvs[idx];
                    const int rx = mv_x(mvr), ry = mv_y(mvr);
                    d0x = (rx * job.dist_scale_factor + 128) >> 8;
                    d0y = (ry * job.dist_scale_factor + 128) >> 8;
                    d1x = d0x - rx; d1y = d0y - ry;
                    d0x = clip3i(d0x, min_x, max_x); d0y = clip3i(d0y, min_y, max_y);
                    d1x = clip3i(d1x, min_x, max_x); d1y = clip3i(d1y, min_y, max_y);
                    if (!job.subme_gt1) { d0x &= ~1; d0y &= ~1; d1x &= ~1; d1y &= ~1; }
                }
                mv0 = job.mvs0[idx]; mv1 = job.mvs1[idx];
                // Three bidirectional candidates around ONE copy of the fetch + average + metric
                // code (code size): k=0 TRY_BIDIR(dmv[0], dmv[1], 0); k=1 the zero vectors, if the
                // direct vectors were not zero; then the two list costs; k=2 TRY_BIDIR(m[0].mv,
                // m[1].mv, 5) if either is non-zero.  Same comparison order as upstream.
#pragma unroll 1
                for (int k = 0; k < 3; k++) {
                    int ax, ay, bx, by, pen; bool en;
                    if (k == 0) { ax = d0x; ay = d0y; bx = d1x; by = d1y; pen = 0; en = true; }
                    else if (k == 1) { ax = ay = bx = by = 0; pen = 0; en = (d0x | d0y | d1x | d1y) != 0; }
                    else { ax = mv_x(mv0); ay = mv_y(mv0); bx = mv_x(mv1); by = mv_y(mv1); pen = 5; en = (mv0 | mv1) != 0; }
                    if (en) {
                        if (job.subme_gt1) {
                            fetch_ref8(ra, job.fref0, stride, pel, ax, ay, w0);
                            fetch_ref8(rb, job.fref1, stride, pel, bx, by, w0);
                        } else {
#pragma unroll
                            for (int r = 0; r < 8; r++) {
                                ra[r] = load8u(job.fref0[((ax & 2) >> 1) + (ay & 2)] + pel + ((ay >> 2) + r) * stride + (ax >> 2));
                                rb[r] = load8u(job.fref1[((bx & 2) >> 1) + (by & 2)] + pel + ((by >> 2) + r) * stride + (bx >> 2));
                            }
                        }
                        const int c = pen + bidir_cost(fenc, ra, rb, job.bipred_weight, job.satd);
                        if (c < bcost) { bcost = c; list_used = 3; }
                    }
                    if (k == 1) {
                        const int c0 = job.mv_costs0[idx], c1 = job.mv_costs1[idx];
                        if (c0 < bcost) { bcost = c0; list_used = 1; }
                        if (c1 < bcost) { bcost = c1; list_used = 2; }
                    }
                }
            } else if (job.mv_costs0) {
                const int c0 = job.mv_costs0[idx];
                if (c0 < bcost) { bcost = c0; list_used = 1; }
            }
            bcost += 4;                                      // lowres_penalty
            if (job.b_p) {
                const int icost = job.intra_cost[idx];
                const int b_intra = icost < bcost;
                if (b_intra) { bcost = icost; list_used = 0; }
                if (!edge || tiny) acc_intra = b_intra;
            }
            int bcost_aq = bcost;
            if (job.aq_on) bcost_aq = (bcost_aq * job.inv_qscale[idx] + 128) >> 8;
            if (job.row_satd) atomicAdd(job.row_satd + my, bcost_aq);
            if (!edge || tiny) { acc_cost = bcost; acc_aq = bcost_aq; }
            job.lowres_costs[idx] = (uint16_t)(min(bcost, LA_LOWRES_COST_MASK) + (list_used << LA_LOWRES_COST_SHIFT));
        }
    }
    acc_cost = warp_sum(acc_cost); acc_aq = warp_sum(acc_aq); acc_intra = warp_sum(acc_intra);
    if ((threadIdx.x & 31) == 0) {
        if (acc_cost) atomicAdd(job.result, acc_cost);
        if (acc_aq) atomicAdd(job.result + 1, acc_aq);
        if (acc_intra) atomicAdd(job.result + 2, acc_intra);
    }
}

// B evaluations: the three bidirectional candidates are most of the work (six block fetches and
// three SATDs per MB), and the host usually waits for this kernel.  Four lanes share an MB (two
// rows each; the SATD butterflies cross lanes with shuffles), which makes the kernel four times
// wider and its per-thread chain four times shorter.  Same arithmetic, same comparison order.
__global__ void __launch_bounds__(128)
finalize_bidir4_kernel(LaGeom g, FinalizeJob job)
{
    const unsigned FULLM = 0xffffffffu;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int idx = tid >> 2, part = tid & 3;
    const bool in_range = idx < g.mb_count;
    const int mx = in_range ? idx % g.mb_w : 0, my = in_range ? idx / g.mb_w : 0;
    const bool edge = mx == 0 || my == 0 || mx == g.mb_w - 1 || my == g.mb_h - 1;
    const bool tiny = g.mb_w <= 2 || g.mb_h <= 2;
    const bool act = in_range && (job.do_edges || !edge);
    const int stride = g.lstride, pel = 8 * (mx + my * stride), r0 = 2 * part;
    const WeightDev w0 = {0, 1, 0, 0};
    int acc_cost = 0, acc_aq = 0;
    int bcost = LA_COST_MAX, list_used = 0;
    uint2 f0 = make_uint2(0, 0), f1 = f0;
    int d0x = 0, d0y = 0, d1x = 0, d1y = 0, mv0 = 0, mv1 = 0;
    if (act) {
        f0 = load8u(job.fenc + pel + r0 * stride); f1 = load8u(job.fenc + pel + (r0 + 1) * stride);
        int min_x, max_x, min_y, max_y;
        mv_limits(mx, my, g.mb_w, g.mb_h, job.mv_range2, min_x, max_x, min_y, max_y);
        if (job.ref1_mvs) {
            const int mvr = job.ref1_mvs[idx];
            const int rx = mv_x(mvr), ry = mv_y(mvr);
            d0x = (rx * job.dist_scale_factor + 128) >> 8;
            d0y = (ry * job.dist_scale_factor + 128) >> 8;
            d1x = d0x - rx; d1y = d0y - ry;
            d0x = clip3i(d0x, min_x, max_x); d0y = clip3i(d0y, min_y, max_y);
            d1x = clip3i(d1x, min_x, max_x); d1y = clip3i(d1y, min_y, max_y);
            if (!job.subme_gt1) { d0x &= ~1; d0y &= ~1; d1x &= ~1; d1y &= ~1; }
        }
        mv0 = job.mvs0[idx]; mv1 = job.mvs1[idx];
    }
#pragma unroll 1
    for (int k = 0; k < 3; k++) {
        int ax, ay, bx, by, pen; bool en;
        if (k == 0) { ax = d0x; ay = d0y; bx = d1x; by = d1y; pen = 0; en = act; }
        else if (k == 1) { ax = ay = bx = by = 0; pen = 0; en = act && (d0x | d0y | d1x | d1y) != 0; }
        else { ax = mv_x(mv0); ay = mv_y(mv0); bx = mv_x(mv1); by = mv_y(mv1); pen = 5; en = act && (mv0 | mv1) != 0; }
        if (__any_sync(FULLM, en)) {
            uint2 a0 = make_uint2(0, 0), a1 = a0, b0 = a0, b1 = a0;
            if (en) {
                if (job.subme_gt1) {
                    a0 = get_ref_row(job.fref0, stride, pel, ax, ay, r0, w0); a1 = get_ref_row(job.fref0, stride, pel, ax, ay, r0 + 1, w0);
                    b0 = get_ref_row(job.fref1, stride, pel, bx, by, r0, w0); b1 = get_ref_row(job.fref1, stride, pel, bx, by, r0 + 1, w0);
                } else {
                    const uint8_t *pa = job.fref0[((ax & 2) >> 1) + (ay & 2)] + pel + ((ay >> 2) + r0) * stride + (ax >> 2);
                    const uint8_t *pb = job.fref1[((bx & 2) >> 1) + (by & 2)] + pel + ((by >> 2) + r0) * stride + (bx >> 2);
                    a0 = load8u(pa); a1 = load8u(pa + stride); b0 = load8u(pb); b1 = load8u(pb + stride);
                }
                if (job.bipred_weight == 32) {
                    a0.x = avg4(a0.x, b0.x); a0.y = avg4(a0.y, b0.y); a1.x = avg4(a1.x, b1.x); a1.y = avg4(a1.y, b1.y);
                } else {
                    a0.x = avg_weighted4(a0.x, b0.x, job.bipred_weight); a0.y = avg_weighted4(a0.y, b0.y, job.bipred_weight);
                    a1.x = avg_weighted4(a1.x, b1.x, job.bipred_weight); a1.y = avg_weighted4(a1.y, b1.y, job.bipred_weight);
                }
            }
            int c;
            if (job.satd) c = satd_rows4(f0, f1, a0, a1, part & 1);
            else {
                c = __vsadu4(a0.x, f0.x) + __vsadu4(a0.y, f0.y) + __vsadu4(a1.x, f1.x) + __vsadu4(a1.y, f1.y);
                c += __shfl_xor_sync(FULLM, c, 1);
                c += __shfl_xor_sync(FULLM, c, 2);
            }
            c += pen;
            if (en && c < bcost) { bcost = c; list_used = 3; }
        }
        if (k == 1 && act) {
            const int c0 = job.mv_costs0[idx], c1 = job.mv_costs1[idx];
            if (c0 < bcost) { bcost = c0; list_used = 1; }
            if (c1 < bcost) { bcost = c1; list_used = 2; }
        }
    }
    if (act && part == 0) {
        bcost += 4;                                      // lowres_penalty
        int bcost_aq = bcost;
        if (job.aq_on) bcost_aq = (bcost_aq * job.inv_qscale[idx] + 128) >> 8;
        if (job.row_satd) atomicAdd(job.row_satd + my, bcost_aq);
        if (!edge || tiny) { acc_cost = bcost; acc_aq = bcost_aq; }
        job.lowres_costs[idx] = (uint16_t)(min(bcost, LA_LOWRES_COST_MASK) + (list_used << LA_LOWRES_COST_SHIFT));
    }
    acc_cost = warp_sum(acc_cost); acc_aq = warp_sum(acc_aq);
    if ((threadIdx.x & 31) == 0) {
        if (acc_cost) atomicAdd(job.result, acc_cost);
        if (acc_aq) atomicAdd(job.result + 1, acc_aq);
    }
}

int launch_finalize(cudaStream_t st, const LaGeom &g, const FinalizeJob &job)
{
    if (job.b_bidir) finalize_bidir4_kernel<<<(4 * g.mb_count + 127) / 128, 128, 0, st>>>(g, job);
    else finalize_kernel<<<(g.mb_count + 63) / 64, 64, 0, st>>>(g, job);
    XV_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------
// Weights ([x264] weight_cost_luma at mv 0, x264_weight_scale_plane)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
weight_cost_kernel(LaGeom g, WeightCostJob job)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int c = 0;
    if (idx < g.mb_count) {
        const int mx = idx % g.mb_w, my = idx / g.mb_w;
        const int pel = 8 * (mx + my * g.lstride);
        uint2 a[8], f[8];
#pragma unroll
        for (int r = 0; r < 8; r++) {
            a[r] = load8u(job.ref + pel + r * g.lstride);
            f[r] = load8u(job.fenc + pel + r * g.lstride);
            if (job.w.on) { a[r].x = weight_word(job.w, a[r].x); a[r].y = weight_word(job.w, a[r].y); }
        }
        c = min(mbcmp_rows(job.satd, a, f), (int)job.intra_cost[idx]);
    }
    c = warp_sum(c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(job.result, (unsigned)c);
}

int launch_weight_cost(cudaStream_t st, const LaGeom &g, const WeightCostJob &job)
{
    weight_cost_kernel<<<(g.mb_count + 127) / 128, 128, 0, st>>>(g, job);
    XV_LAUNCH_CHECK();
    return 0;
}

__global__ void __launch_bounds__(256)
weight_plane_kernel(uint8_t *dst, const uint8_t *src, int nwords, WeightDev w)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nwords) ((uint32_t *)dst)[i] = weight_word(w, ((const uint32_t *)src)[i]);
}

int launch_weight_plane(cudaStream_t st, const LaGeom &g, uint8_t *dst, const uint8_t *src, WeightDev w)
{
    const int nwords = g.lplane / 4;
    weight_plane_kernel<<<(nwords + 255) / 256, 256, 0, st>>>(dst, src, nwords, w);
    XV_LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------
// MB-tree.  The reference accumulates into uint16 with per-add saturation at 32767; all
// addends are >= 0, so that equals min(sum, 32767): accumulate with 32-bit atomics into a
// shadow array and clamp when the value is read.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void propagate_mb(const LaGeom &g, const PropagateJob &job, int idx)
{
    const int i = idx % g.mb_w, mb_y = idx / g.mb_w;
    const int lc = job.lowres_costs[idx];
    const int intra_cost = job.intra_cost[idx];
    const int inter_cost = min(intra_cost, lc & LA_LOWRES_COST_MASK);
    const int pin = job.propagate_in ? min(__ldcg(job.propagate_in + idx), 32767) : 0;
    const float propagate_intra = (float)(intra_cost * (int)job.inv_qscale[idx]);
    const float propagate_amount = (float)pin + propagate_intra * job.fps_factor;
    const float propagate_num = (float)(intra_cost - inter_cost);
    const float propagate_denom = (float)intra_cost;
    const int amount = min((int)(propagate_amount * propagate_num / propagate_denom + 0.5f), 32767);
    const int lists_used = lc >> LA_LOWRES_COST_SHIFT;
    const unsigned width = g.mb_w, height = g.mb_h, stride = g.mb_w;
#pragma unroll
    for (int list = 0; list < 2; list++) {
        if (list == 1 && !job.b_bidir) break;
        if (!(lists_used & (1 << list))) continue;
        int *ref = list ? job.ref1_cost : job.ref0_cost;
        int listamount = amount;
        if (lists_used == 3) listamount = (listamount * (list ? 64 - job.bipred_weight : job.bipred_weight) + 32) >> 6;
        const int mv = (list ? job.mvs1 : job.mvs0)[idx];
        if (!mv) { if (listamount) atomicAdd(ref + mb_y * stride + i, listamount); continue; }
        int x = mv_x(mv), y = mv_y(mv);
        const unsigned mbx = (unsigned)((x >> 5) + i), mby = (unsigned)((y >> 5) + mb_y);
        // unsigned 32-bit index arithmetic on purpose: mbx == -1 wraps so that idx0 + 1 is column 0
        const unsigned idx0 = mbx + mby * stride, idx2 = idx0 + stride;
        x &= 31; y &= 31;
        const int w0 = ((32 - y) * (32 - x) * listamount + 512) >> 10, w1 = ((32 - y) * x * listamount + 512) >> 10;
        const int w2 = (y * (32 - x) * listamount + 512) >> 10, w3 = (y * x * listamount + 512) >> 10;
        if (mbx < width - 1 && mby < height - 1) {
            if (w0) atomicAdd(&ref[idx0], w0);
            if (w1) atomicAdd(&ref[idx0 + 1u], w1);
            if (w2) atomicAdd(&ref[idx2], w2);
            if (w3) atomicAdd(&ref[idx2 + 1u], w3);
        } else {
            if (mby < height) {
                if (mbx < width && w0) atomicAdd(&ref[idx0], w0);
                if (mbx + 1 < width && w1) atomicAdd(&ref[idx0 + 1u], w1);
            }
            if (mby + 1 < height) {
                if (mbx < width && w2) atomicAdd(&ref[idx2], w2);
                if (mbx + 1 < width && w3) atomicAdd(&ref[idx2 + 1u], w3);
            }
        }
    }
}

__global__ void __launch_bounds__(128)
propagate_kernel(LaGeom g, PropagateJob job)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < g.mb_count) propagate_mb(g, job, idx);
}

int launch_propagate(cudaStream_t st, const LaGeom &g, const PropagateJob &job)
{
    propagate_kernel<<<(g.mb_count + 127) / 128, 128, 0, st>>>(g, job);
    XV_LAUNCH_CHECK();
    return 0;
}

__device__ __forceinline__ void tree_finish_mb(const TreeFinishJob &job, int idx)
{
    const int intra_cost = ((int)job.intra_cost[idx] * (int)job.inv_qscale[idx] + 128) >> 8;
    if (intra_cost) {
        const int propagate_cost = (min(__ldcg(job.propagate + idx), 32767) * job.fps_factor + 128) >> 8;
        const float log2_ratio = dev_log2(job.log2_lut, intra_cost + propagate_cost) - dev_log2(job.log2_lut, intra_cost) + job.weightdelta;
        job.qp_offset[idx] = job.qp_offset_aq[idx] - job.strength * log2_ratio;
    }
}

__global__ void __launch_bounds__(128)
tree_finish_kernel(LaGeom g, TreeFinishJob job)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < g.mb_count) tree_finish_mb(job, idx);
}

// A whole macroblock_tree walk in ONE launch: the steps are strictly sequential (each
// propagate reads what the previous ones accumulated), so instead of one tiny kernel per step
// a single 8-CTA thread-block cluster executes the list and separates the steps with the
// hardware cluster barrier (release/acquire at cluster scope orders the global atomics).
#define TREE_CLUSTER 8
#define TREE_THREADS 1024
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(TREE_CLUSTER, 1, 1) __launch_bounds__(TREE_THREADS)
tree_chain_kernel(LaGeom g, const TreeStep *steps, int nsteps, const float *log2_lut)
{
    const int tid = blockIdx.x * TREE_THREADS + threadIdx.x, nthr = TREE_CLUSTER * TREE_THREADS;
    for (int s = 0; s < nsteps; s++) {
        const TreeStep &st = steps[s];          // uniform across the cluster, so is st.sync
        if (st.op == 0) {
            for (int i = tid; i < st.n; i += nthr) st.zero[i] = 0;
        } else if (st.op == 1) {
            for (int i = tid; i < g.mb_count; i += nthr) propagate_mb(g, st.prop, i);
        } else {
            TreeFinishJob fj;
            fj.propagate = st.fin_propagate; fj.intra_cost = st.fin_intra; fj.inv_qscale = st.fin_invq;
            fj.qp_offset_aq = st.fin_qp_aq; fj.qp_offset = st.fin_qp; fj.fps_factor = st.fin_fps_factor;
            fj.weightdelta = st.fin_weightdelta; fj.strength = st.fin_strength; fj.log2_lut = log2_lut;
            for (int i = tid; i < g.mb_count; i += nthr) tree_finish_mb(fj, i);
        }
        if (st.sync) cluster_sync_all();
    }
}

int launch_tree_chain(cudaStream_t stream, const LaGeom &g, const TreeStep *steps_dev, int nsteps, const float *log2_lut)
{
    if (nsteps <= 0) return 0;
    tree_chain_kernel<<<TREE_CLUSTER, TREE_THREADS, 0, stream>>>(g, steps_dev, nsteps, log2_lut);
    XV_LAUNCH_CHECK();
    return 0;
}

int launch_tree_finish(cudaStream_t st, const LaGeom &g, const TreeFinishJob &job)
{
    tree_finish_kernel<<<(g.mb_count + 127) / 128, 128, 0, st>>>(g, job);
    XV_LAUNCH_CHECK();
    return 0;
}

} // namespace xv
