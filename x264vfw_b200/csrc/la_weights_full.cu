// SURVEY 8(f) row 3, remainder: [x264] x264_weights_analyse( h, fenc, ref, 0 ) -- the ENCODER-side explicit weight
// analysis of a P frame against its nearest reference ([x264] encoder/slicetype.c; reached from x264_encoder_encode,
// codec.c:1693, through reference_build_list when weightp >= 1).
//
// Upstream scores one (scale, offset) candidate after the other, each score a pass over the frame:
//   luma    on the LOWRES planes, reference compensated 8x8 by 8x8 with the lookahead's list-0 vectors of that
//           distance (weight_cost_init_luma), score = sum min(mbcmp 8x8, intra cost) (weight_cost_luma);
//   chroma  at full resolution on the NV12 planes, reference compensated by mc_chroma with the same lowres vector
//           (weight_cost_init_chroma), score = sum |sum of block differences| (pixel_asd8, weight_cost_chroma);
// in a window of scales and offsets that grows with subme (weight_check_distance), with an early exit per scale.
// Every candidate of a plane is known before the first score (they follow from the frame statistics), so here ONE launch
// per plane scores all of them -- each thread compensates its block once and applies every candidate to it in registers --
// and the host then walks upstream's loops over the finished scores, early exit included: same decisions, 2 launches and
// 2 synchronisations per frame instead of up to 3 x 45 passes.
#include "common.cuh"
#include "la_kernels.h"
#include "la_common.cuh"
#include "../../include/x264vfw_cuda.h"
#include <math.h>
#include <vector>

namespace xv {

#define WA_MAX_CAND 48          // 1 unweighted + (2*4+1) scales x (2*2+1) offsets at subme 11

struct WaJob {
    // luma
    const uint8_t *fenc_l0, *ref_l0;        // pixel (0,0) of lowres plane 0; the reference's 4 planes are lplane apart
    const int *mvs;                         // packed lowres quarter-pel vectors per MB, or null (never searched)
    const uint16_t *intra_cost;
    int satd;
    // chroma (NV12, padded to mod 16)
    const uint8_t *fenc_uv, *ref_uv;
    int uv_stride, cw, ch;
    int ncand[3];
    WeightDev cand[3][WA_MAX_CAND];
    unsigned *result;                       // [3][WA_MAX_CAND], zeroed by the launcher
};

__device__ __forceinline__ int wa_warp_sum(int v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(128) wa_luma_kernel(const LaGeom g, const __grid_constant__ WaJob job)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = idx < g.mb_count;
    uint2 a[8], f[8];
    int intra = 0;
    if (live) {
        const int mx = idx % g.mb_w, my = idx / g.mb_w;
        const int pel = 8 * (mx + my * g.lstride);
        const WeightDev none = {0, 1, 0, 0};
        const int mv = job.mvs ? job.mvs[idx] : 0;
#pragma unroll
        for (int r = 0; r < 8; r++) {
            // weight_cost_init_luma: mc_luma( ..., mvx + (x << 2), mvy + (y << 2), 8, 8 ) on the reference's lowres planes
            a[r] = job.mvs ? get_ref_row_ps(job.ref_l0, g.lplane, g.lstride, pel, mv_x(mv), mv_y(mv), r, none)
                           : load8u(job.ref_l0 + pel + r * g.lstride);
            f[r] = load8u(job.fenc_l0 + pel + r * g.lstride);
        }
        intra = job.intra_cost[idx];
    }
    for (int c = 0; c < job.ncand[0]; c++) {
        const WeightDev w = job.cand[0][c];
        int cost = 0;
        if (live) {
            uint2 b[8];
#pragma unroll
            for (int r = 0; r < 8; r++) {
                b[r] = a[r];
                if (w.on) { b[r].x = weight_word(w, b[r].x); b[r].y = weight_word(w, b[r].y); }
            }
            cost = min(mbcmp_rows(job.satd, b, f), intra);
        }
        cost = wa_warp_sum(cost);
        if ((threadIdx.x & 31) == 0 && cost) atomicAdd(job.result + c, (unsigned)cost);
    }
}

__device__ __forceinline__ int wa_uv(const uint8_t *uv, int stride, int cw, int ch, int x, int y, int comp)
{
    x = min(max(x, 0), cw - 1);             // == the replicated pairs x264_frame_expand_border_chroma leaves past the edge
    y = min(max(y, 0), ch - 1);
    return __ldg(uv + (ptrdiff_t)y * stride + 2 * x + comp);
}

// thread = (MB, chroma component)
__global__ void __launch_bounds__(128) wa_chroma_kernel(const LaGeom g, const __grid_constant__ WaJob job)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int comp = blockIdx.y;
    const bool live = t < g.mb_count;
    uint2 blk[8];                           // the compensated 8x8 reference block
    int fsum = 0;
    if (live) {
        const int bx = 8 * (t % g.mb_w), by = 8 * (t / g.mb_w);
        if (job.mvs) {
            // weight_cost_init_chroma: mc_chroma( ..., mvx, 2*mvy >> v_shift, 8, 8 ) with the MB's LOWRES vector
            const int mv = job.mvs[t];
            const int mvx = mv_x(mv), mvy = mv_y(mv);
            const int d8x = mvx & 7, d8y = mvy & 7;
            const int cA = (8 - d8x) * (8 - d8y), cB = d8x * (8 - d8y), cC = (8 - d8x) * d8y, cD = d8x * d8y;
            const int x0 = bx + (mvx >> 3), y0 = by + (mvy >> 3);
            int top[9], bot[9];
#pragma unroll
            for (int x = 0; x < 9; x++) top[x] = wa_uv(job.ref_uv, job.uv_stride, job.cw, job.ch, x0 + x, y0, comp);
#pragma unroll
            for (int y = 0; y < 8; y++) {
#pragma unroll
                for (int x = 0; x < 9; x++) bot[x] = wa_uv(job.ref_uv, job.uv_stride, job.cw, job.ch, x0 + x, y0 + y + 1, comp);
                uint32_t lo = 0, hi = 0;
#pragma unroll
                for (int x = 0; x < 8; x++) {
                    const uint32_t p = (uint32_t)(cA * top[x] + cB * top[x + 1] + cC * bot[x] + cD * bot[x + 1] + 32) >> 6;
                    if (x < 4) lo |= p << (8 * x); else hi |= p << (8 * (x - 4));
                }
                blk[y] = make_uint2(lo, hi);
#pragma unroll
                for (int x = 0; x < 9; x++) top[x] = bot[x];
            }
        } else {
#pragma unroll
            for (int y = 0; y < 8; y++) {
                uint32_t lo = 0, hi = 0;
#pragma unroll
                for (int x = 0; x < 8; x++) {
                    const uint32_t p = __ldg(job.ref_uv + (ptrdiff_t)(by + y) * job.uv_stride + 2 * (bx + x) + comp);
                    if (x < 4) lo |= p << (8 * x); else hi |= p << (8 * (x - 4));
                }
                blk[y] = make_uint2(lo, hi);
            }
        }
#pragma unroll
        for (int y = 0; y < 8; y++)
#pragma unroll
            for (int x = 0; x < 8; x++) fsum += __ldg(job.fenc_uv + (ptrdiff_t)(by + y) * job.uv_stride + 2 * (bx + x) + comp);
    }
    const int plane = 1 + comp;
    for (int c = 0; c < job.ncand[plane]; c++) {
        const WeightDev w = job.cand[plane][c];
        int cost = 0;
        if (live) {
            int s = 0;                      // pixel_asd8 of the (weighted) block against the source block
#pragma unroll
            for (int y = 0; y < 8; y++)
#pragma unroll
                for (int x = 0; x < 8; x++) {
                    const int p = px_of(blk[y], x);
                    s += w.on ? weight_px_dev(w, p) : p;
                }
            cost = abs(s - fsum);
        }
        cost = wa_warp_sum(cost);
        if ((threadIdx.x & 31) == 0 && cost) atomicAdd(job.result + plane * WA_MAX_CAND + c, (unsigned)cost);
    }
}

// ---- host: the loops of x264_weights_analyse over precomputed scores ------------------------------------------------

static int ue_bits(unsigned v) { v += 1; int n = 0; while (v >> (n + 1)) n++; return 2 * n + 1; }           // bs_size_ue
static int se_bits(int v) { int t = 1 - v * 2; if (t < 0) t = v * 2; int n = 0; while (t >> (n + 1)) n++; return 2 * n + 1; }   // bs_size_se
static int header_cost(const WeightDev &w, int b_chroma)       // weight_slice_header_cost, lambda(QP 12) = 1, one slice
{
    const int lambda = b_chroma ? 4 : 1;
    return lambda * (10 + ue_bits(w.denom) * (2 - b_chroma) + 2 * (se_bits(w.scale) + se_bits(w.offset)));
}
static int iclip(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }
static float fclip(float v, float lo, float hi) { return v < lo ? lo : v > hi ? hi : v; }

struct ScaleStep { int cur_scale, start_offset, end_offset, first; };   // first: index of start_offset's candidate

struct PlanePlan {
    bool skip = false;          // early termination / invalid: weight stays off without scoring
    int mindenom = 0, minscale = 0;
    std::vector<ScaleStep> steps;
};

int weights_analyse_full(cudaStream_t st, const LaGeom &g, const x264vfw_cuda_weights_in *in, int32_t out[3][4], float *cost_delta,
                         unsigned *d_result, unsigned *h_result)
{
    static const uint8_t check_distance[12][2] = {{0, 0}, {0, 0}, {0, 1}, {0, 1}, {0, 1}, {0, 1}, {0, 1}, {1, 1}, {1, 1}, {2, 1}, {2, 1}, {4, 2}};
    const float epsilon = 1.f / 128.f;
    WeightDev weights[3] = {{0, 1, 0, 0}, {0, 1, 0, 0}, {0, 1, 0, 0}};
    float guess_scale[3], fenc_mean[3], ref_mean[3];
    const int dims[2] = {g.luma_h * g.luma_w, (g.luma_h / 2) * (g.luma_w / 2)};
    for (int p = 0; p < 3; p++) {
        const int zero_bias = !in->ref_ssd[p];
        const float fenc_var = in->fenc_ssd[p] + zero_bias, ref_var = in->ref_ssd[p] + zero_bias;
        guess_scale[p] = sqrtf(fenc_var / ref_var);
        fenc_mean[p] = (float)(in->fenc_sum[p] + zero_bias) / dims[!!p] / 1;
        ref_mean[p] = (float)(in->ref_sum[p] + zero_bias) / dims[!!p] / 1;
    }
    int chroma_denom = 7;
    while (chroma_denom > 0) {
        const float thresh = 127.f / (1 << chroma_denom);
        if (guess_scale[1] < thresh && guess_scale[2] < thresh) break;
        chroma_denom--;
    }
    const int subme = iclip(in->subme, 0, 11);
    const int scale_dist = check_distance[subme][0], offset_dist = check_distance[subme][1];
    if (cost_delta) *cost_delta = 0;

    WaJob job;
    memset(&job, 0, sizeof(job));
    job.fenc_l0 = in->fenc_lowres + g.lorigin; job.ref_l0 = in->ref_lowres + g.lorigin;
    job.mvs = (const int *)in->lowres_mvs; job.intra_cost = in->intra_cost; job.satd = in->subme > 1;   // mbcmp_init
    job.fenc_uv = in->fenc_uv; job.ref_uv = in->ref_uv; job.uv_stride = in->uv_stride;
    job.cw = 8 * g.mb_w; job.ch = 8 * g.mb_h;
    job.result = d_result;

    // candidates of one plane, in upstream's loop order, the unweighted score first
    PlanePlan plan[3];
    auto make_plan = [&](int p) -> bool {            // false: stop looking at further planes (upstream's `break`)
        PlanePlan &pl = plan[p];
        job.ncand[p] = 0;
        if (fabsf(ref_mean[p] - fenc_mean[p]) < 0.5f && fabsf(1.f - guess_scale[p]) < epsilon) { pl.skip = true; return true; }
        if (p) {
            weights[p].denom = chroma_denom;
            weights[p].scale = iclip((int)round(guess_scale[p] * (1 << chroma_denom)), 0, 255);
            if (weights[p].scale > 127) { weights[1].on = weights[2].on = 0; pl.skip = true; return false; }
        } else {
            int scale = (int)round(guess_scale[0] * 128), denom = 7;       // x264_weight_get_h264
            while (denom > 0 && scale > 127) { denom--; scale >>= 1; }
            weights[0].scale = scale < 127 ? scale : 127; weights[0].denom = denom; weights[0].offset = 0;
        }
        pl.mindenom = weights[p].denom; pl.minscale = weights[p].scale;
        int n = 0;
        job.cand[p][n++] = WeightDev{0, 1, 0, 0};
        const int start_scale = iclip(pl.minscale - scale_dist, 0, 127), end_scale = iclip(pl.minscale + scale_dist, 0, 127);
        for (int i_scale = start_scale; i_scale <= end_scale; i_scale++) {
            int cur_scale = i_scale;
            int cur_offset = fenc_mean[p] - ref_mean[p] * cur_scale / (1 << pl.mindenom) + 0.5f * 0;
            if (cur_offset < -128 || cur_offset > 127) {
                cur_offset = iclip(cur_offset, -128, 127);
                cur_scale = fclip((1 << pl.mindenom) * (fenc_mean[p] - cur_offset) / ref_mean[p] + 0.5f, 0, 127);
            }
            ScaleStep s = {cur_scale, iclip(cur_offset - offset_dist, -128, 127), iclip(cur_offset + offset_dist, -128, 127), n};
            for (int i_off = s.start_offset; i_off <= s.end_offset; i_off++) job.cand[p][n++] = WeightDev{1, cur_scale, pl.mindenom, i_off};
            pl.steps.push_back(s);
        }
        job.ncand[p] = n;
        return true;
    };
    auto decide = [&](int p, const unsigned *score) {
        PlanePlan &pl = plan[p];
        if (pl.skip) { weights[p] = WeightDev{0, 1, 0, 0}; return; }
        const unsigned origscore = score[0];
        unsigned minscore = origscore;
        if (!minscore) return;                         // upstream `continue`s with the guess left in place, weightfn unset
        int found = 0, minscale = pl.minscale, mindenom = pl.mindenom, minoff = 0;
        for (const ScaleStep &s : pl.steps)
            for (int i_off = s.start_offset; i_off <= s.end_offset; i_off++) {
                const WeightDev &w = job.cand[p][s.first + i_off - s.start_offset];
                const unsigned sc = score[s.first + i_off - s.start_offset] + header_cost(w, p != 0);
                if (sc < minscore) { minscore = sc; minscale = s.cur_scale; minoff = i_off; found = 1; }
                if (minoff == s.start_offset && i_off != s.start_offset) break;
            }
        if (!p)
            while (mindenom > 0 && !(minscale & 1)) { mindenom--; minscale >>= 1; }
        if (!found || (minscale == 1 << mindenom && minoff == 0) || (float)minscore / origscore > 0.998f) { weights[p] = WeightDev{0, 1, 0, 0}; return; }
        weights[p] = WeightDev{1, minscale, mindenom, minoff};
        if (in->weightp == -1 && !p && cost_delta) *cost_delta = (float)minscore / origscore;       // X264_WEIGHTP_FAKE
    };

    const int nblk = (g.mb_count + 127) / 128;
    // luma
    make_plan(0);
    if (!plan[0].skip) {
        XV_CUDA_OK(cudaMemsetAsync(d_result, 0, sizeof(unsigned) * 3 * WA_MAX_CAND, st));
        wa_luma_kernel<<<nblk, 128, 0, st>>>(g, job);
        XV_LAUNCH_CHECK();
        XV_CUDA_OK(cudaMemcpyAsync(h_result, d_result, sizeof(unsigned) * WA_MAX_CAND, cudaMemcpyDeviceToHost, st));
        XV_CUDA_OK(cudaStreamSynchronize(st));
    }
    decide(0, h_result);
    if (plan[0].skip || !h_result[0]) weights[0].on = 0;
    // chroma: only when luma found a weight
    if (weights[0].on) {
        bool go = make_plan(1);
        if (go) go = make_plan(2); else plan[2].skip = true;
        const bool any = (!plan[1].skip) || (go && !plan[2].skip);
        if (any) {
            if (plan[1].skip) job.ncand[1] = 0;
            if (plan[2].skip) job.ncand[2] = 0;
            wa_chroma_kernel<<<dim3(nblk, 2), 128, 0, st>>>(g, job);
            XV_LAUNCH_CHECK();
            XV_CUDA_OK(cudaMemcpyAsync(h_result, d_result, sizeof(unsigned) * 3 * WA_MAX_CAND, cudaMemcpyDeviceToHost, st));
            XV_CUDA_OK(cudaStreamSynchronize(st));
        }
        decide(1, h_result + WA_MAX_CAND);
        if (!plan[1].skip && !h_result[WA_MAX_CAND]) weights[1].on = 0;
        if (go) {
            decide(2, h_result + 2 * WA_MAX_CAND);
            if (!plan[2].skip && !h_result[2 * WA_MAX_CAND]) weights[2].on = 0;
        } else
            weights[1].on = weights[2].on = 0;
    }
    // optimise and unify the chroma denominator
    if (weights[1].on || weights[2].on) {
        int denom = weights[1].on ? weights[1].denom : weights[2].denom;
        const bool both = weights[1].on && weights[2].on;
        while ((!both && denom == 7) ||
               (denom > 0 && !(weights[1].on && (weights[1].scale & 1)) && !(weights[2].on && (weights[2].scale & 1)))) {
            denom--;
            for (int i = 1; i <= 2; i++)
                if (weights[i].on) { weights[i].scale >>= 1; weights[i].denom = denom; }
        }
    }
    for (int i = 0; i < 3; i++) {
        if (!weights[i].on) weights[i] = WeightDev{0, 1, 0, 0};
        out[i][0] = weights[i].on; out[i][1] = weights[i].scale; out[i][2] = weights[i].denom; out[i][3] = weights[i].offset;
    }
    return 0;
}

} // namespace xv

using namespace xv;

extern "C" int x264vfw_cuda_weights_analyse(x264vfw_cuda_ctx *ctx, const x264vfw_cuda_weights_in *in, int32_t out[3][4], float *cost_delta)
{
    if (!ctx || !in || !out || !in->fenc_lowres || !in->ref_lowres || !in->intra_cost || !in->fenc_uv || !in->ref_uv) { set_error("null argument"); return -1; }
    if (in->width <= 0 || in->height <= 0 || (in->width & 1) || (in->height & 1)) { set_error("bad size"); return -1; }
    Ctx *c = (Ctx *)ctx;
    XV_CUDA_OK(cudaSetDevice(c->device));
    x264vfw_cuda_lowres_geom lg;
    x264vfw_cuda_lowres_geometry(&lg, in->width, in->height);
    LaGeom g;
    g.width = in->width; g.height = in->height; g.mb_w = lg.mb_w; g.mb_h = lg.mb_h; g.mb_count = lg.mb_w * lg.mb_h;
    g.luma_w = lg.luma_w; g.luma_h = lg.luma_h; g.lw = lg.lw; g.lh = lg.lh; g.lstride = lg.lstride; g.lplane = lg.lplane_bytes; g.lorigin = lg.lorigin;
    unsigned *d_res = nullptr, *h_res = nullptr;
    XV_CUDA_OK(cudaMalloc((void **)&d_res, sizeof(unsigned) * 3 * WA_MAX_CAND));
    if (cudaMallocHost((void **)&h_res, sizeof(unsigned) * 3 * WA_MAX_CAND) != cudaSuccess) { cudaFree(d_res); set_error("pinned allocation failed"); return -1; }
    memset(h_res, 0, sizeof(unsigned) * 3 * WA_MAX_CAND);
    const int rc = weights_analyse_full(c->stream, g, in, out, cost_delta, d_res, h_res);
    cudaFree(d_res); cudaFreeHost(h_res);
    return rc;
}
