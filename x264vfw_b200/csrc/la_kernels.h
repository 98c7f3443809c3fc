// Launch descriptors for the lookahead kernels (host <-> la_kernels.cu / la_me_kernel.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

struct x264vfw_cuda_weights_in;     // include/x264vfw_cuda.h

namespace xv {

struct LaGeom {
    int width, height;            // display size
    int mb_w, mb_h, mb_count;
    int luma_w, luma_h;           // mod-16 size
    int lw, lh, lstride, lplane, lorigin;
};

struct WeightDev { int on, scale, denom, offset; };

// ---- adaptive quant statistics ([x264] x264_adaptive_quant_frame) -------------------------
struct AqJob {
    const uint8_t *y, *u, *v; int y_stride, c_stride;
    int c_step;                   // bytes between chroma samples: 1 planar, 2 interleaved (NV12: v = u + 1)
    int chroma_format;            // 1: 4:2:0, 2: 4:2:2, 3: 4:4:4; 0: luma only
    int aq_on; float strength;    // strength: aq-mode 1 (already * 1.0397)
    int aq_mode; float aq_strength; // aq-mode 2 / 3 (auto-variance): raw rc.f_aq_strength
    float *qp_offset, *qp_offset_aq; uint16_t *inv_qscale;
    unsigned long long *stats;    // [6]: sum[3], ssd[3] (raw, before mean removal)
    const float *log2_lut; const uint8_t *exp2_lut;
};
int launch_aq(cudaStream_t st, const LaGeom &g, const AqJob &job);
int launch_aq_auto(cudaStream_t st, const LaGeom &g, const AqJob &job);   // second loop of aq-mode 2 / 3 only

// ---- intra cost ([x264] slicetype_mb_cost, lowres_intra_mb) --------------------------------
struct IntraJob {
    const uint8_t *plane0;        // lowres plane 0 origin
    uint16_t *intra_cost;
    int full;                     // subme > 1: also plane + the six 8x8 directional modes
    int satd;                     // mbcmp = SATD (subme > 1) else SAD
};
int launch_intra(cudaStream_t st, const LaGeom &g, const IntraJob &job, int do_edges);

// ---- motion search ([x264] x264_me_search_ref + refine_subpel through slicetype_mb_cost) ---
struct MeJob {
    const uint8_t *fenc;          // plane 0 origin of the frame being costed
    const uint8_t *fref[4];       // the four phase planes of the reference (origins)
    const uint8_t *fref_w;        // weighted plane 0 origin, or fref[0]
    WeightDev w;
    int *mvs;                     // packed int16x2 per MB (x | y<<16)
    int *mv_costs;
    int2 *rec;                    // per-MB {mv, epoch} records: the inter-row channel of this launch
    int *ticket;                  // row hand-out counter (zeroed by the launcher)
    const int *guess;             // speculative search: a similar MV field (e.g. the same search of the previous frame) or null
    int guess_num, guess_den;     // ... scaled by num/den (a field of another temporal distance)
    int4 *assumed;                // speculative search: per MB, the four neighbour MVs its current result was computed from
};
#define XV_ME_MAX_JOBS 8
struct MeParams {
    int njobs;
    int epoch;                    // unique per launch; a record is valid iff its tag equals it
    int rows_in_flight;           // warps per search (each walks several rows); 0 = one warp per row
    MeJob job[XV_ME_MAX_JOBS];
    int bands;                    // lookahead_threads
    int do_edges;
    int mv_range2;                // 2 * analyse.i_mv_range
    int me_hex;                   // lowres me method: 1 hex, 0 dia
    int subpel_refine;            // 4 or 2 (lowres_context_init)
    int satd;                     // mbcmp is SATD
    int me_range;
    const uint16_t *cost_mv;      // centred table
    int variant;                  // 0: plain wavefront (me_wavefront_kernel); 1: speculative parallel passes + verification wavefront
    int npasses;                  // variant 1: parallel (Jacobi) passes before the verification
    int nrelax;                   // variant 1: row-relaxation passes after them (la_me2_kernel.cu: RELAX)
    int *stats;                   // optional device counters: [0] kept, [1] re-searched in order, [2..5] searched in pass 0..3,
                                  // [6] SAD 8x8 evaluations, [7] SATD 8x8 evaluations
    int force_miss;               // diagnostics: the verification keeps nothing (every MB re-searched in order) = 0 % hit rate
};
int launch_me(cudaStream_t st, const LaGeom &g, const MeParams &p);                    // plain wavefront
int launch_me_pass(cudaStream_t st, const LaGeom &g, const MeParams &p, int pass);     // one speculative parallel pass
int launch_me_verify(cudaStream_t st, const LaGeom &g, const MeParams &p, int relax_pass = -1);   // exact verification wavefront (relax_pass >= 1: a row-relaxation pass)

// ---- per-MB cost selection + frame accumulators ([x264] rest of slicetype_mb_cost) ---------
struct FinalizeJob {
    const uint8_t *fenc;
    const uint8_t *fref0[4], *fref1[4];
    const int *mvs0, *mvs1;       // this frame's (list, dist) results; null if b==p0 / b==p1
    const int *mv_costs0, *mv_costs1;
    const int *ref1_mvs;          // frames[p1]->lowres_mvs[0][p1-p0-1] if already searched, else null
    const uint16_t *intra_cost;
    const uint16_t *inv_qscale;
    uint16_t *lowres_costs;
    int *row_satd;                // [mb_h] or null
    int *result;                  // [0]=cost_est [1]=cost_est_aq [2]=intra_mbs (zeroed by the launcher)
    int b_bidir, b_p;             // b < p1 ; b == p1 (P evaluation)
    int dist_scale_factor, bipred_weight;
    int aq_on, subme_gt1, satd;
    int mv_range2, do_edges;
};
int launch_finalize(cudaStream_t st, const LaGeom &g, const FinalizeJob &job);

// sums of a uint16 per-MB array over interior / all MBs: intra cost_est[0][0]
struct IntraSumJob {
    const uint16_t *intra_cost, *inv_qscale; int aq_on;
    int *result;                  // [0]=cost_est, [1]=cost_est_aq
    int *row_satd;
};
int launch_intra_sum(cudaStream_t st, const LaGeom &g, const IntraSumJob &job);

// ---- weights ([x264] weight_cost_luma, x264_weight_scale_plane) ----------------------------
struct WeightCostJob {
    const uint8_t *fenc, *ref;    // plane 0 origins
    const uint16_t *intra_cost;
    WeightDev w;                  // w.on == 0: unweighted score
    int satd;
    unsigned *result;             // one accumulator (zeroed by the launcher)
};
int launch_weight_cost(cudaStream_t st, const LaGeom &g, const WeightCostJob &job);
int launch_weight_plane(cudaStream_t st, const LaGeom &g, uint8_t *dst, const uint8_t *src_plane_base, WeightDev w);

// ---- encoder-side weight analysis ([x264] x264_weights_analyse(..., 0); la_weights_full.cu) ----
// d_result / h_result: 3 * 48 unsigned each (device / pinned host), h_result zeroed by the caller
int weights_analyse_full(cudaStream_t st, const LaGeom &g, const ::x264vfw_cuda_weights_in *in, int32_t out[3][4], float *cost_delta,
                         unsigned *d_result, unsigned *h_result);

// ---- mb-tree ([x264] mbtree_propagate_cost/_list, macroblock_tree_finish) ------------------
struct PropagateJob {
    const int *propagate_in;      // this frame's accumulated cost (32-bit shadow), null if not referenced
    const uint16_t *intra_cost, *lowres_costs, *inv_qscale;
    const int *mvs0, *mvs1;
    int *ref0_cost, *ref1_cost;   // 32-bit shadows of the references' i_propagate_cost
    int bipred_weight;
    float fps_factor;
    int b_bidir;
};
int launch_propagate(cudaStream_t st, const LaGeom &g, const PropagateJob &job);

// One step of a whole mb-tree walk executed by a single cluster kernel (no launch per step).
struct TreeStep {
    int op;                       // 0: zero an accumulator (n ints), 1: propagate, 2: finish
    int n;                        // op 0: number of ints to clear
    int sync;                     // cluster barrier after this step (0 when the next step is independent of it)
    int *zero;                    // op 0
    PropagateJob prop;            // op 1
    // op 2 (pointers only; the LUT comes with the launch)
    const int *fin_propagate; const uint16_t *fin_intra, *fin_invq; const float *fin_qp_aq; float *fin_qp;
    int fin_fps_factor; float fin_weightdelta, fin_strength;
};
int launch_tree_chain(cudaStream_t st, const LaGeom &g, const TreeStep *steps_dev, int nsteps, const float *log2_lut);

struct TreeFinishJob {
    const int *propagate; const uint16_t *intra_cost, *inv_qscale;
    const float *qp_offset_aq; float *qp_offset;
    int fps_factor; float weightdelta, strength;
    const float *log2_lut;
};
int launch_tree_finish(cudaStream_t st, const LaGeom &g, const TreeFinishJob &job);

} // namespace xv
