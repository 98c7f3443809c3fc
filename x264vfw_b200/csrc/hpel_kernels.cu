// Launcher of the half-pel reference plane kernel (hpel_kernel.cuh holds the warp program and its
// description).  sm_100a device versions of the helpers the program uses.
#include "common.cuh"
#include "csp_kernels.h"

#define XV_DEVICE __device__ __forceinline__
#define XV_SHARED __shared__

namespace xv {

__device__ __forceinline__ uint32_t xv_ld_u32(const uint8_t *p) { return __ldg((const uint32_t *)p); }
__device__ __forceinline__ void xv_ld_u64(const uint8_t *p, uint32_t &x, uint32_t &y) { const uint2 v = __ldg((const uint2 *)p); x = v.x; y = v.y; }
__device__ __forceinline__ uint32_t xv_ld_u8(const uint8_t *p) { return __ldg(p); }
__device__ __forceinline__ void xv_st_u32(uint8_t *p, uint32_t v) { asm volatile("st.global.b32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void xv_st_u64(uint8_t *p, uint32_t x, uint32_t y) { asm volatile("st.global.v2.b32 [%0], {%1, %2};" :: "l"(p), "r"(x), "r"(y) : "memory"); }
// shared-memory addresses are carried as 32-bit shared-window offsets
typedef uint32_t xv_saddr;
__device__ __forceinline__ xv_saddr xv_saddr_of(const void *smem) { return (uint32_t)__cvta_generic_to_shared(smem); }
// 8 bytes global -> shared without a destination register; groups complete in order
__device__ __forceinline__ void xv_cp_async8(xv_saddr smem, const void *gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(smem), "l"(gmem) : "memory");
}
__device__ __forceinline__ void xv_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void xv_cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void xv_lds_u64(xv_saddr smem, uint32_t &x, uint32_t &y)
{
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(smem) : "memory");
}
__device__ __forceinline__ void xv_sts_u64(xv_saddr smem, uint32_t x, uint32_t y)
{
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" :: "r"(smem), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ uint32_t xv_opaque_u32(uint32_t v) { asm volatile("" : "+r"(v)); return v; }
__device__ __forceinline__ xv_saddr xv_opaque_saddr(xv_saddr v) { asm volatile("" : "+r"(v)); return v; }
// keeps the compiler from folding the frame's base back into every address computation
__device__ __forceinline__ uint8_t *xv_opaque(uint8_t *p) { asm volatile("" : "+l"(p)); return p; }
__device__ __forceinline__ const uint8_t *xv_opaque(const uint8_t *p) { asm volatile("" : "+l"(p)); return p; }
__device__ __forceinline__ uint32_t xv_prmt(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }
__device__ __forceinline__ uint32_t xv_shfl_up1(uint32_t v) { return __shfl_up_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ uint32_t xv_shfl_down1(uint32_t v) { return __shfl_down_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ uint32_t xv_shfl_idx(uint32_t v, int src) { return __shfl_sync(0xffffffffu, v, src); }
// per 16-bit lane: max(min(a + b, c), 0)  (DPX, VIADDMNMX-class on sm_90+)
__device__ __forceinline__ uint32_t xv_addmin_relu_s16x2(uint32_t a, uint32_t b, uint32_t c) { return __viaddmin_s16x2_relu(a, b, c); }
// c + sum of four (unsigned byte of a) * (signed byte of b)
__device__ __forceinline__ int xv_dp4a_us(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// c + (signed low half of a) * (signed byte 0 of b) + (signed high half of a) * (signed byte 1 of b)
__device__ __forceinline__ int xv_dp2a_lo(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// bytes (sat(v0), sat(v1), sat(v2), sat(v3)), v0 in the low byte; sat = clamp to [0, 255]
__device__ __forceinline__ uint32_t xv_pack_sat_u8(int v0, int v1, int v2, int v3)
{
    uint32_t t, d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, 0;" : "=r"(t) : "r"(v3), "r"(v2));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(v1), "r"(v0), "r"(t));
    return d;
}

} // namespace xv

#include "hpel_kernel.cuh"

namespace xv {

// one warp per block: tiles at the frame's edges do more work (border), a block of several warps would
// hold its slot until the slowest one is done.  Register budget: asking for 25 blocks (cap 80) lets ptxas settle
// on 72 registers without a spill -- 28 warps per SM all the same; asking for 28 (cap 72) spills the trip counter.
// The byte-gather instantiation (unaligned planes) needs more registers and is bound by its loads anyway.
template <bool ALIGNED>
__global__ void __launch_bounds__(32, ALIGNED ? 25 : 20)
hpel_kernel(HpelJob job)
{
    hpel_unit_any<ALIGNED>(job, blockIdx.x, blockIdx.y, threadIdx.x);
}

int launch_hpel(cudaStream_t st, HpelJob &job, int n_frames)
{
    if (job.w <= 0 || job.h <= 0 || n_frames <= 0) return 0;
    const long long units = hpel_plan(job, n_frames);
    dim3 grid((unsigned)units, (unsigned)n_frames);
    if (job.aligned) hpel_kernel<true><<<grid, 32, 0, st>>>(job);
    else hpel_kernel<false><<<grid, 32, 0, st>>>(job);
    XV_LAUNCH_CHECK();
    return 0;
}

} // namespace xv
