// Lockstep warp shim: lets g++ compile a warp program written against the xv_* helper set
// (x264vfw_b200/csrc/hpel_kernel.cuh) and run it on the CPU, 32 OS threads per warp, shuffles as
// barrier-separated exchanges.  TEST INFRASTRUCTURE: it exists so that the CPU suite can run the very
// source of a CUDA kernel against the oracle before it goes to the GPU box; nothing in the product
// includes it, and it is far too slow to be a fallback (a 64x48 frame takes milliseconds).
//
// Semantics mirrored from the PTX ISA / CUDA headers:
//   prmt (default mode), shfl.sync up/down/idx with full mask and width 32, dp4a.u32.s32,
//   dp2a.lo.s32.s32, cvt.pack.sat.u8.s32.b32 (operand order as in CUTLASS' NumericArrayConverter<uint8_t,int,4>),
//   __viaddmin_s16x2_relu (host emulation in crt/device_functions.hpp).
#pragma once
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>

#define XV_DEVICE static inline
#define XV_SHARED               /* each simulated lane only touches its own column of a "shared" array */

namespace xv {

using std::max;
using std::min;

struct WarpSim {
    pthread_barrier_t bar;
    uint32_t slot[32];
};
static thread_local WarpSim *t_warp = nullptr;
static thread_local int t_lane = 0;

static inline uint32_t sim_exchange(uint32_t v, int src)
{
    WarpSim *w = t_warp;
    w->slot[t_lane] = v;
    pthread_barrier_wait(&w->bar);
    const uint32_t r = w->slot[src & 31];
    pthread_barrier_wait(&w->bar);
    return r;
}

static inline uint32_t xv_shfl_up1(uint32_t v) { return sim_exchange(v, t_lane >= 1 ? t_lane - 1 : t_lane); }
static inline uint32_t xv_shfl_down1(uint32_t v) { return sim_exchange(v, t_lane <= 30 ? t_lane + 1 : t_lane); }
static inline uint32_t xv_shfl_idx(uint32_t v, int src) { return sim_exchange(v, src); }

static inline uint32_t xv_ld_u32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint32_t xv_ld_u8(const uint8_t *p) { return *p; }
static inline void xv_st_u32(uint8_t *p, uint32_t v)
{
    if ((uintptr_t)p & 3) abort();                      // the device store needs 4-byte alignment
    memcpy(p, &v, 4);
}
static inline void xv_ld_u64(const uint8_t *p, uint32_t &x, uint32_t &y)
{
    if ((uintptr_t)p & 7) abort();                      // the device load needs 8-byte alignment
    memcpy(&x, p, 4); memcpy(&y, p + 4, 4);
}
static inline void xv_st_u64(uint8_t *p, uint32_t x, uint32_t y)
{
    if ((uintptr_t)p & 7) abort();
    memcpy(p, &x, 4); memcpy(p + 4, &y, 4);
}
// cp.async: modelled as an immediate copy (the real one lands no later than the matching wait_group)
typedef uint8_t *xv_saddr;
static inline xv_saddr xv_saddr_of(void *smem) { return (uint8_t *)smem; }
static inline void xv_cp_async8(xv_saddr smem, const void *gmem)
{
    if (((uintptr_t)smem | (uintptr_t)gmem) & 7) abort();
    memcpy(smem, gmem, 8);
}
static inline void xv_cp_async_commit() {}
template <int N> static inline void xv_cp_async_wait() {}
static inline void xv_lds_u64(xv_saddr smem, uint32_t &x, uint32_t &y) { memcpy(&x, smem, 4); memcpy(&y, smem + 4, 4); }
static inline void xv_sts_u64(xv_saddr smem, uint32_t x, uint32_t y) { memcpy(smem, &x, 4); memcpy(smem + 4, &y, 4); }
static inline uint32_t xv_opaque_u32(uint32_t v) { return v; }
static inline xv_saddr xv_opaque_saddr(xv_saddr v) { return v; }
static inline uint8_t *xv_opaque(uint8_t *p) { return p; }
static inline const uint8_t *xv_opaque(const uint8_t *p) { return p; }

static inline uint32_t xv_prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    const uint64_t pool = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        const uint32_t s = (sel >> (4 * i)) & 0xF;
        uint32_t byte = (uint32_t)(pool >> (8 * (s & 7))) & 0xFF;
        if (s & 8) byte = (byte & 0x80) ? 0xFF : 0x00;   // msb replication
        r |= byte << (8 * i);
    }
    return r;
}

static inline int xv_dp4a_us(uint32_t a, uint32_t b, int c)
{
    int64_t d = c;
    for (int i = 0; i < 4; i++) d += (int64_t)((a >> (8 * i)) & 0xFF) * (int8_t)((b >> (8 * i)) & 0xFF);
    return (int)(uint32_t)(uint64_t)d;
}

static inline int xv_dp2a_lo(uint32_t a, uint32_t b, int c)
{
    int64_t d = c;
    d += (int64_t)(int16_t)(a & 0xFFFF) * (int8_t)(b & 0xFF);
    d += (int64_t)(int16_t)(a >> 16) * (int8_t)((b >> 8) & 0xFF);
    return (int)(uint32_t)(uint64_t)d;
}

static inline uint32_t sat_u8(int v) { return v < 0 ? 0u : v > 255 ? 255u : (uint32_t)v; }
static inline uint32_t xv_pack_sat_u8(int v0, int v1, int v2, int v3)
{
    return sat_u8(v0) | (sat_u8(v1) << 8) | (sat_u8(v2) << 16) | (sat_u8(v3) << 24);
}

static inline uint32_t xv_addmin_relu_s16x2(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r = 0;
    for (int i = 0; i < 2; i++) {
        const int16_t s = (int16_t)(uint16_t)(((a >> (16 * i)) & 0xFFFF) + ((b >> (16 * i)) & 0xFFFF));
        const int16_t cc = (int16_t)(uint16_t)((c >> (16 * i)) & 0xFFFF);
        const int m = std::max(0, (int)std::min(s, cc));
        r |= (uint32_t)(m & 0xFFFF) << (16 * i);
    }
    return r;
}

// run fn(lane) on the 32 lanes of one warp in lockstep
template <class F>
static void sim_run_warp(F fn)
{
    WarpSim w;
    pthread_barrier_init(&w.bar, nullptr, 32);
    struct Arg { WarpSim *w; int lane; F *fn; } args[32];
    pthread_t th[32];
    for (int l = 0; l < 32; l++) {
        args[l] = {&w, l, &fn};
        pthread_create(&th[l], nullptr, [](void *p) -> void * {
            Arg *a = (Arg *)p;
            t_warp = a->w; t_lane = a->lane;
            (*a->fn)(a->lane);
            return nullptr;
        }, &args[l]);
    }
    for (int l = 0; l < 32; l++) pthread_join(th[l], nullptr);
    pthread_barrier_destroy(&w.bar);
}

} // namespace xv
