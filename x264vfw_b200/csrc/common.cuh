// Shared device/host helpers for the x264vfw B200 front end (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

namespace xv {

// ---- error plumbing (thread-local last error string behind the C ABI) ----------------
void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launch_count;

#define XV_CUDA_OK(expr)                                                          \
    do {                                                                          \
        cudaError_t e__ = (expr);                                                 \
        if (e__ != cudaSuccess) {                                                 \
            xv::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                          __FILE__, __LINE__);                                    \
            return -1;                                                            \
        }                                                                         \
    } while (0)

#define XV_LAUNCH_CHECK()                                                         \
    do {                                                                          \
        xv::g_launch_count.fetch_add(1, std::memory_order_relaxed);               \
        cudaError_t e__ = cudaGetLastError();                                     \
        if (e__ != cudaSuccess) {                                                 \
            xv::set_error("kernel launch failed: %s (%s:%d)",                     \
                          cudaGetErrorString(e__), __FILE__, __LINE__);           \
            return -1;                                                            \
        }                                                                         \
    } while (0)

struct Ctx {
    int device;
    cudaStream_t stream;
    // staging for the host-pointer paths (grown on demand)
    uint8_t *h_pinned = nullptr; size_t h_pinned_bytes = 0;
    uint8_t *d_src = nullptr;    size_t d_src_bytes = 0;
    uint8_t *d_dst = nullptr;    size_t d_dst_bytes = 0;
};

// ---- streaming loads/stores ------------------------------------------------------------
// Inputs are read exactly once: bypass L1 allocation.  Outputs use default policy because
// the next stage (lowres / AQ / lookahead) consumes them out of L2.
__device__ __forceinline__ uint4 ldg_stream128(const void *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ldg_stream64(const void *p)
{
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];"
                 : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldg_stream32(const void *p)
{
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

// rounding byte-wise average (a+b+1)>>1 on 4 packed bytes
// = __vavgu4, which the compiler expands to five integer operations (xor, shift, and, or, sub);
// folding the mask into the xor with one LOP3 leaves four -- these kernels are ALU-pipe bound
__device__ __forceinline__ uint32_t avg4(uint32_t a, uint32_t b)
{
    uint32_t t;
    asm("lop3.b32 %0, %1, %2, 0xfefefefe, 0x28;" : "=r"(t) : "r"(a), "r"(b));      // (a ^ b) & 0xfefefefe
    return (a | b) - (t >> 1);
}

} // namespace xv
