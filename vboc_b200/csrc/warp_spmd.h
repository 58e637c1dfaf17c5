// warp_spmd.h -- the warp-cooperative programming style of the OCP kernels.
//
// One warp owns one OCP.  The solver code (ocp_warp.h) is written as a sequence of LANE REGIONS:
//
//     FOR_LANES   ... code executed by each of the 32 lanes, `lane` in scope ...   END_LANES
//
// Rules of the style:
//   * inside one region a lane never reads shared/global data that another lane writes in the
//     same region; END_LANES is a __syncwarp(), which also orders the warp's memory accesses;
//   * code outside the regions is WARP-UNIFORM: every lane computes the same values from
//     shared memory, uniform global loads or reduction results, and writes nothing;
//   * a per-lane value that must survive a region boundary is declared with LV(type, name) and
//     accessed as L(name); WARP_MIN/MAX/SUM(name) reduce it over the warp (shuffle butterflies);
//   * uniform code whose loads are overwritten by the next region ends with UNIFORM_SYNC().
//
// With nvcc this is ordinary SIMT code.  With a host compiler (VBOC_EMU, tools/emu) the same
// source runs the 32 lanes of a region as a loop -- the lane program is then debuggable on a
// machine without a GPU.  The emulation is a development/test harness only; nothing in the
// product loads it.
#pragma once
#include "fast_math.h"

#if defined(__CUDACC__) && !defined(VBOC_EMU)

#define VB_HD __host__ __device__ __forceinline__
#define VB_DEV __device__ __forceinline__
#define VB_DEV_NOINLINE __device__ __noinline__  // one copy of a large callee shared by several call sites
#define FOR_LANES {                                   \
    const int lane = (int)(threadIdx.x & 31u);        \
    (void)lane;
#define END_LANES }                                   \
    __syncwarp();
#define LV(type, name) type name
#define L(name) name
#define UNIFORM_SYNC() __syncwarp()
// VB_FAST_MATH: hardware seed + two Newton steps instead of the IEEE library sequences (fast_math.h)
#ifndef VB_FAST_MATH
#define VB_FAST_MATH 2  // 1: reciprocals / rsqrt, 2: also the ratio test (measured +1.5 % and +4 %: shorter code)
#endif
#if VB_FAST_MATH
#define VB_RSQRT(x) vb_rsqrt_pos(x)
#define VB_RCP(x) vb_rcp_pos(x)
#else
#define VB_RSQRT(x) rsqrt(x)
#define VB_RCP(x) (1.0 / (x))
#endif
// a / b with b > 0 (step to the boundary in the ratio test)
#if VB_FAST_MATH >= 2
#define VB_RATIO(a, b) ((a) * vb_rcp_pos(b))
#else
#define VB_RATIO(a, b) ((a) / (b))
#endif
// software prefetch of the next lane-strided iteration (the flat passes are latency bound)
__device__ __forceinline__ void vb_prefetch(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

__device__ __forceinline__ void vb_prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#define VB_PREFETCH_L2(p) vb_prefetch_l2(p)
#define VB_PREFETCH(p) vb_prefetch(p)

// out of line on purpose: ~20 call sites share one copy (instruction-cache footprint)
__device__ __noinline__ double vb_warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __noinline__ double vb_warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __noinline__ double vb_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int vb_warp_or(int v) { return __reduce_or_sync(0xffffffffu, (unsigned)v) != 0; }
#define WARP_MAX(name) vb_warp_max(name)
#define WARP_MIN(name) vb_warp_min(name)
#define WARP_SUM(name) vb_warp_sum(name)
#define WARP_ANY(name) vb_warp_or(name)

// ---- stage-record ring: TMA bulk copies global -> shared, completion on an mbarrier per slot ----
// One lane issues `cp.async.bulk` for a whole contiguous stage record; all lanes wait on the slot's
// mbarrier phase.  `ph` is the warp-uniform bitmask of the slots' current phase parities.
struct VbRing {
    unsigned long long *bar;  // shared: one mbarrier per slot
    unsigned bar0;            // its shared-window address (computed once)
    unsigned ph;
};
__device__ __forceinline__ unsigned vb_smem_addr(const void *p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void vb_ring_init(VbRing &r, unsigned long long *bars, int depth) {
    r.bar = bars, r.ph = 0u, r.bar0 = vb_smem_addr(bars);
    if ((threadIdx.x & 31u) == 0) {
        for (int i = 0; i < depth; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(vb_smem_addr(bars + i)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
}
__device__ __forceinline__ void vb_ring_fetch(VbRing &r, int slot, void *dst, const void *src, unsigned bytes,
                                              bool leader) {
    if (leader) {
        unsigned b = r.bar0 + 8u * (unsigned)slot;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                vb_smem_addr(dst)),
            "l"(src), "r"(bytes), "r"(b)
            : "memory");
    }
}
__device__ __forceinline__ void vb_ring_wait(VbRing &r, int slot) {
    unsigned b = r.bar0 + 8u * (unsigned)slot, parity = (r.ph >> slot) & 1u;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "VB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra VB_DONE_%=;\n"
        "bra VB_WAIT_%=;\n"
        "VB_DONE_%=:\n"
        "}\n" ::"r"(b),
        "r"(parity)
        : "memory");
    r.ph ^= 1u << slot;
}
// generic-proxy global writes -> later async-proxy (TMA) reads of the same memory
#define PROXY_FENCE()                                      \
    do {                                                   \
        asm volatile("fence.proxy.async;" ::: "memory");   \
        __syncwarp();                                      \
    } while (0)
#define RING_INIT(r, bars, depth) vb_ring_init(r, bars, depth)
#define RING_FETCH(r, slot, dst, src, bytes) vb_ring_fetch(r, slot, dst, src, bytes, (threadIdx.x & 31u) == 0)
#define RING_WAIT(r, slot) vb_ring_wait(r, slot)

#else  // host emulation

#include <cmath>
#define VB_HD inline
#define VB_DEV inline
#define VB_DEV_NOINLINE inline
#define FOR_LANES for (int lane = 0; lane < 32; ++lane) {
#define END_LANES }
#define LV(type, name) type name[32]
#define L(name) name[lane]
#define UNIFORM_SYNC() ((void)0)
#define VB_RSQRT(x) (1.0 / std::sqrt(x))
#define VB_RCP(x) (1.0 / (x))
#define VB_RATIO(a, b) ((a) / (b))
#define VB_PREFETCH(p) ((void)(p))
#define VB_PREFETCH_L2(p) ((void)(p))

// the same butterfly order as the shuffle versions, so sums round identically
inline double vb_emu_max(const double *a) {
    double t[32], u[32];
    for (int i = 0; i < 32; ++i) t[i] = a[i];
    for (int o = 16; o > 0; o >>= 1) {
        for (int i = 0; i < 32; ++i) u[i] = std::fmax(t[i], t[i ^ o]);
        for (int i = 0; i < 32; ++i) t[i] = u[i];
    }
    return t[0];
}
inline double vb_emu_min(const double *a) {
    double t[32], u[32];
    for (int i = 0; i < 32; ++i) t[i] = a[i];
    for (int o = 16; o > 0; o >>= 1) {
        for (int i = 0; i < 32; ++i) u[i] = std::fmin(t[i], t[i ^ o]);
        for (int i = 0; i < 32; ++i) t[i] = u[i];
    }
    return t[0];
}
inline double vb_emu_sum(const double *a) {
    double t[32], u[32];
    for (int i = 0; i < 32; ++i) t[i] = a[i];
    for (int o = 16; o > 0; o >>= 1) {
        for (int i = 0; i < 32; ++i) u[i] = t[i] + t[i ^ o];
        for (int i = 0; i < 32; ++i) t[i] = u[i];
    }
    return t[0];
}
inline int vb_emu_any(const int *a) {
    int r = 0;
    for (int i = 0; i < 32; ++i) r |= a[i] != 0;
    return r;
}
#define WARP_MAX(name) vb_emu_max(name)
#define WARP_MIN(name) vb_emu_min(name)
#define WARP_SUM(name) vb_emu_sum(name)
#define WARP_ANY(name) vb_emu_any(name)

#include <cstring>
struct VbRing {
    unsigned long long *bar;
    unsigned ph;
};
#define PROXY_FENCE() ((void)0)
#define RING_INIT(r, bars, depth) ((void)(r).ph)
#define RING_FETCH(r, slot, dst, src, bytes) std::memcpy(dst, src, bytes)
#define RING_WAIT(r, slot) ((void)0)

#endif
