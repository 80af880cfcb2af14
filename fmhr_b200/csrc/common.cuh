// Shared device/host helpers for the fmhr_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/fmhr_b200.h"

namespace fmhr {

void set_error(const char* fmt, ...);

#define FMHR_CHECK_ARG(cond)                                                              \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            fmhr::set_error("%s: invalid argument: %s", __func__, #cond);                 \
            return FMHR_EINVAL;                                                           \
        }                                                                                 \
    } while (0)

#define FMHR_CUDA(call)                                                                   \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            fmhr::set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(e_)); \
            return FMHR_ECUDA;                                                            \
        }                                                                                 \
    } while (0)

#define FMHR_LAUNCH_CHECK() FMHR_CUDA(cudaGetLastError())

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------------
// Exactly-rounded fp32 arithmetic: never contracted into FMA, so decisions (coverage depth keys,
// antialias edge selection) are bit-identical to the CPU oracle built with -ffp-contract=off.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float xm(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xa(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xs(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xd(float a, float b) { return __fdiv_rn(a, b); }

constexpr unsigned long long ZB_EMPTY = ~0ull;
constexpr int kTile = 16;  // screen tile edge of the coverage kernels' dirty-tile bitmaps (ham.cu, raster.cu)

__device__ __forceinline__ uint32_t depth_key(float zw) {
    uint32_t b = __float_as_uint(zw);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float depth_from_key(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(b);
}
__device__ __forceinline__ float clamp01x(float x) { return (x > 0.0f) ? ((x < 1.0f) ? x : 1.0f) : 0.0f; }
__device__ __forceinline__ float clampzx(float x) { return (x < 1.0f) ? ((x > -1.0f) ? x : -1.0f) : 1.0f; }

struct Bary {
    float u, v, zw;
};

// Perspective-correct barycentrics and depth of pixel (px,py); same operation order as the oracle's bary_at().
__device__ __forceinline__ Bary bary_at(const float4 p0, const float4 p1, const float4 p2, int px, int py, float invW,
                                        float invH) {
    float fx = xs(xm((float)(2 * px + 1), invW), 1.0f);
    float fy = xs(xm((float)(2 * py + 1), invH), 1.0f);
    float q0x = xs(p0.x, xm(fx, p0.w)), q0y = xs(p0.y, xm(fy, p0.w));
    float q1x = xs(p1.x, xm(fx, p1.w)), q1y = xs(p1.y, xm(fy, p1.w));
    float q2x = xs(p2.x, xm(fx, p2.w)), q2y = xs(p2.y, xm(fy, p2.w));
    float a0 = xs(xm(q1x, q2y), xm(q1y, q2x));
    float a1 = xs(xm(q2x, q0y), xm(q2y, q0x));
    float a2 = xs(xm(q0x, q1y), xm(q0y, q1x));
    float at = xa(xa(a0, a1), a2);
    float iw = xd(1.0f, at);
    Bary b;
    b.u = clamp01x(xm(a0, iw));
    b.v = clamp01x(xm(a1, iw));
    float zn = xa(xa(xm(p0.z, a0), xm(p1.z, a1)), xm(p2.z, a2));
    float wd = xa(xa(xm(p0.w, a0), xm(p1.w, a1)), xm(p2.w, a2));
    b.zw = xa(clampzx(xd(zn, wd)), 0.0f);
    return b;
}

// clip -> 24.8 fixed-point window coordinates (pixel centre at integer + 0.5 -> +128); rejects vertices outside the clip
// volume / guard band.  Same operation order as the oracle's snap_tri().
constexpr float kGuard = 16384.0f;
constexpr int kSnapRejected = INT_MIN;
__device__ __forceinline__ bool snap_vertex(const float4 p, float hw, float hh, int& X, int& Y) {
    if (!(p.w > 0.0f)) return false;
    if (!(p.z >= -p.w && p.z <= p.w)) return false;
    float sx = xa(xm(xd(p.x, p.w), hw), hw);
    float sy = xa(xm(xd(p.y, p.w), hh), hh);
    if (!(fabsf(sx) <= kGuard) || !(fabsf(sy) <= kGuard)) return false;
    X = __float2int_rn(xm(sx, 256.0f));
    Y = __float2int_rn(xm(sy, 256.0f));
    return true;
}

// ------------------------------------------------------------------------------------------------
// Timeline tracing (diagnostic builds only, -DFMHR_TRACE; tools/trace_timeline.py): every kernel of the fused iteration
// records the %globaltimer of its first block entry (atomicMin) and of its last block exit (atomicMax) into slot `id`, so
// the real schedule of a CUDA-graph replay - overlaps, fill / tail bubbles between kernels - can be read back.  The
// product build compiles the scope to nothing.
// ------------------------------------------------------------------------------------------------
constexpr int kTraceSlots = 64;
#ifdef FMHR_TRACE
static __device__ unsigned long long g_trace[kTraceSlots][2];  // only ham.cu's copy is used / read back
__device__ __forceinline__ unsigned long long trace_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
struct TraceScope {
    int id;
    // slot id: [first block entry, last block exit]; slot id + 32: [first block EXIT, last block ENTRY] - the spread of the
    // blocks' entries (fill: blocks that had to wait for a slot) and of their exits (tail: imbalance of the work split)
    __device__ __forceinline__ explicit TraceScope(int i) : id(i) {
        if (threadIdx.x == 0) {
            const unsigned long long t = trace_now();
            atomicMin(&g_trace[id][0], t);
            if (id < 32) atomicMax(&g_trace[id + 32][1], t);
        }
    }
    __device__ __forceinline__ ~TraceScope() {
        if (threadIdx.x == 0) {  // (thread 0's exit stands for the block's)
            const unsigned long long t = trace_now();
            atomicMax(&g_trace[id][1], t);
            if (id < 32) atomicMin(&g_trace[id + 32][0], t);
        }
    }
};
__device__ __forceinline__ void trace_stamp_min(int id, int k) { atomicMin(&g_trace[id][k], trace_now()); }
__device__ __forceinline__ void trace_stamp_max(int id, int k) { atomicMax(&g_trace[id][k], trace_now()); }
#define FMHR_TRACE_SCOPE(id) fmhr::TraceScope trace_scope_(id)
#define FMHR_TRACE_MIN(id, k) do { if (threadIdx.x == 0) fmhr::trace_stamp_min(id, k); } while (0)
#define FMHR_TRACE_MAX(id, k) do { if (threadIdx.x == 0) fmhr::trace_stamp_max(id, k); } while (0)
#else
#define FMHR_TRACE_SCOPE(id)
#define FMHR_TRACE_MIN(id, k)
#define FMHR_TRACE_MAX(id, k)
#endif

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// warp / block reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace fmhr
