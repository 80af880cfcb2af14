// Attribute interpolation forward / backward (replaces dr.interpolate; reference call sites
// mesh_sfs_optim.py:143,214,269, train_mlp.py:184 with A in {4,6,7,10,30}).
#include "common.cuh"

namespace fmhr {

// One thread per OUTPUT float4: a row of A floats is not 16-byte aligned for A = 7 / 6 / 30, but the output plane as a
// whole is, so thread q produces elements 4q .. 4q+3 of the flat [P*A] array (they straddle at most two pixels for
// A >= 4) with one coalesced 128-bit store.  The kernel is a stream: 92 % of the pixels of the HAM workloads are
// empty, so the common case is "read the id word of one or two rast texels, store a zero float4"; A is a template
// parameter for the widths the reference uses (the divisions become multiply-shift), 0 = run-time width.
template <int AT>
__global__ void __launch_bounds__(256) interpolate_fwd_kernel(const float* __restrict__ attr,
                                                              const float4* __restrict__ rast,
                                                              const int32_t* __restrict__ tri, int NA, int V, int T,
                                                              unsigned hw, int A_rt, size_t nquads, size_t nelem,
                                                              float* __restrict__ out) {
    const int A = AT > 0 ? AT : A_rt;
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nquads) return;
    const size_t e0 = 4 * q;
    // 32-bit index arithmetic whenever the plane allows it (the division by a templated width is then a multiply-shift)
    size_t pix0;
    unsigned k;
    if (nelem <= 0xffffffffull) {
        const unsigned e32 = (unsigned)e0, p32 = e32 / (unsigned)A;
        pix0 = p32;
        k = e32 - p32 * (unsigned)A;
    } else {
        pix0 = e0 / (unsigned)A;
        k = (unsigned)(e0 - pix0 * (unsigned)A);
    }
    const bool full = e0 + 3 < nelem;
    // Fast path, the common case by far: a full quad of a plane with A >= 4 touches at most two pixels; when both are
    // empty (92 % of the frame in the HAM workloads) the thread reads two id words and stores a zero float4.
    if (A >= 4 && full) {
        const bool two = k + 4 > (unsigned)A;
        const int t0 = (int)__ldg(reinterpret_cast<const float*>(rast + pix0) + 3) - 1;
        const int t1 = two ? (int)__ldg(reinterpret_cast<const float*>(rast + pix0 + 1) + 3) - 1 : -1;
        if ((t0 < 0 || t0 >= T) && (t1 < 0 || t1 >= T)) {
            reinterpret_cast<float4*>(out)[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            return;
        }
    }
    size_t pix = pix0;
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    float u = 0.f, v = 0.f;
    const float *A0 = nullptr, *A1 = nullptr, *A2 = nullptr;
    bool have = false, covered = false;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (e0 + j < nelem) {
            if (!have) {  // first element of a pixel handled by this thread: fetch its texel
                have = true;
                const float idw = __ldg(reinterpret_cast<const float*>(rast + pix) + 3);
                const int t = (int)idw - 1;
                covered = t >= 0 && t < T;
                if (covered) {
                    const float4 r = __ldg(rast + pix);
                    u = r.x; v = r.y;
                    const unsigned n = (unsigned)(pix / hw);
                    const float* At = attr + (NA == 1 ? 0 : (size_t)n * V * A);
                    A0 = At + (size_t)__ldg(tri + 3 * t) * A;
                    A1 = At + (size_t)__ldg(tri + 3 * t + 1) * A;
                    A2 = At + (size_t)__ldg(tri + 3 * t + 2) * A;
                }
            }
            if (covered) o[j] = u * __ldg(A0 + k) + v * __ldg(A1 + k) + (1.0f - u - v) * __ldg(A2 + k);
            if (++k == (unsigned)A) { k = 0; pix++; have = false; }
        }
    }
    if (full) {
        reinterpret_cast<float4*>(out)[q] = make_float4(o[0], o[1], o[2], o[3]);
    } else {
        for (int j = 0; j < 4 && e0 + j < nelem; j++) out[e0 + j] = o[j];
    }
}

// One thread per pixel; covered pixels scatter u/v/w-weighted dy to the three vertices and reduce d/du, d/dv.
__global__ void __launch_bounds__(256) interpolate_bwd_kernel(const float* __restrict__ attr,
                                                              const float4* __restrict__ rast,
                                                              const int32_t* __restrict__ tri,
                                                              const float* __restrict__ dy, int NA, int V, int T,
                                                              size_t hw, int A, size_t npix,
                                                              float* __restrict__ grad_attr,
                                                              float4* __restrict__ grad_rast) {
    const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= npix) return;
    const float4 r = __ldg(rast + pix);
    const int t = (int)r.w - 1;
    float du = 0.0f, dv = 0.0f;
    if (t >= 0 && t < T) {
        const int n = (int)(pix / hw);
        const size_t base = (NA == 1 ? 0 : (size_t)n * V * A);
        const float* At = attr + base;
        float* Gt = grad_attr + base;
        const size_t i0 = (size_t)__ldg(tri + 3 * t) * A, i1 = (size_t)__ldg(tri + 3 * t + 1) * A,
                     i2 = (size_t)__ldg(tri + 3 * t + 2) * A;
        const float u = r.x, v = r.y, w = 1.0f - r.x - r.y;
        const float* g = dy + pix * A;
        for (int k = 0; k < A; k++) {
            const float gk = __ldg(g + k);
            if (gk == 0.0f) continue;
            atomicAdd(Gt + i0 + k, u * gk);
            atomicAdd(Gt + i1 + k, v * gk);
            atomicAdd(Gt + i2 + k, w * gk);
            const float a2 = __ldg(At + i2 + k);
            du += gk * (__ldg(At + i0 + k) - a2);
            dv += gk * (__ldg(At + i1 + k) - a2);
        }
    }
    grad_rast[pix] = make_float4(du, dv, 0.0f, 0.0f);
}

}  // namespace fmhr

using namespace fmhr;

extern "C" int fmhr_interpolate_fwd(const float* attr, const float* rast, const int32_t* tri, int N, int NA, int V,
                                    int T, int H, int W, int A, float* out, fmhr_stream_t stream) {
    FMHR_CHECK_ARG(attr && rast && tri && out);
    FMHR_CHECK_ARG(N > 0 && V > 0 && T >= 0 && H > 0 && W > 0 && A > 0);
    FMHR_CHECK_ARG(NA == N || NA == 1);
    const size_t nelem = (size_t)N * H * W * A;
    const size_t nquads = (nelem + 3) / 4;
    FMHR_CHECK_ARG(((uintptr_t)out & 15) == 0 && ((uintptr_t)rast & 15) == 0);
    const unsigned hw = (unsigned)((size_t)H * W);
    const dim3 grid((unsigned)cdiv(nquads, 256));
    cudaStream_t st = (cudaStream_t)stream;
#define FMHR_INTERP(AT)                                                                                              \
    interpolate_fwd_kernel<AT><<<grid, 256, 0, st>>>(attr, (const float4*)rast, tri, NA, V, T, hw, A, nquads, nelem, out)
    switch (A) {  // the widths of the reference's call sites
        case 1: FMHR_INTERP(1); break;
        case 3: FMHR_INTERP(3); break;
        case 4: FMHR_INTERP(4); break;
        case 6: FMHR_INTERP(6); break;
        case 7: FMHR_INTERP(7); break;
        case 10: FMHR_INTERP(10); break;
        case 30: FMHR_INTERP(30); break;
        default: FMHR_INTERP(0); break;
    }
#undef FMHR_INTERP
    FMHR_LAUNCH_CHECK();
    return FMHR_OK;
}

extern "C" int fmhr_interpolate_bwd(const float* attr, const float* rast, const int32_t* tri, const float* dy, int N,
                                    int NA, int V, int T, int H, int W, int A, float* grad_attr, float* grad_rast,
                                    fmhr_stream_t stream) {
    FMHR_CHECK_ARG(attr && rast && tri && dy && grad_attr && grad_rast);
    FMHR_CHECK_ARG(N > 0 && V > 0 && T >= 0 && H > 0 && W > 0 && A > 0);
    FMHR_CHECK_ARG(NA == N || NA == 1);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t npix = (size_t)N * H * W;
    FMHR_CUDA(cudaMemsetAsync(grad_attr, 0, (size_t)NA * V * A * sizeof(float), st));
    interpolate_bwd_kernel<<<cdiv(npix, 256), 256, 0, st>>>(attr, (const float4*)rast, tri, dy, NA, V, T,
                                                            (size_t)H * W, A, npix, grad_attr, (float4*)grad_rast);
    FMHR_LAUNCH_CHECK();
    return FMHR_OK;
}
