// Attribute interpolation forward / backward (replaces dr.interpolate; reference call sites
// mesh_sfs_optim.py:143,214,269, train_mlp.py:184 with A in {4,6,7,10,30}).
#include "common.cuh"

namespace fmhr {

// One thread per OUTPUT ELEMENT (pixel, channel): stores are perfectly coalesced for any A (A=7 rows are 28 B,
// not vectorisable per pixel), the 16-byte rast texel is a warp-broadcast L1 hit, attribute gathers are
// contiguous in the channel index.
__global__ void __launch_bounds__(256) interpolate_fwd_kernel(const float* __restrict__ attr,
                                                              const float4* __restrict__ rast,
                                                              const int32_t* __restrict__ tri, int NA, int V, int T,
                                                              size_t hw, int A, size_t nelem, float* __restrict__ out) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nelem) return;
    const size_t pix = e / A;
    const int k = (int)(e - pix * A);
    const float4 r = __ldg(rast + pix);
    const int t = (int)r.w - 1;
    float o = 0.0f;
    if (t >= 0 && t < T) {
        const int n = (int)(pix / hw);
        const float* At = attr + (NA == 1 ? 0 : (size_t)n * V * A);
        const float a0 = __ldg(At + (size_t)__ldg(tri + 3 * t) * A + k);
        const float a1 = __ldg(At + (size_t)__ldg(tri + 3 * t + 1) * A + k);
        const float a2 = __ldg(At + (size_t)__ldg(tri + 3 * t + 2) * A + k);
        o = r.x * a0 + r.y * a1 + (1.0f - r.x - r.y) * a2;
    }
    out[e] = o;
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
    interpolate_fwd_kernel<<<cdiv(nelem, 256), 256, 0, (cudaStream_t)stream>>>(attr, (const float4*)rast, tri, NA, V,
                                                                               T, (size_t)H * W, A, nelem, out);
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
