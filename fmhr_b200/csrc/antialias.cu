// Silhouette antialiasing forward / backward (replaces dr.antialias; reference call sites
// mesh_sfs_optim.py:146-147,217-219,274,287).  The per-mesh topology (`opp`, from fmhr_mesh_topology_build)
// replaces nvdiffrast's per-call edge hash; the backward re-derives the pair analysis instead of storing a
// work queue, so the op needs no scratch memory.
#include "aa_rule.cuh"

namespace fmhr {

constexpr int kAARow = 128;  // threads per row segment
template <bool BWD>
__global__ void __launch_bounds__(kAARow) antialias_kernel(const float* __restrict__ color,
                                                        const float4* __restrict__ rast,
                                                        const float* __restrict__ pos, const int32_t* __restrict__ tri,
                                                        const int32_t* __restrict__ opp,
                                                        const float* __restrict__ dy, int H, int W, int C, int V, int T,
                                                        size_t npix, float* __restrict__ out,
                                                        float* __restrict__ grad_pos) {
    // grid = (row segments, rows, views): no per-thread index divisions in a kernel that is otherwise a stream
    const int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y, n = blockIdx.z;
    if (px >= W) return;
    const size_t pix0 = ((size_t)n * H + py) * W + px;
    const float4 r0 = __ldg(rast + pix0);
    const float* P = pos + (size_t)n * V * 4;
#pragma unroll
    for (int d = 0; d < 2; d++) {
        if (d == 0 ? (px + 1 >= W) : (py + 1 >= H)) continue;
        const size_t pix1 = pix0 + (d ? W : 1);
        const float4 r1 = __ldg(rast + pix1);
        if (r0.w == r1.w) continue;
        AAPair pr;
        const AAProjClip proj{P, 0.5f * (float)W, 0.5f * (float)H};
        if (!aa_analyse((int)r0.w - 1, r0.z, (int)r1.w - 1, r1.z, px, py, d, proj, tri, opp, V, T, H, W, pr)) continue;
        const float* c0 = color + pix0 * C;
        const float* c1 = color + pix1 * C;
        const size_t recv = (pr.alpha > 0.0f) ? pix0 : pix1;
        if (!BWD) {
            float* o = out + recv * C;
            for (int c = 0; c < C; c++) atomicAdd(o + c, pr.alpha * (__ldg(c1 + c) - __ldg(c0 + c)));
        } else {
            const float* g = dy + recv * C;
            float dd = 0.0f;
            for (int c = 0; c < C; c++) {
                const float gc = __ldg(g + c);
                if (gc == 0.0f) continue;
                dd += gc * (__ldg(c1 + c) - __ldg(c0 + c));
                atomicAdd(out + pix0 * C + c, -pr.alpha * gc);
                atomicAdd(out + pix1 * C + c, pr.alpha * gc);
            }
            if (dd == 0.0f || pr.clamped || grad_pos == nullptr) continue;
            float4 g1, g2;
            aa_pos_grad(pr, px, py, d, ldg4(P + 4 * (size_t)pr.i1), ldg4(P + 4 * (size_t)pr.i2), H, W, dd, g1, g2);
            float4* G = reinterpret_cast<float4*>(grad_pos + (size_t)n * V * 4);
            atomicAdd(G + pr.i1, g1);
            atomicAdd(G + pr.i2, g2);
        }
    }
}

}  // namespace fmhr

using namespace fmhr;

extern "C" int fmhr_antialias_fwd(const float* color, const float* rast, const float* pos, const int32_t* tri,
                                  const int32_t* opp, int N, int H, int W, int C, int V, int T, float* out,
                                  fmhr_stream_t stream) {
    FMHR_CHECK_ARG(color && rast && pos && tri && opp && out);
    FMHR_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && V > 0 && T >= 0);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t npix = (size_t)N * H * W;
    FMHR_CUDA(cudaMemcpyAsync(out, color, npix * C * sizeof(float), cudaMemcpyDeviceToDevice, st));
    FMHR_CHECK_ARG(H <= 65535 && N <= 65535);
    antialias_kernel<false><<<dim3(cdiv(W, kAARow), H, N), kAARow, 0, st>>>(color, (const float4*)rast, pos, tri, opp, nullptr, H, W,
                                                             C, V, T, npix, out, nullptr);
    FMHR_LAUNCH_CHECK();
    return FMHR_OK;
}

extern "C" int fmhr_antialias_bwd(const float* color, const float* rast, const float* pos, const int32_t* tri,
                                  const int32_t* opp, const float* dy, int N, int H, int W, int C, int V, int T,
                                  float* grad_color, float* grad_pos, fmhr_stream_t stream) {
    FMHR_CHECK_ARG(color && rast && pos && tri && opp && dy && grad_color);
    FMHR_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && V > 0 && T >= 0);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t npix = (size_t)N * H * W;
    FMHR_CUDA(cudaMemcpyAsync(grad_color, dy, npix * C * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (grad_pos) FMHR_CUDA(cudaMemsetAsync(grad_pos, 0, (size_t)N * V * 4 * sizeof(float), st));
    FMHR_CHECK_ARG(H <= 65535 && N <= 65535);
    antialias_kernel<true><<<dim3(cdiv(W, kAARow), H, N), kAARow, 0, st>>>(color, (const float4*)rast, pos, tri, opp, dy, H, W, C, V,
                                                            T, npix, grad_color, grad_pos);
    FMHR_LAUNCH_CHECK();
    return FMHR_OK;
}
