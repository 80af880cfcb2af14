// NCC photo-consistency term of the HAM iteration (BASELINE.json configs[2]: "16 views x 1024x1024 with NCC loss").
//
// models/ncc_utils.py:4-35 (NCC) has no caller in the reference, so the WIRING is this project's and is
// documented in DESIGN.md: Np points are fixed on the mesh surface (face id + barycentrics), every iteration each point is
// projected into a reference view and Nv source views, an axis-aligned (2 half + 1)^2 patch of the view's gray image is
// sampled bilinearly around the projection, NCC(ref patch, source patches, source mask) is evaluated with the reference's
// arithmetic (fmhr_ncc_fwd / fmhr_ncc_bwd, mesh.cu) and  weight * mean(1 - ncc)  is back-propagated through the SOURCE
// sampling positions to the vertices (the reference patch is a constant of the step).
//   fmhr_ncc_sample_fwd : vertices, points, cameras, gray images, masks -> patches [Nv+1,Np,Npx], patch masks
//   fmhr_ncc_sample_bwd : d loss / d patches -> d loss / d vertices  (+=, world space, three atomics per (view, point))
#include "common.cuh"
#include <stdlib.h>

namespace fmhr {

struct NccCam { float m[16]; };
__device__ __forceinline__ NccCam ncc_view_matrix(const float* __restrict__ w2cs, const float* __restrict__ projs, int view) {
    const float* Wm = w2cs + (size_t)view * 16;
    const float* Pm = projs + (size_t)view * 16;
    NccCam M;
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int j = 0; j < 4; j++)
            M.m[4 * r + j] = __fmaf_rn(Wm[4 * r + 3], Pm[12 + j], __fmaf_rn(Wm[4 * r + 2], Pm[8 + j],
                                       __fmaf_rn(Wm[4 * r + 1], Pm[4 + j], __fmul_rn(Wm[4 * r], Pm[j]))));
    return M;
}

struct NccPoint { float3 X; int i0, i1, i2; float b0, b1, b2; };
__device__ __forceinline__ NccPoint ncc_point(const float* __restrict__ vertices, const int32_t* __restrict__ tri,
                                              const int32_t* __restrict__ pt_face, const float* __restrict__ pt_bary, int p) {
    NccPoint q;
    const int f = __ldg(pt_face + p);
    q.i0 = __ldg(tri + 3 * (size_t)f); q.i1 = __ldg(tri + 3 * (size_t)f + 1); q.i2 = __ldg(tri + 3 * (size_t)f + 2);
    q.b0 = __ldg(pt_bary + 2 * (size_t)p); q.b1 = __ldg(pt_bary + 2 * (size_t)p + 1); q.b2 = 1.0f - q.b0 - q.b1;
    const float* a = vertices + 3 * (size_t)q.i0; const float* b = vertices + 3 * (size_t)q.i1; const float* c = vertices + 3 * (size_t)q.i2;
    q.X = make_float3(q.b0 * a[0] + q.b1 * b[0] + q.b2 * c[0], q.b0 * a[1] + q.b1 * b[1] + q.b2 * c[1],
                      q.b0 * a[2] + q.b1 * b[2] + q.b2 * c[2]);
    return q;
}
// continuous sample coordinates (u, v): integer values are pixel centres; w = clip w
__device__ __forceinline__ void ncc_project(const NccCam& M, float3 X, int H, int W, float& u, float& v, float4& clip) {
    clip = make_float4(X.x * M.m[0] + X.y * M.m[4] + X.z * M.m[8] + M.m[12], X.x * M.m[1] + X.y * M.m[5] + X.z * M.m[9] + M.m[13],
                       X.x * M.m[2] + X.y * M.m[6] + X.z * M.m[10] + M.m[14], X.x * M.m[3] + X.y * M.m[7] + X.z * M.m[11] + M.m[15]);
    const float iw = 1.0f / clip.w;
    u = (clip.x * iw * 0.5f + 0.5f) * (float)W - 0.5f;
    v = (clip.y * iw * 0.5f + 0.5f) * (float)H - 0.5f;
}
// bilinear tap set of sample (su, sv): values of the four taps (0 outside the image), weights, validity
struct NccTaps { float i00, i10, i01, i11, fx, fy; bool inside; int xn, yn; };
__device__ __forceinline__ NccTaps ncc_taps(const float* __restrict__ img, int H, int W, float su, float sv) {
    NccTaps t;
    const float x0f = floorf(su), y0f = floorf(sv);
    const int x0 = (int)x0f, y0 = (int)y0f;
    t.fx = su - x0f; t.fy = sv - y0f;
    const bool xa = x0 >= 0 && x0 < W, xb = x0 + 1 >= 0 && x0 + 1 < W, ya = y0 >= 0 && y0 < H, yb = y0 + 1 >= 0 && y0 + 1 < H;
    t.i00 = (xa && ya) ? __ldg(img + (size_t)y0 * W + x0) : 0.0f;
    t.i10 = (xb && ya) ? __ldg(img + (size_t)y0 * W + x0 + 1) : 0.0f;
    t.i01 = (xa && yb) ? __ldg(img + (size_t)(y0 + 1) * W + x0) : 0.0f;
    t.i11 = (xb && yb) ? __ldg(img + (size_t)(y0 + 1) * W + x0 + 1) : 0.0f;
    t.inside = xa && xb && ya && yb;
    t.xn = x0 + (t.fx >= 0.5f ? 1 : 0); t.yn = y0 + (t.fy >= 0.5f ? 1 : 0);  // nearest pixel (mask lookup)
    return t;
}

// one warp per (view slot, point): lanes stride over the patch
__global__ void __launch_bounds__(256) ncc_sample_fwd_kernel(
    const float* __restrict__ vertices, const int32_t* __restrict__ tri, const int32_t* __restrict__ pt_face,
    const float* __restrict__ pt_bary, const float* __restrict__ w2cs, const float* __restrict__ projs,
    const int32_t* __restrict__ view_idx, int Nv1, const float* __restrict__ gray, const float* __restrict__ masks, int Np,
    int H, int W, int half, float* __restrict__ patches, float* __restrict__ patch_mask) {
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= (long long)Nv1 * Np) return;
    const int s = (int)(warp / Np), p = (int)(warp - (long long)s * Np);
    const int view = __ldg(view_idx + s);
    const NccCam M = ncc_view_matrix(w2cs, projs, view);
    const NccPoint q = ncc_point(vertices, tri, pt_face, pt_bary, p);
    float u, v;
    float4 clip;
    ncc_project(M, q.X, H, W, u, v, clip);
    const int side = 2 * half + 1, npx = side * side;
    const float* img = gray + (size_t)view * H * W;
    const float* msk = masks + (size_t)view * H * W;
    const bool front = clip.w > 0.0f;
    for (int k = lane; k < npx; k += 32) {
        const int dy = k / side - half, dx = k - (k / side) * side - half;
        float val = 0.0f, mk = 0.0f;
        if (front) {
            const NccTaps t = ncc_taps(img, H, W, u + (float)dx, v + (float)dy);
            val = (1.0f - t.fy) * ((1.0f - t.fx) * t.i00 + t.fx * t.i10) + t.fy * ((1.0f - t.fx) * t.i01 + t.fx * t.i11);
            if (t.inside && __ldg(msk + (size_t)t.yn * W + t.xn) > 0.5f) mk = 1.0f;
        }
        patches[(size_t)warp * npx + k] = val;
        patch_mask[(size_t)warp * npx + k] = mk;
    }
}

// one warp per (source view slot >= 1, point): d loss / d (u, v) summed over the patch, chain rule to the vertices
__global__ void __launch_bounds__(256) ncc_sample_bwd_kernel(
    const float* __restrict__ vertices, const int32_t* __restrict__ tri, const int32_t* __restrict__ pt_face,
    const float* __restrict__ pt_bary, const float* __restrict__ w2cs, const float* __restrict__ projs,
    const int32_t* __restrict__ view_idx, int Nv1, const float* __restrict__ gray, int Np, int H, int W, int half,
    const float* __restrict__ grad_patches, float* __restrict__ grad_vertices) {
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= (long long)(Nv1 - 1) * Np) return;
    const int s = 1 + (int)(warp / Np), p = (int)(warp - (long long)(s - 1) * Np);
    const int view = __ldg(view_idx + s);
    const NccCam M = ncc_view_matrix(w2cs, projs, view);
    const NccPoint q = ncc_point(vertices, tri, pt_face, pt_bary, p);
    float u, v;
    float4 clip;
    ncc_project(M, q.X, H, W, u, v, clip);
    if (!(clip.w > 0.0f)) return;
    const int side = 2 * half + 1, npx = side * side;
    const float* img = gray + (size_t)view * H * W;
    const float* g = grad_patches + ((size_t)s * Np + p) * npx;
    float gu = 0.0f, gv = 0.0f;
    for (int k = lane; k < npx; k += 32) {
        const float gk = __ldg(g + k);
        if (gk == 0.0f) continue;
        const int dy = k / side - half, dx = k - (k / side) * side - half;
        const NccTaps t = ncc_taps(img, H, W, u + (float)dx, v + (float)dy);
        gu += gk * ((1.0f - t.fy) * (t.i10 - t.i00) + t.fy * (t.i11 - t.i01));
        gv += gk * ((1.0f - t.fx) * (t.i01 - t.i00) + t.fx * (t.i11 - t.i10));
    }
    gu = warp_sum(gu); gv = warp_sum(gv);
    if (lane != 0 || (gu == 0.0f && gv == 0.0f)) return;
    // u = (cx / cw * 0.5 + 0.5) W - 0.5: d u / d clip = W/2 (1/cw, 0, 0, -cx/cw^2), likewise v with H and cy
    const float iw = 1.0f / clip.w, hu = 0.5f * (float)W * gu * iw, hv = 0.5f * (float)H * gv * iw;
    const float gcx = hu, gcy = hv, gcw = -(hu * clip.x + hv * clip.y) * iw;
    const float3 gX = make_float3(M.m[0] * gcx + M.m[1] * gcy + M.m[3] * gcw, M.m[4] * gcx + M.m[5] * gcy + M.m[7] * gcw,
                                  M.m[8] * gcx + M.m[9] * gcy + M.m[11] * gcw);
    const int vi[3] = {q.i0, q.i1, q.i2};
    const float bw[3] = {q.b0, q.b1, q.b2};
#pragma unroll
    for (int c = 0; c < 3; c++) {
        float* o = grad_vertices + 3 * (size_t)vi[c];
        atomicAdd(o, bw[c] * gX.x); atomicAdd(o + 1, bw[c] * gX.y); atomicAdd(o + 2, bw[c] * gX.z);
    }
}

// Fused form of the whole term for patches of at most 128 samples (11 x 11 = 121 in conf/demo_sfs.conf): one warp per POINT
// keeps its reference patch in registers (4 samples per lane) and walks the source views; per view the patch is sampled
// into registers together with its bilinear gradient coefficients, the masked NCC statistics are warp-reduced, the value
// is stored and the analytic gradient (the arithmetic of ncc_bwd_kernel, mesh.cu) is contracted with the sampling
// gradients on the spot.  Nothing of size [Nv, Np, Npx] is ever written: the unfused chain moves ~2.6 GB of patches, masks
// and patch gradients through HBM per iteration for config 3, this kernel reads image pixels through L1 / L2 and issues
// nine atomics per point.  g = d loss / d ncc (the same constant for every (view, point): -weight / (Nv Np)).
//
// Sampling: the patch offsets are whole pixels, so all samples of a (point, view) share ONE fractional position
// (fx, fy) = (u - floor u, v - floor v) and read a (side + 1)^2 block of pixels: the warp stages that block (image and
// mask, zero outside the frame) in shared memory with ~5 coalesced loads each, and every sample is four shared-memory
// reads with the same weights - instead of four scattered global gathers, a floor and eight bounds tests per sample
// (measured on config 3: the per-sample form was bound by L1 wavefronts and address arithmetic, 1.07 ms per iteration).
constexpr int kNccTile = 12 * 12;  // (side + 1)^2 for side <= 11
constexpr int kNccMaxViews = 64;   // view slots of the fused kernel (reference + sources)
struct NccView { int x0, y0; float fx, fy; int bx, by; bool front; };
template <int HALF>
__device__ __forceinline__ NccView ncc_stage(const float* __restrict__ img, const float* __restrict__ msk, int H, int W,
                                             int half_rt, float u, float v, bool front, float* timg, float* tmsk, int lane) {
    const int half = HALF > 0 ? HALF : half_rt;  // compile-time for the configured patch sizes: t / ts becomes a multiply
    NccView nv;
    const float x0f = floorf(u), y0f = floorf(v);
    nv.x0 = (int)x0f; nv.y0 = (int)y0f; nv.fx = u - x0f; nv.fy = v - y0f;
    nv.bx = nv.fx >= 0.5f ? 1 : 0; nv.by = nv.fy >= 0.5f ? 1 : 0;
    nv.front = front;
    const int ts = 2 * half + 2;
    __syncwarp();  // the previous view's samples have been read
    for (int t = lane; t < ts * ts; t += 32) {
        const int ty = t / ts, tx = t - ty * ts;
        const int px = nv.x0 - half + tx, py = nv.y0 - half + ty;
        const bool in = front && px >= 0 && px < W && py >= 0 && py < H;
        timg[t] = in ? __ldg(img + (size_t)py * W + px) : 0.0f;
        tmsk[t] = (in && msk) ? __ldg(msk + (size_t)py * W + px) : 0.0f;
    }
    __syncwarp();
    return nv;
}
// sample (dx, dy) of the staged block: value, bilinear gradient coefficients, mask (same rules as ncc_taps)
template <int HALF>
__device__ __forceinline__ void ncc_tile_sample(const NccView& nv, const float* timg, const float* tmsk, int H, int W,
                                                int half_rt, int dx, int dy, float& val, float& du, float& dv, float& mk) {
    const int half = HALF > 0 ? HALF : half_rt;
    const int ts = 2 * half + 2, o = (dy + half) * ts + dx + half;
    const float i00 = timg[o], i10 = timg[o + 1], i01 = timg[o + ts], i11 = timg[o + ts + 1];
    val = (1.0f - nv.fy) * ((1.0f - nv.fx) * i00 + nv.fx * i10) + nv.fy * ((1.0f - nv.fx) * i01 + nv.fx * i11);
    du = (1.0f - nv.fy) * (i10 - i00) + nv.fy * (i11 - i01);
    dv = (1.0f - nv.fx) * (i01 - i00) + nv.fx * (i11 - i10);
    const int px = nv.x0 + dx, py = nv.y0 + dy;
    const bool inside = px >= 0 && px + 1 < W && py >= 0 && py + 1 < H;
    mk = (inside && tmsk[o + nv.by * ts + nv.bx] > 0.5f) ? 1.0f : 0.0f;
}

template <int HALF>
__global__ void __launch_bounds__(256, 3) ncc_term_fused_kernel(
    const float* __restrict__ vertices, const int32_t* __restrict__ tri, const int32_t* __restrict__ pt_face,
    const float* __restrict__ pt_bary, const float* __restrict__ w2cs, const float* __restrict__ projs,
    const int32_t* __restrict__ view_idx, int Nv1, const float* __restrict__ gray, const float* __restrict__ masks, int Np,
    int H, int W, int half_rt, float g, float* __restrict__ ncc, float* __restrict__ grad_vertices, int chunk) {
    const int half = HALF > 0 ? HALF : half_rt;
    __shared__ float timg_s[8][kNccTile], tmsk_s[8][kNccTile];
    // world -> clip matrices of the term's views, once per block (every lane of every warp multiplied them per view)
    __shared__ NccCam Ms[kNccMaxViews];
    for (int i = threadIdx.x; i < Nv1 * 16; i += blockDim.x) {
        const int sl = i >> 4, e = i & 15, rr = e >> 2, jj = e & 3;
        const int view = __ldg(view_idx + sl);
        const float* Wm = w2cs + (size_t)view * 16;
        const float* Pm = projs + (size_t)view * 16;
        Ms[sl].m[e] = __fmaf_rn(Wm[4 * rr + 3], Pm[12 + jj], __fmaf_rn(Wm[4 * rr + 2], Pm[8 + jj],
                                __fmaf_rn(Wm[4 * rr + 1], Pm[4 + jj], __fmul_rn(Wm[4 * rr], Pm[jj]))));
    }
    __syncthreads();
    // a warp owns one point and `chunk` consecutive source views (the reference patch is re-sampled per chunk)
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_chunks = (Nv1 - 1 + chunk - 1) / chunk;
    const int p = (int)(wid / n_chunks), ch = (int)(wid - (long long)p * n_chunks);
    const int lane = threadIdx.x & 31;
    if (p >= Np) return;  // (whole warps: no partial-warp exit before the __syncwarp()s below)
    float* timg = timg_s[threadIdx.x >> 5];
    float* tmsk = tmsk_s[threadIdx.x >> 5];
    const int s_begin = 1 + ch * chunk, s_end = min(Nv1, s_begin + chunk);
    const NccPoint q = ncc_point(vertices, tri, pt_face, pt_bary, p);
    const int side = 2 * half + 1, npx = side * side;
    int dxs[4], dys[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int k = min(lane + 32 * j, npx - 1);
        dys[j] = k / side - half;
        dxs[j] = k - (k / side) * side - half;
    }
    // reference patch (view slot 0): a constant of the step
    float r[4] = {0.f, 0.f, 0.f, 0.f};
    {
        const int view = __ldg(view_idx);
        const NccCam& M = Ms[0];
        float u, v;
        float4 clip;
        ncc_project(M, q.X, H, W, u, v, clip);
        const NccView nv = ncc_stage<HALF>(gray + (size_t)view * H * W, nullptr, H, W, half, u, v, clip.w > 0.0f, timg, tmsk, lane);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float du, dv, mk;
            if (lane + 32 * j < npx) ncc_tile_sample<HALF>(nv, timg, tmsk, H, W, half, dxs[j], dys[j], r[j], du, dv, mk);
        }
    }
    float3 gXsum = make_float3(0.f, 0.f, 0.f);
    for (int s = s_begin; s < s_end; s++) {
        const int view = __ldg(view_idx + s);
        const NccCam& M = Ms[s];
        float u, v;
        float4 clip;
        ncc_project(M, q.X, H, W, u, v, clip);
        const bool front = clip.w > 0.0f;
        const NccView nv = ncc_stage<HALF>(gray + (size_t)view * H * W, masks + (size_t)view * H * W, H, W, half, u, v, front, timg,
                                     tmsk, lane);
        float sv[4], mk[4], du[4], dv[4];
        float cnt = 0.f, sr = 0.f, ss = 0.f;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            sv[j] = 0.f; mk[j] = 0.f; du[j] = 0.f; dv[j] = 0.f;
            if (front && lane + 32 * j < npx) ncc_tile_sample<HALF>(nv, timg, tmsk, H, W, half, dxs[j], dys[j], sv[j], du[j], dv[j], mk[j]);
            cnt += mk[j]; sr += r[j] * mk[j]; ss += sv[j] * mk[j];
        }
        cnt = warp_sum(cnt); sr = warp_sum(sr); ss = warp_sum(ss);
        if (cnt == 0.0f) cnt = 1.0f;
        const float rm = sr / cnt, sm = ss / cnt;
        float vr = 0.f, vs = 0.f, cv = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (lane + 32 * j >= npx) continue;
            const float dr = (r[j] - rm) * mk[j], dsv = (sv[j] - sm) * mk[j];
            vr += dr * dr; vs += dsv * dsv; cv += (r[j] - rm) * (sv[j] - sm) * mk[j];
            s1 += dsv * mk[j]; s2 += dr;
        }
        vr = warp_sum(vr) / cnt; vs = warp_sum(vs) / cnt; cv = warp_sum(cv) / cnt;
        s1 = warp_sum(s1); s2 = warp_sum(s2);
        if (vr == 0.0f) vr = 1.0f;
        if (vs == 0.0f) vs = 1.0f;
        if (lane == 0) ncc[(size_t)(s - 1) * Np + p] = cv / (sqrtf(vr) * sqrtf(vs));
        if (!front) continue;
        const float ir = rsqrtf(vr), is = rsqrtf(vs);
        const float c_cv = g * ir * is / cnt, c_vs = -g * cv * ir * is / vs / cnt;
        float gu = 0.f, gv = 0.f;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (lane + 32 * j >= npx) continue;
            const float dcv = (r[j] - rm) * mk[j] - mk[j] / cnt * s2;
            const float dvs = (sv[j] - sm) * mk[j] * mk[j] - mk[j] / cnt * s1;
            const float gk = c_cv * dcv + c_vs * dvs;
            gu += gk * du[j]; gv += gk * dv[j];
        }
        gu = warp_sum(gu); gv = warp_sum(gv);
        const float iw = 1.0f / clip.w, hu = 0.5f * (float)W * gu * iw, hv = 0.5f * (float)H * gv * iw;
        const float gcw = -(hu * clip.x + hv * clip.y) * iw;
        gXsum.x += M.m[0] * hu + M.m[1] * hv + M.m[3] * gcw;
        gXsum.y += M.m[4] * hu + M.m[5] * hv + M.m[7] * gcw;
        gXsum.z += M.m[8] * hu + M.m[9] * hv + M.m[11] * gcw;
    }
    if (lane != 0 || (gXsum.x == 0.0f && gXsum.y == 0.0f && gXsum.z == 0.0f)) return;
    const int vi[3] = {q.i0, q.i1, q.i2};
    const float bw[3] = {q.b0, q.b1, q.b2};
#pragma unroll
    for (int c = 0; c < 3; c++) {
        float* o = grad_vertices + 3 * (size_t)vi[c];
        atomicAdd(o, bw[c] * gXsum.x); atomicAdd(o + 1, bw[c] * gXsum.y); atomicAdd(o + 2, bw[c] * gXsum.z);
    }
}

}  // namespace fmhr

using namespace fmhr;

constexpr int kNccViewChunk = 64;  // source views per warp of the fused kernel: all of them (sweep on config 3, 1 / 3 / 5 / 15
                                   // views per warp: 1.96 / 1.44 / 1.35 / 1.27 ms per iteration; override: FMHR_NCC_CHUNK)
static int ncc_check(const void* vertices, const void* tri, const void* pt_face, const void* pt_bary, const void* w2cs,
                     const void* projs, const void* view_idx, int Nv1, const void* gray, int V, int Np, int H, int W, int half) {
    FMHR_CHECK_ARG(vertices && tri && pt_face && pt_bary && w2cs && projs && view_idx && gray);
    FMHR_CHECK_ARG(Nv1 >= 2 && V > 0 && Np > 0 && H > 0 && W > 0 && half >= 0 && half <= 16);
    FMHR_CHECK_ARG((long long)Nv1 * Np < (1ll << 31) / 32);
    return FMHR_OK;
}

extern "C" int fmhr_ncc_sample_fwd(const float* vertices, const int32_t* tri, const int32_t* pt_face, const float* pt_bary,
                                   const float* w2cs, const float* projs, const int32_t* view_idx, int Nv1,
                                   const float* gray, const float* masks, int V, int Np, int H, int W, int half,
                                   float* patches, float* patch_mask, fmhr_stream_t stream) {
    int rc = ncc_check(vertices, tri, pt_face, pt_bary, w2cs, projs, view_idx, Nv1, gray, V, Np, H, W, half);
    if (rc) return rc;
    FMHR_CHECK_ARG(masks && patches && patch_mask);
    const long long threads = (long long)Nv1 * Np * 32;
    ncc_sample_fwd_kernel<<<cdiv(threads, 256), 256, 0, (cudaStream_t)stream>>>(
        vertices, tri, pt_face, pt_bary, w2cs, projs, view_idx, Nv1, gray, masks, Np, H, W, half, patches, patch_mask);
    FMHR_LAUNCH_CHECK();
    return FMHR_OK;
}

extern "C" int fmhr_ncc_sample_bwd(const float* vertices, const int32_t* tri, const int32_t* pt_face, const float* pt_bary,
                                   const float* w2cs, const float* projs, const int32_t* view_idx, int Nv1,
                                   const float* gray, int V, int Np, int H, int W, int half, const float* grad_patches,
                                   float* grad_vertices, fmhr_stream_t stream) {
    int rc = ncc_check(vertices, tri, pt_face, pt_bary, w2cs, projs, view_idx, Nv1, gray, V, Np, H, W, half);
    if (rc) return rc;
    FMHR_CHECK_ARG(grad_patches && grad_vertices);
    const long long threads = (long long)(Nv1 - 1) * Np * 32;
    ncc_sample_bwd_kernel<<<cdiv(threads, 256), 256, 0, (cudaStream_t)stream>>>(
        vertices, tri, pt_face, pt_bary, w2cs, projs, view_idx, Nv1, gray, Np, H, W, half, grad_patches, grad_vertices);
    FMHR_LAUNCH_CHECK();
    return FMHR_OK;
}

extern "C" int fmhr_ncc_term_fused(const float* vertices, const int32_t* tri, const int32_t* pt_face, const float* pt_bary,
                                   const float* w2cs, const float* projs, const int32_t* view_idx, int Nv1,
                                   const float* gray, const float* masks, int V, int Np, int H, int W, int half,
                                   float grad_ncc, float* ncc, float* grad_vertices, fmhr_stream_t stream) {
    int rc = ncc_check(vertices, tri, pt_face, pt_bary, w2cs, projs, view_idx, Nv1, gray, V, Np, H, W, half);
    if (rc) return rc;
    FMHR_CHECK_ARG(masks && ncc && grad_vertices);
    if ((2 * half + 1) * (2 * half + 1) > 128 || Nv1 > kNccMaxViews) {
        set_error("fmhr_ncc_term_fused: patches of more than 128 samples need the unfused chain (fmhr_ncc_sample_fwd / "
                  "fmhr_ncc_fwd / fmhr_ncc_bwd / fmhr_ncc_sample_bwd)");
        return FMHR_EUNSUPPORTED;
    }
    static const int env_chunk = [] { const char* e = getenv("FMHR_NCC_CHUNK"); return e ? atoi(e) : 0; }();
    const int chunk = max(1, min(Nv1 - 1, env_chunk > 0 ? env_chunk : kNccViewChunk));
    const long long warps = (long long)Np * ((Nv1 - 1 + chunk - 1) / chunk);
#define FMHR_NCC_FUSED(HALF)                                                                                              \
    ncc_term_fused_kernel<HALF><<<cdiv(warps * 32, 256), 256, 0, (cudaStream_t)stream>>>(                                 \
        vertices, tri, pt_face, pt_bary, w2cs, projs, view_idx, Nv1, gray, masks, Np, H, W, half, grad_ncc, ncc,          \
        grad_vertices, chunk)
    if (half == 5) FMHR_NCC_FUSED(5);
    else if (half == 2) FMHR_NCC_FUSED(2);
    else FMHR_NCC_FUSED(0);
#undef FMHR_NCC_FUSED
    FMHR_LAUNCH_CHECK();
    return FMHR_OK;
}
