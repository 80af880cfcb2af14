// Mesh-domain kernels: topology build (once per mesh), vertex normals (models/utils.py:508-548), uniform
// Laplacian regulariser as CSR SpMV (models/utils.py:661-722), SH radiance (models/utils.py:208-226) and
// NCC (models/ncc_utils.py:4-35).  All gathers run over CSR rows, so none of the forward/backward passes here
// needs an atomic.
#include <cub/cub.cuh>

#include "common.cuh"

namespace fmhr {

// ------------------------------------------------------------------------------------------------
// topology
// ------------------------------------------------------------------------------------------------
__global__ void topo_edge_keys_kernel(const int32_t* __restrict__ tri, int T, unsigned long long* __restrict__ keys,
                                      int32_t* __restrict__ vals) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;  // j = 3t + k, edge opposite corner k
    if (j >= 3 * T) return;
    const int t = j / 3, k = j - 3 * t;
    const uint32_t a = (uint32_t)tri[3 * t + (k + 1) % 3], b = (uint32_t)tri[3 * t + (k + 2) % 3];
    keys[j] = ((unsigned long long)min(a, b) << 32) | max(a, b);
    vals[j] = j;
}

__global__ void topo_opp_kernel(const int32_t* __restrict__ tri, int T, const unsigned long long* __restrict__ keys,
                                const int32_t* __restrict__ vals, int32_t* __restrict__ opp) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= 3 * T) return;
    const unsigned long long key = keys[j];
    int s = j;
    while (s > 0 && keys[s - 1] == key) s--;  // run start (runs are 1-2 long on a manifold)
    int cand = -1;
    if (s != j) cand = s;
    else if (j + 1 < 3 * T && keys[j + 1] == key) cand = j + 1;
    // the sort is stable, so `cand` is the lowest-indexed other triangle on this edge
    opp[vals[j]] = (cand < 0) ? -1 : tri[vals[cand]];  // vals = 3t'+k' and tri[3t'+k'] is the corner facing the edge
}

__global__ void topo_v2f_keys_kernel(const int32_t* __restrict__ tri, int T, uint32_t* __restrict__ keys,
                                     int32_t* __restrict__ vals) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= 3 * T) return;
    const int t = j / 3, k = j - 3 * t;
    keys[j] = (uint32_t)tri[j];
    vals[j] = t * 4 + k;
}

__global__ void topo_rowptr32_kernel(const uint32_t* __restrict__ sorted, int n, int V, int32_t* __restrict__ rowptr) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v > V) return;
    int lo = 0, hi = n;  // first index with key >= v
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (sorted[mid] < (uint32_t)v) lo = mid + 1; else hi = mid;
    }
    rowptr[v] = lo;
}

__global__ void topo_v2v_keys_kernel(const int32_t* __restrict__ tri, int T, unsigned long long* __restrict__ keys) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= 3 * T) return;
    const int t = j / 3, k = j - 3 * t;
    const uint32_t a = (uint32_t)tri[3 * t + k], b = (uint32_t)tri[3 * t + (k + 1) % 3];
    keys[2 * j] = ((unsigned long long)a << 32) | b;
    keys[2 * j + 1] = ((unsigned long long)b << 32) | a;
}

__global__ void topo_v2v_finish_kernel(const unsigned long long* __restrict__ uniq, const int* __restrict__ n_uniq,
                                       int V, int32_t* __restrict__ rowptr, int32_t* __restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = *n_uniq;
    if (i < n) idx[i] = (int32_t)(uint32_t)uniq[i];
    if (i <= V) {
        const unsigned long long target = (unsigned long long)(uint32_t)i << 32;
        int lo = 0, hi = n;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (uniq[mid] < target) lo = mid + 1; else hi = mid;
        }
        rowptr[i] = lo;
    }
}

// Derived adjacency that removes one dependent load from every gather loop of the vertex-domain kernels:
//   v2f_nbr[j] = the two other corners (cyclic order) of incident face entry j  -> no tri[] lookup
//   inv_deg[i] = 1 / degree(i) (0 for isolated vertices)                         -> no second row-pointer lookup
__global__ void topo_derive_kernel(const int32_t* __restrict__ tri, const int32_t* __restrict__ v2f_idx,
                                   const int32_t* __restrict__ v2v_ptr, int V, int T, int2* __restrict__ v2f_nbr,
                                   float* __restrict__ inv_deg) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < 3 * T) {
        const int tk = v2f_idx[j], t = tk >> 2, k = tk & 3;
        v2f_nbr[j] = make_int2(tri[3 * t + (k + 1) % 3], tri[3 * t + (k + 2) % 3]);
    }
    if (j < V) {
        const int d = v2v_ptr[j + 1] - v2v_ptr[j];
        inv_deg[j] = d > 0 ? 1.0f / (float)d : 0.0f;
    }
}

// ------------------------------------------------------------------------------------------------
// vertex normals
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float3 ld3(const float* p) { return make_float3(__ldg(p), __ldg(p + 1), __ldg(p + 2)); }
__device__ __forceinline__ float3 sub3(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 add3(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 cross3(float3 a, float3 b) {
    return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// 8 lanes per vertex: each lane takes every 8th incident face, partial sums are shuffle-reduced (V ~ 50k alone would
// leave the GPU at a fraction of a wave of latency-bound gather loops).
__global__ void __launch_bounds__(256) vertex_normals_fwd_kernel(const float* __restrict__ verts,
                                                                 const int32_t* __restrict__ tri,
                                                                 const int32_t* __restrict__ v2f_ptr,
                                                                 const int32_t* __restrict__ v2f_idx,
                                                                 const int2* __restrict__ v2f_nbr, int V,
                                                                 float* __restrict__ normals, float* __restrict__ raw) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, sub = threadIdx.x & 7;
    float3 acc = make_float3(0.f, 0.f, 0.f);
    if (i < V) {
        const int b = __ldg(v2f_ptr + i), e = __ldg(v2f_ptr + i + 1);
        const float3 pk = ld3(verts + 3 * (size_t)i);
        for (int j = b + sub; j < e; j += 8) {
            int ia, ib;
            if (v2f_nbr) { const int2 nb = __ldg(v2f_nbr + j); ia = nb.x; ib = nb.y; }
            else {
                const int tk = __ldg(v2f_idx + j), t = tk >> 2, k = tk & 3;
                ia = __ldg(tri + 3 * t + (k + 1) % 3); ib = __ldg(tri + 3 * t + (k + 2) % 3);
            }
            // corner-k form of the face normal, models/utils.py:517-543
            acc = add3(acc, cross3(sub3(ld3(verts + 3 * (size_t)ia), pk), sub3(ld3(verts + 3 * (size_t)ib), pk)));
        }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
        acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
    }
    if (i >= V || sub != 0) return;
    const float len = sqrtf(dot3(acc, acc));
    const float inv = 1.0f / fmaxf(len, 1e-6f);
    if (raw) { raw[3 * (size_t)i] = acc.x; raw[3 * (size_t)i + 1] = acc.y; raw[3 * (size_t)i + 2] = acc.z; }
    normals[3 * (size_t)i] = acc.x * inv;
    normals[3 * (size_t)i + 1] = acc.y * inv;
    normals[3 * (size_t)i + 2] = acc.z * inv;
}

// gN = d(loss)/d(raw) from d(loss)/d(normalised)
__global__ void __launch_bounds__(128) vertex_normals_bwd_project_kernel(const float* __restrict__ raw,
                                                                        const float* __restrict__ grad_normals, int V,
                                                                        float scale, float* __restrict__ gN) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const float3 N = ld3(raw + 3 * (size_t)i);
    float3 g = ld3(grad_normals + 3 * (size_t)i);
    g = make_float3(g.x * scale, g.y * scale, g.z * scale);
    const float len = sqrtf(dot3(N, N));
    float3 r;
    if (len > 1e-6f) {
        const float inv = 1.0f / len;
        const float3 nh = make_float3(N.x * inv, N.y * inv, N.z * inv);
        const float d = dot3(nh, g);
        r = make_float3((g.x - nh.x * d) * inv, (g.y - nh.y * d) * inv, (g.z - nh.z * d) * inv);
    } else {
        r = make_float3(g.x * 1e6f, g.y * 1e6f, g.z * 1e6f);
    }
    gN[3 * (size_t)i] = r.x; gN[3 * (size_t)i + 1] = r.y; gN[3 * (size_t)i + 2] = r.z;
}

__device__ __forceinline__ float3 normals_bwd_gather(int i, const float* __restrict__ verts,
                                                     const int32_t* __restrict__ tri,
                                                     const int32_t* __restrict__ v2f_ptr,
                                                     const int32_t* __restrict__ v2f_idx,
                                                     const float* __restrict__ gN) {
    float3 acc = make_float3(0.f, 0.f, 0.f);
    const int b = __ldg(v2f_ptr + i), e = __ldg(v2f_ptr + i + 1);
    const float3 gi = ld3(gN + 3 * (size_t)i);
    for (int j = b; j < e; j++) {
        const int tk = __ldg(v2f_idx + j), t = tk >> 2, k = tk & 3;
        const int ia = __ldg(tri + 3 * t + (k + 1) % 3), ib = __ldg(tri + 3 * t + (k + 2) % 3);
        const float3 G = add3(gi, add3(ld3(gN + 3 * (size_t)ia), ld3(gN + 3 * (size_t)ib)));
        acc = add3(acc, cross3(sub3(ld3(verts + 3 * (size_t)ia), ld3(verts + 3 * (size_t)ib)), G));
    }
    return acc;
}

__global__ void __launch_bounds__(128) vertex_normals_bwd_gather_kernel(const float* __restrict__ verts,
                                                                       const int32_t* __restrict__ tri,
                                                                       const int32_t* __restrict__ v2f_ptr,
                                                                       const int32_t* __restrict__ v2f_idx,
                                                                       const float* __restrict__ gN, int V,
                                                                       float* __restrict__ grad_verts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const float3 a = normals_bwd_gather(i, verts, tri, v2f_ptr, v2f_idx, gN);
    grad_verts[3 * (size_t)i] = a.x; grad_verts[3 * (size_t)i + 1] = a.y; grad_verts[3 * (size_t)i + 2] = a.z;
}

// ------------------------------------------------------------------------------------------------
// uniform Laplacian  y = D^-1 A x - x ; loss = sum ||y_i|| / V
// ------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(128) laplacian_fwd_kernel(const float* __restrict__ x,
                                                            const int32_t* __restrict__ v2v_ptr,
                                                            const int32_t* __restrict__ v2v_idx, int V,
                                                            float* __restrict__ yhat, float* __restrict__ loss) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float li = 0.0f;
    if (i < V) {
        const int b = __ldg(v2v_ptr + i), e = __ldg(v2v_ptr + i + 1);
        float s[C];
#pragma unroll
        for (int c = 0; c < C; c++) s[c] = 0.0f;
        for (int j = b; j < e; j++) {
            const size_t nb = (size_t)__ldg(v2v_idx + j) * C;
#pragma unroll
            for (int c = 0; c < C; c++) s[c] += __ldg(x + nb + c);
        }
        const float invd = (e > b) ? 1.0f / (float)(e - b) : 0.0f;
        float n2 = 0.0f;
#pragma unroll
        for (int c = 0; c < C; c++) {
            s[c] = s[c] * invd - __ldg(x + (size_t)i * C + c);
            n2 += s[c] * s[c];
        }
        li = sqrtf(n2);
        const float inv = (li > 0.0f) ? 1.0f / li : 0.0f;
#pragma unroll
        for (int c = 0; c < C; c++) yhat[(size_t)i * C + c] = s[c] * inv;
    }
    li = warp_sum(li);
    __shared__ float ws[4];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = li;
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(loss, (ws[0] + ws[1] + ws[2] + ws[3]) / (float)V);
}

template <int C>
__global__ void __launch_bounds__(128) laplacian_bwd_kernel(const float* __restrict__ yhat,
                                                            const int32_t* __restrict__ v2v_ptr,
                                                            const int32_t* __restrict__ v2v_idx, int V, float scale,
                                                            float* __restrict__ grad_x) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= V) return;
    const int b = __ldg(v2v_ptr + j), e = __ldg(v2v_ptr + j + 1);
    float s[C];
#pragma unroll
    for (int c = 0; c < C; c++) s[c] = 0.0f;
    for (int q = b; q < e; q++) {
        const int i = __ldg(v2v_idx + q);
        const float invd = 1.0f / (float)(__ldg(v2v_ptr + i + 1) - __ldg(v2v_ptr + i));
#pragma unroll
        for (int c = 0; c < C; c++) s[c] += __ldg(yhat + (size_t)i * C + c) * invd;
    }
    const float sc = scale / (float)V;
#pragma unroll
    for (int c = 0; c < C; c++) grad_x[(size_t)j * C + c] = (s[c] - __ldg(yhat + (size_t)j * C + c)) * sc;
}

// ------------------------------------------------------------------------------------------------
// SH radiance, degree 3 (9 coefficients)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sh_radiance_fwd_kernel(const float* __restrict__ coeff, int coeff_rows,
                                                              const float* __restrict__ normal, int n,
                                                              float* __restrict__ radiance) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* c = coeff + (coeff_rows == 1 ? 0 : (size_t)i * 9);
    const float x = normal[3 * (size_t)i], y = normal[3 * (size_t)i + 1], z = normal[3 * (size_t)i + 2];
    float r = c[0];
    r = r + c[1] * y;
    r = r + c[2] * z;
    r = r + c[3] * x;
    r = r + c[4] * x * y;
    r = r + c[5] * y * z;
    r = r + c[6] * (2 * z * z - x * x - y * y);
    r = r + c[7] * z * x;
    r = r + c[8] * (x * x - y * y);
    radiance[i] = r;
}

__global__ void __launch_bounds__(256) sh_radiance_bwd_kernel(const float* __restrict__ coeff, int coeff_rows,
                                                              const float* __restrict__ normal,
                                                              const float* __restrict__ grad_radiance, int n,
                                                              float* __restrict__ grad_coeff,
                                                              float* __restrict__ grad_normal) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float gc[9];
#pragma unroll
    for (int k = 0; k < 9; k++) gc[k] = 0.0f;
    if (i < n) {
        const float* c = coeff + (coeff_rows == 1 ? 0 : (size_t)i * 9);
        const float x = normal[3 * (size_t)i], y = normal[3 * (size_t)i + 1], z = normal[3 * (size_t)i + 2];
        const float g = grad_radiance[i];
        gc[0] = g; gc[1] = g * y; gc[2] = g * z; gc[3] = g * x; gc[4] = g * x * y; gc[5] = g * y * z;
        gc[6] = g * (2 * z * z - x * x - y * y); gc[7] = g * z * x; gc[8] = g * (x * x - y * y);
        if (grad_normal) {
            grad_normal[3 * (size_t)i] = g * (c[3] + c[4] * y - 2 * c[6] * x + c[7] * z + 2 * c[8] * x);
            grad_normal[3 * (size_t)i + 1] = g * (c[1] + c[4] * x + c[5] * z - 2 * c[6] * y - 2 * c[8] * y);
            grad_normal[3 * (size_t)i + 2] = g * (c[2] + c[5] * y + 4 * c[6] * z + c[7] * x);
        }
        if (coeff_rows != 1) {
#pragma unroll
            for (int k = 0; k < 9; k++) grad_coeff[(size_t)i * 9 + k] = gc[k];
        }
    }
    if (coeff_rows == 1) {  // broadcast coefficients: warp-shuffle reduce, one atomic per warp and coefficient
#pragma unroll
        for (int k = 0; k < 9; k++) {
            const float s = warp_sum(gc[k]);
            if ((threadIdx.x & 31) == 0 && s != 0.0f) atomicAdd(grad_coeff + k, s);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// NCC: one warp per (view, point); two passes over the patch with warp-shuffle reductions
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ncc_fwd_kernel(const float* __restrict__ ref, const float* __restrict__ src,
                                                      const float* __restrict__ mask, int Nv, int Np, int Npx,
                                                      float* __restrict__ ncc) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= Nv * Np) return;
    const int p = warp % Np;
    const float* r = ref + (size_t)p * Npx;
    const float* s = src + (size_t)warp * Npx;
    const float* m = mask + (size_t)warp * Npx;
    float cnt = 0.f, sr = 0.f, ss = 0.f;
    for (int k = lane; k < Npx; k += 32) {
        const float mk = m[k];
        cnt += mk; sr += r[k] * mk; ss += s[k] * mk;
    }
    cnt = warp_sum(cnt); sr = warp_sum(sr); ss = warp_sum(ss);
    if (cnt == 0.0f) cnt = 1.0f;
    const float rm = sr / cnt, sm = ss / cnt;
    float vr = 0.f, vs = 0.f, cv = 0.f;
    for (int k = lane; k < Npx; k += 32) {
        const float mk = m[k];
        const float dr = (r[k] - rm) * mk, dsv = (s[k] - sm) * mk;
        vr += dr * dr; vs += dsv * dsv; cv += (r[k] - rm) * (s[k] - sm) * mk;
    }
    vr = warp_sum(vr) / cnt; vs = warp_sum(vs) / cnt; cv = warp_sum(cv) / cnt;
    if (vr == 0.0f) vr = 1.0f;
    if (vs == 0.0f) vs = 1.0f;
    if (lane == 0) ncc[warp] = cv / (sqrtf(vr) * sqrtf(vs));
}

// d(sum g * ncc)/d(src): three warp passes over the patch (means; variances / covariance; per-pixel gradient)
__global__ void __launch_bounds__(256) ncc_bwd_kernel(const float* __restrict__ ref, const float* __restrict__ src,
                                                      const float* __restrict__ mask, const float* __restrict__ gout,
                                                      int Nv, int Np, int Npx, float* __restrict__ grad_src) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= Nv * Np) return;
    const int p = warp % Np;
    const float* r = ref + (size_t)p * Npx;
    const float* s = src + (size_t)warp * Npx;
    const float* m = mask + (size_t)warp * Npx;
    float cnt = 0.f, sr = 0.f, ss = 0.f;
    for (int k = lane; k < Npx; k += 32) {
        const float mk = m[k];
        cnt += mk; sr += r[k] * mk; ss += s[k] * mk;
    }
    cnt = warp_sum(cnt); sr = warp_sum(sr); ss = warp_sum(ss);
    if (cnt == 0.0f) cnt = 1.0f;
    const float rm = sr / cnt, sm = ss / cnt;
    float vr = 0.f, vs = 0.f, cv = 0.f, s1 = 0.f, s2 = 0.f;
    for (int k = lane; k < Npx; k += 32) {
        const float mk = m[k];
        const float dr = (r[k] - rm) * mk, dsv = (s[k] - sm) * mk;
        vr += dr * dr; vs += dsv * dsv; cv += (r[k] - rm) * (s[k] - sm) * mk;
        s1 += dsv * mk; s2 += dr;
    }
    vr = warp_sum(vr) / cnt; vs = warp_sum(vs) / cnt; cv = warp_sum(cv) / cnt;
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (vr == 0.0f) vr = 1.0f;
    if (vs == 0.0f) vs = 1.0f;
    const float g = gout[warp];
    const float ir = rsqrtf(vr), is = rsqrtf(vs);
    const float c_cv = g * ir * is / cnt;                    // d/d(cov) * 1/cnt
    const float c_vs = -g * cv * ir * is / vs / cnt;         // d/d(var_s) * 2/cnt  (the 1/2 and the 2 cancel)
    for (int k = lane; k < Npx; k += 32) {
        const float mk = m[k];
        const float dcv = (r[k] - rm) * mk - mk / cnt * s2;
        const float dvs = (s[k] - sm) * mk * mk - mk / cnt * s1;
        grad_src[(size_t)warp * Npx + k] = c_cv * dcv + c_vs * dvs;
    }
}

// exported to ham.cu
int launch_vertex_normals_fwd(const float* verts, const int32_t* tri, const int32_t* v2f_ptr, const int32_t* v2f_idx,
                              const int32_t* v2f_nbr, int V, float* normals, float* raw, cudaStream_t st) {
    vertex_normals_fwd_kernel<<<cdiv((long long)V * 8, 256), 256, 0, st>>>(verts, tri, v2f_ptr, v2f_idx,
                                                                           (const int2*)v2f_nbr, V, normals, raw);
    FMHR_LAUNCH_CHECK();
    return FMHR_OK;
}

}  // namespace fmhr

using namespace fmhr;

extern "C" size_t fmhr_mesh_topology_workspace_bytes(int V, int T) {
    if (V <= 0 || T <= 0) return 0;
    const size_t n6 = (size_t)6 * T;
    // 2 key buffers (64-bit, 6T) + 2 value buffers (32-bit, 6T) + unique output + CUB temp + counters
    return n6 * 8 * 3 + n6 * 4 * 2 + ((size_t)8 << 20) + n6 * 16 + 1024;
}

extern "C" int fmhr_mesh_topology_build(const int32_t* tri, int V, int T, int32_t* opp, int32_t* v2f_ptr,
                                        int32_t* v2f_idx, int32_t* v2v_ptr, int32_t* v2v_idx, int* n_dir_edges_host,
                                        void* workspace, size_t workspace_bytes, fmhr_stream_t stream) {
    FMHR_CHECK_ARG(tri && workspace && V > 0 && T > 0);
    FMHR_CHECK_ARG(workspace_bytes >= fmhr_mesh_topology_workspace_bytes(V, T));
    FMHR_CHECK_ARG((v2f_ptr == nullptr) == (v2f_idx == nullptr));
    FMHR_CHECK_ARG((v2v_ptr == nullptr) == (v2v_idx == nullptr));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n6 = (size_t)6 * T;
    const int n3 = 3 * T;
    char* w = (char*)workspace;
    unsigned long long* k64a = (unsigned long long*)w; w += n6 * 8;
    unsigned long long* k64b = (unsigned long long*)w; w += n6 * 8;
    unsigned long long* uniq = (unsigned long long*)w; w += n6 * 8;
    int32_t* v32a = (int32_t*)w; w += n6 * 4;
    int32_t* v32b = (int32_t*)w; w += n6 * 4;
    int* counter = (int*)w; w += 1024;
    void* cub_tmp = w;
    const size_t cub_bytes = workspace_bytes - (size_t)(w - (char*)workspace);
    size_t need = 0;
    int vbits = 1;
    while ((1ll << vbits) < (long long)V + 1) vbits++;

    if (opp) {
        topo_edge_keys_kernel<<<cdiv(n3, 256), 256, 0, st>>>(tri, T, k64a, v32a);
        FMHR_LAUNCH_CHECK();
        cub::DoubleBuffer<unsigned long long> dk(k64a, k64b);
        cub::DoubleBuffer<int32_t> dv(v32a, v32b);
        FMHR_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, dk, dv, n3, 0, 32 + vbits, st));
        FMHR_CHECK_ARG(need <= cub_bytes);
        FMHR_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, need, dk, dv, n3, 0, 32 + vbits, st));
        topo_opp_kernel<<<cdiv(n3, 256), 256, 0, st>>>(tri, T, dk.Current(), dv.Current(), opp);
        FMHR_LAUNCH_CHECK();
    }
    if (v2f_ptr) {
        uint32_t* k32a = (uint32_t*)k64a;
        uint32_t* k32b = (uint32_t*)k64b;
        topo_v2f_keys_kernel<<<cdiv(n3, 256), 256, 0, st>>>(tri, T, k32a, v32a);
        FMHR_LAUNCH_CHECK();
        cub::DoubleBuffer<uint32_t> dk(k32a, k32b);
        cub::DoubleBuffer<int32_t> dv(v32a, v32b);
        FMHR_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, dk, dv, n3, 0, vbits, st));
        FMHR_CHECK_ARG(need <= cub_bytes);
        FMHR_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp, need, dk, dv, n3, 0, vbits, st));
        topo_rowptr32_kernel<<<cdiv(V + 1, 256), 256, 0, st>>>(dk.Current(), n3, V, v2f_ptr);
        FMHR_LAUNCH_CHECK();
        FMHR_CUDA(cudaMemcpyAsync(v2f_idx, dv.Current(), (size_t)n3 * 4, cudaMemcpyDeviceToDevice, st));
    }
    int n_dir = 0;
    if (v2v_ptr) {
        topo_v2v_keys_kernel<<<cdiv(n3, 256), 256, 0, st>>>(tri, T, k64a);
        FMHR_LAUNCH_CHECK();
        cub::DoubleBuffer<unsigned long long> dk(k64a, k64b);
        FMHR_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, need, dk, (int)n6, 0, 32 + vbits, st));
        FMHR_CHECK_ARG(need <= cub_bytes);
        FMHR_CUDA(cub::DeviceRadixSort::SortKeys(cub_tmp, need, dk, (int)n6, 0, 32 + vbits, st));
        FMHR_CUDA(cub::DeviceSelect::Unique(nullptr, need, dk.Current(), uniq, counter, (int)n6, st));
        FMHR_CHECK_ARG(need <= cub_bytes);
        FMHR_CUDA(cub::DeviceSelect::Unique(cub_tmp, need, dk.Current(), uniq, counter, (int)n6, st));
        topo_v2v_finish_kernel<<<cdiv(max((long long)n6, (long long)V + 1), 256), 256, 0, st>>>(uniq, counter, V,
                                                                                              v2v_ptr, v2v_idx);
        FMHR_LAUNCH_CHECK();
        FMHR_CUDA(cudaMemcpyAsync(&n_dir, counter, sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    FMHR_CUDA(cudaStreamSynchronize(st));  // setup call: documented sync
    if (n_dir_edges_host) *n_dir_edges_host = n_dir;
    return FMHR_OK;
}

extern "C" int fmhr_mesh_topology_derive(const int32_t* tri, const int32_t* v2f_idx, const int32_t* v2v_ptr, int V, int T,
                                         int32_t* v2f_nbr, float* inv_deg, fmhr_stream_t stream) {
    FMHR_CHECK_ARG(tri && v2f_idx && v2v_ptr && v2f_nbr && inv_deg && V > 0 && T > 0);
    topo_derive_kernel<<<cdiv(max(3 * T, V), 256), 256, 0, (cudaStream_t)stream>>>(tri, v2f_idx, v2v_ptr, V, T,
                                                                                   (int2*)v2f_nbr, inv_deg);
    FMHR_LAUNCH_CHECK();
    return FMHR_OK;
}

extern "C" int fmhr_vertex_normals_fwd(const float* verts, const int32_t* tri, const int32_t* v2f_ptr,
                                       const int32_t* v2f_idx, int V, int T, float* normals, float* raw,
                                       fmhr_stream_t stream) {
    FMHR_CHECK_ARG(verts && tri && v2f_ptr && v2f_idx && normals && V > 0 && T > 0);
    return launch_vertex_normals_fwd(verts, tri, v2f_ptr, v2f_idx, nullptr, V, normals, raw, (cudaStream_t)stream);
}

extern "C" int fmhr_vertex_normals_bwd(const float* verts, const int32_t* tri, const int32_t* v2f_ptr,
                                       const int32_t* v2f_idx, const float* raw, const float* grad_normals, int V,
                                       int T, float* scratch, float* grad_verts, fmhr_stream_t stream) {
    FMHR_CHECK_ARG(verts && tri && v2f_ptr && v2f_idx && raw && grad_normals && scratch && grad_verts);
    FMHR_CHECK_ARG(V > 0 && T > 0);
    cudaStream_t st = (cudaStream_t)stream;
    vertex_normals_bwd_project_kernel<<<cdiv(V, 128), 128, 0, st>>>(raw, grad_normals, V, 1.0f, scratch);
    FMHR_LAUNCH_CHECK();
    vertex_normals_bwd_gather_kernel<<<cdiv(V, 128), 128, 0, st>>>(verts, tri, v2f_ptr, v2f_idx, scratch, V,
                                                                   grad_verts);
    FMHR_LAUNCH_CHECK();
    return FMHR_OK;
}

extern "C" int fmhr_laplacian_fwd(const float* x, const int32_t* v2v_ptr, const int32_t* v2v_idx, int V, int C,
                                  float* yhat, float* loss, fmhr_stream_t stream) {
    FMHR_CHECK_ARG(x && v2v_ptr && v2v_idx && yhat && loss && V > 0);
    FMHR_CHECK_ARG(C >= 1 && C <= 4);
    cudaStream_t st = (cudaStream_t)stream;
    FMHR_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
    const int g = cdiv(V, 128);
    switch (C) {
        case 1: laplacian_fwd_kernel<1><<<g, 128, 0, st>>>(x, v2v_ptr, v2v_idx, V, yhat, loss); break;
        case 2: laplacian_fwd_kernel<2><<<g, 128, 0, st>>>(x, v2v_ptr, v2v_idx, V, yhat, loss); break;
        case 3: laplacian_fwd_kernel<3><<<g, 128, 0, st>>>(x, v2v_ptr, v2v_idx, V, yhat, loss); break;
        default: laplacian_fwd_kernel<4><<<g, 128, 0, st>>>(x, v2v_ptr, v2v_idx, V, yhat, loss); break;
    }
    FMHR_LAUNCH_CHECK();
    return FMHR_OK;
}

extern "C" int fmhr_laplacian_bwd(const float* yhat, const int32_t* v2v_ptr, const int32_t* v2v_idx, int V, int C,
                                  float scale, float* grad_x, fmhr_stream_t stream) {
    FMHR_CHECK_ARG(yhat && v2v_ptr && v2v_idx && grad_x && V > 0);
    FMHR_CHECK_ARG(C >= 1 && C <= 4);
    cudaStream_t st = (cudaStream_t)stream;
    const int g = cdiv(V, 128);
    switch (C) {
        case 1: laplacian_bwd_kernel<1><<<g, 128, 0, st>>>(yhat, v2v_ptr, v2v_idx, V, scale, grad_x); break;
        case 2: laplacian_bwd_kernel<2><<<g, 128, 0, st>>>(yhat, v2v_ptr, v2v_idx, V, scale, grad_x); break;
        case 3: laplacian_bwd_kernel<3><<<g, 128, 0, st>>>(yhat, v2v_ptr, v2v_idx, V, scale, grad_x); break;
        default: laplacian_bwd_kernel<4><<<g, 128, 0, st>>>(yhat, v2v_ptr, v2v_idx, V, scale, grad_x); break;
    }
    FMHR_LAUNCH_CHECK();
    return FMHR_OK;
}

extern "C" int fmhr_sh_radiance_fwd(const float* coeff, int coeff_rows, const float* normal, int n, float* radiance,
                                    fmhr_stream_t stream) {
    FMHR_CHECK_ARG(coeff && normal && radiance && n >= 0);
    FMHR_CHECK_ARG(coeff_rows == 1 || coeff_rows == n);
    if (n == 0) return FMHR_OK;
    sh_radiance_fwd_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(coeff, coeff_rows, normal, n, radiance);
    FMHR_LAUNCH_CHECK();
    return FMHR_OK;
}

extern "C" int fmhr_sh_radiance_bwd(const float* coeff, int coeff_rows, const float* normal,
                                    const float* grad_radiance, int n, float* grad_coeff, float* grad_normal,
                                    fmhr_stream_t stream) {
    FMHR_CHECK_ARG(coeff && normal && grad_radiance && grad_coeff && n >= 0);
    FMHR_CHECK_ARG(coeff_rows == 1 || coeff_rows == n);
    cudaStream_t st = (cudaStream_t)stream;
    if (coeff_rows == 1) FMHR_CUDA(cudaMemsetAsync(grad_coeff, 0, 9 * sizeof(float), st));
    if (n == 0) return FMHR_OK;
    sh_radiance_bwd_kernel<<<cdiv(n, 256), 256, 0, st>>>(coeff, coeff_rows, normal, grad_radiance, n, grad_coeff,
                                                         grad_normal);
    FMHR_LAUNCH_CHECK();
    return FMHR_OK;
}

extern "C" int fmhr_ncc_bwd(const float* ref, const float* src, const float* src_mask, const float* grad_ncc, int Nv,
                            int Np, int Npx, float* grad_src, fmhr_stream_t stream) {
    FMHR_CHECK_ARG(ref && src && src_mask && grad_ncc && grad_src && Nv > 0 && Np > 0 && Npx > 0);
    const long long threads = (long long)Nv * Np * 32;
    ncc_bwd_kernel<<<cdiv(threads, 256), 256, 0, (cudaStream_t)stream>>>(ref, src, src_mask, grad_ncc, Nv, Np, Npx,
                                                                         grad_src);
    FMHR_LAUNCH_CHECK();
    return FMHR_OK;
}

extern "C" int fmhr_ncc_fwd(const float* ref, const float* src, const float* src_mask, int Nv, int Np, int Npx,
                            float* ncc, fmhr_stream_t stream) {
    FMHR_CHECK_ARG(ref && src && src_mask && ncc && Nv > 0 && Np > 0 && Npx > 0);
    const long long threads = (long long)Nv * Np * 32;
    ncc_fwd_kernel<<<cdiv(threads, 256), 256, 0, (cudaStream_t)stream>>>(ref, src, src_mask, Nv, Np, Npx, ncc);
    FMHR_LAUNCH_CHECK();
    return FMHR_OK;
}
