// Meshlets for the fused HAM coverage kernel (setup time, once per mesh; host code).
//
// The coverage kernel used to gather three snapped (view, vertex) records per triangle from a per-view array: 3 x n x F
// random 8-byte gathers per iteration, which made it bound by L1 wavefronts, plus a separate transform kernel that wrote
// 32 bytes per (view, vertex).  With meshlets a block owns <= 1024 spatially coherent triangles and the <= 1024 distinct
// vertices they reference: it transforms those vertices into its view ONCE into shared memory and every triangle then
// indexes shared memory with 10-bit local indices.
//
// Faces are ordered along a Morton curve of their centroids (any input order works; the z-buffer still stores the
// ORIGINAL triangle id, so the equal-depth tie-break "lower triangle index wins" is unchanged).
#include <algorithm>
#include <numeric>
#include <vector>
#include "common.cuh"

namespace {

inline uint32_t spread10(uint32_t x) {  // 10 bits -> every third bit
    x &= 0x3ffu;
    x = (x | (x << 16)) & 0x030000ffu;
    x = (x | (x << 8)) & 0x0300f00fu;
    x = (x | (x << 4)) & 0x030c30c3u;
    x = (x | (x << 2)) & 0x09249249u;
    return x;
}

}  // namespace

using namespace fmhr;

extern "C" int fmhr_meshlets_build_host(const int32_t* tri, const float* verts, int V, int T, int tris_per_meshlet,
                                        int* n_meshlets, int* n_vert_refs, int* max_verts, int32_t* ml_vptr,
                                        int32_t* ml_verts, uint32_t* ml_tri2) {
    FMHR_CHECK_ARG(tri && V > 0 && T > 0 && n_meshlets && n_vert_refs && max_verts);
    FMHR_CHECK_ARG(tris_per_meshlet >= 32 && tris_per_meshlet <= 1024 && tris_per_meshlet % 32 == 0);
    FMHR_CHECK_ARG((ml_vptr == nullptr) == (ml_verts == nullptr) && (ml_vptr == nullptr) == (ml_tri2 == nullptr));
    for (int t = 0; t < 3 * T; t++) FMHR_CHECK_ARG(tri[t] >= 0 && tri[t] < V);
    std::vector<int> order(T);
    std::iota(order.begin(), order.end(), 0);
    if (verts) {
        float lo[3] = {verts[0], verts[1], verts[2]}, hi[3] = {verts[0], verts[1], verts[2]};
        for (int i = 0; i < V; i++)
            for (int c = 0; c < 3; c++) {
                lo[c] = std::min(lo[c], verts[3 * (size_t)i + c]);
                hi[c] = std::max(hi[c], verts[3 * (size_t)i + c]);
            }
        float ext = std::max(hi[0] - lo[0], std::max(hi[1] - lo[1], hi[2] - lo[2]));
        if (!(ext > 0.0f)) ext = 1.0f;
        std::vector<uint32_t> code(T);
        for (int t = 0; t < T; t++) {
            uint32_t q[3];
            for (int c = 0; c < 3; c++) {
                const float m = (verts[3 * (size_t)tri[3 * t] + c] + verts[3 * (size_t)tri[3 * t + 1] + c] +
                                 verts[3 * (size_t)tri[3 * t + 2] + c]) * (1.0f / 3.0f);
                const float u = (m - lo[c]) / ext * 1023.0f;
                q[c] = (uint32_t)std::min(1023.0f, std::max(0.0f, u));
            }
            code[t] = spread10(q[0]) | (spread10(q[1]) << 1) | (spread10(q[2]) << 2);
        }
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return code[a] < code[b]; });
    }
    // greedy cut: close a meshlet when it holds tris_per_meshlet triangles or a triangle would push it past 1024 vertices
    std::vector<int> local(V, -1);
    std::vector<int> cur_verts;
    int M = 0, refs = 0, mx = 0, ntri = 0;
    auto close = [&]() {
        if (ml_vptr) {
            for (int k = ntri; k < tris_per_meshlet; k++) {
                ml_tri2[2 * ((size_t)M * tris_per_meshlet + k)] = 0xffffffffu;
                ml_tri2[2 * ((size_t)M * tris_per_meshlet + k) + 1] = 0xffffffffu;
            }
            ml_vptr[M + 1] = refs;
        }
        mx = std::max(mx, (int)cur_verts.size());
        for (int v : cur_verts) local[v] = -1;
        cur_verts.clear();
        ntri = 0;
        M++;
    };
    if (ml_vptr) ml_vptr[0] = 0;
    for (int k = 0; k < T; k++) {
        const int t = order[k];
        int fresh = 0;
        for (int c = 0; c < 3; c++) {
            const int v = tri[3 * t + c];
            bool seen = local[v] >= 0;
            for (int d = 0; d < c; d++) seen = seen || tri[3 * t + d] == v;
            fresh += seen ? 0 : 1;
        }
        if (ntri == tris_per_meshlet || (int)cur_verts.size() + fresh > 1024) close();
        uint32_t l[3];
        for (int c = 0; c < 3; c++) {
            const int v = tri[3 * t + c];
            if (local[v] < 0) {
                local[v] = (int)cur_verts.size();
                cur_verts.push_back(v);
                if (ml_vptr) ml_verts[refs] = v;
                refs++;
            }
            l[c] = (uint32_t)local[v];
        }
        if (ml_vptr) {
            ml_tri2[2 * ((size_t)M * tris_per_meshlet + ntri)] = l[0] | (l[1] << 10) | (l[2] << 20);
            ml_tri2[2 * ((size_t)M * tris_per_meshlet + ntri) + 1] = (uint32_t)t;
        }
        ntri++;
    }
    if (ntri > 0) close();
    *n_meshlets = M;
    *n_vert_refs = refs;
    *max_verts = mx;
    return FMHR_OK;
}
