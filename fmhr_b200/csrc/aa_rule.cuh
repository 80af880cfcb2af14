// Silhouette-edge analysis of one pixel pair, shared by the stand-alone antialias op (antialias.cu) and the fused
// HAM kernels (ham.cu).  Semantics follow nvdiffrast's antialias; every decision-relevant
// operation is an exactly-rounded, un-contracted fp32 op in the same order as the CPU checker (aa_analyse in the test oracle),
// so the discrete choices (which triangle, which edge, blend or not) are bit-identical to the oracle's.
#pragma once
#include "common.cuh"

namespace fmhr {

struct AAPair {
    int tri;      // chosen triangle
    int di;       // chosen edge: runs from corner (di+1)%3 to corner (di+2)%3
    int from1;    // 1 if the chosen triangle belongs to the second pixel of the pair
    int clamped;  // crossing parameter was clamped -> no position gradient
    int i1, i2;   // vertex ids of the chosen edge
    float alpha;  // blend weight; receiver is pixel0 if alpha > 0 else pixel1
};

__device__ __forceinline__ bool aa_same_sign(float a, float b) {
    return (int)(__float_as_uint(a) ^ __float_as_uint(b)) >= 0;
}
__device__ __forceinline__ bool aa_rational_gt(float n0, float d0, float n1, float d1) {
    const float l = xm(n0, d1), r = xm(n1, d0);
    const bool flip = (d0 < 0.0f) != (d1 < 0.0f);
    return flip ? (l < r) : (l > r);
}
// Window coordinates of a vertex relative to a pixel centre.  Two sources with bit-identical results:
//  AAProjClip   - from clip-space positions [V,4] (the stand-alone op): one IEEE divide per vertex;
//  AAProjScreen - from (x/w*W/2, y/w*H/2) pairs the fused path's transform kernel pre-computed once per (view, vertex)
//                 with the very same operations (aa_window_xy), leaving two subtractions here.
__device__ __forceinline__ float2 aa_window_xy(const float4 p, float xh, float yh) {
    const float iw = xd(1.0f, p.w);
    return make_float2(xm(xm(p.x, iw), xh), xm(xm(p.y, iw), yh));
}
struct AAProjClip {
    const float* P;
    float xh, yh;
    __device__ __forceinline__ void operator()(int v, float fx, float fy, float& x, float& y) const {
        const float2 s = aa_window_xy(ldg4(P + 4 * (size_t)v), xh, yh);
        x = xs(s.x, fx);
        y = xs(s.y, fy);
    }
};
struct AAProjScreen {
    const float2* S;
    __device__ __forceinline__ void operator()(int v, float fx, float fy, float& x, float& y) const {
        const float2 s = __ldg(S + v);
        x = xs(s.x, fx);
        y = xs(s.y, fy);
    }
};

// Projected corners of triangle t relative to the centre of its OWNING pixel (qx,qy), plus the three silhouette-candidate
// bits: bit k set <=> the wing across the edge facing corner k lies on the same side as the triangle itself (or the
// edge is a mesh boundary).  bits == 0 means no edge of this triangle can ever blend, whatever the neighbour.
struct AAGeom {
    int v0, v1, v2;
    float x0, y0, x1, y1, x2, y2;
    int bits;
};

// Core of aa_triangle_geom for callers that already hold the triangle's corners (v*) and opposite-wing vertices (o*).
template <class Proj>
__device__ __forceinline__ bool aa_triangle_geom_idx(int v0, int v1, int v2, int o0, int o1, int o2, int qx, int qy,
                                                     const Proj proj, int V, int H, int W, AAGeom& g) {
    if ((unsigned)v0 >= (unsigned)V || (unsigned)v1 >= (unsigned)V || (unsigned)v2 >= (unsigned)V) return false;
    const float xh = 0.5f * (float)W, yh = 0.5f * (float)H;
    const float fx = xs(xa((float)qx, 0.5f), xh), fy = xs(xa((float)qy, 0.5f), yh);
    float x0, y0, x1, y1, x2, y2;
    proj(v0, fx, fy, x0, y0);
    proj(v1, fx, fy, x1, y1);
    proj(v2, fx, fy, x2, y2);
    float ox0 = x0, oy0 = y0, ox1 = x1, oy1 = y1, ox2 = x2, oy2 = y2;
    if ((unsigned)o0 < (unsigned)V) proj(o0, fx, fy, ox0, oy0);
    if ((unsigned)o1 < (unsigned)V) proj(o1, fx, fy, ox1, oy1);
    if ((unsigned)o2 < (unsigned)V) proj(o2, fx, fy, ox2, oy2);
    const float bb = xs(xm(xs(x1, x0), xs(y2, y0)), xm(xs(x2, x0), xs(y1, y0)));
    const float a0 = xs(xm(xs(x1, ox0), xs(y2, oy0)), xm(xs(x2, ox0), xs(y1, oy0)));
    const float a1 = xs(xm(xs(x2, ox1), xs(y0, oy1)), xm(xs(x0, ox1), xs(y2, oy1)));
    const float a2 = xs(xm(xs(x0, ox2), xs(y1, oy2)), xm(xs(x1, ox2), xs(y0, oy2)));
    g.v0 = v0; g.v1 = v1; g.v2 = v2;
    g.x0 = x0; g.y0 = y0; g.x1 = x1; g.y1 = y1; g.x2 = x2; g.y2 = y2;
    g.bits = (aa_same_sign(a0, bb) ? 1 : 0) | (aa_same_sign(a1, bb) ? 2 : 0) | (aa_same_sign(a2, bb) ? 4 : 0);
    return true;
}

// Same, from absolute window coordinates (aa_window_xy) of the corners (s*) and of the opposite-wing vertices (so*,
// ignored where the wing does not exist): identical arithmetic to AAProjScreen.
__device__ __forceinline__ void aa_triangle_geom_win(int v0, int v1, int v2, float2 s0, float2 s1, float2 s2, int o0, int o1,
                                                     int o2, float2 so0, float2 so1, float2 so2, int qx, int qy, int V,
                                                     int H, int W, AAGeom& g) {
    const float xh = 0.5f * (float)W, yh = 0.5f * (float)H;
    const float fx = xs(xa((float)qx, 0.5f), xh), fy = xs(xa((float)qy, 0.5f), yh);
    const float x0 = xs(s0.x, fx), y0 = xs(s0.y, fy), x1 = xs(s1.x, fx), y1 = xs(s1.y, fy), x2 = xs(s2.x, fx), y2 = xs(s2.y, fy);
    float ox0 = x0, oy0 = y0, ox1 = x1, oy1 = y1, ox2 = x2, oy2 = y2;
    if ((unsigned)o0 < (unsigned)V) { ox0 = xs(so0.x, fx); oy0 = xs(so0.y, fy); }
    if ((unsigned)o1 < (unsigned)V) { ox1 = xs(so1.x, fx); oy1 = xs(so1.y, fy); }
    if ((unsigned)o2 < (unsigned)V) { ox2 = xs(so2.x, fx); oy2 = xs(so2.y, fy); }
    const float bb = xs(xm(xs(x1, x0), xs(y2, y0)), xm(xs(x2, x0), xs(y1, y0)));
    const float a0 = xs(xm(xs(x1, ox0), xs(y2, oy0)), xm(xs(x2, ox0), xs(y1, oy0)));
    const float a1 = xs(xm(xs(x2, ox1), xs(y0, oy1)), xm(xs(x0, ox1), xs(y2, oy1)));
    const float a2 = xs(xm(xs(x0, ox2), xs(y1, oy2)), xm(xs(x1, ox2), xs(y0, oy2)));
    g.v0 = v0; g.v1 = v1; g.v2 = v2;
    g.x0 = x0; g.y0 = y0; g.x1 = x1; g.y1 = y1; g.x2 = x2; g.y2 = y2;
    g.bits = (aa_same_sign(a0, bb) ? 1 : 0) | (aa_same_sign(a1, bb) ? 2 : 0) | (aa_same_sign(a2, bb) ? 4 : 0);
}

template <class Proj>
__device__ __forceinline__ bool aa_triangle_geom(int t, int qx, int qy, const Proj proj,
                                                 const int32_t* __restrict__ tri, const int32_t* __restrict__ opp,
                                                 int V, int T, int H, int W, AAGeom& g) {
    if (t < 0 || t >= T) return false;
    const int v0 = __ldg(tri + 3 * t), v1 = __ldg(tri + 3 * t + 1), v2 = __ldg(tri + 3 * t + 2);
    const int o0 = __ldg(opp + 3 * t), o1 = __ldg(opp + 3 * t + 1), o2 = __ldg(opp + 3 * t + 2);
    return aa_triangle_geom_idx(v0, v1, v2, o0, o1, o2, qx, qy, proj, V, H, W, g);
}

// Edge selection for the pair whose chosen triangle has geometry g (d = 0 right / 1 down neighbour, from1 = the chosen
// triangle belongs to the pair's second pixel).
__device__ __forceinline__ bool aa_select_edge(const AAGeom& g, int t, int d, bool from1, AAPair& out) {
    if (g.bits == 0) return false;
    const bool s0 = g.bits & 1, s1 = g.bits & 2, s2 = g.bits & 4;
    float x0 = g.x0, y0 = g.y0, x1 = g.x1, y1 = g.y1, x2 = g.x2, y2 = g.y2;
    if (d) {
        float tmp;
        tmp = x0; x0 = y0; y0 = tmp;
        tmp = x1; x1 = y1; y1 = tmp;
        tmp = x2; x2 = y2; y2 = tmp;
    }
    float dx0 = xs(x2, x1), dx1 = xs(x0, x2), dx2 = xs(x1, x0);
    float dy0 = xs(y2, y1), dy1 = xs(y0, y2), dy2 = xs(y1, y0);
    const float ds = from1 ? -1.0f : 1.0f;
    const float NEG = -3.402823466e38f;
    float d0 = xm(ds, xs(xm(x1, dy0), xm(y1, dx0)));
    float d1 = xm(ds, xs(xm(x2, dy1), xm(y2, dx1)));
    float d2 = xm(ds, xs(xm(x0, dy2), xm(y0, dx2)));
    if (aa_same_sign(y1, y2)) { d0 = NEG; dy0 = 1.0f; }
    if (aa_same_sign(y2, y0)) { d1 = NEG; dy1 = 1.0f; }
    if (aa_same_sign(y0, y1)) { d2 = NEG; dy2 = 1.0f; }
    const bool g10 = aa_rational_gt(d1, dy1, d0, dy0);
    const bool g20 = aa_rational_gt(d2, dy2, d0, dy0);
    const bool g21 = aa_rational_gt(d2, dy2, d1, dy1);
    const int di = (g20 && g21) ? 2 : (g10 ? 1 : 0);
    float dc = NEG;
    if (di == 0 && s0 && fabsf(dy0) >= fabsf(dx0)) dc = xd(d0, dy0);
    if (di == 1 && s1 && fabsf(dy1) >= fabsf(dx1)) dc = xd(d1, dy1);
    if (di == 2 && s2 && fabsf(dy2) >= fabsf(dx2)) dc = xd(d2, dy2);
    const float eps = 0.0625f;
    if (!(dc > -eps && dc < 1.0f + eps)) return false;
    out.clamped = !(dc > 0.0f && dc < 1.0f);
    dc = fminf(fmaxf(dc, 0.0f), 1.0f);
    out.tri = t;
    out.di = di;
    out.from1 = from1 ? 1 : 0;
    out.alpha = xm(ds, xs(0.5f, dc));
    out.i1 = (di == 0) ? g.v1 : ((di == 1) ? g.v2 : g.v0);
    out.i2 = (di == 0) ? g.v2 : ((di == 1) ? g.v0 : g.v1);
    return true;
}

// (px,py) = first pixel of the pair, d = 0 (right neighbour) or 1 (down neighbour);
// tri0/z0 and tri1/z1 are triangle id (-1 = empty) and z/w of the two pixels; proj = this view's vertex projector.
template <class Proj>
__device__ __forceinline__ bool aa_analyse(int tri0, float z0, int tri1, float z1, int px, int py, int d,
                                           const Proj proj, const int32_t* __restrict__ tri,
                                           const int32_t* __restrict__ opp, int V, int T, int H, int W, AAPair& out) {
    int t = (tri0 >= 0) ? tri0 : tri1;
    if (tri0 >= 0 && tri1 >= 0) t = (z0 < z1) ? tri0 : tri1;
    const bool from1 = (t == tri1);
    if (from1) { px += 1 - d; py += d; }
    AAGeom g;
    if (!aa_triangle_geom(t, px, py, proj, tri, opp, V, T, H, W, g)) return false;
    return aa_select_edge(g, t, d, from1, out);
}

// aa_analyse for callers that already know the chosen triangle's silhouette-candidate bits in its owning pixel's frame
// (the fused path stores them in the z-buffer key during shading): only the three corners are projected, the three
// wing vertices (three more gathers and IEEE divides) are not needed.  Decisions are identical to aa_analyse.
template <class Proj>
__device__ __forceinline__ bool aa_analyse_bits(int tri0, float z0, int bits0, int tri1, float z1, int bits1, int px, int py,
                                                int d, const Proj proj, const int32_t* __restrict__ tri, int V, int T,
                                                int H, int W, AAPair& out) {
    int t = (tri0 >= 0) ? tri0 : tri1;
    if (tri0 >= 0 && tri1 >= 0) t = (z0 < z1) ? tri0 : tri1;
    const bool from1 = (t == tri1);
    if (from1) { px += 1 - d; py += d; }
    if (t < 0 || t >= T) return false;
    AAGeom g;
    g.bits = from1 ? bits1 : bits0;
    if (g.bits == 0) return false;
    g.v0 = __ldg(tri + 3 * t); g.v1 = __ldg(tri + 3 * t + 1); g.v2 = __ldg(tri + 3 * t + 2);
    if ((unsigned)g.v0 >= (unsigned)V || (unsigned)g.v1 >= (unsigned)V || (unsigned)g.v2 >= (unsigned)V) return false;
    const float xh = 0.5f * (float)W, yh = 0.5f * (float)H;
    const float fx = xs(xa((float)px, 0.5f), xh), fy = xs(xa((float)py, 0.5f), yh);
    proj(g.v0, fx, fy, g.x0, g.y0);
    proj(g.v1, fx, fy, g.x1, g.y1);
    proj(g.v2, fx, fy, g.x2, g.y2);
    return aa_select_edge(g, t, d, from1, out);
}

// d(alpha)/d(clip position) of the two edge vertices, scaled by `dd` = d(loss)/d(alpha).
// (px,py) is the pair's FIRST pixel, p1 / p2 the clip positions of the edge's vertices pr.i1 / pr.i2.  Returns the two
// float4 gradients (x, y, 0, w).
__device__ __forceinline__ void aa_pos_grad(const AAPair& pr, int px, int py, int d, const float4 p1, const float4 p2,
                                            int H, int W, float dd, float4& g1, float4& g2) {
    const int qx = px + (pr.from1 ? 1 - d : 0), qy = py + (pr.from1 ? d : 0);
    const float xh = 0.5f * (float)W, yh = 0.5f * (float)H;
    const float fx = (float)qx + 0.5f - xh, fy = (float)qy + 0.5f - yh;
    const float iw1 = 1.0f / p1.w, iw2 = 1.0f / p2.w;
    float x1 = p1.x * iw1 * xh - fx, y1 = p1.y * iw1 * yh - fy;
    float x2 = p2.x * iw2 * xh - fx, y2 = p2.y * iw2 * yh - fy;
    if (d) {
        float tmp;
        tmp = x1; x1 = y1; y1 = tmp;
        tmp = x2; x2 = y2; y2 = tmp;
    }
    const float ex = x2 - x1, ey = y2 - y1, iy = 1.0f / ey;
    float ax1 = -y2 * iy, ax2 = y1 * iy;
    float ay1 = ex * y2 * iy * iy, ay2 = -ex * y1 * iy * iy;
    if (d) {
        float tmp;
        tmp = ax1; ax1 = ay1; ay1 = tmp;
        tmp = ax2; ax2 = ay2; ay2 = tmp;
    }
    g1 = make_float4(dd * ax1 * xh * iw1, dd * ay1 * yh * iw1, 0.0f,
                     -dd * (ax1 * xh * p1.x + ay1 * yh * p1.y) * iw1 * iw1);
    g2 = make_float4(dd * ax2 * xh * iw2, dd * ay2 * yh * iw2, 0.0f,
                     -dd * (ax2 * xh * p2.x + ay2 * yh * p2.y) * iw2 * iw2);
}

}  // namespace fmhr
