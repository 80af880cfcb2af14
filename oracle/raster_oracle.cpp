// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product; nothing under fmhr_b200/ may
// import, link or execute this file.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs use it, as the checker / timed CPU baseline.
//
// Plain C++ restatement of the three nvdiffrast ops the HAM loop calls
//   dr.rasterize  (reference call sites mesh_sfs_optim.py:142,212,267)
//   dr.interpolate(mesh_sfs_optim.py:143,214,269)
//   dr.antialias  (mesh_sfs_optim.py:146-147,217-219,274,287)
// nvdiffrast is an un-vendored, un-pinned third-party dependency of the reference
// (requirements.txt:14, commented out) and is absent from /root/reference, so its published
// algorithm is restated here from SURVEY.md Appendix A.  PARITY UNPINNED for these three ops:
// the reference ships no golden vectors at this boundary and upstream's OpenGL coverage is
// hardware-defined; the deterministic rule below is the spec the CUDA kernels must match
// bit-exactly (triangle id / coverage) or to fp32 tolerance (everything else).
//
// Build: g++ -O2 -ffp-contract=off -fopenmp -shared -fPIC  (see oracle/Makefile).
// -ffp-contract=off matters: every float op below is a separately rounded IEEE op, which is
// what the CUDA side reproduces with __fmul_rn/__fadd_rn/__fdiv_rn.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#include <map>

namespace {

const float kGuard = 16384.0f;  // guard band, pixels

inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline uint32_t depth_key(float zw) {  // order-preserving float -> uint32
    uint32_t b = f2u(zw);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
inline float clamp01(float x) { return (x > 0.0f) ? ((x < 1.0f) ? x : 1.0f) : 0.0f; }   // NaN -> 0
inline float clampz(float x) { return (x < 1.0f) ? ((x > -1.0f) ? x : -1.0f) : 1.0f; }  // NaN -> 1
inline int64_t floordiv256(int64_t a) { return a >> 8; }  // arithmetic shift == floor for /256

struct Bary { float u, v, zw; };

// Perspective-correct barycentrics + depth of pixel (px,py) w.r.t. clip-space triangle p0,p1,p2.
// Fixed operation order, no contraction.
inline Bary bary_at(const float* p0, const float* p1, const float* p2, int px, int py, int W, int H) {
    float invW = 1.0f / (float)W, invH = 1.0f / (float)H;
    float fx = (float)(2 * px + 1) * invW - 1.0f;
    float fy = (float)(2 * py + 1) * invH - 1.0f;
    float q0x = p0[0] - fx * p0[3], q0y = p0[1] - fy * p0[3];
    float q1x = p1[0] - fx * p1[3], q1y = p1[1] - fy * p1[3];
    float q2x = p2[0] - fx * p2[3], q2y = p2[1] - fy * p2[3];
    float a0 = q1x * q2y - q1y * q2x;
    float a1 = q2x * q0y - q2y * q0x;
    float a2 = q0x * q1y - q0y * q1x;
    float at = (a0 + a1) + a2;
    float iw = 1.0f / at;
    Bary b;
    b.u = clamp01(a0 * iw);
    b.v = clamp01(a1 * iw);
    float zn = (p0[2] * a0 + p1[2] * a1) + p2[2] * a2;
    float wd = (p0[3] * a0 + p1[3] * a1) + p2[3] * a2;
    b.zw = clampz(zn / wd) + 0.0f;  // +0 canonicalises -0
    return b;
}

inline bool owns(int64_t dx, int64_t dy) { return dy > 0 || (dy == 0 && dx > 0); }

struct Snap { int X[3], Y[3]; bool ok; };

inline Snap snap_tri(const float* p0, const float* p1, const float* p2, int W, int H) {
    Snap s; s.ok = false;
    const float* p[3] = {p0, p1, p2};
    float hw = (float)W * 0.5f, hh = (float)H * 0.5f;
    for (int k = 0; k < 3; k++) {
        float x = p[k][0], y = p[k][1], z = p[k][2], w = p[k][3];
        if (!(w > 0.0f)) return s;
        if (!(z >= -w && z <= w)) return s;
        float sx = (x / w) * hw + hw;
        float sy = (y / w) * hh + hh;
        if (!(std::fabs(sx) <= kGuard) || !(std::fabs(sy) <= kGuard)) return s;
        s.X[k] = (int)std::nearbyintf(sx * 256.0f);
        s.Y[k] = (int)std::nearbyintf(sy * 256.0f);
    }
    s.ok = true;
    return s;
}

}  // namespace

extern "C" {

// rast[N,H,W,4] = (u, v, z/w, tri_id+1); rast_db[N,H,W,4] = (du/dX, du/dY, dv/dX, dv/dY) or NULL.
// keys (optional, [N,H,W] uint64) receives the raw depth|tri z-buffer.
int orc_rasterize_fwd(const float* pos, const int32_t* tri, int N, int V, int T, int H, int W,
                      float* rast, float* rast_db, uint64_t* keys_out) {
    const uint64_t EMPTY = ~0ull;
    #pragma omp parallel for schedule(dynamic, 1)
    for (int n = 0; n < N; n++) {
        std::vector<uint64_t> zb((size_t)H * W, EMPTY);
        const float* P = pos + (size_t)n * V * 4;
        for (int t = 0; t < T; t++) {
            int i0 = tri[3 * t], i1 = tri[3 * t + 1], i2 = tri[3 * t + 2];
            if (i0 < 0 || i0 >= V || i1 < 0 || i1 >= V || i2 < 0 || i2 >= V) continue;
            const float *p0 = P + 4 * i0, *p1 = P + 4 * i1, *p2 = P + 4 * i2;
            Snap s = snap_tri(p0, p1, p2, W, H);
            if (!s.ok) continue;
            int64_t X0 = s.X[0], Y0 = s.Y[0], X1 = s.X[1], Y1 = s.Y[1], X2 = s.X[2], Y2 = s.Y[2];
            int64_t area2 = (X1 - X0) * (Y2 - Y0) - (X2 - X0) * (Y1 - Y0);
            if (area2 == 0) continue;
            if (area2 < 0) { std::swap(X1, X2); std::swap(Y1, Y2); }
            int64_t minX = std::min(X0, std::min(X1, X2)), maxX = std::max(X0, std::max(X1, X2));
            int64_t minY = std::min(Y0, std::min(Y1, Y2)), maxY = std::max(Y0, std::max(Y1, Y2));
            int64_t px0 = std::max<int64_t>(0, floordiv256(minX - 128 + 255));
            int64_t px1 = std::min<int64_t>(W - 1, floordiv256(maxX - 128));
            int64_t py0 = std::max<int64_t>(0, floordiv256(minY - 128 + 255));
            int64_t py1 = std::min<int64_t>(H - 1, floordiv256(maxY - 128));
            for (int64_t py = py0; py <= py1; py++)
                for (int64_t px = px0; px <= px1; px++) {
                    int64_t Cx = px * 256 + 128, Cy = py * 256 + 128;
                    int64_t e0 = (X2 - X1) * (Cy - Y1) - (Y2 - Y1) * (Cx - X1);
                    int64_t e1 = (X0 - X2) * (Cy - Y2) - (Y0 - Y2) * (Cx - X2);
                    int64_t e2 = (X1 - X0) * (Cy - Y0) - (Y1 - Y0) * (Cx - X0);
                    bool in0 = e0 > 0 || (e0 == 0 && owns(X2 - X1, Y2 - Y1));
                    bool in1 = e1 > 0 || (e1 == 0 && owns(X0 - X2, Y0 - Y2));
                    bool in2 = e2 > 0 || (e2 == 0 && owns(X1 - X0, Y1 - Y0));
                    if (!(in0 && in1 && in2)) continue;
                    Bary b = bary_at(p0, p1, p2, (int)px, (int)py, W, H);
                    uint64_t key = ((uint64_t)depth_key(b.zw) << 32) | (uint32_t)t;
                    uint64_t& slot = zb[(size_t)py * W + px];
                    if (key < slot) slot = key;
                }
        }
        // resolve
        for (int py = 0; py < H; py++)
            for (int px = 0; px < W; px++) {
                size_t pix = ((size_t)n * H + py) * W + px;
                uint64_t key = zb[(size_t)py * W + px];
                if (keys_out) keys_out[pix] = key;
                float* o = rast + pix * 4;
                float* d = rast_db ? rast_db + pix * 4 : nullptr;
                if (key == EMPTY) {
                    o[0] = o[1] = o[2] = o[3] = 0.0f;
                    if (d) d[0] = d[1] = d[2] = d[3] = 0.0f;
                    continue;
                }
                int t = (int)(uint32_t)key;
                const float *p0 = P + 4 * tri[3 * t], *p1 = P + 4 * tri[3 * t + 1], *p2 = P + 4 * tri[3 * t + 2];
                Bary b = bary_at(p0, p1, p2, px, py, W, H);
                o[0] = b.u; o[1] = b.v; o[2] = b.zw; o[3] = (float)(t + 1);
                if (d) {
                    // analytic screen-space derivatives of the (unclamped) barycentrics
                    double fx = (2.0 * px + 1.0) / W - 1.0, fy = (2.0 * py + 1.0) / H - 1.0;
                    double q[3][2], w[3] = {p0[3], p1[3], p2[3]};
                    const float* pp[3] = {p0, p1, p2};
                    for (int k = 0; k < 3; k++) { q[k][0] = pp[k][0] - fx * w[k]; q[k][1] = pp[k][1] - fy * w[k]; }
                    double a0 = q[1][0] * q[2][1] - q[1][1] * q[2][0];
                    double a1 = q[2][0] * q[0][1] - q[2][1] * q[0][0];
                    double a2 = q[0][0] * q[1][1] - q[0][1] * q[1][0];
                    double at = a0 + a1 + a2, iw = 1.0 / at;
                    // d/dfx: dq_k = (-w_k, 0); d/dfy: dq_k = (0, -w_k)
                    double da0x = -w[1] * q[2][1] + q[1][1] * w[2], da0y = -q[1][0] * w[2] + w[1] * q[2][0];
                    double da1x = -w[2] * q[0][1] + q[2][1] * w[0], da1y = -q[2][0] * w[0] + w[2] * q[0][0];
                    double da2x = -w[0] * q[1][1] + q[0][1] * w[1], da2y = -q[0][0] * w[1] + w[0] * q[1][0];
                    double datx = da0x + da1x + da2x, daty = da0y + da1y + da2y;
                    double u = a0 * iw, v = a1 * iw;
                    double sx = 2.0 / W, sy = 2.0 / H;
                    d[0] = (float)((da0x - u * datx) * iw * sx);
                    d[1] = (float)((da0y - u * daty) * iw * sy);
                    d[2] = (float)((da1x - v * datx) * iw * sx);
                    d[3] = (float)((da1y - v * daty) * iw * sy);
                }
            }
    }
    return 0;
}

// grad_pos[N,V,4] += d(u,v)/d(pos) . dy[...,0:2]   (SURVEY.md Appendix A "rasterize bwd").
// grad_pos must be zero-initialised by the caller.
int orc_rasterize_bwd(const float* pos, const int32_t* tri, const float* rast, const float* dy,
                      int N, int V, int T, int H, int W, float* grad_pos) {
    #pragma omp parallel for schedule(dynamic, 1)
    for (int n = 0; n < N; n++) {
        const float* P = pos + (size_t)n * V * 4;
        float* G = grad_pos + (size_t)n * V * 4;
        for (int py = 0; py < H; py++)
            for (int px = 0; px < W; px++) {
                size_t pix = ((size_t)n * H + py) * W + px;
                int t = (int)rast[pix * 4 + 3] - 1;
                if (t < 0 || t >= T) continue;
                float dydu = dy[pix * 4 + 0], dydv = dy[pix * 4 + 1];
                if (dydu == 0.0f && dydv == 0.0f) continue;
                int vi[3] = {tri[3 * t], tri[3 * t + 1], tri[3 * t + 2]};
                const float *p0 = P + 4 * vi[0], *p1 = P + 4 * vi[1], *p2 = P + 4 * vi[2];
                float fx = (float)(2 * px + 1) / (float)W - 1.0f;
                float fy = (float)(2 * py + 1) / (float)H - 1.0f;
                float q0x = p0[0] - fx * p0[3], q0y = p0[1] - fy * p0[3];
                float q1x = p1[0] - fx * p1[3], q1y = p1[1] - fy * p1[3];
                float q2x = p2[0] - fx * p2[3], q2y = p2[1] - fy * p2[3];
                float a0 = q1x * q2y - q1y * q2x;
                float a1 = q2x * q0y - q2y * q0x;
                float a2 = q0x * q1y - q0y * q1x;
                float at = a0 + a1 + a2;
                float iw = 1.0f / (at + std::copysign(1e-6f, at));
                float b0 = a0 * iw, b1 = a1 * iw;
                float gb0 = dydu * iw, gb1 = dydv * iw, gbb = gb0 * b0 + gb1 * b1;
                // d(a0,a1,at)/dq chain; see Appendix A
                float g0x = gbb * (q2y - q1y) - gb1 * q2y;
                float g1x = gbb * (q0y - q2y) + gb0 * q2y;
                float g2x = gbb * (q1y - q0y) - gb0 * q1y + gb1 * q0y;
                float g0y = gbb * (q1x - q2x) + gb1 * q2x;
                float g1y = gbb * (q2x - q0x) - gb0 * q2x;
                float g2y = gbb * (q0x - q1x) + gb0 * q1x - gb1 * q0x;
                float gx[3] = {g0x, g1x, g2x}, gy[3] = {g0y, g1y, g2y};
                for (int k = 0; k < 3; k++) {
                    float* g = G + 4 * vi[k];
                    g[0] += gx[k];
                    g[1] += gy[k];
                    g[3] += -fx * gx[k] - fy * gy[k];
                }
            }
    }
    return 0;
}

// out[N,H,W,A] = u a0 + v a1 + (1-u-v) a2; attr is [NA,V,A] with NA in {1,N} (NA==1 broadcasts).
int orc_interpolate_fwd(const float* attr, const float* rast, const int32_t* tri, int N, int NA, int V, int T,
                        int H, int W, int A, float* out) {
    #pragma omp parallel for schedule(static)
    for (int n = 0; n < N; n++) {
        const float* At = attr + (NA == 1 ? 0 : (size_t)n * V * A);
        for (size_t p = 0; p < (size_t)H * W; p++) {
            size_t pix = (size_t)n * H * W + p;
            float* o = out + pix * A;
            int t = (int)rast[pix * 4 + 3] - 1;
            if (t < 0 || t >= T) { for (int k = 0; k < A; k++) o[k] = 0.0f; continue; }
            float u = rast[pix * 4], v = rast[pix * 4 + 1], w = 1.0f - u - v;
            const float *a0 = At + (size_t)tri[3 * t] * A, *a1 = At + (size_t)tri[3 * t + 1] * A, *a2 = At + (size_t)tri[3 * t + 2] * A;
            for (int k = 0; k < A; k++) o[k] = u * a0[k] + v * a1[k] + w * a2[k];
        }
    }
    return 0;
}

// grad_attr[NA,V,A] (zero-initialised), grad_rast[N,H,W,4] (fully written).
int orc_interpolate_bwd(const float* attr, const float* rast, const int32_t* tri, const float* dy, int N, int NA,
                        int V, int T, int H, int W, int A, float* grad_attr, float* grad_rast) {
    for (int n = 0; n < N; n++) {  // sequential: NA==1 makes views alias the same accumulator
        const float* At = attr + (NA == 1 ? 0 : (size_t)n * V * A);
        float* Gt = grad_attr + (NA == 1 ? 0 : (size_t)n * V * A);
        for (size_t p = 0; p < (size_t)H * W; p++) {
            size_t pix = (size_t)n * H * W + p;
            float* gr = grad_rast + pix * 4;
            gr[0] = gr[1] = gr[2] = gr[3] = 0.0f;
            int t = (int)rast[pix * 4 + 3] - 1;
            if (t < 0 || t >= T) continue;
            float u = rast[pix * 4], v = rast[pix * 4 + 1], w = 1.0f - u - v;
            int i0 = tri[3 * t], i1 = tri[3 * t + 1], i2 = tri[3 * t + 2];
            const float* g = dy + pix * A;
            float du = 0.0f, dv = 0.0f;
            for (int k = 0; k < A; k++) {
                float gk = g[k];
                Gt[(size_t)i0 * A + k] += u * gk;
                Gt[(size_t)i1 * A + k] += v * gk;
                Gt[(size_t)i2 * A + k] += w * gk;
                du += gk * (At[(size_t)i0 * A + k] - At[(size_t)i2 * A + k]);
                dv += gk * (At[(size_t)i1 * A + k] - At[(size_t)i2 * A + k]);
            }
            gr[0] = du; gr[1] = dv;
        }
    }
    return 0;
}

// opp[T,3]: for triangle t and corner k, the vertex opposite to the edge facing corner k (edge between the other
// two corners) in the lowest-indexed OTHER triangle sharing that edge; -1 if the edge is a boundary.
int orc_antialias_topology(const int32_t* tri, int T, int32_t* opp) {
    std::map<std::pair<int, int>, std::vector<std::pair<int, int>>> edges;  // edge -> [(tri, opposite vertex)]
    for (int t = 0; t < T; t++)
        for (int k = 0; k < 3; k++) {
            int a = tri[3 * t + (k + 1) % 3], b = tri[3 * t + (k + 2) % 3], o = tri[3 * t + k];
            edges[{std::min(a, b), std::max(a, b)}].push_back({t, o});
        }
    for (int t = 0; t < T; t++)
        for (int k = 0; k < 3; k++) {
            int a = tri[3 * t + (k + 1) % 3], b = tri[3 * t + (k + 2) % 3];
            const auto& lst = edges[{std::min(a, b), std::max(a, b)}];
            int r = -1;
            for (const auto& e : lst) if (e.first != t) { r = e.second; break; }  // lst is in ascending tri order
            opp[3 * t + k] = r;
        }
    return 0;
}

namespace {
inline bool same_sign(float a, float b) { return (int32_t)(f2u(a) ^ f2u(b)) >= 0; }
// n0/d0 > n1/d1 without dividing (denominators non-zero)
inline bool rational_gt(float n0, float d0, float n1, float d1) {
    float l = n0 * d1, r = n1 * d0;
    bool flip = (d0 < 0.0f) != (d1 < 0.0f);
    return flip ? (l < r) : (l > r);
}
struct AAItem { int pix0, pix1, tri, di, d, from1; float alpha; int clamped; };

// Analyse one pixel pair; returns true if a silhouette edge crossing was found (alpha written).
inline bool aa_analyse(const float* rast, const float* pos, const int32_t* tri, const int32_t* opp,
                       int n, int px, int py, int d, int H, int W, int V, int T, AAItem& it) {
    size_t pix0 = ((size_t)n * H + py) * W + px;
    size_t pix1 = pix0 + (d ? W : 1);
    float z0 = rast[pix0 * 4 + 2], z1 = rast[pix1 * 4 + 2];
    int tri0 = (int)rast[pix0 * 4 + 3] - 1, tri1 = (int)rast[pix1 * 4 + 3] - 1;
    int t = (tri0 >= 0) ? tri0 : tri1;
    if (tri0 >= 0 && tri1 >= 0) t = (z0 < z1) ? tri0 : tri1;
    bool from1 = (t == tri1);
    if (from1) { px += 1 - d; py += d; }
    if (t < 0 || t >= T) return false;
    int vi[3] = {tri[3 * t], tri[3 * t + 1], tri[3 * t + 2]};
    for (int k = 0; k < 3; k++) if (vi[k] < 0 || vi[k] >= V) return false;
    const float* P = pos + (size_t)n * V * 4;
    float xh = 0.5f * (float)W, yh = 0.5f * (float)H;
    float fx = (float)px + 0.5f - xh, fy = (float)py + 0.5f - yh;
    float x[3], y[3], ox[3], oy[3];
    for (int k = 0; k < 3; k++) {
        const float* p = P + 4 * vi[k];
        float iw = 1.0f / p[3];
        x[k] = p[0] * iw * xh - fx;
        y[k] = p[1] * iw * yh - fy;
    }
    for (int k = 0; k < 3; k++) {
        int o = opp[3 * t + k];
        if (o < 0 || o >= V) { ox[k] = x[k]; oy[k] = y[k]; continue; }
        const float* p = P + 4 * o;
        float iw = 1.0f / p[3];
        ox[k] = p[0] * iw * xh - fx;
        oy[k] = p[1] * iw * yh - fy;
    }
    float bb = (x[1] - x[0]) * (y[2] - y[0]) - (x[2] - x[0]) * (y[1] - y[0]);
    float a0 = (x[1] - ox[0]) * (y[2] - oy[0]) - (x[2] - ox[0]) * (y[1] - oy[0]);
    float a1 = (x[2] - ox[1]) * (y[0] - oy[1]) - (x[0] - ox[1]) * (y[2] - oy[1]);
    float a2 = (x[0] - ox[2]) * (y[1] - oy[2]) - (x[1] - ox[2]) * (y[0] - oy[2]);
    bool s0 = same_sign(a0, bb), s1 = same_sign(a1, bb), s2 = same_sign(a2, bb);
    if (!(s0 || s1 || s2)) return false;
    if (d) for (int k = 0; k < 3; k++) std::swap(x[k], y[k]);
    float dx0 = x[2] - x[1], dx1 = x[0] - x[2], dx2 = x[1] - x[0];
    float dy0 = y[2] - y[1], dy1 = y[0] - y[2], dy2 = y[1] - y[0];
    float ds = from1 ? -1.0f : 1.0f;
    const float NEG = -3.402823466e38f;
    float d0 = ds * (x[1] * dy0 - y[1] * dx0);
    float d1 = ds * (x[2] * dy1 - y[2] * dx1);
    float d2 = ds * (x[0] * dy2 - y[0] * dx2);
    if (same_sign(y[1], y[2])) { d0 = NEG; dy0 = 1.0f; }
    if (same_sign(y[2], y[0])) { d1 = NEG; dy1 = 1.0f; }
    if (same_sign(y[0], y[1])) { d2 = NEG; dy2 = 1.0f; }
    bool g10 = rational_gt(d1, dy1, d0, dy0);
    bool g20 = rational_gt(d2, dy2, d0, dy0);
    bool g21 = rational_gt(d2, dy2, d1, dy1);
    int di = (g20 && g21) ? 2 : (g10 ? 1 : 0);
    float dc = NEG;
    if (di == 0 && s0 && std::fabs(dy0) >= std::fabs(dx0)) dc = d0 / dy0;
    if (di == 1 && s1 && std::fabs(dy1) >= std::fabs(dx1)) dc = d1 / dy1;
    if (di == 2 && s2 && std::fabs(dy2) >= std::fabs(dx2)) dc = d2 / dy2;
    const float eps = 0.0625f;
    if (!(dc > -eps && dc < 1.0f + eps)) return false;
    int clamped = !(dc > 0.0f && dc < 1.0f);
    dc = std::min(std::max(dc, 0.0f), 1.0f);
    it.pix0 = (int)pix0; it.pix1 = (int)pix1; it.tri = t; it.di = di; it.d = d; it.from1 = from1 ? 1 : 0;
    it.alpha = ds * (0.5f - dc); it.clamped = clamped;
    return true;
}
}  // namespace

// out[N,H,W,C] = color + silhouette blends.  items (optional) receives up to max_items records of 8 int32:
// (pix0, pix1, tri, di, d, from1, alpha_bits, clamped); *n_items gets the count found.
int orc_antialias_fwd(const float* color, const float* rast, const float* pos, const int32_t* tri, const int32_t* opp,
                      int N, int H, int W, int C, int V, int T, float* out, int32_t* items, int max_items, int* n_items) {
    std::memcpy(out, color, sizeof(float) * (size_t)N * H * W * C);
    int cnt = 0;
    for (int n = 0; n < N; n++)
        for (int py = 0; py < H; py++)
            for (int px = 0; px < W; px++)
                for (int d = 0; d < 2; d++) {
                    if (d == 0 && px + 1 >= W) continue;
                    if (d == 1 && py + 1 >= H) continue;
                    size_t pix0 = ((size_t)n * H + py) * W + px, pix1 = pix0 + (d ? W : 1);
                    if (rast[pix0 * 4 + 3] == rast[pix1 * 4 + 3]) continue;
                    AAItem it;
                    if (!aa_analyse(rast, pos, tri, opp, n, px, py, d, H, W, V, T, it)) continue;
                    const float *c0 = color + (size_t)it.pix0 * C, *c1 = color + (size_t)it.pix1 * C;
                    float* o = out + (size_t)(it.alpha > 0.0f ? it.pix0 : it.pix1) * C;
                    for (int c = 0; c < C; c++) o[c] += it.alpha * (c1[c] - c0[c]);
                    if (items && cnt < max_items) {
                        int32_t* r = items + 8 * (size_t)cnt;
                        r[0] = it.pix0; r[1] = it.pix1; r[2] = it.tri; r[3] = it.di; r[4] = it.d; r[5] = it.from1;
                        r[6] = (int32_t)f2u(it.alpha); r[7] = it.clamped;
                    }
                    cnt++;
                }
    if (n_items) *n_items = cnt;
    return 0;
}

// grad_color[N,H,W,C] (fully written), grad_pos[N,V,4] (zero-initialised by caller).
int orc_antialias_bwd(const float* color, const float* rast, const float* pos, const int32_t* tri, const int32_t* opp,
                      const float* dy, int N, int H, int W, int C, int V, int T, float* grad_color, float* grad_pos) {
    std::memcpy(grad_color, dy, sizeof(float) * (size_t)N * H * W * C);
    for (int n = 0; n < N; n++)
        for (int py = 0; py < H; py++)
            for (int px = 0; px < W; px++)
                for (int d = 0; d < 2; d++) {
                    if (d == 0 && px + 1 >= W) continue;
                    if (d == 1 && py + 1 >= H) continue;
                    size_t pix0 = ((size_t)n * H + py) * W + px, pix1 = pix0 + (d ? W : 1);
                    if (rast[pix0 * 4 + 3] == rast[pix1 * 4 + 3]) continue;
                    AAItem it;
                    if (!aa_analyse(rast, pos, tri, opp, n, px, py, d, H, W, V, T, it)) continue;
                    const float *c0 = color + (size_t)it.pix0 * C, *c1 = color + (size_t)it.pix1 * C;
                    const float* g = dy + (size_t)(it.alpha > 0.0f ? it.pix0 : it.pix1) * C;
                    float dd = 0.0f;
                    for (int c = 0; c < C; c++) {
                        dd += g[c] * (c1[c] - c0[c]);
                        grad_color[(size_t)it.pix0 * C + c] -= it.alpha * g[c];
                        grad_color[(size_t)it.pix1 * C + c] += it.alpha * g[c];
                    }
                    if (dd == 0.0f || it.clamped || !grad_pos) continue;
                    // alpha = ds/2 - (x1*dy - y1*dx)/dy in the (possibly x/y-swapped) frame of edge (v1,v2)
                    int t = it.tri;
                    int i1 = tri[3 * t + (it.di + 1) % 3], i2 = tri[3 * t + (it.di + 2) % 3];
                    int qx = px + (it.from1 ? 1 - d : 0), qy = py + (it.from1 ? d : 0);
                    float xh = 0.5f * (float)W, yh = 0.5f * (float)H;
                    float fx = (float)qx + 0.5f - xh, fy = (float)qy + 0.5f - yh;
                    const float *p1 = pos + ((size_t)n * V + i1) * 4, *p2 = pos + ((size_t)n * V + i2) * 4;
                    float iw1 = 1.0f / p1[3], iw2 = 1.0f / p2[3];
                    float x1 = p1[0] * iw1 * xh - fx, y1 = p1[1] * iw1 * yh - fy;
                    float x2 = p2[0] * iw2 * xh - fx, y2 = p2[1] * iw2 * yh - fy;
                    if (d) { std::swap(x1, y1); std::swap(x2, y2); }
                    float ex = x2 - x1, ey = y2 - y1, iy = 1.0f / ey;
                    // partials of alpha
                    float ax1 = -y2 * iy, ax2 = y1 * iy;
                    float ay1 = ex * y2 * iy * iy, ay2 = -ex * y1 * iy * iy;
                    if (d) { std::swap(ax1, ay1); std::swap(ax2, ay2); }  // back to image x/y
                    float* g1 = grad_pos + ((size_t)n * V + i1) * 4;
                    float* g2 = grad_pos + ((size_t)n * V + i2) * 4;
                    // screen x = p.x/p.w*xh ; y = p.y/p.w*yh
                    g1[0] += dd * ax1 * xh * iw1;
                    g1[1] += dd * ay1 * yh * iw1;
                    g1[3] += -dd * (ax1 * xh * p1[0] + ay1 * yh * p1[1]) * iw1 * iw1;
                    g2[0] += dd * ax2 * xh * iw2;
                    g2[1] += dd * ay2 * yh * iw2;
                    g2[3] += -dd * (ax2 * xh * p2[0] + ay2 * yh * p2[1]) * iw2 * iw2;
                }
    return 0;
}


// Clip positions of the documented rule (rule 0): the reference evaluates
//   rot = [v,1] @ w2c ;  clip = rot @ proj                                    (mesh_sfs_optim.py:262-264)
// with two batched fp32 GEMMs whose summation order is unspecified (it differs between ATen's CPU and CUDA back ends).
// Coverage is a discontinuous function of the positions, so the GPU path and this checker share ONE fixed evaluation
// order: the per-view combined matrix M = w2c @ proj and clip = [v,1] @ M, each a chain of single-rounded fused
// multiply-adds (fmaf is exact: -ffp-contract only concerns implicit contraction).  Agrees with the einsum to 3e-7 relative.
int orc_clip_positions(const float* verts /*[V,3]*/, int V, const float* w2cs /*[n,4,4] transposed*/,
                       const float* projs /*[n,4,4] transposed*/, int n, float* out /*[n,V,4]*/) {
    for (int k = 0; k < n; k++) {
        const float* Wm = w2cs + 16 * (size_t)k;
        const float* Pm = projs + 16 * (size_t)k;
        float M[16];
        for (int r = 0; r < 4; r++)
            for (int j = 0; j < 4; j++)
                M[4 * r + j] = fmaf(Wm[4 * r + 3], Pm[12 + j], fmaf(Wm[4 * r + 2], Pm[8 + j],
                                    fmaf(Wm[4 * r + 1], Pm[4 + j], Wm[4 * r] * Pm[j])));
#pragma omp parallel for
        for (int i = 0; i < V; i++) {
            const float x = verts[3 * (size_t)i], y = verts[3 * (size_t)i + 1], z = verts[3 * (size_t)i + 2];
            float* o = out + ((size_t)k * V + i) * 4;
            for (int j = 0; j < 4; j++) o[j] = fmaf(z, M[8 + j], fmaf(y, M[4 + j], fmaf(x, M[j], M[12 + j])));
        }
    }
    return 0;
}

}  // extern "C"
