// Rasterize forward / backward (replaces dr.rasterize; reference call sites mesh_sfs_optim.py:142,212,267).
//
// Micropolygon regime: a 3x-subdivided hand at 512x334 puts ~0.4 pixel centres in the average
// triangle, so the design is one thread per (view, triangle) with an integer bounding box that is almost always
// <= 2x2 pixels, 32-bit edge functions on that fast path, and a 64-bit atomicMin z-buffer keyed
// (orderable depth << 32 | triangle id).  A second, pixel-parallel pass resolves the winners into the
// nvdiffrast-shaped rast_out plane with fully coalesced float4 stores.
#include "common.cuh"

namespace fmhr {

template <typename I>
__device__ __forceinline__ bool owns_edge(I dx, I dy) { return dy > 0 || (dy == 0 && dx > 0); }

template <typename I>
__device__ __forceinline__ void cover_bbox(int X0, int Y0, int X1, int Y1, int X2, int Y2, int px0, int px1, int py0,
                                           int py1, const float4 p0, const float4 p1, const float4 p2, int W,
                                           float invW, float invH, unsigned long long* __restrict__ zb, uint32_t t) {
    // edge k is opposite vertex k; e_k(C) = dx_k*(Cy - Ya) - dy_k*(Cx - Xa)
    const I dx0 = X2 - X1, dy0 = Y2 - Y1;
    const I dx1 = X0 - X2, dy1 = Y0 - Y2;
    const I dx2 = X1 - X0, dy2 = Y1 - Y0;
    // bias folds the tie rule into a strict comparison: inside <=> e + bias > 0
    const I b0 = owns_edge(dx0, dy0) ? 1 : 0, b1 = owns_edge(dx1, dy1) ? 1 : 0, b2 = owns_edge(dx2, dy2) ? 1 : 0;
    const int Cx0 = px0 * 256 + 128;
    for (int py = py0; py <= py1; py++) {
        const int Cy = py * 256 + 128;
        I e0 = dx0 * (I)(Cy - Y1) - dy0 * (I)(Cx0 - X1);
        I e1 = dx1 * (I)(Cy - Y2) - dy1 * (I)(Cx0 - X2);
        I e2 = dx2 * (I)(Cy - Y0) - dy2 * (I)(Cx0 - X0);
        for (int px = px0; px <= px1; px++) {
            if (e0 + b0 > 0 && e1 + b1 > 0 && e2 + b2 > 0) {
                Bary b = bary_at(p0, p1, p2, px, py, invW, invH);
                unsigned long long key = ((unsigned long long)depth_key(b.zw) << 32) | t;
                atomicMin(&zb[(size_t)py * W + px], key);
            }
            e0 -= dy0 * 256;
            e1 -= dy1 * 256;
            e2 -= dy2 * 256;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Coverage for the fused HAM path.
//  * vertices were snapped once per (view, vertex) by the transform kernel, so the per-triangle work before the
//    (usually failing) bounding-box test is three 8-byte gathers and integer min/max;
//  * the nested bounding-box loops only TEST pixel centres; every hit is pushed to a per-warp shared-memory queue and
//    the expensive part (three float4 position gathers, perspective barycentric depth with two IEEE divides, the 64-bit
//    atomicMin) then runs once per fragment with the queue spread over all 32 lanes.  With ~0.3 fragments per triangle
//    the in-loop version executed that ~120-instruction body up to four times per warp at ~10 % lane utilisation;
//  * tile hits are collected in a per-block shared-memory bitmap; at block end the tiles not yet in the slot's global
//    bitmap are appended to its work list (a few atomics per block instead of one store per fragment).
// ------------------------------------------------------------------------------------------------
constexpr int kFragQueue = 160;  // per-warp capacity (128 triangles per warp); overflowing fragments are resolved in place

__device__ __forceinline__ void resolve_fragment(const float4* __restrict__ P, const int32_t* __restrict__ tri, int t,
                                                 int px, int py, int W, float invW, float invH,
                                                 unsigned long long* __restrict__ zb) {
    const float4 p0 = __ldg(P + __ldg(tri + 3 * t)), p1 = __ldg(P + __ldg(tri + 3 * t + 1)),
                 p2 = __ldg(P + __ldg(tri + 3 * t + 2));
    const Bary b = bary_at(p0, p1, p2, px, py, invW, invH);
    atomicMin(&zb[(size_t)py * W + px], ((unsigned long long)depth_key(b.zw) << 32) | (uint32_t)t);
}

template <typename I>
__device__ __forceinline__ void cover_bbox_queue(int X0, int Y0, int X1, int Y1, int X2, int Y2, int px0, int px1,
                                                 int py0, int py1, const float4* __restrict__ P,
                                                 const int32_t* __restrict__ tri, int W, float invW, float invH,
                                                 unsigned long long* __restrict__ zb, unsigned int* tbits, int tiles_x,
                                                 int t, int* qcount, uint2* queue) {
    const I dx0 = X2 - X1, dy0 = Y2 - Y1;
    const I dx1 = X0 - X2, dy1 = Y0 - Y2;
    const I dx2 = X1 - X0, dy2 = Y1 - Y0;
    const I b0 = owns_edge(dx0, dy0) ? 1 : 0, b1 = owns_edge(dx1, dy1) ? 1 : 0, b2 = owns_edge(dx2, dy2) ? 1 : 0;
    const int Cx0 = px0 * 256 + 128;
    for (int py = py0; py <= py1; py++) {
        const int Cy = py * 256 + 128;
        I e0 = dx0 * (I)(Cy - Y1) - dy0 * (I)(Cx0 - X1);
        I e1 = dx1 * (I)(Cy - Y2) - dy1 * (I)(Cx0 - X2);
        I e2 = dx2 * (I)(Cy - Y0) - dy2 * (I)(Cx0 - X0);
        for (int px = px0; px <= px1; px++) {
            if (e0 + b0 > 0 && e1 + b1 > 0 && e2 + b2 > 0) {
                const int tile = (py >> 4) * tiles_x + (px >> 4);
                atomicOr(tbits + (tile >> 5), 1u << (tile & 31));
                const int slot = atomicAdd(qcount, 1);
                if (slot < kFragQueue) queue[slot] = make_uint2((uint32_t)t, ((uint32_t)py << 16) | (uint32_t)px);
                else resolve_fragment(P, tri, t, px, py, W, invW, invH, zb);  // large triangles only
            }
            e0 -= dy0 * 256;
            e1 -= dy1 * 256;
            e2 -= dy2 * 256;
        }
    }
}

constexpr int kTriPerThread = 4;  // triangles per thread: their 12 index loads and 12 snapped-vertex gathers are issued
                                  // back to back, so a block pays the two dependent L2 round trips once for 1024
                                  // triangles instead of once for 256 (the kernel's time was waves x block latency)

__device__ __forceinline__ void coverage_snapped_test(int t, int2 s0, int2 s1, int2 s2, const float4* __restrict__ P,
                                                      const int32_t* __restrict__ tri, int H, int W, float invW,
                                                      float invH, unsigned long long* __restrict__ zb,
                                                      unsigned int* tbits, int tiles_x, int* qcount, uint2* queue) {
    if (s0.x == kSnapRejected || s1.x == kSnapRejected || s2.x == kSnapRejected) return;
    int X0 = s0.x, Y0 = s0.y, X1 = s1.x, Y1 = s1.y, X2 = s2.x, Y2 = s2.y;
    const int minX = min(X0, min(X1, X2)), maxX = max(X0, max(X1, X2));
    const int minY = min(Y0, min(Y1, Y2)), maxY = max(Y0, max(Y1, Y2));
    const int px0 = max(0, (minX - 128 + 255) >> 8), px1 = min(W - 1, (maxX - 128) >> 8);
    const int py0 = max(0, (minY - 128 + 255) >> 8), py1 = min(H - 1, (maxY - 128) >> 8);
    if (px0 > px1 || py0 > py1) return;  // no pixel centre in the bounding box: the common case
    const bool small = (maxX - minX) < 32768 && (maxY - minY) < 32768 && px1 < 65536 && py1 < 65536;
    bool neg;
    if (small) {
        const int a = (X1 - X0) * (Y2 - Y0) - (X2 - X0) * (Y1 - Y0);  // |factors| < 2^15: exact in 32 bits
        if (a == 0) return;
        neg = a < 0;
    } else {
        const long long a = (long long)(X1 - X0) * (Y2 - Y0) - (long long)(X2 - X0) * (Y1 - Y0);
        if (a == 0) return;
        neg = a < 0;
    }
    if (neg) {  // orient for coverage only; barycentrics keep the original vertex order
        int tx = X1; X1 = X2; X2 = tx;
        int ty = Y1; Y1 = Y2; Y2 = ty;
    }
    if (small) cover_bbox_queue<int>(X0, Y0, X1, Y1, X2, Y2, px0, px1, py0, py1, P, tri, W, invW, invH, zb, tbits, tiles_x, t, qcount, queue);
    else cover_bbox_queue<long long>(X0, Y0, X1, Y1, X2, Y2, px0, px1, py0, py1, P, tri, W, invW, invH, zb, tbits, tiles_x, t, qcount, queue);
}

__global__ void __launch_bounds__(256, 4) raster_coverage_snapped_kernel(const float4* __restrict__ pos,
                                                                      const int2* __restrict__ snap,
                                                                      const int32_t* __restrict__ tri, int V, int T,
                                                                      int H, int W, float invW, float invH,
                                                                      unsigned long long* __restrict__ zbuf,
                                                                      uint32_t* __restrict__ gbits,
                                                                      uint32_t* __restrict__ glist,
                                                                      int* __restrict__ gcount, int tiles_x,
                                                                      int tiles_per_view) {
    extern __shared__ unsigned int tbits[];  // (tiles_per_view + 31) / 32 words
    __shared__ uint2 queue[8][kFragQueue];
    __shared__ int qcount[8];
    const int words = (tiles_per_view + 31) >> 5;
    for (int w = threadIdx.x; w < words; w += blockDim.x) tbits[w] = 0u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) qcount[warp] = 0;
    const int n = blockIdx.y;
    const int2* S = snap + (size_t)n * V;
    // ---- batched loads: indices of all kTriPerThread triangles, then their snapped vertices
    int tt[kTriPerThread], i0[kTriPerThread], i1[kTriPerThread], i2[kTriPerThread];
#pragma unroll
    for (int k = 0; k < kTriPerThread; k++) {
        tt[k] = (blockIdx.x * kTriPerThread + k) * blockDim.x + threadIdx.x;
        const bool ok = tt[k] < T;
        i0[k] = ok ? __ldg(tri + 3 * tt[k]) : -1;
        i1[k] = ok ? __ldg(tri + 3 * tt[k] + 1) : -1;
        i2[k] = ok ? __ldg(tri + 3 * tt[k] + 2) : -1;
    }
    int2 s0[kTriPerThread], s1[kTriPerThread], s2[kTriPerThread];
#pragma unroll
    for (int k = 0; k < kTriPerThread; k++) {
        const bool ok = (unsigned)i0[k] < (unsigned)V && (unsigned)i1[k] < (unsigned)V && (unsigned)i2[k] < (unsigned)V;
        s0[k] = ok ? __ldg(S + i0[k]) : make_int2(kSnapRejected, 0);
        s1[k] = ok ? __ldg(S + i1[k]) : make_int2(kSnapRejected, 0);
        s2[k] = ok ? __ldg(S + i2[k]) : make_int2(kSnapRejected, 0);
    }
    __syncthreads();  // tbits / qcount initialised
    unsigned long long* zb = zbuf + (size_t)n * H * W;
    const float4* P = pos + (size_t)n * V;
#pragma unroll
    for (int k = 0; k < kTriPerThread; k++)
        coverage_snapped_test(tt[k], s0[k], s1[k], s2[k], P, tri, H, W, invW, invH, zb, tbits, tiles_x, &qcount[warp],
                              queue[warp]);
    __syncwarp();
    const int nq = min(qcount[warp], kFragQueue);
    for (int e = lane; e < nq; e += 32) {
        const uint2 f = queue[warp][e];
        resolve_fragment(P, tri, (int)f.x, (int)(f.y & 0xffffu), (int)(f.y >> 16), W, invW, invH, zb);
    }
    __syncthreads();
    // flush: tiles this block touched first (global bitmap de-duplicates) are appended to the slot's work list
    for (int w = threadIdx.x; w < words; w += blockDim.x) {
        const unsigned int bits = tbits[w];
        if (bits == 0) continue;
        unsigned int fresh = bits & ~atomicOr(gbits + (size_t)n * words + w, bits);
        while (fresh) {
            const int bit = __ffs(fresh) - 1;
            fresh &= fresh - 1;
            const int tile = (w << 5) + bit;
            const int by = tile / tiles_x, bx = tile - by * tiles_x;
            glist[atomicAdd(gcount, 1)] = ((uint32_t)n << 20) | ((uint32_t)by << 10) | (uint32_t)bx;
        }
    }
}

__global__ void __launch_bounds__(256) raster_coverage_kernel(const float* __restrict__ pos,
                                                              const int32_t* __restrict__ tri, int V, int T, int H,
                                                              int W, unsigned long long* __restrict__ zbuf) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = blockIdx.y;
    if (t >= T) return;
    const int i0 = __ldg(tri + 3 * t), i1 = __ldg(tri + 3 * t + 1), i2 = __ldg(tri + 3 * t + 2);
    if ((unsigned)i0 >= (unsigned)V || (unsigned)i1 >= (unsigned)V || (unsigned)i2 >= (unsigned)V) return;
    const float* P = pos + (size_t)n * V * 4;
    const float4 p0 = ldg4(P + 4 * (size_t)i0), p1 = ldg4(P + 4 * (size_t)i1), p2 = ldg4(P + 4 * (size_t)i2);
    const float hw = (float)W * 0.5f, hh = (float)H * 0.5f;
    int X0, Y0, X1, Y1, X2, Y2;
    if (!snap_vertex(p0, hw, hh, X0, Y0)) return;
    if (!snap_vertex(p1, hw, hh, X1, Y1)) return;
    if (!snap_vertex(p2, hw, hh, X2, Y2)) return;
    const int minX = min(X0, min(X1, X2)), maxX = max(X0, max(X1, X2));
    const int minY = min(Y0, min(Y1, Y2)), maxY = max(Y0, max(Y1, Y2));
    // pixel centres (px*256+128) inside [min,max]
    const int px0 = max(0, (minX - 128 + 255) >> 8), px1 = min(W - 1, (maxX - 128) >> 8);
    const int py0 = max(0, (minY - 128 + 255) >> 8), py1 = min(H - 1, (maxY - 128) >> 8);
    if (px0 > px1 || py0 > py1) return;  // no pixel centre in the bounding box: the common case
    const long long area2 = (long long)(X1 - X0) * (Y2 - Y0) - (long long)(X2 - X0) * (Y1 - Y0);
    if (area2 == 0) return;
    if (area2 < 0) {  // orient for coverage only; barycentrics keep the original vertex order
        int tx = X1; X1 = X2; X2 = tx;
        int ty = Y1; Y1 = Y2; Y2 = ty;
    }
    const float invW = xd(1.0f, (float)W), invH = xd(1.0f, (float)H);
    unsigned long long* zb = zbuf + (size_t)n * H * W;
    // small triangles (every triangle of the HAM workloads): 32-bit edge functions cannot overflow
    if ((maxX - minX) < 32768 && (maxY - minY) < 32768) {
        cover_bbox<int>(X0, Y0, X1, Y1, X2, Y2, px0, px1, py0, py1, p0, p1, p2, W, invW, invH, zb, t);
    } else {
        cover_bbox<long long>(X0, Y0, X1, Y1, X2, Y2, px0, px1, py0, py1, p0, p1, p2, W, invW, invH, zb, t);
    }
}

// zbuf key -> rast_out texel (+ optional rast_db) of pixel (n, py, px).
__device__ __forceinline__ void resolve_pixel(unsigned long long key, const float* __restrict__ pos,
                                              const int32_t* __restrict__ tri, int V, int H, int W, int n, int py, int px,
                                              size_t pix, float4* __restrict__ rast, float4* __restrict__ rast_db) {
    const int t = (int)(uint32_t)key;
    const float* P = pos + (size_t)n * V * 4;
    const float4 p0 = ldg4(P + 4 * (size_t)__ldg(tri + 3 * t));
    const float4 p1 = ldg4(P + 4 * (size_t)__ldg(tri + 3 * t + 1));
    const float4 p2 = ldg4(P + 4 * (size_t)__ldg(tri + 3 * t + 2));
    const float invW = xd(1.0f, (float)W), invH = xd(1.0f, (float)H);
    const Bary b = bary_at(p0, p1, p2, px, py, invW, invH);
    rast[pix] = make_float4(b.u, b.v, b.zw, (float)(t + 1));
    if (rast_db) {
        // analytic d(u,v)/d(pixel x,y) of the unclamped barycentrics
        const float fx = (2.0f * px + 1.0f) * invW - 1.0f, fy = (2.0f * py + 1.0f) * invH - 1.0f;
        const float q0x = p0.x - fx * p0.w, q0y = p0.y - fy * p0.w;
        const float q1x = p1.x - fx * p1.w, q1y = p1.y - fy * p1.w;
        const float q2x = p2.x - fx * p2.w, q2y = p2.y - fy * p2.w;
        const float a0 = q1x * q2y - q1y * q2x, a1 = q2x * q0y - q2y * q0x, a2 = q0x * q1y - q0y * q1x;
        const float iw = 1.0f / (a0 + a1 + a2);
        const float da0x = -p1.w * q2y + q1y * p2.w, da0y = -q1x * p2.w + p1.w * q2x;
        const float da1x = -p2.w * q0y + q2y * p0.w, da1y = -q2x * p0.w + p2.w * q0x;
        const float da2x = -p0.w * q1y + q0y * p1.w, da2y = -q0x * p1.w + p0.w * q1x;
        const float datx = da0x + da1x + da2x, daty = da0y + da1y + da2y;
        const float u = a0 * iw, v = a1 * iw;
        const float sx = 2.0f * invW, sy = 2.0f * invH;
        rast_db[pix] = make_float4((da0x - u * datx) * iw * sx, (da0y - u * daty) * iw * sy,
                                   (da1x - v * datx) * iw * sx, (da1y - v * daty) * iw * sy);
    }
}

// zbuf -> rast_out (+ optional rast_db).  One thread per pixel, float4 stores.
__global__ void __launch_bounds__(256) raster_resolve_kernel(const float* __restrict__ pos,
                                                             const int32_t* __restrict__ tri, int V, int H, int W,
                                                             size_t npix,
                                                             const unsigned long long* __restrict__ zbuf,
                                                             float4* __restrict__ rast, float4* __restrict__ rast_db) {
    const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= npix) return;
    const unsigned long long key = zbuf[pix];
    if (key == ZB_EMPTY) {
        rast[pix] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rast_db) rast_db[pix] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    const size_t hw_ = (size_t)H * W;
    const int n = (int)(pix / hw_);
    const int rem = (int)(pix - (size_t)n * hw_);
    const int py = rem / W, px = rem - py * W;
    resolve_pixel(key, pos, tri, V, H, W, n, py, px, pix, rast, rast_db);
}

// Tile form of the resolve pass (fmhr_rasterize_fwd_meshlets): one block per 16x16 screen tile.  A tile the coverage
// kernel never touched (bitmap bit clear: ~90 % of the frame in the HAM workloads) is zero-filled without reading the
// z-buffer; a dirty tile resolves its winners and puts the keys it consumed back to EMPTY, so the z-buffer is clean again
// when the call returns and the next call needs no 8-byte-per-pixel clear pass.
__global__ void __launch_bounds__(kTile * kTile) raster_resolve_tiles_kernel(
    const float* __restrict__ pos, const int32_t* __restrict__ tri, int V, int H, int W, int tiles_x, int words,
    const uint32_t* __restrict__ tile_bits, unsigned long long* __restrict__ zbuf, float4* __restrict__ rast,
    float4* __restrict__ rast_db) {
    const int tile = blockIdx.x, n = blockIdx.y;
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int px = tx * kTile + (threadIdx.x & (kTile - 1)), py = ty * kTile + (threadIdx.x / kTile);
    if (px >= W || py >= H) return;
    const size_t pix = ((size_t)n * H + py) * W + px;
    const bool dirty = (__ldg(tile_bits + (size_t)n * words + (tile >> 5)) >> (tile & 31)) & 1u;
    unsigned long long key = ZB_EMPTY;
    if (dirty) {
        key = zbuf[pix];
        if (key != ZB_EMPTY) zbuf[pix] = ZB_EMPTY;
    }
    if (key == ZB_EMPTY) {
        rast[pix] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rast_db) rast_db[pix] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    resolve_pixel(key, pos, tri, V, H, W, n, py, px, pix, rast, rast_db);
}

// d(u,v)/d(pos) scattered to the three vertices (SURVEY.md Appendix A "rasterize bwd").
__global__ void __launch_bounds__(256) raster_bwd_kernel(const float* __restrict__ pos, const int32_t* __restrict__ tri,
                                                         const float4* __restrict__ rast, const float4* __restrict__ dy,
                                                         int V, int T, int H, int W, size_t npix,
                                                         float* __restrict__ grad_pos) {
    const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= npix) return;
    const float4 r = __ldg(rast + pix);
    const int t = (int)r.w - 1;
    if (t < 0 || t >= T) return;
    const float4 g = __ldg(dy + pix);
    if (g.x == 0.0f && g.y == 0.0f) return;
    const size_t hw_ = (size_t)H * W;
    const int n = (int)(pix / hw_);
    const int rem = (int)(pix - (size_t)n * hw_);
    const int py = rem / W, px = rem - py * W;
    const int i0 = __ldg(tri + 3 * t), i1 = __ldg(tri + 3 * t + 1), i2 = __ldg(tri + 3 * t + 2);
    const float* P = pos + (size_t)n * V * 4;
    const float4 p0 = ldg4(P + 4 * (size_t)i0), p1 = ldg4(P + 4 * (size_t)i1), p2 = ldg4(P + 4 * (size_t)i2);
    const float fx = (float)(2 * px + 1) / (float)W - 1.0f;
    const float fy = (float)(2 * py + 1) / (float)H - 1.0f;
    const float q0x = p0.x - fx * p0.w, q0y = p0.y - fy * p0.w;
    const float q1x = p1.x - fx * p1.w, q1y = p1.y - fy * p1.w;
    const float q2x = p2.x - fx * p2.w, q2y = p2.y - fy * p2.w;
    const float a0 = q1x * q2y - q1y * q2x, a1 = q2x * q0y - q2y * q0x, a2 = q0x * q1y - q0y * q1x;
    const float at = a0 + a1 + a2;
    const float iw = 1.0f / (at + copysignf(1e-6f, at));
    const float b0 = a0 * iw, b1 = a1 * iw;
    const float gb0 = g.x * iw, gb1 = g.y * iw, gbb = gb0 * b0 + gb1 * b1;
    const float g0x = gbb * (q2y - q1y) - gb1 * q2y;
    const float g1x = gbb * (q0y - q2y) + gb0 * q2y;
    const float g2x = gbb * (q1y - q0y) - gb0 * q1y + gb1 * q0y;
    const float g0y = gbb * (q1x - q2x) + gb1 * q2x;
    const float g1y = gbb * (q2x - q0x) - gb0 * q2x;
    const float g2y = gbb * (q0x - q1x) + gb0 * q1x - gb1 * q0x;
    float4* G = reinterpret_cast<float4*>(grad_pos + (size_t)n * V * 4);
    atomicAdd(G + i0, make_float4(g0x, g0y, 0.f, -fx * g0x - fy * g0y));
    atomicAdd(G + i1, make_float4(g1x, g1y, 0.f, -fx * g1x - fy * g1y));
    atomicAdd(G + i2, make_float4(g2x, g2y, 0.f, -fx * g2x - fy * g2y));
}

// Launchers shared with the fused HAM path (ham.cu).
int launch_raster_coverage(const float* pos, const int32_t* tri, int N, int V, int T, int H, int W,
                           unsigned long long* zbuf, cudaStream_t st) {
    dim3 grid(cdiv(T, 256), N);
    raster_coverage_kernel<<<grid, 256, 0, st>>>(pos, tri, V, T, H, W, zbuf);
    FMHR_LAUNCH_CHECK();
    return FMHR_OK;
}

int launch_raster_coverage_snapped(const float4* pos, const int2* snap, const int32_t* tri, int N, int V, int T, int H,
                                   int W, unsigned long long* zbuf, uint32_t* tbits, uint32_t* tlist, int* tcount,
                                   int tiles_x, int tiles_per_view, cudaStream_t st) {
    dim3 grid(cdiv(T, 256 * kTriPerThread), N);
    const size_t smem = (size_t)((tiles_per_view + 31) / 32) * sizeof(unsigned int);
    if (smem > 48 * 1024) {
        set_error("launch_raster_coverage_snapped: %d tiles per view exceed the shared-memory bitmap", tiles_per_view);
        return FMHR_EUNSUPPORTED;
    }
    raster_coverage_snapped_kernel<<<grid, 256, smem, st>>>(pos, snap, tri, V, T, H, W, 1.0f / (float)W, 1.0f / (float)H,
                                                            zbuf, tbits, tlist, tcount, tiles_x, tiles_per_view);
    FMHR_LAUNCH_CHECK();
    return FMHR_OK;
}

}  // namespace fmhr

using namespace fmhr;

extern "C" size_t fmhr_rasterize_workspace_bytes(int N, int H, int W) {
    if (N <= 0 || H <= 0 || W <= 0) return 0;
    return (size_t)N * H * W * sizeof(unsigned long long);
}

extern "C" int fmhr_rasterize_fwd(const float* pos, const int32_t* tri, int N, int V, int T, int H, int W, float* rast,
                                  float* rast_db, void* workspace, size_t workspace_bytes, fmhr_stream_t stream) {
    FMHR_CHECK_ARG(pos && (tri || T == 0) && rast && workspace);
    FMHR_CHECK_ARG(N > 0 && V > 0 && T >= 0 && H > 0 && W > 0);
    FMHR_CHECK_ARG(workspace_bytes >= fmhr_rasterize_workspace_bytes(N, H, W));
    FMHR_CHECK_ARG(T < (1 << 24));  // triangle id + 1 must stay exact in fp32 (rast[...,3])
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* zbuf = (unsigned long long*)workspace;
    const size_t npix = (size_t)N * H * W;
    FMHR_CUDA(cudaMemsetAsync(zbuf, 0xFF, npix * sizeof(unsigned long long), st));
    if (T > 0) {
        int rc = launch_raster_coverage(pos, tri, N, V, T, H, W, zbuf, st);
        if (rc) return rc;
    }
    raster_resolve_kernel<<<cdiv(npix, 256), 256, 0, st>>>(pos, tri, V, H, W, npix, zbuf, (float4*)rast,
                                                           (float4*)rast_db);
    FMHR_LAUNCH_CHECK();
    return FMHR_OK;
}

namespace fmhr {
int launch_meshlet_coverage_clip(const float* pos, int N, int V, int H, int W, const int32_t* ml_vptr,
                                 const int32_t* ml_verts, const uint32_t* ml_tri2, int n_meshlets, int ml_tris,
                                 int ml_max_verts, unsigned long long* zbuf, uint32_t* tile_bits, cudaStream_t st);
}

extern "C" size_t fmhr_rasterize_tile_words(int N, int H, int W) {
    if (N <= 0 || H <= 0 || W <= 0) return 0;
    return (size_t)N * (size_t)((cdiv(W, kTile) * cdiv(H, kTile) + 31) / 32);
}

extern "C" int fmhr_rasterize_fwd_meshlets(const float* pos, const int32_t* tri, int N, int V, int T, int H, int W,
                                           float* rast, float* rast_db, const int32_t* ml_vptr, const int32_t* ml_verts,
                                           const uint32_t* ml_tri2, int n_meshlets, int ml_tris, int ml_max_verts,
                                           void* zbuf_ws, size_t zbuf_bytes, int zbuf_is_clean, uint32_t* tile_bits,
                                           fmhr_stream_t stream) {
    FMHR_CHECK_ARG(pos && tri && rast && zbuf_ws && tile_bits && ml_vptr && ml_verts && ml_tri2);
    FMHR_CHECK_ARG(N > 0 && V > 0 && T > 0 && H > 0 && W > 0 && n_meshlets > 0);
    FMHR_CHECK_ARG(zbuf_bytes >= fmhr_rasterize_workspace_bytes(N, H, W));
    FMHR_CHECK_ARG(T < (1 << 24) && N <= 65535);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* zbuf = (unsigned long long*)zbuf_ws;
    const size_t npix = (size_t)N * H * W;
    if (!zbuf_is_clean) FMHR_CUDA(cudaMemsetAsync(zbuf, 0xFF, npix * sizeof(unsigned long long), st));
    const int tiles_x = cdiv(W, kTile), tiles_pv = tiles_x * cdiv(H, kTile), words = (tiles_pv + 31) / 32;
    FMHR_CUDA(cudaMemsetAsync(tile_bits, 0, (size_t)N * words * sizeof(uint32_t), st));
    int rc = launch_meshlet_coverage_clip(pos, N, V, H, W, ml_vptr, ml_verts, ml_tri2, n_meshlets, ml_tris, ml_max_verts,
                                          zbuf, tile_bits, st);
    if (rc) return rc;
    raster_resolve_tiles_kernel<<<dim3(tiles_pv, N), kTile * kTile, 0, st>>>(pos, tri, V, H, W, tiles_x, words, tile_bits,
                                                                            zbuf, (float4*)rast, (float4*)rast_db);
    FMHR_LAUNCH_CHECK();
    return FMHR_OK;
}

extern "C" int fmhr_rasterize_bwd(const float* pos, const int32_t* tri, const float* rast, const float* dy, int N,
                                  int V, int T, int H, int W, float* grad_pos, fmhr_stream_t stream) {
    FMHR_CHECK_ARG(pos && tri && rast && dy && grad_pos);
    FMHR_CHECK_ARG(N > 0 && V > 0 && T >= 0 && H > 0 && W > 0);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t npix = (size_t)N * H * W;
    FMHR_CUDA(cudaMemsetAsync(grad_pos, 0, (size_t)N * V * 4 * sizeof(float), st));
    raster_bwd_kernel<<<cdiv(npix, 256), 256, 0, st>>>(pos, tri, (const float4*)rast, (const float4*)dy, V, T, H, W,
                                                       npix, grad_pos);
    FMHR_LAUNCH_CHECK();
    return FMHR_OK;
}
