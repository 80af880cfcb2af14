// Fused HAM iteration (mesh_sfs_optim.py:198-237 phase A, :253-310 phase B).
//
// One iteration is two C-ABI calls so that a multi-GPU host can all-reduce the packed gradient buffer between them:
//   fmhr_ham_step_render : vertices = vertices_tmp + delta -> vertex normals -> per-view clip positions ->
//                          z-buffer coverage -> shade (interpolate + normalise + SH + albedo) -> antialias + losses ->
//                          pixel backward (antialias / shading / interpolate / rasterize gradients scattered straight
//                          into WORLD-space per-vertex accumulators with float4 vector atomics)
//   fmhr_ham_step_update : regularisers (uniform Laplacian x2, edge hinge, delta), normal backward, loss
//                          normalisation by the global valid-pixel count, fused Adam.
// Nothing here materialises rast_out, the [n,V,7] attribute tensor, the [n,H,W,7] interpolated plane or the
// [n,F,3,3] face-vertex gather of the reference; per pixel the only HBM planes are the 8-byte z-buffer, one float4
// shaded colour and one float4 pixel-gradient (two of each in phase A, where 6 channels are antialiased).
#include <stdlib.h>

#include "aa_rule.cuh"

// resident 256-thread blocks per SM the pixel / coverage kernels are compiled for (register budget = 65536 / 256 / N);
// measured on B200 (tools/run_variants.sh): shade and backward are faster spill-free at 3, coverage at 5
// Checked builds (-DFMHR_CHECKED, tools/build_variant.sh checked -DFMHR_CHECKED; tests/test_gpu_checked.py): every write into a
// per-warp shared buffer or a capacity-bounded work list asserts its index and records a site code in g_dcheck instead of
// writing out of bounds (this pool refuses compute-sanitizer; these are the indices memcheck would have watched).  The
// product build compiles the checks to nothing.
#ifdef FMHR_CHECKED
static __device__ unsigned int g_dcheck;
#define FMHR_DCHECK(cond, site) ((cond) ? true : (atomicOr(&g_dcheck, 1u << (site)), false))
#else
#define FMHR_DCHECK(cond, site) (true)
#endif
#ifndef FMHR_LB_COVERAGE
#define FMHR_LB_COVERAGE 5
#endif
#ifndef FMHR_LB_SHADE
#define FMHR_LB_SHADE 3
#endif
#ifndef FMHR_LB_SHADE_BWD
#define FMHR_LB_SHADE_BWD 2  // shade pass with the speculative backward fused in (phase B training step)
#endif
#ifndef FMHR_LB_AA
#define FMHR_LB_AA 4
#endif
#ifndef FMHR_LB_SCAN
#define FMHR_LB_SCAN 6  // a short latency chain per warp: more resident warps = fewer serial units per warp (4: 4,810, 6: 4,895, 8: 4,875 iters/s)
#endif
#ifndef FMHR_PEER_POST_DEFAULT
// only block 0 posts this rank's step word: 193 blocks x (system fence + remote store) made the rendezvous of two ranks that
// arrive together take 13-15 us; one remote store per peer: 2-3 us (profiles/r2_scaling_exchange.md)
#define FMHR_PEER_POST_DEFAULT 2
#endif
#ifndef FMHR_SIDE_PRIO
#define FMHR_SIDE_PRIO 0  // high priority for the vertex side stream delays the head of the coverage kernel: 4,810 vs 5,047 iters/s
#endif
#ifndef FMHR_DEFAULT_VIEW_GROUPS
#define FMHR_DEFAULT_VIEW_GROUPS 1
#endif
#ifndef FMHR_AA_HOIST
#define FMHR_AA_HOIST 0
#endif
#ifndef FMHR_COV_SMEM_FLOOR
#define FMHR_COV_SMEM_FLOOR 0
#endif
#ifndef FMHR_REG_GRAD_LATE
#define FMHR_REG_GRAD_LATE 0
#endif
#ifndef FMHR_SIDE2
#define FMHR_SIDE2 0  // regulariser forward / gradients on a second side stream behind the normals (beside the triangle records)
#endif
#ifndef FMHR_INTERLEAVE
#define FMHR_INTERLEAVE 0  // bit 0: shade, bit 1: antialias - consecutive warp batches go to different blocks
#endif
#ifndef FMHR_PAIR_SECTORS
#define FMHR_PAIR_SECTORS 1  // shade pass: lane pairs complete whole 32-byte sectors per vector RED (scatter_sector_pairs)
#endif
#ifndef FMHR_LB_BWD
#define FMHR_LB_BWD 3
#endif

namespace fmhr {

// Per-view combined matrix, written once per step by the prep kernel (row-vector convention of get_data.py:96-97):
//   viewM[n][4i+j] = (w2c @ proj)[i][j],  clip_j = sum_i world_i * M[4i+j] + M[12+j]  (mesh_sfs_optim.py:262-264);
// the same 16 floats are d(clip_j)/d(world_i) for the backward pass.  Every kernel derives clip positions from the
// world-space vertex and this matrix with clip_from_world() (fixed FMA order -> bit-identical everywhere), so no
// [n,V] position array exists in HBM.
constexpr int kViewM = 16;

struct HamWs {
    unsigned long long* zbuf[2];  // [n,H,W] x2: `zbuf_slot` is rasterised this step, the other is reset for the next
    float4* plane[4];          // [n,H,W] each
    float* viewM;              // [n,16] combined world->clip matrix per view
    // Compact work lists (rebuilt every iteration):
    uint2* clist;              // [P]   covered pixels (pixel index, triangle | valid << 31), by the scan / shade passes
    uint32_t* rlist;           // [P/2] empty pixels touching a covered one (antialias receivers), by the scan pass
    uint32_t* ringbits;        // [P/32] de-duplication bitmap of rlist; zero outside an iteration
    uint4* plist_a;            // [P/2] blending pixel pairs found by the antialias pass: (pixel0, flags, alpha, i1)
    uint32_t* plist_b;         // [P/2]                                                    i2
    int* ccount;               // inside common_region (zeroed every step)
    int* rcount;
    int* pcount;
    uint2* qlist;              // [P/4] phase B: pixels whose antialiased L1 sign differs from the shade pass' speculation
    int* qcount;               //        (pixel, packed sign deltas), back-propagated by the pair kernel
    int* status;               // bit 0: pair / correction list overflow, bit 1: ring list overflow, bit 2: a peer never arrived
    // Tile work lists (16x16 tiles).  slot[s]: tiles of z-buffer slot s that received fragments (bitmap for
    // de-duplication + compact list + count, filled by the coverage kernel).  The scan pass walks the list of the slot
    // rasterised this step (idle tiles cost nothing) and resets the tiles of the other slot.
    char* slot_region[2];        // [count (256 B) | bitmap] zeroed together when the slot is reused
    int* tcount[2];
    uint32_t* tbits[2];          // [n, words_per_view]
    uint32_t* tlist[2];          // [n * tiles_per_view]
    char* common_region;         // [acc | counters] zeroed every step
    size_t common_bytes, slot_bytes;
    // Vertex-domain records (every gather of the pixel / update kernels is a 128- or 256-bit load of one 32-byte sector):
    float4* vg;                // [V,2]  vg[2i] = (x, y, z, 0) current vertex,   vg[2i+1] = d(loss)/d(raw normal) (update pass 1)
    float4* vattr;             // [V,2]  vattr[2i] = (unit normal, degenerate flag), vattr[2i+1] = (albedo b,g,r, 0)
    float4* raw4;              // [V]    (un-normalised normal sum N, |N|)
    float4* ys;                // [V,2]  ys[2i] = (yhat_v / deg, deg), ys[2i+1] = (yhat_a / deg, 0): Laplacian backward rows
    float4* greg;              // [V,2]  phase B: (d regularisers / d delta, 0), (d albedo Laplacian / d albedo, 0), weighted
    float4* trirec;            // [T,10] per-triangle record of the pixel passes (kTriRec floats), rebuilt every step
    float* gsh;                // [n_sh_rows,9] un-normalised SH gradients by SH row (phase A)
    double* acc;               // [8][32] (32-way spread against same-address atomics):
                               // 0 n_valid, 1 abs_sum, 2 mask_sq correction, 3 lap_v, 4 lap_a, 5 edge, 6 delta
    float* adam_sc;            // [8]: step_size / bias2_sqrt for delta, albedo, sh
};

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// Optional per-kernel timing (fmhr_ham_stage_times): events recorded on the launching stream between launches.
constexpr int kMaxStages = 16;
struct StageTimer {
    cudaEvent_t ev[kMaxStages + 1];
    int n = 0;
    cudaStream_t st;
    void mark() { if (n <= kMaxStages) cudaEventRecord(ev[n++], st); }
};
static thread_local StageTimer* g_timer = nullptr;
#define FMHR_STAGE_MARK() do { if (g_timer) g_timer->mark(); } while (0)

static size_t ham_layout(const fmhr_ham_config* c, char* base, HamWs* ws) {
    // every per-view / per-pixel array is laid out for the configuration's view CAPACITY (>= the step's batch): steps with
    // different batch sizes then share one layout (view slot k sits at the same addresses) and keep each other's
    // z-buffer / tile-list invariants, so the caller does not have to reset between them
    const size_t ncap = (size_t)(c->n_views_capacity > c->n_views ? c->n_views_capacity : c->n_views);
    const size_t P = ncap * c->H * c->W, V = (size_t)c->V;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += align256(bytes); return p; };
    char* p;
    for (int i = 0; i < 2; i++) { p = take(P * 8); if (ws) ws->zbuf[i] = (unsigned long long*)p; }
    const int nplanes = c->phase == 0 ? 4 : 2;
    for (int i = 0; i < 4; i++) {
        p = (i < nplanes) ? take(P * 16) : nullptr;
        if (ws) ws->plane[i] = (float4*)p;
    }
    p = take(ncap * kViewM * 4); if (ws) ws->viewM = (float*)p;
    p = take(P * 8); if (ws) ws->clist = (uint2*)p;
    p = take((P / 2 + 64) * 4); if (ws) ws->rlist = (uint32_t*)p;
    p = take((P / 32 + 64) * 4); if (ws) ws->ringbits = (uint32_t*)p;
    p = take((P / 2 + 64) * 16); if (ws) ws->plist_a = (uint4*)p;
    p = take((P / 2 + 64) * 4); if (ws) ws->plist_b = (uint32_t*)p;
    p = take((P / 4 + 64) * 8); if (ws) ws->qlist = (uint2*)p;
    const size_t tiles_pv = (size_t)((c->W + 15) / 16) * ((c->H + 15) / 16);
    const size_t words = ncap * ((tiles_pv + 31) / 32);
    const size_t slot_bytes = 256 + align256(words * 4);
    for (int i = 0; i < 2; i++) {
        p = take(slot_bytes);
        if (ws) { ws->slot_region[i] = p; ws->tcount[i] = (int*)p; ws->tbits[i] = (uint32_t*)(p + 256); }
        p = take(ncap * tiles_pv * 4); if (ws) ws->tlist[i] = (uint32_t*)p;
    }
    const size_t common_bytes = 8 * 32 * sizeof(double) + 256;
    p = take(common_bytes);
    if (ws) {
        ws->common_region = p; ws->common_bytes = common_bytes; ws->slot_bytes = slot_bytes;
        ws->acc = (double*)p;
        int* cnt = (int*)(p + 8 * 32 * sizeof(double));
        ws->ccount = cnt; ws->rcount = cnt + 1; ws->pcount = cnt + 2; ws->status = cnt + 3; ws->qcount = cnt + 4;
    }
    p = take(V * 32); if (ws) ws->vg = (float4*)p;
    p = take(V * 32); if (ws) ws->vattr = (float4*)p;
    p = take(V * 16); if (ws) ws->raw4 = (float4*)p;
    p = take(V * 32); if (ws) ws->ys = (float4*)p;
    p = take(V * 32); if (ws) ws->greg = (float4*)p;
    p = take((size_t)c->T * 160); if (ws) ws->trirec = (float4*)p;
    p = take((size_t)c->n_sh_rows * 9 * 4); if (ws) ws->gsh = (float*)p;
    p = take(8 * sizeof(float)); if (ws) ws->adam_sc = (float*)p;
    return off;
}

// ------------------------------------------------------------------------------------------------
// vertex-domain prologue
// ------------------------------------------------------------------------------------------------
// 256-bit gather of one 32-byte record (LDG.E.256, sm_100+): one L1 wavefront instead of two.
struct F8 { float4 a, b; };
__device__ __forceinline__ F8 ldg256(const float4* p) {
    F8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w)
                 : "l"(p));
    return r;
}
// same, for records written earlier in the SAME kernel launch sequence but read through the coherent path
__device__ __forceinline__ void st256(float4* p, const float* r) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]),
                 "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7])
                 : "memory");
}
__device__ __forceinline__ F8 ld256(const float4* p) {
    F8 r;
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w)
                 : "l"(p));
    return r;
}

// vertices = vertices_tmp + delta (mesh_sfs_optim.py:253).  The same launch re-arms the step's accumulators (the packed
// gradient buffer = 3V float4, loss / work-list scratch, SH gradients) so the iteration has no memset nodes, and
// computes the per-view combined world->clip matrices.
__global__ void __launch_bounds__(256) ham_vertex_prep_kernel(const float* __restrict__ vtmp,
                                                              const float* __restrict__ delta, int V,
                                                              float4* __restrict__ vg, float4* __restrict__ packed4,
                                                              uint32_t* __restrict__ z0, int n0,
                                                              uint32_t* __restrict__ z1, int n1,
                                                              uint32_t* __restrict__ z2, int n2,
                                                              const float* __restrict__ w2cs,
                                                              const float* __restrict__ projs,
                                                              const int32_t* __restrict__ view_idx, int n_views,
                                                              float* __restrict__ viewM,
                                                              const int32_t* __restrict__ ml_vptr,
                                                              const int32_t* __restrict__ ml_verts, int ml_stride,
                                                              int ml_total, float4* __restrict__ ml_pos) {
    FMHR_TRACE_SCOPE(0);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ml_total) {  // vertices in meshlet order, padded to ml_stride per meshlet (w = 0: the snap rejects the padding)
        const int m = i / ml_stride, l = i - m * ml_stride;
        const int vb = __ldg(ml_vptr + m), nv = __ldg(ml_vptr + m + 1) - vb;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (l < nv) {
            const size_t k = 3 * (size_t)__ldg(ml_verts + vb + l);
            p = make_float4(vtmp[k] + delta[k], vtmp[k + 1] + delta[k + 1], vtmp[k + 2] + delta[k + 2], 0.f);
        }
        ml_pos[i] = p;
    }
    if (i < n_views * kViewM) {
        const int n = i >> 4, r = (i >> 2) & 3, j = i & 3;
        const float* Wm = w2cs + (size_t)view_idx[n] * 16;
        const float* Pm = projs + (size_t)view_idx[n] * 16;
        viewM[i] = __fmaf_rn(Wm[4 * r + 3], Pm[12 + j], __fmaf_rn(Wm[4 * r + 2], Pm[8 + j],
                             __fmaf_rn(Wm[4 * r + 1], Pm[4 + j], __fmul_rn(Wm[4 * r], Pm[j]))));
    }
    if (i < V) {
        const size_t k = 3 * (size_t)i;
        vg[2 * (size_t)i] = make_float4(vtmp[k] + delta[k], vtmp[k + 1] + delta[k + 1], vtmp[k + 2] + delta[k + 2], 0.f);
    }
    if (i <= 3 * V) packed4[i] = make_float4(0.f, 0.f, 0.f, 0.f);  // 3V accumulator float4 + the four scalars behind them
    if (i < n0) z0[i] = 0u;
    if (i < n1) z1[i] = 0u;
    if (i < n2) z2[i] = 0u;
}

// Area-weighted vertex normals (models/utils.py:508-548, corner form of the face normal), 4 lanes per vertex over the
// vertex->face CSR; writes the packed attribute records of the pixel passes.
__global__ void __launch_bounds__(256, 6) ham_normals_kernel(const float4* __restrict__ vg, const float* __restrict__ albedo,
                                                          const int32_t* __restrict__ v2f_ptr,
                                                          const int2* __restrict__ v2f_nbr, int V,
                                                          float4* __restrict__ vattr, float4* __restrict__ raw4) {
    FMHR_TRACE_SCOPE(1);
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 2, sub = threadIdx.x & 3;  // 4 lanes per vertex: ONE wave of blocks
    float ax = 0.f, ay = 0.f, az = 0.f;
    if (i < V) {
        const int b = __ldg(v2f_ptr + i), e = __ldg(v2f_ptr + i + 1);
        const float4 pk = vg[2 * (size_t)i];
        for (int j = b + sub; j < e; j += 4) {
            const int2 nb = __ldg(v2f_nbr + j);
            const float4 pa = vg[2 * (size_t)nb.x], pb = vg[2 * (size_t)nb.y];
            const float ux = pa.x - pk.x, uy = pa.y - pk.y, uz = pa.z - pk.z;
            const float wx = pb.x - pk.x, wy = pb.y - pk.y, wz = pb.z - pk.z;
            ax += uy * wz - uz * wy; ay += uz * wx - ux * wz; az += ux * wy - uy * wx;
        }
    }
#pragma unroll
    for (int o = 2; o > 0; o >>= 1) {
        ax += __shfl_xor_sync(0xffffffffu, ax, o);
        ay += __shfl_xor_sync(0xffffffffu, ay, o);
        az += __shfl_xor_sync(0xffffffffu, az, o);
    }
    if (i >= V || sub != 0) return;
    const float len = sqrtf(ax * ax + ay * ay + az * az);
    const float inv = 1.0f / fmaxf(len, 1e-6f);
    raw4[i] = make_float4(ax, ay, az, len);
    // flag: the normalisation backward of this vertex does not project (models/utils.py:547 clamps the norm at 1e-6)
    vattr[2 * (size_t)i] = make_float4(ax * inv, ay * inv, az * inv, len > 1e-6f ? 0.0f : 1.0f);
    const float* a = albedo + 3 * (size_t)i;
    vattr[2 * (size_t)i + 1] = make_float4(a[0], a[1], a[2], 0.0f);
}

struct ViewM { float m[16]; };
__device__ __forceinline__ ViewM load_viewM(const float* __restrict__ g) {
    ViewM M;
    const float4* g4 = reinterpret_cast<const float4*>(g);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float4 r = __ldg(g4 + k);
        M.m[4 * k] = r.x; M.m[4 * k + 1] = r.y; M.m[4 * k + 2] = r.z; M.m[4 * k + 3] = r.w;
    }
    return M;
}
// clip = [v,1] @ (w2c @ proj); fixed FMA order, shared by every kernel that needs a clip position
__device__ __forceinline__ float4 clip_from_world(const float* M, const float4 v) {
    return make_float4(__fmaf_rn(v.z, M[8], __fmaf_rn(v.y, M[4], __fmaf_rn(v.x, M[0], M[12]))),
                       __fmaf_rn(v.z, M[9], __fmaf_rn(v.y, M[5], __fmaf_rn(v.x, M[1], M[13]))),
                       __fmaf_rn(v.z, M[10], __fmaf_rn(v.y, M[6], __fmaf_rn(v.x, M[2], M[14]))),
                       __fmaf_rn(v.z, M[11], __fmaf_rn(v.y, M[7], __fmaf_rn(v.x, M[3], M[15]))));
}

// ------------------------------------------------------------------------------------------------
// Coverage (replaces dr.rasterize's visibility, mesh_sfs_optim.py:267): one block = one meshlet in one view.
//  1. the meshlet's <= 1024 vertices are gathered (16 B each, from the 1.6 MB vertex record array), transformed to clip
//     space and snapped to the rasteriser's 24.8 grid into SHARED memory - once per (view, meshlet vertex) instead of a
//     [n,V] array in HBM written by a transform kernel and gathered three times per triangle;
//  2. every thread tests TPT triangles against their integer pixel-centre bounding box on the shared snapped coordinates
//     (micropolygons: ~60 % contain no pixel centre); the survivors are compacted per warp;
//  3. the survivors are processed with all lanes busy: integer edge functions; covered pixel centres go to a per-warp
//     queue and are then depth-resolved densely (perspective barycentric depth from the shared clip positions, 64-bit
//     atomicMin on depth | ORIGINAL triangle id); touched 16x16 tiles are collected in a shared bitmap and appended to
//     the slot's global work list at block end.
// Rule and arithmetic are those of raster.cu and of the CPU checker (bit-exact ids and depth).
// ------------------------------------------------------------------------------------------------
template <typename I>
__device__ __forceinline__ bool ml_owns_edge(I dx, I dy) { return dy > 0 || (dy == 0 && dx > 0); }

// Coverage of one candidate triangle (its bounding box holds at least one pixel centre): integer edge functions on the
// snapped corners, every covered pixel is depth-resolved in place from the shared clip positions.
constexpr int kFragQueue = 160;  // per-warp capacity; overflowing fragments (large triangles) are resolved in place

__device__ __forceinline__ void ml_resolve(const float4* pos_s, uint2 rec, int px, int py, int W, float invW, float invH,
                                           unsigned long long* __restrict__ zb) {
    const float4 p0 = pos_s[rec.x & 1023u], p1 = pos_s[(rec.x >> 10) & 1023u], p2 = pos_s[(rec.x >> 20) & 1023u];
    const Bary b = bary_at(p0, p1, p2, px, py, invW, invH);
    if (!FMHR_DCHECK(px >= 0 && px < W && py >= 0 && rec.x != 0xffffffffu, 7)) return;
    atomicMin(&zb[(size_t)py * W + px], ((unsigned long long)depth_key(b.zw) << 32) | rec.y);
}

// Tile bookkeeping of a fragment: the 16x16 tile's bit in the block-wide shared bitmap (read first: after the first few
// fragments of a block the bit is set and no atomic is issued).
__device__ __forceinline__ void ml_mark_tile(unsigned int* tbits, int tiles_x, int px, int py) {
    const int tile = (py >> 4) * tiles_x + (px >> 4);
    const unsigned int bit = 1u << (tile & 31);
    if (!(tbits[tile >> 5] & bit)) atomicOr(tbits + (tile >> 5), bit);
}

template <typename I>
__device__ __forceinline__ void ml_cover(int X0, int Y0, int X1, int Y1, int X2, int Y2, int px0, int px1, int py0, int py1,
                                         const float4* pos_s, uint2 rec, int e, int W, float invW, float invH,
                                         unsigned long long* __restrict__ zb, unsigned int* tbits, int tiles_x,
                                         int* qcount, uint2* queue) {
    const I dx0 = X2 - X1, dy0 = Y2 - Y1;
    const I dx1 = X0 - X2, dy1 = Y0 - Y2;
    const I dx2 = X1 - X0, dy2 = Y1 - Y0;
    // bias folds the tie rule into a strict comparison: inside <=> e + bias > 0
    const I b0 = ml_owns_edge(dx0, dy0) ? 1 : 0, b1 = ml_owns_edge(dx1, dy1) ? 1 : 0, b2 = ml_owns_edge(dx2, dy2) ? 1 : 0;
    const int Cx0 = px0 * 256 + 128;
    for (int py = py0; py <= py1; py++) {
        const int Cy = py * 256 + 128;
        I e0 = dx0 * (I)(Cy - Y1) - dy0 * (I)(Cx0 - X1);
        I e1 = dx1 * (I)(Cy - Y2) - dy1 * (I)(Cx0 - X2);
        I e2 = dx2 * (I)(Cy - Y0) - dy2 * (I)(Cx0 - X0);
        for (int px = px0; px <= px1; px++) {
            if (e0 + b0 > 0 && e1 + b1 > 0 && e2 + b2 > 0) {
                // the ~150-instruction depth resolve (and the tile bookkeeping) runs afterwards with the hits spread over
                // all lanes
                const int q = atomicAdd(qcount, 1);
                if (q < kFragQueue && FMHR_DCHECK(q >= 0, 0)) queue[q] = make_uint2((uint32_t)e, ((uint32_t)py << 16) | (uint32_t)px);
                else {
                    ml_mark_tile(tbits, tiles_x, px, py);
                    ml_resolve(pos_s, rec, px, py, W, invW, invH, zb);
                }
            }
            e0 -= dy0 * 256;
            e1 -= dy1 * 256;
            e2 -= dy2 * 256;
        }
    }
}

// Pixel-centre bounding box of a triangle from its snapped corners; false when a corner was rejected or no pixel centre
// lies inside (the common case for micropolygons).
struct MlBox { int px0, px1, py0, py1; bool small; };
__device__ __forceinline__ bool ml_bbox(const int2 s0, const int2 s1, const int2 s2, int H, int W, MlBox& bx) {
    if (s0.x == kSnapRejected || s1.x == kSnapRejected || s2.x == kSnapRejected) return false;
    const int minX = min(s0.x, min(s1.x, s2.x)), maxX = max(s0.x, max(s1.x, s2.x));
    const int minY = min(s0.y, min(s1.y, s2.y)), maxY = max(s0.y, max(s1.y, s2.y));
    // pixel centres (px*256+128) inside [min,max]
    bx.px0 = max(0, (minX - 128 + 255) >> 8); bx.px1 = min(W - 1, (maxX - 128) >> 8);
    bx.py0 = max(0, (minY - 128 + 255) >> 8); bx.py1 = min(H - 1, (maxY - 128) >> 8);
    bx.small = (maxX - minX) < 32768 && (maxY - minY) < 32768 && bx.px1 < 65536 && bx.py1 < 65536;
    return bx.px0 <= bx.px1 && bx.py0 <= bx.py1;
}

__device__ __forceinline__ void ml_candidate(uint2 rec, int e, const float4* pos_s, const int2* snap_s, int H, int W,
                                             float invW, float invH, unsigned long long* __restrict__ zb,
                                             unsigned int* tbits, int tiles_x, int* qcount, uint2* queue) {
    const int2 s0 = snap_s[rec.x & 1023u], s1 = snap_s[(rec.x >> 10) & 1023u], s2 = snap_s[(rec.x >> 20) & 1023u];
    MlBox bx;
    if (!ml_bbox(s0, s1, s2, H, W, bx)) return;
    int X0 = s0.x, Y0 = s0.y, X1 = s1.x, Y1 = s1.y, X2 = s2.x, Y2 = s2.y;
    bool neg;
    if (bx.small) {
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
    if (bx.small) ml_cover<int>(X0, Y0, X1, Y1, X2, Y2, bx.px0, bx.px1, bx.py0, bx.py1, pos_s, rec, e, W, invW, invH, zb, tbits, tiles_x, qcount, queue);
    else ml_cover<long long>(X0, Y0, X1, Y1, X2, Y2, bx.px0, bx.px1, bx.py0, bx.py1, pos_s, rec, e, W, invW, invH, zb, tbits, tiles_x, qcount, queue);
}

#ifndef FMHR_COV_THREADS
#define FMHR_COV_THREADS 256  // threads per coverage block (tuning: 128 gives smaller barrier domains, more blocks per SM)
#endif
constexpr int kCovThreads = FMHR_COV_THREADS;
// CLIP = true: stand-alone dr.rasterize (fmhr_rasterize_fwd_meshlets): `vg` is the caller's clip-space pos [N,clipV] (float4),
// no transform, and only the tile bitmap is produced (glist / gcount are NULL).
// DRAIN = true: variant for triangles of several pixels (chosen by the launcher when the frame has more than four pixels per
// triangle): the fragment queue is drained between candidate rounds instead of overflowing into the in-place resolve.
template <int TPT, bool CLIP = false, bool DRAIN = false>
__global__ void __launch_bounds__(kCovThreads, FMHR_LB_COVERAGE * (256 / kCovThreads)) ham_coverage_meshlet_kernel(
    const float4* __restrict__ vg, const float* __restrict__ viewM, const int32_t* __restrict__ ml_vptr,
    const int32_t* __restrict__ ml_verts, const uint2* __restrict__ ml_tri2, int max_verts, int H, int W, float invW,
    float invH, unsigned long long* __restrict__ zbuf, uint32_t* __restrict__ gbits, uint32_t* __restrict__ glist,
    int* __restrict__ gcount, int tiles_x, int tiles_per_view, int clipV) {
    FMHR_TRACE_SCOPE(4);
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    float4* pos_s = reinterpret_cast<float4*>(dyn_smem);
    int2* snap_s = reinterpret_cast<int2*>(pos_s + max_verts);
    const int words = (tiles_per_view + 31) >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned int* tbits = reinterpret_cast<unsigned int*>(snap_s + max_verts);  // block-wide tile bitmap
    __shared__ uint2 cand[kCovThreads / 32][TPT * 32];  // per warp: the triangles whose bounding box holds a pixel centre
    __shared__ uint32_t cbox[kCovThreads / 32][TPT * 32];  // small-box candidates: px0 | py0 << 14 | wide << 28 | tall << 29
    __shared__ uint2 queue[kCovThreads / 32][kFragQueue];  // per warp: covered pixel centres (candidate index, py << 16 | px)
    __shared__ int qcount[kCovThreads / 32];
    const int m = blockIdx.x, n = blockIdx.y;
    for (int w = threadIdx.x; w < words; w += blockDim.x) tbits[w] = 0u;  // visible after the vertex-phase barrier
    ViewM Mv;
    if (!CLIP) Mv = load_viewM(viewM + (size_t)n * kViewM);  // uniform loads: L1 broadcast
    // triangle records of this thread: issued before the vertex phase so their latency hides behind it
    uint2 rec[TPT];
    const uint2* recs = ml_tri2 + (size_t)m * (TPT * kCovThreads);
#pragma unroll
    for (int k = 0; k < TPT; k++) rec[k] = __ldg(recs + k * kCovThreads + threadIdx.x);
    const float hw = (float)W * 0.5f, hh = (float)H * 0.5f;
    if (CLIP) {
        const int vb = __ldg(ml_vptr + m), nv = __ldg(ml_vptr + m + 1) - vb;
        for (int i = threadIdx.x; i < nv; i += blockDim.x) {
            const float4 p = __ldg(vg + (size_t)n * clipV + (size_t)__ldg(ml_verts + vb + i));
            pos_s[i] = p;
            int X = kSnapRejected, Y = 0;
            if (!snap_vertex(p, hw, hh, X, Y)) X = kSnapRejected;
            snap_s[i] = make_int2(X, Y);
        }
    } else {
        // fused path: `vg` is the meshlet-ordered, padded copy of the step's vertices (ml_pos, written by the prologue): the
        // block's loads - triangle records above, these, the view matrix - are all independent of each other
        const float4* mp = vg + (size_t)m * max_verts;
        for (int i = threadIdx.x; i < max_verts; i += blockDim.x) {
            const float4 p = clip_from_world(Mv.m, __ldg(mp + i));
            pos_s[i] = p;
            int X = kSnapRejected, Y = 0;
            if (!snap_vertex(p, hw, hh, X, Y)) X = kSnapRejected;
            snap_s[i] = make_int2(X, Y);
        }
    }
    __syncthreads();
    // stage A: cheap bounding-box test of every triangle (all lanes busy), survivors compacted per warp.  Candidates whose
    // box holds at most 2 x 2 pixel centres (all but a handful on the subdivided mesh) go to the FRONT of the list with
    // their box, the others to the back.
    const unsigned lt = (1u << lane) - 1u;
    int n1 = 0, nb = 0;  // warp-uniform: small-box candidates (front), big-box candidates (back)
    const bool res_ok = W <= 16384 && H <= 16384;  // the packed box holds 14-bit pixel coordinates
#pragma unroll
    for (int k = 0; k < TPT; k++) {
        bool pass = false, fast = false;
        uint32_t box = 0;
        if (rec[k].x != 0xffffffffu) {  // padding
            MlBox bx;
            pass = ml_bbox(snap_s[rec[k].x & 1023u], snap_s[(rec[k].x >> 10) & 1023u], snap_s[(rec[k].x >> 20) & 1023u], H, W, bx);
            const int wide = bx.px1 - bx.px0, tall = bx.py1 - bx.py0;
            fast = pass && !DRAIN && bx.small && res_ok && wide <= 1 && tall <= 1;
            box = (uint32_t)bx.px0 | ((uint32_t)bx.py0 << 14) | ((uint32_t)wide << 28) | ((uint32_t)tall << 29);
        }
        const unsigned m1 = __ballot_sync(0xffffffffu, fast), mb = __ballot_sync(0xffffffffu, pass && !fast);
        if (fast) {
            const int at = n1 + __popc(m1 & lt);
            if (FMHR_DCHECK(at >= 0 && at < TPT * 32 - nb, 1)) {
                cand[warp][at] = rec[k];
                cbox[warp][at] = box;
            }
        } else if (pass && FMHR_DCHECK(TPT * 32 - 1 - nb - __popc(mb & lt) >= n1, 2)) cand[warp][TPT * 32 - 1 - nb - __popc(mb & lt)] = rec[k];
        n1 += __popc(m1);
        nb += __popc(mb);
    }
    __syncwarp();
    unsigned long long* zb = zbuf + (size_t)n * H * W;
    // stage B1: small-box candidates, one per lane, straight-line code: the (up to) four pixel centres of the box are
    // tested against the three 32-bit edge functions at once and the hits are appended to the warp's fragment queue at
    // positions from a warp prefix sum (no loops over the box, no shared atomics, no divergence).
    // Orientation: with E_i the edge functions of the ORIGINAL vertex order, the oriented triangle's are s * E_i and its
    // edge vectors s * (dx_i, dy_i), s = sign(area) (exact in integers: swapping two vertices reverses every edge);
    // inside <=> s * E_i + owns(s * dx_i, s * dy_i) > 0 for the three edges, as in ml_cover.
    int qn = 0;  // warp-uniform queue length
    for (int e0 = 0; e0 < n1; e0 += 32) {  // warp-uniform trip count
        const int e = e0 + lane;
        unsigned hits = 0;
        int px0 = 0, py0 = 0;
        uint2 r = make_uint2(0u, 0u);
        if (e < n1) {
            r = cand[warp][e];
            const uint32_t box = cbox[warp][e];
            const int2 s0 = snap_s[r.x & 1023u], s1 = snap_s[(r.x >> 10) & 1023u], s2 = snap_s[(r.x >> 20) & 1023u];
            px0 = (int)(box & 0x3fffu); py0 = (int)((box >> 14) & 0x3fffu);
            const int area = (s1.x - s0.x) * (s2.y - s0.y) - (s2.x - s0.x) * (s1.y - s0.y);  // |factors| < 2^15: exact
            const int sg = area < 0 ? -1 : 1;
            const int Cx = px0 * 256 + 128, Cy = py0 * 256 + 128;
            // edges 1->2 (at vertex 1), 2->0 (at vertex 2), 0->1 (at vertex 0), oriented
            const int dx0 = sg * (s2.x - s1.x), dy0 = sg * (s2.y - s1.y);
            const int dx1 = sg * (s0.x - s2.x), dy1 = sg * (s0.y - s2.y);
            const int dx2 = sg * (s1.x - s0.x), dy2 = sg * (s1.y - s0.y);
            // f_i >= 0 <=> inside w.r.t. edge i (the tie rule folded in: e + owns > 0 <=> e + owns - 1 >= 0)
            const int f0 = dx0 * (Cy - s1.y) - dy0 * (Cx - s1.x) + (ml_owns_edge(dx0, dy0) ? 0 : -1);
            const int f1 = dx1 * (Cy - s2.y) - dy1 * (Cx - s2.x) + (ml_owns_edge(dx1, dy1) ? 0 : -1);
            const int f2 = dx2 * (Cy - s0.y) - dy2 * (Cx - s0.x) + (ml_owns_edge(dx2, dy2) ? 0 : -1);
            // one pixel to the right: e -= 256 dy; one pixel down: e += 256 dx
            const int g0 = f0 - 256 * dy0, g1 = f1 - 256 * dy1, g2 = f2 - 256 * dy2;
            const int ux0 = 256 * dx0, ux1 = 256 * dx1, ux2 = 256 * dx2;
            const unsigned in00 = ~(unsigned)(f0 | f1 | f2) >> 31;
            const unsigned in10 = ~(unsigned)(g0 | g1 | g2) >> 31;
            const unsigned in01 = ~(unsigned)((f0 + ux0) | (f1 + ux1) | (f2 + ux2)) >> 31;
            const unsigned in11 = ~(unsigned)((g0 + ux0) | (g1 + ux1) | (g2 + ux2)) >> 31;
            const unsigned wide = (box >> 28) & 1u, tall = (box >> 29) & 1u;
            hits = in00 | ((in10 & wide) << 1) | ((in01 & tall) << 2) | ((in11 & wide & tall) << 3);
            if (area == 0) hits = 0;
        }
        // exclusive prefix sum of the hit counts over the warp
        const int cnt = __popc(hits);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        int at = qn + incl - cnt;
        qn += __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (hits & (1u << k)) {
                const int px = px0 + (k & 1), py = py0 + (k >> 1);
                if (at < kFragQueue && FMHR_DCHECK(at >= 0 && e < n1, 3)) queue[warp][at] = make_uint2((uint32_t)e, ((uint32_t)py << 16) | (uint32_t)px);
                else {  // (queue full: resolve in place)
                    ml_mark_tile(tbits, tiles_x, px, py);
                    ml_resolve(pos_s, r, px, py, W, invW, invH, zb);
                }
                at++;
            }
        }
    }
    if (lane == 0) qcount[warp] = min(qn, kFragQueue);
    __syncwarp();
    // stage B2: edge functions of the big-box candidates, spread over the lanes; hits go to the same queue
    if (!DRAIN || nb <= 64) {
        for (int e = TPT * 32 - 1 - lane; e >= TPT * 32 - nb; e -= 32)
            ml_candidate(cand[warp][e], e, pos_s, snap_s, H, W, invW, invH, zb, tbits, tiles_x, &qcount[warp], queue[warp]);
        __syncwarp();
    } else {
        // Triangles of a few pixels (configs 1, 3, 5: nearly every triangle is a candidate) fill the queue long before the
        // candidates run out, and what does not fit is resolved in place at one or two active lanes: drain the queue
        // between rounds whenever the next round might not fit.
        for (int e0 = 0; e0 < nb; e0 += 32) {  // warp-uniform trip count
            const int e = TPT * 32 - 1 - (e0 + lane);
            if (e0 + lane < nb)
                ml_candidate(cand[warp][e], e, pos_s, snap_s, H, W, invW, invH, zb, tbits, tiles_x, &qcount[warp], queue[warp]);
            __syncwarp();
            if (e0 + 32 < nb && qcount[warp] > kFragQueue - 96) {
                const int nq0 = min(qcount[warp], kFragQueue);
                for (int f = lane; f < nq0; f += 32) {
                    const uint2 fr = queue[warp][f];
                    const int px = (int)(fr.y & 0xffffu), py = (int)(fr.y >> 16);
                    ml_mark_tile(tbits, tiles_x, px, py);
                    ml_resolve(pos_s, cand[warp][fr.x], px, py, W, invW, invH, zb);
                }
                __syncwarp();
                if (lane == 0) qcount[warp] = 0;
                __syncwarp();
            }
        }
    }
    // stage C: tile bookkeeping + depth resolve of the hits, again with all lanes busy
    const int nq = min(qcount[warp], kFragQueue);
    for (int f = lane; f < nq; f += 32) {
        const uint2 fr = queue[warp][f];
        const int px = (int)(fr.y & 0xffffu), py = (int)(fr.y >> 16);
        ml_mark_tile(tbits, tiles_x, px, py);
        ml_resolve(pos_s, cand[warp][fr.x], px, py, W, invW, invH, zb);
    }
    __syncthreads();
    // flush: tiles this block touched first (global bitmap de-duplicates) are appended to the slot's work list
    for (int w = threadIdx.x; w < words; w += blockDim.x) {
        const unsigned int bits = tbits[w];
        if (bits == 0) continue;
        unsigned int fresh = bits & ~atomicOr(gbits + (size_t)n * words + w, bits);
        if (CLIP) continue;  // the stand-alone resolve pass walks the bitmap
        while (fresh) {
            const int bit = __ffs(fresh) - 1;
            fresh &= fresh - 1;
            const int tile = (w << 5) + bit;
            const int by = tile / tiles_x, bx = tile - by * tiles_x;
            const int gi_ = atomicAdd(gcount, 1);
            if (FMHR_DCHECK(gi_ >= 0 && gi_ < (int)gridDim.y * tiles_per_view && by < 1024 && bx < 1024, 9))
                glist[gi_] = ((uint32_t)n << 20) | ((uint32_t)by << 10) | (uint32_t)bx;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// pixel-domain helpers
// ------------------------------------------------------------------------------------------------
// Per-triangle record of the pixel passes (kTriRec floats = 160 B = five 32-byte sectors, rebuilt every iteration by
// ham_trirec_kernel after the normals): everything a pixel of triangle t needs sits in ONE contiguous record, so the
// shade pass issues 5 and the backward pass 4 independent 256-bit loads per pixel instead of 13 / 7 scattered sector
// gathers through vertex indices (these passes are bound by L1 wavefronts, not by DRAM).
//   [0..3]   i0, i1, i2 (int bits), flags (bit k: corner k has a degenerate normal; bit 4+k: wing k exists)
//   [4..12]  world positions of the corners      [13..21] unit normals       [22..30] albedo (b,g,r)
//   [31..39] world positions of the three opposite-wing vertices (antialias silhouette test; shade pass only)
constexpr int kTriRec = 40;
__global__ void __launch_bounds__(128) ham_trirec_kernel(const int32_t* __restrict__ tri, const int32_t* __restrict__ opp,
                                                         const float4* __restrict__ vg, const float4* __restrict__ vattr,
                                                         int V, int T, float4* __restrict__ trirec) {
    FMHR_TRACE_SCOPE(2);
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    int iv[3], ov[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { iv[k] = __ldg(tri + 3 * (size_t)t + k); ov[k] = __ldg(opp + 3 * (size_t)t + k); }
    float r[kTriRec];
    int flags = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const bool ok = (unsigned)iv[k] < (unsigned)V;
        const float4 p = ok ? vg[2 * (size_t)iv[k]] : make_float4(0.f, 0.f, 0.f, 0.f);
        F8 a;
        a.a = make_float4(0.f, 0.f, 0.f, 0.f); a.b = a.a;
        if (ok) a = ldg256(vattr + 2 * (size_t)iv[k]);
        r[4 + 3 * k] = p.x; r[5 + 3 * k] = p.y; r[6 + 3 * k] = p.z;
        r[13 + 3 * k] = a.a.x; r[14 + 3 * k] = a.a.y; r[15 + 3 * k] = a.a.z;
        r[22 + 3 * k] = a.b.x; r[23 + 3 * k] = a.b.y; r[24 + 3 * k] = a.b.z;
        if (a.a.w != 0.0f) flags |= 1 << k;
        const bool wing = (unsigned)ov[k] < (unsigned)V;
        const float4 w = wing ? vg[2 * (size_t)ov[k]] : make_float4(0.f, 0.f, 0.f, 0.f);
        r[31 + 3 * k] = w.x; r[32 + 3 * k] = w.y; r[33 + 3 * k] = w.z;
        if (wing) flags |= 16 << k;
    }
    r[0] = __int_as_float(iv[0]); r[1] = __int_as_float(iv[1]); r[2] = __int_as_float(iv[2]); r[3] = __int_as_float(flags);
    float4* out = trirec + (size_t)t * (kTriRec / 4);
#pragma unroll
    for (int k = 0; k < kTriRec / 8; k++) st256(out + 2 * k, r + 8 * k);  // five full-sector stores
}

struct PixTri {
    int i0, i1, i2, flags;
    float4 p0, p1, p2;  // clip positions
    float u, v;
    float3 n0, n1, n2;  // vertex normals
    float3 b0, b1, b2;  // vertex albedo
    float w0x;          // first float of the wing block (lives in the record's fourth sector)
};

__device__ __forceinline__ void load_pixtri(int t, int px, int py, const float4* __restrict__ trirec, const float* M,
                                            float invW, float invH, PixTri& q) {
    const float4* rp = trirec + (size_t)t * (kTriRec / 4);
    const F8 r0 = ldg256(rp), r1 = ldg256(rp + 2), r2 = ldg256(rp + 4), r3 = ldg256(rp + 6);
    q.i0 = __float_as_int(r0.a.x); q.i1 = __float_as_int(r0.a.y); q.i2 = __float_as_int(r0.a.z);
    q.flags = __float_as_int(r0.a.w);
    q.p0 = clip_from_world(M, make_float4(r0.b.x, r0.b.y, r0.b.z, 0.f));
    q.p1 = clip_from_world(M, make_float4(r0.b.w, r1.a.x, r1.a.y, 0.f));
    q.p2 = clip_from_world(M, make_float4(r1.a.z, r1.a.w, r1.b.x, 0.f));
    q.n0 = make_float3(r1.b.y, r1.b.z, r1.b.w);
    q.n1 = make_float3(r2.a.x, r2.a.y, r2.a.z);
    q.n2 = make_float3(r2.a.w, r2.b.x, r2.b.y);
    q.b0 = make_float3(r2.b.z, r2.b.w, r3.a.x);
    q.b1 = make_float3(r3.a.y, r3.a.z, r3.a.w);
    q.b2 = make_float3(r3.b.x, r3.b.y, r3.b.z);
    q.w0x = r3.b.w;
    const Bary b = bary_at(q.p0, q.p1, q.p2, px, py, invW, invH);
    q.u = b.u; q.v = b.v;
}

__device__ __forceinline__ float3 interp3(const float3 a0, const float3 a1, const float3 a2, const PixTri& q) {
    const float w = 1.0f - q.u - q.v;
    return make_float3(q.u * a0.x + q.v * a1.x + w * a2.x, q.u * a0.y + q.v * a1.y + w * a2.y,
                       q.u * a0.z + q.v * a1.z + w * a2.z);
}

// Tangent frame of a unit normal (branchless orthonormal basis, Duff et al. 2017).  The normalisation backward of a
// vertex normal only keeps the tangential part of its gradient (models/utils.py:547 F.normalize), so the pixel
// backward accumulates the two tangential components (t1.g, t2.g) instead of three Cartesian ones and the update pass
// rebuilds g_tangential = t1 (t1.g) + t2 (t2.g) from the identical frame: 8 floats = ONE 32-byte sector per vertex.
__device__ __forceinline__ void tangent_frame(const float3 n, float3& t1, float3& t2) {
    const float sg = copysignf(1.0f, n.z);
    const float a = -1.0f / (sg + n.z);
    const float b = n.x * n.y * a;
    t1 = make_float3(1.0f + sg * n.x * n.x * a, sg * b, -sg * n.x);
    t2 = make_float3(b, sg + n.y * n.y * a, -n.y);
}

// Window coordinates for the antialias rule straight from the world-space vertex (same arithmetic as aa_window_xy on
// the clip position every other kernel derives).
struct AAProjWorld {
    const float4* vg;
    const float* M;
    float xh, yh;
    __device__ __forceinline__ void operator()(int v, float fx, float fy, float& x, float& y) const {
        const float2 s = aa_window_xy(clip_from_world(M, __ldg(vg + 2 * (size_t)v)), xh, yh);
        x = xs(s.x, fx);
        y = xs(s.y, fy);
    }
};

// models/utils.py:208-226, same term order
__device__ __forceinline__ float sh_radiance(const float* c, float x, float y, float z) {
    float r = c[0];
    r = r + c[1] * y;
    r = r + c[2] * z;
    r = r + c[3] * x;
    r = r + c[4] * x * y;
    r = r + c[5] * y * z;
    r = r + c[6] * (2 * z * z - x * x - y * y);
    r = r + c[7] * z * x;
    r = r + c[8] * (x * x - y * y);
    return r;
}

// The scan pass walks 16x16 tiles (2-D locality keeps the hand's pixels in few, densely populated tiles); every later
// pixel pass walks the compact pixel lists the scan pass builds.  (kTile = 16 lives in common.cuh.)

// clip-space gradient (x, y, -, w) -> world-space xyz
__device__ __forceinline__ float3 clip_to_world(const float* M, float gx, float gy, float gw) {
    return make_float3(M[0] * gx + M[1] * gy + M[3] * gw, M[4] * gx + M[5] * gy + M[7] * gw,
                       M[8] * gx + M[9] * gy + M[11] * gw);
}

// One 16x16 tile of one view, as handed out by the work lists.
struct TileCtx {
    int n, bx, by, nx, ny;  // view slot, tile coordinates, tiles per row / column
};
__device__ __forceinline__ TileCtx tile_decode(uint32_t e, int nx, int ny) {
    TileCtx tc;
    tc.n = (int)(e >> 20); tc.by = (int)((e >> 10) & 1023u); tc.bx = (int)(e & 1023u); tc.nx = nx; tc.ny = ny;
    return tc;
}

// 32-way spread fp64 accumulators: shuffle-reduce inside the warp, one spread atomic per warp
__device__ __forceinline__ void warp_acc_add(double* acc, int k, float v) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0 && v != 0.0f) atomicAdd(acc + k * 32 + ((blockIdx.x * 8 + (threadIdx.x >> 5)) & 31), (double)v);
}
// Final flush of per-warp work-list buffers: the 8 warps of a block reserve their slots with ONE global atomic (every
// warp of the grid reaching its last flush at the same time meant ~5,000 same-address atomics queued at one L2 slice).
// Must be called by all threads of the block; returns this warp's base slot.
__device__ __forceinline__ int block_reserve(int* counter, int n_warp, int* s_cnt /* [10] shared */) {
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) s_cnt[wib] = n_warp;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < 8; w++) { const int c = s_cnt[w]; s_cnt[w] = tot; tot += c; }
        s_cnt[8] = tot > 0 ? atomicAdd(counter, tot) : 0;
    }
    __syncthreads();
    const int base = s_cnt[8] + s_cnt[wib];
    __syncthreads();  // s_cnt may be reused by the next reservation
    return base;
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// global pixel index -> view slot, coordinates
struct PixAddr { int n, px, py, rem; };
__device__ __forceinline__ PixAddr pix_decode(uint32_t pix, int H, int W) {
    PixAddr a;
    const uint32_t hw = (uint32_t)(H * W);
    a.n = (int)(pix / hw);
    a.rem = (int)(pix - (uint32_t)a.n * hw);
    a.py = a.rem / W;
    a.px = a.rem - a.py * W;
    return a;
}

// Per-vertex accumulator layout in `packed` (12V floats, viewed as 3V float4):
//   G8[2i]   = (photo_pos.xyz, normal.t1)   G8[2i+1] = (normal.t2, albedo.bgr)    one 32-byte sector per vertex, the
//                                                                                   only record the pixel backward hits
//   Gm[i] = packed4[2V + i] = (mask_pos.xyz, normal.z of degenerate vertices)       silhouette pairs only
// photo_* are gradients of the UN-NORMALISED photometric sum  sum |tmp_img - img|, mask_pos of sum (pred - valid)^2 / 2;
// normal.t1/t2 are the components of the vertex-normal gradient in tangent_frame(normal) (x, y for degenerate normals).

// ------------------------------------------------------------------------------------------------
// z-buffer key layout after the shade pass (low word): bits 0..27 triangle id, 28..30 silhouette-candidate bits of
// that triangle in this pixel's frame (aa_triangle_geom), bit 31 "valid" (covered and mask > 0).
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kTriMask = 0x0FFFFFFFu;

struct NbrKeys {
    int tri;   // -1 empty
    float zw;
    int bits;  // silhouette-candidate bits
    bool valid;
};
__device__ __forceinline__ NbrKeys decode_key(unsigned long long key) {
    NbrKeys k;
    if (key == ZB_EMPTY) { k.tri = -1; k.zw = 0.0f; k.bits = 0; k.valid = false; }
    else {
        const uint32_t lo = (uint32_t)key;
        k.tri = (int)(lo & kTriMask);
        k.bits = (int)((lo >> 28) & 7u);
        k.valid = (lo >> 31) != 0;
        k.zw = depth_from_key((uint32_t)(key >> 32));
    }
    return k;
}

// ------------------------------------------------------------------------------------------------
// scan: z-buffer tiles -> compact pixel lists.  Persistent over the tile list of the slot the coverage kernel just
// filled; each warp walks half tiles (16 x 8 pixels) with coalesced key loads and
//   * appends every covered pixel as (pixel index, triangle id) to `clist` (per-warp shared buffer, one global atomic per
//     ~200 pixels): the shade / antialias / backward passes then run with every lane busy and perfectly balanced;
//   * appends every EMPTY pixel that touches a covered one to `rlist` (the only empty pixels antialiasing can blend
//     into; a self-cleaning bitmap de-duplicates);
//   * resets the tiles the OTHER z-buffer slot dirtied in the previous iteration (no separate clear pass).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, FMHR_LB_SCAN) ham_scan_kernel(const unsigned long long* __restrict__ zbuf,
                                                       unsigned long long* __restrict__ zbuf_next,
                                                       const uint32_t* __restrict__ tlist, const int* __restrict__ tcount,
                                                       const uint32_t* __restrict__ tlist_next,
                                                       const int* __restrict__ tcount_next, int tiles_x, int tiles_y,
                                                       int H, int W, uint2* __restrict__ clist, int* __restrict__ ccount,
                                                       uint32_t* __restrict__ ringbits, uint32_t* __restrict__ rlist,
                                                       int* __restrict__ rcount, int rcap, int* __restrict__ status) {
    FMHR_TRACE_SCOPE(5);
    // unit of work = half a tile (16 x 8 pixels): a lane owns column (lane & 15) of rows (lane >> 4) + 2k, k = 0..3, so
    // the four key loads of a unit are independent and issued back to back
    constexpr int kBuf = 288, kRBuf = 96;
    __shared__ uint2 cbuf[8][kBuf];
    __shared__ uint32_t rbuf[8][kRBuf];  // ring pixels: appended to rlist with ONE global atomic per ~64 entries (a
                                         // same-address atomicAdd per ring pixel serialised the kernel at the L2 slice)
    int nbuf = 0, nring = 0;  // warp-uniform
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    auto flush = [&]() {
        int base = 0;
        if (lane == 0) base = atomicAdd(ccount, nbuf);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (int i = lane; i < nbuf; i += 32)
            if (FMHR_DCHECK(base >= 0 && i < kBuf, 8)) clist[base + i] = cbuf[wib][i];
        __syncwarp();
        nbuf = 0;
    };
    auto flush_ring = [&]() {
        int base = 0;
        if (lane == 0) base = atomicAdd(rcount, nring);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (int i = lane; i < nring; i += 32) {
            if (base + i < rcap) rlist[base + i] = rbuf[wib][i];
            else atomicOr(status, 2);
        }
        __syncwarp();
        nring = 0;
    };
    const int u0 = blockIdx.x * 8 + wib, ustride = gridDim.x * 8;
    const int lx = lane & 15, ly = lane >> 4;
    // A warp owns only one or two units of each loop, so the kernel's length is the length of its dependent load chains:
    // both list lengths and the first tile entries of BOTH loops are requested up front, every unit's next tile entry is
    // requested before its keys are consumed, and the border keys travel with the unit's own keys (one level instead of two).
    const int nd = *tcount_next;
    const int nt = *tcount;
    uint32_t te_next = u0 < 2 * nt ? __ldg(tlist + (u0 >> 1)) : 0u;
    for (int u = u0; u < 2 * nd; u += ustride) {
        const TileCtx tc = tile_decode(__ldg(tlist_next + (u >> 1)), tiles_x, tiles_y);
        const int px = tc.bx * kTile + lx, py0 = tc.by * kTile + (u & 1) * 8 + ly;
        if (px < W) {
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (py0 + 2 * k < H) zbuf_next[((size_t)tc.n * H + py0 + 2 * k) * W + px] = ZB_EMPTY;
        }
    }
    for (int u = u0; u < 2 * nt; u += ustride) {
        const TileCtx tc = tile_decode(te_next, tiles_x, tiles_y);
        if (u + ustride < 2 * nt) te_next = __ldg(tlist + ((u + ustride) >> 1));
        const int px = tc.bx * kTile + lx, py0 = tc.by * kTile + (u & 1) * 8 + ly;
        const unsigned long long* zb = zbuf + (size_t)tc.n * H * W;
        unsigned long long key[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int py = py0 + 2 * k;
            key[k] = (px < W && py < H) ? zb[(size_t)py * W + px] : ZB_EMPTY;
        }
        // Emptiness of the 4-neighbours of the covered pixels.  Inside the 16x8 unit it is in the ballot words (bit lx +
        // 16 * (row & 1) of m[row >> 1]); only the unit's border needs keys from memory: ONE load for the rows above /
        // below (lanes 0-15 / 16-31) and ONE for the columns left / right (lanes 0-7 / 8-15) instead of 16 per lane.
        const int uy0 = tc.by * kTile + (u & 1) * 8;  // first row of the unit
        const int qy = ly == 0 ? uy0 - 1 : uy0 + 8;
        const bool in = px < W && qy >= 0 && qy < H;
        const unsigned long long kq = in ? zb[(size_t)qy * W + px] : 0ull;
        const int qx = lane < 8 ? tc.bx * kTile - 1 : tc.bx * kTile + 16, qr = uy0 + (lane & 7);
        const bool in2 = lane < 16 && qx >= 0 && qx < W && qr < H;
        const unsigned long long kq2 = in2 ? zb[(size_t)qr * W + qx] : 0ull;
        unsigned m[4];
        int total = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) { m[k] = __ballot_sync(0xffffffffu, key[k] != ZB_EMPTY); total += __popc(m[k]); }
        if (total == 0) continue;
        const unsigned e_tb = __ballot_sync(0xffffffffu, in && kq == ZB_EMPTY);
        const unsigned e_lr = __ballot_sync(0xffffffffu, in2 && kq2 == ZB_EMPTY);
        unsigned emask = 0u;  // bit 4k + d: neighbour d (right, left, down, up) of this lane's pixel in row 2k + ly is empty
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (key[k] == ZB_EMPTY) continue;
            const int r = 2 * k + ly, py = py0 + 2 * k;
            const unsigned row = m[k] >> (16 * ly);                       // coverage bits of this pixel's row
            // row below / above inside the unit: other half of m[k], or the neighbouring word
            const unsigned below = ly == 0 ? (m[k] >> 16) : (k < 3 ? m[k < 3 ? k + 1 : 3] : 0u);
            const unsigned above = ly == 1 ? m[k] : (k > 0 ? (m[k > 0 ? k - 1 : 0] >> 16) : 0u);
            const bool er = lx < 15 ? (!((row >> (lx + 1)) & 1u) && px + 1 < W) : ((e_lr >> (8 + r)) & 1u);
            const bool el = lx > 0 ? !((row >> (lx - 1)) & 1u) : ((e_lr >> r) & 1u);
            const bool ed = r < 7 ? (!((below >> lx) & 1u) && py + 1 < H) : ((e_tb >> (16 + lx)) & 1u);
            const bool eu = r > 0 ? !((above >> lx) & 1u) : ((e_tb >> lx) & 1u);
            emask |= ((er ? 1u : 0u) | (el ? 2u : 0u) | (ed ? 4u : 0u) | (eu ? 8u : 0u)) << (4 * k);
        }
        const bool any_ring = __any_sync(0xffffffffu, emask != 0u);
        int off = nbuf;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int py = py0 + 2 * k;
            if (key[k] != ZB_EMPTY) {
                const uint32_t pix = (uint32_t)(((size_t)tc.n * H + py) * W + px);
                if (FMHR_DCHECK(off + __popc(m[k] & lt) < kBuf, 4)) cbuf[wib][off + __popc(m[k] & lt)] = make_uint2(pix, (uint32_t)key[k] & kTriMask);
            }
            off += __popc(m[k]);
        }
        if (any_ring) {  // boundary unit: empty 4-neighbours of covered pixels = the ring antialiasing can blend into
            // the bitmap de-duplicates; all (up to 16) atomics of a lane are issued before the first result is consumed
            // (one after the other they were 16 dependent L2 round trips per unit)
            unsigned fresh = 0u;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                if ((emask >> j) & 1u) {
                    const int k = j >> 2, d = j & 3;
                    const int qx = px + (d == 0 ? 1 : (d == 1 ? -1 : 0)), qy = py0 + 2 * k + (d == 2 ? 1 : (d == 3 ? -1 : 0));
                    const uint32_t q = (uint32_t)(((size_t)tc.n * H + qy) * W + qx);
                    const uint32_t bit = 1u << (q & 31);
                    if (!(atomicOr(ringbits + (q >> 5), bit) & bit)) fresh |= 1u << j;
                }
            }
            if (__any_sync(0xffffffffu, fresh != 0u)) {
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const bool f = (fresh >> j) & 1u;
                    const unsigned mr = __ballot_sync(0xffffffffu, f);
                    if (mr) {
                        if (f) {
                            const int k = j >> 2, d = j & 3;
                            const int qx = px + (d == 0 ? 1 : (d == 1 ? -1 : 0)), qy = py0 + 2 * k + (d == 2 ? 1 : (d == 3 ? -1 : 0));
                            if (FMHR_DCHECK(nring + __popc(mr & lt) < kRBuf && qx >= 0 && qx < W && qy >= 0 && qy < H, 5))
                                rbuf[wib][nring + __popc(mr & lt)] = (uint32_t)(((size_t)tc.n * H + qy) * W + qx);
                        }
                        nring += __popc(mr);
                        __syncwarp();
                        if (nring > kRBuf - 32) flush_ring();
                    }
                }
            }
        }
        nbuf = off;
        __syncwarp();
        if (nbuf > kBuf - 128) flush();
    }
    // last flush of both lists: ONE reservation phase (the two block-level atomics are issued by two threads side by side,
    // not as two dependent round trips at the end of the kernel's critical chain)
    __shared__ int s_cnt[20];
    if (lane == 0) { s_cnt[wib] = nbuf; s_cnt[10 + wib] = nring; }
    __syncthreads();
    if (threadIdx.x == 0 || threadIdx.x == 32) {
        int* c = s_cnt + (threadIdx.x == 0 ? 0 : 10);
        int tot = 0;
        for (int w = 0; w < 8; w++) { const int v = c[w]; c[w] = tot; tot += v; }
        c[8] = tot > 0 ? atomicAdd(threadIdx.x == 0 ? ccount : rcount, tot) : 0;
    }
    __syncthreads();
    {
        const int base = s_cnt[8] + s_cnt[wib];
        for (int i = lane; i < nbuf; i += 32) clist[base + i] = cbuf[wib][i];
    }
    {
        const int base = s_cnt[18] + s_cnt[10 + wib];
        for (int i = lane; i < nring; i += 32) {
            if (base + i < rcap) rlist[base + i] = rbuf[wib][i];
            else atomicOr(status, 2);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// shade: one covered pixel per thread over the compact list (grid-stride, every lane busy): triangle -> world-space
// corners -> clip positions, perspective barycentrics, interpolated normal / albedo, SH radiance * albedo (phase B) or
// normals + albedo planes (phase A); tags the z-buffer key with the triangle's silhouette-candidate bits and the valid
// flag (one 32-bit RED), and records the valid flag in the list entry for the backward pass.
// ------------------------------------------------------------------------------------------------
struct DeferredScatter { int vi[3]; float4 lo[3], hi[3]; };  // corner records of one pixel's backward, vi < 0: none
template <int PHASE, bool DEFER>
__device__ __forceinline__ void pixel_backward_q(const PixTri& q, int px, int py, float4 gin, const float* __restrict__ M,
                                                 const float* __restrict__ c, int V, int H, int W, float4* __restrict__ G,
                                                 DeferredScatter* def);
__device__ __forceinline__ void scatter_sector_pairs(const DeferredScatter& d, float4* __restrict__ G);

// BWD (phase B training step): the pixel's own loss gradient is known right here for every pixel that does not RECEIVE an
// antialias blend - tmp_img == pred_img there, so d|tmp - img| = sign(pred - img) - and those are > 95 % of the pixels.
// The pass therefore speculates sign(pred - img) for every valid pixel, back-propagates it immediately while the
// triangle record, barycentrics and shading terms are still in registers (no second gather of the record, no pixel
// backward pass), and leaves the three signs in the colour plane's w channel; the antialias pass compares them with the
// signs of the antialiased image and queues the few pixels that differ for a linear correction (pair kernel).
__device__ __forceinline__ float sgnf(float d) { return (float)((d > 0.f) - (d < 0.f)); }
__device__ __forceinline__ uint32_t pack_signs(float s0, float s1, float s2) {  // 2 bits per channel + a marker bit
    return (uint32_t)((int)s0 + 1) | ((uint32_t)((int)s1 + 1) << 2) | ((uint32_t)((int)s2 + 1) << 4) | 0x40u;
}
template <int PHASE, bool BWD = false>
__global__ void __launch_bounds__(256, BWD ? FMHR_LB_SHADE_BWD : FMHR_LB_SHADE) ham_shade_kernel(uint2* __restrict__ clist, const int* __restrict__ ccount,
                                                        unsigned long long* __restrict__ zbuf,
                                                        const float4* __restrict__ vg, const float* __restrict__ viewM,
                                                        float invW, float invH, const float4* __restrict__ trirec,
                                                        const float* __restrict__ masks,
                                                        const float* __restrict__ sh_coeffs,
                                                        const int32_t* __restrict__ view_idx,
                                                        const int32_t* __restrict__ sh_idx, int V, int H, int W,
                                                        float4* __restrict__ plane0, float4* __restrict__ plane1,
                                                        double* __restrict__ acc, const float* __restrict__ imgs,
                                                        float4* __restrict__ G) {
    FMHR_TRACE_SCOPE(6);
    const int nc = *ccount;
    const int hw = H * W;
    float nvalid = 0.0f;
    // the NEXT list entry is loaded before this entry's gather chain starts (one dependent level less per iteration)
    const int e_stride = gridDim.x * blockDim.x;
    const int e_first = (FMHR_INTERLEAVE & 1) ? (int)(((threadIdx.x >> 5) * gridDim.x + blockIdx.x) * 32 + (threadIdx.x & 31))
                                              : (int)(blockIdx.x * blockDim.x + threadIdx.x);
    uint2 ent_next = e_first < nc ? clist[e_first] : make_uint2(0u, 0u);
#if FMHR_PAIR_SECTORS
    // (the scatter at the end of the body exchanges records between lane pairs: the trip count is the warp's, not the lane's)
    for (int e = e_first; e - (int)(threadIdx.x & 31) < nc; e += e_stride) {
        DeferredScatter dsc;
        dsc.vi[0] = dsc.vi[1] = dsc.vi[2] = -1;
        if (e < nc) {
#else
    for (int e = e_first; e < nc; e += e_stride) {
        {
#endif
        const uint2 ent = ent_next;
        if (e + e_stride < nc) ent_next = clist[e + e_stride];
        const size_t pix = ent.x;
        const int t = (int)ent.y;
        const PixAddr pa = pix_decode(ent.x, H, W);
        const int n = pa.n, px = pa.px, py = pa.py;
        const float* Mv = viewM + (size_t)n * kViewM;
        const int view = __ldg(view_idx + n);
        const bool valid = __ldg(masks + (size_t)view * hw + pa.rem) > 0.0f;
        // BWD: the target pixel is needed only after the whole shading chain - issue its loads now, beside the mask and the
        // triangle record, instead of as one more dependent DRAM round trip at the end of the iteration
        float tg0 = 0.f, tg1 = 0.f, tg2 = 0.f;
        if (BWD) {
            const float* img = imgs + ((size_t)view * hw + pa.rem) * 3;
            tg0 = __ldg(img); tg1 = __ldg(img + 1); tg2 = __ldg(img + 2);
        }
        PixTri q;
        load_pixtri(t, px, py, trirec, Mv, invW, invH, q);
        AAGeom g;
        g.bits = 0;
        {
            // silhouette-candidate bits of this triangle in this pixel's frame: window coordinates of the corners come
            // from the clip positions already in registers, the three wing vertices from the record's last sector
            const F8 r4 = ldg256(trirec + (size_t)t * (kTriRec / 4) + 8);
            const float xh = 0.5f * (float)W, yh = 0.5f * (float)H;
            const float2 so0 = aa_window_xy(clip_from_world(Mv, make_float4(q.w0x, r4.a.x, r4.a.y, 0.f)), xh, yh);
            const float2 so1 = aa_window_xy(clip_from_world(Mv, make_float4(r4.a.z, r4.a.w, r4.b.x, 0.f)), xh, yh);
            const float2 so2 = aa_window_xy(clip_from_world(Mv, make_float4(r4.b.y, r4.b.z, r4.b.w, 0.f)), xh, yh);
            // aa_triangle_geom_win ignores a wing whose vertex id is outside [0, V)
            const int o0 = (q.flags & 16) ? 0 : -1, o1 = (q.flags & 32) ? 0 : -1, o2 = (q.flags & 64) ? 0 : -1;
            aa_triangle_geom_win(q.i0, q.i1, q.i2, aa_window_xy(q.p0, xh, yh), aa_window_xy(q.p1, xh, yh),
                                 aa_window_xy(q.p2, xh, yh), o0, o1, o2, so0, so1, so2, px, py, V, H, W, g);
        }
        const float3 m = interp3(q.n0, q.n1, q.n2, q);
        const float3 a = interp3(q.b0, q.b1, q.b2, q);
        if (valid) nvalid += 1.0f;
        const uint32_t tag = ((uint32_t)g.bits << 28) | (valid ? 0x80000000u : 0u);
        if (tag) atomicOr(reinterpret_cast<unsigned int*>(zbuf + pix), tag);  // low word of the little-endian key
        if (!BWD) clist[e].y = ent.y | (valid ? 0x80000000u : 0u);  // (the separate backward pass reads the flag)
        if (PHASE == 1) {
            float4 col = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {
                const float* sh = sh_coeffs + (size_t)__ldg(sh_idx + n) * 9;
                const float inv = 1.0f / fmaxf(sqrtf(m.x * m.x + m.y * m.y + m.z * m.z), 1e-12f);
                const float r = sh_radiance(sh, m.x * inv, m.y * inv, m.z * inv);
                col = make_float4(r * a.x, r * a.y, r * a.z, 1.0f);
                if (BWD) {
                    const float s0 = sgnf(col.x - tg0), s1 = sgnf(col.y - tg1), s2 = sgnf(col.z - tg2);
                    col.w = __uint_as_float(pack_signs(s0, s1, s2));
#if FMHR_PAIR_SECTORS
                    if (s0 != 0.f || s1 != 0.f || s2 != 0.f)
                        pixel_backward_q<1, true>(q, px, py, make_float4(s0, s1, s2, 0.f), Mv, sh, V, H, W, G, &dsc);
#else
                    pixel_backward_q<1, false>(q, px, py, make_float4(s0, s1, s2, 0.f), Mv, sh, V, H, W, G, nullptr);
#endif
                }
            }
            plane0[pix] = col;
        } else {
            plane0[pix] = make_float4(m.x, m.y, m.z, valid ? 1.0f : 0.0f);
            plane1[pix] = make_float4(a.x, a.y, a.z, 0.0f);
        }
        }
#if FMHR_PAIR_SECTORS
        if (BWD) scatter_sector_pairs(dsc, G);
#endif
    }
    warp_acc_add(acc, 0, nvalid);
}

// ------------------------------------------------------------------------------------------------
// antialias (gather form) + losses over the compact lists: covered pixels, then the ring of empty pixels around them.
// Each pixel inspects its four pixel pairs (self,right) (self,down) (left,self) (up,self).  Pairs that can blend
// (different triangle ids and silhouette bits set on the chosen triangle - exact, the bits come from the identical
// geometry call in the shade pass) are rare (~1 % of pixels), so instead of running the ~400-instruction edge analysis
// under a 1-3 lane mask they are queued per warp in shared memory and analysed with the queue spread over the lanes;
// a pair is analysed from both of its pixels and each takes the blend only if it is the receiver
// (out[alpha > 0 ? first : second] += alpha * (color[second] - color[first])), the pair's first pixel records it for
// the backward pass.  Everything is warp-synchronous.
//   item = (lane << 2) | which,  which: 0 (self,right)  1 (self,down)  2 (left,self)  3 (up,self)
// The mask loss sum_pixels (pred_mask - valid_mask)^2 is accumulated as a correction (pred - valid)^2 - valid^2 over the
// listed pixels; every other pixel has pred_mask == 0 and contributes valid_mask^2, whose total per view is a constant
// of the optimisation (buffers.view_vm2).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool pair_needs_analysis(const NbrKeys& k0, const NbrKeys& k1) {
    if (k0.tri == k1.tri) return false;
    // same triangle choice as aa_analyse
    const bool from1 = (k0.tri >= 0 && k1.tri >= 0) ? !(k0.zw < k1.zw) : (k0.tri < 0);
    return (from1 ? k1.bits : k0.bits) != 0;
}

// SPEC (phase B training step): the shade pass already back-propagated sign(pred - img) of every valid pixel (packed in
// plane0.w); this pass only queues the pixels whose antialiased sign differs (qlist) and writes the pixel-gradient plane
// for the pixels that received a blend (the only entries the pair kernel reads).
template <int PHASE, bool SPEC = false>
__global__ void __launch_bounds__(256, FMHR_LB_AA) ham_aa_loss_kernel(
    const uint2* __restrict__ clist, const int* __restrict__ ccount, const uint32_t* __restrict__ rlist,
    const int* __restrict__ rcount, int rcap, uint32_t* __restrict__ ringbits, const unsigned long long* __restrict__ zbuf,
    const float4* __restrict__ vg, const float* __restrict__ viewM, const int32_t* __restrict__ tri,
    const int32_t* __restrict__ opp, const float* __restrict__ imgs, const float* __restrict__ valid_masks,
    const float* __restrict__ sh_coeffs, const int32_t* __restrict__ view_idx, const int32_t* __restrict__ sh_idx,
    int V, int T, int H, int W, const float4* __restrict__ plane0, const float4* __restrict__ plane1,
    float4* __restrict__ gplane0, float4* __restrict__ gplane1, double* __restrict__ acc, float* __restrict__ gsh,
    float* __restrict__ dbg_image, float* __restrict__ dbg_mask, uint4* __restrict__ plist_a,
    uint32_t* __restrict__ plist_b, int* __restrict__ pcount, int pcap, int* __restrict__ status,
    double* __restrict__ init_acc, uint2* __restrict__ qlist, int* __restrict__ qcount, int qcap) {
    FMHR_TRACE_SCOPE(7);
    // PHASE 2 = HAM initialisation (mesh_sfs_optim.py:124-163): plane0 holds interpolated NORMALS, they and the coverage
    // are antialiased like phase B's colour + coverage; `imgs` is the gray image [num,H,W]; outputs: antialiased coverage
    // (dbg_mask), unit antialiased normal (gplane0) and the per-view normal equations of the SH fit (init_acc).
    constexpr int NC = PHASE != 0 ? 4 : 6;  // blended channels: (b,g,r | normal, coverage) or (normal xyz, albedo bgr)
    __shared__ uint32_t q_items_s[8][128];
    __shared__ int q_n_s[8];
    __shared__ float blend_s[8][32][NC];
    __shared__ uint32_t spix_s[8][32];
    __shared__ uint32_t recv_s[8];  // SPEC: bit L = lane L's pixel received a blend in this batch
    // recorded pairs wait in shared memory and reach plist with ONE global atomic per >= 32 records (a same-address
    // atomicAdd per pair serialised the kernel at the L2 slice)
    constexpr int kPBuf = 64;
    __shared__ uint4 pbuf_a[8][kPBuf];
    __shared__ uint32_t pbuf_b[8][kPBuf];
    int npb = 0;  // warp-uniform
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    auto flush_pairs = [&]() {
        int base = 0;
        if (lane == 0) base = atomicAdd(pcount, npb);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (int i = lane; i < npb; i += 32) {
            if (base + i < pcap) { plist_a[base + i] = pbuf_a[wib][i]; plist_b[base + i] = pbuf_b[wib][i]; }
            else atomicOr(status, 1);
        }
        __syncwarp();
        npb = 0;
    };
    uint32_t* q_items = q_items_s[wib];
    int* q_n = &q_n_s[wib];
    float (*blend)[NC] = blend_s[wib];
    uint32_t* spix = spix_s[wib];
    const int nc = *ccount, total = nc + min(*rcount, rcap);
    const int hw = H * W;
    float abs_acc = 0.0f;
    double msk_acc = 0.0;  // fp64: the corrections cancel against the fp64 view totals (exactly 0 when pred == valid)
    auto list_pixel = [&](int e) -> uint32_t { return e < nc ? clist[e].x : rlist[e - nc]; };
    const int e_stride = gridDim.x * blockDim.x;
    uint32_t pix_next = 0u;  // the NEXT list entry is loaded before this batch's key / analysis chain starts
    const int e_first = (FMHR_INTERLEAVE & 2) ? (int)((wib * gridDim.x + blockIdx.x) * 32 + lane)
                                              : (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (e_first < total) pix_next = list_pixel(e_first);
    for (int e0 = e_first - lane; e0 < total; e0 += e_stride) {
        const int e = e0 + lane;
        const bool active = e < total;
        const uint32_t pix32 = active ? pix_next : 0u;
        if (e + e_stride < total) pix_next = list_pixel(e + e_stride);
        if (active && e >= nc) atomicAnd(ringbits + (pix32 >> 5), ~(1u << (pix32 & 31)));  // the bitmap cleans itself
        const PixAddr pa = pix_decode(pix32, H, W);
        const int n = pa.n, px = pa.px, py = pa.py, rem = pa.rem;
        const size_t base = (size_t)n * hw;
        const unsigned long long* zb = zbuf + base;
        NbrKeys self = decode_key(ZB_EMPTY);
#if FMHR_AA_HOIST
        // loads whose addresses only depend on the pixel (own colour, target, valid_mask) are issued before the key /
        // pair-analysis chain instead of behind it
        float4 h_c0 = make_float4(0.f, 0.f, 0.f, 0.f);
        float h_t0 = 0.f, h_t1 = 0.f, h_t2 = 0.f, h_vm = 0.f;
        if (PHASE == 1 && active) {
            const int hview = __ldg(view_idx + n);
            h_c0 = plane0[pix32];
            const float* himg = imgs + ((size_t)hview * hw + rem) * 3;
            h_t0 = __ldg(himg); h_t1 = __ldg(himg + 1); h_t2 = __ldg(himg + 2);
            h_vm = __ldg(valid_masks + (size_t)hview * hw + rem);
        }
#endif
        if (lane == 0) { *q_n = 0; if (SPEC) recv_s[wib] = 0u; }
        spix[lane] = pix32;
        __syncwarp();
        if (active) {
            self = decode_key(zb[rem]);
            if (px + 1 < W && pair_needs_analysis(self, decode_key(zb[rem + 1]))) q_items[atomicAdd(q_n, 1)] = ((uint32_t)lane << 2) | 0u;
            if (py + 1 < H && pair_needs_analysis(self, decode_key(zb[rem + W]))) q_items[atomicAdd(q_n, 1)] = ((uint32_t)lane << 2) | 1u;
            if (px > 0 && pair_needs_analysis(decode_key(zb[rem - 1]), self)) q_items[atomicAdd(q_n, 1)] = ((uint32_t)lane << 2) | 2u;
            if (py > 0 && pair_needs_analysis(decode_key(zb[rem - W]), self)) q_items[atomicAdd(q_n, 1)] = ((uint32_t)lane << 2) | 3u;
        }
        __syncwarp();
        // analysis of the queued pairs spread over the lanes; the receiver's blend lands in shared memory
        const int nq = *q_n;
        if (nq > 0) {  // warp-uniform
#pragma unroll
            for (int c = 0; c < NC; c++) blend[lane][c] = 0.0f;
            __syncwarp();
            for (int x0 = 0; x0 < nq; x0 += 32) {  // warp-uniform trip count: the record buffer is appended by ballot
                const int x = x0 + lane;
                bool found = false, record = false;
                AAPair pr;
                int L = 0, which = 0, d = 0, r0 = 0, r1 = 0;
                size_t qbase = 0;
                NbrKeys k0 = decode_key(ZB_EMPTY), k1 = k0;
                if (x < nq) {
                    const uint32_t item = q_items[x];
                    L = (int)(item >> 2); which = (int)(item & 3u); d = which & 1;
                    const PixAddr qa = pix_decode(spix[L], H, W);
                    // first pixel of the pair
                    const int qx = which < 2 ? qa.px : qa.px - (1 - d), qy = which < 2 ? qa.py : qa.py - d;
                    qbase = (size_t)qa.n * hw;
                    r0 = qy * W + qx; r1 = r0 + (d ? W : 1);
                    k0 = decode_key(zbuf[qbase + r0]); k1 = decode_key(zbuf[qbase + r1]);
                    const AAProjWorld proj{vg, viewM + (size_t)qa.n * kViewM, 0.5f * (float)W, 0.5f * (float)H};
                    found = aa_analyse_bits(k0.tri, k0.zw, k0.bits, k1.tri, k1.zw, k1.bits, qx, qy, d, proj, tri, V, T, H, W, pr);
                    record = found && which < 2 && PHASE != 2;  // seen from the pair's first pixel: record it for the backward pass
                }
                const unsigned mrec = __ballot_sync(0xffffffffu, record);
                if (mrec) {
                    if (record) {
                        const int slot = npb + __popc(mrec & lt);
                        const uint32_t flags = (uint32_t)d | ((uint32_t)pr.from1 << 1) | ((uint32_t)pr.clamped << 2) |
                                               ((uint32_t)pr.di << 3);
                        if (FMHR_DCHECK(slot >= 0 && slot < kPBuf, 6)) {
                            pbuf_a[wib][slot] = make_uint4((uint32_t)(qbase + r0), flags, __float_as_uint(pr.alpha), (uint32_t)pr.i1);
                            pbuf_b[wib][slot] = (uint32_t)pr.i2;
                        }
                    }
                    npb += __popc(mrec);
                    __syncwarp();
                    if (npb > kPBuf - 32) flush_pairs();
                }
                if (!found) continue;
                const bool recv_is_self = (which < 2) == (pr.alpha > 0.0f);
                if (!recv_is_self) continue;  // the other pixel's own item delivers it
                if (SPEC) atomicOr(&recv_s[wib], 1u << L);
                // out[recv] += alpha * (color[second] - color[first]); empty pixels are zero in every channel
                float4 f0 = make_float4(0.f, 0.f, 0.f, 0.f), f1 = f0, s0 = f0, s1 = f0;
                if (k0.tri >= 0) { f0 = plane0[qbase + r0]; if (PHASE == 0) f1 = plane1[qbase + r0]; }
                if (k1.tri >= 0) { s0 = plane0[qbase + r1]; if (PHASE == 0) s1 = plane1[qbase + r1]; }
                atomicAdd(&blend[L][0], pr.alpha * (s0.x - f0.x));
                atomicAdd(&blend[L][1], pr.alpha * (s0.y - f0.y));
                atomicAdd(&blend[L][2], pr.alpha * (s0.z - f0.z));
                if (PHASE != 0) {
                    atomicAdd(&blend[L][3], pr.alpha * ((k1.tri >= 0 ? 1.0f : 0.0f) - (k0.tri >= 0 ? 1.0f : 0.0f)));
                } else {
                    atomicAdd(&blend[L][3], pr.alpha * (s1.x - f1.x));
                    atomicAdd(&blend[L][4], pr.alpha * (s1.y - f1.y));
                    atomicAdd(&blend[L][5], pr.alpha * (s1.z - f1.z));
                }
            }
            __syncwarp();
        }
        float gc[9];
#pragma unroll
        for (int k = 0; k < 9; k++) gc[k] = 0.0f;
        float init_y = 0.0f;  // PHASE 2: gc = SH basis row of this valid pixel, init_y = its gray value
        bool init_valid = false;
        uint32_t corr_pack = 0u;  // SPEC: sign deltas of this pixel against the shade pass' speculation (0 = none)
        if (active) {
            const size_t pix = pix32;
            const int view = __ldg(view_idx + n);
            // own (pre-antialias) values; empty pixels are zero in every channel
            float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (self.tri >= 0) {
#if FMHR_AA_HOIST
                c0 = PHASE == 1 ? h_c0 : plane0[pix];
#else
                c0 = plane0[pix];
#endif
                if (PHASE == 0) c1 = plane1[pix];
            }
            float4 a0 = c0, a1 = c1;  // antialiased values
            float amask = self.tri >= 0 ? 1.0f : 0.0f;
            if (nq > 0) {
                a0.x += blend[lane][0]; a0.y += blend[lane][1]; a0.z += blend[lane][2];
                if (PHASE != 0) amask += blend[lane][3];
                else { a1.x += blend[lane][3]; a1.y += blend[lane][4]; a1.z += blend[lane][5]; }
            }
            const bool valid = self.valid;
            const float* img = imgs + ((size_t)view * hw + rem) * 3;
            if (PHASE == 2) {
                float4 nrm = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) {  // mesh_sfs_optim.py:148-153: F.normalize of the antialiased normals, get_matrix row
                    const float len = sqrtf(a0.x * a0.x + a0.y * a0.y + a0.z * a0.z);
                    const float inv = 1.0f / fmaxf(len, 1e-12f);
                    const float nx = a0.x * inv, ny = a0.y * inv, nz = a0.z * inv;
                    nrm = make_float4(nx, ny, nz, 1.0f);
                    gc[0] = 1.0f; gc[1] = ny; gc[2] = nz; gc[3] = nx; gc[4] = nx * ny; gc[5] = ny * nz;
                    gc[6] = 2 * nz * nz - nx * nx - ny * ny; gc[7] = nz * nx; gc[8] = nx * nx - ny * ny;
                    init_y = __ldg(imgs + (size_t)view * hw + rem);
                    init_valid = true;
                }
                gplane0[pix] = nrm;
                if (dbg_mask) dbg_mask[pix] = amask;
            } else if (PHASE == 1) {
                float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) {  // mesh_sfs_optim.py:289  l1 over tmp_img[valid_idx]
#if FMHR_AA_HOIST
                    const float d0 = a0.x - h_t0, d1 = a0.y - h_t1, d2 = a0.z - h_t2;
#else
                    const float d0 = a0.x - __ldg(img), d1 = a0.y - __ldg(img + 1), d2 = a0.z - __ldg(img + 2);
#endif
                    abs_acc += fabsf(d0) + fabsf(d1) + fabsf(d2);
                    g.x = (d0 > 0.f) - (d0 < 0.f); g.y = (d1 > 0.f) - (d1 < 0.f); g.z = (d2 > 0.f) - (d2 < 0.f);
                    if (SPEC) {
                        // signs the shade pass speculated (and already back-propagated) vs the antialiased image's
                        const uint32_t now = pack_signs(g.x, g.y, g.z), was = __float_as_uint(c0.w);
                        if (now != was) {
                            // per channel (now - was) in -2..2, stored + 2 in 3 bits
                            corr_pack = 0x200u;
#pragma unroll
                            for (int ch = 0; ch < 3; ch++) {
                                const int dlt = (int)((now >> (2 * ch)) & 3u) - (int)((was >> (2 * ch)) & 3u);
                                corr_pack |= (uint32_t)(dlt + 2) << (3 * ch);
                            }
                        }
                    }
                }
                // mesh_sfs_optim.py:295  mean((pred_mask - valid_mask)^2), as a correction to sum valid_mask^2
#if FMHR_AA_HOIST
                const float vm = h_vm;
#else
                const float vm = __ldg(valid_masks + (size_t)view * hw + rem);
#endif
                const float dm = amask - vm;
                msk_acc += (double)dm * (double)dm - (double)vm * (double)vm;
                g.w = dm;
                if (!SPEC || (nq > 0 && ((recv_s[wib] >> lane) & 1u))) gplane0[pix] = g;
                if (dbg_image) { dbg_image[pix * 3] = a0.x; dbg_image[pix * 3 + 1] = a0.y; dbg_image[pix * 3 + 2] = a0.z; }
                if (dbg_mask) dbg_mask[pix] = amask;
            } else {
                // phase A (mesh_sfs_optim.py:217-230): a0 = antialiased normals, a1 = antialiased albedo
                float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0;
                float3 pred = make_float3(0.f, 0.f, 0.f);
                if (valid) {
                    const float* sh = sh_coeffs + (size_t)__ldg(sh_idx + n) * 9;
                    const float len = sqrtf(a0.x * a0.x + a0.y * a0.y + a0.z * a0.z);
                    const float inv = 1.0f / fmaxf(len, 1e-12f);
                    const float nx = a0.x * inv, ny = a0.y * inv, nz = a0.z * inv;
                    const float r = sh_radiance(sh, nx, ny, nz);
                    pred = make_float3(r * a1.x, r * a1.y, r * a1.z);
                    const float d0 = pred.x - __ldg(img), d1 = pred.y - __ldg(img + 1), d2 = pred.z - __ldg(img + 2);
                    abs_acc += fabsf(d0) + fabsf(d1) + fabsf(d2);
                    const float s0 = (d0 > 0.f) - (d0 < 0.f), s1 = (d1 > 0.f) - (d1 < 0.f), s2 = (d2 > 0.f) - (d2 < 0.f);
                    g1 = make_float4(s0 * r, s1 * r, s2 * r, 0.0f);  // d/d(albedo_aa)
                    const float gr = s0 * a1.x + s1 * a1.y + s2 * a1.z;
                    gc[0] = gr; gc[1] = gr * ny; gc[2] = gr * nz; gc[3] = gr * nx; gc[4] = gr * nx * ny; gc[5] = gr * ny * nz;
                    gc[6] = gr * (2 * nz * nz - nx * nx - ny * ny); gc[7] = gr * nz * nx; gc[8] = gr * (nx * nx - ny * ny);
                    // normals are not trainable in phase A (vertices detached, mesh_sfs_optim.py:191,196): g0 stays 0
                }
                gplane0[pix] = g0;
                gplane1[pix] = g1;
                if (dbg_image) { dbg_image[pix * 3] = pred.x; dbg_image[pix * 3 + 1] = pred.y; dbg_image[pix * 3 + 2] = pred.z; }
            }
        }
        if (SPEC) {
            const unsigned mc = __ballot_sync(0xffffffffu, corr_pack != 0u);
            if (mc) {  // rare (a few 10^3 pixels per iteration): one global atomic per warp batch that has any
                const int leader = __ffs(mc) - 1;
                int qb = 0;
                if (lane == leader) qb = atomicAdd(qcount, __popc(mc));
                qb = __shfl_sync(0xffffffffu, qb, leader);
                if (corr_pack != 0u) {
                    const int slot = qb + __popc(mc & lt);
                    if (slot < qcap) qlist[slot] = make_uint2(pix32, corr_pack);
                    else atomicOr(status, 1);
                }
            }
        }
        if (PHASE == 2) {
            // normal equations of the per-view least-squares SH fit: upper triangle of A^T A (45) + A^T y (9) + count,
            // fp64, warp-reduced when the batch belongs to one view (almost always)
            const int n0 = __shfl_sync(0xffffffffu, n, 0);
            const bool uniform = __all_sync(0xffffffffu, !active || n == n0);
            double* row = init_acc + (size_t)__ldg(view_idx + (uniform ? n0 : n)) * 56;
            int k = 0;
#pragma unroll
            for (int i = 0; i < 9; i++) {
#pragma unroll
                for (int j = i; j <= 9; j++, k++) {  // j == 9: the right-hand side
                    const double v = init_valid ? (double)gc[i] * (double)(j < 9 ? gc[j] : init_y) : 0.0;
                    if (uniform) {
                        const double sv = warp_sum_f64(v);
                        if (lane == 0 && sv != 0.0) atomicAdd(row + k, sv);
                    } else if (v != 0.0) {
                        atomicAdd(row + k, v);
                    }
                }
            }
            const double cv = init_valid ? 1.0 : 0.0;
            if (uniform) {
                const double sc = warp_sum_f64(cv);
                if (lane == 0 && sc != 0.0) atomicAdd(row + 54, sc);
            } else if (init_valid) {
                atomicAdd(row + 54, 1.0);
            }
        }
        if (PHASE == 0) {
            // SH gradient rows: a warp's 32 list entries almost always belong to one view
            const int n0 = __shfl_sync(0xffffffffu, n, 0);
            const bool uniform = __all_sync(0xffffffffu, !active || n == n0);
            if (uniform) {
#pragma unroll
                for (int k = 0; k < 9; k++) {
                    const float sk = warp_sum(gc[k]);
                    if (lane == 0 && sk != 0.0f) atomicAdd(gsh + (size_t)__ldg(sh_idx + n0) * 9 + k, sk);
                }
            } else if (active) {
#pragma unroll
                for (int k = 0; k < 9; k++)
                    if (gc[k] != 0.0f) atomicAdd(gsh + (size_t)__ldg(sh_idx + n) * 9 + k, gc[k]);
            }
        }
        __syncwarp();  // blend[] / q_items / q_n / spix of this warp are rewritten by its next batch
    }
    {
        __shared__ int s_cnt[10];
        const int base = block_reserve(pcount, npb, s_cnt);
        for (int i = lane; i < npb; i += 32) {
            if (base + i < pcap) { plist_a[base + i] = pbuf_a[wib][i]; plist_b[base + i] = pbuf_b[wib][i]; }
            else atomicOr(status, 1);
        }
    }
    warp_acc_add(acc, 1, abs_acc);
    if (PHASE == 1) {
        msk_acc = warp_sum_f64(msk_acc);
        if (lane == 0 && msk_acc != 0.0) atomicAdd(acc + 2 * 32 + ((blockIdx.x * 8 + wib) & 31), msk_acc);
    }
}

// ------------------------------------------------------------------------------------------------
// pixel backward: antialias bwd (gather form for colours, owner-scatter for positions), shading bwd,
// interpolate bwd, rasterize bwd; everything lands in the per-vertex world-space accumulators.
// ------------------------------------------------------------------------------------------------
// Shading backward of ONE covered pixel for an incoming gradient g w.r.t. its pre-antialias value (phase B: shaded
// colour b,g,r; phase A: albedo): SH / normalise / interpolate / rasterize backward from the triangle record, two
// red.global.add.v4.f32 per corner into the world-space vertex accumulators.  LINEAR in g, so the pass-through gradient
// (pixel kernel) and the antialias pair terms (pair kernel) are back-propagated independently.
// Core of the shading backward for a pixel whose triangle record `q` (clip positions, barycentrics, attributes) is
// already in registers: the shade pass calls it right after the forward (phase B), pixel_backward() loads the record first.
template <int PHASE, bool DEFER>
__device__ __forceinline__ void pixel_backward_q(const PixTri& q, int px, int py, float4 gin, const float* __restrict__ M,
                                                 const float* __restrict__ c, int V, int H, int W,
                                                 float4* __restrict__ G, DeferredScatter* def) {
    const float4 g0 = gin, g1 = gin;
    const float w = 1.0f - q.u - q.v;
    if (PHASE == 0) {
        // only the albedo attribute is trainable: interpolate bwd
        if (g1.x == 0.0f && g1.y == 0.0f && g1.z == 0.0f) return;
        atomicAdd(G + 2 * (size_t)q.i0 + 1, make_float4(0.f, q.u * g1.x, q.u * g1.y, q.u * g1.z));
        atomicAdd(G + 2 * (size_t)q.i1 + 1, make_float4(0.f, q.v * g1.x, q.v * g1.y, q.v * g1.z));
        atomicAdd(G + 2 * (size_t)q.i2 + 1, make_float4(0.f, w * g1.x, w * g1.y, w * g1.z));
        return;
    }
    if (g0.x == 0.0f && g0.y == 0.0f && g0.z == 0.0f) return;
    const float3 m = interp3(q.n0, q.n1, q.n2, q);
    const float3 a = interp3(q.b0, q.b1, q.b2, q);
    const float len = sqrtf(m.x * m.x + m.y * m.y + m.z * m.z);
    const float inv = 1.0f / fmaxf(len, 1e-12f);
    const float nx = m.x * inv, ny = m.y * inv, nz = m.z * inv;
    const float r = sh_radiance(c, nx, ny, nz);
    const float3 ga = make_float3(g0.x * r, g0.y * r, g0.z * r);           // d/d(interpolated albedo)
    const float gr = g0.x * a.x + g0.y * a.y + g0.z * a.z;                 // d/d(radiance)
    float3 gn = make_float3(gr * (c[3] + c[4] * ny - 2 * c[6] * nx + c[7] * nz + 2 * c[8] * nx),
                            gr * (c[1] + c[4] * nx + c[5] * nz - 2 * c[6] * ny - 2 * c[8] * ny),
                            gr * (c[2] + c[5] * ny + 4 * c[6] * nz + c[7] * nx));
    float3 gm;  // through F.normalize(eps=1e-12), mesh_sfs_optim.py:273
    if (len > 1e-12f) {
        const float dt = nx * gn.x + ny * gn.y + nz * gn.z;
        gm = make_float3((gn.x - nx * dt) * inv, (gn.y - ny * dt) * inv, (gn.z - nz * dt) * inv);
    } else {
        gm = make_float3(gn.x * inv, gn.y * inv, gn.z * inv);
    }
    // interpolate bwd: d/du, d/dv over the six differentiable attributes
    const float du = gm.x * (q.n0.x - q.n2.x) + gm.y * (q.n0.y - q.n2.y) + gm.z * (q.n0.z - q.n2.z) +
                     ga.x * (q.b0.x - q.b2.x) + ga.y * (q.b0.y - q.b2.y) + ga.z * (q.b0.z - q.b2.z);
    const float dv = gm.x * (q.n1.x - q.n2.x) + gm.y * (q.n1.y - q.n2.y) + gm.z * (q.n1.z - q.n2.z) +
                     ga.x * (q.b1.x - q.b2.x) + ga.y * (q.b1.y - q.b2.y) + ga.z * (q.b1.z - q.b2.z);
    // rasterize bwd
    const float fx = (float)(2 * px + 1) / (float)W - 1.0f;
    const float fy = (float)(2 * py + 1) / (float)H - 1.0f;
    const float q0x = q.p0.x - fx * q.p0.w, q0y = q.p0.y - fy * q.p0.w;
    const float q1x = q.p1.x - fx * q.p1.w, q1y = q.p1.y - fy * q.p1.w;
    const float q2x = q.p2.x - fx * q.p2.w, q2y = q.p2.y - fy * q.p2.w;
    const float e0 = q1x * q2y - q1y * q2x, e1 = q2x * q0y - q2y * q0x, e2 = q0x * q1y - q0y * q1x;
    const float at = e0 + e1 + e2;
    const float iw = 1.0f / (at + copysignf(1e-6f, at));
    const float bb0 = e0 * iw, bb1 = e1 * iw;
    const float gb0 = du * iw, gb1 = dv * iw, gbb = gb0 * bb0 + gb1 * bb1;
    const float g0x = gbb * (q2y - q1y) - gb1 * q2y;
    const float g1x = gbb * (q0y - q2y) + gb0 * q2y;
    const float g2x = gbb * (q1y - q0y) - gb0 * q1y + gb1 * q0y;
    const float g0y = gbb * (q1x - q2x) + gb1 * q2x;
    const float g1y = gbb * (q2x - q0x) - gb0 * q2x;
    const float g2y = gbb * (q0x - q1x) + gb0 * q1x - gb1 * q0x;
    const float3 w0 = clip_to_world(M, g0x, g0y, -fx * g0x - fy * g0y);
    const float3 w1 = clip_to_world(M, g1x, g1y, -fx * g1x - fy * g1y);
    const float3 w2 = clip_to_world(M, g2x, g2y, -fx * g2x - fy * g2y);
    const int vi[3] = {q.i0, q.i1, q.i2};
    const float wt[3] = {q.u, q.v, w};
    const float3 wp[3] = {w0, w1, w2};
    const float3 nn[3] = {q.n0, q.n1, q.n2};
    const bool dg[3] = {(q.flags & 1) != 0, (q.flags & 2) != 0, (q.flags & 4) != 0};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float3 gk = make_float3(wt[k] * gm.x, wt[k] * gm.y, wt[k] * gm.z);  // d/d(vertex normal attribute)
        float a1, a2;
        if (!dg[k]) {
            float3 t1, t2;
            tangent_frame(nn[k], t1, t2);
            a1 = t1.x * gk.x + t1.y * gk.y + t1.z * gk.z;
            a2 = t2.x * gk.x + t2.y * gk.y + t2.z * gk.z;
        } else {  // |N| <= 1e-6: the normalisation does not project, keep all three components
            a1 = gk.x; a2 = gk.y;
            atomicAdd(reinterpret_cast<float*>(G + 2 * (size_t)V + vi[k]) + 3, gk.z);
        }
        if (DEFER) {
            def->vi[k] = vi[k];
            def->lo[k] = make_float4(wp[k].x, wp[k].y, wp[k].z, a1);
            def->hi[k] = make_float4(a2, wt[k] * ga.x, wt[k] * ga.y, wt[k] * ga.z);
        } else {
            float4* Gk = G + 2 * (size_t)vi[k];
            atomicAdd(Gk, make_float4(wp[k].x, wp[k].y, wp[k].z, a1));
            atomicAdd(Gk + 1, make_float4(a2, wt[k] * ga.x, wt[k] * ga.y, wt[k] * ga.z));
        }
    }
}

// Scatter of the deferred corner records with every lane of the warp present (lanes without a record carry vi = -1):
// lane pairs (2j, 2j+1) write BOTH 16-byte halves of one vertex record in the same instruction - first the even lane's
// record, then the odd lane's - so every red.v4 instruction completes whole 32-byte sectors instead of touching 32 half
// sectors (tools/ubench/red_bench.cu form C against form B: 20.5 vs 24.6 us for the iteration's references; in the shade
// pass: 5,223 -> 5,292 iters/s, profiles/r2/README.md).
__device__ __forceinline__ float4 shfl_xor4(float4 v, int m) {
    return make_float4(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m),
                       __shfl_xor_sync(0xffffffffu, v.z, m), __shfl_xor_sync(0xffffffffu, v.w, m));
}
__device__ __forceinline__ void scatter_sector_pairs(const DeferredScatter& d, float4* __restrict__ G) {
    const bool odd = (threadIdx.x & 1) != 0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        // the half the partner will write for this lane: an even lane writes lower halves, an odd lane upper halves
        const float4 got = shfl_xor4(odd ? d.lo[k] : d.hi[k], 1);
        const int vo = __shfl_xor_sync(0xffffffffu, d.vi[k], 1);
        const int v_even = odd ? vo : d.vi[k], v_odd = odd ? d.vi[k] : vo;
        if (v_even >= 0) atomicAdd(G + 2 * (size_t)v_even + (odd ? 1 : 0), odd ? got : d.lo[k]);
        if (v_odd >= 0) atomicAdd(G + 2 * (size_t)v_odd + (odd ? 1 : 0), odd ? d.hi[k] : got);
    }
}

template <int PHASE>
__device__ __forceinline__ void pixel_backward(uint32_t pix32, int tself, float4 gin, const float4* __restrict__ trirec,
                                               float invW, float invH, const float* __restrict__ viewM,
                                               const float* __restrict__ sh_coeffs, const int32_t* __restrict__ sh_idx,
                                               int V, int H, int W, float4* __restrict__ G) {
    if (gin.x == 0.0f && gin.y == 0.0f && gin.z == 0.0f) return;
    const PixAddr pa = pix_decode(pix32, H, W);
    const int n = pa.n, px = pa.px, py = pa.py;
    const float* M = viewM + (size_t)n * kViewM;
    const float* c = sh_coeffs + (size_t)__ldg(sh_idx + n) * 9;
    PixTri q;
    load_pixtri(tself, px, py, trirec, M, invW, invH, q);
    pixel_backward_q<PHASE, false>(q, px, py, gin, M, c, V, H, W, G, nullptr);
}

// ------------------------------------------------------------------------------------------------
// backward, part 1: the blending pairs recorded by the antialias pass (a few 10^4 per iteration, one thread each).
//   * colour / albedo gradient of the two pixels of the pair (-alpha g_recv, +alpha g_recv), back-propagated through each
//     pixel's shading chain right here (pixel_backward is linear in its input, so no per-pixel delta plane is needed
//     and this kernel is independent of the pixel kernel);
//   * phase B: silhouette position gradient -> world-space accumulators.
// ------------------------------------------------------------------------------------------------
template <int PHASE>
__global__ void __launch_bounds__(128) ham_pair_bwd_kernel(
    const uint4* __restrict__ plist_a, const uint32_t* __restrict__ plist_b, const int* __restrict__ pcount, int pcap,
    const unsigned long long* __restrict__ zbuf, const float4* __restrict__ vg, const float* __restrict__ viewM, int V,
    int H, int W, const float4* __restrict__ plane0, const float4* __restrict__ gplane0,
    const float4* __restrict__ gplane1, const float4* __restrict__ trirec, float invW, float invH,
    const float* __restrict__ sh_coeffs, const int32_t* __restrict__ sh_idx, float4* __restrict__ G,
    const uint2* __restrict__ qlist, const int* __restrict__ qcount, int qcap) {
    FMHR_TRACE_SCOPE(8);
    const int np = min(*pcount, pcap);
    const int hw = H * W;
    if (PHASE == 1 && qlist) {
        // corrections of the shade pass' speculative backward: pixels whose antialiased L1 sign differs from sign(pred -
        // img); the shading backward is linear in its input, so the difference of the two signs is propagated here
        // (taken from the FAR end of the thread range: the pair items below start at thread 0, and a thread that had one
        // of each ran the two gather chains one after the other - the kernel's tail)
        const int nq = min(*qcount, qcap);
        const int n_threads = gridDim.x * blockDim.x;
        for (int e = n_threads - 1 - (int)(blockIdx.x * blockDim.x + threadIdx.x); e < nq; e += n_threads) {
            const uint2 it = qlist[e];
            const float4 dg = make_float4((float)((int)(it.y & 7u) - 2), (float)((int)((it.y >> 3) & 7u) - 2),
                                          (float)((int)((it.y >> 6) & 7u) - 2), 0.f);
            const NbrKeys k = decode_key(zbuf[it.x]);
            pixel_backward<PHASE>(it.x, k.tri, dg, trirec, invW, invH, viewM, sh_coeffs, sh_idx, V, H, W, G);
        }
    }
    // three work items per pair - the shading chains of its two pixels and the silhouette position gradient are independent
    // gather chains, so they go to three threads instead of one after the other in one
    for (int wi = blockIdx.x * blockDim.x + threadIdx.x; wi < 3 * np; wi += gridDim.x * blockDim.x) {
        const int e = wi / 3, role = wi - 3 * e;
        const uint4 a = plist_a[e];
        AAPair pr;
        const size_t pix0 = a.x;
        const int d = (int)(a.y & 1u);
        pr.from1 = (int)((a.y >> 1) & 1u); pr.clamped = (int)((a.y >> 2) & 1u); pr.di = (int)((a.y >> 3) & 3u);
        pr.alpha = __uint_as_float(a.z); pr.i1 = (int)a.w; pr.tri = 0;
        const size_t pix1 = pix0 + (d ? W : 1);
        const int n = (int)(pix0 / hw);
        const int rem0 = (int)(pix0 - (size_t)n * hw), qy = rem0 / W, qx = rem0 - qy * W;
        const NbrKeys k0 = decode_key(zbuf[pix0]), k1 = decode_key(zbuf[pix1]);
        const size_t recv = pr.alpha > 0.0f ? pix0 : pix1;
        // phase B blends the shaded colour (gplane0.xyz) and coverage (gplane0.w); phase A the albedo (gplane1.xyz)
        const float4 gr = (PHASE == 1) ? gplane0[recv] : gplane1[recv];
        // out[recv] += alpha*(c_second - c_first): d/dc_first = -alpha*g, d/dc_second = +alpha*g; only pixels on the
        // main pass' list consume it (phase B: valid, phase A: covered)
        const bool t0 = PHASE == 1 ? k0.valid : k0.tri >= 0, t1 = PHASE == 1 ? k1.valid : k1.tri >= 0;
        if (role == 0) {
            if (t0)
                pixel_backward<PHASE>((uint32_t)pix0, k0.tri, make_float4(-pr.alpha * gr.x, -pr.alpha * gr.y, -pr.alpha * gr.z, 0.f),
                                      trirec, invW, invH, viewM, sh_coeffs, sh_idx, V, H, W, G);
            continue;
        }
        if (role == 1) {
            if (t1)
                pixel_backward<PHASE>((uint32_t)pix1, k1.tri, make_float4(pr.alpha * gr.x, pr.alpha * gr.y, pr.alpha * gr.z, 0.f),
                                      trirec, invW, invH, viewM, sh_coeffs, sh_idx, V, H, W, G);
            continue;
        }
        pr.i2 = (int)plist_b[e];
        if (PHASE == 1 && !pr.clamped) {
            float4 f0 = make_float4(0.f, 0.f, 0.f, 0.f), s0 = f0;
            if (k0.tri >= 0) f0 = plane0[pix0];
            if (k1.tri >= 0) s0 = plane0[pix1];
            const float dd_img = gr.x * (s0.x - f0.x) + gr.y * (s0.y - f0.y) + gr.z * (s0.z - f0.z);
            const float dd_msk = gr.w * ((k1.tri >= 0 ? 1.0f : 0.0f) - (k0.tri >= 0 ? 1.0f : 0.0f));
            if (dd_img != 0.0f || dd_msk != 0.0f) {
                const float* M = viewM + (size_t)n * kViewM;
                float4 e1, e2;
                aa_pos_grad(pr, qx, qy, d, clip_from_world(M, __ldg(vg + 2 * (size_t)pr.i1)),
                            clip_from_world(M, __ldg(vg + 2 * (size_t)pr.i2)), H, W, 1.0f, e1, e2);
                const float3 w1 = clip_to_world(M, e1.x, e1.y, e1.w);
                const float3 w2 = clip_to_world(M, e2.x, e2.y, e2.w);
                float4* Gm = G + 2 * (size_t)V;
                if (dd_img != 0.0f) {
                    atomicAdd(G + 2 * (size_t)pr.i1, make_float4(dd_img * w1.x, dd_img * w1.y, dd_img * w1.z, 0.f));
                    atomicAdd(G + 2 * (size_t)pr.i2, make_float4(dd_img * w2.x, dd_img * w2.y, dd_img * w2.z, 0.f));
                }
                if (dd_msk != 0.0f) {
                    atomicAdd(Gm + pr.i1, make_float4(dd_msk * w1.x, dd_msk * w1.y, dd_msk * w1.z, 0.f));
                    atomicAdd(Gm + pr.i2, make_float4(dd_msk * w2.x, dd_msk * w2.y, dd_msk * w2.z, 0.f));
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward, part 2: one thread per covered pixel of the compact list (every lane busy), pass-through gradient of the
// pixel's own loss term.  Runs concurrently with the pair kernel (both only add into the vertex accumulators).
// ------------------------------------------------------------------------------------------------
template <int PHASE>
__global__ void __launch_bounds__(256, FMHR_LB_BWD) ham_pixel_bwd_kernel(
    const uint2* __restrict__ clist, const int* __restrict__ ccount,
    const float4* __restrict__ trirec, float invW, float invH, const float* __restrict__ viewM,
    const float* __restrict__ sh_coeffs, const int32_t* __restrict__ sh_idx, int V, int H, int W,
    const float4* __restrict__ gplane0, const float4* __restrict__ gplane1, float4* __restrict__ G) {
    FMHR_TRACE_SCOPE(9);
    const int nv = *ccount;
    const int e_first = blockIdx.x * blockDim.x + threadIdx.x, e_stride = gridDim.x * blockDim.x;
    uint2 ent_next = e_first < nv ? clist[e_first] : make_uint2(0u, 0u);
    for (int e = e_first; e < nv; e += e_stride) {
        const uint2 ent = ent_next;  // (pixel, triangle | valid << 31) from the scan / shade passes; next one prefetched
        if (e + e_stride < nv) ent_next = clist[e + e_stride];
        // phase B: tmp_img[valid_idx] = pred_img -> only valid pixels feed the shader (mesh_sfs_optim.py:285-286);
        // phase A back-propagates through every covered pixel's albedo
        if (PHASE == 1 && !(ent.y >> 31)) continue;
        const float4 g = PHASE == 1 ? gplane0[ent.x] : gplane1[ent.x];
        pixel_backward<PHASE>(ent.x, (int)(ent.y & kTriMask), g, trirec, invW, invH, viewM, sh_coeffs, sh_idx, V, H, W, G);
    }
}

// HAM initialisation, part 2: per-view and global least-squares SH lighting from the accumulated normal equations
// (np.linalg.lstsq in the reference, mesh_sfs_optim.py:153,166).  Block v < num solves view v, block num the sum over
// all views.  9x9 Cholesky in fp64 by one thread; a view without enough valid normals (non-positive pivot) gets zeros
// in the dependent components (lstsq returns the minimum-norm solution there - outside the reference's use).
// init_acc rows: [45 upper-triangle (i <= j, with the rhs as column 9 interleaved per row) ...] laid out per row i as
// (A_ii .. A_i8, b_i), then [54] = number of valid pixels.
__global__ void ham_init_solve_kernel(const double* __restrict__ init_acc, int num, float* __restrict__ sh_coeffs,
                                      float* __restrict__ sh_global, double* __restrict__ global_row) {
    if (threadIdx.x != 0) return;
    const int v = blockIdx.x;
    double A[9][9], b[9];
    for (int i = 0; i < 9; i++) { b[i] = 0.0; for (int j = 0; j < 9; j++) A[i][j] = 0.0; }
    const int r0 = v < num ? v : 0, r1 = v < num ? v + 1 : num;
    double count = 0.0;
    for (int r = r0; r < r1; r++) {
        const double* row = init_acc + (size_t)r * 56;
        int k = 0;
        for (int i = 0; i < 9; i++)
            for (int j = i; j <= 9; j++, k++) {
                if (j < 9) A[i][j] += row[k]; else b[i] += row[k];
            }
        count += row[54];
    }
    // Cholesky A = L L^T (lower triangle stored in A[j][i], i <= j)
    double L[9][9];
    bool ok[9];
    for (int i = 0; i < 9; i++) {
        for (int j = 0; j <= i; j++) {
            double sum = A[j][i];
            for (int k = 0; k < j; k++) sum -= L[i][k] * L[j][k];
            if (i == j) {
                ok[i] = sum > 1e-12 * (A[i][i] > 0.0 ? A[i][i] : 1.0);
                L[i][i] = ok[i] ? sqrt(sum) : 1.0;
            } else {
                L[i][j] = ok[j] ? sum / L[j][j] : 0.0;
            }
        }
    }
    double y[9], x[9];
    for (int i = 0; i < 9; i++) {
        double sum = b[i];
        for (int k = 0; k < i; k++) sum -= L[i][k] * y[k];
        y[i] = ok[i] ? sum / L[i][i] : 0.0;
    }
    for (int i = 8; i >= 0; i--) {
        double sum = y[i];
        for (int k = i + 1; k < 9; k++) sum -= L[k][i] * x[k];
        x[i] = ok[i] ? sum / L[i][i] : 0.0;
    }
    float* out = v < num ? sh_coeffs + (size_t)v * 9 : sh_global;
    for (int i = 0; i < 9; i++) out[i] = (float)x[i];
    if (v == num) { global_row[54] = count; for (int i = 0; i < 4; i++) global_row[i] = 0.0; }
}

// HAM initialisation, part 3: albedo_mean = mean over the valid pixels of img / radiance(global SH, normal)
// (mesh_sfs_optim.py:173-174); the unit antialiased normals were left in `nplane` by the antialias pass.
__global__ void __launch_bounds__(256) ham_init_albedo_kernel(const uint2* __restrict__ clist, const int* __restrict__ ccount,
                                                              const float4* __restrict__ nplane,
                                                              const float* __restrict__ imgs,
                                                              const int32_t* __restrict__ view_idx, int H, int W,
                                                              const float* __restrict__ sh_global,
                                                              double* __restrict__ global_row) {
    const int nv = *ccount;
    const int hw = H * W;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nv; e += gridDim.x * blockDim.x) {
        const uint2 ent = clist[e];
        if (!(ent.y >> 31)) continue;
        const float4 nr = nplane[ent.x];
        if (nr.w == 0.0f) continue;
        const PixAddr pa = pix_decode(ent.x, H, W);
        const float r = sh_radiance(sh_global, nr.x, nr.y, nr.z);
        const float* img = imgs + ((size_t)__ldg(view_idx + pa.n) * hw + pa.rem) * 3;
        s0 += (double)(__ldg(img) / r); s1 += (double)(__ldg(img + 1) / r); s2 += (double)(__ldg(img + 2) / r);
    }
    s0 = warp_sum_f64(s0); s1 = warp_sum_f64(s1); s2 = warp_sum_f64(s2);
    if ((threadIdx.x & 31) == 0) {
        if (s0 != 0.0) atomicAdd(global_row + 0, s0);
        if (s1 != 0.0) atomicAdd(global_row + 1, s1);
        if (s2 != 0.0) atomicAdd(global_row + 2, s2);
    }
}
__global__ void ham_init_albedo_mean_kernel(const double* __restrict__ global_row, float* __restrict__ albedo_mean) {
    if (threadIdx.x < 3) albedo_mean[threadIdx.x] = (float)(global_row[threadIdx.x] / global_row[54]);
}

// One warp: lane j owns spread slot j of every accumulator (fp64 shuffles), view totals are strided over the lanes.
__global__ void __launch_bounds__(32) ham_finalize_scalars_kernel(const double* __restrict__ acc,
                                                                  const double* __restrict__ view_vm2,
                                                                  const int32_t* __restrict__ view_idx, int n_views,
                                                                  int tiles, int phase, float* __restrict__ scal,
                                                                  double* __restrict__ reg_totals) {
    FMHR_TRACE_SCOPE(10);
    const int lane = threadIdx.x;
    double vm2 = 0.0;  // sum of valid_mask^2 over the batch's views; acc[2] holds the listed pixels' corrections
    if (phase == 1)
        for (int n = lane; n < n_views; n += 32) vm2 += view_vm2[(size_t)view_idx[n] * (tiles + 1) + tiles];
    const double a0 = warp_sum_f64(acc[0 * 32 + lane]), a1 = warp_sum_f64(acc[1 * 32 + lane]);
    const double a2 = warp_sum_f64(acc[2 * 32 + lane] + vm2);
    // regulariser totals (the regulariser kernel finished long ago: the caller orders this kernel behind it): summed here,
    // beside the pair kernel, instead of by one thread of the update kernel (128 dependent fp64 adds = that kernel's tail)
    const double r3 = warp_sum_f64(acc[3 * 32 + lane]), r4 = warp_sum_f64(acc[4 * 32 + lane]);
    const double r5 = warp_sum_f64(acc[5 * 32 + lane]), r6 = warp_sum_f64(acc[6 * 32 + lane]);
    if (lane == 0) {
        scal[0] = (float)a0;
        scal[1] = (float)a1;
        scal[2] = (float)a2;
        scal[3] = 0.0f;
        reg_totals[0] = r3; reg_totals[1] = r4; reg_totals[2] = r5; reg_totals[3] = r6;
    }
}

// view_vm2[view][tile] = sum of valid_mask^2 over the 16x16 tile, view_vm2[view][tiles] = sum over the view (constants
// of the optimisation: valid_masks never change, mesh_sfs_optim.py:163); the iteration only uses the view totals.
__global__ void __launch_bounds__(256) ham_view_vm2_kernel(const float* __restrict__ valid_masks, int H, int W,
                                                           double* __restrict__ out) {
    __shared__ double red[8];
    const int view = blockIdx.z, tiles = gridDim.x * gridDim.y;
    const int px = blockIdx.x * kTile + threadIdx.x, py = blockIdx.y * kTile + threadIdx.y;
    double s = 0.0;
    if (px < W && py < H) {
        const float x = valid_masks[((size_t)view * H + py) * W + px];
        s = (double)x * (double)x;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const int tid = threadIdx.y * kTile + threadIdx.x;
    if ((tid & 31) == 0) red[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; i++) t += red[i];
        out[(size_t)view * (tiles + 1) + blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(32) ham_view_vm2_total_kernel(int tiles, double* __restrict__ out) {
    double* row = out + (size_t)blockIdx.x * (tiles + 1);
    double t = 0.0;
    for (int i = threadIdx.x; i < tiles; i += 32) t += row[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) row[tiles] = t;
}

// zbuf -> rast_out for fmhr_ham_debug_export
__global__ void __launch_bounds__(256) ham_export_rast_kernel(const unsigned long long* __restrict__ zbuf,
                                                              const float4* __restrict__ vg,
                                                              const float* __restrict__ viewM,
                                                              const int32_t* __restrict__ tri, int V, int H, int W,
                                                              float4* __restrict__ rast) {
    const int n = blockIdx.y, hw = H * W;
    const int rem = blockIdx.x * blockDim.x + threadIdx.x;
    if (rem >= hw) return;
    const size_t pix = (size_t)n * hw + rem;
    const unsigned long long key = zbuf[pix];
    if (key == ZB_EMPTY) { rast[pix] = make_float4(0.f, 0.f, 0.f, 0.f); return; }
    const int t = (int)((uint32_t)key & kTriMask), py = rem / W, px = rem - py * W;
    const float* M = viewM + (size_t)n * kViewM;
    const Bary b = bary_at(clip_from_world(M, __ldg(vg + 2 * (size_t)__ldg(tri + 3 * t))),
                           clip_from_world(M, __ldg(vg + 2 * (size_t)__ldg(tri + 3 * t + 1))),
                           clip_from_world(M, __ldg(vg + 2 * (size_t)__ldg(tri + 3 * t + 2))), px, py,
                           xd(1.0f, (float)W), xd(1.0f, (float)H));
    rast[pix] = make_float4(b.u, b.v, b.zw, (float)(t + 1));
}
// clip positions [n,V,4] for fmhr_ham_debug_export (the iteration itself never materialises them)
__global__ void __launch_bounds__(256) ham_export_pos_kernel(const float4* __restrict__ vg, const float* __restrict__ viewM,
                                                             int V, float4* __restrict__ pos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, n = blockIdx.y;
    if (i < V) pos[(size_t)n * V + i] = clip_from_world(viewM + (size_t)n * kViewM, __ldg(vg + 2 * (size_t)i));
}

// ------------------------------------------------------------------------------------------------
// update: regularisers, normal backward, normalisation, Adam
// ------------------------------------------------------------------------------------------------
// The vertex-domain kernels below use 4 lanes per vertex (kLPV): V is only ~50k, so one thread per vertex leaves the
// GPU at <0.3 waves of latency-bound gather loops; splitting each vertex's ~6 neighbours / incident faces over 4 lanes
// (shuffle-reduced) keeps the whole mesh in ONE wave of resident warps (8 lanes needed 1.3 waves: the tail wave
// doubled the kernel's latency-bound run time).  Every neighbour gather is one 128/256-bit record load.
constexpr int kLPV = 4;
__device__ __forceinline__ float sub_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// Regulariser forward (Laplacian for vertices and albedo, edge / delta losses) and the step's Adam scalars.  Depends only
// on the vertices and albedo, not on the rendering: fmhr_ham_step_render runs it on a side stream, concurrently with the
// coverage kernel, so it is off the iteration's critical path.
__global__ void __launch_bounds__(256, 6) ham_regulariser_kernel(
    fmhr_ham_config cfg, const float4* __restrict__ vg, const float* __restrict__ delta, const float4* __restrict__ vattr,
    const int32_t* __restrict__ v2f_ptr, const int2* __restrict__ v2f_nbr, const int32_t* __restrict__ v2v_ptr,
    const int32_t* __restrict__ v2v_idx, float4* __restrict__ ys, double* __restrict__ acc,
    int32_t* __restrict__ adam_step, float* __restrict__ adam_sc) {
    FMHR_TRACE_SCOPE(3);
    __shared__ float red[4][8];
    const int V = cfg.V;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) / kLPV, sub = threadIdx.x & (kLPV - 1);
    float lv = 0.f, la = 0.f, le = 0.f, ld = 0.f;
    float3 sv = make_float3(0.f, 0.f, 0.f), sa = sv, vi = sv;
    int deg = 0;
    if (i < V) {
        const int b = __ldg(v2v_ptr + i), e = __ldg(v2v_ptr + i + 1);
        deg = e - b;
        for (int j = b + sub; j < e; j += kLPV) {
            const size_t nb = 2 * (size_t)__ldg(v2v_idx + j);
            const float4 xv = vg[nb], xa_ = __ldg(vattr + nb + 1);
            sv.x += xv.x; sv.y += xv.y; sv.z += xv.z;
            sa.x += xa_.x; sa.y += xa_.y; sa.z += xa_.z;
        }
        const float4 v4 = vg[2 * (size_t)i];
        vi = make_float3(v4.x, v4.y, v4.z);
        // edge hinge (mesh_sfs_optim.py:296-302): every half-edge is seen from both of its endpoints -> weight 1/2
        const int fb = __ldg(v2f_ptr + i), fe = __ldg(v2f_ptr + i + 1);
        for (int j = fb + sub; j < fe; j += kLPV) {
            const int2 nb = __ldg(v2f_nbr + j);
#pragma unroll
            for (int s = 1; s <= 2; s++) {
                const float4 o = vg[2 * (size_t)(s == 1 ? nb.x : nb.y)];
                const float dx = vi.x - o.x, dy = vi.y - o.y, dz = vi.z - o.z;
                const float x = dx * dx + dy * dy + dz * dz - cfg.edge_length_mean;
                le += 0.5f * fminf(fmaxf(x, 0.0f), 1.0f);
            }
        }
    }
    sv.x = sub_sum(sv.x); sv.y = sub_sum(sv.y); sv.z = sub_sum(sv.z);
    sa.x = sub_sum(sa.x); sa.y = sub_sum(sa.y); sa.z = sub_sum(sa.z);
    if (i < V && sub == 0) {
        const float invd = (deg > 0) ? 1.0f / (float)deg : 0.0f;
        const float4 at = __ldg(vattr + 2 * (size_t)i + 1);
        sv = make_float3(sv.x * invd - vi.x, sv.y * invd - vi.y, sv.z * invd - vi.z);
        sa = make_float3(sa.x * invd - at.x, sa.y * invd - at.y, sa.z * invd - at.z);
        lv = sqrtf(sv.x * sv.x + sv.y * sv.y + sv.z * sv.z);
        la = sqrtf(sa.x * sa.x + sa.y * sa.y + sa.z * sa.z);
        // rows of the Laplacian backward: L^T yhat = sum_j yhat_j / deg_j - yhat_i
        const float iv = lv > 0.f ? invd / lv : 0.f, ia = la > 0.f ? invd / la : 0.f;
        ys[2 * (size_t)i] = make_float4(sv.x * iv, sv.y * iv, sv.z * iv, (float)deg);
        ys[2 * (size_t)i + 1] = make_float4(sa.x * ia, sa.y * ia, sa.z * ia, 0.f);
        const float* dp = delta + 3 * (size_t)i;
        ld = dp[0] * dp[0] + dp[1] * dp[1] + dp[2] * dp[2];
    }
    lv = warp_sum(lv); la = warp_sum(la); le = warp_sum(le); ld = warp_sum(ld);
    if ((threadIdx.x & 31) == 0) {
        const int w = threadIdx.x >> 5;
        red[0][w] = lv; red[1][w] = la; red[2][w] = le; red[3][w] = ld;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        float s = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; w++) s += red[threadIdx.x][w];
        if (s != 0.0f) atomicAdd(acc + (3 + threadIdx.x) * 32 + (blockIdx.x & 31), (double)s);
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x >= 64 && threadIdx.x < 70) {
        // Adam bias corrections (torch.optim.Adam defaults); which parameters step depends on the phase:
        //   phase A: albedo, sh   (mesh_sfs_optim.py:193)     phase B: delta, albedo   (:242-244, sh has no grad)
        // six fp64 pow() calls, one per lane (a single thread doing all six was a ~6 us serial tail)
        const int q = threadIdx.x - 64, k = q >> 1;
        const int step = k == 0 ? (cfg.phase == 1) : (k == 1 ? 1 : (cfg.phase == 0));
        const float lr = k == 0 ? cfg.lr : (k == 1 ? cfg.albedo_lr : cfg.sh_lr);
        const int t = adam_step[k] + step;
        const double bc = 1.0 - pow((double)((q & 1) ? cfg.beta2 : cfg.beta1), (double)max(t, 1));
        adam_sc[q] = (q & 1) ? (float)sqrt(bc) : (float)((double)lr / bc);
        __syncwarp(0x3fu);  // both lanes of a parameter have read the counter before it advances
        if ((q & 1) == 0) adam_step[k] = t;
    }
}

// Regulariser backward (phase B): gradient of  lap_weight * Laplacian(vertices) + edge hinge + delta loss  w.r.t. delta and
// of  albedo_weight * Laplacian(albedo)  w.r.t. albedo, fully weighted.  View-independent like the regulariser forward
// whose rows `ys` it gathers, so it runs behind it on the side stream, under the coverage kernel, and the update kernel on
// the critical path is left with the normal backward and Adam.
__global__ void __launch_bounds__(256, 6) ham_reg_grad_kernel(
    fmhr_ham_config cfg, const float4* __restrict__ vg, const float* __restrict__ delta,
    const int32_t* __restrict__ v2f_ptr, const int2* __restrict__ v2f_nbr, const int32_t* __restrict__ v2v_ptr,
    const int32_t* __restrict__ v2v_idx, const float4* __restrict__ ys, float4* __restrict__ greg) {
    FMHR_TRACE_SCOPE(18);
    const int V = cfg.V;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) / kLPV, sub = threadIdx.x & (kLPV - 1);
    float3 lv = make_float3(0.f, 0.f, 0.f), la = lv, ge = lv;
    if (i < V) {
        const int b = __ldg(v2v_ptr + i), e = __ldg(v2v_ptr + i + 1);
        for (int q = b + sub; q < e; q += kLPV) {
            const F8 y = ld256(ys + 2 * (size_t)__ldg(v2v_idx + q));
            la.x += y.b.x; la.y += y.b.y; la.z += y.b.z;
            lv.x += y.a.x; lv.y += y.a.y; lv.z += y.a.z;
        }
        const float4 v4 = vg[2 * (size_t)i];
        const int fb = __ldg(v2f_ptr + i), fe = __ldg(v2f_ptr + i + 1);
        for (int j = fb + sub; j < fe; j += kLPV) {
            const int2 nb = __ldg(v2f_nbr + j);
#pragma unroll
            for (int s2 = 0; s2 < 2; s2++) {
                const float4 o = vg[2 * (size_t)(s2 == 0 ? nb.x : nb.y)];
                const float dx = v4.x - o.x, dy = v4.y - o.y, dz = v4.z - o.z;
                const float x = dx * dx + dy * dy + dz * dz - cfg.edge_length_mean;
                if (x >= 0.0f && x <= 1.0f) { ge.x += 2.0f * dx; ge.y += 2.0f * dy; ge.z += 2.0f * dz; }
            }
        }
    }
    la.x = sub_sum(la.x); la.y = sub_sum(la.y); la.z = sub_sum(la.z);
    lv.x = sub_sum(lv.x); lv.y = sub_sum(lv.y); lv.z = sub_sum(lv.z);
    ge.x = sub_sum(ge.x); ge.y = sub_sum(ge.y); ge.z = sub_sum(ge.z);
    if (i >= V || sub != 0) return;
    const float iv = 1.0f / (float)V;
    const F8 yself = ld256(ys + 2 * (size_t)i);
    const float deg = yself.a.w;  // yhat_i = (yhat_i / deg_i) * deg_i
    const float s_edge = cfg.edge_weight / (3.0f * (float)cfg.T);
    const float s_delta = 2.0f * cfg.delta_weight / (float)V;
    const float* dp = delta + 3 * (size_t)i;
    const float3 lapv = make_float3((lv.x - yself.a.x * deg) * iv, (lv.y - yself.a.y * deg) * iv, (lv.z - yself.a.z * deg) * iv);
    const float3 lapa = make_float3((la.x - yself.b.x * deg) * iv, (la.y - yself.b.y * deg) * iv, (la.z - yself.b.z * deg) * iv);
    greg[2 * (size_t)i] = make_float4(cfg.lap_weight * lapv.x + s_edge * ge.x + s_delta * dp[0],
                                      cfg.lap_weight * lapv.y + s_edge * ge.y + s_delta * dp[1],
                                      cfg.lap_weight * lapv.z + s_edge * ge.z + s_delta * dp[2], 0.f);
    greg[2 * (size_t)i + 1] = make_float4(cfg.albedo_weight * lapa.x, cfg.albedo_weight * lapa.y, cfg.albedo_weight * lapa.z, 0.f);
}

// Normal backward, step 1: gradient through the per-vertex normalisation (un-normalised photometric scale; linear, scaled
// in the Adam pass), from the (all-reduced) tangential accumulators.  One thread per vertex.
__global__ void __launch_bounds__(256) ham_normal_grad_kernel(int V, float4* __restrict__ vg,
                                                              const float4* __restrict__ vattr,
                                                              const float4* __restrict__ raw4,
                                                              const float4* __restrict__ packed4) {
    FMHR_TRACE_SCOPE(11);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const float4 ga = packed4[2 * (size_t)i], gb = packed4[2 * (size_t)i + 1];
    const float4 nrm = __ldg(vattr + 2 * (size_t)i);
    const float4 N = raw4[i];
    float3 r;
    if (nrm.w == 0.0f) {  // |N| > 1e-6: tangential part of the gradient / |N|
        float3 t1, t2;
        tangent_frame(make_float3(nrm.x, nrm.y, nrm.z), t1, t2);
        const float inv = 1.0f / N.w;
        r = make_float3((t1.x * ga.w + t2.x * gb.x) * inv, (t1.y * ga.w + t2.y * gb.x) * inv,
                        (t1.z * ga.w + t2.z * gb.x) * inv);
    } else {
        const float gz = packed4[2 * (size_t)V + i].w;
        r = make_float3(ga.w * 1e6f, gb.x * 1e6f, gz * 1e6f);
    }
    vg[2 * (size_t)i + 1] = make_float4(r.x, r.y, r.z, 0.f);
}

// ---------------------------------------------------------------------------------------------------------------------
// Multi-GPU exchange over NVLink peer memory, fused into the first kernel of the update (the one exchange of
// the path is the sum of `packed` over the ranks).  Every rank's `packed` lives in a cudaIpc-shared allocation; this
// kernel (a) tells every peer that this rank's accumulators of the step are complete, (b) waits for the same word from
// every peer, then (c) each thread sums ITS vertex's three accumulator float4s straight out of every rank's buffer (one-shot
// gather in rank order, so all replicas get bit-identical sums), stores the sums into the local `reduced` buffer the Adam
// pass reads, and does the normal-gradient step of that vertex on them.  `packed` alternates between two buffers per step
// (the z-buffer slot parity), so a rank never re-arms a buffer a slower peer may still be reading: before it can reach the
// prep kernel two steps later it has waited for that peer's flag of the step in between.
struct HamPeerArgs {
    const float4* packed[FMHR_MAX_PEERS];  // this step's packed buffer of every rank (own rank included)
    uint32_t* signal[FMHR_MAX_PEERS];      // word [rank] of every rank's flag array: where this rank posts its step count
    const uint32_t* flags;                 // this rank's flag array [FMHR_MAX_PEERS], word r is posted by rank r
    const uint32_t* epoch;                 // steps this rank has completed (device counter, bumped by the Adam pass)
    float4* reduced;                       // [3V + 1] sums over the ranks (this rank's copy)
    int world;
    // two-shot form (more than two ranks): rank k sums chunk k of the vertices and stores it into every rank's `reduced`
    float4* reduced_all[FMHR_MAX_PEERS];
    uint32_t* signal_b[FMHR_MAX_PEERS];    // second flag word per rank pair: "my chunk has landed in your `reduced`"
    const uint32_t* flags_b;
    int rank, chunk;                       // chunk = ceil(V / world) vertices per rank
    long long timeout_cycles;              // rendezvous give-up time (SM cycles)
    int post_mode;                         // see peer_rendezvous
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_peer(const float4* p) {  // peer lines must never be served from this SM's L1
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// A peer that never posts its step must not hang the device: after `timeout_cycles` SM cycles (fmhr_ham_peers.timeout_s,
// default 30 s) the rendezvous gives up and raises status bit 2.  That is FATAL and non-destructive: the update kernels
// then leave parameters, Adam state and the step counters untouched, poison the loss record with NaN and latch
// adam_step[3] (sticky across steps: every later update refuses as well) until the host has seen it
// (HamOptimizer.check_health raises).
constexpr long long kPeerCyclesPerSecond = 2000000000ll;
constexpr int kPeerDefaultTimeoutS = 30;

__device__ __forceinline__ void st_peer(float4* p, float4 v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Block-level rendezvous with every peer: lanes 0..world-1 of BLOCK 0 post this rank's step count into the peers' flag
// arrays (release at system scope: everything this rank wrote in earlier kernels of the stream is visible to a peer that
// has seen the word), lanes 0..world-1 of EVERY block wait for the peers' words.  A block only ever waits for a peer's
// block 0, which posts unconditionally when it starts - never for another block of its own grid.
__device__ __forceinline__ void peer_rendezvous(uint32_t* const* signal, const uint32_t* flags, int world, uint32_t want,
                                                int* status, long long timeout_cycles, int trace_slot, int post_mode) {
    FMHR_TRACE_MIN(trace_slot, 0);  // first block of this rank posts
    if (threadIdx.x < world) {
        // post_mode 0: every block posts behind a system-scope fence (round 1); 1: every block posts, the release store alone
        // orders the earlier kernels' writes; 2: only block 0 posts (one remote store per peer instead of one per block)
        if (post_mode == 0) __threadfence_system();
        if (post_mode != 2 || blockIdx.x == 0) st_release_sys(signal[threadIdx.x], want);
        const uint32_t* f = flags + threadIdx.x;
        const long long t0 = clock64();
        while ((int32_t)(ld_acquire_sys(f) - want) < 0) {
            if (clock64() - t0 > timeout_cycles) { atomicOr(status, 4); break; }
        }
    }
    __syncthreads();
    FMHR_TRACE_MIN(trace_slot + 1, 0);  // first block that has seen every peer
    FMHR_TRACE_MAX(trace_slot, 1);      // last block that has seen every peer
}

// Two-shot exchange, kernel 1 (reduce-scatter + all-gather by peer stores): after the rendezvous thread j sums vertex
// rank*chunk + j over every rank's `packed` (rank order) and stores the three sums into EVERY rank's `reduced`.
__global__ void __launch_bounds__(256) ham_peer_reduce_scatter_kernel(int V, HamPeerArgs pa, int* __restrict__ status) {
    FMHR_TRACE_SCOPE(13);
    peer_rendezvous(pa.signal, pa.flags, pa.world, *pa.epoch + 1u, status, pa.timeout_cycles, 20, pa.post_mode);
    // One thread per float4 of this rank's chunk (3 per vertex: the two halves of G8 and Gm - both contiguous ranges of the
    // packed buffer), every rank's copy loaded in ONE round of up to 16 independent NVLink reads, summed in rank order and
    // stored into every rank's `reduced`.  (One thread per vertex with the ranks taken four at a time was 3 float4 x 2
    // dependent round trips deep on 1/3 of the threads: 13 us of the 8-rank step for 300 kB per rank.)
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int v0 = pa.rank * pa.chunk, nv = max(0, min(pa.chunk, V - v0));
    size_t idx;
    if (j < 2 * nv) idx = 2 * (size_t)v0 + j;                          // G8[2 v0 .. 2 (v0 + nv))
    else if (j < 3 * nv) idx = 2 * (size_t)V + v0 + (j - 2 * nv);      // Gm[v0 .. v0 + nv)
    else if (j == 3 * nv && pa.rank == 0) idx = 3 * (size_t)V;         // the four loss scalars: rank 0
    else return;
    float4 x[FMHR_MAX_PEERS];
#pragma unroll
    for (int r = 0; r < FMHR_MAX_PEERS; r++)
        if (r < pa.world) x[r] = ld_peer(pa.packed[r] + idx);
    float4 sum = x[0];
#pragma unroll
    for (int r = 1; r < FMHR_MAX_PEERS; r++)
        if (r < pa.world) sum = add4(sum, x[r]);
#pragma unroll
    for (int r = 0; r < FMHR_MAX_PEERS; r++)
        if (r < pa.world) st_peer(pa.reduced_all[r] + idx, sum);
}

// Two-shot exchange, kernel 2: second rendezvous (this rank's chunk is complete - the kernel boundary - and so is every
// peer's), then the normal-gradient step from the local `reduced`.
__global__ void __launch_bounds__(256) ham_peer_normal_grad_kernel(int V, float4* __restrict__ vg,
                                                                   const float4* __restrict__ vattr,
                                                                   const float4* __restrict__ raw4, HamPeerArgs pa,
                                                                   int* __restrict__ status) {
    FMHR_TRACE_SCOPE(14);
    peer_rendezvous(pa.signal_b, pa.flags_b, pa.world, *pa.epoch + 1u, status, pa.timeout_cycles, 22, pa.post_mode);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const float4 ga = ld_peer(pa.reduced + 2 * (size_t)i), gb = ld_peer(pa.reduced + 2 * (size_t)i + 1);
    const float4 nrm = __ldg(vattr + 2 * (size_t)i);
    const float4 N = raw4[i];
    float3 r;
    if (nrm.w == 0.0f) {
        float3 t1, t2;
        tangent_frame(make_float3(nrm.x, nrm.y, nrm.z), t1, t2);
        const float inv = 1.0f / N.w;
        r = make_float3((t1.x * ga.w + t2.x * gb.x) * inv, (t1.y * ga.w + t2.y * gb.x) * inv,
                        (t1.z * ga.w + t2.z * gb.x) * inv);
    } else {
        const float gz = ld_peer(pa.reduced + 2 * (size_t)V + i).w;
        r = make_float3(ga.w * 1e6f, gb.x * 1e6f, gz * 1e6f);
    }
    vg[2 * (size_t)i + 1] = make_float4(r.x, r.y, r.z, 0.f);
}

__global__ void __launch_bounds__(256) ham_peer_reduce_normal_grad_kernel(int V, float4* __restrict__ vg,
                                                                          const float4* __restrict__ vattr,
                                                                          const float4* __restrict__ raw4, HamPeerArgs pa,
                                                                          int* __restrict__ status) {
    FMHR_TRACE_SCOPE(15);
    peer_rendezvous(pa.signal, pa.flags, pa.world, *pa.epoch + 1u, status, pa.timeout_cycles, 20, pa.post_mode);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > V) return;
    if (i == V) {  // the four loss scalars behind the accumulators
        float4 s = ld_peer(pa.packed[0] + 3 * (size_t)V);
        for (int r = 1; r < pa.world; r++) s = add4(s, ld_peer(pa.packed[r] + 3 * (size_t)V));
        pa.reduced[3 * (size_t)V] = s;
        return;
    }
    // four ranks' lines in flight per round trip; the sums run in rank order on every rank
    float4 ga = make_float4(0.f, 0.f, 0.f, 0.f), gb = ga, gm = ga;
    for (int r0 = 0; r0 < pa.world; r0 += 4) {
        float4 a[4], b[4], m[4];
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (r0 + k < pa.world) {
                const float4* p = pa.packed[r0 + k];
                a[k] = ld_peer(p + 2 * (size_t)i);
                b[k] = ld_peer(p + 2 * (size_t)i + 1);
                m[k] = ld_peer(p + 2 * (size_t)V + i);
            }
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (r0 + k < pa.world) { ga = add4(ga, a[k]); gb = add4(gb, b[k]); gm = add4(gm, m[k]); }
    }
    pa.reduced[2 * (size_t)i] = ga;
    pa.reduced[2 * (size_t)i + 1] = gb;
    pa.reduced[2 * (size_t)V + i] = gm;
    const float4 nrm = __ldg(vattr + 2 * (size_t)i);
    const float4 N = raw4[i];
    float3 r;
    if (nrm.w == 0.0f) {
        float3 t1, t2;
        tangent_frame(make_float3(nrm.x, nrm.y, nrm.z), t1, t2);
        const float inv = 1.0f / N.w;
        r = make_float3((t1.x * ga.w + t2.x * gb.x) * inv, (t1.y * ga.w + t2.y * gb.x) * inv,
                        (t1.z * ga.w + t2.z * gb.x) * inv);
    } else {
        r = make_float3(ga.w * 1e6f, gb.x * 1e6f, gm.w * 1e6f);
    }
    vg[2 * (size_t)i + 1] = make_float4(r.x, r.y, r.z, 0.f);
}

__device__ __forceinline__ float adam_update(float p, float g, float* m, float* v, float b1, float b2, float eps,
                                             float step_size, float bias2_sqrt) {
    const float mm = *m + (g - *m) * (1.0f - b1);  // exp_avg.lerp_(grad, 1 - beta1)
    const float vv = *v * b2 + (1.0f - b2) * g * g;
    *m = mm;
    *v = vv;
    const float denom = sqrtf(vv) / bias2_sqrt + eps;
    return p - step_size * (mm / denom);
}

__device__ __forceinline__ float pick3(const float3 v, int c) { return c == 0 ? v.x : (c == 1 ? v.y : v.z); }

// pass 2: gather every gradient term per vertex (kLPV lanes each), then Adam on delta (phase B) and albedo with the
// components of the vertex spread over lanes 0..2 (coalesced Adam state traffic)
__global__ void __launch_bounds__(256, 6) ham_update_pass2_kernel(
    fmhr_ham_config cfg, const float4* __restrict__ vg, float* __restrict__ delta, float* __restrict__ albedo,
    const int32_t* __restrict__ v2f_ptr, const int2* __restrict__ v2f_nbr, const int32_t* __restrict__ v2v_ptr,
    const int32_t* __restrict__ v2v_idx, const float* __restrict__ packed, const float4* __restrict__ greg,
    float* __restrict__ adam_m, float* __restrict__ adam_v, const float* __restrict__ adam_sc,
    const double* __restrict__ acc, float* __restrict__ losses, float* __restrict__ dbg_grad,
    const int* __restrict__ status, uint32_t* __restrict__ epoch_bump, int32_t* __restrict__ adam_step) {
    FMHR_TRACE_SCOPE(12);
    const int V = cfg.V;
    // a peer never arrived (this step, or latched by an earlier one): the summed gradients are incomplete - refuse the
    // whole update (parameters, Adam moments and step counters stay as they were) and poison the loss record
    const bool fatal = (status && (*status & 4)) || adam_step[3] != 0;
    if (fatal) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            for (int k = 0; k < 8; k++) losses[k] = __int_as_float(0x7fc00000);
            // the regulariser kernel advanced the counters of the parameters that step in this phase: take that back
            adam_step[0] -= (cfg.phase == 1); adam_step[1] -= 1; adam_step[2] -= (cfg.phase == 0);
            adam_step[3] = 1;
        }
        return;
    }
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) / kLPV, sub = threadIdx.x & (kLPV - 1);
    const float* scal = packed + 12 * (size_t)V;
    const float n_valid = scal[0];
    const float P_global = (float)cfg.n_views_global * (float)cfg.H * (float)cfg.W;
    const float s_photo = cfg.sfs_weight / (3.0f * n_valid);            // F.l1_loss mean over [N_valid,3]
    const float s_mask = 2.0f * cfg.mask_weight / P_global;              // F.mse_loss mean over n*H*W
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const float sfs = cfg.sfs_weight * scal[1] / (3.0f * n_valid);
        const double* reg_totals = acc + 7 * 32;  // written by ham_finalize_scalars_kernel
        const float lap = cfg.lap_weight * (float)(reg_totals[0] / (double)V);
        const float alb = cfg.albedo_weight * (float)(reg_totals[1] / (double)V);
        const float msk = cfg.phase == 1 ? cfg.mask_weight * scal[2] / P_global : 0.0f;
        const float edg = cfg.edge_weight * (float)(reg_totals[2] / (3.0 * (double)cfg.T));
        const float del = cfg.delta_weight * (float)(reg_totals[3] / (double)V);
        losses[0] = sfs; losses[1] = cfg.phase == 1 ? lap : 0.0f; losses[2] = alb; losses[3] = msk;
        losses[4] = cfg.phase == 1 ? edg : 0.0f; losses[5] = cfg.phase == 1 ? del : 0.0f; losses[6] = n_valid;
        losses[7] = cfg.phase == 1 ? sfs + lap + alb + msk + edg + del : sfs;
        // pair list (bit 0) or ring list (bit 1) overflow: pixels were dropped, refuse a number
        if (status && (*status & 3)) losses[7] = __int_as_float(0x7fc00000);
        if (epoch_bump) *epoch_bump += 1u;  // peer exchange: this rank has consumed the step's buffers
    }
    // normal backward over the incident faces (the regulariser gradients were summed by ham_reg_grad_kernel)
    float3 gnb = make_float3(0.f, 0.f, 0.f);
    if (i < V && cfg.phase == 1) {
        const F8 self = ldg256(vg + 2 * (size_t)i);
        const float3 gi = make_float3(self.b.x, self.b.y, self.b.z);
        const int fb = __ldg(v2f_ptr + i), fe = __ldg(v2f_ptr + i + 1);
        for (int j = fb + sub; j < fe; j += kLPV) {
            const int2 nb = __ldg(v2f_nbr + j);
            const F8 A = ldg256(vg + 2 * (size_t)nb.x), B = ldg256(vg + 2 * (size_t)nb.y);
            const float3 Gs = make_float3(gi.x + A.b.x + B.b.x, gi.y + A.b.y + B.b.y, gi.z + A.b.z + B.b.z);
            const float3 ed = make_float3(A.a.x - B.a.x, A.a.y - B.a.y, A.a.z - B.a.z);
            gnb.x += ed.y * Gs.z - ed.z * Gs.y; gnb.y += ed.z * Gs.x - ed.x * Gs.z; gnb.z += ed.x * Gs.y - ed.y * Gs.x;
        }
    }
    // butterfly sums: every lane of the vertex's group ends up with the totals
    if (cfg.phase == 1) { gnb.x = sub_sum(gnb.x); gnb.y = sub_sum(gnb.y); gnb.z = sub_sum(gnb.z); }
    if (i >= V || sub >= 3) return;
    // lane c of the vertex's group owns component c of delta and of albedo
    const int c = sub;
    const size_t k = 3 * (size_t)i + c;
    // albedo gradient (phase A: photometric only, mesh_sfs_optim.py:233; phase B adds the albedo Laplacian)
    float ga = s_photo * packed[8 * (size_t)i + 5 + c];
    float gd = 0.0f;
    if (cfg.phase == 1) {
        const float* gr = reinterpret_cast<const float*>(greg + 2 * (size_t)i);
        ga += gr[4 + c];
        gd = s_photo * (packed[8 * (size_t)i + c] + pick3(gnb, c)) + s_mask * packed[8 * (size_t)V + 4 * (size_t)i + c] + gr[c];
    }
    if (dbg_grad) { dbg_grad[6 * (size_t)i + c] = gd; dbg_grad[6 * (size_t)i + 3 + c] = ga; }
    if (cfg.phase == 1)
        delta[k] = adam_update(delta[k], gd, adam_m + k, adam_v + k, cfg.beta1, cfg.beta2, cfg.eps, adam_sc[0], adam_sc[1]);
    const size_t ka = 3 * (size_t)V + k;
    albedo[k] = adam_update(albedo[k], ga, adam_m + ka, adam_v + ka, cfg.beta1, cfg.beta2, cfg.eps, adam_sc[2], adam_sc[3]);
}

// phase A: Adam on sh_coeffs.  torch.optim.Adam steps the WHOLE [num,9] tensor every iteration (rows of views outside
// the batch have zero gradient but keep moving on their momentum, mesh_sfs_optim.py:193,205,237), so this runs over
// every SH row resident on this rank; rows are view-local and never reduced across ranks.
__global__ void ham_update_sh_kernel(fmhr_ham_config cfg, const float* __restrict__ gsh,
                                     const float* __restrict__ packed, float* __restrict__ sh_coeffs,
                                     float* __restrict__ adam_m, float* __restrict__ adam_v,
                                     const float* __restrict__ adam_sc, float* __restrict__ dbg_grad_sh,
                                     const int32_t* __restrict__ adam_step) {
    FMHR_TRACE_SCOPE(17);
    if (adam_step[3] != 0) return;  // latched by ham_update_pass2_kernel (same stream, earlier): a peer never arrived
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cfg.n_sh_rows * 9) return;
    const float n_valid = packed[12 * (size_t)cfg.V];
    const float g = cfg.sfs_weight / (3.0f * n_valid) * gsh[k];
    if (dbg_grad_sh) dbg_grad_sh[k] = g;
    const size_t ks = 6 * (size_t)cfg.V + k;
    sh_coeffs[k] = adam_update(sh_coeffs[k], g, adam_m + ks, adam_v + ks, cfg.beta1, cfg.beta2, cfg.eps, adam_sc[4],
                               adam_sc[5]);
}

// Persistent pixel kernels: the grid is exactly the number of co-resident 256-thread blocks (SMs x occupancy), so the
// static round-robin over the work list is balanced (a grid larger than that runs a second, idle-tailed wave).
template <typename K>
static int persistent_blocks(K kernel) {
    int dev = 0, sms = 148, per_sm = 2;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kTile * kTile, 0) != cudaSuccess || per_sm < 1)
        per_sm = 2;
    return sms * per_sm;
}

static int check_cfg(const fmhr_ham_config* c) {
    FMHR_CHECK_ARG(c != nullptr);
    FMHR_CHECK_ARG(c->V > 0 && c->T > 0 && c->H > 0 && c->W > 0 && c->n_views > 0 && c->n_views_global >= c->n_views);
    FMHR_CHECK_ARG(c->phase == 0 || c->phase == 1);
    FMHR_CHECK_ARG(c->n_sh_rows >= 1);
    FMHR_CHECK_ARG(c->zbuf_slot == 0 || c->zbuf_slot == 1);
    FMHR_CHECK_ARG(c->view_groups >= 0 && c->view_groups <= 4);
    FMHR_CHECK_ARG(c->T < (1 << 28));  // triangle id shares the key's low word with 4 tag bits
    FMHR_CHECK_ARG(c->n_views < 4096 && c->W <= 16384 && c->H <= 16384);  // work-list entry = view:12 | ty:10 | tx:10
    FMHR_CHECK_ARG(c->n_views_capacity == 0 || (c->n_views_capacity >= c->n_views && c->n_views_capacity < 4096));
    FMHR_CHECK_ARG((long long)c->H * c->W < (1ll << 31));
    return FMHR_OK;
}

}  // namespace fmhr

using namespace fmhr;

extern "C" size_t fmhr_ham_workspace_bytes(const fmhr_ham_config* cfg) {
    if (!cfg || check_cfg(cfg) != FMHR_OK) return 0;
    return ham_layout(cfg, nullptr, nullptr);
}

extern "C" size_t fmhr_ham_packed_floats(const fmhr_ham_config* cfg) {
    if (!cfg || cfg->V <= 0) return 0;
    return 12 * (size_t)cfg->V + 4;
}

static int ham_check_buffers(const fmhr_ham_config* cfg, const fmhr_ham_buffers* b) {
    FMHR_CHECK_ARG(b != nullptr);
    FMHR_CHECK_ARG(b->tri && b->opp && b->v2f_ptr && b->v2f_idx && b->v2v_ptr && b->v2v_idx && b->v2f_nbr && b->inv_deg);
    FMHR_CHECK_ARG(b->vertices_tmp && b->delta && b->albedo && b->sh_coeffs && b->adam_m && b->adam_v && b->adam_step);
    FMHR_CHECK_ARG(b->imgs && b->masks && b->valid_masks && b->w2cs && b->projs && b->view_idx && b->view_vm2);
    FMHR_CHECK_ARG(b->packed && b->losses && b->workspace);
    FMHR_CHECK_ARG(b->ml_vptr && b->ml_verts && b->ml_tri2 && b->ml_pos && b->n_meshlets > 0);
    FMHR_CHECK_ARG((b->ml_tris == 256 || b->ml_tris == 512 || b->ml_tris == 1024) && b->ml_max_verts > 0 && b->ml_max_verts <= 1024);
    FMHR_CHECK_ARG(b->workspace_bytes >= fmhr_ham_workspace_bytes(cfg));
    FMHR_CHECK_ARG(((uintptr_t)b->packed & 15) == 0 && ((uintptr_t)b->workspace & 255) == 0);
    return FMHR_OK;
}

// 8-bit host batch -> the float planes the kernels read: img = u8 / 255 (one IEEE divide, bit-identical to the loader's
// `img.astype(float32) / 255`, get_data.py:77-90), mask = u8 > 127 (get_data.py:78-79).  Four bytes per thread.
__global__ void __launch_bounds__(256) ham_u8_to_f32_kernel(const uchar4* __restrict__ src, size_t n_img4, size_t n_all4,
                                                            float4* __restrict__ imgs, float4* __restrict__ masks) {
    FMHR_TRACE_SCOPE(16);
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_all4) return;
    const uchar4 b = src[i];
    if (i < n_img4) {
        imgs[i] = make_float4(__fdiv_rn((float)b.x, 255.0f), __fdiv_rn((float)b.y, 255.0f), __fdiv_rn((float)b.z, 255.0f),
                              __fdiv_rn((float)b.w, 255.0f));
    } else {
        masks[i - n_img4] = make_float4(b.x > 127 ? 1.0f : 0.0f, b.y > 127 ? 1.0f : 0.0f, b.z > 127 ? 1.0f : 0.0f,
                                        b.w > 127 ? 1.0f : 0.0f);
    }
}

// Box rows of a pinned host batch -> device staging, read by the SMs straight from host memory (zero-copy) with 16-byte
// loads: one warp per (view, row), lane l takes 16-byte chunk l, l + 32, ... of the row segment (aligned down / up to 16
// bytes; the few extra bytes are genuine neighbours of the same planes).  tools/ubench/hostread_bench.cu: such a kernel
// reads 270-450-byte row segments at 38-43 GB/s and contiguous memory at the DMA engines' rate, while ~100 separate
// row-pitched cudaMemcpy2DAsync operations per step cost ~4 us of copy-engine time EACH.
__device__ __forceinline__ uint4 ld_host16(const uint8_t* p) {
    uint4 v;
    asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__global__ void __launch_bounds__(256) ham_pull_boxes_kernel(const uint8_t* __restrict__ imgs_host,
                                                             const uint8_t* __restrict__ masks_host,
                                                             const int4* __restrict__ boxes, int H, int W,
                                                             uint8_t* __restrict__ st_img, uint8_t* __restrict__ st_msk) {
    const int v = blockIdx.y;
    const int4 b = boxes[v];  // y0, y1, x0, x1
    if (b.y <= b.x || b.w <= b.z) return;
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = gridDim.x * (blockDim.x >> 5);
    for (int r = b.x + warp; r < b.y; r += nw) {
        const size_t p0 = ((size_t)v * H + r) * W + b.z, p1 = p0 + (size_t)(b.w - b.z);
        for (size_t c = ((3 * p0) & ~(size_t)15) + 16 * (size_t)lane; c < 3 * p1; c += 512)
            *reinterpret_cast<uint4*>(st_img + c) = ld_host16(imgs_host + c);
        for (size_t c = (p0 & ~(size_t)15) + 16 * (size_t)lane; c < p1; c += 512)
            *reinterpret_cast<uint4*>(st_msk + c) = ld_host16(masks_host + c);
    }
}

// Converted-on-arrival form: the same row pull, but every 16-byte chunk is written as 16 floats into the planes the
// kernels read (img = u8 / 255 with one IEEE divide per byte, mask = u8 > 127: bit-identical to ham_u8_to_f32_kernel; the
// byte index of the u8 planes IS the float index of the f32 planes).  blockIdx.x == gridDim.x - 1 of view 0 also pulls the
// step's cameras.  Chunks are aligned down / up to 16 bytes: the extra bytes are genuine neighbours (mask bytes outside a
// box are <= 127 by the definition of the box, so they convert to the 0 the plane already holds).
__device__ __forceinline__ void st_u8x16_as_f32(float* dst, uint4 v, bool mask) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t b0 = w[k] & 255u, b1 = (w[k] >> 8) & 255u, b2 = (w[k] >> 16) & 255u, b3 = w[k] >> 24;
        float4 f;
        if (mask) f = make_float4(b0 > 127u ? 1.f : 0.f, b1 > 127u ? 1.f : 0.f, b2 > 127u ? 1.f : 0.f, b3 > 127u ? 1.f : 0.f);
        else f = make_float4(__fdiv_rn((float)b0, 255.0f), __fdiv_rn((float)b1, 255.0f), __fdiv_rn((float)b2, 255.0f),
                             __fdiv_rn((float)b3, 255.0f));
        reinterpret_cast<float4*>(dst)[k] = f;
    }
}
// Footprint: the pull takes ~110 - 160 us (PCIe-bound) and runs BESIDE the step that is in flight.  Its warps sit on
// system-memory loads for microseconds, and the kernels that share the chip with them slow down by a roughly constant 35 - 40 us
// per step wherever the pull lands (CUPTI timelines, profiles/r2/README.md): 384 x 256 threads beside the shade pass -> shade
// 44 -> 107 us; 96 x 128 threads beside shade / antialias -> 62 / 48 instead of 44 / 35 us; one 64-thread block on EVERY SM
// beside the coverage kernel -> coverage 66 -> 196 us; 48 ... 128 blocks of 256 threads launched ahead of the step -> coverage 92,
// scan 23, shade 47 us, the end-to-end rate flat at 4,190 - 4,210 iters/s; fewer than 48 blocks cannot keep the PCIe link busy
// (16 blocks: 350 us per pull).  So: kPullBlocks blocks of 256 threads, each walking several views, launched AHEAD of the
// step's first kernel (the box table is uploaded before the copy stream waits for the plane set).  No memset node in front of
// it either: the mask plane is zero outside the boxes of the batch that was pulled into this plane set before (prev), so the
// kernel zeroes those rows' chunks itself - except the chunks the new rows write - and then pulls.  Chunk = 16 consecutive
// elements of the flat plane (the byte index of the u8 planes IS the element index of the float planes); a row segment is
// written as the whole chunks it touches.
constexpr int kPullThreads = 256, kPullBlocks = 64;
__device__ __forceinline__ bool pull_owns_chunk(size_t c, size_t vbase, int W, int4 nb) {
    // is chunk [c, c + 16) written by the pull of box nb = (y0, y1, x0, x1) of the view starting at element vbase?
    const size_t e0 = c - vbase;
    const int r0 = (int)(e0 / (size_t)W), r1 = (int)((e0 + 15) / (size_t)W);
    for (int r = r0; r <= r1; r++) {
        if (r < nb.x || r >= nb.y) continue;
        const size_t lo = (size_t)r * W + nb.z, hi = (size_t)r * W + nb.w;
        if (e0 < hi && e0 + 16 > lo) return true;
    }
    return false;
}
__global__ void __launch_bounds__(kPullThreads) ham_pull_boxes_f32_kernel(
    const uint8_t* __restrict__ imgs_host, const uint8_t* __restrict__ masks_host, const int4* __restrict__ boxes,
    const int4* __restrict__ prev_boxes, int H, int W, float* __restrict__ imgs, float* __restrict__ masks,
    const float* __restrict__ w2cs_host, const float* __restrict__ projs_host, float* __restrict__ w2cs,
    float* __restrict__ projs, int n_cam_floats, int n_views) {
    if (blockIdx.x == gridDim.x - 1) {
        for (int i = threadIdx.x; i < n_cam_floats; i += blockDim.x) {
            float a, b;
            asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(a) : "l"(w2cs_host + i) : "memory");
            asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(b) : "l"(projs_host + i) : "memory");
            w2cs[i] = a;
            projs[i] = b;
        }
    }
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;  // a view is walked by the warps of ONE block
    for (int v = blockIdx.x; v < n_views; v += gridDim.x) {
    const int4 b = boxes[v];  // y0, y1, x0, x1
    const size_t vbase = (size_t)v * H * W;
    const bool have = b.y > b.x && b.w > b.z;
    // 1. the pull: two rows per warp iteration, all (up to) six loads in flight before the first store
    if (have) {
        const size_t width = (size_t)(b.w - b.z);
        for (int r = b.x + warp; r < b.y; r += 2 * nw) {
            const bool two = r + nw < b.y;
            const size_t p0 = vbase + (size_t)r * W + b.z, q0 = p0 + (size_t)nw * W;
            if (3 * width <= 1024 - 16) {
                const size_t ci = ((3 * p0) & ~(size_t)15) + 16 * (size_t)lane, cm = (p0 & ~(size_t)15) + 16 * (size_t)lane;
                const size_t di = ((3 * q0) & ~(size_t)15) + 16 * (size_t)lane, dm = (q0 & ~(size_t)15) + 16 * (size_t)lane;
                const size_t ei = 3 * (p0 + width), fi = 3 * (q0 + width);
                const bool l0 = ci < ei, l1 = ci + 512 < ei, lm = cm < p0 + width;
                const bool k0 = two && di < fi, k1 = two && di + 512 < fi, km = two && dm < q0 + width;
                uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0, a2 = a0, a3 = a0, a4 = a0, a5 = a0;
                if (l0) a0 = ld_host16(imgs_host + ci);
                if (l1) a1 = ld_host16(imgs_host + ci + 512);
                if (lm) a2 = ld_host16(masks_host + cm);
                if (k0) a3 = ld_host16(imgs_host + di);
                if (k1) a4 = ld_host16(imgs_host + di + 512);
                if (km) a5 = ld_host16(masks_host + dm);
                if (l0) st_u8x16_as_f32(imgs + ci, a0, false);
                if (l1) st_u8x16_as_f32(imgs + ci + 512, a1, false);
                if (lm) st_u8x16_as_f32(masks + cm, a2, true);
                if (k0) st_u8x16_as_f32(imgs + di, a3, false);
                if (k1) st_u8x16_as_f32(imgs + di + 512, a4, false);
                if (km) st_u8x16_as_f32(masks + dm, a5, true);
            } else {
                for (int k = 0; k < (two ? 2 : 1); k++) {
                    const size_t s0 = k ? q0 : p0, s1 = s0 + width;
                    for (size_t c = ((3 * s0) & ~(size_t)15) + 16 * (size_t)lane; c < 3 * s1; c += 512)
                        st_u8x16_as_f32(imgs + c, ld_host16(imgs_host + c), false);
                    for (size_t c = (s0 & ~(size_t)15) + 16 * (size_t)lane; c < s1; c += 512)
                        st_u8x16_as_f32(masks + c, ld_host16(masks_host + c), true);
                }
            }
        }
    }
    // 2. zero what the previous batch left in this view's mask plane and this batch does not overwrite
    const int4 pb = prev_boxes[v];
    if (pb.y > pb.x && pb.w > pb.z) {
        const int4 nb = have ? b : make_int4(0, 0, 0, 0);
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = pb.x + warp; r < pb.y; r += nw) {
            const size_t p0 = vbase + (size_t)r * W + pb.z, p1 = p0 + (size_t)(pb.w - pb.z);
            for (size_t c = (p0 & ~(size_t)15) + 16 * (size_t)lane; c < p1; c += 512) {
                if (pull_owns_chunk(c, vbase, W, nb)) continue;
                float4* d = reinterpret_cast<float4*>(masks + c);
                d[0] = z; d[1] = z; d[2] = z; d[3] = z;
            }
        }
    }
    }  // views of this block
}

// Side stream for work that is independent of the rendering chain (forked / joined with events, so it is captured into
// the same CUDA graph when the caller's stream is being captured).
struct BoxGraph {
    const void *imgs = nullptr, *masks = nullptr;
    void* staging = nullptr;
    int n = 0, H = 0, W = 0, seen = 0;
    uint64_t hash = 0;
    cudaGraphExec_t exec = nullptr;
};
struct SideStream {
    cudaStream_t st = nullptr;                // vertex-domain kernels beside the coverage kernel (high priority)
    cudaStream_t pix = nullptr;               // pixel passes of the view groups (high priority)
    cudaEvent_t cov_done[4] = {nullptr, nullptr, nullptr, nullptr}, pix_join = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr, records = nullptr;
    cudaStream_t st2 = nullptr;               // FMHR_SIDE2: regulariser forward / gradients, forked behind the normals
    cudaEvent_t normals_done = nullptr, join2 = nullptr;
    cudaStream_t copy = nullptr;              // host-batch uploads of fmhr_ham_step_host_u8
    cudaEvent_t copy_fork = nullptr, ready = nullptr;
    // pipelined host batches (fmhr_ham_host_u8_submit): two staging slots in flight
    const void* slot_ptr[2] = {nullptr, nullptr};
    cudaEvent_t slot_ready[2] = {nullptr, nullptr}, slot_free[2] = {nullptr, nullptr};
    bool slot_used[2] = {false, false};
    BoxGraph box_graph[2];  // captured copy lists of fmhr_ham_host_u8_submit_boxes, one per staging slot
    int4* box_dev[2] = {nullptr, nullptr};  // device copies of the box tables (pull-kernel form)
    int box_cap[2] = {0, 0};
    // converted-on-arrival form: per plane set two box tables (this batch / the batch before it, whose mask rows the pull
    // kernel zeroes) and the plane set for which "mask plane == 0 outside the previous boxes" is known to hold
    int4* dbox[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    int dbox_cap[2] = {0, 0}, dbox_cur[2] = {0, 0};
    const void* dprimed[2] = {nullptr, nullptr};
    int dprimed_n[2] = {0, 0};
    int dev = -1;
};
// Host batch in flight (fmhr_ham_step_host_u8): the render chain converts it right before the first kernel that reads
// images / masks, so the PCIe transfer overlaps vertex prep, coverage and scan.
struct PendingInputs {
    const uint8_t* staging = nullptr;  // device: [n*H*W*3 image bytes | n*H*W mask bytes]
    cudaEvent_t ready = nullptr;
    cudaEvent_t consumed = nullptr;    // recorded after the conversion kernel (pipelined form: the slot may be refilled)
};
static thread_local PendingInputs g_pending;
static int side_stream(SideStream** out) {
    static SideStream side[64];
    int dev = 0;
    FMHR_CUDA(cudaGetDevice(&dev));
    FMHR_CHECK_ARG(dev >= 0 && dev < 64);
    SideStream& s = side[dev];
    if (s.dev != dev) {
        // small latency-bound kernels must not queue behind the thousands of blocks of a coverage kernel: their streams
        // get the highest priority, so the block scheduler serves them first whenever an SM slot frees up
        int prio_lo = 0, prio_hi = 0;
        FMHR_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        FMHR_CUDA(cudaStreamCreateWithPriority(&s.st, cudaStreamNonBlocking, FMHR_SIDE_PRIO ? prio_hi : prio_lo));
        FMHR_CUDA(cudaStreamCreateWithPriority(&s.pix, cudaStreamNonBlocking, prio_hi));
        for (int i = 0; i < 4; i++) FMHR_CUDA(cudaEventCreateWithFlags(&s.cov_done[i], cudaEventDisableTiming));
        FMHR_CUDA(cudaEventCreateWithFlags(&s.pix_join, cudaEventDisableTiming));
        FMHR_CUDA(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
        FMHR_CUDA(cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming));
        FMHR_CUDA(cudaEventCreateWithFlags(&s.records, cudaEventDisableTiming));
        FMHR_CUDA(cudaStreamCreateWithPriority(&s.st2, cudaStreamNonBlocking, FMHR_SIDE_PRIO ? prio_hi : prio_lo));
        FMHR_CUDA(cudaEventCreateWithFlags(&s.normals_done, cudaEventDisableTiming));
        FMHR_CUDA(cudaEventCreateWithFlags(&s.join2, cudaEventDisableTiming));
        FMHR_CUDA(cudaStreamCreateWithFlags(&s.copy, cudaStreamNonBlocking));
        FMHR_CUDA(cudaEventCreateWithFlags(&s.copy_fork, cudaEventDisableTiming));
        FMHR_CUDA(cudaEventCreateWithFlags(&s.ready, cudaEventDisableTiming));
        for (int i = 0; i < 2; i++) {
            FMHR_CUDA(cudaEventCreateWithFlags(&s.slot_ready[i], cudaEventDisableTiming));
            FMHR_CUDA(cudaEventCreateWithFlags(&s.slot_free[i], cudaEventDisableTiming));
        }
        s.dev = dev;
    }
    *out = &s;
    return FMHR_OK;
}

// fmhr_ham_init: the forward chain up to the antialias pass runs in its initialisation mode (PHASE 2 kernel)
struct InitArgs {
    const float* grayimgs;  // [num,H,W]
    double* init_acc;       // [(num + 1) * 56]
};

// View groups (phase-B training step): the batch's views are split into G consecutive groups, each with its own work lists,
// and the issue-bound coverage kernel of group g + 1 runs on the caller's stream WHILE the latency-bound pixel passes of
// group g (scan / shade + backward / antialias / pair) run on a high-priority stream - the block scheduler hands the SM
// slots that coverage blocks free up to the pixel kernels first, so the two kinds of work share every SM instead of
// taking turns.  Everything per group is pointer arithmetic on the one workspace (pixel indices are group-local).
constexpr int kMaxGroups = 4;
struct GroupView {
    int v0, nv;
    size_t p0, Pg;
    unsigned long long *zcur, *znext;
    float4* plane[4];
    const float* viewM;
    const int32_t *view_idx, *sh_idx;
    uint2* clist; uint32_t* rlist; uint32_t* ringbits; uint4* plist_a; uint32_t* plist_b; uint2* qlist;
    int *ccount, *rcount, *pcount, *qcount;
    int* tcount[2]; uint32_t* tbits[2]; uint32_t* tlist[2];
    int rcap, pcap, qcap;
};
static GroupView group_view(const fmhr_ham_config* cfg, const fmhr_ham_buffers* b, const HamWs& ws, int g, int G) {
    GroupView v;
    const int n = cfg->n_views, per = cdiv(n, G);
    const size_t hw = (size_t)cfg->H * cfg->W;
    v.v0 = min(g * per, n);
    v.nv = min(per, n - v.v0);
    v.p0 = (size_t)v.v0 * hw;
    v.Pg = (size_t)v.nv * hw;
    v.zcur = ws.zbuf[cfg->zbuf_slot] + v.p0;
    v.znext = ws.zbuf[cfg->zbuf_slot ^ 1] + v.p0;
    for (int i = 0; i < 4; i++) v.plane[i] = ws.plane[i] ? ws.plane[i] + v.p0 : nullptr;
    v.viewM = ws.viewM + (size_t)v.v0 * kViewM;
    v.view_idx = b->view_idx + v.v0;
    v.sh_idx = (b->sh_idx ? b->sh_idx : b->view_idx) + v.v0;
    v.clist = ws.clist + v.p0;
    v.rlist = ws.rlist + v.p0 / 2;
    v.ringbits = ws.ringbits + v.p0 / 32 + g;
    v.plist_a = ws.plist_a + v.p0 / 2;
    v.plist_b = ws.plist_b + v.p0 / 2;
    v.qlist = ws.qlist + v.p0 / 4;
    int* cnt = ws.ccount + 8 * g;  // group 0: the historical slots (ccount, rcount, pcount, status, qcount)
    v.ccount = cnt; v.rcount = cnt + 1; v.pcount = cnt + 2; v.qcount = cnt + 4;
    const size_t tiles_pv = (size_t)cdiv(cfg->W, kTile) * cdiv(cfg->H, kTile), words_pv = (tiles_pv + 31) / 32;
    for (int s2 = 0; s2 < 2; s2++) {
        v.tcount[s2] = ws.tcount[s2] + g;
        v.tbits[s2] = ws.tbits[s2] + (size_t)v.v0 * words_pv;
        v.tlist[s2] = ws.tlist[s2] + (size_t)v.v0 * tiles_pv;
    }
    v.rcap = (int)(v.Pg / 2); v.pcap = (int)(v.Pg / 2); v.qcap = (int)(v.Pg / 4);
    // Test hook (tests/test_gpu_ham.py::test_work_list_overflow_poisons_the_loss_record): FMHR_TEST_LIST_CAP shrinks the ring
    // and pair lists so that the overflow path - entries dropped, status bit set, losses[7] = NaN - can be exercised on a
    // normal scene.  rcap = P/2 is NOT the worst case (scattered single-pixel coverage has up to 0.8 P ring pixels): such a
    // frame is refused with the NaN flag, not silently mis-evaluated.
    static const int test_cap = [] { const char* e = getenv("FMHR_TEST_LIST_CAP"); return e ? atoi(e) : 0; }();
    if (test_cap > 0) { v.rcap = min(v.rcap, test_cap); v.pcap = min(v.pcap, test_cap); }
    return v;
}

static int launch_coverage(const fmhr_ham_config* cfg, const fmhr_ham_buffers* b, const HamWs& ws, const GroupView& gv,
                           cudaStream_t st) {
    const int H = cfg->H, W = cfg->W, cur = cfg->zbuf_slot;
    const int tiles_x = cdiv(W, kTile), tiles_pv = tiles_x * cdiv(H, kTile);
    const float invW = 1.0f / (float)W, invH = 1.0f / (float)H;
    size_t smem = (size_t)b->ml_max_verts * 24 + (size_t)((tiles_pv + 31) / 32) * sizeof(unsigned int);
    if (smem > 96 * 1024) {
        set_error("fmhr_ham_step_render: %d tiles per view exceed the coverage kernel's shared-memory bitmaps", tiles_pv);
        return FMHR_EUNSUPPORTED;
    }
    // Tuning hook: a larger dynamic shared-memory request lowers the coverage blocks per SM (5 -> 4 at 24 KB with the
    // 1024-triangle meshlets), which leaves a quarter of the register file to the side-stream vertex kernels.
    static const size_t smem_floor = [] { const char* e = getenv("FMHR_COV_SMEM"); return e ? (size_t)atoi(e) : (size_t)FMHR_COV_SMEM_FLOOR; }();
    if (smem < smem_floor && smem_floor <= 96 * 1024) smem = smem_floor;
    const dim3 grid(b->n_meshlets, gv.nv);
    const uint2* tri2 = (const uint2*)b->ml_tri2;
    // more than four frame pixels per triangle (configs 1, 3, 5): triangles span several pixels, use the draining variant
    const bool drain = (long long)H * W > 4ll * cfg->T;
#define FMHR_COVERAGE(TPT)                                                                                             \
    do {                                                                                                               \
        static bool attr_set = false;                                                                                  \
        if (!attr_set) {                                                                                               \
            FMHR_CUDA(cudaFuncSetAttribute(ham_coverage_meshlet_kernel<TPT>,                                           \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));                  \
            FMHR_CUDA(cudaFuncSetAttribute(ham_coverage_meshlet_kernel<TPT, false, true>,                              \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));                  \
            attr_set = true;                                                                                           \
        }                                                                                                              \
        if (drain)                                                                                                     \
            ham_coverage_meshlet_kernel<TPT, false, true><<<grid, kCovThreads, smem, st>>>(                            \
                (const float4*)b->ml_pos, gv.viewM, b->ml_vptr, b->ml_verts, tri2, b->ml_max_verts, H, W, invW, invH,  \
                gv.zcur,                                                                                               \
                gv.tbits[cur], gv.tlist[cur], gv.tcount[cur], tiles_x, tiles_pv, 0);                                   \
        else                                                                                                           \
            ham_coverage_meshlet_kernel<TPT><<<grid, kCovThreads, smem, st>>>(                                         \
                (const float4*)b->ml_pos, gv.viewM, b->ml_vptr, b->ml_verts, tri2, b->ml_max_verts, H, W, invW, invH,  \
                gv.zcur,                                                                                               \
                gv.tbits[cur], gv.tlist[cur], gv.tcount[cur], tiles_x, tiles_pv, 0);                                   \
    } while (0)
    if (b->ml_tris == 1024) FMHR_COVERAGE(1024 / kCovThreads);
    else if (b->ml_tris == 512) FMHR_COVERAGE(512 / kCovThreads);
    else FMHR_COVERAGE(256 / kCovThreads);
#undef FMHR_COVERAGE
    FMHR_LAUNCH_CHECK();
    return FMHR_OK;
}

static int view_groups_for(const fmhr_ham_config* cfg) {
    static const int env = [] { const char* e = getenv("FMHR_VIEW_GROUPS"); return e ? atoi(e) : 0; }();
    int G = cfg->view_groups > 0 ? cfg->view_groups : (env > 0 ? env : FMHR_DEFAULT_VIEW_GROUPS);
    G = max(1, min(G, kMaxGroups));
    while (G > 1 && cfg->n_views < 2 * G) G--;  // groups of at least two views
    return G;
}

template <int PHASE>
static int ham_render_impl(const fmhr_ham_config* cfg, const fmhr_ham_buffers* b, cudaStream_t st, float* dbg_image,
                           float* dbg_mask, bool forward_only, const InitArgs* init = nullptr) {
    HamWs ws;
    ham_layout(cfg, (char*)b->workspace, &ws);
    const int V = cfg->V, T = cfg->T, H = cfg->H, W = cfg->W, n = cfg->n_views;
    const size_t P = (size_t)n * H * W;
    const int cur = cfg->zbuf_slot, nxt = cfg->zbuf_slot ^ 1;
    const int tiles_x = cdiv(W, kTile), tiles_y = cdiv(H, kTile);
    const float invW = 1.0f / (float)W, invH = 1.0f / (float)H;  // IEEE single divides, identical to the device's __fdiv_rn
    // phase-B training step: the shade pass back-propagates the speculated L1 signs itself (no pixel backward pass);
    // FMHR_NO_SPEC=1 keeps the separate backward pass (A/B measurements, tests)
    static const bool no_spec = [] { const char* e = getenv("FMHR_NO_SPEC"); return e && e[0] == '1'; }();
    const bool spec = PHASE == 1 && !forward_only && !init && !no_spec;
    FMHR_STAGE_MARK();  // 0: (no clear pass any more)
    // zeroed by the prep kernel: packed, loss accumulators + dilated work list (common), the work list of the slot
    // rasterised this step, SH gradients (phase A)
    const int n0 = (int)(ws.common_bytes / 4), n1 = (int)(ws.slot_bytes / 4), n2 = PHASE == 0 ? cfg->n_sh_rows * 9 : 0;
    const int ml_total = b->n_meshlets * b->ml_max_verts;
    const int prep_threads = max(max(max(3 * V + 1, n * kViewM), max(n0, max(n1, n2))), ml_total);
    ham_vertex_prep_kernel<<<cdiv(prep_threads, 256), 256, 0, st>>>(
        b->vertices_tmp, b->delta, V, ws.vg, (float4*)b->packed, (uint32_t*)ws.common_region, n0,
        (uint32_t*)ws.slot_region[cfg->zbuf_slot], n1, (uint32_t*)ws.gsh, n2, b->w2cs, b->projs, b->view_idx, n, ws.viewM,
        b->ml_vptr, b->ml_verts, b->ml_max_verts, ml_total, (float4*)b->ml_pos);
    FMHR_LAUNCH_CHECK();
    // The coverage + scan kernels only need the vertices and the view matrices (prep).  Normals -> per-triangle records
    // (needed from the shade pass on) and the regulariser forward / backward + Adam scalars (needed by the update) run on a
    // side stream concurrently with them: latency-bound vertex kernels under the issue-bound coverage kernel.  Inline when
    // the stages are being timed; the regulariser is skipped by the forward-only inspection path (it advances the counters).
    SideStream* side = nullptr;
    cudaStream_t vs = st;
    if (!g_timer) {
        int rc_ = side_stream(&side);
        if (rc_) return rc_;
        FMHR_CUDA(cudaEventRecord(side->fork, st));
        FMHR_CUDA(cudaStreamWaitEvent(side->st, side->fork, 0));
        vs = side->st;
    }
    const int G = (spec && side) ? view_groups_for(cfg) : 1;
    GroupView gv[kMaxGroups];
    for (int g = 0; g < G; g++) gv[g] = group_view(cfg, b, ws, g, G);
    // coverage first in issue order: it is the head of the critical path (the vertex kernels below wait for SM slots anyway)
    if (!g_timer) {
        for (int g = 0; g < G; g++) {
            int rc_ = launch_coverage(cfg, b, ws, gv[g], st);
            if (rc_) return rc_;
            if (G > 1) FMHR_CUDA(cudaEventRecord(side->cov_done[g], st));
        }
    }
    ham_normals_kernel<<<cdiv((long long)V * 4, 256), 256, 0, vs>>>(ws.vg, b->albedo, b->v2f_ptr, (const int2*)b->v2f_nbr, V,
                                                                    ws.vattr, ws.raw4);
    FMHR_LAUNCH_CHECK();
    // FMHR_REG_GRAD_LATE: the regulariser GRADIENTS (only the update needs them) are held back until the antialias pass has
    // finished and run beside the pair kernel: launched right behind the regulariser they trickle into the SMs through the
    // whole shade pass and slow down exactly the blocks that finish last
    const bool late_rg = FMHR_REG_GRAD_LATE && side && G == 1 && PHASE == 1 && !forward_only && !init;
    cudaStream_t rs = vs;  // stream of the regulariser chain
    if (side && FMHR_SIDE2 && !forward_only) {
        FMHR_CUDA(cudaEventRecord(side->normals_done, side->st));
        FMHR_CUDA(cudaStreamWaitEvent(side->st2, side->normals_done, 0));
        rs = side->st2;
    }
    ham_trirec_kernel<<<cdiv(T, 128), 128, 0, vs>>>(b->tri, b->opp, ws.vg, ws.vattr, V, T, ws.trirec);
    FMHR_LAUNCH_CHECK();
    if (side) FMHR_CUDA(cudaEventRecord(side->records, side->st));
    if (!forward_only) {
        ham_regulariser_kernel<<<cdiv((long long)V * kLPV, 256), 256, 0, rs>>>(
            *cfg, ws.vg, b->delta, ws.vattr, b->v2f_ptr, (const int2*)b->v2f_nbr, b->v2v_ptr, b->v2v_idx, ws.ys, ws.acc,
            b->adam_step, ws.adam_sc);
        FMHR_LAUNCH_CHECK();
        if (PHASE == 1 && !late_rg) {
            ham_reg_grad_kernel<<<cdiv((long long)V * kLPV, 256), 256, 0, rs>>>(
                *cfg, ws.vg, b->delta, b->v2f_ptr, (const int2*)b->v2f_nbr, b->v2v_ptr, b->v2v_idx, ws.ys, ws.greg);
            FMHR_LAUNCH_CHECK();
        }
    }
    if (side) FMHR_CUDA(cudaEventRecord(side->join, side->st));
    if (rs != vs) FMHR_CUDA(cudaEventRecord(side->join2, side->st2));
    FMHR_STAGE_MARK();  // 1: vertex prep + normals
    FMHR_STAGE_MARK();  // 2: (the clip transform is fused into the coverage kernel)
    if (g_timer) {
        int rc_ = launch_coverage(cfg, b, ws, gv[0], st);
        if (rc_) return rc_;
    }
    FMHR_STAGE_MARK();  // 3: coverage (transform + visibility)
    // the pixel passes: on the caller's stream, or (view groups) on the high-priority pixel stream behind each group's coverage
    cudaStream_t px = st;
    if (G > 1) {
        px = side->pix;
        FMHR_CUDA(cudaStreamWaitEvent(px, side->cov_done[0], 0));
    }
    if (g_pending.staging) {  // host batch of fmhr_ham_step_host_u8: first use of images / masks is the shade pass
        if (g_pending.ready) FMHR_CUDA(cudaStreamWaitEvent(px, g_pending.ready, 0));
        const size_t n_img4 = P * 3 / 4, n_all4 = n_img4 + P / 4;
        ham_u8_to_f32_kernel<<<cdiv((long long)n_all4, 256), 256, 0, px>>>((const uchar4*)g_pending.staging, n_img4, n_all4,
                                                                          (float4*)b->imgs, (float4*)b->masks);
        FMHR_LAUNCH_CHECK();
        if (g_pending.consumed) FMHR_CUDA(cudaEventRecord(g_pending.consumed, px));
        g_pending.staging = nullptr;
    }
    static const int g_scan = persistent_blocks(ham_scan_kernel);
    static const int g_shade = persistent_blocks(ham_shade_kernel<PHASE>);
    static const int g_aa = persistent_blocks(ham_aa_loss_kernel<PHASE>);
    static const int g_bwd = persistent_blocks(ham_pixel_bwd_kernel<PHASE>);
    const int pblock = 256;  // 8 warps = 8 independent workers (no block barrier in the pixel passes)
    for (int g = 0; g < G; g++) {
        const GroupView& v = gv[g];
        float* dimg = dbg_image ? dbg_image + v.p0 * 3 : nullptr;
        float* dmsk = dbg_mask ? dbg_mask + v.p0 : nullptr;
        if (G > 1 && g > 0) FMHR_CUDA(cudaStreamWaitEvent(px, side->cov_done[g], 0));
        ham_scan_kernel<<<g_scan, pblock, 0, px>>>(v.zcur, v.znext, v.tlist[cur], v.tcount[cur], v.tlist[nxt], v.tcount[nxt],
                                                   tiles_x, tiles_y, H, W, v.clist, v.ccount, v.ringbits, v.rlist, v.rcount,
                                                   v.rcap, ws.status);
        FMHR_LAUNCH_CHECK();
        if (side && g == 0) FMHR_CUDA(cudaStreamWaitEvent(px, side->records, 0));  // normals + triangle records are ready
        if (spec) {
            static const int g_shade_bwd = persistent_blocks(ham_shade_kernel<1, true>);
            ham_shade_kernel<1, true><<<g_shade_bwd, pblock, 0, px>>>(v.clist, v.ccount, v.zcur, ws.vg, v.viewM, invW, invH,
                                                                      ws.trirec, b->masks, b->sh_coeffs, v.view_idx, v.sh_idx,
                                                                      V, H, W, v.plane[0], v.plane[1], ws.acc, b->imgs,
                                                                      (float4*)b->packed);
        } else {
            ham_shade_kernel<PHASE><<<g_shade, pblock, 0, px>>>(v.clist, v.ccount, v.zcur, ws.vg, v.viewM, invW, invH,
                                                              ws.trirec, b->masks, b->sh_coeffs, v.view_idx, v.sh_idx, V, H,
                                                              W, v.plane[0], v.plane[1], ws.acc, nullptr, nullptr);
        }
        FMHR_LAUNCH_CHECK();
        FMHR_STAGE_MARK();  // 4: scan + shade
        float4* g0 = PHASE == 0 ? v.plane[2] : v.plane[1];
        float4* g1 = PHASE == 0 ? v.plane[3] : nullptr;
        if (init) {  // initialisation mode: normals + coverage antialiased, SH normal equations accumulated
            static const int g_init = persistent_blocks(ham_aa_loss_kernel<2>);
            ham_aa_loss_kernel<2><<<g_init, pblock, 0, px>>>(
                v.clist, v.ccount, v.rlist, v.rcount, v.rcap, v.ringbits, v.zcur, ws.vg, v.viewM, b->tri, b->opp,
                init->grayimgs, b->valid_masks, b->sh_coeffs, v.view_idx, v.sh_idx, V, T, H, W, v.plane[0], v.plane[1], g0, g1,
                ws.acc, ws.gsh, dimg, dmsk, v.plist_a, v.plist_b, v.pcount, v.pcap, ws.status, init->init_acc, nullptr,
                nullptr, 0);
        } else if (spec) {
            static const int g_aa_spec = persistent_blocks(ham_aa_loss_kernel<1, true>);
            ham_aa_loss_kernel<1, true><<<g_aa_spec, pblock, 0, px>>>(
                v.clist, v.ccount, v.rlist, v.rcount, v.rcap, v.ringbits, v.zcur, ws.vg, v.viewM, b->tri, b->opp, b->imgs,
                b->valid_masks, b->sh_coeffs, v.view_idx, v.sh_idx, V, T, H, W, v.plane[0], v.plane[1], g0, g1, ws.acc,
                ws.gsh, dimg, dmsk, v.plist_a, v.plist_b, v.pcount, v.pcap, ws.status, nullptr, v.qlist, v.qcount, v.qcap);
        } else {
            ham_aa_loss_kernel<PHASE><<<g_aa, pblock, 0, px>>>(
                v.clist, v.ccount, v.rlist, v.rcount, v.rcap, v.ringbits, v.zcur, ws.vg, v.viewM, b->tri, b->opp, b->imgs,
                b->valid_masks, b->sh_coeffs, v.view_idx, v.sh_idx, V, T, H, W, v.plane[0], v.plane[1], g0, g1, ws.acc,
                ws.gsh, dimg, dmsk, v.plist_a, v.plist_b, v.pcount, v.pcap, ws.status, nullptr, nullptr, nullptr, 0);
        }
        FMHR_LAUNCH_CHECK();
        FMHR_STAGE_MARK();  // 5: antialias + losses
        // After the antialias pass independent kernels remain (all only ADD into `packed`): the pair backward goes to the
        // vertex side stream (one group) or stays in the group's chain (view groups: the next group's coverage is running
        // beside it anyway); the loss-scalar finalize and, without speculation, the pixel backward stay on this stream.
        cudaStream_t ps = px;
        // the finalize kernel reads the regulariser's loss accumulators: order this stream behind the vertex side stream
        // (its last record: after the regulariser gradients, long finished by now)
        if (side && g == G - 1) FMHR_CUDA(cudaStreamWaitEvent(px, side->join, 0));
        if (rs != vs && g == G - 1) FMHR_CUDA(cudaStreamWaitEvent(px, side->join2, 0));
        if (side && G == 1) {
            FMHR_CUDA(cudaEventRecord(side->fork, px));
            FMHR_CUDA(cudaStreamWaitEvent(side->st, side->fork, 0));
            ps = side->st;
        }
        if (late_rg) {
            FMHR_CUDA(cudaStreamWaitEvent(side->st2, side->fork, 0));  // (px is already ordered behind the regulariser)
            ham_reg_grad_kernel<<<cdiv((long long)V * kLPV, 256), 256, 0, side->st2>>>(
                *cfg, ws.vg, b->delta, b->v2f_ptr, (const int2*)b->v2f_nbr, b->v2v_ptr, b->v2v_idx, ws.ys, ws.greg);
            FMHR_LAUNCH_CHECK();
            FMHR_CUDA(cudaEventRecord(side->join2, side->st2));
        }
        if (!forward_only) {
            ham_pair_bwd_kernel<PHASE><<<888, 128, 0, ps>>>(v.plist_a, v.plist_b, v.pcount, v.pcap, v.zcur, ws.vg, v.viewM, V,
                                                           H, W, v.plane[0], g0, g1, ws.trirec, invW, invH, b->sh_coeffs,
                                                           v.sh_idx, (float4*)b->packed, spec ? v.qlist : nullptr, v.qcount,
                                                           v.qcap);
            FMHR_LAUNCH_CHECK();
        }
        if (g == G - 1) {
            // loss scalars of the whole batch: beside the last pair kernel when this stream is otherwise idle, else behind it
            ham_finalize_scalars_kernel<<<1, 32, 0, (spec && G == 1) ? px : ps>>>(
                ws.acc, b->view_vm2, b->view_idx, n, tiles_x * tiles_y, PHASE, b->packed + 12 * (size_t)V, ws.acc + 7 * 32);
            FMHR_LAUNCH_CHECK();
        }
        if (side && G == 1) FMHR_CUDA(cudaEventRecord(side->join, side->st));
        if (!forward_only && !spec) {
            ham_pixel_bwd_kernel<PHASE><<<g_bwd, pblock, 0, px>>>(v.clist, v.ccount, ws.trirec, invW, invH, v.viewM,
                                                                  b->sh_coeffs, v.sh_idx, V, H, W, g0, g1, (float4*)b->packed);
            FMHR_LAUNCH_CHECK();
        }
    }
    if (G > 1) {
        FMHR_CUDA(cudaEventRecord(side->pix_join, px));
        FMHR_CUDA(cudaStreamWaitEvent(st, side->pix_join, 0));
    }
    if (side) FMHR_CUDA(cudaStreamWaitEvent(st, side->join, 0));
    if (rs != vs || late_rg) FMHR_CUDA(cudaStreamWaitEvent(st, side->join2, 0));
    FMHR_STAGE_MARK();  // 6: pixel backward (+ scalar finalize)
    return FMHR_OK;
}

// Coverage launch of the stand-alone rasteriser (raster.cu, fmhr_rasterize_fwd_meshlets): the fused path's meshlet kernel
// on caller-provided clip-space positions.
namespace fmhr {
int launch_meshlet_coverage_clip(const float* pos, int N, int V, int H, int W, const int32_t* ml_vptr,
                                 const int32_t* ml_verts, const uint32_t* ml_tri2, int n_meshlets, int ml_tris,
                                 int ml_max_verts, unsigned long long* zbuf, uint32_t* tile_bits, cudaStream_t st) {
    const int tiles_x = cdiv(W, kTile), tiles_pv = tiles_x * cdiv(H, kTile);
    const size_t smem = (size_t)ml_max_verts * 24 + (size_t)((tiles_pv + 31) / 32) * sizeof(unsigned int);
    if (smem > 96 * 1024 || (ml_tris != 256 && ml_tris != 512 && ml_tris != 1024) || ml_max_verts > 1024) {
        set_error("fmhr_rasterize_fwd_meshlets: unsupported meshlet / tile configuration");
        return FMHR_EUNSUPPORTED;
    }
    const dim3 grid(n_meshlets, N);
    const float invW = 1.0f / (float)W, invH = 1.0f / (float)H;
    const bool drain = (long long)H * W > 4ll * (long long)n_meshlets * ml_tris;  // as in the fused path
#define FMHR_COVERAGE_CLIP(TPT)                                                                                        \
    do {                                                                                                               \
        static bool attr_set = false;                                                                                  \
        if (!attr_set) {                                                                                               \
            FMHR_CUDA(cudaFuncSetAttribute(ham_coverage_meshlet_kernel<TPT, true, false>,                              \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));                  \
            FMHR_CUDA(cudaFuncSetAttribute(ham_coverage_meshlet_kernel<TPT, true, true>,                               \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));                  \
            attr_set = true;                                                                                           \
        }                                                                                                              \
        if (drain)                                                                                                     \
            ham_coverage_meshlet_kernel<TPT, true, true><<<grid, kCovThreads, smem, st>>>(                             \
                (const float4*)pos, nullptr, ml_vptr, ml_verts, (const uint2*)ml_tri2, ml_max_verts, H, W, invW, invH, \
                zbuf, tile_bits, nullptr, nullptr, tiles_x, tiles_pv, V);                                              \
        else                                                                                                           \
            ham_coverage_meshlet_kernel<TPT, true, false><<<grid, kCovThreads, smem, st>>>(                            \
                (const float4*)pos, nullptr, ml_vptr, ml_verts, (const uint2*)ml_tri2, ml_max_verts, H, W, invW, invH, \
                zbuf, tile_bits, nullptr, nullptr, tiles_x, tiles_pv, V);                                              \
    } while (0)
    if (ml_tris == 1024) FMHR_COVERAGE_CLIP(1024 / kCovThreads);
    else if (ml_tris == 512) FMHR_COVERAGE_CLIP(512 / kCovThreads);
    else FMHR_COVERAGE_CLIP(256 / kCovThreads);
#undef FMHR_COVERAGE_CLIP
    FMHR_LAUNCH_CHECK();
    return FMHR_OK;
}
}  // namespace fmhr

// Checked builds: returns and clears the mask of FMHR_DCHECK sites that fired since the last call (0 = clean); -1 in
// product builds (no checks compiled in).
extern "C" int fmhr_debug_checks(void) {
#ifdef FMHR_CHECKED
    unsigned int h = 0, z = 0;
    if (cudaDeviceSynchronize() != cudaSuccess) return -2;
    if (cudaMemcpyFromSymbol(&h, g_dcheck, sizeof(h)) != cudaSuccess) return -2;
    cudaMemcpyToSymbol(g_dcheck, &z, sizeof(z));
    return (int)h;
#else
    return -1;
#endif
}

extern "C" int fmhr_ham_reset(const fmhr_ham_config* cfg, const fmhr_ham_buffers* buf, fmhr_stream_t stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    FMHR_CHECK_ARG(buf && buf->workspace && buf->workspace_bytes >= fmhr_ham_workspace_bytes(cfg));
    HamWs ws;
    ham_layout(cfg, (char*)buf->workspace, &ws);
    const size_t P = (size_t)(cfg->n_views_capacity > cfg->n_views ? cfg->n_views_capacity : cfg->n_views) * cfg->H * cfg->W;
    for (int i = 0; i < 2; i++) {
        FMHR_CUDA(cudaMemsetAsync(ws.zbuf[i], 0xFF, P * 8, (cudaStream_t)stream));
        FMHR_CUDA(cudaMemsetAsync(ws.slot_region[i], 0, ws.slot_bytes, (cudaStream_t)stream));
    }
    FMHR_CUDA(cudaMemsetAsync(ws.common_region, 0, ws.common_bytes, (cudaStream_t)stream));
    FMHR_CUDA(cudaMemsetAsync(ws.ringbits, 0, (P / 32 + 64) * 4, (cudaStream_t)stream));
    return FMHR_OK;
}

extern "C" int fmhr_ham_prepare_views(const float* valid_masks, int num, int H, int W, double* view_vm2,
                                      fmhr_stream_t stream) {
    FMHR_CHECK_ARG(valid_masks && view_vm2 && num > 0 && H > 0 && W > 0);
    const int tx = cdiv(W, kTile), ty = cdiv(H, kTile);
    ham_view_vm2_kernel<<<dim3(tx, ty, num), dim3(kTile, kTile), 0, (cudaStream_t)stream>>>(valid_masks, H, W, view_vm2);
    FMHR_LAUNCH_CHECK();
    ham_view_vm2_total_kernel<<<num, 32, 0, (cudaStream_t)stream>>>(tx * ty, view_vm2);
    FMHR_LAUNCH_CHECK();
    return FMHR_OK;
}

extern "C" int fmhr_ham_step_render(const fmhr_ham_config* cfg, const fmhr_ham_buffers* buf, fmhr_stream_t stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    rc = ham_check_buffers(cfg, buf);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    return cfg->phase == 0 ? ham_render_impl<0>(cfg, buf, st, nullptr, nullptr, false)
                           : ham_render_impl<1>(cfg, buf, st, nullptr, nullptr, false);
}

static int ham_update_impl(const fmhr_ham_config* cfg, const fmhr_ham_buffers* buf, const fmhr_ham_peers* peers,
                           cudaStream_t st) {
    HamWs ws;
    ham_layout(cfg, (char*)buf->workspace, &ws);
    const int V = cfg->V;
    const float* packed = buf->packed;
    uint32_t* bump = nullptr;
    if (peers) {
        HamPeerArgs pa;
        for (int r = 0; r < FMHR_MAX_PEERS; r++) {
            pa.packed[r] = r < peers->world ? (const float4*)peers->packed[r] : nullptr;
            pa.signal[r] = r < peers->world ? peers->flags[r] + peers->rank : nullptr;
        }
        for (int r = 0; r < FMHR_MAX_PEERS; r++) {
            pa.reduced_all[r] = r < peers->world ? (float4*)peers->reduced[r] : nullptr;
            pa.signal_b[r] = r < peers->world ? peers->flags[r] + FMHR_MAX_PEERS + peers->rank : nullptr;
        }
        pa.flags = peers->flags[peers->rank];
        pa.flags_b = peers->flags[peers->rank] + FMHR_MAX_PEERS;
        pa.epoch = peers->epoch;
        pa.reduced = (float4*)peers->reduced[peers->rank];
        pa.world = peers->world;
        pa.rank = peers->rank;
        pa.chunk = cdiv(V, peers->world);
        pa.timeout_cycles = (long long)(peers->timeout_s > 0 ? peers->timeout_s : kPeerDefaultTimeoutS) * kPeerCyclesPerSecond;
        static const int post_mode = [] { const char* e = getenv("FMHR_PEER_POST"); return e ? atoi(e) : FMHR_PEER_POST_DEFAULT; }();
        pa.post_mode = post_mode;
        const bool two_shot = peers->mode == 2 || (peers->mode == 0 && peers->world > 2);
        if (two_shot) {
            // one-shot pulls world x 48 B per vertex into every rank; beyond two ranks the reduce-scatter form moves
            // 1/world of that in and the same out (8 ranks: 16.6 MB -> 2 x 2.1 MB per rank and step)
            ham_peer_reduce_scatter_kernel<<<cdiv(3 * pa.chunk + 1, 256), 256, 0, st>>>(V, pa, ws.status);
            FMHR_LAUNCH_CHECK();
            ham_peer_normal_grad_kernel<<<cdiv(V, 256), 256, 0, st>>>(V, ws.vg, ws.vattr, ws.raw4, pa, ws.status);
        } else {
            ham_peer_reduce_normal_grad_kernel<<<cdiv(V + 1, 256), 256, 0, st>>>(V, ws.vg, ws.vattr, ws.raw4, pa, ws.status);
        }
        packed = peers->reduced[peers->rank];
        bump = peers->epoch;
    } else {
        ham_normal_grad_kernel<<<cdiv(V, 256), 256, 0, st>>>(V, ws.vg, ws.vattr, ws.raw4, (const float4*)buf->packed);
    }
    FMHR_LAUNCH_CHECK();
    ham_update_pass2_kernel<<<cdiv((long long)V * kLPV, 256), 256, 0, st>>>(
        *cfg, ws.vg, buf->delta, buf->albedo, buf->v2f_ptr, (const int2*)buf->v2f_nbr, buf->v2v_ptr, buf->v2v_idx,
        packed, ws.greg, buf->adam_m, buf->adam_v, ws.adam_sc, ws.acc, buf->losses, buf->dbg_grad, ws.status, bump,
        buf->adam_step);
    FMHR_LAUNCH_CHECK();
    FMHR_STAGE_MARK();  // 7: update (regularisers, normal backward, Adam)
    if (cfg->phase == 0) {
        ham_update_sh_kernel<<<cdiv(cfg->n_sh_rows * 9, 128), 128, 0, st>>>(*cfg, ws.gsh, packed, buf->sh_coeffs,
                                                                            buf->adam_m, buf->adam_v, ws.adam_sc,
                                                                            buf->dbg_grad_sh, buf->adam_step);
        FMHR_LAUNCH_CHECK();
    }
    return FMHR_OK;
}

// Extra loss terms (the NCC term of ncc_loop.cu): a fully weighted gradient w.r.t. delta computed outside the fused passes is
// added to the regulariser-gradient slot of the workspace, between fmhr_ham_step_render and fmhr_ham_step_update.
__global__ void __launch_bounds__(256) ham_add_delta_grad_kernel(int V, const float* __restrict__ grad, float4* __restrict__ greg) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    float4 g = greg[2 * (size_t)i];
    g.x += grad[3 * (size_t)i]; g.y += grad[3 * (size_t)i + 1]; g.z += grad[3 * (size_t)i + 2];
    greg[2 * (size_t)i] = g;
}
extern "C" int fmhr_ham_add_delta_grad(const fmhr_ham_config* cfg, const fmhr_ham_buffers* buf, const float* grad_delta,
                                       fmhr_stream_t stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    rc = ham_check_buffers(cfg, buf);
    if (rc) return rc;
    FMHR_CHECK_ARG(grad_delta && cfg->phase == 1);
    HamWs ws;
    ham_layout(cfg, (char*)buf->workspace, &ws);
    ham_add_delta_grad_kernel<<<cdiv(cfg->V, 256), 256, 0, (cudaStream_t)stream>>>(cfg->V, grad_delta, ws.greg);
    FMHR_LAUNCH_CHECK();
    return FMHR_OK;
}

extern "C" int fmhr_ham_step_update(const fmhr_ham_config* cfg, const fmhr_ham_buffers* buf, fmhr_stream_t stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    rc = ham_check_buffers(cfg, buf);
    if (rc) return rc;
    return ham_update_impl(cfg, buf, nullptr, (cudaStream_t)stream);
}

extern "C" int fmhr_ham_step_update_peer(const fmhr_ham_config* cfg, const fmhr_ham_buffers* buf,
                                         const fmhr_ham_peers* peers, fmhr_stream_t stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    rc = ham_check_buffers(cfg, buf);
    if (rc) return rc;
    FMHR_CHECK_ARG(peers && peers->world >= 1 && peers->world <= FMHR_MAX_PEERS && peers->rank >= 0 &&
                   peers->rank < peers->world && peers->epoch && peers->mode >= 0 && peers->mode <= 2);
    for (int r = 0; r < peers->world; r++)
        FMHR_CHECK_ARG(peers->packed[r] && peers->flags[r] && peers->reduced[r] &&
                       ((uintptr_t)peers->packed[r] & 15) == 0 && ((uintptr_t)peers->reduced[r] & 15) == 0);
    FMHR_CHECK_ARG(peers->packed[peers->rank] == buf->packed);  // the render pass accumulated into the shared buffer
    return ham_update_impl(cfg, buf, peers, (cudaStream_t)stream);
}

// cudaIpc plumbing of the peer exchange: one allocation per rank (both packed buffers + the flag words), exported as a
// 64-byte handle the host exchanges out of band (torch.distributed) and the peers map into their address space.
extern "C" int fmhr_peer_alloc(size_t bytes, void** dev_ptr, void* handle64) {
    FMHR_CHECK_ARG(bytes > 0 && dev_ptr && handle64);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size is part of the ABI");
    void* p = nullptr;
    FMHR_CUDA(cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        set_error("fmhr_peer_alloc: %s", cudaGetErrorString(e));
        return FMHR_ECUDA;
    }
    memcpy(handle64, &h, 64);
    *dev_ptr = p;
    return FMHR_OK;
}

extern "C" int fmhr_peer_open(const void* handle64, void** dev_ptr) {
    FMHR_CHECK_ARG(handle64 && dev_ptr);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    FMHR_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return FMHR_OK;
}

extern "C" int fmhr_peer_close(void* dev_ptr) {
    FMHR_CHECK_ARG(dev_ptr);
    FMHR_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return FMHR_OK;
}

extern "C" int fmhr_peer_free(void* dev_ptr) {
    FMHR_CHECK_ARG(dev_ptr);
    FMHR_CUDA(cudaFree(dev_ptr));
    return FMHR_OK;
}

extern "C" int fmhr_ham_stage_times(const fmhr_ham_config* cfg, const fmhr_ham_buffers* buf, float* ms_host,
                                    int* n_stages_host, fmhr_stream_t stream) {
    FMHR_CHECK_ARG(ms_host && n_stages_host);
    cudaStream_t st = (cudaStream_t)stream;
    StageTimer tm;
    tm.st = st;
    for (int i = 0; i <= kMaxStages; i++) FMHR_CUDA(cudaEventCreate(&tm.ev[i]));
    tm.mark();
    g_timer = &tm;
    int rc = fmhr_ham_step_render(cfg, buf, stream);
    if (rc == FMHR_OK) rc = fmhr_ham_step_update(cfg, buf, stream);
    g_timer = nullptr;
    if (rc == FMHR_OK) {
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { set_error("fmhr_ham_stage_times: %s", cudaGetErrorString(e)); rc = FMHR_ECUDA; }
    }
    if (rc == FMHR_OK) {
        *n_stages_host = tm.n - 1;
        for (int i = 0; i + 1 < tm.n; i++) cudaEventElapsedTime(&ms_host[i], tm.ev[i], tm.ev[i + 1]);
    }
    for (int i = 0; i <= kMaxStages; i++) cudaEventDestroy(tm.ev[i]);
    return rc;
}

#ifdef FMHR_TRACE
// steady-state tracing: a one-thread kernel between two steps moves the stamps of the step before it into ring entry `rep`
// and re-arms the slots, so a free-running loop of steps (no host synchronisation in between) can be read back afterwards
constexpr int kTraceRing = 64;
static __device__ unsigned long long g_trace_ring[kTraceRing][kTraceSlots][2];
__global__ void ham_trace_mark_kernel(int rep) {
    for (int i = threadIdx.x; i < kTraceSlots; i += blockDim.x) {
        if (rep >= 0 && rep < kTraceRing) { g_trace_ring[rep][i][0] = g_trace[i][0]; g_trace_ring[rep][i][1] = g_trace[i][1]; }
        g_trace[i][0] = ~0ull; g_trace[i][1] = 0ull;
    }
}
#endif
// Diagnostic builds: enqueue "store the stamps collected since the previous mark as ring entry `rep` (< 64; negative: discard)
// and re-arm" on `stream`; fmhr_trace_read with n_slots = -reps then returns reps x 64 x 2 stamps of the ring.
extern "C" int fmhr_trace_mark(int rep, fmhr_stream_t stream) {
#ifdef FMHR_TRACE
    ham_trace_mark_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(rep);
    FMHR_LAUNCH_CHECK();
    return FMHR_OK;
#else
    (void)rep; (void)stream;
    set_error("fmhr_trace_mark: this build has no timeline tracing (compile with -DFMHR_TRACE)");
    return FMHR_EUNSUPPORTED;
#endif
}

// Diagnostic builds (-DFMHR_TRACE): first-entry / last-exit %globaltimer stamps (ns) of every kernel slot since the last
// reset; the product build returns FMHR_EUNSUPPORTED.  Synchronises the device.
extern "C" int fmhr_trace_read(unsigned long long* stamps_host, int n_slots, int reset) {
#ifdef FMHR_TRACE
    FMHR_CHECK_ARG(n_slots >= -kTraceRing && n_slots <= kTraceSlots && (stamps_host || n_slots == 0));
    FMHR_CUDA(cudaDeviceSynchronize());
    if (n_slots < 0) {  // the ring of fmhr_trace_mark
        FMHR_CUDA(cudaMemcpyFromSymbol(stamps_host, g_trace_ring, (size_t)(-n_slots) * kTraceSlots * 2 * sizeof(unsigned long long)));
        return FMHR_OK;
    }
    if (n_slots > 0)
        FMHR_CUDA(cudaMemcpyFromSymbol(stamps_host, g_trace, (size_t)n_slots * 2 * sizeof(unsigned long long)));
    if (reset) {
        unsigned long long init[kTraceSlots][2];
        for (int i = 0; i < kTraceSlots; i++) { init[i][0] = ~0ull; init[i][1] = 0ull; }
        FMHR_CUDA(cudaMemcpyToSymbol(g_trace, init, sizeof(init)));
    }
    return FMHR_OK;
#else
    (void)stamps_host; (void)n_slots; (void)reset;
    set_error("fmhr_trace_read: this build has no timeline tracing (compile with -DFMHR_TRACE)");
    return FMHR_EUNSUPPORTED;
#endif
}

extern "C" int fmhr_ham_debug_export(const fmhr_ham_config* cfg, const fmhr_ham_buffers* buf, float* pos, float* rast,
                                     float* image, float* pred_mask, float* normals, fmhr_stream_t stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    rc = ham_check_buffers(cfg, buf);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    // re-run the forward with export pointers (leaves `packed` holding forward-only partials); the antialias pass only
    // visits covered pixels and their ring, every other pixel of the exported planes is zero
    {
        const size_t P = (size_t)cfg->n_views * cfg->H * cfg->W;
        if (image) FMHR_CUDA(cudaMemsetAsync(image, 0, P * 3 * sizeof(float), st));
        if (pred_mask) FMHR_CUDA(cudaMemsetAsync(pred_mask, 0, P * sizeof(float), st));
    }
    rc = cfg->phase == 0 ? ham_render_impl<0>(cfg, buf, st, image, pred_mask, true)
                         : ham_render_impl<1>(cfg, buf, st, image, pred_mask, true);
    if (rc) return rc;
    HamWs ws;
    ham_layout(cfg, (char*)buf->workspace, &ws);
    if (pos) {
        ham_export_pos_kernel<<<dim3(cdiv(cfg->V, 256), cfg->n_views), 256, 0, st>>>(ws.vg, ws.viewM, cfg->V, (float4*)pos);
        FMHR_LAUNCH_CHECK();
    }
    if (normals) FMHR_CUDA(cudaMemcpy2DAsync(normals, 12, ws.vattr, 32, 12, (size_t)cfg->V, cudaMemcpyDeviceToDevice, st));
    if (rast) {
        const dim3 pgrid(cdiv((long long)cfg->H * cfg->W, 256), cfg->n_views);
        ham_export_rast_kernel<<<pgrid, 256, 0, st>>>(ws.zbuf[cfg->zbuf_slot], ws.vg, ws.viewM, buf->tri, cfg->V, cfg->H, cfg->W,
                                                      (float4*)rast);
        FMHR_LAUNCH_CHECK();
    }
    return FMHR_OK;
}

extern "C" int fmhr_ham_step_host(const fmhr_ham_config* cfg, const fmhr_ham_buffers* buf, const float* imgs_host,
                                  const float* masks_host, const float* valid_masks_host, const float* w2cs_host,
                                  const float* projs_host, float* losses_host, fmhr_stream_t stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    rc = ham_check_buffers(cfg, buf);
    if (rc) return rc;
    FMHR_CHECK_ARG(imgs_host && masks_host && valid_masks_host && w2cs_host && projs_host && losses_host);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t hw = (size_t)cfg->H * cfg->W, n = cfg->n_views;
    // stage this step's batch into rows [0, n_views) of the device view arrays (view_idx must be 0..n_views-1)
    FMHR_CUDA(cudaMemcpyAsync((void*)buf->imgs, imgs_host, n * hw * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    FMHR_CUDA(cudaMemcpyAsync((void*)buf->masks, masks_host, n * hw * sizeof(float), cudaMemcpyHostToDevice, st));
    FMHR_CUDA(cudaMemcpyAsync((void*)buf->valid_masks, valid_masks_host, n * hw * sizeof(float), cudaMemcpyHostToDevice, st));
    FMHR_CUDA(cudaMemcpyAsync((void*)buf->w2cs, w2cs_host, n * 16 * sizeof(float), cudaMemcpyHostToDevice, st));
    FMHR_CUDA(cudaMemcpyAsync((void*)buf->projs, projs_host, n * 16 * sizeof(float), cudaMemcpyHostToDevice, st));
    rc = fmhr_ham_prepare_views(buf->valid_masks, (int)n, cfg->H, cfg->W, (double*)buf->view_vm2, stream);
    if (rc) return rc;
    rc = fmhr_ham_step_render(cfg, buf, stream);
    if (rc) return rc;
    rc = fmhr_ham_step_update(cfg, buf, stream);
    if (rc) return rc;
    FMHR_CUDA(cudaMemcpyAsync(losses_host, buf->losses, 8 * sizeof(float), cudaMemcpyDeviceToHost, st));
    return FMHR_OK;
}

extern "C" size_t fmhr_ham_host_u8_staging_bytes(const fmhr_ham_config* cfg) {
    if (!cfg || check_cfg(cfg) != FMHR_OK) return 0;
    return (size_t)cfg->n_views * cfg->H * cfg->W * 4;
}

extern "C" int fmhr_ham_step_host_u8(const fmhr_ham_config* cfg, const fmhr_ham_buffers* buf, const uint8_t* imgs_host,
                                     const uint8_t* masks_host, const float* w2cs_host, const float* projs_host,
                                     void* staging, float* losses_host, fmhr_stream_t stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    rc = ham_check_buffers(cfg, buf);
    if (rc) return rc;
    FMHR_CHECK_ARG(imgs_host && masks_host && w2cs_host && projs_host && staging && losses_host);
    const size_t P = (size_t)cfg->n_views * cfg->H * cfg->W;
    FMHR_CHECK_ARG(P % 4 == 0 && ((uintptr_t)staging & 15) == 0 && ((uintptr_t)buf->imgs & 15) == 0 &&
                   ((uintptr_t)buf->masks & 15) == 0);
    cudaStream_t st = (cudaStream_t)stream;
    SideStream* side = nullptr;
    rc = side_stream(&side);
    if (rc) return rc;
    // image / mask bytes travel on the copy stream while the geometry half of the iteration runs on `st`
    FMHR_CUDA(cudaEventRecord(side->copy_fork, st));  // earlier work on `st` may still read the staging buffer
    FMHR_CUDA(cudaStreamWaitEvent(side->copy, side->copy_fork, 0));
    FMHR_CUDA(cudaMemcpyAsync(staging, imgs_host, P * 3, cudaMemcpyHostToDevice, side->copy));
    FMHR_CUDA(cudaMemcpyAsync((char*)staging + P * 3, masks_host, P, cudaMemcpyHostToDevice, side->copy));
    FMHR_CUDA(cudaEventRecord(side->ready, side->copy));
    const size_t n = cfg->n_views;
    FMHR_CUDA(cudaMemcpyAsync((void*)buf->w2cs, w2cs_host, n * 16 * sizeof(float), cudaMemcpyHostToDevice, st));
    FMHR_CUDA(cudaMemcpyAsync((void*)buf->projs, projs_host, n * 16 * sizeof(float), cudaMemcpyHostToDevice, st));
    g_pending.staging = (const uint8_t*)staging;
    g_pending.ready = side->ready;
    g_pending.consumed = nullptr;
    rc = fmhr_ham_step_render(cfg, buf, stream);
    g_pending.staging = nullptr;
    if (rc) return rc;
    rc = fmhr_ham_step_update(cfg, buf, stream);
    if (rc) return rc;
    FMHR_CUDA(cudaMemcpyAsync(losses_host, buf->losses, 8 * sizeof(float), cudaMemcpyDeviceToHost, st));
    return FMHR_OK;
}


// Pipelined form of the host-batch step: the NEXT step's batch is uploaded while the current step computes, so a
// PCIe-bound loop runs at the transfer rate instead of transfer + the post-upload half of the iteration.
// Picks the staging slot of a submission and orders the copy stream behind the conversion of the slot's previous batch.
static int submit_slot(SideStream* side, void* staging, const char* who, int* slot_out) {
    int slot = side->slot_ptr[0] == staging ? 0 : (side->slot_ptr[1] == staging ? 1 : -1);
    if (slot < 0) slot = side->slot_ptr[0] == nullptr ? 0 : (side->slot_ptr[1] == nullptr ? 1 : -1);
    if (slot < 0) slot = side->slot_used[0] ? 0 : (side->slot_used[1] ? 1 : -1);  // a consumed buffer gives way to a new one
    if (slot < 0) {
        set_error("%s: two submitted batches are already waiting for their steps on this device", who);
        return FMHR_EINVAL;
    }
    if (side->slot_ptr[slot] == staging && side->slot_used[slot])  // the conversion of its previous batch has been issued
        FMHR_CUDA(cudaStreamWaitEvent(side->copy, side->slot_free[slot], 0));
    side->slot_ptr[slot] = staging;
    side->slot_used[slot] = false;
    *slot_out = slot;
    return FMHR_OK;
}

extern "C" int fmhr_ham_host_u8_submit(const fmhr_ham_config* cfg, const uint8_t* imgs_host, const uint8_t* masks_host,
                                       void* staging) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    FMHR_CHECK_ARG(imgs_host && masks_host && staging && ((uintptr_t)staging & 15) == 0);
    const size_t P = (size_t)cfg->n_views * cfg->H * cfg->W;
    FMHR_CHECK_ARG(P % 4 == 0);
    SideStream* side = nullptr;
    rc = side_stream(&side);
    if (rc) return rc;
    int slot = -1;
    rc = submit_slot(side, staging, "fmhr_ham_host_u8_submit", &slot);
    if (rc) return rc;
    FMHR_CUDA(cudaMemcpyAsync(staging, imgs_host, P * 3, cudaMemcpyHostToDevice, side->copy));
    FMHR_CUDA(cudaMemcpyAsync((char*)staging + P * 3, masks_host, P, cudaMemcpyHostToDevice, side->copy));
    FMHR_CUDA(cudaEventRecord(side->slot_ready[slot], side->copy));
    return FMHR_OK;
}

// Same, but only the rectangle of every view that can matter travels: boxes_host[v] = (y0, y1, x0, x1), half-open, must
// contain every pixel of view v whose mask byte is > 127 (the loader knows it: the bounding box of the segmentation).
// Outside it mask = 0, so no pixel there is "valid" (mesh_sfs_optim.py:276) and its image bytes are never read; the mask
// staging plane is zero-filled on the device and the image / mask rectangles are copied row by row (cudaMemcpy2DAsync).
extern "C" int fmhr_ham_host_u8_submit_boxes(const fmhr_ham_config* cfg, const uint8_t* imgs_host,
                                             const uint8_t* masks_host, const int32_t* boxes_host, void* staging,
                                             size_t* h2d_bytes) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    FMHR_CHECK_ARG(imgs_host && masks_host && boxes_host && staging && ((uintptr_t)staging & 15) == 0);
    const int n = cfg->n_views, H = cfg->H, W = cfg->W;
    const size_t P = (size_t)n * H * W;
    FMHR_CHECK_ARG(P % 4 == 0);
    for (int v = 0; v < n; v++) {
        const int32_t* b = boxes_host + 4 * v;
        FMHR_CHECK_ARG(b[0] >= 0 && b[1] <= H && b[2] >= 0 && b[3] <= W);
    }
    SideStream* side = nullptr;
    rc = side_stream(&side);
    if (rc) return rc;
    int slot = -1;
    rc = submit_slot(side, staging, "fmhr_ham_host_u8_submit_boxes", &slot);
    if (rc) return rc;
    uint8_t* st_img = (uint8_t*)staging;
    uint8_t* st_msk = st_img + P * 3;
    size_t bytes = 0;
    for (int v = 0; v < n; v++) {
        const int32_t* b = boxes_host + 4 * v;
        if (b[1] > b[0] && b[3] > b[2]) bytes += (size_t)(b[1] - b[0]) * (size_t)(b[3] - b[2]) * 4;
    }
    // Preferred form: ONE kernel pulls every box row out of the mapped host buffers (needs device-visible, 16-byte aligned
    // pinned memory and planes whose byte counts are multiples of 16); otherwise the rectangles go as 2-D DMA copies.
    void *d_imgs = nullptr, *d_masks = nullptr;
    const char* force_dma = getenv("FMHR_BOX_DMA");  // test hook: exercise the 2-D copy form
    const bool can_pull = !(force_dma && force_dma[0] == '1') && (P % 16 == 0) &&
                          (((uintptr_t)imgs_host | (uintptr_t)masks_host) & 15) == 0 &&
                          cudaHostGetDevicePointer(&d_imgs, (void*)imgs_host, 0) == cudaSuccess &&
                          cudaHostGetDevicePointer(&d_masks, (void*)masks_host, 0) == cudaSuccess;
    if (!can_pull) cudaGetLastError();
    if (can_pull) {
        if (side->box_cap[slot] < n) {
            if (side->box_dev[slot]) FMHR_CUDA(cudaFree(side->box_dev[slot]));
            side->box_dev[slot] = nullptr;
            FMHR_CUDA(cudaMalloc(&side->box_dev[slot], (size_t)n * sizeof(int4)));
            side->box_cap[slot] = n;
        }
        FMHR_CUDA(cudaMemcpyAsync(side->box_dev[slot], boxes_host, (size_t)n * sizeof(int4), cudaMemcpyHostToDevice, side->copy));
        FMHR_CUDA(cudaMemsetAsync(st_msk, 0, P, side->copy));
        ham_pull_boxes_kernel<<<dim3(8, n), 256, 0, side->copy>>>((const uint8_t*)d_imgs, (const uint8_t*)d_masks,
                                                                  side->box_dev[slot], H, W, st_img, st_msk);
        FMHR_LAUNCH_CHECK();
        FMHR_CUDA(cudaEventRecord(side->slot_ready[slot], side->copy));
        if (h2d_bytes) *h2d_bytes = bytes;
        return FMHR_OK;
    }
    // A loader cycles through a few pinned buffers, so the same (host batch, boxes, staging) triple comes back: the second
    // time it is seen its 1 + 2 n copy operations are captured into a CUDA graph, afterwards one graph launch replaces
    // ~100 driver calls per step (the submission was host-bound: 0.3 ms of API time for 0.15 ms of transfer).
    uint64_t hash = 1469598103934665603ull;
    for (int i = 0; i < 4 * n; i++) hash = (hash ^ (uint32_t)boxes_host[i]) * 1099511628211ull;
    BoxGraph& bg = side->box_graph[slot];
    const bool match = bg.imgs == imgs_host && bg.masks == masks_host && bg.staging == staging && bg.n == n && bg.H == H &&
                       bg.W == W && bg.hash == hash;
    if (match && bg.exec) {
        FMHR_CUDA(cudaGraphLaunch(bg.exec, side->copy));
    } else {
        if (!match) {
            if (bg.exec) cudaGraphExecDestroy(bg.exec);
            bg = BoxGraph();
            bg.imgs = imgs_host; bg.masks = masks_host; bg.staging = staging; bg.n = n; bg.H = H; bg.W = W; bg.hash = hash;
        }
        const bool capture = match && bg.seen >= 1;
        if (capture) FMHR_CUDA(cudaStreamBeginCapture(side->copy, cudaStreamCaptureModeThreadLocal));
        cudaError_t err = cudaMemsetAsync(st_msk, 0, P, side->copy);
        for (int v = 0; v < n && err == cudaSuccess; v++) {
            const int y0 = boxes_host[4 * v], y1 = boxes_host[4 * v + 1], x0 = boxes_host[4 * v + 2], x1 = boxes_host[4 * v + 3];
            if (y1 <= y0 || x1 <= x0) continue;  // nothing segmented in this view
            const size_t off = ((size_t)v * H + y0) * W + x0;
            const size_t rows = (size_t)(y1 - y0), cols = (size_t)(x1 - x0);
            err = cudaMemcpy2DAsync(st_img + off * 3, (size_t)W * 3, imgs_host + off * 3, (size_t)W * 3, cols * 3, rows,
                                    cudaMemcpyHostToDevice, side->copy);
            if (err == cudaSuccess)
                err = cudaMemcpy2DAsync(st_msk + off, (size_t)W, masks_host + off, (size_t)W, cols, rows,
                                        cudaMemcpyHostToDevice, side->copy);
        }
        if (capture) {
            cudaGraph_t graph = nullptr;
            const cudaError_t e2 = cudaStreamEndCapture(side->copy, &graph);  // always ends the capture, also after an error
            if (err == cudaSuccess) err = e2;
            if (err == cudaSuccess) err = cudaGraphInstantiate(&bg.exec, graph, 0);
            if (graph) cudaGraphDestroy(graph);
            if (err == cudaSuccess) err = cudaGraphLaunch(bg.exec, side->copy);
        }
        if (err != cudaSuccess) {
            set_error("fmhr_ham_host_u8_submit_boxes: %s", cudaGetErrorString(err));
            return FMHR_ECUDA;
        }
        bg.seen++;
    }
    FMHR_CUDA(cudaEventRecord(side->slot_ready[slot], side->copy));
    if (h2d_bytes) *h2d_bytes = bytes;
    return FMHR_OK;
}

extern "C" int fmhr_ham_host_u8_submit_boxes_direct(const fmhr_ham_config* cfg, const uint8_t* imgs_host,
                                                    const uint8_t* masks_host, const int32_t* boxes_host,
                                                    const float* w2cs_host, const float* projs_host, float* imgs_dev,
                                                    float* masks_dev, float* w2cs_dev, float* projs_dev,
                                                    int fresh, size_t* h2d_bytes) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    FMHR_CHECK_ARG(imgs_host && masks_host && boxes_host && w2cs_host && projs_host && imgs_dev && masks_dev && w2cs_dev &&
                   projs_dev);
    FMHR_CHECK_ARG((((uintptr_t)imgs_dev | (uintptr_t)masks_dev) & 15) == 0);
    const int n = cfg->n_views, H = cfg->H, W = cfg->W;
    const size_t P = (size_t)n * H * W;
    for (int v = 0; v < n; v++) {
        const int32_t* b = boxes_host + 4 * v;
        FMHR_CHECK_ARG(b[0] >= 0 && b[1] <= H && b[2] >= 0 && b[3] <= W);
    }
    void *d_imgs = nullptr, *d_masks = nullptr, *d_w2cs = nullptr, *d_projs = nullptr;
    // (H * W % 16 == 0: a 16-element chunk of the flat planes never straddles two views)
    const bool can_pull = (((size_t)H * W) % 16 == 0) && (((uintptr_t)imgs_host | (uintptr_t)masks_host) & 15) == 0 &&
                          cudaHostGetDevicePointer(&d_imgs, (void*)imgs_host, 0) == cudaSuccess &&
                          cudaHostGetDevicePointer(&d_masks, (void*)masks_host, 0) == cudaSuccess &&
                          cudaHostGetDevicePointer(&d_w2cs, (void*)w2cs_host, 0) == cudaSuccess &&
                          cudaHostGetDevicePointer(&d_projs, (void*)projs_host, 0) == cudaSuccess;
    if (!can_pull) {
        cudaGetLastError();
        set_error("fmhr_ham_host_u8_submit_boxes_direct: needs device-mapped, 16-byte aligned pinned host buffers and "
                  "H*W %% 16 == 0; use fmhr_ham_host_u8_submit_boxes");
        return FMHR_EUNSUPPORTED;
    }
    SideStream* side = nullptr;
    rc = side_stream(&side);
    if (rc) return rc;
    // which of the two plane sets (the slot protocol's bookkeeping is done by submit_slot below; here only the index)
    int slot = side->slot_ptr[0] == (void*)imgs_dev ? 0 : (side->slot_ptr[1] == (void*)imgs_dev ? 1 : -1);
    if (slot < 0) slot = side->slot_ptr[0] == nullptr ? 0 : (side->slot_ptr[1] == nullptr ? 1 : -1);
    if (slot < 0) slot = side->slot_used[0] ? 0 : (side->slot_used[1] ? 1 : -1);
    if (slot < 0) {
        set_error("fmhr_ham_host_u8_submit_boxes_direct: two submitted batches are already waiting for their steps on this device");
        return FMHR_EINVAL;
    }
    size_t bytes = (size_t)n * 32 * sizeof(float);
    for (int v = 0; v < n; v++) {
        const int32_t* b = boxes_host + 4 * v;
        if (b[1] > b[0] && b[3] > b[2]) bytes += (size_t)(b[1] - b[0]) * (size_t)(b[3] - b[2]) * 4;
    }
    if (side->dbox_cap[slot] < n) {
        for (int k = 0; k < 2; k++) {
            if (side->dbox[slot][k]) FMHR_CUDA(cudaFree(side->dbox[slot][k]));
            side->dbox[slot][k] = nullptr;
            FMHR_CUDA(cudaMalloc(&side->dbox[slot][k], (size_t)n * sizeof(int4)));
        }
        side->dbox_cap[slot] = n;
        side->dprimed[slot] = nullptr;
    }
    const int cur = side->dbox_cur[slot] ^ 1, prv = cur ^ 1;
    side->dbox_cur[slot] = cur;
    const bool prime = fresh || side->dprimed[slot] != (const void*)masks_dev || side->dprimed_n[slot] != n;
    // The box table goes up BEFORE the copy stream starts waiting for the plane set: table `cur` was last read by the pull
    // two batches ago (finished, same stream), and when the set is released the pull kernel is then the first thing to
    // launch - ahead of the next step's coverage kernel, whose blocks would otherwise fill every SM first.
    if (prime) FMHR_CUDA(cudaMemsetAsync(side->dbox[slot][prv], 0, (size_t)n * sizeof(int4), side->copy));
    FMHR_CUDA(cudaMemcpyAsync(side->dbox[slot][cur], boxes_host, (size_t)n * sizeof(int4), cudaMemcpyHostToDevice, side->copy));
    int slot2 = -1;
    rc = submit_slot(side, imgs_dev, "fmhr_ham_host_u8_submit_boxes_direct", &slot2);
    if (rc) return rc;
    FMHR_CHECK_ARG(slot2 == slot);
    if (prime) {
        // first batch into this plane set (or the caller says its content is undefined): establish "zero outside the
        // previous boxes"
        FMHR_CUDA(cudaMemsetAsync(masks_dev, 0, P * sizeof(float), side->copy));
        side->dprimed[slot] = masks_dev;
        side->dprimed_n[slot] = n;
    }
    static const int env_blocks = [] { const char* e = getenv("FMHR_PULL_BLOCKS"); return e ? atoi(e) : 0; }();
    const int pull_blocks = max(1, min(n, env_blocks > 0 ? env_blocks : kPullBlocks));
    ham_pull_boxes_f32_kernel<<<pull_blocks, kPullThreads, 0, side->copy>>>(
        (const uint8_t*)d_imgs, (const uint8_t*)d_masks, side->dbox[slot][cur], side->dbox[slot][prv], H, W, imgs_dev, masks_dev,
        (const float*)d_w2cs, (const float*)d_projs, w2cs_dev, projs_dev, n * 16, n);
    FMHR_LAUNCH_CHECK();
    FMHR_CUDA(cudaEventRecord(side->slot_ready[slot], side->copy));
    if (h2d_bytes) *h2d_bytes = bytes;
    return FMHR_OK;
}

// The submitted-batch step in three parts, so that a host can capture the device work (body) into a CUDA graph while the
// cross-stream handshakes with the copy stream (acquire / release) stay ordinary stream operations around the replay.
static int submitted_slot(SideStream* side, const void* staging, const char* who) {
    const int slot = side->slot_ptr[0] == staging ? 0 : (side->slot_ptr[1] == staging ? 1 : -1);
    if (slot < 0 || side->slot_used[slot]) {
        set_error("%s: no batch was submitted into this staging buffer", who);
        return -1;
    }
    return slot;
}

extern "C" int fmhr_ham_step_host_u8_acquire(void* staging, fmhr_stream_t stream) {
    FMHR_CHECK_ARG(staging);
    SideStream* side = nullptr;
    int rc = side_stream(&side);
    if (rc) return rc;
    const int slot = submitted_slot(side, staging, "fmhr_ham_step_host_u8_acquire");
    if (slot < 0) return FMHR_EINVAL;
    FMHR_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, side->slot_ready[slot], 0));
    return FMHR_OK;
}

extern "C" int fmhr_ham_step_host_u8_body(const fmhr_ham_config* cfg, const fmhr_ham_buffers* buf, const float* w2cs_host,
                                          const float* projs_host, const void* staging, float* losses_host,
                                          const fmhr_ham_peers* peers, fmhr_stream_t stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    rc = ham_check_buffers(cfg, buf);
    if (rc) return rc;
    FMHR_CHECK_ARG(staging && losses_host);
    // converted-on-arrival batches (fmhr_ham_host_u8_submit_boxes_direct): the planes and cameras `buf` names ARE the batch
    const bool direct = staging == (const void*)buf->imgs;
    FMHR_CHECK_ARG(direct ? (!w2cs_host && !projs_host) : (w2cs_host && projs_host));
    FMHR_CHECK_ARG(((uintptr_t)buf->imgs & 15) == 0 && ((uintptr_t)buf->masks & 15) == 0);
    FMHR_CHECK_ARG(((size_t)cfg->n_views * cfg->H * cfg->W) % 4 == 0);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = cfg->n_views;
    if (!direct) {
        FMHR_CUDA(cudaMemcpyAsync((void*)buf->w2cs, w2cs_host, n * 16 * sizeof(float), cudaMemcpyHostToDevice, st));
        FMHR_CUDA(cudaMemcpyAsync((void*)buf->projs, projs_host, n * 16 * sizeof(float), cudaMemcpyHostToDevice, st));
    }
    g_pending.staging = direct ? nullptr : (const uint8_t*)staging;
    g_pending.ready = nullptr;     // acquire ordered the stream behind the upload
    g_pending.consumed = nullptr;  // release frees the buffer
    rc = fmhr_ham_step_render(cfg, buf, stream);
    g_pending.staging = nullptr;
    if (rc) return rc;
    rc = peers ? fmhr_ham_step_update_peer(cfg, buf, peers, stream) : fmhr_ham_step_update(cfg, buf, stream);
    if (rc) return rc;
    FMHR_CUDA(cudaMemcpyAsync(losses_host, buf->losses, 8 * sizeof(float), cudaMemcpyDeviceToHost, st));
    return FMHR_OK;
}

extern "C" int fmhr_ham_step_host_u8_release(void* staging, fmhr_stream_t stream) {
    FMHR_CHECK_ARG(staging);
    SideStream* side = nullptr;
    int rc = side_stream(&side);
    if (rc) return rc;
    const int slot = submitted_slot(side, staging, "fmhr_ham_step_host_u8_release");
    if (slot < 0) return FMHR_EINVAL;
    FMHR_CUDA(cudaEventRecord(side->slot_free[slot], (cudaStream_t)stream));
    side->slot_used[slot] = true;
    return FMHR_OK;
}

extern "C" int fmhr_ham_step_host_u8_submitted(const fmhr_ham_config* cfg, const fmhr_ham_buffers* buf,
                                               const float* w2cs_host, const float* projs_host, void* staging,
                                               float* losses_host, const fmhr_ham_peers* peers, fmhr_stream_t stream) {
    int rc = fmhr_ham_step_host_u8_acquire(staging, stream);
    if (rc) return rc;
    rc = fmhr_ham_step_host_u8_body(cfg, buf, w2cs_host, projs_host, staging, losses_host, peers, stream);
    if (rc) return rc;
    return fmhr_ham_step_host_u8_release(staging, stream);
}

extern "C" size_t fmhr_ham_init_scratch_bytes(int num) { return num > 0 ? (size_t)(num + 1) * 56 * sizeof(double) : 0; }

extern "C" int fmhr_ham_init(const fmhr_ham_config* cfg, const fmhr_ham_buffers* buf, const float* grayimgs,
                             float* valid_masks_out, float* sh_coeffs_out, float* sh_global_out, float* albedo_mean_out,
                             void* scratch, fmhr_stream_t stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    FMHR_CHECK_ARG(cfg->phase == 0);  // the phase-A workspace layout (normals / albedo planes) is the one used here
    FMHR_CHECK_ARG(buf && grayimgs && valid_masks_out && sh_coeffs_out && sh_global_out && albedo_mean_out && scratch);
    // valid_masks / view_vm2 are OUTPUTS of the initialisation: the checks below only need them non-null
    fmhr_ham_buffers b = *buf;
    if (!b.valid_masks) b.valid_masks = valid_masks_out;
    if (!b.view_vm2) b.view_vm2 = (const double*)scratch;
    rc = ham_check_buffers(cfg, &b);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int num = cfg->n_views;  // the reference initialises from ALL views (mesh_sfs_optim.py:130)
    const size_t P = (size_t)num * cfg->H * cfg->W;
    double* init_acc = (double*)scratch;
    FMHR_CUDA(cudaMemsetAsync(init_acc, 0, fmhr_ham_init_scratch_bytes(num), st));
    FMHR_CUDA(cudaMemsetAsync(valid_masks_out, 0, P * sizeof(float), st));
    InitArgs ia{grayimgs, init_acc};
    rc = ham_render_impl<0>(cfg, &b, st, nullptr, valid_masks_out, true, &ia);
    if (rc) return rc;
    HamWs ws;
    ham_layout(cfg, (char*)b.workspace, &ws);
    double* global_row = init_acc + (size_t)num * 56;
    ham_init_solve_kernel<<<num + 1, 32, 0, st>>>(init_acc, num, sh_coeffs_out, sh_global_out, global_row);
    FMHR_LAUNCH_CHECK();
    ham_init_albedo_kernel<<<296, 256, 0, st>>>(ws.clist, ws.ccount, ws.plane[2], b.imgs, b.view_idx, cfg->H, cfg->W,
                                                sh_global_out, global_row);
    FMHR_LAUNCH_CHECK();
    ham_init_albedo_mean_kernel<<<1, 32, 0, st>>>(global_row, albedo_mean_out);
    FMHR_LAUNCH_CHECK();
    return FMHR_OK;
}
