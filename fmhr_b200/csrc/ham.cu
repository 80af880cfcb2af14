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
#include "aa_rule.cuh"

namespace fmhr {

int launch_raster_coverage_snapped(const float4* pos, const int2* snap, const int32_t* tri, int N, int V, int T, int H,
                                   int W, unsigned long long* zbuf, uint32_t* tbits, uint32_t* tlist, int* tcount,
                                   int tiles_x, int tiles_per_view, cudaStream_t st);
int launch_vertex_normals_fwd(const float* verts, const int32_t* tri, const int32_t* v2f_ptr, const int32_t* v2f_idx,
                              const int32_t* v2f_nbr, int V, float* normals, float* raw, cudaStream_t st);

// Per-view constants of the backward pass, written once per step by the transform kernel:
//   viewM[n][0..11] = rows 0..2 of (w2c @ proj), columns x,y,z,w : d(clip_j)/d(world_i) = M[4i+j]
constexpr int kViewM = 12;

struct HamWs {
    unsigned long long* zbuf[2];  // [n,H,W] x2: `zbuf_slot` is rasterised this step, the other is reset for the next
    float4* plane[4];          // [n,H,W] each
    float4* pos;               // [n,V] clip positions
    int2* snap;                // [n,V] 24.8 fixed-point window coordinates (x == INT_MIN: vertex rejected)
    float2* scr;               // [n,V] (x/w*W/2, y/w*H/2): the antialias rule's window coordinates, divide done once
    float* viewM;              // [n,12] d(clip)/d(world) per view
    // Compact work lists of the backward pass (built by the forward passes of the same iteration):
    uint32_t* vlist;           // [P]   pixels that feed the backward shader (phase B: valid, phase A: covered), by shade
    uint4* plist_a;            // [P/2] blending pixel pairs found by the antialias pass: (pixel0, flags, alpha, i1)
    uint32_t* plist_b;         // [P/2]                                                    i2
    float4* gdelta;            // [P]   pair terms of d(loss)/d(pre-antialias value); zero outside an iteration
    int* vcount;               // inside common_region (zeroed every step)
    int* pcount;
    int* status;               // bit 0: pair list overflow
    int* cursors;              // [8] work cursors of the persistent kernels (inside common_region)
    // Active-tile work lists (16x16 tiles).  slot[s]: tiles of z-buffer slot s that received fragments (bitmap for
    // de-duplication + compact list + count, filled by the coverage kernel); act: those tiles dilated by their four
    // edge neighbours (antialias pairs straddle tile edges), filled by the shade pass.  The pixel passes are persistent
    // kernels that walk these lists, so idle tiles cost nothing (launching a block per tile spent ~80 us per pass on
    // ~25k blocks whose only instruction was a flag load).
    char* slot_region[2];        // [count (256 B) | bitmap] zeroed together when the slot is reused
    int* tcount[2];
    uint32_t* tbits[2];          // [n, words_per_view]
    uint32_t* tlist[2];          // [n * tiles_per_view]
    char* common_region;         // [acc | acount | abits] zeroed every step
    size_t common_bytes, slot_bytes;
    int* acount;
    uint32_t* abits;
    uint32_t* alist;
    float* vertices;           // [V,3]
    float* normals;            // [V,3] normalised
    float* raw;                // [V,3] un-normalised normal sums
    float* gN;                 // [V,3]
    float* yhat_v;             // [V,3]
    float* yhat_a;             // [V,3]
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
    const size_t P = (size_t)c->n_views * c->H * c->W, V = (size_t)c->V;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += align256(bytes); return p; };
    char* p;
    for (int i = 0; i < 2; i++) { p = take(P * 8); if (ws) ws->zbuf[i] = (unsigned long long*)p; }
    const int nplanes = c->phase == 0 ? 4 : 2;
    for (int i = 0; i < 4; i++) {
        p = (i < nplanes) ? take(P * 16) : nullptr;
        if (ws) ws->plane[i] = (float4*)p;
    }
    p = take((size_t)c->n_views * V * 16); if (ws) ws->pos = (float4*)p;
    p = take((size_t)c->n_views * V * 8); if (ws) ws->snap = (int2*)p;
    p = take((size_t)c->n_views * V * 8); if (ws) ws->scr = (float2*)p;
    p = take((size_t)c->n_views * kViewM * 4); if (ws) ws->viewM = (float*)p;
    p = take(P * 4); if (ws) ws->vlist = (uint32_t*)p;
    p = take((P / 2 + 64) * 16); if (ws) ws->plist_a = (uint4*)p;
    p = take((P / 2 + 64) * 4); if (ws) ws->plist_b = (uint32_t*)p;
    p = take(P * 16); if (ws) ws->gdelta = (float4*)p;
    const size_t tiles_pv = (size_t)((c->W + 15) / 16) * ((c->H + 15) / 16);
    const size_t words = (size_t)c->n_views * ((tiles_pv + 31) / 32);
    const size_t slot_bytes = 256 + align256(words * 4);
    for (int i = 0; i < 2; i++) {
        p = take(slot_bytes);
        if (ws) { ws->slot_region[i] = p; ws->tcount[i] = (int*)p; ws->tbits[i] = (uint32_t*)(p + 256); }
        p = take((size_t)c->n_views * tiles_pv * 4); if (ws) ws->tlist[i] = (uint32_t*)p;
    }
    p = take((size_t)c->n_views * tiles_pv * 4); if (ws) ws->alist = (uint32_t*)p;
    const size_t common_bytes = 8 * 32 * sizeof(double) + 256 + align256(words * 4);
    p = take(common_bytes);
    if (ws) {
        ws->common_region = p; ws->common_bytes = common_bytes; ws->slot_bytes = slot_bytes;
        ws->acc = (double*)p; ws->acount = (int*)(p + 8 * 32 * sizeof(double));
        ws->cursors = ws->acount + 8;  // same zeroed 256-byte slot
        ws->vcount = ws->acount + 16; ws->pcount = ws->acount + 17; ws->status = ws->acount + 18;
        ws->abits = (uint32_t*)(p + 8 * 32 * sizeof(double) + 256);
    }
    p = take(V * 12); if (ws) ws->vertices = (float*)p;
    p = take(V * 12); if (ws) ws->normals = (float*)p;
    p = take(V * 12); if (ws) ws->raw = (float*)p;
    p = take(V * 12); if (ws) ws->gN = (float*)p;
    p = take(V * 12); if (ws) ws->yhat_v = (float*)p;
    p = take(V * 12); if (ws) ws->yhat_a = (float*)p;
    p = take((size_t)c->n_sh_rows * 9 * 4); if (ws) ws->gsh = (float*)p;
    p = take(8 * sizeof(float)); if (ws) ws->adam_sc = (float*)p;
    return off;
}

// ------------------------------------------------------------------------------------------------
// vertex-domain prologue
// ------------------------------------------------------------------------------------------------
// vertices = vertices_tmp + delta (mesh_sfs_optim.py:253).  The same launch re-arms the step's accumulators (the packed
// gradient buffer = 3V float4, loss / work-list scratch, SH gradients) so the iteration has no memset nodes.
__global__ void __launch_bounds__(256) ham_vertex_prep_kernel(const float* __restrict__ vtmp,
                                                              const float* __restrict__ delta, int n3,
                                                              float* __restrict__ vertices, float4* __restrict__ packed4,
                                                              uint32_t* __restrict__ z0, int n0,
                                                              uint32_t* __restrict__ z1, int n1,
                                                              uint32_t* __restrict__ z2, int n2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n3) {
        vertices[i] = vtmp[i] + delta[i];
        packed4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (i == 0) packed4[n3] = make_float4(0.f, 0.f, 0.f, 0.f);  // the four scalars behind the 12V floats
    if (i < n0) z0[i] = 0u;
    if (i < n1) z1[i] = 0u;
    if (i < n2) z2[i] = 0u;
}

// clip = ([v,1] @ w2c) @ proj, both matrices stored transposed (mesh_sfs_optim.py:262-264, get_data.py:96-97); also snaps
// every vertex to the rasteriser's fixed-point grid once per (view, vertex) instead of once per (view, triangle corner).
constexpr int kViewsPerBlock = 8;
__global__ void __launch_bounds__(256) ham_transform_kernel(const float* __restrict__ vertices,
                                                            const float* __restrict__ w2cs,
                                                            const float* __restrict__ projs,
                                                            const int32_t* __restrict__ view_idx, int n_views, int V,
                                                            int H, int W, float4* __restrict__ pos,
                                                            int2* __restrict__ snap, float2* __restrict__ scr,
                                                            float* __restrict__ viewM) {
    // A block transforms 256 vertices into kViewsPerBlock views: the dependent preamble (view index -> matrix rows) is
    // paid once per block and hidden behind the vertex load instead of once per (view, vertex) thread - with one view
    // per block the kernel was 8 waves of blocks that lived for two L2 round trips each.
    __shared__ float Wm[kViewsPerBlock][16], Pm[kViewsPerBlock][16];
    const int n0 = blockIdx.y * kViewsPerBlock;
    const int nv = min(kViewsPerBlock, n_views - n0);
    for (int q = threadIdx.x; q < nv * 32; q += blockDim.x) {
        const int slot = q >> 5, e = q & 31;
        const int view = __ldg(view_idx + n0 + slot);
        if (e < 16) Wm[slot][e] = w2cs[(size_t)view * 16 + e];
        else Pm[slot][e - 16] = projs[(size_t)view * 16 + e - 16];
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float x = 0.f, y = 0.f, z = 0.f;
    if (i < V) { x = vertices[3 * (size_t)i]; y = vertices[3 * (size_t)i + 1]; z = vertices[3 * (size_t)i + 2]; }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x < kViewM * nv) {
        const int slot = threadIdx.x / kViewM, k = threadIdx.x - slot * kViewM, r = k >> 2, j = k & 3;
        viewM[(n0 + slot) * kViewM + k] = Wm[slot][4 * r] * Pm[slot][j] + Wm[slot][4 * r + 1] * Pm[slot][4 + j] +
                                          Wm[slot][4 * r + 2] * Pm[slot][8 + j] + Wm[slot][4 * r + 3] * Pm[slot][12 + j];
    }
    if (i >= V) return;
    const float hw = (float)W * 0.5f, hh = (float)H * 0.5f;
    for (int sl = 0; sl < nv; sl++) {
        const float* w = Wm[sl];
        const float* q = Pm[sl];
        const float r0 = x * w[0] + y * w[4] + z * w[8] + w[12], r1 = x * w[1] + y * w[5] + z * w[9] + w[13];
        const float r2 = x * w[2] + y * w[6] + z * w[10] + w[14], r3 = x * w[3] + y * w[7] + z * w[11] + w[15];
        const float4 p = make_float4(r0 * q[0] + r1 * q[4] + r2 * q[8] + r3 * q[12], r0 * q[1] + r1 * q[5] + r2 * q[9] + r3 * q[13],
                                     r0 * q[2] + r1 * q[6] + r2 * q[10] + r3 * q[14], r0 * q[3] + r1 * q[7] + r2 * q[11] + r3 * q[15]);
        const size_t o = (size_t)(n0 + sl) * V + i;
        pos[o] = p;
        int X = kSnapRejected, Y = 0;
        if (!snap_vertex(p, hw, hh, X, Y)) X = kSnapRejected;
        snap[o] = make_int2(X, Y);
        scr[o] = aa_window_xy(p, hw, hh);
    }
}

// ------------------------------------------------------------------------------------------------
// pixel-domain helpers
// ------------------------------------------------------------------------------------------------
struct PixTri {
    int i0, i1, i2;
    float4 p0, p1, p2;
    float u, v;
};

__device__ __forceinline__ void load_pixtri(int t, int px, int py, const float4* __restrict__ P,
                                            const int32_t* __restrict__ tri, float invW, float invH, PixTri& q) {
    q.i0 = __ldg(tri + 3 * t); q.i1 = __ldg(tri + 3 * t + 1); q.i2 = __ldg(tri + 3 * t + 2);
    q.p0 = __ldg(P + q.i0); q.p1 = __ldg(P + q.i1); q.p2 = __ldg(P + q.i2);
    const Bary b = bary_at(q.p0, q.p1, q.p2, px, py, invW, invH);
    q.u = b.u; q.v = b.v;
}

__device__ __forceinline__ float3 interp3(const float* __restrict__ a, const PixTri& q) {
    const float w = 1.0f - q.u - q.v;
    const float* a0 = a + 3 * (size_t)q.i0;
    const float* a1 = a + 3 * (size_t)q.i1;
    const float* a2 = a + 3 * (size_t)q.i2;
    return make_float3(q.u * __ldg(a0) + q.v * __ldg(a1) + w * __ldg(a2),
                       q.u * __ldg(a0 + 1) + q.v * __ldg(a1 + 1) + w * __ldg(a2 + 1),
                       q.u * __ldg(a0 + 2) + q.v * __ldg(a1 + 2) + w * __ldg(a2 + 2));
}

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

// pixel kernels run 16x16 tiles: 2-D locality keeps the hand's pixels in few, densely active blocks
constexpr int kTile = 16;
__device__ __forceinline__ int tile_tid() { return threadIdx.y * kTile + threadIdx.x; }

// clip-space gradient (x, y, -, w) -> world-space xyz
__device__ __forceinline__ float3 clip_to_world(const float* M, float gx, float gy, float gw) {
    return make_float3(M[0] * gx + M[1] * gy + M[3] * gw, M[4] * gx + M[5] * gy + M[7] * gw,
                       M[8] * gx + M[9] * gy + M[11] * gw);
}

__device__ __forceinline__ float block_sum_256(float v, float* sm) {
    const int tid = tile_tid();
    v = warp_sum(v);
    if ((tid & 31) == 0) sm[tid >> 5] = v;
    __syncthreads();
    float r = 0.0f;
    if (tid < 8) r = sm[tid];
    if (tid < 32) {
        r += __shfl_xor_sync(0xffffffffu, r, 4);
        r += __shfl_xor_sync(0xffffffffu, r, 2);
        r += __shfl_xor_sync(0xffffffffu, r, 1);
    }
    __syncthreads();
    return r;  // valid in thread 0
}

// One 16x16 tile of one view, as handed out by the work lists.
struct TileCtx {
    int n, bx, by, nx, ny;  // view slot, tile coordinates, tiles per row / column
};
__device__ __forceinline__ uint32_t tile_encode(int n, int bx, int by) { return ((uint32_t)n << 20) | ((uint32_t)by << 10) | (uint32_t)bx; }
__device__ __forceinline__ TileCtx tile_decode(uint32_t e, int nx, int ny) {
    TileCtx tc;
    tc.n = (int)(e >> 20); tc.by = (int)((e >> 10) & 1023u); tc.bx = (int)(e & 1023u); tc.nx = nx; tc.ny = ny;
    return tc;
}

// Unit of work of the persistent pixel kernels: one warp = one 16x2 strip of a 16x16 tile; warps are independent
// workers (no block barrier anywhere in the pixel passes).
struct Strip {
    TileCtx tc;
    int lane, wib;   // lane in warp, warp in block
    int lx, ly;      // pixel position inside the tile
    int tid;         // ly * 16 + lx = strip * 32 + lane
    int px, py;      // pixel position in the image
};
__device__ __forceinline__ bool next_strip(int& u, const uint32_t* __restrict__ list, int n_tiles, int tiles_x,
                                           int tiles_y, Strip& st) {
    // static striding at strip granularity: unit u, u + (grid warps), ...  (a single global cursor serialises ~10^5
    // same-address atomics per pass; strips of one tile still land on the 8 warps of one block -> shared L1 lines)
    if (u >= n_tiles * 8) return false;
    const int lane = threadIdx.x & 31;
    st.tc = tile_decode(__ldg(list + (u >> 3)), tiles_x, tiles_y);
    const int strip = u & 7;
    st.lane = lane; st.wib = threadIdx.x >> 5;
    st.lx = lane & 15; st.ly = 2 * strip + (lane >> 4);
    st.tid = strip * 32 + lane;
    st.px = st.tc.bx * kTile + st.lx; st.py = st.tc.by * kTile + st.ly;
    u += gridDim.x * 8;
    return true;
}

// 32-way spread accumulators
__device__ __forceinline__ void acc_add(double* acc, int k, float v, const TileCtx& tc) {
    if (v != 0.0f) atomicAdd(acc + k * 32 + ((tc.bx + tc.by * 7 + tc.n * 13) & 31), (double)v);
}
// barrier-free variant: shuffle-reduce inside the warp, one spread fp64 atomic per warp
__device__ __forceinline__ void warp_acc_add(double* acc, int k, float v, const Strip& st) {
    v = warp_sum(v);
    if (st.lane == 0 && v != 0.0f)
        atomicAdd(acc + k * 32 + ((st.tc.bx + st.tc.by * 7 + st.tc.n * 13 + (st.tid >> 5)) & 31), (double)v);
}
__device__ __forceinline__ double acc_total(const double* acc, int k) {
    double s = 0.0;
    for (int j = 0; j < 32; j++) s += acc[k * 32 + j];
    return s;
}
__device__ __forceinline__ int tile_index(const TileCtx& tc) { return tc.by * tc.nx + tc.bx; }
// Appends this tile and its four edge neighbours to the dilated work list (bitmap de-duplicates).
__device__ __forceinline__ void tile_mark_active(const TileCtx& tc, int tid, uint32_t* __restrict__ abits,
                                                 uint32_t* __restrict__ alist, int* __restrict__ acount) {
    int bx = tc.bx, by = tc.by;
    if (tid == 1) bx -= 1; else if (tid == 2) bx += 1; else if (tid == 3) by -= 1; else if (tid == 4) by += 1;
    if (tid > 4 || bx < 0 || by < 0 || bx >= tc.nx || by >= tc.ny) return;
    const int tile = by * tc.nx + bx, words_pv = (tc.nx * tc.ny + 31) >> 5;
    const uint32_t bit = 1u << (tile & 31);
    const uint32_t old = atomicOr(abits + (size_t)tc.n * words_pv + (tile >> 5), bit);
    if (!(old & bit)) alist[atomicAdd(acount, 1)] = tile_encode(tc.n, bx, by);
}

// Per-vertex accumulator layout in `packed` (12 floats = 3 float4):
//   A = (photo_pos.xyz, mask_pos.x)  B = (mask_pos.yz, normal.xy)  C = (normal.z, albedo.bgr)
// photo_* are gradients of the UN-NORMALISED photometric sum  sum |tmp_img - img|, mask_pos of sum (pred - valid)^2 / 2.

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
// Pair work queue (per warp).  A warp owns a 16x2 strip of its block's tile and is responsible for every horizontally /
// vertically adjacent pixel pair with at least one pixel in the strip: the two pairs each of its pixels starts (right,
// down) plus the pairs entering through the strip's top row and the tile's left column.  Pairs that can blend
// (different triangle ids and silhouette bits set on the chosen triangle) are rare (~1 % of pixels), so instead of
// running the ~400-instruction edge analysis under a 3-lane mask inside the per-pixel code they are queued in shared
// memory and analysed afterwards with the queue spread over the lanes.  Everything is warp-synchronous: the pixel passes
// have NO block-level barrier, so the 32+ resident warps of an SM hide each other's gather latency.
//   item = (tid << 2) | which,  which: 0 (self,right)  1 (self,down)  2 (left,self)  3 (up,self)
// ------------------------------------------------------------------------------------------------
constexpr int kPairQueue = 96;  // 2*32 own pairs + 16 top-row pairs + 2 left-column pairs

__device__ __forceinline__ bool pair_needs_analysis(const NbrKeys& k0, const NbrKeys& k1) {
    if (k0.tri == k1.tri) return false;
    // same triangle choice as aa_analyse; the chosen pixel's silhouette bits come from the identical aa_triangle_geom
    // call in the shade pass, so skipping bits == 0 is exact
    const bool from1 = (k0.tri >= 0 && k1.tri >= 0) ? !(k0.zw < k1.zw) : (k0.tri < 0);
    return (from1 ? k1.bits : k0.bits) != 0;
}

__device__ __forceinline__ void enqueue_pairs(const unsigned long long* __restrict__ zb, const Strip& st, int H, int W,
                                              const NbrKeys& self, uint32_t* q_items, int* q_n) {
    const int px = st.px, py = st.py, tid = st.tid;
    if (px >= W || py >= H) return;
    const int rem = py * W + px;
    if (px + 1 < W) {
        const NbrKeys o = decode_key(zb[rem + 1]);
        if (pair_needs_analysis(self, o)) q_items[atomicAdd(q_n, 1)] = ((uint32_t)tid << 2) | 0u;
    }
    if (py + 1 < H) {
        const NbrKeys o = decode_key(zb[rem + W]);
        if (pair_needs_analysis(self, o)) q_items[atomicAdd(q_n, 1)] = ((uint32_t)tid << 2) | 1u;
    }
    if (st.lx == 0 && px > 0) {
        const NbrKeys o = decode_key(zb[rem - 1]);
        if (pair_needs_analysis(o, self)) q_items[atomicAdd(q_n, 1)] = ((uint32_t)tid << 2) | 2u;
    }
    if ((st.ly & 1) == 0 && py > 0) {  // top row of this warp's strip
        const NbrKeys o = decode_key(zb[rem - W]);
        if (pair_needs_analysis(o, self)) q_items[atomicAdd(q_n, 1)] = ((uint32_t)tid << 2) | 3u;
    }
}

struct PairItem {
    int qx, qy, d;        // first pixel of the pair and direction
    int tid0, tid1;       // in-tile thread index of the first / second pixel, -1 when outside this warp's strip
};
__device__ __forceinline__ PairItem decode_pair_item(uint32_t item, const TileCtx& tc) {
    const int tid = (int)(item >> 2), which = (int)(item & 3u);
    const int lx = tid & (kTile - 1), ly = tid >> 4;
    const int px = tc.bx * kTile + lx, py = tc.by * kTile + ly;
    PairItem it;
    it.d = which & 1;
    if (which < 2) {
        it.qx = px; it.qy = py; it.tid0 = tid;
        const int ox = lx + (1 - it.d), oy = ly + it.d;
        it.tid1 = (ox < kTile && (oy >> 1) == (ly >> 1)) ? oy * kTile + ox : -1;  // same 16x2 strip only
    } else {
        it.qx = px - (1 - it.d); it.qy = py - it.d; it.tid0 = -1; it.tid1 = tid;
    }
    return it;
}

// shade:  z-buffer -> shaded colour (phase B) or interpolated normals + albedo (phase A); also tags the key with
// the silhouette bits / valid flag and resets the OTHER z-buffer slot for the next iteration (no separate clear pass).
template <int PHASE>
__device__ __forceinline__ void shade_tile(const Strip& st, unsigned long long* __restrict__ zbuf,
                                           const float4* __restrict__ pos, const float2* __restrict__ scr, float invW,
                                           float invH, const int32_t* __restrict__ tri,
                                           const int32_t* __restrict__ opp, const float* __restrict__ normals,
                                           const float* __restrict__ albedo, const float* __restrict__ masks,
                                           const float* __restrict__ sh_coeffs, const int32_t* __restrict__ view_idx,
                                           const int32_t* __restrict__ sh_idx, int V, int T, int H, int W,
                                           float4* __restrict__ plane0, float4* __restrict__ plane1,
                                           double* __restrict__ acc, bool& feeds_backward, uint32_t& pix32) {
    const TileCtx& tc = st.tc;
    feeds_backward = false;
    pix32 = 0u;
    const int n = tc.n;
    const int view = __ldg(view_idx + n);
    const float* sh = sh_coeffs + (size_t)__ldg(sh_idx + n) * 9;  // block-uniform address: L1 broadcast
    const int hw = H * W;
    const int px = st.px, py = st.py;
    float nvalid = 0.0f;
    if (px < W && py < H) {
        const int rem = py * W + px;
        const size_t pix = (size_t)n * hw + rem;
        const unsigned long long key = zbuf[pix];
        if (key != ZB_EMPTY) {
            const int t = (int)((uint32_t)key & kTriMask);
            const float4* Pv = pos + (size_t)n * V;
            PixTri q;
            load_pixtri(t, px, py, Pv, tri, invW, invH, q);
            AAGeom g;
            g.bits = 0;
            aa_triangle_geom(t, px, py, AAProjScreen{scr + (size_t)n * V}, tri, opp, V, T, H, W, g);
            const float3 m = interp3(normals, q);
            const float3 a = interp3(albedo, q);
            const bool valid = __ldg(masks + (size_t)view * hw + rem) > 0.0f;
            nvalid = valid ? 1.0f : 0.0f;
            feeds_backward = PHASE == 1 ? valid : true;  // phase A back-propagates through every covered pixel's albedo
            pix32 = (uint32_t)pix;
            zbuf[pix] = key | ((unsigned long long)(uint32_t)g.bits << 28) | (valid ? 0x80000000ull : 0ull);
            if (PHASE == 1) {
                float4 col = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) {
                    const float inv = 1.0f / fmaxf(sqrtf(m.x * m.x + m.y * m.y + m.z * m.z), 1e-12f);
                    const float r = sh_radiance(sh, m.x * inv, m.y * inv, m.z * inv);
                    col = make_float4(r * a.x, r * a.y, r * a.z, 1.0f);
                }
                plane0[pix] = col;
            } else {
                plane0[pix] = make_float4(m.x, m.y, m.z, nvalid);
                plane1[pix] = make_float4(a.x, a.y, a.z, 0.0f);
            }
        }
    }
    warp_acc_add(acc, 0, nvalid, st);
}

// shade:  z-buffer -> shaded colour (phase B) or interpolated normals + albedo (phase A); tags every key with the
// silhouette bits / valid flag, resets the tiles the OTHER z-buffer slot dirtied in the previous iteration (no separate
// clear pass) and builds the dilated work list for the two passes that follow.  Persistent: grid = SMs x 4.
template <int PHASE>
__global__ void __launch_bounds__(256) ham_shade_kernel(unsigned long long* __restrict__ zbuf,
                                                        unsigned long long* __restrict__ zbuf_next,
                                                        const uint32_t* __restrict__ tlist, const int* __restrict__ tcount,
                                                        const uint32_t* __restrict__ tlist_next,
                                                        const int* __restrict__ tcount_next,
                                                        uint32_t* __restrict__ abits, uint32_t* __restrict__ alist,
                                                        int* __restrict__ acount, int* __restrict__ cursors,
                                                        int tiles_x, int tiles_y,
                                                        const float4* __restrict__ pos,
                                                        const float2* __restrict__ scr, float invW, float invH,
                                                        const int32_t* __restrict__ tri,
                                                        const int32_t* __restrict__ opp,
                                                        const float* __restrict__ normals,
                                                        const float* __restrict__ albedo,
                                                        const float* __restrict__ masks,
                                                        const float* __restrict__ sh_coeffs,
                                                        const int32_t* __restrict__ view_idx,
                                                        const int32_t* __restrict__ sh_idx, int V, int T, int H, int W,
                                                        float4* __restrict__ plane0, float4* __restrict__ plane1,
                                                        double* __restrict__ acc, uint32_t* __restrict__ vlist,
                                                        int* __restrict__ vcount) {
    // Pixels that feed the backward shader are collected per warp in shared memory and appended to the global compact
    // list with one atomic per ~450 pixels; the backward pass then runs with every lane busy.
    constexpr int kVBuf = 480;
    __shared__ uint32_t vbuf[8][kVBuf];
    int nbuf = 0;  // warp-uniform
    const int wib_ = threadIdx.x >> 5, lane_ = threadIdx.x & 31;
    auto flush = [&]() {
        int base = 0;
        if (lane_ == 0) base = atomicAdd(vcount, nbuf);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (int i = lane_; i < nbuf; i += 32) vlist[base + i] = vbuf[wib_][i];
        __syncwarp();
        nbuf = 0;
    };
    // units = 16x2 strips; two global cursors (reset pass of the other slot, shade pass of this slot)
    Strip st;
    const int u0 = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int nd = *tcount_next;
    int u = u0;
    while (next_strip(u, tlist_next, nd, tiles_x, tiles_y, st))
        if (st.px < W && st.py < H) zbuf_next[((size_t)st.tc.n * H + st.py) * W + st.px] = ZB_EMPTY;
    const int nc = *tcount;
    u = u0;
    while (next_strip(u, tlist, nc, tiles_x, tiles_y, st)) {
        if (st.tid < 5) tile_mark_active(st.tc, st.tid, abits, alist, acount);  // strip 0 of the tile dilates it
        bool fb;
        uint32_t pix32;
        shade_tile<PHASE>(st, zbuf, pos, scr, invW, invH, tri, opp, normals, albedo, masks, sh_coeffs, view_idx, sh_idx, V, T, H,
                          W, plane0, plane1, acc, fb, pix32);
        const unsigned m = __ballot_sync(0xffffffffu, fb);
        if (m) {
            if (fb) vbuf[wib_][nbuf + __popc(m & ((1u << lane_) - 1u))] = pix32;
            nbuf += __popc(m);
            __syncwarp();
            if (nbuf > kVBuf - 32) flush();
        }
    }
    if (nbuf > 0) flush();
}

template <int PHASE>
__device__ __forceinline__ void aa_loss_tile(
    const Strip& st, const unsigned long long* __restrict__ zbuf,
    const float2* __restrict__ scr, const int32_t* __restrict__ tri,
    const int32_t* __restrict__ opp, const float* __restrict__ imgs, const float* __restrict__ valid_masks,
    const float* __restrict__ sh_coeffs, const int32_t* __restrict__ view_idx, const int32_t* __restrict__ sh_idx,
    int V, int T, int H, int W, const float4* __restrict__ plane0, const float4* __restrict__ plane1,
    float4* __restrict__ gplane0, float4* __restrict__ gplane1, double* __restrict__ acc, float* __restrict__ gsh,
    const double* __restrict__ view_vm2, float* __restrict__ dbg_image, float* __restrict__ dbg_mask,
    uint32_t* q_items, int& q_n, float (*blend)[PHASE == 1 ? 4 : 6], uint4* __restrict__ plist_a,
    uint32_t* __restrict__ plist_b, int* __restrict__ pcount, int pcap, int* __restrict__ status) {
    constexpr int NC = PHASE == 1 ? 4 : 6;  // blended channels: (b,g,r,coverage) or (normal xyz, albedo bgr)
    const TileCtx& tc = st.tc;
    const int n = tc.n;
    const int tiles = tc.nx * tc.ny;
    const int view = __ldg(view_idx + n);
    const int tid = st.tid, lane = st.lane;  // blend[] is this warp's: indexed by lane
    const float* sh = sh_coeffs + (size_t)__ldg(sh_idx + n) * 9;  // block-uniform address: L1 broadcast
    const int hw = H * W;
    const int px = st.px, py = st.py;
    const bool inb = px < W && py < H;
    const size_t base = (size_t)n * hw;
    const unsigned long long* zb = zbuf + base;
    const int rem = py * W + px;
    NbrKeys self = decode_key(ZB_EMPTY);
    if (inb) self = decode_key(zb[rem]);
    enqueue_pairs(zb, st, H, W, self, q_items, &q_n);
    __syncwarp();
    // analysis of the queued pairs spread over the lanes; the receiver's blend lands in shared memory
    const int nq = q_n;
    if (nq > 0) {  // warp-uniform
#pragma unroll
        for (int c = 0; c < NC; c++) blend[lane][c] = 0.0f;
        __syncwarp();
        const AAProjScreen proj{scr + (size_t)n * V};
        for (int e = lane; e < nq; e += 32) {
            const PairItem it = decode_pair_item(q_items[e], tc);
            const int r0 = it.qy * W + it.qx, r1 = r0 + (it.d ? W : 1);
            const NbrKeys k0 = decode_key(zb[r0]), k1 = decode_key(zb[r1]);
            AAPair pr;
            if (!aa_analyse(k0.tri, k0.zw, k1.tri, k1.zw, it.qx, it.qy, it.d, proj, tri, opp, V, T, H, W, pr)) continue;
            if (it.tid0 >= 0) {  // the warp whose strip holds the pair's first pixel records it for the backward pass
                const int slot = atomicAdd(pcount, 1);
                if (slot < pcap) {
                    const uint32_t flags = (uint32_t)it.d | ((uint32_t)pr.from1 << 1) | ((uint32_t)pr.clamped << 2) |
                                           ((uint32_t)pr.di << 3);
                    plist_a[slot] = make_uint4((uint32_t)(base + r0), flags, __float_as_uint(pr.alpha), (uint32_t)pr.i1);
                    plist_b[slot] = (uint32_t)pr.i2;
                } else {
                    atomicOr(status, 1);
                }
            }
            const int recv_tid = pr.alpha > 0.0f ? it.tid0 : it.tid1;
            if (recv_tid < 0) continue;  // the receiver belongs to another warp's strip
            const int recv = recv_tid & 31;
            // out[recv] += alpha * (color[second] - color[first]); empty pixels are zero in every channel
            float4 f0 = make_float4(0.f, 0.f, 0.f, 0.f), f1 = f0, s0 = f0, s1 = f0;
            if (k0.tri >= 0) { f0 = plane0[base + r0]; if (PHASE == 0) f1 = plane1[base + r0]; }
            if (k1.tri >= 0) { s0 = plane0[base + r1]; if (PHASE == 0) s1 = plane1[base + r1]; }
            atomicAdd(&blend[recv][0], pr.alpha * (s0.x - f0.x));
            atomicAdd(&blend[recv][1], pr.alpha * (s0.y - f0.y));
            atomicAdd(&blend[recv][2], pr.alpha * (s0.z - f0.z));
            if (PHASE == 1) {
                atomicAdd(&blend[recv][3], pr.alpha * ((k1.tri >= 0 ? 1.0f : 0.0f) - (k0.tri >= 0 ? 1.0f : 0.0f)));
            } else {
                atomicAdd(&blend[recv][3], pr.alpha * (s1.x - f1.x));
                atomicAdd(&blend[recv][4], pr.alpha * (s1.y - f1.y));
                atomicAdd(&blend[recv][5], pr.alpha * (s1.z - f1.z));
            }
        }
        __syncwarp();
        if (lane == 0) q_n = 0;
    }
    float abs_sum = 0.0f, msk_sum = 0.0f;
    float gc[9];
#pragma unroll
    for (int k = 0; k < 9; k++) gc[k] = 0.0f;
    if (inb) {
        const size_t pix = base + rem;
        // own (pre-antialias) values; empty pixels are zero in every channel
        float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (self.tri >= 0) {
            c0 = plane0[pix];
            if (PHASE == 0) c1 = plane1[pix];
        }
        float4 a0 = c0, a1 = c1;  // antialiased values
        float amask = self.tri >= 0 ? 1.0f : 0.0f;
        if (nq > 0) {
            a0.x += blend[lane][0]; a0.y += blend[lane][1]; a0.z += blend[lane][2];
            if (PHASE == 1) amask += blend[lane][3];
            else { a1.x += blend[lane][3]; a1.y += blend[lane][4]; a1.z += blend[lane][5]; }
        }
        const bool valid = self.valid;
        const float* img = imgs + ((size_t)view * hw + rem) * 3;
        if (PHASE == 1) {
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {  // mesh_sfs_optim.py:289  l1 over tmp_img[valid_idx]
                const float d0 = a0.x - __ldg(img), d1 = a0.y - __ldg(img + 1), d2 = a0.z - __ldg(img + 2);
                abs_sum = fabsf(d0) + fabsf(d1) + fabsf(d2);
                g.x = (d0 > 0.f) - (d0 < 0.f); g.y = (d1 > 0.f) - (d1 < 0.f); g.z = (d2 > 0.f) - (d2 < 0.f);
            }
            // mesh_sfs_optim.py:295  mean((pred_mask - valid_mask)^2).  Inactive tiles have pred_mask == 0 and contribute
            // valid_mask^2, a per-tile constant of the view (buffers.view_vm2): they are never visited.
            const float dm = amask - __ldg(valid_masks + (size_t)view * hw + rem);
            msk_sum = dm * dm;
            g.w = dm;
            gplane0[pix] = g;
            if (dbg_image) { dbg_image[pix * 3] = a0.x; dbg_image[pix * 3 + 1] = a0.y; dbg_image[pix * 3 + 2] = a0.z; }
            if (dbg_mask) dbg_mask[pix] = amask;
        } else {
            // phase A (mesh_sfs_optim.py:217-230): a0 = antialiased normals, a1 = antialiased albedo
            float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0;
            float3 pred = make_float3(0.f, 0.f, 0.f);
            if (valid) {
                const float len = sqrtf(a0.x * a0.x + a0.y * a0.y + a0.z * a0.z);
                const float inv = 1.0f / fmaxf(len, 1e-12f);
                const float nx = a0.x * inv, ny = a0.y * inv, nz = a0.z * inv;
                const float r = sh_radiance(sh, nx, ny, nz);
                pred = make_float3(r * a1.x, r * a1.y, r * a1.z);
                const float d0 = pred.x - __ldg(img), d1 = pred.y - __ldg(img + 1), d2 = pred.z - __ldg(img + 2);
                abs_sum = fabsf(d0) + fabsf(d1) + fabsf(d2);
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
    warp_acc_add(acc, 1, abs_sum, st);
    if (PHASE == 1) {
        warp_acc_add(acc, 2, msk_sum, st);
        // this tile is accounted for explicitly: remove its constant share (exact in fp64)
        if (tid == 0)  // strip 0, lane 0: once per tile
            atomicAdd(acc + 7 * 32 + ((tc.bx + tc.by * 7 + tc.n * 13) & 31),
                      view_vm2[(size_t)view * (tiles + 1) + tile_index(tc)]);
    } else {
#pragma unroll
        for (int k = 0; k < 9; k++) {
            const float sk = warp_sum(gc[k]);
            if (lane == 0 && sk != 0.0f) atomicAdd(gsh + (size_t)__ldg(sh_idx + n) * 9 + k, sk);
        }
    }
    __syncwarp();  // blend[] / q_items / q_n of this warp are rewritten by its next tile
}

// antialias (gather form, dense pair queue) + losses; persistent over the dilated work list
template <int PHASE>
__global__ void __launch_bounds__(256, 4) ham_aa_loss_kernel(
    const unsigned long long* __restrict__ zbuf, const uint32_t* __restrict__ alist, const int* __restrict__ acount,
    int* __restrict__ cursors, int tiles_x, int tiles_y, const float2* __restrict__ scr, const int32_t* __restrict__ tri,
    const int32_t* __restrict__ opp, const float* __restrict__ imgs, const float* __restrict__ valid_masks,
    const float* __restrict__ sh_coeffs, const int32_t* __restrict__ view_idx, const int32_t* __restrict__ sh_idx,
    int V, int T, int H, int W, const float4* __restrict__ plane0, const float4* __restrict__ plane1,
    float4* __restrict__ gplane0, float4* __restrict__ gplane1, double* __restrict__ acc, float* __restrict__ gsh,
    const double* __restrict__ view_vm2, float* __restrict__ dbg_image, float* __restrict__ dbg_mask,
    uint4* __restrict__ plist_a, uint32_t* __restrict__ plist_b, int* __restrict__ pcount, int pcap,
    int* __restrict__ status) {
    __shared__ uint32_t q_items[8][kPairQueue];
    __shared__ int q_n[8];
    __shared__ float blend[8][32][PHASE == 1 ? 4 : 6];
    const int wib = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) q_n[wib] = 0;
    __syncwarp();
    const int na = *acount;
    Strip st;
    int u = blockIdx.x * 8 + wib;
    while (next_strip(u, alist, na, tiles_x, tiles_y, st))
        aa_loss_tile<PHASE>(st, zbuf, scr, tri, opp, imgs, valid_masks, sh_coeffs, view_idx, sh_idx, V, T, H, W, plane0,
                            plane1, gplane0, gplane1, acc, gsh, view_vm2, dbg_image, dbg_mask, q_items[wib], q_n[wib],
                            blend[wib], plist_a, plist_b, pcount, pcap, status);
}

// ------------------------------------------------------------------------------------------------
// pixel backward: antialias bwd (gather form for colours, owner-scatter for positions), shading bwd,
// interpolate bwd, rasterize bwd; everything lands in the per-vertex world-space accumulators.
// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// backward, part 1: the blending pairs recorded by the antialias pass (a few 10^4 per iteration, one thread each).
//   * colour / albedo gradient of the two pixels of the pair -> gdelta plane (only pixels the main pass will visit);
//   * phase B: silhouette position gradient -> world-space accumulators.
// ------------------------------------------------------------------------------------------------
template <int PHASE>
__global__ void __launch_bounds__(256) ham_pair_bwd_kernel(
    const uint4* __restrict__ plist_a, const uint32_t* __restrict__ plist_b, const int* __restrict__ pcount, int pcap,
    const unsigned long long* __restrict__ zbuf, const float4* __restrict__ pos, const float* __restrict__ viewM, int V,
    int H, int W, const float4* __restrict__ plane0, const float4* __restrict__ gplane0,
    const float4* __restrict__ gplane1, float4* __restrict__ gdelta, float4* __restrict__ G) {
    const int np = min(*pcount, pcap);
    const int hw = H * W;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < np; e += gridDim.x * blockDim.x) {
        const uint4 a = plist_a[e];
        AAPair pr;
        const size_t pix0 = a.x;
        const int d = (int)(a.y & 1u);
        pr.from1 = (int)((a.y >> 1) & 1u); pr.clamped = (int)((a.y >> 2) & 1u); pr.di = (int)((a.y >> 3) & 3u);
        pr.alpha = __uint_as_float(a.z); pr.i1 = (int)a.w; pr.i2 = (int)plist_b[e]; pr.tri = 0;
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
        if (t0) atomicAdd(gdelta + pix0, make_float4(-pr.alpha * gr.x, -pr.alpha * gr.y, -pr.alpha * gr.z, 0.f));
        if (t1) atomicAdd(gdelta + pix1, make_float4(pr.alpha * gr.x, pr.alpha * gr.y, pr.alpha * gr.z, 0.f));
        if (PHASE == 1 && !pr.clamped) {
            float4 f0 = make_float4(0.f, 0.f, 0.f, 0.f), s0 = f0;
            if (k0.tri >= 0) f0 = plane0[pix0];
            if (k1.tri >= 0) s0 = plane0[pix1];
            const float dd_img = gr.x * (s0.x - f0.x) + gr.y * (s0.y - f0.y) + gr.z * (s0.z - f0.z);
            const float dd_msk = gr.w * ((k1.tri >= 0 ? 1.0f : 0.0f) - (k0.tri >= 0 ? 1.0f : 0.0f));
            if (dd_img != 0.0f || dd_msk != 0.0f) {
                const float* M = viewM + (size_t)n * kViewM;
                float4 e1, e2;
                aa_pos_grad(pr, qx, qy, d, reinterpret_cast<const float*>(pos + (size_t)n * V), H, W, 1.0f, e1, e2);
                const float3 w1 = clip_to_world(M, e1.x, e1.y, e1.w);
                const float3 w2 = clip_to_world(M, e2.x, e2.y, e2.w);
                atomicAdd(G + 3 * (size_t)pr.i1, make_float4(dd_img * w1.x, dd_img * w1.y, dd_img * w1.z, dd_msk * w1.x));
                atomicAdd(G + 3 * (size_t)pr.i1 + 1, make_float4(dd_msk * w1.y, dd_msk * w1.z, 0.f, 0.f));
                atomicAdd(G + 3 * (size_t)pr.i2, make_float4(dd_img * w2.x, dd_img * w2.y, dd_img * w2.z, dd_msk * w2.x));
                atomicAdd(G + 3 * (size_t)pr.i2 + 1, make_float4(dd_msk * w2.y, dd_msk * w2.z, 0.f, 0.f));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward, part 2: one thread per pixel of the compact list (every lane busy): SH / normalise / interpolate /
// rasterize backward, 9 float4 red.global.add per pixel into the world-space per-vertex accumulators.
// ------------------------------------------------------------------------------------------------
template <int PHASE>
__global__ void __launch_bounds__(256, 4) ham_pixel_bwd_kernel(
    const uint32_t* __restrict__ vlist, const int* __restrict__ vcount, const unsigned long long* __restrict__ zbuf,
    const float4* __restrict__ pos, float invW, float invH, const float* __restrict__ viewM,
    const int32_t* __restrict__ tri, const float* __restrict__ normals, const float* __restrict__ albedo,
    const float* __restrict__ sh_coeffs, const int32_t* __restrict__ sh_idx, int V, int H, int W,
    const float4* __restrict__ gplane0, const float4* __restrict__ gplane1, float4* __restrict__ gdelta,
    float4* __restrict__ G) {
    const int nv = *vcount;
    const int hw = H * W;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nv; e += gridDim.x * blockDim.x) {
        const size_t pix = vlist[e];
        const int n = (int)(pix / hw);
        const int rem = (int)(pix - (size_t)n * hw), py = rem / W, px = rem - py * W;
        const NbrKeys self = decode_key(zbuf[pix]);
        const float* M = viewM + (size_t)n * kViewM;
        const float* c = sh_coeffs + (size_t)__ldg(sh_idx + n) * 9;
        const float4* Pv = pos + (size_t)n * V;
        // gradient w.r.t. this pixel's PRE-antialias values: pass-through + pair terms (consumed and re-armed)
        float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0;
        const float4 gd = gdelta[pix];
        if (gd.x != 0.0f || gd.y != 0.0f || gd.z != 0.0f) gdelta[pix] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (PHASE == 1) {
            g0 = gplane0[pix];
            g0.x += gd.x; g0.y += gd.y; g0.z += gd.z;
        } else {
            g1 = gplane1[pix];
            g1.x += gd.x; g1.y += gd.y; g1.z += gd.z;
        }
        PixTri q;
        load_pixtri(self.tri, px, py, Pv, tri, invW, invH, q);
        const float w = 1.0f - q.u - q.v;
        if (PHASE == 0) {
            // only the albedo attribute is trainable: interpolate bwd
            if (g1.x == 0.0f && g1.y == 0.0f && g1.z == 0.0f) continue;
            atomicAdd(G + 3 * (size_t)q.i0 + 2, make_float4(0.f, q.u * g1.x, q.u * g1.y, q.u * g1.z));
            atomicAdd(G + 3 * (size_t)q.i1 + 2, make_float4(0.f, q.v * g1.x, q.v * g1.y, q.v * g1.z));
            atomicAdd(G + 3 * (size_t)q.i2 + 2, make_float4(0.f, w * g1.x, w * g1.y, w * g1.z));
            continue;
        }
        // phase B: tmp_img[valid_idx] = pred_img -> only valid pixels feed the shader (mesh_sfs_optim.py:285-286)
        if (!self.valid) continue;
        if (g0.x == 0.0f && g0.y == 0.0f && g0.z == 0.0f) continue;
        const float3 m = interp3(normals, q);
        const float3 a = interp3(albedo, q);
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
        const float* n0 = normals + 3 * (size_t)q.i0; const float* n1 = normals + 3 * (size_t)q.i1; const float* n2 = normals + 3 * (size_t)q.i2;
        const float* b0 = albedo + 3 * (size_t)q.i0; const float* b1 = albedo + 3 * (size_t)q.i1; const float* b2 = albedo + 3 * (size_t)q.i2;
        const float n2x = __ldg(n2), n2y = __ldg(n2 + 1), n2z = __ldg(n2 + 2);
        const float b2x = __ldg(b2), b2y = __ldg(b2 + 1), b2z = __ldg(b2 + 2);
        const float du = gm.x * (__ldg(n0) - n2x) + gm.y * (__ldg(n0 + 1) - n2y) + gm.z * (__ldg(n0 + 2) - n2z) +
                         ga.x * (__ldg(b0) - b2x) + ga.y * (__ldg(b0 + 1) - b2y) + ga.z * (__ldg(b0 + 2) - b2z);
        const float dv = gm.x * (__ldg(n1) - n2x) + gm.y * (__ldg(n1 + 1) - n2y) + gm.z * (__ldg(n1 + 2) - n2z) +
                         ga.x * (__ldg(b1) - b2x) + ga.y * (__ldg(b1 + 1) - b2y) + ga.z * (__ldg(b1 + 2) - b2z);
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
        float4* G0 = G + 3 * (size_t)q.i0; float4* G1 = G + 3 * (size_t)q.i1; float4* G2 = G + 3 * (size_t)q.i2;
        atomicAdd(G0, make_float4(w0.x, w0.y, w0.z, 0.f));
        atomicAdd(G0 + 1, make_float4(0.f, 0.f, q.u * gm.x, q.u * gm.y));
        atomicAdd(G0 + 2, make_float4(q.u * gm.z, q.u * ga.x, q.u * ga.y, q.u * ga.z));
        atomicAdd(G1, make_float4(w1.x, w1.y, w1.z, 0.f));
        atomicAdd(G1 + 1, make_float4(0.f, 0.f, q.v * gm.x, q.v * gm.y));
        atomicAdd(G1 + 2, make_float4(q.v * gm.z, q.v * ga.x, q.v * ga.y, q.v * ga.z));
        atomicAdd(G2, make_float4(w2.x, w2.y, w2.z, 0.f));
        atomicAdd(G2 + 1, make_float4(0.f, 0.f, w * gm.x, w * gm.y));
        atomicAdd(G2 + 2, make_float4(w * gm.z, w * ga.x, w * ga.y, w * ga.z));
    }
}

__global__ void ham_finalize_scalars_kernel(const double* __restrict__ acc, const double* __restrict__ view_vm2,
                                            const int32_t* __restrict__ view_idx, int n_views, int tiles, int phase,
                                            float* __restrict__ scal) {
    if (threadIdx.x == 0) {
        double vm2 = 0.0;  // sum of valid_mask^2 over the tiles no block visited = view totals - visited tiles
        if (phase == 1) {
            for (int n = 0; n < n_views; n++) vm2 += view_vm2[(size_t)view_idx[n] * (tiles + 1) + tiles];
            vm2 -= acc_total(acc, 7);
        }
        scal[0] = (float)acc_total(acc, 0);
        scal[1] = (float)acc_total(acc, 1);
        scal[2] = (float)(acc_total(acc, 2) + vm2);
        scal[3] = 0.0f;
    }
}

// view_vm2[view][tile] = sum of valid_mask^2 over the 16x16 tile, view_vm2[view][tiles] = sum over the view (constants
// of the optimisation: valid_masks never change, mesh_sfs_optim.py:163).  The view total is the fp64 sum of the
// tile sums, so "total - visited tiles" cancels exactly when every non-zero tile is visited.
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
    const int tid = tile_tid();
    if ((tid & 31) == 0) red[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; i++) t += red[i];
        out[(size_t)view * (tiles + 1) + blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
}
__global__ void ham_view_vm2_total_kernel(int tiles, double* __restrict__ out) {
    double* row = out + (size_t)blockIdx.x * (tiles + 1);
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < tiles; i++) t += row[i];
        row[tiles] = t;
    }
}

// zbuf -> rast_out for fmhr_ham_debug_export
__global__ void __launch_bounds__(256) ham_export_rast_kernel(const unsigned long long* __restrict__ zbuf,
                                                              const float4* __restrict__ pos,
                                                              const int32_t* __restrict__ tri, int V, int H, int W,
                                                              float4* __restrict__ rast) {
    const int n = blockIdx.y, hw = H * W;
    const int rem = blockIdx.x * blockDim.x + threadIdx.x;
    if (rem >= hw) return;
    const size_t pix = (size_t)n * hw + rem;
    const unsigned long long key = zbuf[pix];
    if (key == ZB_EMPTY) { rast[pix] = make_float4(0.f, 0.f, 0.f, 0.f); return; }
    const int t = (int)((uint32_t)key & kTriMask), py = rem / W, px = rem - py * W;
    const float4* P = pos + (size_t)n * V;
    const Bary b = bary_at(__ldg(P + __ldg(tri + 3 * t)), __ldg(P + __ldg(tri + 3 * t + 1)),
                           __ldg(P + __ldg(tri + 3 * t + 2)), px, py, xd(1.0f, (float)W), xd(1.0f, (float)H));
    rast[pix] = make_float4(b.u, b.v, b.zw, (float)(t + 1));
}

// ------------------------------------------------------------------------------------------------
// update: regularisers, normal backward, normalisation, Adam
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float3 ldf3(const float* p) { return make_float3(__ldg(p), __ldg(p + 1), __ldg(p + 2)); }

// The vertex-domain kernels below use 8 lanes per vertex (kLPV): V is only ~50k, so one thread per vertex leaves the
// GPU at <0.3 waves of latency-bound gather loops; splitting each vertex's ~6 neighbours / incident faces over 8 lanes
// (shuffle-reduced) gives 8x the memory-level parallelism.
constexpr int kLPV = 8;
__device__ __forceinline__ float sub_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// pass 1: Laplacian forward for vertices and albedo, edge/delta losses, projector of the normal backward, Adam scalars
__global__ void __launch_bounds__(256) ham_update_pass1_kernel(
    fmhr_ham_config cfg, const float* __restrict__ vertices, const float* __restrict__ delta,
    const float* __restrict__ albedo, const int32_t* __restrict__ tri, const int32_t* __restrict__ v2f_ptr,
    const int2* __restrict__ v2f_nbr, const int32_t* __restrict__ v2v_ptr, const int32_t* __restrict__ v2v_idx,
    const float* __restrict__ raw, const float* __restrict__ packed, float* __restrict__ yhat_v,
    float* __restrict__ yhat_a, float* __restrict__ gN, double* __restrict__ acc, int32_t* __restrict__ adam_step,
    float* __restrict__ adam_sc) {
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
            const size_t nb = 3 * (size_t)__ldg(v2v_idx + j);
            const float3 xv = ldf3(vertices + nb), xa_ = ldf3(albedo + nb);
            sv.x += xv.x; sv.y += xv.y; sv.z += xv.z;
            sa.x += xa_.x; sa.y += xa_.y; sa.z += xa_.z;
        }
        vi = ldf3(vertices + 3 * (size_t)i);
        // edge hinge (mesh_sfs_optim.py:296-302): every half-edge is seen from both of its endpoints -> weight 1/2
        const int fb = __ldg(v2f_ptr + i), fe = __ldg(v2f_ptr + i + 1);
        for (int j = fb + sub; j < fe; j += kLPV) {
            const int2 nb = __ldg(v2f_nbr + j);
#pragma unroll
            for (int s = 1; s <= 2; s++) {
                const float3 o = ldf3(vertices + 3 * (size_t)(s == 1 ? nb.x : nb.y));
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
        const float3 ai = ldf3(albedo + 3 * (size_t)i);
        sv = make_float3(sv.x * invd - vi.x, sv.y * invd - vi.y, sv.z * invd - vi.z);
        sa = make_float3(sa.x * invd - ai.x, sa.y * invd - ai.y, sa.z * invd - ai.z);
        lv = sqrtf(sv.x * sv.x + sv.y * sv.y + sv.z * sv.z);
        la = sqrtf(sa.x * sa.x + sa.y * sa.y + sa.z * sa.z);
        const float iv = lv > 0.f ? 1.0f / lv : 0.f, ia = la > 0.f ? 1.0f / la : 0.f;
        yhat_v[3 * (size_t)i] = sv.x * iv; yhat_v[3 * (size_t)i + 1] = sv.y * iv; yhat_v[3 * (size_t)i + 2] = sv.z * iv;
        yhat_a[3 * (size_t)i] = sa.x * ia; yhat_a[3 * (size_t)i + 1] = sa.y * ia; yhat_a[3 * (size_t)i + 2] = sa.z * ia;
        const float3 di = ldf3(delta + 3 * (size_t)i);
        ld = di.x * di.x + di.y * di.y + di.z * di.z;
        // normal backward, step 1: through the normalisation (un-normalised photometric scale; linear, scaled later)
        const float* Gi = packed + 12 * (size_t)i;
        const float3 g = make_float3(Gi[6], Gi[7], Gi[8]);
        const float3 N = ldf3(raw + 3 * (size_t)i);
        const float len = sqrtf(N.x * N.x + N.y * N.y + N.z * N.z);
        float3 r;
        if (len > 1e-6f) {
            const float inv = 1.0f / len;
            const float3 nh = make_float3(N.x * inv, N.y * inv, N.z * inv);
            const float d = nh.x * g.x + nh.y * g.y + nh.z * g.z;
            r = make_float3((g.x - nh.x * d) * inv, (g.y - nh.y * d) * inv, (g.z - nh.z * d) * inv);
        } else {
            r = make_float3(g.x * 1e6f, g.y * 1e6f, g.z * 1e6f);
        }
        gN[3 * (size_t)i] = r.x; gN[3 * (size_t)i + 1] = r.y; gN[3 * (size_t)i + 2] = r.z;
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
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        // Adam bias corrections (torch.optim.Adam defaults); which parameters step depends on the phase:
        //   phase A: albedo, sh   (mesh_sfs_optim.py:193)     phase B: delta, albedo   (:242-244, sh has no grad)
        const int steps[3] = {cfg.phase == 1, 1, cfg.phase == 0};
        const float lrs[3] = {cfg.lr, cfg.albedo_lr, cfg.sh_lr};
        for (int k = 0; k < 3; k++) {
            const int t = adam_step[k] + steps[k];
            adam_step[k] = t;
            const double b1 = 1.0 - pow((double)cfg.beta1, (double)max(t, 1));
            const double b2 = 1.0 - pow((double)cfg.beta2, (double)max(t, 1));
            adam_sc[2 * k] = (float)((double)lrs[k] / b1);
            adam_sc[2 * k + 1] = (float)sqrt(b2);
        }
    }
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

// pass 2: gather every gradient term per vertex (8 lanes each), then Adam on delta (phase B) and albedo
__global__ void __launch_bounds__(256) ham_update_pass2_kernel(
    fmhr_ham_config cfg, const float* __restrict__ vertices, float* __restrict__ delta, float* __restrict__ albedo,
    const int32_t* __restrict__ tri, const int32_t* __restrict__ v2f_ptr, const int2* __restrict__ v2f_nbr,
    const int32_t* __restrict__ v2v_ptr, const int32_t* __restrict__ v2v_idx, const float* __restrict__ inv_deg,
    const float* __restrict__ packed,
    const float* __restrict__ yhat_v, const float* __restrict__ yhat_a, const float* __restrict__ gN,
    float* __restrict__ adam_m, float* __restrict__ adam_v, const float* __restrict__ adam_sc,
    const double* __restrict__ acc, float* __restrict__ losses, float* __restrict__ dbg_grad,
    const int* __restrict__ status) {
    const int V = cfg.V;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) / kLPV, sub = threadIdx.x & (kLPV - 1);
    const float* scal = packed + 12 * (size_t)V;
    const float n_valid = scal[0];
    const float P_global = (float)cfg.n_views_global * (float)cfg.H * (float)cfg.W;
    const float s_photo = cfg.sfs_weight / (3.0f * n_valid);            // F.l1_loss mean over [N_valid,3]
    const float s_mask = 2.0f * cfg.mask_weight / P_global;              // F.mse_loss mean over n*H*W
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const float sfs = cfg.sfs_weight * scal[1] / (3.0f * n_valid);
        const float lap = cfg.lap_weight * (float)(acc_total(acc, 3) / (double)V);
        const float alb = cfg.albedo_weight * (float)(acc_total(acc, 4) / (double)V);
        const float msk = cfg.phase == 1 ? cfg.mask_weight * scal[2] / P_global : 0.0f;
        const float edg = cfg.edge_weight * (float)(acc_total(acc, 5) / (3.0 * (double)cfg.T));
        const float del = cfg.delta_weight * (float)(acc_total(acc, 6) / (double)V);
        losses[0] = sfs; losses[1] = cfg.phase == 1 ? lap : 0.0f; losses[2] = alb; losses[3] = msk;
        losses[4] = cfg.phase == 1 ? edg : 0.0f; losses[5] = cfg.phase == 1 ? del : 0.0f; losses[6] = n_valid;
        losses[7] = cfg.phase == 1 ? sfs + lap + alb + msk + edg + del : sfs;
        if (status && (*status & 1)) losses[7] = __int_as_float(0x7fc00000);  // pair list overflow: refuse a number
    }
    // Laplacian backward rows (L^T yhat) for vertices and albedo; normal backward + edge hinge over incident faces
    float3 lv = make_float3(0.f, 0.f, 0.f), la = lv, gnb = lv, ge = lv, vi = lv;
    if (i < V) {
        const int b = __ldg(v2v_ptr + i), e = __ldg(v2v_ptr + i + 1);
        for (int q = b + sub; q < e; q += kLPV) {
            const int j = __ldg(v2v_idx + q);
            const float invd = __ldg(inv_deg + j);
            const float3 ya = ldf3(yhat_a + 3 * (size_t)j);
            la.x += ya.x * invd; la.y += ya.y * invd; la.z += ya.z * invd;
            if (cfg.phase == 1) {
                const float3 yv = ldf3(yhat_v + 3 * (size_t)j);
                lv.x += yv.x * invd; lv.y += yv.y * invd; lv.z += yv.z * invd;
            }
        }
        if (cfg.phase == 1) {
            vi = ldf3(vertices + 3 * (size_t)i);
            const float3 gi = ldf3(gN + 3 * (size_t)i);
            const int fb = __ldg(v2f_ptr + i), fe = __ldg(v2f_ptr + i + 1);
            for (int j = fb + sub; j < fe; j += kLPV) {
                const int2 nb = __ldg(v2f_nbr + j);
                const int ia = nb.x, ib = nb.y;
                const float3 pa = ldf3(vertices + 3 * (size_t)ia), pb = ldf3(vertices + 3 * (size_t)ib);
                const float3 ka = ldf3(gN + 3 * (size_t)ia), kb = ldf3(gN + 3 * (size_t)ib);
                const float3 Gs = make_float3(gi.x + ka.x + kb.x, gi.y + ka.y + kb.y, gi.z + ka.z + kb.z);
                const float3 ed = make_float3(pa.x - pb.x, pa.y - pb.y, pa.z - pb.z);
                gnb.x += ed.y * Gs.z - ed.z * Gs.y; gnb.y += ed.z * Gs.x - ed.x * Gs.z; gnb.z += ed.x * Gs.y - ed.y * Gs.x;
                const float3 o[2] = {pa, pb};
#pragma unroll
                for (int s = 0; s < 2; s++) {
                    const float dx = vi.x - o[s].x, dy = vi.y - o[s].y, dz = vi.z - o[s].z;
                    const float x = dx * dx + dy * dy + dz * dz - cfg.edge_length_mean;
                    if (x >= 0.0f && x <= 1.0f) { ge.x += 2.0f * dx; ge.y += 2.0f * dy; ge.z += 2.0f * dz; }
                }
            }
        }
    }
    la.x = sub_sum(la.x); la.y = sub_sum(la.y); la.z = sub_sum(la.z);
    if (cfg.phase == 1) {
        lv.x = sub_sum(lv.x); lv.y = sub_sum(lv.y); lv.z = sub_sum(lv.z);
        gnb.x = sub_sum(gnb.x); gnb.y = sub_sum(gnb.y); gnb.z = sub_sum(gnb.z);
        ge.x = sub_sum(ge.x); ge.y = sub_sum(ge.y); ge.z = sub_sum(ge.z);
    }
    if (i >= V || sub != 0) return;
    const float* Gi = packed + 12 * (size_t)i;
    {
        const float3 yv = ldf3(yhat_v + 3 * (size_t)i), ya = ldf3(yhat_a + 3 * (size_t)i);
        const float iv = 1.0f / (float)V;
        lv = make_float3((lv.x - yv.x) * iv, (lv.y - yv.y) * iv, (lv.z - yv.z) * iv);
        la = make_float3((la.x - ya.x) * iv, (la.y - ya.y) * iv, (la.z - ya.z) * iv);
    }
    // albedo gradient (phase A: photometric only, mesh_sfs_optim.py:233; phase B adds the albedo Laplacian)
    float3 ga = make_float3(s_photo * Gi[9], s_photo * Gi[10], s_photo * Gi[11]);
    if (cfg.phase == 1) { ga.x += cfg.albedo_weight * la.x; ga.y += cfg.albedo_weight * la.y; ga.z += cfg.albedo_weight * la.z; }
    float3 gd = make_float3(0.f, 0.f, 0.f);
    if (cfg.phase == 1) {
        const float s_edge = cfg.edge_weight / (3.0f * (float)cfg.T);
        const float s_delta = 2.0f * cfg.delta_weight / (float)V;
        const float3 di = ldf3(delta + 3 * (size_t)i);
        gd.x = s_photo * (Gi[0] + gnb.x) + s_mask * Gi[3] + cfg.lap_weight * lv.x + s_edge * ge.x + s_delta * di.x;
        gd.y = s_photo * (Gi[1] + gnb.y) + s_mask * Gi[4] + cfg.lap_weight * lv.y + s_edge * ge.y + s_delta * di.y;
        gd.z = s_photo * (Gi[2] + gnb.z) + s_mask * Gi[5] + cfg.lap_weight * lv.z + s_edge * ge.z + s_delta * di.z;
    }
    if (dbg_grad) {
        dbg_grad[6 * (size_t)i] = gd.x; dbg_grad[6 * (size_t)i + 1] = gd.y; dbg_grad[6 * (size_t)i + 2] = gd.z;
        dbg_grad[6 * (size_t)i + 3] = ga.x; dbg_grad[6 * (size_t)i + 4] = ga.y; dbg_grad[6 * (size_t)i + 5] = ga.z;
    }
    const float gds[3] = {gd.x, gd.y, gd.z}, gas[3] = {ga.x, ga.y, ga.z};
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const size_t k = 3 * (size_t)i + c;
        if (cfg.phase == 1)
            delta[k] = adam_update(delta[k], gds[c], adam_m + k, adam_v + k, cfg.beta1, cfg.beta2, cfg.eps, adam_sc[0],
                                   adam_sc[1]);
        const size_t ka = 3 * (size_t)V + k;
        albedo[k] = adam_update(albedo[k], gas[c], adam_m + ka, adam_v + ka, cfg.beta1, cfg.beta2, cfg.eps, adam_sc[2],
                                adam_sc[3]);
    }
}

// phase A: Adam on sh_coeffs.  torch.optim.Adam steps the WHOLE [num,9] tensor every iteration (rows of views outside
// the batch have zero gradient but keep moving on their momentum, mesh_sfs_optim.py:193,205,237), so this runs over
// every SH row resident on this rank; rows are view-local and never reduced across ranks.
__global__ void ham_update_sh_kernel(fmhr_ham_config cfg, const float* __restrict__ gsh,
                                     const float* __restrict__ packed, float* __restrict__ sh_coeffs,
                                     float* __restrict__ adam_m, float* __restrict__ adam_v,
                                     const float* __restrict__ adam_sc, float* __restrict__ dbg_grad_sh) {
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
    FMHR_CHECK_ARG(c->T < (1 << 28));  // triangle id shares the key's low word with 4 tag bits
    FMHR_CHECK_ARG(c->n_views < 4096 && c->W <= 16384 && c->H <= 16384);  // work-list entry = view:12 | ty:10 | tx:10
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
    FMHR_CHECK_ARG(b->workspace_bytes >= fmhr_ham_workspace_bytes(cfg));
    FMHR_CHECK_ARG(((uintptr_t)b->packed & 15) == 0 && ((uintptr_t)b->workspace & 255) == 0);
    return FMHR_OK;
}

template <int PHASE>
static int ham_render_impl(const fmhr_ham_config* cfg, const fmhr_ham_buffers* b, cudaStream_t st, float* dbg_image,
                           float* dbg_mask, bool forward_only) {
    HamWs ws;
    ham_layout(cfg, (char*)b->workspace, &ws);
    const int V = cfg->V, T = cfg->T, H = cfg->H, W = cfg->W, n = cfg->n_views;
    const size_t P = (size_t)n * H * W;
    unsigned long long* zcur = ws.zbuf[cfg->zbuf_slot];
    unsigned long long* znext = ws.zbuf[cfg->zbuf_slot ^ 1];
    const int cur = cfg->zbuf_slot, nxt = cfg->zbuf_slot ^ 1;
    const int tiles_x = cdiv(W, kTile), tiles_y = cdiv(H, kTile);
    const float invW = 1.0f / (float)W, invH = 1.0f / (float)H;  // IEEE single divides, identical to the device's __fdiv_rn
    const int32_t* sh_idx = b->sh_idx ? b->sh_idx : b->view_idx;
    FMHR_STAGE_MARK();  // 0: (no clear pass any more)
    // zeroed by the prep kernel: packed, loss accumulators + dilated work list (common), the work list of the slot
    // rasterised this step, SH gradients (phase A)
    const int n0 = (int)(ws.common_bytes / 4), n1 = (int)(ws.slot_bytes / 4), n2 = PHASE == 0 ? cfg->n_sh_rows * 9 : 0;
    const int prep_threads = max(3 * V, max(n0, max(n1, n2)));
    ham_vertex_prep_kernel<<<cdiv(prep_threads, 256), 256, 0, st>>>(
        b->vertices_tmp, b->delta, 3 * V, ws.vertices, (float4*)b->packed, (uint32_t*)ws.common_region, n0,
        (uint32_t*)ws.slot_region[cfg->zbuf_slot], n1, (uint32_t*)ws.gsh, n2);
    FMHR_LAUNCH_CHECK();
    int rc = launch_vertex_normals_fwd(ws.vertices, b->tri, b->v2f_ptr, b->v2f_idx, b->v2f_nbr, V, ws.normals, ws.raw, st);
    if (rc) return rc;
    FMHR_STAGE_MARK();  // 1: vertex prep + normals
    ham_transform_kernel<<<dim3(cdiv(V, 256), cdiv(n, kViewsPerBlock)), 256, 0, st>>>(
        ws.vertices, b->w2cs, b->projs, b->view_idx, n, V, H, W, ws.pos, ws.snap, ws.scr, ws.viewM);
    FMHR_LAUNCH_CHECK();
    FMHR_STAGE_MARK();  // 2: transform
    rc = launch_raster_coverage_snapped(ws.pos, ws.snap, b->tri, n, V, T, H, W, zcur, ws.tbits[cur], ws.tlist[cur],
                                        ws.tcount[cur], tiles_x, tiles_x * tiles_y, st);
    if (rc) return rc;
    FMHR_STAGE_MARK();  // 3: coverage
    static const int g_shade = persistent_blocks(ham_shade_kernel<PHASE>);
    static const int g_aa = persistent_blocks(ham_aa_loss_kernel<PHASE>);
    static const int g_bwd = persistent_blocks(ham_pixel_bwd_kernel<PHASE>);
    const int pblock = kTile * kTile;  // 8 warps = 8 independent strip workers
    ham_shade_kernel<PHASE><<<g_shade, pblock, 0, st>>>(zcur, znext, ws.tlist[cur], ws.tcount[cur], ws.tlist[nxt],
                                                      ws.tcount[nxt], ws.abits, ws.alist, ws.acount, ws.cursors, tiles_x, tiles_y, ws.pos,
                                                      ws.scr, invW, invH, b->tri, b->opp, ws.normals, b->albedo,
                                                      b->masks, b->sh_coeffs, b->view_idx, sh_idx, V, T, H, W,
                                                      ws.plane[0], ws.plane[1], ws.acc, ws.vlist, ws.vcount);
    FMHR_LAUNCH_CHECK();
    FMHR_STAGE_MARK();  // 4: shade
    float4* g0 = PHASE == 0 ? ws.plane[2] : ws.plane[1];
    float4* g1 = PHASE == 0 ? ws.plane[3] : nullptr;
    ham_aa_loss_kernel<PHASE><<<g_aa, pblock, 0, st>>>(zcur, ws.alist, ws.acount, ws.cursors, tiles_x, tiles_y, ws.scr, b->tri, b->opp, b->imgs, b->valid_masks,
                                                     b->sh_coeffs, b->view_idx, sh_idx, V, T, H, W, ws.plane[0],
                                                     ws.plane[1], g0, g1, ws.acc, ws.gsh, b->view_vm2, dbg_image, dbg_mask,
                                                     ws.plist_a, ws.plist_b, ws.pcount, (int)(P / 2), ws.status);
    FMHR_LAUNCH_CHECK();
    FMHR_STAGE_MARK();  // 5: antialias + losses
    if (!forward_only) {
        ham_pair_bwd_kernel<PHASE><<<64, 256, 0, st>>>(ws.plist_a, ws.plist_b, ws.pcount, (int)(P / 2), zcur, ws.pos,
                                                       ws.viewM, V, H, W, ws.plane[0], g0, g1, ws.gdelta,
                                                       (float4*)b->packed);
        FMHR_LAUNCH_CHECK();
        ham_pixel_bwd_kernel<PHASE><<<g_bwd, pblock, 0, st>>>(ws.vlist, ws.vcount, zcur, ws.pos, invW, invH, ws.viewM, b->tri,
                                                              ws.normals, b->albedo, b->sh_coeffs, sh_idx, V, H, W, g0, g1,
                                                              ws.gdelta, (float4*)b->packed);
        FMHR_LAUNCH_CHECK();
    }
    ham_finalize_scalars_kernel<<<1, 32, 0, st>>>(ws.acc, b->view_vm2, b->view_idx, n, tiles_x * tiles_y, PHASE,
                                                  b->packed + 12 * (size_t)V);
    FMHR_LAUNCH_CHECK();
    FMHR_STAGE_MARK();  // 6: pixel backward (+ scalar finalize)
    return FMHR_OK;
}

extern "C" int fmhr_ham_reset(const fmhr_ham_config* cfg, const fmhr_ham_buffers* buf, fmhr_stream_t stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    FMHR_CHECK_ARG(buf && buf->workspace && buf->workspace_bytes >= fmhr_ham_workspace_bytes(cfg));
    HamWs ws;
    ham_layout(cfg, (char*)buf->workspace, &ws);
    const size_t P = (size_t)cfg->n_views * cfg->H * cfg->W;
    for (int i = 0; i < 2; i++) {
        FMHR_CUDA(cudaMemsetAsync(ws.zbuf[i], 0xFF, P * 8, (cudaStream_t)stream));
        FMHR_CUDA(cudaMemsetAsync(ws.slot_region[i], 0, ws.slot_bytes, (cudaStream_t)stream));
    }
    FMHR_CUDA(cudaMemsetAsync(ws.common_region, 0, ws.common_bytes, (cudaStream_t)stream));
    FMHR_CUDA(cudaMemsetAsync(ws.gdelta, 0, P * 16, (cudaStream_t)stream));
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

extern "C" int fmhr_ham_step_update(const fmhr_ham_config* cfg, const fmhr_ham_buffers* buf, fmhr_stream_t stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    rc = ham_check_buffers(cfg, buf);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    HamWs ws;
    ham_layout(cfg, (char*)buf->workspace, &ws);
    const int V = cfg->V;
    ham_update_pass1_kernel<<<cdiv((long long)V * kLPV, 256), 256, 0, st>>>(*cfg, ws.vertices, buf->delta, buf->albedo, buf->tri,
                                                          buf->v2f_ptr, (const int2*)buf->v2f_nbr, buf->v2v_ptr, buf->v2v_idx,
                                                          ws.raw, buf->packed, ws.yhat_v, ws.yhat_a, ws.gN, ws.acc,
                                                          buf->adam_step, ws.adam_sc);
    FMHR_LAUNCH_CHECK();
    ham_update_pass2_kernel<<<cdiv((long long)V * kLPV, 256), 256, 0, st>>>(*cfg, ws.vertices, buf->delta, buf->albedo, buf->tri,
                                                          buf->v2f_ptr, (const int2*)buf->v2f_nbr, buf->v2v_ptr, buf->v2v_idx,
                                                          buf->inv_deg, buf->packed, ws.yhat_v, ws.yhat_a, ws.gN, buf->adam_m,
                                                          buf->adam_v, ws.adam_sc, ws.acc, buf->losses, buf->dbg_grad, ws.status);
    FMHR_LAUNCH_CHECK();
    FMHR_STAGE_MARK();  // 7: update (regularisers, normal backward, Adam)
    if (cfg->phase == 0) {
        ham_update_sh_kernel<<<cdiv(cfg->n_sh_rows * 9, 128), 128, 0, st>>>(*cfg, ws.gsh, buf->packed, buf->sh_coeffs,
                                                                            buf->adam_m, buf->adam_v, ws.adam_sc,
                                                                            buf->dbg_grad_sh);
        FMHR_LAUNCH_CHECK();
    }
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

extern "C" int fmhr_ham_debug_export(const fmhr_ham_config* cfg, const fmhr_ham_buffers* buf, float* pos, float* rast,
                                     float* image, float* pred_mask, float* normals, fmhr_stream_t stream) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    rc = ham_check_buffers(cfg, buf);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    // re-run the forward with export pointers (leaves `packed` holding forward-only partials)
    rc = cfg->phase == 0 ? ham_render_impl<0>(cfg, buf, st, image, pred_mask, true)
                         : ham_render_impl<1>(cfg, buf, st, image, pred_mask, true);
    if (rc) return rc;
    HamWs ws;
    ham_layout(cfg, (char*)buf->workspace, &ws);
    const size_t nV = (size_t)cfg->n_views * cfg->V;
    if (pos) FMHR_CUDA(cudaMemcpyAsync(pos, ws.pos, nV * 16, cudaMemcpyDeviceToDevice, st));
    if (normals) FMHR_CUDA(cudaMemcpyAsync(normals, ws.normals, (size_t)cfg->V * 12, cudaMemcpyDeviceToDevice, st));
    if (rast) {
        const dim3 pgrid(cdiv((long long)cfg->H * cfg->W, 256), cfg->n_views);
        ham_export_rast_kernel<<<pgrid, 256, 0, st>>>(ws.zbuf[cfg->zbuf_slot], ws.pos, buf->tri, cfg->V, cfg->H, cfg->W,
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
