/* fmhr_b200 — C ABI of the B200-native HAM inverse-rendering hot path.
 *
 * Drop-in boundary for the path BASELINE.json's north_star names: the rasterize / interpolate /
 * antialias operators that /root/reference calls through `nvdiffrast.torch`
 * (mesh_sfs_optim.py:14,120,142-147,212-219,267-287; train_mlp.py:178,184; get_data.py:246-253)
 * plus the fused kernels of the HAM iteration itself (mesh_sfs_optim.py:198-237 phase A,
 * :253-310 phase B; models/utils.py:188-226,508-548,661-722).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *  - fp32 data, int32 indices, row-major / channel-last layouts exactly as the reference's tensors;
 *  - the caller allocates every output and every workspace (sizes from the *_workspace_bytes queries);
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*); no call synchronises the
 *    device or allocates memory, except the *_build setup calls, which say so;
 *  - return value 0 = success, otherwise a negative FMHR_E* code; the message is available from
 *    fmhr_last_error_string() (thread-local).  No C++ exception crosses this boundary.
 */
#ifndef FMHR_B200_H
#define FMHR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FMHR_OK 0
#define FMHR_EINVAL (-1)    /* bad argument (null pointer, non-positive size, too-small workspace) */
#define FMHR_ECUDA (-2)     /* a CUDA runtime call or kernel launch failed */
#define FMHR_EUNSUPPORTED (-3)

typedef void* fmhr_stream_t;

int fmhr_version(void);
const char* fmhr_last_error_string(void);

/* ---------------------------------------------------------------------------------------------
 * dr.rasterize(glctx, pos, tri, resolution)          [mesh_sfs_optim.py:142,212,267]
 *   pos  [N,V,4] clip-space positions (instanced mode), tri [T,3]
 *   rast [N,H,W,4] = (u, v, z/w, triangle_id+1), 0 where empty;  rast_db [N,H,W,4] or NULL
 * Deterministic coverage rule (DESIGN.md "Rasterisation rule"): 1/256-pixel snapping, 64-bit integer
 * edge functions, top-left style tie-break, 64-bit atomicMin z-buffer keyed (depth, triangle id).
 * workspace: N*H*W*8 bytes (the z-buffer).
 * ------------------------------------------------------------------------------------------- */
size_t fmhr_rasterize_workspace_bytes(int N, int H, int W);
int fmhr_rasterize_fwd(const float* pos, const int32_t* tri, int N, int V, int T, int H, int W,
                       float* rast, float* rast_db, void* workspace, size_t workspace_bytes, fmhr_stream_t stream);
/* Same result (bit-exact) on the fused path's meshlet coverage kernel: the meshlet-local vertices are snapped once into
 * shared memory instead of three times per triangle, a dirty-tile bitmap lets the resolve pass zero-fill untouched 16x16
 * tiles without reading the z-buffer, and the resolve pass puts the keys it consumed back to "empty" - so when
 * `zbuf_is_clean` (every byte of zbuf_ws 0xFF on entry: freshly filled, or left by a previous successful call of this
 * function with any N/H/W) no clear pass runs at all.  ml_* from fmhr_meshlets_build_host (any vertex positions, e.g.
 * one view's NDC, or NULL for the input order; tris_per_meshlet 256 / 512 / 1024), uploaded to the device; every index
 * of `tri` must be in [0,V).  tile_bits = fmhr_rasterize_tile_words(N,H,W) uint32 of scratch. */
size_t fmhr_rasterize_tile_words(int N, int H, int W);
int fmhr_rasterize_fwd_meshlets(const float* pos, const int32_t* tri, int N, int V, int T, int H, int W, float* rast,
                                float* rast_db, const int32_t* ml_vptr, const int32_t* ml_verts, const uint32_t* ml_tri2,
                                int n_meshlets, int ml_tris, int ml_max_verts, void* zbuf_ws, size_t zbuf_bytes,
                                int zbuf_is_clean, uint32_t* tile_bits, fmhr_stream_t stream);
/* grad_pos [N,V,4] is overwritten with d(sum dy.rast)/d(pos); only the u,v channels of dy carry gradient. */
int fmhr_rasterize_bwd(const float* pos, const int32_t* tri, const float* rast, const float* dy,
                       int N, int V, int T, int H, int W, float* grad_pos, fmhr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * dr.interpolate(attr, rast, tri)                    [mesh_sfs_optim.py:143,214,269; train_mlp.py:184]
 *   attr [NA,V,A] with NA == N, or NA == 1 to broadcast one attribute set to all N images
 *   out  [N,H,W,A] = u*a0 + v*a1 + (1-u-v)*a2, 0 where empty
 * bwd overwrites grad_attr [NA,V,A] and grad_rast [N,H,W,4] (channels 2,3 are zero).
 * ------------------------------------------------------------------------------------------- */
int fmhr_interpolate_fwd(const float* attr, const float* rast, const int32_t* tri, int N, int NA, int V, int T,
                         int H, int W, int A, float* out, fmhr_stream_t stream);
int fmhr_interpolate_bwd(const float* attr, const float* rast, const int32_t* tri, const float* dy, int N, int NA,
                         int V, int T, int H, int W, int A, float* grad_attr, float* grad_rast, fmhr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Mesh topology (setup; once per `tri` tensor).  Replaces nvdiffrast's per-call topology hash
 * (antialias(..., topology_hash=None)) and the per-call sparse-Laplacian rebuild of
 * models/utils.py:551-571,661-693.
 *   opp      [T,3]  vertex opposite to the edge facing corner k in the lowest-indexed other triangle on
 *                   that edge, -1 on a boundary
 *   v2f_ptr  [V+1], v2f_idx [3T]   vertex -> incident (triangle*4 + corner), ascending
 *   v2v_ptr  [V+1], v2v_idx [6T]   vertex -> unique neighbour vertices, ascending (first v2v_ptr[V] used)
 * This call synchronises `stream` once (it returns the directed-edge count through n_dir_edges_host).
 * Any of the three groups may be NULL to skip it.
 * ------------------------------------------------------------------------------------------- */
size_t fmhr_mesh_topology_workspace_bytes(int V, int T);
int fmhr_mesh_topology_build(const int32_t* tri, int V, int T, int32_t* opp, int32_t* v2f_ptr, int32_t* v2f_idx,
                             int32_t* v2v_ptr, int32_t* v2v_idx, int* n_dir_edges_host, void* workspace,
                             size_t workspace_bytes, fmhr_stream_t stream);

/* Derived adjacency for the fused iteration (setup, after fmhr_mesh_topology_build):
 *   v2f_nbr [3T,2] the two other corners (cyclic order) of each vertex->face entry, inv_deg [V] = 1/degree. */
int fmhr_mesh_topology_derive(const int32_t* tri, const int32_t* v2f_idx, const int32_t* v2v_ptr, int V, int T,
                              int32_t* v2f_nbr, float* inv_deg, fmhr_stream_t stream);

/* Meshlets for the fused iteration's coverage kernel (setup, HOST pointers): faces are ordered along a Morton curve of
 * their centroids (verts [V,3]; NULL keeps the input order) and cut into groups of at most tris_per_meshlet (multiple of
 * 32, <= 1024) triangles and 1024 distinct vertices.  Call once with ml_vptr = ml_verts = ml_tri2 = NULL to obtain
 * the counts, allocate ml_vptr [*n_meshlets+1], ml_verts [*n_vert_refs], ml_tri2 [*n_meshlets*tris_per_meshlet*2] and
 * call again.  The z-buffer keeps ORIGINAL triangle ids, so results do not depend on the meshlet order. */
int fmhr_meshlets_build_host(const int32_t* tri, const float* verts, int V, int T, int tris_per_meshlet,
                             int* n_meshlets, int* n_vert_refs, int* max_verts, int32_t* ml_vptr, int32_t* ml_verts,
                             uint32_t* ml_tri2);

/* ---------------------------------------------------------------------------------------------
 * dr.antialias(color, rast, pos, tri)                [mesh_sfs_optim.py:146-147,217-219,274,287]
 *   color [N,H,W,C], rast [N,H,W,4], pos [N,V,4], tri [T,3], opp [T,3] from fmhr_mesh_topology_build
 * bwd overwrites grad_color [N,H,W,C] and grad_pos [N,V,4] (grad_pos may be NULL).
 * ------------------------------------------------------------------------------------------- */
int fmhr_antialias_fwd(const float* color, const float* rast, const float* pos, const int32_t* tri,
                       const int32_t* opp, int N, int H, int W, int C, int V, int T, float* out, fmhr_stream_t stream);
int fmhr_antialias_bwd(const float* color, const float* rast, const float* pos, const int32_t* tri,
                       const int32_t* opp, const float* dy, int N, int H, int W, int C, int V, int T,
                       float* grad_color, float* grad_pos, fmhr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * models/utils.py operators used by the loop, stand-alone (autograd.Function building blocks).
 * ------------------------------------------------------------------------------------------- */
/* get_normals (models/utils.py:508-548): normals [V,3] = normalize(sum of incident face cross products, eps 1e-6);
 * raw [V,3] receives the un-normalised sums (needed by the backward).  One copy, not n. */
int fmhr_vertex_normals_fwd(const float* verts, const int32_t* tri, const int32_t* v2f_ptr, const int32_t* v2f_idx,
                            int V, int T, float* normals, float* raw, fmhr_stream_t stream);
/* grad_verts [V,3] overwritten; scratch [V,3] floats. */
int fmhr_vertex_normals_bwd(const float* verts, const int32_t* tri, const int32_t* v2f_ptr, const int32_t* v2f_idx,
                            const float* raw, const float* grad_normals, int V, int T, float* scratch,
                            float* grad_verts, fmhr_stream_t stream);
/* laplacian_smoothing(x, faces, "uniform") (models/utils.py:696-722): *loss = sum_i ||(Lx)_i|| / V.
 * yhat [V,C] receives (Lx)_i/||(Lx)_i|| (0 where the norm is 0) for the backward; C in {1..4}. */
int fmhr_laplacian_fwd(const float* x, const int32_t* v2v_ptr, const int32_t* v2v_idx, int V, int C, float* yhat,
                       float* loss, fmhr_stream_t stream);
/* grad_x [V,C] overwritten with scale * L^T yhat / V. */
int fmhr_laplacian_bwd(const float* yhat, const int32_t* v2v_ptr, const int32_t* v2v_idx, int V, int C, float scale,
                       float* grad_x, fmhr_stream_t stream);
/* get_radiance (models/utils.py:208-226), degree 3: radiance[i] = coeff[i or 0] . basis(normal[i]). */
int fmhr_sh_radiance_fwd(const float* coeff, int coeff_rows, const float* normal, int n, float* radiance,
                         fmhr_stream_t stream);
int fmhr_sh_radiance_bwd(const float* coeff, int coeff_rows, const float* normal, const float* grad_radiance, int n,
                         float* grad_coeff, float* grad_normal, fmhr_stream_t stream);
/* NCC (models/ncc_utils.py:4-35): ref [1,Np,Npx], src/mask [Nv,Np,Npx] -> ncc [Nv,Np]. */
int fmhr_ncc_fwd(const float* ref, const float* src, const float* src_mask, int Nv, int Np, int Npx, float* ncc,
                 fmhr_stream_t stream);
/* grad_src [Nv,Np,Npx] = d(sum grad_ncc * ncc)/d(src) (overwritten). */
int fmhr_ncc_bwd(const float* ref, const float* src, const float* src_mask, const float* grad_ncc, int Nv, int Np,
                 int Npx, float* grad_src, fmhr_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Fused HAM iteration (mesh_sfs_optim.py:198-237 phase A, :253-310 phase B).
 * The context owns nothing: every buffer below is caller memory; `workspace` holds the z-buffer, the
 * shaded-colour and pixel-gradient planes and the per-vertex gradient accumulators.
 * ------------------------------------------------------------------------------------------- */
typedef struct fmhr_ham_config {
    int32_t V, T, H, W;
    int32_t n_views;          /* views in this step's batch on THIS rank */
    int32_t n_views_global;   /* views in the step's batch over all ranks (mask-loss denominator) */
    int32_t phase;            /* 0 = phase A (albedo + SH), 1 = phase B (delta + albedo) */
    int32_t n_sh_rows;        /* rows of sh_coeffs resident on this rank (= its number of views) */
    int32_t zbuf_slot;        /* 0/1: z-buffer rasterised by THIS step; the step resets the other one, so the caller
                                 alternates the slot every step (after fmhr_ham_reset) */
    int32_t view_groups;      /* phase-B step: number of consecutive view groups whose pixel passes overlap the next
                                 group's coverage kernel (1..4; 0 = library default, env FMHR_VIEW_GROUPS) */
    float sfs_weight, lap_weight, albedo_weight, mask_weight, edge_weight, delta_weight;
    float lr, albedo_lr, sh_lr;
    float beta1, beta2, eps;
    float edge_length_mean;   /* mesh_sfs_optim.py:188 */
    int32_t n_views_capacity; /* 0, or >= n_views: the workspace is laid out for this many views, so that steps with
                                 DIFFERENT batch sizes (the reference's last short batch of an epoch, mesh_sfs_optim.py:252-254)
                                 share one layout and need no fmhr_ham_reset in between; 0 = laid out for n_views */
} fmhr_ham_config;

typedef struct fmhr_ham_buffers {
    /* mesh + topology (static) */
    const int32_t* tri;       /* [T,3] */
    const int32_t* opp;       /* [T,3] */
    const int32_t* v2f_ptr;   /* [V+1] */
    const int32_t* v2f_idx;   /* [3T] */
    const int32_t* v2v_ptr;   /* [V+1] */
    const int32_t* v2v_idx;   /* [2E] */
    const int32_t* v2f_nbr;   /* [3T,2] from fmhr_mesh_topology_derive */
    const float* inv_deg;     /* [V]    from fmhr_mesh_topology_derive */
    /* optimisation state */
    const float* vertices_tmp; /* [V,3] */
    float* delta;              /* [V,3] */
    float* albedo;             /* [V,3] */
    float* sh_coeffs;          /* [n_sh_rows,9] */
    float* adam_m;             /* [6V + 9*n_sh_rows]  (delta | albedo | sh) */
    float* adam_v;             /* same layout */
    int32_t* adam_step;        /* [4] step counters (delta, albedo, sh) + [3] latched fatal flag (fmhr_ham_step_update_peer) */
    /* per-view data of all views resident on this rank */
    const float* imgs;         /* [num,H,W,3] */
    const float* masks;        /* [num,H,W] */
    const float* valid_masks;  /* [num,H,W] */
    const double* view_vm2;    /* [num, tiles+1] from fmhr_ham_prepare_views, tiles = ceil(W/16)*ceil(H/16) */
    const float* w2cs;         /* [num,4,4] transposed (row-vector) */
    const float* projs;        /* [num,4,4] transposed */
    const int32_t* view_idx;   /* [n_views] rows of the per-view arrays used by this step (the perm slice) */
    const int32_t* sh_idx;     /* [n_views] rows of sh_coeffs for those views; NULL = same as view_idx */
    /* packed reduction buffer: [12V gradient accumulators | 16 scalars]; all-reduced (sum) across ranks
     * between fmhr_ham_step_render and fmhr_ham_step_update when world size > 1 */
    float* packed;
    /* loss record written by fmhr_ham_step_update: sfs, lap, albedo, mask, edge, delta, n_valid, total */
    float* losses;             /* [8] */
    void* workspace;
    size_t workspace_bytes;
    /* optional inspection outputs for parity tests (NULL in production): the gradients Adam consumed */
    float* dbg_grad;           /* [V,6] = (d loss/d delta xyz, d loss/d albedo bgr) */
    float* dbg_grad_sh;        /* [n_sh_rows,9] (phase A) */
    /* meshlets of the coverage kernel (static, from fmhr_meshlets_build_host, uploaded to the device) */
    const int32_t* ml_vptr;    /* [n_meshlets+1] */
    const int32_t* ml_verts;   /* [ml_vptr[n_meshlets]] global vertex id of each meshlet-local vertex */
    const uint32_t* ml_tri2;   /* [n_meshlets*ml_tris,2] (l0 | l1<<10 | l2<<20, original triangle id); padding = ~0 */
    int32_t n_meshlets;
    int32_t ml_tris;           /* triangles per meshlet (multiple of 256, <= 1024) */
    int32_t ml_max_verts;      /* largest meshlet vertex count (<= 1024) */
    int32_t ml_reserved;
    /* caller-owned scratch [n_meshlets * ml_max_verts] float4: the step's world-space vertices in MESHLET order (written by
     * the prologue kernel, padded per meshlet), so a coverage block fetches its vertices with independent coalesced loads
     * instead of the vptr -> index -> record chain */
    float* ml_pos;
} fmhr_ham_buffers;

size_t fmhr_ham_workspace_bytes(const fmhr_ham_config* cfg);
size_t fmhr_ham_packed_floats(const fmhr_ham_config* cfg);
/* Resets both z-buffer slots of `workspace`.  Call once before the first step and whenever the workspace layout
 * changes (different n_views / phase / workspace pointer). */
int fmhr_ham_reset(const fmhr_ham_config* cfg, const fmhr_ham_buffers* buf, fmhr_stream_t stream);
/* Diagnostic builds (-DFMHR_CHECKED): bit mask of the index assertions that fired inside the kernels since the last call
 * (synchronises the device; 0 = clean).  Product builds compile the assertions out and return -1. */
int fmhr_debug_checks(void);
/* Constants of the mask loss: view_vm2[i][t] = sum of valid_masks[i]^2 over 16x16 tile t (row-major tiles), and
 * view_vm2[i][tiles] = the view total, tiles = ceil(W/16)*ceil(H/16) (valid_masks are fixed during the optimisation,
 * mesh_sfs_optim.py:163).  Call once at setup, and again if valid_masks change. */
int fmhr_ham_prepare_views(const float* valid_masks, int num, int H, int W, double* view_vm2, fmhr_stream_t stream);
/* forward + pixel backward: fills `packed` with un-normalised gradient accumulators and loss partials. */
int fmhr_ham_step_render(const fmhr_ham_config* cfg, const fmhr_ham_buffers* buf, fmhr_stream_t stream);
/* normalises with the (all-reduced) counts, adds the regulariser gradients, applies Adam, writes `losses`. */
int fmhr_ham_step_update(const fmhr_ham_config* cfg, const fmhr_ham_buffers* buf, fmhr_stream_t stream);
/* Extra loss terms of a phase-B step (the NCC term below): adds a fully weighted gradient w.r.t. delta [V,3], computed by
 * the caller between fmhr_ham_step_render and fmhr_ham_step_update, to the gradient the update applies. */
int fmhr_ham_add_delta_grad(const fmhr_ham_config* cfg, const fmhr_ham_buffers* buf, const float* grad_delta,
                            fmhr_stream_t stream);
/* NCC photo-consistency term (BASELINE.json configs[2]; arithmetic of models/ncc_utils.py:4-35 = fmhr_ncc_fwd / _bwd above;
 * the wiring has no counterpart in the reference, SURVEY.md F4, and is specified in DESIGN.md): Np points fixed on the
 * surface (pt_face [Np], pt_bary [Np,2]; third weight = 1 - b0 - b1) are projected into the views view_idx[0] (reference)
 * and view_idx[1..Nv1) (sources) of the camera arrays w2cs / projs (transposed, as everywhere), and a (2 half + 1)^2 patch
 * of the gray image [num,H,W] is sampled bilinearly around each projection (0 outside the frame): patches [Nv1,Np,Npx];
 * patch_mask = all four taps inside the frame and masks[view] > 0.5 at the nearest pixel.  The backward takes
 * d loss / d patches (row 0, the reference view, is ignored: the reference patch is a constant of the step) and ADDS
 * d loss / d vertices [V,3]. */
int fmhr_ncc_sample_fwd(const float* vertices, const int32_t* tri, const int32_t* pt_face, const float* pt_bary,
                        const float* w2cs, const float* projs, const int32_t* view_idx, int Nv1, const float* gray,
                        const float* masks, int V, int Np, int H, int W, int half, float* patches, float* patch_mask,
                        fmhr_stream_t stream);
int fmhr_ncc_sample_bwd(const float* vertices, const int32_t* tri, const int32_t* pt_face, const float* pt_bary,
                        const float* w2cs, const float* projs, const int32_t* view_idx, int Nv1, const float* gray, int V,
                        int Np, int H, int W, int half, const float* grad_patches, float* grad_vertices,
                        fmhr_stream_t stream);
/* The whole term in one kernel for patches of at most 128 samples (half <= 5): sampling, masked NCC (models/ncc_utils.py:
 * 4-35) and its gradient through the source sampling positions to the vertices, nothing of size [Nv,Np,Npx] materialised.
 * grad_ncc = d loss / d ncc, the same for every (source view, point); ncc [Nv1-1,Np] receives the values;
 * grad_vertices [V,3] is accumulated into (+=).  FMHR_EUNSUPPORTED for larger patches (use the four calls above). */
int fmhr_ncc_term_fused(const float* vertices, const int32_t* tri, const int32_t* pt_face, const float* pt_bary,
                        const float* w2cs, const float* projs, const int32_t* view_idx, int Nv1, const float* gray,
                        const float* masks, int V, int Np, int H, int W, int half, float grad_ncc, float* ncc,
                        float* grad_vertices, fmhr_stream_t stream);
/* ---- multi-GPU exchange over NVLink peer memory (replaces the ncclAllReduce of `packed` the reference
 * design would place between loss.backward() and optimizer.step(), mesh_sfs_optim.py:309-310, when views shard) ----
 * Every rank allocates its exchange memory with fmhr_peer_alloc (cudaMalloc + cudaIpcGetMemHandle; zero-filled), the
 * host exchanges the 64-byte handles out of band (torch.distributed) and maps the peers' allocations with
 * fmhr_peer_open.  fmhr_ham_step_update_peer is fmhr_ham_step_update with the all-reduce fused into its first kernel:
 * it posts this rank's step count into every peer's flag array, waits for all peers, sums the ranks' `packed` buffers
 * in rank order over NVLink (bit-identical on every rank) into `reduced` and carries on with the update.  Up to two
 * ranks every rank gathers all accumulators itself (one kernel); beyond that rank k sums vertex chunk k and stores it
 * into every rank's `reduced`, and a second rendezvous opens the normal-gradient kernel (two kernels).  Contract:
 *  - packed[r] = rank r's packed buffer of THIS step as mapped on this device, packed[rank] == buf->packed;
 *  - consecutive steps must alternate between two packed buffers per rank (a peer may still read the previous one);
 *  - flags[r]  = rank r's flag array, 2 * FMHR_MAX_PEERS uint32 words, zero before the first step;
 *  - reduced[r] = rank r's sum buffer (one per rank is enough: a peer can only store into it again after this rank
 *    has posted the next step);
 *  - epoch     = one zero-initialised uint32 in local device memory (steps completed; advanced by the call);
 *  - every rank makes the same sequence of calls.  A peer that does not arrive within timeout_s does not hang the
 *    device: the update is REFUSED (delta / albedo / sh_coeffs, Adam moments and step counters stay untouched), the
 *    whole loss record becomes NaN and adam_step[3] is latched to 1 so that every later update refuses as well - the
 *    host must treat adam_step[3] != 0 as fatal (HamOptimizer.check_health raises). */
#define FMHR_MAX_PEERS 16
typedef struct fmhr_ham_peers {
    int32_t rank, world;
    int32_t mode;              /* 0 = by world size, 1 = one-shot gather, 2 = two-shot (reduce-scatter + peer stores) */
    int32_t timeout_s;         /* rendezvous give-up time in seconds (0 = default, 30 s) */
    const float* packed[FMHR_MAX_PEERS];
    uint32_t* flags[FMHR_MAX_PEERS];
    float* reduced[FMHR_MAX_PEERS]; /* every rank's [fmhr_ham_packed_floats] sum buffer (in its shared allocation) */
    uint32_t* epoch;
} fmhr_ham_peers;
int fmhr_peer_alloc(size_t bytes, void** dev_ptr, void* handle64);
int fmhr_peer_open(const void* handle64, void** dev_ptr);
int fmhr_peer_close(void* dev_ptr);
int fmhr_peer_free(void* dev_ptr);
int fmhr_ham_step_update_peer(const fmhr_ham_config* cfg, const fmhr_ham_buffers* buf, const fmhr_ham_peers* peers,
                              fmhr_stream_t stream);
/* Inspection for parity tests: copies internal planes of the last fmhr_ham_step_render into caller buffers
 * (any may be NULL): pos [n,V,4], rast [n,H,W,4], image [n,H,W,3] (antialiased), pred_mask [n,H,W],
 * normals [V,3]. */
int fmhr_ham_debug_export(const fmhr_ham_config* cfg, const fmhr_ham_buffers* buf, float* pos, float* rast,
                          float* image, float* pred_mask, float* normals, fmhr_stream_t stream);
/* Measurement aid: runs ONE render+update with CUDA events between the launches and returns the device time of each
 * stage in milliseconds (ms_host[16]): 0 clears, 1 vertex prep + normals, 2 clip transform, 3 coverage, 4 shade,
 * 5 antialias + losses, 6 pixel backward, 7 update/Adam.  Synchronises the stream; advances the optimiser one step. */
int fmhr_ham_stage_times(const fmhr_ham_config* cfg, const fmhr_ham_buffers* buf, float* ms_host, int* n_stages_host,
                         fmhr_stream_t stream);
/* Diagnostic builds only (-DFMHR_TRACE, tools/trace_timeline.py): per-kernel (first block entry, last warp exit)
 * %globaltimer stamps in ns since the last reset, stamps_host[2 * slot + {0,1}]; slot numbering in csrc/ham.cu.
 * Synchronises the device.  The product build returns FMHR_EUNSUPPORTED. */
int fmhr_trace_read(unsigned long long* stamps_host, int n_slots, int reset);
/* Diagnostic builds: enqueue on `stream` "store the stamps collected since the previous mark as ring entry `rep` (0..63,
 * negative = discard) and re-arm the slots"; fmhr_trace_read(out, -reps, 0) then returns reps x 64 x 2 stamps - the
 * schedule of `reps` consecutive free-running steps without a host synchronisation between them. */
int fmhr_trace_mark(int rep, fmhr_stream_t stream);
/* Host-buffer variant (end-to-end measurement path): copies this step's view batch (img/mask/valid_mask/w2c/proj
 * rows, all HOST pinned pointers, n_views rows each) into the device staging planes named by `buf`, runs
 * render+update, and copies the 8-float loss record back to losses_host. */
int fmhr_ham_step_host(const fmhr_ham_config* cfg, const fmhr_ham_buffers* buf, const float* imgs_host,
                       const float* masks_host, const float* valid_masks_host, const float* w2cs_host,
                       const float* projs_host, float* losses_host, fmhr_stream_t stream);

/* Same, with the batch in its native 8-bit form (what the reference's loader reads from disk, get_data.py:77-90):
 * imgs_host [n,H,W,3] uint8 (BGR, img = u8 / 255), masks_host [n,H,W] uint8 (mask = u8 > 127), pinned HOST memory;
 * `staging` = fmhr_ham_host_u8_staging_bytes(cfg) bytes of DEVICE memory.  The bytes travel on an internal copy stream
 * while vertex prep / coverage / scan run, are converted into buf->imgs / buf->masks right before the shade pass, and
 * the 8-float loss record is copied back to losses_host.  valid_masks / view_vm2 stay resident (the reference derives
 * valid_masks on the device, mesh_sfs_optim.py:146-163).  Requires n*H*W % 4 == 0. */
size_t fmhr_ham_host_u8_staging_bytes(const fmhr_ham_config* cfg);
int fmhr_ham_step_host_u8(const fmhr_ham_config* cfg, const fmhr_ham_buffers* buf, const uint8_t* imgs_host,
                          const uint8_t* masks_host, const float* w2cs_host, const float* projs_host, void* staging,
                          float* losses_host, fmhr_stream_t stream);

/* Pipelined form of the same step (a loader thread decodes batch i+1 while step i runs): fmhr_ham_host_u8_submit starts
 * the upload of a batch into `staging` on the internal copy stream and returns at once; up to two staging buffers may be
 * in flight per device, and a buffer is only overwritten after the conversion kernel of its previous batch.
 * fmhr_ham_step_host_u8_submitted runs the iteration on the batch last submitted into `staging`; `peers` = NULL on a single
 * rank, else the update is fmhr_ham_step_update_peer (buf->packed must be peers->packed[rank], see below). */
int fmhr_ham_host_u8_submit(const fmhr_ham_config* cfg, const uint8_t* imgs_host, const uint8_t* masks_host,
                            void* staging);
struct fmhr_ham_peers;
/* fmhr_ham_step_host_u8_submitted == acquire + body + release.  A host that replays the device work from a CUDA graph
 * captures `body` once per (staging buffer, z-buffer slot) and brackets every replay with acquire (orders `stream` behind
 * the upload of the batch) and release (lets the copy stream refill the staging buffer once the step has consumed it). */
int fmhr_ham_step_host_u8_acquire(void* staging, fmhr_stream_t stream);
int fmhr_ham_step_host_u8_body(const fmhr_ham_config* cfg, const fmhr_ham_buffers* buf, const float* w2cs_host,
                               const float* projs_host, const void* staging, float* losses_host,
                               const struct fmhr_ham_peers* peers, fmhr_stream_t stream);
int fmhr_ham_step_host_u8_release(void* staging, fmhr_stream_t stream);
/* fmhr_ham_host_u8_submit that only moves what can matter: boxes_host [n_views,4] int32 = (y0, y1, x0, x1), half-open, per
 * view SLOT a rectangle containing every pixel whose mask byte is > 127 (loader metadata: the bounding box of the
 * segmentation).  Outside it mask = 0, so no pixel is valid (mesh_sfs_optim.py:276-281) and no image byte is read; the mask
 * staging plane is zero-filled on the device and the rectangles travel as row-pitched 2-D copies.  *h2d_bytes (optional)
 * receives the bytes queued (4 per pixel of the rectangles). */
int fmhr_ham_host_u8_submit_boxes(const fmhr_ham_config* cfg, const uint8_t* imgs_host, const uint8_t* masks_host,
                                  const int32_t* boxes_host, void* staging, size_t* h2d_bytes);
int fmhr_ham_step_host_u8_submitted(const fmhr_ham_config* cfg, const fmhr_ham_buffers* buf, const float* w2cs_host,
                                    const float* projs_host, void* staging, float* losses_host,
                                    const struct fmhr_ham_peers* peers, fmhr_stream_t stream);
/* Converted-on-arrival form of fmhr_ham_host_u8_submit_boxes: the kernel that pulls the box rows out of the mapped pinned
 * host buffers writes them straight into a set of FLOAT planes (img = u8 / 255, mask = u8 > 127; imgs_dev [n,H,W,3],
 * masks_dev [n,H,W], zero-filled outside the boxes first) and the step's cameras (w2cs_host / projs_host [n,4,4], pinned)
 * travel on the same copy stream into w2cs_dev / projs_dev - nothing of the batch is left to upload or convert inside the
 * step.  The caller keeps two such plane sets and alternates them; `imgs_dev` is the handle the slot protocol knows the set
 * by: fmhr_ham_step_host_u8_acquire(imgs_dev) / _release(imgs_dev) bracket the step, and fmhr_ham_step_host_u8_body is
 * called with buf->imgs / masks / w2cs / projs pointing at the set, staging = imgs_dev and w2cs_host = projs_host = NULL.
 * The library keeps masks_dev zero outside the boxes: a batch submitted with fresh != 0 (the set's content is undefined: newly
 * allocated, or written by the caller) clears the plane, every later one zeroes the rows of the batch before it.
 * Needs device-mapped, 16-byte aligned host buffers and H*W % 16 == 0 (FMHR_EUNSUPPORTED otherwise: use the staging
 * form).  *h2d_bytes (optional) receives the bytes pulled (4 per pixel of the rectangles + the cameras). */
int fmhr_ham_host_u8_submit_boxes_direct(const fmhr_ham_config* cfg, const uint8_t* imgs_host, const uint8_t* masks_host,
                                         const int32_t* boxes_host, const float* w2cs_host, const float* projs_host,
                                         float* imgs_dev, float* masks_dev, float* w2cs_dev, float* projs_dev,
                                         int fresh, size_t* h2d_bytes);

/* HAM initialisation (mesh_sfs_optim.py:124-177) on the same fused forward chain: every view of the batch is rendered
 * once (n_views rows of view_idx, normally all views), normals and coverage are antialiased, and
 *   valid_masks_out [n_views,H,W] = antialiased coverage of the initial mesh (:146,163; indexed by view SLOT),
 *   sh_coeffs_out   [n_views,9]   = per-view least-squares SH lighting of the re-normalised normals onto grayimgs (:152-153),
 *   sh_global_out   [9]           = the same fit over all views (:165-166),
 *   albedo_mean_out [3]           = mean over the valid pixels of img / radiance(global SH) (:173-174).
 * grayimgs [num,H,W] (rows addressed through view_idx like imgs).  cfg->phase must be 0 (workspace layout), `scratch` =
 * fmhr_ham_init_scratch_bytes(n_views) bytes of device memory; buf->valid_masks / view_vm2 / sh_coeffs are not read.
 * The normal equations are accumulated and solved in fp64 (9x9 Cholesky) instead of the reference's SVD least squares. */
size_t fmhr_ham_init_scratch_bytes(int num);
int fmhr_ham_init(const fmhr_ham_config* cfg, const fmhr_ham_buffers* buf, const float* grayimgs, float* valid_masks_out,
                  float* sh_coeffs_out, float* sh_global_out, float* albedo_mean_out, void* scratch, fmhr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FMHR_B200_H */
