"""Native HAM optimiser: the reference's two loops (mesh_sfs_optim.py:193-240 phase A, :242-317 phase B) with every
iteration executed by two C-ABI calls (fmhr_ham_step_render / fmhr_ham_step_update) and, on more than one GPU, one
exchange of the packed gradient buffer (views shard across ranks): fused into the update over NVLink peer
memory (fmhr_ham_step_update_peer), or one NCCL all-reduce between the two calls.

State names follow the reference: vertices_tmp, delta, albedo [1,V,3], sh_coeffs [num,9], valid_masks, ...
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import HamBuffers, HamConfig, check, ptr, stream
from .dist import PeerExchange, allreduce_packed
from .dr import Topology


_IDX_COPY_ALWAYS = os.environ.get("FMHR_IDX_COPY_ALWAYS", "0") == "1"  # A/B measurements of the index-copy elision


class HamOptimizer:
    def __init__(self, vertices, faces, imgs, masks, valid_masks, w2cs, projs, sh_coeffs, albedo, conf,
                 process_group=None, n_views_global=None, debug=False, use_graphs=False, exchange=None,
                 view_groups=None):
        """All tensors CUDA.  vertices [V,3], faces [F,3] int32, imgs [num,H,W,3], masks/valid_masks [num,H,W],
        w2cs/projs [num,4,4] (transposed, get_data.py:96-97), sh_coeffs [num,9], albedo [1,V,3] or [V,3];
        conf: dict with the weights and learning rates of conf/*.conf."""
        self.lib = _lib.load()
        _lib.require_cuda(vertices, faces, imgs, masks, valid_masks, w2cs, projs, sh_coeffs, albedo)
        dev = vertices.device
        self.device = dev
        f32 = lambda t: t.detach().to(torch.float32).contiguous()
        self.vertices_tmp = f32(vertices).clone()
        self.faces = faces.detach().to(torch.int32).contiguous()
        self.V, self.T = self.vertices_tmp.shape[0], self.faces.shape[0]
        self.imgs, self.masks, self.valid_masks = f32(imgs), f32(masks), f32(valid_masks)
        self.w2cs, self.projs = f32(w2cs), f32(projs)
        self.num, self.H, self.W = self.imgs.shape[0], self.imgs.shape[1], self.imgs.shape[2]
        self.sh_coeffs = f32(sh_coeffs).clone()
        self.albedo = f32(albedo).reshape(self.V, 3).clone()
        self.delta = torch.zeros_like(self.vertices_tmp)
        self.conf = dict(conf)
        self.topo = Topology(self.faces, self.V)
        self._build_meshlets()
        # mesh_sfs_optim.py:184-188: mean of squared lengths of the 3F half-edges of the initial mesh
        v, f = self.vertices_tmp, self.faces.long()
        a, b, c = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
        self.edge_length_mean = float(torch.cat([((a - b) ** 2).sum(1), ((c - b) ** 2).sum(1), ((a - c) ** 2).sum(1)]).mean())
        n_state = 6 * self.V + 9 * self.num
        self.adam_m = torch.zeros(n_state, dtype=torch.float32, device=dev)
        self.adam_v = torch.zeros(n_state, dtype=torch.float32, device=dev)
        self.adam_step = torch.zeros(4, dtype=torch.int32, device=dev)
        self.packed = torch.zeros(12 * self.V + 4, dtype=torch.float32, device=dev)
        self.losses = torch.zeros(8, dtype=torch.float32, device=dev)
        self.n_tiles = ((self.W + 15) // 16) * ((self.H + 15) // 16)
        self.view_vm2 = torch.zeros(self.num, self.n_tiles + 1, dtype=torch.float64, device=dev)
        check(self.lib.fmhr_ham_prepare_views(ptr(self.valid_masks), self.num, self.H, self.W, ptr(self.view_vm2), stream()),
              "ham_prepare_views")
        self.workspace = None
        self.pg = process_group
        self.world = 1  # process_group=False forces a single-rank optimiser inside a distributed job
        if process_group is not False and torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size(process_group)
        # The one exchange of the path (sum of `packed` over the ranks): by default fused into the update's first kernel
        # over NVLink peer memory (fmhr_ham_step_update_peer); exchange="nccl" (or FMHR_EXCHANGE=nccl, or peers that cannot
        # be mapped) keeps one NCCL all-reduce between the two halves of the step.
        self.peer = None
        forced = exchange is not None and exchange.startswith("peer")  # explicit: also on a single rank (tests)
        exchange = exchange or os.environ.get("FMHR_EXCHANGE", "peer")
        if exchange not in ("peer", "peer-oneshot", "peer-twoshot", "nccl"):
            raise RuntimeError("fmhr_b200: exchange must be 'peer', 'peer-oneshot', 'peer-twoshot' or 'nccl'")
        if exchange.startswith("peer") and (self.world > 1 or forced):
            px = PeerExchange(12 * self.V + 4, dev, process_group, mode=exchange[5:] or None)
            if forced and not px.ok:
                raise RuntimeError("fmhr_b200: peer exchange requested but the ranks' buffers could not be mapped")
            self.peer = px if px.ok else None
        self.n_views_global_override = n_views_global
        self.view_groups = int(view_groups or 0)  # 0 = library default (fmhr_ham_config.view_groups)
        # The loss is a mean over the GLOBAL batch, so every rank must normalise by the same n_views_global.  The default
        # (local batch x world) is only right when every rank holds the same number of views in every step: checked once
        # here (collective); uneven shards (num_views % world != 0, short last batches) must pass n_views_global.
        self._uneven_shards = False
        if self.world > 1:
            nums = [None] * self.world
            torch.distributed.all_gather_object(nums, int(self.num), group=process_group)
            self._uneven_shards = len(set(nums)) > 1
        self.dbg_grad = torch.zeros(self.V, 6, dtype=torch.float32, device=dev) if debug else None
        self.dbg_grad_sh = None
        self.debug = debug
        self.phase = 0
        self.extra_terms = []  # phase-B loss terms outside the fused passes (fmhr_b200.ncc_term.NccTerm)
        self._view_idx_cache = {}
        self.use_graphs = bool(use_graphs) and not debug
        self._graphs = {}
        self._idx_tags = {}
        self._struct_cache = {}
        self._zb_layout = None
        self._zb_slot = 0
        self.batch_capacity = 0  # see set_batch_capacity

    MESHLET_TRIS = int(os.environ.get("FMHR_MESHLET_TRIS", "1024"))  # 256 / 512 / 1024 (tuning override)

    def _build_meshlets(self):
        """Setup: Morton-ordered meshlets for the coverage kernel (fmhr_meshlets_build_host, host-side, once per mesh)."""
        import numpy as np
        tri = np.ascontiguousarray(self.faces.cpu().numpy(), dtype=np.int32)
        verts = np.ascontiguousarray(self.vertices_tmp.cpu().numpy(), dtype=np.float32)
        hp = lambda a: ctypes.c_void_p(a.ctypes.data) if a is not None else ctypes.c_void_p(0)
        nm, nr, mv = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
        call = lambda a, b, c: check(self.lib.fmhr_meshlets_build_host(
            hp(tri), hp(verts), self.V, self.T, self.MESHLET_TRIS, ctypes.byref(nm), ctypes.byref(nr), ctypes.byref(mv),
            hp(a), hp(b), hp(c)), "meshlets_build_host")
        call(None, None, None)
        vptr = np.zeros(nm.value + 1, dtype=np.int32)
        vrefs = np.zeros(max(nr.value, 1), dtype=np.int32)
        tri2 = np.zeros((nm.value * self.MESHLET_TRIS, 2), dtype=np.uint32)
        call(vptr, vrefs, tri2)
        dev = self.device
        self.ml_vptr = torch.from_numpy(vptr).to(dev)
        self.ml_verts = torch.from_numpy(vrefs).to(dev)
        self.ml_tri2 = torch.from_numpy(tri2.view(np.int32)).to(dev)
        self.n_meshlets, self.ml_max_verts = nm.value, mv.value
        self.ml_pos = torch.zeros(nm.value * mv.value, 4, dtype=torch.float32, device=dev)  # meshlet-ordered vertices (scratch)

    # ------------------------------------------------------------------ initialisation (mesh_sfs_optim.py:124-177)
    def initialise(self, grayimgs):
        """HAM initialisation on the fused forward chain (fmhr_ham_init): renders every view of the INITIAL mesh once and
        sets, exactly as the reference does before its loops, `valid_masks` (antialiased coverage, :146,163), `sh_coeffs`
        (per-view least-squares SH lighting, :152-159), and `albedo` (img / radiance of the global SH fit, averaged over
        the valid pixels, broadcast to every vertex, :165-176).  Returns dict(sh_coeff=global [9], albedo_mean=[3])."""
        dev = self.device
        gray = grayimgs.detach().to(device=dev, dtype=torch.float32).contiguous()
        if gray.shape != self.masks.shape:
            raise RuntimeError("fmhr_b200: grayimgs must be [num,H,W]")
        views = torch.arange(self.num, dtype=torch.int32, device=dev)
        cfg = self._cfg(self.num, 0, None)
        buf = self._buffers(cfg, views)
        valid = torch.empty_like(self.masks)
        sh = torch.empty(self.num, 9, dtype=torch.float32, device=dev)
        sh_g = torch.empty(9, dtype=torch.float32, device=dev)
        alb = torch.empty(3, dtype=torch.float32, device=dev)
        scratch = torch.empty(self.lib.fmhr_ham_init_scratch_bytes(self.num), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            self._prepare_zbuf_local(cfg, buf)
            check(self.lib.fmhr_ham_init(ctypes.byref(cfg), ctypes.byref(buf), ptr(gray), ptr(valid), ptr(sh), ptr(sh_g),
                                         ptr(alb), ptr(scratch), stream()), "ham_init")
            self.valid_masks = valid
            self.sh_coeffs.copy_(sh)
            self.albedo.copy_(alb[None].expand(self.V, 3))
            check(self.lib.fmhr_ham_prepare_views(ptr(self.valid_masks), self.num, self.H, self.W, ptr(self.view_vm2),
                                                  stream()), "ham_prepare_views")
        self._struct_cache.clear()
        self._graphs = {}; self._idx_tags = {}
        return dict(sh_coeff=sh_g, albedo_mean=alb)

    # ------------------------------------------------------------------ reference-shaped accessors
    @property
    def vertices(self):
        return self.vertices_tmp + self.delta

    def begin_phase_b(self):
        """mesh_sfs_optim.py:242-244: a fresh Adam over (delta, albedo, sh_coeffs) -> moments and steps restart."""
        self.adam_m.zero_()
        self.adam_v.zero_()
        self.adam_step[:3].zero_()  # [3] is the latched fatal flag of the peer exchange (check_health)
        self.phase = 1

    # ------------------------------------------------------------------ plumbing
    def _cfg(self, n_views, phase, albedo_weight, n_views_global=None):
        c = self.conf
        cfg = HamConfig()
        cfg.V, cfg.T, cfg.H, cfg.W = self.V, self.T, self.H, self.W
        cfg.n_views = n_views
        nvg = n_views_global or self.n_views_global_override
        if nvg is None and self._uneven_shards:
            raise RuntimeError("fmhr_b200: the ranks hold different numbers of views - pass n_views_global (the size of the "
                               "global batch of the step) so that every rank normalises the mask loss identically")
        cfg.n_views_global = int(nvg) if nvg else n_views * self.world
        cfg.phase = phase
        cfg.n_sh_rows = self.num
        cfg.zbuf_slot = 0
        cfg.view_groups = self.view_groups
        cfg.sfs_weight, cfg.lap_weight = c["sfs_weight"], c["lap_weight"]
        cfg.albedo_weight = c["albedo_weight"] if albedo_weight is None else albedo_weight
        cfg.mask_weight, cfg.edge_weight, cfg.delta_weight = c["mask_weight"], c["edge_weight"], c["delta_weight"]
        cfg.lr, cfg.albedo_lr, cfg.sh_lr = c["lr"], c["albedo_lr"], c["sh_lr"]
        cfg.beta1, cfg.beta2, cfg.eps = 0.9, 0.999, 1e-8
        cfg.edge_length_mean = self.edge_length_mean
        if self.batch_capacity and n_views > self.batch_capacity:
            self.set_batch_capacity(n_views)  # a larger batch than announced: the layout grows (and everything is re-armed)
        cfg.n_views_capacity = self.batch_capacity
        return cfg

    def _buffers(self, cfg, view_idx, imgs=None, masks=None, valid_masks=None, w2cs=None, projs=None, sh_idx=None,
                 view_vm2=None):
        need = self.lib.fmhr_ham_workspace_bytes(ctypes.byref(cfg))
        if need == 0:
            raise RuntimeError("fmhr_b200: invalid HAM configuration")
        if self.workspace is None or self.workspace.numel() < need:
            self.workspace = None
            self._graphs = {}; self._idx_tags = {}
            self.workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
        b = HamBuffers()
        t = self.topo
        b.tri, b.opp = ptr(self.faces), ptr(t.opp)
        b.v2f_ptr, b.v2f_idx, b.v2v_ptr, b.v2v_idx = ptr(t.v2f_ptr), ptr(t.v2f_idx), ptr(t.v2v_ptr), ptr(t.v2v_idx)
        b.v2f_nbr, b.inv_deg = ptr(t.v2f_nbr), ptr(t.inv_deg)
        b.ml_vptr, b.ml_verts, b.ml_tri2 = ptr(self.ml_vptr), ptr(self.ml_verts), ptr(self.ml_tri2)
        b.n_meshlets, b.ml_tris, b.ml_max_verts = self.n_meshlets, self.MESHLET_TRIS, self.ml_max_verts
        b.ml_pos = ptr(self.ml_pos)
        b.vertices_tmp, b.delta, b.albedo, b.sh_coeffs = ptr(self.vertices_tmp), ptr(self.delta), ptr(self.albedo), ptr(self.sh_coeffs)
        b.adam_m, b.adam_v, b.adam_step = ptr(self.adam_m), ptr(self.adam_v), ptr(self.adam_step)
        b.imgs = ptr(self.imgs if imgs is None else imgs)
        b.masks = ptr(self.masks if masks is None else masks)
        b.valid_masks = ptr(self.valid_masks if valid_masks is None else valid_masks)
        b.view_vm2 = ptr(self.view_vm2 if view_vm2 is None else view_vm2)
        b.w2cs = ptr(self.w2cs if w2cs is None else w2cs)
        b.projs = ptr(self.projs if projs is None else projs)
        b.view_idx = ptr(view_idx)
        b.sh_idx = ptr(sh_idx)
        b.packed, b.losses = ptr(self.packed), ptr(self.losses)
        b.workspace, b.workspace_bytes = ptr(self.workspace), self.workspace.numel()
        b.dbg_grad = ptr(self.dbg_grad)
        if self.debug and cfg.phase == 0:
            self.dbg_grad_sh = torch.zeros(self.num, 9, dtype=torch.float32, device=self.device)
        b.dbg_grad_sh = ptr(self.dbg_grad_sh if (self.debug and cfg.phase == 0) else None)
        return b

    def set_batch_capacity(self, n_views):
        """Lay the workspace out for batches of UP TO n_views views (fmhr_ham_config.n_views_capacity).  Steps with different
        batch sizes - the reference's loop ends every epoch with a short batch, mesh_sfs_optim.py:252-254 - then share one
        layout and keep each other's z-buffer invariants: no fmhr_ham_reset (two z-buffer memsets) whenever the batch size
        changes.  0 = lay out for each step's own size (a change of size resets).  Invalidates the workspace and the
        captured graphs."""
        n_views = int(n_views)
        if n_views != self.batch_capacity:
            self.batch_capacity = n_views
            self.workspace = None
            self._graphs = {}; self._idx_tags = {}
            self._struct_cache.clear()
            self._zb_layout = None

    def _layout_views(self, n):
        """Number of views the workspace of a step with n views is laid out for."""
        return max(self.batch_capacity, n) if self.batch_capacity else n

    def _prepare_zbuf_local(self, cfg, buf):
        """_prepare_zbuf for rank-local helpers (initialise / stage_times / export): with the peer exchange the slot
        parity selects the shared `packed` buffer and must only advance on steps EVERY rank makes, so these helpers
        put the parity back and force a z-buffer reset before the next real step."""
        if self.peer is None:
            return self._prepare_zbuf(cfg, buf)
        slot = self._zb_slot
        self._prepare_zbuf(cfg, buf)
        self._zb_slot = slot
        self._zb_layout = None

    def _prepare_zbuf(self, cfg, buf):
        """The fused step rasterises z-buffer slot `cfg.zbuf_slot` and resets the other one for the next step, so the
        slot alternates every render; both slots are reset whenever the workspace layout changes."""
        layout = (self._layout_views(cfg.n_views), cfg.phase == 0, self.workspace.data_ptr())
        if layout != self._zb_layout:
            check(self.lib.fmhr_ham_reset(ctypes.byref(cfg), ctypes.byref(buf), stream()), "ham_reset")
            self._zb_layout = layout  # both slots are clean now: the parity simply carries on (see _run_step)
        cfg.zbuf_slot = self._zb_slot
        self._zb_slot ^= 1

    def _views(self, view_idx):
        if torch.is_tensor(view_idx):
            return view_idx.to(device=self.device, dtype=torch.int32).contiguous()
        key = tuple(int(i) for i in view_idx)
        t = self._view_idx_cache.get(key)
        if t is None:
            t = torch.tensor(key, dtype=torch.int32, device=self.device)
            if len(self._view_idx_cache) < 64:
                self._view_idx_cache[key] = t
        return t

    def _step(self, phase, view_idx, albedo_weight=None, n_views_global=None):
        if self.use_graphs:
            return self._step_graph(phase, view_idx, albedo_weight, n_views_global)
        vi = self._views(view_idx)
        key = (phase, vi.data_ptr(), vi.numel(), albedo_weight, n_views_global)
        cb = self._struct_cache.get(key)
        if cb is None or cb[2] != (self.workspace.data_ptr() if self.workspace is not None else 0):
            cfg = self._cfg(vi.numel(), phase, albedo_weight, n_views_global)
            buf = self._buffers(cfg, vi)
            cb = (cfg, buf, self.workspace.data_ptr(), vi)
            if len(self._struct_cache) > 256:
                self._struct_cache.clear()
            self._struct_cache[key] = cb
        cfg, buf = cb[0], cb[1]
        if torch.cuda.current_device() != self.device.index:
            torch.cuda.set_device(self.device)
        self._prepare_zbuf(cfg, buf)
        self._run_step(cfg, buf, stream())
        return self.losses

    def _run_step(self, cfg, buf, sp, part=None):
        """The C-ABI calls of one iteration on stream `sp` (part: None = all, 0 = render, 1 = update)."""
        if self.peer is not None:
            # the shared buffer alternates with the z-buffer slot, which flips on EVERY step of this optimiser (eager,
            # graph replay, host-streaming, workspace resets included): a peer may still be reading the previous one
            buf.packed = self.peer.packed[cfg.zbuf_slot].data_ptr()
        if part in (None, 0):
            check(self.lib.fmhr_ham_step_render(ctypes.byref(cfg), ctypes.byref(buf), sp), "ham_step_render")
        if part in (None, 0) and cfg.phase == 1 and self.extra_terms:
            # extra loss terms work on the step's vertices and add their delta gradient before the update (torch ops inside
            # run on the current stream, which is `sp` in every caller)
            for term in self.extra_terms:
                term.accumulate(self, cfg, buf, sp)
        if part is None and self.world > 1 and self.peer is None:
            allreduce_packed(self.packed, self.pg)  # one NCCL sum per iteration
        if part in (None, 1):
            if self.peer is not None:
                check(self.lib.fmhr_ham_step_update_peer(ctypes.byref(cfg), ctypes.byref(buf),
                                                         ctypes.byref(self.peer.structs[cfg.zbuf_slot]), sp),
                      "ham_step_update_peer")
            else:
                check(self.lib.fmhr_ham_step_update(ctypes.byref(cfg), ctypes.byref(buf), sp), "ham_step_update")

    def _step_graph(self, phase, view_idx, albedo_weight, n_views_global=None):
        """CUDA-graph replay of the iteration: the launches + memsets of render/update are captured once per
        (phase, batch size, albedo_weight, z-buffer slot) and replayed with one launch; the step's view indices are
        copied into a persistent device buffer the captured kernels read.  With more than one rank the peer exchange is
        part of the captured update (still one launch per iteration); the NCCL variant runs its all-reduce eagerly between
        the two captured halves."""
        n = view_idx.numel() if torch.is_tensor(view_idx) else len(view_idx)
        key = (phase, n, None if albedo_weight is None else float(albedo_weight), n_views_global)
        ent = self._graphs.get(key)
        if ent is None:
            ent = self._capture(phase, n, albedo_weight, view_idx, n_views_global)
            self._graphs[key] = ent
        graphs, idx_buf, ws_ptr, cfgs, bufs = ent
        if self.workspace.data_ptr() != ws_ptr:
            raise RuntimeError("fmhr_b200: workspace was reallocated after graph capture")
        layout = (self._layout_views(n), phase == 0, ws_ptr)
        if layout != self._zb_layout:
            check(self.lib.fmhr_ham_reset(ctypes.byref(cfgs[0]), ctypes.byref(bufs[0]), stream()), "ham_reset")
            self._zb_layout = layout
        slot = self._zb_slot
        self._zb_slot ^= 1
        vi = self._views(view_idx)
        if vi.data_ptr() != idx_buf.data_ptr():
            # the captured kernels read the batch's view indices from `idx_buf`; a loop that passes the SAME index tensor
            # again (same object, not written since: the reference's batch == all views case) needs no second copy - the
            # tiny copy kernel between two graph launches costs ~3 us of every iteration
            tag = self._idx_tags.get(key)
            if tag is None or tag[0] is not vi or tag[1] != vi._version or _IDX_COPY_ALWAYS:
                idx_buf.copy_(vi, non_blocking=True)
                self._idx_tags[key] = (vi, vi._version)  # (holds the tensor: its address cannot be recycled)
        graphs[slot][0].replay()
        if len(graphs[slot]) > 1:
            allreduce_packed(self.packed, self.pg)
            graphs[slot][1].replay()
        return self.losses

    def _capture(self, phase, n, albedo_weight, view_idx, n_views_global=None):
        idx_buf = torch.zeros(n, dtype=torch.int32, device=self.device)
        cfgs, bufs = [], []
        for slot in (0, 1):
            cfg = self._cfg(n, phase, albedo_weight, n_views_global)
            cfg.zbuf_slot = slot
            cfgs.append(cfg)
            bufs.append(self._buffers(cfg, idx_buf))
        ws_ptr = self.workspace.data_ptr()
        idx_buf.copy_(self._views(view_idx))
        state = (self.delta, self.albedo, self.sh_coeffs, self.adam_m, self.adam_v, self.adam_step)
        saved = [t.clone() for t in state]  # the warm-up runs below must not advance the optimiser
        split = self.world > 1 and self.peer is None  # NCCL variant: the all-reduce sits between two graphs
        torch.cuda.synchronize(self.device)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        graphs = [[], []]
        with torch.cuda.stream(side):
            sp = _lib.c_p(side.cuda_stream)
            check(self.lib.fmhr_ham_reset(ctypes.byref(cfgs[0]), ctypes.byref(bufs[0]), sp), "ham_reset")
            # warm-up outside capture (lazy module loading): two steps continuing the slot parity of the steps before
            # (the peer exchange relies on `packed` alternating on EVERY step, resets included)
            for slot in (self._zb_slot, self._zb_slot ^ 1):
                self._run_step(cfgs[slot], bufs[slot], sp)
            side.synchronize()
            for slot in (0, 1):
                for part in ((0, 1) if split else (None,)):
                    gr = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gr, stream=side):
                        sp2 = _lib.c_p(torch.cuda.current_stream(self.device).cuda_stream)
                        self._run_step(cfgs[slot], bufs[slot], sp2, part)
                    graphs[slot].append(gr)
            check(self.lib.fmhr_ham_reset(ctypes.byref(cfgs[0]), ctypes.byref(bufs[0]), sp), "ham_reset")
        torch.cuda.current_stream(self.device).wait_stream(side)
        for t, sv in zip(state, saved):
            t.copy_(sv)
        self._zb_layout = (self._layout_views(n), phase == 0, ws_ptr)  # two warm-up steps: the parity is where it was
        return graphs, idx_buf, ws_ptr, cfgs, bufs

    # ------------------------------------------------------------------ the two loops' bodies
    def step_phase_a(self, view_idx, n_views_global=None):
        """One iteration of mesh_sfs_optim.py:198-237 (albedo + SH warm-up).  Returns the device loss record
        [sfs, 0, albedo(display), 0, 0, 0, n_valid, total]; no host sync."""
        return self._step(0, view_idx, None, n_views_global)

    def step_phase_b(self, view_idx, albedo_weight=None, n_views_global=None):
        """One iteration of mesh_sfs_optim.py:253-310.  Returns the device loss record
        [sfs, lap, albedo, mask, edge, delta, n_valid, total]; no host sync.  n_views_global: size of the step's GLOBAL
        batch when the ranks' shards are uneven (default: local batch x world)."""
        if self.phase != 1:
            self.begin_phase_b()
        return self._step(1, view_idx, albedo_weight, n_views_global)

    def check_health(self):
        """Raises if the peer exchange gave up on a rank (fmhr_ham_step_update_peer latched adam_step[3]): the update of
        that step and of every later one was refused on this rank - parameters are intact, but the replicas are no longer
        in step, so the job must stop.  Host sync; call it wherever the loss record is read."""
        if int(self.adam_step[3].item()) != 0:
            raise RuntimeError("fmhr_b200: a peer did not reach the gradient exchange within the timeout "
                               "(FMHR_PEER_TIMEOUT_S); this rank refused the update - parameters are untouched, abort the job")

    STAGES = ("clears", "vertex_normals", "clip_transform", "coverage", "shade", "antialias_loss", "pixel_backward",
              "update_adam")

    def stage_times(self, view_idx, phase=1, repeats=5):
        """Per-kernel device time (ms) of one iteration, averaged over `repeats` (each advances the optimiser)."""
        vi = self._views(view_idx)
        cfg = self._cfg(vi.numel(), phase, None)
        buf = self._buffers(cfg, vi)
        ms = (_lib.c_f * 16)()
        n = _lib.c_i(0)
        acc = [0.0] * len(self.STAGES)
        with torch.cuda.device(self.device):
            for _ in range(repeats):
                self._prepare_zbuf_local(cfg, buf)
                check(self.lib.fmhr_ham_stage_times(ctypes.byref(cfg), ctypes.byref(buf), ms, ctypes.byref(n), stream()),
                      "ham_stage_times")
                for i in range(min(n.value, len(acc))):
                    acc[i] += ms[i] / repeats
        return dict(zip(self.STAGES, acc))

    def export(self, view_idx, phase=1):
        """Forward-only inspection of the fused path (parity tests): pos, rast, antialiased image, antialiased
        coverage, vertex normals."""
        vi = self._views(view_idx)
        n = vi.numel()
        cfg = self._cfg(n, phase, None)
        buf = self._buffers(cfg, vi)
        dev = self.device
        pos = torch.empty(n, self.V, 4, dtype=torch.float32, device=dev)
        rast = torch.empty(n, self.H, self.W, 4, dtype=torch.float32, device=dev)
        image = torch.zeros(n, self.H, self.W, 3, dtype=torch.float32, device=dev)
        pmask = torch.zeros(n, self.H, self.W, dtype=torch.float32, device=dev)
        normals = torch.empty(self.V, 3, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            self._prepare_zbuf_local(cfg, buf)
            check(self.lib.fmhr_ham_debug_export(ctypes.byref(cfg), ctypes.byref(buf), ptr(pos), ptr(rast), ptr(image),
                                                 ptr(pmask), ptr(normals), stream()), "ham_debug_export")
        return dict(pos=pos, rast=rast, image=image, pred_mask=pmask, normals=normals)


class HostStreamingStepper:
    """End-to-end variant used by bench.py's `e2e` leg: the step's view batch lives in PINNED HOST memory and is copied
    to the device inside the call (fmhr_ham_step_host), and the 8-float loss record is copied back."""

    def __init__(self, opt, n_views):
        self.opt = opt
        self.n = n_views
        dev = opt.device
        H, W = opt.H, opt.W
        self.d_imgs = torch.empty(n_views, H, W, 3, dtype=torch.float32, device=dev)
        self.d_masks = torch.empty(n_views, H, W, dtype=torch.float32, device=dev)
        self.d_valid = torch.empty(n_views, H, W, dtype=torch.float32, device=dev)
        self.d_w2cs = torch.empty(n_views, 4, 4, dtype=torch.float32, device=dev)
        self.d_projs = torch.empty(n_views, 4, 4, dtype=torch.float32, device=dev)
        self.rows = torch.arange(n_views, dtype=torch.int32, device=dev)
        self.d_vm2 = torch.zeros(n_views, opt.n_tiles + 1, dtype=torch.float64, device=dev)
        self.losses_host = torch.zeros(8, dtype=torch.float32).pin_memory()
        self.h2d_bytes = 4 * (n_views * H * W * 5 + n_views * 32)
        self.d2h_bytes = 32
        self.staging = None
        self.h2d_bytes_u8 = n_views * H * W * 4 + 4 * n_views * 32

    def step_phase_b(self, h_imgs, h_masks, h_valid, h_w2cs, h_projs, sh_rows, albedo_weight=None):
        """h_* are pinned host tensors holding this step's n_views rows; sh_rows = int32 device tensor of SH rows."""
        o = self.opt
        for t in (h_imgs, h_masks, h_valid, h_w2cs, h_projs):
            if t.is_cuda or not t.is_contiguous() or t.dtype != torch.float32 or not t.is_pinned():
                raise RuntimeError("HostStreamingStepper: host batches must be pinned, contiguous float32 CPU tensors")
        if o.phase != 1:
            o.begin_phase_b()
        cfg = o._cfg(self.n, 1, albedo_weight)
        buf = o._buffers(cfg, self.rows, self.d_imgs, self.d_masks, self.d_valid, self.d_w2cs, self.d_projs, sh_rows,
                         self.d_vm2)
        if o.world > 1:
            raise RuntimeError("HostStreamingStepper is single-process; use HamOptimizer under torch.distributed")
        with torch.cuda.device(o.device):
            o._prepare_zbuf(cfg, buf)
            check(o.lib.fmhr_ham_step_host(ctypes.byref(cfg), ctypes.byref(buf), ptr(h_imgs), ptr(h_masks), ptr(h_valid),
                                           ptr(h_w2cs), ptr(h_projs), ptr(self.losses_host), stream()), "ham_step_host")
        return self.losses_host

    def set_resident_valid_masks(self, valid_masks):
        """valid_masks of the step's views stay on the device for the u8 path (the reference derives them on the GPU,
        mesh_sfs_optim.py:146-163); also prepares their mask-loss constants."""
        o = self.opt
        self.d_valid.copy_(valid_masks)
        with torch.cuda.device(o.device):
            check(o.lib.fmhr_ham_prepare_views(ptr(self.d_valid), self.n, o.H, o.W, ptr(self.d_vm2), stream()),
                  "ham_prepare_views")

    def step_phase_b_u8(self, h_imgs_u8, h_masks_u8, h_w2cs, h_projs, sh_rows, albedo_weight=None):
        """Host batch in its native 8-bit form: h_imgs_u8 [n,H,W,3] uint8 (img = u8/255), h_masks_u8 [n,H,W] uint8
        (mask = u8 > 127), cameras float32 - all pinned; valid_masks resident (set_resident_valid_masks)."""
        o = self.opt
        for t, dt in ((h_imgs_u8, torch.uint8), (h_masks_u8, torch.uint8), (h_w2cs, torch.float32), (h_projs, torch.float32)):
            if t.is_cuda or not t.is_contiguous() or t.dtype != dt or not t.is_pinned():
                raise RuntimeError("HostStreamingStepper: host batches must be pinned, contiguous CPU tensors (uint8 images / "
                                   "masks, float32 cameras)")
        if o.phase != 1:
            o.begin_phase_b()
        if o.world > 1:
            raise RuntimeError("HostStreamingStepper is single-process; use HamOptimizer under torch.distributed")
        cfg = o._cfg(self.n, 1, albedo_weight)
        buf = o._buffers(cfg, self.rows, self.d_imgs, self.d_masks, self.d_valid, self.d_w2cs, self.d_projs, sh_rows,
                         self.d_vm2)
        if self.staging is None:
            self.staging = torch.empty(o.lib.fmhr_ham_host_u8_staging_bytes(ctypes.byref(cfg)), dtype=torch.uint8,
                                       device=o.device)
        with torch.cuda.device(o.device):
            o._prepare_zbuf(cfg, buf)
            check(o.lib.fmhr_ham_step_host_u8(ctypes.byref(cfg), ctypes.byref(buf), ptr(h_imgs_u8), ptr(h_masks_u8),
                                              ptr(h_w2cs), ptr(h_projs), ptr(self.staging), ptr(self.losses_host),
                                              stream()), "ham_step_host_u8")
        return self.losses_host

    # ---- pipelined form: the next step's batch travels while the current step computes
    @staticmethod
    def mask_boxes(masks_u8):
        """Loader metadata for submit_u8(boxes=...): per view the half-open rectangle (y0, y1, x0, x1) that contains every
        pixel with mask byte > 127 ((0,0,0,0) for an empty mask).  masks_u8: [n,H,W] uint8 (CPU tensor or array)."""
        import numpy as np
        m = np.asarray(masks_u8) > 127
        out = np.zeros((m.shape[0], 4), dtype=np.int32)
        for v in range(m.shape[0]):
            ys, xs = np.flatnonzero(m[v].any(axis=1)), np.flatnonzero(m[v].any(axis=0))
            if ys.size:
                out[v] = (ys[0], ys[-1] + 1, xs[0], xs[-1] + 1)
        return torch.from_numpy(out)

    def ensure_direct_planes(self):
        """The two float plane sets (images, masks, cameras) of the converted-on-arrival form.  The library owns their
        content between submits (it keeps the mask planes zero outside the boxes of the last batch)."""
        if getattr(self, "_planes", None) is None:
            o = self.opt
            mk = lambda *shape: torch.zeros(*shape, dtype=torch.float32, device=o.device)
            self._planes = [(mk(self.n, o.H, o.W, 3), mk(self.n, o.H, o.W), mk(self.n, 4, 4), mk(self.n, 4, 4))
                            for _ in range(2)]
            self._next_plane = 0
            self._plane_fresh = [True, True]  # content undefined for the library until its first batch
        return self._planes

    def submit_u8(self, h_imgs_u8, h_masks_u8, boxes=None, cameras=None):
        """Start the upload of a host batch (pinned uint8 images [n,H,W,3] / masks [n,H,W]) into the next free staging
        buffer; returns a ticket for step_submitted_u8.  At most two batches may be in flight.  boxes (mask_boxes(): int32
        CPU tensor [n,4]) restricts the transfer to the rectangle of every view that holds its segmentation.
        cameras = (h_w2cs, h_projs) (pinned float32 [n,4,4], with boxes): converted-on-arrival form - the pull kernel writes
        float planes of the ticket's own plane set and the cameras travel with the batch, so the step itself uploads and
        converts nothing (fmhr_ham_host_u8_submit_boxes_direct); step_submitted_u8 is then called with h_w2cs = h_projs = None."""
        o = self.opt
        for t in (h_imgs_u8, h_masks_u8):
            if t.is_cuda or not t.is_contiguous() or t.dtype != torch.uint8 or not t.is_pinned():
                raise RuntimeError("HostStreamingStepper: host batches must be pinned, contiguous uint8 CPU tensors")
        cfg = self.__dict__.get("_submit_cfg")
        if cfg is None:
            cfg = self._submit_cfg = o._cfg(self.n, 1, None)
        if cameras is not None:
            if boxes is None:
                raise RuntimeError("HostStreamingStepper: the converted-on-arrival form needs the segmentation boxes")
            h_w2cs, h_projs = cameras
            for t in (h_w2cs, h_projs):
                if t.is_cuda or not t.is_contiguous() or t.dtype != torch.float32 or not t.is_pinned():
                    raise RuntimeError("HostStreamingStepper: cameras must be pinned, contiguous float32 CPU tensors")
            if boxes.is_cuda or boxes.dtype != torch.int32 or tuple(boxes.shape) != (self.n, 4) or not boxes.is_contiguous():
                raise RuntimeError("HostStreamingStepper: boxes must be a contiguous int32 CPU tensor [n,4]")
            self.ensure_direct_planes()
            slot = self._next_plane
            self._next_plane ^= 1
            pl = self._planes[slot]
            nb = _lib.c_sz(0)
            with torch.cuda.device(o.device):
                check(o.lib.fmhr_ham_host_u8_submit_boxes_direct(
                    ctypes.byref(cfg), ptr(h_imgs_u8), ptr(h_masks_u8), ptr(boxes), ptr(h_w2cs), ptr(h_projs), ptr(pl[0]),
                    ptr(pl[1]), ptr(pl[2]), ptr(pl[3]), 1 if self._plane_fresh[slot] else 0, ctypes.byref(nb)),
                    "ham_host_u8_submit_boxes_direct")
            self._plane_fresh[slot] = False
            self.last_submit_bytes = int(nb.value)
            return ("direct", slot)
        if getattr(self, "_stagings", None) is None:
            nbytes = o.lib.fmhr_ham_host_u8_staging_bytes(ctypes.byref(cfg))
            self._stagings = [torch.empty(nbytes, dtype=torch.uint8, device=o.device) for _ in range(2)]
            self._next_slot = 0
        slot = self._next_slot
        self._next_slot ^= 1
        with torch.cuda.device(o.device):
            if boxes is None:
                check(o.lib.fmhr_ham_host_u8_submit(ctypes.byref(cfg), ptr(h_imgs_u8), ptr(h_masks_u8),
                                                    ptr(self._stagings[slot])), "ham_host_u8_submit")
                self.last_submit_bytes = self.h2d_bytes_u8 - 4 * self.n * 32
            else:
                if boxes.is_cuda or boxes.dtype != torch.int32 or tuple(boxes.shape) != (self.n, 4) or not boxes.is_contiguous():
                    raise RuntimeError("HostStreamingStepper: boxes must be a contiguous int32 CPU tensor [n,4]")
                nb = _lib.c_sz(0)
                check(o.lib.fmhr_ham_host_u8_submit_boxes(ctypes.byref(cfg), ptr(h_imgs_u8), ptr(h_masks_u8), ptr(boxes),
                                                          ptr(self._stagings[slot]), ctypes.byref(nb)),
                      "ham_host_u8_submit_boxes")
                self.last_submit_bytes = int(nb.value)
        return slot

    def step_submitted_u8(self, ticket, h_w2cs, h_projs, sh_rows, albedo_weight=None, async_record=False):
        """The iteration on the batch submitted under `ticket` (cameras: pinned float32 [n,4,4]).  With use_graphs (the
        optimiser's setting) the device work - camera upload, conversion, render, update, loss read-back - is captured
        once per (staging buffer, z-buffer slot, host pointers) and replayed with one launch; the handshakes with the copy
        stream stay outside the graph (fmhr_ham_step_host_u8_acquire / _release).

        async_record=True: the step's loss record is copied into slot `ticket` of a two-slot pinned ring and an event is
        recorded behind it; nothing waits here.  read_record(ticket) returns it later (typically one step late, after the
        next step has been launched), so the host never stalls the device between steps - the reference's per-iteration
        `.item()` (mesh_sfs_optim.py:312) only feeds a progress bar."""
        o = self.opt
        direct = isinstance(ticket, tuple)
        planes = None
        if direct:
            ticket = ticket[1]
            planes = self._planes[ticket]
            if h_w2cs is not None or h_projs is not None:
                raise RuntimeError("HostStreamingStepper: the cameras of a converted-on-arrival batch travelled with submit_u8")
        if async_record:
            if getattr(self, "_ring", None) is None:
                self._ring = torch.zeros(2, 8, dtype=torch.float32).pin_memory()
                self._ring_events = [torch.cuda.Event(), torch.cuda.Event()]
            record = self._ring[ticket]
        else:
            record = self.losses_host
        for t in (() if direct else (h_w2cs, h_projs)):
            if t.is_cuda or not t.is_contiguous() or t.dtype != torch.float32 or not t.is_pinned():
                raise RuntimeError("HostStreamingStepper: cameras must be pinned, contiguous float32 CPU tensors")
        if o.phase != 1:
            o.begin_phase_b()
        if o.world > 1 and o.peer is None:
            raise RuntimeError("HostStreamingStepper needs the peer-memory exchange on more than one rank")
        # the per-step Python work is on the critical path (the loss record is read back after every step): the
        # configuration / buffer structs are built once per (albedo_weight, SH rows, workspace)
        ckey = (None if albedo_weight is None else float(albedo_weight), sh_rows.data_ptr(),
                0 if o.workspace is None else o.workspace.data_ptr(), ticket if direct else -1)
        cache = self.__dict__.setdefault("_struct_cache", {})
        cb = cache.get(ckey)
        if cb is None or cb[2] != (0 if o.workspace is None else o.workspace.data_ptr()):
            cfg = o._cfg(self.n, 1, albedo_weight)
            src = planes if direct else (self.d_imgs, self.d_masks, self.d_w2cs, self.d_projs)
            buf = o._buffers(cfg, self.rows, src[0], src[1], self.d_valid, src[2], src[3], sh_rows, self.d_vm2)
            if len(cache) > 32:
                cache.clear()
            cb = cache[(ckey[0], ckey[1], o.workspace.data_ptr(), ckey[3])] = (cfg, buf, o.workspace.data_ptr(), sh_rows)
        cfg, buf = cb[0], cb[1]
        staging = planes[0] if direct else self._stagings[ticket]
        with torch.cuda.device(o.device):
            o._prepare_zbuf(cfg, buf)
            peers = None
            if o.peer is not None:  # the exchange is part of the update kernels: still one call per iteration
                buf.packed = o.peer.packed[cfg.zbuf_slot].data_ptr()
                peers = ctypes.byref(o.peer.structs[cfg.zbuf_slot])
            body = lambda sp: check(o.lib.fmhr_ham_step_host_u8_body(
                ctypes.byref(cfg), ctypes.byref(buf), ptr(h_w2cs), ptr(h_projs), ptr(staging), ptr(record), peers,
                sp), "ham_step_host_u8_body")
            check(o.lib.fmhr_ham_step_host_u8_acquire(ptr(staging), stream()), "ham_step_host_u8_acquire")
            if not o.use_graphs:
                body(stream())
            else:
                key = (ticket, direct, cfg.zbuf_slot, 0 if direct else h_w2cs.data_ptr(), 0 if direct else h_projs.data_ptr(),
                       sh_rows.data_ptr(),
                       None if albedo_weight is None else float(albedo_weight), o.workspace.data_ptr(), record.data_ptr())
                graphs = self.__dict__.setdefault("_step_graphs", {})
                gr = graphs.get(key)
                if gr is None:
                    if len(graphs) >= 16:
                        graphs.clear()
                    cur = torch.cuda.current_stream(o.device)
                    side = torch.cuda.Stream(device=o.device)
                    side.wait_stream(cur)
                    gr = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gr, stream=side):  # capture only: nothing runs, the optimiser does not advance
                        body(_lib.c_p(torch.cuda.current_stream(o.device).cuda_stream))
                    cur.wait_stream(side)
                    graphs[key] = gr
                gr.replay()
            check(o.lib.fmhr_ham_step_host_u8_release(ptr(staging), stream()), "ham_step_host_u8_release")
            if async_record:
                self._ring_events[ticket].record(torch.cuda.current_stream(o.device))
        return record

    def read_record(self, ticket):
        """Loss record of the step launched with async_record=True under `ticket` (waits for THAT step only)."""
        self._ring_events[ticket].synchronize()
        return self._ring[ticket].clone()
