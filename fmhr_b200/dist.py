"""Host-side helpers for the multi-GPU HAM iteration (views shard across ranks).

The loss of one iteration is a mean over the GLOBAL batch:
    sfs  = sfs_weight  * sum_{valid px, rgb} |tmp_img - img| / (3 * N_valid_global)       (mesh_sfs_optim.py:289)
    mask = mask_weight * sum_{px} (pred_mask - valid_mask)^2 / (n_views_global * H * W)    (mesh_sfs_optim.py:295)
so every rank accumulates UN-NORMALISED sums (gradient accumulators and the three scalars) over its own views, one
all-reduce (sum) merges them, and the normalisation by the global counts happens afterwards, identically on every
rank (fmhr_ham_step_update).  Regularisers (Laplacian, edge, delta) are view-independent and computed redundantly.
"""
import ctypes

import torch


def shard_views(num_views, rank, world_size):
    """Round-robin view ownership: rank r owns views r, r + world, r + 2*world, ..."""
    return list(range(rank, num_views, world_size))


def allreduce_packed(packed, group=None):
    """One sum all-reduce of the packed [12V gradient accumulators | n_valid, abs_sum, mask_sq_sum, pad] buffer."""
    if torch.distributed.is_available() and torch.distributed.is_initialized() and \
            torch.distributed.get_world_size(group) > 1:
        torch.distributed.all_reduce(packed, op=torch.distributed.ReduceOp.SUM, group=group)
    return packed


def normalisation_scales(conf, n_valid_global, n_views_global, H, W):
    """(s_photo, s_mask): factors that turn the un-normalised accumulators into the reference's gradients."""
    s_photo = conf["sfs_weight"] / (3.0 * float(n_valid_global))
    s_mask = 2.0 * conf["mask_weight"] / (float(n_views_global) * H * W)
    return s_photo, s_mask


class _RawCuda:
    """A device allocation owned by libfmhr_b200 exposed to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, address, n_floats):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (address, False), "version": 2}


class PeerExchange:
    """NVLink peer-memory exchange of the packed buffer (include/fmhr_b200.h, fmhr_ham_step_update_peer).

    Each rank owns ONE cudaIpc-shared allocation [packed slot 0 | packed slot 1 | reduced | flag words]; the 64-byte handles travel
    through torch.distributed once at setup, after which an iteration needs no host-side collective at all: the update's
    first kernel posts / awaits the ranks' step counters and gathers the peers' accumulators itself.  Construction is
    collective (every rank of `group` must call it); `ok` is False on every rank if any rank could not map a peer, in
    which case the caller keeps the NCCL all-reduce."""

    MODES = {"auto": 0, "oneshot": 1, "twoshot": 2}

    def __init__(self, n_floats, device, group=None, mode=None):
        from ._lib import MAX_PEERS, HamPeers, c_p, check, load
        dist = torch.distributed
        self.lib = load()
        self.device = device
        self.solo = not (dist.is_available() and dist.is_initialized()) or group is False  # one rank exchanging with itself
        self.rank, self.world = (0, 1) if self.solo else (dist.get_rank(group), dist.get_world_size(group))
        self.n_floats = n_floats
        self.slot_bytes = (4 * n_floats + 255) // 256 * 256
        self.base, self.opened, self.ok = None, [], False
        if self.world > MAX_PEERS:
            return
        good = True
        handle = ctypes.create_string_buffer(64)
        base = c_p()
        with torch.cuda.device(device):
            try:
                check(self.lib.fmhr_peer_alloc(3 * self.slot_bytes + 4 * MAX_PEERS * 4, ctypes.byref(base), handle),
                      "peer_alloc")
                self.base = base.value
            except RuntimeError:
                good = False
            handles = [None] * self.world
            if self.solo:
                handles[0] = handle.raw if good else None
            else:
                dist.all_gather_object(handles, handle.raw if good else None, group=group)
            bases = [None] * self.world
            if good and all(h is not None for h in handles):
                for r, h in enumerate(handles):
                    if r == self.rank:
                        bases[r] = self.base
                        continue
                    p = c_p()
                    if self.lib.fmhr_peer_open(ctypes.create_string_buffer(h, 64), ctypes.byref(p)) != 0:
                        good = False
                        break
                    self.opened.append(p.value)
                    bases[r] = p.value
            else:
                good = False
            flag = torch.tensor([1 if good else 0], dtype=torch.int32, device=device)
            if not self.solo:  # agreement; also orders every rank's zero-fill before any rank's first post
                dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            self.ok = bool(flag.item())
        if not self.ok:
            self.close()
            return
        self.epoch = torch.zeros(1, dtype=torch.int32, device=device)
        self.reduced = torch.as_tensor(_RawCuda(self.base + 2 * self.slot_bytes, n_floats), device=device)
        import os
        self.mode = self.MODES[mode or os.environ.get("FMHR_PEER_MODE", "auto")]
        self.packed = [torch.as_tensor(_RawCuda(self.base + s * self.slot_bytes, n_floats), device=device) for s in (0, 1)]
        self.structs = []
        for s in (0, 1):
            st = HamPeers()
            st.rank, st.world, st.mode = self.rank, self.world, self.mode
            st.timeout_s = int(os.environ.get("FMHR_PEER_TIMEOUT_S", "0"))  # 0 = the library's default (30 s)
            for r in range(self.world):
                st.packed[r] = bases[r] + s * self.slot_bytes
                st.reduced[r] = bases[r] + 2 * self.slot_bytes
                st.flags[r] = bases[r] + 3 * self.slot_bytes
            st.epoch = self.epoch.data_ptr()
            self.structs.append(st)

    def close(self):
        for p in self.opened:
            self.lib.fmhr_peer_close(ctypes.c_void_p(p))
        self.opened = []
        if self.base is not None:
            self.lib.fmhr_peer_free(ctypes.c_void_p(self.base))
            self.base = None
