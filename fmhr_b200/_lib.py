"""ctypes binding of libfmhr_b200.so (the C ABI declared in include/fmhr_b200.h).

There is deliberately NO fallback: if the library is missing or a call fails, a RuntimeError is raised
(upstream nvdiffrast raises RuntimeError from its C++ ops on bad shapes / devices as well).
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("FMHR_B200_LIB") or os.path.join(_HERE, "libfmhr_b200.so")  # override: tuning builds only
_lib = None

c_p = ctypes.c_void_p
c_i = ctypes.c_int
c_f = ctypes.c_float
c_sz = ctypes.c_size_t


class HamConfig(ctypes.Structure):
    """struct fmhr_ham_config"""
    _fields_ = [("V", ctypes.c_int32), ("T", ctypes.c_int32), ("H", ctypes.c_int32), ("W", ctypes.c_int32),
                ("n_views", ctypes.c_int32), ("n_views_global", ctypes.c_int32), ("phase", ctypes.c_int32),
                ("n_sh_rows", ctypes.c_int32), ("zbuf_slot", ctypes.c_int32), ("view_groups", ctypes.c_int32),
                ("sfs_weight", c_f), ("lap_weight", c_f), ("albedo_weight", c_f), ("mask_weight", c_f),
                ("edge_weight", c_f), ("delta_weight", c_f),
                ("lr", c_f), ("albedo_lr", c_f), ("sh_lr", c_f),
                ("beta1", c_f), ("beta2", c_f), ("eps", c_f), ("edge_length_mean", c_f),
                ("n_views_capacity", ctypes.c_int32)]


class HamBuffers(ctypes.Structure):
    """struct fmhr_ham_buffers"""
    _fields_ = [(n, c_p) for n in (
        "tri", "opp", "v2f_ptr", "v2f_idx", "v2v_ptr", "v2v_idx", "v2f_nbr", "inv_deg",
        "vertices_tmp", "delta", "albedo", "sh_coeffs", "adam_m", "adam_v", "adam_step",
        "imgs", "masks", "valid_masks", "view_vm2", "w2cs", "projs", "view_idx", "sh_idx",
        "packed", "losses", "workspace")] + [("workspace_bytes", c_sz), ("dbg_grad", c_p), ("dbg_grad_sh", c_p),
                                             ("ml_vptr", c_p), ("ml_verts", c_p), ("ml_tri2", c_p),
                                             ("n_meshlets", ctypes.c_int32), ("ml_tris", ctypes.c_int32),
                                             ("ml_max_verts", ctypes.c_int32), ("ml_reserved", ctypes.c_int32),
                                             ("ml_pos", c_p)]


MAX_PEERS = 16


class HamPeers(ctypes.Structure):
    """struct fmhr_ham_peers"""
    _fields_ = [("rank", ctypes.c_int32), ("world", ctypes.c_int32), ("mode", ctypes.c_int32),
                ("timeout_s", ctypes.c_int32), ("packed", c_p * MAX_PEERS), ("flags", c_p * MAX_PEERS),
                ("reduced", c_p * MAX_PEERS), ("epoch", c_p)]


_SIGS = {
    "fmhr_version": (c_i, []),
    "fmhr_last_error_string": (ctypes.c_char_p, []),
    "fmhr_rasterize_workspace_bytes": (c_sz, [c_i, c_i, c_i]),
    "fmhr_rasterize_fwd": (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_sz, c_p]),
    "fmhr_rasterize_tile_words": (c_sz, [c_i, c_i, c_i]),
    "fmhr_rasterize_fwd_meshlets": (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p,
                                          c_sz, c_i, c_p, c_p]),
    "fmhr_rasterize_bwd": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p, c_p]),
    "fmhr_interpolate_fwd": (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p]),
    "fmhr_interpolate_bwd": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p]),
    "fmhr_mesh_topology_workspace_bytes": (c_sz, [c_i, c_i]),
    "fmhr_mesh_topology_build": (c_i, [c_p, c_i, c_i, c_p, c_p, c_p, c_p, c_p, ctypes.POINTER(c_i), c_p, c_sz, c_p]),
    "fmhr_mesh_topology_derive": (c_i, [c_p, c_p, c_p, c_i, c_i, c_p, c_p, c_p]),
    "fmhr_meshlets_build_host": (c_i, [c_p, c_p, c_i, c_i, c_i, ctypes.POINTER(c_i), ctypes.POINTER(c_i),
                                       ctypes.POINTER(c_i), c_p, c_p, c_p]),
    "fmhr_antialias_fwd": (c_i, [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p]),
    "fmhr_antialias_bwd": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p]),
    "fmhr_vertex_normals_fwd": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_p, c_p, c_p]),
    "fmhr_vertex_normals_bwd": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_p, c_p, c_p]),
    "fmhr_laplacian_fwd": (c_i, [c_p, c_p, c_p, c_i, c_i, c_p, c_p, c_p]),
    "fmhr_laplacian_bwd": (c_i, [c_p, c_p, c_p, c_i, c_i, c_f, c_p, c_p]),
    "fmhr_sh_radiance_fwd": (c_i, [c_p, c_i, c_p, c_i, c_p, c_p]),
    "fmhr_sh_radiance_bwd": (c_i, [c_p, c_i, c_p, c_p, c_i, c_p, c_p, c_p]),
    "fmhr_ncc_fwd": (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_p]),
    "fmhr_ncc_bwd": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_p]),
    "fmhr_ncc_sample_fwd": (c_i, [c_p] * 7 + [c_i, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p]),
    "fmhr_ncc_sample_bwd": (c_i, [c_p] * 7 + [c_i, c_p, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p]),
    "fmhr_ncc_term_fused": (c_i, [c_p] * 7 + [c_i, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_f, c_p, c_p, c_p]),
    "fmhr_ham_add_delta_grad": (c_i, [ctypes.POINTER(HamConfig), ctypes.POINTER(HamBuffers), c_p, c_p]),
    "fmhr_ham_workspace_bytes": (c_sz, [ctypes.POINTER(HamConfig)]),
    "fmhr_ham_packed_floats": (c_sz, [ctypes.POINTER(HamConfig)]),
    "fmhr_debug_checks": (c_i, []),
    "fmhr_ham_reset": (c_i, [ctypes.POINTER(HamConfig), ctypes.POINTER(HamBuffers), c_p]),
    "fmhr_ham_prepare_views": (c_i, [c_p, c_i, c_i, c_i, c_p, c_p]),
    "fmhr_ham_step_render": (c_i, [ctypes.POINTER(HamConfig), ctypes.POINTER(HamBuffers), c_p]),
    "fmhr_ham_step_update": (c_i, [ctypes.POINTER(HamConfig), ctypes.POINTER(HamBuffers), c_p]),
    "fmhr_ham_step_update_peer": (c_i, [ctypes.POINTER(HamConfig), ctypes.POINTER(HamBuffers), ctypes.POINTER(HamPeers), c_p]),
    "fmhr_peer_alloc": (c_i, [c_sz, ctypes.POINTER(c_p), c_p]),
    "fmhr_peer_open": (c_i, [c_p, ctypes.POINTER(c_p)]),
    "fmhr_peer_close": (c_i, [c_p]),
    "fmhr_peer_free": (c_i, [c_p]),
    "fmhr_ham_stage_times": (c_i, [ctypes.POINTER(HamConfig), ctypes.POINTER(HamBuffers), ctypes.POINTER(c_f),
                                   ctypes.POINTER(c_i), c_p]),
    "fmhr_trace_read": (c_i, [c_p, c_i, c_i]),
    "fmhr_trace_mark": (c_i, [c_i, c_p]),
    "fmhr_ham_debug_export": (c_i, [ctypes.POINTER(HamConfig), ctypes.POINTER(HamBuffers), c_p, c_p, c_p, c_p, c_p, c_p]),
    "fmhr_ham_step_host": (c_i, [ctypes.POINTER(HamConfig), ctypes.POINTER(HamBuffers), c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "fmhr_ham_init_scratch_bytes": (c_sz, [c_i]),
    "fmhr_ham_init": (c_i, [ctypes.POINTER(HamConfig), ctypes.POINTER(HamBuffers), c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "fmhr_ham_host_u8_staging_bytes": (c_sz, [ctypes.POINTER(HamConfig)]),
    "fmhr_ham_host_u8_submit": (c_i, [ctypes.POINTER(HamConfig), c_p, c_p, c_p]),
    "fmhr_ham_step_host_u8_acquire": (c_i, [c_p, c_p]),
    "fmhr_ham_step_host_u8_body": (c_i, [ctypes.POINTER(HamConfig), ctypes.POINTER(HamBuffers), c_p, c_p, c_p, c_p,
                                         ctypes.POINTER(HamPeers), c_p]),
    "fmhr_ham_step_host_u8_release": (c_i, [c_p, c_p]),
    "fmhr_ham_host_u8_submit_boxes": (c_i, [ctypes.POINTER(HamConfig), c_p, c_p, c_p, c_p, ctypes.POINTER(c_sz)]),
    "fmhr_ham_host_u8_submit_boxes_direct": (c_i, [ctypes.POINTER(HamConfig), c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p,
                                                   c_i, ctypes.POINTER(c_sz)]),
    "fmhr_ham_step_host_u8_submitted": (c_i, [ctypes.POINTER(HamConfig), ctypes.POINTER(HamBuffers), c_p, c_p, c_p, c_p,
                                              ctypes.POINTER(HamPeers), c_p]),
    "fmhr_ham_step_host_u8": (c_i, [ctypes.POINTER(HamConfig), ctypes.POINTER(HamBuffers), c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)


def load():
    """Load the shared library (building is __graft_entry__.build()'s / fmhr_b200._build's job)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError("fmhr_b200: %s is missing - run `python -m fmhr_b200._build` (there is no CPU "
                               "or PyTorch fallback for these operators)" % SO_PATH)
        lib = ctypes.CDLL(SO_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().fmhr_last_error_string()
        raise RuntimeError("fmhr_b200.%s failed (code %d): %s" % (what, rc, msg.decode() if msg else "?"))


def ptr(t):
    return c_p(t.data_ptr()) if t is not None else c_p(0)


def stream():
    return c_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("fmhr_b200: expected CUDA tensors (this build has no CPU path)")
