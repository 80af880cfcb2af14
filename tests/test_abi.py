"""The C-ABI shared library loads without a GPU and exports every symbol include/fmhr_b200.h declares; the ctypes
mirrors of the two structs match the C layout; host-only size queries answer without touching a device."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fmhr_b200.h")


@pytest.fixture(scope="module")
def lib():
    from fmhr_b200 import _build, _lib
    _build.build()
    return _lib.load()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fmhr_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    from fmhr_b200 import _lib
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libfmhr_b200.so does not export %s" % n
    assert sorted(_lib.EXPORTED_SYMBOLS) == names, "ctypes signature table and header disagree"
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.SO_PATH], text=True)
    exported = set(re.findall(r" T (fmhr_[a-z0-9_]+)", out))
    assert exported == set(names)


def test_struct_layout_matches_header(tmp_path):
    from fmhr_b200._lib import HamBuffers, HamConfig, HamPeers
    prog = tmp_path / "layout.c"
    fields_c = [f for f, _ in HamConfig._fields_]
    fields_b = [f for f, _ in HamBuffers._fields_]
    fields_p = [f for f, _ in HamPeers._fields_]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "fmhr_b200.h"', 'int main(void){',
             'printf("%zu %zu %zu\\n", sizeof(fmhr_ham_config), sizeof(fmhr_ham_buffers), sizeof(fmhr_ham_peers));']
    lines += ['printf("%%zu\\n", offsetof(fmhr_ham_config, %s));' % f for f in fields_c]
    lines += ['printf("%%zu\\n", offsetof(fmhr_ham_buffers, %s));' % f for f in fields_b]
    lines += ['printf("%%zu\\n", offsetof(fmhr_ham_peers, %s));' % f for f in fields_p]
    lines += ['return 0;}']
    prog.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).split()
    assert int(out[0]) == ctypes.sizeof(HamConfig) and int(out[1]) == ctypes.sizeof(HamBuffers)
    assert int(out[2]) == ctypes.sizeof(HamPeers)
    offs = [int(x) for x in out[3:]]
    want = [getattr(HamConfig, f).offset for f in fields_c] + [getattr(HamBuffers, f).offset for f in fields_b] + \
           [getattr(HamPeers, f).offset for f in fields_p]
    assert offs == want


def test_host_only_queries(lib):
    from fmhr_b200._lib import HamConfig
    assert lib.fmhr_version() >= 100
    assert lib.fmhr_rasterize_workspace_bytes(2, 3, 4) == 2 * 3 * 4 * 8
    assert lib.fmhr_rasterize_workspace_bytes(0, 3, 4) == 0
    assert lib.fmhr_mesh_topology_workspace_bytes(100, 200) > 6 * 200 * 8 * 3
    cfg = HamConfig()
    cfg.V, cfg.T, cfg.H, cfg.W, cfg.n_views, cfg.n_views_global, cfg.phase, cfg.n_sh_rows = 1000, 2000, 64, 48, 3, 3, 1, 3
    b1 = lib.fmhr_ham_workspace_bytes(ctypes.byref(cfg))
    assert b1 >= 3 * 64 * 48 * (16 + 32) and lib.fmhr_ham_packed_floats(ctypes.byref(cfg)) == 12 * 1000 + 4
    cfg.phase = 0
    assert lib.fmhr_ham_workspace_bytes(ctypes.byref(cfg)) > b1  # phase A antialiases six channels: two more planes
    cfg.n_views = 0
    assert lib.fmhr_ham_workspace_bytes(ctypes.byref(cfg)) == 0
    # n_views_capacity: the workspace is laid out for the capacity, whatever the step's batch size (steps of different sizes
    # share one layout); a capacity below the batch is refused
    cfg.phase, cfg.n_views, cfg.n_views_global = 1, 3, 3
    cfg.n_views_capacity = 3
    assert lib.fmhr_ham_workspace_bytes(ctypes.byref(cfg)) == b1
    cfg.n_views_capacity = 7
    b7 = lib.fmhr_ham_workspace_bytes(ctypes.byref(cfg))
    assert b7 > b1
    for nv in (1, 2, 5, 7):
        cfg.n_views = cfg.n_views_global = nv
        assert lib.fmhr_ham_workspace_bytes(ctypes.byref(cfg)) == b7
    cfg.n_views = cfg.n_views_global = 8
    assert lib.fmhr_ham_workspace_bytes(ctypes.byref(cfg)) == 0


def test_bad_arguments_return_codes_not_crashes(lib):
    # null pointers are rejected before any CUDA call is made
    rc = lib.fmhr_rasterize_fwd(None, None, 1, 1, 1, 8, 8, None, None, None, 0, None)
    assert rc == -1 and b"invalid argument" in lib.fmhr_last_error_string()
    rc = lib.fmhr_interpolate_fwd(None, None, None, 1, 1, 1, 1, 8, 8, 3, None, None)
    assert rc == -1
    # entry points added for the meshlet rasteriser, the peer exchange and the host-batch pipeline
    assert lib.fmhr_rasterize_tile_words(2, 33, 17) == 2 * 1      # 3 x 2 tiles of 16 x 16 -> one 32-bit word per view
    assert lib.fmhr_rasterize_tile_words(3, 512, 334) == 3 * 21   # 32 x 21 = 672 tiles -> 21 words
    assert lib.fmhr_rasterize_tile_words(0, 8, 8) == 0
    rc = lib.fmhr_rasterize_fwd_meshlets(None, None, 1, 1, 1, 8, 8, None, None, None, None, None, 1, 1024, 1, None, 0, 0,
                                         None, None)
    assert rc == -1
    assert lib.fmhr_peer_alloc(0, None, None) == -1 and lib.fmhr_peer_open(None, None) == -1
    assert lib.fmhr_peer_close(None) == -1 and lib.fmhr_peer_free(None) == -1
    assert lib.fmhr_ham_step_update_peer(None, None, None, None) == -1
    assert lib.fmhr_ham_host_u8_submit(None, None, None, None) == -1
    assert lib.fmhr_ham_host_u8_submit_boxes(None, None, None, None, None, None) == -1
    assert lib.fmhr_ham_host_u8_submit_boxes_direct(None, None, None, None, None, None, None, None, None, None, 0, None) == -1
    assert lib.fmhr_ham_step_host_u8_acquire(None, None) == -1 and lib.fmhr_ham_step_host_u8_release(None, None) == -1
    assert lib.fmhr_ham_step_host_u8_body(None, None, None, None, None, None, None, None) == -1


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under fmhr_b200/ or nvdiffrast/ may import or execute it."""
    bad = []
    for pkg in ("fmhr_b200", "nvdiffrast"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, pkg)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    txt = open(os.path.join(dirpath, f)).read()
                    if re.search(r"^\s*(from|import)\s+oracle\b|oracle/|liboracle", txt, flags=re.M):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_ops_fail_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fmhr_b200 import dr, utils
    with pytest.raises(RuntimeError):
        dr.RasterizeGLContext()
    with pytest.raises(RuntimeError):
        utils.get_normals(torch.zeros(1, 3, 3), torch.zeros(1, 3, dtype=torch.int64))
    with pytest.raises(RuntimeError):
        utils.laplacian_smoothing(torch.zeros(3, 3), torch.zeros(1, 3, dtype=torch.int64))
