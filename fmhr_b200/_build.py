"""Builds fmhr_b200/libfmhr_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libfmhr_b200.so")
SOURCES = ["api.cu", "raster.cu", "interpolate.cu", "antialias.cu", "mesh.cu", "meshlet.cu", "ncc_loop.cu", "ham.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source to an object (in parallel) and link the shared library."""
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "fmhr_b200.h"))
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    jobs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(objdir, s.replace(".cu", ".o"))
        if force or _stale(obj, [src] + headers):
            jobs.append((src, obj))

    def run(job):
        src, obj = job
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        return r.stderr

    with ThreadPoolExecutor(max_workers=8) as ex:
        logs = list(ex.map(run, jobs))
    objs = [os.path.join(objdir, s.replace(".cu", ".o")) for s in SOURCES]
    if jobs or not os.path.exists(SO):
        cmd = [nvcc, "-shared", "-o", SO] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    if verbose:
        for l in logs:
            sys.stderr.write(l)
    return SO


def build_variant(name, defines):
    """Diagnostic / tuning build of the same sources with extra -D flags -> <repo>/variants/libfmhr_<name>.so (git-ignored;
    selected with FMHR_B200_LIB).  Rebuilt only when a source is newer."""
    root = os.path.dirname(HERE)
    out = os.path.join(root, "variants", "libfmhr_%s.so" % name)
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(root, "include", "fmhr_b200.h"))
    if not _stale(out, srcs + headers):
        return out
    objdir = os.path.join(root, "variants", "obj_" + name)
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()

    def run(src):
        obj = os.path.join(objdir, os.path.basename(src).replace(".cu", ".o"))
        r = subprocess.run([nvcc] + NVCC_FLAGS + list(defines) + ["-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(run, srcs))
    r = subprocess.run([nvcc, "-shared", "-o", out] + objs + ["-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    for o in objs:
        os.remove(o)
    os.rmdir(objdir)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
