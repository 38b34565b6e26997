"""Build helper: compiles csrc/*.cu for sm_100a with nvcc into rotate-yolov3_b200/libryolo.so (in-tree, so
that the shared object travels to the GPU box).

Used by ``__graft_entry__.build()``; nothing here runs at import time of the package."""
import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
BUILD = os.path.join(REPO, "build")
LIB = os.path.join(PKG_DIR, "libryolo.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]
# NOTE: no --use_fast_math anywhere: rnms.cu's arithmetic is pinned to the reference build (IEEE div/sqrt,
# full-precision sinf/cosf).  TUs that want fast intrinsics call them explicitly.


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (needed to build libryolo.so for sm_100a)")


def _digest(paths, extra=""):
    h = hashlib.sha256(extra.encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    return h.hexdigest()


def _run(cmd, verbose):
    if verbose:
        print("[build]", " ".join(cmd), flush=True)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build step failed: " + " ".join(cmd))
    return r.stdout


def build_cuda(verbose=True, force=False):
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(REPO, "include", "ryolo.h"))
    sources = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    objs = []
    relink = force or not os.path.exists(LIB)
    procs = []
    for src in sources:
        obj = os.path.join(BUILD, os.path.basename(src)[:-3] + ".o")
        stamp = obj + ".sha"
        dig = _digest([src] + headers, " ".join(NVCC_FLAGS))
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
            continue
        cmd = [nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj]
        if verbose:
            print("[build]", " ".join(cmd), flush=True)
        procs.append((subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True),
                      cmd, stamp, dig))
        relink = True
    for p, cmd, stamp, dig in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
        with open(stamp, "w") as f:
            f.write(dig)
    if relink:
        _run([nvcc, "-shared", "-o", LIB] + objs, verbose)
    return LIB


if __name__ == "__main__":
    build_cuda()
