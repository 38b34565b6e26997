"""TEST INFRASTRUCTURE: builds the CPU oracle (oracle/librbox_oracle.so, gcc) and, when /root/reference is
present (this container; never the GPU box), oracle/_ref from the reference's own sources (build_ref.sh).
Called by __graft_entry__.build(); building the checker is not using it."""
import hashlib
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(REPO, "build")


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


def build_oracle(verbose=True):
    """gcc build of the CPU restatement (test infrastructure) and, when /root/reference is present, of
    oracle/_ref (the reference's own device code; see oracle/build_ref.sh)."""
    odir = os.path.join(REPO, "oracle")
    src = os.path.join(odir, "rbox_oracle.c")
    lib = os.path.join(odir, "librbox_oracle.so")
    os.makedirs(BUILD, exist_ok=True)
    dig = _digest([src])
    stamp = os.path.join(BUILD, "rbox_oracle.sha")
    if not (os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read() == dig):
        o1, o2 = os.path.join(BUILD, "rbox_oracle.o"), os.path.join(BUILD, "rbox_oracle_fma.o")
        base = ["gcc", "-O2", "-fPIC", "-ffp-contract=off", "-std=c11", "-D_GNU_SOURCE"]
        _run(base + ["-c", src, "-o", o1], verbose)
        _run(base + ["-mfma", "-DORC_FMA", "-c", src, "-o", o2], verbose)
        _run(["gcc", "-shared", "-o", lib, o1, o2, "-lm"], verbose)
        with open(stamp, "w") as f:
            f.write(dig)
    if os.path.isdir("/root/reference"):
        ref_dir = os.path.join(odir, "_ref")
        need = not all(os.path.exists(os.path.join(ref_dir, f)) for f in
                       ("libref_rnms_host.so", "libref_rnms_cuda.so"))
        wraps = [os.path.join(odir, f) for f in ("ref_host_wrap.cpp", "ref_cuda_wrap.cu", "build_ref.sh")]
        rdig = _digest(wraps)
        rstamp = os.path.join(BUILD, "ref.sha")
        if need or not (os.path.exists(rstamp) and open(rstamp).read() == rdig):
            _run(["bash", os.path.join(odir, "build_ref.sh")], verbose)
            with open(rstamp, "w") as f:
                f.write(rdig)
    return lib



if __name__ == "__main__":
    build_oracle()
