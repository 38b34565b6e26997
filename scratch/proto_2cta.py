"""Build + check + time the CTA-pair GEMM prototype (scratch/proto_2cta_gemm.cu).  GPU box only:
    timeout 120 python scratch/proto_2cta.py            (wrap in a timeout: a protocol mistake in a cluster kernel hangs)"""
import ctypes, os, subprocess, sys
import torch
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_proto"); os.makedirs(OUT, exist_ok=True)
SO = os.path.join(OUT, "libproto2cta.so")


def build():
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr",
           "-Xcompiler", "-fPIC", "-shared", "-o", SO, os.path.join(HERE, "proto_2cta_gemm.cu")]
    subprocess.check_call(cmd)


def main():
    if "--build-only" in sys.argv or not os.path.exists(SO):
        build()
        if "--build-only" in sys.argv:
            return
    lib = ctypes.CDLL(SO)
    lib.proto_gemm.restype = ctypes.c_int
    lib.proto_gemm.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                               ctypes.c_int, ctypes.c_void_p]
    dev = torch.device("cuda")
    for (m, n, k) in ((256, 256, 128), (512, 512, 1152), (37888, 256, 1152), (18944, 256, 18432), (37888, 256, 18432),
                      (37888, 512, 9216)):
        a = torch.randn(m, k, device=dev).to(torch.bfloat16)
        b = torch.randn(n, k, device=dev).to(torch.bfloat16)
        want = (a[:4096].float() @ b.float().t()) if m > 4096 else a.float() @ b.float().t()
        for pair in (1, 2):
            c = torch.zeros(m, n, device=dev)
            st = lib.proto_gemm(a.data_ptr(), b.data_ptr(), c.data_ptr(), m, n, k, pair, None)
            torch.cuda.synchronize()
            err = float((c[:want.shape[0]] - want).abs().max()) / float(want.abs().max())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                lib.proto_gemm(a.data_ptr(), b.data_ptr(), c.data_ptr(), m, n, k, pair, None)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print("m=%d n=%d k=%d pair=%d status=%d rel_err=%.2e  %.1f us  %.0f TFLOP/s" % (m, n, k, pair, st, err, ms * 1e3,
                                                                                          2.0 * m * n * k / ms / 1e9))


if __name__ == "__main__":
    main()
