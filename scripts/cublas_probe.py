"""Does the cuBLAS this process gets support FP32 emulation on the bf16 tensor cores, and what does it buy on the
backward GEMM shapes?  (ctypes on libcublas.so.12, torch only for device memory and events.)"""
import ctypes
import os
import sys
import torch

os.environ.setdefault("CUBLAS_EMULATION_STRATEGY", sys.argv[1] if len(sys.argv) > 1 else "eager")
lib = ctypes.CDLL("libcublas.so.12")
v = ctypes.c_int()
ver = []
for k in range(3):
    lib.cublasGetProperty(k, ctypes.byref(v)); ver.append(v.value)
print("cublas", ver, "has cublasSetEmulationStrategy:", hasattr(lib, "cublasSetEmulationStrategy"))
h = ctypes.c_void_p()
assert lib.cublasCreate_v2(ctypes.byref(h)) == 0
if hasattr(lib, "cublasSetEmulationStrategy"):
    print("set strategy eager ->", lib.cublasSetEmulationStrategy(h, 2))
lib.cublasSetStream_v2(h, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
CUDA_R_32F, OP_N, OP_T = 0, 0, 1
one, zero = ctypes.c_float(1.0), ctypes.c_float(0.0)
for (m, n, k, ta, tb, tag) in [(512, 97280, 448, OP_N, OP_N, "gate pre-activations (nn)"), (320, 5120, 512, OP_N, OP_N, "dX (nn)"),
                              (1024, 5120, 256, OP_N, OP_N, "d hidden1 (nn)"), (1024, 256, 5120, OP_N, OP_T, "dW2 (tn)"),
                              (320, 512, 97280, OP_N, OP_T, "dW_ih (tn)")]:
    A = torch.randn(m * k, device="cuda"); B = torch.randn(k * n, device="cuda"); C = torch.zeros(m * n, device="cuda")
    lda = m if ta == OP_N else k
    ldb = k if tb == OP_N else n
    res = {}
    for name, ct in (("fp32", 68), ("emulated_bf16x9", 78), ("fast_16bf", 75)):
        outs = []
        rc = 0
        for it in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = lib.cublasGemmEx(h, ta, tb, m, n, k, ctypes.byref(one), ctypes.c_void_p(A.data_ptr()), CUDA_R_32F, lda,
                                  ctypes.c_void_p(B.data_ptr()), CUDA_R_32F, ldb, ctypes.byref(zero), ctypes.c_void_p(C.data_ptr()),
                                  CUDA_R_32F, m, ct, -1)
            e1.record(); torch.cuda.synchronize()
            outs.append(e0.elapsed_time(e1))
        res[name] = (rc, min(outs), C.clone())
    ref = res["fp32"][2]
    print("%-28s m %5d n %6d k %6d | " % (tag, m, n, k) + " | ".join(
        "%s rc %d %.3f ms (%.0f TF) err %.1e" % (nm, r[0], r[1], 2.0 * m * n * k / r[1] / 1e9, float((r[2] - ref).abs().max() / ref.abs().max()))
        for nm, r in res.items()))
