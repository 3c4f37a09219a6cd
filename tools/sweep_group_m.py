"""Sweep the pair-GEMM rasterisation group height on the step's main GEMM shapes (TF/s, CUDA events)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from datatunerx_b200 import lib as L
lib = L.load()
def P(t): return C.c_void_p(t.data_ptr())
shapes = [("qkv NT", 16384, 12288, 4096, 0), ("gate|up NT", 16384, 22016, 4096, 0), ("dh2 NN", 16384, 4096, 22016, 1), ("dact NN", 16384, 11008, 4096, 1)]
for name, M, N, K, b_mn in shapes:
    A = (torch.randn(M, K, device="cuda") * 0.05).bfloat16()
    B = (torch.randn((K, N) if b_mn else (N, K), device="cuda") * 0.02).bfloat16()
    Cm = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    out = []
    for g in (2, 4, 8, 16, 32, 64):
        L.set_option("gemm_group_m", g)
        def run():
            L.check(lib.dtx_gemm_bf16(P(A), K, 0, P(B), B.stride(0), b_mn, None, 0, None, 0, 0, P(Cm), N, None, 0, M, N, K, 0, 1, 0, None))
        for _ in range(2): run()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); s.record()
        for _ in range(8): run()
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 8
        out.append((g, round(2.0 * M * N * K / ms / 1e9)))
    print(name, out, flush=True)
