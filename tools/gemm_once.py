"""Launch the bench's dominant GEMM (16384 x 22016 x 4096, the gate|up projection) a few times.

Meant to run under ncu:  ncu --set full --clock-control none -k regex:gemm2_kernel -s 3 -c 1 python tools/gemm_once.py
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from datatunerx_b200 import lib as L  # noqa: E402

lib = L.load()
M, N, K = 16384, 22016, 4096
A = (torch.randn(M, K, device="cuda") * 0.05).to(torch.bfloat16)
B = (torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16)
Cm = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(5):
    L.check(lib.dtx_gemm_bf16(C.c_void_p(A.data_ptr()), K, 0, C.c_void_p(B.data_ptr()), K, 0, None, 0, None, 0, 0,
                              C.c_void_p(Cm.data_ptr()), N, None, 0, M, N, K, 0, 1, 0, stream))
torch.cuda.synchronize()
print("ok")
