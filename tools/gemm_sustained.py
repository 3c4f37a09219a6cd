"""Sustained (power-capped) throughput of the CTA-pair GEMM at the dominant shape, per rasterisation group size.

The bench's roofline line times 10 launches (burst clocks); a training step runs the tensor pipe for ~0.3 s at a time, where
the 1 kW cap sets the clock.  usage: python tools/gemm_sustained.py [seconds_per_setting]
"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from datatunerx_b200 import lib as L  # noqa: E402

lib = L.load()
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 1.5
M, N, K = 16384, 22016, 4096
A = (torch.randn(M, K, device="cuda") * 0.05).to(torch.bfloat16)
B = (torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16)
Cm = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch():
    L.check(lib.dtx_gemm_bf16(C.c_void_p(A.data_ptr()), K, 0, C.c_void_p(B.data_ptr()), K, 0, None, 0, None, 0, 0,
                              C.c_void_p(Cm.data_ptr()), N, None, 0, M, N, K, 0, 1, 0, stream))


for gm in (16, 4, 8, 32, 64, 16):
    L.set_option("gemm_group_m", gm)
    for _ in range(5):
        launch()
    torch.cuda.synchronize()
    iters = int(secs / 0.002)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        launch()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    print(f"group_m={gm:3d}: {ms * 1000:7.1f} us/launch sustained over {iters} launches = {2.0 * M * N * K / ms / 1e9:7.1f} TFLOP/s", flush=True)
    time.sleep(0.2)
