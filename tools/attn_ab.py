"""A/B the two forward attention kernels (two-tile TMEM-accumulating vs one-tile) at a given shape.

usage: python tools/attn_ab.py [B H S [Hkv]]     (default: the Llama-2-7B bench shape 8 32 2048)
"""
import ctypes as C
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from datatunerx_b200 import lib as L  # noqa: E402

B, H, S = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (8, 32, 2048)
Hkv = int(sys.argv[4]) if len(sys.argv) > 4 else H
D = 128
lib = L.load()
torch.manual_seed(0)
qkv = (torch.randn(B * S, (H + 2 * Hkv) * D, device="cuda") * 0.5).to(torch.bfloat16)
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
res = {}
for mode in (0, 1):
    L.set_option("attn_fwd_two_tiles", mode)
    out = torch.full((B * S, H * D), float("nan"), dtype=torch.bfloat16, device="cuda")
    lse = torch.full((B, H, S), float("nan"), dtype=torch.float32, device="cuda")
    for it in range(3):
        L.check(lib.dtx_attn_fwd(C.c_void_p(qkv.data_ptr()), C.c_void_p(out.data_ptr()), C.c_void_p(lse.data_ptr()), B, S, H, Hkv,
                                 1.0 / math.sqrt(D), stream))
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for it in range(5):
        L.check(lib.dtx_attn_fwd(C.c_void_p(qkv.data_ptr()), C.c_void_p(out.data_ptr()), C.c_void_p(lse.data_ptr()), B, S, H, Hkv,
                                 1.0 / math.sqrt(D), stream))
    e.record()
    torch.cuda.synchronize()
    print(f"mode two_tiles={mode}: {s.elapsed_time(e) / 5 * 1000:.1f} us, nan out {int(torch.isnan(out.float()).sum())}")
    res[mode] = (out.float(), lse)
d = (res[0][0] - res[1][0]).norm() / res[0][0].norm()
dl = (res[0][1] - res[1][1]).abs().max()
print(f"two-tile vs one-tile: out rel {float(d):.3e}, lse max abs {float(dl):.3e}")
