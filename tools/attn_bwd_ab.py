"""Time the backward attention kernels (delta + dQ + dK/dV) at a given shape with an option switched off / on.

usage: python tools/attn_bwd_ab.py [option [B H S [Hkv]]]   (default: attn_bwd_warps16 at the Llama-2-7B bench shape 8 32 2048)
"""
import ctypes as C
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from datatunerx_b200 import lib as L  # noqa: E402

opt = sys.argv[1] if len(sys.argv) > 1 else "attn_bwd_warps16"
B, H, S = (int(x) for x in sys.argv[2:5]) if len(sys.argv) >= 5 else (8, 32, 2048)
Hkv = int(sys.argv[5]) if len(sys.argv) > 5 else H
D = 128
lib = L.load()
torch.manual_seed(0)
W = (H + 2 * Hkv) * D
qkv = (torch.randn(B * S, W, device="cuda") * 0.5).to(torch.bfloat16)
dout = (torch.randn(B * S, H * D, device="cuda") * 0.1).to(torch.bfloat16)
out = torch.empty(B * S, H * D, dtype=torch.bfloat16, device="cuda")
lse = torch.empty(B, H, S, dtype=torch.float32, device="cuda")
delta = torch.empty(B, H, S, dtype=torch.float32, device="cuda")
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: C.c_void_p(t.data_ptr())
scale = 1.0 / math.sqrt(D)
L.check(lib.dtx_attn_fwd(P(qkv), P(out), P(lse), B, S, H, Hkv, scale, stream))
res = {}
for mode in (0, 1, 0, 1):
    L.set_option(opt, mode)
    dqkv = torch.zeros(B * S, W, dtype=torch.bfloat16, device="cuda")
    for it in range(2):
        L.check(lib.dtx_attn_bwd(P(qkv), P(out), P(dout), P(lse), P(delta), P(dqkv), B, S, H, Hkv, scale, stream))
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for it in range(5):
        L.check(lib.dtx_attn_bwd(P(qkv), P(out), P(dout), P(lse), P(delta), P(dqkv), B, S, H, Hkv, scale, stream))
    e.record()
    torch.cuda.synchronize()
    print(f"{opt}={mode}: {s.elapsed_time(e) / 5 * 1000:.1f} us per backward (delta + dq + dkv)")
    res[mode] = dqkv.float()
d = (res[0] - res[1]).norm() / res[0].norm()
print(f"{opt} 0 vs 1: dqkv rel diff {float(d):.3e}")
