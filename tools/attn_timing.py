"""Phase timing of the attention kernels at the benchmark geometry (clock64() instrumentation, DTX_ATTN_TIMING build).

  make -C datatunerx_b200/csrc timing && DTX_LIB_PATH=datatunerx_b200/libdtxtune_timing.so python tools/attn_timing.py
Prints, for three CTAs of each kernel, the per-block cycle split of the compute warps (wait / tcgen05.ld / math / st+arrive)
and of the MMA issuer (waits, issue times), plus loop / drain / epilogue totals, and the kernels' durations (CUDA events).
"""
import ctypes as C
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from datatunerx_b200 import lib as L  # noqa: E402


def main(B=8, S=2048, H=32, Hkv=32):
    lib = L.load()
    D = 128
    W = (H + 2 * Hkv) * D
    g = torch.Generator(device="cpu").manual_seed(1)
    qkv = torch.randn(B * S, W, generator=g).to(torch.bfloat16).cuda()
    dout = torch.randn(B * S, H * D, generator=g).to(torch.bfloat16).cuda()
    out = torch.empty(B * S, H * D, dtype=torch.bfloat16, device="cuda")
    lse2 = torch.empty(B, H, S, dtype=torch.float32, device="cuda")
    delta = torch.empty(B, H, S, dtype=torch.float32, device="cuda")
    dqkv = torch.empty(B * S, W, dtype=torch.bfloat16, device="cuda")
    P = lambda t: C.c_void_p(t.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    sc = 1.0 / math.sqrt(D)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for every in (0, 4, 3, 2):  # forward softmax: fraction of exponentials on the FMA pipe
        L.set_option("attn_fwd_exp_fma_every", every)
        ts = []
        for it in range(4):
            ev[0].record()
            L.check(lib.dtx_attn_fwd(P(qkv), P(out), P(lse2), B, S, H, Hkv, sc, None, 0, st))
            ev[1].record()
            torch.cuda.synchronize()
            ts.append(ev[0].elapsed_time(ev[1]) * 1000)
        print(f"ATTN_TIMING fwd exp_fma_every={every}: {min(ts[1:]):.1f} us (runs {[round(t, 1) for t in ts]})", flush=True)
    L.set_option("attn_fwd_exp_fma_every", int(os.environ.get("DTX_FWD_EXP_FMA", "3")))
    for ring in (0, 4, 3, 0):  # dQ kernel: fraction of the exponentials on the FMA pipe
        L.set_option("attn_dq_exp_fma_every", ring)
        ts = []
        for it in range(4):
            ev[1].record()
            L.check(lib.dtx_attn_bwd(P(qkv), P(out), P(dout), P(lse2), P(delta), P(dqkv), B, S, H, Hkv, sc, None, 0, None, 0, st))
            ev[2].record()
            torch.cuda.synchronize()
            ts.append(ev[1].elapsed_time(ev[2]) * 1000)
        print(f"ATTN_TIMING bwd dq_exp_fma_every={ring}: {min(ts[1:]):.1f} us (runs {[round(t, 1) for t in ts]})", flush=True)
    L.set_option("attn_dq_exp_fma_every", int(os.environ.get("DTX_DQ_EXP_FMA", "0")))
    for it in range(2):
        ev[0].record()
        L.check(lib.dtx_attn_fwd(P(qkv), P(out), P(lse2), B, S, H, Hkv, sc, None, 0, st))
        ev[1].record()
        L.check(lib.dtx_attn_bwd(P(qkv), P(out), P(dout), P(lse2), P(delta), P(dqkv), B, S, H, Hkv, sc, None, 0, None, 0, st))
        ev[2].record()
        torch.cuda.synchronize()
        print(f"ATTN_TIMING iter {it}: fwd {ev[0].elapsed_time(ev[1]) * 1000:.1f} us  bwd {ev[1].elapsed_time(ev[2]) * 1000:.1f} us", flush=True)


if __name__ == "__main__":
    main()
