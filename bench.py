#!/usr/bin/env python
"""Benchmark of the hot path: tokens/sec of the Llama-2-7B LoRA (r=16, q_proj,v_proj) SFT step, seq 2048,
batch 8 per GPU, bf16, synthetic instruction pairs, random-init weights (BASELINE.json configs[1]).

  python bench.py --gpus N --steps K --warmup W            native arm (libdtxtune.so, one process per GPU)
  python bench.py --impl reference --gpus N ...             CPU arm: the oracle restatement of the reference step
                                                            on the host cores (the reference itself cannot run here)

A "step" = forward + backward + (NCCL all-reduce) + clip + AdamW on one batch of 8 x 2048 tokens per GPU.
`value`  : device-resident batches (dtx_step_device), wall clock over K steps between barriers, max over ranks.
`e2e`    : the same K steps through the public host API (pinned host int32 batches in, loss/grad-norm out).
`roofline`: the dominant kernel (tcgen05 GEMM at the gate|up projection shape 16384 x 22016 x 4096) timed alone with
            CUDA events against MEASURED_PEAKS.json's burst bf16 figure; `step_roofline` is the whole step against the
            sustained figure with the algorithmic 28.36 GFLOP/token of BASELINE.md.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_TOKEN = 28.36e9  # BASELINE.md §2 (fwd GEMM + bwd dX GEMM + causal attention fwd/bwd + LoRA; no recompute)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"burst": d.get("bf16_tflops", 1590.0), "sustained": d.get("bf16_tflops_sustained", 1400.0),
                "hbm": d.get("hbm_gbs", 6650.0), "src": "MEASURED_PEAKS.json"}
    return {"burst": 1590.0, "sustained": 1400.0, "hbm": 6650.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def pick_cpu_threads(max_threads: int) -> int:
    """Thread count at which torch's fp32 GEMM is fastest on this host (<= max_threads).  On the 128-thread GPU boxes the full
    thread count is several times SLOWER than a moderate one for the step's matrix shapes; the CPU arm should be the best the
    host cores can do, so the count is calibrated in ~2 s on the MLP GEMM shape of one sequence."""
    import torch
    cands = sorted({t for t in (8, 16, 24, 32, 48, 64, 96, 128, max_threads) if 1 <= t <= max_threads})
    a, b = torch.randn(2048, 4096), torch.randn(4096, 11008)
    best, best_t = cands[-1], float("inf")
    for t in cands:
        torch.set_num_threads(t)
        torch.mm(a, b)
        t0 = time.perf_counter()
        for _ in range(2):
            torch.mm(a, b)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = t, dt
    return best


def cpu_reference_sample(layers_sample: int, threads: int, steps: int = 1, warmup: int = 0):
    """Time the oracle (CPU restatement of the reference step) on a bounded sample of the 7B workload:
    one 2048-token sequence through `layers_sample` of the 32 identical Llama-2-7B decoder layers plus embedding,
    lm_head, CE, backward, clip and AdamW (fp32).  tokens/s is extrapolated to 32 layers from the measured split
    (time = head + per_layer * 32); both measured numbers are returned."""
    import torch
    from oracle import llama_lora as O
    torch.set_num_threads(threads)
    S = 2048

    def run(nl):
        cfg = O.OracleConfig.llama2_7b(lora_r=16, lora_alpha=32.0, lr=1e-4, total_steps=100)
        cfg.n_layers = nl
        g = torch.Generator().manual_seed(1)
        d, F, V = cfg.hidden, cfg.ffn, cfg.vocab
        w = {"model.embed_tokens.weight": torch.randn(V, d, generator=g) * 0.02, "lm_head.weight": torch.randn(V, d, generator=g) * 0.02,
             "model.norm.weight": torch.ones(d)}
        for l in range(nl):
            p = f"model.layers.{l}."
            for n, shp in (("self_attn.q_proj", (d, d)), ("self_attn.k_proj", (d, d)), ("self_attn.v_proj", (d, d)),
                           ("self_attn.o_proj", (d, d)), ("mlp.gate_proj", (F, d)), ("mlp.up_proj", (F, d)), ("mlp.down_proj", (d, F))):
                w[p + n + ".weight"] = torch.randn(*shp, generator=g) * 0.02
            w[p + "input_layernorm.weight"] = torch.ones(d)
            w[p + "post_attention_layernorm.weight"] = torch.ones(d)
        tr = O.OracleTrainer(cfg, w, O.init_lora(cfg, 4321))
        ts = []
        for s in range(warmup + steps):
            batch = O.synthetic_batch(s, 0, 1, S, V)
            t0 = time.perf_counter()
            tr.step([batch])
            if s >= warmup:
                ts.append(time.perf_counter() - t0)
        return float(np.mean(ts))

    t_a = run(1)
    t_b = run(layers_sample)
    per_layer = (t_b - t_a) / max(1, layers_sample - 1)
    head = max(0.0, t_a - per_layer)
    t_full = head + 32 * per_layer
    return {"tokens_per_s": S / t_full, "sec_1_layer": t_a, f"sec_{layers_sample}_layers": t_b, "sec_32_layers_extrapolated": t_full}


def run_reference(args, rank: int):
    if rank != 0:
        return
    import torch
    threads = pick_cpu_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    r = cpu_reference_sample(layers_sample=3, threads=threads, steps=max(1, min(args.steps, 2)), warmup=0)
    wall = time.perf_counter() - t0
    sample = ("oracle (fp32 torch CPU restatement of cmd/tuning/train.py:196-299; the reference worker itself needs ray/peft/"
              "deepspeed/CUDA and cannot run): 1 x 2048-token sequence, fwd+bwd+clip+AdamW through 1 and 3 Llama-2-7B decoder "
              "layers + embedding + lm_head + CE; per-layer time extrapolated to 32 layers")
    v = r["tokens_per_s"]
    line = {"impl": "reference", "metric": "tokens/sec Llama-2-7B LoRA SFT seq2048", "value": v, "unit": "tokens/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * 2048 / v,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "Llama-2-7B LoRA r=16 q_proj,v_proj, seq 2048 (CPU: batch 1 sequence per step, bounded sample)"},
            "cpu_baseline": {"value": v, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": sample, "detail": r,
                             "torch_threads": torch.get_num_threads(), "wall_s": wall},
            "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant GEMM at the bench shape, from the
# ncu --set full capture summarised in profiles/r01_ncu_dominant_gemm_final.txt (1.371 GB read + 0.706 GB write).
DOMINANT_GEMM_DRAM_BYTES = 2.0769e9


def time_dominant_gemm(torch, L):
    """CUDA-event timing of the tcgen05 GEMM at the largest per-layer shape (gate|up projection)."""
    import ctypes as C
    lib = L.load()
    M, N, K = 16384, 22016, 4096
    A = (torch.randn(M, K, device="cuda") * 0.05).to(torch.bfloat16)
    B = (torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16)
    Cm = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def launch():
        L.check(lib.dtx_gemm_bf16(C.c_void_p(A.data_ptr()), K, 0, C.c_void_p(B.data_ptr()), K, 0, None, 0, None, 0, 0,
                                  C.c_void_p(Cm.data_ptr()), N, None, 0, M, N, K, 0, 1, 0, stream))
    for _ in range(3):
        launch()
    iters = 10
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(iters):  # A+B+C = 1.04 GB per launch >> 126 MB L2: every launch streams from HBM
        launch()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    del A, B, Cm
    torch.cuda.empty_cache()
    return {"shape": [M, N, K], "ms": ms, "tflops": 2.0 * M * N * K / ms / 1e9}


def run_native(args, rank: int, local_rank: int, world: int):
    import torch  # device memory for the resident batches, gloo rendezvous and the clock; no torch compute
    from datatunerx_b200 import lib as L
    from datatunerx_b200.tuning.synthetic import synthetic_batch

    from datatunerx_b200.dist import Rendezvous
    rv = Rendezvous()
    dist = rv.dist
    torch.cuda.set_device(local_rank)

    lora_r, quant = 16, None
    if args.config == "7b":
        mc = L.ModelConfig.llama2_7b()
        B, S = 8, 2048
    elif args.config == "mistral7b_qlora":  # BASELINE.json configs[2]: Mistral-7B QLoRA nf4 r=32, seq 4096 (not the headline metric)
        mc = L.ModelConfig(vocab=32000, hidden=4096, n_layers=32, n_heads=32, n_kv_heads=8, ffn=14336, max_seq=32768)
        B, S, lora_r, quant = 4, 4096, 32, "int4"
    else:
        mc = L.ModelConfig(vocab=2048, hidden=256, n_layers=2, n_heads=2, ffn=768)
        B, S = 2, 256
    total = args.warmup + 2 * args.steps + 2
    tc = L.TrainConfig(micro_batch=B, seq_len=S, total_steps=max(total, 100), lora_r=lora_r, lora_alpha=32.0, lora_dropout=0.0, lr=1e-4)
    nccl_id = rv.broadcast_bytes(L.nccl_unique_id)
    tr = L.Trainer(mc, tc, device=local_rank, rank=rank, world=world, nccl_id=nccl_id)
    tr.init_random_weights(1234)
    if quant:
        tr.quantize_base(quant)
    tr.init_lora(4321)

    n_batches = 4
    host = [synthetic_batch(i, rank, B, S, mc.vocab) for i in range(n_batches)]
    pinned = [(torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()) for a, b in host]
    dev = [(a.cuda(non_blocking=False), b.cuda(non_blocking=False)) for a, b in pinned]
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        rv.barrier()

    max_over_ranks = rv.max_over_ranks

    losses = []
    for i in range(args.warmup):
        a, b = dev[i % n_batches]
        losses.append(tr.step_ptr(a.data_ptr(), b.data_ptr(), on_device=True)[0])

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- timed region 1: device-resident inputs ----
    launches0 = tr.launch_count
    barrier()
    t0 = time.perf_counter()
    dev_ms = []
    for i in range(args.steps):
        a, b = dev[i % n_batches]
        losses.append(tr.step_ptr(a.data_ptr(), b.data_ptr(), on_device=True)[0])
        dev_ms.append(tr.last_step_ms)
    barrier()
    dt = max_over_ranks(time.perf_counter() - t0)
    launches = tr.launch_count - launches0
    # ---- timed region 2: end to end through the host API (pinned host batches in, loss out) ----
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        a, b = pinned[i % n_batches]
        losses.append(tr.step_ptr(a.data_ptr(), b.data_ptr(), on_device=False)[0])
    barrier()
    dt_e2e = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop() if rank == 0 else None
    ev_ms = max_over_ranks(float(np.mean(dev_ms)))

    tokens_per_step = B * S * world
    value = tokens_per_step * args.steps / dt
    e2e = tokens_per_step * args.steps / dt_e2e
    if rank != 0:
        tr.close()
        return
    peaks = measured_peaks()
    gemm = time_dominant_gemm(torch, L) if args.config == "7b" else None
    tr.close()
    per_gpu_tflops = (value / world) * FLOP_PER_TOKEN / 1e12
    metric = {"7b": "tokens/sec Llama-2-7B LoRA SFT seq2048", "mistral7b_qlora": "tokens/sec Mistral-7B QLoRA nf4 r=32 seq4096",
              "tiny": "tokens/sec tiny-Llama smoke"}[args.config]
    line = {
        "metric": metric, "value": value, "unit": "tokens/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "Llama-2-7B (random-init N(0,0.02)) LoRA r=16 alpha=32 q_proj,v_proj, seq 2048, batch 8/GPU, "
                               "AdamW + clip 1.0 + linear schedule, bf16 compute / fp32 accumulate / fp32 adapters"
                   if args.config == "7b" else ("Mistral-7B shape (GQA 32/8, ffn 14336) QLoRA nf4 r=32, seq 4096, batch 4/GPU "
                                                "(BASELINE.json configs[2]; NOT the headline metric, FLOP/token differs)"
                                                if args.config == "mistral7b_qlora" else
                                                "tiny-Llama smoke config (NOT the benchmark workload)"),
                   "global_batch": B * world, "seq_len": S, "parallelism": f"dp{world}",
                   "l2": "per-step working set (13.5 GB weights + ~55 GB saved activations) >> 126 MB L2; no flush needed",
                   "recompute": "none (activations kept; the reference's gradient checkpointing is a memory knob, not math)"},
        "e2e": {"value": e2e, "unit": "tokens/s", "h2d_bytes_per_step": 2 * B * S * 4, "d2h_bytes_per_step": 8,
                "ms_per_step": 1000.0 * dt_e2e / args.steps},
        "gpu_launches": int(launches),
        "device_ms_per_step": ev_ms,
        "clocks": clocks,
        "step_roofline": ({"bound": "tensor", "achieved": per_gpu_tflops, "peak": peaks["sustained"], "unit": "TFLOP/s",
                           "frac": per_gpu_tflops / peaks["sustained"], "flop_per_token": FLOP_PER_TOKEN,
                           "peak_src": peaks["src"] + " bf16_tflops_sustained"} if args.config == "7b" else None),
        "loss_first_last": [losses[0], losses[-1]] if losses else None,
    }
    if gemm:
        line["roofline"] = {"bound": "tensor", "achieved": gemm["tflops"], "peak": peaks["burst"], "unit": "TFLOP/s",
                            "frac": gemm["tflops"] / peaks["burst"], "traffic": DOMINANT_GEMM_DRAM_BYTES, "traffic_src": "profiles/r01_ncu_dominant_gemm_final.txt (ncu --set full, dram read+write per launch; algorithmic 1.036e9)", "kernel": "gemm2_kernel<NT,bf16> (CTA-pair 256x256x64, cta_group::2)",
                            "shape_mnk": gemm["shape"], "ms": gemm["ms"], "peak_src": peaks["src"] + " bf16_tflops (burst)"}
    if args.cpu_baseline and args.config == "7b":
        threads = pick_cpu_threads(os.cpu_count() or 1)
        r = cpu_reference_sample(layers_sample=3, threads=threads, steps=1)
        line["cpu_baseline"] = {"value": r["tokens_per_s"], "unit": "tokens/s", "cores": threads, "kind": "port",
                                "sample": "oracle fp32: 1 x 2048-token sequence fwd+bwd+AdamW through 1 and 3 Llama-2-7B layers + "
                                          "embed + lm_head + CE, extrapolated to 32 layers", "detail": r}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="7b", choices=["7b", "mistral7b_qlora", "tiny"])
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.warmup < 3 and args.impl == "native":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        print(json.dumps({"error": f"--gpus {args.gpus} needs torchrun (WORLD_SIZE={world})"}))
        sys.exit(2)
    run_native(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
