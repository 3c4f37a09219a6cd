"""Diagnostic for tests/gpu_checks.py:check_layer_7b_shape: per-tensor gradient errors, run-to-run determinism, and the effect of
the forward softmax's FMA-pipe exp2 fraction."""
import json
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from datatunerx_b200 import lib as L  # noqa: E402
from oracle import llama_lora as O  # noqa: E402


def main(B=2, S=2048):
    ocfg = O.OracleConfig(vocab=32000, hidden=4096, n_layers=1, n_heads=32, ffn=11008, lora_r=16, lora_alpha=32.0, lr=1e-4, total_steps=100)
    mc = L.ModelConfig(vocab=32000, hidden=4096, n_layers=1, n_heads=32, ffn=11008)
    tc = L.TrainConfig(micro_batch=B, seq_len=S, total_steps=100, lora_r=16, lora_alpha=32.0, lora_dropout=0.0, lr=1e-4)
    w, lora = O.init_base_weights(ocfg, 1234), O.init_lora(ocfg, 4321)
    g = torch.Generator().manual_seed(99)
    for k in lora:
        if "lora_B" in k:
            lora[k] = torch.randn(lora[k].shape, generator=g) * 0.01
    ids, labels = O.synthetic_batch(0, 0, B, S, ocfg.vocab)
    orc = O.OracleTrainer(ocfg, w, lora)
    ref_loss, g_ref = orc.loss_and_grads(ids, labels)
    prev = None
    for every in (3, 3, 0, 0):
        L.set_option("attn_fwd_exp_fma_every", every)
        tr = L.Trainer(mc, tc)
        tr.load_state_dict({k: v.numpy() for k, v in w.items()})
        tr.load_state_dict({k: v.numpy() for k, v in lora.items()})
        loss, gn, _, _ = tr.step(ids, labels)
        got = tr.export_adapter(grads=True)
        tr.close()
        row = {"exp_fma_every": every, "loss_rel": abs(loss - ref_loss) / ref_loss, "gnorm": gn}
        for k, v in got.items():
            r = g_ref[k.replace("base_model.model.", "")].numpy()
            name = k.split("layers.0.self_attn.")[1].replace(".weight", "")
            row[name] = {"ref_norm": float(np.linalg.norm(r)), "err": float(np.linalg.norm(v - r) / np.linalg.norm(r)),
                         "nan": int(np.isnan(v).sum())}
            if "v_proj.lora_A" in k:  # per adapter row (16 rows of [r, d])
                row["v_A_rows"] = [round(float(np.linalg.norm(v[i] - r[i]) / max(np.linalg.norm(r[i]), 1e-20)), 4) for i in range(v.shape[0])]
        sig = {k: v.tobytes() for k, v in got.items()}
        row["same_as_previous_run"] = (prev == sig) if prev is not None else None
        prev = sig
        print("DIAG7B " + json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
