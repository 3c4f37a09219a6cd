"""Which initialisation path makes a fresh trainer's first step differ from the next fresh trainer's?  (diag for layer_7b)"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from datatunerx_b200 import lib as L  # noqa: E402
from oracle import llama_lora as O  # noqa: E402


def main(B=2, S=2048):
    ocfg = O.OracleConfig(vocab=32000, hidden=4096, n_layers=1, n_heads=32, ffn=11008, lora_r=16, lora_alpha=32.0, lr=0.0, total_steps=100)
    mc = L.ModelConfig(vocab=32000, hidden=4096, n_layers=1, n_heads=32, ffn=11008)
    tc = L.TrainConfig(micro_batch=B, seq_len=S, total_steps=100, lora_r=16, lora_alpha=32.0, lora_dropout=0.0, lr=0.0)
    w, lora = O.init_base_weights(ocfg, 1234), O.init_lora(ocfg, 4321)
    g = torch.Generator().manual_seed(99)
    for k in lora:
        if "lora_B" in k:
            lora[k] = torch.randn(lora[k].shape, generator=g) * 0.01
    wn = {k: v.numpy() for k, v in w.items()}
    ln = {k: v.numpy() for k, v in lora.items()}
    ids, labels = O.synthetic_batch(0, 0, B, S, ocfg.vocab)
    for mode in ("dev_w+init_lora", "host_w+host_lora", "dev_w+host_lora", "host_w+init_lora", "host_w+host_lora"):
        for rep in range(3):
            tr = L.Trainer(mc, tc)
            if mode.startswith("dev_w"):
                tr.init_random_weights(1234)
            else:
                tr.load_state_dict(wn)
            if mode.endswith("init_lora"):
                tr.init_lora(4321)
            else:
                tr.load_state_dict(ln)
            ev = tr.eval_loss(ids, labels)
            out = []
            for step in range(2):
                loss, gn, _, _ = tr.step(ids, labels)
                gr = tr.export_adapter(grads=True)
                out.append([loss, gn, hashlib.md5(b"".join(gr[k].tobytes() for k in sorted(gr))).hexdigest()[:8]])
            tr.close()
            print("LOAD_PATHS " + json.dumps({"mode": mode, "rep": rep, "eval": ev, "steps": out}), flush=True)


if __name__ == "__main__":
    main()
