"""Writes tests/golden/tiny_trace_100.json: the oracle's 100-step loss / grad-norm / lr trace on the tiny-Llama parity
config (seeded weights, seeded synthetic batches).  The GPU parity test compares the native worker with the live oracle AND
with this committed trace (which pins the oracle itself against silent drift of torch / the restatement)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import llama_lora as O  # noqa: E402


def run(steps=100):
    cfg = O.OracleConfig(vocab=2048, hidden=256, n_layers=2, n_heads=2, ffn=768, lora_r=16, lora_alpha=32.0, lr=1e-3,
                         total_steps=steps)
    w, lora = O.init_base_weights(cfg, 1234), O.init_lora(cfg, 4321)
    tr = O.OracleTrainer(cfg, w, lora)
    out = []
    for s in range(steps):
        log = tr.step([O.synthetic_batch(s, 0, 2, 256, cfg.vocab)])
        out.append([log.loss, log.grad_norm, log.lr])
    return cfg, out


if __name__ == "__main__":
    cfg, out = run()
    json.dump({"generator": "oracle/make_golden_trace.py", "config": {k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.__dict__.items()},
               "batch": 2, "seq_len": 256, "weights_seed": 1234, "lora_seed": 4321, "trace_loss_gradnorm_lr": out,
               "window_mean_91_100": float(np.mean([o[0] for o in out[90:]]))},
              open(os.path.join(ROOT, "tests", "golden", "tiny_trace_100.json"), "w"))
    print("loss[0], loss[99], window:", out[0][0], out[99][0], float(np.mean([o[0] for o in out[90:]])))
