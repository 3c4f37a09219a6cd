import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from datatunerx_b200 import lib as L
from oracle import llama_lora as O
for vocab, B in ((400, 4), (400, 2), (2048, 4), (512, 4)):
    ocfg = O.OracleConfig(vocab=vocab, hidden=256, n_layers=2, n_heads=2, ffn=768, lora_r=16, lora_alpha=32.0, lr=1e-3, total_steps=5)
    mc = L.ModelConfig(vocab=vocab, hidden=256, n_layers=2, n_heads=2, ffn=768)
    tc = L.TrainConfig(micro_batch=B, seq_len=256, total_steps=5, lora_r=16, lora_alpha=32.0, lora_dropout=0.0, lr=1e-3)
    w, lora = O.init_base_weights(ocfg, 7), O.init_lora(ocfg, 4321)
    tr = L.Trainer(mc, tc)
    tr.load_state_dict({k: v.numpy() for k, v in w.items()})
    tr.load_state_dict({k: v.numpy() for k, v in lora.items()})
    orc = O.OracleTrainer(ocfg, w, lora)
    ids, labels = O.synthetic_batch(0, 0, B, 256, vocab)
    # emulate right padding of the worker
    ids[:, 100:] = 2; labels[:, 100:] = -100
    print(vocab, B, "eval native/oracle", tr.eval_loss(ids, labels), orc.eval_loss(ids, labels), "step", tr.step(ids, labels)[:2], orc.step([(ids, labels)]).loss, flush=True)
    tr.close()
