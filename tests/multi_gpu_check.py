"""N-rank data-parallel parity on real GPUs (launch: torchrun --nproc-per-node N tests/multi_gpu_check.py).

Every rank drives one dtx_trainer (world = N) on its own shard; the flat adapter gradient is all-reduced by NCCL inside
libdtxtune.  Checked against the CPU oracle run with world = N on the same shards: per-rank loss, global grad-norm,
and the adapters after `steps` optimizer steps (which must also be bitwise identical across ranks)."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from datatunerx_b200 import lib as L  # noqa: E402
from datatunerx_b200.dist import Rendezvous  # noqa: E402
from oracle import llama_lora as O  # noqa: E402


def main_full(steps=4):
    """Full-parameter SFT across N ranks: per-layer reduce-scatter of bf16 gradients, sharded fp32 AdamW, all-gather of the
    updated bf16 weights - against the oracle with world = N and every weight trainable; replicas must end bitwise equal."""
    rv = Rendezvous()
    ocfg = O.OracleConfig(vocab=2048, hidden=256, n_layers=2, n_heads=2, ffn=768, lr=1e-3, total_steps=steps, full_finetune=True,
                          weight_decay=0.01)
    mc = L.ModelConfig(vocab=2048, hidden=256, n_layers=2, n_heads=2, ffn=768)
    tc = L.TrainConfig(micro_batch=2, seq_len=256, total_steps=steps, lr=1e-3, full_finetune=True, weight_decay=0.01, lora_dropout=0.0)
    w = O.init_base_weights(ocfg, 1234)
    nccl_id = rv.broadcast_bytes(L.nccl_unique_id)
    tr = L.Trainer(mc, tc, device=rv.local_rank, rank=rv.rank, world=rv.world, nccl_id=nccl_id)
    tr.load_state_dict({k: v.numpy() for k, v in w.items()})
    orc = O.OracleTrainer(ocfg, w, {}, world=rv.world)
    worst_l = worst_g = 0.0
    for s in range(steps):
        shards = [O.synthetic_batch(s, r, 2, 256, ocfg.vocab) for r in range(rv.world)]
        ref_losses = [orc.eval_loss(*b) for b in shards]
        ref = orc.step(shards)
        loss, gn, lr, stepped = tr.step(*shards[rv.rank])
        assert stepped
        worst_l = max(worst_l, abs(loss - ref_losses[rv.rank]) / ref_losses[rv.rank])
        worst_g = max(worst_g, abs(gn - ref.grad_norm) / ref.grad_norm)
    wts = tr.export_weights()
    f32 = lambda a: (a.astype(np.uint32) << 16).view(np.float32)
    drift = max(float(np.linalg.norm(f32(v) - orc.lora[k].detach().numpy()) / max(np.linalg.norm(orc.lora[k].detach().numpy()), 1e-12))
                for k, v in wts.items())
    digest = hashlib.sha256(b"".join(wts[k].tobytes() for k in sorted(wts))).hexdigest()
    digests = [None] * rv.world
    if rv.dist is not None:
        rv.dist.all_gather_object(digests, digest)
    else:
        digests = [digest]
    tr.close()
    ok = worst_l < 2e-3 and worst_g < 3e-2 and drift < 1e-2 and len(set(digests)) == 1
    if rv.rank == 0:
        print("MULTI_GPU_CHECK_FULL " + json.dumps({"world": rv.world, "ok": ok, "loss_rel": worst_l, "gnorm_rel": worst_g,
                                                    "weight_drift": drift, "replicas_bitwise_equal": len(set(digests)) == 1}), flush=True)
    rv.close()
    sys.exit(0 if ok else 1)


def main(steps=5, ragged=False, packed=False):
    """ragged: every rank's micro-batch has its own row lengths and runs as length groups ("varlen_split" = 2 cuts wherever the
    128-rounded lengths differ): even ranks split into two groups, odd ranks stay one - the all-reduce still happens once per step."""
    ragged = ragged or packed  # packed: the same ragged shards in the default execution mode (sequences back to back, one pass)
    rv = Rendezvous()
    S = 512 if ragged else 256
    ocfg = O.OracleConfig(vocab=2048, hidden=256, n_layers=2, n_heads=2, ffn=768, lora_r=16, lora_alpha=32.0, lr=1e-3,
                          total_steps=steps)
    mc = L.ModelConfig(vocab=2048, hidden=256, n_layers=2, n_heads=2, ffn=768)
    tc = L.TrainConfig(micro_batch=2, seq_len=S, total_steps=steps, lora_r=16, lora_alpha=32.0, lora_dropout=0.0, lr=1e-3)
    if ragged and not packed:
        L.set_option("varlen_split", 2)

    def shard(s, r):
        ids, labels = O.synthetic_batch(s, r, 2, S, ocfg.vocab)
        if not ragged:
            return ids, labels, None
        lens = np.array([S, 90 + 40 * r + 7 * s] if r % 2 == 0 else [S, S - 5 - r], dtype=np.int32)
        ids[1, lens[1]:] = 0
        labels[1, lens[1]:] = -100
        return ids, labels, lens

    w, lora = O.init_base_weights(ocfg, 1234), O.init_lora(ocfg, 4321)
    nccl_id = rv.broadcast_bytes(L.nccl_unique_id)
    tr = L.Trainer(mc, tc, device=rv.local_rank, rank=rv.rank, world=rv.world, nccl_id=nccl_id)
    tr.load_state_dict({k: v.numpy() for k, v in w.items()})
    tr.load_state_dict({k: v.numpy() for k, v in lora.items()})
    orc = O.OracleTrainer(ocfg, w, lora, world=rv.world)
    worst_l = worst_g = 0.0
    groups = []
    for s in range(steps):
        shards = [shard(s, r) for r in range(rv.world)]
        ref_losses = [orc.eval_loss(b[0], b[1]) for b in shards]
        ref = orc.step([(b[0], b[1]) for b in shards])
        loss, gn, lr, stepped = tr.step(*shards[rv.rank])
        assert stepped
        groups.append(tr.last_step_groups)
        worst_l = max(worst_l, abs(loss - ref_losses[rv.rank]) / ref_losses[rv.rank])
        worst_g = max(worst_g, abs(gn - ref.grad_norm) / ref.grad_norm)
    if ragged:  # even ranks: two length groups (or packed = 0); odd ranks: both rows round to the full length, one pass
        assert groups == [(0 if packed else 2) if rv.rank % 2 == 0 else 1] * steps, (rv.rank, groups)
    ad = tr.export_adapter()
    ref_ad = orc.state_dict()
    drift = max(float(np.linalg.norm(v - ref_ad[k.replace("base_model.model.", "")]) /
                      max(np.linalg.norm(ref_ad[k.replace("base_model.model.", "")]), 1e-12)) for k, v in ad.items())
    digest = hashlib.sha256(b"".join(ad[k].tobytes() for k in sorted(ad))).hexdigest()
    digests = [None] * rv.world
    if rv.dist is not None:
        rv.dist.all_gather_object(digests, digest)
    else:
        digests = [digest]
    tr.close()
    # adapters after a few Adam steps: Adam normalises tiny bf16-noisy gradients, so a few % relative drift is expected
    ok = worst_l < 1e-3 and worst_g < 3e-2 and drift < 0.15 and len(set(digests)) == 1
    if rv.rank == 0:
        print(("MULTI_GPU_CHECK_PACKED " if packed else "MULTI_GPU_CHECK_RAGGED " if ragged else "MULTI_GPU_CHECK ") + json.dumps({"world": rv.world, "ok": ok, "groups_rank0": groups, "loss_rel": worst_l, "gnorm_rel": worst_g,
                                               "adapter_drift": drift, "replicas_bitwise_equal": len(set(digests)) == 1}), flush=True)
    rv.close()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    if "--full" in sys.argv:
        main_full()
    else:
        main(ragged="--ragged" in sys.argv, packed="--packed" in sys.argv)
