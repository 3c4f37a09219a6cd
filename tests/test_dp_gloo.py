"""world_size-2 CPU (gloo) coverage of the N>1 host path: id hand-off, barrier, max/mean over ranks, rank-sharded
synthetic batches, and the data-parallel semantics the native all-reduce must reproduce (oracle world=2 ==
mean of per-rank gradients == one rank seeing both micro-batches with grad_accum=2)."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp

from oracle import llama_lora as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(1)  # same reduction order as the parent: thread counts change fp32 summation orders
    from datatunerx_b200.dist import Rendezvous
    rv = Rendezvous()
    blob = rv.broadcast_bytes(lambda: bytes(range(128)))
    rv.barrier()
    mx = rv.max_over_ranks(float(rank + 1))
    mean = rv.mean_over_ranks(float(rank))
    ids, labels = O.synthetic_batch(step=3, rank=rank, batch=2, seq_len=128, vocab=512)
    # per-rank gradient of the oracle on this rank's shard, then gloo mean all-reduce == DeepSpeed ZeRO-0 semantics
    cfg = O.OracleConfig(vocab=512, hidden=256, n_layers=1, n_heads=2, ffn=256, lora_r=8)
    tr = O.OracleTrainer(cfg, O.init_base_weights(cfg, 1), O.init_lora(cfg, 2))
    loss, g = tr.loss_and_grads(ids, labels)
    flat = torch.cat([g[k].flatten() for k in sorted(g)])
    rv.dist.all_reduce(flat, op=rv.dist.ReduceOp.SUM)
    flat /= world
    q.put((rank, blob, mx, mean, int(ids.sum()), loss, flat.numpy()))
    rv.close()


def test_two_rank_gloo_rendezvous_and_dp_semantics():
    torch.set_num_threads(1)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = sorted([q.get(timeout=240) for _ in range(2)], key=lambda t: t[0])
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    (r0, b0, mx0, mean0, s0, l0, f0), (r1, b1, mx1, mean1, s1, l1, f1) = got
    assert b0 == b1 == bytes(range(128))           # 128-byte id reaches every rank intact
    assert mx0 == mx1 == 2.0 and mean0 == mean1 == 0.5
    assert s0 != s1                                # ranks draw different shards
    assert np.array_equal(f0, f1)                  # all-reduced gradient identical on both ranks
    # and equals what the single-process oracle computes for world=2
    cfg = O.OracleConfig(vocab=512, hidden=256, n_layers=1, n_heads=2, ffn=256, lora_r=8)
    tr = O.OracleTrainer(cfg, O.init_base_weights(cfg, 1), O.init_lora(cfg, 2), world=2)
    gs = [tr.loss_and_grads(*O.synthetic_batch(3, r, 2, 128, 512))[1] for r in range(2)]
    ref = torch.cat([((gs[0][k] + gs[1][k]) / 2).flatten() for k in sorted(gs[0])]).numpy()
    assert np.allclose(f0, ref, rtol=1e-4, atol=1e-7)  # fp32 sums in two processes: orders may still differ by a few ulp


def test_oracle_world2_equals_grad_accum2():
    cfg = O.OracleConfig(vocab=512, hidden=256, n_layers=1, n_heads=2, ffn=256, lora_r=8, total_steps=3, lr=1e-3)
    w, lora = O.init_base_weights(cfg, 1), O.init_lora(cfg, 2)
    batches = [O.synthetic_batch(0, r, 2, 128, 512) for r in range(2)]
    a = O.OracleTrainer(cfg, w, lora, world=2)
    la = a.step(batches)
    cfg2 = O.OracleConfig(**{**cfg.__dict__, "grad_accum": 2})
    b = O.OracleTrainer(cfg2, w, lora, world=1)
    lb = b.step(batches)
    assert la.loss == lb.loss and la.grad_norm == lb.grad_norm
    for k in a.lora:
        assert torch.equal(a.lora[k], b.lora[k])
