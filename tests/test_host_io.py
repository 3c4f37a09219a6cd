"""Host-side side channels and file formats (no GPU): Prometheus remote-write encoding, LogCallback JSONL rows,
safetensors / PEFT adapter writer, optimizer-step arithmetic."""
import json
import os
import struct

import numpy as np

from datatunerx_b200.tuning import metrics as M
from datatunerx_b200.tuning import model_io
from datatunerx_b200.tuning.callback import LogCallback
from datatunerx_b200.tuning.train import total_optimizer_steps


def _decode_pb(buf):
    """tiny protobuf walker: returns list of (field, wiretype, value)"""
    out, i = [], 0
    while i < len(buf):
        key, shift = 0, 0
        while True:
            b = buf[i]; i += 1
            key |= (b & 0x7F) << shift; shift += 7
            if not b & 0x80:
                break
        f, wt = key >> 3, key & 7
        if wt == 2:
            n, shift = 0, 0
            while True:
                b = buf[i]; i += 1
                n |= (b & 0x7F) << shift; shift += 7
                if not b & 0x80:
                    break
            out.append((f, wt, buf[i:i + n])); i += n
        elif wt == 1:
            out.append((f, wt, struct.unpack("<d", buf[i:i + 8])[0])); i += 8
        else:
            n, shift = 0, 0
            while True:
                b = buf[i]; i += 1
                n |= (b & 0x7F) << shift; shift += 7
                if not b & 0x80:
                    break
            out.append((f, wt, n))
    return out


def test_remote_write_payload_matches_reference_series_layout():
    # the reference's own smoke payload: cmd/tuning/prometheus/metrics.py:117-124
    m = {"uid": "1", "total_steps": 84, "current_steps": 10, "loss": 3.088, "learning_rate": 4.404761904761905e-05, "epoch": 0.71}
    wr = M.encode_write_request([M.train_series(m, 1700000000000)])
    body = M.snappy_block_literal(wr)
    assert M.snappy_block_decode(body) == wr
    (f, wt, ts), = _decode_pb(wr)
    assert f == 1 and wt == 2
    fields = _decode_pb(ts)
    labels = [tuple(v.decode() for _, _, v in _decode_pb(x)) for fnum, _, x in fields if fnum == 1]
    assert labels == [("__name__", "train_metrics"), ("uid", "1"), ("total_steps", "84"), ("current_steps", "10"), ("loss", "3.088"),
                      ("learning_rate", "4.404761904761905e-05"), ("epoch", "0.71")]
    sample = [_decode_pb(x) for fnum, _, x in fields if fnum == 2][0]
    assert sample[0] == (1, 1, 1.0) and sample[1] == (2, 0, 1700000000000)
    ev = _decode_pb(_decode_pb(M.encode_write_request([M.eval_series({"uid": "u", "eval_loss": 1.5, "eval_perplexity": 4.48}, 5)]))[0][2])
    names = [_decode_pb(x)[0][2].decode() for fnum, _, x in ev if fnum == 1]
    assert names == ["__name__", "uid", "total_steps", "current_steps", "eval_loss", "eval_perplexity", "epoch"]


def test_snappy_literal_stream_handles_large_payloads():
    data = bytes(range(256)) * 1000
    assert M.snappy_block_decode(M.snappy_block_literal(data)) == data
    assert M.snappy_block_decode(M.snappy_block_literal(b"")) == b""


def test_log_callback_rows_have_reference_keys(tmp_path):
    cb = LogCallback(str(tmp_path), max_steps=84, metrics_export_address=None, uid="abc")
    cb.on_step_end(10)
    row = cb.on_log(3.08812345, 4.404761904761905e-05, 0.7142)
    cb.on_eval(1.25, float(np.exp(1.25)), 1.0)
    got = json.loads(open(tmp_path / "watch" / "trainer_log.jsonl").read().splitlines()[0])
    # key set and order of callback.py:120-138
    assert list(got) == ["uid", "current_steps", "total_steps", "loss", "eval_loss", "val_perplexity", "eval_rouge_1", "eval_rouge_2",
                         "eval_rouge_l", "eval_bleu_4", "predict_loss", "reward", "learning_rate", "epoch", "percentage", "elapsed_time",
                         "remaining_time"]
    assert got["loss"] == 3.0881 and got["epoch"] == 0.71 and got["percentage"] == 11.9 and got["uid"] == "abc" and row == got
    ev = json.loads(open(tmp_path / "watch" / "eval_log.jsonl").read())
    assert ev["eval_loss"] == 1.25 and abs(ev["eval_perplexity"] - np.exp(1.25)) < 1e-12


def test_peft_adapter_writer_roundtrip(tmp_path):
    ad = {"base_model.model.model.layers.0.self_attn.q_proj.lora_A.weight": np.arange(32, dtype=np.float32).reshape(2, 16),
          "base_model.model.model.layers.0.self_attn.q_proj.lora_B.weight": np.ones((16, 2), dtype=np.float32)}
    model_io.save_peft_adapter(str(tmp_path), ad, base_model="/tmp/llama2-7b/", r=2, alpha=32.0, dropout=0.0,
                               target_modules=["q_proj", "v_proj"])
    cfg = json.load(open(tmp_path / "adapter_config.json"))
    assert cfg["peft_type"] == "LORA" and cfg["task_type"] == "CAUSAL_LM" and cfg["r"] == 2 and cfg["target_modules"] == ["q_proj", "v_proj"]
    back = {k: (a, b) for k, a, b in model_io.iter_safetensors(str(tmp_path / "adapter_model.safetensors"))}
    for k, v in ad.items():
        assert np.array_equal(back[k][0], v) and back[k][1] is False
    from safetensors.numpy import load_file  # the real reader accepts our writer's output
    real = load_file(str(tmp_path / "adapter_model.safetensors"))
    assert all(np.array_equal(real[k], v) for k, v in ad.items())
    import torch
    sd = torch.load(tmp_path / "adapter_model.bin")
    assert all(np.array_equal(sd[k].numpy(), v) for k, v in ad.items())


def test_total_optimizer_steps_follow_hf_trainer_arithmetic():
    assert total_optimizer_steps(n_examples=672, world=1, batch=8, grad_accum=1, epochs=1, max_steps=-1) == 84
    assert total_optimizer_steps(n_examples=672, world=2, batch=8, grad_accum=2, epochs=3, max_steps=-1) == 63
    assert total_optimizer_steps(n_examples=5, world=1, batch=8, grad_accum=4, epochs=2, max_steps=-1) == 2
    assert total_optimizer_steps(n_examples=672, world=1, batch=8, grad_accum=1, epochs=1, max_steps=7) == 7


def test_eval_mean_follows_hf_batching():
    """SFTTrainer.evaluate (cmd/tuning/trainer.py:324-327 on HF Trainer.evaluation_loop): each eval batch's token-mean loss is
    repeated once per sample, gathered and averaged.  The native step returns per-row (loss sum, valid tokens); the host forms
    the per_device_eval_batch_size batches from them."""
    import numpy as np
    from datatunerx_b200.tuning import train as T

    rows = [(3.0, 3), (8.0, 4), (1.0, 1), (0.0, 0), (10.0, 5)]  # (sum of token losses, valid tokens) per example

    class FakeTrainer:
        def __init__(self):
            self.i = 0

        def eval_rows(self, ids, lab, lens):
            k = int((lens > 0).sum())
            part = rows[self.i:self.i + k]
            self.i += k
            pad = ids.shape[0] - k
            return (np.array([r[0] for r in part] + [0.0] * pad, dtype=np.float32), np.array([r[1] for r in part] + [0] * pad, dtype=np.int32))

        def allreduce_host(self, v):
            return np.asarray(v, dtype=np.float64)

    ds = [([1] * 5, [-100, 1, 1, 1, 1])] * len(rows)
    got = T.evaluate(FakeTrainer(), ds, rank=0, world=1, B=4, seq_len=128, pad_id=0, eval_batch=2, seed=0)
    # HF with eval batch 2: batches (r0,r1) (r2,r3) (r4): token-mean 11/7, 1/1, 10/5, weighted by batch size 2, 2, 1
    want = (2 * (11 / 7) + 2 * 1.0 + 1 * 2.0) / 5
    assert abs(got - want) < 1e-12, (got, want)


def test_child_rank_failure_takes_the_job_down(tmp_path):
    """A rank that dies must not leave its peers blocked in NCCL: the watchdog thread exits the whole job non-zero."""
    import subprocess
    import sys
    code = (
        "import subprocess, sys, threading, time\n"
        "from datatunerx_b200.tuning import train as T\n"
        "kids = [subprocess.Popen([sys.executable, '-c', 'import time; time.sleep(60)']),\n"
        "        subprocess.Popen([sys.executable, '-c', 'import sys, time; time.sleep(1); sys.exit(7)'])]\n"
        "threading.Thread(target=T._watch_children, args=(kids, threading.Event()), daemon=True).start()\n"
        "time.sleep(45)\n"   # stands for rank 0 blocked inside ncclAllReduce
        "sys.exit(0)\n")
    import os
    import time
    t0 = time.time()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=120)
    assert p.returncode == 7, (p.returncode, p.stderr[-500:])
    assert time.time() - t0 < 30, "the watchdog must fire within seconds"
    assert "aborting the job" in p.stderr


def test_full_model_checkpoint_round_trip(tmp_path):
    """Full fine-tune save: sharded bf16 safetensors + index + the base model's config / tokenizer files, readable by the loader."""
    import json
    import numpy as np
    from datatunerx_b200.tuning import model_io
    base = tmp_path / "base"
    base.mkdir()
    (base / "config.json").write_text(json.dumps({"model_type": "llama"}))
    (base / "tokenizer.json").write_text("{}")
    w = {"model.embed_tokens.weight": np.arange(32, dtype=np.uint16).reshape(4, 8), "lm_head.weight": np.arange(32, 64, dtype=np.uint16).reshape(4, 8),
         "model.norm.weight": np.arange(8, dtype=np.uint16)}
    out = tmp_path / "ckpt"
    model_io.save_full_model(str(out), w, str(base), shard_bytes=70)
    names = sorted(p.name for p in out.iterdir())
    assert "config.json" in names and "tokenizer.json" in names and "model.safetensors.index.json" in names
    idx = json.loads((out / "model.safetensors.index.json").read_text())
    assert set(idx["weight_map"]) == set(w) and idx["metadata"]["total_size"] == sum(a.nbytes for a in w.values())
    got = {}
    for f in names:
        if f.endswith(".safetensors"):
            for n, a, bits in model_io.iter_safetensors(str(out / f)):
                assert bits
                got[n] = np.array(a)
    assert all(np.array_equal(got[k], w[k]) for k in w)
