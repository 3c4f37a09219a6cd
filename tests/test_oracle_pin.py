"""Pins the CPU oracle (oracle/llama_lora.py) against the third-party implementations the reference calls
(transformers LlamaForCausalLM, torch.optim.AdamW, clip_grad_norm_, get_scheduler) and against the one numeric
artefact the reference repo holds for this path (cmd/tuning/prometheus/metrics.py:117-124)."""
import math

import numpy as np
import pytest
import torch

from oracle import llama_lora as O


def _hf_model(cfg: O.OracleConfig, weights):
    from transformers import LlamaConfig, LlamaForCausalLM
    hc = LlamaConfig(vocab_size=cfg.vocab, hidden_size=cfg.hidden, intermediate_size=cfg.ffn,
                     num_hidden_layers=cfg.n_layers, num_attention_heads=cfg.n_heads,
                     num_key_value_heads=cfg.n_kv_heads or cfg.n_heads,
                     max_position_embeddings=4096, rms_norm_eps=cfg.rms_eps, rope_theta=cfg.rope_theta,
                     tie_word_embeddings=False, attn_implementation="eager")
    m = LlamaForCausalLM(hc).float()
    missing, unexpected = m.load_state_dict(weights, strict=False)
    assert not [k for k in missing if "rotary" not in k], missing
    assert not unexpected
    return m.eval()


class _PeftLoRA(torch.nn.Module):
    """peft 0.5.0 lora.Linear forward semantics (dropout 0) around a frozen nn.Linear."""

    def __init__(self, base, a, b, scale):
        super().__init__()
        self.base, self.scale = base, scale
        self.lora_A = torch.nn.Parameter(a.clone())
        self.lora_B = torch.nn.Parameter(b.clone())

    def forward(self, x):
        return self.base(x) + (x @ self.lora_A.t()) @ self.lora_B.t() * self.scale


def test_oracle_matches_hf_llama_logits_loss_and_lora_grads():
    cfg = O.OracleConfig(vocab=512, hidden=256, n_layers=2, n_heads=2, ffn=384, lora_r=8, lora_alpha=16.0)
    w = O.init_base_weights(cfg, seed=7)
    lora = O.init_lora(cfg, seed=8)
    g = torch.Generator().manual_seed(9)
    for k in lora:  # non-zero B so the LoRA branch and its A-gradient are exercised
        if "lora_B" in k:
            lora[k] = torch.randn(lora[k].shape, generator=g) * 0.02
    ids, labels = O.synthetic_batch(step=0, rank=0, batch=2, seq_len=64, vocab=cfg.vocab)

    m = _hf_model(cfg, w)
    for p in m.parameters():
        p.requires_grad_(False)
    wrapped = {}
    for l, layer in enumerate(m.model.layers):
        for t in cfg.lora_target:
            key = f"model.layers.{l}.self_attn.{t}."
            mod = _PeftLoRA(getattr(layer.self_attn, t), lora[key + "lora_A.weight"], lora[key + "lora_B.weight"],
                            cfg.lora_alpha / cfg.lora_r)
            setattr(layer.self_attn, t, mod)
            wrapped[key] = mod
    t_ids, t_lab = torch.from_numpy(ids).long(), torch.from_numpy(labels).long()
    out = m(input_ids=t_ids, labels=t_lab)
    out.loss.backward()

    tr = O.OracleTrainer(cfg, w, lora)
    with torch.no_grad():
        logits = O.forward_logits(cfg, w, tr.lora, t_ids)
    assert torch.allclose(logits, out.logits.float(), atol=2e-5, rtol=1e-4)
    loss, grads = tr.loss_and_grads(ids, labels)
    assert abs(loss - float(out.loss)) < 1e-5
    for key, mod in wrapped.items():
        assert torch.allclose(grads[key + "lora_A.weight"], mod.lora_A.grad, atol=1e-6, rtol=1e-4)
        assert torch.allclose(grads[key + "lora_B.weight"], mod.lora_B.grad, atol=1e-6, rtol=1e-4)


def test_oracle_gqa_matches_hf_llama():
    """grouped-query attention (n_kv_heads < n_heads: Mistral-7B, Llama-2-70B) against the installed HF model."""
    cfg = O.OracleConfig(vocab=512, hidden=512, n_layers=1, n_heads=4, n_kv_heads=2, ffn=384, lora_r=8, lora_alpha=16.0)
    w = O.init_base_weights(cfg, seed=17)
    ids, labels = O.synthetic_batch(step=1, rank=0, batch=2, seq_len=48, vocab=cfg.vocab)
    m = _hf_model(cfg, w)
    t_ids, t_lab = torch.from_numpy(ids).long(), torch.from_numpy(labels).long()
    with torch.no_grad():
        out = m(input_ids=t_ids, labels=t_lab)
        logits = O.forward_logits(cfg, w, {}, t_ids)
    assert torch.allclose(logits, out.logits.float(), atol=2e-5, rtol=1e-4)
    assert abs(float(O.causal_lm_loss(logits, t_lab)) - float(out.loss)) < 1e-5


def test_dropout_masks_are_bernoulli_and_independent_per_module():
    key = O.dropout_key(42, 3, 1, 0)
    m0 = O.dropout_mask(key, 0, 512, 256, 0.1)
    m1 = O.dropout_mask(key, 1, 512, 256, 0.1)
    assert abs(float(m0.mean()) - 0.9) < 5e-3 and abs(float(m1.mean()) - 0.9) < 5e-3
    assert abs(float((m0 * m1).mean()) - 0.81) < 5e-3                      # independent masks for q_proj and v_proj
    assert not torch.equal(m0, O.dropout_mask(O.dropout_key(42, 4, 1, 0), 0, 512, 256, 0.1))  # fresh mask every forward pass
    assert torch.equal(O.dropout_mask(key, 0, 512, 256, 0.0), torch.ones(512, 256))


def test_adamw_clip_match_torch():
    torch.manual_seed(0)
    p0 = torch.randn(1000)
    p_ref = p0.clone().requires_grad_(True)
    opt = torch.optim.AdamW([p_ref], lr=3e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    p, m, v = p0.clone(), torch.zeros(1000), torch.zeros(1000)
    for step in range(1, 6):
        g = torch.randn(1000) * (5.0 if step == 2 else 0.01)
        p_ref.grad = g.clone()
        norm = torch.nn.utils.clip_grad_norm_([p_ref], 1.0)
        opt.step()
        coef = O.clip_coef(float(g.norm()), 1.0)
        assert abs(float(norm) - float(g.norm())) < 1e-5
        O.adamw_update(p, g * coef, m, v, step, 3e-4, 0.9, 0.999, 1e-8, 0.01)
        assert torch.allclose(p, p_ref.detach(), atol=1e-7, rtol=1e-6)


@pytest.mark.parametrize("sched", ["linear", "cosine", "constant", "constant_with_warmup"])
def test_schedule_matches_transformers(sched):
    from transformers.optimization import get_scheduler
    p = torch.nn.Parameter(torch.zeros(1))
    for warm in (0, 7):
        opt = torch.optim.SGD([p], lr=1.0)
        s = get_scheduler(sched, opt, num_warmup_steps=warm, num_training_steps=50)
        for step in range(50):
            assert abs(s.get_last_lr()[0] - O.lr_lambda(sched, step, warm, 50)) < 1e-12, (sched, warm, step)
            opt.step()
            s.step()


def test_lr_known_answer_from_reference_metrics_payload():
    # cmd/tuning/prometheus/metrics.py:117-124: total_steps 84, current_steps 10, learning_rate 4.404761904761905e-05
    # HF logs get_last_lr() after scheduler.step(): lambda(10) * 5e-5 with linear decay and zero warm-up
    assert 5e-5 * O.lr_lambda("linear", 10, 0, 84) == pytest.approx(4.404761904761905e-05, rel=1e-12)


def test_native_lr_lambda_matches_oracle():
    from datatunerx_b200 import lib as L
    for sched in ("linear", "cosine", "constant", "constant_with_warmup"):
        for warm in (0, 5):
            for step in range(0, 40):
                assert L.lr_lambda(sched, step, warm, 37) == pytest.approx(O.lr_lambda(sched, step, warm, 37), abs=1e-12)
    assert 5e-5 * L.lr_lambda("linear", 10, 0, 84) == pytest.approx(4.404761904761905e-05, rel=1e-12)


def test_oracle_loss_decreases_and_is_deterministic():
    cfg = O.OracleConfig(vocab=256, hidden=256, n_layers=1, n_heads=2, ffn=256, total_steps=5, lr=1e-2)
    w, lora = O.init_base_weights(cfg, 1), O.init_lora(cfg, 2)
    runs = []
    for _ in range(2):
        tr = O.OracleTrainer(cfg, w, lora)
        batch = O.synthetic_batch(0, 0, 2, 128, cfg.vocab)
        runs.append([tr.step([batch]).loss for _ in range(5)])
    assert runs[0] == runs[1]
    assert runs[0][-1] < runs[0][0]
    assert abs(runs[0][0] - math.log(cfg.vocab)) < 0.2


def test_oracle_sliding_window_matches_hf_mistral():
    """Mistral sliding-window attention against the installed HF MistralForCausalLM.  The installed transformers (5.x) lets query
    i see keys i - sw + 1 .. i; 4.34.0 - what the reference pins, cmd/tuning/requirements.txt:9 - builds
    triu(diagonal=-sliding_window), one key more (i - sw .. i).  The oracle takes the 4.34.0 width as its parameter, so the
    installed model with sliding_window = sw pins oracle.attention(window = sw - 1)."""
    from transformers import MistralConfig, MistralForCausalLM
    sw = 24
    cfg = O.OracleConfig(vocab=512, hidden=512, n_layers=2, n_heads=4, n_kv_heads=2, ffn=384, lora_r=8, sliding_window=sw - 1)
    w = O.init_base_weights(cfg, seed=27)
    hc = MistralConfig(vocab_size=cfg.vocab, hidden_size=cfg.hidden, intermediate_size=cfg.ffn, num_hidden_layers=cfg.n_layers,
                       num_attention_heads=cfg.n_heads, num_key_value_heads=cfg.n_kv_heads, head_dim=cfg.head_dim,
                       max_position_embeddings=4096, rms_norm_eps=cfg.rms_eps, rope_theta=cfg.rope_theta, sliding_window=sw,
                       tie_word_embeddings=False, attn_implementation="eager")
    m = MistralForCausalLM(hc).float()
    missing, unexpected = m.load_state_dict(w, strict=False)
    assert not [k for k in missing if "rotary" not in k] and not unexpected
    ids, _ = O.synthetic_batch(step=2, rank=0, batch=2, seq_len=96, vocab=cfg.vocab)
    t_ids = torch.from_numpy(ids).long()
    with torch.no_grad():
        ref = m.eval()(input_ids=t_ids).logits.float()
        got = O.forward_logits(cfg, w, {}, t_ids)
        plain = O.forward_logits(O.OracleConfig(**{**cfg.__dict__, "sliding_window": 0}), w, {}, t_ids)
    assert torch.allclose(got, ref, atol=3e-5, rtol=1e-4), float((got - ref).abs().max())
    assert float((plain - ref).abs().max()) > 1e-3, "the window must matter at this length"
    # the 4.34.0 width sees exactly one key more than the installed version
    S = 8
    q = torch.randn(1, 1, S, 16)
    a = O.attention(q, q, torch.eye(S).view(1, 1, S, S), window=3)  # V = identity: the output row IS the probability row
    assert [(a[0, 0, i] > 0).sum().item() for i in range(S)] == [min(i + 1, 4) for i in range(S)]
