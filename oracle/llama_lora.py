"""CPU oracle for the data-parallel LoRA-SFT training step.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package.  The product (datatunerx_b200/) never does: it fails loudly without the CUDA library.

What it restates (fp32, plain torch tensor ops, no nn.Module magic):
  * the step the reference reaches through `trainer.train()` (cmd/tuning/train.py:299) with the arguments it
    actually builds (cmd/tuning/train.py:196-217): HF Trainer.training_step -> LlamaForCausalLM.forward ->
    shifted CE (ignore -100, token mean) -> backward -> clip_grad_norm_(1.0) -> torch.optim.AdamW ->
    linear/cosine schedule with 0 warm-up (the reference drops --warmup_ratio: train.py:204);
  * peft 0.5.0 lora.Linear on q_proj,v_proj (cmd/tuning/train.py:266-280; finetune_controller.go:482):
        y = W x + (alpha/r) * B(A(dropout(x))),  A ~ kaiming_uniform(a=sqrt 5),  B = 0,  adapters fp32;
  * fp32 logits (lm_head patch, cmd/tuning/train.py:256-264).

Where the arithmetic lives: NOT in the reference repo.  It is in un-vendored wheels pinned at
cmd/tuning/requirements.txt:1-9 (transformers 4.34.0, peft 0.5.0, torch 2.1.0, deepspeed 0.12.2).  The model
math below follows the installed transformers 5.5.0 copy (the only one on this box; forward math of Llama is
unchanged since 4.34): transformers/models/llama/modeling_llama.py:62-66 (RMSNorm in fp32), :138-168
(rotate_half RoPE), :182-184 (SwiGLU MLP), :212-217 (fp32 softmax attention), :303-334 (pre-norm block);
transformers/loss/loss_utils.py:28-67 (shifted CE).

Pinning status: the reference has NO tests, goldens or fixtures for this path — "parity unpinned"
against the reference's own tests.  What pins this oracle instead (tests/test_oracle_pin.py):
  * logits, loss and adapter gradients == installed HF LlamaForCausalLM + a peft-semantics LoRA wrap (1e-5);
  * AdamW / clip / schedule == torch.optim.AdamW, torch.nn.utils.clip_grad_norm_, transformers.get_scheduler;
  * the LR schedule known-answer hand-typed in cmd/tuning/prometheus/metrics.py:117-124
    (step 10 of 84, base lr 5e-5 -> 4.404761904761905e-05).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

IGNORE_INDEX = -100


@dataclass
class OracleConfig:
    vocab: int = 2048
    hidden: int = 256
    n_layers: int = 2
    n_heads: int = 2
    ffn: int = 768
    n_kv_heads: Optional[int] = None   # grouped-query attention when < n_heads (Mistral, Llama-2-70B)
    sliding_window: int = 0            # Mistral: query i attends keys i - sliding_window .. i (transformers 4.34.0 mask); 0 = none
    head_dim: int = 128
    rms_eps: float = 1e-5
    rope_theta: float = 10000.0
    # training (defaults = what the reference worker ends up with)
    lora_r: int = 16
    lora_alpha: float = 32.0
    lora_dropout: float = 0.0          # reference default 0.1 (cmd/tuning/parser.py:146-149); parity config uses 0
    seed: int = 42
    lora_target: Tuple[str, ...] = ("q_proj", "v_proj")
    lr: float = 1e-4
    weight_decay: float = 0.0
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-8
    max_grad_norm: float = 1.0
    sched: str = "linear"
    warmup_steps: int = 0
    total_steps: int = 100
    grad_accum: int = 1
    # full-parameter SFT (BASELINE.json configs[3]; beyond the reference, which always wraps LoRA - cmd/tuning/train.py:277):
    # every weight trains, no adapters; AdamW weight decay skips the RMSNorm weights like HF Trainer.get_decay_parameter_names
    full_finetune: bool = False

    @staticmethod
    def llama2_7b(**kw) -> "OracleConfig":
        return OracleConfig(vocab=32000, hidden=4096, n_layers=32, n_heads=32, ffn=11008, **kw)


NF4_LEVELS = np.array([-1.0, -0.6961928009986877, -0.5250730514526367, -0.39491748809814453, -0.28444138169288635,
                       -0.18477343022823334, -0.09105003625154495, 0.0, 0.07958029955625534, 0.16093020141124725,
                       0.24611230194568634, 0.33791524171829224, 0.44070982933044434, 0.5626170039176941, 0.7229568362236023,
                       1.0], dtype=np.float32)


def nf4_roundtrip(w: torch.Tensor, blocksize: int = 64) -> torch.Tensor:
    """dequantize(quantize(w)) of bitsandbytes 0.41 `quantize_4bit(quant_type="nf4", blocksize=64)` without double
    quantisation (cmd/tuning/train.py:224-230): per block of 64 consecutive elements absmax in fp32, x/absmax mapped to the
    nearest of the 16 NF4 levels (the library's decision tree uses the midpoints), value = level * absmax.  The levels are the
    published table of the QLoRA paper / bitsandbytes `create_normal_map`.  Result rounded to bf16 like the device copy."""
    x = w.detach().float().numpy().reshape(-1, blocksize)
    amax = np.abs(x).max(axis=1, keepdims=True).astype(np.float32)
    inv = np.where(amax > 0, np.float32(1.0) / np.where(amax > 0, amax, 1), 0).astype(np.float32)
    mids = (np.float32(0.5) * (NF4_LEVELS[:-1] + NF4_LEVELS[1:])).astype(np.float32)
    code = (((x * inv)[..., None] > mids).sum(-1)).astype(np.int64)
    out = NF4_LEVELS[code] * amax
    return bf16_round(torch.from_numpy(out.astype(np.float32)).view(w.shape))


QUANTIZED_SUFFIXES = ("q_proj.weight", "k_proj.weight", "v_proj.weight", "o_proj.weight", "gate_proj.weight", "up_proj.weight",
                      "down_proj.weight")


def quantize_base_nf4(weights: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """bitsandbytes load_in_4bit replaces every decoder nn.Linear (lm_head is skipped, embeddings/norms untouched)."""
    return {k: (nf4_roundtrip(v) if k.endswith(QUANTIZED_SUFFIXES) else v) for k, v in weights.items()}


def bf16_round(t: torch.Tensor) -> torch.Tensor:
    """Round to bf16 and back: both sides of a parity test start from bit-identical bf16 weights."""
    return t.to(torch.bfloat16).to(torch.float32)


def init_base_weights(cfg: OracleConfig, seed: int = 1234, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """HF _init_weights for Llama: Linear/Embedding ~ N(0, 0.02), RMSNorm weight = 1 (bf16-representable values)."""
    g = torch.Generator().manual_seed(seed)
    d, F, V = cfg.hidden, cfg.ffn, cfg.vocab
    dkv = (cfg.n_kv_heads or cfg.n_heads) * cfg.head_dim

    def n(*shape):
        return bf16_round(torch.randn(*shape, generator=g) * 0.02).to(dtype)

    w = {"model.embed_tokens.weight": n(V, d)}
    for l in range(cfg.n_layers):
        p = f"model.layers.{l}."
        w[p + "self_attn.q_proj.weight"] = n(d, d)
        w[p + "self_attn.k_proj.weight"] = n(dkv, d)
        w[p + "self_attn.v_proj.weight"] = n(dkv, d)
        w[p + "self_attn.o_proj.weight"] = n(d, d)
        w[p + "mlp.gate_proj.weight"] = n(F, d)
        w[p + "mlp.up_proj.weight"] = n(F, d)
        w[p + "mlp.down_proj.weight"] = n(d, F)
        w[p + "input_layernorm.weight"] = torch.ones(d, dtype=dtype)
        w[p + "post_attention_layernorm.weight"] = torch.ones(d, dtype=dtype)
    w["model.norm.weight"] = torch.ones(d, dtype=dtype)
    w["lm_head.weight"] = n(V, d)
    return w


def init_lora(cfg: OracleConfig, seed: int = 4321) -> Dict[str, torch.Tensor]:
    """peft 0.5.0 LoraLayer.reset_lora_parameters: kaiming_uniform_(A, a=sqrt(5)); zeros_(B)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    d, r = cfg.hidden, cfg.lora_r
    dkv = (cfg.n_kv_heads or cfg.n_heads) * cfg.head_dim
    bound = 1.0 / math.sqrt(d)  # gain sqrt(2/(1+5)) * sqrt(3/fan_in) = 1/sqrt(fan_in)
    for l in range(cfg.n_layers):
        for t in cfg.lora_target:
            p = f"model.layers.{l}.self_attn.{t}."
            out[p + "lora_A.weight"] = (torch.rand(r, d, generator=g) * 2 - 1) * bound
            out[p + "lora_B.weight"] = torch.zeros(d if t == "q_proj" else dkv, r)
    return out


from datatunerx_b200.tuning.synthetic import synthetic_batch  # noqa: E402,F401  (shared input generator; not oracle math)


# ------------------------------------------------------------------------------------------------
# model math
# ------------------------------------------------------------------------------------------------
def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    # modeling_llama.py:62-66
    var = x.pow(2).mean(-1, keepdim=True)
    return w * (x * torch.rsqrt(var + eps))


def rope_cos_sin(seq_len: int, head_dim: int, theta: float) -> Tuple[torch.Tensor, torch.Tensor]:
    # modeling_llama.py LlamaRotaryEmbedding: inv_freq = 1/theta^(2i/D); emb = cat(freqs, freqs)
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.float32) / head_dim))
    pos = torch.arange(seq_len, dtype=torch.float32)
    freqs = torch.outer(pos, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos(), emb.sin()


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    # modeling_llama.py:138-143
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def apply_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    # x: [B, H, S, D]; modeling_llama.py:146-168
    return x * cos[None, None] + rotate_half(x) * sin[None, None]


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, window: int = 0) -> torch.Tensor:
    # eager_attention_forward, modeling_llama.py:197-221: fp32 softmax(QK^T/sqrt(D) + causal) V.
    # window > 0: Mistral sliding window as transformers 4.34.0 (the version the reference pins, requirements.txt:9) builds
    # it - modeling_mistral.py _make_sliding_window_causal_mask: tril(diagonal=0) & triu(diagonal=-sliding_window), i.e. query
    # i sees keys i - window .. i (window + 1 keys).  Newer transformers (5.x, installed here) use i - window < j: one key
    # fewer; tests/test_oracle_pin.py pins this function against the installed version with window - 1.
    S, D = q.shape[-2], q.shape[-1]
    scores = q @ k.transpose(-1, -2) / math.sqrt(D)
    mask = torch.full((S, S), float("-inf")).triu(1)
    if window > 0:
        mask = mask + torch.full((S, S), float("-inf")).tril(-(window + 1))
    p = torch.softmax(scores + mask, dim=-1, dtype=torch.float32)
    return p @ v


_M64 = (1 << 64) - 1


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def dropout_key(seed: int, fwd_count: int, layer: int, rank: int = 0) -> int:
    """Key of the LoRA-dropout masks of one (forward pass, layer) — the same integer arithmetic as dropout_key() in
    datatunerx_b200/csrc/trainer.cu (torch's own Philox stream cannot be reproduced on the device; what the reference fixes
    is the distribution: independent Bernoulli(1-p) keep masks per wrapped module, kept values scaled by 1/(1-p))."""
    x = (seed * 0xD1B54A32D192ED03 + fwd_count * 0x100000001B3 + layer * 0x9E3779B1 + rank * 0xC2B2AE3D27D4EB4F) & _M64
    with np.errstate(over="ignore"):
        return int(_splitmix64(np.array([x], dtype=np.uint64))[0])


def dropout_mask(key: int, target: int, n_rows: int, d: int, p: float) -> torch.Tensor:
    """keep[m, c] = splitmix64(key + target*G + m*d + c) >> 40 >= p * 2^24   (elementwise.cu drop_keep)."""
    with np.errstate(over="ignore"):
        idx = np.arange(n_rows * d, dtype=np.uint64) + np.uint64((key + target * 0x9E3779B97F4A7C15) & _M64)
        u = (_splitmix64(idx) >> np.uint64(40)).astype(np.int64)
    thresh = min(int(p * 16777216.0), 16777215)
    return torch.from_numpy((u >= thresh).astype(np.float32)).view(n_rows, d)


def lora_linear(x: torch.Tensor, w: torch.Tensor, a: Optional[torch.Tensor], b: Optional[torch.Tensor], scale: float,
                mask: Optional[torch.Tensor] = None, p: float = 0.0):
    # peft 0.5.0 lora.Linear.forward: result = F.linear(x, W) + lora_B(lora_A(lora_dropout(x))) * scaling
    y = x @ w.t()
    if a is not None:
        xd = x if mask is None else x * mask.view(x.shape) / (1.0 - p)
        y = y + (xd @ a.t()) @ b.t() * scale
    return y


def forward_logits(cfg: OracleConfig, w: Dict[str, torch.Tensor], lora: Dict[str, torch.Tensor], ids: torch.Tensor,
                   drop_ctx: Optional[Tuple[int, int]] = None):
    """drop_ctx = (fwd_count, rank) enables LoRA dropout with the counter-based masks; None = eval mode / p = 0."""
    B, S = ids.shape
    H, D = cfg.n_heads, cfg.head_dim
    Hkv = cfg.n_kv_heads or H
    targets = [t for t in ("q_proj", "k_proj", "v_proj") if t in cfg.lora_target]
    scale = cfg.lora_alpha / cfg.lora_r
    cos, sin = rope_cos_sin(S, D, cfg.rope_theta)
    x = w["model.embed_tokens.weight"][ids.long()]
    for l in range(cfg.n_layers):
        p = f"model.layers.{l}."
        h = rmsnorm(x, w[p + "input_layernorm.weight"], cfg.rms_eps)

        def proj(name):
            a = lora.get(p + f"self_attn.{name}.lora_A.weight")
            b = lora.get(p + f"self_attn.{name}.lora_B.weight")
            mask = None
            if a is not None and drop_ctx is not None and cfg.lora_dropout > 0:
                key = dropout_key(cfg.seed, drop_ctx[0], l, drop_ctx[1])
                mask = dropout_mask(key, targets.index(name), B * S, cfg.hidden, cfg.lora_dropout)
            return lora_linear(h, w[p + f"self_attn.{name}.weight"], a, b, scale, mask, cfg.lora_dropout)

        q = proj("q_proj").view(B, S, H, D).transpose(1, 2)
        k = proj("k_proj").view(B, S, Hkv, D).transpose(1, 2)
        v = proj("v_proj").view(B, S, Hkv, D).transpose(1, 2)
        q, k = apply_rope(q, cos, sin), apply_rope(k, cos, sin)
        if Hkv != H:  # HF repeat_kv: kv head i serves query heads i*g .. (i+1)*g-1
            k = k.repeat_interleave(H // Hkv, dim=1)
            v = v.repeat_interleave(H // Hkv, dim=1)
        o = attention(q, k, v, cfg.sliding_window if S > cfg.sliding_window + 1 else 0).transpose(1, 2).reshape(B, S, H * D)
        x = x + o @ w[p + "self_attn.o_proj.weight"].t()
        h = rmsnorm(x, w[p + "post_attention_layernorm.weight"], cfg.rms_eps)
        g = h @ w[p + "mlp.gate_proj.weight"].t()
        u = h @ w[p + "mlp.up_proj.weight"].t()
        x = x + (torch.nn.functional.silu(g) * u) @ w[p + "mlp.down_proj.weight"].t()
    x = rmsnorm(x, w["model.norm.weight"], cfg.rms_eps)
    return (x @ w["lm_head.weight"].t()).float()


def causal_lm_loss(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    # loss_utils.py ForCausalLMLoss: shift, ignore_index=-100, mean over valid tokens
    V = logits.shape[-1]
    shift_logits = logits[:, :-1, :].reshape(-1, V)
    shift_labels = labels[:, 1:].reshape(-1).long()
    valid = shift_labels != IGNORE_INDEX
    lse = torch.logsumexp(shift_logits, dim=-1)
    tgt = shift_logits.gather(1, shift_labels.clamp(min=0)[:, None])[:, 0]
    return ((lse - tgt) * valid).sum() / valid.sum().clamp(min=1)


# ------------------------------------------------------------------------------------------------
# optimiser / schedule (restated; pinned against torch / transformers in tests/test_oracle_pin.py)
# ------------------------------------------------------------------------------------------------
def lr_lambda(sched: str, step: int, warmup: int, total: int) -> float:
    """transformers.optimization get_{linear,cosine,constant}_schedule_with_warmup."""
    if sched == "constant":
        return 1.0
    if step < warmup:
        return step / max(1, warmup)
    if sched == "constant_with_warmup":
        return 1.0
    if sched == "linear":
        return max(0.0, (total - step) / max(1, total - warmup))
    if sched == "cosine":
        prog = (step - warmup) / max(1, total - warmup)
        return max(0.0, 0.5 * (1.0 + math.cos(math.pi * 0.5 * 2.0 * prog)))
    raise ValueError(sched)


def clip_coef(total_norm: float, max_norm: float) -> float:
    """torch.nn.utils.clip_grad_norm_: coef = max_norm / (norm + 1e-6), clamped to 1."""
    return min(1.0, max_norm / (total_norm + 1e-6))


def adamw_update(p, g, m, v, step: int, lr: float, b1: float, b2: float, eps: float, wd: float) -> None:
    """torch.optim.AdamW single-tensor path (no amsgrad, no maximize), in place, step is 1-based."""
    p.mul_(1 - lr * wd)
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))


@dataclass
class StepLog:
    loss: float
    grad_norm: float
    lr: float


class OracleTrainer:
    """fp32 CPU restatement of the reference worker's training loop for one data-parallel group.
    `self.lora` is the dict of TRAINABLE tensors: the adapters, or - with cfg.full_finetune - every model weight (then `lora`
    passed in is ignored and no adapter exists)."""

    def __init__(self, cfg: OracleConfig, weights: Dict[str, torch.Tensor], lora: Dict[str, torch.Tensor], world: int = 1):
        self.cfg, self.w, self.world = cfg, weights, world
        if cfg.full_finetune:
            self.lora = {k: v.clone().float().requires_grad_(True) for k, v in weights.items()}
            self.w = self.lora  # forward_logits reads the live (trainable) weights; it finds no adapter keys in them
        else:
            self.lora = {k: v.clone().float().requires_grad_(True) for k, v in lora.items()}
        self.m = {k: torch.zeros_like(v) for k, v in self.lora.items()}
        self.v = {k: torch.zeros_like(v) for k, v in self.lora.items()}
        self.opt_step = 0
        self.micro = 0
        self.fwd_count = 0  # mirrors dtx_trainer::fwd_count (every forward pass, training or eval, advances it)
        self.rank_of_next = 0
        self._acc: Dict[str, torch.Tensor] = {}

    def loss_and_grads(self, ids: np.ndarray, labels: np.ndarray, rank: int = 0,
                       fwd_count: Optional[int] = None) -> Tuple[float, Dict[str, torch.Tensor]]:
        for p in self.lora.values():
            p.grad = None
        if fwd_count is None:
            self.fwd_count += 1
            fwd_count = self.fwd_count
        logits = forward_logits(self.cfg, self.w, self.lora, torch.from_numpy(np.asarray(ids)).long(), (fwd_count, rank))
        loss = causal_lm_loss(logits, torch.from_numpy(np.asarray(labels)).long())
        loss.backward()
        return float(loss.detach()), {k: p.grad.detach().clone() for k, p in self.lora.items()}

    def eval_loss(self, ids: np.ndarray, labels: np.ndarray) -> float:
        self.fwd_count += 1
        with torch.no_grad():
            logits = forward_logits(self.cfg, self.w, self.lora, torch.from_numpy(np.asarray(ids)).long())
            return float(causal_lm_loss(logits, torch.from_numpy(np.asarray(labels)).long()))

    def step(self, batches: Sequence[Tuple[np.ndarray, np.ndarray]]) -> StepLog:
        """One optimizer step.  `batches` holds world*grad_accum micro-batches (rank-major); gradients are
        averaged over all of them (DeepSpeed ZeRO-0 mean all-reduce, ds_config.json) and the logged loss is
        the mean of the per-micro-batch token-mean losses (HF Trainer behaviour)."""
        cfg = self.cfg
        assert len(batches) == self.world * cfg.grad_accum
        acc = {k: torch.zeros_like(v) for k, v in self.lora.items()}
        losses = []
        base = self.fwd_count
        for i, (ids, labels) in enumerate(batches):  # rank-major: rank i // grad_accum, its (i % grad_accum)-th forward pass
            loss, g = self.loss_and_grads(ids, labels, rank=i // cfg.grad_accum, fwd_count=base + (i % cfg.grad_accum) + 1)
            losses.append(loss)
            for k in acc:
                acc[k] += g[k]
        self.fwd_count = base + cfg.grad_accum
        for k in acc:
            acc[k] /= len(batches)
        total_norm = math.sqrt(sum(float((g.double() ** 2).sum()) for g in acc.values()))
        coef = clip_coef(total_norm, cfg.max_grad_norm) if cfg.max_grad_norm > 0 else 1.0
        lr = cfg.lr * lr_lambda(cfg.sched, self.opt_step, cfg.warmup_steps, cfg.total_steps)
        self.opt_step += 1
        with torch.no_grad():
            for k, p in self.lora.items():
                wd = 0.0 if (cfg.full_finetune and k.endswith("norm.weight") or k.endswith("layernorm.weight")) else cfg.weight_decay
                adamw_update(p, acc[k] * coef, self.m[k], self.v[k], self.opt_step, lr, cfg.beta1, cfg.beta2, cfg.eps, wd)
        return StepLog(loss=float(np.mean(losses)), grad_norm=total_norm, lr=lr)

    def state_dict(self) -> Dict[str, np.ndarray]:
        return {k: v.detach().numpy().copy() for k, v in self.lora.items()}
