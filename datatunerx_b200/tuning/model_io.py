"""Model / adapter file IO of the worker, numpy only (no torch, no safetensors package needed at run time).

  * load_model_config  : HF config.json -> ModelConfig  (AutoConfig.from_pretrained, cmd/tuning/train.py:221)
  * iter_safetensors   : stream tensors of *.safetensors shards (AutoModelForCausalLM.from_pretrained, train.py:236-242);
                         iter_torch_bin does the same for pytorch_model*.bin shards
  * save_peft_adapter  : adapter_config.json + adapter_model.safetensors (+ adapter_model.bin when torch is importable)
                         in the layout peft 0.5.0 `save_pretrained` writes and the inference image loads
                         (train.py:300; pkg/util/generate/generate.go:287-294 CHECKPOINT_DIR)
"""
from __future__ import annotations

import glob
import json
import os
import struct
from typing import Dict, Iterator, Tuple

import numpy as np

from ..lib import ModelConfig

_ST_DTYPES = {"F32": (np.float32, 4, False), "F16": (np.float16, 2, False), "BF16": (np.uint16, 2, True)}


def load_model_config(model_dir: str) -> ModelConfig:
    cfg = json.load(open(os.path.join(model_dir, "config.json")))
    arch = (cfg.get("architectures") or ["LlamaForCausalLM"])[0]
    if "Llama" not in arch and "Mistral" not in arch:
        raise ValueError(f"{arch}: only Llama-family decoders are implemented natively")
    heads = cfg["num_attention_heads"]
    return ModelConfig(vocab=cfg["vocab_size"], hidden=cfg["hidden_size"], n_layers=cfg["num_hidden_layers"], n_heads=heads,
                       n_kv_heads=cfg.get("num_key_value_heads", heads), head_dim=cfg.get("head_dim") or cfg["hidden_size"] // heads,
                       ffn=cfg["intermediate_size"], rms_eps=cfg.get("rms_norm_eps", 1e-5), rope_theta=cfg.get("rope_theta", 10000.0),
                       max_seq=cfg.get("max_position_embeddings", 4096), sliding_window=int(cfg.get("sliding_window") or 0))


def check_attention_window(model_dir: str, seq_len: int) -> None:
    """Kept for callers of the r01 API: sliding-window attention (Mistral `sliding_window`) is implemented natively - the
    window travels in ModelConfig.sliding_window and only changes the mask once seq_len exceeds it."""
    return None


def iter_safetensors(path: str) -> Iterator[Tuple[str, np.ndarray, bool]]:
    """Yields (name, array, is_bf16_bits).  bf16 tensors come back as uint16 bit patterns."""
    with open(path, "rb") as f:
        (hlen,) = struct.unpack("<Q", f.read(8))
        header = json.loads(f.read(hlen))
        base = 8 + hlen
        mm = np.memmap(path, dtype=np.uint8, mode="r")
        for name, meta in header.items():
            if name == "__metadata__":
                continue
            if meta["dtype"] not in _ST_DTYPES:
                raise ValueError(f"{name}: dtype {meta['dtype']} not supported")
            dt, _, bits = _ST_DTYPES[meta["dtype"]]
            b, e = meta["data_offsets"]
            arr = np.frombuffer(mm[base + b: base + e], dtype=dt).reshape(meta["shape"])
            yield name, arr, bits


def iter_torch_bin(path: str) -> Iterator[Tuple[str, np.ndarray, bool]]:
    """pytorch_model*.bin shards (torch pickles, the pre-safetensors HF format AutoModelForCausalLM.from_pretrained also
    accepts, cmd/tuning/train.py:236-242).  Unpickling needs torch on the host - file plumbing only, no torch compute."""
    import torch
    sd = torch.load(path, map_location="cpu", weights_only=True, mmap=True)
    for name, t in sd.items():
        if t.dtype == torch.bfloat16:
            yield name, t.contiguous().view(torch.uint16).numpy(), True
        elif t.dtype in (torch.float16, torch.float32):
            yield name, t.contiguous().numpy(), False
        else:
            yield name, t.float().contiguous().numpy(), False


def load_weights_into(trainer, model_dir: str) -> int:
    """Stream every checkpoint tensor into the trainer.  The library refuses to step while a base tensor is missing, so a
    partial checkpoint fails loudly instead of training on uninitialised memory."""
    files = sorted(glob.glob(os.path.join(model_dir, "*.safetensors")))
    readers = [(f, iter_safetensors) for f in files]
    if not files:
        bins = sorted(glob.glob(os.path.join(model_dir, "pytorch_model*.bin")))
        if not bins:
            raise FileNotFoundError(f"no *.safetensors or pytorch_model*.bin under {model_dir}")
        readers = [(f, iter_torch_bin) for f in bins]
    cfg = json.load(open(os.path.join(model_dir, "config.json")))
    n, embed, have_head = 0, None, False
    for fpath, reader in readers:
        for name, arr, bits in reader(fpath):
            if "rotary_emb.inv_freq" in name:
                continue
            arr = np.ascontiguousarray(arr)
            trainer.load_tensor(name, arr, bf16_bits=bits)
            if name.endswith("embed_tokens.weight"):
                embed = (arr, bits)
            have_head = have_head or name.endswith("lm_head.weight")
            n += 1
    if not have_head and cfg.get("tie_word_embeddings") and embed is not None:
        trainer.load_tensor("lm_head.weight", embed[0], bf16_bits=embed[1])  # tied output embedding: no separate tensor on disk
        n += 1
    return n


def write_safetensors(path: str, tensors: Dict[str, np.ndarray]) -> None:
    header, off, blobs = {}, 0, []
    for k in sorted(tensors):
        a = np.ascontiguousarray(tensors[k], dtype=np.float32)
        header[k] = {"dtype": "F32", "shape": list(a.shape), "data_offsets": [off, off + a.nbytes]}
        blobs.append(a.tobytes())
        off += a.nbytes
    header["__metadata__"] = {"format": "pt"}
    h = json.dumps(header, separators=(",", ":")).encode()
    h += b" " * ((8 - len(h) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(h)))
        f.write(h)
        for b in blobs:
            f.write(b)


def save_full_model(out_dir: str, weights: Dict[str, np.ndarray], model_dir: str, shard_bytes: int = 5 << 30) -> None:
    """trainer.save_model of a full fine-tune: HF checkpoint directory - bf16 `model-0000x-of-0000y.safetensors` shards (uint16 bit
    patterns straight from the device), `model.safetensors.index.json`, and the base model's config / tokenizer files."""
    import shutil
    os.makedirs(out_dir, exist_ok=True)
    shards, cur, size = [], {}, 0
    for k, a in weights.items():
        if cur and size + a.nbytes > shard_bytes:
            shards.append(cur)
            cur, size = {}, 0
        cur[k] = a
        size += a.nbytes
    shards.append(cur)
    index = {"metadata": {"total_size": int(sum(a.nbytes for a in weights.values()))}, "weight_map": {}}
    for i, sh in enumerate(shards):
        name = "model.safetensors" if len(shards) == 1 else f"model-{i + 1:05d}-of-{len(shards):05d}.safetensors"
        header, off, blobs = {}, 0, []
        for k in sh:
            a = np.ascontiguousarray(sh[k], dtype=np.uint16)
            header[k] = {"dtype": "BF16", "shape": list(a.shape), "data_offsets": [off, off + a.nbytes]}
            blobs.append(a)
            off += a.nbytes
            index["weight_map"][k] = name
        header["__metadata__"] = {"format": "pt"}
        h = json.dumps(header, separators=(",", ":")).encode()
        h += b" " * ((8 - len(h) % 8) % 8)
        with open(os.path.join(out_dir, name), "wb") as f:
            f.write(struct.pack("<Q", len(h)))
            f.write(h)
            for b in blobs:
                f.write(b.tobytes())
    if len(shards) > 1:
        json.dump(index, open(os.path.join(out_dir, "model.safetensors.index.json"), "w"), indent=2)
    for fn in os.listdir(model_dir):
        if fn.endswith(".json") and not fn.endswith(".index.json") or fn.endswith(".model") or fn.startswith("tokenizer"):
            src = os.path.join(model_dir, fn)
            if os.path.isfile(src):
                shutil.copy(src, os.path.join(out_dir, fn))


def save_peft_adapter(out_dir: str, adapter: Dict[str, np.ndarray], *, base_model: str, r: int, alpha: float, dropout: float,
                      target_modules) -> None:
    os.makedirs(out_dir, exist_ok=True)
    cfg = {  # peft 0.5.0 LoraConfig.to_dict() as built at cmd/tuning/train.py:268-276
        "auto_mapping": None, "base_model_name_or_path": base_model, "bias": "none", "fan_in_fan_out": False,
        "inference_mode": True, "init_lora_weights": True, "layers_pattern": None, "layers_to_transform": None,
        "lora_alpha": alpha, "lora_dropout": dropout, "modules_to_save": None, "peft_type": "LORA", "r": r, "revision": None,
        "target_modules": list(target_modules), "task_type": "CAUSAL_LM",
    }
    json.dump(cfg, open(os.path.join(out_dir, "adapter_config.json"), "w"), indent=2, sort_keys=True)
    write_safetensors(os.path.join(out_dir, "adapter_model.safetensors"), adapter)
    try:  # peft 0.5.0 wrote adapter_model.bin (torch pickle); emit it too when torch is around (host plumbing only)
        import torch
        torch.save({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in adapter.items()}, os.path.join(out_dir, "adapter_model.bin"))
    except Exception:
        pass
