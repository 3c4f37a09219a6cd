"""argv contract of the tuning worker — a restatement of cmd/tuning/parser.py (HfArgumentParser over
Seq2SeqTrainingArguments + FinetuningArguments + ModelArguments + DataArguments) without transformers/accelerate.

The Finetune controller builds the command line (internal/controller/finetune/finetune_controller.go:451-516):

    python /tuning/train.py --model_name_or_path P --train_path T [--evaluation_path E] [--columns JSON]
        --output_dir result --deepspeed /tuning/ds_config.json --lora_target q_proj,v_proj --lr_scheduler_type S
        --optim O [--quantization int4|int8] --lora_r R --lora_alpha A --lora_dropout D --learning_rate LR
        --num_train_epochs N --block_size B --per_device_train_batch_size  BS --warmup_ratio W --weight_decay WD
        --gradient_accumulation_steps GA --fp16 true|false --num_workers N --storage_path SP
        [--metrics_export_address URL --uid UID]

Behaviours kept: argparse prefix abbreviation (`--lora_r` resolves to `--lora_rank`, parser.py:138), the double space
after `--per_device_train_batch_size` (finetune_controller.go:502) is harmless to a shell-split argv, unknown flags
fail the process (HfArgumentParser raises), `adamw_hf` is rewritten to `adamw_torch` (parser.py:255), `--storage_path`
and `--train_path` are required (parser.py:220-221,246-247), bools accept HF's string_to_bool spellings.
"""
from __future__ import annotations

import argparse
import json
from dataclasses import dataclass, field
from typing import Dict, List, Optional

_TRUE = {"yes", "true", "t", "y", "1"}
_FALSE = {"no", "false", "f", "n", "0"}


def string_to_bool(v):
    # transformers.hf_argparser.string_to_bool
    if isinstance(v, bool):
        return v
    if v.lower() in _TRUE:
        return True
    if v.lower() in _FALSE:
        return False
    raise argparse.ArgumentTypeError(f"Truthy value expected: got {v} but expected one of yes/no, true/false, t/f, y/n, 1/0")


class ArgError(SystemExit):
    pass


class _Parser(argparse.ArgumentParser):
    def error(self, message):
        self.print_usage()
        raise ArgError(f"error: {message}")


@dataclass
class TrainArgs:
    # ---- Seq2SeqTrainingArguments fields the worker reads (cmd/tuning/train.py:196-217) ----
    output_dir: str = "result"
    deepspeed: Optional[str] = None            # accepted; semantics = ZeRO stage 0 data parallel (ds_config.json)
    lr_scheduler_type: str = "linear"
    optim: str = "adamw_torch"
    learning_rate: float = 5e-5
    num_train_epochs: float = 3.0
    per_device_train_batch_size: int = 8
    per_device_eval_batch_size: int = 8
    warmup_ratio: float = 0.0                  # parsed, then DROPPED by the reference (train.py:204 passes warmup_steps)
    warmup_steps: int = 0
    weight_decay: float = 0.0
    gradient_accumulation_steps: int = 1
    fp16: bool = False
    bf16: bool = False
    seed: int = 42
    max_grad_norm: float = 1.0
    adam_beta1: float = 0.9
    adam_beta2: float = 0.999
    adam_epsilon: float = 1e-8
    max_steps: int = -1
    logging_steps: int = 10                    # the worker hard-codes 10 (train.py:198)
    # ---- FinetuningArguments (parser.py:113-221) ----
    finetuning_type: str = "lora"              # ignored by the reference: LoRA is always applied (train.py:277)
    lora_rank: int = 8
    lora_alpha: float = 32.0
    lora_dropout: float = 0.1
    lora_target: List[str] = field(default_factory=lambda: ["q_proj", "v_proj"])
    num_workers: int = 1
    storage_path: Optional[str] = None
    metrics_export_address: Optional[str] = None
    uid: Optional[str] = None
    # ---- ModelArguments (parser.py:13-110) ----
    model_name_or_path: str = ""
    quantization: Optional[str] = None
    # ---- DataArguments (parser.py:225-248) ----
    train_path: Optional[str] = None
    evaluation_path: Optional[str] = None
    columns: Optional[str] = None
    block_size: int = 1024

    def columns_map(self) -> Dict[str, str]:
        # cmd/tuning/train.py:326-335: defaults, then the CR's {"instruction": col, "response": col} inverted
        m = {"instruction": "instruction", "output": "response"}
        if self.columns:
            m.update({v: k for k, v in json.loads(self.columns).items()})
        return m


# flags that the reference dataclasses define but the worker never reads: accepted and ignored
_IGNORED_STR = ["cache_dir", "model_revision", "quantization_type", "rope_scaling", "checkpoint_dir", "reward_model",
                "hf_auth_token", "export_dir", "stage", "name_module_trainable", "additional_target", "ppo_logger",
                "logging_dir", "report_to", "run_name", "save_strategy", "evaluation_strategy", "log_level"]
_IGNORED_BOOL = ["use_fast_tokenizer", "split_special_tokens", "use_auth_token", "double_quantization", "flash_attn",
                 "shift_attn", "plot_loss", "resume_lora_training", "ppo_score_norm", "upcast_layernorm", "do_train", "do_eval",
                 "overwrite_output_dir", "gradient_checkpointing", "predict_with_generate"]
_IGNORED_NUM = ["quantization_bit", "num_layer_trainable", "ppo_target", "dpo_beta", "neft_alpha", "save_steps", "eval_steps",
                "local_rank"]


def build_parser() -> argparse.ArgumentParser:
    p = _Parser(prog="train.py", allow_abbrev=True)
    p.add_argument("--output_dir", type=str, required=True)
    p.add_argument("--deepspeed", type=str, default=None)
    p.add_argument("--lr_scheduler_type", type=str, default="linear")
    p.add_argument("--optim", type=str, default="adamw_torch")
    p.add_argument("--learning_rate", type=float, default=5e-5)
    p.add_argument("--num_train_epochs", type=float, default=3.0)
    p.add_argument("--per_device_train_batch_size", type=int, default=8)
    p.add_argument("--per_device_eval_batch_size", type=int, default=8)
    p.add_argument("--warmup_ratio", type=float, default=0.0)
    p.add_argument("--warmup_steps", type=int, default=0)
    p.add_argument("--weight_decay", type=float, default=0.0)
    p.add_argument("--gradient_accumulation_steps", type=int, default=1)
    p.add_argument("--fp16", type=string_to_bool, nargs="?", const=True, default=False)
    p.add_argument("--bf16", type=string_to_bool, nargs="?", const=True, default=False)
    p.add_argument("--seed", type=int, default=42)
    p.add_argument("--max_grad_norm", type=float, default=1.0)
    p.add_argument("--adam_beta1", type=float, default=0.9)
    p.add_argument("--adam_beta2", type=float, default=0.999)
    p.add_argument("--adam_epsilon", type=float, default=1e-8)
    p.add_argument("--max_steps", type=int, default=-1)
    p.add_argument("--logging_steps", type=float, default=10)
    p.add_argument("--finetuning_type", type=str, default="lora", choices=["lora", "freeze", "full", "none"])
    p.add_argument("--lora_rank", type=int, default=8)
    p.add_argument("--lora_alpha", type=float, default=32.0)
    p.add_argument("--lora_dropout", type=float, default=0.1)
    p.add_argument("--lora_target", type=str, default=None)
    p.add_argument("--num_workers", type=int, default=1)
    p.add_argument("--storage_path", type=str, default=None)
    p.add_argument("--metrics_export_address", type=str, default=None)
    p.add_argument("--uid", type=str, default=None)
    p.add_argument("--model_name_or_path", type=str, required=True)
    p.add_argument("--quantization", type=str, default=None)
    p.add_argument("--train_path", type=str, default=None)
    p.add_argument("--evaluation_path", type=str, default=None)
    p.add_argument("--columns", type=str, default=None)
    p.add_argument("--block_size", type=int, default=1024)
    for n in _IGNORED_STR:
        p.add_argument("--" + n, type=str, default=None)
    for n in _IGNORED_BOOL:
        p.add_argument("--" + n, type=string_to_bool, nargs="?", const=True, default=None)
    for n in _IGNORED_NUM:
        p.add_argument("--" + n, type=float, default=None)
    return p


def get_train_args(argv: Optional[List[str]] = None) -> TrainArgs:
    """Parse the controller's argv.  Raises ArgError (a SystemExit -> non-zero exit status) on unknown or
    malformed flags, mirroring HfArgumentParser.parse_args_into_dataclasses."""
    ns = build_parser().parse_args(argv)
    a = TrainArgs()
    for k in a.__dataclass_fields__:
        if hasattr(ns, k) and getattr(ns, k) is not None:
            setattr(a, k, getattr(ns, k))
    a.logging_steps = 10  # train.py:198 rebuilds TrainingArguments with logging_steps=10 whatever was passed
    if isinstance(ns.lora_target, str):  # parser.py:211-213
        a.lora_target = [t.strip() for t in ns.lora_target.split(",")]
    if a.optim == "adamw_hf":  # parser.py:255
        a.optim = "adamw_torch"
    if not a.storage_path:  # parser.py:220-221
        raise ArgError("ValueError: --storage_path must be specified")
    if a.train_path is None:  # parser.py:246-247
        raise ArgError("ValueError: --train_path must be specified")
    if a.lr_scheduler_type not in ("linear", "cosine", "constant", "constant_with_warmup"):
        raise ArgError(f"ValueError: lr_scheduler_type {a.lr_scheduler_type!r} is not implemented by the native worker")
    if a.optim not in ("adamw_torch", "adamw_torch_fused", "adamw_apex_fused", "adamw_anyprecision"):
        raise ArgError(f"ValueError: optim {a.optim!r} is not implemented by the native worker (AdamW family only)")
    return a


def controller_entrypoint(model_path: str, train_file: str, *, validate_file: str = "", columns: str = "", scheduler: str = "linear",
                          optimizer: str = "adamw_torch", int4: bool = False, int8: bool = False, lora_r: str = "8",
                          lora_alpha: str = "32", lora_dropout: str = "0.1", learning_rate: str = "5e-5", epochs: int = 1,
                          block_size: int = 1024, batch_size: int = 8, warmup_ratio: str = "0.1", weight_decay: str = "0.0",
                          grad_acc_steps: int = 1, fp16: bool = False, num_workers: int = 1, storage_path: str = "",
                          metrics_export_address: str = "", uid: str = "") -> str:
    """The entrypoint string exactly as getRayJobEntrypoint emits it (finetune_controller.go:451-516) — used by the
    tests to prove the native worker accepts the controller's own command line byte for byte."""
    e = ["python", "/tuning/train.py", "--model_name_or_path", model_path, "--train_path", train_file]
    if validate_file:
        e += ["--evaluation_path", validate_file]
    if columns:
        e += ["--columns", json.dumps(columns)]  # strconv.Quote
    e += ["--output_dir", "result", "--deepspeed", "/tuning/ds_config.json", "--lora_target", "q_proj,v_proj",
          "--lr_scheduler_type", scheduler, "--optim", optimizer]
    quant = "int8" if int8 else ("int4" if int4 else "")
    if quant:
        e += ["--quantization", quant]
    e += ["--lora_r", lora_r, "--lora_alpha", lora_alpha, "--lora_dropout", lora_dropout, "--learning_rate", learning_rate,
          "--num_train_epochs", str(epochs), "--block_size", str(block_size), "--per_device_train_batch_size ", str(batch_size),
          "--warmup_ratio", warmup_ratio, "--weight_decay", weight_decay, "--gradient_accumulation_steps", str(grad_acc_steps),
          "--fp16", "true" if fp16 else "false", "--num_workers", str(num_workers), "--storage_path", storage_path]
    if metrics_export_address:
        e += ["--metrics_export_address", metrics_export_address, "--uid", uid]
    return " ".join(e)
