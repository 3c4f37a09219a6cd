"""BASELINE.json configs[4]: a FineTuneExperiment fans out several FineTuneJobs at once (reference:
internal/controller/finetune/finetuneexperiment_controller.go:123-152).  On one 8-GPU box that is J worker processes
running side by side, each with its own GPUs, its own NCCL communicator and its own hyper-parameters.

This harness launches J jobs x G GPUs through the worker's OWN launch path - `python -m datatunerx_b200.tuning.train
<the controller's argv>` with `--num_workers G`, which re-executes itself once per extra rank with DTX_RANK /
DTX_NCCL_ID (datatunerx_b200/tuning/train.py:main) - each job confined to a disjoint CUDA_VISIBLE_DEVICES set, and
reports per-job and aggregate tokens/s.  Weights are random-init on the device (DTX_RANDOM_INIT: no checkpoint on the
box), the model directory only holds config.json + a tokenizer; the data is a synthetic instruction CSV long enough
that every row fills the block size.

  python tools/concurrent_jobs.py --jobs 4 --gpus-per-job 2 --model 7b --steps 12
  python tools/concurrent_jobs.py --jobs 2 --gpus-per-job 1 --model tiny --steps 6        (smoke)
"""
from __future__ import annotations

import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

MODELS = {
    "7b": dict(vocab_size=32000, hidden_size=4096, intermediate_size=11008, num_hidden_layers=32, num_attention_heads=32,
               num_key_value_heads=32, block=2048, batch=8),
    "tiny": dict(vocab_size=400, hidden_size=256, intermediate_size=768, num_hidden_layers=2, num_attention_heads=2,
                 num_key_value_heads=2, block=256, batch=4),
}


def make_model_dir(path: str, m: dict) -> None:
    os.makedirs(path, exist_ok=True)
    gold = os.path.join(ROOT, "tests", "golden")
    shutil.copy(os.path.join(gold, "tiny_tokenizer.json"), os.path.join(path, "tokenizer.json"))
    json.dump({"tokenizer_class": "PreTrainedTokenizerFast", "bos_token": "<s>", "eos_token": "</s>", "unk_token": "<unk>"},
              open(os.path.join(path, "tokenizer_config.json"), "w"))
    cfg = {k: v for k, v in m.items() if k not in ("block", "batch")}
    cfg.update(architectures=["LlamaForCausalLM"], rms_norm_eps=1e-5, rope_theta=10000.0, max_position_embeddings=4096, model_type="llama")
    json.dump(cfg, open(os.path.join(path, "config.json"), "w"))


def make_csv(path: str, rows: int, block: int, ragged: bool = False) -> None:
    # ~0.66 tokens per character with the tiny tokenizer: make every row longer than the block so that truncation fills it;
    # ragged: log-normal row lengths around a quarter of the block (the shape of real instruction data)
    import numpy as np
    rng = np.random.default_rng(11)
    qs, as_ = "Explain step by step how the following numbers add up and why the answer is what it is. ", \
        "The answer follows from adding the numbers one after another and carrying where needed. "
    with open(path, "w") as f:
        f.write("q,a\n")
        for i in range(rows):
            scale = float(np.clip(np.exp(rng.normal(np.log(0.25), 0.6)), 0.02, 1.5)) if ragged else 1.5
            q = qs * max(1, int(scale * block / 120))
            a = as_ * max(1, int(scale * block / 60))
            f.write(f"\"{i}: {q}\",\"{i}: {a}\"\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=4)
    ap.add_argument("--gpus-per-job", type=int, default=2)
    ap.add_argument("--model", default="7b", choices=sorted(MODELS))
    ap.add_argument("--steps", type=int, default=12, help="optimizer steps per job (one epoch)")
    ap.add_argument("--ragged", action="store_true", help="log-normal row lengths instead of rows that fill the block")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "concurrent_jobs.json"))
    args = ap.parse_args()
    from datatunerx_b200.tuning import parser as TP

    m = MODELS[args.model]
    tmp = tempfile.mkdtemp(prefix="dtx_jobs_")
    mdir = os.path.join(tmp, "model")
    make_model_dir(mdir, m)
    rows = args.steps * m["batch"] * args.gpus_per_job
    csv_path = os.path.join(tmp, "train.csv")
    make_csv(csv_path, rows, m["block"], args.ragged)
    # the hyper-parameter sweep of the experiment: one (lr, lora_r, scheduler) per job
    sweep = [("1e-4", "16", "linear"), ("2e-4", "16", "cosine"), ("5e-5", "8", "linear"), ("1e-4", "32", "cosine"),
             ("3e-4", "16", "linear"), ("1e-4", "8", "cosine"), ("2e-4", "32", "linear"), ("5e-5", "16", "cosine")]
    procs, t0 = [], time.time()
    for j in range(args.jobs):
        lr, r, sched = sweep[j % len(sweep)]
        jdir = os.path.join(tmp, f"job{j}")
        os.makedirs(jdir)
        entry = TP.controller_entrypoint(mdir, csv_path, columns='{"instruction":"q","response":"a"}', scheduler=sched, optimizer="adamw_torch",
                                         lora_r=r, lora_alpha="32", lora_dropout="0.0", learning_rate=lr, epochs=1, block_size=m["block"],
                                         batch_size=m["batch"], grad_acc_steps=1, num_workers=args.gpus_per_job,
                                         storage_path=os.path.join(jdir, "storage"), uid=f"job{j}")
        import shlex
        argv = shlex.split(entry)[2:]
        gpus = ",".join(str(j * args.gpus_per_job + g) for g in range(args.gpus_per_job))
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=gpus, DTX_RANDOM_INIT="1234", DTX_CHECKPOINT_PATH_FILE=os.path.join(jdir, "checkpoint_path"),
                   PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
        log = open(os.path.join(jdir, "stdout.log"), "w")
        p = subprocess.Popen([sys.executable, "-m", "datatunerx_b200.tuning.train"] + argv, cwd=jdir, env=env, stdout=log, stderr=subprocess.STDOUT)
        procs.append((j, p, jdir, gpus, dict(lr=lr, lora_r=r, sched=sched)))
    jobs = []
    for j, p, jdir, gpus, hp in procs:
        rc = p.wait()
        out = open(os.path.join(jdir, "stdout.log")).read()
        mt = re.search(r"train_runtime ([0-9.]+)s, (\d+) optimizer steps, (\d+) real tokens on rank 0", out)
        ck = os.path.join(jdir, "checkpoint_path")
        rec = {"job": j, "gpus": gpus, "rc": rc, "hyperparameters": hp, "checkpoint_written": os.path.exists(ck)}
        if mt:
            dt, steps, toks = float(mt.group(1)), int(mt.group(2)), int(mt.group(3))
            rec.update(train_runtime_s=dt, optimizer_steps=steps, tokens_per_s=toks * args.gpus_per_job / dt)
            mg = re.search(r"([0-9.]+) length groups per micro-batch", out)
            if mg:
                rec["length_groups_per_micro_batch"] = float(mg.group(1))
            ms = re.search(r"steady_state ([0-9.]+)s, (\d+) optimizer steps, (\d+) real tokens on rank 0", out)
            if ms:  # from the end of the first optimizer step to the end of the last: no lazy initialisation, no checkpoint writes
                rec.update(steady_state_s=float(ms.group(1)), steady_tokens_per_s=int(ms.group(3)) * args.gpus_per_job / float(ms.group(1)))
            logs = os.path.join(jdir, "result", "watch", "trainer_log.jsonl")
            if os.path.exists(logs):
                rec["logged_losses"] = [json.loads(l)["loss"] for l in open(logs)]
        else:
            rec["tail"] = out[-800:]
        jobs.append(rec)
    wall = time.time() - t0
    ok = all(r["rc"] == 0 and r["checkpoint_written"] for r in jobs)
    res = {"config": f"{args.jobs} concurrent {args.model} LoRA jobs x {args.gpus_per_job} GPUs, hyper-parameter sweep (BASELINE.json configs[4])",
           "ok": ok, "wall_s": wall, "aggregate_tokens_per_s": sum(r.get("tokens_per_s", 0.0) for r in jobs),
           "aggregate_steady_tokens_per_s": sum(r.get("steady_tokens_per_s", 0.0) for r in jobs), "jobs": jobs}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
    print("CONCURRENT_JOBS " + json.dumps(res))
    shutil.rmtree(tmp, ignore_errors=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
