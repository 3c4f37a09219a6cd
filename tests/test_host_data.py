"""Host-side integer path (CSV -> llama2 template -> masked labels -> batches): bit-exact against golden vectors
produced by the reference's own template.py / preprocess_dataset (tests/golden/make_template_golden.py)."""
import json
import os

import numpy as np
import pytest

from datatunerx_b200.tuning import data as D
from datatunerx_b200.tuning import parser as P

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _tokenizer():
    from tokenizers import Tokenizer
    from transformers import PreTrainedTokenizerFast
    tok = Tokenizer.from_file(os.path.join(GOLD, "tiny_tokenizer.json"))
    return PreTrainedTokenizerFast(tokenizer_object=tok, bos_token="<s>", eos_token="</s>", unk_token="<unk>")


def test_llama2_template_and_label_masking_match_reference_goldens():
    gold = json.load(open(os.path.join(GOLD, "llama2_template.json")))
    tok = _tokenizer()
    D.fix_tokenizer(tok)
    assert len(gold["cases"]) >= 5
    for case in gold["cases"]:
        got = D.build_dataset(case["rows"], tok, case["cutoff_len"])
        assert [g[0] for g in got] == case["input_ids"]
        assert [g[1] for g in got] == case["labels"]
        for ids, labels in got:
            assert len(ids) == len(labels) <= case["cutoff_len"]


def test_collate_static_shape_and_padding():
    ex = [([1, 5, 6, 2], [-100, -100, 6, 2]), ([1, 7, 2], [-100, 7, 2])]
    ids, lab = D.collate(ex, 128, pad_id=2)
    assert ids.shape == lab.shape == (2, 128) and ids.dtype == np.int32
    assert ids[0, :4].tolist() == [1, 5, 6, 2] and (ids[0, 4:] == 2).all() and (lab[0, 4:] == -100).all()
    assert D.static_seq_len(1024) == 1024 and D.static_seq_len(1000) == 1024 and D.static_seq_len(16) == 128


def test_sharding_is_disjoint_and_equal_across_ranks():
    for n, world in [(10, 2), (17, 4), (8, 8), (3, 2)]:
        shards = [D.shard_indices(n, r, world) for r in range(world)]
        assert len({len(s) for s in shards}) == 1
        flat = [i for s in shards for i in s]
        assert len(flat) == len(set(flat)) == (n // world) * world


def test_epoch_batches_same_permutation_on_every_rank():
    ds = [([i, i], [i, i]) for i in range(20)]
    seen = []
    for r in range(2):
        for ids, _ in D.epoch_batches(ds, r, 2, 2, 128, 0, epoch=1, seed=42):
            seen += ids[:, 0].tolist()
    assert sorted(seen) == sorted(set(seen)) and len(seen) == 20


def test_last_partial_batch_is_kept_and_step_counts_follow_hf():
    """HF Trainer keeps the last partial batch (dataloader_drop_last=False) and the reference computes ceil(len / batch)
    (cmd/tuning/train.py:192-193): 11 examples on one rank at batch 4 -> 3 batches, the third with one filler row."""
    from datatunerx_b200.tuning.train import total_optimizer_steps
    ds = [([5 + i] * (3 + i), [-100] + [5 + i] * (2 + i)) for i in range(11)]
    batches = list(D.epoch_batches(ds, 0, 1, 4, 256, 0, epoch=0, seed=1, varlen=True))
    assert len(batches) == D.steps_per_epoch(11, 1, 4) == 3
    ids, lab, lens = batches[-1]
    assert ids.shape == (4, 128) and lens.tolist()[3] == 0 and (lab[3] == D.IGNORE_INDEX).all() and (lens[:3] > 0).all()
    assert sorted(int(x[0]) for b in batches for x, n in zip(b[0], b[2]) if n > 0) == [5 + i for i in range(11)]
    # every batch is padded to its own longest row rounded up to the 128-row tile, never beyond the static length
    assert all(b[0].shape[1] == 128 for b in batches)
    long_ds = [(list(range(3, 3 + 300)), list(range(3, 3 + 300))), ([1, 2], [-100, 2])]
    (i2, l2, n2), = list(D.epoch_batches(long_ds, 0, 1, 2, 512, 0, epoch=0, seed=0, varlen=True))
    assert i2.shape == (2, 384) and sorted(n2.tolist()) == [2, 300]
    assert D.batch_seq_len([5000], 2048) == 2048 and D.batch_seq_len([1], 2048) == 128 and D.batch_seq_len([129, 7], 2048) == 256
    # two ranks: Ray's equal split drops the odd example, each rank then has ceil(5 / 4) = 2 batches
    assert D.steps_per_epoch(11, 2, 4) == 2 and len(list(D.epoch_batches(ds, 1, 2, 4, 256, 0, epoch=0, seed=1))) == 2
    # HF: num_update_steps_per_epoch = max(len(dataloader) // GA, 1); max_steps = ceil(epochs * that)
    assert total_optimizer_steps(11, 1, 4, 1, 2, -1) == 6 and total_optimizer_steps(11, 1, 4, 2, 3, -1) == 3
    assert total_optimizer_steps(3, 1, 4, 4, 2, -1) == 2 and total_optimizer_steps(1000, 8, 8, 1, 1, 7) == 7
    # the reference's own known answer: 84 total steps in cmd/tuning/prometheus/metrics.py:117-124 style runs (ceil semantics)
    assert total_optimizer_steps(84 * 8 - 3, 1, 8, 1, 1, -1) == 84


def test_controller_entrypoint_is_accepted_verbatim():
    # the exact string getRayJobEntrypoint builds (finetune_controller.go:451-516), including the double space
    s = P.controller_entrypoint("/tmp/llama2-7b/", "/data/train.csv", validate_file="/data/val.csv",
                                columns='{"instruction":"q","response":"a"}', scheduler="cosine", optimizer="adamw_hf", lora_r="16",
                                lora_alpha="32", lora_dropout="0.05", learning_rate="1e-4", epochs=2, block_size=2048, batch_size=8,
                                warmup_ratio="0.1", weight_decay="0.01", grad_acc_steps=2, fp16=True, num_workers=8,
                                storage_path="s3://bucket/ckpt", metrics_export_address="http://prom:9090", uid="abc-123")
    assert "--per_device_train_batch_size  8" in s  # the controller's trailing-space quirk
    import shlex
    argv = shlex.split(s)[2:]
    a = P.get_train_args(argv)
    assert a.lora_rank == 16 and a.lora_alpha == 32.0 and a.lora_dropout == 0.05  # --lora_r abbreviates --lora_rank
    assert a.lora_target == ["q_proj", "v_proj"] and a.optim == "adamw_torch" and a.lr_scheduler_type == "cosine"
    assert a.per_device_train_batch_size == 8 and a.gradient_accumulation_steps == 2 and a.fp16 is True
    assert a.num_workers == 8 and a.block_size == 2048 and a.num_train_epochs == 2
    assert a.columns_map() == {"instruction": "instruction", "output": "response", "q": "instruction", "a": "response"}
    assert a.logging_steps == 10 and a.warmup_steps == 0 and a.warmup_ratio == 0.1  # ratio parsed, never used


def test_unknown_flag_and_missing_required_fail_like_hf_argparser():
    base = ["--model_name_or_path", "m", "--train_path", "t", "--output_dir", "o", "--storage_path", "s"]
    P.get_train_args(base)
    with pytest.raises(SystemExit):
        P.get_train_args(base + ["--no_such_flag", "1"])
    with pytest.raises(SystemExit):
        P.get_train_args(base[:-2])
    with pytest.raises(SystemExit):
        P.get_train_args(["--model_name_or_path", "m", "--output_dir", "o", "--storage_path", "s"])


def test_empty_ragged_and_oversized_inputs(tmp_path):
    """Edge cases of the integer path: skipped rows, an empty shard, ragged batches, rows longer than the static length,
    a batch whose labels are all masked, CSV column renaming."""
    tok = _tokenizer()
    D.fix_tokenizer(tok)
    # rows with an empty / missing side are skipped exactly like train.py:83-84 (no exception)
    rows = [{"instruction": "", "response": "x"}, {"instruction": "hello", "response": ""}, {"instruction": None, "response": "y"},
            {"instruction": "hello there", "response": "general kenobi"}]
    ds = D.build_dataset(rows, tok, 64)
    assert len(ds) == 1 and len(ds[0][0]) == len(ds[0][1]) <= 64
    assert D.build_dataset([], tok, 64) == []
    # an empty shard yields no batches (the worker refuses to start on it); fewer examples than one batch yield ONE batch
    # filled up with fully masked rows (HF keeps the last partial batch)
    assert list(D.epoch_batches([], 0, 2, 4, 128, tok.eos_token_id, seed=0, epoch=0)) == []
    assert D.steps_per_epoch(3, 2, 4) == 1 and D.steps_per_epoch(1, 2, 4) == 0
    (ids_t, lab_t), = list(D.epoch_batches(ds * 3, 0, 2, 4, 128, tok.eos_token_id, seed=0, epoch=0))
    assert ids_t.shape == (4, 128) and (lab_t[1:] == D.IGNORE_INDEX).all() and (lab_t[0] != D.IGNORE_INDEX).any()
    # ragged lengths collate to one static shape; an example longer than the static length is cut, never overflows
    long_ids = list(range(3, 3 + 300))
    ids, lab = D.collate([(long_ids, long_ids), ([1, 2], [-100, 2])], 128, pad_id=0)
    assert ids.shape == (2, 128) and ids[0].tolist() == long_ids[:128] and lab[0].tolist() == long_ids[:128]
    assert ids[1, :2].tolist() == [1, 2] and (ids[1, 2:] == 0).all() and (lab[1, 2:] == D.IGNORE_INDEX).all()
    # a fully masked example stays fully masked (the step then sees n_valid = 0 for it)
    _, lab2 = D.collate([([1, 5, 6], [-100, -100, -100])], 128, pad_id=0)
    assert (lab2 == D.IGNORE_INDEX).all()
    # --columns renaming: CSV header names -> instruction / response
    p = tmp_path / "d.csv"
    p.write_text("q,a\n\"what, exactly?\",\"this\"\n")
    assert D.read_csv_rows(str(p), {"q": "instruction", "a": "response"}) == [{"instruction": "what, exactly?", "response": "this"}]


def test_collation_matches_the_installed_data_collator():
    """a12: DataCollatorForSeq2Seq(pad_to_multiple_of=4, label_pad_token_id=-100) (cmd/tuning/train.py:282-286) against
    tuning.data.collate on the same examples: identical ids / labels over the collator's own width, and only padding
    (pad id, -100) beyond it.  The installed transformers (5.x) stands in for the reference's pinned 4.34.0."""
    import os
    from transformers import DataCollatorForSeq2Seq, PreTrainedTokenizerFast
    from datatunerx_b200.tuning import data as D
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiny_tokenizer.json")
    tok = PreTrainedTokenizerFast(tokenizer_file=gold, bos_token="<s>", eos_token="</s>", unk_token="<unk>")
    D.fix_tokenizer(tok)
    assert tok.padding_side == "right"
    rng = np.random.default_rng(3)
    examples = []
    for n in (37, 5, 130, 64):
        x = rng.integers(3, 300, size=n).tolist()
        y = [-100] * (n // 3) + x[n // 3:]
        examples.append((x, y))
    hf = DataCollatorForSeq2Seq(tokenizer=tok, pad_to_multiple_of=4, label_pad_token_id=-100)(
        [{"input_ids": x, "attention_mask": [1] * len(x), "labels": y} for x, y in examples], return_tensors="np")
    width = hf["input_ids"].shape[1]
    assert width == 132  # longest row 130 rounded up to a multiple of 4
    cur = D.batch_seq_len([len(x) for x, _ in examples], 2048)
    assert cur == 256  # the native step rounds to the 128-row attention tile instead and passes the true lengths
    ids, lab = D.collate(examples, cur, tok.pad_token_id)
    assert np.array_equal(ids[:, :width], hf["input_ids"]) and np.array_equal(lab[:, :width], hf["labels"])
    assert (ids[:, width:] == tok.pad_token_id).all() and (lab[:, width:] == -100).all()
    assert [int(r.sum()) for r in hf["attention_mask"]] == [len(x) for x, _ in examples]  # = the seq_lens handed to dtx_step
