"""Generates tests/golden/llama2_template.json by running the REFERENCE's own code (only possible in the build
container, where /root/reference exists): cmd/tuning/template.py is imported as-is, and `preprocess_dataset` is
extracted from cmd/tuning/train.py by AST (train.py itself imports ray/peft and cannot be imported) and executed
unmodified against a stand-in dataset object.  The tokenizer is a small byte-level BPE trained here and committed
as tests/golden/tiny_tokenizer.json so that the GPU box can rebuild the same ids without the reference."""
import ast
import json
import os
import sys
import types

REF = "/root/reference/cmd/tuning"
HERE = os.path.dirname(os.path.abspath(__file__))


def build_tokenizer():
    from tokenizers import Tokenizer, decoders, models, pre_tokenizers, trainers
    tok = Tokenizer(models.BPE(unk_token=None))
    tok.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False)
    tok.decoder = decoders.ByteLevel()
    corpus = [
        "You are a helpful, respectful and honest assistant. Always answer as helpfully as possible, while being safe.",
        "[INST] <<SYS>>\n\n<</SYS>>\n\n [/INST] ", "What is the capital of France? The capital of France is Paris.",
        "Write a haiku about GPUs. Silicon rivers / tensor cores hum in the night / gradients descend.",
        "instruction response query system 0123456789 abcdefghijklmnopqrstuvwxyz ABCDEFGHIJKLMNOPQRSTUVWXYZ",
    ] * 4
    trainer = trainers.BpeTrainer(vocab_size=400, special_tokens=["<unk>", "<s>", "</s>"],
                                  initial_alphabet=pre_tokenizers.ByteLevel.alphabet())
    tok.train_from_iterator(corpus, trainer)
    return tok


def wrap(tok):
    from transformers import PreTrainedTokenizerFast

    class Shim(PreTrainedTokenizerFast):
        # transformers 5.x dropped the `replace_additional_special_tokens` kwarg the reference (4.34) passes with an
        # empty stop-word list for llama2 (template.py:218-221): a no-op either way
        def add_special_tokens(self, special_tokens_dict, **kw):
            kw.pop("replace_additional_special_tokens", None)
            if not special_tokens_dict.get("additional_special_tokens"):
                return 0
            return super().add_special_tokens(special_tokens_dict, **kw)

    return Shim(tokenizer_object=tok, bos_token="<s>", eos_token="</s>", unk_token="<unk>")


class FakeDataset:
    def __init__(self, batch):
        self.batch = batch

    def map_batches(self, fn):
        return FakeDataset(fn(self.batch))

    def take(self, n):
        keys = list(self.batch)
        return [{k: self.batch[k][0] for k in keys}]


def reference_preprocess(tokenizer, rows, cutoff_len):
    sys.path.insert(0, REF)
    import template as ref_template  # the reference's own file
    src = open(os.path.join(REF, "train.py")).read()
    tree = ast.parse(src)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "preprocess_dataset"][0]
    mod = types.ModuleType("ref_train_slice")
    g = mod.__dict__
    from typing import Any, Dict, Generator, List, Union
    g.update(dict(Union=Union, Dict=Dict, List=List, Any=Any, Generator=Generator, IGNORE_INDEX=-100, cutoff_len=cutoff_len,
                  get_template_and_fix_tokenizer=ref_template.get_template_and_fix_tokenizer))
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "train.py:preprocess_dataset", "exec"), g)
    batch = {k: [r.get(k) for r in rows] for k in sorted({k for r in rows for k in r})}
    targs = types.SimpleNamespace(should_log=False)
    out = g["preprocess_dataset"](FakeDataset(batch), tokenizer, targs).batch
    return out["input_ids"], out["labels"]


def main():
    tok = build_tokenizer()
    open(os.path.join(HERE, "tiny_tokenizer.json"), "w").write(tok.to_str())
    hf = wrap(tok)
    cases = []
    long_q = "Explain " + "the gradient of the loss with respect to the adapter " * 12
    long_r = "Because " + "the base weights are frozen and only A and B move " * 9
    specs = [
        (256, [{"instruction": "What is the capital of France?", "response": "The capital of France is Paris."}]),
        (64, [{"instruction": long_q, "response": long_r}]),
        (48, [{"instruction": "short", "response": long_r}, {"instruction": long_q, "response": "ok"}]),
        (128, [{"instruction": "Write a haiku", "response": "", }, {"instruction": "", "response": "x"},
               {"instruction": "Write a haiku", "response": "Silicon rivers", "query": "about GPUs"}]),
        (128, [{"instruction": "Translate", "response": "Bonjour", "system": "You are a translator."}]),
    ]
    for cutoff, rows in specs:
        ids, labels = reference_preprocess(hf, rows, cutoff)
        cases.append({"cutoff_len": cutoff, "rows": rows, "input_ids": ids, "labels": labels})
    json.dump({"generator": "tests/golden/make_template_golden.py", "reference": "cmd/tuning/template.py + train.py:58-135 @ 508be30",
               "cases": cases}, open(os.path.join(HERE, "llama2_template.json"), "w"), indent=0)
    print("cases:", [(c["cutoff_len"], [len(x) for x in c["input_ids"]]) for c in cases])


if __name__ == "__main__":
    main()
