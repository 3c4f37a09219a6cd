"""Rank-0 log side channel — restates LogCallback (cmd/tuning/callback.py:20-155): every `logging_steps` optimizer
steps a JSON line is appended to <output_dir>/watch/trainer_log.jsonl (eval: eval_log.jsonl) with the reference's
keys in the reference's order, and the same dict is pushed through the Prometheus exporter when
--metrics_export_address is set."""
from __future__ import annotations

import json
import os
import time
from datetime import timedelta
from typing import Dict, Optional

from . import metrics as M


class LogCallback:
    def __init__(self, output_dir: str, max_steps: int, metrics_export_address: Optional[str] = None, uid: Optional[str] = None,
                 blocking_export: bool = False):
        self.output_dir, self.max_steps = output_dir, max_steps
        self.metrics_export_address, self.uid = metrics_export_address, uid
        self.start_time = time.time()
        self.cur_steps = 0
        self.elapsed_time = self.remaining_time = ""
        self.blocking_export = blocking_export

    def timing(self) -> None:  # callback.py:31-37
        elapsed = time.time() - self.start_time
        avg = elapsed / self.cur_steps if self.cur_steps != 0 else 0
        self.elapsed_time = str(timedelta(seconds=int(elapsed)))
        self.remaining_time = str(timedelta(seconds=int((self.max_steps - self.cur_steps) * avg)))

    def on_step_end(self, global_step: int) -> None:  # callback.py:72-79
        self.cur_steps = global_step
        self.timing()

    def _pct(self) -> float:
        return round(self.cur_steps / self.max_steps * 100, 2) if self.max_steps != 0 else 100

    def _append(self, name: str, row: Dict) -> None:
        d = os.path.join(self.output_dir, "watch")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, name), "a", encoding="utf-8") as f:
            f.write(json.dumps(row) + "\n")

    def on_log(self, loss: float, learning_rate: float, epoch: float) -> Dict:
        """HF rounds the logged loss to 4 decimals and the epoch to 2 (Trainer._maybe_log_save_evaluate / log)."""
        logs = dict(uid=self.uid, current_steps=self.cur_steps, total_steps=self.max_steps, loss=round(loss, 4), eval_loss=None,
                    val_perplexity=None, eval_rouge_1=None, eval_rouge_2=None, eval_rouge_l=None, eval_bleu_4=None, predict_loss=None,
                    reward=None, learning_rate=learning_rate, epoch=round(epoch, 2), percentage=self._pct(),
                    elapsed_time=self.elapsed_time, remaining_time=self.remaining_time)
        print("log_history: ", {"loss": logs["loss"], "learning_rate": learning_rate, "epoch": logs["epoch"]})
        self._append("trainer_log.jsonl", logs)
        if self.metrics_export_address:
            M.export_train_metrics(self.metrics_export_address, logs, blocking=self.blocking_export)
        return logs

    def on_eval(self, eval_loss: float, eval_perplexity: float, epoch: float) -> Dict:
        row = dict(uid=self.uid, current_steps=self.cur_steps, total_steps=self.max_steps, eval_loss=eval_loss,
                   eval_perplexity=eval_perplexity, eval_rouge_1=None, eval_rouge_2=None, eval_rouge_l=None, eval_bleu_4=None,
                   epoch=round(epoch, 2), percentage=self._pct(), elapsed_time=self.elapsed_time, remaining_time=self.remaining_time)
        self._append("eval_log.jsonl", row)
        if self.metrics_export_address:
            M.export_eval_metrics(self.metrics_export_address, row, blocking=self.blocking_export)
        return row
