"""Prometheus remote-write side channel — restates cmd/tuning/prometheus/metrics.py:21-113 without the
`snappy` and generated-protobuf dependencies: the WriteRequest (prometheus.proto:19-57) is encoded by hand and
wrapped in a valid snappy *block* stream made of literal chunks (any snappy decoder accepts it).

Series (labels carry the values, the sample value is the constant 1, timestamp in ms):
  train_metrics{uid,total_steps,current_steps,loss,learning_rate,epoch}    metrics.py:42-76
  eval_metrics{uid,total_steps,current_steps,eval_loss,eval_perplexity,epoch}  metrics.py:79-113
Differences kept deliberate: the POST runs on a daemon thread with a timeout — the reference blocks the rank-0
training thread on `requests.post` with no timeout; errors are swallowed and printed, as there.
"""
from __future__ import annotations

import struct
import threading
import time
import urllib.request
from typing import Dict, List, Tuple
from urllib.parse import urljoin


def _varint(n: int) -> bytes:
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _ld(field: int, payload: bytes) -> bytes:  # length-delimited field
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def encode_timeseries(labels: List[Tuple[str, str]], value: float, timestamp_ms: int) -> bytes:
    body = b"".join(_ld(1, _ld(1, n.encode()) + _ld(2, v.encode())) for n, v in labels)
    sample = _varint((1 << 3) | 1) + struct.pack("<d", value) + _varint((2 << 3) | 0) + _varint(timestamp_ms & (2 ** 64 - 1))
    return body + _ld(2, sample)


def encode_write_request(series: List[bytes]) -> bytes:
    return b"".join(_ld(1, s) for s in series)


def snappy_block_literal(data: bytes) -> bytes:
    """Snappy block format: varint(uncompressed length) then elements; here only literals (tag low bits 00)."""
    out = bytearray(_varint(len(data)))
    i = 0
    while i < len(data):
        chunk = data[i:i + 65536]
        n = len(chunk) - 1
        if n < 60:
            out.append(n << 2)
        elif n < 256:
            out += bytes([60 << 2, n])
        else:
            out += bytes([61 << 2, n & 0xFF, n >> 8])
        out += chunk
        i += len(chunk)
    return bytes(out)


def snappy_block_decode(buf: bytes) -> bytes:
    """Minimal decoder (literals + copies) used by the tests to prove the stream is valid snappy."""
    n, shift, i = 0, 0, 0
    while True:
        b = buf[i]
        i += 1
        n |= (b & 0x7F) << shift
        shift += 7
        if not b & 0x80:
            break
    out = bytearray()
    while i < len(buf):
        tag = buf[i]
        i += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[i:i + nb], "little")
                i += nb
            ln += 1
            out += buf[i:i + ln]
            i += ln
        else:
            if kind == 1:
                ln = ((tag >> 2) & 7) + 4
                off = ((tag >> 5) << 8) | buf[i]
                i += 1
            elif kind == 2:
                ln = (tag >> 2) + 1
                off = int.from_bytes(buf[i:i + 2], "little")
                i += 2
            else:
                ln = (tag >> 2) + 1
                off = int.from_bytes(buf[i:i + 4], "little")
                i += 4
            for _ in range(ln):
                out.append(out[-off])
    assert len(out) == n
    return bytes(out)


def _labels(name: str, metrics: Dict, keys: List[str]) -> List[Tuple[str, str]]:
    out = [("__name__", name), ("uid", str(metrics["uid"]))]
    out += [(k, str(metrics.get(k, ""))) for k in keys]
    return out


def train_series(metrics: Dict, now_ms: int) -> bytes:
    return encode_timeseries(_labels("train_metrics", metrics, ["total_steps", "current_steps", "loss", "learning_rate", "epoch"]),
                             1.0, now_ms)


def eval_series(metrics: Dict, now_ms: int) -> bytes:
    return encode_timeseries(_labels("eval_metrics", metrics, ["total_steps", "current_steps", "eval_loss", "eval_perplexity", "epoch"]),
                             1.0, now_ms)


def write(address: str, series: List[bytes], timeout: float = 5.0, blocking: bool = False) -> None:
    body = snappy_block_literal(encode_write_request(series))
    url = urljoin(address, "/api/v1/write")
    headers = {"Content-Encoding": "snappy", "Content-Type": "application/x-protobuf",
               "X-Prometheus-Remote-Write-Version": "0.1.0", "User-Agent": "metrics-worker"}

    def post():
        try:
            req = urllib.request.Request(url, data=body, headers=headers, method="POST")
            with urllib.request.urlopen(req, timeout=timeout) as r:
                print(f"<Response [{r.status}]>")
        except Exception as e:  # swallowed like metrics.py:35-39
            print(e)

    if blocking:
        post()
    else:
        threading.Thread(target=post, daemon=True).start()


def export_train_metrics(address: str, metrics: Dict, **kw) -> None:
    write(address, [train_series(metrics, int(time.time()) * 1000)], **kw)


def export_eval_metrics(address: str, metrics: Dict, **kw) -> None:
    write(address, [eval_series(metrics, int(time.time()) * 1000)], **kw)
