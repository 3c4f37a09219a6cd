"""Discrete-event model of one SM running a flash-attention CTA: one in-order tensor pipe fed by a single issuer warp,
groups of compute warps, mbarrier hand-offs with the latencies measured in profiles/r01_ubench_mma_latency.txt and
profiles/r01_attn_phase_timing.txt.  Used to compare pipeline layouts (buffers, block width, where operands live) before
spending GPU time on them; the absolute numbers are only as good as the constants, the ranking is what matters.

  python tools/attn_pipe_model.py
"""
from __future__ import annotations

import heapq
import itertools

L_COMMIT = 290    # last MMA of a group done -> a thread waiting on the committed mbarrier runs again
L_ARRIVE = 100    # last thread arrival -> the issuer warp waiting on that mbarrier runs again
T_LD, T_ST = 30, 45  # tcgen05.ld of a score block / tcgen05.st + wait + fence + arrive
ISSUE_SLACK = 40  # an issue call returns this long before its group finishes executing (the pipe's queue is shallow)


class Sim:
    def __init__(self):
        self.t, self.q, self.n = 0.0, [], itertools.count()
        self.bar = {}      # name -> list of completion times (phase k completes at bar[name][k])
        self.waiters = {}  # (name, k) -> [generator]
        self.pipe_free = 0.0
        self.busy = 0.0

    def at(self, t, gen):
        heapq.heappush(self.q, (t, next(self.n), gen))

    def complete(self, name, t):
        self.bar.setdefault(name, []).append(t)
        k = len(self.bar[name]) - 1
        for g in self.waiters.pop((name, k), []):
            self.at(t, g)

    def run(self, procs):
        for p in procs:
            self.at(0.0, p)
        while self.q:
            t, _, g = heapq.heappop(self.q)
            self.t = max(self.t, t)
            try:
                op = next(g)
            except StopIteration:
                continue
            kind = op[0]
            if kind == "sleep":
                self.at(t + op[1], g)
            elif kind == "wait":  # wait for phase k of barrier name
                _, name, k = op
                done = self.bar.get(name, [])
                if k < len(done):
                    self.at(max(t, done[k]), g)
                else:
                    self.waiters.setdefault((name, k), []).append(g)
            elif kind == "mma":  # issue a group of MMAs: (dur, [barriers committed at completion])
                _, dur, commits = op
                start = max(t, self.pipe_free)
                end = start + dur
                self.pipe_free = end
                self.busy += dur
                for name in commits:
                    self.at(end + L_COMMIT, self._signal(name, end + L_COMMIT))
                self.at(max(t, end - ISSUE_SLACK), g)
            elif kind == "signal":  # thread-side arrive: visible to waiters after L_ARRIVE
                self.at(t + L_ARRIVE, self._signal(op[1], t + L_ARRIVE))
                self.at(t, g)
        return self.t

    def _signal(self, name, t):
        def f():
            self.complete(name, t)
            return
            yield
        return f()


# ------------------------------------------------------------------------------------------------------------------
# layouts.  n = number of KV blocks of one tile; all durations in SM cycles per block.
# ------------------------------------------------------------------------------------------------------------------
def two_buffer_inplace(n, t_score, t_acc, t_math, prologue=0, epilogue=0):
    """Current dQ / dK-dV kernels: two score buffers, P / dS written over the scores they came from.
    issuer: S(0) S(1); for j: wait P(j) -> acc(j) -> S(j+2)."""
    def issuer():
        yield ("sleep", prologue)
        yield ("mma", t_score, ["s0"])
        if n > 1:
            yield ("mma", t_score, ["s1"])
        for j in range(n):
            yield ("wait", f"p{j % 2}", j // 2)
            yield ("mma", t_acc, ["fin"] if j == n - 1 else [])
            if j + 2 < n:
                yield ("mma", t_score, [f"s{j % 2}"])

    def threads():
        for j in range(n):
            yield ("wait", f"s{j % 2}", j // 2)
            yield ("sleep", T_LD + t_math + T_ST)
            yield ("signal", f"p{j % 2}")
        yield ("wait", "fin", 0)
        yield ("sleep", epilogue)
    return [issuer(), threads()]


def three_stage(n, t_score, t_acc, t_math, n_sbuf=2, n_pbuf=2, prologue=0, epilogue=0):
    """Score buffers are released as soon as the threads have LOADED them; P / dS go to their own small double buffer.
    issuer serves whichever is ready in program order: acc(j-?) and S(j+n_sbuf)."""
    def issuer():
        yield ("sleep", prologue)
        for j in range(min(n_sbuf, n)):
            yield ("mma", t_score, [f"s{j % n_sbuf}"])
        for j in range(n):
            # refill the score buffer of block j (freed when the threads have loaded it) before waiting for P(j):
            if j + n_sbuf < n:
                yield ("wait", f"r{j % n_sbuf}", j // n_sbuf)
                yield ("mma", t_score, [f"s{j % n_sbuf}"])
            yield ("wait", f"p{j % n_pbuf}", j // n_pbuf)
            yield ("mma", t_acc, [f"c{j % n_pbuf}"] + (["fin"] if j == n - 1 else []))

    def threads():
        for j in range(n):
            yield ("wait", f"s{j % n_sbuf}", j // n_sbuf)
            yield ("sleep", T_LD)
            yield ("signal", f"r{j % n_sbuf}")
            if j >= n_pbuf:
                yield ("wait", f"c{j % n_pbuf}", j // n_pbuf - 1)  # the accumulate MMA of block j - n_pbuf has read this P buffer
            yield ("sleep", t_math + T_ST)
            yield ("signal", f"p{j % n_pbuf}")
        yield ("wait", "fin", 0)
        yield ("sleep", epilogue)
    return [issuer(), threads()]


def two_tiles_inplace(n, t_score, t_acc, t_math, share=1.0, prologue=0, epilogue=0):
    """Current forward: two query tiles per CTA, each with two score buffers, one issuer serving the tiles in a fixed order.
    share > 1 stretches a thread phase when both tiles' softmax warps run at once (they share the MUFU of each SMSP)."""
    def issuer():
        yield ("sleep", prologue)
        for b in range(min(2, n)):
            for t in range(2):
                yield ("mma", t_score, [f"s{t}{b}"])
        for j in range(n):
            for t in range(2):
                yield ("wait", f"p{t}{j % 2}", j // 2)
                yield ("mma", t_acc, [f"fin{t}"] if j == n - 1 else [])
                if j + 2 < n:
                    yield ("mma", t_score, [f"s{t}{j % 2}"])

    def threads(t):
        for j in range(n):
            yield ("wait", f"s{t}{j % 2}", j // 2)
            yield ("sleep", T_LD + t_math * share + T_ST)
            yield ("signal", f"p{t}{j % 2}")
        yield ("wait", f"fin{t}", 0)
        yield ("sleep", epilogue)
    return [issuer(), threads(0), threads(1)]


def report(name, procs, n, tensor_per_block):
    s = Sim()
    total = s.run(procs)
    print(f"{name:72s} {total / n:7.0f} cyc/block   tensor busy {100 * s.busy / total:5.1f} %   (tensor floor {tensor_per_block})")


if __name__ == "__main__":
    n = 17
    print("dQ kernel (64-key blocks, 8 compute warps: math ~700):")
    report("  now: Q,dO in TMEM, 2 buffers in place", two_buffer_inplace(n, 512, 256, 700, 3000, 1200), n, 768)
    report("  now, no prologue/epilogue (in-loop)", two_buffer_inplace(200, 512, 256, 700), 200, 768)
    report("  3-stage: Q in TMEM, dO in smem (S 8x32 + dP 8x48), P double buffer", three_stage(200, 640, 256, 700), 200, 896)
    report("  3-stage, persistent (prologue hidden, epilogue 600)", three_stage(n, 640, 256, 700, 2, 2, 300, 600), n, 896)
    report("  3-stage, 16 compute warps (math ~450)", three_stage(200, 640, 256, 450), 200, 896)
    print("dK/dV kernel (64-query blocks):")
    report("  now: K,V in smem (2 x 8x48), 2 buffers in place, math ~800", two_buffer_inplace(200, 768, 512, 800), 200, 1280)
    report("  now incl. prologue 2500 / epilogue 2500 at n=17", two_buffer_inplace(n, 768, 512, 800, 2500, 2500), n, 1280)
    print("forward (two tiles per CTA, 64-key blocks):")
    report("  now: S 8x48, PV 4x64, softmax ~600 alone", two_tiles_inplace(200, 384, 256, 600), 200, 1280)
    report("  now, MUFU shared (x1.6 when overlapping)", two_tiles_inplace(200, 384, 256, 600, 1.6), 200, 1280)
    report("  128-key blocks: S 8x64, PV 8x64, softmax ~1150", two_tiles_inplace(100, 512, 512, 1150), 100, 2048)
