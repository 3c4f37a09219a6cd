"""Run the GPU parity checks (tests/gpu_checks.py) and write gpurun_out/diag.json.

All checks run in one child process (one torch import); when a check kills the process (a trapped kernel poisons the
CUDA context, a watchdog timeout...) the remaining checks continue in a fresh child.
Usage: python tools/gpu_diag.py [name ...]
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PER_CHECK_TIMEOUT = 180


def main():
    import ast
    src = open(os.path.join(ROOT, "tests", "gpu_checks.py")).read()
    tree = ast.parse(src)
    all_names = []
    for n in tree.body:
        if isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") == "ALL":
            all_names = [k.value for k in n.value.keys]
    names = sys.argv[1:] or all_names
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    results, todo = [], list(names)
    while todo:
        t0 = time.time()
        p = subprocess.Popen([sys.executable, "-m", "tests.gpu_checks"] + todo, cwd=ROOT, stdout=subprocess.PIPE,
                             stderr=subprocess.STDOUT, text=True)
        done_here, tail = [], []
        try:
            out, _ = p.communicate(timeout=PER_CHECK_TIMEOUT * max(1, len(todo)) // 2 + 120)
        except subprocess.TimeoutExpired:
            p.kill()
            out, _ = p.communicate()
        for line in out.splitlines():
            if line.startswith("RESULT "):
                r = json.loads(line[7:])
                results.append(r)
                done_here.append(r["name"])
                print(json.dumps(r)[:700], flush=True)
            else:
                tail.append(line)
        remaining = [n for n in todo if n not in done_here]
        if remaining and (p.returncode != 0 or not done_here):
            if p.returncode != 3 or not done_here or results[-1].get("ok", False):
                # the child died inside remaining[0] without reporting
                bad = remaining.pop(0)
                r = {"name": bad, "ok": False, "rc": p.returncode, "error": "\n".join(tail)[-1500:], "sec": round(time.time() - t0, 1)}
                results.append(r)
                print(json.dumps(r)[:900], flush=True)
        todo = remaining
        json.dump(results, open(os.path.join(ROOT, "gpurun_out", "diag.json"), "w"), indent=1)
    print("FAILED:", [r["name"] for r in results if not r.get("ok")])


if __name__ == "__main__":
    main()
