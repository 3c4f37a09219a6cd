"""Run every GPU parity check in its own subprocess with a timeout; write gpurun_out/diag.json.

A kernel that traps (e.g. the mbarrier watchdog) poisons only its own process, so one bad kernel
does not hide the state of the others.  Usage: python tools/gpu_diag.py [name ...]
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from tests import gpu_checks  # noqa: only for the list of names (imports torch once here)
    names = sys.argv[1:] or list(gpu_checks.ALL)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    results = []
    for n in names:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, "-m", "tests.gpu_checks", n], cwd=ROOT, capture_output=True, text=True,
                               timeout=240)
            out = p.stdout + p.stderr
            line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
            if p.returncode == 0 and line:
                r = json.loads(line[-1][7:])
            else:
                r = {"name": n, "ok": False, "rc": p.returncode, "tail": out[-1500:]}
        except subprocess.TimeoutExpired as e:
            r = {"name": n, "ok": False, "rc": "timeout", "tail": ((e.stdout or b"")[-800:]).decode("utf-8", "replace")
                 if isinstance(e.stdout, bytes) else str(e.stdout)[-800:]}
        r["wall"] = round(time.time() - t0, 1)
        results.append(r)
        print(json.dumps(r)[:600], flush=True)
        json.dump(results, open(os.path.join(ROOT, "gpurun_out", "diag.json"), "w"), indent=1)
    bad = [r["name"] for r in results if not r.get("ok")]
    print("FAILED:", bad)


if __name__ == "__main__":
    main()
