"""Run every GPU test FUNCTION in its own subprocess with a hard timeout (a hung kernel must not take the box down or
hide the other results).  Used during bring-up on the GPU box; the driver's own run is plain `pytest -m gpu`.

    python tests/run_gpu_isolated.py [--jobs 4] [--timeout 150] [-k substr]
"""
import argparse
import json
import os
import re
import subprocess
import sys
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=4)
    ap.add_argument("--timeout", type=int, default=150)
    ap.add_argument("-k", default="")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "isolated.json"))
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    col = subprocess.run([sys.executable, "-m", "pytest", "tests", "-m", "gpu", "--collect-only", "-q"], cwd=ROOT, capture_output=True, text=True)
    funcs = []
    for ln in col.stdout.splitlines():
        m = re.match(r"(tests/\S+?::\w+)", ln)
        if m and m.group(1) not in funcs and a.k in m.group(1):
            funcs.append(m.group(1))

    def run(node):
        t0 = time.time()
        try:
            r = subprocess.run(["timeout", "-s", "KILL", str(a.timeout), sys.executable, "-m", "pytest", node, "-m", "gpu", "-q", "-x",
                                "--no-header", "-p", "no:cacheprovider"], cwd=ROOT, capture_output=True, text=True)
            out = (r.stdout + r.stderr)
            status = "pass" if r.returncode == 0 else ("TIMEOUT" if r.returncode in (137, -9) else "FAIL")
        except Exception as e:  # noqa
            out, status = str(e), "ERROR"
        return {"node": node, "status": status, "secs": round(time.time() - t0, 1), "tail": out[-2500:]}

    with ThreadPoolExecutor(a.jobs) as ex:
        res = list(ex.map(run, funcs))
    json.dump(res, open(a.out, "w"), indent=1)
    for r in res:
        print(f"{r['status']:8s} {r['secs']:6.1f}s  {r['node']}")
    bad = [r for r in res if r["status"] != "pass"]
    for r in bad:
        print("=" * 100)
        print(r["node"], r["status"])
        print(r["tail"][-1800:])
    print(f"\n{len(res) - len(bad)}/{len(res)} test functions passed")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
