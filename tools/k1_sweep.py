"""K1 (dl_select_compact): CUDA-graph-timed sweep of the CTAs-per-sample count R at the sizes VERDICT names (19 / 38 / 84 / 130 MB),
buffers rotated beyond the 126 MB L2.  R = 0 is the launcher's own heuristic.

    python tools/k1_sweep.py [--out gpurun_out/k1_sweep.json]
"""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynamic_llava_b200._lib import lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "k1_sweep.json"))
a = ap.parse_args()
L = lib()
P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
ST = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
peak = 6455.6
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", peak))


def graph_us(fn, n_rot, iters=40):
    for i in range(3):
        fn(i % n_rot)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for i in range(iters):
                fn(i % n_rot)
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters * 1e3)
    return best


rows = []
for name, B, T, H, rate in [("cfg2 7B B=8 keep 0.2 (19 MB)", 8, 32, 4096, 0.2), ("cfg5 7B B=16 keep 0.2 (38 MB)", 16, 32, 4096, 0.2),
                            ("cfg5 7B B=16 keep 0.5 (84 MB)", 16, 32, 4096, 0.5), ("cfg4 13B B=4 P=2048 keep 0.2 (130 MB)", 4, 1473, 5120, 0.2)]:
    n_img, s0 = 576, 5
    Pn = T - 1 + n_img
    k = int(n_img * rate)
    Pp = Pn - n_img + k
    nbytes = 2 * B * Pp * H * 2 + B * n_img * 4 + B * Pp * 4
    n_rot = max(2, int(300e6 // (B * Pn * H * 2)) + 1)
    hid = [torch.randn(B, Pn, H, device="cuda", dtype=torch.bfloat16) for _ in range(n_rot)]
    out = [torch.empty(B, Pp, H, device="cuda", dtype=torch.bfloat16) for _ in range(n_rot)]
    logits = torch.randn(B, n_img, 2, device="cuda", dtype=torch.bfloat16)
    keep = torch.empty(B, k, dtype=torch.int64, device="cuda")
    pos = torch.empty(B, Pp, dtype=torch.int32, device="cuda")
    cu = torch.empty(B + 1, dtype=torch.int32, device="cuda")
    res = {}
    for R in (0, 4, 8, 12, 16, 19, 24, 28, 37, 48, 56, 74, 111, 148):
        if R > Pp:
            continue
        us = graph_us(lambda i: L.select_compact(P(logits), None, P(hid[i]), P(out[i]), P(keep), P(pos), P(cu), B, Pn, H, s0, n_img, k, R, ST()), n_rot)
        res[R] = {"us": round(us, 2), "frac": round(nbytes / us / 1e3 / peak, 3)}
    rows.append({"case": name, "MB": round(nbytes / 1e6, 1), "by_R": res})
    print(name, {r: (v["us"], v["frac"]) for r, v in res.items()}, flush=True)
    del hid, out
os.makedirs(os.path.dirname(a.out), exist_ok=True)
json.dump({"peak_gbs": peak, "rows": rows}, open(a.out, "w"), indent=1)
