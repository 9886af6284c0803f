"""K1 (dl_select_compact) at the sizes DESIGN.md quotes, heuristic grid (R = 0), buffers rotated beyond the 126 MB L2:
config 2 (7B B=8 keep 0.2, 19 MB), config 5 (7B B=16, keep 0.1/0.2/0.5/1.0), config 4 per GPU (13B B=4 P=2048, 130 MB), 13B B=32 (1 GB).

    python tools/k1_points.py [--out gpurun_out/k1_points.json]
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
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "k1_points.json"))
a = ap.parse_args()
L = lib()
P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
ST = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
peak = float(pk.get("hbm_gbs", 6455.6))


def timeit(fn, n_rot, iters=40, warm=5):
    """-> (us per launch with eager launches, us per launch inside one CUDA-graph replay).  A ctypes launch costs ~10+ us of host
    time, so the eager loop is HOST-bound for kernels shorter than that; the graph replay has no host in it."""
    for i in range(warm):
        fn(i % n_rot)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % n_rot)
    e1.record()
    torch.cuda.synchronize()
    eager = e0.elapsed_time(e1) / iters * 1e3
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
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return eager, e0.elapsed_time(e1) / iters * 1e3


rows = []
for name, B, T, H, rate in [("cfg2 7B B=8 keep 0.2", 8, 32, 4096, 0.2), ("cfg5 7B B=16 keep 0.1", 16, 32, 4096, 0.1),
                            ("cfg5 7B B=16 keep 0.2", 16, 32, 4096, 0.2), ("cfg5 7B B=16 keep 0.5", 16, 32, 4096, 0.5),
                            ("cfg5 7B B=16 keep 1.0", 16, 32, 4096, 1.0), ("cfg4 13B B=4 P=2048 keep 0.2", 4, 1473, 5120, 0.2),
                            ("13B B=32 P=2048 keep 0.2", 32, 1473, 5120, 0.2)]:
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
    us_eager, us = timeit(lambda i: L.select_compact(P(logits), None, P(hid[i]), P(out[i]), P(keep), P(pos), P(cu), B, Pn, H, s0, n_img, k, 0, ST()), n_rot)
    rows.append({"case": name, "MB": round(nbytes / 1e6, 1), "us": round(us, 2), "us_eager_launches": round(us_eager, 2), "GBps": round(nbytes / us / 1e3), "frac_of_hbm_peak": round(nbytes / us / 1e3 / peak, 3)})
    print(rows[-1], flush=True)
    del hid, out
os.makedirs(os.path.dirname(a.out), exist_ok=True)
json.dump({"peak_gbs": peak, "rows": rows}, open(a.out, "w"), indent=1)
