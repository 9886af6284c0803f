"""Per-kernel roofline microbenchmarks on the B200 (CUDA events, >=3 warm-ups, buffers rotated so that the working set
exceeds the 126 MB L2).  Writes gpurun_out/microbench.json; the summary lives in profiles/.

    python tools/microbench.py [--only k1,attn_prefill,attn_decode,skinny]
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynamic_llava_b200._lib import lib  # noqa: E402

DEV = "cuda"
BF = torch.bfloat16
L = None


def P(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def ST():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def timeit(fn, n_rot, iters=20, warm=3):
    """fn(i) launches on rotation slot i % n_rot; returns median-free mean ms per call over `iters` back-to-back calls."""
    for i in range(warm):
        fn(i % n_rot)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % n_rot)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    d = json.load(open(p)) if os.path.exists(p) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
    return d


def bench_k1(out):
    pk = peaks()["hbm_gbs"]
    rows = []
    for name, B, T, H, rate in [("7B B=8 keep0.2", 8, 32, 4096, 0.2), ("7B B=16 keep0.1", 16, 32, 4096, 0.1), ("7B B=16 keep0.2", 16, 32, 4096, 0.2),
                                ("7B B=16 keep0.5", 16, 32, 4096, 0.5), ("7B B=16 keep1.0", 16, 32, 4096, 1.0),
                                ("13B cfg4 B=4 P=2048", 4, 1473, 5120, 0.2), ("13B B=32 P=2048", 32, 1473, 5120, 0.2)]:
        n_img, s0 = 576, 5
        Pn = T - 1 + n_img
        k = int(n_img * rate)
        Pp = Pn - n_img + k
        bytes_alg = 2 * B * Pp * H * 2 + B * n_img * 4 + B * Pp * 4
        n_rot = max(2, int(300e6 // (B * Pn * H * 2)) + 1)
        hid = [torch.randn(B, Pn, H, device=DEV, dtype=BF) for _ in range(n_rot)]
        logits = torch.randn(B, n_img, 2, device=DEV, dtype=BF)
        outb = [torch.empty(B, Pp, H, device=DEV, dtype=BF) for _ in range(n_rot)]
        keep = torch.empty(B, k, dtype=torch.int64, device=DEV)
        pos = torch.empty(B, Pp, dtype=torch.int32, device=DEV)
        cu = torch.empty(B + 1, dtype=torch.int32, device=DEV)
        best = None
        for R in (0, 16, 37, 74, 148, 296):
            if R > Pp:
                continue
            ms = timeit(lambda i: L.select_compact(P(logits), None, P(hid[i]), P(outb[i]), P(keep), P(pos), P(cu), B, Pn, H, s0, n_img, k, R, ST()), n_rot)
            gbs = bytes_alg / ms / 1e6
            if best is None or gbs > best["gbs"]:
                best = {"R": R, "ms": ms, "gbs": gbs}
        # torch reference of the same op (argsort + sort + gather + cat), for context
        def torch_ref(i):
            sc = torch.log_softmax(logits, -1)[:, :, 0]
            ki = torch.sort(torch.argsort(sc, dim=1, descending=True)[:, :k], dim=1).values
            g = hid[i][:, s0:s0 + n_img].gather(1, ki[..., None].expand(B, k, H))
            return torch.cat([hid[i][:, :s0], g, hid[i][:, s0 + n_img:]], 1)
        ms_t = timeit(torch_ref, n_rot)
        rows.append({"case": name, "alg_MB": bytes_alg / 1e6, "best_R": best["R"], "us": best["ms"] * 1e3, "GBps": best["gbs"],
                     "frac_of_measured_hbm": best["gbs"] / pk, "torch_eager_us": ms_t * 1e3})
        print("K1", rows[-1], flush=True)
        del hid, outb
        torch.cuda.empty_cache()
    out["k1_select_compact"] = rows


def bench_attn_prefill(out):
    pk = peaks()
    rows = []
    import torch.nn.functional as F
    for name, B, nh, Lq in [("7B L=607 B=8", 8, 32, 607), ("7B L=146 B=8", 8, 32, 146), ("13B L=2048 B=4", 4, 40, 2048), ("13B L=1587 B=4", 4, 40, 1587),
                            ("7B L=2048 B=8", 8, 32, 2048)]:
        hd = 128
        H = nh * hd
        cap = Lq + 8
        N = B * Lq
        qkv = torch.randn(N, 3 * H, device=DEV, dtype=BF)
        K = torch.randn(B, nh, cap, hd, device=DEV, dtype=BF)
        V = torch.randn(B, nh, cap, hd, device=DEV, dtype=BF)
        o = torch.empty(N, H, device=DEV, dtype=BF)
        cu = torch.arange(0, (B + 1) * Lq, Lq, device=DEV, dtype=torch.int32)
        past = torch.zeros(B, device=DEV, dtype=torch.int32)
        flop = 2.0 * Lq * Lq * H * B  # causal: 4*L*L*H/2 per sample
        sc = 1.0 / math.sqrt(hd)
        q4 = qkv[:, :H].view(B, Lq, nh, hd).transpose(1, 2)
        k4, v4 = K[:, :, :Lq], V[:, :, :Lq]
        o_ref = F.scaled_dot_product_attention(q4, k4, v4, is_causal=True).transpose(1, 2).reshape(N, H)  # library result for the max-diff columns
        t3 = {}
        for fl in (0, 1):
            for nm, fn3 in (("tc4", L.attn_prefill_tc4),):
                ms3 = timeit(lambda i: fn3(P(qkv), 3 * H, P(K), P(V), P(o), H, P(cu), P(past), N, B, nh, cap, Lq, sc, fl, ST()), 1)
                t3[f"{nm}_f{fl}_us"] = ms3 * 1e3
                t3[f"{nm}_f{fl}_TFLOPs"] = flop / ms3 / 1e9
                t3[f"{nm}_f{fl}_frac_of_sustained_bf16"] = flop / ms3 / 1e9 / pk["bf16_tflops_sustained"]
                t3[f"{nm}_f{fl}_vs_sdpa_maxdiff"] = (o.float() - o_ref.float()).abs().max().item()
        ms_sdpa = timeit(lambda i: F.scaled_dot_product_attention(q4, k4, v4, is_causal=True), 1)
        rows.append({"case": name, "GFLOP": flop / 1e9, "torch_sdpa_us": ms_sdpa * 1e3, "torch_sdpa_TFLOPs": flop / ms_sdpa / 1e9, **t3})
        print("ATTN_PREFILL", rows[-1], flush=True)
    out["attn_prefill"] = rows
    rows = []
    for name, B, nh, N in [("CLIP B=8 577x16h", 8, 16, 577), ("predictor B=8 576x8h", 8, 8, 576)]:
        hd, C = 64, nh * 64
        qkv = torch.randn(B * N, 3 * C, device=DEV, dtype=BF)
        o = torch.empty(B * N, C, device=DEV, dtype=BF)
        cu = torch.arange(0, (B + 1) * N, N, device=DEV, dtype=torch.int32)
        flop = 4.0 * N * N * C * B
        sc = 1.0 / math.sqrt(hd)
        q4 = qkv.view(B, N, 3, nh, hd).permute(2, 0, 3, 1, 4)
        o1 = F.scaled_dot_product_attention(q4[0], q4[1], q4[2]).transpose(1, 2).reshape(B * N, C)
        t3 = {}
        for fl in (0, 1):
            for nm, fn3 in (("tc4", L.attn_packed_tc4),):
                ms3 = timeit(lambda i: fn3(P(qkv), P(o), P(cu), B * N, B, nh, hd, N, sc, fl, ST()), 1)
                t3[f"{nm}_f{fl}_us"] = ms3 * 1e3
                t3[f"{nm}_f{fl}_TFLOPs"] = flop / ms3 / 1e9
                t3[f"{nm}_f{fl}_vs_sdpa_maxdiff"] = (o.float() - o1.float()).abs().max().item()
        ms_sdpa = timeit(lambda i: F.scaled_dot_product_attention(q4[0], q4[1], q4[2]), 1)
        rows.append({"case": name, "GFLOP": flop / 1e9, "torch_sdpa_us": ms_sdpa * 1e3, **t3})
        print("ATTN_PACKED", rows[-1], flush=True)
    out["attn_packed_hd64"] = rows


def bench_attn_decode(out):
    pk = peaks()["hbm_gbs"]
    rows = []
    for name, B, nh, ln, ns in [("cfg2 hi len=210", 8, 32, 210, 1), ("cfg2 lo len=735", 8, 32, 735, 1), ("cfg2 lo len=735 split2", 8, 32, 735, 2),
                                ("B=64 len=400", 64, 32, 400, 1), ("B=8 len=2048 split4", 8, 32, 2048, 4), ("B=1 len=2048 split8", 1, 32, 2048, 8)]:
        hd = 128
        cap = ln + 8
        per = B * nh * cap * hd * 2 * 2
        n_rot = max(2, int(300e6 // per) + 1)
        Ks = [torch.randn(B, nh, cap, hd, device=DEV, dtype=BF) for _ in range(n_rot)]
        Vs = [torch.randn(B, nh, cap, hd, device=DEV, dtype=BF) for _ in range(n_rot)]
        q = torch.randn(B, nh * hd, device=DEV, dtype=BF)
        ks, vs = torch.randn(B, nh, hd, device=DEV, dtype=BF), torch.randn(B, nh, hd, device=DEV, dtype=BF)
        o = torch.empty(B, nh * hd, device=DEV, dtype=BF)
        kvl = torch.full((B,), ln, device=DEV, dtype=torch.int32)
        ws = torch.empty(max(L.attn_decode_workspace_bytes(B, nh, ns) // 4, 1), device=DEV, dtype=torch.float32)
        ms = timeit(lambda i: L.attn_decode(P(q), nh * hd, P(Ks[i]), P(Vs[i]), P(ks), P(vs), P(kvl), P(o), P(ws), B, nh, hd, cap, ns, 0.088, 0, ST()), n_rot)
        b = B * nh * ln * hd * 2 * 2
        rows.append({"case": name, "alg_MB": b / 1e6, "us": ms * 1e3, "GBps": b / ms / 1e6, "frac_of_measured_hbm": b / ms / 1e6 / pk})
        print("ATTN_DECODE", rows[-1], flush=True)
        del Ks, Vs
        torch.cuda.empty_cache()
    out["attn_decode"] = rows


def bench_skinny(out):
    pk = peaks()["hbm_gbs"]
    rows = []
    for name, M, N, K in [("qkv 7B", 8, 12288, 4096), ("o 7B", 8, 4096, 4096), ("gate_up 7B", 8, 22016, 4096), ("down 7B", 8, 4096, 11008),
                          ("lm_head", 8, 32000, 4096), ("qkv 13B", 4, 15360, 5120), ("gate_up 7B M=16", 16, 22016, 4096)]:
        n_rot = max(2, int(400e6 // (N * K * 2)) + 1)
        Ws = [torch.randn(N, K, device=DEV, dtype=BF) for _ in range(n_rot)]
        x = torch.randn(M, K, device=DEV, dtype=BF)
        best = None
        allv = {}
        KB = K // 64
        for s in (1, 2, 3, 4, 6, 8):
            kbps = -(-KB // s)
            if (s - 1) * kbps >= KB or kbps < 4:
                continue
            part = torch.empty(s, M, N, device=DEV, dtype=torch.float32)
            ms = timeit(lambda i: L.skinny_gemm(P(x), P(Ws[i]), P(part), M, N, K, s, 0, ST()), n_rot)
            gbs = N * K * 2 / ms / 1e6
            allv[s] = round(ms * 1e3, 2)
            if best is None or gbs > best["GBps"]:
                best = {"splits": s, "us": ms * 1e3, "GBps": gbs}
        best["all_splits_us"] = allv
        best["heuristic_splits"] = L.skinny_gemm_splits(N, K, 296)
        import torch.nn.functional as F
        ms_c = timeit(lambda i: F.linear(x, Ws[i]), n_rot)
        rows.append({"case": name, "MB": N * K * 2 / 1e6, **best, "frac_of_measured_hbm": best["GBps"] / pk, "cublas_us": ms_c * 1e3,
                     "cublas_GBps": N * K * 2 / ms_c / 1e6})
        print("SKINNY", rows[-1], flush=True)
        del Ws
        torch.cuda.empty_cache()
    out["skinny_gemm"] = rows


def main():
    global L
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="k1,attn_prefill,attn_decode,skinny")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "microbench.json"))
    a = ap.parse_args()
    L = lib()
    out = {"peaks": peaks(), "gpu": torch.cuda.get_device_name(0)}
    todo = a.only.split(",")
    if "k1" in todo:
        bench_k1(out)
    if "attn_prefill" in todo:
        bench_attn_prefill(out)
    if "attn_decode" in todo:
        bench_attn_decode(out)
    if "skinny" in todo:
        bench_skinny(out)
    # dense GEMMs: tools/bench_linear.py
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(out, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
