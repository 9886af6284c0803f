"""Per-phase clock trace of ONE CTA of the v3 prefill attention kernel (debug build with -DDL_T3_TRACE).

    python tools/trace_attn.py --build     # here (cross-compile): tools/_trace/libt3trace.so
    python tools/trace_attn.py             # on the GPU box: prints per-half cycle deltas of the heaviest causal tile

Stamps (clock64 of the SM) per softmax group g and 64-key half j: [0] S(j) complete seen, [1] S in registers (tcgen05.ld done),
[2] P(j) stored + p_full arrive.  Not part of the product library."""
import ctypes
import json
import math
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "dynamic_llava_b200", "csrc")
SO = os.path.join(ROOT, "tools", "_trace", "libt3trace.so")


def build():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
           "--expt-relaxed-constexpr", "-DDL_T3_TRACE", "-shared", "-cudart", "static", "-o", SO] + [os.path.join(CSRC, f) for f in ("attn_prefill_tc3.cu", "abi.cu", "skinny_gemm.cu")]
    subprocess.run(cmd, check=True)
    print("built", SO)


def main():
    import torch
    dll = ctypes.CDLL(SO)
    B, nh, Lq, hd = int(os.environ.get("B", 1)), int(os.environ.get("NH", 1)), int(os.environ.get("LQ", 2048)), 128
    H = nh * hd
    cap = Lq + 8
    N = B * Lq
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    qkv = torch.randn(N, 3 * H, device="cuda", dtype=torch.bfloat16)
    K = torch.randn(B, nh, cap, hd, device="cuda", dtype=torch.bfloat16)
    V = torch.randn(B, nh, cap, hd, device="cuda", dtype=torch.bfloat16)
    o = torch.empty(N, H, device="cuda", dtype=torch.bfloat16)
    cu = torch.arange(0, (B + 1) * Lq, Lq, device="cuda", dtype=torch.int32)
    past = torch.zeros(B, device="cuda", dtype=torch.int32)
    res = {}
    for flags in (0, 1, 2):
        tr = torch.zeros(16 * 64, device="cuda", dtype=torch.int64)
        assert dll.dl_debug_t3_set_trace(P(tr)) == 0
        fn = dll.dl_attn_prefill_tc3
        fn.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                       ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int,
                       ctypes.c_void_p]
        for _ in range(3):
            rc = fn(P(qkv), 3 * H, P(K), P(V), P(o), H, P(cu), P(past), N, B, nh, cap, Lq, 1.0 / math.sqrt(hd), flags, st)
            assert rc == 0
        torch.cuda.synchronize()
        t = tr.cpu().view(16, 64)
        t0 = int(t[t > 0].min())
        n_h = int((t[4] > 0).sum())  # group B sees the most halves
        print(f"--- flags={flags}: heaviest tile, group B: {n_h} halves; cycles relative to the first stamp")
        rows = []
        for g in (0, 1):
            n = int((t[g * 4] > 0).sum())
            for j in range(n):
                s, l, p = (int(t[g * 4 + u, j]) - t0 for u in range(3))
                rows.append({"g": g, "j": j, "s_seen": s, "ld_done": l, "p_done": p})
            if n > 4:
                per = (int(t[g * 4 + 2, n - 2]) - int(t[g * 4 + 2, 2])) / (n - 4)
                wait = sum(int(t[g * 4 + 0, j]) - int(t[g * 4 + 2, j - 1]) for j in range(3, n - 1)) / (n - 4)
                ld = sum(int(t[g * 4 + 1, j]) - int(t[g * 4 + 0, j]) for j in range(3, n - 1)) / (n - 4)
                sm = sum(int(t[g * 4 + 2, j]) - int(t[g * 4 + 1, j]) for j in range(3, n - 1)) / (n - 4)
                print(f"group {g}: {n} halves, steady-state cycles per half {per:.0f} = wait-for-S {wait:.0f} + tcgen05.ld {ld:.0f} + max/exp/st/arrive {sm:.0f}")
                res[f"flags{flags}_g{g}"] = {"halves": n, "cycles_per_half": per, "wait_s": wait, "ld": ld, "softmax": sm}
        print("MMA thread prologue issue stamps (S_A0,S_B0,S_A1,S_B1):", [int(x) - t0 for x in t[12, :4]])
        for j in range(min(8, n_h)):
            print(f"  mma j={j}: PV_A issued {int(t[8, j]) - t0}, S_A(j+2) issued {int(t[9, j]) - t0}, PV_B issued {int(t[10, j]) - t0}, S_B(j+2) issued {int(t[11, j]) - t0}")
        res[f"flags{flags}_mma"] = [[int(t[u, j]) - t0 for u in (8, 9, 10, 11)] for j in range(n_h)]
        end = [int(t[g * 4 + 3, 0]) - t0 for g in (0, 1)]
        print("epilogue done at", end)
        res[f"flags{flags}_rows"] = rows
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "trace_attn_tc3.json"), "w"), indent=1)


if __name__ == "__main__":
    if "--build" in sys.argv:
        build()
    else:
        main()
