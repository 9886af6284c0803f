"""Per-item clock trace of CTA 0 of the persistent v4 prefill attention kernel (debug build with -DDL_T4_TRACE).
    python tools/trace_attn4.py --build ; python tools/trace_attn4.py   (B, NH, LQ from the environment)"""
import ctypes, json, math, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "dynamic_llava_b200", "csrc")
SO = os.path.join(ROOT, "tools", "_trace", "libt4trace.so")
if "--build" in sys.argv:
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
                    "--expt-relaxed-constexpr", "-DDL_T4_TRACE", "-shared", "-cudart", "static", "-o", SO] +
                   [os.path.join(CSRC, f) for f in ("attn_prefill_tc4.cu", "abi.cu", "skinny_gemm.cu")], check=True)
    print("built", SO)
    sys.exit(0)
import torch
dll = ctypes.CDLL(SO)
B, nh, Lq, hd = int(os.environ.get("B", 8)), int(os.environ.get("NH", 32)), int(os.environ.get("LQ", 2048)), 128
H, cap, N = nh * hd, Lq + 8, B * Lq
P = lambda t: ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
qkv = torch.randn(N, 3 * H, device="cuda", dtype=torch.bfloat16)
K = torch.randn(B, nh, cap, hd, device="cuda", dtype=torch.bfloat16)
V = torch.randn(B, nh, cap, hd, device="cuda", dtype=torch.bfloat16)
o = torch.empty(N, H, device="cuda", dtype=torch.bfloat16)
cu = torch.arange(0, (B + 1) * Lq, Lq, device="cuda", dtype=torch.int32)
past = torch.zeros(B, device="cuda", dtype=torch.int32)
tr = torch.zeros(16 * 64, device="cuda", dtype=torch.int64)
assert dll.dl_debug_t4_set_trace(P(tr)) == 0
fn = dll.dl_attn_prefill_tc4
fn.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
               ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_void_p]
for _ in range(3):
    assert fn(P(qkv), 3 * H, P(K), P(V), P(o), H, P(cu), P(past), N, B, nh, cap, Lq, 1.0 / math.sqrt(hd), int(os.environ.get("FLAGS", 1)), st) == 0
torch.cuda.synchronize()
t = tr.cpu().view(16, 64)
t0 = int(t[0, 0])
res = []
for g in (0, 1):
    n = int((t[g * 4] > 0).sum())
    print(f"group {g}: {n} items on CTA 0;  [halves] enter -> firstS (wait) -> lastP (cycles/half) -> epilogue end (epilogue) | gap to next item's first S")
    for i in range(n):
        e, f, l, d = (int(t[g * 4 + u, i]) - t0 for u in range(4))
        nh_ = int(t[8 + g, i])
        nxt = int(t[g * 4 + 1, i + 1]) - t0 if i + 1 < n else None
        res.append({"g": g, "item": i, "halves": nh_, "enter": e, "first_s": f, "last_p": l, "epi_end": d})
        print(f"  item {i:2d} [{nh_:2d}] {e:8d} -> {f:8d} ({f - e:5d}) -> {l:8d} ({(l - f) / max(nh_, 1):6.0f}) -> {d:8d} ({d - l:5d}) | {'' if nxt is None else nxt - d}")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "trace_attn_tc4.json"), "w"), indent=1)
