"""isolated timings (CUDA-graph replay, weights rotated beyond L2) of the decode-step ops: fused chain (dec_gemm.cu) vs round-1 kernels"""
import ctypes, json, math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynamic_llava_b200._lib import lib
L = lib()
DEV, BF = "cuda", torch.bfloat16
P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
ST = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

def graph_us(fn, n_rot, iters=24, warm=3):
    for i in range(warm): fn(i % n_rot)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for i in range(iters): fn(i % n_rot)
    torch.cuda.current_stream().wait_stream(side)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

M, H, I, V, nh, hd = 8, 4096, 11008, 32000, 32, 128
NR = 6
g = torch.Generator(device=DEV).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=DEV, dtype=BF, generator=g)
x = rn(M, H); lnw = torch.ones(H, device=DEV, dtype=BF)
ssq = (x.float() ** 2).sum(1, keepdim=True).repeat(1, 256) / 256
wqkv = [rn(3 * H, H) * 0.02 for _ in range(NR)]; wo = [rn(H, H) * 0.02 for _ in range(NR * 2)]
wgu = [rn(2 * I, H) * 0.02 for _ in range(NR)]; wd = [rn(H, I) * 0.02 for _ in range(NR)]
wl = [rn(V, H) * 0.02 for _ in range(2)]
cap = 800
K_ = torch.zeros(M, nh, cap, hd, device=DEV, dtype=BF); V_ = torch.zeros_like(K_)
q = torch.empty(M, H, device=DEV, dtype=BF); attn = rn(M, H); act = rn(M, I); xres = rn(M, H)
kvl = torch.full((M,), 210, dtype=torch.int32, device=DEV); pos = torch.full((M,), 700, dtype=torch.int32, device=DEV)
from oracle import dynllava_oracle as O
cos, sin = O.rope_tables(hd, 4096, 10000.0, BF); cos, sin = cos.to(DEV), sin.to(DEV)
ssq_out = torch.empty(M, 256, device=DEV, dtype=torch.float32)
logits = torch.empty(M, V, device=DEV, dtype=BF)
part = torch.empty(8 * M * 3 * H, device=DEV, dtype=torch.float32)
xn = rn(M, H)
out = {}
out["v2 qkv (norm-on-load, cluster reduce, rope, write)"] = graph_us(lambda i: L.dec_gemm_qkv(P(x), P(ssq), 256, P(lnw), 1e-5, P(wqkv[i]), P(q), P(K_), P(V_), P(kvl), P(pos), P(cos), P(sin), M, nh, hd, cap, H, ST()), NR)
s = L.skinny_gemm_splits(3 * H, H, 296)
out[f"v1 qkv skinny splits={s}"] = graph_us(lambda i: L.skinny_gemm(P(xn), P(wqkv[i]), P(part), M, 3 * H, H, s, -1, ST()), NR)
out["v2 o_proj (cluster reduce + residual + ssq)"] = graph_us(lambda i: L.dec_gemm_residual(P(attn), P(wo[i]), P(xres), P(ssq_out), M, H, H, ST()), NR * 2)
s2 = L.skinny_gemm_splits(H, H, 296)
def v1_o(i):
    L.skinny_gemm(P(attn), P(wo[i]), P(part), M, H, H, s2, -1, ST())
    L.dec_reduce_residual_norm(P(part), P(xres), P(lnw), P(xn), M, H, s2, 1e-5, ST())
out[f"v1 o_proj skinny splits={s2} + reduce_residual_norm"] = graph_us(v1_o, NR * 2)
out["v2 gate|up (norm-on-load + swiglu)"] = graph_us(lambda i: L.dec_gemm_swiglu(P(x), P(ssq), 256, P(lnw), 1e-5, P(wgu[i]), P(act), M, I, H, ST()), NR)
out["v1 gate|up skinny swiglu"] = graph_us(lambda i: L.skinny_gemm_swiglu(P(xn), P(wgu[i]), P(act), M, I, H, ST()), NR)
out["v2 down (cluster reduce + residual + ssq)"] = graph_us(lambda i: L.dec_gemm_residual(P(act), P(wd[i]), P(xres), P(ssq_out), M, H, I, ST()), NR)
s3 = L.skinny_gemm_splits(H, I, 296)
def v1_d(i):
    L.skinny_gemm(P(act), P(wd[i]), P(part), M, H, I, s3, -1, ST())
    L.dec_reduce_residual_norm(P(part), P(xres), P(lnw), P(xn), M, H, s3, 1e-5, ST())
out[f"v1 down skinny splits={s3} + reduce_residual_norm"] = graph_us(v1_d, NR)
out["v2 lm_head logits"] = graph_us(lambda i: L.dec_gemm_logits(P(x), P(ssq), 256, P(lnw), 1e-5, P(wl[i]), P(logits), M, V, H, ST()), 2)
def v1_l(i):
    L.skinny_gemm(P(xn), P(wl[i]), P(part), M, V, H, 1, -1, ST())
    L.dec_reduce_bf16(P(part), P(logits), M, V, 1, ST())
out["v1 lm_head skinny + reduce_bf16"] = graph_us(v1_l, 2)
out["v2 attn inplace len 210"] = graph_us(lambda i: L.attn_decode_inplace(P(q), H, P(K_), P(V_), P(kvl), None, P(attn), None, M, nh, hd, cap, 1, 1 / math.sqrt(hd), 0, 128, ST()), 1)
out["v1 attn fused len 210"] = graph_us(lambda i: L.attn_decode_fused(P(part), 3, P(pos), None, P(cos), P(sin), P(K_), P(V_), P(kvl), P(attn), None, M, nh, hd, cap, 1, 1 / math.sqrt(hd), 0, None, 0, 128, ST()), 1)
for k, v in out.items():
    print(f"{v:8.2f} us  {k}")
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_decode_ops.json"), "w"), indent=1)
