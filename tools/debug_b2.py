"""stress the kernels on the shapes a 7B prefill sees at B = 2 / 3 (M = 292 / 438 rows in layers >= 2): every result against torch"""
import ctypes, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from dynamic_llava_b200._lib import lib
L = lib()
DEV, BF = "cuda", torch.bfloat16
P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
ST = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
torch.manual_seed(0)
ws = torch.zeros(L.linear_workspace_bytes(), device=DEV, dtype=torch.uint8)
H, I, nh, hd = 4096, 11008, 32, 128
wqkv = (torch.randn(3 * H, H, device=DEV) * 0.02).to(BF)
wo = (torch.randn(H, H, device=DEV) * 0.02).to(BF)
wgu = (torch.randn(2 * I, H, device=DEV) * 0.02).to(BF)
wd = (torch.randn(H, I, device=DEV) * 0.02).to(BF)
for B in (2, 3, 4):
    M = B * 146
    bad = {"qkv": 0, "o": 0, "gu": 0, "down": 0, "attn": 0}
    worst = dict(bad)
    for it in range(30):
        x = torch.randn(M, H, device=DEV).to(BF)
        qkv = torch.empty(M, 3 * H, device=DEV, dtype=BF)
        L.linear(P(x), P(wqkv), None, None, P(qkv), M, 3 * H, H, 0, P(ws), ST())
        # attention on the projection (no rope: the kernel under test is the attention itself)
        K = qkv[:, H:2 * H].view(B, 146, nh, hd).permute(0, 2, 1, 3).contiguous()
        V = qkv[:, 2 * H:].view(B, 146, nh, hd).permute(0, 2, 1, 3).contiguous()
        att = torch.empty(M, H, device=DEV, dtype=BF)
        cu = torch.arange(0, (B + 1) * 146, 146, device=DEV, dtype=torch.int32)
        past = torch.zeros(B, device=DEV, dtype=torch.int32)
        L.attn_prefill_tc4(P(qkv), 3 * H, P(K), P(V), P(att), H, P(cu), P(past), M, B, nh, 146, 146, 1.0 / math.sqrt(hd), 0, ST())
        res = torch.randn(M, H, device=DEV).to(BF)
        o = res.clone()
        L.linear(P(att), P(wo), None, P(o), P(o), M, H, H, 0, P(ws), ST())
        act = torch.empty(M, I, device=DEV, dtype=BF)
        L.linear_swiglu(P(o), P(wgu), P(act), M, I, H, P(ws), ST())
        d = o.clone()
        L.linear(P(act), P(wd), None, P(d), P(d), M, H, I, 0, P(ws), ST())
        torch.cuda.synchronize()
        r_qkv = (x.float() @ wqkv.float().T)
        q4 = qkv[:, :H].view(B, 146, nh, hd).permute(0, 2, 1, 3).float()
        r_att = F.scaled_dot_product_attention(q4, K.float(), V.float(), is_causal=True).permute(0, 2, 1, 3).reshape(M, H)
        r_o = res.float() + (att.float() @ wo.float().T).to(BF).float()
        gu = (o.float() @ wgu.float().T).to(BF)
        r_act = (F.silu(gu[:, :I]).to(BF).float() * gu[:, I:].float())
        r_d = o.float() + (act.float() @ wd.float().T).to(BF).float()
        for name, got, ref in (("qkv", qkv, r_qkv), ("attn", att, r_att), ("o", o, r_o), ("gu", act, r_act), ("down", d, r_d)):
            e = (got.float() - ref).abs().max().item() / max(1.0, ref.abs().max().item())
            worst[name] = max(worst[name], e)
            if e > 2e-2:
                bad[name] += 1
    print(f"B={B} M={M}: bad launches of 30 {bad}  worst rel err { {k: round(v, 4) for k, v in worst.items()} }", flush=True)
