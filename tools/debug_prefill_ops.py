"""bisect inside Engine._prefill_layer at full 7B size: every op of every layer against torch on the op's ACTUAL inputs"""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from dynamic_llava_b200 import DynamicLlavaConfig, DynamicLlavaLlamaForCausalLM, engine as E
from dynamic_llava_b200.presets import preset
from dynamic_llava_b200.synthetic import device_state_dict, synthetic_inputs
dev = torch.device("cuda")
cfgd = preset("llava-1.5-7b")
cfgd["sparse_config"]["use_output_text_predictor"] = False
sd = device_state_dict(cfgd, seed=0, device=dev)
model = DynamicLlavaLlamaForCausalLM(DynamicLlavaConfig(**cfgd), device=dev)
model.get_vision_tower().load_model()
model.load_state_dict(sd, strict=True)
_p, _stream = E._p, E._stream
BF = torch.bfloat16


def rel(got, ref):
    return (got.float() - ref.float()).abs().max().item() / max(1.0, ref.float().abs().max().item())


def checked_layer(self, l, x, pos, cu, sess, B, Lq):
    c = self.cfg
    W = self.packed["layers"][l]
    H, nh = c.hidden_size, c.num_attention_heads
    hd = H // nh
    N = x.shape[0]
    K, V, cap = sess.arena.layer(l)
    kvl = sess.arena.kv_len[l]
    errs = {}
    x0 = x.clone()
    xn = self.rmsnorm(x, W["ln1"], c.rms_norm_eps)
    v = x0.float().pow(2).mean(-1, keepdim=True)
    ref_xn = (W["ln1"].float() * (x0.float() * torch.rsqrt(v + c.rms_norm_eps)).to(BF).float()).to(BF)
    errs["ln1"] = rel(xn, ref_xn)
    qkv = self.linear(xn, W["wqkv"])
    errs["qkv"] = rel(qkv, xn.float() @ W["wqkv"].float().T)
    qkv_pre = qkv.clone()
    self.L.rope_append(_p(qkv), _p(pos), _p(cu), _p(kvl), None, _p(self.packed["cos"]), _p(self.packed["sin"]), _p(K), _p(V),
                       None, None, N, B, nh, hd, cap, _stream())
    torch.cuda.synchronize()
    kv0 = kvl.clone()
    # arena rows of this call: [b, h, past:past+Lq]
    cos, sin = self.packed["cos"][pos.long()].float(), self.packed["sin"][pos.long()].float()  # [N, hd]
    def rot(t):  # [N, nh, hd]
        t = t.float()
        t1, t2 = t[..., : hd // 2], t[..., hd // 2:]
        return (t * cos[:, None, :]).to(BF).float() + (torch.cat([-t2, t1], -1) * sin[:, None, :]).to(BF).float()
    q_ref = rot(qkv_pre[:, :H].view(N, nh, hd)).to(BF)
    k_ref = rot(qkv_pre[:, H:2 * H].view(N, nh, hd)).to(BF)
    v_ref = qkv_pre[:, 2 * H:].view(N, nh, hd)
    errs["rope_q"] = rel(qkv[:, :H].view(N, nh, hd), q_ref)
    k_got = torch.stack([K[b, :, int(kv0[b]): int(kv0[b]) + Lq] for b in range(B)], 0)  # [B, nh, Lq, hd]
    v_got = torch.stack([V[b, :, int(kv0[b]): int(kv0[b]) + Lq] for b in range(B)], 0)
    errs["arena_k"] = rel(k_got, k_ref.view(B, Lq, nh, hd).permute(0, 2, 1, 3))
    errs["arena_v"] = rel(v_got, v_ref.view(B, Lq, nh, hd).permute(0, 2, 1, 3))
    attn = torch.empty(N, H, device=x.device, dtype=x.dtype)
    self.L.attn_prefill_tc4(_p(qkv), 3 * H, _p(K), _p(V), _p(attn), H, _p(cu), _p(kvl), N, B, nh, cap, Lq, 1.0 / math.sqrt(hd), self.attn_flags, _stream())
    q4 = qkv[:, :H].view(B, Lq, nh, hd).permute(0, 2, 1, 3).float()
    r_att = F.scaled_dot_product_attention(q4, k_got.float(), v_got.float(), is_causal=True).permute(0, 2, 1, 3).reshape(N, H)
    errs["attn"] = rel(attn, r_att)
    x1 = x.clone()
    self.linear(attn, W["wo"], residual=x, out=x)
    errs["o_proj"] = rel(x, x1.float() + (attn.float() @ W["wo"].float().T).to(BF).float())
    x2 = x.clone()
    xn = self.rmsnorm(x, W["ln2"], c.rms_norm_eps, out=xn)
    v = x2.float().pow(2).mean(-1, keepdim=True)
    errs["ln2"] = rel(xn, (W["ln2"].float() * (x2.float() * torch.rsqrt(v + c.rms_norm_eps)).to(BF).float()).to(BF))
    act = self.linear_swiglu(xn, W["wgu"])
    I = W["wgu"].shape[0] // 2
    gu = (xn.float() @ W["wgu"].float().T).to(BF)
    errs["swiglu"] = rel(act, F.silu(gu[:, :I]).to(BF).float() * gu[:, I:].float())
    self.linear(act, W["wd"], residual=x, out=x)
    errs["down"] = rel(x, x2.float() + (act.float() @ W["wd"].float().T).to(BF).float())
    bad = {k: round(v, 4) for k, v in errs.items() if v > 1.5e-2}
    print(f"  layer {l:2d} rows {N:5d}: " + ("ok (max %.4f)" % max(errs.values()) if not bad else f"BAD {bad}"), flush=True)
    return x


E.Engine._prefill_layer = checked_layer
for B in [int(a) for a in (sys.argv[1:] or ["2"])]:
    ids, pix = synthetic_inputs(cfgd, B, 32, seed=1)
    print(f"B={B}")
    model(ids.to(dev), images=pix.to(dev))
