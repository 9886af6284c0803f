"""What does each non-GEMM kernel class cost ON THE CRITICAL PATH of the decode step?  The step is one CUDA-graph replay of ~233 PDL-chained
launches, so per-kernel durations (ncu: serialised, cold) do not add up to it.  This tool re-captures the step with one kernel class replaced
by a no-op (results are garbage, timing is what is read) and reports the step time of each variant (config 2: 7B, B=8, keep 0.2, forced 0.5).

    python tools/decode_whatif.py [--out gpurun_out/decode_whatif.json]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.sweep import build  # noqa: E402
from dynamic_llava_b200.synthetic import forced_text_decisions, synthetic_inputs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "decode_whatif.json"))
ap.add_argument("--new-tokens", type=int, default=128)
a = ap.parse_args()
cfgd, model = build("llava-1.5-7b")
B, NN = 8, a.new_tokens
ids, pix = synthetic_inputs(cfgd, B, 32, seed=1)
ids, pix = ids.cuda(), pix.cuda()
fd = forced_text_decisions(NN - 1, B, keep_rate=0.5, seed=3)
kw = dict(images=pix, max_new_tokens=NN, do_sample=False, eos_token_id=None, forced_text_decisions=fd)
eng = model.engine
lib = eng.L
noop = lambda *args: 0


def step_ms():
    for _ in range(2):
        model.generate(ids, **kw)
    ts = []
    for _ in range(3):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        model.generate(ids, max_new_tokens=1, images=pix, do_sample=False, eos_token_id=None)
        e1.record()
        model.generate(ids, **kw)
        e2.record()
        torch.cuda.synchronize()
        ts.append((e1.elapsed_time(e2) - e0.elapsed_time(e1)) / (NN - 1))
    return sorted(ts)[1]


variants = [("full step", []), ("no ragged attention (dl_attn_decode_fused)", ["attn_decode_fused"]),
            ("no split-K reduce + residual + RMSNorm (dl_dec_reduce_residual_norm)", ["dec_reduce_residual_norm"]),
            ("neither", ["attn_decode_fused", "dec_reduce_residual_norm"]),
            ("neither, no text predictor side stream", ["attn_decode_fused", "dec_reduce_residual_norm", "layernorm", "text_decision"]),
            ("GEMMs only (also no embed/rmsnorm/greedy/kv_advance/lm_head reduce)",
             ["attn_decode_fused", "dec_reduce_residual_norm", "layernorm", "text_decision", "embed_rows", "rmsnorm", "greedy_next", "kv_advance",
              "dec_reduce_bf16"])]
res = []
for name, skip in variants:
    lib.version()  # make sure __getattr__ has populated nothing we are about to shadow incorrectly
    saved = {}
    for k in skip:
        getattr(lib, k)  # populate the cached bound call
        saved[k] = lib.__dict__[k]
        lib.__dict__[k] = noop
    for s in list(model._sessions.values()):
        s.graph = None
    model._sessions.clear()
    ms = step_ms()
    for k, v in saved.items():
        lib.__dict__[k] = v
    res.append({"variant": name, "decode_step_ms": round(ms, 4)})
    print(res[-1], flush=True)
os.makedirs(os.path.dirname(a.out), exist_ok=True)
json.dump(res, open(a.out, "w"), indent=1)
