"""One `generate()` of the bench workload (7B, B=8, keep 0.2, forced 0.5 decisions) with the decode step launched eagerly —
the command profiled by ncu for profiles/*launches*.  Not a benchmark: numbers printed under a profiler are never reported.

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynamic_llava_b200 import DynamicLlavaConfig, DynamicLlavaLlamaForCausalLM  # noqa: E402
from dynamic_llava_b200.presets import preset  # noqa: E402
from dynamic_llava_b200.synthetic import device_state_dict, forced_text_decisions, synthetic_inputs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--preset", default="llava-1.5-7b")
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--new-tokens", type=int, default=6)
ap.add_argument("--graph", action="store_true")
a = ap.parse_args()
cfgd = preset(a.preset)
cfgd["sparse_config"]["use_output_text_predictor"] = True
model = DynamicLlavaLlamaForCausalLM(DynamicLlavaConfig(**cfgd), device="cuda")
model.get_vision_tower().load_model()
# cheap deterministic fill (keeps the profiled launch list free of ~450 RNG kernels): one randn pool, sliced
pool = torch.randn(1 << 24, device="cuda", dtype=torch.bfloat16) * 0.02
with torch.no_grad():
    for n, p in model.named_parameters():
        if "norm" in n and n.endswith("weight"):
            p.fill_(1.0)
        elif n.endswith("bias"):
            p.zero_()
        else:
            flat = p.view(-1)
            for o in range(0, flat.numel(), pool.numel()):
                m = min(pool.numel(), flat.numel() - o)
                flat[o:o + m] = pool[:m]
ids, pix = synthetic_inputs(cfgd, a.batch, 32, seed=1)
fd = forced_text_decisions(max(a.new_tokens - 1, 1), a.batch, 0.5, seed=3)
torch.cuda.synchronize()
out = model.generate(ids.cuda(), images=pix.cuda(), max_new_tokens=a.new_tokens, do_sample=False, eos_token_id=None,
                     forced_text_decisions=fd, use_cuda_graph=a.graph)
torch.cuda.synchronize()
print("generated", tuple(out.shape), "launches", model.engine.launches)
