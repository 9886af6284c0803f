"""how much of the eager CLIP + prefill time is launch gaps?  Time the same prefill (a) normally and (b) queued behind ~25 ms of dummy
work, so that all of its launches are already in the queue when the device gets to them (what a CUDA graph of the prefill would give)"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dynamic_llava_b200 import DynamicLlavaConfig, DynamicLlavaLlamaForCausalLM
from dynamic_llava_b200.presets import preset
from dynamic_llava_b200.synthetic import device_state_dict, synthetic_inputs
dev = torch.device("cuda")
cfgd = preset("llava-1.5-7b")
cfgd["sparse_config"]["use_output_text_predictor"] = True
model = DynamicLlavaLlamaForCausalLM(DynamicLlavaConfig(**cfgd), device=dev)
model.get_vision_tower().load_model()
model.load_state_dict(device_state_dict(cfgd, seed=0, device=dev), strict=True)
B, T = 8, 32
ids, pix = synthetic_inputs(cfgd, B, T, seed=1)
ids_d, pix_d = ids.to(dev), pix.to(dev)
big = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)


def once(queued):
    torch.cuda.synchronize()
    if queued:
        for _ in range(20):
            big @ big
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = model.generate(ids_d, images=pix_d, max_new_tokens=1, do_sample=False, eos_token_id=None)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


for _ in range(3):
    once(False)
res = {"eager_ms": sorted(once(False) for _ in range(7))[3], "queued_ms": sorted(once(True) for _ in range(7))[3]}
print(json.dumps(res))
