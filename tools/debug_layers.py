"""per-layer input error of the B200 prefill vs the device-resident oracle (same image mask), full-size 7B, given batch sizes"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dynamic_llava_b200 import DynamicLlavaConfig, DynamicLlavaLlamaForCausalLM
from dynamic_llava_b200.presets import preset
from dynamic_llava_b200.synthetic import device_state_dict, synthetic_inputs
from oracle import dynllava_oracle as O
dev = torch.device("cuda")
cfgd = preset("llava-1.5-7b")
cfgd["sparse_config"]["use_output_text_predictor"] = False
sd = device_state_dict(cfgd, seed=0, device=dev)
model = DynamicLlavaLlamaForCausalLM(DynamicLlavaConfig(**cfgd), device=dev)
model.get_vision_tower().load_model()
model.load_state_dict(sd, strict=True)
for B in [int(x) for x in (sys.argv[1:] or ["2", "4"])]:
    ids, pix = synthetic_inputs(cfgd, B, 32, seed=1)
    ids, pix = ids.to(dev), pix.to(dev)
    ro = {}
    lo, _ = O.model_forward(sd, cfgd, ids, images=pix, record=ro)
    rg = {}
    out = model(ids, images=pix, forced_keep_index=ro["keep_index"], record=rg)
    print(f"B={B}: logits err {(out.logits.float() - lo.float()).abs().max().item():.4f} scale {lo.float().abs().max().item():.3f}")
    for l, (a, b) in enumerate(zip(rg["layer_inputs"], ro["layer_inputs"])):
        d = (a.float() - b.float()).abs()
        per_b = d.flatten(1).max(1).values.tolist()
        print(f"  layer {l:2d} input: max err {d.max().item():.4f} (scale {b.float().abs().max().item():.2f}) per sample {[round(x, 4) for x in per_b]} rows {a.shape[1]}")
    del out
