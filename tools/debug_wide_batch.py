"""row i of a B-wide prefill vs the same sample alone, same image mask: max |logit diff| (bf16 noise is ~0.3 on a scale of ~6 at 7B)"""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynamic_llava_b200 import DynamicLlavaConfig, DynamicLlavaLlamaForCausalLM
from dynamic_llava_b200.presets import preset
from dynamic_llava_b200.synthetic import device_state_dict, synthetic_inputs
dev = torch.device("cuda")
cfgd = preset("llava-1.5-7b")
cfgd["sparse_config"]["use_output_text_predictor"] = False
model = DynamicLlavaLlamaForCausalLM(DynamicLlavaConfig(**cfgd), device=dev)
model.get_vision_tower().load_model()
model.load_state_dict(device_state_dict(cfgd, seed=0, device=dev), strict=True)
for B in (2, 8, 16, 32):
    ids, pix = synthetic_inputs(cfgd, B, 32, seed=1)
    ids, pix = ids.to(dev), pix.to(dev)
    rec = {}
    out = model(ids, images=pix, record=rec)
    wide = out.logits.float().clone()
    keep = rec["keep_index"].clone()
    vl = rec["vision_logits"].float().clone()
    del out
    for i in (0, B - 1):
        r1 = {}
        o1 = model(ids[i:i+1], images=pix[i:i+1], forced_keep_index=keep[i:i+1], record=r1)
        d = (o1.logits.float()[0] - wide[i]).abs().max().item()
        dv = (r1["vision_logits"].float()[0] - vl[i]).abs().max().item()
        r2 = {}
        o2 = model(ids[i:i+1], images=pix[i:i+1], record=r2)
        same_keep = len(set(r2["keep_index"][0].tolist()) & set(keep[i].tolist()))
        print(json.dumps(dict(B=B, row=i, max_logit_diff_same_mask=round(d, 4), scale=round(wide[i].abs().max().item(), 3), vision_logit_diff=round(dv, 4),
                              natural_keep_overlap=same_keep, k=keep.shape[1], argmax_equal=bool(o1.logits[0, -1].argmax() == wide[i, -1].argmax()))), flush=True)
        del o1, o2
