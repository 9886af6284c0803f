"""every kernel of the prefill is deterministic, so its logits must be BITWISE identical whether the launches overlap (programmatic
serialization, the default) or run fully serialised (DLB200_NO_PDL=15).  Each prefill is queued behind ~20 ms of dummy work so that all of
its ~460 launches execute back to back on the device (a ctypes launch otherwise costs more host time than the small kernels run).

    python tools/pdl_bitwise.py out.pt            # run once per setting, then
    python tools/pdl_bitwise.py --compare a.pt b.pt
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
if sys.argv[1] == "--compare":
    a, b = torch.load(sys.argv[2]), torch.load(sys.argv[3])
    bad = [k for k in a if not torch.equal(a[k], b[k])]
    for k in bad:
        print("DIFF", k, (a[k].float() - b[k].float()).abs().max().item())
    print(f"{len(a) - len(bad)}/{len(a)} configurations bitwise identical")
    sys.exit(1 if bad else 0)
from dynamic_llava_b200 import DynamicLlavaConfig, DynamicLlavaLlamaForCausalLM
from dynamic_llava_b200.presets import preset
from dynamic_llava_b200.synthetic import device_state_dict, synthetic_inputs
dev = torch.device("cuda")
cfgd = preset(os.environ.get("PDL_BITWISE_PRESET", "llava-1.5-7b"))
cfgd["sparse_config"]["use_output_text_predictor"] = False
model = DynamicLlavaLlamaForCausalLM(DynamicLlavaConfig(**cfgd), device=dev)
model.get_vision_tower().load_model()
model.load_state_dict(device_state_dict(cfgd, seed=0, device=dev), strict=True)
big = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
res = {}
quick = os.environ.get("PDL_BITWISE_QUICK") == "1"
for B in ((2, 3, 8) if quick else (1, 2, 3, 4, 5, 6, 8, 11)):
    for T in ((32,) if quick else (8, 32, 100)):
        for rate in ((0.2, 0.5) if quick else (0.1, 0.2, 0.5, 1.0)):
            model.config.sparse_config["vision_keep_rate"] = rate
            ids, pix = synthetic_inputs(cfgd, B, T, seed=1)
            ids, pix = ids.to(dev), pix.to(dev)
            model(ids, images=pix)  # warm (allocations, tensor maps)
            torch.cuda.synchronize()
            for _ in range(16):
                big @ big  # ~20 ms of queued work: the prefill's launches pile up behind it
            out = model(ids, images=pix)
            torch.cuda.synchronize()
            res[f"B={B} T={T} keep={rate}"] = out.logits[:, -1].cpu().clone()
            del out
torch.save(res, sys.argv[1])
print("saved", len(res))
