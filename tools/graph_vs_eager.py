"""generate() at full 7B size: the CUDA-graph-replayed decode must produce exactly the tokens of the eager decode (same kernels, no
host gaps between launches in the graph: any cross-launch race shows up as a difference), for batch sizes around the BASELINE ones"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dynamic_llava_b200 import DynamicLlavaConfig, DynamicLlavaLlamaForCausalLM
from dynamic_llava_b200.presets import preset
from dynamic_llava_b200.synthetic import device_state_dict, forced_text_decisions, synthetic_inputs
dev = torch.device("cuda")
name = sys.argv[1] if len(sys.argv) > 1 else "llava-1.5-7b"
cfgd = preset(name)
cfgd["sparse_config"]["use_output_text_predictor"] = True
model = DynamicLlavaLlamaForCausalLM(DynamicLlavaConfig(**cfgd), device=dev)
model.get_vision_tower().load_model()
model.load_state_dict(device_state_dict(cfgd, seed=0, device=dev), strict=True)
bad = 0
for B, T, n_new, quirk in [(1, 32, 48, False), (2, 32, 48, False), (2, 32, 48, True), (3, 17, 48, True), (5, 64, 48, False), (7, 32, 48, True),
                           (8, 32, 96, False), (11, 8, 48, True), (16, 32, 48, False)]:
    model.config.reference_batched_quirk = quirk
    ids, pix = synthetic_inputs(cfgd, B, T, seed=1)
    fd = forced_text_decisions(n_new - 1, B, keep_rate=0.5, seed=3).to(dev)
    kw = dict(images=pix.to(dev), max_new_tokens=n_new, do_sample=False, eos_token_id=None, forced_text_decisions=fd)
    eager = model.generate(ids.to(dev), use_cuda_graph=False, **kw)
    runs = [model.generate(ids.to(dev), use_cuda_graph=True, **kw) for _ in range(3)]
    same = [bool(torch.equal(r, eager)) for r in runs]
    first = [int((r != eager).any(0).int().argmax()) if not s else -1 for r, s in zip(runs, same)]
    print(json.dumps(dict(preset=name, B=B, T=T, n_new=n_new, quirk=quirk, graph_equals_eager=same, first_diff_step=first)), flush=True)
    bad += sum(not s for s in same)
print("ok" if not bad else f"{bad} mismatching runs")
sys.exit(1 if bad else 0)
