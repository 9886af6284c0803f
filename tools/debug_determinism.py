"""run the 13B config-4 prefill several times and report the first layer input / row that differs between runs"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynamic_llava_b200 import DynamicLlavaConfig, DynamicLlavaLlamaForCausalLM
from dynamic_llava_b200.presets import preset
from dynamic_llava_b200.synthetic import device_state_dict, synthetic_inputs

name = sys.argv[1] if len(sys.argv) > 1 else "llava-1.5-13b"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
T = int(sys.argv[3]) if len(sys.argv) > 3 else 2048 - 576 + 1
cfgd = preset(name)
cfgd["sparse_config"]["use_output_text_predictor"] = False
model = DynamicLlavaLlamaForCausalLM(DynamicLlavaConfig(**cfgd), device="cuda")
model.get_vision_tower().load_model()
model.load_state_dict(device_state_dict(cfgd, seed=0, device="cuda"), strict=True)
ids, pix = synthetic_inputs(cfgd, B, T, seed=1)
ids, pix = ids.cuda(), pix.cuda()
for mode in ("sk", "tiles"):
    if mode == "tiles":
        model.engine.gemm_ws = None
    runs = []
    for r in range(4):
        rec = {}
        feats = model.encode_images(pix)
        out = model(ids, images=pix, record=rec)
        runs.append((feats.clone(), [x.clone() for x in rec["layer_inputs"]], out.logits.clone(), rec["keep_index"].clone(), rec["vision_logits"].clone()))
    torch.cuda.synchronize()
    for r in range(1, 4):
        msgs = []
        if not torch.equal(runs[0][0], runs[r][0]):
            d = (runs[0][0] != runs[r][0]).nonzero()
            msgs.append(f"clip feats differ: {d.shape[0]} elems, first {d[0].tolist()}")
        if not torch.equal(runs[0][4], runs[r][4]):
            msgs.append("vision logits differ")
        if not torch.equal(runs[0][3], runs[r][3]):
            msgs.append("keep_index differ")
        for l, (a, b) in enumerate(zip(runs[0][1], runs[r][1])):
            if not torch.equal(a, b):
                d = (a != b).nonzero()
                rows = sorted(set((int(x[0]), int(x[1])) for x in d[:2000]))
                msgs.append(f"layer {l} input differs: {d.shape[0]} elems; (sample,row) first {rows[:6]} cols {sorted(set(int(x[2]) for x in d[:200]))[:12]}")
                break
        if not torch.equal(runs[0][2], runs[r][2]):
            msgs.append("logits differ")
        print(mode, "run", r, "vs 0:", msgs if msgs else "identical", flush=True)
