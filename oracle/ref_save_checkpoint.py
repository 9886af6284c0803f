"""Write a checkpoint directory with the REFERENCE's own `save_pretrained` (toy dims, numpy-seeded weights, bf16):
    python oracle/ref_save_checkpoint.py OUT_DIR
Used by tests/test_host_logic_cpu.py (build container only: needs /root/reference) to prove that a directory produced by the
reference classes loads through this package's `load_pretrained_model`.  Runs in its own process because the reference's `llava`
package and this repo's `llava` shim share a name."""
import importlib.util
import os
import sys

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("gen_golden", os.path.join(ROOT, "oracle", "gen_golden.py"))
gg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(gg)  # puts /root/reference first on sys.path and applies the transformers-5 shim

import torch  # noqa: E402
from dynamic_llava_b200.presets import preset  # noqa: E402
from dynamic_llava_b200.synthetic import numpy_state_dict  # noqa: E402


def main(out):
    cfg = preset("toy")
    cfg["sparse_config"]["use_output_text_predictor"] = True
    sd = numpy_state_dict(cfg, seed=0)
    os.makedirs(out, exist_ok=True)
    model = gg.build_reference(cfg, sd, torch.bfloat16, out)  # writes OUT/clip (the tower dir the config points at)
    ck = os.path.join(out, "dynamic-llava-toy")
    model.config.save_pretrained(ck)  # config.json as the reference's DynamicLlavaConfig serialises itself
    # the weights as `save_pretrained` would shard them (transformers 5.x cannot run it on this class: `_tied_weights_keys` is the
    # 4.x list form, dynamic_modeling_llama.py:2598): two safetensors shards of the reference model's own state_dict
    from safetensors.torch import save_file
    items = [(k, v.contiguous()) for k, v in model.state_dict().items()]
    half = len(items) // 2
    save_file(dict(items[:half]), os.path.join(ck, "model-00001-of-00002.safetensors"), metadata={"format": "pt"})
    save_file(dict(items[half:]), os.path.join(ck, "model-00002-of-00002.safetensors"), metadata={"format": "pt"})
    print("saved", sorted(os.listdir(ck)))


if __name__ == "__main__":
    main(sys.argv[1])
