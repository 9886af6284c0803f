"""BASELINE.json configs 4 and 5 on one B200 (parity-test-size workloads, not the bench line):

  config 5: keep-rate sweep {0.1, 0.2, 0.5, 1.0} on LLaVA-1.5-7B, batch 16: tokens/s and achieved GB/s of the
            select+compact kernel vs the effective token count B*P'
  config 4: LLaVA-1.5-13B bf16, keep 0.2, P = 2048 prefill (T = 1473 ids), per-GPU batch 4 (= 32 over 8 GPUs), + 1 token

    python tools/sweep.py [--only 5|4] [--out gpurun_out/sweep.json]
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynamic_llava_b200 import DynamicLlavaConfig, DynamicLlavaLlamaForCausalLM  # noqa: E402
from dynamic_llava_b200.presets import preset  # noqa: E402
from dynamic_llava_b200.synthetic import device_state_dict, forced_text_decisions, synthetic_inputs  # noqa: E402


def build(name):
    cfgd = preset(name)
    cfgd["sparse_config"]["use_output_text_predictor"] = True
    model = DynamicLlavaLlamaForCausalLM(DynamicLlavaConfig(**cfgd), device="cuda")
    model.get_vision_tower().load_model()
    model.load_state_dict(device_state_dict(cfgd, seed=0, device="cuda"), strict=True)
    torch.cuda.empty_cache()
    return cfgd, model


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def config5(out, batch=16, new_tokens=64):
    cfgd, model = build("llava-1.5-7b")
    ids, pix = synthetic_inputs(cfgd, batch, 32, seed=1)
    ids, pix = ids.cuda(), pix.cuda()
    rows = []
    for rate in (0.1, 0.2, 0.5, 1.0):
        model.config.sparse_config["vision_keep_rate"] = rate  # live dict, like the reference's bench script
        model.config.sparse_config["use_output_text_predictor"] = False
        k = int(576 * rate)
        Pp = 607 - 576 + k
        kw = dict(images=pix, do_sample=False, eos_token_id=None)
        t_pre = timed(lambda: model.generate(ids, max_new_tokens=1, **kw))
        t_all = timed(lambda: model.generate(ids, max_new_tokens=new_tokens, **kw), reps=3, warm=2)
        rows.append({"vision_keep_rate": rate, "k": k, "P_prime": Pp, "effective_tokens": batch * Pp,
                     "clip_prefill_ms": t_pre, "prefill_tokens_per_s": batch * 607 / (t_pre / 1e3),
                     "total_ms": t_all, "generated_tokens_per_s": batch * new_tokens / (t_all / 1e3),
                     "decode_ms_per_step": (t_all - t_pre) / (new_tokens - 1)})
        print("CFG5", rows[-1], flush=True)
    out["config5_keep_rate_sweep"] = {"model": "llava-1.5-7b", "batch": batch, "new_tokens": new_tokens, "rows": rows}
    del model
    torch.cuda.empty_cache()


def config4(out, batch=4):
    cfgd, model = build("llava-1.5-13b")
    T = 2048 - 576 + 1
    ids, pix = synthetic_inputs(cfgd, batch, T, seed=1)
    ids, pix = ids.cuda(), pix.cuda()
    fd = forced_text_decisions(1, batch, 0.5, seed=3)
    kw = dict(images=pix, do_sample=False, eos_token_id=None)
    t_pre = timed(lambda: model.generate(ids, max_new_tokens=1, **kw), reps=3, warm=1)
    fd = forced_text_decisions(16, batch, 0.5, seed=3)
    t_2 = timed(lambda: model.generate(ids, max_new_tokens=17, forced_text_decisions=fd, **kw), reps=3, warm=1)
    sess = next(iter(model._sessions.values()))
    lens = sess.arena.kv_len.cpu()
    out["config4_13b_prefill2048"] = {"model": "llava-1.5-13b", "batch_per_gpu": batch, "P": 2048, "P_prime": 2048 - 576 + 115,
                                      "clip_prefill_ms": t_pre, "prefill_tokens_per_s": batch * 2048 / (t_pre / 1e3),
                                      "decode_ms_per_step": (t_2 - t_pre) / 16, "kv_len_layer0": lens[0].tolist(), "kv_len_last": lens[-1].tolist()}
    print("CFG4", out["config4_13b_prefill2048"], flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="5,4")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.json"))
    a = ap.parse_args()
    res = {"gpu": torch.cuda.get_device_name(0)}
    if "5" in a.only:
        config5(res)
    if "4" in a.only:
        config4(res)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)
