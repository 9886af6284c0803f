"""The reference's "decode without KV cache" benchmark mode on one B200: the loop of
llava/dynamic_eval/bench_test/dynamic_llava_long_text_time_with_no_cache.py:336-349 — `model(total_ids, images, use_cache=False)` with
the id sequence growing by one (teacher) token per call, CUDA-event time per call — on LLaVA-1.5-7B dims, random-init weights,
batch 4 (the script's default), 32-id prompt + one 336x336 image, keep 0.2.  Three arms: output thinning by the predictor's own
scores, by hook-injected Bernoulli(0.5)-like logits (random-init predictors keep nearly everything or nothing), and switched off
(`use_output_text_predictor=False`: the dense re-feed the reference compares against).

    python tools/nocache_bench.py [--calls 192] [--batch 4] [--out gpurun_out/nocache_bench.json]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.sweep import build  # noqa: E402
from dynamic_llava_b200.synthetic import synthetic_inputs  # noqa: E402


def run(model, ids, pix, teacher, n_calls, hook=None):
    model.model.answer_indice = None
    h = model.model.output_text_score_predictor.register_forward_hook(hook) if hook is not None else None
    total = ids.clone()
    evs, lens = [], []
    for t in range(n_calls):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = model(total, images=pix, use_cache=False)
        e1.record()
        evs.append((e0, e1))
        lens.append(out.logits.shape[1])
        total = torch.cat([total, teacher[:, t:t + 1]], dim=1)
    torch.cuda.synchronize()
    if h is not None:
        h.remove()
    ms = [a.elapsed_time(b) for a, b in evs]
    return ms, lens


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--calls", type=int, default=192)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "nocache_bench.json"))
    a = ap.parse_args()
    cfgd, model = build("llava-1.5-7b")
    ids, pix = synthetic_inputs(cfgd, a.batch, 32, seed=1)
    ids, pix = ids.cuda(), pix.cuda()
    g = torch.Generator().manual_seed(3)
    teacher = torch.randint(3, cfgd["vocab_size"], (a.batch, a.calls), generator=g).cuda()

    def bernoulli_hook(mod, args, out):  # distinct scores, about half of them above logit1 = 0
        n = out.shape[1]
        gg = torch.Generator(device="cuda").manual_seed(1000 + n)
        z = torch.zeros_like(out)
        z[..., 0] = torch.randn(out.shape[:-1], generator=gg, device="cuda").to(out.dtype)
        return z
    res = {"workload": f"llava-1.5-7b bf16 random-init, batch {a.batch}, 32-id prompt + 1 image, keep 0.2, {a.calls} no-cache calls",
           "arms": {}}
    sc = model.config.sparse_config
    for name, flag, hook in (("thinning_natural", True, None), ("thinning_bernoulli_0.5", True, bernoulli_hook), ("dense_refeed", False, None)):
        sc["use_output_text_predictor"] = flag
        run(model, ids, pix, teacher, a.calls, hook)  # warm-up over EVERY sequence length: each new GEMM shape costs a host-side heuristic
        # lookup and new allocator blocks the first time it is seen (several ms per call, which would otherwise land on the first arm)
        ms, lens = run(model, ids, pix, teacher, a.calls, hook)
        pick = sorted({0, 1, 2, a.calls // 4, a.calls // 2, a.calls - 1})
        res["arms"][name] = {"total_ms": sum(ms), "ms_per_call_at": {str(t): round(ms[t], 3) for t in pick},
                             "seq_len_at_sparse_layers_at": {str(t): lens[t] for t in pick},
                             "tokens_per_s": a.batch * a.calls / (sum(ms) / 1e3)}
        print(name, json.dumps(res["arms"][name]), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
