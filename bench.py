#!/usr/bin/env python
"""bench.py — BASELINE.json metric: prefill+decode tokens/sec @ LLaVA-1.5-7B, vision_keep_rate=0.2.

    python bench.py --gpus 1 --steps 5 --warmup 3                 # our arm, config 2 (B=8/GPU, 128 new tokens)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1   # the reference algorithm (CPU oracle port) on host cores

One "step" = one full pass of the hot path over one batch: CLIP encode -> projector -> splice -> sparsified prefill ->
`new_tokens` greedy decode steps with forced seeded keep/drop decisions.  tokens/s counts
GENERATED tokens: value = n_gpus * B * new_tokens * steps / time (weak scaling: B per GPU is fixed = config 3's layout).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "prefill+decode tokens/sec @ LLaVA-1.5-7B, vision_keep_rate=0.2"
UNIT = "tokens/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--preset", default="llava-1.5-7b")
    ap.add_argument("--batch", type=int, default=8, help="samples per GPU")
    ap.add_argument("--prompt-len", type=int, default=32)
    ap.add_argument("--new-tokens", type=int, default=128)
    ap.add_argument("--keep-rate", type=float, default=0.2)
    ap.add_argument("--out-keep", type=float, default=0.5, help="forced Bernoulli rate for output-token KV decisions; <0 disables")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the decode step eagerly (for ncu launch lists)")
    ap.add_argument("--cpu-new-tokens", type=int, default=32,
                    help="greedy tokens per CPU-baseline step (B=1): CLIP + prefill + 31 decode steps, ~5-15 s of host work per step depending on "
                         "the core count, so that `--impl reference --steps K --warmup W` still ends within a few minutes")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference only: 'cuda' times the same oracle port (the reference's PyTorch op sequence: eager torch, cuBLAS, "
                         "SDPA, torch.cat KV cache) on the GPU over the FULL workload — the like-for-like comparator of SURVEY.md §8d; the "
                         "default (and what the driver runs) is the host-core baseline")
    return ap.parse_args()


def kernel_rooflines(dev):
    """untimed extras: live rooflines of the two other kernels the north star names — K2 (prefill attention, tensor pipe) on the
    L=2048 stress shape and on this workload's own shapes, K1 (select + gather/compact, HBM) on the 13B B=32 shape and on this
    workload's 19 MB.  CUDA events around one CUDA-graph replay of back-to-back launches (K2: 50 after 10 warm-ups, K1: 20 after 3); K1
    rotates buffers beyond the 126 MB L2."""
    import ctypes
    import math
    from dynamic_llava_b200._lib import lib
    L = lib()
    P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
    st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    tf_peak = float(pk.get("bf16_tflops_sustained", 1400.0))
    hbm_peak = float(pk.get("hbm_gbs", 6650.0))

    def timeit(fn, n_rot=1, iters=20, warm=3):
        """ms per launch inside ONE CUDA-graph replay of `iters` launches (CUDA events around the replay).  Eager ctypes launches cost
        ~13 us of host time each (more with tensor-map encoding), which hid every kernel shorter than that
        (profiles/r1_k1_warp_select_experiment.md); a replay has no host in the loop."""
        for i in range(warm):
            fn(i % n_rot)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                for i in range(iters):
                    fn(i % n_rot)
        torch.cuda.current_stream().wait_stream(side)
        g.replay()  # warm replay
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    out = []
    for name, B, nh, Lq in (("K2 dl_attn_prefill_tc4 L=2048 B=8 32 heads (stress)", 8, 32, 2048), ("K2 dl_attn_prefill_tc4 L=607 B=8 (layers 0-1)", 8, 32, 607),
                            ("K2 dl_attn_prefill_tc4 L=146 B=8 (layers 2-31)", 8, 32, 146)):
        hd, H, cap, N = 128, nh * 128, Lq + 8, B * Lq
        qkv = torch.randn(N, 3 * H, device=dev, dtype=torch.bfloat16)
        K = torch.randn(B, nh, cap, hd, device=dev, dtype=torch.bfloat16)
        V = torch.randn(B, nh, cap, hd, device=dev, dtype=torch.bfloat16)
        o = torch.empty(N, H, device=dev, dtype=torch.bfloat16)
        cu = torch.arange(0, (B + 1) * Lq, Lq, device=dev, dtype=torch.int32)
        past = torch.zeros(B, device=dev, dtype=torch.int32)
        ms = timeit(lambda i: L.attn_prefill_tc4(P(qkv), 3 * H, P(K), P(V), P(o), H, P(cu), P(past), N, B, nh, cap, Lq, 1.0 / math.sqrt(hd), 0, st()),
                    iters=50, warm=10)
        flop = 2.0 * Lq * Lq * H * B
        out.append({"kernel": name, "bound": "tensor", "achieved": flop / ms / 1e9, "peak": tf_peak, "unit": "TFLOP/s",
                    "frac": flop / ms / 1e9 / tf_peak, "us": ms * 1e3, "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained"})
        del qkv, K, V, o
    for name, B, T, H in (("K1 dl_select_compact 13B B=32 P=2048 keep 0.2", 32, 1473, 5120), ("K1 dl_select_compact 7B B=8 P=607 keep 0.2 (this workload)", 8, 32, 4096)):
        n_img, s0, k = 576, 5, 115
        Pn = T - 1 + n_img
        Pp = Pn - n_img + k
        n_rot = max(2, int(300e6 // (B * Pn * H * 2)) + 1)
        hid = [torch.randn(B, Pn, H, device=dev, dtype=torch.bfloat16) for _ in range(n_rot)]
        outb = [torch.empty(B, Pp, H, device=dev, dtype=torch.bfloat16) for _ in range(n_rot)]
        logits = torch.randn(B, n_img, 2, device=dev, dtype=torch.bfloat16)
        keep = torch.empty(B, k, dtype=torch.int64, device=dev)
        pos = torch.empty(B, Pp, dtype=torch.int32, device=dev)
        cu = torch.empty(B + 1, dtype=torch.int32, device=dev)
        ms = timeit(lambda i: L.select_compact(P(logits), None, P(hid[i]), P(outb[i]), P(keep), P(pos), P(cu), B, Pn, H, s0, n_img, k, 0, st()), n_rot)
        nbytes = 2 * B * Pp * H * 2 + B * n_img * 4 + B * Pp * 4
        out.append({"kernel": name, "bound": "hbm", "achieved": nbytes / ms / 1e6, "peak": hbm_peak, "unit": "GB/s", "frac": nbytes / ms / 1e6 / hbm_peak,
                    "us": ms * 1e3, "algorithmic_bytes": nbytes, "peak_source": "MEASURED_PEAKS.json hbm_gbs"})
        del hid, outb
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------------------------
def run_b200(a):
    import torch.distributed as dist
    from dynamic_llava_b200 import DynamicLlavaConfig, DynamicLlavaLlamaForCausalLM
    from dynamic_llava_b200.presets import preset
    from dynamic_llava_b200.sharding import gather_token_ids, max_over_ranks
    from dynamic_llava_b200.synthetic import device_state_dict, forced_text_decisions, synthetic_inputs

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly one JSON line: keep NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION) off it
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    cfgd = preset(a.preset)
    cfgd["sparse_config"]["vision_keep_rate"] = a.keep_rate
    cfgd["sparse_config"]["use_output_text_predictor"] = True
    cfg = DynamicLlavaConfig(**cfgd)
    model = DynamicLlavaLlamaForCausalLM(cfg, device=dev)
    model.get_vision_tower().load_model()
    sd = device_state_dict(cfgd, seed=0, device=dev)  # same weights on every replica
    model.load_state_dict(sd, strict=True)
    del sd
    torch.cuda.empty_cache()
    eng = model.engine
    eng._check_packed()

    B, T, NN = a.batch, a.prompt_len, a.new_tokens
    ids_h, pix_h = synthetic_inputs(cfgd, B, T, seed=1 + rank)  # each rank owns its own chunk of the global batch
    fd = forced_text_decisions(max(NN - 1, 1), B, keep_rate=a.out_keep, seed=3 + rank) if a.out_keep >= 0 else None
    ids_pin, pix_pin = ids_h.pin_memory(), pix_h.pin_memory()
    ids_d, pix_d = ids_h.to(dev), pix_h.to(dev)
    gen_kw = dict(max_new_tokens=NN, do_sample=False, eos_token_id=None, forced_text_decisions=fd, use_cuda_graph=not a.no_graph)  # EOS off: fixed decode length

    def step_resident():
        return model.generate(ids_d, images=pix_d, **gen_kw)

    def step_e2e():
        i = ids_pin.to(dev, non_blocking=True)
        p = pix_pin.to(dev, non_blocking=True)
        out = model.generate(i, images=p, **gen_kw)
        return out.to("cpu")  # device->host read of the step's result (synchronises)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        toks = step_resident()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1), dev)
    launches = eng.launches - l0
    # end-to-end through the public API with HOST buffers (pinned), copies inside the timed region
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        toks_h = step_e2e()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0, dev)
    clocks = sampler.stop() if rank == 0 else None
    all_toks = gather_token_ids(toks, world * B)  # NCCL all-gather of the generated ids, AFTER the timed region
    assert all_toks.shape == (world * B, NN)

    # phase split + dominant-kernel roofline (rank 0, untimed extra work)
    extra = {}
    if rank == 0:
        extra = phase_and_roofline(model, eng, a, ids_d, pix_d, fd, dev)
    total_tokens = world * B * NN * a.steps
    line = {
        "metric": METRIC, "value": total_tokens / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic (seeded randn 336x336 pixels, random ids, N(0,0.02) random-init weights)",
        "config": {"workload": f"BASELINE config 2: {a.preset} bf16, batch={B}/GPU, {T}-id prompt + 1 image (P=607 -> P'=146 at keep 0.2), "
                               f"vision_keep_rate={a.keep_rate}, output_text_keep_rate={a.out_keep} (forced seeded decisions), "
                               f"{NN}-token greedy decode; step = CLIP+projector+splice+prefill+decode of one batch",
                   "global_batch": world * B, "prompt_ids": T, "new_tokens": NN, "parallelism": f"replicas x{world} (no data-path collective)",
                   "l2": "no flush needed: every step streams the 13.5 GB weight set (>> 126 MB L2)"},
        "e2e": {"value": total_tokens / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(ids_pin.numel() * 8 + pix_pin.numel() * 2),
                "d2h_bytes_per_step": int(B * NN * 8)},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    line.update(extra)
    if rank == 0 and "phases" in line and NN > 1:
        line["phases"]["decode_ms_per_token_step"] = (ms / a.steps - line["phases"]["clip_prefill_ms"]) / (NN - 1)
        line["phases"]["decode_tokens_per_s_per_gpu"] = B * 1e3 / line["phases"]["decode_ms_per_token_step"]
    if rank == 0:
        if world == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(a, steps=1, warmup=0)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def phase_and_roofline(model, eng, a, ids_d, pix_d, fd, dev):
    """untimed diagnostics: prefill vs decode split and the skinny-GEMM (weight streaming) roofline."""
    cfg = model.config
    B, NN = a.batch, a.new_tokens
    ev = lambda: torch.cuda.Event(enable_timing=True)
    # phase split: generate with 1 new token == CLIP + prefill (+ first greedy)
    t = []
    for _ in range(3):
        s, e = ev(), ev()
        s.record()
        model.generate(ids_d, images=pix_d, max_new_tokens=1, do_sample=False, eos_token_id=None)
        e.record()
        torch.cuda.synchronize()
        t.append(s.elapsed_time(e))
    prefill_ms = statistics.median(t)
    # dominant kernel: dl_skinny_gemm over every decoder weight of one decode step (each launch reads its weight once)
    sess = next(iter(model._sessions.values()))
    H, I, V = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size
    shapes = []
    for W in eng.packed["layers"]:
        shapes += [(sess.xn, W["wqkv"], 3 * H, H), (sess.attn, W["wo"], H, H), (sess.xn, W["wgu"], 2 * I, H), (sess.act, W["wd"], H, I)]
    shapes.append((sess.xn, model.lm_head.weight, V, H))
    bytes_total = sum(w.numel() * 2 for _, w, _, _ in shapes)
    for _ in range(2):
        for x, w, N, K in shapes:
            eng._gemm(sess, x, w, B, N, K)
    torch.cuda.synchronize()
    # timed inside a CUDA-graph replay, the way the decode step itself runs: an eager `_gemm` call costs ~15-20 us of host time (two
    # ctypes calls, tensor-map encoding), as much as the average launch lasts, so an eager loop measures the host
    reps = 5
    launches_before = eng.launches
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(reps):
                for x, w, N, K in shapes:
                    eng._gemm(sess, x, w, B, N, K)
    torch.cuda.current_stream().wait_stream(side)
    eng.launches = launches_before  # capture only records
    g.replay()
    torch.cuda.synchronize()
    s, e = ev(), ev()
    s.record()
    g.replay()
    e.record()
    torch.cuda.synchronize()
    gemm_ms = s.elapsed_time(e) / reps
    del g
    peak, how = measured_peaks()
    achieved = bytes_total / (gemm_ms / 1e3) / 1e9
    other = kernel_rooflines(dev)
    return {
        "roofline_kernels": other,
        "phases": {"clip_prefill_ms": prefill_ms, "prefill_tokens_per_s": B * 607 / (prefill_ms / 1e3),
                   "decode_ms_per_token_step": None, "weight_gemm_ms_per_decode_step": gemm_ms},
        "roofline": {"bound": "hbm", "kernel": "skinny_gemm_kernel (decode weight streaming, TMA+tcgen05)", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": 1.033 * bytes_total / len(shapes),
                     "traffic_source": "ncu --set full, QKV launch: dram read+write 104.88 MB vs 101.5 MB algorithmic (x1.033), "
                                       "scaled to this step's average launch (profiles/r1_ncu_skinny_gemm_full.md)",
                     "peak_source": how,
                     "algorithmic_bytes_per_launch": bytes_total / len(shapes), "launches_per_decode_step": len(shapes),
                     "avg_launch_us": gemm_ms * 1e3 / len(shapes)},
    }


# ------------------------------------------------------------------------------------------------------------------
def cpu_baseline(a, steps, warmup):
    """The reference algorithm (oracle/dynllava_oracle.py: CPU restatement pinned bit-exact to the reference) timed on this
    box's host cores on a BOUNDED sample of the same workload: same model, B=1, same prompt shape, few new tokens."""
    import psutil
    from dynamic_llava_b200.presets import preset
    from dynamic_llava_b200.synthetic import synthetic_inputs, tiled_state_dict
    from oracle import dynllava_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfgd = preset(a.preset)
    cfgd["sparse_config"]["vision_keep_rate"] = a.keep_rate
    cfgd["sparse_config"]["use_output_text_predictor"] = False  # config 1: reference's image-bench setting
    avail = psutil.virtual_memory().available
    if avail < 20e9:
        return {"value": None, "unit": UNIT, "cores": cores, "kind": "port", "sample": f"skipped: {avail/1e9:.0f} GB host RAM free < 20 GB"}
    t0 = time.perf_counter()
    sd = tiled_state_dict(cfgd, seed=0)
    build_s = time.perf_counter() - t0
    ids, pix = synthetic_inputs(cfgd, 1, a.prompt_len, seed=1)
    n_new = a.cpu_new_tokens
    vals = []
    t_start = time.perf_counter()
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        res = O.greedy_loop(sd, cfgd, ids, pix, n_new)
        dt = time.perf_counter() - t0
        if it >= warmup:
            vals.append(n_new / dt)
        if vals and time.perf_counter() - t_start > 150.0:  # bounded: a long --steps request must still end within a few minutes
            break
    return {"value": statistics.median(vals), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{a.preset} bf16 random-init (tiled), B=1, {a.prompt_len}-id prompt + 1 image, keep {a.keep_rate}, {n_new} greedy tokens "
                      f"(CLIP+prefill+{n_new - 1} decode steps) per step; weights built in {build_s:.0f}s (untimed)",
            "steps": len(vals)}


def gpu_comparator(a):
    """"the same oracle is also timed on the B200 (reference PyTorch path, torch SDPA) as the like-for-like GPU comparator".
    The oracle port with its tensors on cuda:0: eager torch ops, cuBLAS GEMMs, SDPA, a torch.cat-grown KV cache that is re-sliced,
    zero-padded and re-stacked per layer per step for B > 1 (cache_utils.py:165-241) — what a user of the reference runs today.  Full
    BASELINE config 2 workload (batch, prompt, forced Bernoulli decisions, new tokens), CUDA events, inputs resident on the device."""
    from dynamic_llava_b200.presets import preset
    from dynamic_llava_b200.synthetic import forced_text_decisions, synthetic_inputs, tiled_state_dict
    from oracle import dynllava_oracle as O

    dev = torch.device("cuda:0")
    cfgd = preset(a.preset)
    cfgd["sparse_config"]["vision_keep_rate"] = a.keep_rate
    cfgd["sparse_config"]["use_output_text_predictor"] = a.out_keep >= 0
    sd = {k: v.to(dev) for k, v in tiled_state_dict(cfgd, seed=0).items()}
    ids, pix = synthetic_inputs(cfgd, a.batch, a.prompt_len, seed=1)
    ids, pix = ids.to(dev), pix.to(dev)
    fd = forced_text_decisions(a.new_tokens - 1, a.batch, keep_rate=a.out_keep, seed=3).to(dev) if a.out_keep >= 0 else None
    ms = []
    for it in range(a.warmup + max(a.steps, 1)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        O.greedy_loop(sd, cfgd, ids, pix, a.new_tokens, forced_text_decisions=fd)
        e1.record()
        torch.cuda.synchronize()
        if it >= a.warmup:
            ms.append(e0.elapsed_time(e1))
    t = statistics.median(ms)
    v = a.batch * a.new_tokens / (t / 1e3)
    sample = (f"{a.preset} bf16 random-init (tiled), B={a.batch}, {a.prompt_len}-id prompt + 1 image, keep {a.keep_rate}, forced output decisions "
              f"{a.out_keep}, {a.new_tokens} greedy tokens: the oracle port (reference op sequence, eager PyTorch) on {torch.cuda.get_device_name(0)}")
    line = {"impl": "reference", "ref_device": "cuda", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": t, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "BASELINE config 2, full workload, reference PyTorch path on the GPU: " + sample},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": 0, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_reference(a):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    if a.ref_device == "cuda":
        return gpu_comparator(a)
    cb = cpu_baseline(a, steps=max(a.steps, 1), warmup=a.warmup)
    v = cb["value"]
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": None if not v else 1e3 * a.cpu_new_tokens / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "BASELINE config 2 workload, bounded CPU sample: " + cb["sample"]},
            "cpu_baseline": cb, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
