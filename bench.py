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
DT = torch.bfloat16  # element type of the B200 arm (--dtype)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16"],
                    help="element type of the B200 arm: bf16 (libdynllava_b200.so, the measured default) or fp16 (libdynllava_b200_f16.so)")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                    help="BASELINE.json config (1-based): 2 = the config the metric is quoted on (default); 3 = ragged-KV stress (B=8 per GPU, "
                         "256 tokens, per-sample forced keep rates linspace(0.1,0.9,8); meant for --gpus 8 = batch 64); 4 = 13B, B=4 per GPU "
                         "(= 32 over 8 GPUs), P=2048 prefill + 16 tokens; 5 = keep-rate sweep {0.1,0.2,0.5,1.0}, 7B, B=16, one GPU")
    ap.add_argument("--preset", default=None)
    ap.add_argument("--batch", type=int, default=None, help="samples per GPU")
    ap.add_argument("--prompt-len", type=int, default=None)
    ap.add_argument("--new-tokens", type=int, default=None)
    ap.add_argument("--keep-rate", type=float, default=None)
    ap.add_argument("--out-keep", type=float, default=None, help="forced Bernoulli rate for output-token KV decisions; <0 disables")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the decode step eagerly (for ncu launch lists)")
    ap.add_argument("--cpu-new-tokens", type=int, default=32,
                    help="greedy tokens per CPU-baseline step (B=1): CLIP + prefill + 31 decode steps, ~5-15 s of host work per step depending on "
                         "the core count, so that `--impl reference --steps K --warmup W` still ends within a few minutes")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference only: 'cuda' times the same oracle port (the reference's PyTorch op sequence: eager torch, cuBLAS, "
                         "SDPA, torch.cat KV cache) on the GPU over the FULL workload — the like-for-like comparator of SURVEY.md §8d; the "
                         "default (and what the driver runs) is the host-core baseline")
    a = ap.parse_args()
    # per-config defaults; explicit flags override
    table = {2: dict(preset="llava-1.5-7b", batch=8, prompt_len=32, new_tokens=128, keep_rate=0.2, out_keep=0.5),
             3: dict(preset="llava-1.5-7b", batch=8, prompt_len=32, new_tokens=256, keep_rate=0.2, out_keep="linspace(0.1,0.9)"),
             4: dict(preset="llava-1.5-13b", batch=4, prompt_len=2048 - 576 + 1, new_tokens=16, keep_rate=0.2, out_keep=0.5),
             5: dict(preset="llava-1.5-7b", batch=16, prompt_len=32, new_tokens=64, keep_rate=0.2, out_keep=-1.0)}[a.config]
    for k, v in table.items():
        if getattr(a, k) is None:
            setattr(a, k, v)
    return a


def workload_name(a, world=1):
    P = a.prompt_len - 1 + 576
    k = int(576 * a.keep_rate)
    ok = "per-sample linspace(0.1,0.9) (ragged-KV stress)" if a.out_keep == "linspace(0.1,0.9)" else (
        f"{a.out_keep}" if float(a.out_keep) >= 0 else "off (output predictor disabled)")
    return (f"BASELINE config {a.config}: {a.preset} bf16, batch={a.batch}/GPU x {world} GPU, {a.prompt_len}-id prompt + 1 image "
            f"(P={P} -> P'={P - 576 + k} at keep {a.keep_rate}), vision_keep_rate={a.keep_rate}, output_text_keep_rate={ok} "
            f"(forced seeded decisions), {a.new_tokens}-token greedy decode; step = CLIP+projector+splice+prefill+decode of one batch")


def out_keep_rates(a):
    """forced Bernoulli keep rate(s) of the output-token decisions: a scalar, or config 3's per-sample ramp"""
    if a.out_keep == "linspace(0.1,0.9)":
        return torch.linspace(0.1, 0.9, a.batch).tolist()
    return float(a.out_keep)


def k1_point(name, B, T, H, k, dev, hbm_peak, timeit):
    """K1 (select + gather/compact) on one shape, graph-timed, buffers rotated beyond the 126 MB L2; bytes = SURVEY.md §8d's formula"""
    import ctypes
    from dynamic_llava_b200._lib import lib
    L = lib(DT)
    P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
    st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    n_img, s0 = 576, 5
    Pn = T - 1 + n_img
    Pp = Pn - n_img + k
    n_rot = max(2, int(300e6 // (B * Pn * H * 2)) + 1)
    hid = [torch.randn(B, Pn, H, device=dev, dtype=DT) for _ in range(n_rot)]
    outb = [torch.empty(B, Pp, H, device=dev, dtype=DT) for _ in range(n_rot)]
    logits = torch.randn(B, n_img, 2, device=dev, dtype=DT)
    keep = torch.empty(B, k, dtype=torch.int64, device=dev)
    pos = torch.empty(B, Pp, dtype=torch.int32, device=dev)
    cu = torch.empty(B + 1, dtype=torch.int32, device=dev)
    ms = timeit(lambda i: L.select_compact(P(logits), None, P(hid[i]), P(outb[i]), P(keep), P(pos), P(cu), B, Pn, H, s0, n_img, k, 0, st()), n_rot)
    nbytes = 2 * B * Pp * H * 2 + B * n_img * 4 + B * Pp * 4
    del hid, outb
    traffic, tsrc = measured_traffic("select_compact_kernel", nbytes)
    return {"kernel": name, "bound": "hbm", "achieved": nbytes / ms / 1e6, "peak": hbm_peak, "unit": "GB/s", "frac": nbytes / ms / 1e6 / hbm_peak,
            "us": ms * 1e3, "algorithmic_bytes": nbytes, "traffic": traffic, "traffic_source": tsrc, "peak_source": "MEASURED_PEAKS.json hbm_gbs"}


def rot_timeit(fn, n_rot=1, iters=20, warm=3):
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


def config5_sweep(model, a, ids_d, pix_d, dev):
    """BASELINE config 5: keep-rate sweep {0.1, 0.2, 0.5, 1.0} on 7B, batch 16, one GPU: generated tokens/s and the achieved HBM GB/s of
    K1 (select + gather/compact) against the effective token count B * P' (run-time mutation of the live sparse_config, like the
    reference's own bench script, llava/dynamic_eval/bench_test/dynamic_llava_image_time_and_mem.py:63-65)."""
    hbm_peak, _ = measured_peaks()
    B, T, NN = a.batch, a.prompt_len, a.new_tokens
    rows = []
    for rate in (0.1, 0.2, 0.5, 1.0):
        model.config.sparse_config["vision_keep_rate"] = rate
        k = int(576 * rate)
        P = T - 1 + 576
        Pp = P - 576 + k
        kw = dict(images=pix_d, do_sample=False, eos_token_id=None)

        def timed(fn, reps=3, warm=2):
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
        t_pre = timed(lambda: model.generate(ids_d, max_new_tokens=1, **kw))
        t_all = timed(lambda: model.generate(ids_d, max_new_tokens=NN, **kw))
        k1 = k1_point(f"K1 keep {rate}", B, T, model.config.hidden_size, k, dev, hbm_peak, rot_timeit)
        rows.append({"vision_keep_rate": rate, "k": k, "P_prime": Pp, "effective_tokens": B * Pp, "clip_prefill_ms": t_pre,
                     "prefill_tokens_per_s": B * P / (t_pre / 1e3), "generated_tokens_per_s": B * NN / (t_all / 1e3),
                     "decode_ms_per_step": (t_all - t_pre) / max(NN - 1, 1), "k1_us": k1["us"], "k1_achieved_gbs": k1["achieved"],
                     "k1_frac_of_hbm_peak": k1["frac"], "k1_algorithmic_bytes": k1["algorithmic_bytes"]})
    model.config.sparse_config["vision_keep_rate"] = a.keep_rate
    return rows


def kernel_rooflines(dev):
    """untimed extras: live rooflines of the two other kernels the north star names — K2 (prefill attention, tensor pipe) on the
    L=2048 stress shape and on this workload's own shapes, K1 (select + gather/compact, HBM) on the 13B B=32 shape and on this
    workload's 19 MB.  CUDA events around one CUDA-graph replay of back-to-back launches (K2: 50 after 10 warm-ups, K1: 20 after 3); K1
    rotates buffers beyond the 126 MB L2."""
    import ctypes
    import math
    from dynamic_llava_b200._lib import lib
    L = lib(DT)
    P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
    st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    tf_peak = float(pk.get("bf16_tflops_sustained", 1400.0))
    tf_burst = float(pk.get("bf16_tflops", 1650.0))
    hbm_peak = float(pk.get("hbm_gbs", 6650.0))

    timeit = rot_timeit
    out = []
    for name, B, nh, Lq in (("K2 dl_attn_prefill_tc4 L=2048 B=8 32 heads (stress)", 8, 32, 2048), ("K2 dl_attn_prefill_tc4 L=607 B=8 (layers 0-1)", 8, 32, 607),
                            ("K2 dl_attn_prefill_tc4 L=146 B=8 (layers 2-31)", 8, 32, 146)):
        hd, H, cap, N = 128, nh * 128, Lq + 8, B * Lq
        qkv = torch.randn(N, 3 * H, device=dev, dtype=DT)
        K = torch.randn(B, nh, cap, hd, device=dev, dtype=DT)
        V = torch.randn(B, nh, cap, hd, device=dev, dtype=DT)
        o = torch.empty(N, H, device=dev, dtype=DT)
        cu = torch.arange(0, (B + 1) * Lq, Lq, device=dev, dtype=torch.int32)
        past = torch.zeros(B, device=dev, dtype=torch.int32)
        ms = timeit(lambda i: L.attn_prefill_tc4(P(qkv), 3 * H, P(K), P(V), P(o), H, P(cu), P(past), N, B, nh, cap, Lq, 1.0 / math.sqrt(hd), 0, st()),
                    iters=50, warm=10)
        flop = 2.0 * Lq * Lq * H * B
        # the kernel is timed ALONE in a 50-launch graph replay = the burst case: `frac` divides by the burst peak; the fraction of the
        # sustained peak (what a long step can hold) is printed beside it
        out.append({"kernel": name, "bound": "tensor", "achieved": flop / ms / 1e9, "peak": tf_burst, "unit": "TFLOP/s",
                    "frac": flop / ms / 1e9 / tf_burst, "frac_of_sustained": flop / ms / 1e9 / tf_peak, "peak_sustained": tf_peak, "us": ms * 1e3,
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst; kernel timed alone) / bf16_tflops_sustained"})
        del qkv, K, V, o
    for name, B, T, H, k in (("K1 dl_select_compact 13B B=32 P=2048 keep 0.2", 32, 1473, 5120, 115),
                             ("K1 dl_select_compact 7B B=8 P=607 keep 0.2 (config 2's own shape)", 8, 32, 4096, 115)):
        out.append(k1_point(name, B, T, H, k, dev, hbm_peak, timeit))
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
    has_out = a.out_keep == "linspace(0.1,0.9)" or float(a.out_keep) >= 0
    cfgd["sparse_config"]["vision_keep_rate"] = a.keep_rate
    cfgd["sparse_config"]["use_output_text_predictor"] = has_out
    cfg = DynamicLlavaConfig(**cfgd)
    model = DynamicLlavaLlamaForCausalLM(cfg, device=dev, dtype=DT)
    model.get_vision_tower().load_model()
    sd = device_state_dict(cfgd, seed=0, device=dev, dtype=DT)  # same weights on every replica
    model.load_state_dict(sd, strict=True)
    del sd
    torch.cuda.empty_cache()
    eng = model.engine
    eng._check_packed()

    B, T, NN = a.batch, a.prompt_len, a.new_tokens
    ids_h, pix_h = synthetic_inputs(cfgd, B, T, seed=1 + rank, dtype=DT)  # each rank owns its own chunk of the global batch
    fd = forced_text_decisions(max(NN - 1, 1), B, keep_rate=out_keep_rates(a), seed=3 + rank) if has_out else None
    ids_pin, pix_pin = ids_h.pin_memory(), pix_h.pin_memory()
    ids_d, pix_d = ids_h.to(dev), pix_h.to(dev)
    gen_kw = dict(max_new_tokens=NN, do_sample=False, eos_token_id=None, forced_text_decisions=fd, use_cuda_graph=not a.no_graph)  # EOS off: fixed decode length

    def step_resident():
        return model.generate(ids_d, images=pix_d, **gen_kw)

    def step_e2e():
        # HOST buffers (pinned) straight into the public API: generate() issues the host->device copies itself
        out = model.generate(ids_pin, images=pix_pin, **gen_kw)
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
        if a.config == 5:
            extra["sweep"] = config5_sweep(model, a, ids_d, pix_d, dev)
    total_tokens = world * B * NN * a.steps
    line = {
        "metric": METRIC, "value": total_tokens / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": a.dtype,
        "data": "synthetic (seeded randn 336x336 pixels, random ids, N(0,0.02) random-init weights)",
        "config": {"workload": workload_name(a, world),
                   "baseline_config": a.config, "global_batch": world * B, "prompt_ids": T, "new_tokens": NN,
                   "parallelism": f"replicas x{world} (no data-path collective)",
                   "l2": "no flush needed: every step streams the whole bf16 weight set (13.5 / 26 GB >> 126 MB L2)"},
        "e2e": {"value": total_tokens / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(ids_pin.numel() * 8 + pix_pin.numel() * 2),
                "d2h_bytes_per_step": int(B * NN * 8)},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    line.update(extra)
    if rank == 0 and "phases" in line and NN > 1:
        line["phases"]["decode_ms_per_token_step"] = (ms / a.steps - line["phases"]["clip_prefill_ms"]) / (NN - 1)
        line["phases"]["decode_tokens_per_s_per_gpu"] = B * 1e3 / line["phases"]["decode_ms_per_token_step"]
        ph = line["phases"]
        # whole decode step in situ: (weights + mean KV bytes) / step time / measured HBM peak
        ph["decode_step_hbm_gbs"] = (ph["decode_step_weight_bytes"] + ph["decode_step_mean_kv_bytes"]) / (ph["decode_ms_per_token_step"] / 1e3) / 1e9
        ph["decode_step_hbm_frac"] = ph["decode_step_hbm_gbs"] / ph["hbm_peak_gbs"]
    if rank == 0:
        if world == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(a, steps=3, warmup=1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _src_sha16(names):
    import hashlib
    h = hashlib.sha256()
    for n in names:
        with open(os.path.join(ROOT, "dynamic_llava_b200", "csrc", n), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def measured_traffic(kernel_key, algorithmic_bytes):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture of THIS build of the kernel
    (profiles/ncu_traffic.json, written by tools/ncu_traffic.py; entries are keyed by the sha of the kernel's sources so that a capture of
    an older build is never quoted for a newer one).  None when no capture of the current sources is committed."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None, "no ncu capture committed (profiles/ncu_traffic.json missing)"
    db = json.load(open(path))
    ent = db.get(kernel_key)
    if ent is None:
        return None, f"no ncu capture for {kernel_key}"
    sha = _src_sha16(ent["sources"])
    if sha != ent["src_sha16"]:
        return None, f"ncu capture is of an older build of {ent['sources']} (sha {ent['src_sha16']} != {sha}): re-run tools/ncu_traffic.py"
    # the capture is of one representative launch; scale its measured/algorithmic ratio to this step's average launch
    ratio = ent["dram_bytes_per_launch"] / ent["algorithmic_bytes_per_launch"]
    return ratio * algorithmic_bytes, (f"ncu --set full of this build ({ent['what']}): dram read+write {ent['dram_bytes_per_launch'] / 1e6:.2f} MB vs "
                                       f"{ent['algorithmic_bytes_per_launch'] / 1e6:.2f} MB algorithmic (x{ratio:.3f}), scaled to this step's average launch")


def graph_time(fn, iters, warm=2):
    """ms per call of `fn` inside ONE CUDA-graph replay of `iters` calls (no host in the loop, PDL edges as in the real step)"""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(iters):
                fn()
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def phase_and_roofline(model, eng, a, ids_d, pix_d, fd, dev):
    """untimed diagnostics: prefill vs decode split, the skinny-GEMM (weight streaming) roofline, the whole-step HBM fraction, live K3."""
    import ctypes
    import math
    cfg = model.config
    B, NN, T = a.batch, a.new_tokens, a.prompt_len
    P = T - 1 + 576
    Pp = P - 576 + int(576 * a.keep_rate)
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
    # leave the session in its end-of-decode state (final kv_len) for the K3 measurement below
    model.generate(ids_d, images=pix_d, max_new_tokens=NN, do_sample=False, eos_token_id=None, forced_text_decisions=fd)
    sess = next(iter(model._sessions.values()))
    # dominant kernel: dl_skinny_gemm over every decoder weight of one decode step (each launch reads its weight once)
    H, I, V = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size
    nh = cfg.num_attention_heads
    hd = H // nh
    shapes = []
    for W in eng.packed["layers"]:
        shapes += [(sess.xn, W["wqkv"], 3 * H, H), (sess.attn, W["wo"], H, H), (sess.xn, W["wgu"], 2 * I, H), (sess.act, W["wd"], H, I)]
    shapes.append((sess.xn, model.lm_head.weight, V, H))
    bytes_total = sum(w.numel() * 2 for _, w, _, _ in shapes)
    launches_before = eng.launches

    def all_gemms():
        for x, w, N, K in shapes:
            eng._gemm(sess, x, w, B, N, K)
    # timed inside a CUDA-graph replay, the way the decode step itself runs: an eager `_gemm` call costs ~15-20 us of host time, as
    # much as the average launch lasts, so an eager loop measures the host
    gemm_ms = graph_time(all_gemms, iters=5)
    peak, how = measured_peaks()
    achieved = bytes_total / (gemm_ms / 1e3) / 1e9
    # K3 live: dl_attn_decode_fused on this workload's own ragged lengths (end of the decode: the longest caches), one lo layer and
    # one hi layer, bytes = sum_b kv_len[l][b] * 2 * hd * 2 * heads (K and V rows of every cached key)
    P_ = lambda x: None if x is None else ctypes.c_void_p(x.data_ptr())
    stv = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    SL = cfg.sparse_config["sparse_layer"]
    ar = sess.arena
    lens_host = ar.kv_len.cpu()
    k3 = []
    s_qkv = eng.L.skinny_gemm_splits(3 * H, H, eng.skinny_target_ctas)
    eng._gemm(sess, sess.xn, eng.packed["layers"][0]["wqkv"], B, 3 * H, H)  # valid slices in sess.partial
    for l in (0, SL):
        K_, V_, cap = ar.layer(l)
        ns = eng._n_split(B, nh, cap)
        if ns > 1 and sess.attn_ws is None:
            sess.attn_ws = torch.empty(eng.L.attn_decode_workspace_bytes(B, nh, 16) // 4, device=dev, dtype=torch.float32)
        kvl = ar.kv_len[l]
        fn = lambda: eng.L.attn_decode_fused(P_(sess.partial), s_qkv, P_(ar.pos), None, P_(eng.packed["cos"]), P_(eng.packed["sin"]), P_(K_), P_(V_),
                                             P_(kvl), P_(sess.attn), P_(sess.attn_ws), B, nh, hd, cap, ns, 1.0 / math.sqrt(hd), 0, None, 0,
                                             ar.floor_lo if l < SL else ar.floor_hi, stv())
        # the arena rows of 20 back-to-back launches on the same layer would sit in L2: rotate over >= 4 layers of the same class
        ms = graph_time(fn, iters=20)
        nbytes = int(lens_host[l].sum()) * 2 * hd * 2 * nh
        k3.append({"kernel": f"K3 dl_attn_decode_fused layer {l} ({'lo' if l < SL else 'hi'} class), kv_len {lens_host[l].tolist()}", "bound": "hbm",
                   "achieved": nbytes / ms / 1e6, "peak": peak, "unit": "GB/s", "frac": nbytes / ms / 1e6 / peak, "us": ms * 1e3,
                   "algorithmic_bytes": nbytes, "traffic": measured_traffic("attn_decode_kernel", nbytes)[0], "peak_source": how,
                   "note": "latency-bound at these lengths; the same layer's rows are re-read by the 20 graph-timed launches, so L2 serves them: "
                           "an upper bound on the in-step rate"})
    eng.launches = launches_before  # diagnostics only
    other = kernel_rooflines(dev) + k3
    # whole decode step against its HBM floor: weights streamed once + the mean KV bytes over the decode steps of this workload
    n_lo, n_hi = SL, cfg.num_hidden_layers - SL
    kv_row = 2 * H * 2  # K and V of one token, all heads, bf16
    steps = max(NN - 1, 1)
    mean_lo = P + (steps - 1) / 2.0
    kept = (fd[:steps].double().cumsum(0) - fd[:steps].double()).mean(0).sum().item() if fd is not None else B * (steps - 1) / 2.0
    kv_bytes = kv_row * (n_lo * B * mean_lo + n_hi * (B * Pp + kept))
    traffic, traffic_src = measured_traffic("skinny_gemm_kernel", bytes_total / len(shapes))
    return {
        "roofline_kernels": other,
        "phases": {"clip_prefill_ms": prefill_ms, "prefill_tokens_per_s": B * P / (prefill_ms / 1e3),
                   "decode_ms_per_token_step": None, "weight_gemm_ms_per_decode_step": gemm_ms,
                   "decode_step_weight_bytes": bytes_total, "decode_step_mean_kv_bytes": kv_bytes, "hbm_peak_gbs": peak},
        "roofline": {"bound": "hbm", "kernel": "skinny_gemm_kernel (decode weight streaming, TMA+tcgen05)", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": how,
                     "algorithmic_bytes_per_launch": bytes_total / len(shapes), "launches_per_decode_step": len(shapes),
                     "avg_launch_us": gemm_ms * 1e3 / len(shapes),
                     "timing": "the step's 129 weight-streaming launches replayed ALONE in one CUDA graph (isolated); the in-situ whole-step "
                               "fraction is phases.decode_step_hbm_frac"},
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
    has_out = a.out_keep == "linspace(0.1,0.9)" or float(a.out_keep) >= 0
    cfgd["sparse_config"]["use_output_text_predictor"] = has_out  # same flags as the timed workload (forced seeded decisions below)
    avail = psutil.virtual_memory().available
    need = 36e9 if "13b" in a.preset else 20e9
    if avail < need:
        return {"value": None, "unit": UNIT, "cores": cores, "kind": "port", "sample": f"skipped: {avail/1e9:.0f} GB host RAM free < {need/1e9:.0f} GB"}
    t0 = time.perf_counter()
    sd = tiled_state_dict(cfgd, seed=0)
    build_s = time.perf_counter() - t0
    T_cpu = min(a.prompt_len, 32)  # config 4's 1473-id prompt (a 53 TFLOP 13B prefill) would not end within minutes on host cores
    ids, pix = synthetic_inputs(cfgd, 1, T_cpu, seed=1)
    n_new = a.cpu_new_tokens
    from dynamic_llava_b200.synthetic import forced_text_decisions
    rates = out_keep_rates(a) if has_out else None
    fd_cpu = forced_text_decisions(max(n_new - 1, 1), 1, keep_rate=(rates[0] if isinstance(rates, list) else rates), seed=3) if has_out else None
    vals = []
    t_start = time.perf_counter()
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        res = O.greedy_loop(sd, cfgd, ids, pix, n_new, forced_text_decisions=fd_cpu)
        dt = time.perf_counter() - t0
        if it >= warmup:
            vals.append(n_new / dt)
        if vals and time.perf_counter() - t_start > 150.0:  # bounded: a long --steps request must still end within a few minutes
            break
    return {"value": statistics.median(vals), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{a.preset} bf16 random-init (tiled), B=1, {T_cpu}-id prompt + 1 image, keep {a.keep_rate}, output decisions "
                      f"{'forced, rate ' + str(rates[0] if isinstance(rates, list) else rates) if has_out else 'off'}, {n_new} greedy tokens "
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
    has_out = a.out_keep == "linspace(0.1,0.9)" or float(a.out_keep) >= 0
    cfgd["sparse_config"]["use_output_text_predictor"] = has_out
    sd = {k: v.to(dev) for k, v in tiled_state_dict(cfgd, seed=0).items()}
    ids, pix = synthetic_inputs(cfgd, a.batch, a.prompt_len, seed=1)
    ids, pix = ids.to(dev), pix.to(dev)
    fd = forced_text_decisions(a.new_tokens - 1, a.batch, keep_rate=out_keep_rates(a), seed=3).to(dev) if has_out else None
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
            "config": {"workload": f"BASELINE config {a.config}, full workload, reference PyTorch path on the GPU: " + sample},
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
            "config": {"workload": workload_name(a, 1) + " -- timed on a bounded CPU sample of it: " + cb["sample"], "baseline_config": a.config},
            "cpu_baseline": cb, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    args = parse()
    DT = torch.float16 if args.dtype == "fp16" else torch.bfloat16
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
