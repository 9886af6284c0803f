"""seeded random sweep of full-size 7B configurations (batch, prompt length, keep rate, batching semantics) through the same checker as
tests/test_full_size_parity_gpu.py (device-resident oracle, teacher-forced logits within 2 * floor + 1e-2 * scale, kv_len exact)

    python tools/full_size_sweep.py [n_cases] [first_seed] [preset]
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import test_full_size_parity_gpu as T
from dynamic_llava_b200.synthetic import forced_text_decisions
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
s0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
preset_name = sys.argv[3] if len(sys.argv) > 3 else "llava-1.5-7b"
dtype = torch.float16 if os.environ.get("DLB200_TEST_DTYPE") == "float16" else torch.bfloat16
bad = 0
for seed in range(s0, s0 + n):
    r = np.random.Generator(np.random.PCG64(4000 + seed))
    B = int(r.choice([1, 2, 3, 4, 5, 6, 7, 9, 10, 12]))
    Tn = int(r.choice([8, 17, 32, 33, 64, 100, 191, 300, 450]))
    rate = float(r.choice([0.1, 0.2, 0.35, 0.5, 0.77, 1.0]))
    quirk = bool(r.integers(0, 2)) or B == 1
    n_new = int(os.environ.get("SWEEP_N_NEW", r.choice([2, 3, 5])))
    fd = forced_text_decisions(n_new - 1, B, keep_rate=[float(x) for x in r.choice([0.0, 0.3, 0.5, 0.8, 1.0], size=B)], seed=seed)
    rows = None if quirk else sorted(set([0, B - 1]))
    tag = f"sweep {preset_name} seed={seed} B={B} T={Tn} keep={rate} quirk={quirk} n_new={n_new}"
    try:
        row = T._run_config(tag, preset_name, B, Tn, n_new, rate, fd, quirk=quirk, per_sample_rows=rows, dtype=dtype)
        print("OK  ", tag, f"err32={row['err_vs_oracle_fp32']:.3f} floor={row['oracle_floor_bf16_vs_fp32']:.3f}", flush=True)
    except AssertionError as e:
        bad += 1
        print("FAIL", tag, str(e)[:300], flush=True)
    torch.cuda.empty_cache()
print(f"{n - bad}/{n} ok")
sys.exit(1 if bad else 0)
