"""full-size 7B parity rows for batch sizes the BASELINE configs do not use (B = 2, 3, 4, 5) — same checker as tests/test_full_size_parity_gpu.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_full_size_parity_gpu as T
from dynamic_llava_b200.synthetic import forced_text_decisions
for B in [int(x) for x in (sys.argv[1:] or ["2", "3", "4", "5"])]:
    fd = forced_text_decisions(3, B, keep_rate=0.5, seed=3)
    try:
        r = T._run_config(f"debug B={B}", "llava-1.5-7b", B, 32, 4, 0.2, fd, quirk=True)
        print("OK", B, r["err_vs_oracle_fp32"], r["oracle_floor_bf16_vs_fp32"])
    except AssertionError as e:
        print("FAIL", B, str(e)[:600])
    torch.cuda.empty_cache()
