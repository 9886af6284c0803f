"""Launch the prefill attention kernels a few times on one shape (for `ncu --set full -k regex:attn_prefill_tc`; FLAGS, B, NH, LQ, REPS from the environment)."""
import ctypes
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynamic_llava_b200._lib import lib  # noqa: E402

L = lib()
B, nh, Lq, hd = int(os.environ.get("B", 4)), int(os.environ.get("NH", 32)), int(os.environ.get("LQ", 2048)), 128
H = nh * hd
cap = Lq + 8
N = B * Lq
P = lambda t: ctypes.c_void_p(t.data_ptr())
ST = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
qkv = torch.randn(N, 3 * H, device="cuda", dtype=torch.bfloat16)
K = torch.randn(B, nh, cap, hd, device="cuda", dtype=torch.bfloat16)
V = torch.randn(B, nh, cap, hd, device="cuda", dtype=torch.bfloat16)
o = torch.empty(N, H, device="cuda", dtype=torch.bfloat16)
cu = torch.arange(0, (B + 1) * Lq, Lq, device="cuda", dtype=torch.int32)
past = torch.zeros(B, device="cuda", dtype=torch.int32)
for _ in range(int(os.environ.get("REPS", 4))):
    L.attn_prefill_tc4(P(qkv), 3 * H, P(K), P(V), P(o), H, P(cu), P(past), N, B, nh, cap, Lq, 1.0 / math.sqrt(hd), int(os.environ.get("FLAGS", 0)), ST())
torch.cuda.synchronize()
print("ok")
