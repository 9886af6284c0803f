"""Measured DRAM traffic of the dominant kernels for bench.py's `roofline.traffic` (run ON the GPU box):

    python tools/ncu_traffic.py            # drives itself under `ncu --set full`, parses the report, writes gpurun_out/ncu_traffic.json

One representative launch per kernel (the decode step's QKV weight-streaming GEMM, K1 on the 13B B=32 shape, K3 on config 2's lengths,
the prefill GEMM on the 7B QKV shape); entries are keyed by the sha of the kernel's sources so bench.py never quotes a capture of an
older build.  Copy the JSON to profiles/ncu_traffic.json and commit it with the build it describes.
"""
import csv
import ctypes
import hashlib
import io
import json
import math
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CSRC = os.path.join(ROOT, "dynamic_llava_b200", "csrc")


def sha16(names):
    h = hashlib.sha256()
    for n in names:
        h.update(open(os.path.join(CSRC, n), "rb").read())
    return h.hexdigest()[:16]


def child():
    import torch
    from dynamic_llava_b200._lib import lib
    L = lib()
    DEV, BF = "cuda", torch.bfloat16
    P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
    ST = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    M, H, nh, hd = 8, 4096, 32, 128
    x = torch.randn(M, H, device=DEV, dtype=BF)
    w = torch.randn(3 * H, H, device=DEV, dtype=BF) * 0.02
    s = L.skinny_gemm_splits(3 * H, H, 296)
    part = torch.empty(s * M * 3 * H, device=DEV, dtype=torch.float32)
    for _ in range(3):
        L.skinny_gemm(P(x), P(w), P(part), M, 3 * H, H, s, -1, ST())
    # prefill GEMM, 7B QKV shape
    Mp = 4856
    xp = torch.randn(Mp, H, device=DEV, dtype=BF)
    op = torch.empty(Mp, 3 * H, device=DEV, dtype=BF)
    ws = torch.zeros(L.linear_workspace_bytes(), device=DEV, dtype=torch.uint8)
    for _ in range(3):
        L.linear(P(xp), P(w), None, None, P(op), Mp, 3 * H, H, 0, P(ws), ST())
    # K1, 13B B=32 P=2048 keep 0.2
    B, T, Hk, k = 32, 1473, 5120, 115
    Pn = T - 1 + 576
    Pp = Pn - 576 + k
    hid = torch.randn(B, Pn, Hk, device=DEV, dtype=BF)
    outb = torch.empty(B, Pp, Hk, device=DEV, dtype=BF)
    lg = torch.randn(B, 576, 2, device=DEV, dtype=BF)
    keep = torch.empty(B, k, dtype=torch.int64, device=DEV)
    pos = torch.empty(B, Pp, dtype=torch.int32, device=DEV)
    cu = torch.empty(B + 1, dtype=torch.int32, device=DEV)
    for _ in range(3):
        L.select_compact(P(lg), None, P(hid), P(outb), P(keep), P(pos), P(cu), B, Pn, Hk, 5, 576, k, 0, ST())
    # K3 on config 2's lo class (735 keys)
    cap = 800
    K_ = torch.randn(M, nh, cap, hd, device=DEV, dtype=BF)
    V_ = torch.randn(M, nh, cap, hd, device=DEV, dtype=BF)
    kvl = torch.full((M,), 735, dtype=torch.int32, device=DEV)
    posd = torch.full((M,), 735, dtype=torch.int32, device=DEV)
    cos = torch.randn(4096, hd, device=DEV, dtype=BF)
    attn = torch.empty(M, H, device=DEV, dtype=BF)
    for _ in range(3):
        L.attn_decode_fused(P(part), s, P(posd), None, P(cos), P(cos), P(K_), P(V_), P(kvl), P(attn), None, M, nh, hd, cap, 1, 1 / math.sqrt(hd), 0,
                            None, 0, 704, ST())
    torch.cuda.synchronize()
    print("child ok")


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        return child()
    rep = os.path.join(ROOT, "gpurun_out", "ncu_traffic")
    os.makedirs(os.path.dirname(rep), exist_ok=True)
    subprocess.run(["ncu", "--set", "full", "--clock-control", "none", "-k", "regex:skinny_gemm_kernel|gemm2_kernel|select_compact|attn_decode_kernel",
                    "-f", "-o", rep, sys.executable, os.path.abspath(__file__), "--child"], check=True, timeout=600)
    raw = subprocess.run(["ncu", "-i", rep + ".ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    ik, ir, iw, it = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    units = rows[1]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    last = {}
    for r in rows[2:]:
        last[r[ik].split("(")[0].replace("void ", "")] = (float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]], float(r[it]))
    spec = {
        "skinny_gemm_kernel": (["skinny_gemm.cu", "common.cuh"], 3 * 4096 * 4096 * 2 + 8 * 4096 * 2 + 3 * 8 * 3 * 4096 * 4,
                               "decode-step QKV projection, 7B, M=8, split 3: W once + X + fp32 slices"),
        "gemm2_kernel": (["gemm2.cu", "common.cuh"], (4856 * 4096 + 3 * 4096 * 4096 + 4856 * 3 * 4096) * 2,
                         "prefill QKV projection, 7B, M=4856: X + W + out once"),
        "select_compact_kernel": (["select_compact.cu", "common.cuh"], 2 * 32 * 1587 * 5120 * 2 + 32 * 576 * 4 + 32 * 1587 * 4,
                                  "K1, 13B B=32 P=2048 keep 0.2"),
        "attn_decode_kernel": (["attn_decode.cu", "common.cuh"], 8 * 32 * 735 * 2 * 128 * 2, "K3, config 2 lo class: 8 samples x 32 heads x 735 keys"),
    }
    out = {}
    for name, (dram, us) in last.items():
        for key, (srcs, alg, what) in spec.items():
            if name.startswith(key):
                out[key] = {"sources": srcs, "src_sha16": sha16(srcs), "dram_bytes_per_launch": dram, "algorithmic_bytes_per_launch": alg,
                            "ncu_duration_us": us, "what": what, "kernel": name}
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ncu_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
