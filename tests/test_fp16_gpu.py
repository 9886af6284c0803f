"""float16 build (libdynllava_b200_f16.so): the reference's eval drivers load torch_dtype=float16
(llava/model/dynamic_llava_builder.py:62), so the same kernels are compiled a second time with the element type switched and
must hold the same parity bars.

Two layers of checks:
  * the fp16 fixtures produced by the UNMODIFIED reference in float16 (tests/golden/toy_fp16_*.pt; the CPU oracle is bit-exact
    against them, tests/test_oracle_golden.py) drive the public API directly here;
  * the whole per-kernel and model-level GPU suites are re-run in a child pytest with DLB200_TEST_DTYPE=float16: every tensor,
    the kernel library and the oracle side then use float16 (full-size cases excluded to bound the run time; one full-size
    float16 row lives in tests/test_full_size_parity_gpu.py).
"""
import os
import subprocess
import sys

import pytest
import torch

from dynamic_llava_b200.presets import preset
from dynamic_llava_b200.synthetic import numpy_state_dict, synthetic_inputs
from oracle import dynllava_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
F16 = torch.float16


def test_both_libraries_load_with_their_element_type():
    from dynamic_llava_b200._lib import lib
    assert lib().element_type() == 0 and lib(torch.bfloat16).element_type() == 0 and lib(F16).element_type() == 1
    assert lib(F16) is not lib() and lib(F16).version() == lib().version()
    with pytest.raises(RuntimeError, match="no kernel library"):
        lib(torch.float32)


@pytest.mark.parametrize("name", ["toy_fp16_b1_natural", "toy_fp16_b2_forced"])
def test_generate_against_reference_fp16_fixture(name):
    """greedy generate() in float16 against the reference's own float16 run: forced masks on both sides, logits within
    2 * floor + 1e-2 * scale of the fp32 oracle, tokens equal wherever the margin exceeds 4 * floor, KV lengths exact."""
    from dynamic_llava_b200 import DynamicLlavaConfig, DynamicLlavaLlamaForCausalLM
    g = torch.load(os.path.join(GOLD, name + ".pt"), weights_only=False)
    cfg = preset("toy")
    cfg["sparse_config"]["use_output_text_predictor"] = True
    sd32 = numpy_state_dict(cfg, seed=g["weight_seed"])
    sd16 = {k: v.to(F16) for k, v in sd32.items()}
    model = DynamicLlavaLlamaForCausalLM.from_state_dict(DynamicLlavaConfig(**cfg), sd16, device="cuda", dtype=F16)
    assert model.dtype == F16
    ids, pix = synthetic_inputs(cfg, g["B"], g["T"], seed=g["input_seed"], dtype=F16)
    rec = {}
    O.model_forward(sd16, cfg, ids, images=pix, record=rec)
    keep = rec["keep_index"]
    per_sample = g["B"] > 1  # our default B>1 semantics == each sample alone; the batched quirk has its own bf16/fp16 child-run test
    fd = g["forced_decisions"]
    if fd is None:
        fd = O.greedy_loop(sd16, cfg, ids, pix, g["n_new"], forced_keep_index=keep, teacher_tokens=g["tokens"])["decisions"]
    kw = dict(forced_keep_index=keep, forced_text_decisions=fd, per_sample=per_sample)
    o16 = O.greedy_loop(sd16, cfg, ids, pix, g["n_new"], **kw)
    if not per_sample:  # B == 1: the oracle's float16 run IS the reference's (same masks, same tokens)
        assert torch.equal(o16["tokens"], g["tokens"])
    o32 = O.greedy_loop(sd32, cfg, ids, pix.float(), g["n_new"], teacher_tokens=o16["tokens"], **kw)
    floor = (o16["logits"] - o32["logits"]).abs().max().item()
    toks = model.generate(ids.cuda(), images=pix.cuda(), max_new_tokens=g["n_new"], do_sample=False, eos_token_id=None,
                          forced_keep_index=keep, forced_text_decisions=fd).cpu()
    lens = next(iter(model._sessions.values())).arena.kv_len.cpu().long()
    assert torch.equal(lens, o16["true_len"]), (lens, o16["true_len"])
    top2 = o32["logits"].topk(2, dim=-1).values
    sure = ((top2[..., 0] - top2[..., 1]) > 4 * floor).T  # [B, n_new]
    # a token can only be compared while the GPU run is still on the oracle's trajectory
    same_prefix = torch.cumprod((toks == o16["tokens"]).int(), dim=1).bool()
    on_track = torch.cat([torch.ones_like(same_prefix[:, :1]), same_prefix[:, :-1]], dim=1)
    assert torch.equal(toks[sure & on_track], o16["tokens"][sure & on_track])
    assert float((sure & on_track).float().mean()) > 0.3
    # float16 keeps 3 more mantissa bits than bfloat16: the floor itself must be well below the bf16 one on this model
    assert floor < 0.05, floor


def test_gpu_suites_pass_in_float16():
    """re-run the per-kernel, model-level and random-sweep GPU tests with every tensor, the library and the oracle in float16"""
    env = dict(os.environ, DLB200_TEST_DTYPE="float16")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-k", "not full_size", "-p", "no:cacheprovider",
                        os.path.join(ROOT, "tests", "test_kernels_gpu.py"), os.path.join(ROOT, "tests", "test_model_gpu.py"),
                        os.path.join(ROOT, "tests", "test_fuzz_gpu.py")],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    tail = r.stdout[-3000:] + r.stderr[-1500:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and " failed" not in r.stdout, tail
