"""Host-side logic that needs no GPU: state-dict key contract, ragged arena + legacy view semantics, sharding (gloo,
world_size 2), synthetic input determinism."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dynamic_llava_b200 import DynamicLlavaConfig, DynamicLlavaLlamaForCausalLM
from dynamic_llava_b200.cache import LegacyPastView, RaggedKVArena
from dynamic_llava_b200.presets import SPARSE_DEFAULTS, preset
from dynamic_llava_b200.sharding import chunk_bounds, gather_token_ids, get_chunk, max_over_ranks
from dynamic_llava_b200.synthetic import forced_text_decisions, numpy_state_dict, synthetic_inputs, weight_shapes


def test_state_dict_keys_match_reference_contract():
    cfg = DynamicLlavaConfig.from_preset("toy")
    cfg.sparse_config["use_output_text_predictor"] = True
    m = DynamicLlavaLlamaForCausalLM(cfg, device="cpu")
    m.get_vision_tower().load_model()
    keys = set(m.state_dict())
    want = {k for k, _, _ in weight_shapes(cfg.to_dict())}
    assert keys == want
    for k in ("model.embed_tokens.weight", "model.layers.0.self_attn.q_proj.weight", "model.layers.3.mlp.down_proj.weight",
              "model.norm.weight", "lm_head.weight", "model.mm_projector.0.bias", "model.mm_projector.2.weight",
              "model.image_score_predictor.down_mlp.0.weight", "model.image_score_predictor.transformer.1.attn.qkv.weight",
              "model.image_score_predictor.output_mlp.4.bias", "model.output_text_score_predictor.output_mlp.7.weight",
              "model.vision_tower.vision_tower.vision_model.embeddings.class_embedding",
              "model.vision_tower.vision_tower.vision_model.encoder.layers.0.self_attn.out_proj.bias",
              "model.vision_tower.vision_tower.vision_model.pre_layrnorm.weight"):
        assert k in keys, k
    assert "model.image_score_predictor.transformer.0.attn.qkv.bias" not in keys  # qkv_bias=False
    sd = numpy_state_dict(cfg.to_dict(), 0)
    assert m.load_state_dict(sd, strict=True)


def test_sparse_config_defaults_and_keep_counts():
    assert SPARSE_DEFAULTS["sparse_layer"] == 2 and SPARSE_DEFAULTS["vision_keep_rate"] == 0.2
    assert [int(576 * r) for r in (0.1, 0.2, 0.5, 1.0)] == [57, 115, 288, 576]
    cfg = DynamicLlavaConfig.from_preset("llava-1.5-7b")
    cfg.sparse_config["vision_keep_rate"] = 0.5  # live, mutable dict
    assert cfg.sparse_config["vision_keep_rate"] == 0.5 and preset("llava-1.5-7b")["sparse_config"]["vision_keep_rate"] == 0.2


def test_arena_legacy_view_is_zero_padded_ragged():
    a = RaggedKVArena(n_layers=4, sparse_layer=2, batch=2, heads=2, head_dim=8, cap_lo=16, cap_hi=12, device="cpu", dtype=torch.float32)
    # emulate: prefill 10 / 6 tokens, then one decode step kept by sample 1 only
    for l in range(4):
        k, v, cap = a.layer(l)
        n = 10 if l < 2 else 6
        k[:, :, :n] = 1.0
        v[:, :, :n] = 2.0
        a.kv_len[l] = n
    a.host_len_lo, a.host_ub_hi = 10, 6
    for l in range(4):
        k, v, _ = a.layer(l)
        if l < 2:
            k[:, :, 10] = 3.0
            a.kv_len[l] += 1
        else:
            k[1, :, 6] = 3.0
            a.kv_len[l, 1] += 1
    a.host_len_lo += 1
    a.host_ub_hi += 1
    a.hi_exact = False
    past = LegacyPastView(a)
    assert len(past) == 2 and len(past[0]) == 4
    assert past[0][0][0].shape == (2, 2, 11, 8)            # dynamic_modeling_llama.py:2831 reads .shape[2]
    assert past[0][-1][0].shape[-2] == 7                   # model_lvis_for_meteor.py:177
    assert (past[0][-1][0][0, :, 6] == 0).all()            # shorter sample: zero pad, like cache_utils.py:204-235
    assert (past[0][-1][0][1, :, 6] == 3).all()
    tl = past[1]
    assert [t.tolist() for t in tl] == [[11, 11], [11, 11], [6, 7], [6, 7]] and tl[0].dtype == torch.int64
    a.grow(40, 30)
    assert a.layer(0)[0].shape[2] == 40 and (a.layer(3)[0][1, :, 6] == 3).all() and a.room() == 23
    a.reset()
    assert int(a.kv_len.sum()) == 0 and float(a.k_hi.abs().sum()) == 0


def test_chunk_rule_matches_reference_split_list():
    import math
    for n in (0, 1, 7, 8, 9, 64, 65):
        for w in (1, 2, 4, 8):
            size = math.ceil(n / w) if n else 0
            ref = [list(range(n))[i:i + size] for i in range(0, n, size)] if n else []
            for k in range(w):
                want = ref[k] if k < len(ref) else []
                assert get_chunk(list(range(n)), w, k) == want
                lo, hi = chunk_bounds(n, w, k)
                assert hi - lo == len(want)


def _worker(rank, world, port, n_items, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = chunk_bounds(n_items, world, rank)
    local = torch.arange(lo, hi).reshape(-1, 1).repeat(1, 3) * 10 + torch.arange(3)
    full = gather_token_ids(local, n_items)
    t = max_over_ranks(1.0 + rank, "cpu")
    if rank == 0:
        q.put((full.tolist(), t))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [8, 7])
def test_replica_sharding_gloo_world2(n_items):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + n_items
    ps = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in ps:
        p.start()
    full, t = q.get(timeout=120)
    for p in ps:
        p.join(60)
    want = (torch.arange(n_items).reshape(-1, 1).repeat(1, 3) * 10 + torch.arange(3)).tolist()
    assert full == want and t == 2.0


def test_synthetic_inputs_are_deterministic():
    cfg = preset("toy")
    a, pa = synthetic_inputs(cfg, 2, 12, seed=1)
    b, pb = synthetic_inputs(cfg, 2, 12, seed=1)
    assert torch.equal(a, b) and torch.equal(pa, pb) and (a[:, 5] == -200).all() and (a[:, 0] == 1).all()
    d = forced_text_decisions(11, 2, keep_rate=[0.3, 0.7], seed=3)
    assert d.shape == (11, 2) and d.dtype == torch.bool


def test_load_pretrained_model_roundtrip(tmp_path):
    """reference signature + return tuple (llava/model/dynamic_llava_builder.py:35-45, :249) on a synthetic HF-format dir"""
    import json
    from safetensors.torch import save_file
    from llava.model.dynamic_llava_builder import load_pretrained_model
    cfg = preset("toy")
    cfg["sparse_config"]["use_output_text_predictor"] = True
    sd = {k: v.to(torch.bfloat16).contiguous() for k, v in numpy_state_dict(cfg, 0).items()}
    tower = tmp_path / "clip"
    tower.mkdir()
    json.dump(dict(cfg["vision"]), open(tower / "config.json", "w"))
    clip_sd = {k[len("model.vision_tower.vision_tower."):]: v for k, v in sd.items() if "vision_tower" in k}
    save_file(clip_sd, str(tower / "model.safetensors"))
    ck = tmp_path / "dynamic-llava-toy"
    ck.mkdir()
    llm_sd = {k: v for k, v in sd.items() if "vision_tower" not in k}
    half = len(llm_sd) // 2
    items = list(llm_sd.items())
    save_file(dict(items[:half]), str(ck / "model-00001-of-00002.safetensors"))
    save_file(dict(items[half:]), str(ck / "model-00002-of-00002.safetensors"))
    cj = {k: v for k, v in cfg.items() if k != "vision"}
    cj.update(mm_vision_tower=str(tower), mm_projector_type="mlp2x_gelu", model_type="dynamic_llava_llama")
    json.dump(cj, open(ck / "config.json", "w"))
    with pytest.warns(UserWarning, match="tokenizer"):
        tok, model, proc, ctx = load_pretrained_model(str(ck), None, "dynamic-llava-toy", device="cpu", torch_dtype=torch.bfloat16)
    assert tok is None and ctx == 2048 and proc is not None
    got = model.state_dict()
    assert set(got) == set(sd) and all(torch.equal(got[k], sd[k]) for k in sd)
    with pytest.warns(UserWarning, match="tokenizer"):  # the reference's default dtype (dynamic_llava_builder.py:62)
        m16 = load_pretrained_model(str(ck), None, "dynamic-llava-toy", device="cpu")[1]
    assert m16.dtype == torch.float16 and all(torch.equal(v, sd[k].to(torch.float16)) for k, v in m16.state_dict().items())
    assert model.config.sparse_config["sparse_layer"] == 2 and model.get_vision_tower().is_loaded
    with pytest.raises(NotImplementedError):
        load_pretrained_model(str(ck), None, "dynamic-llava-toy", load_4bit=True)
    with pytest.raises(NotImplementedError):
        load_pretrained_model(str(ck), "base", "dynamic-llava-lora")


class _ToyTok:
    """whitespace tokenizer that prepends BOS like Llama's sentencepiece tokenizer does"""
    bos_token_id = 1

    def __call__(self, text):
        class R:
            pass
        r = R()
        r.input_ids = [1] + [3 + (sum(map(ord, w)) % 500) for w in text.split()]
        return r


def test_host_prompt_and_image_prep():
    """the eval drivers' host prep (llava/dynamic_eval/model_vqa_loader.py:60-80): conv template -> tokenizer_image_token -> process_images"""
    from PIL import Image
    from transformers import CLIPImageProcessor
    from llava.conversation import conv_templates
    from llava.constants import DEFAULT_IMAGE_TOKEN, IMAGE_TOKEN_INDEX
    from llava.mm_utils import expand2square, get_model_name_from_path, process_images, tokenizer_image_token
    conv = conv_templates["vicuna_v1"].copy()
    conv.append_message(conv.roles[0], DEFAULT_IMAGE_TOKEN + "\nWhat is shown?")
    conv.append_message(conv.roles[1], None)
    prompt = conv.get_prompt()
    assert prompt.endswith("USER: <image>\nWhat is shown? ASSISTANT:") and prompt.startswith("A chat between")
    assert conv_templates["vicuna_v1"].messages == []  # copy() does not leak turns into the template
    conv.messages[-1][1] = "A cat."
    assert conv.get_prompt().endswith("ASSISTANT: A cat.</s>")
    ids = tokenizer_image_token(prompt, _ToyTok(), IMAGE_TOKEN_INDEX, return_tensors="pt")
    assert ids.dtype == torch.long and ids[0] == 1 and (ids == 1).sum() == 1 and (ids == IMAGE_TOKEN_INDEX).sum() == 1
    left, right = prompt.split(DEFAULT_IMAGE_TOKEN)
    assert int((ids == IMAGE_TOKEN_INDEX).nonzero()[0]) == len(left.split()) + 1
    assert len(ids) == 1 + len(left.split()) + 1 + len(right.split())
    assert tokenizer_image_token("a <image> b <image> c", _ToyTok()).count(IMAGE_TOKEN_INDEX) == 2
    assert get_model_name_from_path("/x/dynamic-llava-7b/checkpoint-100/") == "dynamic-llava-7b_checkpoint-100"
    assert get_model_name_from_path("results/dynamic-llava-7b") == "dynamic-llava-7b"
    proc = CLIPImageProcessor(size={"shortest_edge": 56}, crop_size={"height": 56, "width": 56})
    img = Image.new("RGB", (80, 40), (255, 0, 0))
    sq = expand2square(img, (0, 0, 255))
    assert sq.size == (80, 80) and sq.getpixel((0, 0)) == (0, 0, 255) and sq.getpixel((40, 40)) == (255, 0, 0)

    class Cfg:
        image_aspect_ratio = "pad"
    px = process_images([img, img], proc, Cfg())
    assert px.shape == (2, 3, 56, 56)
    Cfg.image_aspect_ratio = None
    assert process_images([img], proc, Cfg()).shape == (1, 3, 56, 56)
    Cfg.image_aspect_ratio = "anyres"
    with pytest.raises(NotImplementedError):
        process_images([img], proc, Cfg())


class _FakeTok:
    """whitespace tokenizer: word -> id through a fixed table; bos = 1"""
    bos_token_id = 1
    table = {"</s>": 2, "###": 7, "Human": 8, "stop": 9, "now": 10}

    def __call__(self, text):
        class R:
            pass
        r = R()
        r.input_ids = [1] + [self.table.get(w, 99) for w in text.split()]
        return r

    def batch_decode(self, ids, skip_special_tokens=True):
        inv = {v: k for k, v in self.table.items()}
        return [" ".join(inv.get(int(t), "x") for t in row) for row in ids]


def test_keywords_stopping_criteria():
    """mm_utils.py:239-282: id-suffix match first, decoded-tail substring match second, all rows must be finished"""
    import torch
    from llava.mm_utils import KeywordsStoppingCriteria
    tok = _FakeTok()
    prompt = torch.zeros(1, 0, dtype=torch.long)
    c = KeywordsStoppingCriteria(["stop now", "###"], tok, prompt)
    assert [k.tolist() for k in c.keyword_ids] == [[9, 10], [7]] and c.max_keyword_len == 2 and c.start_len == 0
    assert not c(torch.tensor([[5, 9]]), None)
    assert c(torch.tensor([[5, 9, 10]]), None)       # suffix ids == keyword ids
    assert c(torch.tensor([[5, 6, 7]]), None)
    assert not c(torch.tensor([[5, 6, 7], [5, 6, 6]]), None)   # every row has to be finished
    assert c(torch.tensor([[5, 6, 7], [9, 10, 7]]), None)
    assert not c(torch.tensor([[7, 5, 6]]), None)    # the keyword left the inspected tail


def test_host_prep_matches_reference_golden():
    """conv templates, tokenizer_image_token, KeywordsStoppingCriteria, get_model_name_from_path, expand2square and process_images against
    the outputs of the REFERENCE's own llava/conversation.py and llava/mm_utils.py on the shared fixtures (oracle/cases.py ->
    oracle/gen_golden.py --only-hostprep -> tests/golden/host_prep_cases.pt), error cases included."""
    import numpy as np
    import torch
    from transformers import CLIPImageProcessor
    from llava.conversation import conv_templates
    from llava.mm_utils import KeywordsStoppingCriteria, expand2square, get_model_name_from_path, process_images, tokenizer_image_token
    from oracle import cases as C
    import llava
    assert not llava.__file__.startswith("/root/reference")  # the shim package of THIS repo is under test
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "host_prep_cases.pt"), weights_only=False)
    for (name, turns), want in zip(C.HOST_PROMPTS, g["prompts"]):
        conv = conv_templates[name].copy()
        for q, a in turns:
            conv.append_message(conv.roles[0], q)
            conv.append_message(conv.roles[1], a)
        if isinstance(want, dict):
            with pytest.raises(TypeError):
                conv.get_prompt()
        else:
            assert conv.get_prompt() == want, name
    assert len(g["token_ids"]) == 20
    for (no_bos, prompt), want in g["token_ids"].items():
        assert tokenizer_image_token(prompt, C.ToyTokenizer(no_bos), -200) == want, (no_bos, prompt)
    crit = KeywordsStoppingCriteria(C.STOP_KEYWORDS, C.ToyTokenizer(), torch.zeros(1, 0, dtype=torch.long))
    meta = g["stopping_meta"]
    assert [k.tolist() for k in crit.keyword_ids] == meta["keyword_ids"] and crit.max_keyword_len == meta["max_keyword_len"]
    assert crit.start_len == meta["start_len"]
    assert [bool(crit(torch.tensor(s), None)) for s in C.STOP_SEQUENCES] == g["stopping"]
    for p, want in g["model_names"].items():
        if isinstance(want, dict):
            with pytest.raises(IndexError):
                get_model_name_from_path(p)
        else:
            assert get_model_name_from_path(p) == want
    img = C.host_image()
    assert torch.equal(torch.from_numpy(np.array(expand2square(img, (10, 20, 30)))), g["square"])
    proc = CLIPImageProcessor(size={"shortest_edge": 56}, crop_size={"height": 56, "width": 56})

    class Cfg:
        image_aspect_ratio = "pad"
    assert torch.equal(process_images([img, img.transpose(2)], proc, Cfg()), g["pixels_pad"])
    Cfg.image_aspect_ratio = None
    assert torch.equal(process_images([img], proc, Cfg()), g["pixels_default"])


def test_state_dict_contract_and_constants_match_reference_golden():
    """every state_dict key AND shape of the reference model (toy dims; built by oracle/gen_golden.py from the unmodified reference
    classes) against this package's model, plus llava.constants — the checkpoint / import contract of SURVEY §8b"""
    import torch
    import llava.constants as C
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "prepare_cases_bf16.pt"), weights_only=False)
    cfg = DynamicLlavaConfig.from_preset("toy")
    cfg.sparse_config["use_output_text_predictor"] = True
    m = DynamicLlavaLlamaForCausalLM(cfg, device="cpu")
    m.get_vision_tower().load_model()
    mine = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    ref = {k: v for k, v in g["state_dict_shapes"].items() if "rotary_emb" not in k and not k.endswith("position_ids")}
    assert set(mine) == set(ref), (sorted(set(mine) ^ set(ref))[:8])
    assert mine == ref
    for k, v in g["constants"].items():
        assert getattr(C, k) == v, k


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs the reference tree (build container only)")
def test_reference_written_checkpoint_loads(tmp_path):
    """a checkpoint directory produced by the REFERENCE classes (config.json from its DynamicLlavaConfig, weights = its model's own
    state_dict in safetensors shards, CLIP tower directory; oracle/ref_save_checkpoint.py, run in a subprocess because the two `llava`
    packages share a name) loads through this package's load_pretrained_model with identical tensors and the right configuration"""
    import subprocess
    import sys
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "oracle", "ref_save_checkpoint.py"), str(tmp_path)], capture_output=True, text=True,
                       cwd=root, timeout=600)
    assert r.returncode == 0 and "saved" in r.stdout, r.stderr[-2000:]
    ck = str(tmp_path / "dynamic-llava-toy")
    cj = json.load(open(os.path.join(ck, "config.json")))
    assert cj["model_type"] == "dynamic_llava_llama" and "sparse_config" in cj
    from llava.model.dynamic_llava_builder import load_pretrained_model
    with pytest.warns(UserWarning, match="tokenizer"):
        tok, model, proc, ctx = load_pretrained_model(ck, None, "dynamic-llava-toy", device="cpu", torch_dtype=torch.bfloat16)
    cfg = preset("toy")
    cfg["sparse_config"]["use_output_text_predictor"] = True
    sd = {k: v.to(torch.bfloat16) for k, v in numpy_state_dict(cfg, 0).items()}
    got = model.state_dict()
    assert set(got) == set(sd) and all(torch.equal(got[k], sd[k]) for k in sd)
    assert model.config.rope_theta == 10000.0 and model.config.sparse_config == cfg["sparse_config"]
    assert model.config.mm_vision_select_layer == -2 and model.get_vision_tower().is_loaded and ctx == 2048 and proc is not None


def test_eval_driver_import_surface():
    """every `llava.*` import of the reference's llava/dynamic_eval/*.py drivers resolves against this repo's shim package
    (model_vqa_loader.py:1-20): utils.disable_torch_init, the builder, mm_utils, conversation, constants"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import torch\n"
            "from llava.utils import disable_torch_init\n"
            "from llava.model.dynamic_llava_builder import load_pretrained_model\n"
            "from llava.mm_utils import tokenizer_image_token, process_images, get_model_name_from_path, KeywordsStoppingCriteria\n"
            "from llava.conversation import conv_templates, SeparatorStyle\n"
            "from llava.constants import IMAGE_TOKEN_INDEX, DEFAULT_IMAGE_TOKEN, DEFAULT_IM_START_TOKEN, DEFAULT_IM_END_TOKEN, IMAGE_PLACEHOLDER\n"
            "from llava.model import DynamicLlavaLlamaForCausalLM, LlavaLlamaForCausalLM\n"
            "import llava; assert 'reference' not in llava.__file__\n"
            "disable_torch_init(); l = torch.nn.Linear(4, 4); assert l.weight.shape == (4, 4)\n"
            "print('ok')")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root, timeout=300)  # own process: it patches torch.nn
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-1500:]


def test_auto_classes_registered_and_from_pretrained(tmp_path):
    """the reference registers its classes with transformers' Auto factories at import time (dynamic_llava_llama.py:172-173) and its
    builder loads through `DynamicLlavaLlamaForCausalLM.from_pretrained(path, low_cpu_mem_usage=True, torch_dtype=float16, ...)`
    (dynamic_llava_builder.py:62, :191-193): both entry points exist here and return the same weights; bfloat16 (default) and float16 each have a
    kernel library, anything else is refused"""
    import json
    from safetensors.torch import save_file
    from transformers import AutoConfig, AutoModelForCausalLM
    import llava.model  # noqa: F401  (registration happens on import)
    from llava.model import DynamicLlavaLlamaForCausalLM
    cfg = preset("toy")
    cfg["sparse_config"]["use_output_text_predictor"] = True
    sd = {k: v.to(torch.bfloat16).contiguous() for k, v in numpy_state_dict(cfg, 0).items()}
    ck = tmp_path / "ck"
    ck.mkdir()
    cj = {k: v for k, v in cfg.items() if k != "vision"}
    cj.update(model_type="dynamic_llava_llama", mm_vision_tower=str(tmp_path / "clip"), mm_projector_type="mlp2x_gelu", architectures=["DynamicLlavaLlamaForCausalLM"])
    json.dump(cj, open(ck / "config.json", "w"))
    (tmp_path / "clip").mkdir()
    vc = dict(cfg["vision"], model_type="clip_vision_model")
    json.dump(vc, open(tmp_path / "clip" / "config.json", "w"))
    save_file(sd, str(ck / "model.safetensors"))
    hf_cfg = AutoConfig.from_pretrained(str(ck))
    assert hf_cfg.model_type == "dynamic_llava_llama" and type(hf_cfg).__name__ == "DynamicLlavaConfig"
    assert hf_cfg.sparse_config["sparse_layer"] == cfg["sparse_config"]["sparse_layer"]
    m1 = AutoModelForCausalLM.from_pretrained(str(ck), device_map="cpu")
    assert isinstance(m1, DynamicLlavaLlamaForCausalLM)
    m2 = DynamicLlavaLlamaForCausalLM.from_pretrained(str(ck), low_cpu_mem_usage=True, device_map="cpu", torch_dtype=torch.float16)
    for m, dt in ((m1, torch.bfloat16), (m2, torch.float16)):
        got = m.state_dict()
        assert set(got) == set(sd) and all(torch.equal(got[k], sd[k].to(dt)) for k in sd)
        assert m.dtype == dt and m.config.sparse_config == cfg["sparse_config"]
    with pytest.raises(NotImplementedError, match="bfloat16 or float16 only"):
        DynamicLlavaLlamaForCausalLM.from_pretrained(str(ck), device_map="cpu", torch_dtype=torch.float32)
