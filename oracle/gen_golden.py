"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):   python oracle/gen_golden.py
The GPU box has no /root/reference; tests there read the committed fixtures.

Shim: transformers 5.x dropped `is_torch_fx_available` and moved `rope_theta`/`rope_scaling`
into `rope_parameters`; nothing in the reference's files is edited.  Weights are NOT sampled by torch: they come
from `dynamic_llava_b200.synthetic.numpy_state_dict` (numpy PCG64) and are loaded strictly into the reference
model, so a fixture stores only seeds, inputs and outputs.

Forced decisions are injected without touching reference code: forward hooks on the predictor modules replace
their logits by +/-1 patterns that make the reference's own `argsort` / `logit0 > logit1` produce the wanted mask.
"""
from __future__ import annotations

import copy
import os
import sys
import tempfile

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import torch  # noqa: E402
import transformers.utils.import_utils as iu  # noqa: E402

iu.is_torch_fx_available = lambda: False  # dynamic_modeling_llama.py:53

from transformers import CLIPImageProcessor, CLIPVisionConfig, CLIPVisionModel  # noqa: E402
from llava.model.language_model.dynamic_llava_llama import (  # noqa: E402
    DynamicLlavaConfig,
    DynamicLlavaLlamaForCausalLM,
)

from dynamic_llava_b200.presets import SPARSE_DEFAULTS, preset  # noqa: E402
from dynamic_llava_b200.synthetic import forced_text_decisions, numpy_state_dict, synthetic_inputs  # noqa: E402
from oracle.cases import PREPARE_CASES, prepare_case_inputs  # noqa: E402


class ShimCfg(DynamicLlavaConfig):
    rope_theta = 10000.0  # read at dynamic_modeling_llama.py:368
    rope_scaling = None  # read at :396


def build_reference(cfg: dict, sd: dict, dtype, scratch: str):
    vc = cfg["vision"]
    clip_dir = os.path.join(scratch, "clip")
    ccfg = CLIPVisionConfig(hidden_size=vc["hidden_size"], intermediate_size=vc["intermediate_size"],
                            num_hidden_layers=vc["num_hidden_layers"], num_attention_heads=vc["num_attention_heads"],
                            image_size=vc["image_size"], patch_size=vc["patch_size"], projection_dim=64,
                            layer_norm_eps=vc["layer_norm_eps"], hidden_act="quick_gelu")
    CLIPVisionModel(ccfg).save_pretrained(clip_dir)
    CLIPImageProcessor(size={"shortest_edge": vc["image_size"]},
                       crop_size={"height": vc["image_size"], "width": vc["image_size"]}).save_pretrained(clip_dir)
    hcfg = ShimCfg(hidden_size=cfg["hidden_size"], intermediate_size=cfg["intermediate_size"],
                   num_hidden_layers=cfg["num_hidden_layers"], num_attention_heads=cfg["num_attention_heads"],
                   num_key_value_heads=cfg["num_attention_heads"], vocab_size=cfg["vocab_size"],
                   rms_norm_eps=cfg["rms_norm_eps"], max_position_embeddings=cfg["max_position_embeddings"],
                   pretraining_tp=1, sparse_config=copy.deepcopy(cfg["sparse_config"]),
                   mm_vision_tower=clip_dir, mm_projector_type="mlp2x_gelu", mm_hidden_size=cfg["mm_hidden_size"],
                   mm_vision_select_layer=cfg["mm_vision_select_layer"], mm_vision_select_feature="patch",
                   mm_patch_merge_type="flat", use_cache=True, pad_token_id=0, bos_token_id=1, eos_token_id=2)
    model = DynamicLlavaLlamaForCausalLM(hcfg).eval()
    model.get_vision_tower().load_model()
    missing, unexpected = model.load_state_dict(sd, strict=False)
    # rotary inv_freq / cos / sin are non-persistent buffers; nothing else may be missing
    assert not unexpected, unexpected
    assert all("rotary_emb" in m or "position_ids" in m for m in missing), missing
    model.to(dtype)
    return model


def force_hooks(model, keep_index=None, decisions=None):
    """decisions: iterator yielding bool[B] per decode call."""
    handles = []
    if keep_index is not None:
        def vis_hook(mod, args, out):
            z = torch.zeros_like(out)
            z[..., 0] = -1.0
            z[..., 0].scatter_(1, keep_index, 1.0)
            return z
        handles.append(model.model.image_score_predictor.register_forward_hook(vis_hook))
    if decisions is not None:
        it = iter(decisions)

        def txt_hook(mod, args, out):
            d = next(it)
            z = torch.zeros_like(out)
            z[..., 0] = torch.where(d.reshape(-1, 1), 1.0, -1.0).to(out.dtype)
            return z
        handles.append(model.model.output_text_score_predictor.register_forward_hook(txt_hook))
    return handles


@torch.no_grad()
def run_reference(model, input_ids, pixels, n_new, forced_dec=None, forced_keep=None, capture_layers=False):
    """Loop shape of llava/dynamic_eval/model_lvis_for_meteor.py:145-174 (batched like
    bench_test/dynamic_llava_long_text_mem.py:310-338)."""
    cap = {}
    hs = []
    hs.append(model.model.image_score_predictor.register_forward_hook(
        lambda m, a, o: cap.setdefault("vision_logits", o.detach().clone())))
    if hasattr(model.model, "output_text_score_predictor"):
        cap["text_logits"] = []
        hs.append(model.model.output_text_score_predictor.register_forward_hook(
            lambda m, a, o: cap["text_logits"].append(o.detach().clone())))
    if capture_layers:
        cap["layer_inputs"] = []
        for layer in model.model.layers:
            hs.append(layer.register_forward_pre_hook(
                lambda m, a, kw: cap["layer_inputs"].append((a[0] if a else kw["hidden_states"]).detach().clone()),
                with_kwargs=True))
    hs += force_hooks(model, forced_keep, None if forced_dec is None else list(forced_dec))
    ids, past, images = input_ids, None, pixels
    toks, lgs, tls = [], [], []
    prefill_layers = None
    for t in range(n_new):
        out = model(ids, images=images, past_key_values=past)
        if t == 0 and capture_layers:
            prefill_layers = cap["layer_inputs"]
            cap["layer_inputs"] = []
        images = None
        last = out.logits[:, -1, :]
        nxt = torch.argmax(last, dim=-1)
        toks.append(nxt)
        lgs.append(last.clone())
        past = out.past_key_values
        tls.append(torch.stack([x.clone() for x in past[1]], 0))
        ids = nxt.reshape(-1, 1)
    for h in hs:
        h.remove()
    res = {
        "tokens": torch.stack(toks, 1),
        "logits": torch.stack(lgs, 0),
        "true_len_hist": torch.stack(tls, 0),
        "vision_logits": cap["vision_logits"],
        "kv_shapes": torch.tensor([k.shape[-2] for (k, v) in past[0]]),
        "k_last": [k.clone() for (k, v) in past[0]],
        "v_last": [v.clone() for (k, v) in past[0]],
    }
    if cap.get("text_logits"):
        res["text_logits"] = torch.stack(cap["text_logits"], 0)
    if prefill_layers is not None:
        res["prefill_layer_inputs"] = prefill_layers
    return res


FP16_CASES = [
    # the reference's eval drivers load torch_dtype=float16 (llava/dynamic_eval/model_vqa_loader.py:82 via builder.py:26)
    ("toy_fp16_b1_natural", torch.float16, 1, 12, 8, False, True),
    ("toy_fp16_b2_forced", torch.float16, 2, 12, 12, True, True),
]


def main(cases=None):
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    torch.set_num_threads(8)
    cfg = preset("toy")
    cases = cases or [
        # name, dtype, B, T, n_new, forced decisions?, output predictor on?
        ("toy_fp32_b1_natural", torch.float32, 1, 12, 8, False, True),
        ("toy_fp32_b2_natural", torch.float32, 2, 12, 8, False, True),
        ("toy_bf16_b1_natural", torch.bfloat16, 1, 12, 8, False, True),
        ("toy_bf16_b1_forced", torch.bfloat16, 1, 12, 12, True, True),
        ("toy_bf16_b2_forced", torch.bfloat16, 2, 12, 12, True, True),
        ("toy_bf16_b3_nodrop", torch.bfloat16, 3, 16, 6, False, False),
    ]
    with tempfile.TemporaryDirectory(dir=os.path.join(ROOT, "oracle")) as scratch:
        for name, dtype, B, T, n_new, forced, out_pred in cases:
            c = copy.deepcopy(cfg)
            sd = numpy_state_dict(c, seed=0)
            c["sparse_config"]["use_output_text_predictor"] = True  # module must exist to load every key
            model = build_reference(c, sd, dtype, scratch)
            model.model.config.sparse_config["use_output_text_predictor"] = out_pred  # runtime flag, like the bench script
            ids, pix = synthetic_inputs(c, B, T, seed=1, dtype=dtype)
            fd = forced_text_decisions(n_new - 1, B, keep_rate=[0.3, 0.7, 0.5][:B] if B > 1 else 0.5, seed=3) if forced else None
            res = run_reference(model, ids, pix, n_new, forced_dec=fd, capture_layers=(name == "toy_bf16_b1_natural"))
            res.update(name=name, dtype=str(dtype), B=B, T=T, n_new=n_new, weight_seed=0, input_seed=1,
                       forced_decisions=fd, use_output_text_predictor=out_pred, preset="toy",
                       torch_version=torch.__version__)
            if name not in ("toy_bf16_b2_forced", "toy_fp16_b2_forced"):  # keep the repo small: KV dumps only where a test reads them
                res.pop("k_last"), res.pop("v_last")
            torch.save(res, os.path.join(out_dir, name + ".pt"))
            print(name, "tokens", res["tokens"].tolist(), "kv", res["kv_shapes"].tolist(),
                  "true_len", res["true_len_hist"][-1].tolist())


@torch.no_grad()
def multiround_case(out_dir, scratch):
    """second user turn: q_len > 1 over an existing ragged cache (llava/dynamic_eval/model_lvis_multi_round_for_meteor.py:150-185
    with use_instruct_predictor=False, README.md:110): prefill -> 3 greedy steps -> 5-token continuation -> 3 greedy steps."""
    cfg = preset("toy")
    sd = numpy_state_dict(cfg, seed=0)
    cfg["sparse_config"]["use_output_text_predictor"] = True
    model = build_reference(cfg, sd, torch.bfloat16, scratch)
    ids, pix = synthetic_inputs(cfg, 1, 12, seed=1, dtype=torch.bfloat16)
    cont, _ = synthetic_inputs(cfg, 1, 12, seed=5, dtype=torch.bfloat16)
    cont = cont[:, 6:11].contiguous()  # 5 ordinary token ids
    logits, lens, toks = [], [], []
    out = model(ids, images=pix, past_key_values=None)
    for t in range(3):
        logits.append(out.logits[:, -1].clone()); lens.append(torch.stack([x.clone() for x in out.past_key_values[1]], 0))
        nxt = out.logits[:, -1].argmax(-1).reshape(1, 1); toks.append(int(nxt))
        out = model(nxt, images=None, past_key_values=out.past_key_values)
    logits.append(out.logits[:, -1].clone()); lens.append(torch.stack([x.clone() for x in out.past_key_values[1]], 0))
    out = model(cont, images=None, past_key_values=out.past_key_values)
    cont_logits = out.logits.clone()
    lens.append(torch.stack([x.clone() for x in out.past_key_values[1]], 0))
    for t in range(3):
        nxt = out.logits[:, -1].argmax(-1).reshape(1, 1); toks.append(int(nxt))
        out = model(nxt, images=None, past_key_values=out.past_key_values)
        logits.append(out.logits[:, -1].clone()); lens.append(torch.stack([x.clone() for x in out.past_key_values[1]], 0))
    res = dict(name="toy_bf16_b1_multiround", dtype="torch.bfloat16", B=1, T=12, weight_seed=0, input_seed=1, cont_ids=cont, tokens=toks,
               step_logits=torch.stack(logits, 0), cont_logits=cont_logits, true_len_hist=torch.stack(lens, 0), preset="toy",
               kv_shapes=torch.tensor([k.shape[-2] for (k, v) in out.past_key_values[0]]), torch_version=torch.__version__)
    torch.save(res, os.path.join(out_dir, "multiround_bf16_b1.pt"))
    print("multiround tokens", toks, "kv", res["kv_shapes"].tolist())


@torch.no_grad()
def instruct_case(out_dir, scratch):
    """`use_instruct_predictor=True`: the prefill drops instruction tokens of the last instruct span
    (dynamic_modeling_llama.py:2261-2375, B == 1 asserted) and a new user turn persists K/V only where the instruct
    predictor says so, last token always (:2506-2521 + cache_utils.py:165-241).  Two variants in one fixture: the
    predictor's own decisions ("natural") and +/-1 patterns injected through a forward hook ("forced")."""
    cfg = preset("toy")
    cfg["sparse_config"]["use_output_text_predictor"] = True
    cfg["sparse_config"]["use_instruct_predictor"] = True
    sd = numpy_state_dict(cfg, seed=0)
    model = build_reference(cfg, sd, torch.bfloat16, scratch)
    T = 20
    ids, pix = synthetic_inputs(cfg, 1, T, seed=1, dtype=torch.bfloat16)
    cont, _ = synthetic_inputs(cfg, 1, 12, seed=5, dtype=torch.bfloat16)
    cont = cont[:, 6:11].contiguous()
    n_span = T - 1 - 5 - 1  # instruct tokens minus the last one (never dropped)
    rng = torch.Generator().manual_seed(11)
    forced = [torch.rand(n_span, generator=rng) < 0.6, torch.tensor([True, False, True, False, False])]
    out_all = {}
    for variant in ("natural", "forced"):
        seen, it = [], iter(forced)

        def hook(mod, args, out):
            seen.append(out.detach().clone())
            if variant == "natural":
                return out
            d = next(it)
            z = torch.zeros_like(out)
            z[..., 0] = torch.where(d.reshape(out.shape[:-1]), 1.0, -1.0).to(out.dtype)
            return z
        h = model.model.instruct_score_predictor.register_forward_hook(hook)
        logits, lens, toks = [], [], []
        snap = lambda o: torch.stack([x.clone() for x in o.past_key_values[1]], 0)
        out = model(ids, images=pix, past_key_values=None)
        for t in range(3):
            logits.append(out.logits[:, -1].clone()); lens.append(snap(out))
            nxt = out.logits[:, -1].argmax(-1).reshape(1, 1); toks.append(int(nxt))
            out = model(nxt, images=None, past_key_values=out.past_key_values)
        logits.append(out.logits[:, -1].clone()); lens.append(snap(out))
        out = model(cont, images=None, past_key_values=out.past_key_values)
        cont_logits = out.logits.clone()
        lens.append(snap(out))
        for t in range(3):
            nxt = out.logits[:, -1].argmax(-1).reshape(1, 1); toks.append(int(nxt))
            out = model(nxt, images=None, past_key_values=out.past_key_values)
            logits.append(out.logits[:, -1].clone()); lens.append(snap(out))
        h.remove()
        out_all[variant] = dict(tokens=toks, step_logits=torch.stack(logits, 0), cont_logits=cont_logits,
                                true_len_hist=torch.stack(lens, 0), instruct_logits_prefill=seen[0], instruct_logits_turn=seen[1],
                                kv_shapes=torch.tensor([k.shape[-2] for (k, v) in out.past_key_values[0]]))
        print("instruct", variant, "tokens", toks, "kv", out_all[variant]["kv_shapes"].tolist(),
              "prefill keep", (seen[0][0, :, 0] > seen[0][0, :, 1]).int().tolist() if variant == "natural" else forced[0].int().tolist())
    res = dict(name="toy_bf16_b1_instruct", dtype="torch.bfloat16", B=1, T=T, weight_seed=0, input_seed=1, cont_ids=cont,
               forced_prefill=forced[0], forced_turn=forced[1], preset="toy", torch_version=torch.__version__, **out_all)
    torch.save(res, os.path.join(out_dir, "instruct_bf16_b1.pt"))


def nocache_forced_logits(n_max: int, B: int, seed: int = 13):
    """Hook patterns for the no-KV-cache fixture: for every output-token count n (the hook is keyed by its input length, not by call
    order) bf16-exact logits [B, n, 2] whose channel-0 scores are distinct within a sample (so `argsort` has no ties) and whose
    `logit0 > logit1` counts differ between the samples (so `num_keep_node = max over the batch` is exercised)."""
    rng = torch.Generator().manual_seed(seed)
    out = {}
    for n in range(1, n_max + 1):
        z = torch.zeros(B, n, 2)
        for b in range(B):
            z[b, :, 0] = (torch.randperm(n, generator=rng).float() - (n - 1) / 2.0 + (0.75 if b % 2 else -0.75)) * 0.25
        out[n] = z
    return out


@torch.no_grad()
def nocache_case(out_dir, scratch):
    """`use_cache=False` output sparsification (dynamic_modeling_llama.py:2393-2504): the loop of
    llava/dynamic_eval/bench_test/dynamic_llava_long_text_time_with_no_cache.py:336-349 — the whole id sequence (prompt + tokens so
    far) and the images are re-fed on every call, `model.model.answer_indice` stays at the first call's prompt end.  Greedy tokens take
    the place of the script's label ids.  Variants: the predictor's own logits ("natural") and hook-injected ones ("forced")."""
    cfg = preset("toy")
    cfg["sparse_config"]["use_output_text_predictor"] = True
    sd = numpy_state_dict(cfg, seed=0)
    model = build_reference(cfg, sd, torch.bfloat16, scratch)
    B, T, n_calls = 2, 12, 8
    ids0, pix = synthetic_inputs(cfg, B, T, seed=1, dtype=torch.bfloat16)
    forced = nocache_forced_logits(n_calls, B)
    SL = cfg["sparse_config"]["sparse_layer"]
    out_all = {}
    for variant in ("natural", "forced"):
        model.model.answer_indice = None
        seen, sl_in = [], []

        def hook(mod, args, out):
            seen.append(out.detach().clone())
            n = args[0].shape[1]
            if variant == "natural" or n == 0:
                return out
            return forced[n].to(out.dtype).reshape(out.shape)
        hs = [model.model.output_text_score_predictor.register_forward_hook(hook),
              model.model.layers[SL].register_forward_pre_hook(
                  lambda m, a, kw: sl_in.append((a[0] if a else kw["hidden_states"]).detach().clone()), with_kwargs=True)]
        total = ids0.clone()
        last_logits, lens, full = [], [], None
        for t in range(n_calls):
            out = model(total, images=pix, past_key_values=None, input_embeds_indices=None, use_cache=False)
            assert out.past_key_values is None
            last_logits.append(out.logits[:, -1].clone())
            lens.append(out.logits.shape[1])
            full = out.logits.clone()
            total = torch.cat([total, out.logits[:, -1].argmax(-1).reshape(B, 1)], dim=1)
        for h in hs:
            h.remove()
        out_all[variant] = dict(tokens=total[:, T:].clone(), last_logits=torch.stack(last_logits, 0), seq_lens=lens, final_logits=full,
                                text_logits=seen, sparse_layer_inputs=sl_in, answer_indice=int(model.model.answer_indice))
        print("nocache", variant, "tokens", total[:, T:].tolist(), "lens", lens, "answer_indice", model.model.answer_indice)
    # branch precedence (dynamic_modeling_llama.py:2261 before :2393): with the instruct predictor on, a use_cache=False call drops
    # instruction tokens and never reaches the output-thinning branch (B == 1 asserted by the reference)
    cfg_i = preset("toy")
    cfg_i["sparse_config"]["use_output_text_predictor"] = True
    cfg_i["sparse_config"]["use_instruct_predictor"] = True
    model_i = build_reference(cfg_i, numpy_state_dict(cfg_i, seed=0), torch.bfloat16, scratch)
    ids_i, pix_i = synthetic_inputs(cfg_i, 1, 20, seed=1, dtype=torch.bfloat16)
    tot = torch.cat([ids_i, torch.tensor([[7, 9, 11]])], dim=1)
    out_i = model_i(tot, images=pix_i, past_key_values=None, input_embeds_indices=None, use_cache=False)
    out_all["instruct_precedence"] = dict(ids=tot, logits=out_i.logits.clone(), answer_indice=model_i.model.answer_indice)
    print("nocache + instruct predictor: logits", tuple(out_i.logits.shape), "answer_indice", model_i.model.answer_indice)
    res = dict(name="toy_bf16_b2_nocache", dtype="torch.bfloat16", B=B, T=T, n_calls=n_calls, weight_seed=0, input_seed=1,
               forced_logits=forced, preset="toy", torch_version=torch.__version__, **out_all)
    torch.save(res, os.path.join(out_dir, "nocache_bf16_b2.pt"))


@torch.no_grad()
def prepare_cases(out_dir, scratch):
    """prepare_inputs_labels_for_multimodal alone (dynamic_llava_arch.py:169-601): spliced embeddings + the index dicts for the
    edge cases the greedy fixtures do not reach."""
    cfg = preset("toy")
    cfg["sparse_config"]["use_output_text_predictor"] = True
    sd = numpy_state_dict(cfg, seed=0)
    model = build_reference(cfg, sd, torch.bfloat16, scratch)
    res = {"preset": "toy", "weight_seed": 0, "torch_version": str(torch.__version__), "cases": {}}
    # the checkpoint contract: every key and shape of the reference model's state_dict (toy dims, both text predictors off/on as built)
    res["state_dict_shapes"] = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    import llava.constants as RC
    res["constants"] = {k: getattr(RC, k) for k in dir(RC) if k.isupper()}
    for name in PREPARE_CASES:
        ids, pix, kw = prepare_case_inputs(cfg, name)
        model.config.tokenizer_model_max_length = kw.get("max_len")
        model.config.tokenizer_padding_side = kw.get("padding_side", "right")
        try:
            (_, pos, mask, _, embeds, _), (indices,) = model.prepare_inputs_labels_for_multimodal(ids, None, None, None, None, pix)
        except IndexError as e:  # left padding + a row without a dict: the reference indexes input_embeds_indices[i] past its end (:556-566)
            res["cases"][name] = {"raises": "IndexError"}
            print("prepare", name, "-> reference raises IndexError:", e)
            continue
        assert pos is None and mask is None
        res["cases"][name] = {"embeds": embeds.clone(), "indices": copy.deepcopy(indices)}
        print("prepare", name, tuple(embeds.shape), indices)
    torch.save(res, os.path.join(out_dir, "prepare_cases_bf16.pt"))


def host_prep_cases(out_dir):
    """The reference's own host-side helpers (llava/conversation.py, llava/mm_utils.py) run on the shared fixtures of oracle/cases.py:
    prompt strings, tokenizer_image_token ids (BOS and no-BOS tokenizers), KeywordsStoppingCriteria decisions, get_model_name_from_path,
    expand2square + process_images ("pad" and default) pixel tensors."""
    from llava.conversation import conv_templates
    from llava.mm_utils import KeywordsStoppingCriteria, expand2square, get_model_name_from_path, process_images, tokenizer_image_token
    from oracle import cases as C
    import llava
    assert llava.__file__.startswith("/root/reference"), llava.__file__
    res = {"prompts": [], "token_ids": {}, "stopping": [], "model_names": {}, "torch_version": str(torch.__version__)}
    for name, turns in C.HOST_PROMPTS:
        conv = conv_templates[name].copy()
        for q, a in turns:
            conv.append_message(conv.roles[0], q)
            conv.append_message(conv.roles[1], a)
        try:
            res["prompts"].append(conv.get_prompt())
        except TypeError:  # "plain" with an answer: sep2 is None in the reference's template (conversation.py:106)
            res["prompts"].append({"raises": "TypeError"})
    for no_bos in (False, True):
        tok = C.ToyTokenizer(no_bos)
        for p in C.TOKENIZE_PROMPTS + [q for q in res["prompts"] if isinstance(q, str)]:
            res["token_ids"][(no_bos, p)] = tokenizer_image_token(p, tok, -200)
    tok = C.ToyTokenizer()
    crit = KeywordsStoppingCriteria(C.STOP_KEYWORDS, tok, torch.zeros(1, 0, dtype=torch.long))
    res["stopping_meta"] = {"keyword_ids": [k.tolist() for k in crit.keyword_ids], "max_keyword_len": crit.max_keyword_len, "start_len": crit.start_len}
    for seq in C.STOP_SEQUENCES:
        res["stopping"].append(bool(crit(torch.tensor(seq), None)))
    for p in C.MODEL_PATHS:
        try:
            res["model_names"][p] = get_model_name_from_path(p)
        except IndexError:  # a bare "checkpoint-N" has no parent component (mm_utils.py:233-234)
            res["model_names"][p] = {"raises": "IndexError"}
    proc = CLIPImageProcessor(size={"shortest_edge": 56}, crop_size={"height": 56, "width": 56})
    img = C.host_image()
    sq = expand2square(img, (10, 20, 30))
    res["square"] = torch.from_numpy(__import__("numpy").array(sq))

    class Cfg:
        image_aspect_ratio = "pad"
    res["pixels_pad"] = process_images([img, img.transpose(2)], proc, Cfg())  # 2 = ROTATE_90: a tall image too
    Cfg.image_aspect_ratio = None
    res["pixels_default"] = process_images([img], proc, Cfg())
    torch.save(res, os.path.join(out_dir, "host_prep_cases.pt"))
    print("host prep:", len(res["prompts"]), "prompts,", len(res["token_ids"]), "tokenisations,", res["stopping"], res["model_names"])


@torch.no_grad()
def keep_rate_cases(out_dir, scratch):
    """`vision_keep_rate` mutated on the live sparse_config between calls (bench_test/dynamic_llava_image_time_and_mem.py:63-65) at the
    extremes the GPU tests visit: 1.0 (k = n_img: everything kept, still routed through the gather), 0.07 (k = 1), 0.0 (k = 0: the image
    span vanishes), plus 0.3 (k = 4): prefill logits, keep index, KV lengths and one decode step."""
    cfg = preset("toy")
    cfg["sparse_config"]["use_output_text_predictor"] = True
    sd = numpy_state_dict(cfg, seed=0)
    model = build_reference(cfg, sd, torch.bfloat16, scratch)
    model.model.config.sparse_config["use_output_text_predictor"] = False
    ids, pix = synthetic_inputs(cfg, 2, 12, seed=1, dtype=torch.bfloat16)
    res = {"preset": "toy", "B": 2, "T": 12, "weight_seed": 0, "input_seed": 1, "torch_version": str(torch.__version__), "rates": {}}
    for rate in (1.0, 0.3, 0.07, 0.0):
        model.model.config.sparse_config["vision_keep_rate"] = rate
        try:
            out = model(ids, images=pix, past_key_values=None)
            nxt = out.logits[:, -1].argmax(-1).reshape(-1, 1)
            out2 = model(nxt, images=None, past_key_values=out.past_key_values)
            res["rates"][rate] = dict(logits=out.logits.clone(), step_logits=out2.logits.clone(),
                                      kv_shapes=[k.shape[-2] for (k, v) in out2.past_key_values[0]])
            print("keep rate", rate, tuple(out.logits.shape), res["rates"][rate]["kv_shapes"])
        except Exception as e:  # noqa
            res["rates"][rate] = {"raises": type(e).__name__}
            print("keep rate", rate, "-> reference raises", type(e).__name__, str(e)[:120])
    torch.save(res, os.path.join(out_dir, "keep_rates_bf16_b2.pt"))


@torch.no_grad()
def multiround_b2_case(out_dir, scratch):
    """the batched (B = 2) form of the second user turn: forced per-sample decisions (0.3 / 0.7) make the cached lengths diverge, so the
    reference's B > 1 cache path runs — zero-padded stack per layer (cache_utils.py:165-241), then the 5-token turn appended BEHIND the
    padding (plain `torch.cat`, :242-250) and unmasked attention over it.  Pins the oracle's batched-quirk mode there."""
    cfg = preset("toy")
    cfg["sparse_config"]["use_output_text_predictor"] = True
    sd = numpy_state_dict(cfg, seed=0)
    model = build_reference(cfg, sd, torch.bfloat16, scratch)
    B = 2
    ids, pix = synthetic_inputs(cfg, B, 12, seed=1, dtype=torch.bfloat16)
    cont, _ = synthetic_inputs(cfg, B, 12, seed=5, dtype=torch.bfloat16)
    cont = cont[:, 6:11].contiguous()
    fd = forced_text_decisions(8, B, keep_rate=[0.3, 0.7], seed=3)
    hs = force_hooks(model, None, list(fd))
    snap = lambda o: torch.stack([x.clone() for x in o.past_key_values[1]], 0)
    logits, lens, toks = [], [], []
    out = model(ids, images=pix, past_key_values=None)
    for t in range(4):
        logits.append(out.logits[:, -1].clone()); lens.append(snap(out))
        nxt = out.logits[:, -1].argmax(-1).reshape(B, 1); toks.append(nxt.reshape(-1).clone())
        out = model(nxt, images=None, past_key_values=out.past_key_values)
    logits.append(out.logits[:, -1].clone()); lens.append(snap(out))
    out = model(cont, images=None, past_key_values=out.past_key_values)
    cont_logits = out.logits.clone()
    lens.append(snap(out))
    for t in range(3):
        nxt = out.logits[:, -1].argmax(-1).reshape(B, 1); toks.append(nxt.reshape(-1).clone())
        out = model(nxt, images=None, past_key_values=out.past_key_values)
        logits.append(out.logits[:, -1].clone()); lens.append(snap(out))
    for h in hs:
        h.remove()
    res = dict(name="toy_bf16_b2_multiround", dtype="torch.bfloat16", B=B, T=12, weight_seed=0, input_seed=1, cont_ids=cont,
               forced_decisions=fd, tokens=torch.stack(toks, 1), step_logits=torch.stack(logits, 0), cont_logits=cont_logits,
               true_len_hist=torch.stack(lens, 0), preset="toy", torch_version=str(torch.__version__),
               kv_shapes=torch.tensor([k.shape[-2] for (k, v) in out.past_key_values[0]]))
    torch.save(res, os.path.join(out_dir, "multiround_bf16_b2.pt"))
    print("multiround b2 tokens", res["tokens"].tolist(), "kv", res["kv_shapes"].tolist(), "true_len", res["true_len_hist"][-1].tolist())


if __name__ == "__main__":
    only = [a for a in sys.argv[1:] if a.startswith("--only-")]
    if not only:
        main()
    if not only or "--only-fp16" in only:
        main(FP16_CASES)
    with tempfile.TemporaryDirectory(dir=os.path.join(ROOT, "oracle")) as scratch:
        if not only:
            multiround_case(os.path.join(ROOT, "tests", "golden"), scratch)
        if not only or "--only-instruct" in only:
            instruct_case(os.path.join(ROOT, "tests", "golden"), scratch)
        if not only or "--only-hostprep" in only:
            host_prep_cases(os.path.join(ROOT, "tests", "golden"))
        if not only or "--only-prepare" in only:
            prepare_cases(os.path.join(ROOT, "tests", "golden"), scratch)
        if not only or "--only-multiround-b2" in only:
            multiround_b2_case(os.path.join(ROOT, "tests", "golden"), scratch)
        if not only or "--only-keeprates" in only:
            keep_rate_cases(os.path.join(ROOT, "tests", "golden"), scratch)
        if not only or "--only-nocache" in only:
            nocache_case(os.path.join(ROOT, "tests", "golden"), scratch)
