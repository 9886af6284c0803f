"""Pin the CPU oracle (oracle/dynllava_oracle.py) against vectors produced by the UNMODIFIED reference
(oracle/gen_golden.py, run in the build container).  Same torch, same CPU kernels, same op order -> bit-exact."""
import copy
import glob
import os

import pytest
import torch

from dynamic_llava_b200.presets import preset
from dynamic_llava_b200.synthetic import numpy_state_dict, synthetic_inputs
from oracle import dynllava_oracle as O

CASES = sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "toy_*.pt")))
DT = {"torch.float32": torch.float32, "torch.bfloat16": torch.bfloat16, "torch.float16": torch.float16}


def _load(golden_dir, name):
    g = torch.load(os.path.join(golden_dir, name + ".pt"), weights_only=False)
    cfg = preset(g["preset"])
    cfg["sparse_config"]["use_output_text_predictor"] = True
    dt = DT[g["dtype"]]
    sd = {k: v.to(dt) for k, v in numpy_state_dict(cfg, seed=g["weight_seed"]).items()}
    cfg["sparse_config"]["use_output_text_predictor"] = g["use_output_text_predictor"]
    ids, pix = synthetic_inputs(cfg, g["B"], g["T"], seed=g["input_seed"], dtype=dt)
    return g, cfg, sd, ids, pix


def test_fixtures_present():
    assert len(CASES) >= 6


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_bitexact(golden_dir, name):
    g, cfg, sd, ids, pix = _load(golden_dir, name)
    torch.set_num_threads(8)
    res = O.greedy_loop(sd, cfg, ids, pix, g["n_new"], forced_text_decisions=g["forced_decisions"])
    assert torch.equal(res["tokens"], g["tokens"])
    assert torch.equal(res["logits"], g["logits"]), (res["logits"] - g["logits"]).abs().max()
    assert torch.equal(res["true_len"], g["true_len_hist"][-1])
    kv_shapes = torch.tensor([k.shape[-2] for (k, v) in res["past"][0]])
    assert torch.equal(kv_shapes, g["kv_shapes"])
    if "k_last" in g:
        for l, (k, v) in enumerate(res["past"][0]):
            assert torch.equal(k, g["k_last"][l]) and torch.equal(v, g["v_last"][l])


def test_vision_logits_and_layer_inputs(golden_dir):
    g, cfg, sd, ids, pix = _load(golden_dir, "toy_bf16_b1_natural")
    rec = {}
    O.model_forward(sd, cfg, ids, images=pix, record=rec)
    assert torch.equal(rec["vision_logits"], g["vision_logits"])
    assert len(rec["layer_inputs"]) == len(g["prefill_layer_inputs"])
    for a, b in zip(rec["layer_inputs"], g["prefill_layer_inputs"]):
        assert torch.equal(a, b)
    # fact 2: layers 0-1 see all P tokens, layers >= 2 only the kept ones
    P = g["T"] - 1 + 16
    assert rec["layer_inputs"][0].shape[1] == P and rec["layer_inputs"][2].shape[1] == P - 16 + 8
    # fact 8: RoPE positions are the original (pre-drop) positions
    pos = rec["position_ids"][0]
    assert pos[:5].tolist() == list(range(5)) and pos[-6:].tolist() == list(range(P - 6, P))
    assert torch.equal(pos[5:13], rec["keep_index"][0] + 5)


def test_forced_decisions_realised(golden_dir):
    g, cfg, sd, ids, pix = _load(golden_dir, "toy_bf16_b2_forced")
    fd = g["forced_decisions"]
    P = g["T"] - 1 + 16
    # layers 0-1 always append (fact 4); layers >= 2 append iff decided
    assert g["true_len_hist"][-1][0].tolist() == [P + g["n_new"] - 1] * 2
    assert g["true_len_hist"][-1][2].tolist() == (P - 8 + fd.sum(0)).tolist()


def test_batched_quirk_differs_from_per_sample(golden_dir):
    """for B>1 with dropping, the reference's zero-padded cache changes the logits of the shorter sample;
    per-sample (B==1) semantics are the clean target.  Both oracles must exist and must differ here."""
    g, cfg, sd, ids, pix = _load(golden_dir, "toy_bf16_b2_forced")
    kw = dict(forced_text_decisions=g["forced_decisions"], teacher_tokens=g["tokens"])
    batched = O.greedy_loop(sd, cfg, ids, pix, g["n_new"], **kw)
    single = O.greedy_loop(sd, cfg, ids, pix, g["n_new"], per_sample=True, **kw)
    assert torch.equal(batched["logits"], g["logits"])
    assert torch.equal(single["true_len"], batched["true_len"])
    d = (batched["logits"] - single["logits"]).abs().amax(dim=(0, 2))
    assert d.max() > 1e-2  # the quirk is visible
    # prefill has no padding: B=2 vs B=1 differ only by bf16 GEMM blocking noise (measured ~0.06 on |logit|~10)
    assert (single["logits"][0] - batched["logits"][0]).abs().max() < 0.02 * batched["logits"][0].abs().max()
    # ...while the sample that dropped tokens next to a longer batch-mate deviates far more once padding appears
    assert d.max() > 3 * (single["logits"][0] - batched["logits"][0]).abs().max()


def test_hf_greedy_generate_semantics(golden_dir):
    g, cfg, sd, ids, pix = _load(golden_dir, "toy_bf16_b3_nodrop")
    out = O.hf_greedy_generate(sd, cfg, ids, pix, max_new_tokens=g["n_new"], eos_token_id=None)
    assert torch.equal(out, g["tokens"])
    eos = int(g["tokens"][1, 2])
    out = O.hf_greedy_generate(sd, cfg, ids, pix, max_new_tokens=g["n_new"], eos_token_id=eos, pad_token_id=0)
    assert out[1, 2] == eos and (out[1, 3:] == 0).all()
    assert torch.equal(out[0], g["tokens"][0][: out.shape[1]])


def test_multiround_continuation_matches_reference_bitexact(golden_dir):
    """second user turn: q_len > 1 over a ragged cache (use_instruct_predictor=False as in README.md:110)"""
    g = torch.load(os.path.join(golden_dir, "multiround_bf16_b1.pt"), weights_only=False)
    cfg = preset("toy")
    cfg["sparse_config"]["use_output_text_predictor"] = True
    sd = {k: v.to(torch.bfloat16) for k, v in numpy_state_dict(cfg, seed=0).items()}
    ids, pix = synthetic_inputs(cfg, 1, 12, seed=1, dtype=torch.bfloat16)
    logits, cont_lg, lens, toks, past = O.multiround_loop(sd, cfg, ids, pix, g["cont_ids"])
    assert toks == g["tokens"]
    assert torch.equal(cont_lg, g["cont_logits"]) and torch.equal(logits, g["step_logits"])
    assert torch.equal(lens, g["true_len_hist"])
    assert [k.shape[-2] for (k, v) in past[0]] == g["kv_shapes"].tolist()


@pytest.mark.parametrize("variant", ["natural", "forced"])
def test_instruct_predictor_matches_reference_bitexact(golden_dir, variant):
    """use_instruct_predictor=True: prefill-time instruction dropping (dynamic_modeling_llama.py:2261-2375) and
    per-token persistence on a new user turn (:2506-2521, cache_utils.py:165-241)"""
    g = torch.load(os.path.join(golden_dir, "instruct_bf16_b1.pt"), weights_only=False)
    cfg = preset("toy")
    cfg["sparse_config"]["use_output_text_predictor"] = True
    cfg["sparse_config"]["use_instruct_predictor"] = True
    sd = {k: v.to(torch.bfloat16) for k, v in numpy_state_dict(cfg, seed=0).items()}
    ids, pix = synthetic_inputs(cfg, 1, g["T"], seed=1, dtype=torch.bfloat16)
    forced = variant == "forced"
    rec = {}
    logits, cont_lg, lens, toks, past = O.multiround_loop(sd, cfg, ids, pix, g["cont_ids"],
                                                          forced_prefill=g["forced_prefill"] if forced else None,
                                                          forced_turn=g["forced_turn"] if forced else None, record=rec)
    r = g[variant]
    assert toks == r["tokens"]
    assert torch.equal(rec["instruct_logits_prefill"], r["instruct_logits_prefill"])
    assert torch.equal(rec["instruct_logits_turn"], r["instruct_logits_turn"])
    assert torch.equal(cont_lg, r["cont_logits"]) and torch.equal(logits, r["step_logits"])
    assert torch.equal(lens, r["true_len_hist"])
    assert [k.shape[-2] for (k, v) in past[0]] == r["kv_shapes"].tolist()


@pytest.mark.parametrize("variant", ["natural", "forced"])
def test_nocache_output_sparsification_matches_reference_bitexact(golden_dir, variant):
    """use_cache=False: the re-fed sequence loses generated tokens at the sparse layer by the output predictor's
    top-`max count` scores (dynamic_modeling_llama.py:2393-2504), `answer_indice` sticky from the first call, the first call's
    duplicated last token included (the `[answer_indice:-1]` slice is empty there)."""
    g = torch.load(os.path.join(golden_dir, "nocache_bf16_b2.pt"), weights_only=False)
    cfg = preset("toy")
    cfg["sparse_config"]["use_output_text_predictor"] = True
    sd = {k: v.to(torch.bfloat16) for k, v in numpy_state_dict(cfg, seed=0).items()}
    ids, pix = synthetic_inputs(cfg, g["B"], g["T"], seed=1, dtype=torch.bfloat16)
    r = O.nocache_loop(sd, cfg, ids, pix, g["n_calls"], forced_output_logits=g["forced_logits"] if variant == "forced" else None)
    ref = g[variant]
    assert r["answer_indice"] == ref["answer_indice"] and r["seq_lens"] == ref["seq_lens"]
    assert ref["seq_lens"][0] == ref["answer_indice"] + 1  # the quirk: nothing to thin yet, the last token appears twice
    assert torch.equal(r["tokens"], ref["tokens"])
    assert torch.equal(r["last_logits"], ref["last_logits"]) and torch.equal(r["final_logits"], ref["final_logits"])
    for a, b in zip(r["sparse_layer_inputs"], ref["sparse_layer_inputs"]):
        assert torch.equal(a, b)
    for a, b in zip(r["text_logits"], ref["text_logits"]):
        assert torch.equal(a, b)
    if variant == "forced":  # the hook patterns really thin the sequence, and by different per-sample counts
        assert ref["seq_lens"][-1] < ref["answer_indice"] + g["n_calls"] - 1


def test_prepare_multimodal_edge_cases_match_reference_bitexact(golden_dir):
    """prepare_inputs_labels_for_multimodal alone (dynamic_llava_arch.py:169-601): spliced embeddings and index dicts for
    `tokenizer_model_max_length` truncation, a placeholder right after BOS / as the last id, and a batch with a text-only row."""
    from oracle.cases import PREPARE_CASES, prepare_case_inputs
    g = torch.load(os.path.join(golden_dir, "prepare_cases_bf16.pt"), weights_only=False)
    cfg = preset("toy")
    cfg["sparse_config"]["use_output_text_predictor"] = True
    sd = {k: v.to(torch.bfloat16) for k, v in numpy_state_dict(cfg, seed=0).items()}
    assert set(g["cases"]) == set(PREPARE_CASES)
    for name in PREPARE_CASES:
        ids, pix, kw = prepare_case_inputs(cfg, name)
        ref = g["cases"][name]
        if "raises" in ref:  # left padding with a dict-less row: the reference itself fails (IndexError); nothing to pin
            continue
        embeds, indices = O.prepare_multimodal(sd, cfg, ids, pix, **kw)
        assert indices == ref["indices"], (name, indices, ref["indices"])
        assert torch.equal(embeds, ref["embeds"]), name


def test_hf_greedy_search_bookkeeping_matches_installed_transformers():
    """`O.hf_greedy_search` (EOS -> pad for finished rows, stop when all rows are finished, new tokens only) against the INSTALLED
    transformers' own `generate` (greedy, entered with `inputs_embeds` like dynamic_llava_llama.py:146-152) on a stand-in Llama with a
    ragged EOS pattern.  transformers 4.37.2 (the reference's pin) is not installable here; this pins the restated loop to a real HF
    implementation of the same search."""
    transformers = pytest.importorskip("transformers")
    from transformers import LlamaConfig, LlamaForCausalLM
    torch.manual_seed(0)
    hcfg = LlamaConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=2,
                       vocab_size=48, max_position_embeddings=128, pad_token_id=0, bos_token_id=1, eos_token_id=2)
    model = LlamaForCausalLM(hcfg).eval()
    B, P, n_new = 4, 7, 12
    embeds = torch.randn(B, P, 64) * 0.5

    def run_hf(eos, rows=slice(None)):
        e = embeds[rows]
        kw = dict(inputs_embeds=e, attention_mask=torch.ones(e.shape[0], P, dtype=torch.long), do_sample=False, max_new_tokens=n_new,
                  pad_token_id=0, num_beams=1)
        # "no EOS": an id outside the vocabulary can never fire
        return model.generate(eos_token_id=hcfg.vocab_size + 5 if eos is None else eos, **kw)

    def run_mine(eos, rows=slice(None)):
        st = {"past": None}
        e = embeds[rows]

        def step(t, prev):
            with torch.no_grad():
                if t == 0:
                    o = model(inputs_embeds=e, use_cache=True)
                else:
                    o = model(input_ids=prev.reshape(-1, 1), past_key_values=st["past"], use_cache=True)
            st["past"] = o.past_key_values
            return o.logits[:, -1, :]
        return O.hf_greedy_search(step, e.shape[0], n_new, eos, 0)
    with torch.no_grad():
        free = run_hf(None)
    assert free.shape == (B, n_new)
    assert torch.equal(run_mine(None), free)
    # an EOS id that rows reach at different steps (and one row possibly never): the token row 0 emits at step 3
    eos = int(free[0, 3])
    with torch.no_grad():
        hf = run_hf(eos)
    mine = run_mine(eos)
    assert hf.shape[1] <= n_new
    assert torch.equal(mine, hf), (mine, hf)
    hit = (free == eos).any(1)
    assert bool(hit[0])
    first = int((free[0] == eos).float().argmax())
    assert (hf[0, first + 1:] == 0).all()  # finished rows are padded
    # every row finishes, at different steps: the search stops early and the output is shorter than max_new_tokens
    counts = torch.stack([(free[1:] == tok).any(1).all() for tok in range(hcfg.vocab_size)])
    common = [tok for tok in range(hcfg.vocab_size) if bool(counts[tok])]
    assert common, "seeded stand-in lost its common token; pick another seed"
    eos2 = common[0]
    steps = [(free[r] == eos2).float().argmax().item() for r in range(1, B)]
    with torch.no_grad():
        hf2 = run_hf(eos2, slice(1, None))
    mine2 = run_mine(eos2, slice(1, None))
    assert torch.equal(mine2, hf2), (mine2, hf2)
    assert hf2.shape[1] == int(max(steps)) + 1 < n_new and len(set(steps)) > 1


@pytest.mark.parametrize("temperature,top_k,top_p", [(0.2, 50, 1.0), (1.0, 0, 0.9), (0.7, 50, 0.7), (1.3, 5, 0.5), (0.2, 0, 1.0), (1.0, 1, 0.3)])
def test_hf_warped_probs_match_installed_transformers(temperature, top_k, top_p):
    """`O.hf_warped_probs` (what tests/test_kernels_gpu.py checks dl_sample_next against) == the installed transformers' own
    TemperatureLogitsWarper -> TopKLogitsWarper -> TopPLogitsWarper chain followed by softmax, bit for bit, ties included."""
    pytest.importorskip("transformers")
    from transformers.generation.logits_process import TemperatureLogitsWarper, TopKLogitsWarper, TopPLogitsWarper
    g = torch.Generator().manual_seed(17)
    logits = (torch.randn(5, 512, generator=g) * 3).to(torch.bfloat16).float()  # bf16-valued logits: exact ties occur
    ids = torch.zeros(5, 1, dtype=torch.long)
    s = logits.clone()
    if temperature != 1.0:
        s = TemperatureLogitsWarper(temperature)(ids, s)
    if top_k > 0:
        s = TopKLogitsWarper(top_k=top_k, min_tokens_to_keep=1)(ids, s)
    if top_p < 1.0:
        s = TopPLogitsWarper(top_p=top_p, min_tokens_to_keep=1)(ids, s)
    want = s.softmax(dim=-1)
    got = O.hf_warped_probs(logits, temperature, top_k, top_p)
    assert torch.equal(got, want), (got - want).abs().max()


def test_keep_rate_extremes_match_reference_bitexact(golden_dir):
    """vision_keep_rate mutated on the live config to 1.0 / 0.3 / 0.07 (k = 1) / 0.0 (k = 0): prefill logits at every position, one decode
    step and the KV lengths per layer, bit for bit (the regimes tests/test_model_gpu.py::test_vision_keep_rate_extremes visits)"""
    g = torch.load(os.path.join(golden_dir, "keep_rates_bf16_b2.pt"), weights_only=False)
    cfg = preset("toy")
    cfg["sparse_config"]["use_output_text_predictor"] = True
    sd = {k: v.to(torch.bfloat16) for k, v in numpy_state_dict(cfg, seed=0).items()}
    cfg["sparse_config"]["use_output_text_predictor"] = False
    ids, pix = synthetic_inputs(cfg, g["B"], g["T"], seed=1, dtype=torch.bfloat16)
    assert set(g["rates"]) == {1.0, 0.3, 0.07, 0.0}
    for rate, ref in g["rates"].items():
        assert "raises" not in ref
        cfg["sparse_config"]["vision_keep_rate"] = rate
        logits, past = O.model_forward(sd, cfg, ids, images=pix)
        assert torch.equal(logits, ref["logits"]), rate
        nxt = logits[:, -1].argmax(-1).reshape(-1, 1)
        step, past = O.model_forward(sd, cfg, nxt, past=past)
        assert torch.equal(step, ref["step_logits"]), rate
        assert [k.shape[-2] for (k, v) in past[0]] == ref["kv_shapes"], rate


def test_multiround_b2_batched_quirk_matches_reference_bitexact(golden_dir):
    """B = 2 second user turn over diverged caches in the reference's own batched form: per-layer zero-padded stacks
    (cache_utils.py:165-241), the 5-token turn appended behind the padding (:242-250), unmasked attention over it"""
    g = torch.load(os.path.join(golden_dir, "multiround_bf16_b2.pt"), weights_only=False)
    cfg = preset("toy")
    cfg["sparse_config"]["use_output_text_predictor"] = True
    sd = {k: v.to(torch.bfloat16) for k, v in numpy_state_dict(cfg, seed=0).items()}
    B = g["B"]
    ids, pix = synthetic_inputs(cfg, B, g["T"], seed=1, dtype=torch.bfloat16)
    fd = g["forced_decisions"]
    snap = lambda past: torch.stack(list(past[1]), 0)
    logits, lens, toks, step = [], [], [], 0
    out, past = O.model_forward(sd, cfg, ids, images=pix)
    for t in range(4):
        logits.append(out[:, -1]); lens.append(snap(past))
        nxt = out[:, -1].argmax(-1).reshape(B, 1); toks.append(nxt.reshape(-1))
        out, past = O.model_forward(sd, cfg, nxt, past=past, forced_text_decision=fd[step].reshape(B, 1))
        step += 1
    logits.append(out[:, -1]); lens.append(snap(past))
    cont_logits, past = O.model_forward(sd, cfg, g["cont_ids"], past=past)
    lens.append(snap(past))
    out = cont_logits
    for t in range(3):
        nxt = out[:, -1].argmax(-1).reshape(B, 1); toks.append(nxt.reshape(-1))
        out, past = O.model_forward(sd, cfg, nxt, past=past, forced_text_decision=fd[step].reshape(B, 1))
        step += 1
        logits.append(out[:, -1]); lens.append(snap(past))
    assert torch.equal(torch.stack(toks, 1), g["tokens"])
    assert torch.equal(torch.stack(lens, 0), g["true_len_hist"])
    assert torch.equal(cont_logits, g["cont_logits"]) and torch.equal(torch.stack(logits, 0), g["step_logits"])
    assert [k.shape[-2] for (k, v) in past[0]] == g["kv_shapes"].tolist()
    assert g["true_len_hist"][-1][-1, 0] != g["true_len_hist"][-1][-1, 1]  # the lengths really diverged


def test_nocache_with_instruct_predictor_takes_the_instruct_branch(golden_dir):
    """elif precedence of the reference (dynamic_modeling_llama.py:2261 before :2393): `use_cache=False` with the instruct predictor on drops
    instruction tokens at the prefill and never thins the generated tokens; `answer_indice` stays None"""
    g = torch.load(os.path.join(golden_dir, "nocache_bf16_b2.pt"), weights_only=False)["instruct_precedence"]
    cfg = preset("toy")
    cfg["sparse_config"]["use_output_text_predictor"] = True
    cfg["sparse_config"]["use_instruct_predictor"] = True
    sd = {k: v.to(torch.bfloat16) for k, v in numpy_state_dict(cfg, seed=0).items()}
    _, pix = synthetic_inputs(cfg, 1, 20, seed=1, dtype=torch.bfloat16)
    state = {"answer_indice": None}
    logits, past = O.model_forward(sd, cfg, g["ids"], images=pix, use_cache=False, state=state)
    assert past is None and state["answer_indice"] is None and g["answer_indice"] is None
    assert torch.equal(logits, g["logits"])
