"""CPU oracle for Dynamic-LLaVA's sparsified prefill/decode path.  TEST INFRASTRUCTURE ONLY.

This module is a functional, state-dict-driven restatement (plain torch on CPU) of the reference algorithm
(Osilly/dynamic_llava @ e1469bc).  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import it; the product package `dynamic_llava_b200` never does.

Parity pinning: the reference ships NO tests or golden vectors for this path, so the oracle is pinned
against outputs of the reference itself: `oracle/gen_golden.py` imports the unmodified reference modules from
`/root/reference` (under the 2-line transformers-5 shim of SURVEY.md App. A), runs them on CPU on seeded toy
inputs, and commits the outputs under `tests/golden/`.  `tests/test_oracle_golden.py` checks this restatement against
those vectors bit-for-bit.  HF `generate()` cannot be run against transformers 5.x, so
`hf_greedy_generate` below restates transformers-4.37.2 greedy-search semantics from its documented behaviour:
that one function is "parity unpinned".

Every function cites the reference lines it follows (paths relative to /root/reference/llava/model/).
"""
from __future__ import annotations

import copy
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

IMAGE_TOKEN_INDEX = -200  # llava/constants.py:8
USER_IDS = [11889, 29901]  # dynamic_llava_arch.py:36  special_text["USER:"]

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]


# --------------------------------------------------------------------------------------------------------------
# elementwise blocks
# --------------------------------------------------------------------------------------------------------------
def rms_norm(x: Tensor, w: Tensor, eps: float) -> Tensor:
    """language_model/dynamic_modeling_llama.py:134-139 — fp32 normalise, round to input dtype, THEN scale."""
    dt = x.dtype
    x32 = x.to(torch.float32)
    var = x32.pow(2).mean(-1, keepdim=True)
    x32 = x32 * torch.rsqrt(var + eps)
    return w * x32.to(dt)


def rope_tables(head_dim: int, max_pos: int, theta: float, dtype, device=None) -> Tuple[Tensor, Tensor]:
    """dynamic_modeling_llama.py:152-184 — fp32 tables, half-split layout cat(freqs, freqs), cast on use."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, device=device).float() / head_dim))
    t = torch.arange(max_pos, dtype=inv_freq.dtype, device=device)
    freqs = torch.outer(t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def _rotate_half(x: Tensor) -> Tensor:
    """dynamic_modeling_llama.py:253-257"""
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def apply_rope(q: Tensor, k: Tensor, cos: Tensor, sin: Tensor, position_ids: Tensor) -> Tuple[Tensor, Tensor]:
    """dynamic_modeling_llama.py:260-285 — q,k are [B, heads, L, hd]; position_ids [B or 1, L]."""
    c = cos[position_ids].unsqueeze(1)
    s = sin[position_ids].unsqueeze(1)
    return (q * c) + (_rotate_half(q) * s), (k * c) + (_rotate_half(k) * s)


def swiglu_mlp(x: Tensor, sd: StateDict, p: str) -> Tensor:
    """dynamic_modeling_llama.py:328 — down(silu(gate(x)) * up(x))."""
    g = F.linear(x, sd[p + "gate_proj.weight"])
    u = F.linear(x, sd[p + "up_proj.weight"])
    return F.linear(F.silu(g) * u, sd[p + "down_proj.weight"])


def layer_norm(x: Tensor, sd: StateDict, p: str, eps: float = 1e-5) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[p + "weight"], sd[p + "bias"], eps)


def lin(x: Tensor, sd: StateDict, p: str) -> Tensor:
    return F.linear(x, sd[p + "weight"], sd.get(p + "bias"))


# --------------------------------------------------------------------------------------------------------------
# ragged KV cache (language_model/cache_utils.py:63-318)
# --------------------------------------------------------------------------------------------------------------
class OracleCache:
    """Restates DynamicCachePlus: per-layer K/V [B, heads, len, hd] grown by concatenation plus `true_len`
    (list over layers of int64[B]).  Keeps the reference's batched behaviour, INCLUDING the B>1 zero-pad that later
    attends as score-0 keys, so that it can serve as the "batched reference" oracle."""

    def __init__(self):
        self.k: List[Tensor] = []
        self.v: List[Tensor] = []
        self.true_len: List[Tensor] = []

    def length(self, layer: int = 0) -> int:
        """cache_utils.py:272-276 (+ :49-60, no max length): PADDED tensor length."""
        return 0 if len(self.k) <= layer else self.k[layer].shape[-2]

    def peek(self, k_new: Tensor, v_new: Tensor, layer: int) -> Tuple[Tensor, Tensor]:
        """cache_utils.py:256-268 `get_cache`: cache ++ new, not persisted."""
        if len(self.k) <= layer:
            return k_new, v_new
        return torch.cat([self.k[layer], k_new], dim=-2), torch.cat([self.v[layer], v_new], dim=-2)

    def update(self, k_new: Tensor, v_new: Tensor, layer: int, decision: Optional[Tensor] = None):
        """cache_utils.py:109-253."""
        B, _, N, _ = k_new.shape
        if len(self.k) <= layer:  # :139-149 first write (prefill)
            self.k.append(k_new)
            self.v.append(v_new)
            self.true_len.append(decision.sum(dim=-1) if decision is not None else torch.tensor([N], device=k_new.device).repeat(B))
        elif decision is None:  # :242-250 plain append
            self.k[layer] = torch.cat([self.k[layer], k_new], dim=-2)
            self.v[layer] = torch.cat([self.v[layer], v_new], dim=-2)
            self.true_len[layer] += N
        elif B == 1 and N == 1:  # :153-164 single-sample conditional append
            if bool(decision[0, 0]):
                self.k[layer] = torch.cat([self.k[layer], k_new], dim=-2)
                self.v[layer] = torch.cat([self.v[layer], v_new], dim=-2)
                self.true_len[layer] += N
        else:  # :165-241 per-sample slice / append / zero-pad / stack
            ks, vs = [], []
            for b in range(B):
                keep = decision[b]
                tl = int(self.true_len[layer][b])
                ks.append(torch.cat([self.k[layer][b, :, :tl, :], k_new[b, :, keep, :]], dim=-2))
                vs.append(torch.cat([self.v[layer][b, :, :tl, :], v_new[b, :, keep, :]], dim=-2))
                self.true_len[layer][b] += int(keep.sum())
            mx = max(t.shape[-2] for t in ks)
            pad = lambda t: torch.cat([t, t.new_zeros(t.shape[0], mx - t.shape[-2], t.shape[-1])], dim=-2)
            self.k[layer] = torch.stack([pad(t) for t in ks])
            self.v[layer] = torch.stack([pad(t) for t in vs])
        return self.k[layer], self.v[layer]

    def to_legacy(self):
        """cache_utils.py:295-302."""
        return tuple((self.k[i], self.v[i]) for i in range(len(self.k))), self.true_len

    @classmethod
    def from_legacy(cls, past) -> "OracleCache":
        """cache_utils.py:304-318."""
        c = cls()
        if past is not None:
            for (k, v) in past[0]:
                c.k.append(k)
                c.v.append(v)
            c.true_len = past[1]
        return c


# --------------------------------------------------------------------------------------------------------------
# predictors
# --------------------------------------------------------------------------------------------------------------
def vision_predictor(sd: StateDict, x: Tensor, policy: Tensor, nhead: int, p: str = "model.image_score_predictor.") -> Tensor:
    """dynamic_modeling_llama.py:1348-1359 with blocks custom_transformer_layer.py:153-180, :116-123, :320-323.
    x [B, N, H]; policy [B, N, 1] (all ones at inference, :1779-1785) -> logits [B, N, 2]."""
    y = F.gelu(lin(layer_norm(x, sd, p + "down_mlp.0."), sd, p + "down_mlp.1."))
    y = y * policy
    j = 0
    while (p + f"transformer.{j}.norm1.weight") in sd:
        q = p + f"transformer.{j}."
        h = layer_norm(y, sd, q + "norm1.")
        B, N, C = h.shape
        qkv = F.linear(h, sd[q + "attn.qkv.weight"]).reshape(B, N, 3, nhead, C // nhead).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2], dropout_p=0.0)
        a = lin(a.transpose(1, 2).reshape(B, N, C), sd, q + "attn.proj.")
        y = y + a
        h = layer_norm(y, sd, q + "norm2.")
        y = y + lin(F.gelu(lin(h, sd, q + "mlp.fc1.")), sd, q + "mlp.fc2.")
        j += 1
    B, N, C = y.shape
    local = y[:, :, : C // 2]
    glob = (y[:, :, C // 2:] * policy).sum(dim=1, keepdim=True) / torch.sum(policy, dim=1, keepdim=True)
    z = torch.cat([local, glob.expand(B, N, C // 2)], dim=-1)
    z = F.gelu(lin(z, sd, p + "output_mlp.0."))
    z = F.gelu(lin(z, sd, p + "output_mlp.2."))
    return lin(z, sd, p + "output_mlp.4.")


def text_predictor(sd: StateDict, x: Tensor, p: str = "model.output_text_score_predictor.") -> Tensor:
    """dynamic_modeling_llama.py:1374-1387 — LN, then 4 biased linears with GELU between."""
    q = p + "output_mlp."
    z = F.gelu(lin(layer_norm(x, sd, q + "0."), sd, q + "1."))
    z = F.gelu(lin(z, sd, q + "3."))
    z = F.gelu(lin(z, sd, q + "5."))
    return lin(z, sd, q + "7.")


def topk_keep_index(logits: Tensor, n_keep: int) -> Tensor:
    """dynamic_modeling_llama.py:1867, :1898-1908 — log_softmax in the model dtype, channel 0, top-k by argsort
    (descending), then ascending index order."""
    score = F.log_softmax(logits, dim=-1)[:, :, 0]
    keep, _ = torch.sort(torch.argsort(score, dim=1, descending=True)[:, :n_keep], dim=1, descending=False)
    return keep


# --------------------------------------------------------------------------------------------------------------
# decoder
# --------------------------------------------------------------------------------------------------------------
def _attention(sd: StateDict, cfg: dict, i: int, x: Tensor, position_ids: Tensor, cache: OracleCache,
               text_decision: Optional[Tensor], cos: Tensor, sin: Tensor) -> Tensor:
    """dynamic_modeling_llama.py:1009-1129 (sdpa class; attention_mask is None in every supported flow)."""
    p = f"model.layers.{i}.self_attn."
    B, Lq, H = x.shape
    nh = cfg["num_attention_heads"]
    hd = H // nh
    q = F.linear(x, sd[p + "q_proj.weight"]).view(B, Lq, nh, hd).transpose(1, 2)
    k = F.linear(x, sd[p + "k_proj.weight"]).view(B, Lq, nh, hd).transpose(1, 2)
    v = F.linear(x, sd[p + "v_proj.weight"]).view(B, Lq, nh, hd).transpose(1, 2)
    q, k = apply_rope(q, k, cos, sin, position_ids)
    if text_decision is not None:  # :1061-1076 attend to cache ++ new, persist only where decided
        k_all, v_all = cache.peek(k, v, i)
        cache.update(k, v, i, text_decision)
    else:  # :1078-1083
        k_all, v_all = cache.update(k, v, i)
    past_i = k_all.shape[-2] - Lq
    if Lq > 1 and past_i > 0:
        # multi-token continuation over a cache (new user turn): the per-layer mask helper (:1805-1810, transformers
        # `_prepare_4d_causal_attention_mask_for_sdpa` with attention_mask=None, q_len > 1, kv > q) materialises a causal
        # 4-D additive mask offset by the cached length, and SDPA runs with is_causal=False (:1121)
        m = torch.full((Lq, Lq), torch.finfo(q.dtype).min, dtype=q.dtype, device=q.device)
        c = torch.arange(Lq, device=q.device)
        m.masked_fill_(c < (c + 1).view(Lq, 1), 0)
        m = torch.cat([torch.zeros(Lq, past_i, dtype=q.dtype, device=q.device), m], dim=-1)[None, None].expand(B, 1, Lq, past_i + Lq)
        o = F.scaled_dot_product_attention(q, k_all, v_all, attn_mask=m, dropout_p=0.0, is_causal=False)
    else:
        o = F.scaled_dot_product_attention(q, k_all, v_all, attn_mask=None, dropout_p=0.0, is_causal=Lq > 1)  # :1114-1122
    o = o.transpose(1, 2).contiguous().reshape(B, Lq, H)
    return F.linear(o, sd[p + "o_proj.weight"])


def decoder_layer(sd, cfg, i, x, position_ids, cache, text_decision, cos, sin) -> Tensor:
    """dynamic_modeling_llama.py:1271-1305."""
    p = f"model.layers.{i}."
    eps = cfg["rms_norm_eps"]
    x = x + _attention(sd, cfg, i, rms_norm(x, sd[p + "input_layernorm.weight"], eps), position_ids, cache,
                       text_decision, cos, sin)
    x = x + swiglu_mlp(rms_norm(x, sd[p + "post_attention_layernorm.weight"], eps), sd, p + "mlp.")
    return x


def llama_forward(sd: StateDict, cfg: dict, *, input_ids: Optional[Tensor] = None, inputs_embeds: Optional[Tensor] = None,
                  position_ids: Optional[Tensor] = None, past=None, indices: Optional[List[dict]] = None,
                  forced_keep_index: Optional[Tensor] = None, forced_text_decision: Optional[Tensor] = None,
                  forced_instruct: Optional[Tensor] = None, record: Optional[dict] = None, use_cache: bool = True,
                  state: Optional[dict] = None, forced_output_logits: Optional[dict] = None):
    """DynamicLlamaModel.forward, eval branches only (dynamic_modeling_llama.py:1656-2594), followed by
    lm_head + .float() (:2709-2710).  Returns (logits fp32 [B, Lq', V], legacy past tuple | None if not use_cache).

    `use_cache=False` (:1703-1712, :2393-2504): no past is read or returned and the generated tokens between the sticky
    `state["answer_indice"]` (the reference's `self.answer_indice`, :1644, :2402-2403) and the last token are thinned at the sparse
    layer; `forced_output_logits` {n: [B,n,2]} replaces the output predictor's logits for an n-token span (what a forward hook does).

    `forced_keep_index` [B,k] / `forced_text_decision` [B,Lq] bool replace the predictors' own decisions
    (parity "under the same predictor masks"); `forced_instruct` bool [span] (prefill) / [Lq] (new turn)
    replaces the instruct predictor's `logit0 > logit1`; `record` (dict) receives intermediates."""
    sc = cfg["sparse_config"]
    if (input_ids is None) == (inputs_embeds is None):
        raise ValueError("You have to specify exactly one of input_ids or inputs_embeds")  # :1686-1695
    cache = OracleCache.from_legacy(past if use_cache else None)  # :1703-1712 (use_cache=False: layers see no cache; a fresh one is
    # bit-identical to `past_key_value=None` in :1056-1083 and is dropped at the end)
    x = F.embedding(input_ids, sd["model.embed_tokens.weight"]) if inputs_embeds is None else inputs_embeds
    B, seq = x.shape[:2]
    past_len0 = cache.length(0)
    if position_ids is None:  # :1714-1722
        position_ids = torch.arange(past_len0, seq + past_len0, dtype=torch.long, device=x.device).unsqueeze(0)
    init_n = seq
    nh = cfg["num_attention_heads"]
    cos, sin = rope_tables(cfg["hidden_size"] // nh, cfg["max_position_embeddings"], cfg["rope_theta"], x.dtype, x.device)
    have_idx = indices is not None and B == len(indices)
    text_decision = None
    SL = sc["sparse_layer"]
    if record is not None:
        record.setdefault("layer_inputs", [])
    for i in range(cfg["num_hidden_layers"]):
        past_len_i = cache.length(i)  # :1791-1794
        if sc["use_vision_predictor"] and i == SL and have_idx:  # :1826-1994
            s0, e0 = indices[0]["image"]
            n_img = e0 - s0
            imgs = [x[b, indices[b]["image"][0]: indices[b]["image"][1], :] for b in range(B)]  # :1832-1842
            mx = max(t.shape[0] for t in imgs)
            img_x = torch.stack([torch.cat([t, t.new_zeros(mx - t.shape[0], t.shape[1])], dim=0) for t in imgs])
            policy = torch.ones(B, n_img, 1, dtype=x.dtype, device=x.device)
            logits = vision_predictor(sd, img_x, policy, sc["nhead"]).reshape(B, -1, 2)
            n_keep = int(n_img * sc["vision_keep_rate"])  # :1899-1901
            keep = topk_keep_index(logits, n_keep) if forced_keep_index is None else forced_keep_index
            gathered = x[:, s0:e0, :].gather(1, keep[..., None].expand(B, n_keep, x.shape[2]))  # :1935-1942
            x = torch.cat([x[:, :s0, :], gathered, x[:, e0:, :]], dim=1)  # :1944-1951
            position_ids = torch.cat(
                [torch.arange(0, s0, device=x.device).repeat(B, 1), keep + s0, torch.arange(e0, init_n, device=x.device).repeat(B, 1)], dim=1
            ).to(torch.long)  # :1962-1983
            drop = n_img - n_keep  # :1985-1994 (mutates the caller's dicts, as the reference does)
            for d in indices:
                d["image"][1] -= drop
                for key in ("instruct", "last_instruct", "answer"):
                    d[key][0] -= drop
                    d[key][1] -= drop
            if record is not None:
                record.update(vision_logits=logits, keep_index=keep, position_ids=position_ids)
        if sc["use_text_predictor"] and i == SL:  # :1995-2521, eval branches
            if (not past_len_i) and have_idx and sc["use_instruct_predictor"]:  # :2261-2375 prefill instruct drop
                assert B == 1, "Using text predictor must keep the batch size to 1"
                a, b_ = indices[0]["last_instruct"][0], indices[0]["last_instruct"][1] - 1
                ih = x[:, a:b_, :]
                il = text_predictor(sd, ih, "model.instruct_score_predictor.").reshape(B, -1, 2)
                dec_i = il[0, :, 0] > il[0, :, 1] if forced_instruct is None else forced_instruct.reshape(-1)
                keep_i = torch.where(dec_i)[0].unsqueeze(0)
                x = torch.cat([x[:, :a, :], ih.gather(1, keep_i[..., None].expand(B, -1, x.shape[2])), x[:, b_:, :]], dim=1)
                position_ids = torch.cat([position_ids[:, :a], position_ids[:, a:b_].gather(1, keep_i), position_ids[:, b_:]], dim=1)
                drop = b_ - a - keep_i.shape[1]
                for d in indices:
                    d["instruct"][1] -= drop
                    d["last_instruct"][1] -= drop
                    d["answer"][0] -= drop
                    d["answer"][1] -= drop
                if record is not None:
                    record.update(instruct_keep_index=keep_i, instruct_logits=il)
            elif past_len_i and x.shape[1] == 1 and sc["use_output_text_predictor"]:  # :2377-2391 decode decision
                tl = text_predictor(sd, x).reshape(B, -1, 2)
                text_decision = tl[:, :, 0] > tl[:, :, 1]
                if record is not None:
                    record.update(text_logits=tl, text_decision_natural=text_decision)
                if forced_text_decision is not None:
                    text_decision = forced_text_decision
            elif have_idx and sc["use_output_text_predictor"] and not use_cache:  # :2393-2504 output stage without KV cache
                if state.get("answer_indice") is None:  # :2402-2403 sticky across calls (post-image-drop coordinates)
                    state["answer_indice"] = indices[0]["instruct"][1]
                A = state["answer_indice"]
                oh = x[:, A:-1, :]  # every generated token but the last; empty on the first two calls
                tl = text_predictor(sd, oh).reshape(B, -1, 2)
                if record is not None:
                    record.update(text_logits=tl)
                if forced_output_logits is not None and oh.shape[1] > 0:
                    tl = forced_output_logits[oh.shape[1]].to(tl.dtype).reshape(tl.shape)
                dec = tl[:, :, 0] > tl[:, :, 1]
                n_keep = int(dec.sum(dim=1).max()) if dec.shape[1] else 0  # :2417 (torch: max of an empty-dim-1 sum is 0)
                keep_o, _ = torch.sort(torch.argsort(tl[:, :, 0], dim=1, descending=True)[:, :n_keep], dim=1)  # :2420-2426
                x = torch.cat([x[:, :A, :], oh.gather(1, keep_o[..., None].expand(B, n_keep, x.shape[2])), x[:, -1:, :]], dim=1)
                position_ids = position_ids.expand(B, -1) if position_ids.shape[0] != B else position_ids
                position_ids = torch.cat([position_ids[:, :A], position_ids[:, A:-1].gather(1, keep_o), position_ids[:, -1:]], dim=1)
                if record is not None:
                    record.update(output_keep_index=keep_o, position_ids=position_ids)
            elif past_len_i and x.shape[1] > 1 and sc["use_instruct_predictor"]:  # :2506-2521 new instruct turn
                tl = text_predictor(sd, x, "model.instruct_score_predictor.").reshape(B, -1, 2)
                text_decision = tl[:, :, 0] > tl[:, :, 1] if forced_instruct is None else forced_instruct.reshape(B, -1).clone()
                text_decision[:, -1:] = True
                if record is not None:
                    record.update(instruct_logits=tl)
        if record is not None:
            record["layer_inputs"].append(x)  # what the layer actually receives (post-compaction at i == SL)
        x = decoder_layer(sd, cfg, i, x, position_ids, cache, text_decision, cos, sin)  # :2541-2552
    x = rms_norm(x, sd["model.norm.weight"], cfg["rms_norm_eps"])  # :2565
    if record is not None:
        record.update(final_hidden=x, text_decision=text_decision)
    logits = F.linear(x, sd["lm_head.weight"]).float()  # :2709-2710
    return logits, (cache.to_legacy() if use_cache else None)  # :2570-2576


# --------------------------------------------------------------------------------------------------------------
# vision tower + projector + splice
# --------------------------------------------------------------------------------------------------------------
def clip_features(sd: StateDict, cfg: dict, pixels: Tensor, p: str = "model.vision_tower.vision_tower.vision_model.") -> Tensor:
    """CLIPVisionTower.forward + feature_select (multimodal_encoder/clip_encoder.py:43-71) over HF CLIPVisionModel
    (transformers, unvendored; math per SURVEY.md App. B): hidden_states[select_layer], CLS dropped."""
    vc = cfg["vision"]
    wdt = sd[p + "embeddings.patch_embedding.weight"].dtype
    x = F.conv2d(pixels.to(wdt), sd[p + "embeddings.patch_embedding.weight"], stride=vc["patch_size"])
    x = x.flatten(2).transpose(1, 2)
    B = x.shape[0]
    x = torch.cat([sd[p + "embeddings.class_embedding"].expand(B, 1, -1), x], dim=1)
    x = x + sd[p + "embeddings.position_embedding.weight"].unsqueeze(0)
    eps = vc["layer_norm_eps"]
    x = layer_norm(x, sd, p + "pre_layrnorm.", eps)
    hs = [x]
    nh = vc["num_attention_heads"]
    for i in range(vc["num_hidden_layers"]):
        q_ = p + f"encoder.layers.{i}."
        h = layer_norm(x, sd, q_ + "layer_norm1.", eps)
        Bq, N, C = h.shape
        q = lin(h, sd, q_ + "self_attn.q_proj.").view(Bq, N, nh, C // nh).transpose(1, 2)
        k = lin(h, sd, q_ + "self_attn.k_proj.").view(Bq, N, nh, C // nh).transpose(1, 2)
        v = lin(h, sd, q_ + "self_attn.v_proj.").view(Bq, N, nh, C // nh).transpose(1, 2)
        a = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, is_causal=False, scale=(C // nh) ** -0.5)
        a = a.transpose(1, 2).reshape(Bq, N, C).contiguous()
        x = x + lin(a, sd, q_ + "self_attn.out_proj.")
        h = layer_norm(x, sd, q_ + "layer_norm2.", eps)
        h = lin(h, sd, q_ + "mlp.fc1.")
        h = h * torch.sigmoid(1.702 * h)  # quick_gelu
        x = x + lin(h, sd, q_ + "mlp.fc2.")
        hs.append(x)
    return hs[cfg["mm_vision_select_layer"]][:, 1:].to(pixels.dtype)


def encode_images(sd: StateDict, cfg: dict, pixels: Tensor) -> Tensor:
    """dynamic_llava_arch.py:163-166 — tower then mlp2x_gelu projector (multimodal_projector/builder.py:172-179)."""
    f = clip_features(sd, cfg, pixels)
    return lin(F.gelu(lin(f, sd, "model.mm_projector.0.")), sd, "model.mm_projector.2.")


def prepare_multimodal(sd: StateDict, cfg: dict, input_ids: Tensor, pixels: Optional[Tensor],
                       max_len: Optional[int] = None, padding_side: str = "right"):
    """prepare_inputs_labels_for_multimodal with labels=None, attention_mask=None, one image per sample, "flat" merge
    (dynamic_llava_arch.py:169-601).  Returns (inputs_embeds [B,P,H] | None, indices | None)."""
    if pixels is None or input_ids.shape[1] == 1:  # :179-188
        return None, None
    feats = encode_images(sd, cfg, pixels)  # :267
    emb_w = sd["model.embed_tokens.weight"]
    embeds, indices = [], []
    img_i = 0
    for b in range(input_ids.shape[0]):
        ids = input_ids[b]
        n_img = int((ids == IMAGE_TOKEN_INDEX).sum())
        if n_img == 0:  # :315-324 — no dict appended for this sample
            embeds.append(torch.cat([F.embedding(ids, emb_w), feats[img_i][0:0]], dim=0))
            img_i += 1
            continue
        if n_img != 1:
            raise NotImplementedError("oracle covers one <image> per sample")
        s = int(torch.where(ids == IMAGE_TOKEN_INDEX)[0].item())  # :330-334
        T = ids.shape[0]
        ans = T  # labels are all -100 -> answer segment is empty (:290-291, :334)
        sys_ids, ins_ids, ans_ids = ids[:s], ids[s + 1: ans], ids[ans:T]
        text = F.embedding(torch.cat([sys_ids, ins_ids, ans_ids]), emb_w)  # :364-366
        img = feats[img_i]
        img_i += 1
        sys_e, ins_e, ans_e = text[:s], text[s: ans - 1], text[ans - 1: T - 1]  # :368-396
        e = torch.cat([sys_e, img, ins_e, ans_e])  # :398-405
        embeds.append(e)
        user = torch.tensor(USER_IDS, device=ids.device)
        m = [j for j in range(len(ins_ids) - len(user) + 1) if torch.equal(ins_ids[j: j + len(user)], user)]  # :418-454
        last = m[-1] if len(m) else 0
        i0 = sys_e.shape[0]
        i1 = i0 + img.shape[0]
        i2 = i1 + ins_e.shape[0]
        indices.append({"system": [0, i0], "image": [i0, i1], "instruct": [i1, i2], "answer": [i2, e.shape[0]],
                        "last_instruct": [i1 + last, i2]})  # :473-490
    if max_len is not None:  # :493-506
        embeds = [e[:max_len] for e in embeds]
        for d in indices:
            for v in d.values():
                v[0] = min(v[0], max_len)
                v[1] = min(v[1], max_len)
    mx = max(e.shape[0] for e in embeds)  # :509-579
    out = []
    for i, e in enumerate(embeds):
        z = e.new_zeros(mx - e.shape[0], e.shape[1])
        if padding_side == "left":
            out.append(torch.cat([z, e], dim=0))
            if i < len(indices):
                for v in indices[i].values():
                    v[0] += mx - e.shape[0]
                    v[1] += mx - e.shape[0]
        else:
            out.append(torch.cat([e, z], dim=0))
    return torch.stack(out, dim=0), indices


def model_forward(sd: StateDict, cfg: dict, input_ids: Tensor, *, images: Optional[Tensor] = None, past=None,
                  forced_keep_index=None, forced_text_decision=None, forced_instruct=None, record=None, use_cache=True, state=None,
                  forced_output_logits=None):
    """DynamicLlavaLlamaForCausalLM.forward (language_model/dynamic_llava_llama.py:68-115)."""
    embeds, indices = prepare_multimodal(sd, cfg, input_ids, images)
    if embeds is None:
        return llama_forward(sd, cfg, input_ids=input_ids, past=past, forced_text_decision=forced_text_decision,
                             forced_instruct=forced_instruct, record=record, use_cache=use_cache, state=state)
    return llama_forward(sd, cfg, inputs_embeds=embeds, past=past, indices=indices, forced_keep_index=forced_keep_index,
                         forced_text_decision=forced_text_decision, forced_instruct=forced_instruct, record=record,
                         use_cache=use_cache, state=state, forced_output_logits=forced_output_logits)


# --------------------------------------------------------------------------------------------------------------
# generation loops
# --------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def greedy_loop(sd: StateDict, cfg: dict, input_ids: Tensor, images: Tensor, n_new: int, *, forced_keep_index=None,
                forced_text_decisions: Optional[Tensor] = None, teacher_tokens: Optional[Tensor] = None,
                per_sample: bool = False):
    """The reference's own hand-rolled loop shape (llava/dynamic_eval/model_lvis_for_meteor.py:145-174; batched:
    bench_test/dynamic_llava_long_text_mem.py:310-338): forward; next = argmax(logits[:, -1]); feed back with the
    returned past and images=None.

    forced_text_decisions: bool [n_new-1, B] — row t-1 is applied at decode call t (the prefill makes no decision).
    teacher_tokens: int64 [B, n_new] — if given, these are fed back instead of the argmax (teacher forcing).
    per_sample=True runs each sample alone (B==1 cache semantics, no zero-pad quirk) and stacks the results.
    Returns dict(tokens [B,n_new], logits [n_new,B,V] (last position), true_len [L,B], kv_len_hist, decisions)."""
    if per_sample and input_ids.shape[0] > 1:
        outs = [greedy_loop(sd, cfg, input_ids[b: b + 1], images[b: b + 1], n_new,
                            forced_keep_index=None if forced_keep_index is None else forced_keep_index[b: b + 1],
                            forced_text_decisions=None if forced_text_decisions is None else forced_text_decisions[:, b: b + 1],
                            teacher_tokens=None if teacher_tokens is None else teacher_tokens[b: b + 1])
                for b in range(input_ids.shape[0])]
        res = {"tokens": torch.cat([o["tokens"] for o in outs], 0), "logits": torch.cat([o["logits"] for o in outs], 1),
               "true_len": torch.cat([o["true_len"] for o in outs], 1),
               "decisions": None if outs[0]["decisions"] is None else torch.cat([o["decisions"] for o in outs], 1),
               "keep_index": None if outs[0]["keep_index"] is None else torch.cat([o["keep_index"] for o in outs], 0),
               "past": [o["past"] for o in outs]}
        return res
    past = None
    ids = input_ids
    toks, lgs, decs = [], [], []
    keep = None
    for t in range(n_new):
        rec = {}
        ftd = None
        if t > 0 and forced_text_decisions is not None:
            ftd = forced_text_decisions[t - 1].reshape(-1, 1)
        logits, past = model_forward(sd, cfg, ids, images=images if t == 0 else None, past=past,
                                     forced_keep_index=forced_keep_index if t == 0 else None,
                                     forced_text_decision=ftd, record=rec)
        if t == 0:
            keep = rec.get("keep_index")
        elif rec.get("text_decision") is not None:
            decs.append(rec["text_decision"].reshape(1, -1))
        last = logits[:, -1, :]
        nxt = torch.argmax(last, dim=-1)
        lgs.append(last)
        toks.append(nxt)
        ids = (teacher_tokens[:, t] if teacher_tokens is not None else nxt).reshape(-1, 1)
    return {"tokens": torch.stack(toks, 1), "logits": torch.stack(lgs, 0), "true_len": torch.stack(list(past[1]), 0),
            "decisions": torch.cat(decs, 0) if decs else None, "keep_index": keep, "past": past}


@torch.no_grad()
def hf_warped_probs(logits: Tensor, temperature: float, top_k: int, top_p: float) -> Tensor:
    """The distribution HF's `sample` draws from: TemperatureLogitsWarper -> TopKLogitsWarper -> TopPLogitsWarper -> softmax
    (transformers logits_process.py, min_tokens_to_keep = 1; top_k = 0 / top_p = 1 switch a warper off), on the fp32 logits the
    reference hands to generate() (dynamic_modeling_llama.py:2710, model_vqa_loader.py:169-172).  Restated; pinned against the installed
    transformers' warper classes in tests/test_oracle_golden.py (4.37.2 itself is not installable here)."""
    scores = logits.float() / temperature
    if top_k > 0:
        kth = torch.topk(scores, min(top_k, scores.shape[-1]))[0][..., -1, None]
        scores = scores.masked_fill(scores < kth, float("-inf"))
    if top_p < 1.0:
        srt, idx = torch.sort(scores, descending=False)
        cum = srt.softmax(dim=-1).cumsum(dim=-1)
        rem = cum <= (1 - top_p)
        rem[..., -1:] = False
        scores = scores.masked_fill(rem.scatter(1, idx, rem), float("-inf"))
    return scores.softmax(dim=-1)


def hf_greedy_search(step_fn, batch: int, max_new_tokens: int, eos_token_id: Optional[int] = 2, pad_token_id: int = 0) -> Tensor:
    """The bookkeeping of transformers' greedy search (4.37.2 `GenerationMixin.greedy_search`, restated from documented behaviour,
    SURVEY.md App. C): `step_fn(t, prev)` returns the last-position logits [B, V] of call t (prev = the tokens fed back, None at t = 0).
    Rows that already produced EOS emit `pad_token_id`; the loop stops when every row has finished; only new tokens are returned.
    4.37.2 is not installable here; tests/test_oracle_golden.py pins this loop token for token against the installed transformers'
    own `generate` (greedy, inputs_embeds entry, ragged EOS) on a stand-in Llama."""
    unfinished = torch.ones(batch, dtype=torch.long)
    out, prev = [], None
    for t in range(max_new_tokens):
        nxt = torch.argmax(step_fn(t, prev), dim=-1).cpu()
        if eos_token_id is not None:
            nxt = nxt * unfinished + pad_token_id * (1 - unfinished)
            unfinished = unfinished * (nxt != eos_token_id).long()
        out.append(nxt)
        prev = nxt
        if eos_token_id is not None and int(unfinished.max()) == 0:
            break
    return torch.stack(out, 1)


def hf_greedy_generate(sd: StateDict, cfg: dict, input_ids: Tensor, images: Tensor, max_new_tokens: int,
                       eos_token_id: Optional[int] = 2, pad_token_id: int = 0, **forced):
    """`generate` as entered from language_model/dynamic_llava_llama.py:146-152: the greedy-search bookkeeping above around the
    reference forward (prompt as embeddings + images at t = 0, then one id per call over the returned past).  PARITY UNPINNED against
    transformers 4.37.2 itself (not in this container); the loop semantics are pinned to the installed transformers (see above)."""
    ftd_all = forced.get("forced_text_decisions")
    state = {"past": None}

    def step(t, prev):
        ids = input_ids if t == 0 else prev.reshape(-1, 1).to(input_ids.device)
        ftd = ftd_all[t - 1].reshape(-1, 1) if (t > 0 and ftd_all is not None) else None
        logits, state["past"] = model_forward(sd, cfg, ids, images=images if t == 0 else None, past=state["past"],
                                              forced_keep_index=forced.get("forced_keep_index") if t == 0 else None,
                                              forced_text_decision=ftd)
        return logits[:, -1, :]
    return hf_greedy_search(step, input_ids.shape[0], max_new_tokens, eos_token_id, pad_token_id)


@torch.no_grad()
def multiround_loop(sd, cfg, ids, pix, cont_ids, forced_prefill=None, forced_turn=None, record=None):
    """prefill -> 3 greedy steps -> multi-token continuation (new user turn) -> 3 greedy steps: the call pattern of
    llava/dynamic_eval/model_lvis_multi_round_for_meteor.py:150-185 (mirrors oracle/gen_golden.py:multiround_case and
    :instruct_case; `forced_*` replace the instruct predictor's decisions, `record` gets its logits per call)."""
    logits, lens, toks = [], [], []
    r0, r1 = {}, {}
    lg, past = model_forward(sd, cfg, ids, images=pix, forced_instruct=forced_prefill, record=r0)
    for t in range(3):
        logits.append(lg[:, -1]); lens.append(torch.stack(list(past[1]), 0).clone())
        nxt = lg[:, -1].argmax(-1).reshape(1, 1); toks.append(int(nxt))
        lg, past = model_forward(sd, cfg, nxt, past=past)
    logits.append(lg[:, -1]); lens.append(torch.stack(list(past[1]), 0).clone())
    cont_lg, past = model_forward(sd, cfg, cont_ids, past=past, forced_instruct=forced_turn, record=r1)
    if record is not None:
        record.update(instruct_logits_prefill=r0.get("instruct_logits"), instruct_logits_turn=r1.get("instruct_logits"))
    lens.append(torch.stack(list(past[1]), 0).clone())
    lg = cont_lg
    for t in range(3):
        nxt = lg[:, -1].argmax(-1).reshape(1, 1); toks.append(int(nxt))
        lg, past = model_forward(sd, cfg, nxt, past=past)
        logits.append(lg[:, -1]); lens.append(torch.stack(list(past[1]), 0).clone())
    return torch.stack(logits, 0), cont_lg, torch.stack(lens, 0), toks, past


@torch.no_grad()
def nocache_loop(sd, cfg, ids, pix, n_calls, forced_output_logits=None, forced_keep_index=None, teacher_tokens=None):
    """llava/dynamic_eval/bench_test/dynamic_llava_long_text_time_with_no_cache.py:336-349: `model(total_ids, images, use_cache=False)`
    with the id sequence growing by one token per call (greedy argmax, or `teacher_tokens` [B, n_calls], in place of the script's label
    ids).  Returns per-call last-position logits, output lengths, the sparse layer's inputs and the final call's full logits."""
    state = {"answer_indice": None}
    total = ids.clone()
    B = ids.shape[0]
    last, lens, sl_in, text_logits, keep_o, full = [], [], [], [], [], None
    SL = cfg["sparse_config"]["sparse_layer"]
    for t in range(n_calls):
        rec = {}
        full, past = model_forward(sd, cfg, total, images=pix, use_cache=False, state=state, record=rec,
                                   forced_output_logits=forced_output_logits, forced_keep_index=forced_keep_index)
        assert past is None
        last.append(full[:, -1].clone())
        lens.append(full.shape[1])
        sl_in.append(rec["layer_inputs"][SL])
        text_logits.append(rec["text_logits"])
        keep_o.append(rec["output_keep_index"])
        nxt = full[:, -1].argmax(-1) if teacher_tokens is None else teacher_tokens[:, t]
        total = torch.cat([total, nxt.reshape(B, 1)], dim=1)
    return dict(tokens=total[:, ids.shape[1]:], last_logits=torch.stack(last, 0), seq_lens=lens, final_logits=full,
                sparse_layer_inputs=sl_in, text_logits=text_logits, output_keep_index=keep_o, answer_indice=state["answer_indice"])
