"""Seeded synthetic weights and inputs shared by the tests, the golden generator and bench.py.

There is no network for checkpoints, so every run uses random-init weights under the reference's
state-dict key names (SURVEY.md §8b "State-dict keys"; dumped from the reference on the meta device).
Two generators:

* `numpy_state_dict` — numpy PCG64 (stream stable across numpy versions), used for the toy goldens so that the
  committed fixtures hold only outputs, not weights;
* `device_state_dict` — torch RNG on the target device, used by bench.py at 7B/13B scale where host-side
  sampling would take minutes (247 s).
"""
from __future__ import annotations

from typing import Dict, Iterator, Tuple

import numpy as np
import torch

IMAGE_TOKEN_INDEX = -200  # llava/constants.py:8


def weight_shapes(cfg: dict) -> Iterator[Tuple[str, Tuple[int, ...], str]]:
    """Yield (key, shape, kind) for every tensor of the Dynamic-LLaVA checkpoint.
    kind in {"linear", "embed", "norm_w", "bias", "cls", "pos"} selects the init distribution."""
    H, L, I, V = cfg["hidden_size"], cfg["num_hidden_layers"], cfg["intermediate_size"], cfg["vocab_size"]
    sc = cfg["sparse_config"]
    yield "model.embed_tokens.weight", (V, H), "embed"
    for i in range(L):
        p = f"model.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "o_proj"):
            yield p + f"self_attn.{n}.weight", (H, H), "linear"
        yield p + "mlp.gate_proj.weight", (I, H), "linear"
        yield p + "mlp.up_proj.weight", (I, H), "linear"
        yield p + "mlp.down_proj.weight", (H, I), "linear"
        yield p + "input_layernorm.weight", (H,), "norm_w"
        yield p + "post_attention_layernorm.weight", (H,), "norm_w"
    yield "model.norm.weight", (H,), "norm_w"
    yield "lm_head.weight", (V, H), "linear"
    # projector mlp2x_gelu (multimodal_projector/builder.py:172-179)
    M = cfg["mm_hidden_size"]
    yield "model.mm_projector.0.weight", (H, M), "linear"
    yield "model.mm_projector.0.bias", (H,), "bias"
    yield "model.mm_projector.2.weight", (H, H), "linear"
    yield "model.mm_projector.2.bias", (H,), "bias"
    # predictors (dynamic_modeling_llama.py:1308-1387)
    D, FF, NL = sc["d_model"], sc["dim_feedforward"], sc["num_layers"]
    if sc["use_vision_predictor"]:
        p = "model.image_score_predictor."
        yield p + "down_mlp.0.weight", (H,), "norm_w"
        yield p + "down_mlp.0.bias", (H,), "bias"
        yield p + "down_mlp.1.weight", (D, H), "linear"
        yield p + "down_mlp.1.bias", (D,), "bias"
        for j in range(NL):
            q = p + f"transformer.{j}."
            yield q + "norm1.weight", (D,), "norm_w"
            yield q + "norm1.bias", (D,), "bias"
            yield q + "attn.qkv.weight", (3 * D, D), "linear"
            yield q + "attn.proj.weight", (D, D), "linear"
            yield q + "attn.proj.bias", (D,), "bias"
            yield q + "norm2.weight", (D,), "norm_w"
            yield q + "norm2.bias", (D,), "bias"
            yield q + "mlp.fc1.weight", (FF, D), "linear"
            yield q + "mlp.fc1.bias", (FF,), "bias"
            yield q + "mlp.fc2.weight", (D, FF), "linear"
            yield q + "mlp.fc2.bias", (D,), "bias"
        for idx, (o, i_) in zip((0, 2, 4), ((D // 2, D), (D // 4, D // 2), (2, D // 4))):
            yield p + f"output_mlp.{idx}.weight", (o, i_), "linear"
            yield p + f"output_mlp.{idx}.bias", (o,), "bias"
    if sc["use_text_predictor"]:
        names = []
        if sc["use_output_text_predictor"]:
            names.append("output_text_score_predictor")
        if sc["use_instruct_predictor"]:
            names.append("instruct_score_predictor")
        for name in names:
            p = f"model.{name}.output_mlp."
            yield p + "0.weight", (H,), "norm_w"
            yield p + "0.bias", (H,), "bias"
            for idx, (o, i_) in zip((1, 3, 5, 7), ((D, H), (D // 2, D), (D // 4, D // 2), (2, D // 4))):
                yield p + f"{idx}.weight", (o, i_), "linear"
                yield p + f"{idx}.bias", (o,), "bias"
    # CLIP vision tower (HF names under model.vision_tower.vision_tower.)
    vc = cfg["vision"]
    C, CL, CI = vc["hidden_size"], vc["num_hidden_layers"], vc["intermediate_size"]
    ps, nch = vc["patch_size"], vc["num_channels"]
    npos = (vc["image_size"] // ps) ** 2 + 1
    p = "model.vision_tower.vision_tower.vision_model."
    yield p + "embeddings.class_embedding", (C,), "cls"
    yield p + "embeddings.patch_embedding.weight", (C, nch, ps, ps), "linear"
    yield p + "embeddings.position_embedding.weight", (npos, C), "pos"
    yield p + "pre_layrnorm.weight", (C,), "norm_w"
    yield p + "pre_layrnorm.bias", (C,), "bias"
    for i in range(CL):
        q = p + f"encoder.layers.{i}."
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            yield q + f"self_attn.{n}.weight", (C, C), "linear"
            yield q + f"self_attn.{n}.bias", (C,), "bias"
        yield q + "layer_norm1.weight", (C,), "norm_w"
        yield q + "layer_norm1.bias", (C,), "bias"
        yield q + "mlp.fc1.weight", (CI, C), "linear"
        yield q + "mlp.fc1.bias", (CI,), "bias"
        yield q + "mlp.fc2.weight", (C, CI), "linear"
        yield q + "mlp.fc2.bias", (C,), "bias"
        yield q + "layer_norm2.weight", (C,), "norm_w"
        yield q + "layer_norm2.bias", (C,), "bias"
    yield p + "post_layernorm.weight", (C,), "norm_w"
    yield p + "post_layernorm.bias", (C,), "bias"


def _fan_in(shape):
    n = 1
    for s in shape[1:]:
        n *= s
    return max(n, 1)


def numpy_state_dict(cfg: dict, seed: int = 0, dtype=torch.float32, lm_head_gain: float = 4.0) -> Dict[str, torch.Tensor]:
    """"Lively" init for toy goldens: linears N(0, 1/fan_in) so attention logits and activations are O(1)
    (N(0,0.02) at toy widths makes every softmax uniform and tests nothing), norm weights 1+0.1 N, biases 0.05 N,
    lm_head scaled by `lm_head_gain` so greedy top-1/top-2 margins are wide (SURVEY.md §7 hard parts)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = {}
    for key, shape, kind in weight_shapes(cfg):
        x = rng.standard_normal(size=shape, dtype=np.float32)
        if kind == "linear":
            x *= 1.0 / np.sqrt(_fan_in(shape))
            if key == "lm_head.weight":
                x *= lm_head_gain
        elif kind == "embed":
            x *= 1.0
        elif kind == "norm_w":
            x = 1.0 + 0.1 * x
        elif kind == "bias":
            x *= 0.05
        elif kind == "cls":
            x *= shape[0] ** -0.5
        elif kind == "pos":
            x *= 0.02
        sd[key] = torch.from_numpy(x).to(dtype)
    return sd


def device_state_dict(cfg: dict, seed: int = 0, dtype=torch.bfloat16, device="cuda", std: float = 0.02) -> Dict[str, torch.Tensor]:
    """Reference-style init (N(0, 0.02) linears/embeddings, zero biases, unit norm weights:
    dynamic_modeling_llama.py:1492-1501; CLIP per HF `_init_weights`), sampled directly on `device` in `dtype`."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    sd = {}
    for key, shape, kind in weight_shapes(cfg):
        if kind in ("linear", "embed", "pos"):
            t = torch.empty(shape, dtype=dtype, device=device).normal_(0.0, std, generator=g)
        elif kind == "cls":
            t = torch.empty(shape, dtype=dtype, device=device).normal_(0.0, shape[0] ** -0.5, generator=g)
        elif kind == "norm_w":
            t = torch.ones(shape, dtype=dtype, device=device)
        else:
            t = torch.zeros(shape, dtype=dtype, device=device)
        sd[key] = t
    return sd


def tiled_state_dict(cfg: dict, seed: int = 0, dtype=torch.bfloat16, std: float = 0.02, pool_elems: int = 1 << 24) -> Dict[str, torch.Tensor]:
    """CPU-side init for the timed CPU baseline only: one N(0,std) pool of `pool_elems` values is tiled into every
    large tensor (memcpy-bound, seconds instead of minutes). Values repeat, so this is for TIMING, never for parity."""
    g = torch.Generator()
    g.manual_seed(seed)
    pool = torch.empty(pool_elems, dtype=torch.float32).normal_(0.0, std, generator=g).to(dtype)
    sd = {}
    for key, shape, kind in weight_shapes(cfg):
        n = 1
        for s in shape:
            n *= s
        if kind in ("linear", "embed", "pos", "cls"):
            reps = (n + pool_elems - 1) // pool_elems
            t = pool.repeat(reps)[:n].reshape(shape).contiguous() if reps > 1 else pool[:n].clone().reshape(shape)
        elif kind == "norm_w":
            t = torch.ones(shape, dtype=dtype)
        else:
            t = torch.zeros(shape, dtype=dtype)
        sd[key] = t
    return sd


def synthetic_inputs(cfg: dict, batch: int, prompt_len: int, seed: int = 1, image_pos: int = 5, dtype=torch.bfloat16):
    """SURVEY.md §8d inputs: `randn` pixels, random ids in [3, V), BOS at 0, one `<image>` placeholder after a shared
    `image_pos`-token system prefix. numpy PCG64 so the stream is identical wherever it is generated."""
    rng = np.random.Generator(np.random.PCG64(seed))
    vc = cfg["vision"]
    pix = rng.standard_normal(size=(batch, vc["num_channels"], vc["image_size"], vc["image_size"]), dtype=np.float32)
    ids = rng.integers(3, cfg["vocab_size"], size=(batch, prompt_len), dtype=np.int64)
    ids[:, 0] = 1
    ids[:, image_pos] = IMAGE_TOKEN_INDEX
    return torch.from_numpy(ids), torch.from_numpy(pix).to(dtype)


def forced_text_decisions(n_steps: int, batch: int, keep_rate=0.5, seed: int = 3) -> torch.Tensor:
    """Seeded Bernoulli keep/drop stream [n_steps, batch] (`output_text_keep_rate` is never read at
    inference, so configs that name it force the decision stream identically on both sides).
    `keep_rate` may be a scalar or a per-sample sequence (config 3's ragged-KV stress)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    u = rng.random(size=(n_steps, batch))
    rate = np.broadcast_to(np.asarray(keep_rate, dtype=np.float64), (batch,))
    return torch.from_numpy(u < rate[None, :])
