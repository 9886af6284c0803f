"""`load_pretrained_model` with the reference's signature and return tuple
(llava/model/dynamic_llava_builder.py:35-45, :249): (tokenizer, model, image_processor, context_len).

Covers the plain Dynamic-LLaVA checkpoint branch (:187-193 + :224-247): HF-format directory with `config.json` (carrying
`sparse_config`, train_sparse.py:1007-1008) and `*.safetensors` / `pytorch_model*.bin` shards under the reference's
state-dict keys.  LoRA / `model_base` / 8-bit / 4-bit / MPT / Mistral branches are other model families or quantised
variants and raise NotImplementedError.  Differences, both deliberate: weights are cast to bf16 (the kernels compute in
bf16; the reference's eval scripts use fp16, :62) and nothing is fetched from a hub (offline): the CLIP tower weights come
from the checkpoint itself or from a local `mm_vision_tower` directory.
"""
from __future__ import annotations

import glob
import json
import os
import warnings

import torch

from .modeling import DynamicLlavaConfig, DynamicLlavaLlamaForCausalLM
from .presets import PRESETS

DEFAULT_IMAGE_PATCH_TOKEN = "<im_patch>"  # llava/constants.py
DEFAULT_IM_START_TOKEN = "<im_start>"
DEFAULT_IM_END_TOKEN = "<im_end>"
_CLIP_PREFIX = "model.vision_tower.vision_tower."


def _load_shards(path: str) -> dict:
    sd = {}
    st = sorted(glob.glob(os.path.join(path, "*.safetensors")))
    if st:
        from safetensors.torch import load_file
        for f in st:
            sd.update(load_file(f, device="cpu"))
        return sd
    for f in sorted(glob.glob(os.path.join(path, "pytorch_model*.bin"))):
        sd.update(torch.load(f, map_location="cpu", weights_only=True))
    if not sd:
        raise FileNotFoundError(f"no *.safetensors or pytorch_model*.bin under {path}")
    return sd


def _vision_config(cfg_json: dict) -> dict:
    tower = cfg_json.get("mm_vision_tower", "openai/clip-vit-large-patch14-336")
    local = os.path.join(tower, "config.json") if isinstance(tower, str) else None
    if local and os.path.exists(local):
        c = json.load(open(local))
        c = c.get("vision_config", c)
        return dict(hidden_size=c["hidden_size"], num_hidden_layers=c["num_hidden_layers"],
                    num_attention_heads=c["num_attention_heads"], intermediate_size=c["intermediate_size"],
                    patch_size=c["patch_size"], image_size=c["image_size"], layer_norm_eps=c.get("layer_norm_eps", 1e-5),
                    num_channels=c.get("num_channels", 3))
    if "clip-vit-large-patch14-336" in str(tower):
        return dict(PRESETS["llava-1.5-7b"]["vision"])
    raise ValueError(f"Unknown vision tower: {tower}")  # multimodal_encoder/builder.py:37


def load_model_from_dir(model_path, device="cuda", torch_dtype=None):
    """`DynamicLlavaLlamaForCausalLM.from_pretrained` for an HF-format checkpoint directory (dynamic_llava_builder.py:191-193 +
    :237-242): `config.json` with `sparse_config` + `*.safetensors` / `pytorch_model*.bin` shards under the reference's keys; the CLIP
    tower comes from the checkpoint itself or from a local `mm_vision_tower` directory (no hub access).
    `torch_dtype`: torch.bfloat16 (default) or torch.float16 (what the reference's eval drivers load,
    dynamic_llava_builder.py:62) — each has its own kernel library; any other dtype is refused."""
    dt = {None: torch.bfloat16, "auto": torch.bfloat16, "bfloat16": torch.bfloat16, torch.bfloat16: torch.bfloat16,
          "float16": torch.float16, torch.float16: torch.float16}.get(torch_dtype)
    if dt is None:
        raise NotImplementedError(f"torch_dtype={torch_dtype}: the B200 path runs in bfloat16 or float16 only")
    cfg_json = json.load(open(os.path.join(model_path, "config.json")))
    if "sparse_config" not in cfg_json:
        raise KeyError("config.json has no `sparse_config`: not a Dynamic-LLaVA checkpoint (train_sparse.py:1007-1008)")
    keys = ("hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads", "vocab_size", "rms_norm_eps",
            "rope_theta", "max_position_embeddings", "pad_token_id", "bos_token_id", "eos_token_id", "mm_vision_tower",
            "mm_projector_type", "mm_hidden_size", "mm_vision_select_layer", "mm_vision_select_feature", "mm_patch_merge_type",
            "mm_use_im_start_end", "mm_use_im_patch_token", "image_aspect_ratio", "tokenizer_padding_side",
            "tokenizer_model_max_length", "sparse_config", "max_sequence_length")
    if cfg_json.get("rope_theta") is None and isinstance(cfg_json.get("rope_parameters"), dict):
        # config.json written under transformers 5.x: rope_theta moved into rope_parameters (4.37.2 wrote it at the top level)
        rp = cfg_json["rope_parameters"]
        if rp.get("rope_type", "default") != "default":
            raise NotImplementedError(f"rope scaling ({rp.get('rope_type')}) is not used by LLaVA-1.5 and not supported")
        if rp.get("rope_theta") is not None:
            cfg_json["rope_theta"] = rp["rope_theta"]
    if cfg_json.get("rope_scaling") not in (None, {}):
        raise NotImplementedError("rope scaling is not used by LLaVA-1.5 and not supported")
    cfg = DynamicLlavaConfig(**{k: cfg_json[k] for k in keys if k in cfg_json and cfg_json[k] is not None},
                             vision=_vision_config(cfg_json))
    if cfg.num_key_value_heads != cfg.num_attention_heads:
        raise NotImplementedError("grouped-query attention is not used by LLaVA-1.5 and not supported")
    model = DynamicLlavaLlamaForCausalLM(cfg, device=device, dtype=dt)
    model.get_vision_tower().load_model()
    sd = _load_shards(model_path)
    if not any(k.startswith(_CLIP_PREFIX) for k in sd):  # tower weights live in their own directory (clip_encoder.py:36-38)
        tower_dir = cfg.mm_vision_tower
        if not os.path.isdir(str(tower_dir)):
            raise FileNotFoundError(f"CLIP tower weights are neither in the checkpoint nor in a local directory ({tower_dir}); "
                                    "there is no hub access in this environment")
        sd.update({_CLIP_PREFIX + k: v for k, v in _load_shards(tower_dir).items() if k.startswith("vision_model.")})
    own = set(model.state_dict())
    extra = [k for k in sd if k not in own and "rotary_emb" not in k and "position_ids" not in k]
    missing = [k for k in own if k not in sd]
    if missing:
        raise RuntimeError(f"checkpoint is missing {len(missing)} tensors, e.g. {missing[:4]}")
    if extra:
        warnings.warn(f"ignoring {len(extra)} unexpected tensors, e.g. {extra[:4]}")
    model.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=True)
    del sd
    return model


def load_pretrained_model(model_path, model_base, model_name, load_8bit=False, load_4bit=False, device_map="auto", device="cuda",
                          use_flash_attn=False, torch_dtype=torch.float16, **kwargs):
    """dynamic_llava_builder.py:26-243.  Like the reference (:62) the model is loaded in float16 unless told otherwise;
    pass torch_dtype=torch.bfloat16 for the bf16 kernels (the configuration bench.py measures)."""
    if load_8bit or load_4bit:
        raise NotImplementedError("bitsandbytes 8-/4-bit loading is out of scope (quantised variants)")
    if model_base is not None or "lora" in model_name.lower():
        raise NotImplementedError("LoRA / model_base checkpoints are out of scope; merge the adapter first")
    if "llava" not in model_name.lower():
        raise NotImplementedError("only Dynamic-LLaVA (Llama) checkpoints are supported; MPT/Mistral are other model families")
    dev = device if device != "cuda" or device_map in ("auto", None) else device_map
    model = load_model_from_dir(model_path, device=dev, torch_dtype=torch_dtype)
    cfg = model.config

    tokenizer = None
    if any(os.path.exists(os.path.join(model_path, f)) for f in ("tokenizer.model", "tokenizer.json", "tokenizer_config.json")):
        from transformers import AutoTokenizer
        tokenizer = AutoTokenizer.from_pretrained(model_path, use_fast=False)
        if getattr(cfg, "mm_use_im_patch_token", True):
            tokenizer.add_tokens([DEFAULT_IMAGE_PATCH_TOKEN], special_tokens=True)
        if getattr(cfg, "mm_use_im_start_end", False):
            tokenizer.add_tokens([DEFAULT_IM_START_TOKEN, DEFAULT_IM_END_TOKEN], special_tokens=True)
        if len(tokenizer) != cfg.vocab_size:
            model.resize_token_embeddings(len(tokenizer))  # dynamic_llava_builder.py:235
    else:
        warnings.warn(f"no tokenizer files under {model_path}: returning tokenizer=None")

    image_processor = None
    try:
        from transformers import CLIPImageProcessor
        tower_dir = str(cfg.mm_vision_tower)
        if os.path.isdir(tower_dir) and os.path.exists(os.path.join(tower_dir, "preprocessor_config.json")):
            image_processor = CLIPImageProcessor.from_pretrained(tower_dir)
        else:
            s = cfg.vision["image_size"]
            image_processor = CLIPImageProcessor(size={"shortest_edge": s}, crop_size={"height": s, "width": s})
    except Exception as e:  # pragma: no cover
        warnings.warn(f"no image processor: {e}")
    model.get_vision_tower().image_processor = image_processor
    context_len = getattr(cfg, "max_sequence_length", 2048) if hasattr(cfg, "max_sequence_length") else 2048
    return tokenizer, model, image_processor, context_len
