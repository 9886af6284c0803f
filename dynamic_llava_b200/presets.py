"""Model-shape presets and the reference's sparse_config defaults.

Shapes for LLaVA-1.5-7B / 13B come from the public `llava-v1.5-{7b,13b}/config.json`
(the reference only reads them off `config`: dynamic_modeling_llama.py:361-368, :1231);
CLIP-ViT-L/14-336 shapes from HF `CLIPVisionConfig` for `openai/clip-vit-large-patch14-336`.
`SPARSE_DEFAULTS` restates `SparseArguments` (llava/train/train_sparse.py:145-165), which the
reference persists as `config.sparse_config` (train_sparse.py:1007-1008).

The `toy` preset keeps every kernel-relevant constant of the real models (Llama head_dim 128,
CLIP / predictor head_dim 64, patch 14) at sizes a CPU oracle finishes in milliseconds.
"""
from __future__ import annotations

import copy

SPARSE_DEFAULTS = {
    "use_vision_predictor": True,
    "vision_keep_rate": 0.2,
    "use_text_predictor": True,
    "use_output_text_predictor": True,
    "output_text_keep_rate": 0.5,
    "output_text_len_for_training": 50,
    "use_instruct_predictor": True,
    "instruct_keep_rate": 0.7,
    "instruct_len_for_training": 25,
    "sparse_layer": 2,
    "d_model": 512,
    "nhead": 8,
    "dim_feedforward": 2048,
    "num_layers": 2,
    "mask_loss_weight": 100.0,
}

_CLIP_L_336 = dict(
    hidden_size=1024,
    num_hidden_layers=24,
    num_attention_heads=16,
    intermediate_size=4096,
    patch_size=14,
    image_size=336,
    layer_norm_eps=1e-5,
    num_channels=3,
)

PRESETS = {
    "llava-1.5-7b": dict(
        hidden_size=4096,
        num_hidden_layers=32,
        num_attention_heads=32,
        intermediate_size=11008,
        vocab_size=32000,
        rms_norm_eps=1e-5,
        rope_theta=10000.0,
        max_position_embeddings=4096,
        mm_hidden_size=1024,
        mm_vision_select_layer=-2,
        vision=_CLIP_L_336,
        sparse_config=dict(SPARSE_DEFAULTS, use_instruct_predictor=False),
    ),
    "llava-1.5-13b": dict(
        hidden_size=5120,
        num_hidden_layers=40,
        num_attention_heads=40,
        intermediate_size=13824,
        vocab_size=32000,
        rms_norm_eps=1e-5,
        rope_theta=10000.0,
        max_position_embeddings=4096,
        mm_hidden_size=1024,
        mm_vision_select_layer=-2,
        vision=_CLIP_L_336,
        sparse_config=dict(SPARSE_DEFAULTS, use_instruct_predictor=False),
    ),
    # hidden 256 = 2 heads x 128; CLIP 128 = 2 heads x 64; 56x56 image -> 16 patches.
    "toy": dict(
        hidden_size=256,
        num_hidden_layers=4,
        num_attention_heads=2,
        intermediate_size=512,
        vocab_size=512,
        rms_norm_eps=1e-5,
        rope_theta=10000.0,
        max_position_embeddings=4096,
        mm_hidden_size=128,
        mm_vision_select_layer=-2,
        vision=dict(
            hidden_size=128,
            num_hidden_layers=3,
            num_attention_heads=2,
            intermediate_size=256,
            patch_size=14,
            image_size=56,
            layer_norm_eps=1e-5,
            num_channels=3,
        ),
        sparse_config=dict(
            SPARSE_DEFAULTS,
            use_instruct_predictor=False,
            vision_keep_rate=0.5,
            d_model=128,
            nhead=2,
            dim_feedforward=256,
        ),
    ),
}


def preset(name: str) -> dict:
    """Deep copy of a preset (callers mutate `sparse_config` like the reference's bench scripts do,
    llava/dynamic_eval/bench_test/dynamic_llava_image_time_and_mem.py:63-65)."""
    return copy.deepcopy(PRESETS[name])
