"""dynamic_llava_b200 — B200-native (sm_100a) implementation of Dynamic-LLaVA's sparsified prefill/decode hot path.

Public surface mirrors the reference's `llava.model` exports (llava/model/__init__.py:2-11)."""
from .modeling import (  # noqa: F401
    CausalLMOutputWithPast,
    DynamicLlavaConfig,
    DynamicLlavaLlamaForCausalLM,
    DynamicLlavaLlamaModel,
    LlavaConfig,
    LlavaLlamaForCausalLM,
)
from .presets import PRESETS, SPARSE_DEFAULTS, preset  # noqa: F401

__version__ = "0.1.0"
