"""Same import path as the reference module (llava/model/language_model/dynamic_llava_llama.py:39-173)."""
from dynamic_llava_b200.modeling import (  # noqa: F401
    DynamicLlavaConfig,
    DynamicLlavaLlamaForCausalLM,
    DynamicLlavaLlamaModel,
)
