"""Same import path as the reference module (llava/model/language_model/dynamic_llava_llama.py:39-173)."""
from dynamic_llava_b200.modeling import (  # noqa: F401
    DynamicLlavaConfig,
    DynamicLlavaLlamaForCausalLM,
    DynamicLlavaLlamaModel,
    LlavaConfig,
    LlavaLlamaForCausalLM,
)

from dynamic_llava_b200.modeling import register_auto_classes  # noqa: E402

register_auto_classes()  # AutoConfig / AutoModelForCausalLM registration at import time (reference: dynamic_llava_llama.py:172-173)
