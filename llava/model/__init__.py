from dynamic_llava_b200 import (  # noqa: F401
    DynamicLlavaConfig,
    DynamicLlavaLlamaForCausalLM,
    LlavaConfig,
    LlavaLlamaForCausalLM,
)
