# importing the class module also registers it with transformers' Auto factories, as in the reference
# (llava/model/__init__.py:8-11 -> language_model/dynamic_llava_llama.py:172-173)
from .language_model.dynamic_llava_llama import (  # noqa: F401
    DynamicLlavaConfig,
    DynamicLlavaLlamaForCausalLM,
    LlavaConfig,
    LlavaLlamaForCausalLM,
)
