"""Drop-in import shim: `from llava.model import DynamicLlavaLlamaForCausalLM` resolves to the B200 implementation
(reference: llava/__init__.py:1, llava/model/__init__.py:2-11)."""
from .model import DynamicLlavaLlamaForCausalLM, LlavaLlamaForCausalLM  # noqa: F401
