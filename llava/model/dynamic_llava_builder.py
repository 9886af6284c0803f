"""Same import path as the reference (llava/model/dynamic_llava_builder.py:35)."""
from dynamic_llava_b200.builder import load_pretrained_model  # noqa: F401
