"""Same import path as the reference's llava/mm_utils.py (host-side prompt / image prep)."""
from dynamic_llava_b200.host_prep import (  # noqa: F401
    KeywordsStoppingCriteria,
    expand2square,
    get_model_name_from_path,
    process_images,
    tokenizer_image_token,
)
