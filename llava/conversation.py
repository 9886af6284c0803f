"""Same import path as the reference's llava/conversation.py (prompt templates used by llava/dynamic_eval/*)."""
from dynamic_llava_b200.host_prep import (  # noqa: F401
    Conversation,
    SeparatorStyle,
    conv_templates,
    default_conversation,
)
