"""Same import path as the reference's llava/utils.py.  Every eval driver calls `disable_torch_init()` before loading the model
(e.g. llava/dynamic_eval/model_vqa_loader.py:119); the logging / moderation helpers of that file belong to the Gradio serving stack and are
out of scope."""


def disable_torch_init():
    """llava/utils.py:99-106: skip torch's default parameter initialisation of Linear / LayerNorm modules (the weights come from a
    checkpoint right after).  This package's modules allocate uninitialised parameters already, so only third-party modules built by
    the caller are affected."""
    import torch

    torch.nn.Linear.reset_parameters = lambda self: None
    torch.nn.LayerNorm.reset_parameters = lambda self: None
