"""Seeded inputs of the prepare_inputs_labels_for_multimodal fixtures, shared by oracle/gen_golden.py (which runs the reference on
them) and the tests (which run the oracle and the CUDA path on them).  Test infrastructure, like everything under oracle/."""
import torch

from dynamic_llava_b200.synthetic import synthetic_inputs

PREPARE_CASES = ("base", "truncate", "image_first", "image_last", "text_only_row", "text_only_row_left")


def prepare_case_inputs(cfg, name):
    """Inputs of the prepare_inputs_labels_for_multimodal fixtures (shared with tests/test_oracle_golden.py)."""
    ids, pix = synthetic_inputs(cfg, 3, 12, seed=1, dtype=torch.bfloat16)
    kw = {}
    if name == "truncate":  # tokenizer_model_max_length inside the instruct span (dynamic_llava_arch.py:493-506)
        kw["max_len"] = 20
    elif name == "image_first":  # placeholder right after BOS: a one-token system span
        ids[:, 5] = ids[:, 1]
        ids[:, 1] = -200
    elif name == "image_last":  # placeholder is the last id: empty instruct span
        ids[:, 5] = ids[:, 11]
        ids[:, 11] = -200
    elif name == "text_only_row":  # one sample without a placeholder: shorter row, right padding, no dict for it (:315-324)
        ids[1, 5] = 7
    elif name == "text_only_row_left":  # same with tokenizer_padding_side = "left" (:509-579)
        ids[1, 5] = 7
        kw["padding_side"] = "left"
    return ids, pix, kw


# ---------------------------------------------------------------------------------------------------------------------
# host-prep fixtures (prompt templates, tokenizer_image_token, KeywordsStoppingCriteria, process_images)
# ---------------------------------------------------------------------------------------------------------------------
class ToyTokenizer:
    """whitespace tokenizer that prepends BOS like Llama's sentencepiece tokenizer does; `no_bos=True` drops it (the other branch of
    tokenizer_image_token, mm_utils.py:211-218).  ids: 3 + (sum of code points) % 500, a few named words for the stopping criteria."""
    bos_token_id = 1
    named = {"</s>": 2, "###": 7, "Human": 8, "stop": 9, "now": 10}

    def __init__(self, no_bos=False):
        self.no_bos = no_bos

    def _id(self, w):
        return self.named.get(w, 11 + (sum(map(ord, w)) % 500))

    def __call__(self, text):
        class R:
            pass
        r = R()
        r.input_ids = ([] if self.no_bos else [1]) + [self._id(w) for w in text.split()]
        return r

    def batch_decode(self, ids, skip_special_tokens=True):
        inv = {v: k for k, v in self.named.items()}
        return [" ".join(inv.get(int(t), "x") for t in row) for row in ids]


HOST_PROMPTS = [
    ("vicuna_v1", [("<image>\nWhat is shown?", None)]),
    ("vicuna_v1", [("<image>\nWhat is shown?", "A cat."), ("And what colour is it?", None)]),
    ("llava_v1", [("Describe the picture <image> in detail.", "Two dogs."), ("How many?", "Two.")]),
    ("plain", [("<image>", None)]),
    ("plain", [("<image>", "a photo of a cat")]),
    ("v1", [("no image here", None)]),
]
TOKENIZE_PROMPTS = ["a <image> b <image> c", "<image> leads", "trails <image>", "no placeholder at all", "<image>"]
STOP_KEYWORDS = ["stop now", "###"]
STOP_SEQUENCES = [[[5, 9]], [[5, 9, 10]], [[5, 6, 7]], [[5, 6, 7], [5, 6, 6]], [[5, 6, 7], [9, 10, 7]], [[7, 5, 6]], [[8, 8, 8, 9, 10]],
                  [[9, 10, 5, 5, 5]]]
MODEL_PATHS = ["/x/dynamic-llava-7b/checkpoint-100/", "results/dynamic-llava-7b", "a/b/c/", "checkpoint-5", "m/checkpoint-7"]


def host_image():
    """a deterministic non-square RGB image with structure in both directions"""
    import numpy as np
    from PIL import Image
    y, x = np.mgrid[0:40, 0:90]
    arr = np.stack([(3 * x + y) % 256, (x * y) % 256, (7 * y + 2 * x) % 256], axis=-1).astype("uint8")
    return Image.fromarray(arr, "RGB")
