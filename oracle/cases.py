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
