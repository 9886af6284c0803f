"""Host-side prompt / image preparation the reference's eval drivers call before `model.generate`
(behaviour of llava/mm_utils.py:163-236 and llava/conversation.py:60-70, :267-278, :394-408).

Plain Python on the host — nothing here is on the GPU hot path.  Scope: LLaVA-1.5 single-image prompts
(`image_aspect_ratio` "pad" or default); "anyres" multi-crop inputs belong to LLaVA-1.6 and raise NotImplementedError.
"""
from __future__ import annotations

import copy
import dataclasses
from enum import Enum, auto
from typing import List, Optional, Sequence

import torch

IMAGE_TOKEN_INDEX = -200
DEFAULT_IMAGE_TOKEN = "<image>"


# ---------------------------------------------------------------------------------------------------------------- images
def expand2square(pil_img, background_color):
    """Centre the image on a square canvas of side max(w, h) filled with `background_color` (mm_utils.py:163-174)."""
    from PIL import Image
    w, h = pil_img.size
    if w == h:
        return pil_img
    side = max(w, h)
    canvas = Image.new(pil_img.mode, (side, side), background_color)
    canvas.paste(pil_img, ((side - w) // 2, (side - h) // 2))
    return canvas


def process_images(images, image_processor, model_cfg):
    """mm_utils.py:177-199: "pad" -> pad to square with the processor's mean colour, then CLIP-preprocess each image and
    stack; default -> the processor's own batch call."""
    mode = getattr(model_cfg, "image_aspect_ratio", None)
    if mode == "anyres":
        raise NotImplementedError("anyres (LLaVA-1.6 multi-crop) inputs are out of scope")
    if mode != "pad":
        return image_processor(images, return_tensors="pt")["pixel_values"]
    fill = tuple(int(c * 255) for c in image_processor.image_mean)
    out = [image_processor.preprocess(expand2square(im, fill), return_tensors="pt")["pixel_values"][0] for im in images]
    return torch.stack(out, dim=0) if all(t.shape == out[0].shape for t in out) else out


# ---------------------------------------------------------------------------------------------------------------- prompts
def tokenizer_image_token(prompt: str, tokenizer, image_token_index: int = IMAGE_TOKEN_INDEX, return_tensors: Optional[str] = None):
    """Tokenise the text around every "<image>" marker separately and put `image_token_index` where the marker was; a BOS
    the tokenizer prepends to each piece is kept once, at the very front (mm_utils.py:202-227)."""
    pieces = [tokenizer(p).input_ids for p in prompt.split(DEFAULT_IMAGE_TOKEN)]
    bos = getattr(tokenizer, "bos_token_id", None)
    has_bos = bool(pieces) and len(pieces[0]) > 0 and pieces[0][0] == bos
    ids: List[int] = [pieces[0][0]] if has_bos else []
    for n, piece in enumerate(pieces):
        if n > 0:
            ids.append(image_token_index)
        ids.extend(piece[1:] if has_bos else piece)
    if return_tensors is None:
        return ids
    if return_tensors == "pt":
        return torch.tensor(ids, dtype=torch.long)
    raise ValueError(f"Unsupported tensor type: {return_tensors}")


def get_model_name_from_path(model_path: str) -> str:
    """mm_utils.py:230-236: last path component, prefixed by its parent for `checkpoint-*` directories."""
    parts = model_path.strip("/").split("/")
    return parts[-2] + "_" + parts[-1] if parts[-1].startswith("checkpoint-") else parts[-1]


class SeparatorStyle(Enum):
    SINGLE = auto()
    TWO = auto()
    PLAIN = auto()


@dataclasses.dataclass
class Conversation:
    """Prompt template state (conversation.py:18-70): `get_prompt()` renders system + alternating role turns."""
    system: str
    roles: Sequence[str]
    messages: list
    offset: int = 0
    sep_style: SeparatorStyle = SeparatorStyle.TWO
    sep: str = " "
    sep2: Optional[str] = None
    version: str = "Unknown"

    def append_message(self, role, message):
        self.messages.append([role, message])

    def copy(self):
        return copy.deepcopy(self)

    def get_prompt(self) -> str:
        def text(m):
            return m[0] if isinstance(m, tuple) else m
        if self.sep_style == SeparatorStyle.TWO:
            ends = (self.sep, self.sep2)
            s = self.system + ends[0]
            for i, (role, msg) in enumerate(self.messages):
                s += f"{role}: {text(msg)}{ends[i % 2]}" if msg else f"{role}:"
            return s
        if self.sep_style == SeparatorStyle.SINGLE:
            s = self.system + self.sep
            for role, msg in self.messages:
                s += f"{role}: {text(msg)}{self.sep}" if msg else f"{role}:"
            return s
        if self.sep_style == SeparatorStyle.PLAIN:
            ends = (self.sep, self.sep2)
            s = self.system
            for i, (_, msg) in enumerate(self.messages):
                if msg:
                    s += text(msg) + ends[i % 2]
            return s
        raise ValueError(f"Invalid style: {self.sep_style}")


# ------------------------------------------------------------------------------------------------------- stopping criteria
class KeywordsStoppingCriteria:
    """Behaviour of mm_utils.py:239-282 (a transformers `StoppingCriteria`; only the serving worker passes one).  A sequence is
    finished when its tail equals one of the keywords' id sequences or when the decoded tail contains a keyword string; a batch
    stops when every row is finished.  `generate(stopping_criteria=[...])` evaluates it on the host after every token, on the ids
    generated so far (LLaVA hands `inputs_embeds` to HF, so `output_ids` never holds the prompt)."""

    def __init__(self, keywords, tokenizer, input_ids):
        self.keywords = list(keywords)
        self.keyword_ids, self.max_keyword_len = [], 0
        for kw in self.keywords:
            ids = list(tokenizer(kw).input_ids)
            if len(ids) > 1 and ids[0] == tokenizer.bos_token_id:
                ids = ids[1:]
            self.max_keyword_len = max(self.max_keyword_len, len(ids))
            self.keyword_ids.append(torch.tensor(ids))
        self.tokenizer = tokenizer
        self.start_len = input_ids.shape[1]

    def call_for_batch(self, output_ids, scores=None, **kwargs) -> bool:
        offset = min(output_ids.shape[1] - self.start_len, self.max_keyword_len)
        for kid in self.keyword_ids:
            if torch.equal(output_ids[0, -kid.shape[0]:].cpu(), kid):
                return True
        text = self.tokenizer.batch_decode(output_ids[:, -offset:], skip_special_tokens=True)[0]
        return any(kw in text for kw in self.keywords)

    def __call__(self, output_ids, scores=None, **kwargs) -> bool:
        return all(self.call_for_batch(output_ids[i].unsqueeze(0), scores) for i in range(output_ids.shape[0]))


_VICUNA_SYSTEM = ("A chat between a curious user and an artificial intelligence assistant. "
                  "The assistant gives helpful, detailed, and polite answers to the user's questions.")

conv_vicuna_v1 = Conversation(system=_VICUNA_SYSTEM, roles=("USER", "ASSISTANT"), messages=[], offset=0,
                              sep_style=SeparatorStyle.TWO, sep=" ", sep2="</s>", version="v1")
_LLAVA_V1_SYSTEM = ("A chat between a curious human and an artificial intelligence assistant. "
                    "The assistant gives helpful, detailed, and polite answers to the human's questions.")  # conversation.py:289-299 ("human", not "user")
conv_llava_v1 = Conversation(system=_LLAVA_V1_SYSTEM, roles=("USER", "ASSISTANT"), messages=[], offset=0,
                             sep_style=SeparatorStyle.TWO, sep=" ", sep2="</s>", version="v1")
conv_llava_plain = Conversation(system="", roles=("", ""), messages=[], offset=0, sep_style=SeparatorStyle.PLAIN, sep="\n", sep2=None)

default_conversation = conv_vicuna_v1
conv_templates = {"default": conv_vicuna_v1, "v1": conv_vicuna_v1, "vicuna_v1": conv_vicuna_v1, "llava_v1": conv_llava_v1,
                  "plain": conv_llava_plain}
