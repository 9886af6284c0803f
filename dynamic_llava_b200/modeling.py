"""Drop-in model API: `DynamicLlavaConfig`, `DynamicLlavaLlamaForCausalLM` with the reference's call surface
(llava/model/language_model/dynamic_llava_llama.py:39-173) on top of the B200 engine.

Kept from the reference: `forward(...)` signature and `CausalLMOutputWithPast`-shaped result with fp32 logits and a
legacy-tuple `past_key_values`; `generate(inputs, images, image_sizes, **kwargs)` (greedy and HF `sample`); the
`prepare_inputs_labels_for_multimodal` return shape; `model.model.image_score_predictor` & friends as hookable
modules; the live, mutable `config.sparse_config` dict that callers edit at run time
(llava/dynamic_eval/bench_test/dynamic_llava_image_time_and_mem.py:63-65); the state-dict key names.

Extra (not in the reference) keyword arguments, all optional: `forced_keep_index`, `forced_text_decisions`
(parity "under the same predictor masks"), and `config.reference_batched_quirk`.
"""
from __future__ import annotations

import copy
import weakref
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.nn as nn

from .cache import LegacyPastView
from .engine import IMAGE_TOKEN_INDEX, Engine, Session
from .modules import CLIPVisionTower, DecoderLayer, Embedding, Linear, Norm, Act, TextPredictor, VisionPredictor
from .presets import SPARSE_DEFAULTS, preset


class DynamicLlavaConfig:
    """Attribute bag mirroring `DynamicLlavaConfig(LlamaConfig)` (dynamic_llava_llama.py:39-40)."""

    model_type = "dynamic_llava_llama"

    def __init__(self, **kw):
        d = dict(hidden_size=4096, intermediate_size=11008, num_hidden_layers=32, num_attention_heads=32,
                 num_key_value_heads=None, vocab_size=32000, rms_norm_eps=1e-5, rope_theta=10000.0, rope_scaling=None,
                 max_position_embeddings=4096, pretraining_tp=1, use_cache=True, pad_token_id=0, bos_token_id=1,
                 eos_token_id=2, hidden_act="silu", mm_vision_tower="synthetic-clip", mm_projector_type="mlp2x_gelu",
                 mm_hidden_size=1024, mm_vision_select_layer=-2, mm_vision_select_feature="patch",
                 mm_patch_merge_type="flat", mm_use_im_start_end=False, mm_use_im_patch_token=False,
                 image_aspect_ratio="square", tokenizer_padding_side="right", tokenizer_model_max_length=None,
                 vision=None, sparse_config=None, reference_batched_quirk=False, output_attentions=False,
                 output_hidden_states=False, use_return_dict=True)
        d.update(kw)
        if d["sparse_config"] is None:
            d["sparse_config"] = copy.deepcopy(SPARSE_DEFAULTS)
        if d["vision"] is None:
            d["vision"] = preset("llava-1.5-7b")["vision"]
        if d["num_key_value_heads"] is None:
            d["num_key_value_heads"] = d["num_attention_heads"]
        self.__dict__.update(d)

    @classmethod
    def from_preset(cls, name: str, **overrides):
        p = preset(name)
        p.update(overrides)
        return cls(**p)

    def to_dict(self):
        return copy.deepcopy(self.__dict__)


@dataclass
class CausalLMOutputWithPast:
    """Same fields as transformers.modeling_outputs.CausalLMOutputWithPast (dynamic_modeling_llama.py:2807-2813)."""
    loss: Optional[torch.Tensor] = None
    logits: Optional[torch.Tensor] = None
    past_key_values: Optional[object] = None
    hidden_states: Optional[tuple] = None
    attentions: Optional[tuple] = None

    def __getitem__(self, i):
        return [v for v in (self.loss, self.logits, self.past_key_values, self.hidden_states, self.attentions) if v is not None][i]


@dataclass
class GenerateOutput:
    sequences: torch.Tensor


class DynamicLlavaLlamaModel(nn.Module):
    """Weights of DynamicLlavaLlamaModel = DynamicLlavaMetaModel + DynamicLlamaModel
    (dynamic_llava_arch.py:39-57, dynamic_modeling_llama.py:1586-1647)."""

    def __init__(self, config: DynamicLlavaConfig, kw):
        super().__init__()
        self.config = config
        c, sc = config, config.sparse_config
        self.padding_idx = c.pad_token_id
        self.vocab_size = c.vocab_size
        self.embed_tokens = Embedding(c.vocab_size, c.hidden_size, **kw)
        self.layers = nn.ModuleList([DecoderLayer(c, kw) for _ in range(c.num_hidden_layers)])
        self.norm = Norm(c.hidden_size, eps=c.rms_norm_eps, **kw)
        if sc["use_vision_predictor"]:
            self.image_score_predictor = VisionPredictor(c.hidden_size, sc, kw)
        if sc["use_text_predictor"]:
            if sc["use_output_text_predictor"]:
                self.output_text_score_predictor = TextPredictor(c.hidden_size, sc, kw)
            if sc["use_instruct_predictor"]:
                self.instruct_score_predictor = TextPredictor(c.hidden_size, sc, kw)
        self.gumbel_tau = 1.0
        self.answer_indice = None
        if getattr(c, "mm_vision_tower", None) is not None:
            self.vision_tower = CLIPVisionTower(c.vision, c.mm_vision_select_layer, c.mm_vision_select_feature, kw, name=c.mm_vision_tower)
            if c.mm_projector_type != "mlp2x_gelu":
                raise ValueError(f"Unknown projector type: {c.mm_projector_type}")  # multimodal_projector/builder.py:184
            self.mm_projector = nn.Sequential(Linear(c.mm_hidden_size, c.hidden_size, True, **kw), Act(),
                                              Linear(c.hidden_size, c.hidden_size, True, **kw))

    def get_vision_tower(self):
        return getattr(self, "vision_tower", None)

    def get_input_embeddings(self):
        return self.embed_tokens


class DynamicLlavaLlamaForCausalLM(nn.Module):
    config_class = DynamicLlavaConfig

    def __init__(self, config: DynamicLlavaConfig, device="cuda", dtype=torch.bfloat16):
        super().__init__()
        self.config = config
        kw = dict(device=device, dtype=dtype)
        self.model = DynamicLlavaLlamaModel(config, kw)
        self.pretraining_tp = config.pretraining_tp
        self.vocab_size = config.vocab_size
        self.lm_head = Linear(config.hidden_size, config.vocab_size, False, **kw)
        self._engine = None
        self._sessions = {}
        self.eval()

    # ---- reference attribute surface ---------------------------------------------------------------------------
    def get_model(self):
        return self.model

    def get_vision_tower(self):
        return self.get_model().get_vision_tower()

    @property
    def device(self):
        return self.lm_head.weight.device

    @property
    def dtype(self):
        return self.lm_head.weight.dtype

    def get_input_embeddings(self):
        return self.model.embed_tokens

    def get_output_embeddings(self):
        return self.lm_head

    def resize_token_embeddings(self, n: int):
        """dynamic_llava_builder.py:235: grow/shrink embed_tokens and lm_head rows (new rows zero).
        The weight-streaming and tile GEMMs address rows in 16-byte units: a vocabulary that is not a multiple of 8 is refused by name
        (LLaVA-1.5 checkpoints keep the 32000-entry Vicuna vocabulary: `mm_use_im_patch_token` / `mm_use_im_start_end` are false)."""
        if n % 8 != 0:
            raise NotImplementedError(f"vocab size {n} is not a multiple of 8 (TMA row stride): pad the tokenizer, e.g. "
                                      f"tokenizer.add_tokens([...]) up to {-(-n // 8) * 8} entries, before resize_token_embeddings")
        for holder in (self.model.embed_tokens, self.lm_head):
            old = holder.weight.data
            new = torch.zeros(n, old.shape[1], device=old.device, dtype=old.dtype)
            new[: min(n, old.shape[0])] = old[: min(n, old.shape[0])]
            holder.weight = nn.Parameter(new, requires_grad=False)
        self.config.vocab_size = self.vocab_size = self.model.vocab_size = n
        self._sessions.clear()
        return self.model.embed_tokens

    @classmethod
    def from_state_dict(cls, config: DynamicLlavaConfig, state_dict, device="cuda", dtype=torch.bfloat16):
        """Build on `device` and load a reference-format checkpoint state dict (keys: SURVEY.md §8b)."""
        model = cls(config, device=device, dtype=dtype)
        model.get_vision_tower().load_model()
        missing, unexpected = model.load_state_dict(state_dict, strict=False)
        if unexpected or missing:
            raise RuntimeError(f"state_dict mismatch: missing={missing[:5]} unexpected={unexpected[:5]}")
        return model

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, *model_args, config=None, torch_dtype=None, dtype=None, device_map=None,
                        low_cpu_mem_usage=None, **kwargs):
        """`DynamicLlavaLlamaForCausalLM.from_pretrained(path, low_cpu_mem_usage=True, device_map="auto", torch_dtype=...)` as the
        reference's builder calls it (dynamic_llava_builder.py:191-193), also reached through `AutoModelForCausalLM.from_pretrained`
        once `register_auto_classes()` ran (dynamic_llava_llama.py:172-173).  Local HF-format directories only (no hub access)."""
        from .builder import load_model_from_dir
        dev = "cuda"
        if isinstance(device_map, (str, torch.device)) and str(device_map) not in ("auto", "balanced", "sequential"):
            dev = device_map
        elif isinstance(device_map, dict):
            raise NotImplementedError("per-module device maps (accelerate layer placement) are out of scope: one replica per GPU")
        return load_model_from_dir(str(pretrained_model_name_or_path), device=dev, torch_dtype=torch_dtype if torch_dtype is not None else dtype)

    # ---- engine plumbing ---------------------------------------------------------------------------------------
    @property
    def engine(self) -> Engine:
        if self._engine is None:
            self._engine = Engine(self)
            ref = weakref.ref(self._engine)
            for name in ("image_score_predictor", "output_text_score_predictor", "instruct_score_predictor"):
                mod = getattr(self.model, name, None)
                if mod is not None:
                    mod._engine = ref
            vt = self.get_vision_tower()
            if vt is not None:
                vt._engine = ref
        return self._engine

    def _session(self, B, cap_lo, cap_hi, max_new, scratch=False, keep_slabs=False) -> Session:
        # sparse_layer fixes the arena's lo/hi split: a caller that mutates the live sparse_config gets a new arena, not a stale split
        key = (B, cap_lo, cap_hi, max_new, scratch, self.config.sparse_config["sparse_layer"])
        s = self._sessions.get(key)
        if s is not None and s.view_ref is not None and s.view_ref() is not None:
            # a past_key_values handed out by an earlier forward() is still alive in the caller: the reference returns an independent
            # cache per call, so this conversation gets its own arena (the old one lives on through the view that references it)
            s = None
        if s is None:
            self._sessions.clear()  # one session cached per model: arenas are large
            s = Session(self.engine, B, cap_lo, cap_hi, max_new)
            self._sessions[key] = s
        else:
            # keep_slabs: generate() never hands the cache out, and no kernel reads a slot at or beyond kv_len, so only the lengths
            # are cleared (re-zeroing 1.3 GB of slabs per call is 0.2 ms at 7B, B = 8); forward() returns zero-padded legacy views
            # and therefore starts from zeroed slabs
            s.reset(lengths_only=scratch or keep_slabs)
        return s

    def _past_view(self, sess) -> LegacyPastView:
        v = LegacyPastView(sess.arena, sess)
        sess.view_ref = weakref.ref(v)
        return v

    def encode_images(self, images):
        """dynamic_llava_arch.py:163-166"""
        return self.engine.encode_images(images)

    # ---- prepare_inputs_labels_for_multimodal (dynamic_llava_arch.py:169-601) ----------------------------------
    def prepare_inputs_labels_for_multimodal(self, input_ids, position_ids, attention_mask, past_key_values, labels, images,
                                             image_sizes=None, ids_host=None):
        vt = self.get_vision_tower()
        if vt is None or images is None or input_ids.shape[1] == 1:  # :179-188
            return (input_ids, position_ids, attention_mask, past_key_values, None, labels), (None,)
        if labels is not None:
            raise NotImplementedError("labels (training) are out of scope of the inference hot path")
        if attention_mask is not None and not bool(torch.as_tensor(attention_mask).bool().all()):
            raise NotImplementedError("padded attention masks are not supported (reference contract: SURVEY.md fact 7)")
        if type(images) is list or images.ndim == 5:
            raise NotImplementedError("multi-image / anyres inputs are out of scope (LLaVA-1.5 'flat' single image only)")
        feats = self.encode_images(images)
        embeds, indices = self.engine.splice(input_ids, feats, max_len=getattr(self.config, "tokenizer_model_max_length", None),
                                             padding_side=getattr(self.config, "tokenizer_padding_side", "right"), ids_host=ids_host)
        return (None, position_ids, attention_mask, past_key_values, embeds, None), (indices,)

    # ---- forward (dynamic_llava_llama.py:68-115 -> dynamic_modeling_llama.py:2631-2813) ------------------------
    @torch.no_grad()
    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None, inputs_embeds=None,
                labels=None, use_cache=None, output_attentions=None, output_hidden_states=None, images=None, image_sizes=None,
                return_dict=None, input_embeds_indices=None, forced_keep_index=None, forced_text_decision=None, record=None):
        if input_ids is not None and inputs_embeds is not None:
            raise ValueError("You cannot specify both input_ids and inputs_embeds at the same time")
        if input_ids is None and inputs_embeds is None:
            raise ValueError("You have to specify either input_ids or inputs_embeds")
        if labels is not None or output_attentions or output_hidden_states:
            raise NotImplementedError("labels / output_attentions / output_hidden_states are outside the inference hot path")
        if use_cache is False and past_key_values is not None:
            raise ValueError("use_cache=False re-feeds the whole sequence; it does not take past_key_values")
        sc = self.config.sparse_config
        eng = self.engine
        if inputs_embeds is None:
            (input_ids, position_ids, attention_mask, past_key_values, inputs_embeds, labels), (input_embeds_indices,) = \
                self.prepare_inputs_labels_for_multimodal(input_ids, position_ids, attention_mask, past_key_values, labels, images, image_sizes)
        if past_key_values is None:
            # ---------------- prefill ----------------
            if inputs_embeds is None:  # text-only prompt: dense prefill (dynamic_llava_llama.py:143-144)
                eng._check_packed()
                ids = input_ids.to(self.device)
                inputs_embeds = torch.empty(*ids.shape, self.config.hidden_size, device=self.device, dtype=self.dtype)
                eng.L.embed_rows(ids.contiguous().data_ptr(), self.model.embed_tokens.weight.data_ptr(), inputs_embeds.data_ptr(),
                                 ids.numel(), self.config.hidden_size, torch.cuda.current_stream().cuda_stream)
            B, P, _ = inputs_embeds.shape
            eng.ensure_rope(P + 1)
            k_drop = 0
            if sc["use_vision_predictor"] and input_embeds_indices is not None and len(input_embeds_indices) == B:
                n_img = input_embeds_indices[0]["image"][1] - input_embeds_indices[0]["image"][0]
                k_drop = n_img - int(n_img * sc["vision_keep_rate"])
            room = 64
            if use_cache is False:
                # no-KV-cache mode (bench_test/dynamic_llava_long_text_time_with_no_cache.py:336-349): the sequence grows by
                # one id per call, so the scratch arena is sized in 256-token steps and reused; nothing is returned as a past
                cap = -(-(P + room) // 256) * 256
                sess = self._session(B, cap, cap, 0, scratch=True)
                logits = eng.prefill(sess, inputs_embeds, input_embeds_indices, forced_keep=forced_keep_index, all_logits=True,
                                     record=record, use_cache=False)
                return CausalLMOutputWithPast(logits=logits.float(), past_key_values=None)
            sess = self._session(B, P + room, P - k_drop + room, 0)
            logits = eng.prefill(sess, inputs_embeds, input_embeds_indices, forced_keep=forced_keep_index, all_logits=True, record=record)
            return CausalLMOutputWithPast(logits=logits.float(), past_key_values=self._past_view(sess))
        # ---------------- decode step ----------------
        if not isinstance(past_key_values, LegacyPastView):
            raise NotImplementedError("past_key_values must be the object returned by a previous forward() of this model")
        ids = input_ids if input_ids is not None else None
        if ids is None:
            raise NotImplementedError("continuing a cache needs input_ids (inputs_embeds continuation is not a reference flow)")
        sess = past_key_values.session
        a = sess.arena
        if past_key_values.stamp != (a.generation, a.host_len_lo):
            raise ValueError("stale past_key_values: this cache has advanced (or was reset) since the object was returned; pass back the "
                             "past_key_values of the LATEST forward() on this conversation")
        n_in = ids.shape[1]
        if eng.ensure_rope(a.host_len_lo + n_in):
            sess.graph = None
        if a.room() < n_in:
            a.grow(a.host_len_lo + n_in + 64, a.host_ub_hi + n_in + 64)
            sess.graph = None
        if n_in > 1:
            # a new user turn over the existing cache (model_lvis_multi_round_for_meteor.py:150-185)
            logits = eng.continue_prefill(sess, ids, record=record)
            return CausalLMOutputWithPast(logits=logits.float(), past_key_values=self._past_view(sess))
        sess.next_ids.copy_(ids.reshape(-1).to(self.device, torch.int64))
        use_dec = bool(sc["use_text_predictor"] and sc["use_output_text_predictor"])
        if forced_text_decision is not None:
            sess.forced = forced_text_decision.reshape(1, -1).to(self.device, torch.uint8).contiguous()
            a.step.zero_().add_(1)  # forced row index = step - 1 = 0
            use_dec = True
        else:
            sess.forced = None
        quirk = bool(getattr(self.config, "reference_batched_quirk", False)) and sess.B > 1
        keep = eng.decode_layers(sess, use_dec, quirk)
        if record is not None and keep is not None:
            record.update(text_decision=keep.clone().bool().view(-1, 1), text_logits=sess.tp_logits.clone())
        logits = sess.logits.float().view(sess.B, 1, -1).clone()
        eng.decode_advance(sess, keep)
        return CausalLMOutputWithPast(logits=logits, past_key_values=self._past_view(sess))

    # ---- generate (dynamic_llava_llama.py:117-152 -> HF 4.37.2 greedy search) -----------------
    @torch.no_grad()
    def generate(self, inputs=None, images=None, image_sizes=None, **kwargs):
        kwargs.pop("position_ids", None)
        attention_mask = kwargs.pop("attention_mask", None)
        kwargs.pop("input_embeds_indices", None)
        if "inputs_embeds" in kwargs:
            raise NotImplementedError("`inputs_embeds` is not supported")
        if kwargs.get("use_cache", True) is False:
            raise NotImplementedError("generate(use_cache=False): drive the no-KV-cache mode through forward(..., use_cache=False) like the "
                                      "reference's bench_test/dynamic_llava_long_text_time_with_no_cache.py; generate() always decodes over the cache")
        if kwargs.get("num_beams", 1) not in (1, None):
            raise NotImplementedError("beam search (num_beams > 1) is de-scoped: the reference's own _reorder_cache "
                                      "(dynamic_modeling_llama.py:2883-2895) drops the true-length list its cache is rebuilt from, so beams do "
                                      "not run there either (DESIGN.md §8)")
        # HF generate() kwargs this loop does not implement must not be silently ignored
        unsupported = [k for k in ("max_length", "repetition_penalty", "no_repeat_ngram_size", "length_penalty", "bad_words_ids",
                                   "force_words_ids", "num_beam_groups", "penalty_alpha", "typical_p", "logits_processor",
                                   "prefix_allowed_tokens_fn", "num_return_sequences", "assistant_model", "streamer")
                       if kwargs.get(k) not in (None, 1, 1.0, 0, [])]
        if unsupported:
            raise NotImplementedError(f"generate() kwargs not implemented by the B200 loop: {unsupported}")
        sampling = None  # (temperature, top_k, top_p): HF 4.37.2 `sample` (model_vqa_loader.py:169-172: do_sample = temperature > 0)
        if kwargs.get("do_sample", False):
            temperature = kwargs.get("temperature", None)
            temperature = 1.0 if temperature is None else float(temperature)
            if temperature <= 0:
                raise ValueError("`temperature` has to be a strictly positive float when do_sample=True")
            gc = getattr(self, "generation_config", None)
            top_k = kwargs.get("top_k", getattr(gc, "top_k", 50) if gc is not None else 50)  # GenerationConfig default
            top_k = 0 if top_k is None else int(top_k)
            top_p = kwargs.get("top_p", None)
            top_p = 1.0 if top_p is None else float(top_p)
            if top_k > 1024 or (top_k == 0 and top_p < 1.0):
                raise NotImplementedError("top_k must be in [1, 1024] when top_p < 1 (dl_sample_next keeps <= 1024 candidates)")
            sampling = (temperature, top_k, top_p)
        max_new = kwargs.get("max_new_tokens", 20)
        min_new = kwargs.get("min_new_tokens", 0) or 0
        eos = kwargs.get("eos_token_id", self.config.eos_token_id)
        if isinstance(eos, (list, tuple)):
            if len(eos) > 1:
                raise NotImplementedError("a list of several eos_token_id values is not implemented (one EOS id is tracked on the device)")
            eos = eos[0] if eos else None
        pad = kwargs.get("pad_token_id", self.config.pad_token_id)
        pad = 0 if pad is None else pad
        forced_keep = kwargs.get("forced_keep_index")
        forced_dec = kwargs.get("forced_text_decisions")  # bool [n_steps >= max_new-1, B]
        use_graph = kwargs.get("use_cuda_graph", True)
        # HF 4.37.2 greedy_search/sample: `if stopping_criteria(input_ids, scores): break` after every appended token, on the ids
        # generated so far (LLaVA passes inputs_embeds, so HF's input_ids start empty).  Host callables -> one D2H read per token,
        # paid only when a criterion is given (mm_utils.py:239-282 KeywordsStoppingCriteria; serve/model_worker.py is its one caller)
        criteria = list(kwargs.get("stopping_criteria") or [])
        sc = self.config.sparse_config
        eng = self.engine
        dev = self.device
        ids_host = inputs if inputs.device.type == "cpu" else None  # host ids (tokenizer output): no device read-back for the segment dicts
        inputs = inputs.to(dev, non_blocking=True) if inputs.device != dev else inputs
        if images is not None:
            images = images.to(dev, non_blocking=True) if images.device != dev else images
            (_, _, attention_mask, _, embeds, _), (indices,) = self.prepare_inputs_labels_for_multimodal(
                inputs, None, attention_mask, None, None, images, image_sizes=image_sizes, ids_host=ids_host)
        else:
            embeds, indices = None, None
        if embeds is None:
            eng._check_packed()
            embeds = torch.empty(*inputs.shape, self.config.hidden_size, device=dev, dtype=self.dtype)
            eng.L.embed_rows(inputs.contiguous().data_ptr(), self.model.embed_tokens.weight.data_ptr(), embeds.data_ptr(),
                             inputs.numel(), self.config.hidden_size, torch.cuda.current_stream().cuda_stream)
            indices = None
        B, P, _ = embeds.shape
        if eng.ensure_rope(P + max_new):
            for s_ in self._sessions.values():
                s_.graph = None  # captured steps hold the old table addresses
        k_drop = 0
        if sc["use_vision_predictor"] and indices is not None and len(indices) == B:
            n_img = indices[0]["image"][1] - indices[0]["image"][0]
            k_drop = n_img - int(n_img * sc["vision_keep_rate"])
        sess = self._session(B, P + max_new, P - k_drop + max_new, max_new, keep_slabs=True)
        a = sess.arena
        use_dec = bool(sc["use_text_predictor"] and sc["use_output_text_predictor"]) or forced_dec is not None
        if forced_dec is not None:
            assert forced_dec.shape[0] >= max_new - 1 and forced_dec.shape[1] == B
            n = min(forced_dec.shape[0], sess.forced_buf.shape[0])
            sess.forced_buf[:n].copy_(forced_dec[:n].to(dev, torch.uint8))  # fixed address: the captured graph stays valid
            sess.forced = sess.forced_buf
        quirk = bool(getattr(self.config, "reference_batched_quirk", False)) and B > 1
        eos_k = None if eos is None else int(eos)
        if sampling is not None:
            seed = kwargs.get("seed", None)  # reproducible under torch.manual_seed, like HF's torch.multinomial draw
            seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if seed is None else int(seed)
            sess.rng_seed.fill_(seed)

        def pick(lg, eos_):
            if eos_ is None and eos_k is not None:
                # HF MinNewTokensLengthLogitsProcessor: while fewer than min_new_tokens exist the EOS logit is -inf, so EOS cannot be
                # emitted mid-sequence (and finished-tracking is off)
                lg.select(-1, eos_k).fill_(float("-inf"))
            if sampling is None:
                eng.greedy(sess, lg, eos_, pad)
            else:
                eng.sample(sess, lg, eos_, pad, *sampling)
        logits = eng.prefill(sess, embeds, indices, forced_keep=forced_keep)
        pick(logits, eos_k if min_new <= 0 else None)
        eng.L.kv_advance(None, None, a.step.data_ptr(), None, 0, B, 0, 0, 0, torch.cuda.current_stream().cuda_stream)  # step += 1
        n_done = 1
        stopped = bool(criteria) and any(c(sess.tokens[:, :1].cpu(), None) for c in criteria)
        # everything the captured step bakes in as a host constant is part of its key: the attention kernels' prefill-length floors
        # (data-dependent with the instruct predictor), the sparse layer, the RoPE tables
        gkey = (use_dec, quirk, eos_k, pad, None if sess.forced is None else sess.forced.data_ptr(), sampling, a.floor_lo, a.floor_hi,
                sc["sparse_layer"], eng.packed["cos"].data_ptr())
        while n_done < max_new and not stopped:
            eos_now = eos_k if n_done >= min_new else None
            if use_graph and sess.graph is not None and sess.graph_key == gkey + (eos_now,):
                sess.graph.replay()
                a.host_len_lo += 1
                a.host_ub_hi += 1
                a.hi_exact = a.hi_exact and not use_dec
                eng.launches += eng.launches_per_step
            else:
                def step():
                    keep = eng.decode_layers(sess, use_dec, quirk)
                    pick(sess.logits, eos_now)
                    eng.decode_advance(sess, keep)
                if use_graph and n_done >= 2:
                    # the first decode step ran eagerly (function attributes, workspaces); capture the rest once
                    host = (a.host_len_lo, a.host_ub_hi, a.hi_exact)
                    before = eng.launches
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        step()
                    eng.launches_per_step = eng.launches - before
                    eng.launches = before
                    a.host_len_lo, a.host_ub_hi, a.hi_exact = host  # capture only records; nothing ran
                    sess.graph, sess.graph_key = g, gkey + (eos_now,)
                    continue
                step()
            n_done += 1
            if criteria:
                stopped = any(c(sess.tokens[:, :n_done].cpu(), None) for c in criteria)
            if eos_k is not None and n_done % 16 == 0 and n_done < max_new:
                if bool(sess.finished.all()):  # one host sync every 16 tokens
                    break
        out = sess.tokens[:, :n_done]
        if eos_k is not None:
            # HF stops as soon as every row has finished: trim trailing all-pad columns beyond the first all-finished step
            fin = (sess.tokens[:, :n_done] == eos_k).cumsum(1) > 0
            all_fin = fin.all(0)
            if bool(all_fin.any()):
                out = sess.tokens[:, : int(all_fin.float().argmax()) + 1]
        out = out.clone()
        if kwargs.get("return_dict_in_generate", False):
            return GenerateOutput(sequences=out)
        return out

    def prepare_inputs_for_generation(self, input_ids, past_key_values=None, inputs_embeds=None, **kwargs):
        """dynamic_llava_llama.py:154-169 / dynamic_modeling_llama.py:2815-2881 (kept for callers that drive their own loop)."""
        images = kwargs.pop("images", None)
        image_sizes = kwargs.pop("image_sizes", None)
        if past_key_values is not None:
            input_ids = input_ids[:, -1:]
        model_inputs = {"inputs_embeds": inputs_embeds} if (inputs_embeds is not None and past_key_values is None) else {"input_ids": input_ids}
        model_inputs.update({"position_ids": kwargs.get("position_ids"), "past_key_values": past_key_values,
                             "use_cache": kwargs.get("use_cache"), "attention_mask": kwargs.get("attention_mask"),
                             "input_embeds_indices": kwargs.get("input_embeds_indices")})
        if images is not None:
            model_inputs["images"] = images
        if image_sizes is not None:
            model_inputs["image_sizes"] = image_sizes
        return model_inputs


def register_auto_classes():
    """`AutoConfig.register("dynamic_llava_llama", DynamicLlavaConfig)` / `AutoModelForCausalLM.register(DynamicLlavaConfig,
    DynamicLlavaLlamaForCausalLM)` (dynamic_llava_llama.py:172-173).  The Auto machinery wants a `PretrainedConfig` subclass; ours is a
    plain attribute bag, so a same-named thin PretrainedConfig carries the parsed config.json and the model class re-reads the
    directory in `from_pretrained`.  Called on import of `llava.model` (as the reference registers at import time)."""
    from transformers import AutoConfig, AutoModelForCausalLM, PretrainedConfig

    class _HFConfig(PretrainedConfig):
        model_type = "dynamic_llava_llama"

    _HFConfig.__name__ = _HFConfig.__qualname__ = "DynamicLlavaConfig"
    AutoConfig.register("dynamic_llava_llama", _HFConfig, exist_ok=True)
    AutoModelForCausalLM.register(_HFConfig, DynamicLlavaLlamaForCausalLM, exist_ok=True)
    return _HFConfig


# Export-name alias only: the dense baseline class name that llava/model/__init__.py:2 also exports.
LlavaLlamaForCausalLM = DynamicLlavaLlamaForCausalLM
LlavaConfig = DynamicLlavaConfig
