"""Execution engine: drives the hand-written sm_100a kernels (through the C ABI in include/dynllava_b200.h) for
Dynamic-LLaVA's sparsified prefill and decode.

PyTorch is used here for device memory, streams, CUDA graphs and the plain dense GEMMs (`F.linear` -> cuBLASLt) of the
prefill / CLIP / predictor blocks; everything else on the path — splice, norms, RoPE, KV append, selection + compaction,
all three attention flavours, the decode-step weight streaming (TMA + tcgen05), SwiGLU, greedy — is a `dl_*` kernel.
There is no CPU or eager fallback: if the extension is missing or the device is not sm_100, construction raises.

Reference call stack replaced: DynamicLlavaLlamaForCausalLM.forward/generate ->
prepare_inputs_labels_for_multimodal -> DynamicLlamaModel.forward (layer loop + predictor branches) -> lm_head.
"""
from __future__ import annotations

import ctypes
import math
import os
import weakref
from typing import List, Optional

import torch
import torch.nn.functional as F

from ._lib import lib
from .cache import LegacyPastView, RaggedKVArena

IMAGE_TOKEN_INDEX = -200  # llava/constants.py:8
USER_IDS = [11889, 29901]  # dynamic_llava_arch.py:36


def _p(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class Session:
    """All per-conversation device state: the ragged KV arena plus the fixed-address buffers of the decode step (so the
    step can be captured once as a CUDA graph and replayed)."""

    def __init__(self, eng: "Engine", B: int, cap_lo: int, cap_hi: int, max_new: int):
        c = eng.cfg
        dev, dt = eng.device, eng.dtype
        H, I, V, nh = c.hidden_size, c.intermediate_size, c.vocab_size, c.num_attention_heads
        self.B, self.max_new = B, max_new
        self.arena = RaggedKVArena(c.num_hidden_layers, c.sparse_config["sparse_layer"], B, nh, H // nh, cap_lo, cap_hi, dev, dt)
        e = lambda *s, dtype=dt: torch.empty(*s, device=dev, dtype=dtype)
        self.xres, self.xn, self.q, self.attn = e(B, H), e(B, H), e(B, H), e(B, H)
        self.k_side, self.v_side = e(B, nh, H // nh), e(B, nh, H // nh)
        self.act = e(B, I)
        self.logits = e(B, V)
        self.next_ids = torch.zeros(B, device=dev, dtype=torch.int64)
        self.tokens = torch.zeros(B, max(max_new, 1), device=dev, dtype=torch.int64)
        self.finished = torch.zeros(B, device=dev, dtype=torch.int32)
        self.partial = e(eng.max_partial_elems(B), dtype=torch.float32)
        D = c.sparse_config["d_model"]
        self.tp_ln = e(B, H)
        self.tp_logits = e(B, 2)
        self.tp_partial = e(8 * B * D, dtype=torch.float32)  # split-K slices of the text predictor's first Linear (own buffer: it
        # runs on a side stream next to the QKV GEMM, which owns `partial`)
        self.rng_seed = torch.zeros(1, device=dev, dtype=torch.int64)  # per-generate() seed of dl_sample_next (device: graph-replayable)
        self.forced_buf = torch.zeros(max(max_new, 1), B, device=dev, dtype=torch.uint8)
        self.forced = None  # uint8 [n_steps, B] (forced_buf when a forced decision stream is in use)
        self.attn_ws = None
        self.graph = None
        self.graph_key = None
        self.view_ref = None  # weakref to the last LegacyPastView handed out for this session (modeling._session)
        self.indices = None
        self.n_steps_done = 0

    def reset(self, lengths_only: bool = False):
        self.arena.reset(lengths_only)
        self.finished.zero_()
        self.indices = None
        self.n_steps_done = 0
        self.forced = None


class Engine:
    def __init__(self, model):
        if not torch.cuda.is_available():
            raise RuntimeError("dynamic_llava_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self._model = weakref.ref(model)
        self.cfg = model.config
        self.device = model.lm_head.weight.device
        self.dtype = model.lm_head.weight.dtype
        if self.dtype not in (torch.bfloat16, torch.float16):
            raise RuntimeError("kernels compute in bf16 or fp16: load the model with torch_dtype=torch.bfloat16 / torch.float16")
        self.L = lib(self.dtype)  # libdynllava_b200.so (bf16) or libdynllava_b200_f16.so (fp16): same symbols
        if not self.L.device_is_sm100():
            raise RuntimeError("dynamic_llava_b200 kernels are built for sm_100a (B200) only")
        self.packed = None
        self.side_stream = torch.cuda.Stream(device=self.device)  # text-predictor branch of the decode step
        self._cu_cache = {}
        self.attn_flags = int(os.environ.get("DLB200_ATTN_FLAGS", "0"))  # dl_attn_prefill_tc4 flags (bit 0: FMA-pipe exp2)
        self.skinny_target_ctas = int(os.environ.get("DLB200_SKINNY_CTAS", "296"))
        self.l2_prefetch_mb = int(os.environ.get("DLB200_L2_PREFETCH_MB", "-1"))
        self.attn_early_kv = os.environ.get("DLB200_ATTN_EARLY_KV", "1") == "1"  # fetch prefill-time K/V tiles before the PDL wait
        self.attn_prefetch_wo = os.environ.get("DLB200_ATTN_PREFETCH_WO", "0") == "1"  # o_proj weight -> L2 during decode attention (measured: no gain)
        # stream-K workspace of dl_linear (parked partial accumulators + flags): zeroed once, used by the main stream's launches only
        self.gemm_ws = torch.zeros(self.L.linear_workspace_bytes(), device=self.device, dtype=torch.uint8)
        self.launches = 0  # kernels of OURS launched (eager path; graph replays counted via launches_per_step)
        self.launches_per_step = 0

    # ------------------------------------------------------------------------------------------------------------
    # weight packing
    # ------------------------------------------------------------------------------------------------------------
    @property
    def model(self):
        return self._model()

    def _check_packed(self):
        m = self.model.model
        if self.packed is not None and self.packed["sig"] == m.layers[0].self_attn.q_proj.weight.data_ptr():
            return
        self.pack()

    @torch.no_grad()
    def pack(self):
        """Fuse q|k|v and gate|up weights into single GEMM operands and re-point the nn.Parameters at views of the
        fused storage (state_dict()/load_state_dict() keep working on the reference's key names, nothing is duplicated)."""
        m = self.model.model
        c = self.cfg
        H = c.hidden_size
        I = c.intermediate_size
        layers = []
        for lyr in m.layers:
            a, f = lyr.self_attn, lyr.mlp
            wqkv = torch.cat([a.q_proj.weight.data, a.k_proj.weight.data, a.v_proj.weight.data], 0).contiguous()
            a.q_proj.weight.data, a.k_proj.weight.data, a.v_proj.weight.data = wqkv[:H], wqkv[H:2 * H], wqkv[2 * H:]
            wgu = torch.cat([f.gate_proj.weight.data, f.up_proj.weight.data], 0).contiguous()
            f.gate_proj.weight.data, f.up_proj.weight.data = wgu[:I], wgu[I:]
            layers.append(dict(wqkv=wqkv, wo=a.o_proj.weight.data, wgu=wgu, wd=f.down_proj.weight.data,
                               ln1=lyr.input_layernorm.weight.data, ln2=lyr.post_attention_layernorm.weight.data))
        cos, sin = self._rope_tables(c.max_position_embeddings)
        self.packed = dict(sig=m.layers[0].self_attn.q_proj.weight.data_ptr(), layers=layers, cos=cos, sin=sin, clip=None)
        self._pack_clip()

    def _rope_tables(self, n_pos: int):
        """RoPE tables exactly as the reference builds them: fp32 on the host, rounded to bf16 (dynamic_modeling_llama.py:152-184)"""
        c = self.cfg
        hd = c.hidden_size // c.num_attention_heads
        inv_freq = 1.0 / (c.rope_theta ** (torch.arange(0, hd, 2).float() / hd))
        t = torch.arange(n_pos, dtype=torch.float32)
        emb = torch.cat((torch.outer(t, inv_freq),) * 2, dim=-1)
        return emb.cos().to(self.dtype).to(self.device).contiguous(), emb.sin().to(self.dtype).to(self.device).contiguous()

    def ensure_rope(self, n_pos: int) -> bool:
        """Every kernel indexes cos/sin by the ORIGINAL position (up to prompt + generated tokens).  The reference's
        LlamaRotaryEmbedding regrows its cache when seq_len > max_seq_len_cached (dynamic_modeling_llama.py:176-179); so do we — same
        fp32 formula, so rows below the old length keep their values.  Returns True when the tables moved (captured graphs hold the old
        addresses: the caller drops them)."""
        self._check_packed()
        have = self.packed["cos"].shape[0]
        if n_pos <= have:
            return False
        n = have
        while n < n_pos:
            n *= 2
        self.packed["cos"], self.packed["sin"] = self._rope_tables(n)
        return True

    @torch.no_grad()
    def _pack_clip(self):
        vt = self.model.get_vision_tower()
        if vt is None or not vt.is_loaded:
            return
        vm = vt.vision_tower.vision_model
        C = vt.hidden_size
        w = vm.embeddings.patch_embedding.weight.data
        K = w[0].numel()
        Kpad = (K + 63) // 64 * 64
        wp = torch.zeros(C, Kpad, device=w.device, dtype=w.dtype)
        wp[:, :K] = w.reshape(C, K)
        layers = []
        for lyr in vm.encoder.layers:
            a = lyr.self_attn
            wqkv = torch.cat([a.q_proj.weight.data, a.k_proj.weight.data, a.v_proj.weight.data], 0).contiguous()
            a.q_proj.weight.data, a.k_proj.weight.data, a.v_proj.weight.data = wqkv[:C], wqkv[C:2 * C], wqkv[2 * C:]
            bqkv = torch.cat([a.q_proj.bias.data, a.k_proj.bias.data, a.v_proj.bias.data], 0).contiguous()
            a.q_proj.bias.data, a.k_proj.bias.data, a.v_proj.bias.data = bqkv[:C], bqkv[C:2 * C], bqkv[2 * C:]
            layers.append(dict(wqkv=wqkv, bqkv=bqkv))
        self.packed["clip"] = dict(patch_w=wp, Kpad=Kpad, layers=layers, sig=vm.embeddings.class_embedding.data_ptr())

    def max_partial_elems(self, B: int) -> int:
        c = self.cfg
        H, I, V = c.hidden_size, c.intermediate_size, c.vocab_size
        D = c.sparse_config["d_model"]
        best = 0
        for N, K in ((3 * H, H), (H, H), (2 * I, H), (H, I), (V, H), (D, H)):
            best = max(best, self.L.skinny_gemm_splits(N, K, self.skinny_target_ctas) * B * N)
        return best

    # ------------------------------------------------------------------------------------------------------------
    # small op wrappers
    # ------------------------------------------------------------------------------------------------------------
    def rmsnorm(self, x2d, w, eps, out=None):
        out = torch.empty_like(x2d) if out is None else out
        self.L.rmsnorm(_p(x2d), _p(w), _p(out), x2d.shape[0], x2d.shape[1], eps, _stream())
        self.launches += 1
        return out

    def layernorm(self, x2d, norm, out=None):
        out = torch.empty_like(x2d) if out is None else out
        self.L.layernorm(_p(x2d), _p(norm.weight), _p(norm.bias), _p(out), x2d.shape[0], x2d.shape[1], norm.eps, _stream())
        self.launches += 1
        return out

    def act_(self, x2d, act):
        self.L.bias_act(_p(x2d), None, x2d.shape[0], x2d.shape[1], act, _stream())
        self.launches += 1
        return x2d

    def add_(self, x, y):
        self.L.add(_p(x), _p(y), x.numel(), _stream())
        self.launches += 1
        return x

    def linear(self, x2d, w, bias=None, act=0, residual=None, out=None):
        """nn.Linear at q_len > 1 on the 2-CTA tcgen05 GEMM (csrc/gemm2.cu): out = act(x2d @ w.T + bias) (act 1: GELU-erf, 2: quick_gelu),
        or with `residual` (may alias `out`): out = residual + (x2d @ w.T + bias) with the reference's two bf16 roundings."""
        M, K = x2d.shape
        N = w.shape[0]
        assert x2d.is_contiguous() and w.is_contiguous() and w.shape[1] == K
        if N < 8 or N % 8 or K % 64:
            assert residual is None and act == 0 and N <= 8, (M, N, K)
            out = torch.empty(M, N, device=x2d.device, dtype=x2d.dtype) if out is None else out
            self.L.linear_tiny(_p(x2d), _p(w), _p(bias), _p(out), M, N, K, _stream())
            self.launches += 1
            return out
        out = torch.empty(M, N, device=x2d.device, dtype=x2d.dtype) if out is None else out
        self.L.linear(_p(x2d), _p(w), _p(bias), _p(residual), _p(out), M, N, K, act, _p(self.gemm_ws), _stream())
        self.launches += 1
        return out

    def linear_swiglu(self, x2d, w_gate_up):
        """LlamaMLP gate|up projection with act_fn(gate) * up in the GEMM epilogue (dynamic_modeling_llama.py:328)."""
        M, K = x2d.shape
        I = w_gate_up.shape[0] // 2
        out = torch.empty(M, I, device=x2d.device, dtype=x2d.dtype)
        self.L.linear_swiglu(_p(x2d), _p(w_gate_up), _p(out), M, I, K, _p(self.gemm_ws), _stream())
        self.launches += 1
        return out

    def attn_packed(self, qkv, B, N, heads, hd, causal=False):
        """non-causal attention over a fused [B*N, 3*heads*hd] projection (CLIP, predictor blocks)."""
        C = heads * hd
        out = torch.empty(B * N, C, device=qkv.device, dtype=qkv.dtype)
        cu = self._cu_cache.get((B, N))  # one CSR header per shape, not one arange launch per attention call (57 per prefill)
        if cu is None:
            cu = self._cu_cache[(B, N)] = torch.arange(0, (B + 1) * N, N, device=qkv.device, dtype=torch.int32)
        if hd != 64 or causal:
            raise NotImplementedError("packed attention runs on the tcgen05 kernel for head_dim 64, non-causal (CLIP ViT-L/14, the predictor "
                                      "blocks of the released checkpoints: d_model 512 / nhead 8)")
        self.L.attn_packed_tc4(_p(qkv), _p(out), _p(cu), B * N, B, heads, hd, N, 1.0 / math.sqrt(hd), self.attn_flags, _stream())
        self.launches += 1
        return out

    # ------------------------------------------------------------------------------------------------------------
    # CLIP tower + projector  (clip_encoder.py:53-71, HF CLIPVisionModel, multimodal_projector/builder.py:172-179)
    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def clip_forward(self, tower, images):
        self._check_packed()
        if self.packed["clip"] is None or self.packed["clip"]["sig"] != tower.vision_tower.vision_model.embeddings.class_embedding.data_ptr():
            self._pack_clip()
        pk = self.packed["clip"]
        vc = tower._vc
        vm = tower.vision_tower.vision_model
        in_dtype = images.dtype
        x = images.to(device=self.device, dtype=self.dtype).contiguous()
        B, Cin, S, _ = x.shape
        ps, C, nh = vc["patch_size"], vc["hidden_size"], vc["num_attention_heads"]
        G = S // ps
        NP = G * G
        cols = torch.empty(B * NP, pk["Kpad"], device=self.device, dtype=self.dtype)
        self.L.im2col_patches(_p(x), _p(cols), B, Cin, S, ps, pk["Kpad"], _stream())
        pe = torch.empty(B * NP, C, device=self.device, dtype=self.dtype)
        self.linear(cols, pk["patch_w"], out=pe)  # Conv2d(k = s = 14, bias=False) as an im2col GEMM
        tok = torch.empty(B * (NP + 1), C, device=self.device, dtype=self.dtype)
        self.L.clip_assemble(_p(pe), _p(vm.embeddings.class_embedding), _p(vm.embeddings.position_embedding.weight), _p(tok), B, NP, C, _stream())
        self.launches += 2
        h = self.layernorm(tok, vm.pre_layrnorm)
        n_run = vc["num_hidden_layers"] + 1 + tower.select_layer if tower.select_layer < 0 else tower.select_layer
        for i in range(n_run):  # hidden_states[select_layer]: the layers after it are never needed
            lyr = vm.encoder.layers[i]
            y = self.layernorm(h, lyr.layer_norm1)
            qkv = self.linear(y, pk["layers"][i]["wqkv"], pk["layers"][i]["bqkv"])
            a = self.attn_packed(qkv, B, NP + 1, nh, C // nh)
            self.linear(a, lyr.self_attn.out_proj.weight, lyr.self_attn.out_proj.bias, residual=h, out=h)  # residual in the epilogue
            y = self.layernorm(h, lyr.layer_norm2)
            f = self.linear(y, lyr.mlp.fc1.weight, lyr.mlp.fc1.bias, act=2)  # quick_gelu in the epilogue
            self.linear(f, lyr.mlp.fc2.weight, lyr.mlp.fc2.bias, residual=h, out=h)
        if tower.select_feature == "patch":
            out = torch.empty(B, NP, C, device=self.device, dtype=self.dtype)
            self.L.drop_cls(_p(h), _p(out), B, NP, C, _stream())
            self.launches += 1
        elif tower.select_feature == "cls_patch":
            out = h.view(B, NP + 1, C)
        else:
            raise ValueError(f"Unexpected select feature: {tower.select_feature}")
        return out.to(in_dtype)

    @torch.no_grad()
    def encode_images(self, images):
        m = self.model.model
        f = self.clip_forward(m.vision_tower, images)
        B, N, C = f.shape
        f = f.reshape(B * N, C).to(self.dtype)
        p = m.mm_projector
        y = self.linear(f, p[0].weight, p[0].bias, act=1)  # Linear + GELU(erf) in one launch
        y = self.linear(y, p[2].weight, p[2].bias)
        return y.view(B, N, -1)

    # ------------------------------------------------------------------------------------------------------------
    # predictors
    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def vision_predictor_forward(self, mod, x, image_policy=None):
        """dynamic_modeling_llama.py:1348-1359; x [B,N,H] may be a strided view into the residual stream (rows
        contiguous), so the reference's per-sample re-stacking (:1832-1862) is not needed."""
        B, N, H = x.shape
        assert x.stride(2) == 1 and x.stride(1) == H, "image rows must be contiguous"
        D, nh = mod.d_model, mod.nhead
        ln = mod.down_mlp[0]
        y = torch.empty(B * N, H, device=x.device, dtype=x.dtype)
        self.L.layernorm_strided(_p(x), _p(ln.weight), _p(ln.bias), _p(y), B, N, x.stride(0), H, ln.eps, _stream())
        self.launches += 1
        y = self.linear(y, mod.down_mlp[1].weight, mod.down_mlp[1].bias, act=1)
        # `* image_policy`: the policy is all ones at inference (:1779-1785) -> exact identity in bf16
        for blk in mod.transformer:
            h = self.layernorm(y, blk.norm1)
            qkv = self.linear(h, blk.attn.qkv.weight)
            a = self.attn_packed(qkv, B, N, nh, D // nh)
            self.linear(a, blk.attn.proj.weight, blk.attn.proj.bias, residual=y, out=y)
            h = self.layernorm(y, blk.norm2)
            f = self.linear(h, blk.mlp.fc1.weight, blk.mlp.fc1.bias, act=1)
            self.linear(f, blk.mlp.fc2.weight, blk.mlp.fc2.bias, residual=y, out=y)
        z = torch.empty_like(y)
        g = torch.empty(B, D // 2, device=x.device, dtype=x.dtype)
        self.L.predictor_mix(_p(y), _p(g), _p(z), B, N, D, _stream())
        self.launches += 2
        o = mod.output_mlp
        z = self.linear(z, o[0].weight, o[0].bias, act=1)
        z = self.linear(z, o[2].weight, o[2].bias, act=1)
        return self.linear(z, o[4].weight, o[4].bias).view(B, N, 2)

    @torch.no_grad()
    def text_predictor_forward(self, mod, x):
        """dynamic_modeling_llama.py:1385-1387 (generic entry; the decode step uses the fused dl_text_decision path)."""
        shp = x.shape
        x2 = x.reshape(-1, shp[-1]).contiguous()
        o = mod.output_mlp
        y = self.layernorm(x2, o[0])
        y = self.linear(y, o[1].weight, o[1].bias, act=1)
        y = self.linear(y, o[3].weight, o[3].bias, act=1)
        y = self.linear(y, o[5].weight, o[5].bias, act=1)
        return self.linear(y, o[7].weight, o[7].bias).view(*shp[:-1], 2)

    # ------------------------------------------------------------------------------------------------------------
    # multimodal splice  (dynamic_llava_arch.py:169-601, labels=None / attention_mask=None / one image per sample)
    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def splice(self, input_ids, image_feats, max_len=None, padding_side="right", ids_host=None):
        """-> (inputs_embeds [B,P,H], indices list[dict] (one per sample that has an image)).
        The per-sample segment dicts are host data (the reference builds them with `.item()` reads, dynamic_llava_arch.py:330-334): when the
        caller's ids came from the host (the usual case: tokenizer output), that copy is used and nothing is read back from the device."""
        if ids_host is None:
            ids_host = input_ids if input_ids.device.type == "cpu" else input_ids.to("cpu")
        B, T = ids_host.shape
        n_img = image_feats.shape[1]
        img_pos, seq_len, indices = [], [], []
        for b in range(B):
            row = ids_host[b]
            w = (row == IMAGE_TOKEN_INDEX).nonzero().flatten()
            if len(w) > 1:
                raise NotImplementedError("one <image> placeholder per sample is supported")
            seq_len.append(T)
            if len(w) == 0:
                img_pos.append(-1)
                continue
            s = int(w[0])
            img_pos.append(s)
            ins = row[s + 1:].tolist()
            last = 0
            for j in range(len(ins) - len(USER_IDS) + 1):
                if ins[j: j + len(USER_IDS)] == USER_IDS:
                    last = j
            i0, i1, i2 = s, s + n_img, T - 1 + n_img
            indices.append({"system": [0, i0], "image": [i0, i1], "instruct": [i1, i2], "answer": [i2, i2],
                            "last_instruct": [i1 + last, i2]})
        P = max((T - 1 + n_img) if p >= 0 else T for p in img_pos)
        if padding_side != "right" and any(p < 0 for p in img_pos) and any(p >= 0 for p in img_pos):
            # rows are rectangular ids with one placeholder each, so padding only ever appears next to a row WITHOUT an image; with
            # `tokenizer_padding_side="left"` the reference itself fails there (IndexError on its per-row dict list,
            # dynamic_llava_arch.py:556-566; tests/golden/prepare_cases_bf16.pt "text_only_row_left").  Every other left-padded flow
            # pads nothing and is served as is.
            raise NotImplementedError("left padding next to a row without an image (the reference raises IndexError here)")
        H = self.cfg.hidden_size
        out = torch.empty(B, P, H, device=self.device, dtype=self.dtype)
        ids_dev = input_ids.to(self.device, torch.int64).contiguous()
        sl = torch.tensor(seq_len, device=self.device, dtype=torch.int32)
        ip = torch.tensor(img_pos, device=self.device, dtype=torch.int32)
        self.L.embed_splice(_p(ids_dev), _p(sl), _p(ip), _p(self.model.model.embed_tokens.weight), _p(image_feats.contiguous()),
                            _p(out), B, T, P, n_img, H, self.cfg.vocab_size, _stream())
        self.launches += 1
        if max_len is not None and P > max_len:  # :493-506
            out = out[:, :max_len].contiguous()
            for d in indices:
                for v in d.values():
                    v[0], v[1] = min(v[0], max_len), min(v[1], max_len)
        return out, indices

    # ------------------------------------------------------------------------------------------------------------
    # prefill
    # ------------------------------------------------------------------------------------------------------------
    def _prefill_layer(self, l, x, pos, cu, sess, B, Lq):
        c = self.cfg
        W = self.packed["layers"][l]
        H, nh = c.hidden_size, c.num_attention_heads
        hd = H // nh
        N = x.shape[0]
        K, V, cap = sess.arena.layer(l)
        kvl = sess.arena.kv_len[l]
        xn = self.rmsnorm(x, W["ln1"], c.rms_norm_eps)
        qkv = self.linear(xn, W["wqkv"])
        self.L.rope_append(_p(qkv), _p(pos), _p(cu), _p(kvl), None, _p(self.packed["cos"]), _p(self.packed["sin"]), _p(K), _p(V),
                           None, None, N, B, nh, hd, cap, _stream())
        attn = torch.empty(N, H, device=x.device, dtype=x.dtype)
        scale = 1.0 / math.sqrt(hd)
        self.L.attn_prefill_tc4(_p(qkv), 3 * H, _p(K), _p(V), _p(attn), H, _p(cu), _p(kvl), N, B, nh, cap, Lq, scale, self.attn_flags, _stream())
        self.launches += 2
        self.linear(attn, W["wo"], residual=x, out=x)  # x + o_proj(attn): bf16 projection, bf16 add, in the GEMM epilogue (:1289)
        xn = self.rmsnorm(x, W["ln2"], c.rms_norm_eps, out=xn)
        act = self.linear_swiglu(xn, W["wgu"])  # act_fn(gate) * up in the epilogue of the fused gate|up GEMM (:328)
        self.linear(act, W["wd"], residual=x, out=x)  # (:1295)
        return x

    @torch.no_grad()
    def prefill(self, sess: Session, inputs_embeds, indices, forced_keep=None, all_logits=False, record=None, use_cache=True):
        """DynamicLlamaModel.forward, prefill branch (dynamic_modeling_llama.py:1656-2594) + lm_head (:2709).
        inputs_embeds [B,P,H]; returns bf16 logits [B,V] (last position) or [B,P',V] if all_logits.
        `use_cache=False` (:2393-2504): the sequence is prompt + tokens generated so far, re-fed whole; the generated
        tokens between the sticky `model.model.answer_indice` and the last token are thinned at the sparse layer.  `sess` is scratch."""
        self._check_packed()
        c, sc = self.cfg, self.cfg.sparse_config
        m = self.model.model
        B, P, H = inputs_embeds.shape
        SL, Lyr = sc["sparse_layer"], c.num_hidden_layers
        x = inputs_embeds.reshape(B * P, H).contiguous()
        if x.data_ptr() == inputs_embeds.data_ptr():
            x = x.clone()  # layers update the residual stream in place; never clobber the caller's embeds
        Lq = P
        pos = torch.arange(P, device=self.device, dtype=torch.int32).repeat(B)
        cu = torch.arange(0, (B + 1) * P, P, device=self.device, dtype=torch.int32)
        have_idx = indices is not None and len(indices) == B
        sparsified = False
        for l in range(Lyr):
            if sc["use_vision_predictor"] and l == SL and have_idx:
                s0, e0 = indices[0]["image"]
                if any(d["image"] != indices[0]["image"] for d in indices[1:]):
                    # the reference scores each row's own span (:1833-1842) but gathers and rebuilds positions with SAMPLE 0's bounds
                    # (:1919-1932, :1962-1983): rows with another <image> position are silently mis-sliced there; refuse by name
                    raise NotImplementedError("all samples of a batch must share the <image> position (the reference slices with sample 0's bounds)")
                n_img = e0 - s0
                k = int(n_img * sc["vision_keep_rate"])  # :1899-1901
                x3 = x.view(B, Lq, H)
                logits = m.image_score_predictor(x3[:, s0:e0, :], None)  # via __call__: forward hooks fire
                Pp = Lq - n_img + k
                x_new = torch.empty(B * Pp, H, device=self.device, dtype=self.dtype)
                keep = torch.empty(B, k, device=self.device, dtype=torch.int64)
                pos = torch.empty(B * Pp, device=self.device, dtype=torch.int32)
                cu = torch.empty(B + 1, device=self.device, dtype=torch.int32)
                fk = None if forced_keep is None else forced_keep.to(self.device, torch.int64).contiguous()
                self.L.select_compact(_p(logits.contiguous()), _p(fk), _p(x), _p(x_new), _p(keep), _p(pos), _p(cu), B, Lq, H, s0,
                                      n_img, k, 0, _stream())
                self.launches += 1
                x, Lq, sparsified = x_new, Pp, True
                drop = n_img - k  # :1985-1994 (the reference mutates the caller's dicts in place; so do we)
                for d in indices:
                    d["image"][1] -= drop
                    for key in ("instruct", "last_instruct", "answer"):
                        d[key][0] -= drop
                        d[key][1] -= drop
                if record is not None:
                    record.update(vision_logits=logits, keep_index=keep, position_ids=pos.view(B, Pp))
            if sc["use_text_predictor"] and sc["use_instruct_predictor"] and l == SL and have_idx:
                # prefill-time instruction dropping (:2261-2375): tokens of the last instruct span (its final token excluded)
                # stay only where the instruct predictor's logit0 > logit1
                assert B == 1, "Using text predictor must keep the batch size to 1"
                a0, b0 = indices[0]["last_instruct"][0], indices[0]["last_instruct"][1] - 1
                if b0 < a0:
                    raise NotImplementedError("empty instruct span (the <image> placeholder is the last prompt token)")
                if b0 > a0:
                    n_span = b0 - a0
                    lg = m.instruct_score_predictor(x.view(B, Lq, H)[:, a0:b0, :]).contiguous()  # via __call__: hooks fire
                    dec = torch.empty(n_span, device=self.device, dtype=torch.uint8)
                    self.L.turn_decision(_p(lg), None, _p(dec), 1, n_span, 0, _stream())
                    keep_i = dec.nonzero().flatten()  # one host sync for the new length, like the reference's torch.where (:2280)
                    ar = torch.arange(Lq, device=self.device, dtype=torch.int64)
                    idx = torch.cat([ar[:a0], keep_i + a0, ar[b0:]]).contiguous()
                    Ln = idx.numel()
                    x_new = torch.empty(Ln, H, device=self.device, dtype=self.dtype)
                    pos_new = torch.empty(Ln, device=self.device, dtype=torch.int32)
                    self.L.gather_rows_pos(_p(x), _p(pos), _p(idx), _p(x_new), _p(pos_new), Ln, H, _stream())
                    self.launches += 2
                    drop = Lq - Ln
                    x, pos, Lq = x_new, pos_new, Ln
                    cu = torch.tensor([0, Ln], device=self.device, dtype=torch.int32)
                    for d in indices:  # :2364-2375
                        d["instruct"][1] -= drop
                        d["last_instruct"][1] -= drop
                        d["answer"][0] -= drop
                        d["answer"][1] -= drop
                    if record is not None:
                        record.update(instruct_logits=lg, instruct_keep_index=keep_i.view(1, -1), position_ids=pos.view(B, Lq))
            elif (sc["use_text_predictor"] and sc["use_output_text_predictor"] and not use_cache and l == SL and have_idx):
                # output stage without KV cache (:2393-2504; the instruct branch above precedes it in the reference's elif chain too)
                if m.answer_indice is None:
                    m.answer_indice = indices[0]["instruct"][1]  # :2402-2403, sticky; post-image-drop coordinates
                A = int(m.answer_indice)
                if A > Lq:
                    raise ValueError(f"answer_indice {A} is beyond the sequence ({Lq}): reset model.model.answer_indice = None for a new prompt")
                n_out = max(Lq - 1 - A, 0)  # the reference's [A:-1] slice; empty on the first two calls
                n_keep, rank, keep_o = 0, None, None
                if n_out > 0:
                    if n_out > 8192:
                        raise NotImplementedError("more than 8192 generated tokens in the no-cache mode (dl_output_rank limit)")
                    lg = m.output_text_score_predictor(x.view(B, Lq, H)[:, A:Lq - 1, :]).contiguous()  # via __call__: hooks fire
                    rank = torch.empty(B, n_out, device=self.device, dtype=torch.int32)
                    cnt = torch.empty(B, device=self.device, dtype=torch.int32)
                    self.L.output_rank(_p(lg), _p(rank), _p(cnt), B, n_out, _stream())
                    n_keep = int(cnt.max())  # one host read for the new length, like the reference's slice by `num_keep_node` (:2417-2423)
                    keep_o = torch.empty(B, n_keep, device=self.device, dtype=torch.int64)
                    self.launches += 1
                    if record is not None:
                        record.update(text_logits=lg, output_keep_index=keep_o)
                Ln = A + n_keep + 1  # left ++ kept ++ last: on the first call (Lq == A) the last token appears twice, as in the reference
                rows = torch.empty(B, Ln, device=self.device, dtype=torch.int64)
                self.L.output_keep_index(_p(rank), B, n_out, n_keep, A, Lq, _p(keep_o) if n_keep else None, _p(rows), _stream())
                x_new = torch.empty(B * Ln, H, device=self.device, dtype=self.dtype)
                pos_new = torch.empty(B * Ln, device=self.device, dtype=torch.int32)
                self.L.gather_rows_pos(_p(x), _p(pos), _p(rows), _p(x_new), _p(pos_new), B * Ln, H, _stream())
                self.launches += 2
                x, pos, Lq = x_new, pos_new, Ln
                cu = torch.arange(0, (B + 1) * Ln, Ln, device=self.device, dtype=torch.int32)
                if record is not None:
                    record.update(position_ids=pos.view(B, Lq))
            if record is not None:
                record.setdefault("layer_inputs", []).append(x.view(B, Lq, H).clone())
            x = self._prefill_layer(l, x, pos, cu, sess, B, Lq)
        a = sess.arena
        self.L.kv_advance(_p(a.kv_len), _p(a.pos), None, None, Lyr, B, SL, P, Lq, _stream())
        self.launches += 1
        a.host_len_lo += P
        a.host_ub_hi += Lq
        a.floor_lo, a.floor_hi = a.host_len_lo, a.host_ub_hi  # every sample holds at least the prefill tokens from here on
        if all_logits:
            xn = self.rmsnorm(x, m.norm.weight, c.rms_norm_eps)
            return self.linear(xn, self.model.lm_head.weight).view(B, Lq, -1)
        last = x.view(B, Lq, H)[:, -1, :].contiguous()
        xn = self.rmsnorm(last, m.norm.weight, c.rms_norm_eps)
        return self.linear(xn, self.model.lm_head.weight)

    @torch.no_grad()
    def continue_prefill(self, sess: Session, input_ids, record=None):
        """q_len > 1 over an existing cache (a new user turn): DynamicLlamaModel.forward with a past
        (dynamic_modeling_llama.py:1714-1722 positions from the layer-0 length, per-layer causal mask offset by that layer's cached
        length :1791-1818).  Per-sample cache lengths may differ (earlier dropped tokens); each sample's new keys land right
        after ITS cached keys.  `use_instruct_predictor=False`: plain append in every layer (:1078-1083).  With the instruct
        predictor (:2506-2521) every new token is attended to (cache_utils.py:256-268) but layers >= sparse_layer persist only
        the tokens whose logit0 > logit1, the last one always (cache_utils.py:165-241): the new rows are compacted in place
        after the layer's attention and the per-sample lengths advance by the kept counts."""
        self._check_packed()
        c, sc = self.cfg, self.cfg.sparse_config
        m = self.model.model
        a = sess.arena
        B, Tq = input_ids.shape
        H, nh = c.hidden_size, c.num_attention_heads
        SL = sc["sparse_layer"]
        ids = input_ids.to(self.device, torch.int64).contiguous()
        x = torch.empty(B * Tq, H, device=self.device, dtype=self.dtype)
        self.L.embed_rows(_p(ids), _p(m.embed_tokens.weight), _p(x), B * Tq, H, _stream())
        self.launches += 1
        pos = (a.pos.view(B, 1) + torch.arange(Tq, device=self.device, dtype=torch.int32).view(1, Tq)).reshape(-1).contiguous()
        cu = torch.arange(0, (B + 1) * Tq, Tq, device=self.device, dtype=torch.int32)
        keep = None
        for l in range(c.num_hidden_layers):
            if l == SL and sc["use_text_predictor"] and sc["use_instruct_predictor"]:
                lg = m.instruct_score_predictor(x.view(B, Tq, H)).contiguous()  # via __call__: forward hooks fire
                keep = torch.empty(B * Tq, device=self.device, dtype=torch.uint8)
                self.L.turn_decision(_p(lg), None, _p(keep), B, Tq, 1, _stream())
                self.launches += 1
                if record is not None:
                    record.update(instruct_logits=lg, text_decision=keep.view(B, Tq).bool())
            x = self._prefill_layer(l, x, pos, cu, sess, B, Tq)
            if keep is not None:
                K, V, cap = a.layer(l)
                self.L.kv_compact_turn(_p(K), _p(V), _p(a.kv_len[l]), _p(keep), Tq, B, nh, H // nh, cap, _stream())
                self.launches += 1
        self.L.kv_advance_turn(_p(a.kv_len), _p(a.pos), _p(keep), c.num_hidden_layers, B, SL, Tq, _stream())
        self.launches += 1
        a.host_len_lo += Tq
        a.host_ub_hi += Tq
        if keep is not None:
            a.hi_exact = False
        xn = self.rmsnorm(x, m.norm.weight, c.rms_norm_eps)
        return self.linear(xn, self.model.lm_head.weight).view(B, Tq, -1)

    # ------------------------------------------------------------------------------------------------------------
    # decode step (q_len == 1): weight streaming + ragged attention, no host sync, CUDA-graph replayable
    # ------------------------------------------------------------------------------------------------------------
    def _gemm(self, sess, x, w, M, N, K):
        s = self.L.skinny_gemm_splits(N, K, self.skinny_target_ctas)
        self.L.skinny_gemm(_p(x), _p(w), _p(sess.partial), M, N, K, s, self.l2_prefetch_mb, _stream())
        self.launches += 1
        return s

    def _n_split(self, B, heads, ub_len):
        # split only when the (sample, head) CTAs alone cannot fill the SMs AND every split still gets >= 512 keys: at config 2's
        # layers 0-1 (B*heads = 256, 735 keys) one pass is faster than two splits + combine (22.3 vs 26.0 us, r1_microbench_v1.json)
        want = (2 * 148 + B * heads - 1) // (B * heads)
        return max(1, min(want, ub_len // 512, 16))

    def decode_layers(self, sess: Session, use_decisions: bool, quirk: bool):
        """embed(next_ids) -> L decoder layers -> final norm -> lm_head: writes sess.logits (bf16 [B,V])."""
        c, sc = self.cfg, self.cfg.sparse_config
        m = self.model.model
        L_ = self.L
        B, H, I, V = sess.B, c.hidden_size, c.intermediate_size, c.vocab_size
        nh = c.num_attention_heads
        hd = H // nh
        SL, Lyr, eps = sc["sparse_layer"], c.num_hidden_layers, c.rms_norm_eps
        a = sess.arena
        st = _stream()
        pk = self.packed
        scale = 1.0 / math.sqrt(hd)
        L_.embed_rows(_p(sess.next_ids), _p(m.embed_tokens.weight), _p(sess.xres), B, H, st)
        L_.rmsnorm(_p(sess.xres), _p(pk["layers"][0]["ln1"]), _p(sess.xn), B, H, eps, st)
        self.launches += 2
        keep = None
        join_side = False
        for l in range(Lyr):
            W = pk["layers"][l]
            if l == SL and use_decisions:
                tp = m.output_text_score_predictor.output_mlp
                D = tp[1].weight.shape[0]
                # The predictor (LayerNorm -> Linear(H->D) -> MLP tail -> decision) only needs this layer's INPUT and its result is
                # first consumed by this layer's attention kernel: it runs on a side stream next to the layer's QKV GEMM (fork /
                # join are graph-capturable) instead of ~30 us in series on every decode step.
                cur = torch.cuda.current_stream()
                side = self.side_stream
                side.wait_stream(cur)
                sst = ctypes.c_void_p(side.cuda_stream)
                s_tp = L_.skinny_gemm_splits(D, H, self.skinny_target_ctas)
                L_.layernorm(_p(sess.xres), _p(tp[0].weight), _p(tp[0].bias), _p(sess.tp_ln), B, H, tp[0].eps, sst)
                L_.skinny_gemm(_p(sess.tp_ln), _p(tp[1].weight), _p(sess.tp_partial), B, D, H, s_tp, self.l2_prefetch_mb, sst)
                L_.text_decision(_p(sess.tp_partial), s_tp, B, D, _p(tp[1].bias), _p(tp[3].weight), _p(tp[3].bias), _p(tp[5].weight),
                                 _p(tp[5].bias), _p(tp[7].weight), _p(tp[7].bias), _p(sess.forced), _p(a.step), -1, _p(a.keep),
                                 _p(sess.tp_logits), sst)
                self.launches += 3
                keep = a.keep
                join_side = True
            K_, V_, cap = a.layer(l)
            kvl = a.kv_len[l]
            s = self._gemm(sess, sess.xn, W["wqkv"], B, 3 * H, H)
            ns = self._n_split(B, nh, cap)
            if ns > 1 and sess.attn_ws is None:
                sess.attn_ws = torch.empty(L_.attn_decode_workspace_bytes(B, nh, 16) // 4, device=self.device, dtype=torch.float32)
            if join_side:  # the keep decision must be there before this layer's attention appends (or not) the new token's K/V
                torch.cuda.current_stream().wait_stream(self.side_stream)
                join_side = False
            L_.attn_decode_fused(_p(sess.partial), s, _p(a.pos), _p(keep if l >= SL else None), _p(pk["cos"]), _p(pk["sin"]), _p(K_), _p(V_),
                                 _p(kvl), _p(sess.attn), _p(sess.attn_ws), B, nh, hd, cap, ns, scale,
                                 1 if (quirk and l >= SL and (keep is not None or not a.hi_exact)) else 0,
                                 _p(W["wo"]) if self.attn_prefetch_wo else None, W["wo"].numel() * 2,
                                 (a.floor_lo if l < SL else a.floor_hi) if self.attn_early_kv else 0, st)
            s = self._gemm(sess, sess.attn, W["wo"], B, H, H)
            L_.dec_reduce_residual_norm(_p(sess.partial), _p(sess.xres), _p(W["ln2"]), _p(sess.xn), B, H, s, eps, st)
            L_.skinny_gemm_swiglu(_p(sess.xn), _p(W["wgu"]), _p(sess.act), B, I, H, st)  # gate|up GEMM with SwiGLU in the epilogue
            s = self._gemm(sess, sess.act, W["wd"], B, H, I)
            nxt = pk["layers"][l + 1]["ln1"] if l + 1 < Lyr else m.norm.weight
            L_.dec_reduce_residual_norm(_p(sess.partial), _p(sess.xres), _p(nxt), _p(sess.xn), B, H, s, eps, st)
            self.launches += 4 + (2 if ns > 1 else 1)  # + 3 split-K GEMMs counted in _gemm
        s = self._gemm(sess, sess.xn, self.model.lm_head.weight, B, V, H)
        L_.dec_reduce_bf16(_p(sess.partial), _p(sess.logits), B, V, s, st)
        self.launches += 1
        return keep

    def decode_advance(self, sess: Session, keep):
        a = sess.arena
        self.L.kv_advance(_p(a.kv_len), _p(a.pos), _p(a.step), _p(keep), self.cfg.num_hidden_layers, sess.B,
                          self.cfg.sparse_config["sparse_layer"], 1, 1, _stream())
        self.launches += 1
        a.host_len_lo += 1
        a.host_ub_hi += 1
        if keep is not None:
            a.hi_exact = False

    def sample(self, sess: Session, logits, eos, pad, temperature, top_k, top_p):
        """HF `sample` step (temperature -> top-k -> top-p -> multinomial) on the device; same bookkeeping as greedy()."""
        self.L.sample_next(_p(logits), sess.B, logits.shape[-1], float(temperature), int(top_k), float(top_p), _p(sess.rng_seed),
                           _p(sess.next_ids), _p(sess.tokens), sess.tokens.shape[1], _p(sess.arena.step), _p(sess.finished),
                           -1 if eos is None else int(eos), int(pad), None, _stream())
        self.launches += 1

    def greedy(self, sess: Session, logits, eos, pad):
        self.L.greedy_next(_p(logits), sess.B, logits.shape[-1], _p(sess.next_ids), _p(sess.tokens), sess.tokens.shape[1],
                           _p(sess.arena.step), _p(sess.finished), -1 if eos is None else int(eos), int(pad), _stream())
        self.launches += 1
