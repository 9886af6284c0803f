"""Parameter tree of Dynamic-LLaVA under the reference's state-dict key names (SURVEY.md §8b "State-dict keys").

These nn.Modules only HOLD weights (created uninitialised, on the target device, in the target dtype) and expose the
attributes the reference's callers touch.  Every computation goes through `dynamic_llava_b200.engine` (hand-written
sm_100a kernels + cuBLASLt GEMMs); the two predictor modules are callable so that forward hooks registered on
`model.model.image_score_predictor` (llava/dynamic_eval/visualize.py:72-74) keep firing.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class Linear(nn.Module):
    def __init__(self, in_f, out_f, bias=True, device=None, dtype=None):
        super().__init__()
        self.in_features, self.out_features = in_f, out_f
        self.weight = nn.Parameter(torch.empty(out_f, in_f, device=device, dtype=dtype), requires_grad=False)
        if bias:
            self.bias = nn.Parameter(torch.empty(out_f, device=device, dtype=dtype), requires_grad=False)
        else:
            self.register_parameter("bias", None)


class Norm(nn.Module):
    """weight (+bias) holder for RMSNorm / LayerNorm."""

    def __init__(self, dim, bias=False, eps=1e-5, device=None, dtype=None):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.empty(dim, device=device, dtype=dtype), requires_grad=False)
        if bias:
            self.bias = nn.Parameter(torch.empty(dim, device=device, dtype=dtype), requires_grad=False)


class Act(nn.Module):
    """index placeholder so nn.Sequential numbering matches the reference (nn.GELU slots)."""


class Embedding(nn.Module):
    def __init__(self, n, dim, device=None, dtype=None):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(n, dim, device=device, dtype=dtype), requires_grad=False)


class SelfAttn(nn.Module):
    def __init__(self, H, kw):
        super().__init__()
        self.q_proj = Linear(H, H, False, **kw)
        self.k_proj = Linear(H, H, False, **kw)
        self.v_proj = Linear(H, H, False, **kw)
        self.o_proj = Linear(H, H, False, **kw)


class MLP(nn.Module):
    def __init__(self, H, I, kw):
        super().__init__()
        self.gate_proj = Linear(H, I, False, **kw)
        self.up_proj = Linear(H, I, False, **kw)
        self.down_proj = Linear(I, H, False, **kw)


class DecoderLayer(nn.Module):
    """DynamicLlamaDecoderLayer (dynamic_modeling_llama.py:1221-1234)."""

    def __init__(self, cfg, kw):
        super().__init__()
        H, I, eps = cfg.hidden_size, cfg.intermediate_size, cfg.rms_norm_eps
        self.self_attn = SelfAttn(H, kw)
        self.mlp = MLP(H, I, kw)
        self.input_layernorm = Norm(H, eps=eps, **kw)
        self.post_attention_layernorm = Norm(H, eps=eps, **kw)


class _PredAttn(nn.Module):
    def __init__(self, D, kw):
        super().__init__()
        self.qkv = Linear(D, 3 * D, False, **kw)  # qkv_bias=False (custom_transformer_layer.py:146)
        self.proj = Linear(D, D, True, **kw)


class _PredMlp(nn.Module):
    def __init__(self, D, FF, kw):
        super().__init__()
        self.fc1 = Linear(D, FF, True, **kw)
        self.fc2 = Linear(FF, D, True, **kw)


class PredictorBlock(nn.Module):
    """SelfTransformerEncoderBlock (custom_transformer_layer.py:276-323)."""

    def __init__(self, D, FF, kw):
        super().__init__()
        self.norm1 = Norm(D, bias=True, **kw)
        self.attn = _PredAttn(D, kw)
        self.norm2 = Norm(D, bias=True, **kw)
        self.mlp = _PredMlp(D, FF, kw)


class VisionPredictor(nn.Module):
    """dynamic_modeling_llama.py:1308-1359.  forward(x [B,N,H], image_policy [B,N,1]) -> logits [B,N,2]."""

    def __init__(self, H, sc, kw):
        super().__init__()
        D, FF = sc["d_model"], sc["dim_feedforward"]
        self.input_dim, self.d_model, self.nhead = H, D, sc["nhead"]
        self.dim_feedforward, self.num_layers = FF, sc["num_layers"]
        self.down_mlp = nn.Sequential(Norm(H, bias=True, **kw), Linear(H, D, True, **kw), Act())
        self.transformer = nn.Sequential(*[PredictorBlock(D, FF, kw) for _ in range(self.num_layers)])
        self.output_mlp = nn.Sequential(Linear(D, D // 2, True, **kw), Act(), Linear(D // 2, D // 4, True, **kw), Act(),
                                        Linear(D // 4, 2, True, **kw))
        self._engine = None

    def forward(self, x, image_policy=None, **kw):
        if self._engine is None:
            raise RuntimeError("VisionPredictor is not attached to an engine")
        return self._engine().vision_predictor_forward(self, x, image_policy, **kw)


class TextPredictor(nn.Module):
    """dynamic_modeling_llama.py:1362-1387.  forward(x [..., H]) -> logits [..., 2]."""

    def __init__(self, H, sc, kw):
        super().__init__()
        D = sc["d_model"]
        self.input_dim, self.d_model = H, D
        self.output_mlp = nn.Sequential(Norm(H, bias=True, **kw), Linear(H, D, True, **kw), Act(), Linear(D, D // 2, True, **kw),
                                        Act(), Linear(D // 2, D // 4, True, **kw), Act(), Linear(D // 4, 2, True, **kw))
        self._engine = None

    def forward(self, x):
        if self._engine is None:
            raise RuntimeError("TextPredictor is not attached to an engine")
        return self._engine().text_predictor_forward(self, x)


# ------------------------------------------------------------------------------------------------------------------
# CLIP vision tower (HF CLIPVisionModel key names; multimodal_encoder/clip_encoder.py wraps it)
# ------------------------------------------------------------------------------------------------------------------
class _ClipAttn(nn.Module):
    def __init__(self, C, kw):
        super().__init__()
        self.k_proj = Linear(C, C, True, **kw)
        self.v_proj = Linear(C, C, True, **kw)
        self.q_proj = Linear(C, C, True, **kw)
        self.out_proj = Linear(C, C, True, **kw)


class _ClipMlp(nn.Module):
    def __init__(self, C, CI, kw):
        super().__init__()
        self.fc1 = Linear(C, CI, True, **kw)
        self.fc2 = Linear(CI, C, True, **kw)


class _ClipLayer(nn.Module):
    def __init__(self, C, CI, eps, kw):
        super().__init__()
        self.self_attn = _ClipAttn(C, kw)
        self.layer_norm1 = Norm(C, bias=True, eps=eps, **kw)
        self.mlp = _ClipMlp(C, CI, kw)
        self.layer_norm2 = Norm(C, bias=True, eps=eps, **kw)


class _PatchEmbedding(nn.Module):
    def __init__(self, C, nch, ps, kw):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(C, nch, ps, ps, **kw), requires_grad=False)


class _ClipEmbeddings(nn.Module):
    def __init__(self, vc, kw):
        super().__init__()
        C, ps = vc["hidden_size"], vc["patch_size"]
        self.class_embedding = nn.Parameter(torch.empty(C, **kw), requires_grad=False)
        self.patch_embedding = _PatchEmbedding(C, vc["num_channels"], ps, kw)
        self.position_embedding = Embedding((vc["image_size"] // ps) ** 2 + 1, C, **kw)


class _ClipEncoder(nn.Module):
    def __init__(self, vc, kw):
        super().__init__()
        self.layers = nn.ModuleList([_ClipLayer(vc["hidden_size"], vc["intermediate_size"], vc["layer_norm_eps"], kw)
                                     for _ in range(vc["num_hidden_layers"])])


class _ClipVisionTransformer(nn.Module):
    def __init__(self, vc, kw):
        super().__init__()
        eps = vc["layer_norm_eps"]
        self.embeddings = _ClipEmbeddings(vc, kw)
        self.pre_layrnorm = Norm(vc["hidden_size"], bias=True, eps=eps, **kw)
        self.encoder = _ClipEncoder(vc, kw)
        self.post_layernorm = Norm(vc["hidden_size"], bias=True, eps=eps, **kw)


class ClipVisionModel(nn.Module):
    def __init__(self, vc, kw):
        super().__init__()
        self.vision_model = _ClipVisionTransformer(vc, kw)


class _VisionCfg:
    def __init__(self, vc):
        self.__dict__.update(vc)


class CLIPVisionTower(nn.Module):
    """multimodal_encoder/clip_encoder.py:7-102: same attributes (`is_loaded`, `load_model`, `config`, `hidden_size`,
    `num_patches_per_side`, `num_patches`, `dtype`, `device`), forward(images) -> [B, n_patches, C] with CLS dropped."""

    def __init__(self, vision_cfg: dict, select_layer: int = -2, select_feature: str = "patch", kw=None, name="synthetic-clip"):
        super().__init__()
        self.is_loaded = False
        self.vision_tower_name = name
        self.select_layer = select_layer
        self.select_feature = select_feature
        self._vc = dict(vision_cfg)
        self._kw = kw or {}
        self.image_processor = None
        self._engine = None

    def load_model(self, device_map=None):
        if self.is_loaded:
            print("{} is already loaded, `load_model` called again, skipping.".format(self.vision_tower_name))
            return
        self.vision_tower = ClipVisionModel(self._vc, self._kw)
        self.is_loaded = True

    def forward(self, images):
        if self._engine is None:
            raise RuntimeError("CLIPVisionTower is not attached to an engine")
        return self._engine().clip_forward(self, images)

    @property
    def config(self):
        return _VisionCfg(self._vc)

    @property
    def dtype(self):
        return self.vision_tower.vision_model.embeddings.class_embedding.dtype

    @property
    def device(self):
        return self.vision_tower.vision_model.embeddings.class_embedding.device

    @property
    def hidden_size(self):
        return self._vc["hidden_size"]

    @property
    def num_patches_per_side(self):
        return self._vc["image_size"] // self._vc["patch_size"]

    @property
    def num_patches(self):
        return self.num_patches_per_side ** 2
