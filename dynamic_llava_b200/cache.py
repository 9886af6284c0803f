"""Device-resident ragged KV arena + the legacy-tuple view the reference's callers index.

Replaces `DynamicCachePlus` (llava/model/language_model/cache_utils.py:63-318).  The reference keeps a Python list of
per-layer tensors grown by `torch.cat` (an O(len) copy of K and V per layer per step) and a CPU list of
`true_cache_length`; for B > 1 it rebuilds and zero-pads every layer each step (:165-241).  Here:

* K/V live in two pre-sized slabs `[layers, B, heads, cap, head_dim]` — one for layers below `sparse_layer` (never
  drop: SURVEY.md fact 4) and one for layers at/above it — so an append is a 2*H*2-byte write at slot `kv_len[l][b]`;
* `kv_len int32 [L, B]`, `pos int32 [B]`, the step counter and the keep decisions stay on the device, so a decode step
  has no host sync and replays as a CUDA graph;
* slots at or beyond `kv_len` are never written and the slabs are zero-initialised, hence the strided view
  `slab[l][:, :, :max_b(len), :]` IS the reference's zero-padded `[B, heads, len, head_dim]` tensor — the legacy tuple
  `(((K, V),) * L, [true_len int64[B]] * L)` (cache_utils.py:295-302) is served lazily from views with no copy.
"""
from __future__ import annotations

from typing import List, Optional

import torch


class RaggedKVArena:
    def __init__(self, n_layers: int, sparse_layer: int, batch: int, heads: int, head_dim: int, cap_lo: int, cap_hi: int,
                 device, dtype=torch.bfloat16):
        self.L, self.SL, self.B, self.heads, self.hd = n_layers, min(sparse_layer, n_layers), batch, heads, head_dim
        self.device, self.dtype = device, dtype
        self.cap_lo, self.cap_hi = cap_lo, cap_hi
        self.k_lo = torch.zeros(self.SL, batch, heads, cap_lo, head_dim, device=device, dtype=dtype)
        self.v_lo = torch.zeros_like(self.k_lo)
        self.k_hi = torch.zeros(self.L - self.SL, batch, heads, cap_hi, head_dim, device=device, dtype=dtype)
        self.v_hi = torch.zeros_like(self.k_hi)
        self.kv_len = torch.zeros(self.L, batch, device=device, dtype=torch.int32)
        self.pos = torch.zeros(batch, device=device, dtype=torch.int32)
        self.step = torch.zeros(1, device=device, dtype=torch.int32)
        self.keep = torch.ones(batch, device=device, dtype=torch.uint8)
        # host mirrors of what is host-knowable without a sync
        self.host_len_lo = 0  # layers < SL: every token is appended
        self.host_ub_hi = 0   # layers >= SL: upper bound (exact when nothing has been dropped)
        self.hi_exact = True
        self.floor_lo = self.floor_hi = 0  # host-known lower bounds of kv_len (set at prefill)
        self.generation = 0  # bumped by reset(): a LegacyPastView of an earlier conversation is recognisably stale

    def reset(self, lengths_only: bool = False):
        """lengths_only: a scratch arena (use_cache=False) is never served as a zero-padded legacy tuple, so the slabs may keep stale
        rows beyond the lengths"""
        for t in ((self.kv_len, self.pos, self.step) if lengths_only else
                  (self.k_lo, self.v_lo, self.k_hi, self.v_hi, self.kv_len, self.pos, self.step)):
            t.zero_()
        self.keep.fill_(1)
        self.host_len_lo = self.host_ub_hi = 0
        self.hi_exact = True
        self.floor_lo = self.floor_hi = 0
        self.generation += 1

    def layer(self, l: int):
        """-> (K slab [B, heads, cap, hd], V slab, cap)"""
        if l < self.SL:
            return self.k_lo[l], self.v_lo[l], self.cap_lo
        return self.k_hi[l - self.SL], self.v_hi[l - self.SL], self.cap_hi

    def room(self) -> int:
        return min(self.cap_lo - self.host_len_lo, self.cap_hi - self.host_ub_hi)

    def grow(self, need_lo: int, need_hi: int):
        """Re-slab with larger capacities (only the hand-rolled `forward` loops hit this; `generate` pre-sizes)."""
        def regrow(t, cap):
            n = torch.zeros(*t.shape[:3], cap, t.shape[4], device=t.device, dtype=t.dtype)
            n[:, :, :, : t.shape[3]] = t
            return n
        if need_lo > self.cap_lo:
            self.cap_lo = max(need_lo, 2 * self.cap_lo)
            self.k_lo, self.v_lo = regrow(self.k_lo, self.cap_lo), regrow(self.v_lo, self.cap_lo)
        if need_hi > self.cap_hi:
            self.cap_hi = max(need_hi, 2 * self.cap_hi)
            self.k_hi, self.v_hi = regrow(self.k_hi, self.cap_hi), regrow(self.v_hi, self.cap_hi)

    def memory_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.k_lo, self.v_lo, self.k_hi, self.v_hi))


class LegacyPastView:
    """Quacks like the reference's `(tuple_L((K, V)), list_L(int64[B]))` (cache_utils.py:295-302) without copying.

    `past[0][l][0].shape[-2]` (model_lvis_for_meteor.py:177) and `past[0][0][0].shape[2]`
    (dynamic_modeling_llama.py:2831) work; touching a layer >= sparse_layer or `past[1]` costs one device->host read of
    the lengths.  The object owns the session it came from and is handed back verbatim on the next `forward` call."""

    def __init__(self, arena: RaggedKVArena, session=None):
        self.arena = arena
        self.session = session
        self._lens = None
        self.stamp = (arena.generation, arena.host_len_lo)  # forward() refuses a view the cache has moved past

    def _host_lens(self) -> torch.Tensor:
        if self._lens is None or self._lens_step != self.arena.host_len_lo:
            self._lens = self.arena.kv_len.to("cpu", torch.int64)
            self._lens_step = self.arena.host_len_lo
        return self._lens

    def _layer_view(self, l: int):
        a = self.arena
        k, v, _ = a.layer(l)
        if l < a.SL:
            n = a.host_len_lo
        elif a.hi_exact:
            n = a.host_ub_hi
        else:
            n = int(self._host_lens()[l].max())
        return (k[:, :, :n, :], v[:, :, :n, :])

    def __len__(self):
        return 2

    def __getitem__(self, i):
        if i == 0:
            return _LayerTuple(self)
        if i == 1:
            lens = self._host_lens()
            return [lens[l].clone() for l in range(self.arena.L)]
        raise IndexError(i)

    def __iter__(self):
        yield self[0]
        yield self[1]


class _LayerTuple:
    def __init__(self, view: LegacyPastView):
        self.view = view

    def __len__(self):
        return self.view.arena.L

    def __getitem__(self, l):
        L = self.view.arena.L
        if isinstance(l, slice):
            return tuple(self.view._layer_view(i) for i in range(*l.indices(L)))
        if l < 0:
            l += L
        if not 0 <= l < L:
            raise IndexError(l)
        return self.view._layer_view(l)

    def __iter__(self):
        for l in range(len(self)):
            yield self.view._layer_view(l)
