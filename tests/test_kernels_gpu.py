"""Per-kernel parity on the GPU: every `dl_*` entry point is called THROUGH THE C ABI (ctypes) on seeded inputs and
compared with the CPU oracle (oracle/dynllava_oracle.py, itself pinned bit-exact to the reference).

Bars: integer / index outputs and pure data movement are bit-exact; elementwise bf16 ops are bit-exact except where a
different fp32 reduction order can flip a last-place rounding (<= 1 bf16 ulp on <= 0.5 % of elements, stated per test);
attention / GEMM outputs are compared against an fp32 evaluation of the same bf16 inputs with the tolerance in the test.
"""
import ctypes
import math
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import dynllava_oracle as O

pytestmark = pytest.mark.gpu

DEV = "cuda"
# element type under test: bfloat16 by default; tests/test_fp16_gpu.py re-runs this file with DLB200_TEST_DTYPE=float16 against
# libdynllava_b200_f16.so (the oracle then runs in float16 too, which is what the reference's eval drivers load)
BF = getattr(torch, os.environ.get("DLB200_TEST_DTYPE", "bfloat16"))


@pytest.fixture(scope="module")
def L():
    from dynamic_llava_b200._lib import lib
    lb = lib(BF)
    assert torch.cuda.is_available() and lb.device_is_sm100() == 1
    return lb


_KEEP = []  # device tensors whose raw pointers were handed to the C ABI: keep them alive until the test ends


@pytest.fixture(autouse=True)
def _keepalive():
    yield
    torch.cuda.synchronize()
    _KEEP.clear()


def P(t):
    """raw device pointer for the C ABI; CPU tensors are uploaded first.  The tensor is pinned in _KEEP because a
    temporary such as `P(x.to(DEV))` would otherwise be freed (and its block reused) before the kernel runs."""
    if t is None:
        return None
    if t.device.type == "cpu":
        t = t.to(DEV)
    _KEEP.append(t)
    return ctypes.c_void_p(t.data_ptr())


def ST():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def rnd(*shape, seed=0, scale=1.0, dtype=BF):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype)


def ulp_close(a, b, max_frac=0.005, abs_ok=0.0):
    """bit-exact except <= max_frac of elements off by one bf16 ulp (elements with |a-b| <= abs_ok are exempt: used where
    the reference formula itself cancels catastrophically, e.g. the far negative tail of erf-GELU)"""
    a, b = a.cpu(), b.cpu()
    if torch.equal(a, b):
        return True
    max_ulp = 1
    if BF == torch.float16:
        # a float16 ulp is 8x smaller than a bfloat16 one, so the same fp32 reduction-order noise flips 8x as many last places, and a
        # flipped intermediate that is then scaled into a lower binade (RMSNorm's `weight * x_normed`) shows up as 2 ulp
        max_frac, max_ulp = 8 * max_frac, 2
    ai, bi = a.view(torch.int16).int(), b.view(torch.int16).int()
    d = (ai - bi).abs()
    d[(a.float() - b.float()).abs() <= abs_ok] = 0
    frac = (d > 0).float().mean().item()
    assert d.max().item() <= max_ulp, f"max ulp diff {d.max().item()}"
    assert frac <= max_frac, f"{frac:.4%} elements differ"
    return True


# ---------------------------------------------------------------------------------------------- K4 row kernels
@pytest.mark.parametrize("rows,H", [(7, 256), (33, 4096), (5, 5120), (3, 1024)])
def test_rmsnorm(L, rows, H):
    x, w = rnd(rows, H, seed=1, scale=2.0), 1 + 0.1 * rnd(H, seed=2).float()
    w = w.to(BF)
    ref = O.rms_norm(x, w, 1e-5)
    y = torch.empty(rows, H, device=DEV, dtype=BF)
    L.rmsnorm(P(x.to(DEV)), P(w.to(DEV)), P(y), rows, H, 1e-5, ST())
    torch.cuda.synchronize()
    ulp_close(y, ref)


@pytest.mark.parametrize("rows,H", [(9, 128), (17, 1024), (4, 4096), (6, 512)])
def test_layernorm(L, rows, H):
    x, w, b = rnd(rows, H, seed=3, scale=1.5), (1 + 0.1 * rnd(H, seed=4).float()).to(BF), rnd(H, seed=5, scale=0.1)
    ref = F.layer_norm(x, (H,), w, b, 1e-5)
    y = torch.empty(rows, H, device=DEV, dtype=BF)
    L.layernorm(P(x.to(DEV)), P(w.to(DEV)), P(b.to(DEV)), P(y), rows, H, 1e-5, ST())
    torch.cuda.synchronize()
    ulp_close(y, ref, max_frac=0.01)


def test_layernorm_strided_matches_restack(L):
    B, Pn, H, s0, N = 3, 40, 256, 5, 16
    x = rnd(B, Pn, H, seed=6)
    w, b = (1 + 0.1 * rnd(H, seed=7).float()).to(BF), rnd(H, seed=8, scale=0.1)
    ref = F.layer_norm(x[:, s0:s0 + N], (H,), w, b, 1e-5).reshape(B * N, H)
    xd = x.to(DEV)
    view = xd[:, s0:s0 + N]
    y = torch.empty(B * N, H, device=DEV, dtype=BF)
    L.layernorm_strided(P(view), P(w.to(DEV)), P(b.to(DEV)), P(y), B, N, view.stride(0), H, 1e-5, ST())
    torch.cuda.synchronize()
    ulp_close(y, ref, max_frac=0.01)


def test_swiglu_bitexact(L):
    rows, I = 19, 11008
    gu = rnd(rows, 2 * I, seed=9, scale=2.0)
    ref = F.silu(gu[:, :I]) * gu[:, I:]
    out = torch.empty(rows, I, device=DEV, dtype=BF)
    L.swiglu(P(gu.to(DEV)), P(out), rows, I, ST())
    torch.cuda.synchronize()
    ulp_close(out, ref, max_frac=0.002)  # expf vs the host libm can differ in the last fp32 bit


@pytest.mark.parametrize("act", [1, 2])
def test_activation(L, act):
    x = rnd(13, 2048, seed=10, scale=2.5)
    ref = F.gelu(x) if act == 1 else x * torch.sigmoid(1.702 * x)
    xd = x.to(DEV).clone()
    L.bias_act(P(xd), None, 13, 2048, act, ST())
    torch.cuda.synchronize()
    ulp_close(xd, ref, max_frac=0.002, abs_ok=1e-5)


def test_residual_add_bitexact(L):
    x, y = rnd(11, 4096, seed=11), rnd(11, 4096, seed=12)
    xd = x.to(DEV).clone()
    L.add(P(xd), P(y.to(DEV)), x.numel(), ST())
    torch.cuda.synchronize()
    assert torch.equal(xd.cpu(), x + y)


def test_predictor_mix(L):
    B, N, C = 2, 576, 512
    y = rnd(B, N, C, seed=13)
    pol = torch.ones(B, N, 1, dtype=BF)
    glob = (y[:, :, C // 2:] * pol).sum(dim=1, keepdim=True) / torch.sum(pol, dim=1, keepdim=True)
    ref = torch.cat([y[:, :, : C // 2], glob.expand(B, N, C // 2)], dim=-1)
    z = torch.empty(B, N, C, device=DEV, dtype=BF)
    g = torch.empty(B, C // 2, device=DEV, dtype=BF)
    L.predictor_mix(P(y.to(DEV)), P(g), P(z), B, N, C, ST())
    torch.cuda.synchronize()
    assert torch.equal(z[:, :, : C // 2].cpu(), ref[:, :, : C // 2])
    ulp_close(z[:, :, C // 2:].contiguous(), ref[:, :, C // 2:].contiguous(), max_frac=0.02)


# ---------------------------------------------------------------------------------------------- splice / embed / greedy
def test_embed_splice_matches_oracle(L):
    B, T, H, V, n_img, s = 3, 12, 256, 512, 16, 5
    g = torch.Generator().manual_seed(14)
    ids = torch.randint(3, V, (B, T), generator=g)
    ids[:, s] = -200
    table, feat = rnd(V, H, seed=15), rnd(B, n_img, H, seed=16)
    # oracle splice = cat[embed(sys), image, embed(instruct)]
    ref = torch.stack([torch.cat([table[ids[b, :s]], feat[b], table[ids[b, s + 1:]]]) for b in range(B)])
    Pn = T - 1 + n_img
    out = torch.empty(B, Pn, H, device=DEV, dtype=BF)
    sl = torch.full((B,), T, dtype=torch.int32, device=DEV)
    ip = torch.full((B,), s, dtype=torch.int32, device=DEV)
    L.embed_splice(P(ids.to(DEV)), P(sl), P(ip), P(table.to(DEV)), P(feat.to(DEV)), P(out), B, T, Pn, n_img, H, V, ST())
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), ref)


def test_greedy_next_and_eos(L):
    B, V = 4, 32000
    lg = rnd(B, V, seed=17)
    lg[1, 777] = 50.0
    lg[2, 5] = lg[2, 31000] = 60.0  # tie -> first index
    ref = torch.argmax(lg.float(), dim=-1)
    nxt = torch.zeros(B, dtype=torch.int64, device=DEV)
    toks = torch.full((B, 4), -1, dtype=torch.int64, device=DEV)
    step = torch.tensor([2], dtype=torch.int32, device=DEV)
    fin = torch.tensor([0, 0, 0, 1], dtype=torch.int32, device=DEV)
    L.greedy_next(P(lg.to(DEV)), B, V, P(nxt), P(toks), 4, P(step), P(fin), 777, 0, ST())
    torch.cuda.synchronize()
    assert nxt[:3].tolist() == ref[:3].tolist() and ref[2] == 5
    assert nxt[3].item() == 0  # finished row emits pad
    assert fin.tolist() == [0, 1, 0, 1]
    assert toks[:, 2].tolist() == nxt.tolist() and (toks[:, [0, 1, 3]] == -1).all()


# ---------------------------------------------------------------------------------------------- K1 select + compact
def _tie_aware_same(keep_gpu, logits, k):
    """Two checks of the selection.
    (1) EXACT against the kernel's stated semantics = torch-CUDA's bf16 log_softmax (fp32 accumulate, ONE rounding to bf16), stable
        descending order: the k best (score, lower index first), returned ascending.
    (2) Against the CPU oracle's scores.  torch-CPU's bf16 log_softmax rounds intermediates and lands one bf16 ulp away from the
        singly-rounded value on ~45 % of the entries (checked on these inputs), so that comparison is tie- AND ulp-aware: everything
        clearly above the k-th score must be kept, nothing clearly below it may be."""
    s1 = F.log_softmax(logits.float(), dim=-1)[:, :, 0].to(BF).float()
    s_cpu = F.log_softmax(logits, dim=-1)[:, :, 0].float()
    n = logits.shape[1]
    for b in range(logits.shape[0]):
        got = keep_gpu[b].tolist()
        assert got == sorted(set(got)) and len(got) == k
        order = sorted(range(n), key=lambda i: (-s1[b, i].item(), i))
        want = sorted(order[:k])
        if got != want:  # expf/logf may differ from libm by an ulp exactly at a bf16 rounding boundary: allow a swap among k-th-score ties only
            kth = s1[b, order[k - 1]].item() if k > 0 else float("inf")
            assert set(got) ^ set(want) <= {i for i in range(n) if abs(s1[b, i].item() - kth) <= abs(kth) * 2 ** -7}, f"sample {b}"
        if k > 0:
            kth = torch.sort(s_cpu[b], descending=True).values[k - 1].item()
            band = abs(kth) * 2 ** -6  # two bf16 ulps
            must = set(torch.nonzero(s_cpu[b] > kth + band).flatten().tolist())
            may = set(torch.nonzero(s_cpu[b] >= kth - band).flatten().tolist())
            assert must <= set(got) <= may, f"sample {b}: selection differs from the CPU oracle beyond ties / one bf16 ulp"
    return True


@pytest.mark.parametrize("B,T,n_img,H,rate", [(2, 12, 16, 256, 0.5), (8, 32, 576, 4096, 0.2), (3, 32, 576, 512, 0.1), (2, 32, 576, 512, 1.0),
                                              (4, 1473, 576, 5120, 0.2), (16, 32, 576, 4096, 0.5)])  # + config 4 per GPU, config 5 (keep 0.5)
def test_select_compact(L, B, T, n_img, H, rate):
    s0 = 5
    Pn = T - 1 + n_img
    k = int(n_img * rate)
    logits = rnd(B, n_img, 2, seed=18)
    hid = rnd(B, Pn, H, seed=19)
    Pp = Pn - n_img + k
    out = torch.empty(B, Pp, H, device=DEV, dtype=BF)
    keep = torch.empty(B, k, dtype=torch.int64, device=DEV)
    pos = torch.empty(B, Pp, dtype=torch.int32, device=DEV)
    cu = torch.empty(B + 1, dtype=torch.int32, device=DEV)
    L.select_compact(P(logits.to(DEV)), None, P(hid.to(DEV)), P(out), P(keep), P(pos), P(cu), B, Pn, H, s0, n_img, k, 0, ST())
    torch.cuda.synchronize()
    keep_c = keep.cpu()
    _tie_aware_same(keep_c, logits, k)
    # untied case -> identical to the oracle's argsort-based selection
    ref_keep = O.topk_keep_index(logits, k)
    score = F.log_softmax(logits, dim=-1)[:, :, 0]
    if all(len(torch.unique(score[b])) == n_img for b in range(B)):
        assert torch.equal(keep_c, ref_keep)
    # gather / positions / CSR header follow the GPU's own keep set exactly (dynamic_modeling_llama.py:1935-1983)
    gathered = hid[:, s0:s0 + n_img].gather(1, keep_c[..., None].expand(B, k, H))
    ref_hid = torch.cat([hid[:, :s0], gathered, hid[:, s0 + n_img:]], dim=1)
    ref_pos = torch.cat([torch.arange(0, s0).repeat(B, 1), keep_c + s0, torch.arange(s0 + n_img, Pn).repeat(B, 1)], dim=1)
    assert torch.equal(out.cpu(), ref_hid)
    assert torch.equal(pos.cpu().long(), ref_pos)
    assert cu.tolist() == [b * Pp for b in range(B + 1)]


def test_select_compact_forced_and_ties(L):
    B, n_img, H, s0, T = 2, 576, 256, 5, 32
    Pn = T - 1 + n_img
    k = 115
    logits = torch.zeros(B, n_img, 2, dtype=BF)  # all tied -> stable order keeps the first k
    hid = rnd(B, Pn, H, seed=20)
    out = torch.empty(B, Pn - n_img + k, H, device=DEV, dtype=BF)
    keep = torch.empty(B, k, dtype=torch.int64, device=DEV)
    pos = torch.empty(B, Pn - n_img + k, dtype=torch.int32, device=DEV)
    cu = torch.empty(B + 1, dtype=torch.int32, device=DEV)
    L.select_compact(P(logits.to(DEV)), None, P(hid.to(DEV)), P(out), P(keep), P(pos), P(cu), B, Pn, H, s0, n_img, k, 0, ST())
    torch.cuda.synchronize()
    assert keep.cpu().tolist() == [list(range(k))] * B
    g = torch.Generator().manual_seed(21)
    forced = torch.stack([torch.sort(torch.randperm(n_img, generator=g)[:k]).values for _ in range(B)])
    L.select_compact(P(logits.to(DEV)), P(forced.to(DEV)), P(hid.to(DEV)), P(out), P(keep), P(pos), P(cu), B, Pn, H, s0, n_img, k, 7, ST())
    torch.cuda.synchronize()
    assert torch.equal(keep.cpu(), forced)
    assert torch.equal(out.cpu()[:, s0:s0 + k], hid[:, s0:s0 + n_img].gather(1, forced[..., None].expand(B, k, H)))


# ---------------------------------------------------------------------------------------------- sampling
_hf_warped_probs = O.hf_warped_probs  # pinned on CPU against the installed transformers' warpers (tests/test_oracle_golden.py)


def _sample(L, logits, step, seed, temperature, top_k, top_p, want_probs=True, eos=-1):
    B, V = logits.shape
    lg = logits.to(DEV)
    nxt = torch.full((B,), -7, device=DEV, dtype=torch.int64)
    toks = torch.full((B, 64), -7, device=DEV, dtype=torch.int64)
    stepc = torch.tensor([step], device=DEV, dtype=torch.int32)
    fin = torch.zeros(B, device=DEV, dtype=torch.int32)
    sd = torch.tensor([seed], device=DEV, dtype=torch.int64)
    probs = torch.full((B, V), float("nan"), device=DEV, dtype=torch.float32) if want_probs else None
    L.sample_next(P(lg), B, V, temperature, top_k, top_p, P(sd), P(nxt), P(toks), 64, P(stepc), P(fin), eos, 0, P(probs), ST())
    torch.cuda.synchronize()
    assert torch.equal(toks[:, step], nxt)
    return nxt.cpu(), (probs.cpu() if want_probs else None)


@pytest.mark.parametrize("temperature,top_k,top_p", [(0.7, 50, 0.9), (1.0, 0, 1.0), (0.2, 50, 1.0), (1.3, 1000, 0.5), (1.0, 1, 1.0), (2.0, 7, 0.3)])
def test_sample_next_distribution_matches_hf_warpers(L, temperature, top_k, top_p):
    """the distribution the token is drawn from == temperature / top-k / top-p warpers + softmax of HF 4.37.2, restated in torch"""
    logits = rnd(4, 32000, seed=160, scale=3.0)
    tok, probs = _sample(L, logits, 3, 1234, temperature, top_k, top_p)
    ref = _hf_warped_probs(logits, temperature, top_k, top_p)
    assert torch.isfinite(probs).all() and (probs.sum(-1) - 1).abs().max() < 1e-4
    # bf16 logits tie a lot and HF's ascending torch.sort breaks ties arbitrarily, so WHICH of several equal-score tokens falls under
    # the top-p cut is not defined: the sorted probability vectors must agree, and tokens on which the supports differ must be exact
    # ties (same logit) of a token at the edge of the support
    assert (probs.sort(-1, descending=True).values - ref.sort(-1, descending=True).values).abs().max().item() < 2e-5
    lf = logits.float()
    for r in range(logits.shape[0]):
        sup_k, sup_r = probs[r] > 0, ref[r] > 0
        assert int(sup_k.sum()) == int(sup_r.sum())
        diff = (sup_k ^ sup_r).nonzero().flatten()
        if len(diff):
            edge = lf[r][sup_r].min()
            assert (lf[r][diff] == edge).all(), (r, diff.tolist())
    assert (probs.gather(1, tok[:, None]) > 0).all()
    if top_k == 1:
        assert torch.equal(tok, logits.float().argmax(-1))


def test_sample_next_frequencies_and_determinism(L):
    """10 240 draws (256 identical rows x 40 steps) follow the warped distribution; the stream is a pure function of (seed, step, row)"""
    g = torch.Generator().manual_seed(161)
    row = (torch.randn(1, 1000, generator=g) * 2.0).to(BF)
    logits = row.repeat(256, 1)
    ref = _hf_warped_probs(row, 0.8, 20, 0.95)[0]
    counts = torch.zeros(1000)
    draws = []
    for step in range(40):
        tok, _ = _sample(L, logits, step, 99, 0.8, 20, 0.95, want_probs=False)
        draws.append(tok)
        counts += torch.bincount(tok, minlength=1000).float()
    n = 256 * 40
    assert counts[ref == 0].sum() == 0
    sigma = torch.sqrt(n * ref * (1 - ref)).clamp_min(1.0)
    assert ((counts - n * ref).abs() / sigma).max().item() < 5.0
    again, _ = _sample(L, logits, 7, 99, 0.8, 20, 0.95, want_probs=False)
    assert torch.equal(again, draws[7])
    other, _ = _sample(L, logits, 7, 100, 0.8, 20, 0.95, want_probs=False)
    assert not torch.equal(other, draws[7])
    assert len(set(draws[7].tolist())) > 3  # rows draw independently


def test_sample_next_eos_bookkeeping(L):
    logits = torch.full((3, 4096), -20.0, dtype=BF)
    logits[:, 2] = 20.0  # every row draws token 2 == eos
    lg = logits.to(DEV)
    nxt = torch.zeros(3, device=DEV, dtype=torch.int64)
    toks = torch.zeros(3, 8, device=DEV, dtype=torch.int64)
    stepc = torch.tensor([1], device=DEV, dtype=torch.int32)
    fin = torch.tensor([0, 1, 0], device=DEV, dtype=torch.int32)
    sd = torch.tensor([5], device=DEV, dtype=torch.int64)
    L.sample_next(P(lg), 3, 4096, 1.0, 50, 1.0, P(sd), P(nxt), P(toks), 8, P(stepc), P(fin), 2, 9, None, ST())
    torch.cuda.synchronize()
    assert nxt.tolist() == [2, 9, 2] and fin.tolist() == [1, 1, 1] and toks[:, 1].tolist() == [2, 9, 2]


# ---------------------------------------------------------------------------------------------- K4+K5 rope + append
def test_rope_append_prefill_and_decode(L):
    B, Lq, nh, hd, cap = 2, 9, 2, 128, 32
    H = nh * hd
    qkv = rnd(B * Lq, 3 * H, seed=22)
    posv = torch.tensor([[0, 1, 2, 3, 4, 17, 30, 31, 32], [0, 1, 2, 3, 4, 9, 11, 33, 34]], dtype=torch.int64)
    cos, sin = O.rope_tables(hd, 4096, 10000.0, BF)
    q = qkv[:, :H].view(B, Lq, nh, hd).transpose(1, 2)
    k = qkv[:, H:2 * H].view(B, Lq, nh, hd).transpose(1, 2)
    v = qkv[:, 2 * H:].view(B, Lq, nh, hd).transpose(1, 2)
    qr, kr = O.apply_rope(q, k, cos, sin, posv)
    qd = qkv.to(DEV).clone()
    K = torch.zeros(B, nh, cap, hd, device=DEV, dtype=BF)
    V = torch.zeros_like(K)
    kvl = torch.tensor([0, 0], dtype=torch.int32, device=DEV)
    cu = torch.tensor([0, Lq, 2 * Lq], dtype=torch.int32, device=DEV)
    L.rope_append(P(qd), P(posv.reshape(-1).to(torch.int32).to(DEV)), P(cu), P(kvl), None, P(cos.to(DEV)), P(sin.to(DEV)), P(K), P(V),
                  None, None, B * Lq, B, nh, hd, cap, ST())
    torch.cuda.synchronize()
    assert torch.equal(qd[:, :H].cpu().view(B, Lq, nh, hd).transpose(1, 2), qr)
    assert torch.equal(K[:, :, :Lq].cpu(), kr) and torch.equal(V[:, :, :Lq].cpu(), v)
    assert (K[:, :, Lq:] == 0).all()
    # decode: sample 0 keeps, sample 1 drops; both expose the new token in the side buffers
    qkv1 = rnd(B, 3 * H, seed=23)
    p1 = torch.tensor([[40], [41]], dtype=torch.int64)
    q1 = qkv1[:, :H].view(B, 1, nh, hd).transpose(1, 2)
    k1 = qkv1[:, H:2 * H].view(B, 1, nh, hd).transpose(1, 2)
    v1 = qkv1[:, 2 * H:].view(B, 1, nh, hd).transpose(1, 2)
    q1r, k1r = O.apply_rope(q1, k1, cos, sin, p1)
    kvl = torch.tensor([Lq, Lq], dtype=torch.int32, device=DEV)
    keep = torch.tensor([1, 0], dtype=torch.uint8, device=DEV)
    ks = torch.empty(B, nh, hd, device=DEV, dtype=BF)
    vs = torch.empty_like(ks)
    q1d = qkv1.to(DEV).clone()
    cu1 = torch.tensor([0, 1, 2], dtype=torch.int32, device=DEV)
    L.rope_append(P(q1d), P(p1.reshape(-1).to(torch.int32).to(DEV)), P(cu1), P(kvl), P(keep), P(cos.to(DEV)), P(sin.to(DEV)), P(K), P(V),
                  P(ks), P(vs), B, B, nh, hd, cap, ST())
    torch.cuda.synchronize()
    assert torch.equal(ks.cpu(), k1r[:, :, 0]) and torch.equal(vs.cpu(), v1[:, :, 0])
    assert torch.equal(K[0, :, Lq].cpu(), k1r[0, :, 0]) and (K[1, :, Lq] == 0).all() and (V[1, :, Lq] == 0).all()
    # split-K partial variant gives the same result when the slices sum to the same bf16 values
    part = torch.zeros(2, B, 3 * H, dtype=torch.float32)
    part[0] = qkv1.float() * 0.5
    part[1] = qkv1.float() * 0.5
    qo = torch.empty(B, H, device=DEV, dtype=BF)
    K2, V2 = torch.zeros_like(K), torch.zeros_like(V)
    L.dec_reduce_rope_append(P(part.to(DEV)), 2, P(qo), P(p1.reshape(-1).to(torch.int32).to(DEV)), P(kvl), P(keep), P(cos.to(DEV)),
                             P(sin.to(DEV)), P(K2), P(V2), P(ks), P(vs), B, nh, hd, cap, ST())
    torch.cuda.synchronize()
    assert torch.equal(qo.cpu().view(B, 1, nh, hd).transpose(1, 2), q1r)
    assert torch.equal(K2[0, :, Lq].cpu(), k1r[0, :, 0]) and (K2[1] == 0).all()


def test_kv_advance(L):
    Lr, B, SL = 4, 3, 2
    kv = torch.zeros(Lr, B, dtype=torch.int32, device=DEV)
    pos = torch.zeros(B, dtype=torch.int32, device=DEV)
    step = torch.zeros(1, dtype=torch.int32, device=DEV)
    L.kv_advance(P(kv), P(pos), None, None, Lr, B, SL, 27, 19, ST())  # prefill: P=27, P'=19
    keep = torch.tensor([1, 0, 1], dtype=torch.uint8, device=DEV)
    L.kv_advance(P(kv), P(pos), P(step), P(keep), Lr, B, SL, 1, 1, ST())
    torch.cuda.synchronize()
    assert kv.tolist() == [[28] * 3, [28] * 3, [20, 19, 20], [20, 19, 20]]
    assert pos.tolist() == [28] * 3 and step.item() == 1


# ---------------------------------------------------------------------------------------------- attention
def _sdpa_ref(q, k, v, causal_off=None):
    """fp32 attention on bf16 inputs. q [Lq,hd], k/v [Lk,hd]; causal_off: query i sees keys <= i + off"""
    s = (q.float() @ k.float().T) / math.sqrt(q.shape[-1])
    if causal_off is not None:
        i = torch.arange(q.shape[0])[:, None]
        j = torch.arange(k.shape[0])[None, :]
        s = s.masked_fill(j > i + causal_off, float("-inf"))
    return torch.softmax(s, dim=-1) @ v.float()


@pytest.mark.parametrize("lens,n_split,quirk", [([146, 210, 64, 1], 1, 0), ([735, 650, 129, 300], 2, 0), ([40, 70, 70, 3], 1, 1),
                                                ([0, 5, 0, 9], 1, 0), ([1500, 900, 1100, 64], 4, 1)])
def test_attn_decode(L, lens, n_split, quirk):
    B, nh, hd = len(lens), 4, 128
    cap = max(lens) + 8
    K = torch.zeros(B, nh, cap, hd, dtype=BF)
    V = torch.zeros_like(K)
    for b, n in enumerate(lens):
        K[b, :, :n] = rnd(nh, n, hd, seed=30 + b)
        V[b, :, :n] = rnd(nh, n, hd, seed=40 + b)
    q, ks, vs = rnd(B, nh * hd, seed=50), rnd(B, nh, hd, seed=51), rnd(B, nh, hd, seed=52)
    out = torch.empty(B, nh * hd, device=DEV, dtype=BF)
    ws = torch.empty(max(L.attn_decode_workspace_bytes(B, nh, n_split) // 4, 1), device=DEV, dtype=torch.float32)
    kvl = torch.tensor(lens, dtype=torch.int32, device=DEV)
    L.attn_decode(P(q.to(DEV)), nh * hd, P(K.to(DEV)), P(V.to(DEV)), P(ks.to(DEV)), P(vs.to(DEV)), P(kvl), P(out), P(ws), B, nh, hd, cap,
                  n_split, 1.0 / math.sqrt(hd), quirk, ST())
    torch.cuda.synchronize()
    mx = max(lens)
    for b, n in enumerate(lens):
        for h in range(nh):
            if quirk:  # reference B>1 semantics: cache zero-padded to the batch max, pad NOT masked (cache_utils.py:165-241)
                kk = torch.cat([K[b, h, :mx], ks[b, h][None]])
                vv = torch.cat([V[b, h, :mx], vs[b, h][None]])
            else:
                kk = torch.cat([K[b, h, :n], ks[b, h][None]])
                vv = torch.cat([V[b, h, :n], vs[b, h][None]])
            ref = _sdpa_ref(q[b, h * hd:(h + 1) * hd][None], kk, vv)[0]
            got = out[b, h * hd:(h + 1) * hd].float().cpu()
            assert (got - ref).abs().max() < 2e-2 * max(1.0, ref.abs().max()), (b, h, (got - ref).abs().max())


@pytest.mark.parametrize("seed", range(16))
def test_attn_decode_random_lengths(L, seed):
    """seeded random ragged lengths (0 .. 2100 keys per sample, up to 8 samples), every split count, both batching semantics"""
    import numpy as np
    r = np.random.Generator(np.random.PCG64(700 + seed))
    B = int(r.integers(1, 9))
    lens = [int(r.choice([0, 1, int(r.integers(2, 130)), int(r.integers(130, 2100))])) for _ in range(B)]
    test_attn_decode(L, lens, int(r.choice([1, 2, 3, 4, 8, 16])), int(r.integers(0, 2)))


@pytest.mark.parametrize("lens,n_split,S,quirk", [([146, 210, 64, 1], 1, 3, 0), ([735, 650, 129, 300], 2, 2, 0), ([40, 70, 70, 3], 1, 1, 1)])
def test_attn_decode_fused_equals_unfused(L, lens, n_split, S, quirk):
    """dl_attn_decode_fused (reduce + RoPE + append inside the attention CTA) must be bit-identical to
    dl_dec_reduce_rope_append followed by dl_attn_decode, including the arena contents."""
    B, nh, hd = len(lens), 4, 128
    H = nh * hd
    cap = max(lens) + 8
    K = torch.zeros(B, nh, cap, hd, dtype=BF)
    V = torch.zeros_like(K)
    for b, n in enumerate(lens):
        K[b, :, :n] = rnd(nh, n, hd, seed=230 + b)
        V[b, :, :n] = rnd(nh, n, hd, seed=240 + b)
    part = rnd(S, B, 3 * H, seed=250, dtype=torch.float32).to(DEV)
    cos, sin = O.rope_tables(hd, 4096, 10000.0, BF)
    cos, sin = cos.to(DEV), sin.to(DEV)
    pos = torch.tensor([n + 400 for n in lens], dtype=torch.int32, device=DEV)
    kvl = torch.tensor(lens, dtype=torch.int32, device=DEV)
    keep = torch.tensor([1, 0, 1, 0], dtype=torch.uint8, device=DEV)
    sc = 1.0 / math.sqrt(hd)
    ws = torch.empty(max(L.attn_decode_workspace_bytes(B, nh, n_split) // 4, 1), device=DEV, dtype=torch.float32)
    K1, V1 = K.to(DEV).clone(), V.to(DEV).clone()
    q = torch.empty(B, H, device=DEV, dtype=BF)
    ks, vs = torch.empty(B, nh, hd, device=DEV, dtype=BF), torch.empty(B, nh, hd, device=DEV, dtype=BF)
    o1 = torch.empty(B, H, device=DEV, dtype=BF)
    L.dec_reduce_rope_append(P(part), S, P(q), P(pos), P(kvl), P(keep), P(cos), P(sin), P(K1), P(V1), P(ks), P(vs), B, nh, hd, cap, ST())
    L.attn_decode(P(q), H, P(K1), P(V1), P(ks), P(vs), P(kvl), P(o1), P(ws), B, nh, hd, cap, n_split, sc, quirk, ST())
    K2, V2 = K.to(DEV).clone(), V.to(DEV).clone()
    o2 = torch.empty(B, H, device=DEV, dtype=BF)
    pf = torch.zeros(1 << 20, device=DEV, dtype=torch.uint8)
    L.attn_decode_fused(P(part), S, P(pos), P(keep), P(cos), P(sin), P(K2), P(V2), P(kvl), P(o2), P(ws), B, nh, hd, cap, n_split, sc, quirk,
                        P(pf) if S > 1 else None, pf.numel(), (min(lens) // 64) * 64 if S == 3 else 0, ST())
    torch.cuda.synchronize()
    assert torch.equal(K1, K2) and torch.equal(V1, V2)
    assert torch.equal(o1, o2), (o1.float() - o2.float()).abs().max()
    for b in range(B):  # appended iff kept, at slot len_b
        assert bool((K2[b, :, lens[b]] != 0).any()) == bool(keep[b])


def _varlen_case(L, fn, hd, causal, q_lens, pasts, nh=3):
    """build arena-layout K/V with `past` cached keys + the new q_len keys; compare with fp32 attention per sample"""
    B = len(q_lens)
    cap = max(p + n for p, n in zip(pasts, q_lens)) + 5
    H = nh * hd
    N = sum(q_lens)
    qkv = rnd(N, 3 * H, seed=60)
    K = torch.zeros(B, nh, cap, hd, dtype=BF)
    V = torch.zeros_like(K)
    cu = [0]
    for n in q_lens:
        cu.append(cu[-1] + n)
    for b in range(B):
        tot = pasts[b] + q_lens[b]
        K[b, :, :tot] = rnd(nh, tot, hd, seed=70 + b)
        V[b, :, :tot] = rnd(nh, tot, hd, seed=80 + b)
    out = torch.zeros(N, H, device=DEV, dtype=BF)
    cu_d = torch.tensor(cu, dtype=torch.int32, device=DEV)
    past_d = torch.tensor(pasts, dtype=torch.int32, device=DEV)
    fn(qkv.to(DEV), K.to(DEV), V.to(DEV), out, cu_d, past_d, N, B, nh, hd, cap, max(q_lens), causal)
    torch.cuda.synchronize()
    worst = 0.0
    for b in range(B):
        tot = pasts[b] + q_lens[b]
        for h in range(nh):
            qq = qkv[cu[b]:cu[b + 1], h * hd:(h + 1) * hd]
            ref = _sdpa_ref(qq, K[b, h, :tot], V[b, h, :tot], causal_off=pasts[b] if causal else None)
            got = out[cu[b]:cu[b + 1], h * hd:(h + 1) * hd].float().cpu()
            err = (got - ref).abs().max().item()
            worst = max(worst, err)
            assert err < 2e-2 * max(1.0, ref.abs().max().item()), (b, h, err)
    return worst


def _call_tc4(L, flags):
    def fn(qkv, K, V, out, cu, past, N, B, nh, hd, cap, max_q, causal):
        assert causal and hd == 128
        H = nh * hd
        L.attn_prefill_tc4(P(qkv), 3 * H, P(K), P(V), P(out), H, P(cu), P(past), N, B, nh, cap, max_q, 1.0 / math.sqrt(hd), flags, ST())
    return fn


@pytest.mark.parametrize("flags", [0, 1])
@pytest.mark.parametrize("q_lens,pasts", [([128], [0]), ([256], [0]), ([1], [0]), ([63, 64, 65], [0, 0, 0]), ([146, 146], [0, 0]),
                                          ([607, 33, 129, 300], [0, 0, 0, 0]), ([19, 70], [27, 5]), ([1587], [0]), ([300], [500]),
                                          ([200, 257], [64, 63])])
def test_attn_prefill_tc4_persistent(L, q_lens, pasts, flags):
    """K2 v4 (persistent): 64-key half tiles, two S buffers per group; flags=1 evaluates a quarter of the exp2 with the FMA
    polynomial."""
    _varlen_case(L, _call_tc4(L, flags), 128, True, q_lens, pasts, nh=2)


@pytest.mark.parametrize("seed", range(16))
def test_attn_prefill_tc4_random_ragged(L, seed):
    """seeded random ragged batches: 1-6 samples, 1-700 query rows each, 0-600 cached rows each (second-turn shape), both exp2 modes"""
    import numpy as np
    r = np.random.Generator(np.random.PCG64(800 + seed))
    B = int(r.integers(1, 7))
    q_lens = [int(r.choice([1, int(r.integers(2, 130)), int(r.integers(130, 700))])) for _ in range(B)]
    pasts = [int(r.choice([0, 0, int(r.integers(1, 600))])) for _ in range(B)]
    _varlen_case(L, _call_tc4(L, int(r.integers(0, 2))), 128, True, q_lens, pasts, nh=int(r.integers(1, 4)))


@pytest.mark.parametrize("B,nh,Lq,Lk", [(2, 4, 300, 300), (1, 2, 64, 200), (3, 2, 129, 129)])
def test_integration_stub_sdpa_prefill(L, B, nh, Lq, Lk):
    """the kernel-level binding shown in INTEGRATION.md: the reference's own tensors (query_states as a transpose(1, 2) view of the
    token-major projection, contiguous [B, heads, Lk, 128] keys/values) go in unchanged and the result equals causal SDPA with the
    query block aligned to the END of the keys (dynamic_modeling_llama.py:1114-1122)"""
    hd = 128
    g = torch.Generator().manual_seed(7)
    proj = (torch.randn(B, Lq, nh * hd, generator=g)).to(BF).to(DEV)          # what q_proj returns
    q = proj.view(B, Lq, nh, hd).transpose(1, 2)                              # query_states
    k = torch.randn(B, nh, Lk, hd, generator=g).to(BF).to(DEV)
    v = torch.randn(B, nh, Lk, hd, generator=g).to(BF).to(DEV)
    q2 = q.transpose(1, 2).reshape(B * Lq, nh * hd)
    assert q2.data_ptr() == proj.data_ptr()                                   # no copy
    out = torch.empty_like(q2)
    cu = torch.arange(0, (B + 1) * Lq, Lq, dtype=torch.int32, device=DEV)
    past = torch.full((B,), Lk - Lq, dtype=torch.int32, device=DEV)
    L.attn_prefill_tc4(P(q2), nh * hd, P(k), P(v), P(out), nh * hd, P(cu), P(past), B * Lq, B, nh, Lk, Lq, hd ** -0.5, 0, ST())
    torch.cuda.synchronize()
    got = out.view(B, Lq, nh, hd).transpose(1, 2).float()
    mask = torch.ones(Lq, Lk, dtype=torch.bool, device=DEV).tril(diagonal=Lk - Lq)
    ref = F.scaled_dot_product_attention(q.float(), k.float(), v.float(), attn_mask=mask)
    assert (got - ref).abs().max().item() < 2e-2


def test_integration_stub_linear_and_mlp(L):
    """the kernel-level bindings shown in INTEGRATION.md for nn.Linear and LlamaMLP (+ residual): the reference's own tensors go in
    unchanged ([B, L, H] activations, nn.Linear weights, gate|up concatenated once) and the results equal the reference's op sequence
    `residual + down(act_fn(gate(x)) * up(x))` (dynamic_modeling_llama.py:299-330, :1295) evaluated in fp32 with its bf16 rounding points"""
    B, Lq, H, I = 2, 150, 512, 1024
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, Lq, H, generator=g).to(BF).to(DEV)
    res = torch.randn(B, Lq, H, generator=g).to(BF).to(DEV)
    wg, wu, wd = [(torch.randn(*s_, generator=g) * 0.05).to(BF).to(DEV) for s_ in ((I, H), (I, H), (H, I))]
    bias = (torch.randn(H, generator=g) * 0.1).to(BF).to(DEV)
    ws = _linear_ws(L)
    # linear(x, weight, bias, act=1): Linear + GELU as in the predictors / projector
    x2 = x.reshape(-1, H)
    y = torch.empty(B * Lq, H, device=DEV, dtype=BF)
    wsq = (torch.randn(H, H, generator=g) * 0.05).to(BF).to(DEV)
    L.linear(P(x2), P(wsq), P(bias), None, P(y), B * Lq, H, H, 1, P(ws), ST())
    want = F.gelu((x2.float() @ wsq.float().T + bias.float()).to(BF).float()).to(BF)
    # mlp(x, w_gate_up, w_down, residual)
    wgu = torch.cat([wg, wu], 0).contiguous()
    act = torch.empty(B * Lq, I, device=DEV, dtype=BF)
    L.linear_swiglu(P(x2), P(wgu), P(act), B * Lq, I, H, P(ws), ST())
    out = res.clone().reshape(-1, H)
    L.linear(P(act), P(wd), None, P(out), P(out), B * Lq, H, I, 0, P(ws), ST())
    torch.cuda.synchronize()
    assert (y.float() - want.float()).abs().max() < 2e-2 * max(1.0, want.float().abs().max().item())
    a_ref = (F.silu((x2.float() @ wg.float().T).to(BF).float()).to(BF).float() * (x2.float() @ wu.float().T).to(BF).float()).to(BF)
    o_ref = (res.reshape(-1, H).float() + (a_ref.float() @ wd.float().T).to(BF).float()).to(BF)
    assert (out.float() - o_ref.float()).abs().max() < 2e-2 * max(1.0, o_ref.float().abs().max().item())


def test_attn_prefill_full_size_config4_against_sdpa(L):
    """BASELINE config 4 per GPU at full size (13B: 40 heads, B=4, L=2048 and the post-compaction L=1587): the persistent kernel against
    torch's fp32 causal SDPA of the same bf16 operands on the device, and the size-independent row property softmax-weights-sum-to-one
    (V = 1 -> O = 1)."""
    nh, hd, B = 40, 128, 4
    H = nh * hd
    for Lq in (2048, 1587):
        cap, N = Lq + 3, B * Lq
        g = torch.Generator(device=DEV).manual_seed(Lq)
        qkv = torch.randn(N, 3 * H, device=DEV, dtype=BF, generator=g)
        K = torch.randn(B, nh, cap, hd, device=DEV, dtype=BF, generator=g)
        V = torch.randn(B, nh, cap, hd, device=DEV, dtype=BF, generator=g)
        cu = torch.arange(0, (B + 1) * Lq, Lq, device=DEV, dtype=torch.int32)
        past = torch.zeros(B, device=DEV, dtype=torch.int32)
        o4 = torch.empty(N, H, device=DEV, dtype=BF)
        sc = 1.0 / math.sqrt(hd)
        L.attn_prefill_tc4(P(qkv), 3 * H, P(K), P(V), P(o4), H, P(cu), P(past), N, B, nh, cap, Lq, sc, 0, ST())
        torch.cuda.synchronize()
        q = qkv[:, :H].view(B, Lq, nh, hd).transpose(1, 2).float()
        ref = F.scaled_dot_product_attention(q, K[:, :, :Lq].float(), V[:, :, :Lq].float(), is_causal=True)
        got = o4.view(B, Lq, nh, hd).transpose(1, 2).float()
        assert (got - ref).abs().max().item() < 2e-2
        del q, ref, got
        ones = torch.ones_like(V)
        L.attn_prefill_tc4(P(qkv), 3 * H, P(K), P(ones), P(o4), H, P(cu), P(past), N, B, nh, cap, Lq, sc, 1, ST())
        torch.cuda.synchronize()
        assert (o4.float() - 1.0).abs().max().item() < 1e-2


def test_attn_prefill_tc4_more_than_63_samples(L):
    """B >= 64: the CSR header no longer fits the kernel's shared-memory copy and is read from global memory per item"""
    q_lens = [1 + (7 * i) % 40 for i in range(70)]
    pasts = [(3 * i) % 11 for i in range(70)]
    _varlen_case(L, _call_tc4(L, 0), 128, True, q_lens, pasts, nh=2)


@pytest.mark.timeout(180)
def test_attn_prefill_tc4_repeated_launches_do_not_deadlock(L):
    """regression: the anti-phase start of group B once waited on a parity of a multi-phase barrier and deadlocked when the issuer
    arrived late (about one launch in twenty at full size).  400 launches on the workload's own prefill shape must all finish."""
    B, nh, hd, Lq = 8, 32, 128, 607
    H, cap, N = nh * hd, Lq + 1, B * Lq
    qkv = torch.randn(N, 3 * H, device=DEV, dtype=BF)
    K = torch.randn(B, nh, cap, hd, device=DEV, dtype=BF)
    V = torch.randn(B, nh, cap, hd, device=DEV, dtype=BF)
    o = torch.empty(N, H, device=DEV, dtype=BF)
    cu = torch.arange(0, (B + 1) * Lq, Lq, device=DEV, dtype=torch.int32)
    past = torch.zeros(B, device=DEV, dtype=torch.int32)
    for i in range(400):
        L.attn_prefill_tc4(P(qkv), 3 * H, P(K), P(V), P(o), H, P(cu), P(past), N, B, nh, cap, Lq, 1.0 / math.sqrt(hd), i & 1, ST())
        if i % 50 == 49:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    assert torch.isfinite(o.float()).all()


@pytest.mark.parametrize("flags", [0, 1])
def test_attn_prefill_tc4_persistent_many_items(L, flags):
    """more work items than SMs, ragged lengths and pasts: every CTA of the persistent kernel walks several items, some skipped,
    some with an idle group B"""
    q_lens = [300, 1, 129, 256, 700, 64, 513, 90, 257, 128]
    pasts = [0, 40, 3, 0, 0, 200, 0, 70, 0, 1]
    _varlen_case(L, _call_tc4(L, flags), 128, True, q_lens, pasts, nh=40)


@pytest.mark.parametrize("flags", [0, 1])
@pytest.mark.parametrize("B,N,nh", [(2, 577, 16), (3, 576, 8), (2, 17, 2), (1, 300, 4), (2, 64, 1), (1, 129, 2)])
def test_attn_packed_tc4_noncausal_hd64(L, B, N, nh, flags):
    """CLIP (577 tokens, 16 heads) and predictor (576 tokens, 8 heads) attention on the persistent pipeline vs fp32 attention."""
    hd = 64
    C = nh * hd
    qkv = rnd(B * N, 3 * C, seed=91)
    out = torch.full((B * N, C), float("nan"), device=DEV, dtype=BF)
    cu = torch.arange(0, (B + 1) * N, N, dtype=torch.int32, device=DEV)
    L.attn_packed_tc4(P(qkv.to(DEV)), P(out), P(cu), B * N, B, nh, hd, N, 1.0 / math.sqrt(hd), flags, ST())
    torch.cuda.synchronize()
    got = out.float().cpu()
    assert torch.isfinite(got).all()
    for b in range(B):
        for h in range(nh):
            sl = slice(b * N, (b + 1) * N)
            ref = _sdpa_ref(qkv[sl, h * hd:(h + 1) * hd], qkv[sl, C + h * hd:C + (h + 1) * hd], qkv[sl, 2 * C + h * hd:2 * C + (h + 1) * hd])
            assert (got[sl, h * hd:(h + 1) * hd] - ref).abs().max() < 2e-2, (b, h)


@pytest.mark.parametrize("flags", [0, 1])
def test_attn_prefill_tc4_rescale_path(L, flags):
    """keys whose scores grow by far more than 2^8 along the sequence force the lazy O-rescale branch"""
    B, nh, hd, Lq = 1, 1, 128, 512
    H = nh * hd
    qkv = torch.zeros(Lq, 3 * H, dtype=BF)
    qkv[:, 0] = 8.0  # q = 8 * e0
    K = torch.zeros(B, nh, Lq + 4, hd, dtype=BF)
    V = rnd(B, nh, Lq + 4, hd, seed=95)
    K[0, 0, :Lq, 0] = torch.linspace(-30, 30, Lq).to(BF)  # logits = 64 * k0 / sqrt(128): span ~ +-170
    out = torch.zeros(Lq, H, device=DEV, dtype=BF)
    cu = torch.tensor([0, Lq], dtype=torch.int32, device=DEV)
    past = torch.zeros(1, dtype=torch.int32, device=DEV)
    L.attn_prefill_tc4(P(qkv.to(DEV)), 3 * H, P(K.to(DEV)), P(V.to(DEV)), P(out), H, P(cu), P(past), Lq,
                                                                   B, nh, Lq + 4, Lq, 1.0 / math.sqrt(hd), flags, ST())
    torch.cuda.synchronize()
    ref = _sdpa_ref(qkv[:, :hd], K[0, 0, :Lq], V[0, 0, :Lq], causal_off=0)
    assert (out.float().cpu() - ref).abs().max() < 3e-2 * max(1.0, ref.abs().max().item())


# ---------------------------------------------------------------------------------------------- decode weight streaming
@pytest.mark.parametrize("M,N,K,splits", [(8, 768, 256, 1), (8, 512, 512, 2), (1, 12288, 4096, 3), (8, 4096, 11008, 9), (16, 22016, 4096, 2),
                                          (3, 32000, 4096, 1), (24, 4096, 4096, 8), (64, 1024, 4096, 4)])
def test_skinny_gemm_tcgen05(L, M, N, K, splits):
    x, w = rnd(M, K, seed=100), rnd(N, K, seed=101, scale=0.05)
    part = torch.full((splits, M, N), float("nan"), device=DEV, dtype=torch.float32)
    L.skinny_gemm(P(x.to(DEV)), P(w.to(DEV)), P(part), M, N, K, splits, -1 if splits % 2 else 0, ST())
    torch.cuda.synchronize()
    got = part.sum(0).cpu()
    ref = x.float() @ w.float().T
    assert torch.isfinite(got).all()
    err = (got - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), err  # fp32 accumulation of exact bf16 products


@pytest.mark.parametrize("seed", range(24))
def test_skinny_gemm_random_shapes(L, seed):
    """seeded random decode-GEMM shapes: 1-64 rows, N any multiple of 8, 1-172 k-blocks, split count from the launcher's own heuristic
    or forced"""
    import numpy as np
    r = np.random.Generator(np.random.PCG64(900 + seed))
    M = int(r.integers(1, 65))
    N = 8 * int(r.choice([int(r.integers(1, 64)), int(r.integers(64, 4000))]))
    K = 64 * int(r.choice([int(r.integers(1, 9)), int(r.integers(9, 173))]))
    kb = K // 64
    splits = int(r.choice([1, L.skinny_gemm_splits(N, K, 296), int(r.integers(1, min(kb, 16) + 1))]))
    splits = max(1, min(splits, kb))
    test_skinny_gemm_tcgen05(L, M, N, K, splits)


@pytest.mark.parametrize("M,I,K", [(8, 512, 256), (8, 11008, 4096), (3, 13824, 5120), (24, 1024, 512)])
def test_skinny_gemm_swiglu_epilogue(L, M, I, K):
    """gate|up GEMM with SwiGLU fused into the TMEM epilogue == split-K GEMM + dl_dec_reduce_swiglu (same fp32 sums)."""
    x, w = rnd(M, K, seed=140), rnd(2 * I, K, seed=141, scale=0.05)
    xd, wd = x.to(DEV), w.to(DEV)
    act = torch.empty(M, I, device=DEV, dtype=BF)
    L.skinny_gemm_swiglu(P(xd), P(wd), P(act), M, I, K, ST())
    part = torch.empty(1, M, 2 * I, device=DEV, dtype=torch.float32)
    L.skinny_gemm(P(xd), P(wd), P(part), M, 2 * I, K, 1, 0, ST())
    ref = torch.empty(M, I, device=DEV, dtype=BF)
    L.dec_reduce_swiglu(P(part), P(ref), M, I, 1, ST())
    torch.cuda.synchronize()
    assert torch.equal(act, ref), (act.float() - ref.float()).abs().max()
    gu = (x.float() @ w.float().T)
    want = (F.silu(gu[:, :I].to(BF)) * gu[:, I:].to(BF)).float()
    assert (act.float().cpu() - want).abs().max() < 2e-2 * max(1.0, want.abs().max().item())


def _linear_ws(L):
    return torch.zeros(L.linear_workspace_bytes(), device=DEV, dtype=torch.uint8)


def _linear_ws_flag_offset(L):
    """int32 index where the flag words start: slots * 2*4*256*32 floats of parked accumulators come first (csrc/gemm2.cu)"""
    n = L.linear_workspace_bytes()
    slots = n // (2 * 4 * 256 * 32 * 4 + 16 * 4)
    return slots * 2 * 4 * 256 * 32


def _linear_ref(x, w, b, act, res):
    """fp32 evaluation of the same bf16 inputs with the reference's rounding points (bf16 nn.Linear output, then the next bf16 op)"""
    y = x.float() @ w.float().T
    if b is not None:
        y = y + b.float()
    y = y.to(BF)
    if act == 1:
        return F.gelu(y.float()).to(BF)
    if act == 2:  # HF quick_gelu on bf16 tensors
        return (y * torch.sigmoid((1.702 * y.float()).to(BF).float()).to(BF)).to(BF)
    if res is not None:
        return (res.float() + y.float()).to(BF)
    return y


# (M, N, K, act, bias, residual): the prefill / CLIP / predictor / projector shapes (some shrunk along M), both tile widths (N = 4096 at
# M = 1168 picks BN = 128), ragged M (1, 54, 300: partial and fully out-of-bounds half tiles), N not a multiple of the tile, K = 64
@pytest.mark.parametrize("M,N,K,act,bias,res", [
    (4608, 1024, 640, 0, False, False),    # CLIP patch embedding (im2col GEMM)
    (1154, 3072, 1024, 0, True, False),    # CLIP q|k|v
    (1154, 1024, 1024, 0, True, True),     # CLIP out_proj + residual
    (1154, 4096, 1024, 2, True, False),    # CLIP fc1 + quick_gelu
    (1154, 1024, 4096, 0, True, True),     # CLIP fc2 + residual
    (1152, 4096, 1024, 1, True, False),    # projector fc1 + GELU
    (1168, 12288, 4096, 0, False, False),  # prefill q|k|v (layers >= 2 of config 2)
    (1168, 4096, 4096, 0, False, True),    # o_proj + residual (BN = 128)
    (1168, 4096, 11008, 0, False, True),   # down_proj + residual, long K
    (1214, 32000, 512, 0, False, False),   # lm_head over all positions (K shrunk)
    (1152, 512, 4096, 1, True, False),     # predictor down_mlp
    (1152, 1536, 512, 0, False, False),    # predictor block qkv (no bias)
    (1152, 256, 512, 1, True, False),      # predictor output_mlp.0
    (1152, 128, 256, 1, True, False),      # predictor output_mlp.2 (N = 128: one half-tile wide)
    (300, 1000, 192, 0, True, False),      # nothing aligned
    (54, 768, 256, 0, False, False),       # toy sizes: the second CTA of the pair is fully out of bounds
    (1, 64, 64, 0, True, True),            # one row, one k-block
    (257, 8, 64, 1, True, False),          # N = 8
])
@pytest.mark.parametrize("stream_k", [False, True])
def test_linear_2cta_tcgen05(L, M, N, K, act, bias, res, stream_k):
    ws = _linear_ws(L) if stream_k else None
    x, w = rnd(M, K, seed=150), rnd(N, K, seed=151, scale=0.05)
    b = rnd(N, seed=152, scale=0.1) if bias else None
    r = rnd(M, N, seed=153) if res else None
    if res:  # in place, as the engine uses it (residual stream updated by the epilogue)
        out = r.to(DEV).clone()
        L.linear(P(x), P(w), P(b), P(out), P(out), M, N, K, act, P(ws), ST())
    else:
        out = torch.full((M, N), float("nan"), device=DEV, dtype=BF)
        L.linear(P(x), P(w), P(b), None, P(out), M, N, K, act, P(ws), ST())
    torch.cuda.synchronize()
    if ws is not None:
        assert int(ws.view(torch.int32)[_linear_ws_flag_offset(L):].abs().sum()) == 0  # flags handed back zeroed
    ref = _linear_ref(x, w, b, act, r).float()
    got = out.float().cpu()
    assert torch.isfinite(got).all()
    tol = 1.5e-2 * max(1.0, ref.abs().max().item())
    assert (got - ref).abs().max() < tol, ((got - ref).abs().max().item(), tol)
    # and against the library GEMM on the device: the same bf16 rounding of the same fp32 sums -> at most 1 ulp apart
    lib = F.linear(x.to(DEV), w.to(DEV), None if b is None else b.to(DEV))
    if act == 0 and not res:
        ulp_close(out, lib.cpu(), max_frac=0.02, abs_ok=2e-3)


@pytest.mark.parametrize("seed", range(40))
def test_linear_random_shapes(L, seed):
    """seeded random (M, N, K, epilogue) sweep of dl_linear with the stream-K workspace supplied, so the launcher's own heuristic picks
    the schedule: unaligned M, N down to 8, one to forty k-blocks, every epilogue"""
    import numpy as np
    r = np.random.Generator(np.random.PCG64(500 + seed))
    M = int(r.choice([int(r.integers(1, 64)), int(r.integers(64, 600)), int(r.integers(600, 2600))]))
    N = 8 * int(r.choice([int(r.integers(1, 16)), int(r.integers(16, 200)), int(r.integers(200, 700))]))
    K = 64 * int(r.integers(1, 41))
    act = int(r.integers(0, 3))
    bias = bool(r.integers(0, 2))
    res = act == 0 and bool(r.integers(0, 2))
    x, w = rnd(M, K, seed=seed).to(DEV), rnd(N, K, seed=seed + 1, scale=0.05).to(DEV)
    b = rnd(N, seed=seed + 2, scale=0.1).to(DEV) if bias else None
    rr = rnd(M, N, seed=seed + 3).to(DEV) if res else None
    ws = _linear_ws(L)
    out = rr.clone() if res else torch.full((M, N), float("nan"), device=DEV, dtype=BF)
    L.linear(P(x), P(w), P(b), P(out) if res else None, P(out), M, N, K, act, P(ws), ST())
    torch.cuda.synchronize()
    assert int(ws.view(torch.int32)[_linear_ws_flag_offset(L):].abs().sum()) == 0
    ref = _linear_ref(x, w, b, act, rr).float()
    got = out.float()
    assert torch.isfinite(got).all(), (M, N, K, act, bias, res)
    tol = 1.5e-2 * max(1.0, ref.abs().max().item())
    assert (got - ref).abs().max().item() < tol, (M, N, K, act, bias, res, (got - ref).abs().max().item(), tol)


def test_linear_repeated_launches_and_determinism(L):
    """persistent pairs + double-buffered TMEM: many tiles per pair, launched repeatedly on one stream, bit-identical every time"""
    M, N, K = 4856, 4096, 1024
    x, w = rnd(M, K, seed=160).to(DEV), rnd(N, K, seed=161, scale=0.05).to(DEV)
    outs = []
    ws = _linear_ws(L)
    for sk in (None, ws):
        outs = []
        for _ in range(4):
            o = torch.empty(M, N, device=DEV, dtype=BF)
            L.linear(P(x), P(w), None, None, P(o), M, N, K, 0, P(sk), ST())
            outs.append(o)
        torch.cuda.synchronize()
        for o in outs[1:]:
            assert torch.equal(o, outs[0])
        ref = (x.float() @ w.float().T)
        assert (outs[0].float() - ref).abs().max() < 1.5e-2 * ref.abs().max().item()


@pytest.mark.parametrize("B", [2, 3])
def test_linear_back_to_back_launches_share_the_stream_k_workspace(L, B):
    """regression (found at full 7B size with B = 2 / 3, never at the BASELINE batch sizes): dl_linear is launched with programmatic
    serialization and every launch of a stream shares one stream-K workspace, so the CTAs of launch n+1 can be resident while launch n
    is still parking / gathering through the same slots and flags; a finisher that polled the flags before observing its own accumulator
    consumed the previous launch's flags.  The GEMM sequence of 7B decoder layers with B*146 rows — q|k|v, o_proj (+x), RMSNorm,
    gate|up + SwiGLU, down_proj (+x), RMSNorm, ... — replayed from a CUDA graph (no host gaps between the launches) must equal the same
    sequence with a sync after every launch, bit for bit."""
    M, H, I = B * 146, 4096, 11008
    x0 = rnd(M, H, seed=180)
    wq, wo = rnd(3 * H, H, seed=181, scale=0.02).to(DEV), rnd(H, H, seed=182, scale=0.02).to(DEV)
    wgu, wd = rnd(2 * I, H, seed=183, scale=0.02).to(DEV), rnd(H, I, seed=184, scale=0.02).to(DEV)
    nw = torch.ones(H, dtype=BF, device=DEV)
    ws = _linear_ws(L)
    n_layers = 6

    def layers(x, sync):
        outs = []

        def step():
            if sync:
                torch.cuda.synchronize()
        for _ in range(n_layers):
            xn = torch.empty(M, H, device=DEV, dtype=BF)
            L.rmsnorm(P(x), P(nw), P(xn), M, H, 1e-5, ST()); step()
            qkv = torch.empty(M, 3 * H, device=DEV, dtype=BF)
            L.linear(P(xn), P(wq), None, None, P(qkv), M, 3 * H, H, 0, P(ws), ST()); step()
            attn = qkv[:, 2 * H:].contiguous(); step()  # stands in for the attention output (a torch copy kernel: a normal launch)
            L.linear(P(attn), P(wo), None, P(x), P(x), M, H, H, 0, P(ws), ST()); step()  # x += o_proj(attn), in place like the engine
            L.rmsnorm(P(x), P(nw), P(xn), M, H, 1e-5, ST()); step()
            act = torch.empty(M, I, device=DEV, dtype=BF)
            L.linear_swiglu(P(xn), P(wgu), P(act), M, I, H, P(ws), ST()); step()
            L.linear(P(act), P(wd), None, P(x), P(x), M, H, I, 0, P(ws), ST()); step()  # x += down_proj(act)
            outs += [qkv, act]
        return outs

    x_ref = x0.to(DEV).clone()
    ref = layers(x_ref, True)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        x_g = x0.to(DEV).clone()
        layers(x_g, False)  # warm-up: attribute setup, tensor-map cache
        x_g.copy_(x0.to(DEV))
        with torch.cuda.graph(g, stream=side):
            outs = layers(x_g, False)
    torch.cuda.current_stream().wait_stream(side)
    for rep in range(3):
        x_g.copy_(x0.to(DEV))
        g.replay()
        torch.cuda.synchronize()
        for i, (a, b) in enumerate(zip([x_g] + outs, [x_ref] + ref)):
            assert torch.equal(a, b), (rep, i, (a.float() - b.float()).abs().max().item())
    assert int(ws.view(torch.int32)[_linear_ws_flag_offset(L):].abs().sum()) == 0


@pytest.mark.parametrize("M,N,K,res", [(4608, 512, 4096, False), (4, 32000, 5120, False), (8192, 5120, 5120, True), (6348, 5120, 13824, True),
                                       (1168, 4096, 4096, True), (2308, 1024, 4096, True), (4608, 2048, 512, False), (1168, 12288, 4096, False)])
def test_linear_stream_k_bitwise_repeatable(L, M, N, K, res):
    """stream-K adds parked partials in a fixed order: 12 back-to-back launches (PDL-chained, sharing one workspace) of shapes whose tiles
    are cut once (last wave + remainder) or into several pieces (fewer tiles than pairs) give bit-identical results"""
    x, w = rnd(M, K, seed=190).to(DEV), rnd(N, K, seed=191, scale=0.05).to(DEV)
    r0 = rnd(M, N, seed=192).to(DEV) if res else None
    ws = _linear_ws(L)
    outs = []
    for _ in range(12):
        o = r0.clone() if res else torch.empty(M, N, device=DEV, dtype=BF)
        outs.append(o)
    for o in outs:
        L.linear(P(x), P(w), None, P(o) if res else None, P(o), M, N, K, 0, P(ws), ST())
    torch.cuda.synchronize()
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    ref = x.float() @ w.float().T
    if res:
        ref = ref.to(BF).float() + r0.float()
    assert (outs[0].float() - ref).abs().max() < 1.5e-2 * max(1.0, ref.abs().max().item())
    assert int(ws.view(torch.int32)[_linear_ws_flag_offset(L):].abs().sum()) == 0


def test_linear_stream_k_in_cuda_graph(L):
    """the parked-partial flags are reset by their reader, so a captured launch replays correctly any number of times"""
    M, N, K = 1168, 4096, 4096
    x, w = rnd(M, K, seed=162).to(DEV), rnd(N, K, seed=163, scale=0.05).to(DEV)
    ws = _linear_ws(L)
    o = torch.zeros(M, N, device=DEV, dtype=BF)
    L.linear(P(x), P(w), None, None, P(o), M, N, K, 0, P(ws), ST())
    torch.cuda.synchronize()
    first = o.clone()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(3):
                L.linear(P(x), P(w), None, None, P(o), M, N, K, 0, P(ws), ctypes.c_void_p(side.cuda_stream))
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(3):
        o.zero_()
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(o, first)
    assert int(ws.view(torch.int32)[_linear_ws_flag_offset(L):].abs().sum()) == 0


@pytest.mark.parametrize("M,I,K", [(1168, 11008, 4096), (300, 512, 256), (54, 640, 128), (2048, 13824, 320)])
@pytest.mark.parametrize("stream_k", [False, True])
def test_linear_swiglu_epilogue(L, M, I, K, stream_k):
    """gate|up GEMM with act_fn(gate) * up in the epilogue == the unfused pair (dl_linear + dl_swiglu) bit for bit, and the fp32 reference"""
    x, w = rnd(M, K, seed=170), rnd(2 * I, K, seed=171, scale=0.05)
    xd, wd = x.to(DEV), w.to(DEV)
    act = torch.full((M, I), float("nan"), device=DEV, dtype=BF)
    L.linear_swiglu(P(xd), P(wd), P(act), M, I, K, P(_linear_ws(L)) if stream_k else None, ST())
    gu = torch.empty(M, 2 * I, device=DEV, dtype=BF)
    L.linear(P(xd), P(wd), None, None, P(gu), M, 2 * I, K, 0, None, ST())
    ref = torch.empty(M, I, device=DEV, dtype=BF)
    L.swiglu(P(gu), P(ref), M, I, ST())
    torch.cuda.synchronize()
    assert torch.isfinite(act.float()).all()
    if stream_k:
        # a split K changes the fp32 summation order, so gate / up may round one bf16 ulp apart; silu amplifies that for gates in its
        # tail (d ln silu / dg -> 1: one ulp of g = -4 is 1 % of the product): rare, and bounded by 3 % + 2e-3
        da, dr = act.float(), ref.float()
        assert ((da - dr).abs() <= 0.03 * dr.abs() + 2e-3).all()
        assert (da != dr).float().mean().item() < 0.02
    else:
        # same rounded gate / up on both sides; the epilogue's sigmoid uses the fast exp (a few fp32 ulps from expf), which can flip a
        # last place of the 16-bit product: never seen in bf16 on these shapes, ~1e-4 of the elements in fp16
        ulp_close(act, ref, max_frac=1e-4)
    g32 = x.float() @ w.float().T
    want = (F.silu(g32[:, :I].to(BF)) * g32[:, I:].to(BF)).float()
    assert (act.float().cpu() - want).abs().max() < 2e-2 * max(1.0, want.abs().max().item())


@pytest.mark.parametrize("M,N,K", [(4608, 2, 128), (37, 2, 32), (5, 8, 64)])
def test_linear_tiny_heads(L, M, N, K):
    x, w, b = rnd(M, K, seed=180), rnd(N, K, seed=181, scale=0.1), rnd(N, seed=182, scale=0.1)
    out = torch.empty(M, N, device=DEV, dtype=BF)
    L.linear_tiny(P(x), P(w), P(b), P(out), M, N, K, ST())
    torch.cuda.synchronize()
    ref = (x.float() @ w.float().T + b.float()).to(BF)
    ulp_close(out, ref, max_frac=0.02)


def test_skinny_gemm_split_heuristic(L):
    for N, K in [(12288, 4096), (4096, 4096), (22016, 4096), (4096, 11008), (32000, 4096), (512, 4096), (768, 256)]:
        s = L.skinny_gemm_splits(N, K, 296)
        kb = K // 64
        assert 1 <= s <= min(kb, 8)


def test_decode_fused_consumers(L):
    M, H, I, S = 5, 512, 1024, 3
    part = (rnd(S, M, H, seed=110, dtype=torch.float32))
    x = rnd(M, H, seed=111)
    w = (1 + 0.1 * rnd(H, seed=112).float()).to(BF)
    gemm_out = part.sum(0).to(BF)
    x_ref = x + gemm_out
    xn_ref = O.rms_norm(x_ref, w, 1e-5)
    xd = x.to(DEV).clone()
    xn = torch.empty(M, H, device=DEV, dtype=BF)
    L.dec_reduce_residual_norm(P(part.to(DEV)), P(xd), P(w.to(DEV)), P(xn), M, H, S, 1e-5, ST())
    torch.cuda.synchronize()
    ulp_close(xd, x_ref, max_frac=0.01)
    ulp_close(xn, xn_ref, max_frac=0.02)
    pg = rnd(S, M, 2 * I, seed=113, dtype=torch.float32)
    gu = pg.sum(0).to(BF)
    act_ref = F.silu(gu[:, :I]) * gu[:, I:]
    act = torch.empty(M, I, device=DEV, dtype=BF)
    L.dec_reduce_swiglu(P(pg.to(DEV)), P(act), M, I, S, ST())
    y = torch.empty(M, 2 * I, device=DEV, dtype=BF)
    L.dec_reduce_bf16(P(pg.to(DEV)), P(y), M, 2 * I, S, ST())
    torch.cuda.synchronize()
    ulp_close(act, act_ref, max_frac=0.02)
    ulp_close(y, gu, max_frac=0.01)


def test_text_decision_matches_oracle(L):
    from dynamic_llava_b200.presets import preset
    from dynamic_llava_b200.synthetic import numpy_state_dict
    cfg = preset("toy")
    sd = {k: v.to(BF) for k, v in numpy_state_dict(cfg, 0).items()}
    B, H, D = 6, cfg["hidden_size"], cfg["sparse_config"]["d_model"]
    x = rnd(B, 1, H, seed=120)
    ref_logits = O.text_predictor(sd, x).reshape(B, 2)
    p = "model.output_text_score_predictor.output_mlp."
    ln = F.layer_norm(x.reshape(B, H), (H,), sd[p + "0.weight"], sd[p + "0.bias"], 1e-5)
    part = (ln.float() @ sd[p + "1.weight"].float().T)[None].contiguous()  # what the skinny GEMM hands over (1 split)
    keep = torch.zeros(B, dtype=torch.uint8, device=DEV)
    lo = torch.zeros(B, 2, device=DEV, dtype=BF)
    step = torch.tensor([1], dtype=torch.int32, device=DEV)
    d = lambda k: P(sd[p + k].to(DEV))
    L.text_decision(P(part.to(DEV)), 1, B, D, d("1.bias"), d("3.weight"), d("3.bias"), d("5.weight"), d("5.bias"), d("7.weight"), d("7.bias"),
                    None, P(step), -1, P(keep), P(lo), ST())
    torch.cuda.synchronize()
    assert (lo.float().cpu() - ref_logits.float()).abs().max() < 0.05 * max(1.0, ref_logits.float().abs().max())
    margin = (ref_logits[:, 0] - ref_logits[:, 1]).float().abs()
    ref_keep = ref_logits[:, 0] > ref_logits[:, 1]
    sure = margin > 0.1
    assert torch.equal(keep.cpu().bool()[sure], ref_keep[sure])
    forced = torch.tensor([[1, 0, 1, 0, 1, 0], [0, 0, 1, 1, 0, 1]], dtype=torch.uint8, device=DEV)
    step.fill_(2)
    L.text_decision(None, 0, B, D, None, None, None, None, None, None, None, P(forced), P(step), -1, P(keep), None, ST())
    torch.cuda.synchronize()
    assert keep.tolist() == forced[1].tolist()


# ---------------------------------------------------------------------------------------------- CLIP front end
def test_im2col_patch_embed_matches_conv(L):
    B, C, S, ps, Cout = 2, 3, 56, 14, 128
    pix = rnd(B, C, S, S, seed=130)
    w = rnd(Cout, C, ps, ps, seed=131, scale=0.05)
    ref = F.conv2d(pix.float(), w.float(), stride=ps).flatten(2).transpose(1, 2)  # [B, 16, Cout]
    K = C * ps * ps
    Kpad = (K + 63) // 64 * 64
    cols = torch.empty(B * 16, Kpad, device=DEV, dtype=BF)
    L.im2col_patches(P(pix.to(DEV)), P(cols), B, C, S, ps, Kpad, ST())
    torch.cuda.synchronize()
    wp = torch.zeros(Cout, Kpad)
    wp[:, :K] = w.float().reshape(Cout, K)
    got = (cols.float().cpu() @ wp.T).view(B, 16, Cout)
    assert (got - ref).abs().max() < 1e-3
    assert (cols[:, K:] == 0).all()
    pe = ref.to(BF).contiguous()  # conv output is a transposed view
    cls, pos = rnd(Cout, seed=132), rnd(17, Cout, seed=133, scale=0.02)
    tok = torch.empty(B, 17, Cout, device=DEV, dtype=BF)
    L.clip_assemble(P(pe.to(DEV)), P(cls.to(DEV)), P(pos.to(DEV)), P(tok), B, 16, Cout, ST())
    out = torch.empty(B, 16, Cout, device=DEV, dtype=BF)
    L.drop_cls(P(tok), P(out), B, 16, Cout, ST())
    torch.cuda.synchronize()
    ref_tok = torch.cat([cls.expand(B, 1, -1), pe], dim=1) + pos[None]
    assert torch.equal(tok.cpu(), ref_tok) and torch.equal(out.cpu(), ref_tok[:, 1:])


# ---------------------------------------------------------------------------------------------- instruct predictor (new turn)
@pytest.mark.parametrize("B,Tq,rate", [(1, 5, 0.5), (3, 70, 0.4), (2, 200, 0.9), (2, 1, 0.0), (3, 64, 0.0), (2, 129, 1.0)])
def test_turn_decision_compact_advance(L, B, Tq, rate):
    """cache_utils.py:165-241 via OracleCache.update(decision) per sample: slice to the true length, append the kept rows"""
    nh, hd, Lr, SL = 3, 128, 4, 2
    g = torch.Generator().manual_seed(7 * B + Tq)
    logits = torch.randn(B, Tq, 2, generator=g).to(BF)
    logits[:, :, 1] += (0.5 - rate) * 4
    if Tq > 3:
        logits[0, 2, 1] = logits[0, 2, 0]  # an exact tie is a drop (strict >)
    want = logits[..., 0] > logits[..., 1]
    want[:, -1:] = True  # dynamic_modeling_llama.py:2520
    keep = torch.zeros(B * Tq, dtype=torch.uint8, device=DEV)
    L.turn_decision(P(logits), None, P(keep), B, Tq, 1, ST())
    assert torch.equal(keep.cpu().bool().view(B, Tq), want)
    keep0 = torch.zeros(B * Tq, dtype=torch.uint8, device=DEV)
    L.turn_decision(P(logits), None, P(keep0), B, Tq, 0, ST())
    assert torch.equal(keep0.cpu().bool().view(B, Tq), logits[..., 0] > logits[..., 1])
    forced = (torch.rand(B, Tq, generator=g) < 0.5)
    keepf = torch.zeros(B * Tq, dtype=torch.uint8, device=DEV)
    L.turn_decision(None, P(forced.to(torch.uint8)), P(keepf), B, Tq, 1, ST())
    wf = forced.clone(); wf[:, -1] = True
    assert torch.equal(keepf.cpu().bool().view(B, Tq), wf)
    # arena with ragged cached lengths, new rows behind them
    lens = torch.tensor([11, 0, 37][:B], dtype=torch.int32)
    cap = int(lens.max()) + Tq + 5
    K = rnd(B, nh, cap, hd, seed=3)
    V = rnd(B, nh, cap, hd, seed=4)
    Kd, Vd = K.to(DEV), V.to(DEV)
    L.kv_compact_turn(P(Kd), P(Vd), P(lens), P(keep), Tq, B, nh, hd, cap, ST())
    kv = torch.zeros(Lr, B, dtype=torch.int32, device=DEV) + lens.to(DEV)
    pos = torch.full((B,), 100, dtype=torch.int32, device=DEV)
    L.kv_advance_turn(P(kv), P(pos), P(keep), Lr, B, SL, Tq, ST())
    torch.cuda.synchronize()
    cnt = want.sum(1).to(torch.int32)
    assert torch.equal(kv.cpu()[:SL], (lens + Tq).repeat(SL, 1)) and torch.equal(kv.cpu()[SL:], (lens + cnt).repeat(Lr - SL, 1))
    assert pos.tolist() == [100 + Tq] * B
    for b in range(B):
        n0 = int(lens[b])
        c = O.OracleCache()
        c.k, c.v, c.true_len = [K[b: b + 1, :, :n0].clone()], [V[b: b + 1, :, :n0].clone()], [torch.tensor([n0])]
        k_ref, v_ref = c.update(K[b: b + 1, :, n0: n0 + Tq], V[b: b + 1, :, n0: n0 + Tq], 0, want[b: b + 1])
        n1 = n0 + int(cnt[b])
        assert k_ref.shape[-2] == n1
        assert torch.equal(Kd[b, :, :n1].cpu(), k_ref[0]) and torch.equal(Vd[b, :, :n1].cpu(), v_ref[0])
        assert torch.equal(Kd[b, :, n0 + Tq:].cpu(), K[b, :, n0 + Tq:])  # nothing behind the new rows is touched


def test_gather_rows_pos(L):
    n, H = 57, 256
    x = rnd(n, H, seed=9)
    pos = torch.arange(n, dtype=torch.int32) * 3 + 1
    idx = torch.tensor([0, 1, 2, 5, 9, 10, 30, 56, 56, 3], dtype=torch.int64)
    out = torch.empty(len(idx), H, dtype=BF, device=DEV)
    po = torch.empty(len(idx), dtype=torch.int32, device=DEV)
    L.gather_rows_pos(P(x), P(pos), P(idx), P(out), P(po), len(idx), H, ST())
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), x[idx]) and torch.equal(po.cpu(), pos[idx])


@pytest.mark.parametrize("B,n", [(1, 1), (2, 5), (3, 255), (2, 256), (2, 257), (3, 1000), (1, 4000)])
def test_output_rank_and_keep_index(L, B, n):
    """no-KV-cache output thinning (dynamic_modeling_llama.py:2411-2426): count = #{logit0 > logit1}, n_keep = max over the batch,
    keep_index = sort(argsort(logit0, descending)[:, :n_keep]); ties resolved towards the lower index (a stable sort — torch's
    CPU argsort is not stable, the reference leaves ties undefined).  Exact."""
    g = torch.Generator().manual_seed(31 * B + n)
    logits = torch.randn(B, n, 2, generator=g).to(BF)  # bf16 randn: plenty of exact ties once n > ~100
    if n >= 5:
        logits[0, 3, 0] = logits[0, 1, 0]
        logits[0, 2, 1] = logits[0, 2, 0]  # logit0 == logit1 is a drop (strict >)
    rank = torch.empty(B, n, dtype=torch.int32, device=DEV)
    cnt = torch.empty(B, dtype=torch.int32, device=DEV)
    L.output_rank(P(logits), P(rank), P(cnt), B, n, ST())
    s0 = logits[..., 0].float()
    order = torch.sort(s0, dim=1, descending=True, stable=True).indices
    want_rank = torch.empty_like(order)
    want_rank.scatter_(1, order, torch.arange(n).repeat(B, 1))
    assert torch.equal(rank.cpu().long(), want_rank)
    want_cnt = (logits[..., 0] > logits[..., 1]).sum(1)
    assert torch.equal(cnt.cpu().long(), want_cnt)
    A, tail = 7, 1
    Lq = A + n + tail
    for n_keep in sorted({int(want_cnt.max()), 0, n, max(n // 2, 0)}):
        keep = torch.full((B, max(n_keep, 1)), -1, dtype=torch.int64, device=DEV)
        rows = torch.full((B, A + n_keep + 1), -1, dtype=torch.int64, device=DEV)
        L.output_keep_index(P(rank), B, n, n_keep, A, Lq, P(keep) if n_keep else None, P(rows), ST())
        want_keep = torch.sort(order[:, :n_keep], dim=1).values
        if n_keep:
            assert torch.equal(keep.cpu(), want_keep)
        base = (torch.arange(B) * Lq).view(B, 1)
        want_rows = torch.cat([torch.arange(A).repeat(B, 1), want_keep + A, torch.full((B, 1), Lq - 1)], dim=1) + base
        assert torch.equal(rows.cpu(), want_rows)


def test_output_keep_index_first_call_duplicates_last_token(L):
    """the reference's empty `[answer_indice:-1]` slice on the first call (Lq == A): left ++ last, i.e. the last token twice"""
    B, A = 2, 19
    rows = torch.full((B, A + 1), -1, dtype=torch.int64, device=DEV)
    L.output_keep_index(None, B, 0, 0, A, A, None, P(rows), ST())
    want = torch.cat([torch.arange(A), torch.tensor([A - 1])]).repeat(B, 1) + (torch.arange(B) * A).view(B, 1)
    assert torch.equal(rows.cpu(), want)
    with pytest.raises(RuntimeError):
        L.output_keep_index(None, B, 4, 5, A, A + 5, None, P(rows), ST())  # n_keep > n

