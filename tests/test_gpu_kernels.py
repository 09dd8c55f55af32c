"""GPU parity tests of the individual C-ABI entry points (include/kmbart.h), called through
ctypes exactly as the host engine calls them, against fp32 torch references of the same op on
the same (bf16-rounded) inputs.  Tolerances are the north_star's bf16 gates or tighter and are
written next to each check."""
import ctypes as C
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

import golden_cases as G  # noqa: E402

BF16, F32 = torch.bfloat16, torch.float32


@pytest.fixture(scope="module")
def lib():
    from kmbart import lib as L
    L.require_b200()
    return L


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return 0 if t is None else t.data_ptr()


def run_gemm(L, A, B, M, N, K, a_mn=0, b_mn=0, elt=0, tile_n=0, **kw):
    e = L.GemmEpilogue()
    e.mode, e.act, e.alpha = kw.get("mode", L.EPI_LINEAR), kw.get("act", L.ACT_NONE), kw.get("alpha", 1.0)
    e.accumulate = int(kw.get("accumulate", 0))
    e.bias = _p(kw.get("bias"))
    e.residual, e.ld_res = _p(kw.get("residual")), kw.get("ld_res", N)
    e.aux, e.ld_aux = _p(kw.get("aux")), kw.get("ld_aux", N)
    e.out_f32, e.ld_f32 = _p(kw.get("out_f32")), kw.get("ld_f32", N)
    e.out_bf16, e.ld_bf16 = _p(kw.get("out_bf16")), kw.get("ld_bf16", N)
    e.out_preact = _p(kw.get("out_preact"))
    e.dropout_p, e.dropout_tag, e.dropout_seed = kw.get("dropout_p", 0.0), kw.get("dropout_tag", 0), _p(kw.get("dropout_seed"))
    e.labels, e.ce_max, e.ce_sum = _p(kw.get("labels")), _p(kw.get("ce_max")), _p(kw.get("ce_sum"))
    e.ce_label_logit, e.ce_lse, e.ce_gscale = _p(kw.get("ce_label_logit")), _p(kw.get("ce_lse")), _p(kw.get("ce_gscale"))
    lda, ldb = A.stride(0), B.stride(0)
    L.check(L.load().kmb_gemm(A.data_ptr(), B.data_ptr(), M, N, K, lda, ldb, a_mn, b_mn, elt, C.byref(e), tile_n, _stream()), "kmb_gemm")
    torch.cuda.synchronize()


def rnd(*shape, dtype=BF16, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(dtype)


def pad_cols(t, mult=8):
    """row pitch must be a multiple of 16 bytes: return a view [rows, cols] of a padded buffer"""
    rows, cols = t.shape
    ld = (cols + mult - 1) // mult * mult
    buf = torch.zeros(rows, ld, dtype=t.dtype, device=t.device)
    buf[:, :cols] = t
    return buf[:, :cols]


# ------------------------------------------------------------------ GEMM (tcgen05 / TMEM / TMA)
@pytest.mark.parametrize("M,N,K,a_mn,b_mn,tile_n", [
    (1000, 520, 200, 0, 0, 0), (1000, 520, 200, 0, 1, 128), (1000, 264, 328, 1, 0, 256), (384, 768, 1000, 1, 1, 0),
    (300, 72, 200, 0, 0, 64), (4096, 768, 768, 0, 0, 0), (20000, 256, 64, 0, 0, 128), (17, 40, 24, 0, 0, 32),
    (6144, 2304, 768, 0, 0, 0), (128, 50320, 128, 0, 0, 256),
    # CTA-pair tiles (cta_group::2): 256 x {256, 192, 128}
    (1000, 520, 200, 0, 0, 1256), (1000, 520, 200, 0, 0, 1192), (1000, 520, 200, 0, 1, 1128), (1000, 264, 328, 1, 0, 1256),
    (384, 768, 1000, 1, 1, 1256), (12800, 768, 768, 0, 0, 1192), (300, 1000, 72, 0, 1, 1256), (6144, 768, 3072, 0, 1, 0),
])
def test_gemm_bf16_plain(lib, M, N, K, a_mn, b_mn, tile_n):
    a = rnd(M, K, seed=1)
    b = rnd(N, K, seed=2)
    A = pad_cols(a.t().contiguous()) if a_mn else pad_cols(a)
    B = pad_cols(b.t().contiguous()) if b_mn else pad_cols(b)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=F32)
    run_gemm(lib, A, B, M, N, K, a_mn, b_mn, 0, tile_n, out_f32=out)
    ref = a.float() @ b.float().t()
    # fp32 accumulation of exact bf16 products: only summation-order differences remain
    assert (out - ref).abs().max().item() <= 1e-3 * math.sqrt(K)


@pytest.mark.parametrize("tile_n", [0, 128, 1256, 1192])
def test_gemm_epilogue_bias_gelu_preact_residual(lib, tile_n):
    M, N, K = 1000, 520, 200
    a, w, bias, res = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=0.1), rnd(N, dtype=F32, seed=3), rnd(M, N, dtype=F32, seed=4)
    ref_pre = a.float() @ w.float().t() + bias
    # (a) bias + GELU, bf16 out + pre-activation copy through the TMA-store path
    out, pre = torch.zeros(M, N, device="cuda", dtype=BF16), torch.zeros(M, N, device="cuda", dtype=BF16)
    run_gemm(lib, a, w, M, N, K, tile_n=tile_n, bias=bias, act=lib.ACT_GELU, out_bf16=out, out_preact=pre)
    assert (pre.float() - ref_pre).abs().max().item() <= 2 ** -7 * ref_pre.abs().max().item()      # one bf16 rounding
    ref = torch.nn.functional.gelu(ref_pre)
    assert (out.float() - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item() + 1e-5
    # (b) bias + residual, fp32 out, then accumulate onto it
    o32 = torch.zeros(M, N, device="cuda", dtype=F32)
    run_gemm(lib, a, w, M, N, K, tile_n=tile_n, bias=bias, residual=res, out_f32=o32)
    assert (o32 - (ref_pre + res)).abs().max().item() <= 1e-4
    run_gemm(lib, a, w, M, N, K, tile_n=tile_n, alpha=0.5, out_f32=o32, accumulate=1)
    assert (o32 - (ref_pre + res + 0.5 * (ref_pre - bias))).abs().max().item() <= 2e-4
    # (c) GELU backward epilogue: dgrad * gelu'(u), u read from a bf16 aux matrix
    u = rnd(M, N, seed=7)
    du = torch.zeros(M, N, device="cuda", dtype=BF16)
    run_gemm(lib, a, w, M, N, K, tile_n=tile_n, act=lib.ACT_GELU_GRAD, aux=u, out_bf16=du)
    uf = u.float().requires_grad_(True)
    torch.nn.functional.gelu(uf).sum().backward()
    ref = (a.float() @ w.float().t()) * uf.grad
    assert (du.float() - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item() + 1e-4
    # (d) tanh / tanh-grad (BartClassificationHead)
    t = torch.zeros(M, N, device="cuda", dtype=BF16)
    run_gemm(lib, a, w, M, N, K, tile_n=tile_n, bias=bias, act=lib.ACT_TANH, out_bf16=t)
    assert (t.float() - torch.tanh(ref_pre)).abs().max().item() <= 2 ** -8 + 1e-5


def test_gemm_split_k_weight_gradient_shape(lib):
    """wgrad: dW[N_out, N_in] += dY^T X with tokens as the contraction (both operands MN-major), split-K reds."""
    T, No, Ni = 12800, 768, 768
    dy, x = rnd(T, No, seed=1, scale=0.05), rnd(T, Ni, seed=2)
    dw = torch.ones(No, Ni, device="cuda", dtype=F32)
    run_gemm(lib, dy, x, No, Ni, T, a_mn=1, b_mn=1, out_f32=dw, accumulate=1)
    ref = 1.0 + dy.float().t() @ x.float()
    assert (dw - ref).abs().max().item() <= 1e-3 * ref.abs().max().item()


def test_gemm_tf32_and_3xtf32(lib):
    M, N, K = 333, 200, 160
    a, b = rnd(M, K, dtype=F32, seed=1), rnd(N, K, dtype=F32, seed=2)
    out = torch.zeros(M, N, device="cuda", dtype=F32)
    run_gemm(lib, a, b, M, N, K, elt=1, out_f32=out)
    ref = a.double() @ b.double().t()
    assert (out - ref).abs().max().item() <= 2 ** -10 * 3 * math.sqrt(K) * 4      # tf32: 10-bit mantissa operands
    a3, b3 = torch.empty(M, 3 * K, device="cuda"), torch.empty(N, 3 * K, device="cuda")
    L = lib.load()
    lib.check(L.kmb_split_tf32(a.data_ptr(), K, a3.data_ptr(), M, K, 0, _stream()), "split")
    lib.check(L.kmb_split_tf32(b.data_ptr(), K, b3.data_ptr(), N, K, 1, _stream()), "split")
    run_gemm(lib, a3, b3, M, N, 3 * K, elt=1, out_f32=out)
    assert ((out - ref).abs().max() / ref.abs().max()).item() <= 1e-5               # fp32-class product on tensor cores


def test_gemm_rejects_bad_arguments(lib):
    a, b = rnd(64, 64), rnd(64, 64)
    out = torch.zeros(64, 64, device="cuda")
    with pytest.raises(lib.KmbartError):
        run_gemm(lib, a, b, 64, 64, 64, tile_n=48, out_f32=out)
    with pytest.raises(lib.KmbartError):
        run_gemm(lib, a[:, 1:], b, 64, 64, 63, out_f32=out)          # misaligned operand
    with pytest.raises(lib.KmbartError):
        run_gemm(lib, a, b, 64, 64, 64, out_f32=out, dropout_p=0.1)  # dropout without a device seed


def test_gemm_dropout_epilogue_is_deterministic_and_unbiased(lib):
    M, N, K = 512, 256, 64
    a, w = rnd(M, K, seed=1), rnd(N, K, seed=2)
    seed = torch.tensor([12345], dtype=torch.int64, device="cuda")
    o1, o2, o0 = (torch.zeros(M, N, device="cuda") for _ in range(3))
    run_gemm(lib, a, w, M, N, K, out_f32=o0)
    run_gemm(lib, a, w, M, N, K, out_f32=o1, dropout_p=0.1, dropout_tag=3, dropout_seed=seed)
    run_gemm(lib, a, w, M, N, K, out_f32=o2, dropout_p=0.1, dropout_tag=3, dropout_seed=seed)
    assert torch.equal(o1, o2)
    kept = o1 != 0
    assert abs(kept.float().mean().item() - 0.9) < 0.01
    assert torch.allclose(o1[kept], o0[kept] / 0.9, rtol=1e-6)


# ------------------------------------------------------------------ fused LM head + cross entropy
@pytest.mark.parametrize("tile_n", [256, 1256])
def test_lmhead_ce_forward_backward_never_materialises_logits(lib, tile_n):
    M, V, d = 300, 50320, 128
    h, E = rnd(M, d, seed=1), rnd(V, d, seed=2, scale=0.05)
    flb = rnd(V, dtype=F32, seed=3, scale=0.1)
    g = torch.Generator(device="cuda").manual_seed(4)
    labels = torch.randint(0, V, (M,), device="cuda", generator=g)
    labels[::7] = -100
    L = lib.load()
    nt = L.kmb_gemm_n_tiles(V, tile_n)
    ce_max, ce_sum = torch.empty(M, nt, device="cuda"), torch.empty(M, nt, device="cuda")
    lab_logit, lse = torch.zeros(M, device="cuda"), torch.empty(M, device="cuda")
    acc2, loss, total = torch.zeros(2, device="cuda"), torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    run_gemm(lib, h, E, M, V, d, tile_n=tile_n, mode=lib.EPI_CE_STATS, bias=flb, labels=labels, ce_max=ce_max, ce_sum=ce_sum,
             ce_label_logit=lab_logit)
    lib.check(L.kmb_ce_combine(ce_max.data_ptr(), ce_sum.data_ptr(), lab_logit.data_ptr(), labels.data_ptr(), M, nt, lse.data_ptr(),
                               0, acc2.data_ptr(), 1.0, loss.data_ptr(), total.data_ptr(), 0, _stream()), "ce_combine")
    hf = h.float().requires_grad_(True)
    logits = hf @ E.float().t() + flb
    ref = torch.nn.functional.cross_entropy(logits, labels, ignore_index=-100)
    assert abs(loss.item() - ref.item()) <= 1e-5 * ref.item()
    assert torch.allclose(lse, torch.logsumexp(logits, -1), atol=1e-4)
    # backward: dlogits (bf16) = (softmax - onehot) / n_valid
    gscale, up = torch.empty(1, device="cuda"), torch.full((1,), 2.0, device="cuda")
    lib.check(L.kmb_ce_gscale(acc2.data_ptr(), up.data_ptr(), 1.0, gscale.data_ptr(), _stream()), "gscale")
    dlog = torch.zeros(M, V, device="cuda", dtype=BF16)
    run_gemm(lib, h, E, M, V, d, tile_n=tile_n, mode=lib.EPI_CE_GRAD, bias=flb, labels=labels, ce_lse=lse, ce_gscale=gscale,
             out_bf16=dlog, ld_bf16=V)
    (2.0 * ref).backward()
    dref = torch.autograd.grad(2.0 * torch.nn.functional.cross_entropy(hf @ E.float().t() + flb, labels, ignore_index=-100), hf)[0]
    dh = dlog.float() @ E.float()
    assert ((dh - dref).norm() / dref.norm()).item() <= 1e-2


def test_small_xent_heads(lib):
    L = lib.load()
    n, Cc = 37, 129
    x = rnd(n, Cc, dtype=F32, seed=1)
    g = torch.Generator(device="cuda").manual_seed(2)
    labels = torch.randint(0, Cc, (n,), device="cuda", generator=g)
    soft = torch.softmax(rnd(n, Cc, dtype=F32, seed=3), -1)
    for mode in (0, 1):
        acc = torch.zeros(1, device="cuda")
        ld_d = 136
        d = torch.zeros(n, ld_d, device="cuda", dtype=BF16)
        up = torch.ones(1, device="cuda")
        lib.check(L.kmb_small_xent(x.data_ptr(), Cc, n, Cc, mode, labels.data_ptr(), soft.data_ptr(), Cc, 3.0, acc.data_ptr(),
                                   d.data_ptr(), ld_d, up.data_ptr(), _stream()), "small_xent")
        xr = x.clone().requires_grad_(True)
        if mode == 0:
            ref = 3.0 * torch.nn.functional.cross_entropy(xr, labels)
        else:
            ref = 3.0 * torch.nn.functional.kl_div(torch.log_softmax(xr, 1), soft, reduction="batchmean")
        ref.backward()
        assert abs(acc.item() - ref.item()) <= 1e-5 * abs(ref.item()) + 1e-6
        assert (d[:, :Cc].float() - xr.grad).abs().max().item() <= 2 ** -8 * xr.grad.abs().max().item() + 1e-6
        assert (d[:, Cc:] == 0).all()


# ------------------------------------------------------------------ attention
def ref_attention(q, k, v, pad, causal, scale):
    """q [B,H,Sq,64] fp32 ...; pad [B,Sk] bool"""
    w = (q @ k.transpose(-1, -2)) * scale
    if causal:
        Sq, Sk = w.shape[-2:]
        w = w + torch.triu(torch.full((Sq, Sk), float("-inf"), device=w.device), 1)
    if pad is not None:
        w = w.masked_fill(pad[:, None, None, :], float("-inf"))
    return torch.softmax(w, -1) @ v


@pytest.mark.parametrize("B,H,Sq,Sk,causal,use_pad", [(3, 2, 100, 100, 0, 1), (2, 12, 48, 48, 1, 1), (2, 3, 48, 100, 0, 1),
                                                      (1, 2, 7, 7, 1, 0), (2, 2, 130, 130, 1, 0), (2, 2, 86, 200, 0, 1),
                                                      (40, 12, 128, 128, 1, 1), (2, 4, 17, 33, 0, 1), (30, 12, 100, 100, 0, 1)])
@pytest.mark.parametrize("persist,split,tc05", [(0, 0, 1), (0, 0, 0), (1, 0, 0), (0, 1, 0)])
def test_attention_forward_backward(lib, monkeypatch, B, H, Sq, Sk, causal, use_pad, persist, split, tc05):
    """tc05 = 1: the tcgen05 / TMEM kernels of attention_tc05.cu for every Sq, Sk <= 128 (the product default takes them where
    they are faster: the backward of encoder-sized tiles); tc05 = 0: the mma.sync kernels (which longer sequences always use) — split = 1 forces their three-kernel backward (0: fused persistent
    backward when Sq, Sk <= 128), persist = 1 their opt-in persistent forward instead of the tiled one"""
    monkeypatch.setenv("KMBART_ATTN_BWD_SPLIT", str(split))
    monkeypatch.setenv("KMBART_ATTN_FWD_PERSIST", str(persist))
    monkeypatch.setenv("KMBART_ATTN_TC05", str(tc05))
    L = lib.load()
    d = H * 64
    q, k, v = rnd(B * Sq, d, seed=1), rnd(B * Sk, d, seed=2), rnd(B * Sk, d, seed=3)
    pad = None
    if use_pad:
        pad = torch.zeros(B, Sk, dtype=torch.bool, device="cuda")
        pad[0, Sk - 5:] = True
        pad[-1, Sk // 2:] = True
    pad_u8 = pad.to(torch.uint8).contiguous() if pad is not None else None
    o = torch.zeros(B * Sq, d, device="cuda", dtype=BF16)
    lse = torch.zeros(B * H * Sq, device="cuda")
    lib.check(L.kmb_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), d, d, d, o.data_ptr(), d, lse.data_ptr(), _p(pad_u8), B, H,
                             Sq, Sk, 64, causal, 0.125, _stream()), "attn_fwd")
    def heads(t, S):
        return t.float().view(B, S, H, 64).transpose(1, 2).detach().requires_grad_(True)
    qh, kh, vh = heads(q, Sq), heads(k, Sk), heads(v, Sk)
    ref = ref_attention(qh, kh, vh, pad, causal, 0.125)
    got = o.float().view(B, Sq, H, 64).transpose(1, 2)
    assert (got - ref).abs().max().item() <= 2e-2          # bf16 P and O roundings; values are O(1)
    do = rnd(B * Sq, d, seed=5)
    dq, dk, dv = (torch.zeros_like(t) for t in (q, k, v))
    dscr = torch.zeros(B * H * Sq, device="cuda")
    lib.check(L.kmb_attn_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), d, d, d, o.data_ptr(), d, do.data_ptr(), d, lse.data_ptr(),
                             dscr.data_ptr(), _p(pad_u8), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), d, d, d, B, H, Sq, Sk, 64,
                             causal, 0.125, _stream()), "attn_bwd")
    ref.backward(do.float().view(B, Sq, H, 64).transpose(1, 2))
    for got_t, ref_t, S in ((dq, qh.grad, Sq), (dk, kh.grad, Sk), (dv, vh.grad, Sk)):
        g = got_t.float().view(B, S, H, 64).transpose(1, 2)
        assert ((g - ref_t).norm() / ref_t.norm()).item() <= 2e-2


def test_attention_strided_cache_layout(lib):
    """decode step: q [n,1,H,64] token-major, K/V in the legacy cache layout [n, H, T, 64]"""
    L = lib.load()
    n, H, T = 5, 4, 13
    d = H * 64
    q = rnd(n, d, seed=1)
    K, V = rnd(n, H, T, 64, seed=2), rnd(n, H, T, 64, seed=3)
    o = torch.zeros(n, d, device="cuda", dtype=BF16)
    arr = (C.c_int64 * 12)(d, 64, 64, H * T * 64, T * 64, 64, H * T * 64, T * 64, 64, d, 64, 64)
    lib.check(L.kmb_attn_fwd_strided(q.data_ptr(), K.data_ptr(), V.data_ptr(), o.data_ptr(), arr, 0, n, H, 1, T, 64, 0, 0.125, _stream()), "attn")
    ref = ref_attention(q.float().view(n, 1, H, 64).transpose(1, 2), K.float(), V.float(), None, 0, 0.125)
    assert (o.float().view(n, 1, H, 64).transpose(1, 2) - ref).abs().max().item() <= 2e-2


# ------------------------------------------------------------------ LayerNorm / embedding
@pytest.mark.parametrize("M,d", [(1000, 768), (77, 1024), (300, 128)])
def test_layernorm_forward_backward(lib, M, d):
    L = lib.load()
    z, res = rnd(M, d, seed=1), rnd(M, d, dtype=F32, seed=2)
    gamma, beta = 1 + 0.1 * rnd(d, dtype=F32, seed=3), 0.1 * rnd(d, dtype=F32, seed=4)
    pre, y32, mean, rstd = torch.empty(M, d, device="cuda"), torch.empty(M, d, device="cuda"), torch.empty(M, device="cuda"), torch.empty(M, device="cuda")
    y16 = torch.empty(M, d, device="cuda", dtype=BF16)
    lib.check(L.kmb_layernorm_fwd(z.data_ptr(), res.data_ptr(), gamma.data_ptr(), beta.data_ptr(), pre.data_ptr(), y32.data_ptr(),
                                  y16.data_ptr(), mean.data_ptr(), rstd.data_ptr(), M, d, 0.0, 0, 0, _stream()), "ln_fwd")
    x = (res + z.float()).requires_grad_(True)
    gp, bp = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(x, (d,), gp, bp, 1e-5)
    assert torch.equal(pre, x.detach())
    assert (y32 - ref).abs().max().item() <= 2e-5
    assert (y16.float() - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item()
    dyA, dyB = rnd(M, d, dtype=F32, seed=5), rnd(M, d, seed=6)
    dpre = torch.empty(M, d, device="cuda")
    dz = torch.empty(M, d, device="cuda", dtype=BF16)
    dg, db, dbias = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
    lib.check(L.kmb_layernorm_bwd(dyA.data_ptr(), dyB.data_ptr(), pre.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                                  dpre.data_ptr(), dz.data_ptr(), dg.data_ptr(), db.data_ptr(), dbias.data_ptr(), M, d, 0.0, 0, 0.0, 0,
                                  0, _stream()), "ln_bwd")
    ref.backward(dyA + dyB.float())
    assert (dpre - x.grad).abs().max().item() <= 1e-4 * x.grad.abs().max().item() + 1e-5
    assert (dz.float() - x.grad).abs().max().item() <= 2 ** -7 * x.grad.abs().max().item()
    assert ((dg - gp.grad).norm() / gp.grad.norm()).item() <= 1e-4
    assert ((db - bp.grad).norm() / bp.grad.norm()).item() <= 1e-4
    assert ((dbias - x.grad.sum(0)).norm() / dbias.norm()).item() <= 1e-4   # summed in fp32 before the bf16 rounding of dz


def test_visual_token_embedding_fused(lib):
    """pack_features + slot_index + feature GEMM + embed_ln_fwd == ImageEmbedding + _embed_multi_modal + pos + LN
    (src/model/modules.py:24-41, :89-102, :133-137), boxes kept in fp32 (raw pixels)."""
    from oracle import kmbart_oracle as O
    L = lib.load()
    ocfg = G.small_config()
    sd = G.perturb(O.init_state_dict(ocfg, seed=0))
    batch = O.synthetic_batch(ocfg, batch=5, n_regions=6, n_ctx=14, tgt_len=4, seed=3, ragged=True)
    d, B, S = ocfg.d_model, *batch["input_ids"].shape
    feats = [f.cuda() for f in batch["image_features"]]
    counts = [f.shape[0] for f in feats]
    R = sum(counts)
    off = torch.tensor([0] + list(torch.tensor(counts).cumsum(0)), dtype=torch.int32, device="cuda")
    ptrs = torch.tensor([f.data_ptr() for f in feats], dtype=torch.int64, device="cuda")
    f16, boxes = torch.empty(R, 2048, device="cuda", dtype=BF16), torch.empty(R, 4, device="cuda")
    lib.check(L.kmb_pack_features(ptrs.data_ptr(), off.data_ptr(), B, 0, f16.data_ptr(), boxes.data_ptr(), R, _stream()), "pack")
    allf = torch.cat(feats, 0)
    assert torch.equal(f16, allf[:, :2048].to(BF16)) and torch.equal(boxes, allf[:, 2048:])
    ids = batch["input_ids"].cuda()
    slot = torch.empty(B * S, dtype=torch.int32, device="cuda")
    lib.check(L.kmb_slot_index(ids.data_ptr(), off.data_ptr(), B, S, ocfg.img_feat_id, ocfg.cls_token_id, slot.data_ptr(), _stream()), "slot")
    W = sd["model.encoder.embed_images.linear.weight"].cuda()
    wf16, wbox = torch.empty(d, 2048, device="cuda", dtype=BF16), torch.empty(d, 4, device="cuda")
    lib.check(L.kmb_repack_img_weight(W.data_ptr(), wf16.data_ptr(), wbox.data_ptr(), d, 2052, _stream()), "repack")
    vis = torch.empty(R, d, device="cuda")
    run_gemm(lib, f16, wf16, R, d, 2048, out_f32=vis)
    x32, x16 = torch.empty(B * S, d, device="cuda"), torch.empty(B * S, d, device="cuda", dtype=BF16)
    cu = {k: v.cuda() for k, v in sd.items()}
    lib.check(L.kmb_embed_ln_fwd(ids.data_ptr(), slot.data_ptr(), cu["model.shared.weight"].data_ptr(),
                                 cu["model.encoder.embed_positions.weight"].data_ptr(), vis.data_ptr(), boxes.data_ptr(), wbox.data_ptr(),
                                 cu["model.encoder.embed_images.linear.bias"].data_ptr(),
                                 cu["model.encoder.layernorm_embedding.weight"].data_ptr(),
                                 cu["model.encoder.layernorm_embedding.bias"].data_ptr(), 0, x32.data_ptr(), x16.data_ptr(), 0, 0,
                                 B * S, S, d, 2, 0, 1.0, 0.0, 0, 0, _stream()), "embed")
    torch.cuda.synchronize()
    emb = O.embed_multimodal(sd, ocfg, batch["input_ids"], batch["image_features"])
    pos = sd["model.encoder.embed_positions.weight"][torch.arange(S) + 2]
    ref = torch.nn.functional.layer_norm(emb + pos, (d,), sd["model.encoder.layernorm_embedding.weight"],
                                         sd["model.encoder.layernorm_embedding.bias"], 1e-5)
    err = (x32.cpu().view(B, S, d) - ref).abs().max().item()
    assert err <= 2e-2, err      # bf16 RoI features/weights on the K=2048 contraction (north_star bf16 gate 2e-2)
    tok_rows = (slot.view(B, S) < 0).cpu()
    assert (x32.cpu().view(B, S, d)[tok_rows] - ref[tok_rows]).abs().max().item() <= 1e-5   # token rows are exact fp32


# ------------------------------------------------------------------ AdamW (golden vector from the reference's optimizer)
def test_fused_adamw_matches_reference_golden(lib):
    import os
    from kmbart.optim import AdamW
    golden = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kmbart_reference_golden.pt"),
                        weights_only=False)["adamw"]
    params, grads_seq = G.case_adamw()
    ps = [torch.nn.Parameter(p.cuda()) for p in params]
    opt = AdamW(ps, lr=G.ADAMW["lr"], weight_decay=G.ADAMW["weight_decay"])
    for grads in grads_seq:
        for p, g in zip(ps, grads):
            p.grad = g.cuda()
        opt.step()
    torch.cuda.synchronize()
    for p, ref, m, mref, v, vref in zip(ps, golden["params"], [opt.state[p]["exp_avg"] for p in ps], golden["exp_avg"],
                                        [opt.state[p]["exp_avg_sq"] for p in ps], golden["exp_avg_sq"]):
        assert torch.allclose(p.detach().cpu(), ref, atol=1e-6, rtol=1e-5)
        assert torch.allclose(m.cpu(), mref, atol=1e-7, rtol=1e-5)
        assert torch.allclose(v.cpu(), vref, atol=1e-9, rtol=1e-5)
    assert opt.state[ps[0]]["step"] == 3 and opt.launches_last == 2
