"""GPU parity tests, kernel by kernel, through the C ABI (ctypes) against plain torch fp32
restatements / the oracle.  Run on a B200: `python -m pytest tests -m gpu`."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ctclip_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from ctclip_b200 import _lib
    _lib.require_device()
    return _lib


def dev():
    return torch.device("cuda")


def rnd(*shape, seed=0, scale=1.0, dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dev(), dtype)


def relerr(a, b):
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


# --------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("impl", [0, 1, 2, 3])     # 0 = tcgen05 product path, 1 = SIMT, 2 = single CTAs, 3 = CTA pairs
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 200, 136), (1000, 512, 512), (2048, 1408, 512),
                                   (13824, 512, 4000), (4096, 256, 512), (1536, 2816, 512), (512, 4000, 512),
                                   (777, 64, 256), (256, 512, 2816), (129, 256, 512), (38016, 512, 512)])
def test_gemm_bf16_out(lib, impl, M, N, K):
    a = rnd(M, K, seed=1, dtype=torch.bfloat16)
    w = rnd(N, K, seed=2, scale=1 / math.sqrt(K), dtype=torch.bfloat16)
    out = torch.full((M, N), float("nan"), device=dev(), dtype=torch.bfloat16)
    lib.call("ctc_gemm_bf16", a, K, w, K, out, N, M, N, K, lib.EPI_BF16, None, None, 0, None, 0, impl, lib.stream_ptr())
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t()
    assert torch.isfinite(out.float()).all()
    assert relerr(out, ref) < 1e-2


@pytest.mark.parametrize("impl", [0, 1, 2, 3])     # 0 = tcgen05 product path, 1 = SIMT, 2 = single CTAs, 3 = CTA pairs
@pytest.mark.parametrize("M,N,K", [(256, 512, 256), (999, 512, 1408), (13824, 512, 256), (130, 64, 256), (300, 256, 512), (4161, 1408, 512)])
def test_gemm_f32_bias_resid(lib, impl, M, N, K):
    a = rnd(M, K, seed=3, dtype=torch.bfloat16)
    w = rnd(N, K, seed=4, scale=1 / math.sqrt(K), dtype=torch.bfloat16)
    bias = rnd(N, seed=5)
    resid = rnd(M, N, seed=6)
    ref = a.float() @ w.float().t() + bias + resid
    out = resid.clone()   # in-place residual
    lib.call("ctc_gemm_bf16", a, K, w, K, out, N, M, N, K, lib.EPI_F32, bias, out, N, None, 0, impl, lib.stream_ptr())
    torch.cuda.synchronize()
    assert relerr(out, ref) < 2e-5 * math.sqrt(K)


def test_gemm_persistent_many_tiles_repeat(lib):
    """Many tiles per CTA (pipeline phase wrap) and back-to-back launches give identical results."""
    M, N, K = 110592 // 4, 512, 512
    a = rnd(M, K, seed=7, dtype=torch.bfloat16)
    w = rnd(N, K, seed=8, scale=0.05, dtype=torch.bfloat16)
    outs = []
    for _ in range(2):
        out = torch.empty(M, N, device=dev(), dtype=torch.bfloat16)
        lib.call("ctc_gemm_bf16", a, K, w, K, out, N, M, N, K, lib.EPI_BF16, None, None, 0, None, 0, 0, lib.stream_ptr())
        outs.append(out)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
    ref = a.float() @ w.float().t()
    assert relerr(outs[0], ref) < 1e-2


# --------------------------------------------------------------------------------------- LayerNorm
@pytest.mark.parametrize("R,C", [(1000, 512), (77, 64)])
def test_layernorm_fwd_bwd(lib, R, C):
    x = rnd(R, C, seed=1, scale=2.0) + 0.3
    g = 1 + 0.1 * rnd(C, seed=2)
    b = 0.1 * rnd(C, seed=3)
    y16 = torch.empty(R, C, device=dev(), dtype=torch.bfloat16)
    y32 = torch.empty(R, C, device=dev())
    raw = torch.empty(R, C, device=dev(), dtype=torch.bfloat16)
    lib.call("ctc_layernorm_fwd", x, R, C, g, b, 1e-5, y16, y32, raw, lib.stream_ptr())
    xr = x.clone().requires_grad_()
    ref = F.layer_norm(xr, (C,), g, b)
    assert relerr(y32, ref) < 1e-5
    assert relerr(y16, ref) < 1e-2
    assert torch.equal(raw, x.to(torch.bfloat16))
    dy = rnd(R, C, seed=4)
    (dx_ref,) = torch.autograd.grad(ref, xr, dy)
    prev = rnd(R, C, seed=5)
    out = prev.clone()
    o16 = torch.empty(R, C, device=dev(), dtype=torch.bfloat16)
    lib.call("ctc_layernorm_bwd", dy, x, R, C, g, 1e-5, out, 1, o16, lib.stream_ptr())
    assert relerr(out, dx_ref + prev) < 1e-5
    assert relerr(o16, dx_ref + prev) < 1e-2
    lib.call("ctc_layernorm_bwd", dy, x, R, C, g, 1e-5, out, 0, None, lib.stream_ptr())
    assert relerr(out, dx_ref) < 1e-5
    # incoming gradient in bf16 (what the dgrad GEMMs' bf16 epilogue hands over): exact w.r.t. the rounded dy
    dy16 = dy.to(torch.bfloat16)
    (dx_ref16,) = torch.autograd.grad(F.layer_norm(xr, (C,), g, b), xr, dy16.float())
    out = prev.clone()
    lib.call("ctc_layernorm_bwd_bf16", dy16, x, R, C, g, 1e-5, out, 1, o16, lib.stream_ptr())
    assert relerr(out, dx_ref16 + prev) < 1e-5
    lib.call("ctc_layernorm_bwd_bf16", dy16, x, R, C, g, 1e-5, out, 0, None, lib.stream_ptr())
    assert relerr(out, dx_ref16) < 1e-5


# --------------------------------------------------------------------------------------- PEG
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("B,T,H,W,C", [(2, 6, 6, 6, 64), (1, 24, 24, 24, 512), (2, 5, 5, 5, 66), (1, 7, 7, 7, 2)])
def test_peg_fwd_and_adjoint(lib, mode, B, T, H, W, C):
    x = rnd(B * T * H * W, C, seed=1)
    w = rnd(C, 1, 3, 3, 3, seed=2, scale=0.2)
    bias = rnd(C, seed=3, scale=0.1)
    w27 = w.reshape(C, 27).t().contiguous()
    y = torch.empty_like(x)
    lib.call("ctc_peg", x, B, T, H, W, C, w27, bias, mode, 0, y, None, lib.stream_ptr())
    xr = x.clone().requires_grad_()
    if mode == 0:
        xin = xr.reshape(B * T, H * W, C)
        ref = (O.peg(xin, (B, T, H, W), w, bias) + xin).reshape(-1, C)
    else:  # '(b h w) t d' view, same video_shape (the reference's scramble)
        xin = xr.reshape(B, T, H, W, C).permute(0, 2, 3, 1, 4).reshape(B * H * W, T, C)
        r = O.peg(xin, (B, T, H, W), w, bias) + xin
        ref = r.reshape(B, H, W, T, C).permute(0, 3, 1, 2, 4).reshape(-1, C)
    assert relerr(y, ref) < 1e-5
    dy = rnd(B * T * H * W, C, seed=4)
    (dx_ref,) = torch.autograd.grad(ref, xr, dy)
    dx = torch.empty_like(x)
    d16 = torch.empty(x.shape, device=dev(), dtype=torch.bfloat16)
    lib.call("ctc_peg", dy, B, T, H, W, C, w27, None, mode, 1, dx, d16, lib.stream_ptr())
    assert relerr(dx, dx_ref) < 1e-5
    assert relerr(d16, dx_ref) < 1e-2


@pytest.mark.parametrize("T,H,W,C", [(9, 6, 6, 64), (24, 24, 24, 512)])
def test_peg_frames_matches_dense_peg(lib, T, H, W, C):
    """ctc_peg_frames over a compact frame list (changed frames of two 'windows' + baseline frames for the rest)
    must reproduce exactly the dense spatial PEG of the volumes one would assemble by hand."""
    import numpy as np
    HW = H * W
    base = rnd(T * HW, C, seed=1)                      # baseline stream of one volume
    w27 = rnd(C, 27, seed=2, scale=0.2).t().contiguous()
    bias = rnd(C, seed=3, scale=0.1)
    wins = [(0, 2), (T - 3, 3), (3, 2)]                # (t0, n_prev changed input frames)
    n_prev = [n for _, n in wins]
    off_prev = np.concatenate([[0], np.cumsum(n_prev)[:-1]])
    changed = rnd(sum(n_prev) * HW, C, seed=4)
    src, expect = [], []
    for wi, (t0, n) in enumerate(wins):
        full = base.clone().view(T, HW, C)
        full[t0:t0 + n] = changed.view(-1, HW, C)[off_prev[wi]:off_prev[wi] + n]
        yfull = torch.empty(T * HW, C, device=dev())
        lib.call("ctc_peg", full.reshape(-1, C).contiguous(), 1, T, H, W, C, w27, bias, 0, 0, yfull, None, lib.stream_ptr())
        n_out = min(n + 2, T - t0)
        for i in range(n_out):
            row = []
            for dt in (-2, -1, 0):
                ip, tt = i + dt, t0 + i + dt
                row.append(-2 ** 31 if tt < 0 else (int(off_prev[wi] + ip) if 0 <= ip < n else -1 - tt))
            src.append(row)
            expect.append(yfull.view(T, HW, C)[t0 + i])
    src_t = torch.tensor(src, dtype=torch.int32, device=dev())
    F_ = len(src)
    y = torch.full((F_ * HW, C), float("nan"), device=dev())
    lib.call("ctc_peg_frames", changed, base, src_t, F_, H, W, C, w27, bias, y, lib.stream_ptr())
    torch.cuda.synchronize()
    assert torch.equal(y.view(F_, HW, C), torch.stack(expect))


def test_frames_gather_and_rows_fill(lib):
    a, b = rnd(5 * 96, 8, seed=1), rnd(7 * 96, 8, seed=2)
    src = torch.tensor([0, -1, 4, -7, 2], dtype=torch.int32, device=dev())
    out = torch.empty(5 * 96, 8, device=dev())
    lib.call("ctc_frames_gather", a, b, src, 5, 96 * 8, out, lib.stream_ptr())
    ref = torch.stack([a.view(5, -1)[0], b.view(7, -1)[0], a.view(5, -1)[4], b.view(7, -1)[6], a.view(5, -1)[2]])
    assert torch.equal(out.view(5, -1), ref)
    rows = torch.tensor([3, 100, 479], dtype=torch.int32, device=dev())
    val = rnd(8, seed=3)
    lib.call("ctc_rows_fill", out, rows, 3, 8, val, lib.stream_ptr())
    ref2 = ref.reshape(-1, 8).clone()
    ref2[rows.long()] = val
    assert torch.equal(out, ref2)


# --------------------------------------------------------------------------------------- attention
def _attn_ref(q, kv, qs, ks, scale, bias, B, T, H, W, heads, mode):
    """fp32 torch restatement of attention.py:144-180 on [R, heads*32] rows in canonical order."""
    R = q.shape[0]
    inner = heads * 32
    k, v = kv[:, :inner], kv[:, inner:]

    def seq(x):  # -> [n_seq, heads, n, 32]
        x = x.reshape(B, T, H * W, heads, 32)
        if mode == 0:
            return x.reshape(B * T, H * W, heads, 32).permute(0, 2, 1, 3)
        return x.permute(0, 2, 1, 3, 4).reshape(B * H * W, T, heads, 32).permute(0, 2, 1, 3)

    qh, kh, vh = seq(q), seq(k), seq(v)
    qn = O.l2norm(qh) * qs
    kn = O.l2norm(kh) * ks
    sim = qn @ kn.transpose(-1, -2) * scale
    if bias is not None:
        sim = sim + bias
    p = sim.softmax(dim=-1)
    o = p @ vh                                              # [n_seq, heads, n, 32]
    lse = torch.logsumexp(sim, dim=-1)                      # [n_seq, heads, n]

    def unseq(x, d):  # [n_seq, heads, n, d] -> [R, heads*d]
        if mode == 0:
            return x.permute(0, 2, 1, 3).reshape(R, heads * d)
        return x.permute(0, 2, 1, 3).reshape(B, H * W, T, heads * d).permute(0, 2, 1, 3).reshape(R, heads * d)

    return unseq(o, 32), unseq(lse.unsqueeze(-1), 1), p


@pytest.mark.parametrize("mode,B,T,H,W,heads", [(0, 1, 2, 24, 24, 8), (1, 1, 24, 24, 24, 8), (0, 2, 6, 6, 6, 2),
                                                (1, 2, 6, 6, 6, 2), (0, 1, 1, 10, 10, 4), (0, 1, 3, 8, 8, 2),
                                                (0, 2, 1, 16, 8, 1)])
def test_attention_fwd_bwd_probs(lib, mode, B, T, H, W, heads):
    _attention_fwd_bwd_probs(lib, mode, B, T, H, W, heads)


@pytest.mark.parametrize("B,T,H,W,heads", [(1, 2, 24, 24, 8), (1, 3, 8, 8, 2), (2, 1, 16, 8, 1)])
def test_attention_bwd_dq_on_tcgen05(lib, B, T, H, W, heads):
    """Same checks with dQ computed by the opt-in tcgen05 / TMEM kernel (ctc_attention_set_tc_bwd)."""
    prev = lib.load().ctc_attention_set_tc_bwd(1)
    try:
        _attention_fwd_bwd_probs(lib, 0, B, T, H, W, heads)
    finally:
        lib.load().ctc_attention_set_tc_bwd(prev)


@pytest.mark.parametrize("B,T,H,W,heads", [(1, 2, 24, 24, 8), (1, 3, 8, 8, 2), (2, 1, 16, 8, 1), (1, 1, 16, 24, 3)])
def test_attention_bwd_two_kernel_mma_sync_path(lib, B, T, H, W, heads):
    """The two mma.sync kernels (dQ, dK/dV) - the default until the one-pass tcgen05 kernel overtook them, still the path of
    every geometry that kernel does not take - forced with ctc_attention_set_tc_bwd(0)."""
    prev = lib.load().ctc_attention_set_tc_bwd(0)
    try:
        _attention_fwd_bwd_probs(lib, 0, B, T, H, W, heads)
    finally:
        lib.load().ctc_attention_set_tc_bwd(prev)


@pytest.mark.parametrize("B,T,H,W,heads", [(1, 2, 24, 24, 8), (1, 3, 8, 8, 2), (2, 1, 16, 8, 1), (1, 1, 16, 24, 3)])
def test_attention_bwd_one_pass_on_tcgen05(lib, B, T, H, W, heads):
    """dQ, dK and dV from ONE recomputation of the probabilities on tcgen05 / TMEM (attention_tc_bwd.cu), against the
    same torch fp32 reference as the two-kernel mma.sync path; twice, and bit-identical (no atomics)."""
    prev = lib.load().ctc_attention_set_tc_bwd(2)
    try:
        a = _attention_fwd_bwd_probs(lib, 0, B, T, H, W, heads)
        b = _attention_fwd_bwd_probs(lib, 0, B, T, H, W, heads)
        assert all(torch.equal(x, y) for x, y in zip(a, b))
    finally:
        lib.load().ctc_attention_set_tc_bwd(prev)


def _attention_fwd_bwd_probs(lib, mode, B, T, H, W, heads):
    if mode == 1 and not (T == H == W):
        pytest.skip("temporal mode uses T tokens")
    R, inner = B * T * H * W, heads * 32
    q = rnd(R, inner, seed=1, dtype=torch.bfloat16)
    kv = rnd(R, 2 * inner, seed=2, dtype=torch.bfloat16)
    qs = 1 + 0.1 * rnd(32, seed=3)
    ks = 1 + 0.1 * rnd(32, seed=4)
    table = rnd(heads, (2 * H - 1) * (2 * W - 1), seed=5, scale=0.5) if mode == 0 else None
    bias = None
    if mode == 0:
        ii = torch.arange(H * W, device=dev())
        hi, wi = ii // W, ii % W
        idx = (hi[:, None] - hi[None, :] + H - 1) * (2 * W - 1) + (wi[:, None] - wi[None, :] + W - 1)
        bias = table[:, idx]                                # [heads, n, n]
    o = torch.empty(R, inner, device=dev(), dtype=torch.bfloat16)
    lse = torch.empty(R, heads, device=dev())
    lib.call("ctc_attention_fwd", q, inner, kv, kv.data_ptr() + inner * 2, 2 * inner, B, T, H, W, heads, qs, ks, 8.0,
             table, mode, o, lse, lib.stream_ptr())
    qf = q.float().requires_grad_()
    kvf = kv.float().requires_grad_()
    o_ref, lse_ref, p_ref = _attn_ref(qf, kvf, qs, ks, 8.0, bias, B, T, H, W, heads, mode)
    assert relerr(o, o_ref) < 2e-2
    assert float((lse - lse_ref).abs().max()) < 2e-2
    n = H * W if mode == 0 else T
    n_seq = B * T if mode == 0 else B * H * W
    probs = torch.empty(n_seq, heads, n, n, device=dev())
    lib.call("ctc_attention_probs", q, inner, kv, 2 * inner, lse, B, T, H, W, heads, qs, ks, 8.0, table, mode, probs,
             lib.stream_ptr())
    assert float((probs - p_ref).abs().max()) < 2e-2
    assert float((probs.sum(-1) - 1).abs().max()) < 2e-2
    # backward
    d_o = rnd(R, inner, seed=6, dtype=torch.bfloat16)
    dq_ref, dkv_ref = torch.autograd.grad(o_ref, [qf, kvf], d_o.float())
    dq = torch.empty(R, inner, device=dev(), dtype=torch.bfloat16)
    dkv = torch.empty(R, 2 * inner, device=dev(), dtype=torch.bfloat16)
    delta = torch.empty(R, heads, device=dev())
    # use the reference-precision o for D = rowsum(dO*O) like the product path does (its own bf16 o)
    lib.call("ctc_attention_bwd", q, inner, kv, kv.data_ptr() + inner * 2, 2 * inner, o, d_o, lse, B, T, H, W, heads,
             qs, ks, 8.0, table, mode, dq, inner, dkv, dkv.data_ptr() + inner * 2, 2 * inner, delta, lib.stream_ptr())
    torch.cuda.synchronize()
    assert relerr(dq, dq_ref) < 4e-2
    assert relerr(dkv[:, :inner], dkv_ref[:, :inner]) < 4e-2
    assert relerr(dkv[:, inner:], dkv_ref[:, inner:]) < 4e-2
    return dq, dkv


@pytest.mark.parametrize("B,T,H,W,heads", [(1, 2, 24, 24, 8), (1, 3, 8, 8, 2), (2, 1, 16, 8, 1), (1, 1, 16, 16, 2)])
def test_attention_fwd_tcgen05_matches_reference_and_mma_sync(lib, B, T, H, W, heads):
    """Spatial attention forward on tcgen05/TMEM with the fixed-shift softmax (attention.py:144-180) against the fp32
    torch restatement and against the mma.sync kernel (same bf16 operands: outputs agree to bf16 rounding)."""
    R, inner = B * T * H * W, heads * 32
    q = rnd(R, inner, seed=1, dtype=torch.bfloat16)
    kv = rnd(R, 2 * inner, seed=2, dtype=torch.bfloat16)
    qs = 1 + 0.1 * rnd(32, seed=3)
    ks = 1 + 0.1 * rnd(32, seed=4)
    table = rnd(heads, (2 * H - 1) * (2 * W - 1), seed=5, scale=0.5)
    ii = torch.arange(H * W, device=dev())
    hi, wi = ii // W, ii % W
    idx = (hi[:, None] - hi[None, :] + H - 1) * (2 * W - 1) + (wi[:, None] - wi[None, :] + W - 1)
    bias = table[:, idx]
    bound_t = torch.empty(1, device=dev())
    lib.call("ctc_attention_score_bound", qs, ks, 8.0, table, heads, H, W, bound_t, lib.stream_ptr())
    bound = float(bound_t)
    assert abs(bound - float(8.0 * qs.abs().max() * ks.abs().max() + table.abs().max())) < 1e-4
    o = torch.full((R, inner), float("nan"), device=dev(), dtype=torch.bfloat16)
    lse = torch.full((R, heads), float("nan"), device=dev())
    lib.call("ctc_attention_fwd_tc", q, inner, kv, kv.data_ptr() + inner * 2, 2 * inner, B, T, H, W, heads, qs, ks, 8.0,
             table, bound, o, lse, lib.stream_ptr())
    torch.cuda.synchronize()
    o_ref, lse_ref, _ = _attn_ref(q.float(), kv.float(), qs, ks, 8.0, bias, B, T, H, W, heads, 0)
    assert torch.isfinite(o.float()).all() and torch.isfinite(lse).all()
    assert relerr(o, o_ref) < 2e-2
    assert float((lse - lse_ref).abs().max()) < 2e-2
    o2 = torch.empty_like(o)
    lse2 = torch.empty_like(lse)
    lib.call("ctc_attention_fwd", q, inner, kv, kv.data_ptr() + inner * 2, 2 * inner, B, T, H, W, heads, qs, ks, 8.0,
             table, 0, o2, lse2, lib.stream_ptr())
    assert relerr(o, o2.float()) < 1e-2
    assert float((lse - lse2).abs().max()) < 1e-3


# --------------------------------------------------------------------------------------- GEGLU
def _group(x_part, gate_part):
    """[R,F],[R,F] -> grouped [R, 2F]: 64-wide groups [32 value | 32 gate]"""
    R, Fp = x_part.shape
    return torch.stack([x_part.view(R, Fp // 32, 32), gate_part.view(R, Fp // 32, 32)], dim=2).reshape(R, 2 * Fp)


def test_geglu_fwd_bwd(lib):
    R, Fp = 333, 256
    xp, gp = rnd(R, Fp, seed=1, dtype=torch.bfloat16), rnd(R, Fp, seed=11, dtype=torch.bfloat16)
    u = _group(xp, gp).contiguous()
    h = torch.empty(R, Fp, device=dev(), dtype=torch.bfloat16)
    lib.call("ctc_geglu_fwd", u, R, Fp, h, lib.stream_ptr())
    xf, gf = xp.float().requires_grad_(), gp.float().requires_grad_()
    ref = F.gelu(gf) * xf
    assert relerr(h, ref) < 1e-2
    dh = rnd(R, Fp, seed=2, dtype=torch.bfloat16)
    dx_ref, dg_ref = torch.autograd.grad(ref, [xf, gf], dh.float())
    du = torch.empty(R, 2 * Fp, device=dev(), dtype=torch.bfloat16)
    lib.call("ctc_geglu_bwd", u, dh, R, Fp, du, lib.stream_ptr())
    assert relerr(du, _group(dx_ref, dg_ref)) < 1e-2


@pytest.mark.parametrize("M,Fp,K", [(300, 256, 64), (13824, 1408, 512), (1000, 128, 512)])
def test_gemm_fused_geglu_epilogues(lib, M, Fp, K):
    """Linear + GEGLU (attention.py:47-48) and its adjoint fused into the tcgen05 GEMM epilogues."""
    a = rnd(M, K, seed=1, dtype=torch.bfloat16)
    wv = rnd(Fp, K, seed=2, scale=1 / math.sqrt(K), dtype=torch.bfloat16)
    wg = rnd(Fp, K, seed=3, scale=1 / math.sqrt(K), dtype=torch.bfloat16)
    w = torch.stack([wv.view(Fp // 32, 32, K), wg.view(Fp // 32, 32, K)], dim=1).reshape(2 * Fp, K).contiguous()
    h = torch.full((M, Fp), float("nan"), device=dev(), dtype=torch.bfloat16)
    u = torch.full((M, 2 * Fp), float("nan"), device=dev(), dtype=torch.bfloat16)
    lib.call("ctc_gemm_bf16", a, K, w, K, h, Fp, M, 2 * Fp, K, lib.EPI_GEGLU, None, None, 0, u, 2 * Fp, 0, lib.stream_ptr())
    xv, xg = a.float() @ wv.float().t(), a.float() @ wg.float().t()
    # aux = the adjoint factors [a | b]: a = gelu(gate) = dh/dvalue, b = value * gelu'(gate) = dh/dgate
    xg_ = xg.clone().requires_grad_()
    (dgelu,) = torch.autograd.grad(F.gelu(xg_).sum(), xg_)
    assert relerr(u, _group(F.gelu(xg), xv * dgelu)) < 1e-2
    assert relerr(h, F.gelu(xg) * xv) < 1e-2
    for impl in (2, 3):                                        # single CTAs / CTA pairs: same accumulation order, same bits
        h1, u1 = torch.empty_like(h), torch.empty_like(u)
        lib.call("ctc_gemm_bf16", a, K, w, K, h1, Fp, M, 2 * Fp, K, lib.EPI_GEGLU, None, None, 0, u1, 2 * Fp, impl, lib.stream_ptr())
        assert torch.equal(h, h1) and torch.equal(u, u1)
    h2 = torch.empty_like(h)                                   # without the pre-activation output
    lib.call("ctc_gemm_bf16", a, K, w, K, h2, Fp, M, 2 * Fp, K, lib.EPI_GEGLU, None, None, 0, None, 0, 0, lib.stream_ptr())
    assert torch.equal(h, h2)
    # backward: dh = d @ W2t, du = GEGLU'(u) * dh
    Kd = 256
    d = rnd(M, Kd, seed=4, dtype=torch.bfloat16)
    w2t = rnd(Fp, Kd, seed=5, scale=1 / math.sqrt(Kd), dtype=torch.bfloat16)
    du = torch.full((M, 2 * Fp), float("nan"), device=dev(), dtype=torch.bfloat16)
    lib.call("ctc_gemm_bf16", d, Kd, w2t, Kd, du, 2 * Fp, M, Fp, Kd, lib.EPI_GEGLU_BWD, None, None, 0, u, 2 * Fp, 0,
             lib.stream_ptr())
    for impl in (2, 3):
        du1 = torch.full_like(du, float("nan"))
        lib.call("ctc_gemm_bf16", d, Kd, w2t, Kd, du1, 2 * Fp, M, Fp, Kd, lib.EPI_GEGLU_BWD, None, None, 0, u, 2 * Fp, impl,
                 lib.stream_ptr())
        assert torch.equal(du, du1)
    dh = d.float() @ w2t.float().t()
    xs, gs = xv.clone().requires_grad_(), xg.clone().requires_grad_()
    dx_ref, dg_ref = torch.autograd.grad(F.gelu(gs) * xs, [xs, gs], dh)
    assert relerr(du, _group(dx_ref, dg_ref)) < 1.5e-2


# --------------------------------------------------------------------------------------- patchify
@pytest.mark.parametrize("cfg,B,shared", [(O.TINY, 3, False), (O.TINY, 3, True), (O.FULL, 1, False)])
def test_patchify_ln_fwd_bwd(lib, cfg, B, shared):
    D, Hh = cfg.depth_voxels, cfg.image_size
    pt, p, P = cfg.temporal_patch_size, cfg.patch_size, cfg.patch_dim
    vol = O.synthetic_volume(cfg, 0, batch=1 if shared else B).to(dev())
    g = 1 + 0.1 * rnd(P, seed=1)
    b = 0.05 * rnd(P, seed=2)
    alpha = torch.linspace(0.2, 1.0, B, device=dev())
    occl = torch.tensor([[0, 0, 0, 0, 0, 0]] * B, dtype=torch.int32)
    occl[B - 1] = torch.tensor([pt, p, 2 * p, 2 * pt, 2 * p, 2 * p])
    occl = occl.to(dev())
    R = B * cfg.n_tokens
    # --- occlusion fused on load
    out = torch.empty(R, P, device=dev(), dtype=torch.bfloat16)
    lib.call("ctc_patchify_ln_fwd", vol, 0 if shared else D * Hh * Hh, B, D, Hh, Hh, pt, p, g, b, 1e-5, None, occl,
             -1.0, out, lib.stream_ptr())
    vb = vol.expand(B, -1, -1, -1, -1).clone() if shared else vol.clone()
    o = occl[B - 1].tolist()
    vb[B - 1, :, o[0]:o[0] + o[3], o[1]:o[1] + o[4], o[2]:o[2] + o[5]] = -1
    ref = F.layer_norm(O.patchify(vb, cfg), (P,), g, b).reshape(R, P)
    assert float((out.float() - ref).abs().max()) < 3e-2
    # --- IG interpolation fused on load + backward
    lib.call("ctc_patchify_ln_fwd", vol, 0 if shared else D * Hh * Hh, B, D, Hh, Hh, pt, p, g, b, 1e-5, alpha, None,
             -1.0, out, lib.stream_ptr())
    vb = (vol.expand(B, -1, -1, -1, -1) if shared else vol)
    xa = (1 + alpha.view(B, 1, 1, 1, 1) * (vb - 1)).detach().requires_grad_()
    ref = F.layer_norm(O.patchify(xa, cfg), (P,), g, b).reshape(R, P)
    assert float((out.float() - ref).abs().max()) < 3e-2
    dy = rnd(R, P, seed=3, dtype=torch.bfloat16)
    (gx,) = torch.autograd.grad(ref, xa, dy.float())
    grad = torch.empty(B, 1, D, Hh, Hh, device=dev())
    lib.call("ctc_patchify_ln_bwd", vol, 0 if shared else D * Hh * Hh, B, D, Hh, Hh, pt, p, g, 1e-5, alpha, dy, grad,
             0, 1.0, lib.stream_ptr())
    # rstd is huge on constant patches; compare where the reference gradient is well conditioned
    assert relerr(grad, gx) < 1e-3
    gsum = torch.zeros(D, Hh, Hh, device=dev())
    lib.call("ctc_patchify_ln_bwd", vol, 0 if shared else D * Hh * Hh, B, D, Hh, Hh, pt, p, g, 1e-5, alpha, dy, gsum,
             1, 0.5, lib.stream_ptr())
    assert relerr(gsum, 0.5 * gx.sum(dim=(0, 1))) < 1e-3


# --------------------------------------------------------------------------------------- CPB / VQ / latent
@pytest.mark.parametrize("cfg", [O.TINY, O.FULL])
def test_cpb_table(lib, cfg):
    sd = O.to_device(O.init_state_dict(cfg, 42), dev())
    p = "visual_transformer.spatial_rel_pos_bias.net."
    H = cfg.h
    table = torch.empty(cfg.heads, (2 * H - 1) ** 2, device=dev())
    lib.call("ctc_cpb_table", sd[p + "0.0.weight"], sd[p + "0.0.bias"], sd[p + "1.0.weight"], sd[p + "1.0.bias"],
             sd[p + "2.weight"], sd[p + "2.bias"], cfg.dim, cfg.heads, H, H, table, lib.stream_ptr())
    ref = O.cpb_bias(sd, "visual_transformer.spatial_rel_pos_bias.", H, H)      # [heads, n, n]
    ii = torch.arange(H * H, device=dev())
    hi, wi = ii // H, ii % H
    idx = (hi[:, None] - hi[None, :] + H - 1) * (2 * H - 1) + (wi[:, None] - wi[None, :] + H - 1)
    assert float((table[:, idx] - ref).abs().max()) < 1e-4


def test_vq_argmax_gather_bwd(lib):
    R, C, K = 2000, 512, 8192
    B, T, HW = 2, 10, 100
    x = rnd(R, C, seed=1)
    cb = F.normalize(rnd(K, C, seed=2), dim=-1)
    n_cand = lib.vq_num_candidates(K)
    cv = torch.empty(R, n_cand, device=dev())
    ci = torch.empty(R, n_cand, device=dev(), dtype=torch.int32)
    ind = torch.empty(R, device=dev(), dtype=torch.int32)
    lib.call("ctc_vq_argmax", x, x.to(torch.bfloat16), R, C, cb, cb.to(torch.bfloat16), K, cv, ci, ind, lib.stream_ptr())
    scores = (F.normalize(x, dim=-1).double() @ cb.double().t())
    ref = scores.argmax(dim=-1)
    agree = (ind.long() == ref)
    # disagreements may only be fp32-level ties
    if not agree.all():
        bad = (~agree).nonzero().flatten()
        gap = scores[bad, ref[bad]] - scores[bad, ind.long()[bad]]
        assert float(gap.max()) < 1e-6
    pooled = torch.empty(B, HW * C, device=dev())
    tokens = torch.empty(R, C, device=dev())
    lib.call("ctc_vq_gather_pool", ind, cb, B, T, HW, C, pooled, None, tokens, lib.stream_ptr())
    tref = cb[ind.long()]
    assert torch.equal(tokens, tref)
    assert relerr(pooled.view(B, HW, C), tref.view(B, T, HW, C).mean(dim=1)) < 1e-5
    # straight-through adjoint
    dpooled = rnd(B, HW * C, seed=3)
    xr = x.clone().requires_grad_()
    out, _ = O.vq_cosine(xr.view(B, T * HW, C), cb[None], "ste_l2norm")
    loss = (out.view(B, T, HW, C).mean(dim=1).reshape(B, -1) * dpooled).sum()
    (dx_ref,) = torch.autograd.grad(loss, xr)
    dx = torch.empty(R, C, device=dev())
    lib.call("ctc_vq_bwd", dpooled, None, x, B, T, HW, C, 0, dx, lib.stream_ptr())
    assert relerr(dx, dx_ref) < 1e-4


@pytest.mark.parametrize("B,L,NL,Bt", [(1, 294912, 512, 1), (11, 2304, 32, 3), (20, 8192, 512, 2), (32, 4160, 128, 1),
                                       (37, 2048, 576, 1)])
def test_latent_proj_sim(lib, B, L, NL, Bt):
    pooled = rnd(B, L, seed=1)
    wv = rnd(NL, L, seed=2, scale=1 / math.sqrt(L), dtype=torch.bfloat16)
    n_chunks = (L + 1023) // 1024
    partial = torch.empty(n_chunks, B, NL, device=dev())
    lat = torch.empty(B, NL, device=dev())
    lib.call("ctc_latent_proj", pooled, wv, None, B, L, NL, partial, n_chunks, lat, lib.stream_ptr())
    ref = (pooled.double() @ wv.double().t()).float()
    assert relerr(lat, ref) < 1e-4
    # hi + lo split of an fp32 weight: the product tracks the fp32 weight to ~2^-16
    w32 = wv.float() * (1 + 0.003 * rnd(NL, L, seed=9))
    hi = w32.to(torch.bfloat16)
    lo = (w32 - hi.float()).to(torch.bfloat16)
    lat2 = torch.empty_like(lat)
    lib.call("ctc_latent_proj", pooled, hi, lo, B, L, NL, partial, n_chunks, lat2, lib.stream_ptr())
    ref32 = (pooled.double() @ w32.double().t()).float()
    assert relerr(lat2, ref32) < 3e-5
    lib.call("ctc_latent_proj", pooled, hi, None, B, L, NL, partial, n_chunks, lat2, lib.stream_ptr())
    assert relerr(lat2, ref32) > 1e-4                           # bf16-only weight: visibly coarser
    e = rnd(Bt, 48, seed=3)
    wt = rnd(NL, 48, seed=4)
    tl = torch.empty(Bt, NL, device=dev())
    lib.call("ctc_text_latent", e, wt, Bt, 48, NL, tl, lib.stream_ptr())
    tl_ref = F.normalize(e @ wt.t(), dim=-1)
    assert relerr(tl, tl_ref) < 1e-5
    sim = torch.empty(B, Bt, device=dev())
    il = torch.empty(B, NL, device=dev())
    dl = torch.empty(B, NL, device=dev())
    lib.call("ctc_latent_sim", lat, tl, B, Bt, NL, 2.5, sim, il, dl, lib.stream_ptr())
    lr = ref.clone().requires_grad_()
    iln = lr / lr.norm(dim=-1, keepdim=True)
    sref = iln @ tl_ref.t() * 2.5
    assert relerr(sim, sref) < 1e-4
    assert relerr(il, iln) < 1e-4
    diag = sum(sref[i, i % Bt] for i in range(B))
    (dl_ref,) = torch.autograd.grad(diag, lr)
    assert relerr(dl, dl_ref) < 1e-3
    dp = torch.empty(B, L, device=dev())
    lib.call("ctc_latent_proj_bwd", dl, wv, B, L, NL, dp, lib.stream_ptr())
    assert relerr(dp, (dl.double() @ wv.double()).float()) < 1e-4


# --------------------------------------------------------------------------------------- attribution reductions
def test_rollout_colmean_gradcam(lib):
    heads, n, S = 8, 576, 3
    probs = torch.softmax(rnd(S, heads, n, n, seed=1, scale=2.0), dim=-1).contiguous()
    out = torch.empty(S, n, device=dev())
    lib.call("ctc_rollout_spatial", probs, S, heads, n, out, lib.stream_ptr())
    ref = torch.stack([O.attention_rollout([probs[s]]).sum(dim=0) for s in range(S)])
    assert relerr(out, ref) < 1e-4
    cm = torch.empty(S, heads, n, device=dev())
    lib.call("ctc_attn_colmean", probs, S, heads, n, cm, lib.stream_ptr())
    assert relerr(cm, probs.mean(dim=2)) < 1e-5
    # temporal chain
    Lyr, ntok, T = 4, 50, 24
    pt = torch.softmax(rnd(Lyr, ntok, heads, T, T, seed=2, scale=2.0), dim=-1).contiguous()
    outt = torch.empty(ntok, T, device=dev())
    lib.call("ctc_rollout_temporal", pt, Lyr, ntok, heads, T, outt, lib.stream_ptr())
    reft = torch.stack([O.attention_rollout([pt[l, k] for l in range(Lyr)]).sum(dim=0) for k in range(ntok)])
    assert relerr(outt, reft) < 1e-4
    # Grad-CAM
    R, C = 13824, 512
    g, fa, fb = rnd(R, C, seed=3), rnd(R, C, seed=4), rnd(R, C, seed=5)
    w = torch.empty(C, device=dev())
    ws = torch.empty(lib.load().ctc_colmean_ws_floats(R, C), device=dev())
    lib.call("ctc_colmean", g, R, C, w, ws, lib.stream_ptr())
    assert relerr(w, g.mean(dim=0)) < 1e-3
    cam = torch.empty(R, device=dev())
    lib.call("ctc_gradcam", fa, fb, w, R, C, cam, lib.stream_ptr())
    assert relerr(cam, ((fa - fb) * g.mean(dim=0)).sum(-1).relu()) < 1e-3


@pytest.mark.parametrize("rot", [0, 1])
def test_upsample_trilinear(lib, rot):
    x = rnd(24, 24, 24, seed=1)
    D, H, W = 240, 480, 480
    out = torch.empty((D, W, H) if rot else (D, H, W), device=dev())
    lib.call("ctc_upsample_trilinear", x, 24, 24, 24, out, D, H, W, rot, lib.stream_ptr())
    ref = O.upsample(x, (D, H, W))
    if rot:
        ref = O.rot90(ref).copy()
    assert float(np.abs(out.cpu().numpy() - ref).max()) < 1e-5


def test_ig_combine(lib):
    n = 240 * 480 * 48
    x = rnd(n, seed=1).clamp(-1, 1)
    gs = rnd(n, seed=2)
    ig = torch.empty(n, device=dev())
    mm = torch.tensor([float("inf"), float("-inf")], device=dev())
    lib.call("ctc_ig_combine", x, gs, n, 1 / 50, ig, mm, lib.stream_ptr())
    ref = ((x - 1) * (gs / 50)).relu()
    assert relerr(ig, ref) < 1e-6
    assert float(mm[0]) == float(ref.min()) and abs(float(mm[1]) - float(ref.max())) < 1e-6
