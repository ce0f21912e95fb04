"""Micro-benchmarks of single kernels at the batch-8 shapes of the step (CUDA events, L2 flushed between
iterations by cycling through more buffers than fit the 126 MB L2).  Development aid.
    python tools/kernel_bench.py [gemm|attn_t|attn_s|all]"""
import sys, os
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "ct-clip-ut_b200"))
import torch
from ctclip_b200 import _lib as L

dev = torch.device("cuda")
what = sys.argv[1] if len(sys.argv) > 1 else "all"
B, T, H, W, heads = 8, 24, 24, 24, 8
R = B * T * H * W
bf = torch.bfloat16


def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3   # us


def rnd(*shape, dtype=torch.float32, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).to(dtype)


if what in ("gemm", "all"):
    for name, N, K, epi, resid in [("q", 256, 512, L.EPI_BF16, False), ("kv", 512, 512, L.EPI_BF16, False),
                                   ("out+res", 512, 256, L.EPI_F32, True), ("ff1 geglu (no u)", 2816, 512, L.EPI_GEGLU, False), ("ff1 geglu + factors", 2816, 512, L.EPI_GEGLU, "u"),
                                   ("ff2+res", 512, 1408, L.EPI_F32, True), ("dh", 1408, 512, L.EPI_BF16, False), ("dh + geglu adjoint", 1408, 512, L.EPI_GEGLU_BWD, False),
                                   ("dxn2", 512, 2816, L.EPI_F32, False), ("pe", 512, 4000, L.EPI_F32, False),
                                   ("pe bwd", 4000, 512, L.EPI_BF16, False)]:
        a = rnd(R, K, dtype=bf)
        w = rnd(N, K, dtype=bf, scale=K ** -0.5)
        aux = None
        if epi == L.EPI_GEGLU_BWD:
            out = torch.empty(R, 2 * N, device=dev, dtype=bf)
            aux = rnd(R, 2 * N, dtype=bf)
        elif epi == L.EPI_GEGLU:
            out = torch.empty(R, N // 2, device=dev, dtype=bf)
            if resid == "u":
                aux, resid = torch.empty(R, N, device=dev, dtype=bf), False
        else:
            out = torch.empty(R, N, device=dev, dtype=bf if epi == L.EPI_BF16 else torch.float32)
        res = rnd(R, N) if resid else None
        def f(impl=0):
            L.call("ctc_gemm_bf16", a, K, w, K, out, out.stride(0), R, N, K, epi, None, res, N if resid else 0, aux,
                   aux.stride(0) if aux is not None else 0, impl, L.stream_ptr())
        us = timeit(lambda: f(3))            # CTA pairs (cta_group::2)
        us1 = timeit(lambda: f(2))           # single-CTA kernel (cta_group::1)
        # direct (staging-free) epilogues: timing only - the weight rows are not permuted here, the work is identical
        flag = L.GEMM_BPERM
        usd, usd1 = timeit(lambda: f(3 | flag)), timeit(lambda: f(2 | flag))
        usb = timeit(lambda: torch.matmul(a, w.t()))     # cuBLAS, bare bf16 GEMM (no epilogue work)
        fl = 2.0 * R * N * K / 1e6
        print(f"  gemm {name:20s} N={N:5d} K={K:5d}: staged pair {us:6.1f} 1cta {us1:6.1f} | direct pair {usd:6.1f} us {fl / usd:6.0f} TF"
              f"  1cta {usd1:6.1f} us {fl / usd1:6.0f} TF | cuBLAS {usb:6.1f} us {fl / usb:6.0f} TF")

for mode, tag in ((1, "attn_t"), (0, "attn_s")):
    if what not in (tag, "all"):
        continue
    inner = heads * 32
    q, kv = rnd(R, inner, dtype=bf), rnd(R, 2 * inner, dtype=bf)
    qs, ks = torch.ones(32, device=dev), torch.ones(32, device=dev)
    table = rnd(heads, 47 * 47, scale=0.5) if mode == 0 else None
    o = torch.empty(R, inner, device=dev, dtype=bf)
    lse = torch.empty(R, heads, device=dev)
    d_o = rnd(R, inner, dtype=bf)
    dq, dkv, delta = torch.empty_like(q), torch.empty_like(kv), torch.empty(R, heads, device=dev)
    def fwd():
        L.call("ctc_attention_fwd", q, inner, kv, kv.data_ptr() + inner * 2, 2 * inner, B, T, H, W, heads, qs, ks, 8.0,
               table, mode, o, lse, L.stream_ptr())
    def bwd():
        L.call("ctc_attention_bwd", q, inner, kv, kv.data_ptr() + inner * 2, 2 * inner, o, d_o, lse, B, T, H, W, heads,
               qs, ks, 8.0, table, mode, dq, inner, dkv, dkv.data_ptr() + inner * 2, 2 * inner, delta, L.stream_ptr())
    prev_mode = L.load().ctc_attention_set_tc_bwd(0)
    print(f"  {tag} fwd {timeit(fwd):8.1f} us   bwd (dq + dkv, mma.sync) {timeit(bwd):8.1f} us")
    L.load().ctc_attention_set_tc_bwd(prev_mode)
    if mode == 0:
        lib = L.load()
        lib.ctc_attention_set_tc_bwd(2)
        t1 = timeit(bwd)
        ref = (dq.clone(), dkv.clone())
        lib.ctc_attention_set_tc_bwd(0)
        bwd(); torch.cuda.synchronize()
        lib.ctc_attention_set_tc_bwd(prev_mode)
        err = max(float((dq.float() - ref[0].float()).abs().max() / dq.float().abs().max()),
                  float((dkv.float() - ref[1].float()).abs().max() / dkv.float().abs().max()))
        print(f"  {tag} bwd one pass on tcgen05 (dq, dk, dv) {t1:8.1f} us   (max rel diff to the mma.sync kernels {err:.2e})")
    if mode == 0:
        bt = torch.empty(1, device=dev)
        L.call("ctc_attention_score_bound", qs, ks, 8.0, table, heads, H, W, bt, L.stream_ptr())
        bound = float(bt)
        def fwd_tc():
            L.call("ctc_attention_fwd_tc", q, inner, kv, kv.data_ptr() + inner * 2, 2 * inner, B, T, H, W, heads, qs, ks,
                   8.0, table, bound, o, lse, L.stream_ptr())
        lib = L.load()
        lib.ctc_attention_set_exp2_poly(0)
        t_mufu = timeit(fwd_tc)
        lib.ctc_attention_set_exp2_poly(1)
        t_poly = timeit(fwd_tc)
        print(f"  {tag} fwd tcgen05: all exponentials on MUFU {t_mufu:8.1f} us | half on FMA-pipe polynomials {t_poly:8.1f} us "
              f"(score bound {bound:.2f})")

if what in ("vq", "all"):
    C, K = 512, 8192
    x = rnd(R, C)
    xb = x.to(bf)
    cb = torch.nn.functional.normalize(rnd(K, C), dim=-1)
    cbb = cb.to(bf)
    nc = L.vq_num_candidates(K)
    cv, ci = torch.empty(R, nc, device=dev), torch.empty(R, nc, device=dev, dtype=torch.int32)
    ind = torch.empty(R, device=dev, dtype=torch.int32)
    t = timeit(lambda: L.call("ctc_vq_argmax", x, xb, R, C, cb, cbb, K, cv, ci, ind, L.stream_ptr()))
    print(f"  vq_argmax (score GEMM N=8192 K=512 + refine; CTC_GEMM_ARGMAX_PAIR={os.environ.get('CTC_GEMM_ARGMAX_PAIR', '0')}): {t:8.1f} us")
