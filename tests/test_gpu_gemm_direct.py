"""GPU: the staging-free ("direct") GEMM epilogues - B rows permuted inside every group of 32 at plan time
(ctc_gemm_row_perm, CTC_GEMM_BPERM) - must reproduce the staged epilogues bit for bit."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from ctclip_b200 import _lib
    _lib.require_device()
    return _lib


def _permute_rows(lib, w):
    """Rows of the B operand permuted inside every group of 32 the way Plan._pack does (ctc_gemm_row_perm)."""
    perm = torch.tensor(lib.gemm_row_perm(), device=w.device, dtype=torch.long)
    return w.view(w.shape[0] // 32, 32, w.shape[1])[:, perm, :].reshape(w.shape).contiguous()


def test_gemm_row_perm_is_a_permutation(lib):
    for kind in (16, 32):
        assert sorted(lib.gemm_row_perm()) == list(range(32))


@pytest.mark.parametrize("impl", [0, 1, 2, 3])     # product selection, SIMT comparator, single CTAs, CTA pairs
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (1000, 512, 512), (4161, 1408, 512), (777, 64, 256), (512, 4000, 512),
                                   (129, 256, 512), (38016, 512, 512), (300, 96, 136)])
def test_gemm_direct_epilogue_bf16_and_f32(lib, impl, M, N, K):
    """The staging-free epilogues (B rows permuted at plan time, CTC_GEMM_BPERM) give the bits of the staged ones."""
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    a = (torch.randn(M, K, generator=g)).to(torch.bfloat16).cuda()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(torch.bfloat16).cuda()
    bias = torch.randn(N, generator=g).cuda()
    resid = torch.randn(M, N, generator=g).cuda()
    ref16 = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    out16 = torch.full_like(ref16, float("nan"))
    lib.call("ctc_gemm_bf16", a, K, w, K, ref16, N, M, N, K, lib.EPI_BF16, None, None, 0, None, 0, impl, lib.stream_ptr())
    lib.call("ctc_gemm_bf16", a, K, _permute_rows(lib, w), K, out16, N, M, N, K, lib.EPI_BF16, None, None, 0, None, 0,
             impl | lib.GEMM_BPERM, lib.stream_ptr())
    ref32, out32 = resid.clone(), resid.clone()       # in-place residual
    lib.call("ctc_gemm_bf16", a, K, w, K, ref32, N, M, N, K, lib.EPI_F32, bias, ref32, N, None, 0, impl, lib.stream_ptr())
    lib.call("ctc_gemm_bf16", a, K, _permute_rows(lib, w), K, out32, N, M, N, K, lib.EPI_F32, bias, out32, N, None, 0,
             impl | lib.GEMM_BPERM, lib.stream_ptr())
    torch.cuda.synchronize()
    assert torch.isfinite(out16.float()).all() and torch.equal(out16, ref16)
    assert torch.equal(out32, ref32)
    assert float((ref16.float() - a.float() @ w.float().t()).abs().max()) < 0.1


@pytest.mark.parametrize("impl", [0, 2, 3])
@pytest.mark.parametrize("M,Fp,K", [(300, 256, 64), (13824, 1408, 512), (1000, 128, 512), (4161, 192, 256)])
def test_gemm_direct_epilogue_geglu(lib, impl, M, Fp, K):
    g = torch.Generator(device="cpu").manual_seed(M + Fp + K)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(2 * Fp, K, generator=g) / math.sqrt(K)).to(torch.bfloat16).cuda()   # 64-row groups [32 value | 32 gate]
    h0, u0 = (torch.full((M, n), float("nan"), device="cuda", dtype=torch.bfloat16) for n in (Fp, 2 * Fp))
    h1, u1, h2 = torch.full_like(h0, float("nan")), torch.full_like(u0, float("nan")), torch.full_like(h0, float("nan"))
    lib.call("ctc_gemm_bf16", a, K, w, K, h0, Fp, M, 2 * Fp, K, lib.EPI_GEGLU, None, None, 0, u0, 2 * Fp, impl, lib.stream_ptr())
    wp = _permute_rows(lib, w)
    lib.call("ctc_gemm_bf16", a, K, wp, K, h1, Fp, M, 2 * Fp, K, lib.EPI_GEGLU, None, None, 0, u1, 2 * Fp,
             impl | lib.GEMM_BPERM, lib.stream_ptr())
    lib.call("ctc_gemm_bf16", a, K, wp, K, h2, Fp, M, 2 * Fp, K, lib.EPI_GEGLU, None, None, 0, None, 0,
             impl | lib.GEMM_BPERM, lib.stream_ptr())
    Kd = 256
    d = torch.randn(M, Kd, generator=g).to(torch.bfloat16).cuda()
    w2t = (torch.randn(Fp, Kd, generator=g) / math.sqrt(Kd)).to(torch.bfloat16).cuda()
    du0, du1 = torch.full_like(u0, float("nan")), torch.full_like(u0, float("nan"))
    lib.call("ctc_gemm_bf16", d, Kd, w2t, Kd, du0, 2 * Fp, M, Fp, Kd, lib.EPI_GEGLU_BWD, None, None, 0, u0, 2 * Fp, impl,
             lib.stream_ptr())
    lib.call("ctc_gemm_bf16", d, Kd, _permute_rows(lib, w2t), Kd, du1, 2 * Fp, M, Fp, Kd, lib.EPI_GEGLU_BWD, None, None, 0,
             u0, 2 * Fp, impl | lib.GEMM_BPERM, lib.stream_ptr())
    torch.cuda.synchronize()
    assert torch.isfinite(h1.float()).all() and torch.isfinite(u1.float()).all() and torch.isfinite(du1.float()).all()
    assert torch.equal(h1, h0) and torch.equal(u1, u0) and torch.equal(h2, h0) and torch.equal(du1, du0)


def test_gemm_direct_epilogue_rejects_mismatched_flags(lib):
    a = torch.zeros(128, 64, device="cuda", dtype=torch.bfloat16)
    w = torch.zeros(48, 64, device="cuda", dtype=torch.bfloat16)
    out = torch.zeros(128, 48, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):      # N not a multiple of 32
        lib.call("ctc_gemm_bf16", a, 64, w, 64, out, 48, 128, 48, 64, lib.EPI_BF16, None, None, 0, None, 0, lib.GEMM_BPERM,
                 lib.stream_ptr())
    w = torch.zeros(64, 64, device="cuda", dtype=torch.bfloat16)
    out = torch.zeros(128, 68, device="cuda", dtype=torch.float32)
    with pytest.raises(RuntimeError):      # fp32 rows that are not 32-byte aligned (ldc = 68)
        lib.call("ctc_gemm_bf16", a, 64, w, 64, out, 68, 128, 64, 64, lib.EPI_F32, None, None, 0, None, 0, lib.GEMM_BPERM,
                 lib.stream_ptr())
