"""One launch of each epilogue-heavy GEMM of the step at the batch-8 shape, inside a cudaProfilerStart/Stop range
(for ncu --profile-from-start off -k regex:gemm_tcgen05).  Development aid.
    python tools/prof_gemm.py [ff1u|adj|outres|ff2res|all] [impl]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "ct-clip-ut_b200"))
import torch
from ctclip_b200 import _lib as L

dev = torch.device("cuda")
what = sys.argv[1] if len(sys.argv) > 1 else "all"
impl = (int(sys.argv[2]) if len(sys.argv) > 2 else 0) | L.GEMM_BPERM   # product path: direct epilogues (timing only: B is not permuted here)
R = 8 * 24 * 24 * 24
bf = torch.bfloat16


def rnd(*shape, dtype=torch.float32, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).to(dtype)


cases = {"ff1u": (2816, 512, L.EPI_GEGLU, False), "adj": (1408, 512, L.EPI_GEGLU_BWD, False),
         "outres": (512, 256, L.EPI_F32, True), "ff2res": (512, 1408, L.EPI_F32, True)}
fns = []
for name, (N, K, epi, resid) in cases.items():
    if what not in (name, "all"):
        continue
    a = rnd(R, K, dtype=bf)
    w = rnd(N, K, dtype=bf, scale=K ** -0.5)
    aux = None
    if epi == L.EPI_GEGLU_BWD:
        out = torch.empty(R, 2 * N, device=dev, dtype=bf); aux = rnd(R, 2 * N, dtype=bf)
    elif epi == L.EPI_GEGLU:
        out = torch.empty(R, N // 2, device=dev, dtype=bf); aux = torch.empty(R, N, device=dev, dtype=bf)
    else:
        out = torch.empty(R, N, device=dev, dtype=torch.float32)
    res = rnd(R, N) if resid else None
    def f(a=a, w=w, out=out, res=res, aux=aux, N=N, K=K, epi=epi, resid=resid):
        L.call("ctc_gemm_bf16", a, K, w, K, out, out.stride(0), R, N, K, epi, None, res, N if resid else 0, aux,
               aux.stride(0) if aux is not None else 0, impl, L.stream_ptr())
    fns.append((name, f))
for _, f in fns:
    for _ in range(3): f()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for name, f in fns:
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); f(); e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) * 1e3:.1f} us")
torch.cuda.cudart().cudaProfilerStop()
