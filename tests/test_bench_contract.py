"""CPU: the driver-facing contract of bench.py that can be checked without a GPU — the reference arm
(`--impl reference`: the oracle port on the host cores) prints ONE JSON line with the agreed keys, and the
FLOP bookkeeping of the attribution sub-metric matches SURVEY §8d."""
import importlib.util
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _bench_module():
    spec = importlib.util.spec_from_file_location("bench_mod", ROOT / "bench.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ctvit_fwd_bwd_volumes_per_sec" and d["unit"] == "volumes/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["config"]["workload"] == "ctvit_fwd_bwd_b8_480x480x240"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_attribution_flop_bookkeeping():
    b = _bench_module()
    dense, executed, n = b.occlusion_flops()
    assert n == 12167                                            # 23^3 windows of the reference sweep
    assert abs(dense - 12168 * b.FLOP_FWD) < 1e6                 # + the un-occluded forward
    # dense-equivalent work of one attributed volume (9.683 PFLOP) and the executed share
    total = dense + 50 * (b.FLOP_FWD + b.FLOP_BWD)
    assert abs(total / 1e15 - 9.683) < 5e-3
    assert 0.60 < executed / dense < 0.65


def test_synthetic_scans_prefix_and_file_order():
    """bench.synthetic_scans: (1) the first scan of a batch depends on the seed only - every rank rebuilds rank 0's first
    volume for the sharded latency mode / the N-rank parity check from it (a per-rank volume made that check fail in round
    2); (2) the fp32 volume is clamp(hu) / 1000 of the int16 scan, which is stored in NIfTI file order (H fastest)."""
    import importlib.util
    from pathlib import Path
    import torch
    spec = importlib.util.spec_from_file_location("bench_mod", Path(__file__).resolve().parent.parent / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    shape, border = (16, 48, 32), (2, 4, 4)
    v8, r8 = bench.synthetic_scans(1234, 8, shape, border)
    v1, r1 = bench.synthetic_scans(1234, 1, shape, border)
    assert torch.equal(v8[:1], v1) and torch.equal(r8[:1], r1)
    assert not torch.equal(bench.synthetic_scans(1235, 1, shape, border)[0], v1)
    assert v8.shape == (8, 1, 16, 48, 32) and r8.shape == (8, 16, 32, 48) and r8.dtype == torch.int16
    hu = r8.permute(0, 1, 3, 2)                                   # logical [B, D, H, W]
    assert torch.equal(v8[:, 0], hu.float() / 1000)
    assert float(v8.min()) == -1.0 and float(v8.max()) <= 1.0 and torch.all(v8[:, :, :2] == -1)
