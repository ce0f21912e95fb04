"""Run-time self check of the multi-GPU path: sharding the occlusion windows / IG alpha steps of ONE volume over the
ranks of the default process group must reproduce what a single rank computes — window scores and heat maps bit for
bit (which rank scores a window is not observable), the IG partial sum up to the order of an fp32 addition.
Used by `bench.py --gpus N` (the `parity` field of its JSON line) and `tools/multi_gpu_check.py`."""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import attribution as A
from .engine import Engine


def sharding_parity(eng: Engine, vol: torch.Tensor, tl: torch.Tensor, tl_multi: torch.Tensor = None,
                    ps=(80, 160, 160), st=(80, 160, 160), ig_steps: int = 6) -> dict:
    """Every rank calls this with the SAME volume.  27 windows at the default size: at world 2 / 4 / 8 the reference's
    `total // world` split drops 1 / 3 / 3 of them (visualizations.py:351-361) and so does the parity mode here."""
    dev = vol.device
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    shape = tuple(vol.shape[-3:])
    heat, aux = A.occlusion_sensitivity(eng, vol, tl, ps, st, batch=4, parity_sharding=True)
    ig, iaux = A.integrated_gradients(eng, vol, tl, steps=ig_steps, batch=3)
    # the single-rank computation, on every rank (same deterministic kernels)
    windows = A.occlusion_windows(shape, ps, st)
    orig, scores = A.occlusion_scores(eng, vol, tl, windows, ps, batch=4)
    kept = len(windows) // world * world
    ok = orig == aux["orig"]
    ok &= bool(torch.equal(aux["scores"][:kept], scores[:kept])) and int(aux["included"].sum()) == kept
    inc = torch.zeros(len(windows), dtype=torch.uint8, device=dev)
    inc[:kept] = 1
    heat1 = A.occlusion_heatmap(orig, scores * inc, inc, shape, ps, st)
    ok &= bool(torch.equal(heat, heat1))
    if tl_multi is not None:                 # multi-prompt sweep: one exchange carries the scores of all prompts
        heats, maux = A.occlusion_sensitivity_multi(eng, vol, tl_multi, ps, st, batch=4)
        o3, s3 = A.occlusion_scores(eng, vol, tl_multi, windows, ps, batch=4, all_prompts=True)
        ok &= bool(torch.equal(maux["scores"][:kept], s3[:kept])) and bool(torch.equal(maux["orig"], o3))
        ok &= int(maux["included"].sum()) == kept and len(heats) == tl_multi.shape[0]
    ig1, iaux1 = A.integrated_gradients(eng, vol, tl, steps=ig_steps, batch=3, shard_steps=False)
    rel, map_diff = 0.0, 0.0
    if rank == 0:                            # the reduced sum and the finished map live on rank 0
        rel = float((iaux["gsum"] - iaux1["gsum"]).abs().max() / iaux1["gsum"].abs().max())
        map_diff = float((iaux["pre_threshold"] - iaux1["pre_threshold"]).abs().max())
        ok &= rel < 1e-5 and bool(torch.equal(iaux["scores"], iaux1["scores"]))
    flag = torch.tensor([1 if ok else 0], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return {"pass": bool(int(flag)), "world": world, "windows": len(windows), "windows_kept": kept,
            "occlusion_scores_and_heat_map": "bit-identical to the single-rank sweep" if int(flag) else "MISMATCH",
            "ig_partial_sum_rel_diff": rel, "ig_pre_threshold_map_max_abs_diff": map_diff}
