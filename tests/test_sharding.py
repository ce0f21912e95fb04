"""CPU: host-side sharding logic (window / alpha-step partition) including the world_size-2 gloo path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from hypothesis import given, settings, strategies as st

from oracle import ctclip_oracle as O


def _attr():
    from ctclip_b200 import attribution
    return attribution


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 400), st.integers(1, 9), st.booleans())
def test_shard_range_properties(total, world, parity):
    A = _attr()
    ranges = [A.shard_range(total, r, world, parity) for r in range(world)]
    covered = [i for s, e in ranges for i in range(s, e)]
    assert covered == sorted(set(covered))                        # disjoint, ordered, contiguous per rank
    if parity:   # reference semantics: total // world each, remainder dropped (visualizations.py:351-361)
        assert all(e - s == total // world for s, e in ranges)
        assert len(covered) == (total // world) * world
        windows = list(range(total))
        for r in range(world):
            assert windows[ranges[r][0]:ranges[r][1]] == O.shard_windows(windows, r, world)
    else:
        assert covered == list(range(total))
        assert max(e - s for s, e in ranges) - min(e - s for s, e in ranges) <= 1


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 400), st.integers(1, 9), st.booleans())
def test_shard_indices_cover_the_reference_window_set(total, world, parity):
    """Round-robin dealing: disjoint, balanced to within one window, and the union is exactly the set of windows the
    reference evaluates (its truncated list, visualizations.py:351-361) — who evaluates a window is not observable."""
    A = _attr()
    parts = [A.shard_indices(total, r, world, parity).tolist() for r in range(world)]
    flat = sorted(i for p in parts for i in p)
    assert len(flat) == len(set(flat))
    windows = list(range(total))
    ref = sorted(w for r in range(world) for w in O.shard_windows(windows, r, world)) if parity else windows
    assert flat == ref
    assert max(map(len, parts)) - min(map(len, parts)) <= (0 if parity else 1)
    assert all(p == sorted(p) for p in parts)


@settings(max_examples=25, deadline=None)
@given(st.tuples(st.integers(4, 40), st.integers(4, 40), st.integers(4, 40)),
       st.tuples(st.integers(1, 8), st.integers(1, 8), st.integers(1, 8)),
       st.tuples(st.integers(1, 5), st.integers(1, 5), st.integers(1, 5)))
def test_window_enumeration_matches_oracle(shape, patch, stride):
    A = _attr()
    patch = tuple(min(p, s) for p, s in zip(patch, shape))
    assert A.occlusion_windows(shape, patch, stride) == O.occlusion_windows(shape, patch, stride)


def test_default_window_grid():
    A = _attr()
    w = A.occlusion_windows((240, 480, 480))
    assert len(w) == 12167 and w[:3] == [(0, 0, 0), (0, 0, 20), (0, 0, 40)] and w[-1] == (220, 440, 440)


def _worker(rank, world, port, total, parity, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    A = _attr()
    s, e = A.shard_range(total, rank, world, parity)
    local = torch.arange(s, e, dtype=torch.float32) * 0.5 + 1.0       # stand-in per-window scores
    scores, inc = A.combine_sharded(local, s, e, total)
    # the round-robin assignment occlusion_sensitivity uses must assemble the same vector and mask
    idx = A.shard_indices(total, rank, world, parity)
    scores_rr, inc_rr = A.combine_indexed(torch.from_numpy(idx).float() * 0.5 + 1.0, idx, total)
    assert torch.equal(scores_rr, scores) and torch.equal(inc_rr, inc)
    two = torch.stack([torch.from_numpy(idx).float(), -torch.from_numpy(idx).float()], dim=1)     # [n, P=2] (multi-prompt)
    s2p, _ = A.combine_indexed(two, idx, total)
    assert s2p.shape == (total, 2) and torch.equal(s2p[:, 0], -s2p[:, 1])
    # IG partial sums: each rank owns a slice of steps, the all-reduce restores the full sum
    s2, e2 = A.shard_range(50, rank, world, False)
    part = torch.tensor([float(sum(range(s2, e2)))])
    dist.all_reduce(part)
    q.put((rank, scores.tolist(), inc.tolist(), float(part)))
    dist.destroy_process_group()


@pytest.mark.parametrize("parity", [True, False])
def test_world2_gloo_combine(parity):
    world, total = 2, 27
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, parity, q)) for r in range(world)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in range(world)]
    [p.join(timeout=60) for p in procs]
    kept = (total // world) * world if parity else total
    for rank, scores, inc, part in res:
        assert inc == [1] * kept + [0] * (total - kept)
        assert scores[:kept] == [i * 0.5 + 1.0 for i in range(kept)] and all(v == 0 for v in scores[kept:])
        assert part == float(sum(range(50)))


def _bcast_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from types import SimpleNamespace
    A = _attr()
    acc = SimpleNamespace(is_main_process=rank == 0, process_index=rank, num_processes=world, device=torch.device("cpu"))
    vis = A.Visualizations(torch.nn.Identity(), acc, None, None, 1, "/tmp/unused", "", None)
    sample = None
    if rank == 0:
        g = torch.Generator().manual_seed(3)
        sample = (torch.randn(1, 4, 6, 6, generator=g), "report text", torch.arange(18) % 2, "scan7", "scan7.nii.gz")
    out = vis._broadcast_sample(sample)
    q.put((rank, out[0].shape, float(out[0].double().sum()), out[1], out[2].tolist(), out[3], out[4]))
    dist.destroy_process_group()


def test_world2_gloo_sample_broadcast():
    """Visualizations._broadcast_sample (visualizations.py:296-318): rank 0 loads the sample, every rank receives the
    tensors with a leading batch axis and the strings wrapped in lists, exactly as a DataLoader batch looks."""
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bcast_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    assert res[0][1:] == res[1][1:]
    _, shape, _, text, labels, name, path = res[0]
    assert tuple(shape) == (1, 1, 4, 6, 6) and text == ["report text"] and name == ["scan7"] and path == ["scan7.nii.gz"]
    assert labels == [[i % 2 for i in range(18)]]
