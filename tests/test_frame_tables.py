"""CPU: the index tables of the occlusion fast path (Engine.forward_occluded) against a brute-force
dependency simulation of the spatial transformer's only cross-frame operator, the causal PEG stencil
(reference: src/utils/attention.py:55-83 with ctvit.py:60-61 — output frame t reads frames t-2..t)."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from ctclip_b200.engine import INT_MIN, occlusion_frame_tables


def brute_force_changed_sets(t0, nt, T, n_layers):
    changed = [set(range(t0, t0 + nt))]
    for _ in range(n_layers):
        prev = changed[-1]
        changed.append({t for t in range(T) if any((t - k) in prev for k in range(3))})
    return changed


@settings(max_examples=60, deadline=None)
@given(st.integers(3, 24), st.integers(1, 4), st.integers(1, 3), st.integers(1, 5), st.data())
def test_tables_follow_the_causal_stencil(T, n_layers, nt, n_win, data):
    H = 6
    nt = min(nt, T)
    nh, nw = 2, 1
    cubes = [(data.draw(st.integers(0, T - nt)), data.draw(st.integers(0, H - nh)), data.draw(st.integers(0, H - nw)))
             for _ in range(n_win)]
    tables, layer_F = occlusion_frame_tables(cubes, (nt, nh, nw), T, H, n_layers)
    assert len(tables) == n_layers + 3 and len(layer_F) == n_layers
    sets = [brute_force_changed_sets(c[0], nt, T, n_layers) for c in cubes]
    # identity (window, frame) of every compact frame of the layer-0 input
    ident = [(w, cubes[w][0] + j) for w in range(n_win) for j in range(nt)]
    assert list(tables[0]) == [-1 - t for (_, t) in ident]
    rows = set(int(r) for r in tables[1])
    expect_rows = set()
    for f, (w, t) in enumerate(ident):
        for a in range(nh):
            for b in range(nw):
                expect_rows.add(f * H * H + (cubes[w][1] + a) * H + cubes[w][2] + b)
    assert rows == expect_rows and len(tables[1]) == len(expect_rows)
    for l in range(n_layers):
        out_ident = [(w, t) for w in range(n_win) for t in sorted(sets[w][l + 1])]
        assert layer_F[l] == len(out_ident)
        src = tables[2 + l].reshape(-1, 3)
        assert len(src) == len(out_ident)
        prev_index = {wt: i for i, wt in enumerate(ident)}
        for f, (w, t) in enumerate(out_ident):
            for k, dt in enumerate((-2, -1, 0)):
                tt = t + dt
                if tt < 0:
                    assert src[f, k] == INT_MIN
                elif tt in sets[w][l]:
                    assert src[f, k] == prev_index[(w, tt)]
                else:
                    assert src[f, k] == -1 - tt
        ident = out_ident
    full = tables[-1].reshape(n_win, T)
    last_index = {wt: i for i, wt in enumerate(ident)}
    for w in range(n_win):
        for t in range(T):
            assert full[w, t] == (last_index[(w, t)] if t in sets[w][n_layers] else -1 - t)


def test_reference_sweep_executes_62_percent_of_the_dense_frames():
    """(20,40,40)/(10,20,20) on 240x480x480: t0 = 0..22, 2-frame cubes, 4 spatial layers -> changed frames per
    layer 4,6,8,10 clipped at the last frame."""
    total = 0
    for t0 in range(23):
        _, layer_F = occlusion_frame_tables([(t0, 0, 0)], (2, 2, 2), 24, 24, 4)
        assert layer_F == [min(4 + 2 * l, 24 - t0) for l in range(4)]
        total += sum(layer_F)
    assert total == sum(min(4 + 2 * l, 24 - t0) for t0 in range(23) for l in range(4))
    assert total / (23 * 4 * 24) < 0.28
