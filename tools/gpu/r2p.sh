#!/bin/bash
# GPU session P: one-pass attention backward with two softmax warp groups on alternating tiles: parity tests + micro-benchmark + step A/B
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "attention" > $O/r2p_ops.log 2>&1; echo "rc=$?" >> $O/r2p_ops.log
tail -4 $O/r2p_ops.log
grep -q "rc=0" $O/r2p_ops.log || exit 1
timeout 200 python tools/kernel_bench.py attn_s > $O/r2p_kbench.log 2>&1; cat $O/r2p_kbench.log
for m in 2 0 2 0; do
CTC_ATTN_BWD=$m timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-attribution 2> $O/r2p_bench.err | cut -c1-230 | sed "s/^/attn_bwd=$m /"
done
