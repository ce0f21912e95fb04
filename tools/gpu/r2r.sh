#!/bin/bash
# GPU session R: converged issue warps everywhere + one-pass attention backward as the default: full GPU suite, kernel bench, step A/B
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2r_tests.log 2>&1; echo "rc=$?" >> $O/r2r_tests.log
tail -4 $O/r2r_tests.log
timeout 300 python tools/kernel_bench.py all > $O/r2r_kbench.log 2>&1; cat $O/r2r_kbench.log
for m in 2 0 2; do
CTC_ATTN_BWD=$m timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-attribution 2> $O/r2r_bench.err | cut -c1-230 | sed "s/^/attn_bwd=$m /"
done
