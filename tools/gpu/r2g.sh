#!/bin/bash
# GPU session G: re-measure the kernel micro-benchmarks (one-pass attention backward v2) and the step with either attention backward
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 300 python tools/kernel_bench.py all > $O/r2g_kbench.log 2>&1
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-attribution > $O/r2g_bench_step.json 2> $O/r2g_bench_step.err; echo "bench rc=$?" >> $O/r2g_bench_step.err
CTC_ATTN_BWD=2 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-attribution > $O/r2g_bench_onepass.json 2> $O/r2g_bench_onepass.err; echo "bench rc=$?" >> $O/r2g_bench_onepass.err
cat $O/r2g_kbench.log; tail -2 $O/r2g_bench_step.err; cut -c1-300 $O/r2g_bench_step.json; tail -2 $O/r2g_bench_onepass.err; cut -c1-300 $O/r2g_bench_onepass.json
