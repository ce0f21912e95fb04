#!/bin/bash
# GPU session E: one-pass attention backward v2 + faster GEGLU epilogue: kernel tests, micro-benchmarks, parity, bench
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "gemm or attention or latent" > $O/r2e_ops.log 2>&1; echo "rc=$?" >> $O/r2e_ops.log
timeout 300 python tools/kernel_bench.py all > $O/r2e_kbench.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -s > $O/r2e_parity.log 2>&1; echo "parity rc=$?" >> $O/r2e_parity.log
timeout 700 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r2e_bench.json 2> $O/r2e_bench.err; echo "bench rc=$?" >> $O/r2e_bench.err
CTC_ATTN_BWD=2 timeout 700 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-attribution > $O/r2e_bench_onepass.json 2> $O/r2e_bench_onepass.err; echo "bench rc=$?" >> $O/r2e_bench_onepass.err
tail -4 $O/r2e_ops.log; cat $O/r2e_kbench.log; tail -4 $O/r2e_parity.log; tail -2 $O/r2e_bench.err; tail -2 $O/r2e_bench_onepass.err
