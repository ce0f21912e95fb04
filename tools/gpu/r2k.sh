#!/bin/bash
# GPU session K: direct GEMM epilogues - parity with the staged ones, micro-benchmarks, model tests, step bench
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_direct.py -m gpu -q -x > $O/r2k_direct.log 2>&1; echo "rc=$?" >> $O/r2k_direct.log
tail -15 $O/r2k_direct.log
timeout 300 python tools/kernel_bench.py gemm > $O/r2k_kbench.log 2>&1; cat $O/r2k_kbench.log
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q -x > $O/r2k_model.log 2>&1; echo "rc=$?" >> $O/r2k_model.log
tail -5 $O/r2k_model.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-attribution > $O/r2k_bench_step.json 2> $O/r2k_bench_step.err; echo "bench rc=$?" >> $O/r2k_bench_step.err
tail -2 $O/r2k_bench_step.err; cut -c1-250 $O/r2k_bench_step.json
CTC_GEMM_DIRECT=0 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-attribution > $O/r2k_bench_staged.json 2> $O/r2k_bench_staged.err; echo "bench rc=$?" >> $O/r2k_bench_staged.err
tail -2 $O/r2k_bench_staged.err; cut -c1-250 $O/r2k_bench_staged.json
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-attribution > $O/r2k_bench_step2.json 2> $O/r2k_bench_step2.err
cut -c1-250 $O/r2k_bench_step2.json
