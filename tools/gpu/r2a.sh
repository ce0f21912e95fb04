#!/bin/bash
# GPU session A: pair-GEMM correctness + micro-benchmark, the full GPU suite, the parity suite, the bench line.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2a_smi.log 2>&1
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "gemm" > $O/r2a_gemm_tests.log 2>&1; echo "rc=$?" >> $O/r2a_gemm_tests.log
timeout 300 python tools/kernel_bench.py gemm > $O/r2a_gemm_bench.log 2>&1
if grep -q "rc=0" $O/r2a_gemm_tests.log; then export CTC_GEMM_PAIR=1; else export CTC_GEMM_PAIR=0; fi
echo "PAIR=$CTC_GEMM_PAIR" > $O/r2a_mode.log
timeout 1500 python -m pytest tests -m gpu -q -s --ignore=tests/test_gpu_parity.py > $O/r2a_tests.log 2>&1; echo "tests rc=$?" >> $O/r2a_tests.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s > $O/r2a_parity.log 2>&1; echo "parity rc=$?" >> $O/r2a_parity.log
timeout 700 python bench.py --steps 10 --warmup 3 > $O/r2a_bench.json 2> $O/r2a_bench.err; echo "bench rc=$?" >> $O/r2a_bench.err
tail -3 $O/r2a_gemm_tests.log; cat $O/r2a_gemm_bench.log; tail -5 $O/r2a_tests.log; tail -3 $O/r2a_parity.log; tail -3 $O/r2a_bench.err
