#!/bin/bash
# GPU session I: GEMM micro-benchmarks + GEMM parity tests after an epilogue change
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_round2.py -m gpu -q -x -k "gemm or geglu" > $O/r2i_ops.log 2>&1; echo "rc=$?" >> $O/r2i_ops.log
timeout 300 python tools/kernel_bench.py gemm > $O/r2i_kbench.log 2>&1
tail -4 $O/r2i_ops.log; cat $O/r2i_kbench.log
