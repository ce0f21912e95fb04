#!/bin/bash
# GPU session B: kernel tests of the round's new kernels, micro-benchmarks, parity suite (both checkpoints), bench
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "gemm or layernorm or attention" > $O/r2b_ops.log 2>&1; echo "rc=$?" >> $O/r2b_ops.log
timeout 300 python tools/kernel_bench.py all > $O/r2b_kbench.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -s > $O/r2b_parity.log 2>&1; echo "parity rc=$?" >> $O/r2b_parity.log
timeout 700 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r2b_bench.json 2> $O/r2b_bench.err; echo "bench rc=$?" >> $O/r2b_bench.err
tail -4 $O/r2b_ops.log; cat $O/r2b_kbench.log; tail -4 $O/r2b_parity.log; tail -2 $O/r2b_bench.err
