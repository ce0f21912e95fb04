#!/bin/bash
# GPU session N: pair selection after the remote-arrive fix - VQ score GEMM in pairs or not; GEMM bench; step bench; attention + model tests
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 200 python tools/kernel_bench.py vq > $O/r2n_vq.log 2>&1
CTC_GEMM_ARGMAX_PAIR=1 timeout 200 python tools/kernel_bench.py vq >> $O/r2n_vq.log 2>&1
timeout 200 python tools/kernel_bench.py vq >> $O/r2n_vq.log 2>&1
CTC_GEMM_ARGMAX_PAIR=1 timeout 200 python tools/kernel_bench.py vq >> $O/r2n_vq.log 2>&1
cat $O/r2n_vq.log
timeout 300 python tools/kernel_bench.py gemm > $O/r2n_kbench.log 2>&1; cat $O/r2n_kbench.log
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_gemm_direct.py -m gpu -q -x > $O/r2n_model.log 2>&1; echo "rc=$?" >> $O/r2n_model.log
tail -3 $O/r2n_model.log
for i in 1 2; do
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-attribution > $O/r2n_bench_step$i.json 2> $O/r2n_bench_step$i.err; echo "bench rc=$?" >> $O/r2n_bench_step$i.err
tail -1 $O/r2n_bench_step$i.err; cut -c1-250 $O/r2n_bench_step$i.json
done
