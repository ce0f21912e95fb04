#!/bin/bash
# GPU session D (2 GPUs): NCCL equality check of the sharded attribution + the 2-GPU bench line (with its parity field)
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s > $O/r2d_multi.log 2>&1; echo "rc=$?" >> $O/r2d_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 5 --warmup 3 > $O/r2d_bench_2gpu.json 2> $O/r2d_bench_2gpu.err; echo "bench rc=$?" >> $O/r2d_bench_2gpu.err
tail -5 $O/r2d_multi.log; tail -3 $O/r2d_bench_2gpu.err; cat $O/r2d_bench_2gpu.json | cut -c1-600
