#!/bin/bash
# GPU session W (4 GPUs): bench line of the final code with its parity field
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 4 --steps 5 --warmup 3 > $O/r2w_bench_4gpu.json 2> $O/r2w_bench_4gpu.err; echo "bench rc=$?" >> $O/r2w_bench_4gpu.err
tail -2 $O/r2w_bench_4gpu.err; cut -c1-300 $O/r2w_bench_4gpu.json
