cd $GRAFT_REPO_ROOT
timeout 120 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "tcgen05" 2>&1 | grep -v Warning | tail -n 25 > gpurun_out/r30_ops.log
timeout 120 python tools/kernel_bench.py attn_s > gpurun_out/r30_attn_s.log 2>&1
echo done
