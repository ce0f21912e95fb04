cd $GRAFT_REPO_ROOT
timeout 200 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "gemm" 2>&1 | grep -v Warning | tail -n 4 > gpurun_out/r27_ops.log
timeout 200 python tools/kernel_bench.py gemm > gpurun_out/r27_gemm.log 2>&1
timeout 200 python tools/time_engine.py 8 > gpurun_out/r27_time_b8.log 2>&1
timeout 200 python tools/time_occlusion.py 32 > gpurun_out/r27_occ.log 2>&1
echo done
