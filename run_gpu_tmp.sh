cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "latent or gemm" 2>&1 | grep -v Warning | tail -n 8 > gpurun_out/r22_ops.log
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_attribution.py -q -m gpu -x 2>&1 | grep -v Warning | tail -n 8 > gpurun_out/r22_model.log
timeout 600 python tools/time_engine.py 8 > gpurun_out/r22_time_b8.log 2>&1
timeout 600 python tools/time_occlusion.py 32 > gpurun_out/r22_occ.log 2>&1
echo done
