cd $GRAFT_REPO_ROOT
timeout 200 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "patchify" 2>&1 | grep -v Warning | tail -n 6 > gpurun_out/r42_ops.log
timeout 300 python -m pytest tests/test_gpu_model.py tests/test_gpu_attribution.py -q -m gpu -x 2>&1 | grep -v Warning | tail -n 4 > gpurun_out/r42_model.log
timeout 200 python tools/time_engine.py 8 > gpurun_out/r42_time_b8.log 2>&1
echo done
