cd $GRAFT_REPO_ROOT
timeout 150 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "attention" 2>&1 | grep -v Warning | tail -n 12 > gpurun_out/r24_ops.log
timeout 300 python -m pytest tests/test_gpu_model.py tests/test_gpu_attribution.py -q -m gpu -x 2>&1 | grep -v Warning | tail -n 8 > gpurun_out/r24_model.log
timeout 200 python tools/time_engine.py 8 > gpurun_out/r24_time_b8.log 2>&1
timeout 200 python tools/time_occlusion.py 32 > gpurun_out/r24_occ.log 2>&1
echo done
