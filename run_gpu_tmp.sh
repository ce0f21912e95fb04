cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu 2>&1 | grep -v Warning | tail -n 40 > gpurun_out/r4_ops.log
timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -s 2>&1 | grep -v Warning | tail -n 60 > gpurun_out/r4_model.log
timeout 1200 python -m pytest tests/test_gpu_attribution.py -q -m gpu -s 2>&1 | grep -v Warning | tail -n 120 > gpurun_out/r4_attr.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r4_bench.json 2> gpurun_out/r4_bench.err
timeout 600 python tools/time_engine.py 8 > gpurun_out/r4_time_b8.log 2>&1
echo done
