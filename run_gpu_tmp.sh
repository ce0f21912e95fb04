cd $GRAFT_REPO_ROOT
timeout 200 python -m pytest tests/test_preprocess.py -q -m gpu -x -s 2>&1 | grep -v Warning | tail -n 14 > gpurun_out/r26_pre.log
timeout 300 python -m pytest tests/test_gpu_attribution.py -q -m gpu -x 2>&1 | grep -v Warning | tail -n 8 > gpurun_out/r26_attr.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r26_bench.json 2> gpurun_out/r26_bench.err
echo done
